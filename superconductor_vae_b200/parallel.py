"""Batch-dimension sharding across the GPUs of one box (one process per GPU).

Sequences are independent: each rank decodes a contiguous slice of the rows with
replicated weights and its own KV pages; the only collective is one gather of token ids (+ log-probs /
entropy for RL) at the end.  Works with any torch.distributed backend ("nccl" on GPUs, "gloo" in the
CPU tests of the host logic).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n_rows % world_size ranks get one extra row."""
    base, rem = divmod(n_rows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rloo_shard_rows(base_batch: int, k: int, world_size: int, rank: int) -> torch.Tensor:
    """Row ids (in the reference's sample-major `repeat` layout, row = i*B + b,
    scripts/train_v12_clean.py:2677-2688) owned by `rank` when the *base* batch is sharded."""
    lo, hi = shard_bounds(base_batch, world_size, rank)
    b = torch.arange(lo, hi)
    return (torch.arange(k).unsqueeze(1) * base_batch + b.unsqueeze(0)).reshape(-1)


def gather_rows(local: torch.Tensor, n_rows_total: int, pad_value=0, group=None) -> torch.Tensor:
    """All-gather [n_local, L_local] shards into [n_rows_total, L_max].

    L differs per shard (each stops when *its* rows have all finished): shards are padded to the
    global max before the exchange.  Rows must have been split with `shard_bounds`."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    meta = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=local.device)
    metas = [torch.zeros_like(meta) for _ in range(ws)]
    dist.all_gather(metas, meta, group=group)
    l_max = max(int(m[1]) for m in metas)
    n_max = max(int(m[0]) for m in metas)
    buf = torch.full((n_max, l_max), pad_value, dtype=local.dtype, device=local.device)
    buf[:local.shape[0], :local.shape[1]] = local
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf, group=group)
    out = torch.cat([p[:int(m[0])] for p, m in zip(parts, metas)], dim=0)
    assert out.shape[0] == n_rows_total, (out.shape, n_rows_total)
    return out


def restore_rloo_order(gathered: torch.Tensor, base_batch: int, k: int, world_size: int) -> torch.Tensor:
    """Undo the base-batch sharding so that rows are again sample-major over the full base batch."""
    order = torch.cat([rloo_shard_rows(base_batch, k, world_size, r) for r in range(world_size)])
    out = torch.empty_like(gathered)
    out[order.to(gathered.device)] = gathered
    return out


def generate_sharded(decoder, z: torch.Tensor, *, stoich_pred=None, heads_pred=None, group=None, token_dtype=torch.int32,
                     **gen_kwargs):
    """Every rank holds the same [N, ...] inputs (or at least its slice): decode the local slice, gather tokens.

    Returns (tokens [N, L_max] on every rank, log_probs or None, entropy or None)."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    n = z.shape[0]
    lo, hi = shard_bounds(n, ws, rank)
    sl = slice(lo, hi)
    hp = {k: v[sl] for k, v in heads_pred.items()} if heads_pred is not None else None
    toks, lps, ent = decoder.generate_with_kv_cache(z[sl], stoich_pred=stoich_pred[sl] if stoich_pred is not None else None,
                                                    heads_pred=hp, **gen_kwargs)
    out_t = gather_rows(toks.to(token_dtype), n, 0, group)
    out_l = gather_rows(lps, n, 0.0, group) if lps is not None else None
    out_e = gather_rows(ent, n, 0.0, group) if ent is not None else None
    return out_t, out_l, out_e
