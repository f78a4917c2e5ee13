"""Batch-dimension sharding across the GPUs of one box (one process per GPU).

Sequences are independent: each rank decodes a contiguous slice of the rows with
replicated weights and its own KV pages; the only collective is one gather of token ids (+ log-probs /
entropy for RL) at the end.  Works with any torch.distributed backend ("nccl" on GPUs, "gloo" in the
CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

# keyword arguments of generate_with_kv_cache that carry one row per sequence: every one of them is cut to the
# rank's rows (a full-size tensor handed to a shard would silently condition rows lo..hi on rows 0..n_local)
PER_ROW_KWARGS = ("encoder_skip", "cached_memory", "_forced_tokens")


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


def _shard_meta(local: torch.Tensor, group) -> List[Tuple[int, int]]:
    ws = dist.get_world_size(group)
    meta = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=local.device)
    metas = [torch.zeros_like(meta) for _ in range(ws)]
    dist.all_gather(metas, meta, group=group)
    return [(int(m[0]), int(m[1])) for m in metas]


def _wire(t: torch.Tensor) -> torch.Tensor:
    """int16 has no NCCL / gloo wire type: such rows travel as their bytes."""
    return t.contiguous().view(torch.uint8) if t.dtype == torch.int16 else t


def _unwire(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return t.view(torch.int16) if dtype == torch.int16 else t


def gather_rows(local: torch.Tensor, n_rows_total: int, pad_value=0, group=None) -> torch.Tensor:
    """All-gather [n_local, L_local] shards into [n_rows_total, L_max].

    L differs per shard (each stops when *its* rows have all finished): shards are padded to the
    global max before the exchange.  Rows must have been split with `shard_bounds`."""
    ws = dist.get_world_size(group)
    metas = _shard_meta(local, group)
    l_max = max(m[1] for m in metas)
    n_max = max(m[0] for m in metas)
    buf = torch.full((n_max, l_max), pad_value, dtype=local.dtype, device=local.device)
    buf[:local.shape[0], :local.shape[1]] = local
    wire = _wire(buf)
    parts = [torch.empty_like(wire) for _ in range(ws)]
    dist.all_gather(parts, wire, group=group)
    out = torch.cat([_unwire(p, local.dtype)[:m[0]] for p, m in zip(parts, metas)], dim=0)
    assert out.shape[0] == n_rows_total, (out.shape, n_rows_total)
    return out


def restore_rloo_order(gathered: torch.Tensor, base_batch: int, k: int, world_size: int) -> torch.Tensor:
    """Undo the base-batch sharding so that rows are again sample-major over the full base batch."""
    order = torch.cat([rloo_shard_rows(base_batch, k, world_size, r) for r in range(world_size)])
    out = torch.empty_like(gathered)
    out[order.to(gathered.device)] = gathered
    return out


def _slice_rows(kwargs: Dict, sl, n: int) -> Dict:
    out = dict(kwargs)
    for name in PER_ROW_KWARGS:
        t = out.get(name)
        if t is not None:
            if t.shape[0] != n:
                raise RuntimeError(f"{name} has {t.shape[0]} rows, z has {n}")
            out[name] = t[sl]
    return out


def token_dtype_for(vocab_size: int) -> torch.dtype:
    """Token ids travel as int16 when the vocabulary allows (V = 4752 here): half the gather bytes of int32."""
    return torch.int16 if vocab_size <= 32767 else torch.int32


def generate_sharded(decoder, z: Optional[torch.Tensor], *, stoich_pred=None, heads_pred=None, group=None,
                     token_dtype: Optional[torch.dtype] = None, **gen_kwargs):
    """Every rank holds the same [N, ...] inputs: decode the rank's slice, all-gather tokens (+ log-probs / entropy).

    All per-row inputs are cut to the rank's rows: z, stoich_pred, every heads_pred entry, and the per-row keyword
    arguments encoder_skip, cached_memory, _forced_tokens.  Returns (tokens [N, L_max] on every rank in
    `token_dtype` (default: int16 when the vocabulary fits), log_probs or None, entropy or None).  Differences from
    one unsharded call are the documented per-shard ones: each shard stops when ITS rows have
    finished (positions past a shard's L are PAD / 0.0), and the H2 degenerate flag is per shard."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    ref = z if z is not None else gen_kwargs.get("cached_memory")
    if ref is None:
        raise RuntimeError("generate_sharded needs z or cached_memory")
    n = ref.shape[0]
    lo, hi = shard_bounds(n, ws, rank)
    sl = slice(lo, hi)
    for name, t in (("stoich_pred", stoich_pred),) + tuple((heads_pred or {}).items()):
        if t is not None and t.shape[0] != n:
            raise RuntimeError(f"{name} has {t.shape[0]} rows, z has {n}")
    hp = {k: v[sl] for k, v in heads_pred.items()} if heads_pred is not None else None
    kw = _slice_rows(gen_kwargs, sl, n)
    toks, lps, ent = decoder.generate_with_kv_cache(z[sl] if z is not None else None,
                                                    stoich_pred=stoich_pred[sl] if stoich_pred is not None else None,
                                                    heads_pred=hp, **kw)
    if token_dtype is None:
        token_dtype = token_dtype_for(getattr(decoder, "vocab_size", 1 << 30))
    out_t = gather_rows(toks.to(token_dtype), n, 0, group)
    out_l = gather_rows(lps, n, 0.0, group) if lps is not None else None
    out_e = gather_rows(ent, n, 0.0, group) if ent is not None else None
    return out_t, out_l, out_e


def sample_for_reinforce_sharded(decoder, z: torch.Tensor, k: int, *, stoich_pred=None, heads_pred=None, group=None,
                                 **gen_kwargs):
    """RLOO rollouts over the ranks: `z`, `stoich_pred`, `heads_pred` are the BASE batch [B, ...] (not yet repeated).

    Each rank expands its slice of the base batch k times locally (sample-major, like the reference's `repeat`,
    scripts/train_v12_clean.py:2677-2688), samples, and the gathered rows are put back into the reference's global
    sample-major order (row i*B + b), so `rewards.view(k, B)` (:2778) works unchanged.
    Returns (tokens int64 [k*B, L], log_probs, entropy, mask) like `sample_for_reinforce`."""
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    B = z.shape[0]
    lo, hi = shard_bounds(B, ws, rank)
    sl = slice(lo, hi)
    cut = lambda t: t[sl]
    hp = {n_: cut(v) for n_, v in heads_pred.items()} if heads_pred is not None else None
    kw = dict(gen_kwargs)
    forced = kw.pop("_forced_tokens", None)
    for name in ("encoder_skip", "cached_memory"):
        if kw.get(name) is not None:
            kw[name] = cut(kw[name])
    if forced is not None:                                  # forced tokens are per SAMPLE row: expand like the outputs
        kw["_forced_tokens"] = forced[rloo_shard_rows(B, k, ws, rank).to(forced.device)]
    # the k samples of a latent share its memory tokens inside the engine (_n_samples): nothing is repeated here
    toks, lp, ent, _ = decoder.sample_for_reinforce(cut(z), stoich_pred=cut(stoich_pred) if stoich_pred is not None else None,
                                                    heads_pred=hp, _n_samples=k, **kw)
    n = B * k
    tdt = token_dtype_for(getattr(decoder, "vocab_size", 1 << 30))
    out_t = restore_rloo_order(gather_rows(toks.to(tdt), n, 0, group), B, k, ws).to(torch.int64)
    out_l = restore_rloo_order(gather_rows(lp, n, 0.0, group), B, k, ws)
    out_e = restore_rloo_order(gather_rows(ent, n, 0.0, group), B, k, ws)
    Lq = out_t.shape[1]
    is_end = out_t == 2                                                    # END_IDX; mask as in the reference (:1620-1639)
    end_pos = torch.where(is_end.any(dim=1), torch.argmax(is_end.int(), dim=1), torch.full((n,), Lq, device=out_t.device))
    mask = (torch.arange(Lq, device=out_t.device).unsqueeze(0) <= end_pos.unsqueeze(1)).float()
    return out_t, out_l, out_e, mask


class ChunkedGather:
    """Token ids of a long candidate stream (BASELINE config 4: 1 M latents), gathered chunk by chunk.

    Every rank owns one contiguous slice of the stream (`shard_bounds`) and walks it in chunks.  `push(tokens, counts)`
    hands over the rank's rows of its current chunk; the all-gather is issued asynchronously (NCCL runs it on its own
    stream), so it overlaps the decode of the next chunk.  `finish()` waits for every gather and returns
    [n_total, pad_to] in global row order (rank 0's slice, then rank 1's, ...)."""

    def __init__(self, pad_to: int, dtype: torch.dtype = torch.int16, group=None):
        self.pad_to, self.dtype, self.group = pad_to, dtype, group
        self.ws = dist.get_world_size(group)
        self.pending: List[Tuple[object, List[torch.Tensor], List[int]]] = []

    def push(self, tokens: torch.Tensor, rows_per_rank: Sequence[int]) -> None:
        """tokens [n_local, L <= pad_to]; rows_per_rank = how many rows every rank contributes to this chunk
        (0 for a rank whose slice is already exhausted; every rank still calls push so the collective matches)."""
        n_max = max(max(rows_per_rank), 1)
        buf = torch.zeros((n_max, self.pad_to), dtype=self.dtype, device=tokens.device)
        if tokens.numel() > 0:
            buf[:tokens.shape[0], :tokens.shape[1]] = tokens.to(self.dtype)
        wire = _wire(buf)
        parts = [torch.empty_like(wire) for _ in range(self.ws)]
        work = dist.all_gather(parts, wire, group=self.group, async_op=True)
        self.pending.append((work, parts, list(rows_per_rank)))

    def finish(self) -> torch.Tensor:
        for work, _, _ in self.pending:
            work.wait()
        rows = [_unwire(parts[r], self.dtype)[:counts[r]] for r in range(self.ws) for _, parts, counts in self.pending]
        self.pending = []
        return torch.cat(rows, dim=0)
