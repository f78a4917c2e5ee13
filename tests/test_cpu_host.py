"""CPU-only checks of the host side: C-ABI exports, state_dict layout, tokenizer ids, sharding logic
(gloo, world_size 2), and that the product path fails loudly without a GPU (no CPU fallback)."""
import inspect
import os
import re
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

import superconductor_vae_b200 as S
from superconductor_vae_b200 import _lib, parallel
from oracle import vocab as OV
from oracle import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scvae_b200.h")).read()
    declared = set(re.findall(r"\b(scv_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/scvae_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.scv_abi_version() == 1


@pytest.mark.parametrize("shape", [W.C512, W.C512B, W.C576, W.TINY, W.TINY_SKIP])
def test_decoder_state_dict_layout_matches_reference_names(shape):
    sd = W.make_decoder_state_dict(shape)          # keys verified against the reference with strict=True
    m = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shape.nhead, device="cpu")
    mine = m.state_dict()
    assert set(mine) == set(sd)
    for k in sd:
        assert mine[k].shape == sd[k].shape, k
        assert torch.equal(mine[k], sd[k]), k


def test_encoder_state_dict_layout_matches_reference_names():
    sd = W.make_encoder_state_dict()
    m = S.FullMaterialsVAE.from_state_dict(sd, device="cpu")
    mine = m.state_dict()
    assert set(mine) == set(sd)
    for k in sd:
        assert torch.equal(mine[k], sd[k]), k


def test_compiled_checkpoint_prefix_is_stripped():
    sd = W.make_decoder_state_dict(W.TINY)
    sd2 = {k.replace("transformer_decoder.", "transformer_decoder._orig_mod."): v for k, v in sd.items()}
    m = S.EnhancedTransformerDecoder.from_state_dict(sd2, nhead=W.TINY.nhead, device="cpu")
    assert torch.equal(m.state_dict()["transformer_decoder.layers.1.linear2.weight"],
                       sd["transformer_decoder.layers.1.linear2.weight"])


def test_no_cpu_fallback():
    sd = W.make_decoder_state_dict(W.TINY)
    m = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=W.TINY.nhead, device="cpu")
    with pytest.raises(S.EngineError):
        m.generate_with_kv_cache(W.make_latents(2, W.TINY.latent_dim), temperature=0.001)
    with pytest.raises(S.EngineError):                       # teacher-forced forward: engine only, no CPU path either
        m(torch.zeros(1, W.TINY.latent_dim), torch.zeros(1, 4, dtype=torch.long))
    with pytest.raises(S.EngineError):                       # ... nor has the scheduled-sampling (two-pass) forward
        m(torch.zeros(1, W.TINY.latent_dim), torch.zeros(1, 4, dtype=torch.long), teacher_forcing_ratio=0.5)
    names = list(inspect.signature(S.EnhancedTransformerDecoder.forward).parameters)[1:8]
    assert names == ["z", "target_tokens", "encoder_skip", "teacher_forcing_ratio", "stoich_pred", "cached_memory", "heads_pred"]


def test_positional_signature_order():
    import inspect
    names = list(inspect.signature(S.EnhancedTransformerDecoder.generate_with_kv_cache).parameters)[1:16]
    assert names == ["z", "encoder_skip", "stoich_pred", "temperature", "top_k", "top_p", "max_len",
                     "return_log_probs", "return_entropy", "cached_memory", "stop_boost", "hard_stop_threshold",
                     "heads_pred", "type_masks", "site_dup_threshold"]
    names = list(inspect.signature(S.EnhancedTransformerDecoder.sample_for_reinforce).parameters)[1:12]
    assert names == ["z", "encoder_skip", "stoich_pred", "temperature", "max_len", "cached_memory", "stop_boost",
                     "hard_stop_threshold", "heads_pred", "type_masks", "site_dup_threshold"]


# ----------------------------------------------------------------------------- tokenizer
def _tok(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tokenizer.pt"), weights_only=False)
    # id layout only depends on the list lengths; synthesise names where the golden did not record them
    fr = [f"{i + 1}/{100003}" for i in range(g["n_fractions"])]
    fr[0], fr[-1] = "1/2", "99873/100000"
    iso = [f"{300 + i}Og" for i in range(g["n_isotopes"])]
    iso[0], iso[-1] = "1H", "238U"
    return g, S.FractionAwareTokenizer(max_len=64, fractions=fr, isotopes=iso)


def test_tokenizer_layout_and_masks(golden_dir):
    g, tok = _tok(golden_dir)
    assert tok.vocab_size == g["vocab_size"] == 4752
    m = tok.get_type_masks()
    assert m.dtype == torch.bool and tuple(m.shape) == (5, 4752)
    assert m.sum(dim=1).tolist() == g["mask_row_sums"].tolist()
    assert torch.equal(m, OV.type_masks(g["n_fractions"], g["n_isotopes"]))
    for i, name in g["names"].items():
        if i in (143, 4459, 4461, 4751) or i < 143 or i == 4460:
            assert tok.get_token_name(i) == name, (i, tok.get_token_name(i), name)
    assert tok.encode("YBa2Cu3O7")[:12] == g["encoded_YBCO"][:12]
    assert [tok.decode(ids) for ids in g["ids"]] == g["decoded"]
    assert tok.decode_batch(torch.tensor(g["ids"][0]).unsqueeze(0)) == [g["decoded"][0]]
    assert tok.compute_token_type_targets(torch.tensor([2, 5, 123, 143, 4460, 0])).tolist() == [4, 0, 1, 2, 3, 3]


def test_tokenizer_fraction_canonicalisation():
    tok = S.FractionAwareTokenizer(max_len=16, fractions=["1/2", "3/10"], isotopes=["18O"])
    ids = tok.encode("La(2/4)Sr(3/10){18}O{17}O(7/9)", pad=False)
    assert ids == [1, tok._token_to_id["La"], 143, tok._token_to_id["Sr"], 144, tok.isotope_token_start,
                   tok.iso_unk_idx, 4, 2]
    assert tok.decode(ids) == "La(1/2)Sr(3/10){18}O{?}?(?/?)"


# ----------------------------------------------------------------------------- sharding + gather (gloo)
def test_shard_bounds_cover_rows():
    for n in (1, 7, 64, 1000003):
        for ws in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_rloo_order_roundtrip():
    B, k, ws = 10, 4, 3
    full = torch.arange(B * k).unsqueeze(1)           # row id in repeat layout (i*B + b)
    gathered = torch.cat([full[parallel.rloo_shard_rows(B, k, ws, r)] for r in range(ws)])
    assert torch.equal(parallel.restore_rloo_order(gathered, B, k, ws), full)


def _plain(objs):
    """Tensors as numpy arrays: torch.multiprocessing queues pass tensors by shared-memory handle, which the receiver can only
    open while the sending process is alive - a worker that exits right after q.put() makes the test flaky."""
    return tuple(o.numpy() if isinstance(o, torch.Tensor) else o for o in objs)


def _tensors(objs):
    import numpy as np
    return tuple(torch.from_numpy(o) if isinstance(o, np.ndarray) else o for o in objs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, ws, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    n = 11
    lo, hi = parallel.shard_bounds(n, ws, rank)
    L = 3 + 2 * rank                                   # ragged L per shard (each stops on its own rows)
    local = (torch.arange(lo, hi).unsqueeze(1) * 100 + torch.arange(L).unsqueeze(0)).to(torch.int32)
    out = parallel.gather_rows(local, n, pad_value=0)
    lp = parallel.gather_rows(local.float() * 0.5, n, pad_value=0.0)
    if rank == 0:
        q.put(_plain((out, lp)))
    dist.destroy_process_group()


def test_gather_rows_world_size_2():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out, lp = _tensors(q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tuple(out.shape) == (11, 5)
    lo1, _ = parallel.shard_bounds(11, 2, 1)
    assert out[:lo1, :3].tolist() == [[r * 100 + c for c in range(3)] for r in range(lo1)]
    assert (out[:lo1, 3:] == 0).all()                  # shorter shard padded to the global L
    assert out[lo1:].tolist() == [[r * 100 + c for c in range(5)] for r in range(lo1, 11)]
    assert torch.equal(lp, out.float() * 0.5)


class _FakeDecoder:
    """Host-logic stand-in: token (r, c) = hash of the row's own conditioning, so a wrongly sliced input shows."""
    vocab_size = 4752

    def generate_with_kv_cache(self, z, stoich_pred=None, heads_pred=None, encoder_skip=None, cached_memory=None,
                               _forced_tokens=None, return_log_probs=False, **kw):
        ref = z if z is not None else cached_memory.reshape(cached_memory.shape[0], -1)
        key = ref.sum(dim=1)
        for t in (stoich_pred, encoder_skip, _forced_tokens) + tuple((heads_pred or {}).values()):
            if t is not None:
                assert t.shape[0] == ref.shape[0], "per-row input not cut to the shard"
                key = key + t.reshape(t.shape[0], -1).float().sum(dim=1)
        L = 4
        toks = (key.unsqueeze(1) * 7 + torch.arange(L).unsqueeze(0)).to(torch.int64) % 4000 + 3
        return toks, (toks.float() * 0.25 if return_log_probs else None), None

    def sample_for_reinforce(self, z, stoich_pred=None, heads_pred=None, _n_samples=1, **kw):
        k = _n_samples                                     # the engine expands the base rows itself (repeat layout)
        rep = lambda t: None if t is None else t.repeat(k, *([1] * (t.dim() - 1)))
        t, lp, _ = self.generate_with_kv_cache(rep(z), stoich_pred=rep(stoich_pred),
                                               heads_pred={n: rep(v) for n, v in heads_pred.items()} if heads_pred else None,
                                               return_log_probs=True, **kw)
        return t, lp, lp * 2, torch.ones_like(lp)


def _sharded_worker(rank, ws, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=ws)
    g = torch.Generator().manual_seed(1)
    n = 13
    z = torch.randint(0, 50, (n, 6), generator=g).float()
    st = torch.randint(0, 50, (n, 3), generator=g).float()
    hp = {"tc_pred": torch.randint(0, 50, (n,), generator=g).float()}
    skip = torch.randint(0, 50, (n, 2), generator=g).float()
    dec = _FakeDecoder()
    full, full_lp, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, encoder_skip=skip, return_log_probs=True)
    t, lp, _ = parallel.generate_sharded(dec, z, stoich_pred=st, heads_pred=hp, encoder_skip=skip, return_log_probs=True)
    mem = z.reshape(n, 2, 3)
    fm, _, _ = dec.generate_with_kv_cache(None, cached_memory=mem)
    tm, _, _ = parallel.generate_sharded(dec, None, cached_memory=mem)
    B, k = 5, 3
    tr, lpr, enr, mkr = parallel.sample_for_reinforce_sharded(dec, z[:B], k, stoich_pred=st[:B], heads_pred={"tc_pred": hp["tc_pred"][:B]})
    fr, flr, _, _ = dec.sample_for_reinforce(z[:B].repeat(k, 1), stoich_pred=st[:B].repeat(k, 1),
                                             heads_pred={"tc_pred": hp["tc_pred"][:B].repeat(k)})
    bad = None
    try:
        parallel.generate_sharded(dec, z, stoich_pred=st[:5])
    except RuntimeError as e:
        bad = str(e)
    # chunked asynchronous gather (config 4): every rank walks its slice in chunks of 3 rows
    lo, hi = parallel.shard_bounds(n, ws, rank)
    cg = parallel.ChunkedGather(pad_to=4, dtype=torch.int16)
    bounds = [parallel.shard_bounds(n, ws, r) for r in range(ws)]
    n_chunks = max((h - l + 2) // 3 for l, h in bounds)
    for c in range(n_chunks):
        counts = [max(0, min(3, (h - l) - 3 * c)) for l, h in bounds]
        rows = full[lo + 3 * c: lo + 3 * c + counts[rank]]
        cg.push(rows, counts)
    chunked = cg.finish()
    if rank == 0:
        q.put(_plain((t, lp, full, full_lp, tm, fm, tr, fr, lpr, flr, bad, chunked)))
    dist.destroy_process_group()


def test_generate_sharded_slices_every_per_row_input_world_size_2():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    t, lp, full, full_lp, tm, fm, tr, fr, lpr, flr, bad, chunked = _tensors(q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t.dtype == torch.int16 and torch.equal(t.to(torch.int64), full) and torch.equal(lp, full_lp)
    assert torch.equal(tm.to(torch.int64), fm)
    assert tr.dtype == torch.int64 and torch.equal(tr, fr) and torch.equal(lpr, flr)     # global sample-major order
    assert bad is not None and "stoich_pred" in bad
    assert torch.equal(chunked.to(torch.int64), full)


def test_legacy_vocab_of_generate_formulas_fast(golden_dir):
    from superconductor_vae_b200.decoder import _LEGACY_VOCAB
    g = torch.load(os.path.join(golden_dir, "tokenizer.pt"), weights_only=False)
    assert _LEGACY_VOCAB == g["legacy_vocab"] and len(_LEGACY_VOCAB) == 148
    ids = [1, 58, 141, 48, 140, 2, 5]
    assert "".join(_LEGACY_VOCAB[i] for i in ids[1:5]) == g["legacy_decoded"]


def test_bench_reference_arm_prints_one_contract_line():
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-rows", "2", "--max-len", "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    from oracle import ref_loader
    # the unmodified reference module when oracle/_ref has been built (oracle/make_ref.py), else the golden-pinned port
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0
