"""Parity of the CUDA engine (through the Python shim -> C ABI) against the CPU oracle and the golden
vectors recorded from the reference modules.  Run on the B200 box: pytest -m gpu.

Bars (BASELINE.json north_star): greedy token sequences bit-exact; logits / log-probs / entropy within the
fp32 tolerances written next to each assert; sampling matches the reference distribution statistically.
"""
import math
import os

import pytest
import torch

import superconductor_vae_b200 as S
from superconductor_vae_b200 import _lib
from oracle import decoder_oracle as DO
from oracle import encoder_oracle as EO
from oracle import latent as OL
from oracle import vocab as OV
from oracle import weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
torch.set_num_threads(max(1, os.cpu_count() or 1))

SHAPES = {"tiny": (W.TINY, OV.TINY_LAYOUT), "c512_b32": (W.C512, {}), "c512b_b4": (W.C512B, {}),
          "c576_b4": (W.C576, {})}
_cache = {}


def _golden(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _setup(golden_dir, name):
    if name in _cache:
        return _cache[name]
    shape, layout = SHAPES[name]
    g = _golden(golden_dir, name)
    B, seed = g["meta"]["B"], g["meta"]["seed_in"]
    sd = W.make_decoder_state_dict(shape, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shape.nhead, device=DEV)
    z = W.make_latents(B, shape.latent_dim, seed)
    stoich, heads = W.make_conditioning(B, shape.stoich_input_dim, seed)
    masks = OV.type_masks(**layout)
    _cache[name] = (shape, g, sd, dec, z, stoich, heads, masks)
    return _cache[name]


def _cuda(x):
    if x is None:
        return None
    if isinstance(x, dict):
        return {k: v.to(DEV) for k, v in x.items()}
    return x.to(DEV)


# ------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("M,N,K", [(1, 5, 1), (5, 13, 13), (33, 97, 24), (130, 512, 145), (32, 1536, 512),
                                   (300, 256, 513), (64, 512, 2214), (257, 4752, 512), (1024, 2048, 512),
                                   (49, 512, 2048)])
@pytest.mark.parametrize("act", [0, 1])
def test_op_linear_matches_fp32(M, N, K, act):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x = torch.randn((M, K), generator=g)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(torch.bfloat16).float()
    b = torch.randn((N,), generator=g)
    r = torch.randn((M, N), generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    ref = (ref + r.double()).float()
    L = _lib.lib()
    ldw = (K + 7) // 8 * 8
    xd, wd, bd, rd = x.to(DEV), w.to(DEV).contiguous(), b.to(DEV), r.to(DEV)
    wp = torch.zeros((N, ldw), dtype=torch.bfloat16, device=DEV)
    _lib.check(L.scv_op_pack_bf16(_lib.ptr(wd), _lib.ptr(wp), N, K, ldw, _lib.current_stream()))
    y = torch.empty((M, N), device=DEV)
    _lib.check(L.scv_op_linear(_lib.ptr(xd), K, _lib.ptr(wp), ldw, _lib.ptr(bd), _lib.ptr(rd), N, _lib.ptr(y), N,
                               M, N, K, act, 1, _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(wp[:, :K].float().cpu(), w)
    torch.testing.assert_close(y.cpu(), ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 512, 512), (4096, 1536, 512), (4096, 512, 2048),
                                   (200, 132, 72), (64, 4752, 512), (1000, 2048, 576)])
@pytest.mark.parametrize("act,residual", [(0, True), (1, False)])
def test_op_linear_tcgen05_matches_fp64(M, N, K, act, residual):
    """Tensor-core path (bf16 weights, activations split into bf16 hi + lo, fp32 accumulate in TMEM):
    |diff| <= 5e-5 abs against fp64 math on O(1) outputs, i.e. fp32-accumulation noise only."""
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    x = torch.randn((M, K), generator=g)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(torch.bfloat16).float()
    b, r = torch.randn((N,), generator=g), torch.randn((M, N), generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    ref = torch.nn.functional.gelu(ref) if act == 1 else ref
    ref = ref + r.double() if residual else ref
    L = _lib.lib()
    xd, wd, bd, rd = x.to(DEV), w.to(DEV).contiguous(), b.to(DEV), r.to(DEV)
    wt = torch.zeros(int(L.scv_op_tiled_elems(N, K)), dtype=torch.bfloat16, device=DEV)
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(wd), _lib.ptr(wt), N, K, _lib.current_stream()))
    y = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(L.scv_op_linear(_lib.ptr(xd), K, _lib.ptr(wt), 0, _lib.ptr(bd), _lib.ptr(rd) if residual else None, N,
                               _lib.ptr(y), N, M, N, K, act, 2, _lib.current_stream()))
    torch.cuda.synchronize()
    assert float((y.cpu().double() - ref).abs().max()) <= 5e-5


@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (200, 132, 72), (64, 4752, 512), (384, 512, 2048)])
def test_op_linear_narrow_tile_is_bit_identical(M, N, K):
    """The 128 x 64 tile used for projections with few row tiles walks k in the same order as the 128 x 128 tile."""
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn((M, K), generator=g).to(DEV)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(DEV).contiguous()
    b, r = torch.randn((N,), generator=g).to(DEV), torch.randn((M, N), generator=g).to(DEV)
    L = _lib.lib()
    wt = torch.zeros(int(L.scv_op_tiled_elems(N, K)), dtype=torch.bfloat16, device=DEV)
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(w), _lib.ptr(wt), N, K, _lib.current_stream()))
    ys = []
    try:
        for bn64 in (0, 8):
            _lib.tune(gemm_bn64=bn64)
            for act, res in ((0, True), (1, False)):
                y = torch.full((M, N), float("nan"), device=DEV)
                _lib.check(L.scv_op_linear(_lib.ptr(x), K, _lib.ptr(wt), 0, _lib.ptr(b), _lib.ptr(r) if res else None, N,
                                           _lib.ptr(y), N, M, N, K, act, 2, _lib.current_stream()))
                torch.cuda.synchronize()
                ys.append(y)
    finally:
        _lib.tune(gemm_bn64=8)
    assert torch.equal(ys[0], ys[2]) and torch.equal(ys[1], ys[3]) and not bool(torch.isnan(ys[2]).any())


def test_op_layernorm_matches_fp32():
    g = torch.Generator().manual_seed(3)
    for M, N in ((1, 64), (37, 512), (130, 576), (9, 1024)):
        x = torch.randn((M, N), generator=g) * 3 + 1
        ga, be = torch.randn((N,), generator=g), torch.randn((N,), generator=g)
        ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x, (N,), ga, be, 1e-5))
        xd, gd, bd, y = x.to(DEV), ga.to(DEV), be.to(DEV), torch.empty((M, N), device=DEV)
        _lib.check(_lib.lib().scv_op_layernorm(_lib.ptr(xd), N, _lib.ptr(gd), _lib.ptr(bd),
                                               _lib.ptr(y), N, M, N, 1, _lib.current_stream()))
        torch.cuda.synchronize()
        torch.testing.assert_close(y.cpu(), ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------ memory builder
@pytest.mark.parametrize("name", ["tiny", "c512b_b4", "c576_b4", "c512_b32"])
def test_memory_matches_oracle(golden_dir, name):
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    mem = dec.precompute_memory(_cuda(z), None, _cuda(stoich), _cuda(heads)).cpu()
    ref = DO.build_memory(sd, z, None, stoich, heads)
    assert tuple(mem.shape) == tuple(ref.shape) and mem.shape[1] == int(g["memory_shapes"][0])
    torch.testing.assert_close(mem, ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(mem[:2], g["memory24_rows"], rtol=1e-4, atol=2e-5)
    assert dec.precompute_memory(_cuda(z), None, _cuda(stoich), None).shape[1] == int(g["memory_shapes"][1])
    assert dec.precompute_memory(_cuda(z), None, None, None).shape[1] == int(g["memory_shapes"][2])


def test_memory_with_skip_connection(golden_dir):
    g = _golden(golden_dir, "tiny_skip")
    shape = W.TINY_SKIP
    sd = W.make_decoder_state_dict(shape, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shape.nhead, device=DEV)
    B = g["meta"]["B"]
    z = W.make_latents(B, shape.latent_dim, 99)
    stoich, heads = W.make_conditioning(B, shape.stoich_input_dim, 99)
    mem = dec.precompute_memory(_cuda(z), _cuda(g["skip"]), _cuda(stoich), _cuda(heads))
    torch.testing.assert_close(mem.cpu(), g["memory"], rtol=1e-4, atol=2e-5)
    t, _, _ = dec.generate_with_kv_cache(_cuda(z), encoder_skip=_cuda(g["skip"]), stoich_pred=_cuda(stoich),
                                         temperature=0.001, heads_pred=_cuda(heads))
    assert torch.equal(t.cpu().to(torch.int16), g["tokens"])


def test_heads_batch_mismatch_raises(golden_dir):
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    bad = {k: v[:3] for k, v in heads.items()}
    with pytest.raises(RuntimeError):
        dec.precompute_memory(_cuda(z), None, _cuda(stoich), _cuda(bad))


# ------------------------------------------------------------------------------------------ greedy, bit exact
@pytest.mark.parametrize("name", ["tiny", "c512b_b4", "c576_b4", "c512_b32"])
def test_greedy_tokens_bit_exact_vs_reference_goldens(golden_dir, name):
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    zc, sc, hc, mc = _cuda(z), _cuda(stoich), _cuda(heads), _cuda(masks)
    # (a) type masks + stop head + hard stop (BASELINE config 1/2 settings)
    t, lp, en = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.001, max_len=shape.max_len,
                                           heads_pred=hc, type_masks=mc, stop_boost=10.0, hard_stop_threshold=0.8,
                                           return_log_probs=True, return_entropy=True)
    assert t.dtype == torch.int64 and t.is_cuda
    assert torch.equal(t.cpu().to(torch.int16), g["greedy_masked_tokens"]), "masked greedy tokens differ"
    assert float(lp.abs().max()) == 0.0                         # greedy log-probs are zeros (:1508-1509)
    # H2: any -inf in the batch -> entropy is ln V for every row
    torch.testing.assert_close(en[0].cpu(), g["greedy_masked_entropy_row0"], rtol=1e-5, atol=1e-5)
    # (b) plain greedy: all max_len-1 steps
    t, _, en = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.001, max_len=shape.max_len,
                                          heads_pred=hc, return_entropy=True)
    assert torch.equal(t.cpu().to(torch.int16), g["greedy_plain_tokens"]), "plain greedy tokens differ"
    n = g["greedy_plain_entropy"].shape[0]
    torch.testing.assert_close(en[:n].cpu(), g["greedy_plain_entropy"], rtol=2e-4, atol=2e-4)
    # (c) stop boost only, (d) 20 memory tokens, default max_len
    t, _, _ = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.001, max_len=shape.max_len,
                                         heads_pred=hc, stop_boost=10.0)
    assert torch.equal(t.cpu().to(torch.int16), g["greedy_stopboost_tokens"])
    t, _, _ = dec.generate_with_kv_cache(z=zc, stoich_pred=sc, temperature=0.001)
    assert torch.equal(t.cpu().to(torch.int16), g["greedy_m20_tokens"])
    # (e) H1: temperature 0.0 reproduces the reference's divide-by-zero behaviour
    t, _, _ = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.0, max_len=shape.max_len, heads_pred=hc,
                                         type_masks=mc, stop_boost=10.0, hard_stop_threshold=0.8)
    assert torch.equal(t.cpu().to(torch.int16), g["t0_masked_tokens"])
    t, _, _ = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.0, max_len=min(shape.max_len, 8),
                                         heads_pred=hc)
    assert torch.equal(t.cpu().to(torch.int16), g["t0_plain_tokens"])
    assert dec.training is False


@pytest.mark.parametrize("name", ["tiny", "c512b_b4"])
def test_site_dup_gating_bit_exact_vs_reference_goldens(golden_dir, name):
    """site_dup_threshold > 0 (reference :1424-1435, :1525-1539; SURVEY H5: inert by default, old-vocabulary id range)."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    zc, sc, hc, mc = _cuda(z), _cuda(stoich), _cuda(heads), _cuda(masks)
    t, _, _ = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.001, max_len=shape.max_len, heads_pred=hc,
                                         site_dup_threshold=0.6)
    assert torch.equal(t.cpu().to(torch.int16), g["sitedup_plain_tokens"])
    t, _, _ = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.001, max_len=shape.max_len, heads_pred=hc,
                                         site_dup_threshold=0.99, type_masks=mc, stop_boost=10.0)
    assert torch.equal(t.cpu().to(torch.int16), g["sitedup_masked_tokens"])
    # and the ungated call right after it is unaffected by the gate's state
    t, _, _ = dec.generate_with_kv_cache(zc, stoich_pred=sc, temperature=0.001, max_len=shape.max_len, heads_pred=hc)
    assert torch.equal(t.cpu().to(torch.int16), g["greedy_plain_tokens"])


@pytest.mark.parametrize("name", ["tiny", "c512_b32"])
def test_last_step_logits_within_tolerance(golden_dir, name):
    """Raw logits / hidden state of the final executed step vs the fp32 oracle: |diff| <= 2e-4 abs."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    steps = min(shape.max_len, 9)
    trace = {}
    ref_t, _, _ = DO.generate_with_kv_cache(sd, shape.nhead, z, stoich_pred=stoich, temperature=0.001, max_len=steps,
                                            heads_pred=heads, type_masks=masks, stop_boost=10.0, trace=trace,
                                            stop_when_all_finished=False)
    t, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), temperature=0.001, max_len=steps,
                                         heads_pred=_cuda(heads), type_masks=_cuda(masks), stop_boost=10.0)
    L = t.shape[1]
    assert torch.equal(t.cpu(), ref_t[:, :L])
    torch.testing.assert_close(dec.debug_tap(0).cpu(), trace["hidden"][L - 1], rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(dec.debug_tap(1).cpu(), trace["raw_logits"][L - 1], rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(dec.debug_tap(2).cpu(), trace["type_logits"][L - 1], rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(dec.debug_tap(3).cpu(), trace["stop_logit"][L - 1], rtol=1e-4, atol=2e-4)


def test_max_len_is_clamped_to_pe_buffer(golden_dir):
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    t, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), temperature=0.001, max_len=500,
                                         heads_pred=_cuda(heads))
    assert t.shape[1] == shape.max_len - 1
    assert torch.equal(t.cpu().to(torch.int16), g["greedy_plain_tokens"])


def test_wrong_shapes_raise_like_the_reference(golden_dir):
    """The kernels read raw pointers with the configured strides, so every tensor is checked against the batch and the
    feature dims first: what the reference rejects in F.linear / view / cat (RuntimeError) is rejected here too."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    zc, sc, hc = _cuda(z), _cuda(stoich), _cuda(heads)
    B = z.shape[0]
    with pytest.raises(RuntimeError):
        dec.generate_with_kv_cache(zc[:, :-1], stoich_pred=sc, temperature=0.001)                     # latent_dim
    with pytest.raises(RuntimeError):
        dec.generate_with_kv_cache(zc, stoich_pred=torch.zeros((B, 37), device=DEV), temperature=0.001)   # V12 stoich on a 13-dim model
    with pytest.raises(RuntimeError):
        dec.generate_with_kv_cache(zc, stoich_pred=sc[:-1], temperature=0.001)                        # batch
    bad = dict(hc)
    bad["tc_class_logits"] = hc["tc_class_logits"][:, :3]
    with pytest.raises(RuntimeError):
        dec.precompute_memory(zc, None, sc, bad)                                                      # heads_input width
    bad = dict(hc)
    bad["tc_pred"] = hc["tc_pred"][:-1]
    with pytest.raises(RuntimeError):
        dec.precompute_memory(zc, None, sc, bad)                                                      # stale head tensor (:838-843)
    mem = dec.precompute_memory(zc, None, sc, hc)
    with pytest.raises(RuntimeError):
        dec.generate_with_kv_cache(None, temperature=0.001, cached_memory=mem[:, :, :-1])             # d_model
    with pytest.raises(RuntimeError):
        dec.generate_with_kv_cache(None, temperature=0.001, cached_memory=mem, type_masks=_cuda(masks)[:, :-1])
    with pytest.raises(RuntimeError):
        dec.generate_with_kv_cache(None, temperature=0.001, cached_memory=mem, _forced_tokens=torch.zeros((B + 1, 3), dtype=torch.long))


def test_cached_memory_and_chunking_are_equivalent(golden_dir):
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    mem = dec.precompute_memory(_cuda(z), None, _cuda(stoich), _cuda(heads))
    t1, _, _ = dec.generate_with_kv_cache(None, temperature=0.001, cached_memory=mem, max_len=shape.max_len)
    assert torch.equal(t1.cpu().to(torch.int16), g["greedy_plain_tokens"])
    old = dec.max_rows_per_call
    try:
        dec.max_rows_per_call = 4          # 6 rows -> chunks of 4 + 2
        t2, _, _ = dec.generate_with_kv_cache(None, temperature=0.001, cached_memory=mem, max_len=shape.max_len)
    finally:
        dec.max_rows_per_call = old
    assert torch.equal(t1, t2)


# ------------------------------------------------------------------------------------------ sampling
@pytest.mark.parametrize("name", ["tiny", "c512b_b4"])
def test_sampled_logprobs_and_entropy_match_oracle_replay(golden_dir, name):
    """The engine samples with Philox, so tokens differ from torch.multinomial; replaying the engine's own
    tokens through the oracle must reproduce log-prob and entropy (tolerance 1e-3 abs)."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    t, lp, en, mk = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), temperature=1.2,
                                             max_len=shape.max_len, stop_boost=10.0, heads_pred=_cuda(heads), _seed=11)
    t, lp, en, mk = t.cpu(), lp.cpu(), en.cpu(), mk.cpu()
    rt, rlp, ren = DO.generate_with_kv_cache(sd, shape.nhead, z, stoich_pred=stoich, temperature=1.2,
                                             max_len=shape.max_len, stop_boost=10.0, heads_pred=heads,
                                             return_log_probs=True, return_entropy=True, forced_tokens=t)
    assert rt.shape == t.shape
    torch.testing.assert_close(lp, rlp, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(en, ren, rtol=1e-3, atol=1e-3)
    assert torch.equal(mk, DO.reinforce_mask(t))
    # same seed -> same sample; different seed -> different sample
    t2, _, _, _ = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), temperature=1.2,
                                           max_len=shape.max_len, stop_boost=10.0, heads_pred=_cuda(heads), _seed=11)
    assert torch.equal(t2.cpu(), t)
    t3, _, _, _ = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), temperature=1.2,
                                           max_len=shape.max_len, stop_boost=10.0, heads_pred=_cuda(heads), _seed=12)
    assert not torch.equal(t3.cpu()[:, :min(t3.shape[1], t.shape[1])], t[:, :min(t3.shape[1], t.shape[1])])


@pytest.mark.parametrize("name,kw", [("tiny", dict(top_k=7)), ("tiny", dict(top_p=0.8)), ("tiny", dict(top_k=20, top_p=0.6)),
                                     ("c512b_b4", dict(top_k=50)), ("c512b_b4", dict(top_p=0.9))])
def test_topk_topp_filtering_matches_oracle_replay(golden_dir, name, kw):
    """top-k / nucleus filtering (reference :1489-1503): every sampled token must lie inside the oracle's filtered set
    and carry the oracle's log-prob of the filtered, temperature-scaled distribution (1e-3 abs)."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    t, lp, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), temperature=0.9, heads_pred=_cuda(heads),
                                          max_len=min(shape.max_len, 16), return_log_probs=True, _seed=3, **kw)
    t, lp = t.cpu(), lp.cpu()
    rt, rlp, _ = DO.generate_with_kv_cache(sd, shape.nhead, z, stoich_pred=stoich, temperature=0.9, heads_pred=heads,
                                           max_len=min(shape.max_len, 16), return_log_probs=True, forced_tokens=t, **kw)
    assert rt.shape == t.shape
    assert float(rlp.min()) > math.log(1e-8) + 1e-3, "a sampled token fell outside the oracle's top-k / top-p set"
    torch.testing.assert_close(lp, rlp, rtol=1e-3, atol=1e-3)


def test_topk_histogram_stays_inside_the_top_k(golden_dir):
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    n, k = 20000, 5
    trace = {}
    DO.generate_with_kv_cache(sd, shape.nhead, z[:1], stoich_pred=stoich[:1], temperature=1.0, max_len=2,
                              heads_pred={kk: v[:1] for kk, v in heads.items()}, trace=trace)
    top = set(trace["final_logits"][0][0].topk(k).indices.tolist())
    mem = dec.precompute_memory(_cuda(z[:1]), None, _cuda(stoich[:1]), _cuda({kk: v[:1] for kk, v in heads.items()}))
    old = dec.max_rows_per_call
    dec.max_rows_per_call = n
    try:
        t, _, _ = dec.generate_with_kv_cache(None, temperature=1.0, top_k=k, max_len=2, _seed=9,
                                             cached_memory=mem.expand(n, -1, -1).contiguous())
    finally:
        dec.max_rows_per_call = old
    seen = set(t[:, 0].cpu().tolist())
    assert seen == top, (seen, top)


def test_sampling_distribution_matches_reference_softmax(golden_dir):
    """First-step token histogram over 200k draws of one latent vs softmax(logits/T) of the oracle:
    chi-square over bins with expected count >= 20 must stay below the 99.9% quantile."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    n, T = 200_000, 1.2
    trace = {}
    DO.generate_with_kv_cache(sd, shape.nhead, z[:1], stoich_pred=stoich[:1], temperature=T, max_len=2,
                              heads_pred={k: v[:1] for k, v in heads.items()}, trace=trace)
    p = trace["probs"][0][0].double()
    mem = dec.precompute_memory(_cuda(z[:1]), None, _cuda(stoich[:1]), _cuda({k: v[:1] for k, v in heads.items()}))
    old = dec.max_rows_per_call
    dec.max_rows_per_call = n
    try:
        t, lp, _ = dec.generate_with_kv_cache(None, temperature=T, max_len=2, cached_memory=mem.expand(n, -1, -1).contiguous(),
                                              return_log_probs=True, _seed=2024)
    finally:
        dec.max_rows_per_call = old
    counts = torch.bincount(t[:, 0].cpu(), minlength=shape.vocab_size).double()
    exp = p * n
    big = exp >= 20
    chi2 = float((((counts - exp) ** 2) / exp)[big].sum() + ((counts[~big].sum() - exp[~big].sum()) ** 2) / max(float(exp[~big].sum()), 1e-9))
    dof = int(big.sum())
    # Wilson-Hilferty upper 99.9% quantile of chi-square(dof)
    q = dof * (1 - 2 / (9 * dof) + 3.09 * math.sqrt(2 / (9 * dof))) ** 3
    assert chi2 < q, (chi2, q, dof)
    torch.testing.assert_close(lp[:, 0].cpu().double(), p[t[:, 0].cpu()].clamp(min=1e-8).log(), rtol=1e-3, atol=1e-3)


def test_h2_uniform_fallback_with_masks(golden_dir):
    """With hard masks the reference samples uniformly over the whole vocabulary."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, "tiny")
    t, lp, en, mk = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), temperature=1.2,
                                             max_len=shape.max_len, stop_boost=10.0, hard_stop_threshold=0.8,
                                             heads_pred=_cuda(heads), type_masks=_cuda(masks), _seed=5)
    lnv = math.log(shape.vocab_size)
    torch.testing.assert_close(lp.cpu(), torch.full_like(lp.cpu(), -lnv), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(en.cpu(), torch.full_like(en.cpu(), lnv), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(g["sample_masked_logprobs"][:, :1], torch.full((t.shape[0], 1), -lnv), rtol=1e-5, atol=1e-5)
    # "fixed" mode: masked softmax sampling; every sampled token respects the predicted type's mask
    dec.h2_uniform_fallback = False
    try:
        t, lp, en, mk = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), temperature=1.2,
                                                 max_len=shape.max_len, stop_boost=10.0, heads_pred=_cuda(heads),
                                                 type_masks=_cuda(masks), _seed=5)
    finally:
        dec.h2_uniform_fallback = True
    rt, rlp, _ = DO.generate_with_kv_cache(sd, shape.nhead, z, stoich_pred=stoich, temperature=1.2,
                                           max_len=shape.max_len, stop_boost=10.0, heads_pred=heads, type_masks=masks,
                                           return_log_probs=True, forced_tokens=t.cpu(), trace=(tr := {}))
    for s in range(t.shape[1]):
        fl = tr["final_logits"][s]
        assert torch.isfinite(fl.gather(1, t.cpu()[:, s:s + 1])).all(), "sampled a masked token"
        ref = torch.log_softmax(fl / 1.2, dim=-1).gather(1, t.cpu()[:, s:s + 1]).squeeze(1)
        torch.testing.assert_close(lp.cpu()[:, s], ref, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("rows", [5, 300])
def test_h2_uniform_fallback_without_masks_is_reported_by_the_projections(rows):
    """Without a type mask / hard stop the batch-global "some adjusted logit is NaN or +-inf" test of the reference
    (:1464-1466) is not a separate pass over the logits: the logits projection and the stop head report non-finite outputs
    themselves (LinearArgs::nonfinite_flag).  One NaN vocabulary bias or a NaN stop bias makes EVERY row sample uniformly
    (log-prob = -log V, entropy = log V, like the oracle); an infinite stop bias does not (sigmoid(inf) = 1 is finite);
    clean weights do not.  5 rows: persistent small-batch machinery; 300 rows: tensor-core projections."""
    shape = W.TINY
    lnv = math.log(shape.vocab_size)
    z = W.make_latents(rows, shape.latent_dim, 3)
    stoich, heads = W.make_conditioning(rows, shape.stoich_input_dim, 3)
    kw = dict(temperature=1.2, max_len=shape.max_len, stop_boost=10.0)
    for key, idx, val, degenerate in (("output_proj.4.bias", 7, float("nan"), True), ("output_proj.4.bias", 7, float("inf"), True),
                                      ("stop_head.2.bias", 0, float("nan"), True), ("stop_head.2.bias", 0, float("inf"), False),
                                      (None, 0, 0.0, False)):
        sd = W.make_decoder_state_dict(shape, 0)
        if key is not None:
            sd[key] = sd[key].clone()
            sd[key][idx] = val
        dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shape.nhead, device=DEV)
        t, lp, en, mk = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), _seed=9, **kw)
        rt, rlp, ren = DO.generate_with_kv_cache(sd, shape.nhead, z, stoich_pred=stoich, heads_pred=heads, return_log_probs=True,
                                                 return_entropy=True, forced_tokens=t.cpu(), **kw)
        L = min(t.shape[1], rt.shape[1])
        if degenerate:
            torch.testing.assert_close(lp.cpu(), torch.full_like(lp.cpu(), -lnv), rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(en.cpu(), torch.full_like(en.cpu(), lnv), rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(rlp[:, :L], torch.full_like(rlp[:, :L], -lnv), rtol=1e-5, atol=1e-5)      # the oracle agrees
        else:
            assert float((lp.cpu() + lnv).abs().max()) > 1e-3
            torch.testing.assert_close(lp.cpu()[:, :L], rlp[:, :L], rtol=1e-3, atol=1e-3)
            torch.testing.assert_close(en.cpu()[:, :L], ren[:, :L], rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------------------------------ encoder
def test_encoder_matches_reference_golden(golden_dir):
    g = _golden(golden_dir, "encoder_default")
    sd = W.make_encoder_state_dict(W.ENC_DEFAULT, 1)
    enc = S.FullMaterialsVAE.from_state_dict(sd, device=DEV)
    idx, frac, mask, magpie, tc = W.make_compositions(g["meta"]["B"], g["meta"]["seed_in"])
    out = enc(idx.to(DEV), frac.to(DEV), mask.to(DEV), magpie.to(DEV), tc.to(DEV))
    fused = enc.encode(idx.to(DEV), frac.to(DEV), mask.to(DEV), magpie.to(DEV), tc.to(DEV))["fused_repr"]
    for k, v in g.items():
        if k == "meta":
            continue
        got = fused if k == "fused_repr" else out[k]
        torch.testing.assert_close(got.cpu(), v, rtol=5e-4, atol=5e-5, msg=lambda m, k=k: f"{k}: {m}")
    stoich, heads = enc.conditioning(out["z"])
    ref = EO.forward(sd, idx, frac, mask, magpie, tc)
    rs, rh = EO.conditioning(ref)
    torch.testing.assert_close(stoich.cpu(), rs, rtol=5e-4, atol=5e-5)
    hin = enc.heads_from_latent(out["z"])["heads_input"].cpu()
    torch.testing.assert_close(hin, DO.heads_input_matrix(rh, rs.shape[0]), rtol=5e-4, atol=5e-5)


def test_encoder_52k_rows_relative_l2(golden_dir):
    """BASELINE config 5 at full size: rel-L2 <= 2e-3 on z and memory vs the fp32 oracle on a 512-row sample,
    and full-size run is finite and deterministic."""
    sd_e = W.make_encoder_state_dict(W.ENC_DEFAULT, 1)
    sd_d = W.make_decoder_state_dict(W.C512, 0)
    enc = S.FullMaterialsVAE.from_state_dict(sd_e, device=DEV)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd_d, nhead=8, device=DEV)
    n = 52800
    idx, frac, mask, magpie, tc = W.make_compositions(n, 7)
    z = enc.encode(idx.to(DEV), frac.to(DEV), mask.to(DEV), magpie.to(DEV), tc.to(DEV))["z"]
    stoich, heads = enc.conditioning(z)
    mem = dec.precompute_memory(z, None, stoich, heads)
    assert tuple(mem.shape) == (n, 24, 512) and bool(torch.isfinite(mem).all())
    s = slice(0, 512)
    ref = EO.forward(sd_e, idx[s], frac[s], mask[s], magpie[s], tc[s])
    rs, rh = EO.conditioning(ref)
    rmem = DO.build_memory(sd_d, ref["z"], None, rs, rh)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(z[s].cpu(), ref["z"]) <= 2e-3
    assert rel(mem[s].cpu(), rmem) <= 2e-3
    z2 = enc.encode(idx.to(DEV), frac.to(DEV), mask.to(DEV), magpie.to(DEV), tc.to(DEV))["z"]
    assert torch.equal(z, z2)


# ------------------------------------------------------------------------------------------ latent walks
def test_slerp_matches_reference(golden_dir):
    g = _golden(golden_dir, "slerp")
    out = S.latent.slerp(g["z1"].to(DEV), g["z2"].to(DEV), g["t"].to(DEV))
    torch.testing.assert_close(out.cpu(), g["slerp"], rtol=1e-5, atol=1e-5)
    out = S.latent.slerp(g["z1"].to(DEV), (g["z1"] * 2.0).to(DEV), 0.25)     # near-parallel -> lerp fallback
    torch.testing.assert_close(out.cpu(), g["slerp_parallel"], rtol=1e-5, atol=1e-5)


def test_unique_sequences_and_decode_unique():
    """rows grouped on the device by the ids before their first END; strings built once per distinct row.
    Checked against the reference semantics (decode every row, then group: oracle/latent.py unique_formulas)."""
    from oracle import latent as OL
    tok = S.FractionAwareTokenizer(max_len=24, fractions=[f"{i + 1}/997" for i in range(64)])
    gen = torch.Generator().manual_seed(5)
    n, ln = 6000, 23
    pool = torch.randint(3, tok.vocab_size, (40, ln), generator=gen)              # 40 base rows -> many duplicates
    rows = pool[torch.randint(0, 40, (n,), generator=gen)].clone()
    ends = torch.randint(0, ln + 4, (n,), generator=gen)                          # >= ln: the row has no END
    for r in range(n):
        e = int(ends[r]) % 7 if r % 3 == 0 else int(ends[r])
        if e < ln:
            rows[r, e] = 2
            rows[r, e + 1:] = torch.randint(0, tok.vocab_size, (ln - e - 1,), generator=gen)   # garbage after END
    rows[::50, 0] = 0                                                              # a PAD before the END is skipped by decode
    formulas, inverse, counts = S.latent.decode_unique(tok, rows.to(DEV))
    uniq, inv2, cnt2, lengths = S.latent.unique_sequences(rows.to(DEV))
    assert torch.equal(inverse, inv2) and torch.equal(counts, cnt2)
    assert int(counts.sum()) == n and len(formulas) == counts.numel()
    ref_formulas, ref_inverse, ref_counts = OL.unique_formulas(rows, tok.decode)
    inv = inverse.cpu().tolist()
    assert all(formulas[inv[r]] == ref_formulas[ref_inverse[r]] for r in range(n))
    # per-formula totals agree (distinct id rows may spell the same formula; the device groups by ids)
    agg = {}
    for f, c in zip(formulas, counts.cpu().tolist()):
        agg[f] = agg.get(f, 0) + c
    assert agg == dict(zip(ref_formulas, ref_counts))
    assert counts.numel() < n // 4                                                # duplicates were actually merged
    first_end = [(r.tolist() + [2]).index(2) for r in rows]
    assert lengths.cpu().tolist() == first_end
    # every row equals its representative up to its first END
    rep = uniq[inverse].cpu()
    for r in range(0, n, 37):
        assert rep[r, :first_end[r]].tolist() == rows[r, :first_end[r]].tolist() and int(rep[r, first_end[r]:].abs().sum()) == 0


def test_element_similarity_kernel_matches_reference(golden_dir):
    """SURVEY 8 f2, second half: every candidate x target pair of the reference's element_similarity on the device
    (one kernel, doubles) against the reference function's own outputs (1e-12 abs) and, for 600 x 40 generated formulas,
    against the oracle restatement."""
    g = torch.load(os.path.join(golden_dir, "similarity.pt"), weights_only=False)
    F = g["formulas"]
    sim = S.latent.element_similarity_matrix(F, F, DEV)
    torch.testing.assert_close(sim.cpu(), g["similarity"], rtol=0, atol=1e-12)
    assert abs(S.latent.element_similarity("YBa2Cu3O7", "La(7/10)Sr(3/10)CuO4", DEV) - float(g["similarity"][0, 1])) < 1e-12
    gen = torch.Generator().manual_seed(8)
    els = ["La", "Sr", "Cu", "O", "Y", "Ba", "Fe", "As", "Se", "H", "Mg", "B", "Nb", "Sn", "Bi", "Ca"]

    def make(n):
        out = []
        for _ in range(n):
            k = int(torch.randint(1, 6, (1,), generator=gen))
            parts = []
            for e in torch.randperm(len(els), generator=gen)[:k].tolist():
                kind = int(torch.randint(0, 4, (1,), generator=gen))
                a, b = int(torch.randint(1, 9, (1,), generator=gen)), int(torch.randint(1, 11, (1,), generator=gen))
                parts.append(els[e] + ("", str(a), f"({a}/{b})", f"0.{a}")[kind])
            out.append("".join(parts))
        return out
    cand, targ = make(600), make(40)
    sim = S.latent.element_similarity_matrix(cand, targ, DEV).cpu()
    ref = torch.tensor([[OL.element_similarity(a, b) for b in targ] for a in cand], dtype=torch.float64)
    torch.testing.assert_close(sim, ref, rtol=0, atol=1e-12)


# ------------------------------------------------------------------------------------------ teacher-forced forward
@pytest.mark.parametrize("name,shape", [("tiny", W.TINY), ("c512", W.C512)])
def test_teacher_forced_forward_matches_reference(golden_dir, name, shape):
    """forward(z, target_tokens) with teacher_forcing_ratio = 1 against the reference's own outputs
    (tests/golden/forward_tf.pt): logits / stop / type / site-dup logits at every position, causal + padding masks.
    Tolerance: bf16-exact weights, fp32 accumulate, bf16 hi/lo activations on the tensor-core path -> 2e-3 abs on
    logits of magnitude ~1 (the fp32 oracle meets 2e-4 on CPU)."""
    g = torch.load(os.path.join(golden_dir, "forward_tf.pt"), weights_only=False)[name]
    sd = W.make_decoder_state_dict(shape, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shape.nhead, device=DEV)
    z = W.make_latents(g["B"], shape.latent_dim, g["seed_in"])
    stoich, heads = W.make_conditioning(g["B"], shape.stoich_input_dim, g["seed_in"])
    logits, gen, stop, typ, dup = dec(_cuda(z), g["target_tokens"].to(DEV), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads))
    assert logits.shape == g["logits"].shape and gen.dtype == torch.int64
    torch.testing.assert_close(logits.cpu(), g["logits"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(stop.cpu(), g["stop_logits"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(typ.cpu(), g["type_logits"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(dup.cpu(), g["site_dup_logits"], rtol=2e-3, atol=2e-3)
    assert (gen.cpu().to(torch.int16) == g["generated"]).float().mean() > 0.98
    # the per-(sequence, head) kernel with K / V staged in shared memory (default) and the per-row kernel give the same bits
    try:
        _lib.tune(attn_forward=1, attn_forward_min_ctas=1)
        fw_out = dec(_cuda(z), g["target_tokens"].to(DEV), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads))
        _lib.tune(attn_forward=0)
        ref_out = dec(_cuda(z), g["target_tokens"].to(DEV), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads))
    finally:
        _lib.tune(attn_forward=1, attn_forward_min_ctas=1024)
    for x, y, w in zip((logits, gen, stop, typ, dup), ref_out, fw_out):
        assert torch.equal(x, y) and torch.equal(w, y)
    # consistency with the decode path: teacher forcing on the ids the greedy decode emitted reproduces them
    t, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), temperature=0.001,
                                         max_len=g["target_tokens"].shape[1])
    tgt = torch.cat([torch.ones((t.shape[0], 1), dtype=torch.int64, device=t.device), t], dim=1)
    _, gen2, _, _, _ = dec(_cuda(z), tgt, stoich_pred=_cuda(stoich), heads_pred=_cuda(heads))
    agree = (gen2 == t)[tgt[:, :-1] != 0]          # PAD inputs are masked as keys in forward but not in generation
    assert agree.float().mean() > 0.99


@pytest.mark.parametrize("name,shape", [("tiny", W.TINY), ("c512", W.C512)])
def test_scheduled_sampling_forward_matches_reference(golden_dir, name, shape):
    """SURVEY 8 f3, teacher_forcing_ratio < 1 (reference :987-1082): two engine passes with the argmax of the first mixed
    into the inputs of the second, against the reference's own outputs (tests/golden/forward_ss.pt; plain ratio,
    position-dependent ratio, ratio 0).  The reference's torch.rand mask is reproduced on the CPU from its seed and
    injected; tolerance as for the single pass (2e-3 abs), and the mixed inputs must be the reference's (generated ids
    agree wherever the first pass has no near-tie: >= 98 %)."""
    g = torch.load(os.path.join(golden_dir, "forward_ss.pt"), weights_only=False)[name]
    base = torch.load(os.path.join(golden_dir, "forward_tf.pt"), weights_only=False)[name]
    sd = W.make_decoder_state_dict(shape, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shape.nhead, device=DEV)
    for c in g:
        B = c["B"]
        tgt = base["target_tokens"][:B]
        z = W.make_latents(base["B"], shape.latent_dim, base["seed_in"])[:B]
        stoich, heads = W.make_conditioning(base["B"], shape.stoich_input_dim, base["seed_in"])
        stoich, heads = stoich[:B], {k: v[:B] for k, v in heads.items()}
        torch.manual_seed(c["seed"])
        mask = DO.scheduled_sampling_mask(B, tgt.shape[1] - 1, c["ratio"], c["positional"], c["decay"])
        dec.use_position_dependent_tf, dec.tf_position_decay = c["positional"], c["decay"]
        logits, gen, stop, typ, dup = dec(_cuda(z), tgt.to(DEV), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads),
                                          teacher_forcing_ratio=c["ratio"], _use_gt_mask=mask)
        torch.testing.assert_close(logits.cpu(), c["logits"], rtol=2e-3, atol=2e-3)
        torch.testing.assert_close(stop.cpu(), c["stop_logits"], rtol=2e-3, atol=2e-3)
        torch.testing.assert_close(typ.cpu(), c["type_logits"], rtol=2e-3, atol=2e-3)
        torch.testing.assert_close(dup.cpu(), c["site_dup_logits"], rtol=2e-3, atol=2e-3)
        assert (gen.cpu().to(torch.int16) == c["generated"]).float().mean() > 0.98
    # without an injected mask the module draws it itself (device RNG, like the reference): shapes and determinism per seed
    torch.manual_seed(3)
    a = dec(_cuda(z), tgt.to(DEV), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), teacher_forcing_ratio=0.5)[0]
    torch.manual_seed(3)
    b = dec(_cuda(z), tgt.to(DEV), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), teacher_forcing_ratio=0.5)[0]
    assert a.shape == c["logits"].shape and torch.equal(a, b)


# ------------------------------------------------------------------------------------------ draft verification (f4)
@pytest.mark.parametrize("name", ["tiny", "c512_b32"])
def test_generate_with_draft_is_exactly_greedy(golden_dir, name):
    """greedy decoding by draft verification (one teacher-forced pass checks every position; wrong positions
    are replaced by the pass's own predictions and re-checked) returns exactly the tokens of the step-by-step decode - the
    reference goldens - whatever the draft: the true sequence (1 pass), the true sequence with corrupted positions, pure
    garbage (converges within L passes), and garbage with a 2-pass budget (unconverged rows fall back to the KV-cache decode)."""
    shape, g, sd, dec, z, stoich, heads, masks = _setup(golden_dir, name)
    kw = dict(stoich_pred=_cuda(stoich), max_len=shape.max_len, heads_pred=_cuda(heads), type_masks=_cuda(masks),
              stop_boost=10.0, hard_stop_threshold=0.8)
    ref, _, _ = dec.generate_with_kv_cache(_cuda(z), temperature=0.001, **kw)
    assert torch.equal(ref.cpu().to(torch.int16), g["greedy_masked_tokens"])
    is_end = ref == 2
    after = (torch.cumsum(is_end.int(), dim=1) - is_end.int()) > 0
    want = ref.masked_fill(after, 0)                                   # PAD after each row's first END

    def check(draft, max_passes):
        t, passes, fallback = dec.generate_with_draft(_cuda(z), draft, max_passes=max_passes, **kw)
        assert t.shape[1] <= want.shape[1] and torch.equal(t, want[:, :t.shape[1]]) and int(want[:, t.shape[1]:].abs().sum()) == 0
        return passes, fallback
    assert check(ref, 4) == (1, 0)                                     # a correct draft is accepted in one pass
    gen = torch.Generator().manual_seed(3)
    bad = ref.clone().cpu()
    for r in range(bad.shape[0]):                                      # two wrong tokens per row
        for p in torch.randint(0, bad.shape[1], (2,), generator=gen).tolist():
            bad[r, p] = int(torch.randint(3, shape.vocab_size, (1,), generator=gen))
    passes, fallback = check(bad.to(DEV), shape.max_len)
    assert fallback == 0 and passes <= shape.max_len - 1
    garbage = torch.randint(3, shape.vocab_size, tuple(ref.shape), generator=gen).to(DEV)
    passes, fallback = check(garbage, shape.max_len)
    assert fallback == 0
    passes, fallback = check(garbage, 2)
    assert passes == 2 and fallback > 0
    with pytest.raises(ValueError):
        dec.generate_with_draft(_cuda(z), ref, temperature=1.0, **kw)


# ------------------------------------------------------------------------------------------ full-size properties
def test_config2_4096_latents_properties(golden_dir):
    """BASELINE config 2 at full size (4096 latents, masks + stop head, greedy): bit-exact vs the oracle on the
    first 48 rows (positions up to each row's first END), run-to-run determinism, and batch invariance
    (rows decoded alone equal the same rows inside the 4096 batch)."""
    shape = W.C512
    sd = W.make_decoder_state_dict(shape, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=DEV)
    B = 4096
    z = W.make_latents(B, shape.latent_dim, 1234)
    stoich, heads = W.make_conditioning(B, shape.stoich_input_dim, 1234)
    masks = OV.type_masks()
    kw = dict(temperature=0.001, max_len=64, type_masks=_cuda(masks), stop_boost=10.0, hard_stop_threshold=0.8)
    t, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
    t2, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
    assert torch.equal(t, t2)
    n = 48
    sub = {k: v[:n] for k, v in heads.items()}
    ts, _, _ = dec.generate_with_kv_cache(_cuda(z[:n]), stoich_pred=_cuda(stoich[:n]), heads_pred=_cuda(sub), **kw)
    rt, _, _ = DO.generate_with_kv_cache(sd, 8, z[:n], stoich_pred=stoich[:n], temperature=0.001, max_len=64,
                                         heads_pred=sub, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
    assert torch.equal(ts.cpu(), rt)
    lens = DO.first_end_lengths(rt)
    tc = t.cpu()
    for r in range(n):
        k = int(lens[r])
        assert torch.equal(tc[r, :k], rt[r, :k]), f"row {r} differs inside the 4096 batch"
    assert tc.shape[1] >= rt.shape[1]
    # every row that finished ends with END exactly once up to its length
    fl = DO.first_end_lengths(tc)
    assert int((tc == 2).any(dim=1).sum()) >= 1 and int(fl.max()) <= tc.shape[1]


# ------------------------------------------------------------------------------------------ run-time tunables
@pytest.mark.parametrize("rows", [6, 17, 32])
def test_cluster_parallel_small_batch_kernel_matches_default(rows):
    """The opt-in cluster-parallel small-batch decode (csrc/decode_cluster.cu: rows dealt to 8-CTA clusters, weights streamed
    with cp.async.bulk, activations exchanged through distributed shared memory) decodes the same tokens as the default
    grid-barrier kernel: plain greedy with masks + stop head (whole decode in one launch), plain greedy over all 63 steps,
    and a sampled rollout (one launch per step; log-probs within 1e-4)."""
    sd = W.make_decoder_state_dict(W.C512, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=DEV)
    z = W.make_latents(rows, 2048, 99)
    stoich, heads = W.make_conditioning(rows, 13, 99)
    masks = OV.type_masks()
    cases = (dict(temperature=0.001, max_len=64, type_masks=_cuda(masks), stop_boost=10.0, hard_stop_threshold=0.8),
             dict(temperature=0.001, max_len=64),
             dict(temperature=1.2, max_len=64, stop_boost=10.0, return_log_probs=True, _seed=5))
    try:
        for kw in cases:
            _lib.tune(cluster=0)
            t0, lp0, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
            _lib.tune(cluster=1)
            t1, lp1, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
            assert t0.shape == t1.shape and torch.equal(t0, t1)
            if lp0 is not None:
                torch.testing.assert_close(lp0, lp1, rtol=1e-4, atol=1e-4)
    finally:
        _lib.tune(cluster=0)


@pytest.mark.parametrize("rows", [24, 300, 2304])
def test_compact_finished_matches_reference_semantics_up_to_end(rows):
    """Opt-in retirement of finished rows (decoder.compact_finished): rows that have emitted END leave the batch
    (slots are compacted after every step, projections / attention / sampler skip the empty slots).  Against the default
    (reference semantics: finished rows keep decoding): identical tokens, log-probs and entropy up to and including every
    row's first END, PAD / 0 after it, same executed length; on a sampled rollout whose rows end at very different steps
    (24 rows: persistent small-batch kernel; 300: one stream; 2304: two sub-batch streams)."""
    sd = W.make_decoder_state_dict(W.C512, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=DEV)
    z = W.make_latents(rows, 2048, 4242)
    stoich, heads = W.make_conditioning(rows, 13, 4242)
    for kw in (dict(temperature=1.2, max_len=64, stop_boost=10.0, return_log_probs=True, return_entropy=True, _seed=17),
               dict(temperature=0.001, max_len=64, type_masks=_cuda(OV.type_masks()), stop_boost=10.0, hard_stop_threshold=0.8)):
        t0, lp0, en0 = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
        dec.compact_finished = True
        try:
            t1, lp1, en1 = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
        finally:
            dec.compact_finished = False
        assert t0.shape == t1.shape
        is_end = t0 == 2
        after = (torch.cumsum(is_end.int(), dim=1) - is_end.int()) > 0          # strictly after the first END
        assert torch.equal(t1[~after], t0[~after]) and int(t1[after].abs().sum()) == 0
        if lp0 is not None:
            assert torch.equal(lp1[~after], lp0[~after]) and float(lp1[after].abs().sum()) == 0.0
            assert torch.equal(en1[~after], en0[~after]) and float(en1[after].abs().sum()) == 0.0
            ends = torch.where(is_end.any(dim=1), is_end.int().argmax(dim=1), torch.full((rows,), t0.shape[1], device=t0.device))
            assert int(ends.max()) - int(ends.min()) >= 5                          # the rollout really is ragged


def test_compact_finished_with_shared_memory_tokens():
    """Both opt-ins at once: RLOO rollouts whose samples share a latent's memory tokens (`_n_samples`) while finished rows are
    retired.  Compaction scatters a latent's samples over the slots, so cross-attention goes back to the per-row kernel with
    the slot -> row -> latent lookup; outputs up to every row's first END equal the plain repeated-input rollout."""
    sd = W.make_decoder_state_dict(W.C512, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=DEV)
    base, k = 300, 4
    z = W.make_latents(base, 2048, 77)
    stoich, heads = W.make_conditioning(base, 13, 77)
    kw = dict(temperature=1.2, max_len=64, stop_boost=10.0, _seed=5)
    rep = lambda t: t.repeat(k, *([1] * (t.dim() - 1)))
    t0, lp0, en0, mk0 = dec.sample_for_reinforce(_cuda(rep(z)), stoich_pred=_cuda(rep(stoich)),
                                                 heads_pred=_cuda({n: rep(v) for n, v in heads.items()}), **kw)
    dec.compact_finished = True
    try:
        t1, lp1, en1, mk1 = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), _n_samples=k, **kw)
    finally:
        dec.compact_finished = False
    assert t0.shape == t1.shape and torch.equal(mk0, mk1)
    keep = mk0.bool()                                   # up to and including the first END
    assert torch.equal(t1[keep], t0[keep]) and torch.equal(lp1[keep], lp0[keep]) and torch.equal(en1[keep], en0[keep])
    assert int(t1[~keep].abs().sum()) == 0


@pytest.mark.parametrize("base,k,shape", [(96, 4, "C512"), (1024, 4, "C512"), (700, 3, "C512"), (512, 6, "C512"), (160, 5, "C576"),
                                          (70, 2, "TINY")])
def test_rloo_samples_sharing_memory_match_repeated_inputs(base, k, shape):
    """RLOO rollouts with `_n_samples = k` (the k samples of a latent share its memory tokens and their projected K / V
    inside the engine, include/scvae_b200.h memory_rows) against the reference's way of passing z.repeat(k, 1) etc.
    (scripts/train_v12_clean.py:2677-2688): bit-identical tokens, log-probs and entropy in the same sample-major layout
    (one stream at 384 rows, 4096 rows, a row count that does not split into equal sub-batches, more samples than the
    group kernel serves in one pass), for the group kernel over all rows (default), the group kernel inside two sub-batch
    row ranges (which split a latent's samples, or do not hold whole groups and fall back) and the per-row kernel."""
    shp = getattr(W, shape)
    sd = W.make_decoder_state_dict(shp, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=shp.nhead, device=DEV)
    z = W.make_latents(base, shp.latent_dim, 321)
    stoich, heads = W.make_conditioning(base, shp.stoich_input_dim, 321)
    kw = dict(temperature=1.2, max_len=min(64, shp.max_len), stop_boost=10.0, _seed=23)
    rep = lambda t: t.repeat(k, *([1] * (t.dim() - 1)))
    t0, lp0, en0, mk0 = dec.sample_for_reinforce(_cuda(rep(z)), stoich_pred=_cuda(rep(stoich)),
                                                 heads_pred=_cuda({n: rep(v) for n, v in heads.items()}), **kw)
    try:
        for tune in (dict(attn_shared=1, subbatches=0), dict(attn_shared=1, subbatches=2), dict(attn_shared=0, subbatches=0)):
            _lib.tune(**tune)
            t1, lp1, en1, mk1 = dec.sample_for_reinforce(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), _n_samples=k, **kw)
            assert t1.shape == t0.shape and t1.shape[0] == base * k, tune
            assert torch.equal(t0, t1) and torch.equal(lp0, lp1) and torch.equal(en0, en1) and torch.equal(mk0, mk1), tune
    finally:
        _lib.tune(attn_shared=1, subbatches=0)
    mem = dec.precompute_memory(_cuda(z), None, _cuda(stoich), _cuda(heads))
    t2, _, _ = dec.generate_with_kv_cache(None, temperature=1.2, max_len=kw["max_len"], stop_boost=10.0, cached_memory=mem, _seed=23,
                                          _n_samples=k)
    assert torch.equal(t2, t0)


def test_tunables_never_change_tokens():
    """scv_tune moves work between streams / grids only: the opt-in launch configurations (grid-stride attention grid,
    forced GEMM pipeline depth, 1 / 3 sub-batch streams, bulk-copy staged cross-attention, no graph replay) decode the
    same tokens as the default on 2304 rows of the config-2 workload."""
    sd = W.make_decoder_state_dict(W.C512, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=DEV)
    B = 2304
    z = W.make_latents(B, 2048, 1234)
    stoich, heads = W.make_conditioning(B, 13, 1234)
    kw = dict(temperature=0.001, max_len=64, type_masks=_cuda(OV.type_masks()), stop_boost=10.0, hard_stop_threshold=0.8)
    defaults = dict(attn_ctas_per_sm=0, gemm_stages=0, subbatches=0, graph=1, attn_bulk=0, attn_bulk_piece_kb=0, gemm_bn64=8, gemm_mc=0, gemm_mc_min_row_tiles=9)
    base, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
    try:
        for cfg in (dict(attn_ctas_per_sm=3, gemm_stages=2), dict(subbatches=1), dict(subbatches=3), dict(attn_bulk=1),
                    dict(attn_bulk=1, attn_bulk_piece_kb=16), dict(graph=0), dict(gemm_bn64=0), dict(gemm_bn64=16), dict(gemm_bn64=16, subbatches=3),
                    dict(gemm_mc=1, gemm_mc_min_row_tiles=1), dict(gemm_mc=1, gemm_mc_min_row_tiles=1, subbatches=1, gemm_stages=4)):
            _lib.tune(**{**defaults, **cfg})
            t, _, _ = dec.generate_with_kv_cache(_cuda(z), stoich_pred=_cuda(stoich), heads_pred=_cuda(heads), **kw)
            assert torch.equal(t, base), cfg
    finally:
        _lib.tune(**defaults)


# ------------------------------------------------------------------------------------------ opt-in kernels
@pytest.mark.gpu
def test_optin_kernels_match_default_tokens():
    """The alternative kernels must decode the same greedy tokens as the default path: 32 rows (default: the whole decode
    in the persistent small-batch kernel; SCV_SMALL=0: per-projection kernels; SCV_SMALL_PERSIST=0: one persistent kernel
    per step), 64 rows (SCV_SMALL_MAX_ROWS=64: two row groups per phase of the persistent kernel) and 768 rows (multi-wave projections; opt-in fused residual + LayerNorm cluster projection, persistent
    tcgen05 projection and CTA-pair (cta_group::2) projection), with and without masks / stop head."""
    import subprocess
    import sys as _sys
    script = os.path.join(os.path.dirname(__file__), "optin_check.py")

    def digest(extra):
        env = dict(os.environ)
        env.update(extra)
        r = subprocess.run([_sys.executable, script], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")]
        assert line, r.stdout[-2000:]
        return line[0]

    base = digest({})
    for extra in ({"SCV_SMALL": "0"}, {"SCV_SMALL_PERSIST": "0"}, {"SCV_SMALL_MAX_ROWS": "64"}, {"SCV_SMALL_FUSE_FFN": "1"}, {"SCV_FUSE_LN": "1"},
                  {"SCV_GEMM_PERSISTENT": "1"}, {"SCV_GEMM_2CTA": "1"}):
        assert digest(extra) == base, f"{extra} decodes different tokens"
