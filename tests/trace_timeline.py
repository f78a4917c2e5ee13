"""Probe (not a test): CTA residency timeline of one 4096-latent decode (scv_trace_begin / scv_trace_read).

usage: python tests/trace_timeline.py [rows] [spec ...]      spec as in tests/overlap_sweep.py
For every configuration: which share of the decode's wall time has attention CTAs resident, projection CTAs resident,
both, or neither (GPU-wide and per SM), and the mean number of resident CTAs of each kind.  This is the evidence for
(or against) HBM-bound attention and tensor-bound projections of different sub-batch streams sharing the SMs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as W      # noqa: E402
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer      # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
specs = sys.argv[2:] or ["subbatches=0"]
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
dec.max_rows_per_call = max(dec.max_rows_per_call, rows)
tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                             isotopes=[f"{300 + i}Og" for i in range(291)])
masks = tok.get_type_masks(dev)
z = W.make_latents(rows, 2048, 1234).to(dev)
st, hp = W.make_conditioning(rows, 13, 1234)
st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
kw = dict(temperature=0.001, max_len=64, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
defaults = dict(attn_ctas_per_sm=0, gemm_stages=0, subbatches=0, graph=1, sub_min_rows=2048, attn_bulk=0, attn_bulk_min_rows=256, attn_bulk_piece_kb=0)


def union_time(t0, t1):
    """Total length of the union of intervals."""
    if len(t0) == 0:
        return 0.0
    o = np.argsort(t0)
    t0, t1 = t0[o], t1[o]
    end = np.maximum.accumulate(t1)
    gaps = np.maximum(t0[1:] - end[:-1], 0)
    return float((end[-1] - t0[0]) - gaps.sum())


def both_time(a0, a1, b0, b1):
    """Length of (union A) intersect (union B) = |A| + |B| - |A u B|."""
    return union_time(a0, a1) + union_time(b0, b1) - union_time(np.concatenate([a0, b0]), np.concatenate([a1, b1]))


for spec in specs:
    cfg = dict(defaults)
    for kv in spec.split(","):
        if kv:
            k, v = kv.split("=")
            cfg[k] = int(v)
    _lib.tune(**cfg)
    for _ in range(2):
        dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
    torch.cuda.synchronize()
    _lib.trace_begin(1 << 22)
    t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
    rec = _lib.trace_read(1 << 22)
    rec = rec[rec["t1"] > 0]
    t0 = rec["t0"].astype(np.float64)
    t1 = rec["t1"].astype(np.float64)
    base = t0.min()
    t0, t1 = (t0 - base) / 1e3, (t1 - base) / 1e3          # microseconds
    span = t1.max()
    # look at the middle half of the decode (steady state, away from the memory builder and the tail)
    lo, hi = 0.25 * span, 0.75 * span
    sel = (t0 >= lo) & (t1 <= hi)
    att = sel & (rec["kid"] < 100)
    gem = sel & (rec["kid"] >= 100)
    win = hi - lo
    ua, ug = union_time(t0[att], t1[att]), union_time(t0[gem], t1[gem])
    ub = both_time(t0[att], t1[att], t0[gem], t1[gem])
    print(f"== {spec}: {len(rec)} CTAs traced, decode span {span / 1e3:.2f} ms, {t.shape[1]} steps; window {win / 1e3:.2f} ms")
    print(f"   GPU-wide: attention resident {100 * ua / win:5.1f} %, projections resident {100 * ug / win:5.1f} %, "
          f"both {100 * ub / win:5.1f} %, neither {100 * (win - ua - ug + ub) / win:5.1f} %")
    print(f"   mean resident CTAs: attention {((t1 - t0)[att]).sum() / win:7.1f}, projections {((t1 - t0)[gem]).sum() / win:6.1f}; "
          f"mean CTA lifetime: attention {(t1 - t0)[att].mean():6.2f} us, projections {(t1 - t0)[gem].mean():6.2f} us")
    per_sm = []
    for sm in np.unique(rec["sm"]):
        a_, g_ = att & (rec["sm"] == sm), gem & (rec["sm"] == sm)
        per_sm.append((union_time(t0[a_], t1[a_]), union_time(t0[g_], t1[g_]), both_time(t0[a_], t1[a_], t0[g_], t1[g_])))
    ps = np.array(per_sm) / win
    print(f"   per SM (mean over {len(per_sm)} SMs): attention {100 * ps[:, 0].mean():5.1f} %, projections "
          f"{100 * ps[:, 1].mean():5.1f} %, both on the same SM {100 * ps[:, 2].mean():5.1f} %, idle of both "
          f"{100 * (1 - ps[:, 0] - ps[:, 1] + ps[:, 2]).mean():5.1f} %", flush=True)
    out = os.environ.get("TRACE_OUT")
    if out:
        np.save(out + "_" + spec.replace(",", "_").replace("=", "") + ".npy", rec)
