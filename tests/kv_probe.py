"""Probe (not a test): how many rows of a large greedy batch equal the fp32 CPU oracle token for token (config 2
settings).  Measured on B200 with 2048 rows: 2048/2048 for the tensor-core path (bf16 hi/lo activations, fp32 K/V) and
for the CUDA-core path (SCV_LINEAR_IMPL=1); an experimental 3-byte K/V cache format (fp32 rounded to 16 significant
bits) gave 2046/2048 and was dropped.  usage: python tests/kv_probe.py [rows] [seed] [chunk]
chunk > 0 decodes the rows `chunk` at a time (chunk <= 64: every call goes through the persistent small-batch kernel,
csrc/decode_small.cu), compared against the same oracle rows."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import synthetic as W      # noqa: E402
from oracle import decoder_oracle as DO                 # noqa: E402
from oracle import vocab as OV                          # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1234
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
shape = W.C512
sd = W.make_decoder_state_dict(shape, 0)
dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device="cuda:0")
z = W.make_latents(rows, shape.latent_dim, seed)
stoich, heads = W.make_conditioning(rows, shape.stoich_input_dim, seed)
masks = OV.type_masks()


def cu(t):
    return {k: v.cuda() for k, v in t.items()} if isinstance(t, dict) else t.cuda()


kw = dict(temperature=0.001, max_len=64, stop_boost=10.0, hard_stop_threshold=0.8)
if chunk <= 0:
    t, _, _ = dec.generate_with_kv_cache(cu(z), stoich_pred=cu(stoich), heads_pred=cu(heads), type_masks=cu(masks), **kw)
else:
    t = torch.zeros((rows, 63), dtype=torch.long, device="cuda:0")
    zc, sc, hc, mc = cu(z), cu(stoich), cu(heads), cu(masks)
    for r0 in range(0, rows, chunk):
        sl = slice(r0, min(rows, r0 + chunk))
        tt, _, _ = dec.generate_with_kv_cache(zc[sl], stoich_pred=sc[sl], heads_pred={k: v[sl] for k, v in hc.items()},
                                              type_masks=mc, **kw)
        t[sl, : tt.shape[1]] = tt
t0 = time.time()
torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
rt, _, _ = DO.generate_with_kv_cache(sd, 8, z, stoich_pred=stoich, heads_pred=heads, type_masks=masks, **kw)
dt = time.time() - t0
lens = DO.first_end_lengths(rt)
tc = t.cpu()
bad = [r for r in range(rows) if not torch.equal(tc[r, : int(lens[r])], rt[r, : int(lens[r])])]
print(f"SCV_LINEAR_IMPL={os.environ.get('SCV_LINEAR_IMPL', '0')} "
      f"rows={rows} seed={seed} chunk={chunk}: {rows - len(bad)}/{rows} rows equal to the oracle up to their END "
      f"(oracle {dt:.1f} s); differing rows: {bad[:16]}")
# every differing row: where it leaves the oracle and how close the oracle's decision was there (a near-tie inside the
# engine's logit tolerance of 2e-4 can fall the other way; anything larger would be a bug)
for r in bad[:8]:
    n_cmp = min(int(lens[r]), tc.shape[1], rt.shape[1])
    pos = int((tc[r, :n_cmp] != rt[r, :n_cmp]).nonzero()[0])
    trace = {}
    DO.generate_with_kv_cache(sd, 8, z[r:r + 1], stoich_pred=stoich[r:r + 1], heads_pred={k: v[r:r + 1] for k, v in heads.items()},
                              type_masks=masks, trace=trace, **kw)       # (same max_len: the length boost depends on it)
    lg = trace["final_logits"][pos][0]
    top = lg.topk(2)
    print(f"   row {r} leaves the oracle at step {pos}: oracle picks {int(top.indices[0])} over {int(top.indices[1])} by "
          f"{float(top.values[0] - top.values[1]):.3e} (logits {float(top.values[0]):.4f} / {float(top.values[1]):.4f}); "
          f"engine picked {int(tc[r, pos])}")
    if "type_logits" in trace and not bool(torch.isfinite(lg[int(tc[r, pos])])):
        tl = trace["type_logits"][pos][0]
        tt2 = tl.topk(2)
        print(f"      the engine's token is outside the oracle's type mask: the oracle's type head picks class {int(tt2.indices[0])} over "
              f"{int(tt2.indices[1])} by {float(tt2.values[0] - tt2.values[1]):.3e} (type logits {[round(float(v), 5) for v in tl]})")

