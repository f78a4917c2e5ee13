"""Probe (not a test): the cluster-parallel small-batch decode (csrc/decode_cluster.cu) against the grid-barrier kernel and
the per-projection path, tokens and time per step.   usage: python tests/cluster_check.py [rows ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as W      # noqa: E402
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer      # noqa: E402

rows_list = [int(a) for a in sys.argv[1:]] or [32, 6, 17, 64, 1]
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                             isotopes=[f"{300 + i}Og" for i in range(291)])
masks = tok.get_type_masks(dev)


def run(z, st, hp, kw, reps=3):
    for _ in range(2):
        out = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps


for rows in rows_list:
    z = W.make_latents(rows, 2048, 1234).to(dev)
    st, hp = W.make_conditioning(rows, 13, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    for name, kw in (("masks+stop", dict(temperature=0.001, max_len=64, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)),
                     ("plain63", dict(temperature=0.001, max_len=64)),
                     ("sample", dict(temperature=1.2, max_len=64, stop_boost=10.0, return_log_probs=True, return_entropy=True, _seed=7))):
        _lib.tune(cluster=0)
        (t0, lp0, en0), dt0 = run(z, st, hp, kw)
        _lib.tune(cluster=1)
        (t1, lp1, en1), dt1 = run(z, st, hp, kw)
        L0, L1 = t0.shape[1], t1.shape[1]
        same = L0 == L1 and bool(torch.equal(t0, t1))
        extra = ""
        if lp0 is not None and same:
            extra = f" max|dlogp|={float((lp0 - lp1).abs().max()):.2e} max|dH|={float((en0 - en1).abs().max()):.2e}"
        elif L0 == L1:
            extra = f" rows differing: {int((t0 != t1).any(dim=1).sum())}"
        if not same:
            Lm = min(L0, L1)
            d = (t0[:, :Lm] != t1[:, :Lm])
            bad = d.any(dim=1).nonzero().flatten().tolist()
            ends0 = [(r.tolist() + [2]).index(2) for r in t0.cpu()]
            ends1 = [(r.tolist() + [2]).index(2) for r in t1.cpu()]
            print(f"   rows differing within the first {Lm} steps: {bad}; first END old {ends0[:34]} cluster {ends1[:34]}")
            for r in bad[:3]:
                pos = int(d[r].nonzero()[0])
                print(f"   row {r} first differs at step {pos}: old {t0[r, max(0, pos - 2):pos + 3].tolist()} cluster {t1[r, max(0, pos - 2):pos + 3].tolist()}")
        print(f"rows={rows:3d} {name:10s}: steps {L0}/{L1} same_tokens={same}{extra}  old {1e6 * dt0 / L0:7.1f} us/step  "
              f"cluster {1e6 * dt1 / L1:7.1f} us/step", flush=True)
