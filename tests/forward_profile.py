"""Probe (not a test): per-category kernel time of the teacher-forced forward (f3) at several sizes.
usage: python tests/forward_profile.py [sequences ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as W      # noqa: E402

dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
for Bf in [int(a) for a in sys.argv[1:]] or [256, 1024]:
    Lf = 64
    z = W.make_latents(Bf, 2048, 1234).to(dev)
    st, hp = W.make_conditioning(Bf, 13, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    g = torch.Generator().manual_seed(3)
    tgt = torch.randint(3, dec.vocab_size, (Bf, Lf), generator=g)
    tgt[:, 0] = 1
    tgt[::2, 40:] = 0
    tgt = tgt.to(dev)
    mem = dec.precompute_memory(z, None, st, hp)
    fn = lambda: dec(z, tgt, cached_memory=mem)
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        fn()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 3
    _lib.profile_begin()
    fn()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    R = Bf * (Lf - 1)
    print(f"== {Bf} sequences x {Lf - 1} positions = {R} rows: {ms:.2f} ms per pass, {R / ms / 1e3:.2f} M positions/s; "
          f"projections alone at the sustained tensor rate (hi + lo): {2 * R * 93e6 / 1389.2e12 * 1e3:.2f} ms")
    for c, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {c:24s} {v['launches']:5d} launches {v['ms']:8.3f} ms  {1e3 * v['ms'] / v['launches']:8.1f} us avg  "
              f"{v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f} TFLOP/s  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f} GB/s", flush=True)
