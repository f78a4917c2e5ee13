"""Probe (not a test): where BASELINE config 5 (52,800 compositions -> z, heads, memory tokens) spends its time, by kernel
category (CUDA events around every launch, scv_profile_begin / scv_profile_end).   usage: python tests/enc_profile.py [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as Sy      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 52800
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(Sy.make_decoder_state_dict(Sy.C512, 0), nhead=8, device=dev)
enc = S.FullMaterialsVAE.from_state_dict(Sy.make_encoder_state_dict(Sy.ENC_DEFAULT, 1), device=dev)
idx, frac, mask, magpie, tc = (t.to(dev) for t in Sy.make_compositions(n, 7))


def parts():
    z = enc.encode(idx, frac, mask, magpie, tc)["z"]
    st, hp = enc.conditioning(z)
    return dec.precompute_memory(z, None, st, hp)


for _ in range(2):
    parts()
torch.cuda.synchronize()
for name, fn in (("encode", lambda: enc.encode(idx, frac, mask, magpie, tc)["z"]),):
    pass
z = enc.encode(idx, frac, mask, magpie, tc)["z"]
st, hp = enc.conditioning(z)
stages = (("encode (three branches -> z)", lambda: enc.encode(idx, frac, mask, magpie, tc)),
          ("heads_from_latent (conditioning)", lambda: enc.conditioning(z)),
          ("precompute_memory", lambda: dec.precompute_memory(z, None, st, hp)))
for name, fn in stages:
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        fn()
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 3
    _lib.profile_begin()
    fn()
    prof = _lib.profile_end()
    prof.pop("event_pair_overhead", None)
    print(f"{name:34s} {ms:7.2f} ms   " + "  ".join(f"{k}: {v['launches']} launches {v['ms']:.2f} ms ({v['flops'] / max(v['ms'], 1e-9) / 1e9:.0f} TFLOP/s)"
                                                      for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])), flush=True)
