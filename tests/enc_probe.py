import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
import superconductor_vae_b200 as S
from superconductor_vae_b200 import synthetic as Sy, _lib
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(Sy.make_decoder_state_dict(Sy.C512, 0), nhead=8, device=dev)
enc = S.FullMaterialsVAE.from_state_dict(Sy.make_encoder_state_dict(Sy.ENC_DEFAULT, 1), device=dev)
n = 52800
idx, frac, mask, magpie, tc = (t.to(dev) for t in Sy.make_compositions(n, 7))
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): out = fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it, out
for rep in range(3):
    ms1, o = t(lambda: enc.encode(idx, frac, mask, magpie, tc)["z"])
    z = o
    ms2, (st, hp) = t(lambda: enc.conditioning(z))
    ms3, mem = t(lambda: dec.precompute_memory(z, None, st, hp))
    print(f"encode {ms1:.2f} ms  heads {ms2:.2f} ms  memory {ms3:.2f} ms  total {ms1+ms2+ms3:.2f}", flush=True)
_lib.profile_begin()
z = enc.encode(idx, frac, mask, magpie, tc)["z"]; st, hp = enc.conditioning(z); dec.precompute_memory(z, None, st, hp)
print(_lib.profile_end())
