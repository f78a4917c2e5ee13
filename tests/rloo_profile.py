"""Probe (not a test): one config-3 RLOO rollout (2048 latents x 4 samples, T = 1.2) with a latent's memory tokens repeated
per sample, shared (per-row kernel, samples adjacent) and shared through the one-warp-per-(latent, head) kernel: time per
call, identical tokens, per-category kernel time.   usage: python tests/rloo_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as W      # noqa: E402

dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
B, k = 2048, 4
z = W.make_latents(B, 2048, 1234).to(dev)
st, hp = W.make_conditioning(B, 13, 1234)
st, hp = st.to(dev), {n: v.to(dev) for n, v in hp.items()}
zz, ss = z.repeat(k, 1), st.repeat(k, 1)
hh = {n: v.repeat(k, *([1] * (v.dim() - 1))) for n, v in hp.items()}
kw = dict(temperature=1.2, max_len=64, stop_boost=10.0, _seed=7)
shared = lambda: dec.sample_for_reinforce(z, stoich_pred=st, heads_pred=hp, _n_samples=k, **kw)
repeated = lambda: dec.sample_for_reinforce(zz, stoich_pred=ss, heads_pred=hh, **kw)
ref = None
for name, fn, tune in (("repeated", repeated, dict(attn_shared=0, subbatches=0)),
                       ("shared, per-row kernel", shared, dict(attn_shared=0, subbatches=0)),
                       ("shared, group kernel, 2 sub-batches", shared, dict(attn_shared=1, subbatches=0)),
                       ("shared, group kernel, 1 sub-batch", shared, dict(attn_shared=1, subbatches=1))):
    _lib.tune(**tune)
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        t = fn()
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 3
    if ref is None:
        ref = t
    same = all(torch.equal(x, y) for x, y in zip(ref[:2], t[:2]))
    dlp = float((ref[1] - t[1]).abs().max())
    _lib.profile_begin()
    fn()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    print(f"== {name}: {ms:.1f} ms per call, {t[0].shape[1]} steps, tokens and log-probs identical to repeated: {same} (max |dlogp| {dlp:.1e})")
    for c, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        if c.startswith("attention"):
            print(f"   {c:24s} {v['launches']:6d} launches {v['ms']:8.2f} ms  {1e3 * v['ms'] / v['launches']:8.2f} us avg", flush=True)
