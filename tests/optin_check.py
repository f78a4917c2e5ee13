"""Helper of test_optin_kernels_match (run in a subprocess because the switches are read once per process):
prints a digest of greedy tokens for a 32-row batch (golden-checked), a 64-row batch (two row groups of the
persistent small-batch kernel) and a 768-row batch (six row tiles: multi-wave projections, whole CTA pairs)."""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                      # noqa: E402
from superconductor_vae_b200 import synthetic as W       # noqa: E402
from oracle import vocab as OV                           # noqa: E402

dev = "cuda:0"
g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "c512_b32.pt"), weights_only=False)
shape = W.C512
sd = W.make_decoder_state_dict(shape, 0)
dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=dev)
masks = OV.type_masks().to(dev)
out = []
for B in (32, 64, 768):
    z = W.make_latents(B, shape.latent_dim, 1234).to(dev)
    st, hp = W.make_conditioning(B, shape.stoich_input_dim, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    kw = dict(stoich_pred=st, heads_pred=hp, temperature=0.001, max_len=shape.max_len)
    t1, _, _ = dec.generate_with_kv_cache(z, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8, **kw)
    t2, _, _ = dec.generate_with_kv_cache(z, **kw)
    out.append(hashlib.sha256(t1.cpu().numpy().tobytes() + t2.cpu().numpy().tobytes()).hexdigest()[:16])
print("DIGEST", " ".join(out))
