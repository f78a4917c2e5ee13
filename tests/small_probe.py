"""Probe (not a test): tokens / logits of a small batch through the path selected by SCV_SMALL (1 = persistent step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import superconductor_vae_b200 as S
from superconductor_vae_b200 import synthetic as Sy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
sd = Sy.make_decoder_state_dict(Sy.C512, 0)
dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=dev)
z = Sy.make_latents(B, 2048, 1234).to(dev)
st, hp = Sy.make_conditioning(B, 13, 1234)
st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
kw = {}
if len(sys.argv) > 3:
    from oracle import vocab as OV
    kw = dict(type_masks=OV.type_masks().to(dev), stop_boost=10.0, hard_stop_threshold=0.8)
t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, temperature=0.001, max_len=steps + 1, **kw)
torch.cuda.synchronize()
x = dec.debug_tap(0)
lg = dec.debug_tap(1)
print("SCV_SMALL", os.environ.get("SCV_SMALL", "1"), "tokens", t[:4].tolist())
print("tlog", dec.debug_tap(2)[:2].tolist(), "slog", dec.debug_tap(3)[:4].tolist())
print("x", float(x.abs().sum()), bool(torch.isnan(x).any()), "logits", float(lg.abs().sum()), bool(torch.isnan(lg).any()))
torch.save({"t": t.cpu(), "x": x.cpu(), "lg": lg.cpu()}, f"gpurun_out/small_probe_{os.environ.get('SCV_SMALL', '1')}.pt")
