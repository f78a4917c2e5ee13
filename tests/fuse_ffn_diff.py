"""Probe (not a test): where do the fused feed-forward phases (SCV_SMALL_FUSE_FFN=1) and the two projection phases
decode different greedy tokens on the 32-row golden inputs (plain 63-step decode and masked decode)?"""
import os, subprocess, sys
import torch
here = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(here))
    import superconductor_vae_b200 as S
    from superconductor_vae_b200 import synthetic as W
    from oracle import vocab as OV
    dev = "cuda:0"
    shape = W.C512
    dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(shape, 0), nhead=8, device=dev)
    z = W.make_latents(32, shape.latent_dim, 1234).to(dev)
    st, hp = W.make_conditioning(32, shape.stoich_input_dim, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    kw = dict(stoich_pred=st, heads_pred=hp, temperature=0.001, max_len=shape.max_len)
    t1, _, _ = dec.generate_with_kv_cache(z, **kw)
    t2, _, _ = dec.generate_with_kv_cache(z, type_masks=OV.type_masks().to(dev), stop_boost=10.0, hard_stop_threshold=0.8, **kw)
    torch.save({"plain": t1.cpu(), "masked": t2.cpu()}, sys.argv[2])
    sys.exit(0)
out = {}
for v in ("0", "1"):
    e = dict(os.environ); e["SCV_SMALL_FUSE_FFN"] = v
    f = f"gpurun_out/fuse_{v}.pt"
    r = subprocess.run([sys.executable, __file__, "child", f], env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out[v] = torch.load(f)
g = torch.load(os.path.join(here, "golden", "c512_b32.pt"), weights_only=False)
print("golden keys", [k for k in g.keys()][:12])
for k in ("plain", "masked"):
    a, b = out["0"][k], out["1"][k]
    L = min(a.shape[1], b.shape[1])
    d = (a[:, :L] != b[:, :L])
    rows = d.any(1).nonzero().flatten().tolist()
    print(k, "shapes", tuple(a.shape), tuple(b.shape), "rows differing", rows, "first positions", [int(d[r].float().argmax()) for r in rows])
