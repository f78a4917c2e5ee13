"""Small-batch decode timing (not a pytest file): python tests/bench_small.py B [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import superconductor_vae_b200 as S
from superconductor_vae_b200 import synthetic as Sy

dev = "cuda:0"
sd = Sy.make_decoder_state_dict(Sy.C512, 0)
dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=dev)
for B in [int(a) for a in sys.argv[1:]] or [8, 32, 128]:
    z = Sy.make_latents(B, 2048, 1234).to(dev)
    st, hp = Sy.make_conditioning(B, 13, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    kw = dict(stoich_pred=st, heads_pred=hp, temperature=0.001, max_len=64)      # no masks / stop: all 63 steps run
    for _ in range(2):
        t, _, _ = dec.generate_with_kv_cache(z, **kw)
    torch.cuda.synchronize()
    n0 = S.launch_count()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        t, _, _ = dec.generate_with_kv_cache(z, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"B={B}: {dt*1e3:.2f} ms per decode of {t.shape[1]} steps -> {B/dt:.0f} formulas/s, {dt/t.shape[1]*1e6:.0f} us/step, {(S.launch_count() - n0) / n / t.shape[1]:.1f} launches/step "
          f"(tc_min_rows={os.environ.get('SCV_TC_MIN_ROWS','64')}, graph={os.environ.get('SCV_GRAPH','1')})", flush=True)
