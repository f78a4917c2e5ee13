"""Small-batch RL rollout timing (not a pytest file): sample_for_reinforce on B rows, temperature 1.2, log-probs + entropy."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import superconductor_vae_b200 as S
from superconductor_vae_b200 import synthetic as Sy

dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(Sy.make_decoder_state_dict(Sy.C512, 0), nhead=8, device=dev)
for B in [int(a) for a in sys.argv[1:]] or [32]:
    z = Sy.make_latents(B, 2048, 1234).to(dev)
    st, hp = Sy.make_conditioning(B, 13, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    fn = lambda: dec.sample_for_reinforce(z, stoich_pred=st, temperature=1.2, max_len=64, heads_pred=hp, _seed=7)
    for _ in range(2):
        t = fn()[0]
    torch.cuda.synchronize()
    n0, t0, n = S.launch_count(), time.perf_counter(), 5
    for _ in range(n):
        t = fn()[0]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"B={B}: sampling, {dt*1e3:.2f} ms per rollout of {t.shape[1]} steps, {dt/t.shape[1]*1e6:.0f} us/step, "
          f"{(S.launch_count() - n0) / n / t.shape[1]:.1f} launches/step", flush=True)
