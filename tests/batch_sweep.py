"""Probe (not a test): time per decode step against the batch size (config-2 options, C512).
usage: python tests/batch_sweep.py [rows ...] [tunable=value ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as W      # noqa: E402
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer      # noqa: E402

_lib.tune(**{a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[1:] if "=" in a})
rows_list = [int(a) for a in sys.argv[1:] if "=" not in a] or [1, 8, 32, 33, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
dec.max_rows_per_call = max(dec.max_rows_per_call, max(rows_list))
tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                             isotopes=[f"{300 + i}Og" for i in range(291)])
masks = tok.get_type_masks(dev)
for rows in rows_list:
    z = W.make_latents(rows, 2048, 1234).to(dev)
    st, hp = W.make_conditioning(rows, 13, 1234)
    st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
    for name, kw in (("masks+stop", dict(temperature=0.001, max_len=64, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)),
                     ("plain 63 steps", dict(temperature=0.001, max_len=64))):
        fn = lambda: dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            t, _, _ = fn()
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / reps
        L = t.shape[1]
        print(f"rows={rows:6d} {name:15s}: {ms:8.2f} ms per call, {L:2d} steps, {1e3 * ms / L:8.1f} us per step, "
              f"{1e3 * ms / L / rows:7.3f} us per row-step, {rows / ms:8.1f} K formulas/s", flush=True)
