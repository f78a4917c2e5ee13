"""Probe (not a test): sweep the run-time launch tunables (scv_tune) on the config-2 workload and time whole decodes.

usage: python tests/overlap_sweep.py [rows] [spec ...]
  spec = comma-separated key=value pairs, e.g.  attn_ctas_per_sm=3,subbatches=2,gemm_stages=2
Every configuration decodes the same latents; the tokens must equal the first configuration's (the tunables only move
work between streams / SMs, never change arithmetic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import _lib, synthetic as W      # noqa: E402
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer      # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
specs = sys.argv[2:] or ["subbatches=0"]
dev = "cuda:0"
sd = W.make_decoder_state_dict(W.C512, 0)
dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=dev)
dec.max_rows_per_call = max(dec.max_rows_per_call, rows)
tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                             isotopes=[f"{300 + i}Og" for i in range(291)])
masks = tok.get_type_masks(dev)
z = W.make_latents(rows, 2048, 1234).to(dev)
st, hp = W.make_conditioning(rows, 13, 1234)
st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
mode = os.environ.get("SWEEP_MODE", "greedy")
if mode == "greedy":
    kw = dict(temperature=0.001, max_len=64, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
elif mode == "full":          # no masks / stop head: all 63 steps run
    kw = dict(temperature=0.001, max_len=64)
else:                         # RLOO-style sampling with log-probs and entropy
    kw = dict(temperature=1.2, max_len=64, stop_boost=10.0, return_log_probs=True, return_entropy=True, _seed=7)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
defaults = dict(attn_ctas_per_sm=0, gemm_stages=0, subbatches=0, graph=1, sub_min_rows=2048, attn_bulk=0, attn_bulk_min_rows=256, attn_bulk_piece_kb=0)
first = None
for spec in specs:
    cfg = dict(defaults)
    for kv in spec.split(","):
        if kv:
            k, v = kv.split("=")
            cfg[k] = int(v)
    _lib.tune(**cfg)
    for _ in range(2):
        t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
    ms = []
    for _ in range(3):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    same = True if first is None else bool(torch.equal(first, t))
    if first is None:
        first = t.clone()
    best = min(ms)
    print(f"{spec:60s} steps={t.shape[1]:3d} ms={best:8.2f} (all {[round(m, 2) for m in ms]})  "
          f"{rows / best:8.2f} K formulas/s  same_tokens={same}", flush=True)
