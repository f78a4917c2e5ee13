"""Probe (not a test): what the opt-in retirement of finished rows (decoder.compact_finished) buys on a length distribution
like a trained model's (mean formula length ~9 tokens, scripts/train_v12_clean.py:897).  The END token of every row is
forced at a position drawn from a geometric-like distribution with that mean; everything before it is decoded greedily.
usage: python tests/compact_bench.py [rows] [mean_len]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import synthetic as W      # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mean_len = float(sys.argv[2]) if len(sys.argv) > 2 else 9.0
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
dec.max_rows_per_call = max(dec.max_rows_per_call, rows)
z = W.make_latents(rows, 2048, 1234).to(dev)
st, hp = W.make_conditioning(rows, 13, 1234)
st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
g = torch.Generator().manual_seed(5)
lens = torch.clamp((torch.empty(rows).exponential_(1.0 / (mean_len - 2.0), generator=g) + 3.0).round().long(), 3, 63)
forced = torch.full((rows, 63), -1, dtype=torch.int64)
forced[torch.arange(rows), lens - 1] = 2                                      # END at position len - 1
forced = forced.to(dev)
kw = dict(temperature=0.001, max_len=64, _forced_tokens=forced)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run():
    for _ in range(2):
        t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
    best = 1e9
    for _ in range(3):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return t, best


t0, ms0 = run()
dec.compact_finished = True
t1, ms1 = run()
L = t0.shape[1]
is_end = t0 == 2
after = (torch.cumsum(is_end.int(), dim=1) - is_end.int()) > 0
same = t0.shape == t1.shape and bool(torch.equal(t1[~after], t0[~after]))
row_steps = int((~after).sum())
print(f"rows={rows} forced lengths: mean {float(lens.float().mean()):.2f}, max {int(lens.max())}; executed steps {L}; "
      f"row-steps needed {row_steps} of {rows * L} ({100.0 * row_steps / (rows * L):.1f} %)")
print(f"reference semantics (finished rows keep decoding): {ms0:8.2f} ms = {rows / ms0:7.2f} K formulas/s")
print(f"compact_finished:                                  {ms1:8.2f} ms = {rows / ms1:7.2f} K formulas/s   "
      f"x{ms0 / ms1:.2f}, identical up to END: {same}")
