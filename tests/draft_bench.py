"""Probe (not a test): latency of greedy decoding by draft verification (decoder.generate_with_draft)
against the step-by-step KV-cache decode at small batch, for drafts of varying quality.
usage: python tests/draft_bench.py [rows]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superconductor_vae_b200 as S                     # noqa: E402
from superconductor_vae_b200 import synthetic as W      # noqa: E402
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer      # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
dec = S.EnhancedTransformerDecoder.from_state_dict(W.make_decoder_state_dict(W.C512, 0), nhead=8, device=dev)
tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                             isotopes=[f"{300 + i}Og" for i in range(291)])
masks = tok.get_type_masks(dev)
z = W.make_latents(rows, 2048, 1234).to(dev)
st, hp = W.make_conditioning(rows, 13, 1234)
st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
kw = dict(stoich_pred=st, max_len=64, heads_pred=hp, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, 1e3 * (time.perf_counter() - t0) / reps


(ref, _, _), ms_ref = timed(lambda: dec.generate_with_kv_cache(z, temperature=0.001, **kw))
print(f"rows={rows}: step-by-step KV-cache decode: {ref.shape[1]} steps, {ms_ref:7.2f} ms")
gen = torch.Generator().manual_seed(1)
for name, n_bad in (("exact draft", 0), ("1 wrong token per row", 1), ("3 wrong tokens per row", 3)):
    d = ref.clone().cpu()
    for r in range(rows):
        for p in torch.randint(0, d.shape[1], (n_bad,), generator=gen).tolist():
            d[r, p] = int(torch.randint(3, 4752, (1,), generator=gen))
    d = d.to(dev)
    (t, passes, fb), ms = timed(lambda: dec.generate_with_draft(z, d, **kw))
    ok = bool(torch.equal(t, ref[:, :t.shape[1]].masked_fill(((torch.cumsum((ref == 2).int(), 1) - (ref == 2).int()) > 0)[:, :t.shape[1]], 0)))
    print(f"   draft verification, {name:24s}: {passes} passes, {fb} fallback rows, {ms:7.2f} ms  (x{ms_ref / ms:.1f}), same tokens: {ok}")
# neighbouring latents of a SLERP walk: the previous point's decode is the draft of the next
za, zb = W.make_latents(rows, 2048, 7).to(dev), W.make_latents(rows, 2048, 8).to(dev)
prev = None
for t_ in (0.50, 0.51, 0.52, 0.55):
    zt = S.latent.slerp(za, zb, t_)
    (r_, _, _), ms_r = timed(lambda: dec.generate_with_kv_cache(zt, temperature=0.001, **kw), reps=2)
    if prev is not None:
        (t, passes, fb), ms = timed(lambda: dec.generate_with_draft(zt, prev, **kw), reps=2)
        print(f"   walk t={t_:.2f}: draft = decode at the previous t: {passes} passes, {fb} fallback rows, {ms:7.2f} ms vs {ms_r:7.2f} ms")
    prev = r_
