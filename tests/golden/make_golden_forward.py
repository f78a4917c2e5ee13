"""Golden vectors for the teacher-forced forward: run the REFERENCE EnhancedTransformerDecoder.forward
(models/autoregressive_decoder.py:901-985, teacher_forcing_ratio = 1.0) on the seeded synthetic weights.
Build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_forward.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG      # noqa: E402  (imports the reference modules, defines ref_decoder)
import torch                  # noqa: E402
from oracle import weights as W   # noqa: E402
from oracle import vocab as OV    # noqa: E402


def case(shape, B, masks, max_len):
    dec, _ = MG.ref_decoder(shape)
    z = W.make_latents(B, shape.latent_dim, 4321)
    stoich, heads = W.make_conditioning(B, shape.stoich_input_dim, 4321)
    with torch.no_grad():
        t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.001, max_len=max_len, heads_pred=heads,
                                             type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
        # targets: START + generated ids, PAD after each row's first END (exercises tgt_key_padding_mask)
        tgt = torch.cat([torch.ones((B, 1), dtype=torch.long), t], dim=1)
        for r in range(B):
            e = (tgt[r] == 2).nonzero()
            if e.numel():
                tgt[r, int(e[0]) + 1:] = 0
        tgt[0, 3:] = tgt[0, 3:].clone()          # row 0 untouched
        if tgt.shape[1] > 6:
            tgt[1, 4:] = 0                       # a row that is mostly padding
        logits, generated, stop_logits, type_logits, dup_logits = dec(z, tgt, stoich_pred=stoich, heads_pred=heads)
    return {"target_tokens": tgt, "logits": logits.float(), "generated": generated.to(torch.int16),
            "stop_logits": stop_logits.float(), "type_logits": type_logits.float(), "site_dup_logits": dup_logits.float(),
            "B": B, "seed_in": 4321, "shape": shape.as_dict()}


if __name__ == "__main__":
    out = {"meta": MG.META,
           "tiny": case(W.TINY, 5, OV.type_masks(**OV.TINY_LAYOUT), W.TINY.max_len),
           "c512": case(W.C512, 3, OV.type_masks(), 20)}
    torch.save(out, os.path.join(HERE, "forward_tf.pt"))
    for k in ("tiny", "c512"):
        print(k, {n: (tuple(v.shape) if torch.is_tensor(v) else v) for n, v in out[k].items() if n != "shape"},
              "nan:", bool(torch.isnan(out[k]["logits"]).any()))
