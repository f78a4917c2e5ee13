"""Golden vectors for the scheduled-sampling forward (teacher_forcing_ratio < 1): the REFERENCE
EnhancedTransformerDecoder.forward (models/autoregressive_decoder.py:987-1082, two passes) on the seeded synthetic weights,
for the plain ratio and for use_position_dependent_tf.  The reference draws its keep-ground-truth mask with
torch.rand(B, L) after torch.manual_seed(seed); the same call on the CPU reproduces the mask, which the tests hand to the
engine and to the oracle.  Build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_forward_ss.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG      # noqa: E402  (imports the reference modules, defines ref_decoder)
import torch                  # noqa: E402
from oracle import weights as W   # noqa: E402


def case(shape, B, ratio, positional, seed):
    dec, _ = MG.ref_decoder(shape)
    dec.use_position_dependent_tf = positional
    dec.tf_position_decay = 0.5
    g = torch.load(os.path.join(HERE, "forward_tf.pt"), weights_only=False)["tiny" if shape is W.TINY else "c512"]
    tgt = g["target_tokens"][:B]
    z = W.make_latents(g["B"], shape.latent_dim, g["seed_in"])[:B]
    stoich, heads = W.make_conditioning(g["B"], shape.stoich_input_dim, g["seed_in"])
    stoich, heads = stoich[:B], {k: v[:B] for k, v in heads.items()}
    with torch.no_grad():
        torch.manual_seed(seed)
        logits, generated, stop_logits, type_logits, dup_logits = dec(z, tgt, stoich_pred=stoich, heads_pred=heads,
                                                                     teacher_forcing_ratio=ratio)
    return {"B": B, "ratio": ratio, "positional": positional, "decay": 0.5, "seed": seed, "logits": logits.float(),
            "generated": generated.to(torch.int16), "stop_logits": stop_logits.float(), "type_logits": type_logits.float(),
            "site_dup_logits": dup_logits.float()}


if __name__ == "__main__":
    out = {"meta": MG.META,
           "tiny": [case(W.TINY, 5, 0.5, False, 11), case(W.TINY, 5, 0.3, True, 12), case(W.TINY, 5, 0.0, False, 13)],
           "c512": [case(W.C512, 3, 0.5, False, 21), case(W.C512, 3, 0.4, True, 22)]}
    torch.save(out, os.path.join(HERE, "forward_ss.pt"))
    for k in ("tiny", "c512"):
        for c in out[k]:
            print(k, c["ratio"], c["positional"], tuple(c["logits"].shape), "nan:", bool(torch.isnan(c["logits"]).any()))
