"""Generate golden vectors by running the REFERENCE modules (imported from /root/reference/src)
on the oracle's seeded synthetic state_dicts.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs small fixtures under tests/golden/ (committed).  torch version and seeds are recorded in
each file.  The reference has no tests of its own for this path, so these
vectors are what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA engine.
"""
import io
import os
import sys
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

import torch

with contextlib.redirect_stdout(io.StringIO()):
    from superconductor.models.attention_vae import FullMaterialsVAE
    from superconductor.models.autoregressive_decoder import EnhancedTransformerDecoder
    from superconductor.tokenizer.fraction_tokenizer import FractionAwareTokenizer

from oracle import weights as W
from oracle import vocab as OV

torch.set_num_threads(8)
META = {"torch": torch.__version__}


def ref_decoder(shape: W.DecoderShape, seed=0):
    dec = EnhancedTransformerDecoder(
        latent_dim=shape.latent_dim, d_model=shape.d_model, nhead=shape.nhead, num_layers=shape.num_layers,
        dim_feedforward=shape.dim_feedforward, max_len=shape.max_len, n_memory_tokens=shape.n_memory_tokens,
        encoder_skip_dim=shape.encoder_skip_dim, use_skip_connection=shape.use_skip_connection,
        vocab_size=shape.vocab_size, stoich_input_dim=shape.stoich_input_dim,
        memory_bottleneck_dim=shape.memory_bottleneck_dim)
    sd = W.make_decoder_state_dict(shape, seed)
    dec.load_state_dict(sd, strict=True)
    return dec.eval(), sd


def ref_encoder(seed=1):
    enc = FullMaterialsVAE()
    sd = W.make_encoder_state_dict(W.ENC_DEFAULT, seed)
    enc.load_state_dict(sd, strict=True)
    return enc.eval(), sd


def i16(t):
    return t.to(torch.int16)


def decoder_cases(name, shape, B, masks, seed_in=1234, steps_logits=(0, 1, 5), rows_logits=4, extra=None):
    dec, _ = ref_decoder(shape)
    z = W.make_latents(B, shape.latent_dim, seed_in)
    stoich, heads = W.make_conditioning(B, shape.stoich_input_dim, seed_in)
    out = {"meta": dict(META, shape=shape.as_dict(), B=B, seed_in=seed_in)}
    with torch.no_grad():
        mem24 = dec.precompute_memory(z, None, stoich, heads)
        mem20 = dec.precompute_memory(z, None, stoich, None)
        mem16 = dec.precompute_memory(z, None, None, None)
    out["memory24_rows"] = mem24[:2].clone()
    out["memory_shapes"] = torch.tensor([mem24.shape[1], mem20.shape[1], mem16.shape[1]])
    out["memory24_sum"] = mem24.double().sum(dim=(1, 2)).float()

    # capture per-step raw logits through a forward hook on output_proj
    cap = []
    hook = dec.output_proj.register_forward_hook(lambda m, i, o: cap.append(o.squeeze(1)[:rows_logits].clone()))
    # (a) greedy, masks + stop head + hard stop (train defaults)
    t, lp, en = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.001, max_len=shape.max_len,
                                           heads_pred=heads, type_masks=masks, stop_boost=10.0,
                                           hard_stop_threshold=0.8, return_log_probs=True, return_entropy=True)
    out["greedy_masked_tokens"] = i16(t)
    out["greedy_masked_entropy_row0"] = en[0].clone()
    out["greedy_masked_logprob_absmax"] = lp.abs().max()
    cap.clear()
    # (b) greedy, no masks, no stop -> all max_len-1 steps run
    t, _, en = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.001, max_len=shape.max_len,
                                          heads_pred=heads, return_entropy=True)
    out["greedy_plain_tokens"] = i16(t)
    out["greedy_plain_entropy"] = en[:rows_logits].clone()
    out["greedy_plain_logits"] = torch.stack([cap[s] for s in steps_logits if s < len(cap)])
    out["greedy_plain_logit_steps"] = torch.tensor([s for s in steps_logits if s < len(cap)])
    cap.clear()
    # (c) greedy with stop boost only (additive, no -inf)
    t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.001, max_len=shape.max_len,
                                         heads_pred=heads, stop_boost=10.0)
    out["greedy_stopboost_tokens"] = i16(t)
    # (d) 20-token memory (holdout scripts: no heads, no masks)
    t, _, _ = dec.generate_with_kv_cache(z=z, stoich_pred=stoich, temperature=0.001)
    out["greedy_m20_tokens"] = i16(t)
    # (e) H1: temperature=0.0 compat semantics
    t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.0, max_len=shape.max_len,
                                         heads_pred=heads, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
    out["t0_masked_tokens"] = i16(t)
    t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.0, max_len=min(shape.max_len, 8),
                                         heads_pred=heads)
    out["t0_plain_tokens"] = i16(t)
    cap.clear()
    # (f) sampling with the global generator, no masks (distribution = softmax(logits/T))
    torch.manual_seed(77)
    t, lp, en, mk = dec.sample_for_reinforce(z, stoich_pred=stoich, temperature=1.2, max_len=shape.max_len,
                                             stop_boost=10.0, heads_pred=heads)
    out["sample_tokens"], out["sample_logprobs"], out["sample_entropy"], out["sample_mask"] = i16(t), lp, en, mk
    # (g) H2: sampling with masks -> uniform fallback
    torch.manual_seed(78)
    t, lp, en, mk = dec.sample_for_reinforce(z, stoich_pred=stoich, temperature=1.2, max_len=shape.max_len,
                                             stop_boost=10.0, hard_stop_threshold=0.8, heads_pred=heads,
                                             type_masks=masks)
    out["sample_masked_tokens"], out["sample_masked_logprobs"] = i16(t), lp
    out["sample_masked_entropy"], out["sample_masked_mask"] = en, mk
    # (h) site-duplication gating (:1424-1435, :1525-1539)
    t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.001, max_len=shape.max_len,
                                         heads_pred=heads, site_dup_threshold=0.6)
    out["sitedup_plain_tokens"] = i16(t)
    t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.001, max_len=shape.max_len,
                                         heads_pred=heads, site_dup_threshold=0.99, type_masks=masks, stop_boost=10.0)
    out["sitedup_masked_tokens"] = i16(t)
    if extra:
        extra(dec, out, z, stoich, heads)
    hook.remove()
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print(name, {k: (tuple(v.shape) if torch.is_tensor(v) else "meta") for k, v in out.items()})


def topk_topp_extra(dec, out, z, stoich, heads):
    torch.manual_seed(5)
    t, lp, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.9, top_k=7, heads_pred=heads,
                                          return_log_probs=True)
    out["topk_tokens"], out["topk_logprobs"] = i16(t), lp
    torch.manual_seed(6)
    t, lp, _ = dec.generate_with_kv_cache(z, stoich_pred=stoich, temperature=0.9, top_p=0.8, heads_pred=heads,
                                          return_log_probs=True)
    out["topp_tokens"], out["topp_logprobs"] = i16(t), lp


def skip_case():
    shape = W.TINY_SKIP
    dec, _ = ref_decoder(shape)
    B = 5
    z = W.make_latents(B, shape.latent_dim, 99)
    stoich, heads = W.make_conditioning(B, shape.stoich_input_dim, 99)
    skip = torch.randn((B, shape.encoder_skip_dim), generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        mem = dec.precompute_memory(z, skip, stoich, heads)
    t, _, _ = dec.generate_with_kv_cache(z, encoder_skip=skip, stoich_pred=stoich, temperature=0.001, heads_pred=heads)
    torch.save({"meta": dict(META, shape=shape.as_dict(), B=B), "skip": skip, "memory": mem, "tokens": i16(t)},
               os.path.join(HERE, "tiny_skip.pt"))
    print("tiny_skip", tuple(mem.shape), tuple(t.shape))


def encoder_case():
    enc, _ = ref_encoder()
    B = 8
    idx, frac, mask, magpie, tc = W.make_compositions(B, 4321)
    with torch.no_grad():
        out = enc(idx, frac, mask, magpie, tc)
    keep = ("z", "attention_weights", "tc_pred", "magpie_pred", "attended_input", "competence", "fraction_pred",
            "element_count_pred", "hp_pred", "sc_pred", "tc_class_logits", "family_coarse_logits",
            "family_cuprate_sub_logits", "family_iron_sub_logits", "family_composed_14")
    g = {k: out[k].clone() for k in keep}
    g["fused_repr"] = enc.encode(idx, frac, mask, magpie, tc)["fused_repr"].detach().clone()
    g["meta"] = dict(META, B=B, seed_in=4321)
    torch.save(g, os.path.join(HERE, "encoder_default.pt"))
    print("encoder_default", {k: tuple(v.shape) for k, v in g.items() if torch.is_tensor(v)})


def slerp_case():
    from types import SimpleNamespace
    import importlib.util
    spec = importlib.util.spec_from_file_location("holdout_search", "/root/reference/scripts/holdout/holdout_search.py")
    src = open("/root/reference/scripts/holdout/holdout_search.py").read()
    # pull only the slerp function out of the script (the script has heavy top-level side effects)
    start = src.index("def slerp(")
    end = src.index("\ndef ", start + 1)
    ns = {"torch": torch, "F": torch.nn.functional}
    exec(src[start:end], ns)
    g = torch.Generator().manual_seed(11)
    z1, z2 = torch.randn((6, 64), generator=g), torch.randn((6, 64), generator=g)
    t = torch.rand((6, 1), generator=g)
    out = {"z1": z1, "z2": z2, "t": t, "slerp": ns["slerp"](z1, z2, t),
           "slerp_parallel": ns["slerp"](z1, z1 * 2.0, 0.25), "meta": META}
    torch.save(out, os.path.join(HERE, "slerp.pt"))
    print("slerp ok")


def tokenizer_case():
    tok = FractionAwareTokenizer("/root/reference/data/fraction_vocab.json", max_len=64,
                                 isotope_vocab_path="/root/reference/data/isotope_vocab.json")
    m = tok.get_type_masks()
    assert torch.equal(m, OV.type_masks(tok.n_fraction_tokens, tok.n_isotope_tokens)), "oracle layout mismatch"
    ids = [[1, 5 + 28, 123 + 1, 5 + 7, 123 + 6, 2, 0, 0], [5, 143, 143 + 4316, 4460, 4461, 4751, 3, 4, 2]]
    out = {"vocab_size": tok.vocab_size, "n_fractions": tok.n_fraction_tokens, "n_isotopes": tok.n_isotope_tokens,
           "mask_row_sums": m.sum(dim=1), "ids": ids, "decoded": [tok.decode(i) for i in ids],
           "names": {i: tok.get_token_name(i) for i in (0, 1, 2, 3, 4, 5, 122, 123, 142, 143, 4459, 4460, 4461, 4751)},
           "encoded_YBCO": tok.encode("YBa2Cu3O7"), "meta": META}
    from superconductor.models import autoregressive_decoder as AD      # legacy 148-token vocabulary (generate_formulas_fast)
    out["legacy_vocab"] = list(AD.VOCAB)
    out["legacy_decoded"] = AD.indices_to_formula(torch.tensor([1, 58, 141, 48, 140, 2, 5]))
    torch.save(out, os.path.join(HERE, "tokenizer.pt"))
    print("tokenizer", out["vocab_size"], out["mask_row_sums"].tolist(), out["decoded"], out["encoded_YBCO"][:12])
    return m


if __name__ == "__main__":
    masks_full = tokenizer_case()
    tiny_masks = OV.type_masks(**OV.TINY_LAYOUT)
    decoder_cases("tiny", W.TINY, 6, tiny_masks, steps_logits=(0, 1, 5), rows_logits=6, extra=topk_topp_extra)
    skip_case()
    encoder_case()
    slerp_case()
    decoder_cases("c512_b32", W.C512, 32, masks_full, steps_logits=(0, 1, 31, 62))
    decoder_cases("c512b_b4", W.C512B, 4, masks_full, steps_logits=(0, 7))
    decoder_cases("c576_b4", W.C576, 4, masks_full, steps_logits=(0, 7))
