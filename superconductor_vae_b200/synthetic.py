"""Seeded synthetic state_dicts with the reference's key names and shapes, and synthetic inputs.

Used by the tests, the oracle (re-exported as ``oracle.weights``) and bench.py.  The reference ships no
checkpoint (SURVEY.md "Facts"), so every parity case and benchmark runs on synthetic weights.
They are drawn here from an explicit ``torch.Generator`` so that the build
container (where the real reference modules load them with ``strict=True``)
and the GPU box (where only this oracle exists) see identical tensors.

Key names / shapes follow:
  decoder: src/superconductor/models/autoregressive_decoder.py:617-765
  encoder: src/superconductor/models/attention_vae.py:375-606,
           src/superconductor/encoders/element_attention.py:61-71,137-147
Every floating *parameter* is rounded to bf16 and back so
that an engine holding bf16 weights and the fp32 oracle share exact values; the
sinusoidal ``pos_encoding.pe`` buffer stays fp32 (it is a buffer, not a weight).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List

import torch


@dataclass(frozen=True)
class DecoderShape:
    latent_dim: int = 2048
    d_model: int = 512
    nhead: int = 8
    num_layers: int = 12
    dim_feedforward: int = 2048
    max_len: int = 64
    vocab_size: int = 4752
    n_memory_tokens: int = 16
    stoich_input_dim: int = 13
    n_stoich_tokens: int = 4
    heads_input_dim: int = 24
    heads_n_tokens: int = 4
    memory_bottleneck_dim: int = 0
    use_skip_connection: bool = False
    encoder_skip_dim: int = 256
    skip_n_tokens: int = 8

    def as_dict(self):
        return asdict(self)


@dataclass(frozen=True)
class EncoderShape:
    n_elements: int = 118
    element_embed_dim: int = 128
    n_attention_heads: int = 8
    magpie_dim: int = 145
    fusion_dim: int = 256
    encoder_hidden: tuple = (512, 256)
    latent_dim: int = 2048
    decoder_hidden: tuple = (256, 512)
    max_elements: int = 12


# SURVEY.md section 8 "Configurations"
C512 = DecoderShape()                                           # "V14.3" per README
C512B = DecoderShape(memory_bottleneck_dim=1024)                # class defaults (V15 bottleneck)
C576 = DecoderShape(d_model=576, dim_feedforward=2304)          # script MODEL_CONFIG, head_dim 72
TINY = DecoderShape(latent_dim=64, d_model=64, nhead=4, num_layers=2, dim_feedforward=128,
                    max_len=12, vocab_size=97, n_memory_tokens=4)
TINY_SKIP = DecoderShape(latent_dim=64, d_model=64, nhead=4, num_layers=2, dim_feedforward=128,
                         max_len=12, vocab_size=97, n_memory_tokens=4, use_skip_connection=True,
                         encoder_skip_dim=32)
ENC_DEFAULT = EncoderShape()


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class _Maker:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)
        self.sd: Dict[str, torch.Tensor] = {}

    def _u(self, shape, bound):
        return (torch.rand(shape, generator=self.g, dtype=torch.float32) * 2.0 - 1.0) * bound

    def matrix(self, name, out_f, in_f, gain=1.0):
        bound = gain * math.sqrt(6.0 / (in_f + out_f))       # xavier-uniform scale
        self.sd[name] = _bf16_round(self._u((out_f, in_f), bound))

    def vector(self, name, n, bound):
        self.sd[name] = _bf16_round(self._u((n,), bound))

    def linear(self, prefix, out_f, in_f, gain=1.0):
        self.matrix(prefix + ".weight", out_f, in_f, gain)
        self.vector(prefix + ".bias", out_f, 1.0 / math.sqrt(in_f))

    def layernorm(self, prefix, n):
        self.sd[prefix + ".weight"] = _bf16_round(1.0 + self._u((n,), 0.1))
        self.sd[prefix + ".bias"] = _bf16_round(self._u((n,), 0.05))


def sinusoidal_pe(max_len: int, d_model: int) -> torch.Tensor:
    """pos_encoding.pe buffer, [1, max_len, d_model] fp32 (autoregressive_decoder.py:399-408)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def make_decoder_state_dict(shape: DecoderShape = C512, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic EnhancedTransformerDecoder.state_dict() (same keys, same shapes)."""
    s = shape
    d = s.d_model
    m = _Maker(seed)
    m.matrix("token_embedding.weight", s.vocab_size, d)
    m.sd["pos_encoding.pe"] = sinusoidal_pe(s.max_len, d)
    n_lat = d * s.n_memory_tokens
    if s.memory_bottleneck_dim > 0:
        m.linear("latent_to_memory.0", s.memory_bottleneck_dim, s.latent_dim)
        m.layernorm("latent_to_memory.1", s.memory_bottleneck_dim)
        m.linear("latent_to_memory.3", n_lat, s.memory_bottleneck_dim)
    else:
        m.linear("latent_to_memory.0", n_lat // 2, s.latent_dim)
        m.linear("latent_to_memory.2", n_lat, n_lat // 2)
    if s.use_skip_connection:
        n_skip = d * s.skip_n_tokens
        m.linear("skip_to_memory.0", n_skip // 2, s.encoder_skip_dim)
        m.linear("skip_to_memory.2", n_skip, n_skip // 2)
    m.linear("stoich_to_memory.0", d, s.stoich_input_dim)
    m.layernorm("stoich_to_memory.1", d)
    m.linear("stoich_to_memory.3", d * s.n_stoich_tokens, d)
    for i in range(s.num_layers):
        p = f"transformer_decoder.layers.{i}."
        m.matrix(p + "self_attn.in_proj_weight", 3 * d, d)
        m.vector(p + "self_attn.in_proj_bias", 3 * d, 0.05)
        m.linear(p + "self_attn.out_proj", d, d)
        m.matrix(p + "multihead_attn.in_proj_weight", 3 * d, d)
        m.vector(p + "multihead_attn.in_proj_bias", 3 * d, 0.05)
        m.linear(p + "multihead_attn.out_proj", d, d)
        m.linear(p + "linear1", s.dim_feedforward, d)
        m.linear(p + "linear2", d, s.dim_feedforward)
        for k in (1, 2, 3):
            m.layernorm(p + f"norm{k}", d)
    m.layernorm("output_proj.0", d)
    m.linear("output_proj.1", d, d)
    m.linear("output_proj.4", s.vocab_size, d)
    m.linear("stop_head.0", d // 4, d)
    m.linear("stop_head.2", 1, d // 4)
    m.linear("site_dup_head.0", d // 4, d)
    m.linear("site_dup_head.2", 1, d // 4)
    m.layernorm("token_type_head.0", d)
    m.linear("token_type_head.1", d, d)
    m.linear("token_type_head.4", d // 4, d)
    m.linear("token_type_head.7", 5, d // 4)
    m.linear("heads_to_memory.0", d // 2, s.heads_input_dim)
    m.layernorm("heads_to_memory.1", d // 2)
    m.linear("heads_to_memory.3", d, d // 2)
    m.linear("heads_to_memory.5", d * s.heads_n_tokens, d)
    return m.sd


def make_encoder_state_dict(shape: EncoderShape = ENC_DEFAULT, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Synthetic FullMaterialsVAE.state_dict() (same keys, same shapes)."""
    s = shape
    e = s.element_embed_dim
    f = s.fusion_dim
    m = _Maker(seed)
    p = "element_encoder.element_embedding."
    emb = _bf16_round(torch.randn((s.n_elements + 1, e), generator=m.g))
    emb[0].zero_()                                        # padding_idx=0 row (element_attention.py:61)
    m.sd[p + "element_embed.weight"] = emb
    m.linear(p + "property_encoder.0", e, 11)             # present in the state_dict, unused (element_properties=None)
    m.layernorm(p + "property_encoder.1", e)
    m.linear(p + "combiner", e, 2 * e)
    p = "element_encoder.element_attention."
    m.matrix(p + "query", s.n_attention_heads, e // s.n_attention_heads)
    m.linear(p + "key_proj", e, e)
    m.linear(p + "value_proj", e, e)
    m.linear(p + "output_proj", e, e)
    m.layernorm(p + "layer_norm", e)
    m.linear("element_encoder.output_projection.0", f, e)
    m.layernorm("element_encoder.output_projection.1", f)
    m.linear("magpie_encoder.0", 2 * f, s.magpie_dim)
    m.layernorm("magpie_encoder.1", 2 * f)
    m.linear("magpie_encoder.4", f, 2 * f)
    m.layernorm("magpie_encoder.5", f)
    m.linear("tc_encoder.0", f // 2, 1)
    m.linear("tc_encoder.2", f, f // 2)
    m.layernorm("tc_encoder.3", f)
    m.linear("fusion.0", 3 * f, 3 * f)
    m.layernorm("fusion.1", 3 * f)
    prev = 3 * f
    for j, h in enumerate(s.encoder_hidden):
        m.linear(f"vae_encoder.encoder.{3 * j}", h, prev)
        m.layernorm(f"vae_encoder.encoder.{3 * j + 1}", h)
        prev = h
    m.linear("vae_encoder.fc_mean", s.latent_dim, prev)
    prev = s.latent_dim
    for j, h in enumerate(s.decoder_hidden):
        m.linear(f"decoder_backbone.{4 * j}", h, prev)
        m.layernorm(f"decoder_backbone.{4 * j + 1}", h)
        prev = h
    bb = prev
    m.linear("tc_proj", 256, bb)
    m.linear("tc_res_block.0", 256, 256)
    m.layernorm("tc_res_block.1", 256)
    m.linear("tc_res_block.4", 256, 256)
    m.layernorm("tc_out.0", 256)
    m.linear("tc_out.2", 128, 256)
    m.linear("tc_out.4", 1, 128)
    m.linear("magpie_head.0", bb, bb)
    m.linear("magpie_head.2", s.magpie_dim, bb)
    m.linear("attended_head.0", f, bb)
    m.layernorm("attended_head.1", f)
    m.linear("competence_head.0", s.latent_dim // 4, s.latent_dim)
    m.linear("competence_head.2", 1, s.latent_dim // 4)
    m.linear("fraction_head.0", 256, s.latent_dim)
    m.layernorm("fraction_head.1", 256)
    m.linear("fraction_head.4", 128, 256)
    m.linear("fraction_head.6", s.max_elements + 1, 128)
    m.linear("hp_head.0", 256, s.latent_dim)
    m.linear("hp_head.2", 1, 256)
    m.linear("tc_class_head.0", 256, bb)
    m.linear("tc_class_head.3", 5, 256)
    sc_in = s.latent_dim + 1 + s.magpie_dim + 1 + s.max_elements + 1 + 1 + 5
    m.linear("sc_head.0", 512, sc_in)
    m.layernorm("sc_head.2", 512)
    m.linear("sc_head.4", 128, 512)
    m.linear("sc_head.6", 1, 128)
    p = "hierarchical_family_head."
    m.linear(p + "coarse_head.0", 256, bb + 1)
    m.layernorm(p + "coarse_head.1", 256)
    m.linear(p + "coarse_head.4", 128, 256)
    m.linear(p + "coarse_head.6", 7, 128)
    m.linear(p + "cuprate_sub_head.0", 128, bb + 1)
    m.layernorm(p + "cuprate_sub_head.1", 128)
    m.linear(p + "cuprate_sub_head.4", 64, 128)
    m.linear(p + "cuprate_sub_head.6", 6, 64)
    m.linear(p + "iron_sub_head.0", 64, bb + 1)
    m.layernorm(p + "iron_sub_head.1", 64)
    m.linear(p + "iron_sub_head.4", 2, 64)
    return m.sd


# ---------------------------------------------------------------------------
# Synthetic inputs
# ---------------------------------------------------------------------------

def make_latents(n: int, latent_dim: int = 2048, seed: int = 1234, z_norm: float = 22.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n, latent_dim), generator=g) * (z_norm / math.sqrt(latent_dim))


def make_conditioning(n: int, stoich_dim: int = 13, seed: int = 1234):
    """stoich_pred [n, stoich_dim] and the heads_pred dict (train_v12_clean.py:5288-5296)."""
    g = torch.Generator().manual_seed(seed + 1)
    stoich = torch.rand((n, stoich_dim), generator=g)
    heads = {
        "tc_pred": torch.randn((n,), generator=g),
        "sc_pred": torch.randn((n,), generator=g),
        "hp_pred": torch.randn((n,), generator=g),
        "tc_class_logits": torch.randn((n, 5), generator=g),
        "competence": torch.rand((n,), generator=g),
        "element_count_pred": torch.randn((n,), generator=g),
        "family_composed_14": torch.softmax(torch.randn((n, 14), generator=g), dim=-1),
    }
    return stoich, heads


def make_compositions(n: int, seed: int = 1234, max_elements: int = 12, n_elements: int = 118,
                      magpie_dim: int = 145):
    """Synthetic encoder inputs: 1..8 distinct elements, Dirichlet(1) fractions."""
    g = torch.Generator().manual_seed(seed + 2)
    idx = torch.zeros((n, max_elements), dtype=torch.int64)
    frac = torch.zeros((n, max_elements), dtype=torch.float32)
    mask = torch.zeros((n, max_elements), dtype=torch.bool)
    n_el = torch.randint(1, 9, (n,), generator=g)
    for b in range(n):
        k = int(n_el[b])
        perm = torch.randperm(n_elements, generator=g)[:k] + 1
        e = -torch.log(torch.rand((k,), generator=g).clamp_min(1e-12))   # Dirichlet(1) via exponentials
        idx[b, :k] = perm
        frac[b, :k] = e / e.sum()
        mask[b, :k] = True
    magpie = torch.randn((n, magpie_dim), generator=g)
    tc = torch.randn((n,), generator=g)
    return idx, frac, mask, magpie, tc


# ------------------------------------------------------------------------------------------------ reward inputs (f1)
def make_fraction_values(vocab: int = 4752, fraction_token_start: int = 143, seed: int = 4242) -> torch.Tensor:
    """Per-token float value table of the semantic fraction tokens (scripts/train_v12_clean.py:2556-2564): zero below
    `fraction_token_start`, here uniform in [0, 25) so that the 20.0 clamp of the value penalty is reached."""
    g = torch.Generator().manual_seed(seed)
    fv = torch.zeros(vocab)
    fv[fraction_token_start:] = torch.rand(vocab - fraction_token_start, generator=g) * 25.0
    return fv


def make_reward_rows(n: int, seq_len: int, vocab: int = 4752, seed: int = 4242, old_vocab: bool = False):
    """Seeded (sampled, target, mask) rows for the rollout reward, built to reach every branch of
    `compute_reward_gpu_native`: exact copies, 1-6 substitutions, samples that stop early / run on after a correct
    prefix, samples without END, fully random rows, and masks both "up to the sample's END" (what
    sample_for_reinforce returns) and all-ones.  Targets are START-less formulas `tokens... END PAD...`
    (PAD=0, END=2).  old_vocab=True draws from the pre-V13 ids the digit-level penalties look at
    (parentheses 4/5, slash 16, digits 138-147)."""
    g = torch.Generator().manual_seed(seed)
    if old_vocab:
        pool = torch.tensor([4, 5, 16] * 6 + list(range(138, 148)) * 3 + list(range(20, 60)))
    else:
        pool = torch.cat([torch.arange(5, 123), torch.arange(123, 143), torch.randint(143, vocab, (160,), generator=g)])

    def draw(k):
        return pool[torch.randint(0, pool.numel(), (k,), generator=g)]

    sampled = torch.zeros((n, seq_len), dtype=torch.long)
    target = torch.zeros((n, seq_len), dtype=torch.long)
    mask = torch.zeros((n, seq_len), dtype=torch.bool)
    for b in range(n):
        tl = int(torch.randint(3, seq_len - 2, (1,), generator=g))        # formula tokens before END
        t = torch.zeros(seq_len, dtype=torch.long)
        t[:tl] = draw(tl)
        t[tl] = 2
        s = t.clone()
        kind = b % 10
        if kind in (1, 2, 3):                                              # a few substitutions
            for p in torch.randperm(tl, generator=g)[: int(torch.randint(1, 7, (1,), generator=g))]:
                s[p] = draw(1)[0]
        elif kind == 4:                                                    # correct prefix, stops early
            cut = int(torch.randint(1, tl, (1,), generator=g))
            s[cut] = 2
            s[cut + 1:] = 0
        elif kind == 5:                                                    # correct formula, runs on
            ext = int(torch.randint(1, seq_len - tl - 1, (1,), generator=g)) if seq_len - tl - 1 > 1 else 1
            s[tl:tl + ext] = draw(ext)
            if tl + ext < seq_len:
                s[tl + ext] = 2
        elif kind == 6:                                                    # never emits END
            s[:] = draw(seq_len)
            s[s == 2] = 7
        elif kind == 7:                                                    # unrelated sample
            sl = int(torch.randint(1, seq_len - 1, (1,), generator=g))
            s[:] = 0
            s[:sl] = draw(sl)
            s[sl] = 2
        elif kind == 8:                                                    # substitutions + wrong length
            for p in torch.randperm(tl, generator=g)[:2]:
                s[p] = draw(1)[0]
            s[tl] = draw(1)[0]
            if tl + 1 < seq_len:
                s[tl + 1] = 2
        # kinds 0 and 9: exact copies
        sampled[b], target[b] = s, t
        if b % 3 == 0:
            mask[b] = True
        else:                                                              # 1 up to and including the sample's first END
            e = (s == 2).nonzero()
            mask[b, : (int(e[0]) + 1 if e.numel() else seq_len)] = True
    return sampled, target, mask


# ------------------------------------------------------------------------------------------------ constraint inputs (f1)
_RULE_ELEMENTS = (3, 5, 6, 8, 9, 12, 13, 14, 20, 23, 25, 26, 27, 28, 29, 32, 38, 39, 41, 50, 56, 57, 80, 81, 82, 83)


def make_constraint_rows(n: int, seq_len: int, seed: int = 777, semantic: bool = True, vocab: int = 4752,
                         fraction_token_start: int = 143):
    """Seeded formula-like token rows `El sub El sub ... END PAD...` for the chemistry-constraint rewards, in the V13
    layout (elements 5-122, integer tokens 123-142, fraction tokens 143+) or the pre-V13 one (elements 20-137,
    digits 138-147, '(' 4, ')' 5, '/' 16).  Elements are drawn mostly from the ones the rules name (O, Cu, Sr, ...),
    so duplicates, F+Tl, Cu next to magnetic metals, reducible integer formulas and every family rule occur; some rows
    have malformed groups, no END, or an unmasked position in the middle.  Returns (tokens int64 [n, L], mask bool)."""
    g = torch.Generator().manual_seed(seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))
    e0 = 5 if semantic else 20
    tokens = torch.zeros((n, seq_len), dtype=torch.long)
    mask = torch.zeros((n, seq_len), dtype=torch.bool)
    for b in range(n):
        row = []
        n_el = ri(1, 7)
        all_int = ri(0, 3) == 0                                    # integer-only formulas reach the A4 rule
        for _ in range(n_el):
            z = _RULE_ELEMENTS[ri(0, len(_RULE_ELEMENTS))] if ri(0, 5) else ri(1, 119)
            row.append(e0 - 1 + z)
            kind = ri(0, 6)
            if semantic:
                if kind <= 1 or all_int:
                    if kind != 5:
                        row.append(123 + (2 * ri(0, 5) + 1 if all_int and ri(0, 2) else ri(0, 20)))
                elif kind <= 4:
                    row.append(ri(fraction_token_start, vocab))
            else:
                if all_int or kind <= 1:
                    if kind != 5:
                        row.extend(138 + int(c) for c in str(ri(1, 40) * (2 if all_int and ri(0, 2) else 1)))
                elif kind <= 3:
                    num, den = ri(1, 60), ri(1, 60)
                    row.append(4)
                    row.extend(138 + int(c) for c in str(num))
                    row.append(16)
                    row.extend(138 + int(c) for c in str(den))
                    if ri(0, 8):                                   # sometimes the group is left open
                        row.append(5)
                elif kind == 4:
                    row.extend([4, 138 + ri(0, 10), 138 + ri(0, 10), 5] if ri(0, 2) else [4, 16, 5])    # malformed groups
        if b % 11 != 7:
            row.append(2)                                          # END (missing in a few rows)
        row = row[:seq_len]
        tokens[b, :len(row)] = torch.tensor(row, dtype=torch.long)
        if b % 4 == 0:
            mask[b] = True                                         # PAD positions masked in (id 0 is no element)
        else:
            mask[b, :len(row)] = True
        if b % 13 == 5 and len(row) > 4:
            mask[b, ri(2, len(row))] = False                       # a hole: the scans stop there, A1 does not
    return tokens, mask


def make_family_probs(n: int, seed: int = 778) -> torch.Tensor:
    """[n, 14] composed family probabilities, about half of the rows above the 0.8 confidence threshold."""
    g = torch.Generator().manual_seed(seed)
    return torch.softmax(torch.randn((n, 14), generator=g) * 4.0, dim=1)


def make_constraint_fraction_values(vocab: int = 4752, fraction_token_start: int = 143, seed: int = 779) -> torch.Tensor:
    """Fraction-token value table for the constraint inputs: amounts in [0, 8) so that the family windows
    (0.055-0.27, 0.3, 0.7, 6.35 ...) are hit from both sides."""
    g = torch.Generator().manual_seed(seed)
    fv = torch.zeros(vocab)
    u = torch.rand(vocab - fraction_token_start, generator=g)
    fv[fraction_token_start:] = torch.where(u < 0.5, u * 1.0, (u - 0.5) * 16.0)
    return fv
