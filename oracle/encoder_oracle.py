"""fp32 CPU restatement of FullMaterialsVAE encode + the heads that feed decoder memory.

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by the product.

Functional over a plain ``state_dict``; follows
  encode          src/superconductor/models/attention_vae.py:625-676
                  (ElementEncoder :87-133, ElementEmbedding element_attention.py:73-98,
                   ElementAttention element_attention.py:152-214, AttentionVAEEncoder :164-168)
  heads_from_latent  attention_vae.py:678-709 (decode) + :733-770 (head section of forward)
                  + HierarchicalFamilyHead.forward :236-307
  conditioning    scripts/train_v12_clean.py:5245-5296 (stoich_pred / heads_pred assembly)
Pinned by tests/golden/make_golden.py against the reference module itself.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(sd, prefix, x):
    w = sd[prefix + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[prefix + ".bias"], 1e-5)


def encode(sd, element_indices, element_fractions, element_mask, magpie_features, tc) -> Dict[str, torch.Tensor]:
    if tc.dim() == 1:
        tc = tc.unsqueeze(-1)
    B, E = element_indices.shape
    # --- element branch
    emb = F.embedding(element_indices, sd["element_encoder.element_embedding.element_embed.weight"])
    weighted = emb * element_fractions.unsqueeze(-1)
    pa = "element_encoder.element_attention."
    query = sd[pa + "query"]                                    # [heads, head_dim]
    nh, hd = query.shape
    keys = _lin(sd, pa + "key_proj", weighted).view(B, E, nh, hd).permute(0, 2, 1, 3)
    vals = _lin(sd, pa + "value_proj", weighted).view(B, E, nh, hd).permute(0, 2, 1, 3)
    scores = torch.matmul(query.unsqueeze(0).unsqueeze(2), keys.transpose(-2, -1)) / (hd ** 0.5 * 1.0)
    scores = scores.masked_fill(~element_mask.bool().unsqueeze(1).unsqueeze(2), float("-inf"))
    attn = F.softmax(scores, dim=-1)                            # [B, heads, 1, E]
    attended = torch.matmul(attn, vals).squeeze(2).reshape(B, nh * hd)
    attended = _ln(sd, pa + "layer_norm", _lin(sd, pa + "output_proj", attended))
    element_repr = F.gelu(_ln(sd, "element_encoder.output_projection.1",
                              _lin(sd, "element_encoder.output_projection.0", attended)))
    # --- magpie branch
    h = F.gelu(_ln(sd, "magpie_encoder.1", _lin(sd, "magpie_encoder.0", magpie_features)))
    magpie_repr = F.gelu(_ln(sd, "magpie_encoder.5", _lin(sd, "magpie_encoder.4", h)))
    # --- tc branch
    h = F.gelu(_lin(sd, "tc_encoder.0", tc))
    tc_repr = F.gelu(_ln(sd, "tc_encoder.3", _lin(sd, "tc_encoder.2", h)))
    fused = torch.cat([element_repr, magpie_repr, tc_repr], dim=-1)
    fused = F.gelu(_ln(sd, "fusion.1", _lin(sd, "fusion.0", fused)))
    h = fused
    j = 0
    while f"vae_encoder.encoder.{3 * j}.weight" in sd:
        h = F.gelu(_ln(sd, f"vae_encoder.encoder.{3 * j + 1}", _lin(sd, f"vae_encoder.encoder.{3 * j}", h)))
        j += 1
    z = _lin(sd, "vae_encoder.fc_mean", h)
    return {"z": z, "z_mean": z, "z_logvar": None, "attention_weights": attn.mean(dim=1).squeeze(1),
            "element_embeddings": emb, "fused_repr": fused}


def heads_from_latent(sd, z) -> Dict[str, torch.Tensor]:
    h = z
    j = 0
    while f"decoder_backbone.{4 * j}.weight" in sd:
        h = F.gelu(_ln(sd, f"decoder_backbone.{4 * j + 1}", _lin(sd, f"decoder_backbone.{4 * j}", h)))
        j += 1
    tc_h = _lin(sd, "tc_proj", h)
    r = _lin(sd, "tc_res_block.4", F.gelu(_ln(sd, "tc_res_block.1", _lin(sd, "tc_res_block.0", tc_h))))
    tc_h = tc_h + r
    tc_pred = _lin(sd, "tc_out.4", F.gelu(_lin(sd, "tc_out.2", F.gelu(_ln(sd, "tc_out.0", tc_h))))).squeeze(-1)
    magpie_pred = _lin(sd, "magpie_head.2", F.gelu(_lin(sd, "magpie_head.0", h)))
    attended_input = _ln(sd, "attended_head.1", _lin(sd, "attended_head.0", h))
    tc_class_logits = _lin(sd, "tc_class_head.3", F.gelu(_lin(sd, "tc_class_head.0", h)))
    competence = torch.sigmoid(_lin(sd, "competence_head.2", F.gelu(_lin(sd, "competence_head.0", z)))).squeeze(-1)
    fr = F.gelu(_ln(sd, "fraction_head.1", _lin(sd, "fraction_head.0", z)))
    fraction_output = _lin(sd, "fraction_head.6", F.gelu(_lin(sd, "fraction_head.4", fr)))
    fraction_pred, element_count_pred = fraction_output[:, :-1], fraction_output[:, -1]
    hp_pred = _lin(sd, "hp_head.2", F.relu(_lin(sd, "hp_head.0", z))).squeeze(-1)
    sc_in = torch.cat([z, tc_pred.unsqueeze(-1), magpie_pred, hp_pred.unsqueeze(-1), fraction_pred,
                       element_count_pred.unsqueeze(-1), competence.unsqueeze(-1), tc_class_logits], dim=-1)
    s = _ln(sd, "sc_head.2", F.gelu(_lin(sd, "sc_head.0", sc_in)))
    sc_pred = _lin(sd, "sc_head.6", F.gelu(_lin(sd, "sc_head.4", s))).squeeze(-1)
    # hierarchical family head
    pf = "hierarchical_family_head."
    sc_prob = torch.sigmoid(sc_pred).unsqueeze(-1)
    cond = torch.cat([h, sc_prob], dim=-1)
    c = F.gelu(_ln(sd, pf + "coarse_head.1", _lin(sd, pf + "coarse_head.0", cond)))
    coarse = _lin(sd, pf + "coarse_head.6", F.gelu(_lin(sd, pf + "coarse_head.4", c)))
    c = F.gelu(_ln(sd, pf + "cuprate_sub_head.1", _lin(sd, pf + "cuprate_sub_head.0", cond)))
    cup = _lin(sd, pf + "cuprate_sub_head.6", F.gelu(_lin(sd, pf + "cuprate_sub_head.4", c)))
    c = F.gelu(_ln(sd, pf + "iron_sub_head.1", _lin(sd, pf + "iron_sub_head.0", cond)))
    iron = _lin(sd, pf + "iron_sub_head.4", c)
    cp, up, ip = F.softmax(coarse, -1), F.softmax(cup, -1), F.softmax(iron, -1)
    sp = sc_prob.squeeze(-1)
    comp = torch.zeros(z.shape[0], 14)
    comp[:, 0] = 1.0 - sp
    comp[:, 1] = sp * cp[:, 0]
    comp[:, 2:8] = (sp * cp[:, 1]).unsqueeze(-1) * up
    comp[:, 8:10] = (sp * cp[:, 2]).unsqueeze(-1) * ip
    comp[:, 10] = sp * cp[:, 3]
    comp[:, 11] = sp * cp[:, 4]
    comp[:, 12] = sp * cp[:, 5]
    comp[:, 13] = sp * cp[:, 6]
    return {"tc_pred": tc_pred, "magpie_pred": magpie_pred, "attended_input": attended_input,
            "tc_class_logits": tc_class_logits, "competence": competence, "fraction_pred": fraction_pred,
            "element_count_pred": element_count_pred, "hp_pred": hp_pred, "sc_pred": sc_pred,
            "family_coarse_logits": coarse, "family_cuprate_sub_logits": cup, "family_iron_sub_logits": iron,
            "family_composed_14": comp, "backbone_h": h}


def forward(sd, element_indices, element_fractions, element_mask, magpie_features, tc):
    out = encode(sd, element_indices, element_fractions, element_mask, magpie_features, tc)
    out.update(heads_from_latent(sd, out["z"]))
    return out


def conditioning(out: Dict[str, torch.Tensor]):
    """stoich_pred [B, 13] and heads_pred as the training script assembles them."""
    stoich = torch.cat([out["fraction_pred"], out["element_count_pred"].unsqueeze(-1)], dim=-1)
    heads = {k: out[k] for k in ("tc_pred", "sc_pred", "hp_pred", "tc_class_logits", "competence",
                                 "element_count_pred", "family_composed_14")}
    return stoich, heads
