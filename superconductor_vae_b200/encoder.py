"""Drop-in `FullMaterialsVAE` whose encode / head computations run on the B200 engine.

Mirrors src/superconductor/models/attention_vae.py: constructor :350-362, state_dict names :375-606,
``encode`` :625-676, ``decode`` :678-709, ``forward`` :711-822 (inference outputs; ``kl_loss`` is the
same mean(z^2) regulariser), plus two entry points latent-space callers need:
``heads_from_latent(z)`` (the notebook's ``_build_heads_pred``, notebooks/generative_evaluation.ipynb
cell 14:246-284) and ``conditioning(z)`` returning ``(stoich_pred, heads_pred)`` as the training script
assembles them (scripts/train_v12_clean.py:5245-5296).  No autograd: every method runs under no_grad.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib

HEADS_KEYS = ("tc_pred", "sc_pred", "hp_pred", "tc_class_logits", "competence", "element_count_pred",
              "family_composed_14")


def _mlp(*mods):
    return nn.Sequential(*mods)


class _ElementEmbedding(nn.Module):
    def __init__(self, n_elements: int, dim: int, property_dim: int = 11):
        super().__init__()
        self.element_embed = nn.Embedding(n_elements + 1, dim, padding_idx=0)
        # present in reference checkpoints, not executed (element_properties is always None)
        self.property_encoder = _mlp(nn.Linear(property_dim, dim), nn.LayerNorm(dim), nn.GELU())
        self.combiner = nn.Linear(dim * 2, dim)


class _ElementAttention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.query = nn.Parameter(torch.randn(heads, dim // heads))
        self.key_proj, self.value_proj, self.output_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.layer_norm = nn.LayerNorm(dim)
        nn.init.xavier_uniform_(self.query)


class _ElementEncoder(nn.Module):
    def __init__(self, n_elements: int, dim: int, heads: int, out_dim: int, dropout: float):
        super().__init__()
        self.element_embedding = _ElementEmbedding(n_elements, dim)
        self.element_attention = _ElementAttention(dim, heads)
        self.output_projection = _mlp(nn.Linear(dim, out_dim), nn.LayerNorm(out_dim), nn.GELU(), nn.Dropout(dropout))


class _VAEEncoder(nn.Module):
    def __init__(self, input_dim: int, hidden: List[int], latent_dim: int):
        super().__init__()
        layers, prev = [], input_dim
        for h in hidden:
            layers += [nn.Linear(prev, h), nn.LayerNorm(h), nn.GELU()]
            prev = h
        self.encoder = nn.Sequential(*layers)
        self.fc_mean = nn.Linear(prev, latent_dim)


class _FamilyHead(nn.Module):
    def __init__(self, bb: int, dropout: float):
        super().__init__()
        self.coarse_head = _mlp(nn.Linear(bb + 1, 256), nn.LayerNorm(256), nn.GELU(), nn.Dropout(dropout),
                                nn.Linear(256, 128), nn.GELU(), nn.Linear(128, 7))
        self.cuprate_sub_head = _mlp(nn.Linear(bb + 1, 128), nn.LayerNorm(128), nn.GELU(), nn.Dropout(dropout),
                                     nn.Linear(128, 64), nn.GELU(), nn.Linear(64, 6))
        self.iron_sub_head = _mlp(nn.Linear(bb + 1, 64), nn.LayerNorm(64), nn.GELU(), nn.Dropout(dropout),
                                  nn.Linear(64, 2))


class FullMaterialsVAE(nn.Module):
    def __init__(self, n_elements: int = 118, element_embed_dim: int = 128, n_attention_heads: int = 8,
                 magpie_dim: int = 145, fusion_dim: int = 256, encoder_hidden: List[int] = [512, 256],
                 latent_dim: int = 2048, decoder_hidden: List[int] = [256, 512], dropout: float = 0.1,
                 use_numden_head: bool = False):
        super().__init__()
        if use_numden_head:
            raise NotImplementedError("numden_head (V12.41 only, disabled since V13.0) is outside the engine")
        self.n_elements, self.latent_dim, self.magpie_dim, self.fusion_dim = n_elements, latent_dim, magpie_dim, fusion_dim
        self.element_embed_dim, self.n_attention_heads = element_embed_dim, n_attention_heads
        self.encoder_hidden, self.decoder_hidden = list(encoder_hidden), list(decoder_hidden)
        self.max_elements = 12
        self.use_numden_head = False
        f, L = fusion_dim, latent_dim
        self.element_encoder = _ElementEncoder(n_elements, element_embed_dim, n_attention_heads, f, dropout)
        self.magpie_encoder = _mlp(nn.Linear(magpie_dim, 2 * f), nn.LayerNorm(2 * f), nn.GELU(), nn.Dropout(dropout),
                                   nn.Linear(2 * f, f), nn.LayerNorm(f), nn.GELU())
        self.tc_encoder = _mlp(nn.Linear(1, f // 2), nn.GELU(), nn.Linear(f // 2, f), nn.LayerNorm(f), nn.GELU())
        self.fusion = _mlp(nn.Linear(3 * f, 3 * f), nn.LayerNorm(3 * f), nn.GELU(), nn.Dropout(dropout))
        self.vae_encoder = _VAEEncoder(3 * f, self.encoder_hidden, L)
        layers, prev = [], L
        for h in self.decoder_hidden:
            layers += [nn.Linear(prev, h), nn.LayerNorm(h), nn.GELU(), nn.Dropout(dropout)]
            prev = h
        self.decoder_backbone = nn.Sequential(*layers)
        bb = prev
        self.tc_proj = nn.Linear(bb, 256)
        self.tc_res_block = _mlp(nn.Linear(256, 256), nn.LayerNorm(256), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(256, 256))
        with torch.no_grad():                        # identity init of the residual block (:455-459)
            nn.init.eye_(self.tc_res_block[0].weight); nn.init.zeros_(self.tc_res_block[0].bias)
            nn.init.eye_(self.tc_res_block[4].weight); nn.init.zeros_(self.tc_res_block[4].bias)
        self.tc_out = _mlp(nn.LayerNorm(256), nn.GELU(), nn.Linear(256, 128), nn.GELU(), nn.Linear(128, 1))
        self.magpie_head = _mlp(nn.Linear(bb, bb), nn.GELU(), nn.Linear(bb, magpie_dim))
        self.attended_head = _mlp(nn.Linear(bb, f), nn.LayerNorm(f))
        self.competence_head = _mlp(nn.Linear(L, L // 4), nn.GELU(), nn.Linear(L // 4, 1), nn.Sigmoid())
        self.fraction_head = _mlp(nn.Linear(L, 256), nn.LayerNorm(256), nn.GELU(), nn.Dropout(dropout),
                                  nn.Linear(256, 128), nn.GELU(), nn.Linear(128, self.max_elements + 1))
        self.hp_head = _mlp(nn.Linear(L, 256), nn.ReLU(), nn.Linear(256, 1))
        self.tc_class_head = _mlp(nn.Linear(bb, 256), nn.GELU(), nn.Dropout(dropout), nn.Linear(256, 5))
        sc_in = L + 1 + magpie_dim + 1 + self.max_elements + 1 + 1 + 5
        self.sc_head = _mlp(nn.Linear(sc_in, 512), nn.GELU(), nn.LayerNorm(512), nn.Dropout(dropout),
                            nn.Linear(512, 128), nn.GELU(), nn.Linear(128, 1))
        self.hierarchical_family_head = _FamilyHead(bb, dropout)
        self._engine = None
        self._engine_versions: Dict[str, Tuple[int, int]] = {}

    def get_config(self) -> Dict:
        return {"n_elements": self.n_elements, "latent_dim": self.latent_dim, "magpie_dim": self.magpie_dim,
                "fusion_dim": self.fusion_dim}

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], device="cuda", **overrides):
        sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}
        emb = sd["element_encoder.element_embedding.element_embed.weight"]
        q = sd["element_encoder.element_attention.query"]
        f = sd["element_encoder.output_projection.0.weight"].shape[0]
        enc_h, j = [], 0
        while f"vae_encoder.encoder.{3 * j}.weight" in sd:
            enc_h.append(sd[f"vae_encoder.encoder.{3 * j}.weight"].shape[0]); j += 1
        dec_h, j = [], 0
        while f"decoder_backbone.{4 * j}.weight" in sd:
            dec_h.append(sd[f"decoder_backbone.{4 * j}.weight"].shape[0]); j += 1
        kw = dict(n_elements=emb.shape[0] - 1, element_embed_dim=emb.shape[1], n_attention_heads=q.shape[0],
                  magpie_dim=sd["magpie_encoder.0.weight"].shape[1], fusion_dim=f, encoder_hidden=enc_h,
                  latent_dim=sd["vae_encoder.fc_mean.weight"].shape[0], decoder_hidden=dec_h)
        kw.update(overrides)
        m = cls(**kw)
        m.load_state_dict(sd, strict=False)
        return m.to(device).eval()

    @classmethod
    def from_reference(cls, module: nn.Module, device="cuda"):
        return cls.from_state_dict(module.state_dict(), device=device)

    # ------------------------------------------------------------------ engine plumbing
    def _sync_engine(self):
        L = _lib.lib()
        w = self.vae_encoder.fc_mean.weight
        _lib.require_cuda(w, "FullMaterialsVAE parameters")
        if self._engine is None:
            cfg = _lib.EncoderConfig(
                n_element_rows=self.n_elements + 1, element_embed_dim=self.element_embed_dim,
                n_attention_heads=self.n_attention_heads, max_elements=self.max_elements, magpie_dim=self.magpie_dim,
                fusion_dim=self.fusion_dim, latent_dim=self.latent_dim, n_encoder_hidden=len(self.encoder_hidden),
                encoder_hidden=(C.c_int32 * 4)(*(self.encoder_hidden + [0] * (4 - len(self.encoder_hidden)))),
                n_decoder_hidden=len(self.decoder_hidden),
                decoder_hidden=(C.c_int32 * 4)(*(self.decoder_hidden + [0] * (4 - len(self.decoder_hidden)))))
            h = C.c_void_p()
            with torch.cuda.device(w.device):
                _lib.check(L.scv_encoder_create(C.byref(cfg), C.byref(h)), "scv_encoder_create")
            self._engine, self._engine_device, self._engine_versions = h, w.device, {}
        elif self._engine_device != w.device:
            raise _lib.EngineError("module was moved to another device after its engine was created")
        stream = _lib.current_stream()
        for name, t in self.state_dict(keep_vars=True).items():
            key = (t.data_ptr(), t._version)
            if self._engine_versions.get(name) == key:
                continue
            src = t.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            _lib.check(L.scv_encoder_load_weight(self._engine, name.encode(), _lib.ptr(src), src.numel(), stream),
                       f"load_weight({name})")
            self._engine_versions[name] = key
        return L

    def __del__(self):
        try:
            if getattr(self, "_engine", None) is not None:
                _lib.lib().scv_encoder_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def encode(self, element_indices, element_fractions, element_mask, magpie_features, tc) -> Dict[str, torch.Tensor]:
        L = self._sync_engine()
        for n, t in (("element_indices", element_indices), ("element_fractions", element_fractions),
                     ("element_mask", element_mask), ("magpie_features", magpie_features), ("tc", tc)):
            _lib.require_cuda(t, n)
        dev = element_indices.device
        B, E = element_indices.shape
        if E != self.max_elements:
            raise RuntimeError(f"element slots {E} != max_elements {self.max_elements}")
        idx = element_indices.to(torch.int64).contiguous()
        frac = element_fractions.to(torch.float32).contiguous()
        mask = element_mask.to(torch.uint8).contiguous()
        mag = magpie_features.to(torch.float32).contiguous()
        tcv = tc.to(torch.float32).reshape(B).contiguous()
        z = torch.empty((B, self.latent_dim), dtype=torch.float32, device=dev)
        attn = torch.empty((B, E), dtype=torch.float32, device=dev)
        fused = torch.empty((B, 3 * self.fusion_dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.scv_encoder_encode(self._engine, B, _lib.ptr(idx), _lib.ptr(frac), _lib.ptr(mask),
                                            _lib.ptr(mag), _lib.ptr(tcv), _lib.ptr(z), _lib.ptr(attn), _lib.ptr(fused),
                                            _lib.current_stream()), "encode")
        emb = torch.nn.functional.embedding(idx, self.element_encoder.element_embedding.element_embed.weight)
        return {"z": z, "z_mean": z, "z_logvar": None, "attention_weights": attn, "element_embeddings": emb,
                "fused_repr": fused}

    @torch.no_grad()
    def heads_from_latent(self, z: torch.Tensor) -> Dict[str, torch.Tensor]:
        """All head outputs computed from z alone (decode + head section of forward)."""
        L = self._sync_engine()
        _lib.require_cuda(z, "z")
        z = z.to(torch.float32).contiguous()
        B, dev = z.size(0), z.device
        shapes = {"tc_pred": (B,), "magpie_pred": (B, self.magpie_dim), "attended_input": (B, self.fusion_dim),
                  "tc_class_logits": (B, 5), "competence": (B,), "fraction_pred": (B, self.max_elements),
                  "element_count_pred": (B,), "hp_pred": (B,), "sc_pred": (B,), "family_coarse_logits": (B, 7),
                  "family_cuprate_sub_logits": (B, 6), "family_iron_sub_logits": (B, 2),
                  "family_composed_14": (B, 14), "stoich_pred": (B, self.max_elements + 1), "heads_input": (B, 24)}
        out = {k: torch.empty(s, dtype=torch.float32, device=dev) for k, s in shapes.items()}
        ho = _lib.EncoderHeadsOut(**{k: out[k].data_ptr() for k in _lib.HEADS_OUT_FIELDS})
        with torch.cuda.device(dev):
            _lib.check(L.scv_encoder_heads(self._engine, B, _lib.ptr(z), C.byref(ho), _lib.current_stream()), "heads")
        return out

    def decode(self, z: torch.Tensor) -> Dict[str, torch.Tensor]:
        h = self.heads_from_latent(z)
        return {k: h[k] for k in ("tc_pred", "magpie_pred", "attended_input", "tc_class_logits")}

    def conditioning(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        h = self.heads_from_latent(z)
        return h["stoich_pred"], {k: h[k] for k in HEADS_KEYS}

    @torch.no_grad()
    def forward(self, element_indices, element_fractions, element_mask, magpie_features, tc) -> Dict[str, torch.Tensor]:
        enc = self.encode(element_indices, element_fractions, element_mask, magpie_features, tc)
        h = self.heads_from_latent(enc["z"])
        out = {"z": enc["z"], "z_mean": enc["z_mean"], "z_logvar": None, "kl_loss": torch.mean(enc["z"].pow(2)),
               "attention_weights": enc["attention_weights"], "element_embeddings": enc["element_embeddings"],
               "numden_pred": None}
        for k in ("tc_pred", "magpie_pred", "attended_input", "competence", "fraction_pred", "element_count_pred",
                  "hp_pred", "sc_pred", "tc_class_logits", "family_coarse_logits", "family_cuprate_sub_logits",
                  "family_iron_sub_logits", "family_composed_14"):
            out[k] = h[k]
        return out
