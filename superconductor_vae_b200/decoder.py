"""Drop-in `EnhancedTransformerDecoder` whose KV-cache generation runs on the B200 engine.

Mirrors the reference's public surface for the decode path
(src/superconductor/models/autoregressive_decoder.py):

  * constructor arguments and state_dict key names / shapes            :564-765
  * precompute_memory(z, encoder_skip, stoich_pred, heads_pred)         :875-899
  * generate_with_kv_cache(...)  same parameter order and defaults      :1321-1338
  * sample_for_reinforce(...)                                           :1559-1572

The module holds ordinary fp32 ``nn.Parameter``s under the reference's names (so
``load_state_dict`` of a reference checkpoint works and an optimiser can keep updating them); before
each engine call, parameters whose version changed are re-uploaded and rounded to bf16 once.
``forward`` (teacher forcing and the two-pass scheduled sampling) runs on the engine as an
inference pass; it builds no autograd graph.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib

PAD_IDX, START_IDX, END_IDX = 0, 1, 2
N_TOKEN_TYPES = 5
HEADS_ORDER = ("tc_pred", "sc_pred", "hp_pred", "tc_class_logits", "competence", "element_count_pred")


def _legacy_vocab() -> List[str]:
    """id -> string of the reference's pre-V13 character vocabulary (autoregressive_decoder.py:52-97):
    20 special tokens, the 118 element symbols, digits 0-9."""
    from .tokenizer import ELEMENTS
    special = ["<PAD>", "<START>", "<END>", ".", "(", ")", "[", "]", "{", "}", "·", "•", "+", "-", "_", "^", "/",
               "'", '"', "*"]
    return special + ELEMENTS[1:] + [str(i) for i in range(10)]


_LEGACY_VOCAB = _legacy_vocab()


class _PositionalEncoding(nn.Module):
    """Holds the sinusoidal ``pe`` buffer [1, max_len, d_model] (reference :392-413)."""

    def __init__(self, d_model: int, max_len: int, dropout: float):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe.unsqueeze(0))


class EnhancedTransformerDecoder(nn.Module):
    def __init__(self, latent_dim: int = 2048, d_model: int = 512, nhead: int = 8, num_layers: int = 12,
                 dim_feedforward: int = 2048, dropout: float = 0.1, max_len: int = 80, n_memory_tokens: int = 16,
                 encoder_skip_dim: int = 256, use_skip_connection: bool = True, use_stoich_conditioning: bool = True,
                 max_elements: int = 12, n_stoich_tokens: int = 4, use_gradient_checkpointing: bool = False,
                 use_position_dependent_tf: bool = False, tf_position_decay: float = 0.5,
                 vocab_size: Optional[int] = None, stoich_input_dim: Optional[int] = None,
                 memory_bottleneck_dim: int = 1024):
        super().__init__()
        self.latent_dim, self.d_model, self.nhead, self.num_layers = latent_dim, d_model, nhead, num_layers
        self.dim_feedforward = dim_feedforward
        self.max_len = max_len
        self.vocab_size = vocab_size if vocab_size is not None else 148      # legacy VOCAB_SIZE (:107)
        self.n_memory_tokens = n_memory_tokens
        self.use_skip_connection = use_skip_connection
        self.use_stoich_conditioning = use_stoich_conditioning
        self.max_elements = max_elements
        self.memory_bottleneck_dim = memory_bottleneck_dim
        self.encoder_skip_dim = encoder_skip_dim
        self.use_position_dependent_tf = use_position_dependent_tf      # scheduled sampling (:1037-1043)
        self.tf_position_decay = tf_position_decay

        self.token_embedding = nn.Embedding(self.vocab_size, d_model, padding_idx=PAD_IDX)
        self.pos_encoding = _PositionalEncoding(d_model, max_len, dropout)
        n_lat = d_model * n_memory_tokens
        if memory_bottleneck_dim > 0:
            self.latent_to_memory = nn.Sequential(nn.Linear(latent_dim, memory_bottleneck_dim),
                                                  nn.LayerNorm(memory_bottleneck_dim), nn.GELU(),
                                                  nn.Linear(memory_bottleneck_dim, n_lat))
        else:
            self.latent_to_memory = nn.Sequential(nn.Linear(latent_dim, n_lat // 2), nn.GELU(),
                                                  nn.Linear(n_lat // 2, n_lat))
        if use_skip_connection:
            self.skip_n_tokens = 8
            n_skip = d_model * self.skip_n_tokens
            self.skip_to_memory = nn.Sequential(nn.Linear(encoder_skip_dim, n_skip // 2), nn.GELU(),
                                                nn.Linear(n_skip // 2, n_skip))
        else:
            self.skip_n_tokens = 0
        if use_stoich_conditioning:
            self.stoich_n_tokens = n_stoich_tokens
            if stoich_input_dim is None:
                stoich_input_dim = max_elements * 3 + 1
            self.stoich_input_dim = stoich_input_dim
            self.stoich_to_memory = nn.Sequential(nn.Linear(stoich_input_dim, d_model), nn.LayerNorm(d_model),
                                                  nn.GELU(), nn.Linear(d_model, d_model * n_stoich_tokens))
        else:
            self.stoich_n_tokens = 0
            self.stoich_input_dim = 0
        layer = nn.TransformerDecoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward,
                                           dropout=dropout, activation="gelu", batch_first=True, norm_first=True)
        self.transformer_decoder = nn.TransformerDecoder(layer, num_layers=num_layers)
        self.output_proj = nn.Sequential(nn.LayerNorm(d_model), nn.Linear(d_model, d_model), nn.GELU(),
                                         nn.Dropout(dropout), nn.Linear(d_model, self.vocab_size))
        self.stop_head = nn.Sequential(nn.Linear(d_model, d_model // 4), nn.GELU(), nn.Linear(d_model // 4, 1))
        self.site_dup_head = nn.Sequential(nn.Linear(d_model, d_model // 4), nn.GELU(), nn.Linear(d_model // 4, 1))
        self.token_type_head = nn.Sequential(nn.LayerNorm(d_model), nn.Linear(d_model, d_model), nn.GELU(),
                                             nn.Dropout(dropout), nn.Linear(d_model, d_model // 4), nn.GELU(),
                                             nn.Dropout(dropout), nn.Linear(d_model // 4, N_TOKEN_TYPES))
        self.heads_input_dim = 24
        self.heads_n_tokens = 4
        self.heads_to_memory = nn.Sequential(nn.Linear(self.heads_input_dim, d_model // 2),
                                             nn.LayerNorm(d_model // 2), nn.GELU(), nn.Linear(d_model // 2, d_model),
                                             nn.GELU(), nn.Linear(d_model, d_model * self.heads_n_tokens))
        for p in self.parameters():          # same init rule as the reference (:769-772)
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._engine = None
        self._engine_versions: Dict[str, Tuple[int, int]] = {}
        self.max_rows_per_call = 8192        # rows decoded per engine call; larger batches are chunked
        self.h2_uniform_fallback = True      # reproduce the reference's batch-global degenerate guard
        # Opt-in: retire a row once it has emitted END instead of decoding it until EVERY row has finished
        # like the reference does (:1541-1548).  Tokens / log-probs / entropy are identical up to and including each
        # row's first END - what every caller consumes - and PAD / 0.0 after it; the executed length L is the same.
        self.compact_finished = False
        self.last_steps: Optional[int] = None

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], nhead: int = 8, device="cuda", **overrides):
        """Infer every shape from the checkpoint the way the reference's loaders do
        (scripts/holdout/holdout_search.py:223-252) and load it."""
        sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}      # torch.compile prefixes (:4042-4058)
        d = sd["token_embedding.weight"].shape[1]
        layers = 0
        while f"transformer_decoder.layers.{layers}.self_attn.in_proj_weight" in sd:
            layers += 1
        if "latent_to_memory.3.weight" in sd:
            bottleneck = sd["latent_to_memory.0.weight"].shape[0]
            n_lat = sd["latent_to_memory.3.weight"].shape[0] // d
        else:
            bottleneck = 0
            n_lat = sd["latent_to_memory.2.weight"].shape[0] // d
        kw = dict(latent_dim=sd["latent_to_memory.0.weight"].shape[1], d_model=d, nhead=nhead, num_layers=layers,
                  dim_feedforward=sd["transformer_decoder.layers.0.linear1.weight"].shape[0],
                  max_len=sd["pos_encoding.pe"].shape[1], n_memory_tokens=n_lat,
                  use_skip_connection="skip_to_memory.0.weight" in sd,
                  vocab_size=sd["token_embedding.weight"].shape[0],
                  use_stoich_conditioning="stoich_to_memory.0.weight" in sd,
                  memory_bottleneck_dim=bottleneck)
        if "skip_to_memory.0.weight" in sd:
            kw["encoder_skip_dim"] = sd["skip_to_memory.0.weight"].shape[1]
        if "stoich_to_memory.0.weight" in sd:
            kw["stoich_input_dim"] = sd["stoich_to_memory.0.weight"].shape[1]
            kw["n_stoich_tokens"] = sd["stoich_to_memory.3.weight"].shape[0] // d
        kw.update(overrides)
        m = cls(**kw)
        k_heads = sd["heads_to_memory.0.weight"].shape[1] if "heads_to_memory.0.weight" in sd else m.heads_input_dim
        if k_heads != m.heads_input_dim:      # V14.3 checkpoints predate the 14 family dims (SURVEY 8 "Configurations")
            m.heads_input_dim = k_heads
            m.heads_to_memory[0] = nn.Linear(k_heads, d // 2)
        m.load_state_dict(sd, strict=False)
        return m.to(device).eval()

    @classmethod
    def from_reference(cls, module: nn.Module, device="cuda"):
        """Build from an instance of the reference class (same state_dict layout)."""
        return cls.from_state_dict(module.state_dict(), nhead=module.nhead, device=device)

    def forward(self, z, target_tokens, encoder_skip=None, teacher_forcing_ratio: float = 1.0, stoich_pred=None,
                cached_memory=None, heads_pred=None, *, _use_gt_mask: Optional[torch.Tensor] = None):
        """Teacher-forced forward (reference :901-1082), inference only (no autograd graph): every position
        of every row in one engine pass.

        ``teacher_forcing_ratio >= 1``: the parallel pass on the ground-truth inputs (:947-985).
        ``teacher_forcing_ratio < 1``: the reference's two-pass scheduled sampling (:987-1082): pass 1 on the ground truth,
        its argmax tokens are mixed with the ground truth position by position (``rand < ratio`` keeps the ground truth;
        with ``use_position_dependent_tf`` the ratio is scaled by ``1 + tf_position_decay * (1 - pos / (L - 1))``), pass 2
        runs on START + the mixed tokens.  The mask is drawn with ``torch.rand`` on the module's device like the
        reference does; tests inject it through ``_use_gt_mask`` ([B, L-1] bool) to compare with a CPU run.

        Returns ``(logits [B, L-1, V], generated [B, L-1], stop_logits [B, L-1], type_logits [B, L-1, 5],
        site_dup_logits [B, L-1] or None when the checkpoint has no site_dup_head)``, like the reference."""
        with torch.no_grad():
            self._sync_engine()
            memory = self._checked_memory(cached_memory) if cached_memory is not None else \
                self._create_memory(z, encoder_skip, stoich_pred, heads_pred)
            device = memory.device
            B = memory.size(0)
            tokens = target_tokens.to(device=device, dtype=torch.int64)
            if tokens.dim() != 2 or tokens.size(0) != B or tokens.size(1) < 2:
                raise RuntimeError(f"target_tokens must be [{B}, seq_len >= 2], got {tuple(tokens.shape)}")
            if int(tokens.min()) < 0 or int(tokens.max()) >= self.vocab_size:
                raise IndexError("index out of range in self")                  # nn.Embedding's error
            seq = tokens.size(1) - 1
            if seq > self.pos_encoding.pe.shape[1]:
                raise RuntimeError(f"The size of tensor a ({seq}) must match the size of tensor b "
                                   f"({self.pos_encoding.pe.shape[1]}) at non-singleton dimension 1")
            inputs = tokens[:, :-1].contiguous()
            if teacher_forcing_ratio >= 1.0:
                logits, stop, typ, dup = self._forward_pass(memory, inputs)
                return logits, logits.argmax(dim=-1), stop, typ, dup
            first_logits, _, _, _ = self._forward_pass(memory, inputs, heads=False)             # pass 1 (:1012-1031)
            predicted = first_logits.argmax(dim=-1)
            if _use_gt_mask is not None:
                use_gt = _use_gt_mask.to(device=device, dtype=torch.bool)
                if tuple(use_gt.shape) != (B, seq):
                    raise RuntimeError(f"_use_gt_mask must be [{B}, {seq}], got {tuple(use_gt.shape)}")
            elif self.use_position_dependent_tf:                                               # (:1037-1043)
                pos = torch.arange(seq, device=device).float() / max(seq - 1, 1)
                tf_pos = (teacher_forcing_ratio * (1.0 + self.tf_position_decay * (1.0 - pos))).clamp(0.0, 1.0)
                use_gt = torch.rand(B, seq, device=device) < tf_pos.unsqueeze(0)
            else:
                use_gt = torch.rand(B, seq, device=device) < teacher_forcing_ratio              # (:1045)
            mixed = torch.where(use_gt, tokens[:, 1:], predicted)                               # (:1051)
            mixed_inputs = torch.cat([tokens[:, :1], mixed[:, :-1]], dim=1).contiguous()        # (:1056)
            logits, stop, typ, dup = self._forward_pass(memory, mixed_inputs)                   # pass 2 (:1058-1082)
            return logits, logits.argmax(dim=-1), stop, typ, dup

    def _forward_pass(self, memory: torch.Tensor, inputs: torch.Tensor, heads: bool = True, key_padding: bool = True):
        """One parallel pass of the layer stack over ``inputs`` [B, L] (PAD inputs are masked as keys, :952, unless
        ``key_padding=False``: generation attends every earlier position)."""
        L = _lib.lib()
        device = memory.device
        B, M, seq = memory.size(0), memory.size(1), inputs.size(1)
        logits = torch.empty((B, seq, self.vocab_size), dtype=torch.float32, device=device)
        stop = torch.empty((B, seq), dtype=torch.float32, device=device) if heads else None
        typ = torch.empty((B, seq, N_TOKEN_TYPES), dtype=torch.float32, device=device) if heads else None
        has_dup = heads and hasattr(self, "site_dup_head")
        dup = torch.empty((B, seq), dtype=torch.float32, device=device) if has_dup else None
        args = _lib.ForwardArgs(batch=B, seq_len=seq, n_memory=M, memory=_lib.ptr(memory), tokens=_lib.ptr(inputs),
                                ld_tokens=inputs.size(1), out_logits=_lib.ptr(logits), out_stop=_lib.ptr(stop),
                                out_type=_lib.ptr(typ), out_dup=_lib.ptr(dup),
                                flags=0 if key_padding else _lib.FORWARD_NO_KEY_PADDING)
        with torch.cuda.device(device):
            _lib.check(L.scv_decoder_forward(self._engine, C.byref(args), _lib.current_stream()), "forward")
        return logits, stop, typ, dup

    # ------------------------------------------------------------------ engine plumbing
    def _config(self) -> _lib.DecoderConfig:
        return _lib.DecoderConfig(
            d_model=self.d_model, nhead=self.nhead, num_layers=self.num_layers, dim_feedforward=self.dim_feedforward,
            vocab_size=self.vocab_size, pe_len=self.pos_encoding.pe.shape[1], latent_dim=self.latent_dim,
            n_memory_tokens=self.n_memory_tokens, memory_bottleneck_dim=self.memory_bottleneck_dim,
            stoich_input_dim=self.stoich_input_dim, n_stoich_tokens=self.stoich_n_tokens,
            heads_input_dim=self.heads_input_dim, heads_n_tokens=self.heads_n_tokens,
            encoder_skip_dim=self.encoder_skip_dim if self.use_skip_connection else 0,
            skip_n_tokens=self.skip_n_tokens)

    def _sync_engine(self):
        """Create the engine on first use and upload every tensor whose storage or version changed."""
        L = _lib.lib()
        dev = self.token_embedding.weight.device
        _lib.require_cuda(self.token_embedding.weight, "EnhancedTransformerDecoder parameters")
        if self._engine is None:
            h = C.c_void_p()
            cfg = self._config()
            with torch.cuda.device(dev):
                _lib.check(L.scv_decoder_create(C.byref(cfg), C.byref(h)), "scv_decoder_create")
            self._engine = h
            self._engine_device = dev
            self._engine_versions = {}
        elif self._engine_device != dev:
            raise _lib.EngineError("module was moved to another device after its engine was created")
        stream = _lib.current_stream()
        for name, t in self.state_dict(keep_vars=True).items():
            key = (t.data_ptr(), t._version)
            if self._engine_versions.get(name) == key:
                continue
            src = t.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            _lib.check(L.scv_decoder_load_weight(self._engine, name.encode(), _lib.ptr(src), src.numel(), stream),
                       f"load_weight({name})")
            self._engine_versions[name] = key      # (a temporary `src` is safe to drop: same-stream reuse is ordered)
        return L

    def __del__(self):
        try:
            if getattr(self, "_engine", None) is not None:
                _lib.lib().scv_decoder_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    @staticmethod
    def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
        if t is None:
            return None
        _lib.require_cuda(t, name)
        return t.detach().to(torch.float32).contiguous()

    @staticmethod
    def _check_rows(t: torch.Tensor, name: str, batch: int, features: int, weight: str) -> None:
        """The kernels read raw pointers with the configured strides, so every shape the reference would reject in
        F.linear / view / cat is rejected here with the same exception type (RuntimeError)."""
        if t.dim() != 2 or t.size(1) != features:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({name} {tuple(t.shape)} against "
                               f"{weight} [*, {features}])")
        if t.size(0) != batch:
            raise RuntimeError(f"{name} batch {t.size(0)} != z batch {batch}. Shapes: {name}={tuple(t.shape)}, "
                               f"z=[{batch}, ...]")

    def _heads_matrix(self, heads_pred: Dict[str, torch.Tensor], batch: int, device) -> torch.Tensor:
        """[B, 24] in the order tc, sc, hp, tc_class(5), competence, count, family(14) (:845-858)."""
        fam = heads_pred.get("family_composed_14")
        for name in HEADS_ORDER + (("family_composed_14",) if fam is not None else ()):
            t = heads_pred[name]
            if t.size(0) != batch:
                raise RuntimeError(f"heads_pred['{name}'] batch {t.size(0)} != z batch {batch}. "
                                   f"Shapes: {name}={tuple(t.shape)}, z=[{batch}, ...]")
        parts = [heads_pred["tc_pred"].unsqueeze(-1), heads_pred["sc_pred"].unsqueeze(-1),
                 heads_pred["hp_pred"].unsqueeze(-1), heads_pred["tc_class_logits"],
                 heads_pred["competence"].unsqueeze(-1), heads_pred["element_count_pred"].unsqueeze(-1)]
        if self.heads_input_dim > 10:
            parts.append(fam if fam is not None else torch.zeros(batch, self.heads_input_dim - 10, device=device))
        for p_ in parts:
            if p_.dim() != 2:
                raise RuntimeError(f"heads_pred entries must be [B] scalars or [B, k] rows, got a part of shape {tuple(p_.shape)}")
        out = torch.cat([p.to(device=device, dtype=torch.float32) for p in parts], dim=-1).contiguous()
        if out.size(1) != self.heads_input_dim:          # the reference fails in heads_to_memory[0] (F.linear)
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (heads_input {tuple(out.shape)} against "
                               f"heads_to_memory.0.weight [*, {self.heads_input_dim}])")
        return out

    def _checked_memory(self, cached_memory: torch.Tensor) -> torch.Tensor:
        m = self._f32(cached_memory, "cached_memory")
        if m.dim() != 3 or m.size(2) != self.d_model or m.size(1) < 1:
            raise RuntimeError(f"cached_memory must be [B, n_tokens, {self.d_model}] (precompute_memory's output), "
                               f"got {tuple(m.shape)}")
        return m

    # ------------------------------------------------------------------ reference API
    def _create_memory(self, z, encoder_skip=None, stoich_pred=None, heads_pred=None) -> torch.Tensor:
        with torch.no_grad():
            L = self._sync_engine()
            z = self._f32(z, "z")
            if z.dim() != 2 or z.size(1) != self.latent_dim:
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (z {tuple(z.shape)} against "
                                   f"latent_to_memory.0.weight [*, {self.latent_dim}])")
            B = z.size(0)
            skip = self._f32(encoder_skip, "encoder_skip") if (self.use_skip_connection and encoder_skip is not None) else None
            stoich = self._f32(stoich_pred, "stoich_pred") if (self.use_stoich_conditioning and stoich_pred is not None) else None
            if skip is not None:
                self._check_rows(skip, "encoder_skip", B, self.encoder_skip_dim, "skip_to_memory.0.weight")
            if stoich is not None:
                self._check_rows(stoich, "stoich_pred", B, self.stoich_input_dim, "stoich_to_memory.0.weight")
            heads = self._heads_matrix(heads_pred, B, z.device) if heads_pred is not None else None
            M = self.n_memory_tokens + (self.skip_n_tokens if skip is not None else 0) + \
                (self.stoich_n_tokens if stoich is not None else 0) + (self.heads_n_tokens if heads is not None else 0)
            memory = torch.empty((B, M, self.d_model), dtype=torch.float32, device=z.device)
            m_out = C.c_int32(0)
            with torch.cuda.device(z.device):
                _lib.check(L.scv_decoder_build_memory(self._engine, B, _lib.ptr(z), _lib.ptr(skip), _lib.ptr(stoich),
                                                      _lib.ptr(heads), _lib.ptr(memory), C.byref(m_out),
                                                      _lib.current_stream()), "build_memory")
            assert m_out.value == M
            return memory

    def precompute_memory(self, z, encoder_skip=None, stoich_pred=None, heads_pred=None) -> torch.Tensor:
        return self._create_memory(z, encoder_skip, stoich_pred, heads_pred)

    def generate_with_kv_cache(self, z, encoder_skip=None, stoich_pred=None, temperature: float = 1.0,
                               top_k: Optional[int] = None, top_p: Optional[float] = None,
                               max_len: Optional[int] = None, return_log_probs: bool = False,
                               return_entropy: bool = False, cached_memory: Optional[torch.Tensor] = None,
                               stop_boost: float = 0.0, hard_stop_threshold: float = 0.0,
                               heads_pred: Optional[Dict[str, torch.Tensor]] = None,
                               type_masks: Optional[torch.Tensor] = None, site_dup_threshold: float = 0.0,
                               *, _forced_tokens: Optional[torch.Tensor] = None, _seed: Optional[int] = None,
                               _n_samples: int = 1
                               ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Reference signature and semantics (:1321-1557).  Keyword-only extension ``_n_samples = k`` (RLOO): the
        conditioning inputs (or ``cached_memory``) are the BASE batch [B0, ...] and k * B0 rows are decoded in the
        reference's ``repeat`` layout (row i * B0 + b = sample i of latent b, scripts/train_v12_clean.py:2677-2688) without
        materialising the k copies: memory tokens and their projected K / V are built once per latent and shared by its
        samples.  Same outputs as passing ``z.repeat(k, 1)`` etc."""
        self.eval()                                                         # side effect kept (:1368)
        max_len = max_len or self.max_len
        pe_max = self.pos_encoding.pe.shape[1]
        if max_len > pe_max:
            max_len = pe_max                                                # silent clamp (:1372-1375)
        with torch.no_grad():
            L = self._sync_engine()
            memory = self._checked_memory(cached_memory) if cached_memory is not None else \
                self._create_memory(z, encoder_skip, stoich_pred, heads_pred)
            device = memory.device
            B, M = memory.size(0), memory.size(1)
            memory_rows = 0
            if _n_samples > 1:
                if B * _n_samples > 64 and B * _n_samples <= int(self.max_rows_per_call):
                    memory_rows, B = B, B * _n_samples                      # shared by the samples inside the engine
                else:                                                       # tiny or chunked batches: plain copies
                    memory = memory.repeat(_n_samples, 1, 1)
                    B = memory.size(0)
            steps_max = max_len - 1
            if steps_max < 1:
                raise RuntimeError("max_len leaves no decoding step (the reference fails in torch.cat here)")
            masks_u8 = None
            if type_masks is not None:
                masks_u8 = type_masks.to(device=device).to(torch.uint8).contiguous()
                if tuple(masks_u8.shape) != (N_TOKEN_TYPES, self.vocab_size):
                    raise RuntimeError(f"type_masks must be [{N_TOKEN_TYPES}, {self.vocab_size}], got {tuple(masks_u8.shape)}")
            tokens = torch.zeros((B, steps_max), dtype=torch.int64, device=device)
            lps = torch.zeros((B, steps_max), dtype=torch.float32, device=device) if return_log_probs else None
            ents = torch.zeros((B, steps_max), dtype=torch.float32, device=device) if return_entropy else None
            forced = None
            if _forced_tokens is not None:
                if _forced_tokens.dim() != 2 or _forced_tokens.size(0) != B:
                    raise RuntimeError(f"_forced_tokens must be [{B}, L], got {tuple(_forced_tokens.shape)}")
                # positions beyond the given columns, and negative entries, stay sampled (-1)
                forced = torch.full((B, steps_max), -1, dtype=torch.int64, device=device)
                n = min(steps_max, _forced_tokens.size(1))
                forced[:, :n] = _forced_tokens[:, :n].to(device)
            if _seed is None:
                _seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if not (temperature < 0.01) else 0
            flags = _lib.FLAG_H2_UNIFORM_FALLBACK if self.h2_uniform_fallback else 0
            if self.compact_finished:
                flags |= _lib.FLAG_COMPACT_FINISHED
            steps_done = 0
            chunk = max(1, int(self.max_rows_per_call))
            with torch.cuda.device(device):
                stream = _lib.current_stream()
                for lo in range(0, B, chunk):
                    hi = min(B, lo + chunk)
                    out_steps = C.c_int32(0)
                    args = _lib.GenerateArgs(
                        batch=hi - lo, n_memory=M, max_len=max_len, memory_rows=memory_rows,
                        memory=(memory if memory_rows else memory[lo:hi]).data_ptr(),
                        temperature=float(temperature), top_k=int(top_k) if top_k else 0,
                        top_p=float(top_p) if top_p is not None else 1.0, stop_boost=float(stop_boost),
                        hard_stop_threshold=float(hard_stop_threshold), site_dup_threshold=float(site_dup_threshold or 0.0),
                        type_masks=masks_u8.data_ptr() if masks_u8 is not None else None,
                        want_log_probs=int(return_log_probs), want_entropy=int(return_entropy), flags=flags,
                        seed=_seed, offset=lo,
                        out_tokens=tokens[lo:hi].data_ptr(),
                        out_log_probs=lps[lo:hi].data_ptr() if lps is not None else None,
                        out_entropy=ents[lo:hi].data_ptr() if ents is not None else None,
                        out_steps=C.pointer(out_steps),
                        forced_tokens=forced[lo:hi].data_ptr() if forced is not None else None)
                    _lib.check(L.scv_decoder_generate(self._engine, C.byref(args), stream), "generate")
                    steps_done = max(steps_done, out_steps.value)
                    self._last_B = hi - lo
            self.last_steps = steps_done
            if self.compact_finished:
                # same outputs whichever kernel path decoded the rows: nothing after a row's first END
                t_ = tokens[:, :steps_done]
                after = (torch.cumsum((t_ == END_IDX).to(torch.int32), dim=1) - (t_ == END_IDX).to(torch.int32)) > 0
                tokens[:, :steps_done].masked_fill_(after, PAD_IDX)
                if lps is not None:
                    lps[:, :steps_done].masked_fill_(after, 0.0)
                if ents is not None:
                    ents[:, :steps_done].masked_fill_(after, 0.0)
            gen = tokens[:, :steps_done].contiguous()
            return (gen, lps[:, :steps_done].contiguous() if lps is not None else None,
                    ents[:, :steps_done].contiguous() if ents is not None else None)

    def generate_with_draft(self, z, draft_tokens, encoder_skip=None, stoich_pred=None, temperature: float = 0.001,
                            max_len: Optional[int] = None, cached_memory: Optional[torch.Tensor] = None,
                            stop_boost: float = 0.0, hard_stop_threshold: float = 0.0,
                            heads_pred: Optional[Dict[str, torch.Tensor]] = None, type_masks: Optional[torch.Tensor] = None,
                            max_passes: int = 4) -> Tuple[torch.Tensor, int, int]:
        """Greedy decoding by DRAFT VERIFICATION (replaces the reference's disabled n-gram speculative
        sampler, models/autoregressive_decoder.py:1643-1984): instead of one dependent step per token, every position of a
        drafted sequence is checked in ONE teacher-forced pass (the f3 engine pass with generation's attention semantics
        + the decode's greedy epilogue at every position).  If the pass reproduces the draft up to its END the row is done;
        otherwise the pass's own predictions become the next draft (a Jacobi / fixed-point iteration: the correct prefix
        grows by at least one token per pass, so the result is EXACTLY the greedy sequence however bad the draft is).
        Rows that have not converged after ``max_passes`` are finished by the ordinary KV-cache decode.

        ``draft_tokens`` [B, <= max_len - 1] are z-conditioned guesses, e.g. the sequence decoded for a neighbouring
        latent of a SLERP walk, or the previous epoch's greedy sequence of the same sample (both differ from the true
        sequence in a few positions at most); shorter drafts are padded.  Greedy only (``temperature < 0.01``).
        Returns ``(tokens [B, L], n_passes, n_fallback_rows)``: identical to ``generate_with_kv_cache`` up to and including
        each row's first END, PAD after it (like ``compact_finished``)."""
        if not (0.0 < temperature < 0.01):
            raise ValueError("generate_with_draft is greedy decoding: 0 < temperature < 0.01 (the reference's argmax branch)")
        self.eval()
        max_len = min(max_len or self.max_len, self.pos_encoding.pe.shape[1])
        with torch.no_grad():
            L_ = self._sync_engine()
            memory = self._checked_memory(cached_memory) if cached_memory is not None else \
                self._create_memory(z, encoder_skip, stoich_pred, heads_pred)
            device = memory.device
            B, steps = memory.size(0), max_len - 1
            if steps < 1:
                raise RuntimeError("max_len leaves no decoding step")
            masks_u8 = None
            if type_masks is not None:
                masks_u8 = type_masks.to(device=device).to(torch.uint8).contiguous()
                if tuple(masks_u8.shape) != (N_TOKEN_TYPES, self.vocab_size):
                    raise RuntimeError(f"type_masks must be [{N_TOKEN_TYPES}, {self.vocab_size}], got {tuple(masks_u8.shape)}")
            draft = draft_tokens.to(device=device, dtype=torch.int64)
            if draft.dim() != 2 or draft.size(0) != B:
                raise RuntimeError(f"draft_tokens must be [{B}, <= {steps}], got {tuple(draft.shape)}")
            cur = torch.zeros((B, steps), dtype=torch.int64, device=device)
            n = min(steps, draft.size(1))
            cur[:, :n] = draft[:, :n].clamp(0, self.vocab_size - 1)
            start = torch.full((B, 1), START_IDX, dtype=torch.int64, device=device)
            pos = torch.arange(steps, device=device).unsqueeze(0)
            pred = cur
            done = torch.zeros(B, dtype=torch.bool, device=device)
            passes = 0
            while passes < max(1, int(max_passes)):
                passes += 1
                inputs = torch.cat([start, cur[:, :-1]], dim=1).contiguous()
                logits, stop, typ, _ = self._forward_pass(memory, inputs, key_padding=False)
                is_end = cur == END_IDX
                fin_before = ((torch.cumsum(is_end.to(torch.int32), dim=1) - is_end.to(torch.int32)) > 0).to(torch.uint8).contiguous()
                pred = torch.empty((B, steps), dtype=torch.int64, device=device)
                with torch.cuda.device(device):
                    _lib.check(L_.scv_greedy_positions(
                        _lib.ptr(logits), _lib.ptr(typ) if masks_u8 is not None else None,
                        _lib.ptr(stop) if stop_boost > 0 else None, _lib.ptr(masks_u8), _lib.ptr(fin_before), B * steps, steps,
                        self.vocab_size, max_len, float(temperature), float(stop_boost), float(hard_stop_threshold),
                        _lib.ptr(pred), _lib.current_stream()), "greedy_positions")
                # a row has converged when the pass reproduces its own inputs up to and including its first END
                p_end = pred == END_IDX
                first_end = torch.where(p_end.any(dim=1), p_end.int().argmax(dim=1), torch.full((B,), steps - 1, device=device))
                live = pos <= first_end.unsqueeze(1)
                done = ((pred == cur) | ~live).all(dim=1)
                cur = pred
                if bool(done.all()):
                    break
            tokens = pred.clone()
            rest = (~done).nonzero().flatten()
            if rest.numel() > 0:                      # drafts too far off: finish those rows step by step
                t_, _, _ = self.generate_with_kv_cache(None, temperature=temperature, max_len=max_len,
                                                       cached_memory=memory[rest].contiguous(), stop_boost=stop_boost,
                                                       hard_stop_threshold=hard_stop_threshold, type_masks=type_masks)
                tokens[rest] = 0
                tokens[rest, :t_.shape[1]] = t_
            t_end = tokens == END_IDX
            after = (torch.cumsum(t_end.to(torch.int32), dim=1) - t_end.to(torch.int32)) > 0
            tokens.masked_fill_(after, PAD_IDX)
            ends = torch.where(t_end.any(dim=1), t_end.int().argmax(dim=1) + 1, torch.full((B,), steps, device=device))
            return tokens[:, :int(ends.max())].contiguous(), passes, int(rest.numel())

    def sample_for_reinforce(self, z, encoder_skip=None, stoich_pred=None, temperature: float = 0.8,
                             max_len: Optional[int] = None, cached_memory: Optional[torch.Tensor] = None,
                             stop_boost: float = 0.0, hard_stop_threshold: float = 0.0, heads_pred=None,
                             type_masks=None, site_dup_threshold: float = 0.0, **kw
                             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        tokens, lp, ent = self.generate_with_kv_cache(
            z=z, encoder_skip=encoder_skip, stoich_pred=stoich_pred, temperature=temperature, max_len=max_len,
            return_log_probs=True, return_entropy=True, cached_memory=cached_memory, stop_boost=stop_boost,
            hard_stop_threshold=hard_stop_threshold, heads_pred=heads_pred, type_masks=type_masks,
            site_dup_threshold=site_dup_threshold, **kw)
        B, Lq = tokens.shape                                               # mask: 1 up to and incl. first END (:1620-1639)
        is_end = tokens == END_IDX
        end_pos = torch.argmax(is_end.int(), dim=1)
        end_pos = torch.where(is_end.any(dim=1), end_pos, torch.full_like(end_pos, Lq))
        mask = (torch.arange(Lq, device=tokens.device).unsqueeze(0) <= end_pos.unsqueeze(1)).float()
        return tokens, lp, ent, mask

    @torch.no_grad()
    def rollout_with_rewards(self, z, targets: torch.Tensor, n_samples: int = 1, reward_config=None, constraint_config=None,
                             family_predictions: Optional[torch.Tensor] = None, family_config=None,
                             use_semantic_fractions: bool = False, fraction_token_start: int = 0,
                             fraction_values: Optional[torch.Tensor] = None, **rollout_kwargs):
        """One call for the rollout block of the reference's RL loss (scripts/train_v12_clean.py:2677-2766):
        expand the batch to n_samples rollouts per latent (sample-major, the samples of a latent share its memory tokens),
        sample with log-probs / entropy / mask, pad or truncate to the targets' length, score every rollout with
        `compute_reward_gpu_native` and, when `constraint_config` is given, add `compute_constraint_rewards` - all
        enqueued on the current stream, nothing returns to the host between the decode and the reward kernels.
        `rollout_kwargs` are `sample_for_reinforce`'s (BASE-batch tensors).  Returns
        (sampled_tokens, log_probs, entropy, mask, task_rewards), each with n_samples * batch rows."""
        from . import constraints as _constraints, reward as _reward
        k = int(n_samples)
        if k < 1:
            raise ValueError("n_samples must be >= 1")
        if k > 1:
            rollout_kwargs["_n_samples"] = k
        tokens, lp, ent, mask = self.sample_for_reinforce(z, **rollout_kwargs)
        T = targets.size(1)
        if tokens.size(1) < T:                                             # (:2733-2744)
            pad = (0, T - tokens.size(1))
            tokens, lp = nn.functional.pad(tokens, pad, value=PAD_IDX), nn.functional.pad(lp, pad, value=0.0)
            ent, mask = nn.functional.pad(ent, pad, value=0.0), nn.functional.pad(mask, pad, value=0.0)
        elif tokens.size(1) > T:
            tokens, lp, ent, mask = tokens[:, :T], lp[:, :T], ent[:, :T], mask[:, :T]
        tgt = targets.to(tokens.device)
        tgt = tgt.repeat(k, 1) if k > 1 else tgt                            # (:2688)
        rewards = _reward.compute_reward_gpu_native(tokens, tgt, mask.bool(), config=reward_config, pad_idx=PAD_IDX,
                                                    end_idx=END_IDX, use_semantic_fractions=use_semantic_fractions,
                                                    fraction_token_start=fraction_token_start, fraction_values=fraction_values)
        if constraint_config is not None:                                   # (:2754-2766)
            fam = family_predictions
            if fam is not None and k > 1:
                fam = fam.repeat(k, 1)
            rewards = rewards + _constraints.compute_constraint_rewards(tokens, mask, config=constraint_config,
                                                                        family_predictions=fam, family_config=family_config)
        return tokens, lp, ent, mask, rewards

    def generate_formulas_fast(self, z, encoder_skip=None, stoich_pred=None, temperature: float = 1.0,
                               max_len: Optional[int] = None, cached_memory: Optional[torch.Tensor] = None,
                               tokenizer=None) -> List[str]:
        """Reference :1998-2032: KV-cache generation, then ids -> strings.  Like the reference it uses the legacy
        148-token character vocabulary (20 specials, 118 elements, digits 0-9; ids outside it map to '') unless a
        `FractionAwareTokenizer` is passed, in which case its `decode` is used."""
        tokens, _, _ = self.generate_with_kv_cache(z=z, encoder_skip=encoder_skip, stoich_pred=stoich_pred,
                                                   temperature=temperature, max_len=max_len, cached_memory=cached_memory)
        rows = tokens.cpu().tolist()
        if tokenizer is not None:
            return [tokenizer.decode(r) for r in rows]
        out = []
        for r in rows:
            parts = []
            for i in r:
                if i == END_IDX:
                    break
                if i not in (PAD_IDX, START_IDX):
                    parts.append(_LEGACY_VOCAB[i] if 0 <= i < len(_LEGACY_VOCAB) else "")
            out.append("".join(parts))
        return out

    def debug_tap(self, what: int) -> torch.Tensor:
        """Engine-internal fp32 state of the last executed step of the last engine call (tests only):
        0 hidden [B,d], 1 raw logits [B,V], 2 type logits [B,5], 3 stop logit [B]."""
        L = _lib.lib()
        B = self._last_B
        cols = {0: self.d_model, 1: self.vocab_size, 2: 8, 3: 1}[what]
        out = torch.empty((B, cols), dtype=torch.float32, device=self._engine_device)
        with torch.cuda.device(self._engine_device):
            _lib.check(L.scv_decoder_debug_read(self._engine, what, _lib.ptr(out), out.numel(), _lib.current_stream()),
                       "debug_read")
        torch.cuda.synchronize(self._engine_device)
        return out[:, :5] if what == 2 else (out[:, 0] if what == 3 else out)
