"""Chemistry-constraint rewards of a rollout on the device.

Mirror of the reference's interface for this path, `superconductor.losses.constraint_rewards`
(src/superconductor/losses/constraint_rewards.py): `VocabConfig` (:29-56), `make_v13_vocab_config` (:59-74),
`set_vocab_config` (:80-84), `ConstraintRewardConfig` (:132-149), `FamilyConstraintConfig` (:152-167), the five rule
functions (:270-626) and `compute_constraint_rewards` (:629-676), same names, arguments and module-level "active
vocabulary" convention, so the call sites after each rollout (scripts/train_v12_clean.py:2754-2766, 2990-3007) work
unchanged.  The reference copies the tokens to the host and walks every row in Python; this calls ONE kernel
(csrc/constraints.cu, thread per row over rows staged in shared memory) through the C ABI `scv_constraint_rewards`.
No CPU fallback: tensors must be on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


@dataclass
class VocabConfig:
    """Token layout; defaults = the pre-V13 character-level vocabulary (same fields as the reference's VocabConfig)."""
    element_start: int = 20
    element_end: int = 137
    digit_start: int = 138
    digit_end: int = 147
    lparen_idx: int = 4
    rparen_idx: int = 5
    slash_idx: int = 16
    pad_idx: int = 0
    end_idx: int = 2
    use_semantic_fractions: bool = False
    fraction_token_start: int = 0
    fraction_values: Optional[torch.Tensor] = None

    def elem_idx(self, z: int) -> int:
        return self.element_start - 1 + z


def make_v13_vocab_config(fraction_token_start: int = 143, fraction_values: Optional[torch.Tensor] = None) -> VocabConfig:
    return VocabConfig(element_start=5, element_end=122, digit_start=123, digit_end=142, lparen_idx=-1, rparen_idx=-1,
                       slash_idx=-1, use_semantic_fractions=True, fraction_token_start=fraction_token_start,
                       fraction_values=fraction_values)


_active_vocab: VocabConfig = VocabConfig()


def set_vocab_config(config: VocabConfig):
    """Active vocabulary of the rule functions (module-level, like the reference)."""
    global _active_vocab
    _active_vocab = config


@dataclass
class ConstraintRewardConfig:
    a1_enabled: bool = True
    a1_penalty: float = -50.0
    a2_enabled: bool = True
    a2_penalty_per_violation: float = -5.0
    a4_enabled: bool = True
    a4_penalty: float = -10.0
    a7_enabled: bool = True
    a7_penalty: float = -30.0


@dataclass
class FamilyConstraintConfig:
    enabled: bool = True
    confidence_threshold: float = 0.8
    b1_penalty: float = -40.0
    b2_penalty: float = -40.0
    b3_penalty: float = -40.0
    b4_penalty: float = -30.0
    b5_penalty: float = -30.0
    b6_penalty: float = -30.0
    b7_penalty: float = -30.0
    b8_penalty: float = -30.0


def _run(sampled_tokens, mask, a=None, family_predictions=None, family_config=None) -> torch.Tensor:
    """One launch with the rules in `a` (ConstraintRewardConfig or None = none of A1-A7) and, when both are given
    and enabled, the family rules."""
    _lib.require_cuda(sampled_tokens, "sampled_tokens")
    dev = sampled_tokens.device
    if sampled_tokens.dim() != 2 or mask.shape != sampled_tokens.shape:
        raise RuntimeError(f"sampled_tokens {tuple(sampled_tokens.shape)} and mask {tuple(mask.shape)} must be the same "
                           f"[batch, seq_len]")
    B, L = sampled_tokens.shape
    out = torch.zeros(B, dtype=torch.float32, device=dev)
    if B == 0:
        return out
    v = _active_vocab
    c = _lib.ConstraintConfig()
    for n in ("element_start", "element_end", "digit_start", "digit_end", "lparen_idx", "rparen_idx", "slash_idx",
              "pad_idx", "end_idx", "fraction_token_start"):
        setattr(c, n, int(getattr(v, n)))
    c.use_semantic_fractions = int(bool(v.use_semantic_fractions))
    for n in ("a1", "a2", "a4", "a7"):
        setattr(c, n + "_enabled", int(bool(a is not None and getattr(a, n + "_enabled"))))
    if a is not None:
        c.a1_penalty, c.a2_penalty_per_violation = float(a.a1_penalty), float(a.a2_penalty_per_violation)
        c.a4_penalty, c.a7_penalty = float(a.a4_penalty), float(a.a7_penalty)
    fam = None
    if family_predictions is not None and family_config is not None and family_config.enabled:
        c.family_enabled = 1
        c.confidence_threshold = float(family_config.confidence_threshold)
        for i in range(8):
            c.b_penalty[i] = float(getattr(family_config, f"b{i + 1}_penalty"))
        fam = family_predictions.detach().to(device=dev, dtype=torch.float32).contiguous()
        if fam.dim() != 2 or fam.shape[0] != B:
            raise RuntimeError(f"family_predictions {tuple(fam.shape)} must be [batch={B}, n_families]")
    s = sampled_tokens.to(torch.int64).contiguous()
    m = mask.to(device=dev).bool().to(torch.uint8).contiguous()
    fv = None
    if v.fraction_values is not None:
        fv = v.fraction_values.to(device=dev, dtype=torch.float32).contiguous()
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scv_constraint_rewards(
            s.data_ptr(), m.data_ptr(), B, L, L, C.byref(c), fv.data_ptr() if fv is not None else None,
            fv.numel() if fv is not None else 0, fam.data_ptr() if fam is not None else None,
            fam.shape[1] if fam is not None else 0, out.data_ptr(), C.c_void_p(stream)))
    return out


def _only(**kw) -> ConstraintRewardConfig:
    base = dict(a1_enabled=False, a2_enabled=False, a4_enabled=False, a7_enabled=False)
    base.update(kw)
    return ConstraintRewardConfig(**base)


@torch.no_grad()
def compute_duplicate_element_penalty(sampled_tokens, mask, penalty: float = -50.0) -> torch.Tensor:
    """A1 (:270-303)."""
    return _run(sampled_tokens, mask, _only(a1_enabled=True, a1_penalty=penalty))


@torch.no_grad()
def compute_gcd_canonicality_penalty(sampled_tokens, mask, penalty_per_violation: float = -5.0) -> torch.Tensor:
    """A2 (:306-379); zero with semantic fraction tokens."""
    return _run(sampled_tokens, mask, _only(a2_enabled=True, a2_penalty_per_violation=penalty_per_violation))


@torch.no_grad()
def compute_stoich_normalization_penalty(sampled_tokens, mask, penalty: float = -10.0) -> torch.Tensor:
    """A4 (:382-459)."""
    return _run(sampled_tokens, mask, _only(a4_enabled=True, a4_penalty=penalty))


@torch.no_grad()
def compute_impossible_element_penalty(sampled_tokens, mask, penalty: float = -30.0) -> torch.Tensor:
    """A7 (:462-507)."""
    return _run(sampled_tokens, mask, _only(a7_enabled=True, a7_penalty=penalty))


@torch.no_grad()
def compute_family_constraint_rewards(sampled_tokens, mask, family_predictions, config: FamilyConstraintConfig) -> torch.Tensor:
    """B1-B8 (:510-626)."""
    return _run(sampled_tokens, mask, None, family_predictions, config)


@torch.no_grad()
def compute_constraint_rewards(sampled_tokens: torch.Tensor, mask: torch.Tensor, config: ConstraintRewardConfig,
                               family_predictions: Optional[torch.Tensor] = None,
                               family_config: Optional[FamilyConstraintConfig] = None) -> torch.Tensor:
    """Sum of the enabled penalties per row, float32 [batch] (:629-676)."""
    return _run(sampled_tokens, mask, config, family_predictions, family_config)
