"""Token-level rollout reward on the device.

Mirror of the reference's interface for this path, `superconductor.losses.reward_gpu_native`
(src/superconductor/losses/reward_gpu_native.py): `GPURewardConfig` (:42-79), `GPURewardConfigV14` (:82-131),
`get_default_gpu_reward_config` (:134), `get_v14_gpu_reward_config` (:139) and `compute_reward_gpu_native` (:448-722)
with the same argument names, order and meaning, so the call sites after each rollout
(scripts/train_v12_clean.py:2745-2752, 2829-2836, 2942-2950) work unchanged.  The reference evaluates ~60 whole-batch
tensor expressions; this calls ONE kernel (csrc/reward.cu, warp per row) through the C ABI `scv_reward_tokens`.
No CPU fallback: tensors must be on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields
from typing import Optional

import torch

from . import _lib


@dataclass
class GPURewardConfig:
    """Same fields and defaults as the reference's GPURewardConfig."""
    exact_match: float = 100.0
    near_exact_1: float = 50.0
    near_exact_2: float = 25.0
    near_exact_3: float = 10.0
    token_correct: float = 1.0
    token_penalty: float = -0.5
    length_mismatch_penalty: float = -2.0
    fraction_digit_penalty: float = -10.0
    fraction_structure_penalty: float = -5.0
    use_semantic_digit_penalty: bool = True
    semantic_digit_scale: float = 2.0
    length_only_base_reward: float = 50.0
    length_only_per_extra: float = 5.0
    length_only_floor: float = 10.0


@dataclass
class GPURewardConfigV14(GPURewardConfig):
    """Same fields and defaults as the reference's GPURewardConfigV14 (continuous reward, token-type penalties,
    too-short handling, phased curriculum, token-type boundaries of the V13 vocabulary)."""
    use_continuous_reward: bool = True
    max_reward: float = 100.0
    sharpness: float = 4.0
    element_error_penalty: float = -3.0
    integer_error_penalty: float = -1.0
    fraction_error_penalty: float = -0.5
    special_error_penalty: float = -0.5
    too_short_base_reward: float = 50.0
    too_short_per_missing: float = 5.0
    too_short_floor: float = 10.0
    use_phased_curriculum: bool = False
    reward_phase: int = 3
    phase3_sharpness: float = 6.0
    v14_element_start: int = 5
    v14_element_end: int = 122
    v14_integer_start: int = 123
    v14_integer_end: int = 142
    v14_fraction_start: int = 143


def get_default_gpu_reward_config() -> GPURewardConfig:
    return GPURewardConfig()


def get_v14_gpu_reward_config(**kwargs) -> GPURewardConfigV14:
    return GPURewardConfigV14(**kwargs)


def _pack(config: GPURewardConfig) -> "_lib.RewardConfig":
    v14_defaults = GPURewardConfigV14()
    c = _lib.RewardConfig()
    is_v14 = isinstance(config, GPURewardConfigV14)          # the reference branches on the class (:564)
    for name, _ in _lib.RewardConfig._fields_:
        if name == "v14":
            c.v14 = int(is_v14)
            continue
        v = getattr(config, name) if hasattr(config, name) else getattr(v14_defaults, name)
        setattr(c, name, int(v) if isinstance(getattr(v14_defaults, name), (bool, int)) else float(v))
    return c


@torch.no_grad()
def compute_reward_gpu_native(sampled_tokens: torch.Tensor, target_tokens: torch.Tensor, mask: torch.Tensor,
                              config: Optional[GPURewardConfig] = None, pad_idx: int = 0, end_idx: int = 2,
                              use_semantic_fractions: bool = False, fraction_token_start: int = 0,
                              fraction_values: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rewards [batch] float32 for sampled / target token rows [batch, seq_len] under `mask` (reference :448-722)."""
    if config is None:
        config = get_default_gpu_reward_config()
    _lib.require_cuda(sampled_tokens, "sampled_tokens")
    _lib.require_cuda(target_tokens, "target_tokens")
    dev = sampled_tokens.device
    if sampled_tokens.dim() != 2 or sampled_tokens.shape != target_tokens.shape or mask.shape != sampled_tokens.shape:
        raise RuntimeError(f"sampled_tokens {tuple(sampled_tokens.shape)}, target_tokens {tuple(target_tokens.shape)} and "
                           f"mask {tuple(mask.shape)} must be the same [batch, seq_len]")
    B, L = sampled_tokens.shape
    out = torch.empty(B, dtype=torch.float32, device=dev)
    if B == 0:
        return out
    s = sampled_tokens.to(torch.int64).contiguous()
    t = target_tokens.to(device=dev, dtype=torch.int64).contiguous()
    m = mask.to(device=dev).bool().to(torch.uint8).contiguous()
    fv, nfv = None, 0
    if use_semantic_fractions and fraction_values is not None:
        fv = fraction_values.to(device=dev, dtype=torch.float32).contiguous()      # (:312-313)
        nfv = fv.numel()
    cfg = _pack(config)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().scv_reward_tokens(s.data_ptr(), t.data_ptr(), m.data_ptr(), B, L, L, C.byref(cfg), int(end_idx),
                                                int(bool(use_semantic_fractions)), int(fraction_token_start),
                                                fv.data_ptr() if fv is not None else None, nfv, out.data_ptr(),
                                                C.c_void_p(stream)))
    return out
