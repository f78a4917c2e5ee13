"""CPU restatement of the token-level rollout reward, `compute_reward_gpu_native`
(reference: src/superconductor/losses/reward_gpu_native.py:448-722 and its helpers :144-445), called right after every
RLOO / SCST rollout (scripts/train_v12_clean.py:2745-2752, 2829-2836, 2942-2950).  SURVEY.md section 8 row f1.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
Pinned by tests/golden/reward.pt, produced by the reference function itself (tests/golden/make_golden_reward.py).

Written row by row (one Python loop over the batch, float32 arithmetic in the reference's operation order) instead of
the reference's whole-batch tensor expressions: the CUDA kernel walks rows the same way.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

# pre-V13 vocabulary ids used by the digit-level fraction penalties (reward_gpu_native.py:34-39)
LPAREN_IDX, RPAREN_IDX, SLASH_IDX, DIGIT_START, DIGIT_END = 4, 5, 16, 138, 147


@dataclass
class RewardConfig:
    """Field names and defaults of GPURewardConfig (:42-79) + GPURewardConfigV14 (:82-131).  `v14` says which of the
    two classes the caller passed (the reference branches on isinstance, :564)."""
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
    # GPURewardConfigV14
    v14: bool = False
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


F = np.float32


def _first_true(flags: np.ndarray, fallback: float) -> F:
    """torch.where(x.any(), x.float().argmax(), fallback) (:506-515)."""
    idx = np.flatnonzero(flags)
    return F(idx[0]) if idx.size else F(fallback)


def _continuous(n_correct: F, n_total: F, cfg: RewardConfig) -> F:
    """_compute_continuous_reward (:406-445)."""
    ratio = F(n_correct) / max(F(n_total), F(1.0))
    ratio = min(max(ratio, F(0.0)), F(1.0))
    sharp = cfg.phase3_sharpness if (cfg.use_phased_curriculum and cfg.reward_phase >= 3) else cfg.sharpness
    return F(cfg.max_reward) * F(np.power(F(ratio), F(sharp), dtype=np.float32))


def _row_sum(x: np.ndarray) -> F:
    return F(np.sum(x.astype(np.float32), dtype=np.float32))


def compute_reward(sampled: np.ndarray, target: np.ndarray, mask: np.ndarray, cfg: Optional[RewardConfig] = None,
                   end_idx: int = 2, use_semantic_fractions: bool = False, fraction_token_start: int = 0,
                   fraction_values: Optional[np.ndarray] = None) -> np.ndarray:
    """sampled, target: int [B, L]; mask: bool [B, L]; returns float32 [B]."""
    cfg = cfg or RewardConfig()
    sampled, target, mask = np.asarray(sampled), np.asarray(target), np.asarray(mask).astype(bool)
    B, L = sampled.shape
    pos = np.arange(L)
    semantic = use_semantic_fractions and fraction_values is not None
    out = np.zeros(B, dtype=np.float32)
    for b in range(B):
        s, t, m = sampled[b], target[b], mask[b]
        eq = s == t
        mism = (~eq) & m
        n_matches, n_mism, n_valid = F((eq & m).sum()), F(mism.sum()), F(m.sum())
        exact = n_mism == 0
        s_end = (s == end_idx) & m
        s_end_pos = _first_true(s_end, n_valid)                                  # (:506-515)
        t_end_pos = _first_true((t == end_idx) & m, n_valid)
        length_diff = F(abs(s_end_pos - t_end_pos))
        # ---- fraction penalty (:519-557)
        if semantic:                                                             # compute_fraction_value_penalty (:279-342)
            fv = np.asarray(fraction_values, dtype=np.float32)
            frac_mis = (~eq) & (t >= fraction_token_start) & m
            sv, tv = fv[np.clip(s, 0, fv.shape[0] - 1)], fv[np.clip(t, 0, fv.shape[0] - 1)]
            scale = F(1.0) + F(cfg.semantic_digit_scale) * np.minimum(np.abs(sv - tv), F(20.0)) / F(20.0)
            frac_pen = _row_sum(frac_mis.astype(np.float32) * F(cfg.fraction_digit_penalty) * scale)
        else:                                                                    # digit-level penalties (:144-276)
            lp, rp = (t == LPAREN_IDX) & m, (t == RPAREN_IDX) & m
            depth = np.cumsum(lp.astype(np.int64) - rp.astype(np.int64))
            in_frac = ((depth > 0) | lp) & m
            structure = ((t == LPAREN_IDX) | (t == RPAREN_IDX) | (t == SLASH_IDX)) & m
            structure_errors = F((mism & structure).sum())
            t_digit = (t >= DIGIT_START) & (t <= DIGIT_END)
            if cfg.use_semantic_digit_penalty:
                dm = (~eq) & t_digit & in_frac & m
                sd = np.clip(s - DIGIT_START, 0, 9).astype(np.float32)
                td = np.clip(t - DIGIT_START, 0, 9).astype(np.float32)
                scale = F(1.0) + F(cfg.semantic_digit_scale) * np.abs(sd - td) / F(9.0)
                digit_pen = _row_sum(dm.astype(np.float32) * F(cfg.fraction_digit_penalty) * scale)
            else:
                digit_pen = F((mism & t_digit & m & in_frac).sum()) * F(cfg.fraction_digit_penalty)
            frac_pen = F(digit_pen + structure_errors * F(cfg.fraction_structure_penalty))
        # ---- length-only error: the whole target prefix is right, the sample runs on (:576-588, 667-685)
        before_t_end = pos < int(t_end_pos)
        prefix_ok = bool(np.all(eq | ~before_t_end | ~m))
        length_only = prefix_ok and (s_end_pos > t_end_pos) and not exact
        extra = max(F(s_end_pos - t_end_pos), F(0.0))
        length_only_reward = max(F(cfg.length_only_base_reward) - extra * F(cfg.length_only_per_extra), F(cfg.length_only_floor))
        length_pen = F(length_diff * F(cfg.length_mismatch_penalty))
        r = F(0.0)
        if exact:
            r = F(cfg.exact_match)
        if cfg.v14 and cfg.use_continuous_reward:                                # (:564-660)
            if length_only:
                r = length_only_reward
            before_s_end = pos < int(s_end_pos)
            prefix2_ok = bool(np.all(eq | ~before_s_end | ~m))
            too_short = prefix2_ok and (s_end_pos < t_end_pos) and bool(s_end.any()) and not exact and not length_only
            if too_short:
                missing = max(F(t_end_pos - s_end_pos), F(0.0))
                r = max(F(cfg.too_short_base_reward) - missing * F(cfg.too_short_per_missing), F(cfg.too_short_floor))
            if not exact and not length_only and not too_short:
                if cfg.use_phased_curriculum and cfg.reward_phase < 3:
                    is_el = (t >= cfg.v14_element_start) & (t <= cfg.v14_element_end)
                    if cfg.reward_phase == 1:
                        pm = is_el & m
                    elif cfg.reward_phase == 2:
                        pm = (is_el | ((t >= cfg.v14_integer_start) & (t <= cfg.v14_integer_end)) | (t >= cfg.v14_fraction_start)) & m
                    else:
                        pm = m
                    base = _continuous(F((eq & pm).sum()), F(pm.sum()), cfg)
                else:
                    content_len = max(F(t_end_pos + F(1.0)), F(1.0))
                    base = _continuous(F((eq & (pos <= int(t_end_pos)) & m).sum()), content_len, cfg)
                # _compute_token_type_penalties (:345-403), classified by the TARGET token
                is_el = (t >= cfg.v14_element_start) & (t <= cfg.v14_element_end) & mism
                is_int = (t >= cfg.v14_integer_start) & (t <= cfg.v14_integer_end) & mism
                is_fr = (t >= cfg.v14_fraction_start) & mism
                is_sp = mism & ~is_el & ~is_int & ~is_fr
                tp = F(F(is_el.sum()) * F(cfg.element_error_penalty) + F(is_int.sum()) * F(cfg.integer_error_penalty)
                       + F(is_sp.sum()) * F(cfg.special_error_penalty))
                if not semantic:
                    tp = F(tp + F(is_fr.sum()) * F(cfg.fraction_error_penalty))
                r = max(F(F(F(base + tp) + frac_pen) + length_pen), F(-100.0))
        else:                                                                    # tiered rewards (:662-722)
            if length_only:
                r = length_only_reward
            not_handled = not exact and not length_only
            for k, bonus in ((1, cfg.near_exact_1), (2, cfg.near_exact_2), (3, cfg.near_exact_3)):
                if not_handled and n_mism == k:
                    r = F(F(F(bonus) + frac_pen) + length_pen)
            if not_handled and n_mism > 3:
                tr = F(n_matches * F(cfg.token_correct) + n_mism * F(cfg.token_penalty))
                tr = F(tr + length_diff * F(cfg.length_mismatch_penalty))
                tr = F(tr + frac_pen)
                r = min(max(tr, F(-100.0)), F(5.0))
        out[b] = r
    return out
