"""CPU restatement of the chemistry-constraint rewards added to every rollout reward, `compute_constraint_rewards`
(reference: src/superconductor/losses/constraint_rewards.py:629-676 and the rules it aggregates: A1 duplicates
:270-303, A2 fraction canonicality :306-379, A4 reducible stoichiometry :382-459, A7 impossible combinations :462-507,
B1-B8 family rules :510-626, the formula parser :172-267, the vocabulary layouts :29-77).  Call sites:
scripts/train_v12_clean.py:2754-2766 (RLOO) and :2990-3007 (SCST).  SURVEY.md section 8 row f1.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
Pinned by tests/golden/constraints.pt, produced by the reference functions themselves
(tests/golden/make_golden_constraints.py).

The reference walks each row with Python loops on the host; so does this file, organised differently: one token
cursor (`Row`) shared by the four scans, rules as small functions over the parsed composition.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class Vocab:
    """Token layout (:29-77).  Defaults = the pre-V13 character-level vocabulary; `v13()` = semantic fractions."""
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
    fraction_values: Optional[np.ndarray] = None

    @staticmethod
    def v13(fraction_token_start: int = 143, fraction_values=None) -> "Vocab":
        return Vocab(5, 122, 123, 142, -1, -1, -1, 0, 2, True, fraction_token_start, fraction_values)

    def z(self, atomic_number: int) -> int:            # token id of an element (:51-55)
        return self.element_start - 1 + atomic_number

    def is_element(self, t: int) -> bool:
        return self.element_start <= t <= self.element_end

    def is_digit(self, t: int) -> bool:
        return self.digit_start <= t <= self.digit_end


@dataclass
class Rules:
    """ConstraintRewardConfig (:132-149) + FamilyConstraintConfig (:152-167)."""
    a1_enabled: bool = True
    a1_penalty: float = -50.0
    a2_enabled: bool = True
    a2_penalty_per_violation: float = -5.0
    a4_enabled: bool = True
    a4_penalty: float = -10.0
    a7_enabled: bool = True
    a7_penalty: float = -30.0
    family_enabled: bool = True
    confidence_threshold: float = 0.8
    b1_penalty: float = -40.0
    b2_penalty: float = -40.0
    b3_penalty: float = -40.0
    b4_penalty: float = -30.0
    b5_penalty: float = -30.0
    b6_penalty: float = -30.0
    b7_penalty: float = -30.0
    b8_penalty: float = -30.0


def _digits_value(ds: List[int]) -> int:
    v = 0
    for d in ds:
        v = 10 * v + d
    return v


def parse_composition(tok: np.ndarray, msk: np.ndarray, v: Vocab) -> Tuple[List[int], Dict[int, float]]:
    """Elements in order of appearance and the amount attached to each (last occurrence wins), read up to the first
    END / unmasked position (:172-267).  Amount = 1 without a subscript; V13: an integer token (value id - start + 1)
    or a fraction token's table value; pre-V13: `( digits / digits )` or a run of digits."""
    L, i = len(tok), 0
    elems: List[int] = []
    amount: Dict[int, float] = {}
    while i < L and msk[i] and tok[i] != v.end_idx:
        t = int(tok[i])
        if not v.is_element(t):
            i += 1
            continue
        elems.append(t)
        a, j = 1.0, i + 1
        if not (j < L and msk[j]):
            amount[t] = a
            i += 1
            continue
        n = int(tok[j])
        if v.use_semantic_fractions:
            if v.is_digit(n):
                a, j = float(n - v.digit_start + 1), j + 1
            elif n >= v.fraction_token_start and v.fraction_values is not None:
                if n < len(v.fraction_values):
                    a = float(v.fraction_values[n])
                j += 1
        elif n == v.lparen_idx:
            j += 1
            num, den, in_num = [], [], True
            while j < L:                                   # (the reference does not look at the mask in here)
                c = int(tok[j])
                if c == v.slash_idx:
                    in_num = False
                elif c == v.rparen_idx:
                    j += 1
                    break
                elif v.is_digit(c):
                    (num if in_num else den).append(c - v.digit_start)
                else:
                    break
                j += 1
            if num and den and _digits_value(den) > 0:
                a = _digits_value(num) / _digits_value(den)
        elif v.is_digit(n):
            ds = []
            while j < L and v.is_digit(int(tok[j])):
                ds.append(int(tok[j]) - v.digit_start)
                j += 1
            a = float(_digits_value(ds))
        amount[t] = a
        i = j
    return elems, amount


def a1_duplicates(tok, msk, v: Vocab) -> bool:
    """Any element id at two masked positions of the row, END or not (:270-303)."""
    seen = set()
    for t, m in zip(tok, msk):
        if m and v.is_element(int(t)):
            if int(t) in seen:
                return True
            seen.add(int(t))
    return False


def a2_violations(tok, msk, v: Vocab) -> int:
    """Closed `( num / den )` groups with gcd(num, den) > 1; pre-V13 only (:306-379)."""
    if v.use_semantic_fractions:
        return 0
    L, i, bad = len(tok), 0, 0
    while i < L and msk[i]:
        if int(tok[i]) != v.lparen_idx:
            i += 1
            continue
        j, num, den, in_num, closed = i + 1, [], [], True, False
        while j < L and msk[j]:
            c = int(tok[j])
            if c == v.slash_idx:
                in_num = False
            elif c == v.rparen_idx:
                closed, j = True, j + 1
                break
            elif v.is_digit(c):
                (num if in_num else den).append(c - v.digit_start)
            else:
                break
            j += 1
        if closed and num and den and _digits_value(den) > 0 and math.gcd(_digits_value(num), _digits_value(den)) > 1:
            bad += 1
        i = j if j > i + 1 else i + 1
    return bad


def a4_reducible(tok, msk, v: Vocab) -> bool:
    """All-integer formula (>= 2 elements) whose subscripts share a factor (:382-459)."""
    L, i, subs = len(tok), 0, []
    while i < L and msk[i] and tok[i] != v.end_idx:
        t = int(tok[i])
        if (v.use_semantic_fractions and t >= v.fraction_token_start) or (not v.use_semantic_fractions and t == v.lparen_idx):
            return False
        if not v.is_element(t):
            i += 1
            continue
        j = i + 1
        if v.use_semantic_fractions:
            if j < L and msk[j] and v.is_digit(int(tok[j])):
                subs.append(int(tok[j]) - v.digit_start + 1)
                j += 1
            else:
                subs.append(1)
        else:
            ds = []
            while j < L and msk[j] and v.is_digit(int(tok[j])):
                ds.append(int(tok[j]) - v.digit_start)
                j += 1
            subs.append(_digits_value(ds) if ds else 1)
        i = j
    if len(subs) < 2:
        return False
    g = subs[0]
    for s in subs[1:]:
        g = math.gcd(g, s)
    return g > 1


def a7_impossible(elems, amount, v: Vocab) -> bool:
    """F together with Tl, or Mn / Fe / Co / Ni above 2 % and above half the Cu amount next to Cu (:462-507)."""
    present = set(elems)
    if v.z(9) in present and v.z(81) in present:
        return True
    cu = amount.get(v.z(29), 0.0) if v.z(29) in present else 0.0
    if v.z(29) in present and cu > 0:
        for zz in (25, 26, 27, 28):
            if v.z(zz) in present:
                f = amount.get(v.z(zz), 0.0)
                if f > 0.02 and f > 0.5 * cu:
                    return True
    return False


def family_penalty(elems, amount, fam: int, r: Rules, v: Vocab) -> float:
    """B1-B8 (:549-620) for the predicted family `fam` (index into the 14 composed family probabilities)."""
    present = set(elems)
    get = lambda zz: amount.get(v.z(zz), 0.0)
    has = lambda zz: v.z(zz) in present
    magnetic_over = lambda lim: any(has(zz) and get(zz) > lim for zz in (25, 26, 27, 28))
    p = 0.0
    if fam == 2:                                           # YBCO: oxygen content
        if 0 < get(8) < 6.35:
            p += r.b1_penalty
    elif fam == 3:                                         # LSCO: Sr doping window
        if has(38) and (get(38) < 0.055 or get(38) > 0.27):
            p += r.b2_penalty
    elif fam == 4:                                         # BSCCO: Ca - (Cu - 1)
        if has(20) and has(29) and abs(get(20) - (get(29) - 1)) > 0.3:
            p += r.b3_penalty
    elif fam == 6:                                         # Hg cuprates: V on the Hg site
        if get(23) > 0.30:
            p += r.b4_penalty
    elif fam == 5:                                         # Tl cuprates: V, Li, magnetic 3d
        if get(23) > 0.30:
            p += r.b5_penalty
        if get(3) > 0.10:
            p += r.b5_penalty
        if magnetic_over(0.10):
            p += r.b5_penalty
    elif fam == 8:                                         # iron pnictides: oxygen
        if has(8) and get(8) < 0.7 and get(8) != 1.0:
            p += r.b6_penalty
    elif fam == 10:                                        # MgB2: C, Al, magnetic 3d
        if get(6) > 0.125:
            p += r.b7_penalty
        if get(13) > 0.50:
            p += r.b7_penalty
        if magnetic_over(0.05):
            p += r.b7_penalty
    elif fam == 1:                                         # A15: (Nb + V) : (Sn + Al + Si + Ge) = 3 : 1 within 10 %
        a_tot = sum(get(zz) for zz in (41, 23) if has(zz))
        b_tot = sum(get(zz) for zz in (50, 13, 14, 32) if has(zz))
        if a_tot > 0 and b_tot > 0 and abs(a_tot / b_tot - 3.0) > 0.3:
            p += r.b8_penalty
    return p


def compute_constraint_rewards(sampled: np.ndarray, mask: np.ndarray, rules: Rules, vocab: Vocab,
                               family_predictions: Optional[np.ndarray] = None) -> np.ndarray:
    """float32 [B]: the enabled penalties added in the reference's order (:645-676)."""
    sampled, mask = np.asarray(sampled), np.asarray(mask).astype(bool)
    B = sampled.shape[0]
    out = np.zeros(B, dtype=np.float32)
    for b in range(B):
        tok, msk = sampled[b], mask[b]
        elems, amount = parse_composition(tok, msk, vocab)
        total = np.float32(0.0)
        if rules.a1_enabled:
            total = total + np.float32(float(a1_duplicates(tok, msk, vocab)) * rules.a1_penalty)
        if rules.a2_enabled:
            total = total + np.float32(a2_violations(tok, msk, vocab) * rules.a2_penalty_per_violation)
        if rules.a4_enabled:
            total = total + np.float32(rules.a4_penalty if a4_reducible(tok, msk, vocab) else 0.0)
        if rules.a7_enabled:
            total = total + np.float32(rules.a7_penalty if a7_impossible(elems, amount, vocab) else 0.0)
        if family_predictions is not None and rules.family_enabled:
            probs = np.asarray(family_predictions[b], dtype=np.float32)
            fam = int(np.argmax(probs))
            if not float(probs[fam]) < rules.confidence_threshold:
                p = family_penalty(elems, amount, fam, rules, vocab)
                if p < 0:
                    total = total + np.float32(p)
        out[b] = total
    return out
