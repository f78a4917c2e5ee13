"""Spherical interpolation of latents, restated from scripts/holdout/holdout_search.py:128-146
(identical in notebooks/generative_evaluation.ipynb cell 12).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def slerp(z1: torch.Tensor, z2: torch.Tensor, t) -> torch.Tensor:
    n1, n2 = F.normalize(z1, dim=-1), F.normalize(z2, dim=-1)
    omega = torch.acos(torch.clamp((n1 * n2).sum(dim=-1, keepdim=True), -1.0, 1.0)).clamp(min=1e-6)
    so = torch.sin(omega)
    if so.abs().min() < 1e-6:                       # batch-global fallback to lerp (:137-138)
        return (1 - t) * z1 + t * z2
    mag = (1 - t) * z1.norm(dim=-1, keepdim=True) + t * z2.norm(dim=-1, keepdim=True)
    return (torch.sin((1 - t) * omega) / so * n1 + torch.sin(t * omega) / so * n2) * mag


def unique_formulas(tokens, decode):
    """Reference semantics of candidate post-processing (scripts/holdout/holdout_search.py:88-99 applied to every row,
    then grouped): formulas[i] = decode(row i); returns (distinct formulas in first-occurrence order, index of each
    row's formula in that list, counts).  `decode` is FractionAwareTokenizer.decode (stops at the first END)."""
    seen, order, inverse, counts = {}, [], [], []
    for row in tokens.tolist():
        f = decode(row)
        if f not in seen:
            seen[f] = len(order)
            order.append(f)
            counts.append(0)
        inverse.append(seen[f])
        counts[seen[f]] += 1
    return order, inverse, counts



# ---- candidate scoring (second half) ------------------------------------------------------------------
_ELEMENT_PATTERN = r'([A-Z][a-z]?)(?:\((\d+)/(\d+)\)|\((\d+)\)|(\d+(?:\.\d+)?))?'


def parse_formula_elements(formula: str) -> dict:
    """{element: summed amount} of a formula string, restated from scripts/holdout/holdout_search.py:109-125 over
    data/canonical_ordering.py:110-152 (CanonicalOrderer.parse_formula) and :87-96 (fraction_value): `El(p/q)` -> p / q,
    `El(n)` and `Eln` / `El0.n` -> the number, a bare element -> 1; repeated elements add up; any exception (a zero
    denominator) -> {}.  TEST INFRASTRUCTURE."""
    import re
    try:
        out = {}
        for m in re.finditer(_ELEMENT_PATTERN, formula):
            el = m.group(1)
            if not el:
                continue
            if m.group(2) and m.group(3):
                val = int(m.group(2)) / int(m.group(3))
            elif m.group(4):
                val = int(m.group(4)) / 1
            elif m.group(5):
                try:
                    val = float(m.group(5))
                except ValueError:
                    val = 1.0
            else:
                val = 1.0
            out[el] = out.get(el, 0) + val
        return out
    except Exception:
        return {}


def element_similarity(formula_a: str, formula_b: str) -> float:
    """0.5 * Jaccard(element sets) + 0.5 * sum over shared elements of min(normalised amounts)
    (scripts/holdout/holdout_search.py:149-182)."""
    pa, pb = parse_formula_elements(formula_a), parse_formula_elements(formula_b)
    if not pa or not pb:
        return 0.0
    union, shared = set(pa) | set(pb), set(pa) & set(pb)
    jaccard = len(shared) / len(union)
    frac = 0.0
    if shared:
        ta, tb = sum(pa.values()), sum(pb.values())
        for el in shared:
            frac += min(pa[el] / max(ta, 1e-8), pb[el] / max(tb, 1e-8))
    return 0.5 * jaccard + 0.5 * frac
