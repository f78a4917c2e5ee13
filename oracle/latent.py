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

