"""Latent-space candidate generation: SLERP walks between anchor latents, decoded on the engine.

Reference call convention: scripts/holdout/holdout_search.py:128-146 (slerp), :392-436 (decode_z_batch),
and the V14.3-complete variant in notebooks/generative_evaluation.ipynb cells 12/14/16
(stoich_pred = fraction_head(z), heads_pred = _build_heads_pred(z), type masks, stop head).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib


@torch.no_grad()
def slerp(z1: torch.Tensor, z2: torch.Tensor, t) -> torch.Tensor:
    """Same arithmetic as the reference's slerp for [N, D] rows (t: float or [N, 1])."""
    _lib.require_cuda(z1, "z1")
    n, d = z1.shape
    anchors = torch.cat([z1, z2.expand_as(z1)], dim=0).to(torch.float32).contiguous()
    i1 = torch.arange(n, dtype=torch.int32, device=z1.device)
    i2 = i1 + n
    tt = torch.as_tensor(t, dtype=torch.float32, device=z1.device).reshape(-1).expand(n).contiguous()
    return slerp_rows(anchors, i1, i2, tt)


@torch.no_grad()
def slerp_rows(anchors: torch.Tensor, i1: torch.Tensor, i2: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """out[r] = slerp(anchors[i1[r]], anchors[i2[r]], t[r]) without materialising the gathered rows."""
    L = _lib.lib()
    _lib.require_cuda(anchors, "anchors")
    dev = anchors.device
    anchors = anchors.to(torch.float32).contiguous()
    i1 = i1.to(device=dev, dtype=torch.int32).contiguous()
    i2 = i2.to(device=dev, dtype=torch.int32).contiguous()
    t = t.to(device=dev, dtype=torch.float32).contiguous()
    n = i1.numel()
    out = torch.empty((n, anchors.shape[1]), dtype=torch.float32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.scv_slerp_rows(_lib.ptr(anchors), anchors.shape[1], _lib.ptr(i1), _lib.ptr(i2), _lib.ptr(t), n,
                                    _lib.ptr(out), _lib.ptr(flag), _lib.current_stream()), "slerp_rows")
    return out


@torch.no_grad()
def decode_z_batch(encoder, decoder, z: torch.Tensor, temperature: float = 0.001, type_masks=None,
                   stop_boost: float = 10.0, hard_stop_threshold: float = 0.8, use_heads: bool = True,
                   max_len: Optional[int] = None, return_log_probs: bool = False):
    """z [N, latent] -> tokens [N, L] (+ log-probs): conditioning from z alone, then KV-cache decode.

    ``use_heads=True`` is the notebook's 24-memory-token pipeline; ``False`` the scripts' 20-token one
    (no heads, no masks, no stop head: pass type_masks=None, stop_boost=0)."""
    if use_heads:
        stoich, heads = encoder.conditioning(z)
    else:
        stoich, heads = encoder.heads_from_latent(z)["stoich_pred"], None
    toks, lps, _ = decoder.generate_with_kv_cache(
        z=z, stoich_pred=stoich, temperature=temperature, max_len=max_len, heads_pred=heads, type_masks=type_masks,
        stop_boost=stop_boost, hard_stop_threshold=hard_stop_threshold, return_log_probs=return_log_probs)
    return (toks, lps) if return_log_probs else toks


@torch.no_grad()
def unique_sequences(tokens: torch.Tensor):
    """Group identical generated rows on the device.

    tokens [N, L] int64 on CUDA -> (unique_rows int16 [U, L] (ids up to the first END, PAD after), inverse int64 [N]
    (row -> index into unique_rows), counts int64 [U], lengths int32 [N]).  A row's formula depends only on the ids
    before its first END (tokenizer decode), so rows are canonicalised and hashed by one kernel
    (scv_tokens_canonical_hash), grouped by hash with torch.unique, and every row is then compared with its group's
    representative; a hash collision (never observed) falls back to an exact row-wise unique."""
    L = _lib.lib()
    _lib.require_cuda(tokens, "tokens")
    dev = tokens.device
    tokens = tokens.to(torch.int64).contiguous()
    n, ln = tokens.shape
    canon = torch.empty((n, ln), dtype=torch.int16, device=dev)
    hsh = torch.empty((n,), dtype=torch.int64, device=dev)        # 64-bit pattern; only equality matters
    lengths = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.scv_tokens_canonical_hash(_lib.ptr(tokens), n, ln, _lib.ptr(canon), _lib.ptr(hsh), _lib.ptr(lengths),
                                               _lib.current_stream()), "tokens_canonical_hash")
    _, inverse, counts = torch.unique(hsh, return_inverse=True, return_counts=True)
    first = torch.full((counts.numel(),), n, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inverse, torch.arange(n, device=dev), reduce="amin")
    uniq = canon[first]
    if not bool((canon == uniq[inverse]).all()):                  # two different rows with one hash
        uniq, inverse, counts = torch.unique(canon, dim=0, return_inverse=True, return_counts=True)
    return uniq, inverse, counts, lengths


@torch.no_grad()
def decode_unique(tokenizer, tokens: torch.Tensor):
    """tokens [N, L] -> (formulas of the U distinct rows, inverse [N], counts [U]): strings are built once per distinct
    candidate instead of once per row (the reference loops over every row, scripts/holdout/holdout_search.py:88-99)."""
    uniq, inverse, counts, _ = unique_sequences(tokens)
    return tokenizer.decode_batch(uniq.to(torch.int64)), inverse, counts



# ------------------------------------------------------------------ candidate scoring (second half)
_ELEMENT_PATTERN = re.compile(r'([A-Z][a-z]?)(?:\((\d+)/(\d+)\)|\((\d+)\)|(\d+(?:\.\d+)?))?')


def parse_formula_elements(formula: str) -> Dict[str, float]:
    """{element: summed amount} of a formula string, same result as the reference's parse_formula_elements
    (scripts/holdout/holdout_search.py:109-125): `El(p/q)` -> p / q, `El(n)`, `Eln`, `El0.n` -> the number, a bare element
    -> 1, repeated elements add up, anything that makes the reference's parser raise (a zero denominator) -> {}."""
    out: Dict[str, float] = {}
    for m in _ELEMENT_PATTERN.finditer(formula):
        el, num, den, par, plain = m.groups()
        if num and den:
            if int(den) == 0:
                return {}
            val = int(num) / int(den)
        elif par:
            val = float(int(par))
        elif plain:
            val = float(plain)
        else:
            val = 1.0
        out[el] = out.get(el, 0) + val
    return out


def composition_matrix(formulas: List[str], elements: Optional[List[str]] = None):
    """(float64 [N, E] host tensor, element column names): parsed amounts per formula, -1 where the formula does not
    contain the element (the layout scv_element_similarity takes)."""
    parsed = [parse_formula_elements(f) for f in formulas]
    if elements is None:
        elements = sorted({e for p in parsed for e in p})
    col = {e: i for i, e in enumerate(elements)}
    m = torch.full((len(formulas), max(len(elements), 1)), -1.0, dtype=torch.float64)
    for r, p in enumerate(parsed):
        for e, v in p.items():
            m[r, col[e]] = v
    return m, elements


@torch.no_grad()
def element_similarity_matrix(candidates: List[str], targets: List[str], device="cuda") -> torch.Tensor:
    """similarity[i, j] = element_similarity(candidates[i], targets[j]) for every pair, on the device (one kernel): the
    holdout search scores each of its ~31,000 candidates per target against the target formula
    (scripts/holdout/holdout_search.py:149-182 inside the candidate loops); float64 [len(candidates), len(targets)]."""
    L = _lib.lib()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.EngineError("element_similarity_matrix runs on a CUDA device (sm_100a); this package has no CPU fallback")
    if not candidates or not targets:
        return torch.zeros((len(candidates), len(targets)), dtype=torch.float64, device=dev)
    parsed_elements = sorted({e for f in list(candidates) + list(targets) for e in parse_formula_elements(f)})
    a, _ = composition_matrix(candidates, parsed_elements)
    b, _ = composition_matrix(targets, parsed_elements)
    a, b = a.to(dev).contiguous(), b.to(dev).contiguous()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.scv_element_similarity(_lib.ptr(a), _lib.ptr(b), a.shape[0], b.shape[0], a.shape[1], _lib.ptr(out),
                                            _lib.current_stream()), "element_similarity")
    return out


def element_similarity(formula_a: str, formula_b: str, device="cuda") -> float:
    """The reference's scalar function (same name and arguments) through the device kernel."""
    return float(element_similarity_matrix([formula_a], [formula_b], device)[0, 0])
