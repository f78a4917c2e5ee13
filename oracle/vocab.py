"""Token-id layout and [5, V] type masks of FractionAwareTokenizer, restated by id ranges.

TEST INFRASTRUCTURE (see oracle/__init__.py).
Follows src/superconductor/tokenizer/fraction_tokenizer.py:130-204 (vocab build: 5 specials,
118 elements, integers 1..20, N fractions, ISO_UNK, isotopes) and :306-338 (get_token_type /
get_type_masks: element 0, integer 1, fraction 2, special 3, EOS 4).
"""
from __future__ import annotations

import torch

N_SPECIAL, PAD_IDX, BOS_IDX, EOS_IDX, UNK_IDX, FRAC_UNK_IDX = 5, 0, 1, 2, 3, 4
TYPE_ELEMENT, TYPE_INTEGER, TYPE_FRACTION, TYPE_SPECIAL, TYPE_EOS = 0, 1, 2, 3, 4
N_TOKEN_TYPES = 5


def vocab_size(n_fractions: int = 4317, n_isotopes: int = 291, n_elements: int = 118, n_integers: int = 20) -> int:
    v = N_SPECIAL + n_elements + n_integers + n_fractions
    return v + 1 + n_isotopes if n_isotopes > 0 else v


def token_type(tid: int, n_fractions: int, n_elements: int = 118, n_integers: int = 20) -> int:
    if tid == EOS_IDX:
        return TYPE_EOS
    if N_SPECIAL <= tid < N_SPECIAL + n_elements:
        return TYPE_ELEMENT
    int0 = N_SPECIAL + n_elements
    if int0 <= tid < int0 + n_integers:
        return TYPE_INTEGER
    if int0 + n_integers <= tid < int0 + n_integers + n_fractions:
        return TYPE_FRACTION
    return TYPE_SPECIAL


def type_masks(n_fractions: int = 4317, n_isotopes: int = 291, n_elements: int = 118,
               n_integers: int = 20) -> torch.Tensor:
    v = vocab_size(n_fractions, n_isotopes, n_elements, n_integers)
    m = torch.zeros(N_TOKEN_TYPES, v, dtype=torch.bool)
    for tid in range(v):
        m[token_type(tid, n_fractions, n_elements, n_integers), tid] = True
    return m


# layout used with the TINY decoder shape (vocab 97): 5 special + 40 elements + 10 integers + 30 fractions
# + ISO_UNK + 11 isotopes
TINY_LAYOUT = dict(n_fractions=30, n_isotopes=11, n_elements=40, n_integers=10)
