"""Token-id layout of the reference's FractionAwareTokenizer, table driven.

Same ids as src/superconductor/tokenizer/fraction_tokenizer.py (vocab build :130-204):
  0..4   <PAD> <BOS> <EOS> <UNK> <FRAC_UNK>
  5..122 the 118 element symbols in atomic-number order
  123..142 integers "1".."20"
  143..   one FRAC:p/q token per entry of fraction_vocab.json["fractions"] (file order)
  then   <ISO_UNK> and one ISO:<mass><symbol> token per entry of isotope_vocab.json["isotopes"]
Type classes (:306-338): element 0, integer 1, fraction 2, special 3, EOS 4.

The vocabulary files are data of the reference checkpoint, not part of this package: pass their paths
(or the lists themselves).  ``decode_batch`` turns a [B, L] tensor of generated ids into strings with one
table lookup per token.
"""
from __future__ import annotations

import json
import math
import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

PAD_IDX, BOS_IDX, EOS_IDX, UNK_IDX, FRAC_UNK_IDX = 0, 1, 2, 3, 4
N_SPECIAL, N_ELEMENTS, MAX_INTEGER = 5, 118, 20
TOKEN_TYPE_ELEMENT, TOKEN_TYPE_INTEGER, TOKEN_TYPE_FRACTION, TOKEN_TYPE_SPECIAL, TOKEN_TYPE_EOS = 0, 1, 2, 3, 4
N_TOKEN_TYPES = 5

_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb "
            "Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf "
            "Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og").split()
assert len(_SYMBOLS) == N_ELEMENTS
ELEMENTS = [""] + _SYMBOLS

_SPECIAL_NAMES = ["<PAD>", "<BOS>", "<EOS>", "<UNK>", "<FRAC_UNK>"]
_PIECE = re.compile(r"\{(\d+)\}([A-Z][a-z]?)|\((\d+)/(\d+)\)|([A-Z][a-z]?)|(\d+)")
_ISO = re.compile(r"^(\d+)([A-Z][a-z]?)$")


class FractionAwareTokenizer:
    def __init__(self, fraction_vocab_path: Optional[str] = None, max_len: int = 60,
                 isotope_vocab_path: Optional[str] = None, *, fractions: Optional[Sequence[str]] = None,
                 isotopes: Optional[Sequence[str]] = None):
        self.max_len = max_len
        if fraction_vocab_path is not None:
            with open(fraction_vocab_path) as f:
                fractions = json.load(f)["fractions"]
        if isotope_vocab_path is not None:
            with open(isotope_vocab_path) as f:
                isotopes = json.load(f)["isotopes"]
        self._fraction_list: List[str] = list(fractions or [])
        self._isotope_list: List[str] = list(isotopes or [])
        self._int_offset = N_SPECIAL + N_ELEMENTS
        self._frac_offset = self._int_offset + MAX_INTEGER
        self._iso_unk_idx = self._frac_offset + len(self._fraction_list) if self._isotope_list else None
        self._iso_offset = self._iso_unk_idx + 1 if self._isotope_list else 0
        # id -> token name, id -> surface string used by decode()
        names = list(_SPECIAL_NAMES) + _SYMBOLS + [str(i) for i in range(1, MAX_INTEGER + 1)]
        surface = ["", "", "", "?", "(?/?)"] + _SYMBOLS + [str(i) for i in range(1, MAX_INTEGER + 1)]
        for fr in self._fraction_list:
            names.append("FRAC:" + fr)
            surface.append("(" + fr + ")")
        if self._isotope_list:
            names.append("<ISO_UNK>")
            surface.append("{?}?")
            for iso in self._isotope_list:
                names.append("ISO:" + iso)
                m = _ISO.match(iso)
                surface.append("{%s}%s" % (m.group(1), m.group(2)) if m else "{%s}" % iso)
        self._names, self._surface = names, surface
        self._token_to_id: Dict[str, int] = {n: i for i, n in enumerate(names)}
        self._fraction_to_id = {fr: self._frac_offset + i for i, fr in enumerate(self._fraction_list)}
        self._isotope_to_id = {iso: self._iso_offset + i for i, iso in enumerate(self._isotope_list)}

    # ------------------------------------------------------------------ sizes and ranges
    @property
    def vocab_size(self) -> int:
        return len(self._names)

    pad_idx, bos_idx, eos_idx, unk_idx, frac_unk_idx = PAD_IDX, BOS_IDX, EOS_IDX, UNK_IDX, FRAC_UNK_IDX

    @property
    def n_fraction_tokens(self) -> int:
        return len(self._fraction_list)

    @property
    def n_isotope_tokens(self) -> int:
        return len(self._isotope_list)

    @property
    def fraction_token_start(self) -> int:
        return self._frac_offset

    @property
    def iso_unk_idx(self) -> Optional[int]:
        return self._iso_unk_idx

    @property
    def isotope_token_start(self) -> Optional[int]:
        return self._iso_offset if self._isotope_list else None

    def is_element_token(self, tid: int) -> bool:
        return N_SPECIAL <= tid < N_SPECIAL + N_ELEMENTS

    def is_integer_token(self, tid: int) -> bool:
        return self._int_offset <= tid < self._int_offset + MAX_INTEGER

    def is_fraction_token(self, tid: int) -> bool:
        return self._frac_offset <= tid < self._frac_offset + len(self._fraction_list)

    def is_isotope_token(self, tid: int) -> bool:
        return bool(self._isotope_list) and self._iso_offset <= tid < self._iso_offset + len(self._isotope_list)

    def fraction_token_to_numden(self, tid: int) -> Tuple[int, int]:
        if not self.is_fraction_token(tid):
            raise ValueError(f"Token {tid} is not a fraction token")
        p, q = self._fraction_list[tid - self._frac_offset].split("/")
        return int(p), int(q)

    def fraction_token_to_value(self, tid: int) -> float:
        p, q = self.fraction_token_to_numden(tid)
        return p / q

    def get_token_name(self, tid: int) -> str:
        return self._names[tid] if 0 <= tid < len(self._names) else f"<ID:{tid}>"

    # ------------------------------------------------------------------ token types (V14.3 hard masking)
    def get_token_type(self, tid: int) -> int:
        if tid == EOS_IDX:
            return TOKEN_TYPE_EOS
        if self.is_element_token(tid):
            return TOKEN_TYPE_ELEMENT
        if self.is_integer_token(tid):
            return TOKEN_TYPE_INTEGER
        if self.is_fraction_token(tid):
            return TOKEN_TYPE_FRACTION
        return TOKEN_TYPE_SPECIAL

    def token_type_table(self):
        import torch
        v = self.vocab_size
        lut = torch.full((v,), TOKEN_TYPE_SPECIAL, dtype=torch.long)
        lut[N_SPECIAL:N_SPECIAL + N_ELEMENTS] = TOKEN_TYPE_ELEMENT
        lut[self._int_offset:self._int_offset + MAX_INTEGER] = TOKEN_TYPE_INTEGER
        lut[self._frac_offset:self._frac_offset + len(self._fraction_list)] = TOKEN_TYPE_FRACTION
        lut[EOS_IDX] = TOKEN_TYPE_EOS
        return lut

    def get_type_masks(self, device="cpu"):
        """[5, vocab] bool: mask[type, id] is True when the token belongs to that type class."""
        import torch
        lut = self.token_type_table()
        return (lut.unsqueeze(0) == torch.arange(N_TOKEN_TYPES).unsqueeze(1)).to(device)

    def compute_token_type_targets(self, token_ids):
        lut = self.token_type_table().to(token_ids.device)
        return lut[token_ids.clamp(0, lut.shape[0] - 1)]

    # ------------------------------------------------------------------ text <-> ids
    def encode(self, formula: str, add_bos_eos: bool = True, pad: bool = True) -> List[int]:
        out: List[int] = []
        for m in _PIECE.finditer(formula):
            mass, iso_el, num, den, el, integer = m.groups()
            if mass is not None:
                if not self._isotope_list:      # without an isotope vocabulary "{18}O" reads as 18, O
                    out.append(self._token_to_id.get(mass, UNK_IDX) if 1 <= int(mass) <= MAX_INTEGER else UNK_IDX)
                    out.append(self._token_to_id.get(iso_el, UNK_IDX))
                else:
                    out.append(self._isotope_to_id.get(mass + iso_el, self._iso_unk_idx))
            elif num is not None:
                p, q = int(num), int(den)
                g = math.gcd(p, q)
                out.append(self._fraction_to_id.get(f"{p // g}/{q // g}", FRAC_UNK_IDX))
            elif el is not None:
                tid = self._token_to_id.get(el, UNK_IDX)
                out.append(tid if self.is_element_token(tid) else UNK_IDX)
            else:
                val = int(integer)
                out.append(self._int_offset + val - 1 if 1 <= val <= MAX_INTEGER else UNK_IDX)
        if add_bos_eos:
            out = [BOS_IDX] + out + [EOS_IDX]
        if pad:
            if len(out) < self.max_len:
                out = out + [PAD_IDX] * (self.max_len - len(out))
            elif len(out) > self.max_len:
                out = out[:self.max_len - 1] + [EOS_IDX]
        return out

    def decode(self, token_ids: Iterable[int], strip_special: bool = True) -> str:
        parts: List[str] = []
        n = len(self._surface)
        for tid in token_ids:
            tid = int(tid)
            if strip_special and tid in (PAD_IDX, BOS_IDX, EOS_IDX):
                if tid == EOS_IDX:
                    break
                continue
            if 0 <= tid < n:
                parts.append(self._surface[tid] if tid > EOS_IDX else self._names[tid])
            else:
                parts.append("?")
        return "".join(parts)

    def decode_batch(self, tokens) -> List[str]:
        """[B, L] tensor (any device) -> formulas, stopping at each row's first EOS."""
        rows = tokens.detach().to("cpu").tolist()
        return [self.decode(r) for r in rows]

    def __repr__(self) -> str:
        iso = f", n_isotopes={self.n_isotope_tokens}" if self._isotope_list else ""
        return (f"FractionAwareTokenizer(vocab_size={self.vocab_size}, n_fractions={self.n_fraction_tokens}{iso}, "
                f"max_len={self.max_len})")
