"""Golden vectors of the chemistry-constraint rewards: runs the REFERENCE
`compute_constraint_rewards` (and each rule function on its own) from /root/reference/src on seeded formula-like rows
in both vocabulary layouts.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_constraints.py

Output: tests/golden/constraints.pt (committed; torch version and seeds recorded inside).
"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

import torch

with contextlib.redirect_stdout(io.StringIO()):
    from superconductor.losses import constraint_rewards as CR

from superconductor_vae_b200.synthetic import (make_constraint_rows, make_family_probs,      # noqa: E402
                                               make_constraint_fraction_values)

SEED, B, L = 777, 400, 28


def main():
    out = {"meta": {"torch": torch.__version__, "seed": SEED, "B": B, "L": L}, "cases": {}}
    fv = make_constraint_fraction_values()
    fam = make_family_probs(B, SEED + 1)
    for name, semantic in (("v13", True), ("v12", False)):
        tokens, mask = make_constraint_rows(B, L, SEED + (0 if semantic else 5), semantic)
        CR.set_vocab_config(CR.make_v13_vocab_config(143, fv) if semantic else CR.VocabConfig())
        cfg, fcfg = CR.ConstraintRewardConfig(), CR.FamilyConstraintConfig()
        case = {"semantic": semantic, "tokens": tokens.to(torch.int16), "mask": mask,
                "total": CR.compute_constraint_rewards(tokens, mask, cfg, fam, fcfg),
                "total_no_family": CR.compute_constraint_rewards(tokens, mask, cfg),
                "a1": CR.compute_duplicate_element_penalty(tokens, mask, cfg.a1_penalty),
                "a2": CR.compute_gcd_canonicality_penalty(tokens, mask, cfg.a2_penalty_per_violation),
                "a4": CR.compute_stoich_normalization_penalty(tokens, mask, cfg.a4_penalty),
                "a7": CR.compute_impossible_element_penalty(tokens, mask, cfg.a7_penalty),
                "family": CR.compute_family_constraint_rewards(tokens, mask, fam, fcfg)}
        custom = CR.ConstraintRewardConfig(a1_penalty=-7.5, a2_enabled=False, a4_penalty=-3.25, a7_penalty=-11.0)
        fcustom = CR.FamilyConstraintConfig(confidence_threshold=0.5, b1_penalty=-1.5, b5_penalty=-2.5, b7_penalty=-4.5, b8_penalty=-8.0)
        case["total_custom"] = CR.compute_constraint_rewards(tokens, mask, custom, fam, fcustom)
        out["cases"][name] = case
        print(name, {k: int((v != 0).sum()) for k, v in case.items() if isinstance(v, torch.Tensor) and v.dtype == torch.float32})
    CR.set_vocab_config(CR.VocabConfig())
    out["fraction_values"], out["family_probs"] = fv, fam
    torch.save(out, os.path.join(HERE, "constraints.pt"))


if __name__ == "__main__":
    main()
