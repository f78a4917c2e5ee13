"""Golden vectors of the rollout reward: runs the REFERENCE `compute_reward_gpu_native`
(imported from /root/reference/src) on seeded random token rows built to reach every branch.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_reward.py

Output: tests/golden/reward.pt (committed; torch version and seed recorded inside).
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
    from superconductor.losses.reward_gpu_native import (GPURewardConfig, GPURewardConfigV14, compute_reward_gpu_native)

from superconductor_vae_b200.synthetic import make_reward_rows, make_fraction_values      # noqa: E402

SEED, B, L, V = 4242, 160, 24, 4752
FRAC_START = 143

CASES = {
    "v14_semantic": (dict(cls="v14"), True),
    "v14_digit": (dict(cls="v14"), False),
    "v14_phase1": (dict(cls="v14", use_phased_curriculum=True, reward_phase=1), True),
    "v14_phase2": (dict(cls="v14", use_phased_curriculum=True, reward_phase=2), True),
    "v14_phase3": (dict(cls="v14", use_phased_curriculum=True, reward_phase=3), True),
    "v14_tiered": (dict(cls="v14", use_continuous_reward=False), True),
    "v14_custom": (dict(cls="v14", sharpness=2.5, max_reward=80.0, element_error_penalty=-4.0, too_short_per_missing=7.0,
                        length_only_floor=20.0, length_mismatch_penalty=-1.5, semantic_digit_scale=3.0), True),
    "tiered_digit": (dict(cls="base"), False),
    "tiered_flat_digit": (dict(cls="base", use_semantic_digit_penalty=False), False),
    "tiered_semantic": (dict(cls="base", near_exact_2=30.0, token_penalty=-0.75), True),
    "default_none": (None, False),
}


def main():
    out = {"meta": {"torch": torch.__version__, "seed": SEED, "B": B, "L": L, "V": V, "fraction_token_start": FRAC_START},
           "cases": {}}
    fv = make_fraction_values(V, FRAC_START, SEED)
    for i, (name, (kw, semantic)) in enumerate(CASES.items()):
        sampled, target, mask = make_reward_rows(B, L, V, SEED + i, old_vocab=not semantic)
        if kw is None:
            cfg = None
        else:
            kw = dict(kw)
            cls = GPURewardConfigV14 if kw.pop("cls") == "v14" else GPURewardConfig
            cfg = cls(**kw)
        r = compute_reward_gpu_native(sampled, target, mask, config=cfg, pad_idx=0, end_idx=2,
                                      use_semantic_fractions=semantic, fraction_token_start=FRAC_START if semantic else 0,
                                      fraction_values=fv if semantic else None)
        out["cases"][name] = {"config": CASES[name][0], "semantic": semantic, "seed": SEED + i,
                              "sampled": sampled.to(torch.int16), "target": target.to(torch.int16), "mask": mask,
                              "rewards": r.float()}
        print(name, "rewards: min %.3f max %.3f, exact %d, distinct %d" % (float(r.min()), float(r.max()),
              int((r == (cfg.exact_match if cfg else 100.0)).sum()), int(r.unique().numel())))
    out["fraction_values"] = fv
    torch.save(out, os.path.join(HERE, "reward.pt"))


if __name__ == "__main__":
    main()
