"""Chemistry-constraint rewards: oracle vs the reference's own outputs (CPU), CUDA kernel vs both (GPU).

Goldens: tests/golden/constraints.pt, produced by the reference's compute_constraint_rewards and its rule functions
(tests/golden/make_golden_constraints.py).  The work is integer / comparison logic on doubles: results must be
bit-identical."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import constraints as OC
from superconductor_vae_b200 import _lib, constraints as K, synthetic as Sy

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "constraints.pt")
OFF = dict(a1_enabled=False, a2_enabled=False, a4_enabled=False, a7_enabled=False)
CUSTOM_A = dict(a1_penalty=-7.5, a2_enabled=False, a4_penalty=-3.25, a7_penalty=-11.0)
CUSTOM_F = dict(confidence_threshold=0.5, b1_penalty=-1.5, b5_penalty=-2.5, b7_penalty=-4.5, b8_penalty=-8.0)


def _load():
    return torch.load(GOLDEN, weights_only=False)


def test_oracle_matches_reference_goldens():
    g = _load()
    fv, fam = g["fraction_values"].numpy(), g["family_probs"].numpy()
    for name, c in g["cases"].items():
        v = OC.Vocab.v13(143, fv) if c["semantic"] else OC.Vocab()
        tok, m = c["tokens"].long().numpy(), c["mask"].numpy()
        run = lambda rules, f=None: OC.compute_constraint_rewards(tok, m, rules, v, f)
        got = {"total": run(OC.Rules(), fam), "total_no_family": run(OC.Rules()),
               "a1": run(OC.Rules(**{**OFF, "a1_enabled": True})), "a2": run(OC.Rules(**{**OFF, "a2_enabled": True})),
               "a4": run(OC.Rules(**{**OFF, "a4_enabled": True})), "a7": run(OC.Rules(**{**OFF, "a7_enabled": True})),
               "family": run(OC.Rules(**OFF), fam), "total_custom": run(OC.Rules(**CUSTOM_A, **CUSTOM_F), fam)}
        for k, r in got.items():
            assert np.array_equal(r, c[k].numpy()), (name, k)
        # the fixtures reach the rules
        assert all(int((c[k] != 0).sum()) >= 10 for k in ("a1", "a4", "a7", "family")), name
    assert int((g["cases"]["v12"]["a2"] != 0).sum()) >= 10 and int((g["cases"]["v13"]["a2"] != 0).sum()) == 0


def test_config_mirrors_reference_fields():
    assert [f.name for f in dataclasses.fields(K.VocabConfig)] == [f.name for f in dataclasses.fields(OC.Vocab)]
    v13 = K.make_v13_vocab_config(143)
    o13 = OC.Vocab.v13(143)
    assert all(getattr(v13, f.name) == getattr(o13, f.name) for f in dataclasses.fields(OC.Vocab) if f.name != "fraction_values")
    assert K.VocabConfig().elem_idx(29) == 48 and v13.elem_idx(29) == 33
    a, f, r = K.ConstraintRewardConfig(), K.FamilyConstraintConfig(), OC.Rules()
    assert all(getattr(a, x.name) == getattr(r, x.name) for x in dataclasses.fields(K.ConstraintRewardConfig))
    assert f.enabled == r.family_enabled and f.confidence_threshold == r.confidence_threshold
    assert all(getattr(f, f"b{i}_penalty") == getattr(r, f"b{i}_penalty") for i in range(1, 9))


def test_constraints_have_no_cpu_path():
    s = torch.zeros((2, 4), dtype=torch.long)
    with pytest.raises(_lib.EngineError):
        K.compute_constraint_rewards(s, torch.ones_like(s, dtype=torch.bool), K.ConstraintRewardConfig())


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_kernel_matches_reference_goldens():
    g = _load()
    dev = "cuda:0"
    fv, fam = g["fraction_values"].to(dev), g["family_probs"].to(dev)
    try:
        for name, c in g["cases"].items():
            K.set_vocab_config(K.make_v13_vocab_config(143, fv) if c["semantic"] else K.VocabConfig())
            tok, m = c["tokens"].long().to(dev), c["mask"].to(dev)
            a, f = K.ConstraintRewardConfig(), K.FamilyConstraintConfig()
            got = {"total": K.compute_constraint_rewards(tok, m, a, fam, f),
                   "total_no_family": K.compute_constraint_rewards(tok, m, a),
                   "a1": K.compute_duplicate_element_penalty(tok, m, a.a1_penalty),
                   "a2": K.compute_gcd_canonicality_penalty(tok, m, a.a2_penalty_per_violation),
                   "a4": K.compute_stoich_normalization_penalty(tok, m, a.a4_penalty),
                   "a7": K.compute_impossible_element_penalty(tok, m, a.a7_penalty),
                   "family": K.compute_family_constraint_rewards(tok, m, fam, f),
                   "total_custom": K.compute_constraint_rewards(tok, m, K.ConstraintRewardConfig(**CUSTOM_A), fam,
                                                                K.FamilyConstraintConfig(**CUSTOM_F))}
            for k, r in got.items():
                assert torch.equal(r.cpu(), c[k]), (name, k, int((r.cpu() != c[k]).sum()))
    finally:
        K.set_vocab_config(K.VocabConfig())


@pytest.mark.gpu
@pytest.mark.parametrize("semantic", [True, False])
def test_kernel_matches_oracle_rollout_shape(semantic):
    """8192 x 63 rows (config-3 rollout shape, float mask as sample_for_reinforce returns it): bit-identical to the oracle."""
    dev = "cuda:0"
    B, L = 8192, 63
    tok, m = Sy.make_constraint_rows(B, L, 31, semantic)
    fv = Sy.make_constraint_fraction_values()
    fam = Sy.make_family_probs(B, 32)
    v = OC.Vocab.v13(143, fv.numpy()) if semantic else OC.Vocab()
    ref = OC.compute_constraint_rewards(tok[:1024].numpy(), m[:1024].numpy(), OC.Rules(), v, fam[:1024].numpy())
    try:
        K.set_vocab_config(K.make_v13_vocab_config(143, fv.to(dev)) if semantic else K.VocabConfig())
        r = K.compute_constraint_rewards(tok.to(dev), m.float().to(dev), K.ConstraintRewardConfig(), fam.to(dev),
                                         K.FamilyConstraintConfig())
    finally:
        K.set_vocab_config(K.VocabConfig())
    assert r.shape == (B,) and np.array_equal(r[:1024].cpu().numpy(), ref)
    assert int((r != 0).sum()) > B // 10


@pytest.mark.gpu
@pytest.mark.parametrize("semantic", [True, False])
def test_kernel_long_rows_partial_cta(semantic):
    """130 rows of 131 positions: a partial second CTA and rows wider than the default 48 KB of shared memory."""
    dev = "cuda:0"
    B, L = 130, 131
    tok, m = Sy.make_constraint_rows(B, L, 41, semantic)
    fv = Sy.make_constraint_fraction_values()
    fam = Sy.make_family_probs(B, 42)
    v = OC.Vocab.v13(143, fv.numpy()) if semantic else OC.Vocab()
    ref = OC.compute_constraint_rewards(tok.numpy(), m.numpy(), OC.Rules(), v, fam.numpy())
    try:
        K.set_vocab_config(K.make_v13_vocab_config(143, fv.to(dev)) if semantic else K.VocabConfig())
        r = K.compute_constraint_rewards(tok.to(dev), m.to(dev), K.ConstraintRewardConfig(), fam.to(dev), K.FamilyConstraintConfig())
    finally:
        K.set_vocab_config(K.VocabConfig())
    assert np.array_equal(r.cpu().numpy(), ref)
