"""Rollout reward: oracle vs the reference's own outputs (CPU), CUDA kernel vs both (GPU).

Goldens: tests/golden/reward.pt, produced by the reference's compute_reward_gpu_native
(tests/golden/make_golden_reward.py).  Tolerance 1e-3 absolute on rewards in [-100, 100]: integer / boolean work is
exact, the float parts differ by summation order and pow (measured: <= 8e-6 for the oracle)."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import reward as OR
from superconductor_vae_b200 import _lib, reward as R, synthetic as Sy

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reward.pt")
TOL = 1e-3


def _load():
    return torch.load(GOLDEN, weights_only=False)


def _oracle_cfg(kw):
    if kw is None:
        return None
    kw = dict(kw)
    return OR.RewardConfig(v14=kw.pop("cls") == "v14", **kw)


def _shim_cfg(kw):
    if kw is None:
        return None
    kw = dict(kw)
    return (R.GPURewardConfigV14 if kw.pop("cls") == "v14" else R.GPURewardConfig)(**kw)


def test_oracle_matches_reference_goldens():
    g = _load()
    fv = g["fraction_values"].numpy()
    start = g["meta"]["fraction_token_start"]
    assert len(g["cases"]) >= 11
    for name, c in g["cases"].items():
        sem = c["semantic"]
        r = OR.compute_reward(c["sampled"].long().numpy(), c["target"].long().numpy(), c["mask"].numpy(),
                              _oracle_cfg(c["config"]), 2, sem, start if sem else 0, fv if sem else None)
        ref = c["rewards"].numpy()
        assert np.abs(r - ref).max() <= 1e-4, (name, float(np.abs(r - ref).max()))
        # the fixtures reach the branches they were built for
        assert (ref == 100.0).sum() >= 16 and np.unique(ref).size >= 40, name


def test_config_mirrors_reference_fields_and_abi_struct():
    v14 = {f.name: f.default for f in dataclasses.fields(R.GPURewardConfigV14)}
    orc = {f.name: f.default for f in dataclasses.fields(OR.RewardConfig)}
    orc.pop("v14")
    assert v14 == orc                                   # same names / defaults as the oracle's restatement of :42-131
    abi = [n for n, _ in _lib.RewardConfig._fields_]
    assert set(abi) == set(v14) | {"v14"}
    packed = R._pack(R.GPURewardConfigV14(sharpness=2.5, reward_phase=2, use_phased_curriculum=True))
    assert packed.v14 == 1 and abs(packed.sharpness - 2.5) < 1e-7 and packed.reward_phase == 2 and packed.use_phased_curriculum == 1
    base = R._pack(R.GPURewardConfig(near_exact_2=30.0))
    assert base.v14 == 0 and base.near_exact_2 == 30.0 and base.v14_fraction_start == 143
    assert isinstance(R.get_default_gpu_reward_config(), R.GPURewardConfig)
    assert R.get_v14_gpu_reward_config(max_reward=80.0).max_reward == 80.0


def test_reward_has_no_cpu_path():
    s = torch.zeros((2, 4), dtype=torch.long)
    with pytest.raises(_lib.EngineError):
        R.compute_reward_gpu_native(s, s, torch.ones_like(s, dtype=torch.bool))


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_kernel_matches_reference_goldens():
    g = _load()
    dev = "cuda:0"
    fv = g["fraction_values"].to(dev)
    start = g["meta"]["fraction_token_start"]
    for name, c in g["cases"].items():
        sem = c["semantic"]
        r = R.compute_reward_gpu_native(c["sampled"].long().to(dev), c["target"].long().to(dev), c["mask"].to(dev),
                                        config=_shim_cfg(c["config"]), pad_idx=0, end_idx=2, use_semantic_fractions=sem,
                                        fraction_token_start=start if sem else 0, fraction_values=fv if sem else None)
        d = (r.cpu() - c["rewards"]).abs().max().item()
        assert d <= TOL, (name, d)


@pytest.mark.gpu
@pytest.mark.parametrize("semantic", [True, False])
def test_kernel_matches_oracle_long_rows(semantic):
    """63 positions (two 32-position chunks: the parenthesis depth carries across the chunk boundary), 1024 rows,
    float mask (what sample_for_reinforce returns), plus rows without any valid position."""
    dev = "cuda:0"
    B, L, V = 1024, 63, 4752
    s, t, m = Sy.make_reward_rows(B, L, V, 99, old_vocab=not semantic)
    m[5] = False
    m[6, 40:] = False
    fv = Sy.make_fraction_values(V, 143, 7)
    for cfg_o, cfg_s in ((OR.RewardConfig(v14=True), R.GPURewardConfigV14()),
                         (OR.RewardConfig(v14=True, use_phased_curriculum=True, reward_phase=2),
                          R.GPURewardConfigV14(use_phased_curriculum=True, reward_phase=2)),
                         (OR.RewardConfig(), R.GPURewardConfig())):
        ref = OR.compute_reward(s.numpy(), t.numpy(), m.numpy(), cfg_o, 2, semantic, 143 if semantic else 0,
                                fv.numpy() if semantic else None)
        r = R.compute_reward_gpu_native(s.to(dev), t.to(dev), m.float().to(dev), config=cfg_s, use_semantic_fractions=semantic,
                                        fraction_token_start=143 if semantic else 0, fraction_values=fv.to(dev) if semantic else None)
        d = np.abs(r.cpu().numpy() - ref).max()
        assert d <= TOL, (type(cfg_s).__name__, semantic, float(d))


@pytest.mark.gpu
def test_reward_of_engine_rollout_rlo_shape():
    """Config 3 shape: 2048 x 4 sampled rows against repeated targets, as compute_rloo_autoregressive calls it."""
    dev = "cuda:0"
    B, k, L, V = 2048, 4, 63, 4752
    s, t, m = Sy.make_reward_rows(B, L, V, 5)
    sampled = torch.cat([s, t, s.roll(1, 0), t], 0).to(dev)            # sample-major [k * B, L]
    targets = t.repeat(k, 1).to(dev)
    mask = m.repeat(k, 1).to(dev)
    fv = Sy.make_fraction_values(V, 143, 7).to(dev)
    r = R.compute_reward_gpu_native(sampled, targets, mask, config=R.GPURewardConfigV14(), use_semantic_fractions=True,
                                    fraction_token_start=143, fraction_values=fv)
    assert r.shape == (k * B,) and torch.isfinite(r).all()
    r = r.view(k, B)
    full = m.all(dim=1).to(dev)                     # rows scored over every position: sample == target -> exact match
    assert torch.all(r[1][full] == 100.0) and torch.all(r[3][full] == 100.0)
    ref = OR.compute_reward(sampled[:64].cpu().numpy(), targets[:64].cpu().numpy(), mask[:64].cpu().numpy(),
                            OR.RewardConfig(v14=True), 2, True, 143, fv.cpu().numpy())
    assert np.abs(r[0, :64].cpu().numpy() - ref).max() <= TOL


@pytest.mark.gpu
def test_rollout_with_rewards_is_the_reference_block_in_one_call():
    """decoder.rollout_with_rewards = the rollout block of the reference's RL loss (scripts/train_v12_clean.py:2677-2766):
    k samples per latent (sample-major), pad / truncate to the targets' length, task reward + constraint rewards.  Checked
    against the same steps done one by one with the reference's way of expanding the inputs (repeat k times), and the
    rewards of 64 rows against the oracle."""
    import superconductor_vae_b200 as S
    from superconductor_vae_b200 import constraints as K
    dev = "cuda:0"
    B, k, V = 160, 3, 4752
    dec = S.EnhancedTransformerDecoder.from_state_dict(Sy.make_decoder_state_dict(Sy.C512, 0), nhead=8, device=dev)
    z = Sy.make_latents(B, 2048, 11).to(dev)
    st, hp = Sy.make_conditioning(B, 13, 11)
    st, hp = st.to(dev), {n: v.to(dev) for n, v in hp.items()}
    _, targets, _ = Sy.make_reward_rows(B, 40, V, 5)                   # targets shorter than max_len: the rollout is truncated
    fv = Sy.make_fraction_values(V, 143, 7).to(dev)
    kw = dict(stoich_pred=st, heads_pred=hp, temperature=1.2, max_len=64, stop_boost=10.0, _seed=3)
    ccfg = K.ConstraintRewardConfig()
    tok, lp, ent, mask, rew = dec.rollout_with_rewards(z, targets, n_samples=k, reward_config=R.GPURewardConfigV14(),
                                                       constraint_config=ccfg, use_semantic_fractions=True,
                                                       fraction_token_start=143, fraction_values=fv, **kw)
    assert tok.shape == (k * B, 40) and lp.shape == tok.shape and ent.shape == tok.shape and mask.shape == tok.shape
    rep = lambda t: t.repeat(k, *([1] * (t.dim() - 1)))
    t0, lp0, en0, mk0 = dec.sample_for_reinforce(rep(z), stoich_pred=rep(st), heads_pred={n: rep(v) for n, v in hp.items()},
                                                 temperature=1.2, max_len=64, stop_boost=10.0, _seed=3)
    L0 = t0.shape[1]
    if L0 < 40:
        pad = (0, 40 - L0)
        t0, lp0 = torch.nn.functional.pad(t0, pad, value=0), torch.nn.functional.pad(lp0, pad)
        en0, mk0 = torch.nn.functional.pad(en0, pad), torch.nn.functional.pad(mk0, pad)
    t0, lp0, en0, mk0 = t0[:, :40], lp0[:, :40], en0[:, :40], mk0[:, :40]
    assert torch.equal(tok, t0) and torch.equal(lp, lp0) and torch.equal(ent, en0) and torch.equal(mask, mk0)
    tg = targets.repeat(k, 1).to(dev)
    r0 = R.compute_reward_gpu_native(t0, tg, mk0.bool(), config=R.GPURewardConfigV14(), use_semantic_fractions=True,
                                     fraction_token_start=143, fraction_values=fv)
    r0 = r0 + K.compute_constraint_rewards(t0, mk0, config=ccfg)
    assert torch.equal(rew, r0)
    ref = OR.compute_reward(t0[:64].cpu().numpy(), tg[:64].cpu().numpy(), mk0[:64].bool().cpu().numpy(), OR.RewardConfig(v14=True), 2,
                            True, 143, fv.cpu().numpy())
    task = R.compute_reward_gpu_native(t0[:64], tg[:64], mk0[:64].bool(), config=R.GPURewardConfigV14(), use_semantic_fractions=True,
                                       fraction_token_start=143, fraction_values=fv)
    assert np.abs(task.cpu().numpy() - ref).max() <= TOL


@pytest.mark.gpu
def test_kernel_rows_longer_than_three_chunks():
    """70 rows of 100 positions (four 32-position chunks, ragged last one), V14 continuous and tiered digit-level paths."""
    dev = "cuda:0"
    B, L, V = 70, 100, 4752
    fv = Sy.make_fraction_values(V, 143, 7)
    for semantic, cfg_o, cfg_s in ((True, OR.RewardConfig(v14=True), R.GPURewardConfigV14()), (False, OR.RewardConfig(), R.GPURewardConfig())):
        s, t, m = Sy.make_reward_rows(B, L, V, 123, old_vocab=not semantic)
        ref = OR.compute_reward(s.numpy(), t.numpy(), m.numpy(), cfg_o, 2, semantic, 143 if semantic else 0, fv.numpy() if semantic else None)
        r = R.compute_reward_gpu_native(s.to(dev), t.to(dev), m.to(dev), config=cfg_s, use_semantic_fractions=semantic,
                                        fraction_token_start=143 if semantic else 0, fraction_values=fv.to(dev) if semantic else None)
        assert np.abs(r.cpu().numpy() - ref).max() <= TOL, semantic
