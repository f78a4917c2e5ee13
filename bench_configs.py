#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (3: RLOO rollouts, 4: 1M SLERP latents, 5: fused encoder).
Not the driver's contract bench (that is bench.py = config 2); prints one JSON line per configuration.

    python bench_configs.py [--configs 3,4,5] [--latents 1000000]
    torchrun --nproc-per-node N bench_configs.py --configs 4        # config 4 sharded, NCCL gather of sequences
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry the JSON lines only: NCCL (and anything else in native code) prints its banner to fd 1, so fd 1 is
# pointed at stderr for the whole run and the JSON goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


import torch
import torch.distributed as dist

import superconductor_vae_b200 as S
from superconductor_vae_b200 import latent, parallel, synthetic as Sy
from superconductor_vae_b200.tokenizer import FractionAwareTokenizer


def timed(fn, warmup=1, iters=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--latents", type=int, default=1_000_000)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dec = S.EnhancedTransformerDecoder.from_state_dict(Sy.make_decoder_state_dict(Sy.C512, 0), nhead=8, device=dev)
    enc = S.FullMaterialsVAE.from_state_dict(Sy.make_encoder_state_dict(Sy.ENC_DEFAULT, 1), device=dev)
    tok = FractionAwareTokenizer(max_len=64, fractions=[f"{i + 1}/100003" for i in range(4317)],
                                 isotopes=[f"{300 + i}Og" for i in range(291)])
    masks = tok.get_type_masks(dev)
    todo = [int(c) if c.isdigit() else c for c in args.configs.split(",")]

    if 3 in todo and rank == 0:
        # RLOO: 2048 latents x k = 4 samples (sample-major repeat), temperature 1.2, log-probs + entropy + mask
        B, k = 2048, 4
        z = Sy.make_latents(B, 2048, 1234).to(dev).repeat(k, 1)
        st, hp = Sy.make_conditioning(B, 13, 1234)
        st, hp = st.to(dev).repeat(k, 1), {n: v.to(dev).repeat(k, *([1] * (v.dim() - 1))) for n, v in hp.items()}
        fn = lambda: dec.sample_for_reinforce(z, stoich_pred=st, temperature=1.2, max_len=64, stop_boost=10.0,
                                              heads_pred=hp, _seed=7)
        ms, (t, lp, en, mk) = timed(fn)
        emit(({"config": 3, "what": "RLOO rollouts: 2048 latents x 4 samples, temperature 1.2, per-token log-probs, "
                          "entropy and mask (stop_boost 10, no hard masks so the reference's H2 fallback cannot trip)",
                          "rows": B * k, "executed_decode_steps": int(t.shape[1]), "ms": ms,
                          "formulas_per_s": B * k / (ms / 1e3), "mean_len": float(mk.sum(1).mean())}))
        # the same rollouts with the k samples of a latent sharing its memory tokens inside the engine (_n_samples):
        # identical outputs (tests/test_gpu_parity.py), the projected memory K / V are built and streamed once per latent
        zb, stb = Sy.make_latents(B, 2048, 1234).to(dev), Sy.make_conditioning(B, 13, 1234)[0].to(dev)
        hpb = {n: v.to(dev) for n, v in Sy.make_conditioning(B, 13, 1234)[1].items()}
        fn2 = lambda: dec.sample_for_reinforce(zb, stoich_pred=stb, temperature=1.2, max_len=64, stop_boost=10.0,
                                               heads_pred=hpb, _seed=7, _n_samples=k)
        ms2, (t2, _, _, _) = timed(fn2)
        emit(({"config": "3-shared", "what": "the same RLOO rollouts through `_n_samples=4`: base batch in, the 4 samples of a "
               "latent share its memory tokens and their projected K / V (scv_generate_args.memory_rows)", "rows": B * k,
               "executed_decode_steps": int(t2.shape[1]), "ms": ms2, "formulas_per_s": B * k / (ms2 / 1e3),
               "same_tokens_as_repeated_inputs": bool(torch.equal(t, t2))}))
        # f1: the token-level reward of those rollouts (compute_reward_gpu_native, V14 continuous reward + semantic
        # fraction values, as scripts/train_v12_clean.py:2745-2752 calls it): one kernel
        from superconductor_vae_b200 import reward as R
        L = int(t.shape[1])
        _, targets, _ = Sy.make_reward_rows(B, L, 4752, 11)
        targets = targets.to(dev).repeat(k, 1)
        fv = Sy.make_fraction_values(4752, 143, 7).to(dev)
        cfg = R.GPURewardConfigV14()
        rfn = lambda: R.compute_reward_gpu_native(t, targets, mk.bool(), config=cfg, use_semantic_fractions=True,
                                                  fraction_token_start=143, fraction_values=fv)
        rms, rew = timed(rfn, warmup=2, iters=20)
        emit(({"config": "f1", "what": "token-level rollout reward (V14 continuous + semantic fraction penalty) of the "
                          "config-3 rollouts: one warp-per-row kernel", "rows": B * k, "seq_len": L, "us": 1e3 * rms,
                          "rows_per_s": B * k / (rms / 1e3), "algorithmic_bytes": B * k * L * 17,
                          "gbs": B * k * L * 17 / (rms * 1e6), "mean_reward": float(rew.mean())}))
        # the chemistry-constraint rewards added to it (compute_constraint_rewards, :2754-2766): the reference walks every
        # row on the host with Python loops; here one kernel.  Formula-like rows of the rollout's shape.
        from superconductor_vae_b200 import constraints as K
        ctok, cmask = Sy.make_constraint_rows(B * k, L, 31, True)
        ctok, cmask = ctok.to(dev), cmask.to(dev)
        fam = Sy.make_family_probs(B * k, 32).to(dev)
        K.set_vocab_config(K.make_v13_vocab_config(143, Sy.make_constraint_fraction_values().to(dev)))
        kfn = lambda: K.compute_constraint_rewards(ctok, cmask, K.ConstraintRewardConfig(), fam, K.FamilyConstraintConfig())
        kms, kr = timed(kfn, warmup=2, iters=20)
        K.set_vocab_config(K.VocabConfig())
        emit(({"config": "f1-constraints", "what": "chemistry-constraint rewards (A1, A2, A4, A7, B1-B8) of 8192 formula-like "
                          "rows: one thread-per-row kernel over rows staged in shared memory", "rows": B * k, "seq_len": L,
                          "us": 1e3 * kms, "rows_per_s": B * k / (kms / 1e3), "rows_penalised": int((kr != 0).sum())}))

    if 5 in todo and rank == 0:
        n = 52800
        idx, frac, mask, magpie, tc = (t.to(dev) for t in Sy.make_compositions(n, 7))

        def fn():
            z = enc.encode(idx, frac, mask, magpie, tc)["z"]
            st, hp = enc.conditioning(z)
            return dec.precompute_memory(z, None, st, hp)
        ms, mem = timed(fn)
        emit(({"config": 5, "what": "three-branch encoder + all heads + memory tokens on 52,800 synthetic compositions",
                          "rows": n, "ms": ms, "rows_per_s": n / (ms / 1e3), "memory_shape": list(mem.shape),
                          "algorithmic_tflops": n * 103.4e6 / (ms * 1e9)}))

    if "f3" in todo and rank == 0:
        # teacher-forced forward: 256 sequences x 63 positions in one pass
        Bf, Lf = 256, 64
        z = Sy.make_latents(Bf, 2048, 1234).to(dev)
        st, hp = Sy.make_conditioning(Bf, 13, 1234)
        st, hp = st.to(dev), {k: v.to(dev) for k, v in hp.items()}
        g = torch.Generator().manual_seed(3)
        tgt = torch.randint(3, dec.vocab_size, (Bf, Lf), generator=g)
        tgt[:, 0] = 1
        tgt[::2, 40:] = 0
        tgt = tgt.to(dev)
        mem = dec.precompute_memory(z, None, st, hp)
        ms, out = timed(lambda: dec(z, tgt, cached_memory=mem), warmup=2, iters=5)
        emit(({"config": "f3", "what": "teacher-forced forward, 256 sequences x 63 positions (16,128 rows per projection), "
                                       "logits + stop / type / site-dup heads at every position", "ms": ms,
               "positions_per_s": Bf * (Lf - 1) / (ms / 1e3), "logits_shape": list(out[0].shape)}))
    if 4 in todo:
        # 1M SLERP-interpolated latents between 1024 anchors, conditioning from z alone (notebook pipeline),
        # greedy decode with masks + stop head, rows sharded over the ranks, NCCL gather of token ids only
        N = args.latents
        lo, hi = parallel.shard_bounds(N, world, rank)
        anchors = Sy.make_latents(1024, 2048, 1234).to(dev)
        g = torch.Generator().manual_seed(99)
        i1 = torch.randint(0, 1024, (N,), generator=g)
        i2 = (i1 + torch.randint(1, 1024, (N,), generator=g)) % 1024
        tt = torch.rand((N,), generator=g) * 0.9 + 0.05
        i1, i2, tt = i1[lo:hi].to(dev), i2[lo:hi].to(dev), tt[lo:hi].to(dev)
        chunk = 65536
        bounds = [parallel.shard_bounds(N, world, r) for r in range(world)]
        n_chunks = max((h - l + chunk - 1) // chunk for l, h in bounds)
        out = torch.zeros((hi - lo, 63), dtype=torch.int16, device=dev)
        # the product's multi-GPU plumbing: every chunk's int16 token ids are all-gathered asynchronously (NCCL runs on its
        # own stream), so the gather of chunk k overlaps the decode of chunk k + 1 (parallel.ChunkedGather)
        # untimed warm-up, as for every other line of this file: one decode per distinct chunk size of this rank (workspace
        # allocation, step graphs) and one gather (NCCL communicator), so the line is the steady state of a long search
        sizes = sorted({min(hi - lo, (k + 1) * chunk) - min(hi - lo, k * chunk) for k in range(n_chunks)} - {0})
        for n_w in sizes:
            z = latent.slerp_rows(anchors, i1[:n_w], i2[:n_w], tt[:n_w])
            latent.decode_z_batch(enc, dec, z, temperature=0.001, type_masks=masks, max_len=64)
        if world > 1:
            wg = parallel.ChunkedGather(pad_to=63, dtype=torch.int16)
            wg.push(out[:8], [8] * world)
            wg.finish()
        cg = parallel.ChunkedGather(pad_to=63, dtype=torch.int16) if world > 1 else None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        L = 0
        for k in range(n_chunks):
            c0, c1 = min(hi - lo, k * chunk), min(hi - lo, (k + 1) * chunk)
            if c1 > c0:
                z = latent.slerp_rows(anchors, i1[c0:c1], i2[c0:c1], tt[c0:c1])
                toks = latent.decode_z_batch(enc, dec, z, temperature=0.001, type_masks=masks, max_len=64)
                out[c0:c1, :toks.shape[1]] = toks.to(torch.int16)
                L = max(L, toks.shape[1])
            if cg is not None:
                cg.push(out[c0:c1], [max(0, min(h - l, (k + 1) * chunk) - min(h - l, k * chunk)) for l, h in bounds])
        if world > 1:
            gathered = cg.finish()
            lmax = torch.tensor([L], device=dev)
            dist.all_reduce(lmax, op=dist.ReduceOp.MAX)
            gathered = gathered[:, :int(lmax)]
        else:
            gathered = out[:, :L]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tmax = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax)
        if rank == 0:
            lens = (gathered == 2).int().argmax(dim=1) + 1
            # post-processing: group identical candidates on the device, build strings once per group;
            # the per-row Python loop of the reference is timed on a 20,000-row sample
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            formulas, inverse, counts = latent.decode_unique(tok, gathered.to(torch.int64))
            torch.cuda.synchronize()
            dt_unique = time.perf_counter() - t1
            t1 = time.perf_counter()
            tok.decode_batch(gathered[:20000])
            dt_rows = (time.perf_counter() - t1) * (N / 20000.0)
            emit(({"config": 4, "what": "SLERP latents -> heads_from_latent -> greedy decode (masks + stop head), "
                              "sharded over ranks, chunked asynchronous all-gather of int16 token ids (parallel.ChunkedGather)", "latents": N, "n_gpus": world,
                              "warmup": "one untimed decode per distinct chunk size + one gather", "seconds": dt, "formulas_per_s": N / dt, "max_len_gathered": int(gathered.shape[1]),
                              "mean_formula_len": float(lens.float().mean()), "distinct_candidates": len(formulas),
                              "decode_unique_seconds": dt_unique, "decode_every_row_seconds_extrapolated": dt_rows,
                              "sample": tok.decode_batch(gathered[:2])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
