#!/usr/bin/env python
"""Benchmark of the KV-cache decode hot path (BASELINE.json metric: formulas/sec, KV-cache decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]

One "step" = one full greedy decode of a batch of synthetic latents (BASELINE config 2: 4096 latents per GPU,
max_len 64, type masks + stop head + hard stop, temperature 0.001, 24 memory tokens).  Weak scaling: every
rank decodes its own 4096 latents with replicated weights; the only collective is the gather of token ids.

`value`   : formulas/s with the conditioning inputs already resident in HBM.
`e2e`     : same call through the public Python API with HOST (pinned) inputs: H2D of z / stoich / heads and
            D2H of the token ids inside the timed region.
`roofline`: the kernel category with the largest share of the step, timed live with CUDA events on the launch
            stream in a separate (untimed) profiling pass; peaks from MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference's algorithm (the reference is pure
            Python/PyTorch and cannot travel to the GPU box) on a bounded sample, all host threads.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry the JSON lines only: NCCL (and anything else in native code) prints its banner to fd 1, so fd 1 is
# pointed at stderr for the whole run and the JSON goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


METRIC = "formulas/sec (KV-cache decode, bf16)"
UNIT = "formulas/s"
WORKLOAD = ("BASELINE config 2: greedy batched generation of 4096 synthetic 2048-d latents per GPU, C512 decoder "
            "(d_model 512, 8 heads, 12 layers, ff 2048, vocab 4752), 24 memory tokens, max_len 64, "
            "type-mask + stop head + hard stop, temperature 0.001")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; started before the warm-up (nvidia-smi needs a few
    hundred ms to come up), reported over the timed region only (samples are time-stamped by the reader thread)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,clocks.mem")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 is None or (t0 <= ts <= t1 + 0.05)]
        in_window = len(rows)
        if not rows:                                   # region shorter than a sampling period: everything under load
            rows = [r for _, r in self.rows]
        sm, mx, mem, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            try:
                mem.append(float(r[8]))
            except Exception:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        hot = [v for v in sm if v >= 0.5 * (max(sm) if sm else 0)]
        return {"sm_mhz": hot[len(hot) // 2] if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "mem_mhz": sorted(mem)[len(mem) // 2] if mem else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": in_window}


def algorithmic_bytes(B, executed_steps, kv_bytes_per_pos=49152, weight_bytes=107_091_244):
    """SURVEY 8d cost model with the exactness configuration's fp32 self-attention KV cache:
    bytes_step(B, t) = weights + B * kv_bytes_per_pos * (t + 1) (read t cached + the new position, write one)."""
    return sum(weight_bytes + B * kv_bytes_per_pos * (t + 1) for t in range(executed_steps))


def algorithmic_flops(B, executed_steps):
    return sum(B * (2 * 53_420_800 + 4 * 512 * 12 * (t + 1) + 4 * 24 * 512 * 12) for t in range(executed_steps))


# Steps the full 4096-latent batch executes before every row has emitted END (deterministic for the seeded
# synthetic weights / inputs; measured on the B200 by the engine arm and asserted there).  The CPU sample is run
# for the same number of steps so that its cost per formula is that of the full batch, not of a small batch
# that happens to finish early.
FULL_BATCH_STEPS = 22      # measured on B200 (profiles/bench_r01*.json: "executed_decode_steps": 22)


def cpu_oracle_run(rows, threads, max_len=64, repeats=1, steps=FULL_BATCH_STEPS):
    """Time the CPU oracle port of the reference path (config 2 sample) on `rows` latents for `steps` steps."""
    import torch
    from oracle import decoder_oracle as DO, vocab as OV
    from superconductor_vae_b200 import synthetic as Sy
    torch.set_num_threads(threads)
    sd = Sy.make_decoder_state_dict(Sy.C512, 0)
    z = Sy.make_latents(rows, 2048, 1234)
    stoich, heads = Sy.make_conditioning(rows, 13, 1234)
    masks = OV.type_masks()
    best, L = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        tok, _, _ = DO.generate_with_kv_cache(sd, 8, z, stoich_pred=stoich, temperature=0.001, max_len=max_len,
                                              heads_pred=heads, type_masks=masks, stop_boost=10.0,
                                              hard_stop_threshold=0.8, stop_when_all_finished=False,
                                              max_steps=steps)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        L = tok.shape[1]
    return rows / best, best, L


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = args.cpu_rows
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, L = cpu_oracle_run(rows, threads, args.max_len)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    sample = (f"{rows} latents of the config-2 workload per step (the 4096-latent batch is bounded to {rows} rows "
              f"for the CPU), {L} executed steps, torch {threads} threads, fp32")
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_engine(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import superconductor_vae_b200 as S
    from superconductor_vae_b200 import _lib, synthetic as Sy
    from superconductor_vae_b200.tokenizer import FractionAwareTokenizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, max_len = args.batch, args.max_len
    sd = Sy.make_decoder_state_dict(Sy.C512, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=dev)
    dec.max_rows_per_call = max(dec.max_rows_per_call, B)
    tok = FractionAwareTokenizer(max_len=max_len, fractions=[f"{i + 1}/100003" for i in range(4317)],
                                 isotopes=[f"{300 + i}Og" for i in range(291)])
    masks = tok.get_type_masks(dev)
    # host-side (pinned) inputs: different latents per rank (weak scaling)
    z_h = Sy.make_latents(B, 2048, 1234 + rank).pin_memory()
    st_h, hp_h = Sy.make_conditioning(B, 13, 1234 + rank)
    st_h = st_h.pin_memory()
    hp_h = {k: v.pin_memory() for k, v in hp_h.items()}
    z, st, hp = z_h.to(dev), st_h.to(dev), {k: v.to(dev) for k, v in hp_h.items()}
    kw = dict(temperature=0.001, max_len=max_len, type_masks=masks, stop_boost=10.0, hard_stop_threshold=0.8)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    gather_buf = [torch.empty((B, max_len - 1), dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None
    tok_host = torch.empty((B, max_len - 1), dtype=torch.int32).pin_memory()
    h2d_bytes = z_h.numel() * 4 + st_h.numel() * 4 + sum(v.numel() * 4 for v in hp_h.values())

    def step_resident():
        t, _, _ = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
        if world > 1:      # NCCL gather of sequences only
            pad = torch.zeros((B, max_len - 1), dtype=torch.int32, device=dev)
            pad[:, :t.shape[1]] = t
            dist.all_gather(gather_buf, pad)
        return t

    def step_e2e():
        zd = z_h.to(dev, non_blocking=True)
        sd_ = st_h.to(dev, non_blocking=True)
        hd = {k: v.to(dev, non_blocking=True) for k, v in hp_h.items()}
        t, _, _ = dec.generate_with_kv_cache(zd, stoich_pred=sd_, heads_pred=hd, **kw)
        t32 = t.to(torch.int32)
        if world > 1:
            pad = torch.zeros((B, max_len - 1), dtype=torch.int32, device=dev)
            pad[:, :t.shape[1]] = t32
            dist.all_gather(gather_buf, pad)
        tok_host[:, :t.shape[1]].copy_(t32, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return t

    def timed(fn, n):
        total_ms, last = 0.0, None
        for _ in range(n):
            flush.fill_(1)                                   # evict L2 between timed iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            last = fn()
            b.record()
            b.synchronize()
            total_ms += a.elapsed_time(b)
        return total_ms, last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.ncu:          # profiler-friendly run: `--warmup` decodes then ONE decode, no timing, no JSON contract
        for _ in range(args.warmup):
            step_resident()
        t = step_resident()
        torch.cuda.synchronize()
        emit(({"ncu_run": True, "executed_decode_steps": int(t.shape[1]),
                          "launches_per_decode": _lib.launch_count() // (args.warmup + 1)}))
        return

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        step_resident()
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t_region0 = time.time()
    n0 = _lib.launch_count()
    ms_total, toks = timed(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    barrier()
    ms_e2e, _ = timed(step_e2e, args.steps)
    barrier()
    clk = clocks.stop(t_region0, time.time())
    L = int(toks.shape[1])
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)             # max over ranks
        ms_total, ms_e2e = float(t[0]), float(t[1])
    value = world * B * args.steps / (ms_total / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- untimed profiling pass: per-category kernel time with CUDA events on the launch stream
    roof, kernels, cpu = None, None, None
    if rank == 0:
        pk = peaks()
        _lib.profile_begin()
        dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)
        prof = _lib.profile_end()
        tot = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 4),
                       "tflops": round(v["flops"] / (v["ms"] * 1e9), 3) if v["ms"] > 0 else None,
                       "gbs": round(v["bytes"] / (v["ms"] * 1e6), 1) if v["ms"] > 0 else None}
                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        name, v = top
        if name in ("linear_simt", "gemm_tcgen05"):
            ach = v["flops"] / (v["ms"] * 1e9)
            roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"], "traffic": None,
                    "peak_source": pk["source"] + " (sustained cuBLAS bf16)", "launches": v["launches"],
                    "avg_launch_us": 1e3 * v["ms"] / v["launches"],
                    "algorithmic_flops_per_launch": v["flops"] / v["launches"],
                    # the tensor pipe executes two MMAs (bf16 hi and lo halves of every fp32 activation) per weight
                    # tile, so it is this busy; `achieved` / `frac` count each multiply-add once
                    "frac_executed_on_pipe": (2.0 * ach / pk["bf16_tflops_sustained"]) if name == "gemm_tcgen05" else None}
        else:
            ach = v["bytes"] / (v["ms"] * 1e6)
            roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                    "launches": v["launches"], "avg_launch_us": 1e3 * v["ms"] / v["launches"],
                    "algorithmic_bytes_per_launch": v["bytes"] / v["launches"]}
        # DRAM traffic per launch of the same kernels from the committed ncu capture of this build (profiles/README.md)
        traffic = {}
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)
        except OSError:
            pass
        roof["traffic"] = traffic.get(roof["kernel"], {}).get("dram_bytes_per_launch")
        # the largest HBM-bound kernel beside it (cross-attention streams the projected memory tokens of every row)
        roof_hbm = None
        hb = [(k, v) for k, v in prof.items() if k.startswith("attention") and v["ms"] > 0]
        if hb and roof["bound"] != "hbm":
            k2, v2 = max(hb, key=lambda kv: kv[1]["ms"])
            ach2 = v2["bytes"] / (v2["ms"] * 1e6)
            roof_hbm = {"kernel": k2, "bound": "hbm", "achieved": ach2, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach2 / pk["hbm_gbs"], "traffic": traffic.get(k2, {}).get("dram_bytes_per_launch"),
                        "peak_source": pk["source"], "launches": v2["launches"],
                        "avg_launch_us": 1e3 * v2["ms"] / v2["launches"],
                        "algorithmic_bytes_per_launch": v2["bytes"] / v2["launches"]}
        if world == 1 and args.cpu_rows > 0:
            threads = os.cpu_count() or 1
            cv, cdt, cL = cpu_oracle_run(args.cpu_rows, threads, max_len, steps=L)
            cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_rows} latents of the same workload ({cL} executed steps, {cdt:.1f} s), "
                             f"CPU oracle port, torch fp32, {threads} threads"}
        # small-batch regime (BASELINE config 1 shape: 32 latents, max_len 64): one persistent kernel per decode
        small = None
        if world == 1:
            Bs = 32
            zs, sts = z[:Bs].contiguous(), st[:Bs].contiguous()
            hps = {k: v[:Bs].contiguous() for k, v in hp.items()}
            kws = dict(temperature=0.001, max_len=max_len)          # no masks / stop head: all max_len - 1 steps run
            for _ in range(2):
                dec.generate_with_kv_cache(zs, stoich_pred=sts, heads_pred=hps, **kws)
            n1 = _lib.launch_count()
            reps = 5
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                ts, _, _ = dec.generate_with_kv_cache(zs, stoich_pred=sts, heads_pred=hps, **kws)
            b.record()
            b.synchronize()
            ms_s = a.elapsed_time(b) / reps
            Ls = int(ts.shape[1])
            small = {"workload": f"{Bs} latents, greedy, max_len {max_len}, no masks / stop head ({Ls} executed steps), "
                                 "whole decode in one persistent cooperative kernel (csrc/decode_small.cu)",
                     "ms_per_decode": ms_s, "us_per_step": 1e3 * ms_s / Ls, "formulas_per_s": Bs / (ms_s / 1e3),
                     "launches_per_decode": (_lib.launch_count() - n1) / reps,
                     "hbm_algorithmic_gbs": algorithmic_bytes(Bs, Ls) / (ms_s * 1e6),
                     "hbm_frac": algorithmic_bytes(Bs, Ls) / (ms_s * 1e6) / pk["hbm_gbs"],
                     "note": "bound by the ~100 dependent grid-wide phases of a step (8 per layer), not by HBM: see DESIGN.md section 4"}
        per_gpu_ms = ms_total / args.steps
        step_roof = {
            "hbm_algorithmic_gbs": algorithmic_bytes(B, L) / (per_gpu_ms * 1e6),
            "hbm_frac": algorithmic_bytes(B, L) / (per_gpu_ms * 1e6) / pk["hbm_gbs"],
            "tensor_algorithmic_tflops": algorithmic_flops(B, L) / (per_gpu_ms * 1e9),
            "tensor_frac": algorithmic_flops(B, L) / (per_gpu_ms * 1e9) / pk["bf16_tflops_sustained"],
            "note": "SURVEY 8d cost model over the executed steps, fp32 KV cache (49,152 B per cached position)"}
        emit(({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 weights, fp32 activations/accumulate/KV", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "max_len": max_len,
                       "executed_decode_steps": L, "parallelism": f"dp{world} (latents sharded, weights replicated)",
                       "l2": "working set (KV pages + projected memory > 10 GB) exceeds L2; a 256 MiB buffer is "
                             "also rewritten between timed iterations"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": B * L * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_hbm_kernel": roof_hbm,
            "step_roofline": step_roof, "small_batch": small,
            "kernels": kernels, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="latents per GPU")
    ap.add_argument("--max-len", type=int, default=64)
    ap.add_argument("--cpu-rows", type=int, default=1536, help="rows of the bounded CPU-baseline sample (10-20 s of CPU work)")
    ap.add_argument("--ncu", action="store_true", help="short run for ncu: warm-up decodes + one decode, no timing")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
