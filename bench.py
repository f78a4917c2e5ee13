#!/usr/bin/env python
"""Benchmark of the KV-cache decode hot path (BASELINE.json metric: formulas/sec, KV-cache decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]

One "step" = one full greedy decode of a batch of synthetic latents (BASELINE config 2: 4096 latents per GPU,
max_len 64, type masks + stop head + hard stop, temperature 0.001, 24 memory tokens).  Weak scaling: every
rank decodes its own 4096 latents with replicated weights; the only collective is the gather of token ids, done by
the product's own multi-GPU API (superconductor_vae_b200.parallel.generate_sharded / gather_rows).

`value`   : formulas/s with the conditioning inputs already resident in HBM.
`e2e`     : same call through the public Python API with HOST (pinned) inputs: H2D of z / stoich / heads and
            D2H of the token ids inside the timed region.
`roofline`: the kernel category with the largest share of the step.  Timed in a separate (untimed) profiling pass in
            which the step still replays as a CUDA graph whose event-record nodes bracket every kernel (no host launch
            latency inside an interval); the cost of an empty event pair, measured in the same graph, is subtracted;
            only executed steps are counted.  Peaks from MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the reference's own `generate_with_kv_cache` (the unmodified module from
            oracle/_ref, copied from /root/reference by oracle/make_ref.py at build time) on a bounded sample with all
            host threads; when oracle/_ref is absent the golden-pinned oracle port is timed instead (kind "port").
`gpu_eager_baseline`: info only: the same reference module run eagerly on the B200, fp32 and autocast(bf16).
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
GEN_KW = dict(temperature=0.001, stop_boost=10.0, hard_stop_threshold=0.8)


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


# Steps the full 4096-latent batch executes before every row has emitted END (deterministic for the seeded synthetic
# weights / inputs: 3203 rows end at step 22, 893 at step 21, so any sample of a few hundred rows runs 22 steps too).
FULL_BATCH_STEPS = 22


# --------------------------------------------------------------------------------------------- CPU / reference legs
def _sample_inputs(rows, seed=1234):
    from superconductor_vae_b200 import synthetic as Sy
    z = Sy.make_latents(rows, 2048, seed)
    stoich, heads = Sy.make_conditioning(rows, 13, seed)
    return z, stoich, heads


def reference_available():
    from oracle import ref_loader
    return ref_loader.available()


def cpu_reference_run(rows, threads, max_len=64, device="cpu", autocast=False):
    """Time the reference's OWN generate_with_kv_cache (oracle/_ref, unmodified) on `rows` config-2 latents."""
    import torch
    from oracle import ref_loader, vocab as OV
    from superconductor_vae_b200 import synthetic as Sy
    if device == "cpu":
        torch.set_num_threads(threads)
    shape = Sy.C512
    dec = ref_loader.reference_decoder(shape, Sy.make_decoder_state_dict(shape, 0)).to(device)
    z, stoich, heads = _sample_inputs(rows)
    z, stoich, heads = z.to(device), stoich.to(device), {k: v.to(device) for k, v in heads.items()}
    masks = OV.type_masks().to(device)

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return dec.generate_with_kv_cache(z, stoich_pred=stoich, max_len=max_len, heads_pred=heads, type_masks=masks,
                                              **GEN_KW)[0]
    if device != "cpu":
        run()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    tok = run()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return rows / dt, dt, int(tok.shape[1])


def cpu_oracle_run(rows, threads, max_len=64, steps=FULL_BATCH_STEPS):
    """Fallback when oracle/_ref is absent: the golden-pinned oracle port on the same sample."""
    import torch
    from oracle import decoder_oracle as DO, vocab as OV
    from superconductor_vae_b200 import synthetic as Sy
    torch.set_num_threads(threads)
    sd = Sy.make_decoder_state_dict(Sy.C512, 0)
    z, stoich, heads = _sample_inputs(rows)
    masks = OV.type_masks()
    t0 = time.perf_counter()
    tok, _, _ = DO.generate_with_kv_cache(sd, 8, z, stoich_pred=stoich, max_len=max_len, heads_pred=heads, type_masks=masks,
                                          stop_when_all_finished=False, max_steps=steps, **GEN_KW)
    dt = time.perf_counter() - t0
    return rows / dt, dt, int(tok.shape[1])


def cpu_leg(rows, threads, max_len):
    if reference_available():
        v, dt, L = cpu_reference_run(rows, threads, max_len)
        kind, what = "reference", "the reference's own EnhancedTransformerDecoder.generate_with_kv_cache (oracle/_ref, unmodified)"
    else:
        v, dt, L = cpu_oracle_run(rows, threads, max_len)
        kind, what = "port", "CPU oracle port of the reference path (oracle/_ref not built)"
    sample = (f"{rows} latents of the config-2 workload (the 4096-latent batch bounded to {rows} rows for the CPU), {L} executed "
              f"steps, {dt:.1f} s, {what}, torch fp32, {threads} threads")
    return v, dt, L, kind, sample


def cpu_config1(threads):
    """BASELINE config 1 on the host: V14.3 encode + greedy KV-cache decode, batch 32, max_len 64, fp32, the reference's
    own modules; with masks + stop head, and without (all 63 steps run)."""
    import torch
    from oracle import ref_loader, vocab as OV
    from superconductor_vae_b200 import synthetic as Sy
    torch.set_num_threads(threads)
    _, Enc, _ = ref_loader.load()
    enc = Enc()
    enc.load_state_dict(Sy.make_encoder_state_dict(Sy.ENC_DEFAULT, 1), strict=True)
    enc.eval()
    dec = ref_loader.reference_decoder(Sy.C512, Sy.make_decoder_state_dict(Sy.C512, 0))
    idx, frac, mask, magpie, tc = Sy.make_compositions(32, 7)
    masks = OV.type_masks()
    out = {}
    with torch.no_grad():
        for name, kw in (("masks_and_stop_head", dict(type_masks=masks, **GEN_KW)), ("plain_63_steps", dict(temperature=0.001))):
            t0 = time.perf_counter()
            o = enc(idx, frac, mask, magpie, tc)
            stoich = torch.cat([o["fraction_pred"], o["element_count_pred"].unsqueeze(-1)], dim=-1)
            heads = {k: o[k] for k in ("tc_pred", "sc_pred", "hp_pred", "tc_class_logits", "competence", "element_count_pred",
                                       "family_composed_14")}
            t, _, _ = dec.generate_with_kv_cache(o["z"], stoich_pred=stoich, max_len=64, heads_pred=heads, **kw)
            dt = time.perf_counter() - t0
            out[name] = {"seconds": dt, "formulas_per_s": 32 / dt, "executed_steps": int(t.shape[1])}
    out["what"] = f"BASELINE config 1: reference FullMaterialsVAE encode + greedy KV-cache decode, batch 32, max_len 64, fp32, {threads} host threads"
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = args.cpu_rows
    vals, kind, sample = [], "port", ""
    for i in range(args.warmup + args.steps):
        v, dt, L, kind, sample = cpu_leg(rows, threads, args.max_len)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# --------------------------------------------------------------------------------------------- engine arm
def run_engine(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import superconductor_vae_b200 as S
    from superconductor_vae_b200 import _lib, parallel, synthetic as Sy
    from superconductor_vae_b200.tokenizer import FractionAwareTokenizer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, max_len = args.batch, args.max_len
    N = B * world                                         # weak scaling: B latents per rank, N in the job
    sd = Sy.make_decoder_state_dict(Sy.C512, 0)
    dec = S.EnhancedTransformerDecoder.from_state_dict(sd, nhead=8, device=dev)
    dec.max_rows_per_call = max(dec.max_rows_per_call, B)
    tok = FractionAwareTokenizer(max_len=max_len, fractions=[f"{i + 1}/100003" for i in range(4317)],
                                 isotopes=[f"{300 + i}Og" for i in range(291)])
    masks = tok.get_type_masks(dev)
    # The job's inputs: N rows, block r (seed 1234 + r) is rank r's share.  Every rank holds the whole [N, ...] arrays
    # (what parallel.generate_sharded expects) on the device for `value`, and its own block in pinned host memory for `e2e`.
    blocks = [(Sy.make_latents(B, 2048, 1234 + r),) + Sy.make_conditioning(B, 13, 1234 + r) for r in range(world)]
    z_all = torch.cat([b[0] for b in blocks]).to(dev)
    st_all = torch.cat([b[1] for b in blocks]).to(dev)
    hp_all = {k: torch.cat([b[2][k] for b in blocks]).to(dev) for k in blocks[0][2]}
    z_h, st_h = blocks[rank][0].pin_memory(), blocks[rank][1].pin_memory()
    hp_h = {k: v.pin_memory() for k, v in blocks[rank][2].items()}
    lo = rank * B
    z, st, hp = z_all[lo:lo + B], st_all[lo:lo + B], {k: v[lo:lo + B] for k, v in hp_all.items()}
    kw = dict(max_len=max_len, type_masks=masks, **GEN_KW)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    tok_host = torch.empty((B, max_len - 1), dtype=torch.int32).pin_memory()
    h2d_bytes = z_h.numel() * 4 + st_h.numel() * 4 + sum(v.numel() * 4 for v in hp_h.values())

    def step_resident():
        if world > 1:       # the product's multi-GPU call: decode the rank's rows, NCCL gather of int16 sequences
            t, _, _ = parallel.generate_sharded(dec, z_all, stoich_pred=st_all, heads_pred=hp_all, **kw)
            return t
        return dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)[0]

    def step_e2e():
        zd = z_h.to(dev, non_blocking=True)
        sd_ = st_h.to(dev, non_blocking=True)
        hd = {k: v.to(dev, non_blocking=True) for k, v in hp_h.items()}
        t, _, _ = dec.generate_with_kv_cache(zd, stoich_pred=sd_, heads_pred=hd, **kw)
        if world > 1:
            parallel.gather_rows(t.to(torch.int16), N, 0)
        tok_host[:, :t.shape[1]].copy_(t.to(torch.int32), non_blocking=True)
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
    value = N * args.steps / (ms_total / 1e3)
    e2e = N * args.steps / (ms_e2e / 1e3)

    roof = roof_hbm = kernels = cpu = small = eager = None
    if rank == 0:
        pk = peaks()
        # ---- untimed profiling pass (one decode of this rank's rows): see the module docstring
        _lib.profile_begin()
        tp = dec.generate_with_kv_cache(z, stoich_pred=st, heads_pred=hp, **kw)[0]
        prof = _lib.profile_end()
        Lp = int(tp.shape[1])
        ov = prof.pop("event_pair_overhead", None)
        ov_ms = (ov["ms"] / ov["launches"]) if ov and ov["launches"] else 0.0
        for v in prof.values():                              # subtract what an empty event pair measures in the same graph
            v["ms_raw"] = v["ms"]
            v["ms"] = max(v["ms"] - v["launches"] * ov_ms, 1e-6)
        tot = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "ms_uncorrected": round(v["ms_raw"], 3),
                       "share": round(v["ms"] / tot, 4), "avg_launch_us": round(1e3 * v["ms"] / v["launches"], 2),
                       "tflops": round(v["flops"] / (v["ms"] * 1e9), 3), "gbs_own_traffic": round(v["bytes"] / (v["ms"] * 1e6), 1)}
                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)
        except OSError:
            pass

        def tensor_roof(name, v):
            ach = v["flops"] / (v["ms"] * 1e9)
            return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic.get(name, {}).get("dram_bytes_per_launch"),
                    "peak_source": pk["source"] + " (sustained cuBLAS bf16)", "launches": v["launches"],
                    "avg_launch_us": 1e3 * v["ms"] / v["launches"], "algorithmic_flops_per_launch": v["flops"] / v["launches"],
                    # the tensor pipe executes two MMAs (bf16 hi and lo halves of every fp32 activation) per weight tile:
                    # `achieved` / `frac` count each multiply-add once, the pipe is this busy
                    "frac_executed_on_pipe": (2.0 * ach / pk["bf16_tflops_sustained"]) if name == "gemm_tcgen05" else None}

        def hbm_roof(name, v):
            own = v["bytes"] / v["launches"]
            alg = own
            note = "algorithmic bytes = the K / V rows of every attended position + q, out and the appended row (fp32)"
            if name == "attention_cross":
                # SURVEY 8d counts the memory tokens once per sequence for the WHOLE decode and no projected K / V at all:
                # per launch that leaves q in, out, and the raw tokens amortised over layers x executed steps
                alg = B * 512 * 4 * 2 + B * 24 * 512 * 4 / (12.0 * Lp)
                note = ("algorithmic bytes per SURVEY 8d (q, out, raw memory tokens amortised over the decode); the kernel "
                        "really streams the per-layer projected K / V of the 24 tokens (`own_traffic_bytes_per_launch`), "
                        "which that model counts as zero: DESIGN.md section 4 explains why they are kept")
            ach = alg * v["launches"] / (v["ms"] * 1e6)
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": traffic.get(name, {}).get("dram_bytes_per_launch"), "peak_source": pk["source"],
                    "launches": v["launches"], "avg_launch_us": 1e3 * v["ms"] / v["launches"],
                    "algorithmic_bytes_per_launch": alg, "own_traffic_bytes_per_launch": own,
                    "own_traffic_gbs": own * v["launches"] / (v["ms"] * 1e6), "note": note}

        name, v = max(prof.items(), key=lambda kv: kv[1]["ms"])
        roof = tensor_roof(name, v) if name in ("linear_simt", "gemm_tcgen05") else hbm_roof(name, v)
        # the other regime beside it: largest attention kernel when the projections dominate, else the projections
        if roof["bound"] == "tensor":
            hb = [(k, v2) for k, v2 in prof.items() if k.startswith("attention")]
            if hb:
                k2, v2 = max(hb, key=lambda kv: kv[1]["ms"])
                roof_hbm = hbm_roof(k2, v2)
        elif "gemm_tcgen05" in prof:
            roof_hbm = tensor_roof("gemm_tcgen05", prof["gemm_tcgen05"])

        threads = os.cpu_count() or 1
        if world == 1 and args.cpu_rows > 0:
            cv, cdt, cL, kind, sample = cpu_leg(args.cpu_rows, threads, max_len)
            cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
            if reference_available() and not args.no_extras:
                try:
                    cpu["config1"] = cpu_config1(threads)
                except Exception as e:                       # informational only
                    cpu["config1"] = {"error": repr(e)[:200]}
        # ---- info only: the unmodified reference module, eager PyTorch on this same B200 (SURVEY 8d second baseline)
        if world == 1 and reference_available() and not args.no_extras:
            try:
                rows_e = min(B, 4096)
                f32 = cpu_reference_run(rows_e, threads, max_len, device=dev)
                b16 = cpu_reference_run(rows_e, threads, max_len, device=dev, autocast=True)
                eager = {"what": f"reference EnhancedTransformerDecoder.generate_with_kv_cache, eager PyTorch on the same GPU, {rows_e} "
                                 "latents of the same workload (not the optimisation target; autocast(bf16) is not token-exact)",
                         "fp32": {"formulas_per_s": f32[0], "seconds": f32[1], "executed_steps": f32[2]},
                         "autocast_bf16": {"formulas_per_s": b16[0], "seconds": b16[1], "executed_steps": b16[2]}}
            except Exception as e:
                eager = {"error": repr(e)[:200]}
        # ---- small-batch regime (BASELINE config 1 shape: 32 latents, max_len 64): one persistent kernel per decode
        if world == 1 and not args.no_extras:
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
            small = {"workload": f"{Bs} latents, greedy, max_len {max_len}, no masks / stop head ({Ls} executed steps), whole decode "
                                 "in one persistent kernel (csrc/decode_small*.cu)",
                     "ms_per_decode": ms_s, "us_per_step": 1e3 * ms_s / Ls, "formulas_per_s": Bs / (ms_s / 1e3),
                     "launches_per_decode": (_lib.launch_count() - n1) / reps,
                     "hbm_algorithmic_gbs": algorithmic_bytes(Bs, Ls) / (ms_s * 1e6),
                     "hbm_frac": algorithmic_bytes(Bs, Ls) / (ms_s * 1e6) / pk["hbm_gbs"]}
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
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": N, "max_len": max_len,
                       "executed_decode_steps": L, "parallelism": f"dp{world} (latents sharded, weights replicated; "
                       "parallel.generate_sharded: NCCL all-gather of int16 token ids)",
                       "l2": "working set (KV pages + projected memory > 10 GB) exceeds L2; a 256 MiB buffer is "
                             "also rewritten between timed iterations"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": B * L * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_other_regime": roof_hbm,
            "step_roofline": step_roof, "small_batch": small, "profile_pass": {"executed_steps": Lp, "event_pair_overhead_us": 1e3 * ov_ms},
            "kernels": kernels, "cpu_baseline": cpu, "gpu_eager_baseline": eager}))
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
    ap.add_argument("--cpu-rows", type=int, default=1536, help="rows of the bounded CPU-baseline sample (10-30 s of CPU work)")
    ap.add_argument("--ncu", action="store_true", help="short run for ncu: warm-up decodes + one decode, no timing")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational legs (config 1 on CPU, eager GPU, small batch)")
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
