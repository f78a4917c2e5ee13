"""Micro-benchmark of the tcgen05 projection kernel on the decode-step shapes (not a pytest file).
python tests/gemm_bench.py [M ...]   -- env switches: SCV_GEMM_BN, SCV_PDL"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superconductor_vae_b200 import _lib

DEV = "cuda:0"
L = _lib.lib()
SHAPES = [("qkv", 1536, 512, False, False), ("q", 512, 512, False, False), ("o+res", 512, 512, True, False),
          ("ff1->split", 2048, 512, False, True), ("ff2+res", 512, 2048, True, False), ("logits", 4752, 512, False, False)]


def bench(M, name, N, K, residual, split_out, iters=50):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((M, K), generator=g).to(DEV)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn((N,), generator=g).to(DEV)
    st = _lib.current_stream()
    wt = torch.zeros(int(L.scv_op_tiled_elems(N, K)), dtype=torch.bfloat16, device=DEV)
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(w), _lib.ptr(wt), N, K, st))
    xs = torch.zeros(int(L.scv_op_split_tile_bytes(M, K)), dtype=torch.uint8, device=DEV)
    if K <= 1024:      # wider rows: time with an all-zero SplitTile operand (tensor-core timing is data independent)
        _lib.check(L.scv_op_split_rows(_lib.ptr(x), K, None, None, _lib.ptr(xs), M, K, 0, st))
    y = torch.zeros((M, N), device=DEV)
    ys = torch.zeros(int(L.scv_op_split_tile_bytes(M, N)), dtype=torch.uint8, device=DEV) if split_out else None
    r = y if residual else None

    def run():
        _lib.check(L.scv_op_linear_split(_lib.ptr(xs), _lib.ptr(wt), _lib.ptr(b), _lib.ptr(r) if residual else None, N,
                                         None if split_out else _lib.ptr(y), N, _lib.ptr(ys) if split_out else None,
                                         M, N, K, int(os.environ.get("GB_ACT", "1")) if split_out else 0, st))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    err = None
    if not split_out and not residual and K <= 1024:      # check against fp64 with the bf16-rounded weights
        ref = (x.double() @ w.to(torch.bfloat16).double().t() + b.double())
        err = float((y.double() - ref).abs().max())
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run()
    e.record(); e.synchronize()
    us = a.elapsed_time(e) * 1e3 / iters
    return us, 2.0 * M * N * K / (us * 1e6), err


if __name__ == "__main__":
    if os.environ.get("GB_SHAPES") == "mem":      # the memory builder's two big projections (BASELINE config 5)
        for M in [int(a) for a in sys.argv[1:]] or [52736]:
            for (name, N, K, res, so) in (("l2m.0+gelu->split", 4096, 2048, False, True), ("l2m.2", 8192, 4096, False, False)):
                us, tf, err = bench(M, name, N, K, res, so, iters=5)
                print(f"M={M:6d} {name:18s} N={N:5d} K={K:5d}: {us:9.1f} us  {tf:7.1f} TFLOP/s(alg)", flush=True)
        sys.exit(0)
    Ms = [int(a) for a in sys.argv[1:]] or [2048, 4096]
    for M in Ms:
        tot = 0.0
        for (name, N, K, res, so) in SHAPES:
            us, tf, err = bench(M, name, N, K, res, so)
            w = {"q": 1, "o+res": 2}.get(name, 1)
            tot += us * w
            print(f"M={M:5d} {name:11s} N={N:5d} K={K:5d}: {us:7.2f} us  {tf:7.1f} TFLOP/s(alg)" + (f"  max|err| {err:.2e}" if err is not None else ""), flush=True)
        print(f"M={M}: one layer (qkv + q + 2*o + ff1 + ff2) + logits/12 ~ {tot - bench(M, *SHAPES[5])[0] * 11 / 12:7.1f} us "
              f"[BN={os.environ.get('SCV_GEMM_BN', '128')}]", flush=True)
