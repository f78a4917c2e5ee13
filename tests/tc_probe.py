"""Stand-alone probe of the tcgen05 linear kernel against fp64 math (run under `timeout`; not a pytest file)."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superconductor_vae_b200 import _lib

DEV = "cuda:0"
L = _lib.lib()


def run(M, N, K, act=0, residual=True, seed=0):
    g = torch.Generator().manual_seed(seed + M + 3 * N + 7 * K)
    x = torch.randn((M, K), generator=g)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(torch.bfloat16).float()
    b = torch.randn((N,), generator=g)
    r = torch.randn((M, N), generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    if residual:
        ref = ref + r.double()
    xd, wd, bd, rd = x.to(DEV), w.to(DEV).contiguous(), b.to(DEV), r.to(DEV)
    wt = torch.zeros(int(L.scv_op_tiled_elems(N, K)), dtype=torch.bfloat16, device=DEV)
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(wd), _lib.ptr(wt), N, K, _lib.current_stream()))
    y = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(L.scv_op_linear(_lib.ptr(xd), K, _lib.ptr(wt), 0, _lib.ptr(bd), _lib.ptr(rd) if residual else None, N,
                               _lib.ptr(y), N, M, N, K, act, 2, _lib.current_stream()))
    torch.cuda.synchronize()
    err = (y.cpu().double() - ref).abs().max().item()
    return err


if __name__ == "__main__":
    shapes = [(128, 128, 64), (128, 128, 512), (256, 512, 512), (4096, 1536, 512), (4096, 512, 2048), (200, 132, 72),
              (64, 4752, 512), (1000, 2048, 576), (4096, 4752, 512)]
    worst = 0.0
    for i, (M, N, K) in enumerate(shapes):
        e = run(M, N, K, act=i % 2, residual=bool(i % 3))
        worst = max(worst, e)
        print(f"M={M} N={N} K={K}: max abs err {e:.3e}", flush=True)
    print("WORST", worst)
    if worst >= 5e-5:
        sys.exit(1)


def run_split_chain(M, K, H, N, seed=0):
    """y = W2 * gelu(LN(x) * W1^T + b1) + b2 + r  with LN -> SplitTile -> GEMM(SplitTile out) -> GEMM."""
    g = torch.Generator().manual_seed(seed + M + K + H + N)
    x = torch.randn((M, K), generator=g) * 2 + 0.5
    ga, be = 1 + 0.1 * torch.randn((K,), generator=g), 0.1 * torch.randn((K,), generator=g)
    w1 = (torch.randn((H, K), generator=g) / math.sqrt(K)).to(torch.bfloat16).float()
    w2 = (torch.randn((N, H), generator=g) / math.sqrt(H)).to(torch.bfloat16).float()
    b1, b2, r = torch.randn((H,), generator=g), torch.randn((N,), generator=g), torch.randn((M, N), generator=g)
    xn = torch.nn.functional.layer_norm(x.double(), (K,), ga.double(), be.double(), 1e-5)
    ref = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(xn, w1.double(), b1.double())),
                                     w2.double(), b2.double()) + r.double()
    dev = lambda t: t.to(DEV).contiguous()
    xd, gd, bd, w1d, w2d, b1d, b2d, rd = map(dev, (x, ga, be, w1, w2, b1, b2, r))
    w1t = torch.zeros(int(L.scv_op_tiled_elems(H, K)), dtype=torch.bfloat16, device=DEV)
    w2t = torch.zeros(int(L.scv_op_tiled_elems(N, H)), dtype=torch.bfloat16, device=DEV)
    st = _lib.current_stream()
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(w1d), _lib.ptr(w1t), H, K, st))
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(w2d), _lib.ptr(w2t), N, H, st))
    xs = torch.zeros(int(L.scv_op_split_tile_bytes(M, K)), dtype=torch.uint8, device=DEV)
    hs = torch.zeros(int(L.scv_op_split_tile_bytes(M, H)), dtype=torch.uint8, device=DEV)
    y = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(L.scv_op_split_rows(_lib.ptr(xd), K, _lib.ptr(gd), _lib.ptr(bd), _lib.ptr(xs), M, K, 1, st))
    _lib.check(L.scv_op_linear_split(_lib.ptr(xs), _lib.ptr(w1t), _lib.ptr(b1d), None, 0, None, 0, _lib.ptr(hs), M, H, K, 1, st))
    _lib.check(L.scv_op_linear_split(_lib.ptr(hs), _lib.ptr(w2t), _lib.ptr(b2d), _lib.ptr(rd), N, _lib.ptr(y), N, None, M, N, H, 0, st))
    torch.cuda.synchronize()
    return (y.cpu().double() - ref).abs().max().item()


if __name__ == "__main__":
    worst = 0.0
    for (M, K, H, N) in [(128, 64, 128, 128), (4096, 512, 2048, 512), (300, 576, 144, 4752), (4096, 512, 512, 4752),
                         (64, 512, 128, 8)]:
        e = run_split_chain(M, K, H, N)
        worst = max(worst, e)
        print(f"chain M={M} K={K} H={H} N={N}: max abs err {e:.3e}", flush=True)
    print("WORST chain", worst)
    sys.exit(0 if worst < 1e-4 else 1)
