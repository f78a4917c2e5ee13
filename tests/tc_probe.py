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
    sys.exit(0 if worst < 5e-5 else 1)
