import sys, math, torch
sys.path.insert(0, '.')
from superconductor_vae_b200 import _lib
L = _lib.lib(); DEV = "cuda:0"
for (M, N, K) in ((4096, 512, 512), (4096, 512, 2048)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((M, K), generator=g)
    w = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(torch.bfloat16).float()
    b = torch.randn((N,), generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    xd, wd, bd = x.to(DEV), w.to(DEV).contiguous(), b.to(DEV)
    wt = torch.zeros(int(L.scv_op_tiled_elems(N, K)), dtype=torch.bfloat16, device=DEV)
    _lib.check(L.scv_op_pack_tiled(_lib.ptr(wd), _lib.ptr(wt), N, K, _lib.current_stream()))
    y = torch.empty((M, N), device=DEV)
    _lib.check(L.scv_op_linear(_lib.ptr(xd), K, _lib.ptr(wt), 0, _lib.ptr(bd), None, N, _lib.ptr(y), N, M, N, K, 0, 2, _lib.current_stream()))
    torch.cuda.synchronize()
    e = (y.cpu().double() - ref).abs()
    f32 = (torch.nn.functional.linear(x, w, b).double() - ref).abs()
    print(f"M={M} N={N} K={K}: tensor path max |err| {float(e.max()):.3e} rms {float(e.pow(2).mean().sqrt()):.3e}; torch fp32 CPU max {float(f32.max()):.3e} rms {float(f32.pow(2).mean().sqrt()):.3e}")
