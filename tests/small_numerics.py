"""Probe (not a test): numerical distance of the three projection paths at B rows after `steps` decode steps.
  a: fp32 CUDA-core kernels (SCV_LINEAR_IMPL=1, SCV_SMALL=0)   -- every product exact in the FMA
  b: tcgen05 hi/lo kernels (SCV_SMALL=0, SCV_TC_MIN_ROWS=1)
  c: persistent small-batch kernel (mma.sync hi/lo)
usage: python tests/small_numerics.py [B] [steps]"""
import os, subprocess, sys
import torch
B = sys.argv[1] if len(sys.argv) > 1 else "32"
steps = sys.argv[2] if len(sys.argv) > 2 else "12"
here = os.path.dirname(os.path.abspath(__file__))
cfgs = {"a": {"SCV_LINEAR_IMPL": "1", "SCV_SMALL": "0"}, "b": {"SCV_SMALL": "0", "SCV_TC_MIN_ROWS": "1"},
        "c": {"SCV_SMALL": "1", "SCV_SMALL_PERSIST": "0"}}
out = {}
for k, env in cfgs.items():
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(here, "small_probe.py"), B, steps], env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out[k] = torch.load(f"gpurun_out/small_probe_{env['SCV_SMALL']}.pt")
for k in ("b", "c"):
    dx = (out[k]["x"] - out["a"]["x"]).abs()
    dl = (out[k]["lg"] - out["a"]["lg"]).abs()
    print(f"{k} vs a: tokens equal {torch.equal(out[k]['t'], out['a']['t'])}; x max {dx.max():.3e} mean {dx.mean():.3e}; "
          f"logits max {dl.max():.3e} mean {dl.mean():.3e} (|logits| max {out['a']['lg'].abs().max():.3f})")
