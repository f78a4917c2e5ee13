"""Probe helper (not a test): compares gpurun_out/small_probe_{0,1}.pt written by tests/small_probe.py."""
import torch
a = torch.load("gpurun_out/small_probe_0.pt")
b = torch.load("gpurun_out/small_probe_1.pt")
print("tokens equal:", torch.equal(a["t"], b["t"]), "x max diff", float((a["x"] - b["x"]).abs().max()),
      "logits max diff", float((a["lg"] - b["lg"]).abs().max()), "logits scale", float(a["lg"].abs().max()))
