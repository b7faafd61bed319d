"""BASELINE configs[1] alone (bench.py's event_only pipeline): wall time per pass and the kernels' share."""
import sys
import types
sys.path.insert(0, ".")
import bench

args = types.SimpleNamespace(steps=6, precision=sys.argv[1] if len(sys.argv) > 1 else "fp32")
import torch
tm = bench.Timer(torch.device('cuda', 0), 0, None)
r = bench.event_only_bench(args, 0, tm)
print(f"event_only: {r['ms_per_step']:.1f} ms per pass, kernels {sum(r['kernel_ms_per_step'].values()):.1f} ms: {r['kernel_ms_per_step']}", flush=True)
