"""Latency of small batches (host-side launch cost shows here): greedy search on the raw model, B chunks, S = 33."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import ravvent_basecaller_b200 as rb

bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "raw", 0.).load_weights(seed=22)
for B in (8, 1000, 9472):
    raw = torch.from_numpy(bench.synth_range(0, B)[0]).cuda()
    for _ in range(3):
        bc.greedy_search_prediction(raw, 34)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        bc.greedy_search_prediction(raw, 34)
    torch.cuda.synchronize()
    print(f"B={B}: {1e3 * (time.perf_counter() - t0) / n:.2f} ms per call", flush=True)
