"""Kernel-only timing of the recurrent LSTM (K3) on one full wave: per-layer time per timestep.

    python tools/time_encoder.py [n_chunks]

Depth-1 raw encoder = the layer-0 launch alone; depth 2 adds the projection GEMM and the pre-gate launch.
Times come from the library's own CUDA events (rvb_profile), i.e. kernel time on the launching stream.
"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import model_ref as mr
import ravvent_basecaller_b200 as rb
from ravvent_basecaller_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
raw, ev = mr.synth_chunks(np.random.default_rng(1), n)
raw_d, ev_d = torch.from_numpy(raw).cuda(), torch.from_numpy(ev).cuda()
for kind, x, T in (("raw", raw_d, 200), ("event", ev_d, 30)):
    prev = 0.0
    for depth in (1, 2):
        w = mr.init_weights(22, encoder_depth=depth, random_bias=True)
        bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, kind, 0., encoder_depth=depth).load_weights(w)
        for _ in range(2):
            bc._encode_input(x)
        torch.cuda.synchronize()
        _lib.profile(True)
        reps = 5
        for _ in range(reps):
            bc._encode_input(x)
        prof = _lib.profile_read(); _lib.profile(False)
        rec = prof["recurrent_lstm"]["ms"] / reps
        gemm = prof["projection_gemm"]["ms"] / reps
        layer = rec - prev
        print(f"{kind} depth {depth}: K3 total {rec:.3f} ms, this layer {layer:.3f} ms = {1e3 * layer / T:.2f} us/timestep; K2 {gemm:.3f} ms", flush=True)
        prev = rec
        del bc
