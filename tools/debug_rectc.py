"""Diagnostic for the tensor-core recurrence: depth-1 raw encoder, few timesteps, error maps."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravvent_basecaller_b200 as rb
from oracle import model_ref as mr

w = mr.init_weights(7, encoder_depth=1, random_bias=True)
for kind, F in (("raw", 1), ("event", 5)):
    for T in (1, 2, 3, 8):
        B = 256
        rng = np.random.default_rng(T)
        x = rng.normal(size=(B, T, F)).astype(np.float32)
        bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, kind, 0., encoder_depth=1)
        bc.load_weights(w)
        enc, _ = bc._encode_input(x)
        ref, _ = mr.encode_input(w, x, kind, encoder_depth=1)
        nan = np.isnan(enc)
        err = np.abs(np.nan_to_num(enc, nan=1e3) - ref)
        print(f"== {kind} T={T}: nan frac {nan.mean():.4f}  max err {err.max():.3e}")
        if err.max() > 1e-3:
            for d, dn in ((0, "fwd"), (1, "bwd")):
                e = err[:, :, d * 128:(d + 1) * 128]
                tab = [[e[r0:r0 + 32, :, h * 64:(h + 1) * 64].max() for h in (0, 1)] for r0 in range(0, 256, 32)]
                print(f"   {dn}: max err per (row block of 32 | unit half):", " ".join(f"[{a:.1e} {b:.1e}]" for a, b in tab))
                print(f"   {dn}: err by t:", [float(f"{e[:, t].max():.2e}") for t in range(T)])
                bad = np.argwhere(e > 1e-3)
                print(f"   {dn}: first bad (row,t,unit):", bad[:6].tolist(), " n_bad", len(bad), "of", e.size)
                print("   sample got/ref:", enc[0, 0, d * 128:d * 128 + 4], ref[0, 0, d * 128:d * 128 + 4])
