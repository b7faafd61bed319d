"""One encoder pass over a full wave (for ncu captures of K2 / K3): python tools/run_encoder_once.py [kind] [depth] [n]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import model_ref as mr
import ravvent_basecaller_b200 as rb

kind = sys.argv[1] if len(sys.argv) > 1 else "raw"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 9472
raw, ev = mr.synth_chunks(np.random.default_rng(1), n)
x = {"raw": torch.from_numpy(raw).cuda(), "event": torch.from_numpy(ev).cuda()}
x["joint"] = (x["raw"], x["event"])
w = mr.init_weights(22, encoder_depth=depth, random_bias=True)
bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, kind, 0., encoder_depth=depth).load_weights(w)
enc, mask = bc._encode_input(x[kind])
torch.cuda.synchronize()
print("ok", tuple(enc.shape), float(enc.float().abs().mean()))
