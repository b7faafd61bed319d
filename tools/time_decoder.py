"""Kernel-only time of the decoder's dense phases (cell GEMM, query GEMM, attention-layer GEMM, search) of one beam-search
pass over a full wave:    python tools/time_decoder.py [beam] [n_chunks] [max_output_len] [fp32|bf16]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
import ravvent_basecaller_b200 as rb
from ravvent_basecaller_b200 import _lib

beam = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
L = int(sys.argv[3]) if len(sys.argv) > 3 else 6
precision = sys.argv[4] if len(sys.argv) > 4 else "fp32"
raw, ev = bench.synth_range(0, n)
x = (torch.from_numpy(raw).cuda(), torch.from_numpy(ev).cuda())
bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., precision=precision).load_weights(seed=22)
for _ in range(2):
    bc.beam_search_prediction(x, beam, L)
torch.cuda.synchronize()
_lib.profile(True)
reps = 3
for _ in range(reps):
    bc.beam_search_prediction(x, beam, L)
prof = _lib.profile_read(); _lib.profile(False)
d, a = prof["decoder"], prof["attention"]
steps = a["launches"]
print(f"{precision} beam {beam}, {n} snippets: dense decoder phases {1e3 * d['ms'] / steps:.1f} us per decode step "
      f"({d['launches']} launches over {steps} decode steps); attention {1e3 * a['ms'] / steps:.1f} us", flush=True)
