"""Kernel-only timing of the attention launches of one beam-search pass over a full wave.

    python tools/time_attention.py [beam] [n_chunks] [max_output_len] [fp32|bf16]

Times come from the library's own CUDA events (rvb_profile), i.e. kernel time on the launching stream.
"""
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
a = prof["attention"]
print(f"{precision} beam {beam}, {n} snippets: attention {a['ms'] / a['launches']:.4f} ms per launch over {a['launches']} launches", flush=True)
