"""One beam-search pass over a full wave (for ncu captures of the decoder kernels): python tools/run_beam_once.py [beam] [n] [L]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import ravvent_basecaller_b200 as rb

beam = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
L = int(sys.argv[3]) if len(sys.argv) > 3 else 6
raw, ev = bench.synth_range(0, n)
bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0.).load_weights(seed=22)
ids, sc = bc.beam_search_prediction((torch.from_numpy(raw).cuda(), torch.from_numpy(ev).cuda()), beam, L)
torch.cuda.synchronize()
print("ok", tuple(ids.shape))
