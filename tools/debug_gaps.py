import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravvent_basecaller_b200 as rb
from ravvent_basecaller_b200 import _lib
import bench
C = 100000
raw, ev = bench.synth_chunks(np.random.default_rng(0), C)
rd, ed = torch.from_numpy(raw).cuda(), torch.from_numpy(ev).cuda()
bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0.); bc.load_weights(seed=22)
for _ in range(3): bc.beam_search_prediction((rd, ed), 1, 34)
torch.cuda.synchronize()
for rep in range(12):
    _lib.profile(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ids, sc = bc.beam_search_prediction((rd, ed), 1, 34)
    e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    pr = _lib.profile_read(); _lib.profile(False)
    print(f"rep {rep}: wall {1e3*(t2-t0):.1f} ms, events {e0.elapsed_time(e1):.1f} ms, kernel sum {sum(v['ms'] for v in pr.values()):.1f}",
          {k: round(v['ms'], 1) for k, v in pr.items() if v['ms'] > 0})
