import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravvent_basecaller_b200 as rb
from ravvent_basecaller_b200 import _lib
import bench
C = 100000
raw, ev = bench.synth_chunks(np.random.default_rng(0), C)
rd, ed = torch.from_numpy(raw).cuda(), torch.from_numpy(ev).cuda()
bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0.); bc.load_weights(seed=22)
for _ in range(2): bc.beam_search_prediction((rd, ed), 1, 34)
torch.cuda.synchronize()
for rep in range(3):
    _lib.profile(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ids, sc = bc.beam_search_prediction((rd, ed), 1, 34)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    pr = _lib.profile_read(); _lib.profile(False)
    print(f"rep {rep}: host call {1e3*(t1-t0):.1f} ms, wall {1e3*(t2-t0):.1f} ms, device events {e0.elapsed_time(e1):.1f} ms, kernels",
          {k: round(v['ms'], 1) for k, v in pr.items()}, "sum", round(sum(v['ms'] for v in pr.values()), 1))
# same without the Python class: raw C call with preallocated outputs
import ctypes as Cc
S, W = 33, 1
ids = torch.empty((C, S, W), dtype=torch.int32, device="cuda"); scs = torch.empty((C, S, W), dtype=torch.float32, device="cuda")
a = torch.empty_like(ids); b = torch.empty_like(ids); steps = torch.zeros(1, dtype=torch.int32, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(_lib.lib.rvb_beam(bc._h, rd.data_ptr(), 200, ed.data_ptr(), 30, C, 1, 34, ids.data_ptr(), scs.data_ptr(), a.data_ptr(), b.data_ptr(), steps.data_ptr(), torch.cuda.current_stream().cuda_stream))
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"raw C call: launch phase {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
