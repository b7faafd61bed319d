"""Device time of rvb_merge_reads alone (CUDA events), error-free stride-5 windows: R reads x n snippets."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from ravvent_basecaller_b200 import _lib
for R, n_snip in ((1, 1000), (24, 1000), (100, 1000), (1000, 1000)):
    rng = np.random.default_rng(8)
    W, stride = 30, 5
    L = (n_snip - 1) * stride + W
    reads = rng.integers(0, 4, (R, L)).astype(np.int32)
    idx = (np.arange(n_snip)[:, None] * stride + np.arange(W)[None, :])
    ids = np.ones((R * n_snip, 33), np.int32)
    ids[:, :W] = (reads[:, idx] + 3).reshape(R * n_snip, W)
    d_ids = torch.from_numpy(ids).cuda()
    d_probs = torch.full((R * n_snip, 33), 0.9, dtype=torch.float32, device="cuda")
    off = torch.from_numpy((np.arange(R + 1) * n_snip).astype(np.int32)).cuda()
    seq = torch.zeros(R * n_snip * 33, dtype=torch.uint8, device="cuda")
    pl = torch.zeros(R * n_snip * 33, dtype=torch.float32, device="cuda")
    ln = torch.zeros(R, dtype=torch.int32, device="cuda")
    def run():
        _lib.check(_lib.lib.rvb_merge_reads(d_ids.data_ptr(), d_probs.data_ptr(), R * n_snip, 33, off.data_ptr(), R, 0,
                                            seq.data_ptr(), pl.data_ptr(), ln.data_ptr(), None))
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"{R:5d} reads x {n_snip} snippets: {ms:8.2f} ms  = {1e3 * ms / n_snip:6.2f} us per snippet of a read chain")
