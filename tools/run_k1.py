import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravvent_basecaller_b200 as rb
from oracle.event_ref import synth_read
rng = np.random.default_rng(3)
reads = [synth_read(rng, 60000) for _ in range(8)]
n_reads = 256
sig = torch.from_numpy(np.concatenate([reads[i % 8] for i in range(n_reads)])).cuda()
offs = np.arange(n_reads + 1, dtype=np.int64) * 60000
det = rb.EventDetector(6, 9)
for _ in range(3):
    out = det.detect_batch(sig, offs)
torch.cuda.synchronize()
print(int(out["count"].sum()))
