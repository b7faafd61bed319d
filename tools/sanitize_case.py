"""Tiny end-to-end case for compute-sanitizer: K1 (several chunks + ragged reads), snippet builder, K3/K2 (depth 2),
decoder greedy / beam 1 / beam 5, both precision modes."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravvent_basecaller_b200 as rb
from ravvent_basecaller_b200 import data_loader as dl
rng = np.random.default_rng(0)
lvl = np.repeat(rng.uniform(250, 550, 2000), 2 + rng.geometric(1 / 7., 2000))
sig = np.rint(lvl[:9000] + rng.normal(0, 8, 9000)).astype(np.int32)
det = rb.EventDetector(6, 9)
out = det.detect_batch(np.concatenate([sig, sig[:2500], sig[:17]]), [0, 9000, 11500, 11517])
print("events", out["count"].cpu().numpy())
rs, es = dl.load_data_from_signal(sig, stride=6, detector=det)
print("snippets", tuple(rs.shape))
for prec in ("fp32", "bf16"):
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., precision=prec, wave_snippets=64)
    bc.load_weights(seed=22)
    x = (rs[:70], es[:70])
    enc, mask = bc._encode_input(x)
    g, _ = bc.greedy_search_prediction(x, 8)
    b1, _ = bc.beam_search_prediction(x, 1, 8)
    b5, _ = bc.beam_search_prediction(x, 5, 8)
    print(prec, tuple(enc.shape), g.shape, b1.shape, b5.shape)
print("done")
