import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravvent_basecaller_b200 as rb
from oracle import model_ref as mr
w = mr.init_weights(22, random_bias=True)
x = mr.synth_chunks(np.random.default_rng(0), 96)
ref, rmask = mr.encode_input(w, x, "joint")
for prec in ("fp32", "bf16"):
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0., precision=prec); bc.load_weights(w)
    enc, mask = bc._encode_input(x)
    print(prec, "enc max abs err", np.abs(enc - ref).max(), "rms", np.sqrt(np.mean((enc - ref) ** 2)), "mask ok", np.array_equal(mask, rmask))
    gid, glog = bc.greedy_search_prediction(x, 20)
    rid, rlog = mr.greedy_search(w, ref, rmask, 20)
    same = np.array([np.array_equal(a, b) for a, b in zip(gid, rid)])
    first = [int(np.flatnonzero(a != b)[0]) if (a != b).any() else 99 for a, b in zip(gid, rid)]
    print(prec, "greedy rows identical", same.mean(), "logit max err (identical rows)", np.abs(glog[same] - rlog[same]).max() if same.any() else None,
          "first-step logit err", np.abs(glog[:, 0] - rlog[:, 0]).max(), "min first mismatch", min(first))
    for W in (1, 5):
        bid, bsc = bc.beam_search_prediction(x, W, 20)
        r2, s2 = mr.beam_search(w, ref, rmask, W, 20)
        same = np.array([np.array_equal(a, b) for a, b in zip(bid, r2)])
        print(prec, f"beam{W} rows identical", same.mean(), "score max err", np.abs(bsc[same] - s2[same]).max() if same.any() else None)
