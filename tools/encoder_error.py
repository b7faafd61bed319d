"""Max / rms difference of the GPU encoder outputs and beam-1 logits-free scores against the fp32 oracle (sanity figure for DESIGN.md)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import model_ref as mr
import ravvent_basecaller_b200 as rb
w = mr.init_weights(22, random_bias=True)
x = mr.synth_chunks(np.random.default_rng(3), 96)
bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0.).load_weights(w)
enc, mask = bc._encode_input(x)
renc, rmask = mr.encode_input(w, x, "joint")
d = (enc - renc)[rmask]
print("encoder outputs: max abs err %.3g, rms %.3g, max |ref| %.3g" % (np.abs(d).max(), np.sqrt((d ** 2).mean()), np.abs(renc).max()))
ids, sc = bc.beam_search_prediction(x, 5, 16)
rid, rsc = mr.beam_search(w, renc, rmask, 5, 16)
same = np.array([np.array_equal(a, b) for a, b in zip(ids, rid)])
print("beam 5: %d / %d rows identical, max score err %.3g" % (same.sum(), len(same), np.abs(sc[same] - rsc[same]).max()))
