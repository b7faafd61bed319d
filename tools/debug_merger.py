"""Find minimal 2-snippet cases where the GPU merger and oracle/merger_ref.py differ (debug aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import merger_ref as m
import ravvent_basecaller_b200 as rb
sys.path.insert(0, "tests")
from test_gpu_merger import _pack, _rand_seq

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
reads = []
for _ in range(3000):
    a = _rand_seq(rng, int(rng.integers(1, 30)))
    k = int(rng.integers(0, len(a)))
    b = list(a[k:] + _rand_seq(rng, int(rng.integers(0, 12))))
    for _e in range(int(rng.integers(0, 4))):
        if b:
            j = int(rng.integers(0, len(b)))
            op = rng.integers(0, 3)
            if op == 0: b[j] = rng.choice(list("ACGT"))
            elif op == 1: b.insert(j, rng.choice(list("ACGT")))
            else: del b[j]
    b = "".join(b)[:33]
    reads.append([(a, [0.5] * len(a)), (b, [0.7] * len(b))])
ids, probs, off = _pack(reads)
for ss in (0, 1, 2):
    got = rb.Merger(ss).merge_predictions(ids, None, off, probs=probs)
    bad = [i for i, (r, g) in enumerate(zip(reads, got)) if g.seq != m.merge_read(r, ss)[0]]
    print("score set", ss, "bad", len(bad), "of", len(reads))
    bad.sort(key=lambda i: len(reads[i][0][0]) + len(reads[i][1][0]))
    for i in bad[:4]:
        a, b = reads[i][0][0], reads[i][1][0]
        al = m.local_align(a[-25:], b[:25], ss)
        print(" a", a, "b", b)
        print("   oracle align", al[0][:3] if al else None)
        print("   oracle", m.merge_read(reads[i], ss)[0], " gpu", got[i].seq, [round(x, 1) for x in got[i].logits])
