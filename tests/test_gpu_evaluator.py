"""The reference's timed read loop (RavventPerformanceEvaluator.run, ravvent_performance_evaluator.py:24-87) on the
B200 path: .signal/.label -> event detection -> snippets -> beam search -> read stitching."""
import json

import numpy as np
import pytest

from oracle import merger_ref as mref
from oracle.event_ref import synth_read

pytestmark = pytest.mark.gpu


def _write_read(tmp_path, name, n_samples, seed):
    rng = np.random.default_rng(seed)
    raw = synth_read(rng, n_samples)
    edges = np.arange(0, n_samples + 1, 9)
    edges[-1] = n_samples
    syms = rng.choice(list("ACGT"), size=len(edges) - 1)
    sp = tmp_path / f"{name}.signal"
    np.savetxt(sp, raw.reshape(1, -1), fmt="%d")
    with open(tmp_path / f"{name}.label", "w") as f:
        for a, b, c in zip(edges[:-1], edges[1:], syms):
            f.write(f"{a} {b} {c}\n")
    return sp, len(syms)


@pytest.mark.parametrize("beam", [1, 3])
def test_run_matches_manual_pipeline(tmp_path, beam):
    import ravvent_basecaller_b200 as rb
    from ravvent_basecaller_b200 import data_loader as dl
    from ravvent_basecaller_b200.evaluator import RavventPerformanceEvaluator
    sp, n_bases = _write_read(tmp_path, "r0", 4000, 5)
    ev = RavventPerformanceEvaluator(merger_scores_id=0, beam_width=beam)
    ev.setup_basecaller(None, "joint")
    res = ev.run(str(sp), chunk_size=64)
    assert res["bases_num"] == n_bases and res["samples_num"] == 4000
    for k in ("t_data_loading", "t_predicting", "t_postprocessing", "t_merge", "total", "total_processing"):
        assert res[k] >= 0.0
    assert res["total_processing"] == pytest.approx(res["t_predicting"] + res["t_postprocessing"] + res["t_merge"])
    # the same read by hand: loader -> one beam-search call -> the oracle's stitching of those predictions
    raw_s, ev_s, tok = dl.load_data_from_single_signal_label(str(sp), str(sp.with_suffix(".label")), 6)
    assert raw_s.shape[0] > 64                                   # several predict chunks were exercised
    ids, scores = ev.basecaller.beam_search_prediction((raw_s, ev_s), beam, tok.shape[1])
    want = mref.merge_read(mref.snippets_from_predictions(ids, scores))[0]
    assert res["merged_seq"] == want
    assert set(res["merged_seq"]) <= set("ACGT")


def test_evaluate_specific_and_totals(tmp_path):
    from ravvent_basecaller_b200.evaluator import RavventPerformanceEvaluator
    paths = [str(_write_read(tmp_path, f"r{i}", 2500 + 500 * i, 10 + i)[0]) for i in range(2)]
    info = tmp_path / "files_info.json"
    info.write_text(json.dumps([{"signal_path": p} for p in paths]))
    ev = RavventPerformanceEvaluator(beam_width=1)
    out = tmp_path / "results.json"
    results = ev.evaluate_specific(str(info), str(out), None, "raw")
    assert [r["path"] for r in results] == paths
    saved = json.loads(out.read_text())
    assert len(saved) == 2 and saved[0]["bases_num"] == results[0]["bases_num"]
    mean_b, _, mean_s, _ = ev.compute_total_results(str(out))
    assert mean_b > 0 and mean_s > mean_b
    assert ev._split_into_chunks(np.arange(10), 4)[-1].tolist() == [8, 9]


def test_run_batch_equals_per_read_runs(tmp_path):
    from ravvent_basecaller_b200.evaluator import RavventPerformanceEvaluator
    paths = [str(_write_read(tmp_path, f"b{i}", 2000 + 700 * i, 30 + i)[0]) for i in range(3)]
    ev = RavventPerformanceEvaluator(beam_width=3)
    ev.setup_basecaller(None, "joint")
    single = [ev.run(p, chunk_size=4096) for p in paths]
    both = ev.run_batch(paths)
    # every read decodes to the same length here (random weights never emit the end token early), so batching
    # changes nothing but the number of device calls
    assert both["merged_seqs"] == [r["merged_seq"] for r in single]
    assert both["bases_num"] == sum(r["bases_num"] for r in single)
    assert both["samples_num"] == sum(r["samples_num"] for r in single)
