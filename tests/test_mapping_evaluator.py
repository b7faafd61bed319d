"""File writers / PAF reader of the mapping evaluator (ravvent_mapping_evaluator.py:74-110) and create_files_info
(data_loader.py:129-156) -- SURVEY §8 f-3."""
import json

import numpy as np
import pytest

from ravvent_basecaller_b200.evaluator import RavventMappingEvaluator as RME


def test_fasta_fastq_writers(tmp_path):
    seq = "ACGTACGTACGTTT"
    RME._create_fasta(seq, tmp_path / "r.fasta")
    RME._create_fastq(seq, tmp_path / "p.fastq")
    assert (tmp_path / "r.fasta").read_text() == ">ACGTACGTAC\nACGTACGTACGTTT"
    assert (tmp_path / "p.fastq").read_text() == "@ACGTACGTAC\nACGTACGTACGTTT\n+\n" + "!" * len(seq)


def test_paf_identity(tmp_path):
    paf = tmp_path / "m.paf"
    paf.write_text("q\t1000\t0\t600\t+\tt\t2000\t10\t610\t540\t600\t60\tNM:i:60\n"
                   "short line\n"
                   "q\t1000\t600\t1000\t+\tt\t2000\t700\t1100\t380\t400\t60\n")
    got = RME._read_mapping_identity(paf)
    assert got == {"read_length": 1000, "matches": 920, "total_block_len": 1000, "identity": 0.92}
    paf.write_text("")
    assert RME._read_mapping_identity(paf)["identity"] == 0.0


def test_minimap_missing_is_reported(tmp_path):
    import shutil
    if shutil.which("minimap2"):
        pytest.skip("minimap2 installed")
    with pytest.raises(FileNotFoundError, match="minimap2"):
        RME._run_minimap(tmp_path / "a", tmp_path / "b", tmp_path / "c")


@pytest.mark.gpu
def test_create_files_info(tmp_path):
    from oracle.event_ref import synth_read
    from ravvent_basecaller_b200 import data_loader as dl
    rng = np.random.default_rng(3)
    for i, n in enumerate((2500, 3100)):
        raw = synth_read(rng, n)
        np.savetxt(tmp_path / f"read{i}.signal", raw.reshape(1, -1), fmt="%d")
        edges = np.arange(0, n + 1, 10); edges[-1] = n
        with open(tmp_path / f"read{i}.label", "w") as f:
            for a, b in zip(edges[:-1], edges[1:]):
                f.write(f"{a} {b} A\n")
    path = dl.create_files_info(tmp_path, stride=6, verbose=False)
    info = json.loads(path.read_text())
    assert path.name == "files_info.snippets.stride_6.json" and len(info) == 2
    for i, rec in enumerate(info):
        r, _, _ = dl.load_data_from_single_signal_label(rec["signal_path"], rec["label_path"], 6)
        assert rec["snippets_num"] == r.shape[0] > 0 and rec["signal_path"].endswith(f"read{i}.signal")
