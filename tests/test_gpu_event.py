"""K1 parity: the CUDA event scan (through the EventDetector host class -> C ABI) against the
golden vectors made by the reference's EventDetector and against the oracle on fresh signals.
start / length / mean are bit-exact; stdv is compared through the variance with a bound of
4 ulp(mean^2) because the reference squares the mean with libm pow."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import event_ref
from test_oracle_event import CASES, c_detect, load_c_oracle

pytestmark = pytest.mark.gpu


def _check(ev, start, length, mean, stdv):
    st = np.array([e.start for e in ev], dtype=np.int64)
    ln = np.array([e.length for e in ev], dtype=np.int64)
    mu = np.array([e.mean for e in ev]); sd = np.array([e.stdv for e in ev])
    assert np.array_equal(st, start)
    assert np.array_equal(ln, length)
    assert np.array_equal(mu, mean)                         # bitwise float64
    bound = 4 * np.spacing(mean * mean) + 1e-300
    assert np.all(np.abs(sd * sd - stdv * stdv) <= bound + 8 * np.spacing(stdv * stdv))
    assert np.mean(sd == stdv) > 0.99 or len(sd) < 100      # and bit-identical almost everywhere


@pytest.mark.parametrize("warmup", [-1, 0, 16, 256])
@pytest.mark.parametrize("case", CASES, ids=[f"case{c[0]}" for c in CASES])
def test_matches_reference_golden(case, warmup):
    """warmup = 0 makes every speculative sub-segment start from the wrong state, so the whole
    read goes through the verify-and-rerun path; the answer must not change."""
    import ravvent_basecaller_b200 as rb
    _, raw, par, start, length, mean, stdv = case
    ev = rb.EventDetector(*par, warmup=warmup).run(raw)
    _check(ev, start, length, mean, stdv)


def test_int16_and_int32_inputs_agree():
    import ravvent_basecaller_b200 as rb
    _, raw, par, start, length, mean, stdv = CASES[0]
    det = rb.EventDetector(*par)
    _check(det.run(raw.astype(np.int16)), start, length, mean, stdv)
    _check(det.run(raw.astype(np.int64)), start, length, mean, stdv)


def test_ragged_batch_against_c_oracle():
    """Several reads (incl. empty and shorter-than-window ones) in one launch; chunk chains of
    different reads must not interact."""
    import ravvent_basecaller_b200 as rb
    rng = np.random.default_rng(99)
    lens = [5000, 0, 17, 2048, 2049, 4096, 1, 30000, 6143, 19]
    reads = [event_ref.synth_read(rng, n) if n else np.zeros(0, np.int32) for n in lens]
    offs = np.concatenate(([0], np.cumsum(lens)))
    out = rb.EventDetector(6, 9).detect_batch(np.concatenate(reads), offs)
    lib = load_c_oracle()
    cnt = out["count"].cpu().numpy()
    for r, raw in enumerate(reads):
        st, ln, mu, sd = c_detect(lib, raw, 6, 9, 1.4, 9.0, 0.2)
        o = int(out["event_offsets"][r])
        assert cnt[r] == st.size, (r, cnt[r], st.size)
        assert np.array_equal(out["start"][o:o + cnt[r]].cpu().numpy(), st)
        assert np.array_equal(out["length"][o:o + cnt[r]].cpu().numpy(), ln)
        assert np.array_equal(out["mean"][o:o + cnt[r]].cpu().numpy(), mu)


def test_full_size_properties():
    """BASELINE-size input (1M chunks ~ 57M samples is scaled to 8M samples here to keep the CPU
    checker in seconds): events tile every read, and a checksum of (start,length,mean) equals the
    C oracle's."""
    import ravvent_basecaller_b200 as rb
    rng = np.random.default_rng(5)
    n_reads, n = 128, 60000
    sig = np.concatenate([event_ref.synth_read(rng, n) for _ in range(n_reads)])
    offs = np.arange(n_reads + 1) * n
    out = rb.EventDetector(6, 9).detect_batch(sig, offs)
    cnt = out["count"].cpu().numpy()
    start = out["start"].cpu().numpy(); length = out["length"].cpu().numpy(); mean = out["mean"].cpu().numpy()
    lib = load_c_oracle()
    for r in range(n_reads):
        o = int(out["event_offsets"][r]); c = int(cnt[r])
        st, ln = start[o:o + c], length[o:o + c]
        assert st[0] == 0 and np.array_equal(st[1:], st[:-1] + ln[:-1]) and (ln > 0).all()
        if r % 16 == 0:
            cst, cln, cmu, _ = c_detect(lib, sig[offs[r]:offs[r + 1]], 6, 9, 1.4, 9.0, 0.2)
            assert np.array_equal(st, cst) and np.array_equal(ln, cln) and np.array_equal(mean[o:o + c], cmu)
