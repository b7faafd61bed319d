"""K2-K5 parity: CUDA path (through the Basecaller host class -> C ABI) against the CPU oracle
on the same seeded weights and inputs.

Tolerances (north_star): encoder outputs and logits agree with the fp32 oracle within
rtol 1e-3 (+ atol 2e-5 for values near zero); integer search outputs are identical except for
near-tie argmax flips, which are reported with their logit margins and bounded."""
import numpy as np
import pytest
import torch

from oracle import model_ref as mr
from oracle.parity import check_beam

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 2e-5
W22 = mr.init_weights(22, random_bias=True)


def make(kind, depth=2, wave=0, weights=None, dec_depth=1):
    import ravvent_basecaller_b200 as rb
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, kind, 0., encoder_depth=depth, decoder_depth=dec_depth, wave_snippets=wave)
    bc.compile(optimizer=None)
    bc.load_weights(weights if weights is not None else W22)
    return bc


def inputs(kind, n, seed=0):
    raw, ev = mr.synth_chunks(np.random.default_rng(seed), n)
    return {"raw": raw, "event": ev, "joint": (raw, ev)}[kind]


@pytest.mark.parametrize("kind,n", [("raw", 70), ("event", 130), ("joint", 64), ("joint", 3)])
def test_encode_input_matches_oracle(kind, n):
    x = inputs(kind, n)
    enc, mask = make(kind)._encode_input(x)
    ref, rmask = mr.encode_input(W22, x, kind)
    assert enc.shape == ref.shape and enc.dtype == np.float32
    assert np.array_equal(mask, rmask)
    np.testing.assert_allclose(enc, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("depth", [1, 3])
def test_encoder_depths(depth):
    w = mr.init_weights(7, encoder_depth=depth, random_bias=True)
    x = inputs("joint", 20, seed=3)
    enc, _ = make("joint", depth=depth, weights=w)._encode_input(x)
    ref, _ = mr.encode_input(w, x, "joint", encoder_depth=depth)
    np.testing.assert_allclose(enc, ref, rtol=RTOL, atol=ATOL)


def test_multi_wave_and_device_tensors():
    """B larger than the internal wave; torch device inputs give torch device outputs."""
    x = inputs("joint", 200, seed=4)
    bc = make("joint", wave=64)
    enc, mask = bc._encode_input((torch.from_numpy(x[0]).cuda(), torch.from_numpy(x[1]).cuda()))
    assert enc.is_cuda and mask.dtype == torch.bool
    ref, rmask = mr.encode_input(W22, x, "joint")
    np.testing.assert_allclose(enc.cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
    assert np.array_equal(mask.cpu().numpy(), rmask)


def test_short_time_axis():
    """Time lengths shorter than 200/30 (caller padded differently)."""
    raw, ev = inputs("joint", 10, seed=5)
    x = (raw[:, :57], ev[:, :11])
    enc, mask = make("joint")._encode_input(x)
    ref, rmask = mr.encode_input(W22, x, "joint")
    np.testing.assert_allclose(enc, ref, rtol=RTOL, atol=ATOL)


def _first_diff(a, b):
    d = np.flatnonzero(a != b)
    return int(d[0]) if d.size else None


def _check_greedy(ids, logits, rid, rlog):
    """Greedy rows are identical to the oracle's except for near-tie argmax flips: at the first differing step the
    oracle's top-2 logit margin is < 1e-3; logits agree up to and including that step."""
    n = len(rid)
    assert ids.shape == rid.shape and ids.dtype == np.int32 and logits.shape == rlog.shape
    flips = 0
    for b in range(n):
        t = _first_diff(ids[b], rid[b])
        upto = ids.shape[1] if t is None else t + 1
        np.testing.assert_allclose(logits[b, :upto], rlog[b, :upto], rtol=RTOL, atol=1e-4)
        if t is not None:
            top2 = np.sort(rlog[b, t])[-2:]
            assert top2[1] - top2[0] < 1e-3, f"row {b} step {t}: flip with margin {top2[1] - top2[0]}"
            flips += 1
    assert flips <= max(1, n // 50), f"{flips} near-tie flips in {n} rows"


@pytest.mark.parametrize("kind", ["raw", "joint"])
def test_greedy_matches_oracle(kind):
    n, L = 96, 20
    x = inputs(kind, n, seed=6)
    ids, logits = make(kind).greedy_search_prediction(x, L)
    enc, mask = mr.encode_input(W22, x, kind)
    rid, rlog = mr.greedy_search(W22, enc, mask, L)
    _check_greedy(ids, logits, rid, rlog)


@pytest.mark.parametrize("kind,W", [("joint", 1), ("joint", 5), ("event", 5), ("raw", 3)])
def test_beam_matches_oracle(kind, W):
    """Every row identical to the oracle in every beam slot (ids, parents, gather_tree output, scores), except rows
    that diverge first at a step where the oracle's own top-(W+1) margin is < 1e-3 (oracle/parity.py)."""
    n, L = 70, 16
    x = inputs(kind, n, seed=8)
    bc = make(kind)
    got = bc.beam_search_prediction(x, W, L, return_all_beams=True)
    enc, mask = mr.encode_input(W22, x, kind)
    ties = check_beam(got, W22, enc, mask, W, L, atol=1e-4, label=f" {kind} W={W}")
    ids, sc = bc.beam_search_prediction(x, W, L)                 # the reference's return value: beam slot 0
    assert np.array_equal(ids, got[0][:, :, 0]) and np.array_equal(sc, got[1][:, :, 0])
    assert ids.dtype == np.int32 and sc.dtype == np.float32
    assert ties == 0 or ties <= 2


@pytest.mark.parametrize("W", [1, 5, 9])
def test_beam_early_termination_matches_oracle(W):
    """A large end-token bias makes every beam finish within a few steps: the kernel's early exit must
    reproduce what tfa's dynamic_decode would have produced (T, ids, scores), also when other
    snippets of the batch finish later."""
    w = dict(W22)
    b = w["decoder/fc/bias"].copy(); b[mr.TOKEN_END] += 2.5
    w["decoder/fc/bias"] = b
    n, L = 50, 20
    x = inputs("joint", n, seed=11)
    got = make("joint", weights=w).beam_search_prediction(x, W, L, return_all_beams=True)
    enc, mask = mr.encode_input(w, x, "joint")
    rid, _ = mr.beam_search(w, enc, mask, W, L)
    assert rid.shape[1] < L - 1, "the bias should end decoding early"
    check_beam(got, w, enc, mask, W, L, label=f" early-stop W={W}")
    # a milder bias: some snippets finish (and are skipped by the attention kernel from then on) while others decode on
    b2 = W22["decoder/fc/bias"].copy(); b2[mr.TOKEN_END] += 1.0
    w["decoder/fc/bias"] = b2
    got = make("joint", weights=w).beam_search_prediction(x, W, 34, return_all_beams=True)
    check_beam(got, w, enc, mask, W, 34, label=f" mixed early-stop W={W}")


@pytest.mark.parametrize("enc_depth,W", [(3, 1), (3, 5), (2, 5)])
def test_decoder_depth_2(enc_depth, W):
    """(encoder_depth, decoder_depth) = (3,2) and (2,2): the other configurations of the shipped results
    (accuracy_results_all.lambda.beam5.json:89,118): two stacked decoder LSTM cells."""
    w = mr.init_weights(31, encoder_depth=enc_depth, decoder_depth=2, random_bias=True)
    n, L = 40, 14
    x = inputs("joint", n, seed=12)
    bc = make("joint", depth=enc_depth, weights=w, dec_depth=2)
    enc, mask = mr.encode_input(w, x, "joint", encoder_depth=enc_depth)
    got = bc.beam_search_prediction(x, W, L, return_all_beams=True)
    check_beam(got, w, enc, mask, W, L, decoder_depth=2, label=f" depth ({enc_depth},2) W={W}")
    if W == 1:
        gid, glog = bc.greedy_search_prediction(x, L)
        rgid, rglog = mr.greedy_search(w, enc, mask, L, decoder_depth=2)
        _check_greedy(gid, glog, rgid, rglog)


def test_beam_all_beams_internal_consistency():
    """predicted_ids must equal gather_tree(step_ids, parent_ids) recomputed by the oracle, exactly."""
    x = inputs("joint", 33, seed=9)
    bc = make("joint")
    pred, sc, step_ids, parents = bc.beam_search_prediction(x, 5, 14, return_all_beams=True)
    lengths = np.zeros((33, 5), dtype=np.int64)
    fin = np.zeros((33, 5), dtype=bool); fin[:, 1:] = True
    for t in range(step_ids.shape[1]):
        pf = np.take_along_axis(fin, parents[:, t].astype(np.int64), axis=1)
        lengths = np.take_along_axis(lengths, parents[:, t].astype(np.int64), axis=1) + (~pf)
        fin = pf | (step_ids[:, t] == mr.TOKEN_END)
    ref = mr.gather_tree(np.transpose(step_ids, (1, 0, 2)), np.transpose(parents, (1, 0, 2)), lengths.max(axis=1))
    assert np.array_equal(pred, np.transpose(ref, (1, 0, 2)))
    assert np.all(np.diff(sc, axis=2) <= 0)          # top_k returns scores sorted descending


def test_beam_step_kernel_bit_exact():
    """K5 standalone on identical log-probs, with many exact ties: integer outputs and scores bit-equal."""
    import ctypes as C
    from ravvent_basecaller_b200 import _lib
    rng = np.random.default_rng(3)
    for W in (1, 3, 5, 9):
        B, V = 257, 7
        slp = mr.log_softmax(rng.integers(-2, 3, size=(B, W, V)).astype(np.float32))
        lp = np.sort(rng.normal(-3, 1, size=(B, W)).astype(np.float32))[:, ::-1].copy()
        if W > 2:
            lp[:, 2] = lp[:, 1]
        lp[::7, -1] = -np.inf
        fin = rng.random((B, W)) < 0.3
        lens = rng.integers(0, 9, size=(B, W)).astype(np.int64)
        ref = mr.beam_step(slp, lp, fin, lens)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        t_slp, t_lp, t_fin, t_len = d(slp), d(lp), d(fin.astype(np.uint8)), d(lens)
        o_sc = torch.empty((B, W), dtype=torch.float32, device="cuda"); o_w = torch.empty((B, W), dtype=torch.int32, device="cuda")
        o_p = torch.empty_like(o_w); o_f = torch.empty((B, W), dtype=torch.uint8, device="cuda"); o_l = torch.empty((B, W), dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib.rvb_beam_step(t_slp.data_ptr(), t_lp.data_ptr(), t_fin.data_ptr(), t_len.data_ptr(), B, W, V, 1,
                                          o_sc.data_ptr(), o_w.data_ptr(), o_p.data_ptr(), o_f.data_ptr(), o_l.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(o_w.cpu().numpy(), ref[1]) and np.array_equal(o_p.cpu().numpy(), ref[2])
        assert np.array_equal(o_sc.cpu().numpy(), ref[0])
        assert np.array_equal(o_f.cpu().numpy().astype(bool), ref[4]) and np.array_equal(o_l.cpu().numpy(), ref[5])


def test_gather_tree_kernel_bit_exact():
    from ravvent_basecaller_b200 import _lib
    rng = np.random.default_rng(4)
    T, B, W = 33, 300, 5
    ids = rng.integers(1, 7, size=(T, B, W)).astype(np.int32)
    par = rng.integers(0, W, size=(T, B, W)).astype(np.int32)
    mx = rng.integers(0, T + 3, size=B).astype(np.int32)
    ref = mr.gather_tree(ids, par, mx)
    t_ids, t_par, t_mx = (torch.from_numpy(a).cuda() for a in (ids, par, mx))     # keep the tensors alive
    out = torch.empty((T, B, W), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.rvb_gather_tree(t_ids.data_ptr(), t_par.data_ptr(), t_mx.data_ptr(), T, B, W, 1, out.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("prec", [-1, 0, 1, 2, 3])
def test_projection_kernel(prec):
    """K2 through rvb_project.  prec -1: FFMA validation kernel; 0: tcgen05 3xTF32 (fp32 parity);
    1: tcgen05 single tf32 pass; 2: fp16 hi/lo planes, 3 passes on the fp16 pipe (what the encoders use
    between layers, fp32 parity); 3: fp16 planes, single pass."""
    from ravvent_basecaller_b200 import _lib
    rng = np.random.default_rng(5)
    shapes = [(1000, 1024, 256), (257, 128, 256), (64, 128, 384), (128 * 300 + 5, 1024, 256), (300, 512, 128)]
    if prec >= 2:
        shapes = [sh for sh in shapes if sh[1] % 256 == 0 and sh[2] % 64 == 0]
    for M, N, K in shapes:
        a = rng.normal(size=(M, K)).astype(np.float32); b = (rng.normal(size=(K, N)) * 0.1).astype(np.float32)
        bias = rng.normal(size=N).astype(np.float32)
        ta, tb, tbias = (torch.from_numpy(v).cuda() for v in (a, b, bias))
        c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.rvb_project(ta.data_ptr(), tb.data_ptr(), tbias.data_ptr(), c.data_ptr(), M, N, K, prec, None))
        torch.cuda.synchronize()
        ref = a.astype(np.float64) @ b.astype(np.float64) + bias
        got = c.cpu().numpy()
        err = np.abs(got - ref).max()
        # |a||b| row/col norms ~ 16 * 1.6: fp32-level error ~1e-5, tf32 single pass ~1e-2
        assert np.isfinite(got).all(), (M, N, K, np.isnan(got).mean())
        assert err < (5e-5 if prec in (-1, 0, 2) else 3e-2), (M, N, K, err)
        c2 = torch.empty((M, N), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.rvb_project(ta.data_ptr(), tb.data_ptr(), None, c2.data_ptr(), M, N, K, prec, None))
        torch.cuda.synchronize()
        np.testing.assert_allclose(c2.cpu().numpy() + bias, got, rtol=0, atol=1e-5)


def test_error_paths():
    import ravvent_basecaller_b200 as rb
    with pytest.raises(NotImplementedError):
        rb.Basecaller(128, 128, 128, rb.nuc_tk, 'joint', 0., rnn_type='birnn')
    with pytest.raises(rb.RavventError):
        rb.Basecaller(64, 64, 128, rb.nuc_tk, 'joint', 0.)
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, 'raw', 0.)
    with pytest.raises(rb.RavventError):                      # weights not loaded
        bc._encode_input(inputs("raw", 2))
    bc.load_weights(seed=22)
    with pytest.raises(ValueError):
        bc._encode_input(np.zeros((2, 200, 3), np.float32))
    ids, sc = bc.beam_search_prediction(np.zeros((0, 200, 1), np.float32), 5, 10)
    assert ids.shape[0] == 0


@pytest.mark.parametrize("pair", [1, 2])
def test_projection_cluster_variants(pair):
    """The cluster-of-two forms of the fp16-plane projection (RVB_GEMM_PAIR=1: W halves TMA-multicast, 2: one
    cta_group::2 MMA per pair) are opt-in A/B variants; the mode is read once per process, hence the subprocess."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import numpy as np, torch
        from ravvent_basecaller_b200 import _lib
        rng = np.random.default_rng(5)
        for M in (128 * 300 + 5, 128 * 296):
            a = rng.normal(size=(M, 256)).astype(np.float32); b = (rng.normal(size=(256, 1024)) * 0.1).astype(np.float32)
            bias = rng.normal(size=1024).astype(np.float32)
            ta, tb, tbias = (torch.from_numpy(v).cuda() for v in (a, b, bias))
            c = torch.full((M, 1024), float("nan"), dtype=torch.float32, device="cuda")
            _lib.check(_lib.lib.rvb_project(ta.data_ptr(), tb.data_ptr(), tbias.data_ptr(), c.data_ptr(), M, 1024, 256, 2, None))
            torch.cuda.synchronize()
            err = np.abs(c.cpu().numpy() - (a.astype(np.float64) @ b.astype(np.float64) + bias)).max()
            assert err < 5e-5, err
        print("ok")
    """)
    env = dict(os.environ, RVB_GEMM_PAIR=str(pair))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("W,n", [(2, 5), (4, 1), (8, 6), (9, 7)])
def test_beam_width_and_batch_edges(W, n):
    """Every attention-kernel instantiation of the wave-level decoder (widths 1, <= 5, <= 9) and batches that do not
    fill a CTA; interior masked rows (an exactly-zero sample inside the valid part of a chunk) are honoured."""
    L = 12
    raw, ev = mr.synth_chunks(np.random.default_rng(21 + W), n)
    raw[0, 17, 0] = 0.0                                   # interior row masked out (utils.input_mask: any zero feature)
    ev[n - 1, 3, :] = 0.0
    got = make("joint").beam_search_prediction((raw, ev), W, L, return_all_beams=True)
    enc, mask = mr.encode_input(W22, (raw, ev), "joint")
    assert not mask[0, 17] and not mask[n - 1, 200 + 3]
    check_beam(got, W22, enc, mask, W, L, atol=1e-4, label=f" edges W={W} n={n}")
