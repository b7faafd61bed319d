"""BASELINE-size checks (configs[2]/[3]: 100 000 joint chunks on one B200) through size-independent
properties, since the CPU oracle cannot finish that size in seconds:
  * determinism (two runs bit-identical),
  * permutation equivariance (snippets are independent: shuffling the chunks permutes the outputs),
  * independence from the internal wave partition (a prefix computed alone equals the prefix of the full run),
  * and an oracle spot check on chunks drawn from the far end of the batch."""
import numpy as np
import pytest
import torch

from oracle import model_ref as mr
from oracle.parity import check_beam

pytestmark = pytest.mark.gpu

N = 100_000
L = 34


@pytest.fixture(scope="module")
def setup():
    import ravvent_basecaller_b200 as rb
    w = mr.init_weights(22)
    bc = rb.Basecaller(128, 128, 128, rb.nuc_tk, "joint", 0.)
    bc.load_weights(w)
    rng = np.random.default_rng(2026)
    raw = torch.from_numpy(rng.standard_normal((N, 200, 1), dtype=np.float32))
    ev = torch.from_numpy(rng.standard_normal((N, 30, 5), dtype=np.float32))
    rl = torch.from_numpy(rng.integers(159, 196, size=N)); el = torch.from_numpy(rng.integers(16, 28, size=N))
    raw[torch.arange(200)[None, :] >= rl[:, None]] = 0.0
    ev[torch.arange(30)[None, :] >= el[:, None]] = 0.0
    return bc, w, raw.cuda(), ev.cuda()


@pytest.mark.parametrize("beam", [1, 5])
def test_fullsize_properties(setup, beam):
    bc, w, raw, ev = setup
    ids, sc = bc.beam_search_prediction((raw, ev), beam, L)
    assert ids.shape == (N, L - 1) or ids.shape[0] == N
    ids2, sc2 = bc.beam_search_prediction((raw, ev), beam, L)
    assert torch.equal(ids, ids2) and torch.equal(sc, sc2)                      # deterministic
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(1)).cuda()
    idp, scp = bc.beam_search_prediction((raw[perm], ev[perm]), beam, L)
    T = min(ids.shape[1], idp.shape[1])
    assert torch.equal(idp[:, :T], ids[perm][:, :T]) and torch.equal(scp[:, :T], sc[perm][:, :T])   # equivariant
    k = 9472 + 123                                                              # not a multiple of any tile size
    idk, sck = bc.beam_search_prediction((raw[:k], ev[:k]), beam, L)
    T = min(ids.shape[1], idk.shape[1])
    assert torch.equal(idk[:, :T], ids[:k, :T]) and torch.equal(sck[:, :T], sc[:k, :T])             # wave independent
    # oracle spot check on the last 48 chunks of the batch: every beam slot, near ties explained (oracle/parity.py)
    sel = slice(N - 48, N)
    enc, mask = mr.encode_input(w, (raw[sel].cpu().numpy(), ev[sel].cpu().numpy()), "joint")
    got = bc.beam_search_prediction((raw[sel], ev[sel]), beam, L, return_all_beams=True)
    check_beam([g.cpu().numpy() for g in got], w, enc, mask, beam, L, label=f" fullsize tail W={beam}")
    T = min(ids.shape[1], got[0].shape[1])
    assert torch.equal(got[0][:, :T, 0], ids[sel][:, :T]) and torch.equal(got[1][:, :T, 0], sc[sel][:, :T])   # same as inside the big batch


def test_fullsize_encoder_checksum(setup):
    """Encoder outputs of the full batch: finite, bounded by 1 in magnitude (h = o*tanh(c)), and a
    checksum that is invariant under the wave partition."""
    bc, w, raw, ev = setup
    n = 30000
    enc, mask = bc._encode_input((raw[:n], ev[:n]))
    assert torch.isfinite(enc).all() and enc.abs().max() <= 1.0
    enc2, _ = bc._encode_input((raw[5000:n], ev[5000:n]))
    assert torch.equal(enc[5000:], enc2)
    assert torch.equal(mask[:, :200], (raw[:n, :, 0] != 0)) and torch.equal(mask[:, 200:], (ev[:n] != 0).all(dim=-1))
