"""Snippet -> read stitching behind the reference's merger.py surface.

`SeqLogitsPair` and `Merger(scores_id).merge(list_of_pairs)` keep the reference's
names and argument meaning (merger.py:7-37, 121-248); the alignment (Biopython
pairwise2.align.localms / localds in the reference) and the merge run in
libravvent_b200's CUDA kernel, one warp per read.  `Merger.merge_predictions`
is the device-resident form of ravvent_performance_evaluator.py:66-74: it takes
the beam-search outputs of a batch of reads and never moves them to the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = {"A": 3, "C": 4, "G": 5, "T": 6}            # data_loader.nuc_tk ids


class SeqLogitsPair(object):
    """merger.py:7-37."""

    @classmethod
    def align_logits(cls, seq_gapped, logits_non_gapped):
        out, index = [], 0
        for c in seq_gapped:
            if c == '-':
                out.append(-1.)
            else:
                out.append(logits_non_gapped[index])
                index += 1
        return out

    @property
    def seq(self):
        return self._seq

    @property
    def logits(self):
        return self._logits

    def __init__(self, seq, logits):
        assert len(seq) == len(logits)
        self._seq = seq
        self._logits = logits


class Merger():
    def __init__(self, scores_id=0, device=None):
        if scores_id not in (0, 1, 2):
            raise ValueError("scores_id must be 0, 1 or 2 (merger.py:124-147)")
        self.scores_id = scores_id
        self.overlap_seq_len = 25
        if _lib.device_count() == 0:
            raise _lib.RavventError(_lib.RVB_ERR_CUDA, "no CUDA device: Merger has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))

    # -- device-resident path -------------------------------------------------
    def merge_predictions(self, pred_tokens, beam_scores, read_offsets, probs=None):
        """pred_tokens [N,S] int32 and beam_scores [N,S] f32 as returned by Basecaller.beam_search_prediction
        for the concatenated snippets of several reads; read_offsets [R+1] delimits the reads.
        `probs` overrides exp(score_t - score_{t-1}) (utils.calc_prob_logits_beam_search_scores).
        -> list of SeqLogitsPair, one per read."""
        dev = self.device
        ids = torch.as_tensor(pred_tokens).to(dev, torch.int32).contiguous()
        N, S = int(ids.shape[0]), int(ids.shape[1]) if ids.dim() == 2 else 0
        off = torch.as_tensor(np.asarray(read_offsets, dtype=np.int32)).to(dev)
        R = int(off.numel()) - 1
        if R < 0 or int(off[0]) != 0 or int(off[-1]) != N:
            raise ValueError("read_offsets must start at 0 and end at the number of snippets")
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if probs is None:
                sc = torch.as_tensor(beam_scores).to(dev, torch.float32).contiguous()
                p = torch.empty_like(sc)
                _lib.check(_lib.lib.rvb_beam_scores_to_probs(sc.data_ptr(), N, S, p.data_ptr(), stream))
            else:
                p = torch.as_tensor(probs).to(dev, torch.float32).contiguous()
            if tuple(p.shape) != (N, S):
                raise ValueError("scores / probs must have the shape of pred_tokens")
            seq = torch.zeros((max(N * S, 1),), dtype=torch.uint8, device=dev)
            plog = torch.zeros((max(N * S, 1),), dtype=torch.float32, device=dev)
            lens = torch.zeros((max(R, 1),), dtype=torch.int32, device=dev)
            _lib.check(_lib.lib.rvb_merge_reads(ids.data_ptr(), p.data_ptr(), N, S, off.data_ptr(), R, self.scores_id,
                                                seq.data_ptr(), plog.data_ptr(), lens.data_ptr(), stream))
        seq_h, log_h, len_h, off_h = seq.cpu().numpy(), plog.cpu().numpy(), lens.cpu().numpy(), off.cpu().numpy()
        out = []
        for r in range(R):
            a, n = int(off_h[r]) * S, int(len_h[r])
            out.append(SeqLogitsPair(_BASES[seq_h[a:a + n]].tobytes().decode("ascii"), log_h[a:a + n].tolist()))
        return out

    # -- the reference's call (merger.py:153) -----------------------------------
    def merge(self, nuc_pred_snippets):
        """One read: list of SeqLogitsPair (base strings over ACGT + per-base probabilities) -> SeqLogitsPair."""
        n = len(nuc_pred_snippets)
        if n == 0:
            raise IndexError("merge of an empty snippet list")        # the reference indexes [0]
        S = max(1, max(len(sp.seq) for sp in nuc_pred_snippets))
        ids = np.zeros((n, S), dtype=np.int32)
        probs = np.zeros((n, S), dtype=np.float32)
        for i, sp in enumerate(nuc_pred_snippets):
            ids[i, :len(sp.seq)] = [_CODE[c] for c in sp.seq.upper()]
            probs[i, :len(sp.seq)] = sp.logits
        return self.merge_predictions(ids, None, [0, n], probs=probs)[0]
