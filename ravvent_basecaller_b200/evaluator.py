"""RavventPerformanceEvaluator -- the reference's timed read loop (ravvent_performance_evaluator.py:13-148)
over the B200 path: `.signal` / `.label` in, one merged read and the timing dictionary out -- and
RavventMappingEvaluator (ravvent_mapping_evaluator.py:20-110): the same read written as FASTA / FASTQ and
mapped with minimap2 when that binary is installed.

Same method names, arguments and result keys as the reference, so its result files and
`compute_total_results` keep working.  Differences that do not change results: snippets stay on the
device between the loader, the basecaller and the merger; `beam_width` is a constructor argument
(the reference hard-codes 5 at :55 and edits the file for beam 1).
"""
from __future__ import annotations

import json
import shlex
import shutil
import subprocess
from pathlib import Path
from timeit import default_timer as timer

import numpy as np
import torch

from . import data_loader as dl
from .basecaller import Basecaller
from .merger import Merger


class RavventPerformanceEvaluator():
    def __init__(self, merger_scores_id=0, beam_width=5, device=None, precision='fp32', encoder_depth=2, decoder_depth=1):
        """encoder_depth / decoder_depth: what the reference hard-codes in setup_basecaller (2, 1 at
        ravvent_performance_evaluator.py:99-100) or edits as module constants (ravvent_mapping_evaluator.py:16-17)."""
        self.merger = Merger(scores_id=merger_scores_id, device=device)
        self.stride = 6
        self.basecaller = None
        self.beam_width = int(beam_width)
        self.device = device
        self.precision = precision
        self.encoder_depth, self.decoder_depth = int(encoder_depth), int(decoder_depth)

    def _split_into_chunks(self, arr, def_chunk_size):
        """Chunks of def_chunk_size rows, the last one shorter (ravvent_performance_evaluator.py:19-22)."""
        n = int(arr.shape[0])
        return [arr[i:i + def_chunk_size] for i in range(0, max(n, 1), def_chunk_size)]

    def run(self, signal_data_source, chunk_size=1024):
        """One read.  -> dict with the reference's keys (:77-87) plus 'merged_seq'."""
        label_path = Path(signal_data_source).with_suffix('.label')
        labels = np.loadtxt(label_path, dtype='object', ndmin=2)
        ranges_ids = labels[:, :2].astype(int)
        ref_seq = ''.join(list(labels[:, 2]))
        samples_num = int(ranges_ids[-1, 1] - ranges_ids[0, 0])

        def sync():
            torch.cuda.synchronize(self.basecaller.device)

        start = timer()
        raw_snippets, event_snippets, nuc_tk_snippets = dl.load_data_from_single_signal_label(
            signal_data_source, label_path, self.stride, as_numpy=False, device=self.basecaller.device.index)
        data_chunks = list(zip(self._split_into_chunks(raw_snippets, chunk_size),
                               self._split_into_chunks(event_snippets, chunk_size),
                               self._split_into_chunks(nuc_tk_snippets, chunk_size)))
        sync()
        t_data_loading = timer() - start

        max_len = int(nuc_tk_snippets.shape[1])
        tokens, scores = [], []
        t_predicting = 0.0
        for data in data_chunks:
            start = timer()
            input_data, target_data = dl.unpack_data_to_input_target(data, self.basecaller.input_data_type)
            pred_tokens, beam_scores = self.basecaller.beam_search_prediction(
                input_data, beam_width=self.beam_width, max_output_len=target_data.shape[1])
            sync()
            t_predicting += timer() - start
            tokens.append(pred_tokens)
            scores.append(beam_scores)

        # post-processing (:66-70: scores -> probabilities, tokens -> bases) and merge (:74) are one device call here
        start = timer()
        S = max(max_len - 1, 1)
        n = sum(int(t.shape[0]) for t in tokens)
        ids = torch.full((n, S), int(self.basecaller.output_end_token), dtype=torch.int32, device=self.basecaller.device)
        sc = torch.zeros((n, S), dtype=torch.float32, device=self.basecaller.device)
        row = 0
        for t, s in zip(tokens, scores):          # dynamic_decode may stop early: pad with end tokens / flat scores
            k = int(t.shape[1])
            ids[row:row + t.shape[0], :k] = t
            sc[row:row + t.shape[0], :k] = s
            if k and k < S:
                sc[row:row + t.shape[0], k:] = s[:, -1:]
            row += int(t.shape[0])
        merged = self.merger.merge_predictions(ids, sc, [0, n])[0] if n else None
        t_merge = timer() - start
        t_postprocessing = 0.0

        return {
            'bases_num': len(ref_seq),
            'samples_num': samples_num,
            't_data_loading': t_data_loading,
            't_predicting': t_predicting,
            't_postprocessing': t_postprocessing,
            't_merge': t_merge,
            'total': t_data_loading + t_predicting + t_postprocessing + t_merge,
            'total_processing': t_predicting + t_postprocessing + t_merge,
            'merged_seq': merged.seq if merged is not None else '',
        }

    def run_batch(self, signal_data_sources, chunk_size=1 << 20):
        """Several reads per device call (not in the reference, which loops over run()): the snippets of all reads are
        decoded together and stitched by one merge call, one warp per read.  Same result keys, summed over the reads,
        plus 'merged_seqs'."""
        def sync():
            torch.cuda.synchronize(self.basecaller.device)

        start = timer()
        raws, events, offsets, bases_num, samples_num, max_len = [], [], [0], 0, 0, 2
        for src in signal_data_sources:
            label_path = Path(src).with_suffix('.label')
            labels = np.loadtxt(label_path, dtype='object', ndmin=2)
            ranges_ids = labels[:, :2].astype(int)
            bases_num += int(labels.shape[0])
            samples_num += int(ranges_ids[-1, 1] - ranges_ids[0, 0])
            r, e, tk = dl.load_data_from_single_signal_label(src, label_path, self.stride, as_numpy=False,
                                                             device=self.basecaller.device.index)
            raws.append(r)
            events.append(e)
            offsets.append(offsets[-1] + int(r.shape[0]))
            max_len = max(max_len, int(tk.shape[1]))
        raw_all, ev_all = torch.cat(raws), torch.cat(events)
        sync()
        t_data_loading = timer() - start

        start = timer()
        S = max_len - 1
        n = offsets[-1]
        ids = torch.full((n, S), int(self.basecaller.output_end_token), dtype=torch.int32, device=self.basecaller.device)
        sc = torch.zeros((n, S), dtype=torch.float32, device=self.basecaller.device)
        for a in range(0, n, chunk_size):
            data = (raw_all[a:a + chunk_size], ev_all[a:a + chunk_size], None)
            input_data, _ = dl.unpack_data_to_input_target(data, self.basecaller.input_data_type)
            t, s = self.basecaller.beam_search_prediction(input_data, beam_width=self.beam_width, max_output_len=max_len)
            k = int(t.shape[1])
            ids[a:a + t.shape[0], :k] = t
            sc[a:a + t.shape[0], :k] = s
            if k and k < S:
                sc[a:a + t.shape[0], k:] = s[:, -1:]
        sync()
        t_predicting = timer() - start

        start = timer()
        merged = self.merger.merge_predictions(ids, sc, offsets) if n else []
        t_merge = timer() - start
        return {
            'bases_num': bases_num, 'samples_num': samples_num, 't_data_loading': t_data_loading,
            't_predicting': t_predicting, 't_postprocessing': 0.0, 't_merge': t_merge,
            'total': t_data_loading + t_predicting + t_merge, 'total_processing': t_predicting + t_merge,
            'merged_seqs': [m.seq for m in merged],
        }

    def setup_basecaller(self, weights_path, data_type, mode=1):
        """ravvent_performance_evaluator.py:89-107; weights_path may also be None (seeded random initialisation)."""
        if mode == 1:
            self.basecaller = Basecaller(
                enc_units=128, dec_units=128, batch_sz=128, tokenizer=dl.nuc_tk, input_data_type=data_type,
                input_padding_value=dl.INPUT_PADDING, encoder_depth=self.encoder_depth, decoder_depth=self.decoder_depth, rnn_type='bilstm',
                attention_type='luong', teacher_forcing=0.5, device=self.device, precision=self.precision)
        self.basecaller.compile(optimizer=None)
        self.basecaller.load_weights(weights_path)

    def compute_total_results(self, results_path):
        """:109-129 (the reference returns the running-mean speeds; its unreachable second return is dropped)."""
        with open(results_path, 'rt') as f:
            results = json.load(f)
        bases_num, samples_num, t_processing = 0, 0, 0
        bases_speeds, signals_speeds = [], []
        for res in results:
            bases_num += res['bases_num']
            samples_num += res['samples_num']
            t_processing += res['total_processing']
            bases_speeds.append(bases_num / t_processing)
            signals_speeds.append(samples_num / t_processing)
        return np.mean(bases_speeds), np.std(signals_speeds), np.mean(signals_speeds), np.std(signals_speeds)

    def evaluate_specific(self, files_info_path, results_path, weights_path, data_type):
        """:131-148: every `signal_path` of a files_info JSON, results appended to results_path after each read."""
        results = []
        self.setup_basecaller(weights_path, data_type, mode=1)
        with open(files_info_path, 'rt') as f:
            val_files = [v['signal_path'] for v in json.load(f)]
        for v in val_files:
            res = self.run(v)
            res['path'] = v
            results.append(res)
            with open(results_path, 'wt') as f:
                json.dump(results, f, indent=2)
        return results


class RavventMappingEvaluator(RavventPerformanceEvaluator):
    """ravvent_mapping_evaluator.py:20-110.  run() basecalls and stitches one read, writes the reference as FASTA and
    the prediction as FASTQ, maps them with `minimap2 -x map-ont -c` and returns the identity dictionary.  The file
    writers and the PAF reader are static so they can be used (and tested) without a GPU."""

    def __init__(self, merger_scores_id=0, beam_width=5, device=None, precision='fp32', work_dir='temp',
                 encoder_depth=1, decoder_depth=1):
        # defaults = the reference's module constants ENCODER_DEPTH = DECODER_DEPTH = 1, BEAM_WIDTH = 5 (:15-17)
        super().__init__(merger_scores_id, beam_width, device, precision, encoder_depth, decoder_depth)
        self.work_dir = Path(work_dir)

    @staticmethod
    def _create_fasta(seq, fname):
        with open(fname, 'wt') as f:
            f.write(f'>{seq[:10]}\n{seq}')

    @staticmethod
    def _create_fastq(seq, fname):
        with open(fname, 'wt') as f:
            f.write(f'@{seq[:10]}\n')
            f.write(seq + '\n')
            f.write('+\n')
            f.write('!' * len(seq))

    @staticmethod
    def _run_minimap(ref_path, pred_path, out_path):
        if shutil.which('minimap2') is None:
            raise FileNotFoundError("minimap2 is not installed (ravvent_mapping_evaluator.py:86 shells out to it)")
        cmd = f'minimap2 -x map-ont -c {ref_path} {pred_path}'
        with open(out_path, 'wt') as f:
            subprocess.run(shlex.split(cmd), stdout=f, check=False)

    @staticmethod
    def _read_mapping_identity(mapping_path):
        matches = total_blocks_len = read_length = 0
        with open(mapping_path, 'rt') as paf:
            for line in paf:
                parts = line.strip().split('\t')
                if len(parts) < 11:
                    continue
                read_length = int(parts[1])
                matches += int(parts[9])
                total_blocks_len += int(parts[10])
        return {'read_length': read_length, 'matches': matches, 'total_block_len': total_blocks_len,
                'identity': matches / total_blocks_len if total_blocks_len != 0 else 0.}

    def run(self, signal_data_source, chunk_size=1024):
        label_path = Path(signal_data_source).with_suffix('.label')
        ref_seq = ''.join(list(np.loadtxt(label_path, dtype='object', ndmin=2)[:, 2]))
        merged_seq = super().run(signal_data_source, chunk_size)['merged_seq']
        self.work_dir.mkdir(parents=True, exist_ok=True)
        fasta_path, fastq_path, mapping_path = (self.work_dir / n for n in ('ref.fasta', 'pred.fastq', 'mapping.paf'))
        self._create_fasta(ref_seq, fasta_path)
        self._create_fastq(merged_seq, fastq_path)
        self._run_minimap(fasta_path, fastq_path, mapping_path)
        return self._read_mapping_identity(mapping_path)
