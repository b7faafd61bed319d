/* ravvent_b200.h -- C ABI of libravvent_b200.so (B200 / sm_100a).
 *
 * The reference (adamnapieralski/ravvent-basecaller) is pure Python and has no
 * FFI; its drop-in boundary is the Python class surface of `EventDetector` and
 * `Basecaller`.  This header is what the Python host layer
 * (ravvent_basecaller_b200/*.py, same class / method names as the reference)
 * binds with ctypes; INTEGRATION.md shows the stub a reference maintainer
 * would add.  Each entry point cites the reference code it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns an int status (RVB_OK == 0); on failure a
 *     thread-local message is available from rvb_last_error().  Nothing throws
 *     across the boundary.
 *   - `d_` pointers are device memory on the handle's device, `h_` pointers are
 *     host memory.  Outputs are caller-allocated.  The library never returns
 *     memory the caller must free; a model handle owns its weights/workspace.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls
 *     are asynchronous with respect to the host unless stated otherwise.
 *   - one handle per GPU; calls on one handle must be serialised by the caller
 *     (the reference `Basecaller` is not re-entrant either, basecaller.py:303,326).
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with RVB_ERR_CUDA.
 */
#ifndef RAVVENT_B200_H
#define RAVVENT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVB_OK             0
#define RVB_ERR_ARG        1   /* bad argument / unsupported configuration      */
#define RVB_ERR_CUDA       2   /* CUDA runtime / driver error                    */
#define RVB_ERR_STATE      3   /* call order (e.g. weights not finalised)        */
#define RVB_ERR_OVERFLOW   4   /* caller-provided capacity too small             */
#define RVB_ERR_INTERNAL   5   /* device-side consistency check failed           */

#define RVB_INPUT_RAW      0
#define RVB_INPUT_EVENT    1
#define RVB_INPUT_JOINT    2

#define RVB_PREC_FP32      0   /* parity mode: split-precision (3-pass) tensor-core products, fp32-accurate     */
#define RVB_CELL_LSTM      0   /* rvb_model_set_rnn cell_kind */
#define RVB_CELL_GRU       1
#define RVB_PREC_BF16      1   /* reduced mode: single 16-bit pass, fp32 accumulate / cell state, fp16 pre-gates and attention memory */

typedef struct rvb_model rvb_model_t;

int         rvb_version(void);
const char *rvb_last_error(void);
/* Number of visible CUDA devices (0 and RVB_OK when there is no driver/GPU). */
int         rvb_device_count(int *count);

/* ------------------------------------------------------------------------
 * K1  event detection  -- replaces EventDetector.run / _add_sample /
 *     _compute_tstat / _detect_peak / _create_event
 *     (event_detection/event_detector.py:75-210) for a ragged batch of reads.
 *
 * d_signal        concatenated integer samples of all reads (int32, or int16
 *                 when sample_bytes == 2)
 * h_read_offsets  n_reads+1 sample offsets into d_signal (host)
 * h_event_offsets n_reads+1 offsets into the event arrays: read r may produce
 *                 at most h_event_offsets[r+1]-h_event_offsets[r] events
 *                 (len/2+2 is always enough); RVB_ERR_OVERFLOW otherwise
 * d_ev_*          structure-of-arrays event table (start, length: int32 holding
 *                 the reference's u32 values; mean, stdv: float64)
 * d_ev_count      n_reads event counts
 * warmup          speculation warm-up in samples (<= 256); < 0 selects the default (32)
 *                 when negative.  Results never depend on it (exact verify).
 * Bit-exact vs the reference for start/length/mean; stdv differs only where
 * libm pow(mean,2) != mean*mean.
 * ------------------------------------------------------------------------ */
int rvb_event_detect_workspace_bytes(const int64_t *h_read_offsets, int32_t n_reads, size_t *bytes);
int rvb_event_detect(const void *d_signal, int sample_bytes,
                     const int64_t *h_read_offsets, int32_t n_reads,
                     int window_length1, int window_length2,
                     double threshold1, double threshold2, double peak_height,
                     const int64_t *h_event_offsets,
                     int32_t *d_ev_start, int32_t *d_ev_length,
                     double *d_ev_mean, double *d_ev_stdv, int32_t *d_ev_count,
                     void *d_workspace, size_t workspace_bytes, int warmup, void *stream);

/* ------------------------------------------------------------------------
 * Snippet builder -- replaces data_loader.prepare_snippets /
 * compute_fitting_event_ranges / convert_events_ranges_to_raw_ranges /
 * pad_input_snippets (data_loader.py:29-51, 70-111) for ONE read whose events
 * are already on the device (output of rvb_event_detect).
 *   d_raw_snips   [max_snippets,200,1] f32, d_event_snips [max_snippets,30,5] f32
 *   h_n_snippets  number of snippets produced (host, written after a stream sync)
 *   d_raw_ranges  optional [max_snippets,2] int32: sample range [start,end) of each snippet in the read
 *                 (what the reference uses to cut the label sequence, data_loader.py:101-102); may be NULL
 * ------------------------------------------------------------------------ */
int rvb_build_snippets(const void *d_signal, int sample_bytes, int64_t n_samples,
                       const int32_t *d_ev_start, const int32_t *d_ev_length,
                       const double *d_ev_mean, const double *d_ev_stdv, int32_t n_events,
                       int64_t label_start, int64_t label_end, int32_t stride,
                       float *d_raw_snips, float *d_event_snips, int32_t max_snippets,
                       int32_t *h_n_snippets, int32_t *d_raw_ranges, void *stream);

/* The same for a batch of whole reads (label range = the read, the inference case; the reference loops over reads on the
 * host, ravvent_performance_evaluator.py:60-66): the arguments are what rvb_event_detect took and left --
 *   d_signal            all reads concatenated, read r at samples [h_read_offsets[r], h_read_offsets[r+1])
 *   h_event_offsets     [n_reads+1] capacity offsets of the event arrays, d_counts [n_reads] events detected per read
 * Snippets of read r follow those of read r-1:
 *   d_snippet_offsets   optional [n_reads+1] int64 (device): first snippet of each read, total last
 *   d_raw_snips         [max_snippets,200,1] f32 or NULL (event-only model), d_event_snips [max_snippets,30,5] f32
 *   d_raw_ranges        optional [max_snippets,2] int32, relative to the snippet's own read; may be NULL
 *   h_n_snippets        total number of snippets (host): ONE device-to-host copy and stream sync for the whole batch
 * RVB_ERR_OVERFLOW (nothing beyond the capacity is written) when the batch yields more than max_snippets. */
int rvb_build_snippets_batch(const void *d_signal, int sample_bytes, const int64_t *h_read_offsets, int32_t n_reads,
                             const int64_t *h_event_offsets, const int32_t *d_ev_start, const int32_t *d_ev_length,
                             const double *d_ev_mean, const double *d_ev_stdv, const int32_t *d_counts,
                             int32_t stride, float *d_raw_snips, float *d_event_snips, int64_t max_snippets,
                             int64_t *d_snippet_offsets, int32_t *d_raw_ranges, int64_t *h_n_snippets, void *stream);

/* ------------------------------------------------------------------------
 * Model handle -- replaces Basecaller.__init__ / load_weights
 * (basecaller.py:158-206; ravvent_performance_evaluator.py:89-107).
 * Supported: enc_units == dec_units == 128, encoder_depth 1..3,
 * decoder_depth 1..2, bilstm encoders, Luong attention, vocab 7.
 * ------------------------------------------------------------------------ */
int rvb_model_create(rvb_model_t **out, int device, int enc_units, int dec_units,
                     int encoder_depth, int decoder_depth, int vocab_size,
                     int input_kind, int precision, int wave_snippets);
/* rnn_type of the reference constructor (basecaller.py:25-46, 86-89, 195): 'bi*' -> bidirectional = 1; '*lstm' / '*gru' ->
 * cell_kind.  The default after rvb_model_create is ('bilstm': 1, RVB_CELL_LSTM).  Call before rvb_model_finalize.
 * GRU weights: kernel [in,3u], recurrent_kernel [u,3u], bias [2,3u] (Keras reset_after = True); unidirectional
 * encoders have no ".../backward/..." weights and layers > 0 / the decoder's memory and attention layers take
 * enc_units instead of 2*enc_units inputs. */
int rvb_model_set_rnn(rvb_model_t *m, int bidirectional, int cell_kind);
int rvb_model_destroy(rvb_model_t *m);
/* name as in the .npz interchange (oracle/model_ref.py init_weights), e.g.
 * "encoder_raw/layer0/forward/kernel"; row-major float32 host data. */
int rvb_model_set_weight(rvb_model_t *m, const char *name, const float *h_data,
                         const int64_t *shape, int ndim);
int rvb_model_finalize(rvb_model_t *m);
/* Synchronises the device and reports device-side failures of earlier asynchronous calls
 * (e.g. a tensor-core pipeline that timed out).  RVB_OK when everything completed. */
int rvb_model_check(rvb_model_t *m);

/* K2+K3  Basecaller._encode_input (basecaller.py:395-416; Encoder.call :48-59).
 * d_raw [B,t_raw,1] / d_event [B,t_event,5] float32 (either may be NULL per
 * input_kind).  d_enc_out [B,Tm,256] f32, d_mask [B,Tm] u8, Tm = t_raw+t_event. */
int rvb_encode(rvb_model_t *m, const float *d_raw, int t_raw, const float *d_event, int t_event,
               int64_t batch, float *d_enc_out, uint8_t *d_mask, void *stream);

/* K2+K3+K4  Basecaller.greedy_search_prediction (basecaller.py:317-330).
 * Runs S = max_output_len-1 steps; d_ids [B,S] i32, d_logits [B,S,V] f32.
 * *d_steps receives T (number of steps tfa's dynamic_decode would have run);
 * the reference result is the [:, :T] prefix. */
int rvb_greedy(rvb_model_t *m, const float *d_raw, int t_raw, const float *d_event, int t_event,
               int64_t batch, int max_output_len, int32_t *d_ids, float *d_logits,
               int32_t *d_steps, void *stream);

/* K2+K3+K4+K5  Basecaller.beam_search_prediction (basecaller.py:296-315).
 * d_pred_ids / d_scores [B,S,W] (gather_tree'd ids, per-step beam scores); the
 * reference returns [:, :T, 0].  Optional d_step_ids / d_parent_ids [B,S,W]
 * (raw BeamSearchDecoderOutput) may be NULL. */
int rvb_beam(rvb_model_t *m, const float *d_raw, int t_raw, const float *d_event, int t_event,
             int64_t batch, int beam_width, int max_output_len,
             int32_t *d_pred_ids, float *d_scores, int32_t *d_step_ids, int32_t *d_parent_ids,
             int32_t *d_steps, void *stream);

/* Host-buffer variant (the reference-facing call: numpy in, numpy out).  The batch is processed in waves through
 * two device I/O sets and three streams: while wave k computes, wave k+1 is staged and copied in and wave k-1 is
 * copied out.  Pageable caller memory goes through the handle's pinned staging buffers (chunked memcpy on the
 * calling thread); page-locked caller memory is copied from / to directly.  Returns after the last wave has
 * landed in h_ids [B,S] / h_scores [B,S] (beam slot 0); *h_steps = T. */
int rvb_beam_host(rvb_model_t *m, const float *h_raw, int t_raw, const float *h_event, int t_event,
                  int64_t batch, int beam_width, int max_output_len,
                  int32_t *h_ids, float *h_scores, int32_t *h_steps);

/* K5 standalone (parity tests): one tfa _beam_search_step on log-softmaxed rows
 * and tfa gather_tree.  All device pointers. */
int rvb_beam_step(const float *d_step_log_probs, const float *d_log_probs, const uint8_t *d_finished,
                  const int64_t *d_lengths, int64_t batch, int beam_width, int vocab, int end_token,
                  float *d_scores, int32_t *d_word, int32_t *d_parent, uint8_t *d_next_finished,
                  int64_t *d_next_lengths, void *stream);
int rvb_gather_tree(const int32_t *d_step_ids, const int32_t *d_parent_ids, const int32_t *d_max_len,
                    int steps, int64_t batch, int beam_width, int end_token,
                    int32_t *d_out, void *stream);

/* K2 standalone (parity / roofline tests): C[M,N] = A[M,K] * B[K,N] (+ bias[N])
 * through the same projection kernel the encoders use. */
int rvb_project(const float *d_a, const float *d_b, const float *d_bias, float *d_c,
                int64_t m, int n, int k, int precision, void *stream);

/* ---- snippet -> read stitching ---------------------------------------------
 * rvb_beam_scores_to_probs: utils.calc_prob_logits_beam_search_scores (utils.py:123-128) on the
 *   beam-0 scores of rvb_beam: probs[i,t] = exp(scores[i,t] - scores[i,t-1]), scores[i,-1] = 0.
 * rvb_merge_reads: ravvent_performance_evaluator.py:66-74 -- tokens_to_nuc_sequences + SeqLogitsPair
 *   per snippet, then Merger(scores_id).merge per read (merger.py:121-248, which calls Biopython
 *   pairwise2.align.localms / localds on the 25-base overlaps).  Snippets of read r are rows
 *   [read_offsets[r], read_offsets[r+1]) of d_ids / d_probs ([n_snippets, steps], device).  The merged
 *   read r is written at d_seq_out / d_prob_out + read_offsets[r] * steps (buffers of n_snippets * steps
 *   elements; bases as codes 0..3 = A C G T) with its length in d_len_out[r].  Synchronises the stream. */
int rvb_beam_scores_to_probs(const float *d_scores, int64_t n_snippets, int steps, float *d_probs, void *stream);
int rvb_merge_reads(const int32_t *d_ids, const float *d_probs, int64_t n_snippets, int steps,
                    const int32_t *d_read_offsets, int n_reads, int scores_id,
                    uint8_t *d_seq_out, float *d_prob_out, int32_t *d_len_out, void *stream);

/* Introspection for bench.py: kernels launched by this library since load, and optional
 * per-kernel device timing (CUDA events on the launching stream).  rvb_profile(1) starts a
 * fresh recording, rvb_profile(0) stops; rvb_profile_read fills ms[6] / launches[6] (n >= 6) in the order
 * event scan, projection GEMM, recurrent LSTM, decoder (dense phases + search), other, decoder attention
 * (synchronises the device). */
int64_t rvb_launch_count(void);
int rvb_profile(int enable);
int rvb_profile_read(double *ms, int64_t *launches, int n);

#ifdef __cplusplus
}
#endif
#endif /* RAVVENT_B200_H */
