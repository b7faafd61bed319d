// Internal launch interfaces shared between the kernel translation units and api.cu.
#pragma once
#include "common.cuh"

namespace rvb {

namespace rectc {   // K3 on tcgen05 (cta_group::2), lstm_recurrent_tc.cu
struct Params {
    const float *x;             // [B,T,F] batch-major (layer 0)
    const float *G;             // pre-gates (layer > 0): element (b,t,dir*512 + unit*4 + gate) at G[b*g_bs + t*g_ts + ...]
    long long g_bs, g_ts;
    int g_blocked;              // 1: G is [row tile of 128][column quad][128 rows][4] with row = t * g_rows_per_t + b (coalesced by lane)
    long long g_rows_per_t;     // padded batch (multiple of 128) of the blocked layout
    int g16;                    // 1 (reduced-precision mode only, blocked): the quads are 4 fp16 values (8 bytes), as K2 writes them there
    const uint16_t *bimg;       // [dir][rank] pre-swizzled fp16 hi/lo B-operand images (pack_b_image)
    const float *w0;            // [dir][6][512] layer-0 input rows + bias row, [unit][gate] column order
    const float *state_in;      // [B,2,2,128] or nullptr
    float *state_out;
    float *y;                   // (b,t,dir*128+u) at y[b*y_bs + t*y_ts + dir*128 + u]   (nullptr when planes are written)
    long long y_bs, y_ts;
    uint16_t *y16_hi, *y16_lo;  // optional fp16 hi / lo planes of the same logical tensor (input of the next K2)
    long long y16_bs, y16_ts;
    uint16_t *yv16;             // optional fp16 copy of y with the same (y_bs, y_ts) element strides (decoder memory, reduced-precision mode)
    int B, T;
    int precision;              // RVB_PREC_FP32: 3 split passes, RVB_PREC_BF16: single 16-bit pass
    int *abort_flag;
    int gru;                    // 1: Keras GRUCell on the same four columns per unit (z, r, candidate input part, candidate recurrent part)
};
int run(int feat, const Params &p, cudaStream_t stream);
void pack_b_image(const float *U, int rank, uint16_t *img);
constexpr int B_IMAGE_BYTES = 8 * 16384;
constexpr int W0_FLOATS_PER_DIR = 6 * 512;
}  // namespace rectc

namespace gemm {    // K2, proj_gemm.cu
int run_simt(const float *A, const float *Bm, const float *bias, float *C, long long M, int N, int K, cudaStream_t s);
// tcgen05 path: WhiT / WloT are the [N,K] tf32 hi / remainder parts made by prepare_weights
int prepare_weights(const float *W, float *hiT, float *loT, int K, int N, cudaStream_t s);
// lda: row pitch of A in elements (0 = K)
// Optional fused epilogue of the wave-level decoder's cell GEMM (N = 512 gate columns in [unit][gate] order): instead of
// storing Z, each accumulator row adds the token's kernel row, runs the LSTM cell update against c_in[parent row] and
// writes c_out[row, unit] and h into xa[row, 0:128] (row stride 384).
struct CellEpilogue {
    const float *wtok;          // [vocab][512]: input-kernel row of the token + bias, [unit][gate] order
    const int32_t *tok, *parent;    // tok == nullptr: row 0 of wtok for every row (a plain bias, second stacked cell)
    const float *c_in;
    float *c_out, *xa;          // xa: fp32 h at xa[row*xa_ld + unit] (must be non-null: it selects this epilogue)
    int W;                      // beams per snippet (parent indices are relative to the snippet's first row)
    uint16_t *h_hi, *h_lo;      // optional fp16 hi / lo planes of h at [row*h_ld + unit]: A operand of the next GEMM
    int xa_ld, h_ld;            // row pitches in elements (0 = the depth-1 defaults 384 / 128)
    int gru;                    // 1: GRU cell (columns z, r, candidate input part, candidate recurrent part); c_in / c_out carry h
};
int run_tc(const float *A, const float *WhiT, const float *WloT, const float *bias, float *C, long long M, int N, int K,
           int precision, int *abort_flag, cudaStream_t s, long long lda = 0, const CellEpilogue *cell = nullptr);
bool tc_available();
// fp16-plane path (A produced as hi/lo planes by K3): all operands fp16, 3 passes on the fp16 pipe
int split_planes_f16(const float *X, void *hi, void *lo, long long n, cudaStream_t s);
int prepare_weights_f16(const float *W, void *hiT, void *loT, int K, int N, cudaStream_t s);
int run_tc_f16(const void *Ahi, const void *Alo, const void *WhiT, const void *WloT, const float *bias, float *C, long long M,
               int N, int K, int precision, int *abort_flag, cudaStream_t s, bool blocked_out = false, const CellEpilogue *cell = nullptr,
               long long lda = 0, int n_out = 0, bool blocked_half = false);
// lda: row pitch of the A planes (0 = K); n_out: columns of C (0 = N; W padded beyond it); blocked_half: the blocked output holds
// fp16 quads (8 bytes per row and column quad) instead of fp32 ones
}  // namespace gemm

namespace decw {    // K4 + K5 per decode step over the whole wave (beam width >= 2), decoder_wave.cu
struct Params {
    const float *values;        // [B,Tm,256]
    const uint16_t *values16;   // optional fp16 copy of values (reduced-precision mode): the attention kernel streams it instead
    int att16_tc;               // values16 goes through the single-plane tcgen05 attention (else the FFMA kernel)
    const uint16_t *v_hi, *v_lo;    // optional fp16 hi / lo planes of values, [B,Tm,256] each: beam widths >= 2 run the tcgen05 attention
    const uint8_t *mask;        // [B,Tm]
    const float *wg_hiT, *wg_loT;   // tf32 hi / lo of [att-input rows ; recurrent kernel], transposed [512,256], [unit][gate] columns
    const float *wm_hiT, *wm_loT;   // W_mem^T as a [K=128, N=256] weight, transposed [256,128]
    const float *wa_hiT, *wa_loT;   // attention layer [384,128], transposed [128,384]
    const void *wg16_hi, *wg16_lo, *wm16_hi, *wm16_lo;   // fp16 hi / lo planes of the first two (transposed); nullptr: tf32 path
    const void *wa16_hi, *wa16_lo;  // fp16 hi / lo planes of the attention layer, zero-padded to 256 output columns, transposed [256,384];
                                    // nullptr: the attention-layer GEMM takes fp32 [h | ctx] rows on the tf32 path
    const float *wtok;          // [7][512] kernel row of token v + bias, [unit][gate] columns
    const void *wg1_16_hi, *wg1_16_lo;   // decoder_depth 2: fp16 hi / lo planes of [kernel ; recurrent kernel] of cell 1, transposed [512,256]
    const float *b1;            // decoder_depth 2: bias of cell 1, [512] in [unit][gate] order
    int depth;                  // stacked decoder cells: 1 or 2
    int gru;                    // 1: GRU cells (the weights carry the four-column form: z, r, candidate input part, candidate recurrent part)
    int greedy;                 // 1: greedy search (W == 1): ids = sample_id [B,S], logits [B,S,7]; no beam bookkeeping
    float *logits;              // greedy only
    const float *wfc, *bfc;     // [128][7], [7]
    int B, Tm, W, S;
    int32_t *ids;               // predicted_ids [B,S,W]
    float *scores;              // [B,S,W]
    int32_t *step_ids, *parent_ids;   // [B,S,W]
    int32_t *steps;             // atomicMax of T
    float *ws;                  // workspace_floats(B*W) floats
    int *abort_flag;
};
size_t workspace_floats(long long rows, int depth = 1);
int run(const Params &p, cudaStream_t stream);
}  // namespace decw

namespace atc {     // K4 attention on tcgen05, attention_tc.cu: fp16 hi + lo planes (parity mode, widths >= 2), or v_lo == nullptr:
                    // one fp16 plane (reduced-precision mode, every width)
int run(const uint16_t *v_hi, const uint16_t *v_lo, const uint8_t *mask, const float *Q, float *xa, const int32_t *skip,
        int B, int Tm, int W, int *abort_flag, cudaStream_t s, uint16_t *xp_hi = nullptr, uint16_t *xp_lo = nullptr);
// xp_hi / xp_lo != nullptr: the context goes out as fp16 hi / lo planes at [row*384 + 128 + column] INSTEAD of fp32 into xa
}  // namespace atc

}  // namespace rvb
