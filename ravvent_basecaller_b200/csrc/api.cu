// C ABI of libravvent_b200.so: model handle, weight packing, and the drivers that chain
// K2/K3/K4/K5 into Basecaller._encode_input / greedy_search_prediction /
// beam_search_prediction (reference basecaller.py:296-330, 395-416).
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"

#include <mutex>

namespace rvb {
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};

struct ProfEntry { int kind; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;

void prof_record(int kind, cudaStream_t stream, bool begin) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (begin) {
        ProfEntry e{kind, nullptr, nullptr};
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, stream);
        g_prof.push_back(e);
    } else {
        for (auto it = g_prof.rbegin(); it != g_prof.rend(); ++it)
            if (it->kind == kind) { cudaEventRecord(it->b, stream); break; }
    }
}

// utils.input_mask (utils.py:26-32): mask[b,t] = all_f(x[b,t,f] != 0)
__global__ void input_mask_kernel(const float *x, int F, long long B, int T, uint8_t *mask, int Tm, int t_off) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    long long b = i / T; int t = (int)(i % T);
    bool ok = true;
    for (int f = 0; f < F; ++f) ok = ok && (x[i * F + f] != 0.0f);
    mask[b * Tm + t_off + t] = ok ? 1 : 0;
}
}  // namespace rvb

using namespace rvb;

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

struct rvb_model {
    int device = 0, enc_depth = 2, dec_depth = 1, input_kind = RVB_INPUT_JOINT, precision = RVB_PREC_FP32;
    int wave = 9472;
    bool finalized = false;
    std::map<std::string, HostTensor> hw;
    // packed device weights
    float *d_pw[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    float *d_pb[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    float *d_phi[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // tf32 hi part, [N,K]
    float *d_plo[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // remainder, [N,K]
    void *d_phi16[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // fp16 hi part, [N,K]
    void *d_plo16[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    int *d_abort = nullptr;
    uint16_t *d_bimg[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    float *d_w0[2] = {nullptr, nullptr};
    // wave-level beam decoder (decoder_wave.cu): tf32 hi/lo transposed weights + Keras-order token rows + workspace
    float *dw_wg[2] = {nullptr, nullptr}, *dw_wm[2] = {nullptr, nullptr}, *dw_wa[2] = {nullptr, nullptr}, *dw_wtok = nullptr, *dw_ws = nullptr;
    uint16_t *dw_wg16[2] = {nullptr, nullptr}, *dw_wm16[2] = {nullptr, nullptr}, *dw_wg1_16[2] = {nullptr, nullptr};
    uint16_t *dw_wa16[2] = {nullptr, nullptr};   // attention layer zero-padded to 256 output columns, fp16 hi / lo planes, transposed [256,384]
    float *dw_wg1[2] = {nullptr, nullptr}, *dw_b1 = nullptr;
    size_t dw_ws_rows = 0;
    float *d_wfc = nullptr, *d_bfc = nullptr;
    // workspace for one wave
    size_t ws_raw_t = 0, ws_ev_t = 0, ws_tm = 0, ws_sw = 0;
    float *y_raw[2] = {nullptr, nullptr}, *y_ev[2] = {nullptr, nullptr};
    float *G_raw = nullptr, *G_ev = nullptr;
    float *st[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [encoder][ping-pong]
    float *enc_out = nullptr, *keys = nullptr;
    uint16_t *enc_out16 = nullptr;             // fp16 copy of enc_out (reduced-precision mode only)
    uint16_t *enc_hi = nullptr, *enc_lo = nullptr;   // fp16 hi / lo planes of the memory (tcgen05 attention, beam widths >= 2)
    bool att_tc = false;
    bool att16_tc = false;                     // reduced-precision mode: single-plane tcgen05 attention at every width
    bool bidir = true;                         // rnn_type 'bi*': Bidirectional encoders; otherwise the backward direction has zero weights (outputs exactly 0)
    int cell = RVB_CELL_LSTM;                  // LSTM or GRU cells (encoders and decoder)
    uint8_t *mask = nullptr;
    int32_t *step_ids = nullptr, *parent_ids = nullptr;
    // host-buffer variant: double-buffered I/O sets, pinned staging, copy streams (HostPipe, below)
    struct HostPipe *pipe = nullptr;
    cudaStream_t hstream = nullptr;
    std::vector<void *> owned;
    std::vector<void *> wowned;                // packed weight buffers: released and rebuilt by every rvb_model_finalize
    bool in_finalize = false;
};

template <typename T>
static int dmalloc(rvb_model *m, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) return fail(RVB_ERR_CUDA, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
    (m->in_finalize ? m->wowned : m->owned).push_back(q);
    *p = reinterpret_cast<T *>(q);
    return RVB_OK;
}
static void dfree(rvb_model *m, void *p) {
    if (!p) return;
    for (auto &q : m->owned)
        if (q == p) { cudaFree(q); q = nullptr; }
    for (auto &q : m->wowned)
        if (q == p) { cudaFree(q); q = nullptr; }
}

static void host_pipe_destroy(rvb_model *m);

extern "C" int rvb_version(void) { return 100; }
extern "C" const char *rvb_last_error(void) { return err_buf(); }
extern "C" int64_t rvb_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int rvb_device_count(int *count) {
    if (!count) return fail(RVB_ERR_ARG, "null count");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return RVB_OK;
}

extern "C" int rvb_model_create(rvb_model_t **out, int device, int enc_units, int dec_units, int encoder_depth,
                                int decoder_depth, int vocab_size, int input_kind, int precision, int wave_snippets) {
    if (!out) return fail(RVB_ERR_ARG, "null out");
    if (enc_units != UNITS || dec_units != UNITS)
        return fail(RVB_ERR_ARG, "kernels are specialised for enc_units == dec_units == 128 (got %d, %d)", enc_units, dec_units);
    if (encoder_depth < 1 || encoder_depth > 3) return fail(RVB_ERR_ARG, "encoder_depth must be 1..3");
    if (decoder_depth < 1 || decoder_depth > 2) return fail(RVB_ERR_ARG, "decoder_depth must be 1 or 2");
    if (vocab_size != VOCAB) return fail(RVB_ERR_ARG, "vocab_size must be 7");
    if (input_kind < 0 || input_kind > 2) return fail(RVB_ERR_ARG, "bad input_kind");
    if (precision != RVB_PREC_FP32 && precision != RVB_PREC_BF16) return fail(RVB_ERR_ARG, "bad precision");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RVB_ERR_CUDA, "no CUDA device: libravvent_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(RVB_ERR_ARG, "device %d out of range (%d devices)", device, n);
    RVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RVB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(RVB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    rvb_model *m = new rvb_model();
    m->device = device; m->enc_depth = encoder_depth; m->dec_depth = decoder_depth;
    m->input_kind = input_kind; m->precision = precision;
    if (wave_snippets > 0) m->wave = (wave_snippets + 63) / 64 * 64;
    if (!gemm::tc_available()) { delete m; return fail(RVB_ERR_CUDA, "tcgen05 projection kernel unavailable on this device"); }
    const char *av = getenv("RVB_ATT");
    m->att_tc = !(av && strcmp(av, "ffma") == 0);           // A/B switch: FFMA attention at beam widths >= 2
    const char *av16 = getenv("RVB_ATT16");
    m->att16_tc = !(av16 && strcmp(av16, "ffma") == 0);     // A/B switch: FFMA attention over the fp16 memory (reduced precision)
    *out = m;
    return RVB_OK;
}

extern "C" int rvb_model_set_rnn(rvb_model_t *m, int bidirectional, int cell_kind) {
    if (!m) return fail(RVB_ERR_ARG, "null model");
    if (cell_kind != RVB_CELL_LSTM && cell_kind != RVB_CELL_GRU) return fail(RVB_ERR_ARG, "cell_kind must be RVB_CELL_LSTM or RVB_CELL_GRU");
    m->bidir = bidirectional != 0;
    m->cell = cell_kind;
    m->finalized = false;
    return RVB_OK;
}

extern "C" int rvb_model_destroy(rvb_model_t *m) {
    if (!m) return RVB_OK;
    cudaSetDevice(m->device);
    host_pipe_destroy(m);
    for (void *q : m->owned) if (q) cudaFree(q);
    for (void *q : m->wowned) if (q) cudaFree(q);
    if (m->hstream) cudaStreamDestroy(m->hstream);
    delete m;
    return RVB_OK;
}

extern "C" int rvb_model_set_weight(rvb_model_t *m, const char *name, const float *h_data, const int64_t *shape, int ndim) {
    if (!m || !name || !h_data || !shape || ndim < 1 || ndim > 2) return fail(RVB_ERR_ARG, "set_weight: bad argument");
    size_t n = 1;
    HostTensor t;
    for (int i = 0; i < ndim; ++i) { if (shape[i] <= 0) return fail(RVB_ERR_ARG, "set_weight: bad shape"); n *= (size_t)shape[i]; t.shape.push_back(shape[i]); }
    t.data.assign(h_data, h_data + n);
    m->hw[name] = std::move(t);
    m->finalized = false;
    return RVB_OK;
}

static int get_w(rvb_model *m, const std::string &name, int64_t d0, int64_t d1, const HostTensor **out) {
    auto it = m->hw.find(name);
    if (it == m->hw.end()) return fail(RVB_ERR_STATE, "missing weight '%s'", name.c_str());
    const HostTensor &t = it->second;
    bool ok = d1 < 0 ? (t.shape.size() == 1 && t.shape[0] == d0) : (t.shape.size() == 2 && t.shape[0] == d0 && t.shape[1] == d1);
    if (!ok) return fail(RVB_ERR_ARG, "weight '%s' has the wrong shape", name.c_str());
    *out = &t;
    return RVB_OK;
}

// One recurrent cell in the internal "four gate blocks of `UNITS` columns" form every packer below consumes:
//   LSTM: Keras kernel [F,4u] / recurrent_kernel [u,4u] / bias [4u] as they are (blocks i, f, g, o);
//   GRU (Keras reset_after = True: kernel [F,3u], recurrent_kernel [u,3u], bias [2,3u], blocks z, r, h):
//        W4 = [W_z | W_r | W_h | 0],  U4 = [U_z | U_r | 0 | U_h],  b4 = [b_iz + b_rz | b_ir + b_rr | b_ih | b_rh]
//     so that column block 2 is the INPUT part of the candidate and block 3 its RECURRENT part (bias included): the cell
//     epilogues compute z = s(s0), r = s(s1), h~ = tanh(s2 + r s3) on the same summed pre-activations an LSTM uses;
//   a missing direction / zero-padded input rows (unidirectional encoders: F_in < F): zeros -- a cell with all-zero
//     weights and a zero state outputs exactly 0, which is what the absent backward half of the memory must be.
struct Cell4 { std::vector<float> W, U, b; };      // [F][512], [128][512], [512]
static int get_cell4(rvb_model *m, const std::string &base, int F, int F_in, bool required, Cell4 *out) {
    out->W.assign((size_t)F * GATES, 0.0f); out->U.assign((size_t)UNITS * GATES, 0.0f); out->b.assign(GATES, 0.0f);
    if (!required && m->hw.find(base + "kernel") == m->hw.end()) return RVB_OK;
    const int G = m->cell == RVB_CELL_GRU ? 3 * UNITS : GATES;
    const HostTensor *W, *U, *Bv;
    RVB_CHECK(get_w(m, base + "kernel", F_in, G, &W));
    RVB_CHECK(get_w(m, base + "recurrent_kernel", UNITS, G, &U));
    if (m->cell == RVB_CELL_GRU) RVB_CHECK(get_w(m, base + "bias", 2, G, &Bv)); else RVB_CHECK(get_w(m, base + "bias", G, -1, &Bv));
    if (m->cell != RVB_CELL_GRU) {
        for (int k = 0; k < F_in; ++k) std::copy(W->data.begin() + (size_t)k * GATES, W->data.begin() + (size_t)(k + 1) * GATES, out->W.begin() + (size_t)k * GATES);
        out->U = U->data; out->b = Bv->data;
        return RVB_OK;
    }
    for (int u = 0; u < UNITS; ++u) {
        for (int k = 0; k < F_in; ++k)
            for (int g = 0; g < 3; ++g) out->W[(size_t)k * GATES + g * UNITS + u] = W->data[(size_t)k * G + g * UNITS + u];
        for (int k = 0; k < UNITS; ++k) {
            out->U[(size_t)k * GATES + 0 * UNITS + u] = U->data[(size_t)k * G + 0 * UNITS + u];
            out->U[(size_t)k * GATES + 1 * UNITS + u] = U->data[(size_t)k * G + 1 * UNITS + u];
            out->U[(size_t)k * GATES + 3 * UNITS + u] = U->data[(size_t)k * G + 2 * UNITS + u];
        }
        out->b[0 * UNITS + u] = Bv->data[0 * UNITS + u] + Bv->data[G + 0 * UNITS + u];
        out->b[1 * UNITS + u] = Bv->data[1 * UNITS + u] + Bv->data[G + 1 * UNITS + u];
        out->b[2 * UNITS + u] = Bv->data[2 * UNITS + u];
        out->b[3 * UNITS + u] = Bv->data[G + 2 * UNITS + u];
    }
    return RVB_OK;
}

static int upload(rvb_model *m, float **dst, const std::vector<float> &v) {
    RVB_CHECK(dmalloc(m, dst, v.size()));
    RVB_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return RVB_OK;
}

static int finalize_impl(rvb_model *m);

extern "C" int rvb_model_finalize(rvb_model_t *m) {
    if (!m) return fail(RVB_ERR_ARG, "null model");
    RVB_CUDA(cudaSetDevice(m->device));
    // a second load_weights on the same handle: drop the previous packed copies first (nothing may be in flight)
    RVB_CUDA(cudaDeviceSynchronize());
    for (void *q : m->wowned) if (q) cudaFree(q);
    m->wowned.clear();
    m->d_abort = nullptr;
    m->finalized = false;
    m->in_finalize = true;
    const int st = finalize_impl(m);
    m->in_finalize = false;
    return st;
}

static int finalize_impl(rvb_model *m) {
    const char *enc_name[2] = {"encoder_raw", "encoder_event"};
    const int enc_feat[2] = {1, 5};
    const char *dir_name[2] = {"forward", "backward"};
    for (int e = 0; e < 2; ++e) {
        if (e == 0 && m->input_kind == RVB_INPUT_EVENT) continue;
        if (e == 1 && m->input_kind == RVB_INPUT_RAW) continue;
        for (int l = 0; l < m->enc_depth; ++l) {
            const int F = l == 0 ? enc_feat[e] : ENC_OUT;
            std::vector<float> wcat, bcat;
            if (l > 0) { wcat.resize((size_t)ENC_OUT * 2 * GATES); bcat.resize(2 * GATES); }
            std::vector<uint16_t> bimg((size_t)4 * (rectc::B_IMAGE_BYTES / 2));
            std::vector<float> w0(l == 0 ? (size_t)2 * rectc::W0_FLOATS_PER_DIR : 0, 0.0f);
            for (int d = 0; d < 2; ++d) {
                std::string base = std::string(enc_name[e]) + "/layer" + std::to_string(l) + "/" + dir_name[d] + "/";
                // unidirectional encoders: no backward weights, and layers > 0 see enc_units inputs instead of 2 enc_units
                Cell4 c4;
                const bool present = d == 0 || m->bidir;
                if (present) RVB_CHECK(get_cell4(m, base, F, (l > 0 && !m->bidir) ? UNITS : F, true, &c4));
                else RVB_CHECK(get_cell4(m, base, F, F, false, &c4));
                struct { std::vector<float> &data; } Wr{c4.W}, Ur{c4.U}, Br{c4.b};
                auto *W = &Wr, *U = &Ur, *Bv = &Br;
                if (l > 0) {
                    // gate columns unit-major (unit*4 + gate): the recurrence's TMEM columns are laid out that way
                    for (int k = 0; k < ENC_OUT; ++k)
                        for (int n = 0; n < GATES; ++n)
                            wcat[(size_t)k * 2 * GATES + d * GATES + (n % UNITS) * 4 + n / UNITS] = W->data[(size_t)k * GATES + n];
                    for (int n = 0; n < GATES; ++n) bcat[d * GATES + (n % UNITS) * 4 + n / UNITS] = Bv->data[n];
                }
                {
                    for (int r = 0; r < 2; ++r)
                        rectc::pack_b_image(U->data.data(), r, bimg.data() + (size_t)(d * 2 + r) * (rectc::B_IMAGE_BYTES / 2));
                    if (l == 0) {
                        for (int f = 0; f <= F; ++f)
                            for (int n = 0; n < GATES; ++n)
                                w0[(size_t)d * rectc::W0_FLOATS_PER_DIR + (size_t)f * GATES + (n % UNITS) * 4 + n / UNITS] =
                                    f < F ? W->data[(size_t)f * GATES + n] : Bv->data[n];
                    }
                }
            }
            {
                RVB_CHECK(dmalloc(m, &m->d_bimg[e][l], bimg.size()));
                RVB_CUDA(cudaMemcpy(m->d_bimg[e][l], bimg.data(), bimg.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                if (l == 0) RVB_CHECK(upload(m, &m->d_w0[e], w0));
            }
            if (l > 0) {
                RVB_CHECK(upload(m, &m->d_pw[e][l], wcat));
                RVB_CHECK(upload(m, &m->d_pb[e][l], bcat));
                RVB_CHECK(dmalloc(m, &m->d_phi[e][l], wcat.size()));
                RVB_CHECK(dmalloc(m, &m->d_plo[e][l], wcat.size()));
                RVB_CHECK(gemm::prepare_weights(m->d_pw[e][l], m->d_phi[e][l], m->d_plo[e][l], ENC_OUT, 2 * GATES, nullptr));
                {
                    uint16_t *h16 = nullptr, *l16 = nullptr;
                    RVB_CHECK(dmalloc(m, &h16, wcat.size()));
                    RVB_CHECK(dmalloc(m, &l16, wcat.size()));
                    m->d_phi16[e][l] = h16; m->d_plo16[e][l] = l16;
                    RVB_CHECK(gemm::prepare_weights_f16(m->d_pw[e][l], h16, l16, ENC_OUT, 2 * GATES, nullptr));
                }
            }
        }
    }
    {
        const HostTensor *Wm_in, *Wa_in, *Wf, *Bf;
        Cell4 d0;
        RVB_CHECK(get_cell4(m, "decoder/cell0/", VOCAB + UNITS, VOCAB + UNITS, true, &d0));
        struct { std::vector<float> &data; } Wdr{d0.W}, Udr{d0.U}, Bdr{d0.b};
        auto *Wd = &Wdr, *Ud = &Udr, *Bd = &Bdr;
        // unidirectional encoders: the memory is enc_units wide; its absent backward half is zero, so are the rows that read it
        const int mem_w = m->bidir ? ENC_OUT : UNITS;
        RVB_CHECK(get_w(m, "decoder/memory_layer/kernel", mem_w, UNITS, &Wm_in));
        RVB_CHECK(get_w(m, "decoder/attention_layer/kernel", UNITS + mem_w, UNITS, &Wa_in));
        HostTensor Wm_pad, Wa_pad;
        Wm_pad.data.assign((size_t)ENC_OUT * UNITS, 0.0f); Wa_pad.data.assign((size_t)(UNITS + ENC_OUT) * UNITS, 0.0f);
        std::copy(Wm_in->data.begin(), Wm_in->data.end(), Wm_pad.data.begin());
        std::copy(Wa_in->data.begin(), Wa_in->data.end(), Wa_pad.data.begin());
        const HostTensor *Wm = &Wm_pad, *Wa = &Wa_pad;
        RVB_CHECK(get_w(m, "decoder/fc/kernel", UNITS, VOCAB, &Wf));
        RVB_CHECK(get_w(m, "decoder/fc/bias", VOCAB, -1, &Bf));
        {
            // [att-input rows ; U] as a [256,512] weight with gate columns in [unit][gate] order (the cell update is fused
            // into that GEMM's epilogue, which sees 32 consecutive columns = 8 whole units); W_mem^T [128,256], W_att [384,128]
            std::vector<float> wcat((size_t)2 * UNITS * GATES), wtk((size_t)VOCAB * GATES), wmT((size_t)UNITS * ENC_OUT);
            for (int k = 0; k < 2 * UNITS; ++k)
                for (int n = 0; n < GATES; ++n) {
                    const int col = (n % UNITS) * 4 + n / UNITS;
                    wcat[(size_t)k * GATES + col] = k < UNITS ? Wd->data[(size_t)(VOCAB + k) * GATES + n] : Ud->data[(size_t)(k - UNITS) * GATES + n];
                }
            for (int v = 0; v < VOCAB; ++v)
                for (int n = 0; n < GATES; ++n) wtk[(size_t)v * GATES + (n % UNITS) * 4 + n / UNITS] = Wd->data[(size_t)v * GATES + n] + Bd->data[n];
            for (int e = 0; e < ENC_OUT; ++e)
                for (int d = 0; d < UNITS; ++d) wmT[(size_t)d * ENC_OUT + e] = Wm->data[(size_t)e * UNITS + d];
            float *tmp = nullptr;
            auto prep = [&](const std::vector<float> &w, int K, int N, float **pair, uint16_t **pair16 = nullptr) -> int {
                RVB_CHECK(upload(m, &tmp, w));
                RVB_CHECK(dmalloc(m, &pair[0], w.size()));
                RVB_CHECK(dmalloc(m, &pair[1], w.size()));
                RVB_CHECK(gemm::prepare_weights(tmp, pair[0], pair[1], K, N, nullptr));
                if (pair16 != nullptr) {        // fp16 hi / lo planes of the same weight: the GEMM then runs on the fp16 pipe
                    RVB_CHECK(dmalloc(m, &pair16[0], w.size()));
                    RVB_CHECK(dmalloc(m, &pair16[1], w.size()));
                    RVB_CHECK(gemm::prepare_weights_f16(tmp, pair16[0], pair16[1], K, N, nullptr));
                }
                return RVB_OK;
            };
            RVB_CHECK(prep(wcat, 2 * UNITS, GATES, m->dw_wg, m->dw_wg16));
            RVB_CHECK(prep(wmT, UNITS, ENC_OUT, m->dw_wm, m->dw_wm16));
            RVB_CHECK(prep(Wa->data, UNITS + ENC_OUT, UNITS, m->dw_wa));
            {   // the same layer for the fp16-plane GEMM, whose column tile is 256 wide: columns 128..255 are zero and never stored
                std::vector<float> wa_pad((size_t)(UNITS + ENC_OUT) * 2 * UNITS, 0.0f);
                for (int k = 0; k < UNITS + ENC_OUT; ++k)
                    for (int n = 0; n < UNITS; ++n) wa_pad[(size_t)k * 2 * UNITS + n] = Wa->data[(size_t)k * UNITS + n];
                RVB_CHECK(upload(m, &tmp, wa_pad));
                RVB_CHECK(dmalloc(m, &m->dw_wa16[0], wa_pad.size()));
                RVB_CHECK(dmalloc(m, &m->dw_wa16[1], wa_pad.size()));
                RVB_CHECK(gemm::prepare_weights_f16(tmp, m->dw_wa16[0], m->dw_wa16[1], UNITS + ENC_OUT, 2 * UNITS, nullptr));
            }
            RVB_CHECK(upload(m, &m->dw_wtok, wtk));
            if (m->dec_depth == 2) {
                // second stacked cell: [kernel (input = h of cell 0) ; recurrent kernel] as one [256,512] weight, [unit][gate] columns
                Cell4 d1;
                RVB_CHECK(get_cell4(m, "decoder/cell1/", UNITS, UNITS, true, &d1));
                struct { std::vector<float> &data; } W1r{d1.W}, U1r{d1.U}, B1r{d1.b};
                auto *W1 = &W1r, *U1 = &U1r, *B1 = &B1r;
                std::vector<float> w1cat((size_t)2 * UNITS * GATES), b1v(GATES);
                for (int k = 0; k < 2 * UNITS; ++k)
                    for (int n = 0; n < GATES; ++n)
                        w1cat[(size_t)k * GATES + (n % UNITS) * 4 + n / UNITS] = k < UNITS ? W1->data[(size_t)k * GATES + n] : U1->data[(size_t)(k - UNITS) * GATES + n];
                for (int n = 0; n < GATES; ++n) b1v[(n % UNITS) * 4 + n / UNITS] = B1->data[n];
                RVB_CHECK(prep(w1cat, 2 * UNITS, GATES, m->dw_wg1, m->dw_wg1_16));
                RVB_CHECK(upload(m, &m->dw_b1, b1v));
            }
        }
        RVB_CHECK(upload(m, &m->d_wfc, Wf->data));
        RVB_CHECK(upload(m, &m->d_bfc, Bf->data));
    }
    if (!m->d_abort) { RVB_CHECK(dmalloc(m, &m->d_abort, 1)); }
    RVB_CUDA(cudaMemset(m->d_abort, 0, sizeof(int)));
    RVB_CUDA(cudaDeviceSynchronize());
    m->finalized = true;
    return RVB_OK;
}

static int ensure_workspace(rvb_model *m, int t_raw, int t_ev, int S, int W) {
    const size_t wv = ((size_t)m->wave + 127) / 128 * 128;       // intermediates are padded to whole 128-row tiles
    const int Tm = t_raw + t_ev;
    if ((size_t)t_raw > m->ws_raw_t) {
        for (int i = 0; i < 2; ++i) { dfree(m, m->y_raw[i]); RVB_CHECK(dmalloc(m, &m->y_raw[i], wv * t_raw * ENC_OUT)); }
        dfree(m, m->G_raw);
        if (m->enc_depth > 1) RVB_CHECK(dmalloc(m, &m->G_raw, wv * t_raw * 2 * GATES));
        m->ws_raw_t = t_raw;
    }
    if ((size_t)t_ev > m->ws_ev_t) {
        for (int i = 0; i < 2; ++i) { dfree(m, m->y_ev[i]); RVB_CHECK(dmalloc(m, &m->y_ev[i], wv * t_ev * ENC_OUT)); }
        dfree(m, m->G_ev);
        if (m->enc_depth > 1) RVB_CHECK(dmalloc(m, &m->G_ev, wv * t_ev * 2 * GATES));
        m->ws_ev_t = t_ev;
    }
    if (!m->st[0][0])
        for (int e = 0; e < 2; ++e)
            for (int i = 0; i < 2; ++i) RVB_CHECK(dmalloc(m, &m->st[e][i], wv * 2 * 2 * UNITS));
    if ((size_t)Tm > m->ws_tm) {
        dfree(m, m->enc_out); dfree(m, m->mask);
        RVB_CHECK(dmalloc(m, &m->enc_out, wv * Tm * ENC_OUT));
        if (m->precision == RVB_PREC_BF16) { dfree(m, m->enc_out16); RVB_CHECK(dmalloc(m, &m->enc_out16, wv * Tm * ENC_OUT)); }
        RVB_CHECK(dmalloc(m, &m->mask, wv * Tm));
        if (m->att_tc) {
            dfree(m, m->enc_hi); dfree(m, m->enc_lo);
            RVB_CHECK(dmalloc(m, &m->enc_hi, wv * Tm * ENC_OUT));
            RVB_CHECK(dmalloc(m, &m->enc_lo, wv * Tm * ENC_OUT));
        }
        m->ws_tm = Tm;
    }
    if ((size_t)S * W > m->ws_sw) {
        dfree(m, m->step_ids); dfree(m, m->parent_ids);
        RVB_CHECK(dmalloc(m, &m->step_ids, wv * S * W));
        RVB_CHECK(dmalloc(m, &m->parent_ids, wv * S * W));
        m->ws_sw = (size_t)S * W;
    }
    return RVB_OK;
}

// One encoder (raw: e = 0, event: e = 1) over nb snippets; final layer writes into
// out[b, t_off + t, :] with row stride Tm*256 (basecaller.py:48-59, 405).
// planes: the final layer writes the memory as fp16 hi / lo planes (m->enc_hi / enc_lo, [nb,Tm,256]) instead of fp32 `out`.
static int encode_branch(rvb_model *m, int e, const float *x, int T, int nb, float *out, int Tm, int t_off, cudaStream_t s, bool planes = false) {
    const int feat = e == 0 ? 1 : 5;
    float **yb = e == 0 ? m->y_raw : m->y_ev;
    float *G = e == 0 ? m->G_raw : m->G_ev;
    for (int l = 0; l < m->enc_depth; ++l) {
        // tensor-core recurrence; intermediates are time-major (row = t*nb + b) so that a tile's rows of one
        // timestep are contiguous for both K2 and K3
        const bool last = (l == m->enc_depth - 1);
        const long long nbp = ((long long)nb + 127) / 128 * 128;    // rows per timestep of the fp16 planes and of the blocked G
        rectc::Params p{};
        p.x = x; p.G = G; p.g_bs = 2 * GATES; p.g_ts = nbp * 2 * GATES; p.g_blocked = 1; p.g_rows_per_t = nbp;
        // reduced precision: the pre-gates cross HBM as fp16 (K2 writes, K3 reads half the bytes of the encoder's largest tensor)
        const bool g16 = m->precision != RVB_PREC_FP32;
        p.g16 = (l > 0 && g16) ? 1 : 0;
        p.bimg = m->d_bimg[e][l]; p.w0 = m->d_w0[e];
        p.state_in = l == 0 ? nullptr : m->st[e][(l - 1) & 1];
        p.state_out = m->st[e][l & 1];
        // intermediate layers hand their output to K2 as fp16 hi/lo planes (time-major, same bytes as fp32)
        uint16_t *pl_hi = reinterpret_cast<uint16_t *>(yb[l & 1]);
        uint16_t *pl_lo = pl_hi + (size_t)nbp * T * ENC_OUT;
        p.y = (last && !planes) ? out + (size_t)t_off * ENC_OUT : nullptr;
        p.y_bs = (long long)Tm * ENC_OUT; p.y_ts = ENC_OUT;
        p.y16_hi = last ? nullptr : pl_hi; p.y16_lo = last ? nullptr : pl_lo;
        p.y16_bs = ENC_OUT; p.y16_ts = nbp * ENC_OUT;
        if (last && planes) {       // batch-major planes of the attention memory
            p.y16_hi = m->enc_hi + (size_t)t_off * ENC_OUT; p.y16_lo = m->enc_lo + (size_t)t_off * ENC_OUT;
            p.y16_bs = (long long)Tm * ENC_OUT; p.y16_ts = ENC_OUT;
        }
        p.yv16 = (last && m->enc_out16 != nullptr && out == m->enc_out) ? m->enc_out16 + (size_t)t_off * ENC_OUT : nullptr;
        p.B = nb; p.T = T; p.abort_flag = m->d_abort; p.precision = m->precision; p.gru = m->cell == RVB_CELL_GRU;
        if (l > 0) {
            const uint16_t *a_hi = reinterpret_cast<const uint16_t *>(yb[(l - 1) & 1]);
            // rows b >= nb of a timestep are padding: never written by K3, projected as they are, never read back
            RVB_CHECK(gemm::run_tc_f16(a_hi, a_hi + (size_t)nbp * T * ENC_OUT, m->d_phi16[e][l], m->d_plo16[e][l], m->d_pb[e][l], G,
                                       nbp * T, 2 * GATES, ENC_OUT, m->precision, m->d_abort, s, true, nullptr, 0, 0, g16));
        }
        RVB_CHECK(rectc::run(l == 0 ? feat : 0, p, s));
    }
    const long long n = (long long)nb * T;
    input_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, feat, nb, T, m->mask, Tm, t_off);
    RVB_LAUNCH_CHECK();
    count_launch();
    return RVB_OK;
}

static int check_inputs(rvb_model *m, const float *raw, int t_raw, const float *ev, int t_ev, int64_t batch, int *Tm) {
    if (!m) return fail(RVB_ERR_ARG, "null model");
    if (!m->finalized) return fail(RVB_ERR_STATE, "weights not finalised (call rvb_model_finalize)");
    if (batch < 0) return fail(RVB_ERR_ARG, "negative batch");
    const bool need_raw = m->input_kind != RVB_INPUT_EVENT, need_ev = m->input_kind != RVB_INPUT_RAW;
    if (need_raw && (!raw || t_raw <= 0) && batch > 0) return fail(RVB_ERR_ARG, "raw input required");
    if (need_ev && (!ev || t_ev <= 0) && batch > 0) return fail(RVB_ERR_ARG, "event input required");
    *Tm = (need_raw ? t_raw : 0) + (need_ev ? t_ev : 0);
    if (*Tm > 256) return fail(RVB_ERR_ARG, "memory length %d exceeds 256", *Tm);
    return RVB_OK;
}

// encoders + mask + keys for one wave, into the handle's workspace
static int encode_wave(rvb_model *m, const float *raw, int t_raw, const float *ev, int t_ev, int nb, int Tm, cudaStream_t s, bool planes = false) {
    const bool need_raw = m->input_kind != RVB_INPUT_EVENT, need_ev = m->input_kind != RVB_INPUT_RAW;
    if (need_raw) RVB_CHECK(encode_branch(m, 0, raw, t_raw, nb, m->enc_out, Tm, 0, s, planes));
    if (need_ev) RVB_CHECK(encode_branch(m, 1, ev, t_ev, nb, m->enc_out, Tm, need_raw ? t_raw : 0, s, planes));
    return RVB_OK;
}

extern "C" int rvb_encode(rvb_model_t *m, const float *d_raw, int t_raw, const float *d_event, int t_event, int64_t batch,
                          float *d_enc_out, uint8_t *d_mask, void *stream) {
    int Tm = 0;
    RVB_CHECK(check_inputs(m, d_raw, t_raw, d_event, t_event, batch, &Tm));
    if (batch == 0) return RVB_OK;
    if (!d_enc_out || !d_mask) return fail(RVB_ERR_ARG, "null output");
    RVB_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = (cudaStream_t)stream;
    const bool need_raw = m->input_kind != RVB_INPUT_EVENT, need_ev = m->input_kind != RVB_INPUT_RAW;
    if (!need_raw) t_raw = 0;
    if (!need_ev) t_event = 0;
    RVB_CHECK(ensure_workspace(m, t_raw, t_event, 1, 1));
    for (int64_t b0 = 0; b0 < batch; b0 += m->wave) {
        const int nb = (int)std::min<int64_t>(m->wave, batch - b0);
        RVB_CHECK(encode_wave(m, need_raw ? d_raw + (size_t)b0 * t_raw : nullptr, t_raw,
                              need_ev ? d_event + (size_t)b0 * t_event * 5 : nullptr, t_event, nb, Tm, s));
        RVB_CUDA(cudaMemcpyAsync(d_enc_out + (size_t)b0 * Tm * ENC_OUT, m->enc_out, (size_t)nb * Tm * ENC_OUT * sizeof(float), cudaMemcpyDeviceToDevice, s));
        RVB_CUDA(cudaMemcpyAsync(d_mask + (size_t)b0 * Tm, m->mask, (size_t)nb * Tm, cudaMemcpyDeviceToDevice, s));
    }
    return RVB_OK;
}

static int search(rvb_model *m, const float *d_raw, int t_raw, const float *d_event, int t_event, int64_t batch, int W,
                  int max_output_len, bool beam, int32_t *d_ids, float *d_logits, float *d_scores, int32_t *d_step_ids,
                  int32_t *d_parent_ids, int32_t *d_steps, cudaStream_t s) {
    int Tm = 0;
    RVB_CHECK(check_inputs(m, d_raw, t_raw, d_event, t_event, batch, &Tm));
    const int S = max_output_len - 1;
    if (!d_steps || !d_ids) return fail(RVB_ERR_ARG, "null output");
    if (beam && (W < 1 || W > 9)) return fail(RVB_ERR_ARG, "beam_width must be in [1,9]");
    RVB_CUDA(cudaSetDevice(m->device));
    RVB_CUDA(cudaMemsetAsync(d_steps, 0, sizeof(int32_t), s));
    if (batch == 0 || S <= 0) return RVB_OK;
    const bool need_raw = m->input_kind != RVB_INPUT_EVENT, need_ev = m->input_kind != RVB_INPUT_RAW;
    if (!need_raw) t_raw = 0;
    if (!need_ev) t_event = 0;
    RVB_CHECK(ensure_workspace(m, t_raw, t_event, S, W));
    for (int64_t b0 = 0; b0 < batch; b0 += m->wave) {
        const int nb = (int)std::min<int64_t>(m->wave, batch - b0);
        // beam widths >= 2: the attention runs on tcgen05 and reads the memory as fp16 hi / lo planes, which the last encoder
        // layer then writes INSTEAD of the fp32 rows (same bytes)
        // reduced precision: the memory is ONE fp16 plane (K3 writes it next to the fp32 rows), also read on tcgen05, at every width
        // Both tcgen05 forms walk a snippet as a chain of tile hand-offs with a floor of ~0.1 ms per launch; the FFMA kernels'
        // time is proportional to the memory length and wins below ~100 rows (event-only model, Tm = 30: 0.03 vs 0.11 ms).
        const bool long_memory = Tm >= 128;
        const bool tc16 = m->precision == RVB_PREC_BF16 && m->att16_tc && long_memory;
        const bool tc_att = beam && m->att_tc && W >= 2 && !tc16 && long_memory;
        RVB_CHECK(encode_wave(m, need_raw ? d_raw + (size_t)b0 * t_raw : nullptr, t_raw,
                              need_ev ? d_event + (size_t)b0 * t_event * 5 : nullptr, t_event, nb, Tm, s, tc_att));
        // Wave-level decoder: every beam width, decoder depth and cell kind, and greedy search (a mode of its search kernel)
        {
            if (!beam) W = 1;
            const size_t rows = (size_t)m->wave * W;
            if (rows > m->dw_ws_rows) {
                dfree(m, m->dw_ws);
                RVB_CHECK(dmalloc(m, &m->dw_ws, decw::workspace_floats((long long)rows, m->dec_depth)));
                m->dw_ws_rows = rows;
            }
            decw::Params q{};
            q.values = m->enc_out; q.mask = m->mask;
            // fp16 copy of the memory: the tcgen05 kernel at every width; the FFMA kernel (RVB_ATT16=ffma) only at width 1, where it
            // is bandwidth bound (measured: at width 5 it is issue bound and the widening conversions cost more than the halved
            // bytes save, 199 -> 251 ms per step)
            q.values16 = (m->precision == RVB_PREC_BF16 && (W == 1 || tc16)) ? m->enc_out16 : nullptr;
            q.att16_tc = tc16 ? 1 : 0;
            q.v_hi = tc_att ? m->enc_hi : nullptr; q.v_lo = tc_att ? m->enc_lo : nullptr;
            q.wg_hiT = m->dw_wg[0]; q.wg_loT = m->dw_wg[1]; q.wm_hiT = m->dw_wm[0]; q.wm_loT = m->dw_wm[1];
            q.wa_hiT = m->dw_wa[0]; q.wa_loT = m->dw_wa[1];
            q.wg16_hi = m->dw_wg16[0]; q.wg16_lo = m->dw_wg16[1]; q.wm16_hi = m->dw_wm16[0]; q.wm16_lo = m->dw_wm16[1]; q.wa16_hi = m->dw_wa16[0]; q.wa16_lo = m->dw_wa16[1]; q.wtok = m->dw_wtok; q.wfc = m->d_wfc; q.bfc = m->d_bfc;
            q.wg1_16_hi = m->dw_wg1_16[0]; q.wg1_16_lo = m->dw_wg1_16[1]; q.b1 = m->dw_b1; q.depth = m->dec_depth;
            q.gru = m->cell == RVB_CELL_GRU; q.greedy = beam ? 0 : 1;
            q.logits = beam ? nullptr : d_logits + (size_t)b0 * S * VOCAB;
            q.B = nb; q.Tm = Tm; q.W = W; q.S = S;
            q.ids = d_ids + (size_t)b0 * S * W; q.scores = beam ? d_scores + (size_t)b0 * S * W : nullptr;
            q.step_ids = d_step_ids ? d_step_ids + (size_t)b0 * S * W : m->step_ids;
            q.parent_ids = d_parent_ids ? d_parent_ids + (size_t)b0 * S * W : m->parent_ids;
            q.steps = d_steps; q.ws = m->dw_ws; q.abort_flag = m->d_abort;
            RVB_CHECK(decw::run(q, s));
        }
    }
    return RVB_OK;
}

extern "C" int rvb_greedy(rvb_model_t *m, const float *d_raw, int t_raw, const float *d_event, int t_event, int64_t batch,
                          int max_output_len, int32_t *d_ids, float *d_logits, int32_t *d_steps, void *stream) {
    if (!d_logits) return fail(RVB_ERR_ARG, "null logits output");
    return search(m, d_raw, t_raw, d_event, t_event, batch, 1, max_output_len, false, d_ids, d_logits, nullptr, nullptr,
                  nullptr, d_steps, (cudaStream_t)stream);
}

extern "C" int rvb_beam(rvb_model_t *m, const float *d_raw, int t_raw, const float *d_event, int t_event, int64_t batch,
                        int beam_width, int max_output_len, int32_t *d_pred_ids, float *d_scores, int32_t *d_step_ids,
                        int32_t *d_parent_ids, int32_t *d_steps, void *stream) {
    if (!d_scores) return fail(RVB_ERR_ARG, "null scores output");
    return search(m, d_raw, t_raw, d_event, t_event, batch, beam_width, max_output_len, true, d_pred_ids, nullptr,
                  d_scores, d_step_ids, d_parent_ids, d_steps, (cudaStream_t)stream);
}

// beam slot 0 of [B,S,W] -> [B,S]
__global__ void slot0_kernel(const int32_t *ids, const float *sc, long long n, int W, int32_t *ids0, float *sc0) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ids0[i] = ids[i * W];
    sc0[i] = sc[i * W];
}

extern "C" int rvb_model_check(rvb_model_t *m);

// Host-buffer pipeline of rvb_beam_host: two device I/O sets, two pinned staging sets, three streams.
// Wave k: [CPU copy user -> pinned in (pageable callers only)] -> H2D on s_in -> compute on s_comp -> slot-0 extraction ->
// D2H on s_out -> [CPU copy pinned out -> user].  Wave k+1's staging + H2D and wave k-1's D2H + unstaging overlap wave
// k's compute; the compute itself is serial (one workspace per handle).
struct HostPipe {
    size_t in_floats = 0, out_words = 0;           // per set
    float *d_in[2] = {nullptr, nullptr};
    int32_t *d_out[2] = {nullptr, nullptr};        // [ids0 | scores0 | steps]
    float *p_in[2] = {nullptr, nullptr};           // pinned staging
    int32_t *p_out[2] = {nullptr, nullptr};
    float *d_full_sc = nullptr; int32_t *d_full_ids = nullptr; size_t full_words = 0;   // [wave,S,W] search outputs
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
};

static bool host_ptr_is_pinned(const void *p) {
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int host_pipe_prepare(rvb_model *m, size_t in_floats, size_t out_words, size_t full_words) {
    if (!m->pipe) m->pipe = new HostPipe();
    HostPipe *hp = m->pipe;
    if (!hp->s_in) {
        RVB_CUDA(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
        RVB_CUDA(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            RVB_CUDA(cudaEventCreateWithFlags(&hp->h2d_done[i], cudaEventDisableTiming));
            RVB_CUDA(cudaEventCreateWithFlags(&hp->comp_done[i], cudaEventDisableTiming));
            RVB_CUDA(cudaEventCreateWithFlags(&hp->d2h_done[i], cudaEventDisableTiming));
        }
    }
    if (in_floats > hp->in_floats) {
        for (int i = 0; i < 2; ++i) {
            dfree(m, hp->d_in[i]);
            RVB_CHECK(dmalloc(m, &hp->d_in[i], in_floats));
            if (hp->p_in[i]) cudaFreeHost(hp->p_in[i]);
            RVB_CUDA(cudaMallocHost(&hp->p_in[i], in_floats * sizeof(float)));
        }
        hp->in_floats = in_floats;
    }
    if (out_words > hp->out_words) {
        for (int i = 0; i < 2; ++i) {
            dfree(m, hp->d_out[i]);
            RVB_CHECK(dmalloc(m, &hp->d_out[i], out_words));
            if (hp->p_out[i]) cudaFreeHost(hp->p_out[i]);
            RVB_CUDA(cudaMallocHost(&hp->p_out[i], out_words * sizeof(int32_t)));
        }
        hp->out_words = out_words;
    }
    if (full_words > hp->full_words) {
        dfree(m, hp->d_full_sc); dfree(m, hp->d_full_ids);
        RVB_CHECK(dmalloc(m, &hp->d_full_sc, full_words));
        RVB_CHECK(dmalloc(m, &hp->d_full_ids, full_words));
        hp->full_words = full_words;
    }
    return RVB_OK;
}

static void host_pipe_destroy(rvb_model *m) {
    HostPipe *hp = m->pipe;
    if (!hp) return;
    for (int i = 0; i < 2; ++i) {
        if (hp->p_in[i]) cudaFreeHost(hp->p_in[i]);
        if (hp->p_out[i]) cudaFreeHost(hp->p_out[i]);
        if (hp->h2d_done[i]) cudaEventDestroy(hp->h2d_done[i]);
        if (hp->comp_done[i]) cudaEventDestroy(hp->comp_done[i]);
        if (hp->d2h_done[i]) cudaEventDestroy(hp->d2h_done[i]);
    }
    if (hp->s_in) cudaStreamDestroy(hp->s_in);
    if (hp->s_out) cudaStreamDestroy(hp->s_out);
    delete hp;
    m->pipe = nullptr;
}

extern "C" int rvb_beam_host(rvb_model_t *m, const float *h_raw, int t_raw, const float *h_event, int t_event, int64_t batch,
                             int beam_width, int max_output_len, int32_t *h_ids, float *h_scores, int32_t *h_steps) {
    int Tm = 0;
    if (!m) return fail(RVB_ERR_ARG, "null model");
    const bool need_raw = m->input_kind != RVB_INPUT_EVENT, need_ev = m->input_kind != RVB_INPUT_RAW;
    RVB_CHECK(check_inputs(m, h_raw, t_raw, h_event, t_event, batch, &Tm));
    if (!h_ids || !h_scores || !h_steps) return fail(RVB_ERR_ARG, "null output");
    const int S = max_output_len - 1, W = beam_width;
    *h_steps = 0;
    if (batch == 0 || S <= 0) return RVB_OK;
    if (W < 1 || W > 9) return fail(RVB_ERR_ARG, "beam_width must be in [1,9]");
    RVB_CUDA(cudaSetDevice(m->device));
    if (!m->hstream) RVB_CUDA(cudaStreamCreateWithFlags(&m->hstream, cudaStreamNonBlocking));
    cudaStream_t sc = m->hstream;
    if (!need_raw) t_raw = 0;
    if (!need_ev) t_event = 0;
    const size_t wv = (size_t)m->wave;
    const size_t raw_f = wv * (size_t)t_raw, ev_f = wv * (size_t)t_event * 5;
    const size_t out_w = wv * (size_t)S * 2 + 4;
    RVB_CHECK(host_pipe_prepare(m, raw_f + ev_f + 4, out_w, wv * (size_t)S * W));
    HostPipe *hp = m->pipe;
    // a caller that already holds page-locked memory is copied from / to directly; pageable memory goes through the pinned
    // staging sets (a cudaMemcpyAsync on pageable memory would serialise against the running wave)
    const bool stage_in = !(host_ptr_is_pinned(need_raw ? h_raw : nullptr) && host_ptr_is_pinned(need_ev ? h_event : nullptr));
    const bool stage_out = !(host_ptr_is_pinned(h_ids) && host_ptr_is_pinned(h_scores));
    const int64_t nw = (batch + m->wave - 1) / m->wave;
    auto wave_rows = [&](int64_t k) { return (int)std::min<int64_t>(m->wave, batch - k * m->wave); };

    auto issue_h2d = [&](int64_t k) -> int {
        const int set = (int)(k & 1), nb = wave_rows(k);
        const size_t b0 = (size_t)k * wv;
        if (k >= 2) {
            if (stage_in) RVB_CUDA(cudaEventSynchronize(hp->h2d_done[set]));          // pinned set: its previous H2D has left
            RVB_CUDA(cudaStreamWaitEvent(hp->s_in, hp->comp_done[set], 0));          // device set: wave k-2 consumed it
        }
        float *d_raw = hp->d_in[set], *d_ev = d_raw + raw_f;
        const float *src_raw = need_raw ? h_raw + b0 * t_raw : nullptr;
        const float *src_ev = need_ev ? h_event + b0 * t_event * 5 : nullptr;
        if (stage_in) {
            float *p_raw = hp->p_in[set], *p_ev = p_raw + raw_f;
            if (need_raw) { memcpy(p_raw, src_raw, (size_t)nb * t_raw * sizeof(float)); src_raw = p_raw; }
            if (need_ev) { memcpy(p_ev, src_ev, (size_t)nb * t_event * 5 * sizeof(float)); src_ev = p_ev; }
        }
        if (need_raw) RVB_CUDA(cudaMemcpyAsync(d_raw, src_raw, (size_t)nb * t_raw * sizeof(float), cudaMemcpyHostToDevice, hp->s_in));
        if (need_ev) RVB_CUDA(cudaMemcpyAsync(d_ev, src_ev, (size_t)nb * t_event * 5 * sizeof(float), cudaMemcpyHostToDevice, hp->s_in));
        RVB_CUDA(cudaEventRecord(hp->h2d_done[set], hp->s_in));
        return RVB_OK;
    };
    auto finish = [&](int64_t k) -> int {                 // wave k's results are in the pinned / user buffers after this
        const int set = (int)(k & 1), nb = wave_rows(k);
        const size_t b0 = (size_t)k * wv, n = (size_t)nb * S;
        RVB_CUDA(cudaEventSynchronize(hp->d2h_done[set]));
        const int32_t *po = hp->p_out[set];
        if (stage_out) {
            memcpy(h_ids + b0 * S, po, n * sizeof(int32_t));
            memcpy(h_scores + b0 * S, po + wv * S, n * sizeof(float));
        }
        const int32_t wave_steps = po[2 * wv * S];
        if (wave_steps > *h_steps) *h_steps = wave_steps;
        return RVB_OK;
    };

    RVB_CHECK(issue_h2d(0));
    for (int64_t k = 0; k < nw; ++k) {
        const int set = (int)(k & 1), nb = wave_rows(k);
        const size_t b0 = (size_t)k * wv;
        const long long n = (long long)nb * S;
        if (k + 1 < nw) RVB_CHECK(issue_h2d(k + 1));       // staging + H2D of the next wave run under this wave's compute
        int32_t *d_ids0 = hp->d_out[set];
        float *d_sc0 = reinterpret_cast<float *>(d_ids0 + wv * S);
        int32_t *d_steps = d_ids0 + 2 * wv * S;
        RVB_CUDA(cudaStreamWaitEvent(sc, hp->h2d_done[set], 0));
        if (k >= 2) RVB_CUDA(cudaStreamWaitEvent(sc, hp->d2h_done[set], 0));          // output set: wave k-2 has been read back
        float *d_raw = hp->d_in[set], *d_ev = d_raw + raw_f;
        RVB_CHECK(search(m, need_raw ? d_raw : nullptr, t_raw, need_ev ? d_ev : nullptr, t_event, nb, W, max_output_len, true,
                         hp->d_full_ids, nullptr, hp->d_full_sc, nullptr, nullptr, d_steps, sc));
        slot0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, sc>>>(hp->d_full_ids, hp->d_full_sc, n, W, d_ids0, d_sc0);
        RVB_LAUNCH_CHECK();
        count_launch();
        RVB_CUDA(cudaEventRecord(hp->comp_done[set], sc));
        RVB_CUDA(cudaStreamWaitEvent(hp->s_out, hp->comp_done[set], 0));
        int32_t *po = hp->p_out[set];
        RVB_CUDA(cudaMemcpyAsync(stage_out ? po : h_ids + b0 * S, d_ids0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, hp->s_out));
        RVB_CUDA(cudaMemcpyAsync(stage_out ? reinterpret_cast<void *>(po + wv * S) : reinterpret_cast<void *>(h_scores + b0 * S), d_sc0,
                                 n * sizeof(float), cudaMemcpyDeviceToHost, hp->s_out));
        RVB_CUDA(cudaMemcpyAsync(po + 2 * wv * S, d_steps, sizeof(int32_t), cudaMemcpyDeviceToHost, hp->s_out));
        RVB_CUDA(cudaEventRecord(hp->d2h_done[set], hp->s_out));
        if (k >= 1) RVB_CHECK(finish(k - 1));
    }
    RVB_CHECK(finish(nw - 1));
    return rvb_model_check(m);
}

// Per-kernel device time since rvb_profile(1): ms[k], launches[k] for k in KernelKind order
// (event scan, projection GEMM, recurrent LSTM, decoder, other, decoder attention).  Synchronises the device.
extern "C" int rvb_profile(int enable) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &e : g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g_prof.clear();
    g_prof_on.store(enable ? 1 : 0);
    return RVB_OK;
}
extern "C" int rvb_profile_read(double *ms, int64_t *launches, int n) {
    if (!ms || !launches || n < KK_COUNT) return fail(RVB_ERR_ARG, "profile_read: need %d slots", (int)KK_COUNT);
    RVB_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int k = 0; k < n; ++k) { ms[k] = 0.0; launches[k] = 0; }
    for (auto &e : g_prof) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) { ms[e.kind] += t; launches[e.kind] += 1; }
    }
    return RVB_OK;
}

extern "C" int rvb_model_check(rvb_model_t *m) {
    if (!m) return fail(RVB_ERR_ARG, "null model");
    RVB_CUDA(cudaSetDevice(m->device));
    RVB_CUDA(cudaDeviceSynchronize());
    int flag = 0;
    if (m->d_abort) RVB_CUDA(cudaMemcpy(&flag, m->d_abort, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag != 0) return fail(RVB_ERR_INTERNAL, "a tcgen05 projection kernel timed out on an mbarrier (results invalid)");
    return RVB_OK;
}

extern "C" int rvb_project(const float *d_a, const float *d_b, const float *d_bias, float *d_c, int64_t mrows, int n, int k,
                           int precision, void *stream) {
    if (!d_a || !d_b || !d_c || mrows < 0 || n <= 0 || k <= 0) return fail(RVB_ERR_ARG, "project: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == -1) return gemm::run_simt(d_a, d_b, d_bias, d_c, mrows, n, k, s);
    if (precision == 2 || precision == 3) {
        // fp16-plane variant (what the encoders use between layers): split A and W into fp16 hi/lo planes first
        char *scratch16 = nullptr;
        const size_t na = (size_t)mrows * k, nw = (size_t)n * k;
        RVB_CUDA(cudaMalloc(&scratch16, (2 * na + 2 * nw) * 2 + 64));
        uint16_t *ahi = reinterpret_cast<uint16_t *>(scratch16), *alo = ahi + na, *whi = alo + na, *wlo = whi + nw;
        int *flag16 = reinterpret_cast<int *>(wlo + nw + (nw & 1));
        int st16 = RVB_OK;
        if (cudaMemsetAsync(flag16, 0, sizeof(int), s) != cudaSuccess) st16 = fail(RVB_ERR_CUDA, "memset");
        if (st16 == RVB_OK) st16 = gemm::split_planes_f16(d_a, ahi, alo, (long long)na, s);
        if (st16 == RVB_OK) st16 = gemm::prepare_weights_f16(d_b, whi, wlo, k, n, s);
        if (st16 == RVB_OK) st16 = gemm::run_tc_f16(ahi, alo, whi, wlo, d_bias, d_c, mrows, n, k, precision == 2 ? RVB_PREC_FP32 : RVB_PREC_BF16, flag16, s);
        int hf = 0;
        cudaError_t e16 = cudaStreamSynchronize(s);
        if (st16 == RVB_OK && e16 != cudaSuccess) st16 = fail(RVB_ERR_CUDA, "project: %s", cudaGetErrorString(e16));
        if (st16 == RVB_OK) cudaMemcpy(&hf, flag16, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(scratch16);
        if (st16 == RVB_OK && hf != 0) st16 = fail(RVB_ERR_INTERNAL, "project: tcgen05 kernel timed out on an mbarrier");
        return st16;
    }
    if (precision != RVB_PREC_FP32 && precision != RVB_PREC_BF16) return fail(RVB_ERR_ARG, "project: bad precision");
    // standalone entry (tests / roofline): split + transpose the weights into scratch, run, check the abort flag
    float *scratch = nullptr;
    RVB_CUDA(cudaMalloc(&scratch, ((size_t)2 * n * k + 16) * sizeof(float)));
    float *hiT = scratch, *loT = scratch + (size_t)n * k;
    int *flag = reinterpret_cast<int *>(scratch + (size_t)2 * n * k);
    int st = RVB_OK;
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), s);
    if (e != cudaSuccess) st = fail(RVB_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    if (st == RVB_OK) st = gemm::prepare_weights(d_b, hiT, loT, k, n, s);
    if (st == RVB_OK) st = gemm::run_tc(d_a, hiT, loT, d_bias, d_c, mrows, n, k, precision, flag, s);
    int h_flag = 0;
    e = cudaStreamSynchronize(s);
    if (st == RVB_OK && e != cudaSuccess) st = fail(RVB_ERR_CUDA, "project: %s", cudaGetErrorString(e));
    if (st == RVB_OK) cudaMemcpy(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(scratch);
    if (st == RVB_OK && h_flag != 0) st = fail(RVB_ERR_INTERNAL, "project: tcgen05 kernel timed out on an mbarrier");
    return st;
}
