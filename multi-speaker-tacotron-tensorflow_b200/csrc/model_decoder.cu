// Decoder (teacher-forced): hoisted GEMMs around the attention recurrence (attention.cu) and the two residual GRU
// layers (gru.cu), forward and backward.       reference: models/tacotron.py:127-214, models/helpers.py:35-67
#include "model.h"

namespace taco {

static taco_gemm_desc gd(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc) {
    taco_gemm_desc d{};
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
    d.alpha = 1.f; d.split_k = 1;
    return d;
}
static int wsplit(int M, int N, long long rows) {
    long long tiles = (long long)cdiv(M, 128) * cdiv(N, 128);      // 128x128 tensor-core tiles, two CTAs per SM
    long long want = (2 * 148 + tiles - 1) / tiles;
    long long maxs = rows / 256 > 0 ? rows / 256 : 1;
    long long sp = want < maxs ? want : maxs;
    if (sp < 1) sp = 1;
    if (sp > 64) sp = 64;
    return (int)sp;
}
static taco_gemm_desc wgrad(const float* X, int ldx, const float* dY, int ldy, float* dW, int Kin, int Nout, long long rows) {
    taco_gemm_desc d = gd(X, dY, dW, Kin, Nout, (int)rows, ldx, ldy, Nout);
    d.transA = 1; d.accumulate = 1; d.split_k = wsplit(Kin, Nout, rows);
    return d;
}

struct DecDims { int N, Ti, Td, To, M, r, E, A, HA, Z1, Z, SPK, Y; };
static DecDims dec_dims(const Model& m) {
    DecDims d;
    d.N = m.shape.N; d.Ti = m.shape.Ti; d.Td = m.shape.Td; d.To = m.shape.To;
    d.M = m.cfg.num_mels; d.r = m.cfg.reduction_factor;
    d.E = 2 * m.cfg.enc_rnn_size; d.A = m.cfg.attention_size; d.HA = m.cfg.attention_state_size;
    d.Z1 = m.cfg.dec_prenet_sizes[0]; d.Z = m.cfg.dec_prenet_sizes[1];
    d.SPK = (m.cfg.speaker_mode == TACO_SPK_SIMPLE) ? m.cfg.speaker_embedding_size : 0;
    d.Y = m.cfg.dec_rnn_size;
    return d;
}

static void fill_att_weights(const Model& m, const DecDims& D, AttArgs& a) {
    a.W1c = m.P("dec_prenet/dense_1/kernel") + (long long)D.M * D.Z1;
    a.W2 = m.P("dec_prenet/dense_2/kernel"); a.b2 = m.P("dec_prenet/dense_2/bias");
    a.Wg = m.P("attention_gru/gates_kernel"); a.bg = m.P("attention_gru/gates_bias");
    a.Wc = m.P("attention_gru/cand_kernel"); a.bc = m.P("attention_gru/cand_bias");
    a.Wq = m.P("attention/query_kernel"); a.v = m.P("attention/v");
    a.score_bias = m.has("attention/score_bias") ? m.P("attention/score_bias") : nullptr;
    a.att_g = m.has("attention/g") ? m.P("attention/g") : nullptr;
    a.att_b = m.has("attention/b") ? m.P("attention/b") : nullptr;
    a.Wo = m.P("concat_proj/kernel"); a.bo = m.P("concat_proj/bias");
}

// The fast attention kernels pull their operands (bf16 weight slices per cluster rank, key / memory slices per CTA) as
// ready-made shared-memory images (att_fast.cu).  Weight images depend on the parameters only: both directions are built
// once per step, beside the encoder when the side stream is available.  Images exist in training plans only.
static bool images_applicable(const Model& m, const AttArgs& a) {
    return m.shape.training && a.fast && !a.free_run && att_fast_supported(a) && m.has_region("dec/img_fw");
}
static AttArgs att_dims(const Model& m, const DecDims& D) {
    AttArgs a{};
    a.N = D.N; a.Ti = D.Ti; a.Td = D.Td; a.E = D.E; a.A = D.A; a.HA = D.HA; a.Z1 = D.Z1; a.Z = D.Z; a.SPK = D.SPK; a.Y = D.Y;
    a.att_type = m.cfg.attention_type; a.fast = (m.cfg.precision != TACO_PREC_FP32);
    return a;
}
int decoder_pack_weight_images(Model& m, cudaStream_t s) {
    const DecDims D = dec_dims(m);
    AttArgs a = att_dims(m, D);
    if (!images_applicable(m, a)) return TACO_OK;
    fill_att_weights(m, D, a);
    TACO_TRY(launch_att_fast_pack(a, false, 0, reinterpret_cast<uint8_t*>(m.W("dec/img_fw")), nullptr, s));
    TACO_TRY(launch_att_fast_pack(a, true, 0, reinterpret_cast<uint8_t*>(m.W("dec/img_bw")), nullptr, s));
    TACO_TRY(image_ready(0, s));
    m.img_w_state = 2;
    return TACO_OK;
}

static int dec_gru_layer(Model& m, const DecDims& D, int layer, const float* x, float* y, int training, const float* h0, cudaStream_t s) {
    const std::string gn = "dec_gru_" + std::to_string(layer);
    const std::string rp = "dec/g" + std::to_string(layer) + "_";
    const int Y = D.Y, rows = D.N * D.Td;
    float* gx = m.W(rp + "gx");
    taco_gemm_desc d[2];
    d[0] = gd(x, m.P(gn + "/gates_kernel"), gx, rows, 2 * Y, Y, Y, 2 * Y, 3 * Y); d[0].bias = m.P(gn + "/gates_bias");
    d[1] = gd(x, m.P(gn + "/cand_kernel"), gx + 2 * Y, rows, Y, Y, Y, Y, 3 * Y); d[1].bias = m.P(gn + "/cand_bias");
    TACO_TRY(launch_gemm(d, 2, m.cfg.precision, s));
    GruArgs a{};
    a.N = D.N; a.T = D.Td; a.H = Y; a.ndir = 1; a.fast = (m.cfg.precision != TACO_PREC_FP32);
    a.gx = gx; a.gx_ld = 3 * Y; a.gx_rs_n = D.Td; a.gx_row0 = 0;
    a.Wg[0] = m.P(gn + "/gates_kernel") + (long long)Y * 2 * Y; a.Wc[0] = m.P(gn + "/cand_kernel") + (long long)Y * Y;
    a.h0 = h0; a.res = x; a.res_ld = Y; a.out = y; a.out_ld = Y;
    if (training) { a.st_r = m.W(rp + "st_r"); a.st_u = m.W(rp + "st_u"); a.st_c = m.W(rp + "st_c"); a.st_hprev = m.W(rp + "st_hprev"); }
    return prof_launch_gru(a, false, s);
}

// ---- decoder wavefront ---------------------------------------------------------------------------------------------
// Teacher forcing makes the attention recurrence independent of the two residual GRU layers, and each of the three
// recurrences occupies only a fraction of the SMs (64 / 32 / 32) for a latency-bound time.  Run back to back they cost
// t_att + t_gru1 + t_gru2; cut into time chunks they pipeline: while the attention kernel works on chunk c+1, GRU layer 1
// runs chunk c and layer 2 chunk c-1, each on its own high-priority stream (the reverse order in BPTT).  A chunk launch
// covers steps [t0, t1) and chains to the next through the kernels' carry arguments (kernels.h).  The hoisted x-side
// GEMM of a chunk reads its rows (n, t0..t1) gathered into a contiguous operand and writes straight into the full
// [N, Td, 3Y] layout through the GEMM's row remap, so every other consumer keeps its indexing.
struct Chunk { int t0, t1; };
// Chunk boundaries.  The pipeline is bound by the attention stage (a GRU chunk takes about a third of an attention chunk),
// so what is left to trim is its fill / drain: the chunk at the END of the sequence is the one whose GRU layers trail the
// last attention chunk in the forward pass and lead the first one in BPTT.  It is made short (and the one before it twice
// as long); the rest of the sequence is split evenly.
static std::vector<Chunk> wave_chunks(int Td, int want) {
    std::vector<Chunk> out;
    const int n = std::max(1, std::min(want, Td / 8));            // keep chunks >= 8 steps
    if (n == 1) { out.push_back({0, Td}); return out; }
    const int last = std::max(8, (2 * Td) / (5 * n));
    std::vector<int> len(n, 0);
    len[n - 1] = last;
    int rest = Td - last;
    if (n >= 3) { len[n - 2] = std::min(2 * last, rest - 8 * (n - 2)); if (len[n - 2] < 8) len[n - 2] = 8; rest -= len[n - 2]; }
    const int even = n >= 3 ? n - 2 : 1;
    for (int i = 0; i < even; i++) len[i] = rest / even + (i < rest % even ? 1 : 0);
    int t0 = 0;
    for (int i = 0; i < n; i++) { if (len[i] <= 0) continue; out.push_back({t0, t0 + len[i]}); t0 += len[i]; }
    out.back().t1 = Td;
    return out;
}
static bool wave_applicable(const Model& m, const AttArgs& a, const taco_batch* b) {
    return m.shape.training && !b->rnn_decoder_test_mode && b->mel_targets && a.fast && !a.free_run && att_fast_supported(a) &&
           m.cfg.dec_rnn_size == 256 && m.shape.Td >= 16;
}

// one chunk of a residual GRU layer, forward: gather x rows -> x-side GEMM (remapped into gx) -> recurrence over [t0, t1)
static int dec_gru_chunk_fwd(Model& m, const DecDims& D, int layer, const float* x, float* y, const float* h0, Chunk c, cudaStream_t s) {
    const std::string gn = "dec_gru_" + std::to_string(layer);
    const std::string rp = "dec/g" + std::to_string(layer) + "_";
    const int Y = D.Y, Tc = c.t1 - c.t0, rows = D.N * Tc;
    float* gx = m.W(rp + "gx"); float* xc = m.W(rp + "wv_x"); float* carry = m.W(rp + "wv_h");
    TACO_TRY(launch_copy2d(xc, x + (long long)c.t0 * Y, D.N, Tc * Y, (long long)Tc * Y, (long long)D.Td * Y, s));
    taco_gemm_desc d[2];
    d[0] = gd(xc, m.P(gn + "/gates_kernel"), gx + (long long)c.t0 * 3 * Y, rows, 2 * Y, Y, Y, 2 * Y, 3 * Y); d[0].bias = m.P(gn + "/gates_bias");
    d[1] = gd(xc, m.P(gn + "/cand_kernel"), gx + (long long)c.t0 * 3 * Y + 2 * Y, rows, Y, Y, Y, Y, 3 * Y); d[1].bias = m.P(gn + "/cand_bias");
    for (taco_gemm_desc& g : d) { g.remap_period = Tc; g.remap_outer = (long long)D.Td * 3 * Y; g.remap_inner = 3 * Y; }
    TACO_TRY(launch_gemm(d, 2, m.cfg.precision, s));
    GruArgs a{};
    a.N = D.N; a.T = D.Td; a.H = Y; a.ndir = 1; a.fast = 1;
    a.gx = gx; a.gx_ld = 3 * Y; a.gx_rs_n = D.Td; a.gx_row0 = 0;
    a.Wg[0] = m.P(gn + "/gates_kernel") + (long long)Y * 2 * Y; a.Wc[0] = m.P(gn + "/cand_kernel") + (long long)Y * Y;
    a.h0 = (c.t0 == 0) ? h0 : carry; a.hfinal = carry;
    a.res = x; a.res_ld = Y; a.out = y; a.out_ld = Y;
    a.st_r = m.W(rp + "st_r"); a.st_u = m.W(rp + "st_u"); a.st_c = m.W(rp + "st_c"); a.st_hprev = m.W(rp + "st_hprev");
    a.t_begin = c.t0; a.t_end = c.t1;
    return prof_launch_gru(a, false, s);
}

static int decoder_forward_wave(Model& m, const DecDims& D, AttArgs a, const float* h1, const float* h2, WaveCtx& wv, cudaStream_t s) {
    const std::vector<Chunk> ch = wave_chunks(D.Td, wv.chunks);
    const int C = (int)ch.size();
    for (int c = 0; c < C; c++) {
        a.t_begin = ch[c].t0; a.t_end = ch[c].t1;
        TACO_TRY(prof_launch_att(a, false, s));
        TACO_CHECK_CUDA(cudaEventRecord(wv.ev[0][c], s));
        TACO_CHECK_CUDA(cudaStreamWaitEvent(wv.w[0], wv.ev[0][c], 0));
        TACO_TRY(dec_gru_chunk_fwd(m, D, 1, m.W("dec/y0"), m.W("dec/y1"), h1, ch[c], wv.w[0]));
        TACO_CHECK_CUDA(cudaEventRecord(wv.ev[1][c], wv.w[0]));
        TACO_CHECK_CUDA(cudaStreamWaitEvent(wv.w[1], wv.ev[1][c], 0));
        TACO_TRY(dec_gru_chunk_fwd(m, D, 2, m.W("dec/y1"), m.W("dec/y2"), h2, ch[c], wv.w[1]));
    }
    TACO_CHECK_CUDA(cudaEventRecord(wv.ev[2][0], wv.w[1]));        // layer 2 is ordered behind layer 1, which is behind attention
    TACO_CHECK_CUDA(cudaStreamWaitEvent(s, wv.ev[2][0], 0));
    return TACO_OK;
}

int decoder_forward(Model& m, const taco_batch* b, cudaStream_t s) {
    const DecDims D = dec_dims(m);
    const int prec = m.cfg.precision, training = m.shape.training;
    const bool free_run = (b->mel_targets == nullptr) || b->rnn_decoder_test_mode;      // helpers.py:26-32,63-64
    const int rows = D.N * D.Td;
    const float* memory = m.W("enc_cbhg/rnn_out");
    // attention keys = memory . Wm (no bias, no length mask)   tacotron.py:133-134
    {
        taco_gemm_desc d = gd(memory, m.P("attention/memory_kernel"), m.W("dec/keys"), D.N * D.Ti, D.A, D.E, D.E, D.A, D.A);
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    // teacher-forcing inputs and the hoisted x-side of prenet layer 1   helpers.py:44,60-67; rnn_wrappers.py:249,367-369
    if (!free_run) {
        TACO_TRY(launch_teacher_inputs(b->mel_targets, m.W("dec/x_all"), D.N, D.Td, D.To, D.r, D.M, s));
        taco_gemm_desc d = gd(m.W("dec/x_all"), m.P("dec_prenet/dense_1/kernel"), m.W("dec/px"), rows, D.Z1, D.M, D.M, D.Z1, D.Z1);
        d.bias = m.P("dec_prenet/dense_1/bias");
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    AttArgs a{};
    a.N = D.N; a.Ti = D.Ti; a.Td = D.Td; a.E = D.E; a.A = D.A; a.HA = D.HA; a.Z1 = D.Z1; a.Z = D.Z; a.SPK = D.SPK; a.Y = D.Y;
    a.att_type = m.cfg.attention_type; a.fast = (prec != TACO_PREC_FP32);
    a.px = m.W("dec/px"); a.memory = memory; a.keys = m.W("dec/keys");
    a.spk = D.SPK ? m.W("spk/embed") : nullptr;
    a.ha0 = m.has_region("spk/att_init") ? m.W("spk/att_init") : nullptr;
    a.manual = b->manual_alignments;
    fill_att_weights(m, D, a);
    a.y0 = m.W("dec/y0"); a.align = m.W("alignments");
    const float* h1 = m.has_region("spk/dec_init1") ? m.W("spk/dec_init1") : nullptr;
    const float* h2 = m.has_region("spk/dec_init2") ? m.W("spk/dec_init2") : nullptr;
    if (free_run) {
        // the whole decoder step (prenet x-part, both residual GRUs, mel projection) runs inside the recurrence kernel
        const CbhgGeom& g = m.post;
        a.free_run = 1; a.M = D.M; a.r = D.r;
        a.W1x = m.P("dec_prenet/dense_1/kernel"); a.b1 = m.P("dec_prenet/dense_1/bias");
        a.Wg1 = m.P("dec_gru_1/gates_kernel"); a.bg1 = m.P("dec_gru_1/gates_bias"); a.Wc1 = m.P("dec_gru_1/cand_kernel"); a.bc1 = m.P("dec_gru_1/cand_bias");
        a.Wg2 = m.P("dec_gru_2/gates_kernel"); a.bg2 = m.P("dec_gru_2/gates_bias"); a.Wc2 = m.P("dec_gru_2/cand_kernel"); a.bc2 = m.P("dec_gru_2/cand_bias");
        a.Wmel = m.P("mel_proj/kernel"); a.bmel = m.P("mel_proj/bias");
        a.h1_0 = h1; a.h2_0 = h2;
        a.mel_out = m.W("post_cbhg/xin_p") + (long long)g.PL * D.M; a.mel_bs = (long long)g.Tp * D.M;
        a.y0 = nullptr;
        // low-batch synthesis in the tensor-core modes: every weight of the step resident in the shared memory of one 16-CTA
        // cluster per utterance (att_free.cu); falls through to the general kernel when such a cluster cannot be launched
        static const int use_free = [] { const char* e = getenv("TACO_ATT_FREE"); return e ? atoi(e) : 1; }();
        if (use_free && att_free_supported(a) && m.has_region("dec/frimg")) {
            const int rc = prof_launch_att_free(a, m.W("dec/frimg"), s);
            if (rc != TACO_ENOTSUP) return rc;
        }
        // tensor-core modes: the step's weights are packed once per call as bf16 mma fragments (attention.cu)
        if (att_wfrag_supported(a) && m.has_region("dec/wfrag")) TACO_TRY(launch_att_wfrag_pack(a, m.W("dec/wfrag"), s));
        return prof_launch_att(a, false, s);
    }
    if (training) {
        a.s_z1 = m.W("dec/s_z1"); a.s_z = m.W("dec/s_z"); a.s_r = m.W("dec/s_r"); a.s_u = m.W("dec/s_u"); a.s_c = m.W("dec/s_c");
        a.s_haprev = m.W("dec/s_haprev"); a.s_ha = m.W("dec/s_ha"); a.s_q = m.W("dec/s_q"); a.s_ctxin = m.W("dec/s_ctxin");
        a.s_ctx = m.W("dec/s_ctx"); a.s_e = m.W("dec/s_e"); a.s_a = m.W("dec/s_a");
    }
    if (images_applicable(m, a)) {
        if (m.img_w_state == 0) { TACO_TRY(decoder_pack_weight_images(m, s)); m.img_w_state = 1; }
        else if (m.img_w_state == 2) TACO_TRY(image_wait(0, s));
        uint8_t* fkm = reinterpret_cast<uint8_t*>(m.W("dec/img_fkm"));
        TACO_TRY(launch_att_fast_pack(a, false, 1, nullptr, fkm, s));
        a.img_w = reinterpret_cast<const uint8_t*>(m.W("dec/img_fw")); a.img_km = fkm;
        // the backward pass's key / memory image: same inputs, needed much later -> leaf stream
        cudaStream_t side = fork_side_after(s);
        TACO_TRY(launch_att_fast_pack(a, true, 1, nullptr, reinterpret_cast<uint8_t*>(m.W("dec/img_bkm")), side));
        if (side != s) { TACO_TRY(image_ready(1, side)); m.img_bkm_state = 2; } else m.img_bkm_state = 1;
    }
    WaveCtx* wv = nullptr;
    if (wave_applicable(m, a, b)) TACO_TRY(wave_get(&wv));
    if (wv) {
        TACO_TRY(decoder_forward_wave(m, D, a, h1, h2, *wv, s));
    } else {
        TACO_TRY(prof_launch_att(a, false, s));
        prof_mark("dec_f:gru", s);
        // two ResidualWrapper(GRUCell) layers   tacotron.py:171-175
        TACO_TRY(dec_gru_layer(m, D, 1, m.W("dec/y0"), m.W("dec/y1"), training, h1, s));
        TACO_TRY(dec_gru_layer(m, D, 2, m.W("dec/y1"), m.W("dec/y2"), training, h2, s));
    }
    // r-frame mel projection, written straight into the post-net's padded input (= mel_outputs)   tacotron.py:178-179,213-214
    {
        const CbhgGeom& g = m.post;
        taco_gemm_desc d = gd(m.W("dec/y2"), m.P("mel_proj/kernel"), m.W("post_cbhg/xin_p") + (long long)g.PL * D.M,
                              rows, D.M * D.r, D.Y, D.Y, D.M * D.r, D.M * D.r);
        d.bias = m.P("mel_proj/bias");
        d.remap_period = D.Td; d.remap_outer = (long long)g.Tp * D.M; d.remap_inner = (long long)D.M * D.r;
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    return TACO_OK;
}

static int dec_gru_layer_bwd(Model& m, const DecDims& D, int layer, const float* x, const float* dy, float* dx, bool want_dh0, cudaStream_t s) {
    // y = x + h(x): dx = dy + dgx . Wx^T ; parameter grads accumulated
    const std::string gn = "dec_gru_" + std::to_string(layer);
    const std::string rp = "dec/g" + std::to_string(layer) + "_";
    const int Y = D.Y, rows = D.N * D.Td, prec = m.cfg.precision;
    float* dgx = m.W(rp + "dgx");
    TACO_CHECK_CUDA(cudaMemsetAsync(dgx, 0, sizeof(float) * (size_t)rows * 3 * Y, s));
    GruArgs a{};
    a.N = D.N; a.T = D.Td; a.H = Y; a.ndir = 1; a.fast = (m.cfg.precision != TACO_PREC_FP32);
    a.gx = m.W(rp + "gx"); a.gx_ld = 3 * Y; a.gx_rs_n = D.Td; a.gx_row0 = 0;
    a.Wg[0] = m.P(gn + "/gates_kernel") + (long long)Y * 2 * Y; a.Wc[0] = m.P(gn + "/cand_kernel") + (long long)Y * Y;
    a.st_r = m.W(rp + "st_r"); a.st_u = m.W(rp + "st_u"); a.st_c = m.W(rp + "st_c"); a.st_hprev = m.W(rp + "st_hprev");
    a.dout = dy; a.dout_ld = Y; a.dgx = dgx;
    a.dh0 = want_dh0 ? m.W(rp + "dh0") : nullptr;
    if (a.fast) { a.dbg[0] = m.G(gn + "/gates_bias"); a.dbc[0] = m.G(gn + "/cand_bias"); }      // bias gradients accumulated in the kernel
    TACO_TRY(prof_launch_gru(a, true, s));
    taco_gemm_desc w[4];
    w[0] = wgrad(x, Y, dgx, 3 * Y, m.G(gn + "/gates_kernel"), Y, 2 * Y, rows);
    w[1] = wgrad(x, Y, dgx + 2 * Y, 3 * Y, m.G(gn + "/cand_kernel"), Y, Y, rows);
    w[2] = wgrad(m.W(rp + "st_hprev"), Y, dgx, 3 * Y, m.G(gn + "/gates_kernel") + (long long)Y * 2 * Y, Y, 2 * Y, rows);
    w[3] = wgrad(m.W(rp + "st_r"), Y, dgx + 2 * Y, 3 * Y, m.G(gn + "/cand_kernel") + (long long)Y * Y, Y, Y, rows);
    cudaStream_t leaf = fork_side(s);       // parameter gradients are leaves of the backward graph (model.cu: stream scheduler)
    TACO_TRY(launch_gemm(w, 4, prec, leaf));
    if (!a.fast) {
        TACO_TRY(launch_colsum(dgx, m.G(gn + "/gates_bias"), rows, 2 * Y, 3 * Y, leaf));
        TACO_TRY(launch_colsum(dgx + 2 * Y, m.G(gn + "/cand_bias"), rows, Y, 3 * Y, leaf));
    }
    TACO_CHECK_CUDA(cudaMemcpyAsync(dx, dy, sizeof(float) * (size_t)rows * Y, cudaMemcpyDeviceToDevice, s));
    // dx += dgx . Wx^T in one GEMM (K = 3Y) against the packed x-side rows of both kernels (model.cu: backward_prep)
    taco_gemm_desc e = gd(dgx, m.W(rp + "wxcat"), dx, rows, Y, 3 * Y, 3 * Y, 3 * Y, Y); e.transB = 1; e.accumulate = 1;
    TACO_TRY(launch_gemm(&e, 1, prec, s));
    return TACO_OK;
}

// one chunk of a residual GRU layer, backward: BPTT over [t0, t1) (downwards), then dx rows = dy rows + dgx rows . Wx^T
static int dec_gru_chunk_bwd(Model& m, const DecDims& D, int layer, const float* dy, float* dx, bool last_chunk, bool want_dh0, Chunk c, cudaStream_t s) {
    const std::string gn = "dec_gru_" + std::to_string(layer);
    const std::string rp = "dec/g" + std::to_string(layer) + "_";
    const int Y = D.Y, Tc = c.t1 - c.t0, rows = D.N * Tc, prec = m.cfg.precision;
    float* dgx = m.W(rp + "dgx"); float* carry = m.W(rp + "wv_h");
    GruArgs a{};
    a.N = D.N; a.T = D.Td; a.H = Y; a.ndir = 1; a.fast = 1;
    a.gx = m.W(rp + "gx"); a.gx_ld = 3 * Y; a.gx_rs_n = D.Td; a.gx_row0 = 0;
    a.Wg[0] = m.P(gn + "/gates_kernel") + (long long)Y * 2 * Y; a.Wc[0] = m.P(gn + "/cand_kernel") + (long long)Y * Y;
    a.st_r = m.W(rp + "st_r"); a.st_u = m.W(rp + "st_u"); a.st_c = m.W(rp + "st_c"); a.st_hprev = m.W(rp + "st_hprev");
    a.dout = dy; a.dout_ld = Y; a.dgx = dgx;
    a.t_begin = c.t0; a.t_end = c.t1;
    a.dbg[0] = m.G(gn + "/gates_bias"); a.dbc[0] = m.G(gn + "/cand_bias");      // bias gradients: accumulated over the chunks in the kernel
    a.dh_in = last_chunk ? nullptr : carry;                              // the chunk at the end of the sequence starts from zero
    a.dh0 = (c.t0 == 0) ? (want_dh0 ? m.W(rp + "dh0") : nullptr) : carry;
    TACO_TRY(prof_launch_gru(a, true, s));
    float* gc = m.W(rp + "wv_g"); float* dc = m.W(rp + "wv_x");
    TACO_TRY(launch_copy2d(gc, dgx + (long long)c.t0 * 3 * Y, D.N, Tc * 3 * Y, (long long)Tc * 3 * Y, (long long)D.Td * 3 * Y, s));
    TACO_TRY(launch_copy2d(dc, dy + (long long)c.t0 * Y, D.N, Tc * Y, (long long)Tc * Y, (long long)D.Td * Y, s));
    taco_gemm_desc e = gd(gc, m.W(rp + "wxcat"), dc, rows, Y, 3 * Y, 3 * Y, 3 * Y, Y); e.transB = 1; e.accumulate = 1;
    TACO_TRY(launch_gemm(&e, 1, prec, s));
    TACO_TRY(launch_copy2d(dx + (long long)c.t0 * Y, dc, D.N, Tc * Y, (long long)D.Td * Y, (long long)Tc * Y, s));
    return TACO_OK;
}
// weight / bias gradients of one residual GRU layer over the whole sequence (leaves; operands complete on `producer`)
static int dec_gru_wgrads(Model& m, const DecDims& D, int layer, const float* x, cudaStream_t producer) {
    const std::string gn = "dec_gru_" + std::to_string(layer);
    const std::string rp = "dec/g" + std::to_string(layer) + "_";
    const int Y = D.Y, rows = D.N * D.Td, prec = m.cfg.precision;
    float* dgx = m.W(rp + "dgx");
    taco_gemm_desc w[4];
    w[0] = wgrad(x, Y, dgx, 3 * Y, m.G(gn + "/gates_kernel"), Y, 2 * Y, rows);
    w[1] = wgrad(x, Y, dgx + 2 * Y, 3 * Y, m.G(gn + "/cand_kernel"), Y, Y, rows);
    w[2] = wgrad(m.W(rp + "st_hprev"), Y, dgx, 3 * Y, m.G(gn + "/gates_kernel") + (long long)Y * 2 * Y, Y, 2 * Y, rows);
    w[3] = wgrad(m.W(rp + "st_r"), Y, dgx + 2 * Y, 3 * Y, m.G(gn + "/cand_kernel") + (long long)Y * Y, Y, Y, rows);
    cudaStream_t leaf = fork_side_after(producer);
    TACO_TRY(launch_gemm(w, 4, prec, leaf));
    return TACO_OK;
}

// Input: "post_cbhg/d_xin_p" (grad wrt mel outputs in the padded layout, mel-loss term already added).
// Output: "enc_cbhg/d_rnn_out" (grad wrt encoder memory) and all decoder parameter gradients.
int decoder_backward(Model& m, const taco_batch* b, cudaStream_t s) {
    const DecDims D = dec_dims(m);
    const int prec = m.cfg.precision;
    const int rows = D.N * D.Td, Y = D.Y, MR = D.M * D.r;
    const CbhgGeom& g = m.post;
    const bool deepvoice = m.has_region("spk/att_init");
    // mel projection backward
    float* d_dec = m.W("dec/d_dec");      // [N*Td, M*r] dense
    TACO_TRY(launch_unpad(d_dec, m.W("post_cbhg/d_xin_p"), D.N, D.To, g.Tp, g.PL, D.M, D.M, s));
    {
        taco_gemm_desc w = wgrad(m.W("dec/y2"), Y, d_dec, MR, m.G("mel_proj/kernel"), Y, MR, rows);
        cudaStream_t leaf = fork_side(s);
        TACO_TRY(launch_gemm(&w, 1, prec, leaf));
        TACO_TRY(launch_colsum(d_dec, m.G("mel_proj/bias"), rows, MR, MR, leaf));
        taco_gemm_desc e = gd(d_dec, m.P("mel_proj/kernel"), m.W("dec/d_y2"), rows, Y, MR, MR, MR, Y); e.transB = 1;
        TACO_TRY(launch_gemm(&e, 1, prec, s));
    }
    WaveCtx* wv = nullptr;
    {
        AttArgs probe{};
        probe.E = D.E; probe.A = D.A; probe.HA = D.HA; probe.Z1 = D.Z1; probe.Z = D.Z; probe.SPK = D.SPK; probe.Y = D.Y; probe.Ti = D.Ti;
        probe.att_type = m.cfg.attention_type; probe.fast = (prec != TACO_PREC_FP32);
        if (wave_applicable(m, probe, b)) TACO_TRY(wave_get(&wv));
    }
    if (!wv) {
        prof_mark("dec_b:gru2", s);
        TACO_TRY(dec_gru_layer_bwd(m, D, 2, m.W("dec/y1"), m.W("dec/d_y2"), m.W("dec/d_y1"), deepvoice, s));
        prof_mark("dec_b:gru1", s);
        TACO_TRY(dec_gru_layer_bwd(m, D, 1, m.W("dec/y0"), m.W("dec/d_y1"), m.W("dec/d_y0"), deepvoice, s));
        prof_mark("dec_b:attention", s);
    }

    // (the transposed weight copies the attention BPTT reads are produced by backward_prep, model.cu)
    const int ZS = D.Z + D.SPK;
    AttArgs a{};
    a.N = D.N; a.Ti = D.Ti; a.Td = D.Td; a.E = D.E; a.A = D.A; a.HA = D.HA; a.Z1 = D.Z1; a.Z = D.Z; a.SPK = D.SPK; a.Y = D.Y;
    a.att_type = m.cfg.attention_type; a.fast = (prec != TACO_PREC_FP32);
    a.memory = m.W("enc_cbhg/rnn_out"); a.keys = m.W("dec/keys");
    a.manual = b->manual_alignments;
    fill_att_weights(m, D, a);
    a.s_z1 = m.W("dec/s_z1"); a.s_z = m.W("dec/s_z"); a.s_r = m.W("dec/s_r"); a.s_u = m.W("dec/s_u"); a.s_c = m.W("dec/s_c");
    a.s_haprev = m.W("dec/s_haprev"); a.s_ha = m.W("dec/s_ha"); a.s_q = m.W("dec/s_q"); a.s_ctxin = m.W("dec/s_ctxin");
    a.s_ctx = m.W("dec/s_ctx"); a.s_e = m.W("dec/s_e"); a.s_a = m.W("dec/s_a");
    a.dy0 = m.W("dec/d_y0");
    a.W1cT = m.W("dec/W1cT"); a.W2T = m.W("dec/W2T"); a.WgT = m.W("dec/WgT"); a.WcT = m.W("dec/WcT"); a.WqT = m.W("dec/WqT"); a.WoT = m.W("dec/WoT");
    a.d_G = m.W("dec/d_G"); a.d_zp = m.W("dec/d_zp"); a.d_z1p = m.W("dec/d_z1p"); a.d_ctx = m.W("dec/d_ctx");
    a.d_gq = m.W("dec/d_gq"); a.d_ge = m.W("dec/d_ge");
    a.d_ha0 = deepvoice ? m.W("dec/d_ha0") : nullptr;
    a.d_score_bias = m.has("attention/score_bias") ? m.G("attention/score_bias") : nullptr;
    if (images_applicable(m, a) && m.img_w_state != 0 && m.img_bkm_state != 0) {
        if (m.img_bkm_state == 2) TACO_TRY(image_wait(1, s));
        a.img_w = reinterpret_cast<const uint8_t*>(m.W("dec/img_bw")); a.img_km = reinterpret_cast<const uint8_t*>(m.W("dec/img_bkm"));
    }
    if (!wv) {
        TACO_TRY(prof_launch_att(a, true, s));
    } else {
        // wavefront BPTT, last chunk first: GRU layer 2 -> GRU layer 1 -> attention, one stream each (see decoder_forward_wave)
        const std::vector<Chunk> ch = wave_chunks(D.Td, wv->chunks);
        const int C = (int)ch.size();
        for (int l = 1; l <= 2; l++)
            TACO_CHECK_CUDA(cudaMemsetAsync(m.W("dec/g" + std::to_string(l) + "_dgx"), 0, sizeof(float) * (size_t)rows * 3 * Y, s));
        TACO_CHECK_CUDA(cudaEventRecord(wv->start, s));
        TACO_CHECK_CUDA(cudaStreamWaitEvent(wv->w[1], wv->start, 0));
        TACO_CHECK_CUDA(cudaStreamWaitEvent(wv->w[0], wv->start, 0));
        a.c_dha = m.W("dec/c_dha"); a.c_dctx = m.W("dec/c_dctx"); a.c_dac = m.W("dec/c_dac");
        for (int c = C - 1; c >= 0; c--) {
            const bool last = (c == C - 1);
            TACO_TRY(dec_gru_chunk_bwd(m, D, 2, m.W("dec/d_y2"), m.W("dec/d_y1"), last, deepvoice, ch[c], wv->w[1]));
            TACO_CHECK_CUDA(cudaEventRecord(wv->ev[2][c], wv->w[1]));
            TACO_CHECK_CUDA(cudaStreamWaitEvent(wv->w[0], wv->ev[2][c], 0));
            TACO_TRY(dec_gru_chunk_bwd(m, D, 1, m.W("dec/d_y1"), m.W("dec/d_y0"), last, deepvoice, ch[c], wv->w[0]));
            TACO_CHECK_CUDA(cudaEventRecord(wv->ev[1][c], wv->w[0]));
            TACO_CHECK_CUDA(cudaStreamWaitEvent(s, wv->ev[1][c], 0));
            a.t_begin = ch[c].t0; a.t_end = ch[c].t1; a.c_in = last ? 0 : 1;
            TACO_TRY(prof_launch_att(a, true, s));
        }
        a.t_begin = 0; a.t_end = 0; a.c_in = 0;
        // parameter gradients of the two GRU layers: leaves, ready as soon as their layer's last chunk is done
        TACO_TRY(dec_gru_wgrads(m, D, 2, m.W("dec/y1"), wv->w[1]));
        TACO_TRY(dec_gru_wgrads(m, D, 1, m.W("dec/y0"), wv->w[0]));
    }

    // hoisted parameter gradients of the attention part
    {
        const int HA = D.HA, E = D.E, A = D.A, Z = D.Z, Z1 = D.Z1, M = D.M;
        std::vector<taco_gemm_desc> w;
        float* gWo = m.G("concat_proj/kernel");
        w.push_back(wgrad(a.s_ha, HA, a.dy0, Y, gWo, HA, Y, rows));
        w.push_back(wgrad(a.s_ctx, E, a.dy0, Y, gWo + (long long)HA * Y, E, Y, rows));
        w.push_back(wgrad(a.s_ha, HA, a.d_gq, A, m.G("attention/query_kernel"), HA, A, rows));
        float* gWg = m.G("attention_gru/gates_kernel"); float* gWc = m.G("attention_gru/cand_kernel");
        w.push_back(wgrad(a.s_z, Z, a.d_G, 3 * HA, gWg, Z, 2 * HA, rows));
        w.push_back(wgrad(a.s_haprev, HA, a.d_G, 3 * HA, gWg + (long long)ZS * 2 * HA, HA, 2 * HA, rows));
        w.push_back(wgrad(a.s_z, Z, a.d_G + 2 * HA, 3 * HA, gWc, Z, HA, rows));
        w.push_back(wgrad(a.s_r, HA, a.d_G + 2 * HA, 3 * HA, gWc + (long long)ZS * HA, HA, HA, rows));
        w.push_back(wgrad(a.s_z1, Z1, a.d_zp, Z, m.G("dec_prenet/dense_2/kernel"), Z1, Z, rows));
        float* gW1 = m.G("dec_prenet/dense_1/kernel");
        w.push_back(wgrad(m.W("dec/x_all"), M, a.d_z1p, Z1, gW1, M, Z1, rows));
        w.push_back(wgrad(a.s_ctxin, E, a.d_z1p, Z1, gW1 + (long long)M * Z1, E, Z1, rows));
        cudaStream_t leaf = fork_side(s);
        TACO_TRY(launch_gemm(w.data(), (int)w.size(), prec, leaf));
        TACO_TRY(launch_colsum(a.dy0, m.G("concat_proj/bias"), rows, Y, Y, leaf));
        TACO_TRY(launch_colsum(a.d_G, m.G("attention_gru/gates_bias"), rows, 2 * HA, 3 * HA, leaf));
        TACO_TRY(launch_colsum(a.d_G + 2 * HA, m.G("attention_gru/cand_bias"), rows, HA, 3 * HA, leaf));
        TACO_TRY(launch_colsum(a.d_zp, m.G("dec_prenet/dense_2/bias"), rows, Z, Z, leaf));
        TACO_TRY(launch_colsum(a.d_z1p, m.G("dec_prenet/dense_1/bias"), rows, Z1, Z1, leaf));
        if (m.cfg.attention_type == TACO_ATT_BAH_NORM) TACO_TRY(launch_colsum(a.d_gq, m.G("attention/b"), rows, A, A, leaf));
        if (D.SPK) {
            // 'simple' speaker rows (rnn_wrappers.py:372-376,408-413): the embedding is constant over time, so its kernel rows
            // and its own gradient need only the time sums of the pre-activation gradients
            const int S = D.SPK, Nb = D.N;
            float* sy = m.W("spk/s_y0"); float* sG = m.W("spk/s_G"); float* de = m.W("spk/d_embed");
            TACO_CHECK_CUDA(cudaMemsetAsync(sy, 0, sizeof(float) * (size_t)Nb * Y, s));
            TACO_CHECK_CUDA(cudaMemsetAsync(sG, 0, sizeof(float) * (size_t)Nb * 3 * HA, s));
            TACO_TRY(launch_timesum(a.dy0, sy, Nb, D.Td, D.Td, 0, Y, s));
            TACO_TRY(launch_timesum(a.d_G, sG, Nb, D.Td, D.Td, 0, 3 * HA, s));
            const float* emb = m.W("spk/embed");
            struct Term { const float* sum; int ld; int n; const float* W; float* gW; };
            const Term terms[3] = {
                {sy, Y, Y, m.P("concat_proj/kernel") + (long long)(HA + E) * Y, gWo + (long long)(HA + E) * Y},
                {sG, 3 * HA, 2 * HA, m.P("attention_gru/gates_kernel") + (long long)Z * 2 * HA, gWg + (long long)Z * 2 * HA},
                {sG + 2 * HA, 3 * HA, HA, m.P("attention_gru/cand_kernel") + (long long)Z * HA, gWc + (long long)Z * HA}};
            for (const Term& t : terms) {
                taco_gemm_desc ws = gd(emb, t.sum, t.gW, S, t.n, Nb, S, t.ld, t.n); ws.transA = 1; ws.accumulate = 1;
                TACO_TRY(launch_gemm(&ws, 1, TACO_PREC_FP32, s));
                taco_gemm_desc es = gd(t.sum, t.W, de, Nb, S, t.n, t.ld, t.n, S); es.transB = 1; es.accumulate = 1;
                TACO_TRY(launch_gemm(&es, 1, TACO_PREC_FP32, s));
            }
        }
    }
    prof_mark("dec_b:keys", s);
    // keys / v gradients and the gradient wrt the encoder memory
    {
        const int HA = D.HA, E = D.E, A = D.A;
        (void)HA;
        float* d_mem = m.W("enc_cbhg/d_rnn_out");
        // d_memory[n] = a_seq[n]^T . d_ctx[n] needs only what the attention BPTT left behind, the keys path below needs the
        // score gradients: with the wavefront streams at hand the two run side by side and the keys GEMM accumulates on top
        std::vector<taco_gemm_desc> pm;
        for (int n = 0; n < D.N; n++) {
            taco_gemm_desc d = gd(a.s_a + (long long)n * D.Td * D.Ti, a.d_ctx + (long long)n * D.Td * E, d_mem + (long long)n * D.Ti * E,
                                  D.Ti, E, D.Td, D.Ti, E, E);
            d.transA = 1; d.accumulate = 1;
            pm.push_back(d);
        }
        const bool pm_first = wv != nullptr && !b->manual_alignments;
        if (pm_first) {
            for (taco_gemm_desc& d : pm) d.accumulate = 0;
            TACO_CHECK_CUDA(cudaEventRecord(wv->start, s));
            TACO_CHECK_CUDA(cudaStreamWaitEvent(wv->w[0], wv->start, 0));
            TACO_TRY(launch_gemm(pm.data(), (int)pm.size(), prec, wv->w[0]));
            TACO_CHECK_CUDA(cudaEventRecord(wv->ev[1][0], wv->w[0]));
        }
        if (!b->manual_alignments) {
            if (m.cfg.attention_type == TACO_ATT_BAH_NORM) {
                // e = sum_u (g v_u/||v||) tanh(k + q + b): keys see the normalised vector; its gradient chains to v and g
                float* ve = m.W("dec/v_eff"); float* gve = m.W("dec/g_veff");
                TACO_TRY(launch_att_vnorm(m.P("attention/v"), m.P("attention/g"), ve, nullptr, nullptr, nullptr, A, s));
                TACO_CHECK_CUDA(cudaMemsetAsync(gve, 0, sizeof(float) * A, s));
                TACO_TRY(launch_att_keys_bwd(a.keys, a.s_q, a.d_ge, ve, m.W("dec/d_keys"), gve, D.N, D.Ti, D.Td, A, a.fast, s));
                TACO_TRY(launch_att_vnorm(m.P("attention/v"), m.P("attention/g"), nullptr, gve, m.G("attention/v"), m.G("attention/g"), A, s));
            } else {
                TACO_TRY(launch_att_keys_bwd(a.keys, a.s_q, a.d_ge, m.P("attention/v"), m.W("dec/d_keys"), m.G("attention/v"),
                                             D.N, D.Ti, D.Td, A, a.fast, s));
            }
            taco_gemm_desc w = wgrad(a.memory, E, m.W("dec/d_keys"), A, m.G("attention/memory_kernel"), E, A, (long long)D.N * D.Ti);
            TACO_TRY(launch_gemm(&w, 1, prec, fork_side(s)));
            taco_gemm_desc e = gd(m.W("dec/d_keys"), m.P("attention/memory_kernel"), d_mem, D.N * D.Ti, E, A, A, A, E); e.transB = 1;
            if (pm_first) { TACO_CHECK_CUDA(cudaStreamWaitEvent(s, wv->ev[1][0], 0)); e.accumulate = 1; }
            TACO_TRY(launch_gemm(&e, 1, prec, s));
        } else {
            TACO_CHECK_CUDA(cudaMemsetAsync(d_mem, 0, sizeof(float) * (size_t)D.N * D.Ti * E, s));
        }
        // d_memory[n] += a_seq[n]^T . d_ctx[n]
        if (!pm_first) TACO_TRY(launch_gemm(pm.data(), (int)pm.size(), prec, s));
    }
    return TACO_OK;
}

int debug_wave_chunks(int Td, int want, int* bounds, int cap) {
    const std::vector<Chunk> ch = wave_chunks(Td, want);
    int n = 0;
    for (const Chunk& c : ch) { if (2 * n + 1 < cap) { bounds[2 * n] = c.t0; bounds[2 * n + 1] = c.t1; } n++; }
    return n;
}

}  // namespace taco

// Debug hook (not part of the ABI header, like the other taco_debug_* entries): the time-chunk boundaries the decoder
// wavefront would use for Td decoder steps and a requested chunk count; returns the number of chunks, writes (t0, t1) pairs.
extern "C" int taco_debug_wave_chunks(int32_t Td, int32_t want, int32_t* bounds, int32_t cap) { return taco::debug_wave_chunks(Td, want, bounds, cap); }
