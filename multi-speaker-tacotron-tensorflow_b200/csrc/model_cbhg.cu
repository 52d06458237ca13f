// CBHG block (conv bank -> max-pool -> 2 projections -> residual -> highway x depth -> bi-GRU), forward and
// backward, sequenced over the GEMM / glue / GRU kernels.      reference: models/modules.py:27-96
//
// Every [N,T,C] activation lives in a zero-padded time layout [N, Tp = T + Kb - 1, C] (PL = (Kb-1)/2 zero frames in
// front, the rest behind) so that each 'same' convolution of width k is ONE implicit GEMM over all N*Tp rows with
// tap addressing (A row m reads input rows m-l_k .. m-l_k+k-1), pad rows being forced to zero in the epilogue.
#include "model.h"

namespace taco {

static taco_gemm_desc gemm_desc(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc) {
    taco_gemm_desc d{};
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
    d.alpha = 1.f; d.split_k = 1;
    return d;
}
static void set_mask(taco_gemm_desc& d, const CbhgGeom& g) { d.mask_period = g.Tp; d.mask_lo = g.PL; d.mask_hi = g.PL + g.T; }

// split-K factor for weight-gradient GEMMs (reduction over `rows`): enough CTAs to fill the machine
static int wgrad_split(int M, int N, long long rows) {
    long long tiles = (long long)cdiv(M, 128) * cdiv(N, 128);      // 128x128 tensor-core tiles, two CTAs per SM
    long long want = (2 * 148 + tiles - 1) / tiles;
    long long maxs = rows / 256 > 0 ? rows / 256 : 1;
    long long sp = want < maxs ? want : maxs;
    if (sp < 1) sp = 1;
    if (sp > 64) sp = 64;
    return (int)sp;
}

int cbhg_forward(Model& m, const CbhgGeom& g, const int* lengths, const float* before_highway, const float* rnn_h0,
                 int training, cudaStream_t s) {
    const std::string px = g.prefix + "/";
    auto R = [&](const char* n) { return m.W(px + n); };
    const int prec = m.cfg.precision;
    const int KC = g.Kb * g.Cb, rows = g.rows, H = g.H;
    const double count = (double)g.N * g.T;
    float* xin_p = R("xin_p");
    // bf16 mode: R16 / P16 yield the bf16 mirrors (nullptr in the other modes, which turns every mirror argument off)
    const bool h16 = m.use16();
    auto R16 = [&](const char* n) { return m.W16(px + n); };
    auto off16 = [](void* p, long long elems) -> void* { return p ? static_cast<void*>(static_cast<uint16_t*>(p) + elems) : nullptr; };

    // ---- conv bank: Kb implicit GEMMs, ReLU, masked, column statistics (modules.py:35-44,123-131) ----
    double* bstats = m.Wd(px + "bank_stats");
    if (training) TACO_CHECK_CUDA(cudaMemsetAsync(bstats, 0, sizeof(double) * 2 * KC, s));
    {
        std::vector<taco_gemm_desc> ds;
        for (int k = 1; k <= g.Kb; k++) {
            const int l = (k - 1) / 2;
            const std::string b = px + "bank_" + std::to_string(k);
            // (bf16 mode: the [rows, Kb*Cb] bank output - the widest tensor of the block - exists as bf16 only; the batch-norm
            //  statistics are still taken from the fp32 accumulators in the GEMM epilogue)
            taco_gemm_desc d = gemm_desc(xin_p - (long long)l * g.Cin, m.P(b + "/kernel"), h16 ? nullptr : R("bank_raw") + (k - 1) * g.Cb,
                                         rows, g.Cb, k * g.Cin, g.Cin, g.Cb, KC);
            d.C16 = off16(R16("bank_raw"), (k - 1) * g.Cb);
            d.ctap = g.Cin; d.bias = m.P(b + "/bias"); d.act = ACT_RELU; set_mask(d, g);
            d.A16 = off16(R16("xin_p"), -(long long)l * g.Cin); d.B16 = m.P16(b + "/kernel");
            if (training) { d.colsum = bstats + (k - 1) * g.Cb; d.colsumsq = bstats + KC + (k - 1) * g.Cb; }
            ds.push_back(d);
        }
        TACO_TRY(launch_gemm(ds.data(), (int)ds.size(), prec, s));
    }
    TACO_TRY(launch_bn_finalize(bstats, bstats + KC, count, R("bank_mean"), R("bank_rstd"), R("bank_var"),
                                m.P(px + "bank_1/moving_mean"), m.P(px + "bank_1/moving_var"), KC, training, s));
    // BN + max-pool(2,1,'same')  (modules.py:47-51)
    TACO_TRY(launch_bn_apply(h16 ? nullptr : R("bank_raw"), R("bank_mean"), R("bank_rstd"), m.P(px + "bank_1/gamma"), m.P(px + "bank_1/beta"),
                             nullptr, nullptr, h16 ? nullptr : R("pooled_p"), g.N, g.T, g.Tp, g.PL, KC, 1, s, R16("pooled_p"), R16("bank_raw")));

    prof_mark("cbhg_f:proj", s);
    // ---- projection 1: conv k=pw, ReLU, BN (modules.py:54-59) ----
    const int lp = (g.pw - 1) / 2;
    double* p1s = m.Wd(px + "p1_stats");
    if (training) TACO_CHECK_CUDA(cudaMemsetAsync(p1s, 0, sizeof(double) * 2 * g.P1, s));
    {
        taco_gemm_desc d = gemm_desc(h16 ? nullptr : R("pooled_p") - (long long)lp * KC, m.P(px + "proj_1/kernel"), R("p1_raw"),
                                     rows, g.P1, g.pw * KC, KC, g.P1, g.P1);
        d.ctap = KC; d.bias = m.P(px + "proj_1/bias"); d.act = ACT_RELU; set_mask(d, g);
        d.A16 = off16(R16("pooled_p"), -(long long)lp * KC); d.B16 = m.P16(px + "proj_1/kernel");
        if (training) { d.colsum = p1s; d.colsumsq = p1s + g.P1; }
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    TACO_TRY(launch_bn_finalize(p1s, p1s + g.P1, count, R("p1_mean"), R("p1_rstd"), R("p1_var"),
                                m.P(px + "proj_1/moving_mean"), m.P(px + "proj_1/moving_var"), g.P1, training, s));
    TACO_TRY(launch_bn_apply(R("p1_raw"), R("p1_mean"), R("p1_rstd"), m.P(px + "proj_1/gamma"), m.P(px + "proj_1/beta"),
                             nullptr, nullptr, h16 ? nullptr : R("p1_p"), g.N, g.T, g.Tp, g.PL, g.P1, 0, s, R16("p1_p")));

    // ---- projection 2: conv k=pw, linear, BN; + residual (+ before_highway) (modules.py:54-69) ----
    double* p2s = m.Wd(px + "p2_stats");
    if (training) TACO_CHECK_CUDA(cudaMemsetAsync(p2s, 0, sizeof(double) * 2 * g.P2, s));
    {
        taco_gemm_desc d = gemm_desc(h16 ? nullptr : R("p1_p") - (long long)lp * g.P1, m.P(px + "proj_2/kernel"), R("p2_raw"),
                                     rows, g.P2, g.pw * g.P1, g.P1, g.P2, g.P2);
        d.ctap = g.P1; d.bias = m.P(px + "proj_2/bias"); d.act = ACT_NONE; set_mask(d, g);
        d.A16 = off16(R16("p1_p"), -(long long)lp * g.P1); d.B16 = m.P16(px + "proj_2/kernel");
        if (training) { d.colsum = p2s; d.colsumsq = p2s + g.P2; }
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    TACO_TRY(launch_bn_finalize(p2s, p2s + g.P2, count, R("p2_mean"), R("p2_rstd"), R("p2_var"),
                                m.P(px + "proj_2/moving_mean"), m.P(px + "proj_2/moving_var"), g.P2, training, s));
    TACO_TRY(launch_bn_apply(R("p2_raw"), R("p2_mean"), R("p2_rstd"), m.P(px + "proj_2/gamma"), m.P(px + "proj_2/beta"),
                             xin_p, before_highway, R("hw0"), g.N, g.T, g.Tp, g.PL, g.P2, 0, s, R16("hw0")));

    prof_mark("cbhg_f:highway", s);
    // ---- dimension fix (modules.py:72-73) ----
    float* hw = R("hw0");
    void* hw16 = R16("hw0");
    if (g.has_hin) {
        taco_gemm_desc d = gemm_desc(R("hw0"), m.P(px + "highway_in/kernel"), R("hw_0"), rows, H, g.P2, g.P2, H, H);
        d.bias = m.P(px + "highway_in/bias"); set_mask(d, g);
        d.A16 = hw16; d.B16 = m.P16(px + "highway_in/kernel"); d.C16 = R16("hw_0");
        TACO_TRY(launch_gemm(&d, 1, prec, s));
        hw = R("hw_0"); hw16 = R16("hw_0");
    }
    // ---- highway stack (modules.py:76-77,105-120) ----
    for (int i = 1; i <= g.depth; i++) {
        const std::string hn = px + "highway_" + std::to_string(i);
        float* Hb = m.W(px + "hwH_" + std::to_string(i));
        float* Tb = m.W(px + "hwT_" + std::to_string(i));
        float* out = m.W(px + "hw_" + std::to_string(i));
        taco_gemm_desc d[2];
        d[0] = gemm_desc(hw, m.P(hn + "/H_kernel"), Hb, rows, H, H, H, H, H); d[0].bias = m.P(hn + "/H_bias"); d[0].act = ACT_RELU;
        d[1] = gemm_desc(hw, m.P(hn + "/T_kernel"), Tb, rows, H, H, H, H, H); d[1].bias = m.P(hn + "/T_bias"); d[1].act = ACT_SIGMOID;
        d[0].A16 = hw16; d[0].B16 = m.P16(hn + "/H_kernel"); d[1].A16 = hw16; d[1].B16 = m.P16(hn + "/T_kernel");
        TACO_TRY(launch_gemm(d, 2, prec, s));
        void* out16 = m.W16(px + "hw_" + std::to_string(i));
        TACO_TRY(launch_highway_fwd(Hb, Tb, hw, out, (long long)rows * H, s, out16));
        hw = out; hw16 = out16;
    }
    prof_mark("cbhg_f:gru", s);
    // ---- bi-GRU: hoisted x-side GEMMs, then the cluster-persistent recurrence (modules.py:82-96) ----
    {
        taco_gemm_desc d[4];
        const char* dirs[2] = {"gru_fw", "gru_bw"};
        for (int dd = 0; dd < 2; dd++) {
            const std::string gn = px + dirs[dd];
            d[2 * dd] = gemm_desc(hw, m.P(gn + "/gates_kernel"), R("gx") + dd * 3 * H, rows, 2 * H, H, H, 2 * H, 6 * H);
            d[2 * dd].bias = m.P(gn + "/gates_bias");
            d[2 * dd + 1] = gemm_desc(hw, m.P(gn + "/cand_kernel"), R("gx") + dd * 3 * H + 2 * H, rows, H, H, H, H, 6 * H);
            d[2 * dd + 1].bias = m.P(gn + "/cand_bias");
            d[2 * dd].A16 = hw16; d[2 * dd].B16 = m.P16(gn + "/gates_kernel"); d[2 * dd + 1].A16 = hw16; d[2 * dd + 1].B16 = m.P16(gn + "/cand_kernel");
        }
        TACO_TRY(launch_gemm(d, 4, prec, s));
        GruArgs a{};
        a.N = g.N; a.T = g.T; a.H = H; a.ndir = 2; a.fast = (m.cfg.precision != TACO_PREC_FP32);
        a.gx = R("gx"); a.gx_ld = 6 * H; a.gx_rs_n = g.Tp; a.gx_row0 = g.PL;
        a.Wg[0] = m.P(px + "gru_fw/gates_kernel") + (long long)H * 2 * H; a.Wc[0] = m.P(px + "gru_fw/cand_kernel") + (long long)H * H;
        a.Wg[1] = m.P(px + "gru_bw/gates_kernel") + (long long)H * 2 * H; a.Wc[1] = m.P(px + "gru_bw/cand_kernel") + (long long)H * H;
        a.h0 = rnn_h0; a.lengths = lengths;
        a.out = R("rnn_out"); a.out_ld = 2 * H;
        if (lengths) TACO_CHECK_CUDA(cudaMemsetAsync(a.out, 0, sizeof(float) * (size_t)g.N * g.T * 2 * H, s));
        if (training) { a.st_r = R("st_r"); a.st_u = R("st_u"); a.st_c = R("st_c"); a.st_hprev = R("st_hprev"); a.st_hprev16 = R16("st_hprev"); }
        a.out16 = R16("rnn_out");
        if (lengths && a.out16) TACO_CHECK_CUDA(cudaMemsetAsync(a.out16, 0, 2 * (size_t)g.N * g.T * 2 * H, s));
        TACO_TRY(prof_launch_gru(a, false, s));
    }
    return TACO_OK;
}

// Backward.  Input: region "<prefix>/d_rnn_out" [N*T, 2H].  Output: "<prefix>/d_xin_p" [rows, Cin] (masked), parameter
// gradients accumulated into the flat gradient buffer, optional "<prefix>/d_before" [N,P2] and "<prefix>/d_h0" [N,2H].
int cbhg_backward(Model& m, const CbhgGeom& g, const int* lengths, bool want_dbefore, bool want_dh0, cudaStream_t s) {
    const std::string px = g.prefix + "/";
    auto R = [&](const char* n) { return m.W(px + n); };
    const int prec = m.cfg.precision;
    const int KC = g.Kb * g.Cb, rows = g.rows, H = g.H;
    float* xin_p = R("xin_p");
    float* hw_top = m.W(px + "hw_" + std::to_string(g.depth));
    const bool h16 = m.use16();
    auto R16 = [&](const char* n) { return m.W16(px + n); };
    auto off16 = [](void* p, long long elems) -> void* { return p ? static_cast<void*>(static_cast<uint16_t*>(p) + elems) : nullptr; };
    void* hw_top16 = m.W16(px + "hw_" + std::to_string(g.depth));
    const long long NT = (long long)g.N * g.T;

    // ---- bi-GRU BPTT ----
    // (bf16 mode: the recurrence writes its gate gradients as bf16 only, once in the padded layout - x-side weight gradient,
    //  data gradient - and once dense - recurrent weight gradients; rows it never visits must read as zero)
    if (h16) {
        TACO_CHECK_CUDA(cudaMemsetAsync(R16("dgx"), 0, 2 * (size_t)rows * 6 * H, s));
        if (lengths) TACO_CHECK_CUDA(cudaMemsetAsync(R16("dgx_dense"), 0, 2 * (size_t)2 * NT * 3 * H, s));
    } else {
        TACO_CHECK_CUDA(cudaMemsetAsync(R("dgx"), 0, sizeof(float) * (size_t)rows * 6 * H, s));
    }
    {
        GruArgs a{};
        a.N = g.N; a.T = g.T; a.H = H; a.ndir = 2; a.fast = (m.cfg.precision != TACO_PREC_FP32);
        a.gx = R("gx"); a.gx_ld = 6 * H; a.gx_rs_n = g.Tp; a.gx_row0 = g.PL;
        a.Wg[0] = m.P(px + "gru_fw/gates_kernel") + (long long)H * 2 * H; a.Wc[0] = m.P(px + "gru_fw/cand_kernel") + (long long)H * H;
        a.Wg[1] = m.P(px + "gru_bw/gates_kernel") + (long long)H * 2 * H; a.Wc[1] = m.P(px + "gru_bw/cand_kernel") + (long long)H * H;
        a.lengths = lengths;
        a.st_r = R("st_r"); a.st_u = R("st_u"); a.st_c = R("st_c"); a.st_hprev = R("st_hprev");
        a.dout = R("d_rnn_out"); a.dout_ld = 2 * H; a.dgx = h16 ? nullptr : R("dgx");
        a.dgx16 = R16("dgx"); a.dgx16_dense = R16("dgx_dense"); a.st_rh16 = R16("st_rh");
        if (a.fast) {      // the fast kernels accumulate the gate / candidate bias gradients themselves
            a.dbg[0] = m.G(px + "gru_fw/gates_bias"); a.dbc[0] = m.G(px + "gru_fw/cand_bias");
            a.dbg[1] = m.G(px + "gru_bw/gates_bias"); a.dbc[1] = m.G(px + "gru_bw/cand_bias");
        }
        a.dh0 = want_dh0 ? R("d_h0") : nullptr;
        TACO_TRY(prof_launch_gru(a, true, s));
    }
    prof_mark("cbhg_b:gru_dgrad", s);
    // GRU weight gradients.  Recurrent parts reduce over the unpadded [N*T] stash rows; dgx lives in the padded
    // layout, so gather its valid rows once into a dense [2][N*T, 3H] matrix first.
    cudaStream_t leaf = fork_side(s);       // weight / bias gradients are leaves: they run beside the chain (model.cu)
    float* dgd = h16 ? nullptr : R("dgx_dense");
    if (!h16)
        for (int dd = 0; dd < 2; dd++)
            TACO_TRY(launch_unpad(dgd + (long long)dd * g.N * g.T * 3 * H, R("dgx") + dd * 3 * H, g.N, g.T, g.Tp, g.PL, 3 * H, 6 * H, leaf));
    {
        const char* dirs[2] = {"gru_fw", "gru_bw"};
        std::vector<taco_gemm_desc> ds;
        float* dgx = h16 ? nullptr : R("dgx");
        void* dgx16 = R16("dgx"); void* dgd16 = R16("dgx_dense");
        for (int dd = 0; dd < 2; dd++) {
            const std::string gn = px + dirs[dd];
            const float* dG = h16 ? nullptr : dgd + (long long)dd * NT * 3 * H;
            void* dG16 = off16(dgd16, (long long)dd * NT * 3 * H);
            float* gWg = m.G(gn + "/gates_kernel"); float* gWc = m.G(gn + "/cand_kernel");
            // x-side: dW[0:H] += hw_top^T . dgx (padded rows; pad rows of dgx are zero)
            taco_gemm_desc d = gemm_desc(hw_top, h16 ? nullptr : dgx + dd * 3 * H, gWg, H, 2 * H, rows, H, 6 * H, 2 * H);
            d.A16 = hw_top16; d.B16 = off16(dgx16, dd * 3 * H);
            d.transA = 1; d.accumulate = 1; d.split_k = wgrad_split(H, 2 * H, rows); ds.push_back(d);
            d = gemm_desc(hw_top, h16 ? nullptr : dgx + dd * 3 * H + 2 * H, gWc, H, H, rows, H, 6 * H, H);
            d.A16 = hw_top16; d.B16 = off16(dgx16, dd * 3 * H + 2 * H);
            d.transA = 1; d.accumulate = 1; d.split_k = wgrad_split(H, H, rows); ds.push_back(d);
            // h-side: dWg[H:2H] += hprev^T . [dr|du] ; dWc[H:2H] += (r*hprev)^T . dc
            d = gemm_desc(R("st_hprev") + (long long)dd * NT * H, dG, gWg + (long long)H * 2 * H, H, 2 * H, (int)NT, H, 3 * H, 2 * H);
            d.A16 = off16(R16("st_hprev"), (long long)dd * NT * H); d.B16 = dG16;
            d.transA = 1; d.accumulate = 1; d.split_k = wgrad_split(H, 2 * H, NT); ds.push_back(d);
            d = gemm_desc(R("st_r") + (long long)dd * NT * H, h16 ? nullptr : dG + 2 * H, gWc + (long long)H * H, H, H, (int)NT, H, 3 * H, H);
            d.A16 = off16(R16("st_rh"), (long long)dd * NT * H); d.B16 = off16(dG16, 2 * H);
            d.transA = 1; d.accumulate = 1; d.split_k = wgrad_split(H, H, NT); ds.push_back(d);
        }
        TACO_TRY(launch_gemm(ds.data(), (int)ds.size(), prec, leaf));
        for (int dd = 0; dd < 2 && prec == TACO_PREC_FP32; dd++) {       // (exact kernels: bias gradients as column sums of dgx)
            const std::string gn = px + dirs[dd];
            if (h16) {
                TACO_TRY(launch_colsum16(off16(dgx16, dd * 3 * H), m.G(gn + "/gates_bias"), rows, 2 * H, 6 * H, leaf));
                TACO_TRY(launch_colsum16(off16(dgx16, dd * 3 * H + 2 * H), m.G(gn + "/cand_bias"), rows, H, 6 * H, leaf));
            } else {
                TACO_TRY(launch_colsum(dgx + dd * 3 * H, m.G(gn + "/gates_bias"), rows, 2 * H, 6 * H, leaf));
                TACO_TRY(launch_colsum(dgx + dd * 3 * H + 2 * H, m.G(gn + "/cand_bias"), rows, H, 6 * H, leaf));
            }
        }
    }
    {
        // d hw_top = dgx . Wx^T over all four blocks at once: K = 6H against the packed [H, 6H] x-side weights (backward_prep)
        taco_gemm_desc d = gemm_desc(h16 ? nullptr : R("dgx"), R("gru_wxcat"), R("d_hwA"), rows, H, 6 * H, 6 * H, 6 * H, H);
        d.A16 = R16("dgx"); d.B16 = R16("gru_wxcat");
        d.transB = 1;
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    prof_mark("cbhg_b:highway", s);
    // ---- highway stack backward ----
    float* dcur = R("d_hwA");
    float* dnext = R("d_hwB");
    for (int i = g.depth; i >= 1; i--) {
        const std::string hn = px + "highway_" + std::to_string(i);
        float* Hb = m.W(px + "hwH_" + std::to_string(i));
        float* Tb = m.W(px + "hwT_" + std::to_string(i));
        float* xin = (i == 1) ? (g.has_hin ? R("hw_0") : R("hw0")) : m.W(px + "hw_" + std::to_string(i - 1));
        // pre-activation gradients (dHpre | dTpre) get a buffer per layer: the leaf stream may still be reading layer i's
        // while the chain already produces layer i-1's
        float* dHT = h16 ? nullptr : m.W(px + "d_HT_" + std::to_string(i));
        void* dHT16 = m.W16(px + "d_HT_" + std::to_string(i));
        void* xin16 = (i == 1) ? (g.has_hin ? R16("hw_0") : R16("hw0")) : m.W16(px + "hw_" + std::to_string(i - 1));
        TACO_TRY(launch_highway_bwd(dcur, Hb, Tb, xin, dHT, dnext, rows, H, s, dHT16, m.G(hn + "/H_bias"), m.G(hn + "/T_bias")));    // bias gradients ride along
        cudaStream_t lf = fork_side(s);
        taco_gemm_desc d[2];
        d[0] = gemm_desc(xin, dHT, m.G(hn + "/H_kernel"), H, H, rows, H, 2 * H, H); d[0].transA = 1; d[0].accumulate = 1; d[0].split_k = wgrad_split(H, H, rows);
        d[1] = gemm_desc(xin, h16 ? nullptr : dHT + H, m.G(hn + "/T_kernel"), H, H, rows, H, 2 * H, H); d[1].transA = 1; d[1].accumulate = 1; d[1].split_k = d[0].split_k;
        d[0].A16 = xin16; d[0].B16 = dHT16; d[1].A16 = xin16; d[1].B16 = off16(dHT16, H);
        TACO_TRY(launch_gemm(d, 2, prec, lf));
        // dx += dHpre.WH^T + dTpre.WT^T: one GEMM with K = 2H against the packed [H, 2H] weight
        taco_gemm_desc e = gemm_desc(dHT, m.W(px + "hw_wcat_" + std::to_string(i)), dnext, rows, H, 2 * H, 2 * H, 2 * H, H);
        e.A16 = dHT16; e.B16 = m.W16(px + "hw_wcat_" + std::to_string(i));
        e.transB = 1; e.accumulate = 1;
        TACO_TRY(launch_gemm(&e, 1, prec, s));
        std::swap(dcur, dnext);
    }
    // ---- dimension fix backward ----
    float* d_hw0 = dcur;   // grad wrt hw0 [rows,P2] when no dense; else grad wrt hw_0 [rows,H]
    if (g.has_hin) {
        taco_gemm_desc d = gemm_desc(R("hw0"), dcur, m.G(px + "highway_in/kernel"), g.P2, H, rows, g.P2, H, H);
        d.transA = 1; d.accumulate = 1; d.split_k = wgrad_split(g.P2, H, rows);      // (dcur has no bf16 mirror: fp32-operand kernel)
        cudaStream_t lf = fork_side(s);
        TACO_TRY(launch_gemm(&d, 1, prec, lf));
        TACO_TRY(launch_colsum(dcur, m.G(px + "highway_in/bias"), rows, H, H, lf));
        taco_gemm_desc e = gemm_desc(dcur, m.P(px + "highway_in/kernel"), R("d_hw0"), rows, g.P2, H, H, H, g.P2);
        e.transB = 1; set_mask(e, g);
        TACO_TRY(launch_gemm(&e, 1, prec, s));
        d_hw0 = R("d_hw0");
    }
    if (want_dbefore) {
        TACO_CHECK_CUDA(cudaMemsetAsync(R("d_before"), 0, sizeof(float) * (size_t)g.N * g.P2, s));
        TACO_TRY(launch_timesum(d_hw0, R("d_before"), g.N, g.T, g.Tp, g.PL, g.P2, s));
    }
    prof_mark("cbhg_b:proj2", s);
    // ---- projection 2 backward: BN (no activation), conv ----
    const int lp = (g.pw - 1) / 2, rp = g.pw - 1 - lp;
    TACO_TRY(launch_bn_bwd(d_hw0, R("p2_raw"), R("p2_mean"), R("p2_rstd"), m.P(px + "proj_2/gamma"), m.P(px + "proj_2/beta"),
                           m.G(px + "proj_2/gamma"), m.G(px + "proj_2/beta"), h16 ? nullptr : R("d_p2raw"), g.N, g.T, g.Tp, g.PL, g.P2, 0, 0, s,
                           R16("d_p2raw"), m.G(px + "proj_2/bias")));      // the convolution's bias gradient rides in the BN backward pass
    {
        taco_gemm_desc d = gemm_desc(h16 ? nullptr : R("p1_p") - (long long)lp * g.P1, h16 ? nullptr : R("d_p2raw"), m.G(px + "proj_2/kernel"),
                                     g.pw * g.P1, g.P2, rows, g.P1, g.P2, g.P2);
        d.A16 = off16(R16("p1_p"), -(long long)lp * g.P1); d.B16 = R16("d_p2raw");
        d.transA = 1; d.ctap = g.P1; d.accumulate = 1; d.split_k = wgrad_split(g.pw * g.P1, g.P2, rows);
        cudaStream_t lf = fork_side(s);
        TACO_TRY(launch_gemm(&d, 1, prec, lf));
        // dgrad: d p1_p[s] = sum_j' d_p2raw[s - rp + j'] . Wd[j']  with Wd = flipped+transposed kernel (packed)
        taco_gemm_desc e = gemm_desc(h16 ? nullptr : R("d_p2raw") - (long long)rp * g.P2, m.W(px + "proj_2/wd"), R("d_p1p"),
                                     rows, g.P1, g.pw * g.P2, g.P2, g.P1, g.P1);
        e.A16 = off16(R16("d_p2raw"), -(long long)rp * g.P2); e.B16 = R16("proj_2/wd");
        e.ctap = g.P2; set_mask(e, g);
        TACO_TRY(launch_gemm(&e, 1, prec, s));
    }
    prof_mark("cbhg_b:proj1", s);
    // ---- projection 1 backward: BN + ReLU, conv ----
    TACO_TRY(launch_bn_bwd(R("d_p1p"), R("p1_raw"), R("p1_mean"), R("p1_rstd"), m.P(px + "proj_1/gamma"), m.P(px + "proj_1/beta"),
                           m.G(px + "proj_1/gamma"), m.G(px + "proj_1/beta"), h16 ? nullptr : R("d_p1raw"), g.N, g.T, g.Tp, g.PL, g.P1, 0, 1, s,
                           R16("d_p1raw"), m.G(px + "proj_1/bias")));
    {
        taco_gemm_desc d = gemm_desc(h16 ? nullptr : R("pooled_p") - (long long)lp * KC, h16 ? nullptr : R("d_p1raw"), m.G(px + "proj_1/kernel"),
                                     g.pw * KC, g.P1, rows, KC, g.P1, g.P1);
        d.A16 = off16(R16("pooled_p"), -(long long)lp * KC); d.B16 = R16("d_p1raw");
        d.transA = 1; d.ctap = KC; d.accumulate = 1; d.split_k = wgrad_split(g.pw * KC, g.P1, rows);
        cudaStream_t lf = fork_side(s);
        TACO_TRY(launch_gemm(&d, 1, prec, lf));
        taco_gemm_desc e = gemm_desc(h16 ? nullptr : R("d_p1raw") - (long long)rp * g.P1, m.W(px + "proj_1/wd"), h16 ? nullptr : R("d_pooled"),
                                     rows, KC, g.pw * g.P1, g.P1, KC, KC);
        e.C16 = R16("d_pooled");
        e.A16 = off16(R16("d_p1raw"), -(long long)rp * g.P1); e.B16 = R16("proj_1/wd");
        e.ctap = g.P1; set_mask(e, g);
        TACO_TRY(launch_gemm(&e, 1, prec, s));
    }
    prof_mark("cbhg_b:bank", s);
    // ---- max-pool + BN + ReLU backward of the bank ----
    TACO_TRY(launch_bn_bwd(h16 ? nullptr : R("d_pooled"), h16 ? nullptr : R("bank_raw"), R("bank_mean"), R("bank_rstd"), m.P(px + "bank_1/gamma"), m.P(px + "bank_1/beta"),
                           m.G(px + "bank_1/gamma"), m.G(px + "bank_1/beta"), h16 ? nullptr : R("d_bank"), g.N, g.T, g.Tp, g.PL, KC, 1, 1, s,
                           R16("d_bank"), m.G(px + "bank_1/bias"), R16("bank_raw"), R16("d_pooled")));
    cudaStream_t lfb = fork_side(s);
    {
        std::vector<taco_gemm_desc> wg, dg;
        // residual path first: d_xin_p = d_hw0 (masked already: d_hw0 is zero on pad rows)
        TACO_TRY(launch_copy2d(R("d_xin_p"), d_hw0, rows, g.Cin, g.Cin, g.P2, s));
        const bool merged = h16 ? (g.Cb % 64 == 0) : ((prec != TACO_PREC_FP32) && (g.Cb % 32 == 0) && (KC % 4 == 0));
        float* d_bank = h16 ? nullptr : R("d_bank");
        void* d_bank16 = R16("d_bank");
        for (int k = 1; k <= g.Kb; k++) {
            const int l = (k - 1) / 2, r = k - 1 - l;
            const std::string b = px + "bank_" + std::to_string(k);
            taco_gemm_desc d = gemm_desc(xin_p - (long long)l * g.Cin, h16 ? nullptr : d_bank + (k - 1) * g.Cb, m.G(b + "/kernel"),
                                         k * g.Cin, g.Cb, rows, g.Cin, KC, g.Cb);
            d.A16 = off16(R16("xin_p"), -(long long)l * g.Cin); d.B16 = off16(d_bank16, (k - 1) * g.Cb);
            d.transA = 1; d.ctap = g.Cin; d.accumulate = 1; d.split_k = wgrad_split(k * g.Cin, g.Cb, rows);
            wg.push_back(d);
            if (merged) continue;
            taco_gemm_desc e = gemm_desc(h16 ? nullptr : d_bank + (k - 1) * g.Cb - (long long)r * KC, bank_wd(m, g, k), R("d_xin_p"),
                                         rows, g.Cin, k * g.Cb, KC, g.Cin, g.Cin);
            e.A16 = off16(d_bank16, (k - 1) * g.Cb - (long long)r * KC);
            e.B16 = off16(R16("bank_wd"), (long long)g.Cb * g.Cin * (k - 1) * k / 2);
            e.ctap = g.Cb; e.accumulate = 2; set_mask(e, g);
            dg.push_back(e);
        }
        if (merged) {
            // all Kb members in ONE GEMM: K walks (member, tap, channel block) through the tap table; the packed kernels of
            // the members are contiguous, so B is a single [sum_k k*Cb, Cin] matrix (backward_prep)
            const int KK = g.Cb * g.Kb * (g.Kb + 1) / 2;
            taco_gemm_desc e = gemm_desc(h16 ? nullptr : d_bank - (long long)g.Kb * KC, bank_wd(m, g, 1), R("d_xin_p"), rows, g.Cin, KK, KC, g.Cin, g.Cin);
            e.A16 = off16(d_bank16, -(long long)g.Kb * KC); e.B16 = R16("bank_wd");
            e.tap_table = reinterpret_cast<const int32_t*>(R("bank_taps")); e.tap_rows = rows + 2 * g.Kb;
            e.accumulate = 1; e.split_k = 16; set_mask(e, g);
            dg.push_back(e);
        }
        TACO_TRY(launch_gemm(wg.data(), (int)wg.size(), prec, lfb));
        TACO_TRY(launch_gemm(dg.data(), (int)dg.size(), prec, s));
    }
    return TACO_OK;
}

}  // namespace taco
