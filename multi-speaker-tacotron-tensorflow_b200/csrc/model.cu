// C-ABI entry points, workspace planner and the top-level forward / backward / optimizer sequencing.
// reference: models/tacotron.py:21-336 (initialize, add_loss, add_optimizer); see include/taco_capi.h.
#include "model.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace taco {

static thread_local std::string g_err;
int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

// ---- optional per-class device timing (CUDA events on the launching stream; used by bench.py's roofline leg) ----
static bool g_prof_on = false;
struct ProfSpan { cudaEvent_t a, b; int cls; int tag[4]; double gflop; const char* name; cudaStream_t stream; };
static std::vector<ProfSpan> g_prof_spans;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
    cudaStream_t s; int idx = -1;
    ProfScope(int cls, cudaStream_t st) : s(st) {
        if (!g_prof_on) return;
        ProfSpan sp{prof_event(), prof_event(), cls, {0, 0, 0, 0}, 0.0, "", st};
        cudaEventRecord(sp.a, s);
        g_prof_spans.push_back(sp); idx = (int)g_prof_spans.size() - 1;
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(g_prof_spans[idx].b, s); }
    void tag(int a, int b, int c, int d) { if (idx >= 0) { int* t = g_prof_spans[idx].tag; t[0] = a; t[1] = b; t[2] = c; t[3] = d; } }
    void work(double gflop) { if (idx >= 0) g_prof_spans[idx].gflop = gflop; }
};
// stage marker for the timeline listing (taco_debug_profile_spans): class 9, zero length
void prof_mark(const char* name, cudaStream_t s) {
    if (!g_prof_on) return;
    ProfSpan sp{prof_event(), nullptr, 9, {0, 0, 0, 0}, 0.0, name, s};
    cudaEventRecord(sp.a, s);
    sp.b = sp.a;
    g_prof_spans.push_back(sp);
}
int prof_launch_gru(const GruArgs& a, bool bwd, cudaStream_t s) { ProfScope p(1, s); return bwd ? launch_gru_bwd(a, s) : launch_gru_fwd(a, s); }
int prof_launch_att_free(const AttArgs& a, void* img, cudaStream_t s) { ProfScope p(2, s); return launch_att_free(a, img, s); }
int prof_launch_att(const AttArgs& a, bool bwd, cudaStream_t s) { ProfScope p(2, s); return bwd ? launch_att_bwd(a, s) : launch_att_fwd(a, s); }

// ---- stream scheduler --------------------------------------------------------------------------------------
// The backward pass has one long dependent chain (loss -> post-net BPTT -> decoder BPTT -> encoder BPTT) whose recurrence
// kernels occupy at most 64 of the 148 SMs, while every weight / bias gradient is a leaf of the dependency graph.  So
// backward runs on two internal streams: `crit` (highest priority) carries the chain, `side` (lowest priority) the
// leaves; each leaf is forked behind the producer of its operands (fork_side) and all of them are joined once at the end
// of taco_backward.  Priorities make the block scheduler hand freed SMs to the chain first, so leaves only fill idle SMs.
// TACO_OVERLAP=0 (or an active taco_profile window, whose per-launch times must not overlap) runs everything in order.
static constexpr int kAuxStreams = 4;
static constexpr int kMaxDevices = 16;
struct StreamSet {
    cudaStream_t aux[kAuxStreams]; cudaEvent_t fork, join[kAuxStreams];
};
// Streams and events belong to a device: one scheduler state per CUDA ordinal, selected through the device that the C-ABI
// entry made current (DeviceGuard below), so engines on different devices in one process never share them.
struct DevSched {
    StreamSet set[2];                   // 0: chain (high priority), 1: leaves (low priority)
    cudaStream_t crit = nullptr, side = nullptr;
    cudaEvent_t ev_in, ev_out, ev_fork, ev_side, ev_prep, ev_leaf, ev_img[2], ev_early[2];     // ev_early: gradients of the early bucket complete (chain / leaves)
    bool ready = false;
    WaveCtx wave; bool wave_ready = false;
};
static DevSched g_dev[kMaxDevices];
static DevSched& ds() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
    return g_dev[d];
}
// RAII: make the model's device current for the duration of a C-ABI call (the caller's current device is restored)
struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (dev >= 0 && dev != prev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
};
#define TACO_ON_DEVICE(dev)                                                                     \
    DeviceGuard _guard(dev);                                                                    \
    TACO_REQUIRE(_guard.ok && (dev) < kMaxDevices, TACO_ECUDA, "cannot make CUDA device %d current", (int)(dev))

static int g_overlap = -1;
static bool g_prof_keep_overlap = false;     // TACO_PROF_OVERLAP=1: timeline of the real two-stream schedule (tools/timeline.py)
static int sched_init() {
    DevSched& D = ds();
    if (D.ready) return TACO_OK;
    int lo = 0, hi = 0;
    TACO_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority (numerically greatest)
    for (int k = 0; k < 2; k++) {
        for (int i = 0; i < kAuxStreams; i++) {
            TACO_CHECK_CUDA(cudaStreamCreateWithPriority(&D.set[k].aux[i], cudaStreamNonBlocking, k == 0 ? hi : lo));
            TACO_CHECK_CUDA(cudaEventCreateWithFlags(&D.set[k].join[i], cudaEventDisableTiming));
        }
        TACO_CHECK_CUDA(cudaEventCreateWithFlags(&D.set[k].fork, cudaEventDisableTiming));
    }
    TACO_CHECK_CUDA(cudaStreamCreateWithPriority(&D.crit, cudaStreamNonBlocking, hi));
    TACO_CHECK_CUDA(cudaStreamCreateWithPriority(&D.side, cudaStreamNonBlocking, lo));
    for (cudaEvent_t* e : {&D.ev_in, &D.ev_out, &D.ev_fork, &D.ev_side, &D.ev_prep, &D.ev_leaf, &D.ev_img[0], &D.ev_img[1], &D.ev_early[0], &D.ev_early[1]})
        TACO_CHECK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (g_overlap < 0) {
        const char* env = getenv("TACO_OVERLAP");
        g_overlap = (env && env[0] == '0') ? 0 : 1;
        const char* env2 = getenv("TACO_PROF_OVERLAP");
        g_prof_keep_overlap = (env2 && env2[0] == '1');
    }
    D.ready = true;
    return TACO_OK;
}
static bool overlap_on() { return ds().ready && g_overlap == 1 && (!g_prof_on || g_prof_keep_overlap); }

// Decoder wavefront: two more high-priority streams (one per residual GRU layer) and per-chunk events, so that the GRU
// layers of time chunk c run while the attention recurrence is already on chunk c+1 (and the reverse in BPTT).
static int g_wave_chunks = -1;
int wave_get(WaveCtx** out) {
    *out = nullptr;
    if (!overlap_on()) return TACO_OK;
    if (g_wave_chunks < 0) {
        const char* env = getenv("TACO_DEC_CHUNKS");
        g_wave_chunks = env ? atoi(env) : 4;
        if (g_wave_chunks > WaveCtx::kMaxChunks) g_wave_chunks = WaveCtx::kMaxChunks;
    }
    if (g_wave_chunks <= 1) return TACO_OK;
    DevSched& D = ds();
    if (!D.wave_ready) {
        int lo = 0, hi = 0;
        TACO_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int i = 0; i < 2; i++) TACO_CHECK_CUDA(cudaStreamCreateWithPriority(&D.wave.w[i], cudaStreamNonBlocking, hi));
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < WaveCtx::kMaxChunks; c++) TACO_CHECK_CUDA(cudaEventCreateWithFlags(&D.wave.ev[k][c], cudaEventDisableTiming));
        TACO_CHECK_CUDA(cudaEventCreateWithFlags(&D.wave.start, cudaEventDisableTiming));
        D.wave_ready = true;
    }
    D.wave.chunks = g_wave_chunks;
    *out = &D.wave;
    return TACO_OK;
}
// Leaf stream for work whose operands were produced on `producer` (any stream): the side stream waits for what is
// enqueued there so far.  Falls back to `producer` itself when the two-stream schedule is off.
cudaStream_t fork_side_after(cudaStream_t producer) {
    if (!overlap_on()) return producer;
    if (cudaEventRecord(ds().ev_leaf, producer) != cudaSuccess || cudaStreamWaitEvent(ds().side, ds().ev_leaf, 0) != cudaSuccess) return producer;
    return ds().side;
}
// Hand-over of the attention operand images between streams: which = 0 weight images (built beside the encoder), 1 the
// backward key / memory image (built beside the forward decoder).  image_ready records, image_wait orders `s` behind it.
int image_ready(int which, cudaStream_t producer) { TACO_CHECK_CUDA(cudaEventRecord(ds().ev_img[which], producer)); return TACO_OK; }
int image_wait(int which, cudaStream_t s) { TACO_CHECK_CUDA(cudaStreamWaitEvent(s, ds().ev_img[which], 0)); return TACO_OK; }


// Stream for leaf work whose operands were produced by what is already enqueued on `main`.
cudaStream_t fork_side(cudaStream_t main) {
    if (!overlap_on() || main != ds().crit) return main;
    if (cudaEventRecord(ds().ev_fork, main) != cudaSuccess || cudaStreamWaitEvent(ds().side, ds().ev_fork, 0) != cudaSuccess) return main;
    return ds().side;
}

// Route each problem: tcgen05/TMA kernel in TF32 mode when its operands satisfy TMA's alignment rules, otherwise
// (and always in FP32 mode) the exact fp32 SIMT kernel.
// Problems of one call are independent (or accumulate atomically), so the tensor-core launches are spread over a few
// auxiliary streams (event fork / join on the caller's stream): the 36-tile encoder GEMMs and the conv-bank members
// then overlap instead of running one under-filled grid after another.
int launch_gemm(const taco_gemm_desc* d, int n_problems, int precision, cudaStream_t s) {
    ProfScope prof_scope(0, s);
    prof_scope.tag(d[0].M, d[0].N, d[0].K, n_problems);
    if (g_prof_on) {
        double fl = 0.0;
        for (int i = 0; i < n_problems; i++) fl += 2.0 * d[i].M * d[i].N * d[i].K;
        prof_scope.work(fl * 1e-9);
    }
    if (precision == TACO_PREC_FP32) return launch_gemm_simt(d, n_problems, s);
    std::vector<taco_gemm_desc> rest;
    std::vector<int> tc;
    std::vector<char> is16;
    for (int i = 0; i < n_problems; i++) {
        const bool have16 = precision == TACO_PREC_BF16 && d[i].A16 && d[i].B16;
        if (have16 && gemm_bf16_eligible(d[i])) { tc.push_back(i); is16.push_back(1); continue; }
        TACO_REQUIRE(d[i].A && d[i].B && d[i].C, TACO_EINVAL, "gemm: problem %d (M=%d N=%d K=%d) has bf16-only operands the bf16 kernel cannot address",
                     i, d[i].M, d[i].N, d[i].K);
        if (gemm_tc_eligible(d[i])) { tc.push_back(i); is16.push_back(0); } else rest.push_back(d[i]);
    }
    const bool fan_out = tc.size() >= 2;
    TACO_TRY(sched_init());
    StreamSet& set = ds().set[(s == ds().side) ? 1 : 0];
    if (fan_out) {
        TACO_CHECK_CUDA(cudaEventRecord(set.fork, s));
        for (int i = 0; i < kAuxStreams; i++) TACO_CHECK_CUDA(cudaStreamWaitEvent(set.aux[i], set.fork, 0));
    }
    int rc_all = TACO_OK;
    for (size_t k = 0; k < tc.size(); k++) {
        cudaStream_t st = fan_out ? set.aux[k % kAuxStreams] : s;
        int rc = is16[k] ? launch_gemm_bf16(d[tc[k]], st, s == ds().side) : launch_gemm_tc(d[tc[k]], st);
        if (rc == TACO_ENOTSUP) rest.push_back(d[tc[k]]);
        else if (rc != TACO_OK && rc_all == TACO_OK) rc_all = rc;
    }
    if (fan_out) {
        for (int i = 0; i < kAuxStreams; i++) {
            TACO_CHECK_CUDA(cudaEventRecord(set.join[i], set.aux[i]));
            TACO_CHECK_CUDA(cudaStreamWaitEvent(s, set.join[i], 0));
        }
    }
    if (rc_all != TACO_OK) return rc_all;
    if (!rest.empty()) {
        for (taco_gemm_desc& r : rest) TACO_REQUIRE(r.C16 == nullptr, TACO_EINVAL, "gemm: a bf16 mirror output needs the bf16 kernel (M=%d N=%d K=%d)", r.M, r.N, r.K);
        return launch_gemm_simt(rest.data(), (int)rest.size(), s);
    }
    return TACO_OK;
}

// ---- Model helpers ------------------------------------------------------------------------------------
int Model::lookup(const std::string& name, Entry& e) const {
    auto it = table.find(name);
    if (it == table.end()) { set_error("parameter '%s' is not bound", name.c_str()); return TACO_EINVAL; }
    e = it->second;
    return TACO_OK;
}
float* Model::P(const std::string& name) const {
    auto it = table.find(name);
    if (it == table.end()) { set_error("parameter '%s' is not bound", name.c_str()); return nullptr; }
    return (it->second.trainable ? params : bn_state) + it->second.offset;
}
float* Model::G(const std::string& name) const {
    auto it = table.find(name);
    if (it == table.end() || !it->second.trainable) { set_error("gradient slot '%s' missing", name.c_str()); return nullptr; }
    return grads + it->second.offset;
}
float* Model::W(const std::string& name) const {
    auto it = regions.find(name);
    if (it == regions.end()) { set_error("workspace region '%s' missing", name.c_str()); return nullptr; }
    return reinterpret_cast<float*>(ws + it->second.offset);
}
double* Model::Wd(const std::string& name) const { return reinterpret_cast<double*>(W(name)); }
void* Model::W16(const std::string& name) const {
    auto it = regions.find(name + ".h");
    return it == regions.end() ? nullptr : static_cast<void*>(ws + it->second.offset);
}
void* Model::P16(const std::string& name) const {
    auto it = table.find(name);
    auto r = regions.find("params16");
    if (it == table.end() || !it->second.trainable || r == regions.end()) return nullptr;
    return ws + r->second.offset + (size_t)it->second.offset * 2;
}

static CbhgGeom make_geom(const std::string& prefix, int N, int T, int Cin, int Kb, int Cb, int P1, int P2, int pw, int depth, int H) {
    CbhgGeom g;
    g.prefix = prefix; g.N = N; g.T = T; g.Kb = Kb; g.PL = (Kb - 1) / 2; g.PR = Kb - 1 - g.PL; g.Tp = T + Kb - 1;
    g.rows = N * g.Tp; g.slack = Kb;
    g.Cin = Cin; g.Cb = Cb; g.P1 = P1; g.P2 = P2; g.pw = pw; g.depth = depth; g.H = H; g.has_hin = (P2 != H);
    return g;
}

void Model::plan(const Shape& s) {
    regions.clear();
    size_t off = 0;
    auto add_e = [&](const std::string& name, std::initializer_list<int64_t> dims, int64_t slack_elems, bool dbl, bool half) {
        Region r{};
        r.ndim = (int)dims.size(); r.is_double = dbl; r.is_half = half;
        int64_t n = 1; int i = 0;
        for (int64_t d : dims) { r.dims[i++] = d; n *= d; }
        int64_t st = 1;
        for (int k = r.ndim - 1; k >= 0; k--) { r.strides[k] = st; st *= r.dims[k]; }
        r.numel = n;
        const size_t esz = dbl ? 8 : (half ? 2 : 4);
        off = (off + 255) / 256 * 256;
        off += (size_t)slack_elems * esz;
        off = (off + 255) / 256 * 256;
        r.offset = off;
        off += (size_t)n * esz + (size_t)slack_elems * esz;
        regions[name] = r;
    };
    auto add = [&](const std::string& name, std::initializer_list<int64_t> dims, int64_t slack_elems = 0, bool dbl = false) { add_e(name, dims, slack_elems, dbl, false); };
    const bool h16 = use16();
    // bf16 mirror "<name>.h" of an fp32 region (same dims and slack), only in the bf16 precision mode
    auto add16 = [&](const std::string& name, std::initializer_list<int64_t> dims, int64_t slack_elems = 0) { if (h16) add_e(name + ".h", dims, slack_elems, false, true); };
    // fp32 region that the bf16 mode replaces by its mirror alone
    auto add32 = [&](const std::string& name, std::initializer_list<int64_t> dims, int64_t slack_elems = 0) { if (!h16) add(name, dims, slack_elems); };
    const taco_config& c = cfg;
    const int N = s.N, tr = s.training;
    enc = make_geom("enc_cbhg", N, s.Ti, c.enc_prenet_sizes[1], c.enc_bank_size, c.enc_bank_channels, c.enc_proj_sizes[0],
                    c.enc_proj_sizes[1], c.enc_proj_width, c.enc_highway_depth, c.enc_rnn_size);
    post = make_geom("post_cbhg", N, s.To, c.num_mels, c.post_bank_size, c.post_bank_channels, c.post_proj_sizes[0],
                     c.post_proj_sizes[1], c.post_proj_width, c.post_highway_depth, c.post_rnn_size);
    for (const CbhgGeom* gp : {&enc, &post}) {
        const CbhgGeom& g = *gp;
        const std::string p = g.prefix + "/";
        const int64_t rows = g.rows, KC = (int64_t)g.Kb * g.Cb, H = g.H;
        add(p + "xin_p", {rows, g.Cin}, (int64_t)g.slack * g.Cin); add16(p + "xin_p", {rows, g.Cin}, (int64_t)g.slack * g.Cin);
        add32(p + "bank_raw", {rows, KC}); add16(p + "bank_raw", {rows, KC});
        add(p + "bank_stats", {2 * KC}, 0, true);
        add(p + "bank_mean", {KC}); add(p + "bank_rstd", {KC}); add(p + "bank_var", {KC});
        // tensors only contractions read live as bf16 alone in the bf16 mode (add32 / add16 pairs)
        add32(p + "pooled_p", {rows, KC}, (int64_t)g.slack * KC); add16(p + "pooled_p", {rows, KC}, (int64_t)g.slack * KC);
        add(p + "p1_raw", {rows, g.P1}); add(p + "p1_stats", {2 * (int64_t)g.P1}, 0, true);
        add(p + "p1_mean", {g.P1}); add(p + "p1_rstd", {g.P1}); add(p + "p1_var", {g.P1});
        add32(p + "p1_p", {rows, g.P1}, (int64_t)g.slack * g.P1); add16(p + "p1_p", {rows, g.P1}, (int64_t)g.slack * g.P1);
        add(p + "p2_raw", {rows, g.P2}); add(p + "p2_stats", {2 * (int64_t)g.P2}, 0, true);
        add(p + "p2_mean", {g.P2}); add(p + "p2_rstd", {g.P2}); add(p + "p2_var", {g.P2});
        add(p + "hw0", {rows, g.P2}); add16(p + "hw0", {rows, g.P2});
        if (g.has_hin) { add(p + "hw_0", {rows, H}); add16(p + "hw_0", {rows, H}); }
        for (int i = 1; i <= g.depth; i++) {
            add(p + "hw_" + std::to_string(i), {rows, H}); add16(p + "hw_" + std::to_string(i), {rows, H});
            add(p + "hwH_" + std::to_string(i), {rows, H});
            add(p + "hwT_" + std::to_string(i), {rows, H});
        }
        add(p + "gx", {rows, 6 * H});
        add(p + "rnn_out", {(int64_t)g.N * g.T, 2 * H}); add16(p + "rnn_out", {(int64_t)g.N * g.T, 2 * H});
        if (tr) {
            const int64_t st = 2 * (int64_t)g.N * g.T * H;
            add(p + "st_r", {st}); add(p + "st_u", {st}); add(p + "st_c", {st}); add(p + "st_hprev", {st});
            add16(p + "st_hprev", {st}); add16(p + "st_rh", {st});
            add(p + "d_rnn_out", {(int64_t)g.N * g.T, 2 * H});
            add32(p + "dgx", {rows, 6 * H}); add16(p + "dgx", {rows, 6 * H});
            add32(p + "dgx_dense", {2, (int64_t)g.N * g.T, 3 * H}); add16(p + "dgx_dense", {2, (int64_t)g.N * g.T, 3 * H});
            add(p + "d_hwA", {rows, H}); add(p + "d_hwB", {rows, H});
            // per layer: (dHpre | dTpre) side by side and the packed [H, 2H] weight their shared data gradient uses
            for (int i = 1; i <= g.depth; i++) {
                add32(p + "d_HT_" + std::to_string(i), {rows, 2 * H}); add16(p + "d_HT_" + std::to_string(i), {rows, 2 * H});
                add(p + "hw_wcat_" + std::to_string(i), {H, 2 * H}); add16(p + "hw_wcat_" + std::to_string(i), {H, 2 * H});
            }
            add(p + "gru_wxcat", {H, 6 * H}); add16(p + "gru_wxcat", {H, 6 * H});   // x-side rows of the four GRU kernels: fw r|u, fw c, bw r|u, bw c
            if (g.has_hin) add(p + "d_hw0", {rows, g.P2});
            add32(p + "d_p2raw", {rows, g.P2}, (int64_t)g.slack * g.P2); add16(p + "d_p2raw", {rows, g.P2}, (int64_t)g.slack * g.P2);
            add(p + "d_p1p", {rows, g.P1});
            add32(p + "d_p1raw", {rows, g.P1}, (int64_t)g.slack * g.P1); add16(p + "d_p1raw", {rows, g.P1}, (int64_t)g.slack * g.P1);
            add32(p + "d_pooled", {rows, KC}); add16(p + "d_pooled", {rows, KC});
            add32(p + "d_bank", {rows, KC}, (int64_t)g.slack * KC); add16(p + "d_bank", {rows, KC}, (int64_t)g.slack * KC);
            add(p + "d_xin_p", {rows, g.Cin});
            add(p + "d_before", {g.N, g.P2});
            add(p + "d_h0", {g.N, 2 * H});
            // flipped + transposed bank kernels, stored back to back ([sum_k k*Cb, Cin]) so that the bank's data gradient is
            // ONE GEMM; "bank_taps" is its per-k-tile (column, row offset) table (int32 pairs)
            add(p + "bank_wd", {(int64_t)g.Cb * g.Kb * (g.Kb + 1) / 2, g.Cin}); add16(p + "bank_wd", {(int64_t)g.Cb * g.Kb * (g.Kb + 1) / 2, g.Cin});
            add(p + "bank_taps", {(int64_t)2 * (g.Cb / 32 + 1) * g.Kb * (g.Kb + 1) / 2});
            add(p + "proj_1/wd", {(int64_t)g.pw * g.P1, KC}); add16(p + "proj_1/wd", {(int64_t)g.pw * g.P1, KC});
            add(p + "proj_2/wd", {(int64_t)g.pw * g.P2, g.P1}); add16(p + "proj_2/wd", {(int64_t)g.pw * g.P2, g.P1});
        }
    }
    if (h16) add_e("params16", {(n_trainable + 7) / 8 * 8}, 0, false, true);       // bf16 mirror of the flat trainable buffer
    // encoder prenet as lookup tables over the symbol set (the prenet is position-wise: 80 distinct rows)
    add("enc/table1", {c.num_symbols, c.enc_prenet_sizes[0]});
    add("enc/table2", {c.num_symbols, c.enc_prenet_sizes[1]});
    if (tr) {
        add("enc/d_table2", {c.num_symbols, c.enc_prenet_sizes[1]}); add("enc/d_t2pre", {c.num_symbols, c.enc_prenet_sizes[1]});
        add("enc/d_table1", {c.num_symbols, c.enc_prenet_sizes[0]}); add("enc/d_t1pre", {c.num_symbols, c.enc_prenet_sizes[0]});
    }
    // decoder
    {
        const int64_t rows = (int64_t)N * s.Td;
        const int64_t M = c.num_mels, r = c.reduction_factor, E = 2 * c.enc_rnn_size, A = c.attention_size, HA = c.attention_state_size;
        const int64_t Z1 = c.dec_prenet_sizes[0], Z = c.dec_prenet_sizes[1], Y = c.dec_rnn_size;
        const int64_t SPK = (c.speaker_mode == TACO_SPK_SIMPLE) ? c.speaker_embedding_size : 0;
        add("dec/keys", {(int64_t)N * s.Ti, A});
        add("dec/x_all", {rows, M}); add("dec/px", {rows, Z1});
        add("dec/y0", {rows, Y}); add("dec/y1", {rows, Y}); add("dec/y2", {rows, Y});
        add("dec/g1_gx", {rows, 3 * Y}); add("dec/g2_gx", {rows, 3 * Y});
        add("alignments", {N, s.Ti, s.Td});
        if (c.precision != TACO_PREC_FP32) {
            // free-running decoder (inference, rnn_decoder_test_mode): fragment-packed bf16 weights of one decoder step
            AttArgs pa{};
            pa.free_run = 1; pa.fast = 1; pa.SPK = (int)SPK; pa.E = (int)E; pa.A = (int)A; pa.HA = (int)HA; pa.Z1 = (int)Z1; pa.Z = (int)Z; pa.Y = (int)Y;
            pa.M = (int)M; pa.r = (int)r;
            const size_t wb = att_wfrag_bytes(pa);
            if (wb) add("dec/wfrag", {(int64_t)(wb / 4)});
            // low-batch synthesis: per-rank shared-memory weight images of the resident-weight decoder kernel (att_free.cu)
            pa.N = N; pa.Ti = s.Ti; pa.Td = s.Td; pa.att_type = c.attention_type;
            if (!tr && att_free_supported(pa)) add("dec/frimg", {(int64_t)(att_free_image_bytes() / 4)});
        }
        if (tr) {
            add("dec/s_z1", {rows, Z1}); add("dec/s_z", {rows, Z});
            for (const char* nm : {"dec/s_r", "dec/s_u", "dec/s_c", "dec/s_haprev", "dec/s_ha"}) add(nm, {rows, HA});
            add("dec/s_q", {rows, A}); add("dec/s_ctxin", {rows, E}); add("dec/s_ctx", {rows, E});
            add("dec/s_e", {rows, s.Ti}); add("dec/s_a", {rows, s.Ti});
            for (int l = 1; l <= 2; l++) {
                const std::string rp = "dec/g" + std::to_string(l) + "_";
                for (const char* nm : {"st_r", "st_u", "st_c", "st_hprev"}) add(rp + nm, {rows, Y});
                add(rp + "dgx", {rows, 3 * Y}); add(rp + "dh0", {N, Y}); add(rp + "wxcat", {Y, 3 * Y});
                // decoder wavefront (model_decoder.cu): gathered chunk operands and the state / gradient carried between chunks
                add(rp + "wv_x", {rows, Y}); add(rp + "wv_g", {rows, 3 * Y}); add(rp + "wv_h", {N, Y});
            }
            add("dec/c_dha", {N, HA}); add("dec/c_dctx", {N, E}); add("dec/c_dac", {N, (int64_t)((s.Ti + 15) / 16 * 16)});
            // operand images of the fast attention kernels (att_fast.cu: launch_att_fast_pack), sized in floats
            for (int bwd = 0; bwd < 2; bwd++) {
                size_t wb = 0, kmb = 0;
                att_fast_image_bytes(s.Ti, bwd != 0, &wb, &kmb);
                add(bwd ? "dec/img_bw" : "dec/img_fw", {(int64_t)(wb / 4)});
                add(bwd ? "dec/img_bkm" : "dec/img_fkm", {(int64_t)(kmb / 4) * 16 * ((N + 7) / 8)});
            }
            add("dec/d_dec", {rows, M * r}); add("dec/d_y2", {rows, Y}); add("dec/d_y1", {rows, Y}); add("dec/d_y0", {rows, Y});
            const int64_t ZS = Z + SPK, KIN = ZS + HA, KO = HA + E + SPK;
            add("dec/W1cT", {Z1, E}); add("dec/W2T", {Z, Z1}); add("dec/WgT", {2 * HA, KIN}); add("dec/WcT", {HA, KIN});
            add("dec/WqT", {A, HA}); add("dec/WoT", {Y, KO});
            add("dec/d_G", {rows, 3 * HA}); add("dec/d_zp", {rows, Z}); add("dec/d_z1p", {rows, Z1}); add("dec/d_ctx", {rows, E});
            add("dec/d_gq", {rows, A}); add("dec/d_ge", {rows, s.Ti}); add("dec/d_ha0", {N, HA});
            add("dec/d_keys", {(int64_t)N * s.Ti, A});
            add("dec/v_eff", {A}); add("dec/g_veff", {A});
        }
    }
    // The linear-spectrogram tensors use a row pitch rounded up to 64 elements (1025 -> 1088) so TMA can address them (one 3-D box per weight tile) and
    // their bf16 mirrors (16-byte pitches); the pad columns stay zero.  "linear_outputs" is exposed as a strided [N,To,F] view.
    {
        const int64_t Fp = (c.num_freq + 63) / 64 * 64;
        add("linear_buf", {(int64_t)N * s.To, Fp});
        Region r = regions["linear_buf"];
        r.ndim = 3; r.dims[0] = N; r.dims[1] = s.To; r.dims[2] = c.num_freq;
        r.strides[0] = (int64_t)s.To * Fp; r.strides[1] = Fp; r.strides[2] = 1;
        r.numel = (int64_t)N * s.To * c.num_freq;
        regions["linear_outputs"] = r;
        const int64_t wrows = 2 * (int64_t)c.post_rnn_size + ((c.speaker_mode == TACO_SPK_SIMPLE) ? c.speaker_embedding_size : 0);
        add("linear/w_pad", {wrows, Fp}); add16("linear/w_pad", {wrows, Fp});
        // (bf16 mode: the loss gradient exists as bf16 only, except under 'simple' speaker injection, whose time sums read fp32)
        if (tr) { if (!h16 || c.speaker_mode == TACO_SPK_SIMPLE) add("d_linear", {(int64_t)N * s.To, Fp}); add16("d_linear", {(int64_t)N * s.To, Fp}); }
    }
    if (tr) add("post_cbhg/d_mel_loss", {post.rows, c.num_mels});
    if (c.speaker_mode == TACO_SPK_DEEPVOICE || c.speaker_mode == TACO_SPK_DEEPVOICE_TABLE) {
        const int64_t S = c.speaker_embedding_size, Y = c.dec_rnn_size;
        add("spk/embed", {N, S});
        add("spk/before", {N, c.enc_prenet_sizes[1]}); add("spk/enc_init", {N, 2 * (int64_t)c.enc_rnn_size});
        add("spk/att_init", {N, c.attention_state_size}); add("spk/dec_init1", {N, Y}); add("spk/dec_init2", {N, Y});
        if (tr) { add("spk/d_pre", {N, 2 * (int64_t)c.enc_rnn_size + c.attention_state_size + Y}); add("spk/d_embed", {N, S}); }
    }
    if (c.speaker_mode == TACO_SPK_SIMPLE) {
        const int64_t S = c.speaker_embedding_size, Fp = (c.num_freq + 63) / 64 * 64;
        add("spk/embed", {N, S}); add("spk/lin_bias", {N, Fp});
        if (tr) {
            add("spk/d_embed", {N, S}); add("spk/s_lin", {N, Fp});
            add("spk/s_y0", {N, c.dec_rnn_size}); add("spk/s_G", {N, 3 * (int64_t)c.attention_state_size});
        }
    }
    add("scalars", {8}, 0, true);
    add("scalars_f", {8});
    // zero-copy view of the mel outputs inside the post-net's padded input
    {
        Region r = regions["post_cbhg/xin_p"];
        r.offset += (size_t)post.PL * c.num_mels * 4;
        r.ndim = 3; r.dims[0] = N; r.dims[1] = s.To; r.dims[2] = c.num_mels;
        r.strides[0] = (int64_t)post.Tp * c.num_mels; r.strides[1] = c.num_mels; r.strides[2] = 1;
        r.numel = (int64_t)N * s.To * c.num_mels;
        regions["mel_outputs"] = r;
    }
    plan_bytes = (off + 255) / 256 * 256;
    shape = s; planned = true; tables_ready = false; prep_done = false;
}

static int shape_of(const Model& m, const taco_batch* b, Shape& s) {
    TACO_REQUIRE(b && b->N > 0 && b->T_in > 0, TACO_ESHAPE, "batch: N and T_in must be positive");
    s.N = b->N; s.Ti = b->T_in;
    s.training = (b->linear_targets != nullptr) ? 1 : 0;      // tacotron.py:26
    const int r = m.cfg.reduction_factor;
    if (b->mel_targets) {
        TACO_REQUIRE(b->T_out > 0 && b->T_out % r == 0, TACO_ESHAPE, "batch: T_out=%d must be a positive multiple of r=%d", b->T_out, r);
        s.Td = b->T_out / r;
        if (b->decoder_steps > 0 && b->decoder_steps < s.Td) s.Td = b->decoder_steps;
    } else {
        TACO_REQUIRE(b->decoder_steps > 0, TACO_ESHAPE, "batch: decoder_steps must be given without targets");
        s.Td = b->decoder_steps;
    }
    s.To = s.Td * r;
    return TACO_OK;
}

static int ensure_plan(Model& m, const Shape& s) {
    if (!(m.planned && m.shape == s)) m.plan(s);
    TACO_REQUIRE(m.ws != nullptr, TACO_ESTATE, "no workspace bound");
    TACO_REQUIRE(m.plan_bytes <= m.ws_bytes, TACO_ENOMEM, "workspace too small: need %zu bytes, bound %zu", m.plan_bytes, m.ws_bytes);
    return TACO_OK;
}

static taco_gemm_desc gd0(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc) {
    taco_gemm_desc d{};
    d.A = A; d.B = B; d.C = C; d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.alpha = 1.f; d.split_k = 1;
    return d;
}

float* bank_wd(const Model& m, const CbhgGeom& g, int k) {
    return m.W(g.prefix + "/bank_wd") + (long long)g.Cb * g.Cin * (k - 1) * k / 2;
}

// Operands of the backward pass that depend only on the parameters: flipped + transposed convolution kernels, packed
// x-side GRU / highway weights, transposed attention weights, the bank tap tables.  Runs on the leaf stream beside the
// forward pass (taco_forward), or in line at the start of the backward pass when the two-stream schedule is off.
static int backward_prep(Model& m, cudaStream_t s) {
    const taco_config& c = m.cfg;
    auto off16 = [](void* p, long long elems) -> void* { return p ? static_cast<void*>(static_cast<uint16_t*>(p) + elems) : nullptr; };
    for (const CbhgGeom* gp : {&m.enc, &m.post}) {
        const CbhgGeom& g = *gp; const std::string p = g.prefix + "/";
        const int H = g.H;
        // one table per CBHG (prep_ops_kernel): every packed operand and, in bf16 mode, its mirror in the same pass (the
        // contraction kernels of the bf16 mode read nothing else)
        PrepTable tb;
        bool ok = true;
        for (int k = 1; k <= g.Kb; k++) {
            const long long o = (long long)g.Cb * g.Cin * (k - 1) * k / 2;
            ok = ok && tb.pack_dgrad(m.P(p + "bank_" + std::to_string(k) + "/kernel"), bank_wd(m, g, k), off16(m.W16(p + "bank_wd"), o), k, g.Cin, g.Cb);
        }
        ok = ok && tb.pack_dgrad(m.P(p + "proj_1/kernel"), m.W(p + "proj_1/wd"), m.W16(p + "proj_1/wd"), g.pw, g.Kb * g.Cb, g.P1);
        ok = ok && tb.pack_dgrad(m.P(p + "proj_2/kernel"), m.W(p + "proj_2/wd"), m.W16(p + "proj_2/wd"), g.pw, g.P1, g.P2);
        float* wx = m.W(p + "gru_wxcat"); void* wx16 = m.W16(p + "gru_wxcat");
        const char* dirs[2] = {"gru_fw", "gru_bw"};
        for (int dd = 0; dd < 2; dd++) {
            ok = ok && tb.copy2d(wx + dd * 3 * H, off16(wx16, dd * 3 * H), m.P(p + dirs[dd] + "/gates_kernel"), H, 2 * H, 6 * H, 2 * H);
            ok = ok && tb.copy2d(wx + dd * 3 * H + 2 * H, off16(wx16, dd * 3 * H + 2 * H), m.P(p + dirs[dd] + "/cand_kernel"), H, H, 6 * H, H);
        }
        for (int i = 1; i <= g.depth; i++) {
            const std::string hn = p + "highway_" + std::to_string(i);
            float* wc = m.W(p + "hw_wcat_" + std::to_string(i)); void* wc16 = m.W16(p + "hw_wcat_" + std::to_string(i));
            ok = ok && tb.copy2d(wc, wc16, m.P(hn + "/H_kernel"), H, H, 2 * H, H);
            ok = ok && tb.copy2d(wc + H, off16(wc16, H), m.P(hn + "/T_kernel"), H, H, 2 * H, H);
        }
        TACO_REQUIRE(ok, TACO_EINVAL, "backward_prep: operand table of %s overflows", g.prefix.c_str());
        TACO_TRY(launch_prep_ops(tb, s));
        const int tb_w = m.use16() ? 64 : 32;          // k-tile width of the kernel that walks the table
        if (!m.tables_ready && g.Cb % tb_w == 0) {
            // k-tile i of the merged bank data gradient: member k, tap j, channel block q  ->  column (k-1)*Cb + tb_w*q of
            // d_bank, row offset j - r_k (+ Kb: the operand base sits Kb slack rows before the buffer)
            std::vector<int>& tab = (gp == &m.enc) ? m.taps_enc : m.taps_post;
            tab.clear();
            for (int k = 1; k <= g.Kb; k++) {
                const int l = (k - 1) / 2, r = k - 1 - l;
                for (int j = 0; j < k; j++)
                    for (int q = 0; q < g.Cb / tb_w; q++) { tab.push_back((k - 1) * g.Cb + tb_w * q); tab.push_back(j - r + g.Kb); }
            }
            TACO_CHECK_CUDA(cudaMemcpyAsync(m.W(p + "bank_taps"), tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        }
    }
    m.tables_ready = true;
    {
        const int M = c.num_mels, E = 2 * c.enc_rnn_size, A = c.attention_size, HA = c.attention_state_size;
        const int Z1 = c.dec_prenet_sizes[0], Z = c.dec_prenet_sizes[1], Y = c.dec_rnn_size;
        const int SPK = (c.speaker_mode == TACO_SPK_SIMPLE) ? c.speaker_embedding_size : 0;
        const int KIN = Z + SPK + HA, KO = HA + E + SPK;
        PrepTable tb;
        bool ok = true;
        ok = ok && tb.transpose(m.P("dec_prenet/dense_1/kernel") + (long long)M * Z1, m.W("dec/W1cT"), E, Z1);
        ok = ok && tb.transpose(m.P("dec_prenet/dense_2/kernel"), m.W("dec/W2T"), Z1, Z);
        ok = ok && tb.transpose(m.P("attention_gru/gates_kernel"), m.W("dec/WgT"), KIN, 2 * HA);
        ok = ok && tb.transpose(m.P("attention_gru/cand_kernel"), m.W("dec/WcT"), KIN, HA);
        ok = ok && tb.transpose(m.P("attention/query_kernel"), m.W("dec/WqT"), HA, A);
        ok = ok && tb.transpose(m.P("concat_proj/kernel"), m.W("dec/WoT"), KO, Y);
        for (int l = 1; l <= 2; l++) {
            const std::string gn = "dec_gru_" + std::to_string(l);
            float* wx = m.W("dec/g" + std::to_string(l) + "_wxcat");
            ok = ok && tb.copy2d(wx, nullptr, m.P(gn + "/gates_kernel"), Y, 2 * Y, 3 * Y, 2 * Y);
            ok = ok && tb.copy2d(wx + 2 * Y, nullptr, m.P(gn + "/cand_kernel"), Y, Y, 3 * Y, Y);
        }
        TACO_REQUIRE(ok, TACO_EINVAL, "backward_prep: decoder operand table overflows");
        TACO_TRY(launch_prep_ops(tb, s));
    }
    return TACO_OK;
}

// bf16 mirror of every trainable tensor (18.7 MB written per step; the GEMM B operands of this pass and the next backward pass)
static int refresh_params16(Model& m, cudaStream_t s) {
    if (!m.use16()) return TACO_OK;
    auto it = m.regions.find("params16");
    TACO_REQUIRE(it != m.regions.end() && it->second.numel >= m.n_trainable, TACO_ESTATE, "bf16 mode: bind the parameters before sizing the workspace");
    return launch_cast2d_bf16(m.ws + it->second.offset, m.params, 1, (int)m.n_trainable, m.n_trainable, m.n_trainable, s);
}

// speaker injection vectors of the batch (tacotron.py:41-94): regions spk/*
static int speaker_forward(Model& m, const taco_batch* b, cudaStream_t s) {
    const taco_config& c = m.cfg;
    const bool simple = (c.speaker_mode == TACO_SPK_SIMPLE);
    if (simple) {
        // 'simple' injection: one embedding row per utterance, concatenated at three sites (tacotron.py:44-49,82-86)
        TACO_REQUIRE(b->speaker_id != nullptr, TACO_EINVAL, "speaker_id is required when num_speakers > 1");
        TACO_TRY(launch_gather_rows(m.P("speaker_embedding"), b->speaker_id, m.W("spk/embed"), m.shape.N, 1, 1, 0, c.speaker_embedding_size, c.num_speakers, s));
    }
    const bool spk = (c.speaker_mode == TACO_SPK_DEEPVOICE || c.speaker_mode == TACO_SPK_DEEPVOICE_TABLE);
    struct Site { const char* name; const char* region; int dim; };
    const Site sites[5] = {{"before_highway", "spk/before", c.enc_prenet_sizes[1]}, {"encoder_rnn_init_state", "spk/enc_init", 2 * c.enc_rnn_size},
                           {"attention_rnn_init_state", "spk/att_init", c.attention_state_size},
                           {"decoder_rnn_init_states1", "spk/dec_init1", c.dec_rnn_size}, {"decoder_rnn_init_states2", "spk/dec_init2", c.dec_rnn_size}};
    if (spk) {
        // DeepVoice-2 style injection (tacotron.py:41-81): five vectors per utterance from the speaker id
        TACO_REQUIRE(b->speaker_id != nullptr, TACO_EINVAL, "speaker_id is required when num_speakers > 1");
        const int S = c.speaker_embedding_size;
        if (c.speaker_mode == TACO_SPK_DEEPVOICE) {
            TACO_TRY(launch_gather_rows(m.P("speaker_embedding"), b->speaker_id, m.W("spk/embed"), m.shape.N, 1, 1, 0, S, c.num_speakers, s));
            for (const Site& st : sites) {
                taco_gemm_desc d = gd0(m.W("spk/embed"), m.P(std::string("speaker/") + st.name + "/kernel"), m.W(st.region), m.shape.N, st.dim, S, S, st.dim, st.dim);
                d.bias = m.P(std::string("speaker/") + st.name + "/bias"); d.act = ACT_SOFTSIGN;
                TACO_TRY(launch_gemm(&d, 1, TACO_PREC_FP32, s));
            }
        } else {
            for (const Site& st : sites)
                TACO_TRY(launch_gather_rows(m.P(std::string("speaker/") + st.name + "/table"), b->speaker_id, m.W(st.region), m.shape.N, 1, 1, 0, st.dim, c.num_speakers, s));
        }
    }
    return TACO_OK;
}

static int model_forward(Model& m, const taco_batch* b, cudaStream_t s) {
    const taco_config& c = m.cfg;
    const int prec = c.precision, tr = m.shape.training;
    const bool simple = (c.speaker_mode == TACO_SPK_SIMPLE);
    const bool spk = (c.speaker_mode == TACO_SPK_DEEPVOICE || c.speaker_mode == TACO_SPK_DEEPVOICE_TABLE);
    TACO_TRY(refresh_params16(m, s));
    TACO_TRY(speaker_forward(m, b, s));
    // ---- encoder prenet over the symbol table, then lookup (tacotron.py:34-39,101-103; modules.py:18-25; dropout = identity) ----
    const int V = c.num_symbols, E0 = c.embedding_size, E1 = c.enc_prenet_sizes[0], E2 = c.enc_prenet_sizes[1];
    {
        taco_gemm_desc d = gd0(m.P("embedding"), m.P("enc_prenet/dense_1/kernel"), m.W("enc/table1"), V, E1, E0, E0, E1, E1);
        d.bias = m.P("enc_prenet/dense_1/bias"); d.act = ACT_RELU;
        TACO_TRY(launch_gemm(&d, 1, prec, s));
        d = gd0(m.W("enc/table1"), m.P("enc_prenet/dense_2/kernel"), m.W("enc/table2"), V, E2, E1, E1, E2, E2);
        d.bias = m.P("enc_prenet/dense_2/bias"); d.act = ACT_RELU;
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    TACO_TRY(launch_gather_rows(m.W("enc/table2"), b->inputs, m.W("enc_cbhg/xin_p"), m.enc.N, m.enc.T, m.enc.Tp, m.enc.PL, E2, V, s, m.W16("enc_cbhg/xin_p")));
    prof_mark("fwd:enc_cbhg", s);
    TACO_TRY(cbhg_forward(m, m.enc, b->input_lengths, spk ? m.W("spk/before") : nullptr, spk ? m.W("spk/enc_init") : nullptr, tr, s));
    prof_mark("fwd:decoder", s);
    TACO_TRY(decoder_forward(m, b, s));
    prof_mark("fwd:post_cbhg", s);
    if (m.use16())     // the decoder's small GEMMs stay on the fp32-operand kernels: mirror the mel frames they wrote (pad rows stay zero)
        TACO_TRY(launch_cast2d_bf16(m.W16("post_cbhg/xin_p"), m.W("post_cbhg/xin_p"), m.post.rows, c.num_mels, c.num_mels, c.num_mels, s));
    TACO_TRY(cbhg_forward(m, m.post, nullptr, nullptr, nullptr, tr, s));
    prof_mark("fwd:linear", s);
    // ---- linear-spectrogram projection (tacotron.py:235) ----
    {
        const int Hp2 = 2 * c.post_rnn_size, F = c.num_freq, Fp = (F + 63) / 64 * 64;
        const int S = simple ? c.speaker_embedding_size : 0;
        TACO_TRY(launch_copy2d(m.W("linear/w_pad"), m.P("linear/kernel"), Hp2 + S, F, Fp, F, s));     // 16-byte row pitch for TMA
        if (m.use16()) TACO_TRY(launch_cast2d_bf16(m.W16("linear/w_pad"), m.P("linear/kernel"), Hp2 + S, F, Fp, F, s));
        taco_gemm_desc d = gd0(m.W("post_cbhg/rnn_out"), m.W("linear/w_pad") + (long long)S * Fp, m.W("linear_buf"), m.shape.N * m.shape.To, F, Hp2, Hp2, Fp, Fp);
        if (m.use16()) { d.A16 = m.W16("post_cbhg/rnn_out"); d.B16 = static_cast<uint16_t*>(m.W16("linear/w_pad")) + (long long)S * Fp; }
        if (simple) {
            // concat([tiled speaker_embed, post_outputs]) . W  ==  post . W[S:] + (embed . W[:S] + b) tiled over time   tacotron.py:226-235
            taco_gemm_desc e = gd0(m.W("spk/embed"), m.W("linear/w_pad"), m.W("spk/lin_bias"), m.shape.N, F, S, S, Fp, Fp);
            e.bias = m.P("linear/bias");
            TACO_TRY(launch_gemm(&e, 1, TACO_PREC_FP32, s));
            TACO_TRY(launch_bcast_rows(m.W("spk/lin_bias"), m.W("linear_buf"), m.shape.N, m.shape.To, F, Fp, s));
            d.accumulate = 1;
        } else {
            d.bias = m.P("linear/bias");
        }
        TACO_TRY(launch_gemm(&d, 1, prec, s));
    }
    prof_mark("fwd:end", s);
    return TACO_OK;
}

static int model_backward(Model& m, const taco_batch* b, cudaStream_t s) {
    const taco_config& c = m.cfg;
    const int prec = c.precision;
    const int N = m.shape.N, To = m.shape.To, M = c.num_mels, F = c.num_freq;
    TACO_REQUIRE(m.shape.training && b->mel_targets && b->linear_targets, TACO_ESTATE, "backward needs a training forward with targets");
    TACO_REQUIRE(!b->rnn_decoder_test_mode, TACO_ESTATE, "backward through the free-running decoder is not defined (train.py:158-166 builds that model forward-only)");
    prof_mark("bwd:start", s);
    TACO_CHECK_CUDA(cudaMemsetAsync(m.grads, 0, sizeof(float) * (size_t)m.n_trainable, s));
    double* sc = m.Wd("scalars");
    TACO_CHECK_CUDA(cudaMemsetAsync(sc, 0, sizeof(double) * 8, s));
    if (m.prep_done) {
        TACO_CHECK_CUDA(cudaStreamWaitEvent(s, ds().ev_prep, 0));   // packed operands were produced beside the forward pass
    } else {
        TACO_TRY(backward_prep(m, s));
    }
    m.prep_done = false;
    prof_mark("bwd:loss", s);
    // ---- losses and their gradients (tacotron.py:274-302) ----
    const CbhgGeom& g = m.post;
    const double cnt_mel = (double)N * To * M, cnt_lin = (double)N * To * F;
    float w_all = (float)(1.0 / cnt_lin), w_band = 0.f; int lo = 0, hi = 0;
    if (c.prioritize_loss) {
        lo = c.priority_lo; hi = c.priority_hi;
        w_all = (float)(0.5 / cnt_lin); w_band = (float)(0.5 / ((double)N * To * (hi - lo)));
    }
    const int Fp = (F + 63) / 64 * 64;
    float* d_lin32 = m.has_region("d_linear") ? m.W("d_linear") : nullptr;
    TACO_TRY(launch_l1_loss(m.W("linear_buf"), (long long)To * Fp, Fp, b->linear_targets, b->loss_coeff,
                            d_lin32, (long long)To * Fp, Fp, N, To, F, w_all, w_band, lo, hi, sc + 3, s, m.W16("d_linear"), b->linear_targets_bf16));
    TACO_TRY(launch_l1_loss(m.W("mel_outputs"), (long long)g.Tp * M, M, b->mel_targets, b->loss_coeff,
                            m.W("post_cbhg/d_mel_loss") + (long long)g.PL * M, (long long)g.Tp * M, M, N, To, M,
                            (float)(1.0 / cnt_mel), 0.f, 0, 0, sc + 0, s));
    prof_mark("bwd:linear", s);
    // ---- linear projection backward ----
    {
        const int Hp2 = 2 * c.post_rnn_size; const long long rows = (long long)N * To;
        const int S = (c.speaker_mode == TACO_SPK_SIMPLE) ? c.speaker_embedding_size : 0;
        taco_gemm_desc w = gd0(m.W("post_cbhg/rnn_out"), d_lin32, m.G("linear/kernel") + (long long)S * F, Hp2, F, (int)rows, Hp2, Fp, F);
        w.transA = 1; w.accumulate = 1; w.split_k = 8;
        w.A16 = m.W16("post_cbhg/rnn_out"); w.B16 = m.W16("d_linear");
        cudaStream_t leaf = fork_side(s);
        TACO_TRY(launch_gemm(&w, 1, prec, leaf));
        if (d_lin32) TACO_TRY(launch_colsum(d_lin32, m.G("linear/bias"), rows, F, Fp, leaf));
        else TACO_TRY(launch_colsum16(m.W16("d_linear"), m.G("linear/bias"), rows, F, Fp, leaf));
        if (S) {
            // speaker rows of the kernel and the embedding gradient see d_linear only through its sum over time
            TACO_CHECK_CUDA(cudaMemsetAsync(m.W("spk/s_lin"), 0, sizeof(float) * (size_t)N * Fp, s));
            TACO_CHECK_CUDA(cudaMemsetAsync(m.W("spk/d_embed"), 0, sizeof(float) * (size_t)N * S, s));
            TACO_TRY(launch_timesum(m.W("d_linear"), m.W("spk/s_lin"), N, To, To, 0, Fp, s));
            taco_gemm_desc ws = gd0(m.W("spk/embed"), m.W("spk/s_lin"), m.G("linear/kernel"), S, F, N, S, Fp, F);
            ws.transA = 1; ws.accumulate = 1;
            TACO_TRY(launch_gemm(&ws, 1, TACO_PREC_FP32, s));
            taco_gemm_desc es = gd0(m.W("spk/s_lin"), m.W("linear/w_pad"), m.W("spk/d_embed"), N, S, F, Fp, Fp, S);
            es.transB = 1; es.accumulate = 1;
            TACO_TRY(launch_gemm(&es, 1, TACO_PREC_FP32, s));
        }
        taco_gemm_desc e = gd0(d_lin32, m.W("linear/w_pad") + (long long)S * Fp, m.W("post_cbhg/d_rnn_out"), (int)rows, Hp2, F, Fp, Fp, Hp2);
        e.transB = 1;
        if (m.use16()) { e.A16 = m.W16("d_linear"); e.B16 = static_cast<uint16_t*>(m.W16("linear/w_pad")) + (long long)S * Fp; }
        TACO_TRY(launch_gemm(&e, 1, prec, s));
    }
    prof_mark("bwd:post_cbhg", s);
    TACO_TRY(cbhg_backward(m, m.post, nullptr, false, false, s));
    TACO_TRY(launch_axpy(m.W("post_cbhg/d_xin_p"), m.W("post_cbhg/d_mel_loss"), 1.f, (long long)g.rows * M, s));
    prof_mark("bwd:decoder", s);
    TACO_TRY(decoder_backward(m, b, s));
    // Data parallel: every gradient of the "early" bucket (decoder, post-net, linear: the tail of the flat buffer) is enqueued by
    // now - on this stream and, for the leaves, on the side stream.  taco_dp_wait_bucket lets a communication stream start that
    // bucket's all-reduce here, beside the encoder's backward pass.
    TACO_CHECK_CUDA(cudaEventRecord(ds().ev_early[0], s));
    TACO_CHECK_CUDA(cudaEventRecord(ds().ev_early[1], overlap_on() ? ds().side : s));
    m.early_recorded = true;
    prof_mark("bwd:enc_cbhg", s);
    const bool spk = (c.speaker_mode == TACO_SPK_DEEPVOICE || c.speaker_mode == TACO_SPK_DEEPVOICE_TABLE);
    TACO_TRY(cbhg_backward(m, m.enc, b->input_lengths, spk, spk, s));
    if (spk) {
        // gradients of the five injected vectors -> their dense layers / tables -> the speaker embedding table
        struct Site { const char* name; const char* region; const char* grad; int dim; };
        const Site sites[5] = {{"before_highway", "spk/before", "enc_cbhg/d_before", c.enc_prenet_sizes[1]},
                               {"encoder_rnn_init_state", "spk/enc_init", "enc_cbhg/d_h0", 2 * c.enc_rnn_size},
                               {"attention_rnn_init_state", "spk/att_init", "dec/d_ha0", c.attention_state_size},
                               {"decoder_rnn_init_states1", "spk/dec_init1", "dec/g1_dh0", c.dec_rnn_size},
                               {"decoder_rnn_init_states2", "spk/dec_init2", "dec/g2_dh0", c.dec_rnn_size}};
        const int S = c.speaker_embedding_size, Nb = m.shape.N;
        if (c.speaker_mode == TACO_SPK_DEEPVOICE) {
            TACO_CHECK_CUDA(cudaMemsetAsync(m.W("spk/d_embed"), 0, sizeof(float) * (size_t)Nb * S, s));
            for (const Site& st : sites) {
                float* dpre = m.W("spk/d_pre");
                TACO_TRY(launch_softsign_bwd(m.W(st.grad), m.W(st.region), dpre, (long long)Nb * st.dim, s));
                const std::string pn = std::string("speaker/") + st.name;
                taco_gemm_desc w = gd0(m.W("spk/embed"), dpre, m.G(pn + "/kernel"), S, st.dim, Nb, S, st.dim, st.dim);
                w.transA = 1; w.accumulate = 1;
                TACO_TRY(launch_gemm(&w, 1, TACO_PREC_FP32, s));
                TACO_TRY(launch_colsum(dpre, m.G(pn + "/bias"), Nb, st.dim, st.dim, s));
                taco_gemm_desc e = gd0(dpre, m.P(pn + "/kernel"), m.W("spk/d_embed"), Nb, S, st.dim, st.dim, st.dim, S);
                e.transB = 1; e.accumulate = 1;
                TACO_TRY(launch_gemm(&e, 1, TACO_PREC_FP32, s));
            }
            TACO_TRY(launch_scatter_add_rows(m.W("spk/d_embed"), b->speaker_id, m.G("speaker_embedding"), Nb, 1, 1, 0, S, c.num_speakers, s));
        } else {
            for (const Site& st : sites)
                TACO_TRY(launch_scatter_add_rows(m.W(st.grad), b->speaker_id, m.G(std::string("speaker/") + st.name + "/table"), Nb, 1, 1, 0, st.dim, c.num_speakers, s));
        }
    }
    if (c.speaker_mode == TACO_SPK_SIMPLE)      // "spk/d_embed" now holds the linear, concat-projection and attention-GRU terms
        TACO_TRY(launch_scatter_add_rows(m.W("spk/d_embed"), b->speaker_id, m.G("speaker_embedding"), m.shape.N, 1, 1, 0,
                                         c.speaker_embedding_size, c.num_speakers, s));
    // ---- encoder prenet tables backward ----
    {
        const int V = c.num_symbols, E0 = c.embedding_size, E1 = c.enc_prenet_sizes[0], E2 = c.enc_prenet_sizes[1];
        TACO_CHECK_CUDA(cudaMemsetAsync(m.W("enc/d_table2"), 0, sizeof(float) * (size_t)V * E2, s));
        TACO_TRY(launch_scatter_add_rows(m.W("enc_cbhg/d_xin_p"), b->inputs, m.W("enc/d_table2"), m.enc.N, m.enc.T, m.enc.Tp, m.enc.PL, E2, V, s));
        TACO_TRY(launch_relu_bwd(m.W("enc/d_table2"), m.W("enc/table2"), m.W("enc/d_t2pre"), (long long)V * E2, s));
        taco_gemm_desc w = gd0(m.W("enc/table1"), m.W("enc/d_t2pre"), m.G("enc_prenet/dense_2/kernel"), E1, E2, V, E1, E2, E2);
        w.transA = 1; w.accumulate = 1;
        TACO_TRY(launch_gemm(&w, 1, prec, s));
        TACO_TRY(launch_colsum(m.W("enc/d_t2pre"), m.G("enc_prenet/dense_2/bias"), V, E2, E2, s));
        taco_gemm_desc e = gd0(m.W("enc/d_t2pre"), m.P("enc_prenet/dense_2/kernel"), m.W("enc/d_table1"), V, E1, E2, E2, E2, E1); e.transB = 1;
        TACO_TRY(launch_gemm(&e, 1, prec, s));
        TACO_TRY(launch_relu_bwd(m.W("enc/d_table1"), m.W("enc/table1"), m.W("enc/d_t1pre"), (long long)V * E1, s));
        w = gd0(m.P("embedding"), m.W("enc/d_t1pre"), m.G("enc_prenet/dense_1/kernel"), E0, E1, V, E0, E1, E1);
        w.transA = 1; w.accumulate = 1;
        TACO_TRY(launch_gemm(&w, 1, prec, s));
        TACO_TRY(launch_colsum(m.W("enc/d_t1pre"), m.G("enc_prenet/dense_1/bias"), V, E1, E1, s));
        e = gd0(m.W("enc/d_t1pre"), m.P("enc_prenet/dense_1/kernel"), m.G("embedding"), V, E0, E1, E1, E1, E0); e.transB = 1; e.accumulate = 1;
        TACO_TRY(launch_gemm(&e, 1, prec, s));
    }
    prof_mark("bwd:end", s);
    return TACO_OK;
}

}  // namespace taco

// =============================================== C ABI ===============================================
using namespace taco;

struct taco_model_s { Model m; };

extern "C" {

const char* taco_last_error(void) { return g_err.c_str(); }
int taco_abi_version(void) { return TACO_ABI_VERSION; }
int64_t taco_launch_count(void) { return g_launch_count; }

int taco_create(taco_model* out, const taco_config* cfg) {
    TACO_REQUIRE(out && cfg, TACO_EINVAL, "taco_create: null argument");
    TACO_REQUIRE(cfg->abi_version == TACO_ABI_VERSION, TACO_EINVAL, "taco_create: ABI version %d != %d", cfg->abi_version, TACO_ABI_VERSION);
    TACO_REQUIRE(cfg->device >= 0 && cfg->device < kMaxDevices, TACO_EINVAL, "taco_create: device ordinal %d out of range", cfg->device);
    TACO_REQUIRE(cfg->precision >= TACO_PREC_FP32 && cfg->precision <= TACO_PREC_BF16, TACO_EINVAL, "taco_create: unknown precision %d", cfg->precision);
    TACO_REQUIRE(cfg->attention_type >= 0 && cfg->attention_type <= 2, TACO_EINVAL, " [!] Unkown attention type: %d", cfg->attention_type);
    TACO_REQUIRE(cfg->speaker_mode >= 0 && cfg->speaker_mode <= 3, TACO_EINVAL, " [!] Unkown multi-speaker model type: %d", cfg->speaker_mode);
    TACO_REQUIRE(cfg->reduction_factor >= 1 && cfg->num_mels > 0 && cfg->num_freq > 0, TACO_EINVAL, "taco_create: bad sizes");
    TACO_REQUIRE(cfg->enc_proj_sizes[1] == cfg->enc_prenet_sizes[1], TACO_ESHAPE, "encoder residual needs proj_sizes[-1] == prenet_sizes[-1]");
    TACO_REQUIRE(cfg->post_proj_sizes[1] == cfg->num_mels, TACO_ESHAPE, "post-net residual needs post_proj_sizes[-1] == num_mels");
    taco_model_s* h = new (std::nothrow) taco_model_s();
    TACO_REQUIRE(h, TACO_ENOMEM, "taco_create: out of host memory");
    h->m.cfg = *cfg;
    *out = h;
    return TACO_OK;
}

int taco_destroy(taco_model h) {
    delete h;
    return TACO_OK;
}

int taco_bind_params(taco_model h, const taco_param_entry* table, int32_t n_entries, float* params, float* grads,
                     float* adam_m, float* adam_v, float* bn_state, int64_t n_trainable, int64_t n_state) {
    TACO_REQUIRE(h && table && params && bn_state, TACO_EINVAL, "taco_bind_params: null argument");
    TACO_ON_DEVICE(h->m.cfg.device);
    Model& m = h->m;
    m.table.clear();
    for (int i = 0; i < n_entries; i++) {
        const taco_param_entry& e = table[i];
        TACO_REQUIRE(e.name && e.offset >= 0 && e.numel > 0, TACO_EINVAL, "taco_bind_params: bad entry %d", i);
        TACO_REQUIRE(e.offset + e.numel <= (e.trainable ? n_trainable : n_state), TACO_ESHAPE,
                     "taco_bind_params: '%s' exceeds its buffer", e.name);
        m.table[e.name] = Entry{e.offset, e.numel, e.trainable};
    }
    m.params = params; m.grads = grads; m.adam_m = adam_m; m.adam_v = adam_v; m.bn_state = bn_state;
    m.n_trainable = n_trainable; m.n_state = n_state;
    // data-parallel buckets: the gradients of embedding / speaker / encoder tensors are produced last and sit at the head of the
    // flat buffer (params.py); everything behind them forms the early bucket.  If a caller orders the table differently the
    // early bucket is empty and the whole gradient is reduced after the backward pass.
    {
        auto late = [](const std::string& n) { return n.rfind("embedding", 0) == 0 || n.rfind("speaker", 0) == 0 || n.rfind("enc_", 0) == 0; };
        int64_t first_early = n_trainable, end_late = 0;
        for (const auto& kv : m.table) {
            if (!kv.second.trainable) continue;
            if (late(kv.first)) end_late = std::max<int64_t>(end_late, kv.second.offset + kv.second.numel);
            else first_early = std::min<int64_t>(first_early, kv.second.offset);
        }
        m.early_offset = (first_early >= end_late) ? first_early : n_trainable;
    }
    // every tensor the kernels will dereference must be present with the expected size
    const taco_config& c = m.cfg;
    if (c.precision == TACO_PREC_BF16)
        for (const auto& kv : m.table)
            TACO_REQUIRE(!kv.second.trainable || kv.second.offset % 8 == 0, TACO_EINVAL,
                         "taco_bind_params: bf16 mode needs 16-byte aligned bf16 mirrors: offset of '%s' must be a multiple of 8 elements", kv.first.c_str());
    auto need = [&](const std::string& name, int64_t numel) -> int {
        auto it = m.table.find(name);
        TACO_REQUIRE(it != m.table.end(), TACO_EINVAL, "taco_bind_params: missing tensor '%s'", name.c_str());
        TACO_REQUIRE(it->second.numel == numel, TACO_ESHAPE, "taco_bind_params: '%s' has %lld elements, expected %lld", name.c_str(),
                     (long long)it->second.numel, (long long)numel);
        return TACO_OK;
    };
    TACO_TRY(need("embedding", (int64_t)c.num_symbols * c.embedding_size));
    TACO_TRY(need("enc_prenet/dense_1/kernel", (int64_t)c.embedding_size * c.enc_prenet_sizes[0]));
    TACO_TRY(need("enc_cbhg/bank_1/kernel", (int64_t)c.enc_prenet_sizes[1] * c.enc_bank_channels));
    TACO_TRY(need("enc_cbhg/gru_fw/gates_kernel", (int64_t)2 * c.enc_rnn_size * 2 * c.enc_rnn_size));
    TACO_TRY(need("attention/memory_kernel", (int64_t)2 * c.enc_rnn_size * c.attention_size));
    TACO_TRY(need("mel_proj/kernel", (int64_t)c.dec_rnn_size * c.num_mels * c.reduction_factor));
    TACO_TRY(need("post_cbhg/proj_1/kernel", (int64_t)c.post_proj_width * c.post_bank_size * c.post_bank_channels * c.post_proj_sizes[0]));
    const int64_t spk_cat = (c.speaker_mode == TACO_SPK_SIMPLE) ? c.speaker_embedding_size : 0;
    TACO_TRY(need("linear/kernel", (int64_t)(2 * c.post_rnn_size + spk_cat) * c.num_freq));
    // bank tensors of one kind must be contiguous (see params.py:_cbhg)
    for (const char* pf : {"enc_cbhg", "post_cbhg"}) {
        const int Kb = std::string(pf) == "enc_cbhg" ? c.enc_bank_size : c.post_bank_size;
        const int Cb = std::string(pf) == "enc_cbhg" ? c.enc_bank_channels : c.post_bank_channels;
        for (const char* f : {"bias", "gamma", "beta", "moving_mean", "moving_var"})
            for (int k = 2; k <= Kb; k++) {
                auto a = m.table.find(std::string(pf) + "/bank_" + std::to_string(k - 1) + "/" + f);
                auto b2 = m.table.find(std::string(pf) + "/bank_" + std::to_string(k) + "/" + f);
                TACO_REQUIRE(a != m.table.end() && b2 != m.table.end() && b2->second.offset == a->second.offset + Cb, TACO_ESHAPE,
                             "taco_bind_params: %s bank '%s' tensors must be stored back to back", pf, f);
            }
    }
    return TACO_OK;
}

int taco_workspace_bytes(taco_model h, int32_t N, int32_t T_in, int32_t T_out_or_steps, int32_t training, size_t* bytes) {
    TACO_REQUIRE(h && bytes, TACO_EINVAL, "taco_workspace_bytes: null argument");
    TACO_REQUIRE(N > 0 && T_in > 0 && T_out_or_steps > 0, TACO_ESHAPE, "taco_workspace_bytes: non-positive shape");
    Model& m = h->m;
    Shape s; s.N = N; s.Ti = T_in; s.training = training;
    if (training) { s.To = T_out_or_steps; s.Td = s.To / m.cfg.reduction_factor; }
    else { s.Td = T_out_or_steps; s.To = s.Td * m.cfg.reduction_factor; }
    m.plan(s);
    *bytes = m.plan_bytes;
    return TACO_OK;
}

int taco_bind_workspace(taco_model h, void* ws, size_t bytes) {
    TACO_REQUIRE(h && ws, TACO_EINVAL, "taco_bind_workspace: null argument");
    TACO_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, TACO_EINVAL, "taco_bind_workspace: base must be 256-byte aligned");
    h->m.ws = static_cast<char*>(ws); h->m.ws_bytes = bytes; h->m.tables_ready = false; h->m.prep_done = false;
    return TACO_OK;
}

int taco_ws_region(taco_model h, const char* name, size_t* offset_bytes, int64_t* numel, int64_t dims[4], int64_t strides[4], int32_t* ndim) {
    TACO_REQUIRE(h && name, TACO_EINVAL, "taco_ws_region: null argument");
    auto it = h->m.regions.find(name);
    TACO_REQUIRE(it != h->m.regions.end(), TACO_EINVAL, "taco_ws_region: no region '%s' in the current plan", name);
    const Region& r = it->second;
    if (offset_bytes) *offset_bytes = r.offset;
    if (numel) *numel = r.numel;
    if (ndim) *ndim = r.is_double ? -r.ndim : (r.is_half ? 100 + r.ndim : r.ndim);      // sign / offset encode the element type
    for (int i = 0; i < 4; i++) { if (dims) dims[i] = i < r.ndim ? r.dims[i] : 1; if (strides) strides[i] = i < r.ndim ? r.strides[i] : 0; }
    return TACO_OK;
}

int taco_forward(taco_model h, const taco_batch* b, void* stream) {
    TACO_REQUIRE(h && b, TACO_EINVAL, "taco_forward: null argument");
    TACO_ON_DEVICE(h->m.cfg.device);
    Model& m = h->m;
    TACO_REQUIRE(m.params != nullptr, TACO_ESTATE, "taco_forward: parameters not bound");
    TACO_REQUIRE(b->inputs && b->input_lengths, TACO_EINVAL, "taco_forward: inputs / input_lengths are required");
    Shape s;
    TACO_TRY(shape_of(m, b, s));
    TACO_TRY(ensure_plan(m, s));
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    TACO_TRY(sched_init());
    m.prep_done = false;
    m.img_w_state = 0; m.img_bkm_state = 0;
    if (s.training && overlap_on()) {
        // parameter-only operands of the backward pass are packed on the leaf stream while the forward pass runs
        TACO_CHECK_CUDA(cudaEventRecord(ds().ev_in, user));
        TACO_CHECK_CUDA(cudaStreamWaitEvent(ds().side, ds().ev_in, 0));
        m.img_w_state = 0;
        TACO_TRY(decoder_pack_weight_images(m, ds().side));          // sets img_w_state = 2 (ready behind ds().ev_img[0]) when it applies
        TACO_TRY(backward_prep(m, ds().side));
        TACO_CHECK_CUDA(cudaEventRecord(ds().ev_prep, ds().side));
        m.prep_done = true;
    }
    return model_forward(m, b, user);
}

int taco_backward(taco_model h, const taco_batch* b, void* stream) {
    TACO_REQUIRE(h && b, TACO_EINVAL, "taco_backward: null argument");
    TACO_ON_DEVICE(h->m.cfg.device);
    Model& m = h->m;
    TACO_REQUIRE(m.grads != nullptr, TACO_ESTATE, "taco_backward: gradient buffer not bound");
    Shape s;
    TACO_TRY(shape_of(m, b, s));
    TACO_REQUIRE(m.planned && m.shape == s, TACO_ESTATE, "taco_backward: call taco_forward with the same batch first");
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    TACO_TRY(sched_init());
    if (!overlap_on()) return model_backward(m, b, user);
    // chain on the high-priority stream, leaves on the low-priority one (see "stream scheduler"); both are ordered
    // after everything already enqueued on the caller's stream and joined back into it before returning
    TACO_CHECK_CUDA(cudaEventRecord(ds().ev_in, user));
    TACO_CHECK_CUDA(cudaStreamWaitEvent(ds().crit, ds().ev_in, 0));
    TACO_CHECK_CUDA(cudaStreamWaitEvent(ds().side, ds().ev_in, 0));
    const int rc = model_backward(m, b, ds().crit);
    prof_mark("side:end", ds().side);
    TACO_CHECK_CUDA(cudaEventRecord(ds().ev_side, ds().side));
    TACO_CHECK_CUDA(cudaStreamWaitEvent(user, ds().ev_side, 0));
    TACO_CHECK_CUDA(cudaEventRecord(ds().ev_out, ds().crit));
    TACO_CHECK_CUDA(cudaStreamWaitEvent(user, ds().ev_out, 0));
    return rc;
}

int taco_optimizer_step(taco_model h, int64_t global_step, int64_t adam_step, int32_t is_randomly_initialized, float initial_learning_rate,
                        int32_t decay_mode, float beta1, float beta2, float grad_scale, void* stream) {
    TACO_REQUIRE(h, TACO_EINVAL, "taco_optimizer_step: null model");
    TACO_ON_DEVICE(h->m.cfg.device);
    Model& m = h->m;
    TACO_REQUIRE(m.grads && m.adam_m && m.adam_v, TACO_ESTATE, "taco_optimizer_step: optimizer buffers not bound");
    TACO_REQUIRE(m.planned && m.shape.training, TACO_ESTATE, "taco_optimizer_step: no training step in flight");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // learning rate (tacotron.py:314-326), step = global_step + 1
    const double step = (double)(global_step + 1);
    double lr;
    if (decay_mode == 0) {
        const double w = is_randomly_initialized ? 4000.0 : 40000.0;
        lr = initial_learning_rate * std::sqrt(w) * std::fmin(step * std::pow(w, -1.5), 1.0 / std::sqrt(step));
    } else {
        lr = initial_learning_rate * std::pow(0.95, step / 3000.0);
    }
    const double t = (double)(adam_step + 1);     // AdamOptimizer's own update count (beta powers), not the global step
    const double lr_t = lr * std::sqrt(1.0 - std::pow((double)beta2, t)) / (1.0 - std::pow((double)beta1, t));
    double* sc = m.Wd("scalars");
    float* scf = m.W("scalars_f");
    TACO_CHECK_CUDA(cudaMemsetAsync(sc + 6, 0, sizeof(double), s));
    TACO_TRY(launch_sqnorm(m.grads, m.n_trainable, sc + 6, s));
    TACO_TRY(launch_adam_clip(m.params, m.adam_m, m.adam_v, m.grads, m.n_trainable, sc + 6, grad_scale, 1.0f, (float)lr_t,
                              beta1, beta2, 1e-8f, (float)lr, scf, s));
    // batch-norm moving statistics (the UPDATE_OPS dependency of tacotron.py:332-336)
    for (const CbhgGeom* gp : {&m.enc, &m.post}) {
        const CbhgGeom& g = *gp; const std::string p = g.prefix + "/";
        TACO_TRY(launch_bn_update_moving(m.P(p + "bank_1/moving_mean"), m.P(p + "bank_1/moving_var"), m.W(p + "bank_mean"), m.W(p + "bank_var"), g.Kb * g.Cb, s));
        TACO_TRY(launch_bn_update_moving(m.P(p + "proj_1/moving_mean"), m.P(p + "proj_1/moving_var"), m.W(p + "p1_mean"), m.W(p + "p1_var"), g.P1, s));
        TACO_TRY(launch_bn_update_moving(m.P(p + "proj_2/moving_mean"), m.P(p + "proj_2/moving_var"), m.W(p + "p2_mean"), m.W(p + "p2_var"), g.P2, s));
    }
    return TACO_OK;
}

static void finish_scalars(const Model& m, const double sc[8], const float scf[8], taco_step_scalars* out) {
    const taco_config& c = m.cfg;
    const double cnt_mel = (double)m.shape.N * m.shape.To * c.num_mels, cnt_lin = (double)m.shape.N * m.shape.To * c.num_freq;
    out->loss = (float)(sc[0] + sc[3]);
    out->mel_loss = (float)(sc[1] / cnt_mel);
    if (c.prioritize_loss) {
        const double cnt_band = (double)m.shape.N * m.shape.To * (c.priority_hi - c.priority_lo);
        out->linear_loss = (float)(0.5 * (sc[4] / cnt_lin + sc[5] / cnt_band));
    } else {
        out->linear_loss = (float)(sc[4] / cnt_lin);
    }
    out->loss_without_coeff = out->mel_loss + out->linear_loss;
    out->grad_norm = scf[0];
    out->learning_rate = scf[1];
}

int taco_read_scalars(taco_model h, taco_step_scalars* out, void* stream) {
    TACO_REQUIRE(h && out, TACO_EINVAL, "taco_read_scalars: null argument");
    TACO_ON_DEVICE(h->m.cfg.device);
    Model& m = h->m;
    TACO_REQUIRE(m.planned && m.ws, TACO_ESTATE, "taco_read_scalars: nothing has run");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double sc[8]; float scf[8];
    TACO_CHECK_CUDA(cudaMemcpyAsync(sc, m.Wd("scalars"), sizeof sc, cudaMemcpyDeviceToHost, s));
    TACO_CHECK_CUDA(cudaMemcpyAsync(scf, m.W("scalars_f"), sizeof scf, cudaMemcpyDeviceToHost, s));
    TACO_CHECK_CUDA(cudaStreamSynchronize(s));
    finish_scalars(m, sc, scf, out);
    return TACO_OK;
}

int taco_copy_scalars_async(taco_model h, void* pinned_raw, void* stream) {
    TACO_REQUIRE(h && pinned_raw, TACO_EINVAL, "taco_copy_scalars_async: null argument");
    TACO_ON_DEVICE(h->m.cfg.device);
    Model& m = h->m;
    TACO_REQUIRE(m.planned && m.ws, TACO_ESTATE, "taco_copy_scalars_async: nothing has run");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* raw = static_cast<char*>(pinned_raw);
    static_assert(8 * sizeof(double) + 8 * sizeof(float) <= TACO_SCALARS_RAW_BYTES, "raw scalar block");
    TACO_CHECK_CUDA(cudaMemcpyAsync(raw, m.Wd("scalars"), 8 * sizeof(double), cudaMemcpyDeviceToHost, s));
    TACO_CHECK_CUDA(cudaMemcpyAsync(raw + 8 * sizeof(double), m.W("scalars_f"), 8 * sizeof(float), cudaMemcpyDeviceToHost, s));
    return TACO_OK;
}

int taco_finish_scalars(taco_model h, const void* pinned_raw, taco_step_scalars* out) {
    TACO_REQUIRE(h && pinned_raw && out, TACO_EINVAL, "taco_finish_scalars: null argument");
    const char* raw = static_cast<const char*>(pinned_raw);
    double sc[8]; float scf[8];
    memcpy(sc, raw, sizeof sc); memcpy(scf, raw + sizeof sc, sizeof scf);
    finish_scalars(h->m, sc, scf, out);
    return TACO_OK;
}


// ---- block-level entry points: the same code paths the model runs, reachable one block at a time --------------
static int block_geom(Model& m, int which, int N, int T, const CbhgGeom** g) {
    TACO_REQUIRE(which == 0 || which == 1, TACO_EINVAL, "cbhg: which must be 0 (encoder) or 1 (post-net)");
    TACO_REQUIRE(m.planned && m.ws, TACO_ESTATE, "cbhg: size and bind a workspace first (taco_workspace_bytes / taco_bind_workspace)");
    const CbhgGeom& gg = which ? m.post : m.enc;
    TACO_REQUIRE(gg.N == N && gg.T == T, TACO_ESHAPE, "cbhg: the workspace is planned for N=%d T=%d, not N=%d T=%d", gg.N, gg.T, N, T);
    *g = &gg;
    return TACO_OK;
}

int taco_cbhg_forward(taco_model h, int32_t which, const float* inputs, const int32_t* input_lengths, const float* before_highway,
                      const float* rnn_init_state, int32_t N, int32_t T, int32_t is_training, float* outputs, void* stream) {
    TACO_REQUIRE(h && inputs && outputs, TACO_EINVAL, "taco_cbhg_forward: null argument");
    Model& m = h->m;
    TACO_ON_DEVICE(m.cfg.device);
    const CbhgGeom* gp = nullptr;
    TACO_TRY(block_geom(m, which, N, T, &gp));
    const CbhgGeom& g = *gp;
    TACO_REQUIRE(!is_training || m.shape.training, TACO_ESTATE, "taco_cbhg_forward: training statistics need a training plan");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TACO_TRY(sched_init());
    TACO_TRY(refresh_params16(m, s));
    const std::string px = g.prefix + "/";
    float* xin = m.W(px + "xin_p");
    // [N,T,Cin] -> zero-padded time layout [N,Tp,Cin] (pad frames are never written: they stay zero)
    TACO_TRY(launch_copy2d(xin + (long long)g.PL * g.Cin, inputs, N, T * g.Cin, (long long)g.Tp * g.Cin, (long long)T * g.Cin, s));
    if (m.use16()) TACO_TRY(launch_cast2d_bf16(static_cast<uint16_t*>(m.W16(px + "xin_p")) + (long long)g.PL * g.Cin, inputs, N, T * g.Cin,
                                               (long long)g.Tp * g.Cin, (long long)T * g.Cin, s));
    TACO_TRY(cbhg_forward(m, g, input_lengths, before_highway, rnn_init_state, is_training, s));
    TACO_CHECK_CUDA(cudaMemcpyAsync(outputs, m.W(px + "rnn_out"), sizeof(float) * (size_t)N * T * 2 * g.H, cudaMemcpyDeviceToDevice, s));
    return TACO_OK;
}

int taco_cbhg_backward(taco_model h, int32_t which, const float* d_outputs, const int32_t* input_lengths, int32_t N, int32_t T,
                       float* d_inputs, float* d_before_highway, float* d_rnn_init_state, void* stream) {
    TACO_REQUIRE(h && d_outputs && d_inputs, TACO_EINVAL, "taco_cbhg_backward: null argument");
    Model& m = h->m;
    TACO_ON_DEVICE(m.cfg.device);
    const CbhgGeom* gp = nullptr;
    TACO_TRY(block_geom(m, which, N, T, &gp));
    const CbhgGeom& g = *gp;
    TACO_REQUIRE(m.shape.training && m.grads, TACO_ESTATE, "taco_cbhg_backward: needs a training plan and a bound gradient buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TACO_TRY(sched_init());
    const std::string px = g.prefix + "/";
    TACO_TRY(backward_prep(m, s));
    TACO_CHECK_CUDA(cudaMemcpyAsync(m.W(px + "d_rnn_out"), d_outputs, sizeof(float) * (size_t)N * T * 2 * g.H, cudaMemcpyDeviceToDevice, s));
    TACO_TRY(cbhg_backward(m, g, input_lengths, d_before_highway != nullptr, d_rnn_init_state != nullptr, s));
    TACO_TRY(launch_copy2d(d_inputs, m.W(px + "d_xin_p") + (long long)g.PL * g.Cin, N, T * g.Cin, (long long)T * g.Cin, (long long)g.Tp * g.Cin, s));
    if (d_before_highway) TACO_CHECK_CUDA(cudaMemcpyAsync(d_before_highway, m.W(px + "d_before"), sizeof(float) * (size_t)N * g.P2, cudaMemcpyDeviceToDevice, s));
    if (d_rnn_init_state) TACO_CHECK_CUDA(cudaMemcpyAsync(d_rnn_init_state, m.W(px + "d_h0"), sizeof(float) * (size_t)N * 2 * g.H, cudaMemcpyDeviceToDevice, s));
    return TACO_OK;
}

int taco_decoder_forward(taco_model h, const taco_batch* b, const float* memory, void* stream) {
    TACO_REQUIRE(h && b && memory, TACO_EINVAL, "taco_decoder_forward: null argument");
    Model& m = h->m;
    TACO_ON_DEVICE(m.cfg.device);
    Shape sh;
    TACO_TRY(shape_of(m, b, sh));
    TACO_TRY(ensure_plan(m, sh));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TACO_TRY(sched_init());
    m.img_w_state = 0; m.img_bkm_state = 0; m.prep_done = false;
    TACO_TRY(refresh_params16(m, s));
    TACO_TRY(speaker_forward(m, b, s));
    TACO_CHECK_CUDA(cudaMemcpyAsync(m.W("enc_cbhg/rnn_out"), memory, sizeof(float) * (size_t)sh.N * sh.Ti * 2 * m.cfg.enc_rnn_size, cudaMemcpyDeviceToDevice, s));
    return decoder_forward(m, b, s);
}

int taco_decoder_backward(taco_model h, const taco_batch* b, const float* d_mel_outputs, float* d_memory, void* stream) {
    TACO_REQUIRE(h && b && d_mel_outputs && d_memory, TACO_EINVAL, "taco_decoder_backward: null argument");
    Model& m = h->m;
    TACO_ON_DEVICE(m.cfg.device);
    Shape sh;
    TACO_TRY(shape_of(m, b, sh));
    TACO_REQUIRE(m.planned && m.shape == sh && sh.training && m.grads, TACO_ESTATE, "taco_decoder_backward: call taco_decoder_forward with the same training batch first");
    TACO_REQUIRE(!b->rnn_decoder_test_mode, TACO_ESTATE, "backward through the free-running decoder is not defined");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const CbhgGeom& g = m.post;
    const int M = m.cfg.num_mels;
    TACO_TRY(backward_prep(m, s));
    // gradient wrt the mel outputs, in the post-net's padded input layout (pad frames stay zero)
    TACO_TRY(launch_copy2d(m.W("post_cbhg/d_xin_p") + (long long)g.PL * M, d_mel_outputs, sh.N, sh.To * M, (long long)g.Tp * M, (long long)sh.To * M, s));
    TACO_TRY(decoder_backward(m, b, s));
    TACO_CHECK_CUDA(cudaMemcpyAsync(d_memory, m.W("enc_cbhg/d_rnn_out"), sizeof(float) * (size_t)sh.N * sh.Ti * 2 * m.cfg.enc_rnn_size, cudaMemcpyDeviceToDevice, s));
    return TACO_OK;
}

// y = H*T + x*(1-T) on [rows, C] fp32 matrices (the element-wise half of highwaynet, modules.py:105-120)
int taco_highway_combine(const float* H, const float* T, const float* x, float* y, int64_t rows, int32_t C, void* stream) {
    TACO_REQUIRE(H && T && x && y && rows > 0 && C > 0 && C % 4 == 0, TACO_EINVAL, "taco_highway_combine: bad argument (C must be a multiple of 4)");
    return launch_highway_fwd(H, T, x, y, (long long)rows * C, static_cast<cudaStream_t>(stream));
}

// tf.layers.batch_normalization over the last axis of x [N,T,C] (eps 1e-3; modules.py:131): training: biased batch moments over all
// N*T frames (written to batch_mean / batch_var); otherwise the given moving statistics
int taco_batch_norm(const float* x, const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                    int32_t N, int32_t T, int32_t C, int32_t is_training, float* y, float* batch_mean, float* batch_var, void* scratch, void* stream) {
    TACO_REQUIRE(x && gamma && beta && moving_mean && moving_var && y && scratch && N > 0 && T > 0 && C > 0 && C % 4 == 0, TACO_EINVAL,
                 "taco_batch_norm: bad argument (C must be a multiple of 4; scratch: 4*C doubles)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* st = static_cast<double*>(scratch);                    // [2C] sums, then [C] mean | [C] rstd | [C] var as floats behind them
    float* mean = reinterpret_cast<float*>(st + 2 * C); float* rstd = mean + C; float* var = rstd + C;
    if (is_training) {
        // statistics through the GEMM-free path: x . I is not needed - a [rows, C] column sum / sum of squares in double
        TACO_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(double) * 2 * C, s));
        TACO_TRY(launch_colstats(x, st, st + C, (long long)N * T, C, s));
    }
    TACO_TRY(launch_bn_finalize(st, st + C, (double)N * T, mean, rstd, var, moving_mean, moving_var, C, is_training, s));
    TACO_TRY(launch_bn_apply(x, mean, rstd, gamma, beta, nullptr, nullptr, y, N, T, T, 0, C, 0, s));
    if (is_training && batch_mean) TACO_CHECK_CUDA(cudaMemcpyAsync(batch_mean, mean, sizeof(float) * C, cudaMemcpyDeviceToDevice, s));
    if (is_training && batch_var) TACO_CHECK_CUDA(cudaMemcpyAsync(batch_var, var, sizeof(float) * C, cudaMemcpyDeviceToDevice, s));
    return TACO_OK;
}

int taco_dp_bucket(taco_model h, int32_t bucket, int64_t* offset, int64_t* numel) {
    TACO_REQUIRE(h && offset && numel && (bucket == 0 || bucket == 1), TACO_EINVAL, "taco_dp_bucket: bad argument");
    const Model& m = h->m;
    TACO_REQUIRE(m.grads != nullptr, TACO_ESTATE, "taco_dp_bucket: parameters not bound");
    if (bucket == 0) { *offset = m.early_offset; *numel = m.n_trainable - m.early_offset; }
    else { *offset = 0; *numel = m.early_offset; }
    return TACO_OK;
}

int taco_dp_wait_bucket(taco_model h, int32_t bucket, void* stream) {
    TACO_REQUIRE(h && (bucket == 0 || bucket == 1), TACO_EINVAL, "taco_dp_wait_bucket: bad argument");
    TACO_ON_DEVICE(h->m.cfg.device);
    if (bucket == 1) return TACO_OK;          // complete once taco_backward has joined its streams into the caller's
    TACO_REQUIRE(h->m.early_recorded, TACO_ESTATE, "taco_dp_wait_bucket: call taco_backward first");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TACO_CHECK_CUDA(cudaStreamWaitEvent(s, ds().ev_early[0], 0));
    TACO_CHECK_CUDA(cudaStreamWaitEvent(s, ds().ev_early[1], 0));
    return TACO_OK;
}

int taco_profile(int32_t enable, double ms_out[4], int64_t count_out[4]) {
    // enable=1: start collecting; enable=0: stop, synchronise and return per-class totals (0 GEMM, 1 GRU recurrences, 2 attention recurrences);
    // slot 3 carries the work of the GEMM spans: ms_out[3] = sum of 2MNK in GFLOP, count_out[3] = number of GEMM problems
    if (enable) {
        for (auto& sp : g_prof_spans) { g_prof_pool.push_back(sp.a); if (sp.b != sp.a) g_prof_pool.push_back(sp.b); }
        g_prof_spans.clear(); g_prof_on = true;
        return TACO_OK;
    }
    g_prof_on = false;
    TACO_CHECK_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 4; i++) { if (ms_out) ms_out[i] = 0.0; if (count_out) count_out[i] = 0; }
    for (auto& sp : g_prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.cls >= 0 && sp.cls < 4) {
            if (ms_out) ms_out[sp.cls] += ms;
            if (count_out) count_out[sp.cls] += 1;
            if (sp.cls == 0) { if (ms_out) ms_out[3] += sp.gflop; if (count_out) count_out[3] += sp.tag[3]; }
        }
    }
    return TACO_OK;
}

// debug: per-span listing of the last profile window, one line per span:
//   "cls ms tag0 tag1 tag2 tag3 start_ms stream name"   (GEMM spans: M N K n_problems; start relative to the first span;
//   stream 1 = the low-priority leaf stream, 2 / 3 = the wavefront streams of GRU layer 1 / 2; class 9 = stage markers)
int taco_debug_profile_spans(char* buf, int64_t cap) {
    std::string out;
    for (auto& sp : g_prof_spans) {
        float ms = 0.f, t0 = 0.f;
        if (sp.b != sp.a && cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&t0, g_prof_spans.front().a, sp.a) != cudaSuccess) t0 = -1.f;
        char line[192];
        snprintf(line, sizeof line, "%d %.4f %d %d %d %d %.4f %d %s\n", sp.cls, ms, sp.tag[0], sp.tag[1], sp.tag[2], sp.tag[3], t0,
                 sp.stream == ds().side ? 1 : (ds().wave_ready && sp.stream == ds().wave.w[0]) ? 2 : (ds().wave_ready && sp.stream == ds().wave.w[1]) ? 3 : 0,
                 sp.name[0] ? sp.name : "-");
        out += line;
    }
    if (buf && cap > 0) { snprintf(buf, (size_t)cap, "%s", out.c_str()); }
    return (int)out.size();
}

int taco_gemm(const taco_gemm_desc* d, int32_t n_problems, int32_t precision, void* stream) {
    TACO_REQUIRE(d && n_problems > 0, TACO_EINVAL, "taco_gemm: null argument");
    return launch_gemm(d, n_problems, precision, static_cast<cudaStream_t>(stream));
}

// Debug hook (not part of the ABI header, like the other taco_debug_* entries): number of time chunks of the decoder
// wavefront; <= 1 turns it off.  The default comes from TACO_DEC_CHUNKS (4).  Returns the previous value.
int taco_debug_set_dec_chunks(int32_t n) {
    const int prev = taco::g_wave_chunks;
    taco::g_wave_chunks = n > taco::WaveCtx::kMaxChunks ? taco::WaveCtx::kMaxChunks : n;
    return prev;
}

}  // extern "C"
