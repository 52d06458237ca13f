// Cluster-persistent GRU recurrence (forward and BPTT) for the three GRU sites of the hot path:
//   encoder CBHG bi-GRU, length-aware      reference: models/modules.py:82-96  
//   post-net CBHG bi-GRU, full length      reference: models/tacotron.py:219-224 -> modules.py:82-96 (P1)
//   decoder ResidualWrapper(GRUCell) x2    reference: models/tacotron.py:171-175 (D8; teacher-forced training)
//
// TF r1.4 GRUCell semantics (NOT cuDNN's): [r,u] = sigmoid([x,h].Wg + bg); c = tanh([x, r*h].Wc + bc);
// h' = u*h + (1-u)*c.  The x-side products (and biases) are hoisted into one GEMM over all time steps; this
// kernel runs only the serial part.  bidirectional_dynamic_rnn(sequence_length=L): steps >= L emit 0 and keep
// the state; the backward direction walks t = L-1-s.
//
// Mapping: one thread-block cluster of 8 CTAs owns 8 batch rows of one direction for the whole sequence.
// CTA `rank` keeps the recurrent weight columns of its H/8 hidden units resident in shared memory for all T
// steps; per step the cluster exchanges r*h and h' slices through distributed shared memory (coalesced
// 1 KB pushes) and synchronises with two cluster barriers.  Hidden state and all accumulation are fp32.
#include "common.cuh"
#include "kernels.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace taco {

constexpr int GRU_C = 8;     // CTAs per cluster
constexpr int GRU_R = 8;     // batch rows per cluster
constexpr int GRU_NT = 256;  // threads per CTA

template <int H>
struct GruCfg {
    static constexpr int U = H / GRU_C;         // hidden units owned by one CTA
    static constexpr int GC = 2 * U;            // gate columns (r|u) owned by one CTA
    static constexpr int KS_G = GRU_NT / GC;    // k-slices, gate phase
    static constexpr int KL_G = H / KS_G;
    static constexpr int KS_C = GRU_NT / U;     // k-slices, candidate phase
    static constexpr int KL_C = H / KS_C;
    static constexpr int KL_GT = 2 * H / KS_C;  // backward: K = 2H over U columns
    static constexpr int ACT = U * GRU_R;       // threads that own one (unit,row) state element
    static constexpr size_t smem_floats = (size_t)H * GC + (size_t)H * U + (size_t)H * GRU_R + (size_t)2 * H * GRU_R +
                                          (size_t)GRU_NT * GRU_R + (size_t)2 * ACT;
    static_assert(H % GRU_C == 0 && GRU_NT % GC == 0 && H % KS_G == 0 && H % KS_C == 0 && ACT <= GRU_NT, "bad GRU shape");
};

// acc[r] = sum_{k in [k0,k0+KL)} W_s[k*ld + col] * v_s[k*R + r]
template <int KL>
__device__ __forceinline__ void slice_matvec(const float* __restrict__ W_s, int ld, int col, const float* __restrict__ v_s,
                                             int k0, float acc[GRU_R]) {
#pragma unroll
    for (int r = 0; r < GRU_R; r++) acc[r] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < KL; kk++) {
        const int k = k0 + kk;
        const float w = W_s[k * ld + col];
        const float4 a = *reinterpret_cast<const float4*>(v_s + k * GRU_R);
        const float4 b = *reinterpret_cast<const float4*>(v_s + k * GRU_R + 4);
        acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
        acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
    }
}

// Push this CTA's [U][R] block (staged locally) into every cluster peer's vector buffer at unit offset rank*U.
template <int ACT>
__device__ __forceinline__ void cluster_push(cg::cluster_group& cluster, float* vec_s, const float* stage, int rank, int tid) {
    constexpr int Q = ACT / 4;
    for (int idx = tid; idx < GRU_C * Q; idx += GRU_NT) {
        const int peer = idx / Q, q = idx % Q;
        float* dst = cluster.map_shared_rank(vec_s, peer) + rank * ACT;
        reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(stage)[q];
    }
}

template <int H>
__global__ void __launch_bounds__(GRU_NT, 1) gru_fwd_kernel(const GruArgs a) {
    using Cfg = GruCfg<H>;
    constexpr int U = Cfg::U, GC = Cfg::GC, R = GRU_R;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / GRU_C;
    const int d = cid % a.ndir, grp = cid / a.ndir;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) float smem[];
    float* Wg_s = smem;                      // [H][GC]
    float* Wc_s = Wg_s + H * GC;             // [H][U]
    float* h_s = Wc_s + H * U;               // [H][R]   full hidden vector (unit-major)
    float* rh_s = h_s + H * R;               // [H][R]   r*h
    float* spare = rh_s + H * R;             // [H][R]   (unused in fwd; keeps layout equal to bwd)
    float* red = spare + H * R;              // [R][NT]
    float* stage = red + GRU_NT * R;         // [2][ACT]

    const float* __restrict__ Wg = a.Wg[d];
    const float* __restrict__ Wc = a.Wc[d];
    for (int idx = tid; idx < H * GC; idx += GRU_NT) {
        int k = idx / GC, col = idx % GC;
        int gcol = (col < U) ? rank * U + col : H + rank * U + (col - U);
        Wg_s[idx] = __ldg(Wg + (long long)k * 2 * H + gcol);
    }
    for (int idx = tid; idx < H * U; idx += GRU_NT) {
        int k = idx / U, col = idx % U;
        Wc_s[idx] = __ldg(Wc + (long long)k * H + rank * U + col);
    }
    for (int idx = tid; idx < H * R; idx += GRU_NT) {
        int k = idx / R, r = idx % R, n = grp * R + r;
        h_s[idx] = (a.h0 && n < a.N) ? a.h0[(long long)n * a.ndir * H + d * H + k] : 0.f;
    }
    // per-thread state element (unit i of this CTA, batch row r)
    const bool act = tid < Cfg::ACT;
    const int i = tid % U, r = tid / U;
    const int n = grp * R + r;
    int L = 0;
    if (act && n < a.N) L = a.lengths ? min(max(a.lengths[n], 0), a.T) : a.T;
    int Lmax = 0;
    for (int rr = 0; rr < R; rr++) {
        int nn = grp * R + rr;
        if (nn < a.N) Lmax = max(Lmax, a.lengths ? min(max(a.lengths[nn], 0), a.T) : a.T);
    }
    const int unit = rank * U + i;
    float h_own = (act && a.h0 && n < a.N) ? a.h0[(long long)n * a.ndir * H + d * H + unit] : 0.f;
    const long long st_base = ((long long)d * a.N + n) * a.T;
    __syncthreads();
    cluster.sync();

    const int colG = tid % GC, ksG = tid / GC;
    const int colC = tid % U, ksC = tid / U;

    for (int s = 0; s < Lmax; s++) {
        const bool valid = act && (s < L);
        const int t = (d == 0) ? s : (L - 1 - s);
        float gr = 0.f, gu = 0.f, gc = 0.f;
        if (valid) {
            const float* g = a.gx + ((long long)n * a.gx_rs_n + t + a.gx_row0) * a.gx_ld + (long long)d * 3 * H + unit;
            gr = __ldg(g); gu = __ldg(g + H); gc = __ldg(g + 2 * H);
        }
        // ---- gate phase: [r|u] columns of this CTA -------------------------------------
        {
            float acc[R];
            slice_matvec<Cfg::KL_G>(Wg_s, GC, colG, h_s, ksG * Cfg::KL_G, acc);
#pragma unroll
            for (int rr = 0; rr < R; rr++) red[rr * GRU_NT + tid] = acc[rr];
        }
        __syncthreads();
        float rg = 0.f, ug = 0.f;
        if (act) {
            float sr = gr, su = gu;
#pragma unroll
            for (int ks = 0; ks < Cfg::KS_G; ks++) {
                sr += red[r * GRU_NT + ks * GC + i];
                su += red[r * GRU_NT + ks * GC + U + i];
            }
            rg = sigmoidf_(sr); ug = sigmoidf_(su);
            stage[i * R + r] = rg * h_own;
        }
        __syncthreads();
        cluster_push<Cfg::ACT>(cluster, rh_s, stage, rank, tid);
        cluster.sync();
        // ---- candidate phase ---------------------------------------------------------------
        {
            float acc[R];
            slice_matvec<Cfg::KL_C>(Wc_s, U, colC, rh_s, ksC * Cfg::KL_C, acc);
#pragma unroll
            for (int rr = 0; rr < R; rr++) red[rr * GRU_NT + tid] = acc[rr];
        }
        __syncthreads();
        if (act) {
            float sc = gc;
#pragma unroll
            for (int ks = 0; ks < Cfg::KS_C; ks++) sc += red[r * GRU_NT + ks * U + i];
            const float c = tanhf(sc);
            const float hn = ug * h_own + (1.f - ug) * c;
            if (valid) {
                const long long o = (long long)n * a.T + t;
                float y = hn;
                if (a.res) y += a.res[o * a.res_ld + unit];
                a.out[o * a.out_ld + d * H + unit] = y;
                if (a.st_r) {
                    const long long so = (st_base + t) * H + unit;
                    a.st_r[so] = rg; a.st_u[so] = ug; a.st_c[so] = c; a.st_hprev[so] = h_own;
                }
                h_own = hn;
            }
            stage[Cfg::ACT + i * R + r] = h_own;
        }
        __syncthreads();
        cluster_push<Cfg::ACT>(cluster, h_s, stage + Cfg::ACT, rank, tid);
        cluster.sync();
    }
    if (act && a.hfinal && n < a.N) a.hfinal[(long long)n * a.ndir * H + d * H + unit] = h_own;
}

// BPTT.  Consumes the stash written by the forward kernel; writes grads wrt the x-side pre-activations
// (dgx, same layout as gx) and overwrites st_r with r*h_prev (the operand of the candidate-weight gradient GEMM).
template <int H>
__global__ void __launch_bounds__(GRU_NT, 1) gru_bwd_kernel(const GruArgs a) {
    using Cfg = GruCfg<H>;
    constexpr int U = Cfg::U, R = GRU_R;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / GRU_C;
    const int d = cid % a.ndir, grp = cid / a.ndir;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) float smem[];
    float* WgT_s = smem;                     // [2H][U]   WgT_s[gc*U+i] = Wg[unit_i][gc]
    float* WcT_s = WgT_s + 2 * H * U;        // [H][U]    WcT_s[cu*U+i] = Wc[unit_i][cu]
    float* dcp_s = WcT_s + H * U;            // [H][R]    dc_pre of all units
    float* dg_s = dcp_s + H * R;             // [2H][R]   [dr_pre ; du_pre] of all units
    float* red = dg_s + 2 * H * R;           // [R][NT]
    float* stage = red + GRU_NT * R;         // [2][ACT]

    const float* __restrict__ Wg = a.Wg[d];
    const float* __restrict__ Wc = a.Wc[d];
    for (int idx = tid; idx < 2 * H * U; idx += GRU_NT) {
        int ii = idx / (2 * H), gcol = idx % (2 * H);
        WgT_s[gcol * U + ii] = __ldg(Wg + (long long)(rank * U + ii) * 2 * H + gcol);
    }
    for (int idx = tid; idx < H * U; idx += GRU_NT) {
        int ii = idx / H, cu = idx % H;
        WcT_s[cu * U + ii] = __ldg(Wc + (long long)(rank * U + ii) * H + cu);
    }
    const bool act = tid < Cfg::ACT;
    const int i = tid % U, r = tid / U;
    const int n = grp * R + r;
    int L = 0;
    if (act && n < a.N) L = a.lengths ? min(max(a.lengths[n], 0), a.T) : a.T;
    int Lmax = 0;
    for (int rr = 0; rr < R; rr++) {
        int nn = grp * R + rr;
        if (nn < a.N) Lmax = max(Lmax, a.lengths ? min(max(a.lengths[nn], 0), a.T) : a.T);
    }
    const int unit = rank * U + i;
    const long long st_base = ((long long)d * a.N + n) * a.T;
    float dh_carry = 0.f;
    __syncthreads();
    cluster.sync();

    const int colC = tid % U, ksC = tid / U;

    for (int s = Lmax - 1; s >= 0; s--) {
        const bool valid = act && (s < L);
        const int t = (d == 0) ? s : (L - 1 - s);
        float rg = 0.f, ug = 0.f, c = 0.f, hp = 0.f, dh = dh_carry;
        long long so = 0;
        if (valid) {
            so = (st_base + t) * H + unit;
            rg = a.st_r[so]; ug = a.st_u[so]; c = a.st_c[so]; hp = a.st_hprev[so];
            dh += a.dout[((long long)n * a.T + t) * a.dout_ld + d * H + unit];
        }
        float du_pre = 0.f, dc_pre = 0.f;
        if (valid) {
            du_pre = dh * (hp - c) * ug * (1.f - ug);
            dc_pre = dh * (1.f - ug) * (1.f - c * c);
        }
        if (act) stage[i * R + r] = dc_pre;
        __syncthreads();
        cluster_push<Cfg::ACT>(cluster, dcp_s, stage, rank, tid);
        cluster.sync();
        // d(r*h)[own units] = sum_cu dc_pre[cu] * Wc[unit][cu]
        {
            float acc[R];
            slice_matvec<Cfg::KL_C>(WcT_s, U, colC, dcp_s, ksC * Cfg::KL_C, acc);
#pragma unroll
            for (int rr = 0; rr < R; rr++) red[rr * GRU_NT + tid] = acc[rr];
        }
        __syncthreads();
        float d_rh = 0.f, dr_pre = 0.f;
        if (act) {
#pragma unroll
            for (int ks = 0; ks < Cfg::KS_C; ks++) d_rh += red[r * GRU_NT + ks * U + i];
            if (valid) dr_pre = d_rh * hp * rg * (1.f - rg);
            stage[i * R + r] = dr_pre;
            stage[Cfg::ACT + i * R + r] = du_pre;
        }
        __syncthreads();
        cluster_push<Cfg::ACT>(cluster, dg_s, stage, rank, tid);
        cluster_push<Cfg::ACT>(cluster, dg_s + H * R, stage + Cfg::ACT, rank, tid);
        cluster.sync();
        // dh_prev += sum_gc [dr_pre;du_pre][gc] * Wg[unit][gc]
        {
            float acc[R];
            slice_matvec<Cfg::KL_GT>(WgT_s, U, colC, dg_s, ksC * Cfg::KL_GT, acc);
#pragma unroll
            for (int rr = 0; rr < R; rr++) red[rr * GRU_NT + tid] = acc[rr];
        }
        __syncthreads();
        if (act) {
            float sg = 0.f;
#pragma unroll
            for (int ks = 0; ks < Cfg::KS_C; ks++) sg += red[r * GRU_NT + ks * U + i];
            if (valid) {
                dh_carry = dh * ug + d_rh * rg + sg;
                float* g = a.dgx + ((long long)n * a.gx_rs_n + t + a.gx_row0) * a.gx_ld + (long long)d * 3 * H + unit;
                g[0] = dr_pre; g[H] = du_pre; g[2 * H] = dc_pre;
                a.st_r[so] = rg * hp;
            }
        }
        // `red`/`stage` are rewritten only after the next step's barriers; dcp_s/dg_s hazards are covered by
        // the two cluster barriers of the next step ("GRU exchange protocol").
        __syncthreads();
    }
    if (act && a.dh0 && n < a.N) a.dh0[(long long)n * a.ndir * H + d * H + unit] = dh_carry;
}

template <int H>
static int launch_gru_t(const GruArgs& a, bool bwd, cudaStream_t s) {
    using Cfg = GruCfg<H>;
    const size_t smem = Cfg::smem_floats * sizeof(float);
    auto kern = bwd ? gru_bwd_kernel<H> : gru_fwd_kernel<H>;
    static bool configured[2] = {false, false};
    if (!configured[bwd ? 1 : 0]) {
        TACO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[bwd ? 1 : 0] = true;
    }
    const int groups = cdiv(a.N, GRU_R);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(GRU_C * a.ndir * groups);
    cfg.blockDim = dim3(GRU_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GRU_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TACO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    g_launch_count++;
    return TACO_OK;
}

static int check_gru_args(const GruArgs& a, bool bwd) {
    TACO_REQUIRE(a.N > 0 && a.T > 0, TACO_ESHAPE, "gru: empty batch N=%d T=%d", a.N, a.T);
    TACO_REQUIRE(a.ndir == 1 || a.ndir == 2, TACO_EINVAL, "gru: ndir must be 1 or 2");
    TACO_REQUIRE(a.H == 128 || a.H == 256, TACO_ESHAPE, "gru: hidden size %d not instantiated (128, 256)", a.H);
    TACO_REQUIRE(a.gx && a.Wg[0] && a.Wc[0], TACO_EINVAL, "gru: null operand");
    if (bwd) TACO_REQUIRE(a.dout && (a.dgx || (a.fast && a.dgx16)) && a.st_r && a.st_u && a.st_c && a.st_hprev, TACO_EINVAL, "gru bwd: missing stash/grad buffers");
    else TACO_REQUIRE(a.out, TACO_EINVAL, "gru fwd: null output");
    TACO_REQUIRE(a.fast || !(a.out16 || a.st_hprev16 || a.dgx16 || a.dgx16_dense || a.st_rh16), TACO_EINVAL, "gru: bf16 mirrors are written by the fast kernels only");
    return TACO_OK;
}

int launch_gru_fwd(const GruArgs& a, cudaStream_t s) {
    TACO_TRY(check_gru_args(a, false));
    if (a.fast) return launch_gru_fast_fwd(a, s);
    TACO_REQUIRE(a.t_begin == 0 && a.t_end == 0 && !a.dh_in, TACO_EINVAL, "gru: the exact kernels do not run time chunks");
    return a.H == 128 ? launch_gru_t<128>(a, false, s) : launch_gru_t<256>(a, false, s);
}
int launch_gru_bwd(const GruArgs& a, cudaStream_t s) {
    TACO_TRY(check_gru_args(a, true));
    if (a.fast) return launch_gru_fast_bwd(a, s);
    TACO_REQUIRE(a.t_begin == 0 && a.t_end == 0 && !a.dh_in, TACO_EINVAL, "gru: the exact kernels do not run time chunks");
    return a.H == 128 ? launch_gru_t<128>(a, true, s) : launch_gru_t<256>(a, true, s);
}

}  // namespace taco
