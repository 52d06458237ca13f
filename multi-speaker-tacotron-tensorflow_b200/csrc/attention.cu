// Cluster-persistent attention recurrence of the decoder (teacher-forced path), forward and BPTT.
//   reference: models/rnn_wrappers.py:218-341 (AttentionWrapper.call + _compute_attention),
//              :367-378 (DecoderPrenetWrapper), :405-415 (ConcatOutputAndAttentionWrapper),
//              models/tacotron.py:127-170, TF r1.4 BahdanauMonotonicAttention / BahdanauAttention
//   (SURVEY.md §8a rows D1-D7, Appendix C for the monotonic backward).
//
// Per decoder step t (per batch row):
//   z1 = relu(px[t] + ctx.W1c)            px = x_t.W1[0:M] + b1 hoisted over all t (teacher forcing)
//   z  = relu(z1.W2 + b2) (++ spk)
//   [r,u] = sigmoid([z,ha].Wg + bg); c = tanh([z, r*ha].Wc + bc); ha' = u*ha + (1-u)*c        (TF GRUCell)
//   q = ha'.Wq;  e_j = sum_u v_u tanh(keys_ju + q_u) (+ b);  a = monotonic(sigmoid(e), a_prev) | softmax(e) | manual
//   ctx' = sum_j a_j memory_j;  y0 = [ha', ctx' (,spk)].Wo + bo
// The two ResidualWrapper(GRUCell) layers and the mel projection do not feed back in teacher-forced mode
// (helpers.py:60-67), so they run afterwards as hoisted GEMMs + the GRU kernel of gru.cu.
//
// Mapping: a cluster of 8 CTAs owns 8 batch rows; each CTA owns 1/8 of every layer's output units and 1/8 of the
// memory positions for scoring.  Activations are exchanged through distributed shared memory; 7 cluster barriers
// per step.  Weights are streamed from L2 every step with coalesced loads (they do not fit 8 x 227 KB in fp32).
#include "common.cuh"
#include "kernels.h"
#include <cooperative_groups.h>
#include <cfloat>

namespace cg = cooperative_groups;

namespace taco {

constexpr int AT_C = 8, AT_R = 8, AT_NT = 256;

template <bool FAST> __device__ __forceinline__ float tanh_(float x) {
    if (FAST) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    return tanhf(x);
}
template <bool FAST> __device__ __forceinline__ float sigm_(float x) {
    if (FAST) return __fdividef(1.0f, 1.0f + __expf(-x));
    return 1.0f / (1.0f + expf(-x));
}

// acc[r] += sum_{k in [k0,k1)} Wcol[k*ld] * v_s[k*R + r]      (Wcol already points at this thread's column)
// The weights stream from L2 on every call (nothing is resident in these exact kernels), so the loop is bound by how many
// loads a thread keeps in flight: MV_LB weights are fetched back to back before the first is used (with 4 in flight the
// free-running decoder step took 97 us, almost all of it L2 latency).  The accumulation order is unchanged.
constexpr int MV_LB = 16;
__device__ __forceinline__ void mv_fma(float acc[AT_R], float w, const float* __restrict__ v_s, int k) {
    const float4 a = *reinterpret_cast<const float4*>(v_s + k * AT_R);
    const float4 b = *reinterpret_cast<const float4*>(v_s + k * AT_R + 4);
    acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
    acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
}
__device__ __forceinline__ void mv_acc(float acc[AT_R], const float* __restrict__ Wcol, long long ld, const float* __restrict__ v_s,
                                       int k0, int k1) {
    int k = k0;
    for (; k + MV_LB <= k1; k += MV_LB) {
        float w[MV_LB];
#pragma unroll
        for (int u = 0; u < MV_LB; u++) w[u] = __ldg(Wcol + (long long)(k + u) * ld);
#pragma unroll
        for (int u = 0; u < MV_LB; u++) mv_fma(acc, w[u], v_s, k + u);
    }
#pragma unroll 4
    for (; k < k1; k++) mv_fma(acc, __ldg(Wcol + (long long)k * ld), v_s, k);
}
__device__ __forceinline__ void red_store(float* red, int tid, const float acc[AT_R]) {
#pragma unroll
    for (int r = 0; r < AT_R; r++) red[r * AT_NT + tid] = acc[r];
}
// sum over the k-slices of column `col` for row r (ncols columns per slice row)
__device__ __forceinline__ float red_sum(const float* red, int r, int col, int ncols) {
    float s = 0.f;
    for (int c = col; c < AT_NT; c += ncols) s += red[r * AT_NT + c];
    return s;
}
// push a [U][R] block (unit-major) into every peer's [K][R] vector at unit offset rank*U
__device__ __forceinline__ void push_um(cg::cluster_group& cl, float* vec_s, const float* stage, int rank, int U, int tid) {
    const int Q = U * AT_R / 4;
    for (int idx = tid; idx < AT_C * Q; idx += AT_NT) {
        const int peer = idx / Q, q = idx % Q;
        float* dst = cl.map_shared_rank(vec_s, peer) + rank * U * AT_R;
        reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(stage)[q];
    }
}
// push a [R][U] block (row-major) into every peer's [R][W] matrix at column offset rank*U
__device__ __forceinline__ void push_rm(cg::cluster_group& cl, float* mat_s, int W, const float* stage, int rank, int U, int tid) {
    const int tot = AT_R * U;
    for (int idx = tid; idx < AT_C * tot; idx += AT_NT) {
        const int peer = idx / tot, e = idx % tot;
        const int r = e / U, i = e % U;
        cl.map_shared_rank(mat_s, peer)[r * W + rank * U + i] = stage[e];
    }
}


// ---- fragment-packed bf16 weights of the free-running decoder (inference; tensor-core precision modes) --------------------
// The exact kernels stream fp32 weights from L2 with scalar loads and multiply on the FMA pipe: 75 us per decoder step, bound by
// load latency (4-byte loads in flight) and, at 8 rows, by 25 M FMAs per step.  For synthesis the step's 1.57 M weights are
// packed ONCE per call as bf16 mma.m16n8k16 A fragments, per cluster rank, m-tile (16 output columns) and k-tile: a warp then
// fetches a whole 16x16 weight tile with one coalesced 512-byte load (16 bytes per lane) and multiplies it with the 8 batch
// rows of the cluster on the tensor cores.  Thread/slot layout of the partial sums (`red`) is the one of the scalar path, so
// the activation code of every phase is shared.
constexpr int WF_N = 18;
enum WfId { WF_W1C = 0, WF_W1X, WF_W2, WF_WG_Z, WF_WG_H, WF_WC_Z, WF_WC_H, WF_WQWO, WF_WO_C,
            WF_G1_GX, WF_G1_GH, WF_G1_CX, WF_G1_CH, WF_G2_GX, WF_G2_GH, WF_G2_CX, WF_G2_CH, WF_MEL };
struct WfSpec { const float* WA; const float* WB; int ldA, ldB, UA, UB, offA, offB, split, row0, K, ncols, valid; };
struct WfTable { WfSpec s[WF_N]; long long off[WF_N + 1]; };     // off: start of each image in uint4 units

__device__ __forceinline__ uint32_t wf_pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// image index: ((rank * MT + mt) * nkt + kt) * 32 + lane;  lane (g = lane>>2, t = lane&3) holds A[g][2t..], A[g+8][2t..], A[g][2t+8..], A[g+8][2t+8..]
__global__ void wf_pack_kernel(const __grid_constant__ WfTable tab, uint4* __restrict__ img) {
    const WfSpec& sp = tab.s[blockIdx.y];
    const int MT = sp.ncols / 16, nkt = (sp.K + 15) / 16;
    const long long total = (long long)AT_C * MT * nkt * 32;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(idx & 31); long long q = idx >> 5;
        const int kt = (int)(q % nkt); q /= nkt;
        const int mt = (int)(q % MT), rank = (int)(q / MT);
        const int g = lane >> 2, t = lane & 3;
        auto W = [&](int m, int k) -> float {
            const int col = mt * 16 + m, kk = kt * 16 + k;
            if (col >= sp.valid || kk >= sp.K) return 0.f;
            if (col < sp.split) return __ldg(sp.WA + (long long)(sp.row0 + kk) * sp.ldA + rank * sp.UA + sp.offA + col);
            return __ldg(sp.WB + (long long)(sp.row0 + kk) * sp.ldB + rank * sp.UB + sp.offB + (col - sp.split));
        };
        uint4 o;
        o.x = wf_pack2(W(g, 2 * t), W(g, 2 * t + 1));         o.y = wf_pack2(W(g + 8, 2 * t), W(g + 8, 2 * t + 1));
        o.z = wf_pack2(W(g, 2 * t + 8), W(g, 2 * t + 9));     o.w = wf_pack2(W(g + 8, 2 * t + 8), W(g + 8, 2 * t + 9));
        img[tab.off[blockIdx.y] + idx] = o;
    }
}
__device__ __forceinline__ void wf_mma(float c[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// acc += W_tile(mt; k-tiles ks, ks+KS, ...) . v      img_mt: image of this rank's m-tile ([kt][32] uint4); v_s: fp32 [K][R]
__device__ __forceinline__ void wf_seg(float acc[4], const uint4* __restrict__ img_mt, int nkt, int KS, int ks, const float* __restrict__ v_s, int lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int kt0 = ks; kt0 < nkt; kt0 += 4 * KS) {
        uint4 af[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int kt = kt0 + u * KS; if (kt < nkt) af[u] = __ldg(img_mt + kt * 32 + lane); }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int kt = kt0 + u * KS;
            if (kt < nkt) {
                const float* vp = v_s + (kt * 16 + 2 * t) * AT_R + g;
                wf_mma(acc, af[u], wf_pack2(vp[0], vp[AT_R]), wf_pack2(vp[8 * AT_R], vp[9 * AT_R]));
            }
        }
    }
}
// fragment -> red[r][ks*ncols + col]  (the slot the scalar path's thread (col, ks) writes)
__device__ __forceinline__ void wf_store(float* red, const float acc[4], int ncols, int ks, int mt, int lane) {
    const int g = lane >> 2, t = lane & 3, base = ks * ncols + mt * 16 + g;
    red[(2 * t) * AT_NT + base] = acc[0]; red[(2 * t + 1) * AT_NT + base] = acc[1];
    red[(2 * t) * AT_NT + base + 8] = acc[2]; red[(2 * t + 1) * AT_NT + base + 8] = acc[3];
}

static inline size_t att_fwd_smem_floats(const AttArgs& a, int Tip) {
    const size_t fr = a.free_run ? (size_t)AT_R * (a.M + 6 * a.Y) : 0;
    return fr + (size_t)AT_R * (a.E + a.Z1 + (a.Z + a.SPK) + 2 * a.HA + a.A) + (size_t)4 * AT_R * Tip + (size_t)2 * AT_R * AT_NT + 1024 + a.A;
}
static inline size_t att_bwd_smem_floats(const AttArgs& a, int Tip) {
    return (size_t)AT_R * (a.Y + a.A + a.Z1 + a.Z + 3 * a.HA + a.E) + (size_t)8 * AT_R * Tip + (size_t)2 * AT_R * AT_NT + 1024 + a.A;
}

template <bool FAST>
__global__ void __launch_bounds__(AT_NT, 1) att_fwd_kernel(const AttArgs a) {
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int grp = blockIdx.x / AT_C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = AT_R;
    const int E = a.E, A = a.A, HA = a.HA, Z1 = a.Z1, Z = a.Z, SPK = a.SPK, Y = a.Y, Ti = a.Ti, Td = a.Td;
    const int ZS = Z + SPK;
    const int TJ = (Ti + AT_C - 1) / AT_C, Tip = TJ * AT_C;
    const int Uz1 = Z1 / AT_C, Uz = Z / AT_C, Uh = HA / AT_C, Ua = A / AT_C, Uy = Y / AT_C, Ue = E / AT_C;

    extern __shared__ __align__(16) float smem[];
    float* ctx_s = smem;                    // [E][R]
    float* z1_s = ctx_s + E * R;            // [Z1][R]
    float* z_s = z1_s + Z1 * R;             // [Z+SPK][R]
    float* ha_s = z_s + ZS * R;             // [HA][R]
    float* rha_s = ha_s + HA * R;           // [HA][R]
    float* q_rm = rha_s + HA * R;           // [R][A]
    float* e_s = q_rm + R * A;              // [R][Tip]
    float* a_s = e_s + R * Tip;             // [R][Tip]
    float* p_s = a_s + R * Tip;             // [R][Tip]
    float* cp_s = p_s + R * Tip;            // [R][Tip]
    float* red = cp_s + R * Tip;            // [R][NT]
    float* red2 = red + R * AT_NT;          // [R][NT]
    float* stage = red2 + R * AT_NT;        // [1024]
    float* v_s = stage + 1024;              // [A]
    // free-running mode only
    float* x_s = v_s + A;                   // [M][R]   step input (last frame of the previous step)
    float* y0_s = x_s + (a.free_run ? a.M * R : 0);   // [Y][R]
    float* h1_s = y0_s + Y * R; float* rhd_s = h1_s + Y * R; float* y1_s = rhd_s + Y * R; float* h2_s = y1_s + Y * R; float* y2_s = h2_s + Y * R;

    // ---- init ------------------------------------------------------------------------------
    if (a.att_type == TACO_ATT_BAH_NORM) {
        float ss = 0.f;
        for (int u = 0; u < A; u++) ss += a.v[u] * a.v[u];
        const float sc = a.att_g[0] * rsqrtf(ss);
        for (int u = tid; u < A; u += AT_NT) v_s[u] = a.v[u] * sc;
    } else {
        for (int u = tid; u < A; u += AT_NT) v_s[u] = a.v[u];
    }
    for (int idx = tid; idx < E * R; idx += AT_NT) ctx_s[idx] = 0.f;
    for (int idx = tid; idx < HA * R; idx += AT_NT) {
        int k = idx / R, r = idx % R, n = grp * R + r;
        ha_s[idx] = (a.ha0 && n < a.N) ? a.ha0[(long long)n * HA + k] : 0.f;
    }
    for (int idx = tid; idx < ZS * R; idx += AT_NT) {
        int k = idx / R, r = idx % R, n = grp * R + r;
        z_s[idx] = (k >= Z && a.spk && n < a.N) ? a.spk[(long long)n * SPK + (k - Z)] : 0.f;
    }
    for (int idx = tid; idx < R * Tip; idx += AT_NT) {
        int j = idx % Tip;
        a_s[idx] = (a.att_type == TACO_ATT_BAH_MON && j == 0) ? 1.f : 0.f;
        e_s[idx] = 0.f;
    }
    const float score_bias = (a.att_type == TACO_ATT_BAH_MON) ? a.score_bias[0] : 0.f;
    if (a.free_run) {
        for (int idx = tid; idx < a.M * R; idx += AT_NT) x_s[idx] = 0.f;                    // <GO> frame (helpers.py:70-72)
        for (int idx = tid; idx < Y * R; idx += AT_NT) {
            int k = idx / R, r = idx % R, n = grp * R + r;
            h1_s[idx] = (a.h1_0 && n < a.N) ? a.h1_0[(long long)n * Y + k] : 0.f;
            h2_s[idx] = (a.h2_0 && n < a.N) ? a.h2_0[(long long)n * Y + k] : 0.f;
        }
    }

    // activation-thread coordinates for 32-unit and 16-unit layers
    const int i32 = tid % 32, r32 = tid / 32;            // 256 threads: (unit, row)
    const int n32 = grp * R + r32;
    const bool row_ok32 = n32 < a.N;
    float ha_own = (a.ha0 && row_ok32) ? a.ha0[(long long)n32 * HA + rank * Uh + i32] : 0.f;
    float h1_own = (a.free_run && a.h1_0 && row_ok32) ? a.h1_0[(long long)n32 * Y + rank * Uy + i32] : 0.f;
    float h2_own = (a.free_run && a.h2_0 && row_ok32) ? a.h2_0[(long long)n32 * Y + rank * Uy + i32] : 0.f;
    __syncthreads();
    cl.sync();

    // fragment-packed weights (free-running decoder of the tensor-core modes, see wf_pack_kernel): image of matrix `id`, this rank, m-tile mt
    const bool WFP = FAST && a.wfrag != nullptr;
    auto wf = [&](int id, int MT, int nkt, int mt) { return reinterpret_cast<const uint4*>(a.wfrag) + a.wf_off[id] + ((long long)(rank * MT + mt) * nkt) * 32; };

    for (int t = 0; t < Td; t++) {
        const long long row32 = (long long)n32 * Td + t;
        // ===== P1: z1 (own Uz1 units) = relu(px + ctx.W1c) =====
        if (WFP) {
            for (int task = warp; task < 16; task += AT_NT / 32) {          // 2 m-tiles x 8 k-slices
                const int mt = task & 1, ks = task >> 1;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                wf_seg(acc, wf(WF_W1C, 2, 16, mt), 16, 8, ks, ctx_s, lane);
                wf_seg(acc, wf(WF_W1X, 2, 5, mt), 5, 8, ks, x_s, lane);
                wf_store(red, acc, 32, ks, mt, lane);
            }
        } else {
            const int ncols = Uz1, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = E / KS;
            mv_acc(acc, a.W1c + rank * Uz1 + col, Z1, ctx_s, ks * kl, (ks + 1) * kl);
            if (a.free_run) { const int kx = a.M / KS; mv_acc(acc, a.W1x + rank * Uz1 + col, Z1, x_s, ks * kx, (ks + 1) * kx); }
            red_store(red, tid, acc);
        }
        __syncthreads();
        if (tid < Uz1 * R) {
            const int i = tid % Uz1, r = tid / Uz1, n = grp * R + r;
            float v = 0.f;
            if (n < a.N) {
                const long long row = (long long)n * Td + t;
                const float xb = a.free_run ? __ldg(a.b1 + rank * Uz1 + i) : __ldg(a.px + row * Z1 + rank * Uz1 + i);
                v = fmaxf(red_sum(red, r, i, Uz1) + xb, 0.f);
                if (a.s_z1) {
                    a.s_z1[row * Z1 + rank * Uz1 + i] = v;
                }
            }
            stage[i * R + r] = v;
        }
        if (a.s_ctxin) {   // context consumed by this step (for the hoisted W1c gradient)
            for (int idx = tid; idx < Ue * R; idx += AT_NT) {
                const int i = idx % Ue, r = idx / Ue, n = grp * R + r;
                if (n < a.N) a.s_ctxin[((long long)n * Td + t) * E + rank * Ue + i] = ctx_s[(rank * Ue + i) * R + r];
            }
        }
        __syncthreads();
        push_um(cl, z1_s, stage, rank, Uz1, tid);
        cl.sync();
        // ===== P2: z (own Uz units) = relu(z1.W2 + b2) =====
        if (WFP) {
            for (int task = warp; task < 16; task += AT_NT / 32) {          // 1 m-tile x 16 k-slices
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                wf_seg(acc, wf(WF_W2, 1, 16, 0), 16, 16, task, z1_s, lane);
                wf_store(red, acc, 16, task, 0, lane);
            }
        } else {
            const int ncols = Uz, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = Z1 / KS;
            mv_acc(acc, a.W2 + rank * Uz + col, Z, z1_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        if (tid < Uz * R) {
            const int i = tid % Uz, r = tid / Uz, n = grp * R + r;
            float v = 0.f;
            if (n < a.N) {
                v = fmaxf(red_sum(red, r, i, Uz) + __ldg(a.b2 + rank * Uz + i), 0.f);
                if (a.s_z) a.s_z[((long long)n * Td + t) * Z + rank * Uz + i] = v;
            }
            stage[i * R + r] = v;
        }
        __syncthreads();
        push_um(cl, z_s, stage, rank, Uz, tid);
        cl.sync();
        // ===== P3: gates r|u (own 2*Uh columns) and the z-part of the candidate =====
        if (WFP) {
            for (int task = warp; task < 16; task += AT_NT / 32) {
                {   // gates: 4 m-tiles (r | u columns) x 4 k-slices over [z ; ha]
                    const int mt = task & 3, ks = task >> 2;
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    wf_seg(acc, wf(WF_WG_Z, 4, 8, mt), 8, 4, ks, z_s, lane);
                    wf_seg(acc, wf(WF_WG_H, 4, 16, mt), 16, 4, ks, ha_s, lane);
                    wf_store(red, acc, 64, ks, mt, lane);
                }
                {   // z part of the candidate: 2 m-tiles x 8 k-slices
                    const int mt = task & 1, ks = task >> 1;
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    wf_seg(acc, wf(WF_WC_Z, 2, 8, mt), 8, 8, ks, z_s, lane);
                    wf_store(red2, acc, 32, ks, mt, lane);
                }
            }
        } else {
            {
                const int ncols = 2 * Uh, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
                const int gcol = (col < Uh) ? rank * Uh + col : HA + rank * Uh + (col - Uh);
                float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                int kl = ZS / KS;
                mv_acc(acc, a.Wg + gcol, 2 * HA, z_s, ks * kl, (ks + 1) * kl);
                kl = HA / KS;
                mv_acc(acc, a.Wg + (long long)ZS * 2 * HA + gcol, 2 * HA, ha_s, ks * kl, (ks + 1) * kl);
                red_store(red, tid, acc);
            }
            {
                const int ncols = Uh, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
                float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const int kl = ZS / KS;
                mv_acc(acc, a.Wc + rank * Uh + col, HA, z_s, ks * kl, (ks + 1) * kl);
                red_store(red2, tid, acc);
            }
        }
        __syncthreads();
        float rg = 0.f, ug = 0.f, cz = 0.f;
        {
            const int unit = rank * Uh + i32;
            const float sr = red_sum(red, r32, i32, 2 * Uh) + __ldg(a.bg + unit);
            const float su = red_sum(red, r32, Uh + i32, 2 * Uh) + __ldg(a.bg + HA + unit);
            cz = red_sum(red2, r32, i32, Uh);
            rg = sigm_<FAST>(sr); ug = sigm_<FAST>(su);
            stage[i32 * R + r32] = rg * ha_own;
        }
        __syncthreads();
        push_um(cl, rha_s, stage, rank, Uh, tid);
        cl.sync();
        // ===== P4: candidate, new attention-GRU state =====
        if (WFP) {
            for (int task = warp; task < 16; task += AT_NT / 32) {
                const int mt = task & 1, ks = task >> 1;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                wf_seg(acc, wf(WF_WC_H, 2, 16, mt), 16, 8, ks, rha_s, lane);
                wf_store(red, acc, 32, ks, mt, lane);
            }
        } else {
            const int ncols = Uh, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = HA / KS;
            mv_acc(acc, a.Wc + (long long)ZS * HA + rank * Uh + col, HA, rha_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        {
            const int unit = rank * Uh + i32;
            const float c = tanh_<FAST>(red_sum(red, r32, i32, Uh) + cz + __ldg(a.bc + unit));
            const float hn = ug * ha_own + (1.f - ug) * c;
            if (row_ok32 && a.s_r) {
                const long long o = row32 * HA + unit;
                a.s_r[o] = rg; a.s_u[o] = ug; a.s_c[o] = c; a.s_haprev[o] = ha_own; a.s_ha[o] = hn;
            }
            ha_own = row_ok32 ? hn : 0.f;
            stage[i32 * R + r32] = ha_own;
        }
        __syncthreads();
        push_um(cl, ha_s, stage, rank, Uh, tid);
        cl.sync();
        // ===== P5: query (own Ua columns) and the ha-part of the concat projection (own Uy columns) =====
        if (WFP) {
            for (int task = warp; task < 16; task += AT_NT / 32) {          // 4 m-tiles (Wq | Wo_h columns) x 4 k-slices
                const int mt = task & 3, ks = task >> 2;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                wf_seg(acc, wf(WF_WQWO, 4, 16, mt), 16, 4, ks, ha_s, lane);
                wf_store(red, acc, 64, ks, mt, lane);
            }
        } else {
            const int ncols = Ua + Uy, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = HA / KS;
            if (col < Ua) mv_acc(acc, a.Wq + rank * Ua + col, A, ha_s, ks * kl, (ks + 1) * kl);
            else mv_acc(acc, a.Wo + rank * Uy + (col - Ua), Y, ha_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        float yh = 0.f;
        {
            float q = red_sum(red, r32, i32, Ua + Uy);
            if (a.att_type == TACO_ATT_BAH_NORM) q += __ldg(a.att_b + rank * Ua + i32);
            yh = red_sum(red, r32, Ua + i32, Ua + Uy);
            if (row_ok32 && a.s_q) a.s_q[row32 * A + rank * Ua + i32] = q;
            stage[r32 * Ua + i32] = q;
        }
        __syncthreads();
        push_rm(cl, q_rm, A, stage, rank, Ua, tid);
        cl.sync();
        // ===== P6: scores of the own memory slice: e[r][j] = sum_u v_u tanh(keys[n,j,u] + q[r,u]) =====
        // (row-major pair index: with few valid rows - batch-1 synthesis - the positions of a row still spread over all warps;
        //  the keys of the next pair are fetched while the current one is reduced)
        {
            auto pair_ok = [&](int p) { return p < R * TJ && rank * TJ + (p % TJ) < Ti && grp * R + (p / TJ) < a.N; };
            auto pair_keys = [&](int p) { return a.keys + ((long long)(grp * R + p / TJ) * Ti + rank * TJ + (p % TJ)) * A; };
            float kn[8];
            if (A == 256 && pair_ok(warp)) {
                const float* kp = pair_keys(warp);
#pragma unroll
                for (int q8 = 0; q8 < 8; q8++) kn[q8] = __ldg(kp + lane + 32 * q8);
            }
            for (int p = warp; p < R * TJ; p += AT_NT / 32) {
                const int r = p / TJ, jj = p % TJ;
                const bool ok = pair_ok(p);
                float s = 0.f;
                if (A == 256) {
                    float kv[8];
#pragma unroll
                    for (int q8 = 0; q8 < 8; q8++) kv[q8] = kn[q8];
                    const int pn = p + AT_NT / 32;
                    if (pair_ok(pn)) {
                        const float* kp = pair_keys(pn);
#pragma unroll
                        for (int q8 = 0; q8 < 8; q8++) kn[q8] = __ldg(kp + lane + 32 * q8);
                    }
                    if (ok) {
#pragma unroll
                        for (int q8 = 0; q8 < 8; q8++) { const int u = lane + 32 * q8; s = fmaf(v_s[u], tanh_<FAST>(kv[q8] + q_rm[r * A + u]), s); }
                    }
                } else if (ok) {
                    const float* kp = pair_keys(p);
                    for (int u = lane; u < A; u += 32) s = fmaf(v_s[u], tanh_<FAST>(__ldg(kp + u) + q_rm[r * A + u]), s);
                }
                s = warp_sum(s);
                if (lane == 0) stage[r * TJ + jj] = s + score_bias;
            }
        }
        __syncthreads();
        push_rm(cl, e_s, Tip, stage, rank, TJ, tid);
        cl.sync();
        // ===== P7: alignments (every CTA, warp r = row r), then context (own Ue units) =====
        {
            const int r = warp, n = grp * R + r;
            const int CH = (Tip + 31) / 32;
            const int j0 = lane * CH, j1 = min(j0 + CH, Ti);
            float* er = e_s + r * Tip; float* ar = a_s + r * Tip; float* pr = p_s + r * Tip; float* cr = cp_s + r * Tip;
            if (a.manual) {
                for (int j = lane; j < Ti; j += 32) ar[j] = (n < a.N) ? a.manual[((long long)n * Td + t) * Ti + j] : 0.f;
            } else if (a.att_type == TACO_ATT_BAH_MON) {
                // p = sigmoid(e); cp = exp(cumsum_excl(log(clip(1-p, tiny, 1)))); a = p*cp*cumsum(a_prev/clip(cp,1e-10,1))
                float ls = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float p = sigm_<FAST>(er[j]);
                    pr[j] = p;
                    ls += logf(fminf(fmaxf(1.f - p, FLT_MIN), 1.f));
                }
                float run = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
                run -= ls;   // exclusive prefix of this lane's chunk
                float ws = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float cp = expf(run);
                    cr[j] = cp;
                    run += logf(fminf(fmaxf(1.f - pr[j], FLT_MIN), 1.f));
                    ws += ar[j] / fminf(fmaxf(cp, 1e-10f), 1.f);
                }
                float run2 = ws;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
                run2 -= ws;
                for (int j = j0; j < j1; j++) {
                    run2 += ar[j] / fminf(fmaxf(cr[j], 1e-10f), 1.f);
                    ar[j] = pr[j] * cr[j] * run2;      // in place: a_prev[j] is not needed past this point
                }
            } else {   // softmax over the Ti memory positions (no masking: tacotron.py:133-134 passes no lengths)
                float mx = -INFINITY;
                for (int j = lane; j < Ti; j += 32) mx = fmaxf(mx, er[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float sm = 0.f;
                for (int j = lane; j < Ti; j += 32) { float ex = expf(er[j] - mx); pr[j] = ex; sm += ex; }
                sm = warp_sum(sm);
                for (int j = lane; j < Ti; j += 32) ar[j] = pr[j] / sm;
            }
            __syncwarp();
            if (rank == 0 && n < a.N) {
                for (int j = lane; j < Ti; j += 32) {
                    const float av = ar[j];
                    a.align[((long long)n * Ti + j) * Td + t] = av;
                    if (a.s_a) { a.s_a[((long long)n * Td + t) * Ti + j] = av; a.s_e[((long long)n * Td + t) * Ti + j] = er[j]; }
                }
            }
        }
        __syncthreads();
        {
            const int unit = rank * Ue + i32;
            float cx = 0.f;
            if (row_ok32) {
                const float* mp = a.memory + (long long)n32 * Ti * E + unit;
                const float* ar = a_s + r32 * Tip;
                int j = 0;
                for (; j + MV_LB <= Ti; j += MV_LB) {               // memory rows fetched MV_LB at a time (see mv_acc); same summation order
                    float mv[MV_LB];
#pragma unroll
                    for (int u = 0; u < MV_LB; u++) mv[u] = __ldg(mp + (long long)(j + u) * E);
#pragma unroll
                    for (int u = 0; u < MV_LB; u++) cx = fmaf(ar[j + u], mv[u], cx);
                }
                for (; j < Ti; j++) cx = fmaf(ar[j], __ldg(mp + (long long)j * E), cx);
                if (a.s_ctx) a.s_ctx[row32 * E + unit] = cx;
            }
            stage[i32 * R + r32] = cx;
        }
        __syncthreads();
        push_um(cl, ctx_s, stage, rank, Ue, tid);
        cl.sync();
        // ===== P8: y0 (own Uy columns) = yh + ctx.Wo_c (+ spk.Wo_s) + bo =====
        if (WFP) {
            for (int task = warp; task < 16; task += AT_NT / 32) {
                const int mt = task & 1, ks = task >> 1;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                wf_seg(acc, wf(WF_WO_C, 2, 16, mt), 16, 8, ks, ctx_s, lane);
                wf_store(red2, acc, 32, ks, mt, lane);
            }
        } else {
            const int ncols = Uy, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int kl = E / KS;
            mv_acc(acc, a.Wo + (long long)HA * Y + rank * Uy + col, Y, ctx_s, ks * kl, (ks + 1) * kl);
            if (SPK > 0) {
                kl = SPK / KS;
                if (kl > 0) mv_acc(acc, a.Wo + (long long)(HA + E) * Y + rank * Uy + col, Y, z_s + Z * R, ks * kl, (ks + 1) * kl);
                else if (ks == 0) mv_acc(acc, a.Wo + (long long)(HA + E) * Y + rank * Uy + col, Y, z_s + Z * R, 0, SPK);
            }
            red_store(red2, tid, acc);
        }
        __syncthreads();
        const float y0v = row_ok32 ? yh + red_sum(red2, r32, i32, Uy) + __ldg(a.bo + rank * Uy + i32) : 0.f;
        if (row_ok32 && a.y0) a.y0[row32 * Y + rank * Uy + i32] = y0v;
        // (teacher-forced mode) no barrier: the next step's P1 reads ctx_s (stable) and writes `red`, not `red2`
        if (a.free_run) {
            // ===== decoder RNN stack and mel projection inside the loop (tacotron.py:171-179; helpers.py:26-32) =====
            stage[i32 * R + r32] = y0v;
            __syncthreads();
            push_um(cl, y0_s, stage, rank, Uy, tid);
            cl.sync();
            float yin = y0v;
            for (int layer = 0; layer < 2; layer++) {
                const float* Wg_ = layer ? a.Wg2 : a.Wg1; const float* bg_ = layer ? a.bg2 : a.bg1;
                const float* Wc_ = layer ? a.Wc2 : a.Wc1; const float* bc_ = layer ? a.bc2 : a.bc1;
                float* xin_s = layer ? y1_s : y0_s; float* hst_s = layer ? h2_s : h1_s; float* yout_s = layer ? y2_s : y1_s;
                float& h_own = layer ? h2_own : h1_own;
                const int wf0 = layer ? WF_G2_GX : WF_G1_GX;
                if (WFP) {
                    for (int task = warp; task < 16; task += AT_NT / 32) {
                        {   // gates over [x ; h]: 4 m-tiles (r | u) x 4 k-slices
                            const int mt = task & 3, ks = task >> 2;
                            float acc[4] = {0.f, 0.f, 0.f, 0.f};
                            wf_seg(acc, wf(wf0, 4, 16, mt), 16, 4, ks, xin_s, lane);
                            wf_seg(acc, wf(wf0 + 1, 4, 16, mt), 16, 4, ks, hst_s, lane);
                            wf_store(red, acc, 64, ks, mt, lane);
                        }
                        {   // x part of the candidate: 2 m-tiles x 8 k-slices
                            const int mt = task & 1, ks = task >> 1;
                            float acc[4] = {0.f, 0.f, 0.f, 0.f};
                            wf_seg(acc, wf(wf0 + 2, 2, 16, mt), 16, 8, ks, xin_s, lane);
                            wf_store(red2, acc, 32, ks, mt, lane);
                        }
                    }
                } else {
                    {   // gates over [x ; h]
                        const int ncols = 2 * Uy, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
                        const int gcol = (col < Uy) ? rank * Uy + col : Y + rank * Uy + (col - Uy);
                        float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        const int kl = Y / KS;
                        mv_acc(acc, Wg_ + gcol, 2 * Y, xin_s, ks * kl, (ks + 1) * kl);
                        mv_acc(acc, Wg_ + (long long)Y * 2 * Y + gcol, 2 * Y, hst_s, ks * kl, (ks + 1) * kl);
                        red_store(red, tid, acc);
                    }
                    {   // x part of the candidate
                        const int ncols = Uy, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
                        float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        const int kl = Y / KS;
                        mv_acc(acc, Wc_ + rank * Uy + col, Y, xin_s, ks * kl, (ks + 1) * kl);
                        red_store(red2, tid, acc);
                    }
                }
                __syncthreads();
                const int unit = rank * Uy + i32;
                const float rgd = sigm_<FAST>(red_sum(red, r32, i32, 2 * Uy) + __ldg(bg_ + unit));
                const float ugd = sigm_<FAST>(red_sum(red, r32, Uy + i32, 2 * Uy) + __ldg(bg_ + Y + unit));
                const float cxd = red_sum(red2, r32, i32, Uy);
                stage[i32 * R + r32] = rgd * h_own;
                __syncthreads();
                push_um(cl, rhd_s, stage, rank, Uy, tid);
                cl.sync();
                if (WFP) {
                    for (int task = warp; task < 16; task += AT_NT / 32) {
                        const int mt = task & 1, ks = task >> 1;
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
                        wf_seg(acc, wf(wf0 + 3, 2, 16, mt), 16, 8, ks, rhd_s, lane);
                        wf_store(red, acc, 32, ks, mt, lane);
                    }
                } else {
                    const int ncols = Uy, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
                    float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    const int kl = Y / KS;
                    mv_acc(acc, Wc_ + (long long)Y * Y + rank * Uy + col, Y, rhd_s, ks * kl, (ks + 1) * kl);
                    red_store(red, tid, acc);
                }
                __syncthreads();
                const float cd = tanh_<FAST>(red_sum(red, r32, i32, Uy) + cxd + __ldg(bc_ + unit));
                const float hn = ugd * h_own + (1.f - ugd) * cd;
                h_own = row_ok32 ? hn : 0.f;
                yin = yin + h_own;                                  // ResidualWrapper: y = x + GRU(x, h)
                stage[i32 * R + r32] = h_own;
                stage[256 + i32 * R + r32] = yin;
                __syncthreads();
                push_um(cl, hst_s, stage, rank, Uy, tid);
                push_um(cl, yout_s, stage + 256, rank, Uy, tid);
                cl.sync();
            }
            {   // r-frame mel projection (own MR/C columns, padded to 64 for the thread mapping)
                const int MR = a.M * a.r, UO = MR / AT_C;
                if (WFP) {
                    for (int task = warp; task < 16; task += AT_NT / 32) {      // 4 m-tiles (UO columns padded to 64) x 4 k-slices
                        const int mt = task & 3, ks = task >> 2;
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
                        wf_seg(acc, wf(WF_MEL, 4, 16, mt), 16, 4, ks, y2_s, lane);
                        wf_store(red, acc, 64, ks, mt, lane);
                    }
                } else {
                    const int col = tid % 64, ks = tid / 64;
                    float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    const int kl = Y / 4;
                    if (col < UO) mv_acc(acc, a.Wmel + rank * UO + col, MR, y2_s, ks * kl, (ks + 1) * kl);
                    red_store(red, tid, acc);
                }
                __syncthreads();
                for (int idx = tid; idx < UO * R; idx += AT_NT) {
                    const int i = idx % UO, r = idx / UO, n = grp * R + r;
                    const int c = rank * UO + i;
                    float o = 0.f;
                    if (n < a.N) {
                        o = red_sum(red, r, i, 64) + __ldg(a.bmel + c);
                        a.mel_out[(long long)n * a.mel_bs + (long long)t * MR + c] = o;
                    }
                    if (c >= MR - a.M) {                            // last of the r frames feeds the next step
                        const int xi = c - (MR - a.M);
                        for (int peer = 0; peer < AT_C; peer++) cl.map_shared_rank(x_s, peer)[xi * R + r] = o;
                    }
                }
                cl.sync();
            }
        }
    }
    if (a.ha_final && row_ok32) a.ha_final[(long long)n32 * HA + rank * Uh + i32] = ha_own;
}

// ---------------------------------------------------------------------------------------------------------
// BPTT.  Transposed weight copies (W^T, packed once per step by the host side) let every matvec here use the same
// coalesced "lanes over output columns" form as the forward pass.
template <bool FAST>
__global__ void __launch_bounds__(AT_NT, 1) att_bwd_kernel(const AttArgs a) {
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int grp = blockIdx.x / AT_C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = AT_R;
    const int E = a.E, A = a.A, HA = a.HA, Z1 = a.Z1, Z = a.Z, SPK = a.SPK, Y = a.Y, Ti = a.Ti, Td = a.Td;
    const int ZS = Z + SPK, KIN = ZS + HA;     // attention-GRU input width
    const int TJ = (Ti + AT_C - 1) / AT_C, Tip = TJ * AT_C;
    const int Uz1 = Z1 / AT_C, Uz = Z / AT_C, Uh = HA / AT_C, Ua = A / AT_C, Ue = E / AT_C;

    extern __shared__ __align__(16) float smem[];
    float* dy_s = smem;                     // [Y][R]    dy0[t]
    float* gq_s = dy_s + Y * R;             // [A][R]
    float* dz1p_s = gq_s + A * R;           // [Z1][R]
    float* dzp_s = dz1p_s + Z1 * R;         // [Z][R]
    float* dcp_s = dzp_s + Z * R;           // [HA][R]
    float* dg_s = dcp_s + HA * R;           // [2HA][R]
    float* dctx_rm = dg_s + 2 * HA * R;     // [R][E]
    float* da_s = dctx_rm + R * E;          // [R][Tip]  grad wrt a_t through the context, gathered over the cluster
    float* dac_s = da_s + R * Tip;          // [R][Tip]  grad wrt a_t carried from step t+1's recurrence
    float* ge_s = dac_s + R * Tip;          // [R][Tip]  grad wrt the scores
    float* p_s = ge_s + R * Tip;            // [R][Tip]  scratch rows of the monotonic backward
    float* cp_s = p_s + R * Tip;
    float* s_s = cp_s + R * Tip;
    float* t1_s = s_s + R * Tip;
    float* t2_s = t1_s + R * Tip;
    float* red = t2_s + R * Tip;            // [R][NT]
    float* red2 = red + R * AT_NT;
    float* stage = red2 + R * AT_NT;        // [1024]
    float* v_s = stage + 1024;              // [A]

    if (a.att_type == TACO_ATT_BAH_NORM) {
        float ss = 0.f;
        for (int u = 0; u < A; u++) ss += a.v[u] * a.v[u];
        const float sc = a.att_g[0] * rsqrtf(ss);
        for (int u = tid; u < A; u += AT_NT) v_s[u] = a.v[u] * sc;
    } else {
        for (int u = tid; u < A; u += AT_NT) v_s[u] = a.v[u];
    }
    for (int idx = tid; idx < R * Tip; idx += AT_NT) { dac_s[idx] = 0.f; da_s[idx] = 0.f; ge_s[idx] = 0.f; }
    const int i32 = tid % 32, r32 = tid / 32;
    const int n32 = grp * R + r32;
    const bool row_ok32 = n32 < a.N;
    float dha_carry = 0.f;     // grad wrt ha_t arriving from step t+1 (own unit)
    float dctx_carry = 0.f;    // grad wrt ctx_t arriving from step t+1's prenet (own unit)
    float gbias_acc = 0.f;     // score-bias gradient (warp-row partial, rank 0 only)
    __syncthreads();
    cl.sync();

    for (int t = Td - 1; t >= 0; t--) {
        const long long row32 = (long long)n32 * Td + t;
        // load dy0[t] of the cluster's rows: dy_s[k][r]
        for (int idx = tid; idx < Y * R; idx += AT_NT) {
            const int k = idx % Y, r = idx / Y, n = grp * R + r;
            dy_s[k * R + r] = (n < a.N) ? __ldg(a.dy0 + ((long long)n * Td + t) * Y + k) : 0.f;
        }
        __syncthreads();
        // ===== Bp1: dha += dy0.Wo_h^T (own Uh), dctx = carry + dy0.Wo_c^T (own Ue) =====
        {
            const int ncols = Uh + Ue, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = Y / KS;
            const int gcol = (col < Uh) ? rank * Uh + col : HA + rank * Ue + (col - Uh);
            mv_acc(acc, a.WoT + gcol, HA + E + SPK, dy_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        float dha = dha_carry + red_sum(red, r32, i32, Uh + Ue);
        {
            const float dctx = dctx_carry + red_sum(red, r32, Uh + i32, Uh + Ue);
            if (row_ok32) a.d_ctx[row32 * E + rank * Ue + i32] = dctx;
            stage[r32 * Ue + i32] = row_ok32 ? dctx : 0.f;
        }
        __syncthreads();
        push_rm(cl, dctx_rm, E, stage, rank, Ue, tid);
        cl.sync();
        // ===== Bp2: da[r][own j] = sum_u dctx[r,u] memory[n,j,u] =====
        for (int p = warp; p < R * TJ; p += AT_NT / 32) {
            const int r = p % R, jj = p / R, j = rank * TJ + jj, n = grp * R + r;
            float s = 0.f;
            if (j < Ti && n < a.N) {
                const float* mp = a.memory + ((long long)n * Ti + j) * E;
                for (int u = lane; u < E; u += 32) s = fmaf(dctx_rm[r * E + u], __ldg(mp + u), s);
            }
            s = warp_sum(s);
            if (lane == 0) stage[r * TJ + jj] = s;
        }
        __syncthreads();
        push_rm(cl, da_s, Tip, stage, rank, TJ, tid);
        cl.sync();
        // ===== Bp3: attention-probability backward (every CTA; warp r = row r).  SURVEY.md Appendix C =====
        {
            const int r = warp, n = grp * R + r;
            const int CH = (Tip + 31) / 32;
            const int j0 = min(lane * CH, Ti), j1 = min(j0 + CH, Ti);
            float* ga = da_s + r * Tip; float* gc = dac_s + r * Tip; float* ge = ge_s + r * Tip;
            float* pr = p_s + r * Tip; float* cr = cp_s + r * Tip; float* sr = s_s + r * Tip;
            float* t1 = t1_s + r * Tip; float* t2 = t2_s + r * Tip;
            const bool ok = n < a.N;
            if (a.manual || !ok) {
                for (int j = lane; j < Tip; j += 32) { ge[j] = 0.f; gc[j] = 0.f; }
            } else if (a.att_type == TACO_ATT_BAH_MON) {
                const float* e_row = a.s_e + ((long long)n * Td + t) * Ti;
                const float* ap_row = (t > 0) ? a.s_a + ((long long)n * Td + (t - 1)) * Ti : nullptr;
                // forward recompute: p, l=log(clip(1-p)), L=cumsum_excl(l), cp=exp(L), w=ap/clip(cp), s=cumsum(w)
                float ls = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float p = sigm_<FAST>(e_row[j]);
                    pr[j] = p;
                    ls += logf(fminf(fmaxf(1.f - p, FLT_MIN), 1.f));
                }
                float run = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
                run -= ls;
                float ws = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float cp = expf(run);
                    cr[j] = cp;
                    run += logf(fminf(fmaxf(1.f - pr[j], FLT_MIN), 1.f));
                    const float ap = ap_row ? ap_row[j] : (j == 0 ? 1.f : 0.f);
                    ws += ap / fminf(fmaxf(cp, 1e-10f), 1.f);
                }
                float run2 = ws;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
                run2 -= ws;
                float gs_loc = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float ap = ap_row ? ap_row[j] : (j == 0 ? 1.f : 0.f);
                    run2 += ap / fminf(fmaxf(cr[j], 1e-10f), 1.f);
                    sr[j] = run2;
                    const float gs = (ga[j] + gc[j]) * pr[j] * cr[j];     // grad wrt s_j
                    t1[j] = gs;
                    gs_loc += gs;
                }
                float suf = gs_loc;    // -> sum over lanes strictly above
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
                suf -= gs_loc;
                float gL_loc = 0.f;
                {
                    float acc = suf;
                    for (int j = j1 - 1; j >= j0; j--) {
                        acc += t1[j];                                   // gw_j = sum_{i>=j} gs_i
                        const float g = ga[j] + gc[j];
                        const float p = pr[j], cp = cr[j], sj = sr[j];
                        const float ap = ap_row ? ap_row[j] : (j == 0 ? 1.f : 0.f);
                        const float d = fminf(fmaxf(cp, 1e-10f), 1.f);
                        float gcp = g * p * sj;
                        if (cp >= 1e-10f && cp <= 1.f) gcp -= acc * ap / (d * d);
                        const float gL = gcp * cp;
                        t2[j] = gL;
                        gL_loc += gL;
                        ge[j] = g * cp * sj;                            // direct part of grad wrt p_j
                        gc[j] = acc / d;                                // grad wrt a_{t-1,j}, carried to the next iteration
                    }
                }
                float sufL = gL_loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_down_sync(0xffffffffu, sufL, o); if (lane + o < 32) sufL += v; }
                sufL -= gL_loc;
                float gb = 0.f;
                {
                    float acc = sufL;                                   // gl_j = sum_{i>j} gL_i
                    for (int j = j1 - 1; j >= j0; j--) {
                        const float p = pr[j], omp = 1.f - p;
                        float gp = ge[j];
                        if (omp >= FLT_MIN && omp <= 1.f) gp -= acc / fminf(fmaxf(omp, FLT_MIN), 1.f);
                        acc += t2[j];
                        const float gev = gp * p * (1.f - p);
                        ge[j] = gev;
                        gb += gev;
                    }
                }
                for (int j = Ti + lane; j < Tip; j += 32) { ge[j] = 0.f; gc[j] = 0.f; }
                gb = warp_sum(gb);
                if (rank == 0 && lane == 0) gbias_acc += gb;
            } else {
                // softmax: ge = a * (g - sum a*g); nothing is carried (a_prev does not enter)
                const float* a_row = a.s_a + ((long long)n * Td + t) * Ti;
                float dot = 0.f;
                for (int j = lane; j < Ti; j += 32) dot += a_row[j] * ga[j];
                dot = warp_sum(dot);
                for (int j = lane; j < Tip; j += 32) { ge[j] = (j < Ti) ? a_row[j] * (ga[j] - dot) : 0.f; gc[j] = 0.f; }
            }
            __syncwarp();
            if (rank == 0 && ok && a.d_ge)
                for (int j = lane; j < Ti; j += 32) a.d_ge[((long long)n * Td + t) * Ti + j] = ge[j];
        }
        __syncthreads();
        // ===== Bp4: gq (own Ua units) = v_u * sum_j ge[r][j] * (1 - tanh^2(keys[n,j,u] + q[r,u])) =====
        {
            const int unit = rank * Ua + i32;
            float gq = 0.f;
            if (row_ok32 && !a.manual) {
                const float q = a.s_q[row32 * A + unit];
                const float* kp = a.keys + (long long)n32 * Ti * A + unit;
                const float* ger = ge_s + r32 * Tip;
                float acc = 0.f;
#pragma unroll 2
                for (int j = 0; j < Ti; j++) {
                    const float th = tanh_<FAST>(__ldg(kp + (long long)j * A) + q);
                    acc = fmaf(ger[j], 1.f - th * th, acc);
                }
                gq = acc * v_s[unit];
            }
            if (row_ok32 && a.d_gq) a.d_gq[row32 * A + unit] = gq;
            stage[i32 * R + r32] = gq;
        }
        __syncthreads();
        push_um(cl, gq_s, stage, rank, Ua, tid);
        cl.sync();
        // ===== Bp5: dha += gq.Wq^T; GRU cell backward (elementwise part) =====
        {
            const int ncols = Uh, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = A / KS;
            mv_acc(acc, a.WqT + rank * Uh + col, HA, gq_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        float rg = 0.f, ug = 0.f, cc = 0.f, hp = 0.f, du_pre = 0.f, dc_pre = 0.f;
        {
            dha += red_sum(red, r32, i32, Uh);
            if (row_ok32) {
                const long long o = row32 * HA + rank * Uh + i32;
                rg = a.s_r[o]; ug = a.s_u[o]; cc = a.s_c[o]; hp = a.s_haprev[o];
                du_pre = dha * (hp - cc) * ug * (1.f - ug);
                dc_pre = dha * (1.f - ug) * (1.f - cc * cc);
            }
            stage[i32 * R + r32] = dc_pre;
        }
        __syncthreads();
        push_um(cl, dcp_s, stage, rank, Uh, tid);
        cl.sync();
        // ===== Bp6: d(r*h) (own Uh) = dc_pre . Wc_h^T =====
        {
            const int ncols = Uh, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = HA / KS;
            mv_acc(acc, a.WcT + ZS + rank * Uh + col, KIN, dcp_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        float d_rh = 0.f, dr_pre = 0.f;
        {
            d_rh = red_sum(red, r32, i32, Uh);
            if (row_ok32) dr_pre = d_rh * hp * rg * (1.f - rg);
            stage[i32 * R + r32] = dr_pre;
            stage[256 + i32 * R + r32] = du_pre;
        }
        __syncthreads();
        push_um(cl, dg_s, stage, rank, Uh, tid);
        push_um(cl, dg_s + HA * R, stage + 256, rank, Uh, tid);
        cl.sync();
        // ===== Bp7: dha_prev (own Uh) and dz (own Uz) =====
        {
            const int ncols = Uh, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = 2 * HA / KS;
            mv_acc(acc, a.WgT + ZS + rank * Uh + col, KIN, dg_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        {
            const int ncols = Uz, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int kl = 2 * HA / KS;
            mv_acc(acc, a.WgT + rank * Uz + col, KIN, dg_s, ks * kl, (ks + 1) * kl);
            kl = HA / KS;
            mv_acc(acc, a.WcT + rank * Uz + col, KIN, dcp_s, ks * kl, (ks + 1) * kl);
            red_store(red2, tid, acc);
        }
        __syncthreads();
        if (row_ok32) {
            dha_carry = dha * ug + d_rh * rg + red_sum(red, r32, i32, Uh);
            float* g = a.d_G + row32 * 3 * HA + rank * Uh + i32;
            g[0] = dr_pre; g[HA] = du_pre; g[2 * HA] = dc_pre;
            a.s_r[row32 * HA + rank * Uh + i32] = rg * hp;      // operand of the candidate-weight gradient GEMM
        }
        if (tid < Uz * R) {
            const int i = tid % Uz, r = tid / Uz, n = grp * R + r;
            float v = 0.f;
            if (n < a.N) {
                const long long row = (long long)n * Td + t;
                const float z = a.s_z[row * Z + rank * Uz + i];
                v = (z > 0.f) ? red_sum(red2, r, i, Uz) : 0.f;
                a.d_zp[row * Z + rank * Uz + i] = v;
            }
            stage[i * R + r] = v;
        }
        __syncthreads();
        push_um(cl, dzp_s, stage, rank, Uz, tid);
        cl.sync();
        // ===== Bp8: dz1 (own Uz1) = dz_pre . W2^T, relu' =====
        {
            const int ncols = Uz1, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = Z / KS;
            mv_acc(acc, a.W2T + rank * Uz1 + col, Z1, dzp_s, ks * kl, (ks + 1) * kl);
            red_store(red, tid, acc);
        }
        __syncthreads();
        if (tid < Uz1 * R) {
            const int i = tid % Uz1, r = tid / Uz1, n = grp * R + r;
            float v = 0.f;
            if (n < a.N) {
                const long long row = (long long)n * Td + t;
                const float z1 = a.s_z1[row * Z1 + rank * Uz1 + i];
                v = (z1 > 0.f) ? red_sum(red, r, i, Uz1) : 0.f;
                a.d_z1p[row * Z1 + rank * Uz1 + i] = v;
            }
            stage[i * R + r] = v;
        }
        __syncthreads();
        push_um(cl, dz1p_s, stage, rank, Uz1, tid);
        cl.sync();
        // ===== Bp9: grad wrt ctx_{t-1} (own Ue) = dz1_pre . W1c^T =====
        {
            const int ncols = Ue, col = tid % ncols, ks = tid / ncols, KS = AT_NT / ncols;
            float acc[AT_R] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int kl = Z1 / KS;
            mv_acc(acc, a.W1cT + rank * Ue + col, E, dz1p_s, ks * kl, (ks + 1) * kl);
            red_store(red2, tid, acc);
        }
        __syncthreads();
        dctx_carry = red_sum(red2, r32, i32, Ue);
        __syncthreads();   // red2/stage/dy_s are rewritten at the top of the next iteration
    }
    if (a.d_ha0 && row_ok32) a.d_ha0[(long long)n32 * HA + rank * Uh + i32] = dha_carry;
    if (rank == 0 && lane == 0 && a.d_score_bias && a.att_type == TACO_ATT_BAH_MON) atomicAdd(a.d_score_bias, gbias_acc);
}

// ---------------------------------------------------------------------------------------------------------
// Key / v gradients, fully parallel over (n, j) (not part of the serial chain):
//   dkeys[n,j,u] = v_u * sum_t ge[n,t,j] * (1 - th^2),  gv[u] += sum_{n,t,j} ge * th,  th = tanh(keys[n,j,u] + q[n,t,u])
constexpr int KB_J = 8;   // memory positions per block
template <bool FAST>
__global__ void att_keys_bwd_kernel(const float* __restrict__ keys, const float* __restrict__ q, const float* __restrict__ ge,
                                    const float* __restrict__ v_eff, float* __restrict__ dkeys, float* __restrict__ gv,
                                    int N, int Ti, int Td, int A) {
    const int chunks = (Ti + KB_J - 1) / KB_J;
    const int n = blockIdx.x / chunks, jc = blockIdx.x % chunks;
    const int j0 = jc * KB_J, nj = min(KB_J, Ti - j0);
    // t outer, the block's KB_J positions inner and in registers: q is read once per (t, u) and every load feeds KB_J
    // independent tanh chains
    for (int u = threadIdx.x; u < A; u += blockDim.x) {
        float k[KB_J], acc[KB_J];
#pragma unroll
        for (int jj = 0; jj < KB_J; jj++) { k[jj] = (jj < nj) ? keys[((long long)n * Ti + j0 + jj) * A + u] : 0.f; acc[jj] = 0.f; }
        float accv = 0.f;
        const float vu = v_eff[u];
        const float* gep = ge + (long long)n * Td * Ti + j0;
        const float* qp = q + (long long)n * Td * A + u;
        for (int t = 0; t < Td; t++) {
            const float qv = __ldg(qp + (long long)t * A);
            float g[KB_J];
#pragma unroll
            for (int jj = 0; jj < KB_J; jj++) g[jj] = (jj < nj) ? __ldg(gep + (long long)t * Ti + jj) : 0.f;
#pragma unroll
            for (int jj = 0; jj < KB_J; jj++) {
                const float th = tanh_<FAST>(k[jj] + qv);
                acc[jj] = fmaf(g[jj], 1.f - th * th, acc[jj]);
                accv = fmaf(g[jj], th, accv);
            }
        }
#pragma unroll
        for (int jj = 0; jj < KB_J; jj++) if (jj < nj) dkeys[((long long)n * Ti + j0 + jj) * A + u] = acc[jj] * vu;
        atomicAdd(gv + u, accv);
    }
}

// bah_norm (TF BahdanauAttention normalize=True): v_eff = g * v / ||v||, and the chain back to v and g.
__global__ void att_vnorm_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ v_eff,
                                 const float* __restrict__ gveff, float* __restrict__ gv, float* __restrict__ gg, int A) {
    __shared__ float red[2][32];
    float ss = 0.f, sd = 0.f;
    for (int u = threadIdx.x; u < A; u += blockDim.x) { const float x = v[u]; ss = fmaf(x, x, ss); if (gveff) sd = fmaf(x, gveff[u], sd); }
    for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); sd += __shfl_xor_sync(0xffffffffu, sd, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = sd; }
    __syncthreads();
    ss = 0.f; sd = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { ss += red[0][w]; sd += red[1][w]; }
    const float rn = rsqrtf(ss), gs = g[0];
    if (!gveff) {
        for (int u = threadIdx.x; u < A; u += blockDim.x) v_eff[u] = gs * v[u] * rn;
    } else {
        if (threadIdx.x == 0) gg[0] += sd * rn;
        for (int u = threadIdx.x; u < A; u += blockDim.x) gv[u] += gs * rn * (gveff[u] - v[u] * sd * rn * rn);
    }
}
int launch_att_vnorm(const float* v, const float* g, float* v_eff, const float* gveff, float* gv, float* gg, int A, cudaStream_t s) {
    att_vnorm_kernel<<<1, 256, 0, s>>>(v, g, v_eff, gveff, gv, gg, A);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

static int att_check(const AttArgs& a) {
    TACO_REQUIRE(a.N > 0 && a.Ti > 0 && a.Td > 0, TACO_ESHAPE, "attention: empty shape");
    TACO_REQUIRE(a.E % 256 == 0 || a.E == 256, TACO_ESHAPE, "attention: memory width %d unsupported", a.E);
    TACO_REQUIRE(a.E == 256 && a.A == 256 && a.HA == 256 && a.Z1 == 256 && a.Z == 128 && a.Y == 256, TACO_ESHAPE,
                 "attention kernel instantiated for E=A=HA=Z1=Y=256, Z=128 (got %d %d %d %d %d %d)", a.E, a.A, a.HA, a.Z1, a.Y, a.Z);
    TACO_REQUIRE(a.SPK == 0 || a.SPK == 16, TACO_ESHAPE, "attention: speaker width %d unsupported", a.SPK);
    TACO_REQUIRE(a.Ti <= 1024, TACO_ESHAPE, "attention: T_in %d too long", a.Ti);
    if (a.free_run) TACO_REQUIRE(a.M % 8 == 0 && (a.M * a.r) % AT_C == 0 && (a.M * a.r) / AT_C <= 64 && a.W1x && a.Wg1 && a.Wg2 && a.Wmel && a.mel_out,
                                 TACO_ESHAPE, "free-running decoder: unsupported mel/r sizes or missing weights");
    return TACO_OK;
}

template <typename K>
static int att_launch(K kern, const AttArgs& a, bool bwd, cudaStream_t s) {
    const int TJ = (a.Ti + AT_C - 1) / AT_C, Tip = TJ * AT_C;
    const size_t smem = (bwd ? att_bwd_smem_floats(a, Tip) : att_fwd_smem_floats(a, Tip)) * sizeof(float);
    TACO_REQUIRE(smem <= 227 * 1024, TACO_ESHAPE, "attention: shared memory %zu exceeds 227 KB", smem);
    TACO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(AT_C * cdiv(a.N, AT_R));
    cfg.blockDim = dim3(AT_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = AT_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TACO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    g_launch_count++;
    return TACO_OK;
}


// ---- fragment-packed weights: host side ----------------------------------------------------------------------------------
bool att_wfrag_supported(const AttArgs& a) {
    return a.free_run && a.fast && a.SPK == 0 && a.E == 256 && a.A == 256 && a.HA == 256 && a.Z1 == 256 && a.Z == 128 && a.Y == 256 &&
           a.M == 80 && (a.M * a.r) % AT_C == 0 && (a.M * a.r) / AT_C <= 64;
}
static void wf_table(const AttArgs& a, WfTable& tb) {
    const int U32 = 32, MR = a.M * a.r, UO = MR / AT_C, Y = a.Y, HA = a.HA, E = a.E, Z = a.Z, Z1 = a.Z1, A = a.A, M = a.M;
    auto one = [](const float* W, int ld, int U, int off, int row0, int K, int ncols, int valid) {
        WfSpec s{}; s.WA = W; s.WB = W; s.ldA = ld; s.ldB = ld; s.UA = U; s.UB = U; s.offA = off; s.offB = off; s.split = ncols; s.row0 = row0; s.K = K; s.ncols = ncols; s.valid = valid;
        return s;
    };
    // gate matrices: own r columns [rank*U, +U) then own u columns [H + rank*U, +U)
    auto gates = [](const float* W, int ld, int U, int H, int row0, int K) {
        WfSpec s{}; s.WA = W; s.WB = W; s.ldA = ld; s.ldB = ld; s.UA = U; s.UB = U; s.offA = 0; s.offB = H; s.split = U; s.row0 = row0; s.K = K; s.ncols = 2 * U; s.valid = 2 * U;
        return s;
    };
    const float* W1 = a.W1x;                       // dense_1 kernel [M + E, Z1]: rows [0, M) take the frame, rows [M, M + E) the context
    tb.s[WF_W1C] = one(W1, Z1, U32, 0, M, E, 32, 32);
    tb.s[WF_W1X] = one(W1, Z1, U32, 0, 0, M, 32, 32);
    tb.s[WF_W2] = one(a.W2, Z, 16, 0, 0, Z1, 16, 16);
    tb.s[WF_WG_Z] = gates(a.Wg, 2 * HA, U32, HA, 0, Z);
    tb.s[WF_WG_H] = gates(a.Wg, 2 * HA, U32, HA, Z, HA);
    tb.s[WF_WC_Z] = one(a.Wc, HA, U32, 0, 0, Z, 32, 32);
    tb.s[WF_WC_H] = one(a.Wc, HA, U32, 0, Z, HA, 32, 32);
    {   // query columns, then the ha rows of the concat projection
        WfSpec s{}; s.WA = a.Wq; s.ldA = A; s.UA = U32; s.offA = 0; s.WB = a.Wo; s.ldB = Y; s.UB = U32; s.offB = 0; s.split = 32; s.row0 = 0; s.K = HA; s.ncols = 64; s.valid = 64;
        tb.s[WF_WQWO] = s;
    }
    tb.s[WF_WO_C] = one(a.Wo, Y, U32, 0, HA, E, 32, 32);
    const float* Wg_[2] = {a.Wg1, a.Wg2}; const float* Wc_[2] = {a.Wc1, a.Wc2};
    for (int l = 0; l < 2; l++) {
        const int b = l ? WF_G2_GX : WF_G1_GX;
        tb.s[b] = gates(Wg_[l], 2 * Y, U32, Y, 0, Y);
        tb.s[b + 1] = gates(Wg_[l], 2 * Y, U32, Y, Y, Y);
        tb.s[b + 2] = one(Wc_[l], Y, U32, 0, 0, Y, 32, 32);
        tb.s[b + 3] = one(Wc_[l], Y, U32, 0, Y, Y, 32, 32);
    }
    tb.s[WF_MEL] = one(a.Wmel, MR, UO, 0, 0, Y, 64, UO);
    long long off = 0;
    for (int i = 0; i < WF_N; i++) {
        tb.off[i] = off;
        off += (long long)AT_C * (tb.s[i].ncols / 16) * ((tb.s[i].K + 15) / 16) * 32;
    }
    tb.off[WF_N] = off;
}
size_t att_wfrag_bytes(const AttArgs& a) {
    if (!att_wfrag_supported(a)) return 0;
    AttArgs b = a;
    WfTable tb; wf_table(b, tb);
    return (size_t)tb.off[WF_N] * sizeof(uint4);
}
// packs the step's weights into `buf` (att_wfrag_bytes) and points the arguments at the image
int launch_att_wfrag_pack(AttArgs& a, void* buf, cudaStream_t s) {
    TACO_REQUIRE(att_wfrag_supported(a) && buf, TACO_EINVAL, "attention: fragment-packed weights do not apply to this configuration");
    WfTable tb; wf_table(a, tb);
    wf_pack_kernel<<<dim3(64, WF_N), 256, 0, s>>>(tb, static_cast<uint4*>(buf));
    TACO_CHECK_LAUNCH();
    a.wfrag = buf;
    for (int i = 0; i < WF_N; i++) a.wf_off[i] = tb.off[i];
    return TACO_OK;
}

int launch_att_fwd(const AttArgs& a, cudaStream_t s) {
    TACO_TRY(att_check(a));
    if (a.fast && !a.free_run && att_fast_supported(a)) { int rc = launch_att_fast_fwd(a, s); if (rc != TACO_ENOTSUP) return rc; }
    TACO_REQUIRE(a.t_begin == 0 && (a.t_end == 0 || a.t_end == a.Td), TACO_EINVAL, "attention: the exact kernels do not run time chunks");
    return a.fast ? att_launch(att_fwd_kernel<true>, a, false, s) : att_launch(att_fwd_kernel<false>, a, false, s);
}
int launch_att_bwd(const AttArgs& a, cudaStream_t s) {
    TACO_TRY(att_check(a));
    TACO_REQUIRE(a.dy0 && a.d_G && a.d_zp && a.d_z1p && a.d_ctx && a.s_e && a.s_a && a.s_q, TACO_EINVAL, "attention bwd: missing buffers");
    if (a.fast && att_fast_supported(a)) { int rc = launch_att_fast_bwd(a, s); if (rc != TACO_ENOTSUP) return rc; }
    TACO_REQUIRE(a.t_begin == 0 && (a.t_end == 0 || a.t_end == a.Td) && !a.c_in, TACO_EINVAL, "attention: the exact kernels do not run time chunks");
    return a.fast ? att_launch(att_bwd_kernel<true>, a, true, s) : att_launch(att_bwd_kernel<false>, a, true, s);
}
int launch_att_keys_bwd(const float* keys, const float* q, const float* ge, const float* v_eff, float* dkeys, float* gv,
                        int N, int Ti, int Td, int A, int fast, cudaStream_t s) {
    const int blocks = N * ((Ti + KB_J - 1) / KB_J);
    if (fast) att_keys_bwd_kernel<true><<<blocks, 256, 0, s>>>(keys, q, ge, v_eff, dkeys, gv, N, Ti, Td, A);
    else att_keys_bwd_kernel<false><<<blocks, 256, 0, s>>>(keys, q, ge, v_eff, dkeys, gv, N, Ti, Td, A);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

}  // namespace taco
