// fp32 SIMT GEMM — the exact-precision contraction path (TACO_PREC_FP32) and the fallback for
// shapes the tensor-core path does not take.  One kernel covers every contraction of the hot
// path through its addressing modes:
//   * plain dense layers            (reference: tf.layers.dense call sites, modules.py:23,73,109-120; tacotron.py:235)
//   * conv1d as implicit GEMM        (modules.py:125-129) — "tap" addressing over a zero-padded
//                                    [N*Tp, C] activation buffer: A(m,k) = X[(m + k/ctap)*lda + k%ctap]
//   * weight gradients (A^T B)       (TF autodiff of the above, tacotron.py:328), split-K with fp32 atomics
//   * fused epilogues                bias, activation (activation-before-BN order of modules.py:125-131),
//                                    pad-row masking, output row remap, per-column sum / sum-of-squares for
//                                    training-mode batch-norm statistics (modules.py:131).
#include "common.cuh"

namespace taco {

struct GemmP {
    const float* A; const float* B; float* C;
    int M, N, K, lda, ldb, ldc;
    int transA, ctap, transB;
    float alpha; int accumulate;
    const float* bias; int act;
    int mask_period, mask_lo, mask_hi;
    int remap_period; long long remap_outer, remap_inner;
    double* colsum; double* colsumsq;
    int split_k;
    int vecA, vecB, vecC;
};

constexpr int GEMM_MAX_BATCH = 12;
struct GemmBatch { GemmP p[GEMM_MAX_BATCH]; };

constexpr int BK = 16;
constexpr int NT = 256;

template <int BM, int BN>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const __grid_constant__ GemmBatch batch) {
    constexpr int RM = BM / 64;          // row groups of 64
    constexpr int RN = BN / 64;
    constexpr int TM = 4 * RM, TN = 4 * RN;
    constexpr int LDS_A = BM + 4, LDS_B = BN + 4;
    constexpr int A_VECS = BM * BK / 4 / NT;   // float4 per thread per tile
    constexpr int B_VECS = BN * BK / 4 / NT;
    static_assert(A_VECS >= 1 && B_VECS >= 1, "tile too small");

    const GemmP& p = batch.p[blockIdx.z];
    const int tilesN = (p.N + BN - 1) / BN;
    const int tilesM = (p.M + BM - 1) / BM;
    if ((int)blockIdx.x >= tilesN * tilesM || (int)blockIdx.y >= p.split_k) return;
    const int tn = blockIdx.x % tilesN, tm = blockIdx.x / tilesN;
    const int m0 = tm * BM, n0 = tn * BN;

    int kchunk = ((p.K + p.split_k - 1) / p.split_k + BK - 1) / BK * BK;
    const int k_begin = blockIdx.y * kchunk;
    const int k_end = min(p.K, k_begin + kchunk);
    if (k_begin >= k_end && p.split_k > 1) return;

    __shared__ __align__(16) float As[2][BK][LDS_A];
    __shared__ __align__(16) float Bs[2][BK][LDS_B];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const float* __restrict__ A = p.A;
    const float* __restrict__ B = p.B;
    const int ctapA = p.ctap > 0 ? p.ctap : 0x7fffffff;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    float4 ra[A_VECS], rb[B_VECS];

    auto load_a = [&](int kt) {
#pragma unroll
        for (int i = 0; i < A_VECS; i++) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!p.transA) {
                int m = idx / (BK / 4), k4 = idx % (BK / 4);
                int gm = m0 + m, gk = kt + k4 * 4;
                if (gm < p.M) {
                    if (p.vecA && gk + 3 < k_end) {
                        long long off = (long long)(gm + gk / ctapA) * p.lda + gk % ctapA;
                        v = __ldg(reinterpret_cast<const float4*>(A + off));
                    } else {
                        float t[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            int k = gk + e;
                            t[e] = (k < k_end) ? __ldg(A + (long long)(gm + k / ctapA) * p.lda + k % ctapA) : 0.f;
                        }
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            } else {
                int k = idx / (BM / 4), m4 = idx % (BM / 4);
                int gk = kt + k, gm = m0 + m4 * 4;
                if (gk < k_end) {
                    if (p.vecA && gm + 3 < p.M) {
                        long long off = (long long)(gk + gm / ctapA) * p.lda + gm % ctapA;
                        v = __ldg(reinterpret_cast<const float4*>(A + off));
                    } else {
                        float t[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            int m = gm + e;
                            t[e] = (m < p.M) ? __ldg(A + (long long)(gk + m / ctapA) * p.lda + m % ctapA) : 0.f;
                        }
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
            ra[i] = v;
        }
    };
    auto load_b = [&](int kt) {
#pragma unroll
        for (int i = 0; i < B_VECS; i++) {
            int idx = tid + i * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!p.transB) {
                int k = idx / (BN / 4), n4 = idx % (BN / 4);
                int gk = kt + k, gn = n0 + n4 * 4;
                if (gk < k_end) {
                    if (p.vecB && gn + 3 < p.N) {
                        v = __ldg(reinterpret_cast<const float4*>(B + (long long)gk * p.ldb + gn));
                    } else {
                        float t[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) t[e] = (gn + e < p.N) ? __ldg(B + (long long)gk * p.ldb + gn + e) : 0.f;
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            } else {
                int n = idx / (BK / 4), k4 = idx % (BK / 4);
                int gn = n0 + n, gk = kt + k4 * 4;
                if (gn < p.N) {
                    if (p.vecB && gk + 3 < k_end) {
                        v = __ldg(reinterpret_cast<const float4*>(B + (long long)gn * p.ldb + gk));
                    } else {
                        float t[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) t[e] = (gk + e < k_end) ? __ldg(B + (long long)gn * p.ldb + gk + e) : 0.f;
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
            rb[i] = v;
        }
    };
    auto store_ab = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_VECS; i++) {
            int idx = tid + i * NT;
            if (!p.transA) {
                int m = idx / (BK / 4), k4 = idx % (BK / 4);
                As[buf][k4 * 4 + 0][m] = ra[i].x; As[buf][k4 * 4 + 1][m] = ra[i].y;
                As[buf][k4 * 4 + 2][m] = ra[i].z; As[buf][k4 * 4 + 3][m] = ra[i].w;
            } else {
                int k = idx / (BM / 4), m4 = idx % (BM / 4);
                *reinterpret_cast<float4*>(&As[buf][k][m4 * 4]) = ra[i];
            }
        }
#pragma unroll
        for (int i = 0; i < B_VECS; i++) {
            int idx = tid + i * NT;
            if (!p.transB) {
                int k = idx / (BN / 4), n4 = idx % (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][k][n4 * 4]) = rb[i];
            } else {
                int n = idx / (BK / 4), k4 = idx % (BK / 4);
                Bs[buf][k4 * 4 + 0][n] = rb[i].x; Bs[buf][k4 * 4 + 1][n] = rb[i].y;
                Bs[buf][k4 * 4 + 2][n] = rb[i].z; Bs[buf][k4 * 4 + 3][n] = rb[i].w;
            }
        }
    };

    if (k_begin < k_end) {
        load_a(k_begin); load_b(k_begin);
        store_ab(0);
    }
    __syncthreads();
    int buf = 0;
    for (int kt = k_begin; kt < k_end; kt += BK) {
        const bool has_next = kt + BK < k_end;
        if (has_next) { load_a(kt + BK); load_b(kt + BK); }
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < RM; g++) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * 64 + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < RN; g++) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store_ab(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue --------------------------------------------------------------------
    const bool want_stats = (p.colsum != nullptr);
    float cs[TN], css[TN];
#pragma unroll
    for (int j = 0; j < TN; j++) { cs[j] = 0.f; css[j] = 0.f; }

#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (gm >= p.M) continue;
        bool masked = false;
        if (p.mask_period > 0) {
            int t = gm % p.mask_period;
            masked = (t < p.mask_lo) || (t >= p.mask_hi);
        }
        float* crow = p.remap_period > 0
            ? p.C + (long long)(gm / p.remap_period) * p.remap_outer + (long long)(gm % p.remap_period) * p.remap_inner
            : p.C + (long long)gm * p.ldc;
#pragma unroll
        for (int g = 0; g < RN; g++) {
            const int gn = n0 + g * 64 + tx * 4;
            if (gn >= p.N) continue;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float x = p.alpha * acc[i][g * 4 + e];
                if (p.split_k == 1 && p.accumulate != 2) {
                    if (p.bias && gn + e < p.N) x += __ldg(p.bias + gn + e);
                    x = apply_act(x, p.act);
                } else if (p.bias && blockIdx.y == 0 && gn + e < p.N) {
                    x += __ldg(p.bias + gn + e);
                }
                v[e] = masked ? 0.f : x;
            }
            if (p.split_k > 1 || p.accumulate == 2) {
                if (!masked) {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (gn + e < p.N) atomicAdd(crow + gn + e, v[e]);
                }
            } else {
                if (p.accumulate) {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (gn + e < p.N) v[e] += crow[gn + e];
                }
                if (p.vecC && gn + 3 < p.N) {
                    *reinterpret_cast<float4*>(crow + gn) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (gn + e < p.N) crow[gn + e] = v[e];
                }
                if (want_stats && !masked) {
#pragma unroll
                    for (int e = 0; e < 4; e++) { cs[g * 4 + e] += v[e]; css[g * 4 + e] += v[e] * v[e]; }
                }
            }
        }
    }

    if (want_stats) {   // uniform across the block (kernel-parameter condition)
        float* red_s = &As[0][0][0];    // [16][BN]
        float* red_q = &Bs[0][0][0];
        static_assert(2 * BK * LDS_A >= 16 * BN && 2 * BK * LDS_B >= 16 * BN, "reduction scratch too small");
        __syncthreads();
#pragma unroll
        for (int g = 0; g < RN; g++)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                red_s[ty * BN + g * 64 + tx * 4 + e] = cs[g * 4 + e];
                red_q[ty * BN + g * 64 + tx * 4 + e] = css[g * 4 + e];
            }
        __syncthreads();
        if (tid < BN && n0 + tid < p.N) {
            double s = 0.0, q = 0.0;
#pragma unroll
            for (int r = 0; r < 16; r++) { s += red_s[r * BN + tid]; q += red_q[r * BN + tid]; }
            atomicAdd(p.colsum + n0 + tid, s);
            atomicAdd(p.colsumsq + n0 + tid, q);
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_gemm_simt(const taco_gemm_desc* d, int n, cudaStream_t s) {
    int done = 0;
    while (done < n) {
        GemmBatch b;
        int cnt = 0;
        int max_tiles128 = 0, max_tiles64 = 0, max_split = 1;
        bool small_n = true;
        for (; cnt < GEMM_MAX_BATCH && done + cnt < n; cnt++) {
            const taco_gemm_desc& g = d[done + cnt];
            TACO_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, TACO_ESHAPE, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
            TACO_REQUIRE(g.split_k >= 1, TACO_EINVAL, "gemm: split_k must be >= 1");
            TACO_REQUIRE(g.tap_table == nullptr, TACO_EINVAL, "gemm: tap tables are only served by the tensor-core kernel (TF32 mode, aligned operands)");
            TACO_REQUIRE(!((g.split_k > 1 || g.accumulate == 2) && (g.act != 0 || g.colsum != nullptr)),
                         TACO_EINVAL, "gemm: atomic accumulation cannot carry an activation or column statistics");
            GemmP& p = b.p[cnt];
            p.A = static_cast<const float*>(g.A); p.B = static_cast<const float*>(g.B); p.C = g.C;
            p.M = g.M; p.N = g.N; p.K = g.K; p.lda = g.lda; p.ldb = g.ldb; p.ldc = g.ldc;
            p.transA = g.transA; p.ctap = g.ctap; p.transB = g.transB;
            p.alpha = g.alpha; p.accumulate = g.accumulate; p.bias = g.bias; p.act = g.act;
            p.mask_period = g.mask_period; p.mask_lo = g.mask_lo; p.mask_hi = g.mask_hi;
            p.remap_period = g.remap_period; p.remap_outer = g.remap_outer; p.remap_inner = g.remap_inner;
            p.colsum = g.colsum; p.colsumsq = g.colsumsq; p.split_k = g.split_k;
            const int ct = g.ctap > 0 ? g.ctap : 4;
            if (!g.transA) p.vecA = aligned16(g.A) && g.lda % 4 == 0 && ct % 4 == 0 && g.K % 4 == 0;
            else           p.vecA = aligned16(g.A) && g.lda % 4 == 0 && ct % 4 == 0 && g.M % 4 == 0;
            if (!g.transB) p.vecB = aligned16(g.B) && g.ldb % 4 == 0;
            else           p.vecB = aligned16(g.B) && g.ldb % 4 == 0 && g.K % 4 == 0;
            if (g.remap_period > 0) p.vecC = aligned16(g.C) && g.remap_outer % 4 == 0 && g.remap_inner % 4 == 0;
            else                    p.vecC = aligned16(g.C) && g.ldc % 4 == 0;
            int t128 = cdiv(g.M, 128) * cdiv(g.N, 128), t64 = cdiv(g.M, 64) * cdiv(g.N, 64);
            max_tiles128 = t128 > max_tiles128 ? t128 : max_tiles128;
            max_tiles64 = t64 > max_tiles64 ? t64 : max_tiles64;
            max_split = g.split_k > max_split ? g.split_k : max_split;
            if (g.N > 64) small_n = false;
        }
        // Tile choice: 128x128 when that already fills the 148 SMs, else 64x64 for more CTAs.
        const bool big = !small_n && (long long)max_tiles128 * max_split * cnt >= 148;
        if (big) {
            dim3 grid(max_tiles128, max_split, cnt);
            gemm_simt_kernel<128, 128><<<grid, NT, 0, s>>>(b);
        } else {
            dim3 grid(max_tiles64, max_split, cnt);
            gemm_simt_kernel<64, 64><<<grid, NT, 0, s>>>(b);
        }
        TACO_CHECK_LAUNCH();
        done += cnt;
    }
    return TACO_OK;
}

}  // namespace taco
