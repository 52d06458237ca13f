// HBM-bound glue kernels of the CBHG blocks and the losses.  All operate on row-major
// [rows, C] fp32 matrices in the zero-padded time layout (rows = N*Tp, row m <-> (n, tp),
// frame t = tp - PL valid iff 0 <= t < T); C is always a multiple of 4 so accesses are 128-bit.
//
// reference semantics restated here:
//   batch-norm (train: biased batch moments over all N*T frames incl. pad frames; eval: moving
//   statistics; eps 1e-3, momentum .99)                              models/modules.py:131
//   max_pooling1d(2, 1, 'same') = max(x[t], x[t+1]), last frame alone  models/modules.py:47-51
//   residual (+ tiled before_highway)                                  models/modules.py:62-69
//   highway y = H*T + x*(1-T)                                          models/modules.py:105-120
//   L1 losses with loss_coeff and the 'prioritize' band                models/tacotron.py:274-302
#include "common.cuh"
#include "kernels.h"

namespace taco {

constexpr int EW_THREADS = 256;
typedef __nv_bfloat16 bf16;

// bf16 mirrors (TACO_PREC_BF16): the producers of every tensor that feeds a large GEMM also write it as bf16 (round to nearest
// even), so the contraction kernels never read fp32 activations; a NULL mirror pointer turns the extra store off.
__device__ __forceinline__ void st_bf16x4(bf16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float4 ld_bf16x4(const bf16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

static inline int ew_blocks(long long work, int per_block = EW_THREADS, int cap = 148 * 16) {
    long long b = (work + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (int)(b > cap ? cap : b);
}

// ---------------------------------------------------------------------------------------
__global__ void fill_kernel(float* p, long long n, float v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
int launch_fill(float* p, long long n, float v, cudaStream_t s) {
    if (n <= 0) return TACO_OK;
    fill_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(p, n, v);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// out[m,:] = valid(m) ? table[idx[n, t], :] : 0    (embedding/prenet lookup into the padded layout)
__global__ void gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ idx, float* __restrict__ out, bf16* __restrict__ out16,
                                   int N, int T, int Tp, int PL, int C, int n_rows_table) {
    const int c4 = C / 4;
    long long total = (long long)N * Tp * c4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % c4); long long m = i / c4;
        int tp = (int)(m % Tp), n = (int)(m / Tp), t = tp - PL;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T) {
            int id = idx[n * T + t];
            id = min(max(id, 0), n_rows_table - 1);
            v = __ldg(reinterpret_cast<const float4*>(table + (long long)id * C) + c);
        }
        reinterpret_cast<float4*>(out)[i] = v;
        if (out16) st_bf16x4(out16 + 4 * i, v);
    }
}
int launch_gather_rows(const float* table, const int* idx, float* out, int N, int T, int Tp, int PL, int C, int n_rows_table, cudaStream_t s, void* out16) {
    gather_rows_kernel<<<ew_blocks((long long)N * Tp * C / 4), EW_THREADS, 0, s>>>(table, idx, out, static_cast<bf16*>(out16), N, T, Tp, PL, C, n_rows_table);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// dtable[idx[n,t], :] += dx[m,:]  over valid rows (backward of the lookup)
__global__ void scatter_add_rows_kernel(const float* __restrict__ dx, const int* __restrict__ idx, float* __restrict__ dtable,
                                        int N, int T, int Tp, int PL, int C, int n_rows_table) {
    long long total = (long long)N * T * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C); long long nt = i / C;
        int t = (int)(nt % T), n = (int)(nt / T);
        int id = idx[n * T + t];
        id = min(max(id, 0), n_rows_table - 1);
        atomicAdd(dtable + (long long)id * C + c, dx[((long long)n * Tp + PL + t) * C + c]);
    }
}
int launch_scatter_add_rows(const float* dx, const int* idx, float* dtable, int N, int T, int Tp, int PL, int C, int n_rows_table, cudaStream_t s) {
    scatter_add_rows_kernel<<<ew_blocks((long long)N * T * C), EW_THREADS, 0, s>>>(dx, idx, dtable, N, T, Tp, PL, C, n_rows_table);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// ---------------------------------------------------------------------------------------
// batch-norm statistics -> (mean, rstd, var).  training: biased batch moments; eval: moving statistics.
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double count,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ var_out,
                                   const float* __restrict__ moving_mean, const float* __restrict__ moving_var,
                                   int C, int training, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (training) {
        double mu = sum[c] / count;
        double var = sumsq[c] / count - mu * mu;
        if (var < 0.0) var = 0.0;
        mean[c] = (float)mu;
        var_out[c] = (float)var;
        rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    } else {
        mean[c] = moving_mean[c];
        var_out[c] = moving_var[c];
        rstd[c] = rsqrtf(moving_var[c] + eps);
    }
}
int launch_bn_finalize(const double* sum, const double* sumsq, double count, float* mean, float* rstd, float* var,
                       const float* moving_mean, const float* moving_var, int C, int training, cudaStream_t s) {
    bn_finalize_kernel<<<cdiv(C, 128), 128, 0, s>>>(sum, sumsq, count, mean, rstd, var, moving_mean, moving_var, C, training, 1e-3f);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}
// moving <- moving*momentum + batch*(1-momentum)   (the UPDATE_OPS of tf.layers.batch_normalization; tacotron.py:332-336)
__global__ void bn_update_moving_kernel(float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                        const float* __restrict__ mean, const float* __restrict__ var, int C, float momentum) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    moving_mean[c] = moving_mean[c] * momentum + mean[c] * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + var[c] * (1.f - momentum);
}
int launch_bn_update_moving(float* moving_mean, float* moving_var, const float* mean, const float* var, int C, cudaStream_t s) {
    bn_update_moving_kernel<<<cdiv(C, 128), 128, 0, s>>>(moving_mean, moving_var, mean, var, C, 0.99f);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// BN apply.  mode 0: out = bn(x); mode 1: out = max(bn(x[t]), bn(x[t+1])) (maxpool);
// optional residual res[m,:] and per-batch-row vector rowvec[n,:] are added after.  Pad rows -> 0.
template <bool IN16>
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ res, const float* __restrict__ rowvec,
                                float* __restrict__ out, bf16* __restrict__ out16, const bf16* __restrict__ x16, int N, int T, int Tp, int PL, int C, int mode) {
    const int c4 = C / 4;
    long long total = (long long)N * Tp * c4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int cq = (int)(i % c4); long long m = i / c4;
        int tp = (int)(m % Tp), n = (int)(m / Tp), t = tp - PL;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T) {
            float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + cq);
            float4 rs = __ldg(reinterpret_cast<const float4*>(rstd) + cq);
            float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + cq);
            float4 b = __ldg(reinterpret_cast<const float4*>(beta) + cq);
            float4 v = IN16 ? ld_bf16x4(x16 + 4 * i) : __ldg(reinterpret_cast<const float4*>(x) + i);
            o.x = (v.x - mu.x) * rs.x * g.x + b.x; o.y = (v.y - mu.y) * rs.y * g.y + b.y;
            o.z = (v.z - mu.z) * rs.z * g.z + b.z; o.w = (v.w - mu.w) * rs.w * g.w + b.w;
            if (mode == 1 && t + 1 < T) {
                float4 w = IN16 ? ld_bf16x4(x16 + 4 * (i + c4)) : __ldg(reinterpret_cast<const float4*>(x) + i + c4);
                o.x = fmaxf(o.x, (w.x - mu.x) * rs.x * g.x + b.x); o.y = fmaxf(o.y, (w.y - mu.y) * rs.y * g.y + b.y);
                o.z = fmaxf(o.z, (w.z - mu.z) * rs.z * g.z + b.z); o.w = fmaxf(o.w, (w.w - mu.w) * rs.w * g.w + b.w);
            }
            if (res) {
                float4 r = __ldg(reinterpret_cast<const float4*>(res) + i);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            if (rowvec) {
                float4 r = __ldg(reinterpret_cast<const float4*>(rowvec + (long long)n * C) + cq);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
        }
        if (out) reinterpret_cast<float4*>(out)[i] = o;
        if (out16) st_bf16x4(out16 + 4 * i, o);
    }
}
int launch_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    const float* res, const float* rowvec, float* out, int N, int T, int Tp, int PL, int C, int mode, cudaStream_t s, void* out16, const void* x16) {
    TACO_REQUIRE(out || out16, TACO_EINVAL, "bn_apply: no output");
    TACO_REQUIRE(x || x16, TACO_EINVAL, "bn_apply: no input");
    if (x16) bn_apply_kernel<true><<<ew_blocks((long long)N * Tp * C / 4), EW_THREADS, 0, s>>>(x, mean, rstd, gamma, beta, res, rowvec, out, static_cast<bf16*>(out16),
                                                                                          static_cast<const bf16*>(x16), N, T, Tp, PL, C, mode);
    else bn_apply_kernel<false><<<ew_blocks((long long)N * Tp * C / 4), EW_THREADS, 0, s>>>(x, mean, rstd, gamma, beta, res, rowvec, out, static_cast<bf16*>(out16),
                                                                                        nullptr, N, T, Tp, PL, C, mode);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// ---------------------------------------------------------------------------------------
// BN backward.  dy is either given directly (mode 0) or routed through the max-pool (mode 1):
//   dy[t] = dp[t]*[b[t] >= b[t+1] or t==T-1] + dp[t-1]*[b[t] > b[t-1]]   (first max wins, as TF's CPU MaxPoolGrad)
// where b = bn(x).  Pass 1 reduces s1 = sum dy, s2 = sum dy*xhat per channel (these are dbeta, dgamma).
__device__ __forceinline__ float bn_dy(const float* __restrict__ dyp, const float* __restrict__ x, long long off, int C,
                                       int t, int T, float mu, float rs, float g, float b, int mode) {
    if (mode == 0) return dyp[off];
    float bc = (x[off] - mu) * rs * g + b;
    float d = 0.f;
    bool take_right = (t + 1 >= T);
    if (!take_right) { float bn_ = (x[off + C] - mu) * rs * g + b; take_right = (bc >= bn_); }
    if (take_right) d += dyp[off];
    if (t > 0) { float bp = (x[off - C] - mu) * rs * g + b; if (bc > bp) d += dyp[off - C]; }
    return d;
}

// Both passes use one thread per (4 channels) x (BNB_R consecutive time steps of one utterance): the neighbouring rows the
// max-pool routing needs (mode 1) are then already in registers, every load is a coalesced float4 and all of a thread's
// loads are issued before the arithmetic.  Blocks are (tx channel lanes) x (ty row lanes), 256 threads.
constexpr int BNB_R = 8;

__device__ __forceinline__ float4 f4_ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_bn(float4 x, float4 mu, float4 rs, float4 g, float4 b) {
    return make_float4((x.x - mu.x) * rs.x * g.x + b.x, (x.y - mu.y) * rs.y * g.y + b.y, (x.z - mu.z) * rs.z * g.z + b.z, (x.w - mu.w) * rs.w * g.w + b.w);
}

// IN16: x and dy are read from their bf16 mirrors (compile time: a run-time choice per load turned the 19 batched loads of a
// thread into 19 dependent branches and doubled the kernel's time)
template <int MODE, bool APPLY, bool IN16>
__global__ void __launch_bounds__(256, 2) bn_bwd_kernel(const float* __restrict__ dyp, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx,
                                                     bf16* __restrict__ dx16, float* __restrict__ dbias,
                                                     const bf16* __restrict__ x16, const bf16* __restrict__ dyp16,
                                                     int N, int T, int Tp, int PL, int C, int relu_mask, int chunks, int c_off) {
    const int tx = blockDim.x, ty = blockDim.y;
    const int c = c_off + (blockIdx.x * tx + threadIdx.x) * 4;
    const int n = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
    const bool cok = c < C;
    const int tp0 = (chunk * ty + threadIdx.y) * BNB_R;
    float4 s1 = f4_zero(), s2 = f4_zero();
    if (cok && tp0 < Tp) {
        const float4 mu = f4_ld(mean + c), rs = f4_ld(rstd + c), g = f4_ld(gamma + c), b = f4_ld(beta + c);
        const long long base = ((long long)n * Tp) * C + c;
        // rows tp0-1 .. tp0+BNB_R of x (mode 1) / tp0 .. tp0+BNB_R-1 (mode 0); rows tp0-1 .. tp0+BNB_R-1 of dy (mode 1)
        float4 xv[BNB_R + 2], dv[BNB_R + 1];
#pragma unroll
        for (int j = 0; j < BNB_R + 2; j++) {
            const int tp = tp0 - 1 + j, t = tp - PL;
            const bool need = (MODE == 1) ? (t >= 0 && t < T) : (j >= 1 && j <= BNB_R && t >= 0 && t < T);
            if (IN16) xv[j] = need ? ld_bf16x4(x16 + base + (long long)tp * C) : f4_zero();
            else xv[j] = need ? f4_ld(x + base + (long long)tp * C) : f4_zero();
        }
#pragma unroll
        for (int j = 0; j < BNB_R + 1; j++) {
            const int tp = tp0 - 1 + j, t = tp - PL;
            const bool need = (t >= 0 && t < T) && (MODE == 1 || j >= 1);
            if (IN16) dv[j] = need ? ld_bf16x4(dyp16 + base + (long long)tp * C) : f4_zero();
            else dv[j] = need ? f4_ld(dyp + base + (long long)tp * C) : f4_zero();
        }
        float4 dg4 = f4_zero(), db4 = f4_zero();
        if (APPLY) { dg4 = f4_ld(dgamma + c); db4 = f4_ld(dbeta + c); }
        const float invM = 1.0f / ((float)N * (float)T);
        float4 bnv[BNB_R + 2];
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < BNB_R + 2; j++) bnv[j] = f4_bn(xv[j], mu, rs, g, b);
        }
#pragma unroll
        for (int j = 1; j <= BNB_R; j++) {
            const int tp = tp0 - 1 + j, t = tp - PL;
            if (tp >= Tp) break;
            const bool valid = (t >= 0 && t < T);
            float4 dy = dv[j];
            if (MODE == 1) {
                // y[t] = max(bn[t], bn[t+1]) (right pad -inf): row t receives dy[t] when it wins (ties go to the first
                // element) and dy[t-1] when it beats bn[t-1] strictly
                const bool last = (t + 1 >= T), first = (t <= 0);
                const float4 bc = bnv[j], bn_ = bnv[j + 1], bp = bnv[j - 1], dp = dv[j - 1], dc = dv[j];
                dy.x = ((last || bc.x >= bn_.x) ? dc.x : 0.f) + ((!first && bc.x > bp.x) ? dp.x : 0.f);
                dy.y = ((last || bc.y >= bn_.y) ? dc.y : 0.f) + ((!first && bc.y > bp.y) ? dp.y : 0.f);
                dy.z = ((last || bc.z >= bn_.z) ? dc.z : 0.f) + ((!first && bc.z > bp.z) ? dp.z : 0.f);
                dy.w = ((last || bc.w >= bn_.w) ? dc.w : 0.f) + ((!first && bc.w > bp.w) ? dp.w : 0.f);
            }
            const float4 xc = xv[j];
            const float4 xh = make_float4((xc.x - mu.x) * rs.x, (xc.y - mu.y) * rs.y, (xc.z - mu.z) * rs.z, (xc.w - mu.w) * rs.w);
            if (!APPLY) {
                if (valid) {
                    s1.x += dy.x; s1.y += dy.y; s1.z += dy.z; s1.w += dy.w;
                    s2.x = fmaf(dy.x, xh.x, s2.x); s2.y = fmaf(dy.y, xh.y, s2.y); s2.z = fmaf(dy.z, xh.z, s2.z); s2.w = fmaf(dy.w, xh.w, s2.w);
                }
            } else {
                float4 o = f4_zero();
                if (valid) {
                    o.x = g.x * rs.x * (dy.x - db4.x * invM - xh.x * dg4.x * invM);
                    o.y = g.y * rs.y * (dy.y - db4.y * invM - xh.y * dg4.y * invM);
                    o.z = g.z * rs.z * (dy.z - db4.z * invM - xh.z * dg4.z * invM);
                    o.w = g.w * rs.w * (dy.w - db4.w * invM - xh.w * dg4.w * invM);
                    if (relu_mask) {
                        if (!(xc.x > 0.f)) o.x = 0.f;
                        if (!(xc.y > 0.f)) o.y = 0.f;
                        if (!(xc.z > 0.f)) o.z = 0.f;
                        if (!(xc.w > 0.f)) o.w = 0.f;
                    }
                }
                if (dx) *reinterpret_cast<float4*>(dx + base + (long long)tp * C) = o;
                if (dx16) st_bf16x4(dx16 + base + (long long)tp * C, o);
                s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;       // column sums of dx = gradient of the convolution bias
            }
        }
    }
    if (APPLY && dbias) {
        __shared__ float4 redb[256];
        const int tid = threadIdx.y * tx + threadIdx.x;
        redb[tid] = s1;
        __syncthreads();
        if (threadIdx.y == 0 && cok) {
            for (int y = 1; y < ty; y++) { const float4 a = redb[y * tx + threadIdx.x]; s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w; }
            atomicAdd(dbias + c, s1.x); atomicAdd(dbias + c + 1, s1.y); atomicAdd(dbias + c + 2, s1.z); atomicAdd(dbias + c + 3, s1.w);
        }
    }
    if (!APPLY) {
        __shared__ float4 red[2][256];
        const int tid = threadIdx.y * tx + threadIdx.x;
        red[0][tid] = s1; red[1][tid] = s2;
        __syncthreads();
        if (threadIdx.y == 0 && cok) {
            for (int y = 1; y < ty; y++) {
                const float4 a = red[0][y * tx + threadIdx.x], q = red[1][y * tx + threadIdx.x];
                s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
                s2.x += q.x; s2.y += q.y; s2.z += q.z; s2.w += q.w;
            }
            atomicAdd(dbeta + c, s1.x); atomicAdd(dbeta + c + 1, s1.y); atomicAdd(dbeta + c + 2, s1.z); atomicAdd(dbeta + c + 3, s1.w);
            atomicAdd(dgamma + c, s2.x); atomicAdd(dgamma + c + 1, s2.y); atomicAdd(dgamma + c + 2, s2.z); atomicAdd(dgamma + c + 3, s2.w);
        }
    }
}

// lanes over float4 channel groups (power of two, <= 64) x row lanes, 256 threads
static inline void lanes_2d(int C4, int& tx, int& ty) {
    tx = 1; while (tx < C4 && tx < 64) tx <<= 1;
    if (tx < 8) tx = 8;
    ty = 256 / tx;
}

int launch_bn_bwd(const float* dyp, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  float* dgamma, float* dbeta, float* dx, int N, int T, int Tp, int PL, int C, int mode, int relu_mask, cudaStream_t s,
                  void* dx16v, float* dbias, const void* x16v, const void* dyp16v) {
    TACO_REQUIRE(C % 4 == 0, TACO_ESHAPE, "bn_bwd: channel count %d must be a multiple of 4", C);
    TACO_REQUIRE(dx || dx16v, TACO_EINVAL, "bn_bwd: no output");
    bf16* dx16 = static_cast<bf16*>(dx16v);
    const bf16* x16 = static_cast<const bf16*>(x16v); const bf16* dyp16 = static_cast<const bf16*>(dyp16v);
    TACO_REQUIRE((x && dyp && !x16 && !dyp16) || (x16 && dyp16), TACO_EINVAL, "bn_bwd: inputs must be both fp32 or both bf16");
    const bool in16 = x16 != nullptr;
    int tx, ty; lanes_2d(C / 4, tx, ty);
    const int chunks = cdiv(Tp, ty * BNB_R);
    const int gx = cdiv(C / 4, tx);
    // (walking wide tensors in L2-sized column groups, reduce then apply per group, was measured SLOWER on B200: 1 KB row
    // segments out of 8 KB rows waste DRAM pages; 0.65 ms vs 0.34 ms for the C2 conv bank.  Kept switchable for re-measurement.)
    const bool grouped = false;
    const int ngroups = grouped ? gx : 1;
    dim3 block(tx, ty), grid(grouped ? 1 : gx, (unsigned)(N * chunks));
    for (int gi = 0; gi < ngroups; gi++) {
        const int c_off = gi * 4 * tx;
#define BN_BWD_LAUNCH(MODE_, APPLY_, IN16_)                                                                                             \
        bn_bwd_kernel<MODE_, APPLY_, IN16_><<<grid, block, 0, s>>>(dyp, x, mean, rstd, gamma, beta, dgamma, dbeta, dx, dx16, dbias, x16, dyp16, \
                                                                    N, T, Tp, PL, C, relu_mask, chunks, c_off)
        if (mode == 1 && in16) { BN_BWD_LAUNCH(1, false, true); BN_BWD_LAUNCH(1, true, true); }
        else if (mode == 1) { BN_BWD_LAUNCH(1, false, false); BN_BWD_LAUNCH(1, true, false); }
        else if (in16) { BN_BWD_LAUNCH(0, false, true); BN_BWD_LAUNCH(0, true, true); }
        else { BN_BWD_LAUNCH(0, false, false); BN_BWD_LAUNCH(0, true, false); }
#undef BN_BWD_LAUNCH
    }
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// ---------------------------------------------------------------------------------------
// highway combine: y = H*T + x*(1-T)
__global__ void highway_fwd_kernel(const float* __restrict__ H, const float* __restrict__ Tg, const float* __restrict__ x,
                                   float* __restrict__ y, bf16* __restrict__ y16, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 h = reinterpret_cast<const float4*>(H)[i], t = reinterpret_cast<const float4*>(Tg)[i], v = reinterpret_cast<const float4*>(x)[i];
        float4 o;
        o.x = h.x * t.x + v.x * (1.f - t.x); o.y = h.y * t.y + v.y * (1.f - t.y);
        o.z = h.z * t.z + v.z * (1.f - t.z); o.w = h.w * t.w + v.w * (1.f - t.w);
        reinterpret_cast<float4*>(y)[i] = o;
        if (y16) st_bf16x4(y16 + 4 * i, o);
    }
}
int launch_highway_fwd(const float* H, const float* Tg, const float* x, float* y, long long n, cudaStream_t s, void* y16) {
    highway_fwd_kernel<<<ew_blocks(n / 4), EW_THREADS, 0, s>>>(H, Tg, x, y, static_cast<bf16*>(y16), n / 4);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}
// backward: dHpre = dy*T*[H>0]; dTpre = dy*(H-x)*T*(1-T); dx_direct = dy*(1-T)
// The two pre-activation gradients are written side by side into one [rows, 2C] matrix (dHpre | dTpre) so that the data
// gradient dx += dHpre.WH^T + dTpre.WT^T is ONE GEMM with K = 2C against the packed [C, 2C] weight (model_cbhg.cu).
__global__ void __launch_bounds__(256) highway_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ H, const float* __restrict__ Tg,
                                   const float* __restrict__ x, float* __restrict__ dHT, bf16* __restrict__ dHT16, float* __restrict__ dx,
                                   float* __restrict__ dbH, float* __restrict__ dbT, long long rows, int c4, int rows_per_block) {
    // block = (tx float4 channel lanes) x (ty rows); a block walks its row chunk, every thread keeps its 4 channels: the column sums
    // of the two pre-activation gradients (= the H / T bias gradients) come out of the same pass
    const int tx = blockDim.x, ty = blockDim.y;
    const int cq = blockIdx.x * tx + threadIdx.x;
    const bool cok = cq < c4;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cok) {
        for (long long row = r0 + threadIdx.y; row < r1; row += ty) {
            const long long i = row * c4 + cq;
            const float4 d = reinterpret_cast<const float4*>(dy)[i], h = reinterpret_cast<const float4*>(H)[i];
            const float4 t = reinterpret_cast<const float4*>(Tg)[i], v = reinterpret_cast<const float4*>(x)[i];
            float4 a, b, o;
            a.x = (h.x > 0.f) ? d.x * t.x : 0.f; a.y = (h.y > 0.f) ? d.y * t.y : 0.f;
            a.z = (h.z > 0.f) ? d.z * t.z : 0.f; a.w = (h.w > 0.f) ? d.w * t.w : 0.f;
            b.x = d.x * (h.x - v.x) * t.x * (1.f - t.x); b.y = d.y * (h.y - v.y) * t.y * (1.f - t.y);
            b.z = d.z * (h.z - v.z) * t.z * (1.f - t.z); b.w = d.w * (h.w - v.w) * t.w * (1.f - t.w);
            o.x = d.x * (1.f - t.x); o.y = d.y * (1.f - t.y); o.z = d.z * (1.f - t.z); o.w = d.w * (1.f - t.w);
            if (dHT) { float4* out = reinterpret_cast<float4*>(dHT) + row * (2 * c4) + cq; out[0] = a; out[c4] = b; }
            if (dHT16) { bf16* o16 = dHT16 + (row * (2 * c4) + cq) * 4; st_bf16x4(o16, a); st_bf16x4(o16 + 4 * c4, b); }
            reinterpret_cast<float4*>(dx)[i] = o;
            sa.x += a.x; sa.y += a.y; sa.z += a.z; sa.w += a.w; sb.x += b.x; sb.y += b.y; sb.z += b.z; sb.w += b.w;
        }
    }
    if (dbH) {        // kernel-uniform
        __shared__ float4 red[2][256];
        const int tid = threadIdx.y * tx + threadIdx.x;
        red[0][tid] = sa; red[1][tid] = sb;
        __syncthreads();
        if (threadIdx.y == 0 && cok) {
            for (int y = 1; y < ty; y++) {
                const float4 p = red[0][y * tx + threadIdx.x], q = red[1][y * tx + threadIdx.x];
                sa.x += p.x; sa.y += p.y; sa.z += p.z; sa.w += p.w; sb.x += q.x; sb.y += q.y; sb.z += q.z; sb.w += q.w;
            }
            const int c = cq * 4;
            atomicAdd(dbH + c, sa.x); atomicAdd(dbH + c + 1, sa.y); atomicAdd(dbH + c + 2, sa.z); atomicAdd(dbH + c + 3, sa.w);
            atomicAdd(dbT + c, sb.x); atomicAdd(dbT + c + 1, sb.y); atomicAdd(dbT + c + 2, sb.z); atomicAdd(dbT + c + 3, sb.w);
        }
    }
}
int launch_highway_bwd(const float* dy, const float* H, const float* Tg, const float* x, float* dHT, float* dx,
                       long long rows, int C, cudaStream_t s, void* dHT16, float* dbH, float* dbT) {
    TACO_REQUIRE(C % 4 == 0, TACO_ESHAPE, "highway_bwd: width %d must be a multiple of 4", C);
    TACO_REQUIRE(dHT || dHT16, TACO_EINVAL, "highway_bwd: no pre-activation gradient output");
    TACO_REQUIRE((dbH == nullptr) == (dbT == nullptr), TACO_EINVAL, "highway_bwd: give both bias gradients or none");
    int tx, ty; lanes_2d(C / 4, tx, ty);
    const int bx = cdiv(C / 4, tx);
    long long by = cdiv64(148 * 8, bx);
    long long rpb = cdiv64(rows, by);
    if (rpb < ty) rpb = ty;
    dim3 block(tx, ty), grid(bx, (unsigned)cdiv64(rows, rpb));
    highway_bwd_kernel<<<grid, block, 0, s>>>(dy, H, Tg, x, dHT, static_cast<bf16*>(dHT16), dx, dbH, dbT, rows, C / 4, (int)rpb);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// ---------------------------------------------------------------------------------------
// column sums: out[c] += sum_m x[m*ld + c]   (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long M, int C, int ld, int rows_per_block) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s = 0.f;
    for (long long r = r0; r < r1; r++) s += x[r * ld + c];
    atomicAdd(out + c, s);
}
// float4 variant: (tx channel lanes) x (ty row lanes); needs 16-byte aligned rows (ld % 4 == 0, ld >= 4*ceil(C/4))
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ x, float* __restrict__ out, long long M, int C, int ld,
                                                      int rows_per_block) {
    const int tx = blockDim.x, ty = blockDim.y;
    const int c = (blockIdx.x * tx + threadIdx.x) * 4;
    const bool cok = c < C;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cok) {
        long long r = r0 + threadIdx.y;
        for (; r + 3LL * ty < r1; r += 4LL * ty) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(x + (r + ty) * ld + c));
            const float4 d = __ldg(reinterpret_cast<const float4*>(x + (r + 2LL * ty) * ld + c));
            const float4 e = __ldg(reinterpret_cast<const float4*>(x + (r + 3LL * ty) * ld + c));
            s.x += (a.x + b.x) + (d.x + e.x); s.y += (a.y + b.y) + (d.y + e.y);
            s.z += (a.z + b.z) + (d.z + e.z); s.w += (a.w + b.w) + (d.w + e.w);
        }
        for (; r < r1; r += ty) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    __shared__ float4 red[256];
    red[threadIdx.y * tx + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && cok) {
        for (int y = 1; y < ty; y++) { const float4 a = red[y * tx + threadIdx.x]; s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
        atomicAdd(out + c, s.x);
        if (c + 1 < C) atomicAdd(out + c + 1, s.y);
        if (c + 2 < C) atomicAdd(out + c + 2, s.z);
        if (c + 3 < C) atomicAdd(out + c + 3, s.w);
    }
}
int launch_colsum(const float* x, float* out, long long M, int C, int ld, cudaStream_t s) {
    const int C4 = cdiv(C, 4);
    if (ld % 4 == 0 && ld >= 4 * C4 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        int tx, ty; lanes_2d(C4, tx, ty);
        const int bx = cdiv(C4, tx);
        long long by = cdiv64(148 * 8, bx);
        long long rpb = cdiv64(M, by);
        if (rpb < 4LL * ty) rpb = 4LL * ty;
        rpb = cdiv64(rpb, ty) * ty;
        dim3 block(tx, ty), grid(bx, (unsigned)cdiv64(M, rpb));
        colsum4_kernel<<<grid, block, 0, s>>>(x, out, M, C, ld, (int)rpb);
    } else {
        const int rows_per_block = 128;
        dim3 grid(cdiv(C, 128), (unsigned)cdiv64(M, rows_per_block));
        colsum_kernel<<<grid, 128, 0, s>>>(x, out, M, C, ld, rows_per_block);
    }
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// bf16 input variant (bias gradients over a bf16 gradient mirror): 8 columns (16 bytes) per thread, (tx column lanes) x
// (ty row lanes), four row loads in flight per thread; one fp32 atomic per column and block.
__global__ void __launch_bounds__(256) colsum16_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long M, int C, long long ld,
                                                       int rows_per_block) {
    const int tx = blockDim.x, ty = blockDim.y;
    const int c = (blockIdx.x * tx + threadIdx.x) * 8;
    const bool cok = c < C;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cok) {
        auto add = [&](const uint4 u) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; e++) { const float2 f = __bfloat1622float2(h[e]); s[2 * e] += f.x; s[2 * e + 1] += f.y; }
        };
        long long r = r0 + threadIdx.y;
        for (; r + 3LL * ty < r1; r += 4LL * ty) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(x + r * ld + c));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(x + (r + ty) * ld + c));
            const uint4 d = __ldg(reinterpret_cast<const uint4*>(x + (r + 2LL * ty) * ld + c));
            const uint4 e = __ldg(reinterpret_cast<const uint4*>(x + (r + 3LL * ty) * ld + c));
            add(a); add(b); add(d); add(e);
        }
        for (; r < r1; r += ty) add(__ldg(reinterpret_cast<const uint4*>(x + r * ld + c)));
    }
    __shared__ float red[256][9];
    const int tid = threadIdx.y * tx + threadIdx.x;
#pragma unroll
    for (int e = 0; e < 8; e++) red[tid][e] = s[e];
    __syncthreads();
    if (threadIdx.y == 0 && cok) {
        for (int y = 1; y < ty; y++)
#pragma unroll
            for (int e = 0; e < 8; e++) s[e] += red[y * tx + threadIdx.x][e];
#pragma unroll
        for (int e = 0; e < 8; e++) if (c + e < C) atomicAdd(out + c + e, s[e]);
    }
}
// out[c] += sum_m x[m*ld + c], x bf16 with 16-byte aligned rows (ld % 8 == 0, base aligned, ld >= 8*ceil(C/8))
int launch_colsum16(const void* x, float* out, long long M, int C, long long ld, cudaStream_t s) {
    TACO_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld >= 8LL * cdiv(C, 8), TACO_EINVAL, "colsum16: rows must be 16-byte aligned");
    const int C8 = cdiv(C, 8);
    int tx = 1; while (tx < C8 && tx < 32) tx <<= 1;
    if (tx < 4) tx = 4;
    const int ty = 256 / tx;
    const int bx = cdiv(C8, tx);
    long long by = cdiv64(148 * 8, bx);
    long long rpb = cdiv64(M, by);
    if (rpb < 4LL * ty) rpb = 4LL * ty;
    rpb = cdiv64(rpb, ty) * ty;
    dim3 block(tx, ty), grid(bx, (unsigned)cdiv64(M, rpb));
    colsum16_kernel<<<grid, block, 0, s>>>(static_cast<const bf16*>(x), out, M, C, ld, (int)rpb);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// per-column sum and sum of squares of a dense [rows, C] matrix in double (stand-alone batch-norm statistics, taco_batch_norm)
__global__ void colstats_kernel(const float* __restrict__ x, double* __restrict__ sum, double* __restrict__ sumsq, long long rows, int C, int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, rows);
    double s = 0.0, q = 0.0;
    for (long long r = r0; r < r1; r++) { const double v = x[r * C + c]; s += v; q += v * v; }
    atomicAdd(sum + c, s); atomicAdd(sumsq + c, q);
}
int launch_colstats(const float* x, double* sum, double* sumsq, long long rows, int C, cudaStream_t s) {
    const int rpb = 64;
    dim3 grid(cdiv(C, 128), (unsigned)cdiv64(rows, rpb));
    colstats_kernel<<<grid, 128, 0, s>>>(x, sum, sumsq, rows, C, rpb);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// fp32 -> bf16 strided 2-D cast: dst[r*ldd + c] = bf16(src[r*lds + c])   (parameter mirror, packed weights, recurrence outputs)
__global__ void cast2d_kernel(bf16* __restrict__ dst, const float* __restrict__ src, long long rows, int cols, long long ldd, long long lds) {
    long long total = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / cols; int c = (int)(i % cols);
        dst[r * ldd + c] = __float2bfloat16_rn(src[r * lds + c]);
    }
}
__global__ void cast2d_v4_kernel(bf16* __restrict__ dst, const float* __restrict__ src, long long rows, int cols4, long long ldd, long long lds) {
    long long total = rows * cols4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / cols4; int c = (int)(i % cols4) * 4;
        st_bf16x4(dst + r * ldd + c, __ldg(reinterpret_cast<const float4*>(src + r * lds + c)));
    }
}
int launch_cast2d_bf16(void* dst, const float* src, long long rows, int cols, long long ldd, long long lds, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return TACO_OK;
    if (cols % 4 == 0 && ldd % 4 == 0 && lds % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0)
        cast2d_v4_kernel<<<ew_blocks(rows * (cols / 4)), EW_THREADS, 0, s>>>(static_cast<bf16*>(dst), src, rows, cols / 4, ldd, lds);
    else
        cast2d_kernel<<<ew_blocks(rows * cols), EW_THREADS, 0, s>>>(static_cast<bf16*>(dst), src, rows, cols, ldd, lds);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// out[n, c] += sum_t x[(n*Tp + PL + t)*C + c]   (gradient of a vector tiled over time)
__global__ void timesum_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int Tp, int PL, int C) {
    int n = blockIdx.y;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int t = 0; t < T; t++) s += x[((long long)n * Tp + PL + t) * C + c];
    out[(long long)n * C + c] += s;
}
int launch_timesum(const float* x, float* out, int N, int T, int Tp, int PL, int C, cudaStream_t s) {
    dim3 grid(cdiv(C, 128), N);
    timesum_kernel<<<grid, 128, 0, s>>>(x, out, T, Tp, PL, C);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// dst[(n*T + t)*ld + c] = src[n*ld + c]   (a per-utterance vector tiled over time; the speaker term of a dense layer)
__global__ void bcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int C, long long ld, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long row = i / C; int c = (int)(i % C);
        dst[row * ld + c] = __ldg(src + (row / T) * ld + c);
    }
}
int launch_bcast_rows(const float* src, float* dst, int N, int T, int C, long long ld, cudaStream_t s) {
    const long long total = (long long)N * T * C;
    if (total <= 0) return TACO_OK;
    bcast_rows_kernel<<<ew_blocks(total), EW_THREADS, 0, s>>>(src, dst, T, C, ld, total);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// y[i] += a * x[i]
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}
int launch_axpy(float* y, const float* x, float a, long long n, cudaStream_t s) {
    if (n <= 0) return TACO_OK;
    axpy_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(y, x, a, n);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// strided 2-D copy: dst[r*ldd + c] = src[r*lds + c]
__global__ void copy2d_kernel(float* __restrict__ dst, const float* __restrict__ src, long long rows, int cols, long long ldd, long long lds) {
    long long total = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / cols; int c = (int)(i % cols);
        dst[r * ldd + c] = src[r * lds + c];
    }
}
int launch_copy2d(float* dst, const float* src, long long rows, int cols, long long ldd, long long lds, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return TACO_OK;
    copy2d_kernel<<<ew_blocks(rows * cols), EW_THREADS, 0, s>>>(dst, src, rows, cols, ldd, lds);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// dst[(n*T+t)*C + c] = src[(n*Tp+PL+t)*ld + c]   (valid rows of a padded-layout matrix, dense)
__global__ void unpad_kernel(float* __restrict__ dst, const float* __restrict__ src, int N, int T, int Tp, int PL, int C, long long ld) {
    long long total = (long long)N * T * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C); long long nt = i / C;
        int t = (int)(nt % T), n = (int)(nt / T);
        dst[i] = src[((long long)n * Tp + PL + t) * ld + c];
    }
}
// float4 rows: block = (tx float4 lanes) x (ty rows), grid-stride over rows
__global__ void __launch_bounds__(256) unpad4_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int T, int Tp, int PL,
                                                     int C4, long long ld) {
    const int ty = blockDim.y;
    for (int row = blockIdx.x * ty + threadIdx.y; row < rows; row += gridDim.x * ty) {
        const int n = row / T, t = row - n * T;
        const float* sp = src + ((long long)n * Tp + PL + t) * ld;
        float* dp = dst + (long long)row * C4 * 4;
        for (int c = threadIdx.x; c < C4; c += blockDim.x)
            reinterpret_cast<float4*>(dp)[c] = __ldg(reinterpret_cast<const float4*>(sp) + c);
    }
}
int launch_unpad(float* dst, const float* src, int N, int T, int Tp, int PL, int C, long long ld, cudaStream_t s) {
    if (C % 4 == 0 && ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        int tx, ty; lanes_2d(C / 4, tx, ty);
        const int rows = N * T;
        int blocks = cdiv(rows, ty); if (blocks > 148 * 16) blocks = 148 * 16;
        unpad4_kernel<<<blocks, dim3(tx, ty), 0, s>>>(dst, src, rows, T, Tp, PL, C / 4, ld);
    } else {
        unpad_kernel<<<ew_blocks((long long)N * T * C), EW_THREADS, 0, s>>>(dst, src, N, T, Tp, PL, C, ld);
    }
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// conv data-gradient operand: Wd[j'][co][ci] = W[k-1-j'][ci][co]   (flip taps, transpose channels)
__global__ void pack_dgrad_kernel(const float* __restrict__ W, float* __restrict__ Wd, int k, int Cin, int Cout) {
    long long total = (long long)k * Cin * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ci = (int)(i % Cin); long long r = i / Cin;
        int co = (int)(r % Cout), jp = (int)(r / Cout);
        Wd[i] = W[((long long)(k - 1 - jp) * Cin + ci) * Cout + co];
    }
}
int launch_pack_dgrad(const float* W, float* Wd, int k, int Cin, int Cout, cudaStream_t s) {
    pack_dgrad_kernel<<<ew_blocks((long long)k * Cin * Cout), EW_THREADS, 0, s>>>(W, Wd, k, Cin, Cout);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// ---- table-driven operand preparation ----------------------------------------------------------------------------------
// The backward pass reads ~80 small parameter-only operands (flipped / transposed convolution kernels, concatenated GRU and
// highway weights, transposed attention weights; in bf16 mode also their bf16 mirrors).  One launch per operand and per mirror
// was 78 launches of ~5 us per step beside the forward pass; a table of gather descriptors runs them in one launch per block
// of the model (grid.y = operand), each element written as fp32 and - where the operand has a mirror - as bf16 in the same pass.
__global__ void __launch_bounds__(256) prep_ops_kernel(const __grid_constant__ PrepTable tb) {
    const PrepOp& o = tb.op[blockIdx.y];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < o.total; i += (long long)gridDim.x * blockDim.x) {
        long long si, di;
        if (o.kind == PREP_PACK_DGRAD) {            // Wd[j'][co][ci] = W[k-1-j'][ci][co]       a = k, b = Cin, c = Cout
            const int ci = (int)(i % o.b); const long long r = i / o.b;
            const int co = (int)(r % o.c), jp = (int)(r / o.c);
            si = ((long long)(o.a - 1 - jp) * o.b + ci) * o.c + co; di = i;
        } else if (o.kind == PREP_COPY2D) {          // dst[r*ldd + c] = src[r*lds + c]         b = cols
            const long long r = i / o.b; const int c = (int)(i % o.b);
            si = r * o.lds + c; di = r * o.ldd + c;
        } else {                                     // transpose: out[c*rows + r] = in[r*cols + c]   a = rows, b = cols
            const int c = (int)(i / o.a), r = (int)(i % o.a);
            si = (long long)r * o.b + c; di = i;
        }
        const float v = __ldg(o.src + si);
        o.dst[di] = v;
        if (o.dst16) static_cast<bf16*>(o.dst16)[di] = __float2bfloat16_rn(v);
    }
}
int launch_prep_ops(const PrepTable& tb, cudaStream_t s) {
    if (tb.n <= 0) return TACO_OK;
    TACO_REQUIRE(tb.n <= PREP_MAX_OPS, TACO_EINVAL, "prep table: %d operands exceed %d", tb.n, PREP_MAX_OPS);
    prep_ops_kernel<<<dim3(96, tb.n), 256, 0, s>>>(tb);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// ---------------------------------------------------------------------------------------
// L1 loss + gradient.  out/target rows: out row (n,t) at out + (n*out_bs + t*out_ts), target at tgt + (n*T + t)*C.
// scalars[0] += sum |d|*coeff*w ; scalars[1] += sum |d| (unweighted, all bins) ; scalars[2] += sum |d| over the priority band.
// grad[(n,t),c] = sign(out - tgt) * coeff[n] * (w_all + w_band*[lo<=c<hi])     (pad rows of grad untouched)
template <bool TGT16>
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ out, long long out_bs, long long out_ts,
                               const float* __restrict__ tgt, const bf16* __restrict__ tgt16, const float* __restrict__ coeff,
                               float* __restrict__ grad, bf16* __restrict__ grad16, long long grad_bs, long long grad_ts,
                               int N, int T, int C, float w_all, float w_band, int lo, int hi, double* __restrict__ scalars) {
    // block = (tx column lanes) x (ty rows); rows (n,t) are strided over the grid, columns over tx: no per-element division
    const int tx = blockDim.x, ty = blockDim.y, rows = N * T;
    float acc_w = 0.f, acc_all = 0.f, acc_band = 0.f;
    for (int row = blockIdx.x * ty + threadIdx.y; row < rows; row += gridDim.x * ty) {
        const int n = row / T, t = row - n * T;
        const float* op = out + n * out_bs + t * out_ts;
        const float* tp = tgt ? tgt + (long long)row * C : nullptr;
        const bf16* tp16 = tgt16 ? tgt16 + (long long)row * C : nullptr;
        float* gp = grad ? grad + n * grad_bs + t * grad_ts : nullptr;
        bf16* gp16 = grad16 ? grad16 + n * grad_bs + t * grad_ts : nullptr;
        const float cf = coeff ? __ldg(coeff + n) : 1.f;
        float rw = 0.f;
        for (int c0 = threadIdx.x; c0 < C; c0 += 4 * tx) {       // four columns per trip: all eight loads issued before the arithmetic
            float ov[4], tv[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int c = c0 + e * tx; const bool ok = c < C;
                ov[e] = ok ? op[c] : 0.f; tv[e] = ok ? (TGT16 ? __bfloat162float(tp16[c]) : __ldg(tp + c)) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int c = c0 + e * tx;
                if (c < C) {
                    const float d = ov[e] - tv[e];
                    const float a = fabsf(d);
                    const bool band = (c >= lo && c < hi);
                    const float w = w_all + (band ? w_band : 0.f);
                    rw = fmaf(a, w, rw); acc_all += a; if (band) acc_band += a;
                    const float gv = ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f)) * cf * w;
                    if (gp) gp[c] = gv;
                    if (gp16) gp16[c] = __float2bfloat16_rn(gv);
                }
            }
        }
        acc_w = fmaf(rw, cf, acc_w);
    }
    acc_w = warp_sum(acc_w); acc_all = warp_sum(acc_all); acc_band = warp_sum(acc_band);
    __shared__ float sh[3][8];
    const int tid = threadIdx.y * tx + threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (lane == 0) { sh[0][wid] = acc_w; sh[1][wid] = acc_all; sh[2][wid] = acc_band; }
    __syncthreads();
    if (tid < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sh[tid][w];
        atomicAdd(scalars + tid, s);
    }
}
int launch_l1_loss(const float* out, long long out_bs, long long out_ts, const float* tgt, const float* coeff,
                   float* grad, long long grad_bs, long long grad_ts, int N, int T, int C,
                   float w_all, float w_band, int lo, int hi, double* scalars, cudaStream_t s, void* grad16, int tgt_is_bf16) {
    int tx = 32; while (tx < C && tx < 256) tx <<= 1;
    const int ty = 256 / tx;
    int blocks = cdiv(N * T, ty); if (blocks > 148 * 8) blocks = 148 * 8;
    if (tgt_is_bf16)
        l1_loss_kernel<true><<<blocks, dim3(tx, ty), 0, s>>>(out, out_bs, out_ts, nullptr, reinterpret_cast<const bf16*>(tgt), coeff, grad, static_cast<bf16*>(grad16),
                                                             grad_bs, grad_ts, N, T, C, w_all, w_band, lo, hi, scalars);
    else
        l1_loss_kernel<false><<<blocks, dim3(tx, ty), 0, s>>>(out, out_bs, out_ts, tgt, nullptr, coeff, grad, static_cast<bf16*>(grad16),
                                                              grad_bs, grad_ts, N, T, C, w_all, w_band, lo, hi, scalars);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

}  // namespace taco

namespace taco {
// dx = (y > 0) ? dy : 0   (ReLU backward on a stored post-activation value)
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = (y[i] > 0.f) ? dy[i] : 0.f;
}
int launch_relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t s) {
    if (n <= 0) return TACO_OK;
    relu_bwd_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(dy, y, dx, n);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// dx = dy * (1 - |y|)^2 with y = softsign(x) = x / (1 + |x|)
__global__ void softsign_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float t = 1.f - fabsf(y[i]);
        dx[i] = dy[i] * t * t;
    }
}
int launch_softsign_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t s) {
    if (n <= 0) return TACO_OK;
    softsign_bwd_kernel<<<ew_blocks(n), EW_THREADS, 0, s>>>(dy, y, dx, n);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// teacher-forcing inputs: x_all[(n*Td+t), :] = t==0 ? 0 : mel_targets[n, t*r-1, :]     (helpers.py:44,66,70-72)
__global__ void teacher_inputs_kernel(const float* __restrict__ tgt, float* __restrict__ x, int N, int Td, int To, int r, int M) {
    long long total = (long long)N * Td * M;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % M); long long nt = i / M;
        int t = (int)(nt % Td), n = (int)(nt / Td);
        x[i] = (t == 0) ? 0.f : tgt[((long long)n * To + (long long)t * r - 1) * M + c];
    }
}
int launch_teacher_inputs(const float* tgt, float* x, int N, int Td, int To, int r, int M, cudaStream_t s) {
    teacher_inputs_kernel<<<ew_blocks((long long)N * Td * M), EW_THREADS, 0, s>>>(tgt, x, N, Td, To, r, M);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}
}  // namespace taco

namespace taco {
// out[c*rows + r] = in[r*cols + c]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * cols + c] : 0.f;
    }
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int cc = c0 + i;
        if (r < rows && cc < cols) out[(long long)cc * rows + r] = tile[threadIdx.x][i];
    }
}
int launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t s) {
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}
}  // namespace taco
