// Global-norm clip + TF-style Adam over the flat parameter buffer (HBM-bound, one pass each).
// reference: models/tacotron.py:327-336 — tf.clip_by_global_norm(gradients, 1.0) then
// tf.train.AdamOptimizer(lr, beta1, beta2).apply_gradients:
//     lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr_t * m / (sqrt(v) + eps)
// (epsilon OUTSIDE the bias-corrected sqrt — this differs from torch.optim.Adam).
#include "common.cuh"
#include "kernels.h"

namespace taco {

__global__ void sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
    double acc = 0.0;
    const long long n4 = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 * 4; i < n; i++) acc += (double)g[i] * g[i];
    acc = warp_sum_d(acc);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sh[w];
        atomicAdd(out, s);
    }
}
int launch_sqnorm(const float* g, long long n, double* out, cudaStream_t s) {
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    sqnorm_kernel<<<(int)blocks, 256, 0, s>>>(g, n, out);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

// sqnorm holds sum(g^2) of the UNSCALED gradient; gscale (1/world for data parallel) is applied here.
__global__ void adam_clip_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                                 long long n, const double* __restrict__ sqnorm, float gscale, float clip_norm,
                                 float lr_t, float b1, float b2, float eps, float lr, float* __restrict__ norm_out) {
    const float gn = (float)(sqrt(*sqnorm) * (double)gscale);
    // tf.clip_by_global_norm: g * clip_norm * min(1/global_norm, 1/clip_norm)
    const float sc = gscale * clip_norm * fminf(1.0f / fmaxf(gn, 1e-30f), 1.0f / clip_norm);
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) { norm_out[0] = gn; norm_out[1] = lr; }
    const long long n4 = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i], pv = reinterpret_cast<float4*>(p)[i];
        float ge[4] = {gv.x * sc, gv.y * sc, gv.z * sc, gv.w * sc};
        float me[4] = {mv.x, mv.y, mv.z, mv.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w}, pe[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            me[e] = b1 * me[e] + (1.f - b1) * ge[e];
            ve[e] = b2 * ve[e] + (1.f - b2) * ge[e] * ge[e];
            pe[e] -= lr_t * me[e] / (sqrtf(ve[e]) + eps);
        }
        reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
        reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long i = n4 * 4; i < n; i++) {
            float ge = g[i] * sc;
            m[i] = b1 * m[i] + (1.f - b1) * ge;
            v[i] = b2 * v[i] + (1.f - b2) * ge * ge;
            p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
        }
    }
}
int launch_adam_clip(float* p, float* m, float* v, const float* g, long long n, const double* sqnorm, float gscale,
                     float clip_norm, float lr_t, float b1, float b2, float eps, float lr, float* norm_out, cudaStream_t s) {
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adam_clip_kernel<<<(int)blocks, 256, 0, s>>>(p, m, v, g, n, sqnorm, gscale, clip_norm, lr_t, b1, b2, eps, lr, norm_out);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

}  // namespace taco
