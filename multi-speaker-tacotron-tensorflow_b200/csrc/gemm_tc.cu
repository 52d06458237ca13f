// Tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma kind::tf32
// with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.  It serves the same descriptor (taco_gemm_desc) as the
// SIMT kernel: dense layers, conv1d as implicit GEMM (tap addressing expressed as TMA coordinates over the zero-padded
// activation buffer), data gradients and split-K weight gradients, with bias / activation / pad-row mask / row remap /
// batch-norm column statistics fused in the epilogue.      reference ops covered: see gemm_simt.cu header.
//
// Operands stay fp32 in HBM (TF32 rounding happens inside the tensor core), so no cast or repack pass exists:
//   A(m,k) row-major [M,K]      -> K-major  SW128 tile  (one TMA box  {32 k, 128 m})
//   A^T    stored [K(rows), M]  -> MN-major SW128 tile  (four TMA boxes {32 m, 32 k})     (weight gradients)
//   B[k*ldb+n] (TF [in,out])    -> MN-major SW128 tile  (four TMA boxes {32 n, 32 k})
//   B[n*ldb+k]                  -> K-major  SW128 tile  (one TMA box  {32 k, 128 n})     (data gradients)
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2-9 epilogue (two per TMEM
// lane quarter).  3-stage mbarrier ring; one 128x128 output tile (x one K split) per CTA, two CTAs per SM.
#include "tc_common.cuh"
#include "kernels.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>

namespace taco {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3, TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;                 // 16 KB per operand per stage
constexpr int TC_SMEM = TC_STAGES * 2 * TC_TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TcParams {
    float* C;
    int M, N, K, ldc;
    int a_mn_major, b_mn_major;          // 1: MN-major (four boxes), 0: K-major (one box)
    int a_tap, a_ctap;                   // strided-tap addressing for A (lda > ctap)
    int tap_inner;                       // > 0: number of taps; the K loop then walks (channel block, tap) with the tap innermost
    const int2* tap_table;               // optional per-k-tile (column, row offset) of A (K-major A only)
    float alpha; int accumulate;
    const float* bias; int act;
    int mask_period, mask_lo, mask_hi;
    int remap_period; long long remap_outer, remap_inner;
    double* colsum; double* colsumsq;
    int split_k, vecC;
};

// lane l ends with sum over the warp's lanes of v[l]
__device__ __forceinline__ float warp_transpose_reduce(float v[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < s; i++) {
            const bool up = (lane & s) != 0;
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// One 32x32 chunk of the output tile: rows {4i+lr}, columns gn..gn+3 per lane, read back from the staging buffer.
template <int ACT, bool ATOMIC>
__device__ __forceinline__ void epi_chunk(const TcParams& p, uint32_t stg, float* const crow[8], uint32_t okmask, uint32_t maskmask,
                                          int lr, int lc, int gn, bool add_bias, float cs[4], float cq[4]) {
    const bool full4 = gn + 3 < p.N;
    const bool vec = p.vecC && full4;
    float bz[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias && add_bias) {
#pragma unroll
        for (int e = 0; e < 4; e++) if (gn + e < p.N) bz[e] = __ldg(p.bias + gn + e);
    }
    const float alpha = p.alpha;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (!((okmask >> i) & 1u)) continue;
        const float4 t4 = lds_v4(stg + (uint32_t)(((i * 4 + lr) * 36 + lc) * 4));
        const bool msk = (maskmask >> i) & 1u;
        float x[4];
        x[0] = act_ct<ACT>(fmaf(alpha, t4.x, bz[0])); x[1] = act_ct<ACT>(fmaf(alpha, t4.y, bz[1]));
        x[2] = act_ct<ACT>(fmaf(alpha, t4.z, bz[2])); x[3] = act_ct<ACT>(fmaf(alpha, t4.w, bz[3]));
        if (msk) { x[0] = 0.f; x[1] = 0.f; x[2] = 0.f; x[3] = 0.f; }
        float* dst = crow[i] + gn;
        if (ATOMIC) {
            if (!msk) {
                if (vec) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
                else {
#pragma unroll
                    for (int e = 0; e < 4; e++) if (gn + e < p.N) atomicAdd(dst + e, x[e]);
                }
            }
        } else {
            if (vec) {
                if (p.accumulate == 1) { const float4 o = ldg_v4(dst); x[0] += o.x; x[1] += o.y; x[2] += o.z; x[3] += o.w; }
                stg_v4(dst, x[0], x[1], x[2], x[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (gn + e < p.N) { if (p.accumulate == 1) x[e] += dst[e]; dst[e] = x[e]; } else x[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) { cs[e] += x[e]; cq[e] = fmaf(x[e], x[e], cq[e]); }
        }
    }
}

// debug timeline (ns, %globaltimer) of CTA 0: 0 start, 1 setup done, 3 first tile landed, 4 last MMA issued, 5 accumulator ready,
// 6 epilogue done, 7 teardown.  Read with taco_debug_timeline().
__device__ unsigned long long g_tc_stamp[8];
#define TC_STAMP(i)                                                                                       \
    do {                                                                                                  \
        if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0) {                              \
            unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                  \
            g_tc_stamp[i] = _t;                                                                           \
        }                                                                                                 \
    } while (0)

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                         // [STAGES][16 KB]
    uint8_t* sB = smem + TC_STAGES * TC_TILE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 2 * TC_STAGES * TC_TILE_BYTES);
    uint64_t* empty_bar = full_bar + TC_STAGES;
    uint64_t* tmem_full = empty_bar + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tilesN = (p.N + TC_BN - 1) / TC_BN;
    const int tm = blockIdx.x / tilesN, tn = blockIdx.x % tilesN;
    const int m0 = tm * TC_BM, n0 = tn * TC_BN;
    const int ktiles = (p.K + TC_BK - 1) / TC_BK;
    const int kt_per = (ktiles + p.split_k - 1) / p.split_k;
    const int kt0 = blockIdx.y * kt_per, kt1 = min(ktiles, kt0 + kt_per);
    const int nkt = kt1 - kt0;
    if (nkt <= 0) return;      // uniform per CTA
    TC_STAMP(0);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB));
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    TC_STAMP(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        // lane 0 arms the barrier; the boxes of a stage (one per K-major operand, four per MN-major operand) are then
        // issued by different lanes so that a weight-gradient stage (8 boxes) does not serialise on one thread
        for (int it = 0; it < nkt; it++) {
            const int stage = it % TC_STAGES, round = it / TC_STAGES;
            if (lane == 0) {
                if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1);
                mbar_expect_tx(&full_bar[stage], 2 * TC_TILE_BYTES);
            }
            __syncwarp();
            int k0 = (kt0 + it) * TC_BK;
            if (p.tap_inner) {
                // convolution taps innermost: consecutive k-tiles read the same activation rows shifted by one frame, so the
                // re-read hits L2 (tap-major order re-reads a [128 x C] slab only after the whole wave streamed C channels)
                const int cb = (kt0 + it) / p.tap_inner, j = (kt0 + it) - cb * p.tap_inner;
                k0 = j * p.a_ctap + cb * TC_BK;
            }
            uint8_t* a = sA + stage * TC_TILE_BYTES;
            uint8_t* b = sB + stage * TC_TILE_BYTES;
            if (lane < 4) {
                if (!p.a_mn_major) {
                    if (lane == 0) {
                        if (p.tap_table) { const int2 tc = __ldg(p.tap_table + kt0 + it); tma_load_2d(a, &mapA, tc.x, m0 + tc.y, &full_bar[stage]); }
                        else if (p.a_tap) tma_load_2d(a, &mapA, k0 % p.a_ctap, m0 + k0 / p.a_ctap, &full_bar[stage]);
                        else tma_load_2d(a, &mapA, k0, m0, &full_bar[stage]);
                    }
                } else {
                    const int g = lane, mm = m0 + g * 32;
                    if (p.a_tap) tma_load_2d(a + g * 4096, &mapA, mm % p.a_ctap, k0 + mm / p.a_ctap, &full_bar[stage]);
                    else tma_load_2d(a + g * 4096, &mapA, mm, k0, &full_bar[stage]);
                }
            } else if (lane < 8) {
                if (!p.b_mn_major) {
                    if (lane == 4) tma_load_2d(b, &mapB, k0, n0, &full_bar[stage]);
                } else {
                    const int g = lane - 4;
                    tma_load_2d(b + g * 4096, &mapB, n0 + g * 32, k0, &full_bar[stage]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b=TF32 [7,10)/[10,13), majors [15],[16], N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) | ((uint32_t)p.b_mn_major << 16) |
                               ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        for (int it = 0; it < nkt; it++) {
            const int stage = it % TC_STAGES, round = it / TC_STAGES;
            mbar_wait(&full_bar[stage], round & 1);
            if (it == 0) TC_STAMP(3);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a = smem_u32(sA + stage * TC_TILE_BYTES), b = smem_u32(sB + stage * TC_TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; kk++) {
                    // K-major: 8 tf32 = 32 B along the swizzled 128 B row, 8-row groups 1024 B apart (SBO).
                    // MN-major: one 8-row k-group (1024 B) per MMA; 32-element MN groups 4096 B apart (LBO); SBO = 1024 B.
                    //           (32-bit MN-major uses 4-row swizzle atoms: two 512 B k-groups per MMA, SBO = 512 B)
                    const uint64_t ad = p.a_mn_major ? umma_desc(a + kk * 1024, 4096, 512, 1) : umma_desc(a + kk * 32, 16, 1024, 2);
                    const uint64_t bd = p.b_mn_major ? umma_desc(b + kk * 1024, 4096, 512, 1) : umma_desc(b + kk * 32, 16, 1024, 2);
                    umma_tf32(tmem_base, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);                  // frees the smem slot when these MMAs retire
                if (it == nkt - 1) umma_commit(tmem_full);       // accumulator complete
                if (it == nkt - 1) TC_STAMP(4);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // TMEM -> registers (row per lane) -> 32x32 transpose through shared memory -> coalesced 128-byte row segments.
        // Warp w may touch TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter split the four 32-column chunks.
        const int q = warp & 3, half = (warp - 2) >> 2;
        mbar_wait(tmem_full, 0);
        if (warp == 4) TC_STAMP(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stg = smem_u32(reinterpret_cast<float*>(sA) + (warp - 2) * (32 * 36));   // pipeline smem is idle once the accumulator is complete
        const uint32_t red_s = smem_u32(sB);
        const int lr = lane >> 3, lc = (lane & 7) * 4;
        const bool atomic = (p.split_k > 1) || (p.accumulate == 2);
        float* crow[8];
        uint32_t okmask = 0, maskmask = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int gm = m0 + q * 32 + i * 4 + lr;
            crow[i] = p.C;
            if (gm < p.M) {
                okmask |= 1u << i;
                if (p.mask_period > 0) { const int t = gm % p.mask_period; if (t < p.mask_lo || t >= p.mask_hi) maskmask |= 1u << i; }
                crow[i] = p.remap_period > 0
                    ? p.C + (long long)(gm / p.remap_period) * p.remap_outer + (long long)(gm % p.remap_period) * p.remap_inner
                    : p.C + (long long)gm * p.ldc;
            }
        }
        for (int cc = 0; cc < 2; cc++) {
            const int c0 = (half * 2 + cc) * 32;
            if (n0 + c0 >= p.N) break;                           // warp-uniform
            {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4) sts_v4(stg + (uint32_t)((lane * 36 + j) * 4), v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            __syncwarp();
            const int gn = n0 + c0 + lc;
            float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
            if (atomic) epi_chunk<ACT_NONE, true>(p, stg, crow, okmask, maskmask, lr, lc, gn, blockIdx.y == 0, cs, cq);
            else switch (p.act) {                                // kernel-uniform
                case ACT_RELU:     epi_chunk<ACT_RELU, false>(p, stg, crow, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                case ACT_SIGMOID:  epi_chunk<ACT_SIGMOID, false>(p, stg, crow, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                case ACT_TANH:     epi_chunk<ACT_TANH, false>(p, stg, crow, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                case ACT_SOFTSIGN: epi_chunk<ACT_SOFTSIGN, false>(p, stg, crow, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                default:           epi_chunk<ACT_NONE, false>(p, stg, crow, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
            }
            __syncwarp();                                        // staging buffer is rewritten by the next chunk
            if (p.colsum) {                                      // kernel-uniform
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);  cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 8);
                    cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 16);
                }
                if (lane < 8) {                                  // per-quarter partials, reduced over the CTA below
                    const uint32_t r = red_s + (uint32_t)((q * (2 * TC_BN) + c0 + lc) * 4);
                    sts_v4(r, cs[0], cs[1], cs[2], cs[3]);
                    sts_v4(r + TC_BN * 4, cq[0], cq[1], cq[2], cq[3]);
                }
            }
        }
        if (p.colsum) {
            // one double atomic per column and statistic per CTA (the four lane quarters are summed here first: the
            // statistics land on a few hundred addresses, so their atomics serialise in L2)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int et = threadIdx.x - 64, col = et & (TC_BN - 1), which = et >> 7;
            if (n0 + col < p.N) {
                const uint32_t r = red_s + (uint32_t)((which * TC_BN + col) * 4);
                const float v = (lds_f32(r) + lds_f32(r + 2 * TC_BN * 4)) + (lds_f32(r + 4 * TC_BN * 4) + lds_f32(r + 6 * TC_BN * 4));
                atomicAdd((which ? p.colsumsq : p.colsum) + n0 + col, (double)v);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (warp == 4) TC_STAMP(6);
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_BN));
        TC_STAMP(7);
    }
}

// ---------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int get_encode() {
    static std::once_flag once;
    static int rc = TACO_OK;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { rc = TACO_ECUDA; return; }
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    if (rc != TACO_OK) set_error("cuTensorMapEncodeTiled entry point unavailable");
    return rc;
}

// fp32 2-D tensor map: dim0 contiguous (extent d0), dim1 rows (extent d1, stride ld elements), box {b0, b1}, 128B swizzle
static int make_map(CUtensorMap* map, const float* base, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1, bool mn_major, bool soft = false) {
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && soft) return TACO_ENOTSUP;
    TACO_REQUIRE(r == CUDA_SUCCESS, TACO_ECUDA, "cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu) ld=%llu box=(%u,%u)", (int)r,
                 (const void*)base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    return TACO_OK;
}

// Can this problem run on the tensor-core path?  (TMA needs 16-byte aligned bases and row pitches; taps must align to tiles.)
bool gemm_tc_eligible(const taco_gemm_desc& g) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (!al16(g.A) || !al16(g.B)) return false;
    if (g.lda % 4 != 0 || g.ldb % 4 != 0) return false;
    const bool tap = g.ctap > 0 && (g.lda != g.ctap || g.ctap % 32 == 0);
    if (tap && g.ctap % 32 != 0) return false;
    if (g.tap_table && (g.transA || g.K % TC_BK != 0 || g.ctap != 0 || (reinterpret_cast<uintptr_t>(g.tap_table) & 7u) != 0)) return false;
    if ((long long)g.M * g.N * g.K < (1ll << 24)) return false;      // small problems: launch-bound either way; the grouped fp32 kernel batches them
    return true;
}

int launch_gemm_tc(const taco_gemm_desc& g, cudaStream_t s) {
    TACO_TRY(get_encode());
    int dev = 0; TACO_CHECK_CUDA(cudaGetDevice(&dev));
    TACO_REQUIRE(dev >= 0 && dev < 16, TACO_ECUDA, "gemm: device ordinal %d out of range", dev);
    static bool configured[16] = {};       // function attributes and SM counts are per device
    static int n_sm_dev[16] = {};
    if (!configured[dev]) {
        TACO_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        TACO_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev));
        configured[dev] = true;
    }
    TACO_REQUIRE(!((g.split_k > 1 || g.accumulate == 2) && (g.act != 0 || g.colsum != nullptr)), TACO_EINVAL,
                 "gemm: atomic accumulation cannot carry an activation or column statistics");
    const float* A = static_cast<const float*>(g.A);
    const float* B = static_cast<const float*>(g.B);
    // lda == ctap with ctap % 32 != 0 (80-channel mel input): im2col rows are contiguous, so the map uses overlapping
    // rows (pitch lda < extent K); if the driver rejects that the caller falls back to the SIMT kernel.
    const bool tap = g.ctap > 0 && (g.lda != g.ctap || g.ctap % 32 == 0);
    const bool overlap = g.ctap > 0 && !tap;
    const int ntaps = g.ctap > 0 ? ((g.transA ? g.M : g.K) + g.ctap - 1) / g.ctap : 1;
    CUtensorMap mapA, mapB;
    if (g.tap_table) {
        // table-driven taps: the map spans whole rows of the activation matrix; the producer supplies (column, row) per k-tile
        TACO_TRY(make_map(&mapA, A, (uint64_t)g.lda, (uint64_t)g.tap_rows, (uint64_t)g.lda, TC_BK, TC_BM, false));
    } else if (!g.transA) {
        // K-major A: box {32 k, 128 m}
        if (tap) TACO_TRY(make_map(&mapA, A, (uint64_t)g.ctap, (uint64_t)g.M + ntaps - 1, (uint64_t)g.lda, TC_BK, TC_BM, false));
        else TACO_TRY(make_map(&mapA, A, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda, TC_BK, TC_BM, false, overlap));
    } else {
        // A^T: stored [rows = K, cols = M (taps)]: MN-major, boxes {32 m, 32 k}
        if (tap) TACO_TRY(make_map(&mapA, A, (uint64_t)g.ctap, (uint64_t)g.K + ntaps - 1, (uint64_t)g.lda, 32, TC_BK, true));
        else TACO_TRY(make_map(&mapA, A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda, 32, TC_BK, true, overlap));
    }
    if (!g.transB) TACO_TRY(make_map(&mapB, B, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldb, 32, TC_BK, true));     // MN-major
    else TACO_TRY(make_map(&mapB, B, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldb, TC_BK, TC_BN, false));            // K-major
    TcParams p{};
    p.C = g.C; p.M = g.M; p.N = g.N; p.K = g.K; p.ldc = g.ldc;
    p.a_mn_major = g.transA ? 1 : 0; p.b_mn_major = g.transB ? 0 : 1;
    p.a_tap = tap ? 1 : 0; p.a_ctap = tap ? g.ctap : 1;
    p.tap_table = reinterpret_cast<const int2*>(g.tap_table);
    // Measured on B200 (post-net proj_1, M=25824 N=256 K=3x2048): tap-innermost order cuts the DRAM re-reads of the taps but
    // does not speed the GEMM up (0.300 vs 0.298 ms; narrower convolutions got slower) - the main loop is bound by the per-SM
    // operand ingest of fp32 tiles, not by DRAM.  Kept behind TACO_TAP_INNER=1 for re-measurement.
    static const bool tap_inner_on = [] { const char* e = getenv("TACO_TAP_INNER"); return e && e[0] == '1'; }();
    p.tap_inner = (tap_inner_on && tap && !g.transA && g.ctap % TC_BK == 0 && g.K % g.ctap == 0 && g.K / g.ctap > 1) ? g.K / g.ctap : 0;
    p.alpha = g.alpha; p.accumulate = g.accumulate; p.bias = g.bias; p.act = g.act;
    p.mask_period = g.mask_period; p.mask_lo = g.mask_lo; p.mask_hi = g.mask_hi;
    p.remap_period = g.remap_period; p.remap_outer = g.remap_outer; p.remap_inner = g.remap_inner;
    p.colsum = g.colsum; p.colsumsq = g.colsumsq;
    const int ktiles = cdiv(g.K, TC_BK);
    const int tiles = cdiv(g.M, TC_BM) * cdiv(g.N, TC_BN);
    const int n_sm = n_sm_dev[dev];
    // A caller that allows K splitting (split_k > 1: C is accumulated atomically) gets the split that balances whole waves
    // of CTAs (two resident per SM): waves x (k-blocks per CTA + a fixed per-CTA cost in k-block units).
    p.split_k = 1;
    if (g.split_k > 1) {
        double best = 1e30;
        for (int sp = 1; sp <= 64 && sp <= ktiles; sp++) {
            const int per = cdiv(ktiles, sp);
            if (sp > 1 && per < 8) break;
            const double cost = (double)cdiv(tiles * sp, 2 * n_sm) * (per + 10.0);
            if (cost < best - 1e-9) { best = cost; p.split_k = sp; }
        }
    }
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.vecC = g.remap_period > 0 ? (al16(g.C) && g.remap_outer % 4 == 0 && g.remap_inner % 4 == 0) : (al16(g.C) && g.ldc % 4 == 0);
    dim3 grid(cdiv(g.M, TC_BM) * cdiv(g.N, TC_BN), p.split_k);
    gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(mapA, mapB, p);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

int debug_timeline(unsigned long long out[8]) {
    TACO_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_tc_stamp, sizeof(unsigned long long) * 8));
    return TACO_OK;
}

}  // namespace taco

extern "C" int taco_debug_timeline(unsigned long long out[8]) { return taco::debug_timeline(out); }
