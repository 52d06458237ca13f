// bf16-operand tensor-core GEMM for sm_100a (TACO_PREC_BF16): TMA -> 128B-swizzled shared memory -> tcgen05.mma kind::f16
// (bf16 x bf16, fp32 accumulate in TMEM) -> tcgen05.ld epilogue.  Serves the same descriptor (taco_gemm_desc) as the fp32
// kernels through its bf16 operand mirrors (A16 / B16) and can write a bf16 mirror of the result beside (or instead of) the
// fp32 one, so that GEMM -> GEMM chains never touch fp32 activations.        reference ops covered: see gemm_simt.cu header.
//
// Why this shape (measured in round 1, profiles/r1_ncu_full_gemm_proj1.summary.txt): the fp32-operand kernel is bound by
// the chip-wide L2 -> SM throughput (~6300 B/clk, B300_MICROARCH "LTS throughput cap"; 46 B/clk/SM were measured), not by
// DRAM or the tensor pipe.  The lever is bytes per FLOP:   fp32 128x128x32 tile: 32 KB per 1.05 MFLOP
//                                                           bf16 128x256x64 tile: 48 KB per 4.19 MFLOP  (2.7x fewer)
// Structure: persistent CTAs (one per SM, 192 threads): warp 0 = TMA producer + dynamic tile scheduler (atomic work counter,
// so CTAs that start late - SMs still held by a recurrence kernel of another stream - simply find no work left),
// warp 1 = MMA issuer + TMEM allocator, warps 2-5 = epilogue (one TMEM lane quarter each).  Two accumulators of up to 256
// columns in TMEM: the epilogue of tile i overlaps the main loop of tile i+1.  4-stage mbarrier ring of {A 128x64, B BNx64}.
//   A(m,k) row-major [M,K]      -> K-major  SW128 tile  (one TMA box  {64 k, 128 m})
//   A^T    stored [K(rows), M]  -> MN-major SW128 tile  (two TMA boxes {64 m, 64 k})       (weight gradients)
//   B[k*ldb+n] (TF [in,out])    -> MN-major SW128 tile  (BN/64 TMA boxes {64 n, 64 k})
//   B[n*ldb+k]                  -> K-major  SW128 tile  (one TMA box  {64 k, BN n})         (data gradients)
// Convolution taps are TMA coordinates over the zero-padded activation (walked tap-innermost so the re-read of a row slab
// hits L2), the conv-bank data gradient walks a per-k-tile tap table, out-of-range K / N / M is TMA zero fill.
#include "tc_common.cuh"
#include "kernels.h"
#include <cudaTypedefs.h>
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <cstdlib>
#include <cstring>

namespace taco {

constexpr int BF_BM = 128, BF_BK = 64, BF_BN_MAX = 256, BF_MAX_STAGES = 6, BF_PROD_WARPS = 3, BF_EPI_WARPS = 8;
constexpr int BF_EPI_WARP0 = BF_PROD_WARPS + 1, BF_THREADS = 32 * (BF_PROD_WARPS + 1 + BF_EPI_WARPS);     // warps 0-2 TMA, warp 3 MMA, warps 4-11 epilogue
constexpr int BF_A_BYTES = BF_BM * BF_BK * 2;                    // 16 KB
constexpr int BF_STG_FLOATS = 32 * 36;                           // per epilogue warp: 32x32 transpose buffer, pitch 36
constexpr int BF_FIXED_SMEM = BF_EPI_WARPS * BF_STG_FLOATS * 4 + 4 * 2 * BF_BN_MAX * 4 + 512;   // staging + statistics partials + barriers
constexpr int BF_SMEM = 227 * 1024;                              // all of it: the stage ring takes what the fixed part leaves
constexpr int BF_RING_BYTES = BF_SMEM - 1024 - BF_FIXED_SMEM;    // 1024: alignment slack
constexpr int BF_SCHED_SLOTS = 256;

struct BfParams {
    float* C; __nv_bfloat16* C16;
    int M, N, K, ldc, ldc16;
    int BN, tilesN, units, split_k, kt_per, ktiles;
    int stages, b_stage_bytes;           // ring depth and B stage pitch: ceil(BN/64) x 8 KB
    int one_shot;                        // 1: grid = units, every CTA computes the unit of its index and exits (leaf GEMMs, see launch_gemm_bf16)
    int fast_ok;                         // the lean epilogue applies to interior chunks (vector stores, aligned bias, no read-modify-write)
    int a_mn_major, b_mn_major;
    int b_3d;                            // MN-major B tile through ONE 3-D box {64 n, 64 k, BN/64 n-blocks}
    int a_tap, a_ctap, tap_inner, tap_group;   // tap_inner: number of taps (0: tap-major walk); tap_group: channel blocks per group
    const int2* tap_table;
    float alpha; int accumulate;
    const float* bias; int act;
    int mask_period, mask_lo, mask_hi;
    int remap_period; long long remap_outer, remap_inner;
    double* colsum; double* colsumsq;
    int vecC, vecC16;
    unsigned int* sched;                 // [0] next work unit, [1] CTAs done (the last one resets both)
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// One 32x32 chunk of the output tile: rows {4i+lr}, columns gn..gn+3 per lane, read back from the staging buffer.
template <int ACT, bool ATOMIC>
__device__ __forceinline__ void bf_epi_chunk(const BfParams& p, uint32_t stg, const long long rowoff[8], const long long rowoff16[8], uint32_t okmask, uint32_t maskmask,
                                             int lr, int lc, int gn, bool add_bias, float cs[4], float cq[4]) {
    const bool full4 = gn + 3 < p.N;
    const bool vec = p.vecC && full4;
    const bool vec16 = p.vecC16 && full4;
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
        if (ATOMIC) {
            float* dst = p.C + rowoff[i] + gn;
            if (!msk) {
                if (vec) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
                else {
#pragma unroll
                    for (int e = 0; e < 4; e++) if (gn + e < p.N) atomicAdd(dst + e, x[e]);
                }
            }
        } else {
            if (p.C) {
                float* dst = p.C + rowoff[i] + gn;
                if (vec) {
                    if (p.accumulate == 1) { const float4 o = ldg_v4(dst); x[0] += o.x; x[1] += o.y; x[2] += o.z; x[3] += o.w; }
                    stg_v4(dst, x[0], x[1], x[2], x[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (gn + e < p.N) { if (p.accumulate == 1) x[e] += dst[e]; dst[e] = x[e]; } else x[e] = 0.f;
                }
            } else if (!full4) {
#pragma unroll
                for (int e = 0; e < 4; e++) if (gn + e >= p.N) x[e] = 0.f;
            }
            if (p.C16) {
                __nv_bfloat16* d16 = p.C16 + rowoff16[i] + gn;
                if (vec16) stg_v2_b32(d16, pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]));
                else {
#pragma unroll
                    for (int e = 0; e < 4; e++) if (gn + e < p.N) d16[e] = __float2bfloat16_rn(x[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; e++) { cs[e] += x[e]; cq[e] = fmaf(x[e], x[e], cq[e]); }
        }
    }
}

// Lean epilogue of an interior 32x32 chunk (all rows valid, all columns inside N, 128-bit stores): ~25 instructions per row.
// The epilogue is instruction-bound - each epilogue warp runs alone or in pairs on its scheduler, so every dependent
// instruction costs its full latency (measured: the general path below took ~3 000 clk per chunk).
template <int ACT, bool STATS, bool ACC = false>
__device__ __forceinline__ void bf_epi_fast(const BfParams& p, uint32_t stg_lane, float* const crow[8], __nv_bfloat16* const crow16[8], uint32_t maskmask,
                                            int gn, float4 bz, float cs[4], float cq[4]) {
    const float alpha = p.alpha;
    const bool hc = p.C != nullptr, hc16 = p.C16 != nullptr;       // kernel-uniform
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float4 t4 = lds_v4(stg_lane + (uint32_t)(i * 4 * 36 * 4));
        float x0 = act_ct<ACT>(fmaf(alpha, t4.x, bz.x)), x1 = act_ct<ACT>(fmaf(alpha, t4.y, bz.y));
        float x2 = act_ct<ACT>(fmaf(alpha, t4.z, bz.z)), x3 = act_ct<ACT>(fmaf(alpha, t4.w, bz.w));
        if ((maskmask >> i) & 1u) { x0 = 0.f; x1 = 0.f; x2 = 0.f; x3 = 0.f; }
        if (ACC) { const float4 o = ldg_v4(crow[i] + gn); x0 += o.x; x1 += o.y; x2 += o.z; x3 += o.w; }
        if (hc) stg_v4(crow[i] + gn, x0, x1, x2, x3);
        if (hc16) stg_v2_b32(crow16[i] + gn, pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));
        if (STATS) {
            cs[0] += x0; cs[1] += x1; cs[2] += x2; cs[3] += x3;
            cq[0] = fmaf(x0, x0, cq[0]); cq[1] = fmaf(x1, x1, cq[1]); cq[2] = fmaf(x2, x2, cq[2]); cq[3] = fmaf(x3, x3, cq[3]);
        }
    }
}
// atomic variant (split-K / accumulate == 2): fp32 vector reductions, masked rows contribute nothing
__device__ __forceinline__ void bf_epi_fast_atomic(const BfParams& p, uint32_t stg_lane, float* const crow[8], uint32_t maskmask, int gn, float4 bz) {
    const float alpha = p.alpha;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float4 t4 = lds_v4(stg_lane + (uint32_t)(i * 4 * 36 * 4));
        if ((maskmask >> i) & 1u) continue;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow[i] + gn), "f"(fmaf(alpha, t4.x, bz.x)), "f"(fmaf(alpha, t4.y, bz.y)),
                     "f"(fmaf(alpha, t4.z, bz.z)), "f"(fmaf(alpha, t4.w, bz.w)) : "memory");
    }
}

// bytes producer warp w lands per k-tile (see the producer section of the kernel for the box assignment)
__device__ __forceinline__ uint32_t prod_bytes(const BfParams& p, int w) {
    const uint32_t nboxB = p.b_mn_major ? (uint32_t)((p.BN + 63) >> 6) : 0u;
    const bool b2d = p.b_mn_major && !p.b_3d;
    if (w == 0) return p.a_mn_major ? 8192u : (uint32_t)BF_A_BYTES;
    if (w == 1) return (p.a_mn_major ? 8192u : 0u) + (b2d ? (nboxB / 2) * 8192u : 0u);
    return !p.b_mn_major ? (uint32_t)p.BN * 128u : (p.b_3d ? nboxB * 8192u : ((nboxB + 1) / 2) * 8192u);
}

// debug timeline (ns, %globaltimer) of CTA 0: 0 start, 1 setup done, then per local tile lt < 15: 2+4lt first operands landed,
// 3+4lt last MMA issued, 4+4lt accumulator ready (epilogue warp 2), 5+4lt epilogue done.  Read with taco_debug_timeline_bf16().
__device__ unsigned long long g_bf_stamp[64];
#define BF_STAMP(i)                                                                                       \
    do {                                                                                                  \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (i) < 64) {                                     \
            unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                  \
            g_bf_stamp[i] = _t;                                                                           \
        }                                                                                                 \
    } while (0)

__global__ void __launch_bounds__(BF_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const BfParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SW128 tiles need 1024-byte alignment
    const int STAGES = p.stages;
    uint8_t* sA = smem;                                          // [STAGES][16 KB]
    uint8_t* sB = smem + STAGES * BF_A_BYTES;                    // [STAGES][b_stage_bytes]
    float* stg_all = reinterpret_cast<float*>(smem + BF_RING_BYTES);
    float* red = stg_all + BF_EPI_WARPS * BF_STG_FLOATS;         // [4 quarters][2 stats][256]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 4 * 2 * BF_BN_MAX);      // [producer warp][stage]: one barrier per issuing warp
    uint64_t* empty_bar = full_bar + BF_PROD_WARPS * BF_MAX_STAGES;
    uint64_t* tmem_full = empty_bar + BF_MAX_STAGES;             // [2]
    uint64_t* tmem_empty = tmem_full + 2;                        // [2]
    uint64_t* sched_full = tmem_empty + 2;                       // [2]
    uint64_t* sched_empty = sched_full + 2;                      // [2]
    volatile int* sched_unit = reinterpret_cast<volatile int*>(sched_empty + 2);   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(sched_unit) + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BN = p.BN;
    if (warp == 0) BF_STAMP(0);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB));
        for (int i = 0; i < STAGES; i++) { for (int w = 0; w < BF_PROD_WARPS; w++) mbar_init(&full_bar[w * BF_MAX_STAGES + i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], BF_EPI_WARPS); mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], (BF_PROD_WARPS - 1) + 1 + BF_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == BF_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BF_BN_MAX));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) BF_STAMP(1);

    // work unit u -> (m tile, n tile, k split); n fastest so that concurrently running CTAs share A row slabs in L2
    auto decode = [&](int u, int& m0, int& n0, int& kt0, int& nkt) {
        const int sp = u % p.split_k, t = u / p.split_k;
        m0 = (t / p.tilesN) * BF_BM; n0 = (t % p.tilesN) * BN;
        kt0 = sp * p.kt_per; nkt = min(p.ktiles, kt0 + p.kt_per) - kt0;
    };

    if (warp < BF_PROD_WARPS) {
        // ===================== scheduler (warp 0) + TMA producers (warps 0-2) =====================
        // Measured on B200 (tools/tma_bench*.cu): ONE issuing thread gets a cp.async.bulk.tensor instruction through every
        // ~770 clk whatever the box holds (8 KB or 64 KB), different warps overlap fully (4 warps x 16 KB boxes: 81 B/clk/SM).
        // A k-tile's boxes are therefore spread over three warps - A (or A's first half) on warp 0, A's second half on
        // warp 1, B on warp 2 as one 2-D (K-major) or 3-D (MN-major: all 64-column blocks at once) box; a B tile that
        // needs several 2-D boxes alternates them between warps 2 and 1.
        // Each producer warp completes its boxes on its OWN full barrier (boxes of different warps on one barrier were measured
        // to serialise again); the MMA warp waits for the barriers of all warps that carry bytes in this configuration.
        const uint32_t nboxB = p.b_mn_major ? (uint32_t)((BN + 63) >> 6) : 0u;
        const uint32_t my_bytes = prod_bytes(p, warp);
        uint64_t* my_full = full_bar + warp * BF_MAX_STAGES;
        int it = 0;
        int next = 0;
        if (warp == 0 && lane == 0) next = p.one_shot ? (int)blockIdx.x : (int)atomicAdd(p.sched, 1u);
        for (int lt = 0;; lt++) {
            const int slot = lt & 1, use = lt >> 1;
            int u = 0;
            if (warp == 0) {
                if (lane == 0) {
                    u = next;
                    if (use > 0) mbar_wait_bounded(&sched_empty[slot], (use - 1) & 1);
                    sched_unit[slot] = u;
                    mbar_arrive(&sched_full[slot]);
                    if (u < p.units) next = p.one_shot ? p.units : (int)atomicAdd(p.sched, 1u);  // fetched one tile ahead: its latency hides behind this tile's loads
                }
                u = __shfl_sync(0xffffffffu, u, 0);
            } else {
                mbar_wait_bounded(&sched_full[slot], use & 1);
                u = sched_unit[slot];
                __syncwarp();
                if (lane == 0) mbar_arrive(&sched_empty[slot]);
            }
            if (u >= p.units) break;
            int m0, n0, kt0, nkt;
            decode(u, m0, n0, kt0, nkt);
            if (lane == 0) {
                for (int j = 0; j < nkt; j++, it++) {
                    const int stage = it % STAGES, round = it / STAGES;
                    if (round > 0) mbar_wait_bounded(&empty_bar[stage], (round - 1) & 1);
                    if (my_bytes) mbar_expect_tx(&my_full[stage], my_bytes);
                    const int kt = kt0 + j;
                    int k0 = kt * BF_BK;
                    if (p.tap_inner) {
                        // optional K walk of a convolution: (group of G channel blocks, tap, block in group) - re-reads hit L2, but the
                        // plain tap-major walk measured faster (post proj_1: 160 vs 187 us) and is the default (TACO_BF16_TAPG)
                        const int per = p.tap_inner * p.tap_group;
                        const int gq = kt / per, rem = kt - gq * per, tj = rem / p.tap_group, cb = gq * p.tap_group + (rem - tj * p.tap_group);
                        k0 = tj * p.a_ctap + cb * BF_BK;
                    }
                    uint8_t* a = sA + stage * BF_A_BYTES;
                    uint8_t* b = sB + stage * p.b_stage_bytes;
                    if (warp == 0) {
                        if (!p.a_mn_major) {
                            if (p.tap_table) { const int2 tc = __ldg(p.tap_table + kt); tma_load_2d(a, &mapA, tc.x, m0 + tc.y, &my_full[stage]); }
                            else if (p.a_tap) tma_load_2d(a, &mapA, k0 % p.a_ctap, m0 + k0 / p.a_ctap, &my_full[stage]);
                            else tma_load_2d(a, &mapA, k0, m0, &my_full[stage]);
                        } else {
                            if (p.a_tap) tma_load_2d(a, &mapA, m0 % p.a_ctap, k0 + m0 / p.a_ctap, &my_full[stage]);
                            else tma_load_2d(a, &mapA, m0, k0, &my_full[stage]);
                        }
                    } else if (warp == 1) {
                        if (p.a_mn_major) {
                            const int mm = m0 + 64;
                            if (p.a_tap) tma_load_2d(a + 8192, &mapA, mm % p.a_ctap, k0 + mm / p.a_ctap, &my_full[stage]);
                            else tma_load_2d(a + 8192, &mapA, mm, k0, &my_full[stage]);
                        }
                        if (p.b_mn_major && !p.b_3d)
                            for (uint32_t g = 1; g < nboxB; g += 2) tma_load_2d(b + g * 8192, &mapB, n0 + (int)g * 64, k0, &my_full[stage]);
                    } else {
                        if (!p.b_mn_major) tma_load_2d(b, &mapB, k0, n0, &my_full[stage]);
                        else if (p.b_3d) tma_load_3d(b, &mapB, 0, k0, n0 >> 6, &my_full[stage]);
                        else for (uint32_t g = 0; g < nboxB; g += 2) tma_load_2d(b + g * 8192, &mapB, n0 + (int)g * 64, k0, &my_full[stage]);
                    }
                }
            }
            __syncwarp();
        }
        if (warp == 0 && lane == 0 && !p.one_shot) {
            // self-cleaning work counter: the last CTA to leave resets the slot for a later launch
            __threadfence();
            const unsigned int done = atomicAdd(p.sched + 1, 1u);
            if (done == gridDim.x - 1) { p.sched[0] = 0u; p.sched[1] = 0u; __threadfence(); }
        }
    } else if (warp == BF_PROD_WARPS) {
        // ===================== MMA issuer =====================
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b=BF16 [7,10)/[10,13), majors [15],[16], N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn_major << 15) | ((uint32_t)p.b_mn_major << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BF_BM >> 4) << 24);
        int it = 0;
        const uint32_t pbytes[BF_PROD_WARPS] = {prod_bytes(p, 0), prod_bytes(p, 1), prod_bytes(p, 2)};
        for (int lt = 0;; lt++) {
            const int slot = lt & 1, use = lt >> 1;
            mbar_wait_bounded(&sched_full[slot], use & 1);
            const int u = sched_unit[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (u >= p.units) break;
            int m0, n0, kt0, nkt;
            decode(u, m0, n0, kt0, nkt);
            if (use > 0) mbar_wait_bounded(&tmem_empty[slot], (use - 1) & 1);      // the epilogue drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)(slot * BF_BN_MAX);
            for (int j = 0; j < nkt; j++, it++) {
                const int stage = it % STAGES, round = it / STAGES;
#pragma unroll
                for (int w = 0; w < BF_PROD_WARPS; w++)
                    if (pbytes[w]) mbar_wait_bounded(&full_bar[w * BF_MAX_STAGES + stage], round & 1);
                if (j == 0) BF_STAMP(2 + 4 * lt);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t a = smem_u32(sA + stage * BF_A_BYTES), b = smem_u32(sB + stage * p.b_stage_bytes);
#pragma unroll
                    for (int kk = 0; kk < BF_BK / 16; kk++) {
                        // K-major: 16 bf16 = 32 B along the swizzled 128 B row, 8-row groups 1024 B apart (SBO).
                        // MN-major: 16 k-rows of 128 B (two 8-row groups, SBO = 1024 B) per MMA; 64-element MN groups are
                        //           separate TMA boxes 8192 B apart (LBO).
                        const uint64_t ad = p.a_mn_major ? umma_desc(a + kk * 2048, 8192, 1024, 2) : umma_desc(a + kk * 32, 16, 1024, 2);
                        const uint64_t bd = p.b_mn_major ? umma_desc(b + kk * 2048, 8192, 1024, 2) : umma_desc(b + kk * 32, 16, 1024, 2);
                        umma_f16(tacc, ad, bd, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);                  // frees the smem slot when these MMAs retire
                    if (j == nkt - 1) umma_commit(&tmem_full[slot]); // accumulator complete
                    if (j == nkt - 1) BF_STAMP(3 + 4 * lt);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 4..11) =====================
        // TMEM -> registers (row per lane) -> 32x32 transpose through shared memory -> coalesced 128-byte row segments.
        // Warp w may touch TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter take alternate 32-column chunks.
        const int q = warp & 3, half = (warp - BF_EPI_WARP0) >> 2;
        const uint32_t stg = smem_u32(stg_all + (warp - BF_EPI_WARP0) * BF_STG_FLOATS);
        const uint32_t red_s = smem_u32(red);
        const int lr = lane >> 3, lc = (lane & 7) * 4;
        const uint32_t stg_lane = stg + (uint32_t)((lr * 36 + lc) * 4);
        const bool atomic = (p.split_k > 1) || (p.accumulate == 2);
        for (int lt = 0;; lt++) {
            const int slot = lt & 1, use = lt >> 1;
            mbar_wait_bounded(&sched_full[slot], use & 1);
            const int u = sched_unit[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (u >= p.units) break;
            int m0, n0, kt0, nkt;
            decode(u, m0, n0, kt0, nkt);
            // the bf16 mirror keeps the fp32 buffer's element layout (same remap); only its row pitch may differ
            long long rowoff[8], rowoff16[8];
            float* crow[8]; __nv_bfloat16* crow16[8];
            uint32_t okmask = 0, maskmask = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int gm = m0 + q * 32 + i * 4 + lr;
                rowoff[i] = 0; rowoff16[i] = 0;
                if (gm < p.M) {
                    okmask |= 1u << i;
                    if (p.mask_period > 0) { const int t = gm % p.mask_period; if (t < p.mask_lo || t >= p.mask_hi) maskmask |= 1u << i; }
                    rowoff[i] = p.remap_period > 0
                        ? (long long)(gm / p.remap_period) * p.remap_outer + (long long)(gm % p.remap_period) * p.remap_inner
                        : (long long)gm * p.ldc;
                    rowoff16[i] = p.remap_period > 0 ? rowoff[i] : (long long)gm * p.ldc16;
                }
                crow[i] = p.C + rowoff[i]; crow16[i] = p.C16 + rowoff16[i];
            }
            const bool rows_full = (m0 + q * 32 + 31 < p.M) && p.fast_ok;     // warp-uniform
            const bool with_bias = p.bias != nullptr && (!atomic || kt0 == 0);
            mbar_wait_bounded(&tmem_full[slot], use & 1);
            if (warp == BF_EPI_WARP0) BF_STAMP(4 + 4 * lt);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)(slot * BF_BN_MAX) + ((uint32_t)(q * 32) << 16);
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                if (n0 + c0 >= p.N) break;                           // warp-uniform
                const int gn = n0 + c0 + lc;
                const bool fast = rows_full && (n0 + c0 + 32 <= p.N);
                float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fast && with_bias) bz = __ldg(reinterpret_cast<const float4*>(p.bias + gn));   // in flight across the TMEM load
                {
                    float v[32];
                    tmem_ld32(tacc + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) sts_v4(stg + (uint32_t)((lane * 36 + j) * 4), v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                __syncwarp();
                float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
                if (fast) {
                    if (atomic) bf_epi_fast_atomic(p, stg_lane, crow, maskmask, gn, bz);
                    else if (p.accumulate == 1) bf_epi_fast<ACT_NONE, false, true>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq);   // C += A.B (data gradients)
                    else if (p.colsum) {
                        if (p.act == ACT_RELU) bf_epi_fast<ACT_RELU, true>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq);
                        else if (p.act == ACT_NONE) bf_epi_fast<ACT_NONE, true>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq);
                        else bf_epi_chunk<ACT_NONE, false>(p, stg, rowoff, rowoff16, 0u, maskmask, lr, lc, gn, true, cs, cq);   // rejected on the host
                    } else switch (p.act) {                          // kernel-uniform
                        case ACT_RELU:     bf_epi_fast<ACT_RELU, false>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq); break;
                        case ACT_SIGMOID:  bf_epi_fast<ACT_SIGMOID, false>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq); break;
                        case ACT_TANH:     bf_epi_fast<ACT_TANH, false>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq); break;
                        case ACT_SOFTSIGN: bf_epi_fast<ACT_SOFTSIGN, false>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq); break;
                        default:           bf_epi_fast<ACT_NONE, false>(p, stg_lane, crow, crow16, maskmask, gn, bz, cs, cq); break;
                    }
                } else if (atomic) bf_epi_chunk<ACT_NONE, true>(p, stg, rowoff, rowoff16, okmask, maskmask, lr, lc, gn, kt0 == 0, cs, cq);
                else switch (p.act) {                                // kernel-uniform
                    case ACT_RELU:     bf_epi_chunk<ACT_RELU, false>(p, stg, rowoff, rowoff16, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                    case ACT_SIGMOID:  bf_epi_chunk<ACT_SIGMOID, false>(p, stg, rowoff, rowoff16, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                    case ACT_TANH:     bf_epi_chunk<ACT_TANH, false>(p, stg, rowoff, rowoff16, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                    case ACT_SOFTSIGN: bf_epi_chunk<ACT_SOFTSIGN, false>(p, stg, rowoff, rowoff16, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                    default:           bf_epi_chunk<ACT_NONE, false>(p, stg, rowoff, rowoff16, okmask, maskmask, lr, lc, gn, true, cs, cq); break;
                }
                __syncwarp();                                        // staging buffer is rewritten by the next chunk
                if (p.colsum) {                                      // kernel-uniform
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);  cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 8);
                        cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16); cq[e] += __shfl_xor_sync(0xffffffffu, cq[e], 16);
                    }
                    if (lane < 8) {                                  // per-quarter partials, reduced over the CTA below
                        const uint32_t r = red_s + (uint32_t)((q * (2 * BF_BN_MAX) + c0 + lc) * 4);
                        sts_v4(r, cs[0], cs[1], cs[2], cs[3]);
                        sts_v4(r + BF_BN_MAX * 4, cq[0], cq[1], cq[2], cq[3]);
                    }
                }
            }
            // the accumulator is in registers / written out: hand the TMEM slot back before the statistics
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[slot]);
            if (warp == BF_EPI_WARP0) BF_STAMP(5 + 4 * lt);
            if (p.colsum) {
                // one double atomic per column and statistic per CTA (the four lane quarters are summed here first: the
                // statistics land on a few hundred addresses, so their atomics serialise in L2)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int et = threadIdx.x - 32 * BF_EPI_WARP0;
                for (int idx = et; idx < 2 * BN; idx += 32 * BF_EPI_WARPS) {
                    const int which = idx >= BN ? 1 : 0, col = idx - which * BN;
                    if (n0 + col < p.N) {
                        const uint32_t r = red_s + (uint32_t)((which * BF_BN_MAX + col) * 4);
                        const float v = (lds_f32(r) + lds_f32(r + 2 * BF_BN_MAX * 4)) + (lds_f32(r + 4 * BF_BN_MAX * 4) + lds_f32(r + 6 * BF_BN_MAX * 4));
                        atomicAdd((which ? p.colsumsq : p.colsum) + n0 + col, (double)v);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");       // partials are rewritten by the next tile
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == BF_PROD_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BF_BN_MAX));
    }
}

// ---------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode16 = nullptr;
static int get_encode16() {
    static std::once_flag once;
    static int rc = TACO_OK;
    std::call_once(once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { rc = TACO_ECUDA; return; }
        g_encode16 = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    if (rc != TACO_OK) set_error("cuTensorMapEncodeTiled entry point unavailable");
    return rc;
}

// Tensor maps depend on (base, extents, pitch, box) only and the model re-issues the same few hundred GEMMs every step:
// encode each distinct map once (cuTensorMapEncodeTiled costs ~1-2 us of host time per call).
struct MapKey {
    const void* base; uint64_t d0, d1, ld; uint32_t b0, b1;
    bool operator==(const MapKey& o) const { return base == o.base && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.base);
        for (uint64_t v : {k.d0, k.d1, k.ld, (uint64_t)k.b0, (uint64_t)k.b1}) h = h * 1099511628211ull ^ (size_t)v;
        return h;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

// bf16 2-D tensor map: dim0 contiguous (extent d0), dim1 rows (extent d1, pitch ld elements), box {b0, b1}, 128B swizzle
static int make_map16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1) {
    const MapKey key{base, d0, d1, ld, b0, b1};
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *map = it->second; return TACO_OK; }
    }
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode16(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TACO_REQUIRE(r == CUDA_SUCCESS, TACO_ECUDA, "cuTensorMapEncodeTiled (bf16) failed (%d): base=%p dims=(%llu,%llu) ld=%llu box=(%u,%u)", (int)r,
                 base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps.emplace(key, *map);
    return TACO_OK;
}

// 3-D view of a row-major [K, ld] bf16 matrix as {64 columns, K rows, column blocks of 64}: box {64, bk, nblk}
static int make_map16_3d(CUtensorMap* map, const void* base, uint64_t K, uint64_t nblocks, uint64_t ld, uint32_t bk, uint32_t nblk) {
    const MapKey key{base, K | (1ull << 62), nblocks, ld, bk, nblk};
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *map = it->second; return TACO_OK; }
    }
    cuuint64_t dims[3] = {64, K, nblocks};
    cuuint64_t strides[2] = {ld * 2, 128};
    cuuint32_t box[3] = {64, bk, nblk};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode16(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TACO_REQUIRE(r == CUDA_SUCCESS, TACO_ECUDA, "cuTensorMapEncodeTiled (bf16, 3-D) failed (%d): base=%p K=%llu blocks=%llu ld=%llu", (int)r,
                 base, (unsigned long long)K, (unsigned long long)nblocks, (unsigned long long)ld);
    std::lock_guard<std::mutex> lk(g_maps_mu);
    g_maps.emplace(key, *map);
    return TACO_OK;
}

// TMA needs 16-byte aligned bases and row pitches; taps must align to k-tiles (or be contiguous im2col rows).
bool gemm_bf16_eligible(const taco_gemm_desc& g) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (!g.A16 || !g.B16 || !al16(g.A16) || !al16(g.B16)) return false;
    if (g.lda % 8 != 0 || g.ldb % 8 != 0) return false;
    const bool tap = g.ctap > 0 && (g.lda != g.ctap || g.ctap % BF_BK == 0);
    if (tap && g.ctap % BF_BK != 0) return false;
    if (g.tap_table && (g.transA || g.K % BF_BK != 0 || g.ctap != 0 || (reinterpret_cast<uintptr_t>(g.tap_table) & 7u) != 0)) return false;
    if (!g.C && !g.C16) return false;
    return true;
}

static unsigned int* g_sched_ring[16] = {};      // per device: BF_SCHED_SLOTS x {next unit, CTAs done}; self-cleaning (see the kernel)
static unsigned int g_sched_next = 0;
static int g_n_sm[16] = {};
static std::mutex g_sched_mu;

int launch_gemm_bf16(const taco_gemm_desc& g, cudaStream_t s, bool leaf) {
    TACO_TRY(get_encode16());
    int dev = 0; TACO_CHECK_CUDA(cudaGetDevice(&dev));
    TACO_REQUIRE(dev >= 0 && dev < 16, TACO_ECUDA, "gemm: device ordinal %d out of range", dev);
    {
        std::lock_guard<std::mutex> lk(g_sched_mu);
        if (g_n_sm[dev] == 0) {
            TACO_CHECK_CUDA(cudaDeviceGetAttribute(&g_n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
            TACO_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM));
            TACO_CHECK_CUDA(cudaMalloc(&g_sched_ring[dev], sizeof(unsigned int) * 2 * BF_SCHED_SLOTS));
            TACO_CHECK_CUDA(cudaMemset(g_sched_ring[dev], 0, sizeof(unsigned int) * 2 * BF_SCHED_SLOTS));
        }
    }
    const int n_sm = g_n_sm[dev];
    TACO_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, TACO_ESHAPE, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
    TACO_REQUIRE(!((g.split_k > 1 || g.accumulate == 2) && (g.act != 0 || g.colsum != nullptr || g.C16 != nullptr)), TACO_EINVAL,
                 "gemm: atomic accumulation cannot carry an activation, column statistics or a bf16 mirror");
    TACO_REQUIRE(!(g.accumulate == 1 && !g.C), TACO_EINVAL, "gemm: accumulate needs the fp32 output");
    TACO_REQUIRE(g.tap_table == nullptr || g.K % BF_BK == 0, TACO_EINVAL, "gemm: tap table needs K %% 64 == 0");
    const void* A = g.A16;
    const void* B = g.B16;
    // lda == ctap with ctap % 64 != 0 (80-channel mel input): im2col rows are contiguous, so the map uses overlapping
    // rows (pitch lda < extent K)
    const bool tap = g.ctap > 0 && (g.lda != g.ctap || g.ctap % BF_BK == 0);
    const int ntaps = g.ctap > 0 ? ((g.transA ? g.M : g.K) + g.ctap - 1) / g.ctap : 1;
    BfParams p{};
    // tile width: the whole N when it fits one tile, otherwise the width that balances whole waves of persistent CTAs
    // (cost model: waves x (BN + 128): a tile's L2 -> SM bytes)
    int BN;
    if (g.N <= BF_BN_MAX) {
        BN = (g.N + 15) / 16 * 16;
        if (BN == 256) {
            static const int force = [] { const char* e = getenv("TACO_BF16_BN"); return e ? atoi(e) : 0; }();
            const long long t256 = (long long)cdiv(g.M, BF_BM), t128 = 2 * t256;
            const long long c256 = cdiv64(t256, n_sm) * (256 + 128), c128 = cdiv64(t128, n_sm) * (128 + 128);
            if ((force == 128) || (force == 0 && c128 < c256 && g.split_k <= 1)) BN = 128;
        }
    } else {
        const long long tm = cdiv(g.M, BF_BM);
        const long long c256 = cdiv64(tm * cdiv(g.N, 256), n_sm) * (256 + 128), c128 = cdiv64(tm * cdiv(g.N, 128), n_sm) * (128 + 128);
        BN = c128 < c256 ? 128 : 256;
    }
    CUtensorMap mapA, mapB;
    if (g.tap_table) {
        TACO_TRY(make_map16(&mapA, A, (uint64_t)g.lda, (uint64_t)g.tap_rows, (uint64_t)g.lda, BF_BK, BF_BM));
    } else if (!g.transA) {
        if (tap) TACO_TRY(make_map16(&mapA, A, (uint64_t)g.ctap, (uint64_t)g.M + ntaps - 1, (uint64_t)g.lda, BF_BK, BF_BM));
        else TACO_TRY(make_map16(&mapA, A, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda, BF_BK, BF_BM));
    } else {
        if (tap) TACO_TRY(make_map16(&mapA, A, (uint64_t)g.ctap, (uint64_t)g.K + ntaps - 1, (uint64_t)g.lda, 64, BF_BK));
        else TACO_TRY(make_map16(&mapA, A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda, 64, BF_BK));
    }
    // MN-major B wider than two 64-column blocks: one 3-D box {64 n, 64 k, BN/64 blocks} per k-tile when the blocks stay inside
    // the row pitch (N a multiple of 64, or the pitch padded to one)
    const bool b3d = !g.transB && BN > 128 && (g.N % 64 == 0 || g.ldb >= (g.N + 63) / 64 * 64);
    if (b3d) TACO_TRY(make_map16_3d(&mapB, B, (uint64_t)g.K, (uint64_t)cdiv(g.N, 64), (uint64_t)g.ldb, BF_BK, (uint32_t)cdiv(BN, 64)));
    else if (!g.transB) TACO_TRY(make_map16(&mapB, B, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldb, 64, BF_BK));            // MN-major
    else TACO_TRY(make_map16(&mapB, B, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldb, BF_BK, (uint32_t)BN));           // K-major
    p.C = g.C; p.C16 = static_cast<__nv_bfloat16*>(g.C16);
    p.M = g.M; p.N = g.N; p.K = g.K; p.ldc = g.ldc; p.ldc16 = g.ldc16 > 0 ? g.ldc16 : g.ldc;
    p.BN = BN; p.tilesN = cdiv(g.N, BN);
    p.b_stage_bytes = cdiv(BN, 64) * 8192;
    p.stages = BF_RING_BYTES / (BF_A_BYTES + p.b_stage_bytes);
    if (p.stages > BF_MAX_STAGES) p.stages = BF_MAX_STAGES;
    { static const int cap = [] { const char* e = getenv("TACO_BF16_STAGES"); return e ? atoi(e) : 0; }(); if (cap > 0 && p.stages > cap) p.stages = cap; }
    p.a_mn_major = g.transA ? 1 : 0; p.b_mn_major = g.transB ? 0 : 1; p.b_3d = b3d ? 1 : 0;
    p.a_tap = tap ? 1 : 0; p.a_ctap = tap ? g.ctap : 1;
    p.tap_table = reinterpret_cast<const int2*>(g.tap_table);
    p.tap_inner = (tap && !g.transA && !g.tap_table && g.ctap % BF_BK == 0 && g.K % g.ctap == 0 && g.K / g.ctap > 1) ? g.K / g.ctap : 0;
    {
        static const int want = [] { const char* e = getenv("TACO_BF16_TAPG"); return e ? atoi(e) : 0; }();
        const int blocks = g.ctap > 0 ? g.ctap / BF_BK : 1;
        p.tap_group = 1;
        for (int gsz = want; gsz >= 1; gsz--) if (blocks % gsz == 0) { p.tap_group = gsz; break; }
        if (want <= 0) p.tap_inner = 0;
    }
    p.alpha = g.alpha; p.accumulate = g.accumulate; p.bias = g.bias; p.act = g.act;
    p.mask_period = g.mask_period; p.mask_lo = g.mask_lo; p.mask_hi = g.mask_hi;
    p.remap_period = g.remap_period; p.remap_outer = g.remap_outer; p.remap_inner = g.remap_inner;
    p.colsum = g.colsum; p.colsumsq = g.colsumsq;
    p.ktiles = cdiv(g.K, BF_BK);
    const int tiles = cdiv(g.M, BF_BM) * p.tilesN;
    // A caller that allows K splitting (split_k > 1: C is accumulated atomically) gets the split that balances whole waves
    // of persistent CTAs: waves x (k-blocks per unit + a fixed per-unit cost in k-block units).
    // Leaf GEMMs (weight gradients on the low-priority stream of the backward schedule) must not hold the machine: a persistent
    // CTA owns its SM (225 KB of shared memory, 63 K registers) until its tile loop ends, so the high-priority chain - cluster
    // launches of the recurrences included - would wait behind it (measured: bn_bwd of the post-net bank 0.31 -> 0.60 ms).  They
    // run one short unit per CTA instead (<= 24 k-tiles, grid = units): SMs come free every ~20 us and go to the chain first.
    static const int leaf_mode = [] { const char* e = getenv("TACO_BF16_LEAF"); return e ? atoi(e) : 1; }();
    p.one_shot = (leaf && leaf_mode == 1) ? 1 : 0;
    p.split_k = 1;
    if (g.split_k > 1 && p.one_shot) {
        const int cap = p.ktiles / 4 > 0 ? p.ktiles / 4 : 1;
        p.split_k = std::min(cap, cdiv(p.ktiles, 24));
        if (p.split_k < 1) p.split_k = 1;
    } else if (g.split_k > 1) {
        double best = 1e30;
        for (int sp = 1; sp <= 128 && sp <= p.ktiles; sp++) {
            const int per = cdiv(p.ktiles, sp);
            if (sp > 1 && per < 4) break;
            const double cost = (double)cdiv(tiles * sp, n_sm) * (per + 6.0);
            if (cost < best - 1e-9) { best = cost; p.split_k = sp; }
        }
    }
    p.kt_per = cdiv(p.ktiles, p.split_k);
    p.split_k = cdiv(p.ktiles, p.kt_per);            // no empty splits
    p.units = tiles * p.split_k;
    auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    p.vecC = g.C ? (g.remap_period > 0 ? (al(g.C, 16) && g.remap_outer % 4 == 0 && g.remap_inner % 4 == 0) : (al(g.C, 16) && g.ldc % 4 == 0)) : 0;
    p.vecC16 = g.C16 ? (g.remap_period > 0 ? (al(g.C16, 8) && g.remap_outer % 4 == 0 && g.remap_inner % 4 == 0) : (al(g.C16, 8) && p.ldc16 % 4 == 0)) : 0;
    // lean epilogue: 128-bit stores on every output, 16-byte aligned bias, no read-modify-write, statistics only with relu / none
    p.fast_ok = (!g.C || p.vecC) && (!g.C16 || p.vecC16) && (!g.bias || al(g.bias, 16)) &&
                (g.accumulate != 1 || (g.act == ACT_NONE && !g.colsum)) && (!g.colsum || g.act == ACT_RELU || g.act == ACT_NONE) ? 1 : 0;
    {
        std::lock_guard<std::mutex> lk(g_sched_mu);
        p.sched = g_sched_ring[dev] + 2 * (g_sched_next++ % BF_SCHED_SLOTS);
    }
    const int grid = p.one_shot ? p.units : (p.units < n_sm ? p.units : n_sm);
    gemm_bf16_kernel<<<grid, BF_THREADS, BF_SMEM, s>>>(mapA, mapB, p);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

int debug_timeline_bf16(unsigned long long out[64]) {
    TACO_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_bf_stamp, sizeof(unsigned long long) * 64));
    return TACO_OK;
}

}  // namespace taco

extern "C" int taco_debug_timeline_bf16(unsigned long long out[64]) { return taco::debug_timeline_bf16(out); }
