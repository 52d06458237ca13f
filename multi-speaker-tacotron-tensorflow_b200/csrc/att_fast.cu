// Fast attention recurrence for the tensor-core precision mode (same maths, arguments and stash layout as attention.cu).
//   reference: models/rnn_wrappers.py:218-341,367-378,405-415; models/tacotron.py:127-170; SURVEY.md §8a D1-D7, App. C.
//
// A non-portable cluster of 16 CTAs owns 8 batch rows.  Everything the serial chain touches is shared-memory resident
// for all decoder steps: each CTA's bf16 slice of the decoder weights (1/16 of every layer's output units), its slice
// of the attention keys and of the encoder memory.  Per-step products run on mma.sync.m16n8k16 (weight rows = M, the 8
// batch rows = N); state, scores, the monotonic scan and all accumulation stay fp32.  Slices are exchanged with
// st.async DSMEM stores that complete on the receiver's mbarrier (7 exchanges per step, no cluster barriers).
#include "common.cuh"
#include "kernels.h"
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cfloat>

namespace cg = cooperative_groups;

namespace taco {

constexpr int AF_C = 16, AF_R = 8, AF_NT = 256;
constexpr int AF_E = 256, AF_A = 256, AF_HA = 256, AF_Z1 = 256, AF_Z = 128, AF_Y = 256;   // instantiated sizes
constexpr int AF_U = 16;                 // units per CTA of the 256-wide layers
constexpr int AF_UZ = AF_Z / AF_C;       // 8
// Score / alignment-gradient exchange buffers are blocked [C][R*TJ (+ AF_EPAD)] floats.  Dense blocks (64 words at T_in = 128) put
// every block on the same banks: the alignment scan, whose lanes walk the blocks, saw 16-way conflicts on each load
// (profiles/r1_ncu_full_att_fast.summary.txt).  With 8 words of padding a quarter-warp's 16-byte loads cover all 32 banks.
constexpr int AF_EPAD = 8;

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t af_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void af_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(af_u32(bar)), "r"(count));
}
__device__ __forceinline__ void af_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(af_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void af_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(af_u32(bar)), "r"(parity) : "memory");
}
// send this CTA's staged block (nbytes, multiple of 16) to the same offset `dst` in every peer, completing on the peer's `bar`
__device__ __forceinline__ void af_push(const void* stage, void* dst, uint64_t* bar, int nbytes, int tid) {
    const int nq = nbytes / 16;
    for (int idx = tid; idx < AF_C * nq; idx += AF_NT) {
        const uint32_t peer = idx / nq, q = idx % nq;
        const uint4 v = *(reinterpret_cast<const uint4*>(stage) + q);
        uint32_t d, b;
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(af_u32(dst) + q * 16), "r"(peer));
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(b) : "r"(af_u32(bar)), "r"(peer));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(d), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(b) : "memory");
    }
}
__device__ __forceinline__ void af_ldm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    // not volatile: the weight tiles are never written after the prologue, so the loads may be hoisted and batched
    asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void af_mma(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float af_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float af_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// acc += Wt[m0 .. m0+15][kw0 + 16*kt ...] . vec   for nk k-tiles;  Wt: bf16 [rows][KP]; vec: bf16 blocked [K/UB][R][UB], k index kv0 + ...
// All operand fragments of the NK k-tiles are fetched before the (dependent) mma chain starts.
template <int UB, int NK>
__device__ __forceinline__ void af_mma_run(float acc[4], const bf16* Wt, int KP, int m0, int kw0, const bf16* vec, int kv0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t a_base = af_u32(Wt + (size_t)(m0 + (lane & 15)) * KP + kw0 + (lane >> 4) * 8);
    uint32_t af[NK][4], bfr[NK][2];
#pragma unroll
    for (int kk = 0; kk < NK; kk++) {
        af_ldm4(a_base + (uint32_t)kk * 32, af[kk][0], af[kk][1], af[kk][2], af[kk][3]);
        const int ka = kv0 + kk * 16 + 2 * t, kb = ka + 8;
        bfr[kk][0] = *reinterpret_cast<const uint32_t*>(vec + ((ka / UB) * AF_R + g) * UB + (ka % UB));
        bfr[kk][1] = *reinterpret_cast<const uint32_t*>(vec + ((kb / UB) * AF_R + g) * UB + (kb % UB));
    }
    // one accumulator per k-tile, summed afterwards: a DEPENDENT mma.sync costs ~100 clk here (tools/gru_phase_prof.py), so a chain
    // of NK (or, over two calls on one accumulator, of 6) set the phase time rather than the issue rate
    float c[NK][4];
#pragma unroll
    for (int kk = 0; kk < NK; kk++) {
        c[kk][0] = 0.f; c[kk][1] = 0.f; c[kk][2] = 0.f; c[kk][3] = 0.f;
        af_mma(c[kk], af[kk][0], af[kk][1], af[kk][2], af[kk][3], bfr[kk][0], bfr[kk][1]);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
        float t = c[0][e];
#pragma unroll
        for (int kk = 1; kk < NK; kk++) t += c[kk][e];
        acc[e] += t;
    }
}
// fragment -> red[slot][r][m]  (8 batch rows x 16 output rows per slot, batch rows AF_RP floats apart).  The activation threads -
// 16 consecutive output rows of one batch row per half-warp - then read consecutive words; the [m][r] order this replaces made
// each of their (up to 8) reads per phase a 4-way bank conflict.  Stores: (2t)*20 + g covers 32 different banks.
constexpr int AF_RP = 20, AF_RSLOT = AF_R * AF_RP;
__device__ __forceinline__ void af_red_store(float* red, int slot, const float acc[4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    float* rp = red + slot * AF_RSLOT + g;
    rp[(2 * t) * AF_RP] = acc[0]; rp[(2 * t + 1) * AF_RP] = acc[1];
    rp[(2 * t) * AF_RP + 8] = acc[2]; rp[(2 * t + 1) * AF_RP + 8] = acc[3];
}
__device__ __forceinline__ float af_red_sum(const float* red, int slot0, int nslots, int stride, int m, int r) {
    float s = 0.f;
    for (int k = 0; k < nslots; k++) s += red[(slot0 + k * stride) * AF_RSLOT + r * AF_RP + m];
    return s;
}
// One-time operand loads (weights, keys, memory -> bf16 shared memory).  AF_LB loads are issued back to back before the first
// is consumed: with one load in flight per thread this set-up cost ~50 us (forward) / ~100 us (backward) per launch — paid
// once per sequence before, but once per time chunk under the decoder wavefront.
constexpr int AF_LB = 8;
template <typename Src, typename Dst>
__device__ __forceinline__ void af_batched_fill(int total, int tid, int nt, Src src, Dst dst) {
    for (int base = tid; base < total; base += nt * AF_LB) {
        float v[AF_LB];
#pragma unroll
        for (int u = 0; u < AF_LB; u++) { const int idx = base + u * nt; v[u] = idx < total ? src(idx) : 0.f; }
#pragma unroll
        for (int u = 0; u < AF_LB; u++) { const int idx = base + u * nt; if (idx < total) dst(idx, v[u]); }
    }
}
// load W[(r0 + m)*ld + c0 + k] (m < rows, k < K) -> dst[m*KP + k] as bf16 (natural orientation: rows of W are A rows)
__device__ void af_load_rows(bf16* dst, int KP, const float* W, long long ld, int r0, int rows, int rows_pad, int c0, int K, int tid, int nt = AF_NT) {
    af_batched_fill(rows_pad * K, tid, nt,
        [&](int idx) { const int m = idx / K, k = idx % K; return m < rows ? __ldg(W + (long long)(r0 + m) * ld + c0 + k) : 0.f; },
        [&](int idx, float v) { const int m = idx / K, k = idx % K; dst[m * KP + k] = __float2bfloat16(v); });
}
// load W[(r0 + k)*ld + c0 + m] -> dst[m*KP + k]  (transposed: columns of W become A rows)
__device__ void af_load_cols(bf16* dst, int KP, const float* W, long long ld, int r0, int K, int c0, int cols, int cols_pad, int tid, int nt = AF_NT) {
    af_batched_fill(cols_pad * K, tid, nt,
        [&](int idx) { const int k = idx / cols_pad, m = idx % cols_pad; return m < cols ? __ldg(W + (long long)(r0 + k) * ld + c0 + m) : 0.f; },
        [&](int idx, float v) { const int k = idx / cols_pad, m = idx % cols_pad; dst[m * KP + k] = __float2bfloat16(v); });
}

__device__ long long g_af_prof[40];
#define AF_T(i) do { if (prof) { long long _c = clock64(); pacc[i] += _c - plast; plast = _c; } } while (0)

// Monotonic attention forward for one row by one warp, lane owns the CH contiguous positions [lane*CH, lane*CH+CH):
// p = sigmoid(e); cp = exp(cumsum_excl(log(clip(1-p, tiny, 1)))); a = p*cp*cumsum(a_prev/clip(cp,1e-10,1))   (in place in ar_)
template <int CH>
__device__ __forceinline__ void af_mono_scan_fwd(const float* e_s, const int* eoff, bool vec4, float* ar_, int lane, int Ti) {
    const int j0 = lane * CH;
    float pv[CH], lv[CH], av[CH], wv[CH], cpv[CH], ev[CH];
    // the lane's CH = 4 scores are contiguous inside one exchange block when TJ % 4 == 0: one 16-byte load (conflict-free with the
    // padded block pitch, see AF_EPAD) instead of four 16-way conflicted scalar ones
    if (CH == 4 && vec4) {
        const float4 e4 = *reinterpret_cast<const float4*>(e_s + eoff[0]);
        ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
    } else {
#pragma unroll
        for (int k = 0; k < CH; k++) ev[k] = e_s[eoff[k]];
    }
    float ls = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        const bool ok = j0 + k < Ti;
        const float pk = af_sigmoid(ok ? ev[k] : 0.f);
        av[k] = ok ? ar_[j0 + k] : 0.f;
        pv[k] = ok ? pk : 0.f;
        lv[k] = ok ? __logf(fminf(fmaxf(1.f - pk, FLT_MIN), 1.f)) : 0.f;
        ls += lv[k];
    }
    float run = ls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
    run -= ls;
    float ws = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        cpv[k] = __expf(run);
        run += lv[k];
        wv[k] = __fdividef(av[k], fminf(fmaxf(cpv[k], 1e-10f), 1.f));
        ws += wv[k];
    }
    float run2 = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
    run2 -= ws;
#pragma unroll
    for (int k = 0; k < CH; k++) {
        run2 += wv[k];
        if (j0 + k < Ti) ar_[j0 + k] = pv[k] * cpv[k] * run2;
    }
}

struct AfFwdSmem {
    // bf16 weight slices (A operands, [rows][K+8])
    static constexpr int KP256 = 264, KP384 = 392, KP128 = 136;
    static constexpr size_t W1c = 0;                                   // [16][264]
    static constexpr size_t W2 = W1c + 16 * KP256 * 2;                 // [16][264] (8 valid rows)
    static constexpr size_t Wg = W2 + 16 * KP256 * 2;                  // [32][392]  rows: r units | u units; k: z(128) | ha(256)
    static constexpr size_t Wcz = Wg + 32 * KP384 * 2;                 // [16][136]
    static constexpr size_t Wch = Wcz + 16 * KP128 * 2;                // [16][264]
    static constexpr size_t Wq = Wch + 16 * KP256 * 2;                 // [16][264]
    static constexpr size_t Woh = Wq + 16 * KP256 * 2;                 // [16][264]
    static constexpr size_t Woc = Woh + 16 * KP256 * 2;                // [16][264]
    static constexpr size_t w_end = Woc + 16 * KP256 * 2;
};

// The operand image of one forward CTA = the first bytes of its shared memory: the bf16 weight slices [0, w_end) (they
// depend on the cluster rank only) followed by its key rows and memory columns (rank and row group).  The fill functions
// write an image either straight into shared memory or — once per step, att_fast_pack_kernel — into global memory, from
// where every (chunk) launch pulls it with bulk async copies instead of ~50 us of strided fp32 loads.
__device__ void af_fwd_fill_w(uint8_t* sm, const AttArgs& a, int rank, int tid, int nt) {
    using S = AfFwdSmem;
    constexpr int U = AF_U, UZ = AF_UZ, E = AF_E, A = AF_A, HA = AF_HA, Z1 = AF_Z1, Z = AF_Z, Y = AF_Y;
    bf16* W1c_s = reinterpret_cast<bf16*>(sm + S::W1c); bf16* W2_s = reinterpret_cast<bf16*>(sm + S::W2);
    bf16* Wg_s = reinterpret_cast<bf16*>(sm + S::Wg);   bf16* Wcz_s = reinterpret_cast<bf16*>(sm + S::Wcz);
    bf16* Wch_s = reinterpret_cast<bf16*>(sm + S::Wch); bf16* Wq_s = reinterpret_cast<bf16*>(sm + S::Wq);
    bf16* Woh_s = reinterpret_cast<bf16*>(sm + S::Woh); bf16* Woc_s = reinterpret_cast<bf16*>(sm + S::Woc);
    af_load_cols(W1c_s, S::KP256, a.W1c, Z1, 0, E, rank * U, U, 16, tid, nt);
    af_load_cols(W2_s, S::KP256, a.W2, Z, 0, Z1, rank * UZ, UZ, 16, tid, nt);
    af_load_cols(Wg_s, S::KP384, a.Wg, 2 * HA, 0, Z + HA, rank * U, U, 16, tid, nt);                      // r units
    af_load_cols(Wg_s + 16 * S::KP384, S::KP384, a.Wg, 2 * HA, 0, Z + HA, HA + rank * U, U, 16, tid, nt); // u units
    af_load_cols(Wcz_s, S::KP128, a.Wc, HA, 0, Z, rank * U, U, 16, tid, nt);
    af_load_cols(Wch_s, S::KP256, a.Wc, HA, Z, HA, rank * U, U, 16, tid, nt);
    af_load_cols(Wq_s, S::KP256, a.Wq, A, 0, HA, rank * U, U, 16, tid, nt);
    af_load_cols(Woh_s, S::KP256, a.Wo, Y, 0, HA, rank * U, U, 16, tid, nt);
    af_load_cols(Woc_s, S::KP256, a.Wo, Y, HA, E, rank * U, U, 16, tid, nt);
}
__host__ __device__ inline size_t af_fwd_km_bytes(int Ti) {
    const int TJ = (Ti + AF_C - 1) / AF_C;
    return (size_t)AF_R * TJ * AF_A * 2 + (size_t)AF_R * (Ti * AF_U + 16) * 2;
}
__device__ void af_fwd_fill_km(uint8_t* km, const AttArgs& a, int rank, int grp, int tid, int nt) {
    constexpr int R = AF_R, U = AF_U, E = AF_E, A = AF_A;
    const int Ti = a.Ti, TJ = (Ti + AF_C - 1) / AF_C, MRS = Ti * AF_U + 16;
    bf16* keys_s = reinterpret_cast<bf16*>(km);                                        // [R][TJ][A]
    bf16* mem_s = reinterpret_cast<bf16*>(km + (size_t)R * TJ * A * 2);                // [R][Ti][U] (+pad)
    af_batched_fill(R * TJ * A, tid, nt,
        [&](int idx) { const int u = idx % A, jj = (idx / A) % TJ, r = idx / (A * TJ), n = grp * R + r, j = rank * TJ + jj;
                       return (n < a.N && j < Ti) ? __ldg(a.keys + ((long long)n * Ti + j) * A + u) : 0.f; },
        [&](int idx, float v) { keys_s[idx] = __float2bfloat16(v); });
    af_batched_fill(R * Ti * U, tid, nt,
        [&](int idx) { const int i = idx % U, j = (idx / U) % Ti, r = idx / (U * Ti), n = grp * R + r;
                       return n < a.N ? __ldg(a.memory + ((long long)n * Ti + j) * E + rank * U + i) : 0.f; },
        [&](int idx, float v) { const int i = idx % U, j = (idx / U) % Ti, r = idx / (U * Ti); mem_s[(size_t)r * MRS + j * U + i] = __float2bfloat16(v); });
    for (int idx = tid; idx < R * 16; idx += nt) mem_s[(size_t)(idx / 16) * MRS + Ti * U + (idx % 16)] = __float2bfloat16(0.f);   // row padding
}
// Pull an image from global memory into shared memory: thread 0 issues bulk async copies (<= 32 KB each) that complete on
// `bar`; every thread then waits for the barrier's first phase.  dst / src / bytes are multiples of 16.
__device__ __forceinline__ void af_pull_image(uint8_t* dst, const uint8_t* src, size_t bytes, uint8_t* dst2, const uint8_t* src2, size_t bytes2,
                                              uint64_t* bar, int tid) {
    if (tid == 0) {
        af_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        af_expect(bar, (uint32_t)(bytes + bytes2));
        for (int part = 0; part < 2; part++) {
            uint8_t* d = part ? dst2 : dst; const uint8_t* g = part ? src2 : src; const size_t nb = part ? bytes2 : bytes;
            for (size_t off = 0; off < nb; off += 32768) {
                const uint32_t n = (uint32_t)((nb - off) < 32768 ? (nb - off) : 32768);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(af_u32(d + off)), "l"(g + off), "r"(n), "r"(af_u32(bar)) : "memory");
            }
        }
    }
    __syncthreads();
    af_wait(bar, 0);
}

template <int DUMMY>
__global__ void __launch_bounds__(AF_NT, 1) att_fast_fwd_kernel(const AttArgs a) {
    cg::cluster_group cl = cg::this_cluster();
    const long long c_entry = clock64();
    unsigned long long ns_entry; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_entry));
    const int rank = (int)cl.block_rank();
    const int grp = blockIdx.x / AF_C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int R = AF_R, U = AF_U, UZ = AF_UZ, E = AF_E, A = AF_A, HA = AF_HA, Z1 = AF_Z1, Z = AF_Z, Y = AF_Y;
    const int Ti = a.Ti, Td = a.Td;
    const int TJ = (Ti + AF_C - 1) / AF_C, Tip = TJ * AF_C;
    const int TipP = Tip + 4;                 // row pitch of the alignment rows (bank-conflict padding)
    const int MRS = Ti * AF_U + 16;           // row pitch (elements) of the per-row memory slice
    using S = AfFwdSmem;

    extern __shared__ __align__(128) uint8_t sm[];
    bf16* W1c_s = reinterpret_cast<bf16*>(sm + S::W1c); bf16* W2_s = reinterpret_cast<bf16*>(sm + S::W2);
    bf16* Wg_s = reinterpret_cast<bf16*>(sm + S::Wg);   bf16* Wcz_s = reinterpret_cast<bf16*>(sm + S::Wcz);
    bf16* Wch_s = reinterpret_cast<bf16*>(sm + S::Wch); bf16* Wq_s = reinterpret_cast<bf16*>(sm + S::Wq);
    bf16* Woh_s = reinterpret_cast<bf16*>(sm + S::Woh); bf16* Woc_s = reinterpret_cast<bf16*>(sm + S::Woc);
    uint8_t* p = sm + S::w_end;
    bf16* keys_s = reinterpret_cast<bf16*>(p); p += (size_t)R * TJ * A * 2;          // [R][TJ][A]   own memory positions
    bf16* mem_s = reinterpret_cast<bf16*>(p);  p += (size_t)R * MRS * 2;             // [R][Ti][U] (+pad)   own context units
    p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    bf16* ctx_s = reinterpret_cast<bf16*>(p); p += R * E * 2;                        // blocked [C][R][U]
    bf16* z1_s = reinterpret_cast<bf16*>(p);  p += R * Z1 * 2;
    bf16* z_s = reinterpret_cast<bf16*>(p);   p += R * Z * 2;                        // blocked [C][R][UZ]
    bf16* ha_s = reinterpret_cast<bf16*>(p);  p += R * HA * 2;
    bf16* rha_s = reinterpret_cast<bf16*>(p); p += R * HA * 2;
    float* q_s = reinterpret_cast<float*>(p); p += R * A * 4;                        // fp32 blocked [C][R][U]
    const int EB = R * TJ + AF_EPAD;          // block pitch (floats) of the score exchange buffer
    float* e_s = reinterpret_cast<float*>(p); p += (size_t)AF_C * EB * 4;            // fp32 blocked [C][R][TJ] (+pad)
    float* a_s = reinterpret_cast<float*>(p); p += (size_t)R * TipP * 4;             // [R][Tip+4]
    float* p_s = reinterpret_cast<float*>(p); p += (size_t)R * TipP * 4;
    float* cp_s = reinterpret_cast<float*>(p); p += (size_t)R * TipP * 4;
    float* red = reinterpret_cast<float*>(p); p += 16 * AF_RSLOT * 4;                // 16 slots of [8][16 (+4)]
    float* v_s = reinterpret_cast<float*>(p); p += A * 4;
    uint8_t* stage = p; p += 1024;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);                                 // 7 barriers
    uint64_t *b_z1 = bars, *b_z = bars + 1, *b_rha = bars + 2, *b_ha = bars + 3, *b_q = bars + 4, *b_e = bars + 5, *b_ctx = bars + 6;

    // ---- one-time loads (see af_fwd_fill_w) -------------------------------------------------------------------
    if (a.img_w && a.img_km) {
        const size_t kmb = af_fwd_km_bytes(Ti);
        af_pull_image(sm, a.img_w + (size_t)rank * S::w_end, S::w_end, sm + S::w_end, a.img_km + (size_t)blockIdx.x * kmb, kmb, bars + 7, tid);
    } else {
        af_fwd_fill_w(sm, a, rank, tid, AF_NT);
        af_fwd_fill_km(sm + S::w_end, a, rank, grp, tid, AF_NT);
    }
    for (int u = tid; u < A; u += AF_NT) v_s[u] = a.v[u];
    // time chunk [t_lo, t_hi): a chunk that does not start the sequence restores its state from the stash of step t_lo-1
    // (s_ha / s_ctx / s_a hold exactly the fp32 values the loop carries, so a chunked run equals the unchunked one)
    const int t_lo = a.t_begin, t_hi = a.t_end > 0 ? min(a.t_end, Td) : Td;
    const bool resume = t_lo > 0;
    for (int idx = tid; idx < R * E; idx += AF_NT) {        // blocked [C][R][U]
        const int blk = idx / (R * U), r = (idx / U) % R, i = idx % U, n = grp * R + r;
        ctx_s[idx] = __float2bfloat16((resume && n < a.N) ? a.s_ctx[((long long)n * Td + t_lo - 1) * E + blk * U + i] : 0.f);
    }
    for (int idx = tid; idx < R * HA; idx += AF_NT) {       // blocked [C][R][U]
        const int blk = idx / (R * U), r = (idx / U) % R, i = idx % U, n = grp * R + r;
        float hv = 0.f;
        if (n < a.N) hv = resume ? a.s_ha[((long long)n * Td + t_lo - 1) * HA + blk * U + i] : (a.ha0 ? a.ha0[(long long)n * HA + blk * U + i] : 0.f);
        ha_s[idx] = __float2bfloat16(hv);
    }
    for (int idx = tid; idx < R * TipP; idx += AF_NT) {
        const int j = idx % TipP, r = idx / TipP, n = grp * R + r;
        float av = (a.att_type == TACO_ATT_BAH_MON && j == 0) ? 1.f : 0.f;
        if (resume) av = (n < a.N && j < Ti) ? a.s_a[((long long)n * Td + t_lo - 1) * Ti + j] : 0.f;
        a_s[idx] = av;
    }
    if (tid == 0) {
        for (int i = 0; i < 7; i++) af_mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const float score_bias = (a.att_type == TACO_ATT_BAH_MON) ? a.score_bias[0] : 0.f;
    // activation-thread coordinates: 16-unit layers use threads [0,128): (i, r)
    const bool act = tid < U * R;
    const int ai = tid % U, ar = (tid / U) % R;
    const int an = grp * R + ar;
    const bool aok = act && an < a.N;
    const int unit = rank * U + ai;
    float ha_own = 0.f;
    if (aok) ha_own = resume ? a.s_ha[((long long)an * Td + t_lo - 1) * HA + unit] : (a.ha0 ? a.ha0[(long long)an * HA + unit] : 0.f);
    float px_next = aok ? __ldg(a.px + ((long long)an * Td + t_lo) * Z1 + unit) : 0.f;
    const float bg_r = act ? __ldg(a.bg + unit) : 0.f, bg_u = act ? __ldg(a.bg + HA + unit) : 0.f;
    const float bc_own = act ? __ldg(a.bc + unit) : 0.f, bo_own = act ? __ldg(a.bo + unit) : 0.f;
    const float b2_own = (tid < UZ * R) ? __ldg(a.b2 + rank * UZ + (tid % UZ)) : 0.f;
    float ctx_prev = (resume && aok) ? a.s_ctx[((long long)an * Td + t_lo - 1) * E + unit] : 0.f;   // fp32 context of the previous step (own unit), for the W1c-gradient stash
    const long long c_loaded = clock64();
    __syncthreads();
    cl.sync();

    const bool prof = (blockIdx.x == 0 && tid == 0);
    long long pacc[20]; for (int q = 0; q < 20; q++) pacc[q] = 0; long long plast = clock64();
    pacc[15] = c_loaded - c_entry; pacc[16] = plast - c_loaded;      // set-up: operand loads | first cluster barrier
    // per-lane constants of the alignment scan (warp r = row r, lane owns the contiguous chunk [lane*CH, lane*CH+CH))
    constexpr int CHM = 8;
    const int CH = (Tip + 31) / 32;
    int eoff[CHM];
#pragma unroll
    for (int k = 0; k < CHM; k++) { const int j = min(lane * CH + k, Tip - 1); eoff[k] = (j / TJ) * EB + warp * TJ + (j % TJ); }
    const bool evec4 = (TJ % 4 == 0);
    const bool writer = (rank == (warp % AF_C)) && (grp * R + warp < a.N);     // this warp stores row `warp`'s alignments / stash
    for (int t = t_lo; t < t_hi; t++) {
        const uint32_t par = (t - t_lo) & 1;
        const long long row = (long long)an * Td + t;
        const float px_cur = px_next;
        if (tid == 0) {
            af_expect(b_z1, AF_C * R * U * 2); af_expect(b_z, AF_C * R * UZ * 2); af_expect(b_rha, AF_C * R * U * 2);
            af_expect(b_ha, AF_C * R * U * 2); af_expect(b_q, AF_C * R * U * 4); af_expect(b_e, AF_C * R * TJ * 4);
            af_expect(b_ctx, AF_C * R * U * 2);
        }
        if (aok && t + 1 < Td) px_next = __ldg(a.px + (row + 1) * Z1 + unit);
        // ===== P1: z1 = relu(px + ctx.W1c)   (ctx_s holds ctx_{t-1}) =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, W1c_s, S::KP256, 0, warp * 32, ctx_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float z1v = 0.f;
        if (act) {
            if (aok) z1v = fmaxf(af_red_sum(red, 0, 8, 1, ai, ar) + px_cur, 0.f);
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(z1v);
        }
        __syncthreads();
        af_push(stage, z1_s + rank * R * U, b_z1, R * U * 2, tid);
        if (aok && a.s_z1) {       // stash stores ride in the shadow of the exchange
            a.s_z1[row * Z1 + unit] = z1v;
            a.s_ctxin[row * E + unit] = ctx_prev;
        }
        AF_T(0);
        af_wait(b_z1, par);
        AF_T(1);
        // ===== P2: z = relu(z1.W2 + b2)  (own UZ = 8 units) =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, W2_s, S::KP256, 0, warp * 32, z1_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float zv = 0.f;
        const int zi = tid % UZ, zr = tid / UZ, zn = grp * R + zr;
        if (tid < UZ * R) {
            if (zn < a.N) zv = fmaxf(af_red_sum(red, 0, 8, 1, zi, zr) + b2_own, 0.f);
            reinterpret_cast<bf16*>(stage)[zr * UZ + zi] = __float2bfloat16(zv);
        }
        __syncthreads();
        af_push(stage, z_s + rank * R * UZ, b_z, R * UZ * 2, tid);
        if (tid < UZ * R && zn < a.N && a.s_z) a.s_z[((long long)zn * Td + t) * Z + rank * UZ + zi] = zv;
        AF_T(2);
        af_wait(b_z, par);
        AF_T(3);
        // ===== P3: gates (own 2U columns: r | u) over [z ; ha], and the z part of the candidate =====
        {
            const int mt = warp & 1, ks = warp >> 1;             // 2 m-tiles x 4 k-splits
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            // k-split ks takes z k-tiles [2ks, 2ks+2) and ha k-tiles [4ks, 4ks+4) (weight columns 128 + ...)
            af_mma_run<UZ, 2>(acc, Wg_s, S::KP384, mt * 16, ks * 32, z_s, ks * 32, lane);
            af_mma_run<U, 4>(acc, Wg_s, S::KP384, mt * 16, Z + ks * 64, ha_s, ks * 64, lane);
            af_red_store(red, warp, acc, lane);                  // slot = ks*2 + mt
            float acc2[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<UZ, 1>(acc2, Wcz_s, S::KP128, 0, warp * 16, z_s, warp * 16, lane);
            af_red_store(red, 8 + warp, acc2, lane);
        }
        __syncthreads();
        float rg = 0.f, ug = 0.f, cz = 0.f;
        if (act) {
            const float sr = af_red_sum(red, 0, 4, 2, ai, ar) + bg_r;
            const float su = af_red_sum(red, 1, 4, 2, ai, ar) + bg_u;
            cz = af_red_sum(red, 8, 8, 1, ai, ar);
            rg = af_sigmoid(sr); ug = af_sigmoid(su);
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(rg * ha_own);
        }
        __syncthreads();
        af_push(stage, rha_s + rank * R * U, b_rha, R * U * 2, tid);
        AF_T(4);
        af_wait(b_rha, par);
        AF_T(5);
        // ===== P4: candidate and new attention-GRU state =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, Wch_s, S::KP256, 0, warp * 32, rha_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float cc = 0.f, hprev = ha_own;
        if (act) {
            cc = af_tanh(af_red_sum(red, 0, 8, 1, ai, ar) + cz + bc_own);
            const float hn = ug * ha_own + (1.f - ug) * cc;
            ha_own = aok ? hn : 0.f;
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(ha_own);
        }
        __syncthreads();
        af_push(stage, ha_s + rank * R * U, b_ha, R * U * 2, tid);
        if (aok && a.s_r) {
            const long long o = row * HA + unit;
            a.s_r[o] = rg; a.s_u[o] = ug; a.s_c[o] = cc; a.s_haprev[o] = hprev; a.s_ha[o] = ha_own;
        }
        AF_T(6);
        af_wait(b_ha, par);
        AF_T(7);
        // ===== P5: query (own U columns) and the ha part of the concat projection (own U columns) =====
        {
            const int mt = warp & 1, ks = warp >> 1;             // m-tile 0 = Wq, 1 = Wo_h; 4 k-splits of 4 k-tiles
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 4>(acc, mt ? Woh_s : Wq_s, S::KP256, 0, ks * 64, ha_s, ks * 64, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float yh = 0.f, qv = 0.f;
        if (act) {
            qv = af_red_sum(red, 0, 4, 2, ai, ar);
            yh = af_red_sum(red, 1, 4, 2, ai, ar);
            reinterpret_cast<float*>(stage)[ar * U + ai] = qv;
        }
        __syncthreads();
        af_push(stage, q_s + rank * R * U, b_q, R * U * 4, tid);
        if (aok && a.s_q) a.s_q[row * A + unit] = qv;
        AF_T(8);
        af_wait(b_q, par);
        AF_T(9);
        // ===== P6: scores of the own TJ memory positions (warp r = row r): e[r][j] = sum_u v_u tanh(keys[r,j,u] + q[r,u]) + b =====
        {
            const int r = warp;
            const float* qp = q_s + ((lane >> 1) * R + r) * U + (lane & 1) * 8;
            const float4 q0 = *reinterpret_cast<const float4*>(qp), q1 = *reinterpret_cast<const float4*>(qp + 4);
            const float4 v0 = *reinterpret_cast<const float4*>(v_s + lane * 8), v1 = *reinterpret_cast<const float4*>(v_s + lane * 8 + 4);
            for (int jb = 0; jb < TJ; jb += 4) {
                float sc4[4];
#pragma unroll
                for (int u4 = 0; u4 < 4; u4++) {                 // four independent positions in flight
                    const int jj = min(jb + u4, TJ - 1);
                    const uint4 kk = *reinterpret_cast<const uint4*>(keys_s + ((size_t)(r * TJ + jj)) * A + lane * 8);
                    const __nv_bfloat162* kb = reinterpret_cast<const __nv_bfloat162*>(&kk);
                    const float2 k0 = __bfloat1622float2(kb[0]), k1 = __bfloat1622float2(kb[1]), k2 = __bfloat1622float2(kb[2]), k3 = __bfloat1622float2(kb[3]);
                    float s0 = v0.x * af_tanh(k0.x + q0.x), s1 = v0.y * af_tanh(k0.y + q0.y);
                    s0 = fmaf(v0.z, af_tanh(k1.x + q0.z), s0); s1 = fmaf(v0.w, af_tanh(k1.y + q0.w), s1);
                    s0 = fmaf(v1.x, af_tanh(k2.x + q1.x), s0); s1 = fmaf(v1.y, af_tanh(k2.y + q1.y), s1);
                    s0 = fmaf(v1.z, af_tanh(k3.x + q1.z), s0); s1 = fmaf(v1.w, af_tanh(k3.y + q1.w), s1);
                    sc4[u4] = s0 + s1;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int u4 = 0; u4 < 4; u4++) sc4[u4] += __shfl_xor_sync(0xffffffffu, sc4[u4], o);
                }
                const float mine = lane == 0 ? sc4[0] : (lane == 1 ? sc4[1] : (lane == 2 ? sc4[2] : sc4[3]));
                if (lane < 4 && jb + lane < TJ) reinterpret_cast<float*>(stage)[r * TJ + jb + lane] = mine + score_bias;
            }
        }
        __syncthreads();
        af_push(stage, e_s + rank * EB, b_e, R * TJ * 4, tid);
        AF_T(10);
        af_wait(b_e, par);
        AF_T(11);
        // ===== P7: alignments (every CTA redundantly; warp r = row r), then the own U context units =====
        {
            const int r = warp, n = grp * R + r;
            float* ar_ = a_s + r * TipP;
            if (a.manual) {
                for (int j = lane; j < Ti; j += 32) ar_[j] = (n < a.N) ? a.manual[((long long)n * Td + t) * Ti + j] : 0.f;
            } else if (a.att_type == TACO_ATT_BAH_MON && CH == 4) {
                af_mono_scan_fwd<4>(e_s, eoff, evec4, ar_, lane, Ti);
            } else if (a.att_type == TACO_ATT_BAH_MON && CH <= CHM) {
                // p = sigmoid(e); cp = exp(cumsum_excl(log(clip(1-p, tiny, 1)))); a = p*cp*cumsum(a_prev/clip(cp,1e-10,1)), fully unrolled
                const int j0 = lane * CH;
                float pv[CHM], lv[CHM], av[CHM];
                float ls = 0.f;
#pragma unroll
                for (int k = 0; k < CHM; k++) {
                    const bool ok = (k < CH) && (j0 + k < Ti);
                    const float ev = ok ? e_s[eoff[k]] : 0.f;
                    av[k] = ok ? ar_[j0 + k] : 0.f;
                    const float pk = af_sigmoid(ev);
                    pv[k] = ok ? pk : 0.f;
                    lv[k] = ok ? __logf(fminf(fmaxf(1.f - pk, FLT_MIN), 1.f)) : 0.f;
                    ls += lv[k];
                }
                float run = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
                run -= ls;
                float wv[CHM], cpv[CHM];
                float ws = 0.f;
#pragma unroll
                for (int k = 0; k < CHM; k++) {
                    cpv[k] = __expf(run);
                    run += lv[k];
                    wv[k] = __fdividef(av[k], fminf(fmaxf(cpv[k], 1e-10f), 1.f));
                    ws += wv[k];
                }
                float run2 = ws;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
                run2 -= ws;
#pragma unroll
                for (int k = 0; k < CHM; k++) {
                    run2 += wv[k];
                    if ((k < CH) && (j0 + k < Ti)) ar_[j0 + k] = pv[k] * cpv[k] * run2;
                }
            } else if (a.att_type == TACO_ATT_BAH_MON) {
                const int j0 = min(lane * CH, Ti), j1 = min(j0 + CH, Ti);
                float* pr_ = p_s + r * TipP; float* cr_ = cp_s + r * TipP;
                auto E_ = [&](int j) -> float { return e_s[(j / TJ) * EB + r * TJ + (j % TJ)]; };
                float ls = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float pk = af_sigmoid(E_(j));
                    pr_[j] = pk;
                    ls += __logf(fminf(fmaxf(1.f - pk, FLT_MIN), 1.f));
                }
                float run = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
                run -= ls;
                float ws = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float cp = __expf(run);
                    cr_[j] = cp;
                    run += __logf(fminf(fmaxf(1.f - pr_[j], FLT_MIN), 1.f));
                    ws += __fdividef(ar_[j], fminf(fmaxf(cp, 1e-10f), 1.f));
                }
                float run2 = ws;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
                run2 -= ws;
                for (int j = j0; j < j1; j++) {
                    run2 += __fdividef(ar_[j], fminf(fmaxf(cr_[j], 1e-10f), 1.f));
                    ar_[j] = pr_[j] * cr_[j] * run2;
                }
            } else {
                float* pr_ = p_s + r * TipP;
                auto E_ = [&](int j) -> float { return e_s[(j / TJ) * EB + r * TJ + (j % TJ)]; };
                float mx = -INFINITY;
                for (int j = lane; j < Ti; j += 32) mx = fmaxf(mx, E_(j));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float smv = 0.f;
                for (int j = lane; j < Ti; j += 32) { const float ex = __expf(E_(j) - mx); pr_[j] = ex; smv += ex; }
                smv = warp_sum(smv);
                for (int j = lane; j < Ti; j += 32) ar_[j] = __fdividef(pr_[j], smv);
            }
        }
        __syncthreads();
        {
            // ctx[r][own units]: thread = (j-quarter, row, unit pair); partials reduced through `red`
            const int jq = tid >> 6, r = (tid >> 3) & 7, ip = tid & 7;
            const int jn = (Ti + 3) / 4, ja = jq * jn, jb = min(Ti, ja + jn);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            const float* arow = a_s + r * TipP;
            const __nv_bfloat162* mp = reinterpret_cast<const __nv_bfloat162*>(mem_s + (size_t)r * MRS) + ip;
            int j = ja;
            for (; j + 1 < jb; j += 2) {
                const float2 m2 = __bfloat1622float2(mp[(size_t)j * (U / 2)]), m3 = __bfloat1622float2(mp[(size_t)(j + 1) * (U / 2)]);
                const float av0 = arow[j], av1 = arow[j + 1];
                c0 = fmaf(av0, m2.x, c0); c1 = fmaf(av0, m2.y, c1); c2 = fmaf(av1, m3.x, c2); c3 = fmaf(av1, m3.y, c3);
            }
            if (j < jb) { const float2 m2 = __bfloat1622float2(mp[(size_t)j * (U / 2)]); c0 = fmaf(arow[j], m2.x, c0); c1 = fmaf(arow[j], m2.y, c1); }
            *reinterpret_cast<float2*>(red + jq * AF_RSLOT + r * AF_RP + 2 * ip) = make_float2(c0 + c2, c1 + c3);
        }
        __syncthreads();
        float cx = 0.f;
        if (act) {
            if (aok) cx = af_red_sum(red, 0, 4, 1, ai, ar);
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(cx);
        }
        __syncthreads();
        af_push(stage, ctx_s + rank * R * U, b_ctx, R * U * 2, tid);
        // stores in the shadow of the exchange: context stash, alignments, scores
        ctx_prev = cx;
        if (aok && a.s_ctx) a.s_ctx[row * E + unit] = cx;
        if (writer) {
            const int r = warp, n = grp * R + r;
            const float* ar_ = a_s + r * TipP;
            for (int j = lane; j < Ti; j += 32) {
                const float av = ar_[j];
                a.align[((long long)n * Ti + j) * Td + t] = av;
                if (a.s_a) { a.s_a[((long long)n * Td + t) * Ti + j] = av; a.s_e[((long long)n * Td + t) * Ti + j] = e_s[(j / TJ) * EB + r * TJ + (j % TJ)]; }
            }
        }
        AF_T(12);
        af_wait(b_ctx, par);
        AF_T(13);
        // ===== P8: y0 (own U columns) = yh + ctx.Wo_c + bo =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, Woc_s, S::KP256, 0, warp * 32, ctx_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        if (aok) a.y0[row * Y + unit] = yh + af_red_sum(red, 0, 8, 1, ai, ar) + bo_own;
        __syncthreads();     // `red` is rewritten by the next step's P1
        AF_T(14);
    }
    if (prof) {
        unsigned long long ns_exit; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_exit));
        pacc[17] = (long long)(ns_exit - ns_entry);                     // nanoseconds inside the kernel (CTA 0)
        for (int q = 0; q < 20; q++) g_af_prof[q] = pacc[q];
    }
    if (a.ha_final && aok) a.ha_final[(long long)an * HA + unit] = ha_own;
    cl.sync();
}

struct AfBwdSmem {
    static constexpr int KP256 = 264, KP512 = 520, KP128 = 136;
    static constexpr size_t Wo = 0;                                    // [32][264]  rows: ha units | ctx units ; k over Y
    static constexpr size_t Wq = Wo + 32 * KP256 * 2;                  // [16][264]  k over A
    static constexpr size_t Wch = Wq + 16 * KP256 * 2;                 // [16][264]  rows Z+unit of cand kernel, k over HA
    static constexpr size_t Wcz = Wch + 16 * KP256 * 2;                // [16][264]  rows z units (8 valid), k over HA
    static constexpr size_t W1c = Wcz + 16 * KP256 * 2;                // [16][264]  rows ctx units of dense_1, k over Z1
    static constexpr size_t Wgh = W1c + 16 * KP256 * 2;                // [16][520]  rows Z+unit of gates kernel, k over 2HA
    static constexpr size_t Wgz = Wgh + 16 * KP512 * 2;                // [16][520]  rows z units (8 valid)
    static constexpr size_t W2 = Wgz + 16 * KP512 * 2;                 // [16][136]  rows z1 units of dense_2, k over Z
    static constexpr size_t w_end = W2 + 16 * KP128 * 2;
};

// backward operand image (same scheme as af_fwd_fill_w / af_fwd_fill_km)
__device__ void af_bwd_fill_w(uint8_t* sm, const AttArgs& a, int rank, int tid, int nt) {
    using S = AfBwdSmem;
    constexpr int U = AF_U, UZ = AF_UZ, A = AF_A, HA = AF_HA, Z1 = AF_Z1, Z = AF_Z, Y = AF_Y;
    bf16* Wo_s = reinterpret_cast<bf16*>(sm + S::Wo);   bf16* Wq_s = reinterpret_cast<bf16*>(sm + S::Wq);
    bf16* Wch_s = reinterpret_cast<bf16*>(sm + S::Wch); bf16* Wcz_s = reinterpret_cast<bf16*>(sm + S::Wcz);
    bf16* W1c_s = reinterpret_cast<bf16*>(sm + S::W1c); bf16* Wgh_s = reinterpret_cast<bf16*>(sm + S::Wgh);
    bf16* Wgz_s = reinterpret_cast<bf16*>(sm + S::Wgz); bf16* W2_s = reinterpret_cast<bf16*>(sm + S::W2);
    // natural-orientation weight rows (data-gradient products): A row = weight row of the owned input unit
    af_load_rows(Wo_s, S::KP256, a.Wo, Y, rank * U, U, 16, 0, Y, tid, nt);
    af_load_rows(Wo_s + 16 * S::KP256, S::KP256, a.Wo, Y, HA + rank * U, U, 16, 0, Y, tid, nt);
    af_load_rows(Wq_s, S::KP256, a.Wq, A, rank * U, U, 16, 0, A, tid, nt);
    af_load_rows(Wch_s, S::KP256, a.Wc, HA, Z + rank * U, U, 16, 0, HA, tid, nt);
    af_load_rows(Wcz_s, S::KP256, a.Wc, HA, rank * UZ, UZ, 16, 0, HA, tid, nt);
    af_load_rows(W1c_s, S::KP256, a.W1c, Z1, rank * U, U, 16, 0, Z1, tid, nt);
    af_load_rows(Wgh_s, S::KP512, a.Wg, 2 * HA, Z + rank * U, U, 16, 0, 2 * HA, tid, nt);
    af_load_rows(Wgz_s, S::KP512, a.Wg, 2 * HA, rank * UZ, UZ, 16, 0, 2 * HA, tid, nt);
    af_load_rows(W2_s, S::KP128, a.W2, Z, rank * U, U, 16, 0, Z, tid, nt);
}
__host__ __device__ inline size_t af_bwd_km_bytes(int Ti) {
    const int TJ = (Ti + AF_C - 1) / AF_C;
    return (size_t)AF_R * TJ * AF_E * 2 + (size_t)AF_R * (Ti * AF_U + 16) * 2;
}
__device__ void af_bwd_fill_km(uint8_t* km, const AttArgs& a, int rank, int grp, int tid, int nt) {
    constexpr int R = AF_R, U = AF_U, E = AF_E, A = AF_A;
    const int Ti = a.Ti, TJ = (Ti + AF_C - 1) / AF_C, KRS = Ti * AF_U + 16;
    bf16* memj_s = reinterpret_cast<bf16*>(km);                                        // [R][TJ][E]   own memory positions
    bf16* keyu_s = reinterpret_cast<bf16*>(km + (size_t)R * TJ * E * 2);               // [R][Ti][U] (+pad)   own attention units
    af_batched_fill(R * TJ * E, tid, nt,
        [&](int idx) { const int u = idx % E, jj = (idx / E) % TJ, r = idx / (E * TJ), n = grp * R + r, j = rank * TJ + jj;
                       return (n < a.N && j < Ti) ? __ldg(a.memory + ((long long)n * Ti + j) * E + u) : 0.f; },
        [&](int idx, float v) { memj_s[idx] = __float2bfloat16(v); });
    af_batched_fill(R * Ti * U, tid, nt,
        [&](int idx) { const int i = idx % U, j = (idx / U) % Ti, r = idx / (U * Ti), n = grp * R + r;
                       return n < a.N ? __ldg(a.keys + ((long long)n * Ti + j) * A + rank * U + i) : 0.f; },
        [&](int idx, float v) { const int i = idx % U, j = (idx / U) % Ti, r = idx / (U * Ti); keyu_s[(size_t)r * KRS + j * U + i] = __float2bfloat16(v); });
    for (int idx = tid; idx < R * 16; idx += nt) keyu_s[(size_t)(idx / 16) * KRS + Ti * U + (idx % 16)] = __float2bfloat16(0.f);
}

// Builds the operand images once per step.  what: 0 = weight slices (grid = AF_C blocks, block = cluster rank),
// 1 = key / memory slices (grid = one block per CTA of the recurrence kernel).
template <int BWD>
__global__ void __launch_bounds__(1024) att_fast_pack_kernel(const AttArgs a, uint8_t* img_w, uint8_t* img_km, int what) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (what == 0) {
        const int rank = blockIdx.x;
        if (BWD) af_bwd_fill_w(img_w + (size_t)rank * AfBwdSmem::w_end, a, rank, tid, nt);
        else af_fwd_fill_w(img_w + (size_t)rank * AfFwdSmem::w_end, a, rank, tid, nt);
    } else {
        const int rank = blockIdx.x % AF_C, grp = blockIdx.x / AF_C;
        if (BWD) af_bwd_fill_km(img_km + (size_t)blockIdx.x * af_bwd_km_bytes(a.Ti), a, rank, grp, tid, nt);
        else af_fwd_fill_km(img_km + (size_t)blockIdx.x * af_fwd_km_bytes(a.Ti), a, rank, grp, tid, nt);
    }
}

template <int DUMMY>
__global__ void __launch_bounds__(AF_NT, 1) att_fast_bwd_kernel(const AttArgs a) {
    cg::cluster_group cl = cg::this_cluster();
    const long long c_entry = clock64();
    unsigned long long ns_entry; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_entry));
    const int rank = (int)cl.block_rank();
    const int grp = blockIdx.x / AF_C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int R = AF_R, U = AF_U, UZ = AF_UZ, E = AF_E, A = AF_A, HA = AF_HA, Z1 = AF_Z1, Z = AF_Z, Y = AF_Y;
    const int Ti = a.Ti, Td = a.Td;
    const int TJ = (Ti + AF_C - 1) / AF_C, Tip = TJ * AF_C;
    const int TipP = Tip + 4;                 // row pitch of ge_s (bank-conflict padding)
    const int KRS = Ti * AF_U + 16;           // row pitch (elements) of the per-row key slice
    using S = AfBwdSmem;

    extern __shared__ __align__(128) uint8_t sm[];
    bf16* Wo_s = reinterpret_cast<bf16*>(sm + S::Wo);   bf16* Wq_s = reinterpret_cast<bf16*>(sm + S::Wq);
    bf16* Wch_s = reinterpret_cast<bf16*>(sm + S::Wch); bf16* Wcz_s = reinterpret_cast<bf16*>(sm + S::Wcz);
    bf16* W1c_s = reinterpret_cast<bf16*>(sm + S::W1c); bf16* Wgh_s = reinterpret_cast<bf16*>(sm + S::Wgh);
    bf16* Wgz_s = reinterpret_cast<bf16*>(sm + S::Wgz); bf16* W2_s = reinterpret_cast<bf16*>(sm + S::W2);
    uint8_t* p = sm + S::w_end;
    bf16* memj_s = reinterpret_cast<bf16*>(p); p += (size_t)R * TJ * E * 2;          // [R][TJ][E]   own memory positions
    bf16* keyu_s = reinterpret_cast<bf16*>(p); p += (size_t)R * KRS * 2;             // [R][Ti][U] (+pad)   own attention units
    p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    float* dctx_s = reinterpret_cast<float*>(p); p += R * E * 4;                     // fp32 blocked [C][R][U]
    const int EB = R * TJ + AF_EPAD;          // block pitch (floats) of the alignment-gradient exchange buffer
    float* da_s = reinterpret_cast<float*>(p);   p += (size_t)AF_C * EB * 4;         // fp32 blocked [C][R][TJ] (+pad)
    bf16* gq_s = reinterpret_cast<bf16*>(p);   p += R * A * 2;                       // bf16 blocked [C][R][U]
    bf16* dcp_s = reinterpret_cast<bf16*>(p);  p += R * HA * 2;
    bf16* dg_s = reinterpret_cast<bf16*>(p);   p += 2 * R * HA * 2;                  // [2C][R][U]
    bf16* dzp_s = reinterpret_cast<bf16*>(p);  p += R * Z * 2;                       // blocked [C][R][UZ]
    bf16* dz1p_s = reinterpret_cast<bf16*>(p); p += R * Z1 * 2;
    bf16* dy_s = reinterpret_cast<bf16*>(p);   p += R * Y * 2;                       // blocked [Y/16][R][16]
    float* dac_s = reinterpret_cast<float*>(p); p += (size_t)R * Tip * 4;            // carried grad wrt a_t
    float* ge_s = reinterpret_cast<float*>(p);  p += (size_t)R * TipP * 4;           // grad wrt scores (also scratch t1)
    float* p_s = reinterpret_cast<float*>(p);   p += (size_t)R * Tip * 4;
    float* cp_s = reinterpret_cast<float*>(p);  p += (size_t)R * Tip * 4;
    float* s_s = reinterpret_cast<float*>(p);   p += (size_t)R * Tip * 4;            // (also scratch t2)
    float* red = reinterpret_cast<float*>(p);   p += 16 * AF_RSLOT * 4;
    float* v_s = reinterpret_cast<float*>(p);   p += A * 4;
    uint8_t* stage = p; p += 1024;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);
    uint64_t *b_dctx = bars, *b_da = bars + 1, *b_gq = bars + 2, *b_dcp = bars + 3, *b_dg = bars + 4, *b_dzp = bars + 5, *b_dz1p = bars + 6;

    if (a.img_w && a.img_km) {
        const size_t kmb = af_bwd_km_bytes(Ti);
        af_pull_image(sm, a.img_w + (size_t)rank * S::w_end, S::w_end, sm + S::w_end, a.img_km + (size_t)blockIdx.x * kmb, kmb, bars + 7, tid);
    } else {
        af_bwd_fill_w(sm, a, rank, tid, AF_NT);
        af_bwd_fill_km(sm + S::w_end, a, rank, grp, tid, AF_NT);
    }
    for (int u = tid; u < A; u += AF_NT) v_s[u] = a.v[u];
    // time chunk [t_lo, t_hi), processed downwards; the gradients carried across steps (wrt ha, context, alignments) come in
    // from the chunk above through c_dha / c_dctx / c_dac when c_in is set and are handed on the same way at the end
    const int t_lo = a.t_begin, t_hi = a.t_end > 0 ? min(a.t_end, Td) : Td;
    const bool carry_in = a.c_in != 0 && a.c_dha && a.c_dctx && a.c_dac;
    for (int idx = tid; idx < R * Tip; idx += AF_NT) {
        const int r = idx / Tip, j = idx % Tip, n = grp * R + r;
        dac_s[idx] = (carry_in && n < a.N) ? a.c_dac[(long long)n * Tip + j] : 0.f;
    }
    for (int idx = tid; idx < R * TipP; idx += AF_NT) ge_s[idx] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < 7; i++) af_mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool act = tid < U * R;
    const int ai = tid % U, ar = (tid / U) % R;
    const int an = grp * R + ar;
    const bool aok = act && an < a.N;
    const int unit = rank * U + ai;
    // (j-quarter, row, unit pair) mapping of the score-gradient phase
    const int jq = tid >> 6, jr = (tid >> 3) & 7, jip = tid & 7;
    const int jn_ = grp * R + jr;
    float dha_carry = 0.f, dctx_carry = 0.f, gbias_acc = 0.f;
    if (carry_in && aok) { dha_carry = a.c_dha[(long long)an * HA + unit]; dctx_carry = a.c_dctx[(long long)an * E + unit]; }
    const long long c_loaded = clock64();
    __syncthreads();
    cl.sync();

    const bool prof = (blockIdx.x == 0 && tid == 0);
    long long pacc[20]; for (int q = 0; q < 20; q++) pacc[q] = 0; long long plast = clock64();
    pacc[15] = c_loaded - c_entry; pacc[16] = plast - c_loaded;      // set-up: operand loads | first cluster barrier
    // ---- per-thread constants and the one-step-ahead prefetch of everything this loop reads from global memory ----
    constexpr int CHM = 8;
    const int CH = (Tip + 31) / 32;
    int eoff[CHM];
#pragma unroll
    for (int k = 0; k < CHM; k++) { const int j = min(lane * CH + k, Tip - 1); eoff[k] = (j / TJ) * EB + warp * TJ + (j % TJ); }
    const int wn = grp * R + warp;                                // batch row of this warp in the scan phases
    const bool wok = wn < a.N;
    const bool writer = (rank == (warp % AF_C)) && wok;
    const bool mono_fast = (a.att_type == TACO_ATT_BAH_MON) && (CH <= CHM) && !a.manual;
    const int zi = tid % UZ, zr = tid / UZ, zn = grp * R + zr;
    const bool zok = (tid < UZ * R) && zn < a.N;
    float n_rg = 0.f, n_ug = 0.f, n_cc = 0.f, n_hp = 0.f, n_z1 = 0.f, n_z = 0.f, n_q0 = 0.f, n_q1 = 0.f;
    float n_dy[R], n_e[CHM], n_ap[CHM];
    auto prefetch = [&](int tt) {
        if (tt < t_lo) return;
        if (aok) {
            const long long o = ((long long)an * Td + tt) * HA + unit;
            n_rg = a.s_r[o]; n_ug = a.s_u[o]; n_cc = a.s_c[o]; n_hp = a.s_haprev[o];
            n_z1 = a.s_z1[((long long)an * Td + tt) * Z1 + unit];
        }
        if (zok) n_z = a.s_z[((long long)zn * Td + tt) * Z + rank * UZ + zi];
        if (jn_ < a.N && !a.manual) {
            const float* qp = a.s_q + ((long long)jn_ * Td + tt) * A + rank * U + 2 * jip;
            n_q0 = qp[0]; n_q1 = qp[1];
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int n = grp * R + r;
            n_dy[r] = (n < a.N) ? __ldg(a.dy0 + ((long long)n * Td + tt) * Y + tid) : 0.f;      // Y == AF_NT
        }
        if (mono_fast && wok) {
#pragma unroll
            for (int k = 0; k < CHM; k++) {
                const int j = lane * CH + k;
                const bool ok = (k < CH) && (j < Ti);
                n_e[k] = ok ? a.s_e[((long long)wn * Td + tt) * Ti + j] : 0.f;
                n_ap[k] = ok ? (tt > 0 ? a.s_a[((long long)wn * Td + (tt - 1)) * Ti + j] : (j == 0 ? 1.f : 0.f)) : 0.f;
            }
        }
    };
    prefetch(t_hi - 1);
    int it = 0;
    for (int t = t_hi - 1; t >= t_lo; t--, it++) {
        const uint32_t par = it & 1;
        const long long row = (long long)an * Td + t;
        if (tid == 0) {
            af_expect(b_dctx, AF_C * R * U * 4); af_expect(b_da, AF_C * R * TJ * 4); af_expect(b_gq, AF_C * R * U * 2);
            af_expect(b_dcp, AF_C * R * U * 2); af_expect(b_dg, 2 * AF_C * R * U * 2); af_expect(b_dzp, AF_C * R * UZ * 2);
            af_expect(b_dz1p, AF_C * R * U * 2);
        }
        const float rg = n_rg, ug = n_ug, cc = n_cc, hp = n_hp, z1v = n_z1, zval = n_z, q0 = n_q0, q1 = n_q1;
        float ev[CHM], apv[CHM];
#pragma unroll
        for (int k = 0; k < CHM; k++) { ev[k] = n_e[k]; apv[k] = n_ap[k]; }
#pragma unroll
        for (int r = 0; r < R; r++) dy_s[((tid >> 4) * R + r) * 16 + (tid & 15)] = __float2bfloat16(n_dy[r]);
        __syncthreads();
        // ===== Bp1: dha += dy0.Wo_h^T (own U ha units), dctx = carry + dy0.Wo_c^T (own U ctx units) =====
        {
            const int mt = warp & 1, ks = warp >> 1;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<16, 4>(acc, Wo_s, S::KP256, mt * 16, ks * 64, dy_s, ks * 64, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float dha = 0.f, dctx = 0.f;
        if (act) {
            dha = dha_carry + af_red_sum(red, 0, 4, 2, ai, ar);
            dctx = aok ? dctx_carry + af_red_sum(red, 1, 4, 2, ai, ar) : 0.f;
            reinterpret_cast<float*>(stage)[ar * U + ai] = dctx;
        }
        __syncthreads();
        af_push(stage, dctx_s + rank * R * U, b_dctx, R * U * 4, tid);
        prefetch(t - 1);      // next iteration's inputs (~25 loads and their address arithmetic) are issued in the shadow of the first exchange
        if (aok) a.d_ctx[row * E + unit] = dctx;
        AF_T(0);
        af_wait(b_dctx, par);
        AF_T(1);
        // ===== Bp2: da[r][own j] = sum_u dctx[r,u] memory[r,j,u]   (warp r = row r, four positions in flight) =====
        {
            const int r = warp;
            const float* dp = dctx_s + ((lane >> 1) * R + r) * U + (lane & 1) * 8;
            const float4 d0 = *reinterpret_cast<const float4*>(dp), d1 = *reinterpret_cast<const float4*>(dp + 4);
            for (int jb = 0; jb < TJ; jb += 4) {
                float sc4[4];
#pragma unroll
                for (int u4 = 0; u4 < 4; u4++) {
                    const int jj = min(jb + u4, TJ - 1);
                    const uint4 mm = *reinterpret_cast<const uint4*>(memj_s + ((size_t)(r * TJ + jj)) * E + lane * 8);
                    const __nv_bfloat162* mb = reinterpret_cast<const __nv_bfloat162*>(&mm);
                    const float2 m0 = __bfloat1622float2(mb[0]), m1 = __bfloat1622float2(mb[1]), m2 = __bfloat1622float2(mb[2]), m3 = __bfloat1622float2(mb[3]);
                    float s0 = d0.x * m0.x, s1 = d0.y * m0.y;
                    s0 = fmaf(d0.z, m1.x, s0); s1 = fmaf(d0.w, m1.y, s1); s0 = fmaf(d1.x, m2.x, s0); s1 = fmaf(d1.y, m2.y, s1);
                    s0 = fmaf(d1.z, m3.x, s0); s1 = fmaf(d1.w, m3.y, s1);
                    sc4[u4] = s0 + s1;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int u4 = 0; u4 < 4; u4++) sc4[u4] += __shfl_xor_sync(0xffffffffu, sc4[u4], o);
                }
                const float mine = lane == 0 ? sc4[0] : (lane == 1 ? sc4[1] : (lane == 2 ? sc4[2] : sc4[3]));
                if (lane < 4 && jb + lane < TJ) reinterpret_cast<float*>(stage)[r * TJ + jb + lane] = mine;
            }
        }
        __syncthreads();
        af_push(stage, da_s + rank * EB, b_da, R * TJ * 4, tid);
        AF_T(2);
        af_wait(b_da, par);
        AF_T(3);
        // ===== Bp3: attention-probability backward (every CTA; warp r = row r).  SURVEY.md Appendix C =====
        {
            const int r = warp;
            float* gc = dac_s + r * Tip; float* ge = ge_s + r * TipP;
            if (a.manual || !wok) {
                for (int j = lane; j < Tip; j += 32) { ge[j] = 0.f; gc[j] = 0.f; }
            } else if (mono_fast) {
                // forward: q=clip(1-p); L=cumsum_excl(log q); cp=exp(L); d=clip(cp,1e-10,1); w=aprev/d; s=cumsum(w); a=p*cp*s  -- all in registers
                const int j0 = lane * CH;
                float pv[CHM], lv[CHM], cpv[CHM], sv[CHM], gv[CHM], gsv[CHM], gLv[CHM], gpd[CHM];
                float ls = 0.f;
#pragma unroll
                for (int k = 0; k < CHM; k++) {
                    const bool ok = (k < CH) && (j0 + k < Ti);
                    const float pk = af_sigmoid(ev[k]);
                    pv[k] = ok ? pk : 0.f;
                    lv[k] = ok ? __logf(fminf(fmaxf(1.f - pk, FLT_MIN), 1.f)) : 0.f;
                    gv[k] = ok ? da_s[eoff[k]] + gc[j0 + k] : 0.f;       // total grad wrt a_t[j]
                    ls += lv[k];
                }
                float run = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
                run -= ls;
                float ws = 0.f, wv[CHM];
#pragma unroll
                for (int k = 0; k < CHM; k++) {
                    cpv[k] = __expf(run);
                    run += lv[k];
                    wv[k] = __fdividef(apv[k], fminf(fmaxf(cpv[k], 1e-10f), 1.f));
                    ws += wv[k];
                }
                float run2 = ws;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
                run2 -= ws;
                float gs_loc = 0.f;
#pragma unroll
                for (int k = 0; k < CHM; k++) {
                    run2 += wv[k];
                    sv[k] = run2;
                    gsv[k] = gv[k] * pv[k] * cpv[k];
                    gs_loc += gsv[k];
                }
                float suf = gs_loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
                suf -= gs_loc;
                float gL_loc = 0.f;
                {
                    float acc = suf;
#pragma unroll
                    for (int k = CHM - 1; k >= 0; k--) {
                        acc += gsv[k];                                   // gw_j = sum_{i>=j} gs_i
                        const float d = fminf(fmaxf(cpv[k], 1e-10f), 1.f);
                        const float rd = __fdividef(1.f, d);
                        float gcp = gv[k] * pv[k] * sv[k];
                        if (cpv[k] >= 1e-10f && cpv[k] <= 1.f) gcp -= acc * apv[k] * rd * rd;
                        gLv[k] = gcp * cpv[k];
                        gL_loc += gLv[k];
                        gpd[k] = gv[k] * cpv[k] * sv[k];
                        if ((k < CH) && (j0 + k < Ti)) gc[j0 + k] = acc * rd;   // grad wrt a_{t-1,j}, carried to the next iteration
                    }
                }
                float sufL = gL_loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_down_sync(0xffffffffu, sufL, o); if (lane + o < 32) sufL += v; }
                sufL -= gL_loc;
                float gb = 0.f;
                {
                    float acc = sufL;
#pragma unroll
                    for (int k = CHM - 1; k >= 0; k--) {
                        const float omp = 1.f - pv[k];
                        float gp = gpd[k];
                        if (omp >= FLT_MIN && omp <= 1.f) gp -= __fdividef(acc, fminf(fmaxf(omp, FLT_MIN), 1.f));
                        acc += gLv[k];
                        const float gev = gp * pv[k] * omp;
                        if ((k < CH) && (j0 + k < Ti)) { ge[j0 + k] = gev; gb += gev; }
                    }
                }
                for (int j = Ti + lane; j < Tip; j += 32) { ge[j] = 0.f; gc[j] = 0.f; }
                gb = warp_sum(gb);
                if (rank == 0 && lane == 0) gbias_acc += gb;
            } else if (a.att_type == TACO_ATT_BAH_MON) {
                // generic path for very long memories (T_in > 256)
                const int n = wn;
                const int j0 = min(lane * CH, Ti), j1 = min(j0 + CH, Ti);
                float* pr_ = p_s + r * Tip; float* cr_ = cp_s + r * Tip; float* sr_ = s_s + r * Tip;
                float* t1 = ge; float* t2 = sr_;
                auto GA = [&](int j) -> float { return da_s[(j / TJ) * EB + r * TJ + (j % TJ)]; };
                const float* e_row = a.s_e + ((long long)n * Td + t) * Ti;
                const float* ap_row = (t > 0) ? a.s_a + ((long long)n * Td + (t - 1)) * Ti : nullptr;
                float ls = 0.f;
                for (int j = j0; j < j1; j++) { const float pk = af_sigmoid(e_row[j]); pr_[j] = pk; ls += __logf(fminf(fmaxf(1.f - pk, FLT_MIN), 1.f)); }
                float run = ls;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += v; }
                run -= ls;
                float ws = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float cp = __expf(run); cr_[j] = cp;
                    run += __logf(fminf(fmaxf(1.f - pr_[j], FLT_MIN), 1.f));
                    const float ap = ap_row ? ap_row[j] : (j == 0 ? 1.f : 0.f);
                    ws += __fdividef(ap, fminf(fmaxf(cp, 1e-10f), 1.f));
                }
                float run2 = ws;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_up_sync(0xffffffffu, run2, o); if (lane >= o) run2 += v; }
                run2 -= ws;
                float gs_loc = 0.f;
                for (int j = j0; j < j1; j++) {
                    const float ap = ap_row ? ap_row[j] : (j == 0 ? 1.f : 0.f);
                    run2 += __fdividef(ap, fminf(fmaxf(cr_[j], 1e-10f), 1.f));
                    const float g = GA(j) + gc[j];
                    gc[j] = g; sr_[j] = run2;
                    const float gs = g * pr_[j] * cr_[j];
                    t1[j] = gs; gs_loc += gs;
                }
                float suf = gs_loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += v; }
                suf -= gs_loc;
                float gL_loc = 0.f;
                {
                    float acc = suf;
                    for (int j = j1 - 1; j >= j0; j--) {
                        acc += t1[j];
                        const float g = gc[j], pk = pr_[j], cp = cr_[j], sj = sr_[j];
                        const float ap = ap_row ? ap_row[j] : (j == 0 ? 1.f : 0.f);
                        const float d = fminf(fmaxf(cp, 1e-10f), 1.f);
                        float gcp = g * pk * sj;
                        if (cp >= 1e-10f && cp <= 1.f) gcp -= acc * ap / (d * d);
                        const float gL = gcp * cp;
                        gL_loc += gL; ge[j] = g * cp * sj; t2[j] = gL; gc[j] = acc / d;
                    }
                }
                float sufL = gL_loc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { float v = __shfl_down_sync(0xffffffffu, sufL, o); if (lane + o < 32) sufL += v; }
                sufL -= gL_loc;
                float gb = 0.f;
                {
                    float acc = sufL;
                    for (int j = j1 - 1; j >= j0; j--) {
                        const float pk = pr_[j], omp = 1.f - pk;
                        float gp = ge[j];
                        if (omp >= FLT_MIN && omp <= 1.f) gp -= acc / fminf(fmaxf(omp, FLT_MIN), 1.f);
                        acc += t2[j];
                        const float gev = gp * pk * (1.f - pk);
                        ge[j] = gev; gb += gev;
                    }
                }
                for (int j = Ti + lane; j < Tip; j += 32) { ge[j] = 0.f; gc[j] = 0.f; }
                gb = warp_sum(gb);
                if (rank == 0 && lane == 0) gbias_acc += gb;
            } else {
                auto GA = [&](int j) -> float { return da_s[(j / TJ) * EB + r * TJ + (j % TJ)]; };
                const float* a_row = a.s_a + ((long long)wn * Td + t) * Ti;
                float dot = 0.f;
                for (int j = lane; j < Ti; j += 32) dot += a_row[j] * GA(j);
                dot = warp_sum(dot);
                for (int j = lane; j < Tip; j += 32) { ge[j] = (j < Ti) ? a_row[j] * (GA(j) - dot) : 0.f; gc[j] = 0.f; }
            }
        }
        __syncthreads();
        // ===== Bp4: gq (own U units) = v_u * sum_j ge[r][j] * (1 - tanh^2(keys[r,j,u] + q[r,u])) =====
        {
            const int jn = (Ti + 3) / 4, ja = jq * jn, jb = min(Ti, ja + jn);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            if (!a.manual) {
                const float* ger = ge_s + jr * TipP;
                const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(keyu_s + (size_t)jr * KRS) + jip;
                int j = ja;
                for (; j + 1 < jb; j += 2) {
                    const float2 ka = __bfloat1622float2(kp[(size_t)j * (U / 2)]), kb = __bfloat1622float2(kp[(size_t)(j + 1) * (U / 2)]);
                    const float g0 = ger[j], g1 = ger[j + 1];
                    const float ta = af_tanh(ka.x + q0), tb = af_tanh(ka.y + q1), tc = af_tanh(kb.x + q0), td = af_tanh(kb.y + q1);
                    c0 = fmaf(g0, 1.f - ta * ta, c0); c1 = fmaf(g0, 1.f - tb * tb, c1);
                    c2 = fmaf(g1, 1.f - tc * tc, c2); c3 = fmaf(g1, 1.f - td * td, c3);
                }
                if (j < jb) {
                    const float2 ka = __bfloat1622float2(kp[(size_t)j * (U / 2)]);
                    const float ta = af_tanh(ka.x + q0), tb = af_tanh(ka.y + q1);
                    c0 = fmaf(ger[j], 1.f - ta * ta, c0); c1 = fmaf(ger[j], 1.f - tb * tb, c1);
                }
            }
            *reinterpret_cast<float2*>(red + jq * AF_RSLOT + jr * AF_RP + 2 * jip) = make_float2(c0 + c2, c1 + c3);
        }
        __syncthreads();
        float gq = 0.f;
        if (act) {
            if (aok) gq = af_red_sum(red, 0, 4, 1, ai, ar) * v_s[unit];
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(gq);
        }
        __syncthreads();
        af_push(stage, gq_s + rank * R * U, b_gq, R * U * 2, tid);
        if (aok && a.d_gq) a.d_gq[row * A + unit] = gq;
        if (writer && a.d_ge) {
            const float* ge = ge_s + warp * TipP;
            for (int j = lane; j < Ti; j += 32) a.d_ge[((long long)wn * Td + t) * Ti + j] = ge[j];
        }
        AF_T(4);
        af_wait(b_gq, par);
        AF_T(5);
        // ===== Bp5: dha += gq.Wq^T; GRU cell backward (elementwise part) =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, Wq_s, S::KP256, 0, warp * 32, gq_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float du_pre = 0.f, dc_pre = 0.f;
        if (act) {
            dha += af_red_sum(red, 0, 8, 1, ai, ar);
            if (aok) {
                du_pre = dha * (hp - cc) * ug * (1.f - ug);
                dc_pre = dha * (1.f - ug) * (1.f - cc * cc);
            }
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(dc_pre);
        }
        __syncthreads();
        af_push(stage, dcp_s + rank * R * U, b_dcp, R * U * 2, tid);
        AF_T(6);
        af_wait(b_dcp, par);
        AF_T(7);
        // ===== Bp6: d(r*h) (own U) = dc_pre . Wc_h^T =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, Wch_s, S::KP256, 0, warp * 32, dcp_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float d_rh = 0.f, dr_pre = 0.f;
        if (act) {
            d_rh = af_red_sum(red, 0, 8, 1, ai, ar);
            if (aok) dr_pre = d_rh * hp * rg * (1.f - rg);
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(dr_pre);
            reinterpret_cast<bf16*>(stage + 256)[ar * U + ai] = __float2bfloat16(du_pre);
        }
        __syncthreads();
        af_push(stage, dg_s + rank * R * U, b_dg, R * U * 2, tid);
        af_push(stage + 256, dg_s + (AF_C + rank) * R * U, b_dg, R * U * 2, tid);
        if (aok) {
            float* g = a.d_G + row * 3 * HA + unit;
            g[0] = dr_pre; g[HA] = du_pre; g[2 * HA] = dc_pre;
            a.s_r[row * HA + unit] = rg * hp;                   // operand of the candidate-weight gradient GEMM
        }
        AF_T(8);
        af_wait(b_dg, par);
        AF_T(9);
        // ===== Bp7: dha_prev (own U ha units) and dz (own UZ z units) =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 4>(acc, Wgh_s, S::KP512, 0, warp * 64, dg_s, warp * 64, lane);
            af_red_store(red, warp, acc, lane);
            float acc2[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 4>(acc2, Wgz_s, S::KP512, 0, warp * 64, dg_s, warp * 64, lane);
            af_mma_run<U, 2>(acc2, Wcz_s, S::KP256, 0, warp * 32, dcp_s, warp * 32, lane);
            af_red_store(red, 8 + warp, acc2, lane);
        }
        __syncthreads();
        if (aok) dha_carry = dha * ug + d_rh * rg + af_red_sum(red, 0, 8, 1, ai, ar);
        float dzv = 0.f;
        if (tid < UZ * R) {
            if (zok) dzv = (zval > 0.f) ? af_red_sum(red, 8, 8, 1, zi, zr) : 0.f;
            reinterpret_cast<bf16*>(stage)[zr * UZ + zi] = __float2bfloat16(dzv);
        }
        __syncthreads();
        af_push(stage, dzp_s + rank * R * UZ, b_dzp, R * UZ * 2, tid);
        if (zok) a.d_zp[((long long)zn * Td + t) * Z + rank * UZ + zi] = dzv;
        AF_T(10);
        af_wait(b_dzp, par);
        AF_T(11);
        // ===== Bp8: dz1 (own U) = dz_pre . W2^T, relu' =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<UZ, 1>(acc, W2_s, S::KP128, 0, warp * 16, dzp_s, warp * 16, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        float dz1v = 0.f;
        if (act) {
            if (aok) dz1v = (z1v > 0.f) ? af_red_sum(red, 0, 8, 1, ai, ar) : 0.f;
            reinterpret_cast<bf16*>(stage)[ar * U + ai] = __float2bfloat16(dz1v);
        }
        __syncthreads();
        af_push(stage, dz1p_s + rank * R * U, b_dz1p, R * U * 2, tid);
        if (aok) a.d_z1p[row * Z1 + unit] = dz1v;
        AF_T(12);
        af_wait(b_dz1p, par);
        AF_T(13);
        // ===== Bp9: grad wrt ctx_{t-1} (own U) = dz1_pre . W1c^T =====
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            af_mma_run<U, 2>(acc, W1c_s, S::KP256, 0, warp * 32, dz1p_s, warp * 32, lane);
            af_red_store(red, warp, acc, lane);
        }
        __syncthreads();
        if (act) dctx_carry = af_red_sum(red, 0, 8, 1, ai, ar);
        __syncthreads();
        AF_T(14);
    }
    if (prof) {
        unsigned long long ns_exit; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_exit));
        pacc[17] = (long long)(ns_exit - ns_entry);
        for (int q = 0; q < 20; q++) g_af_prof[20 + q] = pacc[q];
    }
    if (a.d_ha0 && aok && t_lo == 0) a.d_ha0[(long long)an * HA + unit] = dha_carry;
    if (a.c_dha && a.c_dctx && a.c_dac) {                       // hand the carries to the chunk below
        if (aok) { a.c_dha[(long long)an * HA + unit] = dha_carry; a.c_dctx[(long long)an * E + unit] = dctx_carry; }
        if (rank == 0) {
            for (int idx = tid; idx < R * Tip; idx += AF_NT) {
                const int r = idx / Tip, j = idx % Tip, n = grp * R + r;
                if (n < a.N) a.c_dac[(long long)n * Tip + j] = dac_s[idx];
            }
        }
    }
    if (rank == 0 && lane == 0 && a.d_score_bias && a.att_type == TACO_ATT_BAH_MON) atomicAdd(a.d_score_bias, gbias_acc);
    cl.sync();
}

static size_t af_bwd_smem(int Ti) {
    const int TJ = (Ti + AF_C - 1) / AF_C, Tip = TJ * AF_C;
    size_t b = AfBwdSmem::w_end + (size_t)AF_R * TJ * AF_E * 2 + (size_t)AF_R * (Ti * AF_U + 16) * 2 + 16 + (size_t)AF_R * 4 * 4;
    b += (size_t)AF_R * AF_E * 4 + (size_t)(AF_R * Tip + AF_C * AF_EPAD) * 4 + (size_t)AF_R * (AF_A + AF_HA + 2 * AF_HA + AF_Z + AF_Z1 + AF_Y) * 2;
    b += (size_t)5 * AF_R * Tip * 4 + 16 * AF_RSLOT * 4 + AF_A * 4 + 1024 + 64 + 128;
    return b;
}

static size_t af_fwd_smem(int Ti) {
    const int TJ = (Ti + AF_C - 1) / AF_C, Tip = TJ * AF_C;
    size_t b = AfFwdSmem::w_end + (size_t)AF_R * TJ * AF_A * 2 + (size_t)AF_R * (Ti * AF_U + 16) * 2 + 16;
    b += (size_t)AF_R * (AF_E + AF_Z1 + AF_Z + 2 * AF_HA) * 2 + (size_t)AF_R * AF_A * 4 + (size_t)4 * AF_R * (Tip + 4) * 4 + (size_t)AF_C * AF_EPAD * 4;
    b += 16 * AF_RSLOT * 4 + AF_A * 4 + 1024 + 64 + 128;
    return b;
}

bool att_fast_supported(const AttArgs& a) {
    return a.E == AF_E && a.A == AF_A && a.HA == AF_HA && a.Z1 == AF_Z1 && a.Z == AF_Z && a.Y == AF_Y && a.SPK == 0 &&
           a.att_type != TACO_ATT_BAH_NORM && af_fwd_smem(a.Ti) <= 227 * 1024 && af_bwd_smem(a.Ti) <= 227 * 1024;
}

int af_debug_prof(long long out[40]) { TACO_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_af_prof, sizeof(long long) * 40)); return TACO_OK; }

template <typename K>
static int af_launch(K kern, const AttArgs& a, size_t smem, int& configured, cudaStream_t s) {
    if (configured == 0) {       // 0 unknown, 1 ok, -1 unsupported (a cluster of 16 CTAs is not launchable here)
        cudaError_t e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured = (e1 == cudaSuccess && e2 == cudaSuccess) ? 1 : -1;
        if (configured < 0) cudaGetLastError();
    }
    if (configured < 0) return TACO_ENOTSUP;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(AF_C * cdiv(a.N, AF_R));
    cfg.blockDim = dim3(AF_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = AF_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) { cudaGetLastError(); configured = -1; return TACO_ENOTSUP; }
    g_launch_count++;
    return TACO_OK;
}

void att_fast_image_bytes(int Ti, bool bwd, size_t* w_bytes, size_t* km_bytes_per_cta) {
    *w_bytes = (size_t)AF_C * (bwd ? AfBwdSmem::w_end : AfFwdSmem::w_end);
    *km_bytes_per_cta = bwd ? af_bwd_km_bytes(Ti) : af_fwd_km_bytes(Ti);
}
// what: 0 weight slices, 1 key / memory slices (see att_fast_pack_kernel)
int launch_att_fast_pack(const AttArgs& a, bool bwd, int what, uint8_t* img_w, uint8_t* img_km, cudaStream_t s) {
    const int blocks = what == 0 ? AF_C : AF_C * cdiv(a.N, AF_R);
    if (bwd) att_fast_pack_kernel<1><<<blocks, 1024, 0, s>>>(a, img_w, img_km, what);
    else att_fast_pack_kernel<0><<<blocks, 1024, 0, s>>>(a, img_w, img_km, what);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

int launch_att_fast_fwd(const AttArgs& a, cudaStream_t s) {
    static int configured = 0;
    TACO_REQUIRE(a.t_begin >= 0 && (a.t_end == 0 || (a.t_begin < a.t_end && a.t_end <= a.Td)), TACO_EINVAL, "attention: bad time chunk [%d, %d)", a.t_begin, a.t_end);
    TACO_REQUIRE(a.t_begin == 0 || (a.s_ha && a.s_ctx && a.s_a), TACO_EINVAL, "attention: a chunk with t_begin > 0 restores its state from the training stash");
    return af_launch(att_fast_fwd_kernel<0>, a, af_fwd_smem(a.Ti), configured, s);
}
int launch_att_fast_bwd(const AttArgs& a, cudaStream_t s) {
    static int configured = 0;
    TACO_REQUIRE(a.t_begin >= 0 && (a.t_end == 0 || (a.t_begin < a.t_end && a.t_end <= a.Td)), TACO_EINVAL, "attention: bad time chunk [%d, %d)", a.t_begin, a.t_end);
    TACO_REQUIRE((a.t_begin == 0 && (a.t_end == 0 || a.t_end == a.Td) && !a.c_in) || (a.c_dha && a.c_dctx && a.c_dac), TACO_EINVAL,
                 "attention: backward time chunks need the carry buffers");
    return af_launch(att_fast_bwd_kernel<0>, a, af_bwd_smem(a.Ti), configured, s);
}

}  // namespace taco

extern "C" int taco_debug_att_prof(long long out[40]) { return taco::af_debug_prof(out); }
