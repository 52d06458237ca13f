// Fast GRU recurrence for the tensor-core precision mode (same contract and stash layout as gru.cu; TF GRUCell
// semantics, reference: models/modules.py:82-96, models/tacotron.py:171-175).
//
// Differences from the exact fp32 kernel in gru.cu:
//   * the recurrent weight slice of each CTA stays resident in shared memory as bf16 for the whole sequence and the
//     per-step products run on mma.sync.m16n8k16 (weight columns = M, the cluster's 8 batch rows = N=8), fp32 accumulate;
//     the hidden state itself is carried in fp32 registers, only the matvec operand is rounded to bf16;
//   * the r*h / h' slices are exchanged with cp.async.bulk shared::cta -> shared::cluster copies that complete on the
//     receiver's mbarrier (one 512-byte copy per peer and phase) instead of two full cluster barriers per step.
// Single receive buffers are sufficient: a CTA can only send phase p+2 data after it has received every peer's phase p+1
// data, which each peer sends after its last read of the phase-p buffer ("exchange protocol").
#include "common.cuh"
#include "kernels.h"
#include <cooperative_groups.h>
#include <cuda_bf16.h>

namespace cg = cooperative_groups;

namespace taco {

constexpr int GF_C = 8, GF_R = 8, GF_NT = 256;

__device__ __forceinline__ uint32_t gf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gf_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gf_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gf_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gf_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(gf_smem_u32(bar)), "r"(parity) : "memory");
}
// bulk copy local shared -> peer CTA's shared (same offsets), completing `bytes` on the peer's mbarrier
__device__ __forceinline__ void gf_push(const void* src_local, void* dst_local_alias, uint64_t* bar_local_alias, uint32_t bytes, uint32_t peer) {
    uint32_t dst, bar;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(gf_smem_u32(dst_local_alias)), "r"(peer));
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(gf_smem_u32(bar_local_alias)), "r"(peer));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(gf_smem_u32(src_local)), "r"(bytes), "r"(bar) : "memory");
}
// Exchange variant 2: every thread forwards 16 bytes of the staged block to one peer with st.async (data + tx-count in
// one DSMEM transaction, no async-proxy fence needed).  NQ = 16-byte chunks per block; threads [0, C*NQ) participate.
// The staged block is dense [R][U]; the receive buffers keep rows U + 8 elements apart (QR = 16-byte chunks per dense row,
// ROWB = bytes per padded row): see GfCfg::UP.
template <int NQ, int QR, int ROWB>
__device__ __forceinline__ void gf_push_stasync(const void* stage, void* dst_local_alias, uint64_t* bar_local_alias, int tid) {
    if (tid < GF_C * NQ) {
        const uint32_t peer = tid / NQ, q = tid % NQ;
        const uint4 v = *(reinterpret_cast<const uint4*>(stage) + q);
        uint32_t dst, bar;
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(gf_smem_u32(dst_local_alias) + (q / QR) * ROWB + (q % QR) * 16), "r"(peer));
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(gf_smem_u32(bar_local_alias)), "r"(peer));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
    }
}
#ifndef GF_USE_STASYNC
#define GF_USE_STASYNC 1
#endif

__device__ __forceinline__ void gf_ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void gf_mma(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// The weight slice never changes during the sequence: its mma A fragments are loaded into registers ONCE.
template <int NK>
__device__ __forceinline__ void gf_load_afrags(uint32_t (&af)[NK][4], const __nv_bfloat16* Wt, int KP, int m0, int k0, int lane) {
    const uint32_t a_base = gf_smem_u32(Wt + (size_t)(m0 + (lane & 15)) * KP + k0 + (lane >> 4) * 8);
#pragma unroll
    for (int kk = 0; kk < NK; kk++) gf_ldmatrix_x4(a_base + (uint32_t)kk * 32, af[kk][0], af[kk][1], af[kk][2], af[kk][3]);
}
// acc = sum over NK k-tiles; B fragments are fetched up front, two independent accumulator chains
template <int U, int NK>
__device__ __forceinline__ void gf_mma_regs(float acc[4], const uint32_t (&af)[NK][4], const __nv_bfloat16* v, int k0, int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int UP = U + 8;      // padded row pitch of the exchanged blocks (GfCfg::UP): the 8 rows g of a fragment load hit 8 different bank groups
    uint32_t b[NK][2];
#pragma unroll
    for (int kk = 0; kk < NK; kk++) {
        const int ka = k0 + kk * 16 + 2 * t, kb = ka + 8;
        b[kk][0] = *reinterpret_cast<const uint32_t*>(v + ((ka / U) * GF_R + g) * UP + (ka % U));
        b[kk][1] = *reinterpret_cast<const uint32_t*>(v + ((kb / U) * GF_R + g) * UP + (kb % U));
    }
    // four independent accumulator chains: a dependent mma.sync costs ~100 clk on sm_100a (gate phase, NK = 8: 569 clk with
    // two chains of four - tools/gru_phase_prof.py), so the chain depth, not the issue rate, sets the phase time
    float c[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++) { c[q][0] = 0.f; c[q][1] = 0.f; c[q][2] = 0.f; c[q][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < NK; kk++) gf_mma(c[kk & 3], af[kk][0], af[kk][1], af[kk][2], af[kk][3], b[kk][0], b[kk][1]);
#pragma unroll
    for (int e = 0; e < 4; e++) acc[e] = (c[0][e] + c[1][e]) + (c[2][e] + c[3][e]);
}
__device__ __forceinline__ float gf_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gf_tanh(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// acc (16 x 8 fp32 fragment) += Wt[m0..m0+15][k0 .. k0+16*nk) . v[k][n]     Wt: bf16 [rows][KP]; v: bf16 [C][R][U] blocked
template <int U, int KP>
__device__ __forceinline__ void gf_mma_slice(float acc[4], const __nv_bfloat16* Wt, int m0, const __nv_bfloat16* v, int k0, int nk, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t a_base = gf_smem_u32(Wt + (size_t)(m0 + (lane & 15)) * KP + (lane >> 4) * 8);
#pragma unroll 4
    for (int kk = 0; kk < nk; kk++) {
        const int k = k0 + kk * 16;
        uint32_t a0, a1, a2, a3;
        gf_ldmatrix_x4(a_base + (uint32_t)k * 2, a0, a1, a2, a3);
        const int ka = k + 2 * t, kb = ka + 8;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(v + ((ka / U) * GF_R + g) * U + (ka % U));
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(v + ((kb / U) * GF_R + g) * U + (kb % U));
        gf_mma(acc, a0, a1, a2, a3, b0, b1);
    }
}

template <int H>
struct GfCfg {
    static constexpr int U = H / GF_C, GC = 2 * U, KP = H + 8, KP2 = 2 * H + 8;
    static constexpr int ACT = U * GF_R;
    static constexpr int BLK_BYTES = GF_R * U * 2;                 // one CTA's bf16 slice of a vector (dense, as staged and sent)
    // Receive buffers [C][R][UP]: rows U + 8 elements apart.  With dense rows (64 B at U = 32) the 8 batch rows of an mma B-fragment
    // load fall on 2 bank groups - a 4-way conflict on every one of the 16 loads of the gate phase, ~500 LSU cycles per step over
    // the 8 warps (tools/gru_phase_prof.py: 570 clk for 8 mma + their operand loads); 80 B (48 B at U = 16) rows are conflict-free.
    static constexpr int UP = U + 8;
    static constexpr int VEC = GF_C * GF_R * UP;                   // elements of one received vector
    // k-split partial sums: [ks][R][cols + 4] floats (row = batch row) so that the activation threads (consecutive units of one
    // batch row) read consecutive words; [ks][col][R] made those reads 8-way conflicted
    static constexpr int RED_FLOATS = 1280;
    // forward: Wg [GC][KP] + Wc [U][KP]; backward: WcT [U][KP] + WgT [U][KP2]
    static constexpr size_t w_bytes = (size_t)3 * U * KP2 * 2;     // upper bound for both directions
    // receive vectors: every exchanged vector exists twice (even / odd steps, see the kernels): forward 2 x {h, r*h}, backward
    // 2 x {dc_pre, [dr_pre ; du_pre]} = 6 x VEC
    static constexpr size_t smem_bytes = w_bytes + (size_t)6 * VEC * 2 /*recv vectors*/ + 3 * BLK_BYTES /*stages*/ +
                                         RED_FLOATS * 4 /*red*/ + 64 /*barriers*/ + 256;
};

__device__ long long g_gf_prof[16];
#define GF_T(i) do { if (prof) { long long _c = clock64(); pacc[i] += _c - plast; plast = _c; } } while (0)

template <int H>
__global__ void __launch_bounds__(GF_NT, 1) gru_fast_fwd_kernel(const GruArgs a) {
    using Cfg = GfCfg<H>;
    constexpr int U = Cfg::U, GC = Cfg::GC, KP = Cfg::KP, R = GF_R, UP = Cfg::UP;
    constexpr int RPG = GC + 4, RPC = U + 4;                      // row pitches of the partial sums (gate / candidate phase)
    constexpr int QR = U / 8, ROWB = UP * 2;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / GF_C;
    const int d = cid % a.ndir, grp = cid / a.ndir;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(128) uint8_t smem_raw[];
    __nv_bfloat16* Wg_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);            // [GC][KP]
    __nv_bfloat16* Wc_s = Wg_s + GC * KP;                                         // [U][KP]
    // Exchange buffers and barriers alternate with the step parity: bytes of step s+2 are the earliest that can meet the barrier
    // of step s, and a peer can only send those after it has my step-(s+1) data, which I send after my step-s waits - so the
    // protocol holds under ANY delivery order of the remote stores (with one buffer per vector it assumed that a store is never
    // overtaken by a whole exchange phase).
    __nv_bfloat16* hb_s = reinterpret_cast<__nv_bfloat16*>(smem_raw + Cfg::w_bytes);   // [2][C][R][UP]
    __nv_bfloat16* rhb_s = hb_s + 2 * Cfg::VEC;                                   // [2][C][R][UP]
    __nv_bfloat16* stage_rh = rhb_s + 2 * Cfg::VEC;                               // [R][U]
    __nv_bfloat16* stage_h = stage_rh + R * U;
    float* red = reinterpret_cast<float*>(stage_h + 2 * R * U);                   // RED_FLOATS
    uint64_t* bar_rh = reinterpret_cast<uint64_t*>(red + Cfg::RED_FLOATS);      // [2]
    uint64_t* bar_h = bar_rh + 2;                                                 // [2]

    const float* __restrict__ Wg = a.Wg[d];
    const float* __restrict__ Wc = a.Wc[d];
    // (loads issued in batches of 8 before the first use: this set-up is paid once per time chunk under the decoder wavefront)
    for (int base = tid; base < GC * H; base += GF_NT * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = base + u * GF_NT, k = idx / GC, col = idx % GC;
            const int gcol = (col < U) ? rank * U + col : H + rank * U + (col - U);
            v[u] = idx < GC * H ? __ldg(Wg + (long long)k * 2 * H + gcol) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; if (idx < GC * H) Wg_s[(idx % GC) * KP + idx / GC] = __float2bfloat16(v[u]); }
    }
    for (int base = tid; base < U * H; base += GF_NT * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; v[u] = idx < U * H ? __ldg(Wc + (long long)(idx / U) * H + rank * U + idx % U) : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; if (idx < U * H) Wc_s[(idx % U) * KP + idx / U] = __float2bfloat16(v[u]); }
    }
    for (int idx = tid; idx < R * H; idx += GF_NT) {       // hb_s[(k/U)][r][k%U]
        const int blk = idx / (R * U), r = (idx / U) % R, i = idx % U, n = grp * R + r;
        const int k = blk * U + i;
        hb_s[Cfg::VEC + (blk * R + r) * UP + i] = __float2bfloat16(      // the first step (parity 0) reads buffer 1
            (a.h0 && n < a.N) ? a.h0[(long long)n * a.ndir * H + d * H + k] : 0.f);
    }
    if (tid == 0) {
        gf_mbar_init(bar_rh, 1); gf_mbar_init(bar_rh + 1, 1); gf_mbar_init(bar_h, 1); gf_mbar_init(bar_h + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool act = tid < Cfg::ACT;
    const int i = tid % U, r = tid / U;
    const int n = grp * R + r;
    int L = 0;
    if (act && n < a.N) L = a.lengths ? min(max(a.lengths[n], 0), a.T) : a.T;
    int Lmax = 0;
    for (int rr = 0; rr < R; rr++) {
        const int nn = grp * R + rr;
        if (nn < a.N) Lmax = max(Lmax, a.lengths ? min(max(a.lengths[nn], 0), a.T) : a.T);
    }
    const int unit = rank * U + i;
    float h_own = (act && a.h0 && n < a.N) ? a.h0[(long long)n * a.ndir * H + d * H + unit] : 0.f;
    const long long st_base = ((long long)d * a.N + n) * a.T;
    __syncthreads();
    cluster.sync();

    // mma work split: gate phase 4 m-tiles (GC=64) x 2 k-halves (H=256) | for H=128: GC=32 -> 2 m-tiles x 4 k-quarters
    constexpr int MT_G = GC / 16, KS_G = 8 / MT_G, NK_G = H / 16 / KS_G;
    constexpr int MT_C = U / 16, KS_C = 8 / MT_C, NK_C = H / 16 / KS_C;
    const int g4 = lane >> 2, t4 = lane & 3;
    uint32_t afG[NK_G][4], afC[NK_C][4];
    gf_load_afrags<NK_G>(afG, Wg_s, KP, (warp % MT_G) * 16, (warp / MT_G) * NK_G * 16, lane);
    gf_load_afrags<NK_C>(afC, Wc_s, KP, (warp % MT_C) * 16, (warp / MT_C) * NK_C * 16, lane);

    const bool prof = (blockIdx.x == 0 && tid == 0);
    long long pacc[10] = {0,0,0,0,0,0,0,0,0,0}; long long plast = clock64();
    float gr = 0.f, gu = 0.f, gc = 0.f, gres = 0.f;
    auto load_gx = [&](int s) {
        if (act && s < L) {
            const int t = (d == 0) ? s : (L - 1 - s);
            const float* g = a.gx + ((long long)n * a.gx_rs_n + t + a.gx_row0) * a.gx_ld + (long long)d * 3 * H + unit;
            gr = __ldg(g); gu = __ldg(g + H); gc = __ldg(g + 2 * H);
            if (a.res) gres = __ldg(a.res + ((long long)n * a.T + t) * a.res_ld + unit);
        }
    };
    const int s_begin = a.t_begin, s_end = a.t_end > 0 ? min(a.t_end, Lmax) : Lmax;      // chunked: steps [t_begin, t_end)
    load_gx(s_begin);
    for (int s = s_begin; s < s_end; s++) {
        const uint32_t par = (s - s_begin) & 1, ph = ((s - s_begin) >> 1) & 1;      // buffer / barrier of this step, and its phase
        __nv_bfloat16* h_in = hb_s + (par ^ 1) * Cfg::VEC;                            // h of the previous step
        __nv_bfloat16* h_out = hb_s + par * Cfg::VEC;
        __nv_bfloat16* rh_io = rhb_s + par * Cfg::VEC;
        const bool valid = act && (s < L);
        const int t = (d == 0) ? s : (L - 1 - s);
        const float cgr = gr, cgu = gu, cgc = gc, cres = gres;
        if (tid == 0) { gf_mbar_expect_tx(bar_rh + par, GF_C * Cfg::BLK_BYTES); gf_mbar_expect_tx(bar_h + par, GF_C * Cfg::BLK_BYTES); }
        GF_T(0);
        // ---- gate phase ----
        {
            float acc[4];
            const int mt = warp % MT_G, ks = warp / MT_G;
            gf_mma_regs<U, NK_G>(acc, afG, h_in, ks * NK_G * 16, lane);
            float* rp = red + ks * (R * RPG) + mt * 16 + g4;
            rp[(2 * t4) * RPG] = acc[0]; rp[(2 * t4 + 1) * RPG] = acc[1];
            rp[(2 * t4) * RPG + 8] = acc[2]; rp[(2 * t4 + 1) * RPG + 8] = acc[3];
        }
        GF_T(1);
        __syncthreads();
        GF_T(2);
        float rg = 0.f, ug = 0.f;
        if (act) {
            float sr = cgr, su = cgu;
#pragma unroll
            for (int ks = 0; ks < KS_G; ks++) { sr += red[ks * (R * RPG) + r * RPG + i]; su += red[ks * (R * RPG) + r * RPG + U + i]; }
            rg = gf_sigmoid(sr); ug = gf_sigmoid(su);
            stage_rh[r * U + i] = __float2bfloat16(rg * h_own);
        }
        GF_T(3);
#if GF_USE_STASYNC
        __syncthreads();
        GF_T(4);
        gf_push_stasync<Cfg::BLK_BYTES / 16, QR, ROWB>(stage_rh, rh_io + rank * R * UP, bar_rh + par, tid);
        load_gx(s + 1);      // next step's x-side pre-activations: issued in the shadow of the exchange (at the top of the step the
                             // address arithmetic and four load issues cost ~200 clk on the serial path)
        GF_T(5);
#else
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < GF_C) gf_push(stage_rh, rhb_s + rank * R * U, bar_rh, Cfg::BLK_BYTES, tid);
#endif
        gf_mbar_wait(bar_rh + par, ph);
        GF_T(6);
        // ---- candidate phase ----
        {
            float acc[4];
            const int mt = warp % MT_C, ks = warp / MT_C;
            gf_mma_regs<U, NK_C>(acc, afC, rh_io, ks * NK_C * 16, lane);
            float* rp = red + ks * (R * RPC) + mt * 16 + g4;
            rp[(2 * t4) * RPC] = acc[0]; rp[(2 * t4 + 1) * RPC] = acc[1];
            rp[(2 * t4) * RPC + 8] = acc[2]; rp[(2 * t4 + 1) * RPC + 8] = acc[3];
        }
        __syncthreads();
        float cnd = 0.f, hprev = h_own;
        if (act) {
            float sc = cgc;
#pragma unroll
            for (int ks = 0; ks < KS_C; ks++) sc += red[ks * (R * RPC) + r * RPC + i];
            cnd = gf_tanh(sc);
            if (valid) h_own = ug * h_own + (1.f - ug) * cnd;
            stage_h[r * U + i] = __float2bfloat16(h_own);
        }
#if GF_USE_STASYNC
        __syncthreads();
        gf_push_stasync<Cfg::BLK_BYTES / 16, QR, ROWB>(stage_h, h_out + rank * R * UP, bar_h + par, tid);
        if (valid) {      // output and stash stores ride in the shadow of the exchange
            const long long o = (long long)n * a.T + t;
            a.out[o * a.out_ld + d * H + unit] = h_own + cres;
            if (a.out16) static_cast<__nv_bfloat16*>(a.out16)[o * a.out_ld + d * H + unit] = __float2bfloat16_rn(h_own + cres);
            if (a.st_r) {
                const long long so = (st_base + t) * H + unit;
                a.st_r[so] = rg; a.st_u[so] = ug; a.st_c[so] = cnd; a.st_hprev[so] = hprev;
                if (a.st_hprev16) static_cast<__nv_bfloat16*>(a.st_hprev16)[so] = __float2bfloat16_rn(hprev);
            }
        }
#else
#error "bulk-copy exchange variant is not maintained"
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < GF_C) gf_push(stage_h, hb_s + rank * R * U, bar_h, Cfg::BLK_BYTES, tid);
#endif
        gf_mbar_wait(bar_h + par, ph);
        GF_T(7);
    }
    if (prof) { for (int q = 0; q < 10; q++) g_gf_prof[q] = pacc[q]; g_gf_prof[10] = Lmax; }
    if (act && a.hfinal && n < a.N) a.hfinal[(long long)n * a.ndir * H + d * H + unit] = h_own;
    cluster.sync();      // no CTA may exit while a peer's copy into it could still be in flight
}

// BPTT (same maths as gru_bwd_kernel in gru.cu).
template <int H>
__global__ void __launch_bounds__(GF_NT, 1) gru_fast_bwd_kernel(const GruArgs a) {
    using Cfg = GfCfg<H>;
    constexpr int U = Cfg::U, KP = Cfg::KP, KP2 = Cfg::KP2, R = GF_R, UP = Cfg::UP;
    constexpr int RP = U + 4;                                      // row pitch of the partial sums
    constexpr int QR = U / 8, ROWB = UP * 2;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / GF_C;
    const int d = cid % a.ndir, grp = cid / a.ndir;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    extern __shared__ __align__(128) uint8_t smem_raw[];
    __nv_bfloat16* WcT_s = reinterpret_cast<__nv_bfloat16*>(smem_raw);           // [U][KP]   WcT_s[i][cu] = Wc[unit_i][cu]
    __nv_bfloat16* WgT_s = WcT_s + U * KP;                                        // [U][KP2]  WgT_s[i][gc] = Wg[unit_i][gc]
    __nv_bfloat16* dcp_s = reinterpret_cast<__nv_bfloat16*>(smem_raw + Cfg::w_bytes);  // [C][R][U]       dc_pre, all units
    __nv_bfloat16* dg_s = dcp_s + 2 * Cfg::VEC;                                   // [2][2C][R][UP]  [dr_pre ; du_pre]   (both alternate with the
    __nv_bfloat16* stage_c = dg_s + 4 * Cfg::VEC;                                 // [R][U]            iteration parity, as in the forward kernel)
    __nv_bfloat16* stage_r = stage_c + R * U;
    __nv_bfloat16* stage_u = stage_r + R * U;
    float* red = reinterpret_cast<float*>(stage_u + R * U);                       // RED_FLOATS
    uint64_t* bar_c = reinterpret_cast<uint64_t*>(red + Cfg::RED_FLOATS);       // [2]
    uint64_t* bar_g = bar_c + 2;                                                  // [2]

    const float* __restrict__ Wg = a.Wg[d];
    const float* __restrict__ Wc = a.Wc[d];
    for (int base = tid; base < U * H; base += GF_NT * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; v[u] = idx < U * H ? __ldg(Wc + (long long)(rank * U + idx / H) * H + idx % H) : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; if (idx < U * H) WcT_s[(idx / H) * KP + idx % H] = __float2bfloat16(v[u]); }
    }
    for (int base = tid; base < U * 2 * H; base += GF_NT * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; v[u] = idx < U * 2 * H ? __ldg(Wg + (long long)(rank * U + idx / (2 * H)) * 2 * H + idx % (2 * H)) : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; u++) { const int idx = base + u * GF_NT; if (idx < U * 2 * H) WgT_s[(idx / (2 * H)) * KP2 + idx % (2 * H)] = __float2bfloat16(v[u]); }
    }
    if (tid == 0) {
        gf_mbar_init(bar_c, 1); gf_mbar_init(bar_c + 1, 1); gf_mbar_init(bar_g, 1); gf_mbar_init(bar_g + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool act = tid < Cfg::ACT;
    const int i = tid % U, r = tid / U;
    const int n = grp * R + r;
    int L = 0;
    if (act && n < a.N) L = a.lengths ? min(max(a.lengths[n], 0), a.T) : a.T;
    int Lmax = 0;
    for (int rr = 0; rr < R; rr++) {
        const int nn = grp * R + rr;
        if (nn < a.N) Lmax = max(Lmax, a.lengths ? min(max(a.lengths[nn], 0), a.T) : a.T);
    }
    const int unit = rank * U + i;
    const long long st_base = ((long long)d * a.N + n) * a.T;
    float dh_carry = (act && a.dh_in && n < a.N) ? a.dh_in[(long long)n * a.ndir * H + d * H + unit] : 0.f;
    float sb_r = 0.f, sb_u = 0.f, sb_c = 0.f;        // bias gradients: this thread's (unit, row) sums of the gate gradients over time
    __syncthreads();
    cluster.sync();

    constexpr int MT = U / 16, KS = 8 / MT;          // output tiles (own units) and k-splits over the 8 warps
    constexpr int NK_C = H / 16 / KS, NK_G = 2 * H / 16 / KS;
    const int g4 = lane >> 2, t4 = lane & 3;
    uint32_t afC[NK_C][4], afG[NK_G][4];
    gf_load_afrags<NK_C>(afC, WcT_s, KP, (warp % MT) * 16, (warp / MT) * NK_C * 16, lane);
    gf_load_afrags<NK_G>(afG, WgT_s, KP2, (warp % MT) * 16, (warp / MT) * NK_G * 16, lane);

    float rg = 0.f, ug = 0.f, cc = 0.f, hp = 0.f, dout = 0.f;
    auto load_step = [&](int s) {
        if (act && s >= 0 && s < L) {
            const int t = (d == 0) ? s : (L - 1 - s);
            const long long so = (st_base + t) * H + unit;
            rg = a.st_r[so]; ug = a.st_u[so]; cc = a.st_c[so]; hp = a.st_hprev[so];
            dout = a.dout[((long long)n * a.T + t) * a.dout_ld + d * H + unit];
        }
    };
    const int s_begin = a.t_begin, s_end = a.t_end > 0 ? min(a.t_end, Lmax) : Lmax;      // chunked: steps [t_begin, t_end)
    load_step(s_end - 1);
    int it = 0;
    for (int s = s_end - 1; s >= s_begin; s--, it++) {
        const uint32_t par = it & 1, ph = (it >> 1) & 1;
        __nv_bfloat16* dcp_io = dcp_s + par * Cfg::VEC;
        __nv_bfloat16* dg_io = dg_s + par * 2 * Cfg::VEC;
        const bool valid = act && (s < L);
        const int t = (d == 0) ? s : (L - 1 - s);
        const float r_ = rg, u_ = ug, c_ = cc, hp_ = hp;
        float dh = dh_carry + (valid ? dout : 0.f);
        if (tid == 0) { gf_mbar_expect_tx(bar_c + par, GF_C * Cfg::BLK_BYTES); gf_mbar_expect_tx(bar_g + par, 2 * GF_C * Cfg::BLK_BYTES); }
        float du_pre = 0.f, dc_pre = 0.f;
        if (valid) {
            du_pre = dh * (hp_ - c_) * u_ * (1.f - u_);
            dc_pre = dh * (1.f - u_) * (1.f - c_ * c_);
        }
        if (act) stage_c[r * U + i] = __float2bfloat16(dc_pre);
#if GF_USE_STASYNC
        __syncthreads();
        gf_push_stasync<Cfg::BLK_BYTES / 16, QR, ROWB>(stage_c, dcp_io + rank * R * UP, bar_c + par, tid);
        if (s - 1 >= s_begin) load_step(s - 1);      // next step's stash values: five loads issued in the shadow of the exchange
#else
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < GF_C) gf_push(stage_c, dcp_s + rank * R * U, bar_c, Cfg::BLK_BYTES, tid);
#endif
        gf_mbar_wait(bar_c + par, ph);
        // d(r*h)[own units] = sum_cu dc_pre[cu] * Wc[unit][cu]
        {
            float acc[4];
            const int mt = warp % MT, ks = warp / MT;
            gf_mma_regs<U, NK_C>(acc, afC, dcp_io, ks * NK_C * 16, lane);
            float* rp = red + ks * (R * RP) + mt * 16 + g4;
            rp[(2 * t4) * RP] = acc[0]; rp[(2 * t4 + 1) * RP] = acc[1];
            rp[(2 * t4) * RP + 8] = acc[2]; rp[(2 * t4 + 1) * RP + 8] = acc[3];
        }
        __syncthreads();
        float d_rh = 0.f, dr_pre = 0.f;
        if (act) {
#pragma unroll
            for (int ks = 0; ks < KS; ks++) d_rh += red[ks * (R * RP) + r * RP + i];
            if (valid) dr_pre = d_rh * hp_ * r_ * (1.f - r_);
            stage_r[r * U + i] = __float2bfloat16(dr_pre);
            stage_u[r * U + i] = __float2bfloat16(du_pre);
        }
#if GF_USE_STASYNC
        __syncthreads();
        gf_push_stasync<Cfg::BLK_BYTES / 16, QR, ROWB>(stage_r, dg_io + rank * R * UP, bar_g + par, tid);
        gf_push_stasync<Cfg::BLK_BYTES / 16, QR, ROWB>(stage_u, dg_io + (GF_C + rank) * R * UP, bar_g + par, tid);
        if (valid) {      // gradient / stash stores in the shadow of the exchange
            sb_r += dr_pre; sb_u += du_pre; sb_c += dc_pre;
            const long long go = ((long long)n * a.gx_rs_n + t + a.gx_row0) * a.gx_ld + (long long)d * 3 * H + unit;
            if (a.dgx) { float* g = a.dgx + go; g[0] = dr_pre; g[H] = du_pre; g[2 * H] = dc_pre; }
            const __nv_bfloat16 br = __float2bfloat16_rn(dr_pre), bu = __float2bfloat16_rn(du_pre), bc = __float2bfloat16_rn(dc_pre);
            if (a.dgx16) { __nv_bfloat16* g = static_cast<__nv_bfloat16*>(a.dgx16) + go; g[0] = br; g[H] = bu; g[2 * H] = bc; }
            if (a.dgx16_dense) {
                __nv_bfloat16* g = static_cast<__nv_bfloat16*>(a.dgx16_dense) + (st_base + t) * 3 * H + unit;
                g[0] = br; g[H] = bu; g[2 * H] = bc;
            }
            a.st_r[(st_base + t) * H + unit] = r_ * hp_;
            if (a.st_rh16) static_cast<__nv_bfloat16*>(a.st_rh16)[(st_base + t) * H + unit] = __float2bfloat16_rn(r_ * hp_);
        }
#else
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < GF_C) gf_push(stage_r, dg_s + rank * R * U, bar_g, Cfg::BLK_BYTES, tid);
        else if (tid < 2 * GF_C) gf_push(stage_u, dg_s + (GF_C + rank) * R * U, bar_g, Cfg::BLK_BYTES, tid - GF_C);
#endif
        gf_mbar_wait(bar_g + par, ph);
        // dh_prev += sum_gc [dr_pre;du_pre][gc] * Wg[unit][gc]      (dg_s is blocked [2C][R][U]: k = gc, K = 2H)
        {
            float acc[4];
            const int mt = warp % MT, ks = warp / MT;
            gf_mma_regs<U, NK_G>(acc, afG, dg_io, ks * NK_G * 16, lane);
            float* rp = red + ks * (R * RP) + mt * 16 + g4;
            rp[(2 * t4) * RP] = acc[0]; rp[(2 * t4 + 1) * RP] = acc[1];
            rp[(2 * t4) * RP + 8] = acc[2]; rp[(2 * t4 + 1) * RP + 8] = acc[3];
        }
        __syncthreads();
        if (act) {
            float sg = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) sg += red[ks * (R * RP) + r * RP + i];
            if (valid) dh_carry = dh * u_ + d_rh * r_ + sg;
        }
        __syncthreads();      // `red` / stages are rewritten at the top of the next iteration
    }
    if (act && a.dh0 && n < a.N) a.dh0[(long long)n * a.ndir * H + d * H + unit] = dh_carry;
    if (act && a.dbg[d] && n < a.N) {                // gate / candidate bias gradients (8 rows per unit and cluster: a handful of atomics)
        atomicAdd(a.dbg[d] + unit, sb_r); atomicAdd(a.dbg[d] + H + unit, sb_u); atomicAdd(a.dbc[d] + unit, sb_c);
    }
    cluster.sync();
}

template <int H>
static int launch_gru_fast_t(const GruArgs& a, bool bwd, cudaStream_t s) {
    using Cfg = GfCfg<H>;
    const size_t smem = Cfg::smem_bytes;
    auto kern = bwd ? gru_fast_bwd_kernel<H> : gru_fast_fwd_kernel<H>;
    TACO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(GF_C * a.ndir * cdiv(a.N, GF_R));
    cfg.blockDim = dim3(GF_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GF_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TACO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    g_launch_count++;
    return TACO_OK;
}

int launch_gru_fast_fwd(const GruArgs& a, cudaStream_t s) {
    TACO_REQUIRE(a.H == 128 || a.H == 256, TACO_ESHAPE, "gru: hidden size %d not instantiated (128, 256)", a.H);
    TACO_REQUIRE((a.t_begin == 0 && a.t_end == 0) || (a.ndir == 1 && !a.lengths && a.t_begin >= 0 && a.t_begin < a.t_end && a.t_end <= a.T),
                 TACO_EINVAL, "gru: time chunks need ndir == 1, no lengths and 0 <= t_begin < t_end <= T");
    return a.H == 128 ? launch_gru_fast_t<128>(a, false, s) : launch_gru_fast_t<256>(a, false, s);
}
int launch_gru_fast_bwd(const GruArgs& a, cudaStream_t s) {
    TACO_REQUIRE(a.H == 128 || a.H == 256, TACO_ESHAPE, "gru: hidden size %d not instantiated (128, 256)", a.H);
    TACO_REQUIRE((a.t_begin == 0 && a.t_end == 0) || (a.ndir == 1 && !a.lengths && a.t_begin >= 0 && a.t_begin < a.t_end && a.t_end <= a.T),
                 TACO_EINVAL, "gru: time chunks need ndir == 1, no lengths and 0 <= t_begin < t_end <= T");
    return a.H == 128 ? launch_gru_fast_t<128>(a, true, s) : launch_gru_fast_t<256>(a, true, s);
}

int gf_debug_prof(long long out[16]) { TACO_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_gf_prof, sizeof(long long) * 16)); return TACO_OK; }

}  // namespace taco

extern "C" int taco_debug_gru_prof(long long out[16]) { return taco::gf_debug_prof(out); }
