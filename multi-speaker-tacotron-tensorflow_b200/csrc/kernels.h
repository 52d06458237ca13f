// Internal launcher prototypes (host side).  All launchers enqueue on `s` and return TACO_* codes.
#pragma once
#include <cuda_runtime.h>
#include "../../include/taco_capi.h"

namespace taco {

// gemm_simt.cu / gemm_tc.cu
int launch_gemm_simt(const taco_gemm_desc* d, int n, cudaStream_t s);
int launch_gemm(const taco_gemm_desc* d, int n_problems, int precision, cudaStream_t s);
bool gemm_tc_eligible(const taco_gemm_desc& g);
int launch_gemm_tc(const taco_gemm_desc& g, cudaStream_t s);   // TACO_ENOTSUP: caller falls back to SIMT
constexpr int TACO_ENOTSUP = -100;
bool gemm_bf16_eligible(const taco_gemm_desc& g);              // gemm_bf16.cu: both bf16 mirrors given and TMA-addressable
int launch_gemm_bf16(const taco_gemm_desc& g, cudaStream_t s, bool leaf = false);   // leaf: short one-unit CTAs instead of persistent ones

// elementwise.cu
int launch_fill(float* p, long long n, float v, cudaStream_t s);
// (trailing `void* ...16` arguments: optional bf16 mirror of the output, see elementwise.cu)
int launch_gather_rows(const float* table, const int* idx, float* out, int N, int T, int Tp, int PL, int C, int n_rows_table, cudaStream_t s, void* out16 = nullptr);
int launch_scatter_add_rows(const float* dx, const int* idx, float* dtable, int N, int T, int Tp, int PL, int C, int n_rows_table, cudaStream_t s);
int launch_bn_finalize(const double* sum, const double* sumsq, double count, float* mean, float* rstd, float* var,
                       const float* moving_mean, const float* moving_var, int C, int training, cudaStream_t s);
int launch_bn_update_moving(float* moving_mean, float* moving_var, const float* mean, const float* var, int C, cudaStream_t s);
int launch_unpad(float* dst, const float* src, int N, int T, int Tp, int PL, int C, long long ld, cudaStream_t s);
int launch_pack_dgrad(const float* W, float* Wd, int k, int Cin, int Cout, cudaStream_t s);
int launch_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    const float* res, const float* rowvec, float* out, int N, int T, int Tp, int PL, int C, int mode, cudaStream_t s, void* out16 = nullptr,
                    const void* x16 = nullptr /* x as bf16 instead of fp32 */);
int launch_bn_bwd(const float* dyp, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  float* dgamma, float* dbeta, float* dx, int N, int T, int Tp, int PL, int C, int mode, int relu_mask, cudaStream_t s,
                  void* dx16 = nullptr, float* dbias = nullptr /* += column sums of dx: the convolution's bias gradient */,
                  const void* x16 = nullptr, const void* dyp16 = nullptr /* bf16 inputs instead of the fp32 ones */);
int launch_highway_fwd(const float* H, const float* Tg, const float* x, float* y, long long n, cudaStream_t s, void* y16 = nullptr);
int launch_highway_bwd(const float* dy, const float* H, const float* Tg, const float* x, float* dHT /*[rows,2C]: dHpre | dTpre*/, float* dx,
                       long long rows, int C, cudaStream_t s, void* dHT16 = nullptr, float* dbH = nullptr, float* dbT = nullptr /* += column sums of dHpre / dTpre */);
int launch_softsign_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t s);
int launch_relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t s);
int launch_teacher_inputs(const float* tgt, float* x, int N, int Td, int To, int r, int M, cudaStream_t s);
int launch_colsum(const float* x, float* out, long long M, int C, int ld, cudaStream_t s);
int launch_colstats(const float* x, double* sum, double* sumsq, long long rows, int C, cudaStream_t s);
int launch_colsum16(const void* x_bf16, float* out, long long M, int C, long long ld, cudaStream_t s);
int launch_cast2d_bf16(void* dst_bf16, const float* src, long long rows, int cols, long long ldd, long long lds, cudaStream_t s);
int launch_bcast_rows(const float* src, float* dst, int N, int T, int C, long long ld, cudaStream_t s);
int launch_timesum(const float* x, float* out, int N, int T, int Tp, int PL, int C, cudaStream_t s);
int launch_axpy(float* y, const float* x, float a, long long n, cudaStream_t s);
int launch_copy2d(float* dst, const float* src, long long rows, int cols, long long ldd, long long lds, cudaStream_t s);
// table-driven operand preparation (elementwise.cu: prep_ops_kernel): gather descriptors, one launch for all of them
enum { PREP_PACK_DGRAD = 0, PREP_COPY2D = 1, PREP_TRANSPOSE = 2, PREP_MAX_OPS = 40 };
struct PrepOp { const float* src; float* dst; void* dst16; int kind, a, b, c; long long ldd, lds, total; };
struct PrepTable {
    PrepOp op[PREP_MAX_OPS]; int n = 0;
    bool add(const PrepOp& o) { if (n >= PREP_MAX_OPS) return false; op[n++] = o; return true; }
    // Wd[j'][co][ci] = W[k-1-j'][ci][co]
    bool pack_dgrad(const float* W, float* Wd, void* Wd16, int k, int Cin, int Cout) {
        return add(PrepOp{W, Wd, Wd16, PREP_PACK_DGRAD, k, Cin, Cout, 0, 0, (long long)k * Cin * Cout});
    }
    // dst[r*ldd + c] = src[r*lds + c]
    bool copy2d(float* dst, void* dst16, const float* src, long long rows, int cols, long long ldd, long long lds) {
        return add(PrepOp{src, dst, dst16, PREP_COPY2D, 0, cols, 0, ldd, lds, rows * cols});
    }
    // out[c*rows + r] = in[r*cols + c]
    bool transpose(const float* in, float* out, int rows, int cols) {
        return add(PrepOp{in, out, nullptr, PREP_TRANSPOSE, rows, cols, 0, 0, 0, (long long)rows * cols});
    }
};
int launch_prep_ops(const PrepTable& tb, cudaStream_t s);
int launch_l1_loss(const float* out, long long out_bs, long long out_ts, const float* tgt, const float* coeff,
                   float* grad, long long grad_bs, long long grad_ts, int N, int T, int C,
                   float w_all, float w_band, int lo, int hi, double* scalars, cudaStream_t s, void* grad16 = nullptr,
                   int tgt_is_bf16 = 0 /* tgt points at bf16 values */);

// gru.cu — cluster-persistent GRU recurrences (TF GRUCell semantics)
struct GruArgs {
    int N, T, H, ndir;
    int fast;                // 1: bf16-weight mma + mbarrier-exchange kernels (gru_fast.cu); 0: exact fp32 kernels (gru.cu)
    // x-side pre-activations (biases included) from the hoisted GEMM: row(n,t) = n*gx_rs_n + t + gx_row0;
    // direction d uses columns [d*3H, (d+1)*3H) as r | u | c.
    const float* gx; long long gx_ld; long long gx_rs_n; long long gx_row0;
    const float* Wg[2];      // recurrent gate weights  [H, 2H] row-major (= gates_kernel + Cin*2H)
    const float* Wc[2];      // recurrent cand weights  [H, H]  row-major (= cand_kernel + Cin*H)
    const float* h0;         // [N, ndir*H] or NULL
    const int* lengths;      // [N] or NULL (=> T)
    const float* res; long long res_ld;   // optional residual input added to the output (ResidualWrapper), row n*T+t
    float* out; long long out_ld;         // out[(n*T+t)*out_ld + d*H + unit]   (must be pre-zeroed when lengths != NULL)
    // stash for backward, each [ndir][N][T][H]; NULL in inference
    float* st_r; float* st_u; float* st_c; float* st_hprev;
    // backward only
    const float* dout; long long dout_ld; // grad wrt out, same indexing as out
    float* dgx;                           // grad wrt gx, same indexing as gx (pre-zeroed)
    float* dh0;                           // [N, ndir*H] or NULL
    float* hfinal;                        // forward: final state [N, ndir*H] or NULL
    // time-chunked execution (decoder wavefront, model_decoder.cu; fast kernels only, ndir == 1, no lengths): the launch
    // covers steps [t_begin, t_end) of the T-step sequence (t_end == 0: all of it); all buffers keep their full-T indexing.
    // Forward chunks chain through h0 / hfinal, backward chunks (run last to first) through dh_in (carry in) / dh0 (out).
    int t_begin, t_end;
    const float* dh_in;                   // backward: gradient wrt the state after step t_end-1, [N, ndir*H] or NULL (zero)
    // bf16 mirrors (fast kernels, TACO_PREC_BF16; each optional): forward: out16 indexed like out, st_hprev16 like st_hprev;
    // backward: dgx16 indexed like dgx (dgx itself may then be NULL), dgx16_dense [ndir][N*T][3H] (the valid rows, dense: the
    // operand of the recurrent weight gradients), st_rh16 [ndir][N][T][H] = r * h_prev
    void* out16; void* st_hprev16; void* dgx16; void* dgx16_dense; void* st_rh16;
    // backward, fast kernels: bias gradients accumulated in the kernel (+= over the launch's steps): gates_bias [2H] and cand_bias [H] per direction, or NULL
    float* dbg[2]; float* dbc[2];
};
int launch_gru_fwd(const GruArgs& a, cudaStream_t s);
int launch_gru_bwd(const GruArgs& a, cudaStream_t s);
int launch_gru_fast_fwd(const GruArgs& a, cudaStream_t s);
int launch_gru_fast_bwd(const GruArgs& a, cudaStream_t s);

// attention.cu — cluster-persistent attention recurrence of the decoder
struct AttArgs {
    int N, Ti, Td;
    int E, A, HA, Z1, Z, SPK, Y;       // memory width, attention size, attention-GRU size, prenet sizes, speaker width, dec rnn size
    int att_type, fast;
    const float* px;        // [N*Td, Z1] x-side of prenet layer 1 (bias included)
    const float* memory;    // [N*Ti, E]
    const float* keys;      // [N*Ti, A]
    const float* spk;       // [N, SPK] or NULL
    const float* ha0;       // [N, HA] or NULL
    const float* manual;    // [N, Td, Ti] or NULL
    const float *W1c, *W2, *b2, *Wg, *bg, *Wc, *bc, *Wq, *v, *score_bias, *att_g, *att_b, *Wo, *bo;
    float* y0;              // [N*Td, Y]
    float* align;           // [N, Ti, Td]
    float* ha_final;        // [N, HA] or NULL
    // stash (training)
    float *s_z1, *s_z, *s_r, *s_u, *s_c, *s_haprev, *s_ha, *s_q, *s_ctxin, *s_ctx, *s_e, *s_a;
    // free-running decoding (inference / rnn_decoder_test_mode; helpers.py:26-32,63-64): the step input is the last of the r
    // frames the previous step produced, so prenet x-part, both residual GRUs and the mel projection run inside the loop
    int free_run, M, r;
    const float *W1x, *b1;                                   // dense_1 rows [0,M) and bias
    const float *Wg1, *bg1, *Wc1, *bc1, *Wg2, *bg2, *Wc2, *bc2;  // full TF kernels of the two decoder GRUs
    const float *Wmel, *bmel;                                // [Y, M*r], [M*r]
    const float *h1_0, *h2_0;                                // [N, Y] or NULL
    float* mel_out; long long mel_bs;                        // mel_out[n*mel_bs + t*M*r + c]
    // backward
    const float* dy0;       // [N*Td, Y]
    const float *W1cT, *W2T, *WgT, *WcT, *WqT, *WoT;   // transposed weights
    float *d_G, *d_zp, *d_z1p, *d_ctx, *d_gq, *d_ge, *d_ha0, *d_score_bias;
    // time-chunked execution (fast kernels, training): steps [t_begin, t_end) (t_end == 0: all).  A forward chunk with
    // t_begin > 0 restores its state (ha, context, previous alignments) from the stash rows of step t_begin-1; backward
    // chunks run last to first and hand the carried gradients (wrt ha, context, alignments) over in c_dha [N,HA], c_dctx
    // [N,E], c_dac [N, 16*ceil(Ti/16)]: every launch writes them at its end, a launch with c_in != 0 reads them first.
    int t_begin, t_end, c_in;
    float *c_dha, *c_dctx, *c_dac;
    // pre-built operand images of the fast kernels (att_fast.cu: launch_att_fast_pack), or NULL: weight slices per cluster
    // rank and key / memory slices per CTA, in the kernels' shared-memory layout, pulled in with bulk async copies
    const uint8_t *img_w, *img_km;
    // free-running decoder of the tensor-core modes: the step's weights as bf16 mma fragments (attention.cu: launch_att_wfrag_pack)
    const void* wfrag; long long wf_off[18];
};
int launch_att_fwd(const AttArgs& a, cudaStream_t s);
bool att_wfrag_supported(const AttArgs& a);
size_t att_wfrag_bytes(const AttArgs& a);
int launch_att_wfrag_pack(AttArgs& a, void* buf, cudaStream_t s);
int launch_att_bwd(const AttArgs& a, cudaStream_t s);
// att_free.cu: resident-weight free-running decoder for low-batch synthesis (one 16-CTA cluster per utterance, N <= 32)
bool att_free_supported(const AttArgs& a);
size_t att_free_image_bytes();
int launch_att_free(const AttArgs& a, void* img, cudaStream_t s);   // TACO_ENOTSUP when 16-CTA clusters cannot be launched
bool att_fast_supported(const AttArgs& a);
int launch_att_fast_fwd(const AttArgs& a, cudaStream_t s);   // TACO_ENOTSUP when 16-CTA clusters cannot be launched
int launch_att_fast_bwd(const AttArgs& a, cudaStream_t s);
void att_fast_image_bytes(int Ti, bool bwd, size_t* w_bytes, size_t* km_bytes_per_cta);
int launch_att_fast_pack(const AttArgs& a, bool bwd, int what, uint8_t* img_w, uint8_t* img_km, cudaStream_t s);
int launch_att_keys_bwd(const float* keys, const float* q, const float* ge, const float* v_eff, float* dkeys, float* gv,
                        int N, int Ti, int Td, int A, int fast, cudaStream_t s);
// forward (gveff == NULL): v_eff = g v/||v||;  backward: gv += d v_eff/d v . gveff, gg += d v_eff/d g . gveff
int launch_att_vnorm(const float* v, const float* g, float* v_eff, const float* gveff, float* gv, float* gg, int A, cudaStream_t s);
int launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t s);

// optim.cu
int launch_sqnorm(const float* g, long long n, double* out, cudaStream_t s);
int launch_adam_clip(float* p, float* m, float* v, const float* g, long long n, const double* sqnorm, float gscale,
                     float clip_norm, float lr_t, float b1, float b2, float eps, float lr, float* norm_out, cudaStream_t s);

}  // namespace taco
