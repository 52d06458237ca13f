// Model object behind the C ABI: configuration, bound parameter table, workspace plan.
#pragma once
#include <map>
#include <string>
#include <vector>
#include <unordered_map>
#include "common.cuh"
#include "kernels.h"

namespace taco {

struct Entry { int64_t offset; int64_t numel; int trainable; };

struct Region {
    size_t offset;      // bytes from workspace base
    int64_t numel;      // elements (fp32 unless is_double)
    int ndim; int64_t dims[4]; int64_t strides[4];
    bool is_double;
    bool is_half;       // bf16 mirror region (TACO_PREC_BF16)
};

// Geometry of one CBHG block in the zero-padded time layout.
struct CbhgGeom {
    std::string prefix;  // "enc_cbhg" | "post_cbhg"
    int N, T, Tp, PL, PR, rows, slack;
    int Cin, Kb, Cb, P1, P2, pw, depth, H;
    bool has_hin;
};

struct Shape {
    int N = 0, Ti = 0, To = 0, Td = 0;
    int training = 0;
    bool operator==(const Shape& o) const { return N == o.N && Ti == o.Ti && To == o.To && Td == o.Td && training == o.training; }
};

struct Model {
    taco_config cfg;
    float* params = nullptr; float* grads = nullptr; float* adam_m = nullptr; float* adam_v = nullptr; float* bn_state = nullptr;
    int64_t n_trainable = 0, n_state = 0;
    int64_t early_offset = 0;        // data-parallel buckets: gradients [early_offset, n_trainable) are complete before the encoder's backward pass
    bool early_recorded = false;
    std::unordered_map<std::string, Entry> table;

    char* ws = nullptr; size_t ws_bytes = 0;
    Shape shape; bool planned = false;
    std::map<std::string, Region> regions;
    size_t plan_bytes = 0;
    CbhgGeom enc, post;
    // backward operands that depend only on the parameters (model.cu: backward_prep)
    bool prep_done = false;          // produced beside the current forward pass on the leaf stream
    // attention operand images of the current step: 0 not built, 1 built on the consumer's own stream, 2 built on another
    // stream (consumer must image_wait first)
    int img_w_state = 0, img_bkm_state = 0;
    bool tables_ready = false;       // bank tap tables uploaded for the current plan / workspace
    std::vector<int> taps_enc, taps_post;

    // parameter access by name
    int lookup(const std::string& name, Entry& e) const;
    float* P(const std::string& name) const;   // parameter (trainable) or BN state
    float* G(const std::string& name) const;   // gradient slot of a trainable parameter
    bool has(const std::string& name) const { return table.count(name) != 0; }

    // workspace
    void plan(const Shape& s);
    float* W(const std::string& name) const;   // region pointer (fp32)
    // bf16 mirrors (TACO_PREC_BF16): region "<name>.h" next to (or instead of) the fp32 region; W16 yields nullptr when the plan
    // has no such mirror (other precisions), so call sites can pass it straight to the optional kernel arguments.  P16: the
    // parameter's slot in the bf16 mirror of the flat trainable buffer (region "params16", refreshed by every forward pass).
    void* W16(const std::string& name) const;
    void* P16(const std::string& name) const;
    bool use16() const { return cfg.precision == TACO_PREC_BF16; }
    double* Wd(const std::string& name) const; // region pointer (double)
    bool has_region(const std::string& name) const { return regions.count(name) != 0; }
};

// Stream for leaf work (weight / bias gradients) whose operands are produced by what is already enqueued on `main`:
// the low-priority side stream when taco_backward runs its two-stream schedule, otherwise `main` itself.
cudaStream_t fork_side(cudaStream_t main);

// Streams / events of the decoder wavefront (model.cu).  wave_get yields nullptr when the schedule is off (TACO_OVERLAP=0,
// TACO_DEC_CHUNKS<=1, or a taco_profile window, whose per-launch timings must not overlap).
struct WaveCtx {
    static constexpr int kMaxChunks = 16;
    cudaStream_t w[2];                       // one per residual GRU layer
    cudaEvent_t ev[3][kMaxChunks];           // [stage][chunk]: 0 attention, 1 GRU layer 1, 2 GRU layer 2
    cudaEvent_t start;
    int chunks;
};
int wave_get(WaveCtx** out);
cudaStream_t fork_side_after(cudaStream_t producer);
int image_ready(int which, cudaStream_t producer);
int image_wait(int which, cudaStream_t s);

void prof_mark(const char* name, cudaStream_t s);   // timeline stage marker (no-op unless taco_profile is collecting)
int prof_launch_gru(const GruArgs& a, bool bwd, cudaStream_t s);
int prof_launch_att(const AttArgs& a, bool bwd, cudaStream_t s);
int prof_launch_att_free(const AttArgs& a, void* img, cudaStream_t s);

float* bank_wd(const Model& m, const CbhgGeom& g, int k);   // flipped + transposed kernel of bank member k (packed, contiguous over k)

// model_cbhg.cu
int cbhg_forward(Model& m, const CbhgGeom& g, const int* lengths, const float* before_highway, const float* rnn_h0,
                 int training, cudaStream_t s);
int cbhg_backward(Model& m, const CbhgGeom& g, const int* lengths, bool want_dbefore, bool want_dh0, cudaStream_t s);

// model_decoder.cu
int decoder_forward(Model& m, const taco_batch* b, cudaStream_t s);
int decoder_pack_weight_images(Model& m, cudaStream_t s);     // forward + backward weight images of the fast attention kernels
int decoder_backward(Model& m, const taco_batch* b, cudaStream_t s);

}  // namespace taco
