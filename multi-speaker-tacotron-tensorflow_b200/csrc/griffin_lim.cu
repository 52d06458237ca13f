// Griffin-Lim spectrogram inversion on the GPU: cuFFT batched C2R / R2C plans + fused window / overlap-add /
// reflect-pad framing / phase-normalise kernels, and the inverse pre-emphasis IIR as a chunked scan.
//   reference: audio/__init__.py:54-56 (inv_spectrogram), :76-84 (_griffin_lim, 60 iterations), :99-106 (_stft/_istft via
//   librosa 0.5.1: centered, reflect-padded, periodic Hann of win_length zero-padded to n_fft, window-sum-square
//   normalisation), :118-122 (_stft_parameters), :149 (_db_to_amp), :158-159 (inv_preemphasis), :164-165 (_denormalize).
// librosa 0.5.x stores conj(FFT) in stft and conjugates again in istft; the loop is self-consistent under either
// convention, so the kernels use the plain FFT and the caller-supplied initial phase is negated to match.
#include "common.cuh"
#include <cufft.h>
#include <cstdlib>
#include <new>

namespace taco {

struct GlState {
    int n_fft, hop, win, max_frames, device, nbins;
    cufftHandle c2r = 0, r2c = 0; int planned_T = 0;
    cufftHandle ana_r2c = 0; int ana_T = 0;      // analysis front end (taco_audio_spectrogram)
    float* window = nullptr;     // [n_fft] periodic Hann(win) centred in n_fft
    float* wsum = nullptr;       // [n_fft + hop*(max_frames-1)] sum of squared windows
    float* mag = nullptr;        // [max_frames, nbins] target magnitudes (S^power)
    cufftComplex* spec = nullptr;// [max_frames, nbins]
    float* frames = nullptr;     // [max_frames, n_fft]   inverse-FFT frames
    float* frames2 = nullptr;    // [max_frames, n_fft]   re-windowed frames of the next forward transform (gl_ola_scatter_kernel); taps
                                 //                       outside the window stay zero from plan time
    float* y = nullptr;          // [n_fft + hop*(max_frames-1)] untrimmed signal
    float* carry = nullptr;      // IIR chunk states
    size_t bytes = 0;
    // The iteration loop (5 launches x n_iters of ~3 us kernels) is launch-bound: it is captured once per (T, n_iters) into a
    // CUDA graph and replayed (TACO_GL_GRAPH=0 issues the launches one by one).
    cudaGraphExec_t graph = nullptr; int graph_T = 0, graph_iters = -1; cudaStream_t capture_stream = nullptr;
};

#define TACO_CHECK_CUFFT(expr)                                                                    \
    do {                                                                                          \
        cufftResult _r = (expr);                                                                  \
        if (_r != CUFFT_SUCCESS) { set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)_r); return TACO_ECUDA; } \
    } while (0)

__global__ void gl_window_kernel(float* w, int n_fft, int win) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_fft) return;
    int lpad = (n_fft - win) / 2;       // librosa.util.pad_center
    int i = k - lpad;
    w[k] = (i >= 0 && i < win) ? 0.5f - 0.5f * cospif(2.0f * (float)i / (float)win) : 0.f;   // scipy get_window('hann', fftbins=True)
}
__global__ void gl_wsum_kernel(const float* __restrict__ w, float* __restrict__ wsum, int n_fft, int hop, int T, int len) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= len) return;
    float acc = 0.f;
    int i_hi = min(T - 1, s / hop);
    for (int i = i_hi; i >= 0; i--) {
        int k = s - i * hop;
        if (k >= n_fft) break;
        acc += w[k] * w[k];
    }
    wsum[s] = acc;
}
// magnitudes: S = 10^(0.05*(clip(x,0,1)*(-min_db) + min_db + ref_db)); mag = S^power; initial spectrum mag*exp(-2*pi*i*phase)
__global__ void gl_init_kernel(const float* __restrict__ lin, const float* __restrict__ phase, float* __restrict__ mag,
                               cufftComplex* __restrict__ spec, long long n, float min_db, float ref_db, float power) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = fminf(fmaxf(lin[i], 0.f), 1.f);
    float db = x * (-min_db) + min_db + ref_db;
    float S = powf(10.0f, db * 0.05f);
    float m = powf(S, power);
    mag[i] = m;
    float ph = phase ? -2.0f * phase[i] : 0.f;     // in units of pi; negated (see header comment)
    float sn, cs; sincospif(ph, &sn, &cs);
    spec[i] = make_cuFloatComplex(m * cs, m * sn);
}
// overlap-add of windowed inverse FFT frames, normalised by the window sum-square
__global__ void gl_ola_kernel(const float* __restrict__ frames, const float* __restrict__ w, const float* __restrict__ wsum,
                              float* __restrict__ y, int n_fft, int hop, int T, int len, float inv_nfft, float tiny) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= len) return;
    float acc = 0.f;
    int i_hi = min(T - 1, s / hop);
    for (int i = i_hi; i >= 0; i--) {
        int k = s - i * hop;
        if (k >= n_fft) break;
        acc += w[k] * frames[(long long)i * n_fft + k] * inv_nfft;
    }
    float ws = wsum[s];
    y[s] = (ws > tiny) ? acc / ws : acc;
}
// windowed frames of the centre-trimmed signal with reflect padding: frame i covers ypad[i*hop : i*hop+n_fft],
// ypad[j] = reflect(ytrim, j - n_fft/2), ytrim = y[n_fft/2 : n_fft/2 + L]
__global__ void gl_frame_kernel(const float* __restrict__ y, const float* __restrict__ w, float* __restrict__ frames,
                                int n_fft, int hop, int T, int L) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)T * n_fft) return;
    int k = (int)(idx % n_fft), i = (int)(idx / n_fft);
    int j = i * hop + k - n_fft / 2;          // index into ytrim
    if (j < 0) j = -j;                        // numpy 'reflect' (no edge repeat)
    if (j >= L) j = 2 * (L - 1) - j;
    j = min(max(j, 0), L - 1);
    frames[idx] = w[k] * y[n_fft / 2 + j];
}
// gl_ola_kernel + gl_frame_kernel in one pass, for the iterations that feed another forward transform: the thread that rebuilds
// sample s (the same loop, in the same order, as gl_ola_kernel - bit-identical) scatters w[k]*y[s] into every frame tap that reads
// it, so y never goes through memory and the 41 % of taps outside the Hann window (zero once, at plan time) are never touched.
// A tap (i, k) reads ytrim[reflect(i*hop + k - n_fft/2)], ytrim = y[n_fft/2 : n_fft/2 + L]: directly, mirrored at the start
// (index -j) or mirrored at the end (index 2(L-1) - j); L > n_fft/2 (checked by the caller) rules out double reflections.
__global__ void gl_ola_scatter_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ wsum,
                                      float* __restrict__ out, int n_fft, int hop, int T, int L, float inv_nfft, float tiny) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;      // index into ytrim
    if (j >= L) return;
    const int half = n_fft / 2, s = half + j;
    float acc = 0.f;
    for (int f = min(T - 1, s / hop); f >= 0; f--) {
        const int k = s - f * hop;
        if (k >= n_fft) break;
        acc += w[k] * in[(long long)f * n_fft + k] * inv_nfft;
    }
    const float ws = wsum[s];
    const float y = (ws > tiny) ? acc / ws : acc;
    // direct readers: i*hop + k - half = j
    for (int i = min(T - 1, s / hop); i >= 0; i--) {
        const int k = s - i * hop;
        if (k >= n_fft) break;
        const float wk = w[k];
        if (wk != 0.f) out[(long long)i * n_fft + k] = wk * y;
    }
    // readers mirrored at the start: i*hop + k - half = -j  (j > 0)
    if (j > 0 && j <= half) {
        for (int i = 0; i * hop <= half - j; i++) {
            const int k = half - j - i * hop;
            if (i >= T) break;
            const float wk = w[k];
            if (wk != 0.f) out[(long long)i * n_fft + k] = wk * y;
        }
    }
    // readers mirrored at the end: i*hop + k - half = 2(L-1) - j  (>= L)
    if (j <= L - 2) {
        const int jr = 2 * (L - 1) - j + half;                 // = i*hop + k
        for (int i = T - 1; i >= 0; i--) {
            const int k = jr - i * hop;
            if (k >= n_fft) break;
            if (k < 0) continue;
            const float wk = w[k];
            if (wk != 0.f) out[(long long)i * n_fft + k] = wk * y;
        }
    }
}
// spec <- mag * spec / |spec|   (angle(0) = 0 -> unit phase 1)
__global__ void gl_phase_kernel(cufftComplex* __restrict__ spec, const float* __restrict__ mag, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cufftComplex v = spec[i];
    float a = sqrtf(v.x * v.x + v.y * v.y);
    float m = mag[i];
    spec[i] = (a > 0.f) ? make_cuFloatComplex(m * v.x / a, m * v.y / a) : make_cuFloatComplex(m, 0.f);
}
// inverse pre-emphasis y[n] = x[n] + a*y[n-1] as a chunked scan (3 passes)
constexpr int IIR_CHUNK = 256;
__global__ void iir_local_kernel(const float* __restrict__ x, float* __restrict__ yo, float* __restrict__ carry, int L, float a) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int s0 = c * IIR_CHUNK;
    if (s0 >= L) return;
    float st = 0.f;
    for (int s = s0; s < min(L, s0 + IIR_CHUNK); s++) { st = x[s] + a * st; yo[s] = st; }
    carry[c] = st;
}
__global__ void iir_carry_kernel(float* carry, int nchunks, float apow) {
    if (blockIdx.x || threadIdx.x) return;
    float st = 0.f;     // state entering chunk c
    for (int c = 0; c < nchunks; c++) { float end = carry[c] + apow * st; carry[c] = st; st = end; }
}
__global__ void iir_fix_kernel(float* __restrict__ yo, const float* __restrict__ carry, int L, float a) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    int c = s / IIR_CHUNK;
    float st = carry[c];
    if (st != 0.f) yo[s] += st * powf(a, (float)(s - c * IIR_CHUNK + 1));
}

// ---- analysis front end (audio/__init__.py:48-51 spectrogram, :64-67 melspectrogram) ------------------------------------
// windowed frames of the pre-emphasised, reflect-padded signal; pre-emphasis x[n] = y[n] - a*y[n-1] (x[0] = y[0], scipy
// lfilter with zero initial state, :155-156) is applied on the fly so the filtered signal never exists in memory
__global__ void ana_frame_kernel(const float* __restrict__ y, const float* __restrict__ w, float* __restrict__ frames,
                                 int n_fft, int hop, int T, int L, float a) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)T * n_fft) return;
    int k = (int)(idx % n_fft), i = (int)(idx / n_fft);
    float wk = w[k];
    float v = 0.f;
    if (wk != 0.f) {                          // the Hann window covers win of the n_fft taps; the rest is exact zero padding
        int j = i * hop + k - n_fft / 2;
        if (j < 0) j = -j;
        if (j >= L) j = 2 * (L - 1) - j;
        j = min(max(j, 0), L - 1);
        v = wk * (y[j] - (j > 0 ? a * y[j - 1] : 0.f));
    }
    frames[idx] = v;
}
// one block per frame: |D| -> normalised dB linear row, and (optionally) mel = basis . |D| -> normalised dB mel row.
// The magnitudes of the frame are staged in shared memory; one warp per mel channel walks its basis row coalesced.
__global__ void ana_db_mel_kernel(const cufftComplex* __restrict__ spec, const float* __restrict__ basis, float* __restrict__ lin_out,
                                  float* __restrict__ mel_out, int nbins, int num_mels, float ref_db, float min_db) {
    extern __shared__ float mag[];
    const int i = blockIdx.x;
    const float inv = -1.0f / min_db;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
        cufftComplex v = spec[(long long)i * nbins + k];
        float m = sqrtf(v.x * v.x + v.y * v.y);
        mag[k] = m;
        if (lin_out) {
            float S = 20.0f * log10f(fmaxf(1e-5f, m)) - ref_db;
            lin_out[(long long)i * nbins + k] = fminf(fmaxf((S - min_db) * inv, 0.f), 1.f);
        }
    }
    if (!mel_out) return;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int c = warp; c < num_mels; c += nwarps) {
        const float* b = basis + (long long)c * nbins;
        float acc = 0.f;
        for (int k = lane; k < nbins; k += 32) acc = fmaf(b[k], mag[k], acc);
        #pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            float S = 20.0f * log10f(fmaxf(1e-5f, acc));      // no ref_level_db here: the reference's melspectrogram omits it
            mel_out[(long long)i * num_mels + c] = fminf(fmaxf((S - min_db) * inv, 0.f), 1.f);
        }
    }
}

}  // namespace taco

using namespace taco;
struct taco_gl_s { GlState st; };

extern "C" {

int taco_gl_create(taco_gl* out, int32_t n_fft, int32_t hop, int32_t win, int32_t max_frames, int32_t device) {
    TACO_REQUIRE(out && n_fft > 0 && hop > 0 && win > 0 && win <= n_fft && max_frames > 1, TACO_EINVAL, "taco_gl_create: bad arguments");
    taco_gl_s* h = new (std::nothrow) taco_gl_s();
    TACO_REQUIRE(h, TACO_ENOMEM, "taco_gl_create: out of host memory");
    GlState& g = h->st;
    g.n_fft = n_fft; g.hop = hop; g.win = win; g.max_frames = max_frames; g.device = device; g.nbins = n_fft / 2 + 1;
    TACO_CHECK_CUDA(cudaSetDevice(device));
    const size_t len = (size_t)n_fft + (size_t)hop * (max_frames - 1);
    const size_t nb = (size_t)max_frames * g.nbins;
    TACO_CHECK_CUDA(cudaMalloc(&g.window, sizeof(float) * n_fft));
    TACO_CHECK_CUDA(cudaMalloc(&g.wsum, sizeof(float) * len));
    TACO_CHECK_CUDA(cudaMalloc(&g.mag, sizeof(float) * nb));
    TACO_CHECK_CUDA(cudaMalloc(&g.spec, sizeof(cufftComplex) * nb));
    TACO_CHECK_CUDA(cudaMalloc(&g.frames, sizeof(float) * (size_t)max_frames * n_fft));
    TACO_CHECK_CUDA(cudaMalloc(&g.frames2, sizeof(float) * (size_t)max_frames * n_fft));
    TACO_CHECK_CUDA(cudaMalloc(&g.y, sizeof(float) * len));
    TACO_CHECK_CUDA(cudaMalloc(&g.carry, sizeof(float) * (len / IIR_CHUNK + 2)));
    g.bytes = sizeof(float) * (n_fft + 2 * len + nb + 2 * (size_t)max_frames * n_fft) + sizeof(cufftComplex) * nb;
    gl_window_kernel<<<cdiv(n_fft, 256), 256>>>(g.window, n_fft, win);
    TACO_CHECK_LAUNCH();
    TACO_CHECK_CUDA(cudaDeviceSynchronize());
    *out = h;
    return TACO_OK;
}

int taco_gl_destroy(taco_gl h) {
    if (!h) return TACO_OK;
    GlState& g = h->st;
    if (g.graph) cudaGraphExecDestroy(g.graph);
    if (g.capture_stream) cudaStreamDestroy(g.capture_stream);
    if (g.c2r) cufftDestroy(g.c2r);
    if (g.r2c) cufftDestroy(g.r2c);
    if (g.ana_r2c) cufftDestroy(g.ana_r2c);
    cudaFree(g.window); cudaFree(g.wsum); cudaFree(g.mag); cudaFree(g.spec); cudaFree(g.frames); cudaFree(g.frames2); cudaFree(g.y); cudaFree(g.carry);
    delete h;
    return TACO_OK;
}

size_t taco_gl_workspace_bytes(taco_gl h) { return h ? h->st.bytes : 0; }

int taco_gl_inv_spectrogram(taco_gl h, const float* linear_spec, const float* init_phase, int32_t T, int32_t n_iters, float power,
                            float min_level_db, float ref_level_db, float preemphasis, float* wav_out, void* ws, void* stream) {
    (void)ws;
    TACO_REQUIRE(h && linear_spec && wav_out, TACO_EINVAL, "taco_gl_inv_spectrogram: null argument");
    GlState& g = h->st;
    TACO_REQUIRE(T > 1 && T <= g.max_frames, TACO_ESHAPE, "taco_gl_inv_spectrogram: T=%d outside (1, %d]", T, g.max_frames);
    const int L = g.hop * (T - 1);                 // trimmed length
    TACO_REQUIRE(L > g.n_fft / 2, TACO_ESHAPE, "taco_gl_inv_spectrogram: signal too short for reflect padding");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int len = g.n_fft + L;
    if (g.planned_T != T) {
        if (g.graph) { cudaGraphExecDestroy(g.graph); g.graph = nullptr; }       // the graph references the old plans
        if (g.c2r) { cufftDestroy(g.c2r); g.c2r = 0; }
        if (g.r2c) { cufftDestroy(g.r2c); g.r2c = 0; }
        int n[1] = {g.n_fft};
        TACO_CHECK_CUFFT(cufftPlanMany(&g.c2r, 1, n, nullptr, 1, g.nbins, nullptr, 1, g.n_fft, CUFFT_C2R, T));
        TACO_CHECK_CUFFT(cufftPlanMany(&g.r2c, 1, n, nullptr, 1, g.n_fft, nullptr, 1, g.nbins, CUFFT_R2C, T));
        g.planned_T = T;
        gl_wsum_kernel<<<cdiv(len, 256), 256, 0, s>>>(g.window, g.wsum, g.n_fft, g.hop, T, len);
        TACO_CHECK_LAUNCH();
        TACO_CHECK_CUDA(cudaMemsetAsync(g.frames2, 0, sizeof(float) * (size_t)g.max_frames * g.n_fft, s));
    }
    const long long nb = (long long)T * g.nbins;
    gl_init_kernel<<<(unsigned)cdiv64(nb, 256), 256, 0, s>>>(linear_spec, init_phase, g.mag, g.spec, nb, min_level_db, ref_level_db, power);
    TACO_CHECK_LAUNCH();
    const float inv_nfft = 1.0f / (float)g.n_fft;
    // TACO_GL_FUSE=0: overlap-add and re-framing as two kernels (five launches per iteration instead of four)
    static const bool fused = [] { const char* e = getenv("TACO_GL_FUSE"); return !(e && e[0] == '0'); }();
    auto iterate = [&](cudaStream_t st) -> int {
        TACO_CHECK_CUFFT(cufftSetStream(g.c2r, st));
        TACO_CHECK_CUFFT(cufftSetStream(g.r2c, st));
        for (int it = 0; it <= n_iters; it++) {
            TACO_CHECK_CUFFT(cufftExecC2R(g.c2r, g.spec, g.frames));
            if (it == n_iters) {      // the last inverse transform produces the waveform
                gl_ola_kernel<<<cdiv(len, 256), 256, 0, st>>>(g.frames, g.window, g.wsum, g.y, g.n_fft, g.hop, T, len, inv_nfft, 1.17549435e-38f);
                TACO_CHECK_CUDA(cudaGetLastError());
                break;
            }
            if (fused) {
                gl_ola_scatter_kernel<<<cdiv(L, 256), 256, 0, st>>>(g.frames, g.window, g.wsum, g.frames2, g.n_fft, g.hop, T, L, inv_nfft, 1.17549435e-38f);
                TACO_CHECK_CUDA(cudaGetLastError());
                TACO_CHECK_CUFFT(cufftExecR2C(g.r2c, g.frames2, g.spec));
            } else {
                gl_ola_kernel<<<cdiv(len, 256), 256, 0, st>>>(g.frames, g.window, g.wsum, g.y, g.n_fft, g.hop, T, len, inv_nfft, 1.17549435e-38f);
                TACO_CHECK_CUDA(cudaGetLastError());
                gl_frame_kernel<<<(unsigned)cdiv64((long long)T * g.n_fft, 256), 256, 0, st>>>(g.y, g.window, g.frames, g.n_fft, g.hop, T, L);
                TACO_CHECK_CUDA(cudaGetLastError());
                TACO_CHECK_CUFFT(cufftExecR2C(g.r2c, g.frames, g.spec));
            }
            gl_phase_kernel<<<(unsigned)cdiv64(nb, 256), 256, 0, st>>>(g.spec, g.mag, nb);
            TACO_CHECK_CUDA(cudaGetLastError());
        }
        return TACO_OK;
    };
    static const bool use_graph = [] { const char* e = getenv("TACO_GL_GRAPH"); return !(e && e[0] == '0'); }();
    if (use_graph) {
        if (!g.graph || g.graph_T != T || g.graph_iters != n_iters) {
            if (g.graph) { cudaGraphExecDestroy(g.graph); g.graph = nullptr; }
            if (!g.capture_stream) TACO_CHECK_CUDA(cudaStreamCreateWithFlags(&g.capture_stream, cudaStreamNonBlocking));
            cudaGraph_t graph = nullptr;
            TACO_CHECK_CUDA(cudaStreamBeginCapture(g.capture_stream, cudaStreamCaptureModeThreadLocal));
            const int rc = iterate(g.capture_stream);
            const cudaError_t ce = cudaStreamEndCapture(g.capture_stream, &graph);
            if (rc != TACO_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            TACO_CHECK_CUDA(ce);
            TACO_CHECK_CUDA(cudaGraphInstantiate(&g.graph, graph, 0));
            TACO_CHECK_CUDA(cudaGraphDestroy(graph));
            g.graph_T = T; g.graph_iters = n_iters;
        }
        TACO_CHECK_CUDA(cudaGraphLaunch(g.graph, s));
    } else {
        TACO_TRY(iterate(s));
    }
    g_launch_count += (fused ? 4LL : 5LL) * n_iters + 2;           // kernels of the loop (cuFFT counts as one launch per transform)
    // inverse pre-emphasis on the trimmed signal
    const int nchunks = cdiv(L, IIR_CHUNK);
    iir_local_kernel<<<cdiv(nchunks, 128), 128, 0, s>>>(g.y + g.n_fft / 2, wav_out, g.carry, L, preemphasis);
    TACO_CHECK_LAUNCH();
    iir_carry_kernel<<<1, 32, 0, s>>>(g.carry, nchunks, powf(preemphasis, (float)IIR_CHUNK));
    TACO_CHECK_LAUNCH();
    iir_fix_kernel<<<cdiv(L, 256), 256, 0, s>>>(wav_out, g.carry, L, preemphasis);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

int taco_audio_spectrogram(taco_gl h, const float* wav, int32_t n_samples, float preemphasis, float ref_level_db, float min_level_db,
                           const float* mel_basis, int32_t num_mels, float* linear_out, float* mel_out, void* stream) {
    TACO_REQUIRE(h && wav && (linear_out || mel_out), TACO_EINVAL, "taco_audio_spectrogram: null argument");
    TACO_REQUIRE(!mel_out || (mel_basis && num_mels > 0), TACO_EINVAL, "taco_audio_spectrogram: mel_out needs mel_basis[num_mels, 1 + n_fft/2]");
    GlState& g = h->st;
    const int L = n_samples;
    const int T = 1 + L / g.hop;                   // librosa centred stft: 1 + (L + n_fft - n_fft) / hop
    TACO_REQUIRE(L > g.n_fft / 2, TACO_ESHAPE, "taco_audio_spectrogram: %d samples are too few for reflect padding (need > %d)", L, g.n_fft / 2);
    TACO_REQUIRE(T <= g.max_frames, TACO_ESHAPE, "taco_audio_spectrogram: %d frames exceed max_frames=%d", T, g.max_frames);
    TACO_REQUIRE(min_level_db < 0.f, TACO_EINVAL, "taco_audio_spectrogram: min_level_db must be negative");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (g.ana_T != T) {
        if (g.ana_r2c) { cufftDestroy(g.ana_r2c); g.ana_r2c = 0; }
        int n[1] = {g.n_fft};
        TACO_CHECK_CUFFT(cufftPlanMany(&g.ana_r2c, 1, n, nullptr, 1, g.n_fft, nullptr, 1, g.nbins, CUFFT_R2C, T));
        g.ana_T = T;
    }
    TACO_CHECK_CUFFT(cufftSetStream(g.ana_r2c, s));
    ana_frame_kernel<<<(unsigned)cdiv64((long long)T * g.n_fft, 256), 256, 0, s>>>(wav, g.window, g.frames, g.n_fft, g.hop, T, L, preemphasis);
    TACO_CHECK_LAUNCH();
    TACO_CHECK_CUFFT(cufftExecR2C(g.ana_r2c, g.frames, g.spec));
    g_launch_count++;
    ana_db_mel_kernel<<<T, 256, sizeof(float) * g.nbins, s>>>(g.spec, mel_basis, linear_out, mel_out, g.nbins, num_mels, ref_level_db, min_level_db);
    TACO_CHECK_LAUNCH();
    return TACO_OK;
}

}  // extern "C"
