// TMA ingest microbenchmark (sm_100a): how many bytes per clock can one SM pull into shared memory through cp.async.bulk.tensor,
// as a function of box shape, boxes in flight, swizzle and the number of active CTAs?  (Sets the ceiling of gemm_bf16.cu.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tma_bench tools/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

// Each CTA streams `iters` boxes {bw elements, bh rows} through a ring of `depth` slots; one thread issues and waits.
// mode 0: tensor map; mode 1: 1-D bulk copies of the same byte count (contiguous).
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap map, const char* base, int mode, int depth, int box_bytes, int bw, int bh,
                                                     int cols, int rows, int iters, int nthreads_issue, unsigned long long* out_clk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + depth * box_bytes);
    if (threadIdx.x == 0) { for (int i = 0; i < depth; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int cblocks = cols / bw, rblocks = rows / bh;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        // box sequence: walk column blocks of one row block first (like the K loop of a GEMM over a row-major A), CTA-dependent start
        auto issue = [&](int it) {
            const int slot = it % depth;
            const long long idx = (long long)blockIdx.x * 7919 + it;
            const int cb = (int)(idx % cblocks), rb = (int)((idx / cblocks) % rblocks);
            mbar_expect(&bar[slot], box_bytes);
            if (mode == 0) tma2d(smem + slot * box_bytes, &map, cb * bw, rb * bh, &bar[slot]);
            else bulk1d(smem + slot * box_bytes, base + ((size_t)idx * box_bytes) % ((size_t)cols * rows * 2 - box_bytes) / 16 * 16, box_bytes, &bar[slot]);
        };
        for (int it = 0; it < depth && it < iters; it++) issue(it);
        for (int it = 0; it < iters; it++) {
            mbar_wait(&bar[it % depth], (it / depth) & 1);
            if (it + depth < iters) issue(it + depth);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = clock64() - t0;
}

static PFN_cuTensorMapEncodeTiled_v12000 enc;
int main() {
    cudaDriverEntryPointQueryResult q; void* fn = nullptr;
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
    enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    const int rows = 4608, cols = 2048;              // 18.9 MB of bf16: L2 resident after the first pass
    void* buf; cudaMalloc(&buf, (size_t)rows * cols * 2); cudaMemset(buf, 1, (size_t)rows * cols * 2);
    const int big_rows = 65536;                       // 268 MB: streams from DRAM
    void* bigbuf; cudaMalloc(&bigbuf, (size_t)big_rows * cols * 2); cudaMemset(bigbuf, 1, (size_t)big_rows * cols * 2);
    unsigned long long* clk; cudaMalloc(&clk, 148 * 8);
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { const char* name; int bw, bh, swz, mode; };
    const Cfg cfgs[] = {{"box 64x128 sw128", 64, 128, 1, 0}, {"box 64x128 none ", 64, 128, 0, 0}, {"box 64x256 sw128", 64, 256, 1, 0}, {"box 64x64  sw128", 64, 64, 1, 0},
                        {"box 256x32 none (512B rows)", 256, 32, 0, 0}, {"bulk1d 16KB", 64, 128, 0, 1}};
    for (int src = 0; src < 2; src++) {
        const int R = src ? big_rows : rows;
        void* b = src ? bigbuf : buf;
        printf("---- source: %s\n", src ? "268 MB (DRAM stream)" : "18.9 MB (L2 resident)");
        for (const Cfg& c : cfgs) {
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh}; cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
            const int box_bytes = c.bw * c.bh * 2;
            for (int ctas : {148, 36, 1}) {
                for (int depth : {1, 2, 4, 8}) {
                    if (depth * box_bytes > 190 * 1024) continue;
                    const int iters = 2000 * 16384 / box_bytes;
                    const size_t smem = (size_t)depth * box_bytes + 1024 + 128;
                    for (int rep = 0; rep < 2; rep++)
                        tma_kernel<<<ctas, 128, smem>>>(map, (const char*)b, c.mode, depth, box_bytes, c.bw, c.bh, cols, R, iters, 1, clk);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
                    std::vector<unsigned long long> h(ctas);
                    cudaMemcpy(h.data(), clk, ctas * 8, cudaMemcpyDeviceToHost);
                    double avg = 0; for (auto v : h) avg += (double)v; avg /= ctas;
                    printf("%-28s ctas %3d depth %d: %6.1f B/clk/SM  (%.0f clk per box)\n", c.name, ctas, depth, (double)iters * box_bytes / avg, avg / iters);
                }
            }
        }
    }
    return 0;
}
