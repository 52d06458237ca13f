// TMA ingest microbenchmark, part 2: the first part found a constant ~754 clk per cp.async.bulk.tensor instruction whatever the
// box size (one issuing thread).  Here: several issuing warps / lanes, 3-D boxes (several k-tiles per instruction), tensor map
// in global memory, prefetch.tensormap.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
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
__device__ __forceinline__ void tma3d(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}

// `nissue` issuers per CTA, each with its own ring of `depth` slots of box_bytes.  issuer i = thread (i % lanes) of warp (i / lanes).
__global__ void __launch_bounds__(256, 1) tma_kernel(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_gmap, int dims3, int depth, int box_bytes,
                                                     int bw, int bh, int bz, int cols, int rows, int iters, int nissue, int lanes, int prefetch, unsigned long long* out_clk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + nissue * depth * box_bytes);
    const CUtensorMap* map = use_gmap ? gmap : &pmap;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nissue * depth; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (prefetch) asm volatile("prefetch.tensormap [%0];" ::"l"(map));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int me = warp * lanes + lane;
    const bool issuer = lane < lanes && me < nissue;
    const int cblocks = cols / (bw * (dims3 ? bz : 1)), rblocks = rows / bh;
    long long t0 = clock64();
    if (issuer) {
        uint8_t* ring = smem + me * depth * box_bytes;
        uint64_t* mybar = bar + me * depth;
        auto issue = [&](int it) {
            const int slot = it % depth;
            const long long idx = ((long long)blockIdx.x * 7919 + me * 131 + it);
            const int cb = (int)(idx % cblocks), rb = (int)((idx / cblocks) % rblocks);
            mbar_expect(&mybar[slot], box_bytes);
            if (dims3) tma3d(ring + slot * box_bytes, map, 0, rb * bh, cb * bz, &mybar[slot]);
            else tma2d(ring + slot * box_bytes, map, cb * bw, rb * bh, &mybar[slot]);
        };
        for (int it = 0; it < depth && it < iters; it++) issue(it);
        for (int it = 0; it < iters; it++) {
            mbar_wait(&mybar[it % depth], (it / depth) & 1);
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
    const int rows = 4608, cols = 2048;
    void* buf; cudaMalloc(&buf, (size_t)rows * cols * 2); cudaMemset(buf, 1, (size_t)rows * cols * 2);
    unsigned long long* clk; cudaMalloc(&clk, 148 * 8);
    CUtensorMap* gmap; cudaMalloc(&gmap, sizeof(CUtensorMap));
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Cfg { const char* name; int bw, bh, bz, d3, nissue, lanes, gm, pf; };
    const Cfg cfgs[] = {
        {"2D 64x128, 1 issuer", 64, 128, 1, 0, 1, 1, 0, 0},
        {"2D 64x128, 1 issuer, prefetch.tensormap", 64, 128, 1, 0, 1, 1, 0, 1},
        {"2D 64x128, 1 issuer, map in global", 64, 128, 1, 0, 1, 1, 1, 0},
        {"2D 64x128, 2 issuers (2 warps)", 64, 128, 1, 0, 2, 1, 0, 0},
        {"2D 64x128, 4 issuers (4 warps)", 64, 128, 1, 0, 4, 1, 0, 0},
        {"2D 64x128, 4 issuers (4 lanes of 1 warp)", 64, 128, 1, 0, 4, 4, 0, 0},
        {"2D 64x64, 8 issuers (8 warps)", 64, 64, 1, 0, 8, 1, 0, 0},
        {"3D 64x128x2 (32 KB), 1 issuer", 64, 128, 2, 1, 1, 1, 0, 0},
        {"3D 64x128x4 (64 KB), 1 issuer", 64, 128, 4, 1, 1, 1, 0, 0},
        {"3D 64x256x2 (64 KB), 1 issuer", 64, 256, 2, 1, 1, 1, 0, 0},
        {"3D 64x128x2 (32 KB), 2 issuers", 64, 128, 2, 1, 2, 1, 0, 0},
    };
    for (const Cfg& c : cfgs) {
        CUtensorMap map;
        CUresult r;
        if (c.d3) {
            cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)cols / 64}; cuuint64_t strides[2] = {(cuuint64_t)cols * 2, 128};
            cuuint32_t box[3] = {64, (cuuint32_t)c.bh, (cuuint32_t)c.bz}; cuuint32_t es[3] = {1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh}; cuuint32_t es[2] = {1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice);
        const int box_bytes = c.bw * c.bh * c.bz * 2;
        for (int depth : {1, 2, 3}) {
            if ((size_t)c.nissue * depth * box_bytes > 210 * 1024) continue;
            const int iters = 1000 * 16384 / box_bytes;
            const size_t smem = (size_t)c.nissue * depth * box_bytes + 1024 + 512;
            for (int rep = 0; rep < 2; rep++)
                tma_kernel<<<148, 256, smem>>>(map, gmap, c.gm, c.d3, depth, box_bytes, c.bw, c.bh, c.bz, cols, rows, iters, c.nissue, c.lanes, c.pf, clk);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            std::vector<unsigned long long> h(148);
            cudaMemcpy(h.data(), clk, 148 * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (auto v : h) avg += (double)v; avg /= 148;
            printf("%-44s depth %d: %6.1f B/clk/SM  (%.0f clk per box per issuer)\n", c.name, depth, (double)iters * box_bytes * c.nissue / avg, avg / iters);
        }
    }
    return 0;
}
