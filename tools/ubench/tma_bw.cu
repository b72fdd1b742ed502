// Microbenchmark: how fast can one SM's TMA unit stream activation boxes out of HBM?
// Persistent CTAs; warp 0 issues TMA loads into an S-stage ring, warp 1 releases the slots as soon
// as they land.  Reports useful GB/s for conv-shaped 4-D boxes, dense 2-D boxes and 1-D bulk copies.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}

struct Params {
    CUtensorMap map;
    const uint8_t* raw;
    int mode;        // 0: 4-D conv box {64,bw,bh,1}; 1: 2-D {64 ch, 128 px}; 2: 1-D bulk 16 KB; 3: 4-D box + L2 prefetch ahead
    int stages, chunks, tiles_x, tiles_y, nimg, bw, bh;
    uint32_t box_bytes;
    long long total_bytes;
    int ahead;
};

__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ Params p, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * p.box_bytes);
    uint64_t* empty = full + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.nimg;
    if (warp == 0 && lane == 0) {
        int st = 0; uint32_t ph = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
            int tx = (int)(t % p.tiles_x); long long m = t / p.tiles_x; int ty = (int)(m % p.tiles_y); int n = (int)(m / p.tiles_y);
            if (p.mode == 3) {
                long long tp = t + (long long)p.ahead * gridDim.x;
                if (tp < tiles) {
                    int ptx = (int)(tp % p.tiles_x); long long pm = tp / p.tiles_x; int pty = (int)(pm % p.tiles_y); int pn = (int)(pm / p.tiles_y);
                    for (int ch = 0; ch < p.chunks; ++ch)
                        asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(&p.map), "r"(ch * 64), "r"(ptx * p.bw), "r"(pty * p.bh), "r"(pn) : "memory");
                }
            }
            for (int ch = 0; ch < p.chunks; ++ch) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect(&full[st], p.box_bytes);
                uint32_t dst = s32(smem + (size_t)st * p.box_bytes), bar = s32(&full[st]);
                if (p.mode == 0 || p.mode == 3) {
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(&p.map), "r"(bar), "r"(ch * 64), "r"(tx * p.bw), "r"(ty * p.bh), "r"(n) : "memory");
                } else if (p.mode == 1) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(&p.map), "r"(bar), "r"(ch * 64), "r"((int)(t * 128)) : "memory");
                } else {
                    const uint8_t* src = p.raw + ((t * p.chunks + ch) * (long long)p.box_bytes) % p.total_bytes;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(p.box_bytes), "r"(bar) : "memory");
                }
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        int st = 0; uint32_t ph = 0; unsigned long long acc = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x)
            for (int ch = 0; ch < p.chunks; ++ch) {
                mbar_wait(&full[st], ph);
                acc += *(volatile unsigned long long*)(smem + (size_t)st * p.box_bytes);
                mbar_arrive(&empty[st]);
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
        if (acc == 0x1234567ull) *sink = acc;
    }
}

// plain LDG streaming reference: every thread reads 16 B per iteration, grid-stride
__global__ void ldg_kernel(const uint4* src, long long n16, unsigned long long* sink) {
    unsigned long long acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = __ldg(src + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    return (PFN_cuTensorMapEncodeTiled_v12000)ptr;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    auto enc = get_encode();
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    const size_t cap = (size_t)1 << 30;          // 1 GiB arena
    uint8_t* arena; CK(cudaMalloc(&arena, cap)); CK(cudaMemset(arena, 1, cap));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    printf("device %s, %d SMs\n", prop.name, sms);
    {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            ldg_kernel<<<sms * 8, 512>>>((const uint4*)arena, (long long)(cap / 16), sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) printf("LDG.128 stream of 1 GiB: %.1f GB/s\n", cap / (ms * 1e-3) / 1e9);
        }
    }
    struct Case { int mode, C, HW, stages, cps, bw, bh, ahead; };
    Case cases[] = {
        {0, 96, 160, 4, 1, 8, 16, 0}, {0, 96, 160, 8, 1, 8, 16, 0}, {0, 96, 160, 12, 1, 8, 16, 0}, {0, 96, 160, 6, 2, 8, 16, 0},
        {0, 96, 160, 4, 4, 8, 16, 0},
        {0, 64, 160, 8, 1, 8, 16, 0}, {0, 48, 160, 8, 1, 8, 16, 0}, {0, 192, 80, 8, 1, 8, 16, 0}, {0, 192, 80, 12, 1, 8, 16, 0},
        {0, 96, 160, 8, 1, 16, 8, 0}, {0, 96, 160, 8, 1, 32, 4, 0}, {0, 64, 160, 8, 1, 32, 4, 0},
        {1, 64, 160, 8, 1, 8, 16, 0}, {1, 64, 160, 12, 1, 8, 16, 0}, {1, 96, 160, 8, 1, 8, 16, 0}, {1, 192, 80, 8, 1, 8, 16, 0},
        {2, 64, 160, 4, 1, 8, 16, 0}, {2, 64, 160, 8, 1, 8, 16, 0}, {2, 64, 160, 12, 1, 8, 16, 0}, {2, 64, 160, 6, 2, 8, 16, 0},
        {3, 96, 160, 8, 1, 8, 16, 2}, {3, 96, 160, 8, 1, 8, 16, 4}, {3, 96, 160, 8, 1, 8, 16, 8}, {3, 192, 80, 8, 1, 8, 16, 4},
    };
    for (auto& c : cases) {
        Params p; memset(&p, 0, sizeof(p));
        p.mode = c.mode; p.stages = c.stages; p.bw = c.bw; p.bh = c.bh; p.ahead = c.ahead;
        p.chunks = (c.C + 63) / 64;
        p.box_bytes = 128 * 128;
        size_t per_img = (size_t)c.HW * c.HW * c.C * 2;
        p.nimg = (int)((cap * 3 / 4) / per_img);
        p.tiles_x = c.HW / c.bw; p.tiles_y = c.HW / c.bh;
        p.raw = arena; p.total_bytes = (long long)(cap / 2);
        double useful;
        if (c.mode == 0 || c.mode == 3) {
            cuuint64_t gd[4] = {(cuuint64_t)c.C, (cuuint64_t)c.HW, (cuuint64_t)c.HW, (cuuint64_t)p.nimg};
            cuuint64_t gs[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.HW * c.C * 2, (cuuint64_t)per_img};
            cuuint32_t bx[4] = {64, (cuuint32_t)c.bw, (cuuint32_t)c.bh, 1}, es[4] = {1, 1, 1, 1};
            CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, arena, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            useful = (double)p.nimg * per_img;
        } else if (c.mode == 1) {
            cuuint64_t gd[2] = {(cuuint64_t)c.C, (cuuint64_t)p.nimg * c.HW * c.HW};
            cuuint64_t gs[1] = {(cuuint64_t)c.C * 2};
            cuuint32_t bx[2] = {64, 128}, es[2] = {1, 1};
            CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, arena, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            p.tiles_x = c.HW * c.HW / 128; p.tiles_y = 1;
            useful = (double)p.nimg * per_img;
        } else {
            p.tiles_x = c.HW * c.HW / 128; p.tiles_y = 1;
            useful = (double)p.tiles_x * p.nimg * p.chunks * p.box_bytes;
        }
        size_t smem = (size_t)c.stages * p.box_bytes + 1024 + 512;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            stream_kernel<<<sms * c.cps, 64, smem>>>(p, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("mode %d C %3d HW %3d box %2dx%2d stages %2d ctas/sm %d ahead %d: %8.1f us  %7.1f GB/s useful (%.0f MB)\n", c.mode, c.C, c.HW, c.bw, c.bh,
               c.stages, c.cps, c.ahead, best * 1e3, useful / (best * 1e-3) / 1e9, useful / 1e6);
    }
    return 0;
}
