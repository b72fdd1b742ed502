// Microbenchmark: TMA store throughput as a function of box shape (epilogue design input).
// Persistent CTAs; W warps per CTA each own a smem slab and loop { store box; commit; wait_group.read 0 }.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

struct Params { CUtensorMap map; int boxes_per_warp, warps, rows, tiles_total, pending; uint32_t slab_bytes; };

__global__ void __launch_bounds__(256) store_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (int)(p.slab_bytes * p.warps / 4); i += blockDim.x) ((uint32_t*)smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp < p.warps && lane == 0) {
        uint32_t src = (uint32_t)__cvta_generic_to_shared(smem + (size_t)warp * p.slab_bytes);
        // global row index space: rows are consecutive 'pixels'
        for (int t = blockIdx.x * p.warps + warp; t < p.tiles_total; t += gridDim.x * p.warps) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&p.map), "r"(src), "r"(0), "r"(t * p.rows) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (p.pending) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// reference: plain 16-byte stores, one row segment of 32 B per thread (the register-direct epilogue pattern)
__global__ void stg_kernel(uint4* dst, long long rows, int row_bytes) {
    const int per_row = row_bytes / 32;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * per_row; i += (long long)gridDim.x * blockDim.x) {
        // lane -> row (like the accumulator layout): consecutive lanes hit consecutive rows, same 32-byte column segment
        long long seg = i / 32 / rows * 0;  (void)seg;
        long long row = i % rows; int cseg = (int)(i / rows);
        uint4* q = (uint4*)((uint8_t*)dst + row * row_bytes + cseg * 32);
        q[0] = make_uint4(1, 2, 3, 4); q[1] = make_uint4(5, 6, 7, 8);
    }
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    const size_t cap = (size_t)1 << 30;
    uint8_t* arena; CK(cudaMalloc(&arena, cap));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Case { int row_bytes, box_bytes, rows, swz, warps, pending; };   // tensor row pitch, box inner bytes, rows per box
    Case cases[] = {
        {96, 32, 32, 32, 4, 0}, {96, 64, 32, 64, 4, 0}, {96, 96, 32, 0, 4, 0}, {96, 96, 128, 0, 1, 0}, {96, 96, 128, 0, 4, 0},
        {192, 128, 32, 128, 4, 0}, {192, 64, 32, 64, 4, 0}, {192, 192, 32, 0, 4, 0}, {192, 192, 128, 0, 1, 0}, {192, 192, 128, 0, 4, 0},
        {384, 128, 32, 128, 4, 0}, {384, 128, 32, 128, 8, 0}, {384, 128, 128, 128, 4, 0}, {384, 256, 32, 0, 4, 0}, {384, 384, 32, 0, 4, 0},
        {128, 128, 32, 128, 4, 0}, {128, 128, 128, 128, 4, 0}, {128, 128, 256, 128, 2, 0}, {64, 64, 32, 64, 4, 0}, {64, 64, 256, 64, 4, 0},
        {96, 96, 32, 0, 4, 1}, {192, 128, 32, 128, 4, 1}, {384, 128, 32, 128, 4, 1},
    };
    for (auto& c : cases) {
        Params p; memset(&p, 0, sizeof(p));
        long long rows_total = (long long)(cap / 2) / c.row_bytes;
        p.rows = c.rows; p.warps = c.warps; p.pending = c.pending;
        p.tiles_total = (int)(rows_total / c.rows);
        p.slab_bytes = (uint32_t)((c.rows * c.box_bytes + 1023) & ~1023) * (c.pending ? 1 : 1);
        cuuint64_t gd[2] = {(cuuint64_t)(c.row_bytes / 2), (cuuint64_t)rows_total};
        cuuint64_t gs[1] = {(cuuint64_t)c.row_bytes};
        cuuint32_t bx[2] = {(cuuint32_t)(c.box_bytes / 2), (cuuint32_t)c.rows}, es[2] = {1, 1};
        CUtensorMapSwizzle sw = c.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : c.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : c.swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
        CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, arena, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d for pitch %d box %d\n", (int)r, c.row_bytes, c.box_bytes); continue; }
        size_t smem = (size_t)p.slab_bytes * c.warps + 2048;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            store_kernel<<<prop.multiProcessorCount, 256, smem>>>(p);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        double bytes = (double)p.tiles_total * c.rows * c.box_bytes;
        printf("pitch %3d B box %3d B x %3d rows swz %3d warps %d pend %d: %8.1f us %7.1f GB/s  (%.0f MB, %.2f us per box per warp)\n", c.row_bytes, c.box_bytes, c.rows, c.swz,
               c.warps, c.pending, best * 1e3, bytes / (best * 1e-3) / 1e9, bytes / 1e6, best * 1e3 / ((double)p.tiles_total / (prop.multiProcessorCount * c.warps)));
    }
    for (int rb : {96, 192, 384}) {
        long long rows = (long long)(cap / 2) / rb;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            stg_kernel<<<prop.multiProcessorCount * 8, 256>>>((uint4*)arena, rows, rb);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("STG 2x16B per thread, lane=row, pitch %d: %8.1f us %7.1f GB/s\n", rb, best * 1e3, (double)rows * rb / (best * 1e-3) / 1e9);
    }
    return 0;
}
