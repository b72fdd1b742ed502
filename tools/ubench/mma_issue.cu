// Microbenchmark: how few cycles per tcgen05.mma can ONE elected thread sustain, as a function of how the issue loop is
// written?  N = 16 keeps the tensor pipe far from its math floor (8 cycles), so the number is the issue interval.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_p(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_1(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\nsetp.eq.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(ph) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t d64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// fully unrolled issue of one tap for MT tiles x KM k-steps; descriptors advance by immediates
template <int MT, int KM>
__device__ __forceinline__ void issue_tap(uint32_t d_tmem, int n_tile, uint32_t hi_a, uint32_t hi_b, uint32_t a_lo, uint32_t halo_units, uint32_t b_lo, uint32_t idesc, uint32_t first_acc) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int kk = 0; kk < KM; ++kk)
            mma_p(d_tmem + (uint32_t)(m * n_tile), d64(hi_a, a_lo + (uint32_t)m * halo_units + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, (kk == 0) ? first_acc : 1u);
}
template <int MT>
__device__ __forceinline__ void issue_tap_km(int km, uint32_t d_tmem, int n_tile, uint32_t hi_a, uint32_t hi_b, uint32_t a_lo, uint32_t halo_units, uint32_t b_lo, uint32_t idesc, uint32_t first_acc) {
    switch (km) {
        case 1: issue_tap<MT, 1>(d_tmem, n_tile, hi_a, hi_b, a_lo, halo_units, b_lo, idesc, first_acc); break;
        case 2: issue_tap<MT, 2>(d_tmem, n_tile, hi_a, hi_b, a_lo, halo_units, b_lo, idesc, first_acc); break;
        case 3: issue_tap<MT, 3>(d_tmem, n_tile, hi_a, hi_b, a_lo, halo_units, b_lo, idesc, first_acc); break;
        default: issue_tap<MT, 4>(d_tmem, n_tile, hi_a, hi_b, a_lo, halo_units, b_lo, idesc, first_acc); break;
    }
}
// variant 0: tight loop, 4 K offsets (baseline of mma_rate.cu)       1: halo-kernel shape (taps unrolled x9, 4 k, elect per tap, commit per tap)
// variant 2: like 1 but no commit/elect per tap (one elect around everything)   3: like 2 with accumulate as immediate   4: like 1 with N MMAs per tap = 8 (mt = 2)
__global__ void __launch_bounds__(128) k(int variant, int N, int rounds, long long* out, int mt, int chunks, int cin, uint32_t halo_bytes, uint32_t kh_bytes, uint32_t b_bytes, int nstages, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[12];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi_a = (1280u >> 4) | (1u << 14) | (2u << 29), hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a0 = ((s32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b0 = (((s32(smem) + 32768u) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t bar8 = s32(&bar[8]), bar9 = s32(&bar[9]);
    long long nmma = 0, t0 = 0, t1 = 0;
    if (warp == 1) {
        t0 = clock64();
        if (variant == 0) {
            if (lane == 0)
                for (int i = 0; i < rounds * 36; ++i) { const uint32_t st = (uint32_t)(i & 3); mma_p(tm, d64(hi_a, a0 + st * 2), d64(hi_b, b0 + st * 2), idesc, 1u); }
            nmma = rounds * 36;
        } else if (variant == 1 || variant == 4) {
            const int per = variant == 4 ? 2 : 1;
            int stage = 0;
            for (int r = 0; r < rounds; ++r) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    if (elect_one()) {
                        const uint32_t b_lo = b0 + (uint32_t)stage * 384u;
                        for (int m = 0; m < per; ++m) {
                            const uint32_t a_lo = a0 + (uint32_t)m * 1440u + (uint32_t)(((tap / 3) * 10 + tap % 3) * 8);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) mma_p(tm + (uint32_t)(m * N), d64(hi_a, a_lo + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, (kk == 0) ? (uint32_t)((r | tap) != 0) : 1u);
                        }
                        commit(s32(&bar[stage]));
                    }
                    if (++stage == 4) stage = 0;
                }
            }
            nmma = (long long)rounds * 36 * per;
        } else if (variant == 2) {
            if (elect_one()) {
                int stage = 0;
                for (int r = 0; r < rounds; ++r) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t b_lo = b0 + (uint32_t)stage * 384u;
                        const uint32_t a_lo = a0 + (uint32_t)(((tap / 3) * 10 + tap % 3) * 8);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) mma_p(tm, d64(hi_a, a_lo + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, (kk == 0) ? (uint32_t)((r | tap) != 0) : 1u);
                        if (++stage == 4) stage = 0;
                    }
                }
            }
            nmma = rounds * 36;
        } else if (variant == 3) {
            if (elect_one()) {
                int stage = 0;
                for (int r = 0; r < rounds; ++r) {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t b_lo = b0 + (uint32_t)stage * 384u;
                        const uint32_t a_lo = a0 + (uint32_t)(((tap / 3) * 10 + tap % 3) * 8);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) mma_1(tm, d64(hi_a, a_lo + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc);
                        if (++stage == 4) stage = 0;
                    }
                }
            }
            nmma = rounds * 36;
        }
        if (variant >= 32) {
            // proposed structure: leader elected once, taps unrolled, tile count via one uniform branch; 32: always 4 k-steps,
            // 33: per-MMA `kk < kmmas` test, 34: like 32 plus a try_wait on a completed barrier per tap (the real loop has one)
            const int f = variant - 32;
            int stage = 0, hb = 0, it = 0;
            const uint32_t smem_a = s32(smem), smem_b = smem_a + 2u * (uint32_t)mt * halo_bytes;
            const uint32_t b_units = b_bytes >> 4, halo_units = halo_bytes >> 4, kh_units = kh_bytes >> 4;
            const uint32_t a_lo0 = ((smem_a & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_b & 0x3FFFFu) >> 4) | (1u << 16);
            const int n_tile = N;
            const bool leader = elect_one();
            const bool mt2 = mt == 2;
            if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[10])) : "memory");
            __syncwarp();
            for (int rd = blockIdx.x; rd < rounds * (int)gridDim.x; rd += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t d_tmem = tm + (uint32_t)(as * mt * n_tile);
                for (int ch = 0; ch < chunks; ++ch) {
                    const uint32_t a_lo_h = a_lo0 + (uint32_t)(hb * mt) * halo_units;
                    const int kmmas = (f == 1 && ch == chunks - 1) ? ((cin - ch * 64 + 15) >> 4) : 4;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (f == 2) wait(s32(&bar[10]), 0);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (leader) {
                            const uint32_t b_lo = b_lo0 + (uint32_t)stage * b_units;
                            const uint32_t a_lo = a_lo_h + (uint32_t)(tap / 3) * kh_units + (uint32_t)(tap % 3) * 8u;
                            const uint32_t acc0 = tap == 0 ? (uint32_t)(ch != 0) : 1u;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                if (kk < kmmas) mma_p(d_tmem, d64(hi_a, a_lo + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, kk == 0 ? acc0 : 1u);
                            if (mt2) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    if (kk < kmmas) mma_p(d_tmem + (uint32_t)n_tile, d64(hi_a, a_lo + halo_units + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, kk == 0 ? acc0 : 1u);
                            }
                            commit(s32(&bar[stage]));
                            if (tap == 8) commit(bar8);
                        }
                        nmma += kmmas * mt;
                        if (++stage == nstages) { stage = 0; }
                    }
                    if (++hb == 2) { hb = 0; }
                }
            }
        } else if (variant >= 16) {
            const int f = variant - 16;
            const bool leader = elect_one();
            // a barrier that is already complete for parity 0... use bar[7]: arrive once so phase 0 completes
            if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[7])) : "memory");
            __syncwarp();
            int stage = 0;
            for (int r = 0; r < rounds; ++r) {
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    if (f & 4) wait(s32(&bar[7]), 0);
                    if (f & 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_lo = b0 + (uint32_t)stage * 384u;
                    const uint32_t a_lo = a0 + (uint32_t)(((tap / 3) * 10 + tap % 3) * 8);
                    if ((f & 1) ? elect_one() : leader) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) mma_p(tm, d64(hi_a, a_lo + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, (kk == 0) ? (uint32_t)((r | tap) != 0) : 1u);
                        if (f & 2) commit(s32(&bar[stage]));
                    }
                    if (++stage == 4) stage = 0;
                }
            }
            nmma = rounds * 36;
        } else if (variant >= 8) {
            // variant 8: leader elected once, warp-uniform address math, templated unrolled issue; 9: elect per tap around the unrolled issue
            int stage = 0, hb = 0, it = 0;
            const uint32_t smem_a = s32(smem), smem_b = smem_a + 2u * (uint32_t)mt * halo_bytes;
            const uint32_t b_units = b_bytes >> 4, halo_units = halo_bytes >> 4, kh_units = kh_bytes >> 4;
            const uint32_t a_lo0 = ((smem_a & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_b & 0x3FFFFu) >> 4) | (1u << 16);
            const int n_tile = N;
            const bool leader = elect_one();
            for (int rd = blockIdx.x; rd < rounds * (int)gridDim.x; rd += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t d_tmem = tm + (uint32_t)(as * mt * n_tile);
                for (int ch = 0; ch < chunks; ++ch) {
                    const uint32_t a_lo_h = a_lo0 + (uint32_t)(hb * mt) * halo_units;
                    const int kmmas = (ch == chunks - 1) ? ((cin - ch * 64 + 15) >> 4) : 4;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_lo = b_lo0 + (uint32_t)stage * b_units;
                        const uint32_t a_lo = a_lo_h + (uint32_t)(tap / 3) * kh_units + (uint32_t)(tap % 3) * 8u;
                        if (variant == 8 ? leader : elect_one()) {
                            if (mt == 2) issue_tap_km<2>(kmmas, d_tmem, n_tile, hi_a, hi_b, a_lo, halo_units, b_lo, idesc, (uint32_t)((ch | tap) != 0));
                            else issue_tap_km<1>(kmmas, d_tmem, n_tile, hi_a, hi_b, a_lo, halo_units, b_lo, idesc, (uint32_t)((ch | tap) != 0));
                            commit(s32(&bar[stage]));
                            if (tap == 8) commit(bar8);
                        }
                        nmma += kmmas * mt;
                        if (++stage == nstages) { stage = 0; }
                    }
                    if (++hb == 2) { hb = 0; }
                }
            }
        } else if (variant >= 5) {
            // verbatim shape of conv_tc_halo_kernel's MMA loop (no operand waits): runtime mt / chunks / kmmas, elect per tap,
            // commits per tap; variant 6 drops tc_fence, variant 7 additionally drops the per-tap commit
            int stage = 0, hb = 0, it = 0;
            const uint32_t smem_a = s32(smem), smem_b = smem_a + 2u * (uint32_t)mt * halo_bytes;
            const uint32_t b_units = b_bytes >> 4;
            const int n_tile = N;
            for (int rd = blockIdx.x; rd < rounds * (int)gridDim.x; rd += gridDim.x, ++it) {
                const int as = it & 1;
                const int nv = min(mt, total_tiles - rd * mt);
                const uint32_t d_tmem = tm + (uint32_t)(as * mt * n_tile);
                for (int ch = 0; ch < chunks; ++ch) {
                    if (variant < 6) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_base = smem_a + (uint32_t)(hb * mt) * halo_bytes;
                    const int kmmas = (ch == chunks - 1) ? ((cin - ch * 64 + 15) >> 4) : 4;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (variant < 6) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            const uint32_t b_lo = (((smem_b & 0x3FFFFu) >> 4) | (1u << 16)) + (uint32_t)stage * b_units;
                            for (int m = 0; m < nv; ++m) {
                                const uint32_t aa = a_base + (uint32_t)m * halo_bytes + (uint32_t)(tap / 3) * kh_bytes + (uint32_t)(tap % 3) * 128u;
                                const uint32_t a_lo = ((aa & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    if (kk == 0 || kk < kmmas)
                                        mma_p(d_tmem + (uint32_t)(m * n_tile), d64(hi_a, a_lo + 2 * kk), d64(hi_b, b_lo + 2 * kk), idesc, (kk == 0) ? (uint32_t)((ch | tap) != 0) : 1u);
                                nmma += kmmas;
                            }
                            if (variant < 7) commit(s32(&bar[stage]));
                            if (tap == 8) {
                                commit(bar8);
                            }
                        }
                        if (++stage == nstages) { stage = 0; }
                    }
                    if (++hb == 2) { hb = 0; }
                }
            }
            nmma = __shfl_sync(0xffffffffu, nmma, 0);
            long long mx = nmma;
            for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            nmma = mx;
        }
        t1 = clock64();
        if (lane == 0) { commit(bar9); wait(bar9, 0); }
        long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = nmma; }
        (void)bar8;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    long long* out; CK(cudaMalloc(&out, 64)); long long h[4];
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg { int N, mt, chunks, cin; uint32_t halo, kh, bb; int st; };
    Cfg cfgs[] = {{192, 1, 3, 192, 25600, 2560, 24576, 5}, {96, 2, 2, 96, 23552, 1280, 12288, 7}, {48, 2, 1, 48, 23552, 1280, 6144, 8}};
    for (int variant = 32; variant <= 34; ++variant)
        for (auto& c : cfgs) {
            CK(cudaMemset(out, 0, 64));
            k<<<prop.multiProcessorCount, 128, 200 * 1024>>>(variant, c.N, 20, out, c.mt, c.chunks, c.cin, c.halo, c.kh, c.bb, c.st, 1 << 30);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
            printf("variant %d N %3d mt %d chunks %d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (math floor %5.1f), %lld mmas\n", variant, c.N, c.mt, c.chunks, (double)h[0] / h[2], (double)h[1] / h[2], c.N / 2.0, h[2]);
        }
    return 0;
}
