// K2 -- implicit-GEMM convolution + bias + SiLU (+ residual, + channel-offset write) on the
// 5th-generation tensor cores: tcgen05.mma with the accumulator in TMEM, both operands staged
// in shared memory by TMA, one elected thread issuing the MMAs.
//
// Replaces every Conv(+folded BN)+Sigmoid*Mul node onnxruntime executes inside
// `session.run` (reference call sites simple_detector.py:474, :666; _script/gpu_handler.py:165).
//
// GEMM view:  D[M = pixels, N = Cout] = sum over taps (kh,kw) and channel chunks of
//             A_tap[M, kc] * W_tap[N, kc]^T
//   * activations are NHWC bf16; an M tile is a box of bw x bh pixels of bn images (= 128 rows).
//     For tap (kh,kw) the A operand is the *same box shifted by the tap offset*, fetched by one
//     4-D TMA tiled load; the zero padding of the convolution is TMA's out-of-bounds fill, so
//     there is no im2col buffer and no halo branch anywhere.
//   * stride-2 3x3 convs read through four "phase" tensor maps (even/odd rows x even/odd
//     columns of the input), each of which is again a dense tiled map.
//   * weights are pre-packed [Cout][kh][kw][Cin] (K contiguous) and fetched by a 2-D TMA load.
//   * both operands are K-major in shared memory with the hardware swizzle matching the chunk
//     width: 64 channels -> SWIZZLE_128B, 32 -> SWIZZLE_64B, 16 -> SWIZZLE_32B.
//   * the epilogue reads the fp32 accumulator from TMEM (tcgen05.ld 32x32b), adds the bias,
//     applies SiLU, adds the residual, rounds to bf16 and writes at a channel offset of the
//     destination buffer -- Concat / Split never run as ops.
//
// Kernel shape: persistent, one CTA per SM, 256 threads = 8 warps:
//   warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warp 3 idle |
//   warps 4-7 epilogue (warp%4 selects the TMEM lane quarter).
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA; two TMEM accumulator
// stages (tmem_full/tmem_empty) between MMA and epilogue, so tile i's epilogue overlaps
// tile i+1's main loop.
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

constexpr int kThreads = 384;     // 4 control warps + 8 epilogue warps
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(addr), "r"(parity) : "memory");
}
// 32-bit shared-address forms for the warp-uniform producer / MMA loops
__device__ __forceinline__ void mbar_wait_u32(uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// one lane of a converged warp (always the same one for a full mask)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(const CUtensorMap* map, uint32_t bar, uint32_t dst_smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit_u32(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t swizzle_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                   // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8u * swizzle_bytes) >> 4) << 32;         // stride byte offset: 8 rows of one swizzle span
    d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
    uint64_t layout = swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6);
    d |= layout << 61;
    return d;
}

struct TileCoord { int nt, x0, y0, n0; };

// Debug trace (B2D_TRACE=1): CTA 0 records clock64() at role events of its first kTraceTiles tiles.
constexpr int kTraceTiles = 24, kTraceEvents = 4, kTraceRoles = 3;   // roles: 0 producer, 1 MMA, 2 epilogue warp 4
__device__ __forceinline__ void trace(const ConvTcParams& p, int role, int it, int ev) {
#ifdef B2D_ENABLE_TRACE     // compiled out of product builds: even a predicted-off branch in the MMA issue loop costs
    if (p.trace && blockIdx.x == 0 && it < kTraceTiles) p.trace[(role * kTraceTiles + it) * kTraceEvents + ev] = clock64();
#endif
}

__device__ __forceinline__ TileCoord decode_tile(const ConvTcParams& p, int t) {
    TileCoord c;
    c.nt = t % p.n_tiles_n;
    int m = t / p.n_tiles_n;
    int tx = m % p.tiles_x;
    m /= p.tiles_x;
    int ty = m % p.tiles_y;
    int tg = m / p.tiles_y;
    c.x0 = tx * p.bw;
    c.y0 = ty * p.bh;
    c.n0 = tg * p.bn;
    return c;
}

// ===================== epilogue (warps 4-7), shared by both kernels =====================
// Warp q owns accumulator rows [32q, 32q+32) = a sub-box of the tile's pixels.  Per tile it
//   1. (residual layers) TMA-loads the residual sub-box into its staging slab,
//   2. waits for the accumulator, reads it 32 columns at a time (tcgen05.ld 32x32b.x32),
//      adds the bias (smem), applies SiLU, adds the residual read back from the slab, rounds to
//      bf16 (or keeps fp32 for head outputs) and writes the slab in the TMA swizzle pattern
//      (conflict-free 16-byte shared stores),
//   3. releases the TMEM stage, then TMA-stores the slab at the channel offset of the
//      destination buffer.  TMA clips partial tiles and padded channels, so there is no
//      per-thread bounds logic and every global write is a full coalesced row.
// A row of n_tile columns is cut into chunks of 128 / 64 / 32 bytes (EpiChunk), one tensor map
// per chunk width.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src_smem),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_s(const CUtensorMap* map, uint32_t bar, uint32_t dst_smem, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_smem),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }

// SiLU with one MUFU op per element: e = 2^(-v log2 e) on the SFU, 1/(1+e) on the FMA pipe
// (bit-trick seed, two Newton steps: relative error 6e-6, far below the bf16 rounding that
// follows).  The SFU (16 lanes/clk/SM) is the scarce pipe of the epilogue: 2 MUFU per output
// element made the memory-bound 1x1 layers epilogue-bound.
__device__ __forceinline__ float silu(float v) {
#ifdef B2D_SILU_MUFU2
    return __fdividef(v, 1.0f + __expf(-v));
#endif
    float y = fminf(v * -1.4426950408889634f, 64.0f), e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
    const float d = 1.0f + e;
    float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
    r = fmaf(r, fmaf(-d, r, 1.0f), r);
    r = fmaf(r, fmaf(-d, r, 1.0f), r);
    return v * r;
}

// 16 accumulator columns of this thread's row (already in registers) -> staging slab.
// `base` is the swizzled address of the unit's first 16-byte piece; the others are base ^ (j << 4).
template <int ACT, int RES, int F32>
__device__ __forceinline__ void epi_unit(const uint32_t (&r)[16], uint32_t bias_addr, uint32_t base) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 b = lds_f4(bias_addr + 16 * j);
        v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
        v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
        v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
        v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
    }
    if (ACT) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = silu(v[i]);
    }
    if (F32) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            sts128(base ^ (uint32_t)(j << 4),
                   make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t a0 = base ^ (uint32_t)(j << 4);
            if (RES) {
                const uint4 x = lds128(a0);
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[8 * j + 2 * i] += __uint_as_float(w[i] << 16);
                    v[8 * j + 2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
                }
            }
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
                w[i] = *(uint32_t*)&h;
            }
            sts128(a0, make_uint4(w[0], w[1], w[2], w[3]));
        }
    }
}

// Eight epilogue warps: warp w serves TMEM lane quarter q = w % 4 (a hardware rule) and, of that
// quarter's 16-column units, the ones with unit % 2 == (w - 4) / 4.  The two warps of a quarter
// share one staging slab and meet at a 64-thread named barrier before the slab is reused and
// before its TMA store is issued (by lane 0 of the first warp).  Everything that does not depend
// on the tile (slab addresses of this warp's units) is computed once, and the unit loop is fully
// unrolled so those values live in registers.
template <int ACT, int RES, int F32>
__device__ __forceinline__ void epilogue_loop(const ConvTcParams& p, int total_tiles, uint32_t tmem_base, const float* bias_s,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* res_bar, uint8_t* stg_base, int warp,
                                              int lane) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int n_tile = p.n_tile, nchunks = p.epi_nchunks;
    constexpr int esize = F32 ? 4 : 2;
    constexpr int ppu = F32 ? 4 : 2;                   // 16-byte pieces per 16-column unit
    const uint32_t slab0 = smem_u32(stg_base) + (uint32_t)q * 32u * (uint32_t)(n_tile * esize);   // this quarter's staging region
    const uint32_t slab_stride = p.stg_bufs == 2 ? 128u * (uint32_t)(n_tile * esize) : 0u;         // second buffer (if any)
    const uint32_t rbar = smem_u32(&res_bar[q]);
    const uint32_t bias_base = smem_u32(bias_s);
    const bool issuer = (half == 0 && lane == 0);
    // sub-box of the tile covered by this quarter's 32 rows
    const int box_px = p.bw * p.bh;
    const int yq = ((q * 32) % box_px) / p.bw, nq = (q * 32) / box_px;
    const int nunits = n_tile >> 4;
    const int my_units = (nunits - half + 1) >> 1;     // units half, half + 2, ...
    // Swizzled slab address of this lane's row for the 16-column unit starting at byte `b` of the row.  Mirrors
    // the host's chunking (conv_tc_plan): full 128-byte chunks first, then one 64-byte, then one 32-byte chunk.
    const uint32_t row_bytes = (uint32_t)(n_tile * esize);
    const uint32_t n128 = row_bytes >> 7, has64 = (row_bytes >> 6) & 1u;
    const uint32_t row128 = slab0 + (uint32_t)lane * 128u, sw128 = (uint32_t)(lane & 7);
    const uint32_t row64 = slab0 + n128 * 4096u + (uint32_t)lane * 64u, sw64 = (uint32_t)((lane >> 1) & 3);
    const uint32_t row32 = slab0 + n128 * 4096u + has64 * 2048u + (uint32_t)lane * 32u, sw32 = (uint32_t)((lane >> 2) & 1);
    auto unit_base = [&](uint32_t b) -> uint32_t {
        if (b < (n128 << 7)) return row128 + (b >> 7) * 4096u + ((((b & 127u) >> 4) ^ sw128) << 4);
        const uint32_t rem = b - (n128 << 7);
        if (has64 && rem < 64u) return row64 + (((rem >> 4) ^ sw64) << 4);
        return row32 + (sw32 << 4);                     // a 32-byte chunk holds exactly one bf16 unit (piece0 = 0)
    };
    const uint32_t ubytes = 16u * esize;               // bytes of one unit in a row
    uint32_t rphase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const TileCoord tc = decode_tile(p, t);
        const int ch_base = tc.nt * n_tile;
        const int cx = tc.x0, cy = tc.y0 + yq, cn = tc.n0 + nq;
        const uint32_t sboff = (it & 1) ? slab_stride : 0u;
        const uint32_t slab = slab0 + sboff;
        if (issuer) {
            // the stores that last read this slab have drained it (with two slabs the previous tile's may still be in flight)
            if (slab_stride) tma_store_wait_read1(); else tma_store_wait_read();
            if (RES) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar), "r"(32u * (uint32_t)(n_tile * 2)) : "memory");
#pragma unroll 1
                for (int k = 0; k < nchunks; ++k) {
                    const EpiChunk ck = p.epi[k];
                    tma_load_4d_s(&p.tmR[ck.map], rbar, slab + ck.off, ch_base + ck.col0, cx, cy, cn);
                }
            }
        }
        pair_sync(q);                                    // slab is free (and the residual load is in flight)
        if (warp == 4 && lane == 0) trace(p, 2, it, 0);
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        if (warp == 4 && lane == 0) trace(p, 2, it, 1);
        if (RES) {
            mbar_wait(&res_bar[q], rphase);
            rphase ^= 1;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * n_tile + half * 16);
        const uint32_t baddr = bias_base + (uint32_t)(ch_base + half * 16) * 4u;
        uint32_t rbuf[2][16];
        if (my_units > 0) tmem_ld16(taddr, rbuf[0]);
#pragma unroll 1
        for (int i = 0; i < my_units; i += 2) {          // two units per trip: the TMEM load of the next overlaps the math of this one
            tmem_ld_wait();
            if (i + 1 < my_units) tmem_ld16(taddr + (i + 1) * 32, rbuf[1]);
            epi_unit<ACT, RES, F32>(rbuf[0], baddr + i * 128, unit_base((uint32_t)(half + 2 * i) * ubytes) + sboff);
            if (i + 1 < my_units) {
                tmem_ld_wait();
                if (i + 2 < my_units) tmem_ld16(taddr + (i + 2) * 32, rbuf[0]);
                epi_unit<ACT, RES, F32>(rbuf[1], baddr + (i + 1) * 128, unit_base((uint32_t)(half + 2 * i + 2) * ubytes) + sboff);
            }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);                    // accumulator stage free: all tcgen05.ld of this tile have completed
        fence_proxy_async();                             // generic-proxy slab writes -> visible to the TMA store
        if (warp == 4 && lane == 0) trace(p, 2, it, 2);
        pair_sync(q);
        if (issuer) {
#pragma unroll 1
            for (int kk = 0; kk < nchunks; ++kk) {
                const EpiChunk ck = p.epi[kk];
                tma_store_4d(&p.tmO[ck.map], slab + ck.off, ch_base + ck.col0, cx, cy, cn);
            }
            tma_store_commit();
        }
        if (warp == 4 && lane == 0) trace(p, 2, it, 3);
    }
    if (issuer) tma_store_wait_all();
}

template <int ACT, int RES, int F32>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p, int nimg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A | B)] then barriers
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    uint8_t* stg_base = smem + p.stg_off;             // epilogue staging, 4 warps x 32 rows x n_tile columns
    uint64_t* full_bar = (uint64_t*)(smem + p.bar_off);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* res_bar = tempty_bar + 6;               // (+2..+5: halo barriers in the halo kernel)
    uint32_t* tmem_slot = (uint32_t*)(res_bar + 4);
    float* bias_s = (float*)(tmem_slot + 4);          // [n_tile * n_tiles_n] padded bias

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles_m = p.tiles_x * p.tiles_y * ((nimg + p.bn - 1) / p.bn);
    const int total_tiles = tiles_m * p.n_tiles_n;
    const int ksteps = p.taps * p.chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        if (p.stride == 2) {
            tma_prefetch_desc(&p.tmA[1]);
            tma_prefetch_desc(&p.tmA[2]);
            tma_prefetch_desc(&p.tmA[3]);
        }
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 256);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_tile * p.n_tiles_n; i += kThreads) bias_s[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Both control loops run warp-uniformly (all 32 lanes take every branch) and elect one lane only
    // around the asynchronous instructions: addresses and coordinates then live in uniform registers
    // and each k-step is a few dozen cycles of issue instead of a long per-thread dependent chain.
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        const int pad = p.ksz >> 1;
        const int nstages = p.stages, taps = p.taps, chunks = p.chunks, ksz = p.ksz, n_tile = p.n_tile;
        const int cin_pad = chunks * 64;
        const uint32_t tx_bytes = (uint32_t)(kTileM * 64 * 2) + p.b_tx_bytes;
        const uint32_t a_bytes = p.a_bytes;
        const bool s2 = (p.stride == 2);
        int pit = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++pit) {
            const TileCoord tc = decode_tile(p, t);
            const int bn0 = tc.nt * n_tile;
            int kh = 0, kw = 0;
            trace(p, 0, pit, 0);
            for (int tap = 0; tap < taps; ++tap) {
                const CUtensorMap* mapA = &p.tmA[0];
                int cx = tc.x0 + kw - pad, cy = tc.y0 + kh - pad;
                if (s2) {
                    // input pixel = 2*o + d, d in {-1,0,1}: odd phase for d = +-1, even for 0
                    const int dy = kh - 1, dx = kw - 1;
                    const int py = dy & 1, px = dx & 1;
                    mapA = &p.tmA[py * 2 + px];
                    cx = tc.x0 + (dx - px) / 2;
                    cy = tc.y0 + (dy - py) / 2;
                }
                const int kb = tap * cin_pad;
                for (int ch = 0; ch < chunks; ++ch) {
                    mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes, fb = full_u32 + stage * 8;
                        mbar_expect_tx_u32(fb, tx_bytes);
                        tma_load_4d_s(mapA, fb, sa, ch * 64, cx, cy, tc.n0);
                        tma_load_2d_s(&p.tmB, fb, sa + a_bytes, kb + ch * 64, bn0);
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (++kw == ksz) { kw = 0; ++kh; }
            }
            trace(p, 0, pit, 1);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        // descriptor = {lo: start>>4 | LBO<<16, hi: SBO | version | layout}; only `lo` moves
        const uint32_t hi = (uint32_t)(make_smem_desc(0, p.swizzle_bytes) >> 32);
        const uint32_t lo_base = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t stage_units = stage_bytes >> 4, a_units = p.a_bytes >> 4;
        const uint32_t idesc = p.idesc;
        const int nstages = p.stages, n_tile = p.n_tile, chunks = p.chunks;
        const int last_kmmas = (p.cin - (chunks - 1) * 64 + 15) >> 4;      // K=16 MMAs that carry data in the last chunk
        int chk = 0;
        const uint32_t tfull_u32 = smem_u32(tfull_bar), tempty_u32 = smem_u32(tempty_bar);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            trace(p, 1, it, 0);
            mbar_wait_u32(tempty_u32 + as * 8, aphase ^ 1);
            tc_fence_after();
            trace(p, 1, it, 1);
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * n_tile);
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait_u32(full_u32 + stage * 8, phase);
                tc_fence_after();
                if (ks == 0) trace(p, 1, it, 2);
                const bool last_chunk = (++chk == chunks);
                if (last_chunk) chk = 0;
                if (elect_one()) {
                    const uint32_t a_lo = lo_base + (uint32_t)stage * stage_units;
                    const uint32_t b_lo = a_lo + a_units;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // kc = 64 -> four K=16 MMAs, +32 B each; the zero-padded tail of the last chunk is skipped
                        if (k == 0 || !last_chunk || k < last_kmmas) {
                            const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + 2 * k);
                            const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_lo + 2 * k);
                            umma_bf16(d_tmem, ad, bd, idesc, (k == 0) ? (uint32_t)(ks != 0) : 1u);
                        }
                    }
                    umma_commit_u32(empty_u32 + stage * 8);                       // frees the smem slot when these MMAs retire
                    if (ks == ksteps - 1) umma_commit_u32(tfull_u32 + as * 8);     // accumulator complete
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
            trace(p, 1, it, 3);
        }
    } else if (warp >= 4) {
        epilogue_loop<ACT, RES, F32>(p, total_tiles, tmem_base, bias_s, tfull_bar, tempty_bar, res_bar, stg_base, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}


// ---------------------------------------------------------------------------------------------
// Halo variant for 3x3 stride-1 convs on 8 x 16 pixel tiles: the (8+2) x (16+2) input halo of a
// 64-channel chunk is fetched ONCE (180 rows) instead of nine shifted 128-row boxes, and the nine
// taps are nine MMAs over the same shared-memory tile whose A descriptor starts at halo row
// kh*10 + kw with a stride of 10 rows between 8-row groups (one output row of 8 pixels each).
// Weights stream tap by tap through the stage ring.  A-side TMA requests drop 6.4x.
// ---------------------------------------------------------------------------------------------
constexpr int kHaloW = 10, kHaloH = 18;
constexpr uint32_t kHaloBytes = 23 * 1024;       // 180 rows x 128 B = 23040, padded to 1 KiB

template <int ACT, int RES, int F32>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_halo_kernel(const __grid_constant__ ConvTcParams p, int nimg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem + 2 * kHaloBytes;                           // B stages
    uint8_t* stg_base = smem + p.stg_off;
    uint64_t* full_bar = (uint64_t*)(smem + p.bar_off);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* hfull_bar = tempty_bar + 2;
    uint64_t* hempty_bar = hfull_bar + 2;
    uint64_t* res_bar = hempty_bar + 2;
    uint32_t* tmem_slot = (uint32_t*)(res_bar + 4);
    float* bias_s = (float*)(tmem_slot + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_m = p.tiles_x * p.tiles_y * nimg;                // bn == 1
    const int total_tiles = tiles_m * p.n_tiles_n;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA[1]); tma_prefetch_desc(&p.tmB); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 256);
            mbar_init(&hfull_bar[i], 1); mbar_init(&hempty_bar[i], 1);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_tile * p.n_tiles_n; i += kThreads) bias_s[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // warp-uniform control loops, one elected lane around the asynchronous instructions (see conv_tc_kernel)
    const uint32_t smem_a = smem_u32(smem), smem_b = smem_u32(ring);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    const uint32_t hfull_u32 = smem_u32(hfull_bar), hempty_u32 = smem_u32(hempty_bar);
    if (warp == 0) {
        int stage = 0, hb = 0;
        uint32_t phase = 0, hphase = 0;
        const int nstages = p.stages, chunks = p.chunks, n_tile = p.n_tile;
        const int cin_pad = chunks * 64;
        const uint32_t b_tx = p.b_tx_bytes, b_bytes = p.b_bytes;
        int pit = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++pit) {
            const TileCoord tc = decode_tile(p, t);
            const int bn0 = tc.nt * n_tile;
            trace(p, 0, pit, 0);
            for (int ch = 0; ch < chunks; ++ch) {
                mbar_wait_u32(hempty_u32 + hb * 8, hphase ^ 1);
                if (ch == 0) trace(p, 0, pit, 1);
                if (elect_one()) {
                    mbar_expect_tx_u32(hfull_u32 + hb * 8, (uint32_t)(kHaloW * kHaloH * 128));
                    tma_load_4d_s(&p.tmA[1], hfull_u32 + hb * 8, smem_a + hb * kHaloBytes, ch * 64, tc.x0 - 1, tc.y0 - 1, tc.n0);
                }
                if (++hb == 2) { hb = 0; hphase ^= 1; }
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx_u32(full_u32 + stage * 8, b_tx);
                        tma_load_2d_s(&p.tmB, full_u32 + stage * 8, smem_b + (uint32_t)stage * b_bytes, tap * cin_pad + ch * 64, bn0);
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
            trace(p, 0, pit, 2);
        }
    } else if (warp == 1) {
        int stage = 0, hb = 0, it = 0;
        uint32_t phase = 0, hphase = 0;
        // B: canonical SW128 K-major, 8-row groups 1024 B apart.  A: 8-row groups one halo row (10 px) apart.
        const uint32_t hi_b = (uint32_t)(make_smem_desc(0, 128) >> 32);
        const uint32_t hi_a0 = (hi_b & ~0x3FFFu) | (uint32_t)((kHaloW * 128) >> 4);
        const uint32_t b_units = p.b_bytes >> 4;
        const uint32_t idesc = p.idesc;
        const int nstages = p.stages, n_tile = p.n_tile, chunks = p.chunks;
        const bool use_base_offset = (p.halo & 2) != 0;
        const uint32_t tfull_u32 = smem_u32(tfull_bar), tempty_u32 = smem_u32(tempty_bar);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            trace(p, 1, it, 0);
            mbar_wait_u32(tempty_u32 + as * 8, aphase ^ 1);
            tc_fence_after();
            trace(p, 1, it, 1);
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * n_tile);
            for (int ch = 0; ch < chunks; ++ch) {
                mbar_wait_u32(hfull_u32 + hb * 8, hphase);
                tc_fence_after();
                if (ch == 0) trace(p, 1, it, 2);
                const uint32_t a_base = smem_a + hb * kHaloBytes;
                const int kmmas = (ch == chunks - 1) ? ((p.cin - ch * 64 + 15) >> 4) : 4;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait_u32(full_u32 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = a_base + (uint32_t)(((tap / 3) * kHaloW + (tap % 3)) * 128);
                        uint32_t hi_a = hi_a0;
                        if (use_base_offset) hi_a |= ((a_addr >> 7) & 7u) << 17;      // descriptor bits [49,52)
                        const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16);
                        const uint32_t b_lo = (((smem_b & 0x3FFFFu) >> 4) | (1u << 16)) + (uint32_t)stage * b_units;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k == 0 || k < kmmas) {                            // skip the zero-padded tail of the last chunk
                                const uint64_t ad = ((uint64_t)hi_a << 32) | (uint64_t)(a_lo + 2 * k);
                                const uint64_t bd = ((uint64_t)hi_b << 32) | (uint64_t)(b_lo + 2 * k);
                                umma_bf16(d_tmem, ad, bd, idesc, (k == 0) ? (uint32_t)((ch | tap) != 0) : 1u);
                            }
                        }
                        umma_commit_u32(empty_u32 + stage * 8);
                        if (tap == 8) {
                            umma_commit_u32(hempty_u32 + hb * 8);                     // halo tile free once its 36 MMAs retire
                            if (ch == chunks - 1) umma_commit_u32(tfull_u32 + as * 8);
                        }
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (++hb == 2) { hb = 0; hphase ^= 1; }
            }
            trace(p, 1, it, 3);
        }
    } else if (warp >= 4) {
        epilogue_loop<ACT, RES, F32>(p, total_tiles, tmem_base, bias_s, tfull_bar, tempty_bar, res_bar, stg_base, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ---- host side ----------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    }
    return fn;
}

int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, uint32_t swizzle_bytes, bool f32 = false) {
    auto enc = get_encode();
    B2D_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2D_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu box %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return 0;
}

uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
    return (uint16_t)u;
}

// pick (bw, bh, bn), bw*bh*bn == 128, maximising useful rows; ties -> squarer spatial box, smaller bn
void pick_tile(int W, int H, int N, int* bw, int* bh, int* bn) {
    double best = -1.0;
    int best_halo = 1 << 30;
    for (int w = 1; w <= 32; w *= 2)
        for (int h = 1; w * h <= 128; h *= 2) {
            int n = 128 / (w * h);
            long long tiles = (long long)ceil_div(W, w) * ceil_div(H, h) * ceil_div(N, n);
            double eff = (double)W * H * N / (double)(tiles * 128);
            int halo = (w + 2) * (h + 2) * n;
            if (eff > best + 1e-9 || (eff > best - 1e-9 && halo < best_halo)) {
                best = eff; best_halo = halo; *bw = w; *bh = h; *bn = n;
            }
        }
}

}  // namespace

int conv_tc_supported(int cin, int ksz, int stride) {
    if (cin % 8 != 0) return 0;
    if (!((ksz == 1 && stride == 1) || (ksz == 3 && (stride == 1 || stride == 2)))) return 0;
    return 1;
}

int conv_tc_plan(ConvTcPlan* plan, int sm_count, int max_batch, const __nv_bfloat16* src, int src_h, int src_w, int src_cs,
                 int src_c0, int cin, void* dst, int dst_h, int dst_w, int dst_cs, int dst_c0, int cout, int dst_f32, int ksz,
                 int stride, int act, const float* w_host, const float* b_host, const __nv_bfloat16* res, int res_cs,
                 int res_c0) {
    memset(plan, 0, sizeof(*plan));
    B2D_CHECK(conv_tc_supported(cin, ksz, stride), "conv_tc: unsupported shape cin=%d k=%d s=%d", cin, ksz, stride);
    B2D_CHECK(src_cs % 8 == 0 && src_c0 % 8 == 0, "conv_tc: source slice must be 16-byte aligned (cs=%d c0=%d)", src_cs, src_c0);
    B2D_CHECK(stride == 1 || (src_h % 2 == 0 && src_w % 2 == 0), "conv_tc: stride-2 input must have even size");
    ConvTcParams& p = plan->p;
    plan->sm_count = sm_count;
    p.W = dst_w; p.H = dst_h; p.cin = cin; p.cout = cout;
    p.ksz = ksz; p.taps = ksz * ksz; p.stride = stride;
    p.act = act; p.out_f32 = dst_f32;
    p.out = dst; p.out_cs = dst_cs; p.out_c0 = dst_c0;
    p.res = res; p.res_cs = res_cs; p.res_c0 = res_c0;
    p.kc = 64;                          // always a full 128-byte swizzle row; a short last chunk is
    p.chunks = ceil_div(cin, 64);       // zero-filled by TMA (A) and zero-padded in the packed weights (B)
    p.swizzle_bytes = p.kc * 2;
    const int cout_pad = ceil_div(cout, 16) * 16;
    int split = 1;
    while (cout_pad % split != 0 || (cout_pad / split) % 16 != 0 || cout_pad / split > 256) ++split;
    p.n_tile = cout_pad / split;
    p.n_tiles_n = split;
    pick_tile(dst_w, dst_h, max_batch, &p.bw, &p.bh, &p.bn);
    p.tiles_x = ceil_div(dst_w, p.bw);
    p.tiles_y = ceil_div(dst_h, p.bh);
    p.a_bytes = kTileM * p.kc * 2;
    p.b_tx_bytes = p.n_tile * p.kc * 2;
    p.b_bytes = (p.b_tx_bytes + 1023u) & ~1023u;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    // ---- epilogue staging: a row of n_tile columns cut into 128 / 64 / 32-byte chunks ----
    const int esize = dst_f32 ? 4 : 2;
    B2D_CHECK(!(dst_f32 && res), "conv_tc: residual with fp32 output is not supported");
    B2D_CHECK(((size_t)dst_cs * esize) % 16 == 0 && ((size_t)dst_c0 * esize) % 16 == 0,
              "conv_tc: destination slice must be 16-byte aligned (cs=%d c0=%d)", dst_cs, dst_c0);
    B2D_CHECK(!res || (res_cs % 8 == 0 && res_c0 % 8 == 0), "conv_tc: residual slice must be 16-byte aligned");
    B2D_CHECK(p.bw <= 32 && 32 % p.bw == 0, "conv_tc: tile width %d does not divide a warp's 32 rows", p.bw);
    const uint32_t row_bytes = (uint32_t)p.n_tile * esize;
    uint32_t stg_bytes = 128u * row_bytes;
    p.stg_bufs = 1;
    {   // a second staging slab hides the TMA store's smem read behind the next tile's math, if the ring keeps >= 4 stages
        const uint32_t sb = p.a_bytes + p.b_bytes;
        const char* env = getenv("B2D_STG2");
        const int want = env ? atoi(env) : 1;
        if (want && (226u * 1024 - 2048 - 2 * stg_bytes - (uint32_t)cout_pad * 4) / sb >= 4) { p.stg_bufs = 2; stg_bytes *= 2; }
    }
    {
        uint32_t done = 0, off = 0;
        int n = 0;
        const uint32_t spans[3] = {128, 64, 32};
        for (int si = 0; si < 3; ++si)
            while (row_bytes - done >= spans[si]) {
                B2D_CHECK(n < kMaxEpiChunks, "conv_tc: n_tile %d needs too many epilogue chunks", p.n_tile);
                p.epi[n].col0 = (uint16_t)(done / esize);
                p.epi[n].cols = (uint16_t)(spans[si] / esize);
                p.epi[n].span = (uint16_t)spans[si];
                p.epi[n].map = (uint16_t)si;
                p.epi[n].off = off;
                off += 32u * spans[si];
                done += spans[si];
                ++n;
            }
        B2D_CHECK(done == row_bytes, "conv_tc: n_tile %d is not a multiple of 32 bytes", p.n_tile);
        p.epi_nchunks = n;
    }
    p.has_res = res ? 1 : 0;
    p.trace = nullptr;
    if (getenv("B2D_TRACE")) {
        B2D_CUDA(cudaMalloc(&plan->trace_dev, sizeof(long long) * kTraceRoles * kTraceTiles * kTraceEvents));
        p.trace = plan->trace_dev;
    }
    const uint32_t tail_bytes = 256 /*barriers + tmem slot*/ + (uint32_t)cout_pad * 4 /*bias*/;
    const uint32_t budget = 226 * 1024 - 1024 /*align slack*/ - stg_bytes - tail_bytes;
    int stages = (int)(budget / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    const int ksteps = p.taps * p.chunks;
    if (stages > ksteps * 2) stages = ksteps * 2;
    if (stages < 2) stages = 2;
    p.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.n_tile)) cols *= 2;
    p.tmem_cols = cols;
    // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    uint32_t ring_bytes = (uint32_t)stages * stage_bytes;
    // halo variant: 3x3 stride 1, one image per tile, 8-pixel-wide tiles (uniform 10-row group stride)
    p.halo = 0;
    {
        const char* env = getenv("B2D_HALO");       // 0 disables; 1 = verified variant (descriptor base_offset 0)
        const int want = env ? atoi(env) : 1;
        if (want && ksz == 3 && stride == 1 && p.bw == 8 && p.bh == 16 && p.bn == 1 && p.n_tile <= 96) {
            p.halo = want;                                       // 1: base_offset = 0, 3: base_offset from address
            int hs = (int)((budget - 2 * kHaloBytes) / p.b_bytes);
            if (hs > kMaxStages) hs = kMaxStages;
            p.stages = hs;
            ring_bytes = 2 * kHaloBytes + (uint32_t)hs * p.b_bytes;
        }
    }
    p.stg_off = ring_bytes;                                      // 1 KiB aligned: every ring slot is a multiple of 1 KiB
    p.bar_off = p.stg_off + stg_bytes;
    plan->smem_bytes = (size_t)p.bar_off + tail_bytes + 1024 /*align slack*/;
    B2D_CHECK(plan->smem_bytes <= 227 * 1024, "conv_tc: %zu bytes of shared memory needed", plan->smem_bytes);

    // ---- weights: fp32 [cout][cin][k][k] -> bf16 [cout_pad][kh][kw][cin_pad] (zero padded) ----
    const int cin_pad = p.chunks * 64;
    const size_t ktot = (size_t)p.taps * cin_pad;
    std::vector<uint16_t> wp((size_t)cout_pad * ktot, 0);
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
            for (int t = 0; t < p.taps; ++t)
                wp[(size_t)o * ktot + (size_t)t * cin_pad + c] = f2bf(w_host[((size_t)o * cin + c) * p.taps + t]);
    std::vector<float> bp(cout_pad, 0.f);
    for (int o = 0; o < cout; ++o) bp[o] = b_host ? b_host[o] : 0.f;
    B2D_CUDA(cudaMalloc(&plan->w_dev, wp.size() * 2));
    B2D_CUDA(cudaMemcpy(plan->w_dev, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
    B2D_CUDA(cudaMalloc(&plan->bias_dev, bp.size() * 4));
    B2D_CUDA(cudaMemcpy(plan->bias_dev, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    p.bias = plan->bias_dev;

    // ---- tensor maps ----
    {
        uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)cout_pad};
        uint64_t str[1] = {(uint64_t)ktot * 2};
        uint32_t box[2] = {(uint32_t)p.kc, (uint32_t)p.n_tile};
        if (encode_map(&p.tmB, plan->w_dev, 2, dims, str, box, p.swizzle_bytes)) return -1;
    }
    const uint32_t boxA[4] = {(uint32_t)p.kc, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (stride == 1) {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)src_w, (uint64_t)src_h, (uint64_t)max_batch};
        uint64_t str[3] = {(uint64_t)src_cs * 2, (uint64_t)src_w * src_cs * 2, (uint64_t)src_h * src_w * src_cs * 2};
        if (encode_map(&p.tmA[0], (void*)(src + src_c0), 4, dims, str, boxA, p.swizzle_bytes)) return -1;
    } else {
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                uint64_t dims[4] = {(uint64_t)cin, (uint64_t)src_w / 2, (uint64_t)src_h / 2, (uint64_t)max_batch};
                uint64_t str[3] = {(uint64_t)2 * src_cs * 2, (uint64_t)2 * src_w * src_cs * 2,
                                   (uint64_t)src_h * src_w * src_cs * 2};
                const __nv_bfloat16* base = src + ((size_t)py * src_w + px) * src_cs + src_c0;
                if (encode_map(&p.tmA[py * 2 + px], (void*)base, 4, dims, str, boxA, p.swizzle_bytes)) return -1;
            }
    }
    if (p.halo) {
        const uint32_t boxH[4] = {64u, (uint32_t)kHaloW, (uint32_t)kHaloH, 1u};
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)src_w, (uint64_t)src_h, (uint64_t)max_batch};
        uint64_t str[3] = {(uint64_t)src_cs * 2, (uint64_t)src_w * src_cs * 2, (uint64_t)src_h * src_w * src_cs * 2};
        if (encode_map(&p.tmA[1], (void*)(src + src_c0), 4, dims, str, boxH, 128)) return -1;
    }
    {   // output / residual sub-box maps: one warp's 32 rows of the tile, one map per chunk width
        const int box_px = p.bw * p.bh;
        const uint32_t sbh = (uint32_t)(box_px >= 32 ? 32 / p.bw : p.bh);
        const uint32_t sbn = (uint32_t)(box_px >= 32 ? 1 : 32 / box_px);
        const uint32_t spans[3] = {128, 64, 32};
        bool used[3] = {false, false, false};
        for (int k = 0; k < p.epi_nchunks; ++k) used[p.epi[k].map] = true;
        for (int si = 0; si < 3; ++si) {
            if (!used[si]) continue;
            const uint32_t box[4] = {spans[si] / (uint32_t)esize, (uint32_t)p.bw, sbh, sbn};
            uint64_t dims[4] = {(uint64_t)cout, (uint64_t)dst_w, (uint64_t)dst_h, (uint64_t)max_batch};
            uint64_t str[3] = {(uint64_t)dst_cs * esize, (uint64_t)dst_w * dst_cs * esize, (uint64_t)dst_h * dst_w * dst_cs * esize};
            if (encode_map(&p.tmO[si], (uint8_t*)dst + (size_t)dst_c0 * esize, 4, dims, str, box, spans[si], dst_f32 != 0)) return -1;
            if (res) {
                uint64_t rstr[3] = {(uint64_t)res_cs * 2, (uint64_t)dst_w * res_cs * 2, (uint64_t)dst_h * dst_w * res_cs * 2};
                if (encode_map(&p.tmR[si], (void*)(res + res_c0), 4, dims, rstr, box, spans[si], false)) return -1;
            }
        }
    }
    return 0;
}

namespace {
typedef void (*ConvKernel)(const ConvTcParams, int);
// (halo, act, res, f32) -> instantiation; fp32 outputs never carry a residual
ConvKernel pick_kernel(int halo, int act, int res, int f32) {
#define B2D_PICK(K)                                                        \
    if (f32) return act ? K<1, 0, 1> : K<0, 0, 1>;                         \
    if (res) return act ? K<1, 1, 0> : K<0, 1, 0>;                         \
    return act ? K<1, 0, 0> : K<0, 0, 0>;
    if (halo) { B2D_PICK(conv_tc_halo_kernel) }
    B2D_PICK(conv_tc_kernel)
#undef B2D_PICK
}
}  // namespace

int conv_tc_launch(const ConvTcPlan* plan, int n, cudaStream_t stream) {
    const ConvTcParams& p = plan->p;
    const int tiles = p.tiles_x * p.tiles_y * ceil_div(n, p.bn) * p.n_tiles_n;
    int grid = tiles < plan->sm_count ? tiles : plan->sm_count;
    if (grid < 1) return 0;
    ConvKernel k = pick_kernel(p.halo, p.act, p.has_res, p.out_f32);
    static bool attr_done[2][2][2][2];
    bool& done = attr_done[p.halo ? 1 : 0][p.act ? 1 : 0][p.has_res ? 1 : 0][p.out_f32 ? 1 : 0];
    if (!done) {
        B2D_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        done = true;
    }
    if (p.trace) B2D_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * kTraceRoles * kTraceTiles * kTraceEvents, stream));
    k<<<grid, kThreads, plan->smem_bytes, stream>>>(p, n);
    if (p.trace && getenv("B2D_TRACE_DUMP")) {
        static long long h[kTraceRoles * kTraceTiles * kTraceEvents];
        B2D_CUDA(cudaStreamSynchronize(stream));
        B2D_CUDA(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
        long long t0 = h[0] ? h[0] : h[kTraceTiles * kTraceEvents];
        const char* names[kTraceRoles] = {"load", "mma ", "epi "};
        fprintf(stderr, "trace (cycles from first stamp; load: first issue, last issue | mma: start, tmem free, first operand, last commit | epi: slab free, acc ready, math done, store issued)\n");
        for (int it = 0; it < kTraceTiles; ++it) {
            fprintf(stderr, "tile %2d", it);
            for (int r = 0; r < kTraceRoles; ++r) {
                fprintf(stderr, " | %s", names[r]);
                for (int e = 0; e < kTraceEvents; ++e) {
                    long long v = h[(r * kTraceTiles + it) * kTraceEvents + e];
                    if (v) fprintf(stderr, " %7lld", v - t0); else fprintf(stderr, "       -");
                }
            }
            fprintf(stderr, "\n");
        }
    }
    B2D_LAUNCH_CHECK();
    return 0;
}

void conv_tc_free(ConvTcPlan* plan) {
    if (plan->w_dev) cudaFree(plan->w_dev);
    if (plan->bias_dev) cudaFree(plan->bias_dev);
    if (plan->trace_dev) cudaFree(plan->trace_dev);
    plan->trace_dev = nullptr;
    plan->w_dev = nullptr;
    plan->bias_dev = nullptr;
}

int conv_tc_describe(const ConvTcPlan* plan, char* buf, int buflen) {
    const ConvTcParams& p = plan->p;
    return snprintf(buf, buflen,
                    "tcgen05%s conv k%d s%d cin %d cout %d -> %dx%d | tile %dx%dx%d n_tile %d x%d kc %d (SW%u) stages %d tmem %u smem %zu", p.halo ? "-halo" : "",
                    p.ksz, p.stride, p.cin, p.cout, p.H, p.W, p.bw, p.bh, p.bn, p.n_tile, p.n_tiles_n, p.kc, p.swizzle_bytes,
                    p.stages, p.tmem_cols, plan->smem_bytes);
}
