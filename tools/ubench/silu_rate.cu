// Epilogue math microbenchmark: outputs/clk/SM of the bias + SiLU + bf16-pack sequence on register data, for
// different warp counts per SM and SiLU formulations.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ex2(float y) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }
__device__ __forceinline__ float rcpa(float y) { float e; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }
__device__ __forceinline__ float tanha(float y) { float e; asm("tanh.approx.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }

template <int MODE>
__device__ __forceinline__ uint64_t silu2(uint64_t v) {
    if (MODE == 0) {          // product: 1 MUFU + 2 Newton, packed
        float y0, y1, n0, n1;
        upk2(mul2(v, pk2(-1.4426950408889634f, -1.4426950408889634f)), y0, y1);
        y0 = fminf(y0, 64.0f); y1 = fminf(y1, 64.0f);
        const uint64_t nd = sub2(pk2(-1.0f, -1.0f), pk2(ex2(y0), ex2(y1)));
        upk2(nd, n0, n1);
        uint64_t r = pk2u(0xFEF311C7u - __float_as_uint(n0), 0xFEF311C7u - __float_as_uint(n1));
        const uint64_t one = pk2(1.0f, 1.0f);
        r = fma2(r, fma2(nd, r, one), r);
        r = fma2(r, fma2(nd, r, one), r);
        return mul2(v, r);
    } else if (MODE == 1) {   // 2 MUFU (ex2 + rcp)
        float y0, y1, d0, d1;
        upk2(mul2(v, pk2(-1.4426950408889634f, -1.4426950408889634f)), y0, y1);
        upk2(add2(pk2(1.0f, 1.0f), pk2(ex2(y0), ex2(y1))), d0, d1);
        return mul2(v, pk2(rcpa(d0), rcpa(d1)));
    } else if (MODE == 2) {   // tanh form: 1 MUFU, 2 packed FMA-pipe ops
        float h0, h1;
        const uint64_t h = mul2(v, pk2(0.5f, 0.5f));
        upk2(h, h0, h1);
        return fma2(h, pk2(tanha(h0), tanha(h1)), h);
    } else if (MODE == 3) {   // no MUFU at all: Newton only (pipe isolation; wrong values)
        float n0, n1;
        const uint64_t nd = sub2(pk2(-1.0f, -1.0f), mul2(v, v));
        upk2(nd, n0, n1);
        uint64_t r = pk2u(0xFEF311C7u - __float_as_uint(n0), 0xFEF311C7u - __float_as_uint(n1));
        const uint64_t one = pk2(1.0f, 1.0f);
        r = fma2(r, fma2(nd, r, one), r);
        r = fma2(r, fma2(nd, r, one), r);
        return mul2(v, r);
    } else if (MODE == 4) {   // MUFU only
        float y0, y1;
        upk2(v, y0, y1);
        return pk2(ex2(y0), ex2(y1));
    } else {                  // scalar product form (before packing)
        float a[2]; upk2(v, a[0], a[1]);
        for (int i = 0; i < 2; ++i) {
            float y = fminf(a[i] * -1.4426950408889634f, 64.0f), e = ex2(y);
            const float d = 1.0f + e;
            float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
            r = fmaf(r, fmaf(-d, r, 1.0f), r);
            r = fmaf(r, fmaf(-d, r, 1.0f), r);
            a[i] *= r;
        }
        return pk2(a[0], a[1]);
    }
}

template <int MODE, int NP>
__global__ void k(const float* in, uint32_t* out, int iters, long long* cyc) {
    uint64_t v[NP], b[NP];
    for (int i = 0; i < NP; ++i) { v[i] = pk2(in[threadIdx.x + i], in[threadIdx.x + i + 1]); b[i] = pk2(in[i], in[i + 7]); }
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            uint64_t x = silu2<MODE>(add2(v[i], b[i]));
            float lo, hi; upk2(x, lo, hi);
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            acc ^= *(uint32_t*)&h;
            v[i] = add2(v[i], pk2(0.001f, 0.002f));   // keep the inputs changing (one extra packed op per pair)
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int NP>
void run(const char* name, int threads, const float* in, uint32_t* out, long long* cyc) {
    const int iters = 2000;
    k<MODE, NP><<<148, threads>>>(in, out, iters, cyc);
    k<MODE, NP><<<148, threads>>>(in, out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s pairs/thread %2d threads %4d: %6.2f outputs/clk/SM\n", name, NP, threads, (double)threads * NP * 2 * iters / h);
}

int main() {
    float* in; uint32_t* out; long long* cyc;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i % 37) * 0.1f - 1.5f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int threads : {256, 512}) {
        run<0, 8>("product packed (1 MUFU + Newton)", threads, in, out, cyc);
        run<0, 16>("product packed (1 MUFU + Newton)", threads, in, out, cyc);
        run<5, 8>("scalar (1 MUFU + Newton)", threads, in, out, cyc);
        run<1, 8>("ex2 + rcp (2 MUFU)", threads, in, out, cyc);
        run<2, 8>("tanh form (1 MUFU)", threads, in, out, cyc);
        run<3, 8>("Newton only (no MUFU)", threads, in, out, cyc);
        run<4, 8>("MUFU.EX2 only", threads, in, out, cyc);
    }
    return 0;
}
