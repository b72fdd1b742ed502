// Accuracy of SiLU formulations after bf16 rounding, against a double-precision reference.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float y) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }
__device__ __forceinline__ float rcpa(float y) { float e; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }
__device__ __forceinline__ float tanha(float y) { float e; asm("tanh.approx.f32 %0, %1;" : "=f"(e) : "f"(y)); return e; }
__global__ void k(const float* v, float* o, int n, int mode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = v[i], r;
    if (mode == 0) { float y = fminf(x * -1.4426950408889634f, 64.0f), e = ex2(y); const float d = 1.0f + e; float q = __int_as_float(0x7EF311C7 - __float_as_int(d));
                     q = fmaf(q, fmaf(-d, q, 1.0f), q); q = fmaf(q, fmaf(-d, q, 1.0f), q); r = x * q; }
    else if (mode == 1) { float e = ex2(x * -1.4426950408889634f); r = x * rcpa(1.0f + e); }
    else { float h = 0.5f * x; r = fmaf(h, tanha(h), h); }
    o[i] = r;
}
static float bf16r(float f) { uint32_t u; memcpy(&u, &f, 4); u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u; float r; memcpy(&r, &u, 4); return r; }
int main() {
    const int n = 1 << 22;
    std::vector<float> h(n), o(n);
    for (int i = 0; i < n; ++i) h[i] = -16.0f + 32.0f * (i + 0.37f) / n;
    float *dv, *dout; cudaMalloc(&dv, n * 4); cudaMalloc(&dout, n * 4);
    cudaMemcpy(dv, h.data(), n * 4, cudaMemcpyHostToDevice);
    const char* names[3] = {"ex2 + Newton (product)", "ex2 + rcp.approx", "tanh.approx form"};
    for (int mode = 0; mode < 3; ++mode) {
        k<<<n / 256, 256>>>(dv, dout, n, mode);
        cudaMemcpy(o.data(), dout, n * 4, cudaMemcpyDeviceToHost);
        double maxabs = 0, maxrel = 0, maxabs_neg = 0, maxrel_pos = 0; long mism = 0, mism2 = 0; double at = 0;
        for (int i = 0; i < n; ++i) {
            double x = h[i], ref = x / (1.0 + exp(-x));
            double ae = fabs(o[i] - ref), re = ae / fmax(fabs(ref), 1e-30);
            if (ae > maxabs) { maxabs = ae; at = x; }
            if (x < 0 && ae > maxabs_neg) maxabs_neg = ae;
            if (x > 0 && re > maxrel_pos) maxrel_pos = re;
            if (fabs(x) < 8 && re > maxrel) maxrel = re;
            float a = bf16r(o[i]), b = bf16r((float)ref);
            if (a != b) { ++mism; if (fabs(a - b) > 1.01 * fabs(bf16r(b * 1.00390625f + 1e-30f) - b) && fabs((double)a - b) > 1e-6) ++mism2; }
        }
        printf("%-24s max abs err %.3e (at v=%.3f), max abs err v<0 %.3e, max rel err v>0 %.3e, max rel err |v|<8 %.3e, bf16 results differing from the correctly rounded: %.4f%% (more than 1 ulp: %.4f%%)\n",
               names[mode], maxabs, at, maxabs_neg, maxrel_pos, maxrel, 100.0 * mism / n, 100.0 * mism2 / n);
    }
    return 0;
}
