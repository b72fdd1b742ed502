// The non-convolution ops of the detector graphs (K3 in SURVEY.md section 2.2):
//   * max-pool k x k (SPPF 5/1, SPPCSPC 5-9-13/1, YOLOv7 MP 2/2);
//   * nearest 2x upsample written straight into the consumer's concat slice.
// Every convolution -- stem and depthwise included -- runs on the tensor cores (conv_tc.cu); there is no CUDA-core
// convolution in this library, and a shape conv_tc.cu cannot run is a planning error, not a slower path.
// All are HBM/L2-bound elementwise-style kernels: NHWC bf16, 16-byte vector accesses along
// the channel axis, one thread per (pixel, 8-channel group).
#include "common.cuh"


namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f[2 * j] = __uint_as_float(w[j] << 16);
        f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        w[j] = *(uint32_t*)&h;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// the same for fp16 activations (B2D_PREC_FP16)
__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 v = __half22float2(*(const __half2*)&w[j]);
        f[2 * j] = v.x;
        f[2 * j + 1] = v.y;
    }
}
__device__ __forceinline__ uint4 pack8h(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
        w[j] = *(uint32_t*)&h;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- max-pool k x k, padding k/2 for stride 1 / none for stride 2 (-inf padding) -------------
__global__ void __launch_bounds__(256) maxpool_kernel(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0,
                                                       __nv_bfloat16* dst, int oh, int ow, int dst_cs, int dst_c0, int c,
                                                       int k, int stride, int n, int f16) {
    const int groups = c / 8;
    const long long total = (long long)n * oh * ow * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const long long pix = idx / groups;
    const int x = (int)(pix % ow);
    const int y = (int)((pix / ow) % oh);
    const int img = (int)(pix / ((long long)ow * oh));
    const int pad = (stride == 1) ? (k >> 1) : 0;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kh = 0; kh < k; ++kh) {
        const int iy = y * stride + kh - pad;
        if (iy < 0 || iy >= h) continue;
        for (int kw = 0; kw < k; ++kw) {
            const int ix = x * stride + kw - pad;
            if (ix < 0 || ix >= w) continue;
            const uint4 u = __ldg((const uint4*)(src + (((long long)img * h + iy) * w + ix) * src_cs + src_c0 + g * 8));
            float a[8];
            if (f16) unpack8h(u, a); else unpack8(u, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], a[j]);
        }
    }
    *(uint4*)(dst + pix * dst_cs + dst_c0 + g * 8) = f16 ? pack8h(m) : pack8(m);
}

// Split-fp16 storage (B2D_PREC_FP16X2): 8 channels = 32 bytes [hi x 8 | lo x 8]; the value is hi + lo (exact in fp32), the
// winner's pair is copied unchanged.  thread = (output pixel, 8-channel group).
__global__ void __launch_bounds__(256) maxpool_x2_kernel(const __half* src, int h, int w, int src_cs, int src_c0, __half* dst, int oh, int ow,
                                                          int dst_cs, int dst_c0, int c, int k, int stride, int n) {
    const int groups = c / 8;
    const long long total = (long long)n * oh * ow * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const long long pix = idx / groups;
    const int x = (int)(pix % ow);
    const int y = (int)((pix / ow) % oh);
    const int img = (int)(pix / ((long long)ow * oh));
    const int pad = (stride == 1) ? (k >> 1) : 0;
    float m[8];
    uint32_t bh[4] = {0, 0, 0, 0}, bl[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kh = 0; kh < k; ++kh) {
        const int iy = y * stride + kh - pad;
        if (iy < 0 || iy >= h) continue;
        for (int kw = 0; kw < k; ++kw) {
            const int ix = x * stride + kw - pad;
            if (ix < 0 || ix >= w) continue;
            const uint4* sp = (const uint4*)(src + ((((long long)img * h + iy) * w + ix) * src_cs + src_c0) * 2 + g * 16);
            const uint4 uh = __ldg(sp), ul = __ldg(sp + 1);
            float a[8], b[8];
            unpack8h(uh, a);
            unpack8h(ul, b);
            const uint32_t wh[4] = {uh.x, uh.y, uh.z, uh.w}, wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = a[j] + b[j];
                if (v > m[j]) {
                    m[j] = v;
                    const uint32_t mask = (j & 1) ? 0xFFFF0000u : 0x0000FFFFu;
                    bh[j >> 1] = (bh[j >> 1] & ~mask) | (wh[j >> 1] & mask);
                    bl[j >> 1] = (bl[j >> 1] & ~mask) | (wl[j >> 1] & mask);
                }
            }
        }
    }
    uint4* dp = (uint4*)(dst + (pix * dst_cs + dst_c0) * 2 + g * 16);
    dp[0] = make_uint4(bh[0], bh[1], bh[2], bh[3]);
    dp[1] = make_uint4(bl[0], bl[1], bl[2], bl[3]);
}

// thread = (source pixel, 8-channel group): one 16-byte load, four 16-byte stores
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0,
                                                          __nv_bfloat16* dst, int dst_cs, int dst_c0, int c, int n) {
    const int groups = c / 8;
    const long long total = (long long)n * h * w * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const long long pix = idx / groups;
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const long long img = pix / ((long long)w * h);
    const uint4 u = __ldg((const uint4*)(src + pix * src_cs + src_c0 + g * 8));
    __nv_bfloat16* o = dst + ((img * 2 * h + 2 * y) * (2 * w) + 2 * x) * dst_cs + dst_c0 + g * 8;
    const long long row = (long long)2 * w * dst_cs;
    *(uint4*)o = u;
    *(uint4*)(o + dst_cs) = u;
    *(uint4*)(o + row) = u;
    *(uint4*)(o + row + dst_cs) = u;
}

// ---- SPPF / SPPCSPC pooling chain: mp5, mp5(mp5) = mp9, mp5(mp5(mp5)) = mp13 in one pass --------------
// One CTA per (image, group of CG channels): the whole H x W plane of those channels lives in shared
// memory; each stage is a separable 5-max (row pass into a temporary, column pass back), written to its
// destination slice and kept for the next stage.  max() is exact on bf16, so packed __hmax2 is used.
template <int CG>
__global__ void __launch_bounds__(256) poolchain_kernel(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0,
                                                         __nv_bfloat16* d0, __nv_bfloat16* d1, __nv_bfloat16* d2, int dst_cs,
                                                         int c0_0, int c0_1, int c0_2, int c, int stages, int f16) {
    extern __shared__ uint4 pool_smem[];
    constexpr int V = CG / 8;                          // 16-byte vectors per pixel
    uint4* cur = pool_smem;
    uint4* tmp = pool_smem + (size_t)h * w * V;
    const int groups = c / CG;
    const int g = blockIdx.x % groups;
    const long long img = blockIdx.x / groups;
    const int items = h * w * V;
    const __nv_bfloat16* sp = src + img * h * w * src_cs + src_c0 + g * CG;
    for (int i = threadIdx.x; i < items; i += blockDim.x) cur[i] = __ldg((const uint4*)(sp + (long long)(i / V) * src_cs) + (i % V));
    __syncthreads();
    auto vmax = [f16](uint4 a, uint4 b) {
        uint4 r;
        if (f16) {
            __half2 t;
            t = __hmax2(*(__half2*)&a.x, *(__half2*)&b.x); r.x = *(uint32_t*)&t;
            t = __hmax2(*(__half2*)&a.y, *(__half2*)&b.y); r.y = *(uint32_t*)&t;
            t = __hmax2(*(__half2*)&a.z, *(__half2*)&b.z); r.z = *(uint32_t*)&t;
            t = __hmax2(*(__half2*)&a.w, *(__half2*)&b.w); r.w = *(uint32_t*)&t;
            return r;
        }
        __nv_bfloat162 t;
        t = __hmax2(*(__nv_bfloat162*)&a.x, *(__nv_bfloat162*)&b.x); r.x = *(uint32_t*)&t;
        t = __hmax2(*(__nv_bfloat162*)&a.y, *(__nv_bfloat162*)&b.y); r.y = *(uint32_t*)&t;
        t = __hmax2(*(__nv_bfloat162*)&a.z, *(__nv_bfloat162*)&b.z); r.z = *(uint32_t*)&t;
        t = __hmax2(*(__nv_bfloat162*)&a.w, *(__nv_bfloat162*)&b.w); r.w = *(uint32_t*)&t;
        return r;
    };
    for (int s = 0; s < stages; ++s) {
        for (int i = threadIdx.x; i < items; i += blockDim.x) {          // row pass
            const int v = i % V, px = i / V, x = px % w, y = px / w;
            uint4 m = cur[i];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && x + d >= 0 && x + d < w) m = vmax(m, cur[(y * w + x + d) * V + v]);
            tmp[i] = m;
        }
        __syncthreads();
        __nv_bfloat16* dp = (s == 0 ? d0 + c0_0 : s == 1 ? d1 + c0_1 : d2 + c0_2) + img * h * w * dst_cs + g * CG;
        for (int i = threadIdx.x; i < items; i += blockDim.x) {          // column pass
            const int v = i % V, px = i / V, x = px % w, y = px / w;
            uint4 m = tmp[i];
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && y + d >= 0 && y + d < h) m = vmax(m, tmp[((y + d) * w + x) * V + v]);
            cur[i] = m;
            *((uint4*)(dp + (long long)px * dst_cs) + v) = m;
        }
        __syncthreads();
    }
}

}  // namespace

int maxpool_launch(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0, __nv_bfloat16* dst, int oh, int ow,
                   int dst_cs, int dst_c0, int c, int k, int stride, int n, cudaStream_t stream, int f16, int x2) {
    B2D_CHECK(c % 8 == 0 && src_cs % 8 == 0 && src_c0 % 8 == 0 && dst_cs % 8 == 0 && dst_c0 % 8 == 0,
              "maxpool: channel counts/offsets must be multiples of 8");
    const long long total = (long long)n * oh * ow * (c / 8);
    if (x2) {       // channel arguments are real channels; the buffers hold 2 x 16 bits per channel
        maxpool_x2_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>((const __half*)src, h, w, src_cs, src_c0, (__half*)dst, oh, ow, dst_cs,
                                                                          dst_c0, c, k, stride, n);
        B2D_LAUNCH_CHECK();
        return 0;
    }
    maxpool_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, h, w, src_cs, src_c0, dst, oh, ow, dst_cs, dst_c0, c, k,
                                                                   stride, n, f16);
    B2D_LAUNCH_CHECK();
    return 0;
}

int upsample2x_launch(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0, __nv_bfloat16* dst, int dst_cs,
                      int dst_c0, int c, int n, cudaStream_t stream) {
    B2D_CHECK(c % 8 == 0 && src_cs % 8 == 0 && src_c0 % 8 == 0 && dst_cs % 8 == 0 && dst_c0 % 8 == 0,
              "upsample: channel counts/offsets must be multiples of 8");
    const long long total = (long long)n * h * w * (c / 8);
    upsample2x_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, h, w, src_cs, src_c0, dst, dst_cs, dst_c0, c, n);
    B2D_LAUNCH_CHECK();
    return 0;
}

// The pool chain keeps an H x W plane of 8 channels twice in shared memory (row-pass temporary + current stage).
int poolchain_fits(int h, int w) { return (size_t)h * w * 8 * 2 * 2 <= 200 * 1024; }

// Stride-1 5x5 max-pool applied `stages` (<= 3) times in a chain; stage s is written to dst[s] at channel offset c0[s].
int poolchain_launch(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0, __nv_bfloat16* const* dst, const int* dst_c0,
                     int dst_cs, int c, int stages, int n, cudaStream_t stream, int f16) {
    B2D_CHECK(stages >= 1 && stages <= 3, "poolchain: 1..3 stages");
    B2D_CHECK(c % 8 == 0 && src_cs % 8 == 0 && src_c0 % 8 == 0 && dst_cs % 8 == 0, "poolchain: channel counts/offsets must be multiples of 8");
    const int cg = (c % 32 == 0 && (size_t)h * w * 32 * 2 * 2 <= 96 * 1024) ? 32 : 8;
    const size_t smem = (size_t)h * w * cg * 2 * 2;
    B2D_CHECK(poolchain_fits(h, w), "poolchain: a %dx%d plane does not fit in shared memory (plan_finalize must not fuse it)", h, w);
    __nv_bfloat16* d[3];
    int c0[3];
    for (int i = 0; i < 3; ++i) { d[i] = dst[i < stages ? i : stages - 1]; c0[i] = dst_c0[i < stages ? i : stages - 1]; }
    const int blocks = n * (c / cg);
    if (cg == 32) {
        if (b2d_func_smem_optin((const void*)poolchain_kernel<32>, 200 * 1024)) return -2;
        poolchain_kernel<32><<<blocks, 256, smem, stream>>>(src, h, w, src_cs, src_c0, d[0], d[1], d[2], dst_cs, c0[0], c0[1], c0[2], c, stages, f16);
    } else {
        if (b2d_func_smem_optin((const void*)poolchain_kernel<8>, 200 * 1024)) return -2;
        poolchain_kernel<8><<<blocks, 256, smem, stream>>>(src, h, w, src_cs, src_c0, d[0], d[1], d[2], dst_cs, c0[0], c0[1], c0[2], c, stages, f16);
    }
    B2D_LAUNCH_CHECK();
    return 0;
}
