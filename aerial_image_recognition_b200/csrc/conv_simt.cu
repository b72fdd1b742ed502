// CUDA-core kernels of the detector graph (K3 in SURVEY.md section 2.2):
//   * direct convolution for the stem (Cin = 3, K = 27: not a tensor-core shape) -- also the
//     in-library cross-check for the tcgen05 path (B2D_CONV_SIMT);
//   * depthwise 3x3 + bias + SiLU (Ultralytics 8.3.x cls branch);
//   * max-pool k x k (SPPF 5/1, SPPCSPC 5-9-13/1, YOLOv7 MP 2/2);
//   * nearest 2x upsample written straight into the consumer's concat slice.
// All are HBM/L2-bound elementwise-style kernels: NHWC bf16, 16-byte vector accesses along
// the channel axis, one thread per (pixel, 8-channel group).
#include "common.cuh"

#include <string.h>
#include <vector>

namespace {

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

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

// ---- direct conv: thread = (output pixel, group of 16 output channels) --------------------
constexpr int kCoutPerThread = 16;

__global__ void __launch_bounds__(128) conv_simt_kernel(ConvSimtPlan p, int n) {
    const int groups = p.cout_pad / kCoutPerThread;
    const long long total = (long long)n * p.dst_h * p.dst_w * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    long long pix = idx / groups;
    const int ox = (int)(pix % p.dst_w);
    const int oy = (int)((pix / p.dst_w) % p.dst_h);
    const int img = (int)(pix / ((long long)p.dst_w * p.dst_h));
    const int pad = p.ksz >> 1;
    float acc[kCoutPerThread];
#pragma unroll
    for (int j = 0; j < kCoutPerThread; ++j) acc[j] = 0.f;
    for (int kh = 0; kh < p.ksz; ++kh) {
        const int iy = oy * p.stride + kh - pad;
        if (iy < 0 || iy >= p.src_h) continue;
        for (int kw = 0; kw < p.ksz; ++kw) {
            const int ix = ox * p.stride + kw - pad;
            if (ix < 0 || ix >= p.src_w) continue;
            const __nv_bfloat16* sp = p.src + (((long long)img * p.src_h + iy) * p.src_w + ix) * p.src_cs + p.src_c0;
            const float* wp = p.w_dev + ((size_t)(kh * p.ksz + kw) * p.cin) * p.cout_pad + g * kCoutPerThread;
            for (int c = 0; c < p.cin; ++c) {
                const float a = __bfloat162float(sp[c]);
                const float4* w4 = (const float4*)(wp + (size_t)c * p.cout_pad);
#pragma unroll
                for (int j = 0; j < kCoutPerThread / 4; ++j) {
                    const float4 w = __ldg(w4 + j);
                    acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
                }
            }
        }
    }
    const int ch0 = g * kCoutPerThread;
    for (int j = 0; j < kCoutPerThread; ++j) {
        const int ch = ch0 + j;
        if (ch >= p.cout) break;
        float v = acc[j] + p.bias_dev[ch];
        if (p.act) v = silu_f(v);
        if (p.res) v += __bfloat162float(p.res[pix * p.res_cs + p.res_c0 + ch]);
        if (p.dst_f32) ((float*)p.dst)[pix * p.dst_cs + p.dst_c0 + ch] = v;
        else ((__nv_bfloat16*)p.dst)[pix * p.dst_cs + p.dst_c0 + ch] = __float2bfloat16_rn(v);
    }
}

// ---- stem: 3x3 stride-s conv on the NHWC4 input, all output channels per thread -----------
// Weights [tap][4][COUT] fp32 live in shared memory (broadcast reads); each thread owns one
// output pixel and COUT accumulators, reads nine 8-byte pixels and writes COUT*2 contiguous bytes.
template <int COUT>
__global__ void __launch_bounds__(128) stem_kernel(ConvSimtPlan p, int n) {
    __shared__ float ws[9 * 4 * COUT];
    __shared__ float bs[COUT];
    for (int i = threadIdx.x; i < 9 * 4 * COUT; i += blockDim.x) {
        const int tap = i / (4 * COUT), r = i % (4 * COUT), c = r / COUT, o = r % COUT;
        ws[i] = (c < p.cin) ? p.w_dev[((size_t)tap * p.cin + c) * p.cout_pad + o] : 0.f;
    }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = p.bias_dev[i];
    __syncthreads();
    const long long total = (long long)n * p.dst_h * p.dst_w;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int ox = (int)(pix % p.dst_w);
    const int oy = (int)((pix / p.dst_w) % p.dst_h);
    const int img = (int)(pix / ((long long)p.dst_w * p.dst_h));
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = bs[j];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int iy = oy * p.stride + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int ix = ox * p.stride + kw - 1;
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            if (iy >= 0 && iy < p.src_h && ix >= 0 && ix < p.src_w) {
                const uint2 u = __ldg((const uint2*)(p.src + (((long long)img * p.src_h + iy) * p.src_w + ix) * 4));
                a[0] = __uint_as_float(u.x << 16);
                a[1] = __uint_as_float(u.x & 0xFFFF0000u);
                a[2] = __uint_as_float(u.y << 16);
            }
            const float* w = ws + (kh * 3 + kw) * 4 * COUT;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < COUT; ++j) acc[j] = fmaf(a[c], w[c * COUT + j], acc[j]);
        }
    }
    __nv_bfloat16* op = (__nv_bfloat16*)p.dst + pix * p.dst_cs + p.dst_c0;
#pragma unroll
    for (int j = 0; j < COUT; j += 8) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = p.act ? silu_f(acc[j + i]) : acc[j + i];
        *(uint4*)(op + j) = pack8(f);
    }
}

// ---- depthwise 3x3, stride 1: thread = (8 channels, 4 consecutive output pixels of a row) ------
// 18 16-byte activation loads and 20 16-byte weight/bias loads produce 32 outputs, so the kernel
// is bound by HBM/L2 rather than by load-instruction issue.
constexpr int kDwPix = 4;
__global__ void __launch_bounds__(256) dwconv_kernel(DwConvPlan p, int n) {
    const int groups = p.c / 8;
    const int xt = (p.w + kDwPix - 1) / kDwPix;
    const long long total = (long long)n * p.h * xt * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    long long r = idx / groups;
    const int x0 = (int)(r % xt) * kDwPix;
    r /= xt;
    const int y = (int)(r % p.h);
    const int img = (int)(r / p.h);
    const int c0 = g * 8;
    float acc[kDwPix][8];
    {
        const float4 b0 = __ldg((const float4*)(p.bias_dev + c0)), b1 = __ldg((const float4*)(p.bias_dev + c0) + 1);
#pragma unroll
        for (int i = 0; i < kDwPix; ++i) {
            acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
            acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
        }
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int iy = y + kh - 1;
        if (iy < 0 || iy >= p.h) continue;
        float wv[3][8];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const float4* wp = (const float4*)(p.w_dev + (size_t)(kh * 3 + kw) * p.c + c0);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            wv[kw][0] = w0.x; wv[kw][1] = w0.y; wv[kw][2] = w0.z; wv[kw][3] = w0.w;
            wv[kw][4] = w1.x; wv[kw][5] = w1.y; wv[kw][6] = w1.z; wv[kw][7] = w1.w;
        }
        const __nv_bfloat16* rowp = p.src + ((long long)img * p.h + iy) * p.w * p.src_cs + p.src_c0 + c0;
#pragma unroll
        for (int j = 0; j < kDwPix + 2; ++j) {
            const int ix = x0 + j - 1;
            if (ix < 0 || ix >= p.w) continue;
            float a[8];
            unpack8(__ldg((const uint4*)(rowp + (long long)ix * p.src_cs)), a);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int i = j - kw;            // output pixel fed by this input through tap kw
                if (i < 0 || i >= kDwPix) continue;
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(a[q], wv[kw][q], acc[i][q]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kDwPix; ++i) {
        const int x = x0 + i;
        if (x >= p.w) break;
        if (p.act) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[i][q] = silu_f(acc[i][q]);
        }
        const long long pix = ((long long)img * p.h + y) * p.w + x;
        *(uint4*)(p.dst + pix * p.dst_cs + p.dst_c0 + c0) = pack8(acc[i]);
    }
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

float bf16_round(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
    memcpy(&f, &u, 4);
    return f;
}

}  // namespace

int conv_simt_plan(ConvSimtPlan* plan, const __nv_bfloat16* src, int src_h, int src_w, int src_cs, int src_c0, int cin,
                   int cin_w, void* dst, int dst_h, int dst_w, int dst_cs, int dst_c0, int cout, int dst_f32, int ksz,
                   int stride, int act, const float* w_host, const float* b_host, const __nv_bfloat16* res, int res_cs,
                   int res_c0) {
    memset(plan, 0, sizeof(*plan));
    plan->src = src; plan->src_h = src_h; plan->src_w = src_w; plan->src_cs = src_cs; plan->src_c0 = src_c0; plan->cin = cin_w;
    plan->dst = dst; plan->dst_h = dst_h; plan->dst_w = dst_w; plan->dst_cs = dst_cs; plan->dst_c0 = dst_c0;
    plan->cout = cout; plan->dst_f32 = dst_f32; plan->ksz = ksz; plan->stride = stride; plan->act = act;
    plan->res = res; plan->res_cs = res_cs; plan->res_c0 = res_c0;
    (void)cin;
    const int taps = ksz * ksz;
    const int cout_pad = ceil_div(cout, 16) * 16;
    plan->cout_pad = cout_pad;
    // cin_w = channels present in the weight tensor (3 for the stem whose buffer has 4)
    std::vector<float> wp((size_t)taps * cin_w * cout_pad, 0.f), bp(cout_pad, 0.f);
    for (int o = 0; o < cout; ++o) {
        for (int c = 0; c < cin_w; ++c)
            for (int t = 0; t < taps; ++t)
                wp[((size_t)t * cin_w + c) * cout_pad + o] = bf16_round(w_host[((size_t)o * cin_w + c) * taps + t]);
        bp[o] = b_host ? b_host[o] : 0.f;
    }
    B2D_CUDA(cudaMalloc(&plan->w_dev, wp.size() * 4));
    B2D_CUDA(cudaMemcpy(plan->w_dev, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice));
    B2D_CUDA(cudaMalloc(&plan->bias_dev, bp.size() * 4));
    B2D_CUDA(cudaMemcpy(plan->bias_dev, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

int conv_simt_launch(const ConvSimtPlan* plan, int n, cudaStream_t stream) {
    const bool stem = plan->src_cs == 4 && plan->cin == 3 && plan->ksz == 3 && !plan->dst_f32 && plan->res == nullptr &&
                      plan->dst_cs % 8 == 0 && plan->dst_c0 % 8 == 0;
    if (stem && (plan->cout == 48 || plan->cout == 32)) {
        const long long total = (long long)n * plan->dst_h * plan->dst_w;
        const int blocks = (int)((total + 127) / 128);
        if (plan->cout == 48) stem_kernel<48><<<blocks, 128, 0, stream>>>(*plan, n);
        else stem_kernel<32><<<blocks, 128, 0, stream>>>(*plan, n);
    } else {
        const long long total = (long long)n * plan->dst_h * plan->dst_w * (plan->cout_pad / kCoutPerThread);
        const int blocks = (int)((total + 127) / 128);
        conv_simt_kernel<<<blocks, 128, 0, stream>>>(*plan, n);
    }
    B2D_LAUNCH_CHECK();
    return 0;
}

void conv_simt_free(ConvSimtPlan* plan) {
    if (plan->w_dev) cudaFree(plan->w_dev);
    if (plan->bias_dev) cudaFree(plan->bias_dev);
    plan->w_dev = nullptr; plan->bias_dev = nullptr;
}

int dwconv_plan(DwConvPlan* plan, const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0, __nv_bfloat16* dst,
                int dst_cs, int dst_c0, int c, int act, const float* w_host, const float* b_host) {
    memset(plan, 0, sizeof(*plan));
    B2D_CHECK(c % 8 == 0 && src_cs % 8 == 0 && src_c0 % 8 == 0 && dst_cs % 8 == 0 && dst_c0 % 8 == 0,
              "dwconv: channel counts/offsets must be multiples of 8");
    plan->src = src; plan->h = h; plan->w = w; plan->src_cs = src_cs; plan->src_c0 = src_c0;
    plan->dst = dst; plan->dst_cs = dst_cs; plan->dst_c0 = dst_c0; plan->c = c; plan->act = act;
    std::vector<float> wp((size_t)9 * c), bp(c);
    for (int ch = 0; ch < c; ++ch) {
        for (int t = 0; t < 9; ++t) wp[(size_t)t * c + ch] = bf16_round(w_host[(size_t)ch * 9 + t]);
        bp[ch] = b_host ? b_host[ch] : 0.f;
    }
    B2D_CUDA(cudaMalloc(&plan->w_dev, wp.size() * 4));
    B2D_CUDA(cudaMemcpy(plan->w_dev, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice));
    B2D_CUDA(cudaMalloc(&plan->bias_dev, bp.size() * 4));
    B2D_CUDA(cudaMemcpy(plan->bias_dev, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

int dwconv_launch(const DwConvPlan* plan, int n, cudaStream_t stream) {
    const long long total = (long long)n * plan->h * ((plan->w + kDwPix - 1) / kDwPix) * (plan->c / 8);
    dwconv_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(*plan, n);
    B2D_LAUNCH_CHECK();
    return 0;
}

void dwconv_free(DwConvPlan* plan) {
    if (plan->w_dev) cudaFree(plan->w_dev);
    if (plan->bias_dev) cudaFree(plan->bias_dev);
    plan->w_dev = nullptr; plan->bias_dev = nullptr;
}

int maxpool_launch(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0, __nv_bfloat16* dst, int oh, int ow,
                   int dst_cs, int dst_c0, int c, int k, int stride, int n, cudaStream_t stream, int f16) {
    B2D_CHECK(c % 8 == 0 && src_cs % 8 == 0 && src_c0 % 8 == 0 && dst_cs % 8 == 0 && dst_c0 % 8 == 0,
              "maxpool: channel counts/offsets must be multiples of 8");
    const long long total = (long long)n * oh * ow * (c / 8);
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

// Stride-1 5x5 max-pool applied `stages` (<= 3) times in a chain; stage s is written to dst[s] at channel offset c0[s].
// Returns 1 (and launches nothing) when the plane does not fit in shared memory: the caller then runs the pools one by one.
int poolchain_launch(const __nv_bfloat16* src, int h, int w, int src_cs, int src_c0, __nv_bfloat16* const* dst, const int* dst_c0,
                     int dst_cs, int c, int stages, int n, cudaStream_t stream, int f16) {
    B2D_CHECK(stages >= 1 && stages <= 3, "poolchain: 1..3 stages");
    B2D_CHECK(c % 8 == 0 && src_cs % 8 == 0 && src_c0 % 8 == 0 && dst_cs % 8 == 0, "poolchain: channel counts/offsets must be multiples of 8");
    const int cg = (c % 32 == 0 && (size_t)h * w * 32 * 2 * 2 <= 96 * 1024) ? 32 : 8;
    const size_t smem = (size_t)h * w * cg * 2 * 2;
    if (smem > 200 * 1024) return 1;
    __nv_bfloat16* d[3];
    int c0[3];
    for (int i = 0; i < 3; ++i) { d[i] = dst[i < stages ? i : stages - 1]; c0[i] = dst_c0[i < stages ? i : stages - 1]; }
    const int blocks = n * (c / cg);
    if (cg == 32) {
        static bool attr32 = false;
        if (!attr32) { B2D_CUDA(cudaFuncSetAttribute(poolchain_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr32 = true; }
        poolchain_kernel<32><<<blocks, 256, smem, stream>>>(src, h, w, src_cs, src_c0, d[0], d[1], d[2], dst_cs, c0[0], c0[1], c0[2], c, stages, f16);
    } else {
        static bool attr8 = false;
        if (!attr8) { B2D_CUDA(cudaFuncSetAttribute(poolchain_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr8 = true; }
        poolchain_kernel<8><<<blocks, 256, smem, stream>>>(src, h, w, src_cs, src_c0, d[0], d[1], d[2], dst_cs, c0[0], c0[1], c0[2], c, stages, f16);
    }
    B2D_LAUNCH_CHECK();
    return 0;
}
