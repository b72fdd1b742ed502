// Test-time-augmentation variants of the reference, on uint8 RGB tiles resident in HBM:
//
//   CLAHE on L of LAB    cv2.cvtColor(RGB2LAB) -> createCLAHE(clip, grid).apply(l) -> merge -> cvtColor(LAB2RGB)
//                        _script/gpu_handler.py:104-110 (3.0, 8x8), :128-136 (4.0, 4x4); gpu_handler_archive.py:100-117
//   brightness / gamma   PIL ImageEnhance.Brightness.enhance(f) (:113-115), np.power(img/255, 1/gamma)*255 (:118-121):
//                        both are functions of one byte -> a 256-entry table built by the caller (tta.py), applied here
//   contrast             PIL ImageEnhance.Contrast.enhance(f) (gpu_handler_archive.py:82): per-image grey mean, then a table
//
// All of it is HBM-bound byte / integer work: bit-exact against OpenCV / Pillow (tests/test_gpu_tta.py), no tensor cores.
// The Lab tables are restated from OpenCV's published algorithm (tools/gen_lab_tables.py -> lab_tables.h).
//
// Pixels are processed in quads (4 px = 12 bytes = three 32-bit words); a batch [n, h, w, 3] is one contiguous run of
// pixels, so only the last (n*h*w mod 4) pixels take the byte path.
#include "common.cuh"
#include "lab_tables.h"

namespace {

struct LabTabs {                 // 11 776 bytes, copied into shared memory by every CTA that converts colours
    uint16_t gamma[256];         // sRGB byte -> linear, 3 fractional bits
    uint16_t cbrt[LAB_CBRT_TAB_SIZE];
    uint16_t l2y[256], l2fy[256];
    uint8_t inv_gamma[LAB_INV_GAMMA_TAB_SIZE];
};
static_assert(sizeof(LabTabs) % 16 == 0, "LabTabs is copied in 16-byte pieces");
struct LabCoef { int f[9]; int i[9]; };

__device__ LabTabs g_tabs;
bool g_tabs_ready[64] = {};

int ensure_tabs() {
    int dev = 0;
    B2D_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && g_tabs_ready[dev]) return 0;
    static LabTabs h;
    for (int k = 0; k < 256; ++k) { h.gamma[k] = kSrgbGammaTab[k]; h.l2y[k] = kLabToY[k]; h.l2fy[k] = kLabToFy[k]; }
    for (int k = 0; k < LAB_CBRT_TAB_SIZE; ++k) h.cbrt[k] = kLabCbrtTab[k];
    for (int k = 0; k < LAB_INV_GAMMA_TAB_SIZE; ++k) h.inv_gamma[k] = kSrgbInvGammaTab[k];
    B2D_CUDA(cudaMemcpyToSymbol(g_tabs, &h, sizeof(h)));
    if (dev < 64) g_tabs_ready[dev] = true;
    return 0;
}

LabCoef coefs() {
    LabCoef c;
    for (int k = 0; k < 9; ++k) { c.f[k] = kRgb2XyzCoeffs[k]; c.i[k] = kXyz2RgbCoeffs[k]; }
    return c;
}

__device__ __forceinline__ void load_tabs(LabTabs* s) {
    const uint4* g = reinterpret_cast<const uint4*>(&g_tabs);
    uint4* d = reinterpret_cast<uint4*>(s);
    for (int k = threadIdx.x; k < (int)(sizeof(LabTabs) / 16); k += blockDim.x) d[k] = g[k];
    __syncthreads();
}

__device__ __forceinline__ int sat_u8(int v) { return min(max(v, 0), 255); }

// ---- RGB -> Lab (8-bit, OpenCV's fixed-point path: shifts 12 / 3 / 15) ------------------------------------------------
__device__ __forceinline__ int lab_fy(const LabTabs* T, const LabCoef& C, int R, int G, int B) {
    return T->cbrt[(R * C.f[3] + G * C.f[4] + B * C.f[5] + (1 << 11)) >> 12];
}
__device__ __forceinline__ int lab_l_from_fy(int fY) {
    constexpr int kLscale = (116 * 255 + 50) / 100;
    constexpr int kLshift = -((16 * 255 * (1 << 15) + 50) / 100);
    return sat_u8((kLscale * fY + kLshift + (1 << 14)) >> 15);
}
__device__ __forceinline__ void rgb2lab_px(const LabTabs* T, const LabCoef& C, int r, int g, int b, int& L, int& A, int& Bo) {
    const int R = T->gamma[r], G = T->gamma[g], B = T->gamma[b];
    const int fX = T->cbrt[(R * C.f[0] + G * C.f[1] + B * C.f[2] + (1 << 11)) >> 12];
    const int fY = lab_fy(T, C, R, G, B);
    const int fZ = T->cbrt[(R * C.f[6] + G * C.f[7] + B * C.f[8] + (1 << 11)) >> 12];
    L = lab_l_from_fy(fY);
    A = sat_u8((500 * (fX - fY) + 128 * (1 << 15) + (1 << 14)) >> 15);
    Bo = sat_u8((200 * (fY - fZ) + 128 * (1 << 15) + (1 << 14)) >> 15);
}

// ---- Lab -> RGB (8-bit, OpenCV's integer path: base 2^14) ---------------------------------------------------------------
__device__ __forceinline__ int ab_to_xz(int v) {
    constexpr int kBase = 1 << 14;
    // C integer division (truncation toward zero), as the library's table initialiser
    return v <= 3390 ? v * 108 / 841 - kBase * 16 / 116 * 108 / 841 : (v * v / kBase) * v / kBase;
}
__device__ __forceinline__ void lab2rgb_px(const LabTabs* T, const LabCoef& C, int L, int a, int b, int& r, int& g, int& bo) {
    constexpr int kBase = 1 << 14;
    const int y = T->l2y[L], ify = T->l2fy[L];
    const int adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * kBase / 500;
    const int bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * kBase / 200 + 1;
    const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
    const int ro = (C.i[0] * x + C.i[1] * y + C.i[2] * z + (1 << 13)) >> 14;
    const int go = (C.i[3] * x + C.i[4] * y + C.i[5] * z + (1 << 13)) >> 14;
    const int bb = (C.i[6] * x + C.i[7] * y + C.i[8] * z + (1 << 13)) >> 14;
    r = T->inv_gamma[min(max(ro, 0), LAB_INV_GAMMA_TAB_SIZE - 1)];
    g = T->inv_gamma[min(max(go, 0), LAB_INV_GAMMA_TAB_SIZE - 1)];
    bo = T->inv_gamma[min(max(bb, 0), LAB_INV_GAMMA_TAB_SIZE - 1)];
}

// ---- quad load / store ----------------------------------------------------------------------------------------------------
struct Quad { uint32_t w[3]; };
__device__ __forceinline__ int quad_byte(const Quad& q, int k) { return (q.w[k >> 2] >> ((k & 3) * 8)) & 255; }
__device__ __forceinline__ void quad_set(Quad& q, int k, int v) { q.w[k >> 2] |= (uint32_t)v << ((k & 3) * 8); }

struct ClaheGeom {
    int h, w, tw, th, tiles_x, tiles_y;
    float inv_tw, inv_th;
    const uint8_t* luts;     // [n][tiles_y][tiles_x][256]
};

enum { OP_RGB2LAB = 0, OP_LAB2RGB = 1, OP_CLAHE = 2 };

// cv2's CLAHE_Interpolation_Body for one pixel: bilinear blend of the four surrounding tiles' tables (fp32, no contraction)
struct PixPos { int img, y, x; };
__device__ __forceinline__ PixPos pix_pos(const ClaheGeom& g, unsigned int pix) {      // 32-bit: npix < 2^31 is checked on the host
    const unsigned int hw = (unsigned int)(g.h * g.w);
    PixPos p;
    p.img = (int)(pix / hw);
    const unsigned int rem = pix - (unsigned int)p.img * hw;
    p.y = (int)(rem / (unsigned int)g.w);
    p.x = (int)(rem - (unsigned int)p.y * (unsigned int)g.w);
    return p;
}
__device__ __forceinline__ void pix_next(const ClaheGeom& g, PixPos& p) {
    if (++p.x == g.w) { p.x = 0; if (++p.y == g.h) { p.y = 0; ++p.img; } }
}
__device__ __forceinline__ int clahe_interp(const ClaheGeom& g, const PixPos& p, int L) {
    const float txf = __fsub_rn(__fmul_rn((float)p.x, g.inv_tw), 0.5f);
    const float tyf = __fsub_rn(__fmul_rn((float)p.y, g.inv_th), 0.5f);
    int tx1 = __float2int_rd(txf), ty1 = __float2int_rd(tyf);
    const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
    const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
    const int tx2 = min(tx1 + 1, g.tiles_x - 1), ty2 = min(ty1 + 1, g.tiles_y - 1);
    tx1 = max(tx1, 0); ty1 = max(ty1, 0);
    const uint8_t* lp = g.luts + (size_t)(p.img * g.tiles_y * g.tiles_x) * 256 + L;
    const float l11 = (float)__ldg(lp + (ty1 * g.tiles_x + tx1) * 256), l12 = (float)__ldg(lp + (ty1 * g.tiles_x + tx2) * 256);
    const float l21 = (float)__ldg(lp + (ty2 * g.tiles_x + tx1) * 256), l22 = (float)__ldg(lp + (ty2 * g.tiles_x + tx2) * 256);
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    return sat_u8(__float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya))));
}

template <int OP>
__device__ __forceinline__ void map_px(const LabTabs* T, const LabCoef& C, const ClaheGeom& g, const PixPos& pos, int c0, int c1, int c2,
                                       int& o0, int& o1, int& o2) {
    if (OP == OP_RGB2LAB) {
        rgb2lab_px(T, C, c0, c1, c2, o0, o1, o2);
    } else if (OP == OP_LAB2RGB) {
        lab2rgb_px(T, C, c0, c1, c2, o0, o1, o2);
    } else {
        int L, a, b;
        rgb2lab_px(T, C, c0, c1, c2, L, a, b);
        lab2rgb_px(T, C, clahe_interp(g, pos, L), a, b, o0, o1, o2);
    }
}

template <int OP>
__global__ void __launch_bounds__(256) pixel_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long npix, LabCoef C,
                                                    ClaheGeom g) {
    __shared__ __align__(16) LabTabs T;
    load_tabs(&T);
    const long long quads = npix >> 2;
    const long long step = (long long)gridDim.x * blockDim.x;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += step) {
        Quad in, out;
        in.w[0] = __ldg(s32 + q * 3); in.w[1] = __ldg(s32 + q * 3 + 1); in.w[2] = __ldg(s32 + q * 3 + 2);
        out.w[0] = out.w[1] = out.w[2] = 0;
        PixPos pos{0, 0, 0};
        if (OP == OP_CLAHE) pos = pix_pos(g, (unsigned int)(q * 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int o0, o1, o2;
            map_px<OP>(&T, C, g, pos, quad_byte(in, 3 * k), quad_byte(in, 3 * k + 1), quad_byte(in, 3 * k + 2), o0, o1, o2);
            quad_set(out, 3 * k, o0); quad_set(out, 3 * k + 1, o1); quad_set(out, 3 * k + 2, o2);
            if (OP == OP_CLAHE) pix_next(g, pos);
        }
        d32[q * 3] = out.w[0]; d32[q * 3 + 1] = out.w[1]; d32[q * 3 + 2] = out.w[2];
    }
    // the last npix mod 4 pixels
    if (blockIdx.x == 0 && threadIdx.x < (int)(npix & 3)) {
        const long long p = (quads << 2) + threadIdx.x;
        int o0, o1, o2;
        PixPos pos{0, 0, 0};
        if (OP == OP_CLAHE) pos = pix_pos(g, (unsigned int)p);
        map_px<OP>(&T, C, g, pos, src[p * 3], src[p * 3 + 1], src[p * 3 + 2], o0, o1, o2);
        dst[p * 3] = (uint8_t)o0; dst[p * 3 + 1] = (uint8_t)o1; dst[p * 3 + 2] = (uint8_t)o2;
    }
}

// ---- CLAHE tables: one CTA per (tile, image) -- cv2's CLAHE_CalcLut_Body ---------------------------------------------------
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ src, int h, int w, int tw, int th, int tiles_x,
                                                        int tiles_y, int limit, float lut_scale, LabCoef C, uint8_t* __restrict__ luts) {
    __shared__ uint16_t s_gamma[256];
    __shared__ uint16_t s_cbrt[LAB_CBRT_TAB_SIZE];
    __shared__ unsigned int hist[8][256];
    __shared__ int warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_gamma[tid] = g_tabs.gamma[tid];
    for (int k = tid; k < LAB_CBRT_TAB_SIZE; k += 256) s_cbrt[k] = g_tabs.cbrt[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[k][tid] = 0;
    __syncthreads();
    const int tile = blockIdx.x, img = blockIdx.y;
    const int tyi = tile / tiles_x, txi = tile - tyi * tiles_x;
    const uint8_t* base = src + (size_t)img * h * w * 3;
    const int area = tw * th;
    if (((w | tw) & 3) == 0 && (txi + 1) * tw <= w) {
        // quad path: tile rows start on a 4-pixel boundary and lie inside the image in x, so four pixels are three aligned
        // 32-bit words (the rows below the image are reflected as whole rows)
        const int tq = tw >> 2, quads = tq * th;
        const int dq = 256 / tq, dr = 256 - dq * tq;
        int yy = tid / tq, xq = tid - yy * tq;
        for (int q = tid; q < quads; q += 256, yy += dq, xq += dr) {
            if (xq >= tq) { xq -= tq; ++yy; }
            int Y = tyi * th + yy;
            if (Y >= h) Y = 2 * (h - 1) - Y;
            const uint32_t* p = reinterpret_cast<const uint32_t*>(base + ((size_t)Y * w + txi * tw + 4 * xq) * 3);
            Quad in;
            in.w[0] = __ldg(p); in.w[1] = __ldg(p + 1); in.w[2] = __ldg(p + 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int R = s_gamma[quad_byte(in, 3 * k)], G = s_gamma[quad_byte(in, 3 * k + 1)], B = s_gamma[quad_byte(in, 3 * k + 2)];
                const int L = lab_l_from_fy(s_cbrt[(R * C.f[3] + G * C.f[4] + B * C.f[5] + (1 << 11)) >> 12]);
                atomicAdd(&hist[warp][L], 1u);
            }
        }
    } else {
        const int dq = 256 / tw, dr = 256 - dq * tw;       // idx += 256 as (row, column) increments: no division per pixel
        int yy = tid / tw, xx = tid - yy * tw;
        for (int idx = tid; idx < area; idx += 256, yy += dq, xx += dr) {
            if (xx >= tw) { xx -= tw; ++yy; }
            int Y = tyi * th + yy, X = txi * tw + xx;
            if (Y >= h) Y = 2 * (h - 1) - Y;            // BORDER_REFLECT_101 of the ragged bottom / right edge
            if (X >= w) X = 2 * (w - 1) - X;
            const uint8_t* p = base + ((size_t)Y * w + X) * 3;
            const int R = s_gamma[__ldg(p)], G = s_gamma[__ldg(p + 1)], B = s_gamma[__ldg(p + 2)];
            const int L = lab_l_from_fy(s_cbrt[(R * C.f[3] + G * C.f[4] + B * C.f[5] + (1 << 11)) >> 12]);
            atomicAdd(&hist[warp][L], 1u);
        }
    }
    __syncthreads();
    int hv = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) hv += (int)hist[k][tid];
    if (limit > 0) {
        int excess = max(hv - limit, 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, o);
        if (lane == 0) warp_tot[warp] = excess;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) clipped += warp_tot[k];
        __syncthreads();
        const int batch = clipped / 256, resid = clipped - batch * 256;
        hv = min(hv, limit) + batch;
        if (resid != 0) {
            const int stepr = max(256 / resid, 1);
            if (tid % stepr == 0 && tid / stepr < resid) ++hv;
        }
    }
    // inclusive prefix sum over the 256 bins
    int sum = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sum, o);
        if (lane >= o) sum += t;
    }
    if (lane == 31) warp_tot[warp] = sum;
    __syncthreads();
    for (int k = 0; k < warp; ++k) sum += warp_tot[k];
    luts[(((size_t)img * tiles_y + tyi) * tiles_x + txi) * 256 + tid] = (uint8_t)sat_u8(__float2int_rn(__fmul_rn((float)sum, lut_scale)));
}

// ---- per-byte tables -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lut_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long img_bytes,
                                                  const uint8_t* __restrict__ lut, int per_image) {
    __shared__ uint8_t s[256];
    const int img = blockIdx.y;
    s[threadIdx.x] = lut[(per_image ? (size_t)img * 256 : 0) + threadIdx.x];
    __syncthreads();
    const uint8_t* sp = src + (size_t)img * img_bytes;
    uint8_t* dp = dst + (size_t)img * img_bytes;
    const long long step = (long long)gridDim.x * blockDim.x;
    if ((img_bytes & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(sp);
        uint4* d4 = reinterpret_cast<uint4*>(dp);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (img_bytes >> 4); i += step) {
            const uint4 v = __ldg(s4 + i);
            uint32_t in[4] = {v.x, v.y, v.z, v.w}, o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = (uint32_t)s[in[k] & 255] | ((uint32_t)s[(in[k] >> 8) & 255] << 8) | ((uint32_t)s[(in[k] >> 16) & 255] << 16) |
                       ((uint32_t)s[in[k] >> 24] << 24);
            d4[i] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < img_bytes; i += step) dp[i] = s[sp[i]];
    }
}

// ---- contrast: PIL's grey mean (ImageEnhance.Contrast.__init__), then Image.blend(mean, image, factor) as a table -----------
__global__ void __launch_bounds__(256) grey_sum_kernel(const uint8_t* __restrict__ src, long long img_pixels,
                                                       unsigned long long* __restrict__ sums) {
    const int img = blockIdx.y;
    const uint8_t* sp = src + (size_t)img * img_pixels * 3;
    const long long step = (long long)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    const bool vec = ((img_pixels * 3) & 3) == 0;
    if (vec) {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sp);
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (img_pixels >> 2); q += step) {
            Quad in;
            in.w[0] = __ldg(s32 + q * 3); in.w[1] = __ldg(s32 + q * 3 + 1); in.w[2] = __ldg(s32 + q * 3 + 2);
#pragma unroll
            for (int k = 0; k < 4; ++k)      // ImagingConvert rgb2l: (R*19595 + G*38470 + B*7471 + 0x8000) >> 16
                acc += (quad_byte(in, 3 * k) * 19595 + quad_byte(in, 3 * k + 1) * 38470 + quad_byte(in, 3 * k + 2) * 7471 + 0x8000) >> 16;
        }
    } else {
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < img_pixels; p += step)
            acc += (sp[p * 3] * 19595 + sp[p * 3 + 1] * 38470 + sp[p * 3 + 2] * 7471 + 0x8000) >> 16;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sums[img], acc);
}

__global__ void __launch_bounds__(256) blend_lut_kernel(const unsigned long long* __restrict__ sums, long long img_pixels, float alpha,
                                                        uint8_t* __restrict__ luts) {
    const int img = blockIdx.x, v = threadIdx.x;
    // int(stat.mean[0] + 0.5): double division, double add, truncation
    const int mean = (int)__dadd_rn(__ddiv_rn((double)sums[img], (double)img_pixels), 0.5);
    // ImagingBlend: in1 + alpha * (in2 - in1) in float32, separate multiply and add; clip; truncate
    const float t = __fadd_rn((float)mean, __fmul_rn(alpha, (float)(v - mean)));
    luts[(size_t)img * 256 + v] = t <= 0.f ? 0 : (t >= 255.f ? 255 : (uint8_t)t);
}

int grid_for(long long items, int per_block) {
    long long b = (items + per_block - 1) / per_block;
    const long long cap = 148 * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int tta_colour_launch(const uint8_t* src, long long npix, int code, uint8_t* dst, cudaStream_t stream) {
    if (npix <= 0) return 0;
    if (ensure_tabs()) return -2;
    ClaheGeom g{};
    const int grid = grid_for(npix >> 2, 256);
    if (code == 0) pixel_kernel<OP_RGB2LAB><<<grid, 256, 0, stream>>>(src, dst, npix, coefs(), g);
    else pixel_kernel<OP_LAB2RGB><<<grid, 256, 0, stream>>>(src, dst, npix, coefs(), g);
    B2D_LAUNCH_CHECK();
    return 0;
}

int tta_clahe_launch(const uint8_t* src, int n, int h, int w, double clip_limit, int tiles_x, int tiles_y, uint8_t* luts, uint8_t* dst,
                     cudaStream_t stream) {
    if (n <= 0) return 0;
    if (ensure_tabs()) return -2;
    // CLAHE_Impl::apply: a ragged image is extended (reflect-101) on the bottom and right by tiles - size % tiles on BOTH axes
    int eh = h, ew = w;
    if (h % tiles_y != 0 || w % tiles_x != 0) { eh = h + (tiles_y - h % tiles_y); ew = w + (tiles_x - w % tiles_x); }
    const int tw = ew / tiles_x, th = eh / tiles_y;
    B2D_CHECK(eh - h < h && ew - w < w, "clahe: image %dx%d too small for a %dx%d grid", w, h, tiles_x, tiles_y);
    const int area = tw * th;
    const float lut_scale = 255.0f / (float)area;
    int limit = 0;
    if (clip_limit > 0.0) { limit = (int)(clip_limit * area / 256); if (limit < 1) limit = 1; }
    const LabCoef C = coefs();
    clahe_lut_kernel<<<dim3(tiles_x * tiles_y, n), 256, 0, stream>>>(src, h, w, tw, th, tiles_x, tiles_y, limit, lut_scale, C, luts);
    B2D_LAUNCH_CHECK();
    ClaheGeom g{h, w, tw, th, tiles_x, tiles_y, 1.0f / (float)tw, 1.0f / (float)th, luts};
    const long long npix = (long long)n * h * w;
    B2D_CHECK(npix < (1ll << 31), "clahe: %lld pixels in one call (limit 2^31)", npix);
    pixel_kernel<OP_CLAHE><<<grid_for(npix >> 2, 256), 256, 0, stream>>>(src, dst, npix, C, g);
    B2D_LAUNCH_CHECK();
    return 0;
}

int tta_lut_launch(const uint8_t* src, int n, long long img_bytes, const uint8_t* lut, int per_image, uint8_t* dst, cudaStream_t stream) {
    if (n <= 0 || img_bytes <= 0) return 0;
    int gx = grid_for(img_bytes >> 4, 256 * 4);
    if (gx * n > 148 * 16) gx = (148 * 16 + n - 1) / n;
    lut_kernel<<<dim3(gx, n), 256, 0, stream>>>(src, dst, img_bytes, lut, per_image);
    B2D_LAUNCH_CHECK();
    return 0;
}

int tta_contrast_launch(const uint8_t* src, int n, int h, int w, float factor, unsigned long long* sums, uint8_t* luts, uint8_t* dst,
                        cudaStream_t stream) {
    if (n <= 0) return 0;
    const long long px = (long long)h * w;
    B2D_CUDA(cudaMemsetAsync(sums, 0, (size_t)n * sizeof(unsigned long long), stream));
    int gx = grid_for(px >> 2, 256 * 4);
    if (gx * n > 148 * 16) gx = (148 * 16 + n - 1) / n;
    grey_sum_kernel<<<dim3(gx, n), 256, 0, stream>>>(src, px, sums);
    B2D_LAUNCH_CHECK();
    blend_lut_kernel<<<n, 256, 0, stream>>>(sums, px, factor, luts);
    B2D_LAUNCH_CHECK();
    return tta_lut_launch(src, n, px * 3, luts, 1, dst, stream);
}
