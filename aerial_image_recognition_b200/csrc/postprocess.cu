// K4 / K4' / K5 / K6 -- everything between the raw head maps and georeferenced detections.
//
//   decode      v8: DFL softmax-expectation + anchor decode + sigmoid (Ultralytics Detect [EXT],
//                   SURVEY.md section 8a row a5);  v7: in-graph sigmoid/grid/anchor decode (Appendix A.4).
//   filter      `boxes[boxes[:,4] >= thr]`           simple_detector.py:479-481, :672-673;
//                                                    _script/gpu_handler.py:166-170 (top-10 at :173)
//   select      per-tile ordering + (optional) greedy IoU-NMS with Ultralytics' arithmetic
//               (class-offset boxes, suppress iff IoU > thr, max_det)  -- `model(window)` at
//               x_arch/02_analyze_images:1 (cell 6)
//   georef      pixel centre -> lon/lat or CRS metres in fp64 without FMA contraction:
//               simple_detector.py:484-502, _script/gpu_handler.py:178-190, pixel_to_geo
//
// Layout: head maps are NHWC fp32 so one anchor's logits are contiguous; the decode kernels give
// each anchor to 4 lanes (one per box side) so a warp reads 8 anchors x 256 B = 2 KB contiguous.
// Candidates are compacted with warp ballots + one atomic per warp; determinism comes from the
// per-tile sort in `select` (key = confidence, anchor index), never from arrival order.
#include "common.cuh"

#include <math.h>

namespace {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- v8: one anchor handled by 4 consecutive lanes; returns the decoded row on every lane ----
struct Row { float cx, cy, w, h, conf; int cls; };

__device__ __forceinline__ float dfl_side(const float* p) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 q = __ldg((const float4*)p + j);
        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
    float m = v[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) m = fmaxf(m, v[j]);
    float s = 0.f, e = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float t = expf(v[j] - m);
        s += t;
        e += t * (float)j;
    }
    return e / s;
}

__device__ __forceinline__ void v8_scores(const float* pix, int nc, float* conf, int* cls) {
    float best = -1.f;
    int bi = 0;
    for (int k = 0; k < nc; ++k) {
        const float s = sigmoid_f(__ldg(pix + 64 + k));
        if (s > best) { best = s; bi = k; }
    }
    *conf = best; *cls = bi;
}

__device__ __forceinline__ Row v8_box(float side, int lane_side, int gx, int gy, float stride, unsigned mask, int base_lane) {
    // gather l,t,r,b from the 4 lanes of this anchor
    const float l = __shfl_sync(mask, side, base_lane + 0);
    const float t = __shfl_sync(mask, side, base_lane + 1);
    const float r = __shfl_sync(mask, side, base_lane + 2);
    const float b = __shfl_sync(mask, side, base_lane + 3);
    (void)lane_side;
    const float ax = (float)gx + 0.5f, ay = (float)gy + 0.5f;
    const float x1 = __fsub_rn(ax, l), y1 = __fsub_rn(ay, t), x2 = __fadd_rn(ax, r), y2 = __fadd_rn(ay, b);
    Row o;
    o.cx = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), stride);
    o.cy = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), stride);
    o.w = __fmul_rn(__fsub_rn(x2, x1), stride);
    o.h = __fmul_rn(__fsub_rn(y2, y1), stride);
    o.conf = 0.f; o.cls = 0;
    return o;
}

// ---- v7: one thread per (anchor box, cell) ------------------------------------------------------
__device__ __forceinline__ Row v7_row(const float* cell, int nc, int gx, int gy, float stride, float aw, float ah) {
    Row o;
    const float sx = sigmoid_f(__ldg(cell + 0)), sy = sigmoid_f(__ldg(cell + 1));
    const float sw = sigmoid_f(__ldg(cell + 2)), sh = sigmoid_f(__ldg(cell + 3));
    o.cx = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sx, 2.0f), 0.5f), (float)gx), stride);
    o.cy = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(sy, 2.0f), 0.5f), (float)gy), stride);
    const float tw = __fmul_rn(sw, 2.0f), th = __fmul_rn(sh, 2.0f);
    o.w = __fmul_rn(__fmul_rn(tw, tw), aw);
    o.h = __fmul_rn(__fmul_rn(th, th), ah);
    o.conf = sigmoid_f(__ldg(cell + 4));
    float best = -1.f; int bi = 0;
    for (int k = 0; k < nc; ++k) {
        const float s = sigmoid_f(__ldg(cell + 5 + k));
        if (s > best) { best = s; bi = k; }
    }
    o.cls = bi;
    return o;
}

__device__ __forceinline__ bool passes(float conf, float thr, int inclusive) { return inclusive ? (conf >= thr) : (conf > thr); }

// Shared body: MODE 0 = dense rows, MODE 1 = thresholded candidates.
template <int MODE>
__global__ void __launch_bounds__(256) head_kernel(HeadDesc h, int n, float thr, int inclusive, float scale, float* rows, b2d_det* cand,
                                                   int* cand_count, int cand_cap) {
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31;
    if (h.kind == B2D_HEAD_V8_DFL) {
        // 4 lanes per anchor
        const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;    // anchor index inside the tile
        const int side = threadIdx.x & 3;
        const bool in_range = a < h.rows_total;
        int lvl = 0;
        if (in_range) { while (lvl + 1 < h.nlevels && a >= h.lv[lvl + 1].row0) ++lvl; }
        const HeadLevel& L = h.lv[lvl];
        const int cell = in_range ? a - L.row0 : 0;
        const int gy = cell / L.hw, gx = cell - gy * L.hw;
        const float* pix = L.buf + ((size_t)tile * L.hw * L.hw + cell) * L.c;
        float conf = 0.f; int cls = 0;
        bool want = in_range;
        if (in_range) {
            v8_scores(pix, h.nc, &conf, &cls);
            if (MODE == 1) { conf = __fmul_rn(conf, scale); want = passes(conf, thr, inclusive); }   // scale: gpu_handler.py:236-238
        }
        // the 4 lanes of an anchor agree on `want`; decode only where some anchor of the warp needs it
        const unsigned any = __ballot_sync(0xffffffffu, want);
        if (any == 0) return;
        float sidev = 0.f;
        if (want) sidev = dfl_side(pix + side * 16);
        Row r = v8_box(sidev, side, gx, gy, (float)L.stride, 0xffffffffu, lane & ~3);
        r.conf = conf; r.cls = cls;
        if (MODE == 0) {
            if (in_range && side == 0) {
                float* o = rows + ((size_t)tile * h.rows_total + a) * 6;
                o[0] = r.cx; o[1] = r.cy; o[2] = r.w; o[3] = r.h; o[4] = r.conf; o[5] = (float)r.cls;
            }
        } else {
            const bool emit = want && side == 0;
            const unsigned em = __ballot_sync(0xffffffffu, emit);
            if (em) {
                int base = 0;
                const int leader = __ffs(em) - 1;
                if (lane == leader) base = atomicAdd(&cand_count[tile], __popc(em));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (emit) {
                    const int slot = base + __popc(em & ((1u << lane) - 1));
                    if (slot < cand_cap) {
                        b2d_det d;
                        d.cx = r.cx; d.cy = r.cy; d.w = r.w; d.h = r.h; d.conf = r.conf; d.cls = r.cls; d.tile = tile; d.anchor = a;
                        cand[(size_t)tile * cand_cap + slot] = d;
                    }
                }
            }
        }
    } else {
        const int a = blockIdx.x * blockDim.x + threadIdx.x;   // row index inside the tile
        const bool in_range = a < h.rows_total;
        int lvl = 0;
        if (in_range) { while (lvl + 1 < h.nlevels && a >= h.lv[lvl + 1].row0) ++lvl; }
        const HeadLevel& L = h.lv[lvl];
        const int no = h.nc + 5;
        const int rel = in_range ? a - L.row0 : 0;
        const int cells = L.hw * L.hw;
        const int ai = rel / cells, cell = rel - ai * cells;     // rows ordered anchor, y, x inside a level
        const int gy = cell / L.hw, gx = cell - gy * L.hw;
        Row r; r.conf = -1.f;
        if (in_range) {
            const float* p = L.buf + ((size_t)tile * cells + cell) * L.c + ai * no;
            r = v7_row(p, h.nc, gx, gy, (float)L.stride, L.anchors[2 * ai], L.anchors[2 * ai + 1]);
            if (MODE == 1) r.conf = __fmul_rn(r.conf, scale);
        }
        if (MODE == 0) {
            if (in_range) {
                float* o = rows + ((size_t)tile * h.rows_total + a) * 6;
                o[0] = r.cx; o[1] = r.cy; o[2] = r.w; o[3] = r.h; o[4] = r.conf; o[5] = (float)r.cls;
            }
        } else {
            const bool emit = in_range && passes(r.conf, thr, inclusive);
            const unsigned em = __ballot_sync(0xffffffffu, emit);
            if (em) {
                int base = 0;
                const int leader = __ffs(em) - 1;
                if (lane == leader) base = atomicAdd(&cand_count[tile], __popc(em));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (emit) {
                    const int slot = base + __popc(em & ((1u << lane) - 1));
                    if (slot < cand_cap) {
                        b2d_det d;
                        d.cx = r.cx; d.cy = r.cy; d.w = r.w; d.h = r.h; d.conf = r.conf; d.cls = r.cls; d.tile = tile; d.anchor = a;
                        cand[(size_t)tile * cand_cap + slot] = d;
                    }
                }
            }
        }
    }
}

// candidates from already-decoded rows [n][num_rows][ncol]
__global__ void __launch_bounds__(256) rows_filter_kernel(const float* __restrict__ rows, int num_rows, int ncol, float thr,
                                                           int inclusive, float scale, b2d_det* cand, int* cand_count, int cand_cap) {
    const int tile = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_range = a < num_rows;
    const float* p = rows + ((size_t)tile * num_rows + (in_range ? a : 0)) * ncol;
    const float conf = in_range ? __fmul_rn(__ldg(p + 4), scale) : -1.f;
    const bool emit = in_range && passes(conf, thr, inclusive);
    const unsigned em = __ballot_sync(0xffffffffu, emit);
    if (!em) return;
    int base = 0;
    const int leader = __ffs(em) - 1;
    if (lane == leader) base = atomicAdd(&cand_count[tile], __popc(em));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (emit) {
        const int slot = base + __popc(em & ((1u << lane) - 1));
        if (slot < cand_cap) {
            b2d_det d;
            d.cx = __ldg(p); d.cy = __ldg(p + 1); d.w = __ldg(p + 2); d.h = __ldg(p + 3); d.conf = conf;
            d.cls = (ncol > 5) ? (int)__ldg(p + 5) : 0;
            d.tile = tile; d.anchor = a;
            cand[(size_t)tile * cand_cap + slot] = d;
        }
    }
}

// ---- select: per-tile sort (+ NMS) ----------------------------------------------------------------
constexpr int kSelCand = 256;                     // candidates per NMS chunk
constexpr int kSelParts = 4;                      // thread groups sharing a chunk's work (= 64-bit words of a suppression mask)
constexpr int kSelThreads = kSelCand * kSelParts;
static_assert(kSelParts == kSelCand / 64, "one part per mask word");
constexpr int kSmemKeys = 2048;
constexpr int kMaxDet = 512;

__device__ void bitonic_sort(unsigned long long* keys, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ bool iou_gt(const float4& a, float aa, const float4& b, float ab, float thr) {
    // torchvision nms arithmetic: inter / (area_a + area_b - inter) > thr, fp32, no contraction.  Disjoint boxes (the common
    // case) leave before the division: there torchvision computes 0 / union = 0 (or 0 / 0 = NaN), never > thr for thr > 0.
    const float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    if (!(w > 0.f)) return false;
    const float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    if (!(h > 0.f)) return false;
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    return ovr > thr;
}

// Sort kernel: one CTA of 1024 threads per tile builds the 64-bit keys (conf desc, anchor asc | row order), sorts them and
// leaves the sorted list in the scratch.  2048 .. 16384 keys (the usual case: 8400 anchors) are sorted with the keys in
// registers -- element i = m * 1024 + tid lives in r[m]; exchange distances >= 1024 stay inside a thread, distances < 32 are
// warp shuffles, only 32 <= j < 1024 goes through shared memory (35 block-wide steps instead of 105 at 16384 keys).  Smaller
// lists are sorted in shared memory, larger ones in the global scratch.
constexpr int kSortThreads = 1024;
constexpr int kSortSmemKeys = 16384;

__device__ __forceinline__ unsigned long long sort_key(const b2d_det* tc, int i, int cnt, int by_conf) {
    if (i >= cnt) return ~0ull;
    const unsigned a = (unsigned)tc[i].anchor & 0x1FFFFu;
    const unsigned long long lo = ((unsigned long long)a << 15) | (unsigned long long)(i & 0x7FFF);
    return by_conf ? (((unsigned long long)(~__float_as_uint(tc[i].conf)) << 32) | lo)     // conf desc, anchor asc
                   : lo;                                                                    // row order
}

template <int KPT>
__device__ void bitonic_sort_regs(const b2d_det* tc, int cnt, int by_conf, unsigned long long* smem, unsigned long long* out) {
    constexpr int N2 = KPT * kSortThreads;
    const int tid = threadIdx.x;
    unsigned long long r[KPT];
#pragma unroll
    for (int m = 0; m < KPT; ++m) r[m] = sort_key(tc, m * kSortThreads + tid, cnt, by_conf);
    for (int k = 2; k <= N2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= kSortThreads) {
#pragma unroll
                for (int d = KPT / 2; d >= 1; d >>= 1) {
                    if (j == d * kSortThreads) {
#pragma unroll
                        for (int m = 0; m < KPT; ++m) {
                            if ((m & d) == 0) {
                                const bool up = (((m * kSortThreads + tid) & k) == 0);
                                const unsigned long long a = r[m], b = r[m | d];
                                if ((a > b) == up) { r[m] = b; r[m | d] = a; }
                            }
                        }
                    }
                }
            } else if (j >= 32) {
#pragma unroll
                for (int m = 0; m < KPT; ++m) smem[m * kSortThreads + tid] = r[m];
                __syncthreads();
#pragma unroll
                for (int m = 0; m < KPT; ++m) {
                    const int i = m * kSortThreads + tid;
                    const unsigned long long o = smem[i ^ j];
                    const bool take_min = (((i & k) == 0) == ((i & j) == 0));     // ascending run: the lower index keeps the minimum
                    r[m] = take_min ? (r[m] < o ? r[m] : o) : (r[m] > o ? r[m] : o);
                }
                __syncthreads();
            } else {
#pragma unroll
                for (int m = 0; m < KPT; ++m) {
                    const int i = m * kSortThreads + tid;
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, r[m], j);
                    const bool take_min = (((i & k) == 0) == ((i & j) == 0));
                    r[m] = take_min ? (r[m] < o ? r[m] : o) : (r[m] > o ? r[m] : o);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < KPT; ++m) {
        const int i = m * kSortThreads + tid;
        if (i < cnt) out[i] = r[m];
    }
}

__global__ void __launch_bounds__(kSortThreads) sort_keys_kernel(const b2d_det* __restrict__ cand, const int* __restrict__ cand_count,
                                                                  int cand_cap, unsigned long long* keys_scratch, int keys_stride,
                                                                  int by_conf) {
    extern __shared__ unsigned long long dkeys[];
    const int tile = blockIdx.x;
    int cnt = cand_count[tile];
    if (cnt > cand_cap) cnt = cand_cap;
    if (cnt == 0) return;
    const b2d_det* tc = cand + (size_t)tile * cand_cap;
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    unsigned long long* out = keys_scratch + (size_t)tile * keys_stride;
    switch (n2) {
        case 2048: bitonic_sort_regs<2>(tc, cnt, by_conf, dkeys, out); return;
        case 4096: bitonic_sort_regs<4>(tc, cnt, by_conf, dkeys, out); return;
        case 8192: bitonic_sort_regs<8>(tc, cnt, by_conf, dkeys, out); return;
        case 16384: bitonic_sort_regs<16>(tc, cnt, by_conf, dkeys, out); return;
        default: break;
    }
    unsigned long long* keys = (n2 <= kSortSmemKeys) ? dkeys : out;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) keys[i] = sort_key(tc, i, cnt, by_conf);
    __syncthreads();
    bitonic_sort(keys, n2);
    if (keys != out)
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) out[i] = keys[i];
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(const b2d_det* __restrict__ cand, const int* __restrict__ cand_count,
                                                              int cand_cap, const unsigned long long* __restrict__ keys_scratch, int keys_stride,
                                                              float iou_thr, int top_k, int max_det, b2d_det* out, int* out_count,
                                                              int cap) {
    __shared__ float4 kbox[kMaxDet];
    __shared__ float karea[kMaxDet];
    __shared__ float4 cbox[kSelCand];
    __shared__ float carea[kSelCand];
    __shared__ unsigned long long cmask[kSelCand][kSelCand / 64];
    __shared__ unsigned int alive_w[kSelCand / 32];
    __shared__ int keep_slot[kSelCand];
    __shared__ int dead[kSelCand];
    __shared__ int s_nk;

    const int tile = blockIdx.x;
    const int tid = threadIdx.x;
    int cnt = cand_count[tile];
    if (cnt > cand_cap) cnt = cand_cap;
    const b2d_det* tc = cand + (size_t)tile * cand_cap;
    b2d_det* to = out + (size_t)tile * cap;
    if (cnt == 0) {
        if (tid == 0) out_count[tile] = 0;
        return;
    }
    const unsigned long long* keys = keys_scratch + (size_t)tile * keys_stride;     // sorted by sort_keys_kernel

    if (!(iou_thr > 0.f)) {
        int m = cnt;
        if (top_k > 0 && m > top_k) m = top_k;
        if (m > cap) m = cap;
        for (int i = tid; i < m; i += blockDim.x) to[i] = tc[(int)(keys[i] & 0x7FFF)];
        if (tid == 0) out_count[tile] = m;
        return;
    }

    // ---- greedy NMS over the sorted list, 256 candidates at a time ----
    // One tile has one CTA, so the work of a chunk is spread over four thread groups ("parts") of 256: part p tests candidate c
    // against the kept boxes k = p, p + 4, ... and builds word p of c's suppression mask.
    if (max_det > cap) max_det = cap;                       // max_det <= kMaxDet is checked by select_launch
    const int c = tid & (kSelCand - 1), part = tid / kSelCand;
    if (tid == 0) s_nk = 0;
    __syncthreads();
    for (int base = 0; base < cnt; base += kSelCand) {
        const int nk = s_nk;
        if (nk >= max_det) break;
        const int i = base + c;
        const bool valid = i < cnt;
        b2d_det d;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float ar = 0.f;
        if (valid) {
            d = tc[(int)(keys[i] & 0x7FFF)];
            // xywh2xyxy then + cls * max_wh, as Ultralytics
            const float hw = __fmul_rn(d.w, 0.5f), hh = __fmul_rn(d.h, 0.5f);
            const float off = __fmul_rn((float)d.cls, 7680.0f);
            bx.x = __fadd_rn(__fsub_rn(d.cx, hw), off);
            bx.y = __fadd_rn(__fsub_rn(d.cy, hh), off);
            bx.z = __fadd_rn(__fadd_rn(d.cx, hw), off);
            bx.w = __fadd_rn(__fadd_rn(d.cy, hh), off);
            ar = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        }
        if (part == 0) {
            cbox[c] = bx;
            carea[c] = ar;
            dead[c] = valid ? 0 : 1;
            keep_slot[c] = -1;
        }
        __syncthreads();
        if (valid) {
            for (int k = part; k < nk; k += kSelParts)
                if (iou_gt(kbox[k], karea[k], bx, ar, iou_thr)) { dead[c] = 1; break; }
        }
        __syncthreads();
        const bool alive = dead[c] == 0;
        if (part == 0) {
            const unsigned al = __ballot_sync(0xffffffffu, alive);
            if ((c & 31) == 0) alive_w[c >> 5] = al;
        }
        __syncthreads();
        // which later, still alive candidates of this chunk does `c` suppress?  (dead rows and dead columns are skipped:
        // after the test against the kept boxes most of a chunk is already gone)
        {
            const int w = part;
            unsigned long long m = 0;
            if (alive) {
                unsigned long long am = ((unsigned long long)alive_w[2 * w + 1] << 32) | (unsigned long long)alive_w[2 * w];
                if (w * 64 <= c) am &= (c - w * 64 >= 63) ? 0ull : (~0ull << (c - w * 64 + 1));      // only j > c
                while (am) {
                    const int jj = __ffsll((long long)am) - 1;
                    am &= am - 1;
                    const int j = w * 64 + jj;
                    if (iou_gt(bx, ar, cbox[j], carea[j], iou_thr)) m |= (1ull << jj);
                }
            }
            cmask[c][w] = m;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long removed[kSelCand / 64] = {0, 0, 0, 0};
            int k = nk;
#pragma unroll
            for (int w = 0; w < kSelCand / 64; ++w) {
                unsigned long long am = ((unsigned long long)alive_w[2 * w + 1] << 32) | (unsigned long long)alive_w[2 * w];
                while (k < max_det) {
                    am &= ~removed[w];
                    if (!am) break;
                    const int jj = __ffsll((long long)am) - 1;
                    am &= ~(1ull << jj);
                    const int j = w * 64 + jj;
                    keep_slot[j] = k++;
#pragma unroll
                    for (int v = 0; v < kSelCand / 64; ++v) removed[v] |= cmask[j][v];
                }
            }
            s_nk = k;
        }
        __syncthreads();
        if (part == 0) {
            const int slot = keep_slot[c];
            if (slot >= 0) {
                kbox[slot] = bx;
                karea[slot] = ar;
                to[slot] = d;
            }
        }
        __syncthreads();
    }
    if (tid == 0) out_count[tile] = s_nk;
}

// ---- georef -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) georef_kernel(const b2d_det* __restrict__ dets, const int* __restrict__ counts, int cap, int mode,
                                                      const double* __restrict__ params, b2d_geodet* __restrict__ out) {
    const int tile = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[tile] || i >= cap) return;
    const b2d_det d = dets[(size_t)tile * cap + i];
    const double* P = params + (size_t)tile * B2D_GEO_PARAMS;
    b2d_geodet g;
    g.conf = d.conf; g.tile = tile; g.x_yolo = d.cx; g.y_yolo = d.cy; g.x_img = 0.f; g.y_img = 0.f;
    const double x = (double)d.cx, y = (double)d.cy;
    if (mode == B2D_GEO_BOUNDS) {
        // simple_detector.py:487-494 -- P = west, east, south, north, crop_size, model_size
        const double ms = P[5] > 0.0 ? P[5] : 640.0;
        const double xf = __ddiv_rn(x, ms), yf = __ddiv_rn(y, ms);
        g.x_img = (float)__dmul_rn(xf, P[4]);
        g.y_img = (float)__dmul_rn(yf, P[4]);
        g.x = __dadd_rn(P[0], __dmul_rn(xf, __dsub_rn(P[1], P[0])));
        g.y = __dsub_rn(P[3], __dmul_rn(yf, __dsub_rn(P[3], P[2])));
    } else if (mode == B2D_GEO_GPUHANDLER) {
        // gpu_handler.py:182-190 -- P = lon_min, lat_min, lon_max, lat_max; keeps the *864/864 round trip
        const double x864 = __dmul_rn(__ddiv_rn(x, 640.0), 864.0), y864 = __dmul_rn(__ddiv_rn(y, 640.0), 864.0);
        g.x_img = (float)x864; g.y_img = (float)y864;
        g.x = __dadd_rn(P[0], __dmul_rn(__ddiv_rn(x864, 864.0), __dsub_rn(P[2], P[0])));
        g.y = __dsub_rn(P[3], __dmul_rn(__ddiv_rn(y864, 864.0), __dsub_rn(P[3], P[1])));
    } else if (mode == B2D_GEO_TENSOR_F32) {
        // gpu_handler.py:243-253 (_process_tensors): float32 CUDA-tensor arithmetic.  `boxes[:, :2] / 640` on a CUDA tensor is a
        // multiplication by the float reciprocal of the scalar; the Python scalars are cast to float32.  P = lon_min, lat_min,
        // lon_max, lat_max.
        const float inv = __fdiv_rn(1.0f, 640.0f);
        const float cxn = __fmul_rn(d.cx, inv), cyn = __fmul_rn(d.cy, inv);
        const float lon_off = (float)__dsub_rn(P[2], P[0]), lat_off = (float)__dsub_rn(P[3], P[1]);
        g.x = (double)__fadd_rn((float)P[0], __fmul_rn(cxn, lon_off));
        g.y = (double)__fsub_rn((float)P[3], __fmul_rn(cyn, lat_off));
    } else {
        // Ultralytics scale_boxes (fp32) then the notebook's centroid + pixel_to_geo (fp64)
        // P = gt[0..5], win_x, win_y, pad_x, pad_y, gain, w0, h0
        const float hw = __fmul_rn(d.w, 0.5f), hh = __fmul_rn(d.h, 0.5f);
        float x1 = __fsub_rn(d.cx, hw), y1 = __fsub_rn(d.cy, hh), x2 = __fadd_rn(d.cx, hw), y2 = __fadd_rn(d.cy, hh);
        const float px = (float)P[8], py = (float)P[9], gain = (float)P[10], w0 = (float)P[11], h0 = (float)P[12];
        x1 = __fdiv_rn(__fsub_rn(x1, px), gain); x2 = __fdiv_rn(__fsub_rn(x2, px), gain);
        y1 = __fdiv_rn(__fsub_rn(y1, py), gain); y2 = __fdiv_rn(__fsub_rn(y2, py), gain);
        x1 = fminf(fmaxf(x1, 0.f), w0); x2 = fminf(fmaxf(x2, 0.f), w0);
        y1 = fminf(fmaxf(y1, 0.f), h0); y2 = fminf(fmaxf(y2, 0.f), h0);
        // (x1 + x2) is an fp32 add; "/ 2" and "+= x" promote to float64 under the reference's NumPy 1.26
        const double cxp = __dadd_rn(__ddiv_rn((double)__fadd_rn(x1, x2), 2.0), P[6]);
        const double cyp = __dadd_rn(__ddiv_rn((double)__fadd_rn(y1, y2), 2.0), P[7]);
        g.x_img = (float)cxp; g.y_img = (float)cyp;
        g.x = __dadd_rn(__dadd_rn(P[0], __dmul_rn(cxp, P[1])), __dmul_rn(cyp, P[2]));
        g.y = __dadd_rn(__dadd_rn(P[3], __dmul_rn(cxp, P[4])), __dmul_rn(cyp, P[5]));
    }
    out[(size_t)tile * cap + i] = g;
}

}  // namespace

int decode_rows_launch(const HeadDesc* h, int n, float* rows, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int per = (h->kind == B2D_HEAD_V8_DFL) ? 4 : 1;
    dim3 grid(ceil_div(h->rows_total * per, 256), n);
    head_kernel<0><<<grid, 256, 0, stream>>>(*h, n, 0.f, 1, 1.f, rows, nullptr, nullptr, 0);
    B2D_LAUNCH_CHECK();
    return 0;
}

int candidates_from_head_launch(const HeadDesc* h, int n, float thr, int inclusive, float scale, b2d_det* cand, int* cand_count,
                                int cand_cap, cudaStream_t stream) {
    if (n <= 0) return 0;
    B2D_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * n, stream));
    const int per = (h->kind == B2D_HEAD_V8_DFL) ? 4 : 1;
    dim3 grid(ceil_div(h->rows_total * per, 256), n);
    head_kernel<1><<<grid, 256, 0, stream>>>(*h, n, thr, inclusive, scale, nullptr, cand, cand_count, cand_cap);
    B2D_LAUNCH_CHECK();
    return 0;
}

int candidates_from_rows_launch(const float* rows, int n, int num_rows, int ncol, float thr, int inclusive, float scale, b2d_det* cand,
                                int* cand_count, int cand_cap, cudaStream_t stream) {
    if (n <= 0) return 0;
    B2D_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * n, stream));
    dim3 grid(ceil_div(num_rows, 256), n);
    rows_filter_kernel<<<grid, 256, 0, stream>>>(rows, num_rows, ncol, thr, inclusive, scale, cand, cand_count, cand_cap);
    B2D_LAUNCH_CHECK();
    return 0;
}

int select_launch(const b2d_det* cand, const int* cand_count, int cand_cap, int n, unsigned long long* keys_scratch, float iou_thr,
                  int top_k, int max_det, b2d_det* out, int* out_count, int cap, cudaStream_t stream) {
    if (n <= 0) return 0;
    B2D_CHECK(cand_cap <= 32768, "select: candidate capacity %d exceeds the 15-bit slot field", cand_cap);
    B2D_CHECK(!(iou_thr > 0.f) || max_det <= kMaxDet, "select: max_det %d exceeds the NMS kernel's %d kept boxes per tile", max_det, kMaxDet);
    int stride = 1;
    while (stride < cand_cap) stride <<= 1;
    {
        if (b2d_func_smem_optin((const void*)sort_keys_kernel, kSortSmemKeys * 8)) return -2;
        int need = 1;
        while (need < cand_cap && need < kSortSmemKeys) need <<= 1;       // keys that can occur, capped by the smem variant
        sort_keys_kernel<<<n, kSortThreads, (size_t)need * 8, stream>>>(cand, cand_count, cand_cap, keys_scratch, stride, (iou_thr > 0.f) || (top_k > 0));
        B2D_LAUNCH_CHECK();
    }
    select_kernel<<<n, kSelThreads, 0, stream>>>(cand, cand_count, cand_cap, keys_scratch, stride, iou_thr, top_k, max_det, out,
                                                 out_count, cap);
    B2D_LAUNCH_CHECK();
    return 0;
}

int georef_launch(const b2d_det* dets, const int* counts, int n, int cap, int mode, const double* params, b2d_geodet* out,
                  cudaStream_t stream) {
    if (n <= 0 || cap <= 0) return 0;
    dim3 grid(ceil_div(cap, 256), n);
    georef_kernel<<<grid, 256, 0, stream>>>(dets, counts, cap, mode, params, out);
    B2D_LAUNCH_CHECK();
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Segmentation head (config C5): per-pixel argmax + softmax probability over nc <= 8 fp32 logits of an NHWC buffer with c
// channels per pixel (c * 4 bytes a multiple of 16).  HBM-bound: 16 bytes read and 1 (+4) written per pixel.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void segment_kernel(const float4* __restrict__ logits, long long npix, int c4, int nc, uint8_t* __restrict__ labels,
                               float* __restrict__ conf) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j < c4) {
                const float4 q = __ldg(logits + i * c4 + j);
                v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
            }
        }
        int best = 0;
        float m = v[0];
#pragma unroll
        for (int k = 1; k < 8; ++k)
            if (k < nc && v[k] > m) { m = v[k]; best = k; }
        labels[i] = (uint8_t)best;
        if (conf) {
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nc) sum += expf(v[k] - m);
            conf[i] = 1.0f / sum;
        }
    }
}
}  // namespace

int segment_launch(const float* logits, long long npix, int c, int nc, uint8_t* labels, float* conf, cudaStream_t stream) {
    B2D_CHECK(c % 4 == 0 && c <= 8 && nc >= 1 && nc <= c, "segment: %d classes in %d channels not supported", nc, c);
    if (npix <= 0) return 0;
    const int threads = 256;
    long long blocks = (npix + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    segment_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const float4*)logits, npix, c / 4, nc, labels, conf);
    B2D_LAUNCH_CHECK();
    return 0;
}
