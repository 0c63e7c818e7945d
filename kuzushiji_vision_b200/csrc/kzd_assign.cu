// TaskAlignedAssigner (sm_100a), GT-centric and sparse: no dense (b,G,A) tensor is ever formed.
//
// reference: utils/tal.py:39-295 (forward/_forward :39-118, get_pos_mask :120-130, get_box_metrics
// :132-151, select_topk_candidates :157-190, get_targets :192-238, select_candidates_in_gts :241-262,
// select_highest_overlaps :264-295) and CIoU utils/metrics.py:96-130.
//
// Kernels (one stream, no host sync):
//   assign_begin_kernel       per-call flags and the device-side epoch of the persistent state
//   assign_verify_kernel      checks anc_points against the level table (grid accelerator is only
//                             used when every anchor matches bit-for-bit)
//   assign_candidates_kernel  one warp per (image, GT): enumerate the anchors inside the GT (separable
//                             x / y tests), queue them, CIoU + alignment metric 32 pairs at a time,
//                             per-anchor argmax-overlap via 64-bit atomicMax, warp-shuffle top-k,
//                             selecting GTs recorded with two atomicMax words
//   assign_resolve_kernel     one thread per anchor: resolve the owner GT, write labels / boxes /
//                             fg / gt_idx for every anchor, append foreground anchors to a compact list
//   assign_fg_metric_kernel   foreground list: recompute the owner pair's metric, per-GT maxima
//   assign_normalize_kernel   foreground list: target_scores = align * pos_ov / (pos_align+eps)
//
// Bit-exactness: every fp32 operation of the eager reference rounds separately, so the pair
// arithmetic below uses __fadd_rn/__fmul_rn/__fdiv_rn (never contracted into FMAs) and the same
// libdevice atanf / powf / sqrtf that aten's CUDA kernels call.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kzd_common.cuh"

namespace kzd {

typedef unsigned long long u64;

// Per-anchor state words carry an 8-bit epoch in their top bits, so a persistent workspace never has
// to be re-zeroed between calls: entries written under an older epoch lose every atomicMax / CAS.
//   key  = epoch:8 | overlap_bits:30 | (KEY_GMASK - g):26     (overlap <= 1.0f  =>  bits < 2^30)
//   sel  = { epoch:8 | (SEL_GMASK - g):24 ,  epoch:8 | g:24 }  both updated with atomicMax (no return
//          value needed): smallest and largest selecting GT; equal => exactly one selector
constexpr int KEY_GBITS = 26;
constexpr u64 KEY_GMASK = (1ull << KEY_GBITS) - 1ull;
constexpr uint32_t SEL_GMASK = 0xffffffu;

struct AssignWs {
    u64* key;            // (b*A)  max over in-GT pairs of the packed (overlap, ~g) key
    uint2* sel;          // (b*A)  smallest / largest GT whose top-k selected the anchor
    uint32_t* pos_align; // (b*G)  per-GT max alignment metric (float bits)
    uint32_t* pos_ov;    // (b*G)  per-GT max overlap (float bits)
    uint32_t* fg_list;   // (b*A)  compact list of foreground anchors (flat index)
    int32_t* flags;      // [0] != 0: anc_points does not match the level table; [1] = fg count;
                         // [2] = current epoch (1..255), advanced on the device by assign_begin_kernel
    size_t state_bytes;  // key + sel + flags: the part that must have been zeroed once
};

struct AssignParams {
    const float *pd_scores, *pd_bboxes, *anc, *gt_labels, *gt_bboxes, *mask_gt;
    int b, A, G, nc, topk;
    float alpha, beta, eps;
    kzd_level lv[KZD_MAX_LEVELS];
    int n_levels, zero_fill;
    int score_dtype;  // dtype pd_scores had at the call site: 0 fp32, 1 fp16, 2 bf16 (AMP, utils/tal.py:137)
    long long* target_labels;
    float* target_bboxes;
    float* target_scores;
    uint8_t* fg_mask;
    long long* target_gt_idx;
    double* counters;
    AssignWs ws;
};

// ---- arithmetic shared by the candidates and resolve kernels (must be one instruction sequence) ----

// aten pow(tensor, scalar) dispatch: 0.5 -> sqrt, 2 -> x*x, 3 -> x*x*x, 1 -> x, 0 -> 1, -0.5 -> rsqrt,
// -1 -> 1/x, -2 -> 1/(x*x), else ::pow  (ATen/native/cuda/PowKernel.cu)
__device__ __forceinline__ float pow_scalar(float x, float p) {
    if (p == 0.5f) return sqrtf(x);
    if (p == 2.f) return __fmul_rn(x, x);
    if (p == 1.f) return x;
    if (p == 3.f) return __fmul_rn(__fmul_rn(x, x), x);
    if (p == 0.f) return 1.f;
    if (p == -0.5f) return rsqrtf(x);
    if (p == -1.f) return __fdiv_rn(1.f, x);
    if (p == -2.f) return __fdiv_rn(1.f, __fmul_rn(x, x));
    return powf(x, p);
}

// bbox_scores inherits pd_scores' dtype (utils/tal.py:137), so under AMP bbox_scores.pow(alpha) is rounded
// to half precision before it meets the fp32 overlaps (utils/tal.py:150).
__device__ __forceinline__ float score_pow(float s, float alpha, int score_dtype) {
    const float v = pow_scalar(s, alpha);
    if (score_dtype == 1) return __half2float(__float2half_rn(v));
    if (score_dtype == 2) return __bfloat162float(__float2bfloat16_rn(v));
    return v;
}

struct GtBox {
    float x1, y1, x2, y2;
    float w, h, at;  // w1, h1 (+eps) and atan(w1/h1) of utils/metrics.py:107,126 - constant per GT
};

__device__ __forceinline__ GtBox make_gt(float4 g) {
    GtBox r;
    r.x1 = g.x, r.y1 = g.y, r.x2 = g.z, r.y2 = g.w;
    r.w = __fsub_rn(g.z, g.x);
    r.h = __fadd_rn(__fsub_rn(g.w, g.y), 1e-7f);
    r.at = atanf(__fdiv_rn(r.w, r.h));
    return r;
}

// utils/tal.py:259-262: min of the four signed distances strictly above 1e-9
__device__ __forceinline__ bool anchor_in_gt(float ax, float ay, const GtBox& g) {
    const float d = fminf(fminf(__fsub_rn(ax, g.x1), __fsub_rn(ay, g.y1)),
                          fminf(__fsub_rn(g.x2, ax), __fsub_rn(g.y2, ay)));
    return d > 1e-9f;
}

// clamp(CIoU(box1 = GT, box2 = prediction), 0)   (utils/metrics.py:103-130, utils/tal.py:155)
__device__ __forceinline__ float overlap_ciou(const GtBox& g, const float4 p) {
    const float e = 1e-7f;
    const float pw = __fsub_rn(p.z, p.x);
    const float ph = __fadd_rn(__fsub_rn(p.w, p.y), e);
    const float iw = fmaxf(__fsub_rn(fminf(g.x2, p.z), fmaxf(g.x1, p.x)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(g.y2, p.w), fmaxf(g.y1, p.y)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(g.w, g.h), __fmul_rn(pw, ph)), inter), e);
    const float iou = __fdiv_rn(inter, uni);
    const float cw = __fsub_rn(fmaxf(g.x2, p.z), fminf(g.x1, p.x));
    const float ch = __fsub_rn(fmaxf(g.y2, p.w), fminf(g.y1, p.y));
    const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)), e);
    const float dx = __fsub_rn(__fsub_rn(__fadd_rn(p.x, p.z), g.x1), g.x2);
    const float dy = __fsub_rn(__fsub_rn(__fadd_rn(p.y, p.w), g.y1), g.y2);
    const float rho2 = __fmul_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 0.25f);
    const float da = __fsub_rn(atanf(__fdiv_rn(pw, ph)), g.at);
    const float v = __fmul_rn(0.40528473456935109f, __fmul_rn(da, da));  // float(4 / pi^2)
    const float al = __fdiv_rn(v, __fadd_rn(__fsub_rn(v, iou), 1.00000011920928955f));  // float(1 + 1e-7)
    const float ciou = __fsub_rn(iou, __fadd_rn(__fdiv_rn(rho2, c2), __fmul_rn(v, al)));
    return fmaxf(ciou, 0.f);
}

__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_up64(u64 v, int d) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
    const uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
    return ((u64)hi << 32) | lo;
}

// ---- begin: per-call flags and the device-side epoch -----------------------------------------

__global__ void assign_begin_kernel(AssignWs ws, int reset_epoch) {
    ws.flags[0] = 0;
    ws.flags[1] = 0;
    const int e = reset_epoch ? 0 : ws.flags[2];
    ws.flags[2] = (e % 255) + 1;  // 1..255; at 255 the resolve kernel wipes the state for the wrap to 1
}

// ---- verify ----------------------------------------------------------------------------------

__global__ void assign_verify_kernel(const float* __restrict__ anc, int A, AssignParams p) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    int l = 0;
    for (int k = 1; k < p.n_levels; ++k)
        if (a >= p.lv[k].offset) l = k;
    const int q = a - p.lv[l].offset;
    const float s = p.lv[l].stride;
    const float ex = __fmul_rn((float)(q % p.lv[l].w) + 0.5f, s);
    const float ey = __fmul_rn((float)(q / p.lv[l].w) + 0.5f, s);
    if (anc[2 * a] != ex || anc[2 * a + 1] != ey) atomicOr(p.ws.flags, 1);
}

// ---- candidates + top-k ----------------------------------------------------------------------

constexpr int CAND_THREADS = 256;
#ifndef KZD_CAND_MINB
#define KZD_CAND_MINB 4
#endif
constexpr int CAND_WARPS = CAND_THREADS / 32;

struct TopK {
    u64 tk;   // lane i holds the i-th best  metric_bits<<32 | ~anchor  (0 = empty)
    u64 kth;  // entry topk-1: the bar a new candidate has to clear
    bool empty;
};

// 32 keys, one per lane -> descending order across the lanes (bitonic network on shuffles)
__device__ __forceinline__ u64 warp_sort_desc(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, j);
            const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), j);
            const u64 o = ((u64)hi << 32) | lo;
            const bool want_max = (((lane & k) == 0) == ((lane & j) == 0));  // descending overall
            v = (want_max == (o > v)) ? o : v;
        }
    }
    return v;
}

// One in-GT pair per lane (a < 0: idle lane): overlap, metric, argmax key, streaming top-k insert.
__device__ __forceinline__ void eval_pairs(const AssignParams& p, const GtBox& gt, int a, size_t img,
                                           const float4* __restrict__ pdb, const float* __restrict__ pds, int g,
                                           int topk, int lane, uint32_t epoch, TopK& st) {
    u64 key = 0ull;
    if (a >= 0) {
        const float4 pb = __ldg(pdb + a);
        const float sc = __ldg(pds + (size_t)a * p.nc);
        const float o = overlap_ciou(gt, pb);
        if (o > 0.f)
            atomicMax(p.ws.key + img + a, ((u64)epoch << 56) | ((u64)min(__float_as_uint(o), 0x3fffffffu) << KEY_GBITS) |
                                              (KEY_GMASK - (u64)g));
        const float al = __fmul_rn(score_pow(sc, p.alpha, p.score_dtype), pow_scalar(o, p.beta));
        if (al > 0.f) key = ((u64)__float_as_uint(al) << 32) | (0xffffffffu - (uint32_t)a);
    }
    if (st.empty) {  // first batch: one warp sort instead of up to 32 serial insertions
        st.tk = warp_sort_desc(key, lane);
        st.kth = shfl64(st.tk, topk - 1);
        st.empty = false;
        return;
    }
    // The list stays sorted over all 32 lanes, so inserting a key that the k-th entry has meanwhile
    // overtaken is harmless (it lands at a lane >= topk): the bar is refreshed once per batch.
    unsigned m = __ballot_sync(0xffffffffu, key > st.kth);
    if (m == 0u) return;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const u64 ck = shfl64(key, src);
        const int pos = __popc(__ballot_sync(0xffffffffu, st.tk > ck));
        const u64 up = shfl_up64(st.tk, 1);
        if (lane == pos) st.tk = ck;
        else if (lane > pos) st.tk = up;
    }
    st.kth = shfl64(st.tk, topk - 1);
}

// GT `g` selected the anchor: two fire-and-forget atomicMax (smallest and largest selecting GT).
__device__ __forceinline__ void mark_selected(uint2* __restrict__ sel, uint32_t epoch, int g) {
    atomicMax(&sel->x, (epoch << 24) | (SEL_GMASK - (uint32_t)g));
    atomicMax(&sel->y, (epoch << 24) | (uint32_t)g);
}

__global__ void __launch_bounds__(CAND_THREADS, KZD_CAND_MINB) assign_candidates_kernel(const AssignParams p) {
    __shared__ int queue[CAND_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)p.b * p.G) return;
    const int bi = (int)(wid / p.G), g = (int)(wid % p.G);
    if (lane == 0) p.ws.pos_align[wid] = 0u, p.ws.pos_ov[wid] = 0u;
    if (wid == 0 && lane == 0 && p.counters) p.counters[0] = 0.0, p.counters[1] = 0.0;
    if (p.mask_gt[wid] == 0.f) return;  // padded GT: claims nothing, contributes zero overlap

    const GtBox gt = make_gt(*reinterpret_cast<const float4*>(p.gt_bboxes + wid * 4));
    int label = (int)p.gt_labels[wid];
    label = label < 0 ? 0 : (label >= p.nc ? p.nc - 1 : label);
    const bool grid_ok = p.n_levels > 0 && p.ws.flags[0] == 0;
    const uint32_t epoch = (uint32_t)p.ws.flags[2];
    const size_t img = (size_t)bi * p.A;
    const float2* __restrict__ anc = reinterpret_cast<const float2*>(p.anc);
    const float4* __restrict__ pdb = reinterpret_cast<const float4*>(p.pd_bboxes) + img;
    const float* __restrict__ pds = p.pd_scores + img * p.nc + label;
    const int topk = p.topk;
    int* q = queue[wib];
    int qn = 0;  // queued in-GT anchors (warp-uniform)
    TopK st{0ull, 0ull, true};

    // In-GT test (utils/tal.py:259-262): min of the four fp32 differences > 1e-9  <=>  each one is,
    // so it separates into an x test per column and a y test per row of the level's grid.
    auto flush32 = [&]() {
        const int a32 = q[lane];
        const int tail = q[32 + lane];
        __syncwarp();
        q[lane] = tail;
        qn -= 32;
        __syncwarp();
        eval_pairs(p, gt, a32, img, pdb, pds, g, topk, lane, epoch, st);
    };
    if (grid_ok) {
        for (int l = 0; l < p.n_levels; ++l) {
            // superset of the cells whose centre can be inside the GT; the exact fp32 predicates decide.
            // centres inside need index in (x1/s-0.5, x2/s-0.5); [floor, ceil] of the computed bounds
            // contains it even when the quotient is off by an ulp
            const float s = p.lv[l].stride, inv_s = 1.f / s;
            const int W = p.lv[l].w, H = p.lv[l].h, off = p.lv[l].offset;
            int ia = (int)floorf(gt.x1 * inv_s - 0.5f) - 1, ib = (int)ceilf(gt.x2 * inv_s - 0.5f) + 1;
            int ja = (int)floorf(gt.y1 * inv_s - 0.5f) - 1, jb = (int)ceilf(gt.y2 * inv_s - 0.5f) + 1;
            ia = max(ia, 0), ja = max(ja, 0), ib = min(ib, W - 1), jb = min(jb, H - 1);
            if (ia > ib || ja > jb) continue;
            for (int i0 = ia; i0 <= ib; i0 += 32) {  // 32 columns at a time (one pass unless the GT is huge)
                const int ii = i0 + lane;
                // anchor x verified bit-identical to anc_points by assign_verify_kernel
                const float ax = __fmul_rn((float)ii + 0.5f, s);
                const bool x_in = ii <= ib && __fsub_rn(ax, gt.x1) > 1e-9f && __fsub_rn(gt.x2, ax) > 1e-9f;
                const unsigned xm = __ballot_sync(0xffffffffu, x_in);
                if (xm == 0u) continue;
                const int nxm = __popc(xm), my_pos = __popc(xm & ((1u << lane) - 1u));
                for (int jj = ja; jj <= jb; ++jj) {
                    const float ay = __fmul_rn((float)jj + 0.5f, s);
                    if (!(__fsub_rn(ay, gt.y1) > 1e-9f && __fsub_rn(gt.y2, ay) > 1e-9f)) continue;  // uniform
                    if (x_in) q[qn + my_pos] = off + jj * W + ii;
                    qn += nxm;
                    __syncwarp();
                    if (qn >= 32) flush32();
                }
            }
        }
    } else {
        for (int t0 = 0; t0 < p.A; t0 += 32) {  // anchors are not a make_anchors grid: test them all
            const int a = t0 + lane;
            bool inside = false;
            if (a < p.A) {
                const float2 ap = __ldg(anc + a);
                inside = anchor_in_gt(ap.x, ap.y, gt);
            }
            const unsigned m = __ballot_sync(0xffffffffu, inside);
            if (inside) q[qn + __popc(m & ((1u << lane) - 1u))] = a;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) flush32();
        }
    }
    if (qn > 0) eval_pairs(p, gt, lane < qn ? q[lane] : -1, img, pdb, pds, g, topk, lane, epoch, st);

    const bool has = lane < topk && st.tk != 0ull;
    if (has) mark_selected(p.ws.sel + img + (0xffffffffu - (uint32_t)st.tk), epoch, g);
    if (p.zero_fill) {
        // torch.topk pads a short list with zero-metric entries, lowest anchor index first; a padded
        // pick only matters when that anchor lies inside this GT (utils/tal.py:128).
        int need = topk - __popc(__ballot_sync(0xffffffffu, has));
        for (int a = 0; need > 0 && a < p.A; ++a) {
            const bool mine = has && (0xffffffffu - (uint32_t)st.tk) == (uint32_t)a;
            if (__any_sync(0xffffffffu, mine)) continue;  // positive pick, already counted
            --need;
            const float2 ap = __ldg(anc + a);
            if (lane == 0 && anchor_in_gt(ap.x, ap.y, gt)) mark_selected(p.ws.sel + img + a, epoch, g);
        }
    }
}

// ---- resolve -----------------------------------------------------------------------------------

constexpr int RES_THREADS = 256;
constexpr int RES_ITERS = 8;  // anchors per CTA = RES_THREADS * RES_ITERS

// The image's GT table is staged in shared memory: boxes by one TMA bulk copy (always 16 B aligned),
// labels and mask by plain loads. Each thread then owns RES_ITERS anchors (state words loaded up
// front so RES_ITERS independent loads are in flight).
__global__ void __launch_bounds__(RES_THREADS) assign_resolve_kernel(const AssignParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ int warp_tot[RES_THREADS / 32];
    __shared__ int block_base;
    float4* s_box = reinterpret_cast<float4*>(smem_raw);
    float* s_lab = reinterpret_cast<float*>(smem_raw + (size_t)p.G * 16);
    float* s_msk = s_lab + p.G;
    const int bi = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)p.G * 16u;
        mbar_expect_tx(&bar, bytes);
        tma_load_1d(s_box, p.gt_bboxes + (size_t)bi * p.G * 4, bytes, &bar);
    }
    for (int g = tid; g < p.G; g += RES_THREADS) {
        s_lab[g] = p.gt_labels[(size_t)bi * p.G + g];
        s_msk[g] = p.mask_gt[(size_t)bi * p.G + g];
    }
    const uint32_t epoch = (uint32_t)p.ws.flags[2];
    const size_t img = (size_t)bi * p.A;
    const int a_base = blockIdx.x * RES_ITERS * RES_THREADS + tid;
    uint2 sel[RES_ITERS];
#pragma unroll
    for (int it = 0; it < RES_ITERS; ++it) {
        const int a = a_base + it * RES_THREADS;
        sel[it] = a < p.A ? p.ws.sel[img + a] : make_uint2(0u, 0u);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    unsigned fg_bits = 0u;
#pragma unroll
    for (int it = 0; it < RES_ITERS; ++it) {
        const int a = a_base + it * RES_THREADS;
        if (a >= p.A) break;
        const size_t idx = img + a;
        const bool fg = (sel[it].y >> 24) == epoch;
        uint32_t g = 0;
        if (fg) {
            fg_bits |= 1u << it;
            const uint32_t g_min = SEL_GMASK - (sel[it].x & SEL_GMASK), g_max = sel[it].y & SEL_GMASK;
            if (g_min == g_max) {
                g = g_min;
            } else {  // contested: first-max argmax of overlaps over ALL GTs (utils/tal.py:286)
                const u64 k = p.ws.key[idx];
                const bool any = (k >> 56) == epoch && ((k >> KEY_GBITS) & 0x3fffffffull) != 0ull;
                g = any ? (uint32_t)(KEY_GMASK - (k & KEY_GMASK)) : 0u;
            }
        }
        const float4 gb = s_box[g];
        long long lab = (long long)s_lab[g];
        lab = lab < 0 ? 0 : lab;
        p.target_gt_idx[idx] = (long long)g;
        p.target_labels[idx] = lab;
        p.fg_mask[idx] = fg ? 1 : 0;
        reinterpret_cast<float4*>(p.target_bboxes)[idx] = gb;
        // foreground entries get their alignment metric from assign_fg_metric_kernel (compact list)
        if (p.nc == 1) p.target_scores[idx] = 0.f;
        if (epoch == 255u) {  // last epoch before the wrap: leave the state clean for epoch 1
            p.ws.sel[idx] = make_uint2(0u, 0u);
            p.ws.key[idx] = 0ull;
        }
    }
    // compact list of foreground anchors for the normalise pass: one global atomic per CTA
    const int mine = __popc(fg_bits);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wib] = incl;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < RES_THREADS / 32; ++w) {
            const int t = warp_tot[w];
            warp_tot[w] = tot;
            tot += t;
        }
        block_base = tot ? atomicAdd(p.ws.flags + 1, tot) : 0;
    }
    __syncthreads();
    int slot = block_base + warp_tot[wib] + incl - mine;
#pragma unroll
    for (int it = 0; it < RES_ITERS; ++it)
        if (fg_bits & (1u << it)) p.ws.fg_list[slot++] = (uint32_t)(img + a_base + it * RES_THREADS);
}

// ---- foreground metric -------------------------------------------------------------------------

// Compact foreground list, all lanes busy: alignment metric of each (anchor, owner GT) pair, recomputed
// with the same instruction sequence as in assign_candidates_kernel (=> identical bits; also covers
// anchors owned by a GT that did not top-k them), stored in the owner's target_scores slot, and the
// per-GT maxima of utils/tal.py:112-114.
__global__ void __launch_bounds__(256) assign_fg_metric_kernel(const AssignParams p) {
    const int n = p.ws.flags[1];
    const float2* __restrict__ anc = reinterpret_cast<const float2*>(p.anc);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t idx = p.ws.fg_list[i];
        const int bi = (int)(idx / (size_t)p.A);
        const int a = (int)(idx - (size_t)bi * p.A);
        const long long g = p.target_gt_idx[idx];
        long long lab = p.target_labels[idx];
        lab = lab >= p.nc ? p.nc - 1 : lab;
        const size_t gslot = (size_t)bi * p.G + g;
        const GtBox gt = make_gt(*reinterpret_cast<const float4*>(p.gt_bboxes + gslot * 4));
        const float2 ap = __ldg(anc + a);
        float al = 0.f;
        if (p.mask_gt[gslot] != 0.f && anchor_in_gt(ap.x, ap.y, gt)) {
            const float4 pb = __ldg(reinterpret_cast<const float4*>(p.pd_bboxes) + idx);
            const float sc = __ldg(p.pd_scores + idx * p.nc + lab);
            const float o = overlap_ciou(gt, pb);
            al = __fmul_rn(score_pow(sc, p.alpha, p.score_dtype), pow_scalar(o, p.beta));
            atomicMax(p.ws.pos_ov + gslot, __float_as_uint(o));
            atomicMax(p.ws.pos_align + gslot, __float_as_uint(al));
        }
        p.target_scores[idx * p.nc + lab] = al;
    }
}

// ---- normalise ---------------------------------------------------------------------------------

// Foreground anchors only (compact list): target_scores = align * pos_ov / (pos_align + eps).
__global__ void __launch_bounds__(256) assign_normalize_kernel(const AssignParams p) {
    __shared__ double s_sum;
    if (threadIdx.x == 0) s_sum = 0.0;
    __syncthreads();
    const int n = p.ws.flags[1];
    float local = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t idx = p.ws.fg_list[i];
        const int bi = (int)(idx / (size_t)p.A);
        const long long g = p.target_gt_idx[idx];
        long long lab = p.target_labels[idx];
        lab = lab >= p.nc ? p.nc - 1 : lab;
        float* dst = p.target_scores + idx * p.nc + lab;
        const float ov = __uint_as_float(p.ws.pos_ov[(size_t)bi * p.G + g]);
        const float pa = __uint_as_float(p.ws.pos_align[(size_t)bi * p.G + g]);
        const float v = __fdiv_rn(__fmul_rn(*dst, ov), __fadd_rn(pa, p.eps));  // utils/tal.py:115
        *dst = v;
        local += v;
    }
    if (p.counters) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local != 0.f) atomicAdd(&s_sum, (double)local);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_sum != 0.0) atomicAdd(p.counters, s_sum);
            if (blockIdx.x == 0) p.counters[1] = (double)n;
        }
    }
}

__global__ void __launch_bounds__(256) zero_fill_kernel(float4* __restrict__ dst, size_t n4, float* __restrict__ tail,
                                                        int n_tail) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(dst + i, z);
    if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] = 0.f;
}

static void split_ws(void* workspace, int b, int A, int G, AssignWs& ws, size_t& total) {
    unsigned char* p = (unsigned char*)workspace;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* r = p ? p + o : nullptr;
        o += align_up(bytes, 256);
        return r;
    };
    const size_t n = (size_t)b * A;
    ws.key = (u64*)take(n * 8);
    ws.sel = (uint2*)take(n * 8);
    ws.flags = (int32_t*)take(256);
    ws.state_bytes = o;
    ws.pos_align = (uint32_t*)take((size_t)b * G * 4);
    ws.pos_ov = (uint32_t*)take((size_t)b * G * 4);
    ws.fg_list = (uint32_t*)take(n * 4);
    total = o;
}

}  // namespace kzd

using namespace kzd;

extern "C" __attribute__((visibility("default"))) size_t kzd_assign_workspace_bytes(int b, int A, int G) {
    if (b < 0 || A < 0 || G < 0) return 0;
    AssignWs ws;
    size_t total = 0;
    split_ws(nullptr, b, A, G, ws, total);
    return total;
}

extern "C" __attribute__((visibility("default"))) int kzd_tal_assign(const float* pd_scores, const float* pd_bboxes, const float* anc_points,
                              const float* gt_labels, const float* gt_bboxes, const float* mask_gt, int b, int A,
                              int G, int nc, int topk, float alpha, float beta, float eps, const kzd_level* h_levels,
                              int n_levels, int zero_fill, int score_dtype, int64_t* target_labels, float* target_bboxes,
                              float* target_scores, uint8_t* fg_mask, int64_t* target_gt_idx, double* counters,
                              void* workspace, size_t workspace_bytes, int persistent, void* stream) {
    KZD_CHECK_ARG(pd_scores && pd_bboxes && anc_points && gt_labels && gt_bboxes && mask_gt,
                  "kzd_tal_assign: null input");
    KZD_CHECK_ARG(target_labels && target_bboxes && target_scores && fg_mask && target_gt_idx && workspace,
                  "kzd_tal_assign: null output/workspace");
    KZD_CHECK_ARG(b >= 0 && A >= 1 && G >= 1 && nc >= 1, "kzd_tal_assign: bad sizes b=%d A=%d G=%d nc=%d", b, A, G, nc);
    KZD_CHECK_ARG(topk >= 1 && topk <= KZD_MAX_TOPK, "kzd_tal_assign: topk=%d not in [1,%d]", topk, KZD_MAX_TOPK);
    KZD_CHECK_ARG(n_levels >= 0 && n_levels <= KZD_MAX_LEVELS && (n_levels == 0 || h_levels),
                  "kzd_tal_assign: bad level table");
    KZD_CHECK_ARG(b <= 65535, "kzd_tal_assign: b=%d > 65535", b);
    KZD_CHECK_ARG(score_dtype >= 0 && score_dtype <= 2, "kzd_tal_assign: score_dtype=%d not in {0,1,2}", score_dtype);
    KZD_CHECK_ARG((double)b * A < 4294967295.0 && G < (1 << 24) - 2, "kzd_tal_assign: b*A or G too large");
    KZD_CHECK_ARG((size_t)G * 24 <= 200 * 1024, "kzd_tal_assign: G=%d exceeds the shared-memory GT table (8533)", G);
    KZD_CHECK_ARG(((uintptr_t)gt_bboxes) % 16 == 0 && ((uintptr_t)pd_bboxes) % 16 == 0 &&
                      ((uintptr_t)target_bboxes) % 16 == 0 && ((uintptr_t)anc_points) % 8 == 0 &&
                      ((uintptr_t)workspace) % 256 == 0,
                  "kzd_tal_assign: misaligned tensor");
    if (b == 0) return 0;
    AssignParams p;
    size_t need = 0;
    split_ws(workspace, b, A, G, p.ws, need);
    if (workspace_bytes < need) {
        set_error("kzd_tal_assign: workspace %zu B < required %zu B", workspace_bytes, need);
        return KZD_ENOSPACE;
    }
    int off = 0;
    for (int l = 0; l < n_levels; ++l) {
        KZD_CHECK_ARG(h_levels[l].h > 0 && h_levels[l].w > 0 && h_levels[l].offset == off && h_levels[l].stride > 0.f,
                      "kzd_tal_assign: level %d inconsistent", l);
        p.lv[l] = h_levels[l];
        off += h_levels[l].h * h_levels[l].w;
    }
    KZD_CHECK_ARG(n_levels == 0 || off == A, "kzd_tal_assign: levels cover %d anchors, A=%d", off, A);
    p.pd_scores = pd_scores, p.pd_bboxes = pd_bboxes, p.anc = anc_points;
    p.gt_labels = gt_labels, p.gt_bboxes = gt_bboxes, p.mask_gt = mask_gt;
    p.b = b, p.A = A, p.G = G, p.nc = nc, p.topk = topk;
    p.alpha = alpha, p.beta = beta, p.eps = eps;
    p.n_levels = n_levels, p.zero_fill = zero_fill, p.score_dtype = score_dtype;
    p.target_labels = (long long*)target_labels, p.target_bboxes = target_bboxes, p.target_scores = target_scores;
    p.fg_mask = fg_mask, p.target_gt_idx = (long long*)target_gt_idx, p.counters = counters;

    cudaStream_t st = (cudaStream_t)stream;
    if (!persistent) KZD_CUDA(cudaMemsetAsync(workspace, 0, p.ws.state_bytes, st));  // stateless mode
    assign_begin_kernel<<<1, 1, 0, st>>>(p.ws, persistent ? 0 : 1);
    KZD_LAUNCH_CHECK("assign_begin_kernel");
    if (n_levels > 0) {
        assign_verify_kernel<<<(A + 255) / 256, 256, 0, st>>>(anc_points, A, p);
        KZD_LAUNCH_CHECK("assign_verify_kernel");
    }
    if (nc > 1) {  // dense (b,A,nc) target_scores: zero rows now, the owner's entry is written later
        const size_t n = (size_t)b * A * nc;
        const size_t n4 = (((uintptr_t)target_scores) % 16 == 0) ? n / 4 : 0;
        const int n_tail = (int)(n - n4 * 4);
        if (n_tail > 256) {  // unaligned buffer: plain memset
            KZD_CUDA(cudaMemsetAsync(target_scores, 0, n * 4, st));
        } else {
            size_t blocks = (n4 + 255) / 256;
            const size_t cap = (size_t)sm_count() * 32;
            if (blocks > cap) blocks = cap;
            if (blocks < 1) blocks = 1;
            zero_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>((float4*)target_scores, n4, target_scores + n4 * 4,
                                                               n_tail);
            KZD_LAUNCH_CHECK("zero_fill_kernel");
        }
    }
    {
        const long long warps = (long long)b * G;
        const long long blocks = (warps * 32 + CAND_THREADS - 1) / CAND_THREADS;
        assign_candidates_kernel<<<(unsigned)blocks, CAND_THREADS, 0, st>>>(p);
        KZD_LAUNCH_CHECK("assign_candidates_kernel");
    }
    {
        const size_t smem = (size_t)G * 24;
        KZD_CUDA(cudaFuncSetAttribute(assign_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((A + RES_THREADS * RES_ITERS - 1) / (RES_THREADS * RES_ITERS), b);
        assign_resolve_kernel<<<grid, RES_THREADS, smem, st>>>(p);
        KZD_LAUNCH_CHECK("assign_resolve_kernel");
    }
    {
        assign_fg_metric_kernel<<<(unsigned)sm_count() * 2, 256, 0, st>>>(p);
        KZD_LAUNCH_CHECK("assign_fg_metric_kernel");
        assign_normalize_kernel<<<(unsigned)sm_count() * 2, 256, 0, st>>>(p);
        KZD_LAUNCH_CHECK("assign_normalize_kernel");
    }
    return 0;
}
