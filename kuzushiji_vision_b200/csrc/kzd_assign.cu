// TaskAlignedAssigner (sm_100a), GT-centric and sparse: no dense (b,G,A) tensor is ever formed.
//
// reference: utils/tal.py:39-295 (forward/_forward :39-118, get_pos_mask :120-130, get_box_metrics
// :132-151, select_topk_candidates :157-190, get_targets :192-238, select_candidates_in_gts :241-262,
// select_highest_overlaps :264-295) and CIoU utils/metrics.py:96-130.
//
// Kernels (one stream, no host sync; three launches in the steady state):
//   assign_verify_kernel      only when the caller does not vouch for anc_points: checks them against the
//                             level table bit for bit (the grid accelerator is used only when every anchor
//                             matches)
//   assign_candidates_kernel  one half warp (topk <= 16) or warp per (image, GT), the groups of a warp in
//                             lockstep: the anchors inside the GT are an exact rectangle per level (separable
//                             x / y tests); CIoU + alignment metric a batch of pairs at a time, top-k by
//                             group-wide max, selecting GTs recorded with two atomicMax words per picked anchor
//                             and in a pick list; each group also stores the background targets of a slice of
//                             the image's anchors (no dense pass over the anchors remains)
//   assign_patch_kernel       pick list (~10 per GT): resolve each picked anchor's owner GT (a contested
//                             anchor - picked by several GTs - takes the first-max argmax of overlaps over
//                             ALL GTs of the image: a scan of the GT rows, or a strip index of the page when a
//                             CTA meets many of them), overwrite its targets, owner pair's metric + per-GT
//                             maxima; the entry that swaps the anchor's selection words for zero represents
//                             it, so the persistent state is all-zero again when the call ends: no begin
//                             kernel, no epoch, no per-call memset
//   assign_normalize_kernel   one thread per pick slot: target_scores = align * pos_ov / (pos_align + eps),
//                             the counters; the last CTA of an image to leave clears the image's state words
//
// Bit-exactness: every fp32 operation of the eager reference rounds separately, so the pair
// arithmetic below uses __fadd_rn/__fmul_rn/__fdiv_rn (never contracted into FMAs) and the same
// libdevice atanf / powf / sqrtf that aten's CUDA kernels call.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kzd_common.cuh"

namespace kzd {

// Persistent per-anchor state: two words per anchor, all-zero between calls.
//   sel = { max over selecting GTs of (0xffffffff - g),  max over selecting GTs of (g + 1) }
//   => smallest and largest GT whose top-k selected the anchor; equal => exactly one selector.
struct AssignWs {
    uint2* sel;          // (b*A)  see above; set by the candidates kernel, cleared by the patch kernel's tail
    int32_t* flags;      // [0] != 0: anc_points does not match the level table (set by assign_verify_kernel,
                         //           consumed by the candidates kernel, cleared by the patch kernel's tail);
                         // [1] images finished (ticket of the patch kernel's tail, self-resetting);
                         // [2] copy of [0] as the last call left it (the host may poll it, kzd_assign_report_offset)
    int32_t* img;        // (b*2)  per image: [0] side-list length, [1] CTA ticket of the normalise kernel; both self-resetting
    uint32_t* pos_align; // (b*G)  per-GT max alignment metric (float bits)
    uint32_t* pos_ov;    // (b*G)  per-GT max overlap (float bits)
    uint4* fg_list;      // (b*A)  per-image side list (image i at i*A) of foreground anchors owned by another GT
                         //        than the pick slot's: (flat anchor index, flat GT slot, metric bits, class)
    uint2* picks;        // (b*G*topk)  every GT's selected anchors: (flat anchor index, g); x = ~0: empty slot.
                         //        The patch kernel rewrites a slot as (anchor, metric bits) when its GT owns the anchor
    size_t state_bytes;  // sel + flags + img: the part that must be zero when a call starts
};

struct AssignParams {
    const void* pd_scores;  // fp32, or score_dtype elements when scores_native
    const float *pd_bboxes, *anc, *gt_labels, *gt_bboxes, *mask_gt;
    int b, A, G, nc, topk;
    float alpha, beta, eps;
    kzd_level lv[KZD_MAX_LEVELS];
    float inv_stride[KZD_MAX_LEVELS];
    int n_levels, zero_fill;
    int fill_vec;     // A % 16 == 0 and 16-byte aligned outputs: background targets are stored as 16 B words
    int fill_per;     // anchors whose background targets one GT's warp stores
    int score_dtype;  // dtype pd_scores had at the call site: 0 fp32, 1 fp16, 2 bf16 (AMP, utils/tal.py:137)
    int scores_native;  // 0: pd_scores holds fp32; 1 / 2: it holds fp16 / bf16 elements (KZD_ASSIGN_NATIVE_SCORES), read as they are
    int strip_words;  // 32-bit words of one strip's GT mask in the patch kernel's index (0: index not used)
    float inv_strip;  // strips per pixel
    long long* target_labels;
    float* target_bboxes;
    float* target_scores;
    uint8_t* fg_mask;
    long long* target_gt_idx;
    double* counters;
    AssignWs ws;
};

// ---- arithmetic shared by the candidates and patch kernels (must be one instruction sequence) ----

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

// pd_scores element i as fp32: the AMP flow hands the assigner half-precision scores (utils/loss.py:233 under
// autocast), read here without an up-cast pass over the (b,A,nc) tensor
__device__ __forceinline__ float load_score(const void* base, size_t i, int native) {
    if (native == 1) return __half2float(__ldg(reinterpret_cast<const __half*>(base) + i));
    if (native == 2) return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + i));
    return __ldg(reinterpret_cast<const float*>(base) + i);
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

// ---- verify: anchor check (only when the caller does not vouch for anc_points) ----------------------

// One thread per anchor: anc_points against the level table, bit for bit. Any mismatch raises flags[0]; the
// candidates kernel then scans all anchors per GT instead of enumerating grid rectangles, and the patch
// kernel's tail clears the flag again.
__global__ void assign_verify_kernel(const float* __restrict__ anc, int A, AssignParams p) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= A) return;
    const float2 ap = __ldg(reinterpret_cast<const float2*>(anc) + a);
    int l = 0;
    for (int k = 1; k < p.n_levels; ++k)
        if (a >= p.lv[k].offset) l = k;
    const int q = a - p.lv[l].offset;
    const float s = p.lv[l].stride;
    const float ex = __fmul_rn((float)(q % p.lv[l].w) + 0.5f, s);
    const float ey = __fmul_rn((float)(q / p.lv[l].w) + 0.5f, s);
    if (ap.x != ex || ap.y != ey) p.ws.flags[0] = 1;
}

// ---- candidates + top-k ----------------------------------------------------------------------

#ifndef KZD_CAND_THREADS
#define KZD_CAND_THREADS 256
#endif
#ifndef KZD_CAND_MINB
#define KZD_CAND_MINB (1024 / KZD_CAND_THREADS)
#endif
constexpr int CAND_THREADS = KZD_CAND_THREADS;
constexpr int CAND_WARPS = CAND_THREADS / 32;
constexpr int CAND_CAP = 512;  // in-GT pairs a warp holds at a time (anchor + metric bits in shared memory)

// GT `g` selected the anchor: two fire-and-forget atomicMax (smallest and largest selecting GT).
__device__ __forceinline__ void mark_selected(uint2* __restrict__ sel, int g) {
    atomicMax(&sel->x, 0xffffffffu - (uint32_t)g);
    atomicMax(&sel->y, (uint32_t)g + 1u);
}

// One in-GT pair: the alignment metric's bits (0: not positive; positive floats order like their bit patterns).
// STD: alpha = 0.5, beta = 6 as literals, so the compiler specialises the inlined powf for the constant exponent
// (same libdevice algorithm, same bits, fewer instructions than the generic call).
template <bool STD>
__device__ __forceinline__ uint32_t eval_pair(const AssignParams& p, const GtBox& gt, const float4 pb, const float sc) {
    const float o = overlap_ciou(gt, pb);
    float al;
    if (STD) {
        float v = sqrtf(sc);
        if (p.score_dtype == 1) v = __half2float(__float2half_rn(v));
        else if (p.score_dtype == 2) v = __bfloat162float(__float2bfloat16_rn(v));
        al = __fmul_rn(v, powf(o, 6.0f));
    } else {
        al = __fmul_rn(score_pow(sc, p.alpha, p.score_dtype), pow_scalar(o, p.beta));
    }
    return al > 0.f ? __float_as_uint(al) : 0u;
}

// The columns (or rows) of one level whose centre passes both strict fp32 tests of utils/tal.py:259-262
// against [lo_edge, hi_edge]. Each test is monotone in the cell index (the centre (i+0.5)*s is), so the
// passing cells are exactly [first cell passing the low test, last cell passing the high test]; the
// float estimate is only a starting point and is walked to the exact boundary in either direction.
__device__ __forceinline__ void cell_range(float lo_edge, float hi_edge, float s, float inv_s, int n_cells, int& first,
                                           int& last) {
    auto low_ok = [&](int i) { return __fsub_rn(__fmul_rn((float)i + 0.5f, s), lo_edge) > 1e-9f; };
    auto high_ok = [&](int i) { return __fsub_rn(hi_edge, __fmul_rn((float)i + 0.5f, s)) > 1e-9f; };
    int i = (int)fminf(fmaxf(floorf(lo_edge * inv_s - 0.5f), 0.f), (float)(n_cells - 1));
    while (i > 0 && low_ok(i - 1)) --i;
    while (i < n_cells && !low_ok(i)) ++i;
    first = i;  // n_cells: none
    int j = (int)fminf(fmaxf(ceilf(hi_edge * inv_s - 0.5f), 0.f), (float)(n_cells - 1));
    while (j < n_cells - 1 && high_ok(j + 1)) ++j;
    while (j >= 0 && !high_ok(j)) --j;
    last = j;  // -1: none
}

// One GROUP of GW lanes per (image, GT): GW = 32 (a warp per GT) or GW = 16 (two GTs per warp).
//  1. produce: the anchors inside the GT. The in-GT predicate (utils/tal.py:259-262: min of the four fp32
//     differences > 1e-9  <=>  each one is) separates into an x test per column and a y test per row, so on
//     a make_anchors grid the anchors are an exact rectangle per level, written to shared memory GW per
//     step in ascending anchor order (only ~108 of 136,000 anchors are ever touched).
//  2. evaluate: GW pairs at a time with all lanes busy - CIoU, s^alpha * o^beta; metric bits parked next to
//     the anchor.
//  3. select: topk rounds of group-wide max over the lanes' local maxima; ties go to the
//     lowest anchor index = lowest slot, as torch.topk breaks them on the device. Only the lane that owned
//     the winner rescans its slots.
// A GT with more than the group's slot capacity is handled in chunks: the chunk's winners are carried into the
// first slots of the next one (they have the lowest anchor indices, so the tie rule still holds).
// Why two GTs per warp: the kernel is a chain of dependent steps per GT and the register budget allows 32
// resident warps per SM, so 9,600 GTs on a warp each take two "waves" of that chain. Kuzushiji boxes hold
// ~100 anchors: 16 lanes keep the pair batches as full as 32 do, and every instruction the warp issues now
// serves two GTs.
// LOCKSTEP: the two groups of a warp never diverge at a collective. Every vote / shuffle / reduction below is a
// full-warp instruction whose result is read per group (segmented butterflies, shifted ballots), and every loop
// that contains one runs to the larger of the two groups' trip counts with the finished group predicated off.
// (The first two-GTs-per-warp version used half-warp member masks: the compiler then wraps each collective in
// a divergence protocol - MATCH.ANY / BRA.DIV / WARPSYNC.COLLECTIVE - and the groups drift apart, so each issued
// instruction served one GT again: 19.9 M warp instructions for 9,600 GTs against 23.9 M with a warp per GT.)
template <int GW>
__global__ void __launch_bounds__(CAND_THREADS, KZD_CAND_MINB) assign_candidates_kernel(const AssignParams p) {
    constexpr int GROUPS = 32 / GW;          // GTs per warp
    constexpr int CAP = CAND_CAP / GROUPS;   // slots per group
    constexpr int LV = GW / 8;               // level slots one ballot settles: lane = level slot x edge (4) x probe (2)
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned GMASK = GW == 32 ? 0xffffffffu : 0xffffu;
    __shared__ uint32_t s_bits[CAND_WARPS][CAND_CAP];
    __shared__ int s_anc[CAND_WARPS][CAND_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & (GW - 1), grp = lane / GW, gbase = grp * GW;  // lane inside the group, group inside the warp
    const uint32_t total = (uint32_t)p.b * (uint32_t)p.G;  // b * G < 2^32 is checked by the host
    // segmented collectives: all 32 lanes execute them, each group reads its own half
    auto gballot = [&](bool pred) -> unsigned { return (__ballot_sync(FULL, pred) >> gbase) & GMASK; };
    auto gmax = [&](uint32_t v) -> uint32_t {
        if constexpr (GW == 32) {
            return __reduce_max_sync(FULL, v);
        } else {
#pragma unroll
            for (int o = GW / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
            return v;
        }
    };
    auto gmin = [&](int v) -> int {
        if constexpr (GW == 32) {
            return __reduce_min_sync(FULL, v);
        } else {
#pragma unroll
            for (int o = GW / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
            return v;
        }
    };
    auto gor = [&](unsigned v) -> unsigned {
        if constexpr (GW == 32) {
            return __reduce_or_sync(FULL, v);
        } else {
#pragma unroll
            for (int o = GW / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(FULL, v, o);
            return v;
        }
    };
    auto wmax = [&](int v) -> int {  // group-uniform v -> the larger of the warp's groups
        if constexpr (GW == 32) return v;
        else return max(v, __shfl_xor_sync(FULL, v, 16));
    };
    const bool std_hyp = p.alpha == 0.5f && p.beta == 6.0f;  // the reference's hyper-parameters (utils/loss.py:168)
    // (a device-side work queue - groups pulling GTs from an atomic counter - was measured 16 us slower
    // than this fixed assignment: the queue's round trips sit on every group's critical path)
    for (unsigned long long it = 0;; ++it) {
    const unsigned long long wid_first =
        ((unsigned long long)blockIdx.x * CAND_WARPS + wib) * GROUPS + it * ((unsigned long long)gridDim.x * CAND_WARPS * GROUPS);
    if (wid_first >= total) return;  // warp-uniform
    const bool in_range = wid_first + grp < total;  // the second group of the very last warp may have no GT: it only follows
    const uint32_t wid = in_range ? (uint32_t)(wid_first + grp) : total - 1u;
    const int bi = (int)(wid / (uint32_t)p.G), g = (int)(wid - (uint32_t)bi * (uint32_t)p.G);  // 32-bit division
    if (in_range && gl == 0) p.ws.pos_align[wid] = 0u, p.ws.pos_ov[wid] = 0u;
    if (wid == 0 && gl == 0 && p.counters) p.counters[0] = 0.0, p.counters[1] = 0.0;
    const size_t img = (size_t)bi * p.A;
    uint2* __restrict__ picks = p.ws.picks + (size_t)wid * p.topk;
    auto fill_background = [&]() {
        // Background targets for this group's slice of the image's anchors (every anchor that nobody
        // selects takes GT 0's label and box, target_gt_idx 0, utils/tal.py:215-233): plain streaming
        // stores that ride along with the latency-bound pair work instead of a separate dense pass (issued
        // when the group has finished its GT, so they are spread over the kernel's lifetime).
        const int per = p.fill_per;  // ceil(A / G), in 16-anchor granules when fill_vec: every array's slice is whole 16 B words
        const int a_lo = min(p.A, g * per), a_hi = in_range ? min(p.A, a_lo + per) : a_lo;
        const float4 box0 = *reinterpret_cast<const float4*>(p.gt_bboxes + (size_t)bi * p.G * 4);
        long long lab0 = (long long)p.gt_labels[(size_t)bi * p.G];
        lab0 = lab0 < 0 ? 0 : lab0;
        if (p.fill_vec) {
            const size_t first = img + a_lo;
            const int len = a_hi - a_lo;
            auto fill16 = [&](void* base, int elem_bytes, uint4 pat) {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(base) + first * elem_bytes);
                for (int i = gl; i < (len * elem_bytes) >> 4; i += GW) dst[i] = pat;
            };
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            const uint32_t l_lo = (uint32_t)lab0, l_hi = (uint32_t)((unsigned long long)lab0 >> 32);
            fill16(p.target_bboxes, 16, make_uint4(__float_as_uint(box0.x), __float_as_uint(box0.y),
                                                   __float_as_uint(box0.z), __float_as_uint(box0.w)));
            fill16(p.target_labels, 8, make_uint4(l_lo, l_hi, l_lo, l_hi));
            fill16(p.target_gt_idx, 8, zero);
            fill16(p.fg_mask, 1, zero);
            if (p.nc == 1) fill16(p.target_scores, 4, zero);
        } else {
            for (int a = a_lo + gl; a < a_hi; a += GW) {
                const size_t idx = img + a;
                p.target_gt_idx[idx] = 0;
                p.target_labels[idx] = lab0;
                p.fg_mask[idx] = 0;
                reinterpret_cast<float4*>(p.target_bboxes)[idx] = box0;
                if (p.nc == 1) p.target_scores[idx] = 0.f;
            }
        }
    };
    // everything the group needs about its GT, loaded together (one memory round trip instead of three)
    const float gt_mask = p.mask_gt[wid], gt_label = p.gt_labels[wid];
    const float4 gt_box = *reinterpret_cast<const float4*>(p.gt_bboxes + (size_t)wid * 4);
    const int flag_mismatch = p.ws.flags[0];
    // a padded GT (mask 0) claims nothing and contributes zero overlap: its group produces no pairs and follows
    const bool valid = in_range && gt_mask != 0.f;

    const GtBox gt = make_gt(gt_box);
    int label = (int)gt_label;
    label = label < 0 ? 0 : (label >= p.nc ? p.nc - 1 : label);
    const bool grid_ok = p.n_levels > 0 && flag_mismatch == 0;  // warp-uniform
    const float2* __restrict__ anc = reinterpret_cast<const float2*>(p.anc);
    const float4* __restrict__ pdb = reinterpret_cast<const float4*>(p.pd_bboxes) + img;
    const size_t pds = img * p.nc + label;  // element index of (image, anchor 0, this GT's class) in pd_scores
    const int topk = p.topk;
    uint32_t* bits = s_bits[wib] + grp * CAP;
    int* ancs = s_anc[wib] + grp * CAP;
    int n = 0, n_eval = 0;        // slots filled / evaluated (group-uniform)
    uint32_t sel_bits = 0u;       // lane r < n_sel: the r-th best pair so far
    int sel_anc = -1, n_sel = 0;

    auto evaluate = [&]() {  // slots [n_eval, n): anchors -> metric bits
        __syncwarp();
        // the gathers of batch k+1 (predicted box 16 B, score 4 B per pair) are issued before batch k is
        // evaluated, so only the first batch's memory latency is exposed
        int a_nx = 0;
        float4 pb_nx = make_float4(0.f, 0.f, 0.f, 0.f);
        float sc_nx = 0.f;
        auto fetch = [&](int t) {
            if (t < n) {
                a_nx = ancs[t];
                pb_nx = __ldg(pdb + a_nx);
                sc_nx = load_score(p.pd_scores, pds + (size_t)a_nx * p.nc, p.scores_native);
            }
        };
        const int trips = wmax((n - n_eval + GW - 1) / GW);
        int e = n_eval;
        fetch(e + gl);
        for (int k = 0; k < trips; ++k, e += GW) {
            const int t = e + gl;
            const float4 pb = pb_nx;
            const float sc = sc_nx;
            fetch(t + GW);
            if (t < n) bits[t] = std_hyp ? eval_pair<true>(p, gt, pb, sc) : eval_pair<false>(p, gt, pb, sc);
        }
        n_eval = n;
        __syncwarp();
    };
    auto select = [&]() {  // top-k of slots [0, n) -> lanes 0..n_sel-1 of the group
        // every lane tracks the best TWO of its slots (t = gl, gl + GW, ...), so a lane that wins a round
        // moves on to its runner-up without touching shared memory; only a lane winning again rescans
        uint32_t lm = 0u, lm2 = 0u, nb = 0u;
        int lt = 0x7fffffff, lt2 = 0x7fffffff, nt = 0, cnt = 0;
        bool have2 = true;
        auto scan2 = [&]() {
            lm = lm2 = 0u, lt = lt2 = 0x7fffffff;
            for (int t = gl; t < n; t += GW) {
                const uint32_t v = bits[t];
                if (v > lm) lm2 = lm, lt2 = lt, lm = v, lt = t;  // strict: the lowest slot among equals comes first
                else if (v > lm2) lm2 = v, lt2 = t;
            }
            have2 = true;
        };
        scan2();
        bool done = false;  // group-uniform: the group's positive pairs are exhausted
        for (int r = 0; r < topk; ++r) {
            const uint32_t mx = gmax(lm);
            done = done || mx == 0u;
            if (__all_sync(FULL, done)) break;
            const unsigned who = gballot(!done && lm == mx);
            int owner = __ffs(who) - 1;
            const bool tie = (who & (who - 1u)) != 0u;
            if (__any_sync(FULL, tie)) {  // equal metrics in two lanes: the lower slot wins (slot t belongs to lane t % GW)
                const int m = gmin(!done && lm == mx ? lt : 0x7fffffff);
                if (tie) owner = m & (GW - 1);
            }
            const int ts = __shfl_sync(FULL, lt, gbase + (owner & (GW - 1)));
            if (!done) {
                if (gl == r) nb = mx, nt = ts;
                ++cnt;
                if (gl == owner) {
                    bits[ts] = 0u;
                    if (have2) lm = lm2, lt = lt2, have2 = false;
                    else scan2();
                }
            }
            __syncwarp();
        }
        sel_bits = nb, n_sel = cnt;
        sel_anc = gl < cnt ? ancs[nt] : -1;
    };
    auto make_room = [&]() {  // a chunk is full: keep its winners, drop the rest (both groups, whichever one is full)
        if (!__any_sync(FULL, n + GW > CAP)) return;
        evaluate();
        select();
        __syncwarp();
        if (gl < n_sel) bits[gl] = sel_bits, ancs[gl] = sel_anc;
        n = n_eval = n_sel;
        __syncwarp();
    };

    if (grid_ok) {
        // Exact cell ranges, one ballot for LV levels: lane = level slot x edge (x-low, x-high, y-low, y-high) x
        // probe (2). Each probe is the reference's own fp32 test on the (verified) anchor coordinate and the tests
        // are monotone in the cell index, so an edge is settled when its two probes - the cells the fp32 estimate
        // puts on either side of the boundary - come out (fail, pass) for a low edge or (pass, fail) for a high
        // one, or when the page border settles it. Anything else (the estimate off by a cell, about one edge in
        // 10^5) falls back to the scalar walk for that axis.
        const int q_lane = gl >> 3, e_lane = (gl >> 1) & 3, k_lane = gl & 1;
        const bool is_hi = e_lane & 1, is_y = e_lane >> 1;
        const float edge = is_y ? (is_hi ? gt.y2 : gt.y1) : (is_hi ? gt.x2 : gt.x1);
        for (int l0 = 0; l0 < p.n_levels; l0 += LV) {
            int c0 = 0;
            bool ok = false;
            if (l0 + q_lane < p.n_levels) {
                const int l = l0 + q_lane;
                const float s = p.lv[l].stride, x = edge * p.inv_stride[l] - 0.5f;
                const int ncell = is_y ? p.lv[l].h : p.lv[l].w;
                // low edge: last failing cell = floor(x); high edge: last passing cell = ceil(x) - 1
                const float est = is_hi ? ceilf(x) - 1.f : floorf(x);
                c0 = (int)fminf(fmaxf(est, -1.f), (float)(ncell - 1));
                const int c = c0 + k_lane;
                const float ctr = __fmul_rn((float)c + 0.5f, s);
                ok = c >= 0 && c < ncell && (is_hi ? (__fsub_rn(edge, ctr) > 1e-9f) : (__fsub_rn(ctr, edge) > 1e-9f));
            }
            const unsigned m_all = gballot(ok);
            for (int q = 0; q < LV && l0 + q < p.n_levels; ++q) {
            const int l = l0 + q;
            const float s = p.lv[l].stride, inv_s = p.inv_stride[l];
            const int W = p.lv[l].w, H = p.lv[l].h, off = p.lv[l].offset;
            const unsigned m = (m_all >> (8 * q)) & 0xffu;  // bits: x-low {0,1}, x-high {2,3}, y-low {4,5}, y-high {6,7}
            const int cx0 = __shfl_sync(FULL, c0, gbase + 8 * q), cx1 = __shfl_sync(FULL, c0, gbase + 8 * q + 2);
            const int cy0 = __shfl_sync(FULL, c0, gbase + 8 * q + 4), cy1 = __shfl_sync(FULL, c0, gbase + 8 * q + 6);
            int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
            if (valid) {
                // low edge, probes (c0, c0+1): first passing cell
                auto low = [](unsigned pr, int c, int ncell, int& first) {
                    if (pr == 2u) { first = c + 1; return true; }                        // (fail, pass)
                    if (pr & 1u) { first = c; return c == 0; }                            // c passes: settled only at the border
                    first = ncell;                                                        // (fail, fail): none, if the last cell was probed
                    return c + 1 >= ncell - 1;
                };
                // high edge, probes (c0, c0+1): last passing cell
                auto high = [](unsigned pr, int c, int ncell, int& last) {
                    if (pr == 1u) { last = c; return true; }                              // (pass, fail)
                    if (pr & 2u) { last = c + 1; return c + 1 == ncell - 1; }             // c+1 passes: settled only at the border
                    last = -1;                                                            // (fail, fail): none, if cell 0 was probed
                    return c <= 0;
                };
                const bool okx = low(m & 3u, cx0, W, i0) & high((m >> 2) & 3u, cx1, W, i1);
                const bool oky = low((m >> 4) & 3u, cy0, H, j0) & high((m >> 6) & 3u, cy1, H, j1);
                if (!okx) cell_range(gt.x1, gt.x2, s, inv_s, W, i0, i1);  // group-uniform, no collective inside
                if (!oky) cell_range(gt.y1, gt.y2, s, inv_s, H, j0, j1);
            }
            const bool some = valid && i0 <= i1 && j0 <= j1;
            const unsigned nx = some ? (unsigned)(i1 - i0 + 1) : 1u, cnt = some ? nx * (unsigned)(j1 - j0 + 1) : 0u;
            const int base = off + j0 * W + i0;
            // t / nx by multiplication: exact for t < 2^16 (magic = ceil(2^32 / nx), nx >= 2)
            const bool use_magic = cnt <= 65536u && nx >= 2u;
            const unsigned magic = use_magic ? 0xffffffffu / nx + 1u : 0u;
            const unsigned cnt_w = (unsigned)wmax((int)cnt);  // cnt < 2^31: a level has fewer cells
            for (unsigned r = 0; r < cnt_w; r += GW) {
                make_room();
                const unsigned t = r + gl;
                if (t < cnt) {
                    const unsigned row = use_magic ? __umulhi(t, magic) : t / nx;
                    ancs[n + gl] = base + (int)row * W + (int)(t - row * nx);
                }
                if (r < cnt) n += (int)min((unsigned)GW, cnt - r);
            }
            }  // level slot
        }
    } else {
        for (int t0 = 0; t0 < p.A; t0 += GW) {  // anchors are not a make_anchors grid: test them all
            const int a = t0 + gl;
            bool inside = false;
            if (valid && a < p.A) {
                const float2 ap = __ldg(anc + a);
                inside = anchor_in_gt(ap.x, ap.y, gt);
            }
            const unsigned m = gballot(inside);
            if (!__any_sync(FULL, m != 0u)) continue;
            make_room();
            if (inside) ancs[n + __popc(m & ((1u << gl) - 1u))] = a;
            n += __popc(m);
        }
    }
    evaluate();
    select();

    const bool has = valid && gl < n_sel;
    if (has) mark_selected(p.ws.sel + img + sel_anc, g);
    uint32_t pick = has ? (uint32_t)(img + sel_anc) : 0xffffffffu;  // lane r: the r-th entry of this GT's pick list
    if (p.zero_fill) {
        // torch.topk pads a short list with zero-metric entries, lowest anchor index first; a padded
        // pick only matters when that anchor lies inside this GT (utils/tal.py:128). The padding takes the
        // first `need` anchors that are not positive picks; at most n_sel of the anchors below it are, so
        // every padded anchor has index < need + n_sel = topk: lane a tests anchor a.
        const int need = valid ? topk - n_sel : 0;
        if (__any_sync(FULL, need > 0)) {
            const unsigned posm = gor(has && sel_anc < topk ? 1u << sel_anc : 0u);  // positive picks among anchors [0, topk)
            bool inside = false;
            if (need > 0 && gl < topk && gl < p.A && !((posm >> gl) & 1u) && __popc(~posm & ((1u << gl) - 1u)) < need) {
                const float2 ap = __ldg(anc + gl);
                inside = anchor_in_gt(ap.x, ap.y, gt);
            }
            const unsigned im = gballot(inside);  // padded picks that count, ascending anchor index
            if (inside) mark_selected(p.ws.sel + img + gl, g);
            const int k = gl - n_sel;  // the k-th of them goes to pick-list entry n_sel + k
            if (need > 0 && k >= 0 && k < __popc(im)) pick = (uint32_t)(img + __fns(im, 0, k + 1));
        }
    }
    if (in_range && gl < topk) picks[gl] = make_uint2(pick, (uint32_t)g);
    fill_background();  // last: the stores never sit in front of this group's latency-critical loads
    __syncwarp();
    }  // next GT
}

// ---- patch: foreground anchors, normalisation, state clean-up ---------------------------------------

constexpr int PATCH_THREADS = 512;
constexpr int PATCH_STRIPS = 64;               // vertical page strips of the contested-anchor index
#ifndef KZD_PATCH_STRIP_MIN
#define KZD_PATCH_STRIP_MIN 48
#endif
constexpr int PATCH_STRIP_MIN = KZD_PATCH_STRIP_MIN;  // (contested anchors) x (GT rows per thread) from which the index pays
constexpr size_t PATCH_INDEX_MAX = 64 * 1024;  // bytes of shared memory the index may take (GT boxes + strip masks)

__device__ __forceinline__ int strip_of(float x, float inv_strip) {
    return (int)fminf(fmaxf(floorf(x * inv_strip), 0.f), (float)(PATCH_STRIPS - 1));
}
constexpr int SCAN_ILP = 4;  // GT rows per lane requested together by the contested-anchor scan

// Grid (x, image): the CTAs of one image stride over that image's pick-list slots (G*topk, ~85 % filled),
// one thread per slot. An anchor picked by several GTs appears once per GT; the first entry to swap the
// anchor's selection words for zero (one 64-bit atomicExch: it reads both words and leaves the persistent
// state clean for the next call) represents it. The representative resolves the owner - one selector -> it;
// several (contested) -> the first-max argmax of overlaps over ALL GTs of the image (utils/tal.py:283-289;
// all-zero column -> GT 0), found by the whole CTA scanning the image's GT rows once for all contested
// anchors of the round: they are rare (tens per image), so this replaces one 64-bit atomicMax per in-GT
// pair of the whole batch. It then overwrites the
// background targets the candidates kernel stored for that anchor, recomputes the (anchor, owner) pair's
// alignment metric with the same instruction sequence as the candidates kernel (=> identical bits; this
// also covers anchors owned by a GT that did not top-k them) and records the per-GT maxima of
// utils/tal.py:112-114. The slot is rewritten as (anchor, metric bits) when the slot's own GT is the owner
// (the usual case), else the anchor goes to the image's short side list.
// The normalisation (utils/tal.py:111-116) follows in assign_normalize_kernel.
__global__ void __launch_bounds__(PATCH_THREADS, 2) assign_patch_kernel(const AssignParams p) {
    __shared__ int c_n;
    __shared__ float2 c_ap[PATCH_THREADS];              // contested anchors of the current round: centre,
    __shared__ float4 c_pb[PATCH_THREADS];              // predicted box,
    __shared__ unsigned long long c_key[PATCH_THREADS]; // best (overlap bits, ~g) found so far
    extern __shared__ __align__(16) unsigned char patch_dyn[];  // contested-anchor index: G boxes, PATCH_STRIPS x strip_words masks
    bool strips_built = false;
    const int bi = blockIdx.y;
    if (threadIdx.x == 0) c_n = 0;
    __syncthreads();
    const int n_slots = p.G * p.topk;  // of this image
    uint2* __restrict__ picks = p.ws.picks + (size_t)bi * n_slots;
    const float2* __restrict__ anc = reinterpret_cast<const float2*>(p.anc);
    const float4* __restrict__ gtb = reinterpret_cast<const float4*>(p.gt_bboxes) + (size_t)bi * p.G;
    const float* __restrict__ gtl = p.gt_labels + (size_t)bi * p.G;
    const float* __restrict__ gtm = p.mask_gt + (size_t)bi * p.G;
    uint4* __restrict__ side = p.ws.fg_list + (size_t)bi * p.A;
    int32_t* __restrict__ img_state = p.ws.img + 2 * bi;  // [0] side-list length, [1] CTA ticket
    for (int s0 = blockIdx.x * PATCH_THREADS; s0 < n_slots; s0 += gridDim.x * PATCH_THREADS) {  // whole warps iterate together
        const int s = s0 + threadIdx.x;
        bool rep = false, contested = false;
        size_t idx = 0;
        uint32_t g = 0u, g_own = 0u;
        float2 ap = make_float2(0.f, 0.f);
        float4 pb = make_float4(0.f, 0.f, 0.f, 0.f), gb = pb;
        float glab = 0.f, gmsk = 0.f, sc_pre = 0.f;
        if (s < n_slots) {
            const uint2 e = picks[s];
            if (e.x != 0xffffffffu) {
                // Everything that depends only on the pick is requested at once: the selection words, the
                // anchor, its predicted box, and - speculatively - the selecting GT's row (it is the owner
                // unless the anchor is contested). One memory round trip instead of four.
                idx = e.x, g_own = e.y;
                const int a = (int)(idx - (size_t)bi * p.A);
                const unsigned long long sel = atomicExch(reinterpret_cast<unsigned long long*>(p.ws.sel + idx), 0ull);
                ap = __ldg(anc + a);
                pb = __ldg(reinterpret_cast<const float4*>(p.pd_bboxes) + idx);
                gb = gtb[g_own], glab = gtl[g_own], gmsk = gtm[g_own];
                if (p.nc == 1) sc_pre = load_score(p.pd_scores, idx, p.scores_native);  // single class: the score's address does not wait for the label
                rep = sel != 0ull;  // the other entries of this anchor find zero
                const uint32_t g_min = 0xffffffffu - (uint32_t)sel, g_max = (uint32_t)(sel >> 32) - 1u;
                contested = rep && g_min != g_max;
                g = g_own;  // a single selector is this entry's GT
            }
        }
        // Contested anchors of this round, all of them at once and by the whole CTA: they are listed in shared
        // memory, every thread takes a few of the image's GT rows (one batch of independent loads) and tests them
        // against every listed anchor; the argmax over the image's GTs of clamp(CIoU, 0) - restricted to valid
        // GTs that contain the anchor, everything else is 0 in the reference's overlaps tensor - is a shared-memory
        // atomicMax on (overlap bits, ~g): highest overlap first, lowest GT among equals.
        int my_c = -1;
        if (contested) {
            my_c = atomicAdd(&c_n, 1);
            c_ap[my_c] = ap, c_pb[my_c] = pb, c_key[my_c] = 0ull;
        }
        __syncthreads();
        const int n_c = c_n;  // CTA-uniform
        if (n_c * ((p.G + PATCH_THREADS - 1) / PATCH_THREADS) >= PATCH_STRIP_MIN && p.strip_words > 0) {
            // Strip index: the page is cut into PATCH_STRIPS vertical strips; strip s holds a G-bit mask of the valid
            // GTs whose x-range reaches into it. strip(x) = clamp(floor(x * inv_strip)) is monotone in x, so a GT
            // that contains an anchor (x1 < ax < x2) is always listed in the anchor's strip: the argmax below tests
            // ~G * (box width + strip width) / page width GTs per anchor instead of all G (scanning every GT row
            // for every contested anchor, a CTA with 300 of them took 19 us at G = 1000 and set the kernel's
            // duration). Masks and a copy of the image's GT boxes live in shared memory; built once per CTA, the
            // first time a round's scan would cost more than building it (~4 us): the scan runs (contested anchors)
            // x (GT rows per thread) tests per thread. Measured per call, index always / never / from 48 tests:
            // 17.2 / 13.3 / 14.4 us at G = 600 (a CTA lists a dozen or two), 23.6 / 28.6 / 24.2 us at G = 1000.
            const int W = p.strip_words;
            float4* s_box = reinterpret_cast<float4*>(patch_dyn);
            uint32_t* strips = reinterpret_cast<uint32_t*>(s_box + p.G);
            if (!strips_built) {
                for (int i = threadIdx.x; i < PATCH_STRIPS * W; i += PATCH_THREADS) strips[i] = 0u;
                __syncthreads();
                for (int gg = threadIdx.x; gg < p.G; gg += PATCH_THREADS) {
                    const float4 bx = gtb[gg];
                    s_box[gg] = bx;
                    if (gtm[gg] == 0.f) continue;
                    const int s_lo = strip_of(bx.x, p.inv_strip), s_hi = strip_of(bx.z, p.inv_strip);
                    for (int st = s_lo; st <= s_hi; ++st) atomicOr(strips + st * W + (gg >> 5), 1u << (gg & 31));
                }
                strips_built = true;
                __syncthreads();
            }
            // four lanes per contested anchor: lane q takes the words w = q, q + 4, ... of the anchor's strip
            for (int base = 0; base < n_c; base += PATCH_THREADS / 4) {
                const int j = base + (threadIdx.x >> 2), q = threadIdx.x & 3;
                unsigned long long best = 0ull;
                if (j < n_c) {
                    const float2 a_ = c_ap[j];
                    const float4 pb_ = c_pb[j];
                    const uint32_t* row = strips + strip_of(a_.x, p.inv_strip) * W;
                    for (int w = q; w < W; w += 4) {
                        uint32_t bits = row[w];
                        while (bits) {
                            const uint32_t gg = (uint32_t)(w * 32 + __ffs(bits) - 1);
                            bits &= bits - 1u;
                            const float4 bx = s_box[gg];
                            GtBox corners;
                            corners.x1 = bx.x, corners.y1 = bx.y, corners.x2 = bx.z, corners.y2 = bx.w;
                            if (!anchor_in_gt(a_.x, a_.y, corners)) continue;
                            const uint32_t ob = __float_as_uint(overlap_ciou(make_gt(bx), pb_));
                            const unsigned long long k = ((unsigned long long)ob << 32) | (unsigned long long)(0xffffffffu - gg);
                            if (ob != 0u && k > best) best = k;  // highest overlap, lowest GT among equals
                        }
                    }
                }
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)best, o);
                    const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(best >> 32), o);
                    const unsigned long long other = ((unsigned long long)hi << 32) | lo;
                    best = other > best ? other : best;
                }
                if (j < n_c && q == 0) c_key[j] = best;
            }
        } else if (n_c > 0) {  // few contested anchors, no level table (page extent unknown) or no room for the index: every GT row
            for (int g0 = 0; g0 < p.G; g0 += PATCH_THREADS * SCAN_ILP) {
                float4 box[SCAN_ILP];
                float msk[SCAN_ILP];
#pragma unroll
                for (int k = 0; k < SCAN_ILP; ++k) {
                    const int gg = g0 + k * PATCH_THREADS + threadIdx.x;
                    msk[k] = 0.f;
                    box[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gg < p.G) msk[k] = gtm[gg], box[k] = gtb[gg];
                }
#pragma unroll
                for (int k = 0; k < SCAN_ILP; ++k) {
                    if (msk[k] == 0.f) continue;
                    const uint32_t gg = (uint32_t)(g0 + k * PATCH_THREADS + threadIdx.x);
                    GtBox corners;
                    corners.x1 = box[k].x, corners.y1 = box[k].y, corners.x2 = box[k].z, corners.y2 = box[k].w;
                    bool have_gt = false;
                    GtBox gt = corners;
                    for (int j = 0; j < n_c; ++j) {
                        const float2 a_ = c_ap[j];
                        if (!anchor_in_gt(a_.x, a_.y, corners)) continue;
                        if (!have_gt) gt = make_gt(box[k]), have_gt = true;
                        const uint32_t ob = __float_as_uint(overlap_ciou(gt, c_pb[j]));
                        if (ob != 0u) atomicMax(c_key + j, ((unsigned long long)ob << 32) | (unsigned long long)(0xffffffffu - gg));
                    }
                }
            }
        }
        if (n_c > 0) {
            __syncthreads();
            if (threadIdx.x == 0) c_n = 0;  // n_c is in registers; the next round's entries come after the barrier below
            if (my_c >= 0) {
                const unsigned long long k = c_key[my_c];
                const uint32_t owner = (k >> 32) != 0ull ? 0xffffffffu - (uint32_t)k : 0u;  // all-zero column -> GT 0
                if (owner != g) {
                    g = owner;
                    gb = gtb[g], glab = gtl[g], gmsk = gtm[g];
                }
            }
            __syncthreads();  // every reader of the list is done before the next round refills it
        }
        if (rep) {
            long long lab = (long long)glab;
            lab = lab < 0 ? 0 : lab;
            p.target_gt_idx[idx] = (long long)g;
            p.target_labels[idx] = lab;
            p.fg_mask[idx] = 1;
            reinterpret_cast<float4*>(p.target_bboxes)[idx] = gb;
            lab = lab >= p.nc ? p.nc - 1 : lab;
            const size_t gslot = (size_t)bi * p.G + g;
            float al = 0.f;
            if (gmsk != 0.f) {
                const GtBox gt = make_gt(gb);
                if (anchor_in_gt(ap.x, ap.y, gt)) {
                    const float sc = p.nc == 1 ? sc_pre : load_score(p.pd_scores, idx * p.nc + lab, p.scores_native);
                    const float o = overlap_ciou(gt, pb);
                    al = __fmul_rn(score_pow(sc, p.alpha, p.score_dtype), pow_scalar(o, p.beta));
                    atomicMax(p.ws.pos_ov + gslot, __float_as_uint(o));
                    atomicMax(p.ws.pos_align + gslot, __float_as_uint(al));
                }
            }
            if (g == g_own) {
                picks[s] = make_uint2((uint32_t)idx, __float_as_uint(al));  // normalised by assign_normalize_kernel
            } else {  // rare: owned by another GT than the slot's
                picks[s] = make_uint2(0xffffffffu, 0u);
                side[atomicAdd(img_state, 1)] = make_uint4((uint32_t)idx, (uint32_t)gslot, __float_as_uint(al), (uint32_t)lab);
            }
        } else if (s < n_slots) {
            picks[s] = make_uint2(0xffffffffu, 0u);
        }
    }
}

// ---- normalise: target_scores of the foreground anchors ------------------------------------------------

constexpr int NORM_THREADS = 256;

// Grid (x, image), one thread per pick slot of the image: target_scores = align * pos_ov / (pos_align + eps)
// (utils/tal.py:111-116) with the per-GT maxima the patch kernel has finished. The slot and its GT's maxima sit
// at computable addresses, so a thread's loads are all in flight together. CTA 0 of an image also takes the
// image's short side list; the Sum(target_scores) / n_fg counters are what utils/loss.py:243,250 reduce again.
// A separate launch on purpose: as the tail of the patch kernel (run by each image's last CTA after a ticket)
// the same work took 6-9 us of a 20 us kernel on one SM per image; spread over the machine it is one memory round
// trip. The last CTA of an image to leave (ticket, off the critical path) clears the image's state words, the
// last image clears the call's flags: the workspace is all-zero again when the call ends.
__global__ void __launch_bounds__(NORM_THREADS) assign_normalize_kernel(const AssignParams p) {
    __shared__ double s_sum;
    __shared__ int s_fg;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int bi = blockIdx.y;
    const int n_slots = p.G * p.topk;  // of this image
    const uint2* __restrict__ picks = p.ws.picks + (size_t)bi * n_slots;
    const float* __restrict__ gtl = p.gt_labels + (size_t)bi * p.G;
    const uint32_t* __restrict__ pos_ov = p.ws.pos_ov + (size_t)bi * p.G;
    const uint32_t* __restrict__ pos_al = p.ws.pos_align + (size_t)bi * p.G;
    int32_t* __restrict__ img_state = p.ws.img + 2 * bi;  // [0] side-list length, [1] CTA ticket
    if (threadIdx.x == 0) s_sum = 0.0, s_fg = 0;
    __syncthreads();
    float local = 0.f;
    int n_fg = 0;
    for (int s = blockIdx.x * NORM_THREADS + threadIdx.x; s < n_slots; s += gridDim.x * NORM_THREADS) {
        const uint2 e = picks[s];
        const int g = s / p.topk;
        const uint32_t ov = pos_ov[g], pa = pos_al[g];
        long long lab = p.nc > 1 ? (long long)gtl[g] : 0;
        if (e.x == 0xffffffffu) continue;
        lab = lab < 0 ? 0 : (lab >= p.nc ? p.nc - 1 : lab);
        const float v = __fdiv_rn(__fmul_rn(__uint_as_float(e.y), __uint_as_float(ov)),
                                  __fadd_rn(__uint_as_float(pa), p.eps));  // utils/tal.py:115
        p.target_scores[(size_t)e.x * p.nc + lab] = v;
        local += v;
        ++n_fg;
    }
    if (blockIdx.x == 0) {  // anchors owned by another GT than the slot's
        const uint4* __restrict__ side = p.ws.fg_list + (size_t)bi * p.A;
        const int n_side = img_state[0];
        for (int i = threadIdx.x; i < n_side; i += NORM_THREADS) {
            const uint4 e = side[i];
            const float ov = __uint_as_float(p.ws.pos_ov[e.y]);
            const float pa = __uint_as_float(p.ws.pos_align[e.y]);
            const float v = __fdiv_rn(__fmul_rn(__uint_as_float(e.z), ov), __fadd_rn(pa, p.eps));
            p.target_scores[(size_t)e.x * p.nc + e.w] = v;
            local += v;
            ++n_fg;
        }
    }
    if (p.counters) {
        for (int o = 16; o > 0; o >>= 1) {
            local += __shfl_xor_sync(FULL, local, o);
            n_fg += __shfl_xor_sync(FULL, n_fg, o);
        }
        if (lane == 0) {
            if (local != 0.f) atomicAdd(&s_sum, (double)local);
            if (n_fg) atomicAdd(&s_fg, n_fg);
        }
    }
    __syncthreads();  // (also: CTA 0 has read the side-list length before its ticket)
    if (threadIdx.x == 0) {
        if (p.counters) {
            if (s_sum != 0.0) atomicAdd(p.counters, s_sum);
            if (s_fg) atomicAdd(p.counters + 1, (double)s_fg);
        }
        if (atomicAdd(img_state + 1, 1) == (int)gridDim.x - 1) {  // the image's last CTA to leave
            img_state[0] = 0, img_state[1] = 0;
            if (atomicAdd(p.ws.flags + 1, 1) == (int)gridDim.y - 1) {  // the call's last image
                p.ws.flags[2] = p.ws.flags[0];  // sticky report for the host: this call's level table did not match
                p.ws.flags[1] = 0, p.ws.flags[0] = 0;
            }
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
    // state first: its layout does not depend on G, so one workspace sized for the largest G seen serves
    // every smaller G with the same (b, A)
    ws.sel = (uint2*)take(n * 8);
    ws.flags = (int32_t*)take(256);
    ws.img = (int32_t*)take((size_t)b * 2 * 4);
    ws.state_bytes = o;
    ws.pos_align = (uint32_t*)take((size_t)b * G * 4);
    ws.pos_ov = (uint32_t*)take((size_t)b * G * 4);
    ws.fg_list = (uint4*)take(n * 16);
    ws.picks = (uint2*)take((size_t)b * G * KZD_MAX_TOPK * 8);
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

extern "C" __attribute__((visibility("default"))) size_t kzd_assign_report_offset(int b, int A) {
    if (b < 0 || A < 0) return 0;
    return align_up((size_t)b * A * 8, 256) + 2 * sizeof(int32_t);
}

extern "C" __attribute__((visibility("default"))) int kzd_tal_assign(const void* pd_scores, const float* pd_bboxes, const float* anc_points,
                              const float* gt_labels, const float* gt_bboxes, const float* mask_gt, int b, int A,
                              int G, int nc, int topk, float alpha, float beta, float eps, const kzd_level* h_levels,
                              int n_levels, int zero_fill, int score_dtype, int64_t* target_labels, float* target_bboxes,
                              float* target_scores, uint8_t* fg_mask, int64_t* target_gt_idx, double* counters,
                              void* workspace, size_t workspace_bytes, int flags, void* stream) {
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
    KZD_CHECK_ARG((flags & ~(KZD_ASSIGN_PERSISTENT | KZD_ASSIGN_TRUSTED_ANCHORS | KZD_ASSIGN_NATIVE_SCORES | KZD_ASSIGN_WARP_PER_GT)) == 0, "kzd_tal_assign: unknown flags 0x%x", flags);
    KZD_CHECK_ARG((double)b * A < 4294967295.0 && G < (1 << 24) - 2, "kzd_tal_assign: b*A or G too large");
    KZD_CHECK_ARG((double)b * G < 4294967295.0, "kzd_tal_assign: b*G too large");
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
        p.inv_stride[l] = 1.f / h_levels[l].stride;
        off += h_levels[l].h * h_levels[l].w;
    }
    KZD_CHECK_ARG(n_levels == 0 || off == A, "kzd_tal_assign: levels cover %d anchors, A=%d", off, A);
    p.pd_scores = pd_scores, p.pd_bboxes = pd_bboxes, p.anc = anc_points;
    p.gt_labels = gt_labels, p.gt_bboxes = gt_bboxes, p.mask_gt = mask_gt;
    p.b = b, p.A = A, p.G = G, p.nc = nc, p.topk = topk;
    p.alpha = alpha, p.beta = beta, p.eps = eps;
    p.n_levels = n_levels, p.zero_fill = zero_fill, p.score_dtype = score_dtype;
    p.scores_native = (flags & KZD_ASSIGN_NATIVE_SCORES) ? score_dtype : 0;
    p.fill_vec = A % 16 == 0 && ((uintptr_t)target_labels | (uintptr_t)target_gt_idx | (uintptr_t)fg_mask |
                                 (uintptr_t)target_scores | (uintptr_t)target_bboxes) % 16 == 0;
    p.fill_per = (A + G - 1) / G;
    if (p.fill_vec) p.fill_per = (p.fill_per + 15) & ~15;
    p.target_labels = (long long*)target_labels, p.target_bboxes = target_bboxes, p.target_scores = target_scores;
    p.fg_mask = fg_mask, p.target_gt_idx = (long long*)target_gt_idx, p.counters = counters;

    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & KZD_ASSIGN_PERSISTENT)) KZD_CUDA(cudaMemsetAsync(workspace, 0, p.ws.state_bytes, st));  // stateless mode
    if (n_levels > 0 && !(flags & KZD_ASSIGN_TRUSTED_ANCHORS)) {
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
        // two GTs per warp (16 lanes each) whenever the pick list fits the half warp; a warp per GT otherwise
#ifndef KZD_CAND_GW
#define KZD_CAND_GW 16
#endif
        // a warp per GT is the faster form when the call has the machine to itself (36.8 vs 43.5 us at G = 600, 53 vs 63 us at
        // G = 1000); two GTs per warp when other kernels run next to it (the graph pipelines: 153.2 vs 156.8 us per step)
        const int gw = (topk <= 16 && KZD_CAND_GW == 16 && !(flags & KZD_ASSIGN_WARP_PER_GT)) ? 16 : 32;
        const long long warps = ((long long)b * G + (32 / gw) - 1) / (32 / gw);
        long long blocks = (warps + CAND_WARPS - 1) / CAND_WARPS;
        if (blocks > 0x7fffffffLL) blocks = 0x7fffffffLL;  // the kernel strides over the rest
        // persistent grid: no more CTAs than the SMs hold at once (4 per SM); their groups stride over the GTs.
        // Measured with four page batches in flight (tools/pipe_value.py, G = 600 / 1000): 150.7 / 180.4 us per step
        // against 150.4 / 186.3 with one CTA per eight groups, and 159.3 / 182.8 with a warp per GT.
#ifndef KZD_CAND_PERSIST
#define KZD_CAND_PERSIST 4
#endif
        if (KZD_CAND_PERSIST > 0 && blocks > (long long)sm_count() * KZD_CAND_PERSIST) blocks = (long long)sm_count() * KZD_CAND_PERSIST;
        if (gw == 16) assign_candidates_kernel<16><<<(unsigned)blocks, CAND_THREADS, 0, st>>>(p);
        else assign_candidates_kernel<32><<<(unsigned)blocks, CAND_THREADS, 0, st>>>(p);
        KZD_LAUNCH_CHECK("assign_candidates_kernel");
    }
    {
        // the CTAs of an image stride over its G*topk pick slots; enough of them to cover the slots once, but
        // not more than a few waves of the machine in total
        long long per_img = ((long long)G * topk + PATCH_THREADS - 1) / PATCH_THREADS;
        const long long cap = ((long long)sm_count() * 8 + b - 1) / b;
        if (per_img > cap) per_img = cap;
        if (per_img < 1) per_img = 1;
        p.strip_words = 0, p.inv_strip = 0.f;
        float extent = 0.f;
        for (int l = 0; l < n_levels; ++l) extent = fmaxf(extent, (float)h_levels[l].w * h_levels[l].stride);
        const size_t index_bytes = (size_t)G * 16 + (size_t)PATCH_STRIPS * ((G + 31) / 32) * 4;
#ifndef KZD_NO_STRIPS
#define KZD_NO_STRIPS 0
#endif
        if (!KZD_NO_STRIPS && extent > 0.f && index_bytes <= PATCH_INDEX_MAX) {
            p.strip_words = (G + 31) / 32, p.inv_strip = (float)PATCH_STRIPS / extent;
            // (per call: the attribute belongs to the current device's copy of the kernel)
            KZD_CUDA(cudaFuncSetAttribute(assign_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PATCH_INDEX_MAX));
        }
        assign_patch_kernel<<<dim3((unsigned)per_img, (unsigned)b), PATCH_THREADS, p.strip_words ? index_bytes : 0, st>>>(p);
        KZD_LAUNCH_CHECK("assign_patch_kernel");
        long long per_img_n = ((long long)G * topk + NORM_THREADS - 1) / NORM_THREADS;
        const long long cap_n = ((long long)sm_count() * 8 + b - 1) / b;
        if (per_img_n > cap_n) per_img_n = cap_n;
        if (per_img_n < 1) per_img_n = 1;
        assign_normalize_kernel<<<dim3((unsigned)per_img_n, (unsigned)b), NORM_THREADS, 0, st>>>(p);
        KZD_LAUNCH_CHECK("assign_normalize_kernel");
    }
    return 0;
}
