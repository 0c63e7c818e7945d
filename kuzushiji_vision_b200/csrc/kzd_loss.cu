// Loss tail on the assigned targets (sm_100a): BCE over (b,A,nc) + CIoU / DFL losses on the
// foreground anchors, forward and backward, without materialising any per-foreground gather
// (each CTA compacts the foreground anchors of its span of 2048 anchors in shared memory and gives them
// four lanes each).
//
// reference: utils/loss.py:243-258 (v8DetectionLoss.__call__ after the assigner), BboxLoss.forward
// :91-113, DFLoss :65-88, bbox2dist utils/tal.py:360-363, bbox_iou(CIoU) utils/metrics.py:96-130
// (called as bbox_iou(pred, target): box1 = prediction, box2 = target; alpha is a no_grad constant),
// nn.BCEWithLogitsLoss(reduction="none").
//
//   loss[0] = sum_fg (1 - CIoU(pred_box, target_box/stride)) * w / tss          (box, before hyp.box)
//   loss[1] = sum_{b,a,c} BCEWithLogits(pred_scores, target_scores) / tss        (cls)
//   loss[2] = sum_fg mean_side(CE(dist, tl) wl + CE(dist, tr) wr) * w / tss      (dfl)
//   w = target_scores.sum(-1) of the anchor, tss = max(target_scores.sum(), 1)
// The predicted box is re-decoded from pred_dist inside the kernel (softmax expectation), so the
// backward chains through the decode directly into grad_pred_dist.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kzd_common.cuh"

namespace kzd {

struct LossParams {
    const void *pred_scores, *pred_dist;  // fp32, or the AMP flow's fp16 / bf16 tensors themselves (logits_dtype, dist_dtype)
    const float *anchor_points, *stride, *target_bboxes, *target_scores;
    const uint8_t* fg_mask;
    const float* tss;  // device scalar: target_scores.sum() (max(., 1) is applied here)
    int b, A, nc;
    // sparse-target form: the assigner's target_scores row of a foreground anchor is one-hot * norm at its
    // target label, so target_labels tells where the single non-zero sits and the dense (b,A,nc) tensor is
    // only read at those (anchor, label) elements
    const long long* target_labels;  // nullptr: dense form
    int score_dtype;                 // dtype the BCE target is rounded to (utils/loss.py:247 under AMP): 0 fp32, 1 fp16, 2 bf16
    int logits_dtype;                // element type of pred_scores and grad_pred_scores: 0 fp32, 1 fp16, 2 bf16
    int dist_dtype;                  // element type of pred_dist and grad_pred_dist
};

// element i of a tensor of the given dtype code, as fp32 / fp32 value v rounded into element i
__device__ __forceinline__ float load_elem(const void* base, size_t i, int dtype) {
    if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(base)[i]);
    if (dtype == 2) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
    return reinterpret_cast<const float*>(base)[i];
}
__device__ __forceinline__ void store_elem(void* base, size_t i, int dtype, float v) {
    if (dtype == 1) reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
    else if (dtype == 2) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(base)[i] = v;
}


__device__ __forceinline__ float round_to(float v, int dtype) {
    if (dtype == 1) return __half2float(__float2half_rn(v));
    if (dtype == 2) return __bfloat162float(__float2bfloat16_rn(v));
    return v;
}

// ---- BCE ----------------------------------------------------------------------------------------

__device__ __forceinline__ float bce_logits(float x, float t) {
    // aten binary_cross_entropy_with_logits: (1 - t) x + max(-x, 0) + log(exp(-max(-x,0)) + exp(-x - max(-x,0))),
    // whose last two terms are softplus(-x) = max(-x, 0) + log1p(exp(-|x|)); the log1p form keeps the tail of
    // very negative logits that the sum of exponentials rounds away
    return (1.f - t) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
}

// BCE(x, 0) = softplus(x) = max(x, 0) + log1p(exp(-|x|)); BCE(x, t) = softplus(x) - t x
__device__ __forceinline__ float softplus_fast(float x) {
    const float e = __expf(-fabsf(x));
    // log1p(e), e in (0, 1]: short series where 1 + e would round the tail away, fast log elsewhere
    const float l = e < 9.765625e-4f ? e * (1.f - e * (0.5f - e * 0.33333334f)) : __logf(1.f + e);
    return fmaxf(x, 0.f) + l;
}

__device__ __forceinline__ void block_add(double v, double* dst, double* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
        if (t != 0.0) atomicAdd(dst, t);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                      size_t n, double* __restrict__ sums) {
    __shared__ double s_red[8];
    float acc = 0.f;
    double total = 0.0;
    int k = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        acc += bce_logits(__ldcs(x + i), __ldcs(t + i));
        if (++k == 64) total += (double)acc, acc = 0.f, k = 0;  // bound the fp32 partial
    }
    total += (double)acc;
    block_add(total, sums + 1, s_red);
}

// Sparse-target BCE, forward: sum softplus(x) over ALL logits (the only dense read: pred_scores, 128-bit
// streaming loads); the -t x terms of the foreground (anchor, label) elements are added by fg_fwd_kernel.
__global__ void __launch_bounds__(256) softplus_sum_kernel(const float* __restrict__ x, size_t n, size_t n4,
                                                           double* __restrict__ sums) {
    __shared__ double s_red[8];
    float acc = 0.f;
    double total = 0.0;
    int k = 0;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(x4 + i);
        acc += (softplus_fast(v.x) + softplus_fast(v.y)) + (softplus_fast(v.z) + softplus_fast(v.w));
        if (++k == 16) total += (double)acc, acc = 0.f, k = 0;  // bound the fp32 partial
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += softplus_fast(__ldcs(x + i));
    total += (double)acc;
    block_add(total, sums + 1, s_red);
}

// Sparse-target BCE, backward: sigmoid(x) * scale everywhere (read x, write the gradient; target_scores is
// not read); fg_bwd_kernel then subtracts t * scale at the foreground (anchor, label) elements.
__global__ void __launch_bounds__(256) sigmoid_grad_kernel(const float* __restrict__ x, size_t n, size_t n4,
                                                           const float* __restrict__ tss, const float* __restrict__ gout,
                                                           float* __restrict__ gx) {
    const float scale = gout[1] / fmaxf(*tss, 1.f);
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    float4* __restrict__ g4 = reinterpret_cast<float4*>(gx);
    auto sg = [&](float v) { return __fdividef(scale, 1.f + __expf(-v)); };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(x4 + i);
        __stcs(g4 + i, make_float4(sg(v.x), sg(v.y), sg(v.z), sg(v.w)));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        __stcs(gx + i, sg(__ldcs(x + i)));
}

// The same two passes on the AMP flow's half-precision logits, read (and the gradient written) as they are: 16-byte
// words of eight elements, fp32 arithmetic, one rounding of the gradient into the logits' dtype - what autograd
// does behind the reference's fp32-autocast BCE (the gradient comes back through the cast of pred_scores).
template <int DT>
__global__ void __launch_bounds__(256) softplus_sum_half_kernel(const void* __restrict__ x, size_t n, size_t n8,
                                                                double* __restrict__ sums) {
    __shared__ double s_red[8];
    float acc = 0.f;
    double total = 0.0;
    int k = 0;
    const uint4* __restrict__ x8 = reinterpret_cast<const uint4*>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float v[8];
        unpack8<DT>(__ldcs(x8 + i), v);
        acc += ((softplus_fast(v[0]) + softplus_fast(v[1])) + (softplus_fast(v[2]) + softplus_fast(v[3]))) +
               ((softplus_fast(v[4]) + softplus_fast(v[5])) + (softplus_fast(v[6]) + softplus_fast(v[7])));
        if (++k == 8) total += (double)acc, acc = 0.f, k = 0;  // bound the fp32 partial
    }
    for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += softplus_fast(load_elem(x, i, DT));
    total += (double)acc;
    block_add(total, sums + 1, s_red);
}

template <int DT>
__global__ void __launch_bounds__(256) sigmoid_grad_half_kernel(const void* __restrict__ x, size_t n, size_t n8,
                                                                const float* __restrict__ tss, const float* __restrict__ gout,
                                                                void* __restrict__ gx) {
    const float scale = gout[1] / fmaxf(*tss, 1.f);
    const uint4* __restrict__ x8 = reinterpret_cast<const uint4*>(x);
    uint4* __restrict__ g8 = reinterpret_cast<uint4*>(gx);
    auto sg = [&](float v) { return __fdividef(scale, 1.f + __expf(-v)); };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float v[8];
        unpack8<DT>(__ldcs(x8 + i), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = sg(v[k]);
        __stcs(g8 + i, pack8<DT>(v));
    }
    for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        store_elem(gx, i, DT, sg(load_elem(x, i, DT)));
}

__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                      size_t n, const float* __restrict__ tss,
                                                      const float* __restrict__ gout, float* __restrict__ gx) {
    const float scale = gout[1] / fmaxf(*tss, 1.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float xv = __ldcs(x + i);
        const float sg = 1.f / (1.f + expf(-xv));
        __stcs(gx + i, (sg - __ldcs(t + i)) * scale);
    }
}

// ---- foreground anchors: four lanes per anchor, lane q owns side q (its 16 DFL bins in registers) ------
//
// A CTA takes a span of FG_SPAN consecutive anchors: it reads their fg_mask bytes, compacts the foreground
// ones (a few percent of the anchors) into a shared-memory list and works the list off 64 anchors at a time
// - a quad of lanes per anchor, so a warp has eight anchors in flight and nothing is serial but the
// arithmetic of one side. (The first version gave a whole warp to one anchor, two bins per lane, and walked
// a warp's foreground anchors one after the other: 125 us forward / 180 us backward at b = 16, A = 136,000
// with ~79,000 foreground anchors, against ~10 us of memory time.)

constexpr int FG_THREADS = 256;
constexpr int FG_SPAN = 2048;  // anchors per CTA

struct FgTerms {
    float box;      // (1 - ciou) * w                                        (same in the four lanes)
    float dfl;      // mean over sides of the DFL cross entropies * w        (same in the four lanes)
    float gd;       // this lane's side: d(box term)/d(distance of the side)
    float inv_sum;  // 1 / sum_k exp(x_k - m) of this lane's side
    float m;        // max logit of this lane's side
    float d;        // expectation of this lane's side
    float wl, wr;   // DFL weights of this lane's side
    int tl;         // left DFL bin of this lane's side
    float w;        // anchor weight
    int lab;        // sparse form: class index of the row's non-zero
};

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// All 32 lanes call this together (idle quads pass live = false and get zeros back). x[] = this lane's 16 logits.
template <bool GRAD>
__device__ __forceinline__ FgTerms fg_anchor(const LossParams& p, bool live, size_t idx, int a, int side, float (&x)[16]) {
    FgTerms r;
    r.box = r.dfl = r.gd = r.d = r.wl = r.wr = r.w = r.m = 0.f, r.inv_sum = 1.f, r.tl = 0, r.lab = 0;
    const size_t row0 = idx * 64 + side * 16;  // this lane's 16 bins inside pred_dist
    float ax = 0.f, ay = 0.f, st = 1.f;
    float4 tbp = make_float4(0.f, 0.f, 1.f, 1.f);
    float w = 0.f;
    if (live) {
        if (p.dist_dtype == 0) {
            const float4* __restrict__ row = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.pred_dist) + row0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 t = __ldg(row + k);
                x[4 * k] = t.x, x[4 * k + 1] = t.y, x[4 * k + 2] = t.z, x[4 * k + 3] = t.w;
            }
        } else {  // 16 half-precision bins = two 16-byte words
            const uint4* __restrict__ row = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.pred_dist) + row0);
            float lo[8], hi[8];
            if (p.dist_dtype == 1) unpack8<1>(__ldg(row), lo), unpack8<1>(__ldg(row + 1), hi);
            else unpack8<2>(__ldg(row), lo), unpack8<2>(__ldg(row + 1), hi);
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = lo[k], x[8 + k] = hi[k];
        }
        ax = p.anchor_points[2 * a], ay = p.anchor_points[2 * a + 1], st = p.stride[a];
        tbp = *reinterpret_cast<const float4*>(p.target_bboxes + idx * 4);
        // anchor weight: target_scores.sum(-1) (utils/loss.py:93) - in the sparse form the row's single non-zero
        if (p.target_labels) {
            long long lab = p.target_labels[idx];
            lab = lab < 0 ? 0 : (lab >= p.nc ? p.nc - 1 : lab);
            w = p.target_scores[idx * p.nc + lab];
            r.lab = (int)lab;
        } else {
            for (int c = side; c < p.nc; c += 4) w += p.target_scores[idx * p.nc + c];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = 0.f;
    }
    if (!p.target_labels) w = quad_sum(w);
    r.w = w;
    float m = x[0];
#pragma unroll
    for (int k = 1; k < 16; ++k) m = fmaxf(m, x[k]);
    float sum = 0.f, wsum = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float e = expf(x[k] - m);
        sum += e;
        wsum += (float)k * e;
    }
    const float d = wsum / sum;
    const float lse = m + logf(sum);
    r.m = m, r.inv_sum = 1.f / sum, r.d = d;
    // boxes in grid units
    const float tx1 = tbp.x / st, ty1 = tbp.y / st, tx2 = tbp.z / st, ty2 = tbp.w / st;  // utils/loss.py:251
    const float dl = __shfl_sync(0xffffffffu, d, 0, 4), dt = __shfl_sync(0xffffffffu, d, 1, 4);
    const float dr = __shfl_sync(0xffffffffu, d, 2, 4), db = __shfl_sync(0xffffffffu, d, 3, 4);
    const float x1 = ax - dl, y1 = ay - dt, x2 = ax + dr, y2 = ay + db;
    // CIoU(box1 = prediction, box2 = target)
    const float eps = 1e-7f;
    const float w1 = x2 - x1, h1 = y2 - y1 + eps, w2 = tx2 - tx1, h2 = ty2 - ty1 + eps;
    const float iw_raw = fminf(x2, tx2) - fmaxf(x1, tx1), ih_raw = fminf(y2, ty2) - fmaxf(y1, ty1);
    const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
    const float inter = iw * ih;
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(x2, tx2) - fminf(x1, tx1), ch = fmaxf(y2, ty2) - fminf(y1, ty1);
    const float c2 = cw * cw + ch * ch + eps;
    const float dx = tx1 + tx2 - x1 - x2, dy = ty1 + ty2 - y1 - y2;
    const float rho2 = (dx * dx + dy * dy) * 0.25f;
    const float at = atanf(w2 / h2) - atanf(w1 / h1);
    const float kv = 0.40528473456935109f;
    const float v = kv * at * at;
    const float alpha = v / (v - iou + (1.f + eps));
    const float ciou = iou - (rho2 / c2 + v * alpha);
    r.box = (1.f - ciou) * w;
    // DFL targets of this lane's side (bbox2dist, utils/tal.py:360-363; DFLoss utils/loss.py:80-88)
    const float tcoord = side == 0 ? ax - tx1 : side == 1 ? ay - ty1 : side == 2 ? tx2 - ax : ty2 - ay;
    const float tgt = fminf(fmaxf(tcoord, 0.f), 14.99f);
    const int tl = (int)tgt;
    const float wl = (float)(tl + 1) - tgt, wr = 1.f - wl;
    r.tl = tl, r.wl = wl, r.wr = wr;
    float x_tl = x[0], x_tr = x[1];  // logits of bins tl and tl + 1 (tl <= 14)
#pragma unroll
    for (int k = 1; k < 15; ++k)
        if (k == tl) x_tl = x[k], x_tr = x[k + 1];
    const float ce = (lse - x_tl) * wl + (lse - x_tr) * wr;
    r.dfl = quad_sum(ce) * 0.25f * w;
    if (GRAD) {
        // d ciou / d (x1, y1, x2, y2), alpha held constant (utils/metrics.py:127-128)
        const float ox = iw_raw > 0.f ? 1.f : 0.f, oy = ih_raw > 0.f ? 1.f : 0.f;
        const float diw_x1 = (x1 > tx1) ? -ox : 0.f, diw_x2 = (x2 < tx2) ? ox : 0.f;
        const float dih_y1 = (y1 > ty1) ? -oy : 0.f, dih_y2 = (y2 < ty2) ? oy : 0.f;
        const float dcw_x1 = (x1 < tx1) ? -1.f : 0.f, dcw_x2 = (x2 > tx2) ? 1.f : 0.f;
        const float dch_y1 = (y1 < ty1) ? -1.f : 0.f, dch_y2 = (y2 > ty2) ? 1.f : 0.f;
        const float u = w1 / h1;
        const float dv_du = -2.f * kv * at / (1.f + u * u);
        // per coordinate: inter', (w1 h1)', c2', rho2', u'
        float din, dwh, dc2, drho, du;
        if (side == 0)      din = ih * diw_x1, dwh = -h1, dc2 = 2.f * cw * dcw_x1, drho = -0.5f * dx, du = -1.f / h1;
        else if (side == 1) din = iw * dih_y1, dwh = -w1, dc2 = 2.f * ch * dch_y1, drho = -0.5f * dy, du = w1 / (h1 * h1);
        else if (side == 2) din = ih * diw_x2, dwh = h1, dc2 = 2.f * cw * dcw_x2, drho = -0.5f * dx, du = 1.f / h1;
        else                din = iw * dih_y2, dwh = w1, dc2 = 2.f * ch * dch_y2, drho = -0.5f * dy, du = -w1 / (h1 * h1);
        const float duni = dwh - din;
        const float diou = (din * uni - inter * duni) / (uni * uni);
        const float dciou = diou - (drho * c2 - rho2 * dc2) / (c2 * c2) - alpha * dv_du * du;
        // coordinate = anchor -/+ distance of the side
        const float dcoord_dd = side < 2 ? -1.f : 1.f;
        r.gd = -dciou * dcoord_dd * w;  // d[(1-ciou) w] / d(distance)
    }
    if (!live) r.box = r.dfl = r.gd = 0.f;
    return r;
}

// The span's foreground anchors -> s_list (index inside the span, ascending within a warp's 8-byte words); returns
// their number. flags (optional): one byte per anchor of the span, for the backward's zero fill.
__device__ __forceinline__ int fg_compact(const LossParams& p, long long first, int span, uint16_t* s_list, int* s_n,
                                          uint8_t* flags) {
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    const bool wide = (first % 8 == 0) && (((uintptr_t)p.fg_mask) % 8 == 0);
    for (int i0 = threadIdx.x * 8; i0 < span; i0 += FG_THREADS * 8) {  // whole warps iterate together (span is CTA-uniform)
        unsigned long long bits = 0ull;
        if (wide && i0 + 8 <= span) {
            bits = *reinterpret_cast<const unsigned long long*>(p.fg_mask + first + i0);
        } else {
            for (int k = 0; k < 8 && i0 + k < span; ++k) bits |= (unsigned long long)p.fg_mask[first + i0 + k] << (8 * k);
        }
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool f = ((bits >> (8 * k)) & 0xffull) != 0ull;
            cnt += f ? 1 : 0;
            if (flags && i0 + k < span) flags[i0 + k] = f ? 1 : 0;
        }
        if (cnt) {
            int pos = atomicAdd(s_n, cnt);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (((bits >> (8 * k)) & 0xffull) != 0ull) s_list[pos++] = (uint16_t)(i0 + k);
        }
    }
    __syncthreads();
    return *s_n;
}

__global__ void __launch_bounds__(FG_THREADS) fg_fwd_kernel(const LossParams p, double* __restrict__ sums) {
    __shared__ double s_red[8];
    __shared__ uint16_t s_list[FG_SPAN];
    __shared__ int s_n;
    const long long n = (long long)p.b * p.A;
    const long long first = (long long)blockIdx.x * FG_SPAN;
    const int span = (int)min((long long)FG_SPAN, n - first);
    const int n_fg = fg_compact(p, first, span, s_list, &s_n, nullptr);
    const int quad = threadIdx.x >> 2, side = threadIdx.x & 3;
    double box = 0.0, dfl = 0.0, cls = 0.0;
    for (int e0 = 0; e0 < n_fg; e0 += FG_THREADS / 4) {  // CTA-uniform trip count
        const int e = e0 + quad;
        const bool live = e < n_fg;
        const size_t idx = (size_t)(first + (live ? s_list[e] : 0));
        float x[16];
        const FgTerms t = fg_anchor<false>(p, live, idx, (int)(idx % (size_t)p.A), side, x);
        if (live && side == 0) {
            box += (double)t.box, dfl += (double)t.dfl;
            // sparse-target BCE: BCE(x, t) - BCE(x, 0) = -t x at the row's only non-zero target
            if (p.target_labels)
                cls -= (double)(round_to(t.w, p.score_dtype) * load_elem(p.pred_scores, idx * p.nc + t.lab, p.logits_dtype));
        }
    }
    block_add(box, sums + 0, s_red);
    block_add(dfl, sums + 2, s_red);
    if (p.target_labels) block_add(cls, sums + 1, s_red);
}

// grad_pred_dist for every anchor of the span: zeros for the background rows (streamed, the bulk of the
// kernel's traffic: 256 B per anchor), decode + CIoU + DFL chain for the foreground rows.
__global__ void __launch_bounds__(FG_THREADS) fg_bwd_kernel(const LossParams p, const float* __restrict__ gout,
                                                            void* __restrict__ gdist, void* gscores) {
    __shared__ uint16_t s_list[FG_SPAN];
    __shared__ uint8_t s_flag[FG_SPAN];
    __shared__ int s_n;
    const long long n = (long long)p.b * p.A;
    const long long first = (long long)blockIdx.x * FG_SPAN;
    const int span = (int)min((long long)FG_SPAN, n - first);
    const int n_fg = fg_compact(p, first, span, s_list, &s_n, s_flag);
    // background rows: 16 float4 per row, consecutive threads write consecutive float4
    const int w_row = p.dist_dtype == 0 ? 16 : 8;  // 16-byte words per row of grad_pred_dist (64 fp32 / 64 half elements)
    uint4* __restrict__ g4 = reinterpret_cast<uint4*>(gdist) + (size_t)first * w_row;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < span * w_row; i += FG_THREADS)
        if (!s_flag[p.dist_dtype == 0 ? i >> 4 : i >> 3]) __stcs(g4 + i, z);
    const float inv_tss = 1.f / fmaxf(*p.tss, 1.f);
    const float g_box = gout[0] * inv_tss, g_dfl = gout[2] * inv_tss, g_cls = gout[1] * inv_tss;
    const int quad = threadIdx.x >> 2, side = threadIdx.x & 3;
    for (int e0 = 0; e0 < n_fg; e0 += FG_THREADS / 4) {  // CTA-uniform trip count
        const int e = e0 + quad;
        const bool live = e < n_fg;
        const size_t idx = (size_t)(first + (live ? s_list[e] : 0));
        float x[16];
        const FgTerms t = fg_anchor<true>(p, live, idx, (int)(idx % (size_t)p.A), side, x);
        if (!live) continue;  // (no collective below)
        const float gd = g_box * t.gd;          // through the expectation d = sum_k k p_k
        const float gc = g_dfl * 0.25f * t.w;   // through the two cross entropies
        float g[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float pk = expf(x[k] - t.m) * t.inv_sum;
            g[k] = gd * pk * ((float)k - t.d) + gc * (pk - (k == t.tl ? t.wl : 0.f) - (k == t.tl + 1 ? t.wr : 0.f));
        }
        if (p.dist_dtype == 0) {
            float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(gdist) + idx * 64 + side * 16);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) __stcs(dst + k4, make_float4(g[4 * k4], g[4 * k4 + 1], g[4 * k4 + 2], g[4 * k4 + 3]));
        } else {
            uint4* __restrict__ dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(gdist) + idx * 64 + side * 16);
            float lo[8], hi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) lo[k] = g[k], hi[k] = g[8 + k];
            if (p.dist_dtype == 1) __stcs(dst, pack8<1>(lo)), __stcs(dst + 1, pack8<1>(hi));
            else __stcs(dst, pack8<2>(lo)), __stcs(dst + 1, pack8<2>(hi));
        }
        // sparse-target BCE: the dense pass wrote sigmoid(x) * scale; the target's share goes in here
        if (gscores && side == 0) {
            const size_t gi = idx * p.nc + t.lab;
            const float share = round_to(t.w, p.score_dtype) * g_cls;
            if (p.logits_dtype == 0) {
                reinterpret_cast<float*>(gscores)[gi] -= share;
            } else {  // half-precision gradient: recomputed in fp32 and rounded once, not the stored rounded value minus the share
                const float xv = load_elem(p.pred_scores, gi, p.logits_dtype);
                store_elem(gscores, gi, p.logits_dtype, __fdividef(gout[1] / fmaxf(*p.tss, 1.f), 1.f + __expf(-xv)) - share);
            }
        }
    }
}

__global__ void loss_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ tss,
                                     float* __restrict__ loss) {
    const double d = fmax((double)*tss, 1.0);
    if (threadIdx.x < 3) loss[threadIdx.x] = (float)(sums[threadIdx.x] / d);
}

}  // namespace kzd

using namespace kzd;

static int check_loss_args(const LossParams& p, const char* who) {
    KZD_CHECK_ARG(p.pred_scores && p.pred_dist && p.anchor_points && p.stride && p.target_bboxes && p.target_scores &&
                      p.fg_mask && p.tss,
                  "%s: null pointer", who);
    KZD_CHECK_ARG(p.b >= 1 && p.A >= 1 && p.nc >= 1, "%s: bad sizes", who);
    KZD_CHECK_ARG(((uintptr_t)p.pred_dist) % 16 == 0 && ((uintptr_t)p.target_bboxes) % 16 == 0, "%s: misaligned tensor",
                  who);
    return 0;
}

static int loss_fwd(const LossParams& p, double* sums_ws, float* loss3, cudaStream_t st) {
    KZD_CUDA(cudaMemsetAsync(sums_ws, 0, 3 * sizeof(double), st));
    const size_t n = (size_t)p.b * p.A * p.nc;
    const int cap = sm_count() * 8;
    if (p.target_labels && p.logits_dtype != 0) {
        const size_t n8 = ((uintptr_t)p.pred_scores) % 16 == 0 ? n / 8 : 0;
        const size_t work = n8 ? n8 : n;
        const unsigned grid = (unsigned)((work + 255) / 256 < (size_t)cap ? (work + 255) / 256 : cap);
        if (p.logits_dtype == 1) softplus_sum_half_kernel<1><<<grid, 256, 0, st>>>(p.pred_scores, n, n8, sums_ws);
        else softplus_sum_half_kernel<2><<<grid, 256, 0, st>>>(p.pred_scores, n, n8, sums_ws);
        KZD_LAUNCH_CHECK("softplus_sum_half_kernel");
    } else if (p.target_labels) {
        const size_t n4 = ((uintptr_t)p.pred_scores) % 16 == 0 ? n / 4 : 0;
        const size_t work = n4 ? n4 : n;
        softplus_sum_kernel<<<(unsigned)((work + 255) / 256 < (size_t)cap ? (work + 255) / 256 : cap), 256, 0, st>>>(
            reinterpret_cast<const float*>(p.pred_scores), n, n4, sums_ws);
        KZD_LAUNCH_CHECK("softplus_sum_kernel");
    } else {
        bce_fwd_kernel<<<(unsigned)((n + 255) / 256 < (size_t)cap ? (n + 255) / 256 : cap), 256, 0, st>>>(
            reinterpret_cast<const float*>(p.pred_scores), p.target_scores, n, sums_ws);
        KZD_LAUNCH_CHECK("bce_fwd_kernel");
    }
    const long long spans = ((long long)p.b * p.A + FG_SPAN - 1) / FG_SPAN;
    fg_fwd_kernel<<<(unsigned)spans, FG_THREADS, 0, st>>>(p, sums_ws);
    KZD_LAUNCH_CHECK("fg_fwd_kernel");
    loss_finalize_kernel<<<1, 32, 0, st>>>(sums_ws, p.tss, loss3);
    KZD_LAUNCH_CHECK("loss_finalize_kernel");
    return 0;
}

static int loss_bwd(const LossParams& p, const float* grad_loss3, void* grad_pred_scores, void* grad_pred_dist,
                    cudaStream_t st) {
    const size_t n = (size_t)p.b * p.A * p.nc;
    const int cap = sm_count() * 8;
    if (p.target_labels && p.logits_dtype != 0) {
        const size_t n8 = (((uintptr_t)p.pred_scores) | ((uintptr_t)grad_pred_scores)) % 16 == 0 ? n / 8 : 0;
        const size_t work = n8 ? n8 : n;
        const unsigned grid = (unsigned)((work + 255) / 256 < (size_t)cap ? (work + 255) / 256 : cap);
        if (p.logits_dtype == 1) sigmoid_grad_half_kernel<1><<<grid, 256, 0, st>>>(p.pred_scores, n, n8, p.tss, grad_loss3, grad_pred_scores);
        else sigmoid_grad_half_kernel<2><<<grid, 256, 0, st>>>(p.pred_scores, n, n8, p.tss, grad_loss3, grad_pred_scores);
        KZD_LAUNCH_CHECK("sigmoid_grad_half_kernel");
    } else if (p.target_labels) {
        const size_t n4 = (((uintptr_t)p.pred_scores) | ((uintptr_t)grad_pred_scores)) % 16 == 0 ? n / 4 : 0;
        const size_t work = n4 ? n4 : n;
        sigmoid_grad_kernel<<<(unsigned)((work + 255) / 256 < (size_t)cap ? (work + 255) / 256 : cap), 256, 0, st>>>(
            reinterpret_cast<const float*>(p.pred_scores), n, n4, p.tss, grad_loss3, reinterpret_cast<float*>(grad_pred_scores));
        KZD_LAUNCH_CHECK("sigmoid_grad_kernel");
    } else {
        bce_bwd_kernel<<<(unsigned)((n + 255) / 256 < (size_t)cap ? (n + 255) / 256 : cap), 256, 0, st>>>(
            reinterpret_cast<const float*>(p.pred_scores), p.target_scores, n, p.tss, grad_loss3,
            reinterpret_cast<float*>(grad_pred_scores));
        KZD_LAUNCH_CHECK("bce_bwd_kernel");
    }
    const long long spans = ((long long)p.b * p.A + FG_SPAN - 1) / FG_SPAN;
    fg_bwd_kernel<<<(unsigned)spans, FG_THREADS, 0, st>>>(p, grad_loss3, grad_pred_dist, p.target_labels ? grad_pred_scores : nullptr);
    KZD_LAUNCH_CHECK("fg_bwd_kernel");
    return 0;
}

extern "C" __attribute__((visibility("default"))) int kzd_loss_tail_fwd(
    const float* pred_scores, const float* pred_dist, const float* anchor_points, const float* stride_tensor,
    const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask, const float* target_scores_sum,
    int b, int A, int nc, double* sums_ws, float* loss3, void* stream) {
    LossParams p{pred_scores, pred_dist, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
                 target_scores_sum, b, A, nc, nullptr, 0, 0, 0};
    if (int rc = check_loss_args(p, "kzd_loss_tail_fwd")) return rc;
    KZD_CHECK_ARG(sums_ws && loss3, "kzd_loss_tail_fwd: null output");
    return loss_fwd(p, sums_ws, loss3, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int kzd_loss_tail_bwd(
    const float* pred_scores, const float* pred_dist, const float* anchor_points, const float* stride_tensor,
    const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask, const float* target_scores_sum,
    int b, int A, int nc, const float* grad_loss3, float* grad_pred_scores, float* grad_pred_dist, void* stream) {
    LossParams p{pred_scores, pred_dist, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
                 target_scores_sum, b, A, nc, nullptr, 0, 0, 0};
    if (int rc = check_loss_args(p, "kzd_loss_tail_bwd")) return rc;
    KZD_CHECK_ARG(grad_loss3 && grad_pred_scores && grad_pred_dist, "kzd_loss_tail_bwd: null pointer");
    KZD_CHECK_ARG(((uintptr_t)grad_pred_dist) % 16 == 0, "kzd_loss_tail_bwd: misaligned grad_pred_dist");
    return loss_bwd(p, grad_loss3, grad_pred_scores, grad_pred_dist, (cudaStream_t)stream);
}

static int check_native(int logits_dtype, int dist_dtype, const char* who) {
    KZD_CHECK_ARG(logits_dtype >= 0 && logits_dtype <= 2 && dist_dtype >= 0 && dist_dtype <= 2,
                  "%s: logits_dtype=%d / dist_dtype=%d not in {0 fp32, 1 fp16, 2 bf16}", who, logits_dtype, dist_dtype);
    return 0;
}

extern "C" __attribute__((visibility("default"))) int kzd_loss_tail_sparse_fwd(
    const void* pred_scores, const void* pred_dist, const float* anchor_points, const float* stride_tensor,
    const float* target_bboxes, const float* target_scores, const int64_t* target_labels, const uint8_t* fg_mask,
    const float* target_scores_sum, int b, int A, int nc, int score_dtype, int logits_dtype, int dist_dtype,
    double* sums_ws, float* loss3, void* stream) {
    if (int rc = check_native(logits_dtype, dist_dtype, "kzd_loss_tail_sparse_fwd")) return rc;
    LossParams p{pred_scores, pred_dist, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
                 target_scores_sum, b, A, nc, (const long long*)target_labels, score_dtype, logits_dtype, dist_dtype};
    if (int rc = check_loss_args(p, "kzd_loss_tail_sparse_fwd")) return rc;
    KZD_CHECK_ARG(target_labels && sums_ws && loss3, "kzd_loss_tail_sparse_fwd: null pointer");
    KZD_CHECK_ARG(score_dtype >= 0 && score_dtype <= 2, "kzd_loss_tail_sparse_fwd: score_dtype=%d not in {0,1,2}", score_dtype);
    return loss_fwd(p, sums_ws, loss3, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int kzd_loss_tail_sparse_bwd(
    const void* pred_scores, const void* pred_dist, const float* anchor_points, const float* stride_tensor,
    const float* target_bboxes, const float* target_scores, const int64_t* target_labels, const uint8_t* fg_mask,
    const float* target_scores_sum, int b, int A, int nc, int score_dtype, int logits_dtype, int dist_dtype,
    const float* grad_loss3, void* grad_pred_scores, void* grad_pred_dist, void* stream) {
    if (int rc = check_native(logits_dtype, dist_dtype, "kzd_loss_tail_sparse_bwd")) return rc;
    LossParams p{pred_scores, pred_dist, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
                 target_scores_sum, b, A, nc, (const long long*)target_labels, score_dtype, logits_dtype, dist_dtype};
    if (int rc = check_loss_args(p, "kzd_loss_tail_sparse_bwd")) return rc;
    KZD_CHECK_ARG(target_labels && grad_loss3 && grad_pred_scores && grad_pred_dist, "kzd_loss_tail_sparse_bwd: null pointer");
    KZD_CHECK_ARG(((uintptr_t)grad_pred_dist) % 16 == 0, "kzd_loss_tail_sparse_bwd: misaligned grad_pred_dist");
    KZD_CHECK_ARG(score_dtype >= 0 && score_dtype <= 2, "kzd_loss_tail_sparse_bwd: score_dtype=%d not in {0,1,2}", score_dtype);
    return loss_bwd(p, grad_loss3, grad_pred_scores, grad_pred_dist, (cudaStream_t)stream);
}
