// Validator matching (sm_100a): which kept detections count as true positives at each IoU threshold, for all
// images of a batch at once.
//
// reference: DetectionValidator._process_batch models/yolo/detect/val.py:209-228 (box_iou of labels x detections,
// utils/metrics.py:52-71) and BaseValidator.match_predictions engine/validator.py:222-260 (use_scipy=False):
// per threshold, every (label, detection) pair of equal class with IoU >= threshold is a match; matches are
// ordered by IoU descending, each detection keeps its first match (its best label), then - in ascending
// detection order - each label keeps its first detection. A detection is correct at the threshold when it
// survives both steps.
//
// Restated without the sort: a detection's best label is the argmax of its class-masked IoU column whatever
// the threshold, so one pass finds (best IoU, best label) per detection; per (threshold, label) the surviving
// detection is the smallest detection index among those that chose the label and pass the threshold - one
// atomicMin. Kernels: match_best_kernel (labels staged through shared memory, one thread per detection),
// match_flag_kernel (one thread per detection x threshold).
#include "kzd_common.cuh"

namespace kzd {

constexpr int MATCH_THREADS = 128;
constexpr int MATCH_TILE = 256;  // labels staged per round
constexpr int MATCH_MAX_IOU = 32;

struct MatchParams {
    const float* det;        // (b, max_det, row_stride)
    const int32_t* det_cnt;  // (b) or nullptr
    const float* gt;         // (b, G, 5) [cls, x1, y1, x2, y2]
    const int32_t* gt_cnt;   // (b)
    const float* iouv;       // (n_iou)
    int b, max_det, row_stride, G, n_iou;
    uint32_t* best_bits;     // (b, max_det) thresholds the detection's best IoU passes (bit i)
    int32_t* best_label;     // (b, max_det)
    int32_t* first;          // (b, n_iou, G) smallest detection index per (threshold, label); 0x7f7f7f7f = none
};

// utils/metrics.py:67-71, box1 = label, box2 = detection; every operation rounds separately (eager torch)
__device__ __forceinline__ float pair_iou(const float4 a, const float4 d) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, d.z), fmaxf(a.x, d.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, d.w), fmaxf(a.y, d.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_d = __fmul_rn(__fsub_rn(d.z, d.x), __fsub_rn(d.w, d.y));
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_d), inter), 1e-7f));
}

__global__ void __launch_bounds__(MATCH_THREADS) match_best_kernel(const MatchParams p) {
    __shared__ float4 s_box[MATCH_TILE];
    __shared__ float s_cls[MATCH_TILE];
    __shared__ float s_thr[MATCH_MAX_IOU];
    const int bi = blockIdx.y;
    const int d = blockIdx.x * MATCH_THREADS + threadIdx.x;
    const int n_det = p.det_cnt ? min(max(p.det_cnt[bi], 0), p.max_det) : p.max_det;
    const int n_gt = min(max(p.gt_cnt[bi], 0), p.G);
    if (threadIdx.x < p.n_iou) s_thr[threadIdx.x] = p.iouv[threadIdx.x];
    const bool live = d < n_det;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    float cls = -1.f;
    if (live) {
        const float* r = p.det + ((size_t)bi * p.max_det + d) * p.row_stride;
        box = make_float4(r[0], r[1], r[2], r[3]);
        cls = r[5];
    }
    float best = 0.f;
    int best_l = -1;
    for (int l0 = 0; l0 < n_gt; l0 += MATCH_TILE) {
        __syncthreads();
        for (int t = threadIdx.x; t < MATCH_TILE && l0 + t < n_gt; t += MATCH_THREADS) {
            const float* g = p.gt + ((size_t)bi * p.G + l0 + t) * 5;
            s_cls[t] = g[0];
            s_box[t] = make_float4(g[1], g[2], g[3], g[4]);
        }
        __syncthreads();
        const int n_t = min(MATCH_TILE, n_gt - l0);
        if (live)
            for (int t = 0; t < n_t; ++t) {
                if (s_cls[t] != cls) continue;  // iou * correct_class (engine/validator.py:239-240)
                const float v = pair_iou(s_box[t], box);
                // the best label of a detection; equal IoUs: the reference's unstable argsort()[::-1] leaves the
                // order open - the later label wins here, as it does for numpy's small-array (stable) sort
                if (v >= best && v > 0.f) best = v, best_l = l0 + t;
            }
    }
    if (!live) return;
    uint32_t bits = 0u;
    if (best_l >= 0)
        for (int i = 0; i < p.n_iou; ++i)
            if (best >= s_thr[i]) bits |= 1u << i;
    p.best_bits[(size_t)bi * p.max_det + d] = bits;
    p.best_label[(size_t)bi * p.max_det + d] = best_l;
    for (int i = 0; i < p.n_iou; ++i)
        if ((bits >> i) & 1u) atomicMin(p.first + ((size_t)bi * p.n_iou + i) * p.G + best_l, d);
}

__global__ void __launch_bounds__(256) match_flag_kernel(const MatchParams p, uint8_t* __restrict__ correct) {
    const long long n = (long long)p.b * p.max_det * p.n_iou;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % p.n_iou);
        const long long row = t / p.n_iou;
        const int d = (int)(row % p.max_det), bi = (int)(row / p.max_det);
        const int n_det = p.det_cnt ? min(max(p.det_cnt[bi], 0), p.max_det) : p.max_det;
        uint8_t ok = 0;
        if (d < n_det && ((p.best_bits[row] >> i) & 1u))
            ok = p.first[((size_t)bi * p.n_iou + i) * p.G + p.best_label[row]] == d;
        correct[t] = ok;
    }
}

static void split_match_ws(void* workspace, int b, int max_det, int G, int n_iou, MatchParams& p, size_t& first_off,
                           size_t& first_bytes, size_t& total) {
    unsigned char* w = (unsigned char*)workspace;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char* r = w ? w + o : nullptr;
        o += align_up(bytes, 256);
        return r;
    };
    p.best_bits = (uint32_t*)take((size_t)b * max_det * 4);
    p.best_label = (int32_t*)take((size_t)b * max_det * 4);
    first_off = o;
    first_bytes = (size_t)b * n_iou * (G > 0 ? G : 1) * 4;
    p.first = (int32_t*)take(first_bytes);
    total = o;
}

}  // namespace kzd

using namespace kzd;

extern "C" __attribute__((visibility("default"))) size_t kzd_match_workspace_bytes(int b, int max_det, int G, int n_iou) {
    if (b < 0 || max_det < 0 || G < 0 || n_iou < 0) return 0;
    MatchParams p;
    size_t off = 0, fb = 0, total = 0;
    split_match_ws(nullptr, b, max_det, G, n_iou, p, off, fb, total);
    return total;
}

extern "C" __attribute__((visibility("default"))) int kzd_match_predictions(
    const float* det_rows, const int32_t* det_counts, int b, int max_det, int row_stride, const float* gt,
    const int32_t* gt_counts, int G, const float* iouv, int n_iou, uint8_t* correct, void* workspace,
    size_t workspace_bytes, void* stream) {
    KZD_CHECK_ARG(det_rows && gt_counts && iouv && correct && workspace, "kzd_match_predictions: null pointer");
    KZD_CHECK_ARG(gt || G == 0, "kzd_match_predictions: null labels with G=%d", G);
    KZD_CHECK_ARG(b >= 0 && max_det >= 0 && G >= 0 && row_stride >= 6, "kzd_match_predictions: bad sizes");
    KZD_CHECK_ARG(n_iou >= 1 && n_iou <= MATCH_MAX_IOU, "kzd_match_predictions: n_iou=%d not in [1,%d]", n_iou, MATCH_MAX_IOU);
    KZD_CHECK_ARG(b <= 65535, "kzd_match_predictions: b=%d > 65535", b);
    KZD_CHECK_ARG(((uintptr_t)workspace) % 256 == 0, "kzd_match_predictions: workspace must be 256-byte aligned");
    if (b == 0 || max_det == 0) return 0;
    MatchParams p;
    size_t first_off = 0, first_bytes = 0, need = 0;
    split_match_ws(workspace, b, max_det, G, n_iou, p, first_off, first_bytes, need);
    if (workspace_bytes < need) {
        set_error("kzd_match_predictions: workspace %zu B < required %zu B", workspace_bytes, need);
        return KZD_ENOSPACE;
    }
    p.det = det_rows, p.det_cnt = det_counts, p.gt = gt, p.gt_cnt = gt_counts, p.iouv = iouv;
    p.b = b, p.max_det = max_det, p.row_stride = row_stride, p.G = G, p.n_iou = n_iou;
    cudaStream_t st = (cudaStream_t)stream;
    KZD_CUDA(cudaMemsetAsync((unsigned char*)workspace + first_off, 0x7f, first_bytes, st));  // 0x7f7f7f7f: no detection yet
    match_best_kernel<<<dim3((max_det + MATCH_THREADS - 1) / MATCH_THREADS, b), MATCH_THREADS, 0, st>>>(p);
    KZD_LAUNCH_CHECK("match_best_kernel");
    const long long n = (long long)b * max_det * n_iou;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    match_flag_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, correct);
    KZD_LAUNCH_CHECK("match_flag_kernel");
    return 0;
}
