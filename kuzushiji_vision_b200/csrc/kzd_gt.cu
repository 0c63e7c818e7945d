// Ground-truth packing (sm_100a): v8DetectionLoss.preprocess without the per-image Python loop.
//
// reference: utils/loss.py:180-195 - targets (n, 6) = [image index, class, cx, cy, w, h] (normalised)
// -> out (b, G, 5) = [class, x1, y1, x2, y2] in pixels, G = the largest per-image count, zero rows as
// padding; rows keep their order of appearance inside an image (out[j, :n] = targets[i == j, 1:]).
// The reference syncs the host once per image (matches.sum()); here: one count kernel, one read-back of
// (counts, first, sortedness) to size the output, one pack kernel.
#include "kzd_common.cuh"

namespace kzd {

// counts[j] = number of targets of image j; first[j] = smallest row index of image j;
// flags[0] != 0 when the image indices are not non-decreasing (the pack kernel needs them grouped).
__global__ void gt_count_kernel(const float* __restrict__ targets, int n, int b, int32_t* __restrict__ counts,
                                int32_t* __restrict__ first, int32_t* __restrict__ flags) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // whole warps stay alive for the match
    int j = -1;
    if (t < n) {
        j = (int)targets[(size_t)t * 6];
        if (j < 0 || j >= b) {
            atomicOr(flags, 2);
            j = -1;
        } else if (t > 0 && targets[(size_t)(t - 1) * 6] > targets[(size_t)t * 6]) {
            atomicOr(flags, 1);
        }
    }
    // rows of one image sit next to each other, so a warp holds one or two images: the lowest lane of each takes
    // the image's two atomics for all of them (hundreds of rows per image would otherwise queue on one address)
    const unsigned peers = __match_any_sync(0xffffffffu, j);
    if (j >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        atomicAdd(counts + j, __popc(peers));
        atomicMin(first + j, t);  // the lowest lane of the group holds its smallest row index
    }
}

// One thread per output row (image j, slot s): row first[j]+s of the (grouped) targets, or zeros.
__global__ void gt_pack_kernel(const float* __restrict__ targets, int b, int G, const float* __restrict__ scale,
                               const int32_t* __restrict__ counts, const int32_t* __restrict__ first,
                               float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * G) return;
    const int j = idx / G, s = idx - j * G;
    float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (s < counts[j]) {
        const float* r = targets + (size_t)(first[j] + s) * 6;
        // out[..., 1:5].mul_(scale_tensor) then xywh2xyxy (utils/loss.py:194, utils/ops.py:429-432)
        const float cx = __fmul_rn(r[2], scale[0]), cy = __fmul_rn(r[3], scale[1]);
        const float hw = __fmul_rn(__fmul_rn(r[4], scale[2]), 0.5f), hh = __fmul_rn(__fmul_rn(r[5], scale[3]), 0.5f);
        o[0] = r[1];
        o[1] = __fsub_rn(cx, hw), o[2] = __fsub_rn(cy, hh), o[3] = __fadd_rn(cx, hw), o[4] = __fadd_rn(cy, hh);
    }
    float* d = out + (size_t)idx * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) d[k] = o[k];
}

}  // namespace kzd

using namespace kzd;

extern "C" __attribute__((visibility("default"))) int kzd_gt_count(const float* targets, int n, int b,
                                                                     int32_t* counts, int32_t* first,
                                                                     int32_t* flags, void* stream) {
    KZD_CHECK_ARG(targets && counts && first && flags, "kzd_gt_count: null pointer");
    KZD_CHECK_ARG(n >= 0 && b >= 1, "kzd_gt_count: bad sizes n=%d b=%d", n, b);
    cudaStream_t st = (cudaStream_t)stream;
    KZD_CUDA(cudaMemsetAsync(counts, 0, (size_t)b * 4, st));
    KZD_CUDA(cudaMemsetAsync(first, 0x7f, (size_t)b * 4, st));
    KZD_CUDA(cudaMemsetAsync(flags, 0, 4, st));
    if (n == 0) return 0;
    gt_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(targets, n, b, counts, first, flags);
    KZD_LAUNCH_CHECK("gt_count_kernel");
    return 0;
}

extern "C" __attribute__((visibility("default"))) int kzd_gt_pack(const float* targets, int n, int b, int G,
                                                                    const float* scale4, const int32_t* counts,
                                                                    const int32_t* first, float* out, void* stream) {
    KZD_CHECK_ARG(targets && scale4 && counts && first && out, "kzd_gt_pack: null pointer");
    KZD_CHECK_ARG(n >= 0 && b >= 1 && G >= 0, "kzd_gt_pack: bad sizes");
    if (G == 0) return 0;
    const long long rows = (long long)b * G;
    KZD_CHECK_ARG(rows < 0x7fffffffLL, "kzd_gt_pack: b*G too large");
    gt_pack_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(targets, b, G, scale4, counts,
                                                                                   first, out);
    KZD_LAUNCH_CHECK("gt_pack_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// post-NMS tail: scale_boxes + clip_boxes for a whole batch of detections
// reference: utils/ops.py:92-127 (scale_boxes, padding=True, xywh=False) and :319-338 (clip_boxes),
// called per image from models/yolo/detect/predict.py:37-41.
//   x -= pad_x; y -= pad_y;  then  *= 1/gain  (aten divides by a Python scalar as a * (1.0f/b));
//   then clamp x to [0, w0], y to [0, h0].
// params (b, 5) = [pad_x, pad_y, gain, w0, h0]; rows (b, max_det, stride) with the box in columns 0..3.
// ------------------------------------------------------------------------------------------------

namespace kzd {

__global__ void scale_boxes_kernel(float* __restrict__ rows, const int32_t* __restrict__ counts, int b, int max_det,
                                   int stride, const float* __restrict__ params) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * max_det) return;
    const int bi = idx / max_det, k = idx - bi * max_det;
    if (counts && k >= counts[bi]) return;
    const float* p = params + bi * 5;
    // aten divides a tensor by a host scalar as a multiplication with the scalar's reciprocal, taken in double and
    // then rounded to fp32 (BinaryDivTrueKernel.cu, is_cpu_scalar path): the host passes that reciprocal
    const float inv = p[2];
    float* r = rows + (size_t)idx * stride;
    const float x1 = __fmul_rn(__fsub_rn(r[0], p[0]), inv), y1 = __fmul_rn(__fsub_rn(r[1], p[1]), inv);
    const float x2 = __fmul_rn(__fsub_rn(r[2], p[0]), inv), y2 = __fmul_rn(__fsub_rn(r[3], p[1]), inv);
    r[0] = fminf(fmaxf(x1, 0.f), p[3]), r[1] = fminf(fmaxf(y1, 0.f), p[4]);
    r[2] = fminf(fmaxf(x2, 0.f), p[3]), r[3] = fminf(fmaxf(y2, 0.f), p[4]);
}

}  // namespace kzd

extern "C" __attribute__((visibility("default"))) int kzd_scale_boxes(float* rows, const int32_t* counts, int b,
                                                                        int max_det, int row_stride,
                                                                        const float* params, void* stream) {
    KZD_CHECK_ARG(rows && params, "kzd_scale_boxes: null pointer");
    KZD_CHECK_ARG(b >= 0 && max_det >= 0 && row_stride >= 4, "kzd_scale_boxes: bad sizes");
    const long long n = (long long)b * max_det;
    if (n == 0) return 0;
    KZD_CHECK_ARG(n < 0x7fffffffLL, "kzd_scale_boxes: b*max_det too large");
    kzd::scale_boxes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, counts, b, max_det,
                                                                                          row_stride, params);
    KZD_LAUNCH_CHECK("scale_boxes_kernel");
    return 0;
}
