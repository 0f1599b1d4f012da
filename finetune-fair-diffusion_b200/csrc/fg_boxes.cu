// Largest-face selection + box expansion, batched (one thread per image).
//   get_largest_face_app  E1:1292-1304      expand_bbox  E1:238-265
// The arithmetic follows the reference under its pinned NumPy 1.26.4 scalar promotion
// (float32 scalar op float32 scalar -> float32; anything touching a Python number -> float64);
// see oracle/boxes.py.  All operations are single IEEE ops with explicit rounding intrinsics so
// nvcc cannot contract them into FMAs: the corners must round to the same integers.
#include "fg_common.cuh"

namespace {

struct Mixed {           // a value that is either a numpy float32 scalar or already float64
    double v;
    bool f32;
};

__device__ __forceinline__ Mixed clip_hi(float v, int dim_max) {
    // python: min(v, dim_max) keeps v unless dim_max < v
    return ((double)dim_max < (double)v) ? Mixed{(double)dim_max, false} : Mixed{(double)v, true};
}
__device__ __forceinline__ Mixed clip_lo(float v, int dim_min) {
    return ((double)dim_min > (double)v) ? Mixed{(double)dim_min, false} : Mixed{(double)v, true};
}
__device__ __forceinline__ Mixed msub(Mixed a, Mixed b) {
    if (a.f32 && b.f32) return Mixed{(double)__fsub_rn((float)a.v, (float)b.v), true};
    return Mixed{__dsub_rn(a.v, b.v), false};
}
__device__ __forceinline__ Mixed mmul(Mixed a, Mixed b) {
    if (a.f32 && b.f32) return Mixed{(double)__fmul_rn((float)a.v, (float)b.v), true};
    return Mixed{__dmul_rn(a.v, b.v), false};
}

__global__ void select_expand_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ counts,
                                     int n, int max_faces, int dim_max, double expand_coef, double target_ratio,
                                     long long fill, long long* __restrict__ out, uint8_t* __restrict__ ind) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cnt = counts ? counts[i] : max_faces;
    cnt = cnt < 0 ? 0 : (cnt > max_faces ? max_faces : cnt);
    if (cnt == 0) {
        if (ind) ind[i] = 0;
        for (int q = 0; q < 4; q++) out[4 * i + q] = fill;
        return;
    }
    const float* bb = boxes + (size_t)i * max_faces * 4;
    int best = 0;
    if (cnt > 1) {
        double best_area = 0.0;
        for (int k = 0; k < cnt; k++) {
            const float* b = bb + 4 * k;
            Mixed w = msub(clip_hi(b[2], dim_max), clip_lo(b[0], 0));
            Mixed h = msub(clip_hi(b[3], dim_max), clip_lo(b[1], 0));
            double area = mmul(w, h).v;
            if (area > best_area) { best_area = area; best = k; }
        }
    }
    const float* b = bb + 4 * best;
    float w = __fsub_rn(b[2], b[0]);
    float h = __fsub_rn(b[3], b[1]);
    float ratio = __fdiv_rn(h, w);
    double extra_w, extra_h;
    if ((double)ratio > target_ratio) {
        extra_h = __dmul_rn((double)h, expand_coef);
        extra_w = __dsub_rn(__ddiv_rn(__dadd_rn((double)h, extra_h), target_ratio), (double)w);
    } else {
        extra_w = __dmul_rn((double)w, expand_coef);
        extra_h = __dsub_rn(__dmul_rn(__dadd_rn((double)w, extra_w), target_ratio), (double)h);
    }
    double hw = __dmul_rn(extra_w, 0.5), hh = __dmul_rn(extra_h, 0.5);
    out[4 * i + 0] = __double2ll_rn(__dsub_rn((double)b[0], hw));
    out[4 * i + 2] = __double2ll_rn(__dadd_rn((double)b[2], hw));
    out[4 * i + 1] = __double2ll_rn(__dsub_rn((double)b[1], hh));
    out[4 * i + 3] = __double2ll_rn(__dadd_rn((double)b[3], hh));
    if (ind) ind[i] = 1;
}

}  // namespace

extern "C" int fg_select_expand_boxes(const float* boxes, const int32_t* counts, int n, int max_faces, int dim_max,
                                      double expand_coef, double target_ratio, int64_t fill,
                                      int64_t* boxes_out, uint8_t* indicators_out, void* stream) {
    if (n < 0 || max_faces <= 0 || !boxes_out || (!boxes && n > 0)) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    int threads = 128;
    select_expand_kernel<<<(n + threads - 1) / threads, threads, 0, fg_stream(stream)>>>(
        boxes, counts, n, max_faces, dim_max, expand_coef, target_ratio, (long long)fill,
        reinterpret_cast<long long*>(boxes_out), indicators_out);
    FG_LAUNCH_CHECK();
    return FG_OK;
}
