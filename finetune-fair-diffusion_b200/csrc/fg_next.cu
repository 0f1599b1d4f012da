// The two "next" rows either side of the path:
//
//   f2  detector staging     E1:1317, 1326 (same line in E3 / E4)
//         images_np = ((images*0.5 + 0.5)*255).cpu().detach().permute(0,2,3,1).float().numpy().astype(np.uint8)
//         face_app.get(image_np[:,:,[2,1,0]])
//       The reference moves the FLOAT tensor to the host and converts there.  Here the conversion, the NCHW -> NHWC
//       permute and the RGB -> BGR swap run on the device, so the device-to-host copy is 3 bytes per pixel instead of
//       3 * sizeof(dtype) and the host does no arithmetic.  The three operations round through the image dtype one by
//       one exactly like the eager expression (fp16 in the reference), then truncate towards zero.
//
//   f3  bias-gap metrics     E3:1716-1749, E4:1780-1821 (get_evaluate_metrics)
//       class frequencies of the argmax predictions, share of low-confidence predictions, mean pairwise gaps; the
//       reference makes 5-9 blocking .item() calls, this is one launch writing all numbers to one small buffer.
#include "fg_common.cuh"

namespace {

// (x*0.5 + 0.5)*255 with one rounding to T after each operation, then numpy's float -> uint8 cast (truncate; values
// outside [0,256) wrap modulo 256 like the x86 conversion numpy uses; non-finite -> 0).
// The scalar float<->16-bit and float->int conversions (F2F / F2I) run on the quarter-rate pipe and made the first
// version of this kernel conversion-bound at 40 % of HBM speed, so two values are rounded at a time with the packed
// full-rate conversion (F2FP ... PACK_AB) and the truncation is a round-towards-zero add of 2^23.
template <typename T> __device__ __forceinline__ void round2(float& a, float& b);
template <> __device__ __forceinline__ void round2<float>(float&, float&) {}
template <> __device__ __forceinline__ void round2<__nv_bfloat16>(float& a, float& b) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&p);
    a = __uint_as_float(u << 16); b = __uint_as_float(u & 0xffff0000u);
}
template <> __device__ __forceinline__ void round2<__half>(float& a, float& b) {
    const float2 f = __half22float2(__floats2half2_rn(a, b));
    a = f.x; b = f.y;
}
__device__ __forceinline__ uint32_t trunc_u8(float t) {
    if (__float_as_uint(t) < 0x4B000000u)                              // +0 <= t < 2^23: every in-range pixel
        return __float_as_uint(__fadd_rz(t, 8388608.f)) & 0xffu;       // low mantissa bits of t + 2^23 = trunc(t)
    const float m = fabsf(t);                                          // rare: negative, huge or non-finite
    if (!(m < 2147483520.f)) return 0u;
    return (uint32_t)__float2int_rz(t) & 0xffu;
}
template <typename T>
__device__ __forceinline__ void to_u8x2(T x0, T x1, uint32_t& u0, uint32_t& u1) {
    float a = __fmul_rn(to_f32(x0), 0.5f), b = __fmul_rn(to_f32(x1), 0.5f);
    round2<T>(a, b);
    a = __fadd_rn(a, 0.5f); b = __fadd_rn(b, 0.5f);
    round2<T>(a, b);
    a = __fmul_rn(a, 255.f); b = __fmul_rn(b, 255.f);
    round2<T>(a, b);
    u0 = trunc_u8(a); u1 = trunc_u8(b);
}
template <typename T>
__device__ __forceinline__ uint32_t to_u8(T x) { uint32_t u0, u1; to_u8x2<T>(x, x, u0, u1); return u0; }

// one thread: 4 consecutive pixels -> 12 output bytes (B,G,R per pixel).  The 3 KB a CTA produces are contiguous in the
// output, so they go through shared memory (word stride 3: conflict-free) and leave as coalesced 16-byte stores.
// Host guarantees: pixels per image divisible by 1024 (CTAs never straddle images), 16-byte aligned pointers.
template <typename T>
__global__ void __launch_bounds__(256)
stage_u8_bgr_kernel(const T* __restrict__ images, int n, int H, int W, uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint32_t buf[256 * 3];
    const size_t plane = (size_t)H * W;
    const size_t first = (size_t)blockIdx.x * 1024;               // first pixel (over all images) of this CTA
    const size_t img = first / plane, pix = first - img * plane + (size_t)threadIdx.x * 4;
    const T* src = images + img * 3 * plane + pix;
    uint32_t v[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (sizeof(T) == 2) {
            const uint2 w = *reinterpret_cast<const uint2*>(src + c * plane);       // 4 x 16-bit
            const T* e = reinterpret_cast<const T*>(&w);
            to_u8x2<T>(e[0], e[1], v[c][0], v[c][1]); to_u8x2<T>(e[2], e[3], v[c][2], v[c][3]);
        } else {
            const float4 w = *reinterpret_cast<const float4*>(src + c * plane);
            const T* e = reinterpret_cast<const T*>(&w);
            to_u8x2<T>(e[0], e[1], v[c][0], v[c][1]); to_u8x2<T>(e[2], e[3], v[c][2], v[c][3]);
        }
    }
    // bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3      (B = channel 2, G = 1, R = 0)
    uint32_t* o = buf + threadIdx.x * 3;
    o[0] = v[2][0] | (v[1][0] << 8) | (v[0][0] << 16) | (v[2][1] << 24);
    o[1] = v[1][1] | (v[0][1] << 8) | (v[2][2] << 16) | (v[1][2] << 24);
    o[2] = v[0][2] | (v[2][3] << 8) | (v[1][3] << 16) | (v[0][3] << 24);
    __syncthreads();
    if (threadIdx.x < 192)
        reinterpret_cast<uint4*>(out + first * 3)[threadIdx.x] = reinterpret_cast<const uint4*>(buf)[threadIdx.x];
}

// any shape: one thread per pixel
template <typename T>
__global__ void __launch_bounds__(256)
stage_u8_bgr_generic_kernel(const T* __restrict__ images, int n, int H, int W, uint8_t* __restrict__ out) {
    const size_t plane = (size_t)H * W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane * (size_t)n) return;
    const size_t img = i / plane, pix = i - img * plane;
    const T* src = images + img * 3 * plane + pix;
    uint8_t* o = out + i * 3;
    o[0] = (uint8_t)to_u8<T>(src[2 * plane]); o[1] = (uint8_t)to_u8<T>(src[plane]); o[2] = (uint8_t)to_u8<T>(src[0]);
}

// ---- metrics: one CTA; per-thread counters, block reduction through shared-memory atomics (integers: exact, order-free)
constexpr int MC_G = 0, MC_R = 2, MC_GR = 6, MC_A = 14, MC_LOW = 16, MC_N = 19, MC_TOTAL = 22;   // counter layout

template <typename T>
__device__ __forceinline__ bool row_valid(const T* p, int w) {
    bool ok = true;
    for (int q = 0; q < w; q++) ok = ok && to_f32(p[q]) != -1.f;
    return ok;
}
template <typename T>
__device__ __forceinline__ int row_argmax(const T* p, int w, float* mx) {
    int best = 0; float bv = to_f32(p[0]);
    for (int q = 1; q < w; q++) { const float x = to_f32(p[q]); if (x > bv) { bv = x; best = q; } }    // first maximum wins
    *mx = bv;
    return best;
}
__device__ __forceinline__ double mean_pairwise_gap(const float* f, int N) {
    // torch.cdist(f, f, p=1) without its diagonal, .mean(): fp32 differences, fp32 sum, one division
    float s = 0.f;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            if (i != j) s = __fadd_rn(s, fabsf(__fsub_rn(f[i], f[j])));
    return (double)__fdiv_rn(s, (float)(N * (N - 1)));
}

template <typename T>
__global__ void __launch_bounds__(256)
bias_metrics_kernel(const T* __restrict__ pg, const T* __restrict__ pr, const T* __restrict__ pa, int n, double* __restrict__ out) {
    __shared__ int cnt[MC_TOTAL];
    for (int e = threadIdx.x; e < MC_TOTAL; e += blockDim.x) cnt[e] = 0;
    __syncthreads();
    const float thr = round_to<T>(0.8f);          // `tensor < 0.8` compares in the tensor's dtype
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool vg = pg && row_valid(pg + 2 * i, 2), vr = row_valid(pr + 4 * i, 4);
        int g = -1, r = -1; float mx;
        if (vg) { g = row_argmax(pg + 2 * i, 2, &mx); atomicAdd(&cnt[MC_G + g], 1); atomicAdd(&cnt[MC_N + 0], 1); if (mx < thr) atomicAdd(&cnt[MC_LOW + 0], 1); }
        if (vr) { r = row_argmax(pr + 4 * i, 4, &mx); atomicAdd(&cnt[MC_R + r], 1); atomicAdd(&cnt[MC_N + 1], 1); if (mx < thr) atomicAdd(&cnt[MC_LOW + 1], 1); }
        if (vg && vr) atomicAdd(&cnt[MC_GR + g * 4 + r], 1);
        if (pa && row_valid(pa + 2 * i, 2)) {
            const int a = row_argmax(pa + 2 * i, 2, &mx);
            atomicAdd(&cnt[MC_A + a], 1); atomicAdd(&cnt[MC_N + 2], 1); if (mx < thr) atomicAdd(&cnt[MC_LOW + 2], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // a mean of 0/1 floats is count / n in fp32 (the sum is an exact integer)
        const float ng = (float)cnt[MC_N + 0], nr = (float)cnt[MC_N + 1];
        float fg[2], fr[4], fgr[8];
        if (!pg) {
            // exp-6 get_evaluate_metrics(probs_race_all), E6:1624-1638: the four race frequencies, their mean pairwise
            // gap and the share of faces whose top race probability is below 0.8
            for (int q = 0; q < 4; q++) { fr[q] = __fdiv_rn((float)cnt[MC_R + q], nr); out[q] = (double)fr[q]; }
            out[4] = mean_pairwise_gap(fr, 4);
            out[5] = (double)__fdiv_rn((float)cnt[MC_LOW + 1], nr);
            return;
        }
        for (int q = 0; q < 2; q++) fg[q] = __fdiv_rn((float)cnt[MC_G + q], ng);
        for (int q = 0; q < 4; q++) fr[q] = __fdiv_rn((float)cnt[MC_R + q], nr);
        for (int q = 0; q < 8; q++) fgr[q] = __fdiv_rn((float)cnt[MC_GR + q], ng);
        out[0] = (double)fabsf(__fsub_rn(fg[1], fg[0]));                     // gender_gap
        out[1] = (double)__fdiv_rn((float)cnt[MC_LOW + 0], ng);              // gender_pred_below_08
        out[2] = mean_pairwise_gap(fr, 4);                                   // race_gap
        out[3] = (double)__fdiv_rn((float)cnt[MC_LOW + 1], nr);              // race_pred_below_08
        out[4] = mean_pairwise_gap(fgr, 8);                                  // gender_race_gap
        if (pa) {
            const float na = (float)cnt[MC_N + 2];
            const double a0 = (double)__fdiv_rn((float)cnt[MC_A + 0], na), a1 = (double)__fdiv_rn((float)cnt[MC_A + 1], na);
            out[5] = a0; out[6] = a1;                                        // age0_freq, age1_freq
            out[7] = (double)__fdiv_rn((float)cnt[MC_LOW + 2], na);          // age_pred_below_08
            out[8] = (fabs(a0 - 0.75) + fabs(a1 - 0.25)) / 2;                // age_gap (python floats: fp64)
        }
    }
}

}  // namespace

extern "C" int fg_stage_detector_input(const void* images, int n, int C, int H, int W, uint8_t* out_bgr_hwc, int dtype, void* stream) {
    if (n < 0 || C != 3 || H <= 0 || W <= 0 || (n > 0 && (!images || !out_bgr_hwc))) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    const size_t plane = (size_t)H * W;
    const bool fast = plane % 1024 == 0 && ((uintptr_t)images % 16 == 0) && ((uintptr_t)out_bgr_hwc % 16 == 0);
    if (fast) {
        const size_t ctas = plane / 1024 * (size_t)n;
        if (ctas > 0x7fffffffull) return FG_ERR_LIMIT;
        FG_DISPATCH_DTYPE(dtype, T, stage_u8_bgr_kernel<T><<<(unsigned)ctas, 256, 0, fg_stream(stream)>>>((const T*)images, n, H, W, out_bgr_hwc));
    } else {
        const size_t px = plane * (size_t)n;
        if (px > 0x7fffffffull * 256) return FG_ERR_LIMIT;
        FG_DISPATCH_DTYPE(dtype, T, stage_u8_bgr_generic_kernel<T><<<(unsigned)((px + 255) / 256), 256, 0, fg_stream(stream)>>>((const T*)images, n, H, W, out_bgr_hwc));
    }
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_bias_metrics(const void* probs_gender, const void* probs_race, const void* probs_age, int n, double* out,
                               int dtype, void* stream) {
    if (n < 0 || !out || (n > 0 && !probs_race) || (!probs_gender && probs_age)) return FG_ERR_INVALID_ARG;
    FG_DISPATCH_DTYPE(dtype, T, bias_metrics_kernel<T><<<1, 256, 0, fg_stream(stream)>>>((const T*)probs_gender, (const T*)probs_race, (const T*)probs_age, n, out));
    FG_LAUNCH_CHECK();
    return FG_OK;
}
