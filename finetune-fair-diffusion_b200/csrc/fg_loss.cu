// Fairness cross-entropy: CE_loss(logits[idx], targets[idx]) on idx = face & target != -1, with a
// -1 placeholder elsewhere (E1:1912-1915, E3:2114-2122, E4:2238-2251); CE_loss is
// nn.CrossEntropyLoss(reduction="none") (E1:968).  One thread per image, k <= 64 classes.
#include "fg_common.cuh"

namespace {

template <typename T>
__global__ void fair_ce_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ targets,
                                   const uint8_t* __restrict__ face, int n, int k, float fill, T* __restrict__ loss) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long t = targets[i];
    if (!face[i] || t == -1 || t < 0 || t >= k) { loss[i] = from_f32<T>(fill); return; }
    const T* lg = logits + (size_t)i * k;
    float mx = -INFINITY;
    for (int q = 0; q < k; q++) mx = fmaxf(mx, to_f32(lg[q]));
    float den = 0.f;
    for (int q = 0; q < k; q++) den += expf(to_f32(lg[q]) - mx);
    loss[i] = from_f32<T>(logf(den) + mx - to_f32(lg[t]));
}

template <typename T>
__global__ void fair_ce_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ targets,
                                   const uint8_t* __restrict__ face, const T* __restrict__ g_loss,
                                   int n, int k, T* __restrict__ g_logits) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long t = targets[i];
    T* go = g_logits + (size_t)i * k;
    if (!face[i] || t == -1 || t < 0 || t >= k) {
        for (int q = 0; q < k; q++) go[q] = from_f32<T>(0.f);
        return;
    }
    const T* lg = logits + (size_t)i * k;
    float mx = -INFINITY;
    for (int q = 0; q < k; q++) mx = fmaxf(mx, to_f32(lg[q]));
    float den = 0.f;
    for (int q = 0; q < k; q++) den += expf(to_f32(lg[q]) - mx);
    float g = to_f32(g_loss[i]);
    for (int q = 0; q < k; q++) {
        float p = expf(to_f32(lg[q]) - mx) / den;
        go[q] = from_f32<T>(g * (p - (q == t ? 1.f : 0.f)));
    }
}

}  // namespace

extern "C" int fg_fair_ce_fwd(const void* logits, const int64_t* targets, const uint8_t* face_indicators,
                              int n, int k, float fill, void* loss, int dtype, void* stream) {
    if (n < 0 || k <= 0 || k > 64 || !loss || (n > 0 && (!logits || !targets || !face_indicators))) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    FG_DISPATCH_DTYPE(dtype, T,
        fair_ce_fwd_kernel<T><<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(
            (const T*)logits, (const long long*)targets, face_indicators, n, k, fill, (T*)loss));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_fair_ce_bwd(const void* logits, const int64_t* targets, const uint8_t* face_indicators,
                              const void* g_loss, int n, int k, void* g_logits, int dtype, void* stream) {
    if (n < 0 || k <= 0 || k > 64 || !g_logits || (n > 0 && (!logits || !targets || !face_indicators || !g_loss))) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    FG_DISPATCH_DTYPE(dtype, T,
        fair_ce_bwd_kernel<T><<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(
            (const T*)logits, (const long long*)targets, face_indicators, (const T*)g_loss, n, k, (T*)g_logits));
    FG_LAUNCH_CHECK();
    return FG_OK;
}
