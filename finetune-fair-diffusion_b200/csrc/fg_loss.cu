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

// All attributes of one head in one launch: CE per attribute (-1 placeholder where inactive), the per-image loss
// assembly  sum_a CE_a + w_img * dyn_w * (L_clip + L_dino) + w_face * L_face  (E3:2119-2147, E4:2253-2283; evaluated
// left to right in fp32 like the reference expression) and d(mean loss)/d(head logits).  One thread per image.
struct FusedLossParams {
    const void* logits[3]; const long long* targets[3]; int width[3]; int col_start[3]; int n_attr;
    const uint8_t* face; int n, k_head; float fill, g_coef;
    const float* dyn_w; const void* loss_clip; const void* loss_dino; const void* loss_face; float w_img, w_face;
    void* loss_fair; float* loss; float* g_logits;
};

template <typename T>
__global__ void fair_loss_fused_kernel(const FusedLossParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float* go = p.g_logits ? p.g_logits + (size_t)i * p.k_head : nullptr;
    if (go) for (int q = 0; q < p.k_head; q++) go[q] = 0.f;
    const float gc = to_f32(from_f32<T>(p.g_coef));
    float total = 0.f;
    for (int a = 0; a < p.n_attr; a++) {
        const int k = p.width[a];
        const long long t = p.targets[a][i];
        T* lf = reinterpret_cast<T*>(p.loss_fair) + (size_t)a * p.n + i;
        T v = from_f32<T>(p.fill);
        if (p.face[i] && t >= 0 && t < k) {
            const T* lg = reinterpret_cast<const T*>(p.logits[a]) + (size_t)i * k;
            float mx = -INFINITY;
            for (int q = 0; q < k; q++) mx = fmaxf(mx, to_f32(lg[q]));
            float den = 0.f;
            for (int q = 0; q < k; q++) den += expf(to_f32(lg[q]) - mx);
            v = from_f32<T>(logf(den) + mx - to_f32(lg[t]));
            if (go)
                for (int q = 0; q < k; q++) {
                    const float pr = expf(to_f32(lg[q]) - mx) / den;
                    go[p.col_start[a] + q] = to_f32(from_f32<T>(gc * (pr - (q == t ? 1.f : 0.f))));
                }
        }
        *lf = v;
        total = a == 0 ? to_f32(v) : __fadd_rn(total, to_f32(v));
    }
    if (p.loss) {
        if (p.dyn_w && p.loss_clip && p.loss_dino) {
            const float sem = __fadd_rn(to_f32(reinterpret_cast<const T*>(p.loss_clip)[i]), to_f32(reinterpret_cast<const T*>(p.loss_dino)[i]));
            total = __fadd_rn(total, __fmul_rn(__fmul_rn(p.w_img, p.dyn_w[i]), sem));
        }
        if (p.loss_face) total = __fadd_rn(total, __fmul_rn(p.w_face, to_f32(reinterpret_cast<const T*>(p.loss_face)[i])));
        p.loss[i] = total;
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

extern "C" int fg_fair_loss_fused(const void* const* logits_attr, const int64_t* const* targets, const int32_t* width,
                                  const int32_t* col_start, int n_attr, const uint8_t* face_indicators, int n, int k_head,
                                  float fill, float g_coef, const float* dyn_weights, const void* loss_clip,
                                  const void* loss_dino, const void* loss_face, float weight_img, float weight_face,
                                  void* loss_fair, float* loss, float* g_logits, int dtype, void* stream) {
    if (n < 0 || n_attr < 1 || n_attr > 3 || k_head <= 0 || !logits_attr || !targets || !width || !col_start || !loss_fair ||
        (n > 0 && !face_indicators))
        return FG_ERR_INVALID_ARG;
    FusedLossParams p;
    for (int a = 0; a < 3; a++) {
        p.logits[a] = a < n_attr ? logits_attr[a] : nullptr;
        p.targets[a] = a < n_attr ? (const long long*)targets[a] : nullptr;
        p.width[a] = a < n_attr ? width[a] : 0;
        p.col_start[a] = a < n_attr ? col_start[a] : 0;
        if (a < n_attr && (width[a] <= 0 || width[a] > 64 || col_start[a] < 0 || col_start[a] + width[a] > k_head ||
                           (n > 0 && (!logits_attr[a] || !targets[a]))))
            return FG_ERR_INVALID_ARG;
    }
    if (n == 0) return FG_OK;
    p.n_attr = n_attr; p.face = face_indicators; p.n = n; p.k_head = k_head; p.fill = fill; p.g_coef = g_coef;
    p.dyn_w = dyn_weights; p.loss_clip = loss_clip; p.loss_dino = loss_dino; p.loss_face = loss_face;
    p.w_img = weight_img; p.w_face = weight_face; p.loss_fair = loss_fair; p.loss = loss; p.g_logits = g_logits;
    FG_DISPATCH_DTYPE(dtype, T, fair_loss_fused_kernel<T><<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(p));
    FG_LAUNCH_CHECK();
    return FG_OK;
}
