// Face crop / whole-image resize (forward) and the fused image gradient (backward).
//
//   crop_face                E1:267-290   slice -> constant pad -> Resize([224,224])
//   transforms.Resize(224)   E1:1905      same sampler over the whole image
//   apply_grad_hook_face     E1:1584-1617 / E3:1751-1784 / E4:1823-1867 (backward scaling)
//
// Sampler = ATen upsample_bilinear2d, align_corners=False, antialias off (torchvision 0.16.2):
//   scale = in/out (float);  src = max(scale*(dst+0.5)-0.5, 0);  i0 = floor(src);
//   i1 = i0 + (i0 < in-1);   l1 = src - i0;  l0 = 1 - l1
// applied to the VIRTUAL padded box: box pixel (py,px) is image pixel (y0+py, x0+px) when that
// lies inside the image and `fill` otherwise; i1 clamps at the box edge, not the image edge.
//
// Work decomposition.  A "job" is one output plane set: job 2i = small_i (box = whole image),
// job 2i+1 = chip_i.  Both jobs of an image are adjacent in block order, so the second read of
// the image region is served by the 126 MB L2 and HBM sees each image once.
#include "fg_common.cuh"
#include <cstdlib>

namespace {

struct Axis {
    int i0, i1;
    float l0, l1;
};

// identical expression order to ATen's area_pixel_compute_source_index (fp32, no contraction)
__device__ __forceinline__ Axis axis_index(int dst, float scale, int in_size) {
    float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    src = src < 0.f ? 0.f : src;
    Axis a;
    a.i0 = (int)src;
    if (a.i0 > in_size - 1) a.i0 = in_size - 1;
    a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
    a.l1 = __fsub_rn(src, (float)a.i0);
    a.l1 = a.l1 < 0.f ? 0.f : (a.l1 > 1.f ? 1.f : a.l1);
    a.l0 = __fsub_rn(1.f, a.l1);
    return a;
}

struct Box {
    int x0, y0, x1, y1;
    bool ok;      // has a face, non-empty, overlaps the image
};

__device__ __forceinline__ Box load_box(const long long* __restrict__ boxes, const uint8_t* __restrict__ ind, int i, int H, int W) {
    Box b;
    long long x0 = boxes[4 * i + 0], y0 = boxes[4 * i + 1], x1 = boxes[4 * i + 2], y1 = boxes[4 * i + 3];
    const long long lim = 1 << 24;
    b.ok = (!ind || ind[i]) && x1 > x0 && y1 > y0 && x1 > 0 && y1 > 0 && x0 < W && y0 < H &&
           x0 > -lim && y0 > -lim && x1 < lim && y1 < lim;
    b.x0 = (int)x0; b.y0 = (int)y0; b.x1 = (int)x1; b.y1 = (int)y1;
    return b;
}

// ------------------------------------------------------------------------------ forward
// block = (32, 8): 32 output columns x 8 output rows, all C channels per thread.
template <typename T>
__global__ void __launch_bounds__(256)
sample_fwd_kernel(const T* __restrict__ images, int n, int C, int H, int W,
                  const long long* __restrict__ boxes, const uint8_t* __restrict__ ind,
                  T* __restrict__ chips, int ch, int cw, T* __restrict__ small, int sh, int sw,
                  float fill, int tiles_x_chip, int tiles_chip, int tiles_x_small, int tiles_small) {
    const int per_image = tiles_chip + tiles_small;
    int img = blockIdx.x / per_image;
    int t = blockIdx.x - img * per_image;
    bool is_small = t < tiles_small;
    T* out; int oh, ow, tx, ty;
    Box b;
    if (is_small) {
        out = small + (size_t)img * C * sh * sw; oh = sh; ow = sw;
        ty = t / tiles_x_small; tx = t - ty * tiles_x_small;
        b.x0 = 0; b.y0 = 0; b.x1 = W; b.y1 = H; b.ok = true;
    } else {
        t -= tiles_small;
        out = chips + (size_t)img * C * ch * cw; oh = ch; ow = cw;
        ty = t / tiles_x_chip; tx = t - ty * tiles_x_chip;
        b = load_box(boxes, ind, img, H, W);
    }
    int ox = tx * 32 + threadIdx.x;
    int oy = ty * 8 + threadIdx.y;
    if (ox >= ow || oy >= oh) return;
    const size_t oplane = (size_t)oh * ow;
    T* o = out + (size_t)oy * ow + ox;
    if (!b.ok) {
        T f = from_f32<T>(fill);
        for (int c = 0; c < C; c++) o[c * oplane] = f;
        return;
    }
    int bw = b.x1 - b.x0, bh = b.y1 - b.y0;
    Axis ax = axis_index(ox, (float)bw / (float)ow, bw);
    Axis ay = axis_index(oy, (float)bh / (float)oh, bh);
    int xa = b.x0 + ax.i0, xb = b.x0 + ax.i1, ya = b.y0 + ay.i0, yb = b.y0 + ay.i1;
    bool xa_in = xa >= 0 && xa < W, xb_in = xb >= 0 && xb < W;
    bool ya_in = ya >= 0 && ya < H, yb_in = yb >= 0 && yb < H;
    const T* src = images + (size_t)img * C * H * W;
    const size_t iplane = (size_t)H * W;
    for (int c = 0; c < C; c++) {
        const T* p = src + c * iplane;
        float v00 = (ya_in && xa_in) ? to_f32(p[(size_t)ya * W + xa]) : fill;
        float v01 = (ya_in && xb_in) ? to_f32(p[(size_t)ya * W + xb]) : fill;
        float v10 = (yb_in && xa_in) ? to_f32(p[(size_t)yb * W + xa]) : fill;
        float v11 = (yb_in && xb_in) ? to_f32(p[(size_t)yb * W + xb]) : fill;
        float top = ax.l0 * v00 + ax.l1 * v01;
        float bot = ax.l0 * v10 + ax.l1 * v11;
        o[c * oplane] = from_f32<T>(ay.l0 * top + ay.l1 * bot);
    }
}

// ------------------------------------------------------------------------------ backward
// Contribution of one output grid (oh x ow sampled from the virtual box) to image pixel (y,x):
//   sum over outputs whose 2x2 footprint covers the pixel of  wy * wx * g[oy][ox].
// Candidates come from inverting src(o) conservatively; each is re-checked with the exact
// forward index function, so forward and backward agree on every footprint.
struct Span { int lo, hi; };   // inclusive candidate range of output indices

__device__ __forceinline__ Span candidates(int p /*box coord*/, float scale, int out_size) {
    // src(o) in [p-1, p+1)  <=>  o in ((p-0.5)/scale - 0.5, (p+1.5)/scale - 0.5)
    float inv = 1.f / scale;
    int lo = (int)floorf(((float)p - 0.5f) * inv - 0.5f) - 1;
    int hi = (int)ceilf(((float)p + 1.5f) * inv - 0.5f) + 1;
    Span s;
    s.lo = lo < 0 ? 0 : lo;
    s.hi = hi > out_size - 1 ? out_size - 1 : hi;
    return s;
}

__device__ __forceinline__ float axis_weight(int o, int p, float scale, int in_size) {
    Axis a = axis_index(o, scale, in_size);
    return (a.i0 == p ? a.l0 : 0.f) + (a.i1 == p ? a.l1 : 0.f);
}

#define FG_MAXC 4

template <typename T>
__device__ __forceinline__ void gather_grid(const T* __restrict__ g, int C, int oh, int ow,
                                            int px, int py, int bw, int bh, float acc[FG_MAXC]) {
    float sx = (float)bw / (float)ow, sy = (float)bh / (float)oh;
    Span cx = candidates(px, sx, ow), cy = candidates(py, sy, oh);
    const size_t oplane = (size_t)oh * ow;
    for (int oy = cy.lo; oy <= cy.hi; oy++) {
        float wy = axis_weight(oy, py, sy, bh);
        if (wy == 0.f) continue;
        for (int ox = cx.lo; ox <= cx.hi; ox++) {
            float wx = axis_weight(ox, px, sx, bw);
            if (wx == 0.f) continue;
            float w = wy * wx;
            const T* q = g + (size_t)oy * ow + ox;
            for (int c = 0; c < C; c++) acc[c] += w * to_f32(q[c * oplane]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
image_grad_kernel(const T* __restrict__ g_chips, const T* __restrict__ g_small,
                  const long long* __restrict__ boxes, const uint8_t* __restrict__ ind,
                  const int32_t* __restrict__ region, const float* __restrict__ scale,
                  T* __restrict__ g_images, int n, int C, int H, int W, int ch, int cw, int sh, int sw) {
    int x = blockIdx.x * 32 + threadIdx.x;
    int y = blockIdx.y * 8 + threadIdx.y;
    int img = blockIdx.z;
    if (x >= W || y >= H) return;
    float acc_s[FG_MAXC] = {0.f, 0.f, 0.f, 0.f};
    float acc_c[FG_MAXC] = {0.f, 0.f, 0.f, 0.f};
    if (g_small) {
        gather_grid<T>(g_small + (size_t)img * C * sh * sw, C, sh, sw, x, y, W, H, acc_s);
        if (region) {
            const int32_t* r = region + 4 * img;
            if (x >= r[0] && x < r[2] && y >= r[1] && y < r[3]) {
                float s = scale[img];
                for (int c = 0; c < C; c++) acc_s[c] *= s;
            }
        }
    }
    if (g_chips) {
        Box b = load_box(boxes, ind, img, H, W);
        if (b.ok && x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1)
            gather_grid<T>(g_chips + (size_t)img * C * ch * cw, C, ch, cw, x - b.x0, y - b.y0, b.x1 - b.x0, b.y1 - b.y0, acc_c);
    }
    T* o = g_images + (size_t)img * C * H * W + (size_t)y * W + x;
    const size_t iplane = (size_t)H * W;
    for (int c = 0; c < C; c++) o[c * iplane] = from_f32<T>(acc_s[c] + acc_c[c]);
}

template <typename T>
__global__ void region_scale_kernel(const T* __restrict__ g_in, const int32_t* __restrict__ region,
                                    const float* __restrict__ scale, T* __restrict__ g_out, int C, int H, int W) {
    int img = blockIdx.z;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const int32_t* r = region + 4 * img;
    float s = (x >= r[0] && x < r[2] && y >= r[1] && y < r[3]) ? scale[img] : 1.f;
    size_t base = (size_t)img * C * H * W + (size_t)y * W + x;
    for (int c = 0; c < C; c++) {
        size_t k = base + (size_t)c * H * W;
        g_out[k] = s == 1.f ? g_in[k] : from_f32<T>(s * to_f32(g_in[k]));
    }
}

#include "fg_sample_tiled.cuh"
#include "fg_image_grad_staged.cuh"
#include "fg_image_grad_quad.cuh"

// ------------------------------------------------------------------------------ factors
// python slice semantics: a negative stop counts from the end (E1:1594-1597 with a -1 box)
__device__ __forceinline__ int slice_stop(long long stop, int size) {
    if (stop < 0) { stop += size; if (stop < 0) stop = 0; }
    if (stop > size) stop = size;
    return (int)stop;
}
__device__ __forceinline__ long long max3(long long a, long long b, long long c) { return max(max(a, b), c); }
__device__ __forceinline__ long long min3(long long a, long long b, long long c) { return min(min(a, b), c); }

struct FactorArgs {
    const long long* targets[3];
    const long long* preds[3];
    float hook[3];
    float weight[3];
};

__device__ __forceinline__ float pick_factor(const FactorArgs& a, const float* f, int n_attr, int e1_rule, int i) {
    if (e1_rule) {
        long long t = a.targets[0][i], p = a.preds[0][i];
        if (t == -1) return f[0];
        return t == p ? 1.f : f[0];
    }
    float best = 0.f; bool any = false;
    for (int k = 0; k < n_attr; k++) {
        if (a.targets[k][i] != a.preds[k][i]) {
            best = any ? fminf(best, f[k]) : f[k];
            any = true;
        }
    }
    return any ? best : 1.f;
}

__global__ void factors_kernel(const uint8_t* __restrict__ face, const long long* __restrict__ bbox,
                               const long long* __restrict__ bbox_ori, FactorArgs a, int n_attr, int e1_rule,
                               int n, int H, int W, int32_t* __restrict__ region, float* __restrict__ scale,
                               float* __restrict__ weights) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (region || scale) {
        const long long* b = bbox + 4 * i;
        bool none = b[0] == -1 && b[1] == -1 && b[2] == -1 && b[3] == -1;
        int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
        float s = 1.f;
        if (!none) {
            const long long* o = bbox_ori + 4 * i;
            // reference: img_width, img_height = image.shape[1:]  (i.e. H, W -- names swapped)
            long long left = max3(b[0], o[0], 0), right = min3(b[2], o[2], (long long)H);
            long long bottom = max3(b[1], o[1], 0), top = min3(b[3], o[3], (long long)W);
            rx0 = (int)min(left, (long long)W);  rx1 = slice_stop(right, W);
            ry0 = (int)min(bottom, (long long)H); ry1 = slice_stop(top, H);
            if (rx1 <= rx0 || ry1 <= ry0) { rx0 = ry0 = rx1 = ry1 = 0; }
            s = pick_factor(a, a.hook, n_attr, e1_rule, i);
        }
        if (region) { region[4 * i] = rx0; region[4 * i + 1] = ry0; region[4 * i + 2] = rx1; region[4 * i + 3] = ry1; }
        if (scale) scale[i] = s;
    }
    if (weights) {
        if (!face[i]) {
            float w = 1.f;
            if (!e1_rule) { w = a.weight[0]; for (int k = 1; k < n_attr; k++) w = fminf(w, a.weight[k]); }
            weights[i] = w;
        } else {
            weights[i] = pick_factor(a, a.weight, n_attr, e1_rule, i);
        }
    }
}

// ------------------------------------------------------------------------------ tiled launches
constexpr int FG_NOT_TILED = 0x7ffffff0;     // shape not covered by the tiled kernels: use the generic ones

template <typename T, int STAGES>
static int launch_fwd_tiled_t(const FwdParams& p, size_t smem, unsigned grid, cudaStream_t st) {
    cudaError_t e;
    if (p.W == 512) {                              // the BASELINE image width: ring row offsets become immediates
        e = cudaFuncSetAttribute(sample_fwd_tiled_kernel<T, 3, STAGES, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sample_fwd_tiled_kernel<T, 3, STAGES, 512><<<grid, FWD_THREADS, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(sample_fwd_tiled_kernel<T, 3, STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        sample_fwd_tiled_kernel<T, 3, STAGES, 0><<<grid, FWD_THREADS, smem, st>>>(p);
    }
    FG_LAUNCH_CHECK();
    return FG_OK;
}

static int launch_fwd_tiled(const void* images, int n, int C, int H, int W, const int64_t* boxes, const uint8_t* indicators,
                            void* chips, int chip_h, int chip_w, void* small, int small_h, int small_w,
                            float fill_value, int dtype, cudaStream_t st) {
    const size_t esz = dtype == FG_F32 ? 4 : 2;
    const size_t row_bytes = (size_t)W * esz;
    const int stages = dtype == FG_F32 ? 2 : FG_FWD_STAGES16;
    const int rmax = dtype == FG_F32 ? 2 * TOH : FG_FWD_RMAX16;             // ring rows per channel (RMAX in the kernel)
    const size_t meta = ((size_t)stages * sizeof(FwdMeta) + 127) / 128 * 128;
    const size_t smem = 128 + meta + (size_t)stages * rmax * 3 * row_bytes;
    FwdParams p;
    p.tiles_small = small ? (small_h + TOH - 1) / TOH : 0;
    p.tiles_chip = chips ? (chip_h + TOH - 1) / TOH : 0;
    const long long total = (long long)n * (p.tiles_small + p.tiles_chip);
    const bool ok = C == 3 && (row_bytes % 16 == 0) && ((uintptr_t)images % 16 == 0) && smem <= 110 * 1024 && total < 0x7fffffffLL;
    if (!ok) return FG_NOT_TILED;
    p.images = images; p.n = n; p.C = C; p.H = H; p.W = W;
    p.boxes = (const long long*)boxes; p.ind = indicators;
    p.chips = chips; p.ch = chip_h; p.cw = chip_w; p.small = small; p.sh = small_h; p.sw = small_w;
    p.fill = fill_value; p.total_tiles = (int)total;
    static const bool fwd_pair = getenv("FG_FWD_PAIR") != nullptr;      // tuning A/B switch, read once
    p.pair_mode = fwd_pair ? 1 : 2;                    // 1: two columns per thread everywhere
    #ifndef FG_FWD_CTAS_PER_SM
#define FG_FWD_CTAS_PER_SM 4
#endif
    const long long resident = (long long)(dtype == FG_F32 ? 2 : FG_FWD_CTAS_PER_SM) * FG_NUM_SMS;     // persistent CTAs
    const unsigned grid = (unsigned)(total < resident ? total : resident);
    switch (dtype) {
        case FG_F32: return launch_fwd_tiled_t<float, 2>(p, smem, grid, st);
        case FG_BF16: return launch_fwd_tiled_t<__nv_bfloat16, FG_FWD_STAGES16>(p, smem, grid, st);
        case FG_F16: return launch_fwd_tiled_t<__half, FG_FWD_STAGES16>(p, smem, grid, st);
        default: return FG_ERR_DTYPE;
    }
}

template <typename T>
static int launch_bwd_tiled_t(const BwdParams& p, int owp, size_t smem, dim3 grid, cudaStream_t st) {
    const bool spec = p.H == 512 && p.W == 512 && (!p.g_small || (p.sh == 224 && p.sw == 224)) && (!p.g_chips || (p.ch == 224 && p.cw == 224));
    cudaError_t e;
    {
        // the BASELINE shape: gradient rows staged through shared memory by bulk async copies
        static const bool staged_off = getenv("FG_BWD_GATHER") != nullptr;      // A/B switches for kernel tuning runs (read once)
        // 16-bit: the first-generation staged kernel stays the default (0.54 ms vs 0.56 ms for 1024 images on B200: at 2 bytes
        // per element both are bound by instruction issue / latency, not by the L1 data pipe the second generation relieves);
        // FG_BWD_QUAD=1 selects the second generation for A/B runs.  fp32 always takes the second generation (0.76 ms vs 1.43 ms).
        const bool v1 = getenv("FG_BWD_QUAD") == nullptr;                      // (dynamic: the parity tests toggle it in-process)
        const bool aligned = ((uintptr_t)p.g_small % 16 == 0) && ((uintptr_t)p.g_chips % 16 == 0);
        if (spec && aligned && !staged_off) {
            // rows per CTA: per-CTA set-up (tap tables) is amortised over NSUB * 8 rows, but the grid should still be
            // >= ~8 waves of 148 SMs x 3 resident CTAs; FG_BWD_NSUB overrides for tuning runs only
            const int nsub_env = getenv("FG_BWD_NSUB") ? atoi(getenv("FG_BWD_NSUB")) : 0;
            const int nsub = nsub_env ? nsub_env : (p.n >= 896 ? 16 : p.n >= 448 ? 8 : 4);
            if (!v1 || sizeof(T) == 4) {
                // second generation (fg_image_grad_quad.cuh): four columns per thread, packed shared loads, fp32 and 16-bit
#define FG_LAUNCH_QUAD(NS)                                                                                                          \
                do {                                                                                                                \
                    using L = GqLayout<T, NS>;                                                                                      \
                    if (p.g_small) {                                                                                                \
                        e = cudaFuncSetAttribute(image_grad_quad_kernel<T, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total); \
                        if (e != cudaSuccess) return (int)e;                                                                        \
                        image_grad_quad_kernel<T, NS, true><<<dim3(512 / (NS * GQ_ROWS), p.n), 256, L::total, st>>>(p);             \
                    } else {                                                                                                        \
                        e = cudaFuncSetAttribute(image_grad_quad_kernel<T, NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total); \
                        if (e != cudaSuccess) return (int)e;                                                                        \
                        image_grad_quad_kernel<T, NS, false><<<dim3(512 / (NS * GQ_ROWS), p.n), 256, L::total, st>>>(p);            \
                    }                                                                                                               \
                } while (0)
                if (nsub == 16) FG_LAUNCH_QUAD(16); else if (nsub == 8) FG_LAUNCH_QUAD(8); else FG_LAUNCH_QUAD(4);
#undef FG_LAUNCH_QUAD
                FG_LAUNCH_CHECK();
                return FG_OK;
            }
          if constexpr (sizeof(T) == 2) {
            constexpr int F = 8 / GS_ROWS;          // sub-tiles of GS_ROWS rows: the CTA keeps its 128 / 64 / 32 image rows
#define FG_LAUNCH_STAGED(NS)                                                                                                        \
            do {                                                                                                                    \
                using L = GsLayout<NS>;                                                                                             \
                e = cudaFuncSetAttribute(image_grad_staged_kernel<T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total); \
                if (e != cudaSuccess) return (int)e;                                                                                \
                image_grad_staged_kernel<T, NS><<<dim3(512 / (NS * GS_ROWS), p.n), 256, L::total, st>>>(p);                         \
            } while (0)
            if (nsub == 16) FG_LAUNCH_STAGED(16 * F); else if (nsub == 8) FG_LAUNCH_STAGED(8 * F); else FG_LAUNCH_STAGED(4 * F);
#undef FG_LAUNCH_STAGED
            FG_LAUNCH_CHECK();
            return FG_OK;
          }
        }
    }
    if (spec) {
        e = cudaFuncSetAttribute(image_grad_tiled_kernel<T, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        image_grad_tiled_kernel<T, 3, true><<<grid, 256, smem, st>>>(p, owp);
    } else {
        e = cudaFuncSetAttribute(image_grad_tiled_kernel<T, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        image_grad_tiled_kernel<T, 3, false><<<grid, 256, smem, st>>>(p, owp);
    }
    FG_LAUNCH_CHECK();
    return FG_OK;
}

static int launch_bwd_tiled(const void* g_chips, const void* g_small, const int64_t* boxes, const uint8_t* indicators,
                            const int32_t* region, const float* scale, void* g_images, int n, int C, int H, int W,
                            int chip_h, int chip_w, int small_h, int small_w, int dtype, cudaStream_t st) {
    const int sw = g_small ? small_w : 0, cw = g_chips ? chip_w : 0;
    int owp = (sw > cw ? sw : cw) + TPAD;
    if (H == 512 && W == 512 && owp < 224 + TPAD) owp = 224 + TPAD;     // the specialised kernel assumes 224-wide rows
    const size_t smem = (size_t)2 * BSUB * BTH * sizeof(Tab) + (size_t)2 * 3 * BTH * owp * sizeof(float);
    const bool ok = C == 3 && W <= 512 && (W % 2 == 0) && ((uintptr_t)g_images % 16 == 0) && smem <= 100 * 1024 &&
                    (!g_small || (small_w <= W && small_h <= H));
    if (!ok) return FG_NOT_TILED;
    BwdParams p;
    p.g_chips = g_chips; p.g_small = g_small; p.boxes = (const long long*)boxes; p.ind = indicators;
    p.region = region; p.scale = scale; p.g_images = g_images;
    p.n = n; p.C = C; p.H = H; p.W = W; p.ch = chip_h; p.cw = chip_w; p.sh = small_h; p.sw = small_w;
    dim3 grid((H + BSUB * BTH - 1) / (BSUB * BTH), n);
    switch (dtype) {
        case FG_F32: return launch_bwd_tiled_t<float>(p, owp, smem, grid, st);
        case FG_BF16: return launch_bwd_tiled_t<__nv_bfloat16>(p, owp, smem, grid, st);
        case FG_F16: return launch_bwd_tiled_t<__half>(p, owp, smem, grid, st);
        default: return FG_ERR_DTYPE;
    }
}

}  // namespace

extern "C" int fg_crop_resize_fwd(const void* images, int n, int C, int H, int W,
                                  const int64_t* boxes, const uint8_t* indicators,
                                  void* chips, int chip_h, int chip_w,
                                  void* small, int small_h, int small_w,
                                  float fill_value, int dtype, void* stream) {
    if (n < 0 || C <= 0 || C > FG_MAXC || H <= 0 || W <= 0 || !images) return FG_ERR_INVALID_ARG;
    if (chips && (!boxes || chip_h <= 0 || chip_w <= 0)) return FG_ERR_INVALID_ARG;
    if (small && (small_h <= 0 || small_w <= 0)) return FG_ERR_INVALID_ARG;
    if (n == 0 || (!chips && !small)) return FG_OK;
    if (getenv("FG_FORCE_GENERIC") == nullptr) {
        int rc = launch_fwd_tiled(images, n, C, H, W, boxes, indicators, chips, chip_h, chip_w, small, small_h, small_w,
                                  fill_value, dtype, fg_stream(stream));
        if (rc != FG_NOT_TILED) return rc;
    }
    int txc = chips ? (chip_w + 31) / 32 : 0, tc = chips ? txc * ((chip_h + 7) / 8) : 0;
    int txs = small ? (small_w + 31) / 32 : 0, ts = small ? txs * ((small_h + 7) / 8) : 0;
    long long blocks = (long long)n * (tc + ts);
    if (blocks > 0x7fffffffLL) return FG_ERR_LIMIT;
    dim3 block(32, 8);
    FG_DISPATCH_DTYPE(dtype, T,
        sample_fwd_kernel<T><<<(unsigned)blocks, block, 0, fg_stream(stream)>>>(
            (const T*)images, n, C, H, W, (const long long*)boxes, indicators, (T*)chips, chip_h, chip_w,
            (T*)small, small_h, small_w, fill_value, txc, tc, txs, ts));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_image_grad(const void* g_chips, const void* g_small,
                             const int64_t* boxes, const uint8_t* indicators,
                             const int32_t* region, const float* scale,
                             void* g_images, int n, int C, int H, int W,
                             int chip_h, int chip_w, int small_h, int small_w, int dtype, void* stream) {
    if (n < 0 || C <= 0 || C > FG_MAXC || H <= 0 || W <= 0 || !g_images) return FG_ERR_INVALID_ARG;
    if (g_chips && (!boxes || chip_h <= 0 || chip_w <= 0)) return FG_ERR_INVALID_ARG;
    if (g_small && (small_h <= 0 || small_w <= 0)) return FG_ERR_INVALID_ARG;
    if ((region == nullptr) != (scale == nullptr)) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (n > 65535) return FG_ERR_LIMIT;
    if (getenv("FG_FORCE_GENERIC") == nullptr) {
        int rc = launch_bwd_tiled(g_chips, g_small, boxes, indicators, region, scale, g_images, n, C, H, W,
                                  chip_h, chip_w, small_h, small_w, dtype, fg_stream(stream));
        if (rc != FG_NOT_TILED) return rc;
    }
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
    FG_DISPATCH_DTYPE(dtype, T,
        image_grad_kernel<T><<<grid, block, 0, fg_stream(stream)>>>(
            (const T*)g_chips, (const T*)g_small, (const long long*)boxes, indicators, region, scale,
            (T*)g_images, n, C, H, W, chip_h, chip_w, small_h, small_w));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_region_scale(const void* g_in, const int32_t* region, const float* scale, void* g_out,
                               int n, int C, int H, int W, int dtype, void* stream) {
    if (n < 0 || C <= 0 || H <= 0 || W <= 0 || !g_in || !g_out || !region || !scale) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (n > 65535 || H > 65535) return FG_ERR_LIMIT;
    dim3 block(256), grid((W + 255) / 256, H, n);
    FG_DISPATCH_DTYPE(dtype, T,
        region_scale_kernel<T><<<grid, block, 0, fg_stream(stream)>>>((const T*)g_in, region, scale, (T*)g_out, C, H, W));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_guidance_factors(const uint8_t* face_indicators, const int64_t* bbox, const int64_t* bbox_ori,
                                   const int64_t* targets0, const int64_t* targets1, const int64_t* targets2,
                                   const int64_t* preds_ori0, const int64_t* preds_ori1, const int64_t* preds_ori2,
                                   const float* hook_factors, const float* weight_factors, int n_attr, int e1_rule,
                                   int n, int H, int W, int32_t* region, float* scale, float* weights, void* stream) {
    if (n < 0 || n_attr < 1 || n_attr > 3) return FG_ERR_INVALID_ARG;
    if (e1_rule && n_attr != 1) return FG_ERR_INVALID_ARG;
    if ((region || scale) && (!bbox || !bbox_ori || !hook_factors)) return FG_ERR_INVALID_ARG;
    if (weights && (!face_indicators || !weight_factors)) return FG_ERR_INVALID_ARG;
    const int64_t* t[3] = {targets0, targets1, targets2};
    const int64_t* p[3] = {preds_ori0, preds_ori1, preds_ori2};
    FactorArgs a;
    for (int k = 0; k < 3; k++) {
        if (k < n_attr && (!t[k] || !p[k])) return FG_ERR_INVALID_ARG;
        a.targets[k] = (const long long*)t[k];
        a.preds[k] = (const long long*)p[k];
        a.hook[k] = (hook_factors && k < n_attr) ? hook_factors[k] : 1.f;
        a.weight[k] = (weight_factors && k < n_attr) ? weight_factors[k] : 1.f;
    }
    if (n == 0) return FG_OK;
    factors_kernel<<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(face_indicators, (const long long*)bbox,
        (const long long*)bbox_ori, a, n_attr, e1_rule, n, H, W, region, scale, weights);
    FG_LAUNCH_CHECK();
    return FG_OK;
}
