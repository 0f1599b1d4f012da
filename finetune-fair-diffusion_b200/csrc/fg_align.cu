// Next row f1: the aligned 112x112 face chip and the face-realism loss.
//
//   image_pipeline                    E1:292-312     (img+1)/2*255 -> similarity warp onto the 5-point template
//                                                    (skimage SimilarityTransform.estimate + kornia.warp_affine,
//                                                    bilinear, zeros, align_corners=False) -> back to [-1,1]
//   get_face_feats                    E1:1179-1190   net(x) + net(flip x) -> float -> L2 normalise   (net is external)
//   FaceFeatsModel.semantic_search    E1:96-117      top-1 dot-product search in the normalised face database
//   face-realism loss                 E1:1917-1929   1 - <f, target>
//
// Geometry.  skimage's estimate is Umeyama's least-squares similarity; in 2-D the optimal proper rotation has the
// closed form R = [[a,-b],[b,a]] / |(a,b)| with a = A00 + A11, b = A10 - A01 of the 2x2 covariance A, and the sum of
// the (sign-corrected) singular values is |(a,b)|, so no SVD is needed.  kornia normalises pixel coordinates with
// 2/(size-1) on both sides but samples with align_corners=False, so output pixel (j,i) reads the source position
//     u = (j+0.5)(Wd-1)/Wd, v = (i+0.5)(Hd-1)/Hd;  (x,y) = M^-1 (u,v,1);  xs = x*Ws/(Ws-1) - 0.5, ys = y*Hs/(Hs-1) - 0.5,
// an affine map `dst pixel -> src pixel` that one thread per image composes in fp64 (C, 6 numbers).  Sampling in the
// 0..255 domain with zero padding and mapping back equals bilinear sampling of the [-1,1] image with out-of-range
// taps reading -1, which is what the kernels do (fp32 interpolation, no round trip through 0..255).
//
// Forward: one thread per output pixel, three channels, four taps through the read-only path (the face region of an
// image, ~200x200 pixels, is read roughly once and sits in L1/L2 between the rows of a CTA's tile).
// Backward: gather form like fg_image_grad -- every source pixel of the warped quad's bounding box sums the output
// gradients whose taps touch it (the candidates lie in the parallelogram C^-1(p + (-1,1)^2): 2x2 at the usual 2x
// down-scale), so the image gradient is written once, without atomics, deterministic; it either overwrites the whole
// image gradient (zeros outside the box) or accumulates into the one fg_image_grad produced.
#include "fg_common.cuh"
#include <cstdlib>
#include <math.h>

namespace {

constexpr int ALIGN_PARAMS = 16;    // per image: [0..5] C (dst->src), [6..9] inverse linear part, [10..11] inverse offset, [12..15] bbox x0,y0,x1,y1 (as float)

__constant__ double TEMPLATE_112[5][2] = {{38.2946, 51.6963}, {73.5318, 51.5014}, {56.0252, 71.7366}, {41.5493, 92.3655}, {70.7299, 92.2041}};

// one thread per image
template <typename T>
__global__ void align_matrices_kernel(const float* __restrict__ landmarks, const uint8_t* __restrict__ indicators, int n,
                                      int Hs, int Ws, int Hd, int Wd, double* __restrict__ M_out, float* __restrict__ params) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* P = params + (size_t)i * ALIGN_PARAMS;
    const bool face = !indicators || indicators[i];
    if (!face) {
        for (int k = 0; k < ALIGN_PARAMS; k++) P[k] = 0.f;
        P[12] = 1.f; P[13] = 1.f; P[14] = 0.f; P[15] = 0.f;            // empty bbox
        if (M_out) for (int k = 0; k < 6; k++) M_out[(size_t)i * 6 + k] = -1.0;
        return;
    }
    // Umeyama, src = detected landmarks, dst = template (E1:304-305)
    double sx = 0, sy = 0, dx = 0, dy = 0;
    for (int k = 0; k < 5; k++) { sx += landmarks[(size_t)i * 10 + 2 * k]; sy += landmarks[(size_t)i * 10 + 2 * k + 1]; dx += TEMPLATE_112[k][0]; dy += TEMPLATE_112[k][1]; }
    sx /= 5; sy /= 5; dx /= 5; dy /= 5;
    double a = 0, b = 0, var = 0;
    for (int k = 0; k < 5; k++) {
        const double px = landmarks[(size_t)i * 10 + 2 * k] - sx, py = landmarks[(size_t)i * 10 + 2 * k + 1] - sy;
        const double qx = TEMPLATE_112[k][0] - dx, qy = TEMPLATE_112[k][1] - dy;
        a += qx * px + qy * py;            // A00 + A11 (times 5)
        b += qy * px - qx * py;            // A10 - A01 (times 5)
        var += px * px + py * py;
    }
    const double nrm = sqrt(a * a + b * b);
    // scale * R = (|(a,b)| / var) * [[a,-b],[b,a]] / |(a,b)| = [[a,-b],[b,a]] / var   (NaN for coincident landmarks, like skimage)
    double m00 = a / var, m01 = -b / var, m10 = b / var, m11 = a / var;
    if (!(nrm > 0.0)) { m00 = m01 = m10 = m11 = nan(""); }
    double m02 = dx - (m00 * sx + m01 * sy), m12 = dy - (m10 * sx + m11 * sy);
    // M = torch.tensor(tform.params[0:2]).to(img.dtype) (E1:307): fp32 images round the matrix to fp32; 16-bit images keep
    // fp32 here (the reference's half-precision grid carries >= 0.5 px of coordinate noise that is not reproduced)
    m00 = (double)(float)m00; m01 = (double)(float)m01; m02 = (double)(float)m02;
    m10 = (double)(float)m10; m11 = (double)(float)m11; m12 = (double)(float)m12;
    if (M_out) { double* o = M_out + (size_t)i * 6; o[0] = m00; o[1] = m01; o[2] = m02; o[3] = m10; o[4] = m11; o[5] = m12; }
    // inverse of the similarity
    const double det = m00 * m11 - m01 * m10;
    const double i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
    const double i02 = -(i00 * m02 + i01 * m12), i12 = -(i10 * m02 + i11 * m12);
    // dst pixel (j,i) -> src pixel: u = (j+0.5)*ku, v = (i+0.5)*kv; xs = (i00 u + i01 v + i02)*gx - 0.5
    const double ku = (double)(Wd - 1) / Wd, kv = (double)(Hd - 1) / Hd;
    const double gx = (double)Ws / (Ws - 1), gy = (double)Hs / (Hs - 1);
    const double c00 = i00 * ku * gx, c01 = i01 * kv * gx, c02 = (i00 * 0.5 * ku + i01 * 0.5 * kv + i02) * gx - 0.5;
    const double c10 = i10 * ku * gy, c11 = i11 * kv * gy, c12 = (i10 * 0.5 * ku + i11 * 0.5 * kv + i12) * gy - 0.5;
    P[0] = (float)c00; P[1] = (float)c01; P[2] = (float)c02; P[3] = (float)c10; P[4] = (float)c11; P[5] = (float)c12;
    // src pixel -> dst pixel (real valued), for the gather backward: q = D (p - c.2)
    const double dc = c00 * c11 - c01 * c10;
    const double d00 = c11 / dc, d01 = -c01 / dc, d10 = -c10 / dc, d11 = c00 / dc;
    P[6] = (float)d00; P[7] = (float)d01; P[8] = (float)d10; P[9] = (float)d11;
    P[10] = (float)(-(d00 * c02 + d01 * c12)); P[11] = (float)(-(d10 * c02 + d11 * c12));
    // bounding box of the taps: corners of the output grid, one pixel of slack each side
    double x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
    for (int cj = 0; cj < 2; cj++) for (int ci = 0; ci < 2; ci++) {
        const double j = cj ? Wd - 1 : 0, r = ci ? Hd - 1 : 0;
        const double x = c00 * j + c01 * r + c02, y = c10 * j + c11 * r + c12;
        x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
    }
    if (!(dc == dc) || !isfinite(x0 + x1 + y0 + y1)) { P[12] = 1.f; P[13] = 1.f; P[14] = 0.f; P[15] = 0.f; return; }
    P[12] = (float)fmax(floor(x0) - 1.0, 0.0); P[13] = (float)fmax(floor(y0) - 1.0, 0.0);
    P[14] = (float)fmin(ceil(x1) + 1.0, (double)(Ws - 1)); P[15] = (float)fmin(ceil(y1) + 1.0, (double)(Hs - 1));
}

// forward: grid (ceil(Hd*Wd / 256), n)
template <typename T>
__global__ void __launch_bounds__(256)
aligned_warp_fwd_kernel(const T* __restrict__ images, const float* __restrict__ params, const uint8_t* __restrict__ indicators,
                        int C, int Hs, int Ws, int Hd, int Wd, float fill, T* __restrict__ out) {
    const int img = blockIdx.y;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= Hd * Wd) return;
    T* o = out + (size_t)img * C * Hd * Wd + q;
    if (indicators && !indicators[img]) {                                   // E1:1332: all fill_value
        for (int c = 0; c < C; c++) o[(size_t)c * Hd * Wd] = from_f32<T>(fill);
        return;
    }
    const float* P = params + (size_t)img * ALIGN_PARAMS;
    const int i = q / Wd, j = q - i * Wd;
    const float xs = fmaf(P[0], (float)j, fmaf(P[1], (float)i, P[2]));
    const float ys = fmaf(P[3], (float)j, fmaf(P[4], (float)i, P[5]));
    const float fx0 = floorf(xs), fy0 = floorf(ys);
    const bool finite = fabsf(xs) < 1e8f && fabsf(ys) < 1e8f;              // NaN / huge -> every tap out of range
    const float wx1 = finite ? xs - fx0 : 0.f, wy1 = finite ? ys - fy0 : 0.f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = finite ? (int)fx0 : -4, y0 = finite ? (int)fy0 : -4;
    const bool vx0 = x0 >= 0 && x0 < Ws, vx1 = x0 + 1 >= 0 && x0 + 1 < Ws;
    const bool vy0 = y0 >= 0 && y0 < Hs, vy1 = y0 + 1 >= 0 && y0 + 1 < Hs;
    const T* src = images + (size_t)img * C * Hs * Ws;
    for (int c = 0; c < C; c++) {
        const T* s = src + (size_t)c * Hs * Ws;
        // out-of-range taps are 0 in the 0..255 domain = -1 here
        const float v00 = (vx0 && vy0) ? to_f32(__ldg(s + (size_t)y0 * Ws + x0)) : -1.f;
        const float v01 = (vx1 && vy0) ? to_f32(__ldg(s + (size_t)y0 * Ws + x0 + 1)) : -1.f;
        const float v10 = (vx0 && vy1) ? to_f32(__ldg(s + (size_t)(y0 + 1) * Ws + x0)) : -1.f;
        const float v11 = (vx1 && vy1) ? to_f32(__ldg(s + (size_t)(y0 + 1) * Ws + x0 + 1)) : -1.f;
        const float top = fmaf(v01, wx1, v00 * wx0), bot = fmaf(v11, wx1, v10 * wx0);
        o[(size_t)c * Hd * Wd] = from_f32<T>(fmaf(bot, wy1, top * wy0));
    }
}

// backward: grid (AW_SLICES, n): the CTAs of an image share the 32x8-pixel tiles of the bounding box of its taps (the whole
// image when the zeros have to be written as well) round-robin.  thread = one source pixel, all channels.  History: a grid
// over all tiles of all images spent its time launching ~1M empty CTAs (2.4 ms for 1024 images); one persistent CTA per
// image (round-robin over the images) left a CTA alone with the ~175 tiles of a face, i.e. with a long dependent chain of
// L2 gathers, and the last images of the launch had the GPU to themselves (1.7 ms).  Measured slower and dropped: staging
// the tile's output-gradient footprint in shared memory (two block barriers per tile) and four pixels per thread.
constexpr int AW_SLICES = 48;
template <typename T>
__global__ void __launch_bounds__(256)
aligned_warp_bwd_kernel(const T* __restrict__ g_out, const float* __restrict__ params, const uint8_t* __restrict__ indicators,
                        int n, int C, int Hs, int Ws, int Hd, int Wd, int accumulate, T* __restrict__ g_images) {
    const int tx_ = threadIdx.x & 31, ty_ = threadIdx.x >> 5;
    {
        const int img = blockIdx.y;
        const float* P = params + (size_t)img * ALIGN_PARAMS;
        const bool face = !indicators || indicators[img];
        const int bx0 = (int)P[12], by0 = (int)P[13], bx1 = (int)P[14], by1 = (int)P[15];
        if (accumulate && (!face || bx1 < bx0 || by1 < by0)) return;
        // tile range: the bounding box (accumulate) or the whole image (overwrite)
        const int X0 = accumulate ? (bx0 & ~31) : 0, Y0 = accumulate ? (by0 & ~7) : 0;
        const int X1 = accumulate ? bx1 : Ws - 1, Y1 = accumulate ? by1 : Hs - 1;
        const int tiles_x = (X1 - X0) / 32 + 1, tiles_y = (Y1 - Y0) / 8 + 1;
        const float d00 = P[6], d01 = P[7], d10 = P[8], d11 = P[9], d02 = P[10], d12 = P[11];
        const float c00 = P[0], c01 = P[1], c02 = P[2], c10 = P[3], c11 = P[4], c12 = P[5];
        const float hx = fabsf(d00) + fabsf(d01), hy = fabsf(d10) + fabsf(d11);
        const T* go = g_out + (size_t)img * C * Hd * Wd;
        for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
            const int x = X0 + (tile % tiles_x) * 32 + tx_, y = Y0 + (tile / tiles_x) * 8 + ty_;
            if (x >= Ws || y >= Hs) continue;
            T* g = g_images + (size_t)img * C * Hs * Ws + (size_t)y * Ws + x;
            const bool inside = face && x >= bx0 && x <= bx1 && y >= by0 && y <= by1;
            if (!inside) {
                if (!accumulate) for (int c = 0; c < C; c++) g[(size_t)c * Hs * Ws] = from_f32<T>(0.f);
                continue;
            }
            // candidate output pixels: q = D p + d, within the parallelogram D (p + (-1,1)^2)
            const float qx = fmaf(d00, (float)x, fmaf(d01, (float)y, d02));
            const float qy = fmaf(d10, (float)x, fmaf(d11, (float)y, d12));
            const int j0 = max((int)ceilf(qx - hx - 1e-3f), 0), j1 = min((int)floorf(qx + hx + 1e-3f), Wd - 1);
            const int i0 = max((int)ceilf(qy - hy - 1e-3f), 0), i1 = min((int)floorf(qy + hy + 1e-3f), Hd - 1);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i = i0; i <= i1; i++) {
                // Output pixel (j,i) taps this pixel iff its source position lies in [x-1, x+1) x [y-1, y+1); relative to the
                // pixel that position is the affine image C (j - qx, i - qy), so most of the window is rejected by two FMAs
                // (with a margin for the fp32 rounding of q); survivors are decided by the forward's own expressions below.
                const float ri = (float)i - qy;
                const float ex = c01 * ri, ey = c11 * ri;
                for (int j = j0; j <= j1; j++) {
                    const float rj = (float)j - qx;
                    const float dxs = fmaf(c00, rj, ex), dys = fmaf(c10, rj, ey);
                    if (fabsf(dxs) > 1.02f || fabsf(dys) > 1.02f) continue;
                    // the same fp32 expressions as the forward, so the tap weights are bit-identical
                    const float xs = fmaf(c00, (float)j, fmaf(c01, (float)i, c02));
                    const float ys = fmaf(c10, (float)j, fmaf(c11, (float)i, c12));
                    const float fx0 = floorf(xs), fy0 = floorf(ys);
                    const int tx = x - (int)fx0, ty = y - (int)fy0;              // 0 or 1 when this pixel is a tap
                    if ((unsigned)tx > 1u || (unsigned)ty > 1u) continue;
                    const float wx1 = xs - fx0, wy1 = ys - fy0;
                    const float w = (tx ? wx1 : 1.f - wx1) * (ty ? wy1 : 1.f - wy1);
                    for (int c = 0; c < C && c < 4; c++) acc[c] = fmaf(w, to_f32(__ldg(go + (size_t)c * Hd * Wd + (size_t)i * Wd + j)), acc[c]);
                }
            }
            for (int c = 0; c < C && c < 4; c++) {
                T* gc = g + (size_t)c * Hs * Ws;
                *gc = from_f32<T>(accumulate ? to_f32(*gc) + acc[c] : acc[c]);
            }
        }
    }
}

// backward, second generation: CELL MAP.  The gather above tests a window of chip pixels per image pixel (4-9 candidates for
// ~1.2 real taps: ~200 instructions per pixel, 1.36 ms for 1024 images).  When the chip grid is coarser than the image grid
// (the usual case: a 112-pixel chip of a 150-300 pixel face), no two chip pixels fall into the same unit cell of the image,
// so the map  cell (floor xs, floor ys) -> chip pixel  is injective and can be built without atomics or ordering:
//   phase 1  every chip pixel whose source position falls into the tile's cells writes its (row, column) into the map
//            (shared memory; entries carry the tile's tag, so the map is never cleared);
//   phase 2  an image pixel (x, y) looks up the four cells (x - dx, y - dy): a hit is a chip pixel for which this pixel is
//            tap (dx, dy), with the forward's own fp32 expressions for the weight -- fixed summation order, deterministic.
// Tiles are 64 x 16 pixels (four rows per thread), the map is double-buffered: one block barrier per tile.  A second chip pixel
// arriving in an occupied cell (up-sampled faces) flags the tile, whose pixels then take the candidate scan of the first
// generation; images whose matrix cannot be injective skip phase 1 altogether.
constexpr int AT_W = 64, AT_H = 16;
constexpr int AM_W = AT_W + 1, AM_H = AT_H + 1, AM_STRIDE = AM_W + 1;

// the first generation's per-pixel candidate scan (see aligned_warp_bwd_kernel)
template <typename T>
__device__ __forceinline__ void aligned_scan_pixel(const T* __restrict__ go, const float* __restrict__ P, int x, int y, int C, int Hd, int Wd, float* acc) {
    const float d00 = P[6], d01 = P[7], d10 = P[8], d11 = P[9], d02 = P[10], d12 = P[11];
    const float c00 = P[0], c01 = P[1], c02 = P[2], c10 = P[3], c11 = P[4], c12 = P[5];
    const float hx = fabsf(d00) + fabsf(d01), hy = fabsf(d10) + fabsf(d11);
    const float qx = fmaf(d00, (float)x, fmaf(d01, (float)y, d02));
    const float qy = fmaf(d10, (float)x, fmaf(d11, (float)y, d12));
    const int j0 = max((int)ceilf(qx - hx - 1e-3f), 0), j1 = min((int)floorf(qx + hx + 1e-3f), Wd - 1);
    const int i0 = max((int)ceilf(qy - hy - 1e-3f), 0), i1 = min((int)floorf(qy + hy + 1e-3f), Hd - 1);
    for (int i = i0; i <= i1; i++) {
        const float ri = (float)i - qy;
        const float ex = c01 * ri, ey = c11 * ri;
        for (int j = j0; j <= j1; j++) {
            const float rj = (float)j - qx;
            const float dxs = fmaf(c00, rj, ex), dys = fmaf(c10, rj, ey);
            if (fabsf(dxs) > 1.02f || fabsf(dys) > 1.02f) continue;
            const float xs = fmaf(c00, (float)j, fmaf(c01, (float)i, c02));
            const float ys = fmaf(c10, (float)j, fmaf(c11, (float)i, c12));
            const float fx0 = floorf(xs), fy0 = floorf(ys);
            const int tx = x - (int)fx0, ty = y - (int)fy0;
            if ((unsigned)tx > 1u || (unsigned)ty > 1u) continue;
            const float wx1 = xs - fx0, wy1 = ys - fy0;
            const float w = (tx ? wx1 : 1.f - wx1) * (ty ? wy1 : 1.f - wy1);
            for (int c = 0; c < C && c < 4; c++) acc[c] = fmaf(w, to_f32(__ldg(go + (size_t)c * Hd * Wd + (size_t)i * Wd + j)), acc[c]);
        }
    }
}

constexpr int AM_CELLS = AM_H * AM_STRIDE;                  // words per map plane
constexpr int AM_PLANES = 6;                                // tag | wx1 | wy1 | three channel values
constexpr size_t AM_SMEM = (size_t)2 * AM_PLANES * AM_CELLS * sizeof(int);

template <typename T>
__global__ void __launch_bounds__(256)
aligned_warp_bwd_cell_kernel(const T* __restrict__ g_out, const float* __restrict__ params, const uint8_t* __restrict__ indicators,
                             int n, int C, int Hs, int Ws, int Hd, int Wd, int accumulate, T* __restrict__ g_images) {
    // per buffer: the cell map (tagged chip pixel) and, next to it, what that chip pixel contributes: its two fractional
    // weights and its three gradient values (fp32) -- written once by the chip pixel in phase 1, so that an image pixel's
    // phase 2 is shared-memory reads only
    extern __shared__ __align__(16) int am_smem[];
    __shared__ int conflict[2];
    const int tid = threadIdx.x, tx = tid & (AT_W - 1), ty = tid >> 6;        // 64 columns x 4 row groups
    for (int e = tid; e < 2 * AM_PLANES * AM_CELLS; e += 256) am_smem[e] = 0;     // tag 0 = never valid; stale values stay finite
    if (tid < 2) conflict[tid] = 0;
    __syncthreads();
    const size_t gplane = (size_t)Hd * Wd, iplane = (size_t)Hs * Ws;
    int it = 0;                                                                // tiles this CTA has processed (tag source)
    // persistent over the images: the shared-memory set-up above is paid once per CTA, not once per (image, slice)
    for (int img = blockIdx.y; img < n; img += gridDim.y) {
    const float* P = params + (size_t)img * ALIGN_PARAMS;
    const bool face = !indicators || indicators[img];
    const int bx0 = (int)P[12], by0 = (int)P[13], bx1 = (int)P[14], by1 = (int)P[15];
    const bool has_taps = face && bx1 >= bx0 && by1 >= by0;
    if (accumulate && !has_taps) continue;
    const int X0 = accumulate ? (bx0 & ~(AT_W - 1)) : 0, Y0 = accumulate ? (by0 & ~(AT_H - 1)) : 0;
    const int X1 = accumulate ? bx1 : Ws - 1, Y1 = accumulate ? by1 : Hs - 1;
    const int tiles_x = (X1 - X0) / AT_W + 1, tiles_y = (Y1 - Y0) / AT_H + 1;
    const float c00 = P[0], c01 = P[1], c02 = P[2], c10 = P[3], c11 = P[4], c12 = P[5];
    const float d00 = P[6], d01 = P[7], d10 = P[8], d11 = P[9], d02 = P[10], d12 = P[11];
    // injective cell map <=> every non-zero integer step of the chip grid moves the source position by >= 1 in x or y;
    // the four shortest steps decide (longer ones move further)
    bool unique = has_taps && C == 3;
    {
        const float sj[4] = {1.f, 0.f, 1.f, 1.f}, si[4] = {0.f, 1.f, 1.f, -1.f};
#pragma unroll
        for (int q = 0; q < 4; q++)
            unique = unique && fmaxf(fabsf(c00 * sj[q] + c01 * si[q]), fabsf(c10 * sj[q] + c11 * si[q])) >= 1.002f;
    }
    const T* go = g_out + (size_t)img * C * Hd * Wd;
    const unsigned inv_tiles_x = 0xFFFFFFFFu / (unsigned)tiles_x + 1u;       // exact quotient by __umulhi for tile < 2^16
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x, it++) {
        const int b = it & 1;
        const int tag = it + 1;                                  // < 2^17: at most a few hundred tiles per CTA
        int* map = am_smem + b * AM_PLANES * AM_CELLS;
        float* fw = reinterpret_cast<float*>(map + AM_CELLS);   // [2][AM_CELLS] wx1, wy1; then [3][AM_CELLS] values
        const int tyi = (int)__umulhi((unsigned)tile, inv_tiles_x), txi = tile - tyi * tiles_x;     // tile / tiles_x without the divide
        const int X = X0 + txi * AT_W, Y = Y0 + tyi * AT_H;
        const bool touches = has_taps && X <= bx1 && X + AT_W > bx0 && Y <= by1 && Y + AT_H > by0;      // block-uniform
        const bool mapped = touches && unique;
        if (mapped) {
            // phase 1: chip pixels with xs in [X-1, X+AT_W), ys in [Y-1, Y+AT_H): bounding box of that rectangle in chip space
            float jlo = INFINITY, jhi = -INFINITY, ilo = INFINITY, ihi = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float px = (q & 1) ? (float)(X + AT_W) : (float)(X - 1), py = (q & 2) ? (float)(Y + AT_H) : (float)(Y - 1);
                const float qj = fmaf(d00, px, fmaf(d01, py, d02)), qi = fmaf(d10, px, fmaf(d11, py, d12));
                jlo = fminf(jlo, qj); jhi = fmaxf(jhi, qj); ilo = fminf(ilo, qi); ihi = fmaxf(ihi, qi);
            }
            const int j0 = max((int)floorf(jlo) - 1, 0), j1 = min((int)ceilf(jhi) + 1, Wd - 1);
            const int i0 = max((int)floorf(ilo) - 1, 0), i1 = min((int)ceilf(ihi) + 1, Hd - 1);
            const int jw = j1 - j0 + 1, cand = jw > 0 && i1 >= i0 ? jw * (i1 - i0 + 1) : 0;
            const unsigned inv_jw = 0xFFFFFFFFu / (unsigned)(jw > 0 ? jw : 1) + 1u;
#pragma unroll 2
            for (int e = tid; e < cand; e += 256) {
                const int ei = (int)__umulhi((unsigned)e, inv_jw);
                const int i = i0 + ei, j = j0 + e - ei * jw;
                const float xs = fmaf(c00, (float)j, fmaf(c01, (float)i, c02));
                const float ys = fmaf(c10, (float)j, fmaf(c11, (float)i, c12));
                const float fx0 = floorf(xs), fy0 = floorf(ys);
                const int cx = (int)fx0 - (X - 1), cy = (int)fy0 - (Y - 1);
                if (fabsf(xs) < 1e8f && fabsf(ys) < 1e8f && (unsigned)cx < (unsigned)AM_W && (unsigned)cy < (unsigned)AM_H) {
                    const int cell = cy * AM_STRIDE + cx;
                    const T* gp = go + (size_t)i * Wd + j;
                    const float v0 = to_f32(__ldg(gp)), v1 = to_f32(__ldg(gp + gplane)), v2 = to_f32(__ldg(gp + 2 * gplane));
                    const int old = atomicExch(&map[cell], tag);
                    if (old == tag) conflict[b] = tag;                         // a second chip pixel in this cell
                    fw[cell] = xs - fx0; fw[AM_CELLS + cell] = ys - fy0;
                    fw[2 * AM_CELLS + cell] = v0; fw[3 * AM_CELLS + cell] = v1; fw[4 * AM_CELLS + cell] = v2;
                }
            }
        }
        __syncthreads();
        const bool use_map = mapped && conflict[b] != tag;
        const int x = X + tx;
        constexpr int RPT = AT_H / 4;                        // rows per thread
        T* gbase = g_images + (size_t)img * C * iplane + x;
        if (use_map) {
            // Branch-free phase 2: the read-modify-write operands of the four rows are requested first, every cell is a few
            // shared-memory reads, a miss contributes with weight zero.
            bool ok[RPT], any[RPT]; T praw[RPT][3]; float acc[RPT][3]; bool mv[RPT][4];
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int y = Y + ty + 4 * r;
                ok[r] = x < Ws && y < Hs && x >= bx0 && x <= bx1 && y >= by0 && y <= by1;
                bool hit = false;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    mv[r][q] = ok[r] && map[(ty + 4 * r + 1 - (q >> 1)) * AM_STRIDE + tx + 1 - (q & 1)] == tag;
                    hit = hit || mv[r][q];
                }
                any[r] = hit;
#pragma unroll
                for (int c = 0; c < 3; c++) { acc[r][c] = 0.f; praw[r][c] = from_f32<T>(0.f); }
                if (accumulate && any[r]) {                       // raw bits only: the conversion (first use) waits until after the cells
#pragma unroll
                    for (int c = 0; c < 3; c++) praw[r][c] = gbase[(size_t)c * iplane + (size_t)y * Ws];
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int dx = q & 1, dy = q >> 1;
                    const int cell = (ty + 4 * r + 1 - dy) * AM_STRIDE + tx + 1 - dx;
                    const float wx1 = fw[cell], wy1 = fw[AM_CELLS + cell];
                    const float w = mv[r][q] ? (dx ? wx1 : 1.f - wx1) * (dy ? wy1 : 1.f - wy1) : 0.f;
#pragma unroll
                    for (int c = 0; c < 3; c++) acc[r][c] = fmaf(w, fw[(2 + c) * AM_CELLS + cell], acc[r][c]);
                }
            }
#pragma unroll
            for (int r = 0; r < RPT; r++) {
                const int y = Y + ty + 4 * r;
                if (x >= Ws || y >= Hs) continue;
                T* g = gbase + (size_t)y * Ws;
                if (accumulate) {
                    if (any[r]) {
#pragma unroll
                        for (int c = 0; c < 3; c++) g[(size_t)c * iplane] = from_f32<T>(to_f32(praw[r][c]) + acc[r][c]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) g[(size_t)c * iplane] = from_f32<T>(acc[r][c]);
                }
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < RPT; r++) {
                const int y = Y + ty + 4 * r;
                if (x >= Ws || y >= Hs) continue;
                T* g = gbase + (size_t)y * Ws;
                const bool inside = touches && x >= bx0 && x <= bx1 && y >= by0 && y <= by1;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if (inside) aligned_scan_pixel<T>(go, P, x, y, C, Hd, Wd, acc);
                if (accumulate) {
                    if (inside) for (int c = 0; c < C && c < 4; c++) { T* gc = g + (size_t)c * iplane; *gc = from_f32<T>(to_f32(*gc) + acc[c]); }
                } else {
                    for (int c = 0; c < C && c < 4; c++) g[(size_t)c * iplane] = from_f32<T>(acc[c]);
                }
            }
        }
    }
    }   // images
}

// ---------------------------------------------------------------------------------------------
// face-realism loss.  Features are [n,d] rows (d = 512 for the reference's SFNet); one warp per row everywhere.

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// order-preserving 32-bit key of a float
__device__ __forceinline__ unsigned fkey32(float f) { const unsigned b = __float_as_uint(f); return (b >> 31) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float funkey32(unsigned k) { return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k); }

// get_face_feats tail (E1:1186-1189): feats.to(float) -> F.normalize(dim=-1) = x / max(|x|, 1e-12).
// Also clears the search keys of the rows (first kernel of the loss).
template <typename T>
__global__ void __launch_bounds__(256)
feats_normalize_kernel(const T* __restrict__ raw, int n, int d, float* __restrict__ f, float* __restrict__ inv_norm,
                       unsigned long long* __restrict__ keys_to_clear) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const T* x = raw + (size_t)row * d;
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) { const float v = to_f32(x[k]); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    for (int k = lane; k < d; k += 32) f[(size_t)row * d + k] = to_f32(x[k]) * inv;
    if (lane == 0) { if (inv_norm) inv_norm[row] = inv; if (keys_to_clear) keys_to_clear[row] = 0ull; }
}

// backward of the normalisation: g_x = (g_f - f <f, g_f>) * inv_norm
template <typename T>
__global__ void __launch_bounds__(256)
feats_normalize_bwd_kernel(const float* __restrict__ g_f, const float* __restrict__ f, const float* __restrict__ inv_norm,
                           int n, int d, T* __restrict__ g_raw) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    float dot = 0.f;
    for (int k = lane; k < d; k += 32) dot = fmaf(f[(size_t)row * d + k], g_f[(size_t)row * d + k], dot);
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    for (int k = lane; k < d; k += 32)
        g_raw[(size_t)row * d + k] = from_f32<T>((g_f[(size_t)row * d + k] - f[(size_t)row * d + k] * dot) * inv);
}

// which target a row uses (E1:1919-1928 / E3:2126-2143 / E4:2255-2272): 0 none, 1 the original image's features,
// 2 nearest database entry
template <typename T>
__global__ void face_mode_kernel(const uint8_t* __restrict__ face, const long long* __restrict__ t0, const long long* __restrict__ t1,
                                 const long long* __restrict__ t2, const long long* __restrict__ p0, const long long* __restrict__ p1,
                                 const long long* __restrict__ p2, const T* __restrict__ q0, const T* __restrict__ q1,
                                 const T* __restrict__ q2, int w0, int w1, int w2, float level, int search_needs_target, int n,
                                 uint8_t* __restrict__ mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long* ts[3] = {t0, t1, t2}; const long long* ps[3] = {p0, p1, p2};
    const T* qs[3] = {q0, q1, q2}; const int ws[3] = {w0, w1, w2};
    const float lv = round_to<T>(level);
    bool ori = face[i] != 0, has_t = true;
    for (int a = 0; a < 3; a++) {
        if (!ts[a]) continue;
        float mx = -INFINITY;
        for (int c = 0; c < ws[a]; c++) mx = fmaxf(mx, to_f32(qs[a][(size_t)i * ws[a] + c]));
        ori = ori && ts[a][i] != -1 && ts[a][i] == ps[a][i] && mx >= lv;
        if (a == 0) has_t = ts[a][i] != -1;
    }
    mode[i] = !face[i] ? 0 : ori ? 1 : (!search_needs_target || has_t) ? 2 : 0;
}

// FaceFeatsModel.semantic_search (E1:96-117): top-1 dot product of every selected query against the database.
// HBM-bound on the database: every warp streams database rows (coalesced 16-byte loads, the row stays in registers),
// the <= 32 queries of the CTA's group sit in shared memory; the best (score, row) per query is kept per warp and merged
// with one 64-bit atomicMax per query and CTA (key = score key << 32 | ~row: the lowest row wins ties).
constexpr int SEARCH_Q = 32;      // queries per group
constexpr int SEARCH_D = 512;     // feature width handled in registers (4 x float4 per lane)
__global__ void __launch_bounds__(256)
face_search_kernel(const float* __restrict__ queries, const uint8_t* __restrict__ selector, uint8_t want, int m,
                   const float* __restrict__ db, int D, int d, unsigned long long* __restrict__ keys,
                   const unsigned* __restrict__ only_if) {
    if (only_if && *only_if == 0u) return;                             // fallback of the tensor-core search: runs only when flagged
    extern __shared__ __align__(16) float q_s[];                       // [SEARCH_Q][d]
    __shared__ unsigned long long best_s[SEARCH_Q];
    const int group = blockIdx.y, q0 = group * SEARCH_Q;
    const int nq = min(SEARCH_Q, m - q0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool any = false;
    for (int q = 0; q < nq; q++) any = any || !selector || selector[q0 + q] == want;
    if (!any) return;
    for (int e = threadIdx.x; e < nq * d; e += 256) q_s[e] = queries[(size_t)q0 * d + e];
    if (threadIdx.x < SEARCH_Q) best_s[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long my_best = 0ull;                                // lane q keeps query q's best
    const int rows_per_cta = (D + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_cta, r_end = min(D, r_begin + rows_per_cta);
    if (d == SEARCH_D) {
        // the next row is fetched while the current one is multiplied (one row = 4 x 16-byte loads per lane)
        float4 v[4], nx[4];
        int r = r_begin + warp;
        if (r < r_end) {
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = __ldg(reinterpret_cast<const float4*>(db + (size_t)r * d) + lane + 32 * k);
        }
        for (; r < r_end; r += 8) {
            if (r + 8 < r_end) {
#pragma unroll
                for (int k = 0; k < 4; k++) nx[k] = __ldg(reinterpret_cast<const float4*>(db + (size_t)(r + 8) * d) + lane + 32 * k);
            }
            for (int q = 0; q < nq; q++) {
                const float4* qv = reinterpret_cast<const float4*>(q_s + (size_t)q * d);
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 w = qv[lane + 32 * k];
                    acc = fmaf(v[k].x, w.x, acc); acc = fmaf(v[k].y, w.y, acc); acc = fmaf(v[k].z, w.z, acc); acc = fmaf(v[k].w, w.w, acc);
                }
                acc = warp_sum(acc);
                const unsigned long long key = ((unsigned long long)fkey32(acc) << 32) | (unsigned)(~r);
                if (lane == q && key > my_best) my_best = key;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = nx[k];
        }
    } else {
        for (int r = r_begin + warp; r < r_end; r += 8) {
            const float* row = db + (size_t)r * d;
            for (int q = 0; q < nq; q++) {
                float acc = 0.f;
                for (int k = lane; k < d; k += 32) acc = fmaf(__ldg(row + k), q_s[(size_t)q * d + k], acc);
                acc = warp_sum(acc);
                const unsigned long long key = ((unsigned long long)fkey32(acc) << 32) | (unsigned)(~r);
                if (lane == q && key > my_best) my_best = key;
            }
        }
    }
    if (lane < nq && my_best) atomicMax(&best_s[lane], my_best);
    __syncthreads();
    if (threadIdx.x < nq && best_s[threadIdx.x] && (!selector || selector[q0 + threadIdx.x] == want))
        atomicMax(&keys[q0 + threadIdx.x], best_s[threadIdx.x]);
}

// loss_face = 1 - <f, target> (E1:1923, 1928), -1 where the row has no target; also the decoded search result
template <typename T>
__global__ void __launch_bounds__(256)
face_loss_fwd_kernel(const float* __restrict__ f, const float* __restrict__ feats_ori, const float* __restrict__ db,
                     const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ mode, int n, int d, float fill,
                     T* __restrict__ loss, long long* __restrict__ target_row, float* __restrict__ similarity) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const int md = mode ? mode[row] : 2;
    long long tr = -1;
    const float* t = nullptr;
    if (md == 1) { t = feats_ori + (size_t)row * d; tr = -2; }
    else if (md == 2 && keys[row]) { tr = (long long)(unsigned)(~(unsigned)(keys[row] & 0xFFFFFFFFull)); t = db + (size_t)tr * d; }
    float dot = 0.f;
    if (t) for (int k = lane; k < d; k += 32) dot = fmaf(f[(size_t)row * d + k], t[k], dot);
    dot = warp_sum(dot);
    if (lane == 0) {
        if (loss) loss[row] = from_f32<T>(t ? 1.f - dot : fill);
        if (target_row) target_row[row] = tr;
        if (similarity) similarity[row] = (md == 2 && keys[row]) ? funkey32((unsigned)(keys[row] >> 32)) : -1.f;
    }
}

// d loss / d raw features: g_f = -g_loss * target, pushed through the normalisation
template <typename T>
__global__ void __launch_bounds__(256)
face_loss_bwd_kernel(const T* __restrict__ g_loss, const float* __restrict__ f, const float* __restrict__ inv_norm,
                     const float* __restrict__ feats_ori, const float* __restrict__ db, const long long* __restrict__ target_row,
                     int n, int d, T* __restrict__ g_raw) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const long long tr = target_row[row];
    T* g = g_raw + (size_t)row * d;
    if (tr == -1) { for (int k = lane; k < d; k += 32) g[k] = from_f32<T>(0.f); return; }
    const float* t = tr == -2 ? feats_ori + (size_t)row * d : db + (size_t)tr * d;
    const float gl = -to_f32(g_loss[row]);
    float dot = 0.f;
    for (int k = lane; k < d; k += 32) dot = fmaf(f[(size_t)row * d + k], t[k], dot);
    dot = warp_sum(dot) * gl;                                          // <f, g_f>
    const float inv = inv_norm[row];
    for (int k = lane; k < d; k += 32) g[k] = from_f32<T>((gl * t[k] - f[(size_t)row * d + k] * dot) * inv);
}

}  // namespace

extern "C" size_t fg_align_params_bytes(int n) { return (size_t)(n > 0 ? n : 1) * ALIGN_PARAMS * sizeof(float); }

extern "C" int fg_align_matrices(const float* landmarks, const uint8_t* indicators, int n, int Hs, int Ws, int Hd, int Wd,
                                 double* M_out, float* params, void* stream) {
    if (n < 0 || Hs < 2 || Ws < 2 || Hd < 1 || Wd < 1) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!landmarks || !params) return FG_ERR_INVALID_ARG;
    align_matrices_kernel<float><<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(landmarks, indicators, n, Hs, Ws, Hd, Wd, M_out, params);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_aligned_warp_fwd(const void* images, int n, int C, int Hs, int Ws, const float* params, const uint8_t* indicators,
                                   void* out, int Hd, int Wd, float fill, int dtype, void* stream) {
    if (n < 0 || C < 1 || Hs < 2 || Ws < 2 || Hd < 1 || Wd < 1) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!images || !params || !out) return FG_ERR_INVALID_ARG;
    if (n > 65535) return FG_ERR_LIMIT;
    dim3 grid((Hd * Wd + 255) / 256, n);
    FG_DISPATCH_DTYPE(dtype, T,
        aligned_warp_fwd_kernel<T><<<grid, 256, 0, fg_stream(stream)>>>((const T*)images, params, indicators, C, Hs, Ws, Hd, Wd, fill, (T*)out));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_aligned_warp_bwd(const void* g_out, int n, int C, int Hd, int Wd, const float* params, const uint8_t* indicators,
                                   void* g_images, int Hs, int Ws, int accumulate, int dtype, void* stream) {
    if (n < 0 || C < 1 || C > 4 || Hs < 2 || Ws < 2 || Hd < 1 || Wd < 1) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!g_out || !params || !g_images) return FG_ERR_INVALID_ARG;
    if (n > 65535) return FG_ERR_LIMIT;
    // cell-map kernel for chips up to 128 x 128 (7-bit row / column in a map entry); FG_AW_GEN1 / FG_AW_SLICES: A/B switches, read once
    static const bool gen1 = getenv("FG_AW_GEN1") != nullptr;
    static const int slices = getenv("FG_AW_SLICES") ? atoi(getenv("FG_AW_SLICES")) : 12;
    if (!gen1 && Hd <= 128 && Wd <= 128) {
        // `slices` CTAs share an image's tiles.  (A persistent grid looping over the images measured SLOWER, 0.82-0.93 ms vs
        // 0.77 ms: faces differ 4x in area and the hardware scheduler balances (image, slice) CTAs better than a static
        // round-robin; FG_AW_GROUPS caps the image groups for such experiments.)
        const int sl = slices > 0 ? slices : 12;
        static const int groups_env = getenv("FG_AW_GROUPS") ? atoi(getenv("FG_AW_GROUPS")) : 0;
        int groups = groups_env > 0 && groups_env < n ? groups_env : n;
        const dim3 grid(sl, groups);
        FG_DISPATCH_DTYPE(dtype, T,
            cudaError_t e = cudaFuncSetAttribute(aligned_warp_bwd_cell_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AM_SMEM);
            if (e != cudaSuccess) return (int)e;
            aligned_warp_bwd_cell_kernel<T><<<grid, 256, AM_SMEM, fg_stream(stream)>>>((const T*)g_out, params, indicators, n, C, Hs, Ws, Hd, Wd, accumulate, (T*)g_images));
    } else {
        const dim3 grid(AW_SLICES, n);
        FG_DISPATCH_DTYPE(dtype, T,
            aligned_warp_bwd_kernel<T><<<grid, 256, 0, fg_stream(stream)>>>((const T*)g_out, params, indicators, n, C, Hs, Ws, Hd, Wd, accumulate, (T*)g_images));
    }
    FG_LAUNCH_CHECK();
    return FG_OK;
}

/* ---- face-realism loss ---- */
extern "C" int fg_feats_normalize_fwd(const void* raw, int n, int d, float* f, float* inv_norm, int dtype, void* stream) {
    if (n < 0 || d < 1) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!raw || !f) return FG_ERR_INVALID_ARG;
    FG_DISPATCH_DTYPE(dtype, T,
        feats_normalize_kernel<T><<<(n + 7) / 8, 256, 0, fg_stream(stream)>>>((const T*)raw, n, d, f, inv_norm, nullptr));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_feats_normalize_bwd(const float* g_f, const float* f, const float* inv_norm, int n, int d, void* g_raw,
                                      int dtype, void* stream) {
    if (n < 0 || d < 1) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!g_f || !f || !inv_norm || !g_raw) return FG_ERR_INVALID_ARG;
    FG_DISPATCH_DTYPE(dtype, T,
        feats_normalize_bwd_kernel<T><<<(n + 7) / 8, 256, 0, fg_stream(stream)>>>(g_f, f, inv_norm, n, d, (T*)g_raw));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

static int launch_search(const float* queries, const uint8_t* selector, uint8_t want, int m, const float* db, int D, int d,
                         unsigned long long* keys, cudaStream_t st, const unsigned* only_if = nullptr) {
    const size_t smem = (size_t)(m < SEARCH_Q ? m : SEARCH_Q) * d * sizeof(float);      // few queries -> more CTAs per SM
    if (smem > 200 * 1024) return FG_ERR_LIMIT;
    cudaError_t e = cudaFuncSetAttribute(face_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int groups = (m + SEARCH_Q - 1) / SEARCH_Q;
    int slabs = (D + 63) / 64;                                          // >= 64 database rows per CTA
    if (slabs > 4 * FG_NUM_SMS) slabs = 4 * FG_NUM_SMS;
    if (slabs < 1) slabs = 1;
    face_search_kernel<<<dim3(slabs, groups), 256, smem, st>>>(queries, selector, want, m, db, D, d, keys, only_if);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

// the exact streaming search as the overflow fallback of the tensor-core search (fg_head.cu): a no-op launch unless *only_if != 0
int fg_internal_search_exact_if(const float* queries, const uint8_t* selector, int m, const float* db, int D, int d,
                                unsigned long long* keys, const unsigned* only_if, cudaStream_t st) {
    return launch_search(queries, selector, 1, m, db, D, d, keys, st, only_if);
}

extern "C" size_t fg_face_search_workspace_bytes(int m) { return (size_t)(m > 0 ? m : 1) * sizeof(unsigned long long); }

extern "C" int fg_face_search_top1(const float* queries, const uint8_t* selector, int m, const float* db, int D, int d,
                                   int64_t* best_row, float* similarity, void* workspace, size_t workspace_bytes, void* stream) {
    if (m < 0 || D < 1 || d < 1 || (d & 3)) return FG_ERR_INVALID_ARG;
    if (m == 0) return FG_OK;
    if (!queries || !db || !best_row) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_face_search_workspace_bytes(m)) return FG_ERR_WORKSPACE;
    cudaStream_t st = fg_stream(stream);
    unsigned long long* keys = (unsigned long long*)workspace;
    cudaError_t e = cudaMemsetAsync(keys, 0, (size_t)m * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    int rc = launch_search(queries, selector, 1, m, db, D, d, keys, st);
    if (rc) return rc;
    face_loss_fwd_kernel<float><<<(m + 7) / 8, 256, 0, st>>>(queries, nullptr, db, keys, nullptr, m, d, -1.f, nullptr,
                                                               (long long*)best_row, similarity);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" size_t fg_face_loss_workspace_bytes(int n, int d) {
    const size_t nn = n > 0 ? n : 1;
    return fg_align_up(nn * d * sizeof(float), 256) + fg_align_up(nn * sizeof(float), 256) + fg_align_up(nn * 8, 256) +
           fg_align_up(nn * 8, 256) + fg_align_up(nn, 256);
}

struct FaceWs { float* f; float* inv_norm; unsigned long long* keys; long long* target_row; uint8_t* mode; };
static FaceWs face_carve(void* base, int n, int d) {
    FaceWs w; char* p = (char*)base; const size_t nn = n > 0 ? n : 1;
    w.f = (float*)p; p += fg_align_up(nn * d * sizeof(float), 256);
    w.inv_norm = (float*)p; p += fg_align_up(nn * sizeof(float), 256);
    w.keys = (unsigned long long*)p; p += fg_align_up(nn * 8, 256);
    w.target_row = (long long*)p; p += fg_align_up(nn * 8, 256);
    w.mode = (uint8_t*)p;
    return w;
}

extern "C" int fg_face_loss_fwd(const void* raw_feats, const float* feats_ori, const float* db, int n, int d, int D,
                                const uint8_t* face_indicators, const int64_t* const* targets, const int64_t* const* preds_ori,
                                const void* const* probs_ori, const int32_t* widths, int n_attr, float confidence_level,
                                int search_needs_target, float fill, void* loss, void* workspace, size_t workspace_bytes,
                                int dtype, void* stream) {
    if (n < 0 || d < 1 || (d & 3) || D < 1 || n_attr < 1 || n_attr > 3) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!raw_feats || !feats_ori || !db || !face_indicators || !targets || !preds_ori || !probs_ori || !widths || !loss) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_face_loss_workspace_bytes(n, d)) return FG_ERR_WORKSPACE;
    FaceWs w = face_carve(workspace, n, d);
    cudaStream_t st = fg_stream(stream);
    const long long* t[3] = {nullptr, nullptr, nullptr}; const long long* p[3] = {nullptr, nullptr, nullptr};
    const void* q[3] = {nullptr, nullptr, nullptr}; int ws[3] = {0, 0, 0};
    for (int a = 0; a < n_attr; a++) {
        if (!targets[a] || !preds_ori[a] || !probs_ori[a] || widths[a] < 1) return FG_ERR_INVALID_ARG;
        t[a] = (const long long*)targets[a]; p[a] = (const long long*)preds_ori[a]; q[a] = probs_ori[a]; ws[a] = widths[a];
    }
    FG_DISPATCH_DTYPE(dtype, T,
        feats_normalize_kernel<T><<<(n + 7) / 8, 256, 0, st>>>((const T*)raw_feats, n, d, w.f, w.inv_norm, w.keys);
        face_mode_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(face_indicators, t[0], t[1], t[2], p[0], p[1], p[2], (const T*)q[0],
                                                            (const T*)q[1], (const T*)q[2], ws[0], ws[1], ws[2], confidence_level,
                                                            search_needs_target, n, w.mode));
    FG_LAUNCH_CHECK();
    int rc = launch_search(w.f, w.mode, 2, n, db, D, d, w.keys, st);
    if (rc) return rc;
    FG_DISPATCH_DTYPE(dtype, T,
        face_loss_fwd_kernel<T><<<(n + 7) / 8, 256, 0, st>>>(w.f, feats_ori, db, w.keys, w.mode, n, d, fill, (T*)loss, w.target_row, nullptr));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_face_loss_bwd(const void* g_loss, const float* feats_ori, const float* db, int n, int d, void* g_raw_feats,
                                void* workspace, size_t workspace_bytes, int dtype, void* stream) {
    if (n < 0 || d < 1) return FG_ERR_INVALID_ARG;
    if (n == 0) return FG_OK;
    if (!g_loss || !feats_ori || !db || !g_raw_feats) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_face_loss_workspace_bytes(n, d)) return FG_ERR_WORKSPACE;
    FaceWs w = face_carve(workspace, n, d);
    FG_DISPATCH_DTYPE(dtype, T,
        face_loss_bwd_kernel<T><<<(n + 7) / 8, 256, 0, fg_stream(stream)>>>((const T*)g_loss, w.f, w.inv_norm, feats_ori, db, w.target_row,
                                                                            n, d, (T*)g_raw_feats));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

/* Test / API hook: the search result the loss used (target_row: >= 0 database row, -2 original features, -1 none). */
extern "C" int fg_face_loss_target_rows(const void* workspace, int n, int d, int64_t* target_row, void* stream) {
    if (n <= 0 || !workspace || !target_row) return FG_ERR_INVALID_ARG;
    FaceWs w = face_carve(const_cast<void*>(workspace), n, d);
    cudaError_t e = cudaMemcpyAsync(target_row, w.target_row, (size_t)n * 8, cudaMemcpyDeviceToDevice, fg_stream(stream));
    return e == cudaSuccess ? FG_OK : (int)e;
}
