// Tiled, shared-memory-staged versions of the two HBM-bound kernels.  Included by fg_sample.cu
// inside its anonymous namespace (uses Axis / axis_index / Box / load_box / candidates / axis_weight).
//
// FORWARD  sample_fwd_tiled_kernel
//   Persistent CTAs, 8 warps: warps 0-6 compute, warp 7 produces.  A tile is TOH = 4 output rows of
//   one job (small_i or chip_i), full output width, all channels.  An output row needs exactly two
//   source rows (i0, i1), so a tile needs 2*TOH*C row segments; the producer warp fetches each with
//   one bulk async copy (cp.async.bulk global->shared, completion on an mbarrier; UBLKCP in SASS)
//   into a STAGES-deep ring, so the HBM reads of tile k+1.. overlap the arithmetic of tile k.  Rows
//   keep their absolute x position in shared memory; pixels outside the image are detected from
//   coordinates and read `fill`.  Tiles are numbered image-major with small_i and chip_i adjacent,
//   so the second pass over an image's pixels hits L2.
//
// BACKWARD  image_grad_tiled_kernel
//   One CTA per (image, 8 source rows), all columns and channels.  Bilinear resampling is separable,
//   so the gather runs in two stages through shared memory:
//     stage 1 (vertical)   t[g][c][r][ox] = sum_oy wy(oy, y_r) * G_g[c][oy][ox]      coalesced G reads
//     stage 2 (horizontal) out[c][y_r][x] = s(x,y) * sum_ox wx_s * t_small + sum_ox wx_c * t_chip
//   with per-CTA tables of (first output index, count, weights) per source x and per source row,
//   built with the exact forward index function.  Every image-gradient element is written once with
//   16-byte stores; no atomics, so the result is run-to-run deterministic.
#pragma once

constexpr int TOH = 4;                      // output rows per forward tile
constexpr int FWD_CONSUMER_WARPS = 7;       // 224 threads = one 224-wide output row per pass
constexpr int FWD_THREADS = (FWD_CONSUMER_WARPS + 1) * 32;

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct FwdParams {
    const void* images; int n, C, H, W;
    const long long* boxes; const uint8_t* ind;
    void* chips; int ch, cw;
    void* small; int sh, sw;
    float fill;
    int tiles_small, tiles_chip;      // row tiles per job
    long long total_tiles;
    int row_bytes;                    // W * sizeof(T), multiple of 16
};

struct FwdTile {
    int img, oy0, oh, ow;
    bool is_small;
    Box b;
};

__device__ __forceinline__ FwdTile decode_tile(const FwdParams& p, long long t) {
    FwdTile f;
    int per_image = p.tiles_small + p.tiles_chip;
    f.img = (int)(t / per_image);
    int r = (int)(t - (long long)f.img * per_image);
    f.is_small = r < p.tiles_small;
    if (f.is_small) {
        f.oy0 = r * TOH; f.oh = p.sh; f.ow = p.sw;
        f.b.x0 = 0; f.b.y0 = 0; f.b.x1 = p.W; f.b.y1 = p.H; f.b.ok = true;
    } else {
        f.oy0 = (r - p.tiles_small) * TOH; f.oh = p.ch; f.ow = p.cw;
        f.b = load_box(p.boxes, p.ind, f.img, p.H, p.W);
    }
    return f;
}

template <typename T, int STAGES>
__global__ void __launch_bounds__(FWD_THREADS)
sample_fwd_tiled_kernel(const FwdParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + STAGES;
    uint8_t* ring = smem + 128;
    const int slots = 2 * TOH * p.C;
    const int stage_bytes = slots * p.row_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], FWD_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const T* images = reinterpret_cast<const T*>(p.images);
    const size_t iplane = (size_t)p.H * p.W;
    constexpr int EPV = 16 / (int)sizeof(T);          // elements per 16 bytes

    if (warp == FWD_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer warp
        int k = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, k++) {
            int s = k % STAGES;
            uint32_t ph = (uint32_t)((k / STAGES) & 1);
            if (k >= STAGES) mbar_wait(&empty[s], ph ^ 1u);
            FwdTile f = decode_tile(p, t);
            uint32_t bytes = 0; const T* src = nullptr; uint8_t* dst = nullptr;
            if (f.b.ok && lane < slots) {
                int c = lane / (2 * TOH), slot = lane - c * 2 * TOH;
                int r = slot >> 1, which = slot & 1;
                int oy = f.oy0 + r;
                if (oy < f.oh) {
                    int bw = f.b.x1 - f.b.x0, bh = f.b.y1 - f.b.y0;
                    Axis ay = axis_index(oy, (float)bh / (float)f.oh, bh);
                    int yy = f.b.y0 + (which ? ay.i1 : ay.i0);
                    Axis a0 = axis_index(0, (float)bw / (float)f.ow, bw);
                    Axis a1 = axis_index(f.ow - 1, (float)bw / (float)f.ow, bw);
                    int xs = f.b.x0 + a0.i0, xe = f.b.x0 + a1.i1 + 1;
                    xs = xs < 0 ? 0 : xs; xe = xe > p.W ? p.W : xe;
                    xs = (xs / EPV) * EPV; xe = ((xe + EPV - 1) / EPV) * EPV;
                    if (xe > p.W) xe = p.W;                     // W*sizeof(T) is a multiple of 16
                    if (yy >= 0 && yy < p.H && xe > xs) {
                        bytes = (uint32_t)(xe - xs) * sizeof(T);
                        src = images + ((size_t)f.img * p.C + c) * iplane + (size_t)yy * p.W + xs;
                        dst = ring + (size_t)s * stage_bytes + (size_t)lane * p.row_bytes + (size_t)xs * sizeof(T);
                    }
                }
            }
            uint32_t total = bytes;
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], total);
            __syncwarp();
            if (bytes) bulk_g2s(dst, src, bytes, &full[s]);
        }
    } else {
        // ------------------------------------------------------------------ consumer warps
        const int ctid = threadIdx.x;                  // 0..223
        int k = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, k++) {
            int s = k % STAGES;
            uint32_t ph = (uint32_t)((k / STAGES) & 1);
            FwdTile f = decode_tile(p, t);
            T* out = reinterpret_cast<T*>(f.is_small ? p.small : p.chips) + (size_t)f.img * p.C * f.oh * f.ow;
            const size_t oplane = (size_t)f.oh * f.ow;
            mbar_wait(&full[s], ph);
            if (!f.b.ok) {
                T fv = from_f32<T>(p.fill);
                for (int ox = ctid; ox < f.ow; ox += FWD_CONSUMER_WARPS * 32)
                    for (int r = 0; r < TOH; r++)
                        if (f.oy0 + r < f.oh)
                            for (int c = 0; c < p.C; c++) out[c * oplane + (size_t)(f.oy0 + r) * f.ow + ox] = fv;
            } else {
                const int bw = f.b.x1 - f.b.x0, bh = f.b.y1 - f.b.y0;
                const float sx = (float)bw / (float)f.ow, sy = (float)bh / (float)f.oh;
                const uint8_t* st = ring + (size_t)s * stage_bytes;
                for (int ox = ctid; ox < f.ow; ox += FWD_CONSUMER_WARPS * 32) {
                    Axis ax = axis_index(ox, sx, bw);
                    int xa = f.b.x0 + ax.i0, xb = f.b.x0 + ax.i1;
                    bool xa_in = xa >= 0 && xa < p.W, xb_in = xb >= 0 && xb < p.W;
#pragma unroll
                    for (int r = 0; r < TOH; r++) {
                        int oy = f.oy0 + r;
                        if (oy >= f.oh) break;
                        Axis ay = axis_index(oy, sy, bh);
                        int ya = f.b.y0 + ay.i0, yb = f.b.y0 + ay.i1;
                        bool ya_in = ya >= 0 && ya < p.H, yb_in = yb >= 0 && yb < p.H;
                        for (int c = 0; c < p.C; c++) {
                            const T* ra = reinterpret_cast<const T*>(st + (size_t)(c * 2 * TOH + 2 * r) * p.row_bytes);
                            const T* rb = reinterpret_cast<const T*>(st + (size_t)(c * 2 * TOH + 2 * r + 1) * p.row_bytes);
                            float v00 = (ya_in && xa_in) ? to_f32(ra[xa]) : p.fill;
                            float v01 = (ya_in && xb_in) ? to_f32(ra[xb]) : p.fill;
                            float v10 = (yb_in && xa_in) ? to_f32(rb[xa]) : p.fill;
                            float v11 = (yb_in && xb_in) ? to_f32(rb[xb]) : p.fill;
                            float top = ax.l0 * v00 + ax.l1 * v01;
                            float bot = ax.l0 * v10 + ax.l1 * v11;
                            out[c * oplane + (size_t)oy * f.ow + ox] = from_f32<T>(ay.l0 * top + ay.l1 * bot);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
}

// ------------------------------------------------------------------------------------- backward
constexpr int BTH = 8;            // source rows per CTA
constexpr int TABW = 3;           // weights kept per table entry; longer runs take the generic path

struct __align__(16) Tab {
    short lo;      // first contributing output index
    short n;       // number of contributing outputs (lo .. lo+n-1); 0 = none
    float w[TABW];
};

// contributions of an output axis (out_size samples of an in_size-long virtual box) to box coordinate p
__device__ __forceinline__ Tab make_tab(int p, float scale, int in_size, int out_size) {
    Tab t; t.lo = 0; t.n = 0; t.w[0] = t.w[1] = t.w[2] = 0.f;
    if (p < 0 || p >= in_size) return t;
    Span s = candidates(p, scale, out_size);
    int first = -1, last = -1;
    for (int o = s.lo; o <= s.hi; o++) {
        float w = axis_weight(o, p, scale, in_size);
        if (w != 0.f) { if (first < 0) first = o; last = o; }
    }
    if (first < 0) return t;
    t.lo = (short)first; t.n = (short)(last - first + 1);
    for (int q = 0; q < TABW && first + q <= last; q++) t.w[q] = axis_weight(first + q, p, scale, in_size);
    return t;
}

// table entries are 16 B; a thread reads V consecutive entries, so the index is XOR-swizzled within
// groups of 8 to keep the 16-byte shared-memory reads of a warp conflict-free
__device__ __forceinline__ int tswz(int e) { return e ^ ((e >> 3) & 7); }

struct BwdParams {
    const void* g_chips; const void* g_small;
    const long long* boxes; const uint8_t* ind;
    const int32_t* region; const float* scale;
    void* g_images;
    int n, C, H, W, ch, cw, sh, sw;
};

// dynamic smem: Tab xtab[2][W]; Tab ytab[2][BTH]; float t[2][C][BTH][OWMAX]
template <typename T>
__global__ void __launch_bounds__(256)
image_grad_tiled_kernel(const BwdParams p, int owmax) {
    extern __shared__ __align__(16) uint8_t smem[];
    Tab* xtab = reinterpret_cast<Tab*>(smem);                 // [2][W]
    Tab* ytab = xtab + 2 * p.W;                                // [2][BTH]
    float* tb = reinterpret_cast<float*>(ytab + 2 * BTH);      // [2][C][BTH][owmax]
    const int img = blockIdx.y;
    const int y0 = blockIdx.x * BTH;
    const int tid = threadIdx.x;
    const bool has_s = p.g_small != nullptr;
    Box b; b.ok = false;
    if (p.g_chips) b = load_box(p.boxes, p.ind, img, p.H, p.W);
    // tile rows that can receive chip gradient at all
    const bool chip_rows = b.ok && y0 < b.y1 && y0 + BTH > b.y0;
    const int bw = b.x1 - b.x0, bh = b.y1 - b.y0;
    const float ssx = (float)p.W / (float)p.sw, ssy = (float)p.H / (float)p.sh;
    const float csx = chip_rows ? (float)bw / (float)p.cw : 1.f, csy = chip_rows ? (float)bh / (float)p.ch : 1.f;

    for (int x = tid; x < p.W; x += 256) {
        if (has_s) xtab[tswz(x)] = make_tab(x, ssx, p.W, p.sw);
        if (chip_rows) xtab[p.W + tswz(x)] = make_tab(x - b.x0, csx, bw, p.cw);
    }
    if (tid < BTH) {
        int y = y0 + tid;
        if (has_s) ytab[tid] = make_tab(y < p.H ? y : -1, ssy, p.H, p.sh);
        if (chip_rows) ytab[BTH + tid] = make_tab(y < p.H ? y - b.y0 : -1, csy, bh, p.ch);
    }
    __syncthreads();

    // ---- stage 1: vertical pass, coalesced along ox
    for (int g = 0; g < 2; g++) {
        if (g == 0 ? !has_s : !chip_rows) continue;
        const int oh = g == 0 ? p.sh : p.ch, ow = g == 0 ? p.sw : p.cw;
        const T* G = reinterpret_cast<const T*>(g == 0 ? p.g_small : p.g_chips) + (size_t)img * p.C * oh * ow;
        const float scl = g == 0 ? ssy : csy;
        const int in_size = g == 0 ? p.H : bh;
        const int items = BTH * ow;
        for (int it = tid; it < items; it += 256) {
            int r = it / ow, ox = it - r * ow;
            Tab ty = ytab[g * BTH + r];
            float acc[FG_MAXC] = {0.f, 0.f, 0.f, 0.f};
            if (ty.n > 0) {
                if (ty.n <= TABW) {
                    for (int q = 0; q < ty.n; q++) {
                        const T* row = G + (size_t)(ty.lo + q) * ow + ox;
#pragma unroll
                        for (int c = 0; c < FG_MAXC; c++) if (c < p.C) acc[c] += ty.w[q] * to_f32(row[(size_t)c * oh * ow]);
                    }
                } else {
                    int py = (y0 + r) - (g == 0 ? 0 : b.y0);
                    for (int q = 0; q < ty.n; q++) {
                        float w = axis_weight(ty.lo + q, py, scl, in_size);
                        const T* row = G + (size_t)(ty.lo + q) * ow + ox;
#pragma unroll
                        for (int c = 0; c < FG_MAXC; c++) if (c < p.C) acc[c] += w * to_f32(row[(size_t)c * oh * ow]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < FG_MAXC; c++) if (c < p.C) tb[((size_t)(g * p.C + c) * BTH + r) * owmax + ox] = acc[c];
        }
    }
    __syncthreads();

    // ---- stage 2: horizontal pass, V consecutive x per thread, 16-byte stores
    constexpr int V = 16 / (int)sizeof(T);
    int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0; float rs = 1.f;
    if (p.region) {
        rx0 = p.region[4 * img]; ry0 = p.region[4 * img + 1]; rx1 = p.region[4 * img + 2]; ry1 = p.region[4 * img + 3];
        rs = p.scale[img];
    }
    const int vec_per_row = p.W / V;                   // host guarantees W % V == 0
    T* gout = reinterpret_cast<T*>(p.g_images) + (size_t)img * p.C * p.H * p.W;
    for (int it = tid; it < BTH * vec_per_row; it += 256) {
        int r = it / vec_per_row, xv = (it - r * vec_per_row) * V;
        int y = y0 + r;
        if (y >= p.H) continue;
        const bool row_in_region = y >= ry0 && y < ry1;
        float acc[FG_MAXC][V];
#pragma unroll
        for (int c = 0; c < FG_MAXC; c++)
#pragma unroll
            for (int v = 0; v < V; v++) acc[c][v] = 0.f;
#pragma unroll
        for (int v = 0; v < V; v++) {
            int x = xv + v;
            if (has_s) {
                Tab tx = xtab[tswz(x)];
                if (tx.n > 0) {
                    float s = (row_in_region && x >= rx0 && x < rx1) ? rs : 1.f;
                    const float* tr = tb + (size_t)r * owmax + tx.lo;
                    if (tx.n <= TABW) {
                        for (int q = 0; q < tx.n; q++) {
                            float w = tx.w[q] * s;
#pragma unroll
                            for (int c = 0; c < FG_MAXC; c++) if (c < p.C) acc[c][v] += w * tr[(size_t)c * BTH * owmax + q];
                        }
                    } else {
                        for (int q = 0; q < tx.n; q++) {
                            float w = axis_weight(tx.lo + q, x, ssx, p.W) * s;
#pragma unroll
                            for (int c = 0; c < FG_MAXC; c++) if (c < p.C) acc[c][v] += w * tr[(size_t)c * BTH * owmax + q];
                        }
                    }
                }
            }
            if (chip_rows) {
                Tab tx = xtab[p.W + tswz(x)];
                if (tx.n > 0) {
                    const float* tr = tb + ((size_t)p.C * BTH + r) * owmax + tx.lo;
                    if (tx.n <= TABW) {
                        for (int q = 0; q < tx.n; q++) {
                            float w = tx.w[q];
#pragma unroll
                            for (int c = 0; c < FG_MAXC; c++) if (c < p.C) acc[c][v] += w * tr[(size_t)c * BTH * owmax + q];
                        }
                    } else {
                        for (int q = 0; q < tx.n; q++) {
                            float w = axis_weight(tx.lo + q, x - b.x0, csx, bw);
#pragma unroll
                            for (int c = 0; c < FG_MAXC; c++) if (c < p.C) acc[c][v] += w * tr[(size_t)c * BTH * owmax + q];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < FG_MAXC; c++) {
            if (c >= p.C) break;
            T packed[V];
#pragma unroll
            for (int v = 0; v < V; v++) packed[v] = from_f32<T>(acc[c][v]);
            *reinterpret_cast<int4*>(gout + ((size_t)c * p.H + y) * p.W + xv) = *reinterpret_cast<const int4*>(packed);
        }
    }
}
