// BACKWARD for 16-bit gradients at the BASELINE shape (512x512 images, 224x224 chips and resized images).
// Included by fg_sample.cu after fg_sample_tiled.cuh (uses Tab / make_tab / the mbarrier + bulk-copy wrappers).
//
// image_grad_tiled_kernel gathers g_small and g_chip straight from global memory; its 2-byte gathers are the
// only thing a thread waits on and there are too few of them in flight (latency-bound at ~1/3 of HBM speed).
// Here the gradient rows a sub-tile needs are CONTIGUOUS per channel (rows [first, first+count) at full width),
// so one thread fetches them with three bulk async copies per grid (cp.async.bulk -> mbarrier), one sub-tile
// ahead of the arithmetic; every gather then reads shared memory.
//
//   CTA = (image, NSUB sub-tiles of 8 image rows), 256 threads, a thread owns two adjacent image columns.
//   per sub-tile:  [wait chip rows] stage 1: vertical pass  sC -> tb     (threads 0..223, one chip column each)
//                  barrier; thread 0 issues the copies of the next sub-tile
//                  [wait small rows] stage 2: horizontal pass  bufS, tb -> g_images (one 4-byte store per channel)
//                  barrier
//   bufS is double-buffered (stage 2 of sub-tile k reads it while k+1 arrives), sC is single-buffered (free
//   once stage 1 is done).  The 512 -> 224 resize shrinks by more than 2x, so an image pixel receives at
//   most ONE resized pixel per axis: the small branch is a scaled gather with no vertical pass.
//   Rows without a tap point at a zero row, columns without a tap are zeroed by a select: a non-finite
//   gradient never leaks into pixels it does not touch.
//   Boxes the staging buffer cannot hold (more than GS_CROWS chip rows per 8 image rows, i.e. boxes under
//   ~115 px, or more than 4 taps per pixel) take the direct 2-D gather per pixel (correct, slow, rare).
#pragma once

constexpr int GS_ROWS = 8;        // image rows per sub-tile
constexpr int GS_SROWS = 6;       // resized rows a sub-tile can touch (<= 5 at 512 -> 224) ; row GS_SROWS is the zero row
constexpr int GS_CROWS = 16;      // chip rows staged per sub-tile
constexpr int GS_OW = 224;        // width (and height) of both gradient grids
constexpr int GS_TBW = GS_OW + TPAD;

struct GsRowS { int off; float wy; };              // element offset of the resized row inside a bufS channel, y weight
struct GsRowC { int off; int n; float w[TABW]; };  // element offset of the first chip row inside an sC channel, taps
struct GsSub { int s_first, s_count, c_first, c_count; };

template <int NSUB> struct GsLayout {
    static constexpr int ROWS = NSUB * GS_ROWS;
    static constexpr size_t bars = 0;                                      // 3 mbarriers
    static constexpr size_t subs = 32;
    static constexpr size_t rowS = subs + NSUB * sizeof(GsSub);
    static constexpr size_t rowC = rowS + ROWS * sizeof(GsRowS);
    static constexpr size_t tabs_end = rowC + ROWS * sizeof(GsRowC);
    static constexpr size_t bufS = (tabs_end + 127) / 128 * 128;
    static constexpr size_t bufS_bytes = 2 * 3 * (GS_SROWS + 1) * GS_OW * 2;
    static constexpr size_t sC = bufS + bufS_bytes;
    static constexpr size_t sC_bytes = 3 * GS_CROWS * GS_OW * 2;
    static constexpr size_t tb = sC + sC_bytes;
    static constexpr size_t total = tb + 3 * GS_ROWS * GS_TBW * sizeof(float);
};

template <typename T, int NSUB>
__global__ void __launch_bounds__(256, 3)
image_grad_staged_kernel(const BwdParams p) {
    static_assert(sizeof(T) == 2, "16-bit gradients only");
    using L = GsLayout<NSUB>;
    constexpr int C = 3, H = 512, W = 512, OH = GS_OW, OW = GS_OW, ROWS = NSUB * GS_ROWS;
    constexpr int S_CH = (GS_SROWS + 1) * OW;         // bufS channel stride (elements)
    constexpr int C_CH = GS_CROWS * OW;               // sC channel stride
    constexpr int T_CH = GS_ROWS * GS_TBW;            // tb channel stride
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
    GsSub* subs = reinterpret_cast<GsSub*>(smem + L::subs);
    GsRowS* rowS = reinterpret_cast<GsRowS*>(smem + L::rowS);
    GsRowC* rowC = reinterpret_cast<GsRowC*>(smem + L::rowC);
    T* bufS = reinterpret_cast<T*>(smem + L::bufS);   // [2][C][GS_SROWS + 1][OW]
    T* sC = reinterpret_cast<T*>(smem + L::sC);       // [C][GS_CROWS][OW]
    float* tb = reinterpret_cast<float*>(smem + L::tb);   // [C][GS_ROWS][GS_TBW]

    const int img = blockIdx.y, ybase = blockIdx.x * ROWS, tid = threadIdx.x;
    const bool has_s = p.g_small != nullptr;
    Box b; b.ok = false; b.x0 = b.y0 = b.x1 = b.y1 = 0;
    if (p.g_chips) b = load_box(p.boxes, p.ind, img, H, W);
    const int bw = b.x1 - b.x0, bh = b.y1 - b.y0;
    const float ss = (float)W / (float)OW;
    const float csx = b.ok ? (float)bw / (float)OW : 1.f, csy = b.ok ? (float)bh / (float)OH : 1.f;
    bool chip_cold = b.ok && (csx < 0.51f || csy < 0.51f);
    const bool chip_tab = b.ok && !chip_cold;

    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero rows of bufS and the pad columns of tb (read with zero weights, so they must stay finite)
    for (int e = tid; e < 2 * C * OW; e += 256) {
        const int bc = e / OW, x = e - bc * OW;
        bufS[bc * S_CH + GS_SROWS * OW + x] = from_f32<T>(0.f);
    }
    for (int e = tid; e < C * GS_ROWS * TPAD; e += 256) {
        const int cr = e / TPAD, q = e - cr * TPAD;
        tb[cr * GS_TBW + OW + q] = 0.f;
    }
    int too_many = 0;
    for (int e = tid; e < 2 * ROWS; e += 256) {
        const int g = e / ROWS, r = e - g * ROWS;
        const int y = ybase + r;
        if (g == 0) {
            Tab t; t.lo = 0; t.n = 0; t.w[0] = 0.f;
            if (has_s) t = make_tab(y, ss, H, OH);
            GsRowS rs_; rs_.off = t.n ? t.lo : -1; rs_.wy = t.n ? t.w[0] : 0.f;       // absolute row for now
            rowS[r] = rs_;
        } else {
            Tab t; t.lo = 0; t.n = 0; t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
            if (chip_tab) t = make_tab(y - b.y0, csy, bh, OH);
            if (t.n > TABW) too_many = 1;
            GsRowC rc; rc.off = t.lo; rc.n = t.n;
#pragma unroll
            for (int q = 0; q < TABW; q++) rc.w[q] = t.w[q];
            rowC[r] = rc;
        }
    }
    // this thread's two image columns
    const int x_a = 2 * tid;
    int xs_lo[2]; float xs_w[2]; bool xs_k[2];
    Tab xc[2];
    bool in_reg_x[2];
    int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0; float rs = 1.f;
    if (p.region) {
        rx0 = p.region[4 * img]; ry0 = p.region[4 * img + 1]; rx1 = p.region[4 * img + 2]; ry1 = p.region[4 * img + 3];
        rs = p.scale[img];
    }
#pragma unroll
    for (int v = 0; v < 2; v++) {
        const int x = x_a + v;
        Tab t; t.lo = 0; t.n = 0; t.w[0] = 0.f;
        if (has_s) t = make_tab(x, ss, W, OW);
        xs_k[v] = t.n != 0; xs_lo[v] = t.n ? t.lo : 0; xs_w[v] = t.n ? t.w[0] : 0.f;
        xc[v].lo = 0; xc[v].n = 0;
#pragma unroll
        for (int q = 0; q < TABW; q++) xc[v].w[q] = 0.f;
        if (chip_tab) xc[v] = make_tab(x - b.x0, csx, bw, OW);
        if (xc[v].n > TABW) too_many = 1;
        if (xc[v].n == 0) xc[v].lo = 0;
        in_reg_x[v] = x >= rx0 && x < rx1;
    }
    if (__syncthreads_or(too_many)) chip_cold = true;               // also orders the table writes
    // per sub-tile: which rows to stage; turn absolute rows into offsets inside the staging buffers
    too_many = 0;
    if (tid < NSUB) {
        int s_lo = 1 << 30, s_hi = -1, c_lo = 1 << 30, c_hi = -1;
        for (int r = 0; r < GS_ROWS; r++) {
            const GsRowS a = rowS[tid * GS_ROWS + r];
            if (a.off >= 0) { s_lo = min(s_lo, a.off); s_hi = max(s_hi, a.off); }
            const GsRowC c = rowC[tid * GS_ROWS + r];
            if (c.n > 0) { c_lo = min(c_lo, c.off); c_hi = max(c_hi, min(c.off + c.n - 1, OH - 1)); }
        }
        GsSub d;
        d.s_first = s_hi >= 0 ? s_lo : 0; d.s_count = s_hi >= 0 ? min(s_hi - s_lo + 1, GS_SROWS) : 0;
        d.c_first = c_hi >= 0 ? c_lo : 0; d.c_count = c_hi >= 0 ? c_hi - c_lo + 1 : 0;
        if (d.c_count > GS_CROWS) too_many = 1;
        if (chip_cold) d.c_count = 0;
        subs[tid] = d;
        for (int r = 0; r < GS_ROWS; r++) {
            GsRowS& a = rowS[tid * GS_ROWS + r];
            a.off = a.off >= 0 ? (a.off - d.s_first) * OW : GS_SROWS * OW;
            GsRowC& c = rowC[tid * GS_ROWS + r];
            c.off = (c.n > 0 ? c.off - d.c_first : 0) * OW;
        }
    }
    if (__syncthreads_or(too_many)) chip_cold = true;
    const bool chip_fast = b.ok && !chip_cold;

    const T* gs_img = reinterpret_cast<const T*>(p.g_small) + (size_t)img * C * OH * OW;
    const T* gc_img = reinterpret_cast<const T*>(p.g_chips) + (size_t)img * C * OH * OW;
    auto issue = [&](int sub) {        // thread 0 only
        const GsSub d = subs[sub];
        if (has_s && d.s_count > 0) {
            uint64_t* bar = &bars[sub & 1];
            const uint32_t bytes = (uint32_t)d.s_count * OW * (uint32_t)sizeof(T);
            mbar_arrive_expect_tx(bar, C * bytes);
#pragma unroll
            for (int c = 0; c < C; c++)
                bulk_g2s(bufS + ((sub & 1) * C + c) * S_CH, gs_img + c * OH * OW + d.s_first * OW, bytes, bar);
        }
        if (chip_fast && d.c_count > 0) {
            const uint32_t bytes = (uint32_t)d.c_count * OW * (uint32_t)sizeof(T);
            mbar_arrive_expect_tx(&bars[2], C * bytes);
#pragma unroll
            for (int c = 0; c < C; c++)
                bulk_g2s(sC + c * C_CH, gc_img + c * OH * OW + d.c_first * OW, bytes, &bars[2]);
        }
    };
    if (tid == 0) issue(0);

    const float wxs0 = xs_w[0], wxs1 = xs_w[1];
    const float wxr0 = in_reg_x[0] ? wxs0 * rs : wxs0, wxr1 = in_reg_x[1] ? wxs1 * rs : wxs1;
    const bool k0 = xs_k[0], k1 = xs_k[1];
    const bool c_wide = __syncthreads_or((xc[0].n > 2) | (xc[1].n > 2)) != 0;
    const bool warp_has_chip = __any_sync(0xffffffffu, (xc[0].n | xc[1].n) != 0);
    const float* tc0 = tb + xc[0].lo;
    const float* tc1 = tb + xc[1].lo;
    char* go_img = reinterpret_cast<char*>(p.g_images) + (size_t)img * C * H * W * sizeof(T) + (size_t)x_a * sizeof(T);
    constexpr unsigned oplb = (unsigned)(H * W * sizeof(T)), orowb = (unsigned)(W * sizeof(T));
    uint32_t phS = 0u, phC = 0u;          // mbarrier phase parities (bit k of phS: bufS[k])

    for (int sub = 0; sub < NSUB; sub++) {
        const int y0 = ybase + sub * GS_ROWS;
        const GsSub d = subs[sub];
        const bool chip_rows = chip_fast && d.c_count > 0;
        // ---- stage 1: vertical pass over the staged chip rows; a thread owns one chip column
        if (chip_rows) {
            mbar_wait(&bars[2], phC); phC ^= 1u;
            if (tid < OW) {
                const T* col = sC + tid;
                float* tcol = tb + tid;
#pragma unroll 4
                for (int r = 0; r < GS_ROWS; r++) {
                    const GsRowC rc = rowC[sub * GS_ROWS + r];               // warp-uniform
                    float acc[C] = {0.f, 0.f, 0.f};
                    const T* base = col + rc.off;
#pragma unroll
                    for (int q = 0; q < TABW; q++) {
                        if (q < rc.n) {
#pragma unroll
                            for (int c = 0; c < C; c++) acc[c] += rc.w[q] * to_f32(base[c * C_CH + q * OW]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < C; c++) tcol[c * T_CH + r * GS_TBW] = acc[c];
                }
            }
            __syncthreads();
        }
        if (tid == 0 && sub + 1 < NSUB) issue(sub + 1);

        // ---- stage 2: horizontal pass for this thread's two columns
        const bool wait_s = has_s && d.s_count > 0;
        if (wait_s) { mbar_wait(&bars[sub & 1], (phS >> (sub & 1)) & 1u); phS ^= 1u << (sub & 1); }
        const bool do_chip = chip_rows && warp_has_chip;
        const T* sbuf = bufS + (sub & 1) * C * S_CH;
        const T* s0 = sbuf + xs_lo[0];
        const T* s1 = sbuf + xs_lo[1];
FG_UNROLL(BWD_UNROLL)
        for (int r = 0; r < GS_ROWS; r++) {
            const int y = y0 + r;
            float o0[C], o1[C];
            if (wait_s) {
                const GsRowS rr = rowS[sub * GS_ROWS + r];                   // warp-uniform
                const bool row_reg = y >= ry0 && y < ry1;
                const float f0 = rr.wy * (row_reg ? wxr0 : wxs0), f1 = rr.wy * (row_reg ? wxr1 : wxs1);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    o0[c] = k0 ? f0 * to_f32(s0[rr.off + c * S_CH]) : 0.f;
                    o1[c] = k1 ? f1 * to_f32(s1[rr.off + c * S_CH]) : 0.f;
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) { o0[c] = 0.f; o1[c] = 0.f; }
            }
            if (do_chip) {
                const float* t0 = tc0 + r * GS_TBW;
                const float* t1 = tc1 + r * GS_TBW;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    o0[c] += xc[0].w[0] * t0[c * T_CH] + xc[0].w[1] * t0[c * T_CH + 1];
                    o1[c] += xc[1].w[0] * t1[c * T_CH] + xc[1].w[1] * t1[c * T_CH + 1];
                }
                if (c_wide) {
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        o0[c] += xc[0].w[2] * t0[c * T_CH + 2] + xc[0].w[3] * t0[c * T_CH + 3];
                        o1[c] += xc[1].w[2] * t1[c * T_CH + 2] + xc[1].w[3] * t1[c * T_CH + 3];
                    }
                }
            }
            if (chip_cold) {
                // rare: tiny box, direct 2-D gather from global memory (the generic kernel's routine)
#pragma unroll
                for (int v = 0; v < 2; v++) {
                    const int x = x_a + v;
                    if (x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1) {
                        float acc[FG_MAXC] = {0.f, 0.f, 0.f, 0.f};
                        gather_grid_cold<T>(gc_img, C, OH, OW, x - b.x0, y - b.y0, bw, bh, acc);
#pragma unroll
                        for (int c = 0; c < C; c++) { if (v == 0) o0[c] += acc[c]; else o1[c] += acc[c]; }
                    }
                }
            }
            char* orow = go_img + (unsigned)y * orowb;                       // warp-uniform + lane offset
#pragma unroll
            for (int c = 0; c < C; c++) {
                using P2 = Pack2<T>;
                *reinterpret_cast<typename P2::type*>(orow + c * oplb) = P2::make(o0[c], o1[c]);
            }
        }
        __syncthreads();
    }
}
