// BACKWARD, second generation, at the BASELINE shape (512x512 images, 224x224 chips and resized images), 16-bit and fp32 gradients.
// Included by fg_sample.cu after fg_image_grad_staged.cuh (shares Tab / make_tab / the mbarrier + bulk-copy wrappers).
//
// image_grad_staged_kernel (first generation) is bound by the L1 data pipe, not by HBM: ncu (profiles/r01_ncu_summary.txt)
// shows 98.7 M shared-memory wavefronts for 2.2 GB of traffic, because every tap is a 2-byte ld.shared (half-empty
// wavefronts), every row of every warp re-reads a 16-byte record, and adjacent image rows re-read the same resized-gradient
// row.  This kernel keeps the decomposition (CTA = image x band of rows, gradient rows prefetched one 8-row sub-tile ahead
// with cp.async.bulk -> mbarrier, vertical pass of the chip gradient through shared memory, every image-gradient element
// written exactly once, no atomics) and removes those three costs:
//
//   * 512 -> 224 is the fixed ratio 16 : 7, so WHICH resized row / column lands on which image row / column is static:
//     resized index o touches image indices i0(o) = floor((32 o + 9) / 14) and i0(o) + 1, i.e. within every block of 16 the
//     pairs (0,1) (2,3) (5,6) (7,8) (9,10) (12,13) (14,15), and indices 4 and 11 receive nothing.  (The distance of
//     (32 o + 9) / 14 from an integer is at least 1/14, so the fp32 rounding of ATen's index formula cannot move i0.)  The
//     WEIGHTS are still the fp32 values of that formula (axis_index), computed per thread / per CTA.
//   * a thread owns FOUR adjacent image columns (one 8-byte store per row and channel; a warp writes 256 contiguous bytes).
//     Its four columns take their resized-gradient values from exactly two adjacent resized columns: one or two 32-bit
//     shared loads of packed 16-bit pairs + one byte permute, shared by the TWO image rows a resized row feeds.
//   * the vertical pass owns two adjacent chip columns per thread: one 32-bit shared load per tap and one 8-byte store.
//   * per-row records are read in the vertical pass only; the horizontal pass gets its row weights from a 7-entry table
//     per 16 rows and its unit schedule from constant memory (uniform loads).
//
//   CTA = (image, NSUB sub-tiles of 8 image rows), 256 threads.
//   vertical pass:   warps 0-3 rows 0-3, warps 4-7 rows 4-7 of the sub-tile; lane < 28 owns chip columns 2p, 2p+1
//   horizontal pass: threads 0-127 / 128-255 ("halves") take alternate units of the sub-tile's static schedule; a unit is
//                    the pair of image rows fed by one resized row, or a single row (rows 4 and 11 of a block of 16, and
//                    rows 7 / 8 whose partner lies in the other sub-tile); both halves process 4 rows per sub-tile.
#pragma once

#ifndef GQ_CUNROLL
#define GQ_CUNROLL 3              // unroll factor of the channel loop of the horizontal pass
#endif
#ifndef GQ_INTERLEAVE
#define GQ_INTERLEAVE 0           // 1: a warp owns two 64-column blocks 256 columns apart; 0: one 128-column block
#endif
#ifndef GQ_NSC
#define GQ_NSC 1                  // chip-row staging buffers for 16-bit gradients
#endif
constexpr int GQ_ROWS = 8;              // image rows per sub-tile
constexpr int GQ_SROWS = 4;             // resized rows a sub-tile touches (static: o = 7g..7g+3 for rows 0-7, 7g+3..7g+6 for rows 8-15)
constexpr int GQ_CROWS = 16;            // chip rows staged per sub-tile
constexpr int GQ_OW = 224;
constexpr int GQ_TBW = GQ_OW + TPAD;
constexpr int GQ_T_CH = GQ_ROWS * GQ_TBW * 4;            // tb channel stride, bytes
template <typename T> struct GqT {
    static constexpr int ES = (int)sizeof(T);
    static constexpr int ROWB = GQ_OW * ES;              // bytes of one staged gradient row
    static constexpr int S_CH = GQ_SROWS * ROWB;         // bufS channel stride, bytes
    static constexpr int C_CH = GQ_CROWS * ROWB;         // sC channel stride, bytes
};

struct GqSub { int c_first, c_count, pad0, pad1; };
struct __align__(16) GqRowC { int off; int n; float w[TABW]; int pad[2]; };

template <typename T, int NSUB> struct GqLayout {
    static constexpr int ROWS = NSUB * GQ_ROWS;
    static constexpr int NSC = sizeof(T) == 2 ? GQ_NSC : 1;                  // chip-row staging buffers (16-bit: double-buffered)
    static constexpr size_t bars = 0;                                        // 4 mbarriers: bufS[0], bufS[1], sC[0], sC[1]
    static constexpr size_t subs = 32;
    static constexpr size_t wy = subs + NSUB * sizeof(GqSub);                // float2 (l0, l1) per resized row of the band
    static constexpr size_t rowC = (wy + (ROWS / 16) * 7 * 8 + 15) / 16 * 16;    // 16-byte records
    static constexpr size_t w23 = rowC + ROWS * sizeof(GqRowC);              // float2 (w2, w3) per image column
    static constexpr size_t tabs_end = w23 + 512 * 8;
    static constexpr size_t bufS = (tabs_end + 127) / 128 * 128;
    static constexpr size_t sC = bufS + 2 * 3 * GqT<T>::S_CH;
    static constexpr size_t tb = sC + NSC * 3 * GqT<T>::C_CH;
    static constexpr size_t total = tb + 3 * GQ_T_CH;
};

// Unit schedule of the horizontal pass.  A unit is the pair of image rows fed by one resized row, or a single row.  Per
// sub-tile parity (rows 0-7 / rows 8-15 of a block of 16) the five units are split so that both halves do 4 rows AND have a
// fixed structure (the code of a half is straight-line; only row / resized-row offsets and one tap choice are run-time values):
//   half 1 "PP" : two pairs                         rows 0-7: (2,3)<-o1, (5,6)<-o2      rows 8-15: (9,10)<-o4, (12,13)<-o5
//   half 0 "PZS": a pair, a row the resize misses,  rows 0-7: (0,1)<-o0, 4, 7<-o3 (l0)  rows 8-15: (14,15)<-o6, 11, 8<-o3 (l1)
//                 a single row
// rows are relative to the sub-tile, resized rows relative to the staged block of 4 (o0..o3 resp. o3..o6)
struct GqSched { int row[3]; int srow[3]; int tap; };
__constant__ GqSched gq_sched[2][2] = {
    {{{0, 4, 7}, {0, -1, 3}, 0}, {{2, 5, 0}, {1, 2, 0}, 0}},
    {{{6, 3, 0}, {3, -1, 0}, 1}, {{1, 4, 0}, {1, 2, 0}, 0}},
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_v2f32(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_v2f32(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y)); }

template <typename T> __device__ __forceinline__ void unpack_pair(uint32_t w, float& lo, float& hi);
template <> __device__ __forceinline__ void unpack_pair<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack_pair<__half>(uint32_t w, float& lo, float& hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w)); lo = f.x; hi = f.y;
}
template <> __device__ __forceinline__ void unpack_pair<float>(uint32_t, float&, float&) {}          // 16-bit paths only
template <typename T> __device__ __forceinline__ uint32_t pack_pair(float a, float b);
template <> __device__ __forceinline__ uint32_t pack_pair<float>(float, float) { return 0u; }
template <> __device__ __forceinline__ uint32_t pack_pair<__nv_bfloat16>(float a, float b) { const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }
template <> __device__ __forceinline__ uint32_t pack_pair<__half>(float a, float b) { const __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }

// two adjacent elements at a (2-element aligned) shared address, widened to fp32
template <typename T> __device__ __forceinline__ void load_pair(uint32_t a, float& lo, float& hi) {
    if (sizeof(T) == 4) { const float2 f = lds_v2f32(a); lo = f.x; hi = f.y; }
    else unpack_pair<T>(lds_u32(a), lo, hi);
}
// the two resized columns (A, B) of a quad: 16-bit: two words + byte permute; fp32: the two words themselves
template <typename T> __device__ __forceinline__ void load_ab(uint32_t a, uint32_t perm, float& vA, float& vB) {
    if (sizeof(T) == 4) { vA = lds_f32(a); vB = lds_f32(a + 4); }
    else unpack_pair<T>(__byte_perm(lds_u32(a), lds_u32(a + 4), perm), vA, vB);
}
template <typename T> __device__ __forceinline__ void store_quad(char* ptr, float a, float b, float c, float d) {
    if (sizeof(T) == 4) *reinterpret_cast<float4*>(ptr) = make_float4(a, b, c, d);
    else *reinterpret_cast<uint2*>(ptr) = make_uint2(pack_pair<T>(a, b), pack_pair<T>(c, d));
}

// vertical pass over the staged chip rows for 4 image rows: tb[c][r][2p .. 2p+1] = sum_q w[r][q] * sC[c][off_r + q][2p .. 2p+1]
template <typename T>
__device__ __forceinline__ void gq_stage1(uint32_t rec, uint32_t src, uint32_t dst) {
    constexpr int C_CH = GqT<T>::C_CH, ROWB = GqT<T>::ROWB;
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        const uint4 h = lds_v4(rec + rr * 32);                // off, n, w0, w1      (warp-uniform)
        float a0[3] = {0.f, 0.f, 0.f}, a1[3] = {0.f, 0.f, 0.f};
        const int n = (int)h.y;
        if (n > 0) {
            const uint32_t a = src + h.x;
            const float w0 = __uint_as_float(h.z), w1 = __uint_as_float(h.w);
            float v0[3], v1[3];
#pragma unroll
            for (int c = 0; c < 3; c++) load_pair<T>(a + c * C_CH, v0[c], v1[c]);
            if (n > 1) {
                float u0[3], u1[3];
#pragma unroll
                for (int c = 0; c < 3; c++) load_pair<T>(a + c * C_CH + ROWB, u0[c], u1[c]);
#pragma unroll
                for (int c = 0; c < 3; c++) { a0[c] = w1 * u0[c]; a1[c] = w1 * u1[c]; }
                if (n > 2) {
                    const uint4 h2 = lds_v4(rec + rr * 32 + 16);   // w2, w3
                    const float w2 = __uint_as_float(h2.x), w3 = __uint_as_float(h2.y);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float lo, hi; load_pair<T>(a + c * C_CH + 2 * ROWB, lo, hi);
                        a0[c] = fmaf(w2, lo, a0[c]); a1[c] = fmaf(w2, hi, a1[c]);
                    }
                    if (n > 3) {
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            float lo, hi; load_pair<T>(a + c * C_CH + 3 * ROWB, lo, hi);
                            a0[c] = fmaf(w3, lo, a0[c]); a1[c] = fmaf(w3, hi, a1[c]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 3; c++) { a0[c] = fmaf(w0, v0[c], a0[c]); a1[c] = fmaf(w0, v1[c], a1[c]); }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) sts_v2f32(dst + c * GQ_T_CH + rr * GQ_TBW * 4, a0[c], a1[c]);
    }
}

struct GqCols {             // per-thread constants of its four image columns
    // resized-gradient branch: byte offset of the first of the two 32-bit words inside a bufS row and the byte-permute selector
    // that brings the two resized columns (A, B) of this quad into one word.  Which of A / B / nothing a column takes depends on
    // the quad's position j inside its block of 16 columns only: j = 0, 3: A A B B;  j = 1: - A A B;  j = 2: A B B -
    uint32_t s_off, s_perm;
    bool j1, j2;
    float wx[4];                // x weight of the resized-gradient tap
    float rs;                   // hook factor of the scaled region; rin[i]: column i lies inside its x range
    bool rin[4];
    // chip branch, per column pair: shared address of tb[0][0][lo], start distance of the second column (0 | 1), the first two
    // tap weights of every column (taps 2 and 3 of upscaled boxes live in a per-column shared-memory table, w23)
    uint32_t t[2][2];           // [pair][column] address of tb[0][0][lo of the column]
    bool d[2];
    float w[4][2];
    uint32_t w23;               // shared address of this thread's four (w2, w3) pairs
};

// chip contribution of one image row to this thread's four columns.  CHIP 1: every column has <= 2 taps and the two columns of a
// pair start 0 or 1 apart (three shared words cover a pair); CHIP 2: up to 4 taps per column.  FIRST: `o` holds nothing yet.
template <int CHIP, bool FIRST>
__device__ __forceinline__ void gq_chip_row(const GqCols& k, uint32_t row_off /* c * GQ_T_CH + r * GQ_TBW * 4 */, float (&o)[4]) {
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        if (CHIP == 1) {
            const uint32_t a0 = k.t[pr][0] + row_off;
            const float ta = lds_f32(a0), tb_ = lds_f32(a0 + 4), tc = lds_f32(a0 + 8);
            const float ua = k.d[pr] ? tb_ : ta, ub = k.d[pr] ? tc : tb_;
            o[2 * pr] = FIRST ? k.w[2 * pr][0] * ta : fmaf(k.w[2 * pr][0], ta, o[2 * pr]);
            o[2 * pr] = fmaf(k.w[2 * pr][1], tb_, o[2 * pr]);
            o[2 * pr + 1] = FIRST ? k.w[2 * pr + 1][0] * ua : fmaf(k.w[2 * pr + 1][0], ua, o[2 * pr + 1]);
            o[2 * pr + 1] = fmaf(k.w[2 * pr + 1][1], ub, o[2 * pr + 1]);
        } else {
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const uint32_t a = k.t[pr][v] + row_off;
                const float2 hi = lds_v2f32(k.w23 + (2 * pr + v) * 8);
                float acc = FIRST ? k.w[2 * pr + v][0] * lds_f32(a) : fmaf(k.w[2 * pr + v][0], lds_f32(a), o[2 * pr + v]);
                acc = fmaf(k.w[2 * pr + v][1], lds_f32(a + 4), acc);
                acc = fmaf(hi.x, lds_f32(a + 8), acc); acc = fmaf(hi.y, lds_f32(a + 12), acc);
                o[2 * pr + v] = acc;
            }
        }
    }
}

// One unit of the horizontal pass: NR image rows (row, row + 1), all three channels, this thread's four columns.
//   SMALL: the rows take the resized row at shared address s_addr (+ channel stride) with y weights wa (row) / wb (row + 1)
template <typename T, int CHIP, int NR, bool SMALL>
__device__ __forceinline__ void gq_unit(const GqCols& k, float wa, float wb, bool ina, bool inb, uint32_t s_addr, uint32_t t_off, char* orow) {
    constexpr unsigned oplb = 512u * 512u * (unsigned)sizeof(T), orowb = 512u * (unsigned)sizeof(T);
    constexpr int S_CH = GqT<T>::S_CH;
    float ca[4], cb[4];
    if (SMALL) {
#pragma unroll
        for (int i = 0; i < 4; i++) { ca[i] = wa * k.wx[i]; if (NR == 2) cb[i] = wb * k.wx[i]; }
        if (ina || inb) {                                       // warp-uniform: rows of the scaled region (E3:1751-1784) only
            const float fa = ina ? k.rs : 1.f, fb = inb ? k.rs : 1.f;
#pragma unroll
            for (int i = 0; i < 4; i++) { if (k.rin[i]) { ca[i] *= fa; if (NR == 2) cb[i] *= fb; } }
        }
    }
    // not unrolled: the six (chip mode x half) variants of this pass run side by side in one CTA and must share the
    // instruction cache (unrolled, the kernel was 100 KB of code and a fifth of the issue slots waited for instructions)
    FG_UNROLL(GQ_CUNROLL)
    for (int c = 0; c < 3; c++) {
        float oa[4], ob[4];
        if (SMALL) {
            float vA, vB;
            load_ab<T>(s_addr + c * S_CH, k.s_perm, vA, vB);
            // a column without a tap takes a literal zero (not a zero weight): a non-finite gradient stays in the pixels it touches
            const float v0 = k.j1 ? 0.f : vA, v1 = k.j2 ? vB : vA, v2 = k.j1 ? vA : vB, v3 = k.j2 ? 0.f : vB;
            oa[0] = ca[0] * v0; oa[1] = ca[1] * v1; oa[2] = ca[2] * v2; oa[3] = ca[3] * v3;
            if (NR == 2) { ob[0] = cb[0] * v0; ob[1] = cb[1] * v1; ob[2] = cb[2] * v2; ob[3] = cb[3] * v3; }
        }
        if (CHIP != 0) {
            gq_chip_row<CHIP, !SMALL>(k, c * GQ_T_CH + t_off, oa);
            if (NR == 2) gq_chip_row<CHIP, !SMALL>(k, c * GQ_T_CH + t_off + GQ_TBW * 4, ob);
        } else if (!SMALL) {
#pragma unroll
            for (int i = 0; i < 4; i++) { oa[i] = 0.f; ob[i] = 0.f; }
        }
        store_quad<T>(orow + c * oplb, oa[0], oa[1], oa[2], oa[3]);
        if (NR == 2) store_quad<T>(orow + c * oplb + orowb, ob[0], ob[1], ob[2], ob[3]);
    }
}

// horizontal pass of one sub-tile for one half of the CTA (HALF 1: two pairs; HALF 0: pair, tap-less row, single row)
template <typename T, bool DO_S, int CHIP, int HALF>
__device__ __forceinline__ void gq_stage2(const GqCols& k, const GqSched& sc, uint32_t wy_a /* wy table of the staged block */,
                                          uint32_t sbuf /* + k.s_off */, char* out /* row 0 of the sub-tile, this thread's columns */,
                                          int y0, int ry0, int ry1) {
    constexpr unsigned orowb = 512u * (unsigned)sizeof(T);
    constexpr int GQ_ROWB = GqT<T>::ROWB;
    auto in_reg = [&](int r) { const int y = y0 + r; return y >= ry0 && y < ry1; };
    if (HALF == 1) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = sc.row[u];
            float2 l = make_float2(0.f, 0.f);
            if (DO_S) l = lds_v2f32(wy_a + sc.srow[u] * 8);
            gq_unit<T, CHIP, 2, DO_S>(k, l.x, l.y, in_reg(r), in_reg(r + 1), sbuf + sc.srow[u] * GQ_ROWB, (unsigned)r * (GQ_TBW * 4),
                                      out + (unsigned)r * orowb);
        }
    } else {
        {
            const int r = sc.row[0];
            float2 l = make_float2(0.f, 0.f);
            if (DO_S) l = lds_v2f32(wy_a + sc.srow[0] * 8);
            gq_unit<T, CHIP, 2, DO_S>(k, l.x, l.y, in_reg(r), in_reg(r + 1), sbuf + sc.srow[0] * GQ_ROWB, (unsigned)r * (GQ_TBW * 4),
                                      out + (unsigned)r * orowb);
        }
        {
            const int r = sc.row[1];
            gq_unit<T, CHIP, 1, false>(k, 0.f, 0.f, false, false, 0u, (unsigned)r * (GQ_TBW * 4), out + (unsigned)r * orowb);
        }
        {
            const int r = sc.row[2];
            float2 l = make_float2(0.f, 0.f);
            if (DO_S) l = lds_v2f32(wy_a + sc.srow[2] * 8);
            gq_unit<T, CHIP, 1, DO_S>(k, sc.tap ? l.y : l.x, 0.f, in_reg(r), false, sbuf + sc.srow[2] * GQ_ROWB, (unsigned)r * (GQ_TBW * 4),
                                      out + (unsigned)r * orowb);
        }
    }
}

#ifndef GQ_MINB
#define GQ_MINB 2                 // resident CTAs per SM the register allocation targets (shared memory allows 2)
#endif
// HAS_S: the launch carries a resized-image gradient (compile-time, so that a kernel holds only the variants it runs)
template <typename T, int NSUB, bool HAS_S>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : GQ_MINB)
image_grad_quad_kernel(const BwdParams p) {
    static_assert(NSUB % 2 == 0, "a band covers whole blocks of 16 rows");
    using L = GqLayout<T, NSUB>;
    constexpr int GQ_ROWB = GqT<T>::ROWB, GQ_S_CH = GqT<T>::S_CH, GQ_C_CH = GqT<T>::C_CH;
    constexpr int C = 3, H = 512, W = 512, OH = GQ_OW, OW = GQ_OW, ROWS = NSUB * GQ_ROWS;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
    GqSub* subs = reinterpret_cast<GqSub*>(smem + L::subs);
    float2* wyS = reinterpret_cast<float2*>(smem + L::wy);
    GqRowC* rowC = reinterpret_cast<GqRowC*>(smem + L::rowC);
    float2* w23 = reinterpret_cast<float2*>(smem + L::w23);
    T* bufS = reinterpret_cast<T*>(smem + L::bufS);   // [2][C][GQ_SROWS][OW]
    T* sC = reinterpret_cast<T*>(smem + L::sC);       // [C][GQ_CROWS][OW]
    float* tb = reinterpret_cast<float*>(smem + L::tb);   // [C][GQ_ROWS][GQ_TBW]
    const uint32_t smem_a = smem_u32(smem);

    const int img = blockIdx.y, ybase = blockIdx.x * ROWS, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr bool has_s = HAS_S;
    Box b; b.ok = false; b.x0 = b.y0 = b.x1 = b.y1 = 0;
    if (p.g_chips) b = load_box(p.boxes, p.ind, img, H, W);
    const int bw = b.x1 - b.x0, bh = b.y1 - b.y0;
    const float ss = (float)W / (float)OW;
    const float csx = b.ok ? (float)bw / (float)OW : 1.f, csy = b.ok ? (float)bh / (float)OH : 1.f;
    bool chip_cold = b.ok && (csx < 0.51f || csy < 0.51f);
    const bool chip_tab = b.ok && !chip_cold;
    int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0; float rs = 1.f;
    if (p.region) {
        rx0 = p.region[4 * img]; ry0 = p.region[4 * img + 1]; rx1 = p.region[4 * img + 2]; ry1 = p.region[4 * img + 3];
        rs = p.scale[img];
    }

    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // pad columns of tb (read with zero weights, so they must stay finite)
    for (int e = tid; e < C * GQ_ROWS * TPAD; e += 256) {
        const int cr = e / TPAD, q = e - cr * TPAD;
        tb[cr * GQ_TBW + OW + q] = 0.f;
    }
    // y weights of the band's resized rows: o = 7 * (ybase / 16) + e
    const int o_base = 7 * (ybase >> 4);
    if (has_s && tid < (ROWS / 16) * 7) {
        const Axis a = axis_index(o_base + tid, ss, H);
        wyS[tid] = make_float2(a.l0, a.l1);
    }
    int too_many = 0;
    for (int r = tid; r < ROWS; r += 256) {
        Tab t; t.lo = 0; t.n = 0; t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
        if (chip_tab) t = make_tab(ybase + r - b.y0, csy, bh, OH);
        if (t.n > TABW) too_many = 1;
        GqRowC rc; rc.off = t.lo; rc.n = t.n; rc.pad[0] = rc.pad[1] = 0;
#pragma unroll
        for (int q = 0; q < TABW; q++) rc.w[q] = t.w[q];
        rowC[r] = rc;
    }
    // this thread's four image columns
    // Columns: a warp of a half owns two blocks of 64 columns, 256 columns apart (lanes 0-15 / 16-31), so that the columns of
    // a face box spread over all four warps (with contiguous 128-column blocks the middle warps did all the chip work and
    // the others waited at the barrier); a store instruction still writes whole 128-byte lines.
    const int half = tid >> 7;
    const int quad = GQ_INTERLEAVE ? ((lane >> 4) << 6) + ((warp & 3) << 4) + (lane & 15) : (tid & 127);
    const int x_a = 4 * quad;
    GqCols k;
    {
        // the two resized columns that feed these four image columns: block of 16 columns m, quarter j
        const int m = quad >> 2, j = quad & 3;
        const int o_first = 7 * m + (j == 0 ? 0 : j == 1 ? 2 : j == 2 ? 3 : 5);
        const Axis aA = axis_index(o_first, ss, W), aB = axis_index(o_first + 1, ss, W);
        k.s_off = sizeof(T) == 4 ? (uint32_t)o_first * 4u : (uint32_t)(o_first >> 1) * 4u;
        k.s_perm = (o_first & 1) ? 0x5432u : 0x3210u;
        k.j1 = j == 1; k.j2 = j == 2;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = x_a + i;
            // the weight of column x: from A or B, whichever taps it (the source pattern itself is static, see GqCols)
            float w = 0.f;
            if (has_s) {
                if (aA.i0 == x) w = aA.l0; else if (aA.i1 == x) w = aA.l1; else if (aB.i0 == x) w = aB.l0; else if (aB.i1 == x) w = aB.l1;
            }
            k.wx[i] = w;
            k.rin[i] = x >= rx0 && x < rx1;
        }
        k.rs = rs;
    }
    int xc_n[4], xc_lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        Tab u; u.lo = 0; u.n = 0;
#pragma unroll
        for (int q = 0; q < TABW; q++) u.w[q] = 0.f;
        if (chip_tab) u = make_tab(x_a + i - b.x0, csx, bw, OW);
        if (u.n > TABW) too_many = 1;
        if (u.n == 0) u.lo = 0;
        xc_n[i] = u.n; xc_lo[i] = u.lo;
        k.w[i][0] = u.w[0]; k.w[i][1] = u.w[1];
        w23[x_a + i] = make_float2(u.w[2], u.w[3]);
    }
    k.w23 = smem_a + (uint32_t)L::w23 + (uint32_t)x_a * 8u;
    if (__syncthreads_or(too_many)) chip_cold = true;               // also orders the table writes
    // per sub-tile: which chip rows to stage; turn absolute rows into byte offsets inside the staging buffer
    too_many = 0;
    if (tid < NSUB) {
        int c_lo = 1 << 30, c_hi = -1;
        for (int r = 0; r < GQ_ROWS; r++) {
            const int co = rowC[tid * GQ_ROWS + r].off, cn = rowC[tid * GQ_ROWS + r].n;
            if (cn > 0) { c_lo = min(c_lo, co); c_hi = max(c_hi, min(co + cn - 1, OH - 1)); }
        }
        GqSub d;
        d.c_first = c_hi >= 0 ? c_lo : 0; d.c_count = c_hi >= 0 ? c_hi - c_lo + 1 : 0; d.pad0 = d.pad1 = 0;
        if (d.c_count > GQ_CROWS) too_many = 1;
        subs[tid] = d;
        for (int r = 0; r < GQ_ROWS; r++) {
            GqRowC& c = rowC[tid * GQ_ROWS + r];
            c.off = (c.n > 0 ? c.off - d.c_first : 0) * GQ_ROWB;
        }
    }
    if (__syncthreads_or(too_many)) chip_cold = true;
    const bool chip_fast = b.ok && !chip_cold;

    const T* gs_img = reinterpret_cast<const T*>(p.g_small) + (size_t)img * C * OH * OW;
    const T* gc_img = reinterpret_cast<const T*>(p.g_chips) + (size_t)img * C * OH * OW;
    // The gradient rows of sub-tile `sub` are fetched a whole iteration ahead (HBM latency under load exceeds the time one
    // half-iteration takes): the resized rows into bufS[sub & 1], the chip rows into sC[sub % NSC].
    constexpr int NSC = L::NSC;
    auto issue_small = [&](int sub) {        // thread 0 only
        if (has_s) {
            // rows 0-7 of a block of 16 read resized rows 7g..7g+3, rows 8-15 read 7g+3..7g+6
            const int s_first = o_base + 7 * (sub >> 1) + 3 * (sub & 1);
            uint64_t* bar = &bars[sub & 1];
            constexpr uint32_t bytes = GQ_SROWS * GQ_ROWB;
            mbar_arrive_expect_tx(bar, C * bytes);
#pragma unroll
            for (int c = 0; c < C; c++)
                bulk_g2s(reinterpret_cast<char*>(bufS) + ((sub & 1) * C + c) * GQ_S_CH, gs_img + c * OH * OW + s_first * OW, bytes, bar);
        }
    };
    auto issue_chip = [&](int sub) {         // thread 0 only
        const GqSub d = subs[sub];
        if (chip_fast && d.c_count > 0) {
            const uint32_t bytes = (uint32_t)d.c_count * GQ_ROWB;
            uint64_t* bar = &bars[2 + sub % NSC];
            mbar_arrive_expect_tx(bar, C * bytes);
#pragma unroll
            for (int c = 0; c < C; c++)
                bulk_g2s(reinterpret_cast<char*>(sC) + ((sub % NSC) * C + c) * GQ_C_CH, gc_img + c * OH * OW + d.c_first * OW, bytes, bar);
        }
    };
    if (tid == 0) { issue_small(0); issue_chip(0); }

    // chip taps of the two column pairs: a column without a tap borrows its partner's start (its weights are zero); the
    // three-word path needs <= 2 taps per column and starts 0 or 1 apart, in every lane of the warp
    bool narrow = true;
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        int lo0 = xc_lo[2 * pr], lo1 = xc_lo[2 * pr + 1];
        if (xc_n[2 * pr] == 0) lo0 = lo1;
        if (xc_n[2 * pr + 1] == 0) lo1 = lo0;
        k.t[pr][0] = smem_a + (uint32_t)L::tb + (uint32_t)lo0 * 4u;
        k.t[pr][1] = smem_a + (uint32_t)L::tb + (uint32_t)lo1 * 4u;
        k.d[pr] = lo1 != lo0;
        narrow = narrow && xc_n[2 * pr] <= 2 && xc_n[2 * pr + 1] <= 2 && (unsigned)(lo1 - lo0) <= 1u;
    }
    const bool warp_has_chip = __any_sync(0xffffffffu, (xc_n[0] | xc_n[1] | xc_n[2] | xc_n[3]) != 0);
    const bool warp_narrow = __all_sync(0xffffffffu, narrow);
    char* go_img = reinterpret_cast<char*>(p.g_images) + (size_t)img * C * H * W * sizeof(T) + (size_t)x_a * sizeof(T);
    constexpr unsigned orowb = (unsigned)(W * sizeof(T));
    uint32_t phS = 0u, phC = 0u;          // mbarrier phase parities (bit k of phS: bufS[k])
    // vertical pass: warps 0-3 take rows 0-3, warps 4-7 rows 4-7; lanes 0-27 of a warp own chip column pairs 28 * (warp & 3) + lane
    const int s1_pair = 28 * (warp & 3) + lane, s1_row0 = 4 * (warp >> 2);
    const bool s1_active = lane < 28;

#pragma unroll 1
    for (int sub = 0; sub < NSUB; sub++) {
        const int y0 = ybase + sub * GQ_ROWS;
        const GqSub d = subs[sub];
        const bool chip_rows = chip_fast && d.c_count > 0;
        // both buffers of the next sub-tile are free here (the barrier that ended the previous iteration)
        if (tid == 0 && sub + 1 < NSUB) { issue_small(sub + 1); if (NSC == 2) issue_chip(sub + 1); }
        if (chip_rows) {
            mbar_wait(&bars[2 + sub % NSC], (phC >> (sub % NSC)) & 1u); phC ^= 1u << (sub % NSC);
            if (s1_active)
                gq_stage1<T>(smem_a + (uint32_t)L::rowC + (sub * GQ_ROWS + s1_row0) * 32,
                             smem_a + (uint32_t)L::sC + (sub % NSC) * C * GQ_C_CH + s1_pair * 2 * (int)sizeof(T),
                             smem_a + (uint32_t)L::tb + s1_row0 * (GQ_TBW * 4) + s1_pair * 8);
            __syncthreads();
        }
        if (NSC == 1 && tid == 0 && sub + 1 < NSUB) issue_chip(sub + 1);      // single buffer: free once the vertical pass is done

        if (has_s) { mbar_wait(&bars[sub & 1], (phS >> (sub & 1)) & 1u); phS ^= 1u << (sub & 1); }
        const uint32_t sbuf = smem_a + (uint32_t)L::bufS + (sub & 1) * C * GQ_S_CH;
        const uint32_t wy_a = smem_a + (uint32_t)L::wy + (7 * (sub >> 1) + 3 * (sub & 1)) * 8;
        char* out = go_img + (unsigned)y0 * orowb;
        const GqSched& sc = gq_sched[sub & 1][half];
        if (!chip_cold) {
            const int mode = !(chip_rows && warp_has_chip) ? 0 : (warp_narrow ? 1 : 2);
            const uint32_t sb = sbuf + k.s_off;
#define GQ_S2(DS, CH)                                                                             \
            do {                                                                                  \
                if (half) gq_stage2<T, DS, CH, 1>(k, sc, wy_a, sb, out, y0, ry0, ry1);            \
                else gq_stage2<T, DS, CH, 0>(k, sc, wy_a, sb, out, y0, ry0, ry1);                 \
            } while (0)
            if (mode == 0) GQ_S2(HAS_S, 0); else if (mode == 1) GQ_S2(HAS_S, 1); else GQ_S2(HAS_S, 2);
#undef GQ_S2
        } else {
            // rare: a box the staging buffer cannot hold (tiny or very elongated); its pixels take the direct 2-D gather from
            // global memory, the resized-image branch still comes from the staged rows
            const int n_units = half ? 2 : 3;
#pragma unroll 1
            for (int u = 0; u < n_units; u++) {
                const int nrows = (half || u == 0) ? 2 : 1, srow = sc.srow[u];
                const bool small = has_s && !(half == 0 && u == 1);
                float2 l = make_float2(0.f, 0.f);
                if (small) l = lds_v2f32(wy_a + srow * 8);
#pragma unroll 1
                for (int rr = 0; rr < nrows; rr++) {
                    const int r = sc.row[u] + rr, y = y0 + r;
                    const float wy = !small ? 0.f : (nrows == 2 ? (rr ? l.y : l.x) : (sc.tap ? l.y : l.x));
                    const bool in_y = y >= ry0 && y < ry1;
#pragma unroll 1
                    for (int c = 0; c < C; c++) {
                        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                        if (small) {
                            const uint32_t a = sbuf + k.s_off + c * GQ_S_CH + srow * GQ_ROWB;
                            float vA, vB;
                            load_ab<T>(a, k.s_perm, vA, vB);
                            const float f0 = (in_y && k.rin[0]) ? k.rs : 1.f, f1 = (in_y && k.rin[1]) ? k.rs : 1.f;
                            const float f2 = (in_y && k.rin[2]) ? k.rs : 1.f, f3 = (in_y && k.rin[3]) ? k.rs : 1.f;
                            o0 = wy * k.wx[0] * f0 * (k.j1 ? 0.f : vA); o1 = wy * k.wx[1] * f1 * (k.j2 ? vB : vA);
                            o2 = wy * k.wx[2] * f2 * (k.j1 ? vA : vB); o3 = wy * k.wx[3] * f3 * (k.j2 ? 0.f : vB);
                        }
                        float o[4] = {o0, o1, o2, o3};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int x = x_a + i;
                            if (x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1) {
                                float acc[FG_MAXC] = {0.f, 0.f, 0.f, 0.f};
                                gather_grid_cold<T>(gc_img + (size_t)c * OH * OW, 1, OH, OW, x - b.x0, y - b.y0, bw, bh, acc);
                                o[i] += acc[0];
                            }
                        }
                        store_quad<T>(out + (unsigned)r * orowb + (size_t)c * H * W * sizeof(T), o[0], o[1], o[2], o[3]);
                    }
                }
            }
        }
        __syncthreads();
    }
}
