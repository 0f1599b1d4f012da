// Distributional-alignment target assignment on the device.
//
//   generate_dynamic_targets                  E1:1403-1447   rank split + binomial-CDF uncertainty
//   generate_dynamic_targets_gender_race      E3:1459-1569   Monte-Carlo exact transport, K = 8
//   generate_dynamic_targets_gender_race_age  E4:1477-1615   same, K = 16, 75/25 age target
//   thresholding                              E1:1835, E3:2022-2023, E4:2129-2131
//
// The reference solves, per rank, S = 100 transport problems  ot.emd(ones(N), b_s, M)  that share
// the cost matrix M [N,K] and differ only in the demand vector b_s (the class histogram of one
// random draw).  With unit supplies the optimum is a 0/1 matrix, i.e. an assignment of rows to
// classes with prescribed class sizes.  Exact method used here:
//
//   * an assignment that minimises  sum_i (M[i,s(i)] - v[s(i)])  for ANY price vector v is
//     optimal among all assignments with the same class sizes;
//   * from such a state, moving one unit from an over-full class to an under-full class along a
//     SHORTEST path of the K-node class graph (edge k->l costs  min_{i in k} M[i,l]-M[i,k])
//     keeps optimality for the new class sizes (successive shortest paths);
//   * so every draw runs, in its own CTA, a cheap parallel PRICE SEARCH (sign-based dual ascent on the K prices in
//     fp32: any prices are admissible, it only has to bring the class sizes close to the demand) and then repairs
//     the remaining  |sizes - b_s|_1 / 2  units exactly, one shortest path each.
//
// One CTA (512 threads) per problem.  FP64 issues at a small fraction of the FP32 rate on this part and the fp64
// cost matrix stays in global memory, so shared memory holds a class-major FP32 copy (written once by the cost
// kernel) that serves the price search and SCREENS every exact decision: the final argmin of a row and the edge
// minima of the class graph are taken in fp32 first, and fp64 is evaluated only for the winner or, when the runner-up
// is within SCREEN_EPS, for the near-ties.  The class graph lives in shared memory; edge minima use warp-level REDUX
// on order-preserving keys; the shortest-path search (Bellman-Ford over <= 16 nodes) runs in one warp.  Every result
// that leaves the kernel is decided by fp64 arithmetic like POT's; ties resolve to the lowest row index / lowest
// class index, deterministically.  fg_ot_solve_single additionally exercises a warm-started second solve.
#include "fg_common.cuh"
#include <math.h>
#include <cstdlib>
#include <cstdio>

namespace {

constexpr int KP = 16;                 // classes padded to 16
constexpr int SOLVER_THREADS = 512;
constexpr int SOLVER_WARPS = SOLVER_THREADS / 32;
constexpr unsigned long long KEY_INF = 0xFFFFFFFFFFFFFFFFull;

enum { ST_NVALID_MISMATCH = 1, ST_NO_PATH = 2, ST_PATH_OVERFLOW = 4, ST_ITER_CAP = 8, ST_BAD_DEMAND = 16 };

__device__ __forceinline__ unsigned long long dkey(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------------------------------------
// workspace layout
struct OtWs {
    int* status;        // [4]: status bits, n_valid seen on device, augmentations (base), augmentations (draws)
    int* idx;           // [n_all] compacted row -> original row
    int* pos;           // [n_all] original row -> compacted row or -1
    double* M;          // [n_valid, K]
    float* Mf;          // [K, n_valid]  fp32, class-major (the solver's search / screening copy)
    int* Mk;            // [K, n_valid]  integer search keys of the same costs (cost_key), class-major
    int* hist;          // [S, KP]
    double* prices;     // [KP]
    uint8_t* sigma0;    // [n_valid]
    uint16_t* members;  // [S][K][n_valid] member lists of the solver's mode 2 (empty otherwise)
    size_t total;
};

static size_t solver_members_bytes(int N, int K, int S);

static OtWs ot_carve(void* base, int n_all, int K, int S) {
    OtWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += fg_align_up(bytes, 256); return (char*)base + o; };
    w.status = (int*)take(4 * sizeof(int));
    w.idx = (int*)take((size_t)n_all * sizeof(int));
    w.pos = (int*)take((size_t)n_all * sizeof(int));
    w.M = (double*)take((size_t)n_all * K * sizeof(double));
    w.Mf = (float*)take((size_t)n_all * K * sizeof(float));
    w.Mk = (int*)take((size_t)n_all * K * sizeof(int));
    w.hist = (int*)take((size_t)(S > 0 ? S : 1) * KP * sizeof(int));
    w.prices = (double*)take(KP * sizeof(double));
    w.sigma0 = (uint8_t*)take((size_t)n_all);
    w.members = (uint16_t*)take(solver_members_bytes(n_all, K, S));
    w.total = off;
    return w;
}

// Integer search key of a cost for the price search: the cost in units of 2^-22 (costs are norms of probability
// differences, < 4; the fp32 spacing at 2..4 is exactly one unit), shifted left by four with the class index in the low
// bits.  Subtracting a price key (a multiple of 16) keeps the class bits, so the cheapest class of a row is ONE integer
// minimum over its K keys -- value and index together, ties to the lower class -- instead of a compare/select tree that
// carries the index separately (~30 instead of ~70 instructions per row at K = 16).  The search is a heuristic (any
// prices are admissible), so the quantisation only costs an occasional extra repair step.
constexpr float KEY_SCALE = 4194304.f;             // 2^22
__device__ __forceinline__ int cost_key(float x, int l) { return (__float2int_rn(fminf(fmaxf(x, 0.f), 16.f) * KEY_SCALE) << 4) | l; }
__device__ __forceinline__ int price_key(float p) { return __float2int_rn(fminf(fmaxf(p, -8.f), 8.f) * KEY_SCALE) << 4; }
__device__ __forceinline__ float price_of_key(int pk) { return (float)(pk >> 4) * (1.f / KEY_SCALE); }
__device__ __forceinline__ int min3(int a, int b, int c) { return min(min(a, b), c); }
template <int KK>
__device__ __forceinline__ int min_key(const int (&y)[KK]) {
    if (KK == 16) {
        const int a0 = min3(y[0], y[1], y[2]), a1 = min3(y[3], y[4], y[5]), a2 = min3(y[6], y[7], y[8]);
        const int a3 = min3(y[9], y[10], y[11]), a4 = min3(y[12], y[13], y[14]);
        return min(min3(a0, a1, a2), min3(a3, a4, y[15]));
    }
    const int a0 = min3(y[0], y[1], y[2]), a1 = min3(y[3], y[4], y[5]);
    return min3(a0, a1, min(y[6], y[KK - 1]));
}

// ---------------------------------------------------------------------------------------------
// valid-row compaction: idx / pos, one CTA, ordered
template <typename T>
__global__ void __launch_bounds__(1024)
compact_kernel(const T* __restrict__ pg, const T* __restrict__ pr, int n_all, int n_valid_expected,
               int* __restrict__ idx, int* __restrict__ pos, int* __restrict__ status) {
    __shared__ int warp_sums[32];
    __shared__ int base;
    if (threadIdx.x == 0) { base = 0; status[0] = 0; status[2] = 0; status[3] = 0; }      // first kernel of the call: clears the status words
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < n_all; start += 1024) {
        int i = start + threadIdx.x;
        bool v = false;
        if (i < n_all) {
            v = !pg || (to_f32(pg[2 * i]) != -1.f && to_f32(pg[2 * i + 1]) != -1.f);      // race-only variant (E6): pg == nullptr
            for (int q = 0; q < 4; q++) v = v && to_f32(pr[4 * i + q]) != -1.f;
        }
        unsigned m = __ballot_sync(0xffffffffu, v);
        int in_warp = __popc(m & ((1u << lane) - 1));
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            int s = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
            warp_sums[lane] = s;          // inclusive
        }
        __syncthreads();
        int before = base + (warp ? warp_sums[warp - 1] : 0) + in_warp;
        if (i < n_all) {
            pos[i] = v ? before : -1;
            if (v) idx[before] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) base += warp_sums[31];
        __syncthreads();
    }
    // A caller that over-reports n_valid makes the cost kernel walk idx[] past the rows found here: point those entries
    // at row 0 so that nothing is read out of bounds (the mismatch is flagged; the Python side raises on it).
    for (int i = base + threadIdx.x; i < n_all; i += 1024) idx[i] = 0;
    if (threadIdx.x == 0) {
        status[1] = base;
        if (n_valid_expected >= 0 && base != n_valid_expected) atomicOr(&status[0], ST_NVALID_MISMATCH);
    }
}

// ---------------------------------------------------------------------------------------------
// cost matrix (fixed IEEE op order, identical to oracle/emd_rowinsert.c) + per-draw histograms
__device__ __forceinline__ double dsq(double x) { return __dmul_rn(x, x); }

// one element M[i,j]; every thread of a row recomputes the row's norms (a few fp64 square roots) so that the matrix is
// spread over 16x more threads than one-thread-per-row -- the kernel sits on the critical path of the assignment
template <typename T>
__device__ __forceinline__ double cost_elem(const T* __restrict__ pg, const T* __restrict__ pr, const T* __restrict__ pa,
                                            int i, int K, int j) {
    const double g[2] = {(double)to_f32(pg[2 * i]), (double)to_f32(pg[2 * i + 1])};
    double r[4];
    for (int q = 0; q < 4; q++) r[q] = (double)to_f32(pr[4 * i + q]);
    int gi, ri, ai = 0;
    if (K == 8) { gi = j >> 2; ri = j & 3; } else { gi = j >> 3; ri = (j >> 1) & 3; ai = j & 1; }
    const double ng = __dsqrt_rn(__dadd_rn(dsq(__dsub_rn(g[0], gi == 0 ? 1.0 : 0.0)), dsq(__dsub_rn(g[1], gi == 1 ? 1.0 : 0.0))));
    double s4 = dsq(__dsub_rn(r[0], ri == 0 ? 1.0 : 0.0));
    for (int q = 1; q < 4; q++) s4 = __dadd_rn(s4, dsq(__dsub_rn(r[q], ri == q ? 1.0 : 0.0)));
    const double nr = __dsqrt_rn(s4);
    double c = __dadd_rn(dsq(ng), dsq(nr));
    if (K == 16) {
        const double a0 = (double)to_f32(pa[2 * i]), a1 = (double)to_f32(pa[2 * i + 1]);
        const double ca = ai == 0 ? __dsqrt_rn(__dadd_rn(dsq(__dsub_rn(a0, 1.0)), dsq(__dsub_rn(a1, 0.0))))
                                  : __dsqrt_rn(__dadd_rn(dsq(__dmul_rn(__dsub_rn(a0, 0.0), 2.0)), dsq(__dsub_rn(a1, 1.0))));
        c = __dadd_rn(c, dsq(ca));
    }
    return __dsqrt_rn(c);
}

template <typename T>
__global__ void __launch_bounds__(256)
cost_hist_kernel(const T* __restrict__ pg, const T* __restrict__ pr, const T* __restrict__ pa,
                 const int* __restrict__ idx, int N, int K, double* __restrict__ M, float* __restrict__ Mf, int* __restrict__ Mk, int cost_blocks,
                 const T* __restrict__ rg, const T* __restrict__ rr, const T* __restrict__ ra, int S,
                 int* __restrict__ hist, int32_t* __restrict__ counts_to_zero) {
    if ((int)blockIdx.x < cost_blocks) {
        // 256 threads = 256 / K rows x K classes
        const int e = blockIdx.x * 256 + threadIdx.x;
        const int r = e / K, j = e - r * K;
        if (r < N) {
            const double c = cost_elem<T>(pg, pr, pa, idx[r], K, j);
            M[(size_t)r * K + j] = c;
            if (Mf) { Mf[(size_t)j * N + r] = (float)c; Mk[(size_t)j * N + r] = cost_key((float)c, j); }
            if (counts_to_zero) counts_to_zero[(size_t)r * K + j] = 0;
        }
        return;
    }
    int s = blockIdx.x - cost_blocks;
    if (s >= S) return;
    __shared__ int h[KP];
    if (threadIdx.x < KP) h[threadIdx.x] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += 256) {
        size_t e = (size_t)s * N + j;
        float ug = to_f32(rg[e]), ur = to_f32(rr[e]);
        int g = ug > 0.5f ? 1 : 0;                                           // E3:1496
        int r = 0;                                                            // E3:1498-1501
        if (ur > 0.25f && ur <= 0.5f) r = 1;
        if (ur > 0.5f && ur <= 0.75f) r = 2;
        if (ur > 0.75f) r = 3;
        int c;
        if (K == 8) c = g * 4 + r;                                            // E3:1506
        else { int a = to_f32(ra[e]) > 0.75f ? 1 : 0; c = g * 8 + r * 2 + a; }  // E4:1518, 1523
        atomicAdd(&h[c], 1);
    }
    __syncthreads();
    if (threadIdx.x < KP) hist[s * KP + threadIdx.x] = h[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// the solver
struct Demand { int b[KP]; };

struct SolverSmem {
    double w[KP][KP];          // w[k][l] = min over rows i currently in k of M[i,l] - M[i,k]
    int wi[KP][KP];            // the row that attains it (lowest index on ties), -1 if k is empty
    double dist[KP];
    double price[KP];
    float pricef[KP];          // price search state (fp32)
    float best_pricef[KP];     // prices with the smallest count error seen by the price search
    float step[KP];            // per-class adaptive step of the price search
    int prev_sign[KP];
    int best_resid;
    int stop;
    int pred[KP];
    int cnt[KP];
    int b[KP];
    int path[KP + 1];
    int moved[KP + 1];
    unsigned need[KP + 1];
    int path_len;
    int cont[2];               // loop-continue flag, double buffered by iteration parity
    unsigned whist[2][SOLVER_WARPS][KP / 2];   // price search: per-warp class counts (packed), by iteration parity
    __align__(16) float wprice[SOLVER_WARPS][KP];   // price search: the warp's copy of the K prices of this round
    int status;
#ifdef FG_OT_PROFILE
    int prof_pivots, prof_pivot_last, prof_rounds;
#endif
    int rcnt[3][KP];           // price search: class counts of the round (atomic adds of the warp sums), three buffers in rotation
    int n_active;              // trust-region price search: rows in the active list, and whether the list overflowed
    int tr_overflow;
    // first scan of the large modes (sweep_assign_scan): per (class k, target l) the fp32 minimum of M[i,l] - M[i,k] over the
    // rows of k as an order-preserving key, the number of rows within SCREEN_EPS of it and the lowest such row.  Rows of 17
    // words: the 32 lanes of an access hold different k and the same l, so a stride of 16 would put them on two banks.
    unsigned smin[KP][KP + 1];
    int scnt[KP][KP + 1];
    int sidx[KP][KP + 1];
    unsigned sexact[KP];       // targets of class k left to the exact scan (near-ties)
};

// dynamic shared memory behind the control block
struct SolverViews {
    uint8_t* sigma;            // [N]      class of every row
    uint16_t* pos;             // [N]      position of the row inside its class list
    uint16_t* members;         // [K][N]   unordered member list per class
    float* Mf;                 // [K][N]   fp32 copy of the cost matrix for the price search (when it fits)
};

constexpr size_t SOLVER_CTRL = (sizeof(SolverSmem) + 15) / 16 * 16;      // control block, padded so that every view is 16-byte aligned
__host__ __device__ inline size_t solver_off_pos(int N) { return SOLVER_CTRL + (((size_t)N + 15) / 16) * 16; }
__host__ __device__ inline size_t solver_off_members(int N) { return solver_off_pos(N) + (((size_t)N * 2 + 15) / 16) * 16; }
__host__ __device__ inline size_t solver_off_M(int N, int K) { return solver_off_members(N) + (((size_t)N * K * 2 + 15) / 16) * 16; }

// order-preserving 32-bit key of a float (for REDUX min)
__device__ __forceinline__ unsigned fkey(float f) { const unsigned b = __float_as_uint(f); return (b >> 31) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float funkey(unsigned k) { return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k); }
// |fp32 difference of the fp32 copies - fp64 difference| <= 3 half-ulps of values below 8 (costs are norms of probability
// differences): 7.2e-7.  Two candidates further apart than twice that in fp32 are ordered the same way in fp64.
constexpr float SCREEN_EPS = 4e-6f;

// One screening pass over the members of class k for NT target classes: fp32 minimum, runner-up and argmin per target
// and lane.  The loads of U members per lane are issued together: for the large problems the member lists and the cost
// copy sit in L2 / global memory and a pass is a chain of dependent round trips (member id -> its costs), so the number
// of loads in flight, not the arithmetic, sets its duration (one member at a time: 181 k cycles for the first scan of
// all classes at N = 7720).
template <int NT, int U>
__device__ __forceinline__ void screen_pass(const SolverViews& v, int N, int k, int cnt, const int (&ls)[NT],
                                            float (&b1)[NT], float (&b2)[NT], int (&bi)[NT]) {
    const int lane = threadIdx.x & 31;
    const uint16_t* mem = v.members + (size_t)k * N;
    const float* Mk = v.Mf + (size_t)k * N;
    int off[NT];
#pragma unroll
    for (int j = 0; j < NT; j++) { off[j] = ls[j] * N; b1[j] = INFINITY; b2[j] = INFINITY; bi[j] = 0x7fffffff; }
    for (int t0 = 0; t0 < cnt; t0 += 32 * U) {
        int id[U]; bool ok[U]; float mk[U], x[U][NT];
#pragma unroll
        for (int u = 0; u < U; u++) { const int t = t0 + 32 * u + lane; ok[u] = t < cnt; id[u] = ok[u] ? (int)mem[t] : 0; }
#pragma unroll
        for (int u = 0; u < U; u++) {
            mk[u] = ok[u] ? Mk[id[u]] : 0.f;
#pragma unroll
            for (int j = 0; j < NT; j++) x[u][j] = ok[u] ? v.Mf[off[j] + id[u]] : INFINITY;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
                const float d = x[u][j] - mk[u];
                if (d < b1[j]) { b2[j] = b1[j]; b1[j] = d; bi[j] = id[u]; } else if (d < b2[j]) b2[j] = d;
            }
    }
}

// the warp-wide verdict of a screening pass for one target l: a unique fp32 winner clear of every runner-up is the fp64
// winner (its row goes to wi[k][l], the exact difference follows), anything closer is left to the exact scan
__device__ __forceinline__ unsigned screen_resolve(SolverSmem& sm, int k, int l, float b1, float b2, int bi) {
    const int lane = threadIdx.x & 31;
    const float m1 = funkey(__reduce_min_sync(0xffffffffu, fkey(b1)));
    const float thr = m1 + SCREEN_EPS;
    const unsigned cand = __ballot_sync(0xffffffffu, b1 <= thr);
    const unsigned close2 = __ballot_sync(0xffffffffu, b2 <= thr);
    if (cand == 0u) {                     // class k is empty
        if (lane == 0) { sm.w[k][l] = INFINITY; sm.wi[k][l] = -1; }
        return 0u;
    }
    if ((cand & (cand - 1)) == 0u && close2 == 0u) {
        const int win = __shfl_sync(0xffffffffu, bi, __ffs(cand) - 1);
        if (lane == 0) sm.wi[k][l] = win;          // w[k][l] follows in warp_rescan
        return 0u;
    }
    return 1u << l;
}

// One warp recomputes w[k][l], wi[k][l] for the classes l in `mask` from the member list of class k.
// FP64 issues at a small fraction of the FP32 rate on this part and the fp64 matrix lives in global memory, so with the
// fp32 copy (`screen`) a pass over the members finds the fp32 minimum and runner-up of every target:
// when the runner-up is more than SCREEN_EPS away the fp32 winner is the fp64 winner, and its exact fp64 difference is
// computed afterwards by lane l for all targets at once (one round trip to global memory per class).  Only targets
// with a near-tie take the exact scan over all members.
__device__ void warp_rescan(SolverSmem& sm, const SolverViews& v, const double* __restrict__ M, int N, int K, int k, unsigned mask,
                            bool screen) {
    const int lane = threadIdx.x & 31;
    const int cnt = sm.cnt[k];
    const uint16_t* mem = v.members + (size_t)k * N;
    mask &= ~(1u << k);
    unsigned exact_mask = screen ? 0u : mask;        // targets that need the exact scan
    if (screen) {
        unsigned todo = mask;
        while (todo) {                               // warp-uniform
            if (__popc(todo) > 4) {                  // 8 targets per pass, two members per lane in flight
                int ls[8]; int nl = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (todo) { ls[j] = __ffs(todo) - 1; todo &= todo - 1; nl = j + 1; } else ls[j] = ls[0];
                }
                float b1[8], b2[8]; int bi[8];
                screen_pass<8, 2>(v, N, k, cnt, ls, b1, b2, bi);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < nl) exact_mask |= screen_resolve(sm, k, ls[j], b1[j], b2[j], bi[j]);      // warp-uniform
            } else {                                 // up to 4 targets, four members per lane in flight
                int ls[4]; int nl = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (todo) { ls[j] = __ffs(todo) - 1; todo &= todo - 1; nl = j + 1; } else ls[j] = ls[0];
                }
                float b1[4], b2[4]; int bi[4];
                screen_pass<4, 4>(v, N, k, cnt, ls, b1, b2, bi);
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j < nl) exact_mask |= screen_resolve(sm, k, ls[j], b1[j], b2[j], bi[j]);      // warp-uniform
            }
        }
        __syncwarp();
        const unsigned direct = mask & ~exact_mask;
        if (lane < K && ((direct >> lane) & 1u)) {
            const int i = sm.wi[k][lane];
            if (i >= 0) sm.w[k][lane] = __dsub_rn(M[(size_t)i * K + lane], M[(size_t)i * K + k]);
        }
    }
    mask = exact_mask;
    while (mask) {                                   // warp-uniform; up to 4 target classes per pass
        int ls[4]; int nl = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (mask) { ls[j] = __ffs(mask) - 1; mask &= mask - 1; nl = j + 1; } else ls[j] = ls[0];
        }
        // candidates are compared as order-preserving integer keys
        unsigned long long bestk[4]; int besti[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { bestk[j] = KEY_INF; besti[j] = 0x7fffffff; }
        for (int t = lane; t < cnt; t += 32) {
            const int i = mem[t];
            const double* row = M + (size_t)i * K;
            const double mk = row[k];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j < nl) {
                    const unsigned long long key = dkey(__dsub_rn(row[ls[j]], mk));
                    // the list is unordered: ties resolve to the lowest row index explicitly
                    if (key < bestk[j] || (key == bestk[j] && i < besti[j])) { bestk[j] = key; besti[j] = i; }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j >= nl) break;                        // warp-uniform
            const unsigned long long key = bestk[j];
            const unsigned hi = (unsigned)(key >> 32);
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned lo = hi == mhi ? (unsigned)key : 0xffffffffu;
            const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
            const unsigned cand = (hi == mhi && lo == mlo) ? (unsigned)besti[j] : 0x7fffffffu;
            const unsigned mi = __reduce_min_sync(0xffffffffu, cand);
            if (lane == 0) {
                const int l = ls[j];
                if (mi == 0x7fffffffu) { sm.w[k][l] = INFINITY; sm.wi[k][l] = -1; }
                else { sm.w[k][l] = dunkey(((unsigned long long)mhi << 32) | mlo); sm.wi[k][l] = (int)mi; }
            }
        }
    }
    __syncwarp();
}

// warp 0: shortest path from the over-full classes to the cheapest under-full class
__device__ void find_path(SolverSmem& sm, int K) {
    int lane = threadIdx.x & 31;
    int l = lane & 15, half = lane >> 4;
    bool surplus = l < K && sm.cnt[l] > sm.b[l];
    bool deficit = l < K && sm.cnt[l] < sm.b[l];
    unsigned def_mask = __ballot_sync(0xffffffffu, deficit && half == 0);
    if (def_mask == 0) { if (lane == 0) sm.path_len = 0; __syncwarp(); return; }
    // Bellman-Ford, lane l owns node l; each round relaxes only from the nodes whose label changed in
    // the previous round (warp-uniform loop over the set bits), lowest source index wins ties
    double dist = surplus ? 0.0 : INFINITY;
    int pred = -1;
    if (half == 0) sm.dist[l] = dist;
    unsigned changed = __ballot_sync(0xffffffffu, surplus && half == 0) & 0xffffu;
    __syncwarp();
    for (int round = 0; round < KP && changed; round++) {
        double nd = dist; int np = pred;
        for (unsigned m = changed; m; m &= m - 1) {
            const int k = __ffs(m) - 1;
            const double cand = sm.dist[k] + sm.w[k][l];        // w[k][k] = +inf
            if (cand < nd) { nd = cand; np = k; }
        }
        const bool improved = half == 0 && nd < dist;
        changed = __ballot_sync(0xffffffffu, improved) & 0xffffu;
        __syncwarp();
        if (improved) { dist = nd; pred = np; sm.dist[l] = nd; }
        __syncwarp();
    }
    if (half == 0) sm.pred[l] = pred;
    __syncwarp();
    // cheapest under-full class (lowest index on ties)
    unsigned long long key = (deficit && half == 0) ? dkey(sm.dist[l]) : KEY_INF;
    unsigned hi = (unsigned)(key >> 32);
    unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    unsigned lo = hi == mhi ? (unsigned)key : 0xffffffffu;
    unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
    unsigned cand = (hi == mhi && lo == mlo && half == 0 && deficit) ? (unsigned)l : 0xffu;
    unsigned t = __reduce_min_sync(0xffffffffu, cand);
    if (lane == 0) {
        if (t >= (unsigned)K || !(sm.dist[t] < INFINITY)) { sm.status |= ST_NO_PATH; sm.path_len = 0; }
        else {
            int rev[KP + 1]; int len = 0; int node = (int)t;
            while (node >= 0 && len <= KP) { rev[len++] = node; node = sm.pred[node]; }
            if (len > KP || len < 2) { sm.status |= ST_PATH_OVERFLOW; sm.path_len = 0; }
            else {
                for (int q = 0; q < len; q++) sm.path[q] = rev[len - 1 - q];
                sm.path_len = len;
            }
        }
    }
    __syncwarp();
}

// Price search (step 1 of ot_solve_kernel), KK = number of classes (8 or 16).
// Lane l (< KK) of EVERY warp keeps class l's price, step, last sign and best price in registers, and all warps make
// the same update from the same counts, so a round costs one block barrier: a row's cheapest class comes from a
// compare tree over KK independent loads, rows are counted per thread in a packed 4-bit histogram, summed over the
// warp with REDUX, published per warp and read back by every warp.
template <int KK>
__device__ __forceinline__ void price_search(SolverSmem& sm, const float* __restrict__ Mf, const double* __restrict__ M_global,
                                             int N, int dual_iters, float step0, bool m_in_smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float my_price = lane < KP ? sm.pricef[lane] : 0.f, my_best = my_price, my_step = step0;
    int my_prev = 0, best_resid = 0x7fffffff;
    const int my_b = lane < KK ? sm.b[lane] : 0;
    for (int it = 0; it <= dual_iters; it++) {
        float pr[KK];
#pragma unroll
        for (int l = 0; l < KK; l++) pr[l] = __shfl_sync(0xffffffffu, my_price, l);
        unsigned acc[KP / 2];
#pragma unroll
        for (int w = 0; w < KP / 2; w++) acc[w] = 0u;
        unsigned long long h = 0ull; int pending = 0;
#pragma unroll 2
        for (int i = tid; i < N; i += SOLVER_THREADS) {
            float x[KK]; int idx[KK];
            if (m_in_smem) {
#pragma unroll
                for (int l = 0; l < KK; l++) x[l] = Mf[l * N + i];
            } else {
#pragma unroll
                for (int l = 0; l < KK; l++) x[l] = (float)M_global[(size_t)i * KK + l];
            }
#pragma unroll
            for (int l = 0; l < KK; l++) { x[l] -= pr[l]; idx[l] = l; }
#pragma unroll
            for (int st = 1; st < KK; st *= 2)
#pragma unroll
                for (int l = 0; l < KK; l += 2 * st)
                    if (x[l + st] < x[l]) { x[l] = x[l + st]; idx[l] = idx[l + st]; }      // strict: the lower class wins ties
            h += 1ull << (4 * idx[0]);
            if (++pending == 15) {                 // 4-bit fields are full: spill into the 16-bit accumulators
#pragma unroll
                for (int w = 0; w < KP / 2; w++) acc[w] += (unsigned)((h >> (8 * w)) & 0xF) | ((unsigned)((h >> (8 * w + 4)) & 0xF) << 16);
                h = 0ull; pending = 0;
            }
        }
#pragma unroll
        for (int w = 0; w < KK / 2; w++) {
            acc[w] += (unsigned)((h >> (8 * w)) & 0xF) | ((unsigned)((h >> (8 * w + 4)) & 0xF) << 16);
            acc[w] = __reduce_add_sync(0xffffffffu, acc[w]);           // N <= 65535: a 16-bit field cannot overflow
        }
        if (lane == 0) {
#pragma unroll
            for (int w = 0; w < KK / 2; w++) sm.whist[it & 1][warp][w] = acc[w];
        }
        __syncthreads();
        int cnt = 0;
        if (lane < KK) {
#pragma unroll
            for (int w = 0; w < SOLVER_WARPS; w++) cnt += (int)((sm.whist[it & 1][w][lane >> 1] >> ((lane & 1) * 16)) & 0xFFFFu);
        }
        const int err = lane < KK ? my_b - cnt : 0;
        const int resid = (int)(__reduce_add_sync(0xffffffffu, (unsigned)(err < 0 ? -err : err)) >> 1);
        if (resid < best_resid) { best_resid = resid; my_best = my_price; }
        if (resid == 0 || it == dual_iters) break;                     // same decision in every thread
        if (lane < KK) {
            const int sg = err > 0 ? 1 : (err < 0 ? -1 : 0);
            if (sg * my_prev < 0) my_step *= 0.5f; else if (sg * my_prev > 0) my_step *= 1.2f;
            my_prev = sg;
            my_price += my_step * (float)sg;                           // too few rows -> cheaper class
        }
    }
    if (warp == 0 && lane < KP) sm.best_pricef[lane] = my_best;
    __syncthreads();
}

// Register-resident variant of the price search for problems with at most RPT rows per thread (N <= RPT * 512, RPT <= 4):
// the thread's rows are loaded once as integer search keys (cost_key) and stay in registers for all rounds; a row's
// cheapest class is one integer minimum (min_key), the per-thread class histogram uses 4-bit fields of one 64-bit word
// (<= 4 rows per thread), widened to 8-bit fields for the four REDUX warp sums (<= 128 rows per class and warp), and the K
// price keys reach the lanes through a per-warp shared-memory row instead of K shuffles.
template <int KK, int RPT>
__device__ __forceinline__ void price_search_reg(SolverSmem& sm, const float* __restrict__ Mf, int N, int dual_iters, float step0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float my_price = lane < KP ? sm.pricef[lane] : 0.f, my_best = my_price, my_step = step0;
    int my_prev = 0, best_resid = 0x7fffffff;
    const int my_b = lane < KK ? sm.b[lane] : 0;
    int x[RPT][KK];
    bool have[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) {
        const int i = tid + r * SOLVER_THREADS;
        have[r] = i < N;
#pragma unroll
        for (int l = 0; l < KK; l++) x[r][l] = have[r] ? cost_key(Mf[l * N + i], l) : 0;
    }
    int* wp = reinterpret_cast<int*>(sm.wprice[warp]);
    const int rword = ((lane >> 3) << 1) | (lane & 1), rsh = ((lane & 7) >> 1) * 8;      // where lane l finds class l in acc[]
    int rb = 0;                                                                          // counter buffer of this round
    for (int it = 0; it <= dual_iters; it++) {
        const int my_pk = price_key(my_price);
        if (lane < KK) wp[lane] = my_pk;
        __syncwarp();
        int pk[KK];
#pragma unroll
        for (int l = 0; l < KK; l += 4) {
            const int4 q = *reinterpret_cast<const int4*>(wp + l);
            pk[l] = q.x; pk[l + 1] = q.y; pk[l + 2] = q.z; pk[l + 3] = q.w;
        }
        __syncwarp();
        unsigned long long h = 0ull;                         // 4-bit count per class
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            int y[KK];
#pragma unroll
            for (int l = 0; l < KK; l++) y[l] = x[r][l] - pk[l];
            const int m = min_key<KK>(y);
            h += (have[r] ? 1ull : 0ull) << ((m & 15) << 2);
        }
        // widen: even classes / odd classes of the low and high halves -> 8-bit fields, one REDUX each
        const unsigned lo = (unsigned)h, hi = (unsigned)(h >> 32);
        unsigned acc[4];
        acc[0] = __reduce_add_sync(0xffffffffu, lo & 0x0F0F0F0Fu);            // classes 0,2,4,6
        acc[1] = __reduce_add_sync(0xffffffffu, (lo >> 4) & 0x0F0F0F0Fu);     // classes 1,3,5,7
        if (KK > 8) {
            acc[2] = __reduce_add_sync(0xffffffffu, hi & 0x0F0F0F0Fu);        // classes 8,10,12,14
            acc[3] = __reduce_add_sync(0xffffffffu, (hi >> 4) & 0x0F0F0F0Fu); // classes 9,11,13,15
        }
        // lane l adds the warp's count of class l to the round's counter (one shared-memory atomic per warp; every lane then
        // reads ONE word -- summing the 16 per-warp histograms in every lane of every warp was 40 % of the round's instructions)
        if (lane < KK) {
            unsigned a = acc[0];
            a = rword == 1 ? acc[1] : a;
            if (KK > 8) { a = rword == 2 ? acc[2] : a; a = rword == 3 ? acc[3] : a; }
            const int mine = (int)((a >> rsh) & 0xFFu);
            if (mine) atomicAdd(&sm.rcnt[rb][lane], mine);
        }
        __syncthreads();
        const int cnt = lane < KK ? sm.rcnt[rb][lane] : 0;
        { const int rz = rb == 0 ? 2 : rb - 1; if (warp == 0 && lane < KP) sm.rcnt[rz][lane] = 0; }     // the buffer of round it + 2
        rb = rb == 2 ? 0 : rb + 1;
        const int err = lane < KK ? my_b - cnt : 0;
        const int resid = (int)(__reduce_add_sync(0xffffffffu, (unsigned)(err < 0 ? -err : err)) >> 1);
        if (resid < best_resid) { best_resid = resid; my_best = price_of_key(my_pk); }      // the price the rows actually saw
        if (resid == 0 || it == dual_iters) break;                     // same decision in every thread
        if (lane < KK) {
            const int sg = err > 0 ? 1 : (err < 0 ? -1 : 0);
            if (sg * my_prev < 0) my_step *= 0.5f; else if (sg * my_prev > 0) my_step *= 1.2f;
            my_prev = sg;
            my_price += my_step * (float)sg;                           // too few rows -> cheaper class
        }
    }
    if (warp == 0 && lane < KP) sm.best_pricef[lane] = my_best;
    __syncthreads();
}

// Price search for problems whose fp32 cost copy does not fit in shared memory (modes 1 / 2 of ot_solve_kernel).  Streaming
// the whole copy from L2 every round made the search L2-bound (100 CTAs read the same 0.5 MB per round: ~6 TB/s of L2 hits),
// so the rows are split three ways: the first 4 x 512 rows stay in registers for all rounds (as in price_search_reg), the
// next `slice_rows` rows sit in the shared memory the member lists leave free (class-major slice), the rest is read from
// global memory.  Histogram and price update as in price_search.
//
// Trust region (rounds >= pivot_round).  Once the steps are small, the prices stay inside a box  |p_l - p0_l| <= R_l  around
// their current value p0 (R_l = trust * step_l) for many rounds.  A row whose cheapest class s under the box's most
// adverse corner still beats every other class,  key_s - (p0_s - R_s) < key_l - (p0_l + R_l)  for all l != s, takes class s
// under EVERY price vector of the box: it is counted once (`frozen`) and skipped.  One classification sweep over the rows
// outside the registers separates those rows from the ACTIVE ones, whose keys are packed into the shared-memory slice; the
// following rounds sweep the registers and the active list only -- everything on chip, no L2 traffic -- and yield exactly the
// counts a full sweep would.  A price that leaves its box triggers a new classification around the current prices.  If
// the active rows do not fit the slice (nearly tied costs everywhere), the search falls back to full sweeps from L2.
// The search is a heuristic either way: the assignment is made exact by the repair steps that follow it.
template <int KK>
__device__ __forceinline__ void price_search_hybrid(SolverSmem& sm, const int* __restrict__ Mk,
                                                    int* __restrict__ slice, int slice_rows, int N, int dual_iters, float step0,
                                                    int pivot_arg) {
    constexpr int RPT = 4;
    const int pivot_round = pivot_arg & 0xff;            // first round of the trust region
    const float trust = (float)((pivot_arg >> 8) & 0xff);    // R_l = trust * step_l
    const int sample_rounds = min((pivot_arg >> 16) & 0x7f, pivot_round);   // first rounds that sweep the on-chip rows only
    const bool asym = ((pivot_arg >> 23) & 1) != 0;                         // asymmetric boxes (A/B runs only)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float my_price = lane < KP ? sm.pricef[lane] : 0.f, my_best = my_price, my_step = step0;
    int my_prev = 0, best_resid = 0x7fffffff;
    const int my_b = lane < KK ? sm.b[lane] : 0;
    const int reg_rows = min(N, RPT * SOLVER_THREADS);
    const int cap = slice_rows;                          // rows the slice can hold (static prefix, later the active list)
    int static_rows = min(slice_rows, N - reg_rows);     // rows of the static prefix currently in the slice
    auto key_at = [&](int l, int i) { return __ldg(Mk + (size_t)l * N + i); };      // (a float fallback in these loops cost 25 %: code size)
    int x[RPT][KK];
    bool have[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) {
        const int i = tid + r * SOLVER_THREADS;
        have[r] = i < N;
#pragma unroll
        for (int l = 0; l < KK; l++) x[r][l] = have[r] ? key_at(l, i) : 0;
    }
    auto fill_static = [&]() {                           // the rows behind the registers, as many as the slice holds
        static_rows = min(slice_rows, N - reg_rows);
        for (int e = tid; e < KK * static_rows; e += SOLVER_THREADS) {
            const int l = e / static_rows, j = e - l * static_rows;
            slice[l * cap + j] = key_at(l, reg_rows + j);
        }
    };
    fill_static();
    if (tid == 0) { sm.n_active = 0; sm.tr_overflow = 0; }
    __syncthreads();
    // The first rounds only steer the prices towards the right region: they sweep the on-chip rows alone (registers + slice, 2/3
    // of the rows at N = 7720) against the demand scaled to that share -- no L2 traffic, half the time of a full sweep.  Only
    // the exact rounds that follow can set the best prices.  (Sampling all rounds up to the pivot shrinks the steps on the
    // sample's sign flips and costs more repair steps than it saves: 2.0 instead of 0.5 per draw.)
    const bool sample_start = sample_rounds + 8 <= dual_iters && reg_rows + static_rows < N;
    const float sample_share = (float)(reg_rows + static_rows) / (float)N;
    int next_pivot = pivot_round;    // the next round that may classify (moved back after an overflow)
    // trust-region state, identical in every thread (lane l < KK holds class l's pivot key, radius and frozen count)
    int tr_mode = 0;                 // 0: full / sampled sweeps, 1: inside a trust region (registers + active list)
    int n_active = 0, my_pk0 = 0, my_rp = 0, my_rm = 0, my_frozen = 0;     // box of class l: [pk0 - rm, pk0 + rp]
    int rb = 0;                      // counter buffer of this round (sm.rcnt)
    int n_pivots = 0;                // classifications so far
    for (int it = 0; it <= dual_iters; it++) {
        const int my_pk = price_key(my_price);
        // ---- classification sweep (no price update): pivot = the current prices
        if (it >= next_pivot && (tr_mode == 0 || __any_sync(0xffffffffu, lane < KK && (my_pk - my_pk0 > my_rp || my_pk0 - my_pk > my_rm)))) {
            my_pk0 = my_pk;
            // Tried and kept as A/B knobs only (FG_OT_ASYM=1): a box that reaches 4x further in the direction a class's price is
            // drifting, and a radius that doubles on every re-classification.  Both cut the number of classifications (5-8 per draw
            // on peaked, class-biased probabilities) but push the active list over the slice's capacity: 0.42 -> 0.53 ms at
            // N = 7720 on the bench step's probabilities, 0.43 -> 0.57 ms inside the 8-GPU step.
            {
                const int base = max(price_key(trust * my_step), 16);
                const int ahead = asym ? 4 * base : base, behind = asym ? max(base / 2, 16) : base;
                my_rp = my_prev > 0 ? ahead : (my_prev < 0 ? behind : base);
                my_rm = my_prev < 0 ? ahead : (my_prev > 0 ? behind : base);
            }
            n_pivots++;
#ifdef FG_OT_PROFILE
            if (tid == 0) { sm.prof_pivots++; sm.prof_pivot_last = it; }
#endif
            int pkz[KK];
#pragma unroll
            for (int l = 0; l < KK; l++) pkz[l] = __shfl_sync(0xffffffffu, my_pk0 + my_rp, l);
            if (tid == 0) { sm.n_active = 0; sm.tr_overflow = 0; }
            __syncthreads();
            unsigned acc[KP / 2];
#pragma unroll
            for (int w = 0; w < KP / 2; w++) acc[w] = 0u;
            unsigned long long h = 0ull; int pending = 0;
            const int rest = N - reg_rows;
            for (int j0 = 0; j0 < rest; j0 += SOLVER_THREADS) {          // uniform trip count: the loop body shuffles
                const int j = j0 + tid;
                const bool valid = j < rest;
                const int i = reg_rows + (valid ? j : 0);
                int y[KK];
#pragma unroll
                for (int l = 0; l < KK; l++) y[l] = key_at(l, i) - pkz[l];
                const int m1 = min_key<KK>(y);
                const int sc = m1 & 15;
#pragma unroll
                for (int l = 0; l < KK; l++) y[l] = l == sc ? 0x7f000000 : y[l];
                const int m2 = min_key<KK>(y);
                const int rs = __shfl_sync(0xffffffffu, my_rp + my_rm, sc);          // width of the winner's box
                const bool frozen = valid && m2 > m1 + rs + 16;
                const bool active = valid && !frozen;
                if (frozen) {
                    h += 1ull << (sc << 2);
                    if (++pending == 15) {
#pragma unroll
                        for (int w = 0; w < KP / 2; w++) acc[w] += (unsigned)((h >> (8 * w)) & 0xF) | ((unsigned)((h >> (8 * w + 4)) & 0xF) << 16);
                        h = 0ull; pending = 0;
                    }
                }
                const unsigned am = __ballot_sync(0xffffffffu, active);
                int base = 0;
                if (lane == 0 && am) base = atomicAdd(&sm.n_active, __popc(am));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (active) {
                    const int slot = base + __popc(am & ((1u << lane) - 1u));
                    if (slot < cap) {
#pragma unroll
                        for (int l = 0; l < KK; l++) slice[l * cap + slot] = key_at(l, i);
                    } else sm.tr_overflow = 1;
                }
            }
#pragma unroll
            for (int w = 0; w < KK / 2; w++) {
                acc[w] += (unsigned)((h >> (8 * w)) & 0xF) | ((unsigned)((h >> (8 * w + 4)) & 0xF) << 16);
                acc[w] = __reduce_add_sync(0xffffffffu, acc[w]);
            }
            if (lane == 0) {
#pragma unroll
                for (int w = 0; w < KK / 2; w++) sm.whist[it & 1][warp][w] = acc[w];
            }
            __syncthreads();
            my_frozen = 0;
            if (lane < KK) {
#pragma unroll
                for (int w = 0; w < SOLVER_WARPS; w++) my_frozen += (int)((sm.whist[it & 1][w][lane >> 1] >> ((lane & 1) * 16)) & 0xFFFFu);
            }
            n_active = sm.n_active;
            tr_mode = sm.tr_overflow ? 0 : 1;
#ifdef FG_OT_PROFILE
            if (tid == 0) sm.prof_rounds = tr_mode == 0 ? -n_active : n_active;
#endif
            static_rows = 0;                                  // the slice no longer holds the static prefix
            __syncthreads();                                  // every thread has read n_active / tr_overflow; whist is free again
            if (tr_mode == 0) {
                // the active rows do not fit (the steps, hence the radii, are still large, or the costs are nearly tied): back to
                // full sweeps with the static prefix restored, and a new attempt two rounds later
                fill_static();
                next_pivot = it + 2;
                __syncthreads();
            }
        }
        const bool sampled = sample_start && it < sample_rounds;   // block-uniform
        int pk[KK];
#pragma unroll
        for (int l = 0; l < KK; l++) pk[l] = __shfl_sync(0xffffffffu, my_pk, l);
        unsigned acc[KP / 2];
#pragma unroll
        for (int w = 0; w < KP / 2; w++) acc[w] = 0u;
        unsigned long long h = 0ull; int pending = 0;
        auto count_row = [&](int (&y)[KK]) {
#pragma unroll
            for (int l = 0; l < KK; l++) y[l] -= pk[l];
            const int m = min_key<KK>(y);
            h += 1ull << ((m & 15) << 2);
            if (++pending == 15) {                 // 4-bit fields are full: spill into the 16-bit accumulators
#pragma unroll
                for (int w = 0; w < KP / 2; w++) acc[w] += (unsigned)((h >> (8 * w)) & 0xF) | ((unsigned)((h >> (8 * w + 4)) & 0xF) << 16);
                h = 0ull; pending = 0;
            }
        };
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            if (have[r]) {
                int y[KK];
#pragma unroll
                for (int l = 0; l < KK; l++) y[l] = x[r][l];
                count_row(y);
            }
        }
        // shared-memory rows: the static prefix (full sweeps) or the active list (trust region)
        const int smem_rows = tr_mode == 1 ? n_active : static_rows;
        for (int j = tid; j < smem_rows; j += SOLVER_THREADS) {
            int y[KK];
#pragma unroll
            for (int l = 0; l < KK; l++) y[l] = slice[l * cap + j];
            count_row(y);
        }
        if (tr_mode != 1 && !sampled) {
#pragma unroll 2
            for (int i = reg_rows + static_rows + tid; i < N; i += SOLVER_THREADS) {
                int y[KK];
#pragma unroll
                for (int l = 0; l < KK; l++) y[l] = key_at(l, i);
                count_row(y);
            }
        }
#pragma unroll
        for (int w = 0; w < KK / 2; w++) {
            acc[w] += (unsigned)((h >> (8 * w)) & 0xF) | ((unsigned)((h >> (8 * w + 4)) & 0xF) << 16);
            acc[w] = __reduce_add_sync(0xffffffffu, acc[w]);           // N <= 65535: a 16-bit field cannot overflow
        }
        // lane l adds the warp's count of class l to the round's counter (see price_search_reg)
        if (lane < KK) {
            unsigned a = acc[0];
#pragma unroll
            for (int w = 1; w < KK / 2; w++) a = (lane >> 1) == w ? acc[w] : a;
            const int mine = (int)((a >> ((lane & 1) * 16)) & 0xFFFFu);
            if (mine) atomicAdd(&sm.rcnt[rb][lane], mine);
        }
        __syncthreads();
        int cnt = tr_mode == 1 ? my_frozen : 0;
        if (lane < KK) cnt += sm.rcnt[rb][lane];
        { const int rz = rb == 0 ? 2 : rb - 1; if (warp == 0 && lane < KP) sm.rcnt[rz][lane] = 0; }     // the buffer of round it + 2
        rb = rb == 2 ? 0 : rb + 1;
        int sg;
        if (sampled) {
            const float errf = (float)my_b * sample_share - (float)cnt;                      // against the scaled demand; half a row of slack
            sg = errf > 0.5f ? 1 : (errf < -0.5f ? -1 : 0);
        } else {
            const int err = lane < KK ? my_b - cnt : 0;
            const int resid = (int)(__reduce_add_sync(0xffffffffu, (unsigned)(err < 0 ? -err : err)) >> 1);
            if (resid < best_resid) { best_resid = resid; my_best = price_of_key(my_pk); }      // the price the rows actually saw
            if (resid == 0 || it == dual_iters) break;                     // same decision in every thread
            sg = err > 0 ? 1 : (err < 0 ? -1 : 0);
        }
        if (lane < KK) {
            if (sg * my_prev < 0) my_step *= 0.5f; else if (sg * my_prev > 0) my_step *= 1.2f;
            my_prev = sg;
            my_price += my_step * (float)sg;                           // too few rows -> cheaper class
        }
    }
    if (warp == 0 && lane < KP) sm.best_pricef[lane] = my_best;
    __syncthreads();
}

// Final assignment under the best prices + first scan of the class graph for the large modes (cost copy in global memory).
// warp_rescan walks a class's member list and gathers that row's costs: every 4-byte gather pulls its own 32-byte sector
// from L2, and with 100 CTAs doing the same the first scan of all classes took ~115-180 k cycles at N = 7720.  The same
// minima come out of ONE coalesced sweep over the rows (lane = row, like the price search) plus a pass over a short list:
//   sweep:  class s of the row (the assignment), then for every target l the difference d = M[i,l] - M[i,s] is compared with
//           the running minimum smin[s][l]; a row within SCREEN_EPS of the minimum it reads updates it (atomicMin) and is a
//           CANDIDATE: its costs are appended to a list in the shared memory the price search no longer needs.  The running
//           minimum only decreases, so every row within SCREEN_EPS of the FINAL minimum is in the list (~N/6 rows: the
//           first 512 and O(log) record-breakers per pair);
//   list:   rows of the list within SCREEN_EPS of the final smin[s][l] are counted and the lowest of them kept.
// A pair with exactly one such row has its fp64 winner (same rule as screen_resolve); the others (near-ties) go to the exact
// member-list scan of warp_rescan, and so does everything when the list overflows.
template <int KK>
__device__ __forceinline__ void sweep_assign_scan(SolverSmem& sm, const SolverViews& v, const double* __restrict__ M, int N,
                                                  float* __restrict__ cand, int cand_cap) {
    constexpr int CW = KK + 1;                       // list entry: KK costs + the row index; odd stride = conflict-free
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < KP * (KP + 1); e += SOLVER_THREADS) {
        (&sm.smin[0][0])[e] = 0xffffffffu; (&sm.scnt[0][0])[e] = 0; (&sm.sidx[0][0])[e] = 0x7fffffff;
    }
    if (tid < KP) sm.sexact[tid] = 0u;
    if (tid == 0) { sm.n_active = 0; sm.tr_overflow = 0; }
    __syncthreads();
#ifdef FG_OT_PROFILE
    long long tq[5]; tq[0] = clock64();
#define FG_SWEEP_MARK(q) do { __syncthreads(); tq[q] = clock64(); } while (0)
#else
#define FG_SWEEP_MARK(q) do {} while (0)
#endif
    float pf[KK];
#pragma unroll
    for (int l = 0; l < KK; l++) pf[l] = sm.best_pricef[l];
    for (int i0 = 0; i0 < N; i0 += SOLVER_THREADS) {           // uniform trip count: the body uses warp votes
        const int i = i0 + tid;
        const bool valid = i < N;
        const int il = valid ? i : 0;
        float xv[KK];
#pragma unroll
        for (int l = 0; l < KK; l++) xv[l] = __ldg(v.Mf + (size_t)l * N + il);
        // fp32 screen: the prices are fp32 values, so |fp32 reduced cost - fp64 reduced cost| <= 2 ulp(M) ~ 5e-7;
        // when the runner-up is further away than that the fp64 argmin is the fp32 argmin
        float b1 = INFINITY, b2 = INFINITY; int s = 0;
#pragma unroll
        for (int l = 0; l < KK; l++) {
            const float x = xv[l] - pf[l];
            if (x < b1) { b2 = b1; b1 = x; s = l; } else if (x < b2) b2 = x;
        }
        if (valid && !(b2 - b1 > 8e-6f)) {
            const double* row = M + (size_t)i * KK;
            double bv = INFINITY; s = 0;
            for (int l = 0; l < KK; l++) { const double x = __dsub_rn(row[l], sm.price[l]); if (x < bv) { bv = x; s = l; } }
        }
        bool att = false;
        if (valid) {
            v.sigma[i] = (uint8_t)s;
            const int slot = atomicAdd(&sm.cnt[s], 1);
            v.members[(size_t)s * N + slot] = (uint16_t)i;
            v.pos[i] = (uint16_t)slot;
            float xs = xv[0];
#pragma unroll
            for (int l = 1; l < KK; l++) xs = l == s ? xv[l] : xs;
#pragma unroll
            for (int l = 0; l < KK; l++) {
                // one flat predicate per target (a short-circuit `l != s && ...` made the compiler nest 16 divergent regions
                // that never reconverged: 7x the instructions); 0xffffffff reads as NaN, so the first rows always enter
                const float d = xv[l] - xs;
                const float cur = funkey(sm.smin[s][l]);
                const bool hit = (l != s) & !(d > cur + SCREEN_EPS);
                att |= hit;
                if (hit) atomicMin(&sm.smin[s][l], fkey(d));
            }
        }
        const unsigned am = __ballot_sync(0xffffffffu, att);
        if (am) {                                     // warp-uniform
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.n_active, __popc(am));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (att) {
                const int slot = base + __popc(am & ((1u << lane) - 1u));
                if (slot < cand_cap) {
#pragma unroll
                    for (int l = 0; l < KK; l++) cand[slot * CW + l] = xv[l];
                    cand[slot * CW + KK] = __int_as_float(i);
                } else sm.tr_overflow = 1;
            }
        }
    }
    __syncthreads();
    FG_SWEEP_MARK(1);
    const bool overflow = sm.tr_overflow != 0;            // block-uniform
    const int n_cand = overflow ? 0 : sm.n_active;
    for (int j = tid; j < n_cand; j += SOLVER_THREADS) {
        const float* e = cand + j * CW;
        const int i = __float_as_int(e[KK]);
        const int s = v.sigma[i];
        const float xs = e[s];
#pragma unroll
        for (int l = 0; l < KK; l++) {
            const float cur = funkey(sm.smin[s][l]);
            const bool hit = (l != s) & (e[l] - xs <= cur + SCREEN_EPS);
            if (hit) { atomicAdd(&sm.scnt[s][l], 1); atomicMin(&sm.sidx[s][l], i); }
        }
    }
    __syncthreads();
    FG_SWEEP_MARK(2);
    if (tid < KP * KP) {
        const int k = tid >> 4, l = tid & 15;
        if (k < KK && l < KK && k != l) {
            if (sm.cnt[k] == 0) { sm.w[k][l] = INFINITY; sm.wi[k][l] = -1; }
            else if (!overflow && sm.scnt[k][l] == 1) {
                const int i = sm.sidx[k][l];
                sm.wi[k][l] = i;
                sm.w[k][l] = __dsub_rn(M[(size_t)i * KK + l], M[(size_t)i * KK + k]);
            } else atomicOr(&sm.sexact[k], 1u << l);
        }
    }
    __syncthreads();
    FG_SWEEP_MARK(3);
    const int warp = tid >> 5;
    for (int k = warp; k < KK; k += SOLVER_WARPS)
        if (sm.sexact[k]) warp_rescan(sm, v, M, N, KK, k, sm.sexact[k], !overflow ? false : true);       // warp-uniform
#ifdef FG_OT_PROFILE
    FG_SWEEP_MARK(4);
    if (tid == 0 && ((blockIdx.x % 50) == 0 || tq[4] - tq[0] > 120000)) {
        int ne = 0; for (int k = 0; k < KK; k++) ne += __popc(sm.sexact[k]);
        printf("sweepprof blk %d: sweep %lld list(%d rows) %lld resolve %lld exact(%d pairs) %lld cycles\n", blockIdx.x, tq[1] - tq[0], n_cand, tq[2] - tq[1], tq[3] - tq[2], ne, tq[4] - tq[3]);
    }
#endif
}

// One CTA solves one transport problem exactly.
//
//  1. price search (all warps):  `dual_iters` rounds of sign-based dual ascent on the K class prices
//     (each round: every row takes its cheapest class under the current prices, the class counts are
//     compared with the demand, a price moves up/down by its own adaptive step -- grow 1.2x while the sign of
//     the count error persists, halve when it flips).  ANY price vector yields an assignment that is optimal
//     for its own class sizes, so this only has to get the sizes CLOSE to the demand; the best prices seen
//     are kept.  It replaces ~N/5 sequential repair steps by a few fully parallel sweeps.
//  2. exact finish (warp 0):  successive shortest paths on the class graph repair the remaining
//     |count - demand|_1 / 2 units one at a time; shortest path, moves, list updates and the re-scan of the
//     classes that lost a row are all warp-synchronous, so a repair step costs no block barrier besides the
//     one that tells the other warps whether the loop is over.
//
// demand: from `demand_by_value` when hist == nullptr, else hist[blockIdx.x].  prices_in == nullptr starts
// from zero prices (the greedy assignment).  prices_out receives feasible optimal prices of the final state.
//
// MODE selects where the two large views live (solver_mode): 0 = member lists and the fp32 cost copy in shared memory
// (N*K up to ~36k), 1 = member lists in shared memory, the fp32 copy read from global memory (it is written once by
// the cost kernel and stays in L2), 2 = both in global memory (`members_global` holds [gridDim.x][K][N] entries) --
// the shape of the 8-GPU weak-scaling batch (8192 rows x 16 classes).
template <int MODE>
__global__ void __launch_bounds__(SOLVER_THREADS)
ot_solve_kernel(const double* __restrict__ M_global, const float* __restrict__ Mf_global, int N, int K,
                const double* __restrict__ prices_in, double* __restrict__ prices_out, int dual_iters, double step0,
                Demand demand_by_value, const int* __restrict__ hist,
                int32_t* __restrict__ assign_out, int32_t* __restrict__ counts,
                int* __restrict__ status, int status_slot, int m_in_smem, uint16_t* __restrict__ members_global, int slice_rows,
                const int* __restrict__ Mk_global, int pivot_round) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    SolverSmem& sm = *reinterpret_cast<SolverSmem*>(dyn_smem);
    SolverViews v;
    v.sigma = dyn_smem + SOLVER_CTRL;
    v.pos = reinterpret_cast<uint16_t*>(dyn_smem + solver_off_pos(N));
    if (MODE == 2) v.members = members_global + (size_t)blockIdx.x * K * N;
    else v.members = reinterpret_cast<uint16_t*>(dyn_smem + solver_off_members(N));
    if (MODE == 0) v.Mf = reinterpret_cast<float*>(dyn_smem + solver_off_M(N, K));
    else v.Mf = const_cast<float*>(Mf_global);
    uint8_t* sigma = v.sigma;
    const int tid = threadIdx.x, warp = tid >> 5;
    // FP64 issues at a small fraction of the FP32 rate on this part, so the price search (a heuristic: any
    // prices are admissible) runs in fp32 on a shared-memory copy of the costs; only the final assignment and
    // the exact repair steps touch the fp64 matrix.
    const double* M = M_global;
#ifdef FG_OT_PROFILE
    long long tprof[8]; int nprof = 0;
#define FG_MARK() do { __syncthreads(); tprof[nprof++] = clock64(); } while (0)
#else
#define FG_MARK() do {} while (0)
#endif
    FG_MARK();
    // class-major copy: lane i reads Mf[l*N + i], conflict-free (a row-major row per lane is a 16-way conflict)
    if (MODE == 0 && m_in_smem) {
        if (Mf_global && (N * K) % 4 == 0) {       // precomputed by the cost kernel: straight 16-byte copies
            const float4* src = reinterpret_cast<const float4*>(Mf_global);
            float4* dst = reinterpret_cast<float4*>(v.Mf);
            for (int e = tid; e < N * K / 4; e += SOLVER_THREADS) dst[e] = __ldg(src + e);
        } else {
            for (int e = tid; e < N * K; e += SOLVER_THREADS) { const int i = e / K, l = e - i * K; v.Mf[l * N + i] = (float)M_global[e]; }
        }
    }
    if (tid < KP) {
        sm.cnt[tid] = 0;
        sm.b[tid] = hist ? hist[blockIdx.x * KP + tid] : demand_by_value.b[tid];
        if (tid >= K) sm.b[tid] = 0;
        sm.price[tid] = (prices_in && tid < K) ? prices_in[tid] : 0.0;
        sm.pricef[tid] = (float)sm.price[tid];
        sm.best_pricef[tid] = sm.pricef[tid];
        sm.step[tid] = (float)step0;
        sm.prev_sign[tid] = 0;
    }
    if (tid == 0) { sm.status = 0; sm.path_len = 0; sm.best_resid = 0x7fffffff; sm.stop = 0; }
    if (tid < 3 * KP) (&sm.rcnt[0][0])[tid] = 0;
#ifdef FG_OT_PROFILE
    if (tid == 0) { sm.prof_pivots = 0; sm.prof_pivot_last = -1; sm.prof_rounds = 0; }
#endif
    for (int e = tid; e < KP * KP; e += SOLVER_THREADS) { sm.w[e / KP][e % KP] = INFINITY; sm.wi[e / KP][e % KP] = -1; }
    __syncthreads();
    if (tid == 0) {
        long long tot = 0;
        for (int k = 0; k < K; k++) { if (sm.b[k] < 0) sm.status |= ST_BAD_DEMAND; tot += sm.b[k]; }
        if (tot != N) sm.status |= ST_BAD_DEMAND;
    }
    FG_MARK();

    // ---- 1. price search (fp32)
    if (MODE == 0 && m_in_smem && N <= 2 * SOLVER_THREADS) {
        if (K == 16) price_search_reg<16, 2>(sm, v.Mf, N, dual_iters, (float)step0);
        else price_search_reg<8, 2>(sm, v.Mf, N, dual_iters, (float)step0);
    } else if (MODE == 0 && m_in_smem && N <= 4 * SOLVER_THREADS) {
        if (K == 16) price_search_reg<16, 4>(sm, v.Mf, N, dual_iters, (float)step0);
        else price_search_reg<8, 4>(sm, v.Mf, N, dual_iters, (float)step0);
    } else if (MODE != 0 && m_in_smem && Mk_global) {
        // the shared memory behind the last view holds a slice of the cost copy during the search
        int* slice = reinterpret_cast<int*>(dyn_smem + (MODE == 1 ? solver_off_M(N, K) : solver_off_members(N)));
        if (K == 16) price_search_hybrid<16>(sm, Mk_global, slice, slice_rows, N, dual_iters, (float)step0, pivot_round);
        else price_search_hybrid<8>(sm, Mk_global, slice, slice_rows, N, dual_iters, (float)step0, pivot_round);
    } else if (K == 16) price_search<16>(sm, v.Mf, M_global, N, dual_iters, (float)step0, m_in_smem != 0);
    else price_search<8>(sm, v.Mf, M_global, N, dual_iters, (float)step0, m_in_smem != 0);
    if (tid < KP) sm.price[tid] = (double)sm.best_pricef[tid];
    __syncthreads();
    FG_MARK();
    // final assignment under the best prices + member lists (list order is arbitrary; no result depends on it)
    const bool coalesced_scan = MODE != 0 && m_in_smem && !(sm.status & ST_BAD_DEMAND);       // block-uniform
    if (coalesced_scan) {
        // candidate list: the shared memory behind the last view, which held the slice of the cost copy during the search
        float* cand = reinterpret_cast<float*>(dyn_smem + (MODE == 1 ? solver_off_M(N, K) : solver_off_members(N)));
        const int cand_cap = (slice_rows * K) / (K + 1);
        if (K == 16) sweep_assign_scan<16>(sm, v, M, N, cand, cand_cap); else sweep_assign_scan<8>(sm, v, M, N, cand, cand_cap);
    } else {
        for (int i = tid; i < N; i += SOLVER_THREADS) {
            int s = 0;
            bool exact = true;
            if (m_in_smem) {
                // fp32 screen: the prices are fp32 values, so |fp32 reduced cost - fp64 reduced cost| <= 2 ulp(M) ~ 5e-7;
                // when the runner-up is further away than that the fp64 argmin is the fp32 argmin
                float b1 = INFINITY, b2 = INFINITY;
                for (int l = 0; l < K; l++) {
                    const float x = v.Mf[l * N + i] - sm.best_pricef[l];
                    if (x < b1) { b2 = b1; b1 = x; s = l; } else if (x < b2) b2 = x;
                }
                exact = !(b2 - b1 > 8e-6f);
            }
            if (exact) {
                const double* row = M + (size_t)i * K;
                double bv = INFINITY; s = 0;
                for (int l = 0; l < K; l++) { const double x = __dsub_rn(row[l], sm.price[l]); if (x < bv) { bv = x; s = l; } }
            }
            sigma[i] = (uint8_t)s;
            const int slot = atomicAdd(&sm.cnt[s], 1);
            v.members[(size_t)s * N + slot] = (uint16_t)i;
            v.pos[i] = (uint16_t)slot;
        }
    }
    __syncthreads();
    FG_MARK();
    const unsigned all_mask = (1u << K) - 1u;
    int iters = 0;
    if (!(sm.status & ST_BAD_DEMAND)) {
        if (!coalesced_scan)
            for (int k = warp; k < K; k += SOLVER_WARPS) warp_rescan(sm, v, M, N, K, k, all_mask, m_in_smem != 0);
        __syncthreads();
        FG_MARK();
        const int max_iters = N + KP;
        // One repair step = shortest path + row moves (warp 0), barrier, re-scan of the classes that lost a row (ONE WARP PER
        // PATH CLASS, in parallel: the scans are independent and each is a chain of L2 round trips in the large modes),
        // barrier, fold of the arrivals (warp 0).
        for (;; iters++) {
            if (warp == 0) {
                const int lane = tid;
                find_path(sm, K);
                int len = sm.path_len;
                if (len > 0 && iters >= max_iters) { if (lane == 0) { sm.status |= ST_ITER_CAP; sm.path_len = 0; } len = 0; }
                if (len > 0) {
                    for (int e = 0; e + 1 < len; e++) {
                        // the row leaving u, and the targets l whose minimum w[u][l] it was holding
                        const int u = sm.path[e], item = sm.wi[u][sm.path[e + 1]];
                        const unsigned need = __ballot_sync(0xffffffffu, lane < K && lane != u && sm.wi[u][lane] == item);
                        if (lane == 0) { sm.moved[e] = item; sm.need[e] = need; }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        for (int e = 0; e + 1 < len; e++) {
                            const int u = sm.path[e], to = sm.path[e + 1], item = sm.moved[e];
                            // unlink from u (swap with the last member), append to `to`
                            const int p = v.pos[item], lastpos = sm.cnt[u] - 1;
                            const uint16_t last = v.members[(size_t)u * N + lastpos];
                            v.members[(size_t)u * N + p] = last; v.pos[last] = (uint16_t)p;
                            sm.cnt[u] = lastpos;
                            const int q = sm.cnt[to];
                            v.members[(size_t)to * N + q] = (uint16_t)item; v.pos[item] = (uint16_t)q;
                            sm.cnt[to] = q + 1;
                            sigma[item] = (uint8_t)to;
                        }
                        if (MODE == 2) __threadfence_block();      // the member lists live in global memory: publish them to the CTA
                    }
                }
            }
            __syncthreads();
            const int len = sm.path_len;                   // block-uniform
            if (len == 0) break;
            if (warp < len - 1) warp_rescan(sm, v, M, N, K, sm.path[warp], sm.need[warp], m_in_smem != 0);     // len - 1 <= K - 1 < SOLVER_WARPS
            __syncthreads();
            if (warp == 0) {
                const int lane = tid;
                // rows that arrived in path[e+1]: fold their outgoing differences into the minima
                if (lane < K) {
                    const int l = lane;
                    for (int e = 0; e + 1 < len; e++) {
                        const int to = sm.path[e + 1], item = sm.moved[e];
                        if (l == to) continue;
                        const double* row = M + (size_t)item * K;
                        const double d = __dsub_rn(row[l], row[to]);
                        const double cur = sm.w[to][l];
                        if (sm.wi[to][l] < 0 || d < cur || (d == cur && item < sm.wi[to][l])) { sm.w[to][l] = d; sm.wi[to][l] = item; }
                    }
                }
                __syncwarp();
            }
        }
        if (tid == 0 && status) atomicAdd(&status[status_slot], iters);
    }
    __syncthreads();
    FG_MARK();

    // outputs
    if (assign_out) for (int i = tid; i < N; i += SOLVER_THREADS) assign_out[(size_t)blockIdx.x * N + i] = sigma[i];
    if (counts) for (int i = tid; i < N; i += SOLVER_THREADS) atomicAdd(&counts[(size_t)i * K + sigma[i]], 1);
    if (prices_out) {
        // feasible prices for the final state: v_l = min(0, min_k v_k + w[k][l]) (difference constraints)
        if (tid < 32) {
            int lane = tid;
            if (lane < KP) sm.dist[lane] = 0.0;
            __syncwarp();
            for (int round = 0; round < KP; round++) {
                double best = lane < KP ? sm.dist[lane] : 0.0;
                if (lane < K) for (int k = 0; k < K; k++) { double c = sm.dist[k] + sm.w[k][lane]; if (k != lane && c < best) best = c; }
                bool ch = lane < K && best < sm.dist[lane];
                unsigned any = __ballot_sync(0xffffffffu, ch);
                __syncwarp();
                if (ch) sm.dist[lane] = best;
                __syncwarp();
                if (!any) break;
            }
            if (lane < KP) prices_out[lane] = sm.dist[lane];
        }
    }
    if (tid == 0 && sm.status && status) atomicOr(&status[0], sm.status);
#ifdef FG_OT_PROFILE
    FG_MARK();
    if (tid == 0)
        printf("otprof blk %d N %d K %d grid %d pivots %d (last at %d) active %d: copy %lld search %lld assign %lld rescan %lld ssp(%d) %lld out %lld total %lld cycles\n", blockIdx.x, N, K, gridDim.x, sm.prof_pivots, sm.prof_pivot_last, sm.prof_rounds,
               tprof[1] - tprof[0], tprof[2] - tprof[1], tprof[3] - tprof[2], tprof[4] - tprof[3], iters, tprof[5] - tprof[4], tprof[6] - tprof[5], tprof[6] - tprof[0]);
#endif
}

// ---------------------------------------------------------------------------------------------
// epilogue E3:1534-1565 (+ thresholding E3:2022-2023): one thread per image row
template <typename T>
__device__ __forceinline__ float seq_sum_T(const float* tp, const int* cols, int ncols) {
    float s = tp[cols[0]];
    for (int q = 1; q < ncols; q++) s = __fadd_rn(s, tp[cols[q]]);
    return round_to<T>(s);
}

template <typename T>
__global__ void ot_targets_kernel(const int32_t* __restrict__ counts, const int* __restrict__ pos, int n_all, int n_valid, int K,
                                  float threshold, long long* __restrict__ tg, T* __restrict__ ug, long long* __restrict__ tr,
                                  T* __restrict__ ur, long long* __restrict__ ta, T* __restrict__ ua) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    int p = n_valid > 0 ? pos[i] : -1;
    if (p >= n_valid) p = -1;            // the caller under-reported n_valid (flagged by compact_kernel): stay inside counts[]
    const int n_attr = K == 16 ? 3 : 2;
    long long* tout[3] = {tg, tr, ta};
    T* uout[3] = {ug, ur, ua};
    if (p < 0) {
        for (int a = 0; a < n_attr; a++) { if (tout[a]) tout[a][i] = -1; if (uout[a]) uout[a][i] = from_f32<T>(-1.f); }
        return;
    }
    // total = target_probs[0,:].sum() in the probs dtype
    float total = 0.f;
    for (int j = 0; j < K; j++) total = __fadd_rn(total, round_to<T>((float)counts[j]));
    total = round_to<T>(total);
    float tp[KP];
    for (int j = 0; j < K; j++) tp[j] = round_to<T>(__fdiv_rn(round_to<T>((float)counts[(size_t)p * K + j]), total));
    float thr = round_to<T>(threshold);
    for (int a = 0; a < n_attr; a++) {
        int width = (a == 1) ? 4 : 2;
        float best = 0.f; int besti = 0;
        for (int c = 0; c < width; c++) {
            int cols[8]; int nc = 0;
            if (K == 8) {
                if (a == 0) { for (int q = 0; q < 4; q++) cols[nc++] = c * 4 + q; }              // E3:1538-1541
                else { cols[nc++] = c; cols[nc++] = c + 4; }                                    // E3:1542-1549
            } else {
                if (a == 0) { for (int q = 0; q < 8; q++) cols[nc++] = c * 8 + q; }              // E4:1572-1575
                else if (a == 1) { cols[nc++] = 2 * c; cols[nc++] = 2 * c + 1; cols[nc++] = 2 * c + 8; cols[nc++] = 2 * c + 9; }  // E4:1576-1583
                else { for (int q = 0; q < 8; q++) cols[nc++] = 2 * q + c; }                     // E4:1584-1589
            }
            float m = seq_sum_T<T>(tp, cols, nc);
            if (c == 0 || m > best) { best = m; besti = c; }
        }
        float unc = round_to<T>(__fsub_rn(1.f, best));
        long long t = besti;
        if (threshold >= 0.f && unc > thr) t = -1;
        if (tout[a]) tout[a][i] = t;
        if (uout[a]) uout[a][i] = from_f32<T>(unc);
    }
}

// ---------------------------------------------------------------------------------------------
// E1: binomial CDF tables + rank split
struct RankWs { int* nvalid; double* cdf0; double* cdf1; size_t total; };
static RankWs rank_carve(void* base, int n_all) {
    RankWs w; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += fg_align_up(bytes, 256); return (char*)base + o; };
    w.nvalid = (int*)take(sizeof(int));
    w.cdf0 = (double*)take((size_t)(n_all + 1) * sizeof(double));
    w.cdf1 = (double*)take((size_t)(n_all + 1) * sizeof(double));
    w.total = off;
    return w;
}

__device__ __forceinline__ double binom_pmf(int k, int n, double p) {
    if (p <= 0.0) return k == 0 ? 1.0 : 0.0;
    if (p >= 1.0) return k == n ? 1.0 : 0.0;
    double lg = lgamma((double)n + 1.0) - lgamma((double)k + 1.0) - lgamma((double)(n - k) + 1.0)
              + (double)k * log(p) + (double)(n - k) * log1p(-p);
    return exp(lg);
}

// cdf0[k] = P[Bin(N, ratio) <= k], cdf1[k] = P[Bin(N, 1-ratio) <= k]; one CTA, chunked scan
template <typename T>
__global__ void __launch_bounds__(1024)
binom_tables_kernel(const T* __restrict__ probs, int n_all, double ratio, int* __restrict__ nvalid_out,
                    double* __restrict__ cdf0, double* __restrict__ cdf1) {
    __shared__ int s_cnt;
    __shared__ double warp_tot[2][32];
    __shared__ double carry[2];
    if (threadIdx.x == 0) { s_cnt = 0; carry[0] = 0.0; carry[1] = 0.0; }
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < n_all; i += 1024)
        local += (to_f32(probs[2 * i]) != -1.f && to_f32(probs[2 * i + 1]) != -1.f) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, local);
    __syncthreads();
    const int N = s_cnt;
    if (threadIdx.x == 0) *nvalid_out = N;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start <= N; start += 1024) {
        int k = start + threadIdx.x;
        double v[2] = {0.0, 0.0};
        if (k <= N) { v[0] = binom_pmf(k, N, ratio); v[1] = binom_pmf(k, N, 1.0 - ratio); }
        for (int q = 0; q < 2; q++) {
            double s = v[q];
            for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
            if (lane == 31) warp_tot[q][warp] = s;
            v[q] = s;
        }
        __syncthreads();
        if (warp == 0) {
            for (int q = 0; q < 2; q++) {
                double s = warp_tot[q][lane];
                for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
                warp_tot[q][lane] = s;
            }
        }
        __syncthreads();
        if (k <= N) {
            double c0 = carry[0] + (warp ? warp_tot[0][warp - 1] : 0.0) + v[0];
            double c1 = carry[1] + (warp ? warp_tot[1][warp - 1] : 0.0) + v[1];
            cdf0[k] = c0 > 1.0 ? 1.0 : c0;
            cdf1[k] = c1 > 1.0 ? 1.0 : c1;
        }
        __syncthreads();
        if (threadIdx.x == 0) { carry[0] += warp_tot[0][31]; carry[1] += warp_tot[1][31]; }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
rank_split_kernel(const T* __restrict__ probs, int n_all, double ratio, float threshold,
                  const int* __restrict__ nvalid, const double* __restrict__ cdf0, const double* __restrict__ cdf1,
                  long long* __restrict__ targets, T* __restrict__ unc) {
    __shared__ float tile[256];
    __shared__ uint8_t tile_ok[256];
    int i = blockIdx.x * 256 + threadIdx.x;
    float p0 = 0.f, p1 = 0.f; bool ok = false;
    if (i < n_all) { p0 = to_f32(probs[2 * i]); p1 = to_f32(probs[2 * i + 1]); ok = p0 != -1.f && p1 != -1.f; }
    int rank = 0;
    for (int start = 0; start < n_all; start += 256) {
        int j = start + threadIdx.x;
        float q0 = 0.f, q1 = 0.f;
        if (j < n_all) { q0 = to_f32(probs[2 * j]); q1 = to_f32(probs[2 * j + 1]); }
        __syncthreads();
        tile[threadIdx.x] = q1;
        tile_ok[threadIdx.x] = (j < n_all && q0 != -1.f && q1 != -1.f) ? 1 : 0;
        __syncthreads();
        int lim = min(256, n_all - start);
        for (int t = 0; t < lim; t++) {
            float q = tile[t];
            int jj = start + t;
            rank += (tile_ok[t] && (q < p1 || (q == p1 && jj < i))) ? 1 : 0;
        }
    }
    if (i >= n_all) return;
    if (!ok) { targets[i] = -1; if (unc) unc[i] = from_f32<T>(-1.f); return; }
    const int N = *nvalid;
    // (rank >= N*ratio) is evaluated by torch in float32: int64 tensor vs python float (E1:1420)
    float cut = (float)((double)N * ratio);
    long long t = ((float)rank >= cut) ? 1 : 0;
    double u = t == 1 ? 1.0 - cdf1[rank] : cdf0[rank];                        // E1:1425-1440
    float uf = round_to<T>((float)u);
    if (threshold >= 0.f && uf > round_to<T>(threshold)) t = -1;              // E1:1835
    targets[i] = t;
    if (unc) unc[i] = from_f32<T>(uf);
}

// shared memory of one solver CTA: control block, assignment bytes, list positions (+ member lists, + the fp32 cost
// matrix, while they fit); see ot_solve_kernel for the three modes
static int solver_mode(int N, int K) {
    if (solver_off_M(N, K) + (size_t)N * K * sizeof(float) <= 220 * 1024) return 0;
    if (solver_off_M(N, K) <= 220 * 1024) return 1;
    return 2;
}
// modes 1 / 2: rows of the cost copy kept in the shared memory the views leave free during the price search
static int solver_slice_rows(int N, int K) {
    const int mode = solver_mode(N, K);
    if (mode == 0) return 0;
    const size_t base = mode == 1 ? solver_off_M(N, K) : solver_off_members(N);
    long rows = ((long)220 * 1024 - (long)base) / (long)(K * sizeof(float));
    const long want = (long)N - 4 * SOLVER_THREADS;
    if (rows > want) rows = want;
    return rows > 0 ? (int)(rows & ~3L) : 0;
}
static int solver_smem_bytes(int N, int K) {
    const int mode = solver_mode(N, K);
    if (mode == 0) return (int)(solver_off_M(N, K) + (size_t)N * K * sizeof(float));
    const size_t base = mode == 1 ? solver_off_M(N, K) : solver_off_members(N);
    return (int)(base + (size_t)solver_slice_rows(N, K) * K * sizeof(float));
}
static size_t solver_members_bytes(int N, int K, int S) {      // global member lists of mode 2
    return solver_mode(N, K) == 2 ? (size_t)(S > 0 ? S : 1) * K * N * sizeof(uint16_t) : 0;
}

static int solver_prepare(int N, int K) {
    if (N > 65535) return FG_ERR_LIMIT;           // member lists index rows with 16 bits
    if (solver_smem_bytes(N, K) > 227 * 1024) return FG_ERR_LIMIT;
    cudaError_t e = cudaFuncSetAttribute(ot_solve_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ot_solve_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ot_solve_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return e == cudaSuccess ? FG_OK : (int)e;
}

// fp32 class-major copy of a row-major fp64 cost matrix (fg_ot_solve_single in modes 1 / 2; the plan path gets it from
// cost_hist_kernel)
__global__ void mf_from_m_kernel(const double* __restrict__ M, float* __restrict__ Mf, int* __restrict__ Mk, int N, int K) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < N * K) { const int i = e / K, l = e - i * K; const float c = (float)M[e]; Mf[(size_t)l * N + i] = c; Mk[(size_t)l * N + i] = cost_key(c, l); }
}

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static double env_dbl(const char* name, double dflt) { const char* e = getenv(name); return e ? atof(e) : dflt; }
// round of the price search at which the trust region starts (price_search_hybrid); a tuning knob, results do not depend on it
#define SEARCH_PIVOT_ROUND (env_int("FG_OT_PIVOT", 18) | (env_int("FG_OT_TRUST", 6) << 8) | (env_int("FG_OT_SAMPLE", 6) << 16) | (env_int("FG_OT_ASYM", 0) << 23))

#define FG_SOLVE_LAUNCH(N_, K_, GRID_, ST_, MK_, ...)                                                                     \
    do {                                                                                                              \
        const int mode_ = solver_mode(N_, K_);                                                                        \
        const int smem_ = solver_smem_bytes(N_, K_);                                                                  \
        const int slice_ = solver_slice_rows(N_, K_);                                                                 \
        const int pivot_ = SEARCH_PIVOT_ROUND;                                                                        \
        if (mode_ == 0) ot_solve_kernel<0><<<GRID_, SOLVER_THREADS, smem_, ST_>>>(__VA_ARGS__, slice_, MK_, pivot_);   \
        else if (mode_ == 1) ot_solve_kernel<1><<<GRID_, SOLVER_THREADS, smem_, ST_>>>(__VA_ARGS__, slice_, MK_, pivot_); \
        else ot_solve_kernel<2><<<GRID_, SOLVER_THREADS, smem_, ST_>>>(__VA_ARGS__, slice_, MK_, pivot_);              \
    } while (0)

// expected demand for n rows: largest-remainder rounding of n*q
static void expected_demand(int n, int K, Demand* d) {
    double q[KP]; double frac[KP]; long long tot = 0;
    for (int j = 0; j < KP; j++) { q[j] = 0.0; d->b[j] = 0; }
    if (K == 8) for (int j = 0; j < 8; j++) q[j] = 1.0 / 8.0;
    else for (int j = 0; j < 16; j++) q[j] = 0.5 * 0.25 * ((j & 1) ? 0.25 : 0.75);
    for (int j = 0; j < K; j++) { double x = q[j] * n; d->b[j] = (int)floor(x); frac[j] = x - floor(x); tot += d->b[j]; }
    for (long long r = tot; r < n; r++) {          // fewer than K units remain
        int best = 0;
        for (int j = 1; j < K; j++) if (frac[j] > frac[best]) best = j;
        d->b[best]++; frac[best] = -1.0 - (double)(r - tot);
    }
}

// price-search schedule (rounds, first step) of a solve that starts from zero prices, i.e. from the greedy assignment
// defaults; the FG_OT_* environment variables exist for tuning runs only (results do not depend on them)
#define BASE_DUAL_ITERS env_int("FG_OT_BASE_ITERS", 40)
#define BASE_STEP0 env_dbl("FG_OT_BASE_STEP", 0.03)
#define DRAW_DUAL_ITERS env_int("FG_OT_DRAW_ITERS", 40)
#define DRAW_STEP0 env_dbl("FG_OT_DRAW_STEP", 0.03)
#define WARM_DUAL_ITERS 24            // fg_ot_solve_single: second solve warm-started from the first one's prices
#define WARM_STEP0 0.01

// base problem: the expected demand; leaves its optimal prices in w.prices
static int launch_base(const double* M, const float* Mf, int N, int K, OtWs& w, cudaStream_t st) {
    int rc = solver_prepare(N, K);
    if (rc) return rc;
    Demand d; expected_demand(N, K, &d);
    FG_SOLVE_LAUNCH(N, K, 1, st, (Mf ? w.Mk : nullptr), M, Mf, N, K, nullptr, w.prices, BASE_DUAL_ITERS, BASE_STEP0, d, nullptr, nullptr, nullptr, w.status, 2,
                    1, w.members);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

// ---------------------------------------------------------------------------------------------
// E6 (race only): generate_dynamic_targets_race, exp-6-debias-race/1-main-debias.py:1413-1482.
// The demands are the enumerated compositions (n1..n4) of N that cover 95 % of the multinomial mass (host side,
// api.py); each is an exact transport problem on the SAME solver (the 4 classes are padded to 8 with unreachable
// ones: cost RACE_PAD, demand 0), and the plans are accumulated with their weights in the reference's order.
constexpr double RACE_PAD = 1.0e3;

// cost E6:1461 = ot.dist(probs, eye(4), metric="euclidean") as POT 0.9.3 evaluates it on the numpy backend
// (ot/utils.py euclidean_distances): c = -2 * X.Y^T ; c += |x|^2 ; c += |y|^2 ; sqrt(max(c, 0)), where |x|^2 is the
// fp32 einsum of the fp32 probabilities (numpy reduces the 4 products pairwise) and X.Y^T is exact in fp64.
template <typename T>
__global__ void __launch_bounds__(256)
race_cost_kernel(const T* __restrict__ pr, const int* __restrict__ idx, int N, double* __restrict__ M, float* __restrict__ Mf, int* __restrict__ Mk) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= N) return;
    const int i = idx[r];
    float x[4];
    for (int q = 0; q < 4; q++) x[q] = to_f32(pr[4 * i + q]);
    const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(x[0], x[0]), __fmul_rn(x[1], x[1])), __fadd_rn(__fmul_rn(x[2], x[2]), __fmul_rn(x[3], x[3])));
    for (int j = 0; j < 8; j++) {
        double c = RACE_PAD;
        if (j < 4) {
            c = __dadd_rn(__dmul_rn(-2.0, (double)x[j]), (double)a2);
            c = __dadd_rn(c, 1.0);
            c = __dsqrt_rn(c > 0.0 ? c : 0.0);
        }
        M[(size_t)r * 8 + j] = c;
        Mf[(size_t)j * N + r] = (float)c;
        Mk[(size_t)j * N + r] = cost_key((float)c, j);
    }
}

// target_probs += T * prob over the compositions in order (E6:1463-1466), row L1 normalisation (E6:1467), argmax and
// 1 - max (E6:1469-1472), scatter to all rows (E6:1474-1479), optional thresholding in the probs dtype (E6 call site)
template <typename T>
__global__ void race_accumulate_kernel(const int32_t* __restrict__ sigma, const double* __restrict__ weights, int S, int N,
                                       const int* __restrict__ pos, int n_all, float threshold,
                                       long long* __restrict__ targets, T* __restrict__ unc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    const int p = pos[i];
    if (p < 0 || p >= N) { targets[i] = -1; if (unc) unc[i] = from_f32<T>(-1.f); return; }
    double tp[4] = {0.0, 0.0, 0.0, 0.0};
    for (int s = 0; s < S; s++) {
        const int c = sigma[(size_t)s * N + p];
        const double w = weights[s];
#pragma unroll
        for (int q = 0; q < 4; q++) tp[q] = __dadd_rn(tp[q], c == q ? w : 0.0);      // T is 0/1: the other entries add +0.0
    }
    const double nrm = __dadd_rn(__dadd_rn(__dadd_rn(fabs(tp[0]), fabs(tp[1])), fabs(tp[2])), fabs(tp[3]));
    int best = 0; double bv = __ddiv_rn(tp[0], nrm);
    for (int q = 1; q < 4; q++) { const double v = __ddiv_rn(tp[q], nrm); if (v > bv) { bv = v; best = q; } }
    const float uf = round_to<T>((float)__dsub_rn(1.0, bv));
    long long t = best;
    if (threshold >= 0.f && uf > round_to<T>(threshold)) t = -1;
    targets[i] = t;
    if (unc) unc[i] = from_f32<T>(uf);
}

struct RaceWs { OtWs ot; int32_t* sigma; size_t total; };
static RaceWs race_carve(void* base, int n_all, int S) {
    RaceWs w;
    w.ot = ot_carve(base, n_all, 8, S);
    w.sigma = (int32_t*)((char*)base + w.ot.total);
    w.total = w.ot.total + fg_align_up((size_t)(S > 0 ? S : 1) * n_all * sizeof(int32_t), 256);
    return w;
}

}  // namespace

extern "C" size_t fg_ot_workspace_bytes(int n_all, int K, int S) {
    if (n_all < 0 || (K != 8 && K != 16) || S < 0) return 0;
    return ot_carve(nullptr, n_all > 0 ? n_all : 1, K, S).total;
}

extern "C" int fg_ot_plan_counts(const void* probs_gender, const void* probs_race, const void* probs_age, int n_all,
                                 const void* rand_gender, const void* rand_race, const void* rand_age, int S, int n_valid,
                                 int32_t* counts, void* workspace, size_t workspace_bytes, int dtype, void* stream) {
    const int K = probs_age ? 16 : 8;
    if (n_all < 0 || S < 0 || n_valid < 0 || n_valid > n_all) return FG_ERR_INVALID_ARG;
    if (n_all > 0 && (!probs_gender || !probs_race)) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_ot_workspace_bytes(n_all, K, S)) return FG_ERR_WORKSPACE;
    OtWs w = ot_carve(workspace, n_all > 0 ? n_all : 1, K, S);
    cudaStream_t st = fg_stream(stream);
    if (n_all == 0) return cudaMemsetAsync(w.status, 0, 4 * sizeof(int), st) == cudaSuccess ? FG_OK : FG_ERR_INVALID_ARG;
    // the status words are cleared by compact_kernel and the plan counts by cost_hist_kernel (no memset launches)
    FG_DISPATCH_DTYPE(dtype, T,
        compact_kernel<T><<<1, 1024, 0, st>>>((const T*)probs_gender, (const T*)probs_race, n_all, n_valid, w.idx, w.pos, w.status));
    FG_LAUNCH_CHECK();
    if (n_valid == 0) return FG_OK;
    if (!counts || (S > 0 && (!rand_gender || !rand_race || (K == 16 && !rand_age)))) return FG_ERR_INVALID_ARG;
    int cost_blocks = (n_valid * K + 255) / 256;
    FG_DISPATCH_DTYPE(dtype, T,
        cost_hist_kernel<T><<<cost_blocks + S, 256, 0, st>>>((const T*)probs_gender, (const T*)probs_race, (const T*)probs_age,
            w.idx, n_valid, K, w.M, w.Mf, w.Mk, cost_blocks, (const T*)rand_gender, (const T*)rand_race, (const T*)rand_age, S, w.hist, counts));
    FG_LAUNCH_CHECK();
    if (S == 0) return FG_OK;
    // Every draw runs the whole price search from zero prices in its own CTA.  (A serial "base" solve of the expected
    // demand used to warm-start the draws: it saved them ~40% of their search rounds but cost a single-CTA launch
    // longer than the draws themselves.)
    int rc = solver_prepare(n_valid, K);
    if (rc) return rc;
    Demand none = {};
    FG_SOLVE_LAUNCH(n_valid, K, S, st, w.Mk, w.M, w.Mf, n_valid, K, nullptr, nullptr, DRAW_DUAL_ITERS, DRAW_STEP0, none, w.hist, nullptr, counts,
                    w.status, 3, 1, w.members);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_ot_targets(const int32_t* counts, const void* probs_gender, const void* probs_race, int n_all, int n_valid,
                             int K, float threshold,
                             int64_t* targets_gender, void* unc_gender, int64_t* targets_race, void* unc_race,
                             int64_t* targets_age, void* unc_age, void* workspace, size_t workspace_bytes,
                             int dtype, void* stream) {
    (void)probs_gender; (void)probs_race;
    if (n_all < 0 || n_valid < 0 || n_valid > n_all || (K != 8 && K != 16)) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_ot_workspace_bytes(n_all, K, 0)) return FG_ERR_WORKSPACE;
    if (n_valid > 0 && !counts) return FG_ERR_INVALID_ARG;
    if (n_all == 0) return FG_OK;
    OtWs w = ot_carve(workspace, n_all, K, 0);
    FG_DISPATCH_DTYPE(dtype, T,
        ot_targets_kernel<T><<<(n_all + 127) / 128, 128, 0, fg_stream(stream)>>>(counts, w.pos, n_all, n_valid, K, threshold,
            (long long*)targets_gender, (T*)unc_gender, (long long*)targets_race, (T*)unc_race, (long long*)targets_age, (T*)unc_age));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_ot_solve_single(const double* M, int n, int K, const int64_t* b_host, int32_t* assign,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0 || (K != 8 && K != 16) || !b_host || (n > 0 && (!M || !assign))) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_ot_workspace_bytes(n, K, 0)) return FG_ERR_WORKSPACE;
    if (n == 0) return FG_OK;
    Demand d = {};
    for (int j = 0; j < K; j++) { if (b_host[j] < 0 || b_host[j] > n) return FG_ERR_INVALID_ARG; d.b[j] = (int)b_host[j]; }
    OtWs w = ot_carve(workspace, n, K, 0);
    cudaStream_t st = fg_stream(stream);
    cudaError_t e = cudaMemsetAsync(w.status, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    const float* Mf = nullptr;                     // mode 0 builds its shared-memory copy from M
    if (solver_mode(n, K) != 0) {
        mf_from_m_kernel<<<(n * K + 255) / 256, 256, 0, st>>>(M, w.Mf, w.Mk, n, K);
        Mf = w.Mf;
    }
    int rc = launch_base(M, Mf, n, K, w, st);
    if (rc) return rc;
    FG_SOLVE_LAUNCH(n, K, 1, st, (Mf ? w.Mk : nullptr), M, Mf, n, K, w.prices, nullptr, WARM_DUAL_ITERS, WARM_STEP0, d, nullptr, assign, nullptr, w.status, 3,
                    1, w.members);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_ot_cost_matrix(const void* probs_gender, const void* probs_race, const void* probs_age, int n_all,
                                 int n_valid, double* M, void* workspace, size_t workspace_bytes, int dtype, void* stream) {
    const int K = probs_age ? 16 : 8;
    if (n_all <= 0 || n_valid < 0 || !probs_gender || !probs_race || !M) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_ot_workspace_bytes(n_all, K, 0)) return FG_ERR_WORKSPACE;
    OtWs w = ot_carve(workspace, n_all, K, 0);
    cudaStream_t st = fg_stream(stream);
    cudaError_t e = cudaMemsetAsync(w.status, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    FG_DISPATCH_DTYPE(dtype, T,
        compact_kernel<T><<<1, 1024, 0, st>>>((const T*)probs_gender, (const T*)probs_race, n_all, n_valid, w.idx, w.pos, w.status);
        if (n_valid > 0) cost_hist_kernel<T><<<(n_valid * K + 255) / 256, 256, 0, st>>>((const T*)probs_gender, (const T*)probs_race,
            (const T*)probs_age, w.idx, n_valid, K, M, nullptr, nullptr, (n_valid * K + 255) / 256, nullptr, nullptr, nullptr, 0, nullptr, nullptr));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" size_t fg_rank_binom_workspace_bytes(int n_all) {
    return rank_carve(nullptr, n_all > 0 ? n_all : 1).total;
}

extern "C" int fg_assign_rank_binom(const void* probs, int n_all, double target_ratio, float threshold,
                                    int64_t* targets, void* uncertainty, void* workspace, size_t workspace_bytes,
                                    int dtype, void* stream) {
    if (n_all < 0 || !(target_ratio >= 0.0 && target_ratio <= 1.0)) return FG_ERR_INVALID_ARG;
    if (n_all == 0) return FG_OK;
    if (!probs || !targets) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_rank_binom_workspace_bytes(n_all)) return FG_ERR_WORKSPACE;
    RankWs w = rank_carve(workspace, n_all);
    cudaStream_t st = fg_stream(stream);
    FG_DISPATCH_DTYPE(dtype, T,
        binom_tables_kernel<T><<<1, 1024, 0, st>>>((const T*)probs, n_all, target_ratio, w.nvalid, w.cdf0, w.cdf1);
        rank_split_kernel<T><<<(n_all + 255) / 256, 256, 0, st>>>((const T*)probs, n_all, target_ratio, threshold, w.nvalid,
                                                                  w.cdf0, w.cdf1, (long long*)targets, (T*)uncertainty));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" size_t fg_race_workspace_bytes(int n_all, int S) {
    if (n_all < 0 || S < 0) return 0;
    return race_carve(nullptr, n_all > 0 ? n_all : 1, S).total;
}

extern "C" int fg_assign_race_enumerated(const void* probs_race, int n_all, int n_valid, const int32_t* demands,
                                         const double* weights, int S, float threshold, int64_t* targets, void* uncertainty,
                                         void* workspace, size_t workspace_bytes, int dtype, void* stream) {
    if (n_all < 0 || n_valid < 0 || n_valid > n_all || S < 0) return FG_ERR_INVALID_ARG;
    if (n_all == 0) return FG_OK;
    if (!probs_race || !targets || (n_valid > 0 && (S == 0 || !demands || !weights))) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_race_workspace_bytes(n_all, S)) return FG_ERR_WORKSPACE;
    RaceWs w = race_carve(workspace, n_all, S);
    cudaStream_t st = fg_stream(stream);
    FG_DISPATCH_DTYPE(dtype, T,
        compact_kernel<T><<<1, 1024, 0, st>>>((const T*)nullptr, (const T*)probs_race, n_all, n_valid, w.ot.idx, w.ot.pos, w.ot.status));
    FG_LAUNCH_CHECK();
    if (n_valid > 0) {
        FG_DISPATCH_DTYPE(dtype, T,
            race_cost_kernel<T><<<(n_valid + 255) / 256, 256, 0, st>>>((const T*)probs_race, w.ot.idx, n_valid, w.ot.M, w.ot.Mf, w.ot.Mk));
        FG_LAUNCH_CHECK();
        int rc = solver_prepare(n_valid, 8);
        if (rc) return rc;
        Demand none = {};
        FG_SOLVE_LAUNCH(n_valid, 8, S, st, w.ot.Mk, w.ot.M, w.ot.Mf, n_valid, 8, nullptr, nullptr, DRAW_DUAL_ITERS, DRAW_STEP0, none, demands, w.sigma,
                        nullptr, w.ot.status, 3, 1, w.ot.members);
        FG_LAUNCH_CHECK();
    }
    FG_DISPATCH_DTYPE(dtype, T,
        race_accumulate_kernel<T><<<(n_all + 127) / 128, 128, 0, st>>>(w.sigma, weights, S, n_valid, w.ot.pos, n_all, threshold,
                                                                     (long long*)targets, (T*)uncertainty));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

/* Test hook: the E6 cost matrix, M [n_valid, 4] float64 in compacted row order. */
extern "C" int fg_race_cost_matrix(const void* probs_race, int n_all, int n_valid, double* M4, void* workspace,
                                   size_t workspace_bytes, int dtype, void* stream) {
    if (n_all <= 0 || n_valid < 0 || n_valid > n_all || !probs_race || !M4) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_race_workspace_bytes(n_all, 0)) return FG_ERR_WORKSPACE;
    RaceWs w = race_carve(workspace, n_all, 0);
    cudaStream_t st = fg_stream(stream);
    FG_DISPATCH_DTYPE(dtype, T,
        compact_kernel<T><<<1, 1024, 0, st>>>((const T*)nullptr, (const T*)probs_race, n_all, n_valid, w.ot.idx, w.ot.pos, w.ot.status);
        if (n_valid > 0) race_cost_kernel<T><<<(n_valid + 255) / 256, 256, 0, st>>>((const T*)probs_race, w.ot.idx, n_valid, w.ot.M, w.ot.Mf, w.ot.Mk));
    FG_LAUNCH_CHECK();
    if (n_valid > 0) {
        cudaError_t e = cudaMemcpy2DAsync(M4, 4 * sizeof(double), w.ot.M, 8 * sizeof(double), 4 * sizeof(double), n_valid,
                                          cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    return FG_OK;
}
