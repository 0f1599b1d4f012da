// Tensor-core (tcgen05 + TMEM) version of the head's dense projections for bf16 / fp16 inputs.
// Included by fg_head.cu inside its anonymous namespace.
//
//   forward   pre[m, d_hid]  = pooled[m, d_in] * W1^T + b1            (Linear(960,1280), E1:929 classifier[0])
//             part[t, m, k]  = sum_{n in tile t} hardswish(pre[m,n]) * W2[k,n]   (Hardswish + Linear(1280,K) fused
//                                                                      into the epilogue, one partial per N tile)
//   backward  g_pooled[m, d_in] = g_pre[m, d_hid] * W1                (frozen weights: no weight gradient)
//
// One CTA computes a 128 x 64 output tile: M = 128 rows live on the 128 TMEM lanes, the fp32 accumulator takes 64
// TMEM columns, K advances 64 elements per shared-memory stage (4 tcgen05.mma of K = 16 each).  Operands are staged
// by all 128 threads with 16-byte cp.async copies straight into the canonical no-swizzle UMMA layouts (8 x 16-byte
// core matrices; LBO = stride between core matrices along K, SBO = stride between 8-row groups along M/N): K-major
// for A and for the forward's W1 [N,K]; MN-major for the backward's W1 read as [K,N] (a core matrix is then 8 k-rows
// of 8 consecutive n, i.e. the 16-byte chunks of the row-major matrix as they are -- no transpose while staging).
// The ring is STAGES deep with DIST stages in flight, so the DRAM latency of the operands (the step's streaming
// kernels leave nothing of them in L2) is paid once, not per K step; completion of the MMAs that read a slot is
// tracked with tcgen05.commit on mbarriers.  The problem is tiny (2.5 GFLOP at m = 1024), so the kernel is built
// for low latency and few launches, not for peak tensor throughput.
#pragma once

namespace tc {

constexpr int BM = 128, BN = 64, BK = 64;
constexpr int THREADS = 128;
constexpr int A_STAGE = BM * BK * 2;            // bytes (16-bit operands)
constexpr int B_STAGE = BN * BK * 2;
constexpr int A_LBO = BM * 16, B_LBO = BN * 16; // K-direction stride between core matrices
constexpr int SBO = 128;                        // M/N-direction stride between 8-row groups
constexpr int TMEM_COLS = 64;
constexpr int STAGES = 4, DIST = 3;            // shared-memory ring depth, stages in flight (2 CTAs of 96 KB fit an SM)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout NONE [61,64)
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t instr_desc(int a_fmt, int b_fmt, int b_mn_major) {
    // cute::UMMA::InstrDescriptor: c_format F32 = 1 [4,6) | a_format [7,10) | b_format [10,13) | a_major [15] = K |
    // b_major [16]: 0 = K-major, 1 = MN-major | n_dim = N>>3 [17,23) | m_dim = M>>4 [24,29)
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {   // src_bytes = 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
                 :: "r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

template <typename T> struct Fmt;
template <> struct Fmt<__nv_bfloat16> { static constexpr int code = 1; };
template <> struct Fmt<__half> { static constexpr int code = 0; };
template <> struct Fmt<float> { static constexpr int code = 2; };          // TF32 operands (fp32 in shared memory)

// fp32 inputs run on the tensor cores as TF32 with an error-compensating split: x = hi + lo with hi = cvt.rna.tf32(x) and
// lo = x - hi (exact); the three passes  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  accumulate into the same TMEM tile, which keeps
// the result at ~2^-21 relative (plain TF32: ~2^-11 per product, not enough for the 1e-3 bar on logits of magnitude ~10).
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct Params {
    const void* A; int lda;           // [M, K] row-major, 16-bit
    const void* B; int ldb;           // MODE 0: [N, K] row-major (K contiguous); MODE 1: [K, N] row-major (N contiguous)
    int M, N, K;
    const void* bias;                 // MODE 0: b1 [N]
    void* out; int ldo;               // MODE 0: pre [M, N]; MODE 1: g_pooled [M, N]
    const void* w2; int k_head;       // MODE 0: W2 [k_head, N]
    float* part;                      // MODE 0: [N/BN, M, k_head]
};

// epilogue: thread = output row (TMEM lane), 2 chunks of 32 columns
template <typename T, int MODE>
__device__ __forceinline__ void epilogue(const Params& p, uint32_t tmem, const float* w2s, const float* b1s, int m0, int n0) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = m0 + tid;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    T* out = reinterpret_cast<T*>(p.out);
    if (MODE == 0) {
        float part[8];
        for (int kb = 0; kb < p.k_head; kb += 8) {
#pragma unroll
            for (int q = 0; q < 8; q++) part[q] = 0.f;
            for (int ch = 0; ch < BN / 32; ch++) {
                float v[32];
                __align__(16) T prs[32];
                tmem_ld32(lane_base + ch * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float pre = round_to<T>(v[i] + b1s[ch * 32 + i]);
                    const float r6 = fminf(fmaxf(pre + 3.f, 0.f), 6.f);
                    v[i] = round_to<T>(pre * r6 * (1.f / 6.f));
                    prs[i] = from_f32<T>(pre);
                }
                if (kb == 0 && row < p.M) {
                    int4* dst = reinterpret_cast<int4*>(out + (size_t)row * p.ldo + n0 + ch * 32);
#pragma unroll
                    for (int q = 0; q < (int)(32 * sizeof(T) / 16); q++) dst[q] = reinterpret_cast<const int4*>(prs)[q];
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (kb + q < p.k_head) {
                        const float* wr = w2s + (kb + q) * BN + ch * 32;
                        float a = part[q];
#pragma unroll
                        for (int i = 0; i < 32; i++) a = fmaf(v[i], wr[i], a);
                        part[q] = a;
                    }
                }
            }
            if (row < p.M)
                for (int q = 0; q < 8 && kb + q < p.k_head; q++)
                    p.part[((size_t)blockIdx.x * p.M + row) * p.k_head + kb + q] = part[q];
        }
    } else {
        for (int ch = 0; ch < BN / 32; ch++) {
            float v[32];
            tmem_ld32(lane_base + ch * 32, v);
            if (row < p.M) {
                __align__(16) T o16[32];
#pragma unroll
                for (int i = 0; i < 32; i++) o16[i] = from_f32<T>(v[i]);
                int4* dst = reinterpret_cast<int4*>(out + (size_t)row * p.ldo + n0 + ch * 32);
#pragma unroll
                for (int q = 0; q < (int)(32 * sizeof(T) / 16); q++) dst[q] = reinterpret_cast<const int4*>(o16)[q];
            }
        }
    }
}

// dynamic smem: STAGES x (A stage | B stage) | W2 tile fp32 [k_head][BN] | b1 tile [BN] | mbarriers | tmem slot
template <typename T, int MODE>
__global__ void __launch_bounds__(THREADS)
head_gemm_tc_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* w2s = reinterpret_cast<float*>(smem + STAGES * (A_STAGE + B_STAGE));
    float* b1s = w2s + (MODE == 0 ? p.k_head * BN : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(b1s + BN);          // [0, STAGES): slot free, [STAGES]: accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + STAGES + 1);
    const uint32_t ring = s32(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const T* A = reinterpret_cast<const T*>(p.A);
    const T* B = reinterpret_cast<const T*>(p.B);

    if (tid == 0) {
        for (int s = 0; s <= STAGES; s++) bar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // one K step of operands -> ring slot: 12 x 16-byte chunks per thread
    auto load_stage = [&](int ks) {
        const uint32_t a_dst = ring + (ks % STAGES) * (A_STAGE + B_STAGE), b_dst = a_dst + A_STAGE;
        const int k0 = ks * BK;
#pragma unroll
        for (int j = 0; j < (BM * BK / 8) / THREADS; j++) {
            const int c = tid + THREADS * j, r = c >> 3, kc = c & 7;
            const bool in = m0 + r < p.M;
            cp_async16(a_dst + kc * A_LBO + r * 16, A + (size_t)(in ? m0 + r : 0) * p.lda + k0 + kc * 8, in ? 16u : 0u);
        }
#pragma unroll
        for (int j = 0; j < (BN * BK / 8) / THREADS; j++) {
            const int c = tid + THREADS * j, r = c >> 3, q = c & 7;
            if (MODE == 0)    // B [N,K]: row n, 8 consecutive k -> K-major core-matrix row n
                cp_async16(b_dst + q * B_LBO + r * 16, B + (size_t)(n0 + r) * p.ldb + k0 + q * 8, 16u);
            else              // B [K,N]: row k, 8 consecutive n -> MN-major core matrix (k group r>>3, n group q), row k&7
                cp_async16(b_dst + (r >> 3) * B_LBO + q * SBO + (r & 7) * 16, B + (size_t)(k0 + r) * p.ldb + n0 + q * 8, 16u);
        }
    };
    const int ksteps = p.K / BK;
    for (int s = 0; s < DIST; s++) {
        if (s < ksteps) load_stage(s);
        cp_async_commit();
    }
    if (MODE == 0) {
        const T* w2 = reinterpret_cast<const T*>(p.w2);
        const T* b1 = reinterpret_cast<const T*>(p.bias);
        for (int e = tid; e < p.k_head * BN; e += THREADS) { int k = e / BN, n = e - k * BN; w2s[e] = to_f32(w2[(size_t)k * p.N + n0 + n]); }
        for (int e = tid; e < BN; e += THREADS) b1s[e] = to_f32(b1[n0 + e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = instr_desc(Fmt<T>::code, Fmt<T>::code, MODE == 1 ? 1 : 0);

    for (int ks = 0; ks < ksteps; ks++) {
        const int st = ks % STAGES;
        cp_async_wait<DIST - 1>();                                       // this thread's chunks of K step ks have landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // ... and are visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = ring + st * (A_STAGE + B_STAGE), b0 = a0 + A_STAGE;
#pragma unroll
            for (int kk = 0; kk < BK / 16; kk++) {
                const uint64_t ad = smem_desc(a0 + kk * 2 * A_LBO, A_LBO, SBO);
                const uint64_t bd = smem_desc(b0 + kk * 2 * B_LBO, B_LBO, SBO);
                mma_f16(tmem, ad, bd, idesc, (ks | kk) ? 1u : 0u);
            }
            mma_commit(&bars[st]);                                        // implies fence::before_thread_sync
            if (ks == ksteps - 1) mma_commit(&bars[STAGES]);
        }
        const int nxt = ks + DIST;
        if (nxt < ksteps) {
            // the slot of K step nxt was last read by the MMAs of K step nxt - STAGES
            if (nxt >= STAGES) bar_wait(&bars[nxt % STAGES], (uint32_t)((nxt / STAGES - 1) & 1));
            load_stage(nxt);
        }
        cp_async_commit();
    }
    bar_wait(&bars[STAGES], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    epilogue<T, MODE>(p, tmem, w2s, b1s, m0, n0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
}


// ---------------------------------------------------------------------------------------------------------------
// TMA version of the same tile: one elected thread streams the operands with cp.async.bulk.tensor (2-D tensor maps,
// 128-byte swizzle) into a STAGES-slot ring and issues the MMAs; the copies are written through the async proxy, so no
// proxy fence and no block barrier sits between a slot landing and the MMAs that read it (the cp.async version above
// serialises on exactly that).  A tile row is 64 elements = 128 bytes = one swizzle span:
//   A, forward W1 [N,K]:  K-major SWIZZLE_128B, 8-row groups 1024 bytes apart (SBO), K advances 32 bytes per MMA
//   backward W1 as [K,N]: MN-major SWIZZLE_128B, a K row is one 128-byte line of 64 n, 8-row K groups 1024 bytes apart
struct alignas(64) TmaMap { unsigned char bytes[128]; };     // CUtensorMap, opaque to device code

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    // start[0,14) | LBO[16,30) = 1 (unused for one swizzle span) | SBO[32,46) = 1024 >> 4 | version 1 [46,48) | SWIZZLE_128B = 2 [61,64)
    return (uint64_t)((addr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const TmaMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(bytes) : "memory");
}

// dynamic smem (1024-byte aligned inside the kernel): STAGES x (A | B) | W2 tile | b1 tile | full[STAGES] empty[STAGES] done | tmem slot
//
// fp32 (T = float): a row of a tile is 32 elements (still 128 bytes, so stage sizes, swizzle and descriptors are unchanged),
// one MMA covers K = 8, and B is K-major in BOTH modes (the backward gets W1 transposed by the split pre-pass).  The K loop
// runs `passes` = 3 times over the operands with the (hi, hi), (hi, lo), (lo, hi) tensor maps.
template <typename T, int MODE>
__global__ void __launch_bounds__(THREADS)
head_gemm_tma_kernel(const Params p, const __grid_constant__ TmaMap mapA, const __grid_constant__ TmaMap mapB,
                     const __grid_constant__ TmaMap mapAlo, const __grid_constant__ TmaMap mapBlo, int passes) {
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int BKE = 128 / (int)sizeof(T);                  // elements of K per stage
    constexpr bool B_KMAJOR = MODE == 0 || F32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
    float* w2s = reinterpret_cast<float*>(smem + STAGES * (A_STAGE + B_STAGE));
    float* b1s = w2s + (MODE == 0 ? p.k_head * BN : 0);
    uint64_t* full = reinterpret_cast<uint64_t*>(b1s + BN);
    uint64_t* empty = full + STAGES;
    uint64_t* done = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const uint32_t ring = s32(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kper = p.K / BKE, ksteps = kper * passes;
    auto fetch = [&](int ks, int slot) {                      // thread 0: operands of K step ks -> ring slot
        const int pass = ks / kper, k0 = (ks - pass * kper) * BKE;
        const TmaMap* ma = pass == 2 ? &mapAlo : &mapA;
        const TmaMap* mb = pass == 1 ? &mapBlo : &mapB;
        const uint32_t a_dst = ring + slot * (A_STAGE + B_STAGE);
        bar_expect_tx(&full[slot], A_STAGE + B_STAGE);
        tma_load_2d(a_dst, ma, k0, m0, &full[slot]);
        if (B_KMAJOR) tma_load_2d(a_dst + A_STAGE, mb, k0, n0, &full[slot]);
        else          tma_load_2d(a_dst + A_STAGE, mb, n0, k0, &full[slot]);
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
        bar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the whole ring is free: start streaming before anything else happens
        for (int s = 0; s < STAGES && s < ksteps; s++) fetch(s, s);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (MODE == 0) {
        const T* w2 = reinterpret_cast<const T*>(p.w2);
        const T* b1 = reinterpret_cast<const T*>(p.bias);
        for (int e = tid; e < p.k_head * BN; e += THREADS) { int k = e / BN, n = e - k * BN; w2s[e] = to_f32(w2[(size_t)k * p.N + n0 + n]); }
        for (int e = tid; e < BN; e += THREADS) b1s[e] = to_f32(b1[n0 + e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = instr_desc(Fmt<T>::code, Fmt<T>::code, B_KMAJOR ? 0 : 1);
        for (int ks = 0; ks < ksteps; ks++) {
            const int st = ks % STAGES;
            bar_wait(&full[st], (uint32_t)((ks / STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = ring + st * (A_STAGE + B_STAGE), b0 = a0 + A_STAGE;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {                  // 4 MMAs per 128-byte K slab: K = 16 (16-bit) or 8 (tf32) each
                const uint64_t ad = smem_desc_sw128(a0 + kk * 32);
                const uint64_t bd = smem_desc_sw128(B_KMAJOR ? b0 + kk * 32 : b0 + kk * 2048);
                if (F32) mma_tf32(tmem, ad, bd, idesc, (ks | kk) ? 1u : 0u);
                else mma_f16(tmem, ad, bd, idesc, (ks | kk) ? 1u : 0u);
            }
            mma_commit(&empty[st]);
            if (ks == ksteps - 1) mma_commit(done);
            // refill the slot the PREVIOUS K step used (its MMAs have had one step to finish)
            const int prev = ks - 1, nxt = prev + STAGES;
            if (prev >= 0 && nxt < ksteps) {
                const int ps = prev % STAGES;
                bar_wait(&empty[ps], (uint32_t)((prev / STAGES) & 1));
                fetch(nxt, ps);
            }
        }
    }
    __syncwarp();
    bar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue<T, MODE>(p, tmem, w2s, b1s, m0, n0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
}

// ---- TF32 split pre-pass (fp32 inputs): hi = cvt.rna.tf32(x), lo = x - hi; optionally transposed ([rows, cols] -> [cols, rows])
__device__ __forceinline__ float tf32_hi(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[i] = h; reinterpret_cast<float4*>(lo)[i] = l;
}

// x [rows, cols] row-major -> hi, lo [cols, rows]; 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
split_transpose_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int rows, int cols) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) t[j][tx] = (r0 + j < rows && c0 + tx < cols) ? x[(size_t)(r0 + j) * cols + c0 + tx] : 0.f;
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (c < cols && r < rows) {
            const float v = t[tx][j], h = tf32_hi(v);
            hi[(size_t)c * rows + r] = h; lo[(size_t)c * rows + r] = v - h;
        }
    }
}

inline size_t smem_bytes_tma(int mode, int k_head) {
    return 1024 + STAGES * (A_STAGE + B_STAGE) + (mode == 0 ? (size_t)k_head * BN * 4 : 0) + BN * 4 + (2 * STAGES + 1) * 8 + 16;
}

// logits[m,k] = b2[k] + sum_t part[t,m,k]   (fixed order: deterministic)
template <typename T>
__global__ void head_reduce_partials_kernel(const float* __restrict__ part, const T* __restrict__ b2, int tiles, int m, int k_head,
                                            float* __restrict__ logits) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * k_head) return;
    int k = e % k_head;
    float s = 0.f;
    for (int t = 0; t < tiles; t++) s += part[(size_t)t * m * k_head + e];
    logits[e] = s + to_f32(b2[k]);
}

inline size_t smem_bytes(int mode, int k_head) {
    return STAGES * (A_STAGE + B_STAGE) + (mode == 0 ? (size_t)k_head * BN * 4 : 0) + BN * 4 + (STAGES + 1) * 8 + 16;
}

}  // namespace tc
