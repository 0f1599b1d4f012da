// Attribute-classifier head: torchvision MobileNetV3 `classifier`
//   Linear(960,1280) -> Hardswish -> Dropout(eval: identity) -> Linear(1280,K_head)
// (mobilenetv3.py:190-216, instantiated at E1:929-935 / E3:940-941 / E4:931-932), followed by the
// per-attribute slice / softmax / argmax / scatter of get_face_gender* (E1:1355-1401,
// E3:1387-1457, E4:1378-1475).
//
// The 960->1280 projection is the only GEMM-shaped work on the path (2.46 MFLOP per face).  bf16 / fp16 inputs run it
// on the tensor cores (fg_head_tc.cuh: tcgen05.mma with the accumulator in TMEM, bias + Hardswish + the partial
// second projection fused into the epilogue; FG_HEAD_SIMT=1 forces the CUDA-core path for A/B runs); fp32 inputs keep
// exact fp32 arithmetic with the shared-memory tiled fp32-accumulate GEMM below.
#include "fg_common.cuh"
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <cuda.h>             // CUtensorMap types only; the encoder is looked up through the runtime, libcuda is not linked

namespace {

__device__ __forceinline__ float hardswish(float x) {
    float r = fminf(fmaxf(x + 3.f, 0.f), 6.f);
    return x * r * (1.f / 6.f);
}
__device__ __forceinline__ float hardswish_grad(float x) {
    if (x < -3.f) return 0.f;
    if (x <= 3.f) return x * (1.f / 3.f) + 0.5f;
    return 1.f;
}

// C[M,N] = A[M,K] * op(B) (+ bias[N]);  A row-major (K contiguous).
// B_IS_NK: B is [N,K] row-major (a torch Linear weight, y = x W^T); else B is [K,N] row-major.
// EPI 0: C = acc + bias, stored as TC.   EPI 1: C = (acc) stored as TC (no bias).
template <typename TA, typename TB, typename TC, bool B_IS_NK, bool HAS_BIAS>
__global__ void __launch_bounds__(256)
gemm_tile_kernel(const TA* __restrict__ A, const TB* __restrict__ B, const TB* __restrict__ bias,
                 TC* __restrict__ C, int M, int N, int K) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    int tid = threadIdx.x;
    int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 4x4 outputs each
    int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: 64 rows x 16 k
        for (int e = tid; e < BM * BK; e += 256) {
            int r = e / BK, kk = e % BK;
            int gm = m0 + r, gk = k0 + kk;
            As[kk][r] = (gm < M && gk < K) ? to_f32(A[(size_t)gm * K + gk]) : 0.f;
        }
        if (B_IS_NK) {
            for (int e = tid; e < BN * BK; e += 256) {
                int c = e / BK, kk = e % BK;
                int gn = n0 + c, gk = k0 + kk;
                Bs[kk][c] = (gn < N && gk < K) ? to_f32(B[(size_t)gn * K + gk]) : 0.f;
            }
        } else {
            for (int e = tid; e < BN * BK; e += 256) {
                int kk = e / BN, c = e % BN;
                int gn = n0 + c, gk = k0 + kk;
                Bs[kk][c] = (gn < N && gk < K) ? to_f32(B[(size_t)gk * N + gn]) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (HAS_BIAS) v += to_f32(bias[gn]);
            C[(size_t)gm * N + gn] = from_f32<TC>(v);
        }
    }
}

// logits[m, k] = sum_j hardswish(pre[m,j]) * W2[k,j] + b2[k]; one warp per (row, group of 8 outputs)
template <typename T>
__global__ void __launch_bounds__(256)
head_logits_kernel(const T* __restrict__ pre, const T* __restrict__ w2, const T* __restrict__ b2,
                   float* __restrict__ logits, int m, int d_hid, int k_head) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int groups = (k_head + 7) / 8;
    int row = warp / groups, grp = warp - row * groups;
    if (row >= m) return;
    float acc[8] = {};
    const T* p = pre + (size_t)row * d_hid;
    for (int j = lane; j < d_hid; j += 32) {
        float h = round_to<T>(hardswish(to_f32(p[j])));
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int k = grp * 8 + q;
            if (k < k_head) acc[q] = fmaf(h, to_f32(w2[(size_t)k * d_hid + j]), acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        float v = acc[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        int k = grp * 8 + q;
        if (lane == 0 && k < k_head) logits[(size_t)row * k_head + k] = v + to_f32(b2[k]);
    }
}

// g_pre[m,j] = hardswish'(pre[m,j]) * sum_k g_logits[m,k] * W2[k,j]
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
head_bwd_hidden_kernel(const float* __restrict__ g_logits, const T* __restrict__ pre, const T* __restrict__ w2,
                       TO* __restrict__ g_pre, int m, int d_hid, int k_head) {
    extern __shared__ float gl[];            // [k_head] for this row
    int row = blockIdx.x;
    for (int k = threadIdx.x; k < k_head; k += blockDim.x) gl[k] = g_logits[(size_t)row * k_head + k];
    __syncthreads();
    for (int j = threadIdx.x; j < d_hid; j += blockDim.x) {
        // same summation order as before; unrolled so that the loads of 8 rows of W2 are in flight together (k_head = 80 for the
        // E1 head: 400 dependent L2 round trips per thread made this 61 us for 122 rows)
        const float pv = to_f32(pre[(size_t)row * d_hid + j]);
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < k_head; k++) s = fmaf(gl[k], to_f32(__ldg(w2 + (size_t)k * d_hid + j)), s);
        g_pre[(size_t)row * d_hid + j] = from_f32<TO>(s * hardswish_grad(pv));
    }
}

// per image: slice -> softmax -> argmax -> scatter (or fill)
template <typename T>
__global__ void head_attr_kernel(const float* __restrict__ logits, int m, int k_head,
                                 const int32_t* __restrict__ src_row, const uint8_t* __restrict__ selector,
                                 int n, int n_attr, int c0, int c1, int c2, int w0, int w1, int w2, float fill,
                                 long long* __restrict__ preds, T* __restrict__ probs, T* __restrict__ logits_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool sel = !selector || selector[i];
    int row = src_row ? src_row[i] : i;
    if (row < 0 || row >= m) sel = false;
    const int cs[3] = {c0, c1, c2}, ws[3] = {w0, w1, w2};
    size_t off = 0;
    for (int a = 0; a < n_attr; a++) {
        int w = ws[a];
        T* po = probs ? probs + off + (size_t)i * w : nullptr;
        T* lo = logits_out ? logits_out + off + (size_t)i * w : nullptr;
        if (!sel) {
            if (preds) preds[(size_t)a * n + i] = (long long)fill;
            for (int q = 0; q < w; q++) { if (po) po[q] = from_f32<T>(fill); if (lo) lo[q] = from_f32<T>(fill); }
        } else {
            const float* lg = logits + (size_t)row * k_head + cs[a];
            float mx = -INFINITY;
            for (int q = 0; q < w; q++) mx = fmaxf(mx, round_to<T>(lg[q]));
            float den = 0.f;
            for (int q = 0; q < w; q++) den += expf(round_to<T>(lg[q]) - mx);
            int best = 0; float bestp = -1.f;
            for (int q = 0; q < w; q++) {
                float p = round_to<T>(expf(round_to<T>(lg[q]) - mx) / den);
                if (p > bestp) { bestp = p; best = q; }          // first maximum, like torch.max
                if (po) po[q] = from_f32<T>(p);
                if (lo) lo[q] = from_f32<T>(lg[q]);
            }
            if (preds) preds[(size_t)a * n + i] = best;
        }
        off += (size_t)n * w;
    }
}


// backward of head_attr_kernel: one thread per image row; writes the WHOLE row of g_logits_full it owns
// (zeros outside the attribute slices), so no memset is needed.  g_full[row, c_a + q] = g_logits_a[i,q]
// + p_q * (g_probs_a[i,q] - sum_r g_probs_a[i,r] * p_r)   (softmax backward), rows that are not selected contribute nothing.
struct AttrBwdPtrs { const void* g_probs[3]; const void* g_logits[3]; };
template <typename T>
__global__ void head_attr_bwd_kernel(const T* __restrict__ probs, AttrBwdPtrs gp, const int32_t* __restrict__ src_row,
                                     const uint8_t* __restrict__ selector, int n, int m, int k_head, int n_attr,
                                     int c0, int c1, int c2, int w0, int w1, int w2, float* __restrict__ g_full) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (selector && !selector[i]) return;
    const int row = src_row ? src_row[i] : i;
    if (row < 0 || row >= m) return;
    float* out = g_full + (size_t)row * k_head;
    for (int q = 0; q < k_head; q++) out[q] = 0.f;
    const int cs[3] = {c0, c1, c2}, ws[3] = {w0, w1, w2};
    size_t off = 0;
    for (int a = 0; a < n_attr; a++) {
        const int w = ws[a];
        const T* p = probs + off + (size_t)i * w;
        const T* g_p = reinterpret_cast<const T*>(gp.g_probs[a]);
        const T* g_l = reinterpret_cast<const T*>(gp.g_logits[a]);
        float dot = 0.f;
        if (g_p) for (int q = 0; q < w; q++) dot += to_f32(g_p[(size_t)i * w + q]) * to_f32(p[q]);
        for (int q = 0; q < w; q++) {
            float acc = g_l ? to_f32(g_l[(size_t)i * w + q]) : 0.f;
            if (g_p) acc += to_f32(p[q]) * (to_f32(g_p[(size_t)i * w + q]) - dot);
            out[cs[a] + q] += acc;
        }
        off += (size_t)n * w;
    }
}

#include "fg_head_tc.cuh"

static bool env_flag(const char* name) {             // read once per process (tuning / A-B switches only)
    return getenv(name) != nullptr;
}
static bool head_simt() { static const bool v = env_flag("FG_HEAD_SIMT"); return v; }

// tensor-core path: K and N multiples of 64, 16-byte aligned rows; bf16 / fp16 operands as they are, fp32 operands as a
// three-pass TF32 split (needs the TMA kernel, i.e. the driver's tensor-map encoder)
static bool tc_ok(int dtype, int m, int d_in, int d_hid, int k_head, const void* a, const void* b) {
    return m > 0 && d_in % 64 == 0 && d_hid % 64 == 0 && k_head <= 256 &&
           ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && !head_simt();
}

// ---- 2-D tensor maps for the TMA kernel (row-major 16-bit matrix, 64-element x box_rows boxes, 128-byte swizzle)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    static const bool off = env_flag("FG_HEAD_CPASYNC");    // FG_HEAD_CPASYNC=1: A/B runs against the cp.async kernel
    return off ? nullptr : fn;
}

// Encoding a tensor map is pure host arithmetic on (base, shape, stride, box): a step encodes the same handful of maps every
// time (frozen weights, reused activation buffers), so the last few are kept per thread.  Nothing here is observable state:
// a hit returns byte-identical bytes to what the encoder would produce.
struct MapKey { const void* base; int inner, rows, ld, box_inner, box_rows, esize; };
template <typename T>
static bool make_map(tc::TmaMap* out, const void* base, int inner, int rows, int ld, int box_rows) {
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return false;
    constexpr int NCACHE = 16;
    thread_local MapKey keys[NCACHE];
    thread_local tc::TmaMap vals[NCACHE];
    thread_local int used = 0, next = 0;
    const int box_inner = 128 / (int)sizeof(T);                 // one 128-byte swizzle span of K (or of N for an MN-major B)
    const int esize = std::is_same<T, __half>::value ? -2 : (int)sizeof(T);      // fp16 and bf16 differ in the data type field
    const MapKey k = {base, inner, rows, ld, box_inner, box_rows, esize};
    for (int i = 0; i < used; i++)
        if (memcmp(&keys[i], &k, sizeof(k)) == 0) { memcpy(out, &vals[i], sizeof(*out)); return true; }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(T)};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
        : (std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (enc(&m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    static_assert(sizeof(CUtensorMap) == sizeof(tc::TmaMap), "tensor map size");
    memcpy(out, &m, sizeof(m));
    keys[next] = k; memcpy(&vals[next], &m, sizeof(m));
    next = (next + 1) % NCACHE; if (used < NCACHE) used++;
    return true;
}

// A [M,K] and B as the kernel reads them; for fp32 also the low halves of the TF32 split (A_lo, B_lo; same shapes)
template <typename T, int MODE>
static int launch_head_gemm(const tc::Params& p, const void* a_lo, const void* b_lo, dim3 grid, cudaStream_t st) {
    constexpr bool F32 = sizeof(T) == 4;
    tc::TmaMap mapA, mapB, mapAlo, mapBlo;
    // A [M,K]: boxes of 128 rows; B: K-major [N,K] boxes of 64 rows (n); the 16-bit backward reads W1 as [K,N]: 64 k-rows x 64 n
    auto map_b = [&](tc::TmaMap* o, const void* base) {
        return (MODE == 0 || F32) ? make_map<T>(o, base, p.K, p.N, p.ldb, tc::BN) : make_map<T>(o, base, p.N, p.K, p.ldb, tc::BK);
    };
    bool tma = make_map<T>(&mapA, p.A, p.K, p.M, p.lda, tc::BM) && map_b(&mapB, p.B);
    if (tma && F32) tma = make_map<T>(&mapAlo, a_lo, p.K, p.M, p.lda, tc::BM) && map_b(&mapBlo, b_lo);
    if (!F32) { mapAlo = mapA; mapBlo = mapB; }
    cudaError_t e;
    if (tma) {
        const size_t smem = tc::smem_bytes_tma(MODE, p.k_head);
        e = cudaFuncSetAttribute(tc::head_gemm_tma_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        tc::head_gemm_tma_kernel<T, MODE><<<grid, tc::THREADS, smem, st>>>(p, mapA, mapB, mapAlo, mapBlo, F32 ? 3 : 1);
    } else {
        if constexpr (F32) {
            return FG_ERR_INVALID_ARG;          // callers check tma_encoder() first and keep fp32 on the CUDA-core kernels without it
        } else {
            const size_t smem = tc::smem_bytes(MODE, p.k_head);
            e = cudaFuncSetAttribute(tc::head_gemm_tc_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            tc::head_gemm_tc_kernel<T, MODE><<<grid, tc::THREADS, smem, st>>>(p);
        }
    }
    return FG_OK;
}

// workspace carve for the fp32 (TF32 split) path
struct F32Ws { float* part; float* a_hi; float* a_lo; float* b_hi; float* b_lo; size_t total; };
static F32Ws f32_carve(void* base, int m, int d_in, int d_hid, int k_head) {
    F32Ws w; size_t off = 0;
    const size_t rows = (size_t)(m > 0 ? m : 1);
    const size_t big = (size_t)(d_in > d_hid ? d_in : d_hid);
    auto take = [&](size_t bytes) { size_t o = off; off += fg_align_up(bytes, 256); return (float*)((char*)base + o); };
    w.part = take((size_t)((d_hid + 63) / 64) * rows * k_head * sizeof(float));
    w.a_hi = take(rows * big * sizeof(float)); w.a_lo = take(rows * big * sizeof(float));
    w.b_hi = take((size_t)d_in * d_hid * sizeof(float)); w.b_lo = take((size_t)d_in * d_hid * sizeof(float));
    w.total = off;
    return w;
}

template <typename T>
static int head_fwd_tc(const void* pooled, const void* w1, const void* b1, const void* w2, const void* b2, int m, int d_in,
                       int d_hid, int k_head, void* hidden_pre, float* logits, float* part, cudaStream_t st) {
    tc::Params p;
    p.A = pooled; p.lda = d_in; p.B = w1; p.ldb = d_in; p.M = m; p.N = d_hid; p.K = d_in;
    p.bias = b1; p.out = hidden_pre; p.ldo = d_hid; p.w2 = w2; p.k_head = k_head; p.part = part;
    int rc = launch_head_gemm<T, 0>(p, nullptr, nullptr, dim3(d_hid / tc::BN, (m + tc::BM - 1) / tc::BM), st);
    if (rc) return rc;
    const int tot = m * k_head;
    tc::head_reduce_partials_kernel<T><<<(tot + 255) / 256, 256, 0, st>>>(part, (const T*)b2, d_hid / tc::BN, m, k_head, logits);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

// fp32: split pre-pass (pooled and W1 -> hi / lo), three-pass TF32 GEMM with the fused epilogue, partial reduction
static int head_fwd_tf32(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2, int m, int d_in,
                         int d_hid, int k_head, float* hidden_pre, float* logits, void* workspace, cudaStream_t st) {
    F32Ws w = f32_carve(workspace, m, d_in, d_hid, k_head);
    const size_t na = (size_t)m * d_in / 4, nb = (size_t)d_hid * d_in / 4;
    tc::split_tf32_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(pooled, w.a_hi, w.a_lo, na);
    tc::split_tf32_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(w1, w.b_hi, w.b_lo, nb);
    tc::Params p;
    p.A = w.a_hi; p.lda = d_in; p.B = w.b_hi; p.ldb = d_in; p.M = m; p.N = d_hid; p.K = d_in;
    p.bias = b1; p.out = hidden_pre; p.ldo = d_hid; p.w2 = w2; p.k_head = k_head; p.part = w.part;
    int rc = launch_head_gemm<float, 0>(p, w.a_lo, w.b_lo, dim3(d_hid / tc::BN, (m + tc::BM - 1) / tc::BM), st);
    if (rc) return rc;
    const int tot = m * k_head;
    tc::head_reduce_partials_kernel<float><<<(tot + 255) / 256, 256, 0, st>>>(w.part, b2, d_hid / tc::BN, m, k_head, logits);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

template <typename T>
static int head_bwd_tc(const float* g_logits, const void* hidden_pre, const void* w1, const void* w2, int m, int d_in, int d_hid,
                       int k_head, void* g_pooled, void* g_pre16, cudaStream_t st) {
    head_bwd_hidden_kernel<T, T><<<m, 256, k_head * sizeof(float), st>>>(g_logits, (const T*)hidden_pre, (const T*)w2, (T*)g_pre16,
                                                                        m, d_hid, k_head);
    tc::Params p;
    p.A = g_pre16; p.lda = d_hid; p.B = w1; p.ldb = d_in; p.M = m; p.N = d_in; p.K = d_hid;
    p.bias = nullptr; p.out = g_pooled; p.ldo = d_in; p.w2 = nullptr; p.k_head = 0; p.part = nullptr;
    int rc = launch_head_gemm<T, 1>(p, nullptr, nullptr, dim3(d_in / tc::BN, (m + tc::BM - 1) / tc::BM), st);
    if (rc) return rc;
    FG_LAUNCH_CHECK();
    return FG_OK;
}

// fp32 backward: g_pre (fp32) -> hi / lo, W1 [d_hid, d_in] -> transposed hi / lo [d_in, d_hid] (K-major B), three-pass TF32 GEMM
static int head_bwd_tf32(const float* g_logits, const float* hidden_pre, const float* w1, const float* w2, int m, int d_in, int d_hid,
                         int k_head, float* g_pooled, void* workspace, cudaStream_t st) {
    F32Ws w = f32_carve(workspace, m, d_in, d_hid, k_head);
    head_bwd_hidden_kernel<float, float><<<m, 256, k_head * sizeof(float), st>>>(g_logits, hidden_pre, w2, w.a_hi, m, d_hid, k_head);
    const size_t na = (size_t)m * d_hid / 4;
    tc::split_tf32_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(w.a_hi, w.a_hi, w.a_lo, na);          // in place: hi overwrites g_pre
    tc::split_transpose_tf32_kernel<<<dim3((d_in + 31) / 32, (d_hid + 31) / 32), 256, 0, st>>>(w1, w.b_hi, w.b_lo, d_hid, d_in);
    tc::Params p;
    p.A = w.a_hi; p.lda = d_hid; p.B = w.b_hi; p.ldb = d_hid; p.M = m; p.N = d_in; p.K = d_hid;
    p.bias = nullptr; p.out = g_pooled; p.ldo = d_in; p.w2 = nullptr; p.k_head = 0; p.part = nullptr;
    int rc = launch_head_gemm<float, 1>(p, w.a_lo, w.b_lo, dim3(d_in / tc::BN, (m + tc::BM - 1) / tc::BM), st);
    if (rc) return rc;
    FG_LAUNCH_CHECK();
    return FG_OK;
}

}  // namespace

extern "C" size_t fg_head_workspace_bytes(int m, int d_in, int d_hid, int k_head, int dtype) {
    // backward: g_pre [m, d_hid] (fp32, or 16-bit on the tensor-core path); forward: partial logits [d_hid/64, m, k_head];
    // fp32 adds the hi / lo halves of the TF32 split of both operands
    size_t rows = (size_t)(m > 0 ? m : 1);
    size_t a = rows * d_hid * sizeof(float), b = (size_t)((d_hid + 63) / 64) * rows * k_head * sizeof(float);
    size_t base = fg_align_up(a > b ? a : b, 256);
    if (dtype == FG_F32 && d_in > 0 && d_hid > 0 && k_head > 0) {
        size_t f = f32_carve(nullptr, m, d_in, d_hid, k_head).total;
        if (f > base) base = f;
    }
    return base;
}

extern "C" int fg_head_fwd(const void* pooled, const void* w1, const void* b1, const void* w2, const void* b2,
                           int m, int d_in, int d_hid, int k_head, void* hidden_pre, float* logits,
                           void* workspace, size_t workspace_bytes, int dtype, void* stream) {
    if (m < 0 || d_in <= 0 || d_hid <= 0 || k_head <= 0) return FG_ERR_INVALID_ARG;
    if (!pooled || !w1 || !b1 || !w2 || !b2 || !hidden_pre || !logits) return FG_ERR_INVALID_ARG;
    if (m == 0) return FG_OK;
    if (tc_ok(dtype, m, d_in, d_hid, k_head, pooled, w1) && (dtype != FG_F32 || tma_encoder())) {
        if (!workspace || workspace_bytes < fg_head_workspace_bytes(m, d_in, d_hid, k_head, dtype)) return FG_ERR_WORKSPACE;
        if (dtype == FG_BF16) return head_fwd_tc<__nv_bfloat16>(pooled, w1, b1, w2, b2, m, d_in, d_hid, k_head, hidden_pre, logits, (float*)workspace, fg_stream(stream));
        if (dtype == FG_F16) return head_fwd_tc<__half>(pooled, w1, b1, w2, b2, m, d_in, d_hid, k_head, hidden_pre, logits, (float*)workspace, fg_stream(stream));
        if (dtype == FG_F32) return head_fwd_tf32((const float*)pooled, (const float*)w1, (const float*)b1, (const float*)w2, (const float*)b2,
                                                  m, d_in, d_hid, k_head, (float*)hidden_pre, logits, workspace, fg_stream(stream));
        return FG_ERR_DTYPE;
    }
    dim3 grid((d_hid + 63) / 64, (m + 63) / 64);
    int groups = (k_head + 7) / 8;
    long long warps = (long long)m * groups;
    FG_DISPATCH_DTYPE(dtype, T,
        gemm_tile_kernel<T, T, T, true, true><<<grid, 256, 0, fg_stream(stream)>>>(
            (const T*)pooled, (const T*)w1, (const T*)b1, (T*)hidden_pre, m, d_hid, d_in);
        head_logits_kernel<T><<<(unsigned)((warps * 32 + 255) / 256), 256, 0, fg_stream(stream)>>>(
            (const T*)hidden_pre, (const T*)w2, (const T*)b2, logits, m, d_hid, k_head));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_head_bwd(const float* g_logits, const void* hidden_pre, const void* w1, const void* w2,
                           int m, int d_in, int d_hid, int k_head, void* g_pooled,
                           void* workspace, size_t workspace_bytes, int dtype, void* stream) {
    if (m < 0 || d_in <= 0 || d_hid <= 0 || k_head <= 0) return FG_ERR_INVALID_ARG;
    if (!g_logits || !hidden_pre || !w1 || !w2 || !g_pooled) return FG_ERR_INVALID_ARG;
    if (m == 0) return FG_OK;
    if (!workspace || workspace_bytes < fg_head_workspace_bytes(m, d_in, d_hid, k_head, dtype)) return FG_ERR_WORKSPACE;
    if (tc_ok(dtype, m, d_in, d_hid, k_head, hidden_pre, w1) && (dtype != FG_F32 || tma_encoder())) {
        if (dtype == FG_BF16) return head_bwd_tc<__nv_bfloat16>(g_logits, hidden_pre, w1, w2, m, d_in, d_hid, k_head, g_pooled, workspace, fg_stream(stream));
        if (dtype == FG_F16) return head_bwd_tc<__half>(g_logits, hidden_pre, w1, w2, m, d_in, d_hid, k_head, g_pooled, workspace, fg_stream(stream));
        if (dtype == FG_F32) return head_bwd_tf32(g_logits, (const float*)hidden_pre, (const float*)w1, (const float*)w2, m, d_in, d_hid, k_head,
                                                  (float*)g_pooled, workspace, fg_stream(stream));
        return FG_ERR_DTYPE;
    }
    float* g_pre = (float*)workspace;
    dim3 grid((d_in + 63) / 64, (m + 63) / 64);
    FG_DISPATCH_DTYPE(dtype, T,
        head_bwd_hidden_kernel<T, float><<<m, 256, k_head * sizeof(float), fg_stream(stream)>>>(
            g_logits, (const T*)hidden_pre, (const T*)w2, g_pre, m, d_hid, k_head);
        gemm_tile_kernel<float, T, T, false, false><<<grid, 256, 0, fg_stream(stream)>>>(
            g_pre, (const T*)w1, nullptr, (T*)g_pooled, m, d_in, d_hid));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_head_attributes(const float* logits, int m, int k_head, const int32_t* src_row, const uint8_t* selector,
                                  int n, int n_attr, const int32_t* col_start, const int32_t* width, float fill,
                                  int64_t* preds, void* probs, void* logits_out, int dtype, void* stream) {
    if (n < 0 || m < 0 || n_attr < 1 || n_attr > 3 || !col_start || !width || (!logits && m > 0)) return FG_ERR_INVALID_ARG;
    int c[3] = {0, 0, 0}, w[3] = {0, 0, 0};
    for (int a = 0; a < n_attr; a++) {
        c[a] = col_start[a]; w[a] = width[a];
        if (w[a] <= 0 || w[a] > 64 || c[a] < 0 || c[a] + w[a] > k_head) return FG_ERR_INVALID_ARG;
    }
    if (n == 0) return FG_OK;
    FG_DISPATCH_DTYPE(dtype, T,
        head_attr_kernel<T><<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(
            logits, m, k_head, src_row, selector, n, n_attr, c[0], c[1], c[2], w[0], w[1], w[2], fill,
            (long long*)preds, (T*)probs, (T*)logits_out));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_head_attributes_bwd(const void* probs, const void* const* g_probs, const void* const* g_logits_attr,
                                      const int32_t* src_row, const uint8_t* selector, int n, int m, int k_head, int n_attr,
                                      const int32_t* col_start, const int32_t* width, float* g_logits_full, int dtype, void* stream) {
    if (n < 0 || m < 0 || k_head <= 0 || n_attr < 1 || n_attr > 3 || !col_start || !width || !g_probs || !g_logits_attr)
        return FG_ERR_INVALID_ARG;
    if (m == 0 || n == 0) return FG_OK;
    if (!probs || !g_logits_full) return FG_ERR_INVALID_ARG;
    int c[3] = {0, 0, 0}, w[3] = {0, 0, 0};
    AttrBwdPtrs gp;
    for (int a = 0; a < 3; a++) { gp.g_probs[a] = nullptr; gp.g_logits[a] = nullptr; }
    for (int a = 0; a < n_attr; a++) {
        c[a] = col_start[a]; w[a] = width[a];
        if (w[a] <= 0 || c[a] < 0 || c[a] + w[a] > k_head) return FG_ERR_INVALID_ARG;
        gp.g_probs[a] = g_probs[a]; gp.g_logits[a] = g_logits_attr[a];
    }
    if (!src_row && !selector && m != n) return FG_ERR_INVALID_ARG;
    if (src_row || selector) {
        // rows of g_logits_full that no selected image maps to stay untouched: clear them first
        cudaError_t e = cudaMemsetAsync(g_logits_full, 0, (size_t)m * k_head * sizeof(float), fg_stream(stream));
        if (e != cudaSuccess) return (int)e;
    }
    FG_DISPATCH_DTYPE(dtype, T,
        head_attr_bwd_kernel<T><<<(n + 127) / 128, 128, 0, fg_stream(stream)>>>(
            (const T*)probs, gp, src_row, selector, n, m, k_head, n_attr, c[0], c[1], c[2], w[0], w[1], w[2], g_logits_full));
    FG_LAUNCH_CHECK();
    return FG_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// f1: FaceFeatsModel.semantic_search (E1:96-117) for query BATCHES on the tensor cores.
// top-1 of Q [m,d] against db [D,d] is a GEMM with an arg-max epilogue.  The streaming kernel of fg_align.cu is HBM-bound
// for the reference's micro-batches of 3-4 rows and loses to cuBLAS from ~16 queries on (warp-shuffle bound).  Here:
//   1. ONE TF32 pass on tcgen05 (M = 128 queries on the TMEM lanes, N = 128 database rows per CTA, K in 128-byte slabs
//      through the same TMA / mbarrier ring as the head GEMM) gives approximate scores t = <q, row> +- eps, with
//      eps <= 2^-9 |q| max|row| (the tensor core truncates both operands to 10 mantissa bits; db_norm_bound = max|row|);
//   2. the epilogue -- thread = query, reading its 128 TMEM columns -- stores the approximate scores [D, m] (12-50 MB, a
//      fraction of the database) and publishes the tile's maximum per query; a second small kernel then keeps every row
//      with t >= (final best t of that query) - 2 eps: the exact arg-max, every row tied with it and whatever else the
//      TF32 resolution cannot separate -- a handful of rows per query.  (Selecting inside the epilogue against the best t
//      published SO FAR kept ~5 rows per query and tile, 100 k pairs: 782 tiles start at once and none has seen the others.)
//   3. one warp per kept (query, row) recomputes the score with the streaming kernel's own fp32 expression and merges it
//      with the same 64-bit atomicMax key: results are IDENTICAL to the exact search, database untouched, no copy of it.
// A list that overflows flags the exact streaming search to run instead (a no-op launch otherwise).
int fg_internal_search_exact_if(const float* queries, const uint8_t* selector, int m, const float* db, int D, int d,
                                unsigned long long* keys, const unsigned* only_if, cudaStream_t st);

namespace {
namespace tc {
constexpr int SN = 128;                                   // database rows per tile
constexpr int S_STAGE = (BM + SN) * 128;                  // one K slab of 32 floats for 128 query rows + 128 database rows
constexpr int S_STAGES = 3;                               // 96 KB: two CTAs per SM, one's epilogue under the other's loads
constexpr int S_TMEM_COLS = 128;

struct SearchParams {
    const uint8_t* selector; int m, D, K;
    const float* qnorm; float eps_scale;                  // eps = eps_scale * qnorm[q]
    unsigned* best_seen;                                  // [m] order-preserving key of the best approximate score
    float* scores;                                        // [D, m] approximate scores (query-minor: coalesced both ways)
};
__device__ __forceinline__ unsigned skey(float f) { const unsigned b = __float_as_uint(f); return (b >> 31) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float sunkey(unsigned k) { return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k); }

__global__ void __launch_bounds__(THREADS)
face_search_tc_kernel(const SearchParams p, const __grid_constant__ TmaMap mapQ, const __grid_constant__ TmaMap mapD) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S_STAGES * S_STAGE);
    uint64_t* empty = full + S_STAGES;
    uint64_t* done = empty + S_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const uint32_t ring = s32(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * SN;
    const int ksteps = p.K / 32;
    auto fetch = [&](int ks, int slot) {
        const uint32_t a_dst = ring + slot * S_STAGE;
        bar_expect_tx(&full[slot], S_STAGE);
        tma_load_2d(a_dst, &mapQ, ks * 32, m0, &full[slot]);
        tma_load_2d(a_dst + BM * 128, &mapD, ks * 32, n0, &full[slot]);
    };
    if (tid == 0) {
        for (int s = 0; s < S_STAGES; s++) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
        bar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < S_STAGES && s < ksteps; s++) fetch(s, s);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(S_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        // c_format F32 | a, b format TF32 | both K-major | N = 128 | M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(SN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int ks = 0; ks < ksteps; ks++) {
            const int st = ks % S_STAGES;
            bar_wait(&full[st], (uint32_t)((ks / S_STAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = ring + st * S_STAGE, b0 = a0 + BM * 128;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) mma_tf32(tmem, smem_desc_sw128(a0 + kk * 32), smem_desc_sw128(b0 + kk * 32), idesc, (ks | kk) ? 1u : 0u);
            mma_commit(&empty[st]);
            if (ks == ksteps - 1) mma_commit(done);
            const int prev = ks - 1, nxt = prev + S_STAGES;
            if (prev >= 0 && nxt < ksteps) {
                const int ps = prev % S_STAGES;
                bar_wait(&empty[ps], (uint32_t)((prev / S_STAGES) & 1));
                fetch(nxt, ps);
            }
        }
    }
    __syncwarp();
    bar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int q = m0 + tid;
        const bool sel = q < p.m && (!p.selector || p.selector[q] == 1);
        float best = -INFINITY;
        float* out = p.scores + (size_t)n0 * p.m + (sel ? q : 0);        // [D, m]: the 32 lanes of a store are 32 adjacent queries
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int ch = 0; ch < SN / 32; ch++) {
            float v[32];
            tmem_ld32(lane_base + ch * 32, v);                         // warp-collective: every thread takes part
            if (sel) {
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const bool in = n0 + ch * 32 + i < p.D;
                    if (in) out[(size_t)(ch * 32 + i) * p.m] = v[i];
                    best = in ? fmaxf(best, v[i]) : best;
                }
            }
        }
        if (sel && best > -INFINITY) atomicMax(&p.best_seen[q], skey(best));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(S_TMEM_COLS) : "memory");
}

// norms of the queries, cleared keys / counters
__global__ void search_prep_kernel(const float* __restrict__ queries, int m, int d, float* __restrict__ qnorm,
                                   unsigned long long* __restrict__ keys, unsigned* __restrict__ best_seen, unsigned* __restrict__ counters) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0u;
    if (row >= m) return;
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) { const float v = queries[(size_t)row * d + k]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) { qnorm[row] = sqrtf(ss); keys[row] = 0ull; best_seen[row] = 0u; }
}

// rows within 2 eps of the query's best approximate score -> the list of (query, row) pairs to re-score exactly
__global__ void __launch_bounds__(256)
search_select_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ selector, int m, int D, const float* __restrict__ qnorm,
                     float eps_scale, const unsigned* __restrict__ best_seen, unsigned* __restrict__ counters, uint2* __restrict__ cand,
                     unsigned cand_cap) {
    const size_t total = (size_t)m * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / m), q = (int)(e - (size_t)r * m);
        if (selector && selector[q] != 1) continue;
        const unsigned seen = best_seen[q];
        if (!seen) continue;
        if (scores[e] >= sunkey(seen) - 2.f * eps_scale * qnorm[q]) {
            const unsigned slot = atomicAdd(&counters[0], 1u);
            if (slot < cand_cap) cand[slot] = make_uint2((unsigned)q, (unsigned)r); else counters[1] = 1u;
        }
    }
}

// exact score of every kept pair: the streaming kernel's expression (fg_align.cu face_search_kernel, d = 512 path and the
// generic one), merged with its key (score key << 32 | ~row: the lowest row wins ties)
__global__ void __launch_bounds__(256)
search_rescore_kernel(const float* __restrict__ queries, const float* __restrict__ db, int d, const unsigned* __restrict__ counters,
                      const uint2* __restrict__ cand, unsigned cand_cap, unsigned long long* __restrict__ keys) {
    if (counters[1]) return;                                           // overflow: the exact search runs instead
    const unsigned n = min(counters[0], cand_cap);
    const int lane = threadIdx.x & 31;
    for (unsigned c = blockIdx.x * 8 + (threadIdx.x >> 5); c < n; c += gridDim.x * 8) {
        const uint2 pr = cand[c];
        const float* qv = queries + (size_t)pr.x * d;
        const float* row = db + (size_t)pr.y * d;
        float acc = 0.f;
        if (d == 512) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row) + lane + 32 * k);
                const float4 w = __ldg(reinterpret_cast<const float4*>(qv) + lane + 32 * k);
                acc = fmaf(v.x, w.x, acc); acc = fmaf(v.y, w.y, acc); acc = fmaf(v.z, w.z, acc); acc = fmaf(v.w, w.w, acc);
            }
        } else {
            for (int k = lane; k < d; k += 32) acc = fmaf(__ldg(row + k), qv[k], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicMax(&keys[pr.x], ((unsigned long long)skey(acc) << 32) | (unsigned)(~pr.y));
    }
}

__global__ void search_decode_kernel(const unsigned long long* __restrict__ keys, int m, long long* __restrict__ best_row, float* __restrict__ similarity) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const unsigned long long k = keys[q];
    best_row[q] = k ? (long long)(unsigned)(~(unsigned)(k & 0xFFFFFFFFull)) : -1;
    if (similarity) similarity[q] = k ? sunkey((unsigned)(k >> 32)) : -1.f;
}
}  // namespace tc
}  // namespace

struct SearchWs { unsigned long long* keys; unsigned* best_seen; float* qnorm; unsigned* counters; uint2* cand; unsigned cap; float* scores; size_t total; };
static SearchWs search_carve(void* base, int m, int D) {
    SearchWs w; size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
    w.keys = (unsigned long long*)take((size_t)m * 8);
    w.best_seen = (unsigned*)take((size_t)m * 4);
    w.qnorm = (float*)take((size_t)m * 4);
    w.counters = (unsigned*)take(16);
    w.cap = (unsigned)m * 256u + 4096u;
    w.cand = (uint2*)take((size_t)w.cap * 8);
    w.scores = (float*)take((size_t)m * D * 4);
    w.total = off;
    return w;
}

extern "C" size_t fg_face_search_tc_workspace_bytes(int m, int D) { return search_carve(nullptr, m > 0 ? m : 1, D > 0 ? D : 1).total; }

extern "C" int fg_face_search_top1_tc(const float* queries, const uint8_t* selector, int m, const float* db, int D, int d,
                                      float db_norm_bound, int64_t* best_row, float* similarity, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    if (m < 0 || D < 1 || d < 32 || (d % 32) || !(db_norm_bound > 0.f) || !(db_norm_bound < 1e30f)) return FG_ERR_INVALID_ARG;
    if (m == 0) return FG_OK;
    if (!queries || !db || !best_row || ((uintptr_t)queries % 16) || ((uintptr_t)db % 16)) return FG_ERR_INVALID_ARG;
    if (!workspace || workspace_bytes < fg_face_search_tc_workspace_bytes(m, D)) return FG_ERR_WORKSPACE;
    cudaStream_t st = fg_stream(stream);
    SearchWs w = search_carve(workspace, m, D);
    tc::TmaMap mq, md;
    if (!make_map<float>(&mq, queries, d, m, d, tc::BM) || !make_map<float>(&md, db, d, D, d, tc::SN)) return FG_ERR_LIMIT;
    tc::search_prep_kernel<<<(m + 7) / 8, 256, 0, st>>>(queries, m, d, w.qnorm, w.keys, w.best_seen, w.counters);
    FG_LAUNCH_CHECK();
    tc::SearchParams p;
    p.selector = selector; p.m = m; p.D = D; p.K = d; p.qnorm = w.qnorm;
    p.eps_scale = 2.2e-3f * db_norm_bound;            // 2^-9 = 1.95e-3 (two truncations to 10 mantissa bits) + the fp32 accumulation
    p.best_seen = w.best_seen; p.scores = w.scores;
    const size_t smem = (size_t)tc::S_STAGES * tc::S_STAGE + 1024 + 256;
    cudaError_t e = cudaFuncSetAttribute(tc::face_search_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc::face_search_tc_kernel<<<dim3((D + tc::SN - 1) / tc::SN, (m + tc::BM - 1) / tc::BM), tc::THREADS, smem, st>>>(p, mq, md);
    FG_LAUNCH_CHECK();
    {
        size_t bx = ((size_t)m * D + 256 * 8 - 1) / (256 * 8);
        if (bx > (size_t)8 * FG_NUM_SMS) bx = (size_t)8 * FG_NUM_SMS;
        tc::search_select_kernel<<<(unsigned)bx, 256, 0, st>>>(w.scores, selector, m, D, w.qnorm, p.eps_scale, w.best_seen, w.counters, w.cand, w.cap);
        FG_LAUNCH_CHECK();
    }
    tc::search_rescore_kernel<<<2 * FG_NUM_SMS, 256, 0, st>>>(queries, db, d, w.counters, w.cand, w.cap, w.keys);
    FG_LAUNCH_CHECK();
    int rc = fg_internal_search_exact_if(queries, selector, m, db, D, d, w.keys, w.counters + 1, st);
    if (rc) return rc;
    tc::search_decode_kernel<<<(m + 127) / 128, 128, 0, st>>>(w.keys, m, (long long*)best_row, similarity);
    FG_LAUNCH_CHECK();
    return FG_OK;
}
