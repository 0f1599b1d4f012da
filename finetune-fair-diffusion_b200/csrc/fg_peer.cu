// The two exchanges of the guidance path as one-shot NVLink transfers between peer-mapped buffers.  NVSwitch gives every
// GPU a direct path to every peer and both messages are <= 0.5 MB per rank, so what an exchange costs is protocol and launch
// latency, not bandwidth; as kernels of ours the exchanges need no host call, the whole multi-rank step is ONE CUDA graph
// (NCCL collectives could not be captured with it), and the strong-scaling step at 8 GPUs went from 0.384 to 0.356 ms.  The
// weak-scaling step is bound by the solver and by waiting for the slowest rank, not by the transport (1.498 vs 1.496 ms).
//
//   exchange 1 = the all-gather of {face indicator, probabilities}  (customized_all_gather E1:222-235, call sites E3:1978-1986)
//        fg_peer_push     every rank stores its packed row block straight into slot `rank` of EVERY peer's gather region
//                         (16-byte stores over NVLink), then raises its flag on every peer
//        fg_peer_wait_copy waits for all flags and hands the gathered rows to the assignment kernels
//   exchange 2 = the all-reduce (SUM) of the int32 Monte-Carlo plan counts  (E3:1535)
//        fg_peer_push     (self_only) publishes the rank's counts in its own region, raises its flag on every peer
//        fg_peer_wait_sum waits for all flags and sums the `world` count slabs, pulling the peers' over NVLink in rank
//                         order (integers: every rank ends with identical sums, like the all-reduce)
//
// All buffers are the same symmetric allocation on every rank (layout owned by the caller, dist.PeerExchange), addressed
// through a DEVICE array of the `world` base pointers.  Two region sets alternate by the parity of a device-resident epoch
// counter that fg_peer_epoch_advance bumps once per step: a rank can only be one exchange ahead of the slowest peer
// (it needs that peer's flag to get further), so a buffer is never overwritten while a peer still reads it.  Flags carry
// the epoch (monotone, never reset).  Nothing here needs the host: every entry point is a plain launch on the caller's
// stream, so the whole multi-rank step is ONE CUDA graph.  A wait that sees no flag after ~1 s gives up, records
// FG_PEER_TIMEOUT in status[0] and lets the step finish (with garbage the caller must discard) instead of hanging the GPU.
#include "fg_common.cuh"

namespace {

constexpr int PEER_MAX_WORLD = 64;
constexpr int FG_PEER_TIMEOUT = 0x100;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void peer_epoch_kernel(unsigned* epoch) { if (threadIdx.x == 0 && blockIdx.x == 0) *epoch = *epoch + 1u; }

// grid (chunks, targets).  Block (b, t) copies chunk b of src into target t's region; the last block to finish (ticket
// counter) raises this rank's flag on EVERY peer.
__global__ void __launch_bounds__(256)
peer_push_kernel(const uint4* __restrict__ src, size_t n16, uint8_t* const* __restrict__ peer_base, size_t region_off,
                 size_t parity_stride, size_t slot_off, int self_only, size_t flags_off, int flag_index, int rank, int world,
                 const unsigned* __restrict__ epoch_dev, unsigned* __restrict__ done_counter) {
    const unsigned epoch = *epoch_dev;
    const int target = self_only ? rank : (int)blockIdx.y;
    uint4* dst = reinterpret_cast<uint4*>(peer_base[target] + region_off + (size_t)(epoch & 1u) * parity_stride + slot_off);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n16; e += (size_t)gridDim.x * blockDim.x) dst[e] = src[e];
    __threadfence_system();                                   // this thread's stores are visible system-wide ...
    __syncthreads();                                          // ... for every thread of the block, before the ticket
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned ticket = atomicAdd(done_counter, 1u);
        if (ticket == total - 1u) {
            __threadfence_system();
            *done_counter = 0u;                               // ready for the next push on this stream
            for (int p = 0; p < world; p++)
                st_release_sys(reinterpret_cast<unsigned*>(peer_base[p] + flags_off) + flag_index * PEER_MAX_WORLD + rank, epoch);
        }
    }
}

// all `world` flags of one exchange >= epoch?  Called by every thread of a block; true on success.
__device__ __forceinline__ bool peer_wait_flags(const unsigned* flags, int world, unsigned epoch, int* status) {
    __shared__ int ok_s;
    if (threadIdx.x == 0) ok_s = 1;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        bool seen = false;
        for (int spin = 0; spin < (1 << 22); spin++) {        // ~1 s with the back-off below
            if ((int)(ld_acquire_sys(flags + threadIdx.x) - epoch) >= 0) { seen = true; break; }
            __nanosleep(spin < 64 ? 20 : 200);
        }
        if (!seen) { ok_s = 0; if (status) atomicOr(status, FG_PEER_TIMEOUT); }
    }
    __syncthreads();
    return ok_s != 0;
}

__global__ void __launch_bounds__(256)
peer_wait_copy_kernel(uint8_t* const* __restrict__ peer_base, int rank, int world, size_t flags_off, int flag_index,
                      size_t region_off, size_t parity_stride, const unsigned* __restrict__ epoch_dev, uint4* __restrict__ dst,
                      size_t n16, int* __restrict__ status) {
    const unsigned epoch = *epoch_dev;
    const uint8_t* base = peer_base[rank];
    peer_wait_flags(reinterpret_cast<const unsigned*>(base + flags_off) + flag_index * PEER_MAX_WORLD, world, epoch, status);
    const uint4* src = reinterpret_cast<const uint4*>(base + region_off + (size_t)(epoch & 1u) * parity_stride);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n16; e += (size_t)gridDim.x * blockDim.x) dst[e] = src[e];
}

__global__ void __launch_bounds__(256)
peer_wait_sum_kernel(uint8_t* const* __restrict__ peer_base, int rank, int world, size_t flags_off, int flag_index,
                     size_t region_off, size_t parity_stride, const unsigned* __restrict__ epoch_dev, int4* __restrict__ out,
                     size_t n16, int* __restrict__ status) {
    const unsigned epoch = *epoch_dev;
    peer_wait_flags(reinterpret_cast<const unsigned*>(peer_base[rank] + flags_off) + flag_index * PEER_MAX_WORLD, world, epoch, status);
    const size_t off = region_off + (size_t)(epoch & 1u) * parity_stride;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n16; e += (size_t)gridDim.x * blockDim.x) {
        int4 acc = make_int4(0, 0, 0, 0);
        for (int p = 0; p < world; p++) {                     // rank order; the loads of all peers are independent
            const int4 v = reinterpret_cast<const int4*>(peer_base[p] + off)[e];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[e] = acc;
    }
}

// ---- exchange 1 without the intermediate copies: the packed row block {indicator, probs_0, probs_1, probs_2} is assembled in
// shared memory straight from the head's outputs and stored into every peer (push_rows), and the gathered block is taken
// apart into the tensors the assignment kernels read (wait_unpack).  Replaces torch.cat + fg_peer_push and fg_peer_wait_copy +
// four slicing kernels.
struct RowSrc { const void* probs[3]; int w[3]; int n_attr; };
struct RowDst { void* probs[3]; int w[3]; int n_attr; };
constexpr int PUSH_ROWS = 64;            // rows per CTA: 64 * row bytes is a multiple of 16 for every element size and width

template <typename T>
__global__ void __launch_bounds__(256)
peer_push_rows_kernel(const uint8_t* __restrict__ ind, RowSrc src, int n, uint8_t* const* __restrict__ peer_base, size_t region_off,
                      size_t parity_stride, size_t slot_off, size_t flags_off, int rank, int world,
                      const unsigned* __restrict__ epoch_dev, unsigned* __restrict__ done_counter) {
    extern __shared__ __align__(16) uint8_t rows_s[];
    const unsigned epoch = *epoch_dev;
    const int wtot = 1 + src.w[0] + src.w[1] + src.w[2];
    const int r0 = blockIdx.x * PUSH_ROWS, nr = min(PUSH_ROWS, n - r0);
    T* rs = reinterpret_cast<T*>(rows_s);
    for (int e = threadIdx.x; e < nr * wtot; e += blockDim.x) {
        const int r = e / wtot, c = e - r * wtot, i = r0 + r;
        T v;
        if (c == 0) v = from_f32<T>(ind[i] ? 1.f : 0.f);
        else {
            int cc = c - 1, a = 0;
            while (a < 2 && cc >= src.w[a]) { cc -= src.w[a]; a++; }
            v = reinterpret_cast<const T*>(src.probs[a])[(size_t)i * src.w[a] + cc];
        }
        rs[e] = v;
    }
    __syncthreads();
    const size_t bytes = (size_t)nr * wtot * sizeof(T);             // a multiple of 16 (host check: n * row bytes % 16 == 0, PUSH_ROWS % 8 == 0)
    const size_t n16 = bytes / 16, boff = (size_t)r0 * wtot * sizeof(T);
    for (int p = 0; p < world; p++) {
        uint4* dst = reinterpret_cast<uint4*>(peer_base[p] + region_off + (size_t)(epoch & 1u) * parity_stride + slot_off + boff);
        for (size_t e = threadIdx.x; e < n16; e += blockDim.x) dst[e] = reinterpret_cast<const uint4*>(rows_s)[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(done_counter, 1u);
        if (ticket == gridDim.x - 1u) {
            __threadfence_system();
            *done_counter = 0u;
            for (int p = 0; p < world; p++)
                st_release_sys(reinterpret_cast<unsigned*>(peer_base[p] + flags_off) + rank, epoch);      // flag_index 0 = rows
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
peer_wait_unpack_kernel(uint8_t* const* __restrict__ peer_base, int rank, int world, size_t flags_off, size_t region_off,
                        size_t parity_stride, const unsigned* __restrict__ epoch_dev, uint8_t* __restrict__ ind_all, RowDst dst,
                        int n_all, int* __restrict__ status) {
    const unsigned epoch = *epoch_dev;
    const uint8_t* base = peer_base[rank];
    peer_wait_flags(reinterpret_cast<const unsigned*>(base + flags_off), world, epoch, status);
    const int wtot = 1 + dst.w[0] + dst.w[1] + dst.w[2];
    const T* rows = reinterpret_cast<const T*>(base + region_off + (size_t)(epoch & 1u) * parity_stride);
    const size_t total = (size_t)n_all * wtot;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / wtot; const int c = (int)(e - i * wtot);
        const T v = rows[e];
        if (c == 0) ind_all[i] = to_f32(v) != 0.f ? 1 : 0;
        else {
            int cc = c - 1, a = 0;
            while (a < 2 && cc >= dst.w[a]) { cc -= dst.w[a]; a++; }
            reinterpret_cast<T*>(dst.probs[a])[i * dst.w[a] + cc] = v;
        }
    }
}

}  // namespace

extern "C" int fg_peer_epoch_advance(uint32_t* epoch_dev, void* stream) {
    if (!epoch_dev) return FG_ERR_INVALID_ARG;
    peer_epoch_kernel<<<1, 32, 0, fg_stream(stream)>>>(epoch_dev);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_peer_push(const void* src, size_t bytes, const void* peer_base_dev, size_t region_off, size_t parity_stride,
                            size_t slot_off, int self_only, size_t flags_off, int flag_index, int rank, int world,
                            const uint32_t* epoch_dev, uint32_t* done_counter, void* stream) {
    if (!src || !peer_base_dev || !epoch_dev || !done_counter || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world ||
        flag_index < 0 || flag_index > 1)
        return FG_ERR_INVALID_ARG;
    if ((bytes % 16) || ((uintptr_t)src % 16) || (region_off % 16) || (parity_stride % 16) || (slot_off % 16)) return FG_ERR_INVALID_ARG;
    const size_t n16 = bytes / 16;
    unsigned chunks = (unsigned)((n16 + 255) / 256);
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    const dim3 grid(chunks, self_only ? 1 : world);
    peer_push_kernel<<<grid, 256, 0, fg_stream(stream)>>>((const uint4*)src, n16, (uint8_t* const*)peer_base_dev, region_off, parity_stride,
                                                         slot_off, self_only, flags_off, flag_index, rank, world, epoch_dev, done_counter);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_peer_wait_copy(const void* peer_base_dev, int rank, int world, size_t flags_off, int flag_index, size_t region_off,
                                 size_t parity_stride, const uint32_t* epoch_dev, void* dst, size_t bytes, int32_t* status, void* stream) {
    if (!peer_base_dev || !epoch_dev || !dst || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || flag_index < 0 || flag_index > 1)
        return FG_ERR_INVALID_ARG;
    if ((bytes % 16) || ((uintptr_t)dst % 16) || (region_off % 16) || (parity_stride % 16)) return FG_ERR_INVALID_ARG;
    const size_t n16 = bytes / 16;
    unsigned blocks = (unsigned)((n16 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    peer_wait_copy_kernel<<<blocks, 256, 0, fg_stream(stream)>>>((uint8_t* const*)peer_base_dev, rank, world, flags_off, flag_index, region_off,
                                                                parity_stride, epoch_dev, (uint4*)dst, n16, status);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_peer_wait_sum(const void* peer_base_dev, int rank, int world, size_t flags_off, int flag_index, size_t region_off,
                                size_t parity_stride, const uint32_t* epoch_dev, int32_t* out, size_t n_elems, int32_t* status, void* stream) {
    if (!peer_base_dev || !epoch_dev || !out || world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || flag_index < 0 || flag_index > 1)
        return FG_ERR_INVALID_ARG;
    if ((n_elems % 4) || ((uintptr_t)out % 16) || (region_off % 16) || (parity_stride % 16)) return FG_ERR_INVALID_ARG;
    const size_t n16 = n_elems / 4;
    unsigned blocks = (unsigned)((n16 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > 2 * FG_NUM_SMS) blocks = 2 * FG_NUM_SMS;
    peer_wait_sum_kernel<<<blocks, 256, 0, fg_stream(stream)>>>((uint8_t* const*)peer_base_dev, rank, world, flags_off, flag_index, region_off,
                                                               parity_stride, epoch_dev, (int4*)out, n16, status);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_peer_push_rows(const uint8_t* indicators, const void* const* probs, const int32_t* widths, int n_attr, int n,
                                 const void* peer_base_dev, size_t region_off, size_t parity_stride, size_t slot_off, size_t flags_off,
                                 int rank, int world, const uint32_t* epoch_dev, uint32_t* done_counter, int dtype, void* stream) {
    if (!indicators || !probs || !widths || n_attr < 1 || n_attr > 3 || n < 1 || !peer_base_dev || !epoch_dev || !done_counter ||
        world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
        return FG_ERR_INVALID_ARG;
    RowSrc src = {}; src.n_attr = n_attr;
    int wtot = 1;
    for (int a = 0; a < n_attr; a++) { if (!probs[a] || widths[a] < 1) return FG_ERR_INVALID_ARG; src.probs[a] = probs[a]; src.w[a] = widths[a]; wtot += widths[a]; }
    const size_t esz = dtype == FG_F32 ? 4 : 2;
    if (((size_t)n * wtot * esz) % 16 || (n % 8) || (region_off % 16) || (parity_stride % 16) || (slot_off % 16)) return FG_ERR_INVALID_ARG;
    const size_t smem = (size_t)PUSH_ROWS * wtot * esz;
    const unsigned grid = (unsigned)((n + PUSH_ROWS - 1) / PUSH_ROWS);
    FG_DISPATCH_DTYPE(dtype, T,
        peer_push_rows_kernel<T><<<grid, 256, smem, fg_stream(stream)>>>(indicators, src, n, (uint8_t* const*)peer_base_dev, region_off,
                                                                        parity_stride, slot_off, flags_off, rank, world, epoch_dev, done_counter));
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_peer_wait_unpack(const void* peer_base_dev, int rank, int world, size_t flags_off, size_t region_off, size_t parity_stride,
                                   const uint32_t* epoch_dev, uint8_t* indicators_all, void* const* probs_all, const int32_t* widths,
                                   int n_attr, int n_all, int32_t* status, int dtype, void* stream) {
    if (!peer_base_dev || !epoch_dev || !indicators_all || !probs_all || !widths || n_attr < 1 || n_attr > 3 || n_all < 1 ||
        world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || (region_off % 16) || (parity_stride % 16))
        return FG_ERR_INVALID_ARG;
    RowDst dst = {}; dst.n_attr = n_attr;
    int wtot = 1;
    for (int a = 0; a < n_attr; a++) { if (!probs_all[a] || widths[a] < 1) return FG_ERR_INVALID_ARG; dst.probs[a] = probs_all[a]; dst.w[a] = widths[a]; wtot += widths[a]; }
    size_t blocks = ((size_t)n_all * wtot + 255) / 256;
    if (blocks > 2 * FG_NUM_SMS) blocks = 2 * FG_NUM_SMS;
    FG_DISPATCH_DTYPE(dtype, T,
        peer_wait_unpack_kernel<T><<<(unsigned)blocks, 256, 0, fg_stream(stream)>>>((uint8_t* const*)peer_base_dev, rank, world, flags_off, region_off,
                                                                                    parity_stride, epoch_dev, indicators_all, dst, n_all, status));
    FG_LAUNCH_CHECK();
    return FG_OK;
}
