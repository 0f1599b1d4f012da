// Next row f4: the consumer side of the image gradient.
//
//   manual gradient all-reduce   E1:1996-2011   for p in params: isfinite(p.grad).all(); all_reduce(p.grad, SUM);
//                                               p.grad = p.grad / num_processes / N_backward
//
// The reference issues, per LoRA parameter tensor (~500 of them, a few KB each), one blocking reduction to a Python bool,
// one NCCL all-reduce and two element-wise divisions.  Here the gradients are packed into ONE contiguous fp32 bucket
// (+ one trailing slot carrying this rank's count of non-finite values), the bucket is all-reduced once, and one
// kernel writes the averaged values back: 3 launches + 1 collective instead of ~4 launches + 1 collective per tensor.
// Both kernels are table driven: a device array of tensor base pointers and the prefix sums of their sizes; a thread
// finds the tensor of its element by binary search over the prefix sums (log2(500) = 9 steps, the tables stay in L1).
#include "fg_common.cuh"

namespace {

__device__ __forceinline__ int find_tensor(const long long* __restrict__ offsets, int n_tensors, long long e) {
    int lo = 0, hi = n_tensors;                     // offsets[t] <= e < offsets[t+1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offsets[mid] <= e) lo = mid; else hi = mid; }
    return lo;
}

template <typename T>
__global__ void __launch_bounds__(256)
bucket_pack_kernel(const unsigned long long* __restrict__ ptrs, const long long* __restrict__ offsets, int n_tensors, long long total,
                   float* __restrict__ bucket, int* __restrict__ nonfinite) {
    int bad = 0;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int t = find_tensor(offsets, n_tensors, e);
        const float v = to_f32(reinterpret_cast<const T*>(ptrs[t])[e - offsets[t]]);
        bucket[e] = v;
        bad += !isfinite(v);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(nonfinite, bad);
}

// the trailing slot: this rank's count as a float, so the all-reduce also delivers the global count
__global__ void bucket_seal_kernel(const int* __restrict__ nonfinite, float* __restrict__ slot) { *slot = (float)*nonfinite; }

template <typename T>
__global__ void __launch_bounds__(256)
bucket_unpack_kernel(const unsigned long long* __restrict__ ptrs, const long long* __restrict__ offsets, int n_tensors, long long total,
                     const float* __restrict__ bucket, float inv_a, float inv_b) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int t = find_tensor(offsets, n_tensors, e);
        // p.grad / num_processes / N_backward: ATen multiplies by the reciprocal of a scalar divisor, one rounding per step
        const float v = round_to<T>(__fmul_rn(bucket[e], inv_a));
        reinterpret_cast<T*>(ptrs[t])[e - offsets[t]] = from_f32<T>(__fmul_rn(v, inv_b));
    }
}

}  // namespace

extern "C" int fg_grad_bucket_pack(const uint64_t* grad_ptrs, const int64_t* offsets, int n_tensors, int64_t total, float* bucket,
                                   int32_t* nonfinite_count, int dtype, void* stream) {
    if (n_tensors < 0 || total < 0) return FG_ERR_INVALID_ARG;
    if (!bucket || !nonfinite_count || (n_tensors > 0 && (!grad_ptrs || !offsets))) return FG_ERR_INVALID_ARG;
    cudaStream_t st = fg_stream(stream);
    cudaError_t e = cudaMemsetAsync(nonfinite_count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (total > 0) {
        long long blocks = (total + 255) / 256;
        if (blocks > 8 * FG_NUM_SMS) blocks = 8 * FG_NUM_SMS;
        FG_DISPATCH_DTYPE(dtype, T,
            bucket_pack_kernel<T><<<(int)blocks, 256, 0, st>>>((const unsigned long long*)grad_ptrs, (const long long*)offsets, n_tensors,
                                                               total, bucket, nonfinite_count));
        FG_LAUNCH_CHECK();
    }
    bucket_seal_kernel<<<1, 1, 0, st>>>(nonfinite_count, bucket + total);
    FG_LAUNCH_CHECK();
    return FG_OK;
}

extern "C" int fg_grad_bucket_unpack(const uint64_t* grad_ptrs, const int64_t* offsets, int n_tensors, int64_t total,
                                     const float* bucket, double divisor_a, double divisor_b, int dtype, void* stream) {
    if (n_tensors < 0 || total < 0 || divisor_a == 0.0 || divisor_b == 0.0) return FG_ERR_INVALID_ARG;
    if (total == 0) return FG_OK;
    if (!bucket || !grad_ptrs || !offsets) return FG_ERR_INVALID_ARG;
    long long blocks = (total + 255) / 256;
    if (blocks > 8 * FG_NUM_SMS) blocks = 8 * FG_NUM_SMS;
    const float inv_a = 1.0f / (float)divisor_a, inv_b = 1.0f / (float)divisor_b;
    FG_DISPATCH_DTYPE(dtype, T,
        bucket_unpack_kernel<T><<<(int)blocks, 256, 0, fg_stream(stream)>>>((const unsigned long long*)grad_ptrs, (const long long*)offsets,
                                                                            n_tensors, total, bucket, inv_a, inv_b));
    FG_LAUNCH_CHECK();
    return FG_OK;
}
