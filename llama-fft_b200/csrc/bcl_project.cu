// bcl_project.cu -- dense -> block-circulant projection on the GPU (layer construction).
//
// Replaces the Python p*q*B loops of BlockCirculantLinear.from_linear (src/patch_llama_fft.py:214-219) and
// dense_block_to_circulant_column_loss_aware (:110-128): one CTA per BxB block, thread t owns wrapped diagonal t
//   c[t]     = mean_r W[r, (r - t) mod B]
//   alpha    = sum_t c[t]*diag_sum[t] / max(B * sum_t c[t]^2, 1e-12)          (analytically 1)
//   out[t]   = alpha * c[t]
// The reference evaluates this with torch ops in the WEIGHT's dtype, so every intermediate is rounded to that
// dtype here as well (fp32 accumulation inside each reduction, as torch does for 16-bit sums/means).
// For a fixed r the B threads read B consecutive elements of one row of W (reversed and rotated): coalesced, and
// the whole weight is read exactly once.
#include "bcl_common.cuh"
#include "bcl_kernels.h"

namespace bcl {

__device__ inline float load_w(const void* W, int64_t off, int dt) {
    if (dt == BCL_F32) return reinterpret_cast<const float*>(W)[off];
    if (dt == BCL_F16) return __half2float(reinterpret_cast<const __half*>(W)[off]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(W)[off]);
}

template <int B>
__device__ inline float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < B / 32; ++w) s += red[w];
    return s;
}

template <int B>
__global__ void __launch_bounds__(B) project_kernel(const void* __restrict__ W, int dt, int64_t stride, int q, float* __restrict__ c) {
    __shared__ float red[B / 32];
    const int blk = blockIdx.x;
    const int j = blk / q, i = blk - j * q;
    const int t = threadIdx.x;
    const int64_t base = (int64_t)j * B * stride + (int64_t)i * B;
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < B; ++r) s += load_w(W, base + (int64_t)r * stride + ((r - t) & (B - 1)), dt);
    const float diag_sum = round_to(s, dt);
    const float cm = round_to(s / (float)B, dt);
    const float num = round_to(block_sum<B>(round_to(cm * diag_sum, dt), red), dt);
    const float cc = round_to(block_sum<B>(round_to(cm * cm, dt), red), dt);
    const float den = fmaxf(round_to((float)B * cc, dt), 1e-12f);
    const float alpha = round_to(num / den, dt);
    c[(int64_t)blk * B + t] = round_to(alpha * cm, dt);
}

cudaError_t launch_project_from_dense(const void* W, int w_dtype, int64_t w_stride, int p, int q, int B, float* c, cudaStream_t st) {
    const int64_t nb = (int64_t)p * q;
    if (nb == 0) return cudaSuccess;
    switch (B) {
        case 64: project_kernel<64><<<(unsigned)nb, 64, 0, st>>>(W, w_dtype, w_stride, q, c); break;
        case 128: project_kernel<128><<<(unsigned)nb, 128, 0, st>>>(W, w_dtype, w_stride, q, c); break;
        default: project_kernel<256><<<(unsigned)nb, 256, 0, st>>>(W, w_dtype, w_stride, q, c); break;
    }
    return cudaGetLastError();
}

}  // namespace bcl
