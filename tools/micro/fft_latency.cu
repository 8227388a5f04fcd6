// Microbenchmark: cycles of one warp-pass of the register FFT core (bcl_fft.cuh), first (cold instruction cache) vs repeated execution,
// with 1 or 8 warps running it at once.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I llama-fft_b200/csrc tools/micro/fft_latency.cu
#include <cstdio>
#include "bcl_stream_parts.cuh"
using namespace bcl;

template <int B>
__global__ void __launch_bounds__(256) k(const float* x, float* y, long long* out, int iters) {
    using C = FftCfg<B>;
    __shared__ float2 tw[B / 2];
    __shared__ __align__(16) float buf[8][C::G][B];
    __shared__ __align__(16) float2 fscr[8][C::G * C::GROUP_F2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / C::T, t = lane % C::T;
    fill_twiddles<B>(tw, threadIdx.x, blockDim.x);
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float2 a[C::V];
        long long t0 = clock64();
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) a[n1] = __ldcg(reinterpret_cast<const float2*>(x + ((warp * C::G + g) * B + 2 * (C::T * n1 + t))));
        long long t1 = clock64();
        // consume the loads before stamping
        float s = 0.f;
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) s += a[n1].x;
        if (s == 123.456f) y[0] = s;
        long long t2 = clock64();
        group_rfft_packed<B>(a, fscr[warp] + g * C::GROUP_F2, buf[warp][g], tw, t, true);
        long long t3 = clock64();
        float2 o[C::V];
        group_irfft_packed<B>(buf[warp][g], fscr[warp] + g * C::GROUP_F2, tw, t, o);
        long long t4 = clock64();
#pragma unroll
        for (int v = 0; v < C::V; ++v) y[((warp * C::G + g) * C::V + v) * 2 * C::T + 2 * t] = o[v].x + o[v].y;
        long long t5 = clock64();
        if (lane == 0) { long long* r = out + (it * 8 + warp) * 6; r[0] = t1 - t0; r[1] = t2 - t1; r[2] = t3 - t2; r[3] = t4 - t3; r[4] = t5 - t4; r[5] = t0; }
    }
}

int main() {
    float *x, *y; long long* out;
    cudaMalloc(&x, 1 << 20); cudaMalloc(&y, 1 << 20); cudaMalloc(&out, 8 * 8 * 6 * 8);
    cudaMemset(x, 0, 1 << 20);
    long long h[8 * 8 * 6];
    for (int warps : {1, 8}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(out, 0, sizeof(h));
            k<128><<<1, 32 * warps>>>(x, y, out, 4);
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("B=128 warps=%d launch %d\n", warps, rep);
            for (int it = 0; it < 4; ++it) {
                for (int w = 0; w < warps; w += 7) printf("  it %d warp %d: issue-loads %lld, loads-land %lld, rfft %lld, irfft %lld, store %lld\n", it, w, h[(it * 8 + w) * 6], h[(it * 8 + w) * 6 + 1], h[(it * 8 + w) * 6 + 2], h[(it * 8 + w) * 6 + 3], h[(it * 8 + w) * 6 + 4]);
            }
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
