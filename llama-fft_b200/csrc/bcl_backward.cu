// bcl_backward.cu -- gradients of the BC linear for calibration (reference: autograd through
// BlockCirculantLinear.forward, src/patch_llama_fft.py:226-275, driven by src/calibrate_bc.py:281-294).
//
//   y[n,j,r] = sum_i sum_s c[j,i,(r-s) mod B] x[n,i,s]
//   dx       = the SAME forward applied to gy with the transposed, index-reversed weights ct[i,j,t] = c[j,i,(-t) mod B]
//              (conj_transpose_kernel builds ct; the forward kernels of bcl_stream.cu do the rest)
//   dc[j,i,t] = sum_n sum_r gy[n,j,r] x[n,i,(r-t) mod B]     -> in the frequency domain  dC_f[j,i] = sum_n GY_f[n,j] conj(X_f[n,i])
//              act_spectrum_kernel: packed fp32 spectra of x and gy (register FFT core), grad_c_kernel: per (j,i) complex
//              multiply-accumulate over tokens (2x2 register tile per warp) + inverse real FFT.
// All arithmetic fp32, like the reference's autograd path (the forward up-casts to fp32 before the FFTs, :249-257).
#include "bcl_stream_parts.cuh"

namespace bcl {

template <int B>
__global__ void conj_transpose_kernel(const float* __restrict__ c, float* __restrict__ ct, int p, int q) {
    const int blk = blockIdx.x;                                       // output block (i, j)
    const int i = blk / p, j = blk - i * p;
    const float* src = c + ((size_t)j * q + i) * B;
    float* dst = ct + (size_t)blk * B;
    for (int t = threadIdx.x; t < B; t += blockDim.x) dst[t] = src[(B - t) & (B - 1)];
}

// (N, blocks*B) activations of any dtype -> fp32 packed half-spectra [N][blocks][B]
template <int B, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) act_spectrum_kernel(const __grid_constant__ ProblemDev P, float* __restrict__ out) {
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char bw_smem[];
    float2* tw = reinterpret_cast<float2*>(bw_smem);
    float2* fscr_all = reinterpret_cast<float2*>(bw_smem + (size_t)(B / 2) * 8);
    float* xs = reinterpret_cast<float*>(bw_smem + (size_t)(B / 2) * 8 + (size_t)WARPS * C::WARP_SCRATCH_BYTES);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    const int64_t n0 = (int64_t)blockIdx.x * NT;
    load_x_spectrum<B, NT, WARPS>(P, n0, xs, tw, fscr_all + (size_t)warp * (C::WARP_SCRATCH_BYTES / 8), lane, warp);
    __syncthreads();
    const int row_elems = P.q * B;
    for (int n = 0; n < NT; ++n) {
        if (n0 + n >= P.N) break;
        float4* dst = reinterpret_cast<float4*>(out + (size_t)(n0 + n) * row_elems);
        const float4* src = reinterpret_cast<const float4*>(xs + (size_t)n * row_elems);
        for (int e = threadIdx.x; e < row_elems / 4; e += WARPS * 32) dst[e] = src[e];
    }
}

// One warp owns a 2 (j) x 2 (i) tile of blocks.  Lane l holds packed reals LaneMap<B>::offset(l, ch) .. of every block.
template <int B, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) grad_c_kernel(const float* __restrict__ Gs, const float* __restrict__ Xs, int64_t N, int p, int q,
                                                             float* __restrict__ dc, int accumulate) {
    using LM = LaneMap<B>;
    using C = FftCfg<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    __shared__ float2 tw[B / 2];
    __shared__ __align__(16) float spec_s[WARPS][4][B];
    __shared__ __align__(16) float2 fscr[WARPS][C::G * C::GROUP_F2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    const int qt = (q + 1) / 2, pt = (p + 1) / 2;
    const int tile = blockIdx.x * WARPS + warp;
    if (tile >= qt * pt) return;
    const int j0 = (tile / qt) * 2, i0 = (tile % qt) * 2;
    const int j1 = min(j0 + 1, p - 1), i1 = min(i0 + 1, q - 1);       // edge tiles shadow the last row / column (not stored)
    float acc[4][CH][CW];
    float dcr[4], nyr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        dcr[a] = 0.f; nyr[a] = 0.f;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int e = 0; e < CW; ++e) acc[a][ch][e] = 0.f;
    }
    const size_t grow = (size_t)p * B, xrow = (size_t)q * B;
    for (int64_t n = 0; n < N; ++n) {
        float g[2][CH][CW], x[2][CH][CW];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            load_spec<float, CW>(Gs + n * grow + (size_t)j0 * B + LM::offset(lane, ch), g[0][ch]);
            load_spec<float, CW>(Gs + n * grow + (size_t)j1 * B + LM::offset(lane, ch), g[1][ch]);
            load_spec<float, CW>(Xs + n * xrow + (size_t)i0 * B + LM::offset(lane, ch), x[0][ch]);
            load_spec<float, CW>(Xs + n * xrow + (size_t)i1 * B + LM::offset(lane, ch), x[1][ch]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int jj = a >> 1, ii = a & 1;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch)
#pragma unroll
                for (int e = 0; e < CW; e += 2) {                     // G * conj(X)
                    const float gr = g[jj][ch][e], gi = g[jj][ch][e + 1], xr = x[ii][ch][e], xi = x[ii][ch][e + 1];
                    acc[a][ch][e] = fmaf(gr, xr, acc[a][ch][e]);
                    acc[a][ch][e] = fmaf(gi, xi, acc[a][ch][e]);
                    acc[a][ch][e + 1] = fmaf(gi, xr, acc[a][ch][e + 1]);
                    acc[a][ch][e + 1] = fmaf(-gr, xi, acc[a][ch][e + 1]);
                    if (ch == 0 && e == 0) {                         // lane 0: packed slot (DC, Nyquist), both real
                        dcr[a] = fmaf(gr, xr, dcr[a]);
                        nyr[a] = fmaf(gi, xi, nyr[a]);
                    }
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (lane == 0) { acc[a][0][0] = dcr[a]; acc[a][0][1] = nyr[a]; }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            float* d = &spec_s[warp][a][LM::offset(lane, ch)];
            if constexpr (CW == 4) *reinterpret_cast<float4*>(d) = make_float4(acc[a][ch][0], acc[a][ch][1], acc[a][ch][2], acc[a][ch][3]);
            else *reinterpret_cast<float2*>(d) = make_float2(acc[a][ch][0], acc[a][ch][1]);
        }
    }
    __syncwarp();
    const int gsel = lane / C::T, t = lane % C::T;
    const int a = min(gsel, 3);
    float2 out[C::V];
    group_irfft_packed<B>(spec_s[warp][a], fscr[warp] + gsel * C::GROUP_F2, tw, t, out);
    const int jj = a >> 1, ii = a & 1;
    const int j = j0 + jj, i = i0 + ii;
    if (gsel < 4 && j < p && i < q) {
        float* dst = dc + ((size_t)j * q + i) * B;
        const float inv_b = 1.0f / (float)B;
#pragma unroll
        for (int k1 = 0; k1 < C::OPT; ++k1)
#pragma unroll
            for (int k2 = 0; k2 < C::T; ++k2) {
                const int m = (t * C::OPT + k1) + C::V * k2;
                float2 v = out[k1 * C::T + k2];
                v.x *= inv_b; v.y *= inv_b;
                float2* d2 = reinterpret_cast<float2*>(dst + 2 * m);
                if (accumulate) { const float2 o = *d2; v.x += o.x; v.y += o.y; }
                *d2 = v;
            }
    }
}

template <typename K> static cudaError_t bw_set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int B>
static cudaError_t act_spectrum_t(const void* x, int dt, int64_t stride, int64_t N, int blocks, float* out, cudaStream_t st) {
    constexpr int WARPS = 8, NT = 1;
    auto kern = act_spectrum_kernel<B, NT, WARPS>;
    const size_t smem = (size_t)(B / 2) * 8 + (size_t)WARPS * FftCfg<B>::WARP_SCRATCH_BYTES + (size_t)NT * blocks * B * 4;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = bw_set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    ProblemDev P{};
    P.x = x; P.x_dtype = dt; P.x_stride = stride; P.N = N; P.q = blocks; P.p = 1;
    kern<<<(unsigned)((N + NT - 1) / NT), WARPS * 32, smem, st>>>(P, out);
    return cudaGetLastError();
}

size_t backward_workspace_bytes(int64_t N, int p, int q, int B) { return (size_t)N * ((size_t)p + q) * B * 4 + 256; }

cudaError_t launch_conj_transpose(const float* c, int p, int q, int B, float* ct, cudaStream_t st) {
    const unsigned grid = (unsigned)((size_t)p * q);
    switch (B) {
        case 64: conj_transpose_kernel<64><<<grid, 64, 0, st>>>(c, ct, p, q); break;
        case 128: conj_transpose_kernel<128><<<grid, 128, 0, st>>>(c, ct, p, q); break;
        default: conj_transpose_kernel<256><<<grid, 128, 0, st>>>(c, ct, p, q); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_backward_weights(const void* x, int x_dtype, int64_t x_stride, const void* gy, int gy_dtype, int64_t gy_stride, int64_t N, int p, int q,
                                    int B, float* dc, int accumulate, void* workspace, cudaStream_t st) {
    float* Xs = reinterpret_cast<float*>(workspace);
    float* Gs = Xs + (size_t)N * q * B;
    cudaError_t e;
    constexpr int WARPS = 4;
    const int tiles = ((p + 1) / 2) * ((q + 1) / 2);
    const unsigned grid = (unsigned)((tiles + WARPS - 1) / WARPS);
#define BCL_BW(BB)                                                                                       \
    {                                                                                                    \
        if ((e = act_spectrum_t<BB>(x, x_dtype, x_stride, N, q, Xs, st)) != cudaSuccess) return e;        \
        if ((e = act_spectrum_t<BB>(gy, gy_dtype, gy_stride, N, p, Gs, st)) != cudaSuccess) return e;     \
        grad_c_kernel<BB, WARPS><<<grid, WARPS * 32, 0, st>>>(Gs, Xs, N, p, q, dc, accumulate);           \
        return cudaGetLastError();                                                                       \
    }
    switch (B) { case 64: BCL_BW(64) case 128: BCL_BW(128) default: BCL_BW(256) }
#undef BCL_BW
}

}  // namespace bcl
