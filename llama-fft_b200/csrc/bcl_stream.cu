// bcl_stream.cu -- the bandwidth-bound small-N ("decode") BC-linear kernel and the spectrum precompute.
//
// One fused launch replaces steps 2-9 of the reference forward (src/patch_llama_fft.py:246-271):
//   rfft(x) per input block  ->  Y_f[j] = sum_i C_f[j,i] * X_f[i]  ->  irfft  ->  round to x.dtype -> + bias.
// The weight spectrum C_f is streamed exactly once from HBM in the STREAM layout ([p][q][B] packed reals);
// x's spectrum lives in shared memory; accumulators live in registers; y is written once.
//
// Work decomposition: one warp owns one output block-row j at a time and walks all q input blocks, so there is
// no cross-warp reduction and no block barrier inside the streaming loop.  Lane l owns packed reals
// LaneMap<B>::offset(l, chunk) .. +CW of every block, i.e. each warp-wide load is a run of whole 128-byte lines.
#include "bcl_common.cuh"
#include "bcl_kernels.h"

namespace bcl {

// ---------------------------------------------------------------------------------------------------------
// spectrum precompute:  c (p,q,B) fp32  ->  STREAM layout, fp32 or bf16.   One warp per (j,i) block.
// replaces Cf = rfft(c.to(float32))  (src/patch_llama_fft.py:256-257)
// ---------------------------------------------------------------------------------------------------------
template <int B, typename SpecT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spectrum_stream_kernel(const float* __restrict__ c, SpecT* __restrict__ spec, int64_t nblocks) {
    using LM = LaneMap<B>;
    __shared__ float2 tw[B / 2];
    __shared__ __align__(16) float buf[WARPS][B];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    for (int64_t blk = (int64_t)blockIdx.x * WARPS + warp; blk < nblocks; blk += (int64_t)gridDim.x * WARPS) {
        const float* src = c + blk * B;
#pragma unroll
        for (int ch = 0; ch < LM::CH; ++ch) {
            float v[LM::CW];
            load_act<LM::CW>(src, LM::offset(lane, ch), BCL_F32, v);
#pragma unroll
            for (int e = 0; e < LM::CW; ++e) buf[warp][LM::offset(lane, ch) + e] = v[e];
        }
        warp_rfft_packed<B>(buf[warp], tw, lane);
#pragma unroll
        for (int ch = 0; ch < LM::CH; ++ch) {
            float v[LM::CW];
#pragma unroll
            for (int e = 0; e < LM::CW; ++e) v[e] = buf[warp][LM::offset(lane, ch) + e];
            store_act<LM::CW>(spec, blk * B + LM::offset(lane, ch), sizeof(SpecT) == 4 ? BCL_F32 : BCL_BF16, v);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// fused forward
// ---------------------------------------------------------------------------------------------------------
struct SegWork { int g, j0, j1, pad; };

// rfft of NT tokens x q input blocks of problem P into xs[NT][q*B] (packed spectra), cooperatively by all warps.
template <int B, int NT, int WARPS>
__device__ inline void load_x_spectrum(const ProblemDev& P, int64_t n0, float* xs, const float2* tw, int lane, int warp) {
    using LM = LaneMap<B>;
    const int q = P.q;
    for (int idx = warp; idx < NT * q; idx += WARPS) {
        const int n = idx / q, i = idx - n * q;
        float* dst = xs + (size_t)idx * B;
        const bool valid = (n0 + n) < P.N;
#pragma unroll
        for (int ch = 0; ch < LM::CH; ++ch) {
            float v[LM::CW];
            if (valid) load_act<LM::CW>(P.x, (n0 + n) * P.x_stride + (int64_t)i * B + LM::offset(lane, ch), P.x_dtype, v);
            else {
#pragma unroll
                for (int e = 0; e < LM::CW; ++e) v[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < LM::CW; ++e) dst[LM::offset(lane, ch) + e] = v[e];
        }
        if (valid) warp_rfft_packed<B>(dst, tw, lane);
    }
}

// one output block-row: stream q spectrum blocks, accumulate, inverse FFT, epilogue, store.
template <int B, int NT, typename SpecT, int UNROLL>
__device__ inline void process_row(const ProblemDev& P, int j, int64_t n0, const float* xs, float* scratch, const float2* tw, int lane) {
    using LM = LaneMap<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    const int q = P.q;
    const SpecT* crow = reinterpret_cast<const SpecT*>(P.spec) + (size_t)j * q * B;

    float acc[NT][CH][CW];
    float accd[NT], accn[NT];          // DC / Nyquist products (real x real), only meaningful in lane 0
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        accd[n] = 0.f; accn[n] = 0.f;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int e = 0; e < CW; ++e) acc[n][ch][e] = 0.f;
    }

    for (int i0 = 0; i0 < q; i0 += UNROLL) {
        float cv[UNROLL][CH][CW];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int i = min(i0 + u, q - 1);                      // clamp: tail re-reads the last block, masked below
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) load_spec<SpecT, CW>(crow + (size_t)i * B + LM::offset(lane, ch), cv[u][ch]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (i0 + u < q) {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const float* xb = xs + ((size_t)n * q + (i0 + u)) * B;
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                        float xv[CW];
                        if constexpr (CW == 4) { float4 t = *reinterpret_cast<const float4*>(xb + LM::offset(lane, ch)); xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w; }
                        else { float2 t = *reinterpret_cast<const float2*>(xb + LM::offset(lane, ch)); xv[0] = t.x; xv[1] = t.y; }
#pragma unroll
                        for (int e = 0; e < CW; e += 2) {
                            const float cr = cv[u][ch][e], ci = cv[u][ch][e + 1], xr = xv[e], xi = xv[e + 1];
                            acc[n][ch][e] = fmaf(cr, xr, acc[n][ch][e]);
                            acc[n][ch][e] = fmaf(-ci, xi, acc[n][ch][e]);
                            acc[n][ch][e + 1] = fmaf(cr, xi, acc[n][ch][e + 1]);
                            acc[n][ch][e + 1] = fmaf(ci, xr, acc[n][ch][e + 1]);
                            if (ch == 0 && e == 0) {             // packed slot 0/1 of lane 0 is (DC, Nyquist): real x real
                                accd[n] = fmaf(cr, xr, accd[n]);
                                accn[n] = fmaf(ci, xi, accn[n]);
                            }
                        }
                    }
                }
            }
        }
    }

    const float inv_b = 1.0f / (float)B;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (n0 + n >= P.N) break;
        if (lane == 0) { acc[n][0][0] = accd[n]; acc[n][0][1] = accn[n]; }
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int e = 0; e < CW; ++e) scratch[LM::offset(lane, ch) + e] = acc[n][ch][e];
        warp_irfft_packed<B>(scratch, tw, lane);
        const float2* z = reinterpret_cast<const float2*>(scratch);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            const int t0 = LM::offset(lane, ch);                   // first time sample of this chunk
            float v[CW];
#pragma unroll
            for (int e = 0; e < CW; e += 2) {
                const float2 zz = z[brev<B>((t0 + e) >> 1)];
                v[e] = zz.x * inv_b; v[e + 1] = zz.y * inv_b;
            }
            const int64_t col = (int64_t)j * B + t0;
#pragma unroll
            for (int e = 0; e < CW; ++e) v[e] = round_to(v[e], P.x_dtype);      // y32.to(x.dtype) first (:266)
            if (P.bias != nullptr) {                                                // then + bias in the promoted dtype (:270-271)
                float bv[CW];
                load_act<CW>(P.bias, col, P.bias_dtype, bv);
#pragma unroll
                for (int e = 0; e < CW; ++e) v[e] += bv[e];
            }
            store_act<CW>(P.y, (n0 + n) * P.y_stride + col, P.y_dtype, v);
        }
        __syncwarp();
    }
}

template <int B, int NT, typename SpecT, int WARPS>
__device__ inline void run_segment(const ProblemDev& P, int j0, int j1, int64_t n0, float* xs, float* scratch_all, const float2* tw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_x_spectrum<B, NT, WARPS>(P, n0, xs, tw, lane, warp);
    __syncthreads();
    constexpr int UNROLL = (B == 256) ? 4 : (B == 128 ? 8 : 16);
    for (int j = j0 + warp; j < j1; j += WARPS)
        process_row<B, NT, SpecT, UNROLL>(P, j, n0, xs, scratch_all + warp * B, tw, lane);
    __syncthreads();
}

// dynamic shared memory: [ tw: B/2 float2 ][ scratch: WARPS*B float ][ xs: NT*qmax*B float ]
template <int B, int WARPS> __host__ __device__ constexpr size_t smem_fixed_bytes() { return (size_t)(B / 2) * 8 + (size_t)WARPS * B * 4; }

template <int B, int NT, typename SpecT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) stream_single_kernel(const __grid_constant__ ProblemDev P, int rows_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float* scratch = reinterpret_cast<float*>(smem_raw + (size_t)(B / 2) * 8);
    float* xs = scratch + WARPS * B;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    const int j0 = blockIdx.x * rows_per_cta;
    const int j1 = min(P.p, j0 + rows_per_cta);
    const int64_t n0 = (int64_t)blockIdx.y * NT;
    run_segment<B, NT, SpecT, WARPS>(P, j0, j1, n0, xs, scratch, tw);
}

template <int B, int NT, typename SpecT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) stream_grouped_kernel(const ProblemDev* __restrict__ problems, const int* __restrict__ cta_seg_ptr,
                                                                   const SegWork* __restrict__ segs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float* scratch = reinterpret_cast<float*>(smem_raw + (size_t)(B / 2) * 8);
    float* xs = scratch + WARPS * B;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    const int s0 = cta_seg_ptr[blockIdx.x], s1 = cta_seg_ptr[blockIdx.x + 1];
    for (int s = s0; s < s1; ++s) {
        const SegWork w = segs[s];
        const ProblemDev P = problems[w.g];
        run_segment<B, NT, SpecT, WARPS>(P, w.j0, w.j1, 0, xs, scratch, tw);
    }
}

// ---------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------
template <typename K> static cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int B, typename SpecT>
static cudaError_t launch_spectrum_t(const float* c, void* spec, int64_t nblocks, cudaStream_t st) {
    constexpr int WARPS = 8;
    int64_t grid = (nblocks + WARPS - 1) / WARPS;
    if (grid > 148 * 8) grid = 148 * 8;
    spectrum_stream_kernel<B, SpecT, WARPS><<<(unsigned)grid, WARPS * 32, 0, st>>>(c, reinterpret_cast<SpecT*>(spec), nblocks);
    return cudaGetLastError();
}

cudaError_t launch_spectrum_stream(const float* c, void* spec, int p, int q, int B, int spec_dtype, cudaStream_t st) {
    const int64_t nb = (int64_t)p * q;
    if (nb == 0) return cudaSuccess;
#define BCL_SP(BB) (spec_dtype == BCL_F32 ? launch_spectrum_t<BB, float>(c, spec, nb, st) : launch_spectrum_t<BB, __nv_bfloat16>(c, spec, nb, st))
    switch (B) { case 64: return BCL_SP(64); case 128: return BCL_SP(128); default: return BCL_SP(256); }
#undef BCL_SP
}

size_t stream_smem_bytes(int B, int nt, int q, int warps) {
    return (size_t)(B / 2) * 8 + (size_t)warps * B * 4 + (size_t)nt * q * B * 4;
}

// Tokens per pass: as many as fit a ~100 KB x-spectrum tile (so two CTAs still share an SM), at most 8.
int stream_pick_nt(int64_t N, int q, int B) {
    int nt = 1;
    while (nt < 8 && nt < N && (size_t)(2 * nt) * q * B * 4 <= 100 * 1024) nt *= 2;
    return nt;
}

template <int B, int NT, typename SpecT>
static cudaError_t launch_single_t(const ProblemDev& P, cudaStream_t st) {
    constexpr int WARPS = 8;
    auto kern = stream_single_kernel<B, NT, SpecT, WARPS>;
    const size_t smem = stream_smem_bytes(B, NT, P.q, WARPS);
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    // rows per CTA: one row per warp unless that would leave most SMs idle (tiny problems -> fewer rows per CTA)
    const int64_t token_groups = (P.N + NT - 1) / NT;
    int rows = WARPS;
    while (rows > 1 && ((P.p + rows - 1) / rows) * token_groups < 148) rows >>= 1;
    dim3 grid((unsigned)((P.p + rows - 1) / rows), (unsigned)token_groups);
    kern<<<grid, WARPS * 32, smem, st>>>(P, rows);
    return cudaGetLastError();
}

template <int B, typename SpecT>
static cudaError_t launch_single_b(const ProblemDev& P, cudaStream_t st) {
    switch (stream_pick_nt(P.N, P.q, B)) {
        case 1: return launch_single_t<B, 1, SpecT>(P, st);
        case 2: return launch_single_t<B, 2, SpecT>(P, st);
        case 4: return launch_single_t<B, 4, SpecT>(P, st);
        default: return launch_single_t<B, 8, SpecT>(P, st);
    }
}

cudaError_t launch_stream_single(const ProblemDev& P, int B, int spec_dtype, cudaStream_t st) {
#define BCL_ST(BB) (spec_dtype == BCL_F32 ? launch_single_b<BB, float>(P, st) : launch_single_b<BB, __nv_bfloat16>(P, st))
    switch (B) { case 64: return BCL_ST(64); case 128: return BCL_ST(128); default: return BCL_ST(256); }
#undef BCL_ST
}

template <int B, int NT, typename SpecT>
static cudaError_t launch_grouped_t(const GroupPlanHeader& H, const unsigned char* dev_plan, cudaStream_t st) {
    constexpr int WARPS = 8;
    auto kern = stream_grouped_kernel<B, NT, SpecT, WARPS>;
    const size_t smem = stream_smem_bytes(B, NT, H.qmax, WARPS);
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)H.num_ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const ProblemDev*>(dev_plan + H.off_problems),
                                                       reinterpret_cast<const int*>(dev_plan + H.off_cta_ptr),
                                                       reinterpret_cast<const SegWork*>(dev_plan + H.off_segs));
    return cudaGetLastError();
}

template <int B, typename SpecT>
static cudaError_t launch_grouped_b(const GroupPlanHeader& H, const unsigned char* dev_plan, cudaStream_t st) {
    switch (H.nt) {
        case 1: return launch_grouped_t<B, 1, SpecT>(H, dev_plan, st);
        case 2: return launch_grouped_t<B, 2, SpecT>(H, dev_plan, st);
        case 4: return launch_grouped_t<B, 4, SpecT>(H, dev_plan, st);
        default: return launch_grouped_t<B, 8, SpecT>(H, dev_plan, st);
    }
}

cudaError_t launch_stream_grouped(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st) {
    const unsigned char* d = reinterpret_cast<const unsigned char*>(dev_plan);
#define BCL_GR(BB) (H.spec_dtype == BCL_F32 ? launch_grouped_b<BB, float>(H, d, st) : launch_grouped_b<BB, __nv_bfloat16>(H, d, st))
    switch (H.B) { case 64: return BCL_GR(64); case 128: return BCL_GR(128); default: return BCL_GR(256); }
#undef BCL_GR
}

}  // namespace bcl
