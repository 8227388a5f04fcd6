// bcl_stream.cu -- the bandwidth-bound small-N ("decode") BC-linear kernel and the spectrum precompute.
//
// One fused launch replaces steps 2-9 of the reference forward (src/patch_llama_fft.py:246-271):
//   rfft(x) per input block  ->  Y_f[j] = sum_i C_f[j,i] * X_f[i]  ->  irfft  ->  round to x.dtype -> + bias.
// The weight spectrum C_f is streamed exactly once from HBM in the STREAM layout ([p][q][B] packed reals);
// x's spectrum lives in shared memory; accumulators live in registers; y is written once.
//
// Work decomposition (plain, CTA-cooperative version; the grouped decode path prefers bcl_stream_ws.cu): warp w owns
// input blocks i = w, w+WARPS, ... of every row (split-q), partial row spectra meet in shared memory once per group
// of rows, and each finishing warp inverse-transforms G rows at once.  Lane l owns packed reals
// LaneMap<B>::offset(l, chunk) .. +CW of every block, i.e. each warp-wide load is a run of whole 128-byte lines.
#include "bcl_stream_parts.cuh"

namespace bcl {

// ---------------------------------------------------------------------------------------------------------
// spectrum precompute:  c (p,q,B) fp32  ->  STREAM layout, fp32 or bf16.   G = 32/T blocks per warp pass.
// replaces Cf = rfft(c.to(float32))  (src/patch_llama_fft.py:256-257)
// ---------------------------------------------------------------------------------------------------------
template <int B, typename SpecT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spectrum_stream_kernel(const float* __restrict__ c, SpecT* __restrict__ spec, int64_t nblocks) {
    using LM = LaneMap<B>;
    using C = FftCfg<B>;
    __shared__ float2 tw[B / 2];
    __shared__ __align__(16) float buf[WARPS][C::G][B];
    __shared__ __align__(16) float2 fscr[WARPS][C::G * C::GROUP_F2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / C::T, t = lane % C::T;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * C::G; base < nblocks; base += (int64_t)gridDim.x * WARPS * C::G) {
        const int64_t blk = base + g;
        const bool active = blk < nblocks;
        float2 a[C::V];
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) a[n1] = active ? *reinterpret_cast<const float2*>(c + blk * B + 2 * (C::T * n1 + t)) : make_float2(0.f, 0.f);
        group_rfft_packed<B>(a, fscr[warp] + g * C::GROUP_F2, buf[warp][g], tw, t, active);
        for (int gg = 0; gg < C::G; ++gg) {                          // coalesced write-out, one block per iteration
            if (base + gg >= nblocks) break;
#pragma unroll
            for (int ch = 0; ch < LM::CH; ++ch) {
                float v[LM::CW];
#pragma unroll
                for (int e = 0; e < LM::CW; ++e) v[e] = buf[warp][gg][LM::offset(lane, ch) + e];
                store_act<LM::CW>(spec, (base + gg) * B + LM::offset(lane, ch), sizeof(SpecT) == 4 ? BCL_F32 : BCL_BF16, v);
            }
        }
        __syncwarp();
    }
}

// Shared-memory carve-up of the plain kernels: [ tw | fft scratch per warp | part: WARPS pairs x WARPS partials x B | xs ]
template <int B, int WARPS> struct PlainSmem {
    static constexpr size_t tw_bytes = (size_t)(B / 2) * 8;
    static constexpr size_t fscr_bytes = (size_t)WARPS * FftCfg<B>::WARP_SCRATCH_BYTES;
    static constexpr size_t part_bytes = (size_t)WARPS * WARPS * B * 4;
    static constexpr size_t fixed = tw_bytes + fscr_bytes + part_bytes;
};

template <int B, int NT, typename SpecT, int WARPS>
__device__ inline void run_segment(const ProblemDev& P, int j0, int j1, int64_t n0, unsigned char* smem_raw) {
    using C = FftCfg<B>;
    using SM = PlainSmem<B, WARPS>;
    const float2* tw = reinterpret_cast<const float2*>(smem_raw);
    float2* fscr_all = reinterpret_cast<float2*>(smem_raw + SM::tw_bytes);
    float* part = reinterpret_cast<float*>(smem_raw + SM::tw_bytes + SM::fscr_bytes);
    float* xs = reinterpret_cast<float*>(smem_raw + SM::fixed);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* fscr = fscr_all + (size_t)warp * (C::WARP_SCRATCH_BYTES / 8);
    constexpr int UNROLL = (B == 256) ? 2 : (B == 128 ? 4 : 8);
    constexpr int RG = WARPS / NT;                                  // rows per group: RG*NT = WARPS (row, token) pairs
    static_assert(RG >= 1 && RG * NT == WARPS, "NT must divide WARPS");
    const int q = P.q;
    const int slice = (q + WARPS - 1) / WARPS;                      // blocks per warp per row (some warps have one fewer)
    const int nbs = (slice + UNROLL - 1) / UNROLL;                  // batches per row
    const int rows = j1 - j0;
    const int T = rows * nbs;                                        // flattened (row, batch) stream, identical for every warp
    const SpecT* spec = reinterpret_cast<const SpecT*>(P.spec) + (size_t)j0 * q * B;
    Batch<B, SpecT, UNROLL> bA, bB;
    if (T > 0) load_batch<B, SpecT, UNROLL, WARPS>(bA, spec, 0, q, warp, lane);   // in flight during the x FFT below
    // Programmatic dependent launch: everything above (twiddles, the first spectrum batch: constants of the stream) overlaps the
    // previous kernel's tail; x may be that kernel's output, and y may still be read by it, so wait here.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    load_x_spectrum<B, NT, WARPS>(P, n0, xs, tw, fscr, lane, warp);
    __syncthreads();
    RowAcc<B, NT> acc;
    acc.clear();
    int r = 0, b = 0;                                                // current row (relative to j0) and batch
    auto step = [&](Batch<B, SpecT, UNROLL>& cur, Batch<B, SpecT, UNROLL>& nxt) {
        int rn = r, bn = b + 1;
        if (bn == nbs) { bn = 0; ++rn; }
        if (rn < rows) load_batch<B, SpecT, UNROLL, WARPS>(nxt, spec + (size_t)rn * q * B, bn, q, warp, lane);
        mac_batch<B, NT, SpecT, UNROLL, WARPS>(cur, acc, xs, b, q, warp, lane);
        if (bn == 0) {                                               // row r complete for this warp
            const int rg = r % RG;
            store_partial<B, NT, WARPS>(acc, part, rg, warp, lane);
            acc.clear();
            if (rg == RG - 1 || rn == rows) {                        // group complete: G pairs per finishing warp
                __syncthreads();
                const int npairs = (rg + 1) * NT;
                const int p0 = warp * C::G;
                if (p0 < npairs)
                    finish_pairs<B, NT, WARPS>(P, j0 + (r - rg), n0, part + (size_t)p0 * WARPS * B, p0, min(C::G, npairs - p0), tw, fscr, lane);
                __syncthreads();
            }
        }
        r = rn; b = bn;
    };
    for (int t = 0; t < T; t += 2) {
        step(bA, bB);
        if (t + 1 < T) step(bB, bA);
    }
    __syncthreads();
}

template <int B, int NT, typename SpecT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (NT <= 2 ? 2 : 1)) stream_single_kernel(const __grid_constant__ ProblemDev P, int rows_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the next kernel of the stream may start its prologue
    fill_twiddles<B>(reinterpret_cast<float2*>(smem_raw), threadIdx.x, WARPS * 32);
    __syncthreads();
    const int j0 = blockIdx.x * rows_per_cta;
    const int j1 = min(P.p, j0 + rows_per_cta);
    const int64_t n0 = (int64_t)blockIdx.y * NT;
    run_segment<B, NT, SpecT, WARPS>(P, j0, j1, n0, smem_raw);
}

template <int B, int NT, typename SpecT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (NT <= 2 ? 2 : 1)) stream_grouped_kernel(const ProblemDev* __restrict__ problems, const int* __restrict__ cta_seg_ptr,
                                                                   const SegWork* __restrict__ segs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fill_twiddles<B>(reinterpret_cast<float2*>(smem_raw), threadIdx.x, WARPS * 32);
    __syncthreads();
    const int s0 = cta_seg_ptr[blockIdx.x], s1 = cta_seg_ptr[blockIdx.x + 1];
    for (int s = s0; s < s1; ++s) {
        const SegWork w = segs[s];
        const ProblemDev P = problems[w.g];
        run_segment<B, NT, SpecT, WARPS>(P, w.j0, w.j1, 0, smem_raw);
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
    constexpr int WARPS = 4;
    int64_t grid = (nblocks + WARPS * FftCfg<B>::G - 1) / (WARPS * FftCfg<B>::G);
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
    const size_t fscr = B == 64 ? FftCfg<64>::WARP_SCRATCH_BYTES : (B == 128 ? FftCfg<128>::WARP_SCRATCH_BYTES : FftCfg<256>::WARP_SCRATCH_BYTES);
    return (size_t)(B / 2) * 8 + (size_t)warps * fscr + (size_t)warps * warps * B * 4 + (size_t)nt * q * B * 4;
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
    // Every warp works on every row (split-q), so any row count per CTA keeps a CTA busy; fewer CTAs along p means
    // fewer redundant x FFTs.  Aim for >= 2 CTAs per SM, then give each CTA as many rows as that allows.
    const int64_t token_groups = (P.N + NT - 1) / NT;
    const int64_t want_p_ctas = token_groups >= 296 ? 1 : (296 + token_groups - 1) / token_groups;
    int rows = (int)((P.p + want_p_ctas - 1) / want_p_ctas);
    if (rows < 1) rows = 1;
    dim3 grid((unsigned)((P.p + rows - 1) / rows), (unsigned)token_groups);
    // launched with the programmatic-stream-serialization attribute: in a chain of dependent per-module launches (a decoder layer:
    // q/k/v -> o -> gate/up -> down) the prologue of each kernel hides behind the tail of the previous one
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(WARPS * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, P, rows);
    if (e != cudaSuccess) return e;
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
