// bcl_prefill_fused.cu -- the whole prefill forward (rfft -> per-frequency GEMM -> irfft + epilogue) as ONE launch.
//
// reference: src/patch_llama_fft.py:246-271 (the forward), src/bench_all_in_one.py:179-231 (cached spectrum).
// Three separate stage kernels move every intermediate through HBM: x -> Xf -> Yf -> y is three full read+write passes
// (profiles/r1_prefill_notes.md: 151 us for 4096->4096 at 16384 tokens, each stage at ~80 % of the HBM copy peak).  Here the
// 148 persistent CTAs of one grid are split into three groups, one per stage, that work on the SAME token tiles a few
// tiles apart: the forward-DFT CTAs publish tile m of Xf (per-tile completion counter in global memory, release/acquire
// + async-proxy fences around the TMA stores / loads), the GEMM CTAs pick it up while it is still in L2 and publish Yf,
// the inverse-DFT CTAs finish the tile.  Only x is read from and y written to HBM; Xf and Yf live in the 126 MB L2.
// A lag limit (back-pressure counter) keeps a fast group from running ahead of its consumer and spilling the intermediates.
// The group sizes follow the per-stage bytes over the measured L2-resident per-SM rates (BCL_FUSED_SPLIT overrides).
#include <cstdlib>
#include <cstdio>
#include "bcl_prefill_bodies.cuh"

namespace bcl {
namespace gemm {

constexpr int FUSED_THREADS = DFT_THREADS + 32;          // + the signaller warp (SIG_WARP)
static_assert(FUSED_THREADS == 32 * (SIG_WARP + 1), "signaller = last warp");

template <int B>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
prefill_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWf, const __grid_constant__ CUtensorMap tmXfS,
                     const __grid_constant__ CUtensorMap tmXfL, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmYfS,
                     const __grid_constant__ CUtensorMap tmYfL, const __grid_constant__ CUtensorMap tmWi, const __grid_constant__ CUtensorMap tmOut,
                     const __grid_constant__ DftArgs a1, const __grid_constant__ GemmArgs a2, const __grid_constant__ DftArgs a3, const int n1, const int n2) {
    extern __shared__ __align__(1024) unsigned char fused_smem[];
    const int b = (int)blockIdx.x;
    if (b < n1) dft_fwd_body<B, false>(fused_smem, tmX, tmWf, tmXfS, nullptr, a1, b, n1);
    else if (b < n1 + n2) freq_gemm_body<false>(fused_smem, tmXfL, tmC, tmYfS, nullptr, a2, b - n1, n2);
    else dft_inv_body<B, false>(fused_smem, tmYfL, tmWi, tmOut, nullptr, a3, b - n1 - n2, (int)gridDim.x - n1 - n2);
}

static_assert(DFT_THREADS == GEMM_THREADS, "the three stage bodies share one block shape");

// ---------------------------------------------------------------------------------------------------------
// Chain kernel: the three stages one after the other in ONE launch, separated by grid-wide barriers (all CTAs co-resident).
// Same work split as the three stand-alone launches (CTA b runs items b, b + grid, ... of every stage), same results bit for
// bit; what it saves is two launch transitions per forward (drain, launch latency, TMEM / barrier set-up behind a cold start)
// and two thirds of the host launch cost -- which is what bounds prefills of a few hundred to a few thousand tokens.
// B = 256 runs both half-N passes of a DFT stage back to back (no barrier needed between them).
// ---------------------------------------------------------------------------------------------------------
struct ChainLo { DftLoMaps fwd; GemmLoMaps gemm; DftLoMaps inv; };      // lo halves of the split (fp32-spectrum) layout

template <int B, bool SPLIT>
__global__ void __launch_bounds__(DFT_THREADS, 1)
prefill_chain_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWf, const __grid_constant__ CUtensorMap tmXfS,
                     const __grid_constant__ CUtensorMap tmXfL, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmYfS,
                     const __grid_constant__ CUtensorMap tmYfL, const __grid_constant__ CUtensorMap tmWi, const __grid_constant__ CUtensorMap tmOut,
                     const __grid_constant__ ChainLo lo, const __grid_constant__ DftArgs a1, const __grid_constant__ GemmArgs a2,
                     const __grid_constant__ DftArgs a3, int* grid_ctr) {
    extern __shared__ __align__(1024) unsigned char chain_smem[];
    const int b = (int)blockIdx.x, n = (int)gridDim.x;
    constexpr int HALVES = B == 256 ? 2 : 1;
    for (int h = 0; h < HALVES; ++h) {
        DftArgs a = a1;
        a.half = h;
        dft_fwd_body<B, SPLIT>(chain_smem, tmX, tmWf, tmXfS, SPLIT ? &lo.fwd : nullptr, a, b, n);
        __syncthreads();
    }
    if (!grid_barrier(grid_ctr, n)) { if (threadIdx.x == 0) set_status(a1.status, 31); return; }
    freq_gemm_body<SPLIT>(chain_smem, tmXfL, tmC, tmYfS, SPLIT ? &lo.gemm : nullptr, a2, b, n);
    if (!grid_barrier(grid_ctr, 2 * n)) { if (threadIdx.x == 0) set_status(a1.status, 32); return; }
    for (int h = 0; h < HALVES; ++h) {
        DftArgs a = a3;
        a.half = h;
        dft_inv_body<B, SPLIT>(chain_smem, tmYfL, tmWi, tmOut, SPLIT ? &lo.inv : nullptr, a, b, n);
        __syncthreads();
    }
}

template <int B, bool SPLIT>
static cudaError_t launch_chain_t(const StageLaunch<DftArgs>& L1, const StageLaunch<GemmArgs>& L2, const StageLaunch<DftArgs>& L3, const ChainLo& lo,
                                  int* grid_ctr, int ctas, cudaStream_t st) {
    size_t smem = L1.smem > L2.smem ? L1.smem : L2.smem;
    if (L3.smem > smem) smem = L3.smem;
    auto kern = prefill_chain_kernel<B, SPLIT>;
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(kern, (unsigned)ctas, DFT_THREADS, smem, st, L1.tm_in, L1.tm_w, L1.tm_out, L2.tm_in, L2.tm_w, L2.tm_out, L3.tm_in, L3.tm_w, L3.tm_out, lo,
                      L1.a, L2.a, L3.a, grid_ctr);
}

// One token chunk [n_begin, n_begin + nc) through all three tensor-core stages in one launch.  grid_ctr: one int, zeroed in stream
// order before the launch.  Xf_lo / Yf_lo: the lo halves of the split layout (g.split), else NULL.
cudaError_t launch_prefill_chain(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, void* Xf_lo, void* Yf,
                                 void* Yf_lo, int* status, int* grid_ctr, int ctas, cudaStream_t st) {
    StageLaunch<DftArgs> L1, L3;
    StageLaunch<GemmArgs> L2;
    ChainLo lo{};
    const bool split = g.split != 0;
    if (split && g.B > 128) return cudaErrorNotSupported;
    if (!prep_dft_fwd(P, g, spec_gemm, n_begin, nc, Xf, status, L1, split ? Xf_lo : nullptr, split ? &lo.fwd : nullptr)) return cudaErrorNotSupported;
    if (!prep_gemm(Xf, spec_gemm, Yf, status, g, nc, L2)) return cudaErrorNotSupported;
    if (split && !prep_gemm_lo(Xf_lo, spec_gemm, Yf_lo, g, nc, &lo.gemm)) return cudaErrorNotSupported;
    if (!prep_dft_inv(P, g, spec_gemm, n_begin, nc, Yf, status, L3, split ? Yf_lo : nullptr, split ? &lo.inv : nullptr)) return cudaErrorNotSupported;
    switch (g.B) {
        case 64: return split ? launch_chain_t<64, true>(L1, L2, L3, lo, grid_ctr, ctas, st) : launch_chain_t<64, false>(L1, L2, L3, lo, grid_ctr, ctas, st);
        case 128: return split ? launch_chain_t<128, true>(L1, L2, L3, lo, grid_ctr, ctas, st) : launch_chain_t<128, false>(L1, L2, L3, lo, grid_ctr, ctas, st);
        default: return launch_chain_t<256, false>(L1, L2, L3, lo, grid_ctr, ctas, st);
    }
}

size_t fused_counter_bytes(int64_t N) {
    const size_t tiles = (size_t)((N + TILE_M - 1) / TILE_M);
    return (3 * tiles * sizeof(int) + 255) & ~(size_t)255;
}

static bool env_ints(const char* name, int* out, int n) {
    const char* e = getenv(name);
    if (!e) return false;
    for (int i = 0; i < n; ++i) {
        char* end = nullptr;
        out[i] = (int)strtol(e, &end, 10);
        if (end == e) return false;
        e = (*end == ',') ? end + 1 : end;
    }
    return true;
}

// CTAs per stage: proportional to the stage's L2-side bytes per token over its per-SM rate (GB/s per SM on L2-resident data,
// measured with each stage alone under CUDA-graph replay: gpurun_out/graph_scan.log -> profiles/r2 notes).
void fused_split(const Geometry& g, int ctas, int* n1, int* n2, int* n3) {
    int ov[2];
    if (env_ints("BCL_FUSED_SPLIT", ov, 2) && ov[0] > 0 && ov[1] > 0 && ov[0] + ov[1] < ctas) {
        *n1 = ov[0]; *n2 = ov[1]; *n3 = ctas - ov[0] - ov[1];
        return;
    }
    const double in_b = (double)g.q * g.B * 2, xf_b = (double)g.F * g.Kp * 2, yf_b = (double)g.F * 2 * g.p * 2, out_b = (double)g.p * g.B * 2;
    const double t1 = (in_b + xf_b) / 50.0, t2 = (xf_b + yf_b) / 54.0, t3 = (yf_b + out_b) / 82.0;
    const double tot = t1 + t2 + t3;
    int a = (int)(ctas * t1 / tot + 0.5), b = (int)(ctas * t2 / tot + 0.5);
    if (a < 1) a = 1;
    if (b < 1) b = 1;
    if (a + b > ctas - 1) b = ctas - 1 - a;
    *n1 = a; *n2 = b; *n3 = ctas - a - b;
}

// counters: [done1 | done2 | done3], m_tiles ints each, zeroed by the caller (stream-ordered) before this launch
cudaError_t launch_prefill_fused(const ProblemDev& P, const Geometry& g, const void* spec_gemm, void* Xf, void* Yf, int* status, int* counters, int ctas,
                                 cudaStream_t st) {
    const int nc = (int)P.N;
    StageLaunch<DftArgs> L1, L3;
    StageLaunch<GemmArgs> L2;
    if (!prep_dft_fwd(P, g, spec_gemm, 0, nc, Xf, status, L1)) return cudaErrorNotSupported;
    if (!prep_gemm(Xf, spec_gemm, Yf, status, g, nc, L2)) return cudaErrorNotSupported;
    if (!prep_dft_inv(P, g, spec_gemm, 0, nc, Yf, status, L3)) return cudaErrorNotSupported;
    const int m_tiles = L1.a.m_tiles;
    int* done1 = counters;
    int* done2 = counters + m_tiles;
    int* done3 = counters + 2 * m_tiles;
    // keep roughly 48 MB of Xf + Yf in flight between a producer group and its consumer
    int lag = (int)((48.0 * 1024 * 1024) / ((double)TILE_M * g.F * (g.Kp + 2 * g.p) * 2));
    int ov[1];
    if (env_ints("BCL_FUSED_LAG", ov, 1)) lag = ov[0];
    if (lag < 4 && lag != 0) lag = 4;
    int ch[3] = {2, 2, 4};                                          // consecutive items per round: one poll / one signal per run of a tile
    int chov[3];
    if (env_ints("BCL_FUSED_CHUNK", chov, 3)) { ch[0] = chov[0]; ch[1] = chov[1]; ch[2] = chov[2]; }
    L1.a.deps = StageDeps{nullptr, done1, lag > 0 ? done2 : nullptr, 0, g.FG * g.NCH * EPI_WARPS, lag, 1, ch[0]};
    L2.a.deps = StageDeps{done1, done2, lag > 0 ? done3 : nullptr, g.KG * DFT_EPI_WARPS, g.p * DFT_EPI_WARPS, lag, 1, ch[1]};
    L3.a.deps = StageDeps{done2, lag > 0 ? done3 : nullptr, nullptr, g.FG * g.NCH * EPI_WARPS, 0, 0, 1, ch[2]};
    int n1, n2, n3;
    fused_split(g, ctas, &n1, &n2, &n3);
    size_t smem = L1.smem > L2.smem ? L1.smem : L2.smem;
    if (L3.smem > smem) smem = L3.smem;
    cudaError_t e;
    if (g.B == 128) {
        if ((e = cudaFuncSetAttribute(prefill_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        return launch_pdl(prefill_fused_kernel<128>, (unsigned)ctas, FUSED_THREADS, smem, st, L1.tm_in, L1.tm_w, L1.tm_out, L2.tm_in, L2.tm_w, L2.tm_out,
                          L3.tm_in, L3.tm_w, L3.tm_out, L1.a, L2.a, L3.a, n1, n2);
    }
    if ((e = cudaFuncSetAttribute(prefill_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    return launch_pdl(prefill_fused_kernel<64>, (unsigned)ctas, FUSED_THREADS, smem, st, L1.tm_in, L1.tm_w, L1.tm_out, L2.tm_in, L2.tm_w, L2.tm_out,
                      L3.tm_in, L3.tm_w, L3.tm_out, L1.a, L2.a, L3.a, n1, n2);
}

}  // namespace gemm
}  // namespace bcl
