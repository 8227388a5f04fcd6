// bcl_gemm.cuh -- geometry and intermediate layouts of the prefill (large-N) path, shared by bcl_gemm.cu (per-frequency
// tcgen05 GEMM, CUDA-core FFT stages, spectrum precompute, chunked chaining) and bcl_dft_tc.cu (DFT-as-GEMM stages).
//
// Everything between the stages is bf16 and laid out so that BOTH the producer's stores and the consumer's TMA loads are
// runs of whole 16-byte chunks:
//   Xf  [F][KG][ntp][8]       Xf[f][kg][n][k0],  k = 8 kg + k0 = 2 i + {re,im}          (A operand of the frequency GEMM)
//   Yf  [FG][p][ntp][8]       Yf[fg][j][n][2 fl + {re,im}],  f = 4 fg + fl               (A operand of the inverse DFT)
// ntp = tokens of the chunk rounded up to 32 (token_pitch): consecutive tokens of one (f,kg) / (fg,j) slab are contiguous,
// so every TMA box row is 32 tokens x 16 B = 512 contiguous bytes (TMA moves about one box row per clock: 16-byte rows
// capped all three kernels at ~27 GB/s per SM, profiles/r1_prefill_notes.md).
//   GEMM-layout spectrum buffer:
//     Cmat [F][Np][Kp]        K-major rows, n = 2 j + {re,im} (Np = 2p rounded up to 64), k as above (Kp = 2q rounded up to 16)
//     Wfwd fp16, Wfwd bf16    [B][B]  row r = packed spectrum real (0: DC, 1: Nyquist, 2f: Re, 2f+1: Im), column t   (forward DFT)
//     Winv bf16               [B][B]  row t, column r, 1/B scaling folded in                                          (inverse DFT)
// F = B/2 packed frequency slots, FG = F/4, KG = Kp/8.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "bcl_kernels.h"

namespace bcl {
namespace gemm {

constexpr int TILE_M = 128;          // tokens per MMA tile (UMMA M)
constexpr int KATOM = 64;            // 16-bit elements per 128-byte swizzle row
constexpr int NCHUNK = 64;           // columns (= 32 output blocks) of one frequency-GEMM tile

struct Geometry {
    int B, p, q, split;
    int F, FG, Kp, KG, Np, NCH;
    // byte offsets inside the GEMM-layout spectrum buffer; the *_lo parts exist only in the split (fp32-accurate) layout
    size_t cmat_bytes, cmat_lo_off, wf16_off, wbf16_off, winv_off, wf16_lo_off, wbf16_lo_off, winv_lo_off, total_bytes;
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// split = 1 (fp32 spectra on the tensor cores): every bf16 operand of the frequency GEMM is a (hi, lo) pair, v = hi + lo to
// ~2^-17, and A*B is evaluated as Ahi*Bhi + Ahi*Blo + Alo*Bhi with fp32 accumulation -- the 1e-4 parity class of the
// reference's complex64 contraction (src/patch_llama_fft.py:260).  Xf / Yf are (hi, lo) pairs too (lo buffer right after
// hi); the rfft / irfft stages then run on the CUDA cores in fp32 (register FFT of bcl_fft.cuh).
__host__ inline Geometry geometry(int p, int q, int B, int split = 0) {
    Geometry g;
    g.B = B; g.p = p; g.q = q; g.split = split;
    g.F = B / 2;
    g.FG = g.F / 4;
    g.Kp = round_up(2 * q, 16);
    g.KG = g.Kp / 8;
    g.Np = round_up(2 * p, NCHUNK);
    g.NCH = g.Np / NCHUNK;
    g.cmat_bytes = ((size_t)g.F * g.Np * g.Kp * 2 + 1023) & ~(size_t)1023;
    g.cmat_lo_off = g.cmat_bytes;
    g.wf16_off = g.cmat_bytes * (split ? 2 : 1);
    g.wbf16_off = g.wf16_off + (size_t)B * B * 2;
    g.winv_off = g.wbf16_off + (size_t)B * B * 2;
    g.wf16_lo_off = g.winv_off + (size_t)B * B * 2;                 // rounding residues of the three DFT matrices (split only)
    g.wbf16_lo_off = g.wf16_lo_off + (size_t)B * B * 2;
    g.winv_lo_off = g.wbf16_lo_off + (size_t)B * B * 2;
    g.total_bytes = split ? g.winv_lo_off + (size_t)B * B * 2 : g.wf16_lo_off;
    return g;
}

constexpr int TOK_GROUP = 32;        // tokens per TMA box row of the chunked layouts
__host__ __device__ inline int64_t token_pitch(int64_t ntok) { return (ntok + TOK_GROUP - 1) / TOK_GROUP * TOK_GROUP; }
inline size_t xf_bytes(const Geometry& g, int64_t ntok) { return (size_t)g.F * g.Kp * token_pitch(ntok) * 2; }
inline size_t yf_bytes(const Geometry& g, int64_t ntok) { return (size_t)g.F * 2 * g.p * token_pitch(ntok) * 2; }
// tensor maps over the chunked layouts: dims (256 elements = 32 tokens x 8, token groups, slabs (kg or j), f or fg)
bool make_chunk_map(CUtensorMap* m, const void* base, int64_t ntok, int slabs, int outer, uint32_t box_groups, uint32_t box_slabs, uint32_t box_outer);

// Tensor-map helper (cuTensorMapEncodeTiled through the runtime's driver entry point; no libcuda link dependency).
// dims/strides innermost first; strides[i] = byte stride of dimension i+1.
bool make_tensor_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, CUtensorMapSwizzle swz);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of on every launch (~1-2 us of host time each;
// a 256-token forward is three launches and host-bound).  Defined in bcl_gemm.cu.
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel's prologue overlaps the previous kernel's tail.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// DFT-as-GEMM stages on tcgen05 (bcl_dft_tc.cu).  B in {64, 128}; 16-bit activations.
bool dft_tc_supported(int B, int act_dtype);
// g.split: Xf_lo / Yf_lo are the lo halves of the (hi, lo) intermediates (NULL otherwise)
cudaError_t launch_dft_fwd_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, void* Xf_lo, int* status, int max_ctas,
                              cudaStream_t st);
cudaError_t launch_dft_inv_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo, int* status,
                              int max_ctas, cudaStream_t st);

// gate | up stacked along the output-block axis: inverse DFT + SWIGLU epilogue, P.y = h with (g.p / 2) * B columns (B <= 128, bf16 spectra)
cudaError_t launch_dft_inv_swiglu_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo,
                                     int* status, int max_ctas, cudaStream_t st);

// Single-launch pipeline of the three tensor-core stages over L2-resident intermediates (bcl_prefill_fused.cu).
// counters: 3 * ceil(N/128) ints, zeroed in stream order before the launch; Xf / Yf sized for all N tokens.
size_t fused_counter_bytes(int64_t N);
// The three tensor-core stages of one token chunk in ONE launch with grid barriers in between (bcl_prefill_fused.cu).
cudaError_t launch_prefill_chain(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, void* Xf_lo, void* Yf,
                                 void* Yf_lo, int* status, int* grid_ctr, int ctas, cudaStream_t st);
cudaError_t launch_prefill_fused(const ProblemDev& P, const Geometry& g, const void* spec_gemm, void* Xf, void* Yf, int* status, int* counters, int ctas,
                                 cudaStream_t st);


// Wave schedule (bcl_prefill_wave.cu): ONE launch, every CTA interleaves forward-DFT / GEMM / inverse-DFT items `lag` token tiles
// apart over L2-resident rings of Xf / Yf.  area: workspace bytes for the rings; counters: 3 * ceil(N/128) ints, zeroed before.
bool wave_supported(const ProblemDev& P, const Geometry& g);
cudaError_t launch_prefill_wave(const ProblemDev& P, const Geometry& g, const void* spec_gemm, void* area, size_t area_bytes, int* status, int* counters,
                                int ctas, cudaStream_t st);

}  // namespace gemm
}  // namespace bcl
