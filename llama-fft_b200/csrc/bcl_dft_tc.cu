// bcl_dft_tc.cu -- the rfft / irfft stages of the prefill path as dense DFT-matrix GEMMs on the tcgen05 tensor cores.
//
// reference: rfft(x32) src/patch_llama_fft.py:252 and irfft(Yf, n=B) + cast + bias :263-271.
// A length-B real DFT of 128 tokens at once is [128 x B] * [B x B]: 2 B^2 flops per block instead of ~2.5 B log2 B, but on
// a pipe that is ~50x faster than the CUDA cores, fed by TMA directly from the 16-bit activations (no conversion pass)
// and drained by TMA stores, so both stages run at memory speed instead of instruction-issue speed.
//
//   dft_fwd_tc_kernel  item = (128-token tile, group kg of 4 input blocks)
//        D_ib[128 x B] = x[tile, block 4kg+ib][128 x B] * Wfwd^T        4 TMEM accumulators (4 B <= 512 columns)
//        epilogue: thread = token; 16-byte chunk (f, kg, token) = {Re,Im} of the 4 blocks -> smem -> TMA store into Xf
//   dft_inv_tc_kernel  item = (128-token tile, output block j)
//        D[128 x B] = Yf[tile, j][128 x B] * Winv^T                     512/B accumulators in flight
//        epilogue: round to x.dtype (:266), + bias in that dtype (:270-271), 128-byte-swizzled smem tile -> TMA store into y
// Operands: activations fp16 or bf16 (Wfwd in the same format), Yf / Winv bf16, fp32 accumulation in TMEM.
#include <cstdlib>
#include "bcl_common.cuh"
#include "bcl_sm100.cuh"
#include "bcl_gemm.cuh"
#include "bcl_prefill_bodies.cuh"

namespace bcl {
namespace gemm {

using namespace sm100;

template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ CUtensorMap tmXf, const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_fwd_body<B, false>(dft_smem, tmX, tmW, tmXf, nullptr, a, (int)blockIdx.x, (int)gridDim.x);
}
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_tc_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_inv_body<B, false>(dft_smem, tmYf, tmW, tmOut, nullptr, a, (int)blockIdx.x, (int)gridDim.x);
}
// fp32-accurate variants: (hi, lo) DFT matrices and (hi, lo) intermediates
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_fwd_tc_split_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                                           const __grid_constant__ CUtensorMap tmXf, const __grid_constant__ DftLoMaps lo,
                                                                           const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_fwd_body<B, true>(dft_smem, tmX, tmW, tmXf, &lo, a, (int)blockIdx.x, (int)gridDim.x);
}
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_tc_split_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                           const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftLoMaps lo,
                                                                           const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_inv_body<B, true>(dft_smem, tmYf, tmW, tmOut, &lo, a, (int)blockIdx.x, (int)gridDim.x);
}

// multi-destination variants (sharded layer, exchange="p2p"): every finished tile also goes to the peers' y
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_tc_bcast_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                           const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftExtraOut ex,
                                                                           const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_inv_body<B, false>(dft_smem, tmYf, tmW, tmOut, nullptr, a, (int)blockIdx.x, (int)gridDim.x, &ex);
}
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_tc_split_bcast_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftLoMaps lo,
                                                                                 const __grid_constant__ DftExtraOut ex, const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_inv_body<B, true>(dft_smem, tmYf, tmW, tmOut, &lo, a, (int)blockIdx.x, (int)gridDim.x, &ex);
}

// gate | up stacked: inverse DFT of both members of a block pair + SWIGLU epilogue (dft_inv_swiglu_body)
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_swiglu_tc_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_inv_swiglu_body<B, false>(dft_smem, tmYf, tmW, tmOut, nullptr, a, (int)blockIdx.x, (int)gridDim.x);
}
// fp32 spectra: (hi, lo) Yf and inverse DFT matrices
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_swiglu_tc_split_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftLoMaps lo,
                                                                                  const __grid_constant__ DftArgs a) {
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    dft_inv_swiglu_body<B, true>(dft_smem, tmYf, tmW, tmOut, &lo, a, (int)blockIdx.x, (int)gridDim.x);
}

// =========================================================================================================
// host side
// =========================================================================================================
bool dft_tc_supported(int B, int act_dtype) { return (B == 64 || B == 128 || B == 256) && (act_dtype == BCL_F16 || act_dtype == BCL_BF16); }
static int dft_nsplit(int B, bool split = false) { return B == 256 ? (split ? 4 : 2) : 1; }      // launches per stage (DftShape::NSPLIT)
static int dft_nb(int B, bool split = false) { return B / dft_nsplit(B, split); }                // accumulator columns per block and launch (DftShape::NB)
static int dft_parts(int B) { return B == 256 ? 2 : 1; }         // K parts = smem stages per block (DftShape<B>::PARTS)

// nbuf: staging buffers per epilogue warp; wmats: resident DFT matrices (1, or 2 = hi + lo)
static size_t dft_smem_bytes(int B, uint32_t stage_bytes, int stages, int nbuf = DFT_NBUF, int wmats = 1) {
    return 1024 + DFT_EPI_WARPS * nbuf * DFT_EPI_BUF + (size_t)wmats * (B / KATOM) * dft_nb(B, wmats == 2) * 128 + (size_t)stages * stage_bytes;
}
static int dft_pick_stages(int B, uint32_t stage_bytes, int nbuf = DFT_NBUF, int wmats = 1) {
    int s = DFT_MAX_STAGES;
    while (s > 1 && dft_smem_bytes(B, stage_bytes, s, nbuf, wmats) > 225 * 1024) --s;
    return s;
}

static bool map_w(CUtensorMap* m, const void* w, int B, CUtensorMapDataType dt, bool split = false) {   // [B rows][B k], box 64 x NB rows, 128-byte swizzle
    const uint64_t dims[2] = {(uint64_t)B, (uint64_t)B};
    const uint64_t strides[1] = {(uint64_t)B * 2};
    const uint32_t box[2] = {KATOM, (uint32_t)dft_nb(B, split)};
    return make_tensor_map(m, dt, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// debug only: BCL_TRACE_PTR=<device address of 512 int64> BCL_TRACE_STAGE=<1|2|3>
long long* trace_ptr(int stage) {
#ifdef BCL_ENABLE_TRACE
    const char* tp = getenv("BCL_TRACE_PTR");
    const char* ts = getenv("BCL_TRACE_STAGE");
    if (!tp || (ts ? atoi(ts) : 1) != stage) return nullptr;
    return reinterpret_cast<long long*>(strtoull(tp, nullptr, 0));
#else
    (void)stage;
    return nullptr;
#endif
}

bool prep_dft_fwd(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, int* status, StageLaunch<DftArgs>& L,
                  void* Xf_lo, DftLoMaps* lo) {
    const bool split = lo != nullptr;
    const int B = g.B;
    const bool bf = P.x_dtype == BCL_BF16;
    const CUtensorMapDataType adt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    {
        const uint64_t dims[2] = {(uint64_t)g.q * B, (uint64_t)P.N};
        const uint64_t strides[1] = {(uint64_t)P.x_stride * 2};
        const uint32_t box[2] = {KATOM, TILE_M};
        if (!make_tensor_map(&L.tm_in, adt, 2, P.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
    }
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(spec_gemm);
    if (!map_w(&L.tm_w, sb + (bf ? g.wbf16_off : g.wf16_off), B, adt, split)) return false;
    if (!make_chunk_map(&L.tm_out, Xf, nc, g.KG, g.F, 1, 1, 8)) return false;
    DftArgs& a = L.a;
    a = DftArgs{};
    a.status = status; a.bias = nullptr; a.n_begin = n_begin; a.nc = nc;
    a.m_tiles = (nc + TILE_M - 1) / TILE_M; a.groups = g.KG; a.num_items = a.m_tiles * g.KG; a.act_bf16 = bf ? 1 : 0;
    a.trace = trace_ptr(1);
    const uint32_t stage_bytes = (uint32_t)(B / KATOM / dft_parts(B)) * TILE_M * 128;
    a.stages = dft_pick_stages(B, stage_bytes, DFT_NBUF, split ? 2 : 1);
    L.smem = dft_smem_bytes(B, stage_bytes, a.stages, DFT_NBUF, split ? 2 : 1);
    if (split) {
        if (!map_w(&lo->W, sb + (bf ? g.wbf16_lo_off : g.wf16_lo_off), B, adt, true)) return false;
        if (!make_chunk_map(&lo->IO, Xf_lo, nc, g.KG, g.F, 1, 1, 8)) return false;
    }
    return true;
}

bool prep_dft_inv(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, StageLaunch<DftArgs>& L,
                  const void* Yf_lo, DftLoMaps* lo, DftExtraOut* ex) {
    const bool split = lo != nullptr;
    const int B = g.B;
    const bool bf = P.x_dtype == BCL_BF16;
    const CUtensorMapDataType adt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (!make_chunk_map(&L.tm_in, Yf, nc, g.p, g.FG, TILE_M / TOK_GROUP, 1, (uint32_t)(g.FG / dft_parts(B)))) return false;
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(spec_gemm);
    if (!map_w(&L.tm_w, sb + g.winv_off, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, split)) return false;
    {
        const uint64_t dims[2] = {(uint64_t)g.p * B, (uint64_t)P.N};
        const uint64_t strides[1] = {(uint64_t)P.y_stride * 2};
        const uint32_t box[2] = {KATOM, 32};
        if (!make_tensor_map(&L.tm_out, adt, 2, P.y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
        if (ex != nullptr) {
            ex->n = P.n_extra;
            for (int e = 0; e < P.n_extra; ++e)
                if (!make_tensor_map(&ex->out[e], adt, 2, P.y_extra[e], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
        }
    }
    DftArgs& a = L.a;
    a = DftArgs{};
    a.status = status; a.bias = P.bias; a.n_begin = n_begin; a.nc = nc;
    a.m_tiles = (nc + TILE_M - 1) / TILE_M; a.groups = g.p; a.num_items = a.m_tiles * g.p; a.act_bf16 = bf ? 1 : 0;
    a.trace = trace_ptr(3);
    const uint32_t stage_bytes = (uint32_t)(g.FG / dft_parts(B)) * 2048 * (split ? 2 : 1);
    a.stages = dft_pick_stages(B, stage_bytes, split ? 1 : DFT_NBUF, split ? 2 : 1);
    L.smem = dft_smem_bytes(B, stage_bytes, a.stages, split ? 1 : DFT_NBUF, split ? 2 : 1);
    if (split) {
        if (!map_w(&lo->W, sb + g.winv_lo_off, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, true)) return false;
        if (!make_chunk_map(&lo->IO, Yf_lo, nc, g.p, g.FG, TILE_M / TOK_GROUP, 1, (uint32_t)(g.FG / dft_parts(B)))) return false;
    }
    return true;
}

bool prep_dft_inv_swiglu(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, StageLaunch<DftArgs>& L,
                         const void* Yf_lo, DftLoMaps* lo) {
    const int B = g.B;
    const bool split = lo != nullptr;
    if ((B != 64 && B != 128) || (g.p & 1) || (g.split != 0) != split || (split && Yf_lo == nullptr)) return false;
    const bool bf = P.x_dtype == BCL_BF16;
    const CUtensorMapDataType adt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (!make_chunk_map(&L.tm_in, Yf, nc, g.p, g.FG, TILE_M / TOK_GROUP, 1, (uint32_t)g.FG)) return false;
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(spec_gemm);
    if (!map_w(&L.tm_w, sb + g.winv_off, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, split)) return false;
    const uint64_t dims[2] = {(uint64_t)(g.p / 2) * B, (uint64_t)P.N};
    const uint64_t strides[1] = {(uint64_t)P.y_stride * 2};
    const uint32_t box[2] = {KATOM, 32};
    if (!make_tensor_map(&L.tm_out, adt, 2, P.y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
    DftArgs& a = L.a;
    a = DftArgs{};
    a.status = status; a.bias = nullptr; a.n_begin = n_begin; a.nc = nc;
    a.m_tiles = (nc + TILE_M - 1) / TILE_M; a.groups = g.p / 2; a.num_items = a.m_tiles * (g.p / 2); a.act_bf16 = bf ? 1 : 0;
    const uint32_t stage_bytes = (uint32_t)g.FG * 2048 * (split ? 2 : 1);
    a.stages = dft_pick_stages(B, stage_bytes, split ? 1 : DFT_NBUF, split ? 2 : 1) & ~1;   // an item takes two stages
    if (a.stages < 2) return false;
    L.smem = dft_smem_bytes(B, stage_bytes, a.stages, split ? 1 : DFT_NBUF, split ? 2 : 1);
    if (split) {
        if (!map_w(&lo->W, sb + g.winv_lo_off, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, true)) return false;
        if (!make_chunk_map(&lo->IO, Yf_lo, nc, g.p, g.FG, TILE_M / TOK_GROUP, 1, (uint32_t)g.FG)) return false;
    }
    return true;
}

cudaError_t launch_dft_inv_swiglu_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo,
                                     int* status, int max_ctas, cudaStream_t st);

template <typename K, typename... Extra>
static cudaError_t launch_stage(K kernel, const StageLaunch<DftArgs>& L, int max_ctas, cudaStream_t st, const Extra&... extra) {
    const int grid = L.a.num_items < max_ctas ? L.a.num_items : max_ctas;
    if (grid < 1) return cudaSuccess;
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), L.smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(kernel, (unsigned)grid, DFT_THREADS, L.smem, st, L.tm_in, L.tm_w, L.tm_out, extra..., L.a);
}

cudaError_t launch_dft_fwd_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, void* Xf_lo, int* status, int max_ctas,
                              cudaStream_t st) {
    StageLaunch<DftArgs> L;
    if (g.split && Xf_lo != nullptr) {
        DftLoMaps lo;
        if (!prep_dft_fwd(P, g, spec_gemm, n_begin, nc, Xf, status, L, Xf_lo, &lo)) return cudaErrorNotSupported;
        if (g.B == 256) {                                           // four launches: 64 spectrum rows each (DftShape<256, true>)
            for (int part = 0; part < 4; ++part) {
                L.a.half = part;
                cudaError_t e = launch_stage(dft_fwd_tc_split_kernel<256>, L, max_ctas, st, lo);
                if (e != cudaSuccess) return e;
            }
            return cudaSuccess;
        }
        return g.B == 128 ? launch_stage(dft_fwd_tc_split_kernel<128>, L, max_ctas, st, lo) : launch_stage(dft_fwd_tc_split_kernel<64>, L, max_ctas, st, lo);
    }
    if (!prep_dft_fwd(P, g, spec_gemm, n_begin, nc, Xf, status, L)) return cudaErrorNotSupported;
    if (g.B == 256) {                                               // two launches: spectrum rows [0,128) and [128,256) (DftShape)
        for (int half = 0; half < 2; ++half) {
            L.a.half = half;
            cudaError_t e = launch_stage(dft_fwd_tc_kernel<256>, L, max_ctas, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    return g.B == 128 ? launch_stage(dft_fwd_tc_kernel<128>, L, max_ctas, st) : launch_stage(dft_fwd_tc_kernel<64>, L, max_ctas, st);
}

cudaError_t launch_dft_inv_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo, int* status,
                              int max_ctas, cudaStream_t st) {
    StageLaunch<DftArgs> L;
    const bool bcast = P.n_extra > 0;                               // peer destinations: the *_bcast kernels store every tile to all of them
    DftExtraOut ex;
    if (g.split && Yf_lo != nullptr) {
        DftLoMaps lo;
        if (!prep_dft_inv(P, g, spec_gemm, n_begin, nc, Yf, status, L, Yf_lo, &lo, bcast ? &ex : nullptr)) return cudaErrorNotSupported;
        if (g.B == 256) {                                           // four launches: 64 output samples of every block each
            for (int part = 0; part < 4; ++part) {
                L.a.half = part;
                cudaError_t e = bcast ? launch_stage(dft_inv_tc_split_bcast_kernel<256>, L, max_ctas, st, lo, ex) : launch_stage(dft_inv_tc_split_kernel<256>, L, max_ctas, st, lo);
                if (e != cudaSuccess) return e;
            }
            return cudaSuccess;
        }
        if (bcast) return g.B == 128 ? launch_stage(dft_inv_tc_split_bcast_kernel<128>, L, max_ctas, st, lo, ex) : launch_stage(dft_inv_tc_split_bcast_kernel<64>, L, max_ctas, st, lo, ex);
        return g.B == 128 ? launch_stage(dft_inv_tc_split_kernel<128>, L, max_ctas, st, lo) : launch_stage(dft_inv_tc_split_kernel<64>, L, max_ctas, st, lo);
    }
    if (!prep_dft_inv(P, g, spec_gemm, n_begin, nc, Yf, status, L, nullptr, nullptr, bcast ? &ex : nullptr)) return cudaErrorNotSupported;
    if (g.B == 256) {                                               // two launches: output samples [0,128) and [128,256) of every block
        for (int half = 0; half < 2; ++half) {
            L.a.half = half;
            cudaError_t e = bcast ? launch_stage(dft_inv_tc_bcast_kernel<256>, L, max_ctas, st, ex) : launch_stage(dft_inv_tc_kernel<256>, L, max_ctas, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (bcast) return g.B == 128 ? launch_stage(dft_inv_tc_bcast_kernel<128>, L, max_ctas, st, ex) : launch_stage(dft_inv_tc_bcast_kernel<64>, L, max_ctas, st, ex);
    return g.B == 128 ? launch_stage(dft_inv_tc_kernel<128>, L, max_ctas, st) : launch_stage(dft_inv_tc_kernel<64>, L, max_ctas, st);
}


cudaError_t launch_dft_inv_swiglu_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo,
                                     int* status, int max_ctas, cudaStream_t st) {
    StageLaunch<DftArgs> L;
    if (g.split) {
        DftLoMaps lo;
        if (!prep_dft_inv_swiglu(P, g, spec_gemm, n_begin, nc, Yf, status, L, Yf_lo, &lo)) return cudaErrorNotSupported;
        return g.B == 128 ? launch_stage(dft_inv_swiglu_tc_split_kernel<128>, L, max_ctas, st, lo) : launch_stage(dft_inv_swiglu_tc_split_kernel<64>, L, max_ctas, st, lo);
    }
    if (!prep_dft_inv_swiglu(P, g, spec_gemm, n_begin, nc, Yf, status, L)) return cudaErrorNotSupported;
    return g.B == 128 ? launch_stage(dft_inv_swiglu_tc_kernel<128>, L, max_ctas, st) : launch_stage(dft_inv_swiglu_tc_kernel<64>, L, max_ctas, st);
}

}  // namespace gemm
}  // namespace bcl
