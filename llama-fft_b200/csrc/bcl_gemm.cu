// bcl_gemm.cu -- the prefill (large-N) BC-linear path: per-frequency complex GEMM on the tcgen05 tensor cores.
//
// reference: src/patch_llama_fft.py:252-263.  For many tokens the middle step of the forward,
//     Yf[n,o,f] = sum_i Cf[o,i,f] * Xf[n,i,f]            (torch.einsum("oif,nif->nof"), :260)
// is, per frequency f, a dense complex (tokens x q) * (q x p) product.  Embedded in the reals
//     [Yr Yi] = [Xr Xi] * [[Cr, Ci], [-Ci, Cr]]
// it is one bf16 GEMM per frequency with K = 2q, N = 2p and fp32 accumulation -- tensor-core work.  The packed (DC,
// Nyquist) slot of the half-spectrum is the same GEMM with the block-diagonal matrix [[C_dc, 0], [0, C_ny]].
//
// Three stages per token chunk, staged through an L2-resident workspace (layouts: bcl_gemm.cuh):
//   stage 1  x (N,in) -> Xf      rfft (:249-252).  dft_fwd_tc_kernel (DFT as a tcgen05 GEMM, bcl_dft_tc.cu) for 16-bit
//                                activations and B <= 128, else rfft_rows_kernel (CUDA cores, register FFT of bcl_fft.cuh)
//   stage 2  Xf, Cmat -> Yf      freq_gemm_kernel: tcgen05.mma, TMA-fed smem ring, TMEM accumulators, TMA stores (:260)
//   stage 3  Yf -> y (N,out)     irfft + the reference epilogue (:263-271): dft_inv_tc_kernel or irfft_rows_kernel
// The GEMM-layout spectrum (Cmat + the DFT matrices) is built once per parameter version by spectrum_gemm_kernel /
// dft_matrix_kernel (replaces rfft(c), :256-257).
#include <mutex>
#include <map>
#include <utility>
#include <cstdlib>
#include "bcl_stream_parts.cuh"
#include "bcl_sm100.cuh"
#include "bcl_gemm.cuh"
#include "bcl_prefill_bodies.cuh"

namespace bcl {
namespace gemm {

using namespace sm100;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// (hi, lo) split of two floats: hi = bf16(v), lo = bf16(v - hi); hi + lo reproduces v to ~2^-17 relative
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    __nv_bfloat162 h; h.x = ha; h.y = hb;
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = pack_bf16(a - __bfloat162float(ha), b - __bfloat162float(hb));
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

// ---------------------------------------------------------------------------------------------------------
// CUDA-core stage 1: rfft of NT tokens (all q input blocks) -> 16-byte chunks of the per-frequency A operand
// ---------------------------------------------------------------------------------------------------------
template <int B, int NT, int WARPS> struct RowsSmem {
    static constexpr int PITCH = B + 2;                                // floats per block: spreads (block, frequency) over the banks
    static constexpr size_t tw_bytes = (size_t)(B / 2) * 8;
    static constexpr size_t fscr_bytes = (size_t)WARPS * FftCfg<B>::WARP_SCRATCH_BYTES;
    static size_t bytes(int blocks) { return tw_bytes + fscr_bytes + (size_t)NT * blocks * PITCH * 4; }
};

template <int B, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) rfft_rows_kernel(const __grid_constant__ ProblemDev P, int64_t n_begin, int nc,
                                                                __nv_bfloat16* __restrict__ Xf, __nv_bfloat16* __restrict__ Xf_lo, int KG) {
    using SM = RowsSmem<B, NT, WARPS>;
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char rows_smem[];
    float2* tw = reinterpret_cast<float2*>(rows_smem);
    float2* fscr_all = reinterpret_cast<float2*>(rows_smem + SM::tw_bytes);
    float* xs = reinterpret_cast<float*>(rows_smem + SM::tw_bytes + SM::fscr_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    const int q = P.q;
    const int t0 = blockIdx.x * NT;                                     // first token of this CTA, relative to the chunk
    ProblemDev Pc = P;
    Pc.N = n_begin + nc;                                                // tokens past the chunk are inactive
    load_x_spectrum<B, NT, WARPS, SM::PITCH>(Pc, n_begin + t0, xs, tw, fscr_all + (size_t)warp * (C::WARP_SCRATCH_BYTES / 8), lane, warp);
    __syncthreads();
    constexpr int F = B / 2;
    const size_t ntp = (size_t)token_pitch(nc);
    // one 16-byte chunk = (f, kg, token): input blocks 4kg .. 4kg+3, {re,im} each.  Tokens of a CTA are adjacent chunks.
    const int nchunks = F * KG * NT;
    for (int id = threadIdx.x; id < nchunks; id += WARPS * 32) {
        const int n = id % NT;
        const int kg = (id / NT) % KG;
        const int f = id / (NT * KG);
        if (t0 + n >= nc) continue;
        uint32_t w[4], wl[4];
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) {
            const int i = 4 * kg + ib;
            w[ib] = 0u; wl[ib] = 0u;
            if (i < q) {
                const float2 v = *reinterpret_cast<const float2*>(xs + ((size_t)n * q + i) * SM::PITCH + 2 * f);
                split_bf16(v.x, v.y, w[ib], wl[ib]);                    // f = 0: (DC, Nyquist), both real
            }
        }
        const size_t off = (((size_t)f * KG + kg) * ntp + (t0 + n)) * 8;
        *reinterpret_cast<uint4*>(Xf + off) = make_uint4(w[0], w[1], w[2], w[3]);
        if (Xf_lo != nullptr) *reinterpret_cast<uint4*>(Xf_lo + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// GEMM-layout spectrum: c (p,q,B) fp32 -> Cmat[F][Np][Kp] bf16.  One CTA per output block j (rows 2j, 2j+1 of every f).
// ---------------------------------------------------------------------------------------------------------
template <int B, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spectrum_gemm_kernel(const float* __restrict__ c, int p, int q, __nv_bfloat16* __restrict__ Cmat,
                                                                   __nv_bfloat16* __restrict__ Cmat_lo, int Np, int Kp) {
    using SM = RowsSmem<B, 1, WARPS>;
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char rows_smem[];
    float2* tw = reinterpret_cast<float2*>(rows_smem);
    float2* fscr_all = reinterpret_cast<float2*>(rows_smem + SM::tw_bytes);
    float* xs = reinterpret_cast<float*>(rows_smem + SM::tw_bytes + SM::fscr_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x;
    constexpr int F = B / 2;
    if (j < p) {
        fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
        __syncthreads();
        ProblemDev P{};
        P.x = c; P.x_dtype = BCL_F32; P.x_stride = (int64_t)q * B; P.N = p; P.q = q;
        load_x_spectrum<B, 1, WARPS, SM::PITCH>(P, j, xs, tw, fscr_all + (size_t)warp * (C::WARP_SCRATCH_BYTES / 8), lane, warp);
        __syncthreads();
    }
    for (int f = warp; f < F; f += WARPS) {
        uint32_t* row_re = reinterpret_cast<uint32_t*>(Cmat + ((size_t)f * Np + 2 * j) * Kp);       // n = 2j  : Yr
        uint32_t* row_im = row_re + Kp / 2;                                                         // n = 2j+1: Yi
        const size_t lo_shift = Cmat_lo != nullptr ? (size_t)(Cmat_lo - Cmat) / 2 : 0;                // in 32-bit words
        for (int i = lane; i < Kp / 2; i += 32) {
            uint32_t wr = 0u, wi = 0u, lr = 0u, li = 0u;
            if (j < p && i < q) {
                const float2 v = *reinterpret_cast<const float2*>(xs + (size_t)i * SM::PITCH + 2 * f);
                if (f == 0) { split_bf16(v.x, 0.f, wr, lr); split_bf16(0.f, v.y, wi, li); }          // [[C_dc, 0], [0, C_ny]]
                else { split_bf16(v.x, -v.y, wr, lr); split_bf16(v.y, v.x, wi, li); }                // Yr = Xr Cr - Xi Ci ; Yi = Xr Ci + Xi Cr
            }
            row_re[i] = wr;
            row_im[i] = wi;
            if (Cmat_lo != nullptr) { row_re[i + lo_shift] = lr; row_im[i + lo_shift] = li; }
        }
    }
}

// DFT matrices for the tensor-core stages.  grid B, block B: thread (row a, column b).
//   Wfwd[r][t]  r = 0: 1, r = 1: (-1)^t, r = 2f: cos(2 pi f t / B), r = 2f+1: -sin(2 pi f t / B)              (rfft, :252)
//   Winv[t][r]  the transpose, times 1/B for r < 2 and 2/B otherwise                                          (irfft n=B, :263)
__global__ void dft_matrix_kernel(int B, __half* __restrict__ wf16, __nv_bfloat16* __restrict__ wbf16, __nv_bfloat16* __restrict__ winv,
                                  __half* __restrict__ wf16_lo, __nv_bfloat16* __restrict__ wbf16_lo, __nv_bfloat16* __restrict__ winv_lo) {
    const int a = blockIdx.x, b = threadIdx.x;
    auto entry = [B](int r, int t) -> float {
        if (r == 0) return 1.f;
        if (r == 1) return (t & 1) ? -1.f : 1.f;
        const int f = r >> 1;
        const int m = (f * t) & (B - 1);                                // angle 2 pi m / B, m exact
        float s, c;
        sincospif(2.0f * (float)m / (float)B, &s, &c);
        return (r & 1) ? -s : c;
    };
    const size_t o = (size_t)a * B + b;
    const float fw = entry(a, b);                                       // Wfwd[a][b]
    const __half h16 = __float2half_rn(fw);
    const __nv_bfloat16 hbf = __float2bfloat16_rn(fw);
    wf16[o] = h16;
    wbf16[o] = hbf;
    const float iv = entry(b, a) * (b < 2 ? 1.0f : 2.0f) / (float)B;    // Winv[a = t][b = r]
    const __nv_bfloat16 hiv = __float2bfloat16_rn(iv);
    winv[o] = hiv;
    if (wf16_lo != nullptr) {                                           // split layout: the rounding residues
        wf16_lo[o] = __float2half_rn(fw - __half2float(h16));
        wbf16_lo[o] = __float2bfloat16_rn(fw - __bfloat162float(hbf));
        winv_lo[o] = __float2bfloat16_rn(iv - __bfloat162float(hiv));
    }
}

// ---------------------------------------------------------------------------------------------------------
// stage 2 stand-alone: the per-frequency tcgen05 GEMM (body: bcl_prefill_bodies.cuh)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1) freq_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ CUtensorMap tmY, const __grid_constant__ GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char gemm_smem[];
    freq_gemm_body<false>(gemm_smem, tmA, tmB, tmY, nullptr, a, (int)blockIdx.x, (int)gridDim.x);
}
// fp32-accurate variant: (hi, lo) bf16 operands, three MMAs per k-step (bcl_gemm.cuh: geometry(..., split = 1))
__global__ void __launch_bounds__(GEMM_THREADS, 1) freq_gemm_split_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                           const __grid_constant__ CUtensorMap tmY, const __grid_constant__ GemmLoMaps lo,
                                                                           const __grid_constant__ GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char gemm_smem[];
    freq_gemm_body<true>(gemm_smem, tmA, tmB, tmY, &lo, a, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-core stage 3: gather a token's output spectra from the 16-byte chunks, inverse rfft, reference epilogue
// ---------------------------------------------------------------------------------------------------------
template <int B, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) irfft_rows_kernel(const __grid_constant__ ProblemDev P, int64_t n_begin, int nc,
                                                                 const __nv_bfloat16* __restrict__ Yf, const __nv_bfloat16* __restrict__ Yf_lo) {
    using SM = RowsSmem<B, NT, WARPS>;
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char rows_smem[];
    float2* tw = reinterpret_cast<float2*>(rows_smem);
    float2* fscr_all = reinterpret_cast<float2*>(rows_smem + SM::tw_bytes);
    float* ys = reinterpret_cast<float*>(rows_smem + SM::tw_bytes + SM::fscr_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    const int p = P.p;
    const int t0 = blockIdx.x * NT;
    constexpr int FG = B / 8;
    const size_t ntp = (size_t)token_pitch(nc);
    const int nchunks = FG * p * NT;                                    // one 16-byte chunk = (fg, j, token): 4 frequencies x {re,im}
    for (int id = threadIdx.x; id < nchunks; id += WARPS * 32) {
        const int n = id % NT;
        const int j = (id / NT) % p;
        const int fg = id / (NT * p);
        if (t0 + n >= nc) continue;
        const size_t off = (((size_t)fg * p + j) * ntp + (t0 + n)) * 8;
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(Yf + off));
        float2* dst = reinterpret_cast<float2*>(ys + ((size_t)n * p + j) * SM::PITCH + 8 * fg);
        dst[0] = unpack_bf16(w.x); dst[1] = unpack_bf16(w.y); dst[2] = unpack_bf16(w.z); dst[3] = unpack_bf16(w.w);
        if (Yf_lo != nullptr) {                                         // split layout: value = hi + lo
            const uint4 l = __ldg(reinterpret_cast<const uint4*>(Yf_lo + off));
            const float2 l0 = unpack_bf16(l.x), l1 = unpack_bf16(l.y), l2 = unpack_bf16(l.z), l3 = unpack_bf16(l.w);
            dst[0].x += l0.x; dst[0].y += l0.y; dst[1].x += l1.x; dst[1].y += l1.y;
            dst[2].x += l2.x; dst[2].y += l2.y; dst[3].x += l3.x; dst[3].y += l3.y;
        }
    }
    __syncthreads();
    float2* fscr = fscr_all + (size_t)warp * (C::WARP_SCRATCH_BYTES / 8);
    const int g = lane / C::T, t = lane % C::T;
    const int nblk = NT * p;
    for (int b0 = warp * C::G; b0 < nblk; b0 += WARPS * C::G) {         // G = 32/T output blocks per warp pass
        const int blk = min(b0 + g, nblk - 1);
        const int n = blk / p, j = blk - n * p;
        float2 out[C::V];
        group_irfft_packed<B>(ys + (size_t)blk * SM::PITCH, fscr + g * C::GROUP_F2, tw, t, out);
        if (b0 + g < nblk && t0 + n < nc) store_block_epilogue<B>(P, n_begin + t0 + n, j, t, out);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = granted[std::make_pair(kernel, dev)];
    if (have >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

bool make_tensor_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
    return fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_chunk_map(CUtensorMap* m, const void* base, int64_t ntok, int slabs, int outer, uint32_t box_groups, uint32_t box_slabs, uint32_t box_outer) {
    const uint64_t ntp = (uint64_t)token_pitch(ntok);
    const uint64_t dims[4] = {(uint64_t)TOK_GROUP * 8, ntp / TOK_GROUP, (uint64_t)slabs, (uint64_t)outer};
    const uint64_t strides[3] = {(uint64_t)TOK_GROUP * 16, ntp * 16, (uint64_t)slabs * ntp * 16};
    const uint32_t box[4] = {TOK_GROUP * 8, box_groups, box_slabs, box_outer};
    return make_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

static size_t gemm_smem_bytes(int stages, int split) { return 1024 + EPI_WARPS * EPI_NBUF * EPI_BUF_BYTES + (size_t)stages * STAGE_BYTES * (split ? 2 : 1); }
static int gemm_pick_stages(int split) {
    int s = MAX_STAGES;
    while (s > 2 && gemm_smem_bytes(s, split) > 225 * 1024) --s;
    return s;
}

// Xf [F][KG][ntp][8], Cmat [F][Np][Kp] -> Yf [FG][p][ntp][8] (ntp = nc rounded up to 32)
// split layout (lo != nullptr): Xf_lo / Yf_lo follow Xf / Yf at lo_stride_x / lo_stride_y bytes, Cmat_lo at g.cmat_lo_off
static bool map_cmat(CUtensorMap* m, const void* Cmat, const Geometry& g) {
    const uint64_t dims[3] = {(uint64_t)g.Kp, (uint64_t)g.Np, (uint64_t)g.F};
    const uint64_t strides[2] = {(uint64_t)g.Kp * 2, (uint64_t)g.Kp * g.Np * 2};
    const uint32_t box[3] = {KATOM, NCHUNK, 1};
    return make_tensor_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, Cmat, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

bool prep_gemm_lo(const void* Xf_lo, const void* Cmat, void* Yf_lo, const Geometry& g, int nc, GemmLoMaps* lo) {
    if (!make_chunk_map(&lo->A, Xf_lo, nc, g.KG, g.F, TILE_M / TOK_GROUP, 8, 1)) return false;
    if (!map_cmat(&lo->B, reinterpret_cast<const unsigned char*>(Cmat) + g.cmat_lo_off, g)) return false;
    return make_chunk_map(&lo->Y, Yf_lo, nc, g.p, g.FG, 1, 8, 1);
}

bool prep_gemm(const void* Xf, const void* Cmat, void* Yf, int* status, const Geometry& g, int nc, StageLaunch<GemmArgs>& L, GemmLoMaps* lo) {
    (void)lo;
    if (!make_chunk_map(&L.tm_in, Xf, nc, g.KG, g.F, TILE_M / TOK_GROUP, 8, 1)) return false;
    if (!map_cmat(&L.tm_w, Cmat, g)) return false;
    if (!make_chunk_map(&L.tm_out, Yf, nc, g.p, g.FG, 1, 8, 1)) return false;
    GemmArgs& a = L.a;
    a = GemmArgs{};
    a.status = status;
    a.ntok = nc; a.p = g.p; a.KG = g.KG; a.FG = g.FG; a.k_atoms = (g.Kp + KATOM - 1) / KATOM; a.NCH = g.NCH;
    a.m_tiles = (nc + TILE_M - 1) / TILE_M;
    a.num_items = a.m_tiles * g.FG * g.NCH;
    a.stages = gemm_pick_stages(g.split);
    L.smem = gemm_smem_bytes(a.stages, g.split);
    a.trace = trace_ptr(2);
    return true;
}

template <typename K> static cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int B, int NT>
static cudaError_t launch_rfft_rows_t(const ProblemDev& P, int64_t n_begin, int nc, void* Xf, void* Xf_lo, int KG, cudaStream_t st) {
    constexpr int WARPS = 8;
    auto kern = rfft_rows_kernel<B, NT, WARPS>;
    const size_t smem = RowsSmem<B, NT, WARPS>::bytes(P.q);
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((nc + NT - 1) / NT), WARPS * 32, smem, st>>>(P, n_begin, nc, reinterpret_cast<__nv_bfloat16*>(Xf), reinterpret_cast<__nv_bfloat16*>(Xf_lo), KG);
    return cudaGetLastError();
}

template <int B, int NT>
static cudaError_t launch_irfft_rows_t(const ProblemDev& P, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo, cudaStream_t st) {
    constexpr int WARPS = 8;
    auto kern = irfft_rows_kernel<B, NT, WARPS>;
    const size_t smem = RowsSmem<B, NT, WARPS>::bytes(P.p);
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((nc + NT - 1) / NT), WARPS * 32, smem, st>>>(P, n_begin, nc, reinterpret_cast<const __nv_bfloat16*>(Yf), reinterpret_cast<const __nv_bfloat16*>(Yf_lo));
    return cudaGetLastError();
}

// tokens per CTA of the row kernels: two while the spectrum tile stays small enough for >= 3 CTAs per SM
static int rows_nt(int blocks, int B) { return (size_t)2 * blocks * (B + 2) * 4 <= 48 * 1024 ? 2 : 1; }

}  // namespace gemm

size_t gemm_spectrum_bytes(int p, int q, int B, int split) { return gemm::geometry(p, q, B, split).total_bytes; }

// workspace = [256-byte status area | Xf chunk | Yf chunk].  Measured on B200 (profiles/r1_prefill_notes.md): three
// full-size passes through HBM (no chunking, 16384 tokens: 157 us for 4096->4096) beat L2-resident 2048-3072-token chunks
// (228 us), because each per-chunk launch pays ~5 us of fixed cost (TMEM alloc, DFT-matrix load, pipeline fill, tail) for
// ~3 us of data movement.  So the default budget only chunks very large batches; BCL_PREFILL_CHUNK_MB overrides it.
static constexpr size_t kWsHeader = 256;

static size_t chunk_budget() {
    static size_t v = 0;
    if (!v) {
        const char* e = getenv("BCL_PREFILL_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        v = (size_t)(mb > 0 ? mb : 1024) << 20;
    }
    return v;
}

static size_t per_token_ws(const gemm::Geometry& g) { return (size_t)g.F * (g.Kp + 2 * g.p) * 2 * (g.split ? 2 : 1); }   // split: (hi, lo) pairs

// workspace = [256-byte status area | tile counters of the fused launch | Xf | Yf]
static size_t ws_fixed(int64_t N) { return kWsHeader + gemm::fused_counter_bytes(N) + 512; }

int64_t gemm_chunk_tokens(int64_t N, int p, int q, int B, size_t workspace_bytes, int split) {
    const gemm::Geometry g = gemm::geometry(p, q, B, split);
    if (workspace_bytes <= ws_fixed(N)) return 0;
    int64_t nc = (int64_t)((workspace_bytes - ws_fixed(N)) / per_token_ws(g));
    if (nc >= gemm::token_pitch(N)) return N;
    return nc / gemm::TILE_M * gemm::TILE_M;
}

size_t gemm_workspace_bytes(int64_t N, int p, int q, int B, int split) {
    const gemm::Geometry g = gemm::geometry(p, q, B, split);
    int64_t nc = (int64_t)(chunk_budget() / per_token_ws(g)) / gemm::TILE_M * gemm::TILE_M;
    if (nc < gemm::TILE_M) nc = gemm::TILE_M;
    if (nc > N) nc = gemm::token_pitch(N);
    return ws_fixed(N) + (size_t)nc * per_token_ws(g);
}

cudaError_t launch_spectrum_gemm(const float* c, void* spec, int p, int q, int B, int split, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(p, q, B, split);
    constexpr int WARPS = 8;
    unsigned char* base = reinterpret_cast<unsigned char*>(spec);
    gemm::dft_matrix_kernel<<<B, B, 0, st>>>(B, reinterpret_cast<__half*>(base + g.wf16_off), reinterpret_cast<__nv_bfloat16*>(base + g.wbf16_off),
                                            reinterpret_cast<__nv_bfloat16*>(base + g.winv_off),
                                            split ? reinterpret_cast<__half*>(base + g.wf16_lo_off) : nullptr,
                                            split ? reinterpret_cast<__nv_bfloat16*>(base + g.wbf16_lo_off) : nullptr,
                                            split ? reinterpret_cast<__nv_bfloat16*>(base + g.winv_lo_off) : nullptr);
    cudaError_t e0 = cudaGetLastError();
    if (e0 != cudaSuccess) return e0;
#define BCL_SG(BB)                                                                                                         \
    {                                                                                                                      \
        auto kern = gemm::spectrum_gemm_kernel<BB, WARPS>;                                                                  \
        const size_t smem = gemm::RowsSmem<BB, 1, WARPS>::bytes(q);                                                         \
        cudaError_t e = gemm::set_smem(kern, smem);                                                                         \
        if (e != cudaSuccess) return e;                                                                                    \
        kern<<<(unsigned)(g.Np / 2), WARPS * 32, smem, st>>>(c, p, q, reinterpret_cast<__nv_bfloat16*>(spec),                \
                                                              split ? reinterpret_cast<__nv_bfloat16*>(base + g.cmat_lo_off) : nullptr, g.Np, g.Kp); \
        return cudaGetLastError();                                                                                         \
    }
    switch (B) { case 64: BCL_SG(64) case 128: BCL_SG(128) default: BCL_SG(256) }
#undef BCL_SG
}

cudaError_t launch_gemm_rfft(const ProblemDev& P, int B, int64_t n_begin, int nc, void* Xf, void* Xf_lo, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B);
    const int nt = gemm::rows_nt(P.q, B);
#define BCL_R1(BB) (nt == 2 ? gemm::launch_rfft_rows_t<BB, 2>(P, n_begin, nc, Xf, Xf_lo, g.KG, st) : gemm::launch_rfft_rows_t<BB, 1>(P, n_begin, nc, Xf, Xf_lo, g.KG, st))
    switch (B) { case 64: return BCL_R1(64); case 128: return BCL_R1(128); default: return BCL_R1(256); }
#undef BCL_R1
}

cudaError_t launch_gemm_irfft(const ProblemDev& P, int B, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo, cudaStream_t st) {
    const int nt = gemm::rows_nt(P.p, B);
#define BCL_R3(BB) (nt == 2 ? gemm::launch_irfft_rows_t<BB, 2>(P, n_begin, nc, Yf, Yf_lo, st) : gemm::launch_irfft_rows_t<BB, 1>(P, n_begin, nc, Yf, Yf_lo, st))
    switch (B) { case 64: return BCL_R3(64); case 128: return BCL_R3(128); default: return BCL_R3(256); }
#undef BCL_R3
}

// status: device int, set non-zero if a pipeline wait timed out.
// Xf_lo / Yf_lo != NULL: split layout (Cmat then carries Cmat_lo at geometry(.., split = 1).cmat_lo_off)
cudaError_t launch_gemm_mma(const void* Xf, const void* Xf_lo, const void* Cmat, void* Yf, void* Yf_lo, int* status, int p, int q, int B, int nc, int max_ctas,
                            cudaStream_t st) {
    const int split = (Xf_lo != nullptr && Yf_lo != nullptr) ? 1 : 0;
    const gemm::Geometry g = gemm::geometry(p, q, B, split);
    gemm::StageLaunch<gemm::GemmArgs> L;
    if (!gemm::prep_gemm(Xf, Cmat, Yf, status, g, nc, L)) return cudaErrorNotSupported;
    const int grid = L.a.num_items < max_ctas ? L.a.num_items : max_ctas;
    if (grid < 1) return cudaSuccess;
    if (split) {
        gemm::GemmLoMaps lo;
        if (!gemm::prep_gemm_lo(Xf_lo, Cmat, Yf_lo, g, nc, &lo)) return cudaErrorNotSupported;
        cudaError_t e = gemm::ensure_dynamic_smem(reinterpret_cast<const void*>(gemm::freq_gemm_split_kernel), L.smem);
        if (e != cudaSuccess) return e;
        return gemm::launch_pdl(gemm::freq_gemm_split_kernel, (unsigned)grid, gemm::GEMM_THREADS, L.smem, st, L.tm_in, L.tm_w, L.tm_out, lo, L.a);
    }
    cudaError_t e = gemm::ensure_dynamic_smem(reinterpret_cast<const void*>(gemm::freq_gemm_kernel), L.smem);
    if (e != cudaSuccess) return e;
    return gemm::launch_pdl(gemm::freq_gemm_kernel, (unsigned)grid, gemm::GEMM_THREADS, L.smem, st, L.tm_in, L.tm_w, L.tm_out, L.a);
}

static int sm_count() {
    static int cached[64] = {0};                                       // per device; a race only repeats the query
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return n;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

// impl: 0 = pick (tensor-core DFT when supported), 1 = CUDA-core FFT kernels, 2 = tensor-core DFT kernels (error if unsupported)
cudaError_t launch_prefill_stage1(const ProblemDev& P, int B, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, int* status, int impl, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B);
    const bool tc_ok = gemm::dft_tc_supported(B, P.x_dtype) && spec_gemm != nullptr;
    if (impl == 2 && !tc_ok) return cudaErrorNotSupported;
    if (impl != 1 && tc_ok) return gemm::launch_dft_fwd_tc(P, g, spec_gemm, n_begin, nc, Xf, nullptr, status, sm_count(), st);
    return launch_gemm_rfft(P, B, n_begin, nc, Xf, nullptr, st);
}

cudaError_t launch_prefill_stage3(const ProblemDev& P, int B, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, int impl, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B);
    const bool tc_ok = gemm::dft_tc_supported(B, P.x_dtype) && P.y_dtype == P.x_dtype && spec_gemm != nullptr;   // (peer destinations: one TMA store per destination)
    if (impl == 2 && !tc_ok) return cudaErrorNotSupported;
    if (impl != 1 && tc_ok) return gemm::launch_dft_inv_tc(P, g, spec_gemm, n_begin, nc, Yf, nullptr, status, sm_count(), st);
    return launch_gemm_irfft(P, B, n_begin, nc, Yf, nullptr, st);
}

// BCL_PREFILL_CHAIN=1: the single-launch chain kernel (three stages, grid barriers in between) instead of three PDL-chained
// launches per chunk.  Opt-in: bit-identical, but measured 1-4 us SLOWER (4096x4096: 26.9 / 34.1 / 142.1 us against 26.3 / 30.0 /
// 140.3 us at 512 / 2048 / 16384 tokens): programmatic dependent launch already hides the launch transitions, and a grid barrier
// waits for the slowest CTA of every stage.  It only saves host launch cost (one launch instead of three).
static bool chain_enabled() {
    static const bool on = getenv("BCL_PREFILL_CHAIN") && getenv("BCL_PREFILL_CHAIN")[0] == '1';
    return on;
}

// split = 1: fp32-accurate chain (fp32 spectra): CUDA-core fp32 rfft -> (hi, lo) Xf -> split tcgen05 GEMM -> (hi, lo) Yf -> CUDA-core irfft
cudaError_t launch_gemm_forward(const ProblemDev& P, int B, const void* spec_gemm, void* workspace, size_t workspace_bytes, int split, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B, split);
    const int64_t ncmax = gemm_chunk_tokens(P.N, P.p, P.q, B, workspace_bytes, split);
    if (ncmax < 1) return cudaErrorInvalidValue;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    int* status = reinterpret_cast<int*>(ws);
    const size_t ctr_bytes = gemm::fused_counter_bytes(P.N);
    const size_t xfb = (gemm::xf_bytes(g, ncmax) + 255) & ~(size_t)255, yfb = (gemm::yf_bytes(g, ncmax) + 255) & ~(size_t)255;
    unsigned char* Xf = ws + kWsHeader + ctr_bytes;
    unsigned char* Yf = Xf + xfb * (split ? 2 : 1);
    if (split) {
        unsigned char* Xf_lo = Xf + xfb;
        unsigned char* Yf_lo = Yf + yfb;
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
        const int ctas = sm_count();
        static const int impl_s = getenv("BCL_PREFILL_DFT") ? atoi(getenv("BCL_PREFILL_DFT")) : 0;   // 1: force the CUDA-core FFT stages
        // 16-bit activations, B <= 128: the DFT stages run on the tensor cores too, with (hi, lo) DFT matrices
        const bool tc1 = impl_s != 1 && gemm::dft_tc_supported(B, P.x_dtype);               // B = 256: quarter-N launches (the (hi, lo) DFT matrices fit 64 rows at a time)
        const bool tc3 = tc1 && P.y_dtype == P.x_dtype;
        for (int64_t n0 = 0; n0 < P.N; n0 += ncmax) {
            const int nc = (int)((P.N - n0 < ncmax) ? (P.N - n0) : ncmax);
            if (tc1 && tc3 && B <= 128 && chain_enabled() && ctas >= 8 && P.n_extra == 0) {    // all three stages on the tensor cores: one launch
                if ((e = cudaMemsetAsync(ws + kWsHeader, 0, 256, st)) != cudaSuccess) return e;
                if ((e = gemm::launch_prefill_chain(P, g, spec_gemm, n0, nc, Xf, Xf_lo, Yf, Yf_lo, status, reinterpret_cast<int*>(ws + kWsHeader), ctas, st)) != cudaSuccess) return e;
                continue;
            }
            e = tc1 ? gemm::launch_dft_fwd_tc(P, g, spec_gemm, n0, nc, Xf, Xf_lo, status, ctas, st) : launch_gemm_rfft(P, B, n0, nc, Xf, Xf_lo, st);
            if (e != cudaSuccess) return e;
            if ((e = launch_gemm_mma(Xf, Xf_lo, spec_gemm, Yf, Yf_lo, status, P.p, P.q, B, nc, ctas, st)) != cudaSuccess) return e;
            e = tc3 ? gemm::launch_dft_inv_tc(P, g, spec_gemm, n0, nc, Yf, Yf_lo, status, ctas, st) : launch_gemm_irfft(P, B, n0, nc, Yf, Yf_lo, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    static const int impl = getenv("BCL_PREFILL_DFT") ? atoi(getenv("BCL_PREFILL_DFT")) : 0;       // A/B switch for profiling only
    // Single-launch pipeline (bcl_prefill_fused.cu): opt-in.  Measured 155-165 us against 140 us for the three launches
    // (4096->4096, 16384 tokens): every stage's epilogue already runs near the TMEM read rate, so splitting the SMs three ways
    // costs what the saved HBM passes gain (profiles/r1_prefill_notes.md).  BCL_FUSED_MIN_TILES=<tiles> enables it.
    static const int fused_min_tiles = getenv("BCL_FUSED_MIN_TILES") ? atoi(getenv("BCL_FUSED_MIN_TILES")) : 0;    // <= 0: never
    const int ctas = sm_count();
    const bool tc_ok = impl != 1 && spec_gemm != nullptr && gemm::dft_tc_supported(B, P.x_dtype) && P.y_dtype == P.x_dtype;
    if (tc_ok && P.n_extra == 0 && B <= 128 && ncmax >= P.N && fused_min_tiles > 0 && (P.N + gemm::TILE_M - 1) / gemm::TILE_M >= fused_min_tiles && ctas >= 12) {
        // one launch: the three stages pipelined tile by tile over L2-resident Xf / Yf (bcl_prefill_fused.cu)
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);                       // (status[1] is sticky: never cleared here)
        if (e == cudaSuccess) e = cudaMemsetAsync(ws + kWsHeader, 0, ctr_bytes, st);
        if (e != cudaSuccess) return e;
        return gemm::launch_prefill_fused(P, g, spec_gemm, Xf, Yf, status, reinterpret_cast<int*>(ws + kWsHeader), ctas, st);
    }
    // Wave schedule (bcl_prefill_wave.cu): one launch, all three stages interleaved on every SM over L2-resident rings of Xf / Yf.
    // BCL_PREFILL_WAVE=0 disables it, BCL_PREFILL_WAVE_MIN_TILES moves the switch-over (token tiles of 128).
    // (read per call: tests and tools/wave_bench.py switch it inside one process)
    const char* wave_env = getenv("BCL_PREFILL_WAVE");
    const int wave_on = wave_env ? atoi(wave_env) : 0;
    const char* wave_min_env = getenv("BCL_PREFILL_WAVE_MIN_TILES");
    const int wave_min_tiles = wave_min_env ? atoi(wave_min_env) : 48;
    if (wave_on && tc_ok && impl == 0 && ctas >= 8 && gemm::wave_supported(P, g) && (P.N + gemm::TILE_M - 1) / gemm::TILE_M >= wave_min_tiles &&
        workspace_bytes > kWsHeader + ctr_bytes) {
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ws + kWsHeader, 0, ctr_bytes, st);
        if (e != cudaSuccess) return e;
        e = gemm::launch_prefill_wave(P, g, spec_gemm, Xf, workspace_bytes - kWsHeader - ctr_bytes, status, reinterpret_cast<int*>(ws + kWsHeader), ctas, st);
        if (e != cudaErrorInvalidValue) return e;              // (workspace too small for the rings: the three launches below)
    }
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    for (int64_t n0 = 0; n0 < P.N; n0 += ncmax) {
        const int nc = (int)((P.N - n0 < ncmax) ? (P.N - n0) : ncmax);
        if (tc_ok && impl == 0 && chain_enabled() && ctas >= 8 && P.n_extra == 0) {            // all three stages on the tensor cores: one launch
            if ((e = cudaMemsetAsync(ws + kWsHeader, 0, 256, st)) != cudaSuccess) return e;
            if ((e = gemm::launch_prefill_chain(P, g, spec_gemm, n0, nc, Xf, nullptr, Yf, nullptr, status, reinterpret_cast<int*>(ws + kWsHeader), ctas, st)) != cudaSuccess) return e;
            continue;
        }
        if ((e = launch_prefill_stage1(P, B, spec_gemm, n0, nc, Xf, status, impl, st)) != cudaSuccess) return e;
        if ((e = launch_gemm_mma(Xf, nullptr, spec_gemm, Yf, nullptr, status, P.p, P.q, B, nc, ctas, st)) != cudaSuccess) return e;
        if ((e = launch_prefill_stage3(P, B, spec_gemm, n0, nc, Yf, status, impl, st)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}


// gate | up stacked along the output-block axis (P.p = 2 x the MLP's intermediate block count), GEMM-layout spectrum (bf16, or the
// (hi, lo) pairs of fp32 spectra with split = 1): the forward-DFT and frequency-GEMM stages of the chain, then the SWIGLU inverse
// stage: P.y = h = round(silu(round(gate))) * round(up).
// Chunked like launch_gemm_forward.  cudaErrorNotSupported: the caller runs the three modules one by one.
cudaError_t launch_gemm_forward_swiglu(const ProblemDev& P, int B, const void* spec_gemm, void* workspace, size_t workspace_bytes, int split, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B, split);
    if (B > 128 || (P.p & 1) || !gemm::dft_tc_supported(B, P.x_dtype) || P.y_dtype != P.x_dtype || P.bias != nullptr || P.n_extra != 0 || spec_gemm == nullptr)
        return cudaErrorNotSupported;
    const int64_t ncmax = gemm_chunk_tokens(P.N, P.p, P.q, B, workspace_bytes, split);
    if (ncmax < 1) return cudaErrorInvalidValue;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    int* status = reinterpret_cast<int*>(ws);
    const size_t ctr_bytes = gemm::fused_counter_bytes(P.N);
    const size_t xfb = (gemm::xf_bytes(g, ncmax) + 255) & ~(size_t)255, yfb = (gemm::yf_bytes(g, ncmax) + 255) & ~(size_t)255;
    unsigned char* Xf = ws + kWsHeader + ctr_bytes;                 // same layout as launch_gemm_forward: Xf [| Xf_lo] | Yf [| Yf_lo]
    unsigned char* Yf = Xf + xfb * (split ? 2 : 1);
    unsigned char* Xf_lo = split ? Xf + xfb : nullptr;
    unsigned char* Yf_lo = split ? Yf + yfb : nullptr;
    const int ctas = sm_count();
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    for (int64_t n0 = 0; n0 < P.N; n0 += ncmax) {
        const int nc = (int)((P.N - n0 < ncmax) ? (P.N - n0) : ncmax);
        if ((e = gemm::launch_dft_fwd_tc(P, g, spec_gemm, n0, nc, Xf, Xf_lo, status, ctas, st)) != cudaSuccess) return e;
        if ((e = launch_gemm_mma(Xf, Xf_lo, spec_gemm, Yf, Yf_lo, status, P.p, P.q, B, nc, ctas, st)) != cudaSuccess) return e;
        if ((e = gemm::launch_dft_inv_swiglu_tc(P, g, spec_gemm, n0, nc, Yf, Yf_lo, status, ctas, st)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace bcl
