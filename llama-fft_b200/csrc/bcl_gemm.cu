// bcl_gemm.cu -- the prefill (large-N) BC-linear path: per-frequency complex GEMM on the tcgen05 tensor cores.
//
// reference: src/patch_llama_fft.py:252-263.  For many tokens the middle step of the forward,
//     Yf[n,o,f] = sum_i Cf[o,i,f] * Xf[n,i,f]            (torch.einsum("oif,nif->nof"), :260)
// is, per frequency f, a dense complex (tokens x q) * (q x p) product.  Embedded in the reals
//     [Yr Yi] = [Xr Xi] * [[Cr, Ci], [-Ci, Cr]]
// it is one bf16 GEMM per frequency with K = 2q, N = 2p and fp32 accumulation -- tensor-core work.  The packed (DC,
// Nyquist) slot of the half-spectrum is the same GEMM with the block-diagonal matrix [[C_dc, 0], [0, C_ny]].
//
// Three kernels per token chunk, staged through an L2-resident workspace:
//   rfft_rows_kernel   x (N,in)  -> Xf  [F][N][Kp]  bf16   CUDA cores, register FFT core of bcl_fft.cuh    (:249-252)
//   freq_gemm_kernel   Xf, Cmat  -> Yf  [F][N][Np]  bf16   tcgen05.mma, TMA-fed smem ring, TMEM accumulators (:260)
//   irfft_rows_kernel  Yf        -> y (N,out)              CUDA cores + the reference epilogue             (:263-271)
// with F = B/2 packed frequency slots, Kp = 2q and Np = 2p rounded up to 16, k = 2i+{re,im}, n = 2j+{re,im}.
// GEMM-layout spectrum:  Cmat[f][n][k]  (K-major rows, zero padded), built once per parameter version by
// spectrum_gemm_kernel (replaces rfft(c), :256-257).
#include <cuda.h>
#include <mutex>
#include "bcl_stream_parts.cuh"
#include "bcl_sm100.cuh"

namespace bcl {
namespace gemm {

using namespace sm100;

constexpr int TILE_M = 128;          // tokens per MMA tile (UMMA M)
constexpr int KATOM = 64;            // bf16 elements per 128-byte swizzle row
constexpr int ACC_COLS = 256;        // TMEM columns reserved per accumulator (two accumulators: 512 columns)
constexpr int EPI_PITCH = 144;       // bytes per staged output row (128 data + 16 pad: conflict-free 16-byte accesses)
constexpr int GEMM_THREADS = 192;    // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int MAX_STAGES = 8;

struct Geometry {
    int F, Kp, Np, Nn, n_chunks, k_atoms;
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

__host__ inline Geometry geometry(int p, int q, int B) {
    Geometry g;
    g.F = B / 2;
    g.Kp = round_up(2 * q, 16);
    g.Np = round_up(2 * p, 16);
    g.n_chunks = (g.Np + 255) / 256;
    g.Nn = round_up((g.Np + g.n_chunks - 1) / g.n_chunks, 16);
    g.k_atoms = (g.Kp + KATOM - 1) / KATOM;
    return g;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

// ---------------------------------------------------------------------------------------------------------
// stage 1: rfft of NT tokens (all q input blocks) -> rows of the per-frequency A operand
// ---------------------------------------------------------------------------------------------------------
template <int B, int NT, int WARPS> struct RowsSmem {
    static constexpr int PITCH = B + 2;                                // floats per block: (2i + 2f) mod 32 banks -> conflict-free
    static constexpr size_t tw_bytes = (size_t)(B / 2) * 8;
    static constexpr size_t fscr_bytes = (size_t)WARPS * FftCfg<B>::WARP_SCRATCH_BYTES;
    static size_t bytes(int blocks) { return tw_bytes + fscr_bytes + (size_t)NT * blocks * PITCH * 4; }
};

template <int B, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) rfft_rows_kernel(const __grid_constant__ ProblemDev P, int64_t n_begin, int nc,
                                                                __nv_bfloat16* __restrict__ Xf, int Kp) {
    using SM = RowsSmem<B, NT, WARPS>;
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float2* fscr_all = reinterpret_cast<float2*>(smem_raw + SM::tw_bytes);
    float* xs = reinterpret_cast<float*>(smem_raw + SM::tw_bytes + SM::fscr_bytes);
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
    for (int row = warp; row < F * NT; row += WARPS) {                  // one warp writes one (f, token) row: Kp bf16, contiguous
        const int n = row % NT, f = row / NT;
        if (t0 + n >= nc) continue;
        uint32_t* dst = reinterpret_cast<uint32_t*>(Xf + ((size_t)f * nc + (t0 + n)) * Kp);
        for (int i = lane; i < Kp / 2; i += 32) {
            uint32_t w = 0u;
            if (i < q) {
                const float2 v = *reinterpret_cast<const float2*>(xs + ((size_t)n * q + i) * SM::PITCH + 2 * f);
                w = pack_bf16(v.x, v.y);                                // f = 0: (DC, Nyquist), both real
            }
            dst[i] = w;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// GEMM-layout spectrum: c (p,q,B) fp32 -> Cmat[F][Np][Kp] bf16.  One CTA per output block j (rows 2j, 2j+1 of every f).
// ---------------------------------------------------------------------------------------------------------
template <int B, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spectrum_gemm_kernel(const float* __restrict__ c, int p, int q, __nv_bfloat16* __restrict__ Cmat, int Np, int Kp) {
    using SM = RowsSmem<B, 1, WARPS>;
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float2* fscr_all = reinterpret_cast<float2*>(smem_raw + SM::tw_bytes);
    float* xs = reinterpret_cast<float*>(smem_raw + SM::tw_bytes + SM::fscr_bytes);
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
        for (int i = lane; i < Kp / 2; i += 32) {
            uint32_t wr = 0u, wi = 0u;
            if (j < p && i < q) {
                const float2 v = *reinterpret_cast<const float2*>(xs + (size_t)i * SM::PITCH + 2 * f);
                if (f == 0) { wr = pack_bf16(v.x, 0.f); wi = pack_bf16(0.f, v.y); }                  // [[C_dc, 0], [0, C_ny]]
                else { wr = pack_bf16(v.x, -v.y); wi = pack_bf16(v.y, v.x); }                        // Yr = Xr Cr - Xi Ci ; Yi = Xr Ci + Xi Cr
            }
            row_re[i] = wr;
            row_im[i] = wi;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// stage 2: batched per-frequency GEMM on tcgen05.   D[128 tokens x Nn] = A[128 x Kp] * B[Nn x Kp]^T
// ---------------------------------------------------------------------------------------------------------
struct GemmArgs {
    __nv_bfloat16* Yf;
    int* status;
    int ntok, F, Kp, Np, Nn, n_chunks, k_atoms, m_tiles, num_tiles, stages;
};

struct GemmBarriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void decode_tile(const GemmArgs& a, int tile, int& f, int& m, int& nc) {
    nc = tile % a.n_chunks;
    const int r = tile / a.n_chunks;
    m = r % a.m_tiles;
    f = r / a.m_tiles;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) freq_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                     const __grid_constant__ GemmArgs a) {
    extern __shared__ __align__(1024) unsigned char gemm_smem[];
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(gemm_smem);
    unsigned char* epi = gemm_smem + 512;                                         // 4 warps x 32 rows x EPI_PITCH
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem) + 512 + 4 * 32 * EPI_PITCH + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TILE_M * KATOM * 2, b_bytes = (uint32_t)a.Nn * KATOM * 2;
    const uint32_t stage_bytes = a_bytes + b_bytes;                              // both multiples of 1024
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 2 * ACC_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
                int f, m, nc; decode_tile(a, tile, f, m, nc);
                for (int ka = 0; ka < a.k_atoms; ++ka) {
                    if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { atomicExch(a.status, 1); goto done; }
                    unsigned char* sa = ring + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&bars->full[stage], stage_bytes);
                    tma_load_3d(sa, &tmA, ka * KATOM, m * TILE_M, f, &bars->full[stage]);
                    tma_load_3d(sa + a_bytes, &tmB, ka * KATOM, nc * a.Nn, f, &bars->full[stage]);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, a.Nn);
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1; const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
                if (!mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1u)) { atomicExch(a.status, 2); goto done; }   // epilogue drained this accumulator
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
                for (int ka = 0; ka < a.k_atoms; ++ka) {
                    if (!mbar_wait(&bars->full[stage], ph)) { atomicExch(a.status, 3); goto done; }
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(ring + (size_t)stage * stage_bytes);
                    const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + a_bytes);
                    const int ks = min(KATOM / 16, (a.Kp - ka * KATOM) / 16);
                    for (int kk = 0; kk < ks; ++kk)
                        umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (ka | kk) != 0 ? 1u : 0u);
                    umma_commit(&bars->empty[stage]);                            // smem stage reusable once these MMAs retire
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                umma_commit(&bars->tmem_full[acc]);                              // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps: TMEM -> registers -> bf16 -> smem -> coalesced global rows =====================
        const int quarter = warp & 3;                                            // TMEM lanes [32*quarter, +32) are this warp's
        unsigned char* stg = epi + (size_t)(warp - 2) * 32 * EPI_PITCH;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
            int f, m, nc; decode_tile(a, tile, f, m, nc);
            const int acc = it & 1; const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[acc], acc_ph))) { atomicExch(a.status, 4); goto done; }   // warp-uniform outcome
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * ACC_COLS;
            const int tok0 = m * TILE_M + quarter * 32;
            for (int c0 = 0; c0 < a.Nn; c0 += 64) {
                const int width = min(64, a.Nn - c0);                            // 16, 32, 48 or 64 columns (warp-uniform)
                for (int h = 0; h < width; h += 32) {
                    uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)lane * EPI_PITCH + h * 2);
                    if (width - h >= 32) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(c0 + h), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            dst[v] = make_uint4(pack_bf16(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1])), pack_bf16(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3])),
                                                pack_bf16(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5])), pack_bf16(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7])));
                    } else {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(taddr + (uint32_t)(c0 + h), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 2; ++v)
                            dst[v] = make_uint4(pack_bf16(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1])), pack_bf16(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3])),
                                                pack_bf16(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5])), pack_bf16(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7])));
                    }
                }
                __syncwarp();
                const int nch = width / 8;                                       // 16-byte chunks per row
                for (int id = lane; id < 32 * nch; id += 32) {
                    const int r = id / nch, cc = id - r * nch;
                    const int tok = tok0 + r, col = nc * a.Nn + c0 + cc * 8;
                    if (tok < a.ntok && col < a.Np)
                        *reinterpret_cast<uint4*>(a.Yf + ((size_t)f * a.ntok + tok) * a.Np + col) = *reinterpret_cast<const uint4*>(stg + (size_t)r * EPI_PITCH + cc * 16);
                }
                __syncwarp();
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        }
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * ACC_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// stage 3: gather a token's output spectra from the per-frequency rows, inverse rfft, reference epilogue
// ---------------------------------------------------------------------------------------------------------
template <int B, int NT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) irfft_rows_kernel(const __grid_constant__ ProblemDev P, int64_t n_begin, int nc,
                                                                 const __nv_bfloat16* __restrict__ Yf, int Np) {
    using SM = RowsSmem<B, NT, WARPS>;
    using C = FftCfg<B>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);
    float2* fscr_all = reinterpret_cast<float2*>(smem_raw + SM::tw_bytes);
    float* ys = reinterpret_cast<float*>(smem_raw + SM::tw_bytes + SM::fscr_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    const int p = P.p;
    const int t0 = blockIdx.x * NT;
    constexpr int F = B / 2;
    for (int row = warp; row < F * NT; row += WARPS) {                  // one warp reads one (f, token) row: Np bf16, contiguous
        const int n = row % NT, f = row / NT;
        if (t0 + n >= nc) continue;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Yf + ((size_t)f * nc + (t0 + n)) * Np);
        for (int j = lane; j < p; j += 32)
            *reinterpret_cast<float2*>(ys + ((size_t)n * p + j) * SM::PITCH + 2 * f) = unpack_bf16(__ldg(src + j));
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

// bf16 tensor [d2][d1][d0] (d0 contiguous), box 64 x box_rows x 1, 128-byte swizzle, zero fill out of bounds
static bool make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {KATOM, box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t gemm_smem_bytes(int Nn, int stages) { return 512 + 4 * 32 * EPI_PITCH + 1024 + (size_t)stages * (TILE_M * KATOM * 2 + (size_t)Nn * KATOM * 2); }

static int pick_gemm_stages(int Nn) {
    int s = MAX_STAGES;
    while (s > 2 && gemm_smem_bytes(Nn, s) > 220 * 1024) --s;
    return s;
}

template <typename K> static cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int B, int NT>
static cudaError_t launch_rfft_rows_t(const ProblemDev& P, int64_t n_begin, int nc, void* Xf, int Kp, cudaStream_t st) {
    constexpr int WARPS = 8;
    auto kern = rfft_rows_kernel<B, NT, WARPS>;
    const size_t smem = RowsSmem<B, NT, WARPS>::bytes(P.q);
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((nc + NT - 1) / NT), WARPS * 32, smem, st>>>(P, n_begin, nc, reinterpret_cast<__nv_bfloat16*>(Xf), Kp);
    return cudaGetLastError();
}

template <int B, int NT>
static cudaError_t launch_irfft_rows_t(const ProblemDev& P, int64_t n_begin, int nc, const void* Yf, int Np, cudaStream_t st) {
    constexpr int WARPS = 8;
    auto kern = irfft_rows_kernel<B, NT, WARPS>;
    const size_t smem = RowsSmem<B, NT, WARPS>::bytes(P.p);
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((nc + NT - 1) / NT), WARPS * 32, smem, st>>>(P, n_begin, nc, reinterpret_cast<const __nv_bfloat16*>(Yf), Np);
    return cudaGetLastError();
}

// tokens per CTA of the row kernels: two while the spectrum tile stays small enough for >= 3 CTAs per SM
static int rows_nt(int blocks, int B) { return (size_t)2 * blocks * (B + 2) * 4 <= 48 * 1024 ? 2 : 1; }

}  // namespace gemm

size_t gemm_spectrum_bytes(int p, int q, int B) {
    const gemm::Geometry g = gemm::geometry(p, q, B);
    return (size_t)g.F * g.Np * g.Kp * 2;
}

// workspace = [256-byte status area | Xf chunk | Yf chunk]; chunk tokens chosen so both stay L2-resident
static constexpr size_t kWsHeader = 256;
static constexpr size_t kChunkBudget = 48u << 20;

static size_t per_token_ws(const gemm::Geometry& g) { return (size_t)g.F * (g.Kp + g.Np) * 2; }

int64_t gemm_chunk_tokens(int64_t N, int p, int q, int B, size_t workspace_bytes) {
    const gemm::Geometry g = gemm::geometry(p, q, B);
    if (workspace_bytes <= kWsHeader + 512) return 0;
    int64_t nc = (int64_t)((workspace_bytes - kWsHeader - 512) / per_token_ws(g));
    if (nc >= N) return N;
    nc = nc / gemm::TILE_M * gemm::TILE_M;
    return nc;
}

size_t gemm_workspace_bytes(int64_t N, int p, int q, int B) {
    const gemm::Geometry g = gemm::geometry(p, q, B);
    int64_t nc = (int64_t)(kChunkBudget / per_token_ws(g)) / gemm::TILE_M * gemm::TILE_M;
    if (nc < gemm::TILE_M) nc = gemm::TILE_M;
    if (nc > N) nc = N;
    return kWsHeader + 512 + (size_t)nc * per_token_ws(g);
}

cudaError_t launch_spectrum_gemm(const float* c, void* spec, int p, int q, int B, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(p, q, B);
    constexpr int WARPS = 8;
#define BCL_SG(BB)                                                                                                         \
    {                                                                                                                      \
        auto kern = gemm::spectrum_gemm_kernel<BB, WARPS>;                                                                  \
        const size_t smem = gemm::RowsSmem<BB, 1, WARPS>::bytes(q);                                                         \
        cudaError_t e = gemm::set_smem(kern, smem);                                                                         \
        if (e != cudaSuccess) return e;                                                                                    \
        kern<<<(unsigned)(g.Np / 2), WARPS * 32, smem, st>>>(c, p, q, reinterpret_cast<__nv_bfloat16*>(spec), g.Np, g.Kp); \
        return cudaGetLastError();                                                                                         \
    }
    switch (B) { case 64: BCL_SG(64) case 128: BCL_SG(128) default: BCL_SG(256) }
#undef BCL_SG
}

cudaError_t launch_gemm_rfft(const ProblemDev& P, int B, int64_t n_begin, int nc, void* Xf, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B);
    const int nt = gemm::rows_nt(P.q, B);
#define BCL_R1(BB) (nt == 2 ? gemm::launch_rfft_rows_t<BB, 2>(P, n_begin, nc, Xf, g.Kp, st) : gemm::launch_rfft_rows_t<BB, 1>(P, n_begin, nc, Xf, g.Kp, st))
    switch (B) { case 64: return BCL_R1(64); case 128: return BCL_R1(128); default: return BCL_R1(256); }
#undef BCL_R1
}

cudaError_t launch_gemm_irfft(const ProblemDev& P, int B, int64_t n_begin, int nc, const void* Yf, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B);
    const int nt = gemm::rows_nt(P.p, B);
#define BCL_R3(BB) (nt == 2 ? gemm::launch_irfft_rows_t<BB, 2>(P, n_begin, nc, Yf, g.Np, st) : gemm::launch_irfft_rows_t<BB, 1>(P, n_begin, nc, Yf, g.Np, st))
    switch (B) { case 64: return BCL_R3(64); case 128: return BCL_R3(128); default: return BCL_R3(256); }
#undef BCL_R3
}

// Xf [F][nc][Kp], Cmat [F][Np][Kp] -> Yf [F][nc][Np].  status: device int, set non-zero if a pipeline wait timed out.
cudaError_t launch_gemm_mma(const void* Xf, const void* Cmat, void* Yf, int* status, int p, int q, int B, int nc, int max_ctas, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(p, q, B);
    CUtensorMap tmA, tmB;
    if (!gemm::make_map(&tmA, Xf, (uint64_t)g.Kp, (uint64_t)nc, (uint64_t)g.F, gemm::TILE_M)) return cudaErrorNotSupported;
    if (!gemm::make_map(&tmB, Cmat, (uint64_t)g.Kp, (uint64_t)g.Np, (uint64_t)g.F, (uint32_t)g.Nn)) return cudaErrorNotSupported;
    gemm::GemmArgs a{};
    a.Yf = reinterpret_cast<__nv_bfloat16*>(Yf);
    a.status = status;
    a.ntok = nc; a.F = g.F; a.Kp = g.Kp; a.Np = g.Np; a.Nn = g.Nn; a.n_chunks = g.n_chunks; a.k_atoms = g.k_atoms;
    a.m_tiles = (nc + gemm::TILE_M - 1) / gemm::TILE_M;
    a.num_tiles = a.m_tiles * g.F * g.n_chunks;
    a.stages = gemm::pick_gemm_stages(g.Nn);
    const size_t smem = gemm::gemm_smem_bytes(g.Nn, a.stages);
    cudaError_t e = cudaFuncSetAttribute(gemm::freq_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = a.num_tiles < max_ctas ? a.num_tiles : max_ctas;
    if (grid < 1) return cudaSuccess;
    gemm::freq_gemm_kernel<<<(unsigned)grid, gemm::GEMM_THREADS, smem, st>>>(tmA, tmB, a);
    return cudaGetLastError();
}

static int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

cudaError_t launch_gemm_forward(const ProblemDev& P, int B, const void* Cmat, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const gemm::Geometry g = gemm::geometry(P.p, P.q, B);
    const int64_t ncmax = gemm_chunk_tokens(P.N, P.p, P.q, B, workspace_bytes);
    if (ncmax < 1) return cudaErrorInvalidValue;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    int* status = reinterpret_cast<int*>(ws);
    unsigned char* Xf = ws + kWsHeader;
    unsigned char* Yf = Xf + (((size_t)g.F * ncmax * g.Kp * 2 + 255) & ~(size_t)255);
    cudaError_t e = cudaMemsetAsync(status, 0, kWsHeader, st);
    if (e != cudaSuccess) return e;
    const int ctas = sm_count();
    for (int64_t n0 = 0; n0 < P.N; n0 += ncmax) {
        const int nc = (int)((P.N - n0 < ncmax) ? (P.N - n0) : ncmax);
        if ((e = launch_gemm_rfft(P, B, n0, nc, Xf, st)) != cudaSuccess) return e;
        if ((e = launch_gemm_mma(Xf, Cmat, Yf, status, P.p, P.q, B, nc, ctas, st)) != cudaSuccess) return e;
        if ((e = launch_gemm_irfft(P, B, n0, nc, Yf, st)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace bcl
