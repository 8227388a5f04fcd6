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
#include "bcl_common.cuh"
#include "bcl_sm100.cuh"
#include "bcl_gemm.cuh"

namespace bcl {
namespace gemm {

using namespace sm100;

constexpr int DFT_EPI_WARPS = 8;     // two per TMEM lane quarter (one epilogue warp per SMSP is latency-bound)
constexpr int DFT_THREADS = 32 * (2 + DFT_EPI_WARPS);   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int DFT_MAX_STAGES = 6;
#ifndef BCL_EPI_NBUF
#define BCL_EPI_NBUF 2
#endif
constexpr int DFT_NBUF = BCL_EPI_NBUF;   // staging buffers per epilogue warp: TMA stores in flight
constexpr uint32_t DFT_EPI_BUF = 4096;   // fwd: [8 f][32 tokens][16 B];  inv: [32 tokens][128 B] swizzled

struct DftBarriers {
    uint64_t full[DFT_MAX_STAGES], empty[DFT_MAX_STAGES];
    uint64_t tmem_full[4], tmem_empty[4];
    uint64_t wfull;
    uint32_t tmem_base;
    uint32_t pad;
};

struct DftArgs {
    int* status;
    const void* bias;        // inverse only: (out) in the activation dtype, or NULL
    int64_t n_begin;         // first token of the chunk (row coordinate into x / y)
    int nc, m_tiles, groups; // groups: KG (forward) or p (inverse)
    int num_items, stages, act_bf16;
};

__device__ __forceinline__ uint32_t pack16(float a, float b, int bf16) {
    if (bf16) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// =========================================================================================================
// forward: x -> Xf
// =========================================================================================================
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ CUtensorMap tmXf, const __grid_constant__ DftArgs a) {
    constexpr int ATOMS = B / KATOM;                         // 128-byte K atoms per block (K = B time samples)
    constexpr uint32_t A_ATOM = TILE_M * 128;                // 128 tokens x 128 B
    constexpr uint32_t W_ATOM = B * 128;                     // B spectrum rows x 128 B
    constexpr uint32_t STAGE = ATOMS * A_ATOM;               // one input block of the tile
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    DftBarriers* bars = reinterpret_cast<DftBarriers*>(dft_smem);
    unsigned char* epi = dft_smem + 1024;                    // 4 warps x 2 x DFT_EPI_BUF
    unsigned char* wsm = dft_smem + 1024 + DFT_EPI_WARPS * DFT_NBUF * DFT_EPI_BUF;
    unsigned char* ring = wsm + ATOMS * W_ATOM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full[0], 1); mbar_init(&bars->tmem_empty[0], DFT_EPI_WARPS);
        mbar_init(&bars->wfull, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmXf); }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 4 * B >= 32 ? 4 * B : 32);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wfull, ATOMS * W_ATOM);                 // the DFT matrix: resident for the whole kernel
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + at * W_ATOM, &tmW, at * KATOM, 0, &bars->wfull);
            pdl_wait();                                                          // x may come from the previous kernel
            int stage = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
                const int kg = item % a.groups, m = item / a.groups;
                for (int ib = 0; ib < 4; ++ib) {
                    if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { atomicExch(a.status, 11); goto done; }
                    unsigned char* sa = ring + (size_t)stage * STAGE;
                    mbar_arrive_expect_tx(&bars->full[stage], STAGE);
                    for (int at = 0; at < ATOMS; ++at)                           // columns past in_features are zero-filled
                        tma_load_2d_hint(sa + at * A_ATOM, &tmX, (4 * kg + ib) * B + at * KATOM, (int)(a.n_begin + (int64_t)m * TILE_M), &bars->full[stage], L2_EVICT_FIRST);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_16(TILE_M, B, a.act_bf16);
            if (!mbar_wait(&bars->wfull, 0)) { atomicExch(a.status, 12); goto done; }
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
                if (!mbar_wait(&bars->tmem_empty[0], ((uint32_t)it & 1u) ^ 1u)) { atomicExch(a.status, 13); goto done; }
                tc_fence_after_sync();
                for (int ib = 0; ib < 4; ++ib) {
                    if (!mbar_wait(&bars->full[stage], ph)) { atomicExch(a.status, 14); goto done; }
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE), sw = smem_u32(wsm);
#pragma unroll
                    for (int at = 0; at < ATOMS; ++at) {
                        const uint64_t adesc = umma_desc_k_sw128(sa + at * A_ATOM), bdesc = umma_desc_k_sw128(sw + at * W_ATOM);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16_ss(tmem_base + (uint32_t)(ib * B), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (at | kk) != 0 ? 1u : 0u);
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                umma_commit(&bars->tmem_full[0]);
            }
        }
    } else {
        const int quarter = warp & 3;
        unsigned char* stg = epi + (size_t)(warp - 2) * DFT_NBUF * DFT_EPI_BUF;
        pdl_wait();                                                              // Xf may still be read by the previous kernel
        int it = 0, buf = 0;
        for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
            const int kg = item % a.groups, m = item / a.groups;
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[0], (uint32_t)it & 1u))) { atomicExch(a.status, 15); goto done; }
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            constexpr int FC = B / 16;                                           // chunks of 8 frequencies (16 accumulator columns)
            const int fc0 = ((warp - 2) >> 2) * (FC / 2), fc1 = fc0 + FC / 2;     // this warp's half of them
            for (int fc = fc0; fc < fc1; ++fc) {
                uint32_t r[4][16];
#pragma unroll
                for (int ib = 0; ib < 4; ++ib) tmem_ld_32x32b_x16(taddr + (uint32_t)(ib * B + fc * 16), r[ib]);
                tmem_ld_wait();
                if (fc == fc1 - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tmem_empty[0]);
                }
                if (lane == 0) tma_store_wait_read<DFT_NBUF - 1>();
                __syncwarp();
                uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)buf * DFT_EPI_BUF) + lane;
#pragma unroll
                for (int fl = 0; fl < 8; ++fl)                                   // chunk (f = 8 fc + fl, kg, token) = {Re,Im} x 4 blocks
                    dst[fl * 32] = make_uint4(pack_bf(__uint_as_float(r[0][2 * fl]), __uint_as_float(r[0][2 * fl + 1])),
                                              pack_bf(__uint_as_float(r[1][2 * fl]), __uint_as_float(r[1][2 * fl + 1])),
                                              pack_bf(__uint_as_float(r[2][2 * fl]), __uint_as_float(r[2][2 * fl + 1])),
                                              pack_bf(__uint_as_float(r[3][2 * fl]), __uint_as_float(r[3][2 * fl + 1])));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d_hint(&tmXf, 0, m * (TILE_M / TOK_GROUP) + quarter, kg, fc * 8, stg + (size_t)buf * DFT_EPI_BUF, L2_EVICT_LAST);
                    tma_store_commit();
                }
                if (++buf == DFT_NBUF) buf = 0;
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 4 * B >= 32 ? 4 * B : 32);
}

// One token's 64 outputs (fp32 accumulators, 1/B already folded into Winv) -> 8 swizzled 16-byte chunks of the store tile.
// Reference epilogue: round to x.dtype (:266), then + bias in that dtype (:270-271).
template <bool BF16, bool HAS_BIAS>
__device__ __forceinline__ void store_row(unsigned char* row, int lane, const uint32_t (&r)[64], const void* bias, int col0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v0 = __uint_as_float(r[8 * c + 2 * e]), v1 = __uint_as_float(r[8 * c + 2 * e + 1]);
            if constexpr (HAS_BIAS) {
                const uint32_t bw = __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(bias) + col0 + 8 * c + 2 * e));
                float b0, b1;
                if constexpr (BF16) { b0 = __uint_as_float(bw << 16); b1 = __uint_as_float(bw & 0xffff0000u); }
                else { const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&bw)); b0 = bf.x; b1 = bf.y; }
                v0 = round_to(v0, BF16 ? BCL_BF16 : BCL_F16) + b0;
                v1 = round_to(v1, BF16 ? BCL_BF16 : BCL_F16) + b1;
            }
            w[e] = pack16(v0, v1, BF16 ? 1 : 0);
        }
        *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// =========================================================================================================
// inverse: Yf -> y
// =========================================================================================================
template <int B>
__global__ void __launch_bounds__(DFT_THREADS, 1) dft_inv_tc_kernel(const __grid_constant__ CUtensorMap tmYf, const __grid_constant__ CUtensorMap tmW,
                                                                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ DftArgs a) {
    constexpr int ATOMS = B / KATOM;
    constexpr int FG = B / 8;
    constexpr uint32_t W_ATOM = B * 128;                     // B time rows x 128 B
    constexpr uint32_t STAGE = FG * 2048;                    // [fg][128 tokens][16 B]
    constexpr int NACC = 512 / B > 4 ? 4 : 512 / B;
    extern __shared__ __align__(1024) unsigned char dft_smem[];
    DftBarriers* bars = reinterpret_cast<DftBarriers*>(dft_smem);
    unsigned char* epi = dft_smem + 1024;
    unsigned char* wsm = dft_smem + 1024 + DFT_EPI_WARPS * DFT_NBUF * DFT_EPI_BUF;
    unsigned char* ring = wsm + ATOMS * W_ATOM;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], DFT_EPI_WARPS); }
        mbar_init(&bars->wfull, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmYf); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmOut); }
    if (warp == 1) tmem_alloc(&bars->tmem_base, NACC * B);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wfull, ATOMS * W_ATOM);
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + at * W_ATOM, &tmW, at * KATOM, 0, &bars->wfull);
            pdl_wait();                                                          // Yf comes from the previous kernel
            int stage = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < a.num_items; item += gridDim.x) {
                const int j = item % a.groups, m = item / a.groups;
                if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { atomicExch(a.status, 21); goto done; }
                mbar_arrive_expect_tx(&bars->full[stage], STAGE);
                tma_load_4d_hint(ring + (size_t)stage * STAGE, &tmYf, 0, m * (TILE_M / TOK_GROUP), j, 0, &bars->full[stage], L2_EVICT_FIRST);
                if (++stage == S) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, B);
            if (!mbar_wait(&bars->wfull, 0)) { atomicExch(a.status, 22); goto done; }
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
                const int acc = it % NACC; const uint32_t acc_ph = (uint32_t)(it / NACC) & 1u;
                if (!mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1u)) { atomicExch(a.status, 23); goto done; }
                if (!mbar_wait(&bars->full[stage], ph)) { atomicExch(a.status, 24); goto done; }
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE), sw = smem_u32(wsm);
#pragma unroll
                for (int kk = 0; kk < B / 16; ++kk)                              // K = B packed spectrum reals; one k-step = 2 fg slabs
                    umma_bf16_ss(tmem_base + (uint32_t)(acc * B), umma_desc_k_nosw(sa + (uint32_t)kk * 4096u, 128u, 2048u),
                                 umma_desc_k_sw128(sw + (kk / 4) * W_ATOM) + (uint64_t)(2 * (kk % 4)), idesc, kk != 0 ? 1u : 0u);
                umma_commit(&bars->empty[stage]);
                umma_commit(&bars->tmem_full[acc]);
                if (++stage == S) { stage = 0; ph ^= 1u; }
            }
        }
    } else {
        const int quarter = warp & 3;
        unsigned char* stg = epi + (size_t)(warp - 2) * DFT_NBUF * DFT_EPI_BUF;
        pdl_wait();                                                              // y may still be read by the previous kernel
        int it = 0, buf = 0;
        for (int item = blockIdx.x; item < a.num_items; item += gridDim.x, ++it) {
            const int j = item % a.groups, m = item / a.groups;
            const int acc = it % NACC; const uint32_t acc_ph = (uint32_t)(it / NACC) & 1u;
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[acc], acc_ph))) { atomicExch(a.status, 25); goto done; }
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * B);
            const int wh = (warp - 2) >> 2;                                      // second warp of the lane quarter?
            const int h0 = (ATOMS == 2) ? wh : 0, h1 = (ATOMS == 2) ? wh + 1 : (wh == 0 ? 1 : 0);
            if (h0 >= h1) {                                                      // B = 64: one half only, the partner warp just releases
                tc_fence_before_sync();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
            }
            for (int h = h0; h < h1; ++h) {                                      // 64 output columns = one 128-byte row of the store tile
                uint32_t r[64];
                tmem_ld_32x32b_x32(taddr + (uint32_t)(h * 64), *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                tmem_ld_32x32b_x32(taddr + (uint32_t)(h * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                tmem_ld_wait();
                if (h == h1 - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
                }
                if (lane == 0) tma_store_wait_read<DFT_NBUF - 1>();
                __syncwarp();
                unsigned char* row = stg + (size_t)buf * DFT_EPI_BUF + (size_t)lane * 128;
                const int col0 = j * B + h * 64;
                // 16-byte chunks of this token's row, 128-byte swizzle; the dtype / bias variants are separate straight-line code
                if (a.bias == nullptr) {
                    if (a.act_bf16) store_row<true, false>(row, lane, r, nullptr, 0);
                    else store_row<false, false>(row, lane, r, nullptr, 0);
                } else {
                    if (a.act_bf16) store_row<true, true>(row, lane, r, a.bias, col0);
                    else store_row<false, true>(row, lane, r, a.bias, col0);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOut, col0, (int)(a.n_begin + (int64_t)m * TILE_M + quarter * 32), stg + (size_t)buf * DFT_EPI_BUF);
                    tma_store_commit();
                }
                if (++buf == DFT_NBUF) buf = 0;
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, NACC * B);
}

// =========================================================================================================
// host side
// =========================================================================================================
bool dft_tc_supported(int B, int act_dtype) { return (B == 64 || B == 128) && (act_dtype == BCL_F16 || act_dtype == BCL_BF16); }

static size_t dft_smem_bytes(int B, uint32_t stage_bytes, int stages) {
    return 1024 + DFT_EPI_WARPS * DFT_NBUF * DFT_EPI_BUF + (size_t)(B / KATOM) * B * 128 + (size_t)stages * stage_bytes;
}
static int dft_pick_stages(int B, uint32_t stage_bytes) {
    int s = DFT_MAX_STAGES;
    while (s > 2 && dft_smem_bytes(B, stage_bytes, s) > 225 * 1024) --s;
    return s;
}

static bool map_w(CUtensorMap* m, const void* w, int B, CUtensorMapDataType dt) {                  // [B rows][B k], box 64 x B, 128-byte swizzle
    const uint64_t dims[2] = {(uint64_t)B, (uint64_t)B};
    const uint64_t strides[1] = {(uint64_t)B * 2};
    const uint32_t box[2] = {KATOM, (uint32_t)B};
    return make_tensor_map(m, dt, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

cudaError_t launch_dft_fwd_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, int* status, int max_ctas, cudaStream_t st) {
    const int B = g.B;
    const bool bf = P.x_dtype == BCL_BF16;
    const CUtensorMapDataType adt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmX, tmW, tmXf;
    {
        const uint64_t dims[2] = {(uint64_t)g.q * B, (uint64_t)P.N};
        const uint64_t strides[1] = {(uint64_t)P.x_stride * 2};
        const uint32_t box[2] = {KATOM, TILE_M};
        if (!make_tensor_map(&tmX, adt, 2, P.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
    }
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(spec_gemm);
    if (!map_w(&tmW, sb + (bf ? g.wbf16_off : g.wf16_off), B, adt)) return cudaErrorNotSupported;
    if (!make_chunk_map(&tmXf, Xf, nc, g.KG, g.F, 1, 1, 8)) return cudaErrorNotSupported;
    DftArgs a{};
    a.status = status; a.bias = nullptr; a.n_begin = n_begin; a.nc = nc;
    a.m_tiles = (nc + TILE_M - 1) / TILE_M; a.groups = g.KG; a.num_items = a.m_tiles * g.KG; a.act_bf16 = bf ? 1 : 0;
    const uint32_t stage_bytes = (uint32_t)(B / KATOM) * TILE_M * 128;
    a.stages = dft_pick_stages(B, stage_bytes);
    const size_t smem = dft_smem_bytes(B, stage_bytes, a.stages);
    const int grid = a.num_items < max_ctas ? a.num_items : max_ctas;
    if (grid < 1) return cudaSuccess;
    cudaError_t e;
    if (B == 128) {
        if ((e = cudaFuncSetAttribute(dft_fwd_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        return launch_pdl(dft_fwd_tc_kernel<128>, (unsigned)grid, DFT_THREADS, smem, st, tmX, tmW, tmXf, a);
    } else {
        if ((e = cudaFuncSetAttribute(dft_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        return launch_pdl(dft_fwd_tc_kernel<64>, (unsigned)grid, DFT_THREADS, smem, st, tmX, tmW, tmXf, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_dft_inv_tc(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, int max_ctas, cudaStream_t st) {
    const int B = g.B;
    const bool bf = P.x_dtype == BCL_BF16;
    const CUtensorMapDataType adt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmYf, tmW, tmOut;
    if (!make_chunk_map(&tmYf, Yf, nc, g.p, g.FG, TILE_M / TOK_GROUP, 1, (uint32_t)g.FG)) return cudaErrorNotSupported;
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(spec_gemm);
    if (!map_w(&tmW, sb + g.winv_off, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return cudaErrorNotSupported;
    {
        const uint64_t dims[2] = {(uint64_t)g.p * B, (uint64_t)P.N};
        const uint64_t strides[1] = {(uint64_t)P.y_stride * 2};
        const uint32_t box[2] = {KATOM, 32};
        if (!make_tensor_map(&tmOut, adt, 2, P.y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
    }
    DftArgs a{};
    a.status = status; a.bias = P.bias; a.n_begin = n_begin; a.nc = nc;
    a.m_tiles = (nc + TILE_M - 1) / TILE_M; a.groups = g.p; a.num_items = a.m_tiles * g.p; a.act_bf16 = bf ? 1 : 0;
    const uint32_t stage_bytes = (uint32_t)g.FG * 2048;
    a.stages = dft_pick_stages(B, stage_bytes);
    const size_t smem = dft_smem_bytes(B, stage_bytes, a.stages);
    const int grid = a.num_items < max_ctas ? a.num_items : max_ctas;
    if (grid < 1) return cudaSuccess;
    cudaError_t e;
    if (B == 128) {
        if ((e = cudaFuncSetAttribute(dft_inv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        return launch_pdl(dft_inv_tc_kernel<128>, (unsigned)grid, DFT_THREADS, smem, st, tmYf, tmW, tmOut, a);
    } else {
        if ((e = cudaFuncSetAttribute(dft_inv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        return launch_pdl(dft_inv_tc_kernel<64>, (unsigned)grid, DFT_THREADS, smem, st, tmYf, tmW, tmOut, a);
    }
    return cudaGetLastError();
}

}  // namespace gemm
}  // namespace bcl
