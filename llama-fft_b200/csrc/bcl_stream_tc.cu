// bcl_stream_tc.cu -- grouped batch-1 decode with the frequency-domain multiply-accumulate on the TENSOR cores (bf16 spectra).
//
// reference: the same forward as bcl_stream_ws.cu (src/patch_llama_fft.py:246-271, one token per problem), einsum at :260.
// Why: with 16-bit spectra the CUDA-core MAC loop is instruction-bound, not HBM-bound (14 instructions per 8 spectrum bytes and lane:
// unpack, FMA; DESIGN.md 3.3b: 55 % of the HBM peak whatever the warp roles).  Per frequency f the contraction is a small dense
// matrix-vector product  Y_f[j] = sum_i C_f[j,i] X_f[i]  -- on tcgen05 it costs ONE issuing thread:
//     D_f[rows j, 8] = A_f[rows j, K = 2q] * B_f[8, K]^T            (M = 64, N = 8, kind::f16, fp32 accumulation in TMEM)
//     A_f[j, 2i + {0,1}] = (Re, Im) C_f[j,i]                          the spectrum itself, streamed from HBM by 1-D bulk copies
//     B_f row 0 = ( Xr[i], -Xi[i] )_i   -> D[:, 0] = Re Y_f          the x spectrum of the token, bf16, built in shared memory
//     B_f row 1 = ( Xi[i],  Xr[i] )_i   -> D[:, 1] = Im Y_f          (rows 2..7 alias the following bytes: columns never read)
//   slot f = 0 holds the two real bins: A = (C_dc, C_ny), B row 0 = (X_dc, 0), row 1 = (0, X_ny)  -> D = (Y_dc, Y_ny).
// Spectrum layout BCL_LAYOUT_DECODE_TC (tc_spectrum_kernel): bf16, exactly the canonical no-swizzle K-major UMMA tile image
//     [quad of 32 rows][f][8-row group][k/8][row % 8][k % 8]
// so that the bytes of (a run of 8-row groups inside one quad, a run of frequencies) are contiguous: a ring stage is one bulk copy
// and needs no tensor map.  p is padded to 8 rows, 2q to 16 (zeros).  Same algorithmic bytes as the STREAM bf16 layout.
//
// MEASURED (B200, 224 projections of Llama-2-7B per launch, tools/decode_tc_bench.py): parity green (tests/test_gpu_parity.py,
// test_grouped_decode_tensor_core_kernel) but 92 us per step (this tcgen05 variant, BCL_DECODE_TC_IMPL=umma) and 51 us (the
// warp-level HMMA variant further down, the default of this layout) against 29.4 us for the row-per-warp CUDA-core kernel, so the
// plan encoder never picks the layout by itself (BCL_LAYOUT_DECODE_TC is an explicit request).  Why (per-role clock counters, -DBCL_TC_TIMING, and
// builds that skip MMAs): the step costs 42.6 us with NO MMAs issued (B-operand image per segment on the critical path, start-up
// FFT by 4 warps only) and every quarter of the ~700 MMAs per CTA adds 12.5 us = ~140 cycles per M = 64, N = 8, K = 16 MMA whether
// the four k-steps of a frequency are issued back to back or four frequencies are interleaved, whether the descriptors are rebuilt
// or incremented, M = 64 or 128: the dispatch cost of an SS-mode MMA with its A operand in shared memory, ~35x its 4-cycle math.
// A matrix-vector product gives the tensor core nothing to amortise that over; the CUDA-core kernels stay the decode path.
//
//   warp 0          TMA producer (cp.async.bulk + mbarrier complete_tx), runs ahead across tiles, segments and problems
//   warp 1          MMA issuer + TMEM owner: 32 frequencies (256 columns) per accumulator pass, two alternating sets
//   warps 2,3,6,7   X warps: rfft(x) of the NEXT segment (bcl_fft.cuh), then the bf16 B-operand image
//   warps 4,5       drain (TMEM lanes 0-31 / 32-63): columns 0,1 of every D_f -> packed fp32 row spectra in shared memory
//   warps 8..11     finishers: inverse real FFT of 4 rows at once, 1/B, round to x.dtype, + bias, store y (reference :263-271)
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include "bcl_stream_parts.cuh"
#include "bcl_sm100.cuh"

#ifdef BCL_TC_TIMING
#define TCT_DECL long long tct[6] = {0, 0, 0, 0, 0, 0}; const long long tct0 = clock64();
#define TCT(i, ...) { const long long t_ = clock64(); __VA_ARGS__; tct[i] += clock64() - t_; }
#define TCT_PRINT(name) if (blockIdx.x == 3) printf("%s w%d: total %lld  a %lld  b %lld  c %lld  d %lld\n", name, (int)(threadIdx.x >> 5), clock64() - tct0, tct[0], tct[1], tct[2], tct[3]);
#else
#define TCT_DECL
#define TCT(i, ...) __VA_ARGS__;
#define TCT_PRINT(name)
#endif

namespace bcl {
namespace tcd {

using namespace sm100;

constexpr int NX = 4, NF = 4, ND = 2;
constexpr int THREADS = 32 * 12;
constexpr int MAX_STAGES = 10;
constexpr uint32_t STAGE_BYTES = 17408;             // 4 frequencies x 4 groups x (1024 + 16) bytes of a 4096-wide input, rounded up
constexpr int TILE_ROWS = 32;                      // rows of a tile = one quad of the layout (TMEM lanes 0-31)
constexpr int FH = 32;                             // frequencies per accumulator pass: FH * 8 = 256 TMEM columns, two sets alternate
// TMEM lane of accumulator row r for M = 64, cta_group::1.  1: lane r.  2: 16 rows per 32-lane quarter (lane 32 (r / 16) + r % 16).
#ifndef BCL_TC_M64_LAYOUT
#define BCL_TC_M64_LAYOUT 2
#endif
constexpr int PP = 132;                            // row pitch (floats) of the packed row spectra handed to the finishers

__host__ __device__ inline int kp_of(int q) { return (2 * q + 15) / 16 * 16; }
// one 8-row group of one frequency: its core matrices (8 rows x 16 B) back to back + 16 bytes, so that the 8-row groups of an MMA's
// 64 rows start in different banks (without the pad every group starts at the same bank and the operand fetch serialises: measured
// ~220 cycles per MMA)
__host__ __device__ inline int gbytes_of(int q) { return kp_of(q) / 8 * 128 + 16; }
size_t tc_spectrum_bytes(int p, int q, int B) { return (size_t)((p + 7) / 8) * (B / 2) * gbytes_of(q); }

// ---------------------------------------------------------------------------------------------------------
// spectrum precompute: c (p, q, B) fp32 -> BCL_LAYOUT_DECODE_TC (bf16)
// ---------------------------------------------------------------------------------------------------------
template <int B, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) tc_spectrum_kernel(const float* __restrict__ c, __nv_bfloat16* __restrict__ spec, int p, int q) {
    using C = FftCfg<B>;
    constexpr int F = B / 2;
    __shared__ float2 tw[B / 2];
    __shared__ __align__(16) float buf[WARPS][C::G][B];
    __shared__ __align__(16) float2 fscr[WARPS][C::G * C::GROUP_F2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / C::T, t = lane % C::T;
    fill_twiddles<B>(tw, threadIdx.x, WARPS * 32);
    __syncthreads();
    const int P8 = (p + 7) / 8, KG = kp_of(q) / 8;
    const size_t gb = (size_t)gbytes_of(q) / 2;                           // elements of one 8-row group of one frequency (padded)
    const int64_t nblocks = (int64_t)p * q;
    for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * C::G; base < nblocks; base += (int64_t)gridDim.x * WARPS * C::G) {
        const int64_t blk = base + g;
        const bool active = blk < nblocks;
        float2 a[C::V];
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) a[n1] = active ? *reinterpret_cast<const float2*>(c + blk * B + 2 * (C::T * n1 + t)) : make_float2(0.f, 0.f);
        group_rfft_packed<B>(a, fscr[warp] + g * C::GROUP_F2, buf[warp][g], tw, t, active);
        __syncwarp();
        for (int gg = 0; gg < C::G; ++gg) {
            const int64_t b2 = base + gg;
            if (b2 >= nblocks) break;
            const int j = (int)(b2 / q), i = (int)(b2 - (int64_t)j * q);
            const int grp = j / 8, Q = grp / 4, ng = min(4, P8 - 4 * Q);
            for (int f = lane; f < F; f += 32) {                        // packed slot f: (DC, Ny) for f = 0, (Re, Im) otherwise
                const size_t off = ((size_t)4 * Q * F + (size_t)f * ng + (grp - 4 * Q)) * gb + (size_t)(i / 4) * 64 + (size_t)(j % 8) * 8 + (size_t)(i % 4) * 2;
                *reinterpret_cast<__nv_bfloat162*>(spec + off) = __floats2bfloat162_rn(buf[warp][gg][2 * f], buf[warp][gg][2 * f + 1]);
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_spectrum_tc(const float* c, void* spec, int p, int q, int B, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(spec, 0, tc_spectrum_bytes(p, q, B), st);       // row / K padding = zeros
    if (e != cudaSuccess) return e;
    constexpr int WARPS = 4;
    const int64_t nb = (int64_t)p * q;
    int64_t grid = (nb + WARPS * 4 - 1) / (WARPS * 4);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) return cudaSuccess;
    if (B == 128) tc_spectrum_kernel<128, WARPS><<<(unsigned)grid, WARPS * 32, 0, st>>>(c, reinterpret_cast<__nv_bfloat16*>(spec), p, q);
    else if (B == 64) tc_spectrum_kernel<64, WARPS><<<(unsigned)grid, WARPS * 32, 0, st>>>(c, reinterpret_cast<__nv_bfloat16*>(spec), p, q);
    else return cudaErrorNotSupported;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// the decode kernel
// ---------------------------------------------------------------------------------------------------------
struct Barriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t tfull[2], tempty[2];      // accumulator set complete (tcgen05.commit) / drained (ND arrivals)
    uint64_t xbfull, xbempty;          // B-operand image of a segment ready (1 arrival) / its MMAs retired (tcgen05.commit)
    uint64_t pfull[2], pempty[2];      // packed row spectra of a tile ready (ND arrivals) / consumed (NF arrivals)
    uint32_t tmem_base, pad;
};
static_assert(sizeof(Barriers) <= 512, "barrier block");

// A tile = a run of 8-row groups inside one 32-row quad of one segment.
struct Tile { int ga, gb, Q, ng; };
__device__ __forceinline__ int seg_tiles(int j0, int j1) { return (j1 + 31) / 32 - j0 / 32; }
__device__ __forceinline__ Tile seg_tile(int j0, int j1, int P8, int k) {
    Tile t;
    t.Q = j0 / 32 + k;
    t.ga = max(j0 / 8, 4 * t.Q);
    t.gb = min((j1 + 7) / 8, 4 * t.Q + 4);
    t.ng = min(4, P8 - 4 * t.Q);
    return t;
}
// frequencies per ring stage for a tile of ntg groups: a power of two, <= 16, stage <= 16 KB
__device__ __forceinline__ int stage_freqs(int ntg, int gbytes) {
    int fs = 16;
    while (fs > 1 && (uint32_t)(fs * ntg * gbytes) > STAGE_BYTES) fs >>= 1;
    return fs;
}

// tcgen05.mma with the two 64-bit shared-memory descriptors given as (lo, hi) words: inside a tile only the start-address fields (lo)
// change, by constants -- the issuing thread spends ~5 instructions per MMA instead of rebuilding both descriptors (measured: ~220
// cycles per MMA with umma_desc_k_nosw in the loop; the issue rate of this ONE thread bounds the kernel, the tensor pipe idles).
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .b64 ad, bd;\n\t.reg .pred p;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int B>
__global__ void __launch_bounds__(THREADS, 1) stream_tc_kernel(const ProblemDev* __restrict__ problems, const int* __restrict__ cta_seg_ptr,
                                                              const SegWork* __restrict__ segs, int S, int qmax) {
    using FC = FftCfg<B>;
    constexpr int F = B / 2;
    constexpr int NH = F / FH > 0 ? F / FH : 1;                       // accumulator passes per tile
    constexpr int FPH = F / NH;                                      // frequencies per pass
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Barriers* bars = reinterpret_cast<Barriers*>(smem_raw);
    float2* tw = reinterpret_cast<float2*>(smem_raw + 512);
    unsigned char* ring = smem_raw + 2048;
    const int KGmax = kp_of(qmax) / 8;
    const uint32_t xfs_max = (uint32_t)KGmax * 32 + 16;              // byte stride between frequencies of the B image (odd multiple of 16)
    unsigned char* xb = ring + (size_t)S * STAGE_BYTES;              // [F][KG][2][8] bf16 (+ pad)
    float* part = reinterpret_cast<float*>(xb + (((size_t)F * xfs_max + 256 + 127) & ~(size_t)127));   // [2][32][PP]
    float2* fscr_all = reinterpret_cast<float2*>(part + 2 * TILE_ROWS * PP);
    float* xs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(fscr_all) + (size_t)(NX + NF) * FC::WARP_SCRATCH_BYTES);   // [qmax][B]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->tfull[b], 1); mbar_init(&bars->tempty[b], ND); }
        mbar_init(&bars->xbfull, 1); mbar_init(&bars->xbempty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->pfull[b], ND); mbar_init(&bars->pempty[b], NF); }
        fence_barrier_init();
    }
    fill_twiddles<B>(tw, threadIdx.x, THREADS);
    if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    const int s0 = cta_seg_ptr[blockIdx.x], s1 = cta_seg_ptr[blockIdx.x + 1];

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            TCT_DECL
            for (int s = s0; s < s1; ++s) {
                const SegWork w = segs[s];
                const ProblemDev* P = problems + w.g;
                const int q = P->q, P8 = (P->p + 7) / 8, gbytes = gbytes_of(q);
                const unsigned char* spec = reinterpret_cast<const unsigned char*>(P->spec);
                const int nt = seg_tiles(w.j0, w.j1);
                for (int k = 0; k < nt; ++k) {
                    const Tile t = seg_tile(w.j0, w.j1, P8, k);
                    const int ntg = t.gb - t.ga;
                    const int fs = stage_freqs(ntg, gbytes);
                    const unsigned char* qbase = spec + (size_t)4 * t.Q * F * gbytes + (size_t)(t.ga - 4 * t.Q) * gbytes;
                    const size_t fstride = (size_t)t.ng * gbytes;
                    const uint32_t fbytes = (uint32_t)(ntg * gbytes);
                    for (int f0 = 0; f0 < F; f0 += fs) {
                        TCT(0, if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) goto done;)
                        unsigned char* dst = ring + (size_t)stage * STAGE_BYTES;
                        mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)fs * fbytes);
                        if (ntg == t.ng) {                                       // whole quad: consecutive frequencies are contiguous
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                                         "l"(qbase + (size_t)f0 * fstride), "r"((uint32_t)fs * fbytes), "r"(smem_u32(&bars->full[stage])) : "memory");
                        } else {
                            for (int fl = 0; fl < fs; ++fl)
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + (size_t)fl * fbytes)),
                                             "l"(qbase + (size_t)(f0 + fl) * fstride), "r"(fbytes), "r"(smem_u32(&bars->full[stage])) : "memory");
                        }
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
            }
            TCT_PRINT("producer(a=wait_empty)")
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(64, 8);
            int stage = 0; uint32_t ph = 0;
            uint32_t pass = 0;                                                   // accumulator passes so far (tfull / tempty phases)
            TCT_DECL
            for (int s = s0; s < s1; ++s) {
                const SegWork w = segs[s];
                const ProblemDev* P = problems + w.g;
                const int q = P->q, P8 = (P->p + 7) / 8, gbytes = gbytes_of(q);
                const int ksteps = kp_of(q) / 16;
                const uint32_t xfs = (uint32_t)(kp_of(q) / 8) * 32 + 16;
                TCT(0, if (!mbar_wait(&bars->xbfull, (uint32_t)(s - s0) & 1u)) goto done;)
                tc_fence_after_sync();
                const int nt = seg_tiles(w.j0, w.j1);
                for (int k = 0; k < nt; ++k) {
                    const Tile t = seg_tile(w.j0, w.j1, P8, k);
                    const int ntg = t.gb - t.ga;
                    const int fs = stage_freqs(ntg, gbytes);
                    const uint32_t fbytes = (uint32_t)(ntg * gbytes);
                    for (int h = 0; h < NH; ++h, ++pass) {
                        const uint32_t set = pass & 1u;
                        TCT(1, if (!mbar_wait(&bars->tempty[set], ((pass >> 1) & 1u) ^ 1u)) goto done;)   // the drain warps have read this set's previous pass
                        tc_fence_after_sync();
                        // MMAs that accumulate into the same D_f are dependent (the next k-step waits for the previous one: ~170 cycles each
                        // for these tiny shapes), so FOUR frequencies are issued interleaved, k-step by k-step: four independent chains.
                        // A batch is 4 frequencies = a quarter of a wide stage (fs >= 4), or 4 / fs consecutive stages.
                        const uint32_t ahi = ((uint32_t)gbytes >> 4) | (1u << 14), bhi = (128u >> 4) | (1u << 14);
                        for (int f0 = h * FPH; f0 < (h + 1) * FPH; f0 += 4) {
                            uint32_t alo[4], blo[4];
                            const int nst = fs >= 4 ? 1 : 4 / fs;                // stages this batch touches
                            const bool first = fs < 4 || ((f0 - h * FPH) % fs) == 0, last = fs < 4 || ((f0 + 4 - h * FPH) % fs) == 0;
                            if (first) {
                                int st = stage; uint32_t p2 = ph;
                                for (int k = 0; k < nst; ++k) {
                                    TCT(2, if (!mbar_wait(&bars->full[st], p2)) goto done;)
                                    if (++st == S) { st = 0; p2 ^= 1u; }
                                }
                                tc_fence_after_sync();
                            }
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int f = f0 + b;
                                const int stg = fs >= 4 ? stage : (stage + b / fs) % S;
                                const int fl = fs >= 4 ? (f - h * FPH) % fs : b % fs;
                                const uint32_t sa = smem_u32(ring + (size_t)stg * STAGE_BYTES) + (uint32_t)fl * fbytes;
                                alo[b] = ((sa & 0x3ffffu) >> 4) | ((128u >> 4) << 16);
                                blo[b] = (((smem_u32(xb) + (uint32_t)f * xfs) & 0x3ffffu) >> 4) | ((32u >> 4) << 16);
                            }
                            const uint32_t d = tmem_base + set * 256u + (uint32_t)((f0 - h * FPH) * 8);
                            for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
                                for (int b = 0; b < 4; ++b) umma_lohi(d + 8u * b, alo[b] + 16u * ks, ahi, blo[b] + 4u * ks, bhi, idesc, ks != 0 ? 1u : 0u);
                            }
                            if (last) {
                                for (int k = 0; k < nst; ++k) {
                                    umma_commit(&bars->empty[stage]);
                                    if (++stage == S) { stage = 0; ph ^= 1u; }
                                }
                            }
                        }
                        umma_commit(&bars->tfull[set]);
                    }
                }
                umma_commit(&bars->xbempty);                                     // the B image may be rebuilt once these MMAs have retired
            }
            TCT_PRINT("mma(a=wait_xb b=wait_tempty c=wait_full)")
        }
    } else if (warp == 4 || warp == 5) {
        // ===================== drain: columns 0,1 of every D_f -> packed row spectra (warp 4: TMEM lanes 0-31, warp 5: lanes 32-63) =========
        const int d = warp - 4;
        uint32_t pass = 0, tile_no = 0;
        TCT_DECL
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const int nt = seg_tiles(w.j0, w.j1);
            for (int k = 0; k < nt; ++k, ++tile_no) {
                const int buf = (int)(tile_no & 1u);
                TCT(0, if (!__all_sync(0xffffffffu, mbar_wait(&bars->pempty[buf], ((tile_no >> 1) & 1u) ^ 1u))) goto done;)   // finishers are done with this buffer
#if BCL_TC_M64_LAYOUT == 2
                const int row = 16 * d + lane;                                   // lanes 0-15 of quarter d hold rows 16 d .. 16 d + 15
                const bool has = lane < 16;
#else
                const int row = lane;                                            // quarter 0 holds rows 0-31
                const bool has = d == 0;
#endif
                float* prow = part + ((size_t)buf * TILE_ROWS + (has ? row : 0)) * PP;
                for (int h = 0; h < NH; ++h, ++pass) {
                    const uint32_t set = pass & 1u;
                    TCT(1, if (!__all_sync(0xffffffffu, mbar_wait(&bars->tfull[set], (pass >> 1) & 1u))) goto done;)
                    tc_fence_after_sync();
#pragma unroll
                    for (int c8 = 0; c8 < FPH / 8; ++c8) {                       // 64 columns = 8 frequencies per load
                        uint32_t r[64];
                        const uint32_t taddr = tmem_base + ((uint32_t)(32 * d) << 16) + set * 256u + (uint32_t)(c8 * 64);
                        tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                        tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                        tmem_ld_wait();
                        if (has) {
                            float4* dst = reinterpret_cast<float4*>(prow + 2 * (h * FPH + c8 * 8));
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                dst[u] = make_float4(__uint_as_float(r[16 * u]), __uint_as_float(r[16 * u + 1]), __uint_as_float(r[16 * u + 8]), __uint_as_float(r[16 * u + 9]));
                        }
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tempty[set]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->pfull[buf]);
            }
        }
        if (lane == 0) { TCT_PRINT("drain(a=wait_pempty b=wait_tfull)") }
    } else if (warp == 2 || warp == 3 || warp == 6 || warp == 7) {
        // ===================== X warps: rfft(x) of each segment's token, then the bf16 B-operand image =====================
        const int xw = warp < 4 ? warp - 2 : warp - 4;
        const int xt = xw * 32 + lane;
        float2* fscr = fscr_all + (size_t)xw * (FC::WARP_SCRATCH_BYTES / 8);
        pdl_wait();                                                              // x may have been written by the previous grid
        int prev_g = -1;
        TCT_DECL
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const ProblemDev P = problems[w.g];
            const int q = P.q, KG = kp_of(q) / 8;
            const uint32_t xfs = (uint32_t)KG * 32 + 16;
            TCT(0, if (w.g != prev_g) load_x_spectrum<B, 1, NX>(P, 0, xs, tw, fscr, lane, xw);)     // (same problem as the previous segment: xs still holds it)
            prev_g = w.g;
            asm volatile("bar.sync 2, %0;" ::"r"(32 * NX) : "memory");           // xs complete
            TCT(1, if (s > s0 && !__all_sync(0xffffffffu, mbar_wait(&bars->xbempty, (uint32_t)(s - s0 - 1) & 1u))) goto done;)   // MMAs of the previous segment retired
            const long long tc_ = clock64(); (void)tc_;
            // one pass over (input block i < Kp / 2, frequency f): blocks past q are the K padding -- zeros (the A image is zero there too,
            // but 0 * NaN would poison the row)
            const int ip = KG * 4;
            for (int idx = xt; idx < F * ip; idx += 32 * NX) {
                const int i = idx / F, f = idx - i * F;                          // F is a power of two
                __nv_bfloat162 r0 = __floats2bfloat162_rn(0.f, 0.f), r1 = r0;
                if (i < q) {
                    const float2 v = *reinterpret_cast<const float2*>(xs + (size_t)i * B + 2 * f);
                    if (f == 0) { r0 = __floats2bfloat162_rn(v.x, 0.f); r1 = __floats2bfloat162_rn(0.f, v.y); }
                    else { r0 = __floats2bfloat162_rn(v.x, -v.y); r1 = __floats2bfloat162_rn(v.y, v.x); }
                }
                unsigned char* dst = xb + (size_t)f * xfs + (size_t)(i / 4) * 32 + (size_t)(i % 4) * 4;
                *reinterpret_cast<__nv_bfloat162*>(dst) = r0;                    // B row 0
                *reinterpret_cast<__nv_bfloat162*>(dst + 16) = r1;               // B row 1
            }
            fence_proxy_async_smem();                                            // generic-proxy writes -> visible to the tensor core's smem reads
            asm volatile("bar.sync 2, %0;" ::"r"(32 * NX) : "memory");
            if (xt == 0) mbar_arrive(&bars->xbfull);
#ifdef BCL_TC_TIMING
            tct[2] += clock64() - tc_;
#endif
        }
        if (lane == 0 && xw == 0) { TCT_PRINT("xwarp(a=fft b=wait_xbempty c=convert)") }
    } else {
        // ===================== finishers =====================
        const int fi = warp - 8;                                                 // warps 8 .. 11
        float2* fscr = fscr_all + (size_t)(NX + fi) * (FC::WARP_SCRATCH_BYTES / 8);
        const int g = lane / FC::T, t = lane % FC::T;
        pdl_wait();                                                              // y may still be read by the previous grid
        uint32_t tile_no = 0;
        TCT_DECL
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const ProblemDev P = problems[w.g];
            const int P8 = (P.p + 7) / 8;
            const int nt = seg_tiles(w.j0, w.j1);
            for (int k = 0; k < nt; ++k, ++tile_no) {
                const Tile tl = seg_tile(w.j0, w.j1, P8, k);
                const int buf = (int)(tile_no & 1u);
                TCT(0, if (!__all_sync(0xffffffffu, mbar_wait(&bars->pfull[buf], (tile_no >> 1) & 1u))) goto done;)
                const int jbase = tl.ga * 8;                                     // tile row r <-> output block-row jbase + r
                const int nrows = min(w.j1, tl.gb * 8) - jbase;
                const float* pb = part + (size_t)buf * TILE_ROWS * PP;
                for (int r0 = fi * FC::G; r0 < nrows; r0 += NF * FC::G) {
                    const int r = min(r0 + g, nrows - 1);                        // idle lane groups shadow the last row (no stores)
                    float2 out[FC::V];
                    group_irfft_packed<B>(pb + (size_t)r * PP, fscr + g * FC::GROUP_F2, tw, t, out);
                    if (r0 + g < nrows) store_block_epilogue<B>(P, 0, jbase + r, t, out);
                    __syncwarp();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->pempty[buf]);
            }
        }
        if (lane == 0 && fi == 0) { TCT_PRINT("finisher(a=wait_pfull)") }
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

size_t smem_bytes(int B, int qmax, int S) {
    const size_t fscr = B == 64 ? FftCfg<64>::WARP_SCRATCH_BYTES : FftCfg<128>::WARP_SCRATCH_BYTES;
    const size_t xfs = (size_t)(kp_of(qmax) / 8) * 32 + 16;
    const size_t xbb = ((size_t)(B / 2) * xfs + 256 + 127) & ~(size_t)127;
    return 2048 + (size_t)S * STAGE_BYTES + xbb + (size_t)2 * TILE_ROWS * PP * 4 + (size_t)(NX + NF) * fscr + (size_t)qmax * B * 4;
}

int pick_stages(int B, int spec_dtype, int nt, int qmax) {
    if (nt != 1 || spec_dtype != BCL_BF16 || (B != 64 && B != 128)) return 0;
    if (gbytes_of(qmax) * 4 > (int)STAGE_BYTES) return 0;             // one frequency of a full quad must fit a stage (q <= 128)
    const size_t budget = 227 * 1024;
    const size_t fixed = smem_bytes(B, qmax, 0);
    if (fixed + 5 * STAGE_BYTES > budget) return 0;               // the MMA issuer works on up to four stages at once
    int S = (int)((budget - fixed) / STAGE_BYTES);
    return S > MAX_STAGES ? MAX_STAGES : S;
}

template <int B>
static cudaError_t launch_t(const GroupPlanHeader& H, const unsigned char* d, cudaStream_t st) {
    auto kern = stream_tc_kernel<B>;
    const int S = pick_stages(B, H.spec_dtype, H.nt, H.qmax);
    if (S < 5) return cudaErrorNotSupported;
    const size_t smem = smem_bytes(B, H.qmax, S);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)H.num_ctas);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const ProblemDev* probs = reinterpret_cast<const ProblemDev*>(d + H.off_problems);
    const int* cta_ptr = reinterpret_cast<const int*>(d + H.off_cta_ptr);
    const SegWork* segs = reinterpret_cast<const SegWork*>(d + H.off_segs);
    e = cudaLaunchKernelEx(&cfg, kern, probs, cta_ptr, segs, S, (int)H.qmax);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}


// =========================================================================================================
// The same matrix-vector products with warp-level mma.sync (HMMA m16n8k16, bf16 x bf16 -> fp32) instead of tcgen05.
// A tcgen05.mma costs ~140 cycles of dispatch whatever its size (above); an HMMA is an ordinary warp instruction: a 16-row x 16-k
// slice of A_f comes out of the ring with one ldmatrix.x4 (the operand image's 8 x 16-byte core matrices are exactly ldmatrix's 8 x 8
// b16 matrices), B_f with one ldmatrix.x2, the accumulator (16 rows x 8 columns, columns 0-1 = Re / Im Y_f) stays in 4 registers.
// Per 8 spectrum bytes and lane the CUDA-core MAC loop issues ~14 instructions, this one ~2.  No TMEM, no drain warps.
// MEASURED: parity green, 51 us per step (55 us before a MAC warp owned whole stages and the first FFT was shared by all warps).
// The MAC itself is no longer the cost (waits: 6-10k cycles for data per warp and step); the frequency-major image is: the rows of a
// tile complete only with its LAST frequency, so nothing can be inverse-transformed while a tile streams (5-8k cycles of finisher
// tail per tile, in the CUDA-core kernels a row is finished while the next one streams), and the B image of the next segment can
// only be built once every MAC warp has left the current one (5-16k cycles per segment switch with q = 32 / 86).  A row-major
// operand image would need K = frequency, which is not what the contraction sums over.
//   warp 0        TMA producer (as above)
//   warps 1..8    MAC warps: warp m takes the frequencies f = m (mod 8) of every stage; Y_f of the tile's rows -> packed row spectra
//   warps 9..12   X warps: rfft(x) of the next segment, then the bf16 B image
//   warps 13..15  finishers: inverse real FFT of 4 rows at once + the reference epilogue
// =========================================================================================================
namespace hm {

constexpr int NM = 8, NX = 4, NF = 3;
constexpr int THREADS = 32 * (1 + NM + NX + NF);

struct Barriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];      // a stage is consumed by ONE MAC warp (stage number mod NM): empty = 1 arrival
    uint64_t xbfull, xbempty;                          // B image ready (1 arrival) / every MAC warp is done with the segment (NM)
    uint64_t pfull[2], pempty[2];                      // packed row spectra of a tile complete (NM) / consumed (NF); two buffers alternate
    volatile int seq[MAX_STAGES];                      // stage number the producer is filling / has filled the slot with: a MAC warp takes every
                                                       // NM-th stage, so it can be MORE than one phase ahead of a slot's barrier, and a parity wait
                                                       // cannot tell phase n + 2 from phase n (same guard as bcl_stream_rw.cu)
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void hmma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int B>
__global__ void __launch_bounds__(THREADS, 1) stream_hm_kernel(const ProblemDev* __restrict__ problems, const int* __restrict__ cta_seg_ptr,
                                                              const SegWork* __restrict__ segs, int S, int qmax) {
    using FC = FftCfg<B>;
    constexpr int F = B / 2;
    constexpr int XPITCH = B + 2;                                    // floats per block of the fp32 x spectrum (bank spread for the image pass)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Barriers* bars = reinterpret_cast<Barriers*>(smem_raw);
    float2* tw = reinterpret_cast<float2*>(smem_raw + 512);
    unsigned char* ring = smem_raw + 2048;
    const int KGmax = kp_of(qmax) / 8;
    const uint32_t xfs_max = (uint32_t)KGmax * 32 + 16;
    unsigned char* xb = ring + (size_t)S * STAGE_BYTES;              // [F][KG][2][8] bf16 (+ pad)
    float* part = reinterpret_cast<float*>(xb + (((size_t)F * xfs_max + 256 + 127) & ~(size_t)127));   // [2][32][PP]
    float2* fscr_all = reinterpret_cast<float2*>(part + 2 * TILE_ROWS * PP);
    float* xs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(fscr_all) + (size_t)(NX + NF) * FC::WARP_SCRATCH_BYTES);   // [qmax][XPITCH]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); bars->seq[s] = -1; }
        mbar_init(&bars->xbfull, 1); mbar_init(&bars->xbempty, NM);
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->pfull[b], NM); mbar_init(&bars->pempty[b], NF); }
        fence_barrier_init();
    }
    fill_twiddles<B>(tw, threadIdx.x, THREADS);
    __syncthreads();
    pdl_launch_dependents();
    const int s0 = cta_seg_ptr[blockIdx.x], s1 = cta_seg_ptr[blockIdx.x + 1];

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0; int nstage = 0;
            TCT_DECL
            for (int s = s0; s < s1; ++s) {
                const SegWork w = segs[s];
                const ProblemDev* P = problems + w.g;
                const int q = P->q, P8 = (P->p + 7) / 8, gbytes = gbytes_of(q);
                const unsigned char* spec = reinterpret_cast<const unsigned char*>(P->spec);
                const int nt = seg_tiles(w.j0, w.j1);
                for (int k = 0; k < nt; ++k) {
                    const Tile t = seg_tile(w.j0, w.j1, P8, k);
                    const int ntg = t.gb - t.ga;
                    const int fs = stage_freqs(ntg, gbytes);
                    const unsigned char* qbase = spec + (size_t)4 * t.Q * F * gbytes + (size_t)(t.ga - 4 * t.Q) * gbytes;
                    const size_t fstride = (size_t)t.ng * gbytes;
                    const uint32_t fbytes = (uint32_t)(ntg * gbytes);
                    for (int f0 = 0; f0 < F; f0 += fs) {
                        TCT(0, if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) return;)
                        unsigned char* dst = ring + (size_t)stage * STAGE_BYTES;
                        bars->seq[stage] = nstage++;                             // from now on full[stage] is in this stage number's phase
                        mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)fs * fbytes);
                        if (ntg == t.ng) {
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                                         "l"(qbase + (size_t)f0 * fstride), "r"((uint32_t)fs * fbytes), "r"(smem_u32(&bars->full[stage])) : "memory");
                        } else {
                            for (int fl = 0; fl < fs; ++fl)
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + (size_t)fl * fbytes)),
                                             "l"(qbase + (size_t)(f0 + fl) * fstride), "r"(fbytes), "r"(smem_u32(&bars->full[stage])) : "memory");
                        }
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
            }
            TCT_PRINT("hm producer(a=wait_empty)")
        }
    } else if (warp <= NM) {
        // ===================== MAC warps (HMMA) =====================
        const int m = warp - 1;
        uint32_t tile_no = 0;
        int nstage = 0;                                                         // running stage number of the CTA's stream (as the producer counts)
        const int mrow = lane & 7, mid = lane >> 3;                              // ldmatrix: this lane addresses row mrow of 8 x 8 matrix mid
        TCT_DECL
#ifdef BCL_TC_TIMING
        long long st_pdl = 0, st_fft = 0, st_xb[4] = {0, 0, 0, 0}, st_tile[6] = {0, 0, 0, 0, 0, 0};
#endif
        static_assert((size_t)NM * FC::WARP_SCRATCH_BYTES <= (size_t)2 * TILE_ROWS * PP * 4, "first-segment FFT scratch of the MAC warps aliases the part buffers");
        if (s0 < s1) {                                                           // first segment: every compute warp helps with rfft(x)
            const ProblemDev P0 = problems[segs[s0].g];
            pdl_wait();
#ifdef BCL_TC_TIMING
            st_pdl = clock64() - tct0;
#endif
            load_x_spectrum<B, 1, NM + NX + NF, XPITCH>(P0, 0, xs, tw, reinterpret_cast<float2*>(part) + (size_t)m * (FC::WARP_SCRATCH_BYTES / 8), lane, NX + NF + m);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX + NF)) : "memory");
#ifdef BCL_TC_TIMING
            st_fft = clock64() - tct0;
#endif
        }
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const ProblemDev* P = problems + w.g;
            const int q = P->q, P8 = (P->p + 7) / 8, gbytes = gbytes_of(q);
            const int ksteps = kp_of(q) / 16;
            const uint32_t xfs = (uint32_t)(kp_of(q) / 8) * 32 + 16;
            TCT(0, if (!__all_sync(0xffffffffu, mbar_wait(&bars->xbfull, (uint32_t)(s - s0) & 1u))) return;)
#ifdef BCL_TC_TIMING
            if (s - s0 < 4) st_xb[s - s0] = clock64() - tct0;
#endif
            const int nt = seg_tiles(w.j0, w.j1);
            for (int k = 0; k < nt; ++k, ++tile_no) {
                const Tile t = seg_tile(w.j0, w.j1, P8, k);
                const int ntg = t.gb - t.ga;
                const int fs = stage_freqs(ntg, gbytes);
                const uint32_t fbytes = (uint32_t)(ntg * gbytes);
                const int halves = (ntg + 1) / 2;
                const int buf = (int)(tile_no & 1u);
                float* pbuf = part + (size_t)buf * TILE_ROWS * PP;
                TCT(1, if (!__all_sync(0xffffffffu, mbar_wait(&bars->pempty[buf], ((tile_no >> 1) & 1u) ^ 1u))) return;)     // the finishers are done with this buffer
                for (int f0 = 0; f0 < F; f0 += fs, ++nstage) {
                    if (nstage % NM != m) continue;                              // another warp's stage: not even waited for
                    const int stage = nstage % S;
                    {
                        const uint64_t t0 = global_ns();
                        uint32_t spin = 0;
                        while (bars->seq[stage] != nstage) {                     // my stage's turn in this slot
                            __nanosleep(64);
                            if ((++spin & 1023u) == 0u && global_ns() - t0 > 2000000000ull) return;
                        }
                    }
                    TCT(2, if (!__all_sync(0xffffffffu, mbar_wait(&bars->full[stage], (uint32_t)(nstage / S) & 1u))) return;)
                    const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE_BYTES);
                    for (int fl = 0; fl < fs; ++fl) {
                        const int f = f0 + fl;
                        const uint32_t a_f = sa + (uint32_t)fl * fbytes + (uint32_t)(mid & 1) * (uint32_t)gbytes + (uint32_t)(mid >> 1) * 128u + (uint32_t)mrow * 16u;
                        const uint32_t b_f = smem_u32(xb) + (uint32_t)f * xfs + (uint32_t)(mid & 1) * 32u + (uint32_t)mrow * 16u;
                        // Both 16-row halves of the tile and even / odd k-steps accumulate separately: four independent HMMA chains per
                        // frequency instead of one of 2 * ksteps dependent instructions (an HMMA waits ~tens of cycles for its accumulator).
                        float d[2][2][4];
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
#pragma unroll
                                for (int c = 0; c < 4; ++c) d[h][e][c] = 0.f;
                        const uint32_t a_h1 = a_f + 2u * (uint32_t)gbytes;
#pragma unroll 2
                        for (int ks = 0; ks < ksteps; ks += 2) {
                            uint32_t a0[4], a1[4], a2[4], a3[4], b0[2], b1[2];
                            const bool odd = ks + 1 < ksteps;
                            ldmatrix_x2(b_f + (uint32_t)ks * 64u, b0);
                            ldmatrix_x4(a_f + (uint32_t)ks * 256u, a0);
                            if (halves > 1) ldmatrix_x4(a_h1 + (uint32_t)ks * 256u, a1);
                            if (odd) {
                                ldmatrix_x2(b_f + (uint32_t)(ks + 1) * 64u, b1);
                                ldmatrix_x4(a_f + (uint32_t)(ks + 1) * 256u, a2);
                                if (halves > 1) ldmatrix_x4(a_h1 + (uint32_t)(ks + 1) * 256u, a3);
                            }
                            hmma_bf16(d[0][0], a0, b0);
                            if (halves > 1) hmma_bf16(d[1][0], a1, b0);
                            if (odd) {
                                hmma_bf16(d[0][1], a2, b1);
                                if (halves > 1) hmma_bf16(d[1][1], a3, b1);
                            }
                        }
                        if ((lane & 3) == 0) {                                           // accumulator columns 0, 1 = (Re, Im) Y_f; slot 0: (DC, Nyquist)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h >= halves) break;
                                const int r0 = 16 * h + (lane >> 2);
                                if (r0 < ntg * 8) *reinterpret_cast<float2*>(pbuf + (size_t)r0 * PP + 2 * f) = make_float2(d[h][0][0] + d[h][1][0], d[h][0][1] + d[h][1][1]);
                                if (r0 + 8 < ntg * 8) *reinterpret_cast<float2*>(pbuf + (size_t)(r0 + 8) * PP + 2 * f) = make_float2(d[h][0][2] + d[h][1][2], d[h][0][3] + d[h][1][3]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->empty[stage]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->pfull[buf]);
#ifdef BCL_TC_TIMING
                if (tile_no < 6) st_tile[tile_no] = clock64() - tct0;
#endif
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xbempty);
        }
        if (lane == 0 && (m == 0 || m == 5)) { TCT_PRINT("hm mac(a=wait_xb b=wait_pempty c=wait_full)") }
#ifdef BCL_TC_TIMING
        if (lane == 0 && m == 0 && blockIdx.x == 3)
            printf("hm-tl mac: pdl %lld fft %lld xb %lld %lld %lld %lld tiles %lld %lld %lld %lld %lld %lld end %lld\n", st_pdl, st_fft, st_xb[0], st_xb[1], st_xb[2], st_xb[3],
                   st_tile[0], st_tile[1], st_tile[2], st_tile[3], st_tile[4], st_tile[5], clock64() - tct0);
#endif
    } else if (warp <= NM + NX) {
        // ===================== X warps =====================
        const int xw = warp - 1 - NM;
        float2* fscr = fscr_all + (size_t)xw * (FC::WARP_SCRATCH_BYTES / 8);
        ProblemDev P0{};
        if (s0 < s1) P0 = problems[segs[s0].g];
        pdl_wait();                                                              // x may have been written by the previous grid
        for (int s = s0; s < s1; ++s) {
            const ProblemDev P = s == s0 ? P0 : problems[segs[s].g];
            const int q = P.q, KG = kp_of(q) / 8;
            const uint32_t xfs = (uint32_t)KG * 32 + 16;
            if (s == s0) {                                                       // first segment: together with the (still idle) finisher warps
                load_x_spectrum<B, 1, NM + NX + NF, XPITCH>(P, 0, xs, tw, fscr, lane, xw);
                asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX + NF)) : "memory");
            } else {
                load_x_spectrum<B, 1, NX, XPITCH>(P, 0, xs, tw, fscr, lane, xw);
                asm volatile("bar.sync 2, %0;" ::"r"(32 * NX) : "memory");
                if (!__all_sync(0xffffffffu, mbar_wait(&bars->xbempty, (uint32_t)(s - s0 - 1) & 1u))) return;   // the MAC warps have left the previous image
            }
            // B image: warp xw takes frequencies xw, xw + NX, ..; its lanes the input blocks (blocks past q = K padding: zeros)
            const int ip = KG * 4;
            for (int f = xw; f < F; f += NX) {
                unsigned char* xf = xb + (size_t)f * xfs;
                for (int i = lane; i < ip; i += 32) {
                    __nv_bfloat162 r0 = __floats2bfloat162_rn(0.f, 0.f), r1 = r0;
                    if (i < q) {
                        const float2 v = *reinterpret_cast<const float2*>(xs + (size_t)i * XPITCH + 2 * f);
                        if (f == 0) { r0 = __floats2bfloat162_rn(v.x, 0.f); r1 = __floats2bfloat162_rn(0.f, v.y); }
                        else { r0 = __floats2bfloat162_rn(v.x, -v.y); r1 = __floats2bfloat162_rn(v.y, v.x); }
                    }
                    unsigned char* dst = xf + (size_t)(i >> 2) * 32 + (size_t)(i & 3) * 4;
                    *reinterpret_cast<__nv_bfloat162*>(dst) = r0;
                    *reinterpret_cast<__nv_bfloat162*>(dst + 16) = r1;
                }
            }
            asm volatile("bar.sync 2, %0;" ::"r"(32 * NX) : "memory");
            if (xw == 0 && lane == 0) mbar_arrive(&bars->xbfull);
        }
    } else {
        // ===================== finishers =====================
        const int fi = warp - 1 - NM - NX;
        float2* fscr = fscr_all + (size_t)(NX + fi) * (FC::WARP_SCRATCH_BYTES / 8);
        const int g = lane / FC::T, t = lane % FC::T;
        ProblemDev P0{};
        if (s0 < s1) P0 = problems[segs[s0].g];
        pdl_wait();                                                              // x may have been written / y may still be read by the previous grid
        if (s0 < s1) {                                                           // help with the first segment's rfft(x)
            load_x_spectrum<B, 1, NM + NX + NF, XPITCH>(P0, 0, xs, tw, fscr, lane, NX + fi);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX + NF)) : "memory");
        }
        uint32_t tile_no = 0;
        TCT_DECL
#ifdef BCL_TC_TIMING
        long long st_tile[6] = {0, 0, 0, 0, 0, 0};
#endif
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const ProblemDev P = problems[w.g];
            const int P8 = (P.p + 7) / 8;
            const int nt = seg_tiles(w.j0, w.j1);
            for (int k = 0; k < nt; ++k, ++tile_no) {
                const Tile tl = seg_tile(w.j0, w.j1, P8, k);
                const int buf = (int)(tile_no & 1u);
                const float* pbuf = part + (size_t)buf * TILE_ROWS * PP;
                TCT(0, if (!__all_sync(0xffffffffu, mbar_wait(&bars->pfull[buf], (tile_no >> 1) & 1u))) return;)
                const int jbase = tl.ga * 8;
                const int nrows = min(w.j1, tl.gb * 8) - jbase;
                for (int r0 = fi * FC::G; r0 < nrows; r0 += NF * FC::G) {
                    const int r = min(r0 + g, nrows - 1);
                    float2 out[FC::V];
                    group_irfft_packed<B>(pbuf + (size_t)r * PP, fscr + g * FC::GROUP_F2, tw, t, out);
                    if (r0 + g < nrows) store_block_epilogue<B>(P, 0, jbase + r, t, out);
                    __syncwarp();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->pempty[buf]);
#ifdef BCL_TC_TIMING
                if (tile_no < 6) st_tile[tile_no] = clock64() - tct0;
#endif
            }
        }
        if (lane == 0 && fi == 0) { TCT_PRINT("hm finisher(a=wait_pfull)") }
#ifdef BCL_TC_TIMING
        if (lane == 0 && fi == 0 && blockIdx.x == 3)
            printf("hm-tl fin: tiles %lld %lld %lld %lld %lld %lld end %lld\n", st_tile[0], st_tile[1], st_tile[2], st_tile[3], st_tile[4], st_tile[5], clock64() - tct0);
#endif
    }
}

size_t smem_bytes(int B, int qmax, int S) {
    const size_t fscr = B == 64 ? FftCfg<64>::WARP_SCRATCH_BYTES : FftCfg<128>::WARP_SCRATCH_BYTES;
    const size_t xfs = (size_t)(kp_of(qmax) / 8) * 32 + 16;
    const size_t xbb = ((size_t)(B / 2) * xfs + 256 + 127) & ~(size_t)127;
    return 2048 + (size_t)S * STAGE_BYTES + xbb + (size_t)2 * TILE_ROWS * PP * 4 + (size_t)(NX + NF) * fscr + (size_t)qmax * (B + 2) * 4;
}

int pick_stages(int B, int spec_dtype, int nt, int qmax) {
    if (nt != 1 || spec_dtype != BCL_BF16 || (B != 64 && B != 128)) return 0;
    if (gbytes_of(qmax) * 4 > (int)STAGE_BYTES) return 0;
    const size_t budget = 227 * 1024;
    const size_t fixed = smem_bytes(B, qmax, 0);
    if (fixed + 3 * STAGE_BYTES > budget) return 0;
    int S = (int)((budget - fixed) / STAGE_BYTES);
    return S > MAX_STAGES ? MAX_STAGES : S;
}

template <int B>
static cudaError_t launch_t(const GroupPlanHeader& H, const unsigned char* d, cudaStream_t st) {
    auto kern = stream_hm_kernel<B>;
    const int S = pick_stages(B, H.spec_dtype, H.nt, H.qmax);
    if (S < 3) return cudaErrorNotSupported;
    const size_t smem = smem_bytes(B, H.qmax, S);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)H.num_ctas);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const ProblemDev* probs = reinterpret_cast<const ProblemDev*>(d + H.off_problems);
    const int* cta_ptr = reinterpret_cast<const int*>(d + H.off_cta_ptr);
    const SegWork* segs = reinterpret_cast<const SegWork*>(d + H.off_segs);
    e = cudaLaunchKernelEx(&cfg, kern, probs, cta_ptr, segs, S, (int)H.qmax);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace hm

}  // namespace tcd

size_t stream_tc_spectrum_bytes(int p, int q, int B) { return tcd::tc_spectrum_bytes(p, q, B); }
cudaError_t launch_spectrum_tc(const float* c, void* spec, int p, int q, int B, cudaStream_t st) { return tcd::launch_spectrum_tc(c, spec, p, q, B, st); }
// BCL_DECODE_TC_IMPL=umma: the tcgen05 issuer + TMEM drain variant (measured 92 us); default: warp-level HMMA consumers
static bool tc_use_umma() {
    static const bool on = getenv("BCL_DECODE_TC_IMPL") && !strcmp(getenv("BCL_DECODE_TC_IMPL"), "umma");
    return on;
}
bool stream_tc_supported(int B, int spec_dtype, int nt, int qmax) {
    return tc_use_umma() ? tcd::pick_stages(B, spec_dtype, nt, qmax) >= 5 : tcd::hm::pick_stages(B, spec_dtype, nt, qmax) >= 3;
}
cudaError_t launch_stream_tc(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st) {
    const unsigned char* d = reinterpret_cast<const unsigned char*>(dev_plan);
    if (tc_use_umma()) return H.B == 128 ? tcd::launch_t<128>(H, d, st) : tcd::launch_t<64>(H, d, st);
    return H.B == 128 ? tcd::hm::launch_t<128>(H, d, st) : tcd::hm::launch_t<64>(H, d, st);
}

}  // namespace bcl
