// bcl_stream_rw.cu -- row-per-warp variant of the grouped decode kernel, for 16-bit (bf16) spectra.
//
// reference: the same forward as bcl_stream_ws.cu (src/patch_llama_fft.py:246-271, one token per problem).
// In bcl_stream_ws.cu the eight MAC warps split every block-row eight ways (warp m owns input blocks i = m mod 8) and park eight
// partial spectra per row; the per-row hand-offs (barrier wait, release, parking, the finisher's 8-way reduction) cost ~480
// cycles per row and warp whether the row is 16 KB of fp32 or 8 KB of bf16 -- with bf16 spectra the MAC warps, not HBM, bound
// the kernel (clock stamps: profiles/r1_prefill_notes.md).  Here a MAC warp multiplies a WHOLE row: lane l accumulates its
// packed reals of all q blocks of C_f[row] (ring) times X_f (shared memory) in registers, parks ONE finished spectrum per row,
// and the finisher only inverse-transforms.  Eight rows are in flight per CTA; the hand-offs per row drop eightfold.
//   NP = 3 warps     TMA producers (identical ring protocol: one bulk copy per (row, stage-of-row); producer w takes chunks w mod NP)
//   next 8 warps     row warps: rows r = m, m + 8, ... of the CTA's running row sequence
//   next 6 warps     X warps: rfft(x) of the next segment into the other half of a double-buffered X buffer
//   last 3 warps     finishers: inverse real FFT of 4 rows at once, 1/B, round to x.dtype, + bias, store y
#include <cstdlib>
#include <cstdio>
#include <type_traits>
#include "bcl_stream_parts.cuh"

namespace bcl {
namespace rw {

// Warp mix.  17 to 20 warps put five warps on an SM sub-partition, which caps a thread at 96 registers (16 warps: 128); B = 64 / 128 need
// 92-93, so once the third producer is there two more warps are free.  Measured at 20 warps, bf16, 224 projections (NP / NX / NF):
// 3/5/4: 26.1 us, 3/4/4 (19 warps): 26.4, 3/4/5: 26.7, 2/6/4: 27.1, 3/4/3 (18): 27.2, 3/6/3: 27.5, 4/4/4: 28.6, 4/5/3: 28.8, 2/5/5: 28.9.
// Measured again after the finisher path was shortened (no partial-sum pass, convert-and-pack epilogue, lane twiddles in registers):
// three finishers are enough and the sixth X warp pays -- 3/6/3: 23.1 us (70.2 %), 3/6/3 with 6 group buffers: 23.2, 3/7/2: 23.4,
// 3/5/4: 23.9, 3/5/3: 24.0, 3/4/4: 24.2, 4/5/3: 26.1.
#ifndef BCL_RW_NF
#define BCL_RW_NF 3
#endif
#ifndef BCL_RW_NX
#define BCL_RW_NX 6
#endif
#ifndef BCL_RW_NG
#define BCL_RW_NG BCL_RW_NF
#endif
// NG group buffers for NF finishers (NG a multiple of NF, so a buffer always belongs to the same finisher).  With one buffer per
// finisher a row warp whose group's buffer is still being transformed stalls 4-5k cycles per step (-DBCL_RW_TIMING); with NG = 6
// that wait disappears (0.8k) -- and reappears as a wait for ring chunks: measured 28.6 (NG = 3) / 29.0 (6) / 29.0 (9) / 29.2 us (12).
// The default stays one buffer per finisher.
static_assert(BCL_RW_NG % BCL_RW_NF == 0, "group buffers: a multiple of the finisher count");
#ifndef BCL_RW_NP
#define BCL_RW_NP 3
#endif
// NP producer warps: producer w issues the ring chunks idx = w (mod NP).  One thread needs ~370 cycles per bulk copy (~40 dependent
// instructions: wait, sequence number, expect_tx, the elect / R2UR sequence around UBLKCP; -DBCL_RW_TIMING: 32k busy cycles for 88
// chunks), i.e. 8 KB per 370 cycles = 22 B/clk, just the HBM share of an SM -- and it waits in order, so one late stage holds up the
// copies behind it.  The number of stages is a multiple of NP, so a stage always belongs to the same producer and the parity waits
// stay valid.  Measured (bf16, 224 projections): NP = 1: 28.6 us, 2: 28.4, 3: 27.2, 4: 28.0 (every extra warp costs the others: above 16
// warps an SM sub-partition holds five and the register cap drops from 128 to 96).  ONE producer warp whose lane l owns stage l and polls
// its empty barrier (copies of all free stages issued in one pass, built and measured): 29.9 us -- a pass costs the same ~40
// instructions whether one lane or ten issue, and the lanes that are not ready make the warp sleep.
constexpr int NM = 8, RG = 4;
// B = 256 keeps the 16-warp mix (one producer, 4 X warps, 3 finishers, 128 registers): its transforms need more than the 96
// registers of the 20-warp tier and the spills cost more than the extra warps bring (11008^2, B = 256: 26.7 us -> 40.3 us with 3/5/4).
__host__ __device__ constexpr int mix_np(int B) { return B == 256 ? 1 : BCL_RW_NP; }
__host__ __device__ constexpr int mix_nx(int B) { return B == 256 ? 4 : BCL_RW_NX; }
__host__ __device__ constexpr int mix_nf(int B) { return B == 256 ? 3 : BCL_RW_NF; }
__host__ __device__ constexpr int mix_ng(int B) { return B == 256 ? 3 : BCL_RW_NG; }
__host__ __device__ constexpr int mix_threads(int B) { return 32 * (mix_np(B) + NM + mix_nx(B) + mix_nf(B)); }
constexpr int NG_MAX = BCL_RW_NG > 3 ? BCL_RW_NG : 3;
#ifndef BCL_RW_MAX_STAGES
#define BCL_RW_MAX_STAGES 16
#endif
constexpr int MAX_STAGES = BCL_RW_MAX_STAGES;
__host__ __device__ constexpr int stage_blocks(int B) { return 4096 / B; }       // same chunking as bcl_stream_ws.cu

// -DBCL_RW_TIMING: per-role cycle counters of CTA 3 (printf at the end of the launch; read with tools/small_launch_timing.py or any bench run
// against a library built with the flag).  Off by default: the macros expand to nothing.
#ifdef BCL_RW_TIMING
__device__ inline long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
#define TSTART(v) long long v##_t0 = clk()
#define TSTOP(v, acc) acc += clk() - v##_t0
#define TPRINT(...) printf(__VA_ARGS__)
#else
__device__ inline long long clk() { return 0; }
#define TSTART(v)
#define TSTOP(v, acc)
#define TPRINT(...)
#endif

// The row warps drift apart (a warp's consecutive rows are NM rows = up to NM * nst ring chunks and two groups apart), so a warp
// can be MORE than one phase ahead of a barrier -- and a parity wait cannot tell phase n+2 from phase n.  Wherever that can happen
// the waiter first checks a sequence number in shared memory: `chunk_in[stage]` = the chunk the producer is filling / has filled
// the stage with, `accepts[buf]` = the group the buffer may be parked into next.
struct Barriers {
    uint64_t full[MAX_STAGES];      // ring stage filled by TMA                      (1 arrival + tx bytes)
    uint64_t empty[MAX_STAGES];     // ring stage released by the ONE row warp that owns it
    uint64_t gfull[NG_MAX];           // group buffer complete                          (RG arrivals: one per row)
    uint64_t xfull[2];              // x spectrum of a segment ready in X buffer k    (NX arrivals)
    uint64_t xempty[2];             // every row warp is done with X buffer k          (NM arrivals)
    volatile int chunk_in[MAX_STAGES];
    volatile int accepts[NG_MAX];
};
static_assert(sizeof(Barriers) <= 512, "barrier block");

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ inline void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ inline void mbar_arrive_n(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ inline void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ inline bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ inline void mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    if (mbar_try(addr, parity)) return;
    while (!mbar_try(addr, parity)) __nanosleep(32);
}
__device__ inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ inline void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ inline void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__host__ __device__ inline size_t warp_scratch_bytes(int B) {
    return B == 64 ? FftCfg<64>::WARP_SCRATCH_BYTES : (B == 128 ? FftCfg<128>::WARP_SCRATCH_BYTES : FftCfg<256>::WARP_SCRATCH_BYTES);
}
// the row warps' first-segment FFT scratch fits into X buffer 1
__host__ __device__ inline bool scratch_aliases_x(int B, int qmax) { return (size_t)qmax * B * 4 >= (size_t)NM * warp_scratch_bytes(B); }

template <int B, typename SpecT>
__global__ void __launch_bounds__(mix_threads(B), 1) stream_rw_kernel(const ProblemDev* __restrict__ problems, const int* __restrict__ cta_seg_ptr,
                                                              const SegWork* __restrict__ segs, int S, int qmax) {
    using LM = LaneMap<B>;
    using FC = FftCfg<B>;
    constexpr int NT = 1;
    constexpr int NP = mix_np(B), NX = mix_nx(B), NF = mix_nf(B), NG = mix_ng(B), THREADS = mix_threads(B);
    constexpr int CH = LM::CH, CW = LM::CW;
    constexpr int BLOCK_BYTES = B * (int)sizeof(SpecT);
    constexpr int CB = stage_blocks(B);                              // blocks per ring stage
    constexpr int STAGE_BYTES = CB * BLOCK_BYTES;
    constexpr int NW = CW * (int)sizeof(SpecT) / 4;                  // 32-bit words per lane chunk of a spectrum block
    constexpr int UB = 8;                                            // blocks multiplied per unrolled batch (loads first)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    Barriers* bars = reinterpret_cast<Barriers*>(smem_raw);
    float2* tw = reinterpret_cast<float2*>(smem_raw + 512);
    unsigned char* ring = smem_raw + 512 + 1024;
    float* part = reinterpret_cast<float*>(ring + (size_t)S * STAGE_BYTES);                         // [NG][RG][B]: finished row spectra
    // FFT scratch: the X warps and the finishers own theirs; the row warps only transform during the cooperative first segment, when X
    // buffer 1 is still unused (the X warps write it after the bar.sync that ends that phase) -- so their scratch aliases it when it
    // is large enough, which buys two more ring stages for Llama shapes (9 -> 12 stages was worth 3 us of 34: the ring is what
    // bounds this kernel).
    float2* fscr_all = reinterpret_cast<float2*>(part + NG * RG * B);                               // scratch of NX + NF warps (+ NM if not aliased)
    const bool alias_scr = scratch_aliases_x(B, qmax);
    float* xs_all = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(fscr_all) + (size_t)(NX + NF + (alias_scr ? 0 : NM)) * FC::WARP_SCRATCH_BYTES);   // [2][qmax][B]
    const size_t xs_stride = (size_t)qmax * B;
    float2* row_scr = alias_scr ? reinterpret_cast<float2*>(xs_all + xs_stride) : fscr_all + (size_t)(NX + NF) * (FC::WARP_SCRATCH_BYTES / 8);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); bars->chunk_in[s] = -1; }
        for (int g = 0; g < NG; ++g) { mbar_init(&bars->gfull[g], RG); bars->accepts[g] = g; }
        for (int k = 0; k < 2; ++k) { mbar_init(&bars->xfull[k], NX); mbar_init(&bars->xempty[k], NM); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fill_twiddles<B>(tw, threadIdx.x, THREADS);
    __syncthreads();
    pdl_launch_dependents();                                         // see bcl_stream_ws.cu: spectra and plan are constants of the stream

    const int s0 = cta_seg_ptr[blockIdx.x], s1 = cta_seg_ptr[blockIdx.x + 1];

    if (warp < NP) {
        // ===================== TMA producers =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0; int idx = 0;
            long long t_all0 = clk(), t_e = 0; (void)t_all0; (void)t_e;
            for (int s = s0; s < s1; ++s) {
                const SegWork w = segs[s];
                const ProblemDev* P = problems + w.g;
                const int q = P->q;
                const int nst = (q + CB - 1) / CB;
                const unsigned char* rowp = reinterpret_cast<const unsigned char*>(P->spec) + (size_t)w.j0 * q * BLOCK_BYTES;
                for (int r = w.j0; r < w.j1; ++r, rowp += (size_t)q * BLOCK_BYTES) {
                    for (int cs = 0; cs < nst; ++cs, ++idx) {
                        if (NP > 1 && idx % NP != warp) {                 // another producer's chunk
                            if (++stage == S) { stage = 0; ph ^= 1u; }
                            continue;
                        }
                        { TSTART(e); mbar_wait(&bars->empty[stage], ph ^ 1u); TSTOP(e, t_e); }   // in order and one phase at a time: parity is enough here
                        bars->chunk_in[stage] = idx;                     // from now on full[stage] is in this chunk's phase (the store cannot pass the
                                                                         // acquiring wait above; nothing else is ordered by it: no fence, the producer
                                                                         // thread's ~300 cycles per copy are the scarce resource here)
                        const uint32_t bytes = (uint32_t)min(CB, q - cs * CB) * BLOCK_BYTES;
                        mbar_arrive_expect_tx(&bars->full[stage], bytes);
                        bulk_g2s(ring + (size_t)stage * STAGE_BYTES, rowp + (size_t)cs * STAGE_BYTES, bytes, &bars->full[stage]);
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
            }
            if (blockIdx.x == 3) TPRINT("producer %d: total %lld wait_empty %lld chunks %d\n", warp, clk() - t_all0, t_e, idx);
        }
    } else if (warp < NP + NM) {
        // ===================== row warps =====================
        const int m = warp - NP;
        long long t_all0 = clk(), t_x = 0, t_f = 0, t_g = 0, t_mac = 0, t_first = 0; int nrows = 0;
        (void)t_all0; (void)t_x; (void)t_f; (void)t_g; (void)t_mac; (void)t_first; (void)nrows;
        if (s0 < s1) {                                               // first segment: help with rfft(x) (start-up latency, HBM idles)
            const ProblemDev P0 = problems[segs[s0].g];              // plan data: constant, fetched while the previous grid drains
            pdl_wait();
            load_x_spectrum<B, NT, NM + NX>(P0, 0, xs_all, tw, row_scr + (size_t)m * (FC::WARP_SCRATCH_BYTES / 8), lane, m);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX)) : "memory");
            t_first = clk() - t_all0;
        }
        int chunk0 = 0;                                              // ring chunk index of the segment's first chunk (CTA-wide sequence)
        int row0 = 0;                                                // running row index of the segment's first row
        int grp0 = 0;                                                // running group index of the segment's first group
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const int q = problems[w.g].q;
            const int rows = w.j1 - w.j0;
            const int nst = (q + CB - 1) / CB;
            const int k = s - s0;
            const float* xs = xs_all + (size_t)(k & 1) * xs_stride;
            if (k > 0) { TSTART(x); mbar_wait(&bars->xfull[k & 1], (uint32_t)((k - 1) >> 1) & 1u); TSTOP(x, t_x); }      // buffer 1: segments 1,3,..; buffer 0: 2,4,.. (0 was cooperative)
            // my rows of this segment: (row0 + r) % NM == m
            int r = (m - (row0 % NM) + NM) % NM;
            for (; r < rows; r += NM) {
                int idx = chunk0 + r * nst;
                int stage = idx % S; uint32_t ph = (uint32_t)(idx / S) & 1u;
                // Independent accumulator chains: the two products of every complex MAC go to separate sums (pos / neg) and even / odd
                // blocks to separate sets, so consecutive FMAs never depend on each other (a single chain per packed real made the
                // row warp issue one FMA per ~4-5 cycles: 2.3k cycles per row, clock stamps).
                // Packed fp32 FMAs (fma.rn.f32x2 -> FFMA2, sm_100): p1 = (Σ cr xre, Σ cr xim), p2 = (Σ ci xim, Σ ci xre) per complex bin; the
                // broadcast of cr / ci and the (xim, xre) swap are operand modifiers of the SASS instruction (.F32, .LO_HI), so a complex MAC
                // is 2 issue slots instead of 4 -- this kernel is issue-bound (ncu: ALU 42 %, issue 51 %, HBM 34 %), and every lane of the
                // packed FMA rounds like the scalar one: the sums are bit-identical to the four-FMA form.
                uint64_t p1[2][CH][CW / 2], p2[2][CH][CW / 2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
                        for (int e = 0; e < CW / 2; ++e) { p1[h][ch][e] = 0ull; p2[h][ch][e] = 0ull; }
                }
                for (int cs = 0; cs < nst; ++cs, ++idx) {
                    { TSTART(f);
                    while (bars->chunk_in[stage] != idx) __nanosleep(128);   // my chunk's turn in this stage (see Barriers)
                    mbar_wait(&bars->full[stage], ph);
                    TSTOP(f, t_f); }
                    TSTART(mac);
                    const SpecT* cbuf = reinterpret_cast<const SpecT*>(ring + (size_t)stage * STAGE_BYTES);
                    const int nblk = min(CB, q - cs * CB);
                    const float* xrow = xs + (size_t)cs * CB * B;
                    // one batch of UB blocks; FULL: all of them exist (the common case: no per-block predicate, no zero fill)
                    auto batch = [&](int u0, auto fullc) {
                        constexpr bool FULL = decltype(fullc)::value;
                        RawChunk<SpecT, CW> raw[UB][CH];
                        float xv[UB][CH][CW];
#pragma unroll
                        for (int u = 0; u < UB; ++u) {                // all shared loads first: one latency per batch
                            const bool on = FULL || u0 + u < nblk;
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                if (on) {
                                    const uint32_t* src = reinterpret_cast<const uint32_t*>(cbuf + (size_t)(u0 + u) * B + LM::offset(lane, ch));
                                    if constexpr (NW == 4) { uint4 t = *reinterpret_cast<const uint4*>(src); raw[u][ch].w[0] = t.x; raw[u][ch].w[1] = t.y; raw[u][ch].w[2] = t.z; raw[u][ch].w[3] = t.w; }
                                    else if constexpr (NW == 2) { uint2 t = *reinterpret_cast<const uint2*>(src); raw[u][ch].w[0] = t.x; raw[u][ch].w[1] = t.y; }
                                    else { raw[u][ch].w[0] = *src; }
                                    const float* xp = xrow + (size_t)(u0 + u) * B + LM::offset(lane, ch);
                                    if constexpr (CW == 4) { float4 t = *reinterpret_cast<const float4*>(xp); xv[u][ch][0] = t.x; xv[u][ch][1] = t.y; xv[u][ch][2] = t.z; xv[u][ch][3] = t.w; }
                                    else { float2 t = *reinterpret_cast<const float2*>(xp); xv[u][ch][0] = t.x; xv[u][ch][1] = t.y; }
                                } else {
#pragma unroll
                                    for (int e = 0; e < NW; ++e) raw[u][ch].w[e] = 0u;
#pragma unroll
                                    for (int e = 0; e < CW; ++e) xv[u][ch][e] = 0.f;
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < UB; ++u) {
                            const int h = u & 1;
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                float cv[CW];
                                decode_raw<SpecT, CW>(raw[u][ch], cv);
#pragma unroll
                                for (int e = 0; e < CW; e += 2) {
                                    const float cr = cv[e], ci = cv[e + 1], xre = xv[u][ch][e], xim = xv[u][ch][e + 1];
                                    ffma2(p1[h][ch][e / 2], pack2(cr, cr), pack2(xre, xim));    // (cr xre, cr xim)
                                    ffma2(p2[h][ch][e / 2], pack2(ci, ci), pack2(xim, xre));    // (ci xim, ci xre)
                                }
                            }
                            // packed slot 0/1 of lane 0 is (DC, Nyquist): real x real products, already in p1[..][0][0].lo and p2[..][0][0].lo
                        }
                    
                    };
                    for (int u0 = 0; u0 < nblk; u0 += UB) {
                        if (u0 + UB <= nblk) batch(u0, std::true_type{});
                        else batch(u0, std::false_type{});
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->empty[stage]);   // the producer may refill this stage
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                    TSTOP(mac, t_mac);
                }
                ++nrows;
                float sp[2][CH][CW], sn[2][CH][CW];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
                        for (int e = 0; e < CW; e += 2) {
                            unpack2(p1[h][ch][e / 2], sp[h][ch][e], sp[h][ch][e + 1]);
                            unpack2(p2[h][ch][e / 2], sn[h][ch][e], sn[h][ch][e + 1]);
                        }
                RowAcc<B, NT> acc;
#pragma unroll
                for (int ch = 0; ch < CH; ++ch)
#pragma unroll
                    for (int e = 0; e < CW; e += 2) {
                        acc.a[0][ch][e] = (sp[0][ch][e] + sp[1][ch][e]) - (sn[0][ch][e] + sn[1][ch][e]);
                        acc.a[0][ch][e + 1] = (sp[0][ch][e + 1] + sp[1][ch][e + 1]) + (sn[0][ch][e + 1] + sn[1][ch][e + 1]);
                    }
                acc.dc[0] = sp[0][0][0] + sp[1][0][0];               // lane 0: DC = sum C_dc X_dc (= the cr*xre sum of slot 0)
                acc.ny[0] = sn[0][0][0] + sn[1][0][0];               //         Nyquist = sum C_ny X_ny (= the ci*xim sum of slot 0)
                // park the finished row spectrum; the group's finisher starts when all its rows have arrived
                const int gi = grp0 + r / RG;
                const int gb = gi % NG;
                { TSTART(g); while (bars->accepts[gb] != gi) __nanosleep(128); TSTOP(g, t_g); }     // the finisher has drained the buffer's previous group
                __threadfence_block();
                store_partial<B, NT, 1>(acc, part + (size_t)gb * RG * B, r % RG, 0, lane);
                __syncwarp();
                if (lane == 0) {
                    const int nr = min(RG, rows - (r / RG) * RG);    // rows of this (possibly ragged) group
                    mbar_arrive_n(&bars->gfull[gb], (r % RG == nr - 1) ? (uint32_t)(1 + RG - nr) : 1u);   // its last row stands in for the missing ones
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xempty[k & 1]);         // done reading this segment's X buffer
            chunk0 += rows * nst;
            row0 += rows;
            grp0 += (rows + RG - 1) / RG;
        }
        if (blockIdx.x == 3 && lane == 0 && (m == 0 || m == NM - 1))
            TPRINT("row warp %d: total %lld first_x %lld wait_x %lld wait_chunk %lld mac %lld wait_group %lld rows %d segs %d\n", m, clk() - t_all0, t_first, t_x, t_f, t_mac, t_g,
                   nrows, s1 - s0);
    } else if (warp < NP + NM + NX) {
        // ===================== X warps: rfft(x) of each segment into the other X buffer, one segment ahead of the row warps ======
        const int xw = warp - NP - NM;
        float2* fscr = fscr_all + (size_t)xw * (FC::WARP_SCRATCH_BYTES / 8);
        ProblemDev P0{};
        if (s0 < s1) P0 = problems[segs[s0].g];                          // plan data: constant, fetched while the previous grid drains
        // (keeping the lane's twiddles in registers here, as the finishers do, pushes the kernel over its 96 registers: measured slower)
        pdl_wait();                                                      // x may have been written by the previous grid
        if (s0 < s1) {
            load_x_spectrum<B, NT, NM + NX>(P0, 0, xs_all, tw, fscr, lane, NM + xw);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX)) : "memory");
        }
        for (int s = s0 + 1; s < s1; ++s) {
            const ProblemDev P = problems[segs[s].g];
            const int k = s - s0;
            if (k >= 2) mbar_wait(&bars->xempty[k & 1], (uint32_t)((k - 2) >> 1) & 1u);     // the segment that used this buffer is finished
            load_x_spectrum<B, NT, NX>(P, 0, xs_all + (size_t)(k & 1) * xs_stride, tw, fscr, lane, xw);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xfull[k & 1]);
        }
    } else {
        // ===================== finishers =====================
        const int f = warp - NP - NM - NX;
        float2* fscr = fscr_all + (size_t)(NX + f) * (FC::WARP_SCRATCH_BYTES / 8);
        FftTw<B> ftw;                                                    // this lane's twiddles, once (B = 256: 25 pairs are too many registers)
        if constexpr (B <= 128) fft_tw_init<B>(ftw, tw, lane % FC::T);
        pdl_wait();                                                      // y may still be read by the previous grid
        int gi = 0;
        long long t_all0 = clk(), t_w = 0, t_b = 0; int ngr = 0; (void)t_all0; (void)t_w; (void)t_b; (void)ngr;
        for (int s = s0; s < s1; ++s) {
            const SegWork w = segs[s];
            const ProblemDev P = problems[w.g];
            const int rows = w.j1 - w.j0;
            for (int g0 = 0; g0 < rows; g0 += RG, ++gi) {
                if (gi % NF != f) continue;
                const int gb = gi % NG;
                { TSTART(w); mbar_wait(&bars->gfull[gb], (uint32_t)(gi / NG) & 1u); TSTOP(w, t_w); }
                TSTART(b); ++ngr;
                const int nr = min(RG, rows - g0);
                float* gpart = part + (size_t)gb * RG * B;
                for (int p0 = 0; p0 < nr; p0 += FC::G)
                    finish_pairs<B, NT, 1>(P, w.j0 + g0, 0, gpart + (size_t)p0 * B, p0, min(FC::G, nr - p0), tw, fscr, lane, B <= 128 ? &ftw : nullptr);
                __syncwarp();
                if (lane == 0) { __threadfence_block(); bars->accepts[gb] = gi + NG; }    // the buffer takes group gi + NG next
                TSTOP(b, t_b);
            }
        }
        if (blockIdx.x == 3 && lane == 0) TPRINT("finisher %d: total %lld wait_gfull %lld busy %lld groups %d\n", f, clk() - t_all0, t_w, t_b, ngr);
    }
}

size_t smem_bytes(int B, int spec_bytes, int qmax, int S) {
    const size_t fscr = warp_scratch_bytes(B);
    const size_t stage = (size_t)stage_blocks(B) * B * spec_bytes;
    const int scr_warps = mix_nx(B) + mix_nf(B) + (scratch_aliases_x(B, qmax) ? 0 : NM);
    return 512 + 1024 + (size_t)S * stage + (size_t)mix_ng(B) * RG * B * 4 + (size_t)scr_warps * fscr + (size_t)2 * qmax * B * 4;
}

int pick_stages(int B, int spec_dtype, int nt, int qmax) {
    if (nt != 1 || spec_dtype != BCL_BF16) return 0;                  // fp32 spectra are HBM-bound in bcl_stream_ws.cu already
    const size_t budget = 227 * 1024;
    const size_t fixed = smem_bytes(B, 2, qmax, 0);
    const size_t stage = (size_t)stage_blocks(B) * B * 2;
    if (fixed + (size_t)(NM + 1) * stage > budget) return 0;          // eight rows in flight + one stage of prefetch at least
    int S = (int)((budget - fixed) / stage);
    if (S > MAX_STAGES) S = MAX_STAGES;
    return S / mix_np(B) * mix_np(B);                                 // a stage always belongs to the same producer
}

template <int B>
static cudaError_t launch_t(const GroupPlanHeader& H, const unsigned char* d, cudaStream_t st) {
    auto kern = stream_rw_kernel<B, __nv_bfloat16>;
    const int S = pick_stages(B, H.spec_dtype, H.nt, H.qmax);
    const size_t smem = smem_bytes(B, 2, H.qmax, S);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)H.num_ctas);
    cfg.blockDim = dim3(mix_threads(B));
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

}  // namespace rw

// BCL_DECODE_RW=0 keeps the split-q kernel of bcl_stream_ws.cu for bf16 spectra too (A/B switch for profiling)
bool stream_rw_supported(int B, int spec_dtype, int nt, int qmax) {
    static const bool on = !(getenv("BCL_DECODE_RW") && getenv("BCL_DECODE_RW")[0] == '0');
    return on && rw::pick_stages(B, spec_dtype, nt, qmax) >= rw::NM + 1;
}

cudaError_t launch_stream_rw(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st) {
    const unsigned char* d = reinterpret_cast<const unsigned char*>(dev_plan);
    switch (H.B) { case 64: return rw::launch_t<64>(H, d, st); case 128: return rw::launch_t<128>(H, d, st); default: return rw::launch_t<256>(H, d, st); }
}

}  // namespace bcl
