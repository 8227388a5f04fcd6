// bcl_stream_ws.cu -- warp-specialised grouped stream kernel (the decode hot path on B200).
//
// One persistent CTA per SM walks its byte-balanced list of segments (problem, output block-rows [j0,j1)).  In the
// STREAM layout the spectrum of a segment is ONE contiguous byte range, so the CTA's whole HBM stream is a sequence
// of fixed-size chunks:
//
//   warp 0            TMA producer: cp.async.bulk (global -> shared, mbarrier complete_tx) of chunk after chunk into a
//                     ring of S stages; runs ahead across segment boundaries, so HBM stays busy while the other warps
//                     transform the next segment's x or finish rows.
//   warps 1..NM       MAC warps: block k of the segment (row k/q, input block k%q) belongs to warp k mod NM; lane l
//                     multiplies its packed reals of C_f[row,i] (from the ring) with X_f[i] (shared) into a per-row
//                     partial spectrum in registers; when its row index advances it parks the partial in a
//                     double-buffered group buffer and signals the finishers.  Before each segment the MAC warps
//                     compute rfft(x) of the segment's problem into shared memory.
//   warps NM+1..+NF   finishers: wait for a group of RG rows, sum the NM partials, inverse real FFT, 1/B scale,
//                     round to x.dtype, + bias, store y (reference epilogue src/patch_llama_fft.py:263-271).
//
// All hand-offs are mbarriers (full/empty per ring stage, gfull/gempty per group buffer); the only block-wide
// barrier is the one after setup.  Replaces steps 2-9 of the reference forward for up to hundreds of independent
// projections per launch.
#include <cstdlib>
#include <cstdio>
#include "bcl_stream_parts.cuh"

namespace bcl {
namespace ws {

#ifndef BCL_WS_NM
#define BCL_WS_NM 8
#endif
#ifndef BCL_WS_NF
#define BCL_WS_NF 3
#endif
#ifndef BCL_WS_STAGE_KB
#define BCL_WS_STAGE_KB 16
#endif
constexpr int NM = BCL_WS_NM;                // MAC warps
constexpr int NF = BCL_WS_NF;                // finisher warps == group buffers (finisher f owns buffer f, so it sees every phase of it)
#ifndef BCL_WS_NX
#define BCL_WS_NX 4
#endif
constexpr int NX = BCL_WS_NX;                // x-spectrum warps
constexpr int THREADS = 32 * (1 + NM + NX + NF);
#ifndef BCL_WS_MAX_STAGES
#define BCL_WS_MAX_STAGES 16
#endif
constexpr int MAX_STAGES = BCL_WS_MAX_STAGES;
constexpr int PAIRS = 4;                     // rows per group buffer (one finisher warp inverse-transforms up to G = 32/T rows at once)
// blocks per ring stage: 32 blocks of B=128 (16 KB fp32), 16 of B=256, 64 of B=64 -- always BPS = CB/NM blocks per MAC warp.
// 16-bit spectra: twice the blocks, so that a stage is 16 KB for both dtypes: with 8 KB stages the MAC warps' per-stage hand-offs
// (full wait, release) -- not HBM -- bound the kernel.  A stage then holds TWO block-rows of a q <= CB/2 segment (4096-wide inputs).
__host__ __device__ constexpr int stage_blocks(int B, int spec_bytes = 4) { return (spec_bytes == 2 ? 8192 : 4096) / B; }

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ inline void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
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
    while (!mbar_try(addr, parity)) __nanosleep(32);                // back off: waiting warps must not eat the issue slots of working ones
}
__device__ inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ inline void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ inline void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifdef BCL_WS_TRACE
// Per-CTA timeline of three launches in steady state (launch 200-202 of the process): %globaltimer at CTA start, at the release of
// griddepcontrol.wait and at the end of each finisher; read by tools/ws_trace.py.
__device__ int g_trace_count[1024];
__device__ inline unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ inline int smid() { int v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
#endif
#ifdef BCL_WS_TIMING
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

// The first packed pair of a lane chunk is (DC, Nyquist) in lane 0 -- two real x real products, Y_dc = sum C_dc X_dc and Y_ny = sum C_ny
// X_ny -- and an ordinary complex bin (Re, Im) in every other lane.  The MAC loop keeps the two products of that pair's real part in
// separate sums (dc = sum cr xre, ny = sum ci xim): lane 0 needs exactly those, the other lanes need their difference.  That saves the
// two extra FMAs per block every lane used to spend on lane 0's behalf (10 -> 8 per block at B = 128) and halves the dependent chain.
template <typename Acc> __device__ __forceinline__ void park_prepare(Acc& acc) { acc.a[0][0][0] = acc.dc[0] - acc.ny[0]; }

struct Barriers {
    uint64_t full[MAX_STAGES];      // ring stage filled by TMA            (1 arrival + tx bytes)
    uint64_t empty[MAX_STAGES];     // ring stage released by the MAC warps (NM arrivals)
    uint64_t gfull[NF];             // group buffer complete                (NM arrivals)
    uint64_t gempty[NF];            // group buffer drained by its finisher (1 arrival)
    uint64_t xfull[2];              // x spectrum of a segment ready in shared memory     (NX arrivals; slot 0 used)
    uint64_t xempty[2];             // every MAC warp copied its slice into registers      (NM arrivals)
    int dq_n;                       // dynamic items published by the producer so far (release / acquire)
};
// Dynamic tail (round 2): the last part of a plan's rows is not assigned to CTAs but cut into small items that the producers draw from
// a ticket counter in the device plan, so a CTA that finishes its static share early takes more of them and all CTAs end together
// (per-CTA finish times of equal-byte static shares spread over ~6 us: the HBM system does not serve 148 streams evenly).  The producer
// thread publishes every item it draws in a small shared-memory queue that the other roles follow; g < 0 ends the CTA.
// Every item switches its CTA to another problem, and an rfft of that problem's input per item (~3 us on the four X warps) would cost
// more than the item's stream (4 rows ~1.8 us).  So the input spectrum of every problem of the tail is computed ONCE per launch -- by
// the X warps of CTA (pool index mod grid), right after start-up, when they are idle -- into a cache in the device plan, marked ready
// with the launch sequence number; an item's CTA copies it from L2 instead (16 KB, one round of loads).
// An item carries everything its CTA's producer needs (one load, issued one item ahead: the chain ticket -> item -> problem table would
// otherwise stall the producer for two L2 round trips per item, longer than the ring covers).
struct DynItem { int g, j0, j1, pi; int q, pad; const void* rowp; };       // pi: pool index of the problem; rowp: spectrum row j0
static_assert(sizeof(DynItem) == 32, "DynItem is two 16-byte loads");
struct DynArgs {
    const DynItem* items;           // [n]
    int n;                          // 0: no dynamic tail
    int K;                          // problems in the tail
    int* ctr;                       // [0] ticket counter, [1] CTAs done, [2] launch sequence number, [4 + pi] ready marks
    const int* pool_g;              // [K] problem index
    float* xcache;                  // [K][xstride] packed input spectra
    int xstride;                    // floats
};
__device__ inline int ld_acquire_gpu(const int* p) { int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ inline void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
constexpr int DQ = 32;              // queue slots: > ring stages + group buffers + 2 items can be in flight between producer and finishers
constexpr int HEADER_BYTES = 1536;  // Barriers (512) + queue (DQ * 32)
__device__ inline void st_release_shared(int* p, int v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ inline int ld_acquire_shared(const int* p) { int v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }

// RS = ring stages per block-row = ceil(qmax / CB), KX = x blocks per MAC warp (>= ceil(qmax / NM)): compile-time so that each MAC
// warp's slice of the x spectrum (KX blocks) can live in registers and the per-stage multiply is straight-line code.
// rfft(x) of a problem by the NX x-spectrum warps, out of line: the dynamic-tail kernel calls it from two places, and a third inlined
// copy of the transform pushes the kernel past what the instruction cache holds.
template <int B>
__device__ __noinline__ void x_warps_rfft(const ProblemDev* P, float* dst, const float2* tw, float2* fscr, int lane, int xw) {
    load_x_spectrum<B, 1, NX>(*P, 0, dst, tw, fscr, lane, xw);
}

// DYN: the plan has a dynamic tail.  A separate instantiation, so that plans without one run exactly the code they ran before (the
// tail's paths make the kernel 60 % larger, and measurably slower even when they are never taken).
template <int B, typename SpecT, int RS, int KX, bool DYN>
__global__ void __launch_bounds__(THREADS, 1) stream_ws_kernel(const ProblemDev* __restrict__ problems, const int* __restrict__ cta_seg_ptr,
                                                              const SegWork* __restrict__ segs, int S, int qmax, int dbg, const DynArgs dyn) {
    using LM = LaneMap<B>;
    using FC = FftCfg<B>;
    constexpr int NT = 1;
    constexpr int CH = LM::CH, CW = LM::CW;
    constexpr int RG = PAIRS;                                        // rows per group buffer
    constexpr int BLOCK_BYTES = B * (int)sizeof(SpecT);
    constexpr int CB = stage_blocks(B, (int)sizeof(SpecT));          // blocks per ring stage
    constexpr int BPS = CB / NM;                                     // blocks per MAC warp per stage
    constexpr int HB = BPS / 2 > 0 ? BPS / 2 : 1;                    // ... per row when a stage holds two rows
    constexpr int STAGE_BYTES = CB * BLOCK_BYTES;
    constexpr int NW = CW * (int)sizeof(SpecT) / 4;                  // 32-bit words per lane chunk of a spectrum block
    static_assert(CB % NM == 0, "stage must split evenly over the MAC warps");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    Barriers* bars = reinterpret_cast<Barriers*>(smem_raw);
    int4* dq = reinterpret_cast<int4*>(smem_raw + 512);
    float2* tw = reinterpret_cast<float2*>(smem_raw + HEADER_BYTES);
    unsigned char* ring = smem_raw + HEADER_BYTES + 1024;            // tw needs B/2*8 <= 1024 bytes
    float* part = reinterpret_cast<float*>(ring + (size_t)S * STAGE_BYTES);                         // [NF][PAIRS][NM][B]
    float2* fscr_all = reinterpret_cast<float2*>(part + NF * PAIRS * NM * B);                       // FFT scratch of the NX + NF warps
    float* xs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(fscr_all) + (size_t)(NX + NF) * FC::WARP_SCRATCH_BYTES);   // [qmax][B]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], NM); }
        for (int g = 0; g < NF; ++g) { mbar_init(&bars->gfull[g], NM); mbar_init(&bars->gempty[g], 1); }
        mbar_init(&bars->xfull[0], NX); mbar_init(&bars->xempty[0], NM);
        bars->dq_n = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fill_twiddles<B>(tw, threadIdx.x, THREADS);
    __syncthreads();
    // Programmatic dependent launch: let the NEXT grid in the stream start on every SM this grid leaves (no launch bubble,
    // early finishers pick up the next step at once).  The plan, the problem table and the spectra are constants of the
    // stream; activations are not: the warps that read x or write y call pdl_wait() first (full completion + visibility
    // of the previous grid).
    pdl_launch_dependents();

    const int s0 = cta_seg_ptr[blockIdx.x], s1 = cta_seg_ptr[blockIdx.x + 1];
    const int ns = s1 - s0;
    const int n_dyn = DYN ? dyn.n : 0;
    int* __restrict__ dyn_ctr = dyn.ctr;
    // segment k of this CTA: its static share first, then the dynamic items its producer drew (false: no more work)
    auto fetch = [&](int k, SegWork& w, int& q) -> bool {
        if (k < ns) { w = segs[s0 + k]; q = problems[w.g].q; return true; }
        if (n_dyn == 0) return false;
        const int d = k - ns;
        while (ld_acquire_shared(&bars->dq_n) <= d) __nanosleep(32);
        const int4 e = dq[2 * (d % DQ)];
        w.g = e.x; w.j0 = e.y; w.j1 = e.z; w.pad = e.w;
        q = dq[2 * (d % DQ) + 1].x;
        return e.x >= 0;
    };
    long long t_a = 0, t_b = 0, t_c = 0;
    const long long t_all0 = clk();
    (void)t_a; (void)t_b; (void)t_c; (void)t_all0; (void)dbg; (void)qmax;

    if (warp == 0) {
        // ===================== TMA producer: one bulk copy per (row, stage-of-row) =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0; int n = 0;
            bool have_t = false; int t = 0;
            // first ticket of the dynamic tail: drawn once the ring has gone round (the MAC warps have consumed a stage, so the previous
            // grid is complete and has reset the counter -- the wait returns at once), its latency hidden behind the static share
            auto first_ticket = [&]() { pdl_wait(); t = atomicAdd(dyn_ctr, 1); have_t = true; };
            auto stream_rows = [&](const int q, const unsigned char* rowp, const int j0, const int j1) {
                if (BPS >= 2 && 2 * q <= CB) {                          // narrow rows: two block-rows per stage (they are contiguous)
                    for (int r = j0; r < j1; r += 2, rowp += (size_t)2 * q * BLOCK_BYTES) {
                        { TSTART(w); mbar_wait(&bars->empty[stage], ph ^ 1u); TSTOP(w, t_a); }
                        const uint32_t bytes = (uint32_t)min(2, j1 - r) * (uint32_t)q * BLOCK_BYTES;
                        mbar_arrive_expect_tx(&bars->full[stage], bytes);
                        bulk_g2s(ring + (size_t)stage * STAGE_BYTES, rowp, bytes, &bars->full[stage]);
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                        ++n;
                        if (n_dyn > 0 && !have_t && n > S) first_ticket();
                    }
                    return;
                }
                for (int r = j0; r < j1; ++r, rowp += (size_t)q * BLOCK_BYTES) {
#pragma unroll
                    for (int cs = 0; cs < RS; ++cs) {
                        if (cs * CB < q) {
                            { TSTART(w); mbar_wait(&bars->empty[stage], ph ^ 1u); TSTOP(w, t_a); }
                            const uint32_t bytes = (uint32_t)min(CB, q - cs * CB) * BLOCK_BYTES;
                            mbar_arrive_expect_tx(&bars->full[stage], bytes);
                            bulk_g2s(ring + (size_t)stage * STAGE_BYTES, rowp + (size_t)cs * STAGE_BYTES, bytes, &bars->full[stage]);
                            if (++stage == S) { stage = 0; ph ^= 1u; }
                            ++n;
                        }
                    }
                    if (n_dyn > 0 && !have_t && n > S) first_ticket();
                }
            };
            if constexpr (!DYN) {
                for (int s = s0; s < s1; ++s) {
                    const SegWork w = segs[s];
                    const ProblemDev* P = problems + w.g;
                    const int q = P->q;
                    stream_rows(q, reinterpret_cast<const unsigned char*>(P->spec) + (size_t)w.j0 * q * BLOCK_BYTES, w.j0, w.j1);
                }
            } else {
                // Software pipeline over the tickets: while item d streams, the load of item d + 1 and the draw of ticket d + 2 are in
                // flight.  Every CTA draws exactly one ticket past the end; the X warps, who follow the queue to that entry, count the
                // CTA as done.
                auto load_item = [&](int tk) {
                    const int4* p = reinterpret_cast<const int4*>(dyn.items + tk);
                    const int4 a = __ldg(p), b = __ldg(p + 1);
                    DynItem it;
                    it.g = a.x; it.j0 = a.y; it.j1 = a.z; it.pi = a.w; it.q = b.x; it.pad = 0;
                    it.rowp = reinterpret_cast<const void*>(((unsigned long long)(unsigned)b.w << 32) | (unsigned)b.z);
                    return it;
                };
                // (one loop -- one copy of the streaming code -- over the static segments, k < ns, and then the tickets)
                bool live = false;
                DynItem it{};
                int tn = 0;
                for (int k = 0;; ++k) {
                    if (k < ns) {
                        const SegWork w = segs[s0 + k];
                        const ProblemDev* P = problems + w.g;
                        it.q = P->q; it.j0 = w.j0; it.j1 = w.j1;
                        it.rowp = reinterpret_cast<const unsigned char*>(P->spec) + (size_t)w.j0 * it.q * BLOCK_BYTES;
                    } else {
                        if (k == ns) {
                            if (!have_t) first_ticket();
                            live = t < n_dyn;
                            if (live) it = load_item(t);
                            tn = live ? atomicAdd(dyn_ctr, 1) : 0;
                        }
                        const int d = k - ns;
                        dq[2 * (d % DQ)] = live ? make_int4(it.g, it.j0, it.j1, it.pi) : make_int4(-1, 0, 0, 0);
                        dq[2 * (d % DQ) + 1] = make_int4(it.q, 0, 0, 0);
                        st_release_shared(&bars->dq_n, d + 1);
                        if (!live) break;
                    }
                    DynItem itn{};
                    bool live_n = false;
                    int tnn = 0;
                    if (k >= ns) {
                        live_n = tn < n_dyn;
                        if (live_n) itn = load_item(tn);
                        tnn = live_n ? atomicAdd(dyn_ctr, 1) : 0;
                    }
                    stream_rows(it.q, reinterpret_cast<const unsigned char*>(it.rowp), it.j0, it.j1);
                    if (k >= ns) { it = itn; live = live_n; tn = tnn; }
                }
            }
            if (blockIdx.x == 3) TPRINT("producer: total %lld wait_empty %lld chunks %d\n", clk() - t_all0, t_a, n);
        }
    } else if (warp <= NM) {
        // ===================== MAC warps =====================
        const int m = warp - 1;
        int stage = 0; uint32_t ph = 0;
        int gi = 0;
        RowAcc<B, NT> acc;
        SegWork w; int q = 0;
        const bool any = fetch(0, w, q);                                 // (no static share: waits for the producer's first ticket)
        if (any) {
            // First segment: nothing to multiply yet, so the MAC warps help the X warps with rfft(x) (all NM + NX warps, one
            // pass for q <= 48 instead of three): the start-up latency during which the ring is full and HBM idles.
            // Scratch: the (still unused) group buffers.
            const ProblemDev P0 = problems[w.g];                         // plan data: constant, fetched while the previous grid drains
            if (!(dbg & 1)) pdl_wait();                                  // x may have been written by the previous grid (bit 0 of the plan flags: it was not)
            TSTART(x0);
            load_x_spectrum<B, NT, NM + NX>(P0, 0, xs, tw, reinterpret_cast<float2*>(part) + (size_t)m * (FC::WARP_SCRATCH_BYTES / 8), lane, m);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX)) : "memory");
            TSTOP(x0, t_a);
        }
        for (int k = 0; any && fetch(k, w, q); ++k) {
            const int rows = w.j1 - w.j0;
            // my slice of this segment's x spectrum -> registers (blocks i = cs*CB + m + NM*u), then free the X buffer at once
            float xr[RS][BPS][CH][CW];                                // (entries cs * BPS + u >= KX are never touched: no registers)
            if (k > 0) { TSTART(x); mbar_wait(&bars->xfull[0], (uint32_t)(k - 1) & 1u); TSTOP(x, t_a); }
#pragma unroll
            for (int cs = 0; cs < RS; ++cs)
#pragma unroll
                for (int u = 0; u < BPS; ++u) {
                    if (cs * BPS + u >= KX) continue;
                    const int i = cs * CB + m + NM * u;
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                        if (i < q) {
                            const float* xp = xs + (size_t)i * B + LM::offset(lane, ch);
                            if constexpr (CW == 4) { float4 t = *reinterpret_cast<const float4*>(xp); xr[cs][u][ch][0] = t.x; xr[cs][u][ch][1] = t.y; xr[cs][u][ch][2] = t.z; xr[cs][u][ch][3] = t.w; }
                            else { float2 t = *reinterpret_cast<const float2*>(xp); xr[cs][u][ch][0] = t.x; xr[cs][u][ch][1] = t.y; }
                        } else {
#pragma unroll
                            for (int e = 0; e < CW; ++e) xr[cs][u][ch][e] = 0.f;
                        }
                    }
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xempty[0]);
            if (BPS >= 2 && 2 * q <= CB) {
                // ---- two block-rows per stage: rows r and r + 1 (same group: RG is even), x blocks u < HB, independent accumulators
                for (int r = 0; r < rows; r += 2) {
                    const bool two = r + 1 < rows;
                    RowAcc<B, NT> acc1;
                    acc.clear(); acc1.clear();
                    { TSTART(f); mbar_wait(&bars->full[stage], ph); TSTOP(f, t_b); }
                    const SpecT* cbuf = reinterpret_cast<const SpecT*>(ring + (size_t)stage * STAGE_BYTES) + (size_t)m * B;
                    RawChunk<SpecT, CW> raw[2][HB][CH];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int u = 0; u < HB; ++u) {
                            if (u < KX && m + NM * u < q && (rr == 0 || two)) {
#pragma unroll
                                for (int ch = 0; ch < CH; ++ch) {
                                    const uint32_t* src = reinterpret_cast<const uint32_t*>(cbuf + (size_t)(rr * q + NM * u) * B + LM::offset(lane, ch));
                                    if constexpr (NW == 4) { uint4 t = *reinterpret_cast<const uint4*>(src); raw[rr][u][ch].w[0] = t.x; raw[rr][u][ch].w[1] = t.y; raw[rr][u][ch].w[2] = t.z; raw[rr][u][ch].w[3] = t.w; }
                                    else if constexpr (NW == 2) { uint2 t = *reinterpret_cast<const uint2*>(src); raw[rr][u][ch].w[0] = t.x; raw[rr][u][ch].w[1] = t.y; }
                                    else { raw[rr][u][ch].w[0] = *src; }
                                }
                            } else {
#pragma unroll
                                for (int ch = 0; ch < CH; ++ch)
#pragma unroll
                                    for (int e = 0; e < NW; ++e) raw[rr][u][ch].w[e] = 0u;
                            }
                        }
#pragma unroll
                    for (int u = 0; u < HB; ++u) {
                        if (u >= KX) continue;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            RowAcc<B, NT>& ac = rr == 0 ? acc : acc1;
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                float cv[CW];
                                decode_raw<SpecT, CW>(raw[rr][u][ch], cv);
#pragma unroll
                                for (int e = 0; e < CW; e += 2) {
                                    const float cr = cv[e], ci = cv[e + 1], xre = xr[0][u][ch][e], xim = xr[0][u][ch][e + 1];
                                    if (ch == 0 && e == 0) {         // first pair: the two products of Re kept apart (see park_prepare)
                                        ac.dc[0] = fmaf(cr, xre, ac.dc[0]);
                                        ac.ny[0] = fmaf(ci, xim, ac.ny[0]);
                                    } else {
                                        ac.a[0][ch][e] = fmaf(cr, xre, ac.a[0][ch][e]);
                                        ac.a[0][ch][e] = fmaf(-ci, xim, ac.a[0][ch][e]);
                                    }
                                    ac.a[0][ch][e + 1] = fmaf(cr, xim, ac.a[0][ch][e + 1]);
                                    ac.a[0][ch][e + 1] = fmaf(ci, xre, ac.a[0][ch][e + 1]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->empty[stage]);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                    const int slot = r % RG;
                    const int gb = gi % NF;
                    if (slot == 0) { TSTART(g); mbar_wait(&bars->gempty[gb], ((uint32_t)(gi / NF) & 1u) ^ 1u); TSTOP(g, t_c); }
                    park_prepare(acc); park_prepare(acc1);
                    store_partial<B, NT, NM>(acc, part + (size_t)gb * PAIRS * NM * B, slot, m, lane);
                    if (two) store_partial<B, NT, NM>(acc1, part + (size_t)gb * PAIRS * NM * B, slot + 1, m, lane);
                    if (slot + (two ? 1 : 0) == RG - 1 || r + 2 >= rows) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->gfull[gb]);
                        ++gi;
                    }
                }
                continue;
            }
            for (int r = 0; r < rows; ++r) {
                acc.clear();
#pragma unroll
                for (int cs = 0; cs < RS; ++cs) {
                    if (cs * CB < q) {
                        { TSTART(f); mbar_wait(&bars->full[stage], ph); TSTOP(f, t_b); }
                        const SpecT* cbuf = reinterpret_cast<const SpecT*>(ring + (size_t)stage * STAGE_BYTES) + (size_t)m * B;
                        RawChunk<SpecT, CW> raw[BPS][CH];
#pragma unroll
                        for (int u = 0; u < BPS; ++u) {               // all shared loads first: one latency, no branches
                            if (cs * BPS + u < KX && cs * CB + m + NM * u < q) {
#pragma unroll
                                for (int ch = 0; ch < CH; ++ch) {
                                    const uint32_t* src = reinterpret_cast<const uint32_t*>(cbuf + (size_t)(NM * u) * B + LM::offset(lane, ch));
                                    if constexpr (NW == 4) { uint4 t = *reinterpret_cast<const uint4*>(src); raw[u][ch].w[0] = t.x; raw[u][ch].w[1] = t.y; raw[u][ch].w[2] = t.z; raw[u][ch].w[3] = t.w; }
                                    else if constexpr (NW == 2) { uint2 t = *reinterpret_cast<const uint2*>(src); raw[u][ch].w[0] = t.x; raw[u][ch].w[1] = t.y; }
                                    else { raw[u][ch].w[0] = *src; }
                                }
                            } else {
#pragma unroll
                                for (int ch = 0; ch < CH; ++ch)
#pragma unroll
                                    for (int e = 0; e < NW; ++e) raw[u][ch].w[e] = 0u;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < BPS; ++u) {
                            if (cs * BPS + u >= KX) continue;
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                float cv[CW];
                                decode_raw<SpecT, CW>(raw[u][ch], cv);
#pragma unroll
                                for (int e = 0; e < CW; e += 2) {
                                    const float cr = cv[e], ci = cv[e + 1], xre = xr[cs][u][ch][e], xim = xr[cs][u][ch][e + 1];
                                    if (ch == 0 && e == 0) {         // first pair: the two products of Re kept apart (see park_prepare)
                                        acc.dc[0] = fmaf(cr, xre, acc.dc[0]);
                                        acc.ny[0] = fmaf(ci, xim, acc.ny[0]);
                                    } else {
                                        acc.a[0][ch][e] = fmaf(cr, xre, acc.a[0][ch][e]);
                                        acc.a[0][ch][e] = fmaf(-ci, xim, acc.a[0][ch][e]);
                                    }
                                    acc.a[0][ch][e + 1] = fmaf(cr, xim, acc.a[0][ch][e + 1]);
                                    acc.a[0][ch][e + 1] = fmaf(ci, xre, acc.a[0][ch][e + 1]);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->empty[stage]);   // this warp no longer reads the stage
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
                // park the row's partial spectrum in the group buffer; the last row of a group signals the finisher
                const int slot = r % RG;
                const int gb = gi % NF;
                if (slot == 0) { TSTART(g); mbar_wait(&bars->gempty[gb], ((uint32_t)(gi / NF) & 1u) ^ 1u); TSTOP(g, t_c); }
                park_prepare(acc);
                store_partial<B, NT, NM>(acc, part + (size_t)gb * PAIRS * NM * B, slot, m, lane);
                if (slot == RG - 1 || r == rows - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->gfull[gb]);
                    ++gi;
                }
            }
        }
        if (blockIdx.x == 3 && lane == 0 && (m == 0 || m == NM - 1))
            TPRINT("mac %d: total %lld wait_x %lld wait_full %lld wait_gempty %lld segs %d\n", m, clk() - t_all0, t_a, t_b, t_c, s1 - s0);
    } else if (warp <= NM + NX) {
        // ===================== X warps: rfft(x) of each segment into the X buffer, one segment ahead of the MAC warps ======
        const int xw = warp - 1 - NM;
        float2* fscr = fscr_all + (size_t)xw * (FC::WARP_SCRATCH_BYTES / 8);
        ProblemDev P0{};
        SegWork w; int qk = 0;
        const bool any = fetch(0, w, qk);
        if (any) P0 = problems[w.g];                                     // plan data: constant, fetched while the previous grid drains
        if (!(dbg & 1)) pdl_wait();                                      // x may have been written by the previous grid
        // dynamic tail: this launch's mark = the sequence number the last CTA of the previous launch left + 1.  Read before anything of
        // this CTA can count as done (below), so no CTA can see the number this launch will leave.
        int mark = 0;
        if (n_dyn > 0) mark = __ldcg(dyn_ctr + 2) + 1;
        FftTw<B> xtw;                                                    // this lane's twiddles, once (static plans, B <= 128)
        if constexpr (B <= 128 && !DYN) fft_tw_init<B>(xtw, tw, lane % FC::T);
        const FftTw<B>* xtwp = (B <= 128 && !DYN) ? &xtw : nullptr;
        if (any) {                                                       // first segment: together with the MAC warps (see there)
            TSTART(y0);
            load_x_spectrum<B, NT, NM + NX>(P0, 0, xs, tw, fscr, lane, NM + xw, xtwp);
            asm volatile("bar.sync 1, %0;" ::"r"(32 * (NM + NX)) : "memory");
            TSTOP(y0, t_b);
        }
        // dynamic tail: this CTA's share of the tail's input spectra (computed when the X warps have nothing else to do: at once if the
        // static share is a single segment, else behind the second segment's transform)
        auto publish = [&]() {
            for (int pi = blockIdx.x; pi < dyn.K; pi += gridDim.x) {
                x_warps_rfft<B>(problems + dyn.pool_g[pi], dyn.xcache + (size_t)pi * dyn.xstride, tw, fscr, lane, xw);
                __threadfence();
                asm volatile("bar.sync 2, %0;" ::"r"(32 * NX) : "memory");
                if (xw == 0 && lane == 0) st_release_gpu(dyn_ctr + 4 + pi, mark);
            }
        };
        bool published = n_dyn == 0, all_ready = false;
        if (!published && ns <= 1) { publish(); published = true; }
        for (int k = 1; any && fetch(k, w, qk); ++k) {
            // An item of the tail: its input spectrum is in the cache.  The first 8 x 16 bytes per lane (all of a 4096-wide input) are
            // loaded BEFORE waiting for the X buffer, so that behind the wait only the shared-memory stores remain: the item before it
            // lasts ~1.5 us, an L2 round trip behind the wait would not fit.
            float4 pre[8];
            const float4* csrc = nullptr;
            int n4 = 0;
            if (DYN && k >= ns) {
                if (!all_ready) {                                        // first item: every spectrum of the tail was marked ready long ago
                    for (int pi = lane; pi < dyn.K; pi += 32)
                        while (ld_acquire_gpu(dyn_ctr + 4 + pi) != mark) __nanosleep(64);
                    __syncwarp();
                    all_ready = true;
                }
                csrc = reinterpret_cast<const float4*>(dyn.xcache + (size_t)w.pad * dyn.xstride);
                n4 = qk * (B / 4);
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int i = xw * 32 + lane + u * 32 * NX; if (i < n4) pre[u] = __ldcg(csrc + i); }
            }
            { TSTART(x); mbar_wait(&bars->xempty[0], ((uint32_t)k & 1u) ^ 1u); TSTOP(x, t_a); }   // MAC warps copied the previous X out
            if (k < ns) {
                TSTART(y);
                if constexpr (DYN) x_warps_rfft<B>(problems + w.g, xs, tw, fscr, lane, xw);
                else { const ProblemDev P = problems[w.g]; load_x_spectrum<B, NT, NX>(P, 0, xs, tw, fscr, lane, xw, xtwp); }
                TSTOP(y, t_b);
            } else if (DYN) {
                float4* dst = reinterpret_cast<float4*>(xs);
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int i = xw * 32 + lane + u * 32 * NX; if (i < n4) dst[i] = pre[u]; }
                for (int i0 = xw * 32 + lane + 8 * 32 * NX; i0 < n4; i0 += 8 * 32 * NX) {      // (wider inputs: the rest, behind the wait)
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (i0 + u * 32 * NX < n4) v[u] = __ldcg(csrc + i0 + u * 32 * NX);
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (i0 + u * 32 * NX < n4) dst[i0 + u * 32 * NX] = v[u];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xfull[0]);
            if (!published) { publish(); published = true; }
        }
        if (!published) publish();
        if (n_dyn > 0) {
            // This CTA has drawn its one ticket past the end (the X warps followed the queue to that entry), read its mark and published
            // its spectra.  The last CTA of the grid to get here re-arms the counters for the next launch, which touches them only after
            // griddepcontrol.wait, i.e. after this grid has completed.
            asm volatile("bar.sync 2, %0;" ::"r"(32 * NX) : "memory");
            if (xw == 0 && lane == 0 && atomicAdd(dyn_ctr + 1, 1) == (int)gridDim.x - 1) { atomicExch(dyn_ctr, 0); atomicExch(dyn_ctr + 1, 0); atomicAdd(dyn_ctr + 2, 1); }
        }
        if (blockIdx.x == 3 && lane == 0) TPRINT("xw %d: total %lld wait_xempty %lld fft %lld\n", xw, clk() - t_all0, t_a, t_b);
    } else {
        // ===================== finishers: one warp per group, alternating =====================
        const int f = warp - 1 - NM - NX;
        float2* fscr = fscr_all + (size_t)(NX + f) * (FC::WARP_SCRATCH_BYTES / 8);
#ifdef BCL_WS_TRACE
        const unsigned long long tr0 = gtime();
#endif
        if (!(dbg & 1)) pdl_wait();                                      // y may still be read by the previous grid
#ifdef BCL_WS_TRACE
        const unsigned long long tr1 = gtime();
        int trn = 0;
        if (lane == 0) { trn = g_trace_count[blockIdx.x * NF + f]; g_trace_count[blockIdx.x * NF + f] = trn + 1; }
#endif
        int gi = 0;
        SegWork w; int qf = 0;
        FftTw<B> ftw;                                                    // this lane's twiddles, once (B = 256: too many registers)
        if constexpr (B <= 128) fft_tw_init<B>(ftw, tw, lane % FC::T);
        for (int k = 0; fetch(k, w, qf); ++k) {
            const ProblemDev P = problems[w.g];
            const int rows = w.j1 - w.j0;
            for (int g0 = 0; g0 < rows; g0 += RG, ++gi) {
                if (gi % NF != f) continue;
                const int gb = f;
                { TSTART(g); mbar_wait(&bars->gfull[gb], (uint32_t)(gi / NF) & 1u); TSTOP(g, t_a); }
                const int nr = min(RG, rows - g0);
                TSTART(fin);
                float* gpart = part + (size_t)gb * PAIRS * NM * B;
                for (int p0 = 0; p0 < nr; p0 += FC::G)
                    finish_pairs<B, NT, NM>(P, w.j0 + g0, 0, gpart + (size_t)p0 * NM * B, p0, min(FC::G, nr - p0), tw, fscr, lane, B <= 128 ? &ftw : nullptr);
                TSTOP(fin, t_b);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->gempty[gb]);
            }
        }
        if (blockIdx.x == 3 && lane == 0) TPRINT("fin %d: total %lld wait_gfull %lld finish %lld groups %d\n", f, clk() - t_all0, t_a, t_b, gi);
#ifdef BCL_WS_TRACE
        if (lane == 0 && trn >= 200 && trn < 203)
            printf("TL %d %d %d %d %llu %llu %llu %d\n", trn, (int)blockIdx.x, smid(), f, tr0, tr1, gtime(), s1 - s0);
#endif
    }
}

size_t smem_bytes(int B, int spec_bytes, int qmax, int S) {
    const size_t fscr = B == 64 ? FftCfg<64>::WARP_SCRATCH_BYTES : (B == 128 ? FftCfg<128>::WARP_SCRATCH_BYTES : FftCfg<256>::WARP_SCRATCH_BYTES);
    const size_t stage = (size_t)stage_blocks(B, spec_bytes) * B * spec_bytes;
    return HEADER_BYTES + 1024 + (size_t)S * stage + (size_t)NF * PAIRS * NM * B * 4 + (size_t)(NX + NF) * fscr + (size_t)qmax * B * 4;
}

// Number of ring stages that fit next to the fixed buffers (0 = this shape does not fit the warp-specialised kernel).
int pick_stages(int B, int spec_dtype, int nt, int qmax) {
    if (nt != 1) return 0;                                           // decode path: one token per problem
    const int sb = spec_dtype == BCL_F32 ? 4 : 2;
    // x slice must fit the MAC warps' registers: RS <= 4 and KX <= 16 blocks of 4 (B = 128) or 8 (B = 256) reals per lane; a block of
    // B = 64 is 2 reals per lane, so 32 of them are the same 64 registers (11008-wide inputs at B = 64: q = 172, KX = 22)
    if (qmax > 4 * stage_blocks(B, sb) || (qmax + NM - 1) / NM > (B == 64 ? 32 : 16)) return 0;
    if (stage_blocks(B, sb) / NM > 4) {                              // wider stages: the instances launch_b has (RS <= min(3, KXMAX / BPS))
        const int bps = stage_blocks(B, sb) / NM, kxmax = B == 64 ? 32 : 16;
        const int rsmax = kxmax / bps > 0 ? (kxmax / bps > 3 ? 3 : kxmax / bps) : 1;
        if ((qmax + stage_blocks(B, sb) - 1) / stage_blocks(B, sb) > rsmax) return 0;
    }
    const size_t budget = 227 * 1024;
    const size_t fixed = smem_bytes(B, sb, qmax, 0);
    const size_t stage = (size_t)stage_blocks(B, sb) * B * sb;
    if (fixed + 3 * stage > budget) return 0;
    int S = (int)((budget - fixed) / stage);
    if (S > MAX_STAGES) S = MAX_STAGES;
    return S;
}

static bool pdl_enabled() {                                             // BCL_PDL=0: A/B switch for profiling only
    static const bool on = !(getenv("BCL_PDL") && getenv("BCL_PDL")[0] == '0');
    return on;
}

template <int B, typename SpecT, int RS, int KX>
static cudaError_t launch_t(const GroupPlanHeader& H, const unsigned char* d, cudaStream_t st) {
    const bool with_tail = H.num_dyn > 0 && !(H.reserved & 1);
    auto kern = with_tail ? stream_ws_kernel<B, SpecT, RS, KX, true> : stream_ws_kernel<B, SpecT, RS, KX, false>;
    const int S = pick_stages(B, H.spec_dtype, H.nt, H.qmax);
    const size_t smem = smem_bytes(B, (int)sizeof(SpecT), H.qmax, S);
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
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    const ProblemDev* probs = reinterpret_cast<const ProblemDev*>(d + H.off_problems);
    const int* cta_ptr = reinterpret_cast<const int*>(d + H.off_cta_ptr);
    const SegWork* segs = reinterpret_cast<const SegWork*>(d + H.off_segs);
    int qmax = H.qmax, dbg = H.reserved, S_ = S;          // `reserved` = plan flags (bit 0: launches of this plan do not depend on the launch before them)
    // dynamic tail: items behind the static segments, ticket + completion counters in the device plan (the one part of it the kernel writes)
    DynArgs dyn{};
    dyn.items = reinterpret_cast<const DynItem*>(segs + H.num_segs);
    dyn.n = (dbg & 1) ? 0 : H.num_dyn;
    dyn.K = H.dyn_problems;
    dyn.ctr = reinterpret_cast<int*>(const_cast<unsigned char*>(d) + H.off_ctr);
    dyn.pool_g = reinterpret_cast<const int*>(d + H.off_ctr) + 4 + H.dyn_problems;
    dyn.xcache = reinterpret_cast<float*>(const_cast<unsigned char*>(d) + H.off_xcache);
    dyn.xstride = H.dyn_xstride;
    void* args[] = {(void*)&probs, (void*)&cta_ptr, (void*)&segs, (void*)&S_, (void*)&qmax, (void*)&dbg, (void*)&dyn};
    e = cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kern), args);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

template <int B, typename SpecT>
static cudaError_t launch_b(const GroupPlanHeader& H, const unsigned char* d, cudaStream_t st) {
    constexpr int CB = stage_blocks(B, (int)sizeof(SpecT)), BPS = CB / NM;
    const int rs = (H.qmax + CB - 1) / CB;
    const int kx = (H.qmax + NM - 1) / NM;                            // x blocks per MAC warp
    if constexpr (BPS <= 4) {                                        // KX = RS * BPS <= 16: one instance per RS
        switch (rs) {
            case 1: return launch_t<B, SpecT, 1, BPS>(H, d, st);
            case 2: return launch_t<B, SpecT, 2, 2 * BPS>(H, d, st);
            case 3: return launch_t<B, SpecT, 3, 3 * BPS>(H, d, st);
            default: return launch_t<B, SpecT, 4, 4 * BPS>(H, d, st);
        }
    } else {                                                         // wider stages: bound the register slice by the plan's qmax (KX <= KXMAX)
        constexpr int KXMAX = B == 64 ? 32 : 16;                     // (pick_stages: the same 64 registers per lane)
        constexpr int RSMAX = KXMAX / BPS > 0 ? (KXMAX / BPS > 3 ? 3 : KXMAX / BPS) : 1;
        if (rs > RSMAX || kx > KXMAX) return cudaErrorInvalidValue;
        if (rs == 1) return kx <= BPS / 2 ? launch_t<B, SpecT, 1, BPS / 2>(H, d, st) : launch_t<B, SpecT, 1, BPS>(H, d, st);
        if constexpr (RSMAX >= 2) {
            if (rs == 2) {
                if (kx <= BPS + BPS / 2) return launch_t<B, SpecT, 2, BPS + BPS / 2>(H, d, st);
                return launch_t<B, SpecT, 2, 2 * BPS>(H, d, st);
            }
        }
        if constexpr (RSMAX >= 3) return launch_t<B, SpecT, 3, 3 * BPS>(H, d, st);      // B = 64, fp32 spectra, 128 < q <= 192
        return cudaErrorInvalidValue;
    }
}

}  // namespace ws

bool stream_ws_supported(int B, int spec_dtype, int nt, int qmax) { return ws::pick_stages(B, spec_dtype, nt, qmax) >= 3; }

cudaError_t launch_stream_ws(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st) {
    const unsigned char* d = reinterpret_cast<const unsigned char*>(dev_plan);
#define BCL_WS(BB) (H.spec_dtype == BCL_F32 ? ws::launch_b<BB, float>(H, d, st) : ws::launch_b<BB, __nv_bfloat16>(H, d, st))
    switch (H.B) { case 64: return BCL_WS(64); case 128: return BCL_WS(128); default: return BCL_WS(256); }
#undef BCL_WS
}

}  // namespace bcl
