// bcl_prefill_wave.cu -- the prefill forward as ONE persistent launch in which EVERY CTA runs all three tensor-core stages,
// interleaved item by item, a fixed number of token tiles apart ("wave" schedule).
//
// reference: src/patch_llama_fft.py:246-271 (rfft -> einsum -> irfft -> cast -> bias), src/bench_all_in_one.py:179-231 (cached spectrum).
//
// Why: three stage launches move Xf and Yf through HBM (three read+write passes, profiles/r1_prefill_notes.md), and the first
// single-launch pipeline (bcl_prefill_fused.cu) split the SMs statically into three role groups, each bound by its own per-item
// latency.  Here the item lists of the three stages
//        A (m, kg)        forward DFT of 4 input blocks of token tile m           x  -> Xf      4 TMEM slots
//        B (m, fg, nch)   frequency GEMM of 4 frequencies x 32 output blocks      Xf -> Yf      2 TMEM slots
//        C (m, j)         inverse DFT + cast + bias of output block j             Yf -> y       1 TMEM slot
// are dealt round-robin to all CTAs in small chunks, and a CTA merges its three lists by the key  tile + lag * stage : it works on
// A of tile v, B of tile v - lag and C of tile v - 2 lag at the same time.  What a B item reads was written `lag` tiles of work
// earlier by other CTAs (still in L2, long complete), so the per-tile completion counters (StageDeps protocol of the fused launch:
// TMA store -> wait_group -> signaller warp: fence + red; consumer: ld.acquire + fence.proxy.async -> TMA load) are almost never
// waited on, every SM stays busy on all three stages, and only x is read from / y written to HBM.  Xf and Yf are RINGS of `ring`
// token tiles (tile m lives in slot m % ring): a slot is overwritten while its lines are still in L2, so the intermediates are
// not written back either; the overwrite waits for the consumer's counter of the tile that held the slot before.
// One unified pipeline per CTA: a shared-memory ring of 32 KB stages (any item type), four 128-column TMEM slots handed out in
// item order (full / empty mbarrier per slot), warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue, warp 10 = signaller.
// The MMAs, operand layouts and epilogues are those of the stage kernels (bcl_prefill_bodies.cuh): the output is bit-identical.
//
// MEASURED (B200, 4096x4096, 16384 tokens, profiles/r2_prefill_wave_notes.md): 155-180 us against 141 us for the three launches, at
// every lag / ring / chunking tried, so it is OPT-IN (BCL_PREFILL_WAVE=1).  Two DFT matrices (64 KB) + store staging leave 3-4
// smem stages of 32 KB for all three item types (the stage kernels have 128-192 KB of ring each), and the per-role cycle counters
// (make WAVE_PROF=1, tools/wave_prof.py) show the MMA issuer waiting for operands 50-65 % of the time while the epilogue warps are
// busy only 25-30 %.  The same counters and the L2-resident stage timings say why removing HBM passes does not pay here: every
// stage runs at the same speed from L2 as from HBM (L2 throughput + TMEM read rate bound them at about the HBM rate).
#include <cstdlib>
#include <climits>
#include "bcl_prefill_bodies.cuh"

namespace bcl {
namespace gemm {

constexpr int WAVE_THREADS = 32 * (SIG_WARP + 1);
constexpr int WAVE_MAX_STAGES = 6;
constexpr uint32_t WAVE_STAGE = 32768;                   // one smem stage: any item type's largest load unit
constexpr int WAVE_SLOTS = 4;                            // TMEM: 4 slots of 128 columns
#ifndef BCL_WAVE_NBUF
#define BCL_WAVE_NBUF 1
#endif
constexpr int WAVE_NBUF = BCL_WAVE_NBUF;                 // staging buffers per epilogue warp (4 KB each)
// BCL_WAVE_ASPLIT=2 (B = 128): a forward-DFT item computes HALF of the spectrum of 4 blocks (2 TMEM slots instead of all 4), so its
// MMAs overlap the epilogue of the item before; x is loaded once per half (the second time from L2).  Measured slower (the shallow
// smem ring is what bounds this kernel, and the second x load competes for it): default 1.
#ifndef BCL_WAVE_ASPLIT
#define BCL_WAVE_ASPLIT 1
#endif
__host__ __device__ constexpr int WAVE_ASPLIT(int B) { return B == 128 ? BCL_WAVE_ASPLIT : 1; }

struct WaveArgs {
    int* status;
    const void* bias;
    int* done;               // [A | B | C] per-tile completion counters, m_tiles ints each, zeroed before the launch
    int act_bf16, stages;
    int m_tiles, ring, lag;
    int KG, FG, NCH, p, k_atoms;
    int per_tile[3];         // items per token tile: KG, FG * NCH, p
    int n_items[3];
    int chunk[3];            // consecutive items a CTA takes at a time
    int target[3];           // counter value of a complete tile: per_tile * epilogue warps
    long long* prof;         // profiling builds (make WAVE_PROF=1) only: per-role wait / busy cycles of the first 8 CTAs
    int immediate;           // few tiles (latency-bound): publish every item as soon as its stores have completed
};

struct WaveBarriers {
    uint64_t full[WAVE_MAX_STAGES], empty[WAVE_MAX_STAGES];
    uint64_t tfull[WAVE_SLOTS], tempty[WAVE_SLOTS];
    uint64_t wfull;
    uint32_t tmem_base;
    uint32_t pad;
};

// The CTA's merged item order.  nxt[t] = next chunk index of type t for this CTA (starts at blockIdx.x, advances by the grid size).
struct WaveCursor {
    int na, nb, nc;
    __device__ __forceinline__ bool next(const WaveArgs& a, int G, int& type, int& i0, int& i1) {
        int best = -1, key = INT_MAX;
        const int fa = na * a.chunk[0], fb = nb * a.chunk[1], fc = nc * a.chunk[2];
        if (fa < a.n_items[0]) { best = 0; key = fa / a.per_tile[0]; }
        if (fb < a.n_items[1]) { const int k = fb / a.per_tile[1] + a.lag; if (k < key) { best = 1; key = k; } }
        if (fc < a.n_items[2]) { const int k = fc / a.per_tile[2] + 2 * a.lag; if (k < key) { best = 2; key = k; } }
        if (best < 0) return false;
        type = best;
        if (best == 0) { i0 = fa; i1 = min(fa + a.chunk[0], a.n_items[0]); na += G; }
        else if (best == 1) { i0 = fb; i1 = min(fb + a.chunk[1], a.n_items[1]); nb += G; }
        else { i0 = fc; i1 = min(fc + a.chunk[2], a.n_items[2]); nc += G; }
        return true;
    }
};

// Completion bookkeeping of an epilogue warp (lane 0).  Consecutive chunks of a CTA belong to different tiles, so the tracker of the
// stage kernels (wait for the previous tile's bulk stores at every tile change) would stall the warp for a full store round trip
// per chunk.  Here a tile's signal is handed to the signaller two tile changes later, when its stores have long completed:
// wait_group<groups committed since> returns at once.  flush() (warp about to idle) publishes everything.
template <int N> struct WaitLe { static __device__ __forceinline__ void run(int n) { if (n >= N) tma_store_wait<N>(); else WaitLe<N - 1>::run(n); } };
template <> struct WaitLe<0> { static __device__ __forceinline__ void run(int) { tma_store_wait<0>(); } };
struct WaveTracker {
    int m0 = -1, c0 = 0, m1 = -1, c1 = 0, g1 = 0;
    static __device__ __forceinline__ void push(uint32_t* sig, int m, int cnt) { sig_push(sig, ((uint32_t)(m + 1) << 8) | (uint32_t)cnt); }
    __device__ __forceinline__ void commit(uint32_t* sig, int m, int groups) {
        if (m == m1) { ++c1; g1 += groups; return; }
        if (m0 >= 0) { WaitLe<12>::run(g1 + groups); push(sig, m0, c0); }
        m0 = m1; c0 = c1; m1 = m; c1 = 1; g1 = groups;
    }
    __device__ __forceinline__ void flush(uint32_t* sig) {
        if (m1 < 0) return;
        tma_store_wait<0>();
        if (m0 >= 0) push(sig, m0, c0);
        push(sig, m1, c1);
        m0 = m1 = -1;
    }
    __device__ __forceinline__ void finish(uint32_t* sig) { flush(sig); sig_push(sig, 0xffffffffu); }
};

// TMEM slots in item order: an item of k slots (k = 4, 2, 1) starts at the next multiple of k (wrapping at 4).
struct SlotRing {
    int pos = 0;
    uint32_t epar = 0;       // bit s: parity of the number of items that have used slot s
    uint32_t fpar = 0;       // bit s: parity of the number of items that STARTED at slot s
    __device__ __forceinline__ int take(int k) {
        int s = (pos + k - 1) / k * k;
        if (s >= WAVE_SLOTS) s = 0;
        pos = (s + k) % WAVE_SLOTS;
        return s;
    }
};

#ifdef BCL_WAVE_PROF
#define WPROF_DECL long long wp[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long wp_t0 = clock64();
#define WPROF(slot, stmt) { const long long wp_a = clock64(); stmt; wp[slot] += clock64() - wp_a; }
#define WPROF_BEGIN(v) const long long v = clock64();
#define WPROF_END(slot, v) wp[slot] += clock64() - v;
#define WPROF_DUMP(base) if (a.prof != nullptr && blockIdx.x < 8) { wp[7] = clock64() - wp_t0; for (int i_ = 0; i_ < 8; ++i_) a.prof[blockIdx.x * 32 + (base) + i_] = wp[i_]; }
#else
#define WPROF_DECL
#define WPROF(slot, stmt) stmt;
#define WPROF_BEGIN(v)
#define WPROF_END(slot, v)
#define WPROF_DUMP(base)
#endif

template <int B>
__global__ void __launch_bounds__(WAVE_THREADS, 1)
prefill_wave_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWf, const __grid_constant__ CUtensorMap tmXfS,
                    const __grid_constant__ CUtensorMap tmXfL, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmYfS,
                    const __grid_constant__ CUtensorMap tmYfL, const __grid_constant__ CUtensorMap tmWi, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ WaveArgs a) {
    static_assert(B == 64 || B == 128, "wave kernel: B = 64 / 128");
    extern __shared__ __align__(1024) unsigned char wave_smem[];
    constexpr int ATOMS = B / KATOM;                         // 128-byte K atoms of a block / of a DFT matrix
    constexpr uint32_t A_ATOM = TILE_M * 128;                // x: 128 tokens x 128 B
    constexpr uint32_t W_ATOM = B * 128;                     // DFT matrices: B rows x 128 B
    constexpr uint32_t W_BYTES = ATOMS * W_ATOM;             // B x B x 2
    constexpr uint32_t STAGE_A = ATOMS * A_ATOM;             // one input block of the tile
    constexpr uint32_t STAGE_C = (B / 8) * 2048;             // Yf of (tile, j): [B/8 slabs][128 tokens][16 B]
    constexpr int ASPLIT = WAVE_ASPLIT(B);                   // A items per (tile, kg): halves of the spectrum (N = B / ASPLIT per block)
    constexpr int NBH = B / ASPLIT;                          // accumulator columns per block of an A item
    constexpr int SLOTS_A = 4 * NBH / 128, SLOTS_B = 2, SLOTS_C = 1;
    constexpr int FCH = NBH / 16;                            // forward epilogue: chunks of 8 frequencies per A item
    constexpr int OUT_ATOMS = B / KATOM;
    constexpr int KSTEPS_C = B / 16;
    static_assert(STAGE_A <= WAVE_STAGE && STAGE_C <= WAVE_STAGE && STAGE_BYTES <= WAVE_STAGE, "stage size");

    WaveBarriers* bars = reinterpret_cast<WaveBarriers*>(wave_smem);
    uint32_t* sig = reinterpret_cast<uint32_t*>(wave_smem + 512);
    unsigned char* epi = wave_smem + 1024;
    unsigned char* wf = wave_smem + 1024 + EPI_WARPS * WAVE_NBUF * 4096;
    unsigned char* wi = wf + W_BYTES;
    unsigned char* ring = wi + W_BYTES;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages, G = (int)gridDim.x, R = a.ring;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int i = 0; i < WAVE_SLOTS; ++i) { mbar_init(&bars->tfull[i], 1); mbar_init(&bars->tempty[i], EPI_WARPS); }
        mbar_init(&bars->wfull, 1);
        for (int i = 0; i <= SIG_RING; ++i) sig[i] = 0u;
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWf); tma_prefetch_desc(&tmXfS); tma_prefetch_desc(&tmXfL); tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmYfS); tma_prefetch_desc(&tmYfL); tma_prefetch_desc(&tmWi); tma_prefetch_desc(&tmOut);
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    WaveCursor cur{(int)blockIdx.x, (int)blockIdx.x, (int)blockIdx.x};
    int type, i0, i1;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wfull, 2 * W_BYTES);                    // both DFT matrices: resident for the whole kernel
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wf + at * W_ATOM, &tmWf, at * KATOM, 0, &bars->wfull);
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wi + at * W_ATOM, &tmWi, at * KATOM, 0, &bars->wfull);
            pdl_wait();
            WPROF_DECL
            int stage = 0; uint32_t ph = 0;
            int mrA = -1, mrB = -1, mrC = -1;
            while (cur.next(a, G, type, i0, i1)) {
                for (int item = i0; item < i1; ++item) {
                    if (type == 0) {
                        const int kg = (item / ASPLIT) % a.KG, m = (item / ASPLIT) / a.KG;
                        if (m != mrA) {                                          // the Xf slot is free once B has consumed the tile that held it
                            WPROF(0, if (m >= R && !ctr_wait(a.done + a.m_tiles + (m - R), a.target[1])) { set_status(a.status, 41); goto done; })
                            mrA = m;
                        }
                        if (item == i0) {                                        // this CTA's next A chunk: x of its tile from HBM into L2 now
                            const int nx = cur.na * a.chunk[0];
                            if (nx < a.n_items[0]) {
                                const int kg2 = (nx / ASPLIT) % a.KG, m2 = (nx / ASPLIT) / a.KG;
                                for (int ib = 0; ib < 4; ++ib)
                                    for (int at = 0; at < ATOMS; ++at) tma_prefetch_2d(&tmX, (4 * kg2 + ib) * B + at * KATOM, m2 * TILE_M);
                            }
                        }
                        for (int ib = 0; ib < 4; ++ib) {
                            WPROF(3, if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 42); goto done; })
                            unsigned char* sa = ring + (size_t)stage * WAVE_STAGE;
                            mbar_arrive_expect_tx(&bars->full[stage], STAGE_A);
                            for (int at = 0; at < ATOMS; ++at)                   // columns past in_features are zero-filled
                                tma_load_2d_hint(sa + at * A_ATOM, &tmX, (4 * kg + ib) * B + at * KATOM, m * TILE_M, &bars->full[stage], L2_EVICT_FIRST);
                            if (++stage == S) { stage = 0; ph ^= 1u; }
                        }
                    } else if (type == 1) {
                        const int nch = item % a.NCH, r = item / a.NCH, fg = r % a.FG, m = r / a.FG;
                        if (m != mrB) {                                          // Xf of the tile complete; the Yf slot consumed by C
                            WPROF(1, if (!ctr_wait(a.done + m, a.target[0])) { set_status(a.status, 43); goto done; })
                            WPROF(0, if (m >= R && !ctr_wait(a.done + 2 * a.m_tiles + (m - R), a.target[2])) { set_status(a.status, 44); goto done; })
                            mrB = m;
                        }
                        const int tg0 = (m % R) * (TILE_M / TOK_GROUP);
                        for (int fl = 0; fl < 4; ++fl) {
                            for (int ka = 0; ka < a.k_atoms; ++ka) {
                                WPROF(4, if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 45); goto done; })
                                unsigned char* sa = ring + (size_t)stage * WAVE_STAGE;
                                mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
                                tma_load_4d_hint(sa, &tmXfL, 0, tg0, ka * 8, 4 * fg + fl, &bars->full[stage], L2_EVICT_NORMAL);
                                tma_load_3d_hint(sa + A_STAGE_BYTES, &tmC, ka * KATOM, nch * NCHUNK, 4 * fg + fl, &bars->full[stage], L2_EVICT_LAST);
                                if (++stage == S) { stage = 0; ph ^= 1u; }
                            }
                        }
                    } else {
                        const int j = item % a.p, m = item / a.p;
                        if (m != mrC) {
                            WPROF(2, if (!ctr_wait(a.done + a.m_tiles + m, a.target[1])) { set_status(a.status, 46); goto done; })
                            mrC = m;
                        }
                        WPROF(5, if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 47); goto done; })
                        mbar_arrive_expect_tx(&bars->full[stage], STAGE_C);
                        tma_load_4d_hint(ring + (size_t)stage * WAVE_STAGE, &tmYfL, 0, (m % R) * (TILE_M / TOK_GROUP), j, 0, &bars->full[stage], L2_EVICT_NORMAL);
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
            }
            WPROF_DUMP(0)
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idescA = umma_idesc_16(TILE_M, NBH, a.act_bf16);
            const uint32_t idescB = umma_idesc_bf16(TILE_M, NCHUNK);
            const uint32_t idescC = umma_idesc_bf16(TILE_M, B);
            const uint32_t swf = smem_u32(wf), swi = smem_u32(wi);
            if (!mbar_wait(&bars->wfull, 0)) { set_status(a.status, 51); goto done; }
            int stage = 0; uint32_t ph = 0;
            SlotRing sr;
            WPROF_DECL
            while (cur.next(a, G, type, i0, i1)) {
                for (int item = i0; item < i1; ++item) {
                    const int k = type == 0 ? SLOTS_A : (type == 1 ? SLOTS_B : SLOTS_C);
                    const int s0 = sr.take(k);
                    for (int s = s0; s < s0 + k; ++s) {                          // the epilogue has drained these slots
                        WPROF(type, if (!mbar_wait(&bars->tempty[s], ((sr.epar >> s) & 1u) ^ 1u)) { set_status(a.status, 52); goto done; })
                        sr.epar ^= 1u << s;
                    }
                    tc_fence_after_sync();
                    const uint32_t d0 = tmem_base + (uint32_t)(s0 * 128);
                    if (type == 0) {
                        const uint32_t wrow = (uint32_t)(item % ASPLIT) * NBH * 128u;        // this half's rows of the DFT matrix
                        for (int ib = 0; ib < 4; ++ib) {
                            WPROF(3, if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 53); goto done; })
                            tc_fence_after_sync();
                            const uint32_t sa = smem_u32(ring + (size_t)stage * WAVE_STAGE);
#pragma unroll
                            for (int at = 0; at < ATOMS; ++at) {
                                const uint64_t adesc = umma_desc_k_sw128(sa + at * A_ATOM), bdesc = umma_desc_k_sw128(swf + at * W_ATOM + wrow);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_bf16_ss(d0 + (uint32_t)(ib * NBH), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idescA, (at | kk) != 0 ? 1u : 0u);
                            }
                            umma_commit(&bars->empty[stage]);
                            if (++stage == S) { stage = 0; ph ^= 1u; }
                        }
                    } else if (type == 1) {
                        for (int fl = 0; fl < 4; ++fl) {
                            const uint32_t d_tmem = d0 + (uint32_t)(fl * NCHUNK);
                            for (int ka = 0; ka < a.k_atoms; ++ka) {
                                WPROF(4, if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 54); goto done; })
                                tc_fence_after_sync();
                                const uint32_t sa = smem_u32(ring + (size_t)stage * WAVE_STAGE);
                                const uint64_t bdesc = umma_desc_k_sw128(sa + A_STAGE_BYTES);
                                const int ks = min(4, (a.KG - ka * 8) / 2);
                                for (int kk = 0; kk < ks; ++kk)
                                    umma_bf16_ss(d_tmem, umma_desc_k_nosw(sa + (uint32_t)kk * 4096u, 128u, 2048u), bdesc + (uint64_t)(2 * kk), idescB, (ka | kk) != 0 ? 1u : 0u);
                                umma_commit(&bars->empty[stage]);
                                if (++stage == S) { stage = 0; ph ^= 1u; }
                            }
                        }
                    } else {
                        WPROF(5, if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 55); goto done; })
                        tc_fence_after_sync();
                        const uint32_t sa = smem_u32(ring + (size_t)stage * WAVE_STAGE);
#pragma unroll
                        for (int kk = 0; kk < KSTEPS_C; ++kk)
                            umma_bf16_ss(d0, umma_desc_k_nosw(sa + (uint32_t)kk * 4096u, 128u, 2048u),
                                         umma_desc_k_sw128(swi + (kk / 4) * W_ATOM) + (uint64_t)(2 * (kk % 4)), idescC, kk != 0 ? 1u : 0u);
                        umma_commit(&bars->empty[stage]);
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                    umma_commit(&bars->tfull[s0]);
                }
            }
            WPROF_DUMP(8)
        }
    } else if (warp == SIG_WARP) {
        if (lane == 0) signaller_loop(sig, a.done, EPI_WARPS);
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;                                            // TMEM lanes [32 quarter, +32) are this warp's
        const int wh = (warp - 2) >> 2;                                          // first / second warp of the lane quarter
        unsigned char* stg = epi + (size_t)(warp - 2) * WAVE_NBUF * 4096;
        pdl_wait();
        int buf = 0;
        WaveTracker trk;
        SlotRing sr;
        WPROF_DECL
        while (cur.next(a, G, type, i0, i1)) {
            for (int item = i0; item < i1; ++item) {
                const int k = type == 0 ? SLOTS_A : (type == 1 ? SLOTS_B : SLOTS_C);
                const int s0 = sr.take(k);
                const uint32_t fph = (sr.fpar >> s0) & 1u;
                sr.fpar ^= 1u << s0;
                WPROF_BEGIN(wp_w)
                if (!__all_sync(0xffffffffu, epi_idle_poll(&bars->tfull[s0], fph)) && lane == 0) trk.flush(sig);    // about to idle: publish
                if (!__all_sync(0xffffffffu, mbar_wait(&bars->tfull[s0], fph))) { set_status(a.status, 61); goto done; }
                tc_fence_after_sync();
                WPROF_END(type, wp_w)
                WPROF_BEGIN(wp_b)
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s0 * 128);
                int m, groups = 0;
                if (type == 0) {
                    // ---- forward DFT: 16-byte chunk (f, kg, token) = {Re,Im} of the 4 blocks -> Xf
                    const int kg = (item / ASPLIT) % a.KG, hf = item % ASPLIT;
                    m = (item / ASPLIT) / a.KG;
                    const int tg = (m % R) * (TILE_M / TOK_GROUP) + quarter;
                    const int fc0 = wh * (FCH / 2), fc1 = fc0 + FCH / 2;
                    for (int fc = fc0; fc < fc1; ++fc) {
                        uint32_t r[4][16];
#pragma unroll
                        for (int ib = 0; ib < 4; ++ib) tmem_ld_32x32b_x16(taddr + (uint32_t)(ib * NBH + fc * 16), r[ib]);
                        tmem_ld_wait();
                        if (fc == fc1 - 1) {
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) for (int s = s0; s < s0 + SLOTS_A; ++s) mbar_arrive(&bars->tempty[s]);
                        }
                        if (lane == 0) tma_store_wait_read<WAVE_NBUF - 1>();
                        __syncwarp();
                        uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)buf * 4096) + lane;
#pragma unroll
                        for (int fl = 0; fl < 8; ++fl)
                            dst[fl * 32] = make_uint4(pack_bf(__uint_as_float(r[0][2 * fl]), __uint_as_float(r[0][2 * fl + 1])),
                                                      pack_bf(__uint_as_float(r[1][2 * fl]), __uint_as_float(r[1][2 * fl + 1])),
                                                      pack_bf(__uint_as_float(r[2][2 * fl]), __uint_as_float(r[2][2 * fl + 1])),
                                                      pack_bf(__uint_as_float(r[3][2 * fl]), __uint_as_float(r[3][2 * fl + 1])));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d_hint(&tmXfS, 0, tg, kg, (hf * FCH + fc) * 8, stg + (size_t)buf * 4096, L2_EVICT_LAST);
                            tma_store_commit();
                        }
                        if (++buf == WAVE_NBUF) buf = 0;
                        ++groups;
                    }
                } else if (type == 1) {
                    // ---- frequency GEMM: interleave the 4 frequencies into 16-byte chunks (fg, j, token) -> Yf
                    const int nch = item % a.NCH, r_ = item / a.NCH, fg = r_ % a.FG;
                    m = r_ / a.FG;
                    const int jc_n = min(4, (a.p - nch * 32 + 7) / 8);
                    const int jc0 = wh * 2, jc1 = min(jc_n, jc0 + 2);
                    if (jc0 >= jc1) {
                        tc_fence_before_sync();
                        if (lane == 0) for (int s = s0; s < s0 + SLOTS_B; ++s) mbar_arrive(&bars->tempty[s]);
                    }
                    const int tg = (m % R) * (TILE_M / TOK_GROUP) + quarter;
                    for (int jc = jc0; jc < jc1; ++jc) {
                        uint32_t r[4][16];
#pragma unroll
                        for (int fl = 0; fl < 4; ++fl) tmem_ld_32x32b_x16(taddr + (uint32_t)(fl * NCHUNK + jc * 16), r[fl]);
                        tmem_ld_wait();
                        if (jc == jc1 - 1) {
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) for (int s = s0; s < s0 + SLOTS_B; ++s) mbar_arrive(&bars->tempty[s]);
                        }
                        if (lane == 0) tma_store_wait_read<WAVE_NBUF - 1>();
                        __syncwarp();
                        uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)buf * 4096) + lane;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            dst[jj * 32] = make_uint4(pack_bf(__uint_as_float(r[0][2 * jj]), __uint_as_float(r[0][2 * jj + 1])),
                                                      pack_bf(__uint_as_float(r[1][2 * jj]), __uint_as_float(r[1][2 * jj + 1])),
                                                      pack_bf(__uint_as_float(r[2][2 * jj]), __uint_as_float(r[2][2 * jj + 1])),
                                                      pack_bf(__uint_as_float(r[3][2 * jj]), __uint_as_float(r[3][2 * jj + 1])));
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d_hint(&tmYfS, 0, tg, nch * 32 + jc * 8, fg, stg + (size_t)buf * 4096, L2_EVICT_LAST);
                            tma_store_commit();
                        }
                        if (++buf == WAVE_NBUF) buf = 0;
                        ++groups;
                    }
                } else {
                    // ---- inverse DFT: round to x.dtype (:266), + bias in that dtype (:270-271), swizzled tile -> y
                    const int j = item % a.p;
                    m = item / a.p;
                    const int h0 = (OUT_ATOMS == 2) ? wh : 0, h1 = (OUT_ATOMS == 2) ? wh + 1 : (wh == 0 ? 1 : 0);
                    if (h0 >= h1) {
                        tc_fence_before_sync();
                        if (lane == 0) mbar_arrive(&bars->tempty[s0]);
                    }
                    for (int h = h0; h < h1; ++h) {
                        uint32_t r[64];
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(h * 64), *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                        tmem_ld_32x32b_x32(taddr + (uint32_t)(h * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                        tmem_ld_wait();
                        if (h == h1 - 1) {
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bars->tempty[s0]);
                        }
                        if (lane == 0) tma_store_wait_read<WAVE_NBUF - 1>();
                        __syncwarp();
                        unsigned char* row = stg + (size_t)buf * 4096 + (size_t)lane * 128;
                        const int col0 = j * B + h * 64;
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
                            tma_store_2d_hint(&tmOut, col0, m * TILE_M + quarter * 32, stg + (size_t)buf * 4096, L2_EVICT_FIRST);
                            tma_store_commit();
                        }
                        if (++buf == WAVE_NBUF) buf = 0;
                        ++groups;
                    }
                }
                WPROF_END(3 + type, wp_b)
                if (lane == 0) { trk.commit(sig, type * a.m_tiles + m, groups); if (a.immediate) trk.flush(sig); }
            }
        }
        if (warp == 2 && lane == 0) { WPROF_DUMP(16) }
        if (warp == 9 && lane == 0) { WPROF_DUMP(24) }
        if (lane == 0) trk.finish(sig);
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

size_t wave_smem_bytes(int B, int stages) { return 1024 + (size_t)EPI_WARPS * WAVE_NBUF * 4096 + (size_t)2 * B * B * 2 + (size_t)stages * WAVE_STAGE; }

bool wave_supported(const ProblemDev& P, const Geometry& g) {
    return !g.split && (g.B == 64 || g.B == 128) && (P.x_dtype == BCL_F16 || P.x_dtype == BCL_BF16) && P.y_dtype == P.x_dtype && P.n_extra == 0 &&
           (P.bias == nullptr || P.bias_dtype == P.x_dtype);
}

// area: the workspace bytes available for the Xf and Yf rings.  counters: 3 * ceil(N / 128) ints, zeroed in stream order before the launch.
cudaError_t launch_prefill_wave(const ProblemDev& P, const Geometry& g, const void* spec_gemm, void* area, size_t area_bytes, int* status, int* counters,
                                int ctas, cudaStream_t st) {
    const int B = g.B;
    const bool bf = P.x_dtype == BCL_BF16;
    const CUtensorMapDataType adt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(spec_gemm);
    WaveArgs a{};
    a.status = status; a.bias = P.bias; a.done = counters; a.act_bf16 = bf ? 1 : 0;
    a.m_tiles = (int)((P.N + TILE_M - 1) / TILE_M);
    a.KG = g.KG; a.FG = g.FG; a.NCH = g.NCH; a.p = g.p; a.k_atoms = (g.Kp + KATOM - 1) / KATOM;
    a.per_tile[0] = g.KG * WAVE_ASPLIT(B); a.per_tile[1] = g.FG * g.NCH; a.per_tile[2] = g.p;
    for (int t = 0; t < 3; ++t) { a.n_items[t] = a.m_tiles * a.per_tile[t]; a.target[t] = a.per_tile[t] * EPI_WARPS; }
    // chunks: the three lists get (about) the same number of chunks per tile, so a CTA alternates A, B, C
    const int cpt = env_int("BCL_WAVE_CHUNKS_PER_TILE", g.KG);
    for (int t = 0; t < 3; ++t) { a.chunk[t] = (a.per_tile[t] + cpt / 2) / cpt; if (a.chunk[t] < 1) a.chunk[t] = 1; }
    // lag: tiles between the stages a CTA works on at the same time; it has to cover the skew between CTAs (about one chunk of
    // every type = ctas / chunks-per-tile tiles) plus the signalling latency.  ring: tiles of Xf / Yf kept (L2-resident).
    if (a.chunk[0] % WAVE_ASPLIT(B)) a.chunk[0] += WAVE_ASPLIT(B) - a.chunk[0] % WAVE_ASPLIT(B);     // both halves of a (tile, kg) on one CTA: x comes from L2 the second time
    const double tiles_per_round = (double)ctas / (double)((a.per_tile[0] + a.chunk[0] - 1) / a.chunk[0]);
    int lag = env_int("BCL_WAVE_LAG", (int)(tiles_per_round * 1.0 + 4.5));
    if (lag < 1) lag = 1;
    int ring = env_int("BCL_WAVE_RING", lag + (int)tiles_per_round + 2);
    const size_t tile_bytes = (size_t)TILE_M * g.F * (g.Kp + 2 * g.p) * 2;       // Xf + Yf of one token tile
    const int ring_tiles_cap = (int)((area_bytes > 512 ? area_bytes - 512 : 0) / tile_bytes);
    if (ring > ring_tiles_cap) ring = ring_tiles_cap;
    if (ring >= a.m_tiles) ring = a.m_tiles;                 // everything fits: no slot is ever reused
    else if (ring < lag + 2) return cudaErrorInvalidValue;   // a reused slot must be consumed before its overwrite comes up in the order
    a.lag = lag; a.ring = ring;
    a.immediate = env_int("BCL_WAVE_IMMEDIATE", 0);        // (measured at 256-2048 tokens: a hand-off costs 2.6-3.7 us -- bulk-store completion + the gpu-scope fence --
                                                           //  as much as a kernel boundary under PDL: 32 us vs 27 us for the three launches at 256 tokens)
    a.prof = getenv("BCL_WAVE_PROF_PTR") ? reinterpret_cast<long long*>(strtoull(getenv("BCL_WAVE_PROF_PTR"), nullptr, 0)) : nullptr;

    CUtensorMap tmX, tmWf, tmXfS, tmXfL, tmC, tmYfS, tmYfL, tmWi, tmOut;
    {
        const uint64_t dims[2] = {(uint64_t)g.q * B, (uint64_t)P.N};
        const uint64_t strides[1] = {(uint64_t)P.x_stride * 2};
        const uint32_t box[2] = {KATOM, TILE_M};
        if (!make_tensor_map(&tmX, adt, 2, P.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
    }
    {
        const uint64_t dims[2] = {(uint64_t)B, (uint64_t)B};
        const uint64_t strides[1] = {(uint64_t)B * 2};
        const uint32_t box[2] = {KATOM, (uint32_t)B};
        if (!make_tensor_map(&tmWf, adt, 2, sb + (bf ? g.wbf16_off : g.wf16_off), dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
        if (!make_tensor_map(&tmWi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, sb + g.winv_off, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
    }
    const int64_t ring_tok = (int64_t)ring * TILE_M;
    void* Xf = area;
    void* Yf = reinterpret_cast<unsigned char*>(area) + ((xf_bytes(g, ring_tok) + 255) & ~(size_t)255);
    if (!make_chunk_map(&tmXfS, Xf, ring_tok, g.KG, g.F, 1, 1, 8)) return cudaErrorNotSupported;
    if (!make_chunk_map(&tmXfL, Xf, ring_tok, g.KG, g.F, TILE_M / TOK_GROUP, 8, 1)) return cudaErrorNotSupported;
    if (!make_chunk_map(&tmYfS, Yf, ring_tok, g.p, g.FG, 1, 8, 1)) return cudaErrorNotSupported;
    if (!make_chunk_map(&tmYfL, Yf, ring_tok, g.p, g.FG, TILE_M / TOK_GROUP, 1, (uint32_t)g.FG)) return cudaErrorNotSupported;
    {
        const uint64_t dims[3] = {(uint64_t)g.Kp, (uint64_t)g.Np, (uint64_t)g.F};
        const uint64_t strides[2] = {(uint64_t)g.Kp * 2, (uint64_t)g.Kp * g.Np * 2};
        const uint32_t box[3] = {KATOM, NCHUNK, 1};
        if (!make_tensor_map(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, spec_gemm, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
    }
    {
        const uint64_t dims[2] = {(uint64_t)g.p * B, (uint64_t)P.N};
        const uint64_t strides[1] = {(uint64_t)P.y_stride * 2};
        const uint32_t box[2] = {KATOM, 32};
        if (!make_tensor_map(&tmOut, adt, 2, P.y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return cudaErrorNotSupported;
    }
    int stages = env_int("BCL_WAVE_STAGES", WAVE_MAX_STAGES);
    if (stages > WAVE_MAX_STAGES) stages = WAVE_MAX_STAGES;
    while (stages > 2 && wave_smem_bytes(B, stages) > 227 * 1024) --stages;
    a.stages = stages;
    const size_t smem = wave_smem_bytes(B, stages);
    int grid = ctas, most = 0;
    for (int t = 0; t < 3; ++t) { const int c = (a.n_items[t] + a.chunk[t] - 1) / a.chunk[t]; if (c > most) most = c; }
    if (grid > most) grid = most;
    if (grid < 1) return cudaSuccess;
    cudaError_t e;
    if (B == 128) {
        if ((e = ensure_dynamic_smem(reinterpret_cast<const void*>(prefill_wave_kernel<128>), smem)) != cudaSuccess) return e;
        return launch_pdl(prefill_wave_kernel<128>, (unsigned)grid, WAVE_THREADS, smem, st, tmX, tmWf, tmXfS, tmXfL, tmC, tmYfS, tmYfL, tmWi, tmOut, a);
    }
    if ((e = ensure_dynamic_smem(reinterpret_cast<const void*>(prefill_wave_kernel<64>), smem)) != cudaSuccess) return e;
    return launch_pdl(prefill_wave_kernel<64>, (unsigned)grid, WAVE_THREADS, smem, st, tmX, tmWf, tmXfS, tmXfL, tmC, tmYfS, tmYfL, tmWi, tmOut, a);
}

}  // namespace gemm
}  // namespace bcl
