// bcl_prefill_bodies.cuh -- the three tcgen05 stages of the prefill path as device functions, shared by the stand-alone
// stage kernels (bcl_gemm.cu, bcl_dft_tc.cu) and the fused single-launch pipeline (bcl_prefill_fused.cu).
//
// Every stage is a persistent, warp-specialised loop over work items: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2-9 = epilogue (TMEM -> registers -> smem -> TMA store).  A CTA runs items first, first + stride, ...; in the fused
// launch the stages additionally hand token tiles to each other through per-tile completion counters in global memory
// (StageDeps): the producer warp waits for the tile's inputs, the epilogue warps publish the tile's outputs.
#pragma once
#include "bcl_common.cuh"
#include "bcl_sm100.cuh"
#include "bcl_gemm.cuh"

namespace bcl {
namespace gemm {

using namespace sm100;

// ---------------------------------------------------------------------------------------------------------
// cross-CTA hand-off of token tiles (fused launch only; all pointers NULL in the stand-alone kernels)
// ---------------------------------------------------------------------------------------------------------
struct StageDeps {
    const int* wait_ctr;     // producer: tile m may be loaded once wait_ctr[m] >= wait_target (the previous stage's done counter)
    int* done_ctr;           // every epilogue warp adds 1 to done_ctr[m] after its stores of an item of tile m have completed
    const int* lag_ctr;      // back-pressure: tile m also waits for lag_ctr[m - lag] >= lag_target, so that a fast stage cannot
    int wait_target;         //   run so far ahead of its consumer that the intermediates fall out of L2
    int lag_target, lag;
    int ascending;           // frequency GEMM: walk the token tiles upwards (fused) instead of downwards (stand-alone)
    int chunk;               // a CTA takes `chunk` consecutive items per round (same token tile: one counter poll / one signal); 0 = 1
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Bounded like mbar_wait: a protocol bug sets the status word instead of hanging the GPU.  The trailing proxy fence orders the
// acquire (generic proxy) before the TMA loads (async proxy) that read what the other CTA's TMA stores wrote.
static __device__ __noinline__ bool ctr_wait(const int* ctr, int target) {
    if (ld_acquire_gpu(ctr) < target) {
        const uint64_t t0 = global_ns();
#pragma unroll 1
        for (uint32_t spin = 0;; ++spin) {
            __nanosleep(200);
            if (ld_acquire_gpu(ctr) >= target) break;
            if ((spin & 255u) == 255u && global_ns() - t0 > 2000000000ull) return false;
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    return true;
}
// Called by the thread whose bulk-store groups have completed (cp.async.bulk.wait_group, not .read).
__device__ __forceinline__ void ctr_signal(int* ctr, int count) {
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(count) : "memory");
}
// Grid-wide barrier between the stages of the chain kernel (all CTAs co-resident: grid <= SM count, 1 CTA per SM).  Every CTA's bulk
// stores have completed (the epilogue warps wait on their bulk groups before a stage body returns); the release / acquire pair plus
// the async-proxy fences order them before the next stage's TMA loads on every SM.  `target` = CTAs x (barriers passed so far + 1).
__device__ __forceinline__ bool grid_barrier(int* ctr, int target) {
    __shared__ int ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ctr), "r"(1) : "memory");
        ok = ctr_wait(ctr, target) ? 1 : 0;
    }
    __syncthreads();
    return ok != 0;
}
__device__ __forceinline__ bool deps_wait_tile(const StageDeps& d, int m) {
    if (d.wait_ctr != nullptr && !ctr_wait(d.wait_ctr + m, d.wait_target)) return false;
    if (d.lag_ctr != nullptr && m >= d.lag && !ctr_wait(d.lag_ctr + (m - d.lag), d.lag_target)) return false;
    return true;
}
// Publishing a tile.  A gpu-scope release fence costs ~2.5 us on an SM with bulk stores in flight (MEMBAR.ALL.GPU waits for the
// SM's outstanding writes; measured with the per-item clock stamps, profiles/r1_prefill_notes.md), far too much for the
// epilogue warps.  So they only hand (tile, item count) to a dedicated signaller warp through a small shared-memory queue once
// their bulk stores of the tile have completed; the signaller does the fence and the release-adds, off the critical path.
//   queue: sig[0] = tickets handed out, sig[1 + k] = entry k of a 64-deep ring (0 = empty, 0xffffffff = "this warp is done")
constexpr int SIG_WARP = 10;                 // fused launch only: warp 10 of a 352-thread block
constexpr int SIG_RING = 64;
__device__ __forceinline__ void sig_push(uint32_t* sig, uint32_t entry) {
    asm volatile("fence.proxy.async;" ::: "memory");                     // completed bulk stores -> ordered before the hand-off
    const uint32_t slot = atomicAdd(sig, 1u) & (SIG_RING - 1);
    volatile uint32_t* e = sig + 1 + slot;
    while (*e != 0u) __nanosleep(32);                                   // ring full: practically never (8 warps x a few entries)
    __threadfence_block();
    *e = entry;
}
// the signaller (one thread): drains the queue until all epilogue warps have said goodbye; one gpu-scope fence per batch
__device__ __forceinline__ void signaller_loop(uint32_t* sig, int* done_ctr, int epi_warps) {
    uint32_t tail = 0;
    int finished = 0;
    const uint64_t t0 = global_ns();
    while (finished < epi_warps) {
        uint32_t batch[8];
        int nb = 0;
        for (; nb < 8; ++nb) {
            volatile uint32_t* e = sig + 1 + ((tail + nb) & (SIG_RING - 1));
            const uint32_t v = *e;
            if (v == 0u) break;
            batch[nb] = v;
            *e = 0u;
        }
        if (nb == 0) {
            __nanosleep(64);
            if (global_ns() - t0 > 4000000000ull) return;               // a warp bailed out on an error: do not hang
            continue;
        }
        tail += nb;
        __threadfence_block();
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        for (int i = 0; i < nb; ++i) {
            if (batch[i] == 0xffffffffu) { ++finished; continue; }
            asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(done_ctr + (batch[i] >> 8) - 1), "r"((int)(batch[i] & 0xffu)) : "memory");
        }
    }
}
// Epilogue-side bookkeeping (lane 0 of each epilogue warp).  Items of one token tile are counted locally and handed over
// when the warp moves on to another tile: by then the tile's bulk stores are older than the groups of the item just
// committed, so wait_group<groups> returns at once and nothing stalls.  When the warp would go idle for long
// (epi_idle_poll) it flushes instead -- a blocked CTA never holds a signal back, so the tile hand-off has no cycles.
struct DoneTracker {
    int cur_m = -1, cur_cnt = 0;
    __device__ __forceinline__ void flush(uint32_t* sig) {
        if (cur_m >= 0) { tma_store_wait<0>(); sig_push(sig, ((uint32_t)(cur_m + 1) << 8) | (uint32_t)cur_cnt); cur_m = -1; cur_cnt = 0; }
    }
    __device__ __forceinline__ void item_committed(uint32_t* sig, int m, int groups) {   // `groups` bulk groups committed for this item
        if (m == cur_m) { ++cur_cnt; return; }
        if (cur_m >= 0) {
            switch (groups) {
                case 1: tma_store_wait<1>(); break;
                case 2: tma_store_wait<2>(); break;
                case 3: tma_store_wait<3>(); break;
                case 4: tma_store_wait<4>(); break;
                default: tma_store_wait<0>(); break;
            }
            sig_push(sig, ((uint32_t)(cur_m + 1) << 8) | (uint32_t)cur_cnt);
        }
        cur_m = m; cur_cnt = 1;
    }
    __device__ __forceinline__ void finish(uint32_t* sig) { flush(sig); sig_push(sig, 0xffffffffu); }
};
// true: the barrier phase completed within ~1.5 us of (non-blocking) polling
__device__ __forceinline__ bool epi_idle_poll(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (int i = 0; i < 24; ++i) {
        if (mbar_test(addr, parity)) return true;
        __nanosleep(40);
    }
    return false;
}

__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int bf16) {
    if (bf16) return pack_bf(a, b);
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------
// stage 2: batched per-frequency GEMM on tcgen05.  Work item = (frequency group of 4, 128-token tile, 64-column chunk):
//   D_fl[128 tokens x 64] = A_f[128 x Kp] * B_f[64 x Kp]^T   for the 4 frequencies f = 4 fg + fl, 4 TMEM accumulators,
//   then the epilogue interleaves the four accumulators into 16-byte chunks (j, token, {fl, re/im}) and TMA-stores them.
// ---------------------------------------------------------------------------------------------------------
constexpr int EPI_WARPS = 8;         // two per TMEM lane quarter: one epilogue warp per SMSP is latency-bound (profiles/r1_prefill_notes.md)
constexpr int GEMM_THREADS = 32 * (2 + EPI_WARPS);   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int MAX_STAGES = 8;
#ifndef BCL_EPI_NBUF
#define BCL_EPI_NBUF 2
#endif
constexpr int EPI_NBUF = BCL_EPI_NBUF;                       // staging buffers per epilogue warp: TMA stores in flight
constexpr uint32_t A_STAGE_BYTES = TILE_M * KATOM * 2;       // [8 kg][128 tokens][16 B], no swizzle ("interleaved" K-major)
constexpr uint32_t B_STAGE_BYTES = NCHUNK * KATOM * 2;       // 64 rows x 128 B, 128-byte swizzle
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr uint32_t EPI_BUF_BYTES = 8 * 32 * 16;              // [8 j][32 tokens][16 B]
constexpr int SET_COLS = 4 * NCHUNK;                         // TMEM columns of one accumulator set (4 frequencies)

// Debug builds only (make TRACE=1 -> -DBCL_ENABLE_TRACE; tools/trace_stages.py, tools/trace_fused.py): per-item clock64 stamps of
// the first CTA of a stage, 8 slots per item, first 64 items, into the buffer named by BCL_TRACE_PTR.  Compiled out otherwise.
#ifdef BCL_ENABLE_TRACE
#define BCL_TRACE(slot, itv) do { if (a.trace != nullptr && first == 0 && (itv) < 64) a.trace[(itv) * 8 + (slot)] = clock64(); } while (0)
#else
#define BCL_TRACE(slot, itv) do { (void)(itv); } while (0)
#endif

struct GemmArgs {
    int* status;
    int ntok, p, KG, FG, k_atoms, m_tiles, NCH, num_items, stages;
    StageDeps deps;
    long long* trace;
};

struct GemmBarriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// Stand-alone launch: token tiles are walked from the LAST one down: the rfft stage wrote Xf in ascending tile order, so its
// most recent ~100 MB are still in L2 when this kernel starts; the inverse stage then walks up again and finds this kernel's
// last Yf.  Fused launch: upwards, right behind the forward-DFT CTAs.
__device__ __forceinline__ void decode_item(const GemmArgs& a, int item, int& fg, int& m, int& nch) {
    nch = item % a.NCH;
    const int r = item / a.NCH;
    fg = r % a.FG;
    m = a.deps.ascending ? r / a.FG : a.m_tiles - 1 - r / a.FG;
}

// (hi, lo) split of two floats: hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void split_bf(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    __nv_bfloat162 h; h.x = ha; h.y = hb;
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = pack_bf(a - __bfloat162float(ha), b - __bfloat162float(hb));
}
struct GemmLoMaps { CUtensorMap A, B, Y; };      // the lo halves of the split layout (bcl_gemm.cuh: geometry(..., split = 1))

// SPLIT: operands are (hi, lo) bf16 pairs; D += Ahi*Bhi + Ahi*Blo + Alo*Bhi (fp32 accumulate), outputs stored as (hi, lo).
template <bool SPLIT>
__device__ __forceinline__ void freq_gemm_body(unsigned char* gemm_smem, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                                               const GemmLoMaps* lo, const GemmArgs& a, const int first, const int stride) {
    constexpr uint32_t STAGE = SPLIT ? 2 * STAGE_BYTES : STAGE_BYTES;            // [A hi | B hi | A lo | B lo]
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(gemm_smem);
    uint32_t* sig = reinterpret_cast<uint32_t*>(gemm_smem + 512);                // signaller queue (fused launch)
    unsigned char* epi = gemm_smem + 1024;                                       // EPI_WARPS x EPI_NBUF buffers x EPI_BUF_BYTES
    unsigned char* ring = gemm_smem + 1024 + EPI_WARPS * EPI_NBUF * EPI_BUF_BYTES;               // stays 1024-byte aligned
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;
    const int C = a.deps.chunk > 0 ? a.deps.chunk : 1;      // consecutive items per round

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EPI_WARPS); }
        for (int i = 0; i <= SIG_RING; ++i) sig[i] = 0u;
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmY);
        if constexpr (SPLIT) { tma_prefetch_desc(&lo->A); tma_prefetch_desc(&lo->B); tma_prefetch_desc(&lo->Y); }
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 2 * SET_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            pdl_wait();                                                          // Xf comes from the previous kernel
            int stage = 0; uint32_t ph = 0; int m_ready = -1, pit = 0;
            for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++pit) {
                int fg, m, nch; decode_item(a, item, fg, m, nch);
                if (m != m_ready) {                                              // fused launch: Xf of this tile complete?
                    if (!deps_wait_tile(a.deps, m)) { set_status(a.status, 5); goto done; }
                    m_ready = m;
                }
                BCL_TRACE(0, pit);
                for (int fl = 0; fl < 4; ++fl) {
                    for (int ka = 0; ka < a.k_atoms; ++ka) {
                        if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 1); goto done; }
                        unsigned char* sa = ring + (size_t)stage * STAGE;
                        mbar_arrive_expect_tx(&bars->full[stage], STAGE);
                        tma_load_4d_hint(sa, &tmA, 0, m * (TILE_M / TOK_GROUP), ka * 8, 4 * fg + fl, &bars->full[stage], L2_EVICT_FIRST);   // read once
                        tma_load_3d_hint(sa + A_STAGE_BYTES, &tmB, ka * KATOM, nch * NCHUNK, 4 * fg + fl, &bars->full[stage], L2_EVICT_LAST);   // re-read by every tile
                        if constexpr (SPLIT) {
                            tma_load_4d_hint(sa + STAGE_BYTES, &lo->A, 0, m * (TILE_M / TOK_GROUP), ka * 8, 4 * fg + fl, &bars->full[stage], L2_EVICT_FIRST);
                            tma_load_3d_hint(sa + STAGE_BYTES + A_STAGE_BYTES, &lo->B, ka * KATOM, nch * NCHUNK, 4 * fg + fl, &bars->full[stage], L2_EVICT_LAST);
                        }
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, NCHUNK);
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++it) {
                const int set = it & 1; const uint32_t set_ph = (uint32_t)(it >> 1) & 1u;
                if (!mbar_wait(&bars->tmem_empty[set], set_ph ^ 1u)) { set_status(a.status, 2); goto done; }   // epilogue drained this set
                tc_fence_after_sync();
                BCL_TRACE(2, it);
                for (int fl = 0; fl < 4; ++fl) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(set * SET_COLS + fl * NCHUNK);
                    for (int ka = 0; ka < a.k_atoms; ++ka) {
                        if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 3); goto done; }
                        tc_fence_after_sync();
                        const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE);
                        const uint64_t bdesc = umma_desc_k_sw128(sa + A_STAGE_BYTES);
                        const int ks = min(4, (a.KG - ka * 8) / 2);
                        for (int kk = 0; kk < ks; ++kk) {                        // one k-step = 16 elements = 2 kg slabs of 2048 bytes
                            const uint64_t adesc = umma_desc_k_nosw(sa + (uint32_t)kk * 4096u, 128u, 2048u);
                            umma_bf16_ss(d_tmem, adesc, bdesc + (uint64_t)(2 * kk), idesc, (ka | kk) != 0 ? 1u : 0u);
                            if constexpr (SPLIT) {
                                const uint64_t bdesc_lo = umma_desc_k_sw128(sa + STAGE_BYTES + A_STAGE_BYTES);
                                umma_bf16_ss(d_tmem, adesc, bdesc_lo + (uint64_t)(2 * kk), idesc, 1u);
                                umma_bf16_ss(d_tmem, umma_desc_k_nosw(sa + STAGE_BYTES + (uint32_t)kk * 4096u, 128u, 2048u), bdesc + (uint64_t)(2 * kk), idesc, 1u);
                            }
                        }
                        umma_commit(&bars->empty[stage]);                        // smem stage reusable once these MMAs retire
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                }
                umma_commit(&bars->tmem_full[set]);                              // all 4 accumulators complete -> epilogue
                BCL_TRACE(4, it);
            }
        }
    } else if (warp == SIG_WARP) {
        if (lane == 0 && a.deps.done_ctr != nullptr) signaller_loop(sig, a.deps.done_ctr, EPI_WARPS);
    } else {
        // ===================== epilogue warps: TMEM -> registers -> bf16 chunks -> smem -> TMA store =====================
        const int quarter = warp & 3;                                            // TMEM lanes [32*quarter, +32) are this warp's
        const int half = (warp - 2) >> 2;                                        // which two of the four 8-block column groups
        unsigned char* stg = epi + (size_t)(warp - 2) * EPI_NBUF * EPI_BUF_BYTES;
        pdl_wait();                                                              // Yf may still be read by the previous kernel
        int it = 0, buf = 0;
        DoneTracker done_trk;
        for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++it) {
            int fg, m, nch; decode_item(a, item, fg, m, nch);
            const int set = it & 1; const uint32_t set_ph = (uint32_t)(it >> 1) & 1u;
            if (a.deps.done_ctr != nullptr && !__all_sync(0xffffffffu, epi_idle_poll(&bars->tmem_full[set], set_ph)) && lane == 0)
                done_trk.flush(sig);                                 // about to idle: publish what is finished
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[set], set_ph))) { set_status(a.status, 4); goto done; }
            tc_fence_after_sync();
            if (warp == 2 && lane == 0) BCL_TRACE(5, it);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * SET_COLS);
            const int jc_n = min(4, (a.p - nch * 32 + 7) / 8);                   // 8-block column groups with at least one real block
            const int jc0 = half * 2, jc1 = min(jc_n, jc0 + 2);
            if (jc0 >= jc1) {                                                    // nothing to read: hand the set back at once
                tc_fence_before_sync();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[set]);
            }
            for (int jc = jc0; jc < jc1; ++jc) {
                uint32_t r[4][16];
#pragma unroll
                for (int fl = 0; fl < 4; ++fl) tmem_ld_32x32b_x16(taddr + (uint32_t)(fl * NCHUNK + jc * 16), r[fl]);
                tmem_ld_wait();
                if (jc == jc1 - 1) {                                             // this warp's columns are read: hand the set back to the MMA warp
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tmem_empty[set]);
                    if (warp == 2 && lane == 0) BCL_TRACE(6, it);
                }
                const int tg = m * (TILE_M / TOK_GROUP) + quarter, j0 = nch * 32 + jc * 8;
                if constexpr (!SPLIT) {
                    if (lane == 0) tma_store_wait_read<EPI_NBUF - 1>();          // the store that last read this buffer has finished
                    __syncwarp();
                    uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)buf * EPI_BUF_BYTES) + lane;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        dst[jj * 32] = make_uint4(pack_bf(__uint_as_float(r[0][2 * jj]), __uint_as_float(r[0][2 * jj + 1])),
                                                  pack_bf(__uint_as_float(r[1][2 * jj]), __uint_as_float(r[1][2 * jj + 1])),
                                                  pack_bf(__uint_as_float(r[2][2 * jj]), __uint_as_float(r[2][2 * jj + 1])),
                                                  pack_bf(__uint_as_float(r[3][2 * jj]), __uint_as_float(r[3][2 * jj + 1])));
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d_hint(&tmY, 0, tg, j0, fg, stg + (size_t)buf * EPI_BUF_BYTES, L2_EVICT_LAST);   // the next stage reads it
                        tma_store_commit();
                    }
                    if (++buf == EPI_NBUF) buf = 0;
                } else {
                    static_assert(EPI_NBUF >= 2, "split epilogue: one staging buffer for hi, one for lo");
                    if (lane == 0) tma_store_wait_read<0>();                     // both buffers (hi, lo) free again
                    __syncwarp();
                    uint4* dh = reinterpret_cast<uint4*>(stg) + lane;
                    uint4* dl = reinterpret_cast<uint4*>(stg + EPI_BUF_BYTES) + lane;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        uint32_t hi[4], lw[4];
#pragma unroll
                        for (int fl = 0; fl < 4; ++fl) split_bf(__uint_as_float(r[fl][2 * jj]), __uint_as_float(r[fl][2 * jj + 1]), hi[fl], lw[fl]);
                        dh[jj * 32] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        dl[jj * 32] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d_hint(&tmY, 0, tg, j0, fg, stg, L2_EVICT_LAST);
                        tma_store_4d_hint(&lo->Y, 0, tg, j0, fg, stg + EPI_BUF_BYTES, L2_EVICT_LAST);
                        tma_store_commit();
                    }
                }
            }
            if (warp == 2 && lane == 0) BCL_TRACE(7, it);
            if (a.deps.done_ctr != nullptr && lane == 0) done_trk.item_committed(sig, m, max(jc1 - jc0, 0));
        }
        if (lane == 0) {
            tma_store_wait<0>();
            if (a.deps.done_ctr != nullptr) done_trk.finish(sig);
        }
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * SET_COLS);
    if (threadIdx.x == 0) {                                                      // the chain kernel re-initialises this memory for the next stage
        for (int s = 0; s < S; ++s) { mbar_inval(&bars->full[s]); mbar_inval(&bars->empty[s]); }
        for (int i = 0; i < 2; ++i) { mbar_inval(&bars->tmem_full[i]); mbar_inval(&bars->tmem_empty[i]); }
    }
}


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
    int half;                // B = 256 only: which slice of the spectrum rows (forward) / output samples (inverse) this launch computes (DftShape::NB each)
    StageDeps deps;
    long long* trace;
};

// B = 256 on the tensor cores: a 256 x 256 DFT matrix (128 KB) plus four 256-column accumulators do not fit, so a launch computes
// HALF of the N dimension (128 spectrum rows / output samples, `DftArgs::half`) with the full K = 256 fed in two 128-wide parts
// (one smem stage each), and the stage is launched twice.  The activations are read twice; the CUDA-core FFT stages this
// replaces are instruction-bound at 3-5x the HBM time (profiles/r1_prefill_notes.md).
// fp32 spectra (SPLIT): the (hi, lo) pair of DFT-matrix slices must fit as well, so a B = 256 launch computes a QUARTER of the N
// dimension (64 rows / samples; 2 x 32 KB of matrix) and the stage is launched four times.
template <int B, bool SPLIT = false> struct DftShape {
    static constexpr int NSPLIT = B == 256 ? (SPLIT ? 4 : 2) : 1;   // launches per stage = slices of the N dimension
    static constexpr int NB = B / NSPLIT;                     // accumulator columns per block = MMA N
    static constexpr int KATOMS = B / KATOM;                  // 128-byte K atoms over the whole K = B
    static constexpr int PARTS = B == 256 ? 2 : 1;            // K parts = smem stages per block
    static constexpr int APP = KATOMS / PARTS;                // atoms per part
};



// =========================================================================================================
// forward: x -> Xf
// =========================================================================================================
struct DftLoMaps { CUtensorMap W, IO; };         // split layout: the lo DFT matrix, and Xf_lo (forward, store) / Yf_lo (inverse, load)
// Sharded layer, exchange fused into the producer (sharded.py exchange="p2p"): the inverse stage's finished tile is TMA-stored into the
// same rows / columns of every peer's y as well (peer-mapped symmetric memory: the stores travel over NVLink), so no collective follows.
struct DftExtraOut { CUtensorMap out[kMaxExtraDst]; int n; };

// SPLIT (fp32 spectra): the activations are exact in their 16-bit format; W = Whi + Wlo (same format), two MMAs per k-step;
// the fp32 accumulators leave as (hi, lo) bf16 pairs into Xf and Xf_lo.
template <int B, bool SPLIT>
__device__ __forceinline__ void dft_fwd_body(unsigned char* dft_smem, const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmXf,
                                             const DftLoMaps* lo, const DftArgs& a, const int first, const int stride) {
    using SH = DftShape<B, SPLIT>;
    constexpr int NB = SH::NB, PARTS = SH::PARTS, APP = SH::APP;
    constexpr int ATOMS = SH::KATOMS;                        // 128-byte K atoms per block (K = B time samples)
    constexpr uint32_t A_ATOM = TILE_M * 128;                // 128 tokens x 128 B
    constexpr uint32_t W_ATOM = NB * 128;                    // NB spectrum rows x 128 B
    constexpr uint32_t STAGE = APP * A_ATOM;                 // one K part of one input block of the tile
    DftBarriers* bars = reinterpret_cast<DftBarriers*>(dft_smem);
    uint32_t* sig = reinterpret_cast<uint32_t*>(dft_smem + 512);                // signaller queue (fused launch)
    unsigned char* epi = dft_smem + 1024;                    // 4 warps x 2 x DFT_EPI_BUF
    unsigned char* wsm = dft_smem + 1024 + DFT_EPI_WARPS * DFT_NBUF * DFT_EPI_BUF;
    unsigned char* ring = wsm + (SPLIT ? 2 : 1) * ATOMS * W_ATOM;           // split: [W hi | W lo]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;
    const int C = a.deps.chunk > 0 ? a.deps.chunk : 1;      // consecutive items per round

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full[0], 1); mbar_init(&bars->tmem_empty[0], DFT_EPI_WARPS);
        mbar_init(&bars->wfull, 1);
        for (int i = 0; i <= SIG_RING; ++i) sig[i] = 0u;
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmXf);
        if constexpr (SPLIT) { tma_prefetch_desc(&lo->W); tma_prefetch_desc(&lo->IO); }
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 4 * NB);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wfull, (SPLIT ? 2 : 1) * ATOMS * W_ATOM);   // the DFT matrix: resident for the whole kernel
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + at * W_ATOM, &tmW, at * KATOM, a.half * NB, &bars->wfull);
            if constexpr (SPLIT)
                for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + (ATOMS + at) * W_ATOM, &lo->W, at * KATOM, a.half * NB, &bars->wfull);
            pdl_wait();                                                          // x may come from the previous kernel
            int stage = 0; uint32_t ph = 0; int m_ready = -1, pit = 0;
            for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++pit) {
                const int kg = item % a.groups, m = item / a.groups;
                if (m != m_ready) {                                              // fused launch: back-pressure from the frequency GEMM
                    if (!deps_wait_tile(a.deps, m)) { set_status(a.status, 16); goto done; }
                    m_ready = m;
                }
                BCL_TRACE(0, pit);
                for (int ibp = 0; ibp < 4 * PARTS; ++ibp) {                      // (input block, K part)
                    const int ib = ibp / PARTS, part = ibp % PARTS;
                    if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 11); goto done; }
                    unsigned char* sa = ring + (size_t)stage * STAGE;
                    mbar_arrive_expect_tx(&bars->full[stage], STAGE);
                    for (int at = 0; at < APP; ++at)                             // columns past in_features are zero-filled
                        tma_load_2d_hint(sa + at * A_ATOM, &tmX, (4 * kg + ib) * B + (part * APP + at) * KATOM, (int)(a.n_begin + (int64_t)m * TILE_M),
                                         &bars->full[stage], L2_EVICT_FIRST);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                BCL_TRACE(1, pit);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // One accumulator set = 4 blocks x B columns = all 512 TMEM columns: MMA and epilogue of consecutive items alternate.
            // Measured alternatives (per-item clock stamps, profiles/r1_prefill_notes.md): two half-spectrum sets (N = B/2, the
            // smem stages held for both passes) starve the MMAs of prefetch depth; two block-pair sets with the first pair parked
            // in registers make the epilogue slower than the overlap gains.  The epilogue itself sits at the TMEM read rate
            // (64 B/clk/SM: 256 KB of fp32 accumulators per item = 4.1k cycles).
            const uint32_t idesc = umma_idesc_16(TILE_M, NB, a.act_bf16);
            if (!mbar_wait(&bars->wfull, 0)) { set_status(a.status, 12); goto done; }
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++it) {
                if (!mbar_wait(&bars->tmem_empty[0], ((uint32_t)it & 1u) ^ 1u)) { set_status(a.status, 13); goto done; }
                tc_fence_after_sync();
                BCL_TRACE(2, it);
                for (int ibp = 0; ibp < 4 * PARTS; ++ibp) {                      // (input block, K part)
                    const int ib = ibp / PARTS, part = ibp % PARTS;
                    if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 14); goto done; }
                    tc_fence_after_sync();
                    if (ibp == 0) BCL_TRACE(3, it);
                    const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE), sw = smem_u32(wsm);
#pragma unroll
                    for (int at = 0; at < APP; ++at) {
                        const uint64_t adesc = umma_desc_k_sw128(sa + at * A_ATOM), bdesc = umma_desc_k_sw128(sw + (part * APP + at) * W_ATOM);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma_bf16_ss(tmem_base + (uint32_t)(ib * NB), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (part | at | kk) != 0 ? 1u : 0u);
                            if constexpr (SPLIT)
                                umma_bf16_ss(tmem_base + (uint32_t)(ib * NB), adesc + (uint64_t)(2 * kk),
                                             umma_desc_k_sw128(sw + (ATOMS + part * APP + at) * W_ATOM) + (uint64_t)(2 * kk), idesc, 1u);
                        }
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                umma_commit(&bars->tmem_full[0]);
                BCL_TRACE(4, it);
            }
        }
    } else if (warp == SIG_WARP) {
        if (lane == 0 && a.deps.done_ctr != nullptr) signaller_loop(sig, a.deps.done_ctr, DFT_EPI_WARPS);
    } else {
        const int quarter = warp & 3;
        unsigned char* stg = epi + (size_t)(warp - 2) * DFT_NBUF * DFT_EPI_BUF;
        pdl_wait();                                                              // Xf may still be read by the previous kernel
        int it = 0, buf = 0;
        DoneTracker done_trk;
        for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++it) {
            const int kg = item % a.groups, m = item / a.groups;
            if (a.deps.done_ctr != nullptr && !__all_sync(0xffffffffu, epi_idle_poll(&bars->tmem_full[0], (uint32_t)it & 1u)) && lane == 0)
                done_trk.flush(sig);                                             // about to idle: publish what is finished
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[0], (uint32_t)it & 1u))) { set_status(a.status, 15); goto done; }
            tc_fence_after_sync();
            if (warp == 2 && lane == 0) BCL_TRACE(5, it);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            constexpr int FC = NB / 16;                                          // chunks of 8 frequencies (16 accumulator columns)
            const int fc0 = ((warp - 2) >> 2) * (FC / 2), fc1 = fc0 + FC / 2;     // this warp's half of them
            for (int fc = fc0; fc < fc1; ++fc) {
                uint32_t r[4][16];
#pragma unroll
                for (int ib = 0; ib < 4; ++ib) tmem_ld_32x32b_x16(taddr + (uint32_t)(ib * NB + fc * 16), r[ib]);
                tmem_ld_wait();
                if (fc == fc1 - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tmem_empty[0]);
                    if (warp == 2 && lane == 0) BCL_TRACE(6, it);
                }
                if constexpr (!SPLIT) {
                    if (lane == 0) tma_store_wait_read<DFT_NBUF - 1>();
                    __syncwarp();
                    uint4* dst = reinterpret_cast<uint4*>(stg + (size_t)buf * DFT_EPI_BUF) + lane;
#pragma unroll
                    for (int fl = 0; fl < 8; ++fl)                               // chunk (f = 8 fc + fl, kg, token) = {Re,Im} x 4 blocks
                        dst[fl * 32] = make_uint4(pack_bf(__uint_as_float(r[0][2 * fl]), __uint_as_float(r[0][2 * fl + 1])),
                                                  pack_bf(__uint_as_float(r[1][2 * fl]), __uint_as_float(r[1][2 * fl + 1])),
                                                  pack_bf(__uint_as_float(r[2][2 * fl]), __uint_as_float(r[2][2 * fl + 1])),
                                                  pack_bf(__uint_as_float(r[3][2 * fl]), __uint_as_float(r[3][2 * fl + 1])));
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d_hint(&tmXf, 0, m * (TILE_M / TOK_GROUP) + quarter, kg, (a.half * FC + fc) * 8, stg + (size_t)buf * DFT_EPI_BUF, L2_EVICT_LAST);
                        tma_store_commit();
                    }
                    if (++buf == DFT_NBUF) buf = 0;
                } else {
                    static_assert(DFT_NBUF >= 2, "split epilogue: one staging buffer for hi, one for lo");
                    if (lane == 0) tma_store_wait_read<0>();                     // both buffers (hi, lo) free again
                    __syncwarp();
                    uint4* dh = reinterpret_cast<uint4*>(stg) + lane;
                    uint4* dl = reinterpret_cast<uint4*>(stg + DFT_EPI_BUF) + lane;
#pragma unroll
                    for (int fl = 0; fl < 8; ++fl) {
                        uint32_t hi[4], lw[4];
#pragma unroll
                        for (int ib = 0; ib < 4; ++ib) split_bf(__uint_as_float(r[ib][2 * fl]), __uint_as_float(r[ib][2 * fl + 1]), hi[ib], lw[ib]);
                        dh[fl * 32] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        dl[fl * 32] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d_hint(&tmXf, 0, m * (TILE_M / TOK_GROUP) + quarter, kg, (a.half * FC + fc) * 8, stg, L2_EVICT_LAST);
                        tma_store_4d_hint(&lo->IO, 0, m * (TILE_M / TOK_GROUP) + quarter, kg, (a.half * FC + fc) * 8, stg + DFT_EPI_BUF, L2_EVICT_LAST);
                        tma_store_commit();
                    }
                }
            }
            if (warp == 2 && lane == 0) BCL_TRACE(7, it);
            if (a.deps.done_ctr != nullptr && lane == 0) done_trk.item_committed(sig, m, FC / 2);
        }
        if (lane == 0) {
            tma_store_wait<0>();
            if (a.deps.done_ctr != nullptr) done_trk.finish(sig);
        }
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 4 * NB);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_inval(&bars->full[s]); mbar_inval(&bars->empty[s]); }
        mbar_inval(&bars->tmem_full[0]); mbar_inval(&bars->tmem_empty[0]); mbar_inval(&bars->wfull);
    }
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
// SPLIT (fp32 spectra): Yf = Yhi + Ylo, Winv = Whi + Wlo (bf16 pairs); three MMAs per k-step (Yhi Whi + Yhi Wlo + Ylo Whi).
// One staging buffer per epilogue warp instead of two, so that two 64 KB (hi, lo) stages fit next to the two DFT matrices.
template <int B, bool SPLIT>
__device__ __forceinline__ void dft_inv_body(unsigned char* dft_smem, const CUtensorMap& tmYf, const CUtensorMap& tmW, const CUtensorMap& tmOut,
                                             const DftLoMaps* lo, const DftArgs& a, const int first, const int stride, const DftExtraOut* ex = nullptr) {
    using SH = DftShape<B, SPLIT>;
    constexpr int NB = SH::NB, PARTS = SH::PARTS;
    constexpr int ATOMS = SH::KATOMS;                        // K atoms of the DFT matrix (K = B packed spectrum reals)
    constexpr int OUT_ATOMS = NB / KATOM;                    // 64-column pieces of one block's output (of this launch's half)
    constexpr int FG = B / 8 / PARTS;                        // 16-byte frequency slabs per K part = per smem stage
    constexpr int KSTEPS = B / 16 / PARTS;                   // k-steps (16 reals) per part
    constexpr uint32_t W_ATOM = NB * 128;                    // NB time rows x 128 B
    constexpr uint32_t STAGE_HALF = FG * 2048;               // [fg][128 tokens][16 B]
    constexpr uint32_t STAGE = (SPLIT ? 2 : 1) * STAGE_HALF; // split: [Yf hi | Yf lo]
    constexpr int NBUF = SPLIT ? 1 : DFT_NBUF;               // staging buffers per epilogue warp
    constexpr int NACC = 512 / NB > 4 ? 4 : 512 / NB;
    DftBarriers* bars = reinterpret_cast<DftBarriers*>(dft_smem);
    uint32_t* sig = reinterpret_cast<uint32_t*>(dft_smem + 512);                // signaller queue (fused launch)
    unsigned char* epi = dft_smem + 1024;
    unsigned char* wsm = dft_smem + 1024 + DFT_EPI_WARPS * NBUF * DFT_EPI_BUF;
    unsigned char* ring = wsm + (SPLIT ? 2 : 1) * ATOMS * W_ATOM;           // split: [W hi | W lo]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;
    const int C = a.deps.chunk > 0 ? a.deps.chunk : 1;      // consecutive items per round

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], DFT_EPI_WARPS); }
        mbar_init(&bars->wfull, 1);
        for (int i = 0; i <= SIG_RING; ++i) sig[i] = 0u;
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmYf); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmOut);
        if constexpr (SPLIT) { tma_prefetch_desc(&lo->W); tma_prefetch_desc(&lo->IO); }
        if (ex != nullptr) for (int e = 0; e < ex->n; ++e) tma_prefetch_desc(&ex->out[e]);
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, NACC * NB);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wfull, (SPLIT ? 2 : 1) * ATOMS * W_ATOM);
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + at * W_ATOM, &tmW, at * KATOM, a.half * NB, &bars->wfull);
            if constexpr (SPLIT)
                for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + (ATOMS + at) * W_ATOM, &lo->W, at * KATOM, a.half * NB, &bars->wfull);
            pdl_wait();                                                          // Yf comes from the previous kernel
            int stage = 0; uint32_t ph = 0; int m_ready = -1, pit = 0;
            for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++pit) {
                const int j = item % a.groups, m = item / a.groups;
                if (m != m_ready) {                                              // fused launch: Yf of this tile complete?
                    if (!deps_wait_tile(a.deps, m)) { set_status(a.status, 26); goto done; }
                    m_ready = m;
                }
                BCL_TRACE(0, pit);
                for (int part = 0; part < PARTS; ++part) {
                    if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 21); goto done; }
                    mbar_arrive_expect_tx(&bars->full[stage], STAGE);
                    tma_load_4d_hint(ring + (size_t)stage * STAGE, &tmYf, 0, m * (TILE_M / TOK_GROUP), j, part * FG, &bars->full[stage], L2_EVICT_FIRST);
                    if constexpr (SPLIT)
                        tma_load_4d_hint(ring + (size_t)stage * STAGE + STAGE_HALF, &lo->IO, 0, m * (TILE_M / TOK_GROUP), j, part * FG, &bars->full[stage], L2_EVICT_FIRST);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, NB);
            if (!mbar_wait(&bars->wfull, 0)) { set_status(a.status, 22); goto done; }
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++it) {
                const int acc = it % NACC; const uint32_t acc_ph = (uint32_t)(it / NACC) & 1u;
                if (!mbar_wait(&bars->tmem_empty[acc], acc_ph ^ 1u)) { set_status(a.status, 23); goto done; }
                BCL_TRACE(2, it);
                for (int part = 0; part < PARTS; ++part) {
                    if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 24); goto done; }
                    tc_fence_after_sync();
                    if (part == 0) BCL_TRACE(3, it);
                    const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE), sw = smem_u32(wsm);
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) {                        // K = B packed spectrum reals; one k-step = 2 fg slabs
                        const int kg = part * KSTEPS + kk;                       // k-step within the whole K
                        const uint64_t adesc = umma_desc_k_nosw(sa + (uint32_t)kk * 4096u, 128u, 2048u);
                        const uint64_t bdesc = umma_desc_k_sw128(sw + (kg / 4) * W_ATOM) + (uint64_t)(2 * (kg % 4));
                        umma_bf16_ss(tmem_base + (uint32_t)(acc * NB), adesc, bdesc, idesc, (part | kk) != 0 ? 1u : 0u);
                        if constexpr (SPLIT) {
                            umma_bf16_ss(tmem_base + (uint32_t)(acc * NB), adesc, umma_desc_k_sw128(sw + (ATOMS + kg / 4) * W_ATOM) + (uint64_t)(2 * (kg % 4)), idesc, 1u);
                            umma_bf16_ss(tmem_base + (uint32_t)(acc * NB), umma_desc_k_nosw(sa + STAGE_HALF + (uint32_t)kk * 4096u, 128u, 2048u), bdesc, idesc, 1u);
                        }
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                umma_commit(&bars->tmem_full[acc]);
                BCL_TRACE(4, it);
            }
        }
    } else if (warp == SIG_WARP) {
        if (lane == 0 && a.deps.done_ctr != nullptr) signaller_loop(sig, a.deps.done_ctr, DFT_EPI_WARPS);
    } else {
        const int quarter = warp & 3;
        unsigned char* stg = epi + (size_t)(warp - 2) * NBUF * DFT_EPI_BUF;
        pdl_wait();                                                              // y may still be read by the previous kernel
        int it = 0, buf = 0;
        DoneTracker done_trk;
        for (int blk = first * C; blk < a.num_items; blk += stride * C) for (int item = blk; item < min(blk + C, a.num_items); ++item, ++it) {
            const int j = item % a.groups, m = item / a.groups;
            const int acc = it % NACC; const uint32_t acc_ph = (uint32_t)(it / NACC) & 1u;
            if (a.deps.done_ctr != nullptr && !__all_sync(0xffffffffu, epi_idle_poll(&bars->tmem_full[acc], acc_ph)) && lane == 0)
                done_trk.flush(sig);                                 // about to idle: publish what is finished
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[acc], acc_ph))) { set_status(a.status, 25); goto done; }
            tc_fence_after_sync();
            if (warp == 2 && lane == 0) BCL_TRACE(5, it);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NB);
            const int wh = (warp - 2) >> 2;                                      // second warp of the lane quarter?
            const int h0 = (OUT_ATOMS == 2) ? wh : 0, h1 = (OUT_ATOMS == 2) ? wh + 1 : (wh == 0 ? 1 : 0);
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
                    if (warp == 2 && lane == 0) BCL_TRACE(6, it);
                }
                if (lane == 0) tma_store_wait_read<NBUF - 1>();
                __syncwarp();
                unsigned char* row = stg + (size_t)buf * DFT_EPI_BUF + (size_t)lane * 128;
                const int col0 = j * B + a.half * NB + h * 64;
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
                    if (ex != nullptr)                                           // the peers' copies of y: same tile, same coordinates
                        for (int e = 0; e < ex->n; ++e)
                            tma_store_2d(&ex->out[e], col0, (int)(a.n_begin + (int64_t)m * TILE_M + quarter * 32), stg + (size_t)buf * DFT_EPI_BUF);
                    tma_store_commit();
                }
                if (++buf == NBUF) buf = 0;
            }
            if (warp == 2 && lane == 0) BCL_TRACE(7, it);
            if (a.deps.done_ctr != nullptr && lane == 0) done_trk.item_committed(sig, m, max(h1 - h0, 0));
        }
        if (lane == 0) {
            tma_store_wait<0>();
            if (a.deps.done_ctr != nullptr) done_trk.finish(sig);
        }
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, NACC * NB);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_inval(&bars->full[s]); mbar_inval(&bars->empty[s]); }
        for (int i = 0; i < NACC; ++i) { mbar_inval(&bars->tmem_full[i]); mbar_inval(&bars->tmem_empty[i]); }
        mbar_inval(&bars->wfull);
    }
}


// =========================================================================================================
// inverse + SWIGLU: Yf of the stacked (gate | up) layer -> h = round(silu(round(gate))) * round(up)
// =========================================================================================================
// HF LlamaMLP.forward computes down_proj(act_fn(gate_proj(x)) * up_proj(x)) (call sites src/patch_llama_fft.py:379): act_fn and the
// product run on the projections' OUTPUTS, i.e. on values already rounded to x.dtype (:266), silu in fp32 rounded to x.dtype, the
// product rounded again -- the same arithmetic as the decode chain kernel's SWIGLU epilogue (bcl_stream_chain.cu).  Here the inverse
// DFT of output block j of gate (slab j of Yf) and of up (slab j + p/2) go to a PAIR of TMEM accumulators and the epilogue writes
// one block of h: the two 11008-wide projection outputs never reach HBM, and the separate act_fn / multiply kernels disappear.
template <bool BF16>
__device__ __forceinline__ void store_row_swiglu(unsigned char* row, int lane, const uint32_t (&rg)[64], const uint32_t (&ru)[64]) {
    constexpr int DT = BF16 ? BCL_BF16 : BCL_F16;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float hv[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float gv = round_to(__uint_as_float(rg[8 * c + 2 * e + k]), DT);
                const float uv = round_to(__uint_as_float(ru[8 * c + 2 * e + k]), DT);
                const float sv = round_to(__fdividef(gv, 1.0f + __expf(-gv)), DT);
                hv[k] = sv * uv;
            }
            w[e] = pack16(hv[0], hv[1], BF16 ? 1 : 0);
        }
        *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// SPLIT (fp32 spectra): Yf = Yhi + Ylo, Winv = Whi + Wlo (bf16 pairs), three MMAs per k-step as in dft_inv_body<B, true>; a stage holds
// [Yf hi | Yf lo] of one slab and every epilogue warp has one staging buffer, so two 64 KB stages fit next to the two DFT matrices.
template <int B, bool SPLIT>
__device__ __forceinline__ void dft_inv_swiglu_body(unsigned char* dft_smem, const CUtensorMap& tmYf, const CUtensorMap& tmW, const CUtensorMap& tmOut,
                                                    const DftLoMaps* lo, const DftArgs& a, const int first, const int stride) {
    static_assert(B == 64 || B == 128, "SWIGLU inverse stage: B = 64 / 128");
    constexpr int ATOMS = B / KATOM;
    constexpr int OUT_ATOMS = B / KATOM;
    constexpr int FG = B / 8;
    constexpr int KSTEPS = B / 16;
    constexpr uint32_t W_ATOM = B * 128;
    constexpr uint32_t STAGE_HALF = FG * 2048;               // Yf of (tile, slab): [fg][128 tokens][16 B]
    constexpr uint32_t STAGE = (SPLIT ? 2 : 1) * STAGE_HALF; // split: [Yf hi | Yf lo]
    constexpr int NBUF = SPLIT ? 1 : DFT_NBUF;               // staging buffers per epilogue warp
    constexpr int NPAIR = 2;                                 // accumulator pairs in flight (2 x 2 x B <= 512 columns)
    DftBarriers* bars = reinterpret_cast<DftBarriers*>(dft_smem);
    unsigned char* epi = dft_smem + 1024;
    unsigned char* wsm = dft_smem + 1024 + DFT_EPI_WARPS * NBUF * DFT_EPI_BUF;
    unsigned char* ring = wsm + (SPLIT ? 2 : 1) * ATOMS * W_ATOM;           // split: [W hi | W lo]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = a.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int i = 0; i < NPAIR; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], DFT_EPI_WARPS); }
        mbar_init(&bars->wfull, 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmYf); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmOut);
        if constexpr (SPLIT) { tma_prefetch_desc(&lo->W); tma_prefetch_desc(&lo->IO); }
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->wfull, (SPLIT ? 2 : 1) * ATOMS * W_ATOM);
            for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + at * W_ATOM, &tmW, at * KATOM, 0, &bars->wfull);
            if constexpr (SPLIT)
                for (int at = 0; at < ATOMS; ++at) tma_load_2d(wsm + (ATOMS + at) * W_ATOM, &lo->W, at * KATOM, 0, &bars->wfull);
            pdl_wait();                                                          // Yf comes from the previous kernel
            int stage = 0; uint32_t ph = 0;
            for (int item = first; item < a.num_items; item += stride) {
                const int j = item % a.groups, m = item / a.groups;
                for (int half = 0; half < 2; ++half) {                           // gate slab j, up slab j + p/2
                    if (!mbar_wait(&bars->empty[stage], ph ^ 1u)) { set_status(a.status, 71); goto done; }
                    mbar_arrive_expect_tx(&bars->full[stage], STAGE);
                    tma_load_4d_hint(ring + (size_t)stage * STAGE, &tmYf, 0, m * (TILE_M / TOK_GROUP), j + half * a.groups, 0, &bars->full[stage], L2_EVICT_FIRST);
                    if constexpr (SPLIT)
                        tma_load_4d_hint(ring + (size_t)stage * STAGE + STAGE_HALF, &lo->IO, 0, m * (TILE_M / TOK_GROUP), j + half * a.groups, 0, &bars->full[stage],
                                         L2_EVICT_FIRST);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TILE_M, B);
            if (!mbar_wait(&bars->wfull, 0)) { set_status(a.status, 72); goto done; }
            int stage = 0; uint32_t ph = 0; int it = 0;
            for (int item = first; item < a.num_items; item += stride, ++it) {
                const int pair = it % NPAIR; const uint32_t pair_ph = (uint32_t)(it / NPAIR) & 1u;
                if (!mbar_wait(&bars->tmem_empty[pair], pair_ph ^ 1u)) { set_status(a.status, 73); goto done; }
                tc_fence_after_sync();
                for (int half = 0; half < 2; ++half) {
                    if (!mbar_wait(&bars->full[stage], ph)) { set_status(a.status, 74); goto done; }
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(ring + (size_t)stage * STAGE), sw = smem_u32(wsm);
                    const uint32_t acc = tmem_base + (uint32_t)((2 * pair + half) * B);
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) {
                        const uint64_t adesc = umma_desc_k_nosw(sa + (uint32_t)kk * 4096u, 128u, 2048u);
                        const uint64_t bdesc = umma_desc_k_sw128(sw + (kk / 4) * W_ATOM) + (uint64_t)(2 * (kk % 4));
                        umma_bf16_ss(acc, adesc, bdesc, idesc, kk != 0 ? 1u : 0u);
                        if constexpr (SPLIT) {
                            umma_bf16_ss(acc, adesc, umma_desc_k_sw128(sw + (ATOMS + kk / 4) * W_ATOM) + (uint64_t)(2 * (kk % 4)), idesc, 1u);
                            umma_bf16_ss(acc, umma_desc_k_nosw(sa + STAGE_HALF + (uint32_t)kk * 4096u, 128u, 2048u), bdesc, idesc, 1u);
                        }
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == S) { stage = 0; ph ^= 1u; }
                }
                umma_commit(&bars->tmem_full[pair]);
            }
        }
    } else {
        const int quarter = warp & 3;
        unsigned char* stg = epi + (size_t)(warp - 2) * NBUF * DFT_EPI_BUF;
        pdl_wait();                                                              // h may still be read by the previous kernel
        int it = 0, buf = 0;
        for (int item = first; item < a.num_items; item += stride, ++it) {
            const int j = item % a.groups, m = item / a.groups;
            const int pair = it % NPAIR; const uint32_t pair_ph = (uint32_t)(it / NPAIR) & 1u;
            if (!__all_sync(0xffffffffu, mbar_wait(&bars->tmem_full[pair], pair_ph))) { set_status(a.status, 75); goto done; }
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(2 * pair * B);
            const int wh = (warp - 2) >> 2;
            const int h0 = (OUT_ATOMS == 2) ? wh : 0, h1 = (OUT_ATOMS == 2) ? wh + 1 : (wh == 0 ? 1 : 0);
            if (h0 >= h1) {                                                      // B = 64: the partner warp only releases
                tc_fence_before_sync();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[pair]);
            }
            for (int h = h0; h < h1; ++h) {
                uint32_t rg[64], ru[64];
                tmem_ld_32x32b_x32(taddr + (uint32_t)(h * 64), *reinterpret_cast<uint32_t(*)[32]>(&rg[0]));
                tmem_ld_32x32b_x32(taddr + (uint32_t)(h * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(&rg[32]));
                tmem_ld_32x32b_x32(taddr + (uint32_t)(B + h * 64), *reinterpret_cast<uint32_t(*)[32]>(&ru[0]));
                tmem_ld_32x32b_x32(taddr + (uint32_t)(B + h * 64 + 32), *reinterpret_cast<uint32_t(*)[32]>(&ru[32]));
                tmem_ld_wait();
                if (h == h1 - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->tmem_empty[pair]);
                }
                if (lane == 0) tma_store_wait_read<NBUF - 1>();
                __syncwarp();
                unsigned char* row = stg + (size_t)buf * DFT_EPI_BUF + (size_t)lane * 128;
                if (a.act_bf16) store_row_swiglu<true>(row, lane, rg, ru);
                else store_row_swiglu<false>(row, lane, rg, ru);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOut, j * B + h * 64, (int)(a.n_begin + (int64_t)m * TILE_M + quarter * 32), stg + (size_t)buf * DFT_EPI_BUF);
                    tma_store_commit();
                }
                if (++buf == NBUF) buf = 0;
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }
done:
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps + arguments of one stage (filled by prep_*, consumed by the stand-alone and the fused launchers)
// ---------------------------------------------------------------------------------------------------------
template <typename Args> struct StageLaunch {
    CUtensorMap tm_in, tm_w, tm_out;     // the stage's input, its resident / streamed matrix (DFT matrix or Cmat), its output
    Args a;
    size_t smem;
};
long long* trace_ptr(int stage);
bool prep_dft_fwd(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, int* status, StageLaunch<DftArgs>& L,
                  void* Xf_lo = nullptr, DftLoMaps* lo = nullptr);
bool prep_gemm(const void* Xf, const void* Cmat, void* Yf, int* status, const Geometry& g, int nc, StageLaunch<GemmArgs>& L, GemmLoMaps* lo = nullptr);
bool prep_gemm_lo(const void* Xf_lo, const void* Cmat, void* Yf_lo, const Geometry& g, int nc, GemmLoMaps* lo);
bool prep_dft_inv(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, StageLaunch<DftArgs>& L,
                  const void* Yf_lo = nullptr, DftLoMaps* lo = nullptr, DftExtraOut* ex = nullptr);
// g: geometry of the STACKED (gate | up) layer (g.p block-rows, even); P.y = h with (g.p / 2) * B columns
bool prep_dft_inv_swiglu(const ProblemDev& P, const Geometry& g, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, StageLaunch<DftArgs>& L,
                         const void* Yf_lo = nullptr, DftLoMaps* lo = nullptr);

}  // namespace gemm
}  // namespace bcl
