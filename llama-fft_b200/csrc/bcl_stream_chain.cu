// bcl_stream_chain.cu -- latency-oriented decode kernel: a CHAIN of dependent BC-linear steps (one token) in ONE launch.
//
// Why: at batch-1 decode the projections of a layer depend on each other (HF LlamaMLP.forward: down_proj(act_fn(gate_proj(x)) *
// up_proj(x)); reference call sites src/patch_llama_fft.py:379), one projection is 0.5-2.8 MB of spectrum = 0.1-0.4 us of HBM time,
// and a launch per module costs ~5 us of pure latency.  This kernel runs up to kMaxSteps dependent steps with a
// grid-wide barrier between them:
//
//   step s = one or more BC linears that read the SAME already-available activations (problems [step_first[s], step_first[s+1]));
//   a problem may carry the SWIGLU epilogue: its block-rows [0, p/2) are gate_proj, [p/2, p) up_proj (stacked along the output-block
//   axis like shared_input.py does) and the finisher writes  h = silu(round(gate)) * round(up)  -- the 11008-wide intermediate of the
//   MLP exists only as that one rounded vector, gate / up never reach memory.
//
// One persistent CTA per SM (grid <= SM count, all co-resident):
//   warp 0          TMA producer: 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) of the CTA's spectrum rows into a ring.  The
//                   spectra are constants of the stream, so the producer never waits for a step barrier: while the compute warps
//                   wait for the activations of step s+1, the ring already fills with its weights (and with the first step's weights
//                   while the PREVIOUS kernel of the stream drains: programmatic dependent launch).
//   warps 1..8      compute: cooperative rfft(x) of the step's input (bcl_fft.cuh, L2-coherent loads: x was written by other SMs a
//                   moment ago), split-q complex MAC out of the ring, partial spectra parked in shared memory, then warp 1 sums them,
//                   inverse-FFTs up to G rows at once and applies the reference epilogue (src/patch_llama_fft.py:263-271) or SWIGLU.
// Work split: a step's (problem, row) units are cut into gridDim.x contiguous, byte-balanced shares by a closed formula every warp
// evaluates itself -- no plan buffer, no dependent global loads: the problem descriptors are kernel parameters (constant bank).
// Grid barrier: a 64-bit arrival counter in global memory that is never reset: every launch advances it by exactly 2^20 (its
// gridDim.x * (nsteps - 1) arrivals plus one padding add), so a CTA recovers the launch's base value by rounding down whatever it
// reads before its own first arrival, and ONE counter serves every chain launch of a stream whatever their grids and step counts.
#include <cstdlib>
#include <cstdio>
#include "bcl_stream_parts.cuh"

namespace bcl {
namespace chain {

constexpr int NC = 8;                        // compute warps
constexpr int THREADS = 32 * (1 + NC);
constexpr int MAX_STAGES = 16;
constexpr unsigned long long QUANTUM = kChainQuantum;
__host__ __device__ constexpr int stage_blocks(int B) { return 4096 / B; }     // blocks per ring stage: 32 of B=128 (16 KB fp32)

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
    while (!mbar_try(addr, parity)) {}                               // try_wait suspends in hardware; the waits here are short by construction
}
__device__ inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ inline void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ inline void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ inline void compute_sync() { asm volatile("bar.sync 1, %0;" ::"r"(32 * NC) : "memory"); }
__device__ inline unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Weak (default, L1-allocating) global loads.  Inside a chain x was written by other SMs earlier in the SAME launch, so it must not come
// from a stale L1 line or the non-coherent path (__ldg): after the step barrier -- ld.acquire.gpu by one thread (ptxas invalidates the
// SM's L1 for it), then bar.sync -- weak loads are ordered behind the writers' release by causality (PTX memory model), and measured
// ~700 cycles faster than ld.global.cg on this path.
__device__ __forceinline__ uint32_t ld_weak_u32(const void* p) { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint2 ld_weak_u64(const void* p) { uint2 v; asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_weak_f4(const void* p) { float4 v; asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ inline uint64_t global_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct Barriers {
    uint64_t full[MAX_STAGES];      // ring stage filled by TMA              (1 arrival + tx bytes)
    uint64_t empty[MAX_STAGES];     // ring stage released by the compute warps (NC arrivals)
    uint64_t xissued;               // the compute warps have issued the loads of the first step's activations (NC arrivals)
};

// This CTA's share [u0, u1) of the units of problem g of a step (unit = one output block-row, or one (gate, up) row pair).  The
// flattened unit sequence of the step is cut by WEIGHT (spectrum blocks streamed) into gridDim.x equal intervals; a unit belongs to the
// interval its start weight falls into.  Producer and compute warps call this with identical arguments.
struct StepGeom {
    unsigned lo, hi;                // this CTA's weight interval (32-bit on purpose: 64-bit integer division is ~100 instructions a piece,
};                                  //  and this runs on the latency path of every step; the host checks that the products fit)
__device__ inline unsigned unit_weight(const ProblemDev& P) { return (P.epilogue & BCL_EPILOGUE_SWIGLU) ? 2u * (unsigned)P.q : (unsigned)P.q; }
__device__ inline int unit_count(const ProblemDev& P) { return (P.epilogue & BCL_EPILOGUE_SWIGLU) ? P.p / 2 : P.p; }
__device__ inline StepGeom step_geom(const ChainParams& prm, int step) {
    unsigned W = 0;
    for (int g = prm.step_first[step]; g < prm.step_first[step + 1]; ++g) W += unit_weight(prm.prob[g]) * (unsigned)unit_count(prm.prob[g]);
    StepGeom s;
    s.lo = (W * blockIdx.x + gridDim.x - 1) / gridDim.x;             // ceil: a unit starting exactly on a cut belongs to the later CTA
    s.hi = (W * (blockIdx.x + 1) + gridDim.x - 1) / gridDim.x;
    return s;
}
__device__ inline void seg_range(const ProblemDev& P, unsigned cum, const StepGeom& sg, int& u0, int& u1) {
    const unsigned w = unit_weight(P);
    const unsigned n = (unsigned)unit_count(P);
    // units u with  lo <= cum + u * w < hi
    const unsigned a = sg.lo <= cum ? 0u : (sg.lo - cum + w - 1) / w;
    const unsigned b = sg.hi <= cum ? 0u : (sg.hi - cum + w - 1) / w;
    u0 = (int)(a > n ? n : a);
    u1 = (int)(b > n ? n : b);
}
// spectrum row of slot k of a segment that starts at unit u0
__device__ inline int slot_row(const ProblemDev& P, int u0, int k) {
    return (P.epilogue & BCL_EPILOGUE_SWIGLU) ? u0 + (k >> 1) + (k & 1) * (P.p / 2) : u0 + k;
}

#define BCL_CHAIN_STAMP(slot) do { if (sp.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 32) sp.dbg[step * 8 + (slot)] = clock64(); } while (0)

// DT >= 0: the activation dtype of every problem is known at compile time (x_dtype == y_dtype == DT, no bias, no extra destinations):
// the dtype switches of the loads, the rounding and the stores fold away -- at one token per launch the kernel is latency-bound and
// every instruction on the single finishing warp's path counts.  DT < 0: the general path (any mix, bias promotion).
template <int DT> __device__ __forceinline__ int act_dt(int runtime_dt) { if constexpr (DT >= 0) return DT; else return runtime_dt; }

// rfft of the q input blocks of P.x (one token) into xs[q][B], cooperatively by the NC compute warps (this is warp m): a pass of all
// warps covers NC * G blocks, the raw loads of up to four passes are issued before the first transform (one exposed L2 latency).
template <int B, int DT>
__device__ __forceinline__ void chain_load_x(const ProblemDev& P, float* xs, const float2* tw, float2* fscr, int lane, int m, long long* dbgp, uint64_t* issued,
                                             const FftTw<B>* pre) {
    using C = FftCfg<B>;
    constexpr int PASS = NC * C::G;
    constexpr int MAXP = 4;
    const int q = P.q;
    if (P.x_dtype == BCL_X_SPECTRUM) {
        // the producing step already transformed its rounded output (EMIT_SPECTRUM): fetch the packed fp32 spectrum, no FFT here
        const float4* src = reinterpret_cast<const float4*>(P.x);
        float4* dst = reinterpret_cast<float4*>(xs);
        const int n4 = q * B / 4, tid = m * 32 + lane;
        constexpr int U = 12;                                        // 12 x 16 B x 256 threads = 48 KB per round: q <= 96 blocks of 128 in ONE latency
        for (int i0 = tid; i0 < n4; i0 += U * NC * 32) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (i0 + u * NC * 32 < n4) v[u] = ld_weak_f4(src + i0 + u * NC * 32);
            if (issued != nullptr && i0 == tid && lane == 0) mbar_arrive(issued);        // (lane 0 has the smallest index of its warp)
#pragma unroll
            for (int u = 0; u < U; ++u) if (i0 + u * NC * 32 < n4) dst[i0 + u * NC * 32] = v[u];
        }
        return;
    }
    const int dt = act_dt<DT>(P.x_dtype);
    const int g = lane / C::T, t = lane % C::T;
    float2* scr = fscr + g * C::GROUP_F2;
    for (int b0 = 0; b0 < q; b0 += MAXP * PASS) {
        XRaw<C::V> raw[MAXP];
#pragma unroll
        for (int ps = 0; ps < MAXP; ++ps) {
            const int i = b0 + ps * PASS + m * C::G + g;
            if (i < q) {
                if (dt == BCL_F32) {
                    const float* base = reinterpret_cast<const float*>(P.x) + (int64_t)i * B + 2 * t;
#pragma unroll
                    for (int n1 = 0; n1 < C::V; ++n1) raw[ps].w[n1] = ld_weak_u64(base + 2 * C::T * n1);
                } else {
                    const uint16_t* base = reinterpret_cast<const uint16_t*>(P.x) + (int64_t)i * B + 2 * t;
#pragma unroll
                    for (int n1 = 0; n1 < C::V; ++n1) raw[ps].w[n1].x = ld_weak_u32(base + 2 * C::T * n1);
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < C::V; ++n1) raw[ps].w[n1] = make_uint2(0u, 0u);
            }
        }
        if (issued != nullptr && b0 == 0 && lane == 0) mbar_arrive(issued);   // loads are on their way: the producer may stream
#pragma unroll
        for (int ps = 0; ps < MAXP; ++ps) {
            const int i0 = b0 + ps * PASS + m * C::G;                   // warp-uniform: does this warp have a block in this pass?
            if (i0 < q) {
                float2 a[C::V];
                convert_x_raw<C::V>(raw[ps], dt, a);
                if (dbgp != nullptr && ps == 0 && b0 == 0) { float sacc = 0.f; for (int n1 = 0; n1 < C::V; ++n1) sacc += a[n1].x; if (sacc == 1.2345e30f) xs[0] = sacc; dbgp[6] = clock64(); }
                if (pre != nullptr) group_rfft_packed<B>(a, scr, xs + (size_t)(i0 + g) * B, *pre, t, i0 + g < q);   // lane twiddles in registers
                else group_rfft_packed<B>(a, scr, xs + (size_t)(i0 + g) * B, tw, t, i0 + g < q);
                if (dbgp != nullptr && ps == 0 && b0 == 0) dbgp[7] = clock64();
            }
        }
    }
}

// Partial spectra part[slot][w][B] of `ng` row slots -> part[slot][0][B] holds the sum.  The association order is the SAME however many
// rows a CTA owns: e = ((p0 + p2) + p4) + p6, o = ((p1 + p3) + p5) + p7, sum = e + o -- so a row's bits do not depend on the grid or the
// work split (a member of a shared-input group equals its own forward bit for bit, tests/test_gpu_shared_input.py).  With 2-4 slots
// two warps per slot form e and o here (stage 1) and the finisher adds them; otherwise the finisher forms both itself.
template <int B>
__device__ __forceinline__ void sum_partials_stage1(float* part, int ng, int m, int lane) {
    using LM = LaneMap<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    if (ng < 2) return;
    const int cnt = NC / ng >= 2 ? 2 : 1;                            // warps per slot: one for the even partials, one for the odd ones
    if (cnt < 2 || m >= cnt * ng) return;
    const int slot = m % ng, r = m / ng;
    float* base = part + (size_t)slot * NC * B;
    float sum[CH][CW];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int e = 0; e < CW; ++e) sum[ch][e] = 0.f;
    for (int w = r; w < NC; w += cnt) {
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            const float* src = base + (size_t)w * B + LM::offset(lane, ch);
            if constexpr (CW == 4) { float4 v = *reinterpret_cast<const float4*>(src); sum[ch][0] += v.x; sum[ch][1] += v.y; sum[ch][2] += v.z; sum[ch][3] += v.w; }
            else { float2 v = *reinterpret_cast<const float2*>(src); sum[ch][0] += v.x; sum[ch][1] += v.y; }
        }
    }
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {                                // partial r is in this warp's own read set: overwrite it
        float* dst = base + (size_t)r * B + LM::offset(lane, ch);
        if constexpr (CW == 4) *reinterpret_cast<float4*>(dst) = make_float4(sum[ch][0], sum[ch][1], sum[ch][2], sum[ch][3]);
        else *reinterpret_cast<float2*>(dst) = make_float2(sum[ch][0], sum[ch][1]);
    }
}

// One warp: finish the sums, inverse real FFT of all slots at once (one T-lane group per slot), epilogue.
template <int B, int DT>
__device__ __forceinline__ void finish_slots(const ProblemDev& P, int u0, int slot0, float* part, int nslots, const float2* tw, float2* fscr, int lane, long long* dbgp,
                                             const FftTw<B>* pre) {
    using LM = LaneMap<B>;
    using C = FftCfg<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    const bool staged = nslots >= 2 && NC / nslots >= 2;             // stage 1 left e in partial 0 and o in partial 1
    for (int k = 0; k < nslots; ++k) {
        float* slot = part + (size_t)k * NC * B;
        float se[CH][CW], so[CH][CW];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int e = 0; e < CW; ++e) { se[ch][e] = 0.f; so[ch][e] = 0.f; }
        const int nper = staged ? 1 : NC / 2;                        // partials per parity to add
        const int stride = staged ? 1 : 2;
        for (int w = 0; w < nper; ++w) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const float* s0p = slot + (size_t)(stride * w) * B + LM::offset(lane, ch);
                const float* s1p = slot + (size_t)(stride * w + 1) * B + LM::offset(lane, ch);
                if constexpr (CW == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(s0p), u = *reinterpret_cast<const float4*>(s1p);
                    se[ch][0] += v.x; se[ch][1] += v.y; se[ch][2] += v.z; se[ch][3] += v.w;
                    so[ch][0] += u.x; so[ch][1] += u.y; so[ch][2] += u.z; so[ch][3] += u.w;
                } else {
                    const float2 v = *reinterpret_cast<const float2*>(s0p), u = *reinterpret_cast<const float2*>(s1p);
                    se[ch][0] += v.x; se[ch][1] += v.y; so[ch][0] += u.x; so[ch][1] += u.y;
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            float* dst = slot + LM::offset(lane, ch);
            if constexpr (CW == 4) *reinterpret_cast<float4*>(dst) = make_float4(se[ch][0] + so[ch][0], se[ch][1] + so[ch][1], se[ch][2] + so[ch][2], se[ch][3] + so[ch][3]);
            else *reinterpret_cast<float2*>(dst) = make_float2(se[ch][0] + so[ch][0], se[ch][1] + so[ch][1]);
        }
    }
    __syncwarp();
    const int g = lane / C::T, t = lane % C::T;
    const int pg = min(g, nslots - 1);                               // idle groups shadow the last slot (no stores)
    float2 out[C::V];
    if (dbgp != nullptr) dbgp[16] = clock64();
    if (pre != nullptr) group_irfft_packed<B>(part + (size_t)pg * NC * B, fscr + g * C::GROUP_F2, *pre, t, out);
    else group_irfft_packed<B>(part + (size_t)pg * NC * B, fscr + g * C::GROUP_F2, tw, t, out);
    if (dbgp != nullptr) { float sacc = 0.f; for (int v = 0; v < C::V; ++v) sacc += out[v].x; if (sacc == 1.2345e30f) part[0] = sacc; dbgp[17] = clock64(); }
    const float inv_b = 1.0f / (float)B;
    const int dt = act_dt<DT>(P.y_dtype);                            // (== x_dtype unless a bias promotes: then the general path below)
    const bool swiglu = (P.epilogue & BCL_EPILOGUE_SWIGLU) != 0;
    const bool emit = (P.epilogue & BCL_EPILOGUE_EMIT_SPECTRUM) != 0;
    // staging for EMIT_SPECTRUM inside the slot's (now consumed) partial buffers: [B .. 2B) rounded time samples, [2B .. 3B) their spectrum
    float* tbuf = part + (size_t)(swiglu ? (pg & ~1) : pg) * NC * B + B;
    int out_block;                                                   // output block this lane group contributes to
    if (!swiglu) {
        out_block = u0 + slot0 + g;
        if (g < nslots) {
            if (DT >= 0 || (P.bias == nullptr && P.n_extra == 0 && (P.y_dtype == P.x_dtype || P.x_dtype == BCL_X_SPECTRUM))) {
                // reference epilogue without a bias (src/patch_llama_fft.py:263-266): 1/B, round to x.dtype (the store's conversion)
                const int64_t colbase = (int64_t)out_block * B;
#pragma unroll
                for (int jj = 0; jj < C::OPT; ++jj)
#pragma unroll
                    for (int k2 = 0; k2 < C::T; ++k2) {
                        const int mm = (t * C::OPT + jj) + C::V * k2;
                        const float2 v = make_float2(round_to(out[jj * C::T + k2].x * inv_b, dt), round_to(out[jj * C::T + k2].y * inv_b, dt));
                        store_act2(P.y, colbase + 2 * mm, dt, v);
                        if (emit) *reinterpret_cast<float2*>(tbuf + 2 * mm) = v;
                    }
            } else {
                store_block_epilogue<B>(P, 0, out_block, t, out);
            }
        }
    } else {
        // slots come in (gate, up) pairs on neighbouring lane groups: HF LlamaMLP.forward computes act_fn(gate_proj(x)) * up_proj(x) on
        // the projections' OUTPUTS, i.e. on values already rounded to x.dtype (reference :266); silu is evaluated in fp32 and rounded
        // to x.dtype, so is the product.  Each lane holds sample pairs (2m, 2m+1) of its row; one xor exchange hands the gate lane
        // the up row's EVEN sample and the up lane the gate row's ODD sample, so the two lanes of a pair each finish one of the two
        // samples -- no divergence, half the transcendental work per lane -- and a second exchange reunites the pair for the store.
        out_block = u0 + ((slot0 + g) >> 1);
        const int64_t colbase = (int64_t)out_block * B;
        const bool is_up = (g & 1) != 0;
        // four phases over all V sample pairs, so that the V dependent chains (shuffle -> round -> exp -> divide -> round -> shuffle) overlap:
        // with the predicated store inside one loop the compiler keeps the warp-synchronous shuffles in order and every pair costs a
        // full chain latency (measured: 1.9k cycles for 8 pairs).
        float got[C::V], hv[C::V], ho[C::V];
#pragma unroll
        for (int v = 0; v < C::V; ++v) got[v] = __shfl_xor_sync(0xffffffffu, is_up ? out[v].x : out[v].y, C::T);   // up.x for the even sample, gate.y for the odd one
#pragma unroll
        for (int v = 0; v < C::V; ++v) {
            const float gv = round_to((is_up ? got[v] : out[v].x) * inv_b, dt);
            const float uv = round_to((is_up ? out[v].y : got[v]) * inv_b, dt);
            const float sv = round_to(__fdividef(gv, 1.0f + __expf(-gv)), dt);
            hv[v] = round_to(sv * uv, dt);                           // gate lane: h[2m], up lane: h[2m+1]
        }
#pragma unroll
        for (int v = 0; v < C::V; ++v) ho[v] = __shfl_xor_sync(0xffffffffu, hv[v], C::T);
        if (g < nslots && !is_up) {
#pragma unroll
            for (int jj = 0; jj < C::OPT; ++jj)
#pragma unroll
                for (int k2 = 0; k2 < C::T; ++k2) {
                    const int mm = (t * C::OPT + jj) + C::V * k2;
                    const float2 v = make_float2(hv[jj * C::T + k2], ho[jj * C::T + k2]);
                    store_act2(P.y, colbase + 2 * mm, dt, v);
                    if (emit) *reinterpret_cast<float2*>(tbuf + 2 * mm) = v;
                }
        }
    }
    if (dbgp != nullptr) dbgp[18] = clock64();
    if (emit) {
        // rfft of the ROUNDED output block, exactly what the next step would compute from y (same FFT core, same inputs): the consumer
        // reads it with x_dtype = BCL_X_SPECTRUM and skips its own -- per-CTA redundant -- forward transform.
        __syncwarp();
        const bool owner = g < nslots && (!swiglu || (g & 1) == 0);  // one lane group per output block
        float2 a[C::V];
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) a[n1] = *reinterpret_cast<const float2*>(tbuf + 2 * (C::T * n1 + t));
        __syncwarp();
        if (pre != nullptr) group_rfft_packed<B>(a, fscr + g * C::GROUP_F2, tbuf + B, *pre, t, owner);
        else group_rfft_packed<B>(a, fscr + g * C::GROUP_F2, tbuf + B, tw, t, owner);
        float* ys = reinterpret_cast<float*>(P.y_spectrum);
        const int nblk = swiglu ? (nslots + 1) / 2 : nslots;
        for (int b = 0; b < nblk; ++b) {                             // coalesced write-out, one block (B floats) per iteration
            const int gsrc = swiglu ? 2 * b : b;
            const float* sb = part + (size_t)gsrc * NC * B + 2 * B;
            const int ob = swiglu ? u0 + ((slot0 + gsrc) >> 1) : u0 + slot0 + gsrc;
            for (int e = lane * 4; e < B; e += 128) *reinterpret_cast<float4*>(ys + (size_t)ob * B + e) = *reinterpret_cast<const float4*>(sb + e);
        }
    }
    __syncwarp();
    if (dbgp != nullptr) dbgp[19] = clock64();
}

template <int B, typename SpecT, int DT>
__global__ void __launch_bounds__(THREADS, 1) chain_kernel(const __grid_constant__ ChainParams prm, int S, int qmax) {
    using LM = LaneMap<B>;
    using FC = FftCfg<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    constexpr int RG = FC::G;                                        // row slots per finishing pass
    constexpr int BLOCK_BYTES = B * (int)sizeof(SpecT);
    constexpr int CB = stage_blocks(B);
    constexpr int BPS = CB / NC;                                     // blocks per compute warp per stage
    constexpr int STAGE_BYTES = CB * BLOCK_BYTES;
    constexpr int NW = CW * (int)sizeof(SpecT) / 4;
    static_assert(CB % NC == 0, "stage must split evenly over the compute warps");
    static_assert(sizeof(ChainParams) % 4 == 0, "parameter block is copied word by word");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(16) ChainParams sp;                         // the parameter block, copied once: every later read is an LDS
    __shared__ int2 s_seg[kChainMaxProblems];                        // this CTA's unit range [x, y) of every problem
    Barriers* bars = reinterpret_cast<Barriers*>(smem_raw);
    float2* tw = reinterpret_cast<float2*>(smem_raw + 512);
    unsigned char* ring = smem_raw + 512 + 1024;
    float* part = reinterpret_cast<float*>(ring + (size_t)S * STAGE_BYTES);                         // [RG][NC][B]
    float2* fscr_all = reinterpret_cast<float2*>(part + RG * NC * B);                               // FFT scratch of the NC compute warps
    float* xs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(fscr_all) + (size_t)NC * FC::WARP_SCRATCH_BYTES);   // [qmax][B]
    (void)qmax;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], NC); }
        mbar_init(&bars->xissued, NC);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&prm);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sp);
        for (int i = threadIdx.x; i < (int)(sizeof(ChainParams) / 4); i += THREADS) dst[i] = src[i];
    }
    fill_twiddles<B>(tw, threadIdx.x, THREADS);
    __syncthreads();
    // This CTA's share of every problem of every step, worked out once, here -- before the wait on the previous grid -- so that the
    // path behind a step barrier starts with one shared-memory read instead of a chain of integer divisions.
    if ((int)threadIdx.x < sp.nsteps) {
        const int step = threadIdx.x;
        const StepGeom sg = step_geom(sp, step);
        unsigned cum = 0;
        for (int g = sp.step_first[step]; g < sp.step_first[step + 1]; ++g) {
            int u0, u1;
            seg_range(sp.prob[g], cum, sg, u0, u1);
            cum += unit_weight(sp.prob[g]) * (unsigned)unit_count(sp.prob[g]);
            s_seg[g] = make_int2(u0, u1);
        }
    }
    __syncthreads();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // Measured (clock stamps, tools/chain_timing.py): the SM takes responses in order at ~64 B/clk, so activation loads issued
            // while ~100 KB of bulk copies are in flight land ~1.6k cycles later than alone.  When this grid starts on an SM the moment
            // the previous one leaves it (the usual case: one 200 KB CTA per SM), kernel start and the end of the wait on the previous
            // grid coincide: let the first step's x loads go first.  Bounded: if the previous grid is still running after ~1k cycles
            // there is time to prefetch, which is what the ring is for.
            {
                const long long t0 = clock64();
                const uint32_t xaddr = smem_u32(&bars->xissued);
                while (!mbar_try(xaddr, 0u) && clock64() - t0 < 1000) {}
            }
            int stage = 0; uint32_t ph = 0;
            for (int step = 0; step < sp.nsteps; ++step) {
                for (int g = sp.step_first[step]; g < sp.step_first[step + 1]; ++g) {
                    const ProblemDev& P = sp.prob[g];
                    const int u0 = s_seg[g].x, u1 = s_seg[g].y;
                    const int q = P.q;
                    const int nslots = (u1 - u0) * ((P.epilogue & BCL_EPILOGUE_SWIGLU) ? 2 : 1);
                    for (int k = 0; k < nslots; ++k) {
                        const unsigned char* rowp = reinterpret_cast<const unsigned char*>(P.spec) + (size_t)slot_row(P, u0, k) * q * BLOCK_BYTES;
                        for (int c0 = 0; c0 < q; c0 += CB) {
                            mbar_wait(&bars->empty[stage], ph ^ 1u);
                            const uint32_t bytes = (uint32_t)min(CB, q - c0) * BLOCK_BYTES;
                            mbar_arrive_expect_tx(&bars->full[stage], bytes);
                            bulk_g2s(ring + (size_t)stage * STAGE_BYTES, rowp + (size_t)c0 * BLOCK_BYTES, bytes, &bars->full[stage]);
                            if (++stage == S) { stage = 0; ph ^= 1u; }
                        }
                    }
                }
            }
        }
        return;
    }

    // ===================== compute warps =====================
    const int m = warp - 1;
    float2* fscr = fscr_all + (size_t)m * (FC::WARP_SCRATCH_BYTES / 8);
    int stage = 0; uint32_t ph = 0;
    pdl_wait();                                                      // x may have been written / y may still be read by the previous grid
    // Barrier base of this launch.  Every launch advances the counter by exactly QUANTUM (its gridDim.x * (nsteps - 1) arrivals plus
    // the padding CTA 0 adds once it has passed the last barrier), and every earlier launch of the stream has completed (pdl_wait
    // above), so the counter started this launch at a multiple of QUANTUM; before this CTA's own first arrival fewer than QUANTUM
    // arrivals of THIS launch can have happened: rounding down recovers the base.
    unsigned long long base = 0;
    if (sp.nsteps > 1 && m == 0 && lane == 0) base = ld_acquire_u64(sp.counter) & ~(QUANTUM - 1);
    bool first_x = true;                                             // (a CTA without work in step 0 never signals: the producer's wait is bounded)
    FftTw<B> ftw;                                                    // this lane's FFT twiddles, fetched once: every transform of every step sits on
    if constexpr (B <= 128) fft_tw_init<B>(ftw, tw, lane % FftCfg<B>::T);   // the latency path (B = 256: 25 pairs are too many registers)
    const FftTw<B>* pre = B <= 128 ? &ftw : nullptr;
    for (int step = 0; step < sp.nsteps; ++step) {
        BCL_CHAIN_STAMP(0);
        if (step > 0) {
            if (m == 0 && lane == 0) {
                const unsigned long long target = base + (unsigned long long)step * gridDim.x;
                if (ld_acquire_u64(sp.counter) < target) {
                    const uint64_t t0 = global_ns();
                    uint32_t spin = 0;
                    while (ld_acquire_u64(sp.counter) < target) {
                        if ((++spin & 4095u) == 0u && global_ns() - t0 > 4000000000ull) __trap();   // a CTA is missing: fail, never hang the GPU
                    }
                }
                if (step == sp.nsteps - 1 && blockIdx.x == 0) {       // every CTA has arrived at the last barrier: pad the launch to QUANTUM
                    const unsigned long long pad = QUANTUM - (unsigned long long)gridDim.x * (unsigned long long)(sp.nsteps - 1);
                    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(sp.counter), "l"(pad) : "memory");
                }
            }
            compute_sync();
        }
        BCL_CHAIN_STAMP(1);
        for (int g = sp.step_first[step]; g < sp.step_first[step + 1]; ++g) {
            const int u0 = s_seg[g].x, u1 = s_seg[g].y;
            if (u0 >= u1) continue;
            const ProblemDev P = sp.prob[g];                         // registers: the fields the path below uses
            const int q = P.q;
            const int nslots = (u1 - u0) * ((P.epilogue & BCL_EPILOGUE_SWIGLU) ? 2 : 1);
            chain_load_x<B, DT>(P, xs, tw, fscr, lane, m, (sp.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 32) ? sp.dbg + step * 8 : nullptr,
                                first_x ? &bars->xissued : nullptr, pre);
            first_x = false;
            compute_sync();
            BCL_CHAIN_STAMP(2);
            for (int s0 = 0; s0 < nslots; s0 += RG) {
                const int ng = min(RG, nslots - s0);
                for (int k = 0; k < ng; ++k) {
                    RowAcc<B, 1> acc;
                    acc.clear();
                    for (int c0 = 0; c0 < q; c0 += CB) {
                        mbar_wait(&bars->full[stage], ph);
                        const SpecT* cbuf = reinterpret_cast<const SpecT*>(ring + (size_t)stage * STAGE_BYTES);
                        RawChunk<SpecT, CW> raw[BPS][CH];
                        float xv[BPS][CH][CW];
#pragma unroll
                        for (int u = 0; u < BPS; ++u) {
                            const int bi = m + NC * u, i = c0 + bi;
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                if (i < q) {
                                    const uint32_t* src = reinterpret_cast<const uint32_t*>(cbuf + (size_t)bi * B + LM::offset(lane, ch));
                                    if constexpr (NW == 4) { uint4 t4 = *reinterpret_cast<const uint4*>(src); raw[u][ch].w[0] = t4.x; raw[u][ch].w[1] = t4.y; raw[u][ch].w[2] = t4.z; raw[u][ch].w[3] = t4.w; }
                                    else if constexpr (NW == 2) { uint2 t2 = *reinterpret_cast<const uint2*>(src); raw[u][ch].w[0] = t2.x; raw[u][ch].w[1] = t2.y; }
                                    else { raw[u][ch].w[0] = *src; }
                                    const float* xp = xs + (size_t)i * B + LM::offset(lane, ch);
                                    if constexpr (CW == 4) { float4 t4 = *reinterpret_cast<const float4*>(xp); xv[u][ch][0] = t4.x; xv[u][ch][1] = t4.y; xv[u][ch][2] = t4.z; xv[u][ch][3] = t4.w; }
                                    else { float2 t2 = *reinterpret_cast<const float2*>(xp); xv[u][ch][0] = t2.x; xv[u][ch][1] = t2.y; }
                                } else {
#pragma unroll
                                    for (int e = 0; e < NW; ++e) raw[u][ch].w[e] = 0u;
#pragma unroll
                                    for (int e = 0; e < CW; ++e) xv[u][ch][e] = 0.f;
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < BPS; ++u) {
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                float cv[CW];
                                decode_raw<SpecT, CW>(raw[u][ch], cv);
#pragma unroll
                                for (int e = 0; e < CW; e += 2) {
                                    const float cr = cv[e], ci = cv[e + 1], xre = xv[u][ch][e], xim = xv[u][ch][e + 1];
                                    acc.a[0][ch][e] = fmaf(cr, xre, acc.a[0][ch][e]);
                                    acc.a[0][ch][e] = fmaf(-ci, xim, acc.a[0][ch][e]);
                                    acc.a[0][ch][e + 1] = fmaf(cr, xim, acc.a[0][ch][e + 1]);
                                    acc.a[0][ch][e + 1] = fmaf(ci, xre, acc.a[0][ch][e + 1]);
                                    if (ch == 0 && e == 0) {         // packed slot 0/1 of lane 0 is (DC, Nyquist)
                                        acc.dc[0] = fmaf(cr, xre, acc.dc[0]);
                                        acc.ny[0] = fmaf(ci, xim, acc.ny[0]);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->empty[stage]);
                        if (++stage == S) { stage = 0; ph ^= 1u; }
                    }
                    store_partial<B, 1, NC>(acc, part, k, m, lane);
                }
                compute_sync();
                BCL_CHAIN_STAMP(3);
                if (ng >= 2 && NC / ng >= 2) {
                    sum_partials_stage1<B>(part, ng, m, lane);
                    compute_sync();
                }
                if (m == 0) finish_slots<B, DT>(P, u0, s0, part, ng, tw, fscr, lane, (sp.dbg != nullptr && blockIdx.x == 0 && lane == 0) ? sp.dbg + step * 8 : nullptr, pre);
                compute_sync();                                      // part / xs may be overwritten; warp 1's stores are issued
                BCL_CHAIN_STAMP(4);
            }
        }
        if (step + 1 < sp.nsteps && m == 0 && lane == 0)             // release: this CTA's y (ordered before by the bar.sync above) is visible
            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(sp.counter), "l"(1ull) : "memory");   // GPU-wide before the count moves
        BCL_CHAIN_STAMP(5);
    }
}

size_t smem_bytes(int B, int spec_bytes, int qmax, int S) {
    const size_t fscr = B == 64 ? FftCfg<64>::WARP_SCRATCH_BYTES : (B == 128 ? FftCfg<128>::WARP_SCRATCH_BYTES : FftCfg<256>::WARP_SCRATCH_BYTES);
    const int rg = B == 64 ? FftCfg<64>::G : (B == 128 ? FftCfg<128>::G : FftCfg<256>::G);
    const size_t stage = (size_t)stage_blocks(B) * B * spec_bytes;
    return 512 + 1024 + (size_t)S * stage + (size_t)rg * NC * B * 4 + (size_t)NC * fscr + (size_t)qmax * B * 4;
}

int pick_stages(int B, int spec_dtype, int qmax) {
    const int sb = spec_dtype == BCL_F32 ? 4 : 2;
    const size_t budget = 227 * 1024 - 2048;                        // minus the static shared memory (parameter block)
    const size_t fixed = smem_bytes(B, sb, qmax, 0);
    const size_t stage = (size_t)stage_blocks(B) * B * sb;
    if (fixed + 2 * stage > budget) return 0;
    int S = (int)((budget - fixed) / stage);
    if (S > MAX_STAGES) S = MAX_STAGES;
    return S;
}

// the activation dtype every problem shares (x_dtype == y_dtype, no bias, no extra destinations), or -1
static int uniform_dt(const ChainParams& prm) {
    const int dt = prm.prob[0].y_dtype;
    for (int g = 0; g < prm.nproblems; ++g) {
        const ProblemDev& P = prm.prob[g];
        if ((P.x_dtype != dt && P.x_dtype != BCL_X_SPECTRUM) || P.y_dtype != dt || P.bias != nullptr || P.n_extra != 0) return -1;
    }
    return dt;
}

// Ring stages the busiest CTA can use: all the (row, stage-of-row) chunks of all its steps (the producer runs ahead across steps).
static int stages_wanted(const ChainParams& prm, int B, int grid) {
    const int CB = stage_blocks(B);
    long long total = 0;
    for (int s = 0; s < prm.nsteps; ++s) {
        long long step_chunks = 0;
        for (int g = prm.step_first[s]; g < prm.step_first[s + 1]; ++g) {
            const ProblemDev& P = prm.prob[g];
            const int units = (P.epilogue & BCL_EPILOGUE_SWIGLU) ? P.p / 2 : P.p;
            const int rows_per_unit = (P.epilogue & BCL_EPILOGUE_SWIGLU) ? 2 : 1;
            step_chunks += (long long)((units + grid - 1) / grid + 1) * rows_per_unit * ((P.q + CB - 1) / CB);   // + 1: a share may straddle two problems
        }
        total += step_chunks;
    }
    return (int)(total > MAX_STAGES ? MAX_STAGES : total);
}

template <int B, typename SpecT, int DT>
static cudaError_t launch_t(const ChainParams& prm, int spec_dtype, int qmax, int grid, cudaStream_t st) {
    auto kern = chain_kernel<B, SpecT, DT>;
    int S = pick_stages(B, spec_dtype, qmax);
    if (S < 2) return cudaErrorInvalidConfiguration;
    // A small launch asks only for the stages it can fill: with ~100 KB per CTA two CTAs share an SM, so under programmatic dependent
    // launch the NEXT kernel's CTAs start (parameters, twiddles, spectrum prefetch) while this kernel's are still finishing.
    const int want = stages_wanted(prm, B, grid);
    if (want < S) S = want < 2 ? 2 : want;
    const size_t smem = smem_bytes(B, (int)sizeof(SpecT), qmax, S);
    static int granted[64] = {0};                                    // (one table per template instance)                                    // per device: largest dynamic-smem attribute set so far (a race only repeats the call)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || granted[dev] < (int)smem) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) granted[dev] = (int)smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, prm, S, qmax);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace chain

bool chain_supported(int B, int spec_dtype, int qmax) { return chain::pick_stages(B, spec_dtype, qmax) >= 2; }

cudaError_t launch_chain(const ChainParams& prm, int B, int spec_dtype, int qmax, int grid, cudaStream_t st) {
    const int dt = chain::uniform_dt(prm);                           // 16-bit activations (the decode configuration): dtype-specialised kernels
#define BCL_CH2(BB, TT) (dt == BCL_F16 ? chain::launch_t<BB, TT, BCL_F16>(prm, spec_dtype, qmax, grid, st) : dt == BCL_BF16 ? chain::launch_t<BB, TT, BCL_BF16>(prm, spec_dtype, qmax, grid, st) : chain::launch_t<BB, TT, -1>(prm, spec_dtype, qmax, grid, st))
#define BCL_CH(BB) (spec_dtype == BCL_F32 ? BCL_CH2(BB, float) : BCL_CH2(BB, __nv_bfloat16))
    switch (B) { case 64: return BCL_CH(64); case 128: return BCL_CH(128); default: return BCL_CH(256); }
#undef BCL_CH
#undef BCL_CH2
}

}  // namespace bcl
