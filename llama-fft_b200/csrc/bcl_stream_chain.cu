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
__device__ inline uint64_t global_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct Barriers {
    uint64_t full[MAX_STAGES];      // ring stage filled by TMA              (1 arrival + tx bytes)
    uint64_t empty[MAX_STAGES];     // ring stage released by the compute warps (NC arrivals)
};

// This CTA's share [u0, u1) of the units of problem g of a step (unit = one output block-row, or one (gate, up) row pair).  The
// flattened unit sequence of the step is cut by WEIGHT (spectrum blocks streamed) into gridDim.x equal intervals; a unit belongs to the
// interval its start weight falls into.  Producer and compute warps call this with identical arguments.
struct StepGeom {
    long long lo, hi;               // this CTA's weight interval
};
__device__ inline long long unit_weight(const ProblemDev& P) { return P.epilogue == BCL_EPILOGUE_SWIGLU ? 2LL * P.q : (long long)P.q; }
__device__ inline int unit_count(const ProblemDev& P) { return P.epilogue == BCL_EPILOGUE_SWIGLU ? P.p / 2 : P.p; }
__device__ inline StepGeom step_geom(const ChainParams& prm, int step) {
    long long W = 0;
    for (int g = prm.step_first[step]; g < prm.step_first[step + 1]; ++g) W += unit_weight(prm.prob[g]) * unit_count(prm.prob[g]);
    StepGeom s;
    s.lo = (W * blockIdx.x + gridDim.x - 1) / gridDim.x;             // ceil: a unit starting exactly on a cut belongs to the later CTA
    s.hi = (W * (blockIdx.x + 1) + gridDim.x - 1) / gridDim.x;
    return s;
}
__device__ inline void seg_range(const ProblemDev& P, long long cum, const StepGeom& sg, int& u0, int& u1) {
    const long long w = unit_weight(P);
    const int n = unit_count(P);
    long long a = sg.lo - cum, b = sg.hi - cum;                      // units u with  a <= u * w < b
    a = a <= 0 ? 0 : (a + w - 1) / w;
    b = b <= 0 ? 0 : (b + w - 1) / w;
    u0 = (int)(a > n ? n : a);
    u1 = (int)(b > n ? n : b);
}
// spectrum row of slot k of a segment that starts at unit u0
__device__ inline int slot_row(const ProblemDev& P, int u0, int k) {
    return P.epilogue == BCL_EPILOGUE_SWIGLU ? u0 + (k >> 1) + (k & 1) * (P.p / 2) : u0 + k;
}

#define BCL_CHAIN_STAMP(slot) do { if (prm.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 32) prm.dbg[step * 8 + (slot)] = clock64(); } while (0)

// Sum the NC partial spectra of `nslots` row slots, inverse real FFT of all of them at once (one T-lane group per slot), epilogue.
template <int B>
__device__ inline void finish_slots(const ProblemDev& P, int u0, int slot0, float* part, int nslots, const float2* tw, float2* fscr, int lane) {
    using LM = LaneMap<B>;
    using C = FftCfg<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    for (int k = 0; k < nslots; ++k) {
        float* slot = part + (size_t)k * NC * B;
        float sum[CH][CW];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int e = 0; e < CW; ++e) sum[ch][e] = 0.f;
#pragma unroll
        for (int w = 0; w < NC; ++w) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const float* src = slot + (size_t)w * B + LM::offset(lane, ch);
                if constexpr (CW == 4) { float4 v = *reinterpret_cast<const float4*>(src); sum[ch][0] += v.x; sum[ch][1] += v.y; sum[ch][2] += v.z; sum[ch][3] += v.w; }
                else { float2 v = *reinterpret_cast<const float2*>(src); sum[ch][0] += v.x; sum[ch][1] += v.y; }
            }
        }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            float* dst = slot + LM::offset(lane, ch);
            if constexpr (CW == 4) *reinterpret_cast<float4*>(dst) = make_float4(sum[ch][0], sum[ch][1], sum[ch][2], sum[ch][3]);
            else *reinterpret_cast<float2*>(dst) = make_float2(sum[ch][0], sum[ch][1]);
        }
    }
    __syncwarp();
    const int g = lane / C::T, t = lane % C::T;
    const int pg = min(g, nslots - 1);                               // idle groups shadow the last slot (no stores)
    float2 out[C::V];
    group_irfft_packed<B>(part + (size_t)pg * NC * B, fscr + g * C::GROUP_F2, tw, t, out);
    if (P.epilogue != BCL_EPILOGUE_SWIGLU) {
        if (g < nslots) store_block_epilogue<B>(P, 0, u0 + slot0 + g, t, out);
    } else {
        // slots come in (gate, up) pairs on neighbouring lane groups: HF LlamaMLP.forward computes act_fn(gate_proj(x)) * up_proj(x) on
        // the projections' OUTPUTS, i.e. on values already rounded to x.dtype (reference :266), silu in fp32 rounded to x.dtype, product
        // rounded to x.dtype.
        const float inv_b = 1.0f / (float)B;
        const int dt = P.x_dtype;
        const int64_t colbase = (int64_t)(u0 + ((slot0 + g) >> 1)) * B;
#pragma unroll
        for (int jj = 0; jj < C::OPT; ++jj) {
#pragma unroll
            for (int k2 = 0; k2 < C::T; ++k2) {
                const float2 mine = out[jj * C::T + k2];
                float2 other;
                other.x = __shfl_xor_sync(0xffffffffu, mine.x, C::T);
                other.y = __shfl_xor_sync(0xffffffffu, mine.y, C::T);
                if (g < nslots && (g & 1) == 0) {
                    const float g0 = round_to(mine.x * inv_b, dt), g1 = round_to(mine.y * inv_b, dt);
                    const float u0v = round_to(other.x * inv_b, dt), u1v = round_to(other.y * inv_b, dt);
                    const float s0 = round_to(g0 / (1.0f + expf(-g0)), dt), s1 = round_to(g1 / (1.0f + expf(-g1)), dt);
                    const int m = (t * C::OPT + jj) + C::V * k2;
                    store_act2(P.y, colbase + 2 * m, P.y_dtype, make_float2(s0 * u0v, s1 * u1v));
                }
            }
        }
    }
    __syncwarp();
}

template <int B, typename SpecT>
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

    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fill_twiddles<B>(tw, threadIdx.x, THREADS);
    __syncthreads();
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int step = 0; step < prm.nsteps; ++step) {
                const StepGeom sg = step_geom(prm, step);
                long long cum = 0;
                for (int g = prm.step_first[step]; g < prm.step_first[step + 1]; ++g) {
                    const ProblemDev& P = prm.prob[g];
                    int u0, u1;
                    seg_range(P, cum, sg, u0, u1);
                    cum += unit_weight(P) * unit_count(P);
                    const int q = P.q;
                    const int nslots = (u1 - u0) * (P.epilogue == BCL_EPILOGUE_SWIGLU ? 2 : 1);
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
    if (prm.nsteps > 1 && m == 0 && lane == 0) base = ld_acquire_u64(prm.counter) & ~(QUANTUM - 1);
    for (int step = 0; step < prm.nsteps; ++step) {
        BCL_CHAIN_STAMP(0);
        if (step > 0) {
            if (m == 0 && lane == 0) {
                const unsigned long long target = base + (unsigned long long)step * gridDim.x;
                if (ld_acquire_u64(prm.counter) < target) {
                    const uint64_t t0 = global_ns();
                    uint32_t spin = 0;
                    while (ld_acquire_u64(prm.counter) < target) {
                        if ((++spin & 4095u) == 0u && global_ns() - t0 > 4000000000ull) __trap();   // a CTA is missing: fail, never hang the GPU
                    }
                }
            }
            if (step == prm.nsteps - 1 && blockIdx.x == 0 && m == 0 && lane == 0) {   // every CTA has arrived at the last barrier: pad to QUANTUM
                const unsigned long long pad = QUANTUM - (unsigned long long)gridDim.x * (unsigned long long)(prm.nsteps - 1);
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(prm.counter), "l"(pad) : "memory");
            }
            compute_sync();
        }
        BCL_CHAIN_STAMP(1);
        const StepGeom sg = step_geom(prm, step);
        long long cum = 0;
        for (int g = prm.step_first[step]; g < prm.step_first[step + 1]; ++g) {
            const ProblemDev& P = prm.prob[g];
            int u0, u1;
            seg_range(P, cum, sg, u0, u1);
            cum += unit_weight(P) * unit_count(P);
            if (u0 >= u1) continue;
            const int q = P.q;
            const int nslots = (u1 - u0) * (P.epilogue == BCL_EPILOGUE_SWIGLU ? 2 : 1);
            load_x_spectrum<B, 1, NC, B, true>(P, 0, xs, tw, fscr, lane, m);
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
                if (m == 0) finish_slots<B>(P, u0, s0, part, ng, tw, fscr, lane);
                compute_sync();                                      // part / xs may be overwritten; warp 1's stores are issued
                BCL_CHAIN_STAMP(4);
            }
        }
        if (step + 1 < prm.nsteps && m == 0 && lane == 0) {
            __threadfence();                                         // this CTA's y (ordered before by the bar.sync above) -> visible GPU-wide
            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(prm.counter), "l"(1ull) : "memory");
        }
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
    const size_t budget = 227 * 1024;
    const size_t fixed = smem_bytes(B, sb, qmax, 0);
    const size_t stage = (size_t)stage_blocks(B) * B * sb;
    if (fixed + 2 * stage > budget) return 0;
    int S = (int)((budget - fixed) / stage);
    if (S > MAX_STAGES) S = MAX_STAGES;
    return S;
}

template <int B, typename SpecT>
static cudaError_t launch_t(const ChainParams& prm, int spec_dtype, int qmax, int grid, cudaStream_t st) {
    auto kern = chain_kernel<B, SpecT>;
    const int S = pick_stages(B, spec_dtype, qmax);
    if (S < 2) return cudaErrorInvalidConfiguration;
    const size_t smem = smem_bytes(B, (int)sizeof(SpecT), qmax, S);
    static int granted[64] = {0};                                    // per device: largest dynamic-smem attribute set so far (a race only repeats the call)
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
#define BCL_CH(BB) (spec_dtype == BCL_F32 ? chain::launch_t<BB, float>(prm, spec_dtype, qmax, grid, st) : chain::launch_t<BB, __nv_bfloat16>(prm, spec_dtype, qmax, grid, st))
    switch (B) { case 64: return BCL_CH(64); case 128: return BCL_CH(128); default: return BCL_CH(256); }
#undef BCL_CH
}

}  // namespace bcl
