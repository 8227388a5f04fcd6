// bcl_stream_parts.cuh -- device pieces shared by the stream kernels (bcl_stream.cu: plain CTA-cooperative version,
// bcl_stream_ws.cu: warp-specialised TMA pipeline).
#pragma once
#include "bcl_common.cuh"
#include "bcl_fft.cuh"
#include "bcl_kernels.h"

namespace bcl {

// ---------------------------------------------------------------------------------------------------------
// fused forward
// ---------------------------------------------------------------------------------------------------------
struct SegWork { int g, j0, j1, pad; };

// rfft of NT tokens x q input blocks of problem P into xs[NT][q*B] (packed spectra), cooperatively by NWARPS warps
// (this warp is number `warp` of them).  Each warp pass transforms G = 32/T blocks at once (bcl_fft.cuh): the T lanes of
// a block read its time samples straight from global memory into registers (all V loads of a lane are independent, so
// one DRAM latency is exposed per pass), and the packed spectrum lands in shared memory.
// fscr: this warp's FFT scratch (FftCfg<B>::WARP_SCRATCH_BYTES).
// Raw (unconverted) time samples of one lane for one pass: V pairs.  Loading raw words first and converting later keeps
// the V global loads of a pass back to back (one exposed DRAM latency) whatever the activation dtype is.
template <int V> struct XRaw { uint2 w[V]; };

// CG: L2-coherent loads (ld.global.cg) instead of the read-only path: inside the chain kernel x was written by OTHER SMs earlier in the
// same launch, and an address may be reused from step to step, so neither L1 nor the non-coherent cache may serve it.
template <bool CG> __device__ __forceinline__ uint2 ldx_u2(const uint2* p) { if constexpr (CG) return __ldcg(p); else return __ldg(p); }
template <bool CG> __device__ __forceinline__ uint32_t ldx_u1(const uint32_t* p) { if constexpr (CG) return __ldcg(p); else return __ldg(p); }

template <int B, bool CG = false>
__device__ __forceinline__ void load_x_raw(XRaw<FftCfg<B>::V>& r, const ProblemDev& P, int64_t row_off, int i, int t, bool active) {
    using C = FftCfg<B>;
    if (!active) {
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) r.w[n1] = make_uint2(0u, 0u);
        return;
    }
    if (P.x_dtype == BCL_F32) {
        const float* base = reinterpret_cast<const float*>(P.x) + row_off + (int64_t)i * B + 2 * t;
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) r.w[n1] = ldx_u2<CG>(reinterpret_cast<const uint2*>(base + 2 * C::T * n1));
    } else {
        const uint16_t* base = reinterpret_cast<const uint16_t*>(P.x) + row_off + (int64_t)i * B + 2 * t;
#pragma unroll
        for (int n1 = 0; n1 < C::V; ++n1) r.w[n1].x = ldx_u1<CG>(reinterpret_cast<const uint32_t*>(base + 2 * C::T * n1));
    }
}

template <int V>
__device__ __forceinline__ void convert_x_raw(const XRaw<V>& r, int dt, float2 (&a)[V]) {
#pragma unroll
    for (int n1 = 0; n1 < V; ++n1) {
        if (dt == BCL_F32) a[n1] = make_float2(__uint_as_float(r.w[n1].x), __uint_as_float(r.w[n1].y));
        else if (dt == BCL_F16) a[n1] = __half22float2(*reinterpret_cast<const __half2*>(&r.w[n1].x));
        else a[n1] = make_float2(__uint_as_float(r.w[n1].x << 16), __uint_as_float(r.w[n1].x & 0xffff0000u));
    }
}

template <int B, int NT, int NWARPS, int PITCH = B, bool CG = false>
__device__ __forceinline__ void load_x_spectrum(const ProblemDev& P, int64_t n0, float* xs, const float2* tw, float2* fscr, int lane, int warp,
                                                const FftTw<B>* pre = nullptr) {
    using C = FftCfg<B>;
    const int q = P.q;
    const int nb = NT * q;
    const int g = lane / C::T, t = lane % C::T;
    float2* scr = fscr + g * C::GROUP_F2;
    auto locate = [&](int base, int& idx, int& n, int& i, bool& active) {
        idx = base + g; n = idx / q; i = idx - n * q;
        active = idx < nb && (n0 + n) < P.N;
    };
    int idx, n, i; bool active;
    XRaw<C::V> cur, nxt;
    int base = warp * C::G;
    if (base >= nb) return;
    locate(base, idx, n, i, active);
    load_x_raw<B, CG>(cur, P, (n0 + n) * P.x_stride, i, t, active);
    while (true) {
        const int nbase = base + NWARPS * C::G;
        int idx2 = 0, n2 = 0, i2 = 0; bool act2 = false;
        const bool more = nbase < nb;
        if (more) {                                                  // next pass's global loads fly during this pass's FFT
            locate(nbase, idx2, n2, i2, act2);
            load_x_raw<B, CG>(nxt, P, (n0 + n2) * P.x_stride, i2, t, act2);
        }
        float2 a[C::V];
        convert_x_raw<C::V>(cur, P.x_dtype, a);
        if (pre != nullptr) group_rfft_packed<B>(a, scr, xs + (size_t)idx * PITCH, *pre, t, active);   // (resolved at compile time after inlining)
        else group_rfft_packed<B>(a, scr, xs + (size_t)idx * PITCH, tw, t, active);
        if (!more) break;
        cur = nxt; base = nbase; idx = idx2; n = n2; i = i2; active = act2;
    }
}

// ---- streaming pieces -------------------------------------------------------------------------------------------
// A "batch" is UNROLL consecutive spectrum blocks of one row held as raw 32-bit words in registers (fp32: CW words,
// bf16: CW/2 words per chunk).  Two batches ping-pong so the next batch's global loads are always in flight while the
// current one is multiplied -- including across the row boundary and across the row's inverse FFT.
template <typename SpecT, int CW> struct RawChunk { uint32_t w[CW * sizeof(SpecT) / 4]; };

template <typename SpecT, int CW>
__device__ inline void load_raw(const SpecT* p, RawChunk<SpecT, CW>& r) {
    constexpr int NW = CW * sizeof(SpecT) / 4;
    if constexpr (NW == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p));
    else if constexpr (NW == 2) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r.w[0]) : "l"(p));
}

template <typename SpecT, int CW>
__device__ inline void decode_raw(const RawChunk<SpecT, CW>& r, float (&v)[CW]) {
    if constexpr (sizeof(SpecT) == 4) {
#pragma unroll
        for (int e = 0; e < CW; ++e) v[e] = __uint_as_float(r.w[e]);
    } else {
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) { v[2 * k] = __uint_as_float(r.w[k] << 16); v[2 * k + 1] = __uint_as_float(r.w[k] & 0xffff0000u); }
    }
}

template <int B, typename SpecT, int UNROLL> struct Batch { RawChunk<SpecT, LaneMap<B>::CW> c[UNROLL][LaneMap<B>::CH]; };

// Partial row spectrum of one warp: NT tokens x (this lane's packed reals).
template <int B, int NT> struct RowAcc {
    float a[NT][LaneMap<B>::CH][LaneMap<B>::CW];
    float dc[NT], ny[NT];                                          // DC / Nyquist products (real x real), meaningful in lane 0 only
    __device__ inline void clear() {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            dc[n] = 0.f; ny[n] = 0.f;
#pragma unroll
            for (int ch = 0; ch < LaneMap<B>::CH; ++ch)
#pragma unroll
                for (int e = 0; e < LaneMap<B>::CW; ++e) a[n][ch][e] = 0.f;
        }
    }
};

// Split-q decomposition: warp w of the CTA owns input blocks i = w, w+WARPS, w+2*WARPS, ... of EVERY row, so all warps
// stream every row together (consecutive warps read consecutive 512-byte blocks) and the work is balanced at block
// granularity.  A batch is UNROLL of the warp's blocks of one row.
template <int B, typename SpecT, int UNROLL, int WARPS>
__device__ inline void load_batch(Batch<B, SpecT, UNROLL>& bt, const SpecT* crow, int b, int q, int warp, int lane) {
    using LM = LaneMap<B>;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int i = warp + WARPS * (b * UNROLL + u);
        if (i < q) {
#pragma unroll
            for (int ch = 0; ch < LM::CH; ++ch) load_raw<SpecT, LM::CW>(crow + (size_t)i * B + LM::offset(lane, ch), bt.c[u][ch]);
        }
    }
}

template <int B, int NT, typename SpecT, int UNROLL, int WARPS>
__device__ inline void mac_batch(const Batch<B, SpecT, UNROLL>& bt, RowAcc<B, NT>& acc, const float* xs, int b, int q, int warp, int lane) {
    using LM = LaneMap<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
        const int i = warp + WARPS * (b * UNROLL + u);
        if (i < q) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                float cv[CW];
                decode_raw<SpecT, CW>(bt.c[u][ch], cv);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const float* xb = xs + ((size_t)n * q + i) * B + LM::offset(lane, ch);
                    float xv[CW];
                    if constexpr (CW == 4) { float4 t = *reinterpret_cast<const float4*>(xb); xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w; }
                    else { float2 t = *reinterpret_cast<const float2*>(xb); xv[0] = t.x; xv[1] = t.y; }
#pragma unroll
                    for (int e = 0; e < CW; e += 2) {
                        const float cr = cv[e], ci = cv[e + 1], xr = xv[e], xi = xv[e + 1];
                        acc.a[n][ch][e] = fmaf(cr, xr, acc.a[n][ch][e]);
                        acc.a[n][ch][e] = fmaf(-ci, xi, acc.a[n][ch][e]);
                        acc.a[n][ch][e + 1] = fmaf(cr, xi, acc.a[n][ch][e + 1]);
                        acc.a[n][ch][e + 1] = fmaf(ci, xr, acc.a[n][ch][e + 1]);
                        if (ch == 0 && e == 0) {                   // packed slot 0/1 of lane 0 is (DC, Nyquist)
                            acc.dc[n] = fmaf(cr, xr, acc.dc[n]);
                            acc.ny[n] = fmaf(ci, xi, acc.ny[n]);
                        }
                    }
                }
            }
        }
    }
}

// part[(r*NT + n)][w][B]: partial spectra of row r of the current group, token n, written by warp w.
template <int B, int NT, int WARPS>
__device__ inline void store_partial(RowAcc<B, NT>& acc, float* part, int r, int warp, int lane) {
    using LM = LaneMap<B>;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (lane == 0) { acc.a[n][0][0] = acc.dc[n]; acc.a[n][0][1] = acc.ny[n]; }
        float* dst = part + ((size_t)(r * NT + n) * WARPS + warp) * B;
#pragma unroll
        for (int ch = 0; ch < LM::CH; ++ch) {
            if constexpr (LM::CW == 4) *reinterpret_cast<float4*>(dst + LM::offset(lane, ch)) = make_float4(acc.a[n][ch][0], acc.a[n][ch][1], acc.a[n][ch][2], acc.a[n][ch][3]);
            else *reinterpret_cast<float2*>(dst + LM::offset(lane, ch)) = make_float2(acc.a[n][ch][0], acc.a[n][ch][1]);
        }
    }
}

// The reference epilogue for one output block (src/patch_llama_fft.py:263-271): lane t of the block's T-lane group holds
// the UNSCALED inverse-FFT samples out[] (layout of group_irfft_packed); scale by 1/B, round to x.dtype FIRST (:266), then
// add the bias in the promoted dtype (:270-271), store y[n, j*B ...].
template <int B>
__device__ __forceinline__ void store_block_epilogue(const ProblemDev& P, int64_t n, int j, int t, const float2 (&out)[FftCfg<B>::V]) {
    using C = FftCfg<B>;
    const float inv_b = 1.0f / (float)B;
    const int64_t colbase = (int64_t)j * B;
    if (P.bias == nullptr && P.n_extra == 0 && P.y_dtype == P.x_dtype && P.x_dtype != BCL_F32) {
        // the common decode case (16-bit activations, no bias, one destination): the store's round-to-nearest-even conversion IS the
        // reference's y32.to(x.dtype) (:266) -- one address computation, one convert-and-pack and one store per sample pair instead
        // of round / reconvert / promote / dtype switches; every instruction counts on the single finishing warp's path.
        uint16_t* yb = reinterpret_cast<uint16_t*>(P.y) + n * P.y_stride + colbase;
        const bool bf = P.x_dtype == BCL_BF16;
#pragma unroll
        for (int jj = 0; jj < C::OPT; ++jj) {
#pragma unroll
            for (int k2 = 0; k2 < C::T; ++k2) {
                const int m = (t * C::OPT + jj) + C::V * k2;
                const float2 v = out[jj * C::T + k2];
                uint32_t w;
                if (bf) { __nv_bfloat162 h = __floats2bfloat162_rn(v.x * inv_b, v.y * inv_b); w = *reinterpret_cast<uint32_t*>(&h); }
                else { __half2 h = __floats2half2_rn(v.x * inv_b, v.y * inv_b); w = *reinterpret_cast<uint32_t*>(&h); }
                *reinterpret_cast<uint32_t*>(yb + 2 * m) = w;
            }
        }
        return;
    }
#pragma unroll
    for (int jj = 0; jj < C::OPT; ++jj) {
#pragma unroll
        for (int k2 = 0; k2 < C::T; ++k2) {
            const int m = (t * C::OPT + jj) + C::V * k2;
            const int64_t col = colbase + 2 * m;
            float2 v = out[jj * C::T + k2];
            v.x = round_to(v.x * inv_b, P.x_dtype);             // y32.to(x.dtype) first (:266)
            v.y = round_to(v.y * inv_b, P.x_dtype);
            if (P.bias != nullptr) {                            // then + bias in the promoted dtype (:270-271)
                const float2 bv = load_act2(P.bias, col, P.bias_dtype);
                v.x += bv.x; v.y += bv.y;
            }
            store_act2(P.y, n * P.y_stride + col, P.y_dtype, v);
            for (int e = 0; e < P.n_extra; ++e) store_act2(P.y_extra[e], n * P.y_stride + col, P.y_dtype, v);   // fused gather: peers' copies
        }
    }
}

// One warp finishes up to G = 32/T (row, token) pairs at once: sum the NPART partial spectra of each pair (all lanes,
// natural lane layout), then every T-lane group runs the inverse real FFT of one pair, applies the reference epilogue
// (src/patch_llama_fft.py:263-271: 1/B scale, round to x.dtype, + bias in the promoted dtype) and stores y.
//   part : [pair][NPART][B] floats, already offset to this warp's first pair; absolute pair pair0 + p is
//          row (j_first + (pair0+p) / NT), token (n0 + (pair0+p) % NT)
template <int B, int NT, int NPART>
__device__ __forceinline__ void finish_pairs(const ProblemDev& P, int j_first, int64_t n0, float* part, int pair0, int npairs, const float2* tw, float2* fscr, int lane,
                                    const FftTw<B>* pre = nullptr) {
    using LM = LaneMap<B>;
    using C = FftCfg<B>;
    constexpr int CH = LM::CH, CW = LM::CW;
    if constexpr (NPART > 1)                                            // (one partial per pair: it already is the spectrum)
    for (int p = 0; p < npairs; ++p) {
        float* slot = part + (size_t)p * NPART * B;
        float sum[CH][CW];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int e = 0; e < CW; ++e) sum[ch][e] = 0.f;
#pragma unroll
        for (int w = 0; w < NPART; ++w) {
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const float* src = slot + (size_t)w * B + LM::offset(lane, ch);
                if constexpr (CW == 4) { float4 v = *reinterpret_cast<const float4*>(src); sum[ch][0] += v.x; sum[ch][1] += v.y; sum[ch][2] += v.z; sum[ch][3] += v.w; }
                else { float2 v = *reinterpret_cast<const float2*>(src); sum[ch][0] += v.x; sum[ch][1] += v.y; }
            }
        }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {                           // partial 0 of the pair becomes the reduced spectrum
            float* dst = slot + LM::offset(lane, ch);
            if constexpr (CW == 4) *reinterpret_cast<float4*>(dst) = make_float4(sum[ch][0], sum[ch][1], sum[ch][2], sum[ch][3]);
            else *reinterpret_cast<float2*>(dst) = make_float2(sum[ch][0], sum[ch][1]);
        }
    }
    __syncwarp();
    const int g = lane / C::T, t = lane % C::T;
    const int pg = min(g, npairs - 1);                              // idle groups shadow the last pair (no stores)
    float2 out[C::V];
    if (pre != nullptr) group_irfft_packed<B>(part + (size_t)pg * NPART * B, fscr + g * C::GROUP_F2, *pre, t, out);   // (resolved at compile time after inlining)
    else group_irfft_packed<B>(part + (size_t)pg * NPART * B, fscr + g * C::GROUP_F2, tw, t, out);
    const int slot_row = (pair0 + pg) / NT, n = (pair0 + pg) - slot_row * NT;
    if (g < npairs && n0 + n < P.N) store_block_epilogue<B>(P, n0 + n, j_first + slot_row, t, out);
    __syncwarp();
}

}  // namespace bcl
