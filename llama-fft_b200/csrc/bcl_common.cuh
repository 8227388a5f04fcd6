// bcl_common.cuh -- shared device helpers for libbclinear (sm_100a only).
//
// * dtype-generic vector loads/stores for activations (fp32 / fp16 / bf16)
// * the packed half-spectrum convention (DESIGN.md "Data layout")
// * warp-level real FFT / inverse real FFT of one length-B block through a per-warp shared-memory scratch
//
// Math being restated: rfft / irfft of the reference (src/patch_llama_fft.py:252,257,263): forward unnormalised
// with e^{-2 pi i k t / B}, inverse scaled by 1/B and ignoring Im(DC), Im(Nyquist).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/bclinear.h"

namespace bcl {

constexpr int kWarp = 32;

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

__host__ __device__ inline int dtype_size(int dt) { return dt == BCL_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------------------------
// Lane <-> element mapping.  A warp owns one length-B block; lane l owns VPT = B/32 consecutive-ish elements:
//   B = 64 : elements 2l, 2l+1                      (one 8-byte fp32 access per lane, 256 B per warp)
//   B = 128: elements 4l .. 4l+3                    (one 16-byte access, 512 B per warp)
//   B = 256: elements 4l .. 4l+3 and 128+4l .. +3   (two 16-byte accesses, 2 x 512 B per warp)
// The same mapping is used for time samples (x, y) and for packed spectrum reals, so every global access of a
// warp is a run of full 128-byte lines.
// ---------------------------------------------------------------------------------------------------------
template <int B> struct LaneMap {
    static constexpr int VPT = B / 32;
    static constexpr int CH = (VPT >= 4) ? VPT / 4 : 1;   // chunks per lane
    static constexpr int CW = (VPT >= 4) ? 4 : VPT;       // chunk width in elements
    __device__ static inline int offset(int lane, int chunk) { return chunk * 128 + lane * CW; }
};

// load CW activation elements at p (dtype dt) -> float
template <int CW>
__device__ inline void load_act(const void* base, int64_t elem_off, int dt, float (&v)[CW]) {
    if (dt == BCL_F32) {
        const float* p = reinterpret_cast<const float*>(base) + elem_off;
        if constexpr (CW == 4) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
    } else {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + elem_off;
        uint32_t w[CW / 2];
        if constexpr (CW == 4) { uint2 t = *reinterpret_cast<const uint2*>(p); w[0] = t.x; w[1] = t.y; }
        else { w[0] = *reinterpret_cast<const uint32_t*>(p); }
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) {
            if (dt == BCL_F16) {
                float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                v[2 * k] = f.x; v[2 * k + 1] = f.y;
            } else {
                v[2 * k] = __uint_as_float(w[k] << 16);
                v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
            }
        }
    }
}

// round-to-nearest-even to a 16-bit dtype and back (the reference's y32.to(dtype), src/patch_llama_fft.py:266)
__device__ inline float round_to(float v, int dt) {
    if (dt == BCL_F16) return __half2float(__float2half_rn(v));
    if (dt == BCL_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    return v;
}

template <int CW>
__device__ inline void store_act(void* base, int64_t elem_off, int dt, const float (&v)[CW]) {
    if (dt == BCL_F32) {
        float* p = reinterpret_cast<float*>(base) + elem_off;
        if constexpr (CW == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        uint32_t w[CW / 2];
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) {
            if (dt == BCL_F16) {
                __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                w[k] = *reinterpret_cast<uint32_t*>(&h);
            } else {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                w[k] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        uint16_t* p = reinterpret_cast<uint16_t*>(base) + elem_off;
        if constexpr (CW == 4) *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
        else *reinterpret_cast<uint32_t*>(p) = w[0];
    }
}

// spectrum element loads (fp32 or bf16 storage) -> float, CW elements
template <typename SpecT, int CW>
__device__ inline void load_spec(const SpecT* p, float (&v)[CW]) {
    if constexpr (sizeof(SpecT) == 4) {
        if constexpr (CW == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else { float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    } else {
        uint32_t w[CW / 2];
        if constexpr (CW == 4) { uint2 t = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = t.x; w[1] = t.y; }
        else { w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
#pragma unroll
        for (int k = 0; k < CW / 2; ++k) {
            v[2 * k] = __uint_as_float(w[k] << 16);
            v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Twiddles: tw[k] = exp(-2 pi i k / B), k = 0 .. B/2-1, evaluated once per CTA with sincospif (the accurate libm
// routine, not the __sincosf fast path; k/B is a dyadic rational so the argument is exact).
// ---------------------------------------------------------------------------------------------------------
template <int B>
__device__ inline void fill_twiddles(float2* tw, int tid, int nthreads) {
    for (int k = tid; k < B / 2; k += nthreads) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)B, &s, &c);
        tw[k] = make_float2(c, s);
    }
}

__device__ inline float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ inline float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)

// In-place radix-2 decimation-in-frequency FFT of M = B/2 complex points held in shared memory by ONE warp.
// Natural-order input, bit-reversed output.  INV selects e^{+...} twiddles.  Ends with __syncwarp().
template <int B, bool INV>
__device__ inline void warp_cfft_dif(float2* z, const float2* tw, int lane) {
    constexpr int M = B / 2;
    constexpr int LOGM = ilog2(M);
#pragma unroll
    for (int s = 0; s < LOGM; ++s) {
        const int half = M >> (s + 1);
#pragma unroll
        for (int b = lane; b < M / 2; b += kWarp) {
            const int pos = b & (half - 1);
            const int i0 = ((b - pos) << 1) + pos;
            const int i1 = i0 + half;
            const float2 u = z[i0], v = z[i1];
            const float2 w = tw[pos << (s + 1)];               // exp(-2 pi i pos / (2 half)) = tw_B[pos * B/(2 half)]
            const float2 d = make_float2(u.x - v.x, u.y - v.y);
            z[i0] = make_float2(u.x + v.x, u.y + v.y);
            z[i1] = INV ? cmulc(d, w) : cmul(d, w);
        }
        __syncwarp();
    }
}

template <int B> __device__ inline int brev(int k) { return (int)(__brev((unsigned)k) >> (32 - ilog2(B / 2))); }

// Real FFT of one block, in place in shared memory, by one warp.
//   in : buf[0..B) = time samples x[t]
//   out: buf[0..B) = packed half-spectrum {Re X0, Re X_{B/2}, Re X1, Im X1, ..., Re X_{B/2-1}, Im X_{B/2-1}}
template <int B>
__device__ inline void warp_rfft_packed(float* buf, const float2* tw, int lane) {
    constexpr int M = B / 2;
    float2* z = reinterpret_cast<float2*>(buf);                // z[m] = x[2m] + i x[2m+1]
    __syncwarp();
    warp_cfft_dif<B, false>(z, tw, lane);
    constexpr int KI = (M / 2) / kWarp + 1;
    float2 zk[KI], zm[KI];
#pragma unroll
    for (int it = 0; it < KI; ++it) {
        const int k = lane + it * kWarp;
        if (k <= M / 2) { zk[it] = z[brev<B>(k)]; zm[it] = z[brev<B>((M - k) & (M - 1))]; }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < KI; ++it) {
        const int k = lane + it * kWarp;
        if (k > M / 2) continue;
        if (k == 0) {
            buf[0] = zk[it].x + zk[it].y;                       // DC
            buf[1] = zk[it].x - zk[it].y;                       // Nyquist
        } else {
            // X[k] = (Zk + conj(Zm))/2 - (i/2) w^k (Zk - conj(Zm));  X[M-k] = conj((Zk + conj(Zm))/2 + (i/2) w^k (Zk - conj(Zm)))
            const float2 e = make_float2(0.5f * (zk[it].x + zm[it].x), 0.5f * (zk[it].y - zm[it].y));
            const float2 d = make_float2(0.5f * (zk[it].x - zm[it].x), 0.5f * (zk[it].y + zm[it].y));
            const float2 wd = cmul(tw[k], d);                   // w^k * d
            // -i * wd = (wd.y, -wd.x)
            const float2 xk = make_float2(e.x + wd.y, e.y - wd.x);
            const float2 xm = make_float2(e.x - wd.y, -(e.y + wd.x));
            z[k] = xk;
            if (k != M - k) z[M - k] = xm;
        }
    }
    __syncwarp();
}

// Inverse real FFT of one block by one warp.
//   in : buf[0..B) = packed half-spectrum of Y (imaginary parts of DC / Nyquist do not exist in this packing,
//        which is exactly "irfft ignores them", src/patch_llama_fft.py:263)
//   out: buf viewed as float2 z[]: y[2m] = z[brev(m)].x * (1/B), y[2m+1] = z[brev(m)].y * (1/B)   (caller scales)
template <int B>
__device__ inline void warp_irfft_packed(float* buf, const float2* tw, int lane) {
    constexpr int M = B / 2;
    float2* z = reinterpret_cast<float2*>(buf);
    constexpr int KI = (M / 2) / kWarp + 1;
    float2 xk[KI], xm[KI];
    __syncwarp();
#pragma unroll
    for (int it = 0; it < KI; ++it) {
        const int k = lane + it * kWarp;
        if (k <= M / 2) {
            if (k == 0) { xk[it] = make_float2(buf[0], 0.f); xm[it] = make_float2(buf[1], 0.f); }
            else { xk[it] = z[k]; xm[it] = z[M - k]; }
        }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < KI; ++it) {
        const int k = lane + it * kWarp;
        if (k > M / 2) continue;
        // E = Xk + conj(Xm);  O = conj(w^k) (Xk - conj(Xm));  Z[k] = E + i O;  Z[M-k] = conj(E) + i conj(O)     (x2 scale kept, caller divides by B)
        const float2 e = make_float2(xk[it].x + xm[it].x, xk[it].y - xm[it].y);
        const float2 d = make_float2(xk[it].x - xm[it].x, xk[it].y + xm[it].y);
        const float2 o = cmulc(d, tw[k]);                       // d * conj(w^k)
        z[k] = make_float2(e.x - o.y, e.y + o.x);
        if (k != 0 && k != M - k) z[M - k] = make_float2(e.x + o.y, -e.y + o.x);
    }
    __syncwarp();
    warp_cfft_dif<B, true>(z, tw, lane);
}

}  // namespace bcl
