// bcl_common.cuh -- shared device helpers for libbclinear (sm_100a only).
//
// * dtype-generic vector loads/stores for activations (fp32 / fp16 / bf16)
// * the packed half-spectrum convention (DESIGN.md "Data layout")
// * twiddle table and complex helpers (the FFT itself is in bcl_fft.cuh)
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

// Packed fp32 arithmetic (PTX fma.rn.f32x2 -> SASS FFMA2, sm_100): two independent IEEE fma.rn per issue slot.  pack2 / unpack2 are
// register-pair moves that ptxas folds into the instruction's operand modifiers (R.F32 = broadcast of one register, .LO_HI = swapped
// pair), so pack2(c, c) and pack2(x.y, x.x) cost nothing.  Each half rounds exactly like the scalar fmaf.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(uint64_t& acc, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

}  // namespace bcl
