// bcl_fft.cuh -- register-resident real FFT / inverse real FFT of length-B blocks (B = 64, 128, 256) for sm_100a.
//
// Restates rfft / irfft of the reference (src/patch_llama_fft.py:252,257,263): forward unnormalised with
// e^{-2 pi i k t / B}; inverse = (1/B) * sum, Im(DC) and Im(Nyquist) ignored (they do not exist in the packing).
//
// Design: a length-B real transform is a length-M = B/2 complex transform of z[m] = x[2m] + i x[2m+1] plus a
// pairwise (k, M-k) pass.  The complex transform is a two-level Cooley-Tukey M = V * T:
//     T lanes cooperate on one block, each lane holds V points in registers
//     level 1: radix-V FFT entirely in registers (compile-time twiddles), then a twiddle by w_M^{t k1}
//     transpose through a padded, bank-conflict-free shared scratch
//     level 2: radix-T FFTs in registers
// so a warp transforms G = 32 / T blocks at once with 3-4 __syncwarp()s and ~5x fewer instructions than a
// one-warp-per-block radix-2 shared-memory FFT.   (B, V, T, G): (64, 8, 4, 8), (128, 8, 8, 4), (256, 16, 8, 4).
#pragma once
#include "bcl_common.cuh"

namespace bcl {

template <int B> struct FftCfg {
    static constexpr int M = B / 2;
    static constexpr int V = (M == 128) ? 16 : 8;       // points per lane
    static constexpr int T = M / V;                     // lanes per block
    static constexpr int G = 32 / T;                    // blocks per warp pass
    static constexpr int OPT = V / T;                   // level-2 FFTs per lane
    static constexpr int PITCH = T + 1;                 // padded row pitch of the transpose (float2 elements)
    // float2 elements of scratch per block; the stride is chosen so consecutive blocks start 16 banks apart
    static constexpr int GROUP_F2 = (M == 128) ? (V * PITCH + 8) : (V * PITCH);
    static constexpr int WARP_SCRATCH_BYTES = G * GROUP_F2 * 8;
    static constexpr int NPAIR = (M / 2) / T + 1;       // (k, M-k) pairs per lane in the real pass
};

__host__ __device__ constexpr int brev_n(int k, int n) {            // bit reversal of k in log2(n) bits
    int r = 0;
    for (int b = 1; b < n; b <<= 1) { r = (r << 1) | (k & 1); k >>= 1; }
    return r;
}


// ---------------------------------------------------------------------------------------------------------
// Packed complex arithmetic (BCL_FFT_PACKED=1, an A/B build switch, default OFF): a complex value is an aligned register pair and every
// complex add / sub / scale is ONE instruction (add / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2 on sm_100), a complex multiply two.
// The sign flips, the (re, im) swaps and the scalar broadcasts below are written as pack2() of plain floats: ptxas folds them into
// the instruction's operand modifiers (.F32 broadcast, .LO_HI swap, .NP / .PN per-half negation), they are not instructions.
// MEASURED (round 2, B200): parity green (426 GPU tests), 2.7x fewer floating-point instructions in the transforms -- and SLOWER:
// fp32 decode step 36.9 -> 39.0 us, bf16 28.4 -> 29.2 us.  The transforms are dependent chains in single warps (latency-bound, DESIGN
// 3.3), and a packed operation occupies the fp32 pipe for two passes with a longer dependent-issue latency than the scalar one, so
// halving the instruction count lengthens the critical path.  Packed FMAs pay where the work is throughput-bound and independent:
// the MAC loop of bcl_stream_rw.cu (29.4 -> 28.4 us).
// ---------------------------------------------------------------------------------------------------------
#ifndef BCL_FFT_PACKED
#define BCL_FFT_PACKED 0
#endif
__device__ __forceinline__ uint64_t p2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t p2_sub(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t p2_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t p2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float2 p2_out(uint64_t v) { float2 r; unpack2(v, r.x, r.y); return r; }
__device__ __forceinline__ uint64_t p2_in(float2 a) { return pack2(a.x, a.y); }
#if BCL_FFT_PACKED
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return p2_out(p2_add(p2_in(a), p2_in(b))); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return p2_out(p2_sub(p2_in(a), p2_in(b))); }
__device__ __forceinline__ float2 c_scale(float2 a, float s) { return p2_out(p2_mul(p2_in(a), pack2(s, s))); }
// a * w = (ax wr - ay wi, ax wi + ay wr)
__device__ __forceinline__ float2 c_mul(float2 a, float2 w) { return p2_out(p2_fma(pack2(a.y, a.y), pack2(-w.y, w.x), p2_mul(pack2(a.x, a.x), p2_in(w)))); }
// a * conj(w) = (ax wr + ay wi, -ax wi + ay wr)
__device__ __forceinline__ float2 c_mulc(float2 a, float2 w) { return p2_out(p2_fma(pack2(a.y, a.y), pack2(w.y, w.x), p2_mul(pack2(a.x, a.x), pack2(w.x, -w.y)))); }
#else
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 c_mul(float2 a, float2 w) { return cmul(a, w); }
__device__ __forceinline__ float2 c_mulc(float2 a, float2 w) { return cmulc(a, w); }
#endif

// exp(-2 pi i k / 16), k = 0..7 (compile-time: every use is in a fully unrolled loop)
__device__ __forceinline__ constexpr float tw16_re(int k) {
    constexpr float t[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
    return t[k];
}
__device__ __forceinline__ constexpr float tw16_im(int k) {
    constexpr float t[8] = {0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
    return t[k];
}

// In-register radix-2 DIF FFT of N points (N = 4, 8, 16), fully unrolled.  Output a[brev_n(k, N)] = X[k].
template <int N, bool INV>
__device__ __forceinline__ void fft_reg(float2 (&a)[N]) {
#pragma unroll
    for (int half = N / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int b = 0; b < N / 2; ++b) {
            const int pos = b % half;
            const int i0 = (b / half) * 2 * half + pos, i1 = i0 + half;
            const int k16 = pos * (8 / half);                       // w_{2 half}^{pos} = w_16^{k16}
            const float2 u = a[i0], v = a[i1];
            a[i0] = c_add(u, v);
            const float2 d = c_sub(u, v);
            const float dx = d.x, dy = d.y;
            if (k16 == 0) a[i1] = d;
            else if (k16 == 4) a[i1] = INV ? make_float2(-dy, dx) : make_float2(dy, -dx);        // * (-i)  /  * (+i)
#if BCL_FFT_PACKED
            else if (k16 == 2) {                                     // w = s (1 -+ i): one packed add of the rotated pair, one scale
                constexpr float sq = 0.70710678118654752f;
                a[i1] = INV ? c_scale(c_add(d, make_float2(-dy, dx)), sq) : c_scale(c_add(d, make_float2(dy, -dx)), sq);
            } else if (k16 == 6) {                                   // w = -s (1 +- i)
                constexpr float sq = 0.70710678118654752f;
                a[i1] = INV ? c_scale(c_sub(make_float2(-dy, dx), d), sq) : c_scale(c_sub(make_float2(dy, -dx), d), sq);
            }
#endif
            else {
                const float wr = tw16_re(k16), wi = INV ? -tw16_im(k16) : tw16_im(k16);
                a[i1] = c_mul(d, make_float2(wr, wi));
            }
        }
    }
}

// w_M^e = exp(-2 pi i e / M) for 0 <= e < M from the table tw[k] = exp(-2 pi i k / B), k < M  (w_M = w_B^2)
template <int B>
__device__ __forceinline__ float2 twiddle_m(const float2* tw, int e) {
    constexpr int M = B / 2;
    const int k = 2 * e;
    float2 w = tw[k & (M - 1)];
    if (k & M) { w.x = -w.x; w.y = -w.y; }
    return w;
}

// Per-lane twiddles of the two-level transform, fetched once: w1[k1] = w_M^{t k1} (level-1 twiddle of point k1) and wk[it] = w_B^k of
// the lane's (k, M-k) pairs.  A warp that transforms group after group (the decode kernels' finishers) keeps them in registers
// instead of 12 dependent shared-memory lookups (index arithmetic + load + sign fix) on every transform's critical path.
template <int B> struct FftTw {
    float2 w1[FftCfg<B>::V];
    float2 wk[FftCfg<B>::NPAIR];
};
template <int B>
__device__ __forceinline__ void fft_tw_init(FftTw<B>& f, const float2* tw, int t) {
    using C = FftCfg<B>;
#pragma unroll
    for (int k1 = 0; k1 < C::V; ++k1) f.w1[k1] = twiddle_m<B>(tw, t * k1);
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) f.wk[it] = tw[min(t + C::T * it, C::M / 2)];
}

// ---------------------------------------------------------------------------------------------------------
// Forward.  Lane t (0..T-1) of a block's group passes a[n1] = z[T*n1 + t] = (x[2(T n1 + t)], x[2(T n1 + t)+1]).
// Writes the packed half-spectrum {Re X0, Re X_{B/2}, Re X1, Im X1, ...} to dst[0..B) (shared, linear) if active.
// scr: this block's scratch (GROUP_F2 float2).  All 32 lanes must call (warp-level syncs inside).
// ---------------------------------------------------------------------------------------------------------
template <int B, bool PRE>
__device__ __forceinline__ void group_rfft_packed_impl(float2 (&a)[FftCfg<B>::V], float2* scr, float* dst, const float2* tw, const FftTw<B>* pre, int t, bool active) {
    using C = FftCfg<B>;
    constexpr int M = C::M, V = C::V, T = C::T, OPT = C::OPT, PITCH = C::PITCH;
    fft_reg<V, false>(a);
#pragma unroll
    for (int k1 = 0; k1 < V; ++k1) {
        float2 A = a[brev_n(k1, V)];
        if (k1 > 0) { if constexpr (PRE) A = c_mul(A, pre->w1[k1]); else A = c_mul(A, twiddle_m<B>(tw, t * k1)); }
        scr[k1 * PITCH + t] = A;
    }
    __syncwarp();
    float2 zo[OPT][T];
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        const int k1 = t * OPT + j;
#pragma unroll
        for (int n2 = 0; n2 < T; ++n2) zo[j][n2] = scr[k1 * PITCH + n2];
        fft_reg<T, false>(zo[j]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        const int k1 = t * OPT + j;
#pragma unroll
        for (int k2 = 0; k2 < T; ++k2) scr[k1 + V * k2] = zo[j][brev_n(k2, T)];          // Z[k1 + V k2], natural order
    }
    __syncwarp();
    // real pass: X[k] = (Zk + conj(Zm))/2 - (i/2) w^k (Zk - conj(Zm)),  X[M-k] = conj((Zk + conj(Zm))/2 + (i/2) w^k (Zk - conj(Zm)))
    // Three phases (all loads, all arithmetic, then the predicated stores): with loads inside the per-pair predicate the compiler cannot
    // hoist them and the NPAIR pairs run one shared-memory latency after the other -- a single warp is latency-bound here.
    float2 zk[C::NPAIR], zm[C::NPAIR], wk[C::NPAIR];
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) {
        const int k = min(t + T * it, M / 2);                        // clamped: always a valid pair, surplus lanes recompute pair M/2
        zk[it] = scr[k]; zm[it] = scr[(M - k) & (M - 1)];
        if constexpr (PRE) wk[it] = pre->wk[it]; else wk[it] = tw[k];
    }
    float2 xa[C::NPAIR], xb[C::NPAIR];
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) {
        const float2 cz = make_float2(zm[it].x, -zm[it].y);          // conj(Zm)
        const float2 e = c_scale(c_add(zk[it], cz), 0.5f);
        const float2 d = c_scale(c_sub(zk[it], cz), 0.5f);
        const float2 wd = c_mul(wk[it], d);
        xa[it] = c_add(e, make_float2(wd.y, -wd.x));
        xb[it] = c_sub(make_float2(e.x, -e.y), make_float2(wd.y, wd.x));
    }
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) {
        const int k = t + T * it;
        if (k <= M / 2 && active) {
            if (k == 0) {
                *reinterpret_cast<float2*>(dst) = make_float2(zk[it].x + zk[it].y, zk[it].x - zk[it].y);     // (DC, Nyquist)
            } else {
                *reinterpret_cast<float2*>(dst + 2 * k) = xa[it];
                if (k != M - k) *reinterpret_cast<float2*>(dst + 2 * (M - k)) = xb[it];
            }
        }
    }
    __syncwarp();
}

template <int B>
__device__ __forceinline__ void group_rfft_packed(float2 (&a)[FftCfg<B>::V], float2* scr, float* dst, const float2* tw, int t, bool active) {
    group_rfft_packed_impl<B, false>(a, scr, dst, tw, nullptr, t, active);
}
template <int B>
__device__ __forceinline__ void group_rfft_packed(float2 (&a)[FftCfg<B>::V], float2* scr, float* dst, const FftTw<B>& pre, int t, bool active) {
    group_rfft_packed_impl<B, true>(a, scr, dst, nullptr, &pre, t, active);
}

// ---------------------------------------------------------------------------------------------------------
// Inverse.  src[0..B): packed half-spectrum (shared, linear).  Returns, UNSCALED (caller multiplies by 1/B), for
// j < OPT, k2 < T:   out[j*T + k2] = ( y[2m], y[2m+1] )  with  m = (t*OPT + j) + V*k2.
// ---------------------------------------------------------------------------------------------------------
template <int B, bool PRE>
__device__ __forceinline__ void group_irfft_packed_impl(const float* src, float2* scr, const float2* tw, const FftTw<B>* pre, int t, float2 (&out)[FftCfg<B>::V]) {
    using C = FftCfg<B>;
    constexpr int M = C::M, V = C::V, T = C::T, OPT = C::OPT, PITCH = C::PITCH;
    // E = Xk + conj(Xm);  O = conj(w^k) (Xk - conj(Xm));  Z[k] = E + i O;  Z[M-k] = conj(E) + i conj(O)
    // (loads, arithmetic and predicated stores in three phases, as in the forward pass)
    float2 xk[C::NPAIR], xm[C::NPAIR], wk[C::NPAIR];
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) {
        const int k = min(t + T * it, M / 2);
        const float2 a0 = *reinterpret_cast<const float2*>(src + 2 * k), a1 = *reinterpret_cast<const float2*>(src + 2 * ((M - k) & (M - 1)));
        xk[it] = (k == 0) ? make_float2(a0.x, 0.f) : a0;             // packed slot 0 is (DC, Nyquist), both real
        xm[it] = (k == 0) ? make_float2(a0.y, 0.f) : a1;
        if constexpr (PRE) wk[it] = pre->wk[it]; else wk[it] = tw[k];
    }
    float2 za[C::NPAIR], zb[C::NPAIR];
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) {
        const float2 cx = make_float2(xm[it].x, -xm[it].y);          // conj(Xm)
        const float2 e = c_add(xk[it], cx);
        const float2 d = c_sub(xk[it], cx);
        const float2 o = c_mulc(d, wk[it]);
        za[it] = c_add(e, make_float2(-o.y, o.x));
        zb[it] = c_add(make_float2(e.x, -e.y), make_float2(o.y, o.x));
    }
#pragma unroll
    for (int it = 0; it < C::NPAIR; ++it) {
        const int k = t + T * it;
        if (k <= M / 2) {
            scr[k] = za[it];
            if (k != 0 && k != M - k) scr[M - k] = zb[it];
        }
    }
    __syncwarp();
    float2 a[V];
#pragma unroll
    for (int n1 = 0; n1 < V; ++n1) a[n1] = scr[T * n1 + t];
    fft_reg<V, true>(a);
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < V; ++k1) {
        float2 A = a[brev_n(k1, V)];
        if (k1 > 0) { if constexpr (PRE) A = c_mulc(A, pre->w1[k1]); else A = c_mulc(A, twiddle_m<B>(tw, t * k1)); }
        scr[k1 * PITCH + t] = A;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        const int k1 = t * OPT + j;
        float2 b[T];
#pragma unroll
        for (int n2 = 0; n2 < T; ++n2) b[n2] = scr[k1 * PITCH + n2];
        fft_reg<T, true>(b);
#pragma unroll
        for (int k2 = 0; k2 < T; ++k2) out[j * T + k2] = b[brev_n(k2, T)];
    }
    __syncwarp();
}

template <int B>
__device__ __forceinline__ void group_irfft_packed(const float* src, float2* scr, const float2* tw, int t, float2 (&out)[FftCfg<B>::V]) {
    group_irfft_packed_impl<B, false>(src, scr, tw, nullptr, t, out);
}
template <int B>
__device__ __forceinline__ void group_irfft_packed(const float* src, float2* scr, const FftTw<B>& pre, int t, float2 (&out)[FftCfg<B>::V]) {
    group_irfft_packed_impl<B, true>(src, scr, nullptr, &pre, t, out);
}

// time-domain element pair (2m, 2m+1) -> activation dtype helpers -------------------------------------------------
__device__ __forceinline__ float2 load_act2(const void* base, int64_t elem_off, int dt) {
    if (dt == BCL_F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + elem_off);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(base) + elem_off);
    if (dt == BCL_F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

__device__ __forceinline__ void store_act2(void* base, int64_t elem_off, int dt, float2 v) {
    if (dt == BCL_F32) { *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + elem_off) = v; return; }
    uint32_t w;
    if (dt == BCL_F16) { __half2 h = __floats2half2_rn(v.x, v.y); w = *reinterpret_cast<uint32_t*>(&h); }
    else { __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y); w = *reinterpret_cast<uint32_t*>(&h); }
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(base) + elem_off) = w;
}

}  // namespace bcl
