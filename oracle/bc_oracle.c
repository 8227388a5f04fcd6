/* CPU oracle (plain C) for the block-circulant linear forward.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/ and __graft_entry__.smoke() as an FFT-free cross-check, never by the product.
 *
 * Restates the operator the reference defines, not its FFT evaluation:
 *   W[j*B + r, i*B + s] = c[j,i,(r - s) mod B]     (/root/reference/src/patch_llama_fft.py:86-88,
 *                                                    src/validate_implementation.py:166-182)
 *   y[n, :] = W x[n, :] (+ bias)                    (src/patch_llama_fft.py:259-271)
 * evaluated as a direct circular convolution in double precision.
 *
 * Build: gcc -O3 -shared -fPIC -o oracle/libbc_oracle.so oracle/bc_oracle.c   (oracle/Makefile)
 */
#include <stddef.h>
#include <stdint.h>

/* c: (p,q,B) float32; x: (N, q*B) float64; bias: (p*B) float64 or NULL; y: (N, p*B) float64 */
int bc_oracle_forward_f64(const float* c, int p, int q, int B,
                          const double* x, int64_t N, const double* bias, double* y)
{
    if (!c || !x || !y || p <= 0 || q <= 0 || B <= 0 || N < 0) return -1;
    const int64_t in = (int64_t)q * B, out = (int64_t)p * B;
/* single-threaded on purpose: the toolchain here has no libgomp */
    for (int64_t n = 0; n < N; ++n) {
        for (int j = 0; j < p; ++j) {
            const double* xn = x + n * in;
            double* yn = y + n * out + (int64_t)j * B;
            for (int r = 0; r < B; ++r) {
                double acc = 0.0;
                for (int i = 0; i < q; ++i) {
                    const float* cji = c + ((int64_t)j * q + i) * B;
                    const double* xi = xn + (int64_t)i * B;
                    for (int s = 0; s <= r; ++s) acc += (double)cji[r - s] * xi[s];
                    for (int s = r + 1; s < B; ++s) acc += (double)cji[r - s + B] * xi[s];
                }
                yn[r] = acc + (bias ? bias[(int64_t)j * B + r] : 0.0);
            }
        }
    }
    return 0;
}

/* Frobenius projection of one dense BxB block onto circulants, in double:
 * cout[t] = mean_r W[r, (r - t) mod B]   (/root/reference/src/patch_llama_fft.py:110-119; alpha == 1 analytically) */
int bc_oracle_project_f64(const double* W, int B, double* cout)
{
    if (!W || !cout || B <= 0) return -1;
    for (int t = 0; t < B; ++t) {
        double s = 0.0;
        for (int r = 0; r < B; ++r) s += W[(int64_t)r * B + ((r - t) % B + B) % B];
        cout[t] = s / B;
    }
    return 0;
}
