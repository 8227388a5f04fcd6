// bcl_api.cu -- the extern "C" surface declared in include/bclinear.h.  Argument validation, dispatch to the
// kernel launchers, group-plan encoding.  No allocation, no global mutable state (thread-local error text only).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include <algorithm>
#include "bcl_common.cuh"
#include "bcl_kernels.h"
#include "bcl_gemm.cuh"

namespace {

thread_local char g_err[384] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(BCL_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

bool valid_B(int B) { return B == 64 || B == 128 || B == 256; }
bool valid_dt(int dt) { return dt == BCL_F32 || dt == BCL_F16 || dt == BCL_BF16; }
bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// the reference's type promotion for y_flat + bias (src/patch_llama_fft.py:270-271)
int promote(int a, int b) {
    if (a == b) return a;
    return BCL_F32;   // f32 with anything, and f16 + bf16, promote to f32 in PyTorch
}

int check_problem(const void* x, int x_dtype, int64_t N, int64_t xs, const void* spec, int p, int q, int B,
                  const void* bias, int bias_dtype, const void* y, int y_dtype, int64_t ys, int out_blocks = -1) {
    if (out_blocks < 0) out_blocks = p;                              // SWIGLU problems write (p/2)*B columns
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0 || N < 0) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d N=%lld", p, q, (long long)N);
    if (!valid_dt(x_dtype) || !valid_dt(y_dtype) || (bias && !valid_dt(bias_dtype))) return fail(BCL_ERR_DTYPE, "unknown dtype enum");
    if (N == 0) return BCL_OK;
    if (!x || !spec || !y) return fail(BCL_ERR_ARG, "NULL x/spec/y");
    if (xs < (int64_t)q * B || ys < (int64_t)out_blocks * B) return fail(BCL_ERR_SHAPE, "row stride smaller than the feature dimension");
    const int want = bias ? promote(x_dtype, bias_dtype) : x_dtype;
    if (y_dtype != want) return fail(BCL_ERR_DTYPE, "y_dtype %d but the reference's promotion of x_dtype %d (+bias %d) gives %d", y_dtype, x_dtype, bias ? bias_dtype : -1, want);
    if (!aligned(x, 16) || !aligned(y, 16) || !aligned(spec, 16) || (bias && !aligned(bias, 16)))
        return fail(BCL_ERR_ALIGN, "x, y, spec and bias must be 16-byte aligned");
    if ((xs * bcl::dtype_size(x_dtype)) % 16 || (ys * bcl::dtype_size(y_dtype)) % 16)
        return fail(BCL_ERR_ALIGN, "row strides must be multiples of 16 bytes");
    return BCL_OK;
}

__global__ void noop_kernel() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace

extern "C" {

int bcl_launch_noop(int pdl, void* stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, noop_kernel);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "noop launch");
}

int bcl_abi_version(void) { return BCL_ABI_VERSION; }

const char* bcl_last_error(void) { return g_err; }

size_t bcl_spectrum_bytes(int p, int q, int B, int spec_dtype, int layout) {
    if (!valid_B(B) || p <= 0 || q <= 0) return 0;
    if (layout == BCL_LAYOUT_STREAM) {
        if (spec_dtype != BCL_F32 && spec_dtype != BCL_BF16) return 0;
        return (size_t)p * q * B * (spec_dtype == BCL_F32 ? 4 : 2);
    }
    if (layout == BCL_LAYOUT_GEMM)                                  // f32 = the split (hi, lo) bf16 layout of the fp32-accurate tensor path
        return (spec_dtype == BCL_BF16 || spec_dtype == BCL_F32) ? bcl::gemm_spectrum_bytes(p, q, B, spec_dtype == BCL_F32) : 0;
    if (layout == BCL_LAYOUT_DECODE_TC) return (spec_dtype == BCL_BF16 && B <= 128) ? bcl::stream_tc_spectrum_bytes(p, q, B) : 0;
    return 0;
}

int bcl_precompute_spectrum(const float* c, int p, int q, int B, void* spec, int spec_dtype, int layout, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d", p, q);
    if (!c || !spec) return fail(BCL_ERR_ARG, "NULL c/spec");
    if (!aligned(c, 16) || !aligned(spec, 16)) return fail(BCL_ERR_ALIGN, "c and spec must be 16-byte aligned");
    if (layout == BCL_LAYOUT_STREAM) {
        if (spec_dtype != BCL_F32 && spec_dtype != BCL_BF16) return fail(BCL_ERR_DTYPE, "STREAM spectrum must be f32 or bf16");
        cudaError_t e = bcl::launch_spectrum_stream(c, spec, p, q, B, spec_dtype, (cudaStream_t)stream);
        return e == cudaSuccess ? BCL_OK : cuda_fail(e, "spectrum precompute launch");
    }
    if (layout == BCL_LAYOUT_GEMM) {
        if (spec_dtype != BCL_BF16 && spec_dtype != BCL_F32) return fail(BCL_ERR_DTYPE, "GEMM spectrum must be bf16 or f32 (split bf16 pairs)");
        cudaError_t e = bcl::launch_spectrum_gemm(c, spec, p, q, B, spec_dtype == BCL_F32, (cudaStream_t)stream);
        return e == cudaSuccess ? BCL_OK : cuda_fail(e, "GEMM spectrum precompute launch");
    }
    if (layout == BCL_LAYOUT_DECODE_TC) {
        if (spec_dtype != BCL_BF16 || B > 128) return fail(BCL_ERR_UNSUPPORTED, "DECODE_TC spectra are bf16 with B <= 128");
        cudaError_t e = bcl::launch_spectrum_tc(c, spec, p, q, B, (cudaStream_t)stream);
        return e == cudaSuccess ? BCL_OK : cuda_fail(e, "DECODE_TC spectrum precompute launch");
    }
    return fail(BCL_ERR_ARG, "unknown spectrum layout %d", layout);
}

size_t bcl_workspace_bytes(int64_t N, int p, int q, int B, int spec_dtype, int layout) {
    if (layout == BCL_LAYOUT_GEMM && valid_B(B) && p > 0 && q > 0 && N > 0 && (spec_dtype == BCL_BF16 || spec_dtype == BCL_F32))
        return bcl::gemm_workspace_bytes(N, p, q, B, spec_dtype == BCL_F32);
    return 0;
}

static int forward_impl(const void* x, int x_dtype, int64_t N, int64_t x_row_stride,
                        const void* spec, int spec_dtype, int layout, int p, int q, int B,
                        const void* bias, int bias_dtype, void* const* y_dsts, int n_dst, int y_dtype, int64_t y_row_stride,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (n_dst < 1 || n_dst > 1 + bcl::kMaxExtraDst || !y_dsts) return fail(BCL_ERR_ARG, "need 1..%d destinations", 1 + bcl::kMaxExtraDst);
    void* y = y_dsts[0];
    int rc = check_problem(x, x_dtype, N, x_row_stride, spec, p, q, B, bias, bias_dtype, y, y_dtype, y_row_stride);
    if (rc != BCL_OK || N == 0) return rc;
    bcl::ProblemDev P{x, spec, bias, y, N, x_row_stride, y_row_stride, p, q, x_dtype, y_dtype, bias ? bias_dtype : BCL_F32, n_dst - 1, {}};
    for (int e = 1; e < n_dst; ++e) {
        if (!y_dsts[e] || !aligned(y_dsts[e], 16)) return fail(BCL_ERR_ALIGN, "destination %d is NULL or not 16-byte aligned", e);
        P.y_extra[e - 1] = y_dsts[e];
    }
    if (layout == BCL_LAYOUT_STREAM) {
        if (spec_dtype != BCL_F32 && spec_dtype != BCL_BF16) return fail(BCL_ERR_DTYPE, "STREAM spectrum must be f32 or bf16");
        if (N > 65535LL * 8) return fail(BCL_ERR_SHAPE, "N=%lld exceeds the stream kernel's grid limit; split the batch", (long long)N);
        // One token, one destination: the chain kernel with a single step -- the same split-q MAC, but the spectrum arrives through the
        // TMA / mbarrier ring (1-D bulk copies that start before the wait on the previous grid) instead of ld.global.nc, and the work
        // split needs no dependent global loads.  BCL_SINGLE_VIA_CHAIN=0: the plain kernel (A/B).
        static const bool via_chain = !(getenv("BCL_SINGLE_VIA_CHAIN") && getenv("BCL_SINGLE_VIA_CHAIN")[0] == '0');
        if (via_chain && N == 1 && n_dst == 1 && bcl::chain_supported(B, spec_dtype, q) && (long long)q * p * 1024 <= 0x7fffffffLL) {
            bcl::ChainParams prm{};
            prm.nsteps = 1; prm.nproblems = 1; prm.step_first[0] = 0; prm.step_first[1] = 1;
            prm.prob[0] = P;
            prm.counter = nullptr; prm.dbg = nullptr;
            int dev = 0, sms = 0;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) {
                cudaError_t e = bcl::launch_chain(prm, B, spec_dtype, q, p < sms ? p : sms, (cudaStream_t)stream);
                return e == cudaSuccess ? BCL_OK : cuda_fail(e, "stream forward launch (chain kernel)");
            }
        }
        cudaError_t e = bcl::launch_stream_single(P, B, spec_dtype, (cudaStream_t)stream);
        return e == cudaSuccess ? BCL_OK : cuda_fail(e, "stream forward launch");
    }
    if (layout == BCL_LAYOUT_GEMM) {
        if (spec_dtype != BCL_BF16 && spec_dtype != BCL_F32) return fail(BCL_ERR_DTYPE, "GEMM spectrum must be bf16 or f32 (split bf16 pairs)");
        const int split = spec_dtype == BCL_F32;
        if (N > (int64_t)1 << 30) return fail(BCL_ERR_SHAPE, "N=%lld too large; split the batch", (long long)N);
        if (!workspace || !aligned(workspace, 256)) return fail(BCL_ERR_WORKSPACE, "GEMM layout needs a 256-byte aligned workspace");
        if (bcl::gemm_chunk_tokens(N, p, q, B, workspace_bytes, split) < (N < 128 ? N : 128))
            return fail(BCL_ERR_WORKSPACE, "workspace %zu bytes cannot stage one 128-token chunk (bcl_workspace_bytes gives %zu)", workspace_bytes,
                        bcl::gemm_workspace_bytes(N, p, q, B, split));
        cudaError_t e = bcl::launch_gemm_forward(P, B, spec, workspace, workspace_bytes, split, (cudaStream_t)stream);
        return e == cudaSuccess ? BCL_OK : cuda_fail(e, "GEMM forward launch");
    }
    return fail(BCL_ERR_ARG, "unknown spectrum layout %d", layout);
}

int bcl_forward(const void* x, int x_dtype, int64_t N, int64_t x_row_stride,
                const void* spec, int spec_dtype, int layout, int p, int q, int B,
                const void* bias, int bias_dtype, void* y, int y_dtype, int64_t y_row_stride,
                void* workspace, size_t workspace_bytes, void* stream) {
    void* dst[1] = {y};
    return forward_impl(x, x_dtype, N, x_row_stride, spec, spec_dtype, layout, p, q, B, bias, bias_dtype, dst, 1, y_dtype, y_row_stride,
                        workspace, workspace_bytes, stream);
}

int bcl_forward_bcast(const void* x, int x_dtype, int64_t N, int64_t x_row_stride,
                      const void* spec, int spec_dtype, int layout, int p, int q, int B,
                      const void* bias, int bias_dtype, void* const* y_dsts, int n_dst, int y_dtype, int64_t y_row_stride,
                      void* workspace, size_t workspace_bytes, void* stream) {
    return forward_impl(x, x_dtype, N, x_row_stride, spec, spec_dtype, layout, p, q, B, bias, bias_dtype, y_dsts, n_dst, y_dtype, y_row_stride,
                        workspace, workspace_bytes, stream);
}

int bcl_forward_swiglu(const void* x, int x_dtype, int64_t N, int64_t x_row_stride, const void* spec_gemm, int spec_dtype, int p, int q, int B,
                       void* h, int h_dtype, int64_t h_row_stride, void* workspace, size_t workspace_bytes, void* stream) {
    if (p <= 0 || (p & 1)) return fail(BCL_ERR_SHAPE, "SWIGLU needs an even number of block-rows (gate | up stacked), got p=%d", p);
    int rc = check_problem(x, x_dtype, N, x_row_stride, spec_gemm, p, q, B, nullptr, 0, h, h_dtype, h_row_stride, p / 2);
    if (rc != BCL_OK || N == 0) return rc;
    if (spec_dtype != BCL_BF16 && spec_dtype != BCL_F32) return fail(BCL_ERR_DTYPE, "GEMM spectrum must be bf16 or f32 (split bf16 pairs)");
    const int split = spec_dtype == BCL_F32;
    if (B > 128 || x_dtype == BCL_F32) return fail(BCL_ERR_UNSUPPORTED, "bcl_forward_swiglu: B <= 128 and 16-bit activations");
    if (N > (int64_t)1 << 30) return fail(BCL_ERR_SHAPE, "N=%lld too large; split the batch", (long long)N);
    if (!workspace || !aligned(workspace, 256)) return fail(BCL_ERR_WORKSPACE, "GEMM layout needs a 256-byte aligned workspace");
    if (bcl::gemm_chunk_tokens(N, p, q, B, workspace_bytes, split) < (N < 128 ? N : 128))
        return fail(BCL_ERR_WORKSPACE, "workspace %zu bytes cannot stage one 128-token chunk (bcl_workspace_bytes gives %zu)", workspace_bytes,
                    bcl::gemm_workspace_bytes(N, p, q, B, split));
    bcl::ProblemDev P{x, spec_gemm, nullptr, h, N, x_row_stride, h_row_stride, p, q, x_dtype, h_dtype, BCL_F32, 0, {}};
    cudaError_t e = bcl::launch_gemm_forward_swiglu(P, B, spec_gemm, workspace, workspace_bytes, split, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(BCL_ERR_UNSUPPORTED, "bcl_forward_swiglu: configuration not covered by the tensor-core SWIGLU stage");
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "SWIGLU forward launch");
}

void bcl_prefill_geometry(int p, int q, int B, int* F, int* Kp, int* Np, size_t* wf16_off, size_t* wbf16_off, size_t* winv_off) {
    const bcl::gemm::Geometry g = bcl::gemm::geometry(p, q, B);
    if (F) *F = g.F;
    if (Kp) *Kp = g.Kp;
    if (Np) *Np = g.Np;
    if (wf16_off) *wf16_off = g.wf16_off;
    if (wbf16_off) *wbf16_off = g.wbf16_off;
    if (winv_off) *winv_off = g.winv_off;
}

static int check_impl(int impl) { return (impl < 0 || impl > 2) ? fail(BCL_ERR_ARG, "impl must be 0 (auto), 1 (CUDA-core FFT) or 2 (tensor-core DFT)") : BCL_OK; }

int bcl_prefill_rfft(const void* x, int x_dtype, int64_t N, int64_t x_row_stride, int p, int q, int B, const void* spec_gemm, int impl,
                     void* Xf, int* status, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0 || N < 0 || N > (int64_t)1 << 30) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d N=%lld", p, q, (long long)N);
    if (!valid_dt(x_dtype)) return fail(BCL_ERR_DTYPE, "unknown dtype enum");
    if (check_impl(impl) != BCL_OK) return BCL_ERR_ARG;
    if (N == 0) return BCL_OK;
    if (!x || !Xf || !status) return fail(BCL_ERR_ARG, "NULL x/Xf/status");
    if (!aligned(x, 16) || !aligned(Xf, 128) || (x_row_stride * bcl::dtype_size(x_dtype)) % 16) return fail(BCL_ERR_ALIGN, "x (16 B), Xf (128 B) and the row stride (16 B) must be aligned");
    bcl::ProblemDev P{x, nullptr, nullptr, nullptr, N, x_row_stride, 0, p, q, x_dtype, x_dtype, BCL_F32, 0};
    cudaError_t e = bcl::launch_prefill_stage1(P, B, spec_gemm, 0, (int)N, Xf, status, impl, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(BCL_ERR_UNSUPPORTED, "tensor-core DFT needs B <= 128, 16-bit activations and the GEMM-layout spectrum buffer");
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "prefill rfft launch");
}

int bcl_prefill_gemm(const void* Xf, const void* spec_gemm, void* Yf, int64_t N, int p, int q, int B, int* status, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0 || N < 0 || N > (int64_t)1 << 30) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d N=%lld", p, q, (long long)N);
    if (N == 0) return BCL_OK;
    if (!Xf || !spec_gemm || !Yf || !status) return fail(BCL_ERR_ARG, "NULL Xf/spec/Yf/status");
    if (!aligned(Xf, 128) || !aligned(spec_gemm, 128) || !aligned(Yf, 128)) return fail(BCL_ERR_ALIGN, "Xf, spec and Yf must be 128-byte aligned");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = bcl::launch_gemm_mma(Xf, nullptr, spec_gemm, Yf, nullptr, status, p, q, B, (int)N, sms, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "prefill tcgen05 GEMM launch");
}

int bcl_prefill_irfft(const void* Yf, int64_t N, int p, int q, int B, const void* spec_gemm, int impl, int x_dtype,
                      const void* bias, int bias_dtype, void* y, int y_dtype, int64_t y_row_stride, int* status, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0 || N < 0 || N > (int64_t)1 << 30) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d N=%lld", p, q, (long long)N);
    if (!valid_dt(x_dtype) || !valid_dt(y_dtype) || (bias && !valid_dt(bias_dtype))) return fail(BCL_ERR_DTYPE, "unknown dtype enum");
    if (check_impl(impl) != BCL_OK) return BCL_ERR_ARG;
    if (N == 0) return BCL_OK;
    if (!Yf || !y || !status) return fail(BCL_ERR_ARG, "NULL Yf/y/status");
    const int want = bias ? promote(x_dtype, bias_dtype) : x_dtype;
    if (y_dtype != want) return fail(BCL_ERR_DTYPE, "y_dtype %d but the reference's promotion gives %d", y_dtype, want);
    if (!aligned(Yf, 128) || !aligned(y, 16) || (y_row_stride * bcl::dtype_size(y_dtype)) % 16) return fail(BCL_ERR_ALIGN, "Yf (128 B), y (16 B) and the row stride (16 B) must be aligned");
    bcl::ProblemDev P{nullptr, nullptr, bias, y, N, 0, y_row_stride, p, q, x_dtype, y_dtype, bias ? bias_dtype : BCL_F32, 0};
    cudaError_t e = bcl::launch_prefill_stage3(P, B, spec_gemm, 0, (int)N, Yf, status, impl, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(BCL_ERR_UNSUPPORTED, "tensor-core inverse DFT needs B <= 128, 16-bit activations, y_dtype == x_dtype and the GEMM-layout spectrum buffer");
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "prefill irfft launch");
}

static constexpr int kMaxDynPerCta = 16;      // cap of dynamic-tail items per CTA (plan size)

size_t bcl_group_plan_bytes(int count, int num_ctas) {
    if (count <= 0 || num_ctas <= 0) return 0;
    // segments: every CTA boundary and every problem boundary can start one; plus the items of the dynamic tail and its two counters
    size_t segs = (size_t)count + (size_t)num_ctas + 2 * (size_t)kMaxDynPerCta * (size_t)num_ctas;      // (an item is 32 bytes)
    size_t b = sizeof(bcl::GroupPlanHeader);
    b = (b + 15) & ~(size_t)15; b += (size_t)count * sizeof(bcl::ProblemDev);
    b = (b + 15) & ~(size_t)15; b += ((size_t)num_ctas + 1) * sizeof(int);
    b = (b + 15) & ~(size_t)15; b += segs * 16;
    b = (b + 15) & ~(size_t)15; b += 16 + 8 * (size_t)count;          // launch counters, ready marks and problem indices of the tail
    return (b + 255) & ~(size_t)255;
}

static int device_sms() {                                               // the dynamic tail needs every CTA of the grid resident at once
    static const int sms = [] { int dev = 0, n = 0; if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; } return n; }();
    return sms;
}

// Dynamic tail of a grouped plan (see bcl_group_plan_encode_layout): which rows [pstat[g], p) of which problems are cut into items.
// A pure function of the shapes, shared by the encoder and by bcl_group_plan_bytes_for.
struct DynTail { int pct = 0, rows = 4, items = 0, K = 0, xstride = 0; std::vector<int> order, pstat; };
static DynTail select_dyn_tail(const bcl_problem* problems, int count, int B, int num_ctas, bool eligible, int arow) {
    DynTail t;
    t.order.resize(count); t.pstat.resize(count);
    int64_t total_rows = 0, total_w = 0;
    for (int g = 0; g < count; ++g) { t.order[g] = g; t.pstat[g] = problems[g].p; total_rows += problems[g].p; total_w += (int64_t)problems[g].p * (problems[g].q + arow); }
    int pct = 0, rows = 4;                                          // off unless requested: measured slower than the static plan (see the encoder)
    if (const char* dc = getenv("BCL_GROUP_DYNAMIC")) { int a = 0, b = 4; int n = sscanf(dc, "%d,%d", &a, &b); if (n >= 1 && a >= 0 && a <= 90) pct = a; if (n == 2 && b >= 1 && b <= 64) rows = b; }
    if (!eligible || num_ctas > device_sms() || total_rows < (int64_t)4 * num_ctas) pct = 0;
    if (pct == 0) return t;
    std::stable_sort(t.order.begin(), t.order.end(), [&](int a, int b) { return problems[a].q > problems[b].q; });
    for (;;) {
        for (int g = 0; g < count; ++g) t.pstat[g] = problems[g].p;
        int64_t want = total_w * pct / 100;
        t.items = 0; t.K = 0; int qm = 0;
        for (int o = count - 1; o >= 0 && want > 0; --o) {
            const int g = t.order[o], p = problems[g].p, wr = problems[g].q + arow;
            int take = (int)std::min<int64_t>(p, (want + wr - 1) / wr);
            take = std::min(p, (take + rows - 1) / rows * rows);
            t.pstat[g] = p - take;
            want -= (int64_t)take * wr;
            t.items += (take + rows - 1) / rows;
            ++t.K; qm = std::max(qm, problems[g].q);
        }
        t.xstride = qm * B;
        if (t.items <= kMaxDynPerCta * num_ctas) break;
        rows *= 2;
    }
    t.pct = pct; t.rows = rows;
    return t;
}

size_t bcl_group_plan_bytes_for(const bcl_problem* problems, int count, int B, int spec_dtype, int num_ctas) {
    size_t b = bcl_group_plan_bytes(count, num_ctas);
    if (!b || !problems || !valid_B(B)) return b;
    int qmax = 0;
    for (int g = 0; g < count; ++g) { if (problems[g].N != 1 || problems[g].p <= 0 || problems[g].q <= 0) return b; qmax = std::max(qmax, problems[g].q); }
    const bool eligible = bcl::stream_ws_supported(B, spec_dtype, 1, qmax) && !bcl::stream_rw_supported(B, spec_dtype, 1, qmax);
    const DynTail t = select_dyn_tail(problems, count, B, num_ctas, eligible, 0);
    return b + (size_t)t.K * t.xstride * sizeof(float);
}

int bcl_group_plan_encode(const bcl_problem* problems, int count, int B, int spec_dtype, int num_ctas,
                          void* host_plan, size_t host_plan_bytes) {
    return bcl_group_plan_encode_layout(problems, count, B, spec_dtype, BCL_LAYOUT_STREAM, num_ctas, host_plan, host_plan_bytes);
}

int bcl_group_plan_encode_layout(const bcl_problem* problems, int count, int B, int spec_dtype, int layout, int num_ctas,
                                 void* host_plan, size_t host_plan_bytes) {
    if (layout != BCL_LAYOUT_STREAM && layout != BCL_LAYOUT_DECODE_TC) return fail(BCL_ERR_ARG, "grouped plans take STREAM or DECODE_TC spectra");
    if (!problems || !host_plan || count <= 0 || num_ctas <= 0) return fail(BCL_ERR_ARG, "bad plan arguments");
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (spec_dtype != BCL_F32 && spec_dtype != BCL_BF16) return fail(BCL_ERR_DTYPE, "STREAM spectrum must be f32 or bf16");
    if (layout == BCL_LAYOUT_DECODE_TC && spec_dtype != BCL_BF16) return fail(BCL_ERR_UNSUPPORTED, "DECODE_TC spectra are bf16");
    const size_t need = bcl_group_plan_bytes(count, num_ctas);
    if (host_plan_bytes < need) return fail(BCL_ERR_WORKSPACE, "plan buffer %zu < %zu bytes", host_plan_bytes, need);
    int64_t maxN = 0; int qmax = 0; int64_t total_w = 0;
    for (int g = 0; g < count; ++g) {
        const bcl_problem& P = problems[g];
        int rc = check_problem(P.x, P.x_dtype, P.N, P.x_row_stride, P.spec, P.p, P.q, B, P.bias, P.bias_dtype, P.y, P.y_dtype, P.y_row_stride);
        if (rc != BCL_OK) return rc;
        if (P.N < 1 || P.N > 8) return fail(BCL_ERR_SHAPE, "grouped problems need 1 <= N <= 8 (problem %d has %lld)", g, (long long)P.N);
        if (P.N > maxN) maxN = P.N;
        if (P.q > qmax) qmax = P.q;
        total_w += (int64_t)P.p * P.q;                                  // (re-weighted below once the kernel variant is known)
    }
    int nt = 1; while (nt < maxN) nt *= 2;
    const char* force = getenv("BCL_GROUP_VARIANT");                 // "plain" | "ws": A/B switch for profiling only
    int variant = bcl::stream_ws_supported(B, spec_dtype, nt, qmax) ? 1 : 0;
    if (bcl::stream_rw_supported(B, spec_dtype, nt, qmax)) variant = 2;                      // bf16 spectra: row-per-warp kernel (its own shape limits)
    if (force && !strcmp(force, "plain")) variant = 0;
    if (force && !strcmp(force, "ws") && variant == 2 && bcl::stream_ws_supported(B, spec_dtype, nt, qmax)) variant = 1;
    const int gran = layout == BCL_LAYOUT_DECODE_TC ? 8 : 1;        // block-rows per work unit (DECODE_TC: 8-row groups of the operand image)
    if (layout == BCL_LAYOUT_DECODE_TC) {
        if (!bcl::stream_tc_supported(B, spec_dtype, nt, qmax))
            return fail(BCL_ERR_UNSUPPORTED, "tensor-core decode kernel: one token per problem, B <= 128, q <= 128 (got N <= %d, B = %d, qmax = %d)", nt, B, qmax);
        variant = 3;
    }
    if (variant == 0 && bcl::stream_smem_bytes(B, nt, qmax, 8) > 200 * 1024)
        return fail(BCL_ERR_UNSUPPORTED, "x spectrum tile (N=%d, q=%d, B=%d) exceeds shared memory; call bcl_forward per problem", nt, qmax, B);

    unsigned char* base = reinterpret_cast<unsigned char*>(host_plan);
    memset(base, 0, need);
    bcl::GroupPlanHeader H{};
    H.magic = bcl::kPlanMagic; H.count = count; H.num_ctas = num_ctas; H.B = B; H.spec_dtype = spec_dtype; H.nt = nt; H.qmax = qmax; H.variant = variant;
    H.reserved = (getenv("BCL_GROUP_INDEPENDENT") && getenv("BCL_GROUP_INDEPENDENT")[0] == '1') ? 1 : 0;
    size_t off = sizeof(bcl::GroupPlanHeader);
    off = (off + 15) & ~(size_t)15; H.off_problems = off; off += (size_t)count * sizeof(bcl::ProblemDev);
    off = (off + 15) & ~(size_t)15; H.off_cta_ptr = off; off += ((size_t)num_ctas + 1) * sizeof(int);
    off = (off + 15) & ~(size_t)15; H.off_segs = off;
    H.total_bytes = need;

    bcl::ProblemDev* pd = reinterpret_cast<bcl::ProblemDev*>(base + H.off_problems);
    for (int g = 0; g < count; ++g) {
        const bcl_problem& P = problems[g];
        pd[g] = bcl::ProblemDev{P.x, P.spec, P.bias, P.y, P.N, P.x_row_stride, P.y_row_stride, P.p, P.q, P.x_dtype, P.y_dtype,
                                P.bias ? P.bias_dtype : BCL_F32, 0, {}};
    }
    // byte-balanced partition of the flattened (problem, row) sequence: CTA b owns the rows whose cumulative
    // weight (q per row = bytes streamed) falls in [b, b+1) * total / num_ctas.
    int* cta_ptr = reinterpret_cast<int*>(base + H.off_cta_ptr);
    int* segs = reinterpret_cast<int*>(base + H.off_segs);
    // Cost model of a CTA's share, in units of one spectrum block streamed from HBM: q blocks per row (+ `arow` per row, + `aseg` per
    // problem start; BCL_PLAN_COST="arow,aseg").  Default: bytes only.  Measured on B200 (224 Llama-2-7B projections): per-CTA clock
    // counts do correlate with rows and segments (+600 cycles per 4-row group and +2.1k per segment with bf16 spectra, the slowest CTA
    // 15 % behind the median), but shifting bytes by those costs made the launch SLOWER (fp32 36.3 -> 37.6 us, bf16 29.3 -> 32.9 us):
    // the launch is bound by the aggregate HBM stream, a CTA that finishes its bytes early leaves its bandwidth to the others, and the
    // one with the most bytes finishes last whatever else it does.
    int arow = 0, aseg = 0;
    if (const char* pc = getenv("BCL_PLAN_COST")) { int a = 0, b = 0; if (sscanf(pc, "%d,%d", &a, &b) == 2 && a >= 0 && b >= 0) { arow = a; aseg = b; } }
    total_w = 0;
    for (int g = 0; g < count; ++g) total_w += (int64_t)problems[g].p * (problems[g].q + arow) + aseg;
    // Dynamic tail (the warp-specialised fp32 kernel; BCL_GROUP_DYNAMIC="percent[,rows]", 0 = off): equal-byte shares finish up to
    // ~6 us apart on B200 (the HBM system does not serve 148 streams evenly, and differently every launch), and a dependent next launch
    // waits for the slowest.  So the last `percent` of the weight is not assigned: it is cut into items of `rows` block-rows that the
    // CTAs draw from a ticket counter once their static share is streamed.  Every item costs its CTA an rfft of the problem's input, so
    // the problems are walked widest input first and the tail is taken from the narrow ones.
    // Every item switches its CTA to another problem; the rfft of that problem's input is computed once per launch into a cache behind
    // the plan (bcl_stream_ws.cu), so the device plan must have room for it: size the buffers with bcl_group_plan_bytes_for.  A buffer
    // sized by bcl_group_plan_bytes alone gets a plan without the tail (every row in a static share).
    DynTail dt = select_dyn_tail(problems, count, B, num_ctas, variant == 1 && maxN == 1 && !(H.reserved & 1) && aseg == 0, arow);
    if (dt.pct > 0 && host_plan_bytes < need + (size_t)dt.K * dt.xstride * sizeof(float))
        dt = select_dyn_tail(problems, count, B, num_ctas, false, arow);
    const std::vector<int>& order = dt.order;
    const std::vector<int>& pstat = dt.pstat;
    const int dyn_pct = dt.pct, dyn_rows = dt.rows;
    if (dyn_pct > 0) {
        total_w = 0;
        for (int g = 0; g < count; ++g) total_w += (int64_t)pstat[g] * (problems[g].q + arow);
    }
    int nseg = 0, cta = 0;
    int64_t cum = 0;
    cta_ptr[0] = 0;
    for (int o = 0; o < count; ++o) {
        const int g = order[o];
        const int p = pstat[g], q = problems[g].q;
        const int wr = q + arow;                                     // weight of one row
        int j = 0;
        if (p) cum += aseg;
        while (j < p) {
            // rows of this problem that still belong to the current CTA
            const int64_t limit = (total_w * (cta + 1) + num_ctas - 1) / num_ctas;   // ceil
            int64_t room = limit - cum;
            int take = (int)((room + wr - 1) / wr);
            if (take < 1) take = 1;
            take = (take + gran - 1) / gran * gran;
            if (take > p - j) take = p - j;
            segs[4 * nseg + 0] = g; segs[4 * nseg + 1] = j; segs[4 * nseg + 2] = j + take; segs[4 * nseg + 3] = 0;
            ++nseg;
            j += take; cum += (int64_t)take * wr;
            while (cta < num_ctas - 1 && cum >= (total_w * (cta + 1) + num_ctas - 1) / num_ctas) { ++cta; cta_ptr[cta] = nseg; }
        }
    }
    while (cta < num_ctas) { ++cta; cta_ptr[cta] = nseg; }
    H.num_segs = nseg;
    int nd = 0, K = 0;
    off = H.off_segs + (size_t)(nseg + 2 * dt.items) * 16;
    off = (off + 15) & ~(size_t)15; H.off_ctr = off;                 // counters: zero here; the kernel leaves them zero after every launch
    int* pool_g = reinterpret_cast<int*>(base + H.off_ctr) + 4 + dt.K;
    for (int o = 0; o < count && dyn_pct > 0; ++o) {                 // the tail's items, in the order the tickets hand them out
        const int g = order[o];
        if (pstat[g] == problems[g].p) continue;
        for (int j = pstat[g]; j < problems[g].p; j += dyn_rows, ++nd) {
            int* it = segs + 4 * nseg + 8 * nd;                      // bcl::ws::DynItem: g, j0, j1, pool index, q, pad, spectrum row j0
            it[0] = g; it[1] = j; it[2] = std::min(problems[g].p, j + dyn_rows); it[3] = K; it[4] = problems[g].q; it[5] = 0;
            const unsigned char* rowp = reinterpret_cast<const unsigned char*>(problems[g].spec) + (size_t)j * problems[g].q * B * (spec_dtype == BCL_F32 ? 4 : 2);
            memcpy(it + 6, &rowp, 8);
        }
        pool_g[K++] = g;
    }
    H.num_dyn = nd; H.dyn_rows = dyn_rows; H.dyn_problems = K; H.dyn_xstride = dt.xstride;
    H.off_xcache = need;                                             // behind everything bcl_group_plan_bytes counts (256-byte aligned)
    H.total_bytes = need + (size_t)K * dt.xstride * sizeof(float);
    memcpy(base, &H, sizeof(H));
    return BCL_OK;
}

int bcl_forward_grouped(const void* dev_plan, const void* host_plan, void* stream) {
    if (!dev_plan || !host_plan) return fail(BCL_ERR_ARG, "NULL plan");
    bcl::GroupPlanHeader H;
    memcpy(&H, host_plan, sizeof(H));
    if (H.magic != bcl::kPlanMagic) return fail(BCL_ERR_ARG, "host_plan was not written by bcl_group_plan_encode");
    if (!aligned(dev_plan, 16)) return fail(BCL_ERR_ALIGN, "dev_plan must be 16-byte aligned");
    cudaError_t e = H.variant == 3 ? bcl::launch_stream_tc(H, dev_plan, (cudaStream_t)stream)
                  : H.variant == 2 ? bcl::launch_stream_rw(H, dev_plan, (cudaStream_t)stream)
                  : H.variant == 1 ? bcl::launch_stream_ws(H, dev_plan, (cudaStream_t)stream)
                                   : bcl::launch_stream_grouped(H, dev_plan, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "grouped forward launch");
}

int bcl_forward_chain(const bcl_problem* problems, const int* step_first, int nsteps, int B, int spec_dtype, void* barrier_counter, void* stream) {
    if (!problems || !step_first) return fail(BCL_ERR_ARG, "NULL problems / step_first");
    if (nsteps < 1 || nsteps > BCL_CHAIN_MAX_STEPS) return fail(BCL_ERR_ARG, "nsteps %d not in 1..%d", nsteps, BCL_CHAIN_MAX_STEPS);
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (spec_dtype != BCL_F32 && spec_dtype != BCL_BF16) return fail(BCL_ERR_DTYPE, "STREAM spectrum must be f32 or bf16");
    const int count = step_first[nsteps];
    if (step_first[0] != 0 || count < 1 || count > BCL_CHAIN_MAX_PROBLEMS) return fail(BCL_ERR_ARG, "step_first must run from 0 to 1..%d problems", BCL_CHAIN_MAX_PROBLEMS);
    if (nsteps > 1 && (!barrier_counter || !aligned(barrier_counter, 8))) return fail(BCL_ERR_ARG, "a chain of several steps needs an 8-byte aligned barrier counter");
    bcl::ChainParams prm{};
    prm.nsteps = nsteps; prm.nproblems = count;
    int qmax = 0, max_units = 1;
    const long long sms_guess = 512;                                 // upper bound of the grid for the overflow check below
    for (int s = 0; s < nsteps; ++s) {
        if (step_first[s + 1] <= step_first[s]) return fail(BCL_ERR_ARG, "step %d is empty", s);
        prm.step_first[s] = step_first[s];
        int units = 0;
        for (int g = step_first[s]; g < step_first[s + 1]; ++g) {
            const bcl_problem& P = problems[g];
            if (P.epilogue & ~(BCL_EPILOGUE_SWIGLU | BCL_EPILOGUE_EMIT_SPECTRUM)) return fail(BCL_ERR_ARG, "problem %d: unknown epilogue %d", g, P.epilogue);
            const bool swiglu = (P.epilogue & BCL_EPILOGUE_SWIGLU) != 0, emit = (P.epilogue & BCL_EPILOGUE_EMIT_SPECTRUM) != 0;
            if (swiglu && ((P.p & 1) || P.bias)) return fail(BCL_ERR_SHAPE, "problem %d: SWIGLU needs an even number of block-rows (gate | up) and no bias", g);
            if (emit && (P.bias || !P.y_spectrum || !aligned(P.y_spectrum, 16))) return fail(BCL_ERR_ARG, "problem %d: EMIT_SPECTRUM needs a 16-byte aligned y_spectrum and no bias", g);
            const bool xspec = P.x_dtype == BCL_X_SPECTRUM;          // x is the fp32 packed spectrum an earlier step emitted
            if (xspec && P.bias) return fail(BCL_ERR_DTYPE, "problem %d: a spectrum input takes no bias", g);
            int rc = check_problem(P.x, xspec ? P.y_dtype : P.x_dtype, P.N, xspec ? (int64_t)P.q * B : P.x_row_stride, P.spec, P.p, P.q, B, P.bias, P.bias_dtype, P.y,
                                   P.y_dtype, P.y_row_stride, swiglu ? P.p / 2 : P.p);
            if (rc != BCL_OK) return rc;
            if (P.N != 1) return fail(BCL_ERR_SHAPE, "chain problems are single tokens (problem %d has N = %lld)", g, (long long)P.N);
            prm.prob[g] = bcl::ProblemDev{P.x, P.spec, P.bias, P.y, P.N, P.x_row_stride, P.y_row_stride, P.p, P.q, P.x_dtype, P.y_dtype,
                                          P.bias ? P.bias_dtype : BCL_F32, 0, {}, P.epilogue, 0, emit ? P.y_spectrum : nullptr};
            if (P.q > qmax) qmax = P.q;
            units += swiglu ? P.p / 2 : P.p;
            if ((long long)(swiglu ? 2 : 1) * P.q * P.p * (2 * sms_guess) > 0x7fffffffLL) return fail(BCL_ERR_SHAPE, "problem %d too large for the chain kernel's 32-bit work split", g);
        }
        if (units > max_units) max_units = units;
    }
    prm.step_first[nsteps] = count;
    prm.counter = reinterpret_cast<unsigned long long*>(barrier_counter);
    static long long* dbg = getenv("BCL_CHAIN_DBG_PTR") ? reinterpret_cast<long long*>(strtoull(getenv("BCL_CHAIN_DBG_PTR"), nullptr, 0)) : nullptr;   // tools/chain_trace.py only
    prm.dbg = dbg;
    if (!bcl::chain_supported(B, spec_dtype, qmax)) return fail(BCL_ERR_UNSUPPORTED, "x spectrum (q=%d, B=%d) leaves no room for the spectrum ring in shared memory", qmax, B);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
        return fail(BCL_ERR_CUDA, "cannot query the SM count");
    // one CTA per SM at most: the grid barrier needs every CTA resident (and two CTAs per SM measured slower for single steps too)
    const int cap = sms;
    const int grid = max_units < cap ? max_units : cap;
    cudaError_t e = bcl::launch_chain(prm, B, spec_dtype, qmax, grid, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "chain forward launch");
}

int bcl_push_columns(const void* src, void* const* dsts, int n_dst, int64_t N, int64_t row_stride, int64_t col_offset, int64_t width, int dtype, void* stream) {
    if (!src || !dsts || n_dst < 0 || n_dst > bcl::kMaxExtraDst) return fail(BCL_ERR_ARG, "need a source and 0..%d destinations", bcl::kMaxExtraDst);
    if (!valid_dt(dtype)) return fail(BCL_ERR_DTYPE, "unknown dtype enum");
    if (N < 0 || width < 0 || col_offset < 0 || row_stride < col_offset + width) return fail(BCL_ERR_SHAPE, "slice [%lld, +%lld) does not fit rows of %lld elements", (long long)col_offset, (long long)width, (long long)row_stride);
    const int es = bcl::dtype_size(dtype);
    if ((row_stride * es) % 16 || (col_offset * es) % 16 || (width * es) % 16 || !aligned(src, 16)) return fail(BCL_ERR_ALIGN, "rows, slice offset and width must be multiples of 16 bytes");
    if (N == 0 || width == 0 || n_dst == 0) return BCL_OK;
    void* d[bcl::kMaxExtraDst];
    for (int k = 0; k < n_dst; ++k) {
        if (!dsts[k] || !aligned(dsts[k], 16)) return fail(BCL_ERR_ALIGN, "destination %d is NULL or not 16-byte aligned", k);
        d[k] = reinterpret_cast<unsigned char*>(dsts[k]) + col_offset * es;
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = bcl::launch_push_columns(reinterpret_cast<const unsigned char*>(src) + col_offset * es, N, width * es, row_stride * es, d, n_dst, sms, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "push-columns launch");
}

int bcl_forward_host(const void* host_x, int x_dtype, int64_t N,
                     const void* spec, int spec_dtype, int layout, int p, int q, int B,
                     const void* bias, int bias_dtype, void* host_y, int y_dtype,
                     void* dev_x, void* dev_y, void* workspace, size_t workspace_bytes, void* stream) {
    if (!host_x || !host_y || !dev_x || !dev_y) return fail(BCL_ERR_ARG, "NULL host/device staging buffer");
    if (!valid_B(B) || p <= 0 || q <= 0 || N < 0 || !valid_dt(x_dtype) || !valid_dt(y_dtype)) return fail(BCL_ERR_ARG, "bad arguments");
    if (N == 0) return BCL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t xb = (size_t)N * q * B * bcl::dtype_size(x_dtype), yb = (size_t)N * p * B * bcl::dtype_size(y_dtype);
    cudaError_t e = cudaMemcpyAsync(dev_x, host_x, xb, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy of x");
    int rc = bcl_forward(dev_x, x_dtype, N, (int64_t)q * B, spec, spec_dtype, layout, p, q, B, bias, bias_dtype,
                         dev_y, y_dtype, (int64_t)p * B, workspace, workspace_bytes, stream);
    if (rc != BCL_OK) return rc;
    e = cudaMemcpyAsync(host_y, dev_y, yb, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "D2H copy of y");
    e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "stream synchronize");
}

int bcl_forward_grouped_host(const void* dev_plan, const void* host_plan, const void* host_x, void* dev_x, size_t x_bytes,
                             void* host_y, const void* dev_y, size_t y_bytes, void* stream) {
    if (!host_x || !dev_x || !host_y || !dev_y) return fail(BCL_ERR_ARG, "NULL host/device staging buffer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(dev_x, host_x, x_bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy of x");
    int rc = bcl_forward_grouped(dev_plan, host_plan, stream);
    if (rc != BCL_OK) return rc;
    e = cudaMemcpyAsync(host_y, dev_y, y_bytes, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "D2H copy of y");
    e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "stream synchronize");
}

int bcl_forward_grouped_host_pipelined(const void* dev_plan, const void* host_plan, const void* host_x, void* dev_x, size_t x_bytes,
                                       void* host_y, const void* dev_y, size_t y_bytes, void* stream_in, void* stream_compute, void* stream_out,
                                       void* ev_in_done, void* ev_compute_done, void* ev_out_done, int slot_used_before) {
    if (!host_x || !dev_x || !host_y || !dev_y) return fail(BCL_ERR_ARG, "NULL host/device staging buffer");
    if (!ev_in_done || !ev_compute_done || !ev_out_done) return fail(BCL_ERR_ARG, "NULL event");
    cudaStream_t si = (cudaStream_t)stream_in, sc = (cudaStream_t)stream_compute, so = (cudaStream_t)stream_out;
    cudaEvent_t ei = (cudaEvent_t)ev_in_done, ec = (cudaEvent_t)ev_compute_done, eo = (cudaEvent_t)ev_out_done;
    cudaError_t e;
#define BCL_TRY(call, what) if ((e = (call)) != cudaSuccess) return cuda_fail(e, what)
    if (slot_used_before) BCL_TRY(cudaStreamWaitEvent(si, ec, 0), "wait: previous kernel of this slot still reads dev_x");
    BCL_TRY(cudaMemcpyAsync(dev_x, host_x, x_bytes, cudaMemcpyHostToDevice, si), "H2D copy of x");
    BCL_TRY(cudaEventRecord(ei, si), "record H2D done");
    BCL_TRY(cudaStreamWaitEvent(sc, ei, 0), "wait: H2D");
    if (slot_used_before) BCL_TRY(cudaStreamWaitEvent(sc, eo, 0), "wait: previous D2H of this slot still reads dev_y");
    int rc = bcl_forward_grouped(dev_plan, host_plan, stream_compute);
    if (rc != BCL_OK) return rc;
    BCL_TRY(cudaEventRecord(ec, sc), "record kernel done");
    BCL_TRY(cudaStreamWaitEvent(so, ec, 0), "wait: kernel");
    BCL_TRY(cudaMemcpyAsync(host_y, dev_y, y_bytes, cudaMemcpyDeviceToHost, so), "D2H copy of y");
    BCL_TRY(cudaEventRecord(eo, so), "record D2H done");
#undef BCL_TRY
    return BCL_OK;
}

int bcl_forward_grouped_host_async(const void* dev_plan, const void* host_plan, const void* host_x, void* dev_x, size_t x_bytes,
                                   void* host_y, const void* dev_y, size_t y_bytes, void* stream) {
    if (!host_x || !dev_x || !host_y || !dev_y) return fail(BCL_ERR_ARG, "NULL host/device staging buffer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(dev_x, host_x, x_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D copy of x");
    int rc = bcl_forward_grouped(dev_plan, host_plan, stream);
    if (rc != BCL_OK) return rc;
    if ((e = cudaMemcpyAsync(host_y, dev_y, y_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail(e, "D2H copy of y");
    return BCL_OK;
}

int bcl_conj_transpose(const float* c, int p, int q, int B, float* ct, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d", p, q);
    if (!c || !ct || c == ct) return fail(BCL_ERR_ARG, "NULL or aliased c/ct");
    cudaError_t e = bcl::launch_conj_transpose(c, p, q, B, ct, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "conj-transpose launch");
}

size_t bcl_backward_workspace_bytes(int64_t N, int p, int q, int B) {
    if (!valid_B(B) || p <= 0 || q <= 0 || N <= 0) return 0;
    return bcl::backward_workspace_bytes(N, p, q, B);
}

int bcl_backward_weights(const void* x, int x_dtype, int64_t x_row_stride, const void* gy, int gy_dtype, int64_t gy_row_stride,
                         int64_t N, int p, int q, int B, float* dc, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0 || N < 0) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d N=%lld", p, q, (long long)N);
    if (!valid_dt(x_dtype) || !valid_dt(gy_dtype)) return fail(BCL_ERR_DTYPE, "unknown dtype enum");
    if (!dc) return fail(BCL_ERR_ARG, "NULL dc");
    if (N == 0) {
        if (!accumulate) { cudaError_t e = cudaMemsetAsync(dc, 0, (size_t)p * q * B * 4, (cudaStream_t)stream); if (e != cudaSuccess) return cuda_fail(e, "memset dc"); }
        return BCL_OK;
    }
    if (!x || !gy || !workspace) return fail(BCL_ERR_ARG, "NULL x/gy/workspace");
    if (x_row_stride < (int64_t)q * B || gy_row_stride < (int64_t)p * B) return fail(BCL_ERR_SHAPE, "row stride smaller than the feature dimension");
    if (!aligned(x, 16) || !aligned(gy, 16) || !aligned(dc, 16) || !aligned(workspace, 256) || (x_row_stride * bcl::dtype_size(x_dtype)) % 16 ||
        (gy_row_stride * bcl::dtype_size(gy_dtype)) % 16)
        return fail(BCL_ERR_ALIGN, "x, gy, dc (16 B), workspace (256 B) and row strides (16 B) must be aligned");
    if (workspace_bytes < bcl::backward_workspace_bytes(N, p, q, B)) return fail(BCL_ERR_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, bcl::backward_workspace_bytes(N, p, q, B));
    cudaError_t e = bcl::launch_backward_weights(x, x_dtype, x_row_stride, gy, gy_dtype, gy_row_stride, N, p, q, B, dc, accumulate, workspace, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "weight-gradient launch");
}

int bcl_project_from_dense(const void* W, int w_dtype, int64_t w_row_stride, int p, int q, int B, float* c, void* stream) {
    if (!valid_B(B)) return fail(BCL_ERR_BLOCK, "block size %d not supported (64, 128, 256)", B);
    if (p <= 0 || q <= 0) return fail(BCL_ERR_SHAPE, "bad shape p=%d q=%d", p, q);
    if (!W || !c) return fail(BCL_ERR_ARG, "NULL W/c");
    if (!valid_dt(w_dtype)) return fail(BCL_ERR_DTYPE, "unknown dtype enum");
    if (w_row_stride < (int64_t)q * B) return fail(BCL_ERR_SHAPE, "row stride smaller than in_features");
    cudaError_t e = bcl::launch_project_from_dense(W, w_dtype, w_row_stride, p, q, B, c, (cudaStream_t)stream);
    return e == cudaSuccess ? BCL_OK : cuda_fail(e, "projection launch");
}

}  // extern "C"
