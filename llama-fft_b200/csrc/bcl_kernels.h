// bcl_kernels.h -- internal interface between the C-ABI layer (bcl_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace bcl {

// Device-side view of one BC-linear problem (mirrors bcl_problem, plus nothing the host must keep alive).
struct ProblemDev {
    const void* x;
    const void* spec;
    const void* bias;
    void* y;
    int64_t N;
    int64_t x_stride;   // elements
    int64_t y_stride;   // elements
    int32_t p, q;
    int32_t x_dtype, y_dtype, bias_dtype;
    int32_t n_extra;        // additional destinations of y (same shape / stride / dtype), e.g. peer GPUs' buffers
    void* y_extra[7];
    int32_t epilogue;       // chain kernel only: BCL_EPILOGUE_* bits (SWIGLU: rows [0,p/2) gate, [p/2,p) up; EMIT_SPECTRUM: also write y_spectrum)
    int32_t pad_;
    void* y_spectrum;       // chain kernel only: packed fp32 spectrum of the rounded output, (out/B, B)
};
constexpr int kMaxExtraDst = 7;

// Parameters of the chain kernel (bcl_stream_chain.cu): up to kChainMaxSteps dependent steps over kChainMaxProblems problems, all N == 1.
constexpr int kChainMaxProblems = 8;
constexpr int kChainMaxSteps = 8;
constexpr unsigned long long kChainQuantum = 1ull << 20;     // every launch advances the barrier counter by exactly this much
struct ChainParams {
    int32_t nsteps, nproblems;
    int32_t step_first[kChainMaxSteps + 1];
    ProblemDev prob[kChainMaxProblems];
    unsigned long long* counter;      // device, 8 bytes, zeroed once by the caller; never reset by the library (nsteps > 1 only)
    long long* dbg;                   // profiling builds / tools only: clock64 stamps of CTA 0 (NULL otherwise)
};

// Header of an encoded group plan (host and device copies are byte-identical).
struct GroupPlanHeader {
    uint32_t magic;       // 'BCLG'
    int32_t count;        // problems
    int32_t num_ctas;
    int32_t num_segs;
    int32_t B, spec_dtype, nt, qmax;
    int32_t variant;      // 0 = plain CTA-cooperative kernel, 1 = warp-specialised TMA pipeline (one CTA per SM), 2 = its row-per-warp form (bf16 spectra),
                          // 3 = tensor-core MAC over BCL_LAYOUT_DECODE_TC spectra (bcl_stream_tc.cu)
    int32_t reserved;     // plan flags.  bit 0 (BCL_GROUP_INDEPENDENT=1 at encode time, a measurement switch): the launch neither reads what the
                          // previous launch of the stream wrote nor overwrites what it still reads, so it skips griddepcontrol.wait
    uint64_t off_problems, off_cta_ptr, off_segs, total_bytes;
    int32_t num_dyn;      // dynamic tail (variant 1): items segs[num_segs .. num_segs + num_dyn) that the CTAs draw from a ticket counter
    int32_t dyn_rows;     // block-rows per item
    uint64_t off_ctr;     // launch scratch in the DEVICE copy of the plan (ints): ticket counter, CTAs done (both zero between launches), launch
                          // sequence number, pad, dyn_problems ready marks, then (constant) the dyn_problems problem indices
    int32_t dyn_problems; // problems that own items of the tail
    int32_t dyn_xstride;  // floats between their cached input spectra
    uint64_t off_xcache;  // [dyn_problems][dyn_xstride] floats in the device copy: rfft(x) of the tail's problems, computed once per launch
};
constexpr uint32_t kPlanMagic = 0x474c4342u;

cudaError_t launch_spectrum_stream(const float* c, void* spec, int p, int q, int B, int spec_dtype, cudaStream_t st);
cudaError_t launch_stream_single(const ProblemDev& P, int B, int spec_dtype, cudaStream_t st);
cudaError_t launch_stream_grouped(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st);
bool stream_ws_supported(int B, int spec_dtype, int nt, int qmax);
cudaError_t launch_stream_ws(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st);
bool stream_rw_supported(int B, int spec_dtype, int nt, int qmax);
cudaError_t launch_stream_rw(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st);
size_t stream_smem_bytes(int B, int nt, int q, int warps);
int stream_pick_nt(int64_t N, int q, int B);

// grouped decode with the MAC on the tensor cores (bcl_stream_tc.cu): BCL_LAYOUT_DECODE_TC spectra (bf16), B <= 128, one token per problem
size_t stream_tc_spectrum_bytes(int p, int q, int B);
cudaError_t launch_spectrum_tc(const float* c, void* spec, int p, int q, int B, cudaStream_t st);
bool stream_tc_supported(int B, int spec_dtype, int nt, int qmax);
cudaError_t launch_stream_tc(const GroupPlanHeader& H, const void* dev_plan, cudaStream_t st);

bool chain_supported(int B, int spec_dtype, int qmax);
cudaError_t launch_chain(const ChainParams& prm, int B, int spec_dtype, int qmax, int grid, cudaStream_t st);

// prefill path (bcl_gemm.cu): GEMM-layout spectrum, the three stages, and the chunked forward that chains them
// split = 1: the fp32-accurate variant (fp32 spectra): (hi, lo) bf16 pairs for Cmat / Xf / Yf, three MMAs per k-step
size_t gemm_spectrum_bytes(int p, int q, int B, int split);
size_t gemm_workspace_bytes(int64_t N, int p, int q, int B, int split);
int64_t gemm_chunk_tokens(int64_t N, int p, int q, int B, size_t workspace_bytes, int split);
cudaError_t launch_spectrum_gemm(const float* c, void* spec, int p, int q, int B, int split, cudaStream_t st);
cudaError_t launch_gemm_rfft(const ProblemDev& P, int B, int64_t n_begin, int nc, void* Xf, void* Xf_lo, cudaStream_t st);
cudaError_t launch_gemm_mma(const void* Xf, const void* Xf_lo, const void* Cmat, void* Yf, void* Yf_lo, int* status, int p, int q, int B, int nc, int max_ctas,
                            cudaStream_t st);
cudaError_t launch_gemm_irfft(const ProblemDev& P, int B, int64_t n_begin, int nc, const void* Yf, const void* Yf_lo, cudaStream_t st);
cudaError_t launch_prefill_stage1(const ProblemDev& P, int B, const void* spec_gemm, int64_t n_begin, int nc, void* Xf, int* status, int impl, cudaStream_t st);
cudaError_t launch_prefill_stage3(const ProblemDev& P, int B, const void* spec_gemm, int64_t n_begin, int nc, const void* Yf, int* status, int impl, cudaStream_t st);
cudaError_t launch_gemm_forward(const ProblemDev& P, int B, const void* Cmat, void* workspace, size_t workspace_bytes, int split, cudaStream_t st);
cudaError_t launch_gemm_forward_swiglu(const ProblemDev& P, int B, const void* spec_gemm, void* workspace, size_t workspace_bytes, int split, cudaStream_t st);

// calibration gradients (bcl_backward.cu)
size_t backward_workspace_bytes(int64_t N, int p, int q, int B);
cudaError_t launch_conj_transpose(const float* c, int p, int q, int B, float* ct, cudaStream_t st);
cudaError_t launch_backward_weights(const void* x, int x_dtype, int64_t x_stride, const void* gy, int gy_dtype, int64_t gy_stride, int64_t N, int p, int q,
                                    int B, float* dc, int accumulate, void* workspace, cudaStream_t st);

// exchange step of the sharded layer (bcl_exchange.cu)
cudaError_t launch_push_columns(const void* src, int64_t rows, int64_t row_bytes, int64_t pitch_bytes, void* const* dsts, int n_dst, int sms, cudaStream_t st);

cudaError_t launch_project_from_dense(const void* W, int w_dtype, int64_t w_stride, int p, int q, int B, float* c, cudaStream_t st);

}  // namespace bcl
