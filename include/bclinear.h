/* bclinear.h -- C ABI of libbclinear.so: the block-circulant (BC) linear forward for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of Pinguhat/llama-fft: BlockCirculantLinear.forward
 * (reference src/patch_llama_fft.py:226-275; cached-spectrum variant src/bench_all_in_one.py:179-231).
 * The reference has no FFI of its own -- its seam is the Python module object -- so the entry points below
 * are what a binding for that path needs: each one cites the reference lines it replaces.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   B  block size, 64 | 128 | 256          q = in_features / B        p = out_features / B
 *   c[j,i,:] is the first column of the circulant block mapping input block i to output block j,
 *   C_{j,i}[r,s] = c[j,i,(r-s) mod B]                         (src/patch_llama_fft.py:152-154, :86-88)
 *   y[n, jB:(j+1)B] = irfft( sum_i rfft(c[j,i]) * rfft(x[n, iB:(i+1)B]) )   (src/patch_llama_fft.py:252-263)
 *
 * Ownership and threading
 *   Every pointer is caller-owned.  Device pointers unless the name says host.  The library never allocates,
 *   frees or retains caller memory; its only mutable global state is a thread-local error string plus three mutex-guarded /
 *   write-once caches of facts about the device (dynamic-shared-memory attributes already granted per (kernel, device), the SM
 *   count per device, profiling / A-B environment switches, most read once) -- nothing that depends on the data of a call.  It never syncs
 *   the device (except the *_host convenience calls, which say so), and every launch goes to the stream that
 *   is passed in (a cudaStream_t cast to void*; NULL = the legacy default stream).  All calls are CUDA-graph
 *   capturable except the *_host ones.  Errors never abort or print: a negative bcl_status is returned and
 *   bcl_last_error() describes it.  There is no CPU fallback and no other backend.
 */
#ifndef BCLINEAR_H_
#define BCLINEAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCL_ABI_VERSION 14

typedef enum bcl_dtype { BCL_F32 = 0, BCL_F16 = 1, BCL_BF16 = 2 } bcl_dtype;

/* Spectrum layouts in HBM (DESIGN.md "Data layout").
 *   STREAM  [p][q][B] packed half-spectrum, frequency-contiguous: {Re F0, Re F_{B/2}, Re F1, Im F1, ...}.
 *           Exactly the size of c.  Used by the bandwidth-bound small-N (decode) kernel.
 *   GEMM    the tcgen05 prefill path's buffer: per-frequency real-embedded complex matrices
 *           Cmat[f][n][k], f < B/2 packed frequency slots, n = 2j+{re,im} < Np, k = 2i+{re,im} < Kp (K-major rows,
 *           Np = 2p rounded up to 64, Kp = 2q rounded up to 16, zero padded); [[Cr,-Ci],[Ci,Cr]] per (j,i), slot 0 =
 *           diag(C_dc, C_ny); followed by the forward (fp16, bf16) and inverse (bf16) B x B DFT matrices.
 *           spec_dtype BCL_BF16: one bf16 Cmat (parity class 1e-2).  spec_dtype BCL_F32: the SPLIT layout -- Cmat_hi then
 *           Cmat_lo, bf16 each, value = hi + lo to ~2^-17; the frequency GEMM evaluates Ahi*Bhi + Ahi*Blo + Alo*Bhi with fp32
 *           accumulation and the rfft / irfft stages run in fp32 (parity class 1e-4, the reference's complex64 contraction). */
typedef enum bcl_layout {
    BCL_LAYOUT_STREAM = 0,      /* [p][q][B] packed half-spectra (f32 / bf16): the CUDA-core decode kernels, bcl_forward at small N */
    BCL_LAYOUT_GEMM = 1,        /* per-frequency K-major matrices + DFT matrices: the tcgen05 prefill path (bcl_forward at large N) */
    BCL_LAYOUT_DECODE_TC = 2    /* bf16, [32-row quad][f][8-row group][k/8][8][8]: the UMMA operand image streamed by the tensor-core
                                   decode kernel (grouped launches with one token per problem, B <= 128; bcl_group_plan_encode_layout) */
} bcl_layout;

typedef enum bcl_status {
    BCL_OK = 0,
    BCL_ERR_ARG = -1,        /* NULL pointer, negative size, unknown enum */
    BCL_ERR_SHAPE = -2,      /* p, q, N out of range / in_features mismatch */
    BCL_ERR_BLOCK = -3,      /* block size not in {64,128,256} */
    BCL_ERR_ALIGN = -4,      /* pointer or row stride not 16-byte aligned */
    BCL_ERR_DTYPE = -5,      /* unsupported dtype combination for this layout */
    BCL_ERR_WORKSPACE = -6,  /* workspace too small */
    BCL_ERR_CUDA = -7,       /* a CUDA runtime call failed; see bcl_last_error() */
    BCL_ERR_UNSUPPORTED = -8 /* valid request this build has no kernel for (never a silent fallback) */
} bcl_status;

int bcl_abi_version(void);

/* Thread-local, NUL-terminated description of the last failure on this thread ("" if none). */
const char* bcl_last_error(void);

/* ---- spectrum precompute: replaces Cf = rfft(c.to(float32)) --------------------------------------------
 * reference: src/patch_llama_fft.py:256-257 (recomputed every forward) and the memoised
 * _Cf_cache of src/bench_all_in_one.py:204-219.  c is the (p,q,B) float32 parameter, row-major. */
size_t bcl_spectrum_bytes(int p, int q, int B, int spec_dtype, int layout);
int bcl_precompute_spectrum(const float* c, int p, int q, int B,
                            void* spec, int spec_dtype, int layout, void* stream);

/* ---- forward: replaces steps 2-9 of BlockCirculantLinear.forward -----------------------------------------
 * reference: src/patch_llama_fft.py:246-271 (reshape, fp32 up-cast, rfft, einsum, irfft, down-cast, bias).
 *   x        (N, in) row-major, inner dimension contiguous, row stride in ELEMENTS
 *   y        (N, out), dtype y_dtype.  Epilogue order is the reference's: y32 is rounded to x_dtype first
 *            (:266), then the bias is added in the promoted dtype y_dtype (:270-271).  Without a bias
 *            y_dtype must equal x_dtype.
 *   bias     (out) or NULL
 *   workspace  at least bcl_workspace_bytes(...) bytes, 256-byte aligned (may be NULL when that is 0).  Layout GEMM
 *            stages Xf/Yf of one token chunk there (a larger workspace means larger chunks); its first int is a
 *            status word the call zeroes and the tcgen05 kernels set non-zero if their pipeline ever stalls; its second
 *            int is STICKY: the code of the first stall since the caller last zeroed it (zero it once after allocating).
 * N == 0 is a no-op.  */
size_t bcl_workspace_bytes(int64_t N, int p, int q, int B, int spec_dtype, int layout);
int bcl_forward(const void* x, int x_dtype, int64_t N, int64_t x_row_stride,
                const void* spec, int spec_dtype, int layout, int p, int q, int B,
                const void* bias, int bias_dtype,
                void* y, int y_dtype, int64_t y_row_stride,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- gate_proj | up_proj + SWIGLU for many tokens (whole-MLP fusion at prefill) ------------------------------
 * Replaces the first half of what HF's LlamaMLP.forward chains over the three modules the reference swaps in
 * (src/patch_llama_fft.py:379):  act_fn(gate_proj(x)) * up_proj(x).  `spec_gemm` is the GEMM-layout spectrum (spec_dtype BCL_BF16, or
 * BCL_F32 = the (hi, lo) bf16 pairs of the fp32-accurate path: three MMAs per k-step in every stage) of gate_proj and up_proj STACKED along the output-block axis (p = 2 * intermediate/B block-rows: [0, p/2) gate, [p/2, p) up; bcl_precompute_spectrum
 * of the concatenated c).  One forward-DFT pass and one frequency-GEMM pass serve both members; the inverse stage transforms the two
 * members of a block pair into a pair of TMEM accumulators and writes
 *     h = round(silu(round(gate))) * round(up)        (every round() to x_dtype, silu in fp32: the arithmetic of the unfused modules)
 * as (N, (p/2)*B): the two projection outputs never reach HBM and the act_fn / multiply kernels disappear.  B <= 128, 16-bit
 * activations, no bias; BCL_ERR_UNSUPPORTED otherwise (callers then run the modules one by one).  Workspace as for bcl_forward. */
int bcl_forward_swiglu(const void* x, int x_dtype, int64_t N, int64_t x_row_stride,
                       const void* spec_gemm, int spec_dtype, int p, int q, int B,
                       void* h, int h_dtype, int64_t h_row_stride,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- forward with several destinations: the fused gather of output-block-row sharding ---------------------------
 * Same as bcl_forward, but the epilogue stores y to every one of y_dsts[0 .. n_dst) (1 <= n_dst <= 8; same shape, row
 * stride and dtype).  With peer-mapped pointers (symmetric memory over NVLink) a rank that owns block-rows [j0, j1) of a
 * sharded layer writes its slice y[:, j0*B : j1*B] straight into every peer's full y: no collective afterwards.
 * The reference has no distributed path; see llama-fft_b200/sharded.py. */
int bcl_forward_bcast(const void* x, int x_dtype, int64_t N, int64_t x_row_stride,
                      const void* spec, int spec_dtype, int layout, int p, int q, int B,
                      const void* bias, int bias_dtype,
                      void* const* y_dsts, int n_dst, int y_dtype, int64_t y_row_stride,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- prefill stages (layout GEMM): the three steps bcl_forward chains per token chunk, exported so that each can be
 * checked on its own against the reference step it replaces.  Intermediates are bf16 in 16-byte chunks:
 *   Xf [B/2][Kp/8][Nt][8]  Xf[f][kg][n][k0], k = 8 kg + k0 = 2i+{0,1} = {Re,Im} rfft(x[n, iB:(i+1)B])[f]  (f = 0: {DC, Nyquist})   src/patch_llama_fft.py:252
 *   Yf [B/8][p][Nt][8]     Yf[fg][j][n][2 fl+{0,1}] = {Re,Im} sum_i Cf[j,i,f] * Xf[n,i,f],  f = 4 fg + fl                         src/patch_llama_fft.py:260
 * with the token pitch Nt = N rounded up to 32 (rows N..Nt-1 are padding the kernels may write).
 * spec_gemm is the GEMM-layout spectrum buffer of the layer (it also carries the DFT matrices of the tensor-core stages).
 * impl: 0 = pick, 1 = CUDA-core register FFT, 2 = DFT as a tcgen05 GEMM (B <= 128, 16-bit activations; else BCL_ERR_UNSUPPORTED).
 * status: device int zeroed by the caller; a tcgen05 kernel sets it non-zero if its pipeline ever stalls. */
void bcl_prefill_geometry(int p, int q, int B, int* F, int* Kp, int* Np, size_t* wf16_off, size_t* wbf16_off, size_t* winv_off);
int bcl_prefill_rfft(const void* x, int x_dtype, int64_t N, int64_t x_row_stride, int p, int q, int B, const void* spec_gemm, int impl,
                     void* Xf, int* status, void* stream);
int bcl_prefill_gemm(const void* Xf, const void* spec_gemm, void* Yf, int64_t N, int p, int q, int B, int* status, void* stream);
int bcl_prefill_irfft(const void* Yf, int64_t N, int p, int q, int B, const void* spec_gemm, int impl, int x_dtype,
                      const void* bias, int bias_dtype, void* y, int y_dtype, int64_t y_row_stride, int* status, void* stream);

/* ---- grouped forward: many independent BC linears in ONE launch -------------------------------------------
 * The reference calls one module at a time (HF LlamaMLP/LlamaAttention -> forward); at batch-1
 * decode a single projection is 0.5-1.4 MB of spectrum, far below launch latency, so callers that hold
 * several independent projections (q/k/v of one layer, gate/up, or a synthetic sweep over layers) submit
 * them together.  All problems of a group share B, spec_dtype (STREAM layout) and have N <= 8.
 * The plan (problem table + a byte-balanced partition of output block-rows over CTAs) is encoded on the
 * host into caller memory; the caller copies it to the device once and replays it. */
typedef struct bcl_problem {
    const void* x;          /* (N, q*B)  */
    const void* spec;       /* STREAM layout, p*q*B elements of spec_dtype */
    const void* bias;       /* (p*B) or NULL */
    void* y;                /* (N, p*B)  */
    int64_t N;
    int64_t x_row_stride;   /* elements */
    int64_t y_row_stride;   /* elements */
    int32_t p, q;
    int32_t x_dtype, y_dtype, bias_dtype;
    int32_t epilogue;       /* BCL_EPILOGUE_LINEAR (0); bcl_forward_chain also takes BCL_EPILOGUE_SWIGLU / BCL_EPILOGUE_EMIT_SPECTRUM bits */
    void* y_spectrum;       /* bcl_forward_chain with BCL_EPILOGUE_EMIT_SPECTRUM: (output blocks, B) fp32, else ignored (NULL) */
} bcl_problem;

/* Epilogues.  LINEAR is the reference's (src/patch_llama_fft.py:263-271: 1/B, round to x_dtype, + bias in the promoted dtype).
 * SWIGLU (bcl_forward_chain only) treats the problem as gate_proj and up_proj stacked along the output-block axis -- block-rows
 * [0, p/2) gate, [p/2, p) up, as shared_input.py stacks them -- and writes what HF LlamaMLP.forward feeds down_proj
 * (transformers LlamaMLP.forward; the reference swaps exactly those three modules, src/patch_llama_fft.py:379):
 *     h = round(silu(round(gate)) ) * round(up),  every round() to x_dtype, y is (1, (p/2)*B), bias must be NULL. */
#define BCL_EPILOGUE_LINEAR 0
#define BCL_EPILOGUE_SWIGLU 1
/* EMIT_SPECTRUM (bit, bcl_forward_chain only, no bias): besides y, the finisher writes y_spectrum = the packed fp32 half-spectrum
 * (STREAM packing) of every ROUNDED output block -- bit for bit what the next step's rfft(x) (src/patch_llama_fft.py:252) would
 * compute from y.  A later step of the same chain may then take x = that buffer with x_dtype = BCL_X_SPECTRUM and skips its own
 * forward transform, which every CTA would otherwise repeat for the whole input (at one token the transform of a 11008-wide
 * input costs more than streaming the weights). */
#define BCL_EPILOGUE_EMIT_SPECTRUM 2
#define BCL_X_SPECTRUM 3    /* x_dtype of a chain problem whose x is a (q, B) fp32 packed spectrum; y_dtype then names the activation dtype */

size_t bcl_group_plan_bytes(int count, int num_ctas);
/* Bytes of a plan with room for the DYNAMIC TAIL of the fp32 decode kernel (>= bcl_group_plan_bytes; size host_plan and dev_plan
 * with it).  An opt-in experiment (BCL_GROUP_DYNAMIC="percent[,rows]" at encode time; default off = every row in a static share):
 * equal-byte shares of 148 CTAs finish up to ~6 us apart on B200 and a dependent next launch waits for the slowest, so the encoder
 * can leave the last `percent` of the spectrum bytes unassigned: the CTAs draw them in `rows`-row items from a ticket counter in the
 * DEVICE copy of the plan, and rfft(x) of the problems those items belong to (src/patch_llama_fft.py:252) is computed once per launch
 * into a cache behind the plan.  Which CTA computes a row does not change a bit of it (tests), but the tail measured SLOWER than the
 * static plan, hence off.  With it the device copy holds a few words the kernel writes (left zero / monotonic after
 * every launch): launches of ONE plan must then be stream-ordered, not concurrent -- they write the same y anyway.  A buffer sized
 * by bcl_group_plan_bytes alone gets a plan without the tail whatever the environment says. */
size_t bcl_group_plan_bytes_for(const bcl_problem* problems, int count, int B, int spec_dtype, int num_ctas);
/* num_ctas: how many CTAs to partition over (e.g. 2 * SM count); returns BCL_OK and fills host_plan. */
int bcl_group_plan_encode(const bcl_problem* problems, int count, int B, int spec_dtype, int num_ctas,
                          void* host_plan, size_t host_plan_bytes);
/* Same, for problems whose `spec` pointers are in `layout`.  BCL_LAYOUT_DECODE_TC (bf16, one token per problem, B <= 128): the plan
 * selects the tensor-core decode kernel -- the per-frequency contraction (einsum, src/patch_llama_fft.py:260) as tcgen05 matrix-vector
 * products, the spectrum streamed by bulk copies as the UMMA operand image -- and cuts the work at multiples of 8 block-rows.
 * BCL_ERR_UNSUPPORTED when the configuration is not covered (callers then plan with the STREAM layout). */
int bcl_group_plan_encode_layout(const bcl_problem* problems, int count, int B, int spec_dtype, int layout, int num_ctas,
                                 void* host_plan, size_t host_plan_bytes);
/* dev_plan = device copy of host_plan (same bytes).  host_plan is read for the launch geometry only. */
int bcl_forward_grouped(const void* dev_plan, const void* host_plan, void* stream);

/* ---- chain forward: DEPENDENT BC linears of one token in ONE launch (decode) -------------------------------------------------
 * A decoder layer's projections depend on each other (down_proj(act_fn(gate_proj(x)) * up_proj(x)), HF LlamaMLP.forward over the
 * modules the reference swaps, src/patch_llama_fft.py:379); launched one module at a time a token costs ~5 us of latency per module
 * for ~0.2 us of HBM time.  bcl_forward_chain runs nsteps <= BCL_CHAIN_MAX_STEPS steps in one persistent kernel with a grid-wide
 * barrier between steps; step s consists of problems [step_first[s], step_first[s+1]) (count = step_first[nsteps] <=
 * BCL_CHAIN_MAX_PROBLEMS), which may read what earlier steps wrote.  The spectra of later steps are prefetched into shared
 * memory while earlier steps compute.  All problems: N == 1, STREAM layout, the same B and spec_dtype; the descriptors travel as
 * kernel parameters (no plan buffer).  With nsteps == 1 it is simply the low-latency small launch for a few independent problems.
 *   barrier_counter  device memory, 8 bytes, 8-byte aligned, zeroed ONCE by the caller and then owned by the library (never reset:
 *                    every launch advances it by 2^20); one counter may serve every chain launch of a stream, but not launches
 *                    that can run concurrently.  May be NULL when nsteps == 1.
 * Graph-capturable.  BCL_ERR_UNSUPPORTED if the device cannot hold the whole grid at once (the barrier needs co-resident CTAs). */
#define BCL_CHAIN_MAX_PROBLEMS 8
#define BCL_CHAIN_MAX_STEPS 8
int bcl_forward_chain(const bcl_problem* problems, const int* step_first, int nsteps, int B, int spec_dtype,
                      void* barrier_counter, void* stream);

/* ---- end-to-end convenience with HOST buffers (what bench.py's e2e figure times) --------------------------
 * Copies x from (pinned) host memory to the caller's device staging buffers, runs bcl_forward, copies y
 * back and synchronises the stream.  dev_x/dev_y are caller-owned device staging areas of N*in / N*out
 * elements.  Not graph-capturable (it synchronises). */
int bcl_forward_host(const void* host_x, int x_dtype, int64_t N,
                     const void* spec, int spec_dtype, int layout, int p, int q, int B,
                     const void* bias, int bias_dtype,
                     void* host_y, int y_dtype,
                     void* dev_x, void* dev_y, void* workspace, size_t workspace_bytes, void* stream);

/* Grouped variant: all problems' x live back to back in one staging area (and all y in another), so a step is one
 * H2D copy, one launch, one D2H copy, one synchronise. */
int bcl_forward_grouped_host(const void* dev_plan, const void* host_plan,
                             const void* host_x, void* dev_x, size_t x_bytes,
                             void* host_y, const void* dev_y, size_t y_bytes, void* stream);

/* Pipelined variant for a stream of steps: step k's H2D copy, kernel and D2H copy run on three streams, so the copies
 * of neighbouring steps overlap the kernel (PCIe is full duplex) and a step costs max(H2D, kernel, D2H) instead of their sum.
 * A "slot" is one set of staging buffers (host_x, dev_x, dev_y, host_y) plus three events; use >= 2 slots round-robin.
 * The call only enqueues (no synchronisation): it records ev_in_done / ev_compute_done / ev_out_done for this slot and, when
 * slot_used_before != 0, first waits on the slot's previous ev_compute_done (dev_x free) and ev_out_done (dev_y free).
 * host_y of the slot is valid once ev_out_done has completed (cudaEventSynchronize); the caller must not touch host_x /
 * host_y of a slot that is still in flight.  Streams and events are caller-owned (cudaStream_t / cudaEvent_t as void*). */
int bcl_forward_grouped_host_pipelined(const void* dev_plan, const void* host_plan,
                                       const void* host_x, void* dev_x, size_t x_bytes,
                                       void* host_y, const void* dev_y, size_t y_bytes,
                                       void* stream_in, void* stream_compute, void* stream_out,
                                       void* ev_in_done, void* ev_compute_done, void* ev_out_done, int slot_used_before);
/* The same step on ONE stream, asynchronously: H2D(x) -> grouped launch -> D2H(y) are enqueued in order and the call returns at once
 * (three API calls, no events).  With one stream per staging slot the steps of neighbouring slots overlap by themselves -- copy
 * engines and SMs serve different streams concurrently -- which costs the host a third of the calls of the three-stream form.
 * The caller synchronises the stream (or an event recorded on it) before it reads host_y or resubmits the slot's buffers. */
int bcl_forward_grouped_host_async(const void* dev_plan, const void* host_plan, const void* host_x, void* dev_x, size_t x_bytes,
                                   void* host_y, const void* dev_y, size_t y_bytes, void* stream);


/* ---- gradients for calibration -----------------------------------------------------------------------------------
 * reference: torch autograd through BlockCirculantLinear.forward (src/patch_llama_fft.py:226-275), used by the KL
 * calibration loop (src/calibrate_bc.py:281-294).  All arithmetic fp32.
 *   dx  = bcl_forward(gy, spectrum of ct) with ct = bcl_conj_transpose(c):  ct[i,j,t] = c[j,i,(-t) mod B]   (q x p blocks)
 *   dc[j,i,t] = sum_n sum_r gy[n,j,r] * x[n,i,(r-t) mod B]                   bcl_backward_weights (dc (p,q,B) fp32;
 *               accumulate != 0 adds to dc, e.g. over token chunks); workspace >= bcl_backward_workspace_bytes(N,p,q,B)
 *   dbias = sum_n gy[n]  is a plain reduction left to the caller. */
int bcl_conj_transpose(const float* c, int p, int q, int B, float* ct, void* stream);
size_t bcl_backward_workspace_bytes(int64_t N, int p, int q, int B);
int bcl_backward_weights(const void* x, int x_dtype, int64_t x_row_stride, const void* gy, int gy_dtype, int64_t gy_row_stride,
                         int64_t N, int p, int q, int B, float* dc, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- construction: replaces the Python p*q*B loops of from_linear ------------------------------------------
 * reference: src/patch_llama_fft.py:201-219 + :110-128.  W is (p*B, q*B) row-major of dtype w_dtype;
 * c[j,i,t] = alpha * mean_r W[jB+r, iB+(r-t) mod B], alpha as :124-126, arithmetic rounded to w_dtype the way
 * the reference's per-block torch ops round, result stored as float32 (p,q,B). */
int bcl_project_from_dense(const void* W, int w_dtype, int64_t w_row_stride, int p, int q, int B,
                           float* c, void* stream);

/* ---- exchange step of output-block-row sharding (BASELINE.json configs[4]; no reference counterpart: SURVEY 2a / 8e) -----------
 * Rank g of G holds block-rows [j0, j1) of c and computes y[:, j0*B : j1*B] -- with bcl_forward and y_row_stride = out it lands in
 * place in the rank's own (N, out) buffer.  bcl_push_columns then copies that column slice into the SAME columns of n_dst peer
 * buffers (peer-mapped device pointers, e.g. symmetric memory; same shape and row stride): a gather straight into the final layout,
 * one read of the slice, one NVLink store per peer, no collective and no interleave pass afterwards.  src / dsts point at ROW 0,
 * COLUMN 0 of the (N, out) buffers; col_offset / width in elements, 16-byte aligned in bytes.  The caller orders the exchange
 * against the peers' reads (a barrier before and after; sharded.py uses the symmetric-memory handle's). */
int bcl_push_columns(const void* src, void* const* dsts, int n_dst, int64_t N, int64_t row_stride, int64_t col_offset, int64_t width,
                     int dtype, void* stream);

/* ---- measurement aid ----------------------------------------------------------------------------------------------
 * Launches an EMPTY kernel (1 CTA, 32 threads) on `stream`; pdl != 0 launches it with the programmatic-stream-serialization
 * attribute the forward kernels use.  bench.py replays graphs of these to print the launch floor next to every
 * launch-latency-bound figure ("report per-call latency next to the empty-kernel launch floor").
 * No reference counterpart. */
int bcl_launch_noop(int pdl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BCLINEAR_H_ */
