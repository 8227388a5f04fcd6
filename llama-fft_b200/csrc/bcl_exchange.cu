// bcl_exchange.cu -- the exchange step of output-block-row sharding (llama-fft_b200/sharded.py): a rank's column slice of
// y goes STRAIGHT into the final (N, out) layout of every peer's buffer.  The reference has no distributed code; the contract is
// BASELINE.json configs[4] ("output-block-row sharded ... with NCCL all-gather").  An all-gather collective produces (G, N, out/G);
// the row-major (N, out) the next layer reads needs a second, full-size interleave pass.  This kernel is the gather without that pass:
// every rank reads its slice once and writes it, 16 bytes per thread and coalesced along rows, into the same columns of each
// peer-mapped destination (NVLink peer stores; symmetric memory), so the bytes cross the link exactly once.
#include "bcl_common.cuh"
#include "bcl_kernels.h"

namespace bcl {

struct PushArgs {
    const unsigned char* src;       // first byte of the slice in the local buffer (row 0)
    unsigned char* dst[kMaxExtraDst];
    int n_dst;
    int64_t rows, row_bytes;        // slice width in bytes (multiple of 16)
    int64_t pitch;                  // row pitch of all buffers in bytes (multiple of 16)
};

__global__ void __launch_bounds__(256) push_columns_kernel(const __grid_constant__ PushArgs a) {
    const int64_t chunks_per_row = a.row_bytes / 16;
    const int64_t total = a.rows * chunks_per_row;
    constexpr int U = 4;                                             // 16-byte loads in flight per thread
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += U * stride) {
        uint4 v[U];
        int64_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * stride;
            if (i < total) {
                const int64_t r = i / chunks_per_row, c = i - r * chunks_per_row;
                off[u] = r * a.pitch + c * 16;
                v[u] = __ldcs(reinterpret_cast<const uint4*>(a.src + off[u]));      // read once: streaming
            } else {
                off[u] = -1;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (off[u] >= 0) {
                for (int d = 0; d < a.n_dst; ++d) *reinterpret_cast<uint4*>(a.dst[d] + off[u]) = v[u];
            }
        }
    }
}

cudaError_t launch_push_columns(const void* src, int64_t rows, int64_t row_bytes, int64_t pitch_bytes, void* const* dsts, int n_dst, int sms, cudaStream_t st) {
    if (rows <= 0 || row_bytes <= 0 || n_dst <= 0) return cudaSuccess;
    PushArgs a{};
    a.src = reinterpret_cast<const unsigned char*>(src);
    for (int d = 0; d < n_dst; ++d) a.dst[d] = reinterpret_cast<unsigned char*>(dsts[d]);
    a.n_dst = n_dst; a.rows = rows; a.row_bytes = row_bytes; a.pitch = pitch_bytes;
    const int64_t total = rows * (row_bytes / 16);
    int64_t grid = (total + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)sms * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    push_columns_kernel<<<(unsigned)grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bcl
