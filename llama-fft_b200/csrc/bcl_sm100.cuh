// bcl_sm100.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) machinery the prefill path uses:
// mbarrier, TMA tensor loads (cp.async.bulk.tensor), TMEM allocation, tcgen05.mma / commit / ld, and the shared-memory
// + instruction descriptors of the 5th-generation tensor cores.  Nothing here is a library call: every wrapper is one
// PTX instruction (SASS: UTCHMMA / UTMALDG / LDTM / SYNCS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bcl {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
// Invalidate an mbarrier so that its shared-memory location may be initialised again (a CTA that runs several stages in turn).
__device__ __forceinline__ void mbar_inval(uint64_t* b) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a system-dependent time; test_wait returns at once).
__device__ __forceinline__ bool mbar_test(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error flag, never as a hung GPU (a hung box is a lost lease).
// Returns false when the phase has not completed after ~2 s; callers then set the kernel's status word and bail out.
__device__ __forceinline__ uint64_t global_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
static __device__ __noinline__ bool mbar_wait_slow(uint32_t addr, uint32_t parity) {
    const uint64_t t0 = global_ns();
#pragma unroll 1
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try(addr, parity)) return true;                        // try_wait suspends in hardware for a bounded time
        if (spin >= 256u) __nanosleep(64);                              // long waits only: stop competing for issue slots
        if ((spin & 1023u) == 1023u && global_ns() - t0 > 2000000000ull) return false;
    }
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    if (mbar_try(addr, parity)) return true;
    if (mbar_try(addr, parity)) return true;
    return mbar_wait_slow(addr, parity);
}

// Pipeline-stall report.  status[0] = code of the last stall of the current call (the host clears it per call), status[1] = code of
// the FIRST stall since the caller zeroed the workspace header (sticky: a later, healthy call does not erase the evidence).
__device__ __forceinline__ void set_status(int* status, int code) {
    atomicExch(status, code);
    atomicCAS(status + 1, 0, code);
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// launch_dependents: the next grid of the stream may start on SMs this grid has left.  wait: block until the previous
// grid has completed and its memory is visible -- call before the first access to anything it may have produced or still read.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- proxies / fences ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 3-D tiled load: coordinates are (innermost, middle, outermost) in ELEMENTS of the tensor map's dimensions.
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
// L2 prefetch of a 2-D tile (no shared-memory destination, no barrier): brings an HBM tile into L2 ahead of the load that needs it.
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
// L2 eviction-priority hints (the fixed createpolicy encodings): streamed-once data must not push out what the next stage re-reads.
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_hint(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const void* tmap, int c0, int c1, const void* src_smem, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(src_smem)), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_store_4d_hint(const void* tmap, int c0, int c1, int c2, int c3, const void* src_smem, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem)), "l"(policy) : "memory");
}
// Tiled stores shared -> global (bulk async-group completion).  Writes outside the tensor's bounds are dropped.
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(src_smem)) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, int c0, int c1, int c2, int c3, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    // (no relinquish_alloc_permit: the chain kernel allocates again for its next stage; one CTA per SM, nobody else is waiting)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (lane_base + i), columns col .. col+31.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- tcgen05.mma --------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor of a K-major operand tile stored as rows of 64 bf16 (128 bytes) with the 128-byte
// swizzle (exactly what a TMA box of 64 x rows with CU_TENSOR_MAP_SWIZZLE_128B writes; base 1024-byte aligned):
//   bits [0,14)  start address >> 4         bits [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   bits [32,46) stride byte offset >> 4 = 1024 >> 4 (distance between 8-row groups)
//   bits [46,48) descriptor version = 1 (sm_100)      bits [61,64) layout type = 2 (SWIZZLE_128B)
// Advancing by one UMMA_K (16 bf16 = 32 bytes) inside the 128-byte row adds 2 to the start-address field.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major operand WITHOUT swizzle ("interleaved" canonical layout): core matrices of 8 rows x 16 bytes stored contiguously
// (128 bytes); sbo = byte distance between consecutive 8-row groups, lbo = byte distance between the two core matrices
// that are adjacent along K.  One UMMA_K (16 x 16-bit) spans two core matrices, so a k-step advances the start by 2*lbo.
__device__ __forceinline__ uint64_t umma_desc_k_nosw(uint32_t smem_addr, uint32_t sbo, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A and B bf16 (bits 7-9, 10-12 = 1), both K-major
// (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A and B both fp16 (format code 0) or both bf16 (format code 1)
__host__ __device__ constexpr uint32_t umma_idesc_16(int M, int N, int is_bf16) {
    return (1u << 4) | ((uint32_t)is_bf16 << 7) | ((uint32_t)is_bf16 << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when every tcgen05.mma this thread issued so far has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace sm100
}  // namespace bcl
