// Thin inline-PTX wrappers for the sm_100a features the laughter-detection kernels use:
// mbarrier, bulk async copies (UBLKCP), tcgen05 MMA / TMEM, and the matching fences.
// Nothing here is generic infrastructure: only what ld_gemm.cu needs.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace ld {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make barrier initialisation visible to the async proxy (bulk copies, tcgen05.commit)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

// Spin with a wall-clock bound: a protocol bug must trap (kills the context, reported as a CUDA
// error to the host) instead of hanging the GPU.  ~4 s at 2 GHz.
#ifndef LD_WAIT_TIMEOUT_CYCLES
#define LD_WAIT_TIMEOUT_CYCLES (8ll << 30)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > LD_WAIT_TIMEOUT_CYCLES) {
            printf("ld: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
                   blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- bulk async copy (global -> smem)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// Asks L2 to fetch `bytes` (multiple of 16) at a 16-byte aligned global address; no shared-memory destination, no barrier.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Predicated forms for a warp that runs the issue loop in uniform control flow: only the elected lane issues.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_f16_ss_pred(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate, bool issue) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(issue)));
    // (volatile, but no "memory" clobber: the MMA touches no C++-visible memory; its ordering against the mbarrier waits and
    //  commits around it -- all volatile asm -- is kept, while plain loads of the next taps may be scheduled across it)
}
// Programmatic dependent launch: the next kernel of the stream may start its prologue while this grid drains
// (launch_dependents), and must not touch memory other kernels produce or consume before wait() returns.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Dataflow between CTAs of one launch (layer-pipelined conv launches): an acquire load of a completion counter, and the proxy
// fence that orders the writes it made visible (generic proxy) before this thread's following bulk copies (async proxy).
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_cta_shared(uint32_t addr) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void umma_commit_pred(uint32_t bar, bool issue) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar), "r"(static_cast<uint32_t>(issue))
        : "memory");
}


// TMEM -> registers: 32 lanes x 32-bit, 8 consecutive columns per thread.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}

// TMEM -> registers: 32 lanes x 32-bit, 16 consecutive columns per thread.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// TMEM -> registers: 32 lanes x 32-bit, 32 consecutive columns per thread.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// registers -> TMEM: zero 16 consecutive columns of this warp's 32 lanes (accumulators are cleared by the epilogue so
// that every MMA of a tile can accumulate, whichever output columns it covers).
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(z)
        : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_zero_cols(uint32_t taddr) {
    static_assert(N % 8 == 0, "accumulator width");
#pragma unroll
    for (int c = 0; c + 16 <= N; c += 16) tmem_st16_zero(taddr + c);
    if (N % 16) tmem_st8_zero(taddr + N / 16 * 16);
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// N consecutive accumulator columns (N a multiple of 16); the caller issues tmem_wait_ld() once afterwards.
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]) {
    static_assert(N % 8 == 0 && N >= 8 && N <= 128, "accumulator width");
    constexpr int n32 = N / 32 * 32, n16 = (N - n32) / 16 * 16;
#pragma unroll
    for (int c = 0; c < n32; c += 32) tmem_ld32(taddr + c, v + c);
    if (n16) {
        uint32_t(&t)[16] = *reinterpret_cast<uint32_t(*)[16]>(v + n32);
        tmem_ld16(taddr + n32, t);
    }
    if (N - n32 - n16) tmem_ld8(taddr + n32 + n16, v + n32 + n16);
}


// ---------------------------------------------------------------- descriptors
__device__ __forceinline__ uint64_t umma_pack_desc(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved") canonical layout:
// in 16-byte units ((8,n),2):((1,SBO),LBO) -- a core matrix is 8 rows x 16 B stored contiguously
// (rows 16 B apart); SBO = byte distance between 8-row groups, LBO = byte distance between the two
// 16-byte K chunks an MMA (K=16 halfs) consumes.  version=1 (sm_100), base_offset=0, lbo_mode=0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
// Instruction descriptor for kind::f16: D=f32, A=B=f16, both K-major, M=128, N=n.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t n) {
    return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// Sum over the 32 lanes of a warp of N per-lane values, for all N channels at once: a butterfly in which the lanes
// trade halves of the vector (N/2 + N/4 + ... shuffles instead of 5 N).  Afterwards lane l holds, in vals[0 .. N/32),
// the totals of channels  i + bit0(l) N/32 + bit1(l) N/16 + bit2(l) N/8 + bit3(l) N/4 + bit4(l) N/2   (N >= 32).
template <int N>
__device__ __forceinline__ void warp_reduce_channels(float (&vals)[N], int lane) {
    static_assert(N % 32 == 0, "channel count");
    int n = N / 2;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            if (i < n) {
                const float keep = upper ? vals[i + n] : vals[i];
                const float send = upper ? vals[i] : vals[i + n];
                vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
}
__device__ __forceinline__ int warp_reduce_channel_of(int lane, int i, int n_channels) {
    const int u = n_channels / 32;
    return i + (lane & 1) * u + ((lane >> 1) & 1) * 2 * u + ((lane >> 2) & 1) * 4 * u + ((lane >> 3) & 1) * 8 * u +
           ((lane >> 4) & 1) * 16 * u;
}


}  // namespace ld
