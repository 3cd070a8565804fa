// Weight gradient of a conv layer on tcgen05 tensor cores (training path, reference train.py:289 loss.backward()):
//
//   dW[co][ci][tap] = sum over pixels p of  X_tap[p + shift_tap][ci] * dZ[p][co]
//
// is a GEMM whose REDUCTION dimension is the pixel index.  Activations and gradients live in channel-chunk planar
// planes [C/8][pixels][8]; 8 channels x 8 pixels of one chunk are 128 contiguous bytes with the channel index fastest,
// which is exactly a core matrix of the UMMA *MN-major* SWIZZLE_NONE canonical layout
//     ((8 ch, m groups), (8 px, k groups)) : ((1, SBO), (16 B, LBO))      SBO = bytes between channel chunks, LBO = 128 B
// so both operands are used straight from the staged pixel runs, no transposition:
//     A (M x K) = X^T : M = channels of several taps stacked (rows), K = 16 pixels per MMA
//     B (N x K) = dZ^T: N = cout
//     D (M x N) in TMEM, fp32, accumulated over ALL pixel tiles of the CTA (split-K over CTAs), then added to global
//     memory with fp32 atomics.
// M = 128 rows hold 128 / CIN consecutive "segments" (staged pixel runs of different taps, laid out back to back with the
// same chunk stride, so that the chunk index runs uniformly across them); rows that fall beyond the real segments
// multiply garbage and are ignored by the epilogue (rows of D are independent).
//
// Roles (192 threads, one persistent CTA per SM): warp 0 producer (cp.async.bulk), warp 1 MMA issuer,
// warps 2..5 final epilogue (tcgen05.ld -> atomicAdd).
#include <cuda_bf16.h>

#include "ld_ptx.cuh"
#include "ld_train.h"

namespace ld {

constexpr int kWgThreads = 192;
constexpr int kWgTile = 128;      // pixels per tile = 8 K-steps of 16
constexpr int kWgSegPx = 136;     // pixels staged per segment (tile + up to 2 pixels of column shift, rounded to 8)

__host__ __device__ inline uint32_t wg_seg_bytes(int cin) { return static_cast<uint32_t>(cin / 8) * kWgSegPx * 16u; }
__host__ __device__ inline uint32_t wg_dz_bytes(int cout) { return static_cast<uint32_t>(cout / 8) * kWgTile * 16u; }
// A stage holds the real segments, then the dZ tile; the garbage rows of the last MMA of a stack may read up to
// (128 / cin) segments from its first segment, so the stage is at least that long.
__host__ __device__ inline uint32_t wg_stage_bytes(int cin, int cout, int n_seg, int max_seg0) {
    const uint32_t real = n_seg * wg_seg_bytes(cin) + wg_dz_bytes(cout);
    const uint32_t reach = static_cast<uint32_t>(max_seg0 + 128 / cin) * wg_seg_bytes(cin) + 64u;
    const uint32_t b = real > reach ? real : reach;
    return (b + 127u) & ~127u;
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_mma_kernel(const __grid_constant__ WgradLaunch L) {
    extern __shared__ __align__(128) uint8_t smem[];
    // (warp index through a shuffle: the compiler then keeps the issuing warp's descriptor arithmetic in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    constexpr int kSegPerMma = 128 / CIN;
    constexpr uint32_t seg_bytes = (CIN / 8) * kWgSegPx * 16u;
    constexpr uint32_t dz_bytes = (COUT / 8) * kWgTile * 16u;
    const int n_stages = L.n_stages;
    const uint32_t stage_bytes = L.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(n_stages) * stage_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * 4 + 1);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * 4, bar_done = bar_empty + 8 * 4;
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t stage0 = smem_u32(smem);
    const long long tiles = (L.M + kWgTile - 1) / kWgTile;
    const int my_tiles = static_cast<int>((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        int stage = 0;
        uint32_t phase = 0;
        const int n_seg = L.n_seg;
        constexpr int kxc = CIN / 8, kdc = COUT / 8;
        for (int it = 0; it < my_tiles; ++it) {
            const long long p0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x) * kWgTile;
            const uint32_t full = bar_full + 8 * stage;
            if (lane == 0) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                mbar_expect_tx(full, n_seg * seg_bytes + dz_bytes);
            }
            __syncwarp();
            const uint32_t base = stage0 + stage * stage_bytes;
            for (int c = lane; c < n_seg * kxc + kdc; c += 32) {
                if (c < n_seg * kxc) {
                    const int s = c / kxc, kc = c - s * kxc;
                    bulk_g2s(base + s * seg_bytes + kc * (kWgSegPx * 16u),
                             L.seg_src[s] + (p0 + L.seg_shift[s]) * 8 + kc * L.seg_kc_stride[s], kWgSegPx * 16u, full);
                } else {
                    const int kc = c - n_seg * kxc;
                    bulk_g2s(base + n_seg * seg_bytes + kc * (kWgTile * 16u), L.dz + p0 * 8 + kc * L.dz_kc_stride, kWgTile * 16u, full);
                }
            }
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // kind::f16, bf16 A/B, fp32 D, A and B MN-major (bits 15, 16), M = 128, N = COUT
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((COUT >> 3) << 17) | ((128u >> 4) << 24);
        // descriptors: LBO (bits 16..29) = 128 B between 8-pixel groups, SBO (bits 32..45) = bytes between channel chunks
        constexpr uint32_t a_hi = ((kWgSegPx * 16u) >> 4) | (1u << 14);
        constexpr uint32_t b_hi = ((kWgTile * 16u) >> 4) | (1u << 14);
        constexpr uint32_t lo_lbo = (128u >> 4) << 16;
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        const int n_grp = L.n_grp;
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t base = stage0 + stage * stage_bytes;
            const uint32_t dz_addr = base + L.n_seg * seg_bytes;
#pragma unroll
            for (int ks = 0; ks < kWgTile / 16; ++ks) {
                const uint32_t b_lo = lo_lbo | ((dz_addr + ks * 256u) >> 4);
                for (int g = 0; g < n_grp; ++g) {
                    const uint32_t a_addr = base + L.grp_seg0[g] * seg_bytes + (L.grp_px_off[g] + ks * 16) * 16u;
                    umma_f16_ss_pred(tmem_base + g * COUT, umma_pack_desc(lo_lbo | (a_addr >> 4), a_hi), umma_pack_desc(b_lo, b_hi),
                                     idesc, (it > 0 || ks > 0) ? 1u : 0u, leader);
                }
            }
            umma_commit_pred(bar_empty + 8 * stage, leader);
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pred(bar_done, leader);
    } else {
        // ------------------------------------------------------------------ epilogue: D -> global atomics
        const int q = warp & 3;
        mbar_wait(bar_done, 0);
        tc_fence_after();
        if (my_tiles > 0) {
            const int row = q * 32 + lane;              // row of D = stacked (segment, channel)
            const int s = row / CIN, ci = row - s * CIN;
            for (int g = 0; g < L.n_grp; ++g) {
                const int tap = L.grp_tap[g][s < kSegPerMma ? s : 0];
                uint32_t v[COUT];
                tmem_ld_cols<COUT>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * COUT, v);
                tmem_wait_ld();
                if (tap >= 0) {
#pragma unroll
                    for (int co = 0; co < COUT; ++co)
                        atomicAdd(L.dw + (static_cast<long long>(co) * CIN + ci) * L.n_taps + tap, __uint_as_float(v[co]));
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int CIN, int COUT>
static cudaError_t launch_wg(const WgradLaunch& L, int num_sms, cudaStream_t stream) {
    static PerDeviceOnce attr_set;
    if (!attr_set.flag()) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_mma_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set.flag() = true;
    }
    const long long tiles = (L.M + kWgTile - 1) / kWgTile;
    if (tiles <= 0) return cudaSuccess;
    const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
    const size_t smem = static_cast<size_t>(L.n_stages) * L.stage_bytes + (2 * 4 + 1) * 8 + 16;
    wgrad_mma_kernel<CIN, COUT><<<grid, kWgThreads, smem, stream>>>(L);
    return cudaGetLastError();
}

cudaError_t launch_wgrad_mma(const WgradLaunch& L, int num_sms, cudaStream_t stream) {
#define LD_WG(ci, co) if (L.cin == ci && L.cout == co) return launch_wg<ci, co>(L, num_sms, stream)
    LD_WG(64, 64); LD_WG(64, 32); LD_WG(32, 32); LD_WG(32, 16); LD_WG(16, 16);
#undef LD_WG
    return cudaErrorInvalidValue;
}

// Fills n_stages / stage_bytes (host).  Returns false when not even one stage fits.
bool wgrad_plan_smem(WgradLaunch& L) {
    int max_seg0 = 0;
    for (int g = 0; g < L.n_grp; ++g) max_seg0 = L.grp_seg0[g] > max_seg0 ? L.grp_seg0[g] : max_seg0;
    L.stage_bytes = wg_stage_bytes(L.cin, L.cout, L.n_seg, max_seg0);
    const uint32_t budget = 227u * 1024u - 256u;
    L.n_stages = static_cast<int>(budget / L.stage_bytes);
    if (L.n_stages > 4) L.n_stages = 4;
    return L.n_stages >= 1;
}

}  // namespace ld
