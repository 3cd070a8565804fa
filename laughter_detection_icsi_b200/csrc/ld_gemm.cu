// K2: shifted-plane ("tap list") implicit-GEMM convolution on tcgen05 tensor cores.
//
//   out_j[p, :] = act( scale * sum_t  src_t[p + shift_t, :] @ W[wtap_t]^T  + shift (+ res[p + rs, :]) )
//
// for every output pixel p of every job j of one conv layer (a job = one output plane, see ld_plan.cpp).
// Replaces the cuDNN/MKLDNN conv2d + BatchNorm2d + ReLU (+ residual add) calls made by
// ResidualBlock.forward / ResNetBigger.forward (reference models.py:110-115, 222-228).
//
// Data layout: activations are fp16, channel-chunk planar [C/8][pixels][8].  A run of 8 consecutive
// pixels of one chunk is 128 contiguous bytes = exactly one UMMA "core matrix" of the K-major
// SWIZZLE_NONE canonical layout, so an A-operand tile that starts at ANY pixel offset is a legal smem
// matrix descriptor: one bulk load of (128 + span) pixels serves all taps that differ only by a pixel
// shift (the three kx taps of a conv row, or all nine taps of an interior job).
//
// Roles (192 threads, one persistent CTA per SM):
//   warp 0      producer: cp.async.bulk (UBLKCP) global -> smem ring, mbarrier complete_tx
//   warp 1      MMA issuer: tcgen05.mma kind::f16, M=128 N=cout K=16, accumulators in TMEM (2 stages)
//   warps 2..5  epilogue: tcgen05.ld -> BN scale/shift, residual, ReLU, pad masking -> fp16 stores
#include <cstdio>

#include "ld_ptx.cuh"
#include "ld_types.h"

namespace ld {

constexpr int kGemmThreads = 192;
constexpr int kAccStages = 4;    // TMEM accumulator ring (tile i+3 can be multiplied while tile i is still being stored)
constexpr int kAccStride = 64;   // TMEM columns per accumulator stage (cout <= 64)
constexpr int kTmemCols = kAccStages * kAccStride;
constexpr int kMaxStages = 6;

struct GemmSmem {
    uint32_t w_off, stage_off, stage_bytes, param_off, launch_off, bar_off, total;
};

__host__ __device__ inline GemmSmem gemm_smem_layout(int cin, int cout, int n_wtaps, int ext_alloc,
                                                     int n_stages) {
    GemmSmem s;
    s.w_off = 0;
    uint32_t w_bytes = static_cast<uint32_t>(n_wtaps) * cin * cout * 2;
    s.stage_off = (w_bytes + 127u) & ~127u;
    s.stage_bytes = static_cast<uint32_t>(ext_alloc) * 16u * (cin / 8);
    s.param_off = s.stage_off + n_stages * s.stage_bytes;
    s.launch_off = s.param_off + 2u * cout * sizeof(float);
    s.launch_off = (s.launch_off + 15u) & ~15u;
    s.bar_off = s.launch_off + static_cast<uint32_t>(sizeof(GemmLaunch));
    s.bar_off = (s.bar_off + 15u) & ~15u;
    s.total = s.bar_off + (2 * kMaxStages + 2 * kAccStages + 1) * 8 + 16;
    return s;
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_taps_kernel(const GemmLaunch* __restrict__ g_launch, int m_tiles, int M) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kChunks = CIN / 8;   // 16-byte channel chunks per pixel
    constexpr int kSteps = CIN / 16;   // MMAs (K = 16) per tap

    // The header fields needed to lay out shared memory come straight from global memory.
    const int n_wtaps = g_launch->n_wtaps;
    const int ext_alloc = g_launch->ext_alloc, n_stages = g_launch->n_stages;
    const GemmSmem lay = gemm_smem_layout(CIN, COUT, n_wtaps, ext_alloc, n_stages);

    float* s_scale = reinterpret_cast<float*>(smem + lay.param_off);
    float* s_shift = s_scale + COUT;
    GemmLaunch* L = reinterpret_cast<GemmLaunch*>(smem + lay.launch_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kAccStages + 1);

    const uint32_t bar_full = smem_u32(bars);                       // [n_stages]
    const uint32_t bar_empty = smem_u32(bars + kMaxStages);         // [n_stages]
    const uint32_t bar_acc_full = smem_u32(bars + 2 * kMaxStages);  // [kAccStages]
    const uint32_t bar_acc_empty = bar_acc_full + 8 * kAccStages;   // [kAccStages]
    const uint32_t bar_w = bar_acc_empty + 8 * kAccStages;

    {   // launch table + folded BN parameters -> smem
        const uint4* src = reinterpret_cast<const uint4*>(g_launch);
        uint4* dst = reinterpret_cast<uint4*>(L);
        for (int i = threadIdx.x; i < static_cast<int>(sizeof(GemmLaunch) / 16); i += kGemmThreads) dst[i] = src[i];
        for (int i = threadIdx.x; i < COUT; i += kGemmThreads) {
            s_scale[i] = g_launch->scale[i];
            s_shift[i] = g_launch->shift[i];
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(bar_acc_full + 8 * i, 1);
            mbar_init(bar_acc_empty + 8 * i, 4);
        }
        mbar_init(bar_w, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_jobs = L->n_jobs;
    const int total_tiles = m_tiles * n_jobs;
    const uint32_t w_addr = smem_u32(smem + lay.w_off);
    const uint32_t stage_addr0 = smem_u32(smem + lay.stage_off);
    const uint32_t lbo_a = static_cast<uint32_t>(ext_alloc) * 16u;  // bytes between channel chunks in a stage

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            constexpr uint32_t tap_bytes = static_cast<uint32_t>(CIN) * COUT * 2;
            mbar_expect_tx(bar_w, tap_bytes * n_wtaps);
            for (int t = 0; t < n_wtaps; ++t)
                bulk_g2s(w_addr + t * tap_bytes, reinterpret_cast<const uint8_t*>(L->weights) + static_cast<size_t>(t) * tap_bytes,
                         tap_bytes, bar_w);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const GemmJob& job = L->jobs[tile % n_jobs];
            const long long p0 = static_cast<long long>(tile / n_jobs) * kTileM;
            const int n_groups = job.n_groups;
            for (int g = 0; g < n_groups; ++g) {
                const GemmGroup& grp = job.groups[g];
                const uint32_t bytes = static_cast<uint32_t>(grp.ext) * 16u;
                const uint32_t full = bar_full + 8 * stage;
                if (lane == 0) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(full, bytes * kChunks);
                }
                __syncwarp();
                if (lane < kChunks)   // one bulk copy per channel chunk, issued by kChunks lanes in parallel
                    bulk_g2s(stage_addr0 + stage * lay.stage_bytes + lane * lbo_a,
                             grp.src + (p0 + grp.shift) * 8 + lane * grp.kc_stride, bytes, full);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(COUT));
        // smem matrix descriptors (see ld_ptx.cuh): only the 14-bit start address in the low word changes per MMA.
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);            // SBO = 128 B, descriptor version 1
        const uint32_t a_lo0 = (static_cast<uint32_t>(ext_alloc) << 16);  // LBO = ext_alloc * 16 B
        const uint32_t b_lo0 = (static_cast<uint32_t>(COUT) << 16) | (w_addr >> 4);  // LBO = COUT * 16 B
        const uint32_t a_kstep = 2u * static_cast<uint32_t>(ext_alloc);   // two channel chunks per K = 16
        constexpr uint32_t b_kstep = 2u * COUT;
        mbar_wait(bar_w, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const GemmJob& job = L->jobs[tile % n_jobs];
            mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            uint32_t accumulate = 0;
            int t = 0;
            const int n_groups = job.n_groups;
            for (int g = 0; g < n_groups; ++g) {
                const int nt = job.group_taps[g];
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_stage = a_lo0 | ((stage_addr0 + stage * lay.stage_bytes) >> 4);
                    for (int i = 0; i < nt; ++i) {
                        const uint32_t a_lo = a_stage + job.tap_a16[t + i];
                        const uint32_t b_lo = b_lo0 + job.tap_b16[t + i];
#pragma unroll
                        for (int ks = 0; ks < kSteps; ++ks) {
                            umma_f16_ss(d_tmem, umma_pack_desc(a_lo + ks * a_kstep, desc_hi),
                                        umma_pack_desc(b_lo + ks * b_kstep, desc_hi), idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);  // frees the smem stage when these MMAs retire
                }
                t += nt;
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) umma_commit(bar_acc_full + 8 * acc);
            __syncwarp();
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        const int wp = L->wp, wp2 = L->wp2, hp = L->hp, relu = L->relu, out_mode = L->out_mode;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const GemmJob& job = L->jobs[tile % n_jobs];
            const long long p = static_cast<long long>(tile / n_jobs) * kTileM + q * 32 + lane;
            const long long row = p / wp;
            const int col = static_cast<int>(p - row * wp);
            const bool valid = p < M;
            bool inner = col >= 1 && col <= wp - 2;
            if (hp > 0) {
                const int ri = static_cast<int>(row % hp);
                inner = inner && ri >= 1 && ri <= hp - 2;
            }
            __half* dst;
            bool do_store;
            if (out_mode == OUT_PLAIN) {
                dst = job.out0 + p * 8;
                do_store = valid;
            } else {
                const int c0 = col - 1;
                dst = ((c0 & 1) ? job.out1 : job.out0) + (row * wp2 + (c0 >> 1) + 1) * 8;
                do_store = valid && inner;
            }
            const long long out_kc = job.out_kc_stride;
            // residual: fetched while the MMAs of this tile are still running
            uint4 res[COUT / 8];
            const bool has_res = job.res != nullptr && valid && inner;
            if (has_res) {
                const __half* resp = job.res + (p + job.res_shift) * 8;
                const long long rs = job.res_kc_stride;
#pragma unroll
                for (int kc = 0; kc < COUT / 8; ++kc) res[kc] = ld_nc_u4(resp + kc * rs);
            } else {
#pragma unroll
                for (int kc = 0; kc < COUT / 8; ++kc) res[kc] = make_uint4(0, 0, 0, 0);
            }

            mbar_wait(bar_acc_full + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;
            uint32_t v[COUT];
            tmem_ld_cols<COUT>(taddr, v);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);  // accumulator is in registers: hand the stage back
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }

            if (do_store) {
#pragma unroll
                for (int kc = 0; kc < COUT / 8; ++kc) {
                    const __half2* rh = reinterpret_cast<const __half2*>(&res[kc]);
                    uint4 ov;
                    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = kc * 8 + 2 * e;
                        const float2 r = __half22float2(rh[e]);
                        float a = fmaf(__uint_as_float(v[c]), s_scale[c], s_shift[c]) + r.x;
                        float b = fmaf(__uint_as_float(v[c + 1]), s_scale[c + 1], s_shift[c + 1]) + r.y;
                        if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        if (!inner) { a = 0.f; b = 0.f; }
                        oh[e] = __floats2half2_rn(a, b);
                    }
                    *reinterpret_cast<uint4*>(dst + kc * out_kc) = ov;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Host launcher.  `h` is the host copy of the launch description that lives at `d_launch`.
template <int CIN, int COUT>
static cudaError_t launch_typed(const GemmLaunch* d_launch, const GemmLaunch& h, int m_tiles, int M, int num_sms,
                                cudaStream_t stream) {
    static bool attr_set = false;
    const GemmSmem lay = gemm_smem_layout(CIN, COUT, h.n_wtaps, h.ext_alloc, h.n_stages);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_taps_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long total = static_cast<long long>(m_tiles) * h.n_jobs;
    if (total <= 0) return cudaSuccess;
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    gemm_taps_kernel<CIN, COUT><<<grid, kGemmThreads, lay.total, stream>>>(d_launch, m_tiles, M);
    return cudaGetLastError();
}

cudaError_t launch_gemm_taps(const GemmLaunch* d_launch, const GemmLaunch& h, int m_tiles, int M, int num_sms,
                             cudaStream_t stream) {
#define LD_GEMM_CASE(ci, co) \
    if (h.cin == ci && h.cout == co) return launch_typed<ci, co>(d_launch, h, m_tiles, M, num_sms, stream)
    LD_GEMM_CASE(64, 64); LD_GEMM_CASE(64, 48); LD_GEMM_CASE(64, 32); LD_GEMM_CASE(64, 16);
    LD_GEMM_CASE(48, 64); LD_GEMM_CASE(48, 48); LD_GEMM_CASE(48, 32); LD_GEMM_CASE(48, 16);
    LD_GEMM_CASE(32, 64); LD_GEMM_CASE(32, 48); LD_GEMM_CASE(32, 32); LD_GEMM_CASE(32, 16);
    LD_GEMM_CASE(16, 64); LD_GEMM_CASE(16, 48); LD_GEMM_CASE(16, 32); LD_GEMM_CASE(16, 16);
#undef LD_GEMM_CASE
    return cudaErrorInvalidValue;
}

// Chooses the smem ring depth for a launch (host side).
int gemm_pick_stages(int cin, int cout, int n_wtaps, int ext_alloc) {
    for (int n = kMaxStages; n >= 2; --n)
        if (gemm_smem_layout(cin, cout, n_wtaps, ext_alloc, n).total <= 227u * 1024u) return n;
    return 0;
}

}  // namespace ld
