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
constexpr int kAccStride = 64;   // TMEM columns per accumulator stage (cout <= 64)
constexpr int kTmemCols = 128;
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
    s.total = s.bar_off + (2 * kMaxStages + 5) * 8 + 16;
    return s;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_taps_kernel(const GemmLaunch* __restrict__ g_launch, int m_tiles, int M) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // The header fields needed to lay out shared memory come straight from global memory.
    const int cin = g_launch->cin, cout = g_launch->cout, n_wtaps = g_launch->n_wtaps;
    const int ext_alloc = g_launch->ext_alloc, n_stages = g_launch->n_stages;
    const GemmSmem lay = gemm_smem_layout(cin, cout, n_wtaps, ext_alloc, n_stages);

    float* s_scale = reinterpret_cast<float*>(smem + lay.param_off);
    float* s_shift = s_scale + cout;
    GemmLaunch* L = reinterpret_cast<GemmLaunch*>(smem + lay.launch_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);

    const uint32_t bar_full = smem_u32(bars);                       // [n_stages]
    const uint32_t bar_empty = smem_u32(bars + kMaxStages);         // [n_stages]
    const uint32_t bar_acc_full = smem_u32(bars + 2 * kMaxStages);  // [2]
    const uint32_t bar_acc_empty = bar_acc_full + 16;               // [2]
    const uint32_t bar_w = bar_acc_full + 32;

    {   // launch table + folded BN parameters -> smem
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g_launch);
        uint32_t* dst = reinterpret_cast<uint32_t*>(L);
        for (int i = threadIdx.x; i < static_cast<int>(sizeof(GemmLaunch) / 4); i += kGemmThreads) dst[i] = src[i];
        for (int i = threadIdx.x; i < cout; i += kGemmThreads) {
            s_scale[i] = g_launch->scale[i];
            s_shift[i] = g_launch->shift[i];
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
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
    const uint32_t lbo_b = static_cast<uint32_t>(cout) * 16u;       // bytes between channel chunks of a weight tap
    const int kchunks = cin / 8;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const uint32_t tap_bytes = static_cast<uint32_t>(cin) * cout * 2;
            mbar_expect_tx(bar_w, tap_bytes * n_wtaps);
            for (int t = 0; t < n_wtaps; ++t)
                bulk_g2s(w_addr + t * tap_bytes, reinterpret_cast<const uint8_t*>(L->weights) + static_cast<size_t>(t) * tap_bytes,
                         tap_bytes, bar_w);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const GemmJob& job = L->jobs[tile % n_jobs];
                const long long p0 = static_cast<long long>(tile / n_jobs) * kTileM;
                for (int g = 0; g < job.n_groups; ++g) {
                    const GemmGroup& grp = job.groups[g];
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t bytes = static_cast<uint32_t>(grp.ext) * 16u;
                    mbar_expect_tx(bar_full + 8 * stage, bytes * kchunks);
                    const __half* src = grp.src + (p0 + grp.shift) * 8;
                    const uint32_t dst = stage_addr0 + stage * lay.stage_bytes;
                    for (int kc = 0; kc < kchunks; ++kc)
                        bulk_g2s(dst + kc * lbo_a, src + kc * grp.kc_stride, bytes, bar_full + 8 * stage);
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(cout));
        const int ksteps = cin / 16;
        mbar_wait(bar_w, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const GemmJob& job = L->jobs[tile % n_jobs];
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            uint32_t accumulate = 0;
            int t = 0;
            for (int g = 0; g < job.n_groups; ++g) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_base = stage_addr0 + stage * lay.stage_bytes;
                    for (; t < job.n_taps && job.taps[t].group == g; ++t) {
                        const uint32_t a_tap = a_base + static_cast<uint32_t>(job.taps[t].off) * 16u;
                        const uint32_t b_tap = w_addr + static_cast<uint32_t>(job.taps[t].wtap) * kchunks * lbo_b;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            umma_f16_ss(d_tmem, umma_smem_desc(a_tap + ks * 2 * lbo_a, lbo_a, 128),
                                        umma_smem_desc(b_tap + ks * 2 * lbo_b, lbo_b, 128), idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);  // frees the smem stage when these MMAs retire
                }
                t = __shfl_sync(0xffffffffu, t, 0);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) umma_commit(bar_acc_full + 8 * acc);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        const int wp = L->wp, wp2 = L->wp2, hp = L->hp, relu = L->relu, out_mode = L->out_mode;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const GemmJob& job = L->jobs[tile % n_jobs];
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
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
            const __half* resp = (job.res != nullptr && valid && inner) ? job.res + (p + job.res_shift) * 8 : nullptr;

            mbar_wait(bar_acc_full + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;
            for (int c0 = 0; c0 < cout; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kc = (c0 >> 3) + h;
                    float r[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) r[e] = 0.f;
                    if (resp != nullptr) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(resp + kc * job.res_kc_stride);
                        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(rh[e]);
                            r[2 * e] = f.x; r[2 * e + 1] = f.y;
                        }
                    }
                    uint4 ov;
                    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = c0 + h * 8 + 2 * e;
                        float a = fmaf(__uint_as_float(v[h * 8 + 2 * e]), s_scale[c], s_shift[c]) + r[2 * e];
                        float b = fmaf(__uint_as_float(v[h * 8 + 2 * e + 1]), s_scale[c + 1], s_shift[c + 1]) + r[2 * e + 1];
                        if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        if (!inner) { a = 0.f; b = 0.f; }
                        oh[e] = __floats2half2_rn(a, b);
                    }
                    if (do_store) *reinterpret_cast<uint4*>(dst + kc * job.out_kc_stride) = ov;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Host launcher.  `h` is the host copy of the launch description that lives at `d_launch`.
cudaError_t launch_gemm_taps(const GemmLaunch* d_launch, const GemmLaunch& h, int m_tiles, int M, int num_sms,
                             cudaStream_t stream) {
    static bool attr_set = false;
    const GemmSmem lay = gemm_smem_layout(h.cin, h.cout, h.n_wtaps, h.ext_alloc, h.n_stages);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long total = static_cast<long long>(m_tiles) * h.n_jobs;
    if (total <= 0) return cudaSuccess;
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    gemm_taps_kernel<<<grid, kGemmThreads, lay.total, stream>>>(d_launch, m_tiles, M);
    return cudaGetLastError();
}

// Chooses the smem ring depth for a launch (host side).
int gemm_pick_stages(int cin, int cout, int n_wtaps, int ext_alloc) {
    for (int n = kMaxStages; n >= 2; --n)
        if (gemm_smem_layout(cin, cout, n_wtaps, ext_alloc, n).total <= 227u * 1024u) return n;
    return 0;
}

}  // namespace ld
