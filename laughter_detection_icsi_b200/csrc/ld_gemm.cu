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
//   warps 0..1  producers: cp.async.bulk (UBLKCP) global -> smem ring, mbarrier complete_tx
//   warps 2..5  MMA issuers: tcgen05.mma kind::f16, M=128 N=cout K=16, accumulators in TMEM; tile i of the CTA belongs
//               to issuer i % 4 (issuing one MMA costs ~25 dependent instructions of a single thread, far more than
//               the 8..32 cycles a small-N MMA occupies the tensor pipe, so four tiles are issued concurrently)
//   warps 6..9  epilogue: tcgen05.ld -> BN shift (scale is folded into the weights), ReLU, pad masking -> fp16 stores
// The residual add of a ResidualBlock is one more tap: the residual plane times an identity weight slab, accumulated in
// fp32 by the tensor core (exact), so it travels through the same TMA/smem pipeline as every other operand.
// The launch description is a __grid_constant__ parameter, so tile/job/tap bookkeeping runs on the uniform datapath.
#include <cstdio>

#include <cuda_bf16.h>

#include "ld_ptx.cuh"
#include "ld_types.h"

namespace ld {

constexpr int kIssuers = 4;      // MMA-issuing warps: tile i of a CTA is issued by warp 1 + i % 4 into accumulator stage i % 4
constexpr int kProducers = 2;    // producer warps: in tile-stage mode they alternate tiles (stage parity = tile parity)
constexpr int kGemmThreads = 32 * (kProducers + kIssuers + 4);
constexpr int kAccStages = kIssuers;  // TMEM accumulator ring: one stage per issuer warp
constexpr int kAccStride = 64;   // TMEM columns per accumulator stage (cout <= 64)
constexpr int kTmemCols = kAccStages * kAccStride;
constexpr int kMaxStages = 16;

// Optional per-launch cycle counters (GemmLaunch::prof), summed over CTAs:
//   0 producer: waiting for a free smem stage      1 MMA warp: waiting for operands      2 MMA warp: waiting for a free accumulator
//   3 MMA warp: issuing                            4 epilogue warp 0: waiting for MMAs   5 epilogue warp 0: converting + storing
//   6 CTA lifetime                                 7 tiles
enum { PROF_PROD_WAIT = 0, PROF_MMA_WAIT_FULL, PROF_MMA_WAIT_ACC, PROF_MMA_ISSUE, PROF_EPI_WAIT, PROF_EPI_WORK, PROF_CTA, PROF_TILES };

struct GemmSmem {
    uint32_t w_off, stage_off, stage_bytes, param_off, bar_off, total;
};

__host__ __device__ inline GemmSmem gemm_smem_layout(int cin, int cout, int n_wtaps, int ext_alloc, int groups_per_stage,
                                                     int n_stages) {
    GemmSmem s;
    s.w_off = 0;
    uint32_t w_bytes = static_cast<uint32_t>(n_wtaps) * cin * cout * 2;
    s.stage_off = (w_bytes + 127u) & ~127u;
    s.stage_bytes = static_cast<uint32_t>(ext_alloc) * 16u * (cin / 8) * groups_per_stage;
    s.param_off = s.stage_off + n_stages * s.stage_bytes;
    s.bar_off = s.param_off + 2u * static_cast<uint32_t>((cout + 31) / 32 * 32) * sizeof(float);
    s.bar_off = (s.bar_off + 15u) & ~15u;
    s.total = s.bar_off + (2 * kMaxStages + 2 * kAccStages + 1 + kIssuers) * 8 + 16;
    return s;
}

// Walks the tiles of one CTA: tile = blockIdx.x + k * gridDim.x, job = tile % n_jobs, m-tile = tile / n_jobs,
// without a division per step.
struct TileWalk {
    int job, mt, step_q, step_r, n_jobs;
    __device__ TileWalk(int first_tile, int stride, int n_jobs_) : n_jobs(n_jobs_) {
        job = first_tile % n_jobs; mt = first_tile / n_jobs;
        step_q = stride / n_jobs; step_r = stride % n_jobs;
    }
    __device__ void next() {
        job += step_r; mt += step_q;
        if (job >= n_jobs) { job -= n_jobs; ++mt; }
    }
};

// MODE 0: inference -- fp16 operands, epilogue = + folded BatchNorm shift, ReLU, fp16 store (PLAIN or COLSPLIT).
// MODE 1: training  -- bf16 operands, epilogue = raw conv output as bf16 (PLAIN) and, when L.stats is set, per-channel
//                      sum / sum of squares of the fp32 accumulators over the real pixels (BatchNorm batch statistics).
template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_taps_kernel(const __grid_constant__ GemmLaunch L, int m_tiles, int M) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kChunks = CIN / 8;   // 16-byte channel chunks per pixel
    constexpr int kSteps = CIN / 16;   // MMAs (K = 16) per tap
    const long long t_cta0 = clock64();

    const int n_wtaps = L.n_wtaps, ext_alloc = L.ext_alloc, n_stages = L.n_stages, gps = L.groups_per_stage;
    const GemmSmem lay = gemm_smem_layout(CIN, COUT, n_wtaps, ext_alloc, gps, n_stages);

    float* s_shift = reinterpret_cast<float*>(smem + lay.param_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kAccStages + 1 + kIssuers);

    const uint32_t bar_full = smem_u32(bars);                       // [n_stages]
    const uint32_t bar_empty = smem_u32(bars + kMaxStages);         // [n_stages]
    const uint32_t bar_acc_full = smem_u32(bars + 2 * kMaxStages);  // [kAccStages]
    const uint32_t bar_acc_empty = bar_acc_full + 8 * kAccStages;   // [kAccStages]
    const uint32_t bar_w = bar_acc_empty + 8 * kAccStages;
    const uint32_t bar_turn = bar_w + 8;                            // [kIssuers] issuer i may start waiting for operands

    // MODE 0: s_shift[COUT] = folded shift.  MODE 1: s_shift[2 * max(COUT, 32)] = per-CTA channel sums, flushed at the end.
    constexpr int kStatN = (COUT + 31) / 32 * 32;
    if constexpr (MODE == 0) {
        for (int i = threadIdx.x; i < COUT; i += kGemmThreads) s_shift[i] = L.shift[i];
    } else {
        for (int i = threadIdx.x; i < 2 * kStatN; i += kGemmThreads) s_shift[i] = 0.f;
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
        for (int i = 0; i < kIssuers; ++i) mbar_init(bar_turn + 8 * i, 1);
        mbar_fence_init();
    }
    if (warp == kProducers) {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_jobs = L.n_jobs;
    const int total_tiles = m_tiles * n_jobs;
    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t w_addr = smem_u32(smem + lay.w_off);
    const uint32_t stage_addr0 = smem_u32(smem + lay.stage_off);
    const uint32_t box_bytes = static_cast<uint32_t>(ext_alloc) * 16u * kChunks;  // one group's operand block
    unsigned long long* prof = L.prof;
    const bool profiling = prof != nullptr;

    if (warp < kProducers) {
        // ------------------------------------------------------------------ producers
        if (warp == 0 && lane == 0) {
            constexpr uint32_t tap_bytes = static_cast<uint32_t>(CIN) * COUT * 2;
            mbar_expect_tx(bar_w, tap_bytes * n_wtaps);
            for (int t = 0; t < n_wtaps; ++t)
                bulk_g2s(w_addr + t * tap_bytes, reinterpret_cast<const uint8_t*>(L.weights) + static_cast<size_t>(t) * tap_bytes,
                         tap_bytes, bar_w);
        }
        const int loader = L.loader;
        const uint32_t lbo_a = static_cast<uint32_t>(ext_alloc) * 16u;  // bytes between channel chunks of a group
        // Two independent pipelines: producer w fills ring w (stages [w * ring_n, (w + 1) * ring_n)) with the tiles
        // it = w, w + 2, ... of the CTA; issuers w and w + 2 drain it.  A ring is filled and drained in tile order by
        // one producer, so nobody ever waits more than one phase ahead on its parity-tracked barriers.
        // (n_rings == 1 when half the stages could not hold one tile's groups: producer 0 and all four issuers share one ring)
        const int n_rings = L.n_rings;
        const int ring_n = n_stages / n_rings;
        const int ring0 = warp * ring_n;
        int stage = 0;       // position inside this producer's ring
        uint32_t phase = 0;
        long long c_wait = 0;
        TileWalk tw(blockIdx.x + warp * gridDim.x, n_rings * gridDim.x, n_jobs);
        for (int it = warp; it < my_tiles && warp < n_rings; it += n_rings, tw.next()) {
            const GemmJob& job = L.jobs[tw.job];
            const int p0 = tw.mt * kTileM;
            const int n_groups = job.n_groups;
            // lanes work in parallel: lane g opens stage g of the tile (wait until free, arm the byte count), then every
            // lane issues its share of the n_groups x kChunks copies
            const long long t0 = profiling ? clock64() : 0;
            if (gps > 1) {
                if (lane == 0) {
                    mbar_wait(bar_empty + 8 * (ring0 + stage), phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * (ring0 + stage), box_bytes * n_groups);
                }
            } else if (lane < n_groups) {
                int sg = stage + lane;
                uint32_t ph = phase;
                if (sg >= ring_n) { sg -= ring_n; ph ^= 1; }
                mbar_wait(bar_empty + 8 * (ring0 + sg), ph ^ 1);
                mbar_expect_tx(bar_full + 8 * (ring0 + sg), box_bytes);
            }
            if (profiling) c_wait += clock64() - t0;
            __syncwarp();
            const int n_copies = loader == 1 ? n_groups : n_groups * kChunks;
            for (int c = lane; c < n_copies; c += 32) {
                const int g = loader == 1 ? c : c / kChunks;
                const int kc = loader == 1 ? 0 : c - g * kChunks;
                int sg = stage;
                uint32_t dst;
                if (gps > 1) {
                    dst = stage_addr0 + (ring0 + stage) * lay.stage_bytes + g * box_bytes;
                } else {
                    sg += g;
                    if (sg >= ring_n) sg -= ring_n;
                    dst = stage_addr0 + (ring0 + sg) * lay.stage_bytes;
                }
                sg += ring0;
                const GemmGroup& grp = job.groups[g];
                if (loader == 1)   // ONE tensor copy brings the (8 halfs x box pixels x C/8 chunks) box
                    tma_load_3d(dst, grp.tmap, 0, grp.pixel0 + p0 + grp.shift, 0, bar_full + 8 * sg);
                else               // one contiguous bulk copy per channel chunk
                    bulk_g2s(dst + kc * lbo_a, grp.src + static_cast<long long>(p0 + grp.shift) * 8 + kc * grp.kc_stride, lbo_a,
                             bar_full + 8 * sg);
            }
            stage += gps > 1 ? 1 : n_groups;
            if (stage >= ring_n) { stage -= ring_n; phase ^= 1; }
        }
        if (profiling && lane == 0 && warp == 0) atomicAdd(prof + PROF_PROD_WAIT, static_cast<unsigned long long>(c_wait));
    } else if (warp < kProducers + kIssuers) {
        // ------------------------------------------------------------------ MMA issuers (uniform control flow per warp)
        const int iw = warp - kProducers;  // this warp issues tiles iw, iw + 4, ... of the CTA into accumulator stage iw
        constexpr uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(COUT)) | (MODE == 1 ? ((1u << 7) | (1u << 10)) : 0u);  // bf16 A/B
        // smem matrix descriptors (see ld_ptx.cuh): only the 14-bit start address in the low word changes per MMA.
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);            // SBO = 128 B, descriptor version 1
        const uint32_t a_lo0 = (static_cast<uint32_t>(ext_alloc) << 16);  // LBO = ext_alloc * 16 B
        const uint32_t b_lo0 = (static_cast<uint32_t>(COUT) << 16) | (w_addr >> 4);  // LBO = COUT * 16 B
        const uint32_t a_kstep = 2u * static_cast<uint32_t>(ext_alloc);   // two channel chunks per K = 16
        constexpr uint32_t b_kstep = 2u * COUT;
        const uint32_t stage16 = lay.stage_bytes >> 4;
        const bool leader = elect_one();
        mbar_wait(bar_w, 0);
        // issuer iw drains ring iw % 2 together with issuer iw ^ 2: the ring's tiles (it = ring, ring + 2, ...) alternate
        // between the two; the partner's tiles are skipped by advancing the ring position by their stage count
        const int n_rings = L.n_rings;
        const int ipr = kIssuers / n_rings;      // issuers per ring
        const int ring = iw % n_rings, slot = iw / n_rings;
        const int next_issuer = ring + n_rings * ((slot + 1) % ipr);
        const int ring_n = n_stages / n_rings;
        const int ring0 = ring * ring_n;
        int stage = 0;
        uint32_t phase = 0;
        auto advance = [&](int n) {
            stage += n;
            while (stage >= ring_n) { stage -= ring_n; phase ^= 1; }
        };
        const int acc = iw;
        uint32_t acc_phase = 0;
        // "full" barriers are parity-tracked, so a warp must never wait on a stage more than one fill ahead: the two
        // issuers of a ring take turns -- one starts waiting for operands only after the other has seen its last stage
        uint32_t turn_phase = 0;
        long long c_full = 0, c_acc = 0, c_issue = 0;
        TileWalk tw(blockIdx.x + ring * gridDim.x, n_rings * gridDim.x, n_jobs);
        for (int it = ring, k = 0; it < my_tiles; it += n_rings, ++k, tw.next()) {
            const GemmJob& job = L.jobs[tw.job];
            if ((k % ipr) != slot) {
                advance(gps > 1 ? 1 : job.n_groups);
                continue;
            }
            // the whole tap program of the job in registers (three 16-byte constant loads, issued before the waits)
            uint32_t tp[kTapWords];
#pragma unroll
            for (int i = 0; i < kTapWords / 4; ++i) {
                const uint4 w4 = reinterpret_cast<const uint4*>(job.tapw)[i];
                tp[4 * i] = w4.x; tp[4 * i + 1] = w4.y; tp[4 * i + 2] = w4.z; tp[4 * i + 3] = w4.w;
            }
            const int n_taps = job.n_taps;
            long long t0 = profiling ? clock64() : 0;
            if (k > 0) {
                mbar_wait(bar_turn + 8 * iw, turn_phase);
                turn_phase ^= 1;
            }
            mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
            if (profiling) { const long long t1 = clock64(); c_acc += t1 - t0; t0 = t1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            uint32_t accumulate = 0;
            uint32_t a_stage = 0;
#pragma unroll
            for (int t = 0; t < kMaxTaps; ++t) {
                if (t < n_taps) {
                    const uint32_t w = tp[t];
                    if (w & kTapFirst) {
                        mbar_wait(bar_full + 8 * (ring0 + stage), phase);
                        if (profiling) { const long long t1 = clock64(); c_full += t1 - t0; t0 = t1; }
                        tc_fence_after();
                        a_stage = a_lo0 | ((stage_addr0 >> 4) + (ring0 + stage) * stage16);
                        if ((w & kTapPass) && lane == 0) mbar_arrive(bar_turn + 8 * next_issuer);
                    }
                    const uint32_t a_lo = a_stage + (w & 0x3FFFu);
                    const uint32_t b_lo = b_lo0 + ((w >> 14) & 0x3FFFu);
#pragma unroll
                    for (int ks = 0; ks < kSteps; ++ks) {
                        umma_f16_ss_pred(d_tmem, umma_pack_desc(a_lo + ks * a_kstep, desc_hi),
                                         umma_pack_desc(b_lo + ks * b_kstep, desc_hi), idesc, accumulate, leader);
                        accumulate = 1;
                    }
                    if (w & kTapLast) {
                        umma_commit_pred(bar_empty + 8 * (ring0 + stage), leader);  // frees the smem stage when these MMAs retire
                        if (profiling) { const long long t1 = clock64(); c_issue += t1 - t0; t0 = t1; }
                        advance(1);
                    }
                }
            }
            umma_commit_pred(bar_acc_full + 8 * acc, leader);
            acc_phase ^= 1;
        }
        if (profiling && lane == 0 && iw == 0) {
            atomicAdd(prof + PROF_MMA_WAIT_FULL, static_cast<unsigned long long>(c_full));
            atomicAdd(prof + PROF_MMA_WAIT_ACC, static_cast<unsigned long long>(c_acc));
            atomicAdd(prof + PROF_MMA_ISSUE, static_cast<unsigned long long>(c_issue));
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        const int wp = L.wp, wp2 = L.wp2, hp = L.hp, relu = L.relu, out_mode = L.out_mode;
        const uint32_t wp_magic = L.wp_magic;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_wait = 0, c_work = 0;
        // pixel -> (row, col, inside the zero border?) ; all 32-bit: a chunk has far fewer than 2^26 pixels
        auto locate = [&](int p, int& row, int& col) -> bool {
            row = static_cast<int>(__umulhi(static_cast<uint32_t>(p), wp_magic));
            col = p - row * wp;
            bool inner = col >= 1 && col <= wp - 2;
            if (hp > 0) {
                const int ri = row % hp;
                inner = inner && ri >= 1 && ri <= hp - 2;
            }
            return inner;
        };
        TileWalk tw(blockIdx.x, gridDim.x, n_jobs);
        for (int it = 0; it < my_tiles; ++it) {
            const GemmJob& job = L.jobs[tw.job];
            const int p = tw.mt * kTileM + q * 32 + lane;
            int row, col;
            const bool inner = locate(p, row, col);
            const bool valid = p < M;
            __half* dst;
            bool do_store;
            if (out_mode == OUT_PLAIN) {
                dst = job.out0 + static_cast<long long>(p) * 8;
                do_store = valid;
            } else {
                const int c0 = col - 1;
                dst = ((c0 & 1) ? job.out1 : job.out0) + (static_cast<long long>(row) * wp2 + (c0 >> 1) + 1) * 8;
                do_store = valid && inner;
            }
            const long long out_kc = job.out_kc_stride;
            tw.next();

            long long t0 = profiling ? clock64() : 0;
            mbar_wait(bar_acc_full + 8 * acc, acc_phase);
            if (profiling) { const long long t1 = clock64(); c_wait += t1 - t0; t0 = t1; }
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride;
            uint32_t v[COUT];
            tmem_ld_cols<COUT>(taddr, v);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);  // accumulator is in registers: hand the stage back
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }

            if constexpr (MODE == 0) {
                if (do_store) {
#pragma unroll
                    for (int kc = 0; kc < COUT / 8; ++kc) {
                        const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + kc * 8);
                        const float4 sh1 = *reinterpret_cast<const float4*>(s_shift + kc * 8 + 4);
                        const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
                        uint4 ov;
                        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = kc * 8 + 2 * e;
                            float a = __uint_as_float(v[c]) + sh[2 * e];
                            float b = __uint_as_float(v[c + 1]) + sh[2 * e + 1];
                            if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                            if (!inner) { a = 0.f; b = 0.f; }
                            oh[e] = __floats2half2_rn(a, b);
                        }
                        *reinterpret_cast<uint4*>(dst + kc * out_kc) = ov;
                    }
                }
            } else {
                const bool real = valid && inner;
                if (do_store) {
#pragma unroll
                    for (int kc = 0; kc < COUT / 8; ++kc) {
                        uint4 ov;
                        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = kc * 8 + 2 * e;
                            oh[e] = __floats2bfloat162_rn(real ? __uint_as_float(v[c]) : 0.f, real ? __uint_as_float(v[c + 1]) : 0.f);
                        }
                        *reinterpret_cast<uint4*>(dst + kc * out_kc) = ov;
                    }
                }
                if (L.stats != nullptr) {   // BatchNorm batch statistics from the fp32 accumulators
                    float a[kStatN], b[kStatN];
#pragma unroll
                    for (int c = 0; c < kStatN; ++c) {
                        const float z = (c < COUT && real) ? __uint_as_float(v[c < COUT ? c : 0]) : 0.f;
                        a[c] = z; b[c] = z * z;
                    }
                    warp_reduce_channels<kStatN>(a, lane);
                    warp_reduce_channels<kStatN>(b, lane);
#pragma unroll
                    for (int i = 0; i < kStatN / 32; ++i) {
                        const int c = warp_reduce_channel_of(lane, i, kStatN);
                        atomicAdd(s_shift + c, a[i]);
                        atomicAdd(s_shift + kStatN + c, b[i]);
                    }
                }
            }
            if (profiling) c_work += clock64() - t0;
        }
        if (profiling && q == 0 && lane == 0) {
            atomicAdd(prof + PROF_EPI_WAIT, static_cast<unsigned long long>(c_wait));
            atomicAdd(prof + PROF_EPI_WORK, static_cast<unsigned long long>(c_work));
            atomicAdd(prof + PROF_TILES, static_cast<unsigned long long>(my_tiles));
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kProducers) tmem_dealloc(tmem_base, kTmemCols);
    if (MODE == 1 && L.stats != nullptr)   // (the __syncthreads above ordered the shared-memory atomics)
        for (int i = threadIdx.x; i < 2 * COUT; i += kGemmThreads)
            atomicAdd(L.stats + i, s_shift[(i < COUT ? 0 : kStatN) + (i < COUT ? i : i - COUT)]);
    if (profiling && threadIdx.x == 0) atomicAdd(prof + PROF_CTA, static_cast<unsigned long long>(clock64() - t_cta0));
}

// Host launcher.
template <int CIN, int COUT, int MODE>
static cudaError_t launch_typed(const GemmLaunch& h, int m_tiles, int M, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    const GemmSmem lay = gemm_smem_layout(CIN, COUT, h.n_wtaps, h.ext_alloc, h.groups_per_stage, h.n_stages);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_taps_kernel<CIN, COUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long total = static_cast<long long>(m_tiles) * h.n_jobs;
    if (total <= 0) return cudaSuccess;
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    gemm_taps_kernel<CIN, COUT, MODE><<<grid, kGemmThreads, lay.total, stream>>>(h, m_tiles, M);
    return cudaGetLastError();
}

cudaError_t launch_gemm_taps(const GemmLaunch& h, int m_tiles, int M, int num_sms, cudaStream_t stream) {
#define LD_GEMM_CASE(ci, co) \
    if (h.mode == 0 && h.cin == ci && h.cout == co) return launch_typed<ci, co, 0>(h, m_tiles, M, num_sms, stream)
    LD_GEMM_CASE(64, 64); LD_GEMM_CASE(64, 48); LD_GEMM_CASE(64, 32); LD_GEMM_CASE(64, 16);
    LD_GEMM_CASE(48, 64); LD_GEMM_CASE(48, 48); LD_GEMM_CASE(48, 32); LD_GEMM_CASE(48, 16);
    LD_GEMM_CASE(32, 64); LD_GEMM_CASE(32, 48); LD_GEMM_CASE(32, 32); LD_GEMM_CASE(32, 16);
    LD_GEMM_CASE(16, 64); LD_GEMM_CASE(16, 48); LD_GEMM_CASE(16, 32); LD_GEMM_CASE(16, 16);
#undef LD_GEMM_CASE
    // training (bf16): forward cin -> cout and dgrad cout -> cin of resnet_base
#define LD_GEMM_TRAIN(ci, co) \
    if (h.mode == 1 && h.cin == ci && h.cout == co) return launch_typed<ci, co, 1>(h, m_tiles, M, num_sms, stream)
    LD_GEMM_TRAIN(64, 64); LD_GEMM_TRAIN(64, 32); LD_GEMM_TRAIN(32, 64); LD_GEMM_TRAIN(32, 32);
    LD_GEMM_TRAIN(32, 16); LD_GEMM_TRAIN(16, 32); LD_GEMM_TRAIN(16, 16);
#undef LD_GEMM_TRAIN
    return cudaErrorInvalidValue;
}

// Chooses the smem ring depth for a launch (host side).
int gemm_pick_stages(int cin, int cout, int n_wtaps, int ext_alloc, int groups_per_stage, int max_stages) {
    if (max_stages > kMaxStages || max_stages < 2) max_stages = kMaxStages;
    for (int n = max_stages; n >= 2; --n) {
        if (gemm_smem_layout(cin, cout, n_wtaps, ext_alloc, groups_per_stage, n).total <= 227u * 1024u) return n;
    }
    return 0;
}

}  // namespace ld
