// K2: shifted-plane ("tap list") implicit-GEMM convolution on tcgen05 tensor cores.
//
//   out_o[p, :] = act( sum_t  src_t[p + shift_t, :] @ W[wtap_t]^T  + shift )        (BatchNorm scale folded into W)
//
// for every output pixel p of every output plane o of one conv layer (see ld_plan.cpp).  Replaces the cuDNN/MKLDNN
// conv2d + BatchNorm2d + ReLU (+ residual add) calls made by ResidualBlock.forward / ResNetBigger.forward
// (reference models.py:110-115, 222-228).
//
// Data layout: activations are fp16, channel-chunk planar [C/8][pixels][8].  A run of 8 consecutive pixels of one chunk
// is 128 contiguous bytes = exactly one UMMA "core matrix" of the K-major SWIZZLE_NONE canonical layout, so an A-operand
// tile that starts at ANY pixel offset is a legal smem matrix descriptor: one bulk load of (128 + span) pixels serves all
// taps that differ only by a pixel shift (the three kx taps of a conv row).
//
// Input-stationary N stacking (ld_types.h, GemmJob): a job is a chain of output planes; every input operand is loaded once
// and multiplied, per kx, by the weight rows of ALL outputs it feeds in one MMA (N = 1..3 x cout) that accumulates into
// their adjacent TMEM column ranges.  The epilogue clears an accumulator after reading it, so every MMA accumulates.
//
// Roles (448 threads, one persistent CTA per SM):
//   warps 0..1  producers: cp.async.bulk (UBLKCP) global -> smem ring, mbarrier complete_tx
//   warps 2..5  MMA issuers: tcgen05.mma kind::f16, M=128 K=16, accumulators in TMEM; cout = 64 layers use two of them with
//               256 accumulator columns each (chains of up to 4 outputs), the narrow layers -- whose small MMAs are bound by
//               the issue latency of a single thread -- all four with 128 columns each; tile i belongs to issuer i % n
//   warps 6..13 epilogue: tcgen05.ld -> BN shift, ReLU, pad masking -> fp16 stores, tcgen05.st zeros; warp e reads TMEM lane
//               quadrant e % 4 and the channel half e / 4 (one warp per scheduler is latency-bound on ~300 instructions
//               per 128x64 output; with wide-N MMAs the epilogue would otherwise set the pace)
// The residual add of a ResidualBlock is one more tap: the residual plane times an identity weight slab, accumulated in
// fp32 by the tensor core (exact), so it travels through the same TMA/smem pipeline as every other operand.
// The launch description is a __grid_constant__ parameter, so tile/job/tap bookkeeping runs on the uniform datapath.
#include <cstdio>
#include <cstdlib>

#include <cuda_bf16.h>

#include "ld_ptx.cuh"
#include "ld_types.h"

namespace ld {

constexpr int kIssuers = 4;      // MMA-issuing warps; a launch uses L.n_issuers = 2 or 4 of them: tile i of a CTA is issued by
                                 // warp 2 + i % n_issuers into accumulator stage i % n_issuers (512 / n_issuers columns)
constexpr int kProducers = 2;    // producer warps: with two rings each feeds half of the issuers
constexpr int kEpiWarps = 8;      // epilogue warps: two per TMEM lane quadrant, each takes half of the channels of every output
constexpr int kGemmThreads = 32 * (kProducers + kIssuers + kEpiWarps);
constexpr int kPipeThreads = kGemmThreads + 64;   // layer-pipelined launches: one more warp sends the completion signals, another one
                                                  // follows the completion counters of the neighbouring roles
constexpr int kAccStages = kIssuers;  // barrier slots of the TMEM accumulator ring
constexpr int kMaxStages = 16;
constexpr int kTapBatch = 8;      // taps an issuer warp reads from the job table at a time
static_assert(kMaxTaps % kTapBatch == 0, "the tap table is read in whole batches");

// Optional per-launch cycle counters (GemmLaunch::prof), summed over CTAs:
//   0 producer: waiting for a free smem stage      1 MMA warp: waiting for operands      2 MMA warp: waiting for a free accumulator
//   3 MMA warp: issuing                            4 epilogue warp 0: waiting for MMAs   5 epilogue warp 0: converting + storing
//   6 CTA lifetime                                 7 tiles
//   8 pipelined launches: producer warp 0 waiting for the neighbouring roles (dataflow + back-pressure)
enum { PROF_PROD_WAIT = 0, PROF_MMA_WAIT_FULL, PROF_MMA_WAIT_ACC, PROF_MMA_ISSUE, PROF_EPI_WAIT, PROF_EPI_WORK, PROF_CTA, PROF_TILES, PROF_SYNC_WAIT };

struct GemmSmem {
    uint32_t w_off, jobs_off, stage_off, stage_bytes, param_off, bar_off, total;
};

__host__ __device__ inline GemmSmem gemm_smem_layout(int cin, int cout, int n_wtaps, int n_jobs, int ext_alloc, int groups_per_stage,
                                                     int n_stages) {
    GemmSmem s;
    s.w_off = 0;
    uint32_t w_bytes = static_cast<uint32_t>(n_wtaps) * cin * cout * 2;
    s.jobs_off = (w_bytes + 15u) & ~15u;
    s.stage_off = (s.jobs_off + static_cast<uint32_t>(n_jobs) * static_cast<uint32_t>(sizeof(GemmJob)) + 127u) & ~127u;
    s.stage_bytes = static_cast<uint32_t>(ext_alloc) * 16u * (cin / 8) * groups_per_stage;
    s.param_off = s.stage_off + n_stages * s.stage_bytes;
    // shift[cout] (mode 0) / channel sums [2][..] + BatchNorm coefficients xa, xb, ya, yb [4][..] (mode 1)
    s.bar_off = s.param_off + 6u * static_cast<uint32_t>((cout + 31) / 32 * 32) * sizeof(float);
    s.bar_off = (s.bar_off + 15u) & ~15u;
    s.total = s.bar_off + (2 * kMaxStages + 2 * kAccStages + 1 + kIssuers) * 8 + 48;   // + TMEM slot, arrival counter, watermarks
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
// DUAL: the CTA allocates 256 instead of 512 accumulator columns and half of the shared memory, so that two CTAs share an SM
// (narrow inference layers, whose small MMAs leave the tensor pipe, the copy engine and the epilogue warps idle in turn).
// SYNC: the CTA is one of `n_cta` CTAs of role `role` of a layer-pipelined launch P (ld_types.h, GemmMultiParams): its producer
// warps wait for the m-tiles of the upstream roles a tile reads (and for its consumer not to fall more than lead_max m-tiles
// behind), its epilogue warps signal every finished tile.  Otherwise `cta` / `n_cta` are blockIdx.x / gridDim.x and P is unused.
template <int CIN, int COUT, int MODE, bool DUAL, bool SYNC>
__device__ __forceinline__ void gemm_taps_body(const GemmParams& L, const int m_tiles, const int M, const int cta, const int n_cta,
                                               const GemmMultiParams* P, const int role) {
    constexpr int kTmemCols = DUAL ? 256 : ld::kTmemCols;
    constexpr int kThreads = SYNC ? kPipeThreads : kGemmThreads;
    extern __shared__ __align__(128) uint8_t smem[];
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, and everything the issuing thread derives
    // from it and from the kernel parameters stays on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    constexpr int kChunks = CIN / 8;   // 16-byte channel chunks per pixel
    constexpr int kSteps = CIN / 16;   // MMAs (K = 16) per tap
    const long long t_cta0 = clock64();
    // PDL: the next conv launch may be scheduled onto SMs as the CTAs of this grid exit; its prologue (barriers, TMEM, weights)
    // then overlaps this grid's tail.  It reads/writes planes only after griddep_wait() below.
    if (threadIdx.x == 0) griddep_launch_dependents();

    const int n_wtaps = L.n_wtaps, ext_alloc = L.ext_alloc, n_stages = L.n_stages, gps = L.groups_per_stage;
    const GemmSmem lay = gemm_smem_layout(CIN, COUT, n_wtaps, L.n_jobs, ext_alloc, gps, n_stages);
    const GemmJob* s_jobs = reinterpret_cast<const GemmJob*>(smem + lay.jobs_off);

    float* s_shift = reinterpret_cast<float*>(smem + lay.param_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kAccStages + 1 + kIssuers);
    uint32_t* sig_cnt = tmem_slot + 1;   // pipelined launches: epilogue-warp arrivals (8 per finished tile), read by the signalling warp
    uint32_t* prod_done = tmem_slot + 2;   // ... producer warps that have issued their last load
    int32_t* s_wm = reinterpret_cast<int32_t*>(tmem_slot + 4);   // ... [4] watermarks kept by the polling warp: every m-tile up to s_wm[k] of the
                                                                 // role 1, 2, 3 places upstream (k = 0..2) / of the consumer (k = 3) is complete

    const uint32_t bar_full = smem_u32(bars);                       // [n_stages]
    const uint32_t bar_empty = smem_u32(bars + kMaxStages);         // [n_stages]
    const uint32_t bar_acc_full = smem_u32(bars + 2 * kMaxStages);  // [kAccStages]
    const uint32_t bar_acc_empty = bar_acc_full + 8 * kAccStages;   // [kAccStages]
    const uint32_t bar_w = bar_acc_empty + 8 * kAccStages;
    const uint32_t bar_turn = bar_w + 8;                            // [kIssuers] issuer i may start waiting for operands

    // MODE 0: s_shift[COUT] = folded shift.  MODE 1: s_shift[2 * max(COUT, 32)] = per-CTA channel sums, flushed at the end.
    constexpr int kStatAll = (COUT + 31) / 32 * 32;     // smem slots per statistic
    constexpr int kStatN = (COUT / 2 + 31) / 32 * 32;   // channels one epilogue warp reduces at a time (padded to the butterfly width)
    if constexpr (MODE == 0) {
        for (int i = threadIdx.x; i < COUT; i += kThreads) s_shift[i] = L.shift[i];
    } else {
        for (int i = threadIdx.x; i < 2 * kStatAll; i += kThreads) s_shift[i] = 0.f;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(bar_acc_full + 8 * i, 1);
            mbar_init(bar_acc_empty + 8 * i, kEpiWarps);
        }
        mbar_init(bar_w, 1);
        *sig_cnt = 0u;
        *prod_done = 0u;
        s_wm[0] = s_wm[1] = s_wm[2] = s_wm[3] = -1;
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
    if (warp >= kProducers + kIssuers && warp < kProducers + kIssuers + kEpiWarps) {   // accumulators start at zero: every MMA accumulates (see the epilogue)
        tmem_zero_cols<kTmemCols>(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16));
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int n_jobs = L.n_jobs;
    const int total_tiles = m_tiles * n_jobs;
    const int my_tiles = (total_tiles - cta + n_cta - 1) / n_cta;
    const uint32_t w_addr = smem_u32(smem + lay.w_off);
    const uint32_t stage_addr0 = smem_u32(smem + lay.stage_off);
    const uint32_t box_bytes = static_cast<uint32_t>(ext_alloc) * 16u * kChunks;  // one group's operand block
    unsigned long long* prof = L.prof;
    const bool profiling = prof != nullptr;

    if (warp != 0) griddep_wait();   // (warp 0 first issues the loads of the launch constants -- weights, job table -- below)
    if (warp < kProducers) {
        // ------------------------------------------------------------------ producers
        if (warp == 0) {
            // weights -> smem.  Stacked 3x3 layout: [kx][cin/8][ky = 2, 1, 0][cout][8], so that the rows of the outputs an
            // input row feeds (ky = 2 for the row above, 1 for its own, 0 for the row below) are adjacent along N
            // (stride-2 convs: ky = 2, 0, 1 -- an odd input row feeds ky = 2 of one output and ky = 0 of the next).
            constexpr uint32_t tap_bytes = static_cast<uint32_t>(CIN) * COUT * 2, row_bytes = COUT * 16u;
            const uint32_t jobs_bytes = static_cast<uint32_t>(n_jobs) * static_cast<uint32_t>(sizeof(GemmJob));
            if (lane == 0) {
                mbar_expect_tx(bar_w, tap_bytes * n_wtaps + jobs_bytes);
                bulk_g2s(smem_u32(smem + lay.jobs_off), L.jobs_dev, jobs_bytes, bar_w);
            }
            __syncwarp();
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(L.weights);
            if (L.w_stack) {
                const int n_stacked = 9 * L.w_blocks;   // split precision: block 0 = hi weights, block 1 = lo weights
                for (int c = lane; c < n_stacked * kChunks; c += 32) {
                    const int slab = c / kChunks, kc = c - slab * kChunks, blk = slab / 9, r = slab - blk * 9, ky = r / 3, kx = r - ky * 3;
                    bulk_g2s(w_addr + blk * 9 * tap_bytes + ((kx * kChunks + kc) * 3 + w_stack_row(L.w_stack, ky)) * row_bytes,
                             wsrc + static_cast<size_t>(c) * row_bytes, row_bytes, bar_w);
                }
                for (int t = n_stacked + lane; t < n_wtaps; t += 32)
                    bulk_g2s(w_addr + t * tap_bytes, wsrc + static_cast<size_t>(t) * tap_bytes, tap_bytes, bar_w);
            } else {
                for (int t = lane; t < n_wtaps; t += 32)
                    bulk_g2s(w_addr + t * tap_bytes, wsrc + static_cast<size_t>(t) * tap_bytes, tap_bytes, bar_w);
            }
        }
        if (warp == 0) griddep_wait();   // the previous kernel's planes are complete and visible from here on
        mbar_wait(bar_w, 0);   // the job table (and the weights) are in shared memory
        const uint32_t lbo_a = static_cast<uint32_t>(ext_alloc) * 16u;  // bytes between channel chunks of a group
        const uint32_t copy_a = static_cast<uint32_t>(L.ext_copy > 0 ? L.ext_copy : ext_alloc) * 16u;   // bytes one bulk copy transfers
        // Two independent pipelines: producer w fills ring w (stages [w * ring_n, (w + 1) * ring_n)) with the tiles
        // it = w, w + 2, ... of the CTA; issuer w drains it.  A ring is filled and drained in tile order, stage by stage,
        // so nobody ever waits more than one phase ahead on its parity-tracked barriers.
        // (n_rings == 1 when the smem ring is too short to split: producer 0 feeds both issuers, who take turns)
        const int n_rings = L.n_rings;
        const int ring_n = n_stages / n_rings;
        const int ring0 = warp * ring_n;
        int stage = 0;       // position inside this producer's ring
        uint32_t phase = 0;
        long long c_wait = 0, c_sync = 0, c_sync_up = 0;
        TileWalk tw(cta + warp * n_cta, n_rings * n_cta, n_jobs);
        // L2 prefetch: the ring holds about half a tile of operands, less than the DRAM latency under load covers; the operands
        // of this producer's tile `l2pf` steps ahead are requested into L2 now, so that their ring loads find them there
        const int l2pf = L.l2_prefetch;
        TileWalk tw_pf = tw;
        for (int i = 0; i < l2pf; ++i) tw_pf.next();
        for (int it = warp; it < my_tiles && warp < n_rings; it += n_rings, tw.next()) {
            const GemmJob& job = s_jobs[tw.job];
            const int p0 = tw.mt * kTileM;
            const int n_groups = job.n_groups;
            if constexpr (SYNC) {
                // dataflow: the polling warp keeps the neighbours' completion watermarks in shared memory; a tile's loads start once
                // the upstream m-tiles it reads are complete and the consumer is at most lead_max m-tiles behind
                const long long t0 = profiling ? clock64() : 0;
                bool waited = false;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int dep = job.dep_back[k];
                    if (dep != kNoDep && role - 1 - k >= 0 && p0 + dep >= 0 && !(P->sync.dbg & 1)) {
                        const int need = min(m_tiles - 1, (p0 + dep) / kTileM);
                        while (static_cast<int>(ld_acquire_cta_shared(smem_u32(s_wm + k))) < need) { __nanosleep(20); }
                        waited = true;
                    }
                }
                if (waited) fence_proxy_async();   // the acquired (generic-proxy) writes are ordered before this thread's bulk-copy reads
                const long long t1 = profiling ? clock64() : 0;
                if (role + 1 < P->n_roles && tw.mt - P->sync.lead_max >= 0 && !(P->sync.dbg & 2))
                    while (static_cast<int>(ld_acquire_cta_shared(smem_u32(s_wm + 3))) < tw.mt - P->sync.lead_max) { __nanosleep(20); }
                if (profiling) { const long long t2 = clock64(); c_sync += t2 - t0; c_sync_up += t1 - t0; }
            }
            if (l2pf > 0) {
                if (it + l2pf * n_rings < my_tiles) {
                    const GemmJob& pj = s_jobs[tw_pf.job];
                    const int pp0 = tw_pf.mt * kTileM;
                    for (int c = lane; c < pj.n_groups * kChunks; c += 32) {
                        const int g = c / kChunks, kc = c - g * kChunks;
                        const GemmGroup& grp = pj.groups[g];
                        bulk_prefetch_l2(grp.src + static_cast<long long>(pp0 + grp.shift) * 8 + kc * grp.kc_stride, lbo_a);
                    }
                }
                tw_pf.next();
            }
            for (int g0 = 0; g0 < n_groups; g0 += gps) {
                const int ng = min(gps, n_groups - g0);
                const int n_copies = (L.dbg & 1) ? 1 : ng * kChunks;   // (dbg bit 0: timing experiment, one copy per stage)
                const uint32_t full = bar_full + 8 * (ring0 + stage);
                if (lane == 0) {
                    const long long t0 = profiling ? clock64() : 0;
                    mbar_wait(bar_empty + 8 * (ring0 + stage), phase ^ 1);
                    if (profiling) c_wait += clock64() - t0;
                    mbar_expect_tx(full, copy_a * n_copies);
                }
                __syncwarp();
                const uint32_t dst0 = stage_addr0 + (ring0 + stage) * lay.stage_bytes;
                for (int c = lane; c < n_copies; c += 32) {   // one contiguous bulk copy per channel chunk of a group
                    const int g = c / kChunks, kc = c - g * kChunks;
                    const GemmGroup& grp = job.groups[g0 + g];
                    bulk_g2s(dst0 + g * box_bytes + kc * lbo_a, grp.src + static_cast<long long>(p0 + grp.shift) * 8 + kc * grp.kc_stride,
                             copy_a, full);
                }
                if (++stage == ring_n) { stage = 0; phase ^= 1; }
            }
        }
        if constexpr (SYNC) {
            __syncwarp();
            if (lane == 0) atomicAdd(prod_done, 1u);
        }
        if (profiling && lane == 0 && warp == 0) {
            atomicAdd(prof + PROF_PROD_WAIT, static_cast<unsigned long long>(c_wait));
            if (SYNC) {
                atomicAdd(prof + PROF_SYNC_WAIT, static_cast<unsigned long long>(c_sync));
                atomicAdd(prof + PROF_SYNC_WAIT + 1, static_cast<unsigned long long>(c_sync_up));
            }
        }
    } else if (warp < kProducers + kIssuers) {
        // ------------------------------------------------------------------ MMA issuers (uniform control flow per warp)
        const int iw = warp - kProducers;  // this warp issues tiles iw, iw + n_issuers, ... of the CTA into accumulator stage iw
        const int n_issuers = L.n_issuers;
        // kind::f16, fp32 D, K-major A/B, M = 128; N comes with every tap
        constexpr uint32_t idesc0 = umma_idesc_f16(0u) | (MODE == 1 ? ((1u << 7) | (1u << 10)) : 0u);  // bf16 A/B in training
        // smem matrix descriptors (see ld_ptx.cuh): only the low word (start address, LBO) changes per MMA.
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);            // SBO = 128 B, descriptor version 1
        const uint32_t a_lo0 = (static_cast<uint32_t>(ext_alloc) << 16);  // LBO = ext_alloc * 16 B
        const uint32_t a_kstep = 2u * static_cast<uint32_t>(ext_alloc);   // two channel chunks per K = 16
        const uint32_t w16 = w_addr >> 4;
        const uint32_t stage16 = lay.stage_bytes >> 4;
        mbar_wait(bar_w, 0);
        // issuer iw drains ring iw % n_rings together with the ring's other issuers: they alternate its tiles, the partners'
        // tiles are skipped by advancing the ring position by their stage count
        const int n_rings = L.n_rings;
        const int ipr = n_issuers / n_rings;     // issuers per ring
        const int ring = iw % n_rings, slot = iw / n_rings;
        const int next_issuer = ring + n_rings * ((slot + 1) % ipr);
        const int ring_n = n_stages / n_rings;
        const int ring0 = ring * ring_n;
        // The whole warp runs the loop converged -- every value below is warp-uniform, so the compiler keeps the tap words,
        // the descriptor arithmetic and the MMA operands in uniform registers -- and only the elected lane issues.
        const bool leader = elect_one();
        const bool mma_on = leader && !(L.dbg & 2);   // (dbg bit 1: timing experiment, everything but the MMAs themselves)
        long long c_full = 0, c_acc = 0, c_issue = 0;
        if (iw < n_issuers) {
            int stage = 0;
            uint32_t phase = 0;
            const int acc = iw;
            uint32_t acc_phase = 0;
            // "full" barriers are parity-tracked, so a warp must never wait on a stage more than one fill ahead: the issuers
            // of a ring take turns -- one starts waiting for operands only after the other has seen its last stage
            uint32_t turn_phase = 0;
            const uint32_t d_tmem = tmem_base + acc * (kTmemCols / n_issuers);
            TileWalk tw(cta + ring * n_cta, n_rings * n_cta, n_jobs);
            for (int it = ring, k = 0; it < my_tiles; it += n_rings, ++k, tw.next()) {
                const GemmJobTaps jt = L.job_taps[tw.job];
                if ((k % ipr) != slot) {
                    stage += jt.n_stages;
                    while (stage >= ring_n) { stage -= ring_n; phase ^= 1; }
                    continue;
                }
                long long t0c = profiling ? clock64() : 0;
                if (k > 0) {
                    mbar_wait(bar_turn + 8 * iw, turn_phase);
                    turn_phase ^= 1;
                }
                mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
                if (profiling) { const long long t1 = clock64(); c_acc += t1 - t0c; t0c = t1; }
                tc_fence_after();
                uint32_t a_stage = 0;
                const int t_end = jt.tap0 + jt.n_taps;
                for (int t = jt.tap0; t < t_end; ++t) {
                    const uint4 w = L.taps[t];
                    if (w.x & kTapFirst) {
                        mbar_wait(bar_full + 8 * (ring0 + stage), phase);
                        if (profiling) { const long long t1 = clock64(); c_full += t1 - t0c; t0c = t1; }
                        tc_fence_after();
                        a_stage = a_lo0 | ((stage_addr0 >> 4) + (ring0 + stage) * stage16);
                        if ((w.x & kTapPass) && lane == 0) mbar_arrive(bar_turn + 8 * next_issuer);
                    }
                    const uint32_t a_lo = a_stage + (w.x & 0x3FFFu);
                    const uint32_t b_lo = w.y + w16;
                    const uint32_t b_kstep = w.y >> 15;   // two channel chunks per K = 16: 2 * LBO
                    const uint32_t d = d_tmem + w.z;
                    const uint32_t idesc = idesc0 | w.w;
                    const int nks = (w.x & kTapHalfK) ? kSteps / 2 : kSteps;   // split precision: hi half of [hi | lo] x lo weights
#pragma unroll
                    for (int ks = 0; ks < kSteps; ++ks)
                        umma_f16_ss_pred(d, umma_pack_desc(a_lo + ks * a_kstep, desc_hi), umma_pack_desc(b_lo + ks * b_kstep, desc_hi), idesc, 1u,
                                         mma_on && ks < nks);
                    if (w.x & kTapLast) {
                        umma_commit_pred(bar_empty + 8 * (ring0 + stage), leader);  // frees the smem stage when these MMAs retire
                        if (profiling) { const long long t1 = clock64(); c_issue += t1 - t0c; t0c = t1; }
                        if (++stage == ring_n) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit_pred(bar_acc_full + 8 * acc, leader);
                acc_phase ^= 1;
            }
            if (profiling && iw == 0 && lane == 0) {
                atomicAdd(prof + PROF_MMA_WAIT_FULL, static_cast<unsigned long long>(c_full));
                atomicAdd(prof + PROF_MMA_WAIT_ACC, static_cast<unsigned long long>(c_acc));
                atomicAdd(prof + PROF_MMA_ISSUE, static_cast<unsigned long long>(c_issue));
            }
        }
        __syncwarp();
    } else if (warp < kProducers + kIssuers + kEpiWarps) {
        // ------------------------------------------------------------------ epilogue
        const int q = warp & 3;  // TMEM lane quadrant this warp may read
        const int half = (warp - kProducers - kIssuers) >> 2;   // which half of the channels of every output
        constexpr int CH = COUT / 2;                             // accumulator columns per epilogue warp
        static_assert(CH % 8 == 0, "cout must be a multiple of 16");
        const int wp = L.wp, wp2 = L.wp2, hp = L.hp, relu = L.relu, out_mode = L.out_mode;
        const int w_real = L.w_real > 0 ? L.w_real : wp - 2;   // (0: the dense training layout with a pad column on either side)
        const bool split_out = MODE == 0 && L.split_out != 0;
        const uint32_t wp_magic = L.wp_magic;
        const int n_acc = L.n_issuers, acc_cols = kTmemCols / n_acc;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_wait = 0, c_work = 0;
        // pixel -> (row, col, inside the zero border?) ; all 32-bit: a chunk has far fewer than 2^26 pixels
        auto locate = [&](int p, int& row, int& col) -> bool {
            row = static_cast<int>(__umulhi(static_cast<uint32_t>(p), wp_magic));
            col = p - row * wp;
            bool inner = col >= 1 && col <= w_real;
            if (hp > 0) {
                const int ri = row % hp;
                inner = inner && ri >= 1 && ri <= hp - (wp - w_real);   // as many pad rows per image as pad columns per row
            }
            return inner;
        };
        mbar_wait(bar_w, 0);   // the job table is in shared memory
        const bool bwd_stats = MODE == 1 && L.stats != nullptr && L.stats_kind == 1;
        float* s_coef = s_shift + 2 * kStatAll;   // xa | xb | ya | yb, kStatAll floats each (mode 1, backward statistics)
        if constexpr (MODE == 1) {
            if (bwd_stats && lane < CH) {
                // the same expressions as bn_coef_setup (ld_train.cu): the ReLU mask recomputed from z must match the forward pass.
                // Every warp of a channel half writes identical values (after griddep_wait: the sums come from earlier kernels).
                const int c = half * CH + lane;
                const float m = L.bwd.fwd_sums[c] * L.bwd.inv_n;
                const float var = fmaxf(L.bwd.fwd_sums[COUT + c] * L.bwd.inv_n - m * m, 0.f);
                const float inv = rsqrtf(var + 1e-5f);
                const float gam = L.bwd.gamma[c];
                s_coef[c] = inv;
                s_coef[kStatAll + c] = -m * inv;
                s_coef[2 * kStatAll + c] = gam * inv;
                s_coef[3 * kStatAll + c] = fmaf(-m * inv, gam, L.bwd.beta[c]);
            }
            __syncwarp();
        }
        TileWalk tw(cta, n_cta, n_jobs);
        // Completion of a tile (pipelined launches): this warp's share of it has been stored -- one of kEpiWarpsPerTile arrivals on
        // the CTA's shared-memory counter.  The signalling warp turns eight of them into the device-wide signal: a gpu-scope
        // fence right behind the stores would stall every epilogue warp for a trip through the saturated memory system per tile.
        auto signal_done = [&]() {
            if constexpr (SYNC) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    atomicAdd(sig_cnt, 1u);
                }
            }
        };
        if constexpr (SYNC) {   // the other counter buffer belongs to the previous pipelined launch, complete by now: clear it for the next
            unsigned* zn = P->sync.done_next;
            const int n_cnt = P->n_roles * P->sync.m_cap;
            for (int i = static_cast<int>(blockIdx.x) * (kEpiWarps * 32) + (static_cast<int>(threadIdx.x) - 32 * (kProducers + kIssuers)); i < n_cnt;
                 i += static_cast<int>(gridDim.x) * (kEpiWarps * 32))
                zn[i] = 0u;
        }
        for (int it = 0; it < my_tiles; ++it) {
            const GemmJob& job = s_jobs[tw.job];
            const int p = tw.mt * kTileM + q * 32 + lane;
            int row, col;
            const bool inner = locate(p, row, col);
            const bool valid = p < M;
            long long dst_off;   // element offset of this pixel inside an output plane
            bool odd = false, do_store;   // (dbg bit 2: timing experiment without the epilogue's global stores)
            if (out_mode == OUT_PLAIN) {
                dst_off = static_cast<long long>(p) * 8;
                do_store = valid;
            } else {
                const int c0 = col - 1;
                odd = c0 & 1;
                dst_off = (static_cast<long long>(row) * wp2 + (c0 >> 1) + 1) * 8;
                do_store = valid && inner;
            }
            if (L.dbg & 4) do_store = false;
            const long long out_kc = job.out_kc_stride;
            const int n_outs = job.n_outs;
            tw.next();

            // backward statistics: this pixel's z (and y) values are fetched while the MMAs of the tile still run
            uint4 zr[CH / 8], yr[CH / 8];
            if constexpr (MODE == 1) {
                if (bwd_stats) {
                    const bool real_px = valid && inner;
#pragma unroll
                    for (int kc = 0; kc < CH / 8; ++kc) {
                        zr[kc] = yr[kc] = make_uint4(0u, 0u, 0u, 0u);
                        if (real_px) {
                            zr[kc] = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(L.bwd.z) +
                                                                     (half * (CH / 8) + kc) * L.bwd.z_kc + static_cast<long long>(p) * 8);
                            if (L.bwd.mask_mode == 1)
                                yr[kc] = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(L.bwd.y) +
                                                                         (half * (CH / 8) + kc) * L.bwd.y_kc + static_cast<long long>(p) * 8);
                        }
                    }
                }
            }
            long long t0 = profiling ? clock64() : 0;
            mbar_wait(bar_acc_full + 8 * acc, acc_phase);
            if (profiling) { const long long t1 = clock64(); c_wait += t1 - t0; t0 = t1; }
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * acc_cols;
            for (int o = 0; o < n_outs; ++o) {
                const uint32_t taddr = taddr0 + o * COUT + half * CH;
                uint32_t v[CH];
                if (!(L.dbg & 8)) {   // (dbg bit 3: timing experiment without the TMEM traffic of the epilogue)
                    tmem_ld_cols<CH>(taddr, v);
                    tmem_wait_ld();
                    tmem_zero_cols<CH>(taddr);   // the next tile's MMAs accumulate into zeros
                } else {
#pragma unroll
                    for (int c = 0; c < CH; ++c) v[c] = 0u;
                }
                if (o == n_outs - 1) {
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);  // all accumulators are in registers: hand the stage back
                }
                __half* dst = (odd ? job.outs[o].out1 : job.outs[o].out0) + dst_off + (half * (CH / 8)) * out_kc;

                if constexpr (MODE == 0) {
                    if (do_store) {
#pragma unroll
                        for (int kc = 0; kc < CH / 8; ++kc) {
                            const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + half * CH + kc * 8);
                            const float4 sh1 = *reinterpret_cast<const float4*>(s_shift + half * CH + kc * 8 + 4);
                            const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
                            uint4 ov;
                            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = kc * 8 + 2 * e;
                                float a = __uint_as_float(v[c]) + sh[2 * e];
                                float b = __uint_as_float(v[c + 1]) + sh[2 * e + 1];
                                if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                                oh[e] = __floats2half2_rn(a, b);
                            }
                            if (!inner) ov = make_uint4(0u, 0u, 0u, 0u);
                            *reinterpret_cast<uint4*>(dst + kc * out_kc) = ov;
                            if (split_out) {   // the fp16 rounding residual goes to channel chunk cout/8 + kc of the [hi | lo] plane
                                uint4 lv;
                                __half2* lh = reinterpret_cast<__half2*>(&lv);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int c = kc * 8 + 2 * e;
                                    float a = __uint_as_float(v[c]) + sh[2 * e];
                                    float b = __uint_as_float(v[c + 1]) + sh[2 * e + 1];
                                    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                                    const float2 hi = __half22float2(oh[e]);
                                    lh[e] = __floats2half2_rn(a - hi.x, b - hi.y);
                                }
                                if (!inner) lv = make_uint4(0u, 0u, 0u, 0u);
                                *reinterpret_cast<uint4*>(dst + (COUT / 8 + kc) * out_kc) = lv;
                            }
                        }
                    }
                } else {
                    const bool real = valid && inner;
                    if (do_store) {
#pragma unroll
                        for (int kc = 0; kc < CH / 8; ++kc) {
                            uint4 ov;
                            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int c = kc * 8 + 2 * e;
                                oh[e] = __floats2bfloat162_rn(__uint_as_float(v[c]), __uint_as_float(v[c + 1]));
                            }
                            if (!real) ov = make_uint4(0u, 0u, 0u, 0u);
                            *reinterpret_cast<uint4*>(dst + kc * out_kc) = ov;
                        }
                    }
                    if (L.stats != nullptr) {   // BatchNorm batch statistics from the fp32 accumulators
                        float a[kStatN], b[kStatN];
                        if (bwd_stats) {   // backward: sum g, sum g * xhat with g = dy * [y > 0]
                            const int mask_mode = L.bwd.mask_mode;
#pragma unroll
                            for (int c = 0; c < kStatN; ++c) {
                                float g = 0.f, gx = 0.f;
                                if (c < CH && real) {
                                    const __nv_bfloat16* zh = reinterpret_cast<const __nv_bfloat16*>(&zr[(c < CH ? c : 0) / 8]);
                                    const __nv_bfloat16* yh = reinterpret_cast<const __nv_bfloat16*>(&yr[(c < CH ? c : 0) / 8]);
                                    const float zf = __bfloat162float(zh[c % 8]);
                                    const int cc = half * CH + c;
                                    bool on = true;
                                    if (mask_mode == 1) on = __bfloat162float(yh[c % 8]) > 0.f;
                                    if (mask_mode == 2) on = fmaf(zf, s_coef[2 * kStatAll + cc], s_coef[3 * kStatAll + cc]) > 0.f;
                                    g = on ? __uint_as_float(v[c < CH ? c : 0]) : 0.f;
                                    gx = g * fmaf(zf, s_coef[cc], s_coef[kStatAll + cc]);
                                }
                                a[c] = g; b[c] = gx;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < kStatN; ++c) {
                                const float z = (c < CH && real) ? __uint_as_float(v[c < CH ? c : 0]) : 0.f;
                                a[c] = z; b[c] = z * z;
                            }
                        }
                        warp_reduce_channels<kStatN>(a, lane);
                        warp_reduce_channels<kStatN>(b, lane);
#pragma unroll
                        for (int i = 0; i < kStatN / 32; ++i) {
                            const int c = warp_reduce_channel_of(lane, i, kStatN);
                            if (c < CH) {
                                atomicAdd(s_shift + half * CH + c, a[i]);
                                atomicAdd(s_shift + kStatAll + half * CH + c, b[i]);
                            }
                        }
                    }
                }
            }
            if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
            if constexpr (SYNC) signal_done();
            if (profiling) c_work += clock64() - t0;
        }
        if (profiling && q == 0 && half == 0 && lane == 0) {
            atomicAdd(prof + PROF_EPI_WAIT, static_cast<unsigned long long>(c_wait));
            atomicAdd(prof + PROF_EPI_WORK, static_cast<unsigned long long>(c_work));
            atomicAdd(prof + PROF_TILES, static_cast<unsigned long long>(my_tiles));
        }
    }

    else if (warp == kProducers + kIssuers + kEpiWarps + 1) {
        // ------------------------------------------------------------------ polling warp (pipelined launches only)
        // Follows the completion counters of up to three upstream roles and of the consumer, 32 m-tiles per read, and publishes
        // "every m-tile up to here is complete" in shared memory: the producer warps then test a tile's dependencies with one
        // shared-memory read instead of a round trip to L2 per tile and neighbour.
        if constexpr (SYNC) {
            const int m_cap = P->sync.m_cap;
            int mark[4] = {-1, -1, -1, -1};
            while (ld_acquire_cta_shared(smem_u32(prod_done)) < static_cast<uint32_t>(kProducers)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int other = k < 3 ? role - 1 - k : role + 1;
                    if (other < 0 || other >= P->n_roles || mark[k] >= m_tiles - 1) continue;
                    const int idx = mark[k] + 1 + lane;
                    const bool done_l = idx < m_tiles && ld_acquire_gpu(P->sync.done + other * m_cap + idx) >= P->expect[other];
                    const unsigned ok = __ballot_sync(0xffffffffu, done_l);
                    const int n = ok == 0xffffffffu ? 32 : __ffs(~ok) - 1;
                    if (n > 0) {
                        mark[k] += n;
                        if (lane == 0) st_release_cta_shared(smem_u32(s_wm + k), static_cast<uint32_t>(mark[k]));
                    }
                }
                __nanosleep(200);
            }
        }
    } else {
        // ------------------------------------------------------------------ signalling warp (pipelined launches only)
        if constexpr (SYNC) {
            if (lane == 0) {
                TileWalk tw(cta, n_cta, n_jobs);
                unsigned* done = P->sync.done + role * P->sync.m_cap;
                for (int it = 0; it < my_tiles;) {
                    unsigned c;
                    while ((c = ld_acquire_cta_shared(smem_u32(sig_cnt))) < static_cast<uint32_t>(kEpiWarps * (it + 1))) __nanosleep(32);
                    // the eight warps' stores of every tile counted so far were ordered before their arrivals (cta scope); ONE fence
                    // makes them visible device-wide before the m-tiles' counters move (it costs about a tile's time under load:
                    // whatever finished meanwhile is signalled with the next one)
                    const int ready = min(my_tiles, static_cast<int>(c / kEpiWarps));
                    if (!(P->sync.dbg & 4)) __threadfence();
                    for (; it < ready; ++it, tw.next()) atomicAdd(done + tw.mt, static_cast<unsigned>(kEpiWarps));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kProducers) tmem_dealloc(tmem_base, kTmemCols);
    if (MODE == 1 && L.stats != nullptr)   // (the __syncthreads above ordered the shared-memory atomics)
        for (int i = threadIdx.x; i < 2 * COUT; i += kThreads)
            atomicAdd(L.stats + i, s_shift[(i < COUT ? 0 : kStatAll) + (i < COUT ? i : i - COUT)]);
    if (profiling && threadIdx.x == 0) {
        const unsigned long long life = static_cast<unsigned long long>(clock64() - t_cta0);
        atomicAdd(prof + PROF_CTA, life);
        // shortest / longest lifetime per tile of a CTA (x 1024): the spread between the CTAs of a launch
        const unsigned long long per_tile = my_tiles > 0 ? (life << 10) / static_cast<unsigned long long>(my_tiles) : 0ull;
        if (my_tiles > 0) {
            atomicMax(prof + PROF_SYNC_WAIT + 2, per_tile);
            atomicMax(prof + PROF_SYNC_WAIT + 3, ~0ull - per_tile);   // (the minimum, stored inverted: the buffer is cleared with zeros)
        }
    }
}

template <int CIN, int COUT, int MODE, bool DUAL = false>
__global__ void __launch_bounds__(kGemmThreads, DUAL ? 2 : 1)
gemm_taps_kernel(const __grid_constant__ GemmParams L, int m_tiles, int M) {
    gemm_taps_body<CIN, COUT, MODE, DUAL, false>(L, m_tiles, M, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), nullptr, 0);
}

// Layer-pipelined launch (inference): the CTA looks its role up by block index and runs that layer's program.
template <int CIN, int COUT, bool DUAL>
__global__ void __launch_bounds__(kPipeThreads, DUAL ? 2 : 1)
gemm_taps_pipe_kernel(const __grid_constant__ GemmMultiParams P, int m_tiles, int M) {
    int role = 0;
    while (role + 1 < P.n_roles && static_cast<int>(blockIdx.x) >= P.cta0[role + 1]) ++role;
    gemm_taps_body<CIN, COUT, 0, DUAL, true>(P.role[role], m_tiles, M, static_cast<int>(blockIdx.x) - P.cta0[role], P.cta0[role + 1] - P.cta0[role],
                                             &P, role);
}

// Host launcher.
template <int CIN, int COUT, int MODE, bool DUAL = false>
static cudaError_t launch_typed(const GemmLaunch& h, int m_tiles, int M, int num_sms, cudaStream_t stream) {
    if (DUAL) num_sms *= 2;   // two resident CTAs per SM
    static PerDeviceOnce attr_set;
    const GemmSmem lay = gemm_smem_layout(CIN, COUT, h.n_wtaps, h.n_jobs, h.ext_alloc, h.groups_per_stage, h.n_stages);
    if (!attr_set.flag()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_taps_kernel<CIN, COUT, MODE, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(gemm_smem_cap(DUAL ? 256 : ld::kTmemCols)));
        if (e != cudaSuccess) return e;
        attr_set.flag() = true;
    }
    const long long total = static_cast<long long>(m_tiles) * h.n_jobs;
    if (total <= 0) return cudaSuccess;
    int grid = static_cast<int>(total < num_sms ? total : num_sms);
    // tile t is job t % n_jobs and CTA b owns tiles b, b + grid, ...: keep grid and n_jobs coprime so that every CTA sees
    // every job kind (jobs differ in cost: chains of several outputs, the interior plane)
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    while (total > grid && grid > 1 && gcd(grid, h.n_jobs) != 1) --grid;
    static const bool pdl = []() { const char* v = std::getenv("LD_GEMM_PDL"); return v == nullptr || std::atoi(v) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = lay.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, gemm_taps_kernel<CIN, COUT, MODE, DUAL>, static_cast<const GemmParams&>(h), m_tiles, M);
}

cudaError_t launch_gemm_taps(GemmLaunch& h, int m_tiles, int M, int num_sms, cudaStream_t stream) {
    if (h.jobs_dev == nullptr) {   // first use: the job table moves to device memory (synchronous, once per launch description)
        GemmJob* dev = nullptr;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dev), static_cast<size_t>(h.n_jobs) * sizeof(GemmJob));
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(dev, h.jobs, static_cast<size_t>(h.n_jobs) * sizeof(GemmJob), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(dev); return e; }
        h.jobs_dev = dev;
    }
#define LD_GEMM_CASE(ci, co)                                                                                          \
    if (h.mode == 0 && h.cin == ci && h.cout == co) {                                                                 \
        if (co <= 32 && h.tmem_cols == 256) return launch_typed<ci, co, 0, (co <= 32)>(h, m_tiles, M, num_sms, stream); \
        return launch_typed<ci, co, 0>(h, m_tiles, M, num_sms, stream);                                               \
    }
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

// Layer-pipelined launch: `roles[r]` are built launches of one shape (cin, cout, inference, same accumulator budget), `ctas[r]`
// the CTAs each gets (their sum is the grid).
template <int CIN, int COUT, bool DUAL>
static cudaError_t launch_pipe_typed(const GemmMultiParams& P, int grid, int m_tiles, int M, unsigned smem_bytes, cudaStream_t stream) {
    static PerDeviceOnce attr_set;
    if (!attr_set.flag()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_taps_pipe_kernel<CIN, COUT, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(gemm_smem_cap(DUAL ? 256 : ld::kTmemCols)));
        if (e != cudaSuccess) return e;
        attr_set.flag() = true;
    }
    static const bool pdl = []() { const char* v = std::getenv("LD_GEMM_PDL"); return v == nullptr || std::atoi(v) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kPipeThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, gemm_taps_pipe_kernel<CIN, COUT, DUAL>, P, m_tiles, M);
}

cudaError_t launch_gemm_pipe(GemmLaunch* const* roles, int n_roles, const int* ctas, const GemmSync& sync, int m_tiles, int M,
                             cudaStream_t stream) {
    if (n_roles < 1 || n_roles > kMaxRoles || m_tiles > sync.m_cap) return cudaErrorInvalidValue;
    static thread_local GemmMultiParams P;   // 26 KB: assembled per launch, passed by value
    unsigned smem = 0;
    int grid = 0;
    for (int r = 0; r < n_roles; ++r) {
        GemmLaunch& h = *roles[r];
        if (h.mode != 0 || h.cin != roles[0]->cin || h.cout != roles[0]->cout || h.tmem_cols != roles[0]->tmem_cols) return cudaErrorInvalidValue;
        if (h.jobs_dev == nullptr) {
            GemmJob* dev = nullptr;
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dev), static_cast<size_t>(h.n_jobs) * sizeof(GemmJob));
            if (e != cudaSuccess) return e;
            e = cudaMemcpy(dev, h.jobs, static_cast<size_t>(h.n_jobs) * sizeof(GemmJob), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(dev); return e; }
            h.jobs_dev = dev;
        }
        P.role[r] = static_cast<const GemmParams&>(h);
        P.cta0[r] = grid;
        grid += ctas[r];
        P.expect[r] = static_cast<uint32_t>(kEpiWarpsPerTile * h.n_jobs);
        const GemmSmem lay = gemm_smem_layout(h.cin, h.cout, h.n_wtaps, h.n_jobs, h.ext_alloc, h.groups_per_stage, h.n_stages);
        smem = lay.total > smem ? lay.total : smem;
    }
    for (int r = n_roles; r <= kMaxRoles; ++r) P.cta0[r] = grid;
    P.n_roles = n_roles;
    P.sync = sync;
    const GemmLaunch& h0 = *roles[0];
    const bool dual = h0.tmem_cols == 256;
#define LD_PIPE_CASE(ci, co)                                                                                              \
    if (h0.cin == ci && h0.cout == co) {                                                                                  \
        if (co <= 32 && dual) return launch_pipe_typed<ci, co, (co <= 32)>(P, grid, m_tiles, M, smem, stream);             \
        return launch_pipe_typed<ci, co, false>(P, grid, m_tiles, M, smem, stream);                                       \
    }
    LD_PIPE_CASE(64, 64); LD_PIPE_CASE(32, 32); LD_PIPE_CASE(16, 16); LD_PIPE_CASE(64, 32); LD_PIPE_CASE(32, 16);
#undef LD_PIPE_CASE
    return cudaErrorInvalidValue;
}

void gemm_release(GemmLaunch& h) {
    if (h.jobs_dev) cudaFree(const_cast<GemmJob*>(h.jobs_dev));
    h.jobs_dev = nullptr;
}

// Chooses the smem ring depth for a launch (host side).
int gemm_pick_stages(int cin, int cout, int n_wtaps, int n_jobs, int ext_alloc, int groups_per_stage, int max_stages, unsigned smem_cap) {
    if (max_stages > kMaxStages || max_stages < 2) max_stages = kMaxStages;
    for (int n = max_stages; n >= 2; --n) {
        if (gemm_smem_layout(cin, cout, n_wtaps, n_jobs, ext_alloc, groups_per_stage, n).total <= smem_cap) return n;
    }
    return 0;
}

}  // namespace ld
