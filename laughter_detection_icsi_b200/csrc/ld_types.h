// Internal data structures shared by the host planner (ld_plan.cpp), the context (ld_api.cu) and the
// kernels.  Vocabulary:
//   row      one 10 ms feature frame position in a "sequence" (channels laid end to end, separated by
//            >=100 all-zero rows).  A window is the 100 rows starting at row b (datasets.py:85-93).
//   level    the activations after one conv of ResNetBigger, seen per window: H local rows x W cols x C.
//   plane    one stored array of a level: either the INTERIOR plane (value at global row g, identical
//            for every window that contains g far enough from its edges) or a SPECIFIC plane for one
//            local row j (value of local row j of the window starting at row b, indexed by b) --
//            rows whose receptive field touches the window's zero padding.
//   pixel    flattened (row, padded col) index p = row * wp + col, col in [0, wp), wp = W + 1: col 0 is zero
//            padding and serves BOTH as the left pad of its row and as the right pad of the row before
//            (p + 1 of a row's last column is the next row's col 0), so 1 / (W + 1) instead of 2 / (W + 2) of
//            the pixels are padding.  Planes are fp16, channel-chunk planar: [C/8][pixels][8].
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ld {

constexpr int kTileM = 128;      // output pixels per MMA tile (UMMA M)
constexpr int kMaxGroups = 16;   // distinct smem loads per job
constexpr int kMaxTaps = 64;     // MMA taps per job (split precision doubles the taps of a job: hi and lo weight blocks)
constexpr int kMaxLaunchTaps = 384;  // MMA taps per launch (all jobs); the table travels as a kernel parameter (6 KB of the 32 KB)
constexpr int kMaxOuts = 8;      // output planes per job (their accumulators sit side by side in TMEM)
constexpr int kTmemCols = 512;   // accumulator columns per SM, split into n_issuers stages: n_outs * cout <= 512 / n_issuers
constexpr int kMaxJobs = 16;     // jobs per launch (the planner also splits a layer into launches of at most this many output planes)
constexpr int kGuardRows = 104;  // zero guard rows allocated before/after every plane

// cudaFuncSetAttribute applies to the CURRENT device only: one flag per device, so that a second context on another GPU of the
// same process configures its kernels too.
struct PerDeviceOnce {
    bool done[64] = {};
    bool& flag() {
        int d = 0;
        cudaGetDevice(&d);
        return done[d & 63];
    }
};

// dynamic shared memory a CTA of a conv launch may use: 228 KB per SM with 1 KB reserved per CTA; two CTAs per SM share it
inline unsigned gemm_smem_cap(int tmem_cols) { return tmem_cols == 256 ? 113u * 1024u : 227u * 1024u; }

enum OutMode : int32_t { OUT_PLAIN = 0, OUT_COLSPLIT = 1 };

// ---- device-side launch description (passed to the kernel BY VALUE as a __grid_constant__ parameter: every field the
// warp-specialised roles index is then read through the uniform constant path, no shared-memory staging) ----
//
// A JOB is a set of up to kMaxOuts output planes that are computed together for one tile of 128 pixels.  The output planes
// of a ResNet conv layer come in chains (local rows j, j+1, ... of the window): input row i feeds output rows i-1, i, i+1
// with the weight rows ky = 2, 1, 0.  The kernel therefore walks the INPUT planes of the chain: each is loaded into shared
// memory once and multiplied by the weights of all the outputs it feeds in ONE MMA whose N spans their accumulators
// (weights stacked [ky=2 | ky=1 | ky=0] along N in shared memory, accumulators side by side in TMEM in chain order).
// An SS-mode MMA re-reads its 4 KB A operand whatever N is, so this divides the shared-memory operand traffic -- the
// measured limiter of every layer of this net (cout <= 64) -- by up to three.
struct GemmGroup {
    const __half* src;   // pixel 0, chunk 0 of the source plane
    int64_t kc_stride;   // elements between channel chunks of the source plane
    int32_t shift;       // first pixel to load relative to the tile's first output pixel
    int32_t pad_;
};
// One MMA tap as the issuing warp sees it, pre-digested on the host (offsets in 16-byte smem-descriptor units):
//   x: bits 0..13  A offset inside the current smem stage (group slot + pixel shift)
//      bit 28 first tap of a stage: wait for its "full" barrier      bit 29 last tap of a stage: commit "empty"
//      bit 30 this wait is the tile's last one: the ring's other issuer may start waiting
//   y: low word of the B smem descriptor relative to the weights: bits 0..13 offset of the weight rows, bits 16..29 LBO
//      (rows per channel chunk: 3 * cout inside a stacked block, cout for a stand-alone slab [cin/8][cout][8])
//   z: first accumulator column                w: N / 8 << 17 (the N field of the instruction descriptor)
//      bit 27 half K: only the first cin/32 K steps (split precision: the hi half of a [hi | lo] activation against lo weights)
constexpr uint32_t kTapFirst = 1u << 28, kTapLast = 1u << 29, kTapPass = 1u << 30, kTapHalfK = 1u << 27;
struct GemmOut {
    __half* out0;  // PLAIN: the plane; COLSPLIT: even-column plane
    __half* out1;  // COLSPLIT: odd-column plane
};
struct alignas(16) GemmJob {          // what the producers and the epilogue need of a job (copied to shared memory)
    GemmGroup groups[kMaxGroups];
    GemmOut outs[kMaxOuts];
    int32_t n_groups, n_taps, n_outs, n_stages;  // n_stages = ceil(n_groups / groups_per_stage)
    int64_t out_kc_stride;
    int32_t dep_back[3];   // pipelined launches: last pixel (relative to the tile's first) this job reads of a plane written by the
                           // role 1, 2, 3 places before its own in the same launch, or kNoDep: the tile's loads wait until the
                           // m-tiles of that role up to that pixel are complete
    int32_t pad_;
};
constexpr int32_t kNoDep = INT32_MIN;
// Position of weight row ky inside a stacked block.
__host__ __device__ inline int w_stack_row(int order, int ky) { return order == 2 ? (ky == 2 ? 0 : ky + 1) : 2 - ky; }

struct GemmJobTaps {   // where a job's tap program sits in GemmParams::taps
    uint16_t tap0, n_taps, n_stages, pad_;
};
// mode 1: BatchNorm-BACKWARD statistics fused into the data-gradient launch that produces dy (GemmParams::stats_kind = 1): the
// epilogue reduces sum g and sum g * xhat (g = dy * [y > 0], xhat = the BatchNorm's normalised input) from its fp32
// accumulators into `stats`, which saves the separate reduction pass over dy / y / z.  z and y are plain planes with the
// pixel geometry of the launch's output.
struct GemmBwdStats {
    const void* z;          // conv output the BatchNorm normalised (bf16, chunk-planar)
    const void* y;          // activation plane for the ReLU mask (mask_mode 1)
    long long z_kc, y_kc;   // elements between channel chunks
    const float* fwd_sums;  // [2 cout] sum z, sum z^2 of the forward pass
    const float* gamma;     // [cout]
    const float* beta;      // [cout]
    float inv_n;
    int32_t mask_mode;      // 0: g = dy;  1: g = dy * [y > 0];  2: g = dy * [z * ya + yb > 0] (y = relu(bn(z)), no residual)
};
struct GemmParams {         // the kernel's __grid_constant__ parameter
    // The tap programs of all jobs, densely packed.  They stay in parameter (constant) space on purpose: the issuing thread
    // indexes them with warp-uniform values, so the loads, the descriptor arithmetic and the tcgen05.mma operands all live
    // in uniform registers (a table read through shared memory lands in vector registers and costs ~7 R2UR per MMA).
    uint4 taps[kMaxLaunchTaps];
    GemmJobTaps job_taps[kMaxJobs];
    const GemmJob* jobs_dev;  // [n_jobs] in device memory; every CTA copies the table into shared memory (producers and
                              // epilogue index it per lane; the constant path would thrash on it)
    const __half* weights;  // [n_wtaps][cin/8][cout][8] fp16, BatchNorm scale folded in
    const float* shift;     // [cout] folded BatchNorm shift (+ conv bias)
    int32_t n_jobs, cin, cout, n_wtaps;
    int32_t relu, wp, out_mode, wp2;
    int32_t w_real;     // real columns of a row: columns 1..w_real hold data, the others are zero padding.  Inference planes share ONE
                        // pad column between consecutive rows (wp = w_real + 1: pixel p - 1 of a row's first column and pixel
                        // p + 1 of its last column are the same kind of zero); 0 = the dense training layout (wp - 2)
    int32_t ext_alloc;  // pixels per loaded group (>= every group's extent, multiple of 8)
    int32_t ext_copy;   // pixels the bulk copies actually transfer per channel chunk of a group: the largest group extent (the
                        // MMAs never read the rounding slack of ext_alloc)
    int32_t hp;         // >0: rows per image incl. 2 pad rows (dense layout), pad rows forced to zero
    int32_t n_stages;
    int32_t groups_per_stage;  // consecutive groups of a job that share one smem stage (one barrier round trip)
    uint32_t wp_magic;  // floor(2^32 / wp) + 1: row = umulhi(pixel, wp_magic)
    int32_t w_stack;    // 1, 2: 3x3 weights are re-stacked in smem as [kx][cin/8][ky order][cout][8], order ky = 2,1,0 (1) or
                        // 2,0,1 (2, stride-2 convs), see w_stack_row; slabs from 9 * w_blocks on stay slabs.  0: plain slabs
    int32_t w_blocks;   // stacked 3x3 blocks at the head of the weights: 1, or 2 in split precision (hi weights, lo weights)
    int32_t split_out;  // split precision: the epilogue also writes the fp16 rounding residual as channel chunks cout/8.. of the
                        // output plane ([hi | lo] planes)
    int32_t dbg;        // timing experiments only (LD_GEMM_DBG, results are garbage): bit 0 one copy per smem stage, bit 1 no MMAs,
                        // bit 2 no global stores, bit 3 no TMEM reads/clears in the epilogue
    int32_t n_issuers;  // MMA-issuing warps = accumulator stages (2: 256 columns each, 4: 128 columns each)
    int32_t n_rings;    // 2: two producer/issuer pipelines over half the stages each; 1: a single ring
    int32_t l2_prefetch; // producer: tiles ahead of the shared-memory ring whose operands are requested into L2 (0 = off)
    int32_t tmem_cols;  // accumulator columns this CTA allocates: 512 (one CTA per SM) or 256 (narrow layers: two CTAs per SM, whose
                        // barrier / issue / epilogue latencies then overlap)
    int32_t mode;       // 0: inference (fp16, shift + ReLU epilogue); 1: training (bf16, raw output + channel statistics)
    float* stats;       // mode 1: [2 * cout] per-channel sum and sum of squares (atomically accumulated), or null
    int32_t stats_kind; // 0: sum z, sum z^2 of the output (BatchNorm forward);  1: sum g, sum g * xhat (BatchNorm backward, `bwd`)
    GemmBwdStats bwd;
    unsigned long long* prof;  // optional: 8 cycle counters per launch (see ld_gemm.cu), null = off
};
// ---- layer-pipelined launches ----
// Consecutive conv layers of one shape (cin, cout) and one resolution run as ROLES of a single launch: role r owns the CTAs
// [cta0[r], cta0[r + 1]) with its own weights, job table and tap programs, walks its m-tiles in increasing order and signals
// every finished (job, m-tile) in done[r][m]; the producer warps of role r + 1 load a tile only once the m-tiles of role r it
// reads are complete.  The layer's output is then consumed out of L2 a few dozen m-tiles after it was written instead of
// making a round trip through HBM between two launches.
constexpr int kMaxRoles = 4;
constexpr int kEpiWarpsPerTile = 8;   // arrivals per finished tile (one per epilogue warp)
struct GemmSync {
    unsigned* done;       // [n_roles][m_cap] arrivals per m-tile of this launch; all zero when the launch starts
    unsigned* done_next;  // the other buffer (previous pipelined launch's counters): zeroed by this launch for the next one
    int32_t m_cap;        // counters per role
    int32_t lead_max;     // a role runs at most this many m-tiles ahead of its consumer (keeps the hand-over inside L2)
    int32_t dbg;          // timing experiments only (LD_GEMM_PIPE_DBG with LD_GEMM_PROF, results are garbage): bit 0 no dataflow waits,
                          // bit 1 no back-pressure waits, bit 2 no fence before the completion signal
};
struct GemmMultiParams {    // the pipelined kernel's __grid_constant__ parameter
    GemmParams role[kMaxRoles];
    GemmSync sync;
    int32_t n_roles;
    int32_t cta0[kMaxRoles + 1];
    uint32_t expect[kMaxRoles];   // arrivals that complete an m-tile of role r: kEpiWarpsPerTile * its job count
};

struct GemmLaunch : GemmParams {   // host side: the parameters plus the job table gemm_build_launch fills;
    GemmJob jobs[kMaxJobs];        // launch_gemm_taps uploads it on first use (jobs_dev), gemm_release frees it
    const uint4* job_tapw(int j) const { return taps + job_taps[j].tap0; }
};

// ---- host-side plan (plane ids instead of pointers) ----
struct PlaneSpec {
    int id;
    int C;       // channels (multiple of 8)
    bool split = false;  // split precision: stored as [hi | lo] = 2 * C channel chunks
    int wp;      // padded width
    std::string tag;
};
struct TapSpec {
    int plane;   // source plane id
    int shift;   // pixel shift
    int wtap;
};
struct JobSpec {
    std::vector<TapSpec> taps;
    int out0 = -1, out1 = -1;
    int res_plane = -1, res_shift = 0;
    std::string tag;
};
struct ConvLaunchSpec {
    std::string conv;  // state_dict prefix of the conv, e.g. "block2.0.conv1"
    std::string bn;    // state_dict prefix of the BatchNorm that follows
    int cin, cout, ksize, relu, wp, out_mode, wp2, hp;
    int w_real = 0;   // real columns per row (wp - 1: one shared pad column)
    int split_in = 0, split_out = 0, split_w = 0;   // split precision: [hi | lo] input planes / output planes / hi + lo weights
    std::vector<JobSpec> jobs;
};
struct StemJobSpec {
    int out_plane;
    int row_shift;   // feature row = plane row + row_shift
    int mask;        // bit ky set -> tap row ky-1 contributes (others are the window's zero padding)
};
struct HeadRowSpec {
    int plane;
    int row_shift;
};
struct Plan {
    int H = 100, W = 44;             // config.FEAT: num_samples x num_filters
    std::vector<PlaneSpec> planes;
    std::vector<StemJobSpec> stem;   // conv1+bn1+relu on the fp32 features (CUDA cores)
    int stem_wp = 0;
    std::vector<ConvLaunchSpec> convs;
    std::vector<HeadRowSpec> head_rows;  // the 12 local rows AvgPool2d(4) consumes
    int head_wp = 0, head_C = 0, head_pool_w = 0, head_pool_groups = 0;
    double macs_per_row = 0;       // executed MACs per sequence row (all planes), for roofline accounting
    double gemm_macs_per_row = 0;  // the tensor-core share of it (everything but the fp32 stem)
};

// ResNetBigger topology needed by the planner (models.py:181-244 in the reference).
struct NetConfig {
    int H = 100, W = 44;
    int filters[4] = {64, 32, 16, 16};
    int linear_in = 48;
    int precision = 0;   // ld_precision: 1 = blocks 2-4 carry weights and activations as hi + lo fp16 pairs
};

Plan build_stream_plan(const NetConfig& cfg);
std::string plan_to_json(const Plan& plan);

}  // namespace ld
