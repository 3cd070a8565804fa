// Internal data structures shared by the host planner (ld_plan.cpp), the context (ld_api.cu) and the
// kernels.  Vocabulary:
//   row      one 10 ms feature frame position in a "sequence" (channels laid end to end, separated by
//            >=100 all-zero rows).  A window is the 100 rows starting at row b (datasets.py:85-93).
//   level    the activations after one conv of ResNetBigger, seen per window: H local rows x W cols x C.
//   plane    one stored array of a level: either the INTERIOR plane (value at global row g, identical
//            for every window that contains g far enough from its edges) or a SPECIFIC plane for one
//            local row j (value of local row j of the window starting at row b, indexed by b) --
//            rows whose receptive field touches the window's zero padding.
//   pixel    flattened (row, padded col) index p = row * wp + col, col in [0, wp), cols 0 and wp-1 are
//            zero padding.  Planes are fp16, channel-chunk planar: [C/8][pixels][8].
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_fp16.h>

namespace ld {

constexpr int kTileM = 128;      // output pixels per MMA tile (UMMA M)
constexpr int kBoxPixels = 136;  // pixels per TMA box: a tile plus the +-1 pixel halo of a 3-tap row, rounded to 8
constexpr int kMaxGroups = 6;    // distinct smem loads per job
constexpr int kMaxTaps = 10;     // MMA taps per job (3x3 = 9)
constexpr int kMaxJobs = 16;     // jobs (output planes) per launch
constexpr int kGuardRows = 104;  // zero guard rows allocated before/after every plane

enum OutMode : int32_t { OUT_PLAIN = 0, OUT_COLSPLIT = 1 };

// ---- device-side launch description (passed to the kernel BY VALUE as a __grid_constant__ parameter: every field the
// warp-specialised roles index is then read through the uniform constant path, no shared-memory staging) ----
struct GemmGroup {
    const __half* src;   // pixel 0, chunk 0 of the source plane
    int64_t kc_stride;   // elements between channel chunks of the source plane
    const void* tmap;    // CUtensorMap of the source plane: (8 halfs, pixels incl. guards, C/8 chunks), box (8, kBoxPixels, C/8)
    int32_t pixel0;      // tensor coordinate of the plane's pixel 0 (= guard pixels in front of it)
    int32_t shift;       // first pixel to load relative to the tile's first output pixel
};
// One MMA tap as the issuing warp sees it (16-byte smem-descriptor units):
//   bits 0..13  A offset inside the current smem stage (group slot + pixel shift)    bits 14..27  B offset of the weight slab
//   bit 28      first tap of a stage: wait for its "full" barrier                   bit 29       last tap of a stage: commit "empty"
constexpr uint32_t kTapFirst = 1u << 28, kTapLast = 1u << 29;
constexpr uint32_t kTapPass = 1u << 30;  // this tap's wait is the tile's last one: the next issuer warp may start waiting
constexpr int kTapWords = 12;  // kMaxTaps rounded up to whole 16-byte loads
struct alignas(16) GemmJob {
    uint32_t tapw[kTapWords];
    GemmGroup groups[kMaxGroups];
    int32_t n_groups, n_taps;
    __half* out0;  // PLAIN: the plane; COLSPLIT: even-column plane
    __half* out1;  // COLSPLIT: odd-column plane
    int64_t out_kc_stride;
};
struct GemmLaunch {
    GemmJob jobs[kMaxJobs];
    const __half* weights;  // [n_wtaps][cin/8][cout][8] fp16, BatchNorm scale folded in
    const float* shift;     // [cout] folded BatchNorm shift (+ conv bias)
    int32_t n_jobs, cin, cout, n_wtaps;
    int32_t relu, wp, out_mode, wp2;
    int32_t ext_alloc;  // pixels per loaded group (>= every group's extent, multiple of 8)
    int32_t hp;         // >0: rows per image incl. 2 pad rows (dense layout), pad rows forced to zero
    int32_t n_stages;
    int32_t loader;     // 0: one bulk copy per channel chunk; 1: one TMA tensor copy per group
    int32_t groups_per_stage;  // 1: every group has its own smem stage/barrier; >1: a stage holds all groups of a tile
    uint32_t wp_magic;  // floor(2^32 / wp) + 1: row = umulhi(pixel, wp_magic)
    int32_t n_rings;    // 2: two producer/issuer pipelines over half the stages each; 1: a single ring
    int32_t mode;       // 0: inference (fp16, shift + ReLU epilogue); 1: training (bf16, raw output + channel statistics)
    float* stats;       // mode 1: [2 * cout] per-channel sum and sum of squares (atomically accumulated), or null
    unsigned long long* prof;  // optional: 8 cycle counters per launch (see ld_gemm.cu), null = off
};

// ---- host-side plan (plane ids instead of pointers) ----
struct PlaneSpec {
    int id;
    int C;       // channels (multiple of 8)
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
};

Plan build_stream_plan(const NetConfig& cfg);
std::string plan_to_json(const Plan& plan);

}  // namespace ld
