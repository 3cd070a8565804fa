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
constexpr int kMaxGroups = 6;    // distinct smem loads per job
constexpr int kMaxTaps = 10;     // MMA taps per job (3x3 = 9)
constexpr int kMaxJobs = 16;     // jobs (output planes) per launch
constexpr int kGuardRows = 104;  // zero guard rows allocated before/after every plane

enum OutMode : int32_t { OUT_PLAIN = 0, OUT_COLSPLIT = 1 };

// ---- device-side launch description (lives in global memory, copied to smem by the kernel) ----
struct GemmGroup {
    const __half* src;   // pixel 0, chunk 0 of the source plane
    int64_t kc_stride;   // elements between channel chunks of the source plane
    int32_t shift;       // first pixel to load relative to the tile's first output pixel
    int32_t ext;         // pixels to load (128 + span of the taps that share this load)
};
struct GemmTap {
    int16_t group;  // index into groups[]
    int16_t off;    // pixel offset inside the group's load
    int16_t wtap;   // weight tap index (ky*3+kx, or 0 for 1x1)
    int16_t pad_;
};
struct GemmJob {
    GemmGroup groups[kMaxGroups];
    GemmTap taps[kMaxTaps];
    // MMA-issue view of the taps, in smem-descriptor units (16 bytes), ordered by group:
    uint16_t tap_a16[kMaxTaps];      // pixel offset of the tap inside its group's load (1 pixel = 16 B per chunk)
    uint16_t tap_b16[kMaxTaps + 2];  // offset of the tap's weight slab inside the weight block
    uint8_t group_taps[kMaxGroups + 2];  // taps per group
    int32_t n_groups, n_taps;
    __half* out0;  // PLAIN: the plane; COLSPLIT: even-column plane
    __half* out1;  // COLSPLIT: odd-column plane
    int64_t out_kc_stride;
    const __half* res;  // residual plane (same geometry as the job's output pixels) or null
    int64_t res_kc_stride;
    int32_t res_shift;
    int32_t pad_;
};
struct GemmLaunch {
    GemmJob jobs[kMaxJobs];
    const __half* weights;  // [n_wtaps][cin/8][cout][8] fp16
    const float* scale;     // [cout] folded BatchNorm scale
    const float* shift;     // [cout] folded BatchNorm shift (+ conv bias)
    int32_t n_jobs, cin, cout, n_wtaps;
    int32_t relu, wp, out_mode, wp2;
    int32_t ext_alloc;  // pixels reserved per smem stage (>= every group's ext, multiple of 8)
    int32_t hp;         // >0: rows per image incl. 2 pad rows (dense layout), pad rows forced to zero
    int32_t n_stages;
    int32_t pad_;
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
