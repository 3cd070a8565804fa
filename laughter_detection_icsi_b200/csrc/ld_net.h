// Launch descriptions and launcher prototypes of the non-GEMM kernels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "ld_types.h"

namespace ld {

constexpr int kMaxHeadFeat = 64;
constexpr int kMaxHeadRows = 16;
constexpr int kMaxStemJobs = 4;

// Channels laid end to end in a sequence; all arrays live in device memory.
struct ChannelTable {
    const long long* seq_off;   // [n_chan] first sequence row of the channel
    const long long* frames;    // [n_chan] T_c
    const long long* feat_off;  // [n_chan] first row of the channel in the caller's feature/prob arrays
    int n_chan;
};

struct StemJob {
    __half* out;
    long long kc_stride;
    int row_shift;
    int mask;
};
struct StemLaunch {
    StemJob jobs[kMaxStemJobs];
    const float* w;      // [64][9]
    const float* scale;  // [64]
    const float* shift;  // [64]
    int n_jobs, W;
    int wp;              // pixels per stored row: W + 1 (one shared pad column) or W + 2
};

struct HeadRow {
    const __half* plane;
    long long kc_stride;
    int row_shift;
    int pad_;
};
struct HeadLaunch {
    HeadRow rows[kMaxHeadRows];
    const float* params;  // bn2 scale/shift, linear1, bn3 scale/shift, linear2 (see ld_net.cu)
    int n_feat, C, groups, wp;
    int split;            // the planes are [hi | lo]: chunks C/8.. hold the fp16 rounding residual, added back in fp32
};

cudaError_t launch_stem(const StemLaunch& L, const ChannelTable& ct, const float* feats, long long chunk_row0,
                        int rows, cudaStream_t stream);
cudaError_t launch_head(const HeadLaunch& L, const ChannelTable& ct, float* probs, long long chunk_row0, int nb,
                        cudaStream_t stream);

// ld_gemm.cu
cudaError_t launch_gemm_taps(GemmLaunch& h, int m_tiles, int M, int num_sms, cudaStream_t stream);
void gemm_release(GemmLaunch& h);   // frees the device copy of the job table
// layer-pipelined launch of consecutive conv layers of one shape (ld_types.h, GemmMultiParams); ctas[r] = CTAs of role r
cudaError_t launch_gemm_pipe(GemmLaunch* const* roles, int n_roles, const int* ctas, const GemmSync& sync, int m_tiles, int M,
                             cudaStream_t stream);
int gemm_pick_stages(int cin, int cout, int n_wtaps, int n_jobs, int ext_alloc, int groups_per_stage, int max_stages,
                     unsigned smem_cap = 227u * 1024u);

// ld_launch.cpp
struct HostTap {
    const void* src;      // pixel 0, chunk 0 of the source plane
    long long kc_stride;  // elements between its channel chunks
    int shift;            // pixel shift of the tap
    int wslab;            // weight slab index: 3x3 block b (0 = hi, 1 = lo weights): 9 * b + ky * 3 + kx; slabs after the
                          // blocks (identity / 1x1 weights) are stand-alone
    int half_k = 0;       // split precision: multiply only the first half of the channel chunks (the hi half of [hi | lo])
};
struct HostJob {   // one OUTPUT plane; gemm_build_launch packs chains of them into GemmJobs
    std::vector<HostTap> taps;
    void* out0 = nullptr;
    void* out1 = nullptr;
    long long out_kc_stride = 0;
};
struct GemmTuning {
    int group_span, max_stages, stage_bytes, max_outs, n_rings_max, issuers_wide, issuers_narrow;
    int dual_narrow;   // inference layers with cout <= 16 (1, default) / <= 32 (2) run two CTAs per SM (256 accumulator columns each)
    int issuers_dual;  // MMA-issuing warps of such a CTA (2: the SM keeps four issuers and 128 columns per accumulator stage)
};
GemmTuning gemm_tuning_from_env();
bool gemm_build_launch(GemmLaunch& L, const std::vector<HostJob>& jobs, const GemmTuning& tune, std::string& err);

// ld_fbank.cu
struct FbankMel {           // sparse view of the (257, F) filterbank, built from the caller's matrix
    const float* weights;   // packed runs of 4-bin vectors, filter after filter (zero weights on the alignment padding)
    const int* lo;          // [F] first bin of filter k's run (multiple of 4)
    const int* len;         // [F] run length in 4-bin vectors
    const int* off;         // [F] offset of the run inside weights (multiple of 4)
    int n_filters;
};
cudaError_t launch_pcm_sum(const int16_t* pcm, long long n, unsigned long long* sum_biased, cudaStream_t stream);
cudaError_t launch_fbank(const int16_t* pcm, long long n_samples, long long n_frames,
                         const unsigned long long* sum_biased, int per_frame,
                         const FbankMel& mel, const float* tables, float* feats, cudaStream_t stream);
void fbank_host_tables(float* out /* kFbankTableFloats */);
constexpr int kFbankTableFloats = 400 + 2 * 256 + 2 * 257;

// ld_segment.cu
cudaError_t launch_segment_runs(const void* probs, int is_f64, const ChannelTable& ct, long long total_frames,
                                const double* thr_cmp_d, const double* thr_raw_d, int n_thr, int* starts, int* ends,
                                int* chans, int* counts, int cap, int* block_counts, cudaStream_t stream);
size_t segment_scratch_ints(long long total_frames, int n_thr);
cudaError_t launch_filtfilt(const void* probs, int is_f64, long long n, const double* b, const double* a, double* out,
                            double* scratch, cudaStream_t stream);
size_t filtfilt_scratch_doubles(long long n);

// ld_gather.cu
cudaError_t launch_gather_windows(const float* feats, const long long* track_off, const long long* track_len, const int* triples,
                                  int n_windows, int n_frames, int F, float pad_value, float* out, cudaStream_t stream);

}  // namespace ld
