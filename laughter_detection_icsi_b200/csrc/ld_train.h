// Training path of ResNetBigger (reference train.py:261-297 -> models.py forward in .train() mode + autograd):
// dense per-sample evaluation, BatchNorm with batch statistics, dropout by caller-provided masks, bf16 operands with
// fp32 accumulation.  Convolutions (forward and data gradient) run on the same tcgen05 tap-list GEMM kernel as
// inference (ld_gemm.cu, MODE 1); weight gradients, BatchNorm forward/backward, the stem and the head run on CUDA cores.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ld_types.h"

namespace ld {

// A training tensor: B images of H x W x C in bf16, channel-chunk planar [C/8][pixels][8] like the inference planes.
// plain: one plane, image b occupies rows [b*hp, (b+1)*hp) with one zero row above/below and one zero column left/right.
// quad : four planes by (row parity, column parity) of the real coordinates, each a plain plane of the halved size --
//        what a stride-2 conv reads with unit-stride taps.
struct TPlane {
    __nv_bfloat16* base[4];
    long long kc_stride;   // elements between channel chunks
    int H, W, C;
    int hp, wp;            // rows per image incl. the 2 pad rows, padded width (of each stored plane)
    int quad;
};

struct TrainParamInfo {
    std::string name;
    long long offset, numel;
};

// Launch description of the tensor-core weight-gradient kernel (ld_wgrad.cu), passed by value.
struct WgradLaunch {
    const __nv_bfloat16* seg_src[9];   // staged pixel runs of X ("segments"): pixel 0 / chunk 0 of the source plane
    long long seg_kc_stride[9];
    int seg_shift[9];                  // first pixel of the run relative to the tile's first pixel
    int n_seg;
    const __nv_bfloat16* dz;           // gradient of the conv output (plain plane)
    long long dz_kc_stride;
    int grp_seg0[6];                   // MMA group g: A starts at segment grp_seg0[g], pixel grp_px_off[g]
    int grp_px_off[6];
    int grp_tap[6][8];                 // weight tap (ky * k + kx) of row block s of the group's accumulator, or -1 (ignored)
    int n_grp;
    float* dw;                         // (cout, cin, k, k) fp32, accumulated atomically
    int n_taps, cin, cout;
    long long M;                       // pixels of the dZ plane to reduce over
    int n_stages;
    uint32_t stage_bytes;
};
cudaError_t launch_wgrad_mma(const WgradLaunch& L, int num_sms, cudaStream_t stream);
bool wgrad_plan_smem(WgradLaunch& L);

class TrainNet;  // opaque (ld_train.cu)

TrainNet* train_create(int max_batch, int num_sms, const NetConfig& cfg, std::string& err);
void train_destroy(TrainNet* net);
const std::vector<TrainParamInfo>& train_param_table(const TrainNet* net);
long long train_num_params(const TrainNet* net);
const std::vector<TrainParamInfo>& train_bn_table(const TrainNet* net);   // name, offset into bn_stats, C
int train_num_bn_stats(const TrainNet* net);   // floats of the batch-statistics output: per BatchNorm mean[C] then var[C]
// params: flat fp32 device buffer in train_param_table order.  x: (B,100,44) fp32 device.  mask1 (B,48), mask2 (B,32):
// float 0/1 keep masks of the two dropout sites (models.py:232,235).  probs: (B) fp32 out.  bn_stats: device out.
cudaError_t train_forward(TrainNet* net, const float* params, const float* x, int B, const float* mask1, const float* mask2,
                          float dropout_p, float* probs, float* bn_stats, cudaStream_t stream, std::string& err);
// dprobs: dL/dprobs (B) fp32 device.  grads: flat fp32 device buffer (train_param_table order), overwritten.
cudaError_t train_backward(TrainNet* net, const float* dprobs, float* grads, cudaStream_t stream, std::string& err);
long long train_kernel_launches(const TrainNet* net);
// K8: clip_grad_norm_(max_norm) + Adam step on flat fp32 vectors (scratch2: 2 device floats; [1] receives the gradient norm).
cudaError_t clip_adam_step(float* params, const float* grads, float* m, float* v, long long n, float max_norm, float lr, float b1,
                           float b2, float eps, long long step, float* scratch2, cudaStream_t stream);
// Same with the step count in device memory (incremented by the call): replayable from a captured CUDA graph.
cudaError_t clip_adam_step_dev(float* params, const float* grads, float* m, float* v, long long n, float max_norm, float lr, float b1,
                               float b2, float eps, long long* step_d, float* scratch2, cudaStream_t stream);
int train_debug_checksums(TrainNet* net, double* out, int cap);
long long train_debug_read(TrainNet* net, int kind, int index, float* out, int* dims4);

}  // namespace ld
