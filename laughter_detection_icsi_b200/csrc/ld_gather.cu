// LAD window gather: builds a training batch of 1-second windows straight from whole-track log-mel features that are
// resident in HBM -- the GPU counterpart of the reference's cut construction
//     row_track.truncate(offset=sub_start, duration=sub_duration).pad(duration=1.0)        (compute_features.py:167)
// followed by PrecomputedFeatures()(cuts) in LadDataset.__getitem__ (datasets.py:49-68).  A window is an index triple
// (track, first frame, frames); rows past `frames` or past the end of the track hold the pad value (Lhotse pads log-domain
// features with LOG_EPSILON).  One warp copies one 44-float row; HBM-bound: 176 B read + 176 B written per row.
#include <cstdint>
#include <cuda_runtime.h>

#include "ld_net.h"

namespace ld {

__global__ void __launch_bounds__(256)
gather_windows_kernel(const float* __restrict__ feats, const long long* __restrict__ track_off, const long long* __restrict__ track_len,
                      const int* __restrict__ triples, int n_windows, int n_frames, int F, float pad_value, float* __restrict__ out) {
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;   // (window, local row)
    const int lane = threadIdx.x & 31;
    if (row >= static_cast<long long>(n_windows) * n_frames) return;
    const int b = static_cast<int>(row / n_frames), r = static_cast<int>(row - static_cast<long long>(b) * n_frames);
    const int track = triples[3 * b], first = triples[3 * b + 1], n = triples[3 * b + 2];
    const long long src_row = static_cast<long long>(first) + r;
    const bool real = r < n && src_row >= 0 && src_row < track_len[track];
    const float* src = feats + (track_off[track] + src_row) * F;
    float* dst = out + row * F;
    for (int c = lane; c < F; c += 32) dst[c] = real ? __ldg(src + c) : pad_value;
}

cudaError_t launch_gather_windows(const float* feats, const long long* track_off, const long long* track_len, const int* triples,
                                  int n_windows, int n_frames, int F, float pad_value, float* out, cudaStream_t stream) {
    if (n_windows <= 0) return cudaSuccess;
    const long long warps = static_cast<long long>(n_windows) * n_frames;
    gather_windows_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(feats, track_off, track_len, triples, n_windows,
                                                                                               n_frames, F, pad_value, out);
    return cudaGetLastError();
}

}  // namespace ld
