// K4: threshold run extraction for many thresholds in one pass, and K5: zero-phase 2nd-order low-pass.
//
// K4 replaces the per-setting Python loop of laugh_segmenter.get_laughter_instances
// (reference laugh_segmenter.py:87-105): runs of frames whose clamped probability is strictly above the
// threshold, reported as (first_frame, last_frame).  The float64 frame->seconds conversion and the strict
// `end - start > min_length` filter stay on the host in double precision (ld_filter_min_length), because
// their rounding decides which 20-frame runs survive.
// Ordered compaction: warp ballots + block scan, with a per-block count pass and a scan pass in between.
//
// K5 replaces scipy.signal.filtfilt(butter(2, 0.01)) (reference laugh_segmenter.py:49-55): the IIR
// recurrence is linear, so each thread filters a chunk from a zero state, the chunk end states are
// chained through the 2x2 transition matrix power, and the chunk is re-filtered from its true state.
#include <cstdint>
#include <cuda_runtime.h>

#include "ld_net.h"

namespace ld {

constexpr int kSegThreads = 256;
constexpr int kSegIters = 8;
constexpr int kSegSpan = kSegThreads * kSegIters;

template <typename T>
__device__ __forceinline__ bool above(const T* probs, long long i, double thr_cmp, double thr_raw) {
    const T p = probs[i];
    if (p > T(1)) return 1.0 > thr_raw;     // fix_over_underflow: prob > 1 -> 1      (laugh_segmenter.py:64-66)
    if (p <= T(0)) return 1e-7 > thr_raw;   //                     prob <= 0 -> 1e-7  (laugh_segmenter.py:68-70)
    return static_cast<double>(p) > thr_cmp;  // NaN compares false, like Python
}

__device__ __forceinline__ int seg_channel(const ChannelTable& ct, long long i, long long& local) {
    int lo = 0, hi = ct.n_chan - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ct.feat_off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    local = i - ct.feat_off[lo];
    return lo;
}

// mode 0: count run starts / ends per block.  mode 1: write them at their global rank.
template <typename T, int kMode>
__global__ void __launch_bounds__(kSegThreads)
segment_kernel(const T* __restrict__ probs, ChannelTable ct, long long total, const double* __restrict__ thr_cmp,
               const double* __restrict__ thr_raw, int* __restrict__ starts, int* __restrict__ ends,
               int* __restrict__ chans, int cap, int* __restrict__ block_counts, int n_blocks) {
    const int k = blockIdx.y;
    const double tc = thr_cmp[k], tr = thr_raw[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_warp[2][kSegThreads / 32];
    __shared__ int s_base[2];
    int* cnt_s = block_counts + (static_cast<long long>(k) * 2 + 0) * n_blocks;
    int* cnt_e = block_counts + (static_cast<long long>(k) * 2 + 1) * n_blocks;
    if (threadIdx.x == 0) {
        s_base[0] = (kMode == 1) ? cnt_s[blockIdx.x] : 0;  // after the scan pass: exclusive prefix
        s_base[1] = (kMode == 1) ? cnt_e[blockIdx.x] : 0;
    }
    __syncthreads();
    for (int it = 0; it < kSegIters; ++it) {
        const long long i = static_cast<long long>(blockIdx.x) * kSegSpan + it * kSegThreads + threadIdx.x;
        bool is_s = false, is_e = false;
        long long local = 0;
        int c = 0;
        if (i < total) {
            c = seg_channel(ct, i, local);
            if (local < ct.frames[c] && above(probs, i, tc, tr)) {
                is_s = (local == 0) || !above(probs, i - 1, tc, tr);
                is_e = (local == ct.frames[c] - 1) || !above(probs, i + 1, tc, tr);
            }
        }
        const unsigned bs = __ballot_sync(0xffffffffu, is_s), be = __ballot_sync(0xffffffffu, is_e);
        if (lane == 0) { s_warp[0][warp] = __popc(bs); s_warp[1][warp] = __popc(be); }
        __syncthreads();
        int off_s = 0, off_e = 0, tot_s = 0, tot_e = 0;
#pragma unroll
        for (int w = 0; w < kSegThreads / 32; ++w) {
            if (w < warp) { off_s += s_warp[0][w]; off_e += s_warp[1][w]; }
            tot_s += s_warp[0][w]; tot_e += s_warp[1][w];
        }
        if (kMode == 1) {
            const unsigned lt = (1u << lane) - 1u;
            if (is_s) {
                const int r = s_base[0] + off_s + __popc(bs & lt);
                if (r < cap) {
                    starts[static_cast<long long>(k) * cap + r] = static_cast<int>(local);
                    chans[static_cast<long long>(k) * cap + r] = c;
                }
            }
            if (is_e) {
                const int r = s_base[1] + off_e + __popc(be & lt);
                if (r < cap) ends[static_cast<long long>(k) * cap + r] = static_cast<int>(local);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_base[0] += tot_s; s_base[1] += tot_e; }
        __syncthreads();
    }
    if (kMode == 0 && threadIdx.x == 0) { cnt_s[blockIdx.x] = s_base[0]; cnt_e[blockIdx.x] = s_base[1]; }
}

// One block per (threshold, start/end list): in-place exclusive scan of the per-block counts.
__global__ void __launch_bounds__(256) segment_scan_kernel(int* __restrict__ block_counts, int n_blocks,
                                                           int* __restrict__ counts) {
    int* c = block_counts + static_cast<long long>(blockIdx.x) * n_blocks;
    __shared__ int s_part[256];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 256) {
        const int i = base + threadIdx.x;
        const int v = (i < n_blocks) ? c[i] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {  // Hillis-Steele inclusive scan
            const int add = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
            __syncthreads();
            s_part[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n_blocks) c[i] = s_carry + s_part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry += s_part[255];
        __syncthreads();
    }
    if (threadIdx.x == 0 && (blockIdx.x & 1) == 0) counts[blockIdx.x >> 1] = s_carry;  // #runs per threshold
}

size_t segment_scratch_ints(long long total_frames, int n_thr) {
    const long long n_blocks = (total_frames + kSegSpan - 1) / kSegSpan;
    return static_cast<size_t>(2) * n_thr * (n_blocks > 0 ? n_blocks : 1);
}

cudaError_t launch_segment_runs(const void* probs, int is_f64, const ChannelTable& ct, long long total,
                                const double* thr_cmp_d, const double* thr_raw_d, int n_thr, int* starts, int* ends,
                                int* chans, int* counts, int cap, int* block_counts, cudaStream_t stream) {
    const int n_blocks = static_cast<int>((total + kSegSpan - 1) / kSegSpan);
    if (n_blocks == 0) return cudaMemsetAsync(counts, 0, sizeof(int) * n_thr, stream);
    const dim3 grid(n_blocks, n_thr);
    if (is_f64) {
        const double* p = static_cast<const double*>(probs);
        segment_kernel<double, 0><<<grid, kSegThreads, 0, stream>>>(p, ct, total, thr_cmp_d, thr_raw_d, starts, ends, chans, cap, block_counts, n_blocks);
        segment_scan_kernel<<<2 * n_thr, 256, 0, stream>>>(block_counts, n_blocks, counts);
        segment_kernel<double, 1><<<grid, kSegThreads, 0, stream>>>(p, ct, total, thr_cmp_d, thr_raw_d, starts, ends, chans, cap, block_counts, n_blocks);
    } else {
        const float* p = static_cast<const float*>(probs);
        segment_kernel<float, 0><<<grid, kSegThreads, 0, stream>>>(p, ct, total, thr_cmp_d, thr_raw_d, starts, ends, chans, cap, block_counts, n_blocks);
        segment_scan_kernel<<<2 * n_thr, 256, 0, stream>>>(block_counts, n_blocks, counts);
        segment_kernel<float, 1><<<grid, kSegThreads, 0, stream>>>(p, ct, total, thr_cmp_d, thr_raw_d, starts, ends, chans, cap, block_counts, n_blocks);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------- K5 filtfilt
constexpr int kIirChunk = 64;
constexpr int kPadLen = 9;  // scipy: 3 * max(len(a), len(b))

struct IirCoef {
    double b0, b1, b2, a1, a2, zi0, zi1;
};

template <typename T>
__device__ __forceinline__ double odd_ext(const T* x, long long n, long long i) {  // scipy.signal._arraytools.odd_ext
    if (i < kPadLen) return 2.0 * static_cast<double>(x[0]) - static_cast<double>(x[kPadLen - i]);
    if (i >= n + kPadLen) return 2.0 * static_cast<double>(x[n - 1]) - static_cast<double>(x[n - 2 - (i - (n + kPadLen))]);
    return static_cast<double>(x[i - kPadLen]);
}

// pass 0: input = odd extension of x (forward);  pass 1: input = previous output reversed (backward).
// phase 0: end state of every chunk from a zero state;  phase 2: re-filter from the true state and write.
template <typename T, int kPass, int kPhase>
__global__ void __launch_bounds__(128)
iir_chunk_kernel(const T* __restrict__ x, long long n, IirCoef c, const double* __restrict__ yf, double* __restrict__ states,
                 double* __restrict__ out) {
    const long long m = n + 2 * kPadLen;
    const long long chunk = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i0 = chunk * kIirChunk;
    if (i0 >= m) return;
    const long long i1 = (i0 + kIirChunk < m) ? i0 + kIirChunk : m;
    double z0 = 0.0, z1 = 0.0;
    if (kPhase == 2) { z0 = states[2 * chunk]; z1 = states[2 * chunk + 1]; }
    for (long long i = i0; i < i1; ++i) {
        const double u = (kPass == 0) ? odd_ext(x, n, i) : yf[m - 1 - i];
        // scipy lfilter, direct form II transposed
        const double y = c.b0 * u + z0;
        z0 = c.b1 * u - c.a1 * y + z1;
        z1 = c.b2 * u - c.a2 * y;
        if (kPhase == 2) {
            if (kPass == 0) {
                out[i] = y;  // forward result, full padded length
            } else {
                const long long k = m - 1 - i - kPadLen;  // un-reverse and drop the padding
                if (k >= 0 && k < n) out[k] = y;
            }
        }
    }
    if (kPhase == 0) { states[2 * chunk] = z0; states[2 * chunk + 1] = z1; }
}

// phase 1: chain chunk start states: S_{j+1} = A^L S_j + e_j with A = [[-a1, 1], [-a2, 0]].
template <typename T, int kPass>
__global__ void iir_chain_kernel(const T* __restrict__ x, long long n, IirCoef c, const double* __restrict__ yf,
                                 double* __restrict__ states, long long n_chunks) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long m = n + 2 * kPadLen;
    double p00 = 1, p01 = 0, p10 = 0, p11 = 1;  // A^L
    for (int i = 0; i < kIirChunk; ++i) {
        const double q00 = -c.a1 * p00 + p10, q01 = -c.a1 * p01 + p11;
        const double q10 = -c.a2 * p00, q11 = -c.a2 * p01;
        p00 = q00; p01 = q01; p10 = q10; p11 = q11;
    }
    const double u0 = (kPass == 0) ? odd_ext(x, n, 0) : yf[m - 1];
    double s0 = c.zi0 * u0, s1 = c.zi1 * u0;  // zi * ext[0]   (scipy filtfilt)
    for (long long j = 0; j < n_chunks; ++j) {
        const double e0 = states[2 * j], e1 = states[2 * j + 1];
        states[2 * j] = s0; states[2 * j + 1] = s1;
        const long long len = (j + 1) * kIirChunk <= m ? kIirChunk : m - j * kIirChunk;
        if (len == kIirChunk) {
            const double t0 = p00 * s0 + p01 * s1 + e0, t1 = p10 * s0 + p11 * s1 + e1;
            s0 = t0; s1 = t1;
        }
    }
}

size_t filtfilt_scratch_doubles(long long n) {
    const long long m = n + 2 * kPadLen;
    const long long n_chunks = (m + kIirChunk - 1) / kIirChunk;
    return static_cast<size_t>(m + 2 * n_chunks);
}

template <typename T>
static cudaError_t filtfilt_typed(const T* x, long long n, const IirCoef& c, double* out, double* scratch, cudaStream_t stream) {
    const long long m = n + 2 * kPadLen;
    const long long n_chunks = (m + kIirChunk - 1) / kIirChunk;
    double* yf = scratch;
    double* states = scratch + m;
    const unsigned grid = static_cast<unsigned>((n_chunks + 127) / 128);
    iir_chunk_kernel<T, 0, 0><<<grid, 128, 0, stream>>>(x, n, c, yf, states, yf);
    iir_chain_kernel<T, 0><<<1, 32, 0, stream>>>(x, n, c, yf, states, n_chunks);
    iir_chunk_kernel<T, 0, 2><<<grid, 128, 0, stream>>>(x, n, c, yf, states, yf);
    iir_chunk_kernel<T, 1, 0><<<grid, 128, 0, stream>>>(x, n, c, yf, states, out);
    iir_chain_kernel<T, 1><<<1, 32, 0, stream>>>(x, n, c, yf, states, n_chunks);
    iir_chunk_kernel<T, 1, 2><<<grid, 128, 0, stream>>>(x, n, c, yf, states, out);
    return cudaGetLastError();
}

cudaError_t launch_filtfilt(const void* probs, int is_f64, long long n, const double* b, const double* a, double* out,
                            double* scratch, cudaStream_t stream) {
    IirCoef c;
    c.b0 = b[0] / a[0]; c.b1 = b[1] / a[0]; c.b2 = b[2] / a[0];
    c.a1 = a[1] / a[0]; c.a2 = a[2] / a[0];
    // scipy.signal.lfilter_zi: solve (I - companion(a).T) zi = b[1:] - a[1:] * b[0]
    const double B0 = c.b1 - c.a1 * c.b0, B1 = c.b2 - c.a2 * c.b0;
    const double det = (1.0 + c.a1) + c.a2;
    c.zi0 = (B0 + B1) / det;
    c.zi1 = ((1.0 + c.a1) * B1 - c.a2 * B0) / det;
    if (is_f64) return filtfilt_typed(static_cast<const double*>(probs), n, c, out, scratch, stream);
    return filtfilt_typed(static_cast<const float*>(probs), n, c, out, scratch, stream);
}

}  // namespace ld
