// K2a (stem) and K3 (head) of the streaming ResNetBigger evaluation, on CUDA cores in fp32.
//
//  stem:  conv1 (1->64, 3x3, pad 1, no bias) + bn1 + ReLU            reference models.py:186-191,224
//         reads the fp32 log-mel features directly (no fp16 rounding of the network input) and writes
//         the fp16 channel-chunk-planar planes the tensor-core convs consume.
//  head:  AvgPool2d(4) -> view -> bn2 -> linear1 -> bn3 -> ReLU -> linear2 -> sigmoid
//                                                                      reference models.py:229-238
//         (dropout is the identity in eval mode).
//
// A "sequence" is the concatenation of channels with zero gaps; seq rows that fall into a gap or past
// the end read as all-zero features, which is exactly InferenceDataset's right zero padding
// (reference datasets.py:85-93).
#include <cstdlib>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "ld_net.h"

namespace ld {

// Channel of sequence row s, or -1 when s lies in a gap / outside. Binary search over seq_off.
__device__ __forceinline__ int find_channel(const ChannelTable& ct, long long s, long long& local) {
    if (s < 0) return -1;
    int lo = 0, hi = ct.n_chan - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ct.seq_off[mid] <= s) lo = mid; else hi = mid - 1;
    }
    local = s - ct.seq_off[lo];
    return (local >= 0 && local < ct.frames[lo]) ? lo : -1;
}

// One warp computes 32 * kStemPx consecutive pixels of the flattened (global feature row, padded column) index space for all
// 64 channels: lane l owns pixels base + l + 32 i, so every 16-byte store instruction of the warp covers 512 contiguous
// bytes of a plane.  The three kernel-row partial sums s_ky stay apart: every stem plane (interior: s0 + s1 + s2; top edge
// of a window: s1 + s2; bottom edge: s0 + s1, 99 rows up) is a masked sum of them, so the 3 x 3 taps are multiplied once
// for all planes (9 instead of 21 multiply-adds per pixel and channel) and the feature patch is read once.
// The BatchNorm scale is folded into the fp32 weights when they are staged and the shift starts the ky = 1 partial sum, so a
// plane costs one or two additions per channel; the ReLU runs on the packed fp16 pair (max commutes with the rounding).
// kFixed: the launch has exactly the three planes of a window taller than two rows, in the planner's order -- top edge
// (ky = 1, 2), bottom edge (ky = 0, 1), interior (all) -- and the masked sums are compile-time: (s1 + s2), (s0 + s1) and
// (s0 + s1) + s2.  Any other plane set runs the generic instantiation with run-time masks.
template <int kStemPx, bool kFixed>
__global__ void __launch_bounds__(256)
stem_kernel(StemLaunch L, ChannelTable ct, const float* __restrict__ feats, long long chunk_row0, int rows, int row_lo, int rows_total) {
    __shared__ __align__(16) float s_w[64 * 9];
    __shared__ __align__(16) float s_shift[64];
    for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) s_w[i] = L.w[i] * L.scale[i / 9];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_shift[i] = L.shift[i];
    __syncthreads();

    const int wp = L.wp;   // column 0 = the zero pad shared with the row before (ld_types.h); wp = W + 2: also column W + 1
    const long long n_pix = static_cast<long long>(rows_total) * wp;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long base = warp * (32 * kStemPx);
    if (base >= n_pix) return;
    constexpr int kJobs = kFixed ? 3 : kMaxStemJobs;

    // per pixel: the 3 x 3 feature patch (real column = padded column - 1) and where it lands in each plane
    float x[kStemPx][3][3];
    long long dst[kStemPx][kJobs];   // element offset of the pixel inside plane j
    bool ok[kStemPx][kJobs];         // ... and whether the plane has that row
    bool pad[kStemPx];
#pragma unroll
    for (int i = 0; i < kStemPx; ++i) {
        const long long p = base + lane + 32 * i;
        const bool live = p < n_pix;
        const int rr = live ? static_cast<int>(p / wp) : 0;
        const int pc = live ? static_cast<int>(p - static_cast<long long>(rr) * wp) : 0;
        const int G = row_lo + rr;                       // chunk-relative global row of the centre tap
        pad[i] = pc == 0 || pc > L.W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            long long local = 0;
            const int c = live ? find_channel(ct, chunk_row0 + G + ky - 1, local) : -1;
            const float* frow = (c >= 0) ? feats + (ct.feat_off[c] + local) * L.W : nullptr;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int cc = pc - 1 + kx - 1;
                x[i][ky][kx] = (frow != nullptr && cc >= 0 && cc < L.W) ? __ldg(frow + cc) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < kJobs; ++j) {
            const int r = G - (j < L.n_jobs ? L.jobs[j].row_shift : 0);
            ok[i][j] = live && j < L.n_jobs && r >= 0 && r < rows;
            dst[i][j] = (static_cast<long long>(r) * wp + pc) * 8;
        }
    }
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const __half2 zero = __floats2half2_rn(0.f, 0.f);
    auto pack_relu = [&](const float (&o)[8]) {
        uint4 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __hmax2(__floats2half2_rn(o[2 * e], o[2 * e + 1]), zero);
        return ov;
    };
#pragma unroll 1
    for (int kc = 0; kc < 8; ++kc) {
        float part[3][kStemPx][8];
        // the 72 weights and 8 shifts of this channel chunk: warp-uniform 16-byte shared-memory reads
        float w[8][9], sh[8];
        {
            const float4* w4 = reinterpret_cast<const float4*>(s_w + kc * 72);
            float* wf = &w[0][0];
#pragma unroll
            for (int q = 0; q < 18; ++q) {
                const float4 v = w4[q];
                wf[4 * q] = v.x; wf[4 * q + 1] = v.y; wf[4 * q + 2] = v.z; wf[4 * q + 3] = v.w;
            }
            const float4 s0 = *reinterpret_cast<const float4*>(s_shift + kc * 8), s1 = *reinterpret_cast<const float4*>(s_shift + kc * 8 + 4);
            sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
            for (int i = 0; i < kStemPx; ++i)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    float a = ky == 1 ? sh[e] : 0.f;   // the ky = 1 partial sum carries the shift
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) a = fmaf(w[e][ky * 3 + kx], x[i][ky][kx], a);
                    part[ky][i][e] = a;
                }
#pragma unroll
        for (int i = 0; i < kStemPx; ++i) {
            if constexpr (kFixed) {
                float o12[8], o01[8], o012[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o12[e] = part[1][i][e] + part[2][i][e];
                    o01[e] = part[0][i][e] + part[1][i][e];
                    o012[e] = o01[e] + part[2][i][e];
                }
                const uint4 v0 = pad[i] ? zero4 : pack_relu(o12), v1 = pad[i] ? zero4 : pack_relu(o01), v2 = pad[i] ? zero4 : pack_relu(o012);
                if (ok[i][0]) *reinterpret_cast<uint4*>(L.jobs[0].out + dst[i][0] + kc * L.jobs[0].kc_stride) = v0;
                if (ok[i][1]) *reinterpret_cast<uint4*>(L.jobs[1].out + dst[i][1] + kc * L.jobs[1].kc_stride) = v1;
                if (ok[i][2]) *reinterpret_cast<uint4*>(L.jobs[2].out + dst[i][2] + kc * L.jobs[2].kc_stride) = v2;
            } else {
#pragma unroll
                for (int j = 0; j < kJobs; ++j) {
                    if (j >= L.n_jobs) continue;
                    const StemJob job = L.jobs[j];
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float a = (job.mask & 2) ? part[1][i][e] : sh[e];
                        if (job.mask & 1) a += part[0][i][e];
                        if (job.mask & 4) a += part[2][i][e];
                        o[e] = a;
                    }
                    const uint4 ov = pad[i] ? zero4 : pack_relu(o);
                    if (ok[i][j]) *reinterpret_cast<uint4*>(job.out + dst[i][j] + kc * job.kc_stride) = ov;
                }
            }
        }
    }
}

// 128 windows per CTA, 384 threads: thread (g, w) first pools group g (4 rows x 4 columns) of window w -- its 4 x 2 x 4 (x 2 in
// split precision) 16-byte loads are independent and all in flight together -- into shared memory, then threads 0..127 run the
// dense part of one window each.  (One thread per window walking its 96 loads in sequence left the kernel latency-bound.)
constexpr int kHeadWin = 128;
__global__ void __launch_bounds__(3 * kHeadWin)
head_kernel(HeadLaunch L, ChannelTable ct, float* __restrict__ probs, long long chunk_row0, int nb) {
    extern __shared__ float s_par[];
    // layout: bn2 scale[F] shift[F] | W1[32*F] b1[32] | bn3 scale[32] shift[32] | w2[32] b2 | pooled[kHeadWin][F + 1]
    const int F = L.n_feat;
    const int n_par = 2 * F + 32 * F + 32 + 64 + 32 + 1;
    float* s_pool = s_par + ((n_par + 3) & ~3);
    for (int i = threadIdx.x; i < n_par; i += blockDim.x) s_par[i] = L.params[i];
    const float* bn2_s = s_par;
    const float* bn2_b = bn2_s + F;
    const float* w1 = bn2_b + F;
    const float* b1 = w1 + 32 * F;
    const float* bn3_s = b1 + 32;
    const float* bn3_b = bn3_s + 32;
    const float* w2 = bn3_b + 32;

    const int C = L.C, G = L.groups;
    const int w = threadIdx.x % kHeadWin, g = threadIdx.x / kHeadWin;
    const int b = blockIdx.x * kHeadWin + w;
    // AvgPool2d(4): rows 4g..4g+3, real cols 0..3; feature index = c * groups + g   (models.py:229-230)
    if (b < nb && g < G) {
        for (int kc = 0; kc < C / 8; ++kc) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int part = 0; part <= L.split; ++part) {   // [hi | lo] planes: the rounding residual sits C/8 chunks further
                uint4 v[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const HeadRow hr = L.rows[4 * g + r];
                    const __half* base = hr.plane + ((static_cast<long long>(b) + hr.row_shift) * L.wp + 1) * 8 +
                                         static_cast<long long>(kc + part * (C / 8)) * hr.kc_stride;
#pragma unroll
                    for (int cpx = 0; cpx < 4; ++cpx) v[r][cpx] = *reinterpret_cast<const uint4*>(base + cpx * 8);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cpx = 0; cpx < 4; ++cpx) {
                        const __half2* vh = reinterpret_cast<const __half2*>(&v[r][cpx]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(vh[e]);
                            acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
                        }
                    }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) s_pool[w * (F + 1) + (kc * 8 + e) * G + g] = acc[e];
        }
    }
    __syncthreads();
    if (g != 0 || b >= nb) return;
    long long local = 0;
    const int chan = find_channel(ct, chunk_row0 + b, local);
    if (chan < 0) return;  // gap row: no window starts here
    float x[kMaxHeadFeat];
    for (int i = 0; i < F; ++i) x[i] = fmaf(s_pool[w * (F + 1) + i] * (1.f / 16.f), bn2_s[i], bn2_b[i]);
    float z = w2[32];
    for (int o = 0; o < 32; ++o) {
        float h = b1[o];
        for (int i = 0; i < F; ++i) h = fmaf(w1[o * F + i], x[i], h);
        h = fmaxf(fmaf(h, bn3_s[o], bn3_b[o]), 0.f);
        z = fmaf(w2[o], h, z);
    }
    probs[ct.feat_off[chan] + local] = 1.f / (1.f + expf(-z));
}

cudaError_t launch_stem(const StemLaunch& L, const ChannelTable& ct, const float* feats, long long chunk_row0,
                        int rows, cudaStream_t stream) {
    // global rows whose taps reach some plane row: [min row_shift, rows + max row_shift)
    int lo = 0, hi = 0;
    for (int j = 0; j < L.n_jobs; ++j) { lo = j == 0 ? L.jobs[j].row_shift : (L.jobs[j].row_shift < lo ? L.jobs[j].row_shift : lo); hi = L.jobs[j].row_shift > hi ? L.jobs[j].row_shift : hi; }
    const int rows_total = rows + hi - lo;
    // pixels per thread: 2 keeps the kernel under 128 registers (two CTAs per SM), 4 reuses every weight fetch twice as often
    static const int px_env = []() { const char* v = std::getenv("LD_STEM_PX"); return (v && std::atoi(v) == 4) ? 4 : 2; }();
    const bool fixed0 = L.n_jobs == 3 && L.jobs[0].mask == 6 && L.jobs[1].mask == 3 && L.jobs[2].mask == 7;
    const int px = fixed0 ? px_env : 2;
    const long long threads = (static_cast<long long>(rows_total) * L.wp + px - 1) / px;   // 32 * px pixels per warp
    const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
    const bool fixed = L.n_jobs == 3 && L.jobs[0].mask == 6 && L.jobs[1].mask == 3 && L.jobs[2].mask == 7;
    if (!fixed) stem_kernel<2, false><<<grid, 256, 0, stream>>>(L, ct, feats, chunk_row0, rows, lo, rows_total);
    else if (px == 4) stem_kernel<4, true><<<grid, 256, 0, stream>>>(L, ct, feats, chunk_row0, rows, lo, rows_total);
    else stem_kernel<2, true><<<grid, 256, 0, stream>>>(L, ct, feats, chunk_row0, rows, lo, rows_total);
    return cudaGetLastError();
}

cudaError_t launch_head(const HeadLaunch& L, const ChannelTable& ct, float* probs, long long chunk_row0, int nb,
                        cudaStream_t stream) {
    const int F = L.n_feat;
    if (L.groups > 3) return cudaErrorInvalidValue;   // the kernel runs three pooling groups per window (13 rows -> 3 x 4)
    const int n_par = 2 * F + 32 * F + 32 + 64 + 32 + 1;
    const size_t smem = sizeof(float) * (((n_par + 3) & ~3) + kHeadWin * (F + 1));
    static PerDeviceOnce attr_set;
    if (!attr_set.flag()) {
        cudaError_t e = cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        attr_set.flag() = true;
    }
    head_kernel<<<(nb + kHeadWin - 1) / kHeadWin, 3 * kHeadWin, smem, stream>>>(L, ct, probs, chunk_row0, nb);
    return cudaGetLastError();
}

}  // namespace ld
