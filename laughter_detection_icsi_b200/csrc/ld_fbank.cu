// K1: fused log-mel filterbank front-end.  int16 PCM in, fp32 (T, F) log-mel out, one pass:
// reflect-pad framing (snip_edges=False) -> DC removal -> pre-emphasis 0.97 -> Povey window (400) ->
// zero-pad to 512 -> real FFT -> |X|^2 -> mel filterbank -> log(max(., eps)).
// Replaces lhotse Fbank(FbankConfig(num_filters=44, frame_shift=0.01)) = Wav2LogFilterBank as called at
// reference load_data.py:47-49 / utils/utils.py:25 (arithmetic restated in oracle/fbank_oracle.py).
//
// 16 threads own one frame: the 512-point real FFT is a 256-point complex FFT of (even, odd) sample
// pairs, done as two register-resident radix-16 passes with one shared-memory transpose between them,
// followed by the real-input untangling step.  256 threads = 16 frames per CTA share one staged,
// pre-processed span of 15*160+400 samples.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ld_net.h"

namespace ld {

constexpr int kFrameLen = 400, kFrameShift = 160, kFftHalf = 256, kBins = 257;
constexpr int kFramesPerCta = 16;
constexpr int kGroupsPerCta = 8;   // 128 frames per CTA: table staging amortised over 8 groups
constexpr int kSpan = (kFramesPerCta - 1) * kFrameShift + kFrameLen;  // 2800 samples staged per CTA
constexpr int kMaxFilters = 64;
constexpr int kMaxMelWeights = 1024;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place 4-point DFT (forward, W4 = -i).
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
    a0 = make_float2(s02.x + s13.x, s02.y + s13.y);
    a2 = make_float2(s02.x - s13.x, s02.y - s13.y);
    a1 = make_float2(d02.x + d13.y, d02.y - d13.x);  // d02 - i*d13
    a3 = make_float2(d02.x - d13.y, d02.y + d13.x);  // d02 + i*d13
}

// 16-point forward DFT, natural order in and out: n = n1 + 4*n2, k = k2 + 4*k1.
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
    // W16^j = exp(-2*pi*i*j/16)
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
    const float2 w[10] = {{1.f, 0.f}, {c1, -s1}, {r2, -r2}, {s1, -c1}, {0.f, -1.f},
                          {-s1, -c1}, {-r2, -r2}, {-c1, -s1}, {-1.f, 0.f}, {-c1, s1}};
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4(a[n1], a[n1 + 4], a[n1 + 8], a[n1 + 12]);  // over n2 -> k2 at index n1+4*k2
#pragma unroll
    for (int n1 = 1; n1 < 4; ++n1)
#pragma unroll
        for (int k2 = 1; k2 < 4; ++k2) a[n1 + 4 * k2] = cmul(a[n1 + 4 * k2], w[n1 * k2]);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) dft4(a[4 * k2], a[4 * k2 + 1], a[4 * k2 + 2], a[4 * k2 + 3]);  // over n1 -> k1 at 4*k2+k1
    // a[4*k2 + k1] holds X[k2 + 4*k1]: transpose the 4x4 index to natural order
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = k1 + 1; k2 < 4; ++k2) {
            const float2 t = a[4 * k2 + k1];
            a[4 * k2 + k1] = a[4 * k1 + k2];
            a[4 * k1 + k2] = t;
        }
}

constexpr int kSpanPad = kSpan + 16;    // pass 1 reads 416 samples per frame (the window is zero past 400)
constexpr int kWinPad = kFrameLen + 16;
constexpr int kTrFloats = 2 * 16 * 17 + 16;   // floats of one frame's transpose buffer (16 x 17 complex) + 16: the two frames of
                                              // a warp sit 16 banks apart, so their power-spectrum accesses do not collide
constexpr int kStageOff = 264;           // floats: the frame's 44 log-mel outputs wait here (behind its 257 power values)

struct FbankSmem {
    float samples[kSpanPad];                   // pre-processed (utterance mode) or raw (frame mode) samples
    float window[kWinPad];                     // Povey window, zero-padded to 416
    float2 tw512[kBins];
    __align__(16) float tr[kFramesPerCta][kTrFloats];   // per frame: transpose buffer between the two radix-16 passes, then the power
                                               // spectrum (257 floats) and the staged outputs (aliased: the passes are done)
    __align__(16) float melw[kMaxMelWeights];
    int mel_lo[kMaxFilters], mel_len[kMaxFilters], mel_off[kMaxFilters];
};

__device__ __forceinline__ float pcm_to_float(int v) { return static_cast<float>(v) * (1.f / 32768.f); }

// kMinBlocks: CTAs per SM the register allocation aims for (2: 128 registers, 3: 80, 4: 64 with 68 bytes of spills)
template <bool kPerFrame, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
fbank_kernel(const int16_t* __restrict__ pcm, long long n_samples, long long n_frames,
             const unsigned long long* __restrict__ sum_biased, FbankMel mel, const float* __restrict__ tables,
             float* __restrict__ feats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FbankSmem& S = *reinterpret_cast<FbankSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int fl = tid >> 4;  // frame within the group
    const int t = tid & 15;   // thread within the frame
    const unsigned gmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
    const int lane_base = threadIdx.x & 16;   // first lane of this frame's 16 threads inside the warp

    // ---- tables: once per CTA (128 frames) --------------------------------------------------------------
    for (int i = tid; i < kWinPad; i += 256) S.window[i] = i < kFrameLen ? tables[i] : 0.f;
    for (int i = tid; i < kBins; i += 256) S.tw512[i] = make_float2(tables[912 + 2 * i], tables[912 + 2 * i + 1]);
    for (int i = kSpan + tid; i < kSpanPad; i += 256) S.samples[i] = 0.f;
    const int F = mel.n_filters;
    for (int i = tid; i < F; i += 256) { S.mel_lo[i] = mel.lo[i]; S.mel_len[i] = mel.len[i]; S.mel_off[i] = mel.off[i]; }
    {
        const int total_w = mel.off[F - 1] + 4 * mel.len[F - 1];   // run lengths are in 4-bin vectors
        for (int i = tid; i < total_w; i += 256) S.melw[i] = mel.weights[i];
    }
    // inter-pass twiddles W256^(t * k1): fixed per thread, kept in registers
    float2 tw[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        const int j = (t * k1) & 255;
        tw[k1] = make_float2(__ldg(tables + 400 + 2 * j), __ldg(tables + 400 + 2 * j + 1));
    }
    float mu = 0.f;
    if (!kPerFrame) {
        // exact integer sum -> mean of the [-1,1) floats (Wav2Win: x - mean(x) over the whole recording)
        const double s = static_cast<double>(static_cast<long long>(*sum_biased) - 32768ll * n_samples);
        mu = static_cast<float>(s / (32768.0 * static_cast<double>(n_samples)));
    }

    for (int grp = 0; grp < kGroupsPerCta; ++grp) {
        const long long f0 = (static_cast<long long>(blockIdx.x) * kGroupsPerCta + grp) * kFramesPerCta;
        if (f0 >= n_frames) break;
        // ---- stage this group's sample span: 15 * 160 + 400 samples (+ 16 read under the zero tail of the window) ----
        const long long q0 = f0 * kFrameShift - (kFrameLen - kFrameShift) / 2;
        const bool fast = q0 >= 1 && q0 + kSpanPad <= n_samples && ((reinterpret_cast<uintptr_t>(pcm + q0) & 15u) == 0);
        if (fast) {
            // no reflection inside the span: 16-byte vector loads, every sample read once (+ one neighbour per vector)
            for (int v = tid; v < kSpanPad / 8; v += 256) {
                const int16_t* src = pcm + q0 + 8 * v;
                const uint4 d = __ldg(reinterpret_cast<const uint4*>(src));
                const int prev = __ldg(src - 1);
                const uint32_t w4[4] = {d.x, d.y, d.z, d.w};
                float x[9];
                x[0] = pcm_to_float(prev);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    x[1 + 2 * e] = pcm_to_float(static_cast<int16_t>(w4[e] & 0xFFFFu));
                    x[2 + 2 * e] = pcm_to_float(static_cast<int16_t>(w4[e] >> 16));
                }
                float y[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (kPerFrame) {
                        y[e] = x[e + 1];
                    } else {
                        const float a = __fsub_rn(x[e + 1], mu), b = __fsub_rn(x[e], mu);
                        y[e] = __fsub_rn(a, __fmul_rn(0.97f, b));   // pre-emphasis on the whole signal, then framing
                    }
                }
                float4* dst = reinterpret_cast<float4*>(S.samples + 8 * v);
                dst[0] = make_float4(y[0], y[1], y[2], y[3]);
                dst[1] = make_float4(y[4], y[5], y[6], y[7]);
            }
        } else {
            for (int i = tid; i < kSpanPad; i += 256) {
                long long q = q0 + i;
                long long n = q < 0 ? -q - 1 : (q >= n_samples ? 2 * n_samples - 1 - q : q);  // flip-padding
                n = n < 0 ? 0 : (n >= n_samples ? n_samples - 1 : n);
                const float x = pcm_to_float(pcm[n]);
                if (kPerFrame) {
                    S.samples[i] = x;
                } else {
                    const float xp = pcm_to_float(pcm[n > 0 ? n - 1 : 0]);
                    const float a = __fsub_rn(x, mu), b = __fsub_rn(xp, mu);
                    S.samples[i] = __fsub_rn(a, __fmul_rn(0.97f, b));
                }
            }
        }
        __syncthreads();

        const long long frame = f0 + fl;
        if (frame < n_frames) {   // whole 16-thread groups skip together; only __syncwarp(gmask) inside
            const float* xs = S.samples + fl * kFrameShift;
            // ---- pass 1: radix-16 over m of z[t + 16 m] (z[n] = x[2n] + i x[2n+1], windowed), then twiddle W256^(t*k1) ----
            float2 a[16];
            if (!kPerFrame) {
                const float2* xs2 = reinterpret_cast<const float2*>(xs) + t;
                const float2* w2 = reinterpret_cast<const float2*>(S.window) + t;
#pragma unroll
                for (int m = 0; m < 13; ++m) {
                    const float2 x = xs2[16 * m], w = w2[16 * m];
                    a[m] = make_float2(x.x * w.x, x.y * w.y);
                }
            } else {
                float acc = 0.f;
                for (int i = t; i < kFrameLen; i += 16) acc += xs[i];
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
                const float fmean = acc * (1.f / kFrameLen);
#pragma unroll
                for (int m = 0; m < 13; ++m) {
                    const int i = 2 * (t + 16 * m);
                    const float xm = __fsub_rn(xs[i > 0 ? i - 1 : 0], fmean), x0 = __fsub_rn(xs[i], fmean), x1 = __fsub_rn(xs[i + 1], fmean);
                    a[m] = make_float2(__fsub_rn(x0, __fmul_rn(0.97f, xm)) * S.window[i], __fsub_rn(x1, __fmul_rn(0.97f, x0)) * S.window[i + 1]);
                }
            }
            a[13] = a[14] = a[15] = make_float2(0.f, 0.f);   // samples 416.. of the zero-padded frame
            dft16(a);
            float2* tr = reinterpret_cast<float2*>(S.tr[fl]);
            tr[t] = a[0];
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) tr[k1 * 17 + t] = cmul(a[k1], tw[k1]);
            __syncwarp(gmask);
            // ---- pass 2: thread k1 = t, radix-16 over t' -> Z[k1 + 16 k2] -------------------------------------------
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = tr[t * 17 + j];
            dft16(a);
            __syncwarp(gmask);  // every thread of the frame has read its row of tr: the buffer now holds the power spectrum
            // ---- real-input untangle + power spectrum.  a[k2] = Z[t + 16 k2]; bin k = t + 16 k2 also needs Z[256 - k], which
            //      sits in register 15 - k2 of thread 16 - t (thread 0: its own register (16 - k2) & 15) ----------------------
            float* P = S.tr[fl];
            const int partner = lane_base | ((16 - t) & 15);
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                const float2 give = a[15 - k2];
                float2 zn = make_float2(__shfl_sync(gmask, give.x, partner), __shfl_sync(gmask, give.y, partner));
                if (t == 0) zn = a[(16 - k2) & 15];
                const float2 zk = a[k2];
                const int k = t + 16 * k2;
                const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));   // (Z[k] + conj Z[N-k]) / 2
                const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y));   // (Z[k] - conj Z[N-k]) / 2
                const float2 o = make_float2(d.y, -d.x);                                    // d / i
                const float2 wo = cmul(S.tw512[k], o);
                const float re = e.x + wo.x, im = e.y + wo.y;
                P[k] = re * re + im * im;
            }
            if (t == 0) {   // bin 256: X[256] = Re Z[0] - Im Z[0]; bins 257..259 only ever meet zero weights
                const float x = a[0].x - a[0].y;
                P[kFftHalf] = x * x;
                P[257] = P[258] = P[259] = 0.f;
            }
            __syncwarp(gmask);
            // ---- mel filterbank + log, staged behind the power spectrum ---------------------------------------------------
            for (int k = t; k < F; k += 16) {
                const float4* w4 = reinterpret_cast<const float4*>(S.melw + S.mel_off[k]);
                const float4* p4 = reinterpret_cast<const float4*>(P + S.mel_lo[k]);
                const int n4 = S.mel_len[k];
                float acc = 0.f;
                for (int j = 0; j < n4; ++j) {   // 16-byte vectors of power values and weights (runs are padded to 4 bins)
                    const float4 pv = p4[j], wv = w4[j];
                    acc = fmaf(pv.x, wv.x, acc); acc = fmaf(pv.y, wv.y, acc); acc = fmaf(pv.z, wv.z, acc); acc = fmaf(pv.w, wv.w, acc);
                }
                P[kStageOff + k] = logf(fmaxf(acc, 1.1920928955078125e-07f));  // torch.finfo(float32).eps
            }
        }
        __syncthreads();
        // ---- the group's 16 x F outputs are contiguous in the feature matrix: coalesced rows -------------------------------
        {
            const long long left = n_frames - f0;
            const int n_out = static_cast<int>(left < kFramesPerCta ? left : kFramesPerCta) * F;
            float* dst = feats + f0 * F;
            for (int i = tid; i < n_out; i += 256) {
                const int fr = i / F;
                dst[i] = S.tr[fr][kStageOff + (i - fr * F)];
            }
        }
        // (the next group's staging writes S.samples only; its pass 1 follows a __syncthreads)
    }
}

__global__ void __launch_bounds__(256) pcm_sum_kernel(const int16_t* __restrict__ pcm, long long n,
                                                      unsigned long long* __restrict__ out) {
    long long acc = 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x, gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    // head up to the first 16-byte boundary, 8 samples per load in the body, scalar tail
    long long head = ((16 - (reinterpret_cast<uintptr_t>(pcm) & 15u)) & 15u) / 2;
    if (head > n) head = n;
    const long long n_vec = (n - head) / 8;
    for (long long i = gid; i < head; i += stride) acc += pcm[i];
    const uint4* pv = reinterpret_cast<const uint4*>(pcm + head);
    for (long long i = gid; i < n_vec; i += stride) {
        const uint4 d = __ldg(pv + i);
        const uint32_t w4[4] = {d.x, d.y, d.z, d.w};
        int part = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) part += static_cast<int16_t>(w4[e] & 0xFFFFu) + static_cast<int16_t>(w4[e] >> 16);
        acc += part;
    }
    for (long long i = head + 8 * n_vec + gid; i < n; i += stride) acc += pcm[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ long long s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int i = 0; i < 8; ++i) tot += s_part[i];
        atomicAdd(out, static_cast<unsigned long long>(tot));  // two's complement wrap-around is exact
    }
}

// `sum_biased` receives sum(pcm) + 32768*n (so that the kernel's unbiasing is uniform); must be zeroed
// by the caller before the launch and initialised with the bias here.
cudaError_t launch_pcm_sum(const int16_t* pcm, long long n, unsigned long long* sum_biased, cudaStream_t stream) {
    const unsigned long long bias = 32768ull * static_cast<unsigned long long>(n);
    cudaError_t e = cudaMemcpyAsync(sum_biased, &bias, sizeof(bias), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    long long blocks = (n + 256 * 64 - 1) / (256 * 64);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    pcm_sum_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pcm, n, sum_biased);
    return cudaGetLastError();
}

cudaError_t launch_fbank(const int16_t* pcm, long long n_samples, long long n_frames, const unsigned long long* sum_biased,
                         int per_frame, const FbankMel& mel, const float* tables, float* feats, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    static const int occ = []() { const char* v = std::getenv("LD_FBANK_OCC"); const int o = v ? std::atoi(v) : 3; return o < 2 ? 2 : (o > 4 ? 4 : o); }();
    static PerDeviceOnce attr_set;
    const int smem = static_cast<int>(sizeof(FbankSmem));
    if (!attr_set.flag()) {
        cudaError_t e = cudaSuccess;
#define LD_FBANK_ATTR(pf, mb) \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fbank_kernel<pf, mb>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
        LD_FBANK_ATTR(false, 2); LD_FBANK_ATTR(false, 3); LD_FBANK_ATTR(false, 4);
        LD_FBANK_ATTR(true, 2); LD_FBANK_ATTR(true, 3); LD_FBANK_ATTR(true, 4);
#undef LD_FBANK_ATTR
        if (e != cudaSuccess) return e;
        attr_set.flag() = true;
    }
    const unsigned grid = static_cast<unsigned>((n_frames + kFramesPerCta * kGroupsPerCta - 1) / (kFramesPerCta * kGroupsPerCta));
#define LD_FBANK_LAUNCH(pf, mb) \
    fbank_kernel<pf, mb><<<grid, 256, smem, stream>>>(pcm, n_samples, n_frames, sum_biased, mel, tables, feats)
    if (per_frame) {
        if (occ == 2) LD_FBANK_LAUNCH(true, 2); else if (occ == 3) LD_FBANK_LAUNCH(true, 3); else LD_FBANK_LAUNCH(true, 4);
    } else {
        if (occ == 2) LD_FBANK_LAUNCH(false, 2); else if (occ == 3) LD_FBANK_LAUNCH(false, 3); else LD_FBANK_LAUNCH(false, 4);
    }
#undef LD_FBANK_LAUNCH
    return cudaGetLastError();
}

// window (400) | W256^j (256 x (cos, -sin)) | W512^k (257 x (cos, -sin)), computed in double.
void fbank_host_tables(float* out) {
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < kFrameLen; ++i)
        out[i] = static_cast<float>(std::pow(0.5 - 0.5 * std::cos(2.0 * pi * i / (kFrameLen - 1)), 0.85));
    for (int j = 0; j < kFftHalf; ++j) {
        out[400 + 2 * j] = static_cast<float>(std::cos(2.0 * pi * j / 256.0));
        out[400 + 2 * j + 1] = static_cast<float>(-std::sin(2.0 * pi * j / 256.0));
    }
    for (int k = 0; k < kBins; ++k) {
        out[912 + 2 * k] = static_cast<float>(std::cos(2.0 * pi * k / 512.0));
        out[912 + 2 * k + 1] = static_cast<float>(-std::sin(2.0 * pi * k / 512.0));
    }
}

}  // namespace ld
