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

struct FbankSmem {
    float samples[kSpan];                      // pre-processed (utterance mode) or raw (frame mode) samples
    float window[kFrameLen];
    float2 tw256[kFftHalf];
    float2 tw512[kBins];
    float2 tr[kFramesPerCta][16 * 17];          // per frame: transpose buffer, then the FFT output Z
    float pw[kFramesPerCta][kBins + 3];         // per frame: power spectrum
    float melw[kMaxMelWeights];
    int mel_lo[kMaxFilters], mel_len[kMaxFilters], mel_off[kMaxFilters];
};

template <bool kPerFrame>
__global__ void __launch_bounds__(256)
fbank_kernel(const int16_t* __restrict__ pcm, long long n_samples, long long n_frames,
             const unsigned long long* __restrict__ sum_biased, FbankMel mel, const float* __restrict__ tables,
             float* __restrict__ feats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FbankSmem& S = *reinterpret_cast<FbankSmem*>(smem_raw);
    const int tid = threadIdx.x;

    // ---- stage tables and this CTA's sample span ------------------------------------------------------
    for (int i = tid; i < kFrameLen; i += 256) S.window[i] = tables[i];
    for (int i = tid; i < kFftHalf; i += 256) S.tw256[i] = make_float2(tables[400 + 2 * i], tables[400 + 2 * i + 1]);
    for (int i = tid; i < kBins; i += 256) S.tw512[i] = make_float2(tables[912 + 2 * i], tables[912 + 2 * i + 1]);
    const int F = mel.n_filters;
    for (int i = tid; i < F; i += 256) { S.mel_lo[i] = mel.lo[i]; S.mel_len[i] = mel.len[i]; S.mel_off[i] = mel.off[i]; }
    {
        const int total_w = mel.off[F - 1] + mel.len[F - 1];
        for (int i = tid; i < total_w; i += 256) S.melw[i] = mel.weights[i];
    }
    // the tables are staged once per CTA; the CTA then walks kGroupsPerCta consecutive groups of 16 frames
    for (int grp = 0; grp < kGroupsPerCta; ++grp) {
    const long long f0 = (static_cast<long long>(blockIdx.x) * kGroupsPerCta + grp) * kFramesPerCta;
    if (f0 >= n_frames) break;
    if (grp > 0) __syncthreads();   // everyone is done with the previous group's samples / spectra
    float mu = 0.f;
    if (!kPerFrame) {
        // exact integer sum -> mean of the [-1,1) floats (Wav2Win: x - mean(x) over the whole recording)
        const double s = static_cast<double>(static_cast<long long>(*sum_biased) - 32768ll * n_samples);
        mu = static_cast<float>(s / (32768.0 * static_cast<double>(n_samples)));
    }
    const long long q0 = f0 * kFrameShift - (kFrameLen - kFrameShift) / 2;
    for (int i = tid; i < kSpan; i += 256) {
        long long q = q0 + i;
        long long n = q < 0 ? -q - 1 : (q >= n_samples ? 2 * n_samples - 1 - q : q);  // flip-padding
        n = n < 0 ? 0 : (n >= n_samples ? n_samples - 1 : n);
        const float x = static_cast<float>(pcm[n]) * (1.f / 32768.f);
        if (kPerFrame) {
            S.samples[i] = x;
        } else {
            const float xp = static_cast<float>(pcm[n > 0 ? n - 1 : 0]) * (1.f / 32768.f);
            const float a = __fsub_rn(x, mu), b = __fsub_rn(xp, mu);
            S.samples[i] = __fsub_rn(a, __fmul_rn(0.97f, b));  // pre-emphasis on the whole signal, then padding
        }
    }
    __syncthreads();

    const int fl = tid >> 4;  // frame within the CTA
    const int t = tid & 15;   // thread within the frame
    const long long frame = f0 + fl;
    if (frame >= n_frames) continue;  // whole 16-thread groups skip together; only __syncwarp below
    const unsigned gmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
    const float* xs = S.samples + fl * kFrameShift;

    float fmean = 0.f;
    if (kPerFrame) {
        float acc = 0.f;
        for (int i = t; i < kFrameLen; i += 16) acc += xs[i];
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
        fmean = acc * (1.f / kFrameLen);
    }
    auto windowed = [&](int i) -> float {
        if (i >= kFrameLen) return 0.f;
        float v;
        if (kPerFrame) {
            const float a = __fsub_rn(xs[i], fmean), b = __fsub_rn(xs[i > 0 ? i - 1 : 0], fmean);
            v = __fsub_rn(a, __fmul_rn(0.97f, b));
        } else {
            v = xs[i];
        }
        return v * S.window[i];
    };

    // ---- pass 1: radix-16 over m of z[t + 16 m], then twiddle W256^(t*k1) ------------------------------
    float2 a[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + 16 * m;
        a[m] = make_float2(windowed(2 * n), windowed(2 * n + 1));
    }
    dft16(a);
    float2* tr = S.tr[fl];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tr[k1 * 17 + t] = cmul(a[k1], S.tw256[t * k1]);
    __syncwarp(gmask);
    // ---- pass 2: thread k1 = t, radix-16 over t' -> Z[k1 + 16 k2] ----------------------------------------
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = tr[t * 17 + j];
    dft16(a);
    __syncwarp(gmask);  // every thread of the frame has read its row of tr: reuse it for Z
    float2* Z = tr;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) Z[t + 16 * k2] = a[k2];
    __syncwarp(gmask);
    // ---- real-input untangle + power spectrum ------------------------------------------------------------
    float* P = S.pw[fl];
    for (int k = t; k < kBins; k += 16) {
        const float2 zk = Z[k & 255];
        const float2 zn = Z[(kFftHalf - k) & 255];
        const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));   // (Z[k] + conj Z[N-k]) / 2
        const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y));   // (Z[k] - conj Z[N-k]) / 2
        const float2 o = make_float2(d.y, -d.x);                                    // d / i
        const float2 wo = cmul(S.tw512[k], o);
        const float re = e.x + wo.x, im = e.y + wo.y;
        P[k] = re * re + im * im;
    }
    __syncwarp(gmask);
    // ---- mel filterbank + log -----------------------------------------------------------------------------
    for (int k = t; k < F; k += 16) {
        const float* w = S.melw + S.mel_off[k];
        const float* p = P + S.mel_lo[k];
        float acc = 0.f;
        for (int j = 0; j < S.mel_len[k]; ++j) acc = fmaf(p[j], w[j], acc);
        feats[frame * F + k] = logf(fmaxf(acc, 1.1920928955078125e-07f));  // torch.finfo(float32).eps
    }
    }  // groups
}

__global__ void __launch_bounds__(256) pcm_sum_kernel(const int16_t* __restrict__ pcm, long long n,
                                                      unsigned long long* __restrict__ out) {
    long long acc = 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        acc += pcm[i];
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
    long long blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    pcm_sum_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pcm, n, sum_biased);
    return cudaGetLastError();
}

cudaError_t launch_fbank(const int16_t* pcm, long long n_samples, long long n_frames, const unsigned long long* sum_biased,
                         int per_frame, const FbankMel& mel, const float* tables, float* feats, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    static PerDeviceOnce attr_set;
    if (!attr_set.flag()) {
        cudaError_t e = cudaFuncSetAttribute(fbank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(FbankSmem)));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fbank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(FbankSmem)));
        if (e != cudaSuccess) return e;
        attr_set.flag() = true;
    }
    const unsigned grid = static_cast<unsigned>((n_frames + kFramesPerCta * kGroupsPerCta - 1) / (kFramesPerCta * kGroupsPerCta));
    if (per_frame)
        fbank_kernel<true><<<grid, 256, sizeof(FbankSmem), stream>>>(pcm, n_samples, n_frames, sum_biased, mel, tables, feats);
    else
        fbank_kernel<false><<<grid, 256, sizeof(FbankSmem), stream>>>(pcm, n_samples, n_frames, sum_biased, mel, tables, feats);
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
