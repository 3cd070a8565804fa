// Training forward/backward of ResNetBigger on B200 (see ld_train.h).  Follows the reference's models.py:82-115,181-244
// in .train() mode (BatchNorm batch statistics with biased variance, dropout at the two sites of models.py:232,235)
// and what autograd derives from it for train.py:277-289.
//
// Per conv:  forward   z = conv(x, W)               tcgen05 tap-list GEMM (MODE 1: raw bf16 output + channel sums)
//                      y = relu(bn(z) [+ shortcut])  bn_apply_kernel
//            backward  g, dz from dy                 bn_bwd_reduce_kernel + bn_bwd_apply_kernel
//                      dW = sum_p x(p + s) (x) dz(p)  wgrad_kernel (CUDA cores, fp32 accumulation)
//                      dx = convT(dz, W) [+ shortcut] the same GEMM kernel with the transposed weight slabs
// A conv bias in front of a training-mode BatchNorm cancels in the normalisation: it is not applied and its gradient
// is exactly zero.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "ld_net.h"
#include "ld_ptx.cuh"
#include "ld_train.h"

namespace ld {
namespace {

constexpr float kBnEps = 1e-5f;
constexpr int kHeadFeat = 48, kHeadHidden = 32;

// ------------------------------------------------------------------------------------------------ addressing
__device__ __forceinline__ long long tpix(const TPlane& t, int b, int r, int c, int& which) {
    if (!t.quad) {
        which = 0;
        return (static_cast<long long>(b) * t.hp + 1 + r) * t.wp + 1 + c;
    }
    which = (r & 1) * 2 + (c & 1);
    return (static_cast<long long>(b) * t.hp + 1 + (r >> 1)) * t.wp + 1 + (c >> 1);
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 load8raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
}

// BatchNorm description for the element-wise kernels: sums (sum z, sum z^2) over n elements per channel.
struct BnRef {
    const float* sums;    // [2C] forward statistics
    const float* gamma;   // [C]
    const float* beta;    // [C]
    float inv_n;
};
__device__ __forceinline__ void bn_mean_inv(const BnRef& bn, int C, int c, float& mean, float& inv) {
    const float m = bn.sums[c] * bn.inv_n;
    const float var = fmaxf(bn.sums[C + c] * bn.inv_n - m * m, 0.f);
    mean = m;
    inv = rsqrtf(var + kBnEps);
}

// ------------------------------------------------------------------------------------------------ weight packing
// W (cout, cin, taps) fp32 -> bf16 UMMA B-operand slabs [slab][K/8][N][8].
//   transpose = 0 (forward):  K = cin,  N = cout : slab[t][ci/8][co][ci%8] = W[co][ci][t]
//   transpose = 1 (dgrad):    K = cout, N = cin  : slab[t][co/8][ci][co%8] = W[co][ci][t]
// All weight slabs of a step in ONE launch: blockIdx.y selects the item, blockIdx.x strides over its elements.
struct PackItem {
    long long w_off;        // offset of the fp32 weights in the parameter vector, or -1: identity slab of `cin` channels
    __nv_bfloat16* dst;
    int cout, cin, taps, transpose;
};
__global__ void __launch_bounds__(256) pack_all_kernel(const PackItem* __restrict__ items, const float* __restrict__ params) {
    const PackItem it = items[blockIdx.y];
    if (it.w_off < 0) {
        const int C = it.cin;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * C; i += gridDim.x * blockDim.x) {
            const int k = i / C, n = i % C;
            it.dst[(static_cast<long long>(k / 8) * C + n) * 8 + (k % 8)] = __float2bfloat16_rn(k == n ? 1.f : 0.f);
        }
        return;
    }
    const float* w = params + it.w_off;
    const int cout = it.cout, cin = it.cin, taps = it.taps, n = cout * cin * taps;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int t = i % taps, ci = (i / taps) % cin, co = i / (taps * cin);
        const long long o = it.transpose ? ((static_cast<long long>(t) * (cout / 8) + co / 8) * cin + ci) * 8 + (co % 8)
                                         : ((static_cast<long long>(t) * (cin / 8) + ci / 8) * cout + co) * 8 + (ci % 8);
        it.dst[o] = __float2bfloat16_rn(w[i]);
    }
}

// ------------------------------------------------------------------------------------------------ stem
// Both stem kernels walk the image like the element-wise BatchNorm kernels (EwWalk, below): a warp owns one 8-channel chunk for
// good and visits groups of 32 consecutive real pixels, so per-channel sums stay in registers until one reduction at the end.
struct StemWalk {   // (a copy of EwWalk's arithmetic; EwWalk is declared further down)
    int kc, lane, P, W, H;
    unsigned group, group_stride, n_groups, magic_w, magic_h;
    __device__ StemWalk(int B, int H_, int W_) : W(W_), H(H_) {
        const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
        lane = threadIdx.x & 31;
        kc = static_cast<int>(warp % 8);
        group = warp / 8;
        group_stride = n_warps / 8;   // the launchers keep n_warps a multiple of 8
        P = B * H * W;
        n_groups = (static_cast<unsigned>(P) + 31u) / 32u;
        magic_w = 0xFFFFFFFFu / static_cast<unsigned>(W) + 1u;
        magic_h = 0xFFFFFFFFu / static_cast<unsigned>(H) + 1u;
    }
    __device__ bool pixel(unsigned g, int& b, int& r0, int& c0) const {
        const unsigned pix = g * 32u + lane;
        const unsigned row = __umulhi(pix, magic_w);
        c0 = static_cast<int>(pix - row * W);
        b = static_cast<int>(__umulhi(row, magic_h));
        r0 = static_cast<int>(row - b * H);
        return pix < static_cast<unsigned>(P);
    }
};
__device__ __forceinline__ void stem_patch(const float* __restrict__ x, int b, int r0, int c0, int H, int W, float (&in)[9]) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int r = r0 + ky - 1, c = c0 + kx - 1;
            in[ky * 3 + kx] = (r >= 0 && r < H && c >= 0 && c < W) ? __ldg(x + (static_cast<long long>(b) * H + r) * W + c) : 0.f;
        }
}

// conv1 (1 -> 64, 3x3, pad 1, no bias) on the fp32 features: raw output z0 (bf16, plain) + channel sums.
__global__ void __launch_bounds__(256)
stem_fwd_kernel(const float* __restrict__ x, int B, int H, int W, const float* __restrict__ w /*[64][9]*/, TPlane z,
                float* __restrict__ stats /*[128]*/) {
    __shared__ float s_acc[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const StemWalk wk(B, H, W);
    const int kc = wk.kc;
    float wt[8][9], sum[8], sq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        sum[e] = 0.f; sq[e] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) wt[e][t] = w[(kc * 8 + e) * 9 + t];
    }
    for (unsigned g = wk.group; g < wk.n_groups; g += wk.group_stride) {
        int b, r0, c0;
        if (!wk.pixel(g, b, r0, c0)) continue;
        float in[9], o[8];
        stem_patch(x, b, r0, c0, H, W, in);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) a = fmaf(wt[e][t], in[t], a);
            o[e] = a;
            sum[e] += a;
            sq[e] = fmaf(a, a, sq[e]);
        }
        store8(z.base[0] + kc * z.kc_stride + static_cast<long long>((b * z.hp + 1 + r0) * z.wp + 1 + c0) * 8, o);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            sum[e] += __shfl_xor_sync(0xffffffffu, sum[e], o);
            sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], o);
        }
    }
    if (wk.lane == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(s_acc + kc * 8 + e, sum[e]);
            atomicAdd(s_acc + 64 + kc * 8 + e, sq[e]);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 128; k += blockDim.x) atomicAdd(stats + k, s_acc[k]);
}

// dW conv1 [64][9] = sum_p dz0[p][ch] * x(p + tap): 8 channels x 9 taps of partial sums per thread, reduced once at the end.
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x, int B, int H, int W, TPlane dz, float* __restrict__ dw /*[64][9]*/) {
    __shared__ float s_acc[64 * 9];
    for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const StemWalk wk(B, H, W);
    const int kc = wk.kc;
    float acc[8][9];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[e][t] = 0.f;
    for (unsigned g = wk.group; g < wk.n_groups; g += wk.group_stride) {
        int b, r0, c0;
        if (!wk.pixel(g, b, r0, c0)) continue;
        float in[9], gz[8];
        stem_patch(x, b, r0, c0, H, W, in);
        load8(dz.base[0] + kc * dz.kc_stride + static_cast<long long>((b * dz.hp + 1 + r0) * dz.wp + 1 + c0) * 8, gz);
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[e][t] = fmaf(gz[e], in[t], acc[e][t]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float v = acc[e][t];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (wk.lane == 0) atomicAdd(s_acc + (kc * 8 + e) * 9 + t, v);
        }
    __syncthreads();
    for (int k = threadIdx.x; k < 64 * 9; k += blockDim.x) atomicAdd(dw + k, s_acc[k]);
}

// ------------------------------------------------------------------------------------------------ BatchNorm forward
// Per-channel affine form of a batch-statistics BatchNorm, built once per block in shared memory:
//   xhat = z * xa + xb   (xa = inv std, xb = -mean * inv std)        y = z * ya + yb   (ya = gamma * xa, yb = beta + gamma * xb)
struct BnCoef {
    float xa[64], xb[64], ya[64], yb[64];
};
__device__ __forceinline__ void bn_coef_setup(BnCoef& s, const BnRef& bn, int C) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, inv;
        bn_mean_inv(bn, C, c, mean, inv);
        s.xa[c] = inv; s.xb[c] = -mean * inv;
        s.ya[c] = bn.gamma[c] * inv; s.yb[c] = fmaf(-mean * inv, bn.gamma[c], bn.beta[c]);
    }
}
// Work split of the element-wise BatchNorm kernels: warp w of the grid owns channel chunk w % chunks for good and walks
// groups of 32 consecutive real pixels (a 512-byte contiguous run of every plane it touches) with stride n_warps / chunks,
// so the per-channel coefficients sit in registers and the reductions stay in registers until the end.  All index math is
// 32-bit (B * H * W <= 1024 * 100 * 44).
constexpr int kEwUnroll = 4;   // pixel groups whose loads an element-wise BatchNorm kernel keeps in flight per warp
constexpr int kEwUnrollBwdApply = 2;   // (three loads and two stores per group: 4 would spill at two CTAs per SM)
struct EwWalk {
    int kc, lane, P, W, H;
    unsigned group, group_stride, n_groups, magic_w, magic_h;
    __device__ EwWalk(int chunks, int B, int H_, int W_) : W(W_), H(H_) {
        const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
        lane = threadIdx.x & 31;
        kc = static_cast<int>(warp % chunks);
        group = warp / chunks;
        group_stride = n_warps / chunks;   // the launchers keep n_warps a multiple of 8 >= chunks
        P = B * H * W;
        n_groups = (static_cast<unsigned>(P) + 31u) / 32u;
        // n / d = umulhi(n, floor(2^32 / d) + 1), exact for n * d < 2^32 (here n < 2^23, d <= 100)
        magic_w = 0xFFFFFFFFu / static_cast<unsigned>(W) + 1u;
        magic_h = 0xFFFFFFFFu / static_cast<unsigned>(H) + 1u;
    }
    __device__ bool pixel(unsigned g, int& b, int& r0, int& c0) const {
        const unsigned pix = g * 32u + lane;
        const unsigned row = __umulhi(pix, magic_w);
        c0 = static_cast<int>(pix - row * W);
        b = static_cast<int>(__umulhi(row, magic_h));
        r0 = static_cast<int>(row - b * H);
        return pix < static_cast<unsigned>(P);
    }
};
// Element offset of real pixel (b, r, c) in a plain plane / in a quad plane set (32-bit pixel index: planes stay below 2^28 pixels).
__device__ __forceinline__ long long plain_off(const TPlane& t, int b, int r, int c) {
    return static_cast<long long>((b * t.hp + 1 + r) * t.wp + 1 + c) * 8;
}
__device__ __forceinline__ long long any_off(const TPlane& t, int b, int r, int c, long long plain, int& which) {
    if (!t.quad) { which = 0; return plain; }
    which = (r & 1) * 2 + (c & 1);
    return static_cast<long long>((b * t.hp + 1 + (r >> 1)) * t.wp + 1 + (c >> 1)) * 8;
}

// y = act( bn(z) [+ res] ),  res = plane value (res_mode 1) or bn_s(zs) (res_mode 2).  z and res planes are plain; y may be
// plain or quad.
__global__ void __launch_bounds__(256, 2)
bn_apply_kernel(TPlane z, BnRef bn, int relu, int res_mode, TPlane res, BnRef bn_res, TPlane y, int B) {
    __shared__ BnCoef s, sr;
    const int C = z.C;
    bn_coef_setup(s, bn, C);
    if (res_mode == 2) bn_coef_setup(sr, bn_res, C);
    __syncthreads();
    const EwWalk w(C / 8, B, z.H, z.W);
    const int kc = w.kc;
    float ya[8], yb[8], ra[8], rb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        ya[e] = s.ya[kc * 8 + e]; yb[e] = s.yb[kc * 8 + e];
        ra[e] = res_mode == 2 ? sr.ya[kc * 8 + e] : 1.f; rb[e] = res_mode == 2 ? sr.yb[kc * 8 + e] : 0.f;
    }
    // kEwUnroll groups per iteration: all their 16-byte loads are issued before the first is consumed (memory-level parallelism)
    for (unsigned g0 = w.group; g0 < w.n_groups; g0 += kEwUnroll * w.group_stride) {
        bool live[kEwUnroll];
        int bb[kEwUnroll], rr[kEwUnroll], cc[kEwUnroll];
        long long pz[kEwUnroll];
        uint4 qz[kEwUnroll], qr[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const unsigned g = g0 + u * w.group_stride;
            live[u] = g < w.n_groups && w.pixel(g, bb[u], rr[u], cc[u]);
            if (live[u]) {
                pz[u] = plain_off(z, bb[u], rr[u], cc[u]);   // z and res are plain planes of the same geometry
                qz[u] = load8raw(z.base[0] + kc * z.kc_stride + pz[u]);
                if (res_mode != 0) qr[u] = load8raw(res.base[0] + kc * res.kc_stride + pz[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            if (!live[u]) continue;
            float v[8], o[8], rv[8];
            unpack8(qz[u], v);
            if (res_mode != 0) unpack8(qr[u], rv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float a = fmaf(v[e], ya[e], yb[e]);
                if (res_mode != 0) a += fmaf(rv[e], ra[e], rb[e]);
                o[e] = relu ? fmaxf(a, 0.f) : a;
            }
            int which;
            const long long py = any_off(y, bb[u], rr[u], cc[u], pz[u], which);
            store8(y.base[which] + kc * y.kc_stride + py, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------ BatchNorm backward
// g = dy * [y > 0] (relu) ; reductions sum g, sum g*xhat per channel.
// relu = 2: y = relu(bn(z)) without a residual, so [y > 0] = [z * ya + yb > 0] is recomputed from z with the very expression the
// forward pass evaluated (bn_apply_kernel) and the y plane is not read at all.
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(TPlane dy, TPlane y, int relu, TPlane z, BnRef bn, int B, float* __restrict__ sums /*[2C]*/) {
    __shared__ float s_acc[128];
    __shared__ BnCoef s;
    const int C = z.C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    bn_coef_setup(s, bn, C);
    __syncthreads();
    const EwWalk w(C / 8, B, z.H, z.W);
    const int kc = w.kc;
    float xa[8], xb[8], ya[8], yb[8], sg[8], sx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        xa[e] = s.xa[kc * 8 + e]; xb[e] = s.xb[kc * 8 + e]; ya[e] = s.ya[kc * 8 + e]; yb[e] = s.yb[kc * 8 + e];
        sg[e] = 0.f; sx[e] = 0.f;
    }
    for (unsigned g0 = w.group; g0 < w.n_groups; g0 += kEwUnroll * w.group_stride) {
        bool live[kEwUnroll];
        uint4 qg[kEwUnroll], qy[kEwUnroll], qz[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const unsigned g = g0 + u * w.group_stride;
            int b, r0, c0, which;
            live[u] = g < w.n_groups && w.pixel(g, b, r0, c0);
            if (live[u]) {
                const long long pz = plain_off(z, b, r0, c0);   // dy and y share one geometry (plain like z, or quad)
                const long long pd = any_off(dy, b, r0, c0, pz, which);
                qg[u] = load8raw(dy.base[which] + kc * dy.kc_stride + pd);
                if (relu == 1) qy[u] = load8raw(y.base[which] + kc * y.kc_stride + pd);
                qz[u] = load8raw(z.base[0] + kc * z.kc_stride + pz);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            if (!live[u]) continue;
            float gg[8], zz[8], yy[8];
            unpack8(qg[u], gg);
            if (relu == 1) unpack8(qy[u], yy);
            unpack8(qz[u], zz);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const bool off = relu == 1 ? !(yy[e] > 0.f) : (relu == 2 ? !(fmaf(zz[e], ya[e], yb[e]) > 0.f) : false);
                const float ge = off ? 0.f : gg[e];
                sg[e] += ge;
                sx[e] = fmaf(ge, fmaf(zz[e], xa[e], xb[e]), sx[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            sg[e] += __shfl_xor_sync(0xffffffffu, sg[e], o);
            sx[e] += __shfl_xor_sync(0xffffffffu, sx[e], o);
        }
    }
    if (w.lane == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            atomicAdd(s_acc + kc * 8 + e, sg[e]);
            atomicAdd(s_acc + C + kc * 8 + e, sx[e]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, s_acc[i]);
}

// dz = gamma * inv * (g - mean(g) - xhat * mean(g xhat)) -> dz plane (plain); optionally g -> g_out (plain);
// block 0 also writes dgamma = sum g xhat, dbeta = sum g.
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(TPlane dy, TPlane y, int relu, TPlane z, BnRef bn, const float* __restrict__ sums, int B, TPlane dz, int write_g,
                    TPlane g_out, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ BnCoef s;
    const int C = z.C;
    bn_coef_setup(s, bn, C);
    if (blockIdx.x == 0 && threadIdx.x < C) {
        dgamma[threadIdx.x] = sums[C + threadIdx.x];
        dbeta[threadIdx.x] = sums[threadIdx.x];
    }
    __syncthreads();
    const EwWalk w(C / 8, B, z.H, z.W);
    const int kc = w.kc;
    float xa[8], xb[8], ya[8], yb[8], c1[8], c2[8];   // c1 = mean(g), c2 = mean(g xhat)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = kc * 8 + e;
        xa[e] = s.xa[c]; xb[e] = s.xb[c]; ya[e] = s.ya[c]; yb[e] = s.yb[c];
        c1[e] = sums[c] * bn.inv_n; c2[e] = sums[C + c] * bn.inv_n;
    }
    for (unsigned g0 = w.group; g0 < w.n_groups; g0 += kEwUnrollBwdApply * w.group_stride) {
        bool live[kEwUnrollBwdApply];
        long long pzs[kEwUnrollBwdApply];
        uint4 qg[kEwUnrollBwdApply], qy[kEwUnrollBwdApply], qz[kEwUnrollBwdApply];
#pragma unroll
        for (int u = 0; u < kEwUnrollBwdApply; ++u) {
            const unsigned g = g0 + u * w.group_stride;
            int b, r0, c0, which;
            live[u] = g < w.n_groups && w.pixel(g, b, r0, c0);
            if (live[u]) {
                pzs[u] = plain_off(z, b, r0, c0);   // dy and y share one geometry (plain like z, or quad); dz, g_out are plain
                const long long pd = any_off(dy, b, r0, c0, pzs[u], which);
                qg[u] = load8raw(dy.base[which] + kc * dy.kc_stride + pd);
                if (relu == 1) qy[u] = load8raw(y.base[which] + kc * y.kc_stride + pd);
                qz[u] = load8raw(z.base[0] + kc * z.kc_stride + pzs[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnrollBwdApply; ++u) {
            if (!live[u]) continue;
            float gg[8], zz[8], yy[8], o[8];
            unpack8(qg[u], gg);
            if (relu == 1) unpack8(qy[u], yy);
            unpack8(qz[u], zz);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (relu == 1 && !(yy[e] > 0.f)) gg[e] = 0.f;
                if (relu == 2 && !(fmaf(zz[e], ya[e], yb[e]) > 0.f)) gg[e] = 0.f;   // [y > 0] recomputed from z (no residual)
                const float xh = fmaf(zz[e], xa[e], xb[e]);
                o[e] = ya[e] * (gg[e] - c1[e] - xh * c2[e]);
            }
            store8(dz.base[0] + kc * dz.kc_stride + pzs[u], o);
            if (write_g) store8(g_out.base[0] + kc * g_out.kc_stride + pzs[u], gg);
        }
    }
}

// mean / biased variance of every conv BatchNorm from the channel sums, conv bias added back to the mean (the kernels leave
// it out because it cancels in the normalisation, but the module's running_mean tracks the biased conv output).
struct BnFinalizeItem {
    int sums_off, out_off, C;
    long long bias_off;   // into the parameter vector, or -1
    float inv_n;
};
__global__ void bn_finalize_kernel(const BnFinalizeItem* __restrict__ items, int n_items, const float* __restrict__ stats,
                                   const float* __restrict__ params, float* __restrict__ bn_stats) {
    const BnFinalizeItem it = items[blockIdx.x];
    for (int c = threadIdx.x; c < it.C; c += blockDim.x) {
        const float m = stats[it.sums_off + c] * it.inv_n;
        const float var = fmaxf(stats[it.sums_off + it.C + c] * it.inv_n - m * m, 0.f);
        bn_stats[it.out_off + c] = m + (it.bias_off >= 0 ? params[it.bias_off + c] : 0.f);
        bn_stats[it.out_off + it.C + c] = var;
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co][ci][t] += sum_p X_t[p + shift_t][ci] * dZ[p][co] over the pixel range of the block.  256 threads = 16 x 16,
// thread (ti, tj) owns ci in [ti*TI, ti*TI+TI) and co in [tj*TJ, tj*TJ+TJ) for every tap.
struct WgradTaps {
    const __nv_bfloat16* src[9];
    long long kc_stride[9];
    int shift[9];
};
template <int CIN, int COUT, int NT>
__global__ void __launch_bounds__(256)
wgrad_kernel(WgradTaps taps, const __nv_bfloat16* __restrict__ dz, long long dz_kc_stride, long long M, float* __restrict__ dw) {
    constexpr int PT = 32;                 // pixels per smem tile
    constexpr int TI = CIN / 16, TJ = COUT / 16;
    __shared__ __align__(16) __nv_bfloat16 s_x[NT][PT][CIN];
    __shared__ __align__(16) __nv_bfloat16 s_dz[PT][COUT];
    const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
    float acc[NT][TI][TJ];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) acc[t][a][b] = 0.f;
    const long long tiles = (M + PT - 1) / PT;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = tile * PT;
        // stage: 16-byte pieces, piece = (pixel, chunk)
        for (int i = threadIdx.x; i < PT * (COUT / 8); i += 256) {
            const int px = i % PT, kc = i / PT;
            const long long p = p0 + px;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (p < M) v = *reinterpret_cast<const uint4*>(dz + kc * dz_kc_stride + p * 8);
            *reinterpret_cast<uint4*>(&s_dz[px][kc * 8]) = v;
        }
        for (int i = threadIdx.x; i < NT * PT * (CIN / 8); i += 256) {
            const int px = i % PT, kc = (i / PT) % (CIN / 8), t = i / (PT * (CIN / 8));
            const long long p = p0 + px;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (p < M) v = *reinterpret_cast<const uint4*>(taps.src[t] + kc * taps.kc_stride[t] + (p + taps.shift[t]) * 8);
            *reinterpret_cast<uint4*>(&s_x[t][px][kc * 8]) = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int px = 0; px < PT; ++px) {
            float d[TJ];
#pragma unroll
            for (int b = 0; b < TJ; ++b) d[b] = __bfloat162float(s_dz[px][tj * TJ + b]);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
#pragma unroll
                for (int a = 0; a < TI; ++a) {
                    const float xv = __bfloat162float(s_x[t][px][ti * TI + a]);
#pragma unroll
                    for (int b = 0; b < TJ; ++b) acc[t][a][b] = fmaf(xv, d[b], acc[t][a][b]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b)
                atomicAdd(dw + (static_cast<long long>(tj * TJ + b) * CIN + ti * TI + a) * NT + t, acc[t][a][b]);
}

// ------------------------------------------------------------------------------------------------ head
// Scratch layout (floats), B = batch: pooled[B][48] xh2[B][48] d1[B][48] l1[B][32] xh3[B][32] d2[B][32] out[B]
//                                     inv2[48] inv3[32] mean2[48] var2[48] mean3[32] var3[32]
struct HeadScratch {
    float *pooled, *xh2, *d1, *l1, *xh3, *d2, *out, *inv2, *inv3;
};
struct HeadParams {
    const float *g2, *b2, *g3, *b3, *w1, *bias1, *w2, *bias2;   // bn2, bn3, linear1 (32,48), linear2 (1,32)
};
// AvgPool2d(4): rows 4g..4g+3, cols 0..3; feature f = c * 3 + g   (models.py:229-230).  Grid over (sample, feature).
__global__ void __launch_bounds__(256)
head_pool_kernel(TPlane y, int B, float* __restrict__ pooled) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * kHeadFeat) return;
    const int b = i / kHeadFeat, f = i % kHeadFeat, c = f / 3, g = f % 3;
    float a = 0.f;
    for (int r = 4 * g; r < 4 * g + 4; ++r)
        for (int col = 0; col < 4; ++col) {
            int which;
            const long long p = tpix(y, b, r, col, which);
            a += __bfloat162float(y.base[which][(c / 8) * y.kc_stride + p * 8 + (c % 8)]);
        }
    pooled[i] = a * (1.f / 16.f);
}
// Average-pool backward into the (plain) gradient plane of the last block output: every real pixel is written (zeros outside
// the pooled 12 x 4 corner).  Grid over (sample, pixel, 8-channel chunk), one 16-byte store each.
__global__ void __launch_bounds__(256)
head_unpool_kernel(TPlane dy, int B, const float* __restrict__ dpool) {
    const int chunks = dy.C / 8;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dy.H * dy.W * chunks) return;
    const int col = i % dy.W, r = (i / dy.W) % dy.H, kc = (i / (dy.W * dy.H)) % chunks, b = i / (dy.W * dy.H * chunks);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
        v[e] = (r < 12 && col < 4) ? dpool[b * kHeadFeat + (kc * 8 + e) * 3 + r / 4] * (1.f / 16.f) : 0.f;
    int which;
    const long long p = tpix(dy, b, r, col, which);
    store8(dy.base[which] + kc * dy.kc_stride + p * 8, v);
}

// The dense part of the head on (B, 48) / (B, 32) matrices, one CTA of 1024 threads.  Batch reductions per feature: thread t
// owns feature t % F and the samples t / F, t / F + slices, ... (adjacent threads read adjacent features: coalesced), the
// slices are combined through shared memory; everything element-wise runs over all threads.
constexpr int kHeadThreads = 1024;
template <int NV, typename Fn>
__device__ __forceinline__ void column_reduce(int B, int F, float (*s_part)[kHeadThreads], float (*s_out)[kHeadFeat], Fn fn) {
    const int tid = threadIdx.x, slices = kHeadThreads / F, f = tid % F, sl = tid / F;
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
    if (sl < slices)
        for (int b = sl; b < B; b += slices) fn(b, f, acc);
#pragma unroll
    for (int k = 0; k < NV; ++k) s_part[k][tid] = acc[k];
    __syncthreads();
    if (tid < F) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float a = 0.f;
            for (int j = 0; j < slices; ++j) a += s_part[k][j * F + tid];
            s_out[k][tid] = a;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(int B, HeadParams hp, const float* __restrict__ mask1, const float* __restrict__ mask2, float keep_scale,
                HeadScratch s, float* __restrict__ probs, float* __restrict__ stat_out /* mean2 var2 mean3 var3 */) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ float s_part[3][kHeadThreads], s_red[3][kHeadFeat];
    __shared__ float s_mean[kHeadFeat], s_inv[kHeadFeat];
    // bn2 over the batch (biased variance, two passes)
    column_reduce<1>(B, kHeadFeat, s_part, s_red, [&](int b, int f, float* a) { a[0] += s.pooled[b * kHeadFeat + f]; });
    if (tid < kHeadFeat) s_mean[tid] = s_red[0][tid] / B;
    __syncthreads();
    column_reduce<1>(B, kHeadFeat, s_part, s_red, [&](int b, int f, float* a) {
        const float d = s.pooled[b * kHeadFeat + f] - s_mean[f];
        a[0] = fmaf(d, d, a[0]);
    });
    if (tid < kHeadFeat) {
        const float q = s_red[0][tid] / B, inv = rsqrtf(q + kBnEps);
        s_inv[tid] = inv;
        s.inv2[tid] = inv;
        stat_out[tid] = s_mean[tid]; stat_out[kHeadFeat + tid] = q;
    }
    __syncthreads();
    for (int i = tid; i < B * kHeadFeat; i += nt) {   // normalise, affine, dropout 1
        const int f = i % kHeadFeat;
        const float xh = (s.pooled[i] - s_mean[f]) * s_inv[f];
        s.xh2[i] = xh;
        s.d1[i] = fmaf(xh, hp.g2[f], hp.b2[f]) * mask1[i] * keep_scale;
    }
    __syncthreads();
    for (int i = tid; i < B * kHeadHidden; i += nt) {   // linear1
        const int b = i / kHeadHidden, o = i % kHeadHidden;
        float a = hp.bias1[o];
        for (int f = 0; f < kHeadFeat; ++f) a = fmaf(hp.w1[o * kHeadFeat + f], s.d1[b * kHeadFeat + f], a);
        s.l1[i] = a;
    }
    __syncthreads();
    // bn3 over the batch
    column_reduce<1>(B, kHeadHidden, s_part, s_red, [&](int b, int o, float* a) { a[0] += s.l1[b * kHeadHidden + o]; });
    if (tid < kHeadHidden) s_mean[tid] = s_red[0][tid] / B;
    __syncthreads();
    column_reduce<1>(B, kHeadHidden, s_part, s_red, [&](int b, int o, float* a) {
        const float d = s.l1[b * kHeadHidden + o] - s_mean[o];
        a[0] = fmaf(d, d, a[0]);
    });
    if (tid < kHeadHidden) {
        const float q = s_red[0][tid] / B, inv = rsqrtf(q + kBnEps);
        s_inv[tid] = inv;
        s.inv3[tid] = inv;
        stat_out[2 * kHeadFeat + tid] = s_mean[tid]; stat_out[2 * kHeadFeat + kHeadHidden + tid] = q;
    }
    __syncthreads();
    for (int i = tid; i < B * kHeadHidden; i += nt) {   // normalise, affine, dropout 2 (the ReLU is applied by the consumers)
        const int o = i % kHeadHidden;
        const float xh = (s.l1[i] - s_mean[o]) * s_inv[o];
        s.xh3[i] = xh;
        s.d2[i] = fmaf(xh, hp.g3[o], hp.b3[o]) * mask2[i] * keep_scale;
    }
    __syncthreads();
    for (int b = tid; b < B; b += nt) {   // linear2 + sigmoid
        float zacc = hp.bias2[0];
        for (int o = 0; o < kHeadHidden; ++o) zacc = fmaf(hp.w2[o], fmaxf(s.d2[b * kHeadHidden + o], 0.f), zacc);
        const float p = 1.f / (1.f + expf(-zacc));
        s.out[b] = p;
        probs[b] = p;
    }
}

struct HeadGrads {
    float *g2, *b2, *g3, *b3, *w1, *bias1, *w2, *bias2;
};
// Scratch for backward (floats): dl1[B][32] dd1[B][48] dpool[B][48]
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(int B, HeadParams hp, const float* __restrict__ mask1, const float* __restrict__ mask2, float keep_scale,
                HeadScratch s, const float* __restrict__ dprobs, float* __restrict__ dl1, float* __restrict__ dd1,
                float* __restrict__ dpool, HeadGrads hg) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ float s_dz[1024];   // dL/dlogit per sample (B <= 1024)
    __shared__ float s_part[3][kHeadThreads], s_red[3][kHeadFeat];
    for (int b = tid; b < B; b += nt) s_dz[b] = dprobs[b] * s.out[b] * (1.f - s.out[b]);
    __syncthreads();
    // linear2, relu, dropout 2: da3 = dL/d(bn3 output); reductions for dW2, dbeta3, dgamma3 (and dbias2 = sum of dz)
    column_reduce<3>(B, kHeadHidden, s_part, s_red, [&](int b, int o, float* a) {
        const float d2 = s.d2[b * kHeadHidden + o];
        a[0] = fmaf(s_dz[b], fmaxf(d2, 0.f), a[0]);
        const float da3 = (d2 > 0.f ? s_dz[b] * hp.w2[o] : 0.f) * mask2[b * kHeadHidden + o] * keep_scale;
        a[1] += da3;
        a[2] = fmaf(da3, s.xh3[b * kHeadHidden + o], a[2]);
    });
    if (tid < kHeadHidden) { hg.w2[tid] = s_red[0][tid]; hg.b3[tid] = s_red[1][tid]; hg.g3[tid] = s_red[2][tid]; }
    for (int i = tid; i < B * kHeadHidden; i += nt) {   // bn3 backward
        const int b = i / kHeadHidden, o = i % kHeadHidden;
        const float d2 = s.d2[i];
        const float da3 = (d2 > 0.f ? s_dz[b] * hp.w2[o] : 0.f) * mask2[i] * keep_scale;
        dl1[i] = hp.g3[o] * s.inv3[o] * (da3 - s_red[1][o] / B - s.xh3[i] * s_red[2][o] / B);
    }
    __syncthreads();
    // dbias1 = column sums of dl1; dbias2 = sum of dz (column 0 of a one-column reduction)
    column_reduce<1>(B, kHeadHidden, s_part, s_red, [&](int b, int o, float* a) { a[0] += dl1[b * kHeadHidden + o]; });
    if (tid < kHeadHidden) hg.bias1[tid] = s_red[0][tid];
    __syncthreads();
    column_reduce<1>(B, 1, s_part, s_red, [&](int b, int, float* a) { a[0] += s_dz[b]; });
    if (tid == 0) hg.bias2[0] = s_red[0][0];
    for (int i = tid; i < kHeadHidden * kHeadFeat; i += nt) {   // dW1[o][f] = sum_b dl1[b][o] d1[b][f]; adjacent threads: adjacent f
        const int o = i / kHeadFeat, f = i % kHeadFeat;
        float a = 0.f;
#pragma unroll 8
        for (int b = 0; b < B; ++b) a = fmaf(dl1[b * kHeadHidden + o], s.d1[b * kHeadFeat + f], a);
        hg.w1[i] = a;
    }
    for (int i = tid; i < B * kHeadFeat; i += nt) {   // dd1 = W1^T dl1, through dropout 1
        const int b = i / kHeadFeat, f = i % kHeadFeat;
        float a = 0.f;
        for (int o = 0; o < kHeadHidden; ++o) a = fmaf(hp.w1[o * kHeadFeat + f], dl1[b * kHeadHidden + o], a);
        dd1[i] = a * mask1[i] * keep_scale;
    }
    __syncthreads();
    // bn2 backward
    column_reduce<2>(B, kHeadFeat, s_part, s_red, [&](int b, int f, float* a) {
        a[0] += dd1[b * kHeadFeat + f];
        a[1] = fmaf(dd1[b * kHeadFeat + f], s.xh2[b * kHeadFeat + f], a[1]);
    });
    if (tid < kHeadFeat) { hg.b2[tid] = s_red[0][tid]; hg.g2[tid] = s_red[1][tid]; }
    for (int i = tid; i < B * kHeadFeat; i += nt) {
        const int f = i % kHeadFeat;
        dpool[i] = hp.g2[f] * s.inv2[f] * (dd1[i] - s_red[0][f] / B - s.xh2[i] * s_red[1][f] / B);
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct BnHost {
    int C = 0;
    long long gamma_off = 0, beta_off = 0;   // into the flat parameter / gradient buffers
    int fwd_sums = 0, bwd_sums = 0;           // float offsets into the statistics buffer
    int stat_out = 0;                         // float offset into the bn_stats output (mean[C], var[C])
    long long count = 0;                      // elements per channel (B * H * W), set per call
    bool reduce_fused = false;                // sum g / sum g xhat come from the epilogue of the dgrad GEMM that produces dy
};
struct ConvHost {
    std::string name;
    int cin = 0, cout = 0, ksize = 3, stride = 1;
    long long w_off = 0, b_off = -1;
    int in_level = -1;              // index into levels (activation read)
    TPlane z{}, dz{};
    BnHost bn;
    __nv_bfloat16 *w_fwd = nullptr, *w_bwd = nullptr;   // packed slabs
    int bwd_slabs = 0;
    GemmLaunch fwd{}, bwd{};
    bool has_bwd = false;
    std::vector<HostTap> fwd_taps;  // forward taps (also the X operands of wgrad)
    WgradLaunch wg{};               // tensor-core weight gradient
};
struct BlockHost {
    int conv1 = -1, conv2 = -1, sc = -1;   // conv indices
    int in_level = -1, h_level = -1, out_level = -1;
    TPlane g{};      // dL/d(block output pre-activation sum), plain
    TPlane dh{};     // gradient wrt h
};

}  // namespace

class TrainNet {
public:
    int max_batch = 0, num_sms = 148;
    NetConfig cfg;
    uint8_t* workspace = nullptr;
    size_t workspace_bytes = 0, cursor = 0;
    std::vector<TPlane> levels;       // activations y: 0 = stem output, then h / out of every block
    std::vector<TPlane> dlevels;      // gradients wrt block inputs / stem output (same storage as the level)
    std::vector<ConvHost> convs;      // 0 = stem (CUDA cores), then block convs
    std::vector<BlockHost> blocks;
    std::vector<TrainParamInfo> params;
    std::vector<TrainParamInfo> bn_table;   // BatchNorm module name, offset into bn_stats (mean[C] then var[C]), C
    long long n_params = 0;
    float* stats = nullptr;           // forward + backward channel sums of every BatchNorm
    int stats_floats = 0, bn_stats_floats = 0;
    BnHost head_bn2, head_bn3;
    long long head_off[8] = {0};      // bn2.w bn2.b bn3.w bn3.b linear1.w linear1.b linear2.w linear2.b
    float* head_scratch = nullptr;
    BnFinalizeItem* bn_items = nullptr;
    std::vector<BnFinalizeItem> bn_items_host;
    int bn_items_batch = -1;          // batch size the device copy of the table was built for
    // optional: weight gradients on a side stream -- they only feed the parameter gradients, so they could overlap the BatchNorm
    // element-wise passes of the data-gradient chain (tensor-bound next to HBM-bound work)
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> fork_events;
    cudaEvent_t join_event = nullptr;
    bool side_wgrad = false;          // LD_TRAIN_SIDE=1 enables it.  Measured: 5.08-5.11 ms per step against 4.90-5.03 ms with everything on
                                      // the caller's stream (same box, CUDA graph): the persistent weight-gradient CTAs take the SMs from the
                                      // element-wise kernels instead of sharing them -- off by default
    void* pack_items = nullptr;       // device table of the weight slabs pack_all_kernel writes every step
    int n_pack_items = 0;
    float *x_keep = nullptr, *mask1_keep = nullptr, *mask2_keep = nullptr, *params_keep = nullptr;   // inputs of the last forward
    HeadScratch hs{};
    float *dl1 = nullptr, *dd1 = nullptr, *dpool = nullptr;
    TPlane dy_last{};
    GemmTuning tune{};
    bool wgrad_mma = true;            // LD_WGRAD=cuda selects the CUDA-core weight-gradient kernel instead
    bool fuse_bwd_stats = false;      // LD_TRAIN_FUSE_BWD=1: BatchNorm-backward sums in the epilogue of the dgrad GEMMs instead of separate
                                      // bn_bwd_reduce passes.  Measured SLOWER (6.14 vs 5.40 ms per step, profiles/r02): the epilogue of
                                      // the single-output training GEMMs is already their critical path.  Kept for cross-checking.
    int max_batch_seen = 0;
    long long launches = 0;
    // per-call state
    int B = 0;
    const float *params_d = nullptr, *x_d = nullptr, *mask1_d = nullptr, *mask2_d = nullptr;
    float keep_scale = 1.f;

    ~TrainNet() {
        if (workspace) cudaFree(workspace);
        if (stats) cudaFree(stats);
        if (head_scratch) cudaFree(head_scratch);
        if (x_keep) cudaFree(x_keep);
        if (bn_items) cudaFree(bn_items);
        if (pack_items) cudaFree(pack_items);
        for (auto e : fork_events) if (e) cudaEventDestroy(e);
        if (join_event) cudaEventDestroy(join_event);
        if (side) cudaStreamDestroy(side);
        for (auto& c : convs) { gemm_release(c.fwd); gemm_release(c.bwd); }
    }

    // ---- allocation: bf16 chunk-planar planes with zero guards, carved from one zero-initialised workspace
    // pad = 1: ONE zero column between consecutive rows and ONE zero row between consecutive images (the last column / row of an
    // image is followed by the next row's / image's pad: 3 % fewer pixels at the 100 x 44 level, 18 % at 13 x 6); pad = 2: a zero
    // border on every side (LD_TRAIN_PAD=2, the round-1 layout, kept for A/B runs)
    int pad = 1;
    size_t plane_bytes(int H, int W, int C, int quad) const {
        const int h = quad ? (H + 1) / 2 : H, w = quad ? (W + 1) / 2 : W;
        const long long pixels = static_cast<long long>(max_batch) * (h + pad) * (w + pad);
        const long long guard = 2 * (w + pad) + 160;
        const long long alloc = (pixels + 2 * guard + 7) & ~7ll;
        return static_cast<size_t>(alloc) * 16 * (C / 8) * (quad ? 4 : 1);
    }
    TPlane carve(int H, int W, int C, int quad) {
        TPlane t{};
        t.H = H; t.W = W; t.C = C; t.quad = quad;
        const int h = quad ? (H + 1) / 2 : H, w = quad ? (W + 1) / 2 : W;
        t.hp = h + pad; t.wp = w + pad;
        const long long pixels = static_cast<long long>(max_batch) * t.hp * t.wp;
        const long long guard = 2 * t.wp + 160;
        const long long alloc = (pixels + 2 * guard + 7) & ~7ll;
        t.kc_stride = alloc * 8;
        for (int q = 0; q < (quad ? 4 : 1); ++q) {
            t.base[q] = reinterpret_cast<__nv_bfloat16*>(workspace + cursor) + guard * 8;
            cursor += static_cast<size_t>(alloc) * 16 * (C / 8);
        }
        return t;
    }
};

namespace {

void add_param(TrainNet& n, const std::string& name, long long numel, long long* off = nullptr) {
    if (off) *off = n.n_params;
    n.params.push_back({name, n.n_params, numel});
    n.n_params += numel;
}

int new_bn(TrainNet& n, BnHost& bn, int C, const std::string& name) {
    bn.C = C;
    n.bn_table.push_back({name, n.bn_stats_floats, C});
    bn.fwd_sums = n.stats_floats; n.stats_floats += 2 * C;
    bn.bwd_sums = n.stats_floats; n.stats_floats += 2 * C;
    bn.stat_out = n.bn_stats_floats; n.bn_stats_floats += 2 * C;
    return 0;
}

BnRef bn_ref(const TrainNet& n, const BnHost& bn) {
    BnRef r;
    r.sums = n.stats + bn.fwd_sums;
    r.gamma = n.params_d + bn.gamma_off;
    r.beta = n.params_d + bn.beta_off;
    r.inv_n = 1.f / static_cast<float>(bn.count);
    return r;
}

// Forward taps of a conv reading level `x` (plain for stride 1, quad for stride 2), output geometry (hp, wp) of z.
std::vector<HostTap> forward_taps(const TPlane& x, int ksize, int stride, int wp_out) {
    std::vector<HostTap> taps;
    for (int ky = 0; ky < ksize; ++ky)
        for (int kx = 0; kx < ksize; ++kx) {
            const int dy = ksize == 3 ? ky - 1 : 0, dx = ksize == 3 ? kx - 1 : 0;
            HostTap t{};
            t.kc_stride = x.kc_stride;
            t.wslab = ky * ksize + kx;
            if (stride == 1) {
                t.src = x.base[0];
                t.shift = dy * wp_out + dx;
            } else {
                // real row 2j + dy: even -> row-parity 0 index j ; odd: dy=-1 -> parity 1 index j-1, dy=+1 -> parity 1 index j
                const int rp = dy & 1, cp = dx & 1;
                const int di = dy < 0 ? -1 : 0, dk = dx < 0 ? -1 : 0;
                t.src = x.base[rp * 2 + cp];
                t.shift = di * wp_out + dk;
            }
            taps.push_back(t);
        }
    return taps;
}

}  // namespace

TrainNet* train_create(int max_batch, int num_sms, const NetConfig& cfg, std::string& err) {
    if (max_batch < 2 || max_batch > 1024) { err = "training batch must be in [2, 1024]"; return nullptr; }
    if (cfg.linear_in != kHeadFeat || cfg.H != 100 || cfg.W != 44) { err = "training supports the resnet_base geometry only"; return nullptr; }
    TrainNet* n = new TrainNet();
    n->max_batch = max_batch; n->num_sms = num_sms; n->cfg = cfg;
    n->tune = gemm_tuning_from_env();
    if (const char* v = std::getenv("LD_WGRAD")) n->wgrad_mma = std::string(v) != "cuda";
    if (const char* v = std::getenv("LD_TRAIN_FUSE_BWD")) n->fuse_bwd_stats = std::atoi(v) != 0;
    if (const char* v = std::getenv("LD_TRAIN_SIDE")) n->side_wgrad = std::atoi(v) != 0;
    if (const char* v = std::getenv("LD_TRAIN_PAD")) n->pad = std::atoi(v) == 2 ? 2 : 1;
    if (n->side_wgrad) {
        bool ok = cudaStreamCreateWithFlags(&n->side, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&n->join_event, cudaEventDisableTiming) == cudaSuccess;
        n->fork_events.resize(32, nullptr);
        for (auto& e : n->fork_events) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { err = "could not create the side stream for the weight gradients"; delete n; return nullptr; }
    }

    // ---- topology, parameter table (module registration order of the reference's ResNetBigger) and sizes
    struct LevelSpec { int H, W, C, quad; };
    std::vector<LevelSpec> lv;
    auto out_size = [](int v, int s) { return (v - 1) / s + 1; };
    n->convs.emplace_back();
    {
        ConvHost& c = n->convs[0];
        c.name = "conv1"; c.cin = 1; c.cout = 64; c.ksize = 3; c.stride = 1;
        add_param(*n, "conv1.weight", 576, &c.w_off);
        new_bn(*n, c.bn, 64, "bn1");
        add_param(*n, "bn1.weight", 64, &c.bn.gamma_off);
        add_param(*n, "bn1.bias", 64, &c.bn.beta_off);
    }
    lv.push_back({cfg.H, cfg.W, 64, 0});
    int cur = 0, in_c = 64;
    for (int b = 0; b < 4; ++b) {
        const int out_c = cfg.filters[b];
        for (int r = 0; r < 2; ++r) {
            const int s = (b > 0 && r == 0) ? 2 : 1;
            const int ic = r == 0 ? in_c : out_c;
            const std::string pre = "block" + std::to_string(b + 1) + "." + std::to_string(r);
            BlockHost blk;
            blk.in_level = cur;
            const int Ho = out_size(lv[cur].H, s), Wo = out_size(lv[cur].W, s);
            if (s == 2) lv[cur].quad = 1;
            auto add_conv = [&](const std::string& name, const std::string& bn_name, int cin, int cout, int k, int stride, int in_level, bool bias) {
                ConvHost c;
                c.name = name; c.cin = cin; c.cout = cout; c.ksize = k; c.stride = stride; c.in_level = in_level;
                add_param(*n, name + ".weight", static_cast<long long>(cout) * cin * k * k, &c.w_off);
                if (bias) add_param(*n, name + ".bias", cout, &c.b_off);
                new_bn(*n, c.bn, cout, bn_name);
                add_param(*n, bn_name + ".weight", cout, &c.bn.gamma_off);
                add_param(*n, bn_name + ".bias", cout, &c.bn.beta_off);
                n->convs.push_back(c);
                return static_cast<int>(n->convs.size()) - 1;
            };
            blk.conv1 = add_conv(pre + ".conv1", pre + ".bn1", ic, out_c, 3, s, cur, true);
            lv.push_back({Ho, Wo, out_c, 0});
            blk.h_level = static_cast<int>(lv.size()) - 1;
            blk.conv2 = add_conv(pre + ".conv2", pre + ".bn2", out_c, out_c, 3, 1, blk.h_level, true);
            if (s != 1 || ic != out_c) blk.sc = add_conv(pre + ".shortcut.0", pre + ".shortcut.1", ic, out_c, 1, s, cur, false);
            lv.push_back({Ho, Wo, out_c, 0});
            blk.out_level = static_cast<int>(lv.size()) - 1;
            cur = blk.out_level;
            n->blocks.push_back(blk);
        }
        in_c = out_c;
    }
    new_bn(*n, n->head_bn2, kHeadFeat, "bn2");
    new_bn(*n, n->head_bn3, kHeadHidden, "bn3");
    add_param(*n, "bn2.weight", kHeadFeat, &n->head_off[0]);
    add_param(*n, "bn2.bias", kHeadFeat, &n->head_off[1]);
    add_param(*n, "bn3.weight", kHeadHidden, &n->head_off[2]);
    add_param(*n, "bn3.bias", kHeadHidden, &n->head_off[3]);
    add_param(*n, "linear1.weight", kHeadHidden * kHeadFeat, &n->head_off[4]);
    add_param(*n, "linear1.bias", kHeadHidden, &n->head_off[5]);
    add_param(*n, "linear2.weight", kHeadHidden, &n->head_off[6]);
    add_param(*n, "linear2.bias", 1, &n->head_off[7]);
    if (lv[cur].H != 13 || lv[cur].W != 6 || lv[cur].C * 3 != kHeadFeat) { err = "unexpected head geometry"; delete n; return nullptr; }

    // ---- workspace
    size_t total = 0;
    for (const auto& l : lv) total += 2 * n->plane_bytes(l.H, l.W, l.C, l.quad);          // y and dL/dy
    for (size_t i = 0; i < n->convs.size(); ++i) {
        const ConvHost& c = n->convs[i];
        const LevelSpec& in = i == 0 ? lv[0] : lv[c.in_level];
        const int Ho = i == 0 ? cfg.H : out_size(in.H, c.stride), Wo = i == 0 ? cfg.W : out_size(in.W, c.stride);
        total += 2 * n->plane_bytes(Ho, Wo, c.cout, 0);                                    // z and dz
        total += static_cast<size_t>(c.cout) * c.cin * 10 * 2 * 2 + 4096;                  // packed weights (fwd + bwd + extra slab)
    }
    for (const auto& b : n->blocks) total += 2 * n->plane_bytes(lv[b.out_level].H, lv[b.out_level].W, lv[b.out_level].C, 0);  // g, dh
    total += 1 << 20;
    n->workspace_bytes = total;
    if (cudaMalloc(reinterpret_cast<void**>(&n->workspace), total) != cudaSuccess || cudaMemset(n->workspace, 0, total) != cudaSuccess) {
        err = "cudaMalloc of the training workspace (" + std::to_string(total >> 20) + " MB) failed";
        delete n; return nullptr;
    }
    auto carve_weights = [&](size_t elems) {
        n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255);
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(n->workspace + n->cursor);
        n->cursor += elems * 2;
        return p;
    };
    for (const auto& l : lv) { n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255); n->levels.push_back(n->carve(l.H, l.W, l.C, l.quad)); }
    for (const auto& l : lv) { n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255); n->dlevels.push_back(n->carve(l.H, l.W, l.C, l.quad)); }
    for (size_t i = 0; i < n->convs.size(); ++i) {
        ConvHost& c = n->convs[i];
        const LevelSpec& in = i == 0 ? lv[0] : lv[c.in_level];
        const int Ho = i == 0 ? cfg.H : out_size(in.H, c.stride), Wo = i == 0 ? cfg.W : out_size(in.W, c.stride);
        n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255);
        c.z = n->carve(Ho, Wo, c.cout, 0);
        n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255);
        c.dz = n->carve(Ho, Wo, c.cout, 0);
        if (i > 0) {
            c.w_fwd = carve_weights(static_cast<size_t>(c.ksize) * c.ksize * c.cin * c.cout);
            c.w_bwd = carve_weights(static_cast<size_t>(10) * c.cin * c.cout);
        }
    }
    for (auto& b : n->blocks) {
        const LevelSpec& o = lv[b.out_level];
        n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255);
        b.g = n->carve(o.H, o.W, o.C, 0);
        n->cursor = (n->cursor + 255) & ~static_cast<size_t>(255);
        b.dh = n->carve(o.H, o.W, o.C, 0);
    }
    if (n->cursor > n->workspace_bytes) { err = "internal: training workspace under-estimated"; delete n; return nullptr; }
    n->dy_last = n->dlevels[cur];

    if (cudaMalloc(reinterpret_cast<void**>(&n->stats), n->stats_floats * sizeof(float)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&n->bn_items), 64 * sizeof(BnFinalizeItem)) != cudaSuccess) { err = "cudaMalloc failed"; delete n; return nullptr; }
    {
        const size_t Bm = max_batch;
        const size_t floats = Bm * (3 * kHeadFeat + 3 * kHeadHidden + 1) + kHeadFeat + kHeadHidden + Bm * (kHeadHidden + 2 * kHeadFeat) + 64;
        if (cudaMalloc(reinterpret_cast<void**>(&n->head_scratch), floats * sizeof(float)) != cudaSuccess) { err = "cudaMalloc failed"; delete n; return nullptr; }
        float* p = n->head_scratch;
        n->hs.pooled = p; p += Bm * kHeadFeat;
        n->hs.xh2 = p; p += Bm * kHeadFeat;
        n->hs.d1 = p; p += Bm * kHeadFeat;
        n->hs.l1 = p; p += Bm * kHeadHidden;
        n->hs.xh3 = p; p += Bm * kHeadHidden;
        n->hs.d2 = p; p += Bm * kHeadHidden;
        n->hs.out = p; p += Bm;
        n->hs.inv2 = p; p += kHeadFeat;
        n->hs.inv3 = p; p += kHeadHidden;
        n->dl1 = p; p += Bm * kHeadHidden;
        n->dd1 = p; p += Bm * kHeadFeat;
        n->dpool = p; p += Bm * kHeadFeat;
        const size_t keep = Bm * cfg.H * cfg.W + Bm * (kHeadFeat + kHeadHidden) + static_cast<size_t>(n->n_params) + 64;
        if (cudaMalloc(reinterpret_cast<void**>(&n->x_keep), keep * sizeof(float)) != cudaSuccess) { err = "cudaMalloc failed"; delete n; return nullptr; }
        n->mask1_keep = n->x_keep + Bm * cfg.H * cfg.W;
        n->mask2_keep = n->mask1_keep + Bm * kHeadFeat;
        n->params_keep = n->mask2_keep + Bm * kHeadHidden;
    }

    // ---- GEMM launches.  Geometry is fixed at max_batch rows; a smaller batch only shortens M.
    for (size_t i = 1; i < n->convs.size(); ++i) {
        ConvHost& c = n->convs[i];
        const TPlane& x = n->levels[c.in_level];
        c.fwd_taps = forward_taps(x, c.ksize, c.stride, c.z.wp);
        GemmLaunch& L = c.fwd;
        std::memset(&L, 0, sizeof(L));
        L.weights = reinterpret_cast<const __half*>(c.w_fwd);
        L.cin = c.cin; L.cout = c.cout; L.n_wtaps = c.ksize * c.ksize;
        L.wp = c.z.wp; L.hp = c.z.hp; L.w_real = c.z.wp - n->pad; L.out_mode = OUT_PLAIN; L.wp2 = c.z.wp; L.mode = 1;
        L.stats = n->stats + c.bn.fwd_sums;
        HostJob job;
        job.taps = c.fwd_taps;
        job.out0 = c.z.base[0];
        job.out_kc_stride = c.z.kc_stride;
        if (!gemm_build_launch(L, {job}, n->tune, err)) { err = c.name + " forward: " + err; delete n; return nullptr; }
    }
    // weight-gradient launches (tensor cores, ld_wgrad.cu)
    for (size_t i = 1; i < n->convs.size(); ++i) {
        ConvHost& c = n->convs[i];
        const TPlane& x = n->levels[c.in_level];
        WgradLaunch& W = c.wg;
        std::memset(&W, 0, sizeof(W));
        W.cin = c.cin; W.cout = c.cout; W.n_taps = c.ksize * c.ksize;
        W.dz = c.dz.base[0]; W.dz_kc_stride = c.dz.kc_stride;
        const int S = 128 / c.cin, wp = c.z.wp;
        auto add_groups = [&](int seg_first, int n_real, int px_off, int kx) {
            for (int p0 = 0; p0 < n_real; p0 += S) {
                const int g = W.n_grp++;
                W.grp_seg0[g] = seg_first + p0;
                W.grp_px_off[g] = px_off;
                for (int sl = 0; sl < 8; ++sl) W.grp_tap[g][sl] = (sl < S && p0 + sl < n_real) ? (p0 + sl) * c.ksize + kx : -1;
            }
        };
        if (c.ksize == 1) {
            W.n_seg = 1;
            W.seg_src[0] = x.base[0]; W.seg_kc_stride[0] = x.kc_stride; W.seg_shift[0] = 0;
            add_groups(0, 1, 0, 0);
        } else if (c.stride == 1) {
            W.n_seg = 3;
            for (int ky = 0; ky < 3; ++ky) { W.seg_src[ky] = x.base[0]; W.seg_kc_stride[ky] = x.kc_stride; W.seg_shift[ky] = (ky - 1) * wp - 1; }
            for (int kx = 0; kx < 3; ++kx) add_groups(0, 3, kx, kx);
        } else {
            W.n_seg = 6;   // [odd-column planes: ky 0..2][even-column planes: ky 0..2]
            for (int cpi = 0; cpi < 2; ++cpi)
                for (int ky = 0; ky < 3; ++ky) {
                    const int rp = (ky - 1) & 1, di = ky == 0 ? -1 : 0, cp = cpi == 0 ? 1 : 0;
                    const int sgi = cpi * 3 + ky;
                    W.seg_src[sgi] = x.base[rp * 2 + cp]; W.seg_kc_stride[sgi] = x.kc_stride;
                    W.seg_shift[sgi] = di * wp + (cpi == 0 ? -1 : 0);
                }
            add_groups(0, 3, 0, 0);   // kx = 0: odd columns, dk = -1
            add_groups(0, 3, 1, 2);   // kx = 2: odd columns, dk = 0
            add_groups(3, 3, 0, 1);   // kx = 1: even columns
        }
        if (W.n_grp > 6 || !wgrad_plan_smem(W)) { err = c.name + ": weight-gradient launch does not fit"; delete n; return nullptr; }
    }
    // data-gradient launches: conv2 -> dh ; conv1 (+ shortcut / identity) -> gradient of the block input
    for (auto& blk : n->blocks) {
        {   // conv2: stride 1, dh(p) = sum_t W2_t^T dz2(p - shift_t)
            ConvHost& c = n->convs[blk.conv2];
            GemmLaunch& L = c.bwd;
            std::memset(&L, 0, sizeof(L));
            L.weights = reinterpret_cast<const __half*>(c.w_bwd);
            L.cin = c.cout; L.cout = c.cin; L.n_wtaps = 9; c.bwd_slabs = 9;
            L.wp = blk.dh.wp; L.hp = blk.dh.hp; L.w_real = blk.dh.wp - n->pad; L.out_mode = OUT_PLAIN; L.wp2 = blk.dh.wp; L.mode = 1;
            HostJob job;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) {
                    HostTap t{};
                    t.src = c.dz.base[0]; t.kc_stride = c.dz.kc_stride;
                    t.shift = -((ky - 1) * c.dz.wp + (kx - 1));
                    t.wslab = ky * 3 + kx;
                    job.taps.push_back(t);
                }
            job.out0 = blk.dh.base[0]; job.out_kc_stride = blk.dh.kc_stride;
            if (!gemm_build_launch(L, {job}, n->tune, err)) { err = c.name + " dgrad: " + err; delete n; return nullptr; }
            c.has_bwd = true;
        }
        {   // conv1 (+ shortcut conv or identity as a 10th slab) -> dlevels[in_level]
            ConvHost& c = n->convs[blk.conv1];
            const TPlane& dx = n->dlevels[blk.in_level];
            GemmLaunch& L = c.bwd;
            std::memset(&L, 0, sizeof(L));
            L.weights = reinterpret_cast<const __half*>(c.w_bwd);
            L.cin = c.cout; L.cout = c.cin; L.n_wtaps = 10; c.bwd_slabs = 10;
            L.wp = dx.wp; L.hp = dx.hp; L.w_real = dx.wp - n->pad; L.out_mode = OUT_PLAIN; L.wp2 = dx.wp; L.mode = 1;
            std::vector<HostJob> jobs;
            const TPlane& extra = blk.sc >= 0 ? n->convs[blk.sc].dz : blk.g;   // gradient entering through the shortcut
            if (c.stride == 1) {
                HostJob job;
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                        HostTap t{};
                        t.src = c.dz.base[0]; t.kc_stride = c.dz.kc_stride;
                        t.shift = -((ky - 1) * c.dz.wp + (kx - 1));
                        t.wslab = ky * 3 + kx;
                        job.taps.push_back(t);
                    }
                HostTap t{};
                t.src = extra.base[0]; t.kc_stride = extra.kc_stride; t.shift = 0; t.wslab = 9;
                job.taps.push_back(t);
                job.out0 = dx.base[0]; job.out_kc_stride = dx.kc_stride;
                jobs.push_back(job);
            } else {
                // parity plane (rp, cp) of dx: element (i, k) <-> real (2i + rp, 2k + cp) = 2 (r', c') + (ky - 1, kx - 1)
                for (int rp = 0; rp < 2; ++rp)
                    for (int cp = 0; cp < 2; ++cp) {
                        HostJob job;
                        for (int ky = 0; ky < 3; ++ky) {
                            if (((ky - 1) & 1) != rp) continue;
                            const int di = ky == 0 ? 1 : 0;   // ky=0: r' = i + 1 ; ky=1,2: r' = i
                            for (int kx = 0; kx < 3; ++kx) {
                                if (((kx - 1) & 1) != cp) continue;
                                const int dk = kx == 0 ? 1 : 0;
                                HostTap t{};
                                t.src = c.dz.base[0]; t.kc_stride = c.dz.kc_stride;
                                t.shift = di * c.dz.wp + dk;
                                t.wslab = ky * 3 + kx;
                                job.taps.push_back(t);
                            }
                        }
                        if (rp == 0 && cp == 0) {   // the 1x1 stride-2 shortcut reads the even/even plane only
                            HostTap t{};
                            t.src = extra.base[0]; t.kc_stride = extra.kc_stride; t.shift = 0; t.wslab = 9;
                            job.taps.push_back(t);
                        }
                        job.out0 = dx.base[rp * 2 + cp]; job.out_kc_stride = dx.kc_stride;
                        jobs.push_back(job);
                    }
            }
            if (!gemm_build_launch(L, jobs, n->tune, err)) { err = c.name + " dgrad: " + err; delete n; return nullptr; }
            c.has_bwd = true;
        }
    }
    return n;
}

void train_destroy(TrainNet* net) { delete net; }
const std::vector<TrainParamInfo>& train_param_table(const TrainNet* net) { return net->params; }
long long train_num_params(const TrainNet* net) { return net->n_params; }
const std::vector<TrainParamInfo>& train_bn_table(const TrainNet* net) { return net->bn_table; }
int train_num_bn_stats(const TrainNet* net) { return net->bn_stats_floats; }
long long train_kernel_launches(const TrainNet* net) { return net->launches; }

namespace {

#define LD_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e__); return e__; } \
    } while (0)

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }
// grid of an element-wise grid-stride kernel: enough 256-thread blocks to fill the GPU, no more than the work needs
inline unsigned ew_grid(const TrainNet* n, long long items) {
    return static_cast<unsigned>(std::max<long long>(1, std::min<long long>((items + 255) / 256, static_cast<long long>(n->num_sms) * 16)));
}

// Points a data-gradient launch at the BatchNorm whose input gradient it produces: its epilogue then accumulates that
// BatchNorm's backward sums (GemmBwdStats).  mask_mode 1: y = relu(bn(z) + residual), mask from the stored y; 2: y = relu(bn(z)).
void set_bwd_stats(TrainNet* n, GemmLaunch& L, const ConvHost& c, int mask_mode, const TPlane* y) {
    L.stats = n->stats + c.bn.bwd_sums;
    L.stats_kind = 1;
    L.bwd.z = c.z.base[0]; L.bwd.z_kc = c.z.kc_stride;
    L.bwd.y = y ? y->base[0] : nullptr; L.bwd.y_kc = y ? y->kc_stride : 0;
    L.bwd.fwd_sums = n->stats + c.bn.fwd_sums;
    L.bwd.gamma = n->params_d + c.bn.gamma_off;
    L.bwd.beta = n->params_d + c.bn.beta_off;
    L.bwd.inv_n = 1.f / static_cast<float>(c.bn.count);
    L.bwd.mask_mode = mask_mode;
}

cudaError_t run_gemm(TrainNet* n, GemmLaunch& L, const TPlane& geom, cudaStream_t stream, std::string& err) {
    const long long M = static_cast<long long>(n->B) * geom.hp * geom.wp;
    const int m_tiles = static_cast<int>((M + kTileM - 1) / kTileM);
    ++n->launches;
    LD_TRY(launch_gemm_taps(L, m_tiles, static_cast<int>(M), n->num_sms, stream));
    return cudaSuccess;
}

template <int CIN, int COUT, int NT>
cudaError_t run_wgrad_t(TrainNet* n, const ConvHost& c, float* dw, cudaStream_t stream) {
    WgradTaps t{};
    for (int i = 0; i < NT; ++i) {
        t.src[i] = static_cast<const __nv_bfloat16*>(c.fwd_taps[i].src);
        t.kc_stride[i] = c.fwd_taps[i].kc_stride;
        t.shift[i] = c.fwd_taps[i].shift;
    }
    const long long M = static_cast<long long>(n->B) * c.dz.hp * c.dz.wp;
    const long long tiles = (M + 31) / 32;
    const unsigned grid = static_cast<unsigned>(std::min<long long>(tiles, n->num_sms * 2));
    wgrad_kernel<CIN, COUT, NT><<<grid, 256, 0, stream>>>(t, c.dz.base[0], c.dz.kc_stride, M, dw);
    return cudaGetLastError();
}
cudaError_t run_wgrad(TrainNet* n, const ConvHost& c, float* dw, cudaStream_t stream) {
    ++n->launches;
    if (n->wgrad_mma) {
        WgradLaunch W = c.wg;
        W.dw = dw;
        W.M = static_cast<long long>(n->B) * c.dz.hp * c.dz.wp;
        // the last pixel tile reads up to 127 pixels past M: planes of images >= B must be zero there.  They are, unless an
        // earlier step used a larger batch -- then the tail is cleared (rare: training uses one batch size).
        if (n->B < n->max_batch_seen) {
            for (int kc = 0; kc < c.cout / 8; ++kc)
                cudaMemsetAsync(c.dz.base[0] + kc * c.dz.kc_stride + W.M * 8, 0, 128 * 16, stream);
        }
        return launch_wgrad_mma(W, n->num_sms, stream);
    }
#define LD_WG(ci, co, nt) if (c.cin == ci && c.cout == co && c.ksize * c.ksize == nt) return run_wgrad_t<ci, co, nt>(n, c, dw, stream)
    LD_WG(64, 64, 9); LD_WG(64, 32, 9); LD_WG(64, 32, 1); LD_WG(32, 32, 9); LD_WG(32, 16, 9); LD_WG(32, 16, 1);
    LD_WG(16, 16, 9); LD_WG(16, 16, 1);
#undef LD_WG
    return cudaErrorInvalidValue;
}

HeadParams head_params(const TrainNet* n, const float* p) {
    HeadParams h;
    h.g2 = p + n->head_off[0]; h.b2 = p + n->head_off[1]; h.g3 = p + n->head_off[2]; h.b3 = p + n->head_off[3];
    h.w1 = p + n->head_off[4]; h.bias1 = p + n->head_off[5]; h.w2 = p + n->head_off[6]; h.bias2 = p + n->head_off[7];
    return h;
}

}  // namespace

cudaError_t train_forward(TrainNet* n, const float* params, const float* x, int B, const float* mask1, const float* mask2,
                          float dropout_p, float* probs, float* bn_stats, cudaStream_t stream, std::string& err) {
    if (B < 2 || B > n->max_batch) { err = "batch size out of range for this training context"; return cudaErrorInvalidValue; }
    if (!(dropout_p >= 0.f && dropout_p < 1.f)) { err = "dropout rate must be in [0, 1)"; return cudaErrorInvalidValue; }
    // the backward pass needs the inputs again: keep private copies so that the caller may free its tensors after forward
    LD_TRY(cudaMemcpyAsync(n->x_keep, x, static_cast<size_t>(B) * n->cfg.H * n->cfg.W * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    LD_TRY(cudaMemcpyAsync(n->mask1_keep, mask1, static_cast<size_t>(B) * kHeadFeat * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    LD_TRY(cudaMemcpyAsync(n->mask2_keep, mask2, static_cast<size_t>(B) * kHeadHidden * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    LD_TRY(cudaMemcpyAsync(n->params_keep, params, static_cast<size_t>(n->n_params) * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    x = n->x_keep; mask1 = n->mask1_keep; mask2 = n->mask2_keep; params = n->params_keep;
    n->B = B; n->params_d = params; n->x_d = x; n->mask1_d = mask1; n->mask2_d = mask2;
    n->keep_scale = 1.f / (1.f - dropout_p);
    n->max_batch_seen = std::max(n->max_batch_seen, B);
    LD_TRY(cudaMemsetAsync(n->stats, 0, n->stats_floats * sizeof(float), stream));
    // packed bf16 weights from the fp32 master parameters (they change every optimizer step): one launch for all slabs
    if (n->pack_items == nullptr) {
        std::vector<PackItem> items;
        for (size_t i = 1; i < n->convs.size(); ++i) {
            ConvHost& c = n->convs[i];
            items.push_back({c.w_off, c.w_fwd, c.cout, c.cin, c.ksize * c.ksize, 0});
        }
        for (auto& blk : n->blocks) {
            ConvHost& c1 = n->convs[blk.conv1];
            ConvHost& c2 = n->convs[blk.conv2];
            items.push_back({c2.w_off, c2.w_bwd, c2.cout, c2.cin, 9, 1});
            items.push_back({c1.w_off, c1.w_bwd, c1.cout, c1.cin, 9, 1});
            __nv_bfloat16* slab9 = c1.w_bwd + static_cast<size_t>(9) * c1.cin * c1.cout;
            if (blk.sc >= 0) {
                ConvHost& cs = n->convs[blk.sc];
                items.push_back({cs.w_off, slab9, cs.cout, cs.cin, 1, 1});
            } else {
                items.push_back({-1, slab9, c1.cin, c1.cin, 1, 0});
            }
        }
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cap);
        if (cap != cudaStreamCaptureStatusNone) { err = "first training step inside a CUDA graph capture: run one eager step first"; return cudaErrorInvalidValue; }
        LD_TRY(cudaMalloc(reinterpret_cast<void**>(&n->pack_items), items.size() * sizeof(PackItem)));
        LD_TRY(cudaMemcpy(n->pack_items, items.data(), items.size() * sizeof(PackItem), cudaMemcpyHostToDevice));
        n->n_pack_items = static_cast<int>(items.size());
    }
    pack_all_kernel<<<dim3(36, n->n_pack_items), 256, 0, stream>>>(static_cast<const PackItem*>(n->pack_items), params);
    ++n->launches;
    LD_TRY(cudaGetLastError());

    // stem
    {
        ConvHost& c = n->convs[0];
        c.bn.count = static_cast<long long>(B) * n->cfg.H * n->cfg.W;
        stem_fwd_kernel<<<ew_grid(n, c.bn.count * 8), 256, 0, stream>>>(x, B, n->cfg.H, n->cfg.W, params + c.w_off, c.z, n->stats + c.bn.fwd_sums);
        const long long work = c.bn.count * 8;
        bn_apply_kernel<<<ew_grid(n, work), 256, 0, stream>>>(c.z, bn_ref(*n, c.bn), 1, 0, c.z, bn_ref(*n, c.bn), n->levels[0], B);
        n->launches += 2;
    }
    for (auto& blk : n->blocks) {
        ConvHost& c1 = n->convs[blk.conv1];
        ConvHost& c2 = n->convs[blk.conv2];
        const TPlane& xin = n->levels[blk.in_level];
        const TPlane& h = n->levels[blk.h_level];
        const TPlane& out = n->levels[blk.out_level];
        c1.bn.count = c2.bn.count = static_cast<long long>(B) * h.H * h.W;
        if (cudaError_t e = run_gemm(n, c1.fwd, c1.z, stream, err)) return e;
        bn_apply_kernel<<<ew_grid(n, c1.bn.count * (c1.cout / 8)), 256, 0, stream>>>(c1.z, bn_ref(*n, c1.bn), 1, 0, c1.z, bn_ref(*n, c1.bn), h, B);
        if (cudaError_t e = run_gemm(n, c2.fwd, c2.z, stream, err)) return e;
        if (blk.sc >= 0) {
            ConvHost& cs = n->convs[blk.sc];
            cs.bn.count = c2.bn.count;
            if (cudaError_t e = run_gemm(n, cs.fwd, cs.z, stream, err)) return e;
            bn_apply_kernel<<<ew_grid(n, c2.bn.count * (c2.cout / 8)), 256, 0, stream>>>(c2.z, bn_ref(*n, c2.bn), 1, 2, cs.z, bn_ref(*n, cs.bn), out, B);
        } else {
            bn_apply_kernel<<<ew_grid(n, c2.bn.count * (c2.cout / 8)), 256, 0, stream>>>(c2.z, bn_ref(*n, c2.bn), 1, 1, xin, bn_ref(*n, c2.bn), out, B);
        }
        n->launches += 2;
    }
    LD_TRY(cudaGetLastError());
    // head (+ its two BatchNorm1d statistics straight into bn_stats)
    const int head_stat0 = n->head_bn2.stat_out;
    head_pool_kernel<<<blocks_for(static_cast<long long>(B) * kHeadFeat, 256), 256, 0, stream>>>(n->levels.back(), B, n->hs.pooled);
    head_fwd_kernel<<<1, kHeadThreads, 0, stream>>>(B, head_params(n, params), mask1, mask2, n->keep_scale, n->hs, probs,
                                           bn_stats + head_stat0);
    ++n->launches;
    ++n->launches;
    LD_TRY(cudaGetLastError());
    // conv BatchNorm statistics (mean incl. conv bias, biased variance) for the caller's running-statistics update
    {
        // the table depends on the batch size only: uploaded (synchronously, pageable host memory) when B changes, so that the
        // steady-state step has no host->device copy and can be captured into a CUDA graph
        std::vector<BnFinalizeItem>& items = n->bn_items_host;
        if (n->bn_items_batch != B) {
            items.clear();
            for (const auto& c : n->convs)
                items.push_back({c.bn.fwd_sums, c.bn.stat_out, c.bn.C, c.b_off, 1.f / static_cast<float>(c.bn.count)});
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(stream, &cap);
            if (cap != cudaStreamCaptureStatusNone) { err = "first training step of a batch size inside a CUDA graph capture: run one eager step first"; return cudaErrorInvalidValue; }
            LD_TRY(cudaStreamSynchronize(stream));
            LD_TRY(cudaMemcpy(n->bn_items, items.data(), items.size() * sizeof(BnFinalizeItem), cudaMemcpyHostToDevice));
            n->bn_items_batch = B;
        }
        bn_finalize_kernel<<<static_cast<unsigned>(items.size()), 64, 0, stream>>>(n->bn_items, static_cast<int>(items.size()), n->stats, params, bn_stats);
        ++n->launches;
        LD_TRY(cudaGetLastError());
    }
    return cudaSuccess;
}

// Debug: sum of |value| over every z plane, every activation level, every dz plane and every input-gradient level of the
// last call (host copies; slow).  Returns the number of doubles written.
int train_debug_checksums(TrainNet* n, double* out, int cap) {
    cudaDeviceSynchronize();
    std::vector<__nv_bfloat16> h;
    int k = 0;
    auto sum_plane = [&](const TPlane& t) {
        double acc = 0.0;
        const long long pixels = static_cast<long long>(n->B) * t.hp * t.wp;
        h.resize(static_cast<size_t>(pixels) * 8);
        for (int q = 0; q < (t.quad ? 4 : 1); ++q)
            for (int kc = 0; kc < t.C / 8; ++kc) {
                cudaMemcpy(h.data(), t.base[q] + kc * t.kc_stride, h.size() * 2, cudaMemcpyDeviceToHost);
                for (size_t i = 0; i < h.size(); ++i) acc += std::fabs(static_cast<double>(__bfloat162float(h[i])));
            }
        if (k < cap) out[k] = acc;
        ++k;
    };
    for (const auto& c : n->convs) sum_plane(c.z);
    for (const auto& l : n->levels) sum_plane(l);
    for (const auto& c : n->convs) sum_plane(c.dz);
    for (const auto& l : n->dlevels) sum_plane(l);
    return k;
}

// Debug: one tensor of the last step as dense (B, C, H, W) fp32 on the host.  kind: 0 conv output z, 1 activation level y,
// 2 conv-output gradient dz, 3 gradient wrt a level, 4 block g (gradient entering the residual sum), 5 block dh.
// Returns the number of floats (also when out is null), or -1.
long long train_debug_read(TrainNet* n, int kind, int index, float* out, int* dims4) {
    const TPlane* t = nullptr;
    if (kind == 0 && index < static_cast<int>(n->convs.size())) t = &n->convs[index].z;
    if (kind == 1 && index < static_cast<int>(n->levels.size())) t = &n->levels[index];
    if (kind == 2 && index < static_cast<int>(n->convs.size())) t = &n->convs[index].dz;
    if (kind == 3 && index < static_cast<int>(n->dlevels.size())) t = &n->dlevels[index];
    if (kind == 4 && index < static_cast<int>(n->blocks.size())) t = &n->blocks[index].g;
    if (kind == 5 && index < static_cast<int>(n->blocks.size())) t = &n->blocks[index].dh;
    if (!t || index < 0 || n->B <= 0) return -1;
    if (dims4) { dims4[0] = n->B; dims4[1] = t->C; dims4[2] = t->H; dims4[3] = t->W; }
    const long long total = static_cast<long long>(n->B) * t->C * t->H * t->W;
    if (!out) return total;
    cudaDeviceSynchronize();
    const long long pixels = static_cast<long long>(n->B) * t->hp * t->wp;
    std::vector<__nv_bfloat16> h(static_cast<size_t>(pixels) * 8);
    for (int q = 0; q < (t->quad ? 4 : 1); ++q)
        for (int kc = 0; kc < t->C / 8; ++kc) {
            cudaMemcpy(h.data(), t->base[q] + kc * t->kc_stride, h.size() * 2, cudaMemcpyDeviceToHost);
            for (int b = 0; b < n->B; ++b)
                for (int r = 0; r < t->H; ++r)
                    for (int c = 0; c < t->W; ++c) {
                        int which; long long p;
                        if (!t->quad) { which = 0; p = (static_cast<long long>(b) * t->hp + 1 + r) * t->wp + 1 + c; }
                        else { which = (r & 1) * 2 + (c & 1); p = (static_cast<long long>(b) * t->hp + 1 + (r >> 1)) * t->wp + 1 + (c >> 1); }
                        if (which != q) continue;
                        for (int e = 0; e < 8; ++e)
                            out[((static_cast<long long>(b) * t->C + kc * 8 + e) * t->H + r) * t->W + c] = __bfloat162float(h[p * 8 + e]);
                    }
        }
    return total;
}

// ------------------------------------------------------------------------------------------------ K8: clip + Adam
// torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam(...).step() (train.py:292-295,336) on ONE flat
// fp32 parameter / gradient vector: pass 1 = global sum of squares, pass 2 = scale by min(1, max_norm / (norm + 1e-6)) and the
// Adam update with PyTorch's bias-correction arithmetic.
namespace {
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    float acc = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        acc = fmaf(g[i], g[i], acc);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        atomicAdd(out, t);
    }
}
// step_d != nullptr: the step count lives in device memory (it was incremented before this launch), so that a captured CUDA
// graph of the training step replays with the right bias corrections.
__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                 const float* __restrict__ sumsq, float max_norm, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                 float* __restrict__ norm_out, const long long* __restrict__ step_d) {
    if (step_d != nullptr) {
        const float st = static_cast<float>(*step_d);
        bc1 = 1.f - powf(b1, st);
        bc2_sqrt = sqrtf(1.f - powf(b2, st));
    }
    const float norm = sqrtf(*sumsq);
    const float scale = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
    if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    const float step_size = lr / bc1;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * scale;
        const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
        const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}
}  // namespace

cudaError_t clip_adam_step(float* params, const float* grads, float* m, float* v, long long n, float max_norm, float lr, float b1,
                           float b2, float eps, long long step, float* scratch2, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(scratch2, 0, sizeof(float), stream);
    if (e != cudaSuccess) return e;
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 592)));
    sumsq_kernel<<<grid, 256, 0, stream>>>(grads, n, scratch2);
    const float bc1 = 1.f - std::pow(b1, static_cast<float>(step));
    const float bc2 = 1.f - std::pow(b2, static_cast<float>(step));
    clip_adam_kernel<<<grid, 256, 0, stream>>>(params, grads, m, v, n, scratch2, max_norm, lr, b1, b2, eps, bc1, std::sqrt(bc2), scratch2 + 1, nullptr);
    return cudaGetLastError();
}

namespace {
__global__ void step_increment_kernel(long long* step, float* sumsq) { *step += 1; *sumsq = 0.f; }
}  // namespace

cudaError_t clip_adam_step_dev(float* params, const float* grads, float* m, float* v, long long n, float max_norm, float lr, float b1,
                               float b2, float eps, long long* step_d, float* scratch2, cudaStream_t stream) {
    step_increment_kernel<<<1, 1, 0, stream>>>(step_d, scratch2);
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>((n + 255) / 256, 592)));
    sumsq_kernel<<<grid, 256, 0, stream>>>(grads, n, scratch2);
    clip_adam_kernel<<<grid, 256, 0, stream>>>(params, grads, m, v, n, scratch2, max_norm, lr, b1, b2, eps, 1.f, 1.f, scratch2 + 1, step_d);
    return cudaGetLastError();
}

cudaError_t train_backward(TrainNet* n, const float* dprobs, float* grads, cudaStream_t stream, std::string& err) {
    if (n->B <= 0 || n->params_d == nullptr) { err = "train_backward without a preceding train_forward"; return cudaErrorInvalidValue; }
    const int B = n->B;
    const float* params = n->params_d;
    LD_TRY(cudaMemsetAsync(grads, 0, n->n_params * sizeof(float), stream));
    HeadGrads hg;
    hg.g2 = grads + n->head_off[0]; hg.b2 = grads + n->head_off[1]; hg.g3 = grads + n->head_off[2]; hg.b3 = grads + n->head_off[3];
    hg.w1 = grads + n->head_off[4]; hg.bias1 = grads + n->head_off[5]; hg.w2 = grads + n->head_off[6]; hg.bias2 = grads + n->head_off[7];
    head_bwd_kernel<<<1, kHeadThreads, 0, stream>>>(B, head_params(n, params), n->mask1_d, n->mask2_d, n->keep_scale, n->hs, dprobs,
                                           n->dl1, n->dd1, n->dpool, hg);
    head_unpool_kernel<<<blocks_for(static_cast<long long>(B) * n->dy_last.H * n->dy_last.W * (n->dy_last.C / 8), 256), 256, 0, stream>>>(
        n->dy_last, B, n->dpool);
    ++n->launches;
    ++n->launches;
    LD_TRY(cudaGetLastError());

    // weight gradient of conv c on the side stream, after everything queued on `stream` so far (its dz is complete)
    size_t n_fork = 0;
    auto fork_side = [&]() -> cudaStream_t {
        if (!n->side_wgrad || n_fork >= n->fork_events.size()) return stream;
        cudaEvent_t ev = n->fork_events[n_fork++];
        if (cudaEventRecord(ev, stream) != cudaSuccess || cudaStreamWaitEvent(n->side, ev, 0) != cudaSuccess) return stream;
        return n->side;
    };
    auto wgrad = [&](const ConvHost& c) -> cudaError_t { return run_wgrad(n, c, grads + c.w_off, fork_side()); };
    auto bn_backward = [&](const TPlane& dy, const TPlane& y, int relu, ConvHost& c, int write_g, const TPlane& g_out) {
        float* sums = n->stats + c.bn.bwd_sums;
        const long long work = c.bn.count * (c.cout / 8);
        if (!c.bn.reduce_fused) {   // (fused: the dgrad GEMM that produced dy already accumulated the sums in its epilogue)
            const unsigned grid_r = static_cast<unsigned>(std::min<long long>((work + 255) / 256, n->num_sms * 8));
            bn_bwd_reduce_kernel<<<grid_r, 256, 0, stream>>>(dy, y, relu, c.z, bn_ref(*n, c.bn), B, sums);
            ++n->launches;
        }
        bn_bwd_apply_kernel<<<ew_grid(n, work), 256, 0, stream>>>(dy, y, relu, c.z, bn_ref(*n, c.bn), sums, B, c.dz, write_g, g_out,
                                                                      grads + c.bn.gamma_off, grads + c.bn.beta_off);
        ++n->launches;
    };
    for (auto& c : n->convs) c.bn.reduce_fused = false;

    for (int bi = static_cast<int>(n->blocks.size()) - 1; bi >= 0; --bi) {
        BlockHost& blk = n->blocks[bi];
        ConvHost& c1 = n->convs[blk.conv1];
        ConvHost& c2 = n->convs[blk.conv2];
        const TPlane& out = n->levels[blk.out_level];
        const TPlane& dout = n->dlevels[blk.out_level];
        const TPlane& h = n->levels[blk.h_level];
        // y = relu(bn2(z2) + shortcut): g = dy * [y > 0] is the gradient of both summands
        bn_backward(dout, out, 1, c2, 1, blk.g);
        if (cudaError_t e = wgrad(c2)) { err = "wgrad " + c2.name; return e; }
        // dh = conv2^T(dz2); its epilogue also reduces bn1's backward sums (h = relu(bn1(z1)): mask recomputed from z1)
        if (n->fuse_bwd_stats) { set_bwd_stats(n, c2.bwd, c1, 2, nullptr); c1.bn.reduce_fused = true; }
        else { c2.bwd.stats = nullptr; c2.bwd.stats_kind = 0; }
        if (cudaError_t e = run_gemm(n, c2.bwd, blk.dh, stream, err)) return e;
        bn_backward(blk.dh, h, 2, c1, 0, blk.g);   // h = relu(bn1(z1)): the mask comes from z1, h is not read
        if (cudaError_t e = wgrad(c1)) { err = "wgrad " + c1.name; return e; }
        if (blk.sc >= 0) {
            ConvHost& cs = n->convs[blk.sc];
            bn_backward(blk.g, blk.g, 0, cs, 0, blk.g);
            if (cudaError_t e = wgrad(cs)) { err = "wgrad " + cs.name; return e; }
        }
        // gradient of the block input: conv1^T on dz1 plus the shortcut path, one GEMM launch.  With a stride-1 conv1 its output
        // has the geometry of the previous BatchNorm's z plane, so that BatchNorm's backward sums ride in the epilogue too
        // (previous block: y = relu(bn2(z2) + shortcut), mask from the stored y; first block: the stem, mask from z0).
        c1.bwd.stats = nullptr; c1.bwd.stats_kind = 0;
        if (n->fuse_bwd_stats && c1.stride == 1) {
            if (bi > 0) {
                ConvHost& pc2 = n->convs[n->blocks[bi - 1].conv2];
                set_bwd_stats(n, c1.bwd, pc2, 1, &n->levels[blk.in_level]);
                pc2.bn.reduce_fused = true;
            } else {
                set_bwd_stats(n, c1.bwd, n->convs[0], 2, nullptr);
                n->convs[0].bn.reduce_fused = true;
            }
        }
        if (cudaError_t e = run_gemm(n, c1.bwd, n->dlevels[blk.in_level], stream, err)) return e;
    }
    {   // stem
        ConvHost& c = n->convs[0];
        bn_backward(n->dlevels[0], n->levels[0], 2, c, 0, c.dz);
        stem_wgrad_kernel<<<n->num_sms * 8, 256, 0, fork_side()>>>(n->x_d, B, n->cfg.H, n->cfg.W, c.dz, grads + c.w_off);
        ++n->launches;
    }
    LD_TRY(cudaGetLastError());
    if (n_fork > 0) {   // join: the caller's stream continues (optimiser) only after every weight gradient has landed
        LD_TRY(cudaEventRecord(n->join_event, n->side));
        LD_TRY(cudaStreamWaitEvent(stream, n->join_event, 0));
    }
    return cudaSuccess;
}

}  // namespace ld
