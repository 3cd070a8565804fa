// Probe for the next round's plane layout, not part of the product:
// does a tcgen05 K-major SWIZZLE_128B A operand that starts at an ARBITRARY pixel (row) offset inside a pre-swizzled,
// pixel-major plane image multiply correctly, and which descriptor base_offset does it need?
//
//   plane image in smem: pixel p (128 bytes = 64 fp16 channels), 16-byte chunk j stored at chunk position j ^ (p & 7)
//   A tile for output pixels [s, s + 128): descriptor start = base + s * 128 (+ ks * 32 for K step ks), SBO = 1024, SW128
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I laughter_detection_icsi_b200/csrc -o tools/probes/umma_swizzle_probe tools/probes/umma_swizzle_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "ld_ptx.cuh"

using namespace ld;

constexpr int kPix = 256, kN = 64;
// Swizzle<B,4,3> on byte addresses: chunk (16 B) index bits ^= row-address bits [7, 7+B); B = 3 / 2 / 1 for 128 / 64 / 32-byte rows.
__host__ __device__ inline int swz_chunk(int row, int chunk, int C) {
    const int R = 2 * C;                       // bytes per row
    const int addr = row * R + chunk * 16;
    const int B = C == 64 ? 3 : C == 32 ? 2 : 1;
    return ((addr >> 4) ^ ((addr >> 7) & ((1 << B) - 1))) & (R / 16 - 1);
}

__host__ __device__ inline int a_val(int p, int c) { return ((p * 7 + c * 3) % 17) - 8; }
__host__ __device__ inline int w_val(int n, int k) { return ((n * 5 + k * 11) % 13) - 6; }

template <int kC>
__global__ void __launch_bounds__(128) probe_kernel(const __half* a_img, const __half* w_img, int shift, int mode, float* d_out) {
    constexpr int R = 2 * kC;   // bytes per pixel / per weight row
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // kPix * R bytes
    uint8_t* sB = smem + kPix * 128;          // kN * R bytes (1024-aligned)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kN * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kPix * R / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < kN * R / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(w_img)[i];
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(slot), 64); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 4) | ((kN >> 3) << 17) | ((128u >> 4) << 24);   // f32 D, f16 A/B, K-major, N = 64, M = 128
        for (int ks = 0; ks < kC / 16; ++ks) {   // K steps of 16 channels = 32 bytes inside the row
            const uint32_t a_addr = smem_u32(sA) + shift * R + ks * 32;
            const uint32_t b_addr = smem_u32(sB) + ks * 32;
            auto desc = [&](uint32_t addr) {
                uint64_t d = (addr >> 4) & 0x3FFFu;
                d |= static_cast<uint64_t>(1) << 16;                  // LBO (unused for swizzled K-major)
                d |= static_cast<uint64_t>((8 * R) >> 4) << 32;       // SBO: 8 rows
                d |= static_cast<uint64_t>(1) << 46;                  // descriptor version
                uint32_t bo = 0;
                if (mode == 1) bo = (addr >> 7) & 7u;                 // start row inside the 8-row swizzle atom
                d |= static_cast<uint64_t>(bo) << 49;
                d |= static_cast<uint64_t>(kC == 64 ? 2 : kC == 32 ? 4 : 6) << 61;   // SWIZZLE_128B / 64B / 32B
                return d;
            };
            umma_f16_ss_pred(tmem, desc(a_addr), desc(b_addr), idesc, ks > 0 ? 1u : 0u, true);
        }
        umma_commit_pred(smem_u32(bar), true);
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    uint32_t v[64];
    tmem_ld_cols<64>(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_wait_ld();
    for (int n = 0; n < kN; ++n) d_out[(warp * 32 + lane) * kN + n] = __uint_as_float(v[n]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

template <int kC>
int run() {
    std::vector<__half> a(kPix * kC), w(kN * kC);
    for (int p = 0; p < kPix; ++p)
        for (int c = 0; c < kC; ++c) a[p * kC + swz_chunk(p, c >> 3, kC) * 8 + (c & 7)] = __float2half(static_cast<float>(a_val(p, c)));
    for (int n = 0; n < kN; ++n)
        for (int k = 0; k < kC; ++k) w[n * kC + swz_chunk(n, k >> 3, kC) * 8 + (k & 7)] = __float2half(static_cast<float>(w_val(n, k)));
    __half *a_d, *w_d;
    float* d_d;
    cudaMalloc(&a_d, a.size() * 2); cudaMalloc(&w_d, w.size() * 2); cudaMalloc(&d_d, 128 * kN * 4);
    cudaMemcpy(a_d, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(w_d, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = kPix * 128 + kN * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel<kC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    std::vector<float> d(128 * kN);
    const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 24, 46, 93};
    for (int mode = 0; mode < 2; ++mode)
        for (int s : shifts) {
            cudaMemset(d_d, 0, d.size() * 4);
            probe_kernel<kC><<<1, 128, smem>>>(a_d, w_d, s, mode, d_d);
            const cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("C %d mode %d shift %d: CUDA error %s\n", kC, mode, s, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(d.data(), d_d, d.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0; int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < kN; ++n) {
                    double ref = 0;
                    for (int k = 0; k < kC; ++k) ref += static_cast<double>(a_val(s + m, k)) * w_val(n, k);
                    const double err = fabs(ref - d[m * kN + n]);
                    worst = err > worst ? err : worst;
                    bad += err > 0.5;
                }
            printf("C = %2d (%3d-byte rows)  base_offset %s  shift %3d px: max |err| %.1f, %d of %d outputs wrong\n", kC, 2 * kC,
                   mode ? "(addr>>7)&7" : "0          ", s, worst, bad, 128 * kN);
        }
    cudaFree(a_d); cudaFree(w_d); cudaFree(d_d);
    return 0;
}

int main() { return run<64>() || run<32>() || run<16>(); }
