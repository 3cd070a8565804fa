// Probe: how much HBM bandwidth does the conv GEMM's ACCESS PATTERN allow, independent of its MMA / barrier logic?
// One persistent CTA per SM: a producer warp streams "operand groups" of a tile into a shared-memory ring with bulk copies,
// 8 "epilogue" warps write the tile's output planes with 16-byte stores -- the data movement of block1.1.conv2 and nothing else.
//   layout 0 (today):  planes are channel-chunk planar [C/8][pixels][8]: a group = 8 copies of 136 px x 16 B (2 176 B each),
//                      an output = 8 chunk streams, a warp store instruction writes 512 contiguous bytes
//   layout 1 (pixel-major): [pixels][C]: a group = ONE copy of 136 px x 128 B (17 408 B),
//                      an output = one stream, a warp writes 32 px x 64 B halves of 128-byte rows
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o plane_stream_probe plane_stream_probe.cu ; run: ./plane_stream_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kStages = 8, kGroupPx = 136, kTilePx = 128, kC = 64;
constexpr int kGroupBytes = kGroupPx * kC * 2;   // 17 408

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Params {
    const uint8_t* in[16];    // input planes
    uint8_t* out[16];         // output planes
    int n_in, n_out, layout, do_loads, do_stores;
    long long pixels;         // per plane
    int tiles;
};

__global__ void __launch_bounds__(320, 1) probe_kernel(Params P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kGroupBytes);
    const uint32_t full = smem_u32(bars), empty = smem_u32(bars + kStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long plane_bytes = P.pixels * kC * 2, chunk_stride = P.pixels * 16;
    if (warp == 0) {            // producer
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < P.tiles && P.do_loads; tile += gridDim.x) {
            const long long p0 = static_cast<long long>(tile) * kTilePx;
            for (int g = 0; g < P.n_in; ++g) {
                if (lane == 0) { mbar_wait(empty + 8 * stage, phase ^ 1); mbar_expect_tx(full + 8 * stage, kGroupBytes); }
                __syncwarp();
                const uint32_t dst = smem_u32(smem + stage * kGroupBytes);
                if (P.layout == 0) {
                    if (lane < 8) bulk_g2s(dst + lane * (kGroupPx * 16), P.in[g] + lane * chunk_stride + p0 * 16, kGroupPx * 16, full + 8 * stage);
                } else {
                    if (lane == 0) bulk_g2s(dst, P.in[g] + p0 * (kC * 2), kGroupBytes, full + 8 * stage);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {     // consumer stand-in for the MMA warp: frees every stage as soon as it has landed
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < P.tiles && P.do_loads; tile += gridDim.x)
            for (int g = 0; g < P.n_in; ++g) {
                if (lane == 0) { mbar_wait(full + 8 * stage, phase); mbar_arrive(empty + 8 * stage); }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
    } else if (P.do_stores) {   // 8 epilogue warps: warp e writes pixel quadrant e % 4, channel half e / 4 of every output
        const int e = warp - 2, q = e & 3, half = e >> 2;
        for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
            const long long p = static_cast<long long>(tile) * kTilePx + q * 32 + lane;
            const uint4 v = make_uint4(tile, lane, e, 1);
            for (int o = 0; o < P.n_out; ++o)
                for (int kc = 0; kc < 4; ++kc) {
                    const int chunk = half * 4 + kc;
                    uint8_t* dst = P.layout == 0 ? P.out[o] + chunk * chunk_stride + p * 16
                                                 : P.out[o] + p * (kC * 2) + ((chunk ^ static_cast<int>(p & 7)) * 16);
                    *reinterpret_cast<uint4*>(dst) = v;
                }
        }
    }
    (void)plane_bytes;
}

int main(int argc, char** argv) {
    const long long pixels = 32868ll * 46;          // one chunk of block1 planes
    const int tiles = static_cast<int>((pixels - kGroupPx) / kTilePx);
    const long long plane_bytes = pixels * kC * 2;
    Params P{};
    P.n_in = 16; P.n_out = 11; P.pixels = pixels; P.tiles = tiles;   // block1.1.conv2: 9 inputs + 7 residuals, 11 outputs
    for (int i = 0; i < 16; ++i) { cudaMalloc(&P.in[i], plane_bytes); cudaMemset((void*)P.in[i], 1, plane_bytes); cudaMalloc(&P.out[i], plane_bytes); }
    const size_t smem = kStages * kGroupBytes + 2 * kStages * 8 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[3] = {"loads only", "stores only", "loads + stores"};
    for (int layout = 0; layout < 2; ++layout)
        for (int mode = 0; mode < 3; ++mode) {
            P.layout = layout; P.do_loads = mode != 1; P.do_stores = mode != 0;
            const double bytes = (P.do_loads ? 1.0 * P.n_in * tiles * kGroupBytes : 0.0) + (P.do_stores ? 1.0 * P.n_out * tiles * kTilePx * kC * 2 : 0.0);
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(a);
                probe_kernel<<<148, 320, smem>>>(P);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (rep > 0 && ms < best) best = ms;
            }
            cudaError_t e = cudaGetLastError();
            std::printf("layout %d (%s) %-15s %7.3f ms  %7.1f GB/s  %s\n", layout, layout ? "pixel-major, 17 KB copies" : "chunk-planar, 2 KB copies", names[mode],
                        best, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
