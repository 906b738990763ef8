// Floor of the backward's per-hit row reductions: H hits, each adds a 256-byte row to one of N rows of a [N,64] float buffer.
// Variants: one TMA bulk reduction per hit (what trace_backward_flat_kernel does), 16 x red.v4.f32 per hit, and bulk reductions of
// 64 / 128 bytes (how does the cost scale with the row size).  nvcc -arch=sm_100a -O3 -o red_floor red_floor.cu && ./red_floor
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
constexpr int BROW = 68;
template <int MODE, int BYTES>
__global__ void __launch_bounds__(128) k(const int *__restrict__ idx, int64_t H, float *__restrict__ buf) {
    __shared__ __align__(16) float rows[4][32 * BROW];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *r = rows[w];
    const int64_t warp0 = ((int64_t)blockIdx.x * 4 + w) * 128;   // a warp handles 4 rounds of 32 hits like a 32-ray group
    for (int round = 0; round < 4; ++round) {
        const int64_t i = warp0 + round * 32 + lane;
        if (i >= H) break;
        const int g = __ldg(idx + i);
        if (MODE == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        float4 *row = reinterpret_cast<float4 *>(r + lane * BROW);
#pragma unroll
        for (int v = 0; v < 16; ++v) row[v] = make_float4(1.f, 2.f, 3.f, (float)v);
        if (MODE == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const uint32_t src = (uint32_t)__cvta_generic_to_shared(r + lane * BROW);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                         :: "l"(buf + (size_t)g * 64), "r"(src), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else {
            __syncwarp();
            const int col = lane & 15;
            for (int j = 0; j < 32; j += 2) {
                const int rowi = j + (lane >> 4);
                const int g_r = __shfl_sync(0xffffffffu, g, rowi);
                if (col * 16 < BYTES) {
                    const float4 v = *reinterpret_cast<const float4 *>(r + rowi * BROW + 4 * col);
                    atomicAdd(reinterpret_cast<float4 *>(buf + (size_t)g_r * 64) + col, v);
                }
            }
        }
    }
    if (MODE == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
template <int MODE, int BYTES>
float run(const int *idx, int64_t H, float *buf) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((H + 511) / 512);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0); k<MODE, BYTES><<<grid, 128>>>(idx, H, buf); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    const int64_t H = 15500000; const int N = 300000;
    std::vector<int> h(H);
    srand(1);
    // hits of neighbouring rays land on nearby surfels only loosely: random rows, and a variant with runs of 4 equal rows
    for (int64_t i = 0; i < H; ++i) h[i] = (int)(((uint64_t)rand() * 32768u + rand()) % N);
    int *idx; float *buf;
    cudaMalloc(&idx, H * 4); cudaMalloc(&buf, (size_t)N * 256); cudaMemset(buf, 0, (size_t)N * 256);
    cudaMemcpy(idx, h.data(), H * 4, cudaMemcpyHostToDevice);
    printf("random rows: bulk256 %.3f ms | bulk128 %.3f | bulk64 %.3f | red.v4 x16 %.3f | red.v4 x8 %.3f\n", run<0, 256>(idx, H, buf),
           run<0, 128>(idx, H, buf), run<0, 64>(idx, H, buf), run<1, 256>(idx, H, buf), run<1, 128>(idx, H, buf));
    for (int64_t i = 0; i < H; ++i) h[i] = (int)((i / 4 * 2654435761u) % N);
    cudaMemcpy(idx, h.data(), H * 4, cudaMemcpyHostToDevice);
    printf("runs of 4:   bulk256 %.3f ms | red.v4 x16 %.3f\n", run<0, 256>(idx, H, buf), run<1, 256>(idx, H, buf));
    const int N2 = 1000000; float *buf2; cudaMalloc(&buf2, (size_t)N2 * 256);
    for (int64_t i = 0; i < H; ++i) h[i] = (int)(((uint64_t)rand() * 32768u + rand()) % N2);
    cudaMemcpy(idx, h.data(), H * 4, cudaMemcpyHostToDevice);
    printf("1M rows (256 MB > L2): bulk256 %.3f ms | red.v4 x16 %.3f\n", run<0, 256>(idx, H, buf2), run<1, 256>(idx, H, buf2));
    return 0;
}
