// micro-benchmark: throughput / latency of mma.sync.m16n8k8 TF32 (HMMA.1688.F32.TF32) on sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(float* out, int iters, long long* cyc) {
    float d[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma(d[c], a0, a1, a2, a3, b0, b1);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CHAINS>
void run(int warps_per_sm, const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<CHAINS><<<148, warps_per_sm * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    k<CHAINS><<<148, warps_per_sm * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp = (double)h / (iters * CHAINS);
    printf("%s warps/SM %2d chains %d: %.2f cycles per HMMA per warp, %.2f cycles per HMMA per SM sub-partition (4 per SM)\n",
           name, warps_per_sm, CHAINS, per_warp, per_warp / (warps_per_sm / 4.0 > 1 ? warps_per_sm / 4.0 : 1));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(4, "latency ");     // one warp per scheduler, dependent chain
    run<4>(4, "ilp4    ");
    run<1>(16, "tlp4    ");
    run<1>(32, "tlp8    ");
    run<4>(32, "tlp8ilp4");
    return 0;
}
