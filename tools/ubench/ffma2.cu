// micro-benchmark: issue rate of FFMA vs FFMA2 / FADD2 / FMUL2 (packed fp32x2, sm_100) per SM sub-partition
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc, float s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(s, s * 1.0001f), c = make_float2(0.5f, 0.25f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }     // 2 FFMA
            else if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);                                       // 1 FFMA2
            else if (MODE == 2) a[i] = __fadd2_rn(a[i], c);
            else a[i] = __fmul2_rn(a[i], m);
        }
    }
    long long t1 = clock64();
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096, warps = 32;
    for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, warps * 32>>>(out, iters, cyc, 0.999f); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_sub = (double)h / (iters * 8.0 * (warps / 4));   // cycles per warp-level fp32x2 element pair per sub-partition
    printf("%s: %.3f cycles per (warp x 2 fp32 ops) per SM sub-partition\n", name, per_sub);
    cudaFree(out); cudaFree(cyc);
}
int main() { run<0>("2 x FFMA "); run<1>("1 x FFMA2"); run<2>("1 x FADD2"); run<3>("1 x FMUL2"); return 0; }
