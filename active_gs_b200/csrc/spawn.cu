// spawn.cu -- device-side pieces of the per-keyframe spawn step.
//
// ags_smooth_depth replaces get_smooth_depth (/root/reference/utils/operations.py:161-169): the
// reference copies the depth image to the host, runs cv2.bilateralFilter(depth, 15, 0.5, 20) on the
// CPU (~30 ms at 640x480) and copies it back.  This kernel restates OpenCV's float32 bilateral
// filter on the GPU (~30 us):
//   * invalid pixels (depth < 0) enter the filter as 0 (np.nan_to_num of NaN), and are -1 in the output;
//   * radius = d/2, taps with sqrt(i^2+j^2) > radius are excluded (circular support);
//   * border = BORDER_REFLECT_101;
//   * space weight exp(-r^2 / (2 sigma_space^2));
//   * colour weight = OpenCV's 4096-bin exp LUT with linear interpolation over
//     |v - v0| * 4096 / (max - min), max/min taken over the whole (zero-filled) image;
//   * if max - min < FLT_EPSILON the image is copied unchanged.
#include <float.h>
#include "ags_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
depth_minmax_kernel(const float* __restrict__ depth, int n, unsigned* __restrict__ mm) {
    float lo = FLT_MAX, hi = 0.f;                            // values are >= 0 after the zero fill
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float d = __ldg(depth + i);
        const float v = (d < 0.f || d != d) ? 0.f : d;
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {                           // non-negative floats order like their bit patterns
        atomicMin(mm + 0, __float_as_uint(lo));
        atomicMax(mm + 1, __float_as_uint(hi));
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

constexpr int BF_TILE = 16;
constexpr int BF_MAXR = 16;

__global__ void __launch_bounds__(BF_TILE * BF_TILE)
bilateral_kernel(const float* __restrict__ depth, float* __restrict__ out, int H, int W, int radius,
                 float color_coeff, float space_coeff, const unsigned* __restrict__ mm) {
    extern __shared__ float s_img[];                         // (TILE + 2r)^2 zero-filled values
    const int span = BF_TILE + 2 * radius;
    const int x0 = blockIdx.x * BF_TILE - radius, y0 = blockIdx.y * BF_TILE - radius;
    for (int k = threadIdx.y * BF_TILE + threadIdx.x; k < span * span; k += BF_TILE * BF_TILE) {
        const int yy = reflect101(y0 + k / span, H), xx = reflect101(x0 + k % span, W);
        const float d = __ldg(depth + (size_t)yy * W + xx);
        s_img[k] = (d < 0.f || d != d) ? 0.f : d;
    }
    __syncthreads();
    const int x = blockIdx.x * BF_TILE + threadIdx.x, y = blockIdx.y * BF_TILE + threadIdx.y;
    if (x >= W || y >= H) return;
    const float raw = __ldg(depth + (size_t)y * W + x);
    const float vmin = __uint_as_float(mm[0]), vmax = __uint_as_float(mm[1]);
    const float v0 = s_img[(threadIdx.y + radius) * span + threadIdx.x + radius];
    float res = v0;
    if (fabsf(vmin - vmax) >= FLT_EPSILON) {
        const float scale_index = 4096.f / (vmax - vmin);
        const float inv_scale = 1.f / scale_index;
        float sum = 0.f, wsum = 0.f;
        for (int j = -radius; j <= radius; ++j)
            for (int i = -radius; i <= radius; ++i) {
                const float r2 = (float)(i * i + j * j);
                if (r2 > (float)(radius * radius)) continue;
                const float v = s_img[(threadIdx.y + radius + j) * span + threadIdx.x + radius + i];
                float alpha = fabsf(v - v0) * scale_index;
                const float fl = floorf(alpha);
                alpha -= fl;
                const float a0 = fl * inv_scale, a1 = (fl + 1.f) * inv_scale;       // LUT abscissae
                const float e0 = __expf(a0 * a0 * color_coeff), e1 = __expf(a1 * a1 * color_coeff);
                const float wgt = __expf(r2 * space_coeff) * (e0 + alpha * (e1 - e0));
                sum += v * wgt;
                wsum += wgt;
            }
        res = sum / wsum;
    }
    out[(size_t)y * W + x] = (raw < 0.f) ? -1.f : res;
}

}  // namespace

extern "C" int ags_smooth_depth(int32_t H, int32_t W, const float* depth, float* out, int32_t d,
                                float sigma_color, float sigma_space, void* scratch, void* stream) {
    AGS_CHECK_ARG(H > 0 && W > 0 && depth && out && scratch, "bad arguments");
    if (sigma_color <= 0.f) sigma_color = 1.f;
    if (sigma_space <= 0.f) sigma_space = 1.f;
    int radius = d <= 0 ? (int)lrintf(sigma_space * 1.5f) : d / 2;
    if (radius < 1) radius = 1;
    AGS_CHECK_ARG(radius <= BF_MAXR, "bilateral radius %d > %d", radius, BF_MAXR);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* mm = (unsigned*)scratch;
    const unsigned init[2] = {0x7f7fffffu /* FLT_MAX */, 0u};
    AGS_CHECK_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    ags_note_launch(); depth_minmax_kernel<<<148, 256, 0, st>>>(depth, H * W, mm);
    AGS_CHECK_CUDA(cudaGetLastError());
    const int span = BF_TILE + 2 * radius;
    dim3 grid((W + BF_TILE - 1) / BF_TILE, (H + BF_TILE - 1) / BF_TILE), block(BF_TILE, BF_TILE);
    ags_note_launch(); bilateral_kernel<<<grid, block, (size_t)span * span * sizeof(float), st>>>(
        depth, out, H, W, radius, -0.5f / (sigma_color * sigma_color), -0.5f / (sigma_space * sigma_space), mm);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
