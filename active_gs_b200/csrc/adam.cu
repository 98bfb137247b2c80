// adam.cu -- K7: one launch of Adam over the five SoA parameter groups of a GaussianMap.
//
// Replaces torch.optim.Adam(eps=1e-15) of /root/reference/mapping/gaussian_map.py:259-292,126-127
// (5 foreach groups -> 1 kernel).  Update rule (torch single-tensor form):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// HBM roofline: 28 B per parameter (read p,g,m,v; write p,m,v) = 392 B per Gaussian.
#include "ags_common.cuh"

namespace {

struct AdamParams {
    float* p[AGS_ADAM_GROUPS];
    const float* g[AGS_ADAM_GROUPS];
    float* m[AGS_ADAM_GROUPS];
    float* v[AGS_ADAM_GROUPS];
    long long end[AGS_ADAM_GROUPS];   // exclusive prefix of numel
    float lr[AGS_ADAM_GROUPS];
    int groups;
    float b1, b2, eps;
    int step;
    const int* step_dev;
    const int* skip_flag;
};

__global__ void __launch_bounds__(256)
adam_kernel(AdamParams P, long long total) {
    if (P.skip_flag && *P.skip_flag != 0) return;
    const int t = P.step_dev ? (*P.step_dev + 1) : P.step;
    const double bc1 = 1.0 - pow((double)P.b1, (double)t);
    const double bc2 = 1.0 - pow((double)P.b2, (double)t);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int grp = 0;
        long long start = 0;
#pragma unroll
        for (int k = 0; k < AGS_ADAM_GROUPS - 1; ++k)
            if (k < P.groups - 1 && e >= P.end[k]) { grp = k + 1; start = P.end[k]; }
        const long long i = e - start;
        const float g = P.g[grp][i];
        const float m = P.b1 * P.m[grp][i] + (1.f - P.b1) * g;
        const float v = P.b2 * P.v[grp][i] + (1.f - P.b2) * g * g;
        const float step_size = (float)((double)P.lr[grp] / bc1);
        const float denom = sqrtf(v) * inv_sqrt_bc2 + P.eps;
        P.m[grp][i] = m;
        P.v[grp][i] = v;
        P.p[grp][i] -= step_size * (m / denom);
    }
}

__global__ void tick_kernel(int* step_dev, const int* skip_flag) {
    if (skip_flag && *skip_flag != 0) return;
    *step_dev += 1;
}

}  // namespace

extern "C" int ags_adam_step(const AgsAdamArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->num_groups > 0 && a->num_groups <= AGS_ADAM_GROUPS, "bad num_groups %d", a->num_groups);
    AdamParams P;
    long long total = 0;
    for (int k = 0; k < AGS_ADAM_GROUPS; ++k) {
        if (k < a->num_groups) {
            AGS_CHECK_ARG(a->numel[k] >= 0, "negative numel");
            if (a->numel[k] > 0)
                AGS_CHECK_ARG(a->param[k] && a->grad[k] && a->exp_avg[k] && a->exp_avg_sq[k], "NULL tensor in group %d", k);
            total += a->numel[k];
        }
        P.p[k] = k < a->num_groups ? a->param[k] : nullptr;
        P.g[k] = k < a->num_groups ? a->grad[k] : nullptr;
        P.m[k] = k < a->num_groups ? a->exp_avg[k] : nullptr;
        P.v[k] = k < a->num_groups ? a->exp_avg_sq[k] : nullptr;
        P.lr[k] = k < a->num_groups ? a->lr[k] : 0.f;
        P.end[k] = total;
    }
    P.groups = a->num_groups;
    P.b1 = a->beta1; P.b2 = a->beta2; P.eps = a->eps;
    P.step = a->step; P.step_dev = a->step_dev; P.skip_flag = a->skip_flag;
    AGS_CHECK_ARG(a->step_dev != nullptr || a->step >= 1, "step must be >= 1");
    if (total == 0) return 0;
    cudaStream_t st = (cudaStream_t)a->stream;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long max_blocks = 148LL * 16;          // 16 resident CTAs of 256 threads per SM
    if (blocks > max_blocks) blocks = max_blocks;
    adam_kernel<<<(int)blocks, threads, 0, st>>>(P, total);
    AGS_CHECK_CUDA(cudaGetLastError());
    if (a->step_dev) {
        tick_kernel<<<1, 1, 0, st>>>(a->step_dev, a->skip_flag);
        AGS_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}
