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
    float* g[AGS_ADAM_GROUPS];
    float* m[AGS_ADAM_GROUPS];
    float* v[AGS_ADAM_GROUPS];
    long long end[AGS_ADAM_GROUPS];   // exclusive prefix of numel
    float lr[AGS_ADAM_GROUPS];
    int groups;
    float b1, b2, eps;
    int step;
    const int* step_dev;
    const int* skip_flag;
    int zero_grad;
};

// grid = (blocks, groups): blockIdx.y selects the parameter group, every thread owns quads of four
// consecutive elements (128-bit accesses on the 16-byte aligned tensors; scalar tail).  The bias
// corrections are computed once per block (double precision pow) and broadcast through smem.
__global__ void __launch_bounds__(256)
adam_kernel(AdamParams P, long long total) {
    if (P.skip_flag && *P.skip_flag != 0) return;
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const int t = P.step_dev ? (*P.step_dev + 1) : P.step;
        const double bc1 = 1.0 - pow((double)P.b1, (double)t);
        const double bc2 = 1.0 - pow((double)P.b2, (double)t);
        s_c[0] = (float)(1.0 / sqrt(bc2));
        s_c[1] = (float)(1.0 / bc1);
    }
    __syncthreads();
    const float inv_sqrt_bc2 = s_c[0];
    const int grp = blockIdx.y;
    const long long n = P.end[grp] - (grp ? P.end[grp - 1] : 0);
    const float step_size = P.lr[grp] * s_c[1];
    float* __restrict__ p = P.p[grp];
    float* __restrict__ g = P.g[grp];
    const bool zg = P.zero_grad != 0;
    float* __restrict__ m = P.m[grp];
    float* __restrict__ v = P.v[grp];
    const float b1 = P.b1, b2 = P.b2, eps = P.eps;
    const long long nq = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool aligned = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
    if (aligned) {
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
            const float4 g4 = reinterpret_cast<const float4*>(g)[q];
            float4 m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q];
            float4 p4 = reinterpret_cast<float4*>(p)[q];
            const float* gg = &g4.x; float* mm = &m4.x; float* vv = &v4.x; float* pp = &p4.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mm[k] = b1 * mm[k] + (1.f - b1) * gg[k];
                vv[k] = b2 * vv[k] + (1.f - b2) * gg[k] * gg[k];
                pp[k] -= step_size * (mm[k] / (sqrtf(vv[k]) * inv_sqrt_bc2 + eps));
            }
            reinterpret_cast<float4*>(m)[q] = m4;
            reinterpret_cast<float4*>(v)[q] = v4;
            reinterpret_cast<float4*>(p)[q] = p4;
            if (zg) reinterpret_cast<float4*>(g)[q] = make_float4(0.f, 0.f, 0.f, 0.f);   // consumed: ready for the next backward
        }
    }
    const long long tail_from = aligned ? (nq << 2) : 0;
    for (long long i = tail_from + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
        if (zg) g[i] = 0.f;
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
        P.g[k] = k < a->num_groups ? const_cast<float*>(a->grad[k]) : nullptr;
        P.m[k] = k < a->num_groups ? a->exp_avg[k] : nullptr;
        P.v[k] = k < a->num_groups ? a->exp_avg_sq[k] : nullptr;
        P.lr[k] = k < a->num_groups ? a->lr[k] : 0.f;
        P.end[k] = total;
    }
    P.groups = a->num_groups;
    P.b1 = a->beta1; P.b2 = a->beta2; P.eps = a->eps;
    P.step = a->step; P.step_dev = a->step_dev; P.skip_flag = a->skip_flag;
    P.zero_grad = a->zero_grad;
    AGS_CHECK_ARG(a->step_dev != nullptr || a->step >= 1, "step must be >= 1");
    if (total == 0) return 0;
    cudaStream_t st = (cudaStream_t)a->stream;
    const int threads = 256;
    long long biggest = 0;
    for (int k = 0; k < a->num_groups; ++k) biggest = a->numel[k] > biggest ? a->numel[k] : biggest;
    long long blocks = (biggest / 4 + threads - 1) / threads;
    const long long max_blocks = 148LL * 4;           // x groups in grid.y
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, a->num_groups);
    ags_note_launch(); adam_kernel<<<grid, threads, 0, st>>>(P, total);
    AGS_CHECK_CUDA(cudaGetLastError());
    if (a->step_dev) {
        ags_note_launch(); tick_kernel<<<1, 1, 0, st>>>(a->step_dev, a->skip_flag);
        AGS_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}
