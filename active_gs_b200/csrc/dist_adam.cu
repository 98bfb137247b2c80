// dist_adam.cu -- K9: ONE kernel per GPU that fuses the data-parallel exchange with the optimiser:
//     reduce-scatter of the per-Gaussian gradients  ->  Adam on the owned shard  ->  all-gather of
//     the updated parameters,
// over NVLink 5 / NVSwitch peer memory.  Replaces (NCCL all-reduce of 56 B/Gaussian + an Adam step
// replicated on every GPU) in the frame-sharded training loop (the reference
// itself is single-GPU, mapping/gaussian_map.py:126).
//
// Every rank owns the contiguous shard [rank*C, (rank+1)*C) of the flat 14*N parameter vector.
// Gradient, parameter and overflow-flag buffers are symmetric allocations (same offset on every
// GPU); the caller passes the peer pointers.  Two data paths:
//   * multimem (NVLS): multimem.ld_reduce.add sums the shard over all GPUs inside the NVSwitch and
//     multimem.st broadcasts the updated parameters -- one load and one store per 16 bytes;
//   * peer pointers: 128-bit loads from each peer's gradient buffer, 128-bit stores into each peer's
//     parameter buffer.
// Exponential averages live only on the owner (full-size buffers, only the shard is touched).
// The caller brackets the launch with cross-GPU barriers on the same stream (gradients complete
// everywhere before; parameters visible everywhere after).  If ANY rank raised its overflow flag
// the step is skipped on all ranks.
#include "ags_common.cuh"

namespace {

struct DistAdamParams {
    int world, rank;
    const float* grad[AGS_MAX_PEERS];
    float* param[AGS_MAX_PEERS];
    const int* skip[AGS_MAX_PEERS];
    const float* grad_mc;
    float* param_mc;
    float* m;
    float* v;
    long long total, shard_begin, shard_end;
    long long seg_end[AGS_ADAM_GROUPS];
    float lr[AGS_ADAM_GROUPS];
    int groups;
    float b1, b2, eps;
    int step;
    SyncP sync;
};

__device__ __forceinline__ float4 ld_peer4(const float* p) {          // bypass L1: written by another GPU
    float4 r;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_peer4(float* p, float4 v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce4(const float* p) {     // in-switch sum over all GPUs
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void mc_st4(float* p, float4 v) {          // broadcast store to all GPUs
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float lr_of(const DistAdamParams& P, long long e) {
    float lr = P.lr[0];
#pragma unroll
    for (int k = 0; k < AGS_ADAM_GROUPS - 1; ++k)
        if (k < P.groups - 1 && e >= P.seg_end[k]) lr = P.lr[k + 1];
    return lr;
}

#ifndef AGS_DIST_UNROLL
#define AGS_DIST_UNROLL 4            // independent 16-byte elements in flight per thread
#endif

// Measured (profiles/README.md, N=2, 5.6 MB shard): ONE wave of 296 blocks with four elements in flight
// per thread: 43 us (peer loads) / 57 us (multimem).  More, smaller blocks are slower (every block pays
// the remote flag read + bias-correction prologue: 1184 blocks = 62-74 us); wider unrolling over the
// peers costs occupancy (162 registers: 135 us).
__global__ void __launch_bounds__(256)
dist_adam_kernel(DistAdamParams P) {
    // folded barrier: this kernel runs after the rank's own backward (stream order), so block 0 can tell
    // every rank "my gradients are complete" right away; then every block waits for all ranks
    if (P.sync.on) {
        if (blockIdx.x == 0) sync_signal(P.sync, AGS_SYNC_GRADS);
        sync_wait(P.sync, AGS_SYNC_GRADS);
    }
    // any rank's overflow flag skips the step everywhere; the peer loads are issued together
    int skip = 0;
    for (int p = 0; p < P.world; ++p)
        if (P.skip[p]) skip |= *reinterpret_cast<const volatile int*>(P.skip[p]);
    if (skip) {
        if (P.sync.on && sync_last_block(P.sync, AGS_SYNC_PARAMS)) sync_signal(P.sync, AGS_SYNC_PARAMS);
        return;
    }
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {                     // bias corrections once per block (double pow)
        const double bc1 = 1.0 - pow((double)P.b1, (double)P.step);
        const double bc2 = 1.0 - pow((double)P.b2, (double)P.step);
        s_c[0] = (float)(1.0 / sqrt(bc2));
        s_c[1] = (float)(1.0 / bc1);
    }
    __syncthreads();
    const float inv_sqrt_bc2 = s_c[0];
    const float inv_bc1 = s_c[1];
    const long long n4 = (P.shard_end - P.shard_begin + 3) / 4;        // shard is 4-aligned; tail padded
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = AGS_DIST_UNROLL;
    for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < n4; q0 += stride * U) {
        float4 g[U], pr[U], m4[U], v4[U];
        // ---- all loads of the U elements first: the NVLink / switch latency is paid once
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = q0 + u * stride;
            if (q < n4) {
                const long long e = P.shard_begin + 4 * q;
                if (P.grad_mc) {
                    g[u] = mc_ld_reduce4(P.grad_mc + e);
                } else {
                    g[u] = ld_peer4(P.grad[0] + e);
                    for (int p = 1; p < P.world; ++p) {
                        const float4 h = ld_peer4(P.grad[p] + e);
                        g[u].x += h.x; g[u].y += h.y; g[u].z += h.z; g[u].w += h.w;
                    }
                }
                pr[u] = *reinterpret_cast<const float4*>(P.param[P.rank] + e);
                m4[u] = *reinterpret_cast<const float4*>(P.m + e);
                v4[u] = *reinterpret_cast<const float4*>(P.v + e);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = q0 + u * stride;
            if (q < n4) {
                const long long e = P.shard_begin + 4 * q;
                float* gp = &g[u].x; float* pp = &pr[u].x; float* mp = &m4[u].x; float* vp = &v4[u].x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float lr = lr_of(P, e + k);
                    const float m = P.b1 * mp[k] + (1.f - P.b1) * gp[k];
                    const float v = P.b2 * vp[k] + (1.f - P.b2) * gp[k] * gp[k];
                    mp[k] = m; vp[k] = v;
                    pp[k] -= (lr * inv_bc1) * (m / (sqrtf(v) * inv_sqrt_bc2 + P.eps));
                }
                *reinterpret_cast<float4*>(P.m + e) = m4[u];
                *reinterpret_cast<float4*>(P.v + e) = v4[u];
                if (P.param_mc) {
                    mc_st4(P.param_mc + e, pr[u]);
                } else {
                    for (int p = 0; p < P.world; ++p) st_peer4(P.param[p] + e, pr[u]);
                }
            }
        }
    }
    __threadfence_system();
    // every rank's parameter buffer holds this rank's updated shard once the last block is done
    if (P.sync.on && sync_last_block(P.sync, AGS_SYNC_PARAMS)) sync_signal(P.sync, AGS_SYNC_PARAMS);
}

}  // namespace

extern "C" int ags_dist_adam_step(const AgsDistAdamArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->world >= 1 && a->world <= AGS_MAX_PEERS && a->rank >= 0 && a->rank < a->world,
                  "bad world/rank %d/%d", a->world, a->rank);
    AGS_CHECK_ARG(a->num_groups > 0 && a->num_groups <= AGS_ADAM_GROUPS, "bad num_groups %d", a->num_groups);
    AGS_CHECK_ARG(a->numel_padded > 0 && a->numel_padded % (4 * a->world) == 0,
                  "numel_padded must be a positive multiple of 4*world");
    AGS_CHECK_ARG(a->exp_avg && a->exp_avg_sq, "NULL optimiser state");
    AGS_CHECK_ARG(a->step >= 1, "step must be >= 1");
    DistAdamParams P;
    long long tot = 0;
    for (int k = 0; k < AGS_ADAM_GROUPS; ++k) {
        if (k < a->num_groups) { AGS_CHECK_ARG(a->numel[k] >= 0, "negative numel"); tot += a->numel[k]; }
        P.seg_end[k] = tot;
        P.lr[k] = k < a->num_groups ? a->lr[k] : 0.f;
    }
    AGS_CHECK_ARG(tot <= a->numel_padded, "groups exceed the padded buffer");
    for (int p = 0; p < AGS_MAX_PEERS; ++p) {
        const bool in = p < a->world;
        if (in) AGS_CHECK_ARG(a->grad_peers[p] && a->param_peers[p], "NULL peer pointer %d", p);
        P.grad[p] = in ? a->grad_peers[p] : nullptr;
        P.param[p] = in ? a->param_peers[p] : nullptr;
        P.skip[p] = in ? a->skip_peers[p] : nullptr;
    }
    P.world = a->world; P.rank = a->rank;
    P.grad_mc = a->grad_multicast; P.param_mc = a->param_multicast;
    P.m = a->exp_avg; P.v = a->exp_avg_sq;
    P.total = a->numel_padded;
    const long long C = a->numel_padded / a->world;
    P.shard_begin = C * a->rank; P.shard_end = C * (a->rank + 1);
    P.groups = a->num_groups;
    P.b1 = a->beta1; P.b2 = a->beta2; P.eps = a->eps; P.step = a->step;
    P.sync = make_sync(a->sync, a->world, a->rank);
    const long long n4 = C / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;          // one wave (see the kernel's note)
    if (blocks < 1) blocks = 1;
    ags_note_launch(); dist_adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)a->stream>>>(P);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
