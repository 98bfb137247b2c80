// ags_common.cuh -- shared definitions for the sm_100a kernels behind include/ags_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ags_b200.h"

#define TILE AGS_TILE
#define TILE_PIX (TILE * TILE)

// rasterizer constants (oracle/rasterizer_ref.py)
#define AGS_NEAR_CULL 0.2f
#define AGS_LOWPASS 0.3f
#define AGS_ALPHA_MAX 0.99f
#define AGS_ALPHA_MIN (1.0f / 255.0f)
#define AGS_T_EPS 1e-4f
#define AGS_SLOPE_COS_MIN 0.1f
// tiles with at most this many instances are depth-sorted inside composite_fwd's prologue (keys alias
// the 20 KB staging buffer); larger tiles are chunk-sorted and merged by the same CTA (sort_oversize_tile)
#define AGS_FUSED_SORT_MAX 2048

// slots of the 16-float per-(view, Gaussian) gradient record `dsplat` (raw moments written by composite_bwd,
// chain rule applied by project_bwd; the order is what the folded butterfly of composite_bwd ends with)
#define AGS_REC_C0 0      // [0..2] sum w*gC
#define AGS_REC_N0 3      // [3..5] sum w*gN
#define AGS_REC_WD 6      // sum w*gD
#define AGS_REC_PAD 7
#define AGS_REC_PDX 8     // sum dpower*dx
#define AGS_REC_PXX 9     // sum dpower*dx^2
#define AGS_REC_WDX 10    // sum w*gD*dx
#define AGS_REC_PXY 11    // sum dpower*dx*dy
#define AGS_REC_PDY 12    // sum dpower*dy
#define AGS_REC_PYY 13    // sum dpower*dy^2
#define AGS_REC_WDY 14    // sum w*gD*dy
#define AGS_REC_P1 15     // sum dpower          (d opacity = sum / opacity)

// ------------------------------------------------------------------------------------------------
// Workspace layout (all offsets 256-byte aligned). One workspace serves one batch of B views and
// carries everything the backward needs.
struct AgsWorkspace {
    float4* geom0;      // (B*N) x, y, conic_a, conic_b
    float4* geom1;      // (B*N) conic_c, opacity, slope_x, slope_y
    float4* feat0;      // (B*N) r, g, b, depth
    float4* feat1;      // (B*N) nx, ny, nz, confidence
    uint2* rect;        // (B*N) packed tile rect: x = minx | maxx<<16, y = miny | maxy<<16
    int32_t* vis_list;  // (B*N) compact list of visible pair indices v*N+i (count in counters[1])
    float* dsplat;      // (B*N*16) per-view per-Gaussian gradient record (backward)
    int32_t* tile_count;   // (B*tiles)
    int32_t* tile_offset;  // (B*tiles)
    int32_t* tile_fill;    // (B*tiles)
    int32_t* counters;     // (8) [0] = instance allocator, [1] = visible pairs, [2] = unused
    uint64_t* inst_key;    // (inst_cap) depth_bits<<32 | gaussian id, grouped per tile
    uint64_t* inst_key_alt;// (inst_cap) ping-pong buffer for oversize tiles
    int32_t* inst_sorted;  // (inst_cap) gaussian ids, front-to-back per tile
    float* final_T;        // (B*P)
    int32_t* n_contrib;    // (B*P) index+1 of the last instance that contributed
    float4* inst_rec;      // (inst_cap * 5) depth-sorted 80-byte staging records per tile, or NULL: written by
                           // composite_fwd, bulk-copied (TMA) by composite_bwd -- only with AGS_BWD_TMA=1
    size_t total;
};

// AGS_BWD_TMA=1 in the environment (read once): composite_fwd also writes the sorted staging records and
// composite_bwd stages its batches with cp.async.bulk + mbarrier instead of gathering them
bool ags_use_tma();

__host__ __device__ inline size_t ags_align256(size_t x) { return (x + 255) & ~(size_t)255; }

inline AgsWorkspace ags_carve(void* base, int N, int B, int H, int W, int inst_cap) {
    AgsWorkspace w;
    size_t off = 0;
    char* p = (char*)base;
    const size_t BN = (size_t)B * N;
    const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    const size_t BT = (size_t)B * tiles, BP = (size_t)B * H * W;
    auto take = [&](size_t bytes) { char* q = p + off; off += ags_align256(bytes); return (void*)q; };
    w.geom0 = (float4*)take(BN * 16);
    w.geom1 = (float4*)take(BN * 16);
    w.feat0 = (float4*)take(BN * 16);
    w.feat1 = (float4*)take(BN * 16);
    w.rect = (uint2*)take(BN * 8);
    w.vis_list = (int32_t*)take(BN * 4);
    w.dsplat = (float*)take(BN * 64);
    w.tile_count = (int32_t*)take(BT * 4);
    w.tile_offset = (int32_t*)take(BT * 4);
    w.tile_fill = (int32_t*)take(BT * 4);
    w.counters = (int32_t*)take(8 * 4);
    w.inst_key = (uint64_t*)take((size_t)inst_cap * 8);
    w.inst_key_alt = (uint64_t*)take((size_t)inst_cap * 8);
    w.inst_sorted = (int32_t*)take((size_t)inst_cap * 4);
    w.final_T = (float*)take(BP * 4);
    w.n_contrib = (int32_t*)take(BP * 4);
    w.inst_rec = ags_use_tma() ? (float4*)take((size_t)inst_cap * 80) : nullptr;
    w.total = off;
    return w;
}

// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local message, no exceptions across the C boundary)
void ags_set_error(const char* fmt, ...);
#define AGS_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ags_set_error(__VA_ARGS__);     \
            return -1;                      \
        }                                   \
    } while (0)
#define AGS_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ags_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));          \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

// every kernel launch of the library is counted (bench.py reports the number it saw inside the timed region)
void ags_note_launch();

// kernel launchers implemented in the individual .cu files
int ags_launch_project_fwd(const AgsRenderArgs& a, const AgsWorkspace& w, bool for_backward);
int ags_launch_binning(const AgsRenderArgs& a, const AgsWorkspace& w);
int ags_launch_composite_fwd(const AgsRenderArgs& a, const AgsWorkspace& w);
int ags_launch_composite_bwd(const AgsRenderArgs& a, const AgsRenderGradArgs& g, const AgsWorkspace& w);
int ags_launch_project_bwd(const AgsRenderArgs& a, const AgsRenderGradArgs& g, const AgsWorkspace& w);

// ------------------------------------------------------------------------------------------------
// small device helpers
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

struct Cam {
    float V[16];  // viewmatrix, row-vector convention: t_j = sum_k p_k V[4k+j] + V[12+j]
    float M[16];  // projmatrix, same convention
    float tanx, tany;
};

__device__ __forceinline__ void load_cam(Cam& c, const float* vm, const float* pm, const float* tf, int v) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        c.V[k] = __ldg(vm + v * 16 + k);
        c.M[k] = __ldg(pm + v * 16 + k);
    }
    c.tanx = __ldg(tf + v * 2);
    c.tany = __ldg(tf + v * 2 + 1);
}

// ------------------------------------------------------------------------------------------------
// cross-GPU signal / wait folded into the exchange kernels (AgsDistSync, include/ags_b200.h)
struct SyncP {
    int32_t* peers[AGS_MAX_PEERS];
    int epoch, world, rank;
    bool on;
};

inline SyncP make_sync(const AgsDistSync& s, int world, int rank) {
    SyncP p;
    for (int r = 0; r < AGS_MAX_PEERS; ++r) p.peers[r] = r < world ? s.peers[r] : nullptr;
    p.epoch = s.epoch; p.world = world; p.rank = rank;
    p.on = s.peers[0] != nullptr;
    return p;
}

// consumer side: every block, before it touches the peers' data
__device__ __forceinline__ void sync_wait(const SyncP& s, int phase) {
    if (!s.on) return;
    if (threadIdx.x < s.world && threadIdx.y == 0) {
        const int32_t* f = s.peers[s.rank] + phase * AGS_MAX_PEERS + threadIdx.x;
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        } while (v < s.epoch);
    }
    __syncthreads();
}

// producer side, called by ALL threads of ONE block after the data the phase publishes is written
__device__ __forceinline__ void sync_signal(const SyncP& s, int phase) {
    if (!s.on) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < s.world && threadIdx.y == 0) {
        int32_t* f = s.peers[threadIdx.x] + phase * AGS_MAX_PEERS + s.rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(f), "r"(s.epoch) : "memory");
    }
}

// "all blocks of this grid are done": true in exactly one block (the last to arrive); the counter lives in
// the rank's own flags array (word AGS_SYNC_WORDS/2 + phase) and is reset for the next launch
__device__ __forceinline__ bool sync_last_block(const SyncP& s, int phase) {
    __shared__ int s_last;
    // gpu scope is enough here: the data lives in this GPU's memory and the one block that signals the
    // peers issues the system-scope fence (sync_signal) after it has observed every other block's arrival
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int32_t* c = s.peers[s.rank] + AGS_SYNC_WORDS / 2 + phase;
        const int nblk = gridDim.x * gridDim.y * gridDim.z;
        const int t = atomicAdd(c, 1);
        s_last = (t == nblk - 1);
        if (s_last) *c = 0;
    }
    __syncthreads();
    return s_last != 0;
}
