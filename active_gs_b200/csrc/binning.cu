// binning.cu -- K2/K3: per-tile instance lists without a global sort and without a host sync.
//
// Replaces the scan / duplicateWithKeys / global 64-bit radix sort / identifyTileRanges stages of the
// 3DGS-lineage extension (behind /root/reference/utils/operations.py:701-713).  B200-first design:
//   1. K1 already counted instances per (view, tile) with atomics        -> tile_count
//   2. alloc_kernel: block-scan of the counts + ONE atomic per block on a device allocator
//      -> tile_offset (segments are contiguous per tile; their order in memory is irrelevant)
//   3. scatter_kernel: every visible (view, Gaussian) writes key = depth_bits<<32 | id into its
//      tiles' segments (slot claimed with an atomic)                       -> inst_key
//   4. depth sort per tile in the PROLOGUE of composite_fwd (composite.cu), where the barrier latency
//      hides behind other CTAs' compositing: rank sort / bitonic network on 64-bit keys in shared memory
//      (unique keys => deterministic, ties in depth resolved by Gaussian id exactly like the stable radix
//      sort of the lineage); tiles above AGS_FUSED_SORT_MAX instances: chunk sort + rank-by-binary-search
//      merges through a ping-pong buffer, by the same CTA.                            -> inst_sorted
// Everything is sized by device-side counters; when the batch needs more than inst_cap instances the
// overflow flag is raised and all later kernels render empty tiles.
#include "ags_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
alloc_kernel(AgsWorkspace w, int n_tiles_total, int tiles_per_view, int inst_cap, int32_t* stats) {
    __shared__ int warp_sums[8];
    __shared__ int block_base;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = (t < n_tiles_total) ? w.tile_count[t] : 0;
    {   // per-view instance totals
        const int v = t / tiles_per_view;
        constexpr int VMAX = AGS_NUM_STATS - AGS_STAT_VIEW0;
        if (tiles_per_view >= 32) {               // a warp's 32 consecutive tiles span at most two views
            const int v0 = __shfl_sync(0xffffffffu, v, 0);
            const int s0 = __reduce_add_sync(0xffffffffu, v == v0 ? c : 0);
            const int s1 = __reduce_add_sync(0xffffffffu, v == v0 ? 0 : c);
            if (lane == 0) {
                if (s0 && v0 < VMAX) atomicAdd(stats + AGS_STAT_VIEW0 + v0, s0);
                if (s1 && v0 + 1 < VMAX) atomicAdd(stats + AGS_STAT_VIEW0 + v0 + 1, s1);
            }
        } else if (c && v < VMAX) {
            atomicAdd(stats + AGS_STAT_VIEW0 + v, c);
        }
    }
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += n;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = (lane < 8) ? warp_sums[lane] : 0;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += n;
        }
        if (lane < 8) warp_sums[lane] = s;              // inclusive over warps
        if (lane == 7) {
            block_base = atomicAdd(w.counters, s);
            atomicAdd(stats + AGS_STAT_INSTANCES, s);
        }
    }
    __syncthreads();
    const int warp_excl = (wid == 0) ? 0 : warp_sums[wid - 1];
    if (t < n_tiles_total) w.tile_offset[t] = block_base + warp_excl + incl - c;
}

// Warp-cooperative scatter: a warp takes 32 visible (view, Gaussian) pairs, scans their tile counts and
// then emits the union of their instances 32 at a time (lane j of a round finds its source pair with a
// 5-step shuffle search over the scan).  Every round issues 32 independent slot-claiming atomics, so the
// number of dependent round trips per warp is instances/32 (about 3) instead of the largest tile count
// among its 32 pairs (one thread per pair serialised ~16 atomics behind the warp's biggest splat).
__global__ void __launch_bounds__(256)
scatter_kernel(AgsRenderArgs a, AgsWorkspace w) {
    const int total = w.counters[0];
    if (total > a.inst_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.stats[AGS_STAT_OVERFLOW] = 1;
        return;
    }
    const int nvis = w.counters[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) a.stats[AGS_STAT_VISIBLE] = nvis;
    const int tiles_x = (a.W + TILE - 1) / TILE, tiles_y = (a.H + TILE - 1) / TILE;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int e0 = warp * 32; e0 < nvis; e0 += nwarps * 32) {
        const int e = e0 + lane;
        int minx = 0, miny = 0, wx = 0, cnt = 0, tbase = 0;
        unsigned key_hi = 0u, key_lo = 0u;
        if (e < nvis) {
            const size_t idx = (size_t)w.vis_list[e];
            const int v = (int)(idx / a.N);
            key_lo = (uint32_t)(idx - (size_t)v * a.N);
            const uint2 r = w.rect[idx];
            minx = r.x & 0xffff; miny = r.y & 0xffff;
            wx = (int)(r.x >> 16) - minx;
            cnt = wx * ((int)(r.y >> 16) - miny);
            tbase = v * tiles_x * tiles_y;
            key_hi = __float_as_uint(w.feat0[idx].w);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const int total_w = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - cnt;
        for (int j0 = 0; j0 < total_w; j0 += 32) {
            const int j = j0 + lane;
            // source lane = number of lanes whose inclusive count is <= j
            int src = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int t = __shfl_sync(0xffffffffu, incl, min(src + step - 1, 31));
                if (t <= j) src += step;
            }
            src = min(src, 31);
            const int local = j - __shfl_sync(0xffffffffu, excl, src);
            const int sminx = __shfl_sync(0xffffffffu, minx, src), sminy = __shfl_sync(0xffffffffu, miny, src);
            const int swx = __shfl_sync(0xffffffffu, wx, src), stb = __shfl_sync(0xffffffffu, tbase, src);
            const unsigned khi = __shfl_sync(0xffffffffu, key_hi, src), klo = __shfl_sync(0xffffffffu, key_lo, src);
            if (j < total_w) {
                const int ty = sminy + local / swx, tx = sminx + local - (local / swx) * swx;
                const size_t t = (size_t)stb + (size_t)ty * tiles_x + tx;
                const int slot = w.tile_offset[t] + atomicAdd(w.tile_fill + t, 1);
                w.inst_key[slot] = ((uint64_t)khi << 32) | klo;
            }
        }
    }
}

}  // namespace

int ags_launch_binning(const AgsRenderArgs& a, const AgsWorkspace& w) {
    cudaStream_t st = (cudaStream_t)a.stream;
    const int tiles = ((a.W + TILE - 1) / TILE) * ((a.H + TILE - 1) / TILE);
    const int nt = a.B * tiles;
    ags_note_launch(); alloc_kernel<<<(nt + 255) / 256, 256, 0, st>>>(w, nt, tiles, a.inst_cap, a.stats);
    AGS_CHECK_CUDA(cudaGetLastError());
    if (a.N > 0) {
        long long blocks = ((long long)a.N * a.B + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        ags_note_launch(); scatter_kernel<<<(int)blocks, 256, 0, st>>>(a, w);
        AGS_CHECK_CUDA(cudaGetLastError());
    }
    // the depth sort of every tile happens in the prologue of composite_fwd (composite.cu)
    return 0;
}
