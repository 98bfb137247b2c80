// binning.cu -- K2/K3: per-tile instance lists without a global sort and without a host sync.
//
// Replaces the scan / duplicateWithKeys / global 64-bit radix sort / identifyTileRanges stages of the
// 3DGS-lineage extension (behind /root/reference/utils/operations.py:701-713).  B200-first design:
//   1. K1 already counted instances per (view, tile) with atomics        -> tile_count
//   2. alloc_kernel: block-scan of the counts + ONE atomic per block on a device allocator
//      -> tile_offset (segments are contiguous per tile; their order in memory is irrelevant)
//   3. scatter_kernel: every visible (view, Gaussian) writes key = depth_bits<<32 | id into its
//      tiles' segments (slot claimed with an atomic)                       -> inst_key
//   4. depth sort per tile (bitonic on 64-bit keys in shared memory; unique keys => deterministic, ties
//      in depth resolved by Gaussian id exactly like the stable radix sort of the lineage).  Tiles with
//      <= AGS_FUSED_SORT_MAX instances are sorted in the PROLOGUE of composite_fwd (composite.cu), where
//      the barrier latency hides behind other CTAs' compositing; tile_sort_kernel only handles larger
//      tiles: chunk sort + rank-by-binary-search merges through a ping-pong buffer.   -> inst_sorted
// Everything is sized by device-side counters; when the batch needs more than inst_cap instances the
// overflow flag is raised and all later kernels render empty tiles.
#include "ags_common.cuh"

namespace {

constexpr int SORT_CHUNK = 4096;
constexpr int SORT_THREADS = 256;

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
    if (c > AGS_FUSED_SORT_MAX) atomicAdd(w.counters + 2, 1);      // tile_sort_kernel has work to do
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
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nvis; e += stride) {
        const size_t idx = (size_t)w.vis_list[e];
        const int v = (int)(idx / a.N);
        const uint32_t i = (uint32_t)(idx - (size_t)v * a.N);
        const uint2 r = w.rect[idx];
        const int minx = r.x & 0xffff, maxx = r.x >> 16, miny = r.y & 0xffff, maxy = r.y >> 16;
        const size_t tbase = (size_t)v * tiles_x * tiles_y;
        const float depth = w.feat0[idx].w;
        const uint64_t key = ((uint64_t)__float_as_uint(depth) << 32) | i;
        for (int ty = miny; ty < maxy; ++ty)
            for (int tx = minx; tx < maxx; ++tx) {
                const size_t t = tbase + (size_t)ty * tiles_x + tx;
                const int slot = w.tile_offset[t] + atomicAdd(w.tile_fill + t, 1);
                w.inst_key[slot] = key;
            }
    }
}

__device__ __forceinline__ void bitonic_smem(uint64_t* s, int m) {
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (m >> 1); t += SORT_THREADS) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const bool asc = ((lo & k) == 0);
                const uint64_t x = s[lo], y = s[hi];
                if ((x > y) == asc) { s[lo] = y; s[hi] = x; }
            }
            __syncthreads();
        }
    }
}

// one oversize tile (more than AGS_FUSED_SORT_MAX instances): chunk sort + rank merges
__device__ void sort_oversize_tile(const AgsWorkspace& w, int t, uint64_t* s) {
    const int n = w.tile_count[t];
    if (n <= AGS_FUSED_SORT_MAX) return;     // sorted in the prologue of composite_fwd
    const int off = w.tile_offset[t];
    uint64_t* keys = w.inst_key + off;
    int32_t* out = w.inst_sorted + off;
    // phase 1: sort chunks of SORT_CHUNK in shared memory
    for (int cbase = 0; cbase < n; cbase += SORT_CHUNK) {
        const int cn = min(SORT_CHUNK, n - cbase);
        int m = 2;
        while (m < cn) m <<= 1;
        for (int k = threadIdx.x; k < m; k += SORT_THREADS) s[k] = (k < cn) ? keys[cbase + k] : ~0ull;
        __syncthreads();
        bitonic_smem(s, m);
        if (n <= SORT_CHUNK) {
            for (int k = threadIdx.x; k < cn; k += SORT_THREADS) out[k] = (int32_t)(s[k] & 0xffffffffu);
            return;
        }
        for (int k = threadIdx.x; k < cn; k += SORT_THREADS) keys[cbase + k] = s[k];
        __syncthreads();
    }
    // phase 2: pairwise merges through the ping-pong buffer (keys are unique)
    uint64_t* src = keys;
    uint64_t* dst = w.inst_key_alt + off;
    for (int run = SORT_CHUNK; run < n; run <<= 1) {
        for (int e = threadIdx.x; e < n; e += SORT_THREADS) {
            const int r = e / run;
            const int pr = r ^ 1;
            const int ps = pr * run;
            const uint64_t key = src[e];
            int dest = e;
            if (ps < n) {
                const int pe = min(ps + run, n);
                int lo = ps, hi = pe;                       // first partner element > key
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (src[mid] < key) lo = mid + 1; else hi = mid;
                }
                dest = min(r, pr) * run + (e - r * run) + (lo - ps);
            }
            dst[dest] = key;
        }
        __syncthreads();
        uint64_t* tmp = src; src = dst; dst = tmp;
    }
    for (int k = threadIdx.x; k < n; k += SORT_THREADS) out[k] = (int32_t)(src[k] & 0xffffffffu);
}

// Oversize tiles are rare (none at all in the BASELINE workloads): the binning pass counts them in
// counters[2], and a small persistent grid returns at once when there is nothing to do -- one block per
// tile cost ~10 us per iteration in block scheduling alone.
__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(AgsWorkspace w, int n_tiles_total, int inst_cap) {
    __shared__ uint64_t s[SORT_CHUNK];
    if (w.counters[2] == 0 || w.counters[0] > inst_cap) return;
    for (int t = blockIdx.x; t < n_tiles_total; t += gridDim.x) {
        sort_oversize_tile(w, t, s);
        __syncthreads();
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
    // only tiles with more than AGS_FUSED_SORT_MAX instances are sorted here (chunk sort + merge);
    // all others are sorted in the prologue of composite_fwd
    ags_note_launch(); tile_sort_kernel<<<nt < 148 * 2 ? nt : 148 * 2, SORT_THREADS, 0, st>>>(w, nt, a.inst_cap);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
