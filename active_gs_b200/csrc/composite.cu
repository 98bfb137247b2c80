// composite.cu -- K4 composite_fwd and K5 composite_bwd: per-pixel front-to-back alpha compositing
// of rgb / normal / plane depth / opacity / confidence and its backward.
//
// Replaces renderCUDA fwd/bwd of the native extension behind
// /root/reference/utils/operations.py:701-713 (semantics: DESIGN.md section 2, oracle/rasterizer_ref.py).
//
// One CTA = one 16x16 tile of one view (grid.z = view: all B views of an iteration in one launch).
// The tile's depth-sorted splat records (4 x float4 = 64 B each) are staged through shared memory in
// batches of 256 with 128-bit loads; the staging thread also derives the splat's cutoff bounding box
// (the pixels where alpha can reach 1/255).  Each WARP owns an 8x4 pixel block of the tile, tests 32
// splats at a time against its block (one ballot) and walks only the overlapping ones in depth
// order (shared-memory broadcast reads) -- the skipped (warp, splat) pairs cost no evaluation at all.  The backward walks the SAME front-to-back order (no 1/(1-alpha) recurrences:
// the "remaining" sum is total - prefix, computed from the saved forward outputs), reduces the 15
// per-Gaussian partial gradients across the warp through a shared-memory transposition (36
// instructions instead of 150 for a plain shuffle tree) and issues one 15-lane vector RED per
// (warp, Gaussian) into a 64 B gradient record.
#include <stdlib.h>
#include "ags_common.cuh"

namespace {

constexpr int BATCH = 256;
constexpr int FUSED_SORT_MAX = AGS_FUSED_SORT_MAX;   // 64-bit keys that fit the staging buffer (20 KB)

// one staged splat in shared memory: 80 B, read with 128-bit broadcast loads at immediate offsets
struct __align__(16) SplatRec {
    float4 g0, g1, f0, f1, bb;
};

#define AGS_LOG2E 1.4426950408889634f

__device__ __forceinline__ void store_rec(float4* base, size_t i, const SplatRec& r) {
    float4* p = base + i * 5;
    p[0] = r.g0; p[1] = r.g1; p[2] = r.f0; p[3] = r.f1; p[4] = r.bb;
}

// TMA bulk copy (cp.async.bulk, 1-D) of contiguous staging records into shared memory, completion on an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {      // MUFU.EX2, flush-to-zero, no range fix-up
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {      // MUFU.RCP
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The staged record carries the conic pre-multiplied by -0.5*log2(e) (a, c) and -log2(e) (b), the two diagonal terms
// next to each other: g0 = (x, y, a', c'), g1 = (b', opacity, slope_x, slope_y).  The exponent then feeds MUFU.EX2
// directly, G = exp(power) = 2^(power2), power2 = a' dx^2 + c' dy^2 + b' dx dy, and its evaluation is three packed
// fp32x2 instructions on natural register pairs: (dx,dy) = (x,y) - (px,py); (dx^2,dy^2); (a' dx^2, c' dy^2).
// `power` below is power2 (same sign as the natural exponent); `G` is the un-clamped Gaussian.
__device__ __forceinline__ void stage_geom(SplatRec& r, const float4 g0, const float4 g1) {   // from K1's (x,y,a,b) (c,o,sx,sy)
    r.g0 = make_float4(g0.x, g0.y, g0.z * (-0.5f * AGS_LOG2E), g1.x * (-0.5f * AGS_LOG2E));
    r.g1 = make_float4(g0.w * (-AGS_LOG2E), g1.y, g1.z, g1.w);
}
__device__ __forceinline__ float splat_power(const float4 g0, const float4 g1, float2 neg_pix, float& dx, float& dy) {
    const float2 d = __fadd2_rn(make_float2(g0.x, g0.y), neg_pix);
    const float2 q = __fmul2_rn(make_float2(g0.z, g0.w), __fmul2_rn(d, d));
    dx = d.x; dy = d.y;
    return fmaf(g1.x * d.x, d.y, q.x + q.y);
}
struct SplatEval {
    float dx, dy, power, G, alpha;
    bool skip;
};

__device__ __forceinline__ SplatEval eval_alpha(const float4 g0, const float4 g1, float2 neg_pix) {
    SplatEval e;
    e.power = splat_power(g0, g1, neg_pix, e.dx, e.dy);
    e.G = ex2_approx(e.power);
    e.alpha = fminf(AGS_ALPHA_MAX, g1.y * e.G);
    e.skip = (e.power > 0.f) || (e.alpha < AGS_ALPHA_MIN);
    return e;
}

// Cutoff bounding box of a splat: alpha = o*exp(-q/2) >= 1/255  <=>  q <= tau = 2 ln(255 o); the
// extent of {q <= tau} along x is sqrt(tau * Sigma_xx), Sigma = conic^-1.  Returned as
// (xmin, xmax, ymin, ymax) in pixel coordinates, padded by a rounding margin; empty if o < 1/255.
__device__ __forceinline__ float4 splat_bbox(const float4 g0, const float4 g1) {
    const float tau = 2.f * __logf(255.f * g1.y);
    if (!(tau > 0.f)) return make_float4(1e30f, -1e30f, 1e30f, -1e30f);
    const float det = g0.z * g1.x - g0.w * g0.w;
    const float inv = 1.f / det;
    const float hx = sqrtf(tau * g1.x * inv) * 1.0001f + 1e-3f;
    const float hy = sqrtf(tau * g0.z * inv) * 1.0001f + 1e-3f;
    return make_float4(g0.x - hx, g0.x + hx, g0.y - hy, g0.y + hy);
}

struct WarpBlock {       // the 8x4 pixel block a warp owns inside its 16x16 tile
    int px, py;          // this lane's pixel
    float x0, x1, y0, y1;  // block extent in pixel coordinates (inclusive)
};

__device__ __forceinline__ WarpBlock warp_block(int tile_x, int tile_y, int tid) {
    const int w = tid >> 5, lane = tid & 31;
    const int bx = tile_x * TILE + (w & 1) * 8, by = tile_y * TILE + (w >> 1) * 4;
    WarpBlock b;
    b.px = bx + (lane & 7);
    b.py = by + (lane >> 3);
    b.x0 = (float)bx; b.x1 = (float)(bx + 7); b.y0 = (float)by; b.y1 = (float)(by + 3);
    return b;
}

__device__ __forceinline__ bool bbox_hits(const float4 bb, const WarpBlock& b) {
    return (bb.x <= b.x1) && (bb.y >= b.x0) && (bb.z <= b.y1) && (bb.w >= b.y0);
}

// Depth sort of an OVERSIZE tile (more than AGS_FUSED_SORT_MAX instances; none in the BASELINE workloads)
// by the tile's own CTA: bitonic sort of chunks of SORT_CHUNK keys in the 20 KB staging buffer, then
// rank-by-binary-search merges through the ping-pong key buffer (keys are unique).  256 threads.
constexpr int SORT_CHUNK = AGS_FUSED_SORT_MAX;

__device__ __forceinline__ void bitonic_cta(uint64_t* s, int m, int tid) {
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (m >> 1); t += 256) {
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

__device__ void sort_oversize_tile(const AgsWorkspace& w, int n, int off, uint64_t* s, int tid) {
    uint64_t* keys = w.inst_key + off;
    int32_t* out = w.inst_sorted + off;
    for (int cbase = 0; cbase < n; cbase += SORT_CHUNK) {
        const int cn = min(SORT_CHUNK, n - cbase);
        int m = 2;
        while (m < cn) m <<= 1;
        for (int k = tid; k < m; k += 256) s[k] = (k < cn) ? keys[cbase + k] : ~0ull;
        __syncthreads();
        bitonic_cta(s, m, tid);
        for (int k = tid; k < cn; k += 256) keys[cbase + k] = s[k];
        __syncthreads();
    }
    uint64_t* src = keys;
    uint64_t* dst = w.inst_key_alt + off;
    for (int run = SORT_CHUNK; run < n; run <<= 1) {
        __threadfence_block();
        for (int e = tid; e < n; e += 256) {
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
    for (int k = tid; k < n; k += 256) out[k] = (int32_t)(src[k] & 0xffffffffu);
    __syncthreads();
}

// K4 ---------------------------------------------------------------------------------------------
#ifndef AGS_RANKSORT
#define AGS_RANKSORT 1        // single-batch tiles: rank sort fused with the staging (0 = bitonic prologue)
#endif
#ifndef AGS_BWD_PX_DEFAULT
#define AGS_BWD_PX_DEFAULT 1  // pixels per lane of the backward (composite_bwd_kernel<PX>)
#endif
#ifndef AGS_BWD_PX2_MINB
#define AGS_BWD_PX2_MINB 6
#endif
#ifndef AGS_BWD_RED_DEFAULT
#define AGS_BWD_RED_DEFAULT 3 // cross-lane reduction variant of the backward (see bwd_red()): measured 318 / 313 / 301 us for 0 / 1 / 2
#endif
#ifndef AGS_FWD_MINB
#define AGS_FWD_MINB 6
#endif
#ifndef AGS_BWD_MINB
#define AGS_BWD_MINB 5
#endif
#ifndef AGS_BWD_MMA_MINB
#define AGS_BWD_MMA_MINB 4    // the MMA reduction keeps 16 constant words per thread: 64 registers, no spills
#endif
template <bool WANT_IMP>
__global__ void __launch_bounds__(256, AGS_FWD_MINB)
composite_fwd_kernel(AgsRenderArgs a, AgsWorkspace w) {
    // shared memory: the staging records; the same bytes first serve the tile's depth sort
    __shared__ __align__(16) unsigned char s_raw[sizeof(SplatRec) * BATCH];
    SplatRec* s_rec = reinterpret_cast<SplatRec*>(s_raw);
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_raw);
    __shared__ int s_id[BATCH];
    const int v = blockIdx.z;
    const int tiles_x = gridDim.x, tiles_y = gridDim.y;
    const int tile = blockIdx.y * tiles_x + blockIdx.x;
    const size_t gt = (size_t)v * tiles_x * tiles_y + tile;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const WarpBlock wb = warp_block(blockIdx.x, blockIdx.y, tid);
    const int px = wb.px, py = wb.py;
    const bool inside = (px < a.W) && (py < a.H);
    const float2 neg_pix = make_float2(-(float)px, -(float)py);
    const bool overflow = w.counters[0] > a.inst_cap;
    const int n = overflow ? 0 : w.tile_count[gt];
    const int off = overflow ? 0 : w.tile_offset[gt];
    // ---- prologue: depth sort of this tile's instance keys (K3 fused here: its barrier latency hides
    // behind the compositing of the other resident CTAs).  Keys (depth_bits<<32 | id) are unique, so
    // the bitonic network is deterministic; ids go to inst_sorted for the batches below and for the
    // backward.  Tiles above the shared-memory capacity: sort_oversize_tile (chunk sort + merges).
    int32_t first_id = -1;
    bool prestaged = false;
    const size_t vN = (size_t)v * a.N;
    if (AGS_RANKSORT && n > 0 && n <= BATCH) {
        // ---- single-batch tile (the common case): rank sort fused with the staging.  Thread t owns
        // key t, fetches that splat's record while it counts the keys in front of its own
        // (broadcast shared-memory reads, no barrier per stage), and stores the record at its rank.
        const uint64_t my = (tid < n) ? w.inst_key[off + tid] : ~0ull;
        if (tid < n) s_keys[tid] = my;
        const int id = (int)(uint32_t)(my & 0xffffffffu);
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
        if (tid < n) {
            g0 = ldg4(w.geom0 + vN + id);
            g1 = ldg4(w.geom1 + vN + id);
        }
        __syncthreads();
        int rank = 0;
        if (tid < n) {
#pragma unroll 4
            for (int j = 0; j < n; ++j) rank += (s_keys[j] < my) ? 1 : 0;
        }
        __syncthreads();            // every key has been read: the records may overwrite them
        if (tid < n) {
            s_id[rank] = id;
            SplatRec& r = s_rec[rank];
            stage_geom(r, g0, g1);
            r.f0 = ldg4(w.feat0 + vN + id);
            r.f1 = ldg4(w.feat1 + vN + id);
            r.bb = splat_bbox(g0, g1);
            w.inst_sorted[off + rank] = id;                                       // the backward walks the same order
            if (w.inst_rec) store_rec(w.inst_rec, off + rank, r);                 // ... and can bulk-copy the records
        }
        prestaged = true;
    } else if (n > 0 && n <= FUSED_SORT_MAX) {
        const uint64_t* keys = w.inst_key + off;
        int m = 2;
        while (m < n) m <<= 1;
        if (m <= 128) {
            // small tile (the common case): ONE warp runs the network with __syncwarp only; the other
            // seven warps wait at the barrier below without consuming issue slots
            if (tid < 32) {
                for (int k = lane; k < m; k += 32) s_keys[k] = (k < n) ? keys[k] : ~0ull;
                __syncwarp();
                for (int kk = 2; kk <= m; kk <<= 1) {
                    for (int j = kk >> 1; j > 0; j >>= 1) {
                        for (int t = lane; t < (m >> 1); t += 32) {
                            const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int hi = lo | j;
                            const bool asc = ((lo & kk) == 0);
                            const uint64_t x = s_keys[lo], y = s_keys[hi];
                            if ((x > y) == asc) { s_keys[lo] = y; s_keys[hi] = x; }
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
        } else {
            for (int k = tid; k < m; k += 256) s_keys[k] = (k < n) ? keys[k] : ~0ull;
            __syncthreads();
            for (int kk = 2; kk <= m; kk <<= 1) {
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (m >> 1); t += 256) {
                        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const int hi = lo | j;
                        const bool asc = ((lo & kk) == 0);
                        const uint64_t x = s_keys[lo], y = s_keys[hi];
                        if ((x > y) == asc) { s_keys[lo] = y; s_keys[hi] = x; }
                    }
                    __syncthreads();
                }
            }
        }
        // the first batch takes its ids straight from shared memory (register), so the global
        // store of the sorted ids (needed by later batches and by the backward) is off the critical path
        if (tid < n) first_id = (int32_t)(s_keys[tid] & 0xffffffffu);
        for (int k = tid; k < n; k += 256) w.inst_sorted[off + k] = (int32_t)(s_keys[k] & 0xffffffffu);
        __syncthreads();            // ids visible to the whole CTA; s_raw free for the records
    } else if (n > FUSED_SORT_MAX) {
        sort_oversize_tile(w, n, off, s_keys, tid);     // ids in inst_sorted (global), visible after its barrier
    }
    const size_t P = (size_t)a.H * a.W;
    const size_t pix = (size_t)py * a.W + px;
    bool imp_pix = false;
    if (WANT_IMP && inside) imp_pix = a.render_mask ? (a.render_mask[(size_t)v * P + pix] == 1.f) : true;

    // T > 0: the pixel is live; T < 0: finished, |T| is its final transmittance (pixels outside the image start finished)
    float T = inside ? 1.f : -1.f;
    // accumulators as register pairs: the eight weighted sums of a splat are four packed FFMA2 (sm_100 fp32x2: two
    // FMAs per issue slot; the kernel is issue-bound, not FP32-pipe-bound) over the natural halves of the 128-bit
    // record loads: (r,g) (b,plane depth) (nx,ny) (nz,confidence)
    float2 CA = make_float2(0.f, 0.f), CB = CA, NA = CA, NB = CA;     // (C0,C1) (C2,D) (N0,N1) (N2,Cf)
    int last = 0;
    bool warp_done = __all_sync(0xffffffffu, !(T > 0.f));
    for (int base = 0; base < n; base += BATCH) {
        if (__syncthreads_and(warp_done)) break;
        const int j = base + tid;
        if (j < n && !prestaged) {
            const int id = (base == 0 && first_id >= 0) ? first_id : w.inst_sorted[off + j];
            const size_t idx = vN + id;
            const float4 g0 = ldg4(w.geom0 + idx), g1 = ldg4(w.geom1 + idx);
            s_id[tid] = id;
            SplatRec& r = s_rec[tid];
            stage_geom(r, g0, g1);
            r.f0 = ldg4(w.feat0 + idx);
            r.f1 = ldg4(w.feat1 + idx);
            r.bb = splat_bbox(g0, g1);
            if (w.inst_rec) store_rec(w.inst_rec, off + j, r);
        }
        if (!(prestaged && base == 0)) __syncthreads();   // (the rank-sorted first batch was ordered by the barrier above)
        const int cnt = min(BATCH, n - base);
        for (int c = 0; c < cnt && !warp_done; c += 32) {
            const int jj = c + lane;
            unsigned mask = __ballot_sync(0xffffffffu, jj < cnt && bbox_hits(s_rec[jj].bb, wb));
            while (mask) {
                const int k = c + __ffs(mask) - 1;
                mask &= mask - 1;
                const SplatRec& rec = s_rec[k];
                const float4 g0 = rec.g0, g1 = rec.g1;
                const SplatEval e = eval_alpha(g0, g1, neg_pix);
                // branch-free per lane: a lane that is done, skips the splat or stops here runs the same
                // arithmetic with weight 0 (one warp-uniform early-out when nobody takes the splat)
                const bool use = (T > 0.f) && !e.skip;
                if (__ballot_sync(0xffffffffu, use) == 0u) continue;
                const float test_T = T * (1.f - e.alpha);
                const bool stop = use && (test_T < AGS_T_EPS);      // stop BEFORE applying this splat
                const bool take = use && !stop;
                const float wgt = take ? e.alpha * T : 0.f;
                const float4 f0 = rec.f0, f1 = rec.f1;
                const float2 ww = make_float2(wgt, wgt);
                const float dpix = f0.w - g1.z * e.dx - g1.w * e.dy;
                CA = __ffma2_rn(ww, make_float2(f0.x, f0.y), CA);
                CB = __ffma2_rn(ww, make_float2(f0.z, dpix), CB);
                NA = __ffma2_rn(ww, make_float2(f1.x, f1.y), NA);
                NB = __ffma2_rn(ww, make_float2(f1.z, f1.w), NB);
                T = take ? test_T : (stop ? -T : T);
                last = take ? base + k + 1 : last;
                if (WANT_IMP) {
                    if (imp_pix && wgt > a.weight_thres) {
                        atomicAdd(a.count + vN + s_id[k], 1);
                        atomicAdd(a.importance + vN + s_id[k], wgt);
                    }
                }
            }
            warp_done = __all_sync(0xffffffffu, !(T > 0.f));
        }
    }
    if (inside) {
        T = fabsf(T);
        const float A = 1.f - T;
        const float bg0 = __ldg(a.bg), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
        float* o = a.out_rgb + (size_t)v * 3 * P + pix;
        o[0] = CA.x + T * bg0; o[P] = CA.y + T * bg1; o[2 * P] = CB.x + T * bg2;
        o = a.out_normal + (size_t)v * 3 * P + pix;
        o[0] = NA.x; o[P] = NA.y; o[2 * P] = NB.x;
        a.out_depth[(size_t)v * P + pix] = (A > 0.f) ? CB.y / A : 0.f;
        a.out_opacity[(size_t)v * P + pix] = A;
        a.out_confidence[(size_t)v * P + pix] = NB.y;
        w.final_T[(size_t)v * P + pix] = T;
        w.n_contrib[(size_t)v * P + pix] = last;
    }
}

// K5 ---------------------------------------------------------------------------------------------
// Warp reduction of 15 values per lane: lane L ends up with the warp-wide sum of value L>>1.
// Transposition through a per-warp shared buffer: every lane stores its 15 partials (conflict-free,
// row = value, column = lane), lane L then loads the 16 partials of value L>>1 that belong to the
// half-warp (L&1) with four 128-bit loads and one shuffle joins the halves.  36 instructions per
// (warp, splat) instead of 63 for the register butterfly (31 FSEL + 16 SHFL + 16 FADD).
// Row stride 36 floats keeps the 128-bit loads of a quarter-warp on distinct banks.
constexpr int RED_STRIDE = 36;

__device__ __forceinline__ float warp_reduce15(const float (&v)[15], unsigned st_addr, unsigned ld_addr, int lane) {
#define AGS_RED_ST(q) asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(st_addr), "n"((q) * RED_STRIDE * 4), "f"(v[q]) : "memory")
    AGS_RED_ST(0); AGS_RED_ST(1); AGS_RED_ST(2); AGS_RED_ST(3); AGS_RED_ST(4);
    AGS_RED_ST(5); AGS_RED_ST(6); AGS_RED_ST(7); AGS_RED_ST(8); AGS_RED_ST(9);
    AGS_RED_ST(10); AGS_RED_ST(11); AGS_RED_ST(12); AGS_RED_ST(13); AGS_RED_ST(14);
#undef AGS_RED_ST
    __syncwarp();
    float r = 0.f;
    if (lane < 30) {
        float4 x0, x1, x2, x3;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "r"(ld_addr) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(ld_addr) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+32];" : "=f"(x2.x), "=f"(x2.y), "=f"(x2.z), "=f"(x2.w) : "r"(ld_addr) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+48];" : "=f"(x3.x), "=f"(x3.y), "=f"(x3.z), "=f"(x3.w) : "r"(ld_addr) : "memory");
        r = ((x0.x + x0.y) + (x0.z + x0.w)) + ((x1.x + x1.y) + (x1.z + x1.w))
          + ((x2.x + x2.y) + (x2.z + x2.w)) + ((x3.x + x3.y) + (x3.z + x3.w));
    }
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    __syncwarp();
    return r;
}

// Variant 1: ONE register butterfly step first (lanes l and l^16 exchange half of their values), then
// the transposition runs on 8 values per lane over two half-warps: half the shared-memory bytes
// (8 STS.32 + 2 LDS.128 per lane instead of 15 + 4) for 8 SHFL + 16 SEL more.  Layout per warp: row =
// value (stride 24 floats), column = lane & 15, rows 8..15 shifted by 16 floats: conflict-free stores and
// quarter-warp-phased 128-bit loads (checked exhaustively offline).
constexpr int RED1_STRIDE = 24, RED1_HALF = 16, RED1_FLOATS = 16 * RED1_STRIDE + RED1_HALF;

__device__ __forceinline__ float warp_reduce15_half(const float (&v)[15], unsigned st_addr, unsigned ld_addr, int lane) {
    const bool upper = (lane & 16) != 0;
    float r8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float lo = v[i], hi = (i < 7) ? v[i + 8] : 0.f;
        const float send = upper ? lo : hi, keep = upper ? hi : lo;
        r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#define AGS_RED_ST(q) asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(st_addr), "n"((q) * RED1_STRIDE * 4), "f"(r8[q]) : "memory")
    AGS_RED_ST(0); AGS_RED_ST(1); AGS_RED_ST(2); AGS_RED_ST(3); AGS_RED_ST(4); AGS_RED_ST(5); AGS_RED_ST(6); AGS_RED_ST(7);
#undef AGS_RED_ST
    __syncwarp();
    float r = 0.f;
    if (lane < 30) {
        float4 x0, x1;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "r"(ld_addr) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+32];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(ld_addr) : "memory");
        r = ((x0.x + x0.y) + (x0.z + x0.w)) + ((x1.x + x1.y) + (x1.z + x1.w));
    }
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    __syncwarp();
    return r;
}

// Variant 2: the full register butterfly (no shared memory): 16 SHFL + 15 FADD + 30 SEL.
__device__ __forceinline__ float warp_reduce15_shfl(const float (&v)[15], int lane) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 15; ++i) a[i] = v[i];
    a[15] = 0.f;
#pragma unroll
    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = upper ? a[i] : a[i + w], keep = upper ? a[i + w] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// Per-pixel state of the backward walk.
struct BwdPix {
    float gC0, gC1, gC2, gN0, gN1, gN2, gD, gCf, rem, T, pyf;
    int last;
};

__device__ __forceinline__ void bwd_load_pixel(BwdPix& s, const AgsRenderArgs& a, const AgsRenderGradArgs& gr,
                                               const AgsWorkspace& w, int v, int px, int py) {
    s.gC0 = s.gC1 = s.gC2 = s.gN0 = s.gN1 = s.gN2 = s.gD = s.gCf = 0.f;
    s.rem = 0.f; s.T = 1.f; s.last = 0; s.pyf = (float)py;
    if (px >= a.W || py >= a.H) return;
    const size_t P = (size_t)a.H * a.W;
    const size_t pix = (size_t)py * a.W + px;
    const size_t vp = (size_t)v * P + pix;
    s.last = w.n_contrib[vp];
    const float Tf = w.final_T[vp];
    const float A = 1.f - Tf;
    if (gr.d_rgb) { const float* p = gr.d_rgb + (size_t)v * 3 * P + pix; s.gC0 = p[0]; s.gC1 = p[P]; s.gC2 = p[2 * P]; }
    if (gr.d_normal) { const float* p = gr.d_normal + (size_t)v * 3 * P + pix; s.gN0 = p[0]; s.gN1 = p[P]; s.gN2 = p[2 * P]; }
    const float gdep = gr.d_depth ? gr.d_depth[vp] : 0.f;
    float gA = gr.d_opacity ? gr.d_opacity[vp] : 0.f;
    if (gr.d_confidence) s.gCf = gr.d_confidence[vp];
    const float depth_out = a.out_depth[vp];
    if (A > 0.f) { s.gD = gdep / A; gA -= gdep * depth_out / A; }
    const float bg0 = __ldg(a.bg), bg1 = __ldg(a.bg + 1), bg2 = __ldg(a.bg + 2);
    const float* c = a.out_rgb + (size_t)v * 3 * P + pix;
    const float* nn = a.out_normal + (size_t)v * 3 * P + pix;
    const float bgdot = s.gC0 * bg0 + s.gC1 * bg1 + s.gC2 * bg2;
    const float S_all = s.gC0 * (c[0] - Tf * bg0) + s.gC1 * (c[P] - Tf * bg1) + s.gC2 * (c[2 * P] - Tf * bg2)
                      + s.gN0 * nn[0] + s.gN1 * nn[P] + s.gN2 * nn[2 * P]
                      + s.gD * (depth_out * A) + s.gCf * a.out_confidence[vp];
    s.rem = S_all + Tf * (bgdot - gA);
}

// One pixel's contribution to the 15 per-splat partials.  Branch-free: an inactive pixel runs the same
// arithmetic with alpha = G = 0, which leaves its T / rem untouched and adds exact zeros.
// The partials are the RAW MOMENTS of the pixel gradients (the chain rule through the conic, the centre
// and the plane slopes is linear in them and is applied once per splat in project_bwd):
//   [0] sum dpower*dx   [1] sum dpower*dy   [2] sum dpower*dx^2   [3] sum dpower*dx*dy   [4] sum dpower*dy^2
//   [5] sum dpower (d opacity = sum / o: alpha = o*G)   [6..8] sum w*gC   [9..11] sum w*gN   [12] sum w*gD
//   [13] sum w*gD*dx   [14] sum w*gD*dy        (dpower = dL/d(natural exponent), w = alpha*T, dx = x_splat - x_pixel)
// They land in the 16-float gradient record at the slots of AGS_REC_* (ags_common.cuh), shared with project_bwd.
template <bool FIRST, bool HAS_CONF>
__device__ __forceinline__ void bwd_accumulate(float (&val)[15], BwdPix& s, const float4 g1, const float4 f0,
                                               const float4 f1, float dx, float dy, float e_alpha, float e_G,
                                               bool active) {
    const float alpha = active ? e_alpha : 0.f;
    const float G = active ? e_G : 0.f;
    const float wgt = alpha * s.T;
    const float one_m = 1.f - alpha;
    const float dpix = f0.w - g1.z * dx - g1.w * dy;
    float sdot = s.gC0 * f0.x + s.gC1 * f0.y + s.gC2 * f0.z + s.gN0 * f1.x + s.gN1 * f1.y + s.gN2 * f1.z + s.gD * dpix;
    if (HAS_CONF) sdot += s.gCf * f1.w;
    s.rem -= wgt * sdot;
    const float dalpha = s.T * sdot - s.rem * rcp_approx(one_m);         // one_m >= 0.01
    s.T *= one_m;
    // alpha = min(0.99, o*G): clamped -> no gradient
    const bool unclamped = (g1.y * G <= AGS_ALPHA_MAX);
    const float dpower = unclamped ? alpha * dalpha : 0.f;
    const float dop = dpower;
    const float wgD = wgt * s.gD;
    const float pdx = dpower * dx, pdy = dpower * dy;
    if (FIRST) {
        val[0] = pdx; val[1] = pdy; val[2] = pdx * dx; val[3] = pdx * dy; val[4] = pdy * dy; val[5] = dop;
        val[6] = wgt * s.gC0; val[7] = wgt * s.gC1; val[8] = wgt * s.gC2;
        val[9] = wgt * s.gN0; val[10] = wgt * s.gN1; val[11] = wgt * s.gN2;
        val[12] = wgD; val[13] = wgD * dx; val[14] = wgD * dy;
    } else {
        val[0] += pdx; val[1] += pdy; val[2] += pdx * dx; val[3] += pdx * dy; val[4] += pdy * dy; val[5] += dop;
        val[6] += wgt * s.gC0; val[7] += wgt * s.gC1; val[8] += wgt * s.gC2;
        val[9] += wgt * s.gN0; val[10] += wgt * s.gN1; val[11] += wgt * s.gN2;
        val[12] += wgD; val[13] += wgD * dx; val[14] += wgD * dy;
    }
}

// Variant 3 (default): the register butterfly with the lane-dependent selects folded away.
// A butterfly level pairs slot i with slot i+w and every lane keeps one of the two depending on one bit of
// its lane id -- two selects per pair, 30 per reduction.  Seven of the fifteen partials are w * (a per-pixel
// constant): if lane l holds the constants PERMUTED by its role bits, Kp[t] = K[t ^ m(l)], every lane can
// keep slot t and send slot t+w at every level and the right values still meet (partner m differs in exactly
// the bit that swaps the halves): no selects for that group of eight through three levels.  The other
// eight partials (the dpower / depth-slope moments) use the d1/d2 trick at the first level (dx and dy
// swapped per lane once) and plain selects below.  12 selects instead of 30.
//   roles: r1 = lane bit 4 (xor 16), r2 = bit 3 (xor 8), r3 = bit 2 (xor 4), r4 = bit 1 (xor 2); bit 0 is summed.
//   the lane ends with record slot q = r4*8 + r1*4 + r2*2 + r3 (AGS_REC_* in ags_common.cuh).
struct FoldLane {
    float Kp[4];       // permuted per-pixel constants of the four W-slots this lane keeps at level 1: K[t ^ m(lane)],
                       // K = (gC0, gC1, gC2, gN0, gN1, gN2, gD, 0)
    float Kq[4];       // the same four slots of the level-1 PARTNER pixel (lane ^ 16: same column, two rows away)
    float gDq;         // the partner pixel's depth gradient
    float2 cc;         // partner's coordinates relative to this lane's, in the lane's role axes: (0,-2) | (+2,0)
    bool r1, r2, r3, r4;
};

// Level 1 of the butterfly does not exchange PRODUCTS (8 shuffles) but the two scalars they are made of (w = alpha*T and
// dL/d(exponent): 2 shuffles): the partner's share of every kept sum is that scalar times a constant of the partner
// PIXEL -- its gradient constants, fetched once per tile here, and its coordinates, which differ from this lane's by a
// known offset (same column, row +-2).
__device__ __forceinline__ void fold_setup(FoldLane& f, const BwdPix& s, int lane) {
    const unsigned FULL = 0xffffffffu;
    f.r1 = (lane & 16) != 0; f.r2 = (lane & 8) != 0; f.r3 = (lane & 4) != 0; f.r4 = (lane & 2) != 0;
    float K[8] = {s.gC0, s.gC1, s.gC2, s.gN0, s.gN1, s.gN2, s.gD, 0.f};
    float Kp[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) Kp[t] = f.r1 ? K[t ^ 4] : K[t];
#pragma unroll
    for (int t = 0; t < 8; ++t) K[t] = Kp[t];
#pragma unroll
    for (int t = 0; t < 8; ++t) Kp[t] = f.r2 ? K[t ^ 2] : K[t];
#pragma unroll
    for (int t = 0; t < 8; ++t) K[t] = Kp[t];
#pragma unroll
    for (int t = 0; t < 8; ++t) Kp[t] = f.r3 ? K[t ^ 1] : K[t];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        f.Kp[t] = Kp[t];
        f.Kq[t] = __shfl_xor_sync(FULL, Kp[t + 4], 16);      // what the partner used to multiply its w with before sending
    }
    f.gDq = __shfl_xor_sync(FULL, s.gD, 16);
    f.cc = f.r1 ? make_float2(2.f, 0.f) : make_float2(0.f, -2.f);
}

template <bool HAS_CONF>
__device__ __forceinline__ float bwd_pair_fold(BwdPix& s, const FoldLane& f, const float4 g1, const float4 f0,
                                               const float4 f1, float dx, float dy, float e_alpha, float e_G,
                                               bool active) {
    // Packed fp32x2 (FFMA2 / FMUL2 / FADD2, sm_100): two fp32 operations per issue slot on natural register pairs --
    // the halves of the 128-bit record loads, the per-pixel constants, the permuted constants.  The kernel is
    // issue-bound (83 % issue-active, FP32 pipe 32 %), so every pair saves a slot.
    const unsigned FULL = 0xffffffffu;
    const float alpha = active ? e_alpha : 0.f;
    const float G = e_G;          // no select: an inactive lane has alpha = 0, so its dpower = alpha * dalpha is 0 either way
    const float wgt = alpha * s.T;
    const float one_m = 1.f - alpha;
    const float dpix = f0.w - g1.z * dx - g1.w * dy;
    float2 acc = __fmul2_rn(make_float2(s.gC0, s.gC1), make_float2(f0.x, f0.y));
    acc = __ffma2_rn(make_float2(s.gC2, s.gD), make_float2(f0.z, dpix), acc);
    acc = __ffma2_rn(make_float2(s.gN0, s.gN1), make_float2(f1.x, f1.y), acc);
    if (HAS_CONF) acc = __ffma2_rn(make_float2(s.gN2, s.gCf), make_float2(f1.z, f1.w), acc);
    else acc.x = fmaf(s.gN2, f1.z, acc.x);
    const float sdot = acc.x + acc.y;
    s.rem -= wgt * sdot;
    const float dalpha = s.T * sdot - s.rem * rcp_approx(one_m);         // one_m >= 0.01
    s.T *= one_m;
    const bool unclamped = (g1.y * G <= AGS_ALPHA_MAX);                   // alpha = min(0.99, o*G): clamped -> no gradient
    const float dpower = unclamped ? alpha * dalpha : 0.f;
    const float wgD = wgt * s.gD;
    // ---- level 1 (xor 16): two scalars cross the lanes, the partner's products are formed here
    const float wgq = __shfl_xor_sync(FULL, wgt, 16), dpq = __shfl_xor_sync(FULL, dpower, 16);
    const float2 W01 = __ffma2_rn(make_float2(wgq, wgq), make_float2(f.Kq[0], f.Kq[1]),
                                  __fmul2_rn(make_float2(wgt, wgt), make_float2(f.Kp[0], f.Kp[1])));
    const float2 W23 = __ffma2_rn(make_float2(wgq, wgq), make_float2(f.Kq[2], f.Kq[3]),
                                  __fmul2_rn(make_float2(wgt, wgt), make_float2(f.Kp[2], f.Kp[3])));
    float P[4];
    {
        const float2 dd = make_float2(f.r1 ? dy : dx, f.r1 ? dx : dy);   // (d1, d2): this pixel, role axis first
        const float2 dq = __fadd2_rn(dd, f.cc);                          // the partner pixel's (d1, d2)
        const float k0 = dpower * dd.x, k0q = dpq * dq.x;                // dpower*d1 of both pixels
        const float2 B = __fadd2_rn(__fmul2_rn(make_float2(k0, k0), dd), __fmul2_rn(make_float2(k0q, k0q), dq));
        P[0] = k0 + k0q;                                                 // sum dpower*dx   | dpower*dy
        P[1] = B.x;                                                      // sum dpower*dx^2 | dpower*dy^2
        P[2] = fmaf(wgq * f.gDq, dq.x, wgD * dd.x);                      // sum wgD*dx      | wgD*dy
        P[3] = f.r1 ? dpower + dpq : B.y;                                // sum dpower*dx*dy | dpower
    }
    // ---- level 2 (xor 8)
    const float2 W2 = __fadd2_rn(W01, make_float2(__shfl_xor_sync(FULL, W23.x, 8), __shfl_xor_sync(FULL, W23.y, 8)));
    float P2[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float send = f.r2 ? P[t] : P[t + 2], keep = f.r2 ? P[t + 2] : P[t];
        P2[t] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    // ---- level 3 (xor 4)
    const float W1 = W2.x + __shfl_xor_sync(FULL, W2.y, 4);
    const float send3 = f.r3 ? P2[0] : P2[1], keep3 = f.r3 ? P2[1] : P2[0];
    const float P1 = keep3 + __shfl_xor_sync(FULL, send3, 4);
    // ---- level 4 (xor 2): the lane keeps the W-group value (r4 = 0) or the P-group value (r4 = 1)
    const float send4 = f.r4 ? W1 : P1, keep4 = f.r4 ? P1 : W1;
    float r = keep4 + __shfl_xor_sync(FULL, send4, 2);
    r += __shfl_xor_sync(FULL, r, 1);
    return r;
}

// Variant 4: the cross-lane reduction as a TENSOR-CORE matrix product (mma.sync m16n8k8, TF32 split into
// hi + lo words for fp32-level accuracy).  Every one of the 15 per-splat sums is linear in two per-lane scalars,
//      sum_l wgt_l * (gC, gN, gD, gD*px)_l       and      sum_l dpower_l * (1, px, py, px^2, px*py, py^2)_l,
// where the second factors do not depend on the splat once the pixel coordinates are taken relative to the
// warp block's centre (px in +-3.5, py in +-1.5; the moments about the SPLAT centre follow from the binomial
// expansion, applied once per (warp block, splat) after the product).  So a warp parks (wgt, dpower) of 8
// splats in shared memory (2 STS per lane and splat) and then multiplies
//      D (16 sums x 8 splats) = A (16 x 64: the per-lane constants) * B (64 x 8: the parked scalars)
// with 20 MMAs: 2.5 per splat instead of 16 SHFL + 15 FADD + 12 FSEL + 16 FMUL.  Rows of D:
//   0-2 sum w gC | 3-5 sum w gN | 6 sum w gD | 7 sum w gD px | 8 sum dp | 9 sum dp px | 10 sum dp py |
//   11 sum dp px^2 | 12 sum dp px py | 13 sum dp py^2 | 14 sum w gD (py - 0.5)   (k-block j holds row j of the
//   8x4 block, so py - 0.5 = j - 2 scales the TF32 words of row 6 exactly) | 15 unused.
// The position-only constants (rows 8-13) are exact in TF32, so that group needs 2 passes instead of 3.
constexpr int MMA_N = 8;            // splats per product
constexpr int MMA_XS = 36;          // floats per staging row: lane stores conflict-free, fragment loads (4g + 8j + t) too

struct __align__(16) MmaStage {
    float x[2][MMA_N][MMA_XS];     // [wgt | dpower][splat][lane]
    int pend[MMA_N];                // batch slot of every parked splat
};

__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// same conversion, pinned where it is written: the constants' words are recomputed per product on purpose
// (hoisting them out of the splat loop costs 24 registers and with them a resident CTA per SM)
__device__ __forceinline__ uint32_t tf32_hi_pinned(float x) {
    uint32_t r;
    asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct MmaConsts {
    float w[4][2];                  // rows 0-7 (this thread's row g = lane>>2) for lanes 8j+t and 8j+t+4; split into TF32 words per product
};

// the per-lane constants of rows 0-7 travel through the (still unused) staging buffer once per tile
__device__ __forceinline__ void mma_setup(MmaConsts& c, MmaStage& st, const BwdPix& s, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const float pxr = (float)(lane & 7) - 3.5f;
    float* tmp = &st.x[0][0][0];
    const float K[8] = {s.gC0, s.gC1, s.gC2, s.gN0, s.gN1, s.gN2, s.gD, s.gD * pxr};
#pragma unroll
    for (int r = 0; r < 8; ++r) tmp[r * MMA_XS + lane] = K[r];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) c.w[j][h] = tmp[g * MMA_XS + 8 * j + t + 4 * h];
    __syncwarp();
}

// Multiply out the `count` parked splats and add the converted sums to their gradient records.
__device__ __forceinline__ void mma_flush(const MmaConsts& c, MmaStage& st, int count, const SplatRec* s_rec,
                                          const int* s_id, float* dsplat_view, float cx, float cy, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // rows 8-13: a = u + v*py + w2*py^2 with exact small numbers (px = t - 3.5 | t + 0.5, py = j - 1.5)
    const float x1 = (float)t - 3.5f, x3 = (float)t + 0.5f;
    const float u1 = g == 0 ? 1.f : g == 1 ? x1 : g == 3 ? x1 * x1 : 0.f;
    const float u3 = g == 0 ? 1.f : g == 1 ? x3 : g == 3 ? x3 * x3 : 0.f;
    const float v1 = g == 2 ? 1.f : g == 4 ? x1 : 0.f;
    const float v3 = g == 2 ? 1.f : g == 4 ? x3 : 0.f;
    const float w2 = g == 5 ? 1.f : 0.f;
    const float f14 = g == 6 ? 1.f : 0.f;                      // the threads that own row 6 also feed row 14
    __syncwarp();
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        {   // wgt group: rows 0-7 (a0, a2) and row 14 (a1, a3 of the row-6 threads), 3 passes
            const float x0 = st.x[0][g][8 * j + t], xb = st.x[0][g][8 * j + t + 4];
            const uint32_t b0h = tf32_hi(x0), b1h = tf32_hi(xb);
            const uint32_t b0l = __float_as_uint(x0 - __uint_as_float(b0h)), b1l = __float_as_uint(xb - __uint_as_float(b1h));
            const float m = f14 * (float)(j - 2);
            const uint32_t a0h = tf32_hi_pinned(c.w[j][0]), a2h = tf32_hi_pinned(c.w[j][1]);
            const float a0lf = c.w[j][0] - __uint_as_float(a0h), a2lf = c.w[j][1] - __uint_as_float(a2h);
            const uint32_t a0l = __float_as_uint(a0lf), a2l = __float_as_uint(a2lf);
            const uint32_t a1h = __float_as_uint(m * __uint_as_float(a0h)), a3h = __float_as_uint(m * __uint_as_float(a2h));
            const uint32_t a1l = __float_as_uint(m * a0lf), a3l = __float_as_uint(m * a2lf);
            mma_tf32(d, a0h, a1h, a2h, a3h, b0h, b1h);
            mma_tf32(d, a0l, a1l, a2l, a3l, b0h, b1h);
            mma_tf32(d, a0h, a1h, a2h, a3h, b0l, b1l);
        }
        {   // dpower group: rows 8-13 (a1, a3), exact constants, 2 passes
            const float x0 = st.x[1][g][8 * j + t], xb = st.x[1][g][8 * j + t + 4];
            const uint32_t b0h = tf32_hi(x0), b1h = tf32_hi(xb);
            const uint32_t b0l = __float_as_uint(x0 - __uint_as_float(b0h)), b1l = __float_as_uint(xb - __uint_as_float(b1h));
            const float py = (float)j - 1.5f;
            const uint32_t a1 = __float_as_uint(u1 + py * (v1 + py * w2)), a3 = __float_as_uint(u3 + py * (v3 + py * w2));
            mma_tf32(d, 0u, a1, 0u, a3, b0h, b1h);
            mma_tf32(d, 0u, a1, 0u, a3, b0l, b1l);
        }
        asm volatile("" ::: "memory");      // one k-block at a time: keeps the fragment loads from piling up in registers
    }
    // d[0], d[1]: row g of splats 2t, 2t+1;  d[2], d[3]: row g+8.  Moments about the block centre -> about the splat centre
    const int slot_lo = g < 7 ? g : AGS_REC_WDX;               // rows 0-6 sit at their record slots (AGS_REC_C0/N0/WD)
    const int slot_hi = g == 0 ? AGS_REC_P1 : g == 1 ? AGS_REC_PDX : g == 2 ? AGS_REC_PDY : g == 3 ? AGS_REC_PXX
                      : g == 4 ? AGS_REC_PXY : g == 5 ? AGS_REC_PYY : AGS_REC_WDY;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int n = 2 * t + e;
        const float lo = d[e], hi = d[2 + e];
        const float P1 = __shfl_sync(0xffffffffu, hi, t), Ppx = __shfl_sync(0xffffffffu, hi, 4 + t),
                    Ppy = __shfl_sync(0xffffffffu, hi, 8 + t), WD = __shfl_sync(0xffffffffu, lo, 24 + t);
        if (n < count) {
            const int k = st.pend[n];
            const float4 g0 = s_rec[k].g0;
            const float gx = g0.x - cx, gy = g0.y - cy;          // splat centre relative to the block centre
            float out_hi;
            switch (g) {
                case 0: out_hi = hi; break;
                case 1: out_hi = gx * P1 - hi; break;
                case 2: out_hi = gy * P1 - hi; break;
                case 3: out_hi = gx * (gx * P1 - 2.f * Ppx) + hi; break;
                case 4: out_hi = gx * (gy * P1 - Ppy) - gy * Ppx + hi; break;
                case 5: out_hi = gy * (gy * P1 - 2.f * Ppy) + hi; break;
                default: out_hi = (gy - 0.5f) * lo - hi; break;   // row 14 = sum w gD (py - 0.5); only g == 6 stores it
            }
            const float out_lo = g < 7 ? lo : gx * WD - lo;
            float* rec = dsplat_view + (size_t)s_id[k] * 16;
            atomicAdd(rec + slot_lo, out_lo);
            if (g < 7) atomicAdd(rec + slot_hi, out_hi);
        }
    }
    __syncwarp();
}

// K5.  CTA = one 16x16 tile of one view, 256 / PX threads: every warp owns a block of 8 x (4*PX) pixels,
// every lane PX pixels of one column (rows y, y+4, ...).  The partials of a lane's pixels are summed in
// registers BEFORE the cross-lane reduction, so a splat costs one reduction + one 15-lane RED per
// (warp block, splat) -- with PX = 4 half as many as with 8x4 blocks -- and the loop control, the record
// loads and the bounding-box test are shared by the PX pixels.  Sub-blocks of 8x4 pixels the splat's
// cutoff box does not reach are skipped warp-uniformly.
// TMA = true: the batches are NOT gathered by the threads; composite_fwd left the tile's depth-sorted staging
// records contiguous in global memory (w.inst_rec) and one elected thread copies each batch of up to 256
// records (20 KB) into shared memory with ONE cp.async.bulk (TMA, completion on an mbarrier), double
// buffered: the copy of batch b+1 is in flight while the warps composite batch b.
template <int PX, bool HAS_CONF, int RED, bool TMA>
__global__ void __launch_bounds__(256 / PX, PX == 1 ? (RED == 4 ? AGS_BWD_MMA_MINB : AGS_BWD_MINB) : (PX == 2 ? AGS_BWD_PX2_MINB : 8))
composite_bwd_kernel(AgsRenderArgs a, AgsRenderGradArgs gr, AgsWorkspace w) {
    constexpr int THREADS = 256 / PX, WARPS = 8 / PX;
    constexpr int RED_FLOATS = RED == 0 ? 15 * RED_STRIDE : (RED == 1 ? RED1_FLOATS : (RED == 4 ? (int)(sizeof(MmaStage) / 4) : 4));
    static_assert((RED != 3 && RED != 4) || PX == 1, "the folded butterfly and the MMA reduction are written for one pixel per lane");
    __shared__ __align__(128) SplatRec s_rec_all[TMA ? 2 * BATCH : BATCH];
    __shared__ __align__(8) uint64_t s_bar[2];
    SplatRec* s_rec = s_rec_all;
    __shared__ int s_id[BATCH];
    __shared__ int s_max_last;
    __shared__ __align__(16) float s_red[WARPS][RED_FLOATS];   // per-warp transposition buffer
    const int v = blockIdx.z;
    const int tiles_x = gridDim.x, tiles_y = gridDim.y;
    const int tile = blockIdx.y * tiles_x + blockIdx.x;
    const size_t gt = (size_t)v * tiles_x * tiles_y + tile;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    if (w.counters[0] > a.inst_cap) return;
    const int n = w.tile_count[gt];
    if (n == 0) return;
    const int off = w.tile_offset[gt];
    const size_t vN = (size_t)v * a.N;
    const int bx = blockIdx.x * TILE + (wid & 1) * 8, by = blockIdx.y * TILE + (wid >> 1) * 4 * PX;
    WarpBlock wb;
    wb.px = bx + (lane & 7); wb.py = by + (lane >> 3);
    wb.x0 = (float)bx; wb.x1 = (float)(bx + 7); wb.y0 = (float)by; wb.y1 = (float)(by + 4 * PX - 1);
    const float pxf = (float)wb.px;
    BwdPix s[PX];
    int lane_last = 0;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
        bwd_load_pixel(s[j], a, gr, w, v, wb.px, wb.py + 4 * j);
        lane_last = max(lane_last, s[j].last);
    }
    if (tid == 0) s_max_last = 0;
    __syncthreads();
    if (lane_last > 0) atomicMax(&s_max_last, lane_last);
    __syncthreads();
    const int n_eff = min(n, s_max_last);
    const int warp_last = __reduce_max_sync(0xffffffffu, lane_last);
    // loop-invariant shared-space addresses of this lane's slots in the warp's transposition buffer
    unsigned red_st, red_ld;
    if (RED == 1) {
        red_st = (unsigned)__cvta_generic_to_shared(&s_red[wid][(lane >> 4) * (8 * RED1_STRIDE + RED1_HALF) + (lane & 15)]);
        red_ld = (unsigned)__cvta_generic_to_shared(
            &s_red[wid][(lane >> 1) * RED1_STRIDE + (lane >> 4) * RED1_HALF + (lane & 1) * 4]);
    } else {
        red_st = (unsigned)__cvta_generic_to_shared(&s_red[wid][RED == 0 ? lane : 0]);
        red_ld = (unsigned)__cvta_generic_to_shared(&s_red[wid][RED == 0 ? (lane >> 1) * RED_STRIDE + (lane & 1) * 16 : 0]);
    }
    // record slot this lane's reduced value goes to (AGS_REC_*): the folded butterfly ends with slot
    // r4*8 + r1*4 + r2*2 + r3; the others with partial (lane >> 1) in the order of bwd_accumulate
    int rec_slot;
    bool rec_lane;
    FoldLane fold;
    if (RED == 3) {
        fold_setup(fold, s[0], lane);
        rec_slot = ((lane >> 1) & 1) * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        rec_lane = (lane & 1) == 0 && rec_slot != AGS_REC_PAD;
    } else {
        const int q = lane >> 1;
        rec_slot = q == 0 ? AGS_REC_PDX : q == 1 ? AGS_REC_PDY : q == 2 ? AGS_REC_PXX : q == 3 ? AGS_REC_PXY
                 : q == 4 ? AGS_REC_PYY : q == 5 ? AGS_REC_P1 : q == 6 ? AGS_REC_C0 : q == 7 ? AGS_REC_C0 + 1
                 : q == 8 ? AGS_REC_C0 + 2 : q == 9 ? AGS_REC_N0 : q == 10 ? AGS_REC_N0 + 1 : q == 11 ? AGS_REC_N0 + 2
                 : q == 12 ? AGS_REC_WD : q == 13 ? AGS_REC_WDX : AGS_REC_WDY;
        rec_lane = (lane & 1) == 0 && lane < 30;
    }
    float* const dsplat_lane = w.dsplat + vN * 16 + rec_slot;
    MmaConsts mmac;
    MmaStage& mst = *reinterpret_cast<MmaStage*>(&s_red[wid][0]);   // RED == 4: the warp's buffer parks (wgt, dpower) of 8 splats
    int parked = 0;                                            // warp-uniform
    if (RED == 4) mma_setup(mmac, mst, s[0], lane);
    if (TMA) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned bytes = (unsigned)min(BATCH, n_eff) * (unsigned)sizeof(SplatRec);
            mbar_expect_tx(&s_bar[0], bytes);
            bulk_g2s(s_rec_all, w.inst_rec + (size_t)off * 5, bytes, &s_bar[0]);
        }
    }
    for (int base = 0; base < n_eff; base += BATCH) {
        __syncthreads();
        if (TMA) {
            const int b = base / BATCH, buf = b & 1;
            // everybody is done with batch b-1: its buffer takes batch b+1 while batch b is composited
            if (tid == 0 && base + BATCH < n_eff) {
                const unsigned bytes = (unsigned)min(BATCH, n_eff - base - BATCH) * (unsigned)sizeof(SplatRec);
                mbar_expect_tx(&s_bar[buf ^ 1], bytes);
                bulk_g2s(s_rec_all + (buf ^ 1) * BATCH, w.inst_rec + (size_t)(off + base + BATCH) * 5, bytes, &s_bar[buf ^ 1]);
            }
            for (int t = tid; t < BATCH; t += THREADS)
                if (base + t < n_eff) s_id[t] = w.inst_sorted[off + base + t];
            mbar_wait(&s_bar[buf], (unsigned)((b >> 1) & 1));
            s_rec = s_rec_all + buf * BATCH;
        } else {
#pragma unroll
            for (int h = 0; h < PX; ++h) {                       // THREADS threads stage BATCH records
                const int t = tid + h * THREADS;
                const int j = base + t;
                if (j < n_eff) {
                    const int id = w.inst_sorted[off + j];
                    const size_t idx = vN + id;
                    const float4 g0 = ldg4(w.geom0 + idx), g1 = ldg4(w.geom1 + idx);
                    s_id[t] = id;
                    SplatRec& r = s_rec[t];
                    stage_geom(r, g0, g1);
                    r.f0 = ldg4(w.feat0 + idx);
                    r.f1 = ldg4(w.feat1 + idx);
                    r.bb = splat_bbox(g0, g1);
                }
            }
        }
        __syncthreads();
        const int cnt = min(BATCH, min(n_eff, warp_last) - base);   // nothing beyond the warp's last contributor
        for (int c = 0; c < cnt; c += 32) {
            const int jj = c + lane;
            unsigned mask = __ballot_sync(0xffffffffu, jj < cnt && bbox_hits(s_rec[jj].bb, wb));
            while (mask) {
                const int k = c + __ffs(mask) - 1;
                mask &= mask - 1;
                const SplatRec& rec = s_rec[k];
                const float4 g0 = rec.g0, g1 = rec.g1;
                float dx = g0.x - pxf;            // shared by the lane's pixels of one column (PX == 1: replaced by the packed form's)
                // evaluate alpha for the lane's pixels; a sub-block (8x4) nobody is active in costs nothing more
                float eA[PX], eG[PX], eDy[PX];
                bool act[PX];
                unsigned any = 0u;
                float bz = 0.f, bw = 0.f;
                if (PX > 1) { const float4 bb = rec.bb; bz = bb.z; bw = bb.w; }
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    act[j] = false; eA[j] = 0.f; eG[j] = 0.f; eDy[j] = 0.f;
                    if (PX > 1 && !((bz <= wb.y0 + (float)(4 * j + 3)) && (bw >= wb.y0 + (float)(4 * j)))) continue;
                    float dxj, dy;
                    const float power = splat_power(g0, g1, make_float2(-pxf, -s[j].pyf), dxj, dy);
                    if (PX == 1) dx = dxj;
                    const float G = ex2_approx(power);
                    const float alpha = fminf(AGS_ALPHA_MAX, g1.y * G);
                    const bool skip = (power > 0.f) || (alpha < AGS_ALPHA_MIN);
                    act[j] = (base + k < s[j].last) && !skip;
                    eA[j] = alpha; eG[j] = G; eDy[j] = dy;
                    if (__ballot_sync(0xffffffffu, act[j])) any |= 1u << j;
                }
                if (any == 0u) continue;
                const float4 f0 = rec.f0, f1 = rec.f1;
                if (RED == 4) {                              // PX == 1 only (launcher): park the two scalars, multiply every 8 splats
                    BwdPix& sp = s[0];
                    const float alpha = act[0] ? eA[0] : 0.f;
                    const float G = act[0] ? eG[0] : 0.f;
                    const float dy = eDy[0];
                    const float wgt = alpha * sp.T;
                    const float one_m = 1.f - alpha;
                    const float dpix = f0.w - g1.z * dx - g1.w * dy;
                    float sdot = sp.gC0 * f0.x + sp.gC1 * f0.y + sp.gC2 * f0.z + sp.gN0 * f1.x + sp.gN1 * f1.y + sp.gN2 * f1.z + sp.gD * dpix;
                    if (HAS_CONF) sdot += sp.gCf * f1.w;
                    sp.rem -= wgt * sdot;
                    const float dalpha = sp.T * sdot - sp.rem * rcp_approx(one_m);
                    sp.T *= one_m;
                    const bool unclamped = (g1.y * G <= AGS_ALPHA_MAX);
                    mst.x[0][parked][lane] = wgt;
                    mst.x[1][parked][lane] = unclamped ? alpha * dalpha : 0.f;
                    if (lane == 0) mst.pend[parked] = k;
                    if (++parked == MMA_N) {
                        mma_flush(mmac, mst, MMA_N, s_rec, s_id, w.dsplat + vN * 16, wb.x0 + 3.5f, wb.y0 + 1.5f, lane);
                        parked = 0;
                    }
                    continue;
                }
                if (RED == 3) {                              // PX == 1 only (launcher)
                    const float r = bwd_pair_fold<HAS_CONF>(s[0], fold, g1, f0, f1, dx, eDy[0], eA[0], eG[0], act[0]);
                    if (rec_lane) atomicAdd(dsplat_lane + (size_t)s_id[k] * 16, r);
                    continue;
                }
                float val[15];
                if (PX == 1) {
                    bwd_accumulate<true, HAS_CONF>(val, s[0], g1, f0, f1, dx, eDy[0], eA[0], eG[0], act[0]);
                } else {
#pragma unroll
                    for (int q = 0; q < 15; ++q) val[q] = 0.f;
#pragma unroll
                    for (int j = 0; j < PX; ++j)
                        if (any & (1u << j))
                            bwd_accumulate<false, HAS_CONF>(val, s[j], g1, f0, f1, dx, eDy[j], eA[j], eG[j], act[j]);
                }
                const float r = RED == 0 ? warp_reduce15(val, red_st, red_ld, lane)
                              : (RED == 1 ? warp_reduce15_half(val, red_st, red_ld, lane) : warp_reduce15_shfl(val, lane));
                if ((lane & 1) == 0 && lane < 30) atomicAdd(dsplat_lane + (size_t)s_id[k] * 16, r);
            }
        }
        if (RED == 4 && parked > 0) {       // the parked splats refer to this batch's records: multiply them out before it is replaced
            mma_flush(mmac, mst, parked, s_rec, s_id, w.dsplat + vN * 16, wb.x0 + 3.5f, wb.y0 + 1.5f, lane);
            parked = 0;
        }
    }
}

}  // namespace

int ags_launch_composite_fwd(const AgsRenderArgs& a, const AgsWorkspace& w) {
    dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE, a.B);
    dim3 block(TILE * TILE);
    ags_note_launch();
    if (a.require_importance) composite_fwd_kernel<true><<<grid, block, 0, (cudaStream_t)a.stream>>>(a, w);
    else composite_fwd_kernel<false><<<grid, block, 0, (cudaStream_t)a.stream>>>(a, w);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Pixels per lane of the backward (1, 2 or 4; see composite_bwd_kernel).  AGS_BWD_PX in the environment
// overrides the default for tuning runs (read once).
static int bwd_px() {
    static int px = -1;
    if (px < 0) {
        const char* e = getenv("AGS_BWD_PX");
        px = e ? atoi(e) : AGS_BWD_PX_DEFAULT;
        if (px != 1 && px != 2 && px != 4) px = AGS_BWD_PX_DEFAULT;
    }
    return px;
}

// Cross-lane reduction of the 15 partials: 0 = shared-memory transposition, 1 = one shuffle step + half
// transposition, 2 = register butterfly, 3 = register butterfly with folded selects (one pixel per lane).
// AGS_BWD_RED overrides the default for tuning runs.
static int bwd_red() {
    static int red = -1;
    if (red < 0) {
        const char* e = getenv("AGS_BWD_RED");
        red = e ? atoi(e) : AGS_BWD_RED_DEFAULT;
        if (red < 0 || red > 4) red = AGS_BWD_RED_DEFAULT;
    }
    return red;
}

template <int PX, int RED>
static void launch_bwd2(const AgsRenderArgs& a, const AgsRenderGradArgs& g, const AgsWorkspace& w, dim3 grid) {
    ags_note_launch();
    if (PX == 1 && RED == 3 && w.inst_rec != nullptr) {      // TMA-staged batches (AGS_BWD_TMA=1)
        if (g.d_confidence) composite_bwd_kernel<1, true, 3, true><<<grid, 256, 0, (cudaStream_t)a.stream>>>(a, g, w);
        else composite_bwd_kernel<1, false, 3, true><<<grid, 256, 0, (cudaStream_t)a.stream>>>(a, g, w);
        return;
    }
    if (g.d_confidence) composite_bwd_kernel<PX, true, RED, false><<<grid, 256 / PX, 0, (cudaStream_t)a.stream>>>(a, g, w);
    else composite_bwd_kernel<PX, false, RED, false><<<grid, 256 / PX, 0, (cudaStream_t)a.stream>>>(a, g, w);
}

template <int PX>
static void launch_bwd(const AgsRenderArgs& a, const AgsRenderGradArgs& g, const AgsWorkspace& w, dim3 grid) {
    switch (bwd_red()) {
        case 0: launch_bwd2<PX, 0>(a, g, w, grid); break;
        case 1: launch_bwd2<PX, 1>(a, g, w, grid); break;
        case 3: if (PX == 1) { launch_bwd2<1, 3>(a, g, w, grid); break; }   // else: fall through
        case 4: if (PX == 1) { launch_bwd2<1, 4>(a, g, w, grid); break; }   // else: fall through
        default: launch_bwd2<PX, 2>(a, g, w, grid); break;
    }
}

int ags_launch_composite_bwd(const AgsRenderArgs& a, const AgsRenderGradArgs& g, const AgsWorkspace& w) {
    dim3 grid((a.W + TILE - 1) / TILE, (a.H + TILE - 1) / TILE, a.B);
    switch (bwd_px()) {
        case 1: launch_bwd<1>(a, g, w, grid); break;
        case 2: launch_bwd<2>(a, g, w, grid); break;
        default: launch_bwd<4>(a, g, w, grid); break;
    }
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
