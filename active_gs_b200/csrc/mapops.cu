// mapops.cu -- small device-side pieces of the map's per-keyframe bookkeeping that the reference does
// with many tiny ATen launches and host round trips.
#include <string.h>
#include "ags_common.cuh"

namespace {

struct IdList {
    int32_t id[AGS_MAX_BATCH];
};

// one thread per (batch slot, float of the camera row)
__global__ void stage_cameras_kernel(const float* __restrict__ table, IdList ids, int B,
                                     float* __restrict__ view, float* __restrict__ proj,
                                     float* __restrict__ tanfov) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * AGS_CAM_ROW) return;
    const int b = t / AGS_CAM_ROW, j = t - b * AGS_CAM_ROW;
    const float v = __ldg(table + (size_t)ids.id[b] * AGS_CAM_ROW + j);
    if (j < 16) view[b * 16 + j] = v;
    else if (j < 32) proj[b * 16 + j - 16] = v;
    else tanfov[b * 2 + j - 32] = v;
}

}  // namespace

extern "C" int ags_stage_cameras(const float* table, int32_t T, const int32_t* ids_host, int32_t B,
                                 float* viewmatrix, float* projmatrix, float* tanfov, void* stream) {
    AGS_CHECK_ARG(table && ids_host && viewmatrix && projmatrix && tanfov, "NULL argument");
    AGS_CHECK_ARG(B > 0 && B <= AGS_MAX_BATCH, "batch %d outside 1..%d", B, AGS_MAX_BATCH);
    IdList ids;
    for (int b = 0; b < B; ++b) {
        AGS_CHECK_ARG(ids_host[b] >= 0 && ids_host[b] < T, "keyframe id %d outside 0..%d", ids_host[b], T - 1);
        ids.id[b] = ids_host[b];
    }
    const int n = B * AGS_CAM_ROW;
    ags_note_launch(); stage_cameras_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table, ids, B, viewmatrix, projmatrix, tanfov);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// shared helpers
namespace {

constexpr int MO_THREADS = 256;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {          // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {          // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}

// Exclusive prefix (in thread order) of a per-thread flag over a MO_THREADS block; `total` = block sum.
__device__ __forceinline__ int block_scan_flag(bool flag, int& total) {
    __shared__ int s_w[MO_THREADS / 32];
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();                                              // previous use of s_w is over
    if (lane == 0) s_w[wid] = __popc(b);
    __syncthreads();
    int before = 0;
    total = 0;
#pragma unroll
    for (int k = 0; k < MO_THREADS / 32; ++k) {
        const int c = s_w[k];
        if (k < wid) before += c;
        total += c;
    }
    return before + __popc(b & ((1u << lane) - 1u));
}

// Exclusive scan of n block counts by ONE block of 1024 threads; total -> out_total[0] (clamped to
// `room` in out_total[0], unclamped in out_total[1] when `room` >= 0).
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t* __restrict__ counts, int32_t* __restrict__ offsets, int n,
                   int32_t* __restrict__ out_clamped, int32_t* __restrict__ out_total, int room) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = (i < n) ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_w[lane] = w;                                        // inclusive over warps
        }
        __syncthreads();
        const int warp_before = (wid == 0) ? 0 : s_w[wid - 1];
        const int carry = s_carry;
        if (i < n) offsets[i] = carry + warp_before + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + warp_before + incl;
        __syncthreads();
    }
    if (tid == 0) {
        const int total = s_carry;
        if (out_total) *out_total = total;
        if (out_clamped) *out_clamped = (room >= 0 && total > room) ? room : total;
    }
}

// =================================================================================================
// spawn (ags_spawn)
struct SpawnWs {
    unsigned long long* keys;     // (table) voxel keys, ~0 = empty
    unsigned long long* vals;     // (table) priority<<32 | pixel of the current winner
    int32_t* slot;                // (P) table slot of a candidate pixel, -1 otherwise
    float* cand_mean;             // (P,3)
    float* cand_rot;              // (P,4)
    int32_t* blk_count;           // (nblk)
    int32_t* blk_offset;          // (nblk)
    uint32_t table_mask;
    size_t total;
};

inline SpawnWs spawn_carve(void* base, int H, int W) {
    SpawnWs w;
    const size_t P = (size_t)H * W;
    size_t table = 1024;
    while (table < 2 * P) table <<= 1;
    const size_t nblk = (P + MO_THREADS - 1) / MO_THREADS;
    size_t off = 0;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { char* q = p + off; off += ags_align256(bytes); return (void*)q; };
    w.keys = (unsigned long long*)take(table * 8);
    w.vals = (unsigned long long*)take(table * 8);
    w.slot = (int32_t*)take(P * 4);
    w.cand_mean = (float*)take(P * 12);
    w.cand_rot = (float*)take(P * 16);
    w.blk_count = (int32_t*)take(nblk * 4);
    w.blk_offset = (int32_t*)take(nblk * 4);
    w.table_mask = (uint32_t)(table - 1);
    w.total = off;
    return w;
}

struct SpawnCam {
    float c2w[16];
    float Kinv[9];
};

// camera-space position of the (replicate-padded) pixel for depth2normal (quirk Q2 pairing kept)
__device__ __forceinline__ void d2n_pos(const float* __restrict__ ds, const float* __restrict__ draw, int x, int y,
                                        int H, int W, float k00, float k11, float& px, float& py, float& pz, float& m) {
    const int xc = min(max(x, 0), W - 1), yc = min(max(y, 0), H - 1);
    const float d = __ldg(ds + (size_t)yc * W + xc);
    px = ((float)xc - 0.5f * (float)W) * d / k00;
    py = ((float)yc - 0.5f * (float)H) * d / k11;
    pz = d;
    m = (__ldg(draw + (size_t)yc * W + xc) > 0.f) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(MO_THREADS)
spawn_candidates_kernel(AgsSpawnArgs a, SpawnCam cam, SpawnWs w) {
    const int P = a.H * a.W;
    const int pix = blockIdx.x * MO_THREADS + threadIdx.x;
    bool select = false;
    if (pix < P) {
        const int y = pix / a.W, x = pix - y * a.W;
        const float depth = __ldg(a.depth + pix);
        bool valid = depth > 0.f;
        // ---- world ray through the pixel centre and the back-projected point
        const float u = ((float)x + 0.5f) / (float)a.W, v = ((float)y + 0.5f) / (float)a.H;
        const float dcx = cam.Kinv[0] * u + cam.Kinv[1] * v + cam.Kinv[2];
        const float dcy = cam.Kinv[3] * u + cam.Kinv[4] * v + cam.Kinv[5];
        const float dcz = cam.Kinv[6] * u + cam.Kinv[7] * v + cam.Kinv[8];
        const float dwx = cam.c2w[0] * dcx + cam.c2w[1] * dcy + cam.c2w[2] * dcz;
        const float dwy = cam.c2w[4] * dcx + cam.c2w[5] * dcy + cam.c2w[6] * dcz;
        const float dwz = cam.c2w[8] * dcx + cam.c2w[9] * dcy + cam.c2w[10] * dcz;
        const float mx = cam.c2w[3] + dwx * depth, my = cam.c2w[7] + dwy * depth, mz = cam.c2w[11] + dwz * depth;
        // ---- camera-space normal of the smoothed depth, fov hard-wired to 60 x 60 degrees
        const float tan30 = 0.57735026918962576f;
        const float k00 = (float)a.H / (2.f * tan30), k11 = (float)a.W / (2.f * tan30);
        float cx, cy, cz, cm, ux, uy, uz, um, lx, ly, lz, lm, bx, by, bz, bm, rx, ry, rz, rm;
        d2n_pos(a.depth_smooth, a.depth, x, y, a.H, a.W, k00, k11, cx, cy, cz, cm);
        d2n_pos(a.depth_smooth, a.depth, x, y - 1, a.H, a.W, k00, k11, ux, uy, uz, um);
        d2n_pos(a.depth_smooth, a.depth, x - 1, y, a.H, a.W, k00, k11, lx, ly, lz, lm);
        d2n_pos(a.depth_smooth, a.depth, x, y + 1, a.H, a.W, k00, k11, bx, by, bz, bm);
        d2n_pos(a.depth_smooth, a.depth, x + 1, y, a.H, a.W, k00, k11, rx, ry, rz, rm);
        cx *= cm; cy *= cm; cz *= cm;
        ux = (ux - cx) * um; uy = (uy - cy) * um; uz = (uz - cz) * um;
        lx = (lx - cx) * lm; ly = (ly - cy) * lm; lz = (lz - cz) * lm;
        bx = (bx - cx) * bm; by = (by - cy) * bm; bz = (bz - cz) * bm;
        rx = (rx - cx) * rm; ry = (ry - cy) * rm; rz = (rz - cz) * rm;
        float nx = (uy * lz - uz * ly) + (ry * uz - rz * uy) + (by * rz - bz * ry) + (ly * bz - lz * by);
        float ny = (uz * lx - ux * lz) + (rz * ux - rx * uz) + (bz * rx - bx * rz) + (lz * bx - lx * bz);
        float nz = (ux * ly - uy * lx) + (rx * uy - ry * ux) + (bx * ry - by * rx) + (lx * by - ly * bx);
        const float nn = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);
        nx = nx / nn * cm; ny = ny / nn * cm; nz = nz / nn * cm;
        valid = valid && (nx * nx + ny * ny + nz * nz > 0.f);
        // ---- world normal ((0,0,1) where invalid), back-facing test against the unit ray
        float wx = 0.f, wy = 0.f, wz = 1.f;
        if (valid) {
            wx = cam.c2w[0] * nx + cam.c2w[1] * ny + cam.c2w[2] * nz;
            wy = cam.c2w[4] * nx + cam.c2w[5] * ny + cam.c2w[6] * nz;
            wz = cam.c2w[8] * nx + cam.c2w[9] * ny + cam.c2w[10] * nz;
        }
        const float dn = fmaxf(sqrtf(dwx * dwx + dwy * dwy + dwz * dwz), 1e-12f);
        const float cosv = (dwx / dn) * wx + (dwy / dn) * wy + (dwz / dn) * wz;
        valid = valid && (cosv < -0.01f);
        // ---- normal2rotation: frame (x, y, z = normal) -> quaternion (r, x, y, z)
        const float zn = sqrtf(wx * wx + wy * wy + wz * wz);
        const float zx = wx / zn, zy = wy / zn, zz = wz / zn;
        const bool par = fabsf(zx) > 0.99f;
        const float r0 = par ? 0.f : 1.f, r1 = par ? 1.f : 0.f;
        const float pr = r0 * zx + r1 * zy;
        float xx = r0 - pr * zx, xy = r1 - pr * zy, xz = -pr * zz;
        const float xn = sqrtf(xx * xx + xy * xy + xz * xz);
        xx /= xn; xy /= xn; xz /= xn;
        float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
        const float yn = sqrtf(yx * yx + yy * yy + yz * yz);
        yx /= yn; yy /= yn; yz /= yn;
        const float tr = xx + yy + zz + 1e-6f;
        const float qr = sqrtf(1.f + tr) / 2.f;
        float q0 = qr, q1 = (yz - zy) / (4.f * qr), q2 = (zx - xz) / (4.f * qr), q3 = (xy - yx) / (4.f * qr);
        const float qn = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
        q0 /= qn; q1 /= qn; q2 /= qn; q3 /= qn;
        valid = valid && !(isnan(q0) || isnan(q1) || isnan(q2) || isnan(q3));
        // ---- cal_mask against the render of the current map
        bool want = true;
        if (a.pred_rgb) {
            const float e0 = __ldg(a.rgb + pix) - __ldg(a.pred_rgb + pix);
            const float e1 = __ldg(a.rgb + P + pix) - __ldg(a.pred_rgb + P + pix);
            const float e2 = __ldg(a.rgb + 2 * P + pix) - __ldg(a.pred_rgb + 2 * P + pix);
            const float err = (e0 * e0 + e1 * e1 + e2 * e2) / 3.f;
            want = (err > a.error_thres) || (__ldg(a.pred_opacity + pix) < 0.5f) ||
                   ((depth - __ldg(a.pred_depth + pix)) < -0.05f * depth);
        }
        select = valid && want;
        int slot = -1;
        if (select) {
            w.cand_mean[3 * (size_t)pix + 0] = mx; w.cand_mean[3 * (size_t)pix + 1] = my; w.cand_mean[3 * (size_t)pix + 2] = mz;
            reinterpret_cast<float4*>(w.cand_rot)[pix] = make_float4(q0, q1, q2, q3);
            slot = 0;
            if (a.voxel_size > 0.f) {
                // one entry per occupied voxel; the member with the largest (random priority, pixel) wins
                const long long off = 1ll << 20, top = (1ll << 21) - 1;
                long long vx = (long long)floorf(mx / a.voxel_size) + off;
                long long vy = (long long)floorf(my / a.voxel_size) + off;
                long long vz = (long long)floorf(mz / a.voxel_size) + off;
                vx = min(max(vx, 0ll), top); vy = min(max(vy, 0ll), top); vz = min(max(vz, 0ll), top);
                const unsigned long long key = ((unsigned long long)vx << 42) | ((unsigned long long)vy << 21) | (unsigned long long)vz;
                const unsigned long long val = ((unsigned long long)mix32((uint32_t)pix * 0x9e3779b9u + a.seed) << 32) | (uint32_t)pix;
                uint32_t s = (uint32_t)mix64(key) & w.table_mask;
                while (true) {
                    const unsigned long long prev = atomicCAS(w.keys + s, ~0ull, key);
                    if (prev == ~0ull || prev == key) break;
                    s = (s + 1) & w.table_mask;
                }
                atomicMax(w.vals + s, val);
                slot = (int)s;
            }
        }
        w.slot[pix] = slot;
        if (a.select_out) a.select_out[pix] = select ? 1 : 0;
    }
    const int cnt = __syncthreads_count(select);
    if (threadIdx.x == 0 && cnt) atomicAdd(a.counters + 1, cnt);
}

__device__ __forceinline__ bool spawn_is_winner(const AgsSpawnArgs& a, const SpawnWs& w, int pix, int P) {
    if (pix >= P) return false;
    const int s = w.slot[pix];
    if (s < 0) return false;
    if (!(a.voxel_size > 0.f)) return true;
    return (uint32_t)(w.vals[s] & 0xffffffffull) == (uint32_t)pix;
}

__global__ void __launch_bounds__(MO_THREADS)
spawn_count_kernel(AgsSpawnArgs a, SpawnWs w) {
    const int P = a.H * a.W;
    const int pix = blockIdx.x * MO_THREADS + threadIdx.x;
    const int cnt = __syncthreads_count(spawn_is_winner(a, w, pix, P));
    if (threadIdx.x == 0) w.blk_count[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(MO_THREADS)
spawn_append_kernel(AgsSpawnArgs a, SpawnWs w) {
    const int P = a.H * a.W;
    const int pix = blockIdx.x * MO_THREADS + threadIdx.x;
    const bool win = spawn_is_winner(a, w, pix, P);
    int total;
    const int rank = block_scan_flag(win, total);
    if (!win) return;
    const long long row = (long long)a.n_old + w.blk_offset[blockIdx.x] + rank;
    if (row >= a.capacity) return;
    const size_t r = (size_t)row;
    a.means[3 * r + 0] = w.cand_mean[3 * (size_t)pix + 0];
    a.means[3 * r + 1] = w.cand_mean[3 * (size_t)pix + 1];
    a.means[3 * r + 2] = w.cand_mean[3 * (size_t)pix + 2];
    a.scales[3 * r + 0] = 0.f; a.scales[3 * r + 1] = 0.f; a.scales[3 * r + 2] = -1e10f;
    reinterpret_cast<float4*>(a.rotations)[r] = reinterpret_cast<const float4*>(w.cand_rot)[pix];
    a.opacities[r] = 0.f;
    a.harmonics[3 * r + 0] = __ldg(a.rgb + pix);
    a.harmonics[3 * r + 1] = __ldg(a.rgb + P + pix);
    a.harmonics[3 * r + 2] = __ldg(a.rgb + 2 * P + pix);
    a.view_scores[r] = 0.f;
    a.view_supports[r] = 0.f;
    a.view_means[3 * r + 0] = 0.f; a.view_means[3 * r + 1] = 0.f; a.view_means[3 * r + 2] = 0.f;
    if (a.select_out) a.select_out[pix] = 2;
}

// =================================================================================================
// confidence bookkeeping (ags_view_stats_update)
__global__ void __launch_bounds__(MO_THREADS)
view_stats_kernel(int N, const int32_t* __restrict__ count_last, const float* __restrict__ means,
                  const float* __restrict__ rot, float cx, float cy, float cz, float depth_max, int use_vd,
                  float* __restrict__ supports, float* __restrict__ vmeans, float* __restrict__ scores) {
    const int i = blockIdx.x * MO_THREADS + threadIdx.x;
    if (i >= N) return;
    if (!(__ldg(count_last + i) >= 1)) return;
    const float sup = supports[i] + 1.f;
    supports[i] = sup;
    if (!use_vd) return;
    float vx = cx - __ldg(means + 3 * (size_t)i), vy = cy - __ldg(means + 3 * (size_t)i + 1), vz = cz - __ldg(means + 3 * (size_t)i + 2);
    const float dist = sqrtf(vx * vx + vy * vy + vz * vz);
    vx /= dist; vy /= dist; vz /= dist;
    const float4 q = __ldg(reinterpret_cast<const float4*>(rot) + i);
    const float qn = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    const float r = q.x / qn, x = q.y / qn, y = q.z / qn, z = q.w / qn;
    float nx = 2.f * (x * z + r * y), ny = 2.f * (y * z - r * x), nz = 1.f - 2.f * (x * x + y * y);
    const float nn = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);
    nx /= nn; ny /= nn; nz /= nn;
    float* m = vmeans + 3 * (size_t)i;
    m[0] += (vx - m[0]) / sup;
    m[1] += (vy - m[1]) / sup;
    m[2] += (vz - m[2]) / sup;
    const float cosv = fminf(fmaxf(nx * vx + ny * vy + nz * vz, 0.f), 1.f);
    const float dfac = fminf(fmaxf(dist / depth_max, 0.f), 1.f);
    scores[i] += (1.f - dfac) * cosv;
}

// =================================================================================================
// prune (ags_prune_compact)
struct PruneWs {
    uint8_t* keep;          // (N)
    int32_t* blk_count;     // (nblk)
    int32_t* blk_offset;    // (nblk)
    size_t total;
};

inline PruneWs prune_carve(void* base, int N) {
    PruneWs w;
    const size_t nblk = ((size_t)N + MO_THREADS - 1) / MO_THREADS;
    size_t off = 0;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { char* q = p + off; off += ags_align256(bytes); return (void*)q; };
    w.keep = (uint8_t*)take((size_t)N + 1);
    w.blk_count = (int32_t*)take(nblk * 4 + 4);
    w.blk_offset = (int32_t*)take(nblk * 4 + 4);
    w.total = off;
    return w;
}

__global__ void __launch_bounds__(MO_THREADS)
prune_flag_kernel(AgsPruneArgs a, PruneWs w) {
    const int i = blockIdx.x * MO_THREADS + threadIdx.x;
    bool keep = false;
    if (i < a.N) {
        bool drop = a.prune_mask ? (a.prune_mask[i] != 0) : false;
        if (a.counts) {
            long long seen = 0;
            for (int t = 0; t < a.T; ++t) seen += __ldg(a.counts + (size_t)t * a.N + i);
            drop = drop || !(seen >= 1);
        }
        const float o = __ldg(a.src[3] + i);
        drop = drop || (1.f / (1.f + expf(-o)) < 0.1f);
        if (a.prune_mask) a.prune_mask[i] = drop ? 1 : 0;
        keep = !drop;
        w.keep[i] = keep ? 1 : 0;
    }
    const int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0) w.blk_count[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(MO_THREADS)
prune_scatter_kernel(AgsPruneArgs a, PruneWs w) {
    const int i = blockIdx.x * MO_THREADS + threadIdx.x;
    const bool keep = (i < a.N) && w.keep[i];
    int total;
    const int rank = block_scan_flag(keep, total);
    if (!keep) return;
    const size_t d = (size_t)w.blk_offset[blockIdx.x] + rank, s = (size_t)i;
    const int width[8] = {3, 3, 4, 1, 3, 1, 1, 3};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < width[k]) a.dst[k][d * width[k] + c] = __ldg(a.src[k] + s * width[k] + c);
    }
}

// =================================================================================================
// planner utilities (ags_view_utility)
constexpr int UT_THREADS = 512;

__device__ __forceinline__ float block_sum(float v, float* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_buf[wid] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < UT_THREADS / 32) ? s_buf[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;      // valid in thread 0
}

__global__ void __launch_bounds__(UT_THREADS)
view_utility_kernel(AgsUtilityArgs a) {
    __shared__ float s_buf[UT_THREADS / 32];
    const int v = blockIdx.x;
    const int npx = a.h * a.w;
    const float* depth = a.depth + (size_t)v * npx;
    const float* conf = a.confidence + (size_t)v * npx;
    const uint8_t* valid = a.valid ? a.valid + (size_t)v * npx : nullptr;
    // ---- exploitation: mean distance-weighted uncertainty
    float acc = 0.f;
    for (int p = threadIdx.x; p < npx; p += UT_THREADS) {
        const float d = __ldg(depth + p);
        float c = __ldg(conf + p);
        if (d > a.depth_hi) c = 1.f;
        if (valid && !valid[p]) c = 1.f;
        const float ds = (d < 0.001f) ? a.depth_hi * 0.5f : d;
        acc += (1.f - c) * ds / a.depth_hi;
    }
    const float tot = block_sum(acc, s_buf);
    if (threadIdx.x == 0) {
        const float e = tot / (float)npx;
        a.exploit[v] = isnan(e) ? 0.f : e;
    }
    // ---- exploration: unexplored voxels whose centre is visible in this view
    const float* Wm = a.w2c + (size_t)v * 16;
    const float* K = a.K + (size_t)v * 9;
    float Wr[12], Kr[9];
#pragma unroll
    for (int k = 0; k < 12; ++k) Wr[k] = __ldg(Wm + k);
#pragma unroll
    for (int k = 0; k < 9; ++k) Kr[k] = __ldg(K + k);
    float cnt = 0.f;
    for (int m = threadIdx.x; m < a.M; m += UT_THREADS) {
        if (!a.unexplored[m]) continue;
        const float x = __ldg(a.voxel_centers + 3 * (size_t)m), y = __ldg(a.voxel_centers + 3 * (size_t)m + 1),
                    z = __ldg(a.voxel_centers + 3 * (size_t)m + 2);
        const float cx = Wr[0] * x + Wr[1] * y + Wr[2] * z + Wr[3];
        const float cy = Wr[4] * x + Wr[5] * y + Wr[6] * z + Wr[7];
        const float cz = Wr[8] * x + Wr[9] * y + Wr[10] * z + Wr[11];
        const float ix = Kr[0] * cx + Kr[1] * cy + Kr[2] * cz;
        const float iy = Kr[3] * cx + Kr[4] * cy + Kr[5] * cz;
        const float iz = Kr[6] * cx + Kr[7] * cy + Kr[8] * cz;
        const float px = ix / iz * (float)a.w, py = iy / iz * (float)a.h;
        const bool inside = (px >= 0.f) && (px < (float)a.w) && (py >= 0.f) && (py < (float)a.h);
        if (!(cz > 0.f) || !inside) continue;
        const int q = (int)py * a.w + (int)px;
        float dv = __ldg(depth + q);
        if (dv < 0.001f) dv = 10000.f;
        dv = fminf(fmaxf(dv, a.depth_lo), a.depth_hi);
        if (valid && !valid[q]) dv = -1.f;
        if (dv > cz) cnt += 1.f;
    }
    const float ctot = block_sum(cnt, s_buf);
    if (threadIdx.x == 0) {
        const float e = ctot / (float)a.M;
        a.explore[v] = isnan(e) ? 0.f : e;
    }
}

// =================================================================================================
// low-confidence voxels (ags_voxel_roi)
__global__ void __launch_bounds__(MO_THREADS)
voxel_roi_scatter_kernel(AgsVoxelRoiArgs a) {
    const int i = blockIdx.x * MO_THREADS + threadIdx.x;
    if (i >= a.N) return;
    if (!(__ldg(a.confidences + i) < a.confidence_thres)) return;
    const float o = __ldg(a.opacities + i);
    if (!(1.f / (1.f + expf(-o)) > a.opacity_thres)) return;
    int v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        v[k] = (int)floorf((__ldg(a.means + 3 * (size_t)i + k) - a.bbox_min[k]) / a.voxel_size[k]);
        if (v[k] < 0 || v[k] >= a.dim[k]) return;
    }
    const size_t lin = ((size_t)v[0] * a.dim[1] + v[1]) * a.dim[2] + v[2];
    const float4 q = __ldg(reinterpret_cast<const float4*>(a.rotations) + i);
    const float qn = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    const float r = q.x / qn, x = q.y / qn, y = q.z / qn, z = q.w / qn;
    float nx = 2.f * (x * z + r * y), ny = 2.f * (y * z - r * x), nz = 1.f - 2.f * (x * x + y * y);
    const float nn = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);
    atomicAdd(a.voxel_count + lin, 1);
    atomicAdd(a.voxel_normal + 3 * lin + 0, nx / nn);
    atomicAdd(a.voxel_normal + 3 * lin + 1, ny / nn);
    atomicAdd(a.voxel_normal + 3 * lin + 2, nz / nn);
}

__global__ void __launch_bounds__(MO_THREADS)
voxel_roi_finish_kernel(AgsVoxelRoiArgs a, int M) {
    const int m = blockIdx.x * MO_THREADS + threadIdx.x;
    if (m >= M) return;
    const int c = a.voxel_count[m];
    const bool upd = c > a.min_gaussian_per_voxel;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (upd) {
        nx = a.voxel_normal[3 * (size_t)m] / (float)c;
        ny = a.voxel_normal[3 * (size_t)m + 1] / (float)c;
        nz = a.voxel_normal[3 * (size_t)m + 2] / (float)c;
        const float nn = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);
        nx /= nn; ny /= nn; nz /= nn;
    }
    a.voxel_normal[3 * (size_t)m] = nx; a.voxel_normal[3 * (size_t)m + 1] = ny; a.voxel_normal[3 * (size_t)m + 2] = nz;
    a.update_mask[m] = upd ? 1 : 0;
}

}  // namespace

// =================================================================================================
extern "C" int ags_voxel_roi(const AgsVoxelRoiArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->N >= 0, "negative N");
    AGS_CHECK_ARG(a->dim[0] > 0 && a->dim[1] > 0 && a->dim[2] > 0 &&
                  (long long)a->dim[0] * a->dim[1] * a->dim[2] < (1ll << 30), "bad voxel grid %d x %d x %d",
                  a->dim[0], a->dim[1], a->dim[2]);
    AGS_CHECK_ARG(a->voxel_size[0] > 0.f && a->voxel_size[1] > 0.f && a->voxel_size[2] > 0.f, "voxel size must be positive");
    AGS_CHECK_ARG(a->voxel_count && a->voxel_normal && a->update_mask, "NULL output");
    AGS_CHECK_ARG(a->N == 0 || (a->means && a->rotations && a->opacities && a->confidences), "NULL input");
    AGS_CHECK_ARG(((uintptr_t)a->rotations & 15) == 0, "rotations must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)a->stream;
    const int M = a->dim[0] * a->dim[1] * a->dim[2];
    AGS_CHECK_CUDA(cudaMemsetAsync(a->voxel_count, 0, (size_t)M * 4, st));
    AGS_CHECK_CUDA(cudaMemsetAsync(a->voxel_normal, 0, (size_t)M * 12, st));
    if (a->N > 0) { ags_note_launch(); voxel_roi_scatter_kernel<<<(a->N + MO_THREADS - 1) / MO_THREADS, MO_THREADS, 0, st>>>(*a); }
    ags_note_launch(); voxel_roi_finish_kernel<<<(M + MO_THREADS - 1) / MO_THREADS, MO_THREADS, 0, st>>>(*a, M);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t ags_spawn_scratch_bytes(int32_t H, int32_t W) {
    if (H <= 0 || W <= 0) return 0;
    return spawn_carve(nullptr, H, W).total;
}

extern "C" int ags_spawn(const AgsSpawnArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->H > 0 && a->W > 0 && (long long)a->H * a->W < (1ll << 30), "bad image size %d x %d", a->H, a->W);
    AGS_CHECK_ARG(a->rgb && a->depth && a->depth_smooth, "NULL keyframe image");
    AGS_CHECK_ARG((a->pred_rgb != nullptr) == (a->pred_depth != nullptr) && (a->pred_rgb != nullptr) == (a->pred_opacity != nullptr),
                  "pred_rgb / pred_depth / pred_opacity must be given together");
    AGS_CHECK_ARG(a->n_old >= 0 && a->capacity >= a->n_old, "bad n_old %d / capacity %d", a->n_old, a->capacity);
    AGS_CHECK_ARG(a->means && a->scales && a->rotations && a->opacities && a->harmonics && a->view_scores &&
                  a->view_supports && a->view_means, "NULL map buffer");
    AGS_CHECK_ARG(((uintptr_t)a->rotations & 15) == 0, "rotations must be 16-byte aligned");
    AGS_CHECK_ARG(a->counters != nullptr, "NULL counters");
    AGS_CHECK_ARG(a->workspace != nullptr && ((uintptr_t)a->workspace & 255) == 0, "workspace NULL or not 256-byte aligned");
    const size_t need = ags_spawn_scratch_bytes(a->H, a->W);
    AGS_CHECK_ARG(a->workspace_bytes >= need, "workspace too small: %zu < %zu", a->workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)a->stream;
    SpawnWs w = spawn_carve(a->workspace, a->H, a->W);
    SpawnCam cam;
    memcpy(cam.c2w, a->c2w, sizeof(cam.c2w));
    memcpy(cam.Kinv, a->Kinv, sizeof(cam.Kinv));
    const int P = a->H * a->W, nblk = (P + MO_THREADS - 1) / MO_THREADS;
    AGS_CHECK_CUDA(cudaMemsetAsync(w.keys, 0xff, (size_t)(w.table_mask + 1) * 8, st));
    AGS_CHECK_CUDA(cudaMemsetAsync(w.vals, 0, (size_t)(w.table_mask + 1) * 8, st));
    AGS_CHECK_CUDA(cudaMemsetAsync(a->counters, 0, 4 * sizeof(int32_t), st));
    ags_note_launch(); spawn_candidates_kernel<<<nblk, MO_THREADS, 0, st>>>(*a, cam, w);
    ags_note_launch(); spawn_count_kernel<<<nblk, MO_THREADS, 0, st>>>(*a, w);
    ags_note_launch(); scan_counts_kernel<<<1, 1024, 0, st>>>(w.blk_count, w.blk_offset, nblk, a->counters, a->counters + 2,
                                           a->capacity - a->n_old);
    ags_note_launch(); spawn_append_kernel<<<nblk, MO_THREADS, 0, st>>>(*a, w);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ags_view_stats_update(int32_t N, const int32_t* count_last, const float* means, const float* rotations_raw,
                                     float cam_x, float cam_y, float cam_z, float depth_max, int32_t use_view_distribution,
                                     float* view_supports, float* view_means, float* view_scores, void* stream) {
    AGS_CHECK_ARG(N >= 0, "negative N");
    if (N == 0) return 0;
    AGS_CHECK_ARG(count_last && means && rotations_raw && view_supports && view_means && view_scores, "NULL argument");
    AGS_CHECK_ARG(((uintptr_t)rotations_raw & 15) == 0, "rotations must be 16-byte aligned");
    ags_note_launch(); view_stats_kernel<<<(N + MO_THREADS - 1) / MO_THREADS, MO_THREADS, 0, (cudaStream_t)stream>>>(
        N, count_last, means, rotations_raw, cam_x, cam_y, cam_z, depth_max, use_view_distribution, view_supports,
        view_means, view_scores);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t ags_prune_scratch_bytes(int32_t N) {
    if (N < 0) return 0;
    return prune_carve(nullptr, N).total;
}

extern "C" int ags_prune_compact(const AgsPruneArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->N >= 0 && (a->counts == nullptr || a->T > 0), "bad N=%d / T=%d", a->N, a->T);
    AGS_CHECK_ARG(a->n_kept != nullptr, "NULL n_kept");
    cudaStream_t st = (cudaStream_t)a->stream;
    if (a->N == 0) {
        AGS_CHECK_CUDA(cudaMemsetAsync(a->n_kept, 0, sizeof(int32_t), st));
        return 0;
    }
    for (int k = 0; k < 8; ++k) AGS_CHECK_ARG(a->src[k] && a->dst[k] && a->src[k] != a->dst[k], "src/dst %d NULL or aliased", k);
    AGS_CHECK_ARG(a->workspace != nullptr && ((uintptr_t)a->workspace & 255) == 0, "workspace NULL or not 256-byte aligned");
    AGS_CHECK_ARG(a->workspace_bytes >= ags_prune_scratch_bytes(a->N), "workspace too small");
    PruneWs w = prune_carve(a->workspace, a->N);
    const int nblk = (a->N + MO_THREADS - 1) / MO_THREADS;
    ags_note_launch(); prune_flag_kernel<<<nblk, MO_THREADS, 0, st>>>(*a, w);
    ags_note_launch(); scan_counts_kernel<<<1, 1024, 0, st>>>(w.blk_count, w.blk_offset, nblk, a->n_kept, nullptr, -1);
    ags_note_launch(); prune_scatter_kernel<<<nblk, MO_THREADS, 0, st>>>(*a, w);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ags_view_utility(const AgsUtilityArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->V > 0 && a->h > 0 && a->w > 0 && a->M >= 0, "bad sizes V=%d h=%d w=%d M=%d", a->V, a->h, a->w, a->M);
    AGS_CHECK_ARG(a->depth && a->confidence && a->w2c && a->K && a->explore && a->exploit, "NULL argument");
    AGS_CHECK_ARG(a->M == 0 || (a->voxel_centers && a->unexplored), "NULL voxel arrays");
    AGS_CHECK_ARG(a->depth_hi > 0.f, "depth_hi must be positive");
    ags_note_launch(); view_utility_kernel<<<a->V, UT_THREADS, 0, (cudaStream_t)a->stream>>>(*a);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
