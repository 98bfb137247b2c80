// gpu_naive.cu -- BENCH-ONLY stand-in for "the reference CUDA rasterizer" (BASELINE.md section 2).
//
// The reference's native extension (diff_gaussian_rasterization_2d, envs/requirements.txt:15) is an
// un-vendored dependency, so `bench.py --impl gpu_naive` times a straightforward GPU build of the SAME
// specification with the structure of the public 3DGS lineage the reference forks:
//     count tiles per splat -> device-wide inclusive scan -> host read of the instance total ->
//     duplicateWithKeys (tile << 32 | depth) -> ONE global 64-bit cub radix sort -> identifyTileRanges ->
//     render: one 16x16 CTA per tile, 256-splat shared-memory batches, EVERY pixel evaluates EVERY splat of
//     its tile (no per-warp culling) -> backward: the same walk with PER-PIXEL atomics into the per-splat
//     gradient record (no cross-lane reduction).
// The per-Gaussian projection (K1) and its backward (K6) are the product's own kernels (called by the host
// through ags_render_stage): they are 11 % of the product's step and their lineage form has the same shape.
// Nothing in active_gs_b200/ or diff_gaussian_rasterization_2d/ imports or links this file.
#include <stdarg.h>
#include <stdlib.h>
#include <cub/cub.cuh>
#include "../active_gs_b200/csrc/ags_common.cuh"

// ags_carve() (shared workspace layout) needs these two symbols of the product library
static thread_local char n_err[512] = "";
void ags_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(n_err, sizeof(n_err), fmt, ap);
    va_end(ap);
}
bool ags_use_tma() {
    const char* e = getenv("AGS_BWD_TMA");
    return e && atoi(e) != 0;
}
static unsigned long long n_launches = 0;
void ags_note_launch() { ++n_launches; }
extern "C" unsigned long long naive_launch_count(void) { return n_launches; }
extern "C" const char* naive_last_error(void) { return n_err; }

namespace {

struct NaiveWs {
    int32_t* touched;    // (B*N) tiles touched per (view, Gaussian)
    int32_t* offsets;    // (B*N) inclusive scan
    uint64_t* keys_in;   // (cap)
    uint64_t* keys_out;  // (cap)
    uint32_t* vals_in;   // (cap) pair index v*N+i
    uint32_t* vals_out;  // (cap)
    uint2* ranges;       // (B*tiles)
    void* cub_temp;
    size_t cub_bytes;
    size_t total;
};

size_t cub_temp_bytes(int BN, int cap) {
    size_t a = 0, b = 0;
    cub::DeviceScan::InclusiveSum(nullptr, a, (int32_t*)nullptr, (int32_t*)nullptr, BN);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, cap, 0, 64);
    return a > b ? a : b;
}

NaiveWs carve(void* base, int N, int B, int H, int W, int cap) {
    NaiveWs w;
    size_t off = 0;
    char* p = (char*)base;
    const size_t BN = (size_t)B * N;
    const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    auto take = [&](size_t bytes) { char* q = p + off; off += ags_align256(bytes); return (void*)q; };
    w.touched = (int32_t*)take(BN * 4);
    w.offsets = (int32_t*)take(BN * 4);
    w.keys_in = (uint64_t*)take((size_t)cap * 8);
    w.keys_out = (uint64_t*)take((size_t)cap * 8);
    w.vals_in = (uint32_t*)take((size_t)cap * 4);
    w.vals_out = (uint32_t*)take((size_t)cap * 4);
    w.ranges = (uint2*)take((size_t)B * tiles * 8);
    w.cub_bytes = cub_temp_bytes((int)BN, cap);
    w.cub_temp = take(w.cub_bytes);
    w.total = off;
    return w;
}

__global__ void count_kernel(AgsRenderArgs a, AgsWorkspace w, NaiveWs n, int BN) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BN) return;
    int c = 0;
    if (a.radii[idx] > 0) {
        const uint2 r = w.rect[idx];
        c = ((int)(r.x >> 16) - (int)(r.x & 0xffff)) * ((int)(r.y >> 16) - (int)(r.y & 0xffff));
    }
    n.touched[idx] = c;
}

__global__ void duplicate_kernel(AgsRenderArgs a, AgsWorkspace w, NaiveWs n, int BN, int tiles_x, int tiles) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BN || a.radii[idx] <= 0) return;
    int off = idx == 0 ? 0 : n.offsets[idx - 1];
    const int v = idx / a.N;
    const uint2 r = w.rect[idx];
    const uint32_t depth_bits = __float_as_uint(w.feat0[idx].w);
    for (int ty = r.y & 0xffff; ty < (int)(r.y >> 16); ++ty)
        for (int tx = r.x & 0xffff; tx < (int)(r.x >> 16); ++tx) {
            const uint64_t tile = (uint64_t)v * tiles + (uint64_t)ty * tiles_x + tx;
            n.keys_in[off] = (tile << 32) | depth_bits;
            n.vals_in[off] = (uint32_t)idx;
            ++off;
        }
}

__global__ void ranges_kernel(NaiveWs n, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t t = (uint32_t)(n.keys_out[i] >> 32);
    if (i == 0) n.ranges[t].x = 0;
    else {
        const uint32_t p = (uint32_t)(n.keys_out[i - 1] >> 32);
        if (p != t) { n.ranges[p].y = i; n.ranges[t].x = i; }
    }
    if (i == total - 1) n.ranges[t].y = total;
}

#define N_LOG2E 1.4426950408889634f

// lineage renderCUDA: CTA = tile, every thread = one pixel, all splats of the tile in 256-batches
__global__ void __launch_bounds__(256)
render_fwd_kernel(AgsRenderArgs a, AgsWorkspace w, NaiveWs n) {
    __shared__ float4 s_g0[256], s_g1[256], s_f0[256], s_f1[256];
    const int v = blockIdx.z;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int tiles = gridDim.x * gridDim.y;
    const int tid = threadIdx.y * 16 + threadIdx.x;
    const int px = blockIdx.x * TILE + threadIdx.x, py = blockIdx.y * TILE + threadIdx.y;
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = n.ranges[(size_t)v * tiles + tile];
    int todo = (int)range.y - (int)range.x;
    const int rounds = (todo + 255) / 256;
    bool done = !inside;
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f, D = 0.f, Cf = 0.f;
    int contributor = 0, last = 0;
    for (int r = 0; r < rounds; ++r, todo -= 256) {
        if (__syncthreads_count(done) == 256) break;
        const int j = r * 256 + tid;
        if (range.x + j < range.y) {
            const size_t idx = n.vals_out[range.x + j];
            s_g0[tid] = w.geom0[idx]; s_g1[tid] = w.geom1[idx]; s_f0[tid] = w.feat0[idx]; s_f1[tid] = w.feat1[idx];
        }
        __syncthreads();
        for (int k = 0; !done && k < min(256, todo); ++k) {
            ++contributor;
            const float4 g0 = s_g0[k], g1 = s_g1[k];
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = (-0.5f * (g0.z * N_LOG2E * dx * dx + g1.x * N_LOG2E * dy * dy) - g0.w * N_LOG2E * dx * dy);
            if (power > 0.f) continue;
            const float alpha = fminf(AGS_ALPHA_MAX, g1.y * exp2f(power));
            if (alpha < AGS_ALPHA_MIN) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < AGS_T_EPS) { done = true; continue; }
            const float wgt = alpha * T;
            const float4 f0 = s_f0[k], f1 = s_f1[k];
            C0 += wgt * f0.x; C1 += wgt * f0.y; C2 += wgt * f0.z;
            D += wgt * (f0.w - g1.z * dx - g1.w * dy);
            N0 += wgt * f1.x; N1 += wgt * f1.y; N2 += wgt * f1.z;
            Cf += wgt * f1.w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t P = (size_t)a.H * a.W, pix = (size_t)py * a.W + px;
        const float A = 1.f - T;
        float* o = a.out_rgb + (size_t)v * 3 * P + pix;
        o[0] = C0 + T * a.bg[0]; o[P] = C1 + T * a.bg[1]; o[2 * P] = C2 + T * a.bg[2];
        o = a.out_normal + (size_t)v * 3 * P + pix;
        o[0] = N0; o[P] = N1; o[2 * P] = N2;
        a.out_depth[(size_t)v * P + pix] = A > 0.f ? D / A : 0.f;
        a.out_opacity[(size_t)v * P + pix] = A;
        a.out_confidence[(size_t)v * P + pix] = Cf;
        w.final_T[(size_t)v * P + pix] = T;
        w.n_contrib[(size_t)v * P + pix] = last;
    }
}

// lineage backward: same walk, gradient of every (pixel, splat) pair goes to the splat's record with atomics
__global__ void __launch_bounds__(256)
render_bwd_kernel(AgsRenderArgs a, AgsRenderGradArgs gr, AgsWorkspace w, NaiveWs n) {
    __shared__ float4 s_g0[256], s_g1[256], s_f0[256], s_f1[256];
    __shared__ uint32_t s_idx[256];
    const int v = blockIdx.z;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int tiles = gridDim.x * gridDim.y;
    const int tid = threadIdx.y * 16 + threadIdx.x;
    const int px = blockIdx.x * TILE + threadIdx.x, py = blockIdx.y * TILE + threadIdx.y;
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = n.ranges[(size_t)v * tiles + tile];
    int todo = (int)range.y - (int)range.x;
    const int rounds = (todo + 255) / 256;
    const size_t P = (size_t)a.H * a.W, pix = (size_t)py * a.W + px, vp = (size_t)v * P + pix;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gN0 = 0.f, gN1 = 0.f, gN2 = 0.f, gD = 0.f, gCf = 0.f, rem = 0.f, T = 1.f;
    int last = 0;
    if (inside) {
        last = w.n_contrib[vp];
        const float Tf = w.final_T[vp], A = 1.f - Tf;
        if (gr.d_rgb) { const float* p = gr.d_rgb + (size_t)v * 3 * P + pix; gC0 = p[0]; gC1 = p[P]; gC2 = p[2 * P]; }
        if (gr.d_normal) { const float* p = gr.d_normal + (size_t)v * 3 * P + pix; gN0 = p[0]; gN1 = p[P]; gN2 = p[2 * P]; }
        const float gdep = gr.d_depth ? gr.d_depth[vp] : 0.f;
        float gA = gr.d_opacity ? gr.d_opacity[vp] : 0.f;
        if (gr.d_confidence) gCf = gr.d_confidence[vp];
        const float depth_out = a.out_depth[vp];
        if (A > 0.f) { gD = gdep / A; gA -= gdep * depth_out / A; }
        const float* c = a.out_rgb + (size_t)v * 3 * P + pix;
        const float* nn = a.out_normal + (size_t)v * 3 * P + pix;
        const float bgdot = gC0 * a.bg[0] + gC1 * a.bg[1] + gC2 * a.bg[2];
        const float S_all = gC0 * (c[0] - Tf * a.bg[0]) + gC1 * (c[P] - Tf * a.bg[1]) + gC2 * (c[2 * P] - Tf * a.bg[2])
                          + gN0 * nn[0] + gN1 * nn[P] + gN2 * nn[2 * P] + gD * (depth_out * A) + gCf * a.out_confidence[vp];
        rem = S_all + Tf * (bgdot - gA);
    }
    int contributor = 0;
    for (int r = 0; r < rounds; ++r, todo -= 256) {
        __syncthreads();
        const int j = r * 256 + tid;
        if (range.x + j < range.y) {
            const uint32_t idx = n.vals_out[range.x + j];
            s_idx[tid] = idx;
            s_g0[tid] = w.geom0[idx]; s_g1[tid] = w.geom1[idx]; s_f0[tid] = w.feat0[idx]; s_f1[tid] = w.feat1[idx];
        }
        __syncthreads();
        for (int k = 0; k < min(256, todo); ++k) {
            ++contributor;
            if (contributor > last) break;
            const float4 g0 = s_g0[k], g1 = s_g1[k];
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = (-0.5f * (g0.z * N_LOG2E * dx * dx + g1.x * N_LOG2E * dy * dy) - g0.w * N_LOG2E * dx * dy);
            if (power > 0.f) continue;
            const float G = exp2f(power);
            const float alpha = fminf(AGS_ALPHA_MAX, g1.y * G);
            if (alpha < AGS_ALPHA_MIN) continue;
            const float4 f0 = s_f0[k], f1 = s_f1[k];
            const float wgt = alpha * T;
            const float dpix = f0.w - g1.z * dx - g1.w * dy;
            const float sdot = gC0 * f0.x + gC1 * f0.y + gC2 * f0.z + gN0 * f1.x + gN1 * f1.y + gN2 * f1.z + gD * dpix + gCf * f1.w;
            rem -= wgt * sdot;
            const float dalpha = T * sdot - rem / (1.f - alpha);
            T *= (1.f - alpha);
            const float dpower = (g1.y * G <= AGS_ALPHA_MAX) ? alpha * dalpha : 0.f;
            const float wgD = wgt * gD;
            float* rec = w.dsplat + (size_t)s_idx[k] * 16;
            atomicAdd(rec + AGS_REC_PDX, dpower * dx); atomicAdd(rec + AGS_REC_PDY, dpower * dy);
            atomicAdd(rec + AGS_REC_PXX, dpower * dx * dx); atomicAdd(rec + AGS_REC_PXY, dpower * dx * dy);
            atomicAdd(rec + AGS_REC_PYY, dpower * dy * dy); atomicAdd(rec + AGS_REC_P1, dpower);
            atomicAdd(rec + AGS_REC_C0, wgt * gC0); atomicAdd(rec + AGS_REC_C0 + 1, wgt * gC1); atomicAdd(rec + AGS_REC_C0 + 2, wgt * gC2);
            atomicAdd(rec + AGS_REC_N0, wgt * gN0); atomicAdd(rec + AGS_REC_N0 + 1, wgt * gN1); atomicAdd(rec + AGS_REC_N0 + 2, wgt * gN2);
            atomicAdd(rec + AGS_REC_WD, wgD); atomicAdd(rec + AGS_REC_WDX, wgD * dx); atomicAdd(rec + AGS_REC_WDY, wgD * dy);
        }
    }
}

}  // namespace

extern "C" size_t naive_scratch_bytes(int32_t N, int32_t B, int32_t H, int32_t W, int32_t cap) {
    return carve(nullptr, N, B, H, W, cap).total;
}

// Binning + global sort + render of views whose splats ags_render_stage(CLEAR, PROJECT_FWD) has projected into
// a->workspace.  Synchronises the stream once to read the instance total (as the lineage does).  Returns the
// number of instances (> cap: nothing rendered, caller re-allocates), < 0 on error.
extern "C" long long naive_forward(const AgsRenderArgs* a, void* nws, size_t nws_bytes, int32_t cap) {
    AgsWorkspace w = ags_carve(a->workspace, a->N, a->B, a->H, a->W, a->inst_cap);
    NaiveWs n = carve(nws, a->N, a->B, a->H, a->W, cap);
    if (n.total > nws_bytes) { ags_set_error("naive workspace too small"); return -1; }
    cudaStream_t st = (cudaStream_t)a->stream;
    const int BN = a->B * a->N;
    const int tiles_x = (a->W + TILE - 1) / TILE, tiles_y = (a->H + TILE - 1) / TILE, tiles = tiles_x * tiles_y;
    ags_note_launch(); count_kernel<<<(BN + 255) / 256, 256, 0, st>>>(*a, w, n, BN);
    size_t tb = n.cub_bytes;
    ags_note_launch(); cub::DeviceScan::InclusiveSum(n.cub_temp, tb, n.touched, n.offsets, BN, st);
    int32_t total = 0;
    cudaMemcpyAsync(&total, n.offsets + BN - 1, 4, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { ags_set_error("naive_forward: sync failed"); return -1; }
    if (total > cap) return total;
    cudaMemsetAsync(n.ranges, 0, (size_t)a->B * tiles * 8, st);
    if (total > 0) {
        ags_note_launch(); duplicate_kernel<<<(BN + 255) / 256, 256, 0, st>>>(*a, w, n, BN, tiles_x, tiles);
        int bits = 1;
        while ((1ll << bits) < (long long)a->B * tiles) ++bits;
        tb = n.cub_bytes;
        ags_note_launch();
        cub::DeviceRadixSort::SortPairs(n.cub_temp, tb, n.keys_in, n.keys_out, n.vals_in, n.vals_out, total, 0, 32 + bits, st);
        ags_note_launch(); ranges_kernel<<<(total + 255) / 256, 256, 0, st>>>(n, total);
    }
    dim3 grid(tiles_x, tiles_y, a->B), block(16, 16);
    ags_note_launch(); render_fwd_kernel<<<grid, block, 0, st>>>(*a, w, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ags_set_error("naive_forward: %s", cudaGetErrorString(e)); return -1; }
    return total;
}

extern "C" int naive_backward(const AgsRenderArgs* a, const AgsRenderGradArgs* g, void* nws, int32_t cap) {
    AgsWorkspace w = ags_carve(a->workspace, a->N, a->B, a->H, a->W, a->inst_cap);
    NaiveWs n = carve(nws, a->N, a->B, a->H, a->W, cap);
    const int tiles_x = (a->W + TILE - 1) / TILE, tiles_y = (a->H + TILE - 1) / TILE;
    dim3 grid(tiles_x, tiles_y, a->B), block(16, 16);
    ags_note_launch(); render_bwd_kernel<<<grid, block, 0, (cudaStream_t)a->stream>>>(*a, *g, w, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ags_set_error("naive_backward: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
