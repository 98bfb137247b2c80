// api.cu -- the extern "C" surface declared in include/ags_b200.h (argument checking, workspace
// carving, kernel sequencing).  No torch types, no allocation, no synchronisation.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "ags_common.cuh"

static thread_local char g_err[512] = "";

void ags_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool ags_use_tma() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("AGS_BWD_TMA");
        on = (e && atoi(e) != 0) ? 1 : 0;
    }
    return on == 1;
}

static unsigned long long g_launches = 0;   // host-side, one mapper thread per process (SURVEY 8b threading)
void ags_note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" unsigned long long ags_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" const char* ags_last_error(void) { return g_err; }
extern "C" int ags_version(void) { return 100; }

extern "C" size_t ags_scratch_bytes(int32_t N, int32_t B, int32_t H, int32_t W, int32_t inst_cap) {
    if (N < 0 || B <= 0 || H <= 0 || W <= 0 || inst_cap < 0) return 0;
    return ags_carve(nullptr, N, B, H, W, inst_cap).total;
}

// One launch zeroes everything the forward accumulates into: the per-tile tables + device counters (contiguous
// in the workspace), the statistics and -- when present -- the importance / count outputs.
__global__ void __launch_bounds__(256)
clear_kernel(int4* tiles, size_t tile_quads, int32_t* stats, int4* imp, int4* cnt, size_t bn_quads, int bn_tail,
             int32_t* imp_tail, int32_t* cnt_tail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int4 z = make_int4(0, 0, 0, 0);
    for (size_t i = t0; i < tile_quads; i += stride) tiles[i] = z;
    if (t0 < AGS_NUM_STATS) stats[t0] = 0;
    if (imp) for (size_t i = t0; i < bn_quads; i += stride) imp[i] = z;
    if (cnt) for (size_t i = t0; i < bn_quads; i += stride) cnt[i] = z;
    if ((int)t0 < bn_tail) {
        if (imp_tail) imp_tail[t0] = 0;
        if (cnt_tail) cnt_tail[t0] = 0;
    }
}

static int launch_clear(const AgsRenderArgs* a, const AgsWorkspace& w) {
    cudaStream_t st = (cudaStream_t)a->stream;
    const size_t zero_bytes = (char*)w.inst_key - (char*)w.tile_count;        // multiple of 256
    const size_t bn = a->N > 0 ? (size_t)a->B * a->N : 0;
    const bool aligned = (((uintptr_t)a->importance | (uintptr_t)a->count) & 15) == 0;
    const size_t quads = aligned ? bn / 4 : 0;
    const int tail = (int)(bn - quads * 4);
    if (!aligned && bn) {      // unaligned caller buffers (never from torch): plain memsets
        if (a->importance) AGS_CHECK_CUDA(cudaMemsetAsync(a->importance, 0, bn * 4, st));
        if (a->count) AGS_CHECK_CUDA(cudaMemsetAsync(a->count, 0, bn * 4, st));
    }
    size_t work = zero_bytes / 16;
    if ((a->importance || a->count) && quads > work) work = quads;
    long long blocks = (long long)((work + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    ags_note_launch();
    clear_kernel<<<(int)blocks, 256, 0, st>>>((int4*)w.tile_count, zero_bytes / 16, a->stats,
                                              aligned ? (int4*)a->importance : nullptr, aligned ? (int4*)a->count : nullptr, quads,
                                              aligned ? tail : 0,
                                              (aligned && a->importance) ? (int32_t*)a->importance + quads * 4 : nullptr,
                                              (aligned && a->count) ? a->count + quads * 4 : nullptr);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int check_render_args(const AgsRenderArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->N >= 0 && a->B > 0 && a->H > 0 && a->W > 0, "bad sizes N=%d B=%d H=%d W=%d", a->N, a->B, a->H, a->W);
    AGS_CHECK_ARG((a->W + TILE - 1) / TILE < 65536 && (a->H + TILE - 1) / TILE < 65536, "image too large");
    AGS_CHECK_ARG(a->param_mode == AGS_PARAMS_ACTIVATED || a->param_mode == AGS_PARAMS_RAW, "bad param_mode %d", a->param_mode);
    AGS_CHECK_ARG(a->inst_cap >= 0, "negative inst_cap");
    AGS_CHECK_ARG((long long)a->B * a->N < 2147483647LL, "B*N = %lld does not fit the 32-bit pair index",
                  (long long)a->B * a->N);
    if (a->N > 0)
        AGS_CHECK_ARG(a->means3D && a->scales && a->rotations && a->opacities && a->colors,
                      "NULL per-Gaussian input");
    AGS_CHECK_ARG(a->viewmatrix && a->projmatrix && a->tanfov && a->bg, "NULL per-view input");
    AGS_CHECK_ARG(a->out_rgb && a->out_normal && a->out_depth && a->out_opacity && a->out_confidence,
                  "NULL output image");
    AGS_CHECK_ARG(a->stats != nullptr, "NULL stats");
    if (a->N > 0) AGS_CHECK_ARG(a->radii != nullptr, "NULL radii");
    AGS_CHECK_ARG(!a->require_importance || a->N == 0 || (a->importance && a->count), "require_importance needs importance/count");
    AGS_CHECK_ARG(a->workspace != nullptr, "NULL workspace");
    const size_t need = ags_scratch_bytes(a->N, a->B, a->H, a->W, a->inst_cap);
    AGS_CHECK_ARG(a->workspace_bytes >= need, "workspace too small: %zu < %zu", a->workspace_bytes, need);
    AGS_CHECK_ARG(((uintptr_t)a->workspace & 255) == 0, "workspace must be 256-byte aligned");
    return 0;
}

static int render_forward_impl(const AgsRenderArgs* a, bool for_backward) {
    int rc = check_render_args(a);
    if (rc) return rc;
    AgsWorkspace w = ags_carve(a->workspace, a->N, a->B, a->H, a->W, a->inst_cap);
    // tile_count, tile_offset, tile_fill and the counters are contiguous in the workspace; importance /
    // count are all-zero unless config[3] (optional outputs)
    if ((rc = launch_clear(a, w))) return rc;
    if ((rc = ags_launch_project_fwd(*a, w, for_backward))) return rc;
    if ((rc = ags_launch_binning(*a, w))) return rc;
    if ((rc = ags_launch_composite_fwd(*a, w))) return rc;
    return 0;
}

extern "C" int ags_render_forward(const AgsRenderArgs* a) { return render_forward_impl(a, true); }

// Profiling hook: run ONE stage of the pipeline (bench.py times each stage with CUDA events).
extern "C" int ags_render_stage(const AgsRenderArgs* a, const AgsRenderGradArgs* g, int stage) {
    int rc = check_render_args(a);
    if (rc) return rc;
    AgsWorkspace w = ags_carve(a->workspace, a->N, a->B, a->H, a->W, a->inst_cap);
    switch (stage) {
        case AGS_STAGE_CLEAR: return launch_clear(a, w);
        case AGS_STAGE_PROJECT_FWD: return ags_launch_project_fwd(*a, w, true);
        case AGS_STAGE_BINNING: return ags_launch_binning(*a, w);
        case AGS_STAGE_COMPOSITE_FWD: return ags_launch_composite_fwd(*a, w);
        case AGS_STAGE_COMPOSITE_BWD:
            AGS_CHECK_ARG(g != nullptr, "grads is NULL");
            return ags_launch_composite_bwd(*a, *g, w);
        case AGS_STAGE_PROJECT_BWD:
            AGS_CHECK_ARG(g != nullptr, "grads is NULL");
            return ags_launch_project_bwd(*a, *g, w);
        default: ags_set_error("unknown stage %d", stage); return -1;
    }
}

extern "C" int ags_render_backward(const AgsRenderArgs* a, const AgsRenderGradArgs* g) {
    int rc = check_render_args(a);
    if (rc) return rc;
    AGS_CHECK_ARG(g != nullptr, "grads is NULL");
    if (a->N > 0)
        AGS_CHECK_ARG(g->d_means3D && g->d_scales && g->d_rotations && g->d_opacities && g->d_colors,
                      "NULL gradient output");
    AgsWorkspace w = ags_carve(a->workspace, a->N, a->B, a->H, a->W, a->inst_cap);
    if ((rc = ags_launch_composite_bwd(*a, *g, w))) return rc;
    if ((rc = ags_launch_project_bwd(*a, *g, w))) return rc;
    return 0;
}
