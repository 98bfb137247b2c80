/* ags_b200.h -- C ABI of libags_b200.so: the B200-native (sm_100a) rasterize-and-optimise hot path
 * of ActiveGS.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   ags_render_forward   the forward of `GaussianRasterizer(settings)(...)`, i.e. the native call at
 *                        utils/operations.py:701-713 (module diff_gaussian_rasterization_2d,
 *                        envs/requirements.txt:15) -- for B views at once (the Python loop
 *                        utils/operations.py:853-892 collapses into one call).
 *   ags_render_backward  the autograd backward of that same call (triggered by
 *                        mapping/gaussian_map.py:125 `total_loss.backward()`).
 *   ags_loss_forward_backward  the per-pixel post-processing + losses + their gradients:
 *                        utils/operations.py:714-718 (mask, normalise, depth2normal :172-219) and
 *                        mapping/gaussian_map.py:106-124 (mapping/utils.py:14-16,28-62,120-121).
 *   ags_adam_step        torch.optim.Adam(eps=1e-15) over the five parameter groups,
 *                        mapping/gaussian_map.py:259-292,126-127.
 *   ags_scratch_bytes    (new) size query for the caller-owned workspace.
 *   ags_last_error       (new) message of the last failing call on this thread.
 *
 * Conventions
 *   - plain C types only; every pointer below is a DEVICE pointer unless its name ends in _host.
 *   - all tensors are fp32 contiguous; images are planar (B, C, H, W); per-view per-Gaussian
 *     arrays are (B, N).
 *   - ownership: the caller (PyTorch side) allocates and owns every buffer, including the
 *     workspace; the library allocates no persistent device memory and keeps no global state
 *     (re-entrant per process: the GUI process of the reference loads its own copy).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - return value: 0 ok; < 0 invalid argument (message via ags_last_error()); > 0 a cudaError_t.
 *   - instance capacity: the number of (tile, Gaussian) instances is data dependent.  The caller
 *     passes `inst_cap`; if the batch needs more, the forward renders nothing, sets
 *     stats[AGS_STAT_OVERFLOW]=1 and stats[AGS_STAT_INSTANCES]=required, and the caller retries
 *     with a larger workspace.  No host synchronisation happens inside the library.
 */
#ifndef AGS_B200_H
#define AGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGS_TILE 16            /* tile edge (pixels); part of the rasterizer semantics */
#define AGS_STAT_INSTANCES 0   /* stats[0]: instances the batch needs */
#define AGS_STAT_OVERFLOW 1    /* stats[1]: 1 if instances > inst_cap (nothing rendered) */
#define AGS_STAT_VISIBLE 2     /* stats[2]: sum over views of visible Gaussians */
#define AGS_NUM_STATS 8

/* parameter interpretation */
#define AGS_PARAMS_ACTIVATED 0 /* boundary semantics: inputs are the activated tensors of get_attr() */
#define AGS_PARAMS_RAW 1       /* fused path: inputs are GaussianMap's raw tensors; sigmoid /
                                  clamp(scale_factor*exp) / normalize (gaussian_map.py:529-545) are
                                  applied inside the projection kernel and its backward */

typedef struct AgsRenderArgs {
    int32_t N, B, H, W;
    int32_t param_mode;            /* AGS_PARAMS_* */
    int32_t require_importance;    /* config[3] */
    int32_t front_only;            /* config[4] */
    int32_t inst_cap;              /* capacity of the instance arrays inside the workspace */
    float scale_modifier;          /* settings.scale_modifier */
    float weight_thres;            /* settings.weight_thres */
    float scale_factor;            /* RAW mode only: GaussianMap.scale_factor (0.01) */
    float scale_max;               /* RAW mode only: clamp max (0.05) */
    /* per-Gaussian inputs */
    const float* means3D;          /* (N,3) */
    const float* scales;           /* (N,3) */
    const float* rotations;        /* (N,4) r,x,y,z */
    const float* opacities;        /* (N,)  */
    const float* colors;           /* (N,3) colors_precomp */
    const float* confidences;      /* (N,)  */
    /* per-view inputs */
    const float* viewmatrix;       /* (B,4,4) = (w2c)^T, row-vector convention */
    const float* projmatrix;       /* (B,4,4) = viewmatrix @ P^T */
    const float* tanfov;           /* (B,2) tan(fov_x/2), tan(fov_y/2) */
    const float* bg;               /* (3) or (4): first three used */
    const float* render_mask;      /* (B,H,W) 0/1 or NULL (= all ones); gates importance/count */
    /* outputs */
    float* out_rgb;                /* (B,3,H,W) */
    float* out_normal;             /* (B,3,H,W) un-normalised, view space, facing the camera */
    float* out_depth;              /* (B,1,H,W) opacity-normalised per-pixel plane depth */
    float* out_opacity;            /* (B,1,H,W) */
    float* out_confidence;         /* (B,1,H,W) */
    float* importance;             /* (B,N) f32, zero unless require_importance; may be NULL if not required */
    int32_t* count;                /* (B,N) i32, zero unless require_importance; may be NULL if not required */
    int32_t* radii;                /* (B,N) i32, 0 = culled */
    int32_t* stats;                /* (AGS_NUM_STATS) i32 */
    /* workspace (also carries everything the backward needs) */
    void* workspace;
    size_t workspace_bytes;
    void* stream;
} AgsRenderArgs;

typedef struct AgsRenderGradArgs {
    /* upstream gradients, any may be NULL (= zero) */
    const float* d_rgb;            /* (B,3,H,W) */
    const float* d_normal;         /* (B,3,H,W) */
    const float* d_depth;          /* (B,1,H,W) */
    const float* d_opacity;        /* (B,1,H,W) */
    const float* d_confidence;     /* (B,1,H,W) */
    /* outputs: summed over the B views; w.r.t. the tensors named by param_mode */
    float* d_means3D;              /* (N,3) */
    float* d_scales;               /* (N,3) */
    float* d_rotations;            /* (N,4) */
    float* d_opacities;            /* (N,)  */
    float* d_colors;               /* (N,3) */
    float* d_means2D;              /* (B,N,3) pixel-space mean gradient, or NULL */
    int32_t accumulate;            /* 0: overwrite the d_* outputs, 1: add to them */
    int32_t clear_records;         /* 1: leave the workspace ready for another backward on the same forward
                                      (autograd may call it twice); 0: one backward per forward (training loop) */
} AgsRenderGradArgs;

size_t ags_scratch_bytes(int32_t N, int32_t B, int32_t H, int32_t W, int32_t inst_cap);
int ags_render_forward(const AgsRenderArgs* args);
/* must be called with the same AgsRenderArgs (same workspace contents) as the forward */
int ags_render_backward(const AgsRenderArgs* args, const AgsRenderGradArgs* grads);

/* Profiling hook (bench.py): run one stage of the forward/backward pipeline so that each kernel can
 * be bracketed with CUDA events.  Stages must be issued in pipeline order. */
#define AGS_STAGE_CLEAR 0
#define AGS_STAGE_PROJECT_FWD 1
#define AGS_STAGE_BINNING 2
#define AGS_STAGE_COMPOSITE_FWD 3
#define AGS_STAGE_COMPOSITE_BWD 4
#define AGS_STAGE_PROJECT_BWD 5
int ags_render_stage(const AgsRenderArgs* args, const AgsRenderGradArgs* grads, int stage);

typedef struct AgsLossArgs {
    int32_t B, H, W;
    int32_t B_total;               /* frames in the whole (possibly multi-GPU) batch; loss means use it */
    /* rasterizer outputs (B,C,H,W) */
    const float* rgb; const float* normal; const float* depth; const float* opacity;
    /* ground truth */
    const float* rgb_gt;           /* (B,3,H,W) */
    const float* depth_gt;         /* (B,1,H,W) */
    const float* tanfov;           /* (B,2) tan(fov_x/2), tan(fov_y/2): the same tensor the rasterizer takes.
                                      depth2normal pairs fov_x with H and fov_y with W (quirk Q2 kept) */
    const int32_t* vis_count;      /* (H,W) sum over ALL frames of (opacity>1e-3) (quirk Q1), or NULL:
                                      computed from this call's B frames */
    /* outputs */
    float* normal_unit;            /* (B,3,H,W) normalize(normal)*mask   (operations.py:714-715) */
    float* d2n;                    /* (B,3,H,W) depth2normal             (operations.py:718)     */
    float* d_rgb; float* d_normal; float* d_depth;   /* dL/d(rasterizer outputs) */
    float* loss_terms;             /* (4 + 2*B): rgb, depth, cons, tv sums (already normalised),
                                      then per frame rgb-L1 mean, depth-L1 mean (track_performance) */
    float w_depth, w_cons, w_tv;   /* 0.8, 0.1, 0.1 (gaussian_map.py:119-124) */
    void* workspace; size_t workspace_bytes;   /* >= ags_loss_scratch_bytes */
    void* stream;
} AgsLossArgs;

size_t ags_loss_scratch_bytes(int32_t B, int32_t H, int32_t W);
int ags_loss_forward_backward(const AgsLossArgs* args);

/* forward-only post-processing of B rendered views (utils/operations.py:714-718):
 * normal_unit = normalize(normal)*(opacity>1e-2), d2n = depth2normal(depth, mask, fov). */
int ags_postprocess(int32_t B, int32_t H, int32_t W, const float* normal, const float* depth,
                    const float* opacity, const float* tanfov, float* normal_unit, float* d2n,
                    void* stream);

/* Spawn step (mapping/gaussian_map.py:294-322): get_smooth_depth (utils/operations.py:161-169) on the
 * device -- OpenCV's float32 bilateral filter (d, sigmaColor, sigmaSpace; BORDER_REFLECT_101, circular
 * support, 4096-bin exp LUT semantics) over the depth image with invalid (< 0) pixels zero-filled on
 * input and set to -1 on output.  depth/out: (H,W) device; scratch: >= 16 bytes device. */
int ags_smooth_depth(int32_t H, int32_t W, const float* depth, float* out, int32_t d, float sigma_color,
                     float sigma_space, void* scratch, void* stream);

#define AGS_ADAM_GROUPS 5
typedef struct AgsAdamArgs {
    int32_t num_groups;                    /* <= AGS_ADAM_GROUPS */
    float* param[AGS_ADAM_GROUPS];
    const float* grad[AGS_ADAM_GROUPS];
    float* exp_avg[AGS_ADAM_GROUPS];
    float* exp_avg_sq[AGS_ADAM_GROUPS];
    int64_t numel[AGS_ADAM_GROUPS];
    float lr[AGS_ADAM_GROUPS];
    float beta1, beta2, eps;
    int32_t step;                          /* 1-based step used when step_dev == NULL */
    int32_t* step_dev;                     /* optional device counter: the kernel uses *step_dev + 1
                                              and a trailing 1-thread kernel increments it (graph replay) */
    const int32_t* skip_flag;              /* optional: if *skip_flag != 0 the step is a no-op (pass
                                              stats + AGS_STAT_OVERFLOW so an overflowed forward never
                                              reaches the parameters) */
    void* stream;
} AgsAdamArgs;

int ags_adam_step(const AgsAdamArgs* args);

/* K9: fused reduce-scatter(gradients) -> Adam(on the owned shard) -> all-gather(parameters) over
 * NVLink peer memory (new: the reference is single-GPU).  All ranks launch it between two cross-GPU
 * barriers on the same stream.  grad / param / skip buffers are symmetric allocations; pass the peer
 * pointers of all ranks (index = rank), optionally the NVLS multicast addresses (NULL = peer loop).
 * The flat vector holds the AGS_ADAM_GROUPS parameter tensors back to back (numel[], lr[]), padded to
 * numel_padded (multiple of 4*world); rank r updates [r*C, (r+1)*C), C = numel_padded/world. */
#define AGS_MAX_PEERS 8
typedef struct AgsDistAdamArgs {
    int32_t world, rank;
    int32_t num_groups;
    int32_t step;                               /* 1-based */
    const float* grad_peers[AGS_MAX_PEERS];
    float* param_peers[AGS_MAX_PEERS];
    const int32_t* skip_peers[AGS_MAX_PEERS];   /* per-rank overflow flags (may be NULL) */
    const float* grad_multicast;                /* NVLS multicast address of the gradient buffer or NULL */
    float* param_multicast;                     /* NVLS multicast address of the parameter buffer or NULL */
    float* exp_avg;                             /* local, numel_padded */
    float* exp_avg_sq;                          /* local, numel_padded */
    int64_t numel[AGS_ADAM_GROUPS];
    int64_t numel_padded;
    float lr[AGS_ADAM_GROUPS];
    float beta1, beta2, eps;
    void* stream;
} AgsDistAdamArgs;
int ags_dist_adam_step(const AgsDistAdamArgs* args);

/* Per-iteration camera staging (new; replaces the per-view host work of GaussianRenderer.__init__,
 * utils/operations.py:748-762, inside the training loop): gathers the camera blocks of the sampled
 * keyframes from a device-resident table into the (B,16)/(B,16)/(B,2) tensors AgsRenderArgs points
 * at.  The keyframe ids travel BY VALUE as kernel arguments (ids_host is read before the call
 * returns), so no host-to-device copy is enqueued: a copy on the compute stream would share the
 * H2D copy engine with the keyframe uploads and stall the forward behind them.
 * table: (T, AGS_CAM_ROW) floats per keyframe = viewmatrix 16 | projmatrix 16 | tanfov 2. */
#define AGS_CAM_ROW 34
#define AGS_MAX_BATCH 64
int ags_stage_cameras(const float* table, int32_t T, const int32_t* ids_host, int32_t B,
                      float* viewmatrix, float* projmatrix, float* tanfov, void* stream);

const char* ags_last_error(void);
int ags_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGS_B200_H */
