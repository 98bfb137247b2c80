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
#define AGS_STAT_VIEW0 8       /* stats[8 + v]: instances of view v (v < 64): the per-frame cost the
                                  frame-sharded loop balances the ranks with */
#define AGS_NUM_STATS 72

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
    float* normal_unit;            /* (B,3,H,W) normalize(normal)*mask   (operations.py:714-715); NULL = not wanted */
    float* d2n;                    /* (B,3,H,W) depth2normal             (operations.py:718);     NULL = not wanted */
    float* d_rgb; float* d_normal; float* d_depth;   /* dL/d(rasterizer outputs) */
    float* loss_terms;             /* (4 + 2*B): rgb, depth, cons, tv sums (already normalised),
                                      then per frame rgb-L1 mean, depth-L1 mean (track_performance) */
    float w_depth, w_cons, w_tv;   /* 0.8, 0.1, 0.1 (gaussian_map.py:119-124) */
    void* workspace; size_t workspace_bytes;   /* >= ags_loss_scratch_bytes */
    void* stream;
    /* optional (NULL = absent) */
    const float* frame_weight;     /* (B) 1 = real frame, 0 = padding: a padded frame contributes nothing to the
                                      loss, to the visibility sum of quirk Q1 or to any gradient (frame sharding
                                      pads the keyframe batch to a multiple of the world size) */
    const float* const* rgb_gt_frames_host;    /* HOST array of B device pointers, each (3,H,W): ground truth read in */
    const float* const* depth_gt_frames_host;  /* place from the keyframe store instead of a stacked (B,..) copy    */
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
    int32_t zero_grad;                     /* 1: the gradients are set to zero after they are consumed
                                              (optimizer.zero_grad(), gaussian_map.py:127, in the same pass; grad[]
                                              is then written).  Pair with AgsRenderGradArgs.accumulate = 1 */
} AgsAdamArgs;

int ags_adam_step(const AgsAdamArgs* args);

/* K9: fused reduce-scatter(gradients) -> Adam(on the owned shard) -> all-gather(parameters) over
 * NVLink peer memory (new: the reference is single-GPU).  All ranks launch it between two cross-GPU
 * barriers on the same stream.  grad / param / skip buffers are symmetric allocations; pass the peer
 * pointers of all ranks (index = rank), optionally the NVLS multicast addresses (NULL = peer loop).
 * The flat vector holds the AGS_ADAM_GROUPS parameter tensors back to back (numel[], lr[]), padded to
 * numel_padded (multiple of 4*world); rank r updates [r*C, (r+1)*C), C = numel_padded/world. */
#define AGS_MAX_PEERS 8

/* Cross-GPU ordering folded into the exchange kernels (instead of a separate barrier launch between them):
 * every rank owns a symmetric int32 array `flags` of AGS_SYNC_WORDS words.  A producer kernel, once ALL its
 * blocks are done, stores `epoch` into word [phase*AGS_MAX_PEERS + rank] of every peer's array (release,
 * system scope); a consumer kernel spins (acquire) on its OWN array until all `world` words of the phase
 * have reached `epoch`.  `epoch` must grow by one per iteration.  peers[0] == NULL disables the folding
 * (the caller then brackets the kernels with its own barriers). */
#define AGS_SYNC_WORDS 64
#define AGS_SYNC_VIS 0      /* ags_dist_vis_local  -> ags_dist_vis_sum */
#define AGS_SYNC_TERMS 1    /* ags_dist_terms_put  -> ags_dist_wait(AGS_SYNC_TERMS) before the D2H of the gather buffer */
#define AGS_SYNC_GRADS 2    /* start of ags_dist_adam_step (this rank's gradients are complete) -> its reduce phase */
#define AGS_SYNC_PARAMS 3   /* end of ags_dist_adam_step -> ags_dist_wait(AGS_SYNC_PARAMS) before the next forward */
typedef struct AgsDistSync {
    int32_t* peers[AGS_MAX_PEERS];  /* the flags array of every rank (index = rank) */
    int32_t epoch;
} AgsDistSync;
/* one tiny kernel: wait until all ranks have signalled `phase` for sync->epoch */
int ags_dist_wait(const AgsDistSync* sync, int32_t phase, int32_t world, int32_t rank, void* stream);

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
    AgsDistSync sync;                           /* folded ordering (AGS_SYNC_GRADS in, AGS_SYNC_PARAMS out) */
} AgsDistAdamArgs;
int ags_dist_adam_step(const AgsDistAdamArgs* args);

/* The two small exchanges of the frame-sharded iteration over NVLink peer memory (new; csrc/dist_loss.cu).
 * Visibility count of quirk Q1 (mapping/gaussian_map.py:116-117) summed over the frames of ALL ranks:
 *   ags_dist_vis_local  counts this rank's B frames (opacity > 1e-3) into its symmetric plane vis_local;
 *   -- cross-GPU barrier on the stream --
 *   ags_dist_vis_sum    vis_count[p] = sum over ranks of their planes (multimem.ld_reduce if
 *                       vis_multicast != NULL, else one load per peer); feed it to AgsLossArgs.vis_count. */
typedef struct AgsDistVisArgs {
    int32_t world, rank;
    int32_t B, H, W;
    const float* opacity;                       /* (B,1,H,W) this rank's rendered opacity */
    int32_t* vis_local;                         /* symmetric (H*W) int32, this rank's copy */
    const int32_t* vis_peers[AGS_MAX_PEERS];    /* the same buffer on every rank */
    const int32_t* vis_multicast;               /* NVLS multicast address of it, or NULL */
    int32_t* vis_count;                         /* local (H,W) output of ags_dist_vis_sum */
    void* stream;
    const float* frame_weight;                  /* optional (B): frames with weight 0 (padding) are not counted */
    AgsDistSync sync;                           /* folded ordering (AGS_SYNC_VIS) */
} AgsDistVisArgs;
int ags_dist_vis_local(const AgsDistVisArgs* args);
int ags_dist_vis_sum(const AgsDistVisArgs* args);

/* All-gather of the per-rank loss terms / per-frame performance (the sampler on every rank needs all
 * of them, mapping/utils.py:206-218) plus (instances, overflow) of the forward: rank r stores its
 * nterm floats into slot r of every rank's gather buffer; a cross-GPU barrier must follow. */
typedef struct AgsDistTermsArgs {
    int32_t world, rank;
    int32_t nterm;                              /* floats per rank: terms | nview view costs | instances, overflow */
    int32_t nview;                              /* views of this rank (per-view instance counts gathered too) */
    const float* terms;                         /* local (nterm - nview - 2) */
    const int32_t* stats;                       /* local AgsRenderArgs.stats */
    float* gather_peers[AGS_MAX_PEERS];         /* symmetric (world*nterm) float buffer on every rank */
    float* gather_multicast;                    /* its NVLS multicast address or NULL */
    void* stream;
    AgsDistSync sync;                           /* folded ordering (AGS_SYNC_TERMS) */
} AgsDistTermsArgs;
int ags_dist_terms_put(const AgsDistTermsArgs* args);

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

/* ---------------------------------------------------------------------------------------------
 * Per-keyframe map maintenance: the map lives in CAPACITY
 * buffers owned by the caller (row-major SoA, `capacity` rows each); the kernels below append to /
 * compact them on the device, and the caller reads back only the new row count.
 *
 * ags_spawn   mapping/gaussian_map.py:294-468 (add_gaussians) after the bilateral filter and the
 *             optional render of the current map: per pixel back-projection (utils/operations.py:
 *             372-392,464-478,544-569), normal from the smoothed depth with the hard-wired 60 x 60
 *             degree fov (gaussian_map.py:316-322 -> operations.py:172-219), back-facing / NaN
 *             rejection (:324-335,386-387), normal2rotation (operations.py:481-541), cal_mask
 *             (gaussian_map.py:470-489), the voxel filter (operations.py:603-625: ONE randomly chosen
 *             candidate per occupied voxel of edge voxel_size; here a device hash table with a random
 *             priority per candidate instead of unique/randperm/scatter), and the append of the
 *             survivors in pixel order (gaussian_map.py:403-462: raw scales (0,0,-1e10), raw opacity
 *             0, colour = pixel rgb, zero view statistics).
 *             Coordinates must satisfy |p / voxel_size| < 2^20 (20 km at 2 cm).
 *             counters[0] = rows appended (clamped so that n_old + appended <= capacity),
 *             counters[1] = candidates before the voxel filter, counters[2] = survivors wanted
 *             (> counters[0] only if the capacity was too small: grow and call again). */
typedef struct AgsSpawnArgs {
    int32_t H, W;
    const float* rgb;              /* (3,H,W) keyframe colour */
    const float* depth;            /* (1,H,W) sensor depth, <= 0 invalid */
    const float* depth_smooth;     /* (1,H,W) ags_smooth_depth(depth) */
    float c2w[16];                 /* keyframe extrinsic, row-major camera-to-world, BY VALUE */
    float Kinv[9];                 /* inverse of the normalised intrinsic, row-major, BY VALUE */
    const float* pred_rgb;         /* (3,H,W) render of the current map at this pose, or NULL (empty map) */
    const float* pred_depth;       /* (1,H,W) */
    const float* pred_opacity;     /* (1,H,W) */
    float error_thres;             /* cfg.error_thres */
    float voxel_size;              /* 0.02; <= 0 disables the voxel filter */
    uint32_t seed;                 /* random priorities of the voxel filter */
    int32_t n_old, capacity;       /* rows in use / rows allocated in the buffers below */
    float* means;                  /* (capacity,3) */
    float* scales;                 /* (capacity,3) */
    float* rotations;              /* (capacity,4) */
    float* opacities;              /* (capacity)   */
    float* harmonics;              /* (capacity,3) */
    float* view_scores;            /* (capacity)   */
    float* view_supports;          /* (capacity)   */
    float* view_means;             /* (capacity,3) */
    int32_t* counters;             /* (4) device, see above */
    uint8_t* select_out;           /* optional (H*W): 1 = candidate before the voxel filter, 2 = appended */
    void* workspace; size_t workspace_bytes;   /* >= ags_spawn_scratch_bytes(H, W) */
    void* stream;
} AgsSpawnArgs;
size_t ags_spawn_scratch_bytes(int32_t H, int32_t W);
int ags_spawn(const AgsSpawnArgs* args);

/* ags_view_stats_update   mapping/gaussian_map.py:195-227: confidence bookkeeping after a keyframe.
 *             For every Gaussian counted (count_last >= 1) in the newest keyframe: supports += 1,
 *             view_means += (dir - view_means)/supports, view_scores += (1 - clamp(dist/depth_max))
 *             * clamp(normal . dir), dir = normalised vector to the camera, normal = third column
 *             of R(normalize(raw rotation)).  In place, one thread per Gaussian. */
int ags_view_stats_update(int32_t N, const int32_t* count_last, const float* means, const float* rotations_raw,
                          float cam_x, float cam_y, float cam_z, float depth_max, int32_t use_view_distribution,
                          float* view_supports, float* view_means, float* view_scores, void* stream);

/* ags_prune_compact   mapping/gaussian_map.py:229-246 (post_processing's prune + prune()): drops the
 *             Gaussians flagged in prune_mask (which is OR-ed IN PLACE with sigmoid(opacity) < 0.1,
 *             quirk Q5) and, if `counts` is given, those never counted in any of the T views
 *             (sum_t counts[t][i] < 1); the survivors of all eight SoA tensors are written, order
 *             preserved, to the dst buffers (src and dst must not overlap: ping-pong stores).
 *             n_kept (device int32) receives the new row count. */
typedef struct AgsPruneArgs {
    int32_t N, T;
    const int32_t* counts;         /* (T,N) or NULL */
    uint8_t* prune_mask;           /* (N) in/out, or NULL (treated as all zero, not written) */
    const float* src[8];           /* means, scales, rotations, opacities, harmonics, view_scores, view_supports, view_means */
    float* dst[8];
    int32_t* n_kept;               /* device */
    void* workspace; size_t workspace_bytes;   /* >= ags_prune_scratch_bytes(N) */
    void* stream;
} AgsPruneArgs;
size_t ags_prune_scratch_bytes(int32_t N);
int ags_prune_compact(const AgsPruneArgs* args);

/* ags_view_utility   planning/confidence.py:69-101 and planning/exploration.py:62-86 for V rendered
 *             candidate views at once (the reference loops over views with ~40 ATen launches and a
 *             nonzero sync each): explore[v] = |{voxels visible in view v (mapping/voxel_map.py:
 *             226-278) and unexplored}| / M, exploit[v] = mean((1 - conf') * depth' / depth_hi).
 *             One CTA per view; no atomics (deterministic). */
typedef struct AgsUtilityArgs {
    int32_t V, h, w, M;
    const float* depth;            /* (V,h,w) rendered depth, 0 where nothing was rendered */
    const float* confidence;       /* (V,h,w) rendered confidence */
    const uint8_t* valid;          /* (V,h,w) simulator validity mask or NULL (all valid) */
    const float* voxel_centers;    /* (M,3) */
    const uint8_t* unexplored;     /* (M) */
    const float* w2c;              /* (V,16) inverse extrinsics, row-major */
    const float* K;                /* (V,9) normalised intrinsics, row-major */
    float depth_lo, depth_hi;      /* simulator.depth_range */
    float* explore;                /* (V) */
    float* exploit;                /* (V) */
    void* stream;
} AgsUtilityArgs;
int ags_view_utility(const AgsUtilityArgs* args);

/* ags_voxel_roi   mapping/voxel_map.py:70-113 (the Gaussian half of VoxelMap.update_utility): voxels that
 *             hold more than min_gaussian_per_voxel opaque (sigmoid(opacity) > opacity_thres) but
 *             low-confidence (confidence < confidence_thres) Gaussians, and their mean surfel normal.
 *             Activations are applied here (raw rotations / opacities; `confidences` already activated,
 *             GaussianMap.get_confidences).  voxel index = floor((mean - bbox_min) / voxel_size) per
 *             axis (fp32 division, like the reference), linear index x*dim_y*dim_z + y*dim_z + z
 *             (:185-194).  voxel_count / voxel_normal are outputs AND the accumulators (zeroed here). */
typedef struct AgsVoxelRoiArgs {
    int32_t N;
    const float* means;            /* (N,3) */
    const float* rotations;        /* (N,4) raw */
    const float* opacities;        /* (N) raw (logit) */
    const float* confidences;      /* (N) activated */
    float bbox_min[3];
    float voxel_size[3];
    int32_t dim[3];
    float confidence_thres;        /* 0.3 */
    float opacity_thres;           /* 0.7 */
    int32_t min_gaussian_per_voxel;
    int32_t* voxel_count;          /* (M) selected Gaussians per voxel, M = dim[0]*dim[1]*dim[2] */
    float* voxel_normal;           /* (M,3) normalised mean normal where update_mask, else 0 */
    uint8_t* update_mask;          /* (M) voxel_count > min_gaussian_per_voxel */
    void* stream;
} AgsVoxelRoiArgs;
int ags_voxel_roi(const AgsVoxelRoiArgs* args);

const char* ags_last_error(void);
int ags_version(void);
/* (new) number of kernels this library has launched in the calling process so far (bench.py's
 * `gpu_launches` is the difference over the timed region; no reference counterpart) */
unsigned long long ags_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AGS_B200_H */
