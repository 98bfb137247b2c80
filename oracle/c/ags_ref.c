/* ags_ref.c -- CPU oracle #2 of the native half (TEST INFRASTRUCTURE ONLY): plain C + OpenMP
 * restatement of the rasterizer specification in DESIGN.md section 2, forward and hand-derived
 * backward, one view per call.  It is independent of oracle/rasterizer_ref.py (torch autograd) and is
 * cross-checked against it in tests/test_oracle_c.py; bench.py times it as the CPU baseline / the
 * `--impl reference` arm.  Only tests/, __graft_entry__ and bench.py may load it.
 *
 *     *** PARITY UNPINNED *** (same caveat as oracle/rasterizer_ref.py: the reference's own CUDA
 *     extension, diff_gaussian_rasterization_2d, envs/requirements.txt:15, is un-vendored and
 *     un-pinned; call site /root/reference/utils/operations.py:682-713.)
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -fPIC -shared, -DREAL=float or -DREAL=double).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL double
#endif
typedef REAL real;

#define TILE 16
#define LOWPASS ((real)0.3)
/* Hard thresholds of the specification.  g_shift (default 0 = the specification) moves EVERY hard
 * decision of the rasterizer by a relative amount: the parity tests evaluate the float64 oracle at
 * +-shift to find the elements whose value hinges on a comparison that fp32 rounding can flip
 * (tests/parity_util.py); nothing else uses it. */
static double g_shift = 0.0;
void agsref_set_threshold_shift(double rel) { g_shift = rel; }
#define SHIFTED(x) ((real)((x) * (1.0 + g_shift)))
/* The compared quantities carry different fp32 noise: alpha inherits the ~1e-4 px error of the projected
 * centre through the exponent (relative error ~1e-4 near the 1/255 cut-off), T and the blend weight
 * accumulate it over the splats in front; geometric quantities are good to ~1e-6.  The shift of each
 * threshold is a fixed multiple of the base shift so that it exceeds that noise ~10x. */
#define SHIFTED_K(x, k) ((real)((x) * (1.0 + (k) * g_shift)))
#define NEAR_CULL SHIFTED(0.2)
#define ALPHA_MAX SHIFTED_K(0.99, 50.0)
#define ALPHA_MIN SHIFTED_K(1.0 / 255.0, 50.0)
#define T_EPS SHIFTED_K(1e-4, 100.0)
#define SLOPE_COS_MIN SHIFTED(0.1)

typedef struct {
    /* sizes */
    int32_t N, H, W;
    int32_t require_importance, front_only;
    real tanfovx, tanfovy, scale_modifier, weight_thres;
    /* inputs (activated, like the module boundary) */
    const real *means, *scales, *rots, *opac, *colors, *conf; /* (N,3)(N,3)(N,4)(N)(N,3)(N) */
    const real *view, *proj, *bg;                             /* 16, 16, 3 */
    const real *mask;                                         /* H*W or NULL */
    /* outputs */
    real *rgb, *normal, *depth, *opacity, *confidence;        /* planar (C,H,W) */
    real *importance;                                         /* N */
    int32_t *count, *radii;                                   /* N */
} RefFwd;

typedef struct {
    const real *d_rgb, *d_normal, *d_depth, *d_opacity, *d_conf; /* may be NULL */
    real *d_means, *d_means2d, *d_opac, *d_colors, *d_scales, *d_rots; /* (N,3)(N,3)(N)(N,3)(N,3)(N,4) */
} RefBwd;

/* ---------------------------------------------------------------- per-Gaussian projection */
typedef struct {
    real t[3], inv_w, ndcx, ndcy, xg, yg, fx, fy, ux, uy;
    int ux_free, uy_free;
    real T0[3], T1[3], ST0[3], ST1[3], a, b, c, det, ca, cb, cc;
    real nv[3], sigma_n, cosv, c0, Dc;
    int Dc_free;
    real sx, sy;
    real R[9], s[3], Sig[6];
    int radius, valid, minx, miny, maxx, maxy;
} Proj;

static void sym_mul(const real* S, const real* v, real* o) {
    o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}

static void project_one(const RefFwd* a, int i, Proj* o) {
    const real* V = a->view;
    const real* M = a->proj;
    const real px = a->means[3 * i], py = a->means[3 * i + 1], pz = a->means[3 * i + 2];
    const real r = a->rots[4 * i], x = a->rots[4 * i + 1], y = a->rots[4 * i + 2], z = a->rots[4 * i + 3];
    real* R = o->R;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - r * z); R[2] = 2 * (x * z + r * y);
    R[3] = 2 * (x * y + r * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (x * z - r * y); R[7] = 2 * (y * z + r * x); R[8] = 1 - 2 * (x * x + y * y);
    for (int k = 0; k < 3; ++k) o->s[k] = a->scales[3 * i + k] * a->scale_modifier;
    const real s0 = o->s[0] * o->s[0], s1 = o->s[1] * o->s[1], s2 = o->s[2] * o->s[2];
    real* S = o->Sig;
    S[0] = R[0] * R[0] * s0 + R[1] * R[1] * s1 + R[2] * R[2] * s2;
    S[1] = R[0] * R[3] * s0 + R[1] * R[4] * s1 + R[2] * R[5] * s2;
    S[2] = R[0] * R[6] * s0 + R[1] * R[7] * s1 + R[2] * R[8] * s2;
    S[3] = R[3] * R[3] * s0 + R[4] * R[4] * s1 + R[5] * R[5] * s2;
    S[4] = R[3] * R[6] * s0 + R[4] * R[7] * s1 + R[5] * R[8] * s2;
    S[5] = R[6] * R[6] * s0 + R[7] * R[7] * s1 + R[8] * R[8] * s2;
    for (int j = 0; j < 3; ++j) o->t[j] = px * V[j] + py * V[4 + j] + pz * V[8 + j] + V[12 + j];
    const real homx = px * M[0] + py * M[4] + pz * M[8] + M[12];
    const real homy = px * M[1] + py * M[5] + pz * M[9] + M[13];
    const real homw = px * M[3] + py * M[7] + pz * M[11] + M[15];
    o->inv_w = 1 / (homw + (real)1e-7);
    o->ndcx = homx * o->inv_w;
    o->ndcy = homy * o->inv_w;
    o->xg = ((o->ndcx + 1) * a->W - 1) * (real)0.5;
    o->yg = ((o->ndcy + 1) * a->H - 1) * (real)0.5;
    o->fx = a->W / (2 * a->tanfovx);
    o->fy = a->H / (2 * a->tanfovy);
    const real tz = o->t[2];
    const real limx = SHIFTED(1.3) * a->tanfovx, limy = SHIFTED(1.3) * a->tanfovy;
    const real rx = o->t[0] / tz, ry = o->t[1] / tz;
    o->ux = rx < -limx ? -limx : (rx > limx ? limx : rx);
    o->uy = ry < -limy ? -limy : (ry > limy ? limy : ry);
    o->ux_free = (rx >= -limx) && (rx <= limx);
    o->uy_free = (ry >= -limy) && (ry <= limy);
    const real J00 = o->fx / tz, J11 = o->fy / tz;
    const real J02 = -o->fx * (o->ux * tz) / (tz * tz), J12 = -o->fy * (o->uy * tz) / (tz * tz);
    for (int k = 0; k < 3; ++k) {
        o->T0[k] = J00 * V[4 * k + 0] + J02 * V[4 * k + 2];
        o->T1[k] = J11 * V[4 * k + 1] + J12 * V[4 * k + 2];
    }
    sym_mul(S, o->T0, o->ST0);
    sym_mul(S, o->T1, o->ST1);
    o->a = o->T0[0] * o->ST0[0] + o->T0[1] * o->ST0[1] + o->T0[2] * o->ST0[2] + LOWPASS;
    o->b = o->T0[0] * o->ST1[0] + o->T0[1] * o->ST1[1] + o->T0[2] * o->ST1[2];
    o->c = o->T1[0] * o->ST1[0] + o->T1[1] * o->ST1[1] + o->T1[2] * o->ST1[2] + LOWPASS;
    o->det = o->a * o->c - o->b * o->b;
    const real ds = (o->det == 0) ? 1 : o->det;
    o->ca = o->c / ds; o->cb = -o->b / ds; o->cc = o->a / ds;
    const real mid = (real)0.5 * (o->a + o->c);
    real disc = mid * mid - o->det;
    if (disc < (real)0.1) disc = (real)0.1;
    const real lam1 = mid + sqrt(disc);
    o->radius = (int)ceil(SHIFTED(3.0) * sqrt(lam1));
    real nw[3] = {R[2], R[5], R[8]}, nv[3];
    for (int j = 0; j < 3; ++j) nv[j] = nw[0] * V[j] + nw[1] * V[4 + j] + nw[2] * V[8 + j];
    o->cosv = nv[0] * o->t[0] + nv[1] * o->t[1] + nv[2] * o->t[2];
    /* sign decisions on n.t: shifted by a multiple of |t| (the scale of its rounding error) */
    const real cos_thr = (real)(g_shift * sqrt((double)(o->t[0] * o->t[0] + o->t[1] * o->t[1] + o->t[2] * o->t[2])));
    o->sigma_n = (o->cosv > cos_thr) ? -1 : 1;
    for (int j = 0; j < 3; ++j) o->nv[j] = o->sigma_n * nv[j];
    o->c0 = o->nv[0] * o->t[0] + o->nv[1] * o->t[1] + o->nv[2] * o->t[2];
    const real d = o->c0 / tz;
    o->Dc_free = d <= -SLOPE_COS_MIN;
    o->Dc = d < -SLOPE_COS_MIN ? d : -SLOPE_COS_MIN;
    o->sx = -tz * o->nv[0] / (o->Dc * o->fx);
    o->sy = -tz * o->nv[1] / (o->Dc * o->fy);
    int valid = (tz > NEAR_CULL) && (o->det != 0);
    if (a->front_only && o->cosv >= cos_thr) valid = 0;
    const int tiles_x = (a->W + TILE - 1) / TILE, tiles_y = (a->H + TILE - 1) / TILE;
    const real rf = (real)o->radius;
#define CLAMPI(v, lo, hi) ((v) < (lo) ? (lo) : ((v) > (hi) ? (hi) : (v)))
    const real ts = (real)(5.0 * g_shift);   /* tile-rect truncation: shifted in units of a tile */
    o->minx = CLAMPI((int)((o->xg - rf) / TILE + ts), 0, tiles_x);
    o->miny = CLAMPI((int)((o->yg - rf) / TILE + ts), 0, tiles_y);
    o->maxx = CLAMPI((int)((o->xg + rf + TILE - 1) / TILE + ts), 0, tiles_x);
    o->maxy = CLAMPI((int)((o->yg + rf + TILE - 1) / TILE + ts), 0, tiles_y);
    if ((o->maxx - o->minx) * (o->maxy - o->miny) <= 0) valid = 0;
    if (!(o->xg == o->xg) || !(o->yg == o->yg)) valid = 0;
    o->valid = valid;
    if (!valid) o->radius = 0;
}

/* ---------------------------------------------------------------- state kept between fwd and bwd */
typedef struct {
    int32_t N, H, W, tiles_x, tiles_y, n_inst;
    Proj* P;            /* N */
    int32_t* tile_off;  /* tiles+1 */
    int32_t* inst;      /* n_inst: Gaussian ids sorted per tile by (float32 depth, id) */
    real* final_T;      /* H*W */
    int32_t* n_contrib; /* H*W */
} RefState;

typedef struct { float d; int32_t id; } KeyT;
static int cmp_key(const void* pa, const void* pb) {
    const KeyT* a = (const KeyT*)pa; const KeyT* b = (const KeyT*)pb;
    if (a->d < b->d) return -1;
    if (a->d > b->d) return 1;
    return (a->id > b->id) - (a->id < b->id);
}

void agsref_free(RefState* s) {
    if (!s) return;
    free(s->P); free(s->tile_off); free(s->inst); free(s->final_T); free(s->n_contrib); free(s);
}

static real eval_alpha(const Proj* g, real o, real pxf, real pyf, real* dx, real* dy, real* power, real* G) {
    *dx = g->xg - pxf;
    *dy = g->yg - pyf;
    *power = (real)-0.5 * (g->ca * *dx * *dx + g->cc * *dy * *dy) - g->cb * *dx * *dy;
    *G = exp(*power);
    const real al = o * *G;
    return al > ALPHA_MAX ? ALPHA_MAX : al;
}

/* ---------------------------------------------------------------- forward */
RefState* agsref_forward(const RefFwd* a) {
    const int N = a->N, H = a->H, W = a->W;
    RefState* s = (RefState*)calloc(1, sizeof(RefState));
    s->N = N; s->H = H; s->W = W;
    s->tiles_x = (W + TILE - 1) / TILE; s->tiles_y = (H + TILE - 1) / TILE;
    const int tiles = s->tiles_x * s->tiles_y;
    s->P = (Proj*)malloc(sizeof(Proj) * (size_t)(N > 0 ? N : 1));
    s->tile_off = (int32_t*)calloc((size_t)tiles + 1, sizeof(int32_t));
    s->final_T = (real*)malloc(sizeof(real) * (size_t)H * W);
    s->n_contrib = (int32_t*)calloc((size_t)H * W, sizeof(int32_t));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        project_one(a, i, &s->P[i]);
        a->radii[i] = s->P[i].radius;
        a->importance[i] = 0;
        a->count[i] = 0;
    }
    /* counting sort by tile, then (depth as float32, id) inside each tile */
    int32_t* cnt = (int32_t*)calloc((size_t)tiles + 1, sizeof(int32_t));
    for (int i = 0; i < N; ++i) {
        const Proj* p = &s->P[i];
        if (!p->valid) continue;
        for (int ty = p->miny; ty < p->maxy; ++ty)
            for (int tx = p->minx; tx < p->maxx; ++tx) cnt[ty * s->tiles_x + tx]++;
    }
    for (int t = 0; t < tiles; ++t) s->tile_off[t + 1] = s->tile_off[t] + cnt[t];
    s->n_inst = s->tile_off[tiles];
    KeyT* keys = (KeyT*)malloc(sizeof(KeyT) * (size_t)(s->n_inst > 0 ? s->n_inst : 1));
    memset(cnt, 0, sizeof(int32_t) * ((size_t)tiles + 1));
    for (int i = 0; i < N; ++i) {
        const Proj* p = &s->P[i];
        if (!p->valid) continue;
        for (int ty = p->miny; ty < p->maxy; ++ty)
            for (int tx = p->minx; tx < p->maxx; ++tx) {
                const int t = ty * s->tiles_x + tx;
                /* depth key: float32 as specified; under a threshold shift every key moves by +-1.5 % of the
                 * shift (sign from a hash of the id) so that near-ties, whose order fp32 rounding of t_z
                 * decides, show up as flip-prone too */
                const double jit = g_shift * 0.015 * ((((uint32_t)i * 2654435761u) >> 31) ? 1.0 : -1.0);
                KeyT k; k.d = (float)((double)p->t[2] * (1.0 + jit)); k.id = i;
                keys[s->tile_off[t] + cnt[t]++] = k;
            }
    }
    s->inst = (int32_t*)malloc(sizeof(int32_t) * (size_t)(s->n_inst > 0 ? s->n_inst : 1));
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < tiles; ++t) {
        const int o0 = s->tile_off[t], n = s->tile_off[t + 1] - o0;
        if (n > 1) qsort(keys + o0, (size_t)n, sizeof(KeyT), cmp_key);
        for (int k = 0; k < n; ++k) s->inst[o0 + k] = keys[o0 + k].id;
    }
    free(keys); free(cnt);
    const size_t P = (size_t)H * W;
    /* compositing: one tile at a time, pixels inside */
#pragma omp parallel for schedule(dynamic, 2)
    for (int t = 0; t < tiles; ++t) {
        const int ty0 = (t / s->tiles_x) * TILE, tx0 = (t % s->tiles_x) * TILE;
        const int o0 = s->tile_off[t], n = s->tile_off[t + 1] - o0;
        for (int py = ty0; py < ty0 + TILE && py < H; ++py)
            for (int px = tx0; px < tx0 + TILE && px < W; ++px) {
                const size_t pix = (size_t)py * W + px;
                real T = 1, C0 = 0, C1 = 0, C2 = 0, N0 = 0, N1 = 0, N2 = 0, D = 0, Cf = 0;
                int last = 0;
                const int imp = a->require_importance && (!a->mask || a->mask[pix] == 1);
                for (int k = 0; k < n; ++k) {
                    const int id = s->inst[o0 + k];
                    const Proj* g = &s->P[id];
                    real dx, dy, power, G;
                    const real alpha = eval_alpha(g, a->opac[id], (real)px, (real)py, &dx, &dy, &power, &G);
                    if (power > 0 || alpha < ALPHA_MIN) continue;
                    const real test_T = T * (1 - alpha);
                    if (test_T < T_EPS) break;
                    const real w = alpha * T;
                    C0 += w * a->colors[3 * id]; C1 += w * a->colors[3 * id + 1]; C2 += w * a->colors[3 * id + 2];
                    D += w * (g->t[2] - g->sx * dx - g->sy * dy);
                    N0 += w * g->nv[0]; N1 += w * g->nv[1]; N2 += w * g->nv[2];
                    Cf += w * a->conf[id];
                    T = test_T;
                    last = k + 1;
                    if (imp && w > SHIFTED_K(a->weight_thres, 100.0)) {
#pragma omp atomic
                        a->count[id] += 1;
#pragma omp atomic
                        a->importance[id] += w;
                    }
                }
                const real A = 1 - T;
                a->rgb[pix] = C0 + T * a->bg[0]; a->rgb[P + pix] = C1 + T * a->bg[1]; a->rgb[2 * P + pix] = C2 + T * a->bg[2];
                a->normal[pix] = N0; a->normal[P + pix] = N1; a->normal[2 * P + pix] = N2;
                a->depth[pix] = A > 0 ? D / A : 0;
                a->opacity[pix] = A;
                a->confidence[pix] = Cf;
                s->final_T[pix] = T;
                s->n_contrib[pix] = last;
            }
    }
    return s;
}

/* ---------------------------------------------------------------- backward */
void agsref_backward(const RefFwd* a, const RefState* s, const RefBwd* g) {
    const int N = a->N, H = a->H, W = a->W;
    const size_t P = (size_t)H * W;
    const int tiles = s->tiles_x * s->tiles_y;
    real* ds = (real*)calloc((size_t)(N > 0 ? N : 1) * 16, sizeof(real)); /* per-Gaussian screen-space record */
#pragma omp parallel for schedule(dynamic, 2)
    for (int t = 0; t < tiles; ++t) {
        const int ty0 = (t / s->tiles_x) * TILE, tx0 = (t % s->tiles_x) * TILE;
        const int o0 = s->tile_off[t];
        for (int py = ty0; py < ty0 + TILE && py < H; ++py)
            for (int px = tx0; px < tx0 + TILE && px < W; ++px) {
                const size_t pix = (size_t)py * W + px;
                const int last = s->n_contrib[pix];
                if (last == 0) continue;
                const real Tf = s->final_T[pix], A = 1 - Tf;
                real gC[3] = {0, 0, 0}, gN[3] = {0, 0, 0}, gD = 0, gCf = 0, gA = 0, gdep = 0;
                if (g->d_rgb) for (int c = 0; c < 3; ++c) gC[c] = g->d_rgb[c * P + pix];
                if (g->d_normal) for (int c = 0; c < 3; ++c) gN[c] = g->d_normal[c * P + pix];
                if (g->d_depth) gdep = g->d_depth[pix];
                if (g->d_opacity) gA = g->d_opacity[pix];
                if (g->d_conf) gCf = g->d_conf[pix];
                const real depth_out = a->depth[pix];
                if (A > 0) { gD = gdep / A; gA -= gdep * depth_out / A; }
                const real bgdot = gC[0] * a->bg[0] + gC[1] * a->bg[1] + gC[2] * a->bg[2];
                real rem = gC[0] * (a->rgb[pix] - Tf * a->bg[0]) + gC[1] * (a->rgb[P + pix] - Tf * a->bg[1])
                         + gC[2] * (a->rgb[2 * P + pix] - Tf * a->bg[2])
                         + gN[0] * a->normal[pix] + gN[1] * a->normal[P + pix] + gN[2] * a->normal[2 * P + pix]
                         + gD * (depth_out * A) + gCf * a->confidence[pix] + Tf * (bgdot - gA);
                real T = 1;
                for (int k = 0; k < last; ++k) {
                    const int id = s->inst[o0 + k];
                    const Proj* p = &s->P[id];
                    real dx, dy, power, G;
                    const real o = a->opac[id];
                    const real alpha = eval_alpha(p, o, (real)px, (real)py, &dx, &dy, &power, &G);
                    if (power > 0 || alpha < ALPHA_MIN) continue;
                    const real w = alpha * T, one_m = 1 - alpha;
                    const real dpix = p->t[2] - p->sx * dx - p->sy * dy;
                    const real sdot = gC[0] * a->colors[3 * id] + gC[1] * a->colors[3 * id + 1] + gC[2] * a->colors[3 * id + 2]
                                    + gN[0] * p->nv[0] + gN[1] * p->nv[1] + gN[2] * p->nv[2] + gD * dpix + gCf * a->conf[id];
                    rem -= w * sdot;
                    const real dalpha = T * sdot - rem / one_m;
                    T *= one_m;
                    const int unclamped = (o * G <= ALPHA_MAX);
                    const real dpower = unclamped ? alpha * dalpha : 0;
                    const real wgD = w * gD;
                    real v[15];
                    v[0] = dpower * (-p->ca * dx - p->cb * dy) - wgD * p->sx;
                    v[1] = dpower * (-p->cc * dy - p->cb * dx) - wgD * p->sy;
                    v[2] = (real)-0.5 * dx * dx * dpower; v[3] = -dx * dy * dpower; v[4] = (real)-0.5 * dy * dy * dpower;
                    v[5] = unclamped ? G * dalpha : 0;
                    v[6] = w * gC[0]; v[7] = w * gC[1]; v[8] = w * gC[2];
                    v[9] = w * gN[0]; v[10] = w * gN[1]; v[11] = w * gN[2];
                    v[12] = wgD; v[13] = -wgD * dx; v[14] = -wgD * dy;
                    real* d = ds + (size_t)id * 16;
                    for (int q = 0; q < 15; ++q) {
#pragma omp atomic
                        d[q] += v[q];
                    }
                }
            }
    }
    /* chain to the boundary tensors */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        for (int k = 0; k < 3; ++k) { g->d_means[3 * i + k] = 0; g->d_means2d[3 * i + k] = 0; g->d_colors[3 * i + k] = 0; g->d_scales[3 * i + k] = 0; }
        for (int k = 0; k < 4; ++k) g->d_rots[4 * i + k] = 0;
        g->d_opac[i] = 0;
        const Proj* p = &s->P[i];
        if (!p->valid) continue;
        const real* V = a->view; const real* M = a->proj;
        const real* d = ds + (size_t)i * 16;
        const real dxg = d[0], dyg = d[1], dca = d[2], dcb = d[3], dcc = d[4];
        g->d_opac[i] = d[5];
        for (int k = 0; k < 3; ++k) g->d_colors[3 * i + k] = d[6 + k];
        real dnv[3] = {d[9], d[10], d[11]};
        const real dz = d[12], dsx = d[13], dsy = d[14];
        g->d_means2d[3 * i] = dxg; g->d_means2d[3 * i + 1] = dyg;
        const real tz = p->t[2];
        real dp[3], dt[3] = {0, 0, 0};
        {
            const real dndcx = dxg * (real)0.5 * W, dndcy = dyg * (real)0.5 * H;
            const real dhx = dndcx * p->inv_w, dhy = dndcy * p->inv_w;
            const real dhw = -(p->ndcx * dndcx + p->ndcy * dndcy) * p->inv_w;
            dp[0] = M[0] * dhx + M[1] * dhy + M[3] * dhw;
            dp[1] = M[4] * dhx + M[5] * dhy + M[7] * dhw;
            dp[2] = M[8] * dhx + M[9] * dhy + M[11] * dhw;
        }
        const real id2 = 1 / (p->det * p->det);
        const real da = (-p->c * p->c * dca + p->b * p->c * dcb - p->b * p->b * dcc) * id2;
        const real db = (2 * p->b * p->c * dca - (p->a * p->c + p->b * p->b) * dcb + 2 * p->a * p->b * dcc) * id2;
        const real dc = (-p->b * p->b * dca + p->a * p->b * dcb - p->a * p->a * dcc) * id2;
        const real* T0 = p->T0; const real* T1 = p->T1;
        real Gs[6];
        Gs[0] = 2 * da * T0[0] * T0[0] + 2 * db * T0[0] * T1[0] + 2 * dc * T1[0] * T1[0];
        Gs[1] = 2 * da * T0[0] * T0[1] + db * (T0[0] * T1[1] + T0[1] * T1[0]) + 2 * dc * T1[0] * T1[1];
        Gs[2] = 2 * da * T0[0] * T0[2] + db * (T0[0] * T1[2] + T0[2] * T1[0]) + 2 * dc * T1[0] * T1[2];
        Gs[3] = 2 * da * T0[1] * T0[1] + 2 * db * T0[1] * T1[1] + 2 * dc * T1[1] * T1[1];
        Gs[4] = 2 * da * T0[1] * T0[2] + db * (T0[1] * T1[2] + T0[2] * T1[1]) + 2 * dc * T1[1] * T1[2];
        Gs[5] = 2 * da * T0[2] * T0[2] + 2 * db * T0[2] * T1[2] + 2 * dc * T1[2] * T1[2];
        real dT0[3], dT1[3];
        for (int k = 0; k < 3; ++k) {
            dT0[k] = 2 * da * p->ST0[k] + db * p->ST1[k];
            dT1[k] = 2 * dc * p->ST1[k] + db * p->ST0[k];
        }
        real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; ++k) {
            dJ00 += dT0[k] * V[4 * k + 0]; dJ02 += dT0[k] * V[4 * k + 2];
            dJ11 += dT1[k] * V[4 * k + 1]; dJ12 += dT1[k] * V[4 * k + 2];
        }
        const real itz = 1 / tz, itz2 = itz * itz;
        dt[2] += -dJ00 * p->fx * itz2 - dJ11 * p->fy * itz2;
        dt[2] += dJ02 * p->fx * p->ux * itz2 + dJ12 * p->fy * p->uy * itz2;
        const real dux = -dJ02 * p->fx * itz, duy = -dJ12 * p->fy * itz;
        if (p->ux_free) { dt[0] += dux * itz; dt[2] += -dux * p->t[0] * itz2; }
        if (p->uy_free) { dt[1] += duy * itz; dt[2] += -duy * p->t[1] * itz2; }
        dt[2] += dz;
        dnv[0] += -tz / (p->Dc * p->fx) * dsx;
        dnv[1] += -tz / (p->Dc * p->fy) * dsy;
        dt[2] += -p->nv[0] / (p->Dc * p->fx) * dsx - p->nv[1] / (p->Dc * p->fy) * dsy;
        const real dDc = -(p->sx * dsx + p->sy * dsy) / p->Dc;
        if (p->Dc_free) {
            const real dc0 = dDc / tz;
            dt[2] += -dDc * p->c0 / (tz * tz);
            for (int j = 0; j < 3; ++j) { dnv[j] += dc0 * p->t[j]; dt[j] += dc0 * p->nv[j]; }
        }
        real dnw[3];
        for (int k = 0; k < 3; ++k)
            dnw[k] = p->sigma_n * (V[4 * k] * dnv[0] + V[4 * k + 1] * dnv[1] + V[4 * k + 2] * dnv[2]);
        for (int k = 0; k < 3; ++k) dp[k] += V[4 * k] * dt[0] + V[4 * k + 1] * dt[1] + V[4 * k + 2] * dt[2];
        const real* R = p->R;
        real M3[9], dM3[9], dR[9], dsc[3];
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) M3[3 * r + k] = R[3 * r + k] * p->s[k];
        for (int k = 0; k < 3; ++k) {
            real col[3] = {M3[k], M3[3 + k], M3[6 + k]}, o3[3];
            sym_mul(Gs, col, o3);
            dM3[k] = o3[0]; dM3[3 + k] = o3[1]; dM3[6 + k] = o3[2];
        }
        for (int k = 0; k < 3; ++k) {
            dsc[k] = dM3[k] * R[k] + dM3[3 + k] * R[3 + k] + dM3[6 + k] * R[6 + k];
            for (int r = 0; r < 3; ++r) dR[3 * r + k] = dM3[3 * r + k] * p->s[k];
        }
        dR[2] += dnw[0]; dR[5] += dnw[1]; dR[8] += dnw[2];
        const real r = a->rots[4 * i], x = a->rots[4 * i + 1], y = a->rots[4 * i + 2], z = a->rots[4 * i + 3];
        g->d_rots[4 * i + 0] = 2 * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
        g->d_rots[4 * i + 1] = 2 * (y * dR[1] + z * dR[2] + y * dR[3] - 2 * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2 * x * dR[8]);
        g->d_rots[4 * i + 2] = 2 * (-2 * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2 * y * dR[8]);
        g->d_rots[4 * i + 3] = 2 * (-2 * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2 * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
        for (int k = 0; k < 3; ++k) { g->d_means[3 * i + k] = dp[k]; g->d_scales[3 * i + k] = dsc[k] * a->scale_modifier; }
    }
    free(ds);
}

int agsref_num_instances(const RefState* s) { return s->n_inst; }
int agsref_sizeof_real(void) { return (int)sizeof(real); }
