// project.cu -- K1 project_fwd (+ per-tile counting) and K6 project_bwd.
//
// Replaces the per-Gaussian stage of the native extension behind
// /root/reference/utils/operations.py:701-713 and, in RAW mode, the activations of
// /root/reference/mapping/gaussian_map.py:529-545 (fused here and in the backward).
// K1: a CTA owns 512 Gaussians for all views (conservative cull per (Gaussian, view) pair, exact
// projection on the compacted survivors); K6 one thread per visible pair (compact list).
//
// HBM roofline: K1 reads 60 B/Gaussian and writes 72+8 B per (view, Gaussian); K6 reads 64 B grad
// record + 56 B params per visible (view, Gaussian) and writes 56 B/Gaussian. Pure streaming.
#include "ags_common.cuh"

namespace {

struct GaussAct {          // activated parameters of one Gaussian + what the activation backward needs
    float px, py, pz;
    float s[3];
    float q[4];            // unit quaternion r,x,y,z (as used)
    float o;
    float R[9];
    float Sig[6];          // xx xy xz yy yz zz
    // RAW mode bookkeeping
    float qn;              // |raw q|
    bool s_pass[3];        // clamp passes gradient
};

__device__ __forceinline__ void activate(GaussAct& g, const AgsRenderArgs& a, int i) {
    g.px = __ldg(a.means3D + 3 * i);
    g.py = __ldg(a.means3D + 3 * i + 1);
    g.pz = __ldg(a.means3D + 3 * i + 2);
    float sr[3] = {__ldg(a.scales + 3 * i), __ldg(a.scales + 3 * i + 1), __ldg(a.scales + 3 * i + 2)};
    float qr[4] = {__ldg(a.rotations + 4 * i), __ldg(a.rotations + 4 * i + 1),
                   __ldg(a.rotations + 4 * i + 2), __ldg(a.rotations + 4 * i + 3)};
    float orw = __ldg(a.opacities + i);
    if (a.param_mode == AGS_PARAMS_RAW) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float e = a.scale_factor * expf(sr[k]);
            g.s_pass[k] = (e <= a.scale_max);       // torch.clamp passes the gradient on [min,max]
            g.s[k] = fminf(fmaxf(e, 0.f), a.scale_max) * a.scale_modifier;
        }
        float n = sqrtf(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
        g.qn = fmaxf(n, 1e-12f);
#pragma unroll
        for (int k = 0; k < 4; ++k) g.q[k] = qr[k] / g.qn;
        g.o = 1.f / (1.f + expf(-orw));
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { g.s[k] = sr[k] * a.scale_modifier; g.s_pass[k] = true; }
#pragma unroll
        for (int k = 0; k < 4; ++k) g.q[k] = qr[k];
        g.qn = 1.f;
        g.o = orw;
    }
    const float r = g.q[0], x = g.q[1], y = g.q[2], z = g.q[3];
    g.R[0] = 1.f - 2.f * (y * y + z * z); g.R[1] = 2.f * (x * y - r * z); g.R[2] = 2.f * (x * z + r * y);
    g.R[3] = 2.f * (x * y + r * z); g.R[4] = 1.f - 2.f * (x * x + z * z); g.R[5] = 2.f * (y * z - r * x);
    g.R[6] = 2.f * (x * z - r * y); g.R[7] = 2.f * (y * z + r * x); g.R[8] = 1.f - 2.f * (x * x + y * y);
    const float s0 = g.s[0] * g.s[0], s1 = g.s[1] * g.s[1], s2 = g.s[2] * g.s[2];
    const float* R = g.R;
    g.Sig[0] = R[0] * R[0] * s0 + R[1] * R[1] * s1 + R[2] * R[2] * s2;
    g.Sig[1] = R[0] * R[3] * s0 + R[1] * R[4] * s1 + R[2] * R[5] * s2;
    g.Sig[2] = R[0] * R[6] * s0 + R[1] * R[7] * s1 + R[2] * R[8] * s2;
    g.Sig[3] = R[3] * R[3] * s0 + R[4] * R[4] * s1 + R[5] * R[5] * s2;
    g.Sig[4] = R[3] * R[6] * s0 + R[4] * R[7] * s1 + R[5] * R[8] * s2;
    g.Sig[5] = R[6] * R[6] * s0 + R[7] * R[7] * s1 + R[8] * R[8] * s2;
}

struct ViewProj {          // forward intermediates of one (view, Gaussian)
    float t[3];
    float homx, homy, inv_w, ndcx, ndcy;
    float xg, yg;
    float fx, fy;
    float ux, uy;          // clamped t.x/t.z , t.y/t.z
    bool ux_free, uy_free; // not clamped
    float J00, J02, J11, J12;
    float T0[3], T1[3];
    float ST0[3], ST1[3];  // Sigma*T0, Sigma*T1
    float a, b, c, det;
    float ca, cb, cc;
    float nv[3];           // flipped view-space normal
    float sigma_n;         // +1 / -1 flip
    float cosv;            // n.t before the flip
    float c0, Dc;          // after the flip / clamped cos
    bool Dc_free;
    float sx, sy;
    int radius;
    bool valid;
    int minx, miny, maxx, maxy;
};

__device__ __forceinline__ void sym_mul(const float* S, const float* v, float* o) {
    o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}

__device__ __forceinline__ void project_view(ViewProj& o, const GaussAct& g, const Cam& c, int H, int W,
                                              bool front_only) {
    const float* V = c.V;
    const float* M = c.M;
#pragma unroll
    for (int j = 0; j < 3; ++j) o.t[j] = g.px * V[j] + g.py * V[4 + j] + g.pz * V[8 + j] + V[12 + j];
    o.homx = g.px * M[0] + g.py * M[4] + g.pz * M[8] + M[12];
    o.homy = g.px * M[1] + g.py * M[5] + g.pz * M[9] + M[13];
    const float homw = g.px * M[3] + g.py * M[7] + g.pz * M[11] + M[15];
    o.inv_w = 1.f / (homw + 1e-7f);
    o.ndcx = o.homx * o.inv_w;
    o.ndcy = o.homy * o.inv_w;
    o.xg = ((o.ndcx + 1.f) * W - 1.f) * 0.5f;
    o.yg = ((o.ndcy + 1.f) * H - 1.f) * 0.5f;
    o.fx = W / (2.f * c.tanx);
    o.fy = H / (2.f * c.tany);
    const float tz = o.t[2];
    const float limx = 1.3f * c.tanx, limy = 1.3f * c.tany;
    const float rx = o.t[0] / tz, ry = o.t[1] / tz;
    o.ux = fminf(fmaxf(rx, -limx), limx);
    o.uy = fminf(fmaxf(ry, -limy), limy);
    o.ux_free = (rx >= -limx) && (rx <= limx);
    o.uy_free = (ry >= -limy) && (ry <= limy);
    o.J00 = o.fx / tz;
    o.J11 = o.fy / tz;
    o.J02 = -o.fx * (o.ux * tz) / (tz * tz);
    o.J12 = -o.fy * (o.uy * tz) / (tz * tz);
    // Wr[j][k] = V[4k + j]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.T0[k] = o.J00 * V[4 * k + 0] + o.J02 * V[4 * k + 2];
        o.T1[k] = o.J11 * V[4 * k + 1] + o.J12 * V[4 * k + 2];
    }
    sym_mul(g.Sig, o.T0, o.ST0);
    sym_mul(g.Sig, o.T1, o.ST1);
    o.a = o.T0[0] * o.ST0[0] + o.T0[1] * o.ST0[1] + o.T0[2] * o.ST0[2] + AGS_LOWPASS;
    o.b = o.T0[0] * o.ST1[0] + o.T0[1] * o.ST1[1] + o.T0[2] * o.ST1[2];
    o.c = o.T1[0] * o.ST1[0] + o.T1[1] * o.ST1[1] + o.T1[2] * o.ST1[2] + AGS_LOWPASS;
    o.det = o.a * o.c - o.b * o.b;
    const float det_safe = (o.det == 0.f) ? 1.f : o.det;
    o.ca = o.c / det_safe;
    o.cb = -o.b / det_safe;
    o.cc = o.a / det_safe;
    const float mid = 0.5f * (o.a + o.c);
    const float lam1 = mid + sqrtf(fmaxf(mid * mid - o.det, 0.1f));
    o.radius = (int)ceilf(3.f * sqrtf(lam1));
    // normal: third column of R, to view space, flipped toward the camera
    float nw[3] = {g.R[2], g.R[5], g.R[8]};
    float nv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) nv[j] = nw[0] * V[j] + nw[1] * V[4 + j] + nw[2] * V[8 + j];
    o.cosv = nv[0] * o.t[0] + nv[1] * o.t[1] + nv[2] * o.t[2];
    o.sigma_n = (o.cosv > 0.f) ? -1.f : 1.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) o.nv[j] = o.sigma_n * nv[j];
    o.c0 = o.nv[0] * o.t[0] + o.nv[1] * o.t[1] + o.nv[2] * o.t[2];
    const float d = o.c0 / tz;
    o.Dc_free = (d <= -AGS_SLOPE_COS_MIN);
    o.Dc = fminf(d, -AGS_SLOPE_COS_MIN);
    o.sx = -tz * o.nv[0] / (o.Dc * o.fx);
    o.sy = -tz * o.nv[1] / (o.Dc * o.fy);
    // culling + tile rect
    bool valid = (tz > AGS_NEAR_CULL) && (o.det != 0.f);
    if (front_only && o.cosv >= 0.f) valid = false;
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    const float rf = (float)o.radius;
    o.minx = min(tiles_x, max(0, (int)((o.xg - rf) / TILE)));
    o.miny = min(tiles_y, max(0, (int)((o.yg - rf) / TILE)));
    o.maxx = min(tiles_x, max(0, (int)((o.xg + rf + TILE - 1) / TILE)));
    o.maxy = min(tiles_y, max(0, (int)((o.yg + rf + TILE - 1) / TILE)));
    if ((o.maxx - o.minx) * (o.maxy - o.miny) <= 0) valid = false;
    if (!(o.xg == o.xg) || !(o.yg == o.yg) || !(rf == rf)) valid = false;   // NaN guard
    o.valid = valid;
    if (!valid) o.radius = 0;
    // Tight tile range (output preserving): inside the 3-sigma rect only tiles that the splat's cutoff
    // box reaches can hold a pixel with alpha >= 1/255 (alpha = o*exp(-q/2) >= 1/255 <=> q <= 2 ln(255 o),
    // extent along x = sqrt(tau * cov_xx)); the other tiles never see a contribution, so they get no
    // instance.  Conservative margins as in the compositing kernels' warp-level test.
    if (valid) {
        const float tau = 2.f * __logf(255.f * g.o);
        if (!(tau > 0.f)) {
            o.maxx = o.minx; o.maxy = o.miny;                // contributes nowhere: zero instances
        } else {
            const float hx = sqrtf(tau * o.a) * 1.0001f + 1e-2f, hy = sqrtf(tau * o.c) * 1.0001f + 1e-2f;
            // tile t covers pixel centres [16t, 16t+15]
            const int bx0 = (int)floorf((o.xg - hx) / TILE), bx1 = (int)floorf((o.xg + hx) / TILE) + 1;
            const int by0 = (int)floorf((o.yg - hy) / TILE), by1 = (int)floorf((o.yg + hy) / TILE) + 1;
            o.minx = max(o.minx, min(bx0, o.maxx)); o.maxx = min(o.maxx, max(bx1, o.minx));
            o.miny = max(o.miny, min(by0, o.maxy)); o.maxy = min(o.maxy, max(by1, o.miny));
        }
    }
}

// K1 ---------------------------------------------------------------------------------------------
// One CTA owns K1_G (256) consecutive Gaussians for ALL views of the batch: the means are loaded once (two
// Gaussians per thread) and the views are walked in chunks of K1_VCHUNK cameras staged in shared memory.
// Per chunk:
//   1. every (Gaussian, view) pair is tested against a CONSERVATIVE screen-space radius bound that
//      needs only the mean (and, in ACTIVATED mode, the scales):
//      a,c <= s_max^2 (f/t_z)^2 (1 + lim^2) + 0.3,  lambda_1 <= a + c + sqrt(0.1)  =>  R_b.
//      A pair whose tile rect is empty even with R_b cannot be visible (C2: ~85 % of the pairs); the
//      survivors are compacted into a shared-memory candidate list (one ballot + one shared atomic per
//      warp and view),
//   2. the exact projection (activations fused in RAW mode) runs on the compacted candidates only, so the
//      warps that execute the heavy path are dense even when the visible Gaussians are scattered over the
//      index range; per-tile counting with fire-and-forget REDs,
//   3. the visible pairs of the chunk are appended to the global compact list (one atomic per CTA and
//      chunk) that drives the scatter and K6.
#ifndef AGS_K1_THREADS
#define AGS_K1_THREADS 128
#endif
#ifndef AGS_K1_MINB
#define AGS_K1_MINB 8
#endif
constexpr int K1_THREADS = AGS_K1_THREADS;         // small CTAs: the kernel is a chain of dependent latencies
constexpr int K1_GPT = 2;                          // Gaussians per thread           (load -> cull -> barrier -> load ->
constexpr int K1_G = K1_THREADS * K1_GPT;          // Gaussians per CTA               project -> store), so many independent
constexpr int K1_VCHUNK = 8;                       // views per pass                  CTAs per SM hide it
constexpr int K1_LIST = K1_G * K1_VCHUNK;          // worst case: every pair of the pass is a candidate

__global__ void __launch_bounds__(K1_THREADS, AGS_K1_MINB)
project_fwd_kernel(AgsRenderArgs a, AgsWorkspace w, int for_backward) {
    __shared__ Cam s_cam[K1_VCHUNK];
    __shared__ int s_cand[K1_LIST], s_vis[K1_LIST];
    __shared__ int s_ncand, s_nvis, s_base;
    const int tid = threadIdx.x, lane = tid & 31;
    const int tiles_x = (a.W + TILE - 1) / TILE, tiles_y = (a.H + TILE - 1) / TILE;
    const int g0 = blockIdx.x * K1_G;
    // the CTA's means, once for all views
    float mx[K1_GPT], my[K1_GPT], mz[K1_GPT], smax[K1_GPT];
    bool have[K1_GPT];
#pragma unroll
    for (int r = 0; r < K1_GPT; ++r) {
        const int i = g0 + r * K1_THREADS + tid;
        have[r] = i < a.N;
        mx[r] = my[r] = mz[r] = 0.f; smax[r] = 0.f;
        if (have[r]) {
            mx[r] = __ldg(a.means3D + 3 * i); my[r] = __ldg(a.means3D + 3 * i + 1); mz[r] = __ldg(a.means3D + 3 * i + 2);
            if (a.param_mode == AGS_PARAMS_RAW) {
                smax[r] = a.scale_max * a.scale_modifier;
            } else {
                smax[r] = fmaxf(fmaxf(fabsf(__ldg(a.scales + 3 * i)), fabsf(__ldg(a.scales + 3 * i + 1))),
                                fabsf(__ldg(a.scales + 3 * i + 2))) * a.scale_modifier;
            }
        }
    }
    for (int v0 = 0; v0 < a.B; v0 += K1_VCHUNK) {
        const int nv = min(K1_VCHUNK, a.B - v0);
        __syncthreads();                                   // previous pass done with s_cam / lists
        for (int q = tid; q < nv * 34; q += K1_THREADS) {
            const int vl = q / 34, k = q - vl * 34, v = v0 + vl;
            if (k < 16) s_cam[vl].V[k] = __ldg(a.viewmatrix + v * 16 + k);
            else if (k < 32) s_cam[vl].M[k - 16] = __ldg(a.projmatrix + v * 16 + k - 16);
            else if (k == 32) s_cam[vl].tanx = __ldg(a.tanfov + v * 2);
            else s_cam[vl].tany = __ldg(a.tanfov + v * 2 + 1);
        }
        if (tid == K1_THREADS - 1) { s_ncand = 0; s_nvis = 0; }
        __syncthreads();
        // ---- phase 1: conservative cull, compaction of the candidates
        for (int vl = 0; vl < nv; ++vl) {
            const Cam& cam = s_cam[vl];
            const float* V = cam.V;
            const float* M = cam.M;
            const float fx = a.W / (2.f * cam.tanx), fy = a.H / (2.f * cam.tany);
            const float kx = fx * fx * (1.f + 1.69f * cam.tanx * cam.tanx), ky = fy * fy * (1.f + 1.69f * cam.tany * cam.tany);
            const size_t vN = (size_t)(v0 + vl) * a.N;
#pragma unroll
            for (int r = 0; r < K1_GPT; ++r) {
                bool maybe = false;
                if (have[r]) {
                    const float tz = mx[r] * V[2] + my[r] * V[6] + mz[r] * V[10] + V[14];
                    maybe = tz > AGS_NEAR_CULL;
                    if (maybe) {
                        const float homx = mx[r] * M[0] + my[r] * M[4] + mz[r] * M[8] + M[12];
                        const float homy = mx[r] * M[1] + my[r] * M[5] + mz[r] * M[9] + M[13];
                        const float homw = mx[r] * M[3] + my[r] * M[7] + mz[r] * M[11] + M[15];
                        const float iw = 1.f / (homw + 1e-7f);
                        const float xg = ((homx * iw + 1.f) * a.W - 1.f) * 0.5f, yg = ((homy * iw + 1.f) * a.H - 1.f) * 0.5f;
                        const float sz = smax[r] / tz;
                        const float Rb = ceilf(3.f * sqrtf(sz * sz * (kx + ky) + 2.f * AGS_LOWPASS + 0.3163f)) * 1.001f + 2.f;
                        if (xg == xg && yg == yg && fabsf(xg) < 1e9f && fabsf(yg) < 1e9f && Rb < 1e9f) {
                            const int minx = min(tiles_x, max(0, (int)((xg - Rb) / TILE)));
                            const int miny = min(tiles_y, max(0, (int)((yg - Rb) / TILE)));
                            const int maxx = min(tiles_x, max(0, (int)((xg + Rb + TILE - 1) / TILE)));
                            const int maxy = min(tiles_y, max(0, (int)((yg + Rb + TILE - 1) / TILE)));
                            maybe = (maxx - minx) * (maxy - miny) > 0;
                        }
                    }
                    if (!maybe) a.radii[vN + g0 + r * K1_THREADS + tid] = 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, maybe);
                if (m) {
                    int base = 0;
                    const int leader = __ffs(m) - 1;
                    if (lane == leader) base = atomicAdd(&s_ncand, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (maybe) s_cand[base + __popc(m & ((1u << lane) - 1u))] = ((r * K1_THREADS + tid) << 4) | vl;
                }
            }
        }
        __syncthreads();
        // ---- phase 2: exact projection of the candidates
        const int ncand = s_ncand;
        for (int c = tid; c < ncand; c += K1_THREADS) {
            const int code = s_cand[c];
            const int vl = code & 15, v = v0 + vl;
            const int i = g0 + (code >> 4);
            const size_t idx = (size_t)v * a.N + i;
            GaussAct g;
            activate(g, a, i);
            ViewProj p;
            project_view(p, g, s_cam[vl], a.H, a.W, a.front_only != 0);
            a.radii[idx] = p.valid ? p.radius : 0;
            if (!p.valid) continue;
            const float conf = a.confidences ? __ldg(a.confidences + i) : 0.f;
            w.geom0[idx] = make_float4(p.xg, p.yg, p.ca, p.cb);
            w.geom1[idx] = make_float4(p.cc, g.o, p.sx, p.sy);
            w.feat0[idx] = make_float4(__ldg(a.colors + 3 * i), __ldg(a.colors + 3 * i + 1),
                                       __ldg(a.colors + 3 * i + 2), p.t[2]);
            w.feat1[idx] = make_float4(p.nv[0], p.nv[1], p.nv[2], conf);
            w.rect[idx] = make_uint2((unsigned)p.minx | ((unsigned)p.maxx << 16),
                                     (unsigned)p.miny | ((unsigned)p.maxy << 16));
            if (for_backward) {
                float4* d = reinterpret_cast<float4*>(w.dsplat + idx * 16);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                d[0] = z; d[1] = z; d[2] = z; d[3] = z;
            }
            int32_t* tc = w.tile_count + (size_t)v * tiles_x * tiles_y;
            for (int ty = p.miny; ty < p.maxy; ++ty)
                for (int tx = p.minx; tx < p.maxx; ++tx) atomicAdd(tc + ty * tiles_x + tx, 1);
            s_vis[atomicAdd(&s_nvis, 1)] = (int)idx;
        }
        __syncthreads();
        // ---- phase 3: append to the global visible list
        const int nvis = s_nvis;
        if (tid == 0) s_base = nvis ? atomicAdd(w.counters + 1, nvis) : 0;
        __syncthreads();
        for (int c = tid; c < nvis; c += K1_THREADS) w.vis_list[s_base + c] = s_vis[c];
    }
}

// K6 ---------------------------------------------------------------------------------------------
// One thread per VISIBLE (view, Gaussian) pair (compact list from K1).  The whole chain down to the
// raw parameters is linear in the incoming gradient record, so every pair finishes its own
// contribution and adds 14 floats to the (pre-zeroed) parameter gradients with atomics; a Gaussian
// is visible in about one of the B views, so collisions are rare.
__global__ void __launch_bounds__(128)
zero_grads_kernel(AgsRenderGradArgs gr, int N, int B) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)N * 4; e += stride) {
        if (e < (size_t)N * 3) { gr.d_means3D[e] = 0.f; gr.d_scales[e] = 0.f; gr.d_colors[e] = 0.f; }
        gr.d_rotations[e] = 0.f;
        if (e < (size_t)N) gr.d_opacities[e] = 0.f;
    }
    if (gr.d_means2D)
        for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)B * N * 3; e += stride)
            gr.d_means2D[e] = 0.f;
}

__global__ void __launch_bounds__(128)
project_bwd_kernel(AgsRenderArgs a, AgsRenderGradArgs gr, AgsWorkspace w) {
    const int nvis = w.counters[1];
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nvis; e += stride) {
        const size_t idx = (size_t)w.vis_list[e];
        const int v = (int)(idx / a.N);
        const int i = (int)(idx - (size_t)v * a.N);
        GaussAct g;
        activate(g, a, i);
        Cam cam;
        load_cam(cam, a.viewmatrix, a.projmatrix, a.tanfov, v);
        ViewProj p;
        project_view(p, g, cam, a.H, a.W, false);
        float4* dr = reinterpret_cast<float4*>(w.dsplat + idx * 16);
        float rec[16];
        {
            const float4 d0 = dr[0], d1 = dr[1], d2 = dr[2], d3 = dr[3];
            rec[0] = d0.x; rec[1] = d0.y; rec[2] = d0.z; rec[3] = d0.w; rec[4] = d1.x; rec[5] = d1.y; rec[6] = d1.z; rec[7] = d1.w;
            rec[8] = d2.x; rec[9] = d2.y; rec[10] = d2.z; rec[11] = d2.w; rec[12] = d3.x; rec[13] = d3.y; rec[14] = d3.z; rec[15] = d3.w;
        }
        if (gr.clear_records) {   // consume-and-clear: a second backward on the same forward starts from zero again
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            dr[0] = z; dr[1] = z; dr[2] = z; dr[3] = z;
        }
        // the record holds the raw moments accumulated by composite_bwd (slots AGS_REC_*, ags_common.cuh);
        // the chain rule through the conic, the centre and the plane slopes is applied here, once per splat:
        //   power = -0.5 (a dx^2 + c dy^2) - b dx dy,  dx = x_splat - x_pixel,  depth(pix) = z - sx dx - sy dy,
        //   alpha = o * G  =>  dL/do = (sum dpower) / o
        const float M0 = rec[AGS_REC_PDX], M1 = rec[AGS_REC_PDY], M20 = rec[AGS_REC_PXX], M11 = rec[AGS_REC_PXY],
                    M02 = rec[AGS_REC_PYY];
        const float d_o = g.o > 0.f ? rec[AGS_REC_P1] / g.o : 0.f;
        const float dcol[3] = {rec[AGS_REC_C0], rec[AGS_REC_C0 + 1], rec[AGS_REC_C0 + 2]};
        float dnv[3] = {rec[AGS_REC_N0], rec[AGS_REC_N0 + 1], rec[AGS_REC_N0 + 2]};
        const float dz = rec[AGS_REC_WD], dsx = -rec[AGS_REC_WDX], dsy = -rec[AGS_REC_WDY];
        const float dxg = -(p.ca * M0 + p.cb * M1) - p.sx * dz;
        const float dyg = -(p.cc * M1 + p.cb * M0) - p.sy * dz;
        const float dca = -0.5f * M20, dcb = -M11, dcc = -0.5f * M02;
        if (gr.d_means2D) { gr.d_means2D[idx * 3 + 0] = dxg; gr.d_means2D[idx * 3 + 1] = dyg; }
        const float* V = cam.V;
        const float* M = cam.M;
        const float tz = p.t[2];
        float dp[3], dt[3] = {0.f, 0.f, 0.f};
        // ---- screen position -> mean
        {
            const float dndcx = dxg * 0.5f * a.W, dndcy = dyg * 0.5f * a.H;
            const float dhx = dndcx * p.inv_w, dhy = dndcy * p.inv_w;
            const float dhw = -(p.ndcx * dndcx + p.ndcy * dndcy) * p.inv_w;
            dp[0] = M[0] * dhx + M[1] * dhy + M[3] * dhw;
            dp[1] = M[4] * dhx + M[5] * dhy + M[7] * dhw;
            dp[2] = M[8] * dhx + M[9] * dhy + M[11] * dhw;
        }
        // ---- conic -> cov2
        float da, db, dc;
        {
            const float id2 = 1.f / (p.det * p.det);
            da = (-p.c * p.c * dca + p.b * p.c * dcb - p.b * p.b * dcc) * id2;
            db = (2.f * p.b * p.c * dca - (p.a * p.c + p.b * p.b) * dcb + 2.f * p.a * p.b * dcc) * id2;
            dc = (-p.b * p.b * dca + p.a * p.b * dcb - p.a * p.a * dcc) * id2;
        }
        // ---- cov2 = T Sigma T^T ; Gs = symmetric (G + G^T) of dL/dSigma: xx xy xz yy yz zz
        float Gs[6];
        {
            const float* T0 = p.T0; const float* T1 = p.T1;
            Gs[0] = 2.f * da * T0[0] * T0[0] + 2.f * db * T0[0] * T1[0] + 2.f * dc * T1[0] * T1[0];
            Gs[1] = 2.f * da * T0[0] * T0[1] + db * (T0[0] * T1[1] + T0[1] * T1[0]) + 2.f * dc * T1[0] * T1[1];
            Gs[2] = 2.f * da * T0[0] * T0[2] + db * (T0[0] * T1[2] + T0[2] * T1[0]) + 2.f * dc * T1[0] * T1[2];
            Gs[3] = 2.f * da * T0[1] * T0[1] + 2.f * db * T0[1] * T1[1] + 2.f * dc * T1[1] * T1[1];
            Gs[4] = 2.f * da * T0[1] * T0[2] + db * (T0[1] * T1[2] + T0[2] * T1[1]) + 2.f * dc * T1[1] * T1[2];
            Gs[5] = 2.f * da * T0[2] * T0[2] + 2.f * db * T0[2] * T1[2] + 2.f * dc * T1[2] * T1[2];
            float dT0[3], dT1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dT0[k] = 2.f * da * p.ST0[k] + db * p.ST1[k];
                dT1[k] = 2.f * dc * p.ST1[k] + db * p.ST0[k];
            }
            // T0 = J00*Wr[0,:] + J02*Wr[2,:],  Wr[j][k] = V[4k+j]
            float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dJ00 += dT0[k] * V[4 * k + 0];
                dJ02 += dT0[k] * V[4 * k + 2];
                dJ11 += dT1[k] * V[4 * k + 1];
                dJ12 += dT1[k] * V[4 * k + 2];
            }
            const float itz = 1.f / tz, itz2 = itz * itz;
            // J00 = fx/tz ; J02 = -fx*u/tz (u = clamp(tx/tz))
            dt[2] += -dJ00 * p.fx * itz2 - dJ11 * p.fy * itz2;
            dt[2] += dJ02 * p.fx * p.ux * itz2 + dJ12 * p.fy * p.uy * itz2;
            const float dux = -dJ02 * p.fx * itz, duy = -dJ12 * p.fy * itz;
            if (p.ux_free) { dt[0] += dux * itz; dt[2] += -dux * p.t[0] * itz2; }
            if (p.uy_free) { dt[1] += duy * itz; dt[2] += -duy * p.t[1] * itz2; }
        }
        // ---- depth + plane slopes + normal
        float dnw[3];
        {
            dt[2] += dz;
            // sx = -tz*nv.x/(Dc*fx)
            dnv[0] += -tz / (p.Dc * p.fx) * dsx;
            dnv[1] += -tz / (p.Dc * p.fy) * dsy;
            dt[2] += -p.nv[0] / (p.Dc * p.fx) * dsx - p.nv[1] / (p.Dc * p.fy) * dsy;
            const float dDc = -(p.sx * dsx + p.sy * dsy) / p.Dc;
            if (p.Dc_free) {
                const float dc0 = dDc / tz;
                dt[2] += -dDc * p.c0 / (tz * tz);
#pragma unroll
                for (int j = 0; j < 3; ++j) { dnv[j] += dc0 * p.t[j]; dt[j] += dc0 * p.nv[j]; }
            }
            // nv = sigma * Wr nw  -> dnw_k = sigma * sum_j V[4k+j] dnv_j
#pragma unroll
            for (int k = 0; k < 3; ++k)
                dnw[k] = p.sigma_n * (V[4 * k] * dnv[0] + V[4 * k + 1] * dnv[1] + V[4 * k + 2] * dnv[2]);
        }
        // ---- t = Wr p + tr
#pragma unroll
        for (int k = 0; k < 3; ++k) dp[k] += V[4 * k] * dt[0] + V[4 * k + 1] * dt[1] + V[4 * k + 2] * dt[2];
        // ---- Sigma = M3 M3^T, M3 = R diag(s):  dM3 = Gs * M3
        const float* R = g.R;
        float dR[9], ds[3];
        {
            float M3[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) M3[3 * r + k] = R[3 * r + k] * g.s[k];
            float dM3[9];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float col[3] = {M3[k], M3[3 + k], M3[6 + k]}, o3[3];
                sym_mul(Gs, col, o3);
                dM3[k] = o3[0]; dM3[3 + k] = o3[1]; dM3[6 + k] = o3[2];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ds[k] = dM3[k] * R[k] + dM3[3 + k] * R[3 + k] + dM3[6 + k] * R[6 + k];
#pragma unroll
                for (int r = 0; r < 3; ++r) dR[3 * r + k] = dM3[3 * r + k] * g.s[k];
            }
            dR[2] += dnw[0]; dR[5] += dnw[1]; dR[8] += dnw[2];
        }
        float dq[4];
        {
            const float r = g.q[0], x = g.q[1], y = g.q[2], z = g.q[3];
            dq[0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
            dq[1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
            dq[2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
            dq[3] = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
        }
        float dsr[3], dqr[4], dor;
        if (a.param_mode == AGS_PARAMS_RAW) {
#pragma unroll
            for (int k = 0; k < 3; ++k) dsr[k] = g.s_pass[k] ? ds[k] * g.s[k] : 0.f;   // d/draw of clamp(sf*exp)
            const float qd = g.q[0] * dq[0] + g.q[1] * dq[1] + g.q[2] * dq[2] + g.q[3] * dq[3];
#pragma unroll
            for (int k = 0; k < 4; ++k) dqr[k] = (dq[k] - g.q[k] * qd) / g.qn;
            dor = d_o * g.o * (1.f - g.o);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) dsr[k] = ds[k] * a.scale_modifier;
#pragma unroll
            for (int k = 0; k < 4; ++k) dqr[k] = dq[k];
            dor = d_o;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicAdd(gr.d_means3D + 3 * i + k, dp[k]);
            if (dsr[k] != 0.f) atomicAdd(gr.d_scales + 3 * i + k, dsr[k]);
            atomicAdd(gr.d_colors + 3 * i + k, dcol[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(gr.d_rotations + 4 * i + k, dqr[k]);
        atomicAdd(gr.d_opacities + i, dor);
    }
}

}  // namespace

int ags_launch_project_fwd(const AgsRenderArgs& a, const AgsWorkspace& w, bool for_backward) {
    if (a.N == 0) return 0;
    dim3 grid((a.N + K1_G - 1) / K1_G);
    ags_note_launch(); project_fwd_kernel<<<grid, K1_THREADS, 0, (cudaStream_t)a.stream>>>(a, w, for_backward ? 1 : 0);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int ags_launch_project_bwd(const AgsRenderArgs& a, const AgsRenderGradArgs& g, const AgsWorkspace& w) {
    if (a.N == 0) return 0;
    const int threads = 128;
    cudaStream_t st = (cudaStream_t)a.stream;
    if (!g.accumulate) {
        ags_note_launch(); zero_grads_kernel<<<148 * 4, threads, 0, st>>>(g, a.N, a.B);
        AGS_CHECK_CUDA(cudaGetLastError());
    }
    ags_note_launch(); project_bwd_kernel<<<148 * 8, threads, 0, st>>>(a, g, w);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
