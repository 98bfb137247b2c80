// loss.cu -- K8: fused per-pixel post-processing + training loss + gradients w.r.t. the rasterizer
// outputs.  Replaces ~150 small ATen launches per iteration of the reference:
//   /root/reference/utils/operations.py:714-718   mask = opacity>1e-2, normalize(normal)*mask,
//                                                 depth2normal (:172-219, quirk Q2 kept)
//   /root/reference/mapping/gaussian_map.py:106-124  masked L1 rgb/depth, normal TV, consistency
//                                                 (quirk Q1: (B,H,W)*(B,1,H,W) -> (B,B,H,W) mean)
//   /root/reference/mapping/utils.py:14-16,28-62,120-121
//   /root/reference/mapping/gaussian_map.py:132-139  track_performance per-frame means
// and the autograd backward of all of it down to d rgb / d depth / d normal(raw).
//
// ONE tiled kernel (loss_fused_kernel): a CTA owns a 32x16 pixel tile of one frame and stages, in shared
// memory, depth / opacity mask of the tile + a halo of 2 and the unit normals / visibility sum / depth
// mask of the tile + a halo of 1.  Phase 1 evaluates depth2normal and its adjoint (the pixel's own depth
// gradient and what it pushes to its four neighbours) for the tile + halo 1; phase 2 gathers the
// neighbours' contributions from shared memory, adds the L1 terms and the TV / consistency normal
// gradient and writes d_rgb / d_depth / d_normal ONCE.  No intermediate planes in global memory, no
// atomics on images, deterministic.
// HBM roofline: reads 12 planes (8 predicted + 4 ground truth) + the (H,W) visibility sum, writes 7 planes
// of B*H*W floats (+6 when the caller wants normal_unit / d2n): 76 B per pixel and frame.  The halo
// (1.20x / 1.41x of the tile) is served by L2.
#include <string.h>
#include "ags_common.cuh"

namespace {

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct FrameGeom {       // per-frame constants of depth2normal (quirk Q2)
    float ik00, ik11;    // 1/fov2focal(fov[0], H), 1/fov2focal(fov[1], W)
    float cx, cy;        // W/2, H/2
};

__device__ __forceinline__ FrameGeom frame_geom(const float* tanfov, int f, int H, int W) {
    FrameGeom g;
    g.ik00 = 2.f * __ldg(tanfov + 2 * f) * (1.f / (float)H);      // 1/H, 1/W fold to constants per launch
    g.ik11 = 2.f * __ldg(tanfov + 2 * f + 1) * (1.f / (float)W);
    g.cx = 0.5f * W;
    g.cy = 0.5f * H;
    return g;
}

// camera-space point of pixel (y,x) and its d2n mask (opacity > 1e-2); read-only path so the
// loads can be issued ahead of the kernel's stores
__device__ __forceinline__ F3 cam_point(const float* depth, const float* opac, const FrameGeom& g, int W,
                                        int y, int x, float& m) {
    const float d = __ldg(depth + y * W + x);
    m = (__ldg(opac + y * W + x) > 1e-2f) ? 1.f : 0.f;
    return f3((x - g.cx) * d * g.ik00, (y - g.cy) * d * g.ik11, d);
}

// the four masked difference vectors of depth2normal at pixel (y,x); replicate padding makes the
// out-of-image neighbour equal to the pixel itself, whose difference is exactly zero for a 0/1 mask.
// Branch-free: neighbours are fetched at clamped coordinates and the border is folded into the mask,
// so all ten loads are independent and issue back to back.
struct D2N {
    F3 pu, pl, pb, pr;
    float mu, ml, mb, mr, mc;
};

__device__ __forceinline__ D2N d2n_vectors(const float* depth, const float* opac, const FrameGeom& g,
                                           int H, int W, int y, int x) {
    D2N r;
    const int yu = max(y - 1, 0), yb = min(y + 1, H - 1), xl = max(x - 1, 0), xr = min(x + 1, W - 1);
    float mu, ml, mb, mr;
    const F3 c = cam_point(depth, opac, g, W, y, x, r.mc);
    const F3 qu = cam_point(depth, opac, g, W, yu, x, mu);
    const F3 ql = cam_point(depth, opac, g, W, y, xl, ml);
    const F3 qb = cam_point(depth, opac, g, W, yb, x, mb);
    const F3 qr = cam_point(depth, opac, g, W, y, xr, mr);
    r.mu = (y > 0) ? mu : 0.f;
    r.ml = (x > 0) ? ml : 0.f;
    r.mb = (y < H - 1) ? mb : 0.f;
    r.mr = (x < W - 1) ? mr : 0.f;
    const F3 pc = c * r.mc;
    r.pu = (qu - pc) * r.mu;
    r.pl = (ql - pc) * r.ml;
    r.pb = (qb - pc) * r.mb;
    r.pr = (qr - pc) * r.mr;
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// block-level accumulate of K values into global with one atomic per block per value
template <int K>
__device__ __forceinline__ void block_accumulate(float (&v)[K], float* const (&dst)[K]) {
    __shared__ float sm[K][8];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float s = warp_sum(v[k]);
        if (lane == 0) sm[k][wid] = s;
    }
    __syncthreads();
    if (tid < K) {
        float s = 0.f;
        const int nw = (blockDim.x * blockDim.y + 31) >> 5;
        for (int w = 0; w < nw; ++w) s += sm[tid][w];
        if (s != 0.f) atomicAdd(dst[tid], s);
    }
    __syncthreads();
}

__device__ __forceinline__ float vis_sum(const AgsLossArgs& a, size_t P, int p) {
    if (a.vis_count) return (float)__ldg(a.vis_count + p);
    float m = 0.f;
    for (int f = 0; f < a.B; ++f) {
        const float wf = a.frame_weight ? __ldg(a.frame_weight + f) : 1.f;       // padded frames do not count
        m += (wf != 0.f && __ldg(a.opacity + (size_t)f * P + p) > 1e-3f) ? 1.f : 0.f;
    }
    return m;
}

// pre-pass: per-pixel visibility count over the frames of this call (quirk Q1), once per pixel.  Four pixels per
// thread with 128-bit loads when the planes allow it (B loads in flight per thread; the scalar form ran at 0.9 TB/s)
__global__ void __launch_bounds__(256)
loss_vis_count(AgsLossArgs a, float* __restrict__ msum_plane, int vec4) {
    const size_t P = (size_t)a.H * a.W;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec4) {
        const size_t p = 4 * t;
        if (p >= P) return;
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.vis_count) {
            const int4 c = __ldg(reinterpret_cast<const int4*>(a.vis_count + p));
            m = make_float4((float)c.x, (float)c.y, (float)c.z, (float)c.w);
        } else {
            for (int f = 0; f < a.B; ++f) {
                const float wf = a.frame_weight ? __ldg(a.frame_weight + f) : 1.f;       // padded frames do not count
                const float4 o = __ldg(reinterpret_cast<const float4*>(a.opacity + (size_t)f * P + p));
                if (wf != 0.f) {
                    m.x += o.x > 1e-3f ? 1.f : 0.f; m.y += o.y > 1e-3f ? 1.f : 0.f;
                    m.z += o.z > 1e-3f ? 1.f : 0.f; m.w += o.w > 1e-3f ? 1.f : 0.f;
                }
            }
        }
        *reinterpret_cast<float4*>(msum_plane + p) = m;
        return;
    }
    if (t >= P) return;
    msum_plane[t] = vis_sum(a, P, (int)t);
}

#ifndef AGS_LOSS_MINB
#define AGS_LOSS_MINB 4
#endif

constexpr int LT_W = 32, LT_H = 16;                // output tile: 256 threads, two rows (y, y + 8) per thread
constexpr int LE_W = LT_W + 2, LE_H = LT_H + 2;    // tile + halo 1 (d2n adjoint, unit normals)
constexpr int LR_W = LT_W + 4, LR_H = LT_H + 4;    // tile + halo 2 (depth, opacity mask)
constexpr int LE_N = LE_W * LE_H, LR_N = LR_W * LR_H;
#define AGS_LOSS_MAX_FRAMES 64

struct LossFrames {        // per-frame ground-truth pointers (kernel parameter; NULL list -> stacked tensors)
    const float* rgb[AGS_LOSS_MAX_FRAMES];
    const float* depth[AGS_LOSS_MAX_FRAMES];
};

template <bool USE_LIST>
__global__ void __launch_bounds__(256, AGS_LOSS_MINB)
loss_fused_kernel(AgsLossArgs a, LossFrames fr, const float* __restrict__ msum_plane) {
    __shared__ float sD[LR_N], sM[LR_N];                         // depth, d2n mask (opacity > 1e-2); 0 outside
    __shared__ float sNU[3][LE_N], sMS[LE_N], sMD[LE_N], sIN[LE_N];   // unit normal, visibility sum, depth_gt > 0, inside the image
    __shared__ float sDN[3][LE_N];                               // d2n (= u * m2)
    __shared__ float sAdj[5][LE_N];                              // own, up, left, bottom, right depth adjoints
    const int H = a.H, W = a.W;
    const unsigned P = (unsigned)H * (unsigned)W;
    const unsigned f = blockIdx.z;
    const int tid = threadIdx.y * LT_W + threadIdx.x;          // block = (32, 8)
    const int x0 = blockIdx.x * LT_W, y0 = blockIdx.y * LT_H;
    const float wf = a.frame_weight ? __ldg(a.frame_weight + f) : 1.f;
    const float Bt = (float)a.B_total;
    const float inv_rgb = wf / (Bt * 3.f * (float)P);
    const float inv_d = wf / (Bt * (float)P);
    const float inv_cons = wf / (Bt * Bt * (float)P);
    const float inv_tv = wf / (Bt * 4.f * (float)P);
    const float inv2s2 = 1.f / (2.f * 0.3f * 0.3f);
    const float* depth = a.depth + f * P;
    const float* opac = a.opacity + f * P;
    const float* normal = a.normal + f * 3u * P;
    const float* depth_gt = USE_LIST ? fr.depth[f] : a.depth_gt + f * P;
    const float* rgb_gt = USE_LIST ? fr.rgb[f] : a.rgb_gt + f * 3u * P;
    const FrameGeom g = frame_geom(a.tanfov, f, H, W);
    // ---- stage: depth + mask with halo 2
    for (int i = tid; i < LR_N; i += 256) {
        const int ry = i / LR_W, rx = i - ry * LR_W;
        const int y = y0 - 2 + ry, x = x0 - 2 + rx;
        float d = 0.f, m = 0.f;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            d = __ldg(depth + y * W + x);
            m = (__ldg(opac + y * W + x) > 1e-2f) ? 1.f : 0.f;
        }
        sD[i] = d; sM[i] = m;
    }
    // ---- stage: unit normal, visibility sum, depth mask with halo 1
    for (int i = tid; i < LE_N; i += 256) {
        const int ey = i / LE_W, ex = i - ey * LE_W;
        const int y = y0 - 1 + ey, x = x0 - 1 + ex;
        F3 nu = f3(0.f, 0.f, 0.f);
        float ms = 0.f, md = 0.f, inside = 0.f;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            inside = 1.f;
            const unsigned p = (unsigned)y * W + x;
            const F3 n = f3(__ldg(normal + p), __ldg(normal + P + p), __ldg(normal + 2u * P + p));
            const float m2 = (__ldg(opac + p) > 1e-2f) ? 1.f : 0.f;
            nu = n * (m2 * rsqrtf(fmaxf(dot(n, n), 1e-24f)));
            ms = __ldg(msum_plane + p);
            md = (__ldg(depth_gt + p) > 0.f) ? 1.f : 0.f;
        }
        sNU[0][i] = nu.x; sNU[1][i] = nu.y; sNU[2][i] = nu.z; sMS[i] = ms; sMD[i] = md; sIN[i] = inside;
    }
    __syncthreads();
    // ---- phase 1: depth2normal + adjoint for the tile + halo 1
    for (int i = tid; i < LE_N; i += 256) {
        const int ey = i / LE_W, ex = i - ey * LE_W;
        const int y = y0 - 1 + ey, x = x0 - 1 + ex;
        F3 d2n = f3(0.f, 0.f, 0.f);
        float a_own = 0.f, a_up = 0.f, a_left = 0.f, a_bottom = 0.f, a_right = 0.f;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const int r = (ey + 1) * LR_W + (ex + 1);            // same pixel in the halo-2 region
            const float rx = (x - g.cx) * g.ik00, ry = (y - g.cy) * g.ik11;   // point = depth * (rx, ry, 1)
            // replicate padding: an out-of-image neighbour equals the pixel itself; with a 0/1 mask its
            // masked difference is exactly zero, which the zero mask staged outside the image reproduces
            const float dc = sD[r], mc = sM[r];
            const float du = sD[r - LR_W], mu = sM[r - LR_W];
            const float dl = sD[r - 1], ml = sM[r - 1];
            const float db = sD[r + LR_W], mb = sM[r + LR_W];
            const float dr = sD[r + 1], mr = sM[r + 1];
            const F3 pc = f3(rx * dc, ry * dc, dc) * mc;
            const F3 pu = (f3(rx * du, (ry - g.ik11) * du, du) - pc) * mu;
            const F3 pl = (f3((rx - g.ik00) * dl, ry * dl, dl) - pc) * ml;
            const F3 pb = (f3(rx * db, (ry + g.ik11) * db, db) - pc) * mb;
            const F3 pr = (f3((rx + g.ik00) * dr, ry * dr, dr) - pc) * mr;
            // cross(pu,pl) + cross(pr,pu) + cross(pb,pr) + cross(pl,pb) = (pu - pb) x (pl - pr): one cross product
            const F3 ev = pu - pb, eh = pl - pr;
            const F3 ns = cross(ev, eh);
            const float insn = rsqrtf(fmaxf(dot(ns, ns), 1e-24f));
            const F3 u = ns * insn;
            d2n = u * mc;
            const float msum = sMS[i];
            if (mc > 0.f && msum > 0.f) {
                const F3 nu = f3(sNU[0][i], sNU[1][i], sNU[2][i]);
                const float wc = -a.w_cons * msum * inv_cons;                // dL/d(nu . d2n)
                const F3 gd = nu * wc;                                       // dL/d u  (d2n = u*m2, m2 = 1)
                const F3 gq = (gd - u * dot(u, gd)) * insn;
                // adjoint of ns = ev x eh:  d ev = eh x gq,  d eh = gq x ev;  ev = pu - pb, eh = pl - pr
                const F3 dev = cross(eh, gq), deh = cross(gq, ev);
                const F3 dpu = dev * mu;
                const F3 dpl = deh * ml;
                const F3 dpb = dev * (-mb);
                const F3 dpr = deh * (-mr);
                const F3 dpc = (dpu + dpl + dpb + dpr) * (-mc);
                a_own = dpc.x * rx + dpc.y * ry + dpc.z;
                a_up = dpu.x * rx + dpu.y * (ry - g.ik11) + dpu.z;           // zero when mu == 0
                a_left = dpl.x * (rx - g.ik00) + dpl.y * ry + dpl.z;
                a_bottom = dpb.x * rx + dpb.y * (ry + g.ik11) + dpb.z;
                a_right = dpr.x * (rx + g.ik00) + dpr.y * ry + dpr.z;
            }
        }
        sDN[0][i] = d2n.x; sDN[1][i] = d2n.y; sDN[2][i] = d2n.z;
        sAdj[0][i] = a_own; sAdj[1][i] = a_up; sAdj[2][i] = a_left; sAdj[3][i] = a_bottom; sAdj[4][i] = a_right;
    }
    __syncthreads();
    // ---- phase 2: every thread finishes two pixels of the tile (rows ty and ty + 8)
    float fr_rgb = 0.f, fr_d = 0.f, acc_cons = 0.f, acc_tv = 0.f;
    const float ctv2 = 2.f * a.w_tv * inv_tv;
#pragma unroll
    for (int half = 0; half < LT_H / 8; ++half) {
        const int ty = threadIdx.y + 8 * half;
        const int x = x0 + threadIdx.x, y = y0 + ty;
        if (x >= W || y >= H) continue;
        const unsigned p = (unsigned)y * (unsigned)W + (unsigned)x;
        const unsigned o1 = f * P + p, o3 = f * 3u * P + p;
        const int e = (ty + 1) * LE_W + (threadIdx.x + 1);
        const int r = (ty + 2) * LR_W + (threadIdx.x + 2);
        const float msum = sMS[e];
        const float A = __ldg(opac + p);
        const float mvis = (A > 1e-3f) ? 1.f : 0.f;
        const float m2 = sM[r];
        // ---- L1 rgb + gradient
#pragma unroll
        for (unsigned c = 0; c < 3; ++c) {
            const float er = (__ldg(a.rgb + o3 + c * P) - __ldg(rgb_gt + p + c * P)) * mvis;
            fr_rgb += fabsf(er);
            a.d_rgb[o3 + c * P] = (er > 0.f ? inv_rgb : (er < 0.f ? -inv_rgb : 0.f));
        }
        // ---- L1 depth + gradient, plus the depth2normal adjoint gathered from the neighbours
        const float dp = sD[r];
        const float md_p = sMD[e];
        const float ed = (dp - __ldg(depth_gt + p)) * md_p;
        fr_d += fabsf(ed);
        float dd = a.w_depth * (ed > 0.f ? inv_d : (ed < 0.f ? -inv_d : 0.f));
        dd += sAdj[0][e];
        dd += sAdj[1][e + LE_W];        // "up" share of the pixel below
        dd += sAdj[2][e + 1];           // "left" share of the pixel to the right
        dd += sAdj[3][e - LE_W];        // "bottom" share of the pixel above
        dd += sAdj[4][e - 1];           // "right" share of the pixel to the left   (zero outside the image)
        a.d_depth[o1] = dd;
        // ---- consistency loss + normal gradient (consistency + TV, through normalize*mask)
        const F3 nu = f3(sNU[0][e], sNU[1][e], sNU[2][e]);
        const F3 d2n = f3(sDN[0][e], sDN[1][e], sDN[2][e]);
        acc_cons += (1.f - dot(nu, d2n)) * msum;
        F3 gnu = d2n * (-a.w_cons * msum * inv_cons);
        // TV: the one-sided difference towards neighbour q enters twice -- as p's own term (mask of p) and
        // as q's mirrored term that references p (mask of q); both share the difference vector, its
        // squared norm, the depth gate and the exponential, so they are evaluated once.
        // (the staged depth mask is 0 outside the image, so no border test is needed)
#define AGS_TV_NEIGHBOUR(EO, RO)                                                            \
        {                                                                                   \
            const F3 dl = nu - f3(sNU[0][e + (EO)], sNU[1][e + (EO)], sNU[2][e + (EO)]);    \
            const float nd = dot(dl, dl);                                                   \
            const float ddq = dp - sD[r + (RO)];                                            \
            const float mq = sMD[e + (EO)];                                                 \
            const float in_q = sIN[e + (EO)];                                               \
            const float gate = (ddq * ddq <= 1e-4f) ? 1.f : 0.f;                            \
            const float ex = __expf(-nd * inv2s2) * gate;                                   \
            acc_tv += md_p * in_q * ex * nd;                                                \
            const float coef = ex * (1.f - nd * inv2s2) * (md_p * in_q + mq);               \
            gnu = gnu + dl * (coef * ctv2);                                                 \
        }
        AGS_TV_NEIGHBOUR(1, 1)
        AGS_TV_NEIGHBOUR(-1, -1)
        AGS_TV_NEIGHBOUR(LE_W, LR_W)
        AGS_TV_NEIGHBOUR(-LE_W, -LR_W)
#undef AGS_TV_NEIGHBOUR
        const F3 n = f3(__ldg(normal + p), __ldg(normal + P + p), __ldg(normal + 2u * P + p));
        const float inn = rsqrtf(fmaxf(dot(n, n), 1e-24f));
        const F3 uh = n * inn;
        const F3 gn = (gnu - uh * dot(uh, gnu)) * (m2 * inn);
        a.d_normal[o3] = gn.x; a.d_normal[o3 + P] = gn.y; a.d_normal[o3 + 2u * P] = gn.z;
        if (a.normal_unit) { a.normal_unit[o3] = nu.x; a.normal_unit[o3 + P] = nu.y; a.normal_unit[o3 + 2u * P] = nu.z; }
        if (a.d2n) { a.d2n[o3] = d2n.x; a.d2n[o3 + P] = d2n.y; a.d2n[o3 + 2u * P] = d2n.z; }
    }
    // loss sums: every term carries the frame weight through its normaliser; the per-frame performance
    // (track_performance, gaussian_map.py:132-139) is the plain mean
    float v6[6] = {fr_rgb * inv_rgb, fr_d * inv_d, acc_cons * inv_cons, acc_tv * inv_tv,
                   fr_rgb / (3.f * (float)P), fr_d / (float)P};
    float* const d6[6] = {a.loss_terms + 0, a.loss_terms + 1, a.loss_terms + 2, a.loss_terms + 3,
                          a.loss_terms + 4 + 2 * f, a.loss_terms + 4 + 2 * f + 1};
    block_accumulate<6>(v6, d6);
}

// forward-only post-processing of rendered views (planners / eval / GUI): unit normal + d2n
__global__ void __launch_bounds__(256)
postprocess_kernel(int B, int H, int W, const float* normal, const float* depth_, const float* opacity,
                   const float* fov, float* normal_unit, float* d2n_out) {
    const size_t P = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (p >= P) return;
    const int y = (int)(p / W), x = (int)(p % W);
    const float* opac = opacity + (size_t)f * P;
    const float* depth = depth_ + (size_t)f * P;
    const float m2 = (opac[p] > 1e-2f) ? 1.f : 0.f;
    const float* np_ = normal + (size_t)f * 3 * P + p;
    const F3 n = f3(np_[0], np_[P], np_[2 * P]);
    const float nn = fmaxf(sqrtf(dot(n, n)), 1e-12f);
    const F3 nu = n * (m2 / nn);
    float* no = normal_unit + (size_t)f * 3 * P + p;
    no[0] = nu.x; no[P] = nu.y; no[2 * P] = nu.z;
    const FrameGeom g = frame_geom(fov, f, H, W);
    const D2N v = d2n_vectors(depth, opac, g, H, W, y, x);
    const F3 ns = cross(v.pu, v.pl) + cross(v.pr, v.pu) + cross(v.pb, v.pr) + cross(v.pl, v.pb);
    const float nsn = fmaxf(sqrtf(dot(ns, ns)), 1e-12f);
    const F3 d2n = ns * (m2 / nsn);
    float* dn = d2n_out + (size_t)f * 3 * P + p;
    dn[0] = d2n.x; dn[P] = d2n.y; dn[2 * P] = d2n.z;
}

}  // namespace

extern "C" int ags_postprocess(int32_t B, int32_t H, int32_t W, const float* normal, const float* depth,
                               const float* opacity, const float* fov, float* normal_unit, float* d2n,
                               void* stream) {
    AGS_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad sizes");
    AGS_CHECK_ARG(normal && depth && opacity && fov && normal_unit && d2n, "NULL pointer");
    dim3 grid((unsigned)(((size_t)H * W + 255) / 256), B);
    ags_note_launch(); postprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, H, W, normal, depth, opacity, fov,
                                                              normal_unit, d2n);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t ags_loss_scratch_bytes(int32_t B, int32_t H, int32_t W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return ags_align256((size_t)H * W * sizeof(float));       // the (H,W) visibility-sum plane
}

extern "C" int ags_loss_forward_backward(const AgsLossArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0 && a->B_total >= 1, "bad sizes B=%d H=%d W=%d B_total=%d",
                  a->B, a->H, a->W, a->B_total);
    const bool list = a->rgb_gt_frames_host != nullptr || a->depth_gt_frames_host != nullptr;
    AGS_CHECK_ARG(a->rgb && a->normal && a->depth && a->opacity && a->tanfov, "NULL input");
    AGS_CHECK_ARG(list ? (a->rgb_gt_frames_host && a->depth_gt_frames_host) : (a->rgb_gt && a->depth_gt),
                  "ground truth: pass rgb_gt/depth_gt or both *_frames_host lists");
    AGS_CHECK_ARG(!list || a->B <= AGS_LOSS_MAX_FRAMES, "at most %d frames with per-frame ground-truth pointers", AGS_LOSS_MAX_FRAMES);
    AGS_CHECK_ARG(a->d_rgb && a->d_normal && a->d_depth && a->loss_terms, "NULL output");
    AGS_CHECK_ARG(a->workspace && a->workspace_bytes >= ags_loss_scratch_bytes(a->B, a->H, a->W),
                  "loss workspace too small");
    AGS_CHECK_ARG((unsigned long long)a->B * 4ull * a->H * a->W < 4294967295ull, "B*4*H*W exceeds 32-bit indexing");
    cudaStream_t st = (cudaStream_t)a->stream;
    AGS_CHECK_CUDA(cudaMemsetAsync(a->loss_terms, 0, (4 + 2 * (size_t)a->B) * sizeof(float), st));
    const size_t P = (size_t)a->H * a->W;
    float* msum_plane = (float*)a->workspace;             // (H,W) visibility count (quirk Q1)
    LossFrames fr;
    memset(&fr, 0, sizeof(fr));
    if (list)
        for (int f = 0; f < a->B; ++f) {
            AGS_CHECK_ARG(a->rgb_gt_frames_host[f] && a->depth_gt_frames_host[f], "NULL ground-truth frame %d", f);
            fr.rgb[f] = a->rgb_gt_frames_host[f];
            fr.depth[f] = a->depth_gt_frames_host[f];
        }
    dim3 grid((a->W + LT_W - 1) / LT_W, (a->H + LT_H - 1) / LT_H, a->B), block(LT_W, 8);
    const int vec4 = (P % 4 == 0) && ((((uintptr_t)a->opacity | (uintptr_t)a->vis_count | (uintptr_t)msum_plane) & 15) == 0);
    const size_t vc_threads = vec4 ? P / 4 : P;
    ags_note_launch(); loss_vis_count<<<(unsigned)((vc_threads + 255) / 256), 256, 0, st>>>(*a, msum_plane, vec4);
    AGS_CHECK_CUDA(cudaGetLastError());
    ags_note_launch();
    if (list) loss_fused_kernel<true><<<grid, block, 0, st>>>(*a, fr, msum_plane);
    else loss_fused_kernel<false><<<grid, block, 0, st>>>(*a, fr, msum_plane);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
