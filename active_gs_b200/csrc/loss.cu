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
// Two stencil passes, one thread per (pixel, frame):
//   pass A: unit normal, d2n, rgb gradient, loss sums, the pixel's own depth gradient (L1 term +
//           its share of the depth2normal adjoint) and, as four planes, what it contributes to the
//           depth gradients of its 4 neighbours
//   pass B: depth gradient += gather of the neighbours' planes; normal gradient (consistency +
//           gather of the TV terms, then through normalize*mask); TV loss sum
// HBM roofline: reads 15 planes + writes 13 planes of B*H*W floats (+3 scratch planes twice).
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
    for (int f = 0; f < a.B; ++f) m += (__ldg(a.opacity + (size_t)f * P + p) > 1e-3f) ? 1.f : 0.f;
    return m;
}

// pre-pass: per-pixel visibility count over the frames of this call (quirk Q1), once per pixel
__global__ void __launch_bounds__(256)
loss_vis_count(AgsLossArgs a, float* __restrict__ msum_plane) {
    const size_t P = (size_t)a.H * a.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    msum_plane[p] = vis_sum(a, P, (int)p);
}

#ifndef AGS_LOSS_MINB
#define AGS_LOSS_MINB 4
#endif

// Addressing: every tensor is indexed as  kernel-parameter pointer + 32-bit unsigned element offset
// (B*3*H*W < 2^32, checked by the launcher), which lets the compiler use the uniform-base +
// 32-bit-offset addressing mode instead of 64-bit pointer arithmetic per access.

// pass A: one thread per (pixel, frame).  Writes normal_unit, d2n, d_rgb, the pixel's own depth
// gradient (L1 term + its share of the depth2normal adjoint) and the four contributions it makes
// to its neighbours' depth gradients as four planes (up, left, bottom, right) that pass B gathers:
// no atomics, deterministic.
__global__ void __launch_bounds__(256, AGS_LOSS_MINB)
loss_pass_a(AgsLossArgs a, float* __restrict__ nb, const float* __restrict__ msum_plane) {
    const int H = a.H, W = a.W;
    const unsigned P = (unsigned)H * (unsigned)W;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const unsigned f = blockIdx.z;
    const bool in = (x < W) && (y < H);
    const unsigned p = (unsigned)y * (unsigned)W + (unsigned)x;
    const float Bt = (float)a.B_total;
    const float inv_rgb = 1.f / (Bt * 3.f * (float)P);
    const float inv_d = 1.f / (Bt * (float)P);
    const float inv_cons = 1.f / (Bt * Bt * (float)P);
    float fr_rgb = 0.f, fr_d = 0.f, acc_cons = 0.f;
    if (in) {
        const unsigned o1 = f * P + p;             // single-channel planes
        const unsigned o3 = f * 3u * P + p;        // three-channel planes
        const unsigned o4 = f * 4u * P + p;        // neighbour planes
        const float msum = __ldg(msum_plane + p);
        const float* opac = a.opacity + f * P;     // frame bases for the stencil helpers
        const float* depth = a.depth + f * P;
        const float A = __ldg(a.opacity + o1);
        const float mvis = (A > 1e-3f) ? 1.f : 0.f;
        const float m2 = (A > 1e-2f) ? 1.f : 0.f;
        // ---- L1 rgb + gradient
#pragma unroll
        for (unsigned c = 0; c < 3; ++c) {
            const float e = (__ldg(a.rgb + o3 + c * P) - __ldg(a.rgb_gt + o3 + c * P)) * mvis;
            fr_rgb += fabsf(e);
            a.d_rgb[o3 + c * P] = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * mvis * inv_rgb;
        }
        // ---- L1 depth + gradient
        const float dg = __ldg(a.depth_gt + o1);
        const float md = (dg > 0.f) ? 1.f : 0.f;
        const float ed = (__ldg(a.depth + o1) - dg) * md;
        fr_d = fabsf(ed);
        float dd_self = a.w_depth * (ed > 0.f ? 1.f : (ed < 0.f ? -1.f : 0.f)) * md * inv_d;
        // ---- unit normal
        const F3 n = f3(__ldg(a.normal + o3), __ldg(a.normal + o3 + P), __ldg(a.normal + o3 + 2u * P));
        const F3 nu = n * (m2 * rsqrtf(fmaxf(dot(n, n), 1e-24f)));
        a.normal_unit[o3] = nu.x; a.normal_unit[o3 + P] = nu.y; a.normal_unit[o3 + 2u * P] = nu.z;
        // ---- depth2normal
        const FrameGeom g = frame_geom(a.tanfov, f, H, W);
        const D2N v = d2n_vectors(depth, opac, g, H, W, y, x);
        const F3 ns = cross(v.pu, v.pl) + cross(v.pr, v.pu) + cross(v.pb, v.pr) + cross(v.pl, v.pb);
        const float insn = rsqrtf(fmaxf(dot(ns, ns), 1e-24f));
        const F3 u = ns * insn;
        const F3 d2n = u * m2;
        a.d2n[o3] = d2n.x; a.d2n[o3 + P] = d2n.y; a.d2n[o3 + 2u * P] = d2n.z;
        // ---- consistency loss; adjoint of the un-normalised d2n vector, pushed to the five depths
        acc_cons = (1.f - dot(nu, d2n)) * msum;
        float c_up = 0.f, c_left = 0.f, c_bottom = 0.f, c_right = 0.f;
        if (m2 > 0.f && msum > 0.f) {
            const float wc = -a.w_cons * msum * inv_cons;                // dL/d(nu . d2n)
            const F3 gd = nu * wc;                                       // dL/d u  (d2n = u*m2, m2 = 1)
            const F3 gq = (gd - u * dot(u, gd)) * insn;
            const F3 dpu = (cross(v.pl, gq) + cross(gq, v.pr)) * v.mu;
            const F3 dpl = (cross(gq, v.pu) + cross(v.pb, gq)) * v.ml;
            const F3 dpb = (cross(v.pr, gq) + cross(gq, v.pl)) * v.mb;
            const F3 dpr = (cross(v.pu, gq) + cross(gq, v.pb)) * v.mr;
            const F3 dpc = (dpu + dpl + dpb + dpr) * (-v.mc);
            const float rx = (x - g.cx) * g.ik00, ry = (y - g.cy) * g.ik11;   // c = depth * (rx, ry, 1)
            dd_self += dpc.x * rx + dpc.y * ry + dpc.z;
            c_up = dpu.x * rx + dpu.y * (ry - g.ik11) + dpu.z;          // zero when v.mu == 0
            c_left = dpl.x * (rx - g.ik00) + dpl.y * ry + dpl.z;
            c_bottom = dpb.x * rx + dpb.y * (ry + g.ik11) + dpb.z;
            c_right = dpr.x * (rx + g.ik00) + dpr.y * ry + dpr.z;
        }
        a.d_depth[o1] = dd_self;
        nb[o4] = c_up; nb[o4 + P] = c_left; nb[o4 + 2u * P] = c_bottom; nb[o4 + 3u * P] = c_right;
    }
    float v5[5] = {fr_rgb * inv_rgb, fr_d * inv_d, acc_cons * inv_cons, fr_rgb / (3.f * (float)P), fr_d / (float)P};
    float* const d5[5] = {a.loss_terms + 0, a.loss_terms + 1, a.loss_terms + 2,
                          a.loss_terms + 4 + 2 * f, a.loss_terms + 4 + 2 * f + 1};
    block_accumulate<5>(v5, d5);
}

// TV helper: value and derivative factor of one one-sided difference
__device__ __forceinline__ void tv_term(F3 np_, F3 nq, float dp, float dq, float md, float inv2s2,
                                        float& val, float& coef) {
    const F3 dl = np_ - nq;
    const float nd = dot(dl, dl);
    const float dd = (dp - dq) * (dp - dq);
    const float gate = (dd <= 1e-4f) ? md : 0.f;
    const float e = __expf(-nd * inv2s2);
    val = gate * e * nd;
    coef = gate * e * (1.f - nd * inv2s2);       // d val / d nd
}

// pass B: one thread per (pixel, frame): depth gradient += the neighbours' planes; normal gradient
// (consistency + TV gather, through normalize*mask) and the TV loss sum.
__global__ void __launch_bounds__(256, AGS_LOSS_MINB)
loss_pass_b(AgsLossArgs a, const float* __restrict__ nb, const float* __restrict__ msum_plane) {
    const int H = a.H, W = a.W;
    const unsigned P = (unsigned)H * (unsigned)W;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const unsigned f = blockIdx.z;
    const bool in = (x < W) && (y < H);
    const unsigned p = (unsigned)y * (unsigned)W + (unsigned)x;
    const float Bt = (float)a.B_total;
    const float inv_cons = 1.f / (Bt * Bt * (float)P);
    const float inv_tv = 1.f / (Bt * 4.f * (float)P);
    const float inv2s2 = 1.f / (2.f * 0.3f * 0.3f);
    float acc_tv = 0.f;
    if (in) {
        const unsigned b1 = f * P, b3 = f * 3u * P, b4 = f * 4u * P;
        const float msum = __ldg(msum_plane + p);
        {   // depth gradient: own term (pass A) + what the four neighbours push to this pixel
            float dd = 0.f;
            if (y < H - 1) dd += __ldg(nb + b4 + p + (unsigned)W);              // "up" plane of the pixel below
            if (x < W - 1) dd += __ldg(nb + b4 + P + p + 1u);                   // "left" plane of the pixel to the right
            if (y > 0) dd += __ldg(nb + b4 + 2u * P + p - (unsigned)W);         // "bottom" plane of the pixel above
            if (x > 0) dd += __ldg(nb + b4 + 3u * P + p - 1u);                  // "right" plane of the pixel to the left
            a.d_depth[b1 + p] += dd;
        }
        auto NU = [&](unsigned q) { return f3(__ldg(a.normal_unit + b3 + q), __ldg(a.normal_unit + b3 + P + q),
                                              __ldg(a.normal_unit + b3 + 2u * P + q)); };
        // all loads of the 5-point stencil up front (clamped coordinates, validity folded into flags)
        const unsigned qs[4] = {(unsigned)(y * W + min(x + 1, W - 1)), (unsigned)(y * W + max(x - 1, 0)),
                                (unsigned)(min(y + 1, H - 1) * W + x), (unsigned)(max(y - 1, 0) * W + x)};
        const float ok[4] = {x < W - 1 ? 1.f : 0.f, x > 0 ? 1.f : 0.f, y < H - 1 ? 1.f : 0.f, y > 0 ? 1.f : 0.f};
        const F3 nu = NU(p);
        F3 nq[4];
        float dq[4], mdq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            nq[k] = NU(qs[k]);
            dq[k] = __ldg(a.depth + b1 + qs[k]);
            mdq[k] = (__ldg(a.depth_gt + b1 + qs[k]) > 0.f) ? ok[k] : 0.f;
        }
        const float dp = __ldg(a.depth + b1 + p);
        const float md_p = (__ldg(a.depth_gt + b1 + p) > 0.f) ? 1.f : 0.f;
        const float m2 = (__ldg(a.opacity + b1 + p) > 1e-2f) ? 1.f : 0.f;
        F3 gnu = f3(__ldg(a.d2n + b3 + p), __ldg(a.d2n + b3 + P + p), __ldg(a.d2n + b3 + 2u * P + p))
                 * (-a.w_cons * msum * inv_cons);
        const F3 n = f3(__ldg(a.normal + b3 + p), __ldg(a.normal + b3 + P + p), __ldg(a.normal + b3 + 2u * P + p));
        const float ctv = a.w_tv * inv_tv;
        // each neighbour q contributes the own one-sided difference (mask of p) and the mirrored
        // difference of q that references p (mask of q)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float val, coef;
            tv_term(nu, nq[k], dp, dq[k], md_p * ok[k], inv2s2, val, coef);
            acc_tv += val;
            gnu = gnu + (nu - nq[k]) * (2.f * coef * ctv);
            tv_term(nq[k], nu, dq[k], dp, mdq[k], inv2s2, val, coef);
            gnu = gnu - (nq[k] - nu) * (2.f * coef * ctv);
        }
        const float inn = rsqrtf(fmaxf(dot(n, n), 1e-24f));
        const F3 uh = n * inn;
        const F3 gn = (gnu - uh * dot(uh, gnu)) * (m2 * inn);
        a.d_normal[b3 + p] = gn.x; a.d_normal[b3 + P + p] = gn.y; a.d_normal[b3 + 2u * P + p] = gn.z;
    }
    float v1[1] = {acc_tv * inv_tv};
    float* const d1[1] = {a.loss_terms + 3};
    block_accumulate<1>(v1, d1);
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
    return ags_align256(((size_t)B * 4 + 1) * H * W * sizeof(float));
}

extern "C" int ags_loss_forward_backward(const AgsLossArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0 && a->B_total >= a->B, "bad sizes B=%d H=%d W=%d B_total=%d",
                  a->B, a->H, a->W, a->B_total);
    AGS_CHECK_ARG(a->rgb && a->normal && a->depth && a->opacity && a->rgb_gt && a->depth_gt && a->tanfov,
                  "NULL input");
    AGS_CHECK_ARG(a->normal_unit && a->d2n && a->d_rgb && a->d_normal && a->d_depth && a->loss_terms,
                  "NULL output");
    AGS_CHECK_ARG(a->workspace && a->workspace_bytes >= ags_loss_scratch_bytes(a->B, a->H, a->W),
                  "loss workspace too small");
    AGS_CHECK_ARG((unsigned long long)a->B * 4ull * a->H * a->W < 4294967295ull, "B*4*H*W exceeds 32-bit indexing");
    cudaStream_t st = (cudaStream_t)a->stream;
    AGS_CHECK_CUDA(cudaMemsetAsync(a->loss_terms, 0, (4 + 2 * (size_t)a->B) * sizeof(float), st));
    const size_t P = (size_t)a->H * a->W;
    float* nb = (float*)a->workspace;                     // (B,4,H,W) neighbour contributions
    float* msum_plane = nb + (size_t)a->B * 4 * P;        // (H,W) visibility count (quirk Q1)
    dim3 grid((a->W + 31) / 32, (a->H + 7) / 8, a->B), block(32, 8);
    ags_note_launch(); loss_vis_count<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(*a, msum_plane);
    AGS_CHECK_CUDA(cudaGetLastError());
    ags_note_launch(); loss_pass_a<<<grid, block, 0, st>>>(*a, nb, msum_plane);
    AGS_CHECK_CUDA(cudaGetLastError());
    ags_note_launch(); loss_pass_b<<<grid, block, 0, st>>>(*a, nb, msum_plane);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
