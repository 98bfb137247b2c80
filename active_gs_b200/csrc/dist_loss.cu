// dist_loss.cu -- the two small exchanges of the frame-sharded training iteration, done over NVLink
// peer memory by our own kernels instead of NCCL collectives (the reference is
// single-GPU):
//   * quirk Q1 couples the frames: the consistency term of every frame is weighted by the number of
//     frames of the WHOLE batch in which the pixel is visible (mapping/gaussian_map.py:116-117).
//     Each rank counts its own frames into a symmetric (H*W) int32 plane (ags_dist_vis_local), a
//     cross-GPU barrier follows, and ags_dist_vis_sum adds the planes of all ranks -- in the switch
//     with multimem.ld_reduce when NVLS is available, else with one load per peer.
//   * the sampler on every rank needs the loss terms / per-frame performance of ALL frames
//     (mapping/utils.py:206-218): ags_dist_terms_put stores this rank's terms (+ per-view instance counts, instance total
//     and overflow flag) into slot `rank` of every peer's gather buffer (multimem.st or peer stores).
// Buffers are symmetric allocations; the caller passes the peer pointers and brackets the kernels
// with cross-GPU barriers on the same stream.
#include "ags_common.cuh"

namespace {

struct VisParams {
    int world, B;
    unsigned P;
    const float* opacity;
    const float* frame_weight;
    int32_t* vis_local;
    const int32_t* peers[AGS_MAX_PEERS];
    const int32_t* mc;
    int32_t* out;
    SyncP sync;
};

__global__ void __launch_bounds__(256)
dist_vis_local_kernel(VisParams a, int vec4) {
    if (vec4) {      // four pixels per thread, 128-bit loads: B independent loads in flight per thread
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; 4u * q < a.P; q += gridDim.x * blockDim.x) {
            const unsigned p = 4u * q;
            int4 c = make_int4(0, 0, 0, 0);
            for (int f = 0; f < a.B; ++f) {
                const bool real = !a.frame_weight || __ldg(a.frame_weight + f) != 0.f;  // padded frames do not count
                const float4 o = __ldg(reinterpret_cast<const float4*>(a.opacity + (size_t)f * a.P + p));
                if (real) { c.x += o.x > 1e-3f; c.y += o.y > 1e-3f; c.z += o.z > 1e-3f; c.w += o.w > 1e-3f; }
            }
            *reinterpret_cast<int4*>(a.vis_local + p) = c;
        }
    } else
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += gridDim.x * blockDim.x) {
        int c = 0;
        for (int f = 0; f < a.B; ++f) {
            const bool real = !a.frame_weight || __ldg(a.frame_weight + f) != 0.f;      // padded frames do not count
            c += (real && __ldg(a.opacity + (size_t)f * a.P + p) > 1e-3f) ? 1 : 0;
        }
        a.vis_local[p] = c;
    }
    // the plane is complete when the last block is: tell every rank (folded barrier, AGS_SYNC_VIS)
    if (a.sync.on && sync_last_block(a.sync, AGS_SYNC_VIS)) sync_signal(a.sync, AGS_SYNC_VIS);
}

__global__ void __launch_bounds__(256)
dist_vis_sum_kernel(VisParams a) {
    sync_wait(a.sync, AGS_SYNC_VIS);                 // every rank's plane is complete
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += gridDim.x * blockDim.x) {
        int s;
        if (a.mc) {
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.s32 %0, [%1];" : "=r"(s) : "l"(a.mc + p) : "memory");
        } else {
            s = 0;
            for (int r = 0; r < a.world; ++r) {
                int v;
                asm volatile("ld.global.relaxed.sys.s32 %0, [%1];" : "=r"(v) : "l"(a.peers[r] + p) : "memory");
                s += v;
            }
        }
        a.out[p] = s;
    }
}

struct TermsParams {
    int world, rank, nterm, nview;
    const float* terms;
    const int32_t* stats;
    float* peers[AGS_MAX_PEERS];
    float* mc;
    SyncP sync;
};

__global__ void dist_terms_put_kernel(TermsParams a) {
    const int k = threadIdx.x;
    if (k < a.nterm) {
        const int nt = a.nterm - a.nview - 2;            // layout: terms | per-view instances | instances, overflow
        const float v = (k < nt) ? a.terms[k]
                      : (k < nt + a.nview) ? (float)a.stats[AGS_STAT_VIEW0 + (k - nt)]
                      : (float)a.stats[k - (nt + a.nview)];
        const size_t slot = (size_t)a.rank * a.nterm + k;
        if (a.mc) {
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(a.mc + slot), "f"(v) : "memory");
        } else {
            for (int r = 0; r < a.world; ++r)
                asm volatile("st.global.relaxed.sys.f32 [%0], %1;" :: "l"(a.peers[r] + slot), "f"(v) : "memory");
        }
    }
    __threadfence_system();
    sync_signal(a.sync, AGS_SYNC_TERMS);             // single block: its stores are out
}

__global__ void dist_wait_kernel(SyncP s, int phase) { sync_wait(s, phase); }

int fill_vis(const AgsDistVisArgs* a, VisParams& P, bool need_peers) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->world >= 1 && a->world <= AGS_MAX_PEERS, "bad world %d", a->world);
    AGS_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0 && (long long)a->H * a->W < (1ll << 31), "bad sizes");
    AGS_CHECK_ARG(a->vis_local != nullptr, "NULL vis_local");
    P.world = a->world; P.B = a->B; P.P = (unsigned)a->H * (unsigned)a->W;
    P.sync = make_sync(a->sync, a->world, a->rank);
    P.opacity = a->opacity; P.frame_weight = a->frame_weight; P.vis_local = a->vis_local; P.mc = a->vis_multicast; P.out = a->vis_count;
    for (int r = 0; r < AGS_MAX_PEERS; ++r) {
        P.peers[r] = r < a->world ? a->vis_peers[r] : nullptr;
        if (need_peers && !a->vis_multicast && r < a->world) AGS_CHECK_ARG(a->vis_peers[r] != nullptr, "NULL peer pointer %d", r);
    }
    return 0;
}

}  // namespace

extern "C" int ags_dist_vis_local(const AgsDistVisArgs* a) {
    VisParams P;
    int rc = fill_vis(a, P, false);
    if (rc) return rc;
    AGS_CHECK_ARG(a->opacity != nullptr, "NULL opacity");
    const int vec4 = (P.P % 4u == 0) && ((((uintptr_t)a->opacity | (uintptr_t)a->vis_local) & 15) == 0);
    unsigned blocks = ((vec4 ? P.P / 4 : P.P) + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;             // grid-stride: every block pays a fence + a counter atomic
    ags_note_launch(); dist_vis_local_kernel<<<blocks, 256, 0, (cudaStream_t)a->stream>>>(P, vec4);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ags_dist_vis_sum(const AgsDistVisArgs* a) {
    VisParams P;
    int rc = fill_vis(a, P, true);
    if (rc) return rc;
    AGS_CHECK_ARG(a->vis_count != nullptr, "NULL vis_count");
    unsigned blocks = (P.P + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    ags_note_launch(); dist_vis_sum_kernel<<<blocks, 256, 0, (cudaStream_t)a->stream>>>(P);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ags_dist_terms_put(const AgsDistTermsArgs* a) {
    AGS_CHECK_ARG(a != nullptr, "args is NULL");
    AGS_CHECK_ARG(a->world >= 1 && a->world <= AGS_MAX_PEERS && a->rank >= 0 && a->rank < a->world,
                  "bad world/rank %d/%d", a->world, a->rank);
    AGS_CHECK_ARG(a->nview >= 0 && a->nview <= AGS_NUM_STATS - AGS_STAT_VIEW0, "bad nview %d", a->nview);
    AGS_CHECK_ARG(a->nterm > a->nview + 2 && a->nterm <= 1024, "nterm %d outside %d..1024", a->nterm, a->nview + 3);
    AGS_CHECK_ARG(a->terms && a->stats, "NULL terms / stats");
    TermsParams P;
    P.world = a->world; P.rank = a->rank; P.nterm = a->nterm; P.nview = a->nview; P.terms = a->terms; P.stats = a->stats;
    P.mc = a->gather_multicast;
    P.sync = make_sync(a->sync, a->world, a->rank);
    for (int r = 0; r < AGS_MAX_PEERS; ++r) {
        P.peers[r] = r < a->world ? a->gather_peers[r] : nullptr;
        if (!a->gather_multicast && r < a->world) AGS_CHECK_ARG(a->gather_peers[r] != nullptr, "NULL peer pointer %d", r);
    }
    ags_note_launch(); dist_terms_put_kernel<<<1, ((a->nterm + 31) / 32) * 32, 0, (cudaStream_t)a->stream>>>(P);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ags_dist_wait(const AgsDistSync* sync, int32_t phase, int32_t world, int32_t rank, void* stream) {
    AGS_CHECK_ARG(sync != nullptr && sync->peers[0] != nullptr, "NULL sync");
    AGS_CHECK_ARG(world >= 1 && world <= AGS_MAX_PEERS && rank >= 0 && rank < world, "bad world/rank %d/%d", world, rank);
    AGS_CHECK_ARG(phase >= 0 && phase < 4, "bad phase %d", phase);
    for (int r = 0; r < world; ++r) AGS_CHECK_ARG(sync->peers[r] != nullptr, "NULL sync pointer %d", r);
    ags_note_launch(); dist_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(make_sync(*sync, world, rank), phase);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
