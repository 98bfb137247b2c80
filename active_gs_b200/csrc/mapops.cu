// mapops.cu -- small device-side pieces of the map's per-keyframe bookkeeping that the reference does
// with many tiny ATen launches and host round trips.
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
    stage_cameras_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table, ids, B, viewmatrix, projmatrix, tanfov);
    AGS_CHECK_CUDA(cudaGetLastError());
    return 0;
}
