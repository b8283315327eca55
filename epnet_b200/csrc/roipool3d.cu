// 3D RoI point pooling for B200 -- first "next" row of SURVEY.md section 8(f).  Replaces roipool3dLauncher
// (/root/reference/lib/utils/roipool3d/src/roipool3d_kernel.cu:207-236): for every box, the first `sampled` points (in index
// order) lying inside the rotated box, cyclically repeated when fewer (:152-158), their coordinates and features copied out;
// an empty box only raises its flag (:149-150).
//
// Reference: three kernels and two cudaMalloc/cudaFree per call (implicit device syncs): a (B,N,M) int mask in global memory,
// then one THREAD per box compacting N mask entries serially, then the copy.  Here: one CTA per (box, scene), no scratch
// memory, no allocation: each of the 4 warps scans a quarter of the cloud 32 points at a time (ballot + popc compaction keeps
// index order), the per-warp hit lists are concatenated in shared memory, and the same CTA copies the selected rows
// (lane-strided: coalesced reads of the point-major (B,N,C) features and coalesced writes).
// The in-box test reproduces the reference's arithmetic as compiled by nvcc -O2 (PTX-checked): cosf/sinf of the box angle,
// x_rot = fma(dx, cosa, -(dz*sina)), z_rot = fma(dz, cosa, dx*sina), |dx|,|dz| pre-test against 10 m.  (x_rot is read off the SASS of
// the reference build: its PTX shows mul, mul, sub without rounding modifiers, which ptxas contracts -- see csrc/iou3d.cu.)
#include "common.cuh"

namespace epnet {

constexpr int kRpWarps = 4;
constexpr int kRpMaxSampled = 512;

__global__ void __launch_bounds__(kRpWarps * 32)
roipool3d_kernel(int n, int m, int c, int sampled, const float *__restrict__ xyz, const float *__restrict__ boxes3d,
                 const float *__restrict__ pts_feature, float *__restrict__ pooled, int *__restrict__ empty_flag)
{
    __shared__ int seg[kRpWarps][kRpMaxSampled];
    __shared__ int seg_cnt[kRpWarps];
    __shared__ int final_idx[kRpMaxSampled];

    const int box = blockIdx.x, scene = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    xyz += (size_t)scene * n * 3;
    pts_feature += (size_t)scene * n * c;
    const float *bp = boxes3d + ((size_t)scene * m + box) * 7;
    const float cx = __ldg(bp), bottom_y = __ldg(bp + 1), cz = __ldg(bp + 2), h = __ldg(bp + 3), w = __ldg(bp + 4), l = __ldg(bp + 5);
    const float angle = __ldg(bp + 6);
    const float half_h = __fmul_rn(h, 0.5f), half_w = __fmul_rn(w, 0.5f), half_l = __fmul_rn(l, 0.5f);
    const float cy = __fsub_rn(bottom_y, half_h);
    const float cosa = cosf(angle), sina = sinf(angle);

    // ---- phase 1: each warp scans its quarter of the cloud in index order
    const int per_warp = ((n + kRpWarps - 1) / kRpWarps + 31) & ~31;
    const int k_begin = warp * per_warp, k_end = min(n, k_begin + per_warp);
    int cnt = 0;
    for (int k0 = k_begin; k0 < k_end && cnt < sampled; k0 += 32) {
        const int k = k0 + lane;
        bool in = false;
        if (k < k_end) {
            const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
            const float dx = __fsub_rn(x, cx), dz = __fsub_rn(z, cz);
            if (!(fabsf(dx) > 10.0f) && !(fabsf(__fsub_rn(y, cy)) > half_h) && !(fabsf(dz) > 10.0f)) {
                const float x_rot = __fmaf_rn(dx, cosa, -__fmul_rn(dz, sina));
                const float z_rot = __fmaf_rn(dz, cosa, __fmul_rn(dx, sina));
                in = (x_rot >= -half_l) & (x_rot <= half_l) & (z_rot >= -half_w) & (z_rot <= half_w);
            }
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, in);
        const int pos = cnt + __popc(mask & lanemask_lt());
        if (in && pos < sampled) seg[warp][pos] = k;
        cnt += __popc(mask);
    }
    if (lane == 0) seg_cnt[warp] = min(cnt, sampled);
    __syncthreads();

    // ---- concatenate the warps' lists in index order, truncate, repeat cyclically
    int off[kRpWarps + 1];
    off[0] = 0;
#pragma unroll
    for (int v = 0; v < kRpWarps; ++v) off[v + 1] = off[v] + seg_cnt[v];
    const int total = min(off[kRpWarps], sampled);
    if (total == 0) {
        if (threadIdx.x == 0) empty_flag[(size_t)scene * m + box] = 1;
        return;
    }
    for (int k = threadIdx.x; k < total; k += kRpWarps * 32) {
        int v = 0;
#pragma unroll
        for (int u = 1; u < kRpWarps; ++u) v += (k >= off[u]);
        final_idx[k] = seg[v][k - off[v]];
    }
    __syncthreads();
    for (int k = total + threadIdx.x; k < sampled; k += kRpWarps * 32) final_idx[k] = final_idx[k % total];
    __syncthreads();

    // ---- phase 2: copy xyz | features of the selected points; one warp per output row, lanes across the row
    const int row_len = 3 + c;
    float *out = pooled + ((size_t)scene * m + box) * sampled * row_len;
    for (int k = warp; k < sampled; k += kRpWarps) {
        const int src = final_idx[k];
        const float *px = xyz + 3 * (size_t)src;
        const float *pf = pts_feature + (size_t)src * c;
        float *o = out + (size_t)k * row_len;
        for (int j = lane; j < row_len; j += 32) o[j] = j < 3 ? __ldg(px + j) : __ldg(pf + (j - 3));
    }
}

}  // namespace epnet

EPNET_API int epnet_roipool3d(int b, int n, int m, int c, int sampled, const float *xyz, const float *boxes3d, const float *pts_feature,
                              float *pooled_features, int *pooled_empty_flag, void *stream)
{
    using namespace epnet;
    if (b < 0 || n < 0 || m < 0 || c < 0 || sampled < 1 || sampled > kRpMaxSampled || !xyz || !boxes3d || !pooled_features || !pooled_empty_flag ||
        (c > 0 && !pts_feature) || b > 65535)
        return EPNET_ERR_BAD_ARG;
    if (b == 0 || m == 0) return EPNET_OK;
    roipool3d_kernel<<<dim3(m, b), kRpWarps * 32, 0, (cudaStream_t)stream>>>(n, m, c, sampled, xyz, boxes3d, pts_feature, pooled_features,
                                                                            pooled_empty_flag);
    EPNET_RETURN_LAUNCH_STATUS();
}
