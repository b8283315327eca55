// oracle/ref_shim.cu -- TEST INFRASTRUCTURE.  extern "C" doors onto the reference's own, unmodified
// kernel launchers (declared in /root/reference/pointnet2_lib/pointnet2/src/*_gpu.h), so they can
// be driven through ctypes without torch's C++ API (the reference's .cpp wrappers need THC, which
// torch >= 1.11 no longer ships).  Built only by oracle/build_ref.sh into oracle/_ref/.
#include <cuda_runtime_api.h>
#include "ball_query_gpu.h"
#include "group_points_gpu.h"
#include "interpolate_gpu.h"
#include "sampling_gpu.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void ref_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream)
{ furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, (cudaStream_t)stream); }

REF_API void ref_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out, void *stream)
{ gather_points_kernel_launcher_fast(b, c, n, m, points, idx, out, (cudaStream_t)stream); }

REF_API void ref_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points, void *stream)
{ gather_points_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, grad_points, (cudaStream_t)stream); }

// first pointer = query centres (new_xyz), second = cloud (xyz): ball_query_gpu.cu:48-49 / ball_query.cpp:23
REF_API void ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, void *stream)
{ ball_query_kernel_launcher_fast(b, n, m, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)stream); }

REF_API void ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out, void *stream)
{ group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out, (cudaStream_t)stream); }

REF_API void ref_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points, void *stream)
{ group_points_grad_kernel_launcher_fast(b, c, n, npoints, nsample, grad_out, idx, grad_points, (cudaStream_t)stream); }

REF_API void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *stream)
{ three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx, (cudaStream_t)stream); }

REF_API void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, void *stream)
{ three_interpolate_kernel_launcher_fast(b, c, m, n, points, idx, weight, out, (cudaStream_t)stream); }

REF_API void ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, void *stream)
{ three_interpolate_grad_kernel_launcher_fast(b, c, n, m, grad_out, idx, weight, grad_points, (cudaStream_t)stream); }

// lib/utils/roipool3d/src/roipool3d_kernel.cu:207 (declared in roipool3d.cpp:8-9); runs on the legacy default stream and
// cudaMalloc/cudaFrees its scratch on every call.
void roipool3dLauncher(int batch_size, int pts_num, int boxes_num, int feature_in_len, int sampled_pts_num, const float *xyz,
                       const float *boxes3d, const float *pts_feature, float *pooled_features, int *pooled_empty_flag);
REF_API void ref_roipool3d(int b, int n, int m, int c, int sampled, const float *xyz, const float *boxes3d, const float *pts_feature,
                           float *pooled_features, int *pooled_empty_flag)
{ roipool3dLauncher(b, n, m, c, sampled, xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag); }

// lib/utils/iou3d/src/iou3d_kernel.cu:352-387 (declared in iou3d.cpp:22-25): legacy default stream.  The host half of the
// reference's nms_gpu (mask D2H + greedy loop, iou3d.cpp:95-113) is replayed by oracle/ref_cuda.py on the mask these produce.
void boxesoverlapLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_overlap);
void boxesioubevLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_iou);
void nmsLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void nmsNormalLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
REF_API void ref_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_overlap)
{ boxesoverlapLauncher(num_a, boxes_a, num_b, boxes_b, ans_overlap); }
REF_API void ref_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou)
{ boxesioubevLauncher(num_a, boxes_a, num_b, boxes_b, ans_iou); }
REF_API void ref_nms_mask(const float *boxes, unsigned long long *mask, int boxes_num, float thresh)
{ nmsLauncher(boxes, mask, boxes_num, thresh); }
REF_API void ref_nms_normal_mask(const float *boxes, unsigned long long *mask, int boxes_num, float thresh)
{ nmsNormalLauncher(boxes, mask, boxes_num, thresh); }

// The reference's nms_gpu / nms_normal_gpu wrappers (iou3d.cpp:74-121, :124-170) need torch's C++ API (at::Tensor); their
// bodies are restated here on raw pointers so the reference ARM pays what the reference pays: a cudaMalloc of the
// N x ceil(N/64) mask, the reference's own mask kernel, a blocking D2H copy, the greedy host loop, a cudaFree.
#include <cstring>
#include <vector>
static int ref_nms_host(const float *boxes, long long *keep, int n, float thresh, bool rotated)
{
    const int col_blocks = (n + 63) / 64;
    unsigned long long *mask = nullptr;
    if (cudaMalloc((void **)&mask, sizeof(unsigned long long) * (size_t)n * col_blocks) != cudaSuccess) return -1;
    if (rotated) nmsLauncher(boxes, mask, n, thresh); else nmsNormalLauncher(boxes, mask, n, thresh);
    std::vector<unsigned long long> mask_cpu((size_t)n * col_blocks);
    cudaMemcpy(mask_cpu.data(), mask, sizeof(unsigned long long) * (size_t)n * col_blocks, cudaMemcpyDeviceToHost);
    cudaFree(mask);
    std::vector<unsigned long long> remv(col_blocks, 0ull);
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const int nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & (1ull << inblock))) {
            keep[kept++] = i;
            const unsigned long long *p = mask_cpu.data() + (size_t)i * col_blocks;
            for (int j = nblock; j < col_blocks; ++j) remv[j] |= p[j];
        }
    }
    return kept;
}
REF_API int ref_nms_gpu(const float *boxes, long long *keep, int n, float thresh) { return ref_nms_host(boxes, keep, n, thresh, true); }
REF_API int ref_nms_normal_gpu(const float *boxes, long long *keep, int n, float thresh) { return ref_nms_host(boxes, keep, n, thresh, false); }
