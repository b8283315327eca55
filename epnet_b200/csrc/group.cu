// gather_points / group_points (+grads) for B200.  Replace gather_points_kernel_fast,
// gather_points_grad_kernel_fast (/root/reference/pointnet2_lib/pointnet2/src/sampling_gpu.cu:8-24, 46-63)
// and group_points_kernel_fast, group_points_grad_kernel_fast (group_points_gpu.cu:47-66, 8-25).
//
// Both ops are the same indexed row gather: out[b,c,e] = points[b,c,idx[b,e]] with e running over M
// (gather) or M*nsample (group) output elements.  The reference spends one thread, one 4-byte index
// load and one 4-byte store per (c,e).  Here a thread owns 4 consecutive e: the indices are read once
// as one 128-bit load and reused for kRowsPerThread channel rows, each row costing 4 gathers (which hit
// L1/L2: a row is N*4 B <= 64 KB) and one 128-bit fully coalesced store -- the store stream is the
// HBM-bound part and runs at full sector efficiency.
#include "common.cuh"

namespace epnet {

constexpr int kGatherThreads = 128;
constexpr int kRowsPerThread = 8;

__global__ void __launch_bounds__(kGatherThreads)
row_gather_kernel(int c, int n, long long e_total, const float *__restrict__ points, const int *__restrict__ idx,
                  float *__restrict__ out, int vec_ok)
{
    const int scene = blockIdx.z;
    points += (size_t)scene * c * n;
    idx += (size_t)scene * e_total;
    out += (size_t)scene * c * e_total;

    const long long e0 = ((long long)blockIdx.x * kGatherThreads + threadIdx.x) * 4;
    if (e0 >= e_total) return;
    const int c_begin = blockIdx.y * kRowsPerThread;
    const int c_end = min(c, c_begin + kRowsPerThread);

    if (vec_ok) {  // e_total % 4 == 0 and 16-byte aligned bases
        const int4 id = __ldg(reinterpret_cast<const int4 *>(idx + e0));
#pragma unroll 4
        for (int ch = c_begin; ch < c_end; ++ch) {
            const float *row = points + (size_t)ch * n;
            const float4 v = make_float4(__ldg(row + id.x), __ldg(row + id.y), __ldg(row + id.z), __ldg(row + id.w));
            __stcs(reinterpret_cast<float4 *>(out + (size_t)ch * e_total + e0), v);  // streaming: written once, read by the next op
        }
    } else {
        int id[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = (e0 + k < e_total) ? __ldg(idx + e0 + k) : 0;
        for (int ch = c_begin; ch < c_end; ++ch) {
            const float *row = points + (size_t)ch * n;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (e0 + k < e_total) out[(size_t)ch * e_total + e0 + k] = __ldg(row + id[k]);
        }
    }
}

// grad_points[b,c,idx[b,e]] += grad_out[b,c,e]   (fp32 atomics like the reference; order unspecified)
__global__ void __launch_bounds__(kGatherThreads)
row_scatter_add_kernel(int c, int n, long long e_total, const float *__restrict__ grad_out, const int *__restrict__ idx,
                       float *__restrict__ grad_points, int vec_ok)
{
    const int scene = blockIdx.z;
    grad_points += (size_t)scene * c * n;
    idx += (size_t)scene * e_total;
    grad_out += (size_t)scene * c * e_total;

    const long long e0 = ((long long)blockIdx.x * kGatherThreads + threadIdx.x) * 4;
    if (e0 >= e_total) return;
    const int c_begin = blockIdx.y * kRowsPerThread;
    const int c_end = min(c, c_begin + kRowsPerThread);

    int id[4];
    if (vec_ok) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(idx + e0));
        id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = (e0 + k < e_total) ? __ldg(idx + e0 + k) : -1;
    }
    for (int ch = c_begin; ch < c_end; ++ch) {
        float *row = grad_points + (size_t)ch * n;
        float g[4];
        if (vec_ok) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(grad_out + (size_t)ch * e_total + e0));
            g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] = (e0 + k < e_total) ? grad_out[(size_t)ch * e_total + e0 + k] : 0.f;
        }
        // ball-query padding repeats one index many times in a row: fold equal neighbours before the atomic
        float acc = g[0];
        int cur = id[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            if (id[k] == cur) {
                acc += g[k];
            } else {
                if (cur >= 0) atomicAdd(row + cur, acc);
                cur = id[k];
                acc = g[k];
            }
        }
        if (cur >= 0) atomicAdd(row + cur, acc);
    }
}

static int launch_gather(int b, int c, int n, long long e_total, const float *points, const int *idx, float *out, cudaStream_t st)
{
    if (b == 0 || c == 0 || e_total == 0) return EPNET_OK;
    const int turned = launch_transposed_gather(false, staged_row_fits(n), b, c, n, e_total, points, idx, nullptr, out, st);
    if (turned != kStagedNotApplicable) return turned;
    const int staged = launch_staged_rows(false, b, c, n, e_total, points, idx, nullptr, out, st);
    if (staged != kStagedNotApplicable) return staged;
    const uintptr_t al = reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(out);
    const int vec_ok = (e_total % 4 == 0) && ((al & 15) == 0);
    dim3 grid((unsigned)(((e_total + 3) / 4 + kGatherThreads - 1) / kGatherThreads), (c + kRowsPerThread - 1) / kRowsPerThread, b);
    row_gather_kernel<<<grid, kGatherThreads, 0, st>>>(c, n, e_total, points, idx, out, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}

static int launch_scatter(int b, int c, int n, long long e_total, const float *grad_out, const int *idx, float *grad_points,
                          cudaStream_t st)
{
    if (b == 0 || c == 0 || e_total == 0) return EPNET_OK;
    const int turned = launch_transposed_scatter(false, b, c, n, e_total, grad_out, idx, nullptr, grad_points, st);
    if (turned != kStagedNotApplicable) return turned;
    const uintptr_t al = reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(grad_out);
    const int vec_ok = (e_total % 4 == 0) && ((al & 15) == 0);
    dim3 grid((unsigned)(((e_total + 3) / 4 + kGatherThreads - 1) / kGatherThreads), (c + kRowsPerThread - 1) / kRowsPerThread, b);
    row_scatter_add_kernel<<<grid, kGatherThreads, 0, st>>>(c, n, e_total, grad_out, idx, grad_points, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet

EPNET_API int epnet_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, void *stream)
{
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || !points || !idx || !out) return EPNET_ERR_BAD_ARG;
    return epnet::launch_gather(b, c, n, npoints, points, idx, out, (cudaStream_t)stream);
}

EPNET_API int epnet_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points,
                                       void *stream)
{
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || !grad_out || !idx || !grad_points) return EPNET_ERR_BAD_ARG;
    return epnet::launch_scatter(b, c, n, npoints, grad_out, idx, grad_points, (cudaStream_t)stream);
}

EPNET_API int epnet_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out,
                                 void *stream)
{
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0 || !points || !idx || !out) return EPNET_ERR_BAD_ARG;
    return epnet::launch_gather(b, c, n, (long long)npoints * nsample, points, idx, out, (cudaStream_t)stream);
}

EPNET_API int epnet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                                      float *grad_points, void *stream)
{
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0 || !grad_out || !idx || !grad_points) return EPNET_ERR_BAD_ARG;
    return epnet::launch_scatter(b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, (cudaStream_t)stream);
}
