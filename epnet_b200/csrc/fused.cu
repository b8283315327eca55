// Fused data-movement kernels of the set-abstraction / feature-propagation levels (include/epnet_b200.h,
// "fused entry points").  Each one replaces a chain of reference launches + torch glue with a single pass
// that writes the next GEMM's operand directly; all are HBM/L2-bandwidth kernels: 128-bit coalesced
// stores, indices read once per thread and reused across channel rows.
#include "common.cuh"

namespace epnet {

constexpr int kFuThreads = 128;
constexpr int kFuRows = 8;

// ---------------------------------------------------------------------------------------------
// out (B,3+C,M,ns): rows 0..2 = xyz[idx] - new_xyz, rows 3.. = features[:, idx]
// (QueryAndGroup.forward, pointnet2_utils.py:250-257: transpose + 2 x group_points + subtract + cat)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFuThreads)
group_concat_kernel(int c, int n, int m, int ns, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                    const float *__restrict__ features, const int *__restrict__ idx, float *__restrict__ out, int vec_ok)
{
    const int scene = blockIdx.z;
    const long long e_total = (long long)m * ns;
    xyz += (size_t)scene * n * 3;
    new_xyz += (size_t)scene * m * 3;
    if (features) features += (size_t)scene * c * n;
    idx += (size_t)scene * e_total;
    out += (size_t)scene * (3 + c) * e_total;

    const long long e0 = ((long long)blockIdx.x * kFuThreads + threadIdx.x) * 4;
    if (e0 >= e_total) return;
    const int r_begin = blockIdx.y * kFuRows;
    const int r_end = min(3 + c, r_begin + kFuRows);

    int id[4];
    if (vec_ok) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(idx + e0));
        id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = (e0 + k < e_total) ? __ldg(idx + e0 + k) : 0;
    }
    for (int r = r_begin; r < r_end; ++r) {
        float v[4];
        if (r < 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = (int)(min(e0 + k, e_total - 1) / ns);
                v[k] = __fsub_rn(__ldg(xyz + 3 * id[k] + r), __ldg(new_xyz + 3 * p + r));
            }
        } else {
            const float *row = features + (size_t)(r - 3) * n;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __ldg(row + id[k]);
        }
        float *o = out + (size_t)r * e_total + e0;
        if (vec_ok) {
            *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (e0 + k < e_total) o[k] = v[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// x[b,c,:] = max(x[b,c,:] + bias[c], 0), in place, x (B*C, L)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bias_relu_kernel(int c, long long l, float *__restrict__ x, const float *__restrict__ bias, int vec_ok)
{
    const int row = blockIdx.y;  // b*C + c
    const float bv = __ldg(bias + row % c);
    float *p = x + (size_t)row * l;
    if (vec_ok) {
        const long long l4 = l >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < l4; i += (long long)gridDim.x * blockDim.x) {
            float4 v = reinterpret_cast<float4 *>(p)[i];
            v.x = fmaxf(v.x + bv, 0.f); v.y = fmaxf(v.y + bv, 0.f); v.z = fmaxf(v.z + bv, 0.f); v.w = fmaxf(v.w + bv, 0.f);
            reinterpret_cast<float4 *>(p)[i] = v;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < l; i += (long long)gridDim.x * blockDim.x)
            p[i] = fmaxf(p[i] + bv, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// out[b,c,p] = max(max_s x[b,c,p,s] + bias[c], 0)   (== max_s relu(x + bias): rounding and relu are monotone)
// A thread owns one (b,c,p): ns consecutive floats, read as 128-bit loads when ns % 4 == 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bias_relu_maxpool_kernel(int c, int m, int ns, const float *__restrict__ x, const float *__restrict__ bias,
                         float *__restrict__ out, long long out_batch_stride, int vec_ok)
{
    const int scene = blockIdx.z, ch = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const float *src = x + (((size_t)scene * c + ch) * m + p) * ns;
    float mx = -__int_as_float(0x7f800000);
    if (vec_ok) {
        for (int s = 0; s < ns; s += 4) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(src + s));
            mx = fmaxf(fmaxf(fmaxf(mx, v.x), fmaxf(v.y, v.z)), v.w);
        }
    } else {
        for (int s = 0; s < ns; ++s) mx = fmaxf(mx, src[s]);
    }
    out[(size_t)scene * out_batch_stride + (size_t)ch * m + p] = fmaxf(mx + __ldg(bias + ch), 0.f);
}

// ---------------------------------------------------------------------------------------------
// out (B,C2+C1,n): rows 0..C2-1 = three_interpolate(known_feats, idx, w), w from squared distances
// as PointnetFPModule does (pointnet2_utils.py:98 sqrt; pointnet2_modules.py:157-159); rows C2.. = skip.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFuThreads)
three_interpolate_concat_kernel(int c2, int m, int n, int c1, const float *__restrict__ known_feats, const int *__restrict__ idx,
                                const float *__restrict__ dist2, const float *__restrict__ skip, float *__restrict__ out, int vec_ok)
{
    const int scene = blockIdx.z;
    known_feats += (size_t)scene * c2 * m;
    idx += (size_t)scene * n * 3;
    dist2 += (size_t)scene * n * 3;
    if (skip) skip += (size_t)scene * c1 * n;
    out += (size_t)scene * (c2 + c1) * n;

    const int i0 = (blockIdx.x * kFuThreads + threadIdx.x) * 4;
    if (i0 >= n) return;
    const int r_begin = blockIdx.y * kFuRows;
    const int r_end = min(c2 + c1, r_begin + kFuRows);

    int id[12];
    float w[12];
    if (r_begin < c2) {
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const bool ok = 3 * i0 + e < 3 * n;
            id[e] = ok ? __ldg(idx + 3 * i0 + e) : 0;
            w[e] = ok ? __ldg(dist2 + 3 * i0 + e) : 1.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(w[3 * q]), 1e-8f));
            const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(w[3 * q + 1]), 1e-8f));
            const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(w[3 * q + 2]), 1e-8f));
            const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
            w[3 * q] = __fdiv_rn(r0, norm);
            w[3 * q + 1] = __fdiv_rn(r1, norm);
            w[3 * q + 2] = __fdiv_rn(r2, norm);
        }
    }
    for (int r = r_begin; r < r_end; ++r) {
        float v[4];
        if (r < c2) {
            const float *row = known_feats + (size_t)r * m;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                v[q] = __fmaf_rn(w[3 * q + 2], __ldg(row + id[3 * q + 2]),
                                 __fmaf_rn(w[3 * q], __ldg(row + id[3 * q]), __fmul_rn(w[3 * q + 1], __ldg(row + id[3 * q + 1]))));
        } else {
            const float *row = skip + (size_t)(r - c2) * n + i0;
            if (vec_ok) {
                const float4 s = __ldg(reinterpret_cast<const float4 *>(row));
                v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = (i0 + q < n) ? __ldg(row + q) : 0.f;
            }
        }
        float *o = out + (size_t)r * n + i0;
        if (vec_ok) {
            *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < n) o[q] = v[q];
        }
    }
}

}  // namespace epnet

EPNET_API int epnet_group_concat(int b, int c, int n, int m, int nsample, const float *xyz, const float *new_xyz, const float *features,
                                 const int *idx, float *out, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || n < 0 || m < 0 || nsample < 0 || !xyz || !new_xyz || !idx || !out || (c > 0 && !features)) return EPNET_ERR_BAD_ARG;
    const long long e_total = (long long)m * nsample;
    if (b == 0 || e_total == 0) return EPNET_OK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(out);
    const int vec_ok = (e_total % 4 == 0) && (nsample % 4 == 0) && ((al & 15) == 0);
    dim3 grid((unsigned)(((e_total + 3) / 4 + kFuThreads - 1) / kFuThreads), (3 + c + kFuRows - 1) / kFuRows, b);
    group_concat_kernel<<<grid, kFuThreads, 0, (cudaStream_t)stream>>>(c, n, m, nsample, xyz, new_xyz, c ? features : nullptr, idx, out, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_bias_relu(int b, int c, long long l, float *x, const float *bias, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || l < 0 || !x || !bias) return EPNET_ERR_BAD_ARG;
    if (b == 0 || c == 0 || l == 0) return EPNET_OK;
    if ((long long)b * c > 65535) return EPNET_ERR_BAD_ARG;
    const int vec_ok = (l % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const long long work = vec_ok ? l / 4 : l;
    const unsigned gx = (unsigned)min((long long)64, (work + 255) / 256);
    bias_relu_kernel<<<dim3(gx, b * c), 256, 0, (cudaStream_t)stream>>>(c, l, x, bias, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_bias_relu_maxpool(int b, int c, int m, int nsample, const float *x, const float *bias, float *out,
                                      long long out_batch_stride, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || m < 0 || nsample <= 0 || !x || !bias || !out || c > 65535 || b > 65535) return EPNET_ERR_BAD_ARG;
    if (b == 0 || c == 0 || m == 0) return EPNET_OK;
    const int vec_ok = (nsample % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    bias_relu_maxpool_kernel<<<dim3((m + 255) / 256, c, b), 256, 0, (cudaStream_t)stream>>>(c, m, nsample, x, bias, out, out_batch_stride, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_three_interpolate_concat(int b, int c2, int m, int n, int c1, const float *known_feats, const int *idx,
                                             const float *dist2, const float *skip_feats, float *out, void *stream)
{
    using namespace epnet;
    if (b < 0 || c2 < 0 || c1 < 0 || m < 0 || n < 0 || !known_feats || !idx || !dist2 || !out || (c1 > 0 && !skip_feats)) return EPNET_ERR_BAD_ARG;
    if (b == 0 || n == 0 || c1 + c2 == 0) return EPNET_OK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(skip_feats);
    const int vec_ok = (n % 4 == 0) && ((al & 15) == 0);
    dim3 grid(((n + 3) / 4 + kFuThreads - 1) / kFuThreads, (c2 + c1 + kFuRows - 1) / kFuRows, b);
    three_interpolate_concat_kernel<<<grid, kFuThreads, 0, (cudaStream_t)stream>>>(c2, m, n, c1, known_feats, idx, dist2,
                                                                                   c1 ? skip_feats : nullptr, out, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}
