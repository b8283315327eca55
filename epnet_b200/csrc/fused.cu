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

// =================================================================================================
// Point-major variants: rows are points (or (centre,sample) pairs), channels are contiguous.  This is the
// layout the tcgen05 GEMM (gemm_tf32x3.cu) consumes as its K-major A operand, and it turns every indexed
// gather into a contiguous C*4-byte row read (full 32-byte sectors) instead of C separate 4-byte reads
// from a channel-major (B,C,N) tensor -- the reference layout's gathers are L2-sector-bound at 1/8 efficiency.
// =================================================================================================
namespace epnet {

// Thread mapping shared by the point-major kernels: a group of G = 2^g lanes (G >= chunks per row, capped at 32) owns one
// output row and walks its 16-byte chunks; a warp therefore covers 32/G rows, and every gathered row is read as one
// contiguous run by adjacent lanes.  Row-level quantities (index, weights) are computed once per lane, no 64-bit divides.
struct PmMap {
    long long row;   // output row handled by this lane
    int chunk0;      // first chunk of this lane
    int step;        // chunk stride (= G)
};
__device__ __forceinline__ PmMap pm_map(int group_shift)
{
    const int G = 1 << group_shift;
    const long long lane_global = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PmMap m;
    m.row = lane_global >> group_shift;
    m.chunk0 = (int)(lane_global & (G - 1));
    m.step = G;
    return m;
}

// out[(b,p,s)][0..C) = feats[b, idx[b,p,s], 0..C) ; out[..][C..C+3) = xyz[b, idx] - new_xyz[b,p]
// A lane owns one 16-byte chunk position of FOUR consecutive rows (a "quad": same centre when nsample % 4 == 0): one 128-bit index
// load, one division, four independent row gathers in flight.
__global__ void __launch_bounds__(256)
group_concat_pm_kernel(int c, int n, int m, int ns, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                       const float *__restrict__ feats, int ldf, const int *__restrict__ idx, float *__restrict__ out, int ldo,
                       int quads_total, int chunks, int vec_ok, int group_shift)
{
    const unsigned lane_global = blockIdx.x * blockDim.x + threadIdx.x;
    const int quad = (int)(lane_global >> group_shift);
    if (quad >= quads_total) return;
    const int chunk0 = (int)(lane_global & ((1u << group_shift) - 1u)), step = 1 << group_shift;
    const int row0 = quad * 4;
    const int centre = row0 / ns;  // b*m + p, shared by the quad
    const int scene = centre / m;
    const int4 src = __ldg(reinterpret_cast<const int4 *>(idx + row0));
    const int srcs[4] = {src.x, src.y, src.z, src.w};
    const float *fbase = feats + (size_t)scene * n * ldf;
    const float *pbase = xyz + (size_t)scene * n * 3;
    const float *crow = new_xyz + (size_t)centre * 3;
    float *orow = out + (size_t)row0 * ldo;
    int chunks_here = chunks;
    if (vec_ok && (c & 3) == 0 && step >= 4) {
        // the last chunk holds only the three re-centred coordinates: lanes 0..3 of the group write it for rows 0..3 of the quad in the same
        // pass, instead of one lane walking it in a second pass of the whole group (33 chunks over 32 lanes at C = 128)
        chunks_here = c >> 2;
        if (chunk0 < 4) {
            const float *p = pbase + (size_t)srcs[chunk0] * 3;
            const float4 t = make_float4(__fsub_rn(__ldg(p), __ldg(crow)), __fsub_rn(__ldg(p + 1), __ldg(crow + 1)),
                                         __fsub_rn(__ldg(p + 2), __ldg(crow + 2)), 0.f);
            __stcs(reinterpret_cast<float4 *>(orow + (size_t)chunk0 * ldo + c), t);
        }
    }
    for (int ch = chunk0; ch < chunks_here; ch += step) {
        const int k0 = ch * 4;
        float4 v[4];
        if (vec_ok && k0 + 4 <= c) {
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = __ldg(reinterpret_cast<const float4 *>(fbase + (size_t)srcs[r] * ldf + k0));
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + j;
                    if (k < c) e[j] = __ldg(fbase + (size_t)srcs[r] * ldf + k);
                    else if (k < c + 3) e[j] = __fsub_rn(__ldg(pbase + (size_t)srcs[r] * 3 + (k - c)), __ldg(crow + (k - c)));
                    else e[j] = 0.f;
                }
                v[r] = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) __stcs(reinterpret_cast<float4 *>(orow + (size_t)r * ldo + k0), v[r]);  // streamed: read once by the GEMM
    }
}

// out[(b,i)][0..C2) = sum_k w_k * known[b, idx[b,i,k], :] ; out[..][C2..C2+C1) = skip[b,i,:]
// `weight` = the normalised inverse-distance weights (epnet_three_nn_weights) when from_dist2 == 0, else squared distances
__global__ void __launch_bounds__(256)
three_interpolate_concat_pm_kernel(int c2, int m, int n, int c1, const float *__restrict__ known, int ldk, const int *__restrict__ idx,
                                   const float *__restrict__ weight, int from_dist2, const float *__restrict__ skip, int lds,
                                   float *__restrict__ out, int ldo, long long rows_total, int chunks, int group_shift)
{
    const PmMap map = pm_map(group_shift);
    if (map.row >= rows_total) return;
    const long long row = map.row;
    const int scene = (int)(row / n);
    const int i0 = __ldg(idx + row * 3), i1 = __ldg(idx + row * 3 + 1), i2 = __ldg(idx + row * 3 + 2);
    float w0 = __ldg(weight + row * 3), w1 = __ldg(weight + row * 3 + 1), w2 = __ldg(weight + row * 3 + 2);
    if (from_dist2) {
        const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(w0), 1e-8f));
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(w1), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(w2), 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        w0 = __fdiv_rn(r0, norm); w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm);
    }
    const float *kb = known + (size_t)scene * m * ldk;
    const float *ka = kb + (size_t)i0 * ldk, *kbb = kb + (size_t)i1 * ldk, *kc = kb + (size_t)i2 * ldk;
    float *orow = out + (size_t)row * ldo;
    for (int ch = map.chunk0; ch < chunks; ch += map.step) {
        const int k0 = ch * 4;
        float4 o;
        if (k0 < c2) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(ka + k0));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(kbb + k0));
            const float4 c = __ldg(reinterpret_cast<const float4 *>(kc + k0));
            o.x = __fmaf_rn(w2, c.x, __fmaf_rn(w0, a.x, __fmul_rn(w1, b.x)));
            o.y = __fmaf_rn(w2, c.y, __fmaf_rn(w0, a.y, __fmul_rn(w1, b.y)));
            o.z = __fmaf_rn(w2, c.z, __fmaf_rn(w0, a.z, __fmul_rn(w1, b.z)));
            o.w = __fmaf_rn(w2, c.w, __fmaf_rn(w0, a.w, __fmul_rn(w1, b.w)));
        } else {
            o = __ldg(reinterpret_cast<const float4 *>(skip + (size_t)row * lds + (k0 - c2)));
        }
        __stcs(reinterpret_cast<float4 *>(orow + k0), o);
    }
}

// out[(b,i)][col_off + ch] = bilinear(fmap[b,ch], xy[b,i]); one thread per (point, 4 channels)
__device__ __forceinline__ void pm_taps(float gx, float gy, int h, int w, int align_corners, int *o, float *wt)
{
    float ix, iy;
    if (align_corners) {
        ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w - 1));
        iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h - 1));
    } else {
        ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), (float)w, -1.f), 0.5f);
        iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), (float)h, -1.f), 0.5f);
    }
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = __fsub_rn(ix, fx), wy1 = __fsub_rn(iy, fy);
    const float wx0 = __fsub_rn((float)x1, ix), wy0 = __fsub_rn((float)y1, iy);
    const bool xin0 = x0 >= 0 && x0 < w, xin1 = x1 >= 0 && x1 < w, yin0 = y0 >= 0 && y0 < h, yin1 = y1 >= 0 && y1 < h;
    wt[0] = (xin0 && yin0) ? __fmul_rn(wx0, wy0) : 0.f; o[0] = (xin0 && yin0) ? y0 * w + x0 : 0;
    wt[1] = (xin1 && yin0) ? __fmul_rn(wx1, wy0) : 0.f; o[1] = (xin1 && yin0) ? y0 * w + x1 : 0;
    wt[2] = (xin0 && yin1) ? __fmul_rn(wx0, wy1) : 0.f; o[2] = (xin0 && yin1) ? y1 * w + x0 : 0;
    wt[3] = (xin1 && yin1) ? __fmul_rn(wx1, wy1) : 0.f; o[3] = (xin1 && yin1) ? y1 * w + x1 : 0;
}

__global__ void __launch_bounds__(256)
grid_gather_pm_kernel(int c, int h, int w, int n, const float *__restrict__ fmap, const float *__restrict__ xy, int align_corners,
                      float *__restrict__ out, int ldo, long long rows_total, int chunks, int group_shift)
{
    const PmMap map = pm_map(group_shift);
    if (map.row >= rows_total) return;
    const long long row = map.row;
    const int scene = (int)(row / n);
    const float2 g = __ldg(reinterpret_cast<const float2 *>(xy + row * 2));
    int o[4];
    float wt[4];
    pm_taps(g.x, g.y, h, w, align_corners, o, wt);
    const size_t plane = (size_t)h * w;
    for (int ch = map.chunk0; ch < chunks; ch += map.step) {
        const int ch0 = ch * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int chn = ch0 + j;
            if (chn < c) {
                const float *p = fmap + ((size_t)scene * c + chn) * plane;
                float acc = __fmul_rn(__ldg(p + o[0]), wt[0]);
                acc = __fmaf_rn(__ldg(p + o[1]), wt[1], acc);
                acc = __fmaf_rn(__ldg(p + o[2]), wt[2], acc);
                acc = __fmaf_rn(__ldg(p + o[3]), wt[3], acc);
                v[j] = acc;
            } else {
                v[j] = 0.f;
            }
        }
        *reinterpret_cast<float4 *>(out + (size_t)row * ldo + ch0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

static inline int pm_group_shift(int chunks)
{
    int g = 0;
    while ((1 << g) < chunks && g < 5) ++g;
    return g;
}

}  // namespace epnet

EPNET_API int epnet_group_concat_pm(int b, int c, int n, int m, int nsample, const float *xyz, const float *new_xyz, const float *feats,
                                    int ldf, const int *idx, float *out, int ldo, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || n < 0 || m < 0 || nsample < 0 || !xyz || !new_xyz || !idx || !out || (c > 0 && !feats)) return EPNET_ERR_BAD_ARG;
    if (ldo < c + 3 || (ldo & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) return EPNET_ERR_BAD_ARG;
    if ((nsample & 3) || (reinterpret_cast<uintptr_t>(idx) & 15)) return EPNET_ERR_BAD_ARG;  // quads of 4 samples share a centre
    const long long rows = (long long)b * m * nsample;
    if (rows == 0) return EPNET_OK;
    if (rows >= (1ll << 31)) return EPNET_ERR_BAD_ARG;
    const int chunks = (c + 3 + 3) / 4;
    const int vec_ok = c > 0 && (ldf % 4 == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15) == 0);
    const int gs = pm_group_shift(chunks);
    const long long threads = (rows / 4) << gs;
    group_concat_pm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, n, m, nsample, xyz, new_xyz, feats, ldf, idx,
                                                                                            out, ldo, (int)(rows / 4), chunks, vec_ok, gs);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_three_interpolate_concat_pm(int b, int c2, int m, int n, int c1, const float *known, int ldk, const int *idx,
                                                const float *weight, int from_dist2, const float *skip, int lds, float *out, int ldo,
                                                void *stream)
{
    using namespace epnet;
    if (b < 0 || c2 <= 0 || c1 < 0 || m < 0 || n < 0 || !known || !idx || !weight || !out || (c1 > 0 && !skip)) return EPNET_ERR_BAD_ARG;
    if ((c2 & 3) || (c1 & 3) || (ldk & 3) || (ldo & 3) || (c1 > 0 && (lds & 3)) || ldo < c1 + c2) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(known) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(skip)) & 15) return EPNET_ERR_BAD_ARG;
    const long long rows = (long long)b * n;
    if (rows == 0) return EPNET_OK;
    const int chunks = (c1 + c2) / 4;
    const int gs = pm_group_shift(chunks);
    const long long threads = rows << gs;
    three_interpolate_concat_pm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        c2, m, n, c1, known, ldk, idx, weight, from_dist2, skip, lds, out, ldo, rows, chunks, gs);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_grid_gather_pm(int b, int c, int h, int w, int n, const float *fmap, const float *xy, int align_corners, float *out,
                                   int ldo, void *stream)
{
    using namespace epnet;
    if (b < 0 || c <= 0 || h <= 0 || w <= 0 || n < 0 || !fmap || !xy || !out) return EPNET_ERR_BAD_ARG;
    if ((ldo & 3) || ldo < c || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(xy) & 7)) return EPNET_ERR_BAD_ARG;
    const long long rows = (long long)b * n;
    if (rows == 0) return EPNET_OK;
    const int chunks = (c + 3) / 4;
    const int gs = pm_group_shift(chunks);
    const long long threads = rows << gs;
    grid_gather_pm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, xy, align_corners, out, ldo,
                                                                                           rows, chunks, gs);
    EPNET_RETURN_LAUNCH_STATUS();
}

// =================================================================================================
// NHWC image-stream helpers (the image stream runs on the tcgen05 implicit-GEMM convolution, whose activations are
// NHWC = "pixel-major" rows)
// =================================================================================================
namespace epnet {

// LI-Fusion gather from an NHWC feature map: out[(b,i)][0..C) = bilinear(fmap[b,:,:,0..C), xy[b,i]).  Each tap is one
// contiguous C*4-byte run; lane groups walk it in 16-byte chunks (same mapping as the other point-major kernels).
__global__ void __launch_bounds__(256)
grid_gather_nhwc_pm_kernel(int c, int h, int w, int n, const float *__restrict__ fmap, int ldc, const float *__restrict__ xy,
                           int align_corners, float *__restrict__ out, int ldo, long long rows_total, int chunks, int group_shift)
{
    const PmMap map = pm_map(group_shift);
    if (map.row >= rows_total) return;
    const long long row = map.row;
    const int scene = (int)(row / n);
    const float2 g = __ldg(reinterpret_cast<const float2 *>(xy + row * 2));
    int o[4];
    float wt[4];
    pm_taps(g.x, g.y, h, w, align_corners, o, wt);
    const float *base = fmap + (size_t)scene * h * w * ldc;
    const float *t0 = base + (size_t)o[0] * ldc, *t1 = base + (size_t)o[1] * ldc, *t2 = base + (size_t)o[2] * ldc, *t3 = base + (size_t)o[3] * ldc;
    for (int ch = map.chunk0; ch < chunks; ch += map.step) {
        const int k0 = ch * 4;
        const float4 a = __ldg(reinterpret_cast<const float4 *>(t0 + k0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(t1 + k0));
        const float4 cc = __ldg(reinterpret_cast<const float4 *>(t2 + k0));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(t3 + k0));
        float4 r;
        r.x = __fmaf_rn(d.x, wt[3], __fmaf_rn(cc.x, wt[2], __fmaf_rn(b.x, wt[1], __fmul_rn(a.x, wt[0]))));
        r.y = __fmaf_rn(d.y, wt[3], __fmaf_rn(cc.y, wt[2], __fmaf_rn(b.y, wt[1], __fmul_rn(a.y, wt[0]))));
        r.z = __fmaf_rn(d.z, wt[3], __fmaf_rn(cc.z, wt[2], __fmaf_rn(b.z, wt[1], __fmul_rn(a.z, wt[0]))));
        r.w = __fmaf_rn(d.w, wt[3], __fmaf_rn(cc.w, wt[2], __fmaf_rn(b.w, wt[1], __fmul_rn(a.w, wt[0]))));
        *reinterpret_cast<float4 *>(out + (size_t)row * ldo + k0) = r;
    }
}

// ConvTranspose2d with kernel == stride (non-overlapping, lib/net/pointnet2_msg.py:163-165) computed as a GEMM leaves
// y[(b,y,x)][(ky*k + kx)*co + o]; this scatters it to the full-resolution NHWC concat:
// out[(b, y*k+ky, x*k+kx)][col_off + o].  One thread per (output pixel, 16-byte chunk).
__global__ void __launch_bounds__(256)
deconv_shuffle_nhwc_kernel(int h, int w, int k, int co, const float *__restrict__ y, float *__restrict__ out, int ldo, int col_off,
                           long long total)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int chunks = co >> 2;
    const int ch = (int)(t % chunks);
    const long long opix = t / chunks;           // (b, Y, X) over the full-resolution grid
    const int W = w * k, H = h * k;
    const int X = (int)(opix % W);
    const long long rest = opix / W;
    const int Y = (int)(rest % H);
    const int b = (int)(rest / H);
    const int yy = Y / k, ky = Y - yy * k, xx = X / k, kx = X - xx * k;
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(y + (((size_t)b * h + yy) * w + xx) * (size_t)(k * k * co) + (size_t)(ky * k + kx) * co + ch * 4));
    *reinterpret_cast<float4 *>(out + (size_t)opix * ldo + col_off + ch * 4) = v;
}

}  // namespace epnet

EPNET_API int epnet_grid_gather_nhwc_pm(int b, int c, int h, int w, int n, const float *fmap, int ldc, const float *xy, int align_corners,
                                        float *out, int ldo, void *stream)
{
    using namespace epnet;
    if (b < 0 || c <= 0 || (c & 3) || h <= 0 || w <= 0 || n < 0 || !fmap || !xy || !out || ldc < c || (ldc & 3) || ldo < c || (ldo & 3))
        return EPNET_ERR_BAD_ARG;
    if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(fmap)) & 15) || (reinterpret_cast<uintptr_t>(xy) & 7)) return EPNET_ERR_BAD_ARG;
    const long long rows = (long long)b * n;
    if (rows == 0) return EPNET_OK;
    const int chunks = c / 4;
    const int gs = pm_group_shift(chunks);
    const long long threads = rows << gs;
    grid_gather_nhwc_pm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, ldc, xy, align_corners, out,
                                                                                                ldo, rows, chunks, gs);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_deconv_shuffle_nhwc(int b, int h, int w, int k, int co, const float *y, float *out, int ldo, int col_off, void *stream)
{
    using namespace epnet;
    if (b < 0 || h <= 0 || w <= 0 || k <= 0 || co <= 0 || (co & 3) || !y || !out || (ldo & 3) || (col_off & 3) || ldo < col_off + co)
        return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(y)) & 15) return EPNET_ERR_BAD_ARG;
    const long long total = (long long)b * h * k * w * k * (co / 4);
    if (total == 0) return EPNET_OK;
    deconv_shuffle_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, w, k, co, y, out, ldo, col_off, total);
    EPNET_RETURN_LAUNCH_STATUS();
}

// LI-Fusion attention tail (IA_Layer.forward, lib/net/pointnet2_msg.py:79-96) on point-major rows, one pass:
//   att = sigmoid(w3 . tanh(r1 + r2) + b3)        r1 = fc1(img) (+ both biases), r2 = fc2(point): (rows, rc)
//   out[row][0..c) = x[row][0..c) * att            x = conv1(img) after BN+ReLU: (rows, c)
// replacing the add / tanh / gemv / add / sigmoid / mul launches of the op-by-op version.  A warp per row; rc % 4 == c % 4 == 0.
namespace epnet {
__global__ void __launch_bounds__(256)
attention_scale_pm_kernel(int rows, int rc, int c, const float *__restrict__ r1, int ld1, const float *__restrict__ r2, int ld2,
                          const float *__restrict__ w3, const float *__restrict__ b3, const float *__restrict__ x, int ldx,
                          float *__restrict__ out, int ldo)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4 *a = reinterpret_cast<const float4 *>(r1 + (size_t)row * ld1);
    const float4 *b = reinterpret_cast<const float4 *>(r2 + (size_t)row * ld2);
    const float4 *w = reinterpret_cast<const float4 *>(w3);
    float dot = 0.f;
    for (int j = lane; j < rc / 4; j += 32) {
        const float4 p = __ldg(a + j), q = __ldg(b + j), ww = __ldg(w + j);
        dot += ww.x * tanhf(p.x + q.x) + ww.y * tanhf(p.y + q.y) + ww.z * tanhf(p.z + q.z) + ww.w * tanhf(p.w + q.w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float att = 1.0f / (1.0f + expf(-(dot + __ldg(b3))));
    const float4 *xs = reinterpret_cast<const float4 *>(x + (size_t)row * ldx);
    float4 *os = reinterpret_cast<float4 *>(out + (size_t)row * ldo);
    for (int j = lane; j < c / 4; j += 32) {
        const float4 v = __ldg(xs + j);
        os[j] = make_float4(v.x * att, v.y * att, v.z * att, v.w * att);
    }
}
}  // namespace epnet

EPNET_API int epnet_attention_scale_pm(int rows, int rc, int c, const float *r1, int ld1, const float *r2, int ld2, const float *w3,
                                       const float *b3, const float *x, int ldx, float *out, int ldo, void *stream)
{
    using namespace epnet;
    if (rows < 0 || rc <= 0 || c <= 0 || !r1 || !r2 || !w3 || !b3 || !x || !out) return EPNET_ERR_BAD_ARG;
    if ((rc & 3) || (c & 3) || (ld1 & 3) || (ld2 & 3) || (ldx & 3) || (ldo & 3) || ld1 < rc || ld2 < rc || ldx < c || ldo < c) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(r1) | reinterpret_cast<uintptr_t>(r2) | reinterpret_cast<uintptr_t>(w3) | reinterpret_cast<uintptr_t>(x) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        return EPNET_ERR_BAD_ARG;
    if (rows == 0) return EPNET_OK;
    attention_scale_pm_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(rows, rc, c, r1, ld1, r2, ld2, w3, b3, x, ldx, out, ldo);
    EPNET_RETURN_LAUNCH_STATUS();
}
