// three_nn / three_interpolate (+grad) for B200.  Replace three_nn_kernel_fast,
// three_interpolate_kernel_fast and three_interpolate_grad_kernel_fast
// (/root/reference/pointnet2_lib/pointnet2/src/interpolate_gpu.cu:9-52, 77-97, 120-142).
#include "common.cuh"

namespace epnet {

// ---------------------------------------------------------------------------------------------
// three_nn: for each query the three smallest squared distances to the known set, ordered, ties to
// the lower index (strict '<' cascade, interpolate_gpu.cu:37-48).  The reference keeps the running
// bests in double initialised to 1e40 but only ever compares exactly-widened floats, so float
// compares against +inf are equivalent, and an unfilled slot is written as +inf with index 0.
// Known points are staged through shared memory by bulk (TMA) copies, double buffered; every lane
// reads the same staged point (a broadcast), so the (m,3) array needs no transpose.
// ---------------------------------------------------------------------------------------------
constexpr int kNnThreads = 128;
constexpr int kNnTile = 960;   // 22.5 KB for the two stages: 8 CTAs per SM (1920 measured 110 us vs 81 us at 16384 x 4096, 480: 86 us)
constexpr int kNnSlices = 4;  // lanes per query: each scans every 4th group of 4 staged points, partial top-3s merged by shuffle

struct Top3 {
    float d0, d1, d2;
    int i0, i1, i2;
};

// lexicographic (d, k) insertion: used to merge partial results, where index order is no longer implied by arrival order
__device__ __forceinline__ void top3_insert_lex(Top3 &t, float d, int k)
{
    if (d < t.d2 || (d == t.d2 && k < t.i2)) {
        if (d < t.d1 || (d == t.d1 && k < t.i1)) {
            t.d2 = t.d1; t.i2 = t.i1;
            if (d < t.d0 || (d == t.d0 && k < t.i0)) {
                t.d1 = t.d0; t.i1 = t.i0;
                t.d0 = d; t.i0 = k;
            } else {
                t.d1 = d; t.i1 = k;
            }
        } else {
            t.d2 = d; t.i2 = k;
        }
    }
}

__device__ __forceinline__ void top3_insert(Top3 &t, float d, int k)
{
    if (d < t.d2) {
        if (d < t.d1) {
            t.d2 = t.d1; t.i2 = t.i1;
            if (d < t.d0) {
                t.d1 = t.d0; t.i1 = t.i0;
                t.d0 = d; t.i0 = k;
            } else {
                t.d1 = d; t.i1 = k;
            }
        } else {
            t.d2 = d; t.i2 = k;
        }
    }
}

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known, float *__restrict__ dist2,
                int *__restrict__ idx, int use_bulk, float *__restrict__ weight)
{
    __shared__ __align__(128) float tile[2][kNnTile * 3];
    __shared__ __align__(8) uint64_t full[2];

    const int scene = blockIdx.y;
    unknown += (size_t)scene * n * 3;
    known += (size_t)scene * m * 3;
    dist2 += (size_t)scene * n * 3;
    idx += (size_t)scene * n * 3;
    if (weight) weight += (size_t)scene * n * 3;

    const int q = blockIdx.x * (kNnThreads / kNnSlices) + (threadIdx.x / kNnSlices);
    const int slice = threadIdx.x % kNnSlices;
    const bool live = q < n;
    const float ux = live ? __ldg(unknown + 3 * q) : 0.f;
    const float uy = live ? __ldg(unknown + 3 * q + 1) : 0.f;
    const float uz = live ? __ldg(unknown + 3 * q + 2) : 0.f;

    const float inf = __int_as_float(0x7f800000);
    Top3 best = {inf, inf, inf, 0, 0, 0};

    const int ntiles = (m + kNnTile - 1) / kNnTile;
    if (use_bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&full[0], 1);
            mbar_init(&full[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)min(kNnTile, m) * 12u;
            mbar_arrive_expect_tx(&full[0], bytes);
            bulk_g2s(tile[0], known, bytes, &full[0]);
        }
    }
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const int base = t * kNnTile;
        const int count = min(kNnTile, m - base);
        if (use_bulk) {
            if (threadIdx.x == 0 && t + 1 < ntiles) {
                const uint32_t bytes = (uint32_t)min(kNnTile, m - base - kNnTile) * 12u;
                mbar_arrive_expect_tx(&full[buf ^ 1], bytes);
                bulk_g2s(tile[buf ^ 1], known + (size_t)(base + kNnTile) * 3, bytes, &full[buf ^ 1]);
            }
            mbar_wait(&full[buf], (t >> 1) & 1);
        } else {
            for (int f = threadIdx.x; f < count * 3; f += kNnThreads) tile[buf][f] = __ldg(known + (size_t)base * 3 + f);
            __syncthreads();
        }
        const float *tp = tile[buf];
        const int full = count & ~3;
        for (int p = slice * 4; p < full; p += 4 * kNnSlices) {
            // 12 consecutive floats = 3 x LDS.128 (tile rows are 48 B, p % 4 == 0 keeps 16-byte alignment)
            const float4 a = *reinterpret_cast<const float4 *>(tp + 3 * p);
            const float4 b = *reinterpret_cast<const float4 *>(tp + 3 * p + 4);
            const float4 c = *reinterpret_cast<const float4 *>(tp + 3 * p + 8);
            const float d0 = sqdist_ref(ux, uy, uz, a.x, a.y, a.z);
            const float d1 = sqdist_ref(ux, uy, uz, a.w, b.x, b.y);
            const float d2 = sqdist_ref(ux, uy, uz, b.z, b.w, c.x);
            const float d3 = sqdist_ref(ux, uy, uz, c.y, c.z, c.w);
            const float lo = fminf(fminf(d0, d1), fminf(d2, d3));
            if (lo < best.d2) {  // rare once the three bests have settled; within a slice indices arrive in ascending order
                top3_insert(best, d0, base + p);
                top3_insert(best, d1, base + p + 1);
                top3_insert(best, d2, base + p + 2);
                top3_insert(best, d3, base + p + 3);
            }
        }
        if (slice == 0)
            for (int p = full; p < count; ++p) top3_insert_lex(best, sqdist_ref(ux, uy, uz, tp[3 * p], tp[3 * p + 1], tp[3 * p + 2]), base + p);
        __syncthreads();  // buffer `buf` may be refilled from the next iteration on
    }
    // merge the kNnSlices partial top-3s of a query (adjacent lanes) -- lexicographic (d, k) is exactly the reference's order
#pragma unroll
    for (int step = 1; step < kNnSlices; step <<= 1) {
        const float e0 = __shfl_xor_sync(0xffffffffu, best.d0, step), e1 = __shfl_xor_sync(0xffffffffu, best.d1, step),
                    e2 = __shfl_xor_sync(0xffffffffu, best.d2, step);
        const int j0 = __shfl_xor_sync(0xffffffffu, best.i0, step), j1 = __shfl_xor_sync(0xffffffffu, best.i1, step),
                  j2 = __shfl_xor_sync(0xffffffffu, best.i2, step);
        top3_insert_lex(best, e0, j0);
        top3_insert_lex(best, e1, j1);
        top3_insert_lex(best, e2, j2);
    }
    if (live && slice == 0) {
        dist2[3 * q] = best.d0; dist2[3 * q + 1] = best.d1; dist2[3 * q + 2] = best.d2;
        idx[3 * q] = best.i0;   idx[3 * q + 1] = best.i1;   idx[3 * q + 2] = best.i2;
        if (weight) {
            // inverse-distance weights of PointnetFPModule (pointnet2_utils.py:98 sqrt, pointnet2_modules.py:157-159), once per query
            const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(best.d0), 1e-8f));
            const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(best.d1), 1e-8f));
            const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(best.d2), 1e-8f));
            const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
            weight[3 * q] = __fdiv_rn(r0, norm); weight[3 * q + 1] = __fdiv_rn(r1, norm); weight[3 * q + 2] = __fdiv_rn(r2, norm);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// three_interpolate: out[b,c,i] = fma(w2,p2, fma(w0,p0, w1*p1))  (the contraction nvcc -O2 applies to
// interpolate_gpu.cu:96; reproducing it makes the result bit-identical, not merely within 1e-5).
// A thread owns 4 consecutive targets: indices/weights are read once as 3+3 128-bit loads and reused
// for kIpChannels channel rows; each row costs 12 L1/L2 gathers and one 128-bit coalesced store.
// ---------------------------------------------------------------------------------------------
constexpr int kIpThreads = 128;
constexpr int kIpChannels = 8;

__global__ void __launch_bounds__(kIpThreads)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                         const float *__restrict__ weight, float *__restrict__ out, int vec_ok)
{
    const int scene = blockIdx.z;
    points += (size_t)scene * c * m;
    idx += (size_t)scene * n * 3;
    weight += (size_t)scene * n * 3;
    out += (size_t)scene * c * n;

    const int i0 = (blockIdx.x * kIpThreads + threadIdx.x) * 4;
    if (i0 >= n) return;
    const int c_begin = blockIdx.y * kIpChannels;
    const int c_end = min(c, c_begin + kIpChannels);

    int id[12];
    float w[12];
    const bool vec = vec_ok && (i0 + 4 <= n);
    if (vec) {
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            const int4 a = __ldg(reinterpret_cast<const int4 *>(idx + 3 * i0) + v);
            const float4 f = __ldg(reinterpret_cast<const float4 *>(weight + 3 * i0) + v);
            id[4 * v] = a.x; id[4 * v + 1] = a.y; id[4 * v + 2] = a.z; id[4 * v + 3] = a.w;
            w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const bool ok = 3 * i0 + e < 3 * n;
            id[e] = ok ? __ldg(idx + 3 * i0 + e) : 0;
            w[e] = ok ? __ldg(weight + 3 * i0 + e) : 0.f;
        }
    }
    for (int ch = c_begin; ch < c_end; ++ch) {
        const float *row = points + (size_t)ch * m;
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            r[e] = __fmaf_rn(w[3 * e + 2], __ldg(row + id[3 * e + 2]),
                             __fmaf_rn(w[3 * e], __ldg(row + id[3 * e]), __fmul_rn(w[3 * e + 1], __ldg(row + id[3 * e + 1]))));
        float *o = out + (size_t)ch * n + i0;
        if (vec) {
            *reinterpret_cast<float4 *>(o) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i0 + e < n) o[e] = r[e];
        }
    }
}

// grad_points[b,c,idx[b,i,k]] += grad_out[b,c,i] * weight[b,i,k]   (fp32 atomics, as the reference)
__global__ void __launch_bounds__(kIpThreads)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx,
                              const float *__restrict__ weight, float *__restrict__ grad_points)
{
    const int scene = blockIdx.z;
    grad_out += (size_t)scene * c * n;
    idx += (size_t)scene * n * 3;
    weight += (size_t)scene * n * 3;
    grad_points += (size_t)scene * c * m;

    const int i = blockIdx.x * kIpThreads + threadIdx.x;
    if (i >= n) return;
    const int c_begin = blockIdx.y * kIpChannels;
    const int c_end = min(c, c_begin + kIpChannels);
    const int i0 = __ldg(idx + 3 * i), i1 = __ldg(idx + 3 * i + 1), i2 = __ldg(idx + 3 * i + 2);
    const float w0 = __ldg(weight + 3 * i), w1 = __ldg(weight + 3 * i + 1), w2 = __ldg(weight + 3 * i + 2);
    for (int ch = c_begin; ch < c_end; ++ch) {
        const float g = __ldg(grad_out + (size_t)ch * n + i);
        float *row = grad_points + (size_t)ch * m;
        atomicAdd(row + i0, g * w0);
        atomicAdd(row + i1, g * w1);
        atomicAdd(row + i2, g * w2);
    }
}

}  // namespace epnet

EPNET_API int epnet_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *stream)
{
    using namespace epnet;
    if (b < 0 || n < 0 || m < 0 || !unknown || !known || !dist2 || !idx) return EPNET_ERR_BAD_ARG;
    if (b == 0 || n == 0) return EPNET_OK;
    const int use_bulk = m > 0 && ((reinterpret_cast<uintptr_t>(known) & 15) == 0) && (m % 4 == 0);
    dim3 grid((n + kNnThreads / kNnSlices - 1) / (kNnThreads / kNnSlices), b);
    three_nn_kernel<<<grid, kNnThreads, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx, use_bulk, nullptr);
    EPNET_RETURN_LAUNCH_STATUS();
}

// three_nn that also emits the normalised inverse-distance weights the feature-propagation module derives from the
// distances (pointnet2_modules.py:157-159), so the interpolation kernel does not recompute them per channel chunk
EPNET_API int epnet_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, float *weight,
                                     void *stream)
{
    using namespace epnet;
    if (b < 0 || n < 0 || m < 0 || !unknown || !known || !dist2 || !idx || !weight) return EPNET_ERR_BAD_ARG;
    if (b == 0 || n == 0) return EPNET_OK;
    const int use_bulk = m > 0 && ((reinterpret_cast<uintptr_t>(known) & 15) == 0) && (m % 4 == 0);
    dim3 grid((n + kNnThreads / kNnSlices - 1) / (kNnThreads / kNnSlices), b);
    three_nn_kernel<<<grid, kNnThreads, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx, use_bulk, weight);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out,
                                      void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || m < 0 || n < 0 || !points || !idx || !weight || !out) return EPNET_ERR_BAD_ARG;
    if (b == 0 || c == 0 || n == 0) return EPNET_OK;
    const int turned = launch_transposed_gather(true, staged_row_fits(m), b, c, m, n, points, idx, weight, out, (cudaStream_t)stream);
    if (turned != kStagedNotApplicable) return turned;
    const int staged = launch_staged_rows(true, b, c, m, n, points, idx, weight, out, (cudaStream_t)stream);
    if (staged != kStagedNotApplicable) return staged;
    dim3 grid(((n + 3) / 4 + kIpThreads - 1) / kIpThreads, (c + kIpChannels - 1) / kIpChannels, b);
    const uintptr_t al = reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out);
    const int vec_ok = (n % 4 == 0) && ((al & 15) == 0);
    three_interpolate_kernel<<<grid, kIpThreads, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out, vec_ok);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                           float *grad_points, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || m < 0 || n < 0 || !grad_out || !idx || !weight || !grad_points) return EPNET_ERR_BAD_ARG;
    if (b == 0 || c == 0 || n == 0) return EPNET_OK;
    const int turned = launch_transposed_scatter(true, b, c, m, n, grad_out, idx, weight, grad_points, (cudaStream_t)stream);
    if (turned != kStagedNotApplicable) return turned;
    dim3 grid((n + kIpThreads - 1) / kIpThreads, (c + kIpChannels - 1) / kIpChannels, b);
    three_interpolate_grad_kernel<<<grid, kIpThreads, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight, grad_points);
    EPNET_RETURN_LAUNCH_STATUS();
}
