// Ball query through a spatially sorted copy of the cloud.
//
// ball_query_kernel (ball_query.cu) tests every centre against every point: M*N distance evaluations, 67 M per scene at the
// first set-abstraction level, almost all of them misses (r = 0.1 m in an 80 m scene).  The result, however, is a pure function
// of the SET of points inside the ball -- "the first nsample indices in ascending order" (ball_query_gpu.cu:23-44) -- so the
// scan order is free.  Here:
//   bucket_cloud_kernel   sorts a scene once by a 30-bit Morton key (cubic cells, bitonic sort in shared memory) and emits the
//                         sorted points as (x, y, z, original index) plus the exact bounding box of every 64 consecutive points;
//   ball_query_sorted_kernel  gives each centre a warp: the lanes test the bucket boxes against the ball (conservatively), only
//                         overlapping buckets are scanned with the reference's exact distance expression, and the hits pass
//                         through a register-resident sorted list of the nsample smallest original indices (insertion by
//                         ballot/popc rank + shuffle shift).  Hits that cannot enter the list any more are rejected per lane by
//                         comparing against the list's current maximum, so a dense ball costs little more than a sparse one.
// Same set, same order, same padding rule => bit-identical output; work drops from N to (#buckets + 64 * overlapping buckets).
#include "common.cuh"

namespace epnet {

constexpr int kBucket = 64;             // sorted points per bounding box
constexpr int kSortThreads = 1024;
constexpr int kSortMaxPoints = 16384;   // 8-byte (key, index) pairs of one scene in shared memory: 128 KB
constexpr float kBoxShrink = 0.999996f; // box distance is compared deflated: covers the few-ulp difference between the two roundings

__device__ __forceinline__ uint32_t spread10(uint32_t v)  // 10 bits -> every third bit
{
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// one CTA per scene; npad = power of two >= n (<= kSortMaxPoints); sorted: (B, npad) float4; boxes: (B, npad / 64, 2) float4
__global__ void __launch_bounds__(kSortThreads, 1)
bucket_cloud_kernel(int n, int npad, const float *__restrict__ xyz, float4 *__restrict__ sorted, float4 *__restrict__ boxes)
{
    extern __shared__ __align__(16) unsigned long long s_key[];
    __shared__ float s_red[6][32];
    __shared__ float s_box[kSortThreads / 32][6];
    const int scene = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    xyz += (size_t)scene * n * 3;
    sorted += (size_t)scene * npad;
    boxes += (size_t)scene * (npad / kBucket) * 2;

    // scene bounding box
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = tid; k < n; k += kSortThreads)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(xyz + 3 * k + a);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { s_red[a][warp] = lo[a]; s_red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = s_red[a][lane];
        hi[a] = s_red[3 + a][lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    const float extent = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), fmaxf(hi[2] - lo[2], 1e-20f));
    const float scale = 1023.0f / extent;  // cubic cells: the short axes simply use fewer of their 10 bits

    // (Morton key << 32 | original index); padding sorts last.  NaN/inf coordinates land in cell 0 or 1023: still a valid order.
    for (int k = tid; k < npad; k += kSortThreads) {
        unsigned long long e = 0xffffffff00000000ull | (unsigned)k;
        if (k < n) {
            uint32_t q[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) q[a] = (uint32_t)fminf(fmaxf((__ldg(xyz + 3 * k + a) - lo[a]) * scale, 0.f), 1023.f);
            const uint32_t key = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
            e = ((unsigned long long)key << 32) | (unsigned)k;
        }
        s_key[k] = e;
    }
    __syncthreads();
    for (int size = 2; size <= npad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < npad / 2; t += kSortThreads) {  // one compare-exchange per thread and step: no idle half
                const int i = 2 * t - (t & (stride - 1)), l = i + stride;
                const unsigned long long a = s_key[i], b = s_key[l];
                if ((a > b) == ((i & size) == 0)) { s_key[i] = b; s_key[l] = a; }
            }
            __syncthreads();
        }

    // sorted points and the exact box of every 64 of them (two warps per bucket)
    for (int base = 0; base < npad; base += kSortThreads) {
        const int i = base + tid;  // npad is a multiple of 32; kSortThreads | npad or npad < kSortThreads
        float x = 1.0e30f, y = 1.0e30f, z = 1.0e30f;
        int k = -1;
        float bl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bh[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        if (i < npad) {
            const unsigned idx = (unsigned)(s_key[i] & 0xffffffffull);
            if ((int)idx < n) {
                k = (int)idx;
                x = __ldg(xyz + 3 * k); y = __ldg(xyz + 3 * k + 1); z = __ldg(xyz + 3 * k + 2);
                bl[0] = bh[0] = x; bl[1] = bh[1] = y; bl[2] = bh[2] = z;
            }
            sorted[i] = make_float4(x, y, z, __int_as_float(k));
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                bl[a] = fminf(bl[a], __shfl_xor_sync(0xffffffffu, bl[a], o));
                bh[a] = fmaxf(bh[a], __shfl_xor_sync(0xffffffffu, bh[a], o));
            }
            if (lane == 0) { s_box[warp][a] = bl[a]; s_box[warp][3 + a] = bh[a]; }
        }
        __syncthreads();
        if (tid < kSortThreads / kBucket) {  // bucket = warps 2*tid, 2*tid + 1 of this pass
            const int g = base / kBucket + tid;
            if (g < npad / kBucket) {
                const float *p = s_box[2 * tid], *q = s_box[2 * tid + 1];
                boxes[2 * g] = make_float4(fminf(p[0], q[0]), fminf(p[1], q[1]), fminf(p[2], q[2]), 0.f);
                boxes[2 * g + 1] = make_float4(fmaxf(p[3], q[3]), fmaxf(p[4], q[4]), fmaxf(p[5], q[5]), 0.f);
            }
        }
        __syncthreads();
    }
}

constexpr int kBqsWarps = 8;
constexpr int kBqsCentresPerWarp = 4;
constexpr int kBqsMaxBuckets = kSortMaxPoints / kBucket;  // 256

// one warp per centre (kBqsCentresPerWarp centres in turn); nsample <= 64
__global__ void __launch_bounds__(kBqsWarps * 32)
ball_query_sorted_kernel(int npad, int m, float radius, int nsample, const float *__restrict__ new_xyz, const float4 *__restrict__ sorted,
                         const float4 *__restrict__ boxes, int *__restrict__ idx)
{
    __shared__ float4 s_boxes[kBqsMaxBuckets * 2];
    const int scene = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = npad / kBucket;
    sorted += (size_t)scene * npad;
    boxes += (size_t)scene * nb * 2;
    new_xyz += (size_t)scene * m * 3;
    idx += (size_t)scene * m * nsample;
    for (int e = threadIdx.x; e < nb * 2; e += kBqsWarps * 32) s_boxes[e] = __ldg(boxes + e);
    __syncthreads();

    const float r2 = __fmul_rn(radius, radius);
    const int INF = 0x7fffffff;
    for (int c = 0; c < kBqsCentresPerWarp; ++c) {
        const int j = (blockIdx.x * kBqsWarps + warp) * kBqsCentresPerWarp + c;
        if (j >= m) break;  // warp-uniform
        const float cx = __ldg(new_xyz + 3 * j), cy = __ldg(new_xyz + 3 * j + 1), cz = __ldg(new_xyz + 3 * j + 2);
        // the nsample smallest original indices so far, ascending: slot s lives in lane s % 32 of l0 (s < 32) or l1
        int l0 = INF, l1 = INF;
        int bound = INF;  // value a new index must beat: INF until the list is full, then its largest entry
        for (int g0 = 0; g0 < nb; g0 += 32) {
            const int g = g0 + lane;
            bool near = false;
            if (g < nb) {
                const float4 lo = s_boxes[2 * g], hi = s_boxes[2 * g + 1];
                const float ex = fmaxf(fmaxf(lo.x - cx, cx - hi.x), 0.f), ey = fmaxf(fmaxf(lo.y - cy, cy - hi.y), 0.f),
                            ez = fmaxf(fmaxf(lo.z - cz, cz - hi.z), 0.f);
                near = (ex * ex + ey * ey + ez * ez) * kBoxShrink < r2;  // an empty (all-padding) bucket has lo > hi: never near
            }
            uint32_t todo = __ballot_sync(0xffffffffu, near);
            while (todo) {
                const int gb = g0 + __ffs(todo) - 1;
                todo &= todo - 1;
#pragma unroll
                for (int h = 0; h < kBucket / 32; ++h) {
                    const float4 p = __ldg(sorted + gb * kBucket + h * 32 + lane);
                    const int k = __float_as_int(p.w);
                    const float d2 = sqdist_ref(cx, cy, cz, p.x, p.y, p.z);
                    // padding points sit at 1e30: never inside; an index that cannot enter the list any more is dropped here
                    uint32_t hits = __ballot_sync(0xffffffffu, d2 < r2 && k < bound);
                    while (hits) {
                        const int src = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const int v = __shfl_sync(0xffffffffu, k, src);
                        if (v < bound) {  // warp-uniform; bound may have dropped since the ballot
                            const int pos = __popc(__ballot_sync(0xffffffffu, l0 < v)) + __popc(__ballot_sync(0xffffffffu, l1 < v));
                            const int up0 = __shfl_up_sync(0xffffffffu, l0, 1);
                            int up1 = __shfl_up_sync(0xffffffffu, l1, 1);
                            const int carry = __shfl_sync(0xffffffffu, l0, 31);
                            if (lane == 0) up1 = carry;
                            l1 = (32 + lane > pos) ? up1 : ((32 + lane == pos) ? v : l1);
                            l0 = (lane > pos) ? up0 : ((lane == pos) ? v : l0);
                            // entries past nsample fall off the end of the window that is read back
                            const int last = nsample - 1;
                            bound = __shfl_sync(0xffffffffu, last < 32 ? l0 : l1, last & 31);
                        }
                    }
                }
            }
        }
        // write-out: ascending hits, then the first hit repeated (ball_query_gpu.cu:36-40); no hit: the caller's zeros stay
        const int first = __shfl_sync(0xffffffffu, l0, 0);
        if (first != INF) {
            int *o = idx + (size_t)j * nsample;
            if (lane < nsample) o[lane] = l0 != INF ? l0 : first;
            if (32 + lane < nsample) o[32 + lane] = l1 != INF ? l1 : first;
        }
    }
}

}  // namespace epnet

EPNET_API int epnet_bucket_cloud(int b, int n, int npad, const float *xyz, float *sorted, float *boxes, void *stream)
{
    using namespace epnet;
    if (b < 0 || n <= 0 || !xyz || !sorted || !boxes) return EPNET_ERR_BAD_ARG;
    if (npad < n || npad < kBucket || npad > kSortMaxPoints || (npad & (npad - 1)) != 0) return EPNET_ERR_BAD_ARG;
    if (((reinterpret_cast<uintptr_t>(sorted) | reinterpret_cast<uintptr_t>(boxes)) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    const size_t smem = (size_t)npad * sizeof(unsigned long long);
    cudaError_t e = cudaFuncSetAttribute(bucket_cloud_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSortMaxPoints * sizeof(unsigned long long)));
    if (e != cudaSuccess) return (int)e;
    bucket_cloud_kernel<<<b, kSortThreads, smem, (cudaStream_t)stream>>>(n, npad, xyz, reinterpret_cast<float4 *>(sorted),
                                                                         reinterpret_cast<float4 *>(boxes));
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_ball_query_sorted(int b, int npad, int m, float radius, int nsample, const float *new_xyz, const float *sorted,
                                      const float *boxes, int *idx, void *stream)
{
    using namespace epnet;
    if (b < 0 || m < 0 || nsample < 1 || nsample > 64 || !new_xyz || !sorted || !boxes || !idx) return EPNET_ERR_BAD_ARG;
    if (npad < kBucket || npad > kSortMaxPoints || (npad & (npad - 1)) != 0) return EPNET_ERR_BAD_ARG;
    if (b == 0 || m == 0) return EPNET_OK;
    dim3 grid((m + kBqsWarps * kBqsCentresPerWarp - 1) / (kBqsWarps * kBqsCentresPerWarp), b);
    ball_query_sorted_kernel<<<grid, kBqsWarps * 32, 0, (cudaStream_t)stream>>>(npad, m, radius, nsample, new_xyz,
                                                                                reinterpret_cast<const float4 *>(sorted),
                                                                                reinterpret_cast<const float4 *>(boxes), idx);
    EPNET_RETURN_LAUNCH_STATUS();
}
