// Furthest point sampling for B200.  Replaces furthest_point_sampling_kernel
// (/root/reference/pointnet2_lib/pointnet2/src/sampling_gpu.cu:93-209) bit-exactly.
//
// What has to be reproduced: every iteration j picks arg-max over k of temp[k] = min(temp[k], d(k, last)).
// The reference resolves equal maxima through (a) a strided per-thread scan that keeps the first
// maximum and (b) a shared-memory tree that keeps the lower slot unless the upper is strictly larger.
// The net effect is: among tied k the winner minimises  tie(k) = (bitreverse_L(k mod BS), k div BS),
// BS = 2^L = min(1024, 2^floor(log2 N))  (cuda_utils.h:10-14).  We compute that order directly:
// a candidate is the pair (distance bits, tie key); distances are >= 0 so their IEEE bits order as
// unsigned ints, and the winner is  max distance, then min tie key -- two warp `redux.sync`
// instructions per level instead of a 10-barrier shared-memory tree.
//
// Layout (resident variant, N <= 16384): one CTA per scene, BS threads; coordinates are transposed once
// into shared memory as three float planes (192 KB at N=16384, conflict-free: lane l reads word l), the
// running distances live in registers (thread t owns points t, t+BS, ...: the reference's ownership, so
// its "first maximum per thread" rule carries over unchanged).  One __syncthreads per iteration:
// warp winners go to a double-buffered slot array that every warp re-reduces for itself.
// Streaming variant (N > 16384): same reduction, coordinates and temp stay in global memory / L2.
#include "common.cuh"

namespace epnet {

constexpr int kFpsMaxResident = 16384;  // 3 planes * 4 B * 16384 = 192 KB of the 227 KB shared memory

struct FpsCand {
    uint32_t dist_bits;
    uint32_t tie;
};

__device__ __forceinline__ uint32_t fps_tie_key(uint32_t k, uint32_t tid, int L)
{
    // tid < 2^L so __brev(tid) occupies the top L bits; k >> L < 2^(32-L).
    return __brev(tid) | (k >> L);
}
__device__ __forceinline__ uint32_t fps_tie_decode(uint32_t tie, int L)
{
    const uint32_t low = L ? (0xffffffffu >> L) : 0xffffffffu;
    const uint32_t tid = __brev(tie & ~low);
    return ((tie & low) << L) | tid;
}

// Block-wide arg-max with the reference's tie rule.  Returns the winning point index to every thread.
// slots: 2 * 32 entries, parity flips per call so one barrier per call is enough.
__device__ __forceinline__ uint32_t fps_block_argmax(uint32_t dist_bits, uint32_t tie, FpsCand *slots, int parity, int nwarps, int L)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t mx = warp_max_u32(dist_bits);
    uint32_t mt = warp_min_u32(dist_bits == mx ? tie : 0xffffffffu);
    if (nwarps > 1) {
        FpsCand *s = slots + parity * 32;
        if (lane == 0) {
            s[warp].dist_bits = mx;
            s[warp].tie = mt;
        }
        __syncthreads();
        FpsCand c;
        c.dist_bits = 0u;
        c.tie = 0xffffffffu;
        if (lane < nwarps) c = s[lane];
        mx = warp_max_u32(c.dist_bits);
        mt = warp_min_u32(c.dist_bits == mx ? c.tie : 0xffffffffu);
    }
    return fps_tie_decode(mt, L);
}

template <int P>
__global__ void __launch_bounds__(1024, 1)
fps_resident_kernel(int n, int m, int L, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx)
{
    extern __shared__ __align__(16) float fps_smem[];
    __shared__ FpsCand slots[64];
    float *xs = fps_smem, *ys = fps_smem + n, *zs = fps_smem + 2 * n;

    const int scene = blockIdx.x;
    xyz += (size_t)scene * n * 3;
    temp += (size_t)scene * n;
    idx += (size_t)scene * m;

    const int tid = threadIdx.x;
    const int bs = 1 << L;
    const int nthreads = blockDim.x;  // max(bs, 32)
    const int nwarps = nthreads >> 5;

    // transpose (N,3) -> three planes; the flat read is fully coalesced
    for (int f = tid; f < 3 * n; f += nthreads) {
        const int p = f / 3, c = f - 3 * p;
        fps_smem[c * n + p] = xyz[f];
    }
    float t[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = tid + i * bs;
        t[i] = (tid < bs && k < n) ? temp[k] : 0.f;
    }
    if (tid == 0) idx[0] = 0;
    __syncthreads();

    uint32_t last = 0;
    for (int j = 1; j < m; ++j) {
        const float lx = xs[last], ly = ys[last], lz = zs[last];
        float best = -1.f;
        uint32_t besti = 0;
        if (tid < bs) {
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int k = tid + i * bs;
                if (k < n) {
                    const float d = sqdist_ref(xs[k], ys[k], zs[k], lx, ly, lz);
                    const float d2 = fminf(d, t[i]);
                    t[i] = d2;
                    if (d2 > best) {
                        best = d2;
                        besti = k;
                    }
                }
            }
        }
        const bool has = tid < bs;  // every thread below bs owns point k = tid < n
        last = fps_block_argmax(has ? __float_as_uint(best) : 0u, has ? fps_tie_key(besti, tid, L) : 0xffffffffu, slots, j & 1,
                                nwarps, L);
        if (tid == 0) idx[j] = (int)last;
    }

    // temp is an in/out buffer in the reference; leave the final running distances behind.
    if (tid < bs) {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int k = tid + i * bs;
            if (k < n) temp[k] = t[i];
        }
    }
}

// N too large for one SM's shared memory: coordinates and running distances stream from L2.
__global__ void __launch_bounds__(1024, 1)
fps_streaming_kernel(int n, int m, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx)
{
    __shared__ FpsCand slots[64];
    const int scene = blockIdx.x;
    xyz += (size_t)scene * n * 3;
    temp += (size_t)scene * n;
    idx += (size_t)scene * m;
    const int tid = threadIdx.x;
    constexpr int L = 10;
    if (tid == 0) idx[0] = 0;

    uint32_t last = 0;
    for (int j = 1; j < m; ++j) {
        const float lx = __ldg(xyz + 3 * last), ly = __ldg(xyz + 3 * last + 1), lz = __ldg(xyz + 3 * last + 2);
        float best = -1.f;
        uint32_t besti = 0;
        for (int k = tid; k < n; k += 1024) {
            const float d = sqdist_ref(__ldg(xyz + 3 * k), __ldg(xyz + 3 * k + 1), __ldg(xyz + 3 * k + 2), lx, ly, lz);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > best) {
                best = d2;
                besti = k;
            }
        }
        last = fps_block_argmax(__float_as_uint(best), fps_tie_key(besti, tid, L), slots, j & 1, 32, L);
        if (tid == 0) idx[j] = (int)last;
    }
}

template <int P>
static int launch_resident(int b, int n, int m, int L, const float *xyz, float *temp, int *idx, cudaStream_t st)
{
    const size_t smem = (size_t)3 * n * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_resident_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int threads = (1 << L) < 32 ? 32 : (1 << L);
    fps_resident_kernel<P><<<b, threads, smem, st>>>(n, m, L, xyz, temp, idx);
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet

EPNET_API int epnet_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream)
{
    using namespace epnet;
    if (b < 0 || n <= 0 || m < 0 || !xyz || !temp || !idx) return EPNET_ERR_BAD_ARG;
    if (b == 0 || m == 0) return EPNET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int L = 0;
    while ((2 << L) <= n && L < 10) ++L;  // BS = 2^L = min(1024, 2^floor(log2 n))
    if (n <= kFpsMaxResident) {
        const int per = (n + (1 << L) - 1) >> L;
        if (per <= 1) return launch_resident<1>(b, n, m, L, xyz, temp, idx, st);
        if (per <= 2) return launch_resident<2>(b, n, m, L, xyz, temp, idx, st);
        if (per <= 4) return launch_resident<4>(b, n, m, L, xyz, temp, idx, st);
        if (per <= 8) return launch_resident<8>(b, n, m, L, xyz, temp, idx, st);
        return launch_resident<16>(b, n, m, L, xyz, temp, idx, st);
    }
    fps_streaming_kernel<<<b, 1024, 0, st>>>(n, m, xyz, temp, idx);
    EPNET_RETURN_LAUNCH_STATUS();
}
