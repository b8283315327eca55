// Ball query for B200.  Replaces ball_query_kernel_fast
// (/root/reference/pointnet2_lib/pointnet2/src/ball_query_gpu.cu:9-45) bit-exactly: for each centre
// the first `nsample` cloud indices, in ascending order, with d2 < r*r; remaining slots repeat the
// first hit; a centre with no hit keeps the caller's zeros.
//
// Reference: one thread per centre, a serial scan of the whole cloud from L2 per thread.
// Here: the cloud is staged through shared memory in tiles by 1-D bulk (TMA) copies, double buffered
// on mbarriers; a warp owns kCentresPerWarp centres held in registers, each lane tests ONE staged
// point against all of them, and hits are compacted in index order with ballot + popc.  The (N,3)
// array is copied verbatim: lane l reads words 3l..3l+2, and 3 is coprime with 32, so the AoS tile
// is bank-conflict-free without a transpose.  Centres that are full stop costing instructions; a CTA
// whose centres are all full stops staging tiles.
#include "common.cuh"

namespace epnet {

constexpr int kBqWarps = 8;
constexpr int kBqCentresPerWarp = 2;
constexpr int kBqCentresPerCta = kBqWarps * kBqCentresPerWarp;  // 16
constexpr int kBqTile = 1920;                                  // points per stage (22.5 KB); two stages stay under the 48 KB static limit

__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                  int *__restrict__ idx, int use_bulk)
{
    __shared__ __align__(128) float tile[2][kBqTile * 3];
    __shared__ __align__(8) uint64_t full[2];

    const int scene = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    xyz += (size_t)scene * n * 3;
    new_xyz += (size_t)scene * m * 3;
    idx += (size_t)scene * m * nsample;

    const float r2 = __fmul_rn(radius, radius);
    const int c0 = blockIdx.x * kBqCentresPerCta + warp * kBqCentresPerWarp;

    float cx[kBqCentresPerWarp], cy[kBqCentresPerWarp], cz[kBqCentresPerWarp];
    int cnt[kBqCentresPerWarp], first[kBqCentresPerWarp];
#pragma unroll
    for (int c = 0; c < kBqCentresPerWarp; ++c) {
        const int j = c0 + c;
        const bool live = j < m;
        cx[c] = live ? __ldg(new_xyz + 3 * j) : 0.f;
        cy[c] = live ? __ldg(new_xyz + 3 * j + 1) : 0.f;
        cz[c] = live ? __ldg(new_xyz + 3 * j + 2) : 0.f;
        cnt[c] = live ? 0 : nsample;  // out-of-range centres count as already full
        first[c] = 0;
    }

    const int ntiles = (n + kBqTile - 1) / kBqTile;
    if (use_bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&full[0], 1);
            mbar_init(&full[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)min(kBqTile, n) * 12u;
            mbar_arrive_expect_tx(&full[0], bytes);
            bulk_g2s(tile[0], xyz, bytes, &full[0]);
        }
    }

    int t = 0;
    for (; t < ntiles; ++t) {
        const int buf = t & 1;
        const int base = t * kBqTile;
        const int count = min(kBqTile, n - base);
        if (use_bulk) {
            // prefetch the next tile into the other buffer (its readers finished before the barrier below)
            if (threadIdx.x == 0 && t + 1 < ntiles) {
                const uint32_t bytes = (uint32_t)min(kBqTile, n - base - kBqTile) * 12u;
                mbar_arrive_expect_tx(&full[buf ^ 1], bytes);
                bulk_g2s(tile[buf ^ 1], xyz + (size_t)(base + kBqTile) * 3, bytes, &full[buf ^ 1]);
            }
            mbar_wait(&full[buf], (t >> 1) & 1);
        } else {
            for (int f = threadIdx.x; f < count * 3; f += kBqWarps * 32) tile[buf][f] = __ldg(xyz + (size_t)base * 3 + f);
            __syncthreads();
        }

        const float *tp = tile[buf];
        bool all_full = true;
#pragma unroll
        for (int c = 0; c < kBqCentresPerWarp; ++c) all_full = all_full && (cnt[c] >= nsample);
        if (!all_full) {
            for (int off = 0; off < count; off += 32) {
                const int p = off + lane;
                const bool valid = p < count;
                const float x = valid ? tp[3 * p] : 0.f, y = valid ? tp[3 * p + 1] : 0.f, z = valid ? tp[3 * p + 2] : 0.f;
                const int k = base + p;
#pragma unroll
                for (int c = 0; c < kBqCentresPerWarp; ++c) {
                    if (cnt[c] < nsample) {  // warp-uniform
                        const float d2 = sqdist_ref(cx[c], cy[c], cz[c], x, y, z);
                        const bool hit = valid && (d2 < r2);
                        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                        if (mask) {
                            if (cnt[c] == 0) first[c] = base + off + (__ffs(mask) - 1);
                            const int pos = cnt[c] + __popc(mask & lanemask_lt());
                            if (hit && pos < nsample) idx[(size_t)(c0 + c) * nsample + pos] = k;
                            cnt[c] += __popc(mask);
                        }
                    }
                }
            }
        }
        // everyone is done with `buf` (it is refilled two tiles later) -- and vote on early exit
        all_full = true;
#pragma unroll
        for (int c = 0; c < kBqCentresPerWarp; ++c) all_full = all_full && (cnt[c] >= nsample);
        if (__syncthreads_and(all_full)) break;
    }
    // left early with a prefetch still in flight: it must land before this CTA's shared memory is released
    if (use_bulk && threadIdx.x == 0 && t + 1 < ntiles) mbar_wait(&full[(t + 1) & 1], ((t + 1) >> 1) & 1);

    // pad with the first hit (ball_query_gpu.cu:36-40); centres without a hit keep the caller's zeros
#pragma unroll
    for (int c = 0; c < kBqCentresPerWarp; ++c) {
        const int j = c0 + c;
        if (j < m && cnt[c] > 0 && cnt[c] < nsample)
            for (int l = cnt[c] + lane; l < nsample; l += 32) idx[(size_t)j * nsample + l] = first[c];
    }
}

}  // namespace epnet

EPNET_API int epnet_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx,
                               void *stream)
{
    using namespace epnet;
    if (b < 0 || n < 0 || m < 0 || nsample < 0 || !new_xyz || !xyz || !idx) return EPNET_ERR_BAD_ARG;
    if (b == 0 || m == 0 || n == 0 || nsample == 0) return EPNET_OK;
    // bulk copies need 16-byte aligned sources and sizes: every scene base (n*12 B) and every tile
    // (kBqTile*12 B) must be; the tail tile's size is n*12 - k*tile bytes.
    const int use_bulk = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) && (n % 4 == 0);
    dim3 grid((m + kBqCentresPerCta - 1) / kBqCentresPerCta, b);
    ball_query_kernel<<<grid, kBqWarps * 32, 0, (cudaStream_t)stream>>>(n, m, radius, nsample, new_xyz, xyz, idx, use_bulk);
    EPNET_RETURN_LAUNCH_STATUS();
}
