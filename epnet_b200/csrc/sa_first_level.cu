// First set-abstraction level in one kernel: group -> re-centre -> three-layer shared MLP -> max over the ball.
//
// Replaces, for a level without input features (RPN.USE_INTENSITY false: the grouped tensor is the three re-centred coordinates),
// QueryAndGroup + SharedMLP + max_pool2d of PointnetSAModuleMSG (/root/reference/pointnet2_lib/pointnet2/pointnet2_modules.py:44-58,
// pointnet2_utils.py:241-264, pytorch_utils.py:20-32 with BatchNorm(eval) folded) for ONE scale: rows = B*npoint*nsample (393 216
// at the published configuration), widths 3 -> 16 -> 16 -> 32 and 3 -> 32 -> 32 -> 64.  With K and N this small a tensor-core tile
// is all fixed cost: the tcgen05 path (grouped-operand GEMM + two narrow-tile GEMMs per scale, csrc/gemm_tf32x3.cu) spends 162 us on
// 1.9 GFLOP and writes/re-reads two intermediate activations of 33 MB each.  Here a thread owns one sample of FOUR balls (two in the 16-16-32 instance): its
// coordinates, then the activations, live in registers; weights (13 KB) sit in shared memory and are read as broadcast 128-bit
// words; the ReLU'd outputs are non-negative, so the max over the ball is one redux.sync.max.u32 per output on their bit patterns.
// Only the pooled row is written.  What shapes the loops (measured): a broadcast 128-bit shared-memory load occupies the load/store
// write-back path for four cycles, the time of 16 FFMA warp-instructions.  With two balls per thread (8 FFMAs per load) the kernel ran
// at a third of the FFMA rate (68 us for the 32-32-64 scale); four balls per thread and layers 1+2 as an outer-product sweep (layer 1
// never stored, every word of W2 spent on 16 FFMAs) balance the two pipes.  Passing the weights as a __grid_constant__ parameter instead
// (constant bank, uniform datapath) was slower (119 us: a 13 KB working set streams through the constant cache every pass).
// Arithmetic is plain fp32 fma (no operand split); the bias starts each chain, k ascending.
#include "common.cuh"

namespace epnet {

constexpr float kSaGuard = 6.0e4f;  // == kF16Guard of gemm_tf32x3.cu: what an FP16-split consumer can read

template <int N1, int N2, int N3>
struct SaPack {  // float offsets into the packed weights (host: epnet_b200/gemm.py FusedFirstLevel)
    // W1: N1 rows of (wx, wy, wz, bias); W2T: N1 rows of N2 weights (k-major: row k feeds every output of layer 2); W3: N3 rows of N2
    static constexpr int W1 = 0, W2T = W1 + N1 * 4, B2 = W2T + N1 * N2, W3 = B2 + N2, B3 = W3 + N3 * N2, TOTAL = B3 + N3;
};

// R: balls per thread (a broadcast weight word feeds 4 R FFMAs).  32-32-64: R = 4, 256 threads, one CTA per SM (its 4 x 32 layer-2
// activations need ~250 registers; 68 -> 58 us against R = 2).  16-16-32: R = 2, 256 threads x 2 CTAs (one pass over the 8192 balls; R = 4 was
// slower there, 27 vs 21 us: too few warps for a kernel that short)
template <int N1, int N2, int N3, int NS, int R, int kSaThreads, int kSaMinBlocks>
__global__ void __launch_bounds__(kSaThreads, kSaMinBlocks)
sa_first_level_kernel(int n, int m, long long groups, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                      const int *__restrict__ idx, const float *__restrict__ pack, float *__restrict__ out, int ldo,
                      unsigned int *__restrict__ overflow)
{
    using P = SaPack<N1, N2, N3>;
    static_assert(N1 % 4 == 0 && N2 % 4 == 0 && N3 % 8 == 0 && N3 == 2 * NS && (NS == 16 || NS == 32), "supported shapes");
    __shared__ __align__(16) float s[P::TOTAL];
    for (int i = threadIdx.x; i < P::TOTAL; i += kSaThreads) s[i] = __ldg(pack + i);
    __syncthreads();

    constexpr int GPW = 32 / NS;  // balls per warp and row slot
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / NS, smp = lane % NS;
    const unsigned mask = NS == 32 ? 0xffffffffu : (0xffffu << (16 * sub));
    const long long stride = (long long)gridDim.x * (kSaThreads / 32) * R * GPW;
    float amax = 0.f;

    // the R samples this thread owns in the balls at g (re-centred coordinates, pointnet2_utils.py:252); a ball past the end is
    // computed on the last one and not stored
    auto fetch = [&](long long g, float (&x)[R][3]) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long gc = min(g + r * GPW + sub, groups - 1);
            const long long scene = gc / m;
            const int i = __ldg(idx + gc * NS + smp);
            const float *p = xyz + (scene * n + i) * 3, *c = new_xyz + gc * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) x[r][k] = __fsub_rn(__ldg(p + k), __ldg(c + k));
        }
    };
    long long g0 = ((long long)blockIdx.x * (kSaThreads / 32) + warp) * R * GPW;
    float x[R][3], xn[R][3];
    if (g0 < groups) fetch(g0, x);
    for (; g0 < groups; g0 += stride) {
        if (g0 + stride < groups) fetch(g0 + stride, xn);  // the next pass's gathers are in flight during this pass's arithmetic
        // layers 1 and 2 as one outer-product sweep over k: unit k of layer 1 is computed for the R balls (3 fma + ReLU each) and
        // immediately spent on all N2 accumulators of layer 2, so layer 1 is never stored and every broadcast word of W2T feeds 4R FFMAs
        float h2[R][N2];
#pragma unroll
        for (int jq = 0; jq < N2 / 4; ++jq) {
            const float4 b = *reinterpret_cast<const float4 *>(s + P::B2 + 4 * jq);
#pragma unroll
            for (int r = 0; r < R; ++r) { h2[r][4 * jq] = b.x; h2[r][4 * jq + 1] = b.y; h2[r][4 * jq + 2] = b.z; h2[r][4 * jq + 3] = b.w; }
        }
#pragma unroll 2
        for (int k = 0; k < N1; ++k) {
            const float4 w1 = *reinterpret_cast<const float4 *>(s + P::W1 + 4 * k);
            float h[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                h[r] = fmaxf(__fmaf_rn(w1.z, x[r][2], __fmaf_rn(w1.y, x[r][1], __fmaf_rn(w1.x, x[r][0], w1.w))), 0.f);
#pragma unroll
            for (int jq = 0; jq < N2 / 4; ++jq) {
                const float4 w = *reinterpret_cast<const float4 *>(s + P::W2T + k * N2 + 4 * jq);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    h2[r][4 * jq] = __fmaf_rn(w.x, h[r], h2[r][4 * jq]);
                    h2[r][4 * jq + 1] = __fmaf_rn(w.y, h[r], h2[r][4 * jq + 1]);
                    h2[r][4 * jq + 2] = __fmaf_rn(w.z, h[r], h2[r][4 * jq + 2]);
                    h2[r][4 * jq + 3] = __fmaf_rn(w.w, h[r], h2[r][4 * jq + 3]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < N2; ++j) h2[r][j] = fmaxf(h2[r][j], 0.f);
        // last layer, 8 outputs at a time; output j of a ball ends up in lane (j % NS) of the ball's lanes, slot j / NS
        uint32_t keep[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) keep[r][0] = keep[r][1] = 0u;
#pragma unroll 1
        for (int jc = 0; jc < N3; jc += 8) {
            float a[R][8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float b = s[P::B3 + jc + jj];
#pragma unroll
                for (int r = 0; r < R; ++r) a[r][jj] = b;
            }
#pragma unroll
            for (int kq = 0; kq < N2 / 4; ++kq) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 w = *reinterpret_cast<const float4 *>(s + P::W3 + (jc + jj) * N2 + 4 * kq);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        a[r][jj] = __fmaf_rn(w.x, h2[r][4 * kq], a[r][jj]);
                        a[r][jj] = __fmaf_rn(w.y, h2[r][4 * kq + 1], a[r][jj]);
                        a[r][jj] = __fmaf_rn(w.z, h2[r][4 * kq + 2], a[r][jj]);
                        a[r][jj] = __fmaf_rn(w.w, h2[r][4 * kq + 3], a[r][jj]);
                    }
                }
            }
            const int slot = jc / NS;          // uniform
            const int lane0 = jc % NS;         // outputs jc .. jc+7 belong to lanes lane0 .. lane0+7 of the ball
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    // ReLU'd values are non-negative (NaN maps above every number and trips the guard): bit patterns order like the values
                    const uint32_t u = __reduce_max_sync(mask, __float_as_uint(fmaxf(a[r][jj], 0.f)));
                    const bool mine = smp == lane0 + jj;
                    keep[r][0] = (mine && slot == 0) ? u : keep[r][0];
                    keep[r][1] = (mine && slot == 1) ? u : keep[r][1];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float v0 = __uint_as_float(keep[r][0]), v1 = __uint_as_float(keep[r][1]);
            amax = fmaxf(amax, fmaxf(v0, v1));
            if (!(v0 <= kSaGuard) || !(v1 <= kSaGuard)) amax = __int_as_float(0x7f800000);
            const long long g = g0 + r * GPW + sub;
            if (g < groups) {
                float *o = out + g * ldo;
                o[smp] = v0;
                o[NS + smp] = v1;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) x[r][k] = xn[r][k];
    }
    if (!(amax <= kSaGuard)) atomicOr(overflow, 1u);
}

template <int N1, int N2, int N3, int NS, int R, int kSaThreads, int kSaMinBlocks>
static int launch_sa(int n, int m, long long groups, const float *xyz, const float *new_xyz, const int *idx, const float *pack, float *out,
                     int ldo, cudaStream_t st)
{
    unsigned int *flag = gemm_overflow_flag();
    if (!flag) return (int)cudaErrorInvalidSymbol;
    constexpr int per_cta = (kSaThreads / 32) * R * (32 / NS);
    const long long want = (groups + per_cta - 1) / per_cta;
    // persistent: every resident CTA walks its passes with the next one's gathers in flight
    const int grid = (int)min(want, (long long)kSmCount * kSaMinBlocks);
    sa_first_level_kernel<N1, N2, N3, NS, R, kSaThreads, kSaMinBlocks><<<grid, kSaThreads, 0, st>>>(n, m, groups, xyz, new_xyz, idx, pack, out, ldo, flag);
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet

EPNET_API int epnet_sa_first_level(int b, int n, int m, int nsample, int n1, int n2, int n3, const float *xyz, const float *new_xyz,
                                   const int *idx, const float *pack, float *out, int ldo, void *stream)
{
    using namespace epnet;
    if (b < 0 || n <= 0 || m < 0 || !xyz || !new_xyz || !idx || !pack || !out || ldo < n3) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(pack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    const long long groups = (long long)b * m;
    cudaStream_t st = (cudaStream_t)stream;
    if (nsample == 32 && n1 == 32 && n2 == 32 && n3 == 64) {
        if (groups == 0) return EPNET_OK;
        return launch_sa<32, 32, 64, 32, 4, 256, 1>(n, m, groups, xyz, new_xyz, idx, pack, out, ldo, st);
    }
    if (nsample == 16 && n1 == 16 && n2 == 16 && n3 == 32) {
        if (groups == 0) return EPNET_OK;
        return launch_sa<16, 16, 32, 16, 2, 256, 2>(n, m, groups, xyz, new_xyz, idx, pack, out, ldo, st);
    }
    return EPNET_ERR_BAD_ARG;  // other widths run on the tcgen05 path
}
