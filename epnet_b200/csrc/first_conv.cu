// The FIRST convolution of the image stream (Img_Block[0].conv1, /root/reference/lib/net/pointnet2_msg.py:17-24 with BatchNorm + ReLU folded:
// 3 -> 64 channels, 3x3, stride 1, pad 1, over the whole 384 x 1280 canvas) as a dedicated fp32 SIMT kernel.
//
// Why not the tensor-core GEMM family: with 3 input channels the contraction is K = 27 -- 4.5 GFLOP per batch of 2 against 252 MB of
// output.  As an implicit GEMM (16-byte gathers of 4-channel taps into a K = 64 tile) it measured 179 us, instruction-bound; with the
// im2col operand materialised by the preparation kernel 114 + 166 us (the 252 MB operand is written and read once more).  Here a thread
// owns two horizontally adjacent output pixels: 12 float4 taps in registers, the 27 x 64 weights in shared memory read as broadcast
// 128-bit words, 2 x 64 fp32 accumulators, plain FFMA (exact fp32 accumulation, no operand split at all), and the epilogue writes the
// result directly as the two FP16 planes the next convolution reads with TMA (x = h1 + 2^-11 h2) and/or as fp32 NHWC.
#include "common.cuh"
#include <cuda_fp16.h>

namespace epnet {

constexpr int kFcThreads = 128;
constexpr int kFcCout = 64;

__device__ __forceinline__ void fc_split2(float a, float b, uint32_t &h1, uint32_t &h2)
{
    const __half2 p = __floats2half2_rn(a, b);
    const float2 f = __half22float2(p);
    // (x - h1) is exact in fp32 and so is the scaling by 2^11
    const __half2 q = __floats2half2_rn(__fmul_rn(__fsub_rn(a, f.x), 2048.0f), __fmul_rn(__fsub_rn(b, f.y), 2048.0f));
    h1 = *reinterpret_cast<const uint32_t *>(&p);
    h2 = *reinterpret_cast<const uint32_t *>(&q);
}

__global__ void __launch_bounds__(kFcThreads)
first_conv_kernel(int b, int H, int W, const float4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, int relu,
                  float *__restrict__ y, int ldy, uint4 *__restrict__ yh1, uint4 *__restrict__ yh2, int ldh)
{
    // w_s[(tap * 3 + c) * 64 + o]: the 27 x 64 weights, K-major so that the 64 outputs of one k are contiguous
    __shared__ __align__(16) float w_s[27 * kFcCout];
    __shared__ float b_s[kFcCout];
    for (int i = threadIdx.x; i < 27 * kFcCout; i += kFcThreads) {
        const int k = i / kFcCout, o = i - k * kFcCout;
        const int tap = k / 3, c = k - 3 * tap;
        w_s[i] = __ldg(w + ((size_t)o * 9 + tap) * 4 + c);  // w: (64, 3, 3, 4) = (o, ky, kx, c padded to 4)
    }
    if (threadIdx.x < kFcCout) b_s[threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.f;
    __syncthreads();
    const int Wp = W >> 1;  // pixel pairs per row (W even)
    const long long pairs = (long long)b * H * Wp;
    for (long long t = (long long)blockIdx.x * kFcThreads + threadIdx.x; t < pairs; t += (long long)gridDim.x * kFcThreads) {
        const int xp = (int)(t % Wp);
        const long long r = t / Wp;
        const int yy = (int)(r % H), s = (int)(r / H);
        const int x0 = xp * 2;
        // the 3 x 4 input window of the pixel pair: columns x0-1 .. x0+2, rows yy-1 .. yy+1 (zeros outside the canvas)
        float4 in[3][4];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = yy + dy - 1;
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) {
                const int ix = x0 + dx - 1;
                in[dy][dx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(x + ((size_t)s * H + iy) * W + ix) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float acc0[kFcCout], acc1[kFcCout];
#pragma unroll
        for (int o = 0; o < kFcCout; ++o) acc0[o] = acc1[o] = b_s[o];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 a = in[dy][dx], c = in[dy][dx + 1];  // tap (dy, dx) of pixel 0 and of pixel 1
                const float av[3] = {a.x, a.y, a.z}, cv[3] = {c.x, c.y, c.z};
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float4 *wk = reinterpret_cast<const float4 *>(w_s + ((dy * 3 + dx) * 3 + ch) * kFcCout);
#pragma unroll
                    for (int o4 = 0; o4 < kFcCout / 4; ++o4) {
                        const float4 ww = wk[o4];  // same address in every lane: a broadcast
                        acc0[4 * o4 + 0] = __fmaf_rn(av[ch], ww.x, acc0[4 * o4 + 0]); acc1[4 * o4 + 0] = __fmaf_rn(cv[ch], ww.x, acc1[4 * o4 + 0]);
                        acc0[4 * o4 + 1] = __fmaf_rn(av[ch], ww.y, acc0[4 * o4 + 1]); acc1[4 * o4 + 1] = __fmaf_rn(cv[ch], ww.y, acc1[4 * o4 + 1]);
                        acc0[4 * o4 + 2] = __fmaf_rn(av[ch], ww.z, acc0[4 * o4 + 2]); acc1[4 * o4 + 2] = __fmaf_rn(cv[ch], ww.z, acc1[4 * o4 + 2]);
                        acc0[4 * o4 + 3] = __fmaf_rn(av[ch], ww.w, acc0[4 * o4 + 3]); acc1[4 * o4 + 3] = __fmaf_rn(cv[ch], ww.w, acc1[4 * o4 + 3]);
                    }
                }
            }
        }
        const size_t pix = ((size_t)s * H + yy) * W + x0;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float *acc = p == 0 ? acc0 : acc1;
            if (relu) {
#pragma unroll
                for (int o = 0; o < kFcCout; ++o) acc[o] = fmaxf(acc[o], 0.f);
            }
            if (y) {
                float4 *d = reinterpret_cast<float4 *>(y + (pix + p) * ldy);
#pragma unroll
                for (int o4 = 0; o4 < kFcCout / 4; ++o4) __stcs(d + o4, make_float4(acc[4 * o4], acc[4 * o4 + 1], acc[4 * o4 + 2], acc[4 * o4 + 3]));
            }
            if (yh1) {
                uint4 *d1 = yh1 + (pix + p) * (ldh >> 3), *d2 = yh2 + (pix + p) * (ldh >> 3);
#pragma unroll
                for (int j = 0; j < kFcCout / 8; ++j) {
                    uint32_t a1[4], a2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) fc_split2(acc[8 * j + 2 * e], acc[8 * j + 2 * e + 1], a1[e], a2[e]);
                    d1[j] = make_uint4(a1[0], a1[1], a1[2], a1[3]);
                    d2[j] = make_uint4(a2[0], a2[1], a2[2], a2[3]);
                }
            }
        }
    }
}

}  // namespace epnet

// x (b, H, W, 4) fp32 NHWC with channel 3 ignored; w (64, 3, 3, 4) fp32 = (o, ky, kx, c), bias (64) or NULL (BatchNorm folded by the caller);
// 3x3, stride 1, pad 1 -> y (b*H*W, ldy) fp32 and/or the FP16 planes yh1 / yh2 (b*H*W, ldh) (either may be NULL, not both).  W even.
EPNET_API int epnet_conv3x3_c3_planes(int b, int H, int W, int cout, const float *x, const float *w, const float *bias, int relu, float *y,
                                      int ldy, void *yh1, void *yh2, int ldh, void *stream)
{
    using namespace epnet;
    if (b < 0 || H <= 0 || W <= 0 || (W & 1) || cout != kFcCout || !x || !w || (!y && !yh1)) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (y && (ldy < cout || (ldy & 3) || (reinterpret_cast<uintptr_t>(y) & 15)))) return EPNET_ERR_BAD_ARG;
    if (yh1 && (!yh2 || ldh < cout || (ldh & 7) || ((reinterpret_cast<uintptr_t>(yh1) | reinterpret_cast<uintptr_t>(yh2)) & 15))) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    const long long pairs = (long long)b * H * (W / 2);
    const long long want = (pairs + kFcThreads - 1) / kFcThreads;
    const int blocks = (int)(want < (long long)kSmCount * 32 ? want : (long long)kSmCount * 32);
    first_conv_kernel<<<blocks, kFcThreads, 0, (cudaStream_t)stream>>>(b, H, W, reinterpret_cast<const float4 *>(x), w, bias, relu, y, ldy,
                                                                    reinterpret_cast<uint4 *>(yh1), reinterpret_cast<uint4 *>(yh2), ldh);
    EPNET_RETURN_LAUNCH_STATUS();
}
