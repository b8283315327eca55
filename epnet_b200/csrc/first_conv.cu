// The FIRST convolution of the image stream (Img_Block[0].conv1, /root/reference/lib/net/pointnet2_msg.py:17-24 with BatchNorm + ReLU folded:
// 3 -> 64 channels, 3x3, stride 1, pad 1, over the whole 384 x 1280 canvas) as a dedicated fp32 SIMT kernel.
//
// Why not the tensor-core GEMM family: with 3 input channels the contraction is K = 27 -- 4.5 GFLOP per batch of 2 against 252 MB of
// output.  As an implicit GEMM (16-byte gathers of 4-channel taps into a K = 64 tile) it measured 179 us, instruction-bound; with the
// im2col operand materialised by the preparation kernel 114 + 166 us (the 252 MB operand is written and read once more).
// A first SIMT version (two pixels x 64 channels per thread) measured 205 us: a broadcast 128-bit shared-memory load occupies the
// load/store write-back path for four cycles -- 16 FFMA issue slots -- and it fed only 8 FFMAs; and every thread wrote its pixels'
// 128-byte plane rows alone, one line per lane and store.  This version: a thread owns EIGHT horizontally adjacent pixels and 8 of
// the 64 output channels (the eight lanes of a group share the pixels and split the channels), so one weight word (4 channels of one
// k) feeds 32 FFMAs, the eight lanes' weight words are one 128-byte multicast shared-memory access, and a group's store of one pixel
// is one whole 128-byte row of a plane.  (Four pixels x 16 channels: 117 us; this: 111 us.)  Plain FFMA (exact fp32 accumulation,
// bias first, k ascending: ky, kx, c); the epilogue writes the two FP16 planes the next convolution reads with TMA
// (x = h1 + 2^-11 h2) and/or fp32 NHWC.
#include "common.cuh"
#include <cuda_fp16.h>

namespace epnet {

constexpr int kFcThreads = 128;
constexpr int kFcCout = 64;
constexpr int kFcPix = 8;     // pixels per thread
constexpr int kFcChunk = 8;   // output channels per thread
constexpr int kFcLanes = kFcCout / kFcChunk;  // lanes sharing a pixel group: 8

__device__ __forceinline__ void fc_split2(float a, float b, uint32_t &h1, uint32_t &h2)
{
    const __half2 p = __floats2half2_rn(a, b);
    const float2 f = __half22float2(p);
    // (x - h1) is exact in fp32 and so is the scaling by 2^11
    const __half2 q = __floats2half2_rn(__fmul_rn(__fsub_rn(a, f.x), 2048.0f), __fmul_rn(__fsub_rn(b, f.y), 2048.0f));
    h1 = *reinterpret_cast<const uint32_t *>(&p);
    h2 = *reinterpret_cast<const uint32_t *>(&q);
}

constexpr int kFcTilePix = kFcThreads / kFcLanes * kFcPix;  // 128 pixels of one image row per CTA and trip
constexpr int kFcTileCols = kFcTilePix + 2;           // + the left and right halo column

__device__ __forceinline__ void fc_cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

// grid: persistent CTAs over tiles (scene, row, 128-pixel strip).  The three input rows of the NEXT tile (3 x 130 pixels x 16 bytes) are
// copied into the other half of a double buffer with cp.async while this tile is computed: without it every thread waited for its 18
// global loads at the top of each trip (ncu: long-scoreboard stalls 1.5 per issued instruction, issue slots 54 % busy with 12 warps per SM).
__global__ void __launch_bounds__(kFcThreads, 4)
first_conv_kernel(int b, int H, int W, const float4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, int relu,
                  float *__restrict__ y, int ldy, uint4 *__restrict__ yh1, uint4 *__restrict__ yh2, int ldh, unsigned int *__restrict__ overflow)
{
    // the 27 x 64 weights, K-major; a thread (channel chunk ck of 8) owns channels 8 ck .. 8 ck + 7, i.e. 16 bytes of a plane row: the eight
    // lanes of a pixel group write one whole 128-byte row per store instruction.  w_s[k][o4][ck][e] holds channel 8 ck + 4 o4 + e: the words the
    // eight lanes read in the same instruction are adjacent -- 128 contiguous bytes per warp, one conflict-free multicast access
    __shared__ __align__(16) float w_s[27 * kFcCout];
    __shared__ __align__(16) float b_s[kFcCout];
    __shared__ __align__(16) float4 tile[2][3][kFcTileCols + 2];
    for (int i = threadIdx.x; i < 27 * kFcCout; i += kFcThreads) {
        const int k = i / kFcCout, r = i - k * kFcCout;
        const int o4 = r >> 5, ck = (r >> 2) & 7, e = r & 3;
        const int o = 8 * ck + 4 * o4 + e;
        const int tap = k / 3, c = k - 3 * tap;
        w_s[i] = __ldg(w + ((size_t)o * 9 + tap) * 4 + c);  // w: (64, 3, 3, 4) = (o, ky, kx, c padded to 4)
    }
    if (threadIdx.x < kFcCout) {
        const int r = threadIdx.x, o4 = r >> 5, ck = (r >> 2) & 7, e = r & 3;
        b_s[r] = bias ? __ldg(bias + 8 * ck + 4 * o4 + e) : 0.f;
    }
    const int chunk = threadIdx.x & (kFcLanes - 1), quad = threadIdx.x / kFcLanes;  // quad: the pixel group of this thread
    const float *wc = w_s + chunk * 4;
    const int tiles_x = (W + kFcTilePix - 1) / kFcTilePix;
    const long long total = (long long)b * H * tiles_x;

    auto issue = [&](int buf, long long t) {  // the 3 x 130 input pixels of tile t -> tile[buf]; zeros outside the canvas
        const int xt = (int)(t % tiles_x);
        const long long r = t / tiles_x;
        const int yy = (int)(r % H), sc = (int)(r / H);
        for (int i = threadIdx.x; i < 3 * kFcTileCols; i += kFcThreads) {
            const int dy = i / kFcTileCols, dx = i - dy * kFcTileCols;
            const int iy = yy + dy - 1, ix = xt * kFcTilePix + dx - 1;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) fc_cp_async16(&tile[buf][dy][dx], x + ((size_t)sc * H + iy) * W + ix);
            else tile[buf][dy][dx] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    long long t = blockIdx.x;
    if (t < total) issue(0, t);
    asm volatile("cp.async.commit_group;" ::: "memory");
    float amax = 0.f;
    for (int it = 0; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (t + gridDim.x < total) issue(buf ^ 1, t + gridDim.x);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // this tile's copies (the older group) have landed
        __syncthreads();                                        // ... for every thread, and the weights on the first trip
        const int xt = (int)(t % tiles_x);
        const long long r = t / tiles_x;
        const int yy = (int)(r % H), s = (int)(r / H);
        const int x0 = xt * kFcTilePix + quad * kFcPix;
        if (x0 < W) {
        float in[3][kFcPix + 2][3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < kFcPix + 2; ++dx) {
                const float4 v = tile[buf][dy][quad * kFcPix + dx];  // the four lanes of a quad read the same word; channel 3 is dropped
                in[dy][dx][0] = v.x; in[dy][dx][1] = v.y; in[dy][dx][2] = v.z;
            }
        float acc[kFcPix][kFcChunk];
#pragma unroll
        for (int o4 = 0; o4 < kFcChunk / 4; ++o4) {
            const float4 bb = *reinterpret_cast<const float4 *>(b_s + (o4 * kFcLanes + chunk) * 4);
#pragma unroll
            for (int p = 0; p < kFcPix; ++p) { acc[p][4 * o4] = bb.x; acc[p][4 * o4 + 1] = bb.y; acc[p][4 * o4 + 2] = bb.z; acc[p][4 * o4 + 3] = bb.w; }
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float4 *wk = reinterpret_cast<const float4 *>(wc + ((dy * 3 + dx) * 3 + ch) * kFcCout);
#pragma unroll
                    for (int o4 = 0; o4 < kFcChunk / 4; ++o4) {
                        const float4 ww = wk[kFcLanes * o4];  // adjacent words per warp (one per channel chunk): one multicast access
#pragma unroll
                        for (int p = 0; p < kFcPix; ++p) {
                            const float a = in[dy][p + dx][ch];  // tap (dy, dx) of pixel p
                            acc[p][4 * o4 + 0] = __fmaf_rn(a, ww.x, acc[p][4 * o4 + 0]);
                            acc[p][4 * o4 + 1] = __fmaf_rn(a, ww.y, acc[p][4 * o4 + 1]);
                            acc[p][4 * o4 + 2] = __fmaf_rn(a, ww.z, acc[p][4 * o4 + 2]);
                            acc[p][4 * o4 + 3] = __fmaf_rn(a, ww.w, acc[p][4 * o4 + 3]);
                        }
                    }
                }
            }
        }
        // acc[p][0..7] = channels 8 chunk .. 8 chunk + 7
        const size_t pix = ((size_t)s * H + yy) * W + x0;
#pragma unroll
        for (int p = 0; p < kFcPix; ++p) {
            if (x0 + p >= W) break;  // W is a multiple of 4, a pixel group has 8: the last group of a row may be half empty
#pragma unroll
            for (int o = 0; o < kFcChunk; ++o) {
                if (relu) acc[p][o] = fmaxf(acc[p][o], 0.f);
                amax = fmaxf(amax, fabsf(acc[p][o]));
            }
            if (y) {
                float4 *d = reinterpret_cast<float4 *>(y + (pix + p) * ldy + 8 * chunk);
                __stcs(d, make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
                __stcs(d + 1, make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]));
            }
            if (yh1) {  // 8 channels = 16 bytes per plane: the eight lanes of the group write the whole 128-byte row together
                uint32_t a1[4], a2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) fc_split2(acc[p][2 * e], acc[p][2 * e + 1], a1[e], a2[e]);
                yh1[(pix + p) * (ldh >> 3) + chunk] = make_uint4(a1[0], a1[1], a1[2], a1[3]);
                yh2[(pix + p) * (ldh >> 3) + chunk] = make_uint4(a2[0], a2[1], a2[2], a2[3]);
            }
        }
        }
        __syncthreads();  // every thread is done with tile[buf] before the next trip refills it
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (!(amax <= 6.0e4f)) atomicOr(overflow, 1u);  // the FP16 range guard of the GEMM epilogues (gemm_tf32x3.cu: kF16Guard); NaN/inf too
}

}  // namespace epnet

// x (b, H, W, 4) fp32 NHWC with channel 3 ignored; w (64, 3, 3, 4) fp32 = (o, ky, kx, c), bias (64) or NULL (BatchNorm folded by the caller);
// 3x3, stride 1, pad 1 -> y (b*H*W, ldy) fp32 and/or the FP16 planes yh1 / yh2 (b*H*W, ldh) (either may be NULL, not both).  W % 4 == 0.
EPNET_API int epnet_conv3x3_c3_planes(int b, int H, int W, int cout, const float *x, const float *w, const float *bias, int relu, float *y,
                                      int ldy, void *yh1, void *yh2, int ldh, void *stream)
{
    using namespace epnet;
    if (b < 0 || H <= 0 || W <= 0 || (W % 4) || cout != kFcCout || !x || !w || (!y && !yh1)) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (y && (ldy < cout || (ldy & 3) || (reinterpret_cast<uintptr_t>(y) & 15)))) return EPNET_ERR_BAD_ARG;
    if (yh1 && (!yh2 || ldh < cout || (ldh & 7) || ((reinterpret_cast<uintptr_t>(yh1) | reinterpret_cast<uintptr_t>(yh2)) & 15))) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    unsigned int *flag = gemm_overflow_flag();
    if (!flag) return (int)cudaErrorInvalidSymbol;
    const long long want = (long long)b * H * ((W + kFcTilePix - 1) / kFcTilePix);
    const int blocks = (int)(want < (long long)kSmCount * 4 ? want : (long long)kSmCount * 4);  // persistent: four CTAs per SM
    first_conv_kernel<<<blocks, kFcThreads, 0, (cudaStream_t)stream>>>(b, H, W, reinterpret_cast<const float4 *>(x), w, bias, relu, y, ldy,
                                                                    reinterpret_cast<uint4 *>(yh1), reinterpret_cast<uint4 *>(yh2), ldh, flag);
    EPNET_RETURN_LAUNCH_STATUS();
}
