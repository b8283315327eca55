// Sparse evaluation of the final image fusion.
//
// The reference (lib/net/pointnet2_msg.py:237-246) up-samples the four image-stream maps with ConvTranspose2d (kernel == stride
// = 2, 4, 8, 16) to full resolution, concatenates them (64 x 384 x 1280 per scene), runs a 1x1 conv + BN + ReLU over ALL 491520
// pixels and then bilinearly samples the result at the 16384 points: 4 * 16384 taps, 13 % of the pixels it computed.  Here only
// the taps are computed.  A transposed convolution with kernel == stride is, per output pixel (Y, X), a matrix-vector product of
// the input pixel (Y / k, X / k) with the weight slice of phase (Y % k, X % k); pixels of equal phase share weights.  So the taps
// are counting-sorted by their phase at the coarsest kernel (k_max = 16: 256 bins; every finer kernel's phase is a function of
// it), each bin padded to a multiple of 128 rows, and every level becomes ONE row-gather GEMM (epnet_gemm_tf32x3_rows) whose
// 128-row tiles pick their weight slice by the tile's phase.  The 1x1 conv + ReLU runs on the sorted rows, and a last kernel
// blends each point's four taps with the bilinear weights -- the same weights, tap order and fma chain as the dense gather.
//
// Kernels here: taps (per point: tap pixels, weights, phase histogram), plan (one CTA: padded bin offsets, per-tile phase, tile
// count), scatter (slot -> sorted position; per level the row of the input map each sorted row reads), blend.
#include "common.cuh"

namespace epnet {

constexpr int kTailPhases = 256;  // (Y % 16, X % 16)

// identical arithmetic to pm_taps of fused.cu (grid_sampler_2d, zeros padding): tap pixel offsets inside the scene and weights
__device__ __forceinline__ void tail_taps(float gx, float gy, int h, int w, int align_corners, int *o, float *wt)
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

// one thread per point: its four taps -> tap_pix[4p+t] = scene * H * W + pixel, tap_w[4p+t]; phase histogram (shared, then global)
__global__ void __launch_bounds__(256)
tail_taps_kernel(int points, int n, int H, int W, int align_corners, const float *__restrict__ xy, int *__restrict__ tap_pix,
                 float *__restrict__ tap_w, int *__restrict__ hist)
{
    __shared__ int h_s[kTailPhases];
    h_s[threadIdx.x] = 0;
    __syncthreads();
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p < points) {
        const float2 g = __ldg(reinterpret_cast<const float2 *>(xy) + p);
        int o[4];
        float wt[4];
        tail_taps(g.x, g.y, H, W, align_corners, o, wt);
        const int scene = p / n;
        int4 pix;
        int *pp = &pix.x;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int Y = o[t] / W, X = o[t] - Y * W;
            atomicAdd(&h_s[(Y & 15) * 16 + (X & 15)], 1);
            pp[t] = scene * H * W + o[t];
        }
        reinterpret_cast<int4 *>(tap_pix)[p] = pix;
        reinterpret_cast<float4 *>(tap_w)[p] = make_float4(wt[0], wt[1], wt[2], wt[3]);
    }
    __syncthreads();
    if (h_s[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h_s[threadIdx.x]);
}

// one CTA of 256 threads: bins padded to multiples of 128 rows, exclusive scan -> start[bin]; tile_phase[tile] for every tile;
// *n_tiles; cursor[bin] = 0
__global__ void __launch_bounds__(kTailPhases)
tail_plan_kernel(const int *__restrict__ hist, int *__restrict__ start, int *__restrict__ cursor, int *__restrict__ tile_phase,
                 int *__restrict__ n_tiles, int max_tiles)
{
    __shared__ int scan[kTailPhases];
    const int t = threadIdx.x;
    const int tiles = (hist[t] + 127) >> 7;
    scan[t] = tiles;
    __syncthreads();
    for (int d = 1; d < kTailPhases; d <<= 1) {
        const int v = t >= d ? scan[t - d] : 0;
        __syncthreads();
        scan[t] += v;
        __syncthreads();
    }
    const int first = scan[t] - tiles;  // exclusive prefix, in tiles
    start[t] = first << 7;
    cursor[t] = 0;
    for (int j = 0; j < tiles; ++j)
        if (first + j < max_tiles) tile_phase[first + j] = t;
    if (t == kTailPhases - 1) *n_tiles = min(scan[t], max_tiles);
}

// one thread per tap slot: its position in the phase-sorted order, and per level the input-map row that sorted row reads
struct TailLevels {
    int levels;
    int k[4], h[4], w[4];  // kernel (= stride) and input-map size per level
};
__global__ void __launch_bounds__(256)
tail_scatter_kernel(int slots, int H, int W, TailLevels lv, const int *__restrict__ tap_pix, const int *__restrict__ start,
                    int *__restrict__ cursor, int *__restrict__ pos_of_slot, int *__restrict__ row_idx, int row_stride)
{
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= slots) return;
    const int gp = __ldg(tap_pix + s);
    const int scene = gp / (H * W), pix = gp - scene * (H * W);
    const int Y = pix / W, X = pix - Y * W;
    const int ph = (Y & 15) * 16 + (X & 15);
    const int pos = __ldg(start + ph) + atomicAdd(cursor + ph, 1);
    pos_of_slot[s] = pos;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i < lv.levels) row_idx[(size_t)i * row_stride + pos] = (scene * lv.h[i] + Y / lv.k[i]) * lv.w[i] + X / lv.k[i];
}

// out[p][0..c) = sum_t w[4p+t] * F[pos[4p+t]][0..c): 2^gs lanes per point walk 16-byte chunks; tap order and fma chain of the dense gather
__global__ void __launch_bounds__(256)
tail_blend_kernel(int points, int c, const float *__restrict__ F, int ldf, const int *__restrict__ pos_of_slot, const float *__restrict__ tap_w,
                  float *__restrict__ out, int ldo, int gs)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const int p = (int)(t >> gs), sub = (int)(t & ((1 << gs) - 1));
    if (p >= points) return;
    const int4 pos = __ldg(reinterpret_cast<const int4 *>(pos_of_slot) + p);
    const float4 w = __ldg(reinterpret_cast<const float4 *>(tap_w) + p);
    const float *t0 = F + (size_t)pos.x * ldf, *t1 = F + (size_t)pos.y * ldf, *t2 = F + (size_t)pos.z * ldf, *t3 = F + (size_t)pos.w * ldf;
    for (int k0 = sub * 4; k0 < c; k0 += 4 << gs) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(t0 + k0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(t1 + k0));
        const float4 cc = __ldg(reinterpret_cast<const float4 *>(t2 + k0));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(t3 + k0));
        float4 r;
        r.x = __fmaf_rn(d.x, w.w, __fmaf_rn(cc.x, w.z, __fmaf_rn(b.x, w.y, __fmul_rn(a.x, w.x))));
        r.y = __fmaf_rn(d.y, w.w, __fmaf_rn(cc.y, w.z, __fmaf_rn(b.y, w.y, __fmul_rn(a.y, w.x))));
        r.z = __fmaf_rn(d.z, w.w, __fmaf_rn(cc.z, w.z, __fmaf_rn(b.z, w.y, __fmul_rn(a.z, w.x))));
        r.w = __fmaf_rn(d.w, w.w, __fmaf_rn(cc.w, w.z, __fmaf_rn(b.w, w.y, __fmul_rn(a.w, w.x))));
        *reinterpret_cast<float4 *>(out + (size_t)p * ldo + k0) = r;
    }
}

}  // namespace epnet

// xy (b, n, 2) normalised to [-1, 1] -> per point four taps of the H x W canvas: tap_pix (b*n*4) = scene*H*W + y*W + x (0 within the
// scene for taps outside the canvas, whose weight is 0), tap_w (b*n*4); hist (256) += taps per phase (Y % 16) * 16 + (X % 16)
// (the caller zeroes hist).  16-byte aligned buffers.
EPNET_API int epnet_tail_taps(int b, int n, int H, int W, int align_corners, const float *xy, int *tap_pix, float *tap_w, int *hist, void *stream)
{
    using namespace epnet;
    if (b < 0 || n < 0 || H <= 0 || W <= 0 || !xy || !tap_pix || !tap_w || !hist || (long long)b * H * W >= (1ll << 31)) return EPNET_ERR_BAD_ARG;
    if (((reinterpret_cast<uintptr_t>(tap_pix) | reinterpret_cast<uintptr_t>(tap_w)) & 15) || (reinterpret_cast<uintptr_t>(xy) & 7)) return EPNET_ERR_BAD_ARG;
    const int points = b * n;
    if (points == 0) return EPNET_OK;
    tail_taps_kernel<<<(points + 255) / 256, 256, 0, (cudaStream_t)stream>>>(points, n, H, W, align_corners, xy, tap_pix, tap_w, hist);
    EPNET_RETURN_LAUNCH_STATUS();
}

// hist (256) -> start (256): first sorted row of each phase (bins padded to multiples of 128 rows); cursor (256) = 0;
// tile_phase (max_tiles): phase of every 128-row tile; *n_tiles = tiles in use (<= max_tiles)
EPNET_API int epnet_tail_plan(const int *hist, int *start, int *cursor, int *tile_phase, int *n_tiles, int max_tiles, void *stream)
{
    using namespace epnet;
    if (!hist || !start || !cursor || !tile_phase || !n_tiles || max_tiles <= 0) return EPNET_ERR_BAD_ARG;
    tail_plan_kernel<<<1, kTailPhases, 0, (cudaStream_t)stream>>>(hist, start, cursor, tile_phase, n_tiles, max_tiles);
    EPNET_RETURN_LAUNCH_STATUS();
}

// slots = taps; levels <= 4 transposed convolutions with kernel == stride k[i] (each dividing 16) over input maps h[i] x w[i]
// (h[i]*k[i] == H, w[i]*k[i] == W): pos_of_slot (slots) = sorted row of every tap; row_idx (levels, row_stride): for sorted row r of
// level i the row of the (b*h[i]*w[i], C) input map it reads.  Rows never written (bin padding) keep the caller's zeros.
EPNET_API int epnet_tail_scatter(int slots, int H, int W, int levels, const int *k, const int *h, const int *w, const int *tap_pix,
                                 const int *start, int *cursor, int *pos_of_slot, int *row_idx, int row_stride, void *stream)
{
    using namespace epnet;
    if (slots < 0 || H <= 0 || W <= 0 || levels < 1 || levels > 4 || !k || !h || !w || !tap_pix || !start || !cursor || !pos_of_slot || !row_idx)
        return EPNET_ERR_BAD_ARG;
    TailLevels lv = {};
    lv.levels = levels;
    for (int i = 0; i < levels; ++i) {
        if (k[i] < 1 || k[i] > 16 || (16 % k[i]) != 0 || h[i] * k[i] != H || w[i] * k[i] != W) return EPNET_ERR_BAD_ARG;
        lv.k[i] = k[i]; lv.h[i] = h[i]; lv.w[i] = w[i];
    }
    if (slots == 0) return EPNET_OK;
    tail_scatter_kernel<<<(slots + 255) / 256, 256, 0, (cudaStream_t)stream>>>(slots, H, W, lv, tap_pix, start, cursor, pos_of_slot, row_idx, row_stride);
    EPNET_RETURN_LAUNCH_STATUS();
}

// out (points, ldo)[0..c) = bilinear blend of the four sorted rows of F (rows, ldf) of every point; c % 4 == 0
EPNET_API int epnet_tail_blend(int points, int c, const float *F, int ldf, const int *pos_of_slot, const float *tap_w, float *out, int ldo,
                               void *stream)
{
    using namespace epnet;
    if (points < 0 || c <= 0 || (c & 3) || !F || !pos_of_slot || !tap_w || !out || ldf < c || (ldf & 3) || ldo < c || (ldo & 3)) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(pos_of_slot) | reinterpret_cast<uintptr_t>(tap_w)) & 15)
        return EPNET_ERR_BAD_ARG;
    if (points == 0) return EPNET_OK;
    int gs = 0;
    while ((4 << gs) < c && gs < 3) ++gs;
    const long long threads = (long long)points << gs;
    tail_blend_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, c, F, ldf, pos_of_slot, tap_w, out, ldo, gs);
    EPNET_RETURN_LAUNCH_STATUS();
}
