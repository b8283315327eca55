// LI-Fusion point-wise image-feature gather for B200.  Replaces the torch.nn.functional.grid_sample
// call in Feature_Gather (/root/reference/lib/net/pointnet2_msg.py:107-120): bilinear interpolation,
// zero padding, grid of shape (B,1,N,2) -- i.e. one row of N sample points per scene.
//
// ATen's generic grid_sampler_2d recomputes the four tap addresses and weights for every (point,
// channel) pair.  Here a thread owns one point: tap offsets/weights are computed once and reused for
// kGgChannels channel planes; the per-channel work is 4 FMAs and one store that is coalesced across the
// warp (consecutive threads = consecutive points of the (B,C,N) output).  The taps of different points lie
// in different 32-byte sectors, so the kernel is bound by the SM's sector-request rate, not by bytes: the
// two taps of an image row are therefore fetched with ONE aligned 128-bit load (the 4-pixel group holding
// x0; only when x0 is the last pixel of its group, one lane in four, does x1 need a second, scalar load) --
// 2.5 requests per output instead of 4.  Needs W % 4 == 0 and a 16-byte aligned map; otherwise four scalar loads.
// Arithmetic follows ATen (unnormalise, floor, weights as products of corner distances, taps added in
// the order nw, ne, sw, se); results agree with grid_sample to ~1e-7, the test tolerance is 1e-5.
#include "common.cuh"

namespace epnet {

constexpr int kGgThreads = 128;
constexpr int kGgChannels = 16;

struct Taps {
    int o[4];    // plane offsets (clamped to 0 when the tap is outside; its weight is then 0)
    float w[4];  // nw, ne, sw, se
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int h, int w, int align_corners)
{
    float ix, iy;
    if (align_corners) {
        ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w - 1));
        iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h - 1));
    } else {  // ((x+1)*W - 1)/2 with the multiply-add contracted, as nvcc does for ATen's expression
        ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), (float)w, -1.f), 0.5f);
        iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), (float)h, -1.f), 0.5f);
    }
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = __fsub_rn(ix, fx), wy1 = __fsub_rn(iy, fy);  // distance to the west / north corner
    const float wx0 = __fsub_rn((float)x1, ix), wy0 = __fsub_rn((float)y1, iy);
    const bool xin0 = x0 >= 0 && x0 < w, xin1 = x1 >= 0 && x1 < w;
    const bool yin0 = y0 >= 0 && y0 < h, yin1 = y1 >= 0 && y1 < h;
    Taps t;
    t.w[0] = (xin0 && yin0) ? __fmul_rn(wx0, wy0) : 0.f;
    t.w[1] = (xin1 && yin0) ? __fmul_rn(wx1, wy0) : 0.f;
    t.w[2] = (xin0 && yin1) ? __fmul_rn(wx0, wy1) : 0.f;
    t.w[3] = (xin1 && yin1) ? __fmul_rn(wx1, wy1) : 0.f;
    t.o[0] = (xin0 && yin0) ? y0 * w + x0 : 0;
    t.o[1] = (xin1 && yin0) ? y0 * w + x1 : 0;
    t.o[2] = (xin0 && yin1) ? y1 * w + x0 : 0;
    t.o[3] = (xin1 && yin1) ? y1 * w + x1 : 0;
    return t;
}

// the two horizontally adjacent taps (x0, x0+1) of row `row` (a pointer to the start of that image row): out-of-image taps read as 0
__device__ __forceinline__ void row_pair(const float *__restrict__ row, int x0, int w, bool y_in, float &west, float &east)
{
    west = east = 0.f;
    if (!y_in) return;
    if (x0 >= 0 && x0 < w) {
        const int q = x0 & 3;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(row + (x0 - q)));
        west = q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w;
        if (q < 3) east = q == 0 ? v.y : q == 1 ? v.z : v.w;
        else if (x0 + 1 < w) east = __ldg(row + x0 + 1);
    } else if (x0 == -1) {
        east = __ldg(row);
    }
}

template <bool kVec>
__global__ void __launch_bounds__(kGgThreads)
grid_gather_kernel(int c, int h, int w, int n, const float *__restrict__ fmap, const float *__restrict__ xy, int align_corners,
                   float *__restrict__ out)
{
    const int scene = blockIdx.z;
    const int i = blockIdx.x * kGgThreads + threadIdx.x;
    if (i >= n) return;
    const size_t plane = (size_t)h * w;
    fmap += (size_t)scene * c * plane;
    out += (size_t)scene * c * n;
    const float2 g = __ldg(reinterpret_cast<const float2 *>(xy + ((size_t)scene * n + i) * 2));
    const Taps t = make_taps(g.x, g.y, h, w, align_corners);
    const int c_begin = blockIdx.y * kGgChannels;
    const int c_end = min(c, c_begin + kGgChannels);
    if (kVec) {
        float ix, iy;  // the same unnormalisation as make_taps, for the integer corner
        if (align_corners) {
            ix = __fmul_rn(__fmul_rn(__fadd_rn(g.x, 1.f), 0.5f), (float)(w - 1));
            iy = __fmul_rn(__fmul_rn(__fadd_rn(g.y, 1.f), 0.5f), (float)(h - 1));
        } else {
            ix = __fmul_rn(__fmaf_rn(__fadd_rn(g.x, 1.f), (float)w, -1.f), 0.5f);
            iy = __fmul_rn(__fmaf_rn(__fadd_rn(g.y, 1.f), (float)h, -1.f), 0.5f);
        }
        // a point far outside the image has no tap: clamp before the int conversion so that x0 + 1 cannot overflow
        const int x0 = (int)fminf(fmaxf(floorf(ix), -2.f), (float)w), y0 = (int)fminf(fmaxf(floorf(iy), -2.f), (float)h);
        const bool yin0 = y0 >= 0 && y0 < h, yin1 = y0 + 1 >= 0 && y0 + 1 < h;
        const size_t r0 = (size_t)(yin0 ? y0 : 0) * w, r1 = (size_t)(yin1 ? y0 + 1 : 0) * w;
#pragma unroll 4
        for (int ch = c_begin; ch < c_end; ++ch) {
            const float *p = fmap + (size_t)ch * plane;
            float nw, ne, sw, se;
            row_pair(p + r0, x0, w, yin0, nw, ne);
            row_pair(p + r1, x0, w, yin1, sw, se);
            float acc = __fmul_rn(nw, t.w[0]);
            acc = __fmaf_rn(ne, t.w[1], acc);
            acc = __fmaf_rn(sw, t.w[2], acc);
            acc = __fmaf_rn(se, t.w[3], acc);
            out[(size_t)ch * n + i] = acc;
        }
        return;
    }
#pragma unroll 4
    for (int ch = c_begin; ch < c_end; ++ch) {
        const float *p = fmap + (size_t)ch * plane;
        float acc = __fmul_rn(__ldg(p + t.o[0]), t.w[0]);
        acc = __fmaf_rn(__ldg(p + t.o[1]), t.w[1], acc);
        acc = __fmaf_rn(__ldg(p + t.o[2]), t.w[2], acc);
        acc = __fmaf_rn(__ldg(p + t.o[3]), t.w[3], acc);
        out[(size_t)ch * n + i] = acc;
    }
}

__global__ void __launch_bounds__(kGgThreads)
grid_gather_grad_kernel(int c, int h, int w, int n, const float *__restrict__ grad_out, const float *__restrict__ xy,
                        int align_corners, float *__restrict__ grad_fmap)
{
    const int scene = blockIdx.z;
    const int i = blockIdx.x * kGgThreads + threadIdx.x;
    if (i >= n) return;
    const size_t plane = (size_t)h * w;
    grad_fmap += (size_t)scene * c * plane;
    grad_out += (size_t)scene * c * n;
    const float2 g = __ldg(reinterpret_cast<const float2 *>(xy + ((size_t)scene * n + i) * 2));
    const Taps t = make_taps(g.x, g.y, h, w, align_corners);
    const int c_begin = blockIdx.y * kGgChannels;
    const int c_end = min(c, c_begin + kGgChannels);
    for (int ch = c_begin; ch < c_end; ++ch) {
        float *p = grad_fmap + (size_t)ch * plane;
        const float go = __ldg(grad_out + (size_t)ch * n + i);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t.w[k] != 0.f) atomicAdd(p + t.o[k], go * t.w[k]);
    }
}

}  // namespace epnet

EPNET_API int epnet_grid_gather_bilinear(int b, int c, int h, int w, int n, const float *fmap, const float *xy, int align_corners,
                                         float *out, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || h <= 0 || w <= 0 || n < 0 || !fmap || !xy || !out) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(xy) & 7) != 0) return EPNET_ERR_BAD_ARG;
    if (b == 0 || c == 0 || n == 0) return EPNET_OK;
    dim3 grid((n + kGgThreads - 1) / kGgThreads, (c + kGgChannels - 1) / kGgChannels, b);
    if (w % 4 == 0 && (reinterpret_cast<uintptr_t>(fmap) & 15) == 0)
        grid_gather_kernel<true><<<grid, kGgThreads, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, xy, align_corners, out);
    else
        grid_gather_kernel<false><<<grid, kGgThreads, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, xy, align_corners, out);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_grid_gather_bilinear_grad(int b, int c, int h, int w, int n, const float *grad_out, const float *xy,
                                              int align_corners, float *grad_fmap, void *stream)
{
    using namespace epnet;
    if (b < 0 || c < 0 || h <= 0 || w <= 0 || n < 0 || !grad_out || !xy || !grad_fmap) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(xy) & 7) != 0) return EPNET_ERR_BAD_ARG;
    if (b == 0 || c == 0 || n == 0) return EPNET_OK;
    dim3 grid((n + kGgThreads - 1) / kGgThreads, (c + kGgChannels - 1) / kGgChannels, b);
    grid_gather_grad_kernel<<<grid, kGgThreads, 0, (cudaStream_t)stream>>>(c, h, w, n, grad_out, xy, align_corners, grad_fmap);
    EPNET_RETURN_LAUNCH_STATUS();
}
