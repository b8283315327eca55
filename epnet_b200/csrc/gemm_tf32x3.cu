// Shared-MLP layer as a tcgen05 / TMEM GEMM with fp32-grade accuracy ("3xTF32").
//
//     Y[l][n] = act( sum_k X[l][k] * W[n][k] + bias[n] )        X: (L, K) point-major rows, W: (N, K)
//
// Replaces the 1x1 Conv2d + BatchNorm(eval) + ReLU units of SharedMLP (reference
// pointnet2_lib/pointnet2/pytorch_utils.py:20-32; cuDNN SIMT fp32 there).  BASELINE.json asks for 1e-5
// agreement with fp32, which plain TF32 (10-bit mantissa) cannot give, and tcgen05 has no fp32 MMA; so each
// fp32 operand is split x = hi + lo with hi = x with the 13 low mantissa bits cleared (exactly representable
// in TF32) and lo = x - hi (exact in fp32), and three MMAs accumulate hi*hi + hi*lo + lo*hi in the fp32 TMEM
// accumulator (the dropped lo*lo term is ~2^-22 relative).
//
// Mapping: UMMA M = 128 rows of X (points or (centre,sample) pairs) = the 128 TMEM lanes; UMMA N = a tile of
// output channels (<= 256 TMEM columns); K is walked in blocks of 32 (= one 128-byte swizzle row of TF32).
//   * W is split and laid out ONCE on the host side (pack_weights in epnet_b200/gemm.py) in exactly the shared
//     memory image the MMA wants (K-major, SWIZZLE_128B, 8-row atoms), so a k-block of it is one contiguous
//     1-D bulk copy (cp.async.bulk + mbarrier complete_tx) -- no tensor map, no SIMT work;
//   * X is split on the fly: 4 producer warps read 128-bit chunks (8 lanes = one 128-byte row segment), form
//     hi/lo, and store them with the same swizzle (conflict-free: the 8 lanes of a row cover its 8 chunks);
//   * one elected thread issues the 12 tcgen05.mma (3 terms x 4 k-steps of 8) per k-block and commits to the
//     stage's "empty" mbarrier; after the last k-block it commits to the accumulator barrier;
//   * the producer warps then become the epilogue: tcgen05.ld (32 lanes x 16 columns per warp per step), bias,
//     ReLU, 128-bit stores; with pool > 1 the max over `pool` consecutive rows (the nsample axis of a grouped
//     tensor, F.max_pool2d in pointnet2_modules.py:59-61) is taken across lanes with redux.sync before storing.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

namespace epnet {

constexpr int kGmBM = 128;       // rows per CTA tile (TMEM lanes)
constexpr int kGmBK = 32;        // k-block: 32 tf32 = 128 bytes per row
constexpr int kGmProducers = 128;
constexpr int kGmThreads = 192;  // 4 producer/epilogue warps + MMA warp + weight-loader warp
constexpr int kGmMaxStages = 4;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start address >> 4, [16,30) leading byte offset >> 4 (1: unused for swizzled K-major),
// [32,46) stride byte offset >> 4 (1024 B between 8-row atoms), [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr)
{
    uint64_t d = (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, single CTA.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive columns of the accumulator -> 32 registers per thread (thread = lane = tile row); no wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

struct GemmArgs {
    const float *x;        // (L, ldx) rows; columns >= K are never read
    const float *wpack;    // packed weights: [n_tile][k_block][hi|lo][BN rows * 128 B]
    const float *bias;     // (N) or null
    float *y;              // (L / pool, ldy)
    int L, K, N;           // logical sizes
    int ldx, ldy;
    int BN;                // columns per n-tile (multiple of 16, <= 256)
    int n_kblocks;
    int relu;
    int pool;              // 1, or the number of consecutive rows pooled by max: 2, 4, 8, 16 or 32
    int stages;
    int depth;             // A-from-TMEM kernel: k-blocks of X in flight per warp (cp.async ring slots)
    int x_vec_ok;          // x base 16-byte aligned and ldx % 4 == 0
    // implicit 3x3 convolution (pad 1): x is an NHWC image (B, H, W, Cin), a GEMM row is an output pixel (b, yo, xo),
    // k = (ky*3 + kx)*Cin + c; Cin is a power of two >= 4, so a 16-byte chunk never straddles two taps.
    int conv;              // 0: plain rows, 1: conv gather
    int H, W, cin_shift, stride, Ho, Wo;
    // transposed convolution with kernel == stride == dk: a GEMM row is an input pixel (b, yi, xi) of an (H, W) map, column
    // n = (ky*dk + kx)*dco + o; the epilogue writes the dco channels to output pixel (yi*dk+ky, xi*dk+kx) of the NHWC buffer y
    // (pixel stride ldy) instead of to row-major (L, ldy).  dco % 4 == 0.
    int dk, dco;
    // channel-major output: rows are (scene, point) pairs with `tr` points per scene and y is (scenes, N, tr) -- the interface
    // layout (B, C, N) of the reference -- written straight from the epilogue (lanes = consecutive points: coalesced per column)
    int tr;
    // grouped-gather A operand (first shared-MLP layer of a set-abstraction scale; narrow-tile kernel only): GEMM row r is the
    // (scene, centre, sample) triple r, its columns are [feats[scene, g_idx[r], 0..g_c) | xyz[scene, g_idx[r]] - centre | 0...]:
    // the grouped tensor of QueryAndGroup (pointnet2_utils.py:241-264) is never materialised.  x = feats (point-major rows of
    // ldx floats, NULL when g_c == 0), K = g_c + 3, kcopy = g_c columns come from x.
    const int *g_idx;
    const float *g_xyz, *g_centre;
    int g_n, g_ns, g_rows_scene;
    int kcopy;             // columns of a row that are read from x (== K unless grouped)
    int f16;               // weights are packed as FP16 planes (k-blocks of 64): wide-tile FP16-split kernel
    float corr_scale;      // the correction accumulator is multiplied by this in the epilogue (1 for the TF32 split, 2^-11 for FP16)
    // Optional second output of the plain-rows epilogue: the result already split into the two FP16 planes the next FP16-split
    // layer consumes (h1 = fp16(v), h2 = fp16((v - h1) * 2^11)), rows ldh halfs apart.  A consumer that reads the planes with TMA
    // (gemm_f16x3_tma_kernel) does no conversion work at all: every activation is split ONCE, by its producer, instead of once per
    // tap and column tile by its consumers.  y may then be NULL (planes only).
    __half *yh1, *yh2;
    int ldh;
    // gemm_f16x3_tma_kernel, convolution mode: a CTA's 128 rows are a th x tw patch of output pixels (tw * th == 128), patches in
    // raster order over (image, patch row, patch column); 0 = rows are linear
    int tw, th, tiles_x, tiles_y;
    // Narrow-tile kernel, sparse image tail (csrc/sparse_tail.cu): rows gathered by g_idx without offset columns (g_xyz == NULL);
    // the number of 128-row tiles actually present is read from the device (*m_tiles_dev <= the tiles of L); the weight set of an
    // m-tile is selected by its transposed-convolution phase: tile_phase[mt] = (Y % 16) * 16 + (X % 16) of the tile's pixels, weight
    // n-tile ((ph >> 4) % phase_k) * phase_k + ((ph & 15) % phase_k).
    const int *m_tiles_dev;
    const int *tile_phase;
    int phase_k;
};

// 16-byte global -> shared copy without register staging; bytes beyond src_bytes (0..16) are written as zero
__device__ __forceinline__ void cp_async16_cg(void *dst_smem, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(void *dst_smem, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}

// The 8 rows of X a producer thread serves (first, first + step, ...), resolved once per tile: plain rows -> pointer to the row,
// mask = row exists; conv -> pointer to the top-left tap (dy = dx = -1; only dereferenced where the mask allows) and a 9-bit
// mask of the taps that fall inside the image.  Per k-block only a thread-uniform offset is added, so the k loop carries no
// per-row index arithmetic.
struct RowSource {
    const float *rowp[8];
    uint32_t rmask[8];

    __device__ __forceinline__ void init(const GemmArgs &a, int first, int step)
    {
        if (a.conv) {
            int row = first;
            int b = row / (a.Ho * a.Wo);
            const int rem = row - b * (a.Ho * a.Wo);
            int yo = rem / a.Wo, xo = rem - yo * a.Wo;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int y0 = yo * a.stride - 1, x0 = xo * a.stride - 1;
                uint32_t vy = 0, vx = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    vy |= (uint32_t)(y0 + d >= 0 && y0 + d < a.H) << d;
                    vx |= (uint32_t)(x0 + d >= 0 && x0 + d < a.W) << d;
                }
                const uint32_t m = ((vy & 1u) ? vx : 0u) | ((vy & 2u) ? vx << 3 : 0u) | ((vy & 4u) ? vx << 6 : 0u);
                rmask[i] = row < a.L ? m : 0u;
                rowp[i] = a.x + (((long long)(b * a.H + y0) * a.W + x0) << a.cin_shift);
                row += step;  // the next row of this thread is `step` output pixels further along the scan
                xo += step;
                while (xo >= a.Wo) { xo -= a.Wo; ++yo; }
                while (yo >= a.Ho) { yo -= a.Ho; ++b; }
            }
        } else if (a.g_idx) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = first + step * i;
                const bool ok = row < a.L && a.kcopy > 0;
                rmask[i] = ok ? 1u : 0u;
                rowp[i] = ok ? a.x + ((size_t)(row / a.g_rows_scene) * a.g_n + __ldg(a.g_idx + row)) * a.ldx : a.x;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = first + step * i;
                rmask[i] = row < a.L ? 1u : 0u;
                rowp[i] = a.x + (size_t)row * a.ldx;
            }
        }
    }

    // this thread's 16-byte chunk (4 floats from column k0) of each of its rows; zeros where nothing exists
    __device__ __forceinline__ void load(const GemmArgs &a, int k0, float4 (&v)[8]) const
    {
        if (a.conv) {
            const int tap = k0 >> a.cin_shift;            // (ky*3 + kx); >= 9 in the zero padding of the last k-block
            const int ky = tap / 3, kx = tap - 3 * ky;
            const int off = ((ky * a.W + kx) << a.cin_shift) + (k0 & ((1 << a.cin_shift) - 1));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((rmask[i] >> tap) & 1u) v[i] = __ldg(reinterpret_cast<const float4 *>(rowp[i] + off));
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rmask[i] && k0 < a.kcopy) {
                const float *src = rowp[i] + k0;
                if (a.x_vec_ok && k0 + 4 <= a.kcopy) {
                    v[i] = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    v[i].x = __ldg(src + 0);
                    if (k0 + 1 < a.kcopy) v[i].y = __ldg(src + 1);
                    if (k0 + 2 < a.kcopy) v[i].z = __ldg(src + 2);
                    if (k0 + 3 < a.kcopy) v[i].w = __ldg(src + 3);
                }
            }
        }
    }

    // 8 consecutive floats from column k0 of this thread's first four rows (two 16-byte chunks each)
    __device__ __forceinline__ void load4x8(const GemmArgs &a, int k0, float4 (&v)[8]) const
    {
        float4 lo[8], hi[8];
        load(a, k0, lo);      // rows 4..7 of the source are not used by the caller: their loads are predicated off below
        load(a, k0 + 4, hi);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = lo[i]; v[2 * i + 1] = hi[i]; }
    }

    // the same chunks, copied asynchronously to base + dst_off[i] (cp.async group of the caller); rows whose base or stride
    // is not 16-byte aligned are loaded and stored synchronously instead
    __device__ __forceinline__ void copy_async(const GemmArgs &a, int k0, uint8_t *base, const uint32_t (&dst_off)[8]) const
    {
        if (a.conv) {
            const int tap = k0 >> a.cin_shift;
            const int ky = tap / 3, kx = tap - 3 * ky;
            const int off = ((ky * a.W + kx) << a.cin_shift) + (k0 & ((1 << a.cin_shift) - 1));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool ok = (rmask[i] >> tap) & 1u;
                cp_async16_ca(base + dst_off[i], ok ? rowp[i] + off : a.wpack, ok ? 16u : 0u);  // neighbouring pixels share taps: keep L1
            }
            return;
        }
        if (a.x_vec_ok) {
            const int valid = min(4, a.kcopy - k0);  // floats of this chunk that exist in x (<= 0: none)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool ok = rmask[i] && valid > 0;
                cp_async16_cg(base + dst_off[i], ok ? rowp[i] + k0 : a.wpack, ok ? (uint32_t)valid * 4u : 0u);  // wpack: a valid dummy address
            }
            return;
        }
        float4 v[8];
        load(a, k0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4 *>(base + dst_off[i]) = v[i];
    }
};

__device__ __forceinline__ void split_f16(float x, __half &h1, __half &h2)
{
    h1 = __float2half_rn(x);
    h2 = __float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(h1)), 2048.0f));
}

// Range guard of the FP16 operand split: gemm_f16x3_kernel needs |x| < 65504.  Every value such a kernel can read is bounded by
// a value some GEMM epilogue of this file wrote (activations of the previous layer; interpolation, bilinear gathering, max-pooling
// and the sigmoid attention scale are convex combinations or contractions of them), so every epilogue raises this per-device flag
// when it writes a magnitude above kF16Guard (or a non-finite value).  The host reads it after the forward
// (epnet_gemm_overflow_read) and re-runs on the TF32 split, whose range is fp32's (epnet_b200/runner.py).
__device__ unsigned int g_gemm_overflow = 0u;
constexpr float kF16Guard = 6.0e4f;

// Accumulator -> global memory: main + correction, bias, ReLU, then plain rows / pooled rows / the transposed convolution's
// patch scatter.  Called by the four warps whose warp index selects the TMEM lane quarter (thread = lane = tile row).
__device__ __forceinline__ void gemm_epilogue(const GemmArgs &a, uint32_t tmem_acc, uint32_t corr_off, int warp, int lane, int row0,
                                              int ntile, const float *bias_s, int c_first = 0, int c_step = 32, int row_override = -1)
{
    const int BN = a.BN;
    const int r = warp * 32 + lane;  // TMEM lane == tile row
    const int row = row_override >= 0 ? row_override : row0 + r;  // override: the caller maps tile rows to output rows itself
    const int pool = a.pool;
    float amax = 0.f;
    for (int c0 = c_first; c0 < BN; c0 += c_step) {
        uint32_t v[32], w[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + corr_off, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int ncols = min(32, BN - c0);
        const int n0 = ntile * BN + c0;
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c0 + j4);  // c0 is a multiple of 32, bias_s 16-byte aligned
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j4 + e;
                float f = __fmaf_rn(__uint_as_float(w[j]), a.corr_scale, __uint_as_float(v[j])) + bb[e];  // scale 1: an exact add
                if (a.relu) f = fmaxf(f, 0.f);
                if (!(a.relu && pool > 1)) amax = fmaxf(amax, fabsf(f));  // pooled after ReLU: the guard looks at the pooled maxima instead
                v[j] = __float_as_uint(f);
            }
        }
        if (a.tr) {
            if (row < a.L) {
                const int scene = row / a.tr, pt = row - scene * a.tr;
                float *dst = a.y + ((size_t)scene * a.N + n0) * a.tr + pt;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < ncols && n0 + j < a.N) dst[(size_t)j * a.tr] = __uint_as_float(v[j]);
            }
        } else if (a.dk) {
            if (row < a.L) {
                const int byi = row / a.W, xi = row - byi * a.W;
                const size_t wo = (size_t)a.W * a.dk;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = n0 + j;
                    if (j < ncols && n < a.N) {
                        const int q = n / a.dco, o = n - q * a.dco;
                        const int ky = q / a.dk, kx = q - ky * a.dk;
                        float *dst = a.y + (((size_t)byi * a.dk + ky) * wo + (size_t)xi * a.dk + kx) * a.ldy + o;
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
        } else if (pool == 1) {
            if (row < a.L && a.yh1) {  // FP16 planes for the next layer (n0, ncols, ldh multiples of 8: checked by the host entry)
                __half *d1 = a.yh1 + (size_t)row * a.ldh + n0, *d2 = a.yh2 + (size_t)row * a.ldh + n0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (j < ncols) {
                        uint32_t q1[4], q2[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {  // two values per packed conversion; (x - h1) and the scaling by 2^11 are exact in fp32
                            const float x0 = __uint_as_float(v[j + 2 * e]), x1 = __uint_as_float(v[j + 2 * e + 1]);
                            const __half2 p1 = __floats2half2_rn(x0, x1);
                            const float2 f1 = __half22float2(p1);
                            const __half2 p2 = __floats2half2_rn(__fmul_rn(__fsub_rn(x0, f1.x), 2048.0f), __fmul_rn(__fsub_rn(x1, f1.y), 2048.0f));
                            q1[e] = *reinterpret_cast<const uint32_t *>(&p1);
                            q2[e] = *reinterpret_cast<const uint32_t *>(&p2);
                        }
                        *reinterpret_cast<uint4 *>(d1 + j) = make_uint4(q1[0], q1[1], q1[2], q1[3]);
                        *reinterpret_cast<uint4 *>(d2 + j) = make_uint4(q2[0], q2[1], q2[2], q2[3]);
                    }
                }
            }
            if (row < a.L && a.y) {
                float *dst = a.y + (size_t)row * a.ldy + n0;
                if (n0 + ncols <= a.N && (a.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (j < ncols) *reinterpret_cast<uint4 *>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols && n0 + j < a.N) dst[j] = __uint_as_float(v[j]);
                }
            }
        } else {
            // max over `pool` consecutive rows (= lanes).  Values are compared as unsigned words: after ReLU they are non-negative floats
            // (whose bit patterns order like the values; rows past the end contribute 0), otherwise through the order-preserving
            // float -> uint map.  Reduction by recursive halving inside aligned groups of `pool` lanes: at step d a lane keeps one half
            // of its live columns and hands the other half to lane ^ d, so 32 columns cost 16 + 8 + ... shuffles (31 for pool = 32)
            // instead of 32 per step, and each lane ends up owning 32 / pool distinct columns of its group's result.
            const bool mapped = !a.relu;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                uint32_t u = v[j];
                if (mapped) {
                    u = row < a.L ? u : 0xff800000u;  // -inf for rows past the end
                    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                } else {
                    u = row < a.L ? u : 0u;
                }
                v[j] = u;
            }
            int col_base = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int d = 1 << k, half = 16 >> k;  // live columns before this step: 2 * half
                if (d < pool) {  // warp-uniform
                    const bool upper = (lane & d) != 0;
#pragma unroll
                    for (int i = 0; i < half; ++i) {
                        const uint32_t mine = upper ? v[i + half] : v[i];
                        const uint32_t give = upper ? v[i] : v[i + half];
                        v[i] = max(mine, __shfl_xor_sync(0xffffffffu, give, d));
                    }
                    col_base += upper ? half : 0;
                }
            }
            // live: v[0 .. 32/pool), columns col_base + i of this lane's group
            const int live = 32 / pool;
            if (!mapped) amax = fmaxf(amax, __uint_as_float(v[0]));  // ReLU'd, non-negative: the pooled maxima bound every value pooled away
            if (row < a.L) {
                float *dst = a.y + (size_t)(row / pool) * a.ldy + n0 + col_base;
                if (!mapped && live == 1 && ncols == 32 && n0 + 32 <= a.N) {
                    dst[0] = __uint_as_float(v[0]);  // the common case (nsample = 32, full chunk): one coalesced store per warp
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (i < live && col_base + i < ncols && n0 + col_base + i < a.N) {
                            const uint32_t u = v[i];
                            dst[i] = __uint_as_float(mapped ? ((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u) : u);
                            if (!mapped && i > 0) amax = fmaxf(amax, __uint_as_float(u));
                        }
                    }
                }
            }
        }
    }
    if (!(amax <= kF16Guard)) atomicOr(&g_gemm_overflow, 1u);  // also true for NaN/inf
}

__global__ void __launch_bounds__(kGmThreads, 2)
gemm_tf32x3_kernel(const GemmArgs a)
{
    extern __shared__ __align__(1024) uint8_t gm_smem[];
    __shared__ __align__(8) uint64_t full_a[kGmMaxStages], full_b[kGmMaxStages], empty[kGmMaxStages], accum_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[256 + 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kGmBM;
    const int ntile = blockIdx.y;
    const int BN = a.BN;
    const uint32_t a_bytes = kGmBM * 128;        // one plane (hi or lo) of the X tile
    const uint32_t b_bytes = (uint32_t)BN * 128; // one plane of the W tile
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    // stage layout: [A_hi | A_lo | B_hi | B_lo], every plane 1024-byte aligned (BN % 8 == 0)
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(gm_smem) + 1023) & ~uintptr_t(1023));

    // two fp32 accumulators: columns [0,BN) take hi*hi, columns [corr_off, corr_off+BN) take the two correction terms.  The tensor
    // core's accumulate truncates, so its error grows with the number of read-modify-writes; keeping the small terms apart
    // leaves the main accumulator with K/8 updates instead of 3K/8 and the sum is formed once, in fp32 RN, in the epilogue.
    const uint32_t BNP = (uint32_t)((BN + 31) & ~31);
    const uint32_t corr_off = BN <= 128 ? (uint32_t)BN : BNP;  // see the MMA issuer
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2 * BNP) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_a[s], kGmProducers);
            mbar_init(&full_b[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 4) {  // TMEM allocation is a warp-wide instruction; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = tid; j < 256 + 32; j += kGmThreads) {
        const int n = ntile * BN + j;
        bias_s[j] = (a.bias && j < BN && n < a.N) ? __ldg(a.bias + n) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp < 4) {
        // ===================== X producers =====================
        const int chunk = tid & 7;          // 16-byte chunk of the 128-byte k-row
        const int rbase = tid >> 3;         // 0..15; rows rbase + 16*i
        // global loads of k-block kb+1 are issued before k-block kb is converted and stored: one block of latency is hidden
        RowSource src;
        src.init(a, row0 + rbase, 16);
        auto load_block = [&](int kb, float4 (&v)[8]) { src.load(a, kb * kGmBK + chunk * 4, v); };
        uint32_t slot[8];  // byte offset of this thread's chunk of row rbase + 16*i inside a plane (swizzled)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rbase + 16 * i;
            slot[i] = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
        }
        float4 cur[8], nxt[8];
        load_block(0, cur);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < a.n_kblocks; ++kb) {
            if (kb + 1 < a.n_kblocks) load_block(kb + 1, nxt);
            mbar_wait(&empty[s], ph ^ 1u);
            uint8_t *a_hi = smem + (size_t)s * stage_bytes;
            uint8_t *a_lo = a_hi + a_bytes;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = cur[i];
                uint4 hi, lo;
                hi.x = __float_as_uint(v.x) & 0xffffe000u; hi.y = __float_as_uint(v.y) & 0xffffe000u;
                hi.z = __float_as_uint(v.z) & 0xffffe000u; hi.w = __float_as_uint(v.w) & 0xffffe000u;
                lo.x = __float_as_uint(__fsub_rn(v.x, __uint_as_float(hi.x))); lo.y = __float_as_uint(__fsub_rn(v.y, __uint_as_float(hi.y)));
                lo.z = __float_as_uint(__fsub_rn(v.z, __uint_as_float(hi.z))); lo.w = __float_as_uint(__fsub_rn(v.w, __uint_as_float(hi.w)));
                *reinterpret_cast<uint4 *>(a_hi + slot[i]) = hi;
                *reinterpret_cast<uint4 *>(a_lo + slot[i]) = lo;
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive(&full_a[s]);
#pragma unroll
            for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
            if (++s == a.stages) { s = 0; ph ^= 1u; }
        }

        // ===================== epilogue =====================
        mbar_wait(&accum_bar, 0u);
        tc_fence_after();
        gemm_epilogue(a, tmem_acc, corr_off, warp, lane, row0, ntile, bias_s);
        tc_fence_before();
    } else if (warp == 4) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // BN <= 128: the two weight planes are adjacent in the stage ([B_hi | B_lo], 8-row atoms), so ONE MMA with N = 2*BN forms
            // hi*hi into columns [0,BN) and hi*lo into [BN,2BN) while reading A_hi once; lo*hi then accumulates onto [BN,2BN).
            // Shared-memory operand reads per k-step drop from 3 to 2 A-planes (the kernel is shared-memory-bandwidth bound
            // for small N).  BN > 128 keeps three MMAs (N = 2*BN would exceed 256).
            const bool fused_b = BN <= 128;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&full_a[s], ph);
                mbar_wait(&full_b[s], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t d_ah = umma_desc_k_sw128(base), d_al = umma_desc_k_sw128(base + a_bytes);
                const uint64_t d_bh = umma_desc_k_sw128(base + 2 * a_bytes), d_bl = umma_desc_k_sw128(base + 2 * a_bytes + b_bytes);
#pragma unroll
                for (int ks = 0; ks < kGmBK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);  // 8 tf32 = 32 bytes = 2 x 16 B along the swizzled row
                    if (fused_b) {
                        umma_tf32(tmem_acc, d_ah + adv, d_bh + adv, idesc2, (kb | ks) ? 1u : 0u);
                        umma_tf32(tmem_acc + corr_off, d_al + adv, d_bh + adv, idesc, 1u);
                    } else {
                        umma_tf32(tmem_acc + corr_off, d_al + adv, d_bh + adv, idesc, (kb | ks) ? 1u : 0u);
                        umma_tf32(tmem_acc + corr_off, d_ah + adv, d_bl + adv, idesc, 1u);
                        umma_tf32(tmem_acc, d_ah + adv, d_bh + adv, idesc, (kb | ks) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);  // implies tcgen05.fence::before_thread_sync
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(&accum_bar);
        }
        __syncwarp();
    } else {
        // ===================== weight loader (one thread) =====================
        if (lane == 0) {
            const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(a.wpack) + (size_t)ntile * a.n_kblocks * 2 * b_bytes;
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_b[s], 2 * b_bytes);
                bulk_g2s(smem + (size_t)s * stage_bytes + 2 * a_bytes, wsrc + (size_t)kb * 2 * b_bytes, 2 * b_bytes, &full_b[s]);
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Wide tiles (BN > 64) with a two-term FP16 split instead of the TF32 split:  x = h1 + 2^-11 h2,  h1 = fp16(x),
// h2 = fp16((x - h1) * 2^11)  (22 significand bits, like hi/lo in TF32), and likewise for W.  Three kind::f16 MMAs per 16
// k-values (h1 g1 -> main accumulator; h1 g2 + h2 g1 -> correction accumulator, scaled back by 2^-11 in the epilogue) run at
// twice the TF32 rate and every operand byte in shared memory carries twice as many k-values: the wide-tile kernel is bound by
// exactly these two resources.  Inputs must stay inside fp16's range (|x| < 65504; the backbone's activations and folded
// weights are below 10, tools/fp16_split_probe.py); values below 2^-24 lose relative but not absolute accuracy.
// 10 warps, one CTA per SM (a 256-column tile owns all 512 TMEM columns anyway): warps 0-7 split X (a k-block is 64 k-values:
// 256 bytes of fp32 per row in, 128 bytes per plane out) and run the epilogue, warp 8 issues the MMAs, warp 9 loads W.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kHfThreads = 320;
constexpr int kHfProducers = 256;
constexpr int kHfBK = 64;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__global__ void __launch_bounds__(kHfThreads, 1)
gemm_f16x3_kernel(const GemmArgs a)
{
    extern __shared__ __align__(1024) uint8_t gm_smem[];
    __shared__ __align__(8) uint64_t full_a[kGmMaxStages], full_b[kGmMaxStages], empty[kGmMaxStages], accum_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[256 + 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kGmBM;
    const int ntile = blockIdx.y;
    const int BN = a.BN;
    const uint32_t a_bytes = kGmBM * 128;        // one plane (h1 or h2) of the X tile: 128 rows x 64 halfs
    const uint32_t b_bytes = (uint32_t)BN * 128; // one plane of the W tile
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(gm_smem) + 1023) & ~uintptr_t(1023));
    const uint32_t BNP = (uint32_t)((BN + 31) & ~31);
    const uint32_t corr_off = BN <= 128 ? (uint32_t)BN : BNP;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2 * BNP) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_a[s], kHfProducers);
            mbar_init(&full_b[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = tid; j < 256 + 32; j += kHfThreads) {
        const int n = ntile * BN + j;
        bias_s[j] = (a.bias && j < BN && n < a.N) ? __ldg(a.bias + n) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;

    if (warp < 8) {
        // ===================== X producers =====================
        const int chunk = tid & 7;   // 16-byte chunk of a plane row = 8 consecutive k-values = 32 bytes of fp32 input
        const int rbase = tid >> 3;  // 0..31; rows rbase + 32*i, i < 4
        RowSource src;
        src.init(a, row0 + rbase, 32);
#pragma unroll
        for (int i = 4; i < 8; ++i) src.rmask[i] = 0u;  // those rows belong to other threads
        uint32_t slot[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rbase + 32 * i;
            slot[i] = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
        }
        float4 cur[8], nxt[8];  // [2*i], [2*i+1]: the 8 floats of row i
        src.load4x8(a, chunk * 8, cur);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < a.n_kblocks; ++kb) {
            if (kb + 1 < a.n_kblocks) src.load4x8(a, (kb + 1) * kHfBK + chunk * 8, nxt);
            mbar_wait(&empty[s], ph ^ 1u);
            uint8_t *p1 = smem + (size_t)s * stage_bytes;
            uint8_t *p2 = p1 + a_bytes;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x[8] = {cur[2 * i].x, cur[2 * i].y, cur[2 * i].z, cur[2 * i].w, cur[2 * i + 1].x, cur[2 * i + 1].y, cur[2 * i + 1].z, cur[2 * i + 1].w};
                __half h1[8], h2[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_f16(x[e], h1[e], h2[e]);
                uint4 q1, q2;
                q1.x = (uint32_t)__half_as_ushort(h1[0]) | ((uint32_t)__half_as_ushort(h1[1]) << 16);
                q1.y = (uint32_t)__half_as_ushort(h1[2]) | ((uint32_t)__half_as_ushort(h1[3]) << 16);
                q1.z = (uint32_t)__half_as_ushort(h1[4]) | ((uint32_t)__half_as_ushort(h1[5]) << 16);
                q1.w = (uint32_t)__half_as_ushort(h1[6]) | ((uint32_t)__half_as_ushort(h1[7]) << 16);
                q2.x = (uint32_t)__half_as_ushort(h2[0]) | ((uint32_t)__half_as_ushort(h2[1]) << 16);
                q2.y = (uint32_t)__half_as_ushort(h2[2]) | ((uint32_t)__half_as_ushort(h2[3]) << 16);
                q2.z = (uint32_t)__half_as_ushort(h2[4]) | ((uint32_t)__half_as_ushort(h2[5]) << 16);
                q2.w = (uint32_t)__half_as_ushort(h2[6]) | ((uint32_t)__half_as_ushort(h2[7]) << 16);
                *reinterpret_cast<uint4 *>(p1 + slot[i]) = q1;
                *reinterpret_cast<uint4 *>(p2 + slot[i]) = q2;
            }
            fence_proxy_async();
            mbar_arrive(&full_a[s]);
#pragma unroll
            for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
            if (++s == a.stages) { s = 0; ph ^= 1u; }
        }

        // ===================== epilogue: warps w and w+4 share a TMEM lane quarter and alternate 32-column chunks =====================
        mbar_wait(&accum_bar, 0u);
        tc_fence_after();
        gemm_epilogue(a, tmem_acc, corr_off, warp & 3, lane, row0, ntile, bias_s, (warp >> 2) * 32, 64);
        tc_fence_before();
    } else if (warp == 8) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const bool fused_b = BN <= 128;
            // kind::f16, A and B fp16 (format 0), D fp32 (bit 4), K-major both
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&full_a[s], ph);
                mbar_wait(&full_b[s], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t d_a1 = umma_desc_k_sw128(base), d_a2 = umma_desc_k_sw128(base + a_bytes);
                const uint64_t d_b1 = umma_desc_k_sw128(base + 2 * a_bytes), d_b2 = umma_desc_k_sw128(base + 2 * a_bytes + b_bytes);
#pragma unroll
                for (int ks = 0; ks < kHfBK / 16; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);  // 16 halfs = 32 bytes = 2 x 16 B along the swizzled row
                    if (fused_b) {
                        umma_f16(tmem_acc, d_a1 + adv, d_b1 + adv, idesc2, (kb | ks) ? 1u : 0u);
                        umma_f16(tmem_acc + corr_off, d_a2 + adv, d_b1 + adv, idesc, 1u);
                    } else {
                        umma_f16(tmem_acc + corr_off, d_a2 + adv, d_b1 + adv, idesc, (kb | ks) ? 1u : 0u);
                        umma_f16(tmem_acc + corr_off, d_a1 + adv, d_b2 + adv, idesc, 1u);
                        umma_f16(tmem_acc, d_a1 + adv, d_b1 + adv, idesc, (kb | ks) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(&accum_bar);
        }
        __syncwarp();
    } else {
        // ===================== weight loader (one thread) =====================
        if (lane == 0) {
            const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(a.wpack) + (size_t)ntile * a.n_kblocks * 2 * b_bytes;
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_b[s], 2 * b_bytes);
                bulk_g2s(smem + (size_t)s * stage_bytes + 2 * a_bytes, wsrc + (size_t)kb * 2 * b_bytes, 2 * b_bytes, &full_b[s]);
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// A-from-TMEM variant for narrow tiles (BN <= 64).  With few output columns the SS kernel above is bound by shared-memory
// bandwidth: per k-block it writes 32 KB of X planes and the tensor core reads them back twice.  Here X never becomes a
// shared-memory operand: a producer warp copies its 32 rows into a warp-private ring with cp.async (coalesced: 8 lanes per
// 128-byte row segment, several k-blocks ahead), reads them back turned so that lane l holds row l, and writes hi (the raw fp32 words: kind::tf32 ignores the 13 low mantissa
// bits, so they ARE x & 0xffffe000) and lo = x - hi into TMEM columns with tcgen05.st; the MMAs take A from TMEM
// (tcgen05.mma [d], [a], b-desc).  Shared memory then carries only the weights (ring of k-blocks).
//
// TMEM columns (256 per CTA, two CTAs per SM): [0, 2*BN) accumulators (main | correction), then a_stages x 64 columns of X
// (hi 32 | lo 32), a_stages = 2 or 3.
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, single CTA.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

constexpr int kTsMaxBStages = 3;
constexpr int kTsMaxAStages = 3;

__global__ void __launch_bounds__(kGmThreads, 2)
gemm_tf32x3_ts_kernel(const GemmArgs a)
{
    extern __shared__ __align__(1024) uint8_t gm_smem[];
    __shared__ __align__(8) uint64_t full_a[kTsMaxAStages], empty_a[kTsMaxAStages], full_b[kTsMaxBStages], empty_b[kTsMaxBStages], accum_bar,
        acc_empty;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[256 + 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int m_tiles = (a.L + kGmBM - 1) / kGmBM;  // this CTA walks m-tiles blockIdx.x, blockIdx.x + gridDim.x, ... of one n-tile
    if (a.m_tiles_dev) m_tiles = min(m_tiles, __ldg(a.m_tiles_dev));
    const int ntile = blockIdx.y;
    const int BN = a.BN;                           // <= 64
    const uint32_t b_bytes = (uint32_t)BN * 128;   // one plane of the W tile
    // shared memory: [4 warps x depth x 4 KB of fp32 rows | ring of a.stages x (B_hi | B_lo)]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(gm_smem) + 1023) & ~uintptr_t(1023));
    uint8_t *b_ring = smem + (size_t)4 * a.depth * 4096;

    const uint32_t acc_cols = (uint32_t)((2 * BN + 31) & ~31);
    const int a_stages = (256 - (int)acc_cols) / 64 > kTsMaxAStages ? kTsMaxAStages : (256 - (int)acc_cols) / 64;  // 2 (BN = 64) or 3
    const uint32_t tmem_cols = 256;
    const uint32_t corr_off = (uint32_t)BN;

    if (tid == 0) {
        for (int s = 0; s < a_stages; ++s) {
            mbar_init(&full_a[s], kGmProducers);
            mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full_b[s], 1);
            mbar_init(&empty_b[s], 1);
        }
        mbar_init(&accum_bar, 1);
        mbar_init(&acc_empty, kGmProducers);
        mbar_fence_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = tid; j < 256 + 32; j += kGmThreads) {
        const int n = ntile * BN + j;
        bias_s[j] = (a.bias && j < BN && n < a.N) ? __ldg(a.bias + n) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base_slot;
    const uint32_t tmem_a0 = tmem_acc + acc_cols;  // first X stage

    if (warp < 4) {
        // ===================== X producers =====================
        const int chunk = lane & 7;   // 16-byte chunk of the 128-byte k-row
        const int rsub = lane >> 3;   // loads: rows 32*warp + 4*i + rsub (i < 8), i.e. the warp's own TMEM lane quarter
        RowSource src;
        // warp-private ring of `depth` slots of 32 rows x 128 B (chunk c of row r at c ^ (r & 7)): k-blocks are copied in with
        // cp.async `depth` ahead (no register staging, no cross-warp synchronisation), then each lane reads back its own row
        uint8_t *ring = smem + (size_t)warp * a.depth * 4096;
        uint32_t st_off[8], ld_off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            st_off[i] = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
            ld_off[i] = (uint32_t)lane * 128u + (uint32_t)((i ^ (lane & 7)) << 4);  // chunk i of this lane's row
        }
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        int t = 0;
        uint32_t ph = 0;
        // the first `depth` k-blocks of a tile are requested before the previous tile's epilogue, so their latency hides behind it
        auto prefetch_tile = [&](int mt) {
            src.init(a, mt * kGmBM + 32 * warp + rsub, 4);
            for (int j = 0; j < a.depth; ++j) {  // always `depth` groups, empty ones past the end: group index == k-block index
                if (j < a.n_kblocks) src.copy_async(a, j * kGmBK + chunk * 4, ring + j * 4096, st_off);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        };
        if ((int)blockIdx.x < m_tiles) prefetch_tile(blockIdx.x);
        for (int mt = blockIdx.x, it = 0; mt < m_tiles; mt += gridDim.x, ++it) {
        const int row0 = mt * kGmBM;
        int slot = 0;
        // grouped operand: this lane's row gets three computed columns, the sample's offset from its centre
        float goff[3] = {0.f, 0.f, 0.f};
        if (a.g_idx && a.g_xyz) {
            const int row = row0 + 32 * warp + lane;
            if (row < a.L) {
                const float *pp = a.g_xyz + ((size_t)(row / a.g_rows_scene) * a.g_n + __ldg(a.g_idx + row)) * 3;
                const float *cc = a.g_centre + (size_t)(row / a.g_ns) * 3;
#pragma unroll
                for (int j = 0; j < 3; ++j) goff[j] = __fsub_rn(__ldg(pp + j), __ldg(cc + j));  // as pointnet2_utils.py:252
            }
        }
        for (int kb = 0; kb < a.n_kblocks; ++kb) {
            // k-block kb has landed when at most depth-1 newer groups are pending
            if (a.depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else if (a.depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else if (a.depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else asm volatile("cp.async.wait_group 3;" ::: "memory");
            __syncwarp();  // every lane's copies of this k-block are complete and visible to the warp
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 q = *reinterpret_cast<const uint4 *>(ring + slot * 4096 + ld_off[i]);
                v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
            }
            __syncwarp();  // the slot has been read by all lanes: refill it
            if (a.g_idx && a.g_xyz && (kb + 1) * kGmBK > a.kcopy && kb * kGmBK < a.kcopy + 3) {  // this k-block holds offset columns
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int col = a.kcopy + j - kb * kGmBK;
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (e == col) v[e] = __float_as_uint(goff[j]);
                }
            }
            if (kb + a.depth < a.n_kblocks) src.copy_async(a, (kb + a.depth) * kGmBK + chunk * 4, ring + slot * 4096, st_off);
            asm volatile("cp.async.commit_group;" ::: "memory");
            mbar_wait(&empty_a[t], ph ^ 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_a0 + (uint32_t)t * 64u + lane_addr;
            tmem_st32(taddr, v);  // hi: the raw words
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float x = __uint_as_float(v[j]);
                v[j] = __float_as_uint(__fsub_rn(x, __uint_as_float(v[j] & 0xffffe000u)));
            }
            tmem_st32(taddr + 32u, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&full_a[t]);
            if (++t == a_stages) { t = 0; ph ^= 1u; }
            if (++slot == a.depth) slot = 0;
        }

        if (mt + (int)gridDim.x < m_tiles) prefetch_tile(mt + gridDim.x);  // every ring slot of this tile has been read

        // ===================== epilogue =====================
        mbar_wait(&accum_bar, (uint32_t)it & 1u);
        tc_fence_after();
        gemm_epilogue(a, tmem_acc, corr_off, warp, lane, row0, ntile, bias_s);
        tc_fence_before();
        mbar_arrive(&acc_empty);  // the accumulator may be overwritten by the next tile
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            int t = 0, s = 0;
            uint32_t pha = 0, phb = 0;
            for (int mt = blockIdx.x, it = 0; mt < m_tiles; mt += gridDim.x, ++it) {
            mbar_wait(&acc_empty, ((uint32_t)it & 1u) ^ 1u);  // the epilogue of the previous tile has read the accumulator
            tc_fence_after();
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                mbar_wait(&full_a[t], pha);
                mbar_wait(&full_b[s], phb);
                tc_fence_after();
                const uint64_t d_bh = umma_desc_k_sw128(smem_u32(b_ring + (size_t)s * 2 * b_bytes));  // [B_hi | B_lo]: 2*BN rows
                const uint32_t a_hi = tmem_a0 + (uint32_t)t * 64u, a_lo = a_hi + 32u;
#pragma unroll
                for (int ks = 0; ks < kGmBK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);  // 8 tf32 = 32 bytes = 2 x 16 B along the swizzled row
                    umma_tf32_ts(tmem_acc, a_hi + ks * 8, d_bh + adv, idesc2, (kb | ks) ? 1u : 0u);  // hi*hi | hi*lo
                    umma_tf32_ts(tmem_acc + corr_off, a_lo + ks * 8, d_bh + adv, idesc, 1u);           // lo*hi
                }
                umma_commit(&empty_a[t]);
                umma_commit(&empty_b[s]);
                if (++t == a_stages) { t = 0; pha ^= 1u; }
                if (++s == a.stages) { s = 0; phb ^= 1u; }
            }
            umma_commit(&accum_bar);
            }
        }
        __syncwarp();
    } else {
        // ===================== weight loader (one thread) =====================
        if (lane == 0) {
            const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(a.wpack) + (size_t)ntile * a.n_kblocks * 2 * b_bytes;
            int s = 0;
            uint32_t ph = 0;
            for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x)
            for (int kb = 0; kb < a.n_kblocks; ++kb) {
                if (a.tile_phase && kb == 0) {  // this m-tile's weight set (sparse image tail)
                    const int p16 = __ldg(a.tile_phase + mt);
                    const int wt = ((p16 >> 4) % a.phase_k) * a.phase_k + ((p16 & 15) % a.phase_k);
                    wsrc = reinterpret_cast<const uint8_t *>(a.wpack) + (size_t)wt * a.n_kblocks * 2 * b_bytes;
                }
                mbar_wait(&empty_b[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_b[s], 2 * b_bytes);
                bulk_g2s(b_ring + (size_t)s * 2 * b_bytes, wsrc + (size_t)kb * 2 * b_bytes, 2 * b_bytes, &full_b[s]);
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// FP16-split GEMM / 3x3 convolution whose A operand arrives ALREADY SPLIT (two FP16 planes written by the producing layer's
// epilogue, GemmArgs::yh1/yh2) and is moved by TMA TENSOR loads (cp.async.bulk.tensor, SASS UTMALDG) straight into the swizzled
// shared-memory image the MMA reads.  Compared with gemm_f16x3_kernel there are no producer warps at all: nothing is converted,
// nothing is addressed per thread -- one elected thread issues two tensor copies (planes h1, h2) and one bulk copy (weights) per
// k-block; image borders (padding 1) and ragged tile edges are the TMA unit's out-of-bounds zero fill; stride-2 convolutions use
// the tensor map's traversal stride.  An implicit-GEMM convolution re-reads every input pixel for 9 taps and every column tile;
// with SIMT producers it also re-CONVERTED it each time (measured: the N <= 128 convolutions were bound by that work).
//
// Convolution mode: x planes are NHWC FP16 images (B, H, W, ldx >= Cin), Cin % 64 == 0; a k-block is one tap x 64 channels:
// box {64 ch, tw px, th px, 1 image} at coordinates {c0, ox0*stride + kx - 1, oy0*stride + ky - 1, b} -> 128 rows of 128 bytes,
// row r = (y_local * tw + x_local), SWIZZLE_128B -- exactly the K-major UMMA layout.  Plain mode: planes are (L, ldx) rows, box
// {64, 128}.  6 warps: 0-3 epilogue (TMEM lane quarters), 4 MMA issuer, 5 TMA producer.  One CTA per SM.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kTmaThreads = 192;

__device__ __forceinline__ void tma_load_4d(void *dst_smem, const void *tmap, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const void *tmap, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

struct TmaMaps {
    alignas(64) unsigned char h1[128];  // CUtensorMap of plane h1
    alignas(64) unsigned char h2[128];  // CUtensorMap of plane h2
};

__global__ void __launch_bounds__(kTmaThreads, 1)
gemm_f16x3_tma_kernel(const GemmArgs a, const __grid_constant__ TmaMaps maps, int m_tiles, int n_tiles)
{
    // PERSISTENT: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the (m-tile, n-tile) grid, n fastest
    // (neighbouring CTAs read the same activation tile: it is fetched from HBM once and served from L2).  TMEM allocation, barrier
    // set-up and the tensor-map fetch are paid once per CTA; the shared-memory ring never drains between tiles (the producer runs
    // ahead into the next tile while the epilogue drains the accumulator); with BN <= 128 the accumulator is double-buffered
    // (2 x (main | correction) = 512 TMEM columns) so the epilogue of tile i overlaps the MMAs of tile i + 1.
    extern __shared__ __align__(1024) uint8_t gm_smem[];
    __shared__ __align__(8) uint64_t full[kGmMaxStages], empty[kGmMaxStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float bias_s[2][256 + 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = a.BN;
    const uint32_t a_bytes = kGmBM * 128;
    const uint32_t b_bytes = (uint32_t)BN * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(gm_smem) + 1023) & ~uintptr_t(1023));
    const uint32_t BNP = (uint32_t)((BN + 31) & ~31);
    const uint32_t corr_off = BN <= 128 ? (uint32_t)BN : BNP;
    const int nbuf = 4 * BNP <= 512 ? 2 : 1;        // accumulator buffers
    const uint32_t buf_cols = 2 * BNP;              // main | correction
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)nbuf * buf_cols) tmem_cols <<= 1;
    const int total = m_tiles * n_tiles;

    if (tid == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 5 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps.h1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(maps.h2)) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc0 = tmem_base_slot;

    if (warp < 4) {
        // ===================== epilogue (128 threads; named barrier 1 among them) =====================
        for (int t = blockIdx.x, it = 0; t < total; t += gridDim.x, ++it) {
            const int mt = t / n_tiles, ntile = t - mt * n_tiles;
            const int buf = nbuf == 2 ? (it & 1) : 0;
            const uint32_t par = nbuf == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
            float *bs = bias_s[it & 1];
            for (int j = tid; j < 256 + 32; j += 128) {
                const int n = ntile * BN + j;
                bs[j] = (a.bias && j < BN && n < a.N) ? __ldg(a.bias + n) : 0.f;
            }
            int row = mt * kGmBM + warp * 32 + lane;
            if (a.tw) {
                const int per_img = a.tiles_x * a.tiles_y;
                const int img = mt / per_img, tt = mt - img * per_img;
                const int r = warp * 32 + lane;
                const int oy = (tt / a.tiles_x) * a.th + r / a.tw, ox = (tt - (tt / a.tiles_x) * a.tiles_x) * a.tw + (r - (r / a.tw) * a.tw);
                row = (oy < a.Ho && ox < a.Wo) ? (img * a.Ho + oy) * a.Wo + ox : a.L;  // a.L: outside the image, nothing is stored
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");  // this tile's bias is in shared memory
            mbar_wait(&acc_full[buf], par);
            tc_fence_after();
            gemm_epilogue(a, tmem_acc0 + (uint32_t)buf * buf_cols, corr_off, warp, lane, mt * kGmBM, ntile, bs, 0, 32, row);
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (one thread): same MMA sequence as gemm_f16x3_kernel =====================
        if (lane == 0) {
            const bool fused_b = BN <= 128;
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | ((uint32_t)(BN >> 2) << 17) | ((uint32_t)(kGmBM >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x, it = 0; t < total; t += gridDim.x, ++it) {
                const int buf = nbuf == 2 ? (it & 1) : 0;
                const uint32_t par = nbuf == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
                mbar_wait(&acc_empty[buf], par ^ 1u);  // the epilogue has drained this buffer (passes at once the first time round)
                tc_fence_after();
                const uint32_t tmem_acc = tmem_acc0 + (uint32_t)buf * buf_cols;
                for (int kb = 0; kb < a.n_kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t d_a1 = umma_desc_k_sw128(base), d_a2 = umma_desc_k_sw128(base + a_bytes);
                    const uint64_t d_b1 = umma_desc_k_sw128(base + 2 * a_bytes), d_b2 = umma_desc_k_sw128(base + 2 * a_bytes + b_bytes);
#pragma unroll
                    for (int ks = 0; ks < kHfBK / 16; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);
                        if (fused_b) {
                            umma_f16(tmem_acc, d_a1 + adv, d_b1 + adv, idesc2, (kb | ks) ? 1u : 0u);
                            umma_f16(tmem_acc + corr_off, d_a2 + adv, d_b1 + adv, idesc, 1u);
                        } else {
                            umma_f16(tmem_acc + corr_off, d_a2 + adv, d_b1 + adv, idesc, (kb | ks) ? 1u : 0u);
                            umma_f16(tmem_acc + corr_off, d_a1 + adv, d_b2 + adv, idesc, 1u);
                            umma_f16(tmem_acc, d_a1 + adv, d_b1 + adv, idesc, (kb | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[s]);
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ===================== TMA producer (one thread): planes by tensor copies, weights by a bulk copy =====================
        if (lane == 0) {
            const int kpc = a.conv ? (1 << a.cin_shift) / kHfBK : 1;  // k-blocks per tap
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int mt = t / n_tiles, ntile = t - mt * n_tiles;
                const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(a.wpack) + (size_t)ntile * a.n_kblocks * 2 * b_bytes;
                int img = 0, oy0 = 0, ox0 = 0;
                if (a.tw) {
                    const int per_img = a.tiles_x * a.tiles_y;
                    img = mt / per_img;
                    const int tt = mt - img * per_img;
                    oy0 = (tt / a.tiles_x) * a.th;
                    ox0 = (tt - (tt / a.tiles_x) * a.tiles_x) * a.tw;
                }
                const int row0 = mt * kGmBM;
                for (int kb = 0; kb < a.n_kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full[s], 2 * a_bytes + 2 * b_bytes);
                    uint8_t *st = smem + (size_t)s * stage_bytes;
                    if (a.conv) {
                        const int tap = kb / kpc, c0 = (kb - tap * kpc) * kHfBK;
                        const int ky = tap / 3, kx = tap - 3 * ky;
                        const int cx = ox0 * a.stride + kx - 1, cy = oy0 * a.stride + ky - 1;
                        tma_load_4d(st, maps.h1, &full[s], c0, cx, cy, img);
                        tma_load_4d(st + a_bytes, maps.h2, &full[s], c0, cx, cy, img);
                    } else {
                        tma_load_2d(st, maps.h1, &full[s], kb * kHfBK, row0);
                        tma_load_2d(st + a_bytes, maps.h2, &full[s], kb * kHfBK, row0);
                    }
                    bulk_g2s(st + 2 * a_bytes, wsrc + (size_t)kb * 2 * b_bytes, 2 * b_bytes, &full[s]);
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc0), "r"(tmem_cols) : "memory");
    }
}

}  // namespace epnet

namespace epnet {
static int gemm_launch(GemmArgs &a, cudaStream_t st)
{
    a.n_kblocks = (a.K + kGmBK - 1) / kGmBK;
    if (!a.g_idx || !a.g_xyz) a.kcopy = a.K;
    if (a.corr_scale == 0.f) a.corr_scale = 1.f;
    const int n_tiles = (a.N + a.BN - 1) / a.BN;
    dim3 grid((a.L + kGmBM - 1) / kGmBM, n_tiles);
    // the opt-in shared-memory limit is a property of the function, not of a launch: always raise it to the hardware maximum so
    // that a kernel node captured in a CUDA graph with a large request stays launchable after later, smaller launches
    auto raise_limit = [](const void *kernel) -> int {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return (int)e;
        return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
    };
    if (a.f16) {  // FP16 split, k-blocks of 64, one CTA per SM
        if (a.g_idx || a.BN <= 64) return EPNET_ERR_BAD_ARG;  // narrow tiles stay on the TF32 A-from-TMEM kernel
        a.n_kblocks = (a.K + kHfBK - 1) / kHfBK;
        a.corr_scale = 1.0f / 2048.0f;
        const size_t sb = 2 * (size_t)kGmBM * 128 + 2 * (size_t)a.BN * 128;
        int st_ = (int)((200 * 1024) / sb);
        if (st_ > kGmMaxStages) st_ = kGmMaxStages;
        if (st_ > a.n_kblocks) st_ = a.n_kblocks;
        if (st_ < 2 && a.n_kblocks > 1) st_ = 2;
        if (st_ < 1) st_ = 1;
        a.stages = st_;
        const int e16 = raise_limit((const void *)gemm_f16x3_kernel);
        if (e16) return e16;
        gemm_f16x3_kernel<<<grid, kHfThreads, sb * st_ + 1024, st>>>(a);
        EPNET_RETURN_LAUNCH_STATUS();
    }
    if (a.BN <= 64) {  // narrow tile: X through TMEM, shared memory holds only a ring of weight k-blocks
        int stages = a.n_kblocks < kTsMaxBStages ? a.n_kblocks : kTsMaxBStages;
        a.stages = stages < 1 ? 1 : stages;
        a.depth = a.n_kblocks < 3 ? a.n_kblocks : 3;  // 48 KB of X + <= 48 KB of W per CTA: two CTAs per SM
        const size_t smem = (size_t)4 * a.depth * 4096 + (size_t)a.stages * 2 * a.BN * 128 + 1024;
        const int e = raise_limit((const void *)gemm_tf32x3_ts_kernel);
        if (e) return e;
        // persistent over m-tiles: two CTAs per SM in total, each walking its share of the tiles of one n-tile, so TMEM
        // allocation, barrier set-up and CTA launch/exit are paid once per CTA instead of once per 128 rows
        const int per_ntile = (2 * kSmCount + n_tiles - 1) / n_tiles;
        if ((int)grid.x > per_ntile) grid.x = per_ntile;
        gemm_tf32x3_ts_kernel<<<grid, kGmThreads, smem, st>>>(a);
        EPNET_RETURN_LAUNCH_STATUS();
    }
    const size_t stage_bytes = 2 * (size_t)kGmBM * 128 + 2 * (size_t)a.BN * 128;
    int stages = (int)((100 * 1024) / stage_bytes);  // <= ~100 KB per CTA: two CTAs per SM overlap each other's epilogue
    if (stages > kGmMaxStages) stages = kGmMaxStages;
    if (stages > a.n_kblocks) stages = a.n_kblocks;
    if (stages < 2 && a.n_kblocks > 1) stages = 2;
    if (stages < 1) stages = 1;
    a.stages = stages;
    const size_t smem = stage_bytes * stages + 1024;
    const int e = raise_limit((const void *)gemm_tf32x3_kernel);
    if (e) return e;
    gemm_tf32x3_kernel<<<grid, kGmThreads, smem, st>>>(a);
    EPNET_RETURN_LAUNCH_STATUS();
}
}  // namespace epnet

// x (L, ldx) fp32 rows, wpack from pack_weights (N_tiles x n_kblocks x 2 x BN x 32), bias (N) or NULL -> y (L/pool, ldy).
static int gemm_entry(int f16, int L, int K, int N, const float *x, int ldx, const float *wpack, int BN, const float *bias, int relu,
                      int pool, float *y, int ldy, void *stream)
{
    using namespace epnet;
    if (L < 0 || K <= 0 || N <= 0 || !x || !wpack || !y || ldx < K || ldy < N) return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 256 || (BN % 16) != 0) return EPNET_ERR_BAD_ARG;
    if (pool < 1 || pool > 32 || (32 % pool) != 0 || (L % pool) != 0) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(wpack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (L == 0) return EPNET_OK;
    GemmArgs a = {};
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = y;
    a.L = L; a.K = K; a.N = N; a.ldx = ldx; a.ldy = ldy; a.BN = BN;
    a.relu = relu; a.pool = pool; a.f16 = f16;
    a.x_vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0);
    return gemm_launch(a, (cudaStream_t)stream);
}
EPNET_API int epnet_gemm_tf32x3(int L, int K, int N, const float *x, int ldx, const float *wpack, int BN, const float *bias, int relu,
                                int pool, float *y, int ldy, void *stream)
{
    return gemm_entry(0, L, K, N, x, ldx, wpack, BN, bias, relu, pool, y, ldy, stream);
}
// Same contract with the FP16 two-term split (wpack = FP16 planes, k-blocks of 64; BN > 64; |x|, |w| < 65504)
EPNET_API int epnet_gemm_f16x3(int L, int K, int N, const float *x, int ldx, const float *wpack, int BN, const float *bias, int relu,
                               int pool, float *y, int ldy, void *stream)
{
    return gemm_entry(1, L, K, N, x, ldx, wpack, BN, bias, relu, pool, y, ldy, stream);
}

// First shared-MLP layer of a set-abstraction scale with the grouping fused into the operand load (QueryAndGroup + Conv2d 1x1 +
// BN + ReLU of pointnet2_modules.py:47-52): row (scene, centre p, sample s) = [feats[scene, idx[scene,p,s], :c] | xyz[scene, idx] -
// new_xyz[scene, p]], never written to memory.  feats point-major (scenes, n, ldf) or NULL when c == 0; BN <= 64.
EPNET_API int epnet_gemm_tf32x3_grouped(int scenes, int n, int m, int nsample, int c, const float *feats, int ldf, const float *xyz,
                                        const float *new_xyz, const int *idx, const float *wpack, int BN, int N, const float *bias,
                                        int relu, int pool, float *y, int ldy, void *stream)
{
    using namespace epnet;
    if (scenes < 0 || n <= 0 || m < 0 || nsample <= 0 || c < 0 || (c > 0 && (!feats || ldf < c)) || !xyz || !new_xyz || !idx || !wpack || !y)
        return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 64 || (BN % 16) != 0 || N <= 0 || ldy < N || (reinterpret_cast<uintptr_t>(wpack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    const long long rows = (long long)scenes * m * nsample;
    if (rows >= (1ll << 31) || pool < 1 || pool > 32 || (32 % pool) != 0 || (rows % pool) != 0) return EPNET_ERR_BAD_ARG;
    if (rows == 0) return EPNET_OK;
    GemmArgs a = {};
    a.x = feats; a.wpack = wpack; a.bias = bias; a.y = y;
    a.L = (int)rows; a.K = c + 3; a.N = N; a.ldx = ldf; a.ldy = ldy; a.BN = BN;
    a.relu = relu; a.pool = pool;
    a.g_idx = idx; a.g_xyz = xyz; a.g_centre = new_xyz; a.g_n = n; a.g_ns = nsample; a.g_rows_scene = m * nsample; a.kcopy = c;
    a.x_vec_ok = c > 0 && ((reinterpret_cast<uintptr_t>(feats) & 15) == 0) && (ldf % 4 == 0);
    return gemm_launch(a, (cudaStream_t)stream);
}

// Narrow-tile GEMM (BN <= 64) over rows GATHERED from a table: row r of the operand is x[row_idx[r]] (K floats, ldx apart); with
// tile_phase the weight set (an n-tile of wpack, i.e. BN output columns) is chosen per 128-row tile -- the transposed-convolution
// phase of the tile's pixels; with m_tiles_dev only the first *m_tiles_dev tiles are computed (device-side row count).  Any of the
// three may be NULL (plain rows / one weight set / all tiles of L).  Used by the sparse evaluation of the image-fusion tail
// (epnet_b200/sparse_tail.py; replaces the dense ConvTranspose2d + concat + 1x1 conv of /root/reference/lib/net/pointnet2_msg.py:237-243 at the
// sampled taps only).  N <= BN: a single column tile.
EPNET_API int epnet_gemm_tf32x3_rows(int L, int K, int N, const float *x, int ldx, const int *row_idx, const int *tile_phase, int phase_k,
                                     const int *m_tiles_dev, const float *wpack, int BN, const float *bias, int relu, float *y, int ldy,
                                     void *stream)
{
    using namespace epnet;
    if (L < 0 || K <= 0 || N <= 0 || !x || !wpack || !y || ldx < K || ldy < N) return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 64 || (BN % 16) != 0 || N > BN || (reinterpret_cast<uintptr_t>(wpack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (tile_phase && (phase_k < 1 || phase_k > 16 || (16 % phase_k) != 0)) return EPNET_ERR_BAD_ARG;
    if (L == 0) return EPNET_OK;
    GemmArgs a = {};
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = y;
    a.L = L; a.K = K; a.N = N; a.ldx = ldx; a.ldy = ldy; a.BN = BN;
    a.relu = relu; a.pool = 1;
    a.x_vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0);
    if (row_idx) {
        a.g_idx = row_idx; a.g_xyz = nullptr; a.g_centre = nullptr; a.g_n = 0; a.g_ns = 1; a.g_rows_scene = 0x7fffffff;
    }
    a.kcopy = K;
    a.tile_phase = tile_phase; a.phase_k = phase_k; a.m_tiles_dev = m_tiles_dev;
    return gemm_launch(a, (cudaStream_t)stream);
}

// The same product with the result written channel-major: x rows are (scene, point) pairs, `pts` points per scene (L % pts == 0),
// y is (L / pts, N, pts) -- the (B, C, N) layout the reference's modules exchange -- so the last layer of a point-major chain
// needs no transposing pass.
EPNET_API int epnet_gemm_tf32x3_cm(int L, int K, int N, int pts, const float *x, int ldx, const float *wpack, int BN, const float *bias,
                                   int relu, float *y, void *stream)
{
    using namespace epnet;
    if (L < 0 || K <= 0 || N <= 0 || pts <= 0 || (L % pts) != 0 || !x || !wpack || !y || ldx < K) return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 256 || (BN % 16) != 0 || (reinterpret_cast<uintptr_t>(wpack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (L == 0) return EPNET_OK;
    GemmArgs a = {};
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = y;
    a.L = L; a.K = K; a.N = N; a.ldx = ldx; a.ldy = N; a.BN = BN;
    a.relu = relu; a.pool = 1; a.tr = pts;
    a.x_vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0);
    return gemm_launch(a, (cudaStream_t)stream);
}

// 3x3 convolution, padding 1, stride 1 or 2, as an implicit GEMM on an NHWC image: x (B, H, W, Cin) with Cin a power of two
// >= 4; wpack packs W reordered to (Cout, ky, kx, Cin) (K = 9*Cin); y (B*Ho*Wo, ldy) = NHWC output, bias/ReLU optional.
static int conv_entry(int f16, int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                      const float *bias, int relu, float *y, int ldy, void *stream, void *yh1 = nullptr, void *yh2 = nullptr, int ldh = 0)
{
    using namespace epnet;
    if (b < 0 || h <= 0 || w <= 0 || cin < 4 || (cin & (cin - 1)) != 0 || cout <= 0 || (stride != 1 && stride != 2) || !x || !wpack || (!y && !yh1))
        return EPNET_ERR_BAD_ARG;
    if (yh1 && (!yh2 || ((reinterpret_cast<uintptr_t>(yh1) | reinterpret_cast<uintptr_t>(yh2)) & 15) != 0 || ldh < cout || (ldh % 8) != 0 || (cout % 8) != 0))
        return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 256 || (BN % 16) != 0 || (y && ldy < cout) || h >= 32768 || w >= 32768) return EPNET_ERR_BAD_ARG;
    if (((reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(x)) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    GemmArgs a = {};
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = y;
    a.conv = 1; a.H = h; a.W = w; a.stride = stride;
    a.Ho = (h + 2 - 3) / stride + 1; a.Wo = (w + 2 - 3) / stride + 1;
    a.cin_shift = 0;
    while ((1 << a.cin_shift) < cin) ++a.cin_shift;
    const long long rows = (long long)b * a.Ho * a.Wo;
    if (rows >= (1ll << 31)) return EPNET_ERR_BAD_ARG;
    a.L = (int)rows; a.K = 9 * cin; a.N = cout; a.ldx = cin; a.ldy = ldy; a.BN = BN;
    a.relu = relu; a.pool = 1; a.x_vec_ok = 1; a.f16 = f16;
    a.yh1 = reinterpret_cast<__half *>(yh1); a.yh2 = reinterpret_cast<__half *>(yh2); a.ldh = ldh;
    return gemm_launch(a, (cudaStream_t)stream);
}
// The TF32-split convolution (fp32 NHWC input) that ALSO writes its result as the FP16 planes the next layer reads with TMA
// (epnet_conv3x3_planes_tma); y may be NULL.  Used for the first convolution of the image stream (Cin = 3 -> 4).
EPNET_API int epnet_conv3x3_nhwc_tf32x3_planes(int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                                               const float *bias, int relu, float *y, int ldy, void *yh1, void *yh2, int ldh, void *stream)
{
    if (!yh1) return EPNET_ERR_BAD_ARG;
    return conv_entry(0, b, h, w, cin, cout, stride, x, wpack, BN, bias, relu, y, ldy, stream, yh1, yh2, ldh);
}
EPNET_API int epnet_conv3x3_nhwc_tf32x3(int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                                        const float *bias, int relu, float *y, int ldy, void *stream)
{
    return conv_entry(0, b, h, w, cin, cout, stride, x, wpack, BN, bias, relu, y, ldy, stream);
}
EPNET_API int epnet_conv3x3_nhwc_f16x3(int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                                       const float *bias, int relu, float *y, int ldy, void *stream)
{
    return conv_entry(1, b, h, w, cin, cout, stride, x, wpack, BN, bias, relu, y, ldy, stream);
}

// Transposed convolution with kernel == stride == k (ConvTranspose2d, no overlap between patches) on an NHWC map: x rows are the
// B*h*w input pixels (ldx floats apart, cin read), wpack packs W reordered to rows (ky, kx, o) x cin; the epilogue scatters each
// pixel's k x k x co patch straight into out, an NHWC buffer of (B, h*k, w*k) pixels ldo floats apart (out already offset to the
// first channel of this map's slice of a concatenation); bias (k*k*co, one value per GEMM column) or NULL.  Replaces a GEMM +
// pixel-shuffle pass.
static int deconv_entry(int f16, int b, int h, int w, int cin, int k, int co, const float *x, int ldx, const float *wpack, int BN,
                        const float *bias, int relu, float *out, int ldo, void *stream)
{
    using namespace epnet;
    if (b < 0 || h <= 0 || w <= 0 || cin <= 0 || k <= 0 || co <= 0 || (co % 4) != 0 || !x || !wpack || !out || ldx < cin || ldo < co)
        return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 256 || (BN % 16) != 0 || (ldo % 4) != 0) return EPNET_ERR_BAD_ARG;
    if (((reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return EPNET_ERR_BAD_ARG;
    const long long rows = (long long)b * h * w, cols = (long long)k * k * co;
    if (rows >= (1ll << 31) || cols >= (1ll << 24)) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    GemmArgs a = {};
    a.x = x; a.wpack = wpack; a.bias = bias; a.y = out;
    a.L = (int)rows; a.K = cin; a.N = (int)cols; a.ldx = ldx; a.ldy = ldo; a.BN = BN;
    a.relu = relu; a.pool = 1; a.H = h; a.W = w; a.dk = k; a.dco = co; a.f16 = f16;
    a.x_vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0);
    return gemm_launch(a, (cudaStream_t)stream);
}
EPNET_API int epnet_deconv_nhwc_tf32x3(int b, int h, int w, int cin, int k, int co, const float *x, int ldx, const float *wpack, int BN,
                                       const float *bias, int relu, float *out, int ldo, void *stream)
{
    return deconv_entry(0, b, h, w, cin, k, co, x, ldx, wpack, BN, bias, relu, out, ldo, stream);
}
EPNET_API int epnet_deconv_nhwc_f16x3(int b, int h, int w, int cin, int k, int co, const float *x, int ldx, const float *wpack, int BN,
                                      const float *bias, int relu, float *out, int ldo, void *stream)
{
    return deconv_entry(1, b, h, w, cin, k, co, x, ldx, wpack, BN, bias, relu, out, ldo, stream);
}


// ---- gemm_f16x3_tma_kernel: host side ---------------------------------------------------------------------------------------
namespace epnet {
// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library does not link libcuda
static PFN_cuTensorMapEncodeTiled tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

constexpr int kTmaTileW = 16, kTmaTileH = 8;  // output pixels per CTA tile: 16 x 8 = 128 GEMM rows

// planes xh1/xh2: convolution mode (a.conv) NHWC FP16 images (b, H, W, ldx), else (L, ldx) rows
static int gemm_launch_tma(GemmArgs &a, const void *xh1, const void *xh2, int ldx, int batch, cudaStream_t st)
{
    static_assert(sizeof(CUtensorMap) == 128, "TmaMaps holds raw CUtensorMap bytes");
    PFN_cuTensorMapEncodeTiled encode = tensor_map_encoder();
    if (!encode) return (int)cudaErrorNotSupported;
    a.n_kblocks = (a.K + kHfBK - 1) / kHfBK;
    a.kcopy = a.K;
    a.f16 = 1;
    a.corr_scale = 1.0f / 2048.0f;
    const int n_tiles = (a.N + a.BN - 1) / a.BN;
    TmaMaps maps;
    const void *planes[2] = {xh1, xh2};
    for (int i = 0; i < 2; ++i) {
        CUtensorMap *m = reinterpret_cast<CUtensorMap *>(i == 0 ? maps.h1 : maps.h2);
        CUresult r;
        if (a.conv) {
            const cuuint64_t gdim[4] = {(cuuint64_t)(1 << a.cin_shift), (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)batch};
            const cuuint64_t gstr[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)a.W * ldx * 2, (cuuint64_t)a.H * a.W * ldx * 2};
            const cuuint32_t box[4] = {(cuuint32_t)kHfBK, (cuuint32_t)(kTmaTileW * a.stride), (cuuint32_t)(kTmaTileH * a.stride), 1u};
            const cuuint32_t estr[4] = {1u, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1u};
            r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(planes[i]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t gdim[2] = {(cuuint64_t)a.K, (cuuint64_t)a.L};
            const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
            const cuuint32_t box[2] = {(cuuint32_t)kHfBK, (cuuint32_t)kGmBM};
            const cuuint32_t estr[2] = {1u, 1u};
            r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(planes[i]), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) return EPNET_ERR_BAD_ARG;
    }
    int m_tiles;
    if (a.conv) {
        a.tw = kTmaTileW; a.th = kTmaTileH;
        a.tiles_x = (a.Wo + kTmaTileW - 1) / kTmaTileW;
        a.tiles_y = (a.Ho + kTmaTileH - 1) / kTmaTileH;
        m_tiles = batch * a.tiles_x * a.tiles_y;
    } else {
        m_tiles = (a.L + kGmBM - 1) / kGmBM;
    }
    const long long total_tiles = (long long)m_tiles * n_tiles;
    dim3 grid((unsigned)(total_tiles < kSmCount ? total_tiles : kSmCount));  // persistent: one CTA per SM
    const size_t sb = 2 * (size_t)kGmBM * 128 + 2 * (size_t)a.BN * 128;
    int st_ = (int)((200 * 1024) / sb);
    if (st_ > kGmMaxStages) st_ = kGmMaxStages;
    if (st_ < 1) st_ = 1;  // the ring runs across tiles: keep every stage even when a tile has fewer k-blocks
    a.stages = st_;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, (const void *)gemm_f16x3_tma_kernel);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute((const void *)gemm_f16x3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return (int)e;
    gemm_f16x3_tma_kernel<<<grid, kTmaThreads, sb * st_ + 1024, st>>>(a, maps, m_tiles, n_tiles);
    EPNET_RETURN_LAUNCH_STATUS();
}

static bool planes_ok(const void *h1, const void *h2, int ld)
{
    return h1 && h2 && ((reinterpret_cast<uintptr_t>(h1) | reinterpret_cast<uintptr_t>(h2)) & 15) == 0 && ld > 0 && (ld % 8) == 0;
}
}  // namespace epnet

// 3x3 convolution, padding 1, stride 1 or 2, whose input is the pair of FP16 planes a previous layer's epilogue wrote
// (xh1, xh2: NHWC (b, h, w, ldx) halfs, cin % 64 == 0 channels read).  wpack = FP16 planes of W reordered to (Cout, ky, kx, Cin) in
// k-blocks of 64 (epnet_b200/gemm.py).  Outputs: y (b*Ho*Wo, ldy) fp32 NHWC and/or the FP16 planes yh1/yh2 (b*Ho*Wo, ldh) for the next
// layer; either may be NULL, not both.
EPNET_API int epnet_conv3x3_planes_tma(int b, int h, int w, int cin, int cout, int stride, const void *xh1, const void *xh2, int ldx,
                                       const float *wpack, int BN, const float *bias, int relu, float *y, int ldy, void *yh1, void *yh2,
                                       int ldh, void *stream)
{
    using namespace epnet;
    if (b < 0 || h <= 0 || w <= 0 || cin < 64 || (cin & (cin - 1)) != 0 || cout <= 0 || (stride != 1 && stride != 2) || !wpack) return EPNET_ERR_BAD_ARG;
    if (!planes_ok(xh1, xh2, ldx) || ldx < cin || BN < 16 || BN > 256 || (BN % 16) != 0 || h >= 32768 || w >= 32768) return EPNET_ERR_BAD_ARG;
    if ((!y && !yh1) || (y && ldy < cout) || (yh1 && (!planes_ok(yh1, yh2, ldh) || ldh < cout || (cout % 8) != 0))) return EPNET_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(wpack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    GemmArgs a = {};
    a.wpack = wpack; a.bias = bias; a.y = y;
    a.yh1 = reinterpret_cast<__half *>(yh1); a.yh2 = reinterpret_cast<__half *>(yh2); a.ldh = ldh;
    a.conv = 1; a.H = h; a.W = w; a.stride = stride;
    a.Ho = (h + 2 - 3) / stride + 1; a.Wo = (w + 2 - 3) / stride + 1;
    while ((1 << a.cin_shift) < cin) ++a.cin_shift;
    const long long rows = (long long)b * a.Ho * a.Wo;
    if (rows >= (1ll << 31)) return EPNET_ERR_BAD_ARG;
    a.L = (int)rows; a.K = 9 * cin; a.N = cout; a.ldx = ldx; a.ldy = ldy; a.BN = BN;
    a.relu = relu; a.pool = 1;
    return gemm_launch_tma(a, xh1, xh2, ldx, b, (cudaStream_t)stream);
}

// Y = act(X W^T + b) with X given as FP16 planes (L, ldx) (K read, ldx % 8 == 0): plain rows out (y fp32 and/or planes), optional
// max-pool over `pool` consecutive rows (fp32 output only).
EPNET_API int epnet_gemm_planes_tma(int L, int K, int N, const void *xh1, const void *xh2, int ldx, const float *wpack, int BN,
                                    const float *bias, int relu, int pool, float *y, int ldy, void *yh1, void *yh2, int ldh, void *stream)
{
    using namespace epnet;
    if (L < 0 || K <= 0 || N <= 0 || !wpack || !planes_ok(xh1, xh2, ldx) || ldx < K) return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 256 || (BN % 16) != 0 || (reinterpret_cast<uintptr_t>(wpack) & 15) != 0) return EPNET_ERR_BAD_ARG;
    if (pool < 1 || pool > 32 || (32 % pool) != 0 || (L % pool) != 0) return EPNET_ERR_BAD_ARG;
    if ((!y && !yh1) || (y && ldy < N) || (yh1 && (pool != 1 || !planes_ok(yh1, yh2, ldh) || ldh < N || (N % 8) != 0))) return EPNET_ERR_BAD_ARG;
    if (L == 0) return EPNET_OK;
    GemmArgs a = {};
    a.wpack = wpack; a.bias = bias; a.y = y;
    a.yh1 = reinterpret_cast<__half *>(yh1); a.yh2 = reinterpret_cast<__half *>(yh2); a.ldh = ldh;
    a.L = L; a.K = K; a.N = N; a.ldx = ldx; a.ldy = ldy; a.BN = BN; a.relu = relu; a.pool = pool;
    return gemm_launch_tma(a, xh1, xh2, ldx, 1, (cudaStream_t)stream);
}

// Transposed convolution (kernel == stride == k) whose input map is given as FP16 planes (b*h*w, ldx): see epnet_deconv_nhwc_tf32x3.
EPNET_API int epnet_deconv_planes_tma(int b, int h, int w, int cin, int k, int co, const void *xh1, const void *xh2, int ldx,
                                      const float *wpack, int BN, const float *bias, int relu, float *out, int ldo, void *stream)
{
    using namespace epnet;
    if (b < 0 || h <= 0 || w <= 0 || cin <= 0 || k <= 0 || co <= 0 || (co % 4) != 0 || !wpack || !out || !planes_ok(xh1, xh2, ldx) || ldx < cin || ldo < co)
        return EPNET_ERR_BAD_ARG;
    if (BN < 16 || BN > 256 || (BN % 16) != 0 || (ldo % 4) != 0) return EPNET_ERR_BAD_ARG;
    if (((reinterpret_cast<uintptr_t>(wpack) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return EPNET_ERR_BAD_ARG;
    const long long rows = (long long)b * h * w, cols = (long long)k * k * co;
    if (rows >= (1ll << 31) || cols >= (1ll << 24)) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    GemmArgs a = {};
    a.wpack = wpack; a.bias = bias; a.y = out;
    a.L = (int)rows; a.K = cin; a.N = (int)cols; a.ldx = ldx; a.ldy = ldo; a.BN = BN;
    a.relu = relu; a.pool = 1; a.H = h; a.W = w; a.dk = k; a.dco = co;
    return gemm_launch_tma(a, xh1, xh2, ldx, 1, (cudaStream_t)stream);
}

namespace epnet {
// device address of the range-guard flag for kernels of other translation units that write activations (sa_first_level.cu)
unsigned int *gemm_overflow_flag()
{
    void *p = nullptr;
    return cudaGetSymbolAddress(&p, g_gemm_overflow) == cudaSuccess ? static_cast<unsigned int *>(p) : nullptr;
}
}  // namespace epnet

// FP16-split range guard (see g_gemm_overflow): asynchronous read of the per-device flag into host memory (pinned for a truly
// asynchronous copy) and reset, both ordered on `stream`.  No reference counterpart: the reference computes in fp32.
EPNET_API int epnet_gemm_overflow_read(unsigned int *host_dst, void *stream)
{
    if (!host_dst) return EPNET_ERR_BAD_ARG;
    return (int)cudaMemcpyFromSymbolAsync(host_dst, epnet::g_gemm_overflow, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
EPNET_API int epnet_gemm_overflow_reset(void *stream)
{
    void *p = nullptr;
    cudaError_t e = cudaGetSymbolAddress(&p, epnet::g_gemm_overflow);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemsetAsync(p, 0, sizeof(unsigned int), (cudaStream_t)stream);
}
