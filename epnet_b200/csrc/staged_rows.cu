// Channel-major indexed row reads with the source rows resident in shared memory.
//
// gather_points / group_points (out[b,c,e] = points[b,c,idx[b,e]]) and three_interpolate
// (out[b,c,i] = sum_k w[b,i,k] * points[b,c,idx[b,i,k]]) keep the reference's (B,C,N) layout at the
// C ABI (/root/reference/pointnet2_lib/pointnet2/src/group_points_gpu.cu:47-66, sampling_gpu.cu:8-24,
// interpolate_gpu.cu:77-97).  In that layout every looked-up value is one 4-byte word of a different
// 32-byte L2 sector, so a kernel that gathers from global memory is bound by the SM's one-sector-per-clock
// request rate (281 G lookups/s = 0.17 of HBM for 4-byte outputs), however the threads are arranged.  Here a
// CTA first copies R whole channel rows (they are contiguous: R*len floats) into shared memory with bulk (TMA)
// copies, then serves every lookup from shared memory at 4-byte granularity (random banks: ~9 lookups per clock);
// the indices (and weights) of 4 consecutive outputs are loaded once as 128-bit words and reused for the R rows,
// and each row costs one 128-bit streaming store.  What is left is the HBM write stream plus idx/weight re-reads
// of 1/R per row; those come from L2 with ~1 us of latency, so the loads of the NEXT group of outputs are issued
// before the lookups of the current one (software pipeline, one group ahead for interpolate, two for the gathers)
// and a CTA that has the SM to itself runs 1024 (gather) / 512 (interpolate) threads.
// A row longer than shared memory (N > 57344) is staged in part: lookups below the staged length come from shared memory, the rest
// from global memory, which removes that fraction of the sector traffic; with enough work such rows go through transposed_rows.cu
// instead.  (Measured and rejected: staging a long row part by part and writing, for each part, the outputs whose index falls
// inside it -- 0.18 of HBM at N = 65536 against 0.47 for partial staging: the indices are read once per part and the predicated
// 4-byte stores cost more than the global lookups they avoid.)
// Arithmetic is the same fma chain as the plain kernels, so results are bit-identical to them.
#include "common.cuh"

namespace epnet {

constexpr int kStagedChunk = 32768;  // bytes per bulk copy
constexpr size_t kStagedSmall = 100 * 1024, kStagedLarge = 224 * 1024;  // two CTAs per SM / one CTA per SM

template <bool kPart>
__device__ __forceinline__ float row_at(const float *s_row, const float *g_row, int staged_len, int i)
{
    if (kPart && i >= staged_len) return __ldg(g_row + i);
    return s_row[i];
}

__device__ __forceinline__ void stage_bytes(float *s_rows, const float *src, size_t bytes, uint64_t *bar)
{
    mbar_arrive_expect_tx(bar, (uint32_t)bytes);
    for (size_t off = 0; off < bytes; off += kStagedChunk)
        bulk_g2s(reinterpret_cast<char *>(s_rows) + off, reinterpret_cast<const char *>(src) + off,
                 (uint32_t)min((size_t)kStagedChunk, bytes - off), bar);
}

// ---- gather_points / group_points -----------------------------------------------------------------------------------------------
// grid (splits, row groups, scenes); a thread owns groups of 4 consecutive outputs, two groups per trip.
template <int kThreads, int kMinBlocks, bool kPart>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
staged_gather_kernel(int c, int len, int staged_len, long long e_total, int rows, long long e_per_cta, const float *__restrict__ src,
                     const int *__restrict__ idx, float *__restrict__ out)
{
    extern __shared__ __align__(128) float s_rows[];
    __shared__ uint64_t bar;

    const int scene = blockIdx.z;
    const int c_begin = blockIdx.y * rows;
    const int r_count = min(rows, c - c_begin);
    src += ((size_t)scene * c + c_begin) * len;
    out += ((size_t)scene * c + c_begin) * e_total;
    idx += (size_t)scene * e_total;

    if (threadIdx.x == 0) {  // sizes are multiples of 16 (launcher)
        mbar_init(&bar, 1);
        mbar_fence_init();
        stage_bytes(s_rows, src, (kPart ? (size_t)staged_len : (size_t)r_count * len) * sizeof(float), &bar);
    }
    __syncthreads();  // the barrier is initialised before anyone polls it

    const long long e_begin = (long long)blockIdx.x * e_per_cta;
    const long long e_end = min(e_total, e_begin + e_per_cta);
    const long long stride = (long long)kThreads * 4;  // the two groups of a trip are one stride apart: every warp access stays contiguous
    const int4 none = make_int4(0, 0, 0, 0);
    long long e0 = e_begin + threadIdx.x * 4;
    int4 cur0 = e0 < e_end ? __ldg(reinterpret_cast<const int4 *>(idx + e0)) : none;
    int4 cur1 = e0 + stride < e_end ? __ldg(reinterpret_cast<const int4 *>(idx + e0 + stride)) : none;
    mbar_wait(&bar, 0);  // never leave with copies into this CTA's shared memory in flight: every thread waits
    for (; e0 < e_end; e0 += 2 * stride) {
        const long long n0 = e0 + 2 * stride, n1 = e0 + 3 * stride;
        const int4 nxt0 = n0 < e_end ? __ldg(reinterpret_cast<const int4 *>(idx + n0)) : none;
        const int4 nxt1 = n1 < e_end ? __ldg(reinterpret_cast<const int4 *>(idx + n1)) : none;
        const bool two = e0 + stride < e_end;
        const float *row = s_rows, *g = src;
        float *o = out + e0;
        for (int r = 0; r < r_count; ++r, row += len, g += len, o += e_total) {
            __stcs(reinterpret_cast<float4 *>(o), make_float4(row_at<kPart>(row, g, staged_len, cur0.x), row_at<kPart>(row, g, staged_len, cur0.y),
                                                             row_at<kPart>(row, g, staged_len, cur0.z), row_at<kPart>(row, g, staged_len, cur0.w)));
            if (two)
                __stcs(reinterpret_cast<float4 *>(o + stride),
                       make_float4(row_at<kPart>(row, g, staged_len, cur1.x), row_at<kPart>(row, g, staged_len, cur1.y),
                                   row_at<kPart>(row, g, staged_len, cur1.z), row_at<kPart>(row, g, staged_len, cur1.w)));
        }
        cur0 = nxt0;
        cur1 = nxt1;
    }
}

// ---- three_interpolate ------------------------------------------------------------------------------------------------------------
template <int kThreads, int kMinBlocks, bool kPart>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
staged_interp_kernel(int c, int len, int staged_len, long long e_total, int rows, long long e_per_cta, const float *__restrict__ src,
                     const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out)
{
    extern __shared__ __align__(128) float s_rows[];
    __shared__ uint64_t bar;

    const int scene = blockIdx.z;
    const int c_begin = blockIdx.y * rows;
    const int r_count = min(rows, c - c_begin);
    src += ((size_t)scene * c + c_begin) * len;
    out += ((size_t)scene * c + c_begin) * e_total;
    idx += (size_t)scene * e_total * 3;
    weight += (size_t)scene * e_total * 3;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        stage_bytes(s_rows, src, (kPart ? (size_t)staged_len : (size_t)r_count * len) * sizeof(float), &bar);
    }
    __syncthreads();  // the barrier is initialised before anyone polls it

    const long long e_begin = (long long)blockIdx.x * e_per_cta;
    const long long e_end = min(e_total, e_begin + e_per_cta);
    const long long stride = (long long)kThreads * 4;
    long long e0 = e_begin + threadIdx.x * 4;
    int4 ci[3];
    float4 cw[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        ci[v] = e0 < e_end ? __ldg(reinterpret_cast<const int4 *>(idx + 3 * e0) + v) : make_int4(0, 0, 0, 0);
        cw[v] = e0 < e_end ? __ldg(reinterpret_cast<const float4 *>(weight + 3 * e0) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mbar_wait(&bar, 0);  // never leave with copies into this CTA's shared memory in flight: every thread waits
    for (; e0 < e_end; e0 += stride) {
        const long long e1 = e0 + stride;
        int4 ni[3];
        float4 nw[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            ni[v] = e1 < e_end ? __ldg(reinterpret_cast<const int4 *>(idx + 3 * e1) + v) : ci[v];
            nw[v] = e1 < e_end ? __ldg(reinterpret_cast<const float4 *>(weight + 3 * e1) + v) : cw[v];
        }
        const int id[12] = {ci[0].x, ci[0].y, ci[0].z, ci[0].w, ci[1].x, ci[1].y, ci[1].z, ci[1].w, ci[2].x, ci[2].y, ci[2].z, ci[2].w};
        const float w[12] = {cw[0].x, cw[0].y, cw[0].z, cw[0].w, cw[1].x, cw[1].y, cw[1].z, cw[1].w, cw[2].x, cw[2].y, cw[2].z, cw[2].w};
        const float *row = s_rows, *g = src;
        float *o = out + e0;
        for (int r = 0; r < r_count; ++r, row += len, g += len, o += e_total) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)  // interpolate_gpu.cu:96 as compiled: fma(w2,p2, fma(w0,p0, w1*p1))
                v[e] = __fmaf_rn(w[3 * e + 2], row_at<kPart>(row, g, staged_len, id[3 * e + 2]),
                                 __fmaf_rn(w[3 * e], row_at<kPart>(row, g, staged_len, id[3 * e]),
                                           __fmul_rn(w[3 * e + 1], row_at<kPart>(row, g, staged_len, id[3 * e + 1]))));
            __stcs(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            ci[v] = ni[v];
            cw[v] = nw[v];
        }
    }
}

template <typename K>
static cudaError_t allow_large_smem(K kernel)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    if (fa.maxDynamicSharedSizeBytes < (int)kStagedLarge)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStagedLarge);
    return e;
}

bool staged_row_fits(int len) { return len > 0 && len % 4 == 0 && (size_t)len * sizeof(float) <= kStagedLarge; }

// Returns EPNET_OK after launching, a cudaError, or kStagedNotApplicable when the plain kernel should run:
// unaligned operands, rows far longer than shared memory, or so little work that staging cannot pay.
int launch_staged_rows(bool interp, int b, int c, int len, long long e_total, const float *src, const int *idx, const float *weight,
                       float *out, cudaStream_t st)
{
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(weight);
    if ((al & 15) || (len % 4) || (e_total % 4) || len <= 0) return kStagedNotApplicable;
    if (e_total < 2 * (long long)len || (long long)c * e_total < (1 << 20)) return kStagedNotApplicable;
    const size_t row_bytes = (size_t)len * sizeof(float);
    // two CTAs per SM when the rows allow it; interpolate re-reads 24 bytes of idx/weight per output and row group: prefer more rows per CTA
    const bool two_per_sm = row_bytes <= kStagedSmall && !(interp && kStagedSmall / row_bytes < 4);
    const size_t budget = two_per_sm ? kStagedSmall : kStagedLarge;
    const bool partial = row_bytes > budget;
    if (partial && row_bytes > 4 * budget) return kStagedNotApplicable;  // less than a quarter of the lookups would be served
    const int rows = partial ? 1 : (int)min((size_t)min(c, 16), budget / row_bytes);
    const int staged_len = partial ? (int)(budget / sizeof(float)) : len;
    const int row_groups = (c + rows - 1) / rows;
    const int threads = interp ? (two_per_sm ? 256 : 512) : (two_per_sm ? 512 : 1024);
    // split the outputs of a row group over several CTAs until the GPU is covered twice, but keep each CTA's
    // output at least as large as what it stages
    const long long resident = (two_per_sm ? 2LL : 1LL) * kSmCount;
    long long splits = (2 * resident + (long long)b * row_groups - 1) / ((long long)b * row_groups);
    splits = max(1LL, min(splits, e_total / max(len, 2048)));
    long long e_per_cta = (e_total + splits - 1) / splits;
    const int quantum = threads * 4 * (interp ? 1 : 2);
    e_per_cta = (e_per_cta + quantum - 1) / quantum * quantum;
    splits = (e_total + e_per_cta - 1) / e_per_cta;
    if (splits > 65535 || row_groups > 65535 || b > 65535) return kStagedNotApplicable;

    const size_t smem = partial ? budget : (size_t)rows * row_bytes;
    dim3 grid((unsigned)splits, row_groups, b);
    cudaError_t e = cudaSuccess;
#define EPNET_LAUNCH_STAGED(KERNEL, ...)                                                  \
    do {                                                                                  \
        e = allow_large_smem(KERNEL);                                                     \
        if (e != cudaSuccess) return (int)e;                                              \
        KERNEL<<<grid, threads, smem, st>>>(c, len, staged_len, e_total, rows, e_per_cta, __VA_ARGS__); \
    } while (0)
    if (interp) {
        if (two_per_sm) EPNET_LAUNCH_STAGED((staged_interp_kernel<256, 2, false>), src, idx, weight, out);
        else if (!partial) EPNET_LAUNCH_STAGED((staged_interp_kernel<512, 1, false>), src, idx, weight, out);
        else EPNET_LAUNCH_STAGED((staged_interp_kernel<512, 1, true>), src, idx, weight, out);
    } else {
        if (two_per_sm) EPNET_LAUNCH_STAGED((staged_gather_kernel<512, 2, false>), src, idx, out);
        else if (!partial) EPNET_LAUNCH_STAGED((staged_gather_kernel<1024, 1, false>), src, idx, out);
        else EPNET_LAUNCH_STAGED((staged_gather_kernel<1024, 1, true>), src, idx, out);
    }
#undef EPNET_LAUNCH_STAGED
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet
