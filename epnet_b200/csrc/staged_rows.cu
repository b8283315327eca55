// Channel-major indexed row reads with the source rows resident in shared memory.
//
// gather_points / group_points (out[b,c,e] = points[b,c,idx[b,e]]) and three_interpolate
// (out[b,c,i] = sum_k w[b,i,k] * points[b,c,idx[b,i,k]]) keep the reference's (B,C,N) layout at the
// C ABI (/root/reference/pointnet2_lib/pointnet2/src/group_points_gpu.cu:47-66, sampling_gpu.cu:8-24,
// interpolate_gpu.cu:77-97).  In that layout every looked-up value is one 4-byte word of a different
// 32-byte L2 sector, so a kernel that gathers from global memory is bound by L2 sector throughput at
// 1/8 efficiency, however the threads are arranged.  Here a CTA first copies R whole channel rows
// (they are contiguous: R*len floats) into shared memory with bulk (TMA) copies, then serves every
// lookup from shared memory at 4-byte granularity; the indices (and weights) of 4 consecutive outputs
// are loaded once as 128-bit words and reused for the R rows, and each row costs one 128-bit
// streaming store.  What is left is the HBM write stream plus idx/weight re-reads of 1/R per row.
// A row longer than shared memory (N > ~51k) is staged in part: lookups below the staged length come from
// shared memory, the rest from global memory, which removes that fraction of the sector traffic.
// Arithmetic is the same fma chain as the plain kernels, so results are bit-identical to them.
#include "common.cuh"

namespace epnet {

constexpr int kStagedThreads = 512;        // gather: 64 registers per thread, two CTAs per SM
constexpr int kStagedThreadsInterp = 256;  // interpolate keeps 12 indices + 12 weights live: 128 registers, two CTAs per SM
constexpr int kStagedChunk = 32768;  // bytes per bulk copy

template <bool kPartial>
__device__ __forceinline__ float row_at(const float *s_row, const float *g_row, int staged_len, int i)
{
    if (kPartial && i >= staged_len) return __ldg(g_row + i);
    return s_row[i];
}

// kPartial: rows == 1 and only the first staged_len floats of the row are in shared memory
template <bool kInterp, int kThreads, bool kPartial>
__global__ void __launch_bounds__(kThreads, 2)
staged_rows_kernel(int c, int len, int staged_len, long long e_total, int rows, long long e_per_cta, const float *__restrict__ src,
                   const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out)
{
    extern __shared__ __align__(128) float s_rows[];
    __shared__ uint64_t bar;

    const int scene = blockIdx.z;
    const int c_begin = blockIdx.y * rows;
    const int r_count = min(rows, c - c_begin);
    src += ((size_t)scene * c + c_begin) * len;
    out += ((size_t)scene * c + c_begin) * e_total;
    idx += (size_t)scene * e_total * (kInterp ? 3 : 1);
    if (kInterp) weight += (size_t)scene * e_total * 3;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        const size_t bytes = (kPartial ? (size_t)staged_len : (size_t)r_count * len) * sizeof(float);  // a multiple of 16 (launcher)
        mbar_arrive_expect_tx(&bar, (uint32_t)bytes);
        for (size_t off = 0; off < bytes; off += kStagedChunk)
            bulk_g2s(reinterpret_cast<char *>(s_rows) + off, reinterpret_cast<const char *>(src) + off,
                     (uint32_t)min((size_t)kStagedChunk, bytes - off), &bar);
    }
    __syncthreads();  // the barrier is initialised before anyone polls it

    const long long e_begin = (long long)blockIdx.x * e_per_cta;
    const long long e_end = min(e_total, e_begin + e_per_cta);
    bool staged = false;
    for (long long e0 = e_begin + threadIdx.x * 4; e0 < e_end; e0 += kThreads * 4) {
        if (!kInterp) {
            const int4 id = __ldg(reinterpret_cast<const int4 *>(idx + e0));
            if (!staged) { mbar_wait(&bar, 0); staged = true; }
            const float *row = s_rows, *g = src;
            float *o = out + e0;
            for (int r = 0; r < r_count; ++r, row += len, g += len, o += e_total)
                __stcs(reinterpret_cast<float4 *>(o),
                       make_float4(row_at<kPartial>(row, g, staged_len, id.x), row_at<kPartial>(row, g, staged_len, id.y),
                                   row_at<kPartial>(row, g, staged_len, id.z), row_at<kPartial>(row, g, staged_len, id.w)));
        } else {
            int id[12];
            float w[12];
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const int4 a = __ldg(reinterpret_cast<const int4 *>(idx + 3 * e0) + v);
                const float4 f = __ldg(reinterpret_cast<const float4 *>(weight + 3 * e0) + v);
                id[4 * v] = a.x; id[4 * v + 1] = a.y; id[4 * v + 2] = a.z; id[4 * v + 3] = a.w;
                w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
            }
            if (!staged) { mbar_wait(&bar, 0); staged = true; }
            const float *row = s_rows, *g = src;
            float *o = out + e0;
            for (int r = 0; r < r_count; ++r, row += len, g += len, o += e_total) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)  // interpolate_gpu.cu:96 as compiled: fma(w2,p2, fma(w0,p0, w1*p1))
                    v[e] = __fmaf_rn(w[3 * e + 2], row_at<kPartial>(row, g, staged_len, id[3 * e + 2]),
                                     __fmaf_rn(w[3 * e], row_at<kPartial>(row, g, staged_len, id[3 * e]),
                                               __fmul_rn(w[3 * e + 1], row_at<kPartial>(row, g, staged_len, id[3 * e + 1]))));
                __stcs(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
            }
        }
    }
    if (!staged) mbar_wait(&bar, 0);  // never leave with copies into this CTA's shared memory in flight
}

// Returns EPNET_OK after launching, a cudaError, or kStagedNotApplicable when the plain kernel should run:
// rows that do not fit in shared memory, unaligned operands, or so little work that staging cannot pay.
int launch_staged_rows(bool interp, int b, int c, int len, long long e_total, const float *src, const int *idx, const float *weight,
                       float *out, cudaStream_t st)
{
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(weight);
    if ((al & 15) || (len % 4) || (e_total % 4) || len <= 0) return kStagedNotApplicable;
    if (e_total < 2 * (long long)len || (long long)c * e_total < (1 << 20)) return kStagedNotApplicable;
    const size_t row_bytes = (size_t)len * sizeof(float);
    const size_t small = 100 * 1024, large = 200 * 1024;  // two CTAs per SM when the rows allow it
    // interpolate re-reads 24 bytes of idx/weight per output and row group: prefer more rows per CTA over a second CTA per SM
    const size_t budget = (row_bytes <= small && !(interp && small / row_bytes < 4)) ? small : large;
    const bool partial = row_bytes > budget;
    if (partial && row_bytes > 4 * budget) return kStagedNotApplicable;  // less than a quarter of the lookups would be served
    const int rows = partial ? 1 : (int)min((size_t)min(c, 16), budget / row_bytes);
    const int staged_len = partial ? (int)(budget / sizeof(float)) : len;
    const int row_groups = (c + rows - 1) / rows;
    // split the outputs of a row group over several CTAs until the GPU is covered twice, but keep each CTA's
    // output at least as large as what it stages
    long long splits = (2LL * kSmCount + (long long)b * row_groups - 1) / ((long long)b * row_groups);
    splits = max(1LL, min(splits, e_total / max(len, 2048)));
    long long e_per_cta = (e_total + splits - 1) / splits;
    const int quantum = (interp ? kStagedThreadsInterp : kStagedThreads) * 4;
    e_per_cta = (e_per_cta + quantum - 1) / quantum * quantum;
    splits = (e_total + e_per_cta - 1) / e_per_cta;
    if (splits > 65535 || row_groups > 65535 || b > 65535) return kStagedNotApplicable;

    const size_t smem = partial ? budget : (size_t)rows * row_bytes;
    auto kernel = interp ? (partial ? staged_rows_kernel<true, kStagedThreadsInterp, true> : staged_rows_kernel<true, kStagedThreadsInterp, false>)
                         : (partial ? staged_rows_kernel<false, kStagedThreads, true> : staged_rows_kernel<false, kStagedThreads, false>);
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return (int)e;
    if (fa.maxDynamicSharedSizeBytes < (int)large) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)large);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)splits, row_groups, b);
    kernel<<<grid, interp ? kStagedThreadsInterp : kStagedThreads, smem, st>>>(c, len, staged_len, e_total, rows, e_per_cta, src, idx, weight, out);
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet
