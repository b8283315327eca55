// Channel-major gathers and scatters served through a point-major copy of the table.
//
// group_points / gather_points / three_interpolate and their gradients keep the reference's (B,C,N) layout at the C ABI
// (/root/reference/pointnet2_lib/pointnet2/src/group_points_gpu.cu:8-66, sampling_gpu.cu:8-63, interpolate_gpu.cu:77-142).  In that
// layout a looked-up value is 4 bytes of a 32-byte sector and the SM issues one sector request per clock: 281 G lookups/s (0.17 of HBM)
// forward, and 212 G reductions/s backward whatever their width (tools/probes/red_probe.cu: red.global.add.f32 / .v2 / .v4 all run at
// 212-215 G requests/s, i.e. 860 G floats/s for .v4).  staged_rows.cu answers that for rows that fit one SM's shared memory.  Here, for
// any row length, the layout is changed inside the call:
//   forward : the table is transposed once into a stream-ordered scratch T[N][Cp] (Cp = C rounded up to 4); a CTA then gathers 32
//             outputs x 128 channels as whole 512-byte rows (128-bit loads, full sectors), turns the tile in shared memory (XOR-swizzled
//             128-bit chunks, conflict-free both ways) and writes 128-byte runs of 32 consecutive outputs per channel.
//   backward: a CTA reads 128-byte runs of grad_out per channel, turns the tile the same way and adds whole rows into a zeroed scratch
//             S[N][Cp] with red.global.add.v4.f32 (a quarter of the requests of the scalar scatter; equal neighbouring indices -- ball-query
//             padding -- are folded first); one transposing pass adds S into grad_points.
// Arithmetic: forward values are copied (gathers) or use the same fma chain as interpolate_gpu.cu:96 as compiled (bit-identical to the
// plain kernels); gradients are fp32 atomic sums in unspecified order like the reference's.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace epnet {

constexpr int kTrTile = 32;      // outputs per tile
constexpr int kTrChunk = 128;    // channels per tile
constexpr int kTrThreads = 256;  // 8 warps: 4 rows / 4 channel quads each

// ---- stream-ordered scratch from a per-device pool that keeps its memory between calls ----------------------------------------------
static cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t st)
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&pools[dev], &props);
            if (e != cudaSuccess) return e;
            unsigned long long keep = ~0ull;  // never hand the pages back at a synchronisation point
            e = cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaMallocFromPoolAsync(p, bytes, pools[dev], st);
}

// ---- (B,C,N) -> (B,N,Cp), zero in the padding channels ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cm_to_pm_kernel(int c, int cp, int n, const float *__restrict__ src, float *__restrict__ dst)
{
    __shared__ float tile[32][33];
    const int scene = blockIdx.z;
    src += (size_t)scene * c * n;
    dst += (size_t)scene * n * cp;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ch = c0 + r, i = n0 + tx;
        tile[r][tx] = (ch < c && i < n) ? __ldg(src + (size_t)ch * n + i) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = n0 + r, ch = c0 + tx;
        if (i < n && ch < cp) dst[(size_t)i * cp + ch] = tile[tx][r];
    }
}

// grad_points (B,C,N) += S (B,N,Cp)
__global__ void __launch_bounds__(256)
pm_add_to_cm_kernel(int c, int cp, int n, const float *__restrict__ src, float *__restrict__ dst)
{
    __shared__ float tile[32][33];
    const int scene = blockIdx.z;
    src += (size_t)scene * n * cp;
    dst += (size_t)scene * c * n;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = n0 + r, ch = c0 + tx;
        tile[r][tx] = (i < n && ch < cp) ? __ldcs(src + (size_t)i * cp + ch) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int ch = c0 + r, i = n0 + tx;
        if (ch < c && i < n) dst[(size_t)ch * n + i] += tile[tx][r];
    }
}

// ---- forward: out[c][e] = T[idx[e]][c]  or  fma(w2,T[i2][c], fma(w0,T[i0][c], w1*T[i1][c])) ------------------------------------------
// grid (tiles of 32 outputs, chunks of 128 channels, scenes)
template <bool kInterp>
__global__ void __launch_bounds__(kTrThreads)
pm_rows_to_cm_kernel(int c, int cp, long long e_total, int len, const float *__restrict__ table, const int *__restrict__ idx,
                     const float *__restrict__ weight, float *__restrict__ out)
{
    __shared__ float4 tile[kTrTile][kTrChunk / 4];  // [output][channel quad ^ output]
    const int scene = blockIdx.z;
    table += (size_t)scene * len * cp;
    out += (size_t)scene * c * e_total;
    idx += (size_t)scene * e_total * (kInterp ? 3 : 1);
    if (kInterp) weight += (size_t)scene * e_total * 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long e_base = (long long)blockIdx.x * kTrTile;
    const int c_base = blockIdx.y * kTrChunk;
    const int cq = c_base + 4 * lane;  // this lane's channel quad while gathering

#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = warp * 4 + k;
        const long long e = e_base + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < e_total && cq < cp) {
            if (!kInterp) {
                const int i = __ldg(idx + e);
                v = __ldg(reinterpret_cast<const float4 *>(table + (size_t)i * cp + cq));
            } else {
                const int i0 = __ldg(idx + 3 * e), i1 = __ldg(idx + 3 * e + 1), i2 = __ldg(idx + 3 * e + 2);
                const float w0 = __ldg(weight + 3 * e), w1 = __ldg(weight + 3 * e + 1), w2 = __ldg(weight + 3 * e + 2);
                const float4 p0 = __ldg(reinterpret_cast<const float4 *>(table + (size_t)i0 * cp + cq));
                const float4 p1 = __ldg(reinterpret_cast<const float4 *>(table + (size_t)i1 * cp + cq));
                const float4 p2 = __ldg(reinterpret_cast<const float4 *>(table + (size_t)i2 * cp + cq));
                v.x = __fmaf_rn(w2, p2.x, __fmaf_rn(w0, p0.x, __fmul_rn(w1, p1.x)));
                v.y = __fmaf_rn(w2, p2.y, __fmaf_rn(w0, p0.y, __fmul_rn(w1, p1.y)));
                v.z = __fmaf_rn(w2, p2.z, __fmaf_rn(w0, p0.z, __fmul_rn(w1, p1.z)));
                v.w = __fmaf_rn(w2, p2.w, __fmaf_rn(w0, p0.w, __fmul_rn(w1, p1.w)));
            }
        }
        tile[r][lane ^ r] = v;
    }
    __syncthreads();
    const long long e = e_base + lane;  // this lane's output while writing
    if (e >= e_total) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int q = warp + 8 * k;  // channel quad
        const int ch = c_base + 4 * q;
        if (ch >= c) break;
        const float4 v = tile[lane][q ^ lane];
        float *o = out + (size_t)ch * e_total + e;
        __stcs(o, v.x);
        if (ch + 1 < c) __stcs(o + e_total, v.y);
        if (ch + 2 < c) __stcs(o + 2 * e_total, v.z);
        if (ch + 3 < c) __stcs(o + 3 * e_total, v.w);
    }
}

__device__ __forceinline__ void red_add_v4(float *p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- backward: S[idx[e]][c] += g[c][e]  or  S[i_k][c] += g[c][e] * w_k ----------------------------------------------------------------
template <bool kInterp>
__global__ void __launch_bounds__(kTrThreads)
cm_scatter_to_pm_kernel(int c, int cp, long long e_total, int len, const float *__restrict__ grad_out, const int *__restrict__ idx,
                        const float *__restrict__ weight, float *__restrict__ acc_table)
{
    __shared__ float4 tile[kTrTile][kTrChunk / 4];
    const int scene = blockIdx.z;
    acc_table += (size_t)scene * len * cp;
    grad_out += (size_t)scene * c * e_total;
    idx += (size_t)scene * e_total * (kInterp ? 3 : 1);
    if (kInterp) weight += (size_t)scene * e_total * 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long e_base = (long long)blockIdx.x * kTrTile;
    const int c_base = blockIdx.y * kTrChunk;

    {
        const long long e = e_base + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int q = warp + 8 * k;
            const int ch = c_base + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < e_total && ch < c) {
                const float *g = grad_out + (size_t)ch * e_total + e;
                v.x = __ldcs(g);
                if (ch + 1 < c) v.y = __ldcs(g + e_total);
                if (ch + 2 < c) v.z = __ldcs(g + 2 * e_total);
                if (ch + 3 < c) v.w = __ldcs(g + 3 * e_total);
            }
            tile[lane][q ^ lane] = v;
        }
    }
    __syncthreads();
    const int cq = c_base + 4 * lane;
    if (cq >= cp) return;
    const int r0 = warp * 4;
    if (!kInterp) {
        // four consecutive outputs per warp: equal neighbouring indices (ball-query padding) are folded before the reduction
        int cur = -1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long e = e_base + r0 + k;
            if (e >= e_total) break;
            const int i = __ldg(idx + e);
            const float4 v = tile[r0 + k][lane ^ (r0 + k)];
            if (i == cur) {
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            } else {
                if (cur >= 0) red_add_v4(acc_table + (size_t)cur * cp + cq, acc);
                cur = i;
                acc = v;
            }
        }
        if (cur >= 0) red_add_v4(acc_table + (size_t)cur * cp + cq, acc);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long e = e_base + r0 + k;
            if (e >= e_total) break;
            const float4 v = tile[r0 + k][lane ^ (r0 + k)];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int i = __ldg(idx + 3 * e + t);
                const float w = __ldg(weight + 3 * e + t);
                red_add_v4(acc_table + (size_t)i * cp + cq, make_float4(v.x * w, v.y * w, v.z * w, v.w * w));
            }
        }
    }
}

static bool grid_ok(long long x, long long y, long long z) { return x <= 0x7fffffffLL && y <= 65535 && z <= 65535; }

// experiment switch (tools/, tests/perf): EPNET_CM_ROWS=staged keeps this path off, =transposed takes it whenever it can run
static int path_override()
{
    static const int v = [] { const char *s = getenv("EPNET_CM_ROWS"); return !s ? 0 : s[0] == 's' ? 1 : s[0] == 't' ? 2 : 0; }();
    return v;
}

// out (B,C,E) from the (B,C,len) table.  Returns kStagedNotApplicable when another kernel should run.
int launch_transposed_gather(bool interp, bool staged_fits, int b, int c, int len, long long e_total, const float *src, const int *idx,
                             const float *weight, float *out, cudaStream_t st)
{
    const int force = path_override();
    if (force == 1 || len <= 0) return kStagedNotApplicable;
    if (force != 2) {
        // the transposition moves the table twice: worth it when the outputs outnumber the table and the work is not tiny;
        // gathers whose rows fit shared memory are better served there (one lookup per output), interpolation (three) is not
        if (e_total < (long long)len) return kStagedNotApplicable;
        if (!interp && (staged_fits || (long long)c * e_total < (1 << 21))) return kStagedNotApplicable;
        // short rows (feature propagation at 1024 -> 4096 points) are served faster from shared memory: 14.8 vs 18.8 us at B=2, C=512
        if (interp && (long long)b * c * e_total < (1 << 23)) return kStagedNotApplicable;
    }
    const int cp = (c + 3) / 4 * 4;
    const long long tiles = (e_total + kTrTile - 1) / kTrTile;
    const int chunks = (c + kTrChunk - 1) / kTrChunk;
    if (!grid_ok(tiles, chunks, b) || !grid_ok((len + 31) / 32, (cp + 31) / 32, b)) return kStagedNotApplicable;
    float *table = nullptr;
    cudaError_t e = scratch_alloc(reinterpret_cast<void **>(&table), (size_t)b * len * cp * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    cm_to_pm_kernel<<<dim3((len + 31) / 32, (cp + 31) / 32, b), 256, 0, st>>>(c, cp, len, src, table);
    if (interp)
        pm_rows_to_cm_kernel<true><<<dim3((unsigned)tiles, chunks, b), kTrThreads, 0, st>>>(c, cp, e_total, len, table, idx, weight, out);
    else
        pm_rows_to_cm_kernel<false><<<dim3((unsigned)tiles, chunks, b), kTrThreads, 0, st>>>(c, cp, e_total, len, table, idx, weight, out);
    e = cudaGetLastError();
    const cudaError_t f = cudaFreeAsync(table, st);
    return e != cudaSuccess ? (int)e : f != cudaSuccess ? (int)f : EPNET_OK;
}

// grad_points (B,C,len) += scatter of grad_out (B,C,E)
int launch_transposed_scatter(bool interp, int b, int c, int len, long long e_total, const float *grad_out, const int *idx,
                              const float *weight, float *grad_points, cudaStream_t st)
{
    const int force = path_override();
    if (force == 1 || len <= 0) return kStagedNotApplicable;
    // the scratch costs a memset and a transposing add over ALL len rows (16 bytes of traffic per table element) to save three
    // quarters of the reduction requests: worth it when the entries outnumber the table rows (measured at N = 131072, E = 16384:
    // 78 us through the scratch against 17 us for the scalar scatter)
    const long long entries = e_total * (interp ? 3 : 1);
    // (four launches and a scratch allocation: below ~4 M values the scalar scatter's single launch wins -- 21 vs 43 us at C=128, E=16384)
    if (force != 2 && ((long long)b * c * e_total < (1 << 22) || entries < (long long)len)) return kStagedNotApplicable;
    const int cp = (c + 3) / 4 * 4;
    const long long tiles = (e_total + kTrTile - 1) / kTrTile;
    const int chunks = (c + kTrChunk - 1) / kTrChunk;
    if (!grid_ok(tiles, chunks, b) || !grid_ok((len + 31) / 32, (c + 31) / 32, b)) return kStagedNotApplicable;
    float *acc = nullptr;
    const size_t bytes = (size_t)b * len * cp * sizeof(float);
    cudaError_t e = scratch_alloc(reinterpret_cast<void **>(&acc), bytes, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(acc, 0, bytes, st);
    if (e == cudaSuccess) {
        if (interp)
            cm_scatter_to_pm_kernel<true><<<dim3((unsigned)tiles, chunks, b), kTrThreads, 0, st>>>(c, cp, e_total, len, grad_out, idx, weight, acc);
        else
            cm_scatter_to_pm_kernel<false><<<dim3((unsigned)tiles, chunks, b), kTrThreads, 0, st>>>(c, cp, e_total, len, grad_out, idx, weight, acc);
        pm_add_to_cm_kernel<<<dim3((len + 31) / 32, (c + 31) / 32, b), 256, 0, st>>>(c, cp, len, acc, grad_points);
        e = cudaGetLastError();
    }
    const cudaError_t f = cudaFreeAsync(acc, st);
    return e != cudaSuccess ? (int)e : f != cudaSuccess ? (int)f : EPNET_OK;
}

}  // namespace epnet
