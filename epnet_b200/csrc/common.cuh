// Shared device helpers for the EPNet B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/epnet_b200.h"

#define EPNET_API extern "C" __attribute__((visibility("default")))

// Launch epilogue shared by every entry point: report, never exit() (cf. reference sampling_gpu.cu:248-252).
#define EPNET_RETURN_LAUNCH_STATUS()              \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        return e__ == cudaSuccess ? EPNET_OK : (int)e__; \
    } while (0)

namespace epnet {

constexpr int kSmCount = 148;  // B200

// staged_rows.cu: channel-major gathers served from shared memory; kStagedNotApplicable = run the plain kernel instead
constexpr int kStagedNotApplicable = -2;
int launch_staged_rows(bool interp, int b, int c, int len, long long e_total, const float *src, const int *idx, const float *weight,
                       float *out, cudaStream_t st);
// does a (len)-float row fit the staged kernels' shared memory whole?
bool staged_row_fits(int len);
// transposed_rows.cu: the same channel-major ops through a point-major scratch copy (any row length); kStagedNotApplicable likewise
int launch_transposed_gather(bool interp, bool staged_fits, int b, int c, int len, long long e_total, const float *src, const int *idx,
                             const float *weight, float *out, cudaStream_t st);
int launch_transposed_scatter(bool interp, int b, int c, int len, long long e_total, const float *grad_out, const int *idx,
                              const float *weight, float *grad_points, cudaStream_t st);

// gemm_tf32x3.cu: device address (current device) of the FP16-split range-guard flag every activation-writing epilogue raises
unsigned int *gemm_overflow_flag();

// Squared distance with the exact rounding sequence of the reference kernels as compiled by nvcc -O2
// (PTX: sub,sub,mul,fma,sub,fma): d = fma(dz,dz, fma(dx,dx, dy*dy)), each difference taken as (a - b).
// Written with intrinsics so that no compiler flag can change the contraction.
__device__ __forceinline__ float sqdist_ref(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = __fsub_rn(ax, bx);
    const float dy = __fsub_rn(ay, by);
    const float dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v)
{
    uint32_t r;
    asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v)
{
    uint32_t r;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t r;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D bulk (TMA) copy: global -> shared, completion counted in bytes -----------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes (or the hint expires)
// instead of re-issuing the poll.  Without the hint the bare try_wait loop of the single-lane MMA / loader warps and of the
// producer warps accounted for ~40 % of all issued instructions of the narrow-tile GEMM (ncu source page, r02k), stealing issue slots
// from the epilogue warps on the same schedulers.  (A software poll COUNTER, to trap on protocol errors, measured ~50 ns slower per
// wait on the MMA-issue critical path in round 1 and is still not used.)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace epnet
