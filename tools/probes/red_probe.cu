// Scattered global reductions on B200: requests per second for red.global.add.f32 (scalar), .v2.f32 and .v4.f32 with one random,
// naturally aligned address per lane over a 32 MB table (L2 resident).  Decides whether a point-major scratch + vector reductions can
// beat the scalar scatter of group_points_grad (csrc/group.cu: row_scatter_add_kernel, one RED per (channel, entry)).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/_bin/red_probe tools/probes/red_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int V>
__global__ void red_kernel(float *table, const uint32_t *slots, long long n_req, uint32_t mask)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_req) return;
    const uint32_t s = slots[i] & mask;  // slot of V floats
    float *p = table + (size_t)s * V;
    if (V == 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
    if (V == 2) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(2.0f) : "memory");
    if (V == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
}

__global__ void fill(uint32_t *slots, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t x = (uint32_t)i * 2654435761u;
        x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        slots[i] = x;
    }
}

template <int V>
static void run(float *table, uint32_t *slots, long long n_req, size_t table_floats)
{
    const uint32_t mask = (uint32_t)(table_floats / V - 1);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int threads = 256;
    const unsigned grid = (unsigned)((n_req + threads - 1) / threads);
    red_kernel<V><<<grid, threads>>>(table, slots, n_req, mask);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) red_kernel<V><<<grid, threads>>>(table, slots, n_req, mask);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    printf("red.add.v%d.f32  %lld requests  %.1f us  %.1f G requests/s  %.1f G floats/s  (%s)\n", V, n_req, ms * 1e3, n_req / ms / 1e6,
           n_req * (double)V / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t table_floats = 8u << 20;  // 32 MB
    const long long n_req = 64ll << 20;
    float *table; uint32_t *slots;
    cudaMalloc(&table, table_floats * 4);
    cudaMemset(table, 0, table_floats * 4);
    cudaMalloc(&slots, n_req * 4);
    fill<<<(unsigned)((n_req + 255) / 256), 256>>>(slots, n_req);
    run<1>(table, slots, n_req, table_floats);
    run<2>(table, slots, n_req / 2, table_floats);
    run<4>(table, slots, n_req / 4, table_floats);
    run<4>(table, slots, n_req, table_floats);
    return 0;
}
