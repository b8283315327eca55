// Furthest point sampling for B200.  Replaces furthest_point_sampling_kernel
// (/root/reference/pointnet2_lib/pointnet2/src/sampling_gpu.cu:93-209) bit-exactly.
//
// What has to be reproduced: iteration j picks arg-max over k of temp[k] = min(temp[k], d(k, last)).
// The reference resolves equal maxima through (a) a strided per-thread scan that keeps the first
// maximum and (b) a shared-memory tree that keeps the lower slot unless the upper is strictly larger.
// Net effect: among tied k the winner minimises  (bitreverse_L(k mod BS), k div BS),
// BS = 2^L = min(1024, 2^floor(log2 N))  (cuda_utils.h:10-14).  We compute that order directly: a
// candidate is (distance bits, tie word); distances are >= 0 so their IEEE bits order as unsigned
// ints; the winner is max distance, then min tie word -- two warp `redux.sync` per level instead of
// a 10-barrier shared-memory tree.
//
// The reference does N distance updates per iteration (M*N in total, 67 M for 16384->4096) on one
// SM.  Here the work per iteration is made proportional to the points that can actually change:
//   * points are counting-sorted once, in shared memory, into Morton order of a near-cubic cell grid;
//     128 consecutive sorted points form a bucket with an exact bounding box;
//   * each bucket caches its current (max running distance, tie word);
//   * an iteration first tests the new sample against every bucket's box: if the (slightly deflated)
//     box distance is >= the bucket's max running distance, no min() in that bucket can change, its
//     cached winner stays valid and the bucket is skipped -- which is what happens to almost all
//     buckets once a few dozen samples exist.  Skipping never changes a result: every distance that
//     IS computed uses the reference's exact FMA sequence, every skipped one is provably >= temp.
// Sorted coordinates live in shared memory as three planes (192 KB at N = 16384), running distances and
// tie words in registers (lane l of warp w owns sorted positions (((b*NW+w)*kSlots+i)*32+l), bucket boxes and
// cached winners in the registers of lanes 0..7.  One __syncthreads per iteration (none when a single
// warp holds the whole cloud, N <= 256).  16384 < N <= 131072 spreads a scene over a thread-block cluster (DSMEM exchange of
// the per-CTA winners); beyond that coordinates stream from L2 (fps_streaming_kernel).
#include "common.cuh"
#include <cstdlib>

namespace epnet {

constexpr int kFpsMaxResident = 16384;  // 3 planes * 4 B * 16384 = 192 KB of the 227 KB shared memory
constexpr float kBoxDeflate = 0.999996f;  // > (1 - 2^-18): covers the few-ulp rounding of both distance evaluations

struct FpsCand {
    uint32_t dist_bits;
    uint32_t tie;
};
struct FpsClusterCand {  // 32 bytes: one v4 + one scalar remote store
    uint32_t bits, key;
    float x, y;
    float z;
    uint32_t pad[3];
};

// ---- tie words -------------------------------------------------------------------------------
// compact key ck(k) = (bitreverse_L(k mod BS) << qbits) | (k >> L)   (order == the reference's tie order)
// tie word = (ck << 16) | sorted position            (N <= 16384: ck < 2^14, position < 2^14)
__device__ __forceinline__ uint32_t fps_compact_key(uint32_t k, int L, int qbits)
{
    const uint32_t rev = L ? (__brev(k & ((1u << L) - 1u)) >> (32 - L)) : 0u;
    return (rev << qbits) | (k >> L);
}
__device__ __forceinline__ uint32_t fps_key_to_index(uint32_t ck, int L, int qbits)
{
    const uint32_t hi = ck & ((1u << qbits) - 1u);
    const uint32_t rev = ck >> qbits;
    const uint32_t low = L ? (__brev(rev) >> (32 - L)) : 0u;
    return (hi << L) | low;
}

__device__ __forceinline__ float warp_min_f(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// (max distance bits, then min tie word) over the warp
__device__ __forceinline__ void warp_argmax(uint32_t &bits, uint32_t &tie)
{
    const uint32_t mx = warp_max_u32(bits);
    tie = warp_min_u32(bits == mx ? tie : 0xffffffffu);
    bits = mx;
}

// NW warps, kBuckets buckets per warp (lane b keeps bucket b's box), kSlots points per lane per bucket (bucket = 32*kSlots sorted points).
// Buckets are dealt to warps round-robin (bucket g -> warp g % NW): buckets that are neighbours in Morton
// order, i.e. the ones a new sample touches together, are updated by different warps in parallel.
// kCluster: one scene is spread over a thread-block cluster of `csize` CTAs (N up to 8 x 16384): CTA r owns the original
// indices [r*slice, (r+1)*slice), sorts and buckets them on its own, and after the CTA-level arg-max the csize candidates
// (distance bits, tie word, coordinates) are exchanged through distributed shared memory -- every CTA stores its candidate
// into every peer's slot array (st.shared::cluster), one barrier.cluster per iteration, every warp re-reduces locally.
template <int NW, int kBuckets, int kSlots, bool kCluster>
__global__ void __launch_bounds__(NW * 32, 1)
fps_bucket_kernel(int n_total, int m, int L, int qbits, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx,
                  float *__restrict__ new_xyz, const float *__restrict__ aux_in, float *__restrict__ aux_out, int aux_dim, int csize,
                  int slice, const int *__restrict__ skip)
{
    constexpr int T = NW * 32;
    constexpr int CAP = T * kBuckets * kSlots;  // sorted positions this CTA can hold
    constexpr int PPT = CAP / T;                // points per thread during the sort
    constexpr int NCELL = CAP;                  // one cell per sorted position on average; the histogram aliases one plane
    constexpr int CB = CAP == 256 ? 8 : CAP == 1024 ? 10 : CAP == 4096 ? 12 : 14;
    constexpr int PB = kCluster ? 14 : 16;  // position bits of a tie word (cluster: ck needs up to 17 bits)
    constexpr uint32_t PMASK = (1u << PB) - 1u;
    static_assert((1 << CB) == NCELL, "cell bits");
    static_assert(kBuckets <= 32, "one lane per bucket");
#define FPS_POS(b, i) ((((b) * NW + warp) * kSlots + (i)) * 32 + lane)

    extern __shared__ __align__(16) float fps_smem[];
    float *xs = fps_smem, *ys = fps_smem + CAP, *zs = fps_smem + 2 * CAP;
    // aliases used before the planes are filled
    uint32_t *hist = reinterpret_cast<uint32_t *>(fps_smem);             // NCELL counters
    float *carry_t = fps_smem;                                            // CAP running distances by sorted position
    uint32_t *carry_k = reinterpret_cast<uint32_t *>(fps_smem + CAP);     // CAP original indices by sorted position

    __shared__ float red[32][6];
    __shared__ float box[6];          // scene bounding box: min xyz, max xyz
    __shared__ float inv_cell[3];
    __shared__ int axis_bits[3];
    __shared__ int bit_axis[16];      // axis providing Morton bit s (MSB first)
    __shared__ uint32_t warp_tot[32];
    __shared__ FpsCand slots[2][32];
    __shared__ __align__(16) FpsClusterCand cslots[2][8];

    const int scene = kCluster ? blockIdx.x / csize : blockIdx.x;
    if (skip && __ldg(skip + scene)) return;  // answered by fps_prefix_fill_kernel (every CTA of a cluster takes the same branch)
    const int rank = kCluster ? blockIdx.x % csize : 0;   // == %cluster_ctarank for a 1-D cluster
    const int first = rank * slice;                        // first original index owned by this CTA
    const int n = kCluster ? min(slice, n_total - first) : n_total;  // points owned by this CTA
    const float *xyz_scene = xyz + (size_t)scene * n_total * 3;
    xyz = xyz_scene + (size_t)first * 3;
    float *temp_scene = temp + (size_t)scene * n_total;
    temp = temp_scene + first;
    idx += (size_t)scene * m;
    if (new_xyz) new_xyz += (size_t)scene * m * 3;
    if (aux_out) {
        aux_in += (size_t)scene * n_total * aux_dim;
        aux_out += (size_t)scene * m * aux_dim;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inf = __int_as_float(0x7f800000);

    if (tid == 0 && rank == 0) {
        idx[0] = 0;
        if (new_xyz) {
            new_xyz[0] = __ldg(xyz_scene);
            new_xyz[1] = __ldg(xyz_scene + 1);
            new_xyz[2] = __ldg(xyz_scene + 2);
        }
    }
    if (m == 1) {
        if (aux_out && rank == 0 && tid < aux_dim) aux_out[tid] = __ldg(aux_in + tid);
        return;
    }

    // ---------------- 1. scene bounding box ----------------
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + i * T;
        if (k < n) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __ldg(xyz + 3 * k + a);
                lo[a] = fminf(lo[a], v);
                hi[a] = fmaxf(hi[a], v);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min_f(lo[a]);
        hi[a] = warp_max_f(hi[a]);
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            red[warp][a] = lo[a];
            red[warp][3 + a] = hi[a];
        }
    }
    __syncthreads();
    if (tid == 0) {
        float ext[3];
        for (int a = 0; a < 3; ++a) {
            float l = inf, h = -inf;
            for (int w = 0; w < NW; ++w) {
                l = fminf(l, red[w][a]);
                h = fmaxf(h, red[w][3 + a]);
            }
            box[a] = l;
            box[3 + a] = h;
            ext[a] = (h > l && h - l < inf) ? h - l : 0.f;  // empty / degenerate / non-finite axes get no bits
            axis_bits[a] = 0;
        }
        // hand out CB Morton bits, MSB first, always to the axis whose cells are currently longest
        float cell[3] = {ext[0], ext[1], ext[2]};
        for (int s = 0; s < CB; ++s) {
            int best = 0;
            if (cell[2] > cell[best]) best = 2;   // prefer x, then z, then y on ties
            if (cell[1] > cell[best]) best = 1;
            bit_axis[s] = best;
            axis_bits[best] += 1;
            cell[best] *= 0.5f;
        }
        for (int a = 0; a < 3; ++a) inv_cell[a] = ext[a] > 0.f ? (float)(1 << axis_bits[a]) / ext[a] : 0.f;
    }
    for (int c = tid; c < NCELL; c += T) hist[c] = 0u;
    __syncthreads();

    // ---------------- 2. cell of every point, rank inside the cell ----------------
    // cell | rank << CB, later the sorted position.  Indexed dynamically on purpose (not unrolled): it lives in
    // local memory (L1), is touched four times per point in this one-off prologue, and keeps the code small.
    uint32_t code[PPT];
    {
        const int nb0 = axis_bits[0], nb1 = axis_bits[1], nb2 = axis_bits[2];
        const float bx = box[0], by = box[1], bz = box[2];
        const float ix = inv_cell[0], iy = inv_cell[1], iz = inv_cell[2];
#pragma unroll 1
        for (int i = 0; i < PPT; ++i) {
            const int k = tid + i * T;
            code[i] = 0xffffffffu;
            if (k < n) {
                const float x = __ldg(xyz + 3 * k), y = __ldg(xyz + 3 * k + 1), z = __ldg(xyz + 3 * k + 2);
                int q[3];
                q[0] = min(max((int)((x - bx) * ix), 0), (1 << nb0) - 1);   // NaN -> 0
                q[1] = min(max((int)((y - by) * iy), 0), (1 << nb1) - 1);
                q[2] = min(max((int)((z - bz) * iz), 0), (1 << nb2) - 1);
                int left[3] = {nb0, nb1, nb2};
                uint32_t cell = 0;
                for (int s = 0; s < CB; ++s) {
                    const int a = bit_axis[s];
                    const int l = (a == 0 ? left[0] : (a == 1 ? left[1] : left[2])) - 1;
                    const int qa = a == 0 ? q[0] : (a == 1 ? q[1] : q[2]);
                    cell = (cell << 1) | ((qa >> l) & 1);
                    if (a == 0) left[0] = l; else if (a == 1) left[1] = l; else left[2] = l;
                }
                const uint32_t rank = atomicAdd(&hist[cell], 1u);
                code[i] = cell | (rank << CB);
            }
        }
    }
    __syncthreads();

    // ---------------- 3. exclusive scan of the histogram (PPT consecutive cells per thread) ----------------
    {
        uint32_t sum = 0;
#pragma unroll 4
        for (int i = 0; i < PPT; ++i) sum += hist[tid * PPT + ((i + lane) & (PPT - 1))];  // rotated: conflict-free
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll 4
        for (int i = 0; i < PPT; ++i) {
            const uint32_t c = hist[tid * PPT + i];
            hist[tid * PPT + i] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < PPT; ++i)
        if (code[i] != 0xffffffffu) code[i] = hist[code[i] & (NCELL - 1)] + (code[i] >> CB);  // sorted position
    __syncthreads();

    // ---------------- 4. carry running distance + original index to the owner of each sorted position ----------------
#pragma unroll 1
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + i * T;
        if (k < n) {
            carry_t[code[i]] = temp[k];
            carry_k[code[i]] = (uint32_t)(first + k);  // ORIGINAL (scene-wide) index: the tie rule is defined on it
        }
    }
    __syncthreads();
    float t[kBuckets][kSlots];
    uint32_t tw[kBuckets][kSlots];
#pragma unroll
    for (int b = 0; b < kBuckets; ++b)
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
            const int pos = FPS_POS(b, i);
            if (pos < n) {
                t[b][i] = carry_t[pos];
                tw[b][i] = (fps_compact_key(carry_k[pos], L, qbits) << PB) | (uint32_t)pos;
            } else {  // empty slot: never changes, loses every tie
                t[b][i] = 0.f;
                tw[b][i] = 0xffffffffu;
            }
        }
    __syncthreads();

    // ---------------- 5. scatter coordinates into the sorted planes ----------------
#pragma unroll 1
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + i * T;
        if (k < n) {
            const uint32_t pos = code[i];
            xs[pos] = __ldg(xyz + 3 * k);
            ys[pos] = __ldg(xyz + 3 * k + 1);
            zs[pos] = __ldg(xyz + 3 * k + 2);
        }
    }
    for (int pos = n + tid; pos < CAP; pos += T) xs[pos] = ys[pos] = zs[pos] = 0.f;
    __syncthreads();

    // ---------------- 6. bucket boxes and initial cached winners (lane b keeps bucket b) ----------------
    float blo_x = inf, blo_y = inf, blo_z = inf, bhi_x = -inf, bhi_y = -inf, bhi_z = -inf;
    uint32_t bmax = 0u, bkey = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < kBuckets; ++b) {
        float l0 = inf, l1 = inf, l2 = inf, h0 = -inf, h1 = -inf, h2 = -inf;
        float bt = t[b][0];
        uint32_t bk = tw[b][0];
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
            const int pos = FPS_POS(b, i);
            if (pos < n) {
                const float x = xs[pos], y = ys[pos], z = zs[pos];
                l0 = fminf(l0, x); h0 = fmaxf(h0, x);
                l1 = fminf(l1, y); h1 = fmaxf(h1, y);
                l2 = fminf(l2, z); h2 = fmaxf(h2, z);
            }
            if (i > 0) {
                const bool take = (t[b][i] > bt) || (t[b][i] == bt && tw[b][i] < bk);
                bt = take ? t[b][i] : bt;
                bk = take ? tw[b][i] : bk;
            }
        }
        l0 = warp_min_f(l0); l1 = warp_min_f(l1); l2 = warp_min_f(l2);
        h0 = warp_max_f(h0); h1 = warp_max_f(h1); h2 = warp_max_f(h2);
        uint32_t bits = __float_as_uint(bt);
        warp_argmax(bits, bk);
        if (lane == b) {
            blo_x = l0; blo_y = l1; blo_z = l2;
            bhi_x = h0; bhi_y = h1; bhi_z = h2;
            bmax = bits; bkey = bk;
        }
    }

    // ---------------- 7. the sampling loop ----------------
    float cx = __ldg(xyz_scene), cy = __ldg(xyz_scene + 1), cz = __ldg(xyz_scene + 2);  // sample 0 is point 0 of the scene
    uint32_t warp_bits = lane < kBuckets ? bmax : 0u, warp_key = lane < kBuckets ? bkey : 0xffffffffu;
    warp_argmax(warp_bits, warp_key);
    const uint32_t slot_wr = smem_u32(&slots[0][warp]), slot_rd = smem_u32(&slots[0][lane & 31]);
    for (int j = 1; j < m; ++j) {
        // which buckets can change?  box distance (deflated) vs the bucket's largest running distance
        const float ex = fmaxf(fmaxf(blo_x - cx, cx - bhi_x), 0.f);
        const float ey = fmaxf(fmaxf(blo_y - cy, cy - bhi_y), 0.f);
        const float ez = fmaxf(fmaxf(blo_z - cz, cz - bhi_z), 0.f);
        const float lb = (ex * ex + ey * ey + ez * ez) * kBoxDeflate;
        const bool active = lane < kBuckets && lb < __uint_as_float(bmax);
        const uint32_t mask = __ballot_sync(0xffffffffu, active);
        if (mask) {  // most warps, most iterations: nothing to update -> straight to the exchange
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) {
            if (mask & (1u << b)) {  // warp-uniform
                // lane-local best of the bucket as a tournament (depth log2(kSlots)), ties to the smaller tie word
                float bt[kSlots];
                uint32_t bk[kSlots];
#pragma unroll
                for (int i = 0; i < kSlots; ++i) {
                    const int pos = FPS_POS(b, i);
                    bt[i] = fminf(sqdist_ref(xs[pos], ys[pos], zs[pos], cx, cy, cz), t[b][i]);
                    t[b][i] = bt[i];
                    bk[i] = tw[b][i];
                }
#pragma unroll
                for (int w = 1; w < kSlots; w <<= 1) {
#pragma unroll
                    for (int i = 0; i + w < kSlots; i += 2 * w) {
                        const bool take = (bt[i + w] > bt[i]) || (bt[i + w] == bt[i] && bk[i + w] < bk[i]);
                        bt[i] = take ? bt[i + w] : bt[i];
                        bk[i] = take ? bk[i + w] : bk[i];
                    }
                }
                uint32_t bits = __float_as_uint(bt[0]), key = bk[0];
                warp_argmax(bits, key);
                if (lane == b) {
                    bmax = bits;
                    bkey = key;
                }
            }
        }
        // winner of the warp's buckets (unchanged if none of them was touched), then of the CTA
            warp_bits = lane < kBuckets ? bmax : 0u;
            warp_key = lane < kBuckets ? bkey : 0xffffffffu;
            warp_argmax(warp_bits, warp_key);
        }
        uint32_t wbits = warp_bits, wkey = warp_key;
        if (NW > 1) {
            // shared-space addresses computed once outside the loop (no generic->shared conversion per iteration)
            const uint32_t par_off = (uint32_t)(j & 1) * 32u * 8u;
            if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot_wr + par_off), "r"(wbits), "r"(wkey) : "memory");
            __syncthreads();
            wbits = 0u;
            wkey = 0xffffffffu;
            if (lane < NW) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wbits), "=r"(wkey) : "r"(slot_rd + par_off) : "memory");
            warp_argmax(wbits, wkey);
        }
        const uint32_t pos = wkey & PMASK;
        cx = xs[pos];
        cy = ys[pos];
        cz = zs[pos];
        if (kCluster) {
            // publish this CTA's candidate in every CTA of the cluster (lane r of warp 0 writes to rank r), one cluster barrier,
            // then every warp picks the winner among the csize candidates it finds in its OWN shared memory
            const int par = j & 1;
            if (warp == 0 && lane < csize) {
                const uint32_t local = smem_u32(&cslots[par][rank]);
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(lane));
                asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"(wbits), "r"(wkey), "r"(__float_as_uint(cx)),
                             "r"(__float_as_uint(cy))
                             : "memory");
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote + 16), "r"(__float_as_uint(cz)) : "memory");
            }
            asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
            uint32_t gb = lane < csize ? cslots[par][lane].bits : 0u;
            uint32_t gk = lane < csize ? cslots[par][lane].key : 0xffffffffu;
            const uint32_t mine = gk;
            warp_argmax(gb, gk);
            const int src = __ffs(__ballot_sync(0xffffffffu, mine == gk && lane < csize)) - 1;  // tie words are unique
            cx = cslots[par][src].x;
            cy = cslots[par][src].y;
            cz = cslots[par][src].z;
            wkey = gk;
        }
        if (tid == 0 && rank == 0) {
            idx[j] = (int)fps_key_to_index(wkey >> PB, L, qbits);
            if (new_xyz) {
                new_xyz[3 * j] = cx;
                new_xyz[3 * j + 1] = cy;
                new_xyz[3 * j + 2] = cz;
            }
        }
    }

    // per-sample payload (LI-Fusion pixel coordinates) follows the samples: aux_out[j] = aux_in[idx[j]]
    if (aux_out && rank == 0) {
        __syncthreads();  // idx[] written by thread 0 is visible to the CTA
        for (int e = tid; e < m * aux_dim; e += T) {
            const int j = e / aux_dim, a = e - j * aux_dim;
            aux_out[e] = __ldg(aux_in + (size_t)idx[j] * aux_dim + a);
        }
    }

    // temp is an in/out buffer in the reference: leave the final running distances behind
#pragma unroll
    for (int b = 0; b < kBuckets; ++b)
#pragma unroll
        for (int i = 0; i < kSlots; ++i)
            if (tw[b][i] != 0xffffffffu) temp_scene[fps_key_to_index(tw[b][i] >> PB, L, qbits)] = t[b][i];
}

// ---- N too large for one SM's shared memory: coordinates and running distances stream from L2 ----
__device__ __forceinline__ uint32_t fps_tie_key(uint32_t k, uint32_t tid, int L) { return __brev(tid) | (k >> L); }
__device__ __forceinline__ uint32_t fps_tie_decode(uint32_t tie, int L)
{
    const uint32_t low = L ? (0xffffffffu >> L) : 0xffffffffu;
    return ((tie & low) << L) | __brev(tie & ~low);
}

__global__ void __launch_bounds__(1024, 1)
fps_streaming_kernel(int n, int m, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx,
                     float *__restrict__ new_xyz, const float *__restrict__ aux_in, float *__restrict__ aux_out, int aux_dim,
                     const int *__restrict__ skip)
{
    __shared__ FpsCand slots[2][32];
    const int scene = blockIdx.x;
    if (skip && __ldg(skip + scene)) return;
    xyz += (size_t)scene * n * 3;
    temp += (size_t)scene * n;
    idx += (size_t)scene * m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int L = 10;
    if (new_xyz) new_xyz += (size_t)scene * m * 3;
    if (aux_out) {
        aux_in += (size_t)scene * n * aux_dim;
        aux_out += (size_t)scene * m * aux_dim;
    }
    if (tid == 0) idx[0] = 0;

    uint32_t last = 0;
    for (int j = 1; j < m; ++j) {
        const float lx = __ldg(xyz + 3 * last), ly = __ldg(xyz + 3 * last + 1), lz = __ldg(xyz + 3 * last + 2);
        float best = -1.f;
        uint32_t besti = 0;
        for (int k = tid; k < n; k += 1024) {  // the reference's ownership: "first maximum per thread" carries over
            const float d = sqdist_ref(__ldg(xyz + 3 * k), __ldg(xyz + 3 * k + 1), __ldg(xyz + 3 * k + 2), lx, ly, lz);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > best) {
                best = d2;
                besti = k;
            }
        }
        uint32_t bits = __float_as_uint(best), tie = fps_tie_key(besti, tid, L);
        warp_argmax(bits, tie);
        FpsCand *s = slots[j & 1];
        if (lane == 0) {
            s[warp].dist_bits = bits;
            s[warp].tie = tie;
        }
        __syncthreads();
        bits = s[lane].dist_bits;
        tie = s[lane].tie;
        warp_argmax(bits, tie);
        last = fps_tie_decode(tie, L);
        if (tid == 0) idx[j] = (int)last;
    }
    __syncthreads();
    for (int e = tid; e < m * 3 && new_xyz; e += 1024) new_xyz[e] = __ldg(xyz + 3 * idx[e / 3] + e % 3);
    for (int e = tid; e < m * aux_dim && aux_out; e += 1024) aux_out[e] = __ldg(aux_in + (size_t)idx[e / aux_dim] * aux_dim + e % aux_dim);
}

// ---- small clouds (N <= 512: RCNN-stage RoIs, the last backbone level): one warp per scene, several scenes per CTA ----
// Lane l keeps points l, l+32, ... (coordinates, running distance) in registers; coordinates also sit in shared memory so the
// winner's can be broadcast-read.  No sort, no buckets, no barrier: an iteration is PPL distance updates per lane and one warp
// arg-max (max distance bits, then min compact key = the reference's tie order).
template <int PPL>
__global__ void __launch_bounds__(128)
fps_warp_kernel(int scenes, int n, int m, int L, int qbits, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx,
                float *__restrict__ new_xyz, const float *__restrict__ aux_in, float *__restrict__ aux_out, int aux_dim,
                const int *__restrict__ skip)
{
    __shared__ float s_xyz[4][PPL * 32 * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int scene = blockIdx.x * 4 + warp;
    if (scene >= scenes) return;
    if (skip && __ldg(skip + scene)) return;  // no block barrier in this kernel: a warp may leave on its own
    xyz += (size_t)scene * n * 3;
    temp += (size_t)scene * n;
    idx += (size_t)scene * m;
    float *sx = s_xyz[warp];
    for (int e = lane; e < n * 3; e += 32) sx[e] = __ldg(xyz + e);
    __syncwarp();

    float px[PPL], py[PPL], pz[PPL], t[PPL];
    uint32_t ck[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
        const int k = lane + 32 * i;
        const bool ok = k < n;
        px[i] = ok ? sx[3 * k] : 0.f;
        py[i] = ok ? sx[3 * k + 1] : 0.f;
        pz[i] = ok ? sx[3 * k + 2] : 0.f;
        t[i] = ok ? temp[k] : -1.f;  // a slot without a point never wins: every real running distance is >= 0
        ck[i] = ok ? fps_compact_key((uint32_t)k, L, qbits) : 0xffffffffu;
    }
    if (lane == 0) idx[0] = 0;
    uint32_t last = 0;
    for (int j = 1; j < m; ++j) {
        const float lx = sx[3 * last], ly = sx[3 * last + 1], lz = sx[3 * last + 2];
        // candidate = (distance bits, inverted compact key) as one 64-bit integer: larger distance wins, then smaller key; the
        // lane's PPL candidates meet in a tournament tree (depth log2 PPL) instead of a serial compare chain
        unsigned long long cand[PPL];
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            t[i] = fminf(t[i], sqdist_ref(px[i], py[i], pz[i], lx, ly, lz));  // padding slots: fminf(-1, d) stays -1
            cand[i] = t[i] < 0.f ? 0ull : ((unsigned long long)__float_as_uint(t[i]) << 32) | (unsigned long long)(~ck[i]);
        }
#pragma unroll
        for (int w = 1; w < PPL; w <<= 1)
#pragma unroll
            for (int i = 0; i + w < PPL; i += 2 * w) cand[i] = cand[i] > cand[i + w] ? cand[i] : cand[i + w];
        uint32_t bits = (uint32_t)(cand[0] >> 32), tie = ~(uint32_t)cand[0];
        warp_argmax(bits, tie);
        last = fps_key_to_index(tie, L, qbits);
        if (lane == 0) idx[j] = (int)last;
    }
#pragma unroll
    for (int i = 0; i < PPL; ++i)  // temp is an in/out buffer in the reference: leave the final running distances behind
        if (lane + 32 * i < n) temp[lane + 32 * i] = t[i];
    if (new_xyz || aux_out) {
        __syncwarp();  // idx[] written by lane 0 is visible to the warp
        for (int e = lane; e < m * 3 && new_xyz; e += 32) new_xyz[(size_t)scene * m * 3 + e] = sx[3 * idx[e / 3] + e % 3];
        for (int e = lane; e < m * aux_dim && aux_out; e += 32)
            aux_out[(size_t)scene * m * aux_dim + e] = __ldg(aux_in + ((size_t)scene * n + idx[e / aux_dim]) * aux_dim + e % aux_dim);
    }
}

template <int PPL>
static int launch_warp(int b, int n, int m, int L, int qbits, const float *xyz, float *temp, int *idx, float *new_xyz,
                       const float *aux_in, float *aux_out, int aux_dim, const int *skip, cudaStream_t st)
{
    fps_warp_kernel<PPL><<<(b + 3) / 4, 128, 0, st>>>(b, n, m, L, qbits, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, skip);
    EPNET_RETURN_LAUNCH_STATUS();
}

template <int NW, int kBuckets, int kSlots>
static int launch_bucket(int b, int n, int m, int L, int qbits, const float *xyz, float *temp, int *idx, float *new_xyz,
                         const float *aux_in, float *aux_out, int aux_dim, const int *skip, cudaStream_t st)
{
    const size_t smem = (size_t)3 * NW * 32 * kBuckets * kSlots * sizeof(float);
    auto kernel = fps_bucket_kernel<NW, kBuckets, kSlots, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kernel<<<b, NW * 32, smem, st>>>(n, m, L, qbits, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, 1, n, skip);
    EPNET_RETURN_LAUNCH_STATUS();
}

// 16384 < N <= 8 * 16384: a thread-block cluster of ceil(N / 16384) CTAs per scene, each with the full resident layout
static int launch_cluster(int b, int n, int m, int L, int qbits, const float *xyz, float *temp, int *idx, float *new_xyz,
                          const float *aux_in, float *aux_out, int aux_dim, const int *skip, cudaStream_t st, int force_csize = 0)
{
    constexpr int NW = 16, KB = 8, KS = 4;
    const int csize = force_csize > 0 ? force_csize : (n + kFpsMaxResident - 1) / kFpsMaxResident;
    const int slice = (n + csize - 1) / csize;
    const size_t smem = (size_t)3 * NW * 32 * KB * KS * sizeof(float);
    auto kernel = fps_bucket_kernel<NW, KB, KS, true>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * csize));
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, n, m, L, qbits, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, csize, slice, skip);
    if (e != cudaSuccess) return (int)e;
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet

static int fps_dispatch(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, const float *aux_in,
                        float *aux_out, int aux_dim, const int *skip, void *stream)
{
    using namespace epnet;
    if (b < 0 || n <= 0 || m < 0 || !xyz || !temp || !idx) return EPNET_ERR_BAD_ARG;
    if ((aux_out != nullptr) != (aux_in != nullptr) || (aux_out && (aux_dim < 1 || aux_dim > 4))) return EPNET_ERR_BAD_ARG;
    if (b == 0 || m == 0) return EPNET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int L = 0;
    while ((2 << L) <= n && L < 10) ++L;  // BS = 2^L = min(1024, 2^floor(log2 n))
    if (n <= kFpsMaxResident) {
        const int q = (n + (1 << L) - 1) >> L;  // points per reference thread
        int qbits = 0;
        while ((1 << qbits) < q) ++qbits;
#define EPNET_FPS_ARGS b, n, m, L, qbits, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, skip, st
        if (n <= 128) return launch_warp<4>(EPNET_FPS_ARGS);            // one warp per scene, points in registers, no buckets
        if (n <= 256) return launch_warp<8>(EPNET_FPS_ARGS);
        if (n <= 512) return launch_warp<16>(EPNET_FPS_ARGS);
        if (n <= 1024) return launch_bucket<4, 8, 1>(EPNET_FPS_ARGS);   //  32 buckets of  32
        if (n <= 4096) return launch_bucket<16, 8, 1>(EPNET_FPS_ARGS);  // 128 buckets of  32
        {   // experiment switch: spread one scene of <= 16384 points over a cluster of 2 / 4 CTAs (DSMEM exchange per iteration)
            static const int forced = [] { const char *e = getenv("EPNET_FPS_CLUSTER"); return e ? atoi(e) : 0; }();
            if (forced > 1 && forced <= 8 && n >= 8192)
                return launch_cluster(b, n, m, L, qbits, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, skip, st, forced);
        }
        return launch_bucket<16, 8, 4>(EPNET_FPS_ARGS);                 // 128 buckets of 128
        // (measured alternatives: 8 warps x 16 buckets is 27 % slower at N = 16384 -- touched buckets of one warp update
        //  serially -- and 4 warps x 8 buckets x 128 points is 25 % slower at N = 4096)
#undef EPNET_FPS_ARGS
    }
    if (n <= 8 * kFpsMaxResident) {
        const int q = (n + 1023) >> 10;
        int qbits = 0;
        while ((1 << qbits) < q) ++qbits;
        return launch_cluster(b, n, m, 10, qbits, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, skip, st);
    }
    fps_streaming_kernel<<<b, 1024, 0, st>>>(n, m, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, skip);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream)
{
    return fps_dispatch(b, n, m, xyz, temp, idx, nullptr, nullptr, nullptr, 0, nullptr, stream);
}

EPNET_API int epnet_fps_sample(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, const float *aux_in,
                               float *aux_out, int aux_dim, void *stream)
{
    return fps_dispatch(b, n, m, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, nullptr, stream);
}

// ---- sampling a cloud that is already in furthest-point order ------------------------------------------------------------------------
// The set-abstraction levels sample 16384 -> 4096 -> 1024 -> 256 -> 64: every level after the first samples the previous level's OUTPUT,
// which is in furthest-point order, so its answer is the identity 0..m-1 -- the arg-max of step s over the sub-cloud is the point the
// first level took at step s over the whole cloud -- unless two points of the sub-cloud tie for a maximum (then the tie rules of the two
// levels, which depend on array positions, may differ).  epnet_fps_prefix_check decides that exactly, in O(n*m) parallel work instead
// of m dependent arg-max steps: with T_j(s) = min(1e10, d(p_j,p_0), ..., d(p_j,p_{s-1})) (the running distances furthest-point sampling
// holds at step s if it has returned 0..s-1 so far, evaluated with the same fma chain) it sets flag = 1 iff T_j(s) < T_s(s) for every
// step 1 <= s < m and every j != s: the arg-max of every step is unique and is s, so the identity is what the sampling kernel returns,
// whatever its tie rule.  Because the condition for (n', m') with n' <= n, m' <= m is a subset of the one for (n, m), one check of the
// second level also covers the later ones.  epnet_fps_sample_guarded then answers flagged scenes with the prefix and runs the real
// sampling for the others (lattices, duplicated points, anything not in furthest-point order): bit-exact either way.
namespace epnet {

constexpr int kPrefixThreads = 128;
constexpr int kPrefixMaxM = 2048;  // the first m points and their winning distances sit in shared memory

__global__ void __launch_bounds__(kPrefixThreads)
fps_prefix_winners_kernel(int n, int m, const float *__restrict__ xyz, float *__restrict__ winners, int *__restrict__ flag)
{
    __shared__ float4 pts[kPrefixMaxM];
    const int scene = blockIdx.y;
    xyz += (size_t)scene * n * 3;
    const int s = blockIdx.x * kPrefixThreads + threadIdx.x;
    const int need = min(m, (blockIdx.x + 1) * kPrefixThreads);  // this CTA's steps look at points below their own index
    for (int i = threadIdx.x; i < need; i += kPrefixThreads)
        pts[i] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), 0.f);
    if (blockIdx.x == 0 && threadIdx.x == 0) flag[scene] = 1;  // fps_prefix_check_kernel (next in the stream) clears it
    __syncthreads();
    if (s >= m) return;
    const float4 p = pts[s];
    float t = 1e10f;
#pragma unroll 8
    for (int i = 0; i < s; ++i) {
        const float4 q = pts[i];
        t = fminf(sqdist_ref(p.x, p.y, p.z, q.x, q.y, q.z), t);
    }
    winners[(size_t)scene * m + s] = t;
}

__global__ void __launch_bounds__(kPrefixThreads)
fps_prefix_check_kernel(int n, int m, const float *__restrict__ xyz, const float *__restrict__ winners, int *__restrict__ flag)
{
    __shared__ float4 pts[kPrefixMaxM];  // .w = the winning distance of the step that picks this point
    const int scene = blockIdx.y;
    xyz += (size_t)scene * n * 3;
    winners += (size_t)scene * m;
    for (int i = threadIdx.x; i < m; i += kPrefixThreads)
        pts[i] = make_float4(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2), __ldg(winners + i));
    __syncthreads();
    const int j = blockIdx.x * kPrefixThreads + threadIdx.x;
    if (j >= n) return;
    const float px = __ldg(xyz + 3 * j), py = __ldg(xyz + 3 * j + 1), pz = __ldg(xyz + 3 * j + 2);
    float t = 1e10f;
    bool ok = true;
    float4 prev = pts[0];
#pragma unroll 8
    for (int s = 1; s < m; ++s) {
        const float4 cur = pts[s];
        t = fminf(sqdist_ref(px, py, pz, prev.x, prev.y, prev.z), t);  // T_j(s)
        ok = ok && (j == s || t < cur.w);                               // NaN fails the comparison: no shortcut
        prev = cur;
    }
    if (!ok) flag[scene] = 0;
}

__global__ void __launch_bounds__(256)
fps_prefix_fill_kernel(int n, int m, const float *__restrict__ xyz, int *__restrict__ idx, float *__restrict__ new_xyz,
                       const float *__restrict__ aux_in, float *__restrict__ aux_out, int aux_dim, const int *__restrict__ flag)
{
    const int scene = blockIdx.y;
    if (!__ldg(flag + scene)) return;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= m) return;
    idx[(size_t)scene * m + k] = k;
    if (new_xyz)
        for (int a = 0; a < 3; ++a) new_xyz[((size_t)scene * m + k) * 3 + a] = __ldg(xyz + ((size_t)scene * n + k) * 3 + a);
    if (aux_out)
        for (int a = 0; a < aux_dim; ++a) aux_out[((size_t)scene * m + k) * aux_dim + a] = __ldg(aux_in + ((size_t)scene * n + k) * aux_dim + a);
}

}  // namespace epnet

EPNET_API int epnet_fps_prefix_check(int b, int n, int m, const float *xyz, float *winners, int *flag, void *stream)
{
    using namespace epnet;
    if (b < 0 || n <= 0 || m <= 0 || m > n || !xyz || !winners || !flag) return EPNET_ERR_BAD_ARG;
    if (b == 0) return EPNET_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (m > kPrefixMaxM || b > 65535) {  // no shortcut: every scene runs the real sampling
        cudaError_t e = cudaMemsetAsync(flag, 0, (size_t)b * sizeof(int), st);
        return e == cudaSuccess ? EPNET_OK : (int)e;
    }
    fps_prefix_winners_kernel<<<dim3((m + kPrefixThreads - 1) / kPrefixThreads, b), kPrefixThreads, 0, st>>>(n, m, xyz, winners, flag);
    fps_prefix_check_kernel<<<dim3((n + kPrefixThreads - 1) / kPrefixThreads, b), kPrefixThreads, 0, st>>>(n, m, xyz, winners, flag);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_fps_sample_guarded(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, const float *aux_in,
                                       float *aux_out, int aux_dim, const int *identity, void *stream)
{
    using namespace epnet;
    if (!identity || m > n) return EPNET_ERR_BAD_ARG;
    const int rc = fps_dispatch(b, n, m, xyz, temp, idx, new_xyz, aux_in, aux_out, aux_dim, identity, stream);
    if (rc != EPNET_OK || b == 0 || m == 0) return rc;
    if (b > 65535) return EPNET_ERR_BAD_ARG;
    fps_prefix_fill_kernel<<<dim3((m + 255) / 256, b), 256, 0, (cudaStream_t)stream>>>(n, m, xyz, idx, new_xyz, aux_in, aux_out, aux_dim, identity);
    EPNET_RETURN_LAUNCH_STATUS();
}
