// Rotated bird's-eye-view box overlap / IoU and NMS for B200 -- SURVEY.md section 8(f), rank 2.  Replaces the four launchers of
// /root/reference/lib/utils/iou3d/src/iou3d_kernel.cu (:352-387) AND the host half of nms_gpu / nms_normal_gpu
// (iou3d.cpp:74-170: cudaMalloc of the mask, blocking D2H copy of N x N/64 words, serial greedy loop on the host, cudaFree).
//
// What is different here:
//  * everything that depends on ONE box (centre, the four rotated corners, cos/sin of the inverse rotation, the bounds with
//    the 1e-5 margin, a bounding radius) is computed once per box and CTA (BoxGeo) instead of once per PAIR;
//  * pairs whose bounding circles are farther apart than a conservative slack are answered 0 without entering the clipping
//    code (the reference walks 16 segment tests + 8 in-box tests to find the same 0);
//  * NMS evaluates only the 64x64 tiles on or above the diagonal (the greedy scan never reads the others: iou3d.cpp:107-109
//    starts at j = nblock) and the greedy scan itself runs on the device: one CTA per problem, a 64-box diagonal tile is
//    resolved by one thread from shared memory, the kept rows are OR-ed into the removal words by the whole CTA; it stops as
//    soon as `max_out` boxes are kept (the proposal layer truncates right after: lib/rpn/proposal_layer.py:111);
//  * several independent problems (scenes x distance bands) go in one launch (grid.z / grid.x), no allocation, no sync.
//
// What is NOT different, because the keep-set is an index output under the bit-exact bar: the floating-point expression of
// the overlap.  Each step below follows the SASS (cuobjdump) of the unmodified reference file built with its own flags
// (nvcc -O2) -- the SASS, not the PTX: the PTX still shows most `a*b - c*d` as mul, mul, sub without rounding modifiers, and
// ptxas contracts those afterwards (a first version written from the PTX was off by one ulp on 74 % of the overlapping pairs).
// In all four reference kernels: a*b - c*d = fma(a, b, -(c*d)) -- second product rounded, first fused -- for the cross
// products s1/s3/s4, both intersection formulas, the rotated y and the polygon-area terms; s2 and s5 = -s2 share their two
// products across a branch and keep them as rounded multiplies; rotated x = cx + fma(cos, dx, sin*dy); polygon centre by
// div.rn with float(cnt); atan2f ordering; |area| * 0.5f; union = fma(wa, ha, wb*hb) - overlap for the rotated IoU and
// fma(wb, hb, wa*ha) - inter for the axis-aligned one.  Written with __f*_rn intrinsics so that no compiler flag or ptxas
// version can change the contraction.  The reference bubble-sorts the
// polygon vertices with a strict `>` on the angle = a stable ascending sort; we compute each angle once and insertion-sort
// stably, which yields the same permutation.
#include "common.cuh"
#include <stdlib.h>

namespace epnet {

constexpr int kNmsTile = 64;          // boxes per mask word (reference THREADS_PER_BLOCK_NMS)
constexpr int kMaxPoly = 24;          // 8 crossings + 8 corners is the geometric maximum; head-room instead of the reference's 16
constexpr float kEpsIou = 1e-8f;      // reference EPS (iou3d_kernel.cu:13)
constexpr float kMargin = 1e-5f;      // reference MARGIN (:51)

struct BoxGeo {
    float x1, y1, x2, y2;     // the box as given (axis-aligned extents before rotation)
    float cx, cy;             // centre
    float px[4], py[4];       // corners rotated about the centre, order (x1,y1) (x2,y1) (x2,y2) (x1,y2)
    float ic, is;             // cos(-angle), sin(-angle): inverse rotation for the in-box test
    float rad;                // >= circumscribed radius (early-out only, never enters a reported number)
};

// a*b - c*d the way the reference binary evaluates it: c*d rounded, a*b fused into the subtraction
__device__ __forceinline__ float pmp(float a, float b, float c, float d) { return __fmaf_rn(a, b, -__fmul_rn(c, d)); }

__device__ __forceinline__ void rotate_ref(float cx, float cy, float c, float s, float x, float y, float &ox, float &oy)
{
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy);
    ox = __fadd_rn(cx, __fmaf_rn(c, dx, __fmul_rn(s, dy)));
    oy = __fadd_rn(cy, pmp(c, dy, s, dx));
}

__device__ __forceinline__ void box_prepare(const float *__restrict__ b, BoxGeo &g)
{
    g.x1 = __ldg(b), g.y1 = __ldg(b + 1), g.x2 = __ldg(b + 2), g.y2 = __ldg(b + 3);
    const float angle = __ldg(b + 4);
    g.cx = __fmul_rn(__fadd_rn(g.x1, g.x2), 0.5f);
    g.cy = __fmul_rn(__fadd_rn(g.y1, g.y2), 0.5f);
    const float c = cosf(angle), s = sinf(angle);
    rotate_ref(g.cx, g.cy, c, s, g.x1, g.y1, g.px[0], g.py[0]);
    rotate_ref(g.cx, g.cy, c, s, g.x2, g.y1, g.px[1], g.py[1]);
    rotate_ref(g.cx, g.cy, c, s, g.x2, g.y2, g.px[2], g.py[2]);
    rotate_ref(g.cx, g.cy, c, s, g.x1, g.y2, g.px[3], g.py[3]);
    g.ic = cosf(-angle), g.is = sinf(-angle);
    const float hw = fabsf(g.x2 - g.x1) * 0.5f, hh = fabsf(g.y2 - g.y1) * 0.5f;
    g.rad = sqrtf(hw * hw + hh * hh);
}

// reference `cross(p1, p2, p0)` (:38-40): (p1-p0) x (p2-p0)
__device__ __forceinline__ float cross3(float p1x, float p1y, float p2x, float p2y, float p0x, float p0y)
{
    return pmp(__fsub_rn(p1x, p0x), __fsub_rn(p2y, p0y), __fsub_rn(p2x, p0x), __fsub_rn(p1y, p0y));
}

// reference `intersection(p1, p0, q1, q0, ans)` (:66-96)
__device__ __forceinline__ bool segment_cross(float p1x, float p1y, float p0x, float p0y, float q1x, float q1y, float q0x, float q0y,
                                              float &ax, float &ay)
{
    if (!(fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) && fminf(p0y, p1y) <= fmaxf(q0y, q1y) &&
          fminf(q0y, q1y) <= fmaxf(p0y, p1y)))
        return false;
    const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
    const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
    const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
    // s2 = cross(p1, q1, p0) and s5 = cross(q1, p1, p0) = -s2 are built from the same two rounded products
    const float m73 = __fmul_rn(__fsub_rn(p1x, p0x), __fsub_rn(q1y, p0y)), m74 = __fmul_rn(__fsub_rn(q1x, p0x), __fsub_rn(p1y, p0y));
    const float s2 = __fsub_rn(m73, m74);
    if (!(__fmul_rn(s1, s2) > 0.0f && __fmul_rn(s3, s4) > 0.0f)) return false;
    const float s5 = __fsub_rn(m74, m73);
    const float den = __fsub_rn(s5, s1);
    if (fabsf(den) > kEpsIou) {
        ax = __fdiv_rn(pmp(s5, q0x, s1, q1x), den);
        ay = __fdiv_rn(pmp(s5, q0y, s1, q1y), den);
    } else {
        const float a0 = __fsub_rn(p0y, p1y), b0 = __fsub_rn(p1x, p0x), c0 = pmp(p0x, p1y, p1x, p0y);
        const float a1 = __fsub_rn(q0y, q1y), b1 = __fsub_rn(q1x, q0x), c1 = pmp(q0x, q1y, q1x, q0y);
        const float D = pmp(a0, b1, a1, b0);
        ax = __fdiv_rn(pmp(b0, c1, b1, c0), D);
        ay = __fdiv_rn(pmp(a1, c0, a0, c1), D);
    }
    return true;
}

// reference `check_in_box2d(box, p)` (:49-64): the point turned back into the box frame, open interval widened by MARGIN
__device__ __forceinline__ bool inside_box(const BoxGeo &g, float x, float y)
{
    const float dx = __fsub_rn(x, g.cx), dy = __fsub_rn(y, g.cy);
    const float rx = __fadd_rn(g.cx, __fmaf_rn(g.ic, dx, __fmul_rn(dy, g.is)));
    const float ry = __fadd_rn(g.cy, pmp(g.ic, dy, g.is, dx));
    return rx > __fadd_rn(g.x1, -kMargin) && rx < __fadd_rn(g.x2, kMargin) && ry > __fadd_rn(g.y1, -kMargin) && ry < __fadd_rn(g.y2, kMargin);
}

// Boxes that cannot touch: centre distance beyond both circumscribed radii plus a slack that dwarfs every rounding error of
// the corner arithmetic (a few ulp of the coordinates).  The reference finds no crossing and no contained corner for such a
// pair and returns |0| / 2 = 0.
__device__ __forceinline__ bool surely_disjoint(const BoxGeo &A, const BoxGeo &B)
{
    const float dx = A.cx - B.cx, dy = A.cy - B.cy;
    const float reach = A.rad + B.rad;
    const float slack = 0.05f + 1e-4f * (fabsf(A.cx) + fabsf(A.cy) + fabsf(B.cx) + fabsf(B.cy) + reach);
    const float lim = reach + slack;
    return dx * dx + dy * dy > lim * lim;   // false for NaN: those pairs take the full path like the reference
}

// reference `box_overlap(box_a, box_b)` (:108-225)
__device__ float overlap_ref(const BoxGeo &A, const BoxGeo &B)
{
    if (surely_disjoint(A, B)) return 0.0f;
    float qx[kMaxPoly], qy[kMaxPoly];
    int cnt = 0;
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x, y;
            if (segment_cross(A.px[(i + 1) & 3], A.py[(i + 1) & 3], A.px[i], A.py[i], B.px[(j + 1) & 3], B.py[(j + 1) & 3], B.px[j], B.py[j], x,
                              y)) {
                if (cnt < kMaxPoly) qx[cnt] = x, qy[cnt] = y;
                sx = __fadd_rn(sx, x), sy = __fadd_rn(sy, y);
                ++cnt;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (inside_box(A, B.px[k], B.py[k])) {
            if (cnt < kMaxPoly) qx[cnt] = B.px[k], qy[cnt] = B.py[k];
            sx = __fadd_rn(sx, B.px[k]), sy = __fadd_rn(sy, B.py[k]);
            ++cnt;
        }
        if (inside_box(B, A.px[k], A.py[k])) {
            if (cnt < kMaxPoly) qx[cnt] = A.px[k], qy[cnt] = A.py[k];
            sx = __fadd_rn(sx, A.px[k]), sy = __fadd_rn(sy, A.py[k]);
            ++cnt;
        }
    }
    if (cnt < 3) return 0.0f;   // the reference's area loop sums terms that all carry the factor (x0 - x0): 0
    cnt = min(cnt, kMaxPoly);
    const float fc = (float)cnt;
    const float mx = __fdiv_rn(sx, fc), my = __fdiv_rn(sy, fc);

    // stable ascending order of the angle about the centre (== the reference's bubble sort with a strict `>`)
    float ang[kMaxPoly];
    for (int k = 0; k < cnt; ++k) ang[k] = atan2f(__fsub_rn(qy[k], my), __fsub_rn(qx[k], mx));
    for (int k = 1; k < cnt; ++k) {
        const float a = ang[k], x = qx[k], y = qy[k];
        int t = k - 1;
        while (t >= 0 && ang[t] > a) {
            ang[t + 1] = ang[t], qx[t + 1] = qx[t], qy[t + 1] = qy[t];
            --t;
        }
        ang[t + 1] = a, qx[t + 1] = x, qy[t + 1] = y;
    }

    float area = 0.0f;
    const float x0 = qx[0], y0 = qy[0];
    for (int k = 0; k < cnt - 1; ++k) {
        const float ux = __fsub_rn(qx[k], x0), uy = __fsub_rn(qy[k], y0);
        const float vx = __fsub_rn(qx[k + 1], x0), vy = __fsub_rn(qy[k + 1], y0);
        area = __fadd_rn(area, pmp(ux, vy, uy, vx));
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

// reference `iou_bev(box_a, box_b)` (:227-234)
__device__ __forceinline__ float iou_rotated_ref(const BoxGeo &A, const BoxGeo &B)
{
    const float sb = __fmul_rn(__fsub_rn(B.x2, B.x1), __fsub_rn(B.y2, B.y1));
    const float sum = __fmaf_rn(__fsub_rn(A.x2, A.x1), __fsub_rn(A.y2, A.y1), sb);
    const float ov = overlap_ref(A, B);
    return __fdiv_rn(ov, fmaxf(__fsub_rn(sum, ov), kEpsIou));
}

// reference `iou_normal(a, b)` (:295-303); a = the row box
__device__ __forceinline__ float iou_axis_ref(const BoxGeo &A, const BoxGeo &B)
{
    const float left = fmaxf(A.x1, B.x1), right = fminf(A.x2, B.x2), top = fmaxf(A.y1, B.y1), bottom = fminf(A.y2, B.y2);
    const float w = fmaxf(__fsub_rn(right, left), 0.0f), h = fmaxf(__fsub_rn(bottom, top), 0.0f);
    const float inter = __fmul_rn(w, h);
    const float sa = __fmul_rn(__fsub_rn(A.x2, A.x1), __fsub_rn(A.y2, A.y1));
    const float sum = __fmaf_rn(__fsub_rn(B.x2, B.x1), __fsub_rn(B.y2, B.y1), sa);
    return __fdiv_rn(inter, fmaxf(__fsub_rn(sum, inter), kEpsIou));
}

// ---- pairwise matrices -------------------------------------------------------------------------------------------------------

constexpr int kPairTile = 16;

template <bool kIou>
__global__ void __launch_bounds__(kPairTile *kPairTile)
pair_matrix_kernel(int num_a, const float *__restrict__ boxes_a, int num_b, const float *__restrict__ boxes_b, float *__restrict__ out)
{
    __shared__ BoxGeo ga[kPairTile], gb[kPairTile];
    const int a0 = blockIdx.y * kPairTile, b0 = blockIdx.x * kPairTile;
    const int t = threadIdx.x;
    if (t < kPairTile) {
        if (a0 + t < num_a) box_prepare(boxes_a + (size_t)(a0 + t) * 5, ga[t]);
    } else if (t < 2 * kPairTile) {
        if (b0 + t - kPairTile < num_b) box_prepare(boxes_b + (size_t)(b0 + t - kPairTile) * 5, gb[t - kPairTile]);
    }
    __syncthreads();
    const int ia = t / kPairTile, ib = t % kPairTile;
    if (a0 + ia >= num_a || b0 + ib >= num_b) return;
    const BoxGeo A = ga[ia], B = gb[ib];
    out[(size_t)(a0 + ia) * num_b + (b0 + ib)] = kIou ? iou_rotated_ref(A, B) : overlap_ref(A, B);
}

// ---- NMS: suppression words for the tiles on/above the diagonal ----------------------------------------------------------------

template <bool kRotated>
__global__ void __launch_bounds__(kNmsTile)
nms_mask_kernel(int nmax, const int *__restrict__ counts, float thresh, const float *__restrict__ boxes, unsigned long long *__restrict__ mask)
{
    const int col = blockIdx.x, row = blockIdx.y, seg = blockIdx.z;
    if (col < row) return;
    const int n = counts ? min(max(__ldg(counts + seg), 0), nmax) : nmax;
    if (row * kNmsTile >= n || col * kNmsTile >= n) return;
    const int stride = (nmax + kNmsTile - 1) / kNmsTile;
    boxes += (size_t)seg * nmax * 5;
    mask += (size_t)seg * nmax * stride;

    __shared__ BoxGeo gcol[kNmsTile];
    const int t = threadIdx.x;
    const int col_size = min(n - col * kNmsTile, kNmsTile), row_size = min(n - row * kNmsTile, kNmsTile);
    if (t < col_size) box_prepare(boxes + (size_t)(col * kNmsTile + t) * 5, gcol[t]);
    __syncthreads();
    if (t >= row_size) return;
    const int i = row * kNmsTile + t;
    BoxGeo A;
    if (row == col) A = gcol[t];
    else box_prepare(boxes + (size_t)i * 5, A);
    unsigned long long word = 0;
    for (int j = (row == col) ? t + 1 : 0; j < col_size; ++j) {
        const float v = kRotated ? iou_rotated_ref(A, gcol[j]) : iou_axis_ref(A, gcol[j]);
        if (v > thresh) word |= 1ULL << j;
    }
    mask[(size_t)i * stride + col] = word;
}

// Rotated variant with the work re-dealt.  ncu on the kernel above (6300 clustered proposals): 5.4 of 32 lanes active per
// instruction -- a lane whose pair passes the bounding-circle test walks the clipping code while its 31 neighbours wait.
// Here a tile is done in two passes: (1) all 64x64 pairs take the circle test, survivors are appended to a list in shared
// memory (one atomicAdd per warp); (2) the list is dealt densely to the lanes, each evaluating the full IoU of its pairs and
// OR-ing the verdict into the row's word in shared memory.  Same arithmetic per pair, hence the same words.
constexpr int kCompactThreads = 128;

__global__ void __launch_bounds__(kCompactThreads)
nms_mask_compact_kernel(int nmax, const int *__restrict__ counts, float thresh, const float *__restrict__ boxes, unsigned long long *__restrict__ mask)
{
    const int col = blockIdx.x, row = blockIdx.y, seg = blockIdx.z;
    if (col < row) return;
    const int n = counts ? min(max(__ldg(counts + seg), 0), nmax) : nmax;
    if (row * kNmsTile >= n || col * kNmsTile >= n) return;
    const int stride = (nmax + kNmsTile - 1) / kNmsTile;
    boxes += (size_t)seg * nmax * 5;
    mask += (size_t)seg * nmax * stride;

    __shared__ BoxGeo geo[2 * kNmsTile];                 // [0,64): column boxes, [64,128): row boxes (unused on the diagonal)
    __shared__ unsigned short cand[kNmsTile * kNmsTile]; // (row << 6) | col of the pairs that may touch
    __shared__ unsigned long long words[kNmsTile];
    __shared__ int ncand;

    const int t = threadIdx.x, lane = t & 31;
    const bool diagonal = row == col;
    const int col_size = min(n - col * kNmsTile, kNmsTile), row_size = min(n - row * kNmsTile, kNmsTile);
    if (t < kNmsTile) {
        if (t < col_size) box_prepare(boxes + (size_t)(col * kNmsTile + t) * 5, geo[t]);
        words[t] = 0;
    } else if (!diagonal && t - kNmsTile < row_size) {
        box_prepare(boxes + (size_t)(row * kNmsTile + t - kNmsTile) * 5, geo[t]);
    }
    if (t == 0) ncand = 0;
    __syncthreads();
    const BoxGeo *gcol = geo, *grow = diagonal ? geo : geo + kNmsTile;

    // pass 1: a warp covers 32 consecutive columns of one row per step
    for (int p = t; p < kNmsTile * kNmsTile; p += kCompactThreads) {
        const int i = p >> 6, j = p & 63;
        const bool live = i < row_size && j < col_size && (!diagonal || j > i) && !surely_disjoint(grow[i], gcol[j]);
        const uint32_t ballot = __ballot_sync(0xffffffffu, live);
        if (ballot) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&ncand, __popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (live) cand[base + __popc(ballot & lanemask_lt())] = (unsigned short)p;
        }
    }
    __syncthreads();

    // pass 2: dense
    const int total = ncand;
    for (int c = t; c < total; c += kCompactThreads) {
        const int p = cand[c], i = p >> 6, j = p & 63;
        if (iou_rotated_ref(grow[i], gcol[j]) > thresh) atomicOr(&words[i], 1ULL << j);
    }
    __syncthreads();
    if (t < row_size) mask[(size_t)(row * kNmsTile + t) * stride + col] = words[t];
}

// ---- NMS: the greedy scan (reference: host loop iou3d.cpp:100-113) -----------------------------------------------------------------

constexpr int kReduceThreads = 256;

__global__ void __launch_bounds__(kReduceThreads)
nms_scan_kernel(int nmax, const int *__restrict__ counts, const unsigned long long *__restrict__ mask, int max_out, long long *__restrict__ keep,
                int *__restrict__ num_out)
{
    extern __shared__ unsigned long long removed[];   // one bit per box of this problem
    __shared__ unsigned long long diag[kNmsTile];
    __shared__ unsigned long long kept_word;
    __shared__ int kept_rows[kNmsTile];

    const int seg = blockIdx.x, t = threadIdx.x;
    const int n = counts ? min(max(__ldg(counts + seg), 0), nmax) : nmax;
    const int stride = (nmax + kNmsTile - 1) / kNmsTile, tiles = (n + kNmsTile - 1) / kNmsTile;
    const int limit = (max_out > 0 && max_out < n) ? max_out : n;
    mask += (size_t)seg * nmax * stride;
    keep += (size_t)seg * nmax;

    for (int j = t; j < tiles; j += kReduceThreads) removed[j] = 0;
    int kept = 0;
    unsigned long long diag_next = 0;
    __syncthreads();
    for (int tile = 0; tile < tiles && kept < limit; ++tile) {
        const int base = tile * kNmsTile, size = min(kNmsTile, n - base);
        if (t < kNmsTile) {
            diag[t] = tile == 0 ? (t < size ? __ldg(mask + (size_t)t * stride) : 0ULL) : diag_next;
            // the next tile's diagonal words travel while this tile is resolved and its rows are OR-ed
            const int nb = base + kNmsTile + t;
            diag_next = (tile + 1 < tiles && nb < n) ? __ldg(mask + (size_t)nb * stride + tile + 1) : 0ULL;
        }
        __syncthreads();
        if (t == 0) {
            // boxes of this tile in index order; only the survivors cost an iteration
            const unsigned long long valid = size == kNmsTile ? ~0ULL : ((1ULL << size) - 1ULL);
            unsigned long long r = removed[tile], kw = 0, cand = ~r & valid;
            while (cand) {
                const int i = __ffsll((long long)cand) - 1;
                kw |= 1ULL << i;
                r |= diag[i];
                cand &= ~r & ~((2ULL << i) - 1ULL);
            }
            kept_word = kw;
        }
        __syncthreads();
        const unsigned long long kw = kept_word;
        const int kc = __popcll(kw);
        if (t < kNmsTile && ((kw >> t) & 1ULL)) {
            const int rank = __popcll(kw & ((1ULL << t) - 1ULL));
            kept_rows[rank] = base + t;
            if (kept + rank < limit) keep[kept + rank] = base + t;
        }
        kept += kc;
        __syncthreads();
        // suppression words of the kept rows for the tiles still to come: (row, word) pairs dealt to all threads, loads in
        // flight four at a time, the (mostly zero) words OR-ed into shared memory
        const int rem = tiles - tile - 1;
        if (rem > 0 && kept < limit) {
            const int total = kc * rem;
            for (int e = t; e < total; e += 4 * kReduceThreads) {
                unsigned long long v[4];
                int w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ee = e + u * kReduceThreads;
                    v[u] = 0;
                    w[u] = 0;
                    if (ee < total) {
                        const int r = ee / rem;
                        w[u] = tile + 1 + (ee - r * rem);
                        v[u] = __ldg(mask + (size_t)kept_rows[r] * stride + w[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (v[u]) atomicOr(&removed[w[u]], v[u]);
            }
        }
        __syncthreads();
    }
    if (t == 0) num_out[seg] = min(kept, limit);
}

// EPNET_NMS_PLAIN_MASK=1 selects the thread-per-row rotated mask kernel (kept for comparison; both produce the same words)
static const bool g_plain_rotated_mask = [] {
    const char *e = getenv("EPNET_NMS_PLAIN_MASK");
    return e && e[0] == '1';
}();

template <bool kRotated>
static int launch_nms(int s, int n, const float *boxes, const int *counts, float thresh, int max_out, void *workspace, long long *keep, int *num_out,
                      cudaStream_t st)
{
    if (s < 0 || n < 0 || s > 65535 || (s > 0 && n > 0 && (!boxes || !workspace || !keep)) || (s > 0 && !num_out)) return EPNET_ERR_BAD_ARG;
    if (s == 0) return EPNET_OK;
    const int tiles = (n + kNmsTile - 1) / kNmsTile;
    if (tiles > 65535 || (size_t)tiles * sizeof(unsigned long long) > 200 * 1024) return EPNET_ERR_BAD_ARG;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_out, 0, sizeof(int) * s, st);
        return e == cudaSuccess ? EPNET_OK : (int)e;
    }
    if (kRotated && !g_plain_rotated_mask)
        nms_mask_compact_kernel<<<dim3(tiles, tiles, s), kCompactThreads, 0, st>>>(n, counts, thresh, boxes, (unsigned long long *)workspace);
    else
        nms_mask_kernel<kRotated><<<dim3(tiles, tiles, s), kNmsTile, 0, st>>>(n, counts, thresh, boxes, (unsigned long long *)workspace);
    const size_t smem = (size_t)tiles * sizeof(unsigned long long);
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    nms_scan_kernel<<<s, kReduceThreads, smem, st>>>(n, counts, (const unsigned long long *)workspace, max_out, keep, num_out);
    EPNET_RETURN_LAUNCH_STATUS();
}

}  // namespace epnet

EPNET_API int epnet_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_overlap, void *stream)
{
    using namespace epnet;
    if (num_a < 0 || num_b < 0) return EPNET_ERR_BAD_ARG;
    if (num_a == 0 || num_b == 0) return EPNET_OK;
    if (!boxes_a || !boxes_b || !ans_overlap || (num_a + kPairTile - 1) / kPairTile > 65535) return EPNET_ERR_BAD_ARG;
    pair_matrix_kernel<false><<<dim3((num_b + kPairTile - 1) / kPairTile, (num_a + kPairTile - 1) / kPairTile), kPairTile * kPairTile, 0,
                                (cudaStream_t)stream>>>(num_a, boxes_a, num_b, boxes_b, ans_overlap);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou, void *stream)
{
    using namespace epnet;
    if (num_a < 0 || num_b < 0) return EPNET_ERR_BAD_ARG;
    if (num_a == 0 || num_b == 0) return EPNET_OK;
    if (!boxes_a || !boxes_b || !ans_iou || (num_a + kPairTile - 1) / kPairTile > 65535) return EPNET_ERR_BAD_ARG;
    pair_matrix_kernel<true><<<dim3((num_b + kPairTile - 1) / kPairTile, (num_a + kPairTile - 1) / kPairTile), kPairTile * kPairTile, 0,
                               (cudaStream_t)stream>>>(num_a, boxes_a, num_b, boxes_b, ans_iou);
    EPNET_RETURN_LAUNCH_STATUS();
}

EPNET_API int epnet_nms_workspace_bytes(int s, int n, unsigned long long *bytes)
{
    if (s < 0 || n < 0 || !bytes) return EPNET_ERR_BAD_ARG;
    const unsigned long long tiles = ((unsigned long long)n + epnet::kNmsTile - 1) / epnet::kNmsTile;
    *bytes = (unsigned long long)s * (unsigned long long)n * tiles * sizeof(unsigned long long);
    return EPNET_OK;
}

EPNET_API int epnet_nms_rotated(int s, int n, const float *boxes, const int *counts, float thresh, int max_out, void *workspace, long long *keep,
                                int *num_out, void *stream)
{
    return epnet::launch_nms<true>(s, n, boxes, counts, thresh, max_out, workspace, keep, num_out, (cudaStream_t)stream);
}

EPNET_API int epnet_nms_normal(int s, int n, const float *boxes, const int *counts, float thresh, int max_out, void *workspace, long long *keep,
                               int *num_out, void *stream)
{
    return epnet::launch_nms<false>(s, n, boxes, counts, thresh, max_out, workspace, keep, num_out, (cudaStream_t)stream);
}
