/*
 * oracle/pointnet2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the arithmetic of EPNet's pointnet2 CUDA ops and of the
 * LI-Fusion bilinear image gather.  It exists so that the CUDA kernels in
 * epnet_b200/csrc can be checked bit-for-bit (indices) and to 1e-5 (floats).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: the reference ships no tests or golden vectors, so this
 * file is pinned against the reference's own, unmodified kernels compiled into
 * oracle/_ref/libpointnet2_ref.so and executed on a B200 (tests/test_ref_pin.py, `-m gpu`),
 * and against fixtures under tests/golden/ produced by that library.
 *
 * All distance computations reproduce the FMA contraction nvcc -O2 emits for the reference
 * source (checked in PTX): d = fmaf(dz,dz, fmaf(dx,dx, dy*dy)).  Build with -ffp-contract=off.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/pointnet2_lib/pointnet2/src/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline float sqdist_contracted(float ax, float ay, float az, float bx, float by, float bz)
{
    /* (a-b) per axis, then y*y first, x fused on top, z fused last: the order nvcc picks. */
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* cuda_utils.h:10-14 -- thread count of the FPS block: largest power of two <= n, capped at 1024.
 * The reference computes it through a double log; this integer form is equal for every n we use
 * (asserted against the log formula in tests/test_oracle.py). */
ORACLE_API int oracle_fps_block_size(int n)
{
    int p = 1;
    while (p * 2 <= n && p * 2 <= 1024) p *= 2;
    return p;
}

/* sampling_gpu.cu:93-209 -- furthest point sampling, including the shared-memory tree reduction
 * whose tie handling (keep the lower slot unless the upper one is strictly larger) decides which
 * of several equally distant points is sampled.  temp is read AND written, as in the reference:
 * the caller pre-fills it with 1e10 (pointnet2_utils.py:26). */
ORACLE_API void oracle_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx)
{
    if (m <= 0) return;
    const int bs = oracle_fps_block_size(n);
    float *slot_val = (float *)malloc(sizeof(float) * (size_t)bs);
    int *slot_idx = (int *)malloc(sizeof(int) * (size_t)bs);

    for (int s = 0; s < b; ++s) {
        const float *p = xyz + (size_t)s * n * 3;
        float *t = temp + (size_t)s * n;
        int *out = idx + (size_t)s * m;
        int last = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float lx = p[last * 3 + 0], ly = p[last * 3 + 1], lz = p[last * 3 + 2];
            /* per-"thread" strided scan: strict '>' keeps the first maximum (:129-138) */
            for (int tid = 0; tid < bs; ++tid) {
                float best = -1.0f;
                int besti = 0;
                for (int k = tid; k < n; k += bs) {
                    const float d = sqdist_contracted(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], lx, ly, lz);
                    const float d2 = fminf(d, t[k]);
                    t[k] = d2;
                    if (d2 > best) { best = d2; besti = k; }
                }
                slot_val[tid] = best;
                slot_idx[tid] = besti;
            }
            /* tree reduction (:143-203) with __update (:86-91) */
            for (int half = bs / 2; half >= 1; half /= 2) {
                for (int tid = 0; tid < half; ++tid) {
                    const float v1 = slot_val[tid], v2 = slot_val[tid + half];
                    if (v2 > v1) { slot_val[tid] = v2; slot_idx[tid] = slot_idx[tid + half]; }
                }
            }
            last = slot_idx[0];
            out[j] = last;
        }
    }
    free(slot_val);
    free(slot_idx);
}

/* sampling_gpu.cu:8-24 */
ORACLE_API void oracle_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out)
{
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *src = points + ((size_t)s * c + ch) * n;
            float *dst = out + ((size_t)s * c + ch) * m;
            const int *id = idx + (size_t)s * m;
            for (int j = 0; j < m; ++j) dst[j] = src[id[j]];
        }
}

/* sampling_gpu.cu:46-63 -- scatter-add; the reference's atomic order is unspecified, so callers
 * compare with a tolerance.  grad_points must arrive zeroed (pointnet2_utils.py:67). */
ORACLE_API void oracle_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points)
{
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *g = grad_out + ((size_t)s * c + ch) * m;
            float *dst = grad_points + ((size_t)s * c + ch) * n;
            const int *id = idx + (size_t)s * m;
            for (int j = 0; j < m; ++j) dst[id[j]] += g[j];
        }
}

/* ball_query_gpu.cu:9-45 -- ascending scan, strict d2 < r*r (r*r rounded in float, :23); the first
 * hit fills every slot, later hits overwrite slots 1..; idx arrives zeroed (pointnet2_utils.py:218)
 * so a centre without any hit keeps zeros. */
ORACLE_API void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx)
{
    const float r2 = radius * radius;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < b; ++s)
        for (int j = 0; j < m; ++j) {
            const float *q = new_xyz + ((size_t)s * m + j) * 3;
            const float *p = xyz + (size_t)s * n * 3;
            int *o = idx + ((size_t)s * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sqdist_contracted(q[0], q[1], q[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < r2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt++] = k;
                }
            }
        }
}

/* group_points_gpu.cu:47-66 */
ORACLE_API void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out)
{
    const size_t per = (size_t)npoints * nsample;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *src = points + ((size_t)s * c + ch) * n;
            float *dst = out + ((size_t)s * c + ch) * per;
            const int *id = idx + (size_t)s * per;
            for (size_t e = 0; e < per; ++e) dst[e] = src[id[e]];
        }
}

/* group_points_gpu.cu:8-25 */
ORACLE_API void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points)
{
    const size_t per = (size_t)npoints * nsample;
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *g = grad_out + ((size_t)s * c + ch) * per;
            float *dst = grad_points + ((size_t)s * c + ch) * n;
            const int *id = idx + (size_t)s * per;
            for (size_t e = 0; e < per; ++e) dst[id[e]] += g[e];
        }
}

/* interpolate_gpu.cu:9-52 -- three smallest squared distances by a strict-'<' cascade: equal
 * distances keep the lower index first.  The reference holds the running bests in double (init
 * 1e40) but compares exactly-widened floats, so float compares with an "unset" marker are
 * equivalent; unset slots come out as +inf / index 0, the float conversion of 1e40. */
ORACLE_API void oracle_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < b; ++s)
        for (int j = 0; j < n; ++j) {
            const float *u = unknown + ((size_t)s * n + j) * 3;
            const float *kn = known + (size_t)s * m * 3;
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                const double d = (double)sqdist_contracted(u[0], u[1], u[2], kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < b1)      { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            float *od = dist2 + ((size_t)s * n + j) * 3;
            int *oi = idx + ((size_t)s * n + j) * 3;
            od[0] = (float)b1; od[1] = (float)b2; od[2] = (float)b3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

/* interpolate_gpu.cu:77-97 -- contraction order from PTX: fma(w2,p2, fma(w0,p0, w1*p1)). */
ORACLE_API void oracle_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *src = points + ((size_t)s * c + ch) * m;
            float *dst = out + ((size_t)s * c + ch) * n;
            for (int j = 0; j < n; ++j) {
                const int *id = idx + ((size_t)s * n + j) * 3;
                const float *w = weight + ((size_t)s * n + j) * 3;
                dst[j] = fmaf(w[2], src[id[2]], fmaf(w[0], src[id[0]], w[1] * src[id[1]]));
            }
        }
}

/* interpolate_gpu.cu:120-142 */
ORACLE_API void oracle_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points)
{
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *g = grad_out + ((size_t)s * c + ch) * n;
            float *dst = grad_points + ((size_t)s * c + ch) * m;
            for (int j = 0; j < n; ++j) {
                const int *id = idx + ((size_t)s * n + j) * 3;
                const float *w = weight + ((size_t)s * n + j) * 3;
                dst[id[0]] += g[j] * w[0];
                dst[id[1]] += g[j] * w[1];
                dst[id[2]] += g[j] * w[2];
            }
        }
}

/* LI-Fusion gather: lib/net/pointnet2_msg.py:107-120 calls torch.nn.functional.grid_sample
 * (ATen grid_sampler_2d; third-party, torch 2.11.0 in this image, not under /root/reference) with
 * mode='bilinear', padding_mode='zeros'.  Restated from its published algorithm: unnormalise
 *   align_corners: ix = (x+1)/2*(W-1)      else: ix = ((x+1)*W-1)/2
 * floor to the NW corner, weights nw=(ixe-ix)(iye-iy) ..., out-of-range taps contribute zero,
 * taps summed in the order nw, ne, sw, se.  Anchored on torch's own CPU grid_sample in
 * tests/test_oracle.py; float, not bit-exact.  Where device code contracts a multiply-add into one FMA
 * the same FMA is written out here, so that CPU and GPU agree to the last bits on the sampling position
 * (a 1-ulp difference in a pixel coordinate of ~1000 would otherwise show up as ~1e-4 in the output). */
ORACLE_API void oracle_grid_gather_bilinear(int b, int c, int h, int w, int n, const float *fmap, const float *xy, int align_corners, float *out)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            const float *img = fmap + ((size_t)s * c + ch) * h * w;
            float *dst = out + ((size_t)s * c + ch) * n;
            for (int j = 0; j < n; ++j) {
                const float gx = xy[((size_t)s * n + j) * 2 + 0], gy = xy[((size_t)s * n + j) * 2 + 1];
                float ix, iy;
                if (align_corners) {
                    ix = ((gx + 1.f) / 2.f) * (float)(w - 1);
                    iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
                } else { /* (x+1)*W-1 contracts to one FMA in device code (nvcc -fmad=true, ATen's build too) */
                    ix = fmaf(gx + 1.f, (float)w, -1.f) / 2.f;
                    iy = fmaf(gy + 1.f, (float)h, -1.f) / 2.f;
                }
                const float fx = floorf(ix), fy = floorf(iy);
                const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
                const float wne = (ix - (float)x0) * ((float)y1 - iy);
                const float wsw = ((float)x1 - ix) * (iy - (float)y0);
                const float wse = (ix - (float)x0) * (iy - (float)y0);
                float acc = 0.f; /* acc += v*w is one FMA per tap in device code */
                if (x0 >= 0 && x0 < w && y0 >= 0 && y0 < h) acc = fmaf(img[(size_t)y0 * w + x0], wnw, acc);
                if (x1 >= 0 && x1 < w && y0 >= 0 && y0 < h) acc = fmaf(img[(size_t)y0 * w + x1], wne, acc);
                if (x0 >= 0 && x0 < w && y1 >= 0 && y1 < h) acc = fmaf(img[(size_t)y1 * w + x0], wsw, acc);
                if (x1 >= 0 && x1 < w && y1 >= 0 && y1 < h) acc = fmaf(img[(size_t)y1 * w + x1], wse, acc);
                dst[j] = acc;
            }
        }
}

/* Backward of the gather w.r.t. the feature map (xy carries no gradient on the hot path). */
ORACLE_API void oracle_grid_gather_bilinear_grad(int b, int c, int h, int w, int n, const float *grad_out, const float *xy, int align_corners, float *grad_fmap)
{
    for (int s = 0; s < b; ++s)
        for (int ch = 0; ch < c; ++ch) {
            float *img = grad_fmap + ((size_t)s * c + ch) * h * w;
            const float *g = grad_out + ((size_t)s * c + ch) * n;
            for (int j = 0; j < n; ++j) {
                const float gx = xy[((size_t)s * n + j) * 2 + 0], gy = xy[((size_t)s * n + j) * 2 + 1];
                float ix, iy;
                if (align_corners) {
                    ix = ((gx + 1.f) / 2.f) * (float)(w - 1);
                    iy = ((gy + 1.f) / 2.f) * (float)(h - 1);
                } else { /* (x+1)*W-1 contracts to one FMA in device code (nvcc -fmad=true, ATen's build too) */
                    ix = fmaf(gx + 1.f, (float)w, -1.f) / 2.f;
                    iy = fmaf(gy + 1.f, (float)h, -1.f) / 2.f;
                }
                const float fx = floorf(ix), fy = floorf(iy);
                const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
                const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
                const float wne = (ix - (float)x0) * ((float)y1 - iy);
                const float wsw = ((float)x1 - ix) * (iy - (float)y0);
                const float wse = (ix - (float)x0) * (iy - (float)y0);
                if (x0 >= 0 && x0 < w && y0 >= 0 && y0 < h) img[(size_t)y0 * w + x0] += g[j] * wnw;
                if (x1 >= 0 && x1 < w && y0 >= 0 && y0 < h) img[(size_t)y0 * w + x1] += g[j] * wne;
                if (x0 >= 0 && x0 < w && y1 >= 0 && y1 < h) img[(size_t)y1 * w + x0] += g[j] * wsw;
                if (x1 >= 0 && x1 < w && y1 >= 0 && y1 < h) img[(size_t)y1 * w + x1] += g[j] * wse;
            }
        }
}

/* ---- next row: 3D RoI point pooling ------------------------------------------------------------
 * /root/reference/lib/utils/roipool3d/src/roipool3d_kernel.cu: pt_in_box3d (:14-28), assign_pts_to_box3d (:97-120),
 * get_pooled_idx (:123-160: first `sampled` inside points in index order, cyclic repetition k % cnt, empty flag) and the copy
 * (:163-195).  Arithmetic as nvcc -O2 compiles it (PTX-checked): the double-typed literals only widen exactly-representable
 * halves, so float compares are equivalent; x_rot = fma(dx, cosa, -(dz*sina)) and z_rot = fma(dz, cosa, dx*sina) (the PTX shows x_rot as
 * mul, mul, sub without rounding modifiers; ptxas contracts it -- read off the SASS).
 * cosf/sinf come from the C library here and from libdevice on the GPU: they can differ in the last bit, which can only flip
 * a point lying within ~1 ulp of a box face; tests compare this oracle with the reference kernel on the same inputs. */
static int oracle_pt_in_box3d(float x, float y, float z, const float *box)
{
    const float cx = box[0], bottom_y = box[1], cz = box[2], h = box[3], w = box[4], l = box[5], angle = box[6];
    const float half_h = h * 0.5f, half_w = w * 0.5f, half_l = l * 0.5f;
    const float cy = bottom_y - half_h;
    const float dx = x - cx, dz = z - cz;
    if (fabsf(dx) > 10.0f || fabsf(y - cy) > half_h || fabsf(dz) > 10.0f) return 0;
    const float cosa = cosf(angle), sina = sinf(angle);
    const float p1 = dz * sina;
    const float x_rot = fmaf(dx, cosa, -p1);   /* SASS of the reference build: FMUL dz*sina, FFMA dx*cosa - that */
    const float z_rot = fmaf(dz, cosa, dx * sina);
    return (x_rot >= -half_l) & (x_rot <= half_l) & (z_rot >= -half_w) & (z_rot <= half_w);
}

ORACLE_API void oracle_roipool3d(int b, int n, int m, int c, int sampled, const float *xyz, const float *boxes3d, const float *pts_feature,
                                 float *pooled_features, int *pooled_empty_flag)
{
    int *sel = (int *)malloc(sizeof(int) * (size_t)sampled);
    for (int s = 0; s < b; ++s)
        for (int j = 0; j < m; ++j) {
            const float *box = boxes3d + ((size_t)s * m + j) * 7;
            int cnt = 0;
            for (int k = 0; k < n && cnt < sampled; ++k) {
                const float *p = xyz + ((size_t)s * n + k) * 3;
                if (oracle_pt_in_box3d(p[0], p[1], p[2], box)) sel[cnt++] = k;
            }
            if (cnt == 0) {
                pooled_empty_flag[(size_t)s * m + j] = 1;
                continue;
            }
            for (int k = cnt; k < sampled; ++k) sel[k] = sel[k % cnt];
            float *out = pooled_features + ((size_t)s * m + j) * sampled * (3 + c);
            for (int k = 0; k < sampled; ++k) {
                const float *p = xyz + ((size_t)s * n + sel[k]) * 3;
                float *o = out + (size_t)k * (3 + c);
                o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
                if (c > 0) memcpy(o + 3, pts_feature + ((size_t)s * n + sel[k]) * c, sizeof(float) * (size_t)c);
            }
        }
    free(sel);
}

/* ---- next row: rotated BEV overlap / IoU / NMS ---------------------------------------------------
 * /root/reference/lib/utils/iou3d/src/iou3d_kernel.cu: cross (:34-40), check_rect_cross (:42-48), check_in_box2d (:50-65),
 * intersection (:67-97), rotate_around_center (:99-103), point_cmp (:105-107), box_overlap (:109-226), iou_bev (:228-235),
 * iou_normal (:296-304), the mask kernels (:262-293, :307-338) and the host greedy loop of iou3d.cpp:100-113 / :150-163.
 * Arithmetic as nvcc -O2 compiles the reference, read off the SASS (cuobjdump) of the unmodified file -- NOT the PTX: the
 * PTX leaves most `a*b - c*d` as mul, mul, sub without rounding modifiers and ptxas then contracts them.  In all four
 * kernels: a*b - c*d = fma(a, b, -(c*d)) (second product rounded, first fused) for the cross products s1/s3/s4, the
 * intersection numerators, c0/c1/D, the rotated y and the area terms; s2/s5 are the exception (shared products, see below);
 * rotated x = fma(cos, dx, sin*dy); sa + sb = fma(wa, ha, wb*hb) in iou_bev and fma(wb, hb, wa*ha) in iou_normal;
 * fabs(area) / 2.0 is a float multiply by 0.5.  cosf / sinf / atan2f come from the C library here and
 * from libdevice on the GPU and may differ in the last bit, so overlaps agree to ~1e-6 relative rather than bit for bit and
 * a keep-set can differ only where an IoU lies within that distance of the threshold; the bit-exact pin of the product is
 * the reference's own kernels in oracle/_ref (tests/test_iou3d.py). */
typedef struct { float x, y; } o_pt;

/* a*b - c*d as the reference's binary evaluates it: the second product rounded, the first fused into the subtraction */
static float o_pmp(float a, float b, float c, float d)
{
    const float cd = c * d;
    return fmaf(a, b, -cd);
}

static float o_cross3(o_pt p1, o_pt p2, o_pt p0)
{
    return o_pmp(p1.x - p0.x, p2.y - p0.y, p2.x - p0.x, p1.y - p0.y);
}

static int o_intersection(o_pt p1, o_pt p0, o_pt q1, o_pt q0, o_pt *ans)
{
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) && fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) &&
          fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return 0;
    const float s1 = o_cross3(q0, p1, p0), s3 = o_cross3(p0, q1, q0), s4 = o_cross3(q1, p1, q0);
    /* the two products of s2 = cross(p1, q1, p0) are shared with s5 = cross(q1, p1, p0) = -s2, which lives behind a branch:
     * both stay rounded multiplies */
    const float m73 = (p1.x - p0.x) * (q1.y - p0.y), m74 = (q1.x - p0.x) * (p1.y - p0.y);
    const float s2 = m73 - m74;
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = m74 - m73;
    if (fabsf(s5 - s1) > 1e-8f) {
        ans->x = o_pmp(s5, q0.x, s1, q1.x) / (s5 - s1);
        ans->y = o_pmp(s5, q0.y, s1, q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = o_pmp(p0.x, p1.y, p1.x, p0.y);
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = o_pmp(q0.x, q1.y, q1.x, q0.y);
        const float D = o_pmp(a0, b1, a1, b0);
        ans->x = o_pmp(b0, c1, b1, c0) / D;
        ans->y = o_pmp(a1, c0, a0, c1) / D;
    }
    return 1;
}

static o_pt o_rotate(o_pt c, float cosa, float sina, o_pt p)
{
    const float dx = p.x - c.x, dy = p.y - c.y;
    const float t = sina * dy;
    o_pt r;
    r.x = c.x + fmaf(cosa, dx, t);
    r.y = c.y + o_pmp(cosa, dy, sina, dx);
    return r;
}

static int o_in_box2d(const float *box, o_pt p)
{
    o_pt c;
    c.x = (box[0] + box[2]) * 0.5f;
    c.y = (box[1] + box[3]) * 0.5f;
    const o_pt r = o_rotate(c, cosf(-box[4]), sinf(-box[4]), p);
    return r.x > box[0] + -1e-5f && r.x < box[2] + 1e-5f && r.y > box[1] + -1e-5f && r.y < box[3] + 1e-5f;
}

static float o_box_overlap(const float *a, const float *b)
{
    o_pt ca, cb, pa[5], pb[5], pts[32], centre = {0.0f, 0.0f};
    ca.x = (a[0] + a[2]) * 0.5f; ca.y = (a[1] + a[3]) * 0.5f;
    cb.x = (b[0] + b[2]) * 0.5f; cb.y = (b[1] + b[3]) * 0.5f;
    const float xs_a[4] = {a[0], a[2], a[2], a[0]}, ys_a[4] = {a[1], a[1], a[3], a[3]};
    const float xs_b[4] = {b[0], b[2], b[2], b[0]}, ys_b[4] = {b[1], b[1], b[3], b[3]};
    const float cos_a = cosf(a[4]), sin_a = sinf(a[4]), cos_b = cosf(b[4]), sin_b = sinf(b[4]);
    for (int k = 0; k < 4; ++k) {
        o_pt p = {xs_a[k], ys_a[k]}, q = {xs_b[k], ys_b[k]};
        pa[k] = o_rotate(ca, cos_a, sin_a, p);
        pb[k] = o_rotate(cb, cos_b, sin_b, q);
    }
    pa[4] = pa[0];
    pb[4] = pb[0];
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (o_intersection(pa[i + 1], pa[i], pb[j + 1], pb[j], &pts[cnt])) {
                centre.x += pts[cnt].x; centre.y += pts[cnt].y;
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (o_in_box2d(a, pb[k])) { centre.x += pb[k].x; centre.y += pb[k].y; pts[cnt++] = pb[k]; }
        if (o_in_box2d(b, pa[k])) { centre.x += pa[k].x; centre.y += pa[k].y; pts[cnt++] = pa[k]; }
    }
    centre.x /= (float)cnt;
    centre.y /= (float)cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - centre.y, pts[i].x - centre.x) > atan2f(pts[i + 1].y - centre.y, pts[i + 1].x - centre.x)) {
                const o_pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.0f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ux = pts[k].x - pts[0].x, uy = pts[k].y - pts[0].y, vx = pts[k + 1].x - pts[0].x, vy = pts[k + 1].y - pts[0].y;
        area += o_pmp(ux, vy, uy, vx);
    }
    return fabsf(area) * 0.5f;
}

static float o_iou_bev(const float *a, const float *b)
{
    const float sb = (b[2] - b[0]) * (b[3] - b[1]);
    const float sum = fmaf(a[2] - a[0], a[3] - a[1], sb);
    const float ov = o_box_overlap(a, b);
    return ov / fmaxf(sum - ov, 1e-8f);
}

static float o_iou_normal(const float *a, const float *b)
{
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]), top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float w = fmaxf(right - left, 0.0f), h = fmaxf(bottom - top, 0.0f), inter = w * h;
    const float sa = (a[2] - a[0]) * (a[3] - a[1]);
    const float sum = fmaf(b[2] - b[0], b[3] - b[1], sa);
    return inter / fmaxf(sum - inter, 1e-8f);
}

/* mode 0: overlap area, 1: rotated IoU, 2: axis-aligned IoU; out (num_a, num_b) */
ORACLE_API void oracle_boxes_pairwise_bev(int mode, int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < num_a; ++i)
        for (int j = 0; j < num_b; ++j) {
            const float *a = boxes_a + (size_t)i * 5, *b = boxes_b + (size_t)j * 5;
            out[(size_t)i * num_b + j] = mode == 0 ? o_box_overlap(a, b) : mode == 1 ? o_iou_bev(a, b) : o_iou_normal(a, b);
        }
}

/* boxes (n,5) already sorted by descending score; keep (n) receives the kept indices; returns how many.  The suppression
 * bits are the reference's (row i against every later box j, `iou > thresh`), the scan is iou3d.cpp:100-113. */
ORACLE_API int oracle_nms_bev(int rotated, int n, const float *boxes, float thresh, long long *keep)
{
    unsigned char *removed = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        if (removed[i]) continue;
        keep[kept++] = i;
#pragma omp parallel for schedule(static)
        for (int j = i + 1; j < n; ++j) {
            if (removed[j]) continue;
            const float v = rotated ? o_iou_bev(boxes + (size_t)i * 5, boxes + (size_t)j * 5) : o_iou_normal(boxes + (size_t)i * 5, boxes + (size_t)j * 5);
            if (v > thresh) removed[j] = 1;
        }
    }
    free(removed);
    return kept;
}
