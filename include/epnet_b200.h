/*
 * epnet_b200.h -- C ABI of libepnet_b200.so, the B200 (sm_100a) implementation of EPNet's RPN-backbone
 * point ops.  This is the drop-in boundary: every "reference-signature" entry point below takes exactly
 * the arguments of the reference's kernel launcher it replaces (ints, a float radius, raw device
 * pointers, a trailing CUDA stream) so that the reference's five .cpp wrappers -- or any FFI -- can bind
 * it one for one.  Differences from the reference launchers, all deliberate:
 *   - extern "C" (the reference's are C++-mangled);
 *   - they return int: 0 on success, otherwise the cudaError_t of the failed launch or
 *     EPNET_ERR_BAD_ARG; they never call exit() (the reference prints and exit(-1)s,
 *     e.g. sampling_gpu.cu:248-252);
 *   - `stream` is passed as void* (a cudaStream_t) so the header needs no CUDA include.
 * Contract kept from the reference: the caller owns every buffer (outputs, the FPS `temp` scratch,
 * zero-filled gradient buffers, zero-filled ball-query idx); nothing here allocates, frees or
 * synchronises; all tensors are contiguous fp32 / int32; the library is re-entrant.
 *
 * Reference paths are relative to /root/reference/pointnet2_lib/pointnet2/src/.
 */
#ifndef EPNET_B200_H
#define EPNET_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EPNET_OK 0
#define EPNET_ERR_BAD_ARG (-1)

/* ABI version of this header; bumped when a signature changes. */
int epnet_abi_version(void);
/* Human-readable text for a code returned by any entry point. */
const char *epnet_error_string(int code);

/* ---- reference-signature entry points ------------------------------------------------------ */

/* replaces furthest_point_sampling_kernel_launcher (sampling_gpu.h:26, sampling_gpu.cu:211-253).
 * xyz (B,N,3); temp (B,N) in/out running min squared distance, pre-filled by the caller with 1e10
 * (pointnet2_utils.py:26), must be >= 0 and not NaN; idx (B,M) out.  Bit-exact with the reference,
 * including which of several equally distant points wins. */
int epnet_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream);

/* The channel-major gather / group / interpolate entry points below and their gradients may take a scratch table of B*N*C floats from
 * the library's own stream-ordered pool (cudaMallocFromPoolAsync / cudaFreeAsync on `stream`; csrc/transposed_rows.cu) when the
 * problem is large enough for the point-major path; nothing else in this header allocates. */
/* replaces gather_points_kernel_launcher_fast (sampling_gpu.h:12): out[b,c,j] = points[b,c,idx[b,j]] */
int epnet_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, void *stream);
/* replaces gather_points_grad_kernel_launcher_fast (sampling_gpu.h:19); grad_points arrives zeroed */
int epnet_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points, void *stream);

/* replaces ball_query_kernel_launcher_fast (ball_query_gpu.h:12; argument ORDER as in
 * ball_query_gpu.cu:48-49: query centres first, then the cloud).  new_xyz (B,M,3), xyz (B,N,3),
 * idx (B,M,nsample) arrives zeroed. */
int epnet_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, void *stream);

/* The same result (bit-identical idx) through a spatially sorted copy of the cloud, for callers that can provide scratch:
 * epnet_bucket_cloud sorts each scene by Morton key and writes `sorted` (B, npad, 4) = (x, y, z, original index as int bits;
 * padding entries hold index -1) and `boxes` (B, npad/64, 8) = exact min/max corner of every 64 sorted points; npad is a
 * power of two with n <= npad <= 16384.  epnet_ball_query_sorted then answers any number of radius / nsample <= 64 queries
 * against it (ball_query_gpu.cu:23-44 semantics: first nsample indices in ascending order, first-hit padding, zeros kept when
 * the ball is empty), scanning only the buckets whose box meets the ball. */
int epnet_bucket_cloud(int b, int n, int npad, const float *xyz, float *sorted, float *boxes, void *stream);
int epnet_ball_query_sorted(int b, int npad, int m, float radius, int nsample, const float *new_xyz, const float *sorted,
                            const float *boxes, int *idx, void *stream);

/* replaces group_points_kernel_launcher_fast (group_points_gpu.h:13): out[b,c,p,s] = points[b,c,idx[b,p,s]] */
int epnet_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out, void *stream);
/* replaces group_points_grad_kernel_launcher_fast (group_points_gpu.h:19); grad_points arrives zeroed */
int epnet_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points, void *stream);

/* replaces three_nn_kernel_launcher_fast (interpolate_gpu.h:13). unknown (B,n,3), known (B,m,3) ->
 * dist2 (B,n,3) SQUARED distances (the Python layer takes the sqrt, pointnet2_utils.py:98), idx (B,n,3) */
int epnet_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *stream);
/* replaces three_interpolate_kernel_launcher_fast (interpolate_gpu.h:20). points (B,C,m), idx/weight (B,n,3) -> out (B,C,n) */
int epnet_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, void *stream);
/* replaces three_interpolate_grad_kernel_launcher_fast (interpolate_gpu.h:27); grad_points (B,C,m) arrives zeroed */
int epnet_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, void *stream);

/* ---- LI-Fusion gather ---------------------------------------------------------------------- */

/* replaces the torch.nn.functional.grid_sample call of Feature_Gather
 * (/root/reference/lib/net/pointnet2_msg.py:107-120): bilinear, zero padding, one row of N sample
 * points per scene.  fmap (B,C,H,W) NCHW, xy (B,N,2) in [-1,1] (x->W, y->H) -> out (B,C,N). */
int epnet_grid_gather_bilinear(int b, int c, int h, int w, int n, const float *fmap, const float *xy, int align_corners, float *out, void *stream);
/* gradient w.r.t. fmap; grad_fmap (B,C,H,W) arrives zeroed */
int epnet_grid_gather_bilinear_grad(int b, int c, int h, int w, int n, const float *grad_out, const float *xy, int align_corners, float *grad_fmap, void *stream);

/* ---- fused entry points (no reference launcher has these signatures; each names the reference
 *      Python lines whose op-by-op composition it replaces) ----------------------------------- */

/* furthest point sampling that also emits what the set-abstraction level needs next, replacing
 * FPS + transpose + gather_operation + transpose (pointnet2_modules.py:41-45) and, for LI-Fusion, the
 * index widening + torch.gather of the pixel coordinates (lib/net/pointnet2_msg.py:218-219):
 *   idx (B,M); new_xyz (B,M,3) = xyz[idx] (may be NULL); aux_out (B,M,aux_dim) = aux_in (B,N,aux_dim)[idx]
 *   (aux_in/aux_out may be NULL, aux_dim <= 4).  Same sampling, same temp contract as
 *   epnet_furthest_point_sampling. */
int epnet_fps_sample(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, const float *aux_in,
                     float *aux_out, int aux_dim, void *stream);

/* The levels of the backbone sample each other's output (pointnet2_msg.py:141-153 stacks PointnetSAModuleMSG with npoint 4096, 1024,
 * 256, 64; each calls furthest_point_sample on the previous level's new_xyz, pointnet2_modules.py:38-42), and a cloud that is already
 * in furthest-point order samples to the identity unless two of its points tie for a maximum.
 *   epnet_fps_prefix_check: flag[s] = 1 iff sampling m of the n points of xyz[s] from temp = 1e10 returns exactly 0..m-1 with every
 *     arg-max unique (decided in O(n*m) parallel work); one check for (n, m) also holds for every (n' <= n, m' <= m) prefix.
 *     winners (B,m) scratch; flag (B) int; m <= 2048, larger m clears the flags.
 *   epnet_fps_sample_guarded: epnet_fps_sample, except that scenes with identity[s] != 0 get idx = 0..m-1, new_xyz / aux_out = the
 *     first m rows of xyz / aux_in (their temp is left untouched); the others run the real sampling.  Bit-exact either way. */
int epnet_fps_prefix_check(int b, int n, int m, const float *xyz, float *winners, int *flag, void *stream);
int epnet_fps_sample_guarded(int b, int n, int m, const float *xyz, float *temp, int *idx, float *new_xyz, const float *aux_in,
                             float *aux_out, int aux_dim, const int *identity, void *stream);

/* One scale of a set-abstraction level without input features, in one launch: QueryAndGroup (re-centred coordinates only) +
 * SharedMLP (three 1x1 convolutions, BatchNorm(eval) folded, ReLU) + max over nsample
 * (pointnet2_modules.py:44-58, pointnet2_utils.py:241-264, pytorch_utils.py:20-32).  pack = [W1 (n1 rows of wx, wy, wz, bias) |
 * W2 transposed (n1 x n2) | b2 | W3 (n3 x n2) | b3] floats, 16-byte aligned; out (B*m, n3) with row stride ldo.  Instantiated for
 * (n1, n2, n3, nsample) = (16,16,32,16) and (32,32,64,32); anything else returns EPNET_ERR_BAD_ARG (use the GEMM entry points). */
int epnet_sa_first_level(int b, int n, int m, int nsample, int n1, int n2, int n3, const float *xyz, const float *new_xyz,
                         const int *idx, const float *pack, float *out, int ldo, void *stream);

/* QueryAndGroup.forward minus the ball query (pointnet2_utils.py:250-257) in one launch:
 *   out (B,3+C,M,ns): rows 0..2 = xyz[idx] - new_xyz (re-centred coordinates), rows 3.. = features[:, idx].
 * xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or NULL with c = 0, idx (B,M,ns). */
int epnet_group_concat(int b, int c, int n, int m, int nsample, const float *xyz, const float *new_xyz, const float *features,
                       const int *idx, float *out, void *stream);

/* in place x[b,c,:] = max(x[b,c,:] + bias[c], 0) over x (B,C,L): the BatchNorm(eval)+ReLU tail of a
 * SharedMLP layer (pytorch_utils.py:20-32) once BN is folded into the 1x1 convolution. */
int epnet_bias_relu(int b, int c, long long l, float *x, const float *bias, void *stream);

/* out[b,c,p] = max_s max(x[b,c,p,s] + bias[c], 0): last SharedMLP layer's tail fused with
 * F.max_pool2d(kernel=[1,nsample]) + squeeze (pointnet2_modules.py:59-68).  x (B,C,M,ns) -> out with
 * batch stride out_batch_stride floats (so it can write straight into the MSG concat, :72). */
int epnet_bias_relu_maxpool(int b, int c, int m, int nsample, const float *x, const float *bias, float *out,
                            long long out_batch_stride, void *stream);

/* PointnetFPModule.forward between three_nn and the MLP (pointnet2_modules.py:157-166) in one launch:
 * inverse-distance weights from the SQUARED distances (sqrt, 1/(d+1e-8), normalise), three_interpolate of
 * known_feats (B,C2,m), and the concat with the skip features (B,C1,n) (NULL with c1 = 0):
 *   out (B,C2+C1,n). */
int epnet_three_interpolate_concat(int b, int c2, int m, int n, int c1, const float *known_feats, const int *idx,
                                   const float *dist2, const float *skip_feats, float *out, void *stream);

/* One SharedMLP layer as a tcgen05/TMEM GEMM with fp32-grade accuracy (3xTF32 split), point-major operands:
 *   y[l / pool][n] = max over pool consecutive rows l of act(sum_k x[l][k] * W[n][k] + bias[n])
 * x (L, ldx) fp32 rows; wpack = W (N,K) split into TF32 hi/lo parts and pre-swizzled by the host into n-tiles of BN
 * rows x k-blocks of 32 (epnet_b200/gemm.py:pack_weights); bias (N) or NULL; relu 0/1; pool in {1,2,4,8,16,32} (the
 * nsample max-pool of pointnet2_modules.py:59-61 fused into the epilogue); y (L / pool, ldy).
 * Replaces the Conv2d(1x1)+BatchNorm(eval)+ReLU units of pytorch_utils.py:20-32 (BN folded into W and bias). */
int epnet_gemm_tf32x3(int L, int K, int N, const float *x, int ldx, const float *wpack, int BN, const float *bias, int relu,
                      int pool, float *y, int ldy, void *stream);

/* Point-major ("pm": rows = points, channels contiguous) variants of the fused kernels; they produce the K-major A
 * operand of epnet_gemm_tf32x3 directly and read each gathered point as one contiguous row.
 * group_concat_pm: out[(b,p,s)] = [ feats[b, idx[b,p,s], 0..C) | xyz[b, idx] - new_xyz[b,p] | 0-pad ], row stride ldo
 *   (QueryAndGroup.forward, pointnet2_utils.py:250-257, with the xyz channels AFTER the features).
 * three_interpolate_concat_pm: out[(b,i)] = [ interpolated known feats (C2) | skip feats (C1) ] (pointnet2_modules.py:157-166);
 *   `weight` holds the normalised weights, or the squared distances when from_dist2 != 0.
 * grid_gather_pm: out[(b,i)][0..C) = bilinear(fmap[b,:,.,.], xy[b,i]) (Feature_Gather, lib/net/pointnet2_msg.py:107-120). */
int epnet_group_concat_pm(int b, int c, int n, int m, int nsample, const float *xyz, const float *new_xyz, const float *feats, int ldf,
                          const int *idx, float *out, int ldo, void *stream);
int epnet_three_interpolate_concat_pm(int b, int c2, int m, int n, int c1, const float *known, int ldk, const int *idx,
                                      const float *weight, int from_dist2, const float *skip, int lds, float *out, int ldo,
                                      void *stream);
/* three_nn (same search, same outputs as epnet_three_nn) that also writes weight (B,n,3): the normalised inverse-distance
 * weights of pointnet2_modules.py:157-159, consumed by epnet_three_interpolate_concat_pm with from_dist2 = 0. */
int epnet_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, float *weight,
                           void *stream);
int epnet_grid_gather_pm(int b, int c, int h, int w, int n, const float *fmap, const float *xy, int align_corners, float *out, int ldo,
                         void *stream);

/* The three wide-tile entry points (BN > 64) with the operands split into two FP16 terms instead of TF32 hi/lo:
 * x = h1 + 2^-11 h2 carries the same 22 significand bits, the MMAs run at twice the TF32 rate and every shared-memory byte holds
 * twice as many k-values.  wpack = FP16 planes in k-blocks of 64 (epnet_b200/gemm.py).  |x| and |w| must stay below 65504. */
int epnet_gemm_f16x3(int L, int K, int N, const float *x, int ldx, const float *wpack, int BN, const float *bias, int relu, int pool,
                     float *y, int ldy, void *stream);
int epnet_conv3x3_nhwc_f16x3(int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                             const float *bias, int relu, float *y, int ldy, void *stream);
int epnet_deconv_nhwc_f16x3(int b, int h, int w, int cin, int k, int co, const float *x, int ldx, const float *wpack, int BN,
                            const float *bias, int relu, float *out, int ldo, void *stream);

/* FP16-split layers whose A operand arrives ALREADY SPLIT into two FP16 planes (x = h1 + 2^-11 h2, written by the producing layer's
 * epilogue through the yh1/yh2 outputs below) and is moved by TMA tensor loads (cp.async.bulk.tensor) -- no conversion work, no
 * per-thread addressing; image borders and ragged tiles are the TMA unit's zero fill, stride 2 its traversal stride.  Same arithmetic
 * as the *_f16x3 entry points (same split, same MMA sequence).  Replace the same reference layers: the image stream's 3x3
 * convolutions (/root/reference/lib/net/pointnet2_msg.py:17-33), Conv2d(1x1)+BN+ReLU units (pytorch_utils.py:20-32), transposed
 * convolutions (pointnet2_msg.py:163-168).  Planes: FP16, 16-byte aligned, row/pixel stride (ldx, ldh) a multiple of 8 halfs.
 * Outputs: y fp32 and/or planes yh1/yh2 (either may be NULL, not both). */
int epnet_conv3x3_planes_tma(int b, int h, int w, int cin, int cout, int stride, const void *xh1, const void *xh2, int ldx,
                             const float *wpack, int BN, const float *bias, int relu, float *y, int ldy, void *yh1, void *yh2, int ldh,
                             void *stream);
int epnet_gemm_planes_tma(int L, int K, int N, const void *xh1, const void *xh2, int ldx, const float *wpack, int BN, const float *bias,
                          int relu, int pool, float *y, int ldy, void *yh1, void *yh2, int ldh, void *stream);
int epnet_deconv_planes_tma(int b, int h, int w, int cin, int k, int co, const void *xh1, const void *xh2, int ldx, const float *wpack,
                            int BN, const float *bias, int relu, float *out, int ldo, void *stream);
/* epnet_conv3x3_nhwc_tf32x3 (fp32 NHWC input, TF32 split) that also writes its result as FP16 planes for a following *_planes_tma layer;
 * y may be NULL. */
int epnet_conv3x3_nhwc_tf32x3_planes(int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                                     const float *bias, int relu, float *y, int ldy, void *yh1, void *yh2, int ldh, void *stream);

/* ---- sparse evaluation of the final image fusion ------------------------------------------------------
 * The reference up-samples the four image-stream maps with ConvTranspose2d (kernel == stride = 2, 4, 8, 16), concatenates them,
 * applies a 1x1 conv + BN + ReLU to ALL 384 x 1280 pixels and bilinearly samples the result at the points
 * (/root/reference/lib/net/pointnet2_msg.py:237-246 with Feature_Gather :107-120).  These entry points compute the fused image feature
 * only at the <= 4 taps of every point: taps are counting-sorted by transposed-convolution phase (Y % 16, X % 16), every level is one
 * row-gather GEMM whose 128-row tiles select their weight slice by phase, the 1x1 conv runs on the sorted rows, and a blend kernel
 * applies the bilinear weights (same taps, weights and fma chain as epnet_grid_gather_nhwc_pm).  See epnet_b200/sparse_tail.py.
 * epnet_tail_taps: xy (b,n,2) in [-1,1] -> tap_pix (b*n*4) = scene*H*W + y*W + x, tap_w (b*n*4), hist (256) += taps per phase.
 * epnet_tail_plan: hist -> start (256; bins padded to 128 rows), cursor (256) = 0, tile_phase (max_tiles), *n_tiles.
 * epnet_tail_scatter: pos_of_slot (taps) and, per level i, row_idx[i*row_stride + sorted row] = row of the level's input map.
 * epnet_tail_blend: out (points, ldo)[0..c) = sum_t tap_w * F[pos_of_slot].
 * epnet_gemm_tf32x3_rows: narrow-tile GEMM over rows gathered by row_idx, weight n-tile per m-tile from tile_phase, tile count from
 * the device (any of the three may be NULL). */
int epnet_tail_taps(int b, int n, int H, int W, int align_corners, const float *xy, int *tap_pix, float *tap_w, int *hist, void *stream);
int epnet_tail_plan(const int *hist, int *start, int *cursor, int *tile_phase, int *n_tiles, int max_tiles, void *stream);
int epnet_tail_scatter(int slots, int H, int W, int levels, const int *k, const int *h, const int *w, const int *tap_pix, const int *start,
                       int *cursor, int *pos_of_slot, int *row_idx, int row_stride, void *stream);
int epnet_tail_blend(int points, int c, const float *F, int ldf, const int *pos_of_slot, const float *tap_w, float *out, int ldo,
                     void *stream);
int epnet_gemm_tf32x3_rows(int L, int K, int N, const float *x, int ldx, const int *row_idx, const int *tile_phase, int phase_k,
                           const int *m_tiles_dev, const float *wpack, int BN, const float *bias, int relu, float *y, int ldy,
                           void *stream);

/* The FIRST convolution of the image stream (Img_Block[0].conv1 + folded BatchNorm + ReLU, /root/reference/lib/net/pointnet2_msg.py:17-24: 3 -> 64
 * channels, 3x3, stride 1, pad 1) as a dedicated fp32 FFMA kernel: K = 27 is too short for the tensor-core tile.  x (b,H,W,4) fp32 NHWC, channel 3
 * ignored; w (64,3,3,4) = (o,ky,kx,c); bias (64) or NULL; outputs y (b*H*W, ldy) fp32 and/or FP16 planes yh1/yh2 (b*H*W, ldh) for a following
 * *_planes_tma layer (either may be NULL, not both).  cout must be 64, W a multiple of 4. */
int epnet_conv3x3_c3_planes(int b, int H, int W, int cout, const float *x, const float *w, const float *bias, int relu, float *y, int ldy,
                            void *yh1, void *yh2, int ldh, void *stream);

/* Range guard of the FP16 operand split (no reference counterpart: pytorch_utils.py:20-32 computes in fp32).  Every GEMM
 * epilogue of this library raises a per-device flag when it writes a magnitude above 6e4 or a non-finite value -- i.e. whenever
 * a following FP16-split layer could overflow.  _read copies the flag (0/1) to host memory asynchronously on `stream` (pinned
 * memory for a truly asynchronous copy), _reset clears it on `stream`.  The caller re-runs on the TF32 split when it is set. */
int epnet_gemm_overflow_read(unsigned int *host_dst, void *stream);
int epnet_gemm_overflow_reset(void *stream);

/* Image input preparation on the device.  Replaces the reference's HOST-side float64 preparation
 * (/root/reference/lib/datasets/kitti_dataset.py:37-57: uint8 RGB -> /255 -> -mean -> /std -> zero-padded (384,1280,3) canvas) and its
 * upload + conversion (/root/reference/lib/net/train_functions.py:37: .cuda().float().permute(0,3,1,2)).
 * src: b decoded images, uint8 RGB interleaved; scene s starts at src + s*scene_stride, row y at + y*pitch (bytes); h_in x w_in pixels
 * are allocated per scene; sizes (b,2) int32 {rows, columns} decoded per scene (device), or NULL = all h_in x w_in.  mean, std: 3
 * doubles each (host).  Outputs, either may be NULL: nhwc4 (b,H,W,4) fp32, channel 3 = 0 (operand of the first convolution);
 * nchw (b,3,H,W) fp32 (the reference's `img` tensor).  Values are bit-identical to the reference's (float64 arithmetic, one
 * rounding to fp32); pixels outside the decoded image are exactly 0. */
int epnet_image_prep_u8(int b, int h_in, int w_in, long long pitch, long long scene_stride, const unsigned char *src, const int *sizes,
                        int H, int W, const double *mean, const double *std, float *nhwc4, float *nchw, void *stream);
/* (b,3,H,W) fp32 -> (b,H,W,4) fp32, channel 3 = 0: the layout change in front of the first convolution, one pass. */
int epnet_image_nchw_to_nhwc4(int b, int H, int W, const float *src, float *dst, void *stream);

/* First shared-MLP layer of a set-abstraction scale with QueryAndGroup fused into the operand load (pointnet2_utils.py:241-264 +
 * pointnet2_modules.py:47-52): GEMM row (scene, centre p, sample s) = [feats[scene, idx[scene,p,s], 0..c) | xyz[scene, idx] -
 * new_xyz[scene, p]] is gathered straight into the tensor-core operand, the grouped tensor is never written.  feats point-major
 * (scenes, n, ldf) or NULL when c == 0; wpack packs W (N, c+3) with the three offset columns LAST; BN <= 64. */
int epnet_gemm_tf32x3_grouped(int scenes, int n, int m, int nsample, int c, const float *feats, int ldf, const float *xyz,
                              const float *new_xyz, const int *idx, const float *wpack, int BN, int N, const float *bias, int relu,
                              int pool, float *y, int ldy, void *stream);

/* epnet_gemm_tf32x3 with the result written channel-major: x rows are (scene, point) pairs, `pts` points per scene, y is
 * (L/pts, N, pts), i.e. the (B, C, N) feature layout of the reference's modules (pointnet2_modules.py:72). */
int epnet_gemm_tf32x3_cm(int L, int K, int N, int pts, const float *x, int ldx, const float *wpack, int BN, const float *bias, int relu,
                         float *y, void *stream);

/* 3x3 convolution (padding 1, stride 1 or 2) as an implicit tcgen05 3xTF32 GEMM on an NHWC image -- the image-stream
 * convolutions of BasicBlock (lib/net/pointnet2_msg.py:17-33; cuDNN fp32 there).  x (B,H,W,Cin), Cin a power of two >= 4;
 * wpack = the weight reordered to (Cout, ky, kx, Cin) and packed like epnet_gemm_tf32x3's; y (B*Ho*Wo, ldy) NHWC. */
int epnet_conv3x3_nhwc_tf32x3(int b, int h, int w, int cin, int cout, int stride, const float *x, const float *wpack, int BN,
                              const float *bias, int relu, float *y, int ldy, void *stream);

/* NHWC image-stream helpers.  grid_gather_nhwc_pm: Feature_Gather (lib/net/pointnet2_msg.py:107-120) from an NHWC map
 * (B,H,W,ldc) -> out[(b,i)][0..C).  deconv_shuffle_nhwc: the pixel shuffle of a ConvTranspose2d with kernel == stride
 * (pointnet2_msg.py:163-165,239-241) computed as a GEMM y (B*h*w, k*k*co) -> out (B, h*k, w*k, ldo)[..., col_off:col_off+co],
 * i.e. straight into the channel concat of :242. */
/* LI-Fusion attention tail (IA_Layer.forward, lib/net/pointnet2_msg.py:79-96) on point-major rows in one pass:
 * out[row] = x[row] * sigmoid(w3 . tanh(r1[row] + r2[row]) + b3), r1/r2 (rows, rc), x/out (rows, c); all 16-byte aligned. */
int epnet_attention_scale_pm(int rows, int rc, int c, const float *r1, int ld1, const float *r2, int ld2, const float *w3,
                             const float *b3, const float *x, int ldx, float *out, int ldo, void *stream);
int epnet_grid_gather_nhwc_pm(int b, int c, int h, int w, int n, const float *fmap, int ldc, const float *xy, int align_corners,
                              float *out, int ldo, void *stream);
int epnet_deconv_shuffle_nhwc(int b, int h, int w, int k, int co, const float *y, float *out, int ldo, int col_off, void *stream);

/* ConvTranspose2d with kernel == stride == k (lib/net/pointnet2_msg.py:170-172 builds them that way) on an NHWC map as one
 * tcgen05 GEMM whose epilogue scatters every input pixel's k x k x co patch into the NHWC output (pixels ldo floats apart, `out`
 * pre-offset to this map's channel slice of the concatenation at pointnet2_msg.py:242).  x: B*h*w rows ldx apart, cin read;
 * wpack: rows ordered (ky, kx, o); bias: k*k*co values or NULL. */
int epnet_deconv_nhwc_tf32x3(int b, int h, int w, int cin, int k, int co, const float *x, int ldx, const float *wpack, int BN,
                             const float *bias, int relu, float *out, int ldo, void *stream);

/* ---- next row of the scope table: 3D RoI point pooling ------------------------------------------------------------ */

/* replaces roipool3dLauncher (/root/reference/lib/utils/roipool3d/src/roipool3d.cpp:8-9, roipool3d_kernel.cu:207-236), same
 * arguments plus the stream.  xyz (B,N,3), boxes3d (B,M,7) [x, bottom y, z, h, w, l, ry], pts_feature (B,N,C) ->
 * pooled_features (B,M,sampled,3+C) and pooled_empty_flag (B,M), both arriving zeroed (roipool3d_utils.py:20-22);
 * sampled <= 512.  Unlike the reference it allocates nothing and does not synchronise. */
int epnet_roipool3d(int b, int n, int m, int c, int sampled, const float *xyz, const float *boxes3d, const float *pts_feature,
                    float *pooled_features, int *pooled_empty_flag, void *stream);

/* ---- next row of the scope table: rotated BEV overlap / IoU and NMS -------------------------- */

/* replaces boxesoverlapLauncher (/root/reference/lib/utils/iou3d/src/iou3d_kernel.cu:352-361, called from iou3d.cpp:34-52).
 * boxes_a (num_a,5), boxes_b (num_b,5) as [x1, y1, x2, y2, ry] -> ans_overlap (num_a,num_b): area of the intersection of
 * the two rotated rectangles.  Same floating-point expression as the reference (bit-identical values). */
int epnet_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_overlap, void *stream);

/* replaces boxesioubevLauncher (iou3d_kernel.cu:363-369, iou3d.cpp:54-72): ans_iou (num_a,num_b) = overlap / max(sa + sb - overlap, 1e-8). */
int epnet_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou, void *stream);

/* bytes of scratch the two NMS entry points need for `s` problems of up to `n` boxes: s * n * ceil(n/64) 64-bit words
 * (the reference cudaMallocs and frees this per call: iou3d.cpp:86-88,98). */
int epnet_nms_workspace_bytes(int s, int n, unsigned long long *bytes);

/* replaces nmsLauncher (iou3d_kernel.cu:372-377) PLUS the host half of nms_gpu (iou3d.cpp:74-121: D2H copy of the mask and
 * the serial greedy loop).  `s` independent problems in one launch: boxes (s,n,5), each problem sorted by descending score by
 * the caller (iou3d_utils.py:62-64); counts (s) = boxes actually present per problem, or NULL for n everywhere;
 * a box is suppressed by an earlier kept box when their rotated BEV IoU > thresh.  keep (s,n) int64: indices of the kept
 * boxes in ascending order (first num_out[p] entries valid, the rest untouched); num_out (s).  max_out > 0 stops each problem
 * after that many kept boxes (what the proposal layer truncates to, lib/rpn/proposal_layer.py:111); <= 0 = no limit.
 * workspace: epnet_nms_workspace_bytes(s, n) bytes, 8-byte aligned, contents irrelevant.  Keep-sets are identical to the reference's. */
int epnet_nms_rotated(int s, int n, const float *boxes, const int *counts, float thresh, int max_out, void *workspace, long long *keep,
                      int *num_out, void *stream);

/* replaces nmsNormalLauncher (iou3d_kernel.cu:380-385) plus the host half of nms_normal_gpu (iou3d.cpp:124-170): the same
 * with the axis-aligned IoU of the [x1,y1,x2,y2] extents (the angle is ignored). */
int epnet_nms_normal(int s, int n, const float *boxes, const int *counts, float thresh, int max_out, void *workspace, long long *keep,
                     int *num_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EPNET_B200_H */
