"""Inference runner for the RPN backbone: the same network as Pointnet2MSG.forward (the mirror of
/root/reference/lib/net/pointnet2_msg.py:201-248), scheduled B200-first.

What changes relative to the op-by-op module path (nothing changes in WHAT is computed):
  * three streams inside ONE captured CUDA graph: the furthest-point-sampling chain of all four levels (it
    depends on coordinates only), the image stream, and the point stream; the ~300 tiny
    launches of a forward become one graph replay with parallel branches;
  * eval-mode BatchNorm is folded into the preceding convolution / 1x1 conv (w' = w*g/sqrt(v+eps),
    b' = beta - mean*g/sqrt(v+eps));
  * fused kernels from libepnet_b200.so: FPS emits new_xyz and the gathered pixel coordinates itself;
    ball-query indices feed one group+recentre+concat launch that writes the GEMM operand; the last MLP
    layer's bias+ReLU is fused with the max-pool and writes into the multi-scale concat; three_nn distances
    feed one weights+interpolate+concat launch;
  * every shared-MLP layer, 3x3 convolution, transposed convolution and fusion layer runs on the hand-written tcgen05 GEMM
    kernels (csrc/gemm_tf32x3.cu) over point-major / NHWC activations -- no cuBLAS or cuDNN call anywhere in the runner.
Only eval mode is supported here (train mode needs batch statistics: use the module path).
"""
import torch

from . import image_prep
from . import pointnet2_cuda as pc
from .sparse_tail import SparseImageTail
from .gemm import FusedFirstLevel, OverflowFlag, PackedConv3x3, PackedDeconv, PackedLinear, f16_split, grouped_first_layer, tile_policy


class _nvtx:
    """NVTX range around a level of the schedule (the reference has no profiler hooks).  Host-side markers:
    they label the eager pass and the capture under Nsight Systems / Compute (`--nvtx --nvtx-include "SA1/"`), cost nothing in a replay."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


def _fold_bn(weight2d, conv_bias, bn):
    """conv(1x1) [+bias] followed by eval-mode BatchNorm -> (W', b')."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = weight2d * scale[:, None]
    b = bn.bias - bn.running_mean * scale
    if conv_bias is not None:
        b = b + conv_bias * scale
    return w.contiguous(), b.contiguous()


def _fold_shared_mlp(mlp):
    layers = []
    for unit in mlp:  # layer0, layer1, ...: Sequential(conv, bn(bn), activation)
        conv = unit.conv
        w2d = conv.weight.reshape(conv.out_channels, conv.in_channels)
        if hasattr(unit, "bn"):
            layers.append(_fold_bn(w2d, conv.bias, unit.bn.bn))
        else:
            bias = conv.bias if conv.bias is not None else torch.zeros(conv.out_channels, device=w2d.device)
            layers.append((w2d.contiguous(), bias.contiguous()))
    return layers


class _FusionPM:
    """Atten_Fusion_Conv / Fusion_Conv on point-major rows with tcgen05 GEMMs.  `cat` is a (rows, 2*Cp) buffer whose left
    half already holds the point features (written there by the producing GEMM) -- the concat costs nothing."""

    def __init__(self, mod):
        self.attention = hasattr(mod, "IA_Layer")
        if self.attention:
            ia = mod.IA_Layer
            self.fc1 = PackedLinear(ia.fc1.weight, ia.fc1.bias + ia.fc2.bias)  # the two biases are summed before tanh
            self.fc2 = PackedLinear(ia.fc2.weight, None)
            self.w3, self.b3 = ia.fc3.weight.detach().reshape(-1).contiguous(), ia.fc3.bias.detach().reshape(1).contiguous()
            conv, bn = ia.conv1[0], ia.conv1[1]
            self.conv = PackedLinear(*_fold_bn(conv.weight.squeeze(-1), conv.bias, bn))
        self.fuse = PackedLinear(*_fold_bn(mod.conv1.weight.squeeze(-1), mod.conv1.bias, mod.bn1))
        self.cp = mod.conv1.out_channels
        self.cat_width = mod.conv1.in_channels

    def __call__(self, cat, img, out=None, out_cm=None):
        """cat (rows, Cp + Ci') with point features in [:, :Cp]; img (rows, Ci) gathered image features -> (rows, Cp), or
        channel-major into out_cm (B, Cp, pts)"""
        cp = self.cp
        if self.attention:
            point = cat[:, :cp]
            # att = sigmoid(fc3(tanh(fc1(img) + fc2(point)))) and the scaling of conv1(img) in one pass over the rows
            pc.attention_scale_pm_wrapper(self.fc1(img, relu=False), self.fc2(point, relu=False), self.w3, self.b3,
                                          self.conv(img, relu=True), cat[:, cp:])
        else:
            cat[:, cp:].copy_(img)
        return self.fuse(cat, relu=True, out=out, out_cm=out_cm)

class BackboneRunner:
    def __init__(self, model, batch, npoints, device, image_hw=(384, 1280), use_graph=True, tiles="latency", f16=True, sparse_tail=True,
                 prefix_fps=True, fused_first_level=True):
        """Point-major activations + tcgen05 fp32-grade GEMMs.
        f16: the wide GEMM tiles split operands into two FP16 terms (gemm.F16_WIDE), which needs activations and folded weights
        below 65504 in magnitude (they are below 10 for the published configuration).  Guarded twice: a layer whose folded
        weights leave the range is packed for the TF32 split (PackedLinear.f16_ok), and every GEMM epilogue raises a device flag
        on |y| > 6e4 / non-finite values, read back with every call (`overflowed()`); Pointnet2MSG.forward and PipelinedRunner
        callers then switch to f16=False (TF32 split everywhere: fp32's range).  EPNET_F16_WIDE=0 forces that from the start.
        tiles: gemm.tile_policy for every GEMM launch of this runner ("latency" for one batch at a time)."""
        self.tiles = tiles
        self.f16 = bool(f16)
        # sparse_tail: evaluate the final image fusion (transposed convolutions + 1x1 conv, pointnet2_msg.py:237-243) only at the
        # <= 4 taps of every point (sparse_tail.py) instead of over the whole up-sampled canvas; False keeps the dense form
        self.want_sparse_tail = bool(sparse_tail)
        # prefix_fps: levels 2.. sample the previous level's output, which is in furthest-point order: one parallel check after level 1
        # (csrc/fps.cu: epnet_fps_prefix_check) proves that their answer is the identity prefix -- or, on a tie, lets the sampling
        # kernels run; bit-exact either way.  False always runs the sampling kernels.
        self.prefix_fps = bool(prefix_fps)
        # fused_first_level: a set-abstraction scale without input features (level 1 of the published configuration) runs as ONE
        # kernel -- group, three-layer MLP in registers, max over the ball (csrc/sa_first_level.cu) -- instead of three tcgen05 GEMM
        # launches whose tiles are all fixed cost at K <= 32; False keeps the GEMM path
        self.fused_first_level = bool(fused_first_level)
        self._fused_sa = {}
        self.image_hw = tuple(image_hw)
        self.overflow = OverflowFlag(device)
        self._last_stream = None
        if model.training:
            raise RuntimeError("BackboneRunner folds BatchNorm: call model.eval() first (train mode: use model(...) itself)")
        c = model.config
        if c.input_channels != 0:
            raise NotImplementedError("runner covers the published config (xyz-only input, RPN.USE_INTENSITY False)")
        self.model, self.cfg, self.device = model, c, device
        self.B, self.N = batch, npoints
        H, W = image_hw
        f32 = dict(dtype=torch.float32, device=device)
        self.points = torch.zeros(batch, npoints, 3, **f32)
        # the graph's image input: the NHWC canvas, channels padded to 4, that the first convolution reads.  It is produced in
        # front of the replay by one kernel from whatever the caller holds (image_prep.py): the decoded uint8 image, or the
        # reference's fp32 (B,3,H,W) tensor.  Staging buffers for host inputs are allocated on first use.
        self.image4 = torch.zeros(batch, H, W, 4, **f32)
        self._stage_f32 = None
        self._stage_u8 = None
        self.xy = torch.zeros(batch, npoints, 2, **f32)
        self._xy_scale = torch.tensor([c.image_size[0] - 1.0, c.image_size[1] - 1.0], **f32)

        with torch.no_grad():
            self.sa = []
            for k, sa in enumerate(model.SA_modules):
                scales = []
                for grouper, mlp in zip(sa.groupers, sa.mlps):
                    scales.append((float(grouper.radius), int(grouper.nsample), _fold_shared_mlp(mlp)))
                self.sa.append((int(sa.npoint), scales))
            self.fp = [_fold_shared_mlp(fp.mlp) for fp in model.FP_modules]
            if c.li_fusion:
                self.img_blocks = []
                for blk in model.Img_Block:
                    w1, b1 = _fold_bn(blk.conv1.weight.flatten(1), None, blk.bn1)
                    self.img_blocks.append((w1.view_as(blk.conv1.weight).contiguous(), b1, blk.conv1.stride, blk.conv2))

            self._build_pm()

        # the FPS chain is the critical path of a batch and needs a whole SM's shared memory per scene: high priority, so that its CTAs
        # are placed as soon as an SM drains instead of queueing behind the image stream's grids (one batch at a time: 2.98 -> 2.90 ms;
        # throughput with 8 batches in flight unchanged)
        self.s_fps = torch.cuda.Stream(device=device, priority=-1)
        self.s_img = torch.cuda.Stream(device=device)
        self.s_scale = [torch.cuda.Stream(device=device)]  # second grouping scale of a set-abstraction level
        self.s_geo = [torch.cuda.Stream(device=device) for _ in range(3)]  # geometry-only work: ball queries (2 scales), three_nn
        self.kernel_launches_per_replay = 0
        self.fps_identity = None  # (B,) int32 after a call: 1 where levels 2.. were answered by the prefix
        self.graph = None
        self.out = None
        if use_graph:
            self._capture()

    # ------------------------------------------------------------------------------------ pieces
    def _forward(self):
        with tile_policy(self.tiles), f16_split(self.f16):
            return self._forward_impl()

    def _forward_impl(self):
        c, B, N, dev = self.cfg, self.B, self.N, self.device
        main = torch.cuda.current_stream(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        xyz0 = self.points
        xyn = self.xy / self._xy_scale * 2.0 - 1.0 if c.li_fusion else None  # pointnet2_msg.py:208-210

        # ---- FPS chain of all levels (coordinates only) on its own stream ----
        ready = torch.cuda.Event()
        ready.record(main)
        self._ready = ready
        l_xyz, l_xy, fps_done = [xyz0], [xyn], []
        with torch.cuda.stream(self.s_fps), _nvtx("fps_chain"):
            self.s_fps.wait_event(ready)
            cur_xyz, cur_xy = xyz0, xyn
            npoints = [npoint for npoint, _ in self.sa]
            nested = all(a >= b for a, b in zip([N] + npoints, npoints))
            identity = None
            for li, (npoint, _) in enumerate(self.sa):
                n = cur_xyz.shape[1]
                temp = torch.full((B, n), 1e10, **f32)
                idx = torch.empty((B, npoint), dtype=torch.int32, device=dev)
                new_xyz = torch.empty((B, npoint, 3), **f32)
                new_xy = torch.empty((B, npoint, 2), **f32) if c.li_fusion else None
                if identity is None:
                    pc.fps_sample_wrapper(B, n, npoint, cur_xyz, temp, idx, new_xyz, cur_xy, new_xy)
                else:
                    pc.fps_sample_guarded_wrapper(B, n, npoint, cur_xyz, temp, idx, identity, new_xyz, cur_xy, new_xy)
                if li == 0 and self.prefix_fps and nested and len(npoints) > 1 and npoints[1] <= 2048:
                    # the check of level 2 (n = npoints[0], m = npoints[1]) contains the conditions of every later level
                    identity = torch.empty((B,), dtype=torch.int32, device=dev)
                    pc.fps_prefix_check_wrapper(B, npoint, npoints[1], new_xyz, torch.empty((B, npoints[1]), **f32), identity)
                    self.fps_identity = identity
                ev = torch.cuda.Event()
                ev.record(self.s_fps)
                fps_done.append(ev)
                l_xyz.append(new_xyz)
                l_xy.append(new_xy)
                cur_xyz, cur_xy = new_xyz, new_xy

        # ---- image stream ----
        imgs, img_done, img_fusion, img_fusion_done = [], [], None, None
        if c.li_fusion:
            # NHWC activations; 3x3 convolutions, transposed convolutions and the 1x1 fusion conv on the tcgen05 3xTF32 GEMM
            with torch.cuda.stream(self.s_img), _nvtx("image_stream"):
                self.s_img.wait_event(ready)
                H, W = self.image4.shape[1], self.image4.shape[2]
                x = self.image4
                def run_conv(conv, inp, relu):
                    # the next convolution wants a power-of-two channel count: pad the buffer (zeros) when Cout is not one
                    cp = 4
                    while cp < conv.cout:
                        cp *= 2
                    if cp == conv.cout:
                        return conv(inp, relu=relu)
                    ho, wo = (inp.shape[1] - 1) // conv.stride + 1, (inp.shape[2] - 1) // conv.stride + 1
                    buf = torch.zeros((B, ho, wo, cp), **f32)
                    conv(inp, relu=relu, out=buf[..., :conv.cout])
                    return buf

                img_channels = []
                # Planes chain: every convolution after the first reads its input as the two FP16 planes the previous one's epilogue
                # wrote, through TMA tensor loads (gemm.Planes, csrc gemm_f16x3_tma_kernel): each activation is split once, by its
                # producer, instead of once per tap and column tile by the consumers' SIMT producer warps.  Needs the FP16 split
                # to be on (range guard: overflowed()) and channel counts that are multiples of 64 (a k-block = one tap x 64 channels).
                chain = self.f16 and all(c1.cout % 64 == 0 and c2.planes_capable() for c1, c2 in self.img_blocks_pm) and \
                    all(c1.planes_capable() for c1, _ in self.img_blocks_pm[1:])
                planes = None
                for bi, (conv1, conv2) in enumerate(self.img_blocks_pm):
                    if chain:
                        _, mid = conv1(x if bi == 0 else planes, relu=True, planes_out=True, f32_out=False)
                        last = bi + 1 == len(self.img_blocks_pm)
                        if last:
                            x = conv2(mid, relu=False)
                        else:
                            x, planes = conv2(mid, relu=False, planes_out=True)
                    else:
                        x = run_conv(conv2, run_conv(conv1, x, True), False)
                    ev = torch.cuda.Event()
                    ev.record(self.s_img)
                    imgs.append(x)
                    img_channels.append(conv2.cout)
                    img_done.append(ev)
                self._img_channels = img_channels
                if self.sparse_tail is not None:
                    # final image fusion at the sampled taps only; its result is the gathered feature row of every point
                    img_fusion = torch.empty((B * N, self.img_fuse_pm.N), **f32)
                    self.sparse_tail(imgs, xyn, img_fusion)
                else:
                    cat = torch.empty((B, H, W, self.deconv_cat_width), **f32)
                    col = 0
                    for i, de in enumerate(self.deconv_pm):  # each input pixel's k x k patch goes straight into its slice of the concat
                        de(imgs[i], cat[..., col:col + de.cout])
                        col += de.cout
                    img_fusion = self.img_fuse_pm(cat.view(-1, self.deconv_cat_width), relu=True).view(B, H, W, -1)
                img_fusion_done = torch.cuda.Event()
                img_fusion_done.record(self.s_img)
        feats = self._point_stream_pm(main, l_xyz, l_xy, fps_done, imgs, img_done, img_fusion, img_fusion_done, xyn)
        main.wait_stream(self.s_fps)
        main.wait_stream(self.s_img)
        for st in self.s_geo + self.s_scale:
            main.wait_stream(st)
        return xyz0, feats

    # ------------------------------------------------------------------------------------ point-major path
    def _build_pm(self):
        """tcgen05 GEMM parameters: BN folded, weights split/swizzled once (gemm.PackedLinear)."""
        self.sa_pm = []
        for npoint, scales in self.sa:
            packed = []
            for radius, ns, layers in scales:
                w0, b0 = layers[0]
                w0 = torch.cat([w0[:, 3:], w0[:, :3]], dim=1)  # group_concat_pm puts the 3 re-centred xyz channels last
                lins = [PackedLinear(w0, b0)] + [PackedLinear(w, b) for w, b in layers[1:]]
                packed.append((radius, ns, lins))
            self.sa_pm.append((npoint, packed))
        self.fp_pm = [[PackedLinear(w, b) for w, b in layers] for layers in self.fp]
        if self.cfg.li_fusion:
            self.fusion_pm = [_FusionPM(m) for m in self.model.Fusion_Conv]
            self.final_fusion_pm = _FusionPM(self.model.final_fusion_img_point)
            # image stream on the tcgen05 GEMM, NHWC
            self.img_blocks_pm = [(PackedConv3x3(w1, b1, stride=stride[0]), PackedConv3x3(conv2.weight, None, stride=conv2.stride[0]))
                                  for (w1, b1, stride, conv2) in self.img_blocks]
            self.deconv_pm, biases = [], []
            for de in self.model.DeConv:  # ConvTranspose2d weight: (Cin, Cout, k, k), kernel == stride
                k, co = de.kernel_size[0], de.out_channels
                assert de.stride[0] == k and de.kernel_size[1] == k and de.stride[1] == k
                self.deconv_pm.append(PackedDeconv(de.weight, None))
                biases.append(de.bias.detach() if de.bias is not None else torch.zeros(co, device=de.weight.device))
            self.deconv_cat_width = sum(de.cout for de in self.deconv_pm)
            fc, fbn = self.model.image_fusion_conv, self.model.image_fusion_bn
            wq = fc.weight.detach().flatten(1)
            # the transposed convolutions' biases pass linearly through the 1x1 fusion conv: fold them into its bias
            bq = (fc.bias.detach() if fc.bias is not None else 0) + wq @ torch.cat(biases)
            wf, bf = _fold_bn(wq, bq, fbn)
            self.img_fuse_pm = PackedLinear(wf, bf)
            self.sparse_tail = None
            if self.want_sparse_tail:
                try:
                    self.sparse_tail = SparseImageTail(list(self.model.DeConv), wf, bf, self.B, self.N, self.image_hw, self.device,
                                                       self.cfg.align_corners)
                except NotImplementedError:
                    self.sparse_tail = None  # a configuration outside its scope (kernel != stride, > 4 levels ...): dense tail

    def _point_stream_pm(self, main, l_xyz, l_xy, fps_done, imgs, img_done, img_fusion, img_fusion_done, xyn):
        c, B, N, dev = self.cfg, self.B, self.N, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        # Geometry-only work depends on the sampled coordinates alone, not on features: every level's ball queries and
        # three_nn searches are issued on side streams as soon as the FPS level they need is done, off the feature path.
        bq, nn = {}, {}
        # large clouds are Morton-sorted into 64-point buckets once per level (as soon as the level's coordinates exist, i.e. in
        # parallel with the FPS that samples them); both radii of the level then search the buckets instead of the whole cloud
        buckets = {}
        with torch.cuda.stream(self.s_geo[0]):
            for k in range(len(self.sa_pm)):
                if pc.SORTED_QUERY_MIN_N <= l_xyz[k].shape[1] <= pc.SORTED_QUERY_MAX_N:
                    self.s_geo[0].wait_event(self._ready if k == 0 else fps_done[k - 1])
                    ev = torch.cuda.Event()
                    buckets[k] = (pc.bucket_cloud(l_xyz[k]), ev)
                    ev.record(self.s_geo[0])
        for si in range(2):
            with torch.cuda.stream(self.s_geo[si]):
                for k, (npoint, scales) in enumerate(self.sa_pm):
                    if si >= len(scales):
                        continue
                    radius, ns, _ = scales[si]
                    self.s_geo[si].wait_event(fps_done[k])
                    bidx = torch.zeros((B, npoint, ns), dtype=torch.int32, device=dev)
                    if k in buckets and ns <= pc.SORTED_QUERY_MAX_NSAMPLE:
                        self.s_geo[si].wait_event(buckets[k][1])
                        pc.ball_query_sorted_wrapper(B, npoint, radius, ns, l_xyz[k + 1], buckets[k][0], bidx)
                    else:
                        pc.ball_query_wrapper(B, l_xyz[k].shape[1], npoint, radius, ns, l_xyz[k + 1], l_xyz[k], bidx)
                    ev = torch.cuda.Event()
                    ev.record(self.s_geo[si])
                    bq[(k, si)] = (bidx, ev)
        with torch.cuda.stream(self.s_geo[2]):
            for k in range(len(self.fp_pm)):
                self.s_geo[2].wait_event(fps_done[k])
                unknown, known = l_xyz[k], l_xyz[k + 1]
                dist2 = torch.empty((B, unknown.shape[1], 3), **f32)
                idx3 = torch.empty((B, unknown.shape[1], 3), dtype=torch.int32, device=dev)
                wts = torch.empty((B, unknown.shape[1], 3), **f32)
                pc.three_nn_weights_wrapper(B, unknown.shape[1], known.shape[1], unknown, known, dist2, idx3, wts)
                ev = torch.cuda.Event()
                ev.record(self.s_geo[2])
                nn[k] = (wts, idx3, ev)

        l_feat = [None]  # (B*n_k, C_k) point-major
        for k, (npoint, scales) in enumerate(self.sa_pm):
            torch.cuda.nvtx.range_push("SA%d" % (k + 1))
            main.wait_event(fps_done[k])
            xyz, new_xyz, feats = l_xyz[k], l_xyz[k + 1], l_feat[k]
            n = xyz.shape[1]
            cin = 0 if feats is None else feats.shape[1]
            ctot = sum(lins[-1].N for _, _, lins in scales)
            width = self.fusion_pm[k].cat_width if c.li_fusion else ctot
            cat = torch.empty((B * npoint, width), **f32)  # [ SA output | (attended) image features ]
            # the scales of a level are independent: they run as parallel branches (main stream + side streams), each
            # writing its own column slice of the level's output
            fork = torch.cuda.Event()
            fork.record(main)
            c_off, joins = 0, []
            for si, (radius, ns, lins) in enumerate(scales):
                st = main if si == 0 else self.s_scale[(si - 1) % len(self.s_scale)]
                with torch.cuda.stream(st):
                    if st is not main:
                        st.wait_event(fork)
                    if (k, si) in bq:
                        bidx, bq_ev = bq[(k, si)]
                        st.wait_event(bq_ev)
                    else:  # more than two scales: query here
                        bidx = torch.zeros((B, npoint, ns), dtype=torch.int32, device=dev)
                        pc.ball_query_wrapper(B, n, npoint, radius, ns, new_xyz, xyz, bidx)
                    fpm = None if feats is None else feats.view(B, n, cin)
                    if feats is None and self.fused_first_level and FusedFirstLevel.supports(lins, ns):
                        if (k, si) not in self._fused_sa:
                            self._fused_sa[(k, si)] = FusedFirstLevel(lins, ns)
                        self._fused_sa[(k, si)](xyz, new_xyz, bidx, cat[:, c_off:c_off + lins[-1].N])
                        if st is not main:
                            ev = torch.cuda.Event()
                            ev.record(st)
                            joins.append(ev)
                        c_off += lins[-1].N
                        continue
                    # first layer: the grouped rows go from the feature table straight into the tensor-core operand
                    x = grouped_first_layer(lins[0], xyz, new_xyz, fpm, bidx, relu=True) if len(lins) > 1 else None
                    rest = lins[1:-1]
                    if x is None:  # a one-layer MLP (its only layer carries the pooled epilogue): materialise the grouped rows
                        kp = (cin + 3 + 3) // 4 * 4
                        x = torch.empty((B * npoint * ns, kp), **f32)
                        pc.group_concat_pm_wrapper(B, cin, n, npoint, ns, xyz, new_xyz, fpm, bidx, x)
                        rest = lins[:-1]
                    for lin in rest:
                        x = lin(x, relu=True)
                    lins[-1](x, relu=True, pool=ns, out=cat[:, c_off:c_off + lins[-1].N])  # ReLU + max over nsample in the epilogue
                    if st is not main:
                        ev = torch.cuda.Event()
                        ev.record(st)
                        joins.append(ev)
                c_off += lins[-1].N
            for ev in joins:
                main.wait_event(ev)
            if c.li_fusion:
                main.wait_event(img_done[k])
                img = imgs[k]  # NHWC, possibly channel-padded
                ci = self._img_channels[k]
                g = torch.empty((B * npoint, ci), **f32)
                pc.grid_gather_nhwc_pm_wrapper(B, ci, img.shape[1], img.shape[2], npoint, img, l_xy[k + 1], c.align_corners, g)
                l_feat.append(self.fusion_pm[k](cat, g))
            else:
                l_feat.append(cat)
            torch.cuda.nvtx.range_pop()

        for i in range(-1, -(len(self.fp_pm) + 1), -1):
            unknown, known = l_xyz[i - 1], l_xyz[i]
            n, m = unknown.shape[1], known.shape[1]
            skip, kf = l_feat[i - 1], l_feat[i]
            c1 = 0 if skip is None else skip.shape[1]
            c2 = kf.shape[1]
            wts, idx3, nn_ev = nn[len(self.fp_pm) + i]  # searched on the geometry stream right after the FPS level
            torch.cuda.nvtx.range_push("FP%d" % (len(self.fp_pm) + i))
            main.wait_event(nn_ev)
            x = torch.empty((B * n, c2 + c1), **f32)
            pc.three_interpolate_concat_pm_wrapper(B, c2, m, n, c1, kf, idx3, wts, skip, x)
            lins = self.fp_pm[i]
            last_out = None
            if i - 1 == -(len(self.fp_pm) + 1) and c.li_fusion:  # level 0: write straight into the final fusion's concat buffer
                final_cat = torch.empty((B * n, self.final_fusion_pm.cat_width), **f32)
                last_out = final_cat[:, :lins[-1].N]
            for lin in lins[:-1]:
                x = lin(x, relu=True)
            l_feat[i - 1] = lins[-1](x, relu=True, out=last_out)
            torch.cuda.nvtx.range_pop()

        feats = l_feat[0]
        if c.li_fusion:
            main.wait_event(img_fusion_done)
            if self.sparse_tail is not None:
                g = img_fusion  # (B*N, Cf): already gathered
            else:
                ci = img_fusion.shape[3]  # NHWC
                g = torch.empty((B * N, ci), **f32)
                pc.grid_gather_nhwc_pm_wrapper(B, ci, img_fusion.shape[1], img_fusion.shape[2], N, img_fusion, xyn, c.align_corners, g)
            # the last GEMM writes the interface layout (B, C, N) itself: no transposing pass
            return self.final_fusion_pm(final_cat, g, out_cm=torch.empty((B, self.final_fusion_pm.cp, N), **f32))
        return feats.view(B, N, -1).transpose(1, 2).contiguous()  # interface layout (B, C, N)

    # ------------------------------------------------------------------------------------ graph
    def _capture(self):
        dev = self.device
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(warm), torch.no_grad():
            for _ in range(3):  # cuDNN/cuBLAS autotuning and workspace allocation happen outside capture
                self._forward()
        torch.cuda.current_stream(dev).wait_stream(warm)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        before = pc.LAUNCHES[0]
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._forward()
        self.kernel_launches_per_replay = pc.LAUNCHES[0] - before + 1  # + the image layout/normalisation kernel in front of the replay

    def _load(self, points, image, xy, sizes=None):
        self.points.copy_(points, non_blocking=True)
        self.xy.copy_(xy, non_blocking=True)
        if image.dtype == torch.uint8:
            # decoded camera image (B,h,w,3) uint8, h <= H, w <= W: normalised, zero-padded and laid out on the device
            if not image.is_cuda:
                if self._stage_u8 is None or self._stage_u8.shape != image.shape:
                    self._stage_u8 = torch.empty(image.shape, dtype=torch.uint8, device=self.device)
                self._stage_u8.copy_(image, non_blocking=True)
                image = self._stage_u8
            if sizes is not None and not sizes.is_cuda:
                sizes = sizes.to(self.device, non_blocking=True)
            image_prep.normalise_pad(image, sizes, out_hw=tuple(self.image4.shape[1:3]), nhwc4=self.image4)
            return
        if not image.is_cuda or not image.is_contiguous():
            if self._stage_f32 is None:
                self._stage_f32 = torch.empty(self.image4.shape[0], 3, self.image4.shape[1], self.image4.shape[2], dtype=torch.float32,
                                              device=self.device)
            self._stage_f32.copy_(image, non_blocking=True)
            image = self._stage_f32
        image_prep.nchw_to_nhwc4(image, self.image4)

    def __call__(self, points, image, xy, sizes=None):
        """points (B,N,3); image: the reference's fp32 (B,3,H,W) tensor, or the decoded uint8 RGB image (B,h,w,3) with optional
        `sizes` (B,2) int32 {rows, cols} per scene (normalised and zero-padded on the device, image_prep.py); xy (B,N,2) pixel
        coordinates.  Inputs may be host-pinned or on the device and are NOT modified -> (xyz (B,N,3), features (B,128,N)) in
        buffers owned by the runner, valid until the next call."""
        self._load(points, image, xy, sizes)
        if self.graph is None:
            with torch.no_grad():
                out = self._forward()
        else:
            self.graph.replay()
            out = self.out
        self._last_stream = torch.cuda.current_stream(self.device)
        self.overflow.read_async()  # 4 bytes, ordered after the forward on this stream
        return out

    def overflowed(self):
        """True when a GEMM of ANY forward on this device since the last reset wrote a value an FP16-split layer cannot represent
        (the flag is per device and sticky).  Synchronises with this runner's last call.  The results of such a forward are
        not trustworthy when f16 is on: rebuild the runner with f16=False (after `overflow.reset()`) and run again."""
        if self._last_stream is None:
            return False
        self._last_stream.synchronize()
        return self.overflow.value() != 0

    def eager(self, points, image, xy, single_stream=False):
        """Same schedule without the graph (profiling: events around individual launches); single_stream=True also
        serialises the three branches on the current stream so that per-kernel durations are not inflated by overlap."""
        self._load(points, image, xy)
        saved = (self.s_fps, self.s_img, self.s_scale, self.s_geo)
        if single_stream:
            self.s_fps = self.s_img = torch.cuda.current_stream(self.device)
            self.s_scale = [self.s_fps]
            self.s_geo = [self.s_fps] * 3
        try:
            with torch.no_grad():
                return self._forward()
        finally:
            self.s_fps, self.s_img, self.s_scale, self.s_geo = saved


class PipelinedRunner:
    """`depth` BackboneRunners (own static buffers, own captured graph, shared weights) replayed round-robin on their own
    streams, so that successive batches overlap on the GPU.  One forward is latency-bound -- its critical path is the serial
    FPS chain, which occupies one SM per scene -- so a single batch leaves most of the 148 SMs idle most of the time; a
    serving loop that keeps 2-3 batches in flight fills them.  Results of call i stay valid until call i + depth."""

    def __init__(self, model, batch, npoints, device, depth=2, **kw):
        self.device = device
        kw.setdefault("tiles", "throughput")  # several batches in flight fill the GPU: widest tiles, least traffic per flop
        self.runners = [BackboneRunner(model, batch, npoints, device, **kw) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(depth)]
        self.kernel_launches_per_replay = self.runners[0].kernel_launches_per_replay
        self.calls = 0
        self._copy_stream = None
        self._staging = {}
        self._last = None

    def read_back(self, key, host_xyz, host_feats):
        """Deliver the results of the LAST call to pinned host tensors without holding its slot for the duration of the transfer:
        a device-side copy into a staging pair (on the slot's stream, microseconds) frees the slot, the device->host copy runs
        on a dedicated copy stream.  `key` names the staging pair (reuse a key only after the event returned for it has completed:
        a ring of keys as deep as the host may lag).  Returns the event that marks the host tensors valid.  With eight ranks
        sharing one host link a 17 MB read-back takes as long as a whole step; on the slot's own stream it made every slot's
        cycle that much longer (tools/e2e_scaling_probe.py)."""
        i = (self.calls - 1) % len(self.runners)
        xyz, feats = self._last
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        if key not in self._staging:
            self._staging[key] = (torch.empty_like(xyz), torch.empty_like(feats))
        sx, sf = self._staging[key]
        st = self.streams[i]
        with torch.cuda.stream(st):
            sx.copy_(xyz, non_blocking=True)
            sf.copy_(feats, non_blocking=True)
        self._copy_stream.wait_stream(st)
        with torch.cuda.stream(self._copy_stream):
            host_xyz.copy_(sx, non_blocking=True)
            host_feats.copy_(sf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return ev

    def __call__(self, points, image, xy, sizes=None):
        i = self.calls % len(self.runners)
        self.calls += 1
        st = self.streams[i]
        st.wait_stream(torch.cuda.current_stream(self.device))  # inputs produced on the caller's stream
        with torch.cuda.stream(st):
            self._last = self.runners[i](points, image, xy, sizes)
        return self._last

    def stream_of_last_call(self):
        return self.streams[(self.calls - 1) % len(self.runners)]

    def overflowed(self):
        """see BackboneRunner.overflowed; waits for every in-flight batch"""
        return any([r.overflowed() for r in self.runners])

    @property
    def overflow(self):
        return self.runners[0].overflow

    def join(self):
        """make the caller's current stream wait for every in-flight batch (and every read_back in flight)"""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)
        if self._copy_stream is not None:
            cur.wait_stream(self._copy_stream)

    def eager(self, *a, **kw):
        return self.runners[0].eager(*a, **kw)
