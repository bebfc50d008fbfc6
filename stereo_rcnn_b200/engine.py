"""The Stereo R-CNN test-mode forward on hand-written sm_100a kernels.

Host-side schedule of `_StereoRCNN.forward` (lib/model/stereo_rcnn/stereo_rcnn.py:141-324):
weights are packed once (NHWC / K-major, frozen BN folded to per-channel scale+shift --
resnet.py:300-309 freezes every BN), left and right images run through the trunk as one
batch of two, and every stage is a libstereo_b200 kernel launched on torch's current stream.
There is no PyTorch operator on the compute path (torch only allocates device memory).
"""
import torch

from . import ops

LAYERS = [3, 4, 23, 3]
PLANES = [64, 128, 256, 512]
STRIDES = [1, 2, 2, 2]
EPS = 1e-5


def _pack_conv(w):
    """[Cout,Cin,kh,kw] -> [Cout,kh,kw,Cin] contiguous (K-major rows, one per output channel)"""
    return w.permute(0, 2, 3, 1).contiguous()


class PackedConv(object):
    __slots__ = ("w", "w16", "scale", "shift", "Cin", "Cout", "kh", "kw", "pad")

    round_weights = True      # class-wide switch: False keeps exact fp32 weights (SIMT yardstick runs)
    make_half = False         # class-wide switch: also keep an fp16 copy (fp16-operand mode)

    def __init__(self, w, scale, shift, pad):
        self.Cout, self.Cin, self.kh, self.kw = w.shape
        wp = _pack_conv(w).clone()
        self.w16 = wp.to(torch.float16) if PackedConv.make_half else None    # kind::f16 operand (round to nearest)
        if PackedConv.round_weights:   # round to TF32 (nearest) once: the tensor core would otherwise truncate
            ops.round_tf32_(wp)
        self.w, self.scale, self.shift, self.pad = wp, scale, shift, pad


class StereoRCNNEngine(object):
    """state_dict uses the reference's key names (RCNN_layer1.0.0.conv1.weight, ...)"""

    def __init__(self, state_dict, device="cuda", n_classes=2, conv_impl="auto", precision=None, lr_streams=None):
        """precision: "tf32" (fp32 storage, kind::tf32) or "fp16" (fp16 conv operands, kind::f16, fp32 accumulate
        and fp32 residual stream); default from $SB_PRECISION, else "fp16" (same measured accuracy as tf32 --
        both round operands to an 11-bit significand -- at twice the MMA rate and half the activation bytes).  conv_impl="simt" forces exact fp32."""
        import os
        precision = precision or os.environ.get("SB_PRECISION", "fp16")
        assert precision in ("tf32", "fp16")
        self.half = precision == "fp16" and conv_impl != "simt"
        self.precision = "fp32-simt" if conv_impl == "simt" else precision
        self.keep32 = False
        self.chain_ctas = int(os.environ.get("SB_CHAIN_CTAS", "0"))
        self.stem_fused = os.environ.get("SB_STEM_FUSED", "1") != "0"
        self.rpn_streams = os.environ.get("SB_RPN_STREAMS", "1") != "0"
        self.head_streams = os.environ.get("SB_HEAD_STREAMS", "1") != "0"
        # lr_streams: run layers 3-4 of the left and the right image as two concurrent chains (lowest latency of a
        # single pair).  With several independent pairs in flight the SMs are already full and the batched chain
        # (fewer, wider tiles) costs less SM time: callers that pipeline pairs pass lr_streams=False.
        if lr_streams is None:
            lr_streams = os.environ.get("SB_LR_STREAMS", "1") != "0"
        self.lr_streams = bool(lr_streams)
        self.side = torch.cuda.Stream(device=torch.device(device)) if torch.device(device).type == "cuda" else None
        self.device = torch.device(device)
        self.n_classes = n_classes
        self.conv_impl = conv_impl
        self.impl_used = {}
        self.max_ctas = 0         # grid cap for convs issued while a second chain runs on the side stream
        self.record = None        # bench: list collecting (desc, live tensors) of every conv launch of a forward
        # conv_impl="simt" is the exact-fp32 yardstick: exact weights, exact stores, no TF32 hygiene modes
        self.exact = conv_impl == "simt"
        PackedConv.round_weights = not self.exact
        PackedConv.make_half = self.half
        # Weights are folded and packed ON THE HOST (one-off, ~1 s) and reach the device by plain copies: no torch
        # kernel runs on the GPU for it, so a profiler's launch window of a process that builds an engine shows this
        # library's kernels, not ~950 elementwise launches of weight preparation.
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in state_dict.items()
              if not k.endswith("num_batches_tracked")}
        self.p = {}

        def bn_fold(k):
            s = sd[k + ".weight"] / torch.sqrt(sd[k + ".running_var"] + EPS)
            return s.contiguous(), (sd[k + ".bias"] - sd[k + ".running_mean"] * s).contiguous()

        def conv_bn(ck, bk, pad):
            s, b = bn_fold(bk)
            self.p[ck] = PackedConv(sd[ck + ".weight"], s, b, pad)

        def conv_bias(ck, pad):
            self.p[ck] = PackedConv(sd[ck + ".weight"], None, sd[ck + ".bias"].contiguous(), pad)

        # stem: [64,3,7,7] -> [64][7][7][3]
        s, b = bn_fold("RCNN_layer0.1")
        self.stem = (_pack_conv(sd["RCNN_layer0.0.weight"]), s, b)
        wst = torch.zeros(64, 160, 1, 1)                              # stem as a GEMM over the padded patch matrix
        wst[:, :147, 0, 0] = sd["RCNN_layer0.0.weight"].reshape(64, 147)      # (ci, r, s) order of sb_stem_im2col
        self.p["stem_gemm"] = PackedConv(wst, s, b, 0)
        wst = torch.zeros(64, 192, 1, 1)                              # fp16 variant: three 64-wide K-steps
        wst[:, :147, 0, 0] = sd["RCNN_layer0.0.weight"].reshape(64, 147)
        self.p["stem_gemm16"] = PackedConv(wst, s, b, 0)
        self.stem_w16 = ops.pack_stem_w16(sd["RCNN_layer0.0.weight"])   # K layout of the fused (implicit-GEMM) stem
        for li, nb in enumerate(LAYERS):
            for bi in range(nb):
                p = "RCNN_layer%d.0.%d" % (li + 1, bi)
                conv_bn(p + ".conv1", p + ".bn1", 0)
                conv_bn(p + ".conv2", p + ".bn2", 1)
                conv_bn(p + ".conv3", p + ".bn3", 0)
                if bi == 0:
                    conv_bn(p + ".downsample.0", p + ".downsample.1", 0)
        for k in ("RCNN_toplayer", "RCNN_latlayer1", "RCNN_latlayer2", "RCNN_latlayer3"):
            conv_bias(k, 0)
        for k in ("RCNN_smooth1", "RCNN_smooth2", "RCNN_smooth3", "RCNN_rpn.RPN_Conv"):
            conv_bias(k, 1)
        # RPN 1x1 heads fused: rows 0..5 cls logits, 6..23 box deltas, 24..31 zero padding
        wh = torch.zeros(32, 1024, 1, 1)
        bh = torch.zeros(32)
        wh[0:6] = sd["RCNN_rpn.RPN_cls_score.weight"]
        wh[6:24] = sd["RCNN_rpn.RPN_bbox_pred_left_right.weight"]
        bh[0:6] = sd["RCNN_rpn.RPN_cls_score.bias"]
        bh[6:24] = sd["RCNN_rpn.RPN_bbox_pred_left_right.bias"]
        self.p["rpn_heads"] = PackedConv(wh, None, bh, 0)
        # box head: Conv2d(512,2048,k=7,s=7) on a 7x7 map == FC over (ph,pw,c) of the NHWC pooled tile
        w0 = sd["RCNN_top.0.weight"]                                   # [2048,512,7,7]
        self.p["RCNN_top.0"] = PackedConv(_pack_conv(w0).reshape(2048, 7 * 7 * 512, 1, 1), None,
                                          sd["RCNN_top.0.bias"].contiguous(), 0)
        conv_bias("RCNN_top.3", 0)
        for i in range(0, 12, 2):
            conv_bias("RCNN_kpts.%d" % i, 1)
        # ConvTranspose2d(256,256,2,2): out[2i+a,2j+b] = W[:,:,a,b]^T x[i,j]  -> four 1x1 convs
        wd = sd["RCNN_kpts.12.weight"]                                 # [Cin,Cout,2,2]
        self.deconv = [[PackedConv(wd[:, :, a, b].t().contiguous().reshape(256, 256, 1, 1), None,
                                   sd["RCNN_kpts.12.bias"].contiguous(), 0) for b in range(2)] for a in range(2)]
        self.kpts_class = (sd["kpts_class.weight"].reshape(6, 256).contiguous(), sd["kpts_class.bias"].contiguous())
        self.fc = [sd[k].contiguous() for k in ("RCNN_cls_score.weight", "RCNN_cls_score.bias",
                                                "RCNN_bbox_pred.weight", "RCNN_bbox_pred.bias",
                                                "RCNN_dim_orien_pred.weight", "RCNN_dim_orien_pred.bias")]
        # ---- host -> device (copies only)
        mv = lambda t: None if t is None else t.contiguous().to(self.device)
        for pc in list(self.p.values()) + [c for row in self.deconv for c in row]:
            pc.w, pc.w16, pc.scale, pc.shift = mv(pc.w), mv(pc.w16), mv(pc.scale), mv(pc.shift)
        self.stem = tuple(mv(t) for t in self.stem)
        self.stem_w16 = mv(self.stem_w16)
        self.kpts_class = tuple(mv(t) for t in self.kpts_class)
        self.fc = [mv(t) for t in self.fc]

    # ------------------------------------------------------------------ helpers
    def _conv(self, x, pc, relu=False, stride=1, residual=None, up_src=None, out=None, out_coff=0,
              out_strides=None, Cin=None, tag=None, out_mode=ops.EXACT, res_biased=False, in_biased=False,
              f32=True, f16=False, out16=None):
        """x fp32 (tf32 / simt modes) or fp16 (fp16 mode).  Returns the fp32 output, the fp16 twin, or both."""
        N, H, W = x.shape[:3]
        Ho = (H + 2 * pc.pad - pc.kh) // stride + 1
        Wo = (W + 2 * pc.pad - pc.kw) // stride + 1
        half_in = x.dtype == torch.float16
        if out is None and f32:
            out = torch.empty(N, Ho, Wo, pc.Cout, dtype=torch.float32, device=x.device)
        if out16 is None and f16:
            out16 = torch.empty(N, Ho, Wo, pc.Cout, dtype=torch.float16, device=x.device)
        if self.exact or half_in:
            out_mode, res_biased, in_biased = ops.EXACT, False, False
        d = ops.conv_desc(x, pc.w16 if half_in else pc.w, out, pc.Cin if Cin is None else Cin, pc.Cout, pc.kh,
                          pc.kw, stride, pc.pad, Ho, Wo, scale=pc.scale, shift=pc.shift, residual=residual,
                          up_src=up_src, relu=relu, out_coff=out_coff, out_strides=out_strides, out_mode=out_mode,
                          res_biased=res_biased, in_biased=in_biased, out16=out16, max_ctas=self.max_ctas)
        impl = ops.conv2d(d, "tc" if half_in else self.conv_impl)
        if self.record is not None:
            self.record.append((d, impl, (x, out, out16, residual, up_src)))
        if tag is not None:
            self.impl_used[tag] = impl + ("16" if half_in else "")
        if out is not None and out16 is not None:
            return out, out16
        return out if out is not None else out16

    def _bottleneck(self, x, prefix, stride, has_ds):
        """resnet.py:82-102; the stride sits on the 1x1 conv1 and on the downsample (Q1)"""
        # TF32 hygiene: the residual stream x is stored "pre-biased" (exact fp32 + 0x1000): the tensor core's
        # truncation of it is then round-to-nearest, and the residual add un-biases it exactly.  Tensors
        # read only by the next conv (o1, o2) are stored rounded to TF32.
        xin = ops.subsample2(x) if stride == 2 else x
        o = self._conv(xin, self.p[prefix + ".conv1"], relu=True, tag=prefix + ".conv1", out_mode=ops.ROUND_TF32,
                       in_biased=True)
        o = self._conv(o, self.p[prefix + ".conv2"], relu=True, tag=prefix + ".conv2", out_mode=ops.ROUND_TF32)
        if has_ds:
            res = self._conv(xin, self.p[prefix + ".downsample.0"], tag=prefix + ".ds", in_biased=True)
        else:
            res = x
        return self._conv(o, self.p[prefix + ".conv3"], relu=True, residual=res, tag=prefix + ".conv3",
                          out_mode=ops.BIASED, res_biased=not has_ds)

    def _bottleneck16(self, x32, x16, prefix, stride, has_ds, out=None):
        """fp16-operand variant: convs read fp16 twins, the residual stream itself stays exact fp32.
        out = (fp32 view, fp16 view) to write the block output into (slices of a batched tensor)"""
        xin = ops.subsample2(x16) if stride == 2 else x16
        o = self._conv(xin, self.p[prefix + ".conv1"], relu=True, tag=prefix + ".conv1", f32=False, f16=True)
        o = self._conv(o, self.p[prefix + ".conv2"], relu=True, tag=prefix + ".conv2", f32=False, f16=True)
        res = self._conv(xin, self.p[prefix + ".downsample.0"], tag=prefix + ".ds") if has_ds else x32
        if out is not None:
            return self._conv(o, self.p[prefix + ".conv3"], relu=True, residual=res, tag=prefix + ".conv3",
                              out=out[0], out16=out[1])
        return self._conv(o, self.p[prefix + ".conv3"], relu=True, residual=res, tag=prefix + ".conv3", f16=True)

    def trunk_fpn(self, im_nchw):
        """images [N,3,H,W] NCHW -> dict of NHWC C2..C5, P2..P6 (stereo_rcnn.py:155-168)"""
        if self.half:
            return self._trunk_fpn16(im_nchw)
        if self.exact:
            c0 = ops.stem_conv(im_nchw, *self.stem, out_mode=ops.EXACT)
        else:       # patch matrix (pixels rounded to TF32) + tcgen05 GEMM
            c0 = self._conv(ops.stem_im2col(im_nchw), self.p["stem_gemm"], relu=True, out_mode=ops.BIASED, tag="stem")
        c1 = ops.maxpool3x3s2_ceil(c0)                      # max() commutes with the monotone pre-bias
        feats = {"c1": c1}
        x = c1
        for li, nb in enumerate(LAYERS):
            for bi in range(nb):
                x = self._bottleneck(x, "RCNN_layer%d.0.%d" % (li + 1, bi), STRIDES[li] if bi == 0 else 1, bi == 0)
            feats["c%d" % (li + 2)] = x
        R = ops.ROUND_TF32
        p5 = self._conv(feats["c5"], self.p["RCNN_toplayer"], tag="toplayer", in_biased=True)
        t = self._conv(feats["c4"], self.p["RCNN_latlayer1"], up_src=p5, tag="lat1", out_mode=R, in_biased=True)
        p4 = self._conv(t, self.p["RCNN_smooth1"], tag="smooth1")          # lateral + upsample-add fused above
        t = self._conv(feats["c3"], self.p["RCNN_latlayer2"], up_src=p4, tag="lat2", out_mode=R, in_biased=True)
        p3 = self._conv(t, self.p["RCNN_smooth2"], tag="smooth2")
        t = self._conv(feats["c2"], self.p["RCNN_latlayer3"], up_src=p3, tag="lat3", out_mode=R, in_biased=True)
        p2 = self._conv(t, self.p["RCNN_smooth3"], tag="smooth3")
        p6 = ops.subsample2(p5)                                                           # Q5
        feats.update(p2=p2, p3=p3, p4=p4, p5=p5, p6=p6)
        return feats

    def _trunk_fpn16(self, im_nchw):
        """fp16-operand trunk: every conv input is an fp16 tensor (half the HBM bytes, kind::f16 = 2x the TF32
        MMA rate, same 11-bit significand); fp32 copies exist only where a non-conv consumer needs them
        (residual adds, FPN upsample source, RoIAlign, the user-visible P levels)."""
        if self.stem_fused:     # patches gathered inside the GEMM kernel: no 229 MB patch matrix
            pc = self.p["stem_gemm16"]
            c0 = ops.stem_conv_tc(im_nchw, self.stem_w16, pc.scale, pc.shift)
            self.impl_used["stem"] = "tc16-fused"
        else:
            c0 = self._conv(ops.stem_im2col16(im_nchw), self.p["stem_gemm16"], relu=True, tag="stem", f32=False, f16=True)
        c1 = ops.maxpool3x3s2_ceil(c0)
        feats = {"c1": c1.float() if self.keep32 else None, "c1_16": c1}
        x32, x16 = None, c1
        for li in (0, 1):                           # layers 1-2: left and right batched (hundreds of tiles)
            for bi in range(LAYERS[li]):
                x32, x16 = self._bottleneck16(x32, x16, "RCNN_layer%d.0.%d" % (li + 1, bi),
                                              STRIDES[li] if bi == 0 else 1, bi == 0)
            feats["c%d" % (li + 2)], feats["c%d_16" % (li + 2)] = x32, x16
        # layers 3-4 have 38 / 10 row tiles per image: far too few to fill 148 SMs, and every launch is latency
        # bound (prologue + load latency + main loop + epilogue back to back).  The left and the right image are
        # independent, so they run as two concurrent chains on two streams (two parallel branches of the CUDA
        # graph): the chains' latencies overlap and together they still need only ~76 SMs.
        N = x16.shape[0]
        B = N // 2
        dev = x16.device

        def out_bufs(li):
            h, w = x16.shape[1], x16.shape[2]
            for _ in range(li - 1):
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            c = PLANES[li] * 4
            return (torch.empty(N, h, w, c, dtype=torch.float32, device=dev),
                    torch.empty(N, h, w, c, dtype=torch.float16, device=dev))
        c4 = out_bufs(2)
        c5 = out_bufs(3)

        def chain(sl):
            y32, y16 = x32[sl], x16[sl]
            for li, dst in ((2, c4), (3, c5)):
                for bi in range(LAYERS[li]):
                    last = bi == LAYERS[li] - 1
                    y32, y16 = self._bottleneck16(y32, y16, "RCNN_layer%d.0.%d" % (li + 1, bi),
                                                  STRIDES[li] if bi == 0 else 1, bi == 0,
                                                  out=(dst[0][sl], dst[1][sl]) if last else None)
        if self.side is not None and self.lr_streams:
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            self.max_ctas = self.chain_ctas      # each chain keeps to half of the SMs so that both really co-run
            with torch.cuda.stream(self.side):
                chain(slice(B, N))
            chain(slice(0, B))
            self.max_ctas = 0
            main.wait_stream(self.side)
        else:
            chain(slice(0, N))
        feats.update(c4=c4[0], c4_16=c4[1], c5=c5[0], c5_16=c5[1])
        p5, p5h = self._conv(feats["c5_16"], self.p["RCNN_toplayer"], tag="toplayer", f16=True)
        t = self._conv(feats["c4_16"], self.p["RCNN_latlayer1"], up_src=p5, tag="lat1", f32=False, f16=True)
        p4, p4h = self._conv(t, self.p["RCNN_smooth1"], tag="smooth1", f16=True)
        t = self._conv(feats["c3_16"], self.p["RCNN_latlayer2"], up_src=p4, tag="lat2", f32=False, f16=True)
        p3, p3h = self._conv(t, self.p["RCNN_smooth2"], tag="smooth2", f16=True)
        t = self._conv(feats["c2_16"], self.p["RCNN_latlayer3"], up_src=p3, tag="lat3", f32=False, f16=True)
        p2, p2h = self._conv(t, self.p["RCNN_smooth3"], tag="smooth3", f16=True)
        feats.update(p2=p2, p3=p3, p4=p4, p5=p5, p6=ops.subsample2(p5), p2_16=p2h, p3_16=p3h, p4_16=p4h,
                     p5_16=p5h, p6_16=ops.subsample2(p5h))
        return feats

    def rpn(self, feats, B):
        """stereo_rpn.py:73-95 -> cls_prob [B,A,2], bbox_pred [B,A,6], level shapes"""
        sfx = "_16" if self.half else ""
        levels = [feats[k + sfx] for k in ("p2", "p3", "p4", "p5", "p6")]
        shapes = [[f.shape[1], f.shape[2]] for f in levels]
        P = sum(h * w for h, w in shapes)
        dev = self.device
        head = torch.empty(B, P, 32, dtype=torch.float32, device=dev)
        def level(f, h, w, off):
            cat = torch.empty(B, h, w, 1024, dtype=f.dtype, device=dev)
            kw = dict(f32=False, out16=cat) if self.half else dict(out=cat, out_mode=ops.ROUND_TF32)
            if B == 1:      # L and R in one launch: image index n = side -> channel offset n * 512
                self._conv(f, self.p["RCNN_rpn.RPN_Conv"], relu=True, out_strides=(512, w * 1024, 1024),
                           tag="rpn_conv", **kw)
            else:
                for side in range(2):   # shared RPN_Conv on L then R, channel-concatenated (Q6)
                    self._conv(f[side * B:(side + 1) * B], self.p["RCNN_rpn.RPN_Conv"], relu=True,
                               out_coff=side * 512, out_strides=(h * w * 1024, w * 1024, 1024), tag="rpn_conv", **kw)
            self._conv(cat, self.p["rpn_heads"], out=head[:, off:off + h * w],
                       out_strides=(P * 32, w * 32, 32), tag="rpn_heads")
        offs = [0]
        for h, w in shapes:
            offs.append(offs[-1] + h * w)
        fork = self.side is not None and self.rpn_streams
        if fork:    # the three coarse levels are a handful of tiles each: run them beside the p2 / p3 convs
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                for i in (2, 3, 4):
                    level(levels[i], shapes[i][0], shapes[i][1], offs[i])
        for i in ((0, 1) if fork else range(5)):
            level(levels[i], shapes[i][0], shapes[i][1], offs[i])
        if fork:
            main.wait_stream(self.side)
        cls_prob, bbox = ops.rpn_head_epilogue(head, B, P)
        return cls_prob, bbox, shapes

    def heads(self, feats, B, rois_l, rois_r, im_h):
        """stereo_rcnn.py:240-271 on flattened rois [R,5]"""
        mk = ("p2", "p3", "p4", "p5")
        fl = [feats[k][:B] for k in mk]
        fr = [feats[k][B:] for k in mk]
        R = rois_l.shape[0]
        dev = self.device
        h = self.half
        adt = torch.float16 if h else torch.float32
        rt = not self.exact and not h
        c16 = dict(f32=False, f16=True) if h else {}
        def box_head():
            pooled = torch.empty(R, 7, 7, 512, dtype=adt, device=dev)
            ops.roi_align_pyramid_nhwc(fl, im_h, rois_l, 7, out=pooled, out_coff=0, round_tf32=rt, half=h)
            ops.roi_align_pyramid_nhwc(fr, im_h, rois_r, 7, out=pooled, out_coff=256, round_tf32=rt, half=h)
            x = self._conv(pooled.view(R, 1, 1, 7 * 7 * 512), self.p["RCNN_top.0"], relu=True, tag="top0",
                           out_mode=ops.ROUND_TF32, **c16)
            fc7 = self._conv(x, self.p["RCNN_top.3"], relu=True, tag="top3").view(R, 2048)
            return (pooled, fc7) + tuple(ops.box_tail(fc7, *self.fc, n_classes=self.n_classes))
        fork = self.side is not None and self.head_streams
        if fork:    # the box head (two FCs with 3 row tiles) is latency-bound: run it beside the keypoint convs
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                pooled, fc7, cls_prob, bbox, dim = box_head()
        else:
            pooled, fc7, cls_prob, bbox, dim = box_head()
        pk = ops.roi_align_pyramid_nhwc(fl, im_h, rois_l, 14, round_tf32=rt, half=h)
        x = pk
        for i in range(0, 12, 2):
            x = self._conv(x, self.p["RCNN_kpts.%d" % i], relu=True, tag="kpts%d" % i, out_mode=ops.ROUND_TF32, **c16)
        up = torch.empty(R, 28, 28, 256, dtype=adt, device=dev)
        for a in range(2):
            for b in range(2):
                kw = dict(f32=False, out16=up[:, a:, b:]) if h else dict(out=up[:, a:, b:])
                self._conv(x, self.deconv[a][b], relu=True, out_strides=(28 * 28 * 256, 2 * 28 * 256, 2 * 256),
                           tag="deconv", **kw)
        kp, lb, rb, ka = ops.kpts_tail(up, *self.kpts_class, want_pred_all=True)
        if fork:
            main.wait_stream(self.side)
        return dict(pooled_box=pooled, pooled_kpts=pk, fc7=fc7, cls_prob=cls_prob, bbox_pred=bbox,
                    dim_orien_pred=dim, kpts_prob=kp, left_border_prob=lb, right_border_prob=rb,
                    kpts_pred_all=ka)

    @torch.no_grad()
    def forward(self, im_left, im_right, im_info, cfg_key="TEST", keep_features=False, im_h=None,
                before_proposals=None):
        """im_left/right [B,3,H,W] NCHW fp32 (BGR - means), im_info [B,3] (device) -> dict (the reference's
        tuple order is produced by model.stereo_rcnn.resnet.resnet.forward).  `im_h` (= im_info[0][0],
        stereo_rcnn.py:128) is taken from the tensor shape so that no device->host read is needed."""
        B = im_left.shape[0]
        im_h = float(im_left.shape[2]) if im_h is None else float(im_h)
        im = torch.cat((im_left, im_right), 0).contiguous()
        feats = self.trunk_fpn(im)
        cls_prob, bbox, shapes = self.rpn(feats, B)
        if before_proposals is not None:
            # the proposal stage is a chain of small, partly single-CTA kernels: callers fork independent work
            # (e.g. dense_align of the previous detections) onto a second stream here so that it fills the idle SMs
            before_proposals()
        rl, rr = ops.proposal_layer(cls_prob, bbox, im_info, cfg_key, shapes)
        out = self.heads(feats, B, rl.view(-1, 5), rr.view(-1, 5), im_h)
        n = rl.shape[1]
        out.update(rois_left=rl, rois_right=rr, rpn_cls_prob=cls_prob, rpn_bbox_pred=bbox, rpn_shapes=shapes)
        out["cls_prob"] = out["cls_prob"].view(B, n, -1)
        out["bbox_pred"] = out["bbox_pred"].view(B, n, -1)
        out["dim_orien_pred"] = out["dim_orien_pred"].view(B, n, -1)
        if keep_features:      # debugging / tests: exact-fp32 views of the pre-biased trunk tensors
            ub = not self.exact and not self.half
            out["feats"] = {k: (ops.unbias(v) if (k[0] == "c" and ub and v is not None and v.dtype == torch.float32)
                                else v) for k, v in feats.items()}
            out["feats_raw"] = feats
        return out


class GraphRunner(object):
    """Capture a launch-bound sequence of libstereo_b200 kernels once into a CUDA graph and replay it.

    `fn(*static_inputs)` must be free of host<->device syncs (every entry point of the C ABI is) and is run
    `warmup` times eagerly first so that workspaces, function attributes and TMA descriptors exist.  Inputs are
    fixed tensors: copy new data into them (`runner.inputs[i].copy_(...)`) before `runner()`.
    """

    def __init__(self, fn, inputs, warmup=2):
        self.fn, self.inputs = fn, list(inputs)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn(*self.inputs)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*self.inputs)

    def __call__(self):
        self.graph.replay()
        return self.outputs
