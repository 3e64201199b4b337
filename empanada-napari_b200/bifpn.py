"""PanopticBiFPN-PointRend (MitoNet_v1_mini: ResNet-50 at output stride 32, two 3-layer BiFPNs of
128 channels, BiFPN decoders) on the sm_100a kernels: weight ingestion from the reference's
TorchScript export and the recorded launch list.

Network structure follows the reference (file:line under /root/reference/empanada/models):
  forward            quantization/panoptic_bifpn.py:147-161 (render_steps=2, interpolate_ins=False)
  BiFPN              decoders/bifpn.py:14-196 (TopDownFPN, BottomUpFPN, BiFPNLayer, BiFPN)
  decoder            decoders/bifpn.py:198-236 (5 x ConvTranspose2d k2 s2 + BN + ReLU, skip concat,
                     5x5 separable fusion)
  blocks             blocks.py:52-171 (Resample2d, Resize2d, separable_conv_bn_act, conv_bn_act,
                     conv_transpose_bn_act)
  heads / PointRend  as PanopticDeepLab (pdl.py)

Kernel mapping: every 1x1 convolution (channel resampling, pointwise halves of the separable
blocks, the transposed convolutions as N = 4*Cout GEMMs with a pixel-shuffle epilogue, PointRend
MLP) is the tcgen05 implicit GEMM with BatchNorm folded into weights/bias and SiLU/ReLU in the
epilogue; the fast-normalised fusion (with its nearest x2 / max-pool resize) is one elementwise
kernel; depthwise 3x3 / 5x5 use the shared-memory depthwise kernel. Skip tensors are written
straight into their concat slot by the producing epilogue (channel offset + pixel stride).
"""
import torch

from . import _lib
from ._lib import call, ptr
from .pdl import ACT_NONE, ACT_RELU, ACT_SILU, _NetModel, _PlanBase, _WeightsBase

EPS = 1e-4  # TopDownFPN.eps / BottomUpFPN.eps


def is_bifpn_state_dict(sd):
    return ("semantic_fpn.p6_resample.conv.0.weight" in sd and "p2_resample.conv.0.weight" in sd
            and "semantic_pr.point_head.predictor.weight" in sd)


class _BiFPNWeights(_WeightsBase):
    def __init__(self, sd, device):
        super().__init__(sd, device)
        self.load_encoder()
        self.D = int(sd["p2_resample.conv.0.weight"].shape[0])
        if self.D % 64 != 0 or self.D > 256:
            raise _lib.B200EmpanadaError(f"unsupported fpn_dim {self.D}")
        self._conv_bn("p2_resample", "p2_resample.conv")
        self.branches = ["semantic"] + (["instance"] if "instance_fpn.p6_resample.conv.0.weight" in sd else [])
        self.n_layers = 0
        while f"semantic_fpn.bifpns.{self.n_layers}.top_down_fpn.weights" in sd:
            self.n_layers += 1
        self.fuse_w = {}
        for br in self.branches:
            fpn = br + "_fpn"
            self._conv_bn(fpn + ".p6", fpn + ".p6_resample.conv")
            for li in range(self.n_layers):
                for side in ("top_down_fpn", "bottom_up_fpn"):
                    pre = f"{fpn}.bifpns.{li}.{side}"
                    # fast-fusion weights: relu, then normalised by (sum + eps), all in fp32
                    w = torch.relu(self.f32(sd[pre + ".weights"]))
                    w = w / (w.sum() + EPS)
                    self.fuse_w[pre] = [float(v) for v in w]
                    self.fuse_w[pre + ".t"] = w
                    for i in range(4):
                        if f"{pre}.resamplings.{i}.conv.0.weight" in sd:
                            self._conv_bn(f"{pre}.rs{i}", f"{pre}.resamplings.{i}.conv")
                        self._sepconv(f"{pre}.ac{i}", f"{pre}.after_combines.{i}")
            dec = br + "_decoder"
            self.n_up = 0
            while f"{dec}.upsamplings.{self.n_up}.0.weight" in sd:
                i = self.n_up
                w = self.f32(sd[f"{dec}.upsamplings.{i}.0.weight"])      # [Cin, Cout, 2, 2]
                scale, shift = self.bn_scale_shift(f"{dec}.upsamplings.{i}.1")
                # rows n = (2*dy + dx) * Cout + co
                wg = (w * scale[None, :, None, None]).permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0])
                self.put(f"{dec}.up{i}.w", wg.to(torch.bfloat16))
                self.put(f"{dec}.up{i}.b", shift.repeat(4))
                self.n_up += 1
            self._sepconv(dec + ".fusion", dec + ".fusion")
        self.load_heads()
        self.load_pointrend()
        del self.sd

    def _conv_bn(self, name, key):
        """conv_bn_act(kernel 1, no activation): BatchNorm folded into weight rows + bias."""
        sd, f32 = self.sd, self.f32
        w = f32(sd[key + ".0.weight"])
        scale, shift = self.bn_scale_shift(key + ".1")
        self.put(name + ".w", (w.reshape(w.shape[0], -1) * scale[:, None]).to(torch.bfloat16))
        self.put(name + ".b", shift)


class _BiFPNPlan(_PlanBase):
    def __init__(self, W, B, h, w, H, Wd, mean255, den, render_steps=2, num_points=8192, fused_stem=True, elem=0):
        super().__init__(W, B)
        D = W.D
        levels = self.record_encoder(h, w, H, Wd, mean255, den, 32, fused_stem, elem)
        p2, H4, W4, _ = levels[2]
        self.p2, self.p5 = p2, levels[5][0]
        # P2 skip goes straight into the second half of both decoders' last concat buffers
        cats = {}
        for br in W.branches:
            cats[br] = [self.buf(B, (H >> (6 - i)), (Wd >> (6 - i)), 2 * D) for i in range(W.n_up)]
        first = cats[W.branches[0]][-1]
        self.conv(p2, H4, W4, 256, "p2_resample", D, out=first, out_ld=2 * D, coff=D, act=ACT_NONE)
        self.p2f = first
        for br in W.branches[1:]:
            # same tensor for the other branch: identity "fusion" copy (w1 = 1) into its slot
            self._fuse((first[..., D:], 2 * D, H4, W4), 0, (first[..., D:], 2 * D), None, (1.0, 0.0, 0.0, 1.0),
                       H4, W4, out=cats[br][-1][..., D:], out_ld=2 * D)
        feats = {}
        for br in W.branches:
            pyr = self._bifpn(br + "_fpn", levels, cats[br])
            feats[br] = self._decoder(br + "_decoder", pyr, cats[br], H4, W4)
        semantic_x = feats["semantic"]
        instance_x = feats.get("instance", semantic_x)
        self.record_heads_pointrend(semantic_x, instance_x, H4, W4, D, render_steps, num_points)

    # (tensor, ld, H, W) helpers ---------------------------------------------------------------
    def _fuse(self, a, mode, b, c, ws, H, W_, out=None, out_ld=None):
        D = self.W.D
        if out is None:
            out, out_ld = self.buf(self.B, H, W_, D), D
        w1, w2, w3, denom = ws
        self._rec("be_op_bifpn_fuse", self.handle, ptr(a[0]), a[1], mode, a[2], a[3], ptr(b[0]), b[1],
                  ptr(c[0]) if c is not None else None, c[1] if c is not None else 8, float(w1), float(w2),
                  float(w3), float(denom), self.B, H, W_, D, ptr(out), out_ld, None)
        return out

    def _after_combine(self, name, x, H, W_, out=None, out_ld=None, coff=0):
        """separable_conv_bn_act(D, D, 3, SiLU): depthwise 3x3 -> pointwise (+BN) -> SiLU."""
        D = self.W.D
        dw = self.buf(self.B, H, W_, D)
        self._rec("be_op_dwconv", self.handle, ptr(x), D, self.B, H, W_, D, 3, ptr(self.W[name + ".dw"]), ptr(dw), D,
                  None, 0, 0, 0, None)
        o, _, _ = self.conv(dw, H, W_, D, name + ".pw", D, act=ACT_SILU, out=out, out_ld=out_ld, coff=coff)
        return o

    def _bifpn(self, fpn, levels, cat):
        """BiFPN.forward. Returns [(tensor, ld, H, W)] of P3..P7 after the last layer; those
        outputs are written directly into the decoder's concat buffers `cat` (second half)."""
        W, B, D = self.W, self.B, self.W.D
        p5, H5, W5, C5 = levels[5]
        p6r, _, _ = self.conv(p5, H5, W5, C5, fpn + ".p6", D, act=ACT_NONE)
        H6, W6 = (H5 + 1) // 2, (W5 + 1) // 2
        H7, W7 = (H6 + 1) // 2, (W6 + 1) // 2
        p6 = self.buf(B, H6, W6, D)
        self._rec("be_op_maxpool", self.handle, ptr(p6r), B, H5, W5, D, ptr(p6), H6, W6, None)
        p7 = self.buf(B, H7, W7, D)
        self._rec("be_op_maxpool", self.handle, ptr(p6), B, H6, W6, D, ptr(p7), H7, W7, None)
        # pyramid large -> small: (tensor, ld, H, W, C)
        pyr = [(levels[l][0], levels[l][3], levels[l][1], levels[l][2], levels[l][3]) for l in (3, 4, 5)]
        pyr += [(p6, D, H6, W6, D), (p7, D, H7, W7, D)]
        for li in range(W.n_layers):
            last = li == W.n_layers - 1
            # ---- top-down over [P7, P6, P5, P4, P3]
            tp = f"{fpn}.bifpns.{li}.top_down_fpn"
            w = W.fuse_w[tp]
            rev = pyr[::-1]
            td = [rev[0]]
            for i in range(4):
                t, ld, Hh, Ww, C = rev[i + 1]
                if (tp + f".rs{i}.w") in W.t:
                    t, _, _ = self.conv(t, Hh, Ww, C, tp + f".rs{i}", D, act=ACT_NONE, in_ld=ld)
                    ld = D
                low = td[-1]
                fused = self._fuse((low[0], low[1], low[2], low[3]), 1, (t, ld), None,
                                   (w[i], w[i + 1], 0.0, w[i] + w[i + 1] + EPS), Hh, Ww)
                if last and i == 3:   # top-down P3 is also the bottom-up output P3: decoder skip slot
                    dst = cat[3]
                    o = self._after_combine(tp + f".ac{i}", fused, Hh, Ww, out=dst, out_ld=2 * D, coff=D)
                    td.append((o[..., D:], 2 * D, Hh, Ww, D))
                else:
                    o = self._after_combine(tp + f".ac{i}", fused, Hh, Ww)
                    td.append((o, D, Hh, Ww, D))
            # ---- bottom-up over [P4, P5, P6, P7] with the top-down features large -> small
            bp = f"{fpn}.bifpns.{li}.bottom_up_fpn"
            w = W.fuse_w[bp]
            tdr = td[::-1]
            bu = [tdr[0]]
            for i in range(4):
                t, ld, Hh, Ww, C = pyr[1 + i]
                if (bp + f".rs{i}.w") in W.t:
                    t, _, _ = self.conv(t, Hh, Ww, C, bp + f".rs{i}", D, act=ACT_NONE, in_ld=ld)
                    ld = D
                high = bu[-1]
                if i < 3:
                    tdl = tdr[i + 1]
                    ws = (w[i], w[i + 1], w[i + 2], w[i] + w[i + 1] + w[i + 2] + EPS)
                    fused = self._fuse((high[0], high[1], high[2], high[3]), 2, (t, ld), (tdl[0], tdl[1]), ws, Hh, Ww)
                else:
                    ws = (w[i], w[i + 1], 0.0, w[i] + w[i + 1] + EPS)
                    fused = self._fuse((high[0], high[1], high[2], high[3]), 2, (t, ld), None, ws, Hh, Ww)
                if last and i < 3:    # P4, P5, P6 outputs are decoder skips (cat[2], cat[1], cat[0])
                    dst = cat[2 - i]
                    o = self._after_combine(bp + f".ac{i}", fused, Hh, Ww, out=dst, out_ld=2 * D, coff=D)
                    bu.append((o[..., D:], 2 * D, Hh, Ww, D))
                else:
                    o = self._after_combine(bp + f".ac{i}", fused, Hh, Ww)
                    bu.append((o, D, Hh, Ww, D))
            pyr = bu
        return pyr

    def _decoder(self, dec, pyr, cat, H4, W4):
        """BiFPNDecoder.forward: x = P7; 5 x (ConvTranspose2d k2 s2 + BN + ReLU, concat skip)
        -> separable 5x5 fusion (ReLU)."""
        W, B, D = self.W, self.B, self.W.D
        x, ld, Hh, Ww, C = pyr[-1]
        for i in range(W.n_up):
            dst = cat[i]
            self._rec("be_op_convt2x2", self.handle, ptr(x), ld, B, Hh, Ww, C, ptr(W[f"{dec}.up{i}.w"]), D,
                      ptr(dst), 2 * D, 0, ptr(W[f"{dec}.up{i}.b"]), ACT_RELU, None)
            self.op_info[-1] = ("conv", 2.0 * B * Hh * Ww * 4 * D * C)
            Hh, Ww = 2 * Hh, 2 * Ww
            if tuple(dst.shape[1:3]) != (Hh, Ww):
                raise _lib.B200EmpanadaError("BiFPN decoder: pyramid sizes do not double (pad to a multiple of 128)")
            x, ld, C = dst, 2 * D, 2 * D
        dw = self.buf(B, H4, W4, 2 * D)
        self._rec("be_op_dwconv", self.handle, ptr(x), 2 * D, B, H4, W4, 2 * D, 5, ptr(W[dec + ".fusion.dw"]), ptr(dw),
                  2 * D, None, 0, 0, 0, None)
        o, _, _ = self.conv(dw, H4, W4, 2 * D, dec + ".fusion.pw", D)
        return o


class BiFPNModel(_NetModel):
    weights_cls = _BiFPNWeights
    plan_cls = _BiFPNPlan
    min_factor = 128
