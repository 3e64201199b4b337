"""PanopticDeepLab-PointRend (ResNet-50, output stride 16) on the sm_100a kernels: weight
ingestion from the reference's fused TorchScript export and the recorded launch list.

Network structure follows the reference (file:line under /root/reference/empanada/models):
  encoder            encoders/resnet.py:143-229, quantization/encoders/resnet.py:45-77
  ASPP + decoder     decoders/aspp.py:51-102, decoders/panoptic_deeplab.py:68-80
  heads              heads.py:9-19
  PointRend (eval)   point_rend.py:110-137,241-269
  forward            quantization/panoptic_deeplab.py:194-250 (render_steps=2, interpolate_ins=False)

Layout decisions: NHWC bf16 activations; conv weights [Cout][R*S*Cin] bf16 with
BatchNorm folded in fp32 before rounding; every 1x1/3x3 convolution is the tcgen05 implicit GEMM
(csrc/conv_gemm.cu) with bias/ReLU/residual in the epilogue; concat buffers are written in
place by their producers; the ASPP image-pool branch is a per-image bias of the projection;
the 256->1/2 head convolutions are fused into the epilogue of the preceding pointwise conv.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

P, I, LL, F = c_void_p, c_int, c_longlong, c_float
_lib.declare("be_oplist_create", [ctypes.POINTER(c_void_p)])
_lib.declare("be_oplist_destroy", [P])
_lib.declare("be_oplist_launches", [P])
_lib.declare("be_oplist_run", [P, P, LL, LL, LL, I, P])
_lib.declare("be_oplist_set_graph", [P, I])
_lib.declare("be_oplist_run_timed", [P, P, LL, LL, LL, I, P, I, P])
_lib.declare("be_oplist_size", [P])
_lib.declare("be_op_conv", [P, P, LL, I, I, I, I, P, I, I, I, I, I, I, I, I, P, LL, I, P, LL, I, P, LL,
                            P, LL, I, P, P, P, I, P])
_lib.declare("be_op_stem", [P, I, I, I, I, I, F, F, P, P, P, P, LL, LL, LL, I, I, P])
_lib.declare("be_op_maxpool", [P, P, I, I, I, I, P, I, I, P])
_lib.declare("be_op_stem_pool", [P, I, I, I, I, I, F, F, P, P, P, P, LL, LL, LL, I, I, P])
_lib.declare("be_op_dwconv", [P, P, LL, I, I, I, I, I, P, P, LL, P, I, I, I, P])
_lib.declare("be_op_bilinear", [P, P, LL, I, I, I, I, P, LL, I, I, I, P])
_lib.declare("be_op_convt2x2", [P, P, LL, I, I, I, I, P, I, P, LL, I, P, I, P])
_lib.declare("be_op_bifpn_fuse", [P, P, LL, I, I, I, P, LL, P, LL, F, F, F, F, I, I, I, I, P, LL, P])
_lib.declare("be_op_aspp_pool_bias", [P, P, I, I, I, P, I, P, P, I, P, P, P, P])
_lib.declare("be_op_up2", [P, P, I, I, I, P, P])
_lib.declare("be_op_topk", [P, P, I, I, I, P, P, P, P])
_lib.declare("be_op_pr_sample", [P, P, I, I, I, I, P, P, I, I, I, P, P, I, P, P])
_lib.declare("be_op_pr_predict", [P, P, I, I, P, P, F, P, I, I, I, P, P])

ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2
# decoder `F.interpolate` + `torch.cat` fused into the depthwise kernel (1) or materialised (0)
FUSED_UPSAMPLE = os.environ.get("B200_EMPANADA_FUSED_UPSAMPLE", "0") == "1"
USE_GRAPHS = os.environ.get("B200_EMPANADA_GRAPHS", "1") == "1"
GRAPH_MAX_PIXELS = 4 << 20
BN_EPS = 1e-5


def is_pdl_state_dict(sd):
    return ("encoder.layer4.2.conv3.weight" in sd and "semantic_decoder.aspp.project.0.0.weight" in sd
            and "semantic_pr.point_head.predictor.weight" in sd)


def _conv_w(w):
    """[Cout, Cin, R, S] fp32 -> [Cout][R*S*Cin] bf16 (K-major rows, tap-major K)."""
    return w.permute(0, 2, 3, 1).contiguous().reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()


class _WeightsBase:
    """Device-resident, kernel-layout weights built from the export's state_dict (shared parts:
    fused ResNet-50 encoder, separable-conv heads, PointRend MLP)."""

    def __init__(self, sd, device):
        self.dev = device
        self.sd = sd
        self.t = {}

    @staticmethod
    def f32(x):
        return x.detach().to(torch.float32)

    def put(self, name, tensor):
        self.t[name] = tensor.contiguous().to(self.dev)

    def conv(self, name, key):
        sd = self.sd
        self.put(name + ".w", _conv_w(self.f32(sd[key + ".weight"])))
        if key + ".bias" in sd:
            self.put(name + ".b", self.f32(sd[key + ".bias"]))

    def bn_scale_shift(self, key):
        sd, f32 = self.sd, self.f32
        scale = f32(sd[key + ".weight"]) / torch.sqrt(f32(sd[key + ".running_var"]) + BN_EPS)
        return scale, f32(sd[key + ".bias"]) - f32(sd[key + ".running_mean"]) * scale

    def load_encoder(self):
        sd, f32 = self.sd, self.f32
        w = f32(sd["encoder.conv1.0.weight"])  # [64,1,7,7]
        if w.shape[1] != 1 or tuple(w.shape[2:]) != (7, 7) or w.shape[0] != 64:
            raise _lib.B200EmpanadaError(f"unsupported stem convolution {tuple(w.shape)}")
        self.put("stem.w", w.reshape(64, 49).t())   # [49][64]
        self.put("stem.b", f32(sd["encoder.conv1.0.bias"]))
        self.blocks = []
        for li in range(1, 5):
            b = 0
            while f"encoder.layer{li}.{b}.conv1.0.weight" in sd:
                p = f"encoder.layer{li}.{b}"
                for c, k in (("c1", ".conv1.0"), ("c2", ".conv2.0"), ("c3", ".conv3")):
                    self.conv(f"{p}.{c}", p + k)
                has_ds = (p + ".downsample.0.weight") in sd
                if has_ds:
                    self.conv(p + ".ds", p + ".downsample.0")
                self.blocks.append((li, b, has_ds, sd[p + ".conv1.0.weight"].shape[0]))
                b += 1
        if [sum(1 for x in self.blocks if x[0] == l) for l in (1, 2, 3, 4)] != [3, 4, 6, 3]:
            raise _lib.B200EmpanadaError("encoder is not a ResNet-50 (unsupported export)")

    def load_heads(self):
        sd, f32 = self.sd, self.f32
        for head in ("semantic_head", "ins_center", "ins_xy"):
            self._sepconv(head + ".sep", head + ".head.0")
            self.put(head + ".out.w", f32(sd[head + ".head.1.weight"]).reshape(sd[head + ".head.1.weight"].shape[0], -1))
            self.put(head + ".out.b", f32(sd[head + ".head.1.bias"]))
        if sd["semantic_head.head.1.weight"].shape[0] != 1:
            raise _lib.B200EmpanadaError("multi-class semantic heads are not built yet")

    def load_pointrend(self):
        """PointRend MLP: K padded C+1 -> C+8 (16-byte rows for TMA). Fused exports (PDL) nest the
        conv one level deeper than unfused ones (BiFPN)."""
        sd, f32 = self.sd, self.f32
        fmt = "semantic_pr.point_head.fc_layers.{}.0.0"
        if fmt.format(0) + ".weight" not in sd:
            fmt = "semantic_pr.point_head.fc_layers.{}.0"
        self.num_fc = 0
        while fmt.format(self.num_fc) + ".weight" in sd:
            w = f32(sd[fmt.format(self.num_fc) + ".weight"])[:, :, 0]  # [C, C+1]
            wpad = torch.zeros(w.shape[0], w.shape[0] + 8)
            if w.shape[1] != w.shape[0] + 1:
                raise _lib.B200EmpanadaError("unsupported PointRend MLP shape")
            wpad[:, : w.shape[1]] = w
            self.put(f"pr.fc{self.num_fc}.w", wpad.to(torch.bfloat16))
            self.put(f"pr.fc{self.num_fc}.b", f32(sd[fmt.format(self.num_fc) + ".bias"]))
            self.num_fc += 1
        self.put("pr.pred.w", f32(sd["semantic_pr.point_head.predictor.weight"]).reshape(-1))  # [C+1]
        self.pr_pred_b = float(sd["semantic_pr.point_head.predictor.bias"].reshape(-1)[0])

    def _sepconv(self, name, key):
        """separable_conv_bn_act (blocks.py): depthwise k x k, pointwise 1x1, live BatchNorm
        folded into the pointwise weights in fp32."""
        sd, f32 = self.sd, self.f32
        dw = f32(sd[key + ".0.sepconv.0.weight"])          # [C,1,k,k]
        C, k = dw.shape[0], dw.shape[-1]
        self.put(name + ".dw", dw.reshape(C, k * k).t())        # [k*k][C]
        pw = f32(sd[key + ".0.sepconv.1.weight"]).reshape(-1, C)
        scale, shift = self.bn_scale_shift(key + ".1")
        self.put(name + ".pw.w", (pw * scale[:, None]).to(torch.bfloat16))
        self.put(name + ".pw.b", shift)

    def __getitem__(self, k):
        return self.t[k]


class _Weights(_WeightsBase):
    """PanopticDeepLab-PointRend export."""

    def __init__(self, sd, device):
        super().__init__(sd, device)
        f32, put, conv = self.f32, self.put, self.conv
        self.load_encoder()
        # decoders
        self.has_ins_decoder = "instance_decoder.aspp.project.0.0.weight" in sd
        for dec in ("semantic_decoder", "instance_decoder") if self.has_ins_decoder else ("semantic_decoder",):
            for i in range(4):
                conv(f"{dec}.aspp{i}", f"{dec}.aspp.convs.{i}.0.0")
            put(dec + ".pool.w", f32(sd[dec + ".aspp.convs.4.aspp_pooling.1.0.weight"]).reshape(-1, 2048))
            wp = f32(sd[dec + ".aspp.project.0.0.weight"]).reshape(256, -1)  # [256, 1280]
            n_main = wp.shape[1] - self.t[dec + ".pool.w"].shape[0]
            put(dec + ".proj.w", wp[:, :n_main].to(torch.bfloat16))
            put(dec + ".proj.wpool", wp[:, n_main:])
            put(dec + ".proj.b", f32(sd[dec + ".aspp.project.0.0.bias"]))
            conv(dec + ".low", dec + ".project.0.0.0")
            self._sepconv(dec + ".fuse", dec + ".fuse.0")
        self.load_heads()
        self.load_pointrend()
        del self.sd


class _PlanBase:
    """Recorded launch list + its activation buffers for one (B, h, w, H, W). Subclasses record
    the decoder between `record_encoder` and `record_heads_pointrend`."""

    def __init__(self, W, B):
        self.W, self.B, self.dev = W, B, W.dev
        self.handle = c_void_p()
        _lib.lib().be_oplist_create(ctypes.byref(self.handle))
        self.bufs = []
        self.op_info = []  # (kind, flops) per recorded op, same order as the launch list
        self.op_desc = []

    def buf(self, *shape, dtype=torch.bfloat16):
        t = torch.empty(shape, dtype=dtype, device=self.dev)
        self.bufs.append(t)
        return t

    def conv(self, x, Hi, Wi, Cin, wname, Cout, k=1, stride=1, dil=1, out=None, out_ld=None, coff=0,
             act=ACT_RELU, res=None, bias=None, bias_img_stride=0, in_ld=None, head=None,
             Bn=None, alg_cin=None):
        W, L = self.W, self.handle
        Bn = self.B if Bn is None else Bn
        pad = dil * (k - 1) // 2
        Ho = (Hi + 2 * pad - dil * (k - 1) - 1) // stride + 1
        Wo = (Wi + 2 * pad - dil * (k - 1) - 1) // stride + 1
        if out is None and head is None:
            out = self.buf(Bn, Ho, Wo, Cout)
        if out_ld is None:
            out_ld = Cout
        if bias is None and (wname + ".b") in W.t:
            bias = W[wname + ".b"]
        hw_, hb_, ho_, hn_ = (None, None, None, 0) if head is None else head
        call("be_op_conv", L, ptr(x), in_ld or Cin, Bn, Hi, Wi, Cin, ptr(W[wname + ".w"]), Cout, k, k,
             stride, dil, pad, Ho, Wo, ptr(out), out_ld, coff, None, 0, 0, ptr(bias),
             bias_img_stride, ptr(res), Cout if res is not None else 0, act, ptr(hw_), ptr(hb_),
             ptr(ho_), hn_, None)
        self.op_info.append(("conv", 2.0 * Bn * Ho * Wo * Cout * k * k * (alg_cin or Cin)))
        nbytes = 2.0 * Bn * (Hi * Wi * Cin + (Ho * Wo * Cout if out is not None else 0) + (Ho * Wo * Cout if res is not None else 0)) + 2.0 * Cout * k * k * Cin
        self.op_desc.append(f"{wname} {Cin}->{Cout} k{k} s{stride} d{dil} {Hi}x{Wi} bytes={nbytes/1e6:.0f}MB")
        return out, Ho, Wo

    def record_encoder(self, h, w, H, Wd, mean255, den, output_stride=16, fused_stem=True, elem=0):
        """Stem + max-pool + the 16 bottlenecks. Returns {level: (tensor, H, W, C)} for the
        outputs of layer1..layer4 (levels 2..5)."""
        W, L, B = self.W, self.handle, self.B
        H2, W2, H4, W4 = H // 2, Wd // 2, H // 4, Wd // 4
        x = self.buf(B, H4, W4, 64)
        if fused_stem:
            self._rec("be_op_stem_pool", L, B, h, w, H, Wd, mean255, den, ptr(W["stem.w"]), ptr(W["stem.b"]),
                      ptr(x), None, 0, 0, 0, 0, elem, None)
        else:
            stem = self.buf(B, H2, W2, 64)
            self._rec("be_op_stem", L, B, h, w, H, Wd, mean255, den, ptr(W["stem.w"]), ptr(W["stem.b"]),
                      ptr(stem), None, 0, 0, 0, 0, elem, None)
            self._rec("be_op_maxpool", L, ptr(stem), B, H2, W2, 64, ptr(x), H4, W4, None)
        Hc, Wc, Cin = H4, W4, 64
        levels = {}
        for (li, b, has_ds, planes) in W.blocks:
            pfx = f"encoder.layer{li}.{b}"
            if output_stride == 16:
                stride = 2 if (b == 0 and li in (2, 3)) else 1
                dil = 2 if li == 4 else 1
            else:
                stride = 2 if (b == 0 and li in (2, 3, 4)) else 1
                dil = 1
            t1, _, _ = self.conv(x, Hc, Wc, Cin, pfx + ".c1", planes)
            t2, Ho, Wo = self.conv(t1, Hc, Wc, planes, pfx + ".c2", planes, k=3, stride=stride, dil=dil)
            if has_ds:
                idt, _, _ = self.conv(x, Hc, Wc, Cin, pfx + ".ds", planes * 4, stride=stride, act=ACT_NONE)
            else:
                idt = x
            x, _, _ = self.conv(t2, Ho, Wo, planes, pfx + ".c3", planes * 4, res=idt)
            Hc, Wc, Cin = Ho, Wo, planes * 4
            levels[li + 1] = (x, Hc, Wc, Cin)
        return levels

    def record_heads_pointrend(self, semantic_x, instance_x, H4, W4, D, render_steps, num_points):
        """The three PanopticDeepLabHeads (heads.py:9-19; the 1x1 output convs are fused into the
        pointwise epilogue) and the PointRend refinement (point_rend.py:110-137,241-269) on
        D-channel decoder features."""
        W, L, B = self.W, self.handle, self.B
        self.coarse = self.buf(B, 1, H4, W4, dtype=torch.float32)
        self.ctr = self.buf(B, H4, W4, dtype=torch.float32)
        self.off = self.buf(B, 2, H4, W4, dtype=torch.float32)
        for head, src, out, n in (("semantic_head", semantic_x, self.coarse, 1),
                                  ("ins_center", instance_x, self.ctr, 1), ("ins_xy", instance_x, self.off, 2)):
            dw = self.buf(B, H4, W4, D)
            self._rec("be_op_dwconv", L, ptr(src), D, B, H4, W4, D, 5, ptr(W[head + ".sep.dw"]), ptr(dw), D,
                      None, 0, 0, 0, None)
            self.conv(dw, H4, W4, D, head + ".sep.pw", D, head=(W[head + ".out.w"], W[head + ".out.b"], out, n))
        sem, Hs, Ws = self.coarse, H4, W4
        ldp = D + 8
        for step in range(render_steps):
            up = self.buf(B, 2 * Hs, 2 * Ws, dtype=torch.float32)
            self._rec("be_op_up2", L, ptr(sem), B, Hs, Ws, ptr(up), None)
            Hs, Ws = 2 * Hs, 2 * Ws
            k = min(Hs * Ws, num_points)
            state = self.buf(B, 8, dtype=torch.int32)
            hist = self.buf(B, 3, 2048, dtype=torch.int32)
            idx = self.buf(B, k, dtype=torch.int32)
            self._rec("be_op_topk", L, ptr(up), B, Hs * Ws, k, ptr(state), ptr(hist), ptr(idx), None)
            Pa, Pb = self.buf(B * k, ldp), self.buf(B * k, ldp)
            cpts = self.buf(B * k, dtype=torch.float32)
            self._rec("be_op_pr_sample", L, ptr(idx), B, k, Hs, Ws, ptr(self.coarse), ptr(semantic_x), H4, W4, D,
                      ptr(Pa), ptr(Pb), ldp, ptr(cpts), None)
            src, dst = Pa, Pb
            for l in range(W.num_fc):
                self.conv(src, 1, B * k, ldp, f"pr.fc{l}", D, out=dst, out_ld=ldp, in_ld=ldp, Bn=1, alg_cin=D + 1)
                src, dst = dst, src
            self._rec("be_op_pr_predict", L, ptr(src), ldp, D, ptr(cpts), ptr(W["pr.pred.w"]), W.pr_pred_b,
                      ptr(idx), B, k, Hs * Ws, ptr(up), None)
            sem = up
        self.sem = sem.view(B, Hs, Ws)
        self.semantic_x, self.instance_x = semantic_x, instance_x
        self.launches = int(_lib.lib().be_oplist_launches(L))

    def _rec(self, name, *args):
        call(name, *args)
        self.op_info.append((name[len("be_op_"):], 0.0))
        self.op_desc.append(name[len("be_op_"):])

    def run_timed(self, vol_d, strides, s0):
        """Per-op device times (ms) of one replay, measured with CUDA events between ops."""
        n = int(_lib.lib().be_oplist_size(self.handle))
        ms = np.zeros(n, dtype=np.float32)
        call("be_oplist_run_timed", self.handle, ptr(vol_d), strides[0], strides[1], strides[2], s0,
             ptr(ms), n, stream_ptr())
        assert n == len(self.op_info)
        return ms

    def run(self, vol_d, strides, s0):
        call("be_oplist_run", self.handle, ptr(vol_d), strides[0], strides[1], strides[2], s0, stream_ptr())

    def __del__(self):
        try:
            _lib.lib().be_oplist_destroy(self.handle)
        except Exception:
            pass


class _Plan(_PlanBase):
    """PanopticDeepLab-PointRend launch list."""

    def __init__(self, W, B, h, w, H, Wd, mean255, den, render_steps=2, num_points=8192, fused_stem=True, elem=0):
        super().__init__(W, B)
        L, buf, conv = self.handle, self.buf, self.conv
        H4, W4 = H // 4, Wd // 4
        levels = self.record_encoder(h, w, H, Wd, mean255, den, 16, fused_stem, elem)
        p2 = levels[2][0]
        p5, H16, W16, _ = levels[5]
        feats = {}
        decs = ("semantic_decoder", "instance_decoder") if W.has_ins_decoder else ("semantic_decoder",)
        for dec in decs:
            cat = buf(B, H16, W16, 1024)
            for i, r in enumerate((1, 2, 4, 6)):
                conv(p5, H16, W16, 2048, f"{dec}.aspp{i}", 256, k=1 if i == 0 else 3, dil=1 if i == 0 else r,
                     out=cat, out_ld=1024, coff=256 * i)
            pooled, mid, pbias = buf(B, 2048, dtype=torch.float32), buf(B, 256, dtype=torch.float32), buf(B, 256, dtype=torch.float32)
            self._rec("be_op_aspp_pool_bias", L, ptr(p5), B, H16 * W16, 2048, ptr(W[dec + ".pool.w"]), 256,
                 ptr(W[dec + ".proj.wpool"]), ptr(W[dec + ".proj.b"]), 256, ptr(pooled), ptr(mid), ptr(pbias), None)
            aspp, _, _ = conv(cat, H16, W16, 1024, dec + ".proj", 256, bias=pbias, bias_img_stride=256)
            clow = W[dec + ".low.w"].shape[0]
            cf = 256 + clow
            dw = buf(B, H4, W4, cf)
            if FUSED_UPSAMPLE:
                # bilinear upsampling + concat fused into the depthwise producer
                low, _, _ = conv(p2, H4, W4, 256, dec + ".low", clow)
                self._rec("be_op_dwconv", L, ptr(low), clow, B, H4, W4, cf, 5, ptr(W[dec + ".fuse.dw"]), ptr(dw), cf,
                          ptr(aspp), 256, H16, W16, None)
            else:
                # concat buffer materialised once (bilinear producer + low-level projection write
                # their channel slots), then the TMA-fed depthwise kernel
                cat4 = buf(B, H4, W4, cf)
                self._rec("be_op_bilinear", L, ptr(aspp), 256, B, H16, W16, 256, ptr(cat4), cf, 0, H4, W4, None)
                conv(p2, H4, W4, 256, dec + ".low", clow, out=cat4, out_ld=cf, coff=256)
                self._rec("be_op_dwconv", L, ptr(cat4), cf, B, H4, W4, cf, 5, ptr(W[dec + ".fuse.dw"]), ptr(dw), cf,
                          None, 0, 0, 0, None)
            feats[dec], _, _ = conv(dw, H4, W4, cf, dec + ".fuse.pw", 256)
        semantic_x = feats["semantic_decoder"]
        instance_x = feats.get("instance_decoder", semantic_x)
        self.p5, self.p2 = p5, p2
        self.record_heads_pointrend(semantic_x, instance_x, H4, W4, 256, render_steps, num_points)


# element type codes of the C-ABI (include/b200_empanada.h be_op_stem): any integer dtype the
# reference's Preprocessor accepts (empanada_napari/utils.py:189-201)
ELEM_CODES = {"uint8": 0, "int8": 1, "uint16": 2, "int16": 3, "uint32": 4, "int32": 5, "uint64": 6, "int64": 7,
              "float32": 8}   # 8: image already normalised (engine-level API, norms=None)


def elem_code(dtype):
    name = str(dtype).replace("torch.", "")
    if name not in ELEM_CODES:
        raise _lib.B200EmpanadaError(f"unsupported volume dtype {dtype} (integer dtypes only)")
    return ELEM_CODES[name]


def norm_constants(norms, dtype):
    """fp32 (mean * max, 1 / (std * max)) exactly as `normalize` computes them
    (empanada_napari/utils.py:170-185) with max = np.iinfo(dtype).max."""
    name = str(dtype).replace("torch.", "")
    if name == "float32" or norms is None:
        if name != "float32" or norms is not None:
            raise _lib.B200EmpanadaError("float volumes are only accepted already normalised (norms=None)")
        return np.float32(0.0), np.float32(1.0)
    maxv = np.float32(np.iinfo(np.dtype(name)).max)
    mean = np.float32(np.float32(norms["mean"]) * maxv)
    den = np.reciprocal(np.float32(np.float32(norms["std"]) * maxv), dtype=np.float32)
    return mean, den


class _NetModel:
    """A network = kernel-layout weights + one recorded launch list per (batch, slice shape).
    Launch lists own their activation buffers, so the cache is bounded: least-recently-used lists
    are dropped beyond `max_plans` entries or `max_plan_bytes` of activations, and a shorter tail
    batch re-uses the full-batch list of the same slice shape (shifted back over slices that were
    already computed) instead of recording a second one."""
    weights_cls = None
    plan_cls = None
    min_factor = 16      # the padded slice size must be a multiple of this
    max_plans = 4

    def __init__(self, sd, device):
        self.dev = device
        with torch.cuda.device(device):
            self.W = self.weights_cls(sd, device)
        self.plans = {}          # insertion order = recency (re-inserted on use)
        self.launches = 0
        self.render_steps = 2
        total = torch.cuda.get_device_properties(device).total_memory
        self.max_plan_bytes = int(0.45 * total)

    @staticmethod
    def _plan_bytes(plan):
        return sum(t.numel() * t.element_size() for t in plan.bufs)

    def _evict(self, need_bytes):
        while self.plans and (len(self.plans) >= self.max_plans or
                              sum(self._plan_bytes(p) for p in self.plans.values()) + need_bytes > self.max_plan_bytes):
            oldest = next(iter(self.plans))
            del self.plans[oldest]

    def forward_slices(self, vol_d, axis, s0, s1, norms, pf, render_steps=None, plane_axis=None):
        """Slices [s0, s1) of the (D,H,W) integer device volume along `axis` -> (sem_logits (B,H,W),
        ctr_hmp (B,H/4,W/4), offsets (B,2,H/4,W/4)) fp32 device tensors owned by the plan. With
        `render_steps` = 2 + k PointRend steps the semantic map is (B, 2^k H, 2^k W)."""
        render_steps = self.render_steps if render_steps is None else int(render_steps)
        D, Hv, Wv = vol_d.shape
        h, w = [(Hv, Wv), (D, Wv), (D, Hv)][axis]
        strides = [(Hv * Wv, Wv, 1), (Wv, Hv * Wv, 1), (1, Hv * Wv, Wv)][axis]
        H = h + (pf - h % pf) % pf
        Wd = w + (pf - w % pf) % pf
        if H % self.min_factor or Wd % self.min_factor:
            raise _lib.B200EmpanadaError(f"padded slice size must be a multiple of {self.min_factor}")
        B = s1 - s0
        elem = elem_code(vol_d.dtype)
        mean255, den = norm_constants(norms, vol_d.dtype)
        shape_key = (h, w, H, Wd, float(mean255), float(den), elem, render_steps)
        key = (B,) + shape_key
        plan = self.plans.pop(key, None)
        skip = 0
        if plan is None:
            # tail batch: replay a longer list of the same slice shape over [s1 - Bp, s1) when that
            # recomputes little (at most a third more slices); a short tail gets its own list -
            # with G ranks a plane has G tails, and replaying the full list for each would waste
            # up to a whole batch per rank (+15 % forward time at 8 ranks on 1024 slices)
            for k2 in list(self.plans.keys()):
                if k2[1:] == shape_key and k2[0] > B and s1 - k2[0] >= 0 and 3 * (k2[0] - B) <= k2[0]:
                    key, plan = k2, self.plans.pop(k2)
                    skip = k2[0] - B
                    break
        if plan is None:
            self._evict(0)
            with torch.cuda.device(self.dev):
                plan = self.plan_cls(self.W, B, h, w, H, Wd, float(mean255), float(den), render_steps, elem=elem)
            # small batches are launch bound (hundreds of kernels of a few microseconds): replay
            # them as one CUDA graph; large batches keep plain launches (kernels >> launch cost)
            if USE_GRAPHS and B * H * Wd <= GRAPH_MAX_PIXELS:
                call("be_oplist_set_graph", plan.handle, 1)
        self.plans[key] = plan
        plan.run(vol_d, strides, s0 - skip)
        self.launches += plan.launches
        self.last_plan = plan
        return plan.sem[skip:], plan.ctr[skip:], plan.off[skip:]

    def release_plans(self):
        """Drop every recorded launch list and its activation buffers."""
        self.plans.clear()


class PDLModel(_NetModel):
    weights_cls = _Weights
    plan_cls = _Plan
    min_factor = 16
