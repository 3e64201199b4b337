"""Seeded synthetic inputs for tests and benchmarks (no file or network I/O).

* `make_volume`   : EM-like uint8 volume of dark ellipsoids on a bright noisy background
                     plus the ellipsoid table.
* `analytic_heads`: head maps a perfect network would emit for one slice of that volume
                    (sem logit +-4, Gaussian centre heatmap at /4, exact offsets at /4).
* `make_pdl_state_dict`: random weights with the key names / shapes of the reference's fused
                    PanopticDeepLab-PointRend TorchScript export (what
                    `torch.jit.load(path).state_dict()` returns; empanada_napari/_train.py:59-73).
"""
import math

import numpy as np


def make_ellipsoids(shape, n_objects=None, seed=0, scale=None):
    d, h, w = shape
    rng = np.random.default_rng(seed)
    if scale is None:
        scale = max(1.0, min(shape) / 256.0)
    if n_objects is None:
        n_objects = max(1, int(round(24 * (d * h * w) / (64.0 ** 3 * scale ** 3))))
    centers = rng.uniform([0, 0, 0], [d, h, w], size=(n_objects, 3))
    radii = rng.uniform(4.0, 12.0, size=(n_objects, 3)) * scale
    return np.concatenate([centers, radii], axis=1).astype(np.float32)  # (n, 6): cz cy cx rz ry rx


def make_separated_ellipsoids(shape, n_objects, seed=0, rmin=5.0, rmax=9.0, gap=5.0):
    """Ellipsoid table (as `make_ellipsoids`) whose objects keep at least `gap` voxels between
    their bounding spheres and stay inside the volume: every object is one well-separated
    instance (used by the end-to-end agreement tests)."""
    d, h, w = shape
    rng = np.random.default_rng(seed)
    out = []
    tries = 0
    while len(out) < n_objects and tries < 200 * n_objects:
        tries += 1
        r = rng.uniform(rmin, rmax, size=3)
        c = rng.uniform(r + 1.0, np.array([d, h, w]) - r - 1.0)
        if all(np.linalg.norm(c - o[:3]) >= r.max() + o[3:].max() + gap for o in out):
            out.append(np.concatenate([c, r]))
    return np.array(out, dtype=np.float32)


def make_separated_volume(shape, n_objects, seed=0, **kw):
    """(uint8 volume, int32 labels, ellipsoid table) of `make_separated_ellipsoids`."""
    ell = make_separated_ellipsoids(shape, n_objects, seed, **kw)
    lab = label_volume(shape, ell)
    rng = np.random.default_rng(seed + 1)
    img = np.where(lab > 0, 70.0, 170.0) + rng.normal(0.0, 8.0, size=shape)
    return np.clip(img, 0, 255).astype(np.uint8), lab, ell


def label_volume(shape, ell):
    """Dense int32 ground truth: voxel -> 1-based index of the LAST ellipsoid containing it."""
    d, h, w = shape
    vol = np.zeros(shape, dtype=np.int32)
    for i, (cz, cy, cx, rz, ry, rx) in enumerate(ell, start=1):
        z0, z1 = max(0, int(math.floor(cz - rz))), min(d, int(math.ceil(cz + rz)) + 1)
        y0, y1 = max(0, int(math.floor(cy - ry))), min(h, int(math.ceil(cy + ry)) + 1)
        x0, x1 = max(0, int(math.floor(cx - rx))), min(w, int(math.ceil(cx + rx)) + 1)
        if z0 >= z1 or y0 >= y1 or x0 >= x1:
            continue
        zz, yy, xx = np.ogrid[z0:z1, y0:y1, x0:x1]
        m = ((zz - cz) / rz) ** 2 + ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        sub = vol[z0:z1, y0:y1, x0:x1]
        sub[m] = i
    return vol


def make_volume(shape, seed=0, n_objects=None, scale=None):
    """Returns (uint8 volume, int32 label volume, ellipsoid table)."""
    ell = make_ellipsoids(shape, n_objects, seed, scale)
    lab = label_volume(shape, ell)
    rng = np.random.default_rng(seed + 1)
    img = np.where(lab > 0, 70.0, 170.0) + rng.normal(0.0, 8.0, size=shape)
    return np.clip(img, 0, 255).astype(np.uint8), lab, ell


def analytic_heads(label_slice, pad_to=None, sigma=8.0, step=4):
    """Head maps for one (h, w) slice of the label volume.

    Returns sem_logits (1, H, W) fp32, ctr_hmp (H/4, W/4) fp32, offsets (2, H/4, W/4) fp32 for
    the padded size H x W (`pad_to` = padding factor)."""
    h, w = label_slice.shape
    H, W = h, w
    if pad_to:
        H = h + (pad_to - h % pad_to) % pad_to
        W = w + (pad_to - w % pad_to) % pad_to
    lab = np.zeros((H, W), dtype=np.int32)
    lab[:h, :w] = label_slice
    sem = np.where(lab > 0, 4.0, -4.0).astype(np.float32)[None]
    h4, w4 = H // step, W // step
    ctr = np.zeros((h4, w4), dtype=np.float32)
    off = np.zeros((2, h4, w4), dtype=np.float32)
    ys4 = (np.arange(h4) * step).astype(np.float32)
    xs4 = (np.arange(w4) * step).astype(np.float32)
    lab4 = lab[::step, ::step]
    ids = np.unique(lab)
    for i in ids[ids > 0]:
        yy, xx = np.nonzero(lab == i)
        cy, cx = np.float32(yy.mean()), np.float32(xx.mean())
        # snap the peak to the /4 grid so the heat-map maximum is unique
        py, px = int(round(float(cy) / step)), int(round(float(cx) / step))
        py, px = min(max(py, 0), h4 - 1), min(max(px, 0), w4 - 1)
        y0, y1 = max(0, py - 12), min(h4, py + 13)
        x0, x1 = max(0, px - 12), min(w4, px + 13)
        gy = (np.arange(y0, y1) - py).astype(np.float32)[:, None]
        gx = (np.arange(x0, x1) - px).astype(np.float32)[None, :]
        g = np.exp(-(gy * gy + gx * gx) * (step * step) / (2.0 * sigma * sigma)).astype(np.float32)
        ctr[y0:y1, x0:x1] = np.maximum(ctr[y0:y1, x0:x1], g)
        m = lab4 == i
        off[0][m] = (np.float32(py * step) - ys4[:, None].repeat(w4, 1))[m]
        off[1][m] = (np.float32(px * step) - xs4[None, :].repeat(h4, 0))[m]
    return sem, ctr, off


# --------------------------------------------------------------------------- random weights
def _pdl_shapes(num_classes=1, decoder_channels=256, low_proj=32, ins_ratio=0.5, num_fc=3):
    """(key, shape, kind) in export order for PanopticDeepLabPR / resnet50 / stride 16."""
    out = []

    def conv(name, co, ci, k, bias=True, kind="conv"):
        out.append((name + ".weight", (co, ci, k, k), kind))
        if bias:
            out.append((name + ".bias", (co,), "bias"))

    conv("encoder.conv1.0", 64, 1, 7)
    inpl = 64
    for li, (planes, blocks) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)], start=1):
        for b in range(blocks):
            p = f"encoder.layer{li}.{b}"
            conv(p + ".conv1.0", planes, inpl, 1)
            conv(p + ".conv2.0", planes, planes, 3)
            conv(p + ".conv3", planes * 4, planes, 1, kind="conv_res")
            if b == 0:
                conv(p + ".downsample.0", planes * 4, inpl, 1, kind="conv_res")
            inpl = planes * 4

    def bn(name, c):
        out.append((name + ".weight", (c,), "bn_w"))
        out.append((name + ".bias", (c,), "bn_b"))
        out.append((name + ".running_mean", (c,), "bn_m"))
        out.append((name + ".running_var", (c,), "bn_v"))
        out.append((name + ".num_batches_tracked", (), "bn_n"))

    for dec, proj in (("semantic_decoder", low_proj), ("instance_decoder", int(low_proj * ins_ratio))):
        conv(dec + ".aspp.convs.0.0.0", decoder_channels, 2048, 1)
        for i in (1, 2, 3):
            conv(f"{dec}.aspp.convs.{i}.0.0", decoder_channels, 2048, 3)
        conv(dec + ".aspp.convs.4.aspp_pooling.1.0", decoder_channels, 2048, 1, bias=False)
        conv(dec + ".aspp.project.0.0", decoder_channels, 5 * decoder_channels, 1)
        conv(dec + ".project.0.0.0", proj, 256, 1)
        cin = decoder_channels + proj
        out.append((dec + ".fuse.0.0.sepconv.0.weight", (cin, 1, 5, 5), "dw"))
        out.append((dec + ".fuse.0.0.sepconv.1.weight", (decoder_channels, cin, 1, 1), "conv"))
        bn(dec + ".fuse.0.1", decoder_channels)
    for head, nout in (("semantic_head", num_classes), ("ins_center", 1), ("ins_xy", 2)):
        out.append((head + ".head.0.0.sepconv.0.weight", (decoder_channels, 1, 5, 5), "dw"))
        out.append((head + ".head.0.0.sepconv.1.weight", (decoder_channels, decoder_channels, 1, 1), "conv"))
        bn(head + ".head.0.1", decoder_channels)
        out.append((head + ".head.1.weight", (nout, decoder_channels, 1, 1), "conv_out"))
        out.append((head + ".head.1.bias", (nout,), "bias"))
    for l in range(num_fc):
        out.append((f"semantic_pr.point_head.fc_layers.{l}.0.0.weight",
                    (decoder_channels, decoder_channels + num_classes, 1), "fc"))
        out.append((f"semantic_pr.point_head.fc_layers.{l}.0.0.bias", (decoder_channels,), "bias"))
    out.append(("semantic_pr.point_head.predictor.weight", (num_classes, decoder_channels + num_classes, 1), "fc_out"))
    out.append(("semantic_pr.point_head.predictor.bias", (num_classes,), "bias"))
    return out


def _bifpn_shapes(num_classes=1, fpn_dim=128, fpn_layers=3, num_fc=3):
    """(key, shape, kind) in export order for PanopticBiFPNPR / resnet50 (output stride 32),
    `ins_decoder=True`, `depthwise=True` (empanada_napari/training/bifpn_model.yaml). The encoder
    is exported fused (conv+BN folded), everything else keeps live BatchNorm."""
    out = [t for t in _pdl_shapes() if t[0].startswith("encoder.")]
    d = fpn_dim

    def bn(name, c):
        out.append((name + ".weight", (c,), "bn_w"))
        out.append((name + ".bias", (c,), "bn_b"))
        out.append((name + ".running_mean", (c,), "bn_m"))
        out.append((name + ".running_var", (c,), "bn_v"))
        out.append((name + ".num_batches_tracked", (), "bn_n"))

    def conv_bn(name, co, ci):
        out.append((name + ".0.weight", (co, ci, 1, 1), "conv_lin"))
        bn(name + ".1", co)

    def sep(name, ci, co, k):
        out.append((name + ".0.sepconv.0.weight", (ci, 1, k, k), "dw"))
        out.append((name + ".0.sepconv.1.weight", (co, ci, 1, 1), "conv"))
        bn(name + ".1", co)

    conv_bn("p2_resample.conv", d, 256)
    for branch in ("semantic", "instance"):
        fpn = branch + "_fpn"
        conv_bn(fpn + ".p6_resample.conv", d, 2048)
        for li in range(fpn_layers):
            for side, nins in (("top_down_fpn", [d, 2048, 1024, 512]), ("bottom_up_fpn", [1024, 2048, d, d])):
                pre = f"{fpn}.bifpns.{li}.{side}"
                out.append((pre + ".weights", (5,), "fuse_w"))
                if li == 0:
                    for i, nin in enumerate(nins):
                        if nin != d:
                            conv_bn(f"{pre}.resamplings.{i}.conv", d, nin)
                for i in range(4):  # one shared block, exported under four aliases
                    sep(f"{pre}.after_combines.{i}", d, d, 3)
        dec = branch + "_decoder"
        for i in range(5):
            out.append((f"{dec}.upsamplings.{i}.0.weight", (d if i == 0 else 2 * d, d, 2, 2), "convT"))
            bn(f"{dec}.upsamplings.{i}.1", d)
        sep(dec + ".fusion", 2 * d, d, 5)
    for head, nout in (("semantic_head", num_classes), ("ins_center", 1), ("ins_xy", 2)):
        sep(head + ".head.0", d, d, 5)
        out.append((head + ".head.1.weight", (nout, d, 1, 1), "conv_out"))
        out.append((head + ".head.1.bias", (nout,), "bias"))
    for l in range(num_fc):
        out.append((f"semantic_pr.point_head.fc_layers.{l}.0.weight", (d, d + num_classes, 1), "fc"))
        out.append((f"semantic_pr.point_head.fc_layers.{l}.0.bias", (d,), "bias"))
    out.append(("semantic_pr.point_head.predictor.weight", (num_classes, d + num_classes, 1), "fc_out"))
    out.append(("semantic_pr.point_head.predictor.bias", (num_classes,), "bias"))
    return out


def make_bifpn_state_dict(seed=0, num_classes=1):
    """Random fp32 state_dict of the PanopticBiFPN-PointRend export with O(1) activations. The
    four `after_combines.{i}` aliases of one shared block hold identical tensors, as in a real
    export (decoders/bifpn.py:34-42,90-98)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in _bifpn_shapes(num_classes=num_classes):
        if ".after_combines." in key and ".after_combines.0." not in key:
            i = key.index(".after_combines.") + len(".after_combines.")
            sd[key] = sd[key[:i] + "0" + key[i + 1:]].clone()
            continue
        if kind in ("conv", "conv_res", "conv_out", "conv_lin", "dw", "fc", "fc_out", "convT"):
            fan_in = int(np.prod(shape[1:])) if kind != "convT" else shape[0]
            gain = {"conv": math.sqrt(2.0), "conv_res": 0.5, "conv_out": 1.0, "conv_lin": 1.0,
                    "dw": math.sqrt(2.0), "fc": math.sqrt(2.0), "fc_out": 1.0, "convT": math.sqrt(2.0)}[kind]
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "bn_w":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif kind in ("bn_b", "bn_m"):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_v":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif kind == "fuse_w":
            t = torch.rand(shape, generator=g) + 0.5
            t[int(torch.randint(0, shape[0], (1,), generator=g))] -= 0.2
        else:
            t = torch.tensor(0, dtype=torch.long)
        sd[key] = t
    return sd


def make_pdl_state_dict(seed=0, num_classes=1):
    """Random fp32 state_dict (torch tensors) with O(1) activations through the whole net."""
    import torch

    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in _pdl_shapes(num_classes=num_classes):
        if kind in ("conv", "conv_res", "conv_out", "dw", "fc", "fc_out"):
            fan_in = int(np.prod(shape[1:]))
            gain = {"conv": math.sqrt(2.0), "conv_res": 0.5, "conv_out": 1.0, "dw": math.sqrt(2.0),
                    "fc": math.sqrt(2.0), "fc_out": 1.0}[kind]
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "bn_w":
            t = torch.rand(shape, generator=g) + 0.5
        elif kind in ("bn_b", "bn_m"):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_v":
            t = torch.rand(shape, generator=g) + 0.5
        else:
            t = torch.tensor(0, dtype=torch.long)
        sd[key] = t
    return sd
