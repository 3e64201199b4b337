"""fp32 PyTorch restatement of the reference's deployed PanopticDeepLab-PointRend forward pass,
driven by the state_dict of the fused TorchScript export (key names exactly as
`torch.jit.load(path).state_dict()` gives them; export recipe empanada_napari/_train.py:59-73).
TEST INFRASTRUCTURE ONLY (see oracle/post.py header). This is the "plain PyTorch fp32 reference"
the bf16 tcgen05 path is tolerance-checked against; it is itself pinned against the reference's
own classes (QuantizablePanopticDeepLabPR, models/quantization/panoptic_deeplab.py:148) by
tests/test_oracle_reference.py (runs where /root/reference exists) and by
tests/golden/model_pdl_tiny.npz (travels).
"""
import torch
import torch.nn.functional as F


def _conv(sd, name, x, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride, padding, dilation, groups)


def _bn(sd, name, x, eps=1e-5):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, eps)


def _bottleneck(sd, p, x, stride, dilation):
    """models/quantization/encoders/resnet.py:45-77 after fuse_model (BN folded)."""
    out = F.relu(_conv(sd, p + ".conv1.0", x))
    out = F.relu(_conv(sd, p + ".conv2.0", out, stride=stride, padding=dilation, dilation=dilation))
    out = _conv(sd, p + ".conv3", out)
    if (p + ".downsample.0.weight") in sd:
        x = _conv(sd, p + ".downsample.0", x, stride=stride)
    return F.relu(out + x)


def resnet50_encoder(sd, x, output_stride=16):
    """models/encoders/resnet.py:143-229."""
    x = F.relu(_conv(sd, "encoder.conv1.0", x, stride=2, padding=3))
    p1 = F.max_pool2d(x, 3, 2, 1)
    feats = [p1]
    x = p1
    last_stride = 1 if output_stride == 16 else 2
    last_dil = 2 if output_stride == 16 else 1
    for li, (nblocks, stride, dil) in enumerate([(3, 1, 1), (4, 2, 1), (6, 2, 1),
                                                 (3, last_stride, last_dil)], start=1):
        for b in range(nblocks):
            x = _bottleneck(sd, f"encoder.layer{li}.{b}", x, stride if b == 0 else 1, dil)
        feats.append(x)
    return feats


def _sepconv_bn_relu(sd, p, x, k=5):
    """models/blocks.py separable_conv_bn_act: depthwise k x k -> 1x1 -> live BN -> ReLU."""
    x = F.conv2d(x, sd[p + ".0.sepconv.0.weight"], None, 1, (k - 1) // 2, 1, x.shape[1])
    x = F.conv2d(x, sd[p + ".0.sepconv.1.weight"], None)
    return F.relu(_bn(sd, p + ".1", x))


def pdl_decoder(sd, p, feats, rates=(2, 4, 6)):
    """models/decoders/aspp.py:51-102 + decoders/panoptic_deeplab.py:68-80 (low_level_stages=[1])."""
    x = feats[-1]
    size = x.shape[-2:]
    branches = [F.relu(_conv(sd, p + ".aspp.convs.0.0.0", x))]
    for i, r in enumerate(rates, start=1):
        branches.append(F.relu(_conv(sd, f"{p}.aspp.convs.{i}.0.0", x, padding=r, dilation=r)))
    pooled = F.adaptive_avg_pool2d(x, 1)
    pooled = F.relu(F.conv2d(pooled, sd[p + ".aspp.convs.4.aspp_pooling.1.0.weight"]))
    branches.append(F.interpolate(pooled, size=size, mode="bilinear", align_corners=True))
    x = F.relu(_conv(sd, p + ".aspp.project.0.0", torch.cat(branches, dim=1)))
    low = F.relu(_conv(sd, p + ".project.0.0.0", feats[1]))
    x = F.interpolate(x, size=low.shape[-2:], mode="bilinear", align_corners=True)
    return _sepconv_bn_relu(sd, p + ".fuse.0", torch.cat((x, low), dim=1))


def pdl_head(sd, p, x):
    """models/heads.py:9-19."""
    x = _sepconv_bn_relu(sd, p + ".head.0", x)
    return _conv(sd, p + ".head.1", x)


def point_rend(sd, coarse, features, render_steps, num_points=8192, collect=None):
    """models/point_rend.py:241-269 (eval branch). `collect` (a list) receives, per render step,
    (predictor input (R,C,k), point indices (R,k), H, W): used by oracle/probe.py to fit the
    predictor layer."""
    sem = coarse.clone()
    nfc = 0
    # fused exports (PDL) nest conv+relu one level deeper than unfused ones (BiFPN)
    fc_fmt = "semantic_pr.point_head.fc_layers.{}.0.0"
    if fc_fmt.format(0) + ".weight" not in sd:
        fc_fmt = "semantic_pr.point_head.fc_layers.{}.0"
    while fc_fmt.format(nfc) + ".weight" in sd:
        nfc += 1
    for _ in range(render_steps):
        sem = F.interpolate(sem, scale_factor=2.0, mode="bilinear", align_corners=False)
        if sem.shape[1] == 1:
            unc = -torch.abs(sem)
        else:
            top2 = torch.topk(sem, k=2, dim=1)[0]
            unc = (top2[:, 1] - top2[:, 0]).unsqueeze(1)
        R, _, H, W = unc.shape
        k = min(H * W, num_points)
        idx = torch.topk(unc.view(R, H * W), k=k, dim=1)[1]
        coords = torch.zeros(R, k, 2, dtype=torch.float)
        coords[:, :, 0] = 0.5 / float(W) + (1.0 / float(W)) * (idx % W).float()
        coords[:, :, 1] = 0.5 / float(H) + (1.0 / float(H)) * torch.div(idx, W, rounding_mode="floor").float()
        grid = (2.0 * coords - 1.0).unsqueeze(2)
        cpts = F.grid_sample(coarse, grid, mode="bilinear", align_corners=False).squeeze(3)
        fpts = F.grid_sample(features, grid, mode="bilinear", align_corners=False).squeeze(3)
        x = torch.cat([fpts, cpts], dim=1)
        for l in range(nfc):
            pre = fc_fmt.format(l)
            x = F.relu(F.conv1d(x, sd[pre + ".weight"], sd[pre + ".bias"]))
            x = torch.cat([x, cpts], dim=1)
        if collect is not None:
            collect.append((x, idx, sem.shape[2], sem.shape[3]))
        logits = F.conv1d(x, sd["semantic_pr.point_head.predictor.weight"],
                          sd["semantic_pr.point_head.predictor.bias"])
        N, C, H, W = sem.shape
        sem = sem.reshape(N, C, H * W).scatter_(2, idx.unsqueeze(1).expand(-1, C, -1), logits).view(N, C, H, W)
    return sem


@torch.no_grad()
def pdl_forward(sd, x, render_steps=2, interpolate_ins=False):
    """QuantizablePanopticDeepLabPR.forward (models/quantization/panoptic_deeplab.py:194-250),
    eval mode, float path. x: (N,1,H,W) fp32. Returns dict of sem_logits, ctr_hmp, offsets
    (+ the /4 coarse logits and semantic_x for layer-wise checks)."""
    feats = resnet50_encoder(sd, x, output_stride=16)
    semantic_x = pdl_decoder(sd, "semantic_decoder", feats)
    if "instance_decoder.aspp.project.0.0.weight" in sd:
        instance_x = pdl_decoder(sd, "instance_decoder", feats)
    else:
        instance_x = semantic_x
    coarse = pdl_head(sd, "semantic_head", semantic_x)
    ctr = pdl_head(sd, "ins_center", instance_x)
    off = pdl_head(sd, "ins_xy", instance_x)
    sem = point_rend(sd, coarse, semantic_x, render_steps)
    if interpolate_ins:
        ctr = F.interpolate(ctr, scale_factor=4.0, mode="bilinear", align_corners=True)
        off = F.interpolate(off, scale_factor=4.0, mode="bilinear", align_corners=True)
    return {"sem_logits": sem, "ctr_hmp": ctr, "offsets": off, "coarse_logits": coarse,
            "semantic_x": semantic_x, "instance_x": instance_x, "p5": feats[-1], "p2": feats[1]}


# ------------------------------------------------------------------------- PanopticBiFPN-PR
def _conv_bn(sd, p, x):
    """blocks.py conv_bn_act(kernel_size=1, activation=None): p.0 conv (no bias), p.1 live BN."""
    return _bn(sd, p + ".1", F.conv2d(x, sd[p + ".0.weight"], None))


def _sepconv_bn_act(sd, p, x, k, act):
    """blocks.py separable_conv_bn_act: depthwise k x k -> 1x1 (both bias-free) -> BN -> act."""
    x = F.conv2d(x, sd[p + ".0.sepconv.0.weight"], None, 1, (k - 1) // 2, 1, x.shape[1])
    x = F.conv2d(x, sd[p + ".0.sepconv.1.weight"], None)
    return act(_bn(sd, p + ".1", x))


def _resample(sd, p, x):
    """blocks.py Resample2d: identity when no conv was created (nin == nout)."""
    return _conv_bn(sd, p + ".conv", x) if (p + ".conv.0.weight") in sd else x


def _fusion_weights(w, eps=1e-4):
    w = F.relu(w)
    return w / (w.sum() + eps)


def bifpn_layer(sd, p, pyr, eps=1e-4):
    """decoders/bifpn.py:14-158 (TopDownFPN, BottomUpFPN, BiFPNLayer). pyr: large -> small."""
    up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")
    down = lambda t: F.max_pool2d(t, 3, 2, 1)
    # top-down over [small ... large]
    tp = p + ".top_down_fpn"
    rev = pyr[::-1]
    w = _fusion_weights(sd[tp + ".weights"], eps)
    td = [rev[0]]
    for i in range(len(rev) - 1):
        high = _resample(sd, f"{tp}.resamplings.{i}", rev[i + 1])
        w1, w2 = w[i], w[i + 1]
        fused = (w1 * up(td[-1]) + w2 * high) / (w1 + w2 + eps)
        td.append(_sepconv_bn_act(sd, f"{tp}.after_combines.{i}", fused, 3, F.silu))
    # bottom-up over pyr[1:] with the top-down features large -> small
    bp = p + ".bottom_up_fpn"
    tdr = td[::-1]
    w = _fusion_weights(sd[bp + ".weights"], eps)
    bu = [tdr[0]]
    n = len(pyr) - 1
    for i in range(n):
        low = _resample(sd, f"{bp}.resamplings.{i}", pyr[1 + i])
        if i < n - 1:
            w1, w2, w3 = w[i], w[i + 1], w[i + 2]
            fused = (w1 * down(bu[-1]) + w2 * low + w3 * tdr[i + 1]) / (w1 + w2 + w3 + eps)
        else:
            w1, w2 = w[i], w[i + 1]
            fused = (w1 * down(bu[-1]) + w2 * low) / (w1 + w2 + eps)
        bu.append(_sepconv_bn_act(sd, f"{bp}.after_combines.{i}", fused, 3, F.silu))
    return bu


def bifpn(sd, p, feats):
    """decoders/bifpn.py:160-196. feats: [P3, P4, P5] encoder maps."""
    p6 = F.max_pool2d(_resample(sd, p + ".p6_resample", feats[-1]), 3, 2, 1)
    p7 = F.max_pool2d(p6, 3, 2, 1)
    pyr = list(feats) + [p6, p7]
    i = 0
    while f"{p}.bifpns.{i}.top_down_fpn.weights" in sd:
        pyr = bifpn_layer(sd, f"{p}.bifpns.{i}", pyr)
        i += 1
    return pyr


def bifpn_decoder(sd, p, fpn_features):
    """decoders/bifpn.py:198-236. fpn_features: small -> large, last one is the P2 skip."""
    x = fpn_features[0]
    for i, skip in enumerate(fpn_features[1:]):
        x = F.conv_transpose2d(x, sd[f"{p}.upsamplings.{i}.0.weight"], None, stride=2)
        x = F.relu(_bn(sd, f"{p}.upsamplings.{i}.1", x))
        x = torch.cat([x, skip], dim=1)
    return _sepconv_bn_act(sd, p + ".fusion", x, 5, F.relu)


@torch.no_grad()
def bifpn_forward(sd, x, render_steps=2, interpolate_ins=False):
    """QuantizablePanopticBiFPNPR.forward (models/quantization/panoptic_bifpn.py:147-161), eval
    mode, float path; encoder = fused ResNet-50 at output stride 32."""
    feats = resnet50_encoder(sd, x, output_stride=32)
    p2f = _resample(sd, "p2_resample", feats[1])
    sem_pyr = [p2f] + bifpn(sd, "semantic_fpn", feats[2:])
    semantic_x = bifpn_decoder(sd, "semantic_decoder", sem_pyr[::-1])
    if "instance_fpn.p6_resample.conv.0.weight" in sd:
        ins_pyr = [p2f] + bifpn(sd, "instance_fpn", feats[2:])
        instance_x = bifpn_decoder(sd, "instance_decoder", ins_pyr[::-1])
    else:
        instance_x = semantic_x
    coarse = pdl_head(sd, "semantic_head", semantic_x)
    ctr = pdl_head(sd, "ins_center", instance_x)
    off = pdl_head(sd, "ins_xy", instance_x)
    sem = point_rend(sd, coarse, semantic_x, render_steps)
    if interpolate_ins:
        ctr = F.interpolate(ctr, scale_factor=4.0, mode="bilinear", align_corners=True)
        off = F.interpolate(off, scale_factor=4.0, mode="bilinear", align_corners=True)
    return {"sem_logits": sem, "ctr_hmp": ctr, "offsets": off, "coarse_logits": coarse,
            "semantic_x": semantic_x, "instance_x": instance_x, "p5": feats[-1], "p2": feats[1],
            "sem_pyr": sem_pyr, "p2f": p2f}


def forward_any(sd, x, render_steps=2, interpolate_ins=False):
    if "semantic_fpn.p6_resample.conv.0.weight" in sd:
        return bifpn_forward(sd, x, render_steps, interpolate_ins)
    return pdl_forward(sd, x, render_steps, interpolate_ins)
