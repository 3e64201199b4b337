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


def point_rend(sd, coarse, features, render_steps, num_points=8192):
    """models/point_rend.py:241-269 (eval branch)."""
    sem = coarse.clone()
    nfc = 0
    while f"semantic_pr.point_head.fc_layers.{nfc}.0.0.weight" in sd:
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
            pre = f"semantic_pr.point_head.fc_layers.{l}.0.0"
            x = F.relu(F.conv1d(x, sd[pre + ".weight"], sd[pre + ".bias"]))
            x = torch.cat([x, cpts], dim=1)
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
