"""Linear-probe weights for the end-to-end agreement tests. TEST INFRASTRUCTURE ONLY (see
oracle/post.py header).

Real MitoNet weights are not available offline, and a randomly initialised network emits logits
that hover around the decision thresholds, so a bf16-vs-fp32 comparison of the hardened output
would measure coin flips, not the implementation. `fit_probe_heads` keeps every randomly
initialised layer of the synthetic PanopticDeepLab-PointRend state_dict and fits ONLY the four
final linear layers (semantic head, centre head, offset head, PointRend predictor) by ridge
regression on fp32 oracle features of a seeded training volume, against the analytic targets of
`empanada_napari_b200.synthetic.analytic_heads`. The result behaves like a (weakly) trained
model: confident semantic logits, one heat-map peak per object, offsets that point to centroids.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import model as om
from . import post


def _ridge(feats, target, lam=1e-2):
    """feats (n, c) float64, target (n, k) -> weight (k, c), bias (k,)."""
    mu, tm = feats.mean(0), target.mean(0)
    x = feats - mu
    a = x.T @ x / len(x) + lam * np.eye(x.shape[1])
    w = np.linalg.solve(a, x.T @ (target - tm) / len(x)).T
    return w, tm - w @ mu


def _block_mean(a, s):
    h, w = a.shape
    return a.reshape(h // s, s, w // s, s).mean((1, 3))


@torch.no_grad()
def fit_probe_heads(sd, volume, labels, norms, padding_factor=16, sem_logit=6.0, lam=1e-2, axes=(0, 1, 2)):
    """Returns a copy of `sd` whose head.1 layers and PointRend predictor are fitted on the
    slices of (volume, labels) along `axes`."""
    import empanada_napari_b200.synthetic as syn
    sd = {k: v.clone() for k, v in sd.items()}
    slices = [(np.take(volume, i, axis=a), np.take(labels, i, axis=a)) for a in axes for i in range(volume.shape[a])]
    hs, hc, ho, ts, tc, to = [], [], [], [], [], []
    cache = []
    for img, lab in slices:
        x = torch.from_numpy(post.factor_pad(post.normalize(img, norms["mean"], norms["std"]), padding_factor)[None, None])
        H, W = x.shape[-2:]
        labp = np.zeros((H, W), dtype=np.int32)
        labp[:lab.shape[0], :lab.shape[1]] = lab
        feats = om.resnet50_encoder(sd, x, 16)
        sx = om.pdl_decoder(sd, "semantic_decoder", feats)
        ix = om.pdl_decoder(sd, "instance_decoder", feats) if "instance_decoder.aspp.project.0.0.weight" in sd else sx
        _, ctr, off = syn.analytic_heads(lab, pad_to=padding_factor)
        fg4 = _block_mean((labp > 0).astype(np.float64), 4)
        for p, src, store in (("semantic_head", sx, hs), ("ins_center", ix, hc), ("ins_xy", ix, ho)):
            h = om._sepconv_bn_relu(sd, p + ".head.0", src)[0]
            store.append(h.reshape(h.shape[0], -1).T.double().numpy())
        ts.append((sem_logit * (2 * fg4 - 1)).reshape(-1, 1))
        tc.append(ctr.reshape(-1, 1).astype(np.float64))
        to.append(off.reshape(2, -1).T.astype(np.float64))
        cache.append((sx, labp))
    for p, h, t in (("semantic_head", hs, ts), ("ins_center", hc, tc), ("ins_xy", ho, to)):
        w, b = _ridge(np.concatenate(h), np.concatenate(t), lam)
        sd[p + ".head.1.weight"] = torch.from_numpy(w).float()[:, :, None, None].contiguous()
        sd[p + ".head.1.bias"] = torch.from_numpy(b).float()
    # PointRend predictor: inputs collected at the points the fitted coarse head marks uncertain
    px, pt = [], []
    for sx, labp in cache:
        coarse = om.pdl_head(sd, "semantic_head", sx)
        got = []
        om.point_rend(sd, coarse, sx, 2, collect=got)
        for xin, idx, h, w in got:
            s = labp.shape[0] // h
            tgt = sem_logit * (2 * _block_mean((labp > 0).astype(np.float64), s) - 1) if s > 1 else sem_logit * (2.0 * (labp > 0) - 1)
            px.append(xin[0].T.double().numpy())
            pt.append(tgt.reshape(-1)[idx[0].numpy()][:, None])
    w, b = _ridge(np.concatenate(px), np.concatenate(pt), lam)
    sd["semantic_pr.point_head.predictor.weight"] = torch.from_numpy(w).float()[:, :, None].contiguous()
    sd["semantic_pr.point_head.predictor.bias"] = torch.from_numpy(b).float()
    return sd
