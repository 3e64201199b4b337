"""Model boundary: `load_model` mirrors `empanada_napari.utils.load_model_to_device`
(/root/reference/empanada_napari/utils.py:80-106): it takes the `model` entry of a model config
(a local TorchScript archive exported by the reference, a plain state_dict file, or an
already-constructed model object) and returns an object with

    forward_slices(volume_u8 (D,H,W) cuda, axis, s0, s1, norms, padding_factor)
        -> sem_logits (B,Hp,Wp) fp32, ctr_hmp (B,Hp/4,Wp/4) fp32, offsets (B,2,Hp/4,Wp/4) fp32

i.e. `model(image, render_steps, interpolate_ins=False)` of engines.py:250 for a batch of slices,
with slice extraction, normalisation and padding (volume_dataset.py:37-53, utils.py:170-201,
postprocess.py:26-36) fused in front.
"""
import os

import torch

from . import _lib


class SyntheticHeadsModel:
    """Test/benchmark stand-in for the network: head maps come from a callable
    `heads_fn(axis, s0, s1) -> (sem_logits, ctr_hmp, offsets)` (cuda tensors). Optionally runs a
    real model first so that the forward pass is still executed and timed (bench.py)."""

    def __init__(self, heads_fn, inner=None):
        self.heads_fn = heads_fn
        self.inner = inner
        self.launches = 0

    def forward_slices(self, vol_d, axis, s0, s1, norms, pf, **kw):
        if self.inner is not None:
            self.inner.forward_slices(vol_d, axis, s0, s1, norms, pf, **kw)
            self.launches = self.inner.launches
        # `plane_axis`: the plane the slices were taken from when they arrive as a re-sampled stack
        return self.heads_fn(kw.get("plane_axis", axis), s0, s1)


class HostHeadsModel:
    """Picklable variant of `SyntheticHeadsModel` for multi-process tests: `heads[axis]` =
    (sem_logits (N,H,W), ctr_hmp (N,h4,w4), offsets (N,2,h4,w4)) numpy arrays, uploaded to the
    current device on first use."""

    def __init__(self, heads):
        self.heads = heads
        self.launches = 0
        self._dev = None

    def __getstate__(self):
        return {"heads": self.heads, "launches": 0, "_dev": None}

    def forward_slices(self, vol_d, axis, s0, s1, norms, pf, **kw):
        if self._dev is None:
            self._dev = {a: tuple(torch.from_numpy(t).to(vol_d.device) for t in h) for a, h in self.heads.items()}
        sem, ctr, off = self._dev[kw.get("plane_axis", axis)]
        return sem[s0:s1], ctr[s0:s1], off[s0:s1]


def load_state_dict_any(path):
    """state_dict of a TorchScript archive (the reference's deployment format) or a plain
    `torch.save(state_dict)` file."""
    try:
        m = torch.jit.load(path, map_location="cpu")
        return {k: v.detach() for k, v in m.state_dict().items()}
    except Exception:
        sd = torch.load(path, map_location="cpu", weights_only=True)
        if not isinstance(sd, dict):
            raise _lib.B200EmpanadaError(f"{path}: neither a TorchScript archive nor a state_dict")
        return sd


def resolve_model_file(fpath_or_url, download=True):
    """Local file for a model URL: the reference's cache location `torch.hub.get_dir()/<basename
    of the URL path>` (empanada_napari/utils.py:84-104), downloaded there when absent exactly as
    the reference does; `~/.empanada/<basename>` is accepted as a second cache location."""
    from urllib.parse import urlparse
    filename = os.path.basename(urlparse(fpath_or_url).path)
    hub_dir = torch.hub.get_dir()
    cached = os.path.join(hub_dir, filename)
    if os.path.exists(cached):
        return cached
    alt = os.path.join(os.path.expanduser("~"), ".empanada", filename)
    if os.path.isfile(alt):
        return alt
    if not download or not (fpath_or_url.startswith("http://") or fpath_or_url.startswith("https://")):
        raise _lib.B200EmpanadaError(f"model file {fpath_or_url} not found (looked in {cached} and {alt})")
    os.makedirs(hub_dir, exist_ok=True)
    try:
        import sys
        sys.stderr.write(f'Downloading: "{fpath_or_url}" to {cached}\n')
        torch.hub.download_url_to_file(fpath_or_url, cached, None, progress=True)
    except Exception as e:
        raise _lib.B200EmpanadaError(f"model URL {fpath_or_url} is not cached at {cached} and the download failed: {e}")
    return cached


def load_model(model, device, model_config=None):
    if hasattr(model, "forward_slices"):
        return model
    if isinstance(model, dict):
        sd = model
    elif isinstance(model, str):
        path = model
        if not os.path.isfile(path):
            path = resolve_model_file(path)
        sd = load_state_dict_any(path)
    else:
        raise _lib.B200EmpanadaError(f"unsupported model specification: {type(model)}")
    from .bifpn import BiFPNModel, is_bifpn_state_dict
    from .pdl import PDLModel, is_pdl_state_dict
    if is_pdl_state_dict(sd):
        return PDLModel(sd, device)       # MitoNet_v1, NucleoNet_base_v2, DropNet_base_v1
    if is_bifpn_state_dict(sd):
        return BiFPNModel(sd, device)     # MitoNet_v1_mini
    raise _lib.B200EmpanadaError(
        "unsupported network: neither a PanopticDeepLab-PointRend nor a PanopticBiFPN-PointRend "
        "ResNet-50 export")
