"""The `empanada.inference.engines` object API on the sm_100a kernels: per-slice engines with the
reference's constructor keywords, methods, attributes and return types
(/root/reference/empanada/inference/engines.py:223-394).

    PanopticDeepLabRenderEngine(model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                                nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5,
                                padding_factor=16, coarse_boundaries=True)
        engine(image, size, upsampling=1) -> pan_seg (1, h, w) int64 tensor on the model's device
    PanopticDeepLabRenderEngine3d(..., median_kernel_size=3, ...)
        engine(image, size, upsampling=1) -> pan_seg or None while the median queue builds
        engine.end(upsampling=1) -> list of the remaining pan_segs;  engine.reset()

`image` is what the reference's data pipeline hands over: a (1, 1, h, w) fp32 tensor that is already
normalised (Preprocessor, empanada_napari/utils.py:187-201). `model` is anything `model.load_model`
accepts (TorchScript archive path, state_dict, or a loaded network object). These engines make one
launch sequence per slice; `inference.Engine3d` / `Engine2d` are the batched, full-throughput
drivers over the same kernels and read their parameters from an engine object of this module, so
mutating `engine.ks`, `engine.nms_kernel`, ... behaves as it does in the reference
(empanada_napari/inference.py:439-455).
"""
import math
from collections import deque

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .model import load_model

__all__ = ["PanopticDeepLabRenderEngine", "PanopticDeepLabRenderEngine3d", "logits_to_prob"]


def logits_to_prob(logits):
    """engines.py:22-30 (sigmoid for one class, softmax over classes otherwise)."""
    if logits.size(1) > 1:
        return torch.softmax(logits, dim=1)
    return torch.sigmoid(logits)


def _unsupported(name):
    raise NotImplementedError(
        f"{name} is outside the hot path built so far (SURVEY.md section 8f 'next' rows)")


class PanopticDeepLabRenderEngine:
    """engines.py:223-325."""

    def __init__(self, model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, padding_factor=16,
                 coarse_boundaries=True, **kwargs):
        if not torch.cuda.is_available():
            raise _lib.B200EmpanadaError("a CUDA device (B200) is required; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = load_model(model, self.device)
        self.thing_list = thing_list
        self.label_divisor = label_divisor
        self.stuff_area = stuff_area
        self.void_label = void_label
        self.nms_threshold = nms_threshold
        self.nms_kernel = nms_kernel
        self.confidence_thr = confidence_thr
        self.padding_factor = padding_factor
        self.coarse_boundaries = coarse_boundaries
        self.center_cap = 4096

    # ------------------------------------------------------------------ pieces of the reference API
    def to_model_device(self, tensor):
        return tensor.to(self.device, non_blocking=True)

    def _check(self, upsampling):
        assert math.log(upsampling, 2).is_integer(), "Upsampling factor not log base 2!"
        if len(self.thing_list) > 1:
            _unsupported("multi-class models")

    def infer(self, image, render_steps=2):
        """model(image, render_steps, interpolate_ins=not coarse_boundaries) + logits_to_prob
        (engines.py:248-256). `image`: padded (1, 1, H, W) fp32 on the device."""
        x = image.reshape(1, image.shape[-2], image.shape[-1]).to(torch.float32).contiguous()
        sem, ctr, off = self.model.forward_slices(x, 0, 0, 1, None, self.padding_factor, render_steps=int(render_steps))
        sem, ctr, off = sem[None].clone(), ctr[None].clone(), off.clone()
        if not self.coarse_boundaries:
            from .inference import upsample_instance_heads
            c, o = upsample_instance_heads(ctr[0], off)
            ctr, off = c[None], o
        return {"sem_logits": sem, "ctr_hmp": ctr, "offsets": off, "sem": logits_to_prob(sem)}

    def _cells_lowres(self, ctr_hmp, offsets):
        """find_instance_center + group_pixels on the head grid: int32 (h', w') ids, 0 = none."""
        hh, ww = ctr_hmp.shape[-2:]
        ctr = ctr_hmp.reshape(1, hh, ww).to(torch.float32).contiguous()
        off = offsets.reshape(1, 2, hh, ww).to(torch.float32).contiguous()
        dev = ctr.device
        while True:
            centers = torch.zeros((1, self.center_cap), dtype=torch.int32, device=dev)
            counts = torch.zeros(1, dtype=torch.int32, device=dev)
            scratch = torch.empty((hh * ww + 1023) // 1024, dtype=torch.int32, device=dev)
            call("be_centers", ptr(ctr), 1, hh, ww, float(self.nms_threshold), int(self.nms_kernel),
                 ptr(centers), self.center_cap, ptr(counts), ptr(scratch), stream_ptr())
            k = int(counts.item())
            if k <= self.center_cap:
                break
            self.center_cap = 1 << (k - 1).bit_length()
        cells = torch.zeros((1, hh, ww), dtype=torch.int32, device=dev)
        step = 4.0 if self.coarse_boundaries else 1.0
        call("be_group_pixels", ptr(off), ptr(centers), self.center_cap, ptr(counts), 1, hh, ww, step,
             ptr(cells), stream_ptr())
        return cells[0], k

    def get_instance_cells(self, ctr_hmp, offsets, upsampling=1):
        """engines.py:258-275: (1, 1, H, W) fp32 tensor of instance ids (nearest-upsampled x step)."""
        cells, _ = self._cells_lowres(ctr_hmp, offsets)
        step = (4 if self.coarse_boundaries else 1) * int(upsampling)
        out = cells.to(torch.float32)
        if step > 1:
            out = out.repeat_interleave(step, 0).repeat_interleave(step, 1)
        return out[None, None]

    def _pan_from(self, hard_u8, cells_i32, scale):
        """merge_semantic_and_instance for one thing class (postprocess.py:224-296) on the padded
        slice: hard (H, W) uint8, cells (H/scale, W/scale) int32 -> pan (H, W) int32."""
        H, W = hard_u8.shape
        dev = hard_u8.device
        cap = self.center_cap
        pan = torch.empty((1, H, W), dtype=torch.int32, device=dev)
        present = torch.empty((1, cap + 1), dtype=torch.int32, device=dev)
        call("be_merge_pan", ptr(hard_u8), ptr(cells_i32), 1, H, W, H, W, int(scale), cap,
             int(self.label_divisor), int(self.thing_list[0]), int(self.void_label), ptr(present),
             ptr(pan), stream_ptr())
        return pan

    def _harden_seg(self, sem):
        """engines.py:115-121 (single class)."""
        if sem.size(1) > 1:
            _unsupported("multi-class semantic heads")
        return (sem >= self.confidence_thr).long()

    def get_panoptic_seg(self, sem, instance_cells):
        """engines.py:277-293: sem (1, H, W) hardened, instance_cells (1, 1, H, W) -> (1, H, W) int64."""
        if len(self.thing_list) > 1:
            _unsupported("multi-class models")
        if len(self.thing_list) == 0:
            # no thing classes: every class of the hardened map is pasted as stuff where it
            # covers at least `stuff_area` pixels (postprocess.py:283-294)
            pan = torch.full_like(sem, self.void_label, dtype=torch.int64)
            for class_id in torch.unique(sem).tolist():
                mask = (sem == class_id).to(torch.uint8).contiguous()
                area = torch.zeros(1, dtype=torch.int32, device=sem.device)
                call("be_slice_area", ptr(mask), 1, mask.shape[-2], mask.shape[-1], ptr(area), stream_ptr())
                if int(area.item()) >= self.stuff_area:
                    pan[mask.bool()] = int(class_id) * self.label_divisor
            return pan
        hard = (sem[0] == self.thing_list[0]).to(torch.uint8).contiguous()
        cells = instance_cells[0, 0].to(torch.int32).contiguous()
        return self._pan_from(hard, cells, 1).long()

    def postprocess(self, sem, instance_cells):
        """engines.py:295-299."""
        return self.get_panoptic_seg(self._harden_seg(sem)[0], instance_cells)

    def _pad(self, image):
        from torch.nn.functional import pad
        h, w = image.shape[-2:]
        pf = self.padding_factor
        ph, pw = (pf - h % pf) % pf, (pf - w % pf) % pf
        return pad(image, (0, pw, 0, ph)) if (ph or pw) else image

    def __call__(self, image, size, upsampling=1):
        self._check(upsampling)
        assert image.ndim == 4 and image.size(0) == 1
        h, w = (int(v) for v in size)
        image = self.to_model_device(self._pad(image))
        out = self.infer(image, int(2 + math.log(upsampling, 2)))
        cells = self.get_instance_cells(out["ctr_hmp"], out["offsets"], upsampling)
        pan = self.postprocess(out["sem"], cells)
        return pan[..., :h, :w]


class PanopticDeepLabRenderEngine3d(PanopticDeepLabRenderEngine):
    """engines.py:327-394 with the `_MedianQueue` (engines.py:47-90): the filtered semantic map
    REPLACES the queued middle item, so later medians see already filtered slices."""

    def __init__(self, model, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, median_kernel_size=3,
                 padding_factor=16, coarse_boundaries=True, **kwargs):
        super().__init__(model, thing_list, label_divisor, stuff_area, void_label, nms_threshold,
                         nms_kernel, confidence_thr, padding_factor, coarse_boundaries)
        assert median_kernel_size % 2 == 1, "Kernel size must be odd integer!"
        self.ks = median_kernel_size
        self.mid_idx = (median_kernel_size - 1) // 2
        self.median_queue = deque(maxlen=median_kernel_size)

    def reset(self):
        self.median_queue = deque(maxlen=self.ks)

    def enqueue(self, item):
        self.median_queue.append(item)

    def get_median(self, key):
        """Per-pixel median over the queued items (engines.py:60-66) on the median kernel: pushing
        the ks queued maps through an empty queue emits their median at slot `mid_idx`."""
        stack = torch.cat([o[key] for o in self.median_queue], dim=0).to(torch.float32).contiguous()
        ks = stack.shape[0]
        stack = stack.reshape(ks, stack.shape[-2], stack.shape[-1])
        _, H, W = stack.shape
        dev = stack.device
        hist = torch.empty((max(ks - 1, 1), H, W), dtype=torch.float32, device=dev)
        hard = torch.empty((ks, H, W), dtype=torch.uint8, device=dev)
        out = torch.empty((ks, H, W), dtype=torch.float32, device=dev)
        call("be_median_push", ptr(stack), ks, H, W, ks, ptr(hist), 0, 0, float(self.confidence_thr), 1,
             ptr(hard), ptr(out), stream_ptr())
        return out[self.mid_idx][None, None]

    def get_next(self, keys):
        nq = len(self.median_queue)
        if nq <= self.mid_idx:
            return self.median_queue[-1]
        if nq < self.ks:
            return None
        output = self.median_queue[self.mid_idx]
        for key in keys:
            output[key] = self.get_median(key)
        return output

    def _finish(self, model_out, upsampling):
        h, w = model_out["size"]
        cells = self.get_instance_cells(model_out["ctr_hmp"], model_out["offsets"], upsampling)
        return self.postprocess(model_out["sem"], cells)[..., :h, :w]

    def end(self, upsampling=1):
        return [self._finish(o, upsampling) for o in list(self.median_queue)[self.mid_idx + 1:]]

    def __call__(self, image, size, upsampling=1):
        self._check(upsampling)
        assert image.ndim == 4 and image.size(0) == 1
        image = self.to_model_device(self._pad(image))
        out = self.infer(image, int(2 + math.log(upsampling, 2)))
        out["size"] = tuple(int(v) for v in size)
        self.enqueue(out)
        median_out = self.get_next(keys=["sem"])
        if median_out is None:
            return None
        return self._finish(median_out, upsampling)
