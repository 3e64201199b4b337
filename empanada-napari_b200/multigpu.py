"""Multi-GPU stack inference: one process per GPU (torchrun / torch.distributed, NCCL over
NVLink). Replaces the reference's `MultiGPUEngine3d` (empanada_napari/multigpu.py:121-260), which
round-robins slices over ranks and all_gathers full-resolution `sem` and `instance_cells` on
every step (patterns.py:226-240, multigpu.py:90-91).

Here each rank runs the network on a CONTIGUOUS slice range of the plane and the head maps are
gathered once per plane to rank 0 (`dist.gather` into the plane buffers), which runs the
sequential part (recursive median, tracker replay) and the remaining post-processing.
Round-1 scope: the conv stack (>90 % of the single-GPU time) is what is sharded; sharding the
post-processing by slice range with halo exchange is the next step (multi-GPU).
"""
import numpy as np
import torch
import torch.distributed as dist

from .inference import Engine3d


def slice_ranges(n, world):
    """Contiguous, balanced [lo, hi) per rank."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_slices_to_root(local, ranges, rank, world, group=None):
    """Gathers per-rank slice blocks (first dim padded to the longest range) to rank 0 and returns
    the list of per-rank blocks trimmed to their true length (None on other ranks). Works on any
    backend (NCCL on GPUs, gloo in the CPU tests)."""
    if rank == 0:
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, bufs, dst=0, group=group)
        return [bufs[r][: ranges[r][1] - ranges[r][0]] for r in range(world)]
    dist.gather(local, None, dst=0, group=group)
    return None


class DistributedEngine3d(Engine3d):
    """Engine3d whose forward pass is sharded by slice range across the ranks of the default
    process group. `infer_on_axis` must be called by every rank; trackers are complete on rank 0
    (other ranks return empty trackers)."""

    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def _forward_all(self, post, vol_d, axis, n, norms, pf):
        ranges = slice_ranges(n, self.world)
        lo, hi = ranges[self.rank]
        H, W = post.H, post.W
        dev = vol_d.device
        nmax = max(b - a for a, b in ranges)
        sem = torch.empty((nmax, H, W), dtype=torch.float32, device=dev)
        ctr = torch.empty((nmax, H // 4, W // 4), dtype=torch.float32, device=dev)
        off = torch.empty((nmax, 2, H // 4, W // 4), dtype=torch.float32, device=dev)
        for s0 in range(lo, hi, self.batch_size):
            s1 = min(hi, s0 + self.batch_size)
            a, b, c = self.model.forward_slices(vol_d, axis, s0, s1, norms, pf)
            sem[s0 - lo:s1 - lo].copy_(a)
            ctr[s0 - lo:s1 - lo].copy_(b)
            off[s0 - lo:s1 - lo].copy_(c)
        gathered = [gather_slices_to_root(t, ranges, self.rank, self.world, self.group) for t in (sem, ctr, off)]
        if self.rank == 0:
            for r, (a, b) in enumerate(ranges):
                for s0 in range(a, b, 64):
                    s1 = min(b, s0 + 64)
                    post.push_heads(gathered[0][r][s0 - a:s1 - a], gathered[1][r][s0 - a:s1 - a],
                                    gathered[2][r][s0 - a:s1 - a], is_prob=False)
        else:
            # keep the per-rank state machine consistent: nothing to post-process here
            post.pushed = post.N
            post.n_hist = min(post.N, post.ks - 1)

    def infer_on_axis(self, volume, axis_name):
        if self.rank == 0:
            return super().infer_on_axis(volume, axis_name)
        # non-zero ranks: run the sharded forward only
        self._check_supported()
        axis = self.axes[axis_name]
        vol_d = self._cache.get(volume, self.device)
        shape3d = tuple(int(s) for s in vol_d.shape)
        n = shape3d[axis]
        h, w = [s for i, s in enumerate(shape3d) if i != axis]
        pf = self.padding_factor
        H, W = h + (pf - h % pf) % pf, w + (pf - w % pf) % pf

        class _Shape:
            pass
        post = _Shape()
        post.H, post.W, post.N, post.ks, post.pushed, post.n_hist = H, W, n, self.median_kernel_size, 0, 0
        launches0 = getattr(self.model, "launches", 0)
        self._forward_all(post, vol_d, axis, n, self.model_config["norms"], pf)
        self.last_stats = {"kernel_launches": getattr(self.model, "launches", 0) - launches0}
        return None, self.create_trackers(shape3d, axis_name)
