"""Multi-GPU stack inference: one process per GPU (torchrun / torch.distributed, NCCL over
NVLink). Replaces the reference's `MultiGPUEngine3d` (empanada_napari/multigpu.py:121-260), which
round-robins slices over ranks and all_gathers full-resolution `sem` and `instance_cells` on
every step (patterns.py:226-240, multigpu.py:90-91).

Here each rank runs the network on a CONTIGUOUS slice range of every plane and the head maps are
gathered once per plane to that plane's leader rank (`dist.gather`), which runs the sequential
part (recursive median, tracker replay) and the remaining post-processing; the three planes have
different leaders, so their post-processing runs concurrently, and rank 0 receives the finished
label volumes for the consensus. Round-1 scope: the conv stack (90 % of the single-GPU time) is
what is sharded by slice; sharding the post-processing by slice range with halo exchange is the
next step.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .inference import Engine3d, upsample_instance_heads

_PROFILE = os.environ.get("B200_EMPANADA_PROFILE") == "1"


class _Timer:
    """Per-rank phase times (B200_EMPANADA_PROFILE=1): device-synchronised wall clock."""

    def __init__(self):
        self.t = {}
        if _PROFILE:
            torch.cuda.synchronize()
        self.t0 = time.perf_counter()

    def mark(self, name):
        if _PROFILE:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            self.t[name] = self.t.get(name, 0.0) + t1 - self.t0
            self.t0 = t1


def slice_ranges(n, world):
    """Contiguous, balanced [lo, hi) per rank."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_slices_to(local, ranges, rank, world, dst=0, group=None):
    """Gathers per-rank slice blocks (first dim padded to the longest range) to rank `dst` and
    returns the list of per-rank blocks trimmed to their true length (None on other ranks). Works
    on any backend (NCCL on GPUs, gloo in the CPU tests)."""
    if rank == dst:
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, bufs, dst=dst, group=group)
        return [bufs[r][: ranges[r][1] - ranges[r][0]] for r in range(world)]
    dist.gather(local, None, dst=dst, group=group)
    return None


def gather_slices_to_root(local, ranges, rank, world, group=None):
    return gather_slices_to(local, ranges, rank, world, 0, group)


class DistributedEngine3d(Engine3d):
    """Engine3d for one process per GPU. Every rank must call `infer_on_axis` for the same planes
    in the same order and then `finalize(trackers)`.

    * the network forward of each plane is sharded by contiguous slice range over ALL ranks;
    * the head maps of plane p are gathered (NCCL) to that plane's LEADER rank, which owns the
      sequential part (recursive median, components, tracker replay, RLE) - the three planes'
      leaders work concurrently, and the post-processing is deferred until `finalize` so that no
      rank's forward pass waits behind another plane's post-processing;
    * `finalize` ships each leader's result (dense label volume + instance table) to rank 0, where
      `tracker_consensus` runs.
    """

    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._pending = {}

    def leader_of(self, axis_name):
        # keep rank 0 (consensus) as free as the world size allows
        return (self.axes[axis_name] + 1) % self.world

    def infer_on_axis(self, volume, axis_name):
        self._check_supported()
        axis, vol_d, shape3d, n, h, w, H, W, pf = self._plane_setup(volume, axis_name)
        leader = self.leader_of(axis_name)
        ranges = slice_ranges(n, self.world)
        lo, hi = ranges[self.rank]
        dev = vol_d.device
        nmax = max(b - a for a, b in ranges)
        sem = torch.empty((nmax, H, W), dtype=torch.float32, device=dev)
        ctr = torch.empty((nmax, H // 4, W // 4), dtype=torch.float32, device=dev)
        off = torch.empty((nmax, 2, H // 4, W // 4), dtype=torch.float32, device=dev)
        launches0 = getattr(self.model, "launches", 0)
        tm = _Timer()
        bs = self.slice_batch(H, W)
        for s0 in range(lo, hi, bs):
            s1 = min(hi, s0 + bs)
            a, b, c = self.model.forward_slices(vol_d, axis, s0, s1, self.model_config["norms"], pf)
            sem[s0 - lo:s1 - lo].copy_(a)
            ctr[s0 - lo:s1 - lo].copy_(b)
            off[s0 - lo:s1 - lo].copy_(c)
        tm.mark("forward")
        gathered = [gather_slices_to(t, ranges, self.rank, self.world, leader, self.group) for t in (sem, ctr, off)]
        tm.mark("gather heads")
        n_launch = getattr(self.model, "launches", 0) - launches0
        trackers = self.create_trackers(shape3d, axis_name)
        if self.rank == leader:
            post = self._make_post(n, h, w, H, W)
            for r, (a, b) in enumerate(ranges):
                for s0 in range(a, b, 64):
                    s1 = min(b, s0 + 64)
                    c_, o_ = gathered[1][r][s0 - a:s1 - a], gathered[2][r][s0 - a:s1 - a]
                    if self.fine_boundaries:
                        c_, o_ = upsample_instance_heads(c_.contiguous(), o_.contiguous())
                    post.push_heads(gathered[0][r][s0 - a:s1 - a], c_, o_, is_prob=False)
            self._pending[axis_name] = (post, shape3d)
        tm.mark("leader push_heads")
        if _PROFILE:
            print(f"[rank {self.rank}] {axis_name} " + " ".join(f"{k}={v:.3f}" for k, v in tm.t.items()), flush=True)
        self.last_stats = {"kernel_launches": n_launch}
        return None, trackers

    def finalize(self, trackers):
        """Collective. Completes the deferred post-processing on the plane leaders and assembles
        all trackers on rank 0 (dense volume + instance table; the per-instance RLE arrays stay on
        the leader). Returns the trackers dict (complete on rank 0)."""
        n_launch = 0
        tm = _Timer()
        # a leader that owns several planes (world < 4) runs their matcher replays concurrently on
        # worker threads while the next plane's component kernels are queued
        for axis_name, (post, shape3d) in list(self._pending.items()):
            trackers[axis_name] = self._finish_plane(post, axis_name, shape3d, defer=len(self._pending) > 1)
        for axis_name, (post, shape3d) in list(self._pending.items()):
            _ = trackers[axis_name][0].instances      # completes a deferred tracker
            n_launch += post.launches
        self._pending = {}
        tm.mark("leader post-processing")
        self.last_stats = {"kernel_launches": n_launch}
        for axis_name in trackers.keys():
            leader = self.leader_of(axis_name)
            if leader == 0:
                continue
            tr = trackers[axis_name][0]
            shape3d = tuple(int(s) for s in tr.shape3d)
            if self.rank == leader:
                labels = np.array(list(tr.instances.keys()), dtype=np.int64)
                meta = np.zeros((len(labels), 8), dtype=np.int64)
                for i, l in enumerate(labels):
                    meta[i, 0] = l
                    meta[i, 1] = tr._b200_sizes[int(l)]
                    meta[i, 2:8] = tr.instances[int(l)]["box"]
                cnt = torch.tensor([len(labels)], dtype=torch.int64, device=self.device)
                dist.send(cnt, dst=0, group=self.group)
                if len(labels):
                    dist.send(torch.from_numpy(meta).to(self.device), dst=0, group=self.group)
                dist.send(tr._b200_dense, dst=0, group=self.group)
            elif self.rank == 0:
                cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
                dist.recv(cnt, src=leader, group=self.group)
                k = int(cnt.item())
                meta = torch.zeros((k, 8), dtype=torch.int64, device=self.device)
                if k:
                    dist.recv(meta, src=leader, group=self.group)
                dense = torch.empty(shape3d, dtype=torch.int32, device=self.device)
                dist.recv(dense, src=leader, group=self.group)
                meta = meta.cpu().numpy()
                empty = np.zeros(0, dtype=np.int64)
                tr.instances = {int(m[0]): {"box": tuple(int(v) for v in m[2:8]), "starts": empty, "runs": empty}
                                for m in meta}
                tr._b200_sizes = {int(m[0]): int(m[1]) for m in meta}
                tr._b200_dense = dense
                tr.finish()
        tm.mark("ship volumes to rank 0")
        if _PROFILE:
            print(f"[rank {self.rank}] finalize " + " ".join(f"{k}={v:.3f}" for k, v in tm.t.items()), flush=True)
        return trackers
