"""Multi-GPU stack inference: one process per GPU (torch.distributed, NCCL over NVLink).
Replaces the reference's `MultiGPUEngine3d` (empanada_napari/multigpu.py:121-260), which
round-robins slices over ranks and all_gathers full-resolution `sem` and `instance_cells` on
every step (patterns.py:226-240, multigpu.py:90-91).

* `ShardedEngine3d` (the engine under torchrun, and what `MultiGPUEngine3d` runs on every rank):
  the WHOLE per-plane path - network, median queue, centres, grouping, components, overlap
  tables, painting - sharded by contiguous slice range; NVLink carries the median-queue wavefront,
  one component slice per shard boundary, the sparse tables to the plane's leader and the label
  table back.
* `MultiGPUEngine3d`: the widget-facing front end, constructed in one process with the reference's
  keywords; it starts the other ranks itself.
* `DistributedEngine3d`: the earlier scheme (only the network sharded, head maps gathered to the
  plane leaders), kept selectable with B200_EMPANADA_MULTIGPU=gather.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .inference import Engine3d, _VolumeCache, _as_device_volume, upsample_instance_heads

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
        self._cache.ready()
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
                    if not self.engine.coarse_boundaries:
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


class _SharedHostVolume:
    """A (D,H,W) int32 result volume in POSIX shared memory that every rank of one host maps: each
    rank copies its own painted z-slab device->host over its own PCIe link (its range of the
    mapping is page-locked once with cudaHostRegister), instead of funnelling the whole volume
    through rank 0's GPU and one link. Rank 0 owns the segment and hands out a numpy view; the
    segment is reused for the next result of the same shape once the caller has released the
    previous array (as the single-GPU page-locked pool does)."""

    def __init__(self, group, rank, world):
        self.group, self.rank, self.world = group, rank, world
        self.shm, self.name, self.shape = None, None, None
        self.array = None        # rank 0: the numpy view handed to the caller
        self.registered = None   # (address, bytes) of the page-locked range of this rank
        self.serial = 0

    def _unregister(self):
        if self.registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self.registered[0])
            self.registered = None

    def close(self):
        self._unregister()
        self.array = None
        if self.shm is not None:
            try:
                self.shm.close()
                if self.rank == 0:
                    self.shm.unlink()
            except Exception:
                pass
            self.shm = None

    def acquire(self, shape, z_range):
        """Collective. Returns (torch int32 view of the whole volume, numpy view on rank 0)."""
        import sys
        from multiprocessing import shared_memory
        shape = tuple(int(v) for v in shape)
        nbytes = int(np.prod(shape)) * 4
        name = [None]
        if self.rank == 0:
            busy = self.array is not None and sys.getrefcount(self.array) > 2
            if self.shm is None or self.shape != shape or busy:
                if busy:      # the caller still holds the previous result: leave that segment to it
                    self._unregister()
                    self.shm, self.array = None, None
                else:
                    self.close()
                self.serial += 1
                try:     # a too-small /dev/shm would only fail later, with SIGBUS on first touch
                    st = os.statvfs("/dev/shm")
                    room = st.f_bavail * st.f_frsize
                except OSError:
                    room = 0
                if room > nbytes + (64 << 20):
                    self.shm = shared_memory.SharedMemory(create=True, size=nbytes,
                                                          name=f"b200emp_{os.getpid()}_{id(self):x}_{self.serial}")
                    self.shape = shape
            name[0] = self.shm.name if self.shm is not None else None
        dist.broadcast_object_list(name, src=0, group=self.group)
        if name[0] is None:      # no room in shared memory: the caller gathers through rank 0
            return None, None
        if self.rank != 0 and (self.shm is None or self.shm.name != name[0]):
            self.close()
            self.shm = _attach_shm(name[0])
            self.shape = shape
        full = np.ndarray(shape, dtype=np.int32, buffer=self.shm.buf)
        z0, z1 = z_range
        if z1 > z0:
            part = full[z0:z1]
            addr, size = part.ctypes.data, part.nbytes
            if self.registered != (addr, size):
                self._unregister()
                err = torch.cuda.cudart().cudaHostRegister(addr, size, 0)
                if int(err) != 0:
                    raise _lib_error(f"cudaHostRegister failed ({err})")
                self.registered = (addr, size)
        if self.rank == 0:
            self.array = full
        return torch.from_numpy(full), (full if self.rank == 0 else None)


class _ReplicatedVolumeCache(_VolumeCache):
    """Upload of a host volume that EVERY rank holds (one process per GPU under torchrun): each
    rank copies 1/G of it over its own PCIe link and NCCL all-gathers the pieces over NVLink,
    instead of G full pageable host->device copies competing for host memory bandwidth."""

    def __init__(self, group, rank, world):
        super().__init__()
        self.group, self.rank, self.world = group, rank, world

    def _upload(self, volume, device):
        G, r = self.world, self.rank
        D = volume.shape[0]
        per = -(-D // G)
        if not np.issubdtype(volume.dtype, np.integer):
            return _as_device_volume(volume, device)      # raises the reference's error
        tdtype = torch.from_numpy(np.zeros(1, dtype=volume.dtype)).dtype
        full = torch.empty((per * G,) + tuple(volume.shape[1:]), dtype=tdtype, device=device)
        piece = torch.zeros((per,) + tuple(volume.shape[1:]), dtype=tdtype, device=device)
        lo, hi = min(D, r * per), min(D, (r + 1) * per)
        if hi > lo:
            piece[:hi - lo].copy_(torch.from_numpy(np.ascontiguousarray(volume[lo:hi])))
        dist.all_gather_into_tensor(full, piece, group=self.group)
        return full[:D]


def owned_ranges(n, world, mid):
    """Slice-sharded planes: rank r runs the network on F_r = slice_ranges(n, world)[r] and OWNS
    (emits, labels, paints) E_r = F_r shifted down by the median latency `mid` - pushing slice t
    into the recursive median queue emits slice t - mid (engines.py:68-82). Rank 0 starts at 0 and
    the last rank also owns the unfiltered tail (engines.py:351-361). Returns (F, E) lists."""
    F = slice_ranges(n, world)
    E = []
    for r, (lo, hi) in enumerate(F):
        e_lo = lo - mid if r > 0 else 0
        e_hi = hi - mid if r < world - 1 else n
        E.append((e_lo, e_hi))
    return F, E


def merge_shard_tables(parts, owned):
    """Leader side: per-rank (n_cc [n_r], table [n_r, cap_r, 5], pair_keys u64, pair_vals) with
    LOCAL slice indices -> plane-wide tables with absolute slice indices (keys carry the slice in
    bits 40+)."""
    cap = max(1, max(int(p[1].shape[1]) for p in parts))
    n_cc = np.concatenate([p[0] for p in parts]).astype(np.int32)
    tables = []
    for p in parts:
        t = p[1]
        if t.shape[1] < cap:
            t = np.concatenate([t, np.zeros((t.shape[0], cap - t.shape[1], 5), dtype=t.dtype)], axis=1)
        tables.append(t)
    table = np.concatenate(tables, axis=0).astype(np.int32)
    keys = np.concatenate([p[2].astype(np.uint64) + (np.uint64(e_lo) << np.uint64(40)) for p, (e_lo, _) in zip(parts, owned)])
    vals = np.concatenate([p[3] for p in parts]).astype(np.int32)
    return n_cc, table, keys, vals


class ShardedEngine3d(Engine3d):
    """Engine3d for one process per GPU with the WHOLE per-plane path sharded by slice range
    (SURVEY.md section 8e): every rank runs the network, centres, grouping, merge, connected
    components and overlap tables for its own slices; what crosses NVLink is

    * the median queue state ((ks-1) filtered/raw fp32 slices) handed from rank r to r+1 - the
      recursion makes this a wavefront, but only the cheap median kernel is serialised;
    * one component-label slice per shard boundary for the cross-boundary overlap table;
    * the sparse per-slice tables (component areas / boxes, overlap pairs) gathered to the plane's
      leader, which replays the sequential tracker on a worker thread WHILE the next plane's
      forward pass runs, and broadcasts the (slice, component) -> label table; each rank paints
      its own slab;
    * for the consensus (`sharded_consensus`): one all-to-all that turns the y-slabs of the xz
      plane and the x-slabs of the yz plane into z-slabs, then only sparse tables (pair counts,
      cluster membership, sizes, run-length ranges) between the ranks and rank 0;
    * optionally (`finalize(gather_dense=True)`) the painted slabs, gathered to rank 0.

    Rank r also runs the network on the `mid` slices before its range so that it holds the
    instance heads of every slice it owns (no halo exchange for them). Every rank must call
    `infer_on_axis` for the same planes in the same order and then `finalize(trackers)`.
    Raises when a shard would be shorter than the median kernel (use fewer ranks, or
    `DistributedEngine3d`, for very short stacks)."""

    def __init__(self, *args, group=None, replicated_input=True, gather_dense=True, **kwargs):
        super().__init__(*args, **kwargs)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # gather_dense=False: the painted plane volumes stay sharded (sharded_consensus needs
        # nothing else), and each plane is completed - label table broadcast, slabs painted - at
        # the end of the NEXT plane's infer_on_axis, by when its tracker replay has long finished
        self.gather_dense = gather_dense
        self._finalize_launches = 0
        self._pending = {}
        self._slabs = {}
        self.front = None        # the MultiGPUEngine3d that drives this engine, if any
        self._host_out = _SharedHostVolume(group, self.rank, self.world)
        if replicated_input:     # host volumes are passed to every rank (torchrun)
            self._cache = _ReplicatedVolumeCache(group, self.rank, self.world)

    def leader_of(self, axis_name):
        return (self.axes[axis_name] + 1) % self.world

    def close(self):
        """Releases the shared host result volume (rank 0 unlinks the segment)."""
        self._host_out.close()

    def gather_plane(self, name):
        """Collective: the painted slabs of plane `name` assembled on rank 0 (None elsewhere)."""
        G, r = self.world, self.rank
        slab, E, shape3d = self._slabs[name]
        ax = self.axes[name]
        if r != 0:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, slab.contiguous(), 0, group=self.group)]):
                req.wait()
            return None
        dense = torch.empty(shape3d, dtype=torch.int32, device=self.device)
        dense.narrow(ax, E[0][0], E[0][1] - E[0][0]).copy_(slab)
        tmps, ops = [], []
        for src in range(1, G):
            shp = list(shape3d)
            shp[ax] = E[src][1] - E[src][0]
            tmps.append(torch.empty(shp, dtype=torch.int32, device=self.device))
            ops.append(dist.P2POp(dist.irecv, tmps[-1], src, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for src in range(1, G):
            dense.narrow(ax, E[src][0], E[src][1] - E[src][0]).copy_(tmps[src - 1])
        return dense

    def infer_on_axis(self, volume, axis_name):
        from .inference import _Async
        self._check_supported(sharded=True)
        if len(self.engine.thing_list) == 0:
            raise _lib_error("semantic-only inference is not built in the slice-sharded multi-GPU engine")
        axis, vol_d, shape3d, n, h, w, H, W, pf = self._plane_setup(volume, axis_name)
        self._cache.ready()
        ks = self.median_kernel_size
        mid = (ks - 1) // 2
        G, r = self.world, self.rank
        F, E = owned_ranges(n, G, mid)
        if min(b - a for a, b in F) < ks + mid or min(b - a for a, b in E) < ks:
            raise _lib_error(f"plane of {n} slices is too short to shard over {G} ranks with a median kernel of {ks}")
        (lo, hi), (e_lo, e_hi) = F[r], E[r]
        c_lo = lo - mid if r > 0 else 0
        dev = vol_d.device
        tm = _Timer()
        launches0 = getattr(self.model, "launches", 0)
        post = self._make_post(e_hi - e_lo, h, w, H, W)
        raw = torch.empty((hi - lo, H, W), dtype=torch.float32, device=dev)
        bs = self.slice_batch(H, W)
        norms = self.model_config["norms"]
        for s0 in range(c_lo, hi, bs):
            s1 = min(hi, s0 + bs)
            sem, ctr, off = self.model.forward_slices(vol_d, axis, s0, s1, norms, pf)
            a = max(s0, lo)
            if a < s1:
                raw[a - lo:s1 - lo].copy_(sem[a - s0:])
            a, b = max(s0, e_lo), min(s1, e_hi)
            if a < b:
                c_, o_ = ctr[a - s0:b - s0], off[a - s0:b - s0]
                if not self.engine.coarse_boundaries:
                    c_, o_ = upsample_instance_heads(c_.contiguous(), o_.contiguous())
                post.push_instance(c_, o_, a - e_lo)
        tm.mark("forward + centres + grouping")
        # recursive median: wavefront over the ranks (state = the queue after the previous shard)
        if ks > 1 and r > 0:
            dist.recv(post.hist, src=r - 1, group=self.group)
            post.n_hist = ks - 1
        for i in range(0, hi - lo, 64):
            j = min(hi - lo, i + 64)
            post.push_semantic(raw[i:j], (lo + i) - e_lo)
        if ks > 1 and r < G - 1:
            dist.send(post.hist, dst=r + 1, group=self.group)
        if r == G - 1:
            post.flush_semantic()
        del raw
        post.pushed = post.N
        post.check_centers()
        tm.mark("median wavefront")
        post.run_cc()
        # cross-boundary overlaps: previous shard's last component slice vs this shard's first
        ops, prev = [], None
        if r < G - 1:
            last = post.cc_images(post.N - 1, post.N)[0].contiguous()
            ops.append(dist.P2POp(dist.isend, last, r + 1, group=self.group))
        if r > 0:
            prev = torch.empty((h, w), dtype=torch.int32, device=dev)
            ops.append(dist.P2POp(dist.irecv, prev, r - 1, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if prev is not None:
            bk, bv = post.boundary_pairs(prev)
            post.pair_keys = np.concatenate([post.pair_keys, bk])
            post.pair_vals = np.concatenate([post.pair_vals, bv])
        tm.mark("components + tables + boundary overlap")
        # sparse tables -> the plane's leader, whose tracker replay runs on a worker thread while
        # every rank goes on with the next plane
        leader = self.leader_of(axis_name)
        n_cc, table = post.replay_inputs()
        parts = [None] * G if r == leader else None
        dist.gather_object((n_cc, table, post.pair_keys, post.pair_vals), parts, dst=leader, group=self.group)
        job = None
        if r == leader:
            merged = merge_shard_tables(parts, E)
            job = _Async(tracking_replay, merged, post.cls, post.div, axis_name, self.merge_iou_thr,
                         self.merge_ioa_thr, self.min_size, self.min_extent)
        tm.mark("gather tables")
        plane_trackers = self.create_trackers(shape3d, axis_name)
        if not self.gather_dense:
            for prev in list(self._pending.keys()):     # earlier planes: replay done behind this forward pass
                self._finalize_plane(prev, False)
            tm.mark("previous plane: broadcast + paint")
        if _PROFILE:
            print(f"[rank {self.rank}] {axis_name} " + " ".join(f"{k}={v:.3f}" for k, v in tm.t.items()), flush=True)
        self._pending[axis_name] = (post, shape3d, (F, E), job, plane_trackers)
        self.last_stats = {"kernel_launches": getattr(self.model, "launches", 0) - launches0 + post.launches}
        return None, plane_trackers

    def _finalize_plane(self, name, gather_dense):
        """Collective: the leader's label table of plane `name` is broadcast, every rank paints
        its slab, rank 0's tracker of the plane receives the instance table."""
        from .postproc import LazyPlane
        G, r = self.world, self.rank
        post, shape3d, (F, E), job, plane_trackers = self._pending.pop(name)
        leader = self.leader_of(name)
        payload = [job.result() if r == leader else None]
        dist.broadcast_object_list(payload, src=leader, group=self.group)
        lut_f, kept_labels, kept_boxes, kept_sizes = payload[0]
        e_lo, e_hi = E[r]
        D, Hv, Wv = shape3d
        local_shape = {"xy": (e_hi - e_lo, Hv, Wv), "xz": (D, e_hi - e_lo, Wv), "yz": (D, Hv, e_hi - e_lo)}[name]
        n0 = post.launches
        slab = post.relabel(np.ascontiguousarray(lut_f[e_lo:e_hi]), name, local_shape)
        self._finalize_launches += post.launches - n0
        self._slabs[name] = (slab, E, shape3d)
        dense = self.gather_plane(name) if gather_dense else None
        if r == 0:
            tr = plane_trackers[0]
            if dense is not None:
                plane = LazyPlane(dense, name, kept_labels, kept_boxes)
                tr._b200_dense = dense
            else:
                plane = ShardedPlane(self, tr, name, kept_labels, kept_boxes)
            tr.instances = plane.attrs
            tr._b200_sizes = dict(zip(np.asarray(kept_labels).tolist(), np.asarray(kept_sizes).tolist()))
            tr._b200_sharded = self
            tr.finish()

    def finalize(self, trackers, gather_dense=None):
        """Collective: the plane leaders' label tables are broadcast and every rank paints its
        own slab of every plane (kept in `self._slabs` for `sharded_consensus`). With
        `gather_dense` (default: the constructor's value) the slabs are also assembled into dense
        label volumes on rank 0, as the single-GPU engine leaves them (`tracker_consensus`,
        per-plane RLE); without it rank 0's trackers carry only the instance tables (boxes, sizes).
        Planes that were already completed at the end of a later `infer_on_axis` call (the engine
        was constructed with `gather_dense=False`) are not touched again. Returns the trackers
        dict (complete on rank 0; the tracker objects are the ones `infer_on_axis` returned)."""
        tm = _Timer()
        gather_dense = self.gather_dense if gather_dense is None else gather_dense
        for name in list(self._pending.keys()):
            self._finalize_plane(name, gather_dense)
        tm.mark("replay + broadcast + paint" + (" + slabs to rank 0" if gather_dense else ""))
        if _PROFILE:
            print(f"[rank {self.rank}] finalize " + " ".join(f"{k}={v:.3f}" for k, v in tm.t.items()), flush=True)
        self.last_stats = {"kernel_launches": self._finalize_launches}
        self._finalize_launches = 0
        return trackers

    # ------------------------------------------------------------------ sharded consensus
    def _z_slabs(self):
        """All-to-all re-shard of the painted plane slabs into this rank's z-slab of all three
        planes: z ranges = the xy plane's slice ownership (no exchange for xy); xz arrives as
        (dz, Y_src, W) blocks, yz as (dz, H, X_src) blocks."""
        G, r = self.world, self.rank
        xy_slab, Ez, shape3d = self._slabs["xy"]
        D, Hv, Wv = shape3d
        z0, z1 = Ez[r]
        out = [xy_slab]
        for name, ax in (("xz", 1), ("yz", 2)):
            slab, E, _ = self._slabs[name]
            send = [slab[Ez[q][0]:Ez[q][1]] for q in range(G)]              # leading-dim slices: contiguous
            recv = []
            for src in range(G):
                shp = [z1 - z0, Hv, Wv]
                shp[ax] = E[src][1] - E[src][0]
                recv.append(torch.empty(shp, dtype=torch.int32, device=self.device))
            dist.all_to_all(recv, send, group=self.group)
            full = torch.empty((z1 - z0, Hv, Wv), dtype=torch.int32, device=self.device)
            for src in range(G):
                full.narrow(ax, E[src][0], E[src][1] - E[src][0]).copy_(recv[src])
            out.append(full)
        return out, z0

    def sharded_consensus(self, trackers, model_config, pixel_vote_thr=2, cluster_iou_thr=0.75,
                          allow_one_view=False, min_size=200, min_extent=4, gather_volume=True, to_host=False):
        """Collective form of `tracker_consensus` for ONE thing class over the slabs `finalize`
        left on every rank: each rank votes on its own z-slab; rank 0 runs the graph decisions on
        the summed sparse tables (consensus.consensus_driver). Returns on rank 0 (volume,
        class_name, instances); (None, None, None) elsewhere. The volume is the device tensor
        assembled on rank 0 (`gather_volume`), or with `to_host` a numpy int32 array in shared
        host memory that every rank filled with its own slab in parallel."""
        from . import consensus
        G, r = self.world, self.rank
        tm = _Timer()
        vols, z0 = self._z_slabs()
        tm.mark("consensus: slab all-to-all")
        shard = consensus.ConsensusShard(vols, [None, None, None], z0=z0)
        if r == 0:
            class_id = model_config["thing_list"][0]
            class_name = model_config["class_names"][class_id]
            trs = [trackers[n][0] for n in ("xy", "xz", "yz")]
            n_nodes, node_sizes, node_boxes, luts = consensus.tracker_node_tables(trs)
            min_cluster = 1 if allow_one_view else 2
            if pixel_vote_thr < min_cluster:
                cluster_iou_thr = 0

            def each(method, *args):
                dist.broadcast_object_list([(method, args)], src=0, group=self.group)
                res = [None] * G
                dist.gather_object(_guarded(shard, method, args), res, dst=0, group=self.group)
                for x in res:
                    if isinstance(x, BaseException):
                        dist.broadcast_object_list([("done", ())], src=0, group=self.group)
                        raise x
                return [_import_arrays(x) for x in res]
            if n_nodes == 0:
                each("zero")
                instances = {}
            else:
                instances = consensus.consensus_driver(each, G, n_nodes, node_sizes, node_boxes, luts, pixel_vote_thr,
                                                       cluster_iou_thr, min_cluster, min_size, min_extent, tm.mark)
            dist.broadcast_object_list([("done", ())], src=0, group=self.group)
        else:
            exports = []      # shared-memory segments holding this rank's large results
            while True:
                cmd = [None]
                dist.broadcast_object_list(cmd, src=0, group=self.group)
                method, args = cmd[0]
                if method == "done":
                    break
                dist.gather_object(_guarded(shard, method, args, exports), None, dst=0, group=self.group)
            _release_exports(exports)
        self.consensus_launches = shard.launches
        vol = None
        full_t = None
        if to_host:
            _, Ez, shape3d = self._slabs["xy"]
            full_t, vol = self._host_out.acquire(shape3d, Ez[r])
        if full_t is not None:
            full_t[Ez[r][0]:Ez[r][1]].copy_(shard.painted, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)
            tm.mark("consensus: painted slabs to shared host memory")
        elif to_host:            # fallback: assemble on rank 0, one page-locked copy from there
            from .inference import _PINNED
            vol = self._gather_z(shard.painted)
            if r == 0:
                vol = _PINNED.to_host(vol, np.int32)
            tm.mark("consensus: painted slabs to rank 0 + copy to the host")
        elif gather_volume:
            vol = self._gather_z(shard.painted)
            tm.mark("consensus: painted slabs to rank 0")
        if _PROFILE:
            print(f"[rank {self.rank}] consensus " + " ".join(f"{k}={v:.3f}" for k, v in tm.t.items()), flush=True)
        self._last_painted = shard.painted
        if r != 0:
            return None, None, None
        return vol, class_name, instances

    def _gather_z(self, painted):
        """Painted z-slabs -> the whole volume on rank 0 (None elsewhere)."""
        G, r = self.world, self.rank
        _, Ez, shape3d = self._slabs["xy"]
        if r == 0:
            vol = torch.empty(shape3d, dtype=torch.int32, device=self.device)
            vol[Ez[0][0]:Ez[0][1]].copy_(painted)
            ops = [dist.P2POp(dist.irecv, vol[Ez[src][0]:Ez[src][1]], src, group=self.group) for src in range(1, G)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            return vol
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, painted, 0, group=self.group)]):
            req.wait()
        return None


def _guarded(shard, method, args, exports=None):
    """One shard call of the consensus driver; an exception travels to rank 0 as the result.
    `exports` (a list, non-root ranks): large numpy results go through a shared-memory segment of
    this rank instead of the pickled object gather; the segment is appended for later clean-up."""
    try:
        res = getattr(shard, method)(*args)
    except Exception as e:  # re-raised on rank 0
        return e
    if exports is not None and isinstance(res, tuple) and res and all(isinstance(a, np.ndarray) for a in res) \
            and sum(a.nbytes for a in res) > (1 << 20):
        try:
            return _export_arrays(res, exports)
        except Exception:
            return res
    return res


def _attach_shm(name):
    """Attach to a peer's segment without making this process's resource tracker its owner
    (before Python 3.13 attaching registers the name, and the tracker would unlink it - with a
    warning - when this process exits)."""
    from multiprocessing import resource_tracker, shared_memory
    shm = shared_memory.SharedMemory(name=name)
    try:
        resource_tracker.unregister(shm._name, "shared_memory")
    except Exception:
        pass
    return shm


def _export_arrays(arrays, exports):
    from multiprocessing import shared_memory
    total = sum(a.nbytes for a in arrays)
    st = os.statvfs("/dev/shm")
    if st.f_bavail * st.f_frsize < total + (64 << 20):
        raise OSError("no room in /dev/shm")
    shm = shared_memory.SharedMemory(create=True, size=total, name=f"b200emp_t{os.getpid()}_{len(exports)}_{time.monotonic_ns()}")
    exports.append(shm)
    meta, off = [], 0
    for a in arrays:
        a = np.ascontiguousarray(a)
        np.ndarray(a.shape, a.dtype, buffer=shm.buf, offset=off)[...] = a
        meta.append((a.shape, a.dtype.str, off))
        off += a.nbytes
    return ("__b200_shm__", shm.name, meta)


def _import_arrays(res):
    """Root side of `_export_arrays`: copies the arrays out of the peer's segment."""
    if not (isinstance(res, tuple) and len(res) == 3 and isinstance(res[0], str) and res[0] == "__b200_shm__"):
        return res
    shm = _attach_shm(res[1])
    try:
        return tuple(np.array(np.ndarray(shape, np.dtype(dt), buffer=shm.buf, offset=off)) for shape, dt, off in res[2])
    finally:
        shm.close()


def _release_exports(exports):
    for shm in exports:
        try:
            shm.close()
            shm.unlink()
        except Exception:
            pass
    exports.clear()


class ShardedPlane:
    """Instance table of one plane on rank 0 when the dense label volume stays sharded over the
    ranks (`finalize(gather_dense=False)`): boxes and sizes are known, the per-instance RLE is
    fetched from the ranks on first use when a `MultiGPUEngine3d` front end drives them."""

    def __init__(self, engine, tracker, axis_name, labels, boxes):
        from .postproc import LazyAttrs
        self.engine, self.tracker, self.axis_name = engine, tracker, axis_name
        self.labels, self.boxes = labels, boxes
        self.attrs = {l: LazyAttrs(tuple(b), self) for l, b in zip(np.asarray(labels).tolist(), np.asarray(boxes).tolist())}

    def dense(self):
        front = self.engine.front
        if front is None:
            raise _lib_error(f"the {self.axis_name} label volume is sharded over the ranks: call "
                             "finalize(trackers, gather_dense=True) to read per-plane volumes or run-length tables")
        vol = front.gather_plane(self.axis_name)
        self.tracker._b200_dense = vol
        return vol

    def materialize(self):
        from .postproc import instances_from_dense
        full = instances_from_dense(self.dense(), self.axis_name, self.labels, self.boxes)
        for l, a in full.items():
            attrs = self.attrs.get(l)
            if attrs is not None:
                attrs._owner = None
                dict.__setitem__(attrs, "starts", a["starts"])
                dict.__setitem__(attrs, "runs", a["runs"])


def tracking_replay(merged, cls, div, axis_name, iou_thr, ioa_thr, min_size, min_extent):
    """Leader side (worker thread): tracker replay on the merged tables + the two tracker filters
    (inference.py:556-558). Returns (lut [N, stride] with dropped instances zeroed, kept labels,
    kept boxes, kept sizes)."""
    from . import tracking
    n_cc, table, keys, vals = merged
    lut, labels, sizes, boxes = tracking.match_replay(n_cc, table, keys, vals, cls, div, axis_name, iou_thr, ioa_thr)
    spans = boxes[:, 3:] - boxes[:, :3] if len(boxes) else np.zeros((0, 3), np.int32)
    keep = (sizes >= min_size) & (spans >= min_extent).all(axis=1) if len(labels) else np.zeros(0, bool)
    kept_labels = labels[keep]
    max_label = int(lut.max()) if lut.size else 0
    keep_lut = np.zeros(max_label + 1, dtype=np.int32)
    keep_lut[kept_labels] = kept_labels
    return keep_lut[lut], kept_labels, boxes[keep], sizes[keep]


def _lib_error(msg):
    from ._lib import B200EmpanadaError
    return B200EmpanadaError(msg)


# ------------------------------------------------------------------------------------------
# One-process front end: what the widget constructs (empanada_napari/_volume_inference.py:17,
# empanada_napari/multigpu.py:121-260)
# ------------------------------------------------------------------------------------------
# a collective that a dead rank never joins must not hang the caller's thread for ever: the ranks
# only meet inside one `infer_on_axis` / `consensus` call (seconds), idle waits are on queues
_FRONT_TIMEOUT = __import__("datetime").timedelta(seconds=int(os.environ.get("B200_EMPANADA_NCCL_TIMEOUT", "300")))


def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_main(rank, world, port, model_config, engine_kwargs, cmd_q, err_q):
    """Rank r > 0 of a `MultiGPUEngine3d`: a persistent process bound to GPU r that mirrors every
    call of the parent (rank 0) on its own `ShardedEngine3d`."""
    import traceback
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=dev, timeout=_FRONT_TIMEOUT)
        eng = ShardedEngine3d(model_config, replicated_input=False, **engine_kwargs)
        vol_version, vol_d = None, None
        while True:
            cmd = cmd_q.get()
            if cmd[0] == "close":
                break
            if cmd[0] == "update_params":
                eng.update_params(*cmd[1])
            elif cmd[0] == "infer":
                _, axis_name, shape, dtype_name, version, gather_dense = cmd
                if version != vol_version:
                    vol_d = None
                    vol_d = torch.empty(shape, dtype=getattr(torch, dtype_name), device=dev)
                    dist.broadcast(vol_d, src=0)
                    vol_version = version
                _, trackers = eng.infer_on_axis(vol_d, axis_name)
                eng.finalize({axis_name: trackers}, gather_dense=gather_dense)
            elif cmd[0] == "gather_plane":
                eng.gather_plane(cmd[1])
            elif cmd[0] == "consensus":
                eng.sharded_consensus(None, None, **cmd[1])
            elif cmd[0] == "release":
                vol_version, vol_d = None, None
        eng._host_out.close()
        dist.destroy_process_group()
    except BaseException:
        err_q.put((rank, traceback.format_exc()))
        raise


class MultiGPUEngine3d:
    r"""Drop-in for `empanada_napari.multigpu.MultiGPUEngine3d` (multigpu.py:121-260): constructed
    in ONE process with the reference's keywords, raises below two GPUs, runs one process per GPU
    and returns `(stack, trackers)` from `infer_on_axis`.

    The calling process is rank 0 (GPU 0); ranks 1..G-1 are persistent worker processes started
    once in the constructor (the reference re-spawns and re-loads the model for every plane,
    multigpu.py:216-220). Each call runs `ShardedEngine3d` on every rank: the whole per-plane path
    sharded by slice range, the uint8 volume uploaded once by rank 0 and broadcast over NVLink."""

    def __init__(self, model_config, inference_scale=1, label_divisor=1000, median_kernel_size=5,
                 stuff_area=64, void_label=0, nms_threshold=0.1, nms_kernel=3, confidence_thr=0.3,
                 force_connected=True, min_size=500, min_extent=4, fine_boundaries=False,
                 semantic_only=False, store_url=None, chunk_size=(256, 256, 256), save_panoptic=False,
                 world_size=None, batch_size=None):
        import torch.multiprocessing as mp
        if not torch.cuda.device_count() > 1:
            raise Exception("MultiGPU inference requires multiple GPUs! Run torch.cuda.device_count()")
        if dist.is_initialized():
            raise _lib_error("MultiGPUEngine3d starts its own process group; under torchrun use ShardedEngine3d")
        self.world = int(world_size or torch.cuda.device_count())
        if not 2 <= self.world <= torch.cuda.device_count():
            raise _lib_error(f"world_size must be between 2 and {torch.cuda.device_count()}")
        self.labels = model_config["labels"]
        self.config = model_config
        self.axes = {"xy": 0, "xz": 1, "yz": 2}
        self.min_size, self.min_extent = min_size, min_extent
        self.save_panoptic, self.chunk_size = save_panoptic, chunk_size
        self.zarr_store = None
        self.dtype = np.int32
        kwargs = dict(inference_scale=inference_scale, label_divisor=label_divisor,
                      median_kernel_size=median_kernel_size, stuff_area=stuff_area, void_label=void_label,
                      nms_threshold=nms_threshold, nms_kernel=nms_kernel, confidence_thr=confidence_thr,
                      force_connected=force_connected, min_size=min_size, min_extent=min_extent,
                      fine_boundaries=fine_boundaries, semantic_only=semantic_only, store_url=store_url,
                      chunk_size=chunk_size, save_panoptic=False, batch_size=batch_size)
        port = _free_port()
        ctx = mp.get_context("spawn")
        self._err_q = ctx.Queue()
        self._cmd_qs, self._procs = [], []
        for r in range(1, self.world):
            q = ctx.Queue()
            p = ctx.Process(target=_worker_main, args=(r, self.world, port, model_config, kwargs, q, self._err_q),
                            daemon=True)
            p.start()
            self._cmd_qs.append(q)
            self._procs.append(p)
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=self.world,
                                device_id=torch.device("cuda", 0), timeout=_FRONT_TIMEOUT)
        self._engine = ShardedEngine3d(model_config, replicated_input=False, **kwargs)
        self._engine.front = self
        self.engine = self._engine.engine
        self._version = 0
        self._uploaded = None
        self._planes_version = {"xy": None, "xz": None, "yz": None}

    def _send(self, *cmd):
        if not self._err_q.empty():
            rank, tb = self._err_q.get()
            raise _lib_error(f"worker rank {rank} failed:\n{tb}")
        dead = [i + 1 for i, p in enumerate(self._procs) if not p.is_alive()]
        if dead:
            raise _lib_error(f"worker rank(s) {dead} exited; create a new MultiGPUEngine3d")
        for q in self._cmd_qs:
            q.put(cmd)

    def create_trackers(self, shape3d, axis_name):
        return self._engine.create_trackers(shape3d, axis_name)

    def update_params(self, *args):
        """Same positional arguments as `Engine3d.update_params` (inference.py:414-432)."""
        self._send("update_params", args)
        self._engine.update_params(*args)

    def infer_on_axis(self, volume, axis_name):
        eng = self._engine
        vol_d = eng._cache.get(volume, eng.device)
        if vol_d is not self._uploaded:          # a new device copy: every rank needs it
            self._uploaded = vol_d
            self._version += 1
            fresh = True
        else:
            fresh = False
        # the dense plane volume comes to rank 0 only when the caller wants the stack; otherwise it
        # stays sharded (fetched on demand if someone reads the per-plane run-length tables)
        self._send("infer", axis_name, tuple(vol_d.shape), str(vol_d.dtype).replace("torch.", ""), self._version,
                   bool(self.save_panoptic))
        if fresh:
            eng._cache.ready()           # the chunked upload has to be complete before it is broadcast
            dist.broadcast(vol_d, src=0)
        _, trackers = eng.infer_on_axis(vol_d, axis_name)
        trackers = eng.finalize({axis_name: trackers}, gather_dense=bool(self.save_panoptic))[axis_name]
        stack = trackers[0]._b200_dense.cpu().numpy() if self.save_panoptic else None
        self._planes_version[axis_name] = self._version
        return stack, trackers

    def gather_plane(self, axis_name):
        """Dense (D,H,W) label volume of a plane on GPU 0 (collective, driven from here)."""
        self._send("gather_plane", axis_name)
        return self._engine.gather_plane(axis_name)

    def consensus(self, trackers, model_config, to_host=False, **params):
        """`tracker_consensus` over the slabs the ranks still hold (called by
        `inference.tracker_consensus` when it is handed this engine's trackers)."""
        if len(set(self._planes_version.get(n) for n in ("xy", "xz", "yz"))) != 1 or None in self._planes_version.values():
            raise _lib_error("the xy, xz and yz trackers must come from the same volume")
        params = dict(params, to_host=bool(to_host), gather_volume=not to_host)
        self._send("consensus", params)
        return self._engine.sharded_consensus(trackers, model_config, **params)

    def release(self):
        self._send("release")
        self._engine.release()
        self._uploaded = None

    def close(self):
        """Stops the worker processes (also done when the object is collected)."""
        if getattr(self, "_procs", None):
            try:
                for q in self._cmd_qs:
                    q.put(("close",))
                for p in self._procs:
                    p.join(timeout=30)
                self._engine._host_out.close()
                dist.destroy_process_group()
            finally:
                self._procs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
