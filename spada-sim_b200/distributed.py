"""Multi-GPU plumbing for the row-sharded path: one process per GPU,
torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests).

  * B (and A) replicated from rank 0 with ``dist.broadcast``            -> ``broadcast_csr``
  * A row-sharded into contiguous ranges of equal intermediate-product
    count: the C ABI's ``spada_b200_plan_shards`` on rank 0, broadcast   -> ``plan_bounds``
  * C gathered on every rank by the store itself: the kernel that places a shard's rows writes
    them into every rank's C buffers over NVLink peer mappings (CUDA IPC)  -> ``PeerGather``
  * the NCCL-only exchange kept as the baseline to compare against (and
    for the gloo tests): padded all-gather / grouped send+recv           -> ``allgather_csr``

torch is plumbing only (device memory views, streams, collectives); every SpGEMM kernel is in
libspada_b200.so.
"""
from __future__ import annotations

import collections
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr: int, n: int, typestr: str, owner=None):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr if n else 0, False),
                                         "version": 2, "strides": None}
        self._owner = owner


def device_view(ptr: int, n: int, dtype: torch.dtype, device, owner=None) -> torch.Tensor:
    """Zero-copy torch view of ``n`` elements at device address ``ptr`` (kept alive by ``owner``)."""
    if n == 0:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.float64: "<f8"}[dtype]
    t = torch.as_tensor(_DevArray(ptr, n, typestr, owner), device=device)
    t._spada_owner = owner
    return t


def result_views(result, device):
    """(row_ptr int64 [rows+1], col int32 [nnz], val float64 [nnz]) views of an engine Result."""
    p, c, v = result.device_ptrs()
    rows = result.shape[0]
    return (device_view(p, rows + 1, torch.int64, device, result), device_view(c, result.nnz, torch.int32, device, result),
            device_view(v, result.nnz, torch.float64, device, result))


class PinnedCsr:
    """A host CSR in pinned memory (int64 row_ptr, int32 col, float64 val): uploads run at full PCIe speed."""

    def __init__(self, mat):
        self.shape, self.nnz = mat.shape, int(mat.nnz)
        self.ptr = torch.from_numpy(np.ascontiguousarray(mat.indptr, dtype=np.int64)).pin_memory()
        self.col = torch.from_numpy(np.ascontiguousarray(mat.indices, dtype=np.int32)).pin_memory()
        self.val = torch.from_numpy(np.ascontiguousarray(mat.data, dtype=np.float64)).pin_memory()


def broadcast_csr(engine, mat, device, src: int = 0, group=None):
    """Replicate a host scipy CSR held by ``src`` to every rank's device.  Returns
    (DeviceCsr, (ptr, col, val) tensors).  Device layout: i64 row_ptr, i32 col, f64 val."""
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        meta = [(int(mat.shape[0]), int(mat.shape[1]), int(mat.nnz))]
    dist.broadcast_object_list(meta, src=src, group=group)
    rows, cols, nnz = meta[0]
    if rank == src and isinstance(mat, PinnedCsr):
        ptr, col, val = (t.to(device, non_blocking=True) for t in (mat.ptr, mat.col, mat.val))
    elif rank == src:
        ptr = torch.from_numpy(np.ascontiguousarray(mat.indptr, dtype=np.int64)).to(device)
        col = torch.from_numpy(np.ascontiguousarray(mat.indices, dtype=np.int32)).to(device)
        val = torch.from_numpy(np.ascontiguousarray(mat.data, dtype=np.float64)).to(device)
    else:
        ptr = torch.empty(rows + 1, dtype=torch.int64, device=device)
        col = torch.empty(nnz, dtype=torch.int32, device=device)
        val = torch.empty(nnz, dtype=torch.float64, device=device)
    for t in (ptr, col, val):
        dist.broadcast(t, src=src, group=group)
    if engine is None:  # CPU tests
        return None, (ptr, col, val)
    torch.cuda.current_stream().synchronize()
    d = engine.wrap_device(rows, cols, nnz, ptr.data_ptr(), col.data_ptr(), val.data_ptr(), keepalive=(ptr, col, val))
    return d, (ptr, col, val)


def balanced_bounds(weights: np.ndarray, n_shards: int) -> np.ndarray:
    """Host restatement of spada_b200_plan_shards' split rule (weights = flops + 1 per row); used by
    the CPU tests to check the C ABI's answer and by the gloo tests to shard without a GPU."""
    total = int(weights.sum())
    csum = np.concatenate([[0], np.cumsum(weights.astype(np.int64))])
    bounds = [0]
    for s in range(1, n_shards):
        target = total * s // n_shards
        bounds.append(int(np.searchsorted(csum, target, side="right") - 1))
    bounds.append(len(weights))
    return np.asarray(bounds, dtype=np.int64)


def plan_bounds(engine, da, db, world: int, device, src: int = 0, group=None) -> np.ndarray:
    rank = dist.get_rank(group)
    t = torch.zeros(world + 1, dtype=torch.int64, device=device)
    if rank == src:
        t.copy_(torch.from_numpy(engine.plan_shards(da, db, world)))
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


# (event, owners): engine Results whose pool blocks torch's stream may still be reading
_inflight: "collections.deque" = collections.deque()


def _keep_until_done(tensors) -> None:
    """The views of ``result_views`` alias pool blocks of the engine; the copies and collectives enqueued on torch's
    stream return before they have run.  Hold the owning Results until an event recorded behind that work has
    completed, so that dropping the Result right after the call cannot hand its blocks to the next product early."""
    while _inflight and _inflight[0][0].query():
        _inflight.popleft()
    owners = [o for o in (getattr(t, "_spada_owner", None) for t in tensors) if o is not None]
    if owners:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        _inflight.append((ev, owners))


def allgather_csr(local_ptr: torch.Tensor, local_col: torch.Tensor, local_val: torch.Tensor,
                  group=None, out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None):
    """All ranks end with the whole C.  Shard r holds rows [b_r, b_{r+1}) with a row_ptr that
    starts at 0; the global row_ptr is local + (nnz of the shards before).  The work is enqueued on torch's current
    stream; inputs that are ``result_views`` keep their Result alive until that work has run (``_keep_until_done``)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_ptr.device
    mine = torch.tensor([local_ptr.numel() - 1, local_col.numel()], dtype=torch.int64, device=dev)
    sizes = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    sizes = sizes.view(world, 2).cpu()
    rows = sizes[:, 0].tolist()
    nnzs = sizes[:, 1].tolist()
    row_off = np.concatenate([[0], np.cumsum(rows)])
    nnz_off = np.concatenate([[0], np.cumsum(nnzs)])
    if out is None:
        out = (torch.empty(int(row_off[-1]) + 1, dtype=torch.int64, device=dev),
               torch.empty(int(nnz_off[-1]), dtype=torch.int32, device=dev),
               torch.empty(int(nnz_off[-1]), dtype=torch.float64, device=dev))
    g_ptr, g_col, g_val = out
    g_ptr[:1] = 0
    r0, n0 = int(row_off[rank]), int(nnz_off[rank])
    g_ptr[r0 + 1:r0 + 1 + rows[rank]] = local_ptr[1:] + n0
    g_col[n0:n0 + nnzs[rank]] = local_col
    g_val[n0:n0 + nnzs[rank]] = local_val
    # NCCL has no all-gather-v.  More than two ranks on GPUs: shards are nearly equal in nnz (they are
    # balanced by product count), so pad them to the largest, run NCCL's own all-gather in place on the
    # padded buffers and compact into the final arrays with local copies (measured at N=4: 2x faster than
    # grouped send/recv, 1.4x faster than one broadcast per shard).
    if world > 2 and dev.type == "cuda":
        mx = max(max(nnzs), 1)
        mr = max(max(rows), 1)
        pad_col = torch.empty((world, mx), dtype=torch.int32, device=dev)
        pad_val = torch.empty((world, mx), dtype=torch.float64, device=dev)
        pad_ptr = torch.empty((world, mr), dtype=torch.int64, device=dev)
        pad_col[rank, :nnzs[rank]] = local_col
        pad_val[rank, :nnzs[rank]] = local_val
        pad_ptr[rank, :rows[rank]] = local_ptr[1:] + n0
        dist.all_gather_into_tensor(pad_col.view(-1), pad_col[rank], group=group)
        dist.all_gather_into_tensor(pad_val.view(-1), pad_val[rank], group=group)
        dist.all_gather_into_tensor(pad_ptr.view(-1), pad_ptr[rank], group=group)
        for r in range(world):
            if r == rank:
                continue
            rs, ns = int(row_off[r]), int(nnz_off[r])
            g_ptr[rs + 1:rs + 1 + rows[r]] = pad_ptr[r, :rows[r]]
            g_col[ns:ns + nnzs[r]] = pad_col[r, :nnzs[r]]
            g_val[ns:ns + nnzs[r]] = pad_val[r, :nnzs[r]]
        _keep_until_done((local_ptr, local_col, local_val))
        return g_ptr, g_col, g_val
    # Two ranks (or CPU/gloo): every rank sends its shard to every peer and receives theirs in ONE
    # grouped batch (ncclGroupStart/End).
    ops = []
    my_ptr = g_ptr[r0 + 1:r0 + 1 + rows[rank]]
    my_col = g_col[n0:n0 + nnzs[rank]]
    my_val = g_val[n0:n0 + nnzs[rank]]
    for r in range(world):
        if r == rank:
            continue
        peer = dist.get_global_rank(group, r) if group else r
        rs, ns = int(row_off[r]), int(nnz_off[r])
        if rows[rank]:
            ops.append(dist.P2POp(dist.isend, my_ptr, peer, group))
        if rows[r]:
            ops.append(dist.P2POp(dist.irecv, g_ptr[rs + 1:rs + 1 + rows[r]], peer, group))
        if nnzs[rank]:
            ops.append(dist.P2POp(dist.isend, my_col, peer, group))
            ops.append(dist.P2POp(dist.isend, my_val, peer, group))
        if nnzs[r]:
            ops.append(dist.P2POp(dist.irecv, g_col[ns:ns + nnzs[r]], peer, group))
            ops.append(dist.P2POp(dist.irecv, g_val[ns:ns + nnzs[r]], peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if dev.type == "cuda":
        _keep_until_done((local_ptr, local_col, local_val))
    return g_ptr, g_col, g_val


class PeerGather:
    """Sharded product with the all-gather of C fused into the store (one process per GPU).

    Every rank owns a full-size set of C buffers (``spada_b200_cbuf``) sized by an upper bound of nnz(C) (the
    intermediate-product count); their CUDA IPC handles are exchanged ONCE through the process group, so every rank
    holds peer mappings of all of them.  One product = ``shard_begin`` on this rank's rows (first pass into scratch
    rows, local row_ptr, nnz left on the device) -> ``all_gather_into_tensor`` of the per-rank nnz (8 bytes each,
    NCCL, no host read-back) -> ``shard_finish``: the kernel that places the scratch rows reads the shard's global
    offset from that device array and stores every row into ALL ranks' buffers -- its own through HBM, the peers'
    over NVLink -- so the exchange overlaps the placement instead of following it.  A one-element all-reduce
    afterwards is the barrier that makes every rank's C complete before anybody reads it.

    Ordering between the engine's kernels and the two collectives: when the engine was created on torch's current
    stream (``Engine(stream=torch.cuda.current_stream().cuda_stream)`` inside a ``torch.cuda.stream`` context) the
    stream orders them.  Otherwise -- an engine-owned stream, which is also what handle 0, torch's default stream,
    selects -- the engine's stream and torch's do not see each other, and ``step`` waits on the host at the two
    hand-overs (the local nnz before the all-gather reads it, the gathered nnz before the placement reads them).
    """

    def __init__(self, engine, rows: int, cols: int, capacity: int, device, group=None):
        self.engine, self.group, self.dev = engine, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows, self.cols, self.capacity = rows, cols, capacity
        self.own = engine.cbuf_create(rows, cols, capacity)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.own.export(), group=group)
        self.peers = [engine.cbuf_import(handles[r], rows, cols, capacity) for r in range(self.world) if r != self.rank]
        self.bufs = [self.own] + self.peers           # own first: spada_b200_shard_finish's convention
        self.shard_nnz = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.local_nnz = torch.zeros(1, dtype=torch.int64, device=device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        es = getattr(engine, "stream", None)
        self.stream_shared = es is not None and es == torch.cuda.current_stream(device).cuda_stream
        dist.barrier(group=group)

    def step(self, da, db, row_begin: int, row_end: int, compute_only: bool = False) -> dict:
        """One sharded product; returns the engine's stats of this rank's shard.  On return (after the barrier
        collective, stream-ordered) every rank's buffers hold the whole C."""
        shard = self.engine.shard_begin(da, db, row_begin, row_end, d_nnz_local=self.local_nnz.data_ptr())
        if not self.stream_shared:
            self.engine.synchronize()                       # local_nnz is written before NCCL reads it
        dist.all_gather_into_tensor(self.shard_nnz, self.local_nnz, group=self.group)
        if not self.stream_shared:
            torch.cuda.current_stream(self.dev).synchronize()   # shard_nnz is complete before the placement reads it
        bufs = [self.own] if compute_only else self.bufs
        st = shard.finish(bufs, 0, self.shard_nnz.data_ptr(), self.rank)
        dist.all_reduce(self.flag, group=self.group)   # every rank's stores have landed when this completes
        return st

    def nvlink_bytes(self) -> int:
        """Bytes this rank pushed to its peers in the last step (12 per entry of its shard + 8 per row, per peer)."""
        nnz = int(self.local_nnz.item())
        return (self.world - 1) * 12 * nnz

    def result(self):
        """(row_ptr, col, val) torch views of this rank's copy of the gathered C."""
        p, c, v = self.own.device_ptrs()
        nnz = self.own.nnz
        return (device_view(p, self.rows + 1, torch.int64, self.dev, self.own), device_view(c, nnz, torch.int32, self.dev, self.own),
                device_view(v, nnz, torch.float64, self.dev, self.own))

    def checksum(self):
        """(nnz, sum of row_ptr, sum of col ids, sum of values) of this rank's gathered C, as a device tensor."""
        rp, col, val = self.result()
        return torch.stack([rp[-1].double(), rp.double().sum(), col.double().sum(), val.sum()])

    def close(self):
        for b in self.peers:
            b.free()
        self.peers = []
        dist.barrier(group=self.group)     # nobody frees buffers a peer still has mapped
        self.own.free()
