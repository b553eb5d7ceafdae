"""Multi-GPU sharding of the hot path: one process per GPU (``torch.distributed``).

The path shards in the two ways it does naturally:

* rays / scan positions: contiguous chunks per rank, model descriptor replicated,
  NO communication (``shard_range`` + ``run_to_end_abcd`` on the local chunk);
* field sum: detector ROWS are sharded (tile-aligned blocks), every rank needs every
  beamlet -> the (nb, 12) coefficient table is broadcast from the source rank, each rank
  sums its rows, and the row blocks are all-gathered into the full image.  Sharding the
  beamlets instead would need a reduce-scatter and change the summation order.

Collectives go through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the
CPU tests, where the per-rank compute is injected).

``PeerImage`` is the fused form of the field-sum exchange: every rank owns a full image in
IPC-shared device memory, maps its peers' images over NVLink, and the kernels that produce the
final values (GEMM epilogue / split reduce) store each rank's row block straight into ALL the
images -- no all-gather; a device-side barrier over peer memory closes the step.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

ROW_ALIGN = 32  # the field kernel's tile height: aligned shards keep tile origins identical


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPU cores NVML reports as local to CUDA device ``device_index`` (one process
    per GPU).  Pinned host buffers allocated afterwards land on the GPU's NUMA node, so the host<->device
    copies of the host-buffer entry points (``tg_*_host``) do not cross the socket interconnect when eight
    ranks share the box.  Returns the CPU list, or ``None`` when NVML / the affinity call is unavailable."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ``[begin, end)`` of ``n`` independent units for ``rank``; sizes differ by
    at most one; every unit is owned by exactly one rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def row_shards(H: int, world: int, align: int = ROW_ALIGN):
    """Tile-aligned detector row blocks ``[(row0, nrows), ...]`` for every rank (some may be
    empty when H is small).  Blocks are whole multiples of ``align`` rows except the last."""
    nblocks = (H + align - 1) // align
    out = []
    for r in range(world):
        b0, b1 = shard_range(nblocks, r, world)
        r0, r1 = min(H, b0 * align), min(H, b1 * align)
        out.append((r0, r1 - r0))
    return out


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_table(poly, nb: int, src: int = 0, group=None):
    """Broadcast the beamlet coefficient table (96 B per beamlet) from ``src``."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(poly, src=src, group=group)
    return poly


def gather_rows(local_rows, H: int, W: int, group=None):
    """All-gather the per-rank row blocks into the full ``(H, W)`` image on every rank."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    shards = row_shards(H, world)
    max_rows = max(n for _, n in shards)
    pad = torch.zeros((max_rows, W), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    # complex tensors travel as their real view (gloo / older NCCL builds)
    buf = torch.view_as_real(pad).contiguous() if pad.is_complex() else pad
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for (r0, n), t in zip(shards, out):
        t = torch.view_as_complex(t) if local_rows.is_complex() else t
        parts.append(t[:n])
    return torch.cat(parts, dim=0)


class _CudaBuf:
    """Raw device pointer exposed through ``__cuda_array_interface__`` (zero-copy torch view)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerImage:
    """A full ``(H, W)`` complex image per rank in IPC-shared device memory, with every peer's
    image (and barrier flags) mapped into this process over NVLink (``tg_peer_alloc`` /
    ``tg_peer_open``; the 64-byte handles travel through ``dist.all_gather_object``).

    ``field_sum(poly, nb, grid, ...)`` computes this rank's row block and the producing kernels
    store it into ALL ranks' images (``tg_field_sum_peers``), then ``barrier()`` runs the
    device-side barrier; afterwards ``self.image`` holds the complete image on every rank.
    The image buffer is reused by the next ``field_sum`` (which first runs a barrier so that no rank
    overwrites a peer's image that is still being consumed): use or copy ``image`` on the same stream
    before calling it again.
    One process per GPU, all ranks on one node.  World size 1 works (and is what the 1-GPU tests run).
    """

    FLAG_BYTES = 256

    def __init__(self, H: int, W: int, dtype=None, group=None, device=None):
        import ctypes as C
        import torch
        from . import _lib as L
        dist = _dist()
        self._L, self._lib = L, L.load()
        self.group = group
        inited = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if inited else 1
        self.rank = dist.get_rank(group) if inited else 0
        if self.world > L.TG_MAX_PEERS:
            raise ValueError(f"PeerImage supports up to {L.TG_MAX_PEERS} ranks (one NVLink node)")
        self.dtype = torch.complex128 if dtype is None else dtype
        if self.dtype not in (torch.complex128, torch.complex64):
            raise ValueError("dtype must be complex128 or complex64")
        self.H, self.W = int(H), int(W)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        elt = 16 if self.dtype == torch.complex128 else 8
        self.image_bytes = ((self.H * self.W * elt + 255) // 256) * 256
        with torch.cuda.device(self.device):
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            L.check(self._lib.tg_peer_alloc(self.image_bytes + self.FLAG_BYTES, C.byref(ptr), handle), "tg_peer_alloc")
            self._own = int(ptr.value)
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, bytes(handle), group=group)
            else:
                handles[0] = bytes(handle)
            self._ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._ptrs.append(self._own)
                    continue
                q = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(h)
                L.check(self._lib.tg_peer_open(hb, C.byref(q)), f"tg_peer_open(rank {r})")
                self._ptrs.append(int(q.value))
            if self.world > 1:
                dist.barrier(group=group)      # every rank has zeroed and mapped before anyone signals
        self.epoch = 0
        typestr = "<c16" if self.dtype == torch.complex128 else "<c8"
        self.image = torch.as_tensor(_CudaBuf(self._own, (self.H, self.W), typestr, self), device=self.device)

    def _images(self):
        return self._L.ptr_array(self._ptrs)

    def _flags(self):
        return self._L.ptr_array([p + self.image_bytes for p in self._ptrs])

    def field_sum(self, poly, nb: int, grid, *, cull_bits=None, method="auto"):
        """This rank's tile-aligned row block of the grid field sum, written into every rank's image."""
        import torch
        from .gaussian import DEFAULT_CULL_BITS
        L = self._L
        r0, nr = row_shards(self.H, self.world)[self.rank]
        cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
        if self.epoch > 0 and self.world > 1:
            # the image is reused: no rank may store into its peers' images before every rank has consumed the
            # previous result (their consumers precede this barrier on their streams)
            self.barrier()
        with torch.cuda.device(self.device):
            L.check(self._lib.tg_field_sum_peers(
                int(nb), poly.data_ptr() if nb else None, L.dbl_array(grid.px2m_affine), self.H, self.W, r0, nr,
                self._images(), self.world, self.rank, int(self.dtype == torch.complex128), cull,
                L.TG_METHOD[method], torch.cuda.current_stream().cuda_stream), "tg_field_sum_peers")
        return r0, nr

    def barrier(self):
        """Device-side barrier over peer memory on the current stream (asynchronous for the host)."""
        import torch
        self.epoch += 1
        with torch.cuda.device(self.device):
            self._L.check(self._lib.tg_peer_barrier(self._flags(), self.world, self.rank, self.epoch,
                                                    torch.cuda.current_stream().cuda_stream), "tg_peer_barrier")

    def close(self):
        import torch
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist = _dist()
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)     # nobody unmaps while a peer may still store into it
        with torch.cuda.device(self.device):
            for r, p in enumerate(self._ptrs):
                if r != self.rank:
                    self._lib.tg_peer_close(p)
            self.image = None
            self._lib.tg_peer_free(self._own)
        self._own = None


def make_gaussian_image_sharded(gaussian_rays, model, *, cull_bits=None, out_dtype=None,
                                method="auto", group=None, src: int = 0,
                                rows_fn: Optional[Callable] = None,
                                table_fn: Optional[Callable] = None,
                                peer_image: Optional["PeerImage"] = None):
    """Row-sharded ``make_gaussian_image`` over the ranks of ``group``.

    ``src`` traces the central rays and builds the coefficient table; the table is
    broadcast; every rank sums its row block; the blocks are all-gathered.  ``rows_fn`` /
    ``table_fn`` exist so the plumbing can be exercised on CPU with gloo.

    With ``peer_image`` (a ``PeerImage`` of the detector's shape) the exchange is fused into the
    compute kernels over NVLink peer memory instead: no broadcast, no all-gather; returns
    ``peer_image.image`` (valid on the current stream after the device-side barrier).
    """
    import torch
    dist = _dist()
    from .gaussian import _field_sum_grid, beamlet_polynomials
    grid = model[-1]
    H, W = int(grid.shape[0]), int(grid.shape[1])
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    table_fn = table_fn or (lambda: beamlet_polynomials(gaussian_rays, model))
    poly, nb, dev = table_fn()          # every rank holds the inputs; src's table wins
    if peer_image is not None:
        # fused path: every rank builds the (deterministic) table from the inputs it holds -- no broadcast --
        # and its row block lands in every rank's image through NVLink stores issued by the compute kernels
        peer_image.field_sum(poly, nb, grid, cull_bits=cull_bits, method=method)
        peer_image.barrier()
        return peer_image.image
    poly = broadcast_table(poly, nb, src=src, group=group)
    r0, nr = row_shards(H, world)[rank]
    rows_fn = rows_fn or (lambda p, n, row0, nrows: _field_sum_grid(
        p, n, grid, dev, row0=row0, nrows=nrows, out_dtype=out_dtype, cull_bits=cull_bits,
        method=method))
    local = rows_fn(poly, nb, r0, nr)
    return gather_rows(local, H, W, group=group)
