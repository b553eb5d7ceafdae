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
    """Two full ``(H, W)`` complex images per rank (ping-pong) in IPC-shared device memory, with every peer's
    images and barrier flags mapped into this process over NVLink (``tg_peer_alloc`` / ``tg_peer_open``; the
    64-byte handles travel through ``dist.all_gather_object``).

    ``field_sum(poly, nb, grid, ...)`` computes this rank's row block and the producing kernels store it into
    the current buffer of ALL ranks (``tg_field_sum_peers``); ``barrier()`` runs the device-side barrier and
    flips the buffers; afterwards ``image`` holds the complete image of that step on every rank.
    ``step(...)`` is both.  Because consecutive steps write different buffers, ONE barrier per step is enough:
    a rank can only start storing step k+1 into a peer's other buffer after it has passed barrier k, which the
    peer enters only after (in stream order) it has consumed the result of step k-1 -- the last user of that
    buffer.  So: use or copy ``image`` on the same stream before the next step.
    The barrier keeps its epoch in device memory (``tg_peer_barrier_auto``), so a whole step can be captured
    into a CUDA graph (``PeerImagePlan``).  A barrier timeout (``barrier_timeout_s``, default
    ``TG_PEER_BARRIER_TIMEOUT_S`` or 10 s) does not trap; ``status()`` reports it.
    One process per GPU, all ranks on one node.  World size 1 works (and is what the 1-GPU tests run).
    """

    FLAG_BYTES = 256

    def __init__(self, H: int, W: int, dtype=None, group=None, device=None, barrier_timeout_s: float = 0.0):
        import ctypes as C
        import torch
        from . import _lib as L
        dist = _dist()
        self._L, self._lib = L, L.load()
        self._own = None
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
        self.barrier_timeout_s = float(barrier_timeout_s)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        elt = 16 if self.dtype == torch.complex128 else 8
        self.image_bytes = ((self.H * self.W * elt + 255) // 256) * 256
        # layout: image 0 | image 1 | flags (npeers uint64) | state {epoch, status} at +128 of the flag block
        total = 2 * self.image_bytes + self.FLAG_BYTES
        with torch.cuda.device(self.device):
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            L.check(self._lib.tg_peer_alloc(total, C.byref(ptr), handle), "tg_peer_alloc")
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
        self.parity = 0
        typestr = "<c16" if self.dtype == torch.complex128 else "<c8"
        self.images = [torch.as_tensor(_CudaBuf(self._own + b * self.image_bytes, (self.H, self.W), typestr, self),
                                       device=self.device) for b in range(2)]
        self._status = torch.as_tensor(_CudaBuf(self._own + 2 * self.image_bytes + 128, (2,), "<u8", self),
                                       device=self.device)
        self.image = self.images[0]

    # -- plumbing
    def _images(self, parity):
        return self._L.ptr_array([p + parity * self.image_bytes for p in self._ptrs])

    def _flags(self):
        return self._L.ptr_array([p + 2 * self.image_bytes for p in self._ptrs])

    def _state_ptr(self):
        return self._own + 2 * self.image_bytes + 128

    def check_grid(self, grid, out_dtype=None):
        H, W = int(grid.shape[0]), int(grid.shape[1])
        if (H, W) != (self.H, self.W):
            raise ValueError(f"PeerImage is {self.H}x{self.W} but the detector is {H}x{W}")
        if out_dtype is not None and out_dtype != self.dtype:
            raise ValueError(f"PeerImage holds {self.dtype} but out_dtype={out_dtype} was requested")

    def rows(self):
        return row_shards(self.H, self.world)[self.rank]

    def field_sum(self, poly, nb: int, grid, *, cull_bits=None, method="auto", parity=None):
        """This rank's tile-aligned row block of the grid field sum, written into buffer ``parity`` (default: the
        current one) of every rank."""
        import torch
        from .gaussian import DEFAULT_CULL_BITS
        L = self._L
        self.check_grid(grid)
        r0, nr = self.rows()
        cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
        par = self.parity if parity is None else int(parity)
        with torch.cuda.device(self.device):
            L.check(self._lib.tg_field_sum_peers(
                int(nb), poly.data_ptr() if nb else None, L.dbl_array(grid.px2m_affine), self.H, self.W, r0, nr,
                self._images(par), self.world, self.rank, int(self.dtype == torch.complex128), cull,
                L.TG_METHOD[method], torch.cuda.current_stream().cuda_stream), "tg_field_sum_peers")
        return r0, nr

    def image_step(self, g_arrays, cm, grid, *, cull_bits=None, method="auto", parity=None):
        """The whole of ``make_gaussian_image`` for this rank's rows with peer stores
        (``tg_make_gaussian_image_peers``); ``g_arrays`` = flat fp64 CUDA tensors of the GaussianRay fields."""
        import ctypes as C
        import torch
        from .gaussian import DEFAULT_CULL_BITS
        from .ray import RAY_FIELDS
        L = self._L
        self.check_grid(grid)
        r0, nr = self.rows()
        cull = DEFAULT_CULL_BITS if cull_bits is None else int(cull_bits)
        par = self.parity if parity is None else int(parity)
        g = g_arrays
        with torch.cuda.device(self.device):
            L.check(self._lib.tg_make_gaussian_image_peers(
                C.byref(cm), g["n"], L.ptr_array([g[f].data_ptr() for f in RAY_FIELDS]), g["amplitude"].data_ptr(),
                g["waist_xy"].data_ptr(), g["radii_of_curv"].data_ptr(), g["wavelength"].data_ptr(),
                g["theta"].data_ptr(), L.dbl_array(grid.px2m_affine), self.H, self.W, r0, nr, self._images(par),
                self.world, self.rank, int(self.dtype == torch.complex128), cull, L.TG_METHOD[method],
                torch.cuda.current_stream().cuda_stream), "tg_make_gaussian_image_peers")

    def barrier(self, flip: bool = True):
        """Device-side barrier over peer memory on the current stream (asynchronous for the host); afterwards
        ``image`` is the buffer the step just completed and the next step writes the other one."""
        import torch
        with torch.cuda.device(self.device):
            self._L.check(self._lib.tg_peer_barrier_auto(self._flags(), self.world, self.rank, self._state_ptr(),
                                                         self.barrier_timeout_s,
                                                         torch.cuda.current_stream().cuda_stream),
                          "tg_peer_barrier_auto")
        if flip:
            self.image = self.images[self.parity]
            self.parity ^= 1

    def step(self, poly, nb: int, grid, *, cull_bits=None, method="auto"):
        self.field_sum(poly, nb, grid, cull_bits=cull_bits, method=method)
        self.barrier()
        return self.image

    def status(self):
        """``(barriers completed, timeout code)`` read back from the device (synchronises): the code is 0, or
        1 + the rank a barrier gave up waiting for."""
        st = self._status.cpu()
        return int(st[0]), int(st[1])

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
            self.images = None
            self._status = None
            self._lib.tg_peer_free(self._own)
        self._own = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        # last resort (interpreter shutdown may already have torn torch.distributed down): a single-rank image
        # can always be released; a multi-rank one is only unmapped by an explicit, collective close()
        try:
            if getattr(self, "_own", None) is not None and self.world == 1:
                self.close()
        except Exception:  # noqa: BLE001
            pass


class PeerImagePlan:
    """Row-sharded ``make_gaussian_image`` of ONE image over all ranks, captured into two CUDA graphs (one per
    ping-pong buffer) and replayed: per step every rank runs the ray kernel, the coefficient kernel and the
    field sum of its row block -- whose final stores go into every rank's image over NVLink -- and ONE
    device-side barrier, with no host work between the kernels.  All ranks must call ``run()`` the same number
    of times.  ``update(rays)`` copies new beamlet parameters into the static buffers (same on all ranks)."""

    def __init__(self, gaussian_rays, model, peer_image: "PeerImage", *, cull_bits=None, method="auto",
                 freeze_dispatch: bool = True):
        import torch
        from .gaussian import _beamlet_arrays, _device_for, auto_dispatch
        if method == "auto" and freeze_dispatch:
            # resolved once, from the inputs every rank holds (same verdict on every rank): the graph then holds the
            # launches of one path only -- see GaussianImagePlan
            method = auto_dispatch(gaussian_rays, list(model), cull_bits)
        self.method = method
        from .gaussian import compile_model
        self.pimg = peer_image
        grid = model[-1]
        peer_image.check_grid(grid)
        dev = _device_for(gaussian_rays)
        self.device = dev
        garr = _beamlet_arrays(gaussian_rays, dev)
        self._g = {k: (v.clone() if hasattr(v, "clone") else v) for k, v in garr.items()}
        cm = compile_model(model)
        kw = dict(cull_bits=cull_bits, method=method)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside the capture (both buffers)
            for par in (0, 1):
                peer_image.image_step(self._g, cm, grid, parity=par, **kw)
                peer_image.barrier(flip=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graphs = []
        for par in (0, 1):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                peer_image.image_step(self._g, cm, grid, parity=par, **kw)
                peer_image.barrier(flip=False)
            self._graphs.append(gr)
        self._parity = 0

    def update(self, gaussian_rays):
        from dataclasses import fields
        from . import _arrays as A
        for f in fields(gaussian_rays):
            dst = self._g[f.name]
            dst.copy_(A.to_device_f64(getattr(gaussian_rays, f.name), self.device).reshape(dst.shape),
                      non_blocking=True)
        return self

    def run(self):
        """Replay one step; returns the complete ``(H, W)`` image (valid on the current stream; overwritten
        by the step after next)."""
        self._graphs[self._parity].replay()
        img = self.pimg.images[self._parity]
        self._parity ^= 1
        return img

    __call__ = run


# Below this many nominal evaluations (beamlets x pixels) one GPU finishes the tensor-core path before the
# exchange of a sharded step would (measured on 8 x B200, bench.py "row_sharded_single_image"): `auto` then
# computes the whole image on every rank instead of sharding it.
SHARD_MIN_EVALS = 4.0e9


def make_gaussian_image_sharded(gaussian_rays, model, *, cull_bits=None, out_dtype=None,
                                method="auto", group=None, src: int = 0,
                                rows_fn: Optional[Callable] = None,
                                table_fn: Optional[Callable] = None,
                                peer_image: Optional["PeerImage"] = None):
    """Row-sharded ``make_gaussian_image`` over the ranks of ``group``.

    ``src`` traces the central rays and builds the coefficient table; the table is
    broadcast; every rank sums its row block; the blocks are all-gathered.  ``rows_fn`` /
    ``table_fn`` exist so the plumbing can be exercised on CPU with gloo.

    With ``peer_image`` (a ``PeerImage`` of the detector's shape and of ``out_dtype``) the exchange is fused
    into the compute kernels over NVLink peer memory instead: no broadcast (every rank builds the same table
    from the inputs it holds; ``src`` is not used), no all-gather; returns ``peer_image.image`` (valid on the
    current stream after the device-side barrier).
    """
    import torch
    dist = _dist()
    from .gaussian import _field_sum_grid, beamlet_polynomials
    grid = model[-1]
    H, W = int(grid.shape[0]), int(grid.shape[1])
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    if peer_image is not None and rows_fn is not None:
        raise ValueError("rows_fn injects the per-rank compute of the NCCL path; the peer-image path runs "
                         "tg_field_sum_peers itself")
    table_fn = table_fn or (lambda: beamlet_polynomials(gaussian_rays, model))
    poly, nb, dev = table_fn()          # every rank holds the inputs; src's table wins
    if peer_image is not None:
        # fused path: every rank builds the (deterministic) table from the inputs it holds -- no broadcast --
        # and its row block lands in every rank's image through NVLink stores issued by the compute kernels
        peer_image.check_grid(grid, out_dtype)
        return peer_image.step(poly, nb, grid, cull_bits=cull_bits, method=method)
    poly = broadcast_table(poly, nb, src=src, group=group)
    r0, nr = row_shards(H, world)[rank]
    rows_fn = rows_fn or (lambda p, n, row0, nrows: _field_sum_grid(
        p, n, grid, dev, row0=row0, nrows=nrows, out_dtype=out_dtype, cull_bits=cull_bits,
        method=method))
    local = rows_fn(poly, nb, r0, nr)
    return gather_rows(local, H, W, group=group)
