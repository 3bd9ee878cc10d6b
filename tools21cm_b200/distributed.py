"""Multi-GPU forms of the estimators (one process per GPU, ``torch.distributed``).

The reference has no distributed path; two partitions of its power-spectrum
path exist and both are here:

* **Batches of independent cubes** (coeval boxes at many redshifts, light-cone chunks): cube i goes
  to rank i mod P, no data-path collective, results gathered at the end -- ``batched``.
* **One cube too large or too slow for one GPU**, sharded as x slabs: every rank transforms its
  slab along z and y (``t21_slab_stage1``), ONE all-to-all re-partitions the half spectrum into y
  slabs with x complete, every rank transforms along x fused with the binning (``t21_slab_stage2``),
  and a tiny all-reduce sums the per-bin sums and counts -- ``sharded_power_spectrum_1d``,
  ``sharded_cross_power_spectrum_1d``, ``sharded_power_spectrum_mu``.

The exchange has two implementations (``T21_EXCHANGE=p2p|nccl``):

* ``p2p`` (default): the re-partition is FUSED into the Y pass.  Every rank maps the other ranks' y-slab
  buffers into its address space (CUDA IPC) and the Y-pass kernel stores each transformed row straight
  into the buffer of the rank that owns it, over NVLink 5 / NVSwitch: no send buffer, no collective, the
  transfer overlaps the transform tile by tile.  The three small exchanges of a step -- the common mean
  offset, "every rank has finished its peer stores", the per-bin sums and counts -- are few-KB all-gathers over the
  same peer-mapped memory (``t21_peer_signal`` / ``t21_peer_wait``), summed in rank order on every rank: no NCCL
  call on the step at all, and bit-identical results on all ranks.  Chosen only when every rank sits on one host,
  the world is at most 16 and the peer mappings open; otherwise the ``nccl`` path runs.
* ``nccl``: ``all_to_all`` over NCCL in 4 pieces pipelined against the Z/Y passes; rank r sends rank s the block
  W[x in slab r, y in slab s, :], packed by the Y pass so that the receive buffer *is* the sequence of x slabs
  stage 2 reads (no pack or unpack pass).

Device half spectra are CHUNK-MAJOR: [pitch/LW][x][y][LW] with LW = 8 (fp32) or 4 (fp64)
complex values, i.e. 64-byte pieces of a z row.  A y slab is handed to stage 2 as ``n_xslabs`` consecutive x
slabs, each chunk-major on its own: 1 for the fused exchange, P x pieces for buffers filled by the all-to-all.
Tensors carrying them are shaped [x][y][pitch] for their sizes only (``YSlab``).

``backend`` lets the CPU tests (gloo, world_size 2) replace the two CUDA stages by a numpy model
while exercising exactly this partitioning, packing, offset and reduction code.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
from collections import OrderedDict

import numpy as np

from . import _lib
from . import binspec as _bs


def partition(n_items, world, rank):
    """Indices of the items rank ``rank`` owns: item i -> rank i mod world."""
    return list(range(rank, n_items, world))


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def batched(fn, items, group=None):
    """Run ``fn(item)`` for this rank's share of ``items`` and return every result, in item order,
    on every rank.  ``fn`` is one of the single-GPU estimators (closed over its keyword arguments)."""
    world, rank = _world(group)
    mine = [(i, fn(items[i])) for i in partition(len(items), world, rank)]
    if world == 1:
        return [r for _, r in mine]
    gathered = [None] * world
    _dist().all_gather_object(gathered, mine, group=group)
    out = [None] * len(items)
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out


class YSlab:
    """This rank's y slab of one field after the re-partition: ``data`` holds nx * nyl * pitch complex values
    as ``n_xslabs`` consecutive x slabs (each chunk-major for the CUDA backend, row-major for the CPU model)."""

    def __init__(self, data, nx, nyl, pitch, n_xslabs=1):
        self.data, self.nx, self.nyl, self.pitch, self.n_xslabs = data, nx, nyl, pitch, n_xslabs


# ------------------------------------------------------------------------------------------
# backends: the two local stages
# ------------------------------------------------------------------------------------------
class CudaBackend:
    """Stages run by libt21ps.so on the current CUDA device."""

    def __init__(self):
        from . import power_spectrum as _ps
        self._ps = _ps

    def mean_estimate(self, x):
        """c ~ mean(x) from a strided sample (one kernel of the library), as a float64 device scalar."""
        import torch
        lib = _lib.load()
        out = torch.empty(1, dtype=x.dtype, device=x.device)
        rc = lib.t21_mean_estimate(x.data_ptr(), x.numel(), self._ps._dtype_code(x), out.data_ptr(),
                                   self._ps._stream_ptr(torch, x.device.index))
        _lib.check(rc, "t21_mean_estimate")
        return out.double().reshape(())

    def pitch(self, nz):
        return (nz // 2 + 1 + 15) // 16 * 16

    def stage1(self, x, offset, send_world=0):
        """Z and Y passes of an x slab.  send_world = 0: returns W, nxl * ny * pitch values, chunk-major.
        send_world = P: returns the packed all-to-all send buffer [P][nxl * ny/P * pitch], block s = the
        chunk-major part for rank s (pack fused into the Y pass)."""
        import torch
        lib = _lib.load()
        dev = self._ps._device(x.device.index)
        code = self._ps._dtype_code(x)
        plan = dev.plan(tuple(x.shape), code)
        pitch = lib.t21_slab_pitch(plan)
        cdt = torch.complex64 if x.dtype == torch.float32 else torch.complex128
        w = torch.empty((x.shape[0], x.shape[1], pitch), dtype=cdt, device=x.device)
        send = None
        if send_world:
            send = torch.empty((send_world, x.shape[0] * (x.shape[1] // send_world) * pitch), dtype=cdt,
                               device=x.device)
        rc = lib.t21_slab_stage1(plan, x.data_ptr(), offset.data_ptr() if offset is not None else None,
                                 w.data_ptr(), send.data_ptr() if send is not None else None, int(send_world),
                                 self._ps._stream_ptr(torch, x.device.index))
        _lib.check(rc, "t21_slab_stage1")
        return send if send is not None else w

    def stage1_p2p(self, x, offset, peer_ptrs, world, x0_global):
        """Z and Y passes of an x slab with the re-partition fused into the Y pass: rows are stored straight
        into the peers' y-slab buffers (``peer_ptrs``: ctypes array of ``world`` device pointers)."""
        import torch
        lib = _lib.load()
        dev = self._ps._device(x.device.index)
        code = self._ps._dtype_code(x)
        plan = dev.plan(tuple(x.shape), code)
        pitch = lib.t21_slab_pitch(plan)
        cdt = torch.complex64 if x.dtype == torch.float32 else torch.complex128
        scratch = dev.scratch((x.shape[0], x.shape[1], pitch), cdt)
        rc = lib.t21_slab_stage1_p2p(plan, x.data_ptr(), offset.data_ptr() if offset is not None else None,
                                     scratch.data_ptr(), peer_ptrs, int(world), int(x0_global),
                                     self._ps._stream_ptr(torch, x.device.index))
        _lib.check(rc, "t21_slab_stage1_p2p")

    def stage2(self, cols, iy0, ny_global, nz, offsets, spec_key, spec_builder):
        import torch
        lib = _lib.load()
        w = cols[0].data
        dev = self._ps._device(w.device.index)
        code = _lib.T21_F32 if w.dtype == torch.complex64 else _lib.T21_F64
        nx, nyl = cols[0].nx, cols[0].nyl
        plan = dev.plan((nx, nyl, nz), code)
        spec, cspec, _own = dev.spec(((nx, ny_global, nz), spec_key), spec_builder, segments=True)
        nb = spec.nbins
        need = lib.t21_workspace_bytes(plan, 0, nb)
        ws = dev.workspace(need)
        sums = torch.empty(nb, dtype=torch.float64, device=w.device)
        counts = torch.empty(nb, dtype=torch.int64, device=w.device)
        o = [t.data_ptr() if t is not None else None for t in offsets] + [None]
        rc = lib.t21_slab_stage2(plan, w.data_ptr(), cols[1].data.data_ptr() if len(cols) == 2 else None,
                                 int(cols[0].n_xslabs), int(iy0), int(ny_global), o[0], o[1], C.byref(cspec), sums.data_ptr(),
                                 counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                 self._ps._stream_ptr(torch, w.device.index))
        _lib.check(rc, "t21_slab_stage2")
        return spec, sums, counts


    def stage2_nd(self, cols, iy0, ny_global, nz, offsets, scale, out_dtype):
        """Scaled power of this rank's rows, half spectrum, FFT order: [nx][ny/P][nz/2+1] (a view of rows padded to
        t21_half_pitch elements, which keeps the X pass's 32-byte runs inside DRAM sectors)."""
        import torch
        lib = _lib.load()
        w = cols[0].data
        dev = self._ps._device(w.device.index)
        code = _lib.T21_F32 if w.dtype == torch.complex64 else _lib.T21_F64
        nx, nyl = cols[0].nx, cols[0].nyl
        plan = dev.plan((nx, nyl, nz), code)
        out = torch.empty((nx, nyl, lib.t21_half_pitch(plan)), dtype=out_dtype, device=w.device)
        o = [t.data_ptr() if t is not None else None for t in offsets] + [None]
        rc = lib.t21_slab_stage2_nd(plan, w.data_ptr(), cols[1].data.data_ptr() if len(cols) == 2 else None,
                                    int(cols[0].n_xslabs), int(iy0), int(ny_global), o[0], o[1], float(scale),
                                    out.data_ptr(),
                                    _lib.T21_F32 if out_dtype == torch.float32 else _lib.T21_F64,
                                    self._ps._stream_ptr(torch, w.device.index))
        _lib.check(rc, "t21_slab_stage2_nd")
        return out[:, :, :nz // 2 + 1]


# ------------------------------------------------------------------------------------------
# the sharded estimator
# ------------------------------------------------------------------------------------------
def exchange(w_slab, world, group=None):
    """Row-major [nx/P][ny][pitch] on every rank  ->  row-major [nx][ny/P][pitch] on every rank (pack + one
    all-to-all).  Generic form used by backends without a fused pack (the CPU model)."""
    nxl, ny, pitch = w_slab.shape
    if ny % world:
        raise ValueError("the y axis (%d) must be divisible by the number of ranks (%d)" % (ny, world))
    nyl = ny // world
    if world == 1:
        return YSlab(w_slab, nxl, ny, pitch)
    send = w_slab.view(nxl, world, nyl, pitch).permute(1, 0, 2, 3).contiguous()   # [P][nx/P][ny/P][pitch]
    return YSlab(all_to_all_packed(send, group).view(world * nxl, nyl, pitch), world * nxl, nyl, pitch)


def all_to_all_packed(send, group=None):
    """send [P][...] (block s goes to rank s) -> recv [P][...] (block s came from rank s).  Complex tensors
    travel as (re, im) pairs."""
    import torch
    recv = torch.empty_like(send)
    _dist().all_to_all_single(torch.view_as_real(recv).view(-1), torch.view_as_real(send).view(-1), group=group)
    return recv


def pipelined_exchange(backend, x, offset, world, group=None, chunks=4):
    """Stage 1 and the re-partition of one field, overlapped: the slab is cut into `chunks` pieces along
    x; while NCCL moves piece c over NVLink, the Z and Y passes of piece c+1 run.  The receive buffer
    [P][chunks][piece] is the y slab as P * chunks consecutive x slabs, which is what stage 2 reads."""
    import torch
    nxl, ny, nz = x.shape
    nyl = ny // world
    while chunks > 1 and (nxl % chunks or nxl // chunks < 4):
        chunks //= 2
    nc = nxl // chunks
    pitch = backend.pitch(nz)
    recv = None
    works, keep = [], []
    for c in range(chunks):
        send = backend.stage1(x[c * nc:(c + 1) * nc], offset, send_world=world)     # [P][nc * nyl * pitch]
        if recv is None:
            recv = send.new_empty((world, chunks, send.shape[1]))
        outs = [torch.view_as_real(recv[s, c]) for s in range(world)]
        ins = [torch.view_as_real(send[s]) for s in range(world)]
        works.append(_dist().all_to_all(outs, ins, group=group, async_op=True))
        keep.append(send)
    for w in works:
        w.wait()
    return YSlab(recv, world * nxl, nyl, pitch, n_xslabs=world * chunks)


class _RawCuda:
    """Zero-copy torch view of a raw device allocation (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _group_key(group):
    """A name that is unique per process group for the life of the process (c10d numbers its groups; ids of the Python
    objects are recycled after destroy_process_group + init_process_group)."""
    dist = _dist()
    pg = group if group is not None else dist.group.WORLD
    return str(getattr(pg, "group_name", None) or id(pg))


class _PeerState:
    """What the ranks of ONE process group share over peer-mapped memory (cudaMalloc + CUDA IPC, outside torch's
    allocator): the mail / flag buffer of the small exchanges and the y-slab receive buffers of the fused re-partition.

    Lifetime: states are kept per group name, at most two (creating a third closes the oldest); buffer sets (y slabs,
    n-D output rows) per (fields, geometry, dtype), at most one geometry per set -- a new geometry frees the previous
    buffers after a barrier, so a long-running job that changes cube size does not accumulate 17 GB allocations.  ``close()`` unmaps
    the peers' buffers and frees this rank's own; ``reset_peer_state()`` closes everything (call it before
    ``destroy_process_group``).  All exchanges must be issued on one stream per rank (they are ordered by it)."""
    _states = OrderedDict()
    SLOT = 16384          # bytes per rank and exchange: per-bin sums and counts of up to 1024 bins

    @classmethod
    def get(cls, group, device):
        key = (_group_key(group), device.index)
        st = cls._states.get(key)
        if st is None:
            while len(cls._states) >= 2:
                _, old = cls._states.popitem(last=False)
                old.close()
            st = cls._states[key] = cls(group, device)
        else:
            cls._states.move_to_end(key)
        return st

    def __init__(self, group, device):
        import torch
        self.group, self.device = group, device
        self.world, self.rank = _world(group)
        self.lib = _lib.load()
        self.epoch = 0
        self.sets, self.views, self.parity = {}, {}, 0
        self.ok = self._probe()
        if not self.ok:
            return
        nbytes = int(self.lib.t21_peer_comm_bytes(self.world, self.SLOT))
        with torch.cuda.device(device):
            self.comm_ptr, handle = self._alloc(nbytes)
            self.comm = torch.as_tensor(_RawCuda(self.comm_ptr, (nbytes,), "|u1"), device=device)
            self.comm.zero_()
            torch.cuda.synchronize(device)
            gathered = [None] * self.world
            _dist().all_gather_object(gathered, handle, group=group)        # (also: every rank has zeroed its buffer)
            self.comm_opened, ptrs, good = [], [], True
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(self.comm_ptr)
                    continue
                q = C.c_void_p()
                if self.lib.t21_peer_open(gathered[r], C.byref(q)) != 0:
                    good = False
                    ptrs.append(0)
                else:
                    self.comm_opened.append(q.value)
                    ptrs.append(q.value)
            flags = [None] * self.world
            _dist().all_gather_object(flags, good, group=group)
            if not all(flags):                                               # no peer access somewhere: NCCL path
                self.ok = False
                self.close()
                return
            self.comm_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
            self.timed_out = torch.zeros(1, dtype=torch.int32, device=device)
            off = int(self.lib.t21_peer_comm_mail_offset())
            self.mail = self.comm[off:].view(torch.float64).view(2, self.world, self.SLOT // 8)

    def _probe(self):
        """Peer stores need every rank on one host and at most 16 ranks (kMaxPeers of the Y pass)."""
        if self.world < 2 or self.world > 16:
            return False
        hosts = [None] * self.world
        _dist().all_gather_object(hosts, socket.gethostname(), group=self.group)
        return len(set(hosts)) == 1

    def _alloc(self, nbytes):
        ptr = C.c_void_p()
        hbuf = C.create_string_buffer(64)
        _lib.check(self.lib.t21_peer_alloc(nbytes, C.byref(ptr), hbuf), "t21_peer_alloc")
        return ptr.value, hbuf.raw

    def gather(self, payload=None):
        """All-gather of a float64 vector (or a pure barrier) over the peer mail boxes, ordered on the current stream
        after everything queued so far.  Returns the [world, n] view of this rank's mail area for the exchange (valid
        until the exchange after the next one)."""
        import torch
        self.epoch += 1
        n = 0 if payload is None else int(payload.numel())
        if n * 8 > self.SLOT:
            raise ValueError("payload of %d values does not fit a %d-byte mail slot" % (n, self.SLOT))
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if payload is not None:
            payload = payload.to(torch.float64).contiguous()
        rc = self.lib.t21_peer_signal(self.comm_ptrs.data_ptr(), self.world, self.rank, self.epoch,
                                      payload.data_ptr() if n else None, n * 8, self.SLOT, stream)
        _lib.check(rc, "t21_peer_signal")
        rc = self.lib.t21_peer_wait(self.comm_ptr, self.world, self.epoch, self.timed_out.data_ptr(), 20.0, stream)
        _lib.check(rc, "t21_peer_wait")
        return self.mail[self.epoch & 1, :, :n]

    def _shared(self, name, key, nbytes, count):
        """`count` buffers of `nbytes` on every rank, each mapped by all peers, cached under `name` until `key`
        changes (then the old set is freed after a barrier).  Returns (local pointers [count], per-buffer ctypes
        arrays of the world's addresses [count], the same as device int64 tensors [count])."""
        import torch
        cur = self.sets.get(name)
        if cur is not None and cur["key"] == key:
            return cur["local"], cur["tabs"], cur["tabs_dev"]
        if cur is not None:
            self._free_set(name)
        with torch.cuda.device(self.device):
            local, handles = [], []
            for _ in range(count):
                ptr, h = self._alloc(nbytes)
                local.append(ptr)
                handles.append(h)
            gathered = [None] * self.world
            _dist().all_gather_object(gathered, handles, group=self.group)
            opened, tabs, tabs_dev = [], [], []
            for i in range(count):
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(local[i])
                    else:
                        q = C.c_void_p()
                        _lib.check(self.lib.t21_peer_open(gathered[r][i], C.byref(q)), "t21_peer_open")
                        opened.append(q.value)
                        row.append(q.value)
                tabs.append((C.c_void_p * self.world)(*row))
                tabs_dev.append(torch.tensor(row, dtype=torch.int64, device=self.device))
        self.sets[name] = dict(key=key, local=local, tabs=tabs, tabs_dev=tabs_dev, opened=opened)
        return local, tabs, tabs_dev

    def _free_set(self, name):
        import torch
        cur = self.sets.pop(name, None)
        if cur is None:
            return
        if self.ok and self.epoch > 0:
            self.gather(None)                        # nobody is still storing into the buffers about to go
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for q in cur["opened"]:
                self.lib.t21_peer_close(q)
            for ptr in cur["local"]:
                self.lib.t21_peer_free(ptr)

    def slabs(self, nfields, shape, real_dtype):
        """Ping-pong pairs of y-slab receive buffers for `nfields` fields of this geometry.  Returns
        (complex views [field][parity] of this rank's buffers, pointer tables [field][parity] of all ranks')."""
        import torch
        esz = 4 if real_dtype == torch.float32 else 8
        nbytes = int(np.prod(shape)) * 2 * esz
        key = (nfields, tuple(shape), str(real_dtype))
        local, tabs, _ = self._shared("slabs", key, nbytes, 2 * nfields)
        views = self.views.get("slabs")
        if views is None or views[0] != key:
            v = [[torch.view_as_complex(torch.as_tensor(_RawCuda(local[2 * f + par], tuple(shape) + (2,), "<f%d" % esz),
                                                        device=self.device)) for par in range(2)] for f in range(nfields)]
            self.views["slabs"] = views = (key, v)
        return views[1], [[tabs[2 * f + par] for par in range(2)] for f in range(nfields)]

    def nd_out(self, shape, real_dtype):
        """This rank's rows of a sharded n-D output (peer-mapped: the other ranks store the mirrored half into it) and
        the device table of every rank's."""
        import torch
        esz = 4 if real_dtype == torch.float32 else 8
        key = (tuple(shape), str(real_dtype))
        local, _, tabs_dev = self._shared("nd_out", key, int(np.prod(shape)) * esz, 1)
        views = self.views.get("nd_out")
        if views is None or views[0] != key:
            v = torch.as_tensor(_RawCuda(local[0], tuple(shape), "<f%d" % esz), device=self.device)
            self.views["nd_out"] = views = (key, v)
        return views[1], tabs_dev[0]

    def close(self):
        import torch
        try:
            for name in list(self.sets):
                self._free_set(name)
            self.views = {}
        finally:
            if getattr(self, "comm_ptr", None):
                torch.cuda.synchronize(self.device)
                with torch.cuda.device(self.device):
                    for q in getattr(self, "comm_opened", []):
                        self.lib.t21_peer_close(q)
                    self.lib.t21_peer_free(self.comm_ptr)
                self.comm_ptr = None


def reset_peer_state():
    """Free every peer-mapped buffer (call on all ranks, before destroy_process_group)."""
    while _PeerState._states:
        _, st = _PeerState._states.popitem()
        st.close()


def _norm_slab(x, backend):
    """What the stages read as raw memory: a contiguous, 16-byte aligned float32 / float64 tensor (on a CUDA device for
    the CUDA backend).  Anything else is converted here rather than misread there."""
    import torch
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.ascontiguousarray(x))
    if isinstance(backend, CudaBackend):
        if not x.is_cuda:
            x = x.to("cuda", non_blocking=True)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float64)
    x = x.contiguous()
    if x.dim() != 3:
        raise ValueError("a slab is a 3-D array [nx/P, ny, nz]; got %d dimensions" % x.dim())
    if x.data_ptr() % 16:
        x = x.clone()
    return x


def _forward(slabs, group, backend):
    """Common mean offset, stage 1 and the re-partition.  Returns (cols, offsets, world, rank, shape, peer state)."""
    import torch
    world, rank = _world(group)
    slabs = [_norm_slab(x, backend) for x in slabs]
    x0 = slabs[0]
    if any(x.dtype != x0.dtype or x.shape != x0.shape or x.device != x0.device for x in slabs):
        raise ValueError("the slabs of a cross spectrum must agree in shape, dtype and device")
    nxl, ny, nz = x0.shape
    shape = (nxl * world, ny, nz)
    if ny % world:
        raise ValueError("the y axis (%d) must be divisible by the number of ranks (%d)" % (ny, world))
    nyl = ny // world
    cuda = world > 1 and isinstance(backend, CudaBackend)
    mode = os.environ.get("T21_EXCHANGE", "p2p")      # p2p: fused peer stores; nccl: pipelined all-to-all
    peer = None
    if cuda and mode == "p2p":
        peer = _PeerState.get(group, x0.device)
        if not peer.ok:
            peer = None
    single = x0.dtype == torch.float32
    offsets = [None] * len(slabs)
    if single:
        # the SAME offset on every rank (per-slab offsets would not be a pure DC change): the mean of the ranks' estimates
        m = torch.stack([backend.mean_estimate(x).to(torch.float64) for x in slabs])
        if peer is not None:
            m = peer.gather(m).sum(0) / world                 # summed in rank order: identical bits on every rank
        elif world > 1:
            _dist().all_reduce(m, group=group)
            m = m / world
        offsets = [m[i].reshape(1).to(x0.dtype) for i in range(len(slabs))]
    cols = []
    if peer is not None:
        pitch = backend.pitch(nz)
        local, ptrs = peer.slabs(len(slabs), (world * nxl, nyl, pitch), x0.dtype)
        peer.parity ^= 1
        par = peer.parity
        with torch.cuda.device(x0.device):
            for f, (x, off) in enumerate(zip(slabs, offsets)):
                backend.stage1_p2p(x, off, ptrs[f][par], world, rank * nxl)
        peer.gather(None)                                     # every rank's peer stores have landed
        cols = [YSlab(local[f][par], world * nxl, nyl, pitch) for f in range(len(slabs))]
    else:
        for x, off in zip(slabs, offsets):
            if cuda:
                cols.append(pipelined_exchange(backend, x, off, world, group))
            else:
                w = backend.stage1(x, off)
                cols.append(exchange(w, world, group) if world > 1 else YSlab(w, nxl, ny, w.shape[2]))
    return cols, offsets, world, rank, shape, peer


def _sharded_binned(slabs, spec_key, make_spec, group=None, backend=None, timings=None):
    import torch
    backend = backend or CudaBackend()
    cols, offsets, world, rank, shape, peer = _forward(slabs, group, backend)
    ny, nz = shape[1], shape[2]
    nyl = ny // world
    spec, sums, counts = backend.stage2(cols, rank * nyl, ny, nz, offsets, spec_key, lambda: make_spec(shape))
    nb = sums.numel()
    if world > 1:
        both = torch.cat([sums.to(torch.float64), counts.to(torch.float64)])     # counts < 2^53: exact in float64
        if peer is not None and 2 * nb * 8 <= peer.SLOT:
            both = torch.cat([peer.gather(both).sum(0), peer.timed_out.to(torch.float64)])
        else:
            _dist().all_reduce(both, group=group)
            both = torch.cat([both, both.new_zeros(1)])
        host = both.cpu().numpy()
        if host[-1] != 0.0:
            raise RuntimeError("a peer exchange timed out: some rank never signalled (did every rank make the call?)")
        return spec, shape, host[:nb].copy(), host[nb:2 * nb].copy()
    return spec, shape, sums.cpu().numpy(), counts.cpu().numpy().astype("float")


def _swap(send, peer, rank, group):
    """Exchange one tensor with `peer` (same shape both ways); a local copy when the peer is this rank."""
    import torch
    if peer == rank:
        return send.clone()
    dist = _dist()
    recv = torch.empty_like(send)
    ops = [dist.P2POp(dist.isend, send, peer, group), dist.P2POp(dist.irecv, recv, peer, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return recv


def _sharded_nd(slabs, box_dims, out_dtype, group, backend, copy=True):
    """fftshifted power of a slab-sharded cube, sharded by rows of y.  Returns (rows, y_index): rows is
    [nx][ny/P][nz] (x and z complete and shifted), y_index the fftshifted y position of each local row."""
    import torch
    backend = backend or CudaBackend()
    cols, offsets, world, rank, shape, peer = _forward(slabs, group, backend)
    nx, ny, nz = shape
    nyl = ny // world
    half = backend.stage2_nd(cols, rank * nyl, ny, nz, offsets, _bs.volume_scale(shape, box_dims), out_dtype)
    y_index = (torch.arange(rank * nyl, (rank + 1) * nyl) + ny // 2) % ny
    # The kz < 0 half of row ky is the kz > 0 half of row -ky read at -kx (Hermitian symmetry of a real cube).
    if peer is not None:
        # one kernel: own kz >= 0 part written locally, the mirrored part stored into the owner of row -ky over NVLink
        out, tab = peer.nd_out((nx, nyl, nz), out_dtype)
        base = half._base if half._base is not None else half
        rc = _lib.load().t21_nd_assemble(base.data_ptr(), _lib.T21_F32 if out_dtype == torch.float32 else _lib.T21_F64,
                                         nx, nyl, nz, int(base.shape[2]), rank * nyl, ny, tab.data_ptr(), world, rank, 0,
                                         C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
        _lib.check(rc, "t21_nd_assemble")
        peer.gather(None)                                          # every rank's mirrored rows have landed
        # `out` is the peer-mapped buffer the next sharded n-D call on this group writes again
        return (out.clone() if copy else out), y_index
    # generic path (one rank, no peer access, the CPU model): exchange the mirror rows, assemble with tensor ops.
    # Row 0 of this slab mirrors into row 0 of slab (P - g) mod P, rows 1.. into slab P-1-g reversed.
    nneg = (nz - 1) // 2
    pos = half[:, :, 1:nneg + 1]
    m0 = _swap(pos[:, 0:1].contiguous(), (world - rank) % world, rank, group)
    if nyl > 1:
        m1 = _swap(pos[:, 1:].flip(1).contiguous(), world - 1 - rank, rank, group)
        mirror = torch.cat([m0, m1], dim=1)
    else:
        mirror = m0
    out = torch.empty((nx, nyl, nz), dtype=half.dtype, device=half.device)
    cz, hx = nz // 2, nx // 2
    hs = torch.roll(half, hx, 0)
    out[:, :, cz:cz + nneg + 1] = hs[:, :, 0:nneg + 1]
    if nz % 2 == 0:
        out[:, :, 0] = hs[:, :, nz // 2]                                   # Nyquist plane sits at index 0
    if nneg:
        neg = torch.roll(torch.roll(mirror.flip(0), 1, 0), hx, 0)          # value at -kx, then the x shift
        out[:, :, cz - nneg:cz] = neg.flip(2)
    return out, y_index


def sharded_power_spectrum_nd(x_slab, box_dims=None, group=None, backend=None, out_dtype=None, copy=True, **kwargs):
    """``power_spectrum_nd`` of a cube held as x slabs.  Returns ``(rows, y_index)``: this rank's rows
    ``fftshifted_power[:, y_index, :]`` (all of x and z, already shifted) -- the output stays sharded."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    world, _ = _world(group)
    gshape = (x_slab.shape[0] * world,) + tuple(x_slab.shape[1:])
    return _sharded_nd([x_slab], _bs.get_dims(box_dims, gshape), out_dtype or x_slab.dtype, group, backend, copy)


def sharded_cross_power_spectrum_nd(a_slab, b_slab, box_dims, group=None, backend=None, out_dtype=None, copy=True, **kwargs):
    """``cross_power_spectrum_nd`` of two cubes held as x slabs (BASELINE.json config 5).  ``copy=False`` returns the
    peer-mapped output buffer itself (no 4 N^3 / P byte copy); it is overwritten by the next sharded n-D call."""
    assert tuple(a_slab.shape) == tuple(b_slab.shape)
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    world, _ = _world(group)
    gshape = (a_slab.shape[0] * world,) + tuple(a_slab.shape[1:])
    return _sharded_nd([a_slab, b_slab], _bs.get_dims(box_dims, gshape), out_dtype or a_slab.dtype, group, backend, copy)


def _finish(sums, counts, scale):
    with np.errstate(all="ignore"):
        return sums * scale / counts


def sharded_power_spectrum_1d(x_slab, kbins=100, box_dims=None, return_n_modes=False, binning="log",
                              breakpoint=0.1, group=None, backend=None, timings=None, **kwargs):
    """``power_spectrum_1d`` of a cube held as x slabs: ``x_slab`` is this rank's [nx/P, ny, nz] part
    (rank r owns x in [r nx/P, (r+1) nx/P)).  Same return tuple as the single-GPU function, on every rank."""
    return _sharded_1d([x_slab], kbins, box_dims, return_n_modes, binning, breakpoint, group, backend, timings,
                       kwargs)


def sharded_cross_power_spectrum_1d(a_slab, b_slab, kbins=100, box_dims=None, return_n_modes=False,
                                    binning="log", breakpoint=0.1, group=None, backend=None, timings=None, **kwargs):
    assert tuple(a_slab.shape) == tuple(b_slab.shape)
    return _sharded_1d([a_slab, b_slab], kbins, box_dims, return_n_modes, binning, breakpoint, group, backend,
                       timings, kwargs)


def _sharded_1d(slabs, kbins, box_dims, return_n_modes, binning, breakpoint, group, backend, timings, kwargs):
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    world, _ = _world(group)
    gshape = (slabs[0].shape[0] * world,) + tuple(slabs[0].shape[1:])
    box_dims = _bs.get_dims(box_dims, gshape)
    from .power_spectrum import _freeze
    key = ("1d", _freeze(list(box_dims)), _freeze(kbins), binning, breakpoint)
    spec, shape, sums, counts = _sharded_binned(
        slabs, key, lambda shp: _bs.build_binspec(shp, box_dims, kbins, binning=binning, breakpoint=breakpoint),
        group, backend, timings)
    ps = _finish(sums, counts, _bs.volume_scale(shape, box_dims))
    if return_n_modes:
        return ps, spec.k_centres(), counts
    return ps, spec.k_centres()


def sharded_power_spectrum_mu(x_slab, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None,
                              exclude_zero_modes=True, return_n_modes=False, absolute_mus=True, group=None,
                              backend=None, timings=None, **kwargs):
    """``power_spectrum_mu`` of a cube held as x slabs."""
    if weights is not None:
        raise ValueError("weights are not usable in the reference either (power_spectrum.py:386,433)")
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    world, _ = _world(group)
    gshape = (x_slab.shape[0] * world,) + tuple(x_slab.shape[1:])
    box_dims = _bs.get_dims(box_dims, gshape)
    from .power_spectrum import _freeze
    key = ("mu", _freeze(list(box_dims)), _freeze(kbins), _freeze(mubins), los_axis, bool(absolute_mus),
           bool(exclude_zero_modes))
    spec, shape, sums, counts = _sharded_binned(
        [x_slab], key, lambda shp: _bs.build_binspec(
            shp, box_dims, kbins, binning="log", mubins=mubins, los_axis=los_axis, absolute_mus=absolute_mus,
            exclude_zero_modes=exclude_zero_modes, last_bin_closed=False), group, backend, timings)
    nmu, nk = spec.nmu, spec.nk
    ps = _finish(sums, counts, _bs.volume_scale(shape, box_dims)).reshape(nmu, nk)
    counts = counts.reshape(nmu, nk)
    if return_n_modes:
        return ps, spec.mu_centres(), spec.k_centres(), counts
    return ps, spec.mu_centres(), spec.k_centres()
