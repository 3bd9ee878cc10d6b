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
  transfer overlaps the transform tile by tile.  A one-element all-reduce orders the stages.
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
        flat = x.reshape(-1)
        step = max(1, flat.numel() // 65536) | 1
        return flat[::step].double().mean()

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
        scratch = torch.empty((x.shape[0], x.shape[1], pitch), dtype=cdt, device=x.device)
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
        spec, cspec, _own = dev.spec(((nx, ny_global, nz), spec_key), spec_builder)
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
        """Scaled power of this rank's rows, half spectrum, FFT order: [nx][ny/P][nz/2+1]."""
        import torch
        lib = _lib.load()
        w = cols[0].data
        dev = self._ps._device(w.device.index)
        code = _lib.T21_F32 if w.dtype == torch.complex64 else _lib.T21_F64
        nx, nyl = cols[0].nx, cols[0].nyl
        plan = dev.plan((nx, nyl, nz), code)
        out = torch.empty((nx, nyl, nz // 2 + 1), dtype=out_dtype, device=w.device)
        o = [t.data_ptr() if t is not None else None for t in offsets] + [None]
        rc = lib.t21_slab_stage2_nd(plan, w.data_ptr(), cols[1].data.data_ptr() if len(cols) == 2 else None,
                                    int(cols[0].n_xslabs), int(iy0), int(ny_global), o[0], o[1], float(scale),
                                    out.data_ptr(),
                                    _lib.T21_F32 if out_dtype == torch.float32 else _lib.T21_F64,
                                    self._ps._stream_ptr(torch, w.device.index))
        _lib.check(rc, "t21_slab_stage2_nd")
        return out


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


class _PeerSlabs:
    """This rank's y-slab receive buffers (a ping-pong pair, cudaMalloc'ed by the library) and the peers'
    pairs mapped into this process with CUDA IPC, so the Y pass can store into them over NVLink."""
    _cache = {}

    @classmethod
    def get(cls, key, shape, real_dtype, device, group):
        hit = cls._cache.get(key)
        if hit is None:
            hit = cls._cache[key] = cls(shape, real_dtype, device, group)
        return hit

    def __init__(self, shape, real_dtype, device, group):
        import torch
        lib = _lib.load()
        world, rank = _world(group)
        esz = 4 if real_dtype == torch.float32 else 8
        nbytes = int(np.prod(shape)) * 2 * esz
        self.local, self.local_ptr, handles = [], [], []
        with torch.cuda.device(device):
            for _ in range(2):
                ptr = C.c_void_p()
                hbuf = C.create_string_buffer(64)
                _lib.check(lib.t21_peer_alloc(nbytes, C.byref(ptr), hbuf), "t21_peer_alloc")
                self.local_ptr.append(ptr.value)
                handles.append(hbuf.raw)
                view = torch.as_tensor(_RawCuda(ptr.value, tuple(shape) + (2,), "<f%d" % esz), device=device)
                self.local.append(view)
            gathered = [None] * world
            _dist().all_gather_object(gathered, handles, group=group)
            self.ptrs = []
            for par in range(2):
                row = []
                for r in range(world):
                    if r == rank:
                        row.append(self.local_ptr[par])
                    else:
                        ptr = C.c_void_p()
                        _lib.check(lib.t21_peer_open(gathered[r][par], C.byref(ptr)), "t21_peer_open")
                        row.append(ptr.value)
                self.ptrs.append((C.c_void_p * world)(*row))
        self.parity = 0

    def next(self):
        self.parity ^= 1
        return self.parity


def fused_exchange(backend, x, offset, world, rank, field, group=None):
    """Stage 1 with the all-to-all fused into the Y pass (peer stores over NVLink), then a one-element
    all-reduce that orders "every rank has finished writing" before stage 2 reads.  Returns the y slab."""
    import torch
    nxl, ny, nz = x.shape
    nyl = ny // world
    pitch = backend.pitch(nz)
    key = (field, world * nxl, nyl, pitch, str(x.dtype), x.device.index, id(group))
    bufs = _PeerSlabs.get(key, (world * nxl, nyl, pitch), x.dtype, x.device, group)
    par = bufs.next()
    backend.stage1_p2p(x, offset, bufs.ptrs[par], world, rank * nxl)
    flag = torch.zeros(1, device=x.device)
    _dist().all_reduce(flag, group=group)
    return YSlab(torch.view_as_complex(bufs.local[par]), world * nxl, nyl, pitch)      # chunk-major, one x slab


def _forward(slabs, group, backend):
    """Common mean offset, stage 1 and the re-partition.  Returns (cols, offsets, world, rank, shape)."""
    import torch
    world, rank = _world(group)
    x0 = slabs[0]
    nxl, ny, nz = x0.shape
    shape = (nxl * world, ny, nz)
    if ny % world:
        raise ValueError("the y axis (%d) must be divisible by the number of ranks (%d)" % (ny, world))
    single = x0.dtype == torch.float32
    offsets = []
    for x in slabs:
        if single:
            # the SAME offset on every rank: per-slab offsets would not be a pure DC change
            m = backend.mean_estimate(x).reshape(1).to(torch.float64)
            if world > 1:
                _dist().all_reduce(m, group=group)
                m = m / world
            offsets.append(m.to(x.dtype))
        else:
            offsets.append(None)
    import os
    cols = []
    cuda = world > 1 and isinstance(backend, CudaBackend)
    mode = os.environ.get("T21_EXCHANGE", "p2p")      # p2p: fused peer stores; nccl: pipelined all-to-all
    for f, (x, off) in enumerate(zip(slabs, offsets)):
        if cuda and mode == "p2p":
            cols.append(fused_exchange(backend, x, off, world, rank, f, group))
        elif cuda:
            cols.append(pipelined_exchange(backend, x, off, world, group))
        else:
            w = backend.stage1(x, off)
            cols.append(exchange(w, world, group) if world > 1 else YSlab(w, nxl, ny, w.shape[2]))
    return cols, offsets, world, rank, shape


def _sharded_binned(slabs, spec_key, make_spec, group=None, backend=None, timings=None):
    backend = backend or CudaBackend()
    cols, offsets, world, rank, shape = _forward(slabs, group, backend)
    ny, nz = shape[1], shape[2]
    nyl = ny // world
    spec, sums, counts = backend.stage2(cols, rank * nyl, ny, nz, offsets, spec_key, lambda: make_spec(shape))
    if world > 1:
        _dist().all_reduce(sums, group=group)
        _dist().all_reduce(counts, group=group)
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


def _sharded_nd(slabs, box_dims, out_dtype, group, backend):
    """fftshifted power of a slab-sharded cube, sharded by rows of y.  Returns (rows, y_index): rows is
    [nx][ny/P][nz] (x and z complete and shifted), y_index the fftshifted y position of each local row."""
    import torch
    backend = backend or CudaBackend()
    cols, offsets, world, rank, shape = _forward(slabs, group, backend)
    nx, ny, nz = shape
    nyl = ny // world
    half = backend.stage2_nd(cols, rank * nyl, ny, nz, offsets, _bs.volume_scale(shape, box_dims), out_dtype)
    # The kz < 0 half of row ky is the kz > 0 half of row -ky read at -kx (Hermitian symmetry of a real
    # cube).  Row 0 of this slab mirrors into row 0 of slab (P - g) mod P, rows 1.. into slab P-1-g reversed.
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
    y_index = (torch.arange(rank * nyl, (rank + 1) * nyl) + ny // 2) % ny
    return out, y_index


def sharded_power_spectrum_nd(x_slab, box_dims=None, group=None, backend=None, out_dtype=None, **kwargs):
    """``power_spectrum_nd`` of a cube held as x slabs.  Returns ``(rows, y_index)``: this rank's rows
    ``fftshifted_power[:, y_index, :]`` (all of x and z, already shifted) -- the output stays sharded."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    world, _ = _world(group)
    gshape = (x_slab.shape[0] * world,) + tuple(x_slab.shape[1:])
    return _sharded_nd([x_slab], _bs.get_dims(box_dims, gshape), out_dtype or x_slab.dtype, group, backend)


def sharded_cross_power_spectrum_nd(a_slab, b_slab, box_dims, group=None, backend=None, out_dtype=None, **kwargs):
    """``cross_power_spectrum_nd`` of two cubes held as x slabs (BASELINE.json config 5)."""
    assert tuple(a_slab.shape) == tuple(b_slab.shape)
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    world, _ = _world(group)
    gshape = (a_slab.shape[0] * world,) + tuple(a_slab.shape[1:])
    return _sharded_nd([a_slab, b_slab], _bs.get_dims(box_dims, gshape), out_dtype or a_slab.dtype, group, backend)


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
