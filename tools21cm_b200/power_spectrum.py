"""Drop-in replacements for the estimators of tools21cm's ``power_spectrum.py``.

Same names, positional order, defaults, ``**kwargs(boxsize)`` alias and return tuples as
the reference (``/root/reference/src/tools21cm/power_spectrum.py``, cited per function).
The arithmetic runs on the GPU through the C ABI in ``include/t21ps.h`` (hand-written
sm_100a kernels, no cuFFT, no CPU fallback): this module only validates arguments,
builds the float64 bin tables (``binspec.py``), owns device memory through torch
tensors, and divides the returned sums by the mode counts.

Inputs may be numpy arrays (results come back as numpy float64, exactly like the
reference) or torch tensors on the CPU or on a CUDA device (no host round trip for the
cube; small outputs still come back as numpy float64, n-D outputs stay on the device).
float32 input is transformed in fp32, everything else in fp64 (the reference always
uses float64; ``set_precision`` overrides the choice).
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import numpy as np

from . import _lib
from . import binspec as _bs
from ._lib import T21_F32, T21_F64

__all__ = ["power_spectrum_nd", "cross_power_spectrum_nd", "radial_average", "power_spectrum_1d", "power_spectrum_2d",
           "cross_power_spectrum_1d", "power_spectrum_mu", "cross_power_spectrum_mu", "mu_binning",
           "dimensionless_ps", "anisotropy_ratio_r_mu", "get_k", "get_kbins", "apply_window",
           "set_precision"]

_PRECISION = None  # None: follow the input dtype; 'float32' / 'float64': force


def set_precision(precision=None):
    """Force the transform precision ('float32', 'float64') or follow the input dtype (None)."""
    global _PRECISION
    if precision not in (None, "float32", "float64"):
        raise ValueError("precision must be None, 'float32' or 'float64'")
    _PRECISION = precision


# ----------------------------------------------------------------------------------------
# device state: plans, workspace, uploaded bin specs (one per CUDA device)
# ----------------------------------------------------------------------------------------
def _torch():
    import torch
    return torch


class _Device:
    def __init__(self, index):
        self.index = index
        self.plans = {}
        self.specs = OrderedDict()
        self.cyls = OrderedDict()
        self.cylsegs = OrderedDict()
        self._seg_lru = []
        self.ws = {}
        self._scratch = {}

    def plan(self, shape, dtype):
        key = (tuple(shape), dtype)
        p = self.plans.get(key)
        if p is None:
            lib = _lib.load()
            h = C.c_void_p()
            _lib.check(lib.t21_plan_create(C.byref(h), shape[0], shape[1], shape[2], dtype, self.index),
                       "t21_plan_create")
            p = self.plans[key] = h
        return p

    def workspace(self, nbytes):
        """The scratch buffer of the calls issued on the CURRENT stream: one per stream, so that estimators running on
        different streams (or host threads with their own streams) never share half spectra or partial histograms."""
        torch = _torch()
        key = torch.cuda.current_stream(self.index).cuda_stream
        ws = self.ws.get(key)
        if ws is None or ws.numel() < nbytes:
            self.ws.pop(key, None)
            ws = self.ws[key] = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda:%d" % self.index)
        return ws

    def scratch(self, shape, dtype):
        """A reusable per-stream tensor for stage outputs nobody reads after the call (the Y pass's in-place copy of a
        slab whose result goes to the peers)."""
        torch = _torch()
        key = (torch.cuda.current_stream(self.index).cuda_stream, dtype)
        n = int(np.prod(shape))
        t = self._scratch.get(key)
        if t is None or t.numel() < n:
            self._scratch.pop(key, None)
            t = self._scratch[key] = torch.empty(n, dtype=dtype, device="cuda:%d" % self.index)
        return t[:n].view(shape)

    def spec(self, key, builder, segments=False):
        hit = self.specs.get(key)
        if hit is not None:
            self.specs.move_to_end(key)
            if segments:
                self._attach_segments(hit)
            return hit
        torch = _torch()
        spec = builder()

        def upload(blob):
            t = torch.from_numpy(blob).to("cuda:%d" % self.index)
            return t.data_ptr(), t

        cspec, owner = _lib.pack_binspec(spec, upload)
        hit = (spec, cspec, [owner])
        self.specs[key] = hit
        while len(self.specs) > 64:
            self.specs.popitem(last=False)
        if segments:
            self._attach_segments(hit)
        return hit

    def _segment_table(self, entries_per_stride, build, max_overflow=0.0):
        """Run a segment-table builder with 32 entries per pencil, then again with 8 or 16 if that is all any pencil
        needs.  Returns (int32 tensor, stride, overflow pencils) or (None, 0, 0) when more than ``max_overflow`` of the
        pencils need over 32 ranges (the few that do are marked in the table and classified mode by mode)."""
        torch = _torch()
        dev = "cuda:%d" % self.index
        stats = torch.zeros(2, dtype=torch.int32, device=dev)
        stream = _stream_ptr(torch, self.index)
        stride = 32
        for _ in range(2):
            table = torch.empty(entries_per_stride * stride, dtype=torch.int32, device=dev)
            build(table.data_ptr(), stride, stats.data_ptr(), stream)
            m, novf = (int(v) for v in stats.tolist())
            if m > stride:
                if stride == 32 and novf <= max_overflow * entries_per_stride:
                    return table, stride, novf
                return None, 0, 0
            want = 8 if m <= 8 else (16 if m <= 16 else 32)
            if want == stride:
                break
            stride = want
        return table, stride, 0

    def _attach_segments(self, hit):
        """The pencil segment table of the warp-per-pencil X pass (``t21_build_pencil_segments``): along an x pencil the
        bins are index ranges, found once per (shape, bin spec) on the device with the per-mode classification of the
        other kernels and kept with the cached spec -- plan data like the twiddle tables, 8-32 entries per pencil.
        nx from 32 to 4096 and at most 128 bins get one (every X pass reads it: warp-per-pencil, CTA-per-tile, mixed
        radix); four for signed mu bins, which are not even in kx."""
        spec, cspec, owners = hit
        if getattr(spec, "_segments_tried", False):
            return
        spec._segments_tried = True
        if os.environ.get("T21_X32_SEGMENTS", "1") == "0":
            return
        nx = int(spec.shape[0])
        if nx < 32 or nx > 4096 or spec.nbins > 128 or int(spec.zero_idx[0]) != 0 or len(spec.shape) != 3:
            return
        ntab = 4 if (spec.nmu and not spec.absolute_mus) else 1      # signed mu: own / twin bins for kx >= 0 and kx < 0
        lib = _lib.load()
        npencils = int(spec.shape[1]) * (int(spec.shape[2]) // 2 + 1)
        # (pencils next to the kx axis cross every bin: with a fine binning a few need more than 32 ranges; the
        # warp-per-pencil kernel classifies those mode by mode, the other X passes only take tables without any)
        table, stride, novf = self._segment_table(ntab * npencils, lambda ptr, st, mx, stream: _lib.check(
            lib.t21_build_pencil_segments(C.byref(cspec), ptr, st, mx, stream), "t21_build_pencil_segments"),
            max_overflow=float(os.environ.get("T21_SEG_MAX_OVERFLOW", "0.02")) if nx == 1024 else 0.0)
        if table is None:
            return                          # too many pencils need more ranges than a table holds: classify per mode
        cspec.seg_overflow = novf
        owners.append(table)
        cspec.seg_tab = table.data_ptr()
        cspec.seg_stride = stride
        cspec.seg_ntab = ntab
        # tables are tens of MB (hundreds at 2048^3): only the most recently built few stay attached; an older spec that
        # comes back rebuilds its table
        self._seg_lru.append(hit)
        while len(self._seg_lru) > 8:
            o_spec, o_cspec, o_owners = self._seg_lru.pop(0)
            o_cspec.seg_tab, o_cspec.seg_stride, o_cspec.seg_ntab, o_cspec.seg_overflow = None, 0, 0, 0
            del o_owners[1:]
            o_spec._segments_tried = False


    def cyl_segments(self, key, cyl, perp, par, nu_axis):
        """(table-only BinSpecC, table) for the fused power_spectrum_2d, or None: the pencil segment table of the
        (k_perp, k_par) bins (``t21_build_pencil_segments_cyl``), cached beside the separable tables it is made from."""
        hit = self.cylsegs.get(key)
        if hit is None:
            if os.environ.get("T21_X32_SEGMENTS", "1") == "0":
                return None
            lib = _lib.load()
            n0, n1, n2 = (int(v) for v in key[0])
            table, stride, _novf = self._segment_table(n1 * (n2 // 2 + 1), lambda ptr, st, mx, stream: _lib.check(
                lib.t21_build_pencil_segments_cyl(n0, n1, n2, nu_axis, perp.data_ptr(), par.data_ptr(), cyl.nperp, cyl.npar,
                                                  ptr, st, mx, stream), "t21_build_pencil_segments_cyl"))
            if table is None:
                hit = (None, None)
            else:
                c = _lib.BinSpecC()
                for d in range(3):
                    c.n[d] = key[0][d]
                    c.nyq_idx[d] = -1
                c.nk, c.nmu, c.los_axis = cyl.nperp * cyl.npar, 0, -1
                c.seg_tab, c.seg_stride, c.seg_ntab = table.data_ptr(), stride, 1
                hit = (c, table)
            self.cylsegs[key] = hit
            while len(self.cylsegs) > 8:
                self.cylsegs.popitem(last=False)
        else:
            self.cylsegs.move_to_end(key)
        return hit if hit[0] is not None else None

    def cyl(self, key, builder):
        """(CylSpec, perp table, par table) of power_spectrum_2d, tables resident on the device."""
        hit = self.cyls.get(key)
        if hit is None:
            torch = _torch()
            c = builder()
            hit = (c, torch.from_numpy(c.perp_tab).to("cuda:%d" % self.index),
                   torch.from_numpy(c.par_tab).to("cuda:%d" % self.index))
            self.cyls[key] = hit
            while len(self.cyls) > 8:
                self.cyls.popitem(last=False)
        else:
            self.cyls.move_to_end(key)
        return hit


_DEVICES = {}


def _device(index):
    d = _DEVICES.get(index)
    if d is None:
        d = _DEVICES[index] = _Device(index)
    return d


def _require_cuda():
    torch = _torch()
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("tools21cm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


class _HostCube:
    """A cube still in host memory, handed to the fused estimator as it is: ``t21_ps_binned_host`` copies it in groups of
    x planes while the Z and Y passes already run on the groups that have landed.  Carries what the estimators read from
    a staged tensor (shape, dtype, device)."""
    is_host = True

    def __init__(self, owner, ptr, shape, dtype, pinned, device):
        self.owner, self.ptr, self.shape, self.dtype, self.pinned, self.device = owner, ptr, tuple(shape), dtype, pinned, device

    def dim(self):
        return len(self.shape)


def _host_feed_ok(shape, nbytes):
    """Cubes worth overlapping: 3-D, power-of-two axes the Stockham passes take, at least 64 MB."""
    if os.environ.get("T21_HOST_FEED", "1") == "0" or len(shape) != 3 or nbytes < (64 << 20):
        return False
    return all(8 <= n <= 4096 and (n & (n - 1)) == 0 for n in shape)


def _stage(arr, lazy=False):
    """-> (contiguous CUDA tensor of float32/float64, was_numpy).  Never copies back.  ``lazy``: a large host cube is
    returned as a ``_HostCube`` for the estimators that overlap its copy with their first passes."""
    torch = _require_cuda()
    was_numpy = not isinstance(arr, torch.Tensor)
    if was_numpy:
        a = np.asarray(arr)
        if _PRECISION is not None:
            want = np.dtype(_PRECISION)
        else:
            want = np.dtype(np.float32) if a.dtype == np.float32 else np.dtype(np.float64)
        a = np.ascontiguousarray(a, dtype=want)
        if lazy and _host_feed_ok(a.shape, a.nbytes) and a.ctypes.data % 16 == 0:
            dev = torch.device("cuda", torch.cuda.current_device())
            return _HostCube(a, a.ctypes.data, a.shape, torch.float32 if want == np.float32 else torch.float64, False, dev), True
        t = _h2d(a)
    else:
        t = arr
        if _PRECISION is not None:
            want = torch.float32 if _PRECISION == "float32" else torch.float64
        else:
            want = torch.float32 if t.dtype == torch.float32 else torch.float64
        if (lazy and not t.is_cuda and t.dtype == want and t.is_contiguous() and t.is_pinned() and t.data_ptr() % 16 == 0
                and _host_feed_ok(tuple(t.shape), t.numel() * t.element_size())):
            dev = torch.device("cuda", torch.cuda.current_device())
            return _HostCube(t, t.data_ptr(), t.shape, want, True, dev), False
        if not t.is_cuda:
            t = t.to("cuda", non_blocking=True)
        if t.dtype != want:
            t = t.to(want)
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t, was_numpy


def _copy_threads():
    """Host threads that fill the pinned ring: all cores for one process, a fair share when torchrun put several ranks on
    the host (LOCAL_WORLD_SIZE)."""
    import os
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(2, min(32, (os.cpu_count() or 8) // ranks))


def _h2d(a):
    """A contiguous numpy array -> a CUDA tensor.  numpy memory is pageable: large arrays go through the library's pinned
    ring filled by several host threads (t21_h2d_staged) instead of one blocking cudaMemcpy from pageable memory."""
    torch = _torch()
    if a.nbytes < (8 << 20) or not a.flags.c_contiguous:
        return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", non_blocking=False)
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype if a.ndim else torch.from_numpy(a.reshape(1)).dtype, device="cuda")
    rc = _lib.load().t21_h2d_staged(t.data_ptr(), a.ctypes.data, a.nbytes, _copy_threads(), _stream_ptr(torch, t.device.index))
    _lib.check(rc, "t21_h2d_staged")
    return t


def _dtype_code(t):
    return T21_F32 if t.dtype == _torch().float32 else T21_F64


def _freeze(v):
    if isinstance(v, np.ndarray):
        return ("nd", v.dtype.str, v.shape, v.tobytes())
    if isinstance(v, (list, tuple)):
        return tuple(_freeze(x) for x in v)
    return v


def _stream_ptr(torch, dev_index):
    return C.c_void_p(torch.cuda.current_stream(dev_index).cuda_stream)


def _fused_binned(fields, spec_key, spec_builder):
    """Run t21_ps_binned on one or two staged cubes.  Returns (spec, sums, counts) as numpy."""
    torch = _torch()
    x = fields[0]
    if x.dim() != 3:
        raise NotImplementedError("the fused estimator handles 3-dimensional cubes; got %d-D" % x.dim())
    dev = _device(x.device.index)
    with torch.cuda.device(x.device):
        lib = _lib.load()
        code = _dtype_code(x)
        plan = dev.plan(tuple(x.shape), code)
        spec, cspec, _owner = dev.spec((tuple(x.shape), spec_key), spec_builder, segments=True)
        nb = spec.nbins
        need = lib.t21_workspace_bytes(plan, len(fields), nb)
        ws = dev.workspace(need)
        # sums (float64) and counts (int64) share one buffer: one device->host read per call, no glue kernels
        both = torch.empty(2 * nb, dtype=torch.float64, device=x.device)
        sums, counts = both[:nb], both[nb:].view(torch.int64)
        if getattr(x, "is_host", False):
            # the cube is still on the host: copy it in groups of x planes while Z and Y run on the ones that have landed
            xdev = torch.empty(x.shape, dtype=x.dtype, device=x.device)
            rc = lib.t21_ps_binned_host(plan, x.ptr, int(x.pinned), 1 if code == T21_F32 else 0, C.byref(cspec),
                                        sums.data_ptr(), counts.data_ptr(), xdev.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _copy_threads(), _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_ps_binned_host")
        else:
            x2 = fields[1].data_ptr() if len(fields) == 2 else None
            rc = lib.t21_ps_binned(plan, x.data_ptr(), x2, 1 if code == T21_F32 else 0, C.byref(cspec),
                                   sums.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_ps_binned")
        host = both.cpu().numpy()
    return spec, host[:nb].copy(), host[nb:].view(np.int64).astype(np.float64)


def _fused_nd(fields, scale, out_dtype):
    torch = _torch()
    shape_in = tuple(fields[0].shape)
    if len(shape_in) > 3 or len(shape_in) < 1:
        raise NotImplementedError("power spectra of %d-dimensional arrays are not supported" % len(shape_in))
    if len(shape_in) < 3:       # 1-D / 2-D arrays are 3-D arrays with leading axes of length 1
        fields = [f.reshape((1,) * (3 - len(shape_in)) + shape_in) for f in fields]
        return _fused_nd(fields, scale, out_dtype).reshape(shape_in)
    x = fields[0]
    dev = _device(x.device.index)
    with torch.cuda.device(x.device):
        lib = _lib.load()
        code = _dtype_code(x)
        plan = dev.plan(tuple(x.shape), code)
        need = lib.t21_workspace_bytes(plan, len(fields), 0)
        ws = dev.workspace(need)
        out = torch.empty(tuple(x.shape), dtype=out_dtype, device=x.device)
        x2 = fields[1].data_ptr() if len(fields) == 2 else None
        ocode = T21_F32 if out_dtype == torch.float32 else T21_F64
        if code == T21_F32 and len(fields) == 1 and x.shape[0] == 1024 and os.environ.get("T21_PENCIL32_ND", "1") != "0" and \
                all(8 <= n <= 4096 and (n & (n - 1)) == 0 for n in x.shape[1:]):
            # warp-per-pencil X pass: the half spectrum is written pencil-major (128-byte runs), the assembly pass
            # transposes 32 x 32 tiles of it into the fftshifted cube (both sides 128-byte runs)
            half = torch.empty((x.shape[1], x.shape[2] // 2 + 1, x.shape[0]), dtype=out_dtype, device=x.device)
            rc = lib.t21_ps_half_t(plan, x.data_ptr(), 1, float(scale), half.data_ptr(), ocode, ws.data_ptr(), ws.numel(),
                                   _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_ps_half_t")
            rc = lib.t21_nd_assemble_t(half.data_ptr(), ocode, x.shape[0], x.shape[1], x.shape[2], out.data_ptr(),
                                       _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_nd_assemble_t")
            return out
        if x.shape[2] >= 64:
            # The X pass stores the computed half spectrum (sector-aligned 32-byte runs), then one pass writes the
            # fftshifted array with 128-byte runs both for the kz >= 0 half and for the mirrored kz < 0 half (the
            # same kernel that assembles a sharded output over NVLink, with one rank).  Writing the Hermitian twins from
            # the X pass itself costs 6.9 ms at 1024^3 (misaligned 32-byte runs); this route 4 ms.
            half = torch.empty((x.shape[0], x.shape[1], lib.t21_half_pitch(plan)), dtype=out_dtype, device=x.device)
            rc = lib.t21_ps_half(plan, x.data_ptr(), x2, 1 if code == T21_F32 else 0, float(scale), half.data_ptr(), ocode,
                                 ws.data_ptr(), ws.numel(), _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_ps_half")
            tab = torch.tensor([out.data_ptr()], dtype=torch.int64, device=x.device)
            rc = lib.t21_nd_assemble(half.data_ptr(), ocode, x.shape[0], x.shape[1], x.shape[2], int(half.shape[2]), 0,
                                     x.shape[1], tab.data_ptr(), 1, 0, 1, _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_nd_assemble")
            return out
        rc = lib.t21_ps_nd(plan, x.data_ptr(), x2, 1 if code == T21_F32 else 0, float(scale), out.data_ptr(),
                           ocode, ws.data_ptr(), ws.numel(), _stream_ptr(torch, x.device.index))
        _lib.check(rc, "t21_ps_nd")
    return out


def _fused_half(fields, scale, out_dtype):
    """t21_ps_half: the scaled power of the computed half spectrum, FFT order, [nx][ny][row pitch >= nz//2+1] (device
    tensor; the columns past nz//2 are padding that keeps the rows sector-aligned)."""
    torch = _torch()
    x = fields[0]
    if x.dim() != 3:
        raise NotImplementedError("this estimator handles 3-dimensional cubes; got %d-D" % x.dim())
    dev = _device(x.device.index)
    with torch.cuda.device(x.device):
        lib = _lib.load()
        code = _dtype_code(x)
        plan = dev.plan(tuple(x.shape), code)
        ws = dev.workspace(lib.t21_workspace_bytes(plan, len(fields), 0))
        out = torch.empty((x.shape[0], x.shape[1], lib.t21_half_pitch(plan)), dtype=out_dtype, device=x.device)
        x2 = fields[1].data_ptr() if len(fields) == 2 else None
        rc = lib.t21_ps_half(plan, x.data_ptr(), x2, 1 if code == T21_F32 else 0, float(scale), out.data_ptr(),
                             T21_F32 if out_dtype == torch.float32 else T21_F64, ws.data_ptr(), ws.numel(),
                             _stream_ptr(torch, x.device.index))
        _lib.check(rc, "t21_ps_half")
    return out


def _bin_only(arr_t, spec_key, spec_builder):
    torch = _torch()
    dev = _device(arr_t.device.index)
    with torch.cuda.device(arr_t.device):
        lib = _lib.load()
        spec, cspec, _owner = dev.spec((tuple(arr_t.shape), spec_key), spec_builder)
        nb = spec.nbins
        need = lib.t21_bin_only_workspace_bytes(C.byref(cspec))
        ws = dev.workspace(need)
        sums = torch.empty(nb, dtype=torch.float64, device=arr_t.device)
        counts = torch.empty(nb, dtype=torch.int64, device=arr_t.device)
        rc = lib.t21_bin_only(arr_t.data_ptr(), _dtype_code(arr_t), C.byref(cspec), sums.data_ptr(),
                              counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(torch, arr_t.device.index))
        _lib.check(rc, "t21_bin_only")
        s = sums.cpu().numpy()
        c = counts.cpu().numpy().astype("float")
    return spec, s, c


def _finish(sums, counts, scale):
    with np.errstate(all="ignore"):   # reference __init__.py:94-95: empty bins are NaN, silently
        return sums * scale / counts


# ----------------------------------------------------------------------------------------
# host helpers the reference exposes (numpy, not on the hot path)
# ----------------------------------------------------------------------------------------
_get_dims = _bs.get_dims


def apply_window(input_array, window):
    """Taper along the last axis, in place like the reference (power_spectrum.py:12-21)."""
    if window is None:
        return input_array
    from scipy.signal import windows
    name = window.lower()   # an ndarray window fails here, exactly as in the reference
    if name == "blackmanharris":
        w = windows.blackmanharris(input_array.shape[-1])[None, None, :]
    elif name == "tukey":
        w = windows.tukey(input_array.shape[-1])[None, None, :]
    else:
        w = window
    if isinstance(input_array, np.ndarray):
        input_array *= w
    else:
        torch = _torch()
        input_array *= torch.as_tensor(w, dtype=input_array.dtype, device=input_array.device)
    return input_array


def _get_k(input_array, box_dims):
    """Full k grids on the host, for the sibling modules that import it (power_spectrum.py:449-478)."""
    shape = tuple(input_array.shape)
    nd = len(shape)
    kax = [_bs.k_axis_shifted(shape[d], box_dims[d], nd) for d in range(nd)]
    if nd == 1:
        return [kax[0]], kax[0]
    if nd == 2:
        kx, ky = np.broadcast_arrays(kax[0][:, None], kax[1][None, :])
        return [kx.copy(), ky.copy()], np.sqrt(kx ** 2 + ky ** 2)
    kx, ky, kz = np.broadcast_arrays(kax[0][:, None, None], kax[1][None, :, None], kax[2][None, None, :])
    return [kx.copy(), ky.copy(), kz.copy()], np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)


def _get_mu(k_comp, k, los_axis, absolute_mus):
    """power_spectrum.py:481-502."""
    if los_axis not in (0, 1, 2):
        raise Exception("Your space is not %d-dimensional!" % los_axis)
    with np.errstate(all="ignore"):
        mu = np.abs(k_comp[los_axis] / k) if absolute_mus else k_comp[los_axis] / np.abs(k)
    mu[np.where(k < 0.001)] = np.nan
    return mu


def _get_kbins(kbins, box_dims, k, binning="log", breakpoint=0.1, kmin=None, kmax=None):
    """power_spectrum.py:505-519."""
    if isinstance(kbins, int):
        if kmin is None and kmax is None:
            return _bs.make_k_edges(kbins, box_dims, k.max(), binning, breakpoint)
        kmin = 2.0 * np.pi / min(box_dims) if kmin is None else kmin
        kmax = k.max() if kmax is None else kmax
        if binning == "linear":
            return np.linspace(kmin, kmax, kbins + 1)
        if binning == "log":
            return 10 ** np.linspace(np.log10(kmin), np.log10(kmax), kbins + 1)
        lo = np.linspace(kmin, kmax, kbins + 1)
        hi = 10 ** np.linspace(np.log10(kmin), np.log10(kmax), kbins + 1)
        return np.hstack((lo[lo < breakpoint], hi[hi > breakpoint]))
    return kbins


def _get_nonzero_idx(ps_shape, los_axis):
    """power_spectrum.py:563-575."""
    idx = np.indices(ps_shape)
    perp = [a for a in (0, 1, 2) if a != los_axis] if los_axis in (0, 1) else [0, 1]
    return np.invert((idx[perp[0]] == ps_shape[perp[0]] / 2) * (idx[perp[1]] == ps_shape[perp[1]] / 2))


def get_k(input_array, box_dims):
    """power_spectrum.py:437-439."""
    return _get_k(input_array, _get_dims(box_dims, input_array.shape))


def get_kbins(kbins, box_dims, k=None, array=None, binning="log"):
    """power_spectrum.py:441-445."""
    assert k is not None or array is not None
    box_dims = _get_dims(box_dims, array.shape)
    if k is None:
        _, k = _get_k(array, box_dims)
    return _get_kbins(kbins, box_dims, k, binning=binning)


# ----------------------------------------------------------------------------------------
# n-D spectra
# ----------------------------------------------------------------------------------------
def power_spectrum_nd(input_array, box_dims=None, verbose=False, **kwargs):
    """Power spectrum on the fftshifted grid (reference power_spectrum.py:24-54).

    numpy in -> numpy float64 out.  torch tensor in -> CUDA tensor out (dtype of the
    transform unless ``out_dtype=`` is given)."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array.shape)
    if verbose:
        print("Calculating power spectrum...")
    x, was_numpy = _stage(input_array)
    torch = _torch()
    out_dtype = kwargs.get("out_dtype") or (torch.float64 if was_numpy else x.dtype)
    out = _fused_nd([x], _bs.volume_scale(tuple(x.shape), box_dims), out_dtype)
    if verbose:
        print("...done")
    return out.cpu().numpy() if was_numpy else out


def cross_power_spectrum_nd(input_array1, input_array2, box_dims, **kwargs):
    """Cross power spectrum on the fftshifted grid (reference power_spectrum.py:57-95)."""
    assert (tuple(input_array1.shape) == tuple(input_array2.shape))
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array1.shape)
    a, was_numpy = _stage(input_array1)
    b, _ = _stage(input_array2)
    if b.dtype != a.dtype:
        torch = _torch()
        a, b = a.to(torch.float64), b.to(torch.float64)
    torch = _torch()
    out_dtype = kwargs.get("out_dtype") or (torch.float64 if was_numpy else a.dtype)
    out = _fused_nd([a, b], _bs.volume_scale(tuple(a.shape), box_dims), out_dtype)
    return out.cpu().numpy() if was_numpy else out


# ----------------------------------------------------------------------------------------
# stand-alone binning of caller-supplied fftshifted arrays
# ----------------------------------------------------------------------------------------
def radial_average(input_array, box_dims, kbins=10, binning="log", breakpoint=0.1, **kwargs):
    """Shell average of an fftshifted 1-, 2- or 3-D array (reference power_spectrum.py:98-134).
    Returns (mean, bin centres, n_modes)."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array.shape)
    t, _ = _stage(input_array)
    shape = tuple(t.shape)
    key = ("ravg", _freeze(list(box_dims)), _freeze(kbins), binning, breakpoint)
    spec, sums, counts = _bin_only(t, key, lambda: _bs.build_binspec(
        shape, box_dims, kbins, binning=binning, breakpoint=breakpoint, order="shifted"))
    return _finish(sums, counts, 1.0), spec.k_centres(), counts


def mu_binning(powerspectrum, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None,
               exclude_zero_modes=True, binning="log", absolute_mus=True):
    """(k, mu) averages of an fftshifted 3-D array (reference power_spectrum.py:380-435).
    Like the reference it zeroes the centre sample of its input in place."""
    _reject_weights(weights)
    assert (len(powerspectrum.shape) == 3)
    box_dims = _get_dims(box_dims, powerspectrum.shape)
    centre = tuple((np.array(powerspectrum.shape) / 2).astype(int))
    powerspectrum[centre] = 0.0                                   # :414
    t, _ = _stage(powerspectrum)
    shape = tuple(t.shape)
    key = ("mubin", _freeze(list(box_dims)), _freeze(kbins), _freeze(mubins), los_axis, binning,
           bool(absolute_mus), bool(exclude_zero_modes))
    spec, sums, counts = _bin_only(t, key, lambda: _bs.build_binspec(
        shape, box_dims, kbins, binning=binning, mubins=mubins, los_axis=los_axis, absolute_mus=absolute_mus,
        exclude_zero_modes=exclude_zero_modes, last_bin_closed=False, order="shifted"))
    nmu, nk = spec.nmu, spec.nk
    return (_finish(sums, counts, 1.0).reshape(nmu, nk), spec.mu_centres(), spec.k_centres(),
            counts.reshape(nmu, nk))


# ----------------------------------------------------------------------------------------
# fused estimators
# ----------------------------------------------------------------------------------------
def _shell_estimate(fields, kbins, box_dims, binning, breakpoint):
    shape = tuple(fields[0].shape)
    if len(shape) < 3:
        # 1-D / 2-D: the reference centres k at (n-1)/2 there (power_spectrum.py:458,463) while fftshift
        # puts DC at n//2, so k(-f) != k(f) for even n and the half-spectrum shortcut does not apply.
        # Do exactly what the reference does: full fftshifted power on the device, then radial_average.
        torch = _torch()
        nd = _fused_nd(fields, _bs.volume_scale(shape, box_dims), torch.float64)
        return radial_average(nd, box_dims, kbins=kbins, binning=binning, breakpoint=breakpoint)
    key = ("1d", _freeze(list(box_dims)), _freeze(kbins), binning, breakpoint)
    spec, sums, counts = _fused_binned(fields, key, lambda: _bs.build_binspec(
        shape, box_dims, kbins, binning=binning, breakpoint=breakpoint))
    return _finish(sums, counts, _bs.volume_scale(shape, box_dims)), spec.k_centres(), counts


def _reject_weights(weights):
    """The reference cannot use weights either: an array fails at `weights != None`
    (power_spectrum.py:386), a scalar at `weights[idx]` (:433).  Same exception types here."""
    if weights is not None:
        if np.ndim(weights) > 0:
            raise ValueError("The truth value of an array with more than one element is ambiguous. "
                             "Use a.any() or a.all()")
        raise TypeError("'%s' object is not subscriptable" % type(weights).__name__)


def _mu_estimate(fields, los_axis, mubins, kbins, box_dims, weights, exclude_zero_modes, absolute_mus):
    _reject_weights(weights)
    shape = tuple(fields[0].shape)
    assert (len(shape) == 3)
    key = ("mu", _freeze(list(box_dims)), _freeze(kbins), _freeze(mubins), los_axis, bool(absolute_mus),
           bool(exclude_zero_modes))
    spec, sums, counts = _fused_binned(fields, key, lambda: _bs.build_binspec(
        shape, box_dims, kbins, binning="log", mubins=mubins, los_axis=los_axis, absolute_mus=absolute_mus,
        exclude_zero_modes=exclude_zero_modes, last_bin_closed=False))
    nmu, nk = spec.nmu, spec.nk
    ps = _finish(sums, counts, _bs.volume_scale(shape, box_dims)).reshape(nmu, nk)
    return ps, spec.mu_centres(), spec.k_centres(), counts.reshape(nmu, nk)


def power_spectrum_1d(input_array_nd, kbins=100, box_dims=None, return_n_modes=False, binning="log",
                      breakpoint=0.1, window=None, **kwargs):
    """Spherically averaged power spectrum (reference power_spectrum.py:137-174).
    Returns (Pk, k centres) or (Pk, k centres, n_modes)."""
    input_array_nd = apply_window(input_array_nd, window)
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array_nd.shape)
    x, _ = _stage(input_array_nd, lazy=True)
    ps, bins, n_modes = _shell_estimate([x], kbins, box_dims, binning, breakpoint)
    if return_n_modes:
        return ps, bins, n_modes
    return ps, bins


def cross_power_spectrum_1d(input_array1_nd, input_array2_nd, kbins=100, box_dims=None, return_n_modes=False,
                            binning="log", breakpoint=0.1, **kwargs):
    """Spherically averaged cross power spectrum (reference power_spectrum.py:257-289)."""
    assert (tuple(input_array1_nd.shape) == tuple(input_array2_nd.shape))
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array1_nd.shape)
    a, _ = _stage(input_array1_nd)
    b, _ = _stage(input_array2_nd)
    if a.dtype != b.dtype:
        torch = _torch()
        a, b = a.to(torch.float64), b.to(torch.float64)
    ps, bins, n_modes = _shell_estimate([a, b], kbins, box_dims, binning, breakpoint)
    if return_n_modes:
        return ps, bins, n_modes
    return ps, bins


def power_spectrum_mu(input_array, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None,
                      exclude_zero_modes=True, return_n_modes=False, absolute_mus=True, **kwargs):
    """P(k, mu) (reference power_spectrum.py:292-332; always log k bins, as there).
    Returns (Pk[n_mu, n_k], mu centres, k centres[, n_modes])."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array.shape)
    _reject_weights(weights)
    x, _ = _stage(input_array, lazy=True)
    ps, mu_bins, k_bins, n_modes = _mu_estimate([x], los_axis, mubins, kbins, box_dims, weights,
                                                exclude_zero_modes, absolute_mus)
    if return_n_modes:
        return ps, mu_bins, k_bins, n_modes
    return ps, mu_bins, k_bins


def cross_power_spectrum_mu(input_array1, input_array2, los_axis=0, mubins=20, kbins=10, box_dims=None,
                            weights=None, exclude_zero_modes=True, return_n_modes=False, absolute_mus=True,
                            **kwargs):
    """Cross P(k, mu) (reference power_spectrum.py:335-377)."""
    assert (tuple(input_array1.shape) == tuple(input_array2.shape))
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array1.shape)
    _reject_weights(weights)
    a, _ = _stage(input_array1)
    b, _ = _stage(input_array2)
    if a.dtype != b.dtype:
        torch = _torch()
        a, b = a.to(torch.float64), b.to(torch.float64)
    ps, mu_bins, k_bins, n_modes = _mu_estimate([a, b], los_axis, mubins, kbins, box_dims, weights,
                                                exclude_zero_modes, absolute_mus)
    if return_n_modes:
        return ps, mu_bins, k_bins, n_modes
    return ps, mu_bins, k_bins


# ----------------------------------------------------------------------------------------
# thin host post-processing on top of the estimators
# ----------------------------------------------------------------------------------------
def power_spectrum_2d(input_array, kbins=10, binning="log", box_dims=244 / .7, return_modes=False, nu_axis=2,
                      window=None, **kwargs):
    """Power spectrum binned in (k_perp, k_par) (reference power_spectrum.py:177-254).
    Returns (Pk[n_kper, n_kpar], kper_mid, kpar_mid[, n_modes])."""
    input_array = apply_window(input_array, window)
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = _get_dims(box_dims, input_array.shape)
    x, _ = _stage(input_array)
    torch = _torch()
    shape = tuple(x.shape)
    if len(shape) != 3:
        raise ValueError("power_spectrum_2d needs a 3-D array")
    dev = _device(x.device.index)
    ckey = (shape, _freeze(list(box_dims)), _freeze(kbins), binning, int(nu_axis))
    cyl, perp, par = dev.cyl(ckey, lambda: _bs.build_cyl_tables(shape, box_dims, kbins, binning=binning, nu_axis=nu_axis,
                                                                order="fft"))
    nb = cyl.nperp * cyl.npar
    scale = _bs.volume_scale(shape, box_dims)
    code = _dtype_code(x)
    cseg = dev.cyl_segments(ckey, cyl, perp, par, int(nu_axis)) if (code == T21_F32 and shape[0] == 1024 and nb <= 128) else None
    with torch.cuda.device(x.device):
        lib = _lib.load()
        sums = torch.empty(nb, dtype=torch.float64, device=x.device)
        if cseg is not None:
            # fused: along an x pencil the (k_perp, k_par) bins are index ranges too, so the warp-per-pencil X pass adds
            # them from its registers (t21_ps_binned with a table-only spec); no power array is written
            plan = dev.plan(shape, code)
            ws = dev.workspace(lib.t21_workspace_bytes(plan, 1, nb))
            counts = torch.empty(nb, dtype=torch.int64, device=x.device)
            rc = lib.t21_ps_binned(plan, x.data_ptr(), None, 1, C.byref(cseg[0]), sums.data_ptr(), counts.data_ptr(),
                                   ws.data_ptr(), ws.numel(), _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_ps_binned")
            s = sums.cpu().numpy().reshape(cyl.nperp, cyl.npar) * scale
            n = counts.cpu().numpy().astype(float).reshape(cyl.nperp, cyl.npar)
        else:
            # the power of the computed half spectrum in the transform's own precision (a quarter of the reference's
            # fftshifted float64 cube, :229), then the separable (k_perp, k_par) tables applied to it
            power = _fused_half([x], scale, x.dtype)
            ws = dev.workspace(lib.t21_bin_cyl_workspace_bytes(cyl.nperp, cyl.npar))
            rc = lib.t21_bin_cyl_half(power.data_ptr(), _dtype_code(power), shape[0], shape[1], shape[2], int(power.shape[2]),
                                      int(nu_axis), perp.data_ptr(), par.data_ptr(), cyl.nperp, cyl.npar, sums.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_bin_cyl_half")
            s = sums.cpu().numpy().reshape(cyl.nperp, cyl.npar)
            n = cyl.counts()          # separable bins: (index pairs in the perp bin) x (indices in the par bin), exact
    ps = _finish(s, n, 1.0)
    if return_modes:
        return ps, cyl.kper_mid, cyl.kpar_mid, n
    return ps, cyl.kper_mid, cyl.kpar_mid


def dimensionless_ps(data, kbins=100, box_dims=None, binning="log", factor=10):
    """Delta^2(k) = P k^3 / 2 pi^2, interpolated from a finer binning (reference :536-561)."""
    from scipy.interpolate import interp1d
    pk, ks = power_spectrum_1d(data, kbins=kbins * factor, box_dims=box_dims, binning=binning)
    f = interp1d(ks, pk * ks ** 3 / 2 / np.pi ** 2)
    if binning == "log":
        knew = 10 ** np.linspace(np.log10(ks[1]), np.log10(ks[-1]), kbins)
    else:
        knew = np.linspace(ks[1], ks[-1], kbins)
    return f(knew), knew


def anisotropy_ratio_r_mu(input_array, los_axis=0, mu_cut=0.5, mubins=20, kbins=10, box_dims=None, weights=None,
                          exclude_zero_modes=True, return_n_modes=False, absolute_mus=True, **kwargs):
    """r_mu(k) = <P(mu >= cut)> / <P(mu < cut)> - 1 (reference :578-613).  Returns (r_mu, k centres)."""
    res = power_spectrum_mu(input_array, los_axis=los_axis, mubins=mubins, kbins=kbins, box_dims=box_dims,
                            weights=weights, exclude_zero_modes=exclude_zero_modes, return_n_modes=False,
                            absolute_mus=absolute_mus, **kwargs)
    pk, mus, ks = res
    with np.errstate(all="ignore"):
        up = np.array([p[np.isfinite(p)].mean() for p in pk[mus >= mu_cut, :].T])
        dn = np.array([p[np.isfinite(p)].mean() for p in pk[mus < mu_cut, :].T])
        rmu = up / dn - 1.0
    return rmu, ks
