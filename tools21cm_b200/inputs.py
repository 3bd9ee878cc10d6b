"""The input side of the power-spectrum path: C2Ray / CubeP3M files straight to the device and
the brightness temperature formed there.

Mirrors the pieces of the reference a power-spectrum run starts with --
``xfrac_file.XfracFile`` (``xfrac_file.py:34-73``), ``density_file.DensityFile`` (``density_file.py:36-70``),
``temperature.calc_dt`` / ``mean_dt`` / ``_dt`` (``temperature.py:51-80,200-222``) and
``statistics.subtract_mean_signal`` (``statistics.py:62-85``) -- with the same names, arguments and numbers, but the
cubes live on the GPU: a file's payload is copied host -> device through the library's pinned ring (``t21_h2d_staged``)
exactly as stored (Fortran order), and ``calc_dt`` of an ``XfracFile`` and a ``DensityFile`` is ONE kernel
(``t21_c2ray_dt``) that reads the two blobs, applies ``dT = C(z) (1 - x_i) rho / <rho>`` and writes the C-ordered cube the
first FFT pass reads.  The reference's float64 copies, its cgs-density array and its two transposed arrays are never
built, and the cube never returns to the host.

Constants are the reference's defaults (``const.py:37-111``, ``conv.py:10-12``).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from . import power_spectrum as _ps

__all__ = ["XfracFile", "DensityFile", "calc_dt", "mean_dt", "subtract_mean_signal"]

# ---- const.py / conv.py defaults -------------------------------------------------------------------------------
h = 0.7
Omega0 = 0.27
OmegaB = 0.044
lam = 1.0 - Omega0
pc = 3.086e18
Mpc = 1e6 * pc
G_grav = 6.6732e-8
m_p = 1.672661e-24
c = 3.0e5
abu_he_mass = 0.2486
H0 = 100.0 * h
H0cgs = H0 * 1e5 / Mpc
rho_crit_0 = 3.0 * H0cgs * H0cgs / (8.0 * np.pi * G_grav)
A10 = 2.85e-15
nu0 = 1.42e3
Tstar = 0.068
lambda0 = c * 1.0e5 / (nu0 * 1.0e6)
num_h_0 = (1 - abu_he_mass) * OmegaB * rho_crit_0 / m_p
meandt = 3.0 * lambda0 ** 3 / (32. * np.pi) * A10 * Tstar * num_h_0 / (H0cgs / h) * 1000.      # const.py:111
nbox_fine = 8000                                                                               # conv.py:12


def mean_dt(z):
    """Mean brightness temperature at redshift z in mK (temperature.py:200-212)."""
    Ez = np.sqrt(Omega0 * (1.0 + z) ** 3 + lam + (1.0 - Omega0 - lam) * (1.0 + z) ** 2)
    return meandt / h * (1.0 + z) ** 2 / Ez


def _redshift_from_name(filename, kind):
    try:
        name = os.path.split(filename)[1]
        return float(name.split('_')[1][:-4]) if kind == "xfrac" else float(name.split('n_')[0])
    except Exception:
        return -1


def _blob_to_device(filename, offset, count, dtype):
    """`count` items of `dtype` at byte `offset` of a file -> a flat CUDA tensor, through the pinned ring."""
    torch = _ps._require_cuda()
    mm = np.memmap(filename, dtype=dtype, mode="r", offset=offset, shape=(count,))
    t = torch.empty(count, dtype=torch.float32 if np.dtype(dtype) == np.float32 else torch.float64, device="cuda")
    rc = _lib.load().t21_h2d_staged(t.data_ptr(), mm.ctypes.data, mm.nbytes, _ps._copy_threads(), _ps._stream_ptr(torch, t.device.index))
    _lib.check(rc, "t21_h2d_staged")
    torch.cuda.current_stream().synchronize()          # the mapping goes away with `mm`
    del mm
    return t


def _f2c(blob, shape, out_dtype=None):
    """Fortran-ordered flat device blob -> C-ordered [nx][ny][nz] tensor (t21_f2c)."""
    torch = _ps._torch()
    out_dtype = out_dtype or blob.dtype
    out = torch.empty(shape, dtype=out_dtype, device=blob.device)
    code = lambda d: _lib.T21_F32 if d == torch.float32 else _lib.T21_F64
    with torch.cuda.device(blob.device):
        rc = _lib.load().t21_f2c(blob.data_ptr(), code(blob.dtype), shape[0], shape[1], shape[2], out.data_ptr(), code(out_dtype),
                                 _ps._stream_ptr(torch, blob.device.index))
    _lib.check(rc, "t21_f2c")
    return out


class XfracFile:
    """A C2Ray ionised-fraction file on the device (reference xfrac_file.py:8-73).  ``xi`` is the C-ordered cube
    [mesh_x][mesh_y][mesh_z] (formed on first use); ``blob`` is the payload as stored (Fortran order)."""

    def __init__(self, filename=None, old_format=False, neutral=False, binary_format=False):
        self.blob = None
        self._xi = None
        self.neutral = bool(neutral)
        if filename:
            self.read_from_file(filename, old_format, neutral=neutral, binary_format=binary_format)

    def read_from_file(self, filename, old_format=False, neutral=False, binary_format=False):
        self.filename = filename
        self.neutral = bool(neutral)
        with open(filename, 'rb') as f:
            if binary_format:                                            # :50-52
                temp_mesh = np.fromfile(f, count=3, dtype='int32')
                self.mesh_x, self.mesh_y, self.mesh_z = (int(v) for v in temp_mesh)
                offset = 12
            else:                                                        # Fortran record markers, :53-55
                temp_mesh = np.fromfile(f, count=6, dtype='int32')
                self.mesh_x, self.mesh_y, self.mesh_z = (int(v) for v in temp_mesh[1:4])
                offset = 24
        n = self.mesh_x * self.mesh_y * self.mesh_z
        # np.fromfile reads to the end of the file and reshape() then demands exactly n items (:57-62)
        dtype = np.float32 if old_format else np.float64
        have = (os.path.getsize(filename) - offset) // np.dtype(dtype).itemsize
        if have != n:
            raise ValueError("cannot reshape array of size %d into shape (%d,%d,%d)" % (have, self.mesh_x, self.mesh_y, self.mesh_z))
        self.blob = _blob_to_device(filename, offset, n, dtype)
        self._xi = None
        self.z = _redshift_from_name(filename, "xfrac")

    @property
    def xi(self):
        if self._xi is None:
            x = _f2c(self.blob, (self.mesh_x, self.mesh_y, self.mesh_z))
            self._xi = 1.0 - x if self.neutral else x                   # :64-65
        return self._xi


class DensityFile:
    """A CubeP3M / C2Ray density file on the device (reference density_file.py:9-70): ``raw_density`` in simulation
    units and ``cgs_density`` in g/cm^3 (comoving), C ordered, formed on first use."""

    def __init__(self, filename=None, old_format=False):
        self.blob = None
        self._raw = None
        if filename:
            self.read_from_file(filename, old_format)

    def read_from_file(self, filename, old_format=False):
        self.filename = filename
        if old_format:                                                   # :45-48
            self.mesh_x = self.mesh_y = self.mesh_z = 203
            offset = 0
        else:
            with open(filename, 'rb') as f:
                temp_mesh = np.fromfile(f, count=3, dtype='int32')
            self.mesh_x, self.mesh_y, self.mesh_z = (int(v) for v in temp_mesh)
            offset = 12
        n = self.mesh_x * self.mesh_y * self.mesh_z
        have = (os.path.getsize(filename) - offset) // 4
        if have != n:
            raise ValueError("cannot reshape array of size %d into shape (%d,%d,%d)" % (have, self.mesh_x, self.mesh_y, self.mesh_z))
        self.blob = _blob_to_device(filename, offset, n, np.float32)
        self._raw = None
        self.conv_factor = rho_crit_0 * (float(self.mesh_x) / float(nbox_fine)) ** 3 * OmegaB       # :56
        self.z = _redshift_from_name(filename, "density")

    @property
    def raw_density(self):
        if self._raw is None:
            self._raw = _f2c(self.blob, (self.mesh_x, self.mesh_y, self.mesh_z))
        return self._raw

    @property
    def cgs_density(self):
        return self.raw_density * self.conv_factor                        # float32, like the reference (:57)


def _as_device(a):
    torch = _ps._require_cuda()
    if isinstance(a, torch.Tensor):
        return a if a.is_cuda else a.to("cuda", non_blocking=True)
    return _ps._h2d(np.ascontiguousarray(a))


def calc_dt(xfrac, dens, z=-1, out_dtype=None):
    """Differential brightness temperature in mK assuming T_s >> T_CMB (reference temperature.py:51-80, _dt :215-222):
    ``dT = mean_dt(z) * (1 - x_i) * rho / rho_mean``, rho_mean = rho_crit_0 * OmegaB when the density is non-negative
    (cgs), else 1 (an overdensity).  Returns a CUDA tensor (float32 unless ``out_dtype`` says otherwise; the reference
    returns float64).  With an ``XfracFile`` and a ``DensityFile`` the whole step is one fused kernel on the raw blobs."""
    torch = _ps._require_cuda()
    out_dtype = out_dtype or torch.float32
    if z < 0:
        z = getattr(xfrac, "z", -1)
        if z < 0:
            z = getattr(dens, "z", -1)
        if z < 0:
            raise Exception('No redshift specified. Could not determine from file.')
    Cdt = mean_dt(z)
    if isinstance(xfrac, XfracFile) and isinstance(dens, DensityFile):
        shape = (xfrac.mesh_x, xfrac.mesh_y, xfrac.mesh_z)
        assert shape == (dens.mesh_x, dens.mesh_y, dens.mesh_z)
        # rho = raw * conv_factor (cgs, float32 in the reference) and rho_mean = rho_crit_0 * OmegaB (raw densities are
        # non-negative): the factor on (1 - x_i) * raw
        factor = Cdt * dens.conv_factor / (rho_crit_0 * OmegaB)
        out = torch.empty(shape, dtype=out_dtype, device=xfrac.blob.device)
        code = lambda d: _lib.T21_F32 if d == torch.float32 else _lib.T21_F64
        with torch.cuda.device(out.device):
            rc = _lib.load().t21_c2ray_dt(xfrac.blob.data_ptr(), code(xfrac.blob.dtype), int(xfrac.neutral), dens.blob.data_ptr(),
                                          code(dens.blob.dtype), float(factor), shape[0], shape[1], shape[2], out.data_ptr(),
                                          code(out_dtype), _ps._stream_ptr(torch, out.device.index))
        _lib.check(rc, "t21_c2ray_dt")
        return out
    xi = _as_device(xfrac.xi if isinstance(xfrac, XfracFile) else xfrac).to(torch.float64)
    rho = _as_device(dens.cgs_density if isinstance(dens, DensityFile) else dens).to(torch.float64)
    rho_mean = rho_crit_0 * OmegaB if float(rho.min().item()) >= 0 else 1
    return (Cdt * (1.0 - xi) * rho / rho_mean).to(out_dtype)


def subtract_mean_signal(signal, los_axis=2):
    """Subtract the mean of every slice along the line of sight (reference statistics.py:62-85), on the device."""
    torch = _ps._torch()
    t = _as_device(signal)
    ax = {0: 0, -3: 0, 1: 1, -2: 1, 2: 2, -1: 2}[los_axis]
    dims = tuple(d for d in (0, 1, 2) if d != ax)
    return t - t.mean(dim=dims, keepdim=True, dtype=torch.float64).to(t.dtype)
