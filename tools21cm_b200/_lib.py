"""ctypes binding of libt21ps.so (the C ABI declared in include/t21ps.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) and lives at
``tools21cm_b200/lib/libt21ps.so``.  There is no CPU fallback: if the library is
missing, or no CUDA device is present when a compute entry point is called, the
caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

T21_F32, T21_F64 = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("T21PS_LIB") or os.path.join(_HERE, "lib", "libt21ps.so")   # override: kernel experiments


class BinSpecC(C.Structure):
    """Mirror of ``t21_binspec`` in include/t21ps.h (field order matters)."""
    _fields_ = [
        ("n", C.c_int32 * 3),
        ("nk", C.c_int32),
        ("nmu", C.c_int32),
        ("los_axis", C.c_int32),
        ("absolute_mus", C.c_int32),
        ("exclude_zero", C.c_int32),
        ("zero_idx", C.c_int32 * 3),
        ("nyq_idx", C.c_int32 * 3),
        ("tab_s", C.c_void_p * 3),
        ("tab_sf", C.c_void_p * 3),
        ("tab_los", C.c_void_p),
        ("tab_losf", C.c_void_p),
        ("thr", C.c_void_p),
        ("mu_edges", C.c_void_p),
        ("s_nan", C.c_double),
        ("k_lut", C.c_void_p),
        ("k_cells", C.c_int32),
        ("k_lut_c0", C.c_float),
        ("k_lut_inv", C.c_float),
        ("mu_lut", C.c_void_p),
        ("mu_cells", C.c_int32),
        ("k_lo_f", C.c_void_p),
        ("k_hi_f", C.c_void_p),
        ("mu_lo_f", C.c_void_p),
        ("mu_hi_f", C.c_void_p),
        ("seg_tab", C.c_void_p),
        ("seg_stride", C.c_int32),
        ("seg_ntab", C.c_int32),
        ("seg_overflow", C.c_int32),
    ]


def pack_binspec(spec, upload):
    """Flatten a ``binspec.BinSpec`` into (BinSpecC, keepalive).

    All tables go into one 256-byte-aligned blob so a spec costs a single host->device copy.
    ``upload(np.ndarray[uint8]) -> (base_address, owner)`` places the blob where the consumer
    reads it: identity for the CPU emulator, one H2D copy for the CUDA library."""
    parts = []
    size = 0

    def put(a, dtype):
        nonlocal size
        a = np.ascontiguousarray(a, dtype=dtype)
        off = size
        parts.append((off, a))
        size = (off + a.nbytes + 255) // 256 * 256
        return off

    offs = {}
    for d in range(3):
        offs["s%d" % d] = put(spec.tab_s[d], np.float64)
        offs["sf%d" % d] = put(spec.tab_s[d], np.float32)
    offs["los"] = put(spec.tab_los, np.float64)
    offs["losf"] = put(spec.tab_los, np.float32)
    offs["thr"] = put(spec.thr, np.float64)
    offs["mu"] = put(spec.mu_edges if spec.nmu else np.zeros(1), np.float64)
    offs["klut"] = put(spec.k_lut, np.uint16)
    offs["mulut"] = put(spec.mu_lut if spec.nmu else np.zeros(1, np.uint16), np.uint16)
    offs["klo"] = put(spec.k_lo_f, np.float32)
    offs["khi"] = put(spec.k_hi_f, np.float32)
    offs["mlo"] = put(spec.mu_lo_f, np.float32)
    offs["mhi"] = put(spec.mu_hi_f, np.float32)
    blob = np.zeros(max(size, 256), dtype=np.uint8)
    for off, a in parts:
        blob[off:off + a.nbytes] = a.view(np.uint8).reshape(-1)
    base, owner = upload(blob)

    c = BinSpecC()
    for d in range(3):
        c.n[d] = int(spec.shape[d])
        c.zero_idx[d] = int(spec.zero_idx[d])
        c.nyq_idx[d] = int(spec.nyq_idx[d])
        c.tab_s[d] = base + offs["s%d" % d]
        c.tab_sf[d] = base + offs["sf%d" % d]
    c.nk, c.nmu, c.los_axis = int(spec.nk), int(spec.nmu), int(spec.los_axis)
    c.absolute_mus, c.exclude_zero = int(spec.absolute_mus), int(spec.exclude_zero)
    c.tab_los = base + offs["los"]
    c.tab_losf = base + offs["losf"]
    c.thr = base + offs["thr"]
    c.mu_edges = base + offs["mu"]
    c.s_nan = float(spec.s_nan)
    c.k_lut = base + offs["klut"]
    c.k_cells = int(spec.k_lut.size)
    c.k_lut_inv = float(np.float32(spec.k_lut_inv))
    c.k_lut_c0 = float(np.float32(-spec.k_lut_u0 * spec.k_lut_inv))
    c.mu_lut = base + offs["mulut"]
    c.mu_cells = int(spec.mu_lut.size) if spec.nmu else 1
    c.k_lo_f = base + offs["klo"]
    c.k_hi_f = base + offs["khi"]
    c.mu_lo_f = base + offs["mlo"]
    c.mu_hi_f = base + offs["mhi"]
    return c, owner


def host_buffer(a):
    return a.ctypes.data, a


_lib = None


def load():
    """Load libt21ps.so and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "tools21cm_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, sz, dbl = C.c_void_p, C.c_int, C.c_size_t, C.c_double
    lib.t21_plan_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32]
    lib.t21_plan_create.restype = i32
    lib.t21_plan_destroy.argtypes = [vp]
    lib.t21_plan_destroy.restype = i32
    lib.t21_workspace_bytes.argtypes = [vp, i32, i32]
    lib.t21_workspace_bytes.restype = sz
    lib.t21_ps_binned.argtypes = [vp, vp, vp, i32, C.POINTER(BinSpecC), vp, vp, vp, sz, vp]
    lib.t21_ps_binned.restype = i32
    lib.t21_ps_binned_host.argtypes = [vp, vp, i32, i32, C.POINTER(BinSpecC), vp, vp, vp, vp, sz, i32, vp]
    lib.t21_ps_binned_host.restype = i32
    lib.t21_ps_binned_windows.argtypes = [vp, vp, C.POINTER(C.c_int), i32, C.POINTER(BinSpecC), vp, vp, vp, sz, vp]
    lib.t21_ps_binned_windows.restype = i32
    lib.t21_ps_nd.argtypes = [vp, vp, vp, i32, dbl, vp, i32, vp, sz, vp]
    lib.t21_ps_nd.restype = i32
    lib.t21_slab_pitch.argtypes = [vp]
    lib.t21_slab_pitch.restype = i32
    lib.t21_slab_stage1.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    lib.t21_slab_stage1.restype = i32
    lib.t21_peer_alloc.argtypes = [sz, C.POINTER(vp), C.c_char_p]
    lib.t21_peer_alloc.restype = i32
    lib.t21_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.t21_peer_open.restype = i32
    lib.t21_peer_close.argtypes = [vp]
    lib.t21_peer_close.restype = i32
    lib.t21_peer_free.argtypes = [vp]
    lib.t21_peer_free.restype = i32
    lib.t21_peer_comm_bytes.argtypes = [i32, i32]
    lib.t21_peer_comm_bytes.restype = sz
    lib.t21_peer_comm_mail_offset.argtypes = []
    lib.t21_peer_comm_mail_offset.restype = sz
    lib.t21_peer_signal.argtypes = [vp, i32, i32, C.c_ulonglong, vp, i32, i32, vp]
    lib.t21_peer_signal.restype = i32
    lib.t21_peer_wait.argtypes = [vp, i32, C.c_ulonglong, vp, dbl, vp]
    lib.t21_peer_wait.restype = i32
    lib.t21_nd_assemble.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.t21_nd_assemble.restype = i32
    lib.t21_mean_estimate.argtypes = [vp, C.c_longlong, i32, vp, vp]
    lib.t21_mean_estimate.restype = i32
    lib.t21_slab_stage1_p2p.argtypes = [vp, vp, vp, vp, C.POINTER(vp), i32, C.c_longlong, vp]
    lib.t21_slab_stage1_p2p.restype = i32
    lib.t21_slab_stage2.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, C.POINTER(BinSpecC), vp, vp, vp, sz, vp]
    lib.t21_slab_stage2.restype = i32
    lib.t21_slab_stage2_nd.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, dbl, vp, i32, vp]
    lib.t21_slab_stage2_nd.restype = i32
    lib.t21_bin_only.argtypes = [vp, i32, C.POINTER(BinSpecC), vp, vp, vp, sz, vp]
    lib.t21_bin_only.restype = i32
    lib.t21_bin_only_workspace_bytes.argtypes = [C.POINTER(BinSpecC)]
    lib.t21_bin_only_workspace_bytes.restype = sz
    lib.t21_bin_cyl.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp, sz, vp]
    lib.t21_bin_cyl.restype = i32
    lib.t21_bin_cyl_workspace_bytes.argtypes = [i32, i32]
    lib.t21_bin_cyl_workspace_bytes.restype = sz
    lib.t21_ps_half.argtypes = [vp, vp, vp, i32, dbl, vp, i32, vp, sz, vp]
    lib.t21_ps_half.restype = i32
    lib.t21_ps_half_t.argtypes = [vp, vp, i32, dbl, vp, i32, vp, sz, vp]
    lib.t21_ps_half_t.restype = i32
    lib.t21_nd_assemble_t.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.t21_nd_assemble_t.restype = i32
    lib.t21_half_pitch.argtypes = [vp]
    lib.t21_half_pitch.restype = i32
    lib.t21_bin_cyl_half.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, sz, vp]
    lib.t21_bin_cyl_half.restype = i32
    lib.t21_bin_multipoles_half.argtypes = [vp, i32, i32, C.POINTER(BinSpecC), i32, i32, vp, vp, vp, sz, vp]
    lib.t21_bin_multipoles_half.restype = i32
    lib.t21_bin_multipoles_workspace_bytes.argtypes = [i32]
    lib.t21_bin_multipoles_workspace_bytes.restype = sz
    lib.t21_ps_multipoles.argtypes = [vp, vp, i32, C.POINTER(BinSpecC), i32, vp, vp, vp, sz, vp]
    lib.t21_ps_multipoles.restype = i32
    lib.t21_h2d_staged.argtypes = [vp, vp, sz, i32, vp]
    lib.t21_h2d_staged.restype = i32
    lib.t21_f2c.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp]
    lib.t21_f2c.restype = i32
    lib.t21_c2ray_dt.argtypes = [vp, i32, i32, vp, i32, dbl, i32, i32, i32, vp, i32, vp]
    lib.t21_c2ray_dt.restype = i32
    lib.t21_build_pencil_segments.argtypes = [C.POINTER(BinSpecC), vp, i32, vp, vp]
    lib.t21_build_pencil_segments.restype = i32
    lib.t21_build_pencil_segments_cyl.argtypes = [i32, i32, i32, i32, vp, vp, i32, i32, vp, i32, vp, vp]
    lib.t21_build_pencil_segments_cyl.restype = i32
    lib.t21_profile.argtypes = [i32]
    lib.t21_profile.restype = i32
    lib.t21_profile_last.argtypes = [C.POINTER(C.c_float), i32]
    lib.t21_profile_last.restype = i32
    lib.t21_last_error.argtypes = []
    lib.t21_last_error.restype = C.c_char_p
    lib.t21_launch_count.argtypes = []
    lib.t21_launch_count.restype = C.c_ulonglong
    lib.t21_version.argtypes = []
    lib.t21_version.restype = i32
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().t21_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("%s: %s" % (what, msg))
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg))
