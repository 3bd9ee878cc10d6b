"""CPU replay of the CUDA pass bodies (tests/emu/emu.cpp) against numpy and the golden vectors.

The kernels are phase-structured host/device templates; the emulator runs those very
phases thread by thread.  This covers the FFT index arithmetic, twiddle tables, real-FFT
untangling, Hermitian twin weights, DC repair and the classification code without a GPU.
The GPU parity tests proper are in test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cases import CASES, build_inputs, build_kwargs
from conftest import golden_outputs, ROOT
from test_host_model import spec_for
from tools21cm_b200 import _lib, binspec as bs

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libt21emu.so")
POW2 = (8, 16, 32, 64, 128, 256, 1024)


@pytest.fixture(scope="session")
def emu():
    src = os.path.join(EMU_DIR, "emu.cpp")
    deps = [src] + [os.path.join(ROOT, "tools21cm_b200", "csrc", f)
                    for f in os.listdir(os.path.join(ROOT, "tools21cm_b200", "csrc")) if f.endswith(("cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "t21ps.h"))
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", EMU_SO, src])
    lib = C.CDLL(EMU_SO)
    lib.emu_run.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p,
                                            C.POINTER(_lib.BinSpecC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_double]
    lib.emu_run.restype = C.c_int
    lib.emu_bin_only.argtypes = [C.c_void_p, C.POINTER(_lib.BinSpecC), C.c_void_p, C.c_void_p]
    lib.emu_pitch.argtypes = [C.c_int]
    lib.emu_bin_segments.argtypes = [C.c_void_p, C.POINTER(_lib.BinSpecC), C.c_void_p, C.c_void_p]
    lib.emu_bin_segments.restype = C.c_int
    return lib


def run_emu(lib, cubes, upto=3, spec=None, offsets=(0.0, 0.0), nd=None, scale=1.0):
    x0 = np.ascontiguousarray(cubes[0])
    x1 = np.ascontiguousarray(cubes[1]) if len(cubes) > 1 else None
    dt = _lib.T21_F32 if x0.dtype == np.float32 else _lib.T21_F64
    nx, ny, nz = x0.shape
    pitch = lib.emu_pitch(nz)
    half = np.zeros((nx, ny, pitch), dtype=np.complex64 if dt == 0 else np.complex128)
    sums = counts = None
    cspec = None
    keep = None
    if spec is not None:
        cspec, keep = _lib.pack_binspec(spec, _lib.host_buffer)
        sums = np.zeros(spec.nbins)
        counts = np.zeros(spec.nbins, dtype=np.int64)
    rc = lib.emu_run(dt, nx, ny, nz, x0.ctypes.data, x1.ctypes.data if x1 is not None else None,
                     offsets[0], offsets[1], upto, half.ctypes.data,
                     C.byref(cspec) if cspec is not None else None,
                     sums.ctypes.data if sums is not None else None,
                     counts.ctypes.data if counts is not None else None,
                     nd.ctypes.data if nd is not None else None, int(nd is not None and nd.dtype == np.float64), scale)
    assert rc == 0
    return half[:, :, :nz // 2 + 1], sums, counts


@pytest.mark.parametrize("shape", [(16, 16, 16), (8, 16, 32), (32, 8, 64), (64, 128, 16), (16, 16, 256), (16, 16, 1024),
                                   (1024, 16, 32), (16, 1024, 32)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fft_passes_match_numpy(emu, shape, dtype):
    x = np.random.default_rng(3).standard_normal(shape).astype(dtype)
    tol = 2e-6 if dtype == np.float32 else 1e-13
    ref_z = np.fft.rfft(x.astype(np.float64), axis=2)
    got, _, _ = run_emu(emu, [x], upto=1)
    assert np.abs(got - ref_z).max() <= tol * np.abs(ref_z).max() * np.log2(shape[2])
    ref_zy = np.fft.fft(ref_z, axis=1)
    got, _, _ = run_emu(emu, [x], upto=2)
    assert np.abs(got - ref_zy).max() <= tol * np.abs(ref_zy).max() * np.log2(shape[1] * shape[2])


FUSED = [c for c in CASES if c["fn"] in ("power_spectrum_1d", "cross_power_spectrum_1d", "power_spectrum_mu",
                                         "cross_power_spectrum_mu")
         and all(n in POW2 for n in c["shape"]) and max(c["shape"]) <= 128]


@pytest.mark.parametrize("case", FUSED, ids=[c["id"] for c in FUSED])
def test_emulated_kernels_vs_golden(emu, case, golden):
    cubes = build_inputs(case)
    spec, box = spec_for(case, cubes[0].shape)
    want = golden_outputs(golden, case["id"])
    single = cubes[0].dtype == np.float32
    offs = [float(np.float32(c.mean())) if single else 0.0 for c in cubes] + [0.0]
    _, sums, counts = run_emu(emu, cubes, spec=spec, offsets=offs[:2])
    with np.errstate(all="ignore"):
        ps = (sums * bs.volume_scale(cubes[0].shape, box) / counts).reshape(want[0].shape)
    assert np.array_equal(counts.reshape(want[-1].shape).astype(float), want[-1]), "n_modes differ"
    if single:
        atol = 1e-5 * np.nanmax(np.abs(want[0])) if len(cubes) == 2 else 0.0
        np.testing.assert_allclose(ps, want[0], rtol=1e-5, atol=atol, equal_nan=True)
    else:
        np.testing.assert_allclose(ps, want[0], rtol=1e-10, equal_nan=True)


ND = [c for c in CASES if c["fn"] in ("power_spectrum_nd", "cross_power_spectrum_nd")
      and all(n in POW2 for n in c["shape"])]


@pytest.mark.parametrize("case", ND, ids=[c["id"] for c in ND])
def test_emulated_nd_vs_golden(emu, case, golden):
    cubes = build_inputs(case)
    kw = build_kwargs(case)
    box = bs.get_dims(kw.get("box_dims"), cubes[0].shape)
    want = golden_outputs(golden, case["id"])[0]
    single = cubes[0].dtype == np.float32
    offs = [float(np.float32(c.mean())) if single else 0.0 for c in cubes] + [0.0]
    out = np.full(cubes[0].shape, np.nan, dtype=np.float64)
    run_emu(emu, cubes, offsets=offs[:2], nd=out, scale=bs.volume_scale(cubes[0].shape, box))
    assert not np.isnan(out).any(), "some samples of the n-D output were never written"
    tol = 2e-5 if single else 1e-11
    assert np.abs(out - want).max() <= tol * np.abs(want).max()


BINONLY = [c for c in CASES if c["fn"] in ("radial_average", "mu_binning")]


@pytest.mark.parametrize("case", BINONLY, ids=[c["id"] for c in BINONLY])
def test_emulated_bin_only_vs_golden(emu, case, golden):
    arr = build_inputs(case)[0]
    kw = build_kwargs(case)
    want = golden_outputs(golden, case["id"])
    if case["fn"] == "radial_average":
        spec = bs.build_binspec(arr.shape, kw["box_dims"], kw["kbins"], binning=kw.get("binning", "log"),
                                order="shifted")
    else:
        spec = bs.build_binspec(arr.shape, kw["box_dims"], kw["kbins"], mubins=kw["mubins"],
                                los_axis=kw["los_axis"], absolute_mus=kw.get("absolute_mus", True),
                                last_bin_closed=False, order="shifted")
        arr = arr.copy()
        arr[tuple((np.array(arr.shape) / 2).astype(int))] = 0.0
    cspec, keep = _lib.pack_binspec(spec, _lib.host_buffer)
    sums = np.zeros(spec.nbins)
    counts = np.zeros(spec.nbins, dtype=np.int64)
    a64 = np.ascontiguousarray(arr, dtype=np.float64)
    emu.emu_bin_only(a64.ctypes.data, C.byref(cspec), sums.ctypes.data, counts.ctypes.data)
    with np.errstate(all="ignore"):
        ps = (sums / counts).reshape(want[0].shape)
    assert np.array_equal(counts.reshape(want[-1].shape).astype(float), want[-1])
    np.testing.assert_allclose(ps, want[0], rtol=1e-12, equal_nan=True)


SEG = [c for c in CASES if c["fn"] in ("power_spectrum_1d", "cross_power_spectrum_1d", "power_spectrum_mu", "cross_power_spectrum_mu")
       and len(c["shape"]) == 3]


@pytest.mark.parametrize("case", SEG, ids=[c["id"] for c in SEG])
def test_pencil_segment_tables_vs_golden(emu, case, golden):
    """The table-driven epilogue of the warp-per-pencil X pass (t21_xpass32.cu) as plain loops: build_pencil_segments()
    turns every x pencil into index ranges (one table for shells and |mu| bins, four for signed mu), the power of the
    numpy half spectrum is added range by range, the mode count of a range is the kernel's closed form -- and the
    result must be the reference's P and n_modes for every golden case of the fused estimators."""
    cubes = [np.asarray(c, dtype=np.float64) for c in build_inputs(case)]
    spec, box = spec_for(case, cubes[0].shape)
    want = golden_outputs(golden, case["id"])
    f0 = np.fft.rfftn(cubes[0])
    f1 = np.fft.rfftn(cubes[1]) if len(cubes) == 2 else f0
    phalf = np.ascontiguousarray((f0 * np.conj(f1)).real)
    cspec, keep = _lib.pack_binspec(spec, _lib.host_buffer)
    sums = np.zeros(spec.nbins)
    counts = np.zeros(spec.nbins, dtype=np.int64)
    rc = emu.emu_bin_segments(phalf.ctypes.data, C.byref(cspec), sums.ctypes.data, counts.ctypes.data)
    assert rc == 0, "more than 32 ranges on a pencil" if rc == 1 else "closed-form range count is wrong"
    with np.errstate(all="ignore"):
        ps = (sums * bs.volume_scale(cubes[0].shape, box) / counts).reshape(want[0].shape)
    assert np.array_equal(counts.reshape(want[-1].shape).astype(float), want[-1]), "n_modes differ"
    atol = 1e-9 * np.nanmax(np.abs(want[0])) if len(cubes) == 2 else 0.0
    np.testing.assert_allclose(ps, want[0], rtol=2e-7, atol=atol, equal_nan=True)
