"""world_size-2 `gloo` runs of the multi-GPU host logic on the CPU: item partitioning and result
gathering for batches, and for slab-sharded cubes the common mean offset, the pack + all-to-all
re-partition, the global-row offset of stage 2 and the final all-reduce.  The two CUDA stages are
replaced by a numpy model (tests/kernel_model.py); the parity bar against the oracle is the same."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyBackend:
    """Stand-in for libt21ps.so's slab stages (same contracts as distributed.CudaBackend)."""

    def mean_estimate(self, x):
        return x.double().mean()

    def stage1(self, x, offset):
        a = x.numpy().astype(np.float64) - (float(offset.item()) if offset is not None else 0.0)
        f = np.fft.fft(np.fft.rfft(a, axis=2), axis=1)
        nzh = a.shape[2] // 2 + 1
        pitch = (nzh + 15) // 16 * 16
        w = np.zeros(a.shape[:2] + (pitch,), dtype=np.complex128)
        w[:, :, :nzh] = f
        return torch.from_numpy(w)

    def stage2(self, cols, iy0, ny_global, nz, offsets, spec_key, spec_builder):
        from kernel_model import bin_half_spectrum
        spec = spec_builder()
        nzh = nz // 2 + 1
        ntot = cols[0].nx * ny_global * nz
        fs = []
        for c, off in zip(cols, offsets):
            f = np.fft.fft(c.data.numpy()[:, :, :nzh], axis=0)
            if iy0 == 0 and off is not None:
                f[0, 0, 0] += float(off.item()) * ntot
            fs.append(f)
        power = (fs[0].real ** 2 + fs[0].imag ** 2) if len(fs) == 1 else \
            (fs[0].real * fs[1].real + fs[0].imag * fs[1].imag)
        sums, counts = bin_half_spectrum(power, spec, check_lut=False, iy0=iy0)
        return spec, torch.from_numpy(sums), torch.from_numpy(counts)


def _stage2_nd(self, cols, iy0, ny_global, nz, offsets, scale, out_dtype):
    nzh = nz // 2 + 1
    ntot = cols[0].nx * ny_global * nz
    fs = []
    for c, off in zip(cols, offsets):
        f = np.fft.fft(c.data.numpy()[:, :, :nzh], axis=0)
        if iy0 == 0 and off is not None:
            f[0, 0, 0] += float(off.item()) * ntot
        fs.append(f)
    power = (fs[0].real ** 2 + fs[0].imag ** 2) if len(fs) == 1 else \
        (fs[0].real * fs[1].real + fs[0].imag * fs[1].imag)
    return torch.from_numpy(power * scale).to(out_dtype)


NumpyBackend.stage2_nd = _stage2_nd


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tools21cm_b200 import distributed as D
        from oracle import ps_oracle
        out = {}
        # ---- batches: item i -> rank i mod P, gathered in order ----
        items = list(range(7))
        out["batched"] = D.batched(lambda i: (i, i * i, rank), items)
        # ---- slab-sharded cube ----
        shape = (16, 8, 12)
        cube = np.random.default_rng(5).standard_normal(shape).astype(np.float32) + 3.0
        nxl = shape[0] // world
        slab = torch.from_numpy(cube[rank * nxl:(rank + 1) * nxl].copy())
        ps, ks, nm = D.sharded_power_spectrum_1d(slab, kbins=5, box_dims=[50.0, 40.0, 60.0], return_n_modes=True,
                                                 backend=NumpyBackend())
        ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(cube, kbins=5, box_dims=[50.0, 40.0, 60.0], return_n_modes=True)
        out["1d"] = (np.array_equal(nm, nm_o), np.array_equal(ks, ks_o), float(np.nanmax(np.abs(ps / ps_o - 1))))
        other = np.random.default_rng(6).standard_normal(shape)
        a = torch.from_numpy(cube.astype(np.float64)[rank * nxl:(rank + 1) * nxl].copy())
        b = torch.from_numpy(other[rank * nxl:(rank + 1) * nxl].copy())
        xs = D.sharded_cross_power_spectrum_1d(a, b, kbins=4, box_dims=30.0, return_n_modes=True,
                                               backend=NumpyBackend())
        xs_o = ps_oracle.cross_power_spectrum_1d(cube.astype(np.float64), other, kbins=4, box_dims=30.0,
                                                 return_n_modes=True)
        out["x1d"] = (np.array_equal(xs[2], xs_o[2]), float(np.nanmax(np.abs(xs[0] - xs_o[0])) / np.nanmax(np.abs(xs_o[0]))))
        mu = D.sharded_power_spectrum_mu(a, los_axis=1, mubins=4, kbins=3, box_dims=30.0, return_n_modes=True,
                                         absolute_mus=False, backend=NumpyBackend())
        mu_o = ps_oracle.power_spectrum_mu(cube.astype(np.float64), los_axis=1, mubins=4, kbins=3, box_dims=30.0,
                                           return_n_modes=True, absolute_mus=False)
        out["mu"] = (np.array_equal(mu[3], mu_o[3]), bool(np.allclose(mu[0], mu_o[0], rtol=1e-9, equal_nan=True)))
        # ---- n-D output, left sharded by rows of y ----
        rows, yidx = D.sharded_cross_power_spectrum_nd(a, b, 30.0, backend=NumpyBackend())
        full = ps_oracle.cross_power_spectrum_nd(cube.astype(np.float64), other, 30.0)
        out["nd"] = float(np.abs(rows.numpy() - full[:, yidx.numpy(), :]).max() / np.abs(full).max())
        odd = np.random.default_rng(9).standard_normal((8, 12, 7))         # odd z: no Nyquist plane
        nx8 = odd.shape[0] // world
        rows, yidx = D.sharded_power_spectrum_nd(torch.from_numpy(odd[rank * nx8:(rank + 1) * nx8].copy()),
                                                 [10.0, 20.0, 30.0], backend=NumpyBackend())
        full = ps_oracle.power_spectrum_nd(odd, [10.0, 20.0, 30.0])
        out["nd_odd"] = float(np.abs(rows.numpy() - full[:, yidx.numpy(), :]).max() / np.abs(full).max())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_partition_is_round_robin():
    from tools21cm_b200.distributed import partition
    assert partition(7, 2, 0) == [0, 2, 4, 6] and partition(7, 2, 1) == [1, 3, 5]
    assert sorted(sum((partition(64, 8, r) for r in range(8)), [])) == list(range(64))


@pytest.mark.parametrize("world", [2, 4])
def test_world_size_n_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        out = results[rank]
        assert [r[0] for r in out["batched"]] == list(range(7))
        assert [r[2] for r in out["batched"]] == [i % world for i in range(7)]      # who computed what
        counts_ok, centres_ok, err = out["1d"]
        assert counts_ok and centres_ok and err < 1e-5
        assert out["x1d"][0] and out["x1d"][1] < 1e-10
        assert out["mu"] == (True, True)
        assert out["nd"] < 1e-12 and out["nd_odd"] < 1e-12
