"""GPU checks of the slab-sharded path.

On one GPU the P ranks are emulated one after another in a single process (no NCCL, no kernels that
wait on each other): stage 1 per x slab, the re-partition done by tensor indexing, stage 2 per y slab,
results summed.  That exercises the CUDA stages and the global-row offset exactly as a P-rank run does.
With >= 2 GPUs the real thing runs under NCCL (one process per GPU)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def emulate_ranks(cubes, world, kw_spec, mu=False, fused=False):
    """fused=False: the pack fused into the Y pass + the all-to-all done by indexing the send blocks (the y slab
    is then P consecutive x slabs).  fused=True: the Y pass stores into the "peer" buffers directly (all on
    this GPU here), one chunk-major array per rank."""
    import ctypes as C
    import torch
    from tools21cm_b200 import binspec as bs, distributed as D
    from tools21cm_b200.power_spectrum import _freeze
    be = D.CudaBackend()
    shape = tuple(cubes[0].shape)
    nx, ny, nz = shape
    nxl, nyl = nx // world, ny // world
    pitch = be.pitch(nz)
    single = cubes[0].dtype == torch.float32
    cdt = torch.complex64 if single else torch.complex128
    offs = [c.reshape(-1)[::7].double().mean().reshape(1).to(c.dtype) if single else None for c in cubes]
    cols_all = []
    for c, off in zip(cubes, offs):
        slabs = [c[r * nxl:(r + 1) * nxl].contiguous() for r in range(world)]
        if fused:
            bufs = [torch.zeros(nx * nyl * pitch, dtype=cdt, device=c.device) for _ in range(world)]
            ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
            for r in range(world):
                be.stage1_p2p(slabs[r], off, ptrs, world, r * nxl)
            cols_all.append([D.YSlab(bufs[s], nx, nyl, pitch) for s in range(world)])
        else:
            send = [be.stage1(slabs[r], off, send_world=world) for r in range(world)]       # [P][block] each
            cols_all.append([D.YSlab(torch.stack([send[r][s] for r in range(world)]), nx, nyl, pitch, n_xslabs=world)
                             for s in range(world)])
    sums = counts = None
    box = kw_spec.pop("box")
    for s in range(world):
        spec, sm, ct = be.stage2([ca[s] for ca in cols_all], s * nyl, ny, nz, offs, ("emu", _freeze(box), mu, world),
                                 lambda: bs.build_binspec(shape, box, **kw_spec))
        sums = sm.clone() if sums is None else sums + sm
        counts = ct.clone() if counts is None else counts + ct
    with np.errstate(all="ignore"):
        ps = sums.cpu().numpy() * bs.volume_scale(shape, box) / counts.cpu().numpy()
    return spec, ps, counts.cpu().numpy().astype(float)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_stages_emulated_on_one_gpu(world, fused):
    import torch
    from oracle import ps_oracle
    from tools21cm_b200 import synthetic
    cube = synthetic.toy_dt_cube((64, 64, 32), seed=3)
    box = [200.0, 200.0, 100.0]
    spec, ps, nm = emulate_ranks([torch.from_numpy(cube).cuda()], world, dict(box=box, kbins=10), fused=fused)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(cube, kbins=10, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o) and np.array_equal(spec.k_centres(), ks_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-5, equal_nan=True)


def test_slab_stages_cross_and_mu_fp64():
    import torch
    from oracle import ps_oracle
    rng = np.random.default_rng(8)
    a, b = rng.standard_normal((32, 32, 32)), rng.standard_normal((32, 32, 32))
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    spec, ps, nm = emulate_ranks([ta, tb], 4, dict(box=[90.0] * 3, kbins=8))
    ps_o, _, nm_o = ps_oracle.cross_power_spectrum_1d(a, b, kbins=8, box_dims=90.0, return_n_modes=True)
    assert np.array_equal(nm, nm_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-9, atol=1e-12 * np.nanmax(np.abs(ps_o)), equal_nan=True)
    spec, ps, nm = emulate_ranks([ta], 2, dict(box=[90.0] * 3, kbins=5, mubins=4, los_axis=1, absolute_mus=False,
                                               last_bin_closed=False), mu=True, fused=True)
    pm_o, _, _, nmm_o = ps_oracle.power_spectrum_mu(a, los_axis=1, mubins=4, kbins=5, box_dims=90.0,
                                                    return_n_modes=True, absolute_mus=False)
    assert np.array_equal(nm.reshape(nmm_o.shape), nmm_o)
    np.testing.assert_allclose(ps.reshape(pm_o.shape), pm_o, rtol=1e-10, equal_nan=True)


@pytest.mark.parametrize("fused", [False, True])
def test_slab_stages_long_x_cross_fp32(fused):
    """nx = 2048, two fields, fp32: the half-width cross tiles of the X pass reading a y slab that is either one
    chunk-major array (fused exchange) or P consecutive x slabs (NCCL exchange)."""
    import torch
    from oracle import ps_oracle
    rng = np.random.default_rng(12)
    a = rng.standard_normal((2048, 16, 32)).astype(np.float32)
    b = (0.3 * a + rng.standard_normal(a.shape)).astype(np.float32)
    box = [400.0, 30.0, 60.0]
    spec, ps, nm = emulate_ranks([torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()], 4, dict(box=box, kbins=8),
                                 fused=fused)
    ps_o, _, nm_o = ps_oracle.cross_power_spectrum_1d(a, b, kbins=8, box_dims=box, return_n_modes=True)
    auto, _ = ps_oracle.power_spectrum_1d(a, kbins=8, box_dims=box)
    assert np.array_equal(nm, nm_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-5, atol=1e-5 * np.nanmax(np.abs(auto)), equal_nan=True)


def test_sharded_nd_single_rank_equals_nd():
    """world = 1: the slab stages, the half-spectrum store and the mirror-row assembly on one GPU."""
    import torch
    import tools21cm_b200 as t2c
    from tools21cm_b200 import distributed as D, synthetic
    x = synthetic.toy_dt_cube_torch((64, 32, 128), seed=4)
    y = synthetic.gaussian_cube_torch((64, 32, 128), seed=5)
    rows, yidx = D.sharded_power_spectrum_nd(x, [100.0, 50.0, 200.0])
    full = t2c.power_spectrum_nd(x, box_dims=[100.0, 50.0, 200.0])
    assert torch.equal(yidx, (torch.arange(32) + 16) % 32)
    assert (rows - full[:, yidx.to(full.device), :]).abs().max() <= 2e-6 * full.max()
    rows, yidx = D.sharded_cross_power_spectrum_nd(x.double(), y.double(), [100.0, 50.0, 200.0])
    full = t2c.cross_power_spectrum_nd(x.double(), y.double(), [100.0, 50.0, 200.0])
    assert (rows - full[:, yidx.to(full.device), :]).abs().max() <= 1e-12 * full.abs().max()


def _nccl_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import ps_oracle
        from tools21cm_b200 import distributed as D, synthetic
        cube = synthetic.toy_dt_cube((128, 64, 64), seed=21)
        nxl = cube.shape[0] // world
        slab = torch.from_numpy(cube[rank * nxl:(rank + 1) * nxl].copy()).cuda()
        ps, ks, nm = D.sharded_power_spectrum_1d(slab, kbins=12, box_dims=[300.0, 150.0, 150.0], return_n_modes=True)
        ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(cube, kbins=12, box_dims=[300.0, 150.0, 150.0],
                                                       return_n_modes=True)
        ok = bool(np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o) and
                  np.allclose(ps, ps_o, rtol=1e-5, equal_nan=True))
        # n-D cross spectrum left sharded by rows of y (BASELINE.json config 5, small)
        other = synthetic.gaussian_cube(cube.shape, seed=22)
        oslab = torch.from_numpy(other[rank * nxl:(rank + 1) * nxl].copy()).cuda()
        rows, yidx = D.sharded_cross_power_spectrum_nd(slab, oslab, [300.0, 150.0, 150.0])
        full = ps_oracle.cross_power_spectrum_nd(cube, other, [300.0, 150.0, 150.0])
        err = np.abs(rows.cpu().numpy() - full[:, yidx.numpy(), :]).max() / np.abs(full).max()
        ok = ok and bool(err < 2e-5)      # fp32; the largest sample is the DC mode
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_over_nccl():
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if world < 4 else 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(res[r] for r in range(world))
