"""Input side on the GPU: C2Ray-format files -> device, fused calc_dt, subtract_mean_signal, and the
staged host -> device copy behind numpy inputs, against the oracle's restatement of the reference readers / formulas."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def write_xfrac(path, xi, old_format=False, binary_format=False):
    """the layout XfracFile.read_from_file expects (xfrac_file.py:48-62)"""
    mesh = np.array(xi.shape, dtype='int32')
    with open(path, 'wb') as f:
        if binary_format:
            mesh.tofile(f)
        else:
            np.array([12, *xi.shape, 12, xi.size * (4 if old_format else 8)], dtype='int32').tofile(f)
        xi.astype('float32' if old_format else 'float64').ravel(order='F').tofile(f)


def write_density(path, raw):
    with open(path, 'wb') as f:
        np.array(raw.shape, dtype='int32').tofile(f)
        raw.astype('float32').ravel(order='F').tofile(f)


@pytest.fixture(scope="module")
def t2c():
    import torch
    assert torch.cuda.is_available()
    import tools21cm_b200
    return tools21cm_b200


@pytest.mark.parametrize("shape", [(64, 64, 64), (40, 24, 56)])
@pytest.mark.parametrize("old_format,binary_format,neutral", [(False, False, False), (True, True, True)])
def test_readers_and_fused_calc_dt(t2c, tmp_path, shape, old_format, binary_format, neutral):
    from oracle import ps_oracle
    rng = np.random.default_rng(3)
    xi = rng.random(shape)
    raw = (1.0 + 0.3 * rng.standard_normal(shape)).clip(0.05).astype(np.float32) * 50.0
    fx, fd = str(tmp_path / "xfrac3d_8.515.bin"), str(tmp_path / "8.515n_all.dat")
    write_xfrac(fx, xi, old_format, binary_format)
    write_density(fd, raw)
    xf = t2c.XfracFile(fx, old_format=old_format, neutral=neutral, binary_format=binary_format)
    df = t2c.DensityFile(fd)
    xi_o = ps_oracle.read_xfrac(fx, old_format, neutral, binary_format)
    raw_o, cgs_o = ps_oracle.read_density(fd)
    assert (xf.mesh_x, xf.mesh_y, xf.mesh_z) == shape and abs(xf.z - 8.515) < 1e-12 and abs(df.z - 8.515) < 1e-12
    assert np.array_equal(xf.xi.cpu().numpy(), xi_o)                       # bit for bit: a copy and a transpose
    assert np.array_equal(df.raw_density.cpu().numpy(), raw_o)
    np.testing.assert_allclose(df.cgs_density.cpu().numpy(), cgs_o, rtol=1e-6)
    want = ps_oracle.calc_dt(xi_o, cgs_o, 8.515)
    import torch
    got64 = t2c.calc_dt(xf, df, out_dtype=torch.float64).cpu().numpy()     # fused kernel, z from the file name
    np.testing.assert_allclose(got64, want, rtol=2e-7, atol=1e-12)          # (the reference rounds cgs density to float32)
    got32 = t2c.calc_dt(xf, df).cpu().numpy()
    assert got32.dtype == np.float32
    np.testing.assert_allclose(got32, want, rtol=1e-6, atol=1e-9)
    arr = t2c.calc_dt(xi_o, cgs_o, z=8.515, out_dtype=torch.float64).cpu().numpy()   # plain arrays: the elementwise route
    np.testing.assert_allclose(arr, want, rtol=1e-13)
    # file -> device -> power spectrum without touching the host again
    dT = t2c.subtract_mean_signal(t2c.calc_dt(xf, df), 2)
    dT_o = ps_oracle.subtract_mean_signal(want, 2)
    np.testing.assert_allclose(dT.cpu().numpy(), dT_o, rtol=0, atol=2e-5 * np.abs(want).max())
    ps, ks, nm = t2c.power_spectrum_1d(dT, kbins=8, box_dims=244 / 0.7, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(dT_o, kbins=8, box_dims=244 / 0.7, return_n_modes=True)
    assert np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o)
    np.testing.assert_allclose(ps, ps_o, rtol=2e-5, equal_nan=True)


def test_truncated_file_fails_like_the_reference(t2c, tmp_path):
    p = str(tmp_path / "xfrac3d_7.000.bin")
    write_xfrac(p, np.zeros((8, 8, 8)))
    with open(p, "r+b") as f:
        f.truncate(24 + 8 * 500)
    with pytest.raises(ValueError):
        t2c.XfracFile(p)


def test_staged_host_to_device_copy_is_exact(t2c):
    """numpy inputs above 8 MB go through t21_h2d_staged (pinned ring, several host threads, chunk tails)."""
    import torch
    from tools21cm_b200 import power_spectrum as P
    rng = np.random.default_rng(0)
    for n in ((1 << 21) + 12345, 3 * (32 << 20) // 4 + 7, 1 << 25):
        a = rng.random(n, dtype=np.float32)
        t = P._h2d(a)
        torch.cuda.synchronize()
        assert torch.equal(t.cpu(), torch.from_numpy(a))
    cube = rng.standard_normal((160, 160, 160)).astype(np.float32)            # 16 MB: the estimator's own staging
    from oracle import ps_oracle
    ps, ks, nm = t2c.power_spectrum_1d(cube, kbins=8, box_dims=100.0, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(cube, kbins=8, box_dims=100.0, return_n_modes=True)
    assert np.array_equal(nm, nm_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-5, equal_nan=True)


@pytest.mark.gpu
def test_host_fed_estimator_equals_the_resident_one(t2c, monkeypatch):
    """A numpy cube (or a pinned CPU tensor) handed to power_spectrum_1d / power_spectrum_mu is copied in groups of x planes
    while the Z and Y passes run on the groups that have landed (t21_ps_binned_host).  Same mode counts, and powers equal
    to the device-resident call up to the rounding of a different mean offset (it is estimated on the first group)."""
    import torch
    from tools21cm_b200 import synthetic
    monkeypatch.setenv("T21_HOST_FEED_GROUP_MB", "8")          # (read once per process: eight groups for this cube)
    a = synthetic.toy_dt_cube((256, 256, 256), 3, np.dtype("float32"))
    dev = torch.from_numpy(a).cuda()
    ref = t2c.power_spectrum_1d(dev, kbins=12, box_dims=300.0, return_n_modes=True)
    for host in (a, torch.from_numpy(a).pin_memory()):
        got = t2c.power_spectrum_1d(host, kbins=12, box_dims=300.0, return_n_modes=True)
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[1], ref[1])
        np.testing.assert_allclose(got[0], ref[0], rtol=2e-6)
    refm = t2c.power_spectrum_mu(dev, los_axis=1, mubins=4, kbins=6, box_dims=300.0, return_n_modes=True, absolute_mus=False)
    gotm = t2c.power_spectrum_mu(a, los_axis=1, mubins=4, kbins=6, box_dims=300.0, return_n_modes=True, absolute_mus=False)
    assert np.array_equal(gotm[3], refm[3])
    np.testing.assert_allclose(gotm[0], refm[0], rtol=2e-6, equal_nan=True)
