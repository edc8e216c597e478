"""GPU parity of the surface-node projection ("Advect Nodes", set3d.f90:465-501) through the C ABI:
positions, phiSurf and gradPhiSurf bit-identical to the oracle's restatement of the reference loop."""
import numpy as np
import pytest

from conftest import GOLDEN, load_mesh

pytestmark = pytest.mark.gpu
DX = 0.05


@pytest.fixture(scope="module")
def S(lsf):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from levelsetfortran_b200 import set_subs
    return set_subs


def test_advect_nodes_host_buffers_cube40_bit_exact(S, oracle):
    z = np.load(f"{GOLDEN}/cube40_fields.npz")
    phi = np.asfortranarray(z["minmax"])
    sb = np.asfortranarray(z["phiSB"].astype(np.int32))
    X, E = load_mesh("cube40")
    g = oracle.grid_from_surface(X, DX)
    st, XX, ps, gs, mv = oracle.advect_nodes(phi, sb, g["xLo"], DX, X, 1000)
    assert st == 0
    XX2, ps2, gs2, mv2 = S.advectNodes(phi, sb, g["nx"], g["ny"], g["nz"], g["xLo"], DX, X, 1000)
    assert mv2 == mv > 0
    assert np.array_equal(XX2, XX) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)
    # iter = 0 is setPhiSurf alone (subs.f90:1057)
    st, XX, ps, gs, mv = oracle.advect_nodes(phi, sb, g["xLo"], DX, X, 0)
    XX2, ps2, gs2, mv2 = S.advectNodes(phi, sb, g["nx"], g["ny"], g["nz"], g["xLo"], DX, X, 0)
    assert mv2 == 0 and np.array_equal(XX2, X) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)


@pytest.mark.parametrize("tol", [1.0e-7, 0.0], ids=["converged", "iteration-limit"])
def test_advect_nodes_device_pipeline(S, oracle, tol):
    """Device-resident: min/max flow, then the projection on the resident phi.  The stencil band is the one the
    reference holds at that point -- the band of the field the last narrowBand call saw (set3d.f90:460)."""
    rng = np.random.default_rng(5)
    n = 56
    x = (np.arange(n) - n / 2.0 + 0.3) * DX
    Xg, Yg, Zg = np.meshgrid(x, x, x, indexing="ij")
    phi0 = np.asfortranarray(np.sqrt(Xg ** 2 + Yg ** 2 + Zg ** 2) - 0.71 + 0.003 * rng.standard_normal((n, n, n)))
    xLo = np.array([x[0], x[0], x[0]])
    d = rng.standard_normal((2000, 3))
    P = np.asfortranarray(d / np.linalg.norm(d, axis=1)[:, None] * (0.71 + 0.08 * (rng.random((2000, 1)) - 0.3)))
    a = phi0.copy(order="F")
    st, ne, hist, nb, sb = oracle.minmax(a, 40, DX, 0.01 * DX, tol=tol)
    sto, XX, ps, gs, mv = oracle.advect_nodes(a, sb, xLo, DX, P, 1000)
    assert sto == 0 and mv > 0
    g = S.DeviceGrid(n - 1, n - 1, n - 1)
    g.upload(phi0)
    rc, ne2, hist2 = g.minMaxFlow(40, DX, 0.01 * DX, tol=tol)
    assert rc == 0 and ne2 == ne
    XX2, ps2, gs2, mv2 = g.advectNodes(xLo, DX, P, 1000)
    g.close()
    assert mv2 == mv
    assert np.array_equal(XX2, XX) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)


def test_advect_nodes_errors(S, oracle):
    from levelsetfortran_b200 import _lib
    n = 24
    x = (np.arange(n) - n / 2.0) * DX
    Xg, Yg, Zg = np.meshgrid(x, x, x, indexing="ij")
    phi = np.asfortranarray(np.sqrt(Xg ** 2 + Yg ** 2 + Zg ** 2) - 0.3)
    nb, sb = oracle.narrowband(phi, DX)
    xLo = np.array([x[0], x[0], x[0]])
    with pytest.raises(_lib.LsfError) as e:
        S.advectNodes(phi, sb, n - 1, n - 1, n - 1, xLo, DX, np.array([[5.0, 0.0, 0.0]]), 10)
    assert e.value.code == _lib.LSF_ERR_NODE_OFF_GRID
    phi2 = np.asfortranarray(Xg - x[2] - 0.01)
    nb2, sb2 = oracle.narrowband(phi2, DX)
    with pytest.raises(_lib.LsfError) as e:
        S.advectNodes(phi2, sb2, n - 1, n - 1, n - 1, xLo, DX, np.array([[x[2] + 0.02, 0.0, 0.0]]), 10)
    assert e.value.code == _lib.LSF_ERR_BAND_ON_BOUNDARY


def test_advect_nodes_on_an_f32_grid(S, oracle):
    """fp32 grid: the projection is evaluated in fp64 on the widened field -> equals the oracle on that field."""
    n = 40
    x = (np.arange(n) - n / 2.0 + 0.3) * DX
    Xg, Yg, Zg = np.meshgrid(x, x, x, indexing="ij")
    phi = np.asfortranarray((np.sqrt(Xg ** 2 + Yg ** 2 + Zg ** 2) - 0.5).astype(np.float32).astype(np.float64))
    nb, sb = oracle.narrowband(phi, DX)
    xLo = np.array([x[0], x[0], x[0]])
    rng = np.random.default_rng(2)
    d = rng.standard_normal((300, 3))
    P = np.asfortranarray(d / np.linalg.norm(d, axis=1)[:, None] * 0.52)
    st, XX, ps, gs, mv = oracle.advect_nodes(phi, sb, xLo, DX, P, 1000)
    g = S.DeviceGrid(n - 1, n - 1, n - 1, f32=True)
    g.upload(phi)
    XX2, ps2, gs2, mv2 = g.advectNodes(xLo, DX, P, 1000)
    g.close()
    assert st == 0 and mv2 == mv and np.array_equal(XX2, XX) and np.array_equal(ps2, ps)
