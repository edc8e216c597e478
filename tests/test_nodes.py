"""Surface-node projection ("Advect Nodes", set3d.f90:465-501) without a GPU:
  * the oracle's literal loop (setPhiSurf over ALL nodes after every single move, as the reference is written)
    equals its per-node form -- the re-ordering the product relies on;
  * the product's per-node code (levelsetfortran_b200/csrc/lsf_nodes.cuh, compiled for the CPU by tests/emu) is
    bit-identical to the oracle on the reference's own input and on synthetic fields;
  * the out-of-bounds cases the reference does not guard are reported, not computed."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh
from test_march_emu import emu  # noqa: F401  (fixture: the CPU build of the product headers)

DX = 0.05
dp = C.POINTER(C.c_double)


def _twin(emu, phi, sbsrc, xLo, X, iters=1000):
    emu.emu_advect_nodes.restype = C.c_int
    emu.emu_advect_nodes.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_double, dp, C.c_int, dp, dp, C.c_int,
                                     C.POINTER(C.c_longlong)]
    nx, ny, nz = (s - 1 for s in phi.shape)
    XX = np.asfortranarray(X, dtype=np.float64).copy(order="F")
    n = len(XX)
    ps, gs, mv = np.zeros(n), np.zeros((n, 3), order="F"), C.c_longlong(0)
    xLo = np.ascontiguousarray(xLo, dtype=np.float64)
    rc = emu.emu_advect_nodes(phi.ctypes.data_as(dp), sbsrc.ctypes.data_as(dp), nx, ny, nz, xLo.ctypes.data_as(dp), DX,
                              XX.ctypes.data_as(dp), n, ps.ctypes.data_as(dp), gs.ctypes.data_as(dp), iters, C.byref(mv))
    return rc, XX, ps, gs, mv.value


def _cube40(oracle):
    z = np.load(f"{GOLDEN}/cube40_fields.npz")
    phi = np.asfortranarray(z["minmax"])
    sb = np.asfortranarray(z["phiSB"].astype(np.int32))
    X, E = load_mesh("cube40")
    g = oracle.grid_from_surface(X, DX)
    return phi, sb, X, g


def test_literal_all_nodes_loop_equals_per_node_loop(oracle):
    phi, sb, X, g = _cube40(oracle)
    sub = np.asfortranarray(X[::37])
    a = oracle.advect_nodes(phi, sb, g["xLo"], DX, sub, 1000, literal=True)
    b = oracle.advect_nodes(phi, sb, g["xLo"], DX, sub, 1000, literal=False)
    assert a[0] == b[0] == 0 and a[4] == b[4] > 0
    for u, v in zip(a[1:4], b[1:4]):
        assert np.array_equal(u, v)


def test_projection_on_the_reference_input(oracle, emu):  # noqa: F811
    """cube40 after the min/max flow (golden field): every node ends with phiSurf <= 1E-13, exactly the nodes
    that started outside the surface (phiSurf > 1E-13) have moved, nodes inside do not move (set3d.f90:493)."""
    phi, sb, X, g = _cube40(oracle)
    st, XX, ps, gs, mv = oracle.advect_nodes(phi, sb, g["xLo"], DX, X, 1000)
    assert st == 0 and mv > len(X)
    assert (ps <= 1e-13).all()
    st0, _, ps0, _, _ = oracle.advect_nodes(phi, sb, g["xLo"], DX, X, 0)
    moved = np.abs(XX - X).max(axis=1) > 0
    assert np.array_equal(moved, ps0 > 1e-13)
    # (a moved node may overshoot to phiSurf < 0 and then stays there: the reference's behaviour, kept)
    nrm = np.sqrt((gs * gs).sum(axis=1))
    assert np.all((np.abs(nrm - 1) < 1e-12) | (nrm == 0))
    # the product's per-node code, on the CPU: bit-identical
    sbsrc = np.asfortranarray(np.where(sb == 1, 0.0, 1.0e300))
    rc, XX2, ps2, gs2, mv2 = _twin(emu, phi, sbsrc, g["xLo"], X)
    assert rc == 0 and mv2 == mv
    assert np.array_equal(XX2, XX) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)


@pytest.mark.parametrize("seed", [0, 1])
def test_projection_synthetic_sphere(oracle, emu, seed):  # noqa: F811
    """Random points around a sphere's zero level set (band from abs(phi) itself, as lsf_grid_advect_nodes uses it)."""
    rng = np.random.default_rng(seed)
    n = 48
    x = (np.arange(n) - n / 2.0 + 0.3) * DX
    Xg, Yg, Zg = np.meshgrid(x, x, x, indexing="ij")
    phi = np.asfortranarray(np.sqrt(Xg ** 2 + Yg ** 2 + Zg ** 2) - 0.61 + 0.002 * rng.standard_normal((n, n, n)))
    nb, sb = oracle.narrowband(phi, DX)
    xLo = np.array([x[0], x[0], x[0]])
    d = rng.standard_normal((500, 3))
    P = np.asfortranarray(d / np.linalg.norm(d, axis=1)[:, None] * (0.61 + 0.1 * (rng.random((500, 1)) - 0.3)))
    st, XX, ps, gs, mv = oracle.advect_nodes(phi, sb, xLo, DX, P, 1000)
    assert st == 0 and mv > 0
    rc, XX2, ps2, gs2, mv2 = _twin(emu, phi, phi, xLo, P)
    assert rc == 0 and mv2 == mv
    assert np.array_equal(XX2, XX) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)
    # iter = 0: setPhiSurf only
    st, XX, ps, gs, mv = oracle.advect_nodes(phi, sb, xLo, DX, P, 0)
    rc, XX2, ps2, gs2, mv2 = _twin(emu, phi, phi, xLo, P, 0)
    assert mv == mv2 == 0 and np.array_equal(XX2, P) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)


def test_out_of_bounds_cases_are_reported(oracle, emu):  # noqa: F811
    n = 24
    x = (np.arange(n) - n / 2.0) * DX
    Xg, Yg, Zg = np.meshgrid(x, x, x, indexing="ij")
    phi = np.asfortranarray(np.sqrt(Xg ** 2 + Yg ** 2 + Zg ** 2) - 0.3)
    nb, sb = oracle.narrowband(phi, DX)
    xLo = np.array([x[0], x[0], x[0]])
    off = np.asfortranarray(np.array([[5.0, 0.0, 0.0]]))                 # outside the grid
    assert oracle.advect_nodes(phi, sb, xLo, DX, off, 10)[0] == -5
    assert _twin(emu, phi, phi, xLo, off, 10)[0] == 1                    # NODE_OFF_GRID
    # a band that reaches the boundary next to a node: the order-8 stencil would leave the array
    phi2 = np.asfortranarray(Xg - x[2] - 0.01)                           # zero level set 2 points from the i = 0 face
    nb2, sb2 = oracle.narrowband(phi2, DX)
    node = np.asfortranarray(np.array([[x[2] + 0.02, 0.0, 0.0]]))
    assert oracle.advect_nodes(phi2, sb2, xLo, DX, node, 10)[0] == -3
    assert _twin(emu, phi2, phi2, xLo, node, 10)[0] == 2                 # NODE_BAND_ON_BOUNDARY


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_projection_through_the_slab_view(oracle, emu, nranks):  # noqa: F811
    """z-slabs: every value is fetched through lsf::SlabView from the local array of the rank that OWNS its plane (ghost planes
    poisoned with NaN here).  Nodes near and across the slab boundaries; result bit-identical to the whole-grid code."""
    rng = np.random.default_rng(7)
    n = 48
    x = (np.arange(n) - n / 2.0 + 0.3) * DX
    Xg, Yg, Zg = np.meshgrid(x, x, x, indexing="ij")
    phi = np.asfortranarray(np.sqrt(Xg ** 2 + Yg ** 2 + Zg ** 2) - 0.61 + 0.002 * rng.standard_normal((n, n, n)))
    xLo = np.array([x[0], x[0], x[0]])
    d = rng.standard_normal((400, 3))
    P = np.asfortranarray(d / np.linalg.norm(d, axis=1)[:, None] * (0.61 + 0.1 * (rng.random((400, 1)) - 0.3)))
    rc, XX, ps, gs, mv = _twin(emu, phi, phi, xLo, P)
    assert rc == 0 and mv > 0
    emu.emu_advect_nodes_slabs.restype = C.c_int
    emu.emu_advect_nodes_slabs.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_double, dp, C.c_int, dp, dp, C.c_int,
                                           C.POINTER(C.c_longlong)]
    XX2 = P.copy(order="F")
    ps2, gs2, mv2 = np.zeros(len(P)), np.zeros((len(P), 3), order="F"), C.c_longlong(0)
    rc2 = emu.emu_advect_nodes_slabs(phi.ctypes.data_as(dp), phi.ctypes.data_as(dp), n - 1, n - 1, n - 1, nranks, xLo.ctypes.data_as(dp), DX,
                                     XX2.ctypes.data_as(dp), len(P), ps2.ctypes.data_as(dp), gs2.ctypes.data_as(dp), 1000, C.byref(mv2))
    assert rc2 == 0 and mv2.value == mv
    assert np.array_equal(XX2, XX) and np.array_equal(ps2, ps) and np.array_equal(gs2, gs)
