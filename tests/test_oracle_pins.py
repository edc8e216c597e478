"""Pins the CPU oracle (oracle/lsf_oracle.c).  The reference ships no tests or golden vectors and
cannot be compiled here (no Fortran compiler), so parity is UNPINNED against the gfortran binary;
the pins are (a) the known answers of the independent survey-time transcription,
(b) the committed golden fixtures (tests/golden/make_golden.py), (c) internal consistency: literal
BC loop == closed form, lexicographic sweep == hyperplane sweep."""
import numpy as np
import pytest

from conftest import GOLDEN, load_mesh

DX = 0.05


@pytest.fixture(scope="module")
def cube40(oracle):
    X, E = load_mesh("cube40")
    g = oracle.grid_from_surface(X, DX)
    return X, E, g, np.load(f"{GOLDEN}/cube40_fields.npz")


def test_cube40_sign_search_known_answers(oracle, cube40):
    X, E, g, gold = cube40
    phi = np.ones((62, 62, 62), order="F")
    oracle.sign_init(phi, g["xLo"], DX, X, E, g["box"])
    assert repr(float(phi.sum())) == "133852.64945568278"
    assert repr(float(phi.min())) == "-0.05252437440599863"
    assert int((phi < 0).sum()) == 59319 and int((phi == 0).sum()) == 9602
    assert np.array_equal(phi, gold["sign"])


def test_cube40_reinit_full_run_known_answers(oracle, cube40):
    """2155 sweeps, ~85 s: exit iteration, last RMS and field values of SURVEY.md section 6."""
    X, E, g, gold = cube40
    phi = np.asfortranarray(gold["sign"].copy())
    st, n, hist = oracle.reinit(phi, 10000, DX, 0.1 * g["dxx"])
    assert (st, n) == (0, 2154)
    assert abs(hist[-1] - 9.997035e-06) < 1e-12
    assert repr(float(phi.sum())) == "43431.239785368976"
    assert repr(float(phi.min())) == "-0.9628290001261777" and repr(float(phi.max())) == "1.027632914780677"
    assert repr(float(phi[31, 31, 31])) == "-0.915413945378742" and repr(float(phi[0, 0, 0])) == "0.8395727865558302"
    assert np.array_equal(phi, gold["reinit1"]) and np.array_equal(hist, gold["rms_reinit1"])


def test_cube40_minmax_and_second_reinit_known_answers(oracle, cube40):
    X, E, g, gold = cube40
    phi = np.asfortranarray(gold["reinit1"].copy())
    st, n, hist, nb, sb = oracle.minmax(phi, 10000, DX, 0.01 * g["dxx"])
    assert (st, n) == (0, 406) and abs(hist[-1] - 9.8381e-08) < 1e-11
    assert repr(float(phi.sum())) == "44075.247828405096"
    assert int((phi < 0).sum()) == 59311 and int((phi == 0).sum()) == 0
    assert np.array_equal(phi, gold["minmax"])
    st, n, hist = oracle.reinit(phi, 2000, DX, 0.001 * g["dxx"])
    assert (st, n) == (0, 0) and abs(hist[0] - 1.62e-06) < 1e-8


def test_twocube10_nan_stop_at_272(oracle):
    X, E = load_mesh("twoCube10")
    g = oracle.grid_from_surface(X, DX)
    gold = np.load(f"{GOLDEN}/twoCube10_fields.npz")
    phi = np.ones((262, 42, 42), order="F")
    oracle.sign_init(phi, g["xLo"], DX, X, E, g["box"])
    assert np.array_equal(phi, gold["sign"])
    st, n, hist = oracle.reinit(phi, 10000, DX, 0.1 * g["dxx"])
    assert (st, n) == (1, 272) and np.isnan(hist[272]) and not np.isnan(hist[:272]).any()
    assert np.array_equal(hist[:272], gold["rms_reinit1"][:272])


@pytest.mark.parametrize("shape", [(6, 6, 6), (7, 6, 5), (12, 13, 14), (16, 12, 11), (3, 3, 3)])
def test_bc_closed_form_equals_literal_loop(oracle, shape):
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.standard_normal(shape))
    b = a.copy(order="F")
    oracle.bc(a, DX, literal=True)
    oracle.bc(b, DX, literal=False)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(9, 9, 9), (14, 11, 12), (20, 13, 17)])
def test_hyperplane_order_is_bitwise_identical(oracle, shape):
    rng = np.random.default_rng(4)
    p = np.asfortranarray(rng.standard_normal(shape) * 0.3)
    pS, q = p.copy(order="F"), p.copy(order="F")
    for r in range(1, 9):
        oracle.reinit_sweep(p, pS, DX, 0.001, r, hyperplane=False)
        oracle.reinit_sweep(q, pS, DX, 0.001, r, hyperplane=True)
        assert np.array_equal(p, q)


def test_reinit_literal_bc_full_loop_small(oracle):
    rng = np.random.default_rng(5)
    p = np.asfortranarray(rng.standard_normal((11, 10, 12)) * 0.2)
    q = p.copy(order="F")
    r1 = oracle.reinit(p, 20, DX, 0.002, bc=0, order=0)
    r2 = oracle.reinit(q, 20, DX, 0.002, bc=1, order=1)
    assert r1[:2] == r2[:2] and np.array_equal(r1[2], r2[2]) and np.array_equal(p, q)
