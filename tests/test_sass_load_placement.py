"""Guard on the SASS of the built sweep kernels (no GPU needed: cuobjdump reads the object file).

The per-step global loads of the sweep kernels (look-ahead cell of the own row, one halo cell) end in a shared-memory deposit
(STS).  Where ptxas issues such a load relative to that deposit decides whether its latency is covered: builds in which the two
were < 30 instructions apart ran 3-7x slower with bit-identical results (sessions 18, 20, 35), and nothing
in the functional tests notices.  This test fails if a change to the step body (or a different compiler) produces such a build.
Skipped when the object file or cuobjdump is not there (the library is then not built here either)."""
import os
import shutil
import sys

import pytest

from conftest import ROOT

OBJ = os.path.join(ROOT, "levelsetfortran_b200", "csrc", "lsf_march.o")
sys.path.insert(0, os.path.join(ROOT, "tools"))

FAMILIES = {
    "fp64 single GPU": "k_reinit_marchINS_9FastArithELb{fa}ELb{fb}ELb{fc}ELb0",
    "fp64 z-slab": "k_reinit_marchINS_9FastArithELb{fa}ELb{fb}ELb{fc}ELb1",
    "fp32 single GPU": "k_reinit_march_f32ILb{fa}ELb{fb}ELb{fc}ELb0",
    "fp32 z-slab": "k_reinit_march_f32ILb{fa}ELb{fb}ELb{fc}ELb1",
}
MIN_DISTANCE = 100      # instructions between a per-step load and its deposit; the shipped build has 209-390


@pytest.fixture(scope="module")
def sass_text():
    from sass_ldg_distance import dump_sass
    return dump_sass(OBJ)


@pytest.mark.skipif(not os.path.exists(OBJ) or shutil.which("cuobjdump") is None, reason="lsf_march.o or cuobjdump not available")
@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_per_step_loads_are_issued_well_before_their_deposits(family, sass_text):
    from sass_ldg_distance import kernel_sass, ldg_distances
    seen = 0
    for o in range(8):
        key = FAMILIES[family].format(fa=o & 1, fb=(o >> 1) & 1, fc=(o >> 2) & 1)
        ins = kernel_sass(OBJ, key, sass_text)
        assert ins, f"kernel {key} not found in {OBJ}"
        deposits = [(a, s, d, s2) for a, s, d, s2 in ldg_distances(ins) if s2.split()[0 if not s2.startswith("@") else 1].startswith("STS")]
        assert len(deposits) >= 4, f"{key}: expected the look-ahead and halo loads of both step bodies, found {len(deposits)}"
        for a, s, d, s2 in deposits:
            assert d >= MIN_DISTANCE, f"{key}: load at {a:#x} is only {d} instructions before its deposit ({s} -> {s2})"
        seen += len(deposits)
    assert seen >= 32
