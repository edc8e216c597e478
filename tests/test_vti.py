"""The .vti writer of the Python mirror against the reference's WRITE statements (set3d.f90:316-351): a known-answer
header typed out by hand from those statements, the raw block layout, and a round trip."""
import numpy as np

from levelsetfortran_b200 import vti


def test_header_known_answer():
    # cube40.stl: nx = ny = nz = 61, xLo = (-1.5, -1.5, -1.5) (stl.grid_from_surface), dx = 0.05
    h = vti.header(61, 61, 61, (-1.5, -1.5, -1.5), 0.05).decode()
    ext = " 0     61 0     61 0     61"
    org = "         -1.50000000          -1.50000000          -1.50000000"
    spc = "          0.05000000           0.05000000           0.05000000"
    want = ('<?xml version="1.0"?>\n'
            '<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian">\n'
            f'<ImageData WholeExtent="{ext}" Origin="{org}" Spacing="{spc}">\n'
            f'<Piece Extent="{ext}">\n'
            '<PointData Scalars="phi">\n'
            '<DataArray type="Float64" Name="phi" format="appended" offset="               0"/>\n'
            '</PointData>\n</Piece>\n</ImageData>\n<AppendedData encoding="raw">\n_')
    assert h == want


def test_length_field_keeps_the_reference_quirk():
    assert np.frombuffer(vti.nbyte_field(61), dtype="<i4")[0] == 62 ** 3 * 24            # set3d.f90:330, not 8*npts
    assert np.frombuffer(vti.nbyte_field(1023), dtype="<i4")[0] == np.int64(1024 ** 3 * 24).astype(np.int32)   # wraps


def test_block_layout_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    phi = np.asfortranarray(rng.standard_normal((5, 4, 3)))
    p = tmp_path / "f.vti"
    vti.write_vti(p, phi, (0.25, -1.0, 3.0), 0.05)
    raw = p.read_bytes()
    h = vti.header(4, 3, 2, (0.25, -1.0, 3.0), 0.05)
    assert raw.startswith(h) and raw.endswith(b"\n</AppendedData>\n</VTKFile>\n")
    assert len(raw) == len(h) + 4 + 8 * phi.size + len(vti.TRAILER)
    body = np.frombuffer(raw, dtype="<f8", count=phi.size, offset=len(h) + 4)
    assert body[1] == phi[1, 0, 0] and body[5] == phi[0, 1, 0] and body[20] == phi[0, 0, 1]      # i fastest, then j, then k
    back, origin, dx = vti.read_vti(p)
    assert np.array_equal(back, phi) and np.allclose(origin, (0.25, -1.0, 3.0)) and dx == 0.05
