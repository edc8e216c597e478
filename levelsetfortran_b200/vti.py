"""The reference's ParaView writer, byte for byte (set3d.f90:316-351 `signedDistanceFunction.vti`, :546-569
`smoothedDistanceFunction.vti`).  Host-side, pure numpy: the Fortran driver keeps its own writer, this
is the same file for drivers of the Python mirror.

The file is a stream of the literal strings of the WRITE statements, the 1024-character work strings TRIMmed
(trailing blanks only: the leading blank of ' 0 ' and the left padding of I16 / F20.8 stay), one raw block
`nbytePhi, (((phi(i,j,k),i=0,nx),j=0,ny),k=0,nz)` and the closing tags.  Two quirks are kept because the output has to
be unchanged: `nbytePhi = (nx+1)**3*24` (set3d.f90:330 -- neither the cube nor the factor 24 matches the 8*(nx+1)(ny+1)
(nz+1) bytes that follow, and the default INTEGER wraps for nx >= 446), and the 4-byte length field."""
from __future__ import annotations

import numpy as np

LF = b"\n"


def _i6(n):
    s = "%6d" % n
    return s if len(s) == 6 else "*" * 6          # Fortran I6 overflow


def _f20_8(v):
    s = "%20.8f" % v
    return s if len(s) == 20 else "*" * 20


def header(nx, ny, nz, xLo, dx):
    """The bytes up to and including the '_' that starts the appended block."""
    extent = "".join(" 0 " + _i6(n) for n in (nx, ny, nz)).rstrip(" ")                  # '(3(A3,I6))'
    origin = "".join(_f20_8(float(v)) + " " for v in xLo).rstrip(" ")                   # '(3(F20.8,A1))'
    spacing = "".join(_f20_8(float(dx)) + " " for _ in range(3)).rstrip(" ")
    coffset = ("%16d" % 0).rstrip(" ")                                                   # '(I16)'
    parts = ['<?xml version="1.0"?>', "\n",
             '<VTKFile type="ImageData" version="0.1" byte_order="LittleEndian">', "\n",
             '<ImageData WholeExtent="', extent, '" Origin="', origin, '" Spacing="', spacing, '">', "\n",
             '<Piece Extent="', extent, '">', "\n",
             '<PointData Scalars="phi">', "\n",
             '<DataArray type="Float64" Name="phi" format="appended" offset="', coffset, '"/>', "\n",
             "</PointData>", "\n", "</Piece>", "\n", "</ImageData>", "\n",
             '<AppendedData encoding="raw">', "\n", "_"]
    return "".join(parts).encode("ascii")


def nbyte_field(nx):
    """set3d.f90:330 in default-kind INTEGER arithmetic (two's-complement wrap, as gfortran -O3 produces)."""
    v = ((nx + 1) ** 3 * 24) & 0xFFFFFFFF
    return np.array([v], dtype="<u4").tobytes()


TRAILER = LF + b"</AppendedData>" + LF + b"</VTKFile>" + LF


def write_vti(path, phi, xLo, dx):
    """phi: Fortran-ordered float64 (nx+1, ny+1, nz+1) == phi(0:nx,0:ny,0:nz)."""
    phi = np.asarray(phi)
    if phi.dtype != np.float64 or phi.ndim != 3 or not phi.flags.f_contiguous:
        raise ValueError("phi must be a Fortran-ordered float64 array phi(0:nx,0:ny,0:nz)")
    nx, ny, nz = (s - 1 for s in phi.shape)
    with open(path, "wb") as f:
        f.write(header(nx, ny, nz, xLo, dx))
        f.write(nbyte_field(nx))
        f.write(phi.astype("<f8", copy=False).tobytes(order="F"))      # i fastest
        f.write(TRAILER)


def read_vti(path):
    """Inverse of write_vti (extents from the header, the length field ignored as ParaView's raw reader would not)."""
    import re
    raw = open(path, "rb").read()
    m = re.search(rb'WholeExtent="\s*0\s+(\d+)\s+0\s+(\d+)\s+0\s+(\d+)"\s+Origin="([^"]*)"\s+Spacing="([^"]*)"', raw)
    if not m:
        raise IOError(f"{path}: not a LevelSetFortran .vti")
    nx, ny, nz = (int(m.group(q)) for q in (1, 2, 3))
    origin = np.array([float(v) for v in m.group(4).split()])
    dx = float(m.group(5).split()[0])
    start = raw.index(b'<AppendedData encoding="raw">\n_') + len(b'<AppendedData encoding="raw">\n_') + 4
    n = (nx + 1) * (ny + 1) * (nz + 1)
    phi = np.frombuffer(raw, dtype="<f8", count=n, offset=start).reshape((nx + 1, ny + 1, nz + 1), order="F")
    return np.asfortranarray(phi), origin, dx


def write_vti_native(path, phi, xLo, dx):
    """The same file through the library's host entry point lsf_write_vti (what a Fortran driver would bind)."""
    import ctypes as C
    from ._lib import c_double_p, check, lib
    phi = np.asarray(phi)
    if phi.dtype != np.float64 or phi.ndim != 3 or not phi.flags.f_contiguous:
        raise ValueError("phi must be a Fortran-ordered float64 array phi(0:nx,0:ny,0:nz)")
    nx, ny, nz = (s - 1 for s in phi.shape)
    x = np.ascontiguousarray(xLo, dtype=np.float64)
    check(lib().lsf_write_vti(str(path).encode(), phi.ctypes.data_as(c_double_p), nx, ny, nz, x.ctypes.data_as(c_double_p), float(dx)))


def write_s3d_native(path, surfOrder, surfElem0, surfElemTag, surfXX, bndNormal, nBndElem=0):
    """set3d.f90:604-614 through lsf_write_s3d.  surfElem0: (nSurfElem,3) ZERO-based (the reference decrements at :590-594)."""
    import ctypes as C
    from ._lib import c_double_p, c_i32_p, check, lib
    so = np.ascontiguousarray(surfOrder, dtype=np.int32)
    se = np.asfortranarray(surfElem0, dtype=np.int32)
    st = np.ascontiguousarray(surfElemTag, dtype=np.int32)
    xx = np.asfortranarray(surfXX, dtype=np.float64)
    bn = np.asfortranarray(bndNormal, dtype=np.float64).reshape(-1, 3, order="F")
    check(lib().lsf_write_s3d(str(path).encode(), se.shape[0], xx.shape[0], int(nBndElem), bn.shape[0], so.ctypes.data_as(c_i32_p),
                              se.ctypes.data_as(c_i32_p), st.ctypes.data_as(c_i32_p), xx.ctypes.data_as(c_double_p),
                              bn.ctypes.data_as(c_double_p) if bn.size else None))
