"""Host-side surface I/O and grid set-up: mirror of stlRead (subs.f90:17-121), the bounding-box /
grid definition block of set3d.f90:86-186, a binary STL writer and the synthetic closed surfaces
the benchmark configurations use (sphere; torus + cube).  Pure numpy; no device work."""
from __future__ import annotations

import math
import struct

import numpy as np


def stl_write(path, tris, header=b"levelsetfortran_b200 synthetic"):
    """tris: (ntri,3,3) float32 vertices.  Normals are written as the (unnormalised -> normalised)
    right-hand-rule normal; the reference ignores them (subs.f90:48)."""
    tris = np.asarray(tris, dtype=np.float32)
    n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]).astype(np.float64)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.where(ln > 0, ln, 1), 0).astype(np.float32)
    rec = np.zeros(len(tris), dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("pad", "<u2")])
    rec["n"], rec["v"] = n, tris
    with open(path, "wb") as f:
        f.write(header.ljust(80, b" ")[:80])
        f.write(struct.pack("<i", len(tris)))
        f.write(rec.tobytes())


def stl_triangles(path):
    """Raw triangles (ntri,3,3) float32 of a binary STL (80-byte header, int32 count, 50-byte records)."""
    with open(path, "rb") as f:
        f.seek(80)
        (ntri,) = struct.unpack("<i", f.read(4))
        rec = np.frombuffer(f.read(50 * ntri), dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("pad", "<u2")])
    if len(rec) != ntri:
        raise IOError(f"{path}: truncated STL")
    return rec["v"].copy()


def dedup_nodes(tris):
    """Vertex de-duplication with stlRead's numbering (subs.f90:68-93): nodes are numbered by first
    occurrence; a vertex is matched against nodes 1..nSurfNode where nSurfNode is refreshed only after
    each triangle (3 during the first), so a vertex repeated inside one later triangle is stored twice.
    The reference compares REAL*4 coordinates with abs(diff) < 1.e-13, which for float32 data is
    equality (+0 == -0); sub-1e-13 near-zero distinct values are treated as distinct here.
    Returns surfX (nSurfNode,3) float64 F-order and surfElem (ntri,3) int32 F-order, 1-based."""
    tris = np.asarray(tris, dtype=np.float32) + np.float32(0.0)   # canonicalise -0.0
    ntri = len(tris)
    first = {}
    nodes = []
    elem = np.zeros((ntri, 3), dtype=np.int32, order="F")
    window = 3
    keys = tris.view(np.uint32).reshape(ntri, 3, 3)
    for n in range(ntri):
        for p in range(3):
            key = (int(keys[n, p, 0]), int(keys[n, p, 1]), int(keys[n, p, 2]))
            idx = first.get(key, 0)
            if idx and idx <= window:
                elem[n, p] = idx
            else:
                nodes.append(tris[n, p])
                if not idx:
                    first[key] = len(nodes)
                elem[n, p] = len(nodes)
        window = len(nodes)
    surfX = np.asfortranarray(np.asarray(nodes, dtype=np.float32).astype(np.float64))
    return surfX, elem


def stlRead(path):
    """SUBROUTINE stlRead (subs.f90:17-121): surfX, nSurfNode, surfElem, nSurfElem."""
    surfX, surfElem = dedup_nodes(stl_triangles(path))
    return surfX, surfX.shape[0], surfElem, surfElem.shape[0]


def stlRead_native(path):
    """stlRead through the library's host entry points (lsf_stl_count / lsf_stl_read_triangles / lsf_stl_dedup, the ones
    a Fortran driver binds): same four results, the reference's exact match predicate, O(n)."""
    import ctypes as C
    from ._lib import check, lib
    L = lib()
    ntri = C.c_int(0)
    check(L.lsf_stl_count(str(path).encode(), C.byref(ntri)))
    tri = np.empty(9 * ntri.value, dtype=np.float32)
    check(L.lsf_stl_read_triangles(str(path).encode(), ntri.value, tri.ctypes.data_as(C.POINTER(C.c_float))))
    return dedup_native(tri.reshape(ntri.value, 3, 3))


def dedup_native(tris):
    import ctypes as C
    from ._lib import check, lib
    tri = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1)
    ntri = tri.size // 9
    nodes = np.empty(max(9 * ntri, 1), dtype=np.float32)
    elem = np.zeros((ntri, 3), dtype=np.int32, order="F")
    nn = C.c_int(0)
    check(lib().lsf_stl_dedup(tri.ctypes.data_as(C.POINTER(C.c_float)), ntri, nodes.ctypes.data_as(C.POINTER(C.c_float)),
                              elem.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nn)))
    surfX = np.asfortranarray(nodes[: 3 * nn.value].reshape(nn.value, 3).astype(np.float64))    # surfX(k,c) = nodesT(c,k), subs.f90:99-103
    return surfX, nn.value, elem, ntri


def grid_from_surface(surfX, dx=0.05, dd=10):
    """Grid definition of set3d.f90:90-186 and the normalised step of :301.
    Returns dict(nx, ny, nz, xLo, xHi, box=(im,ip,jm,jp,km,kp), dx, dxx)."""
    surfX = np.asarray(surfX, dtype=np.float64)
    mn, mx = surfX.min(axis=0), surfX.max(axis=0)
    ext = mx - mn                                                     # ddx,ddy,ddz :135-137
    n = [int(math.ceil(float(e) / dx)) + 1 + 2 * dd for e in ext]    # :143-153
    xLo = np.array([float(m) - dd * dx for m in mn])                  # :156
    xHi = np.array([float(m) + dd * dx for m in mx])
    box = []
    for c in range(3):
        box.append(int(math.floor((float(mn[c]) - xLo[c]) / dx)) - 3)   # :180-182
        box.append(int(math.floor((float(mx[c]) - xLo[c]) / dx)) + 3)   # :184-186
    dxx = dx / math.sqrt(float(ext[0]) * float(ext[0]) + float(ext[1]) * float(ext[1]) + float(ext[2]) * float(ext[2]))
    return dict(nx=n[0], ny=n[1], nz=n[2], xLo=xLo, xHi=xHi, box=tuple(box), dx=dx, dxx=dxx)


# ------------------------------------------------------------------------------- synthetic surfaces
def sphere_tris(diameter, n_lat=101, n_lon=100, center=(0.0, 0.0, 0.0)):
    """Closed UV sphere, outward (counter-clockwise from outside) triangles:
    2*n_lon*(n_lat-1) of them (defaults: 20 000)."""
    R = 0.5 * diameter
    th = np.pi * np.arange(1, n_lat) / n_lat
    ph = 2.0 * np.pi * np.arange(n_lon) / n_lon
    ring = np.stack([np.outer(np.sin(th), np.cos(ph)), np.outer(np.sin(th), np.sin(ph)),
                     np.repeat(np.cos(th)[:, None], n_lon, axis=1)], axis=-1) * R    # (n_lat-1, n_lon, 3)
    north, south = np.array([0.0, 0.0, R]), np.array([0.0, 0.0, -R])
    tris = []
    nxt = lambda j: (j + 1) % n_lon
    for j in range(n_lon):
        tris.append([north, ring[0, j], ring[0, nxt(j)]])
        tris.append([south, ring[-1, nxt(j)], ring[-1, j]])
    for r in range(n_lat - 2):
        for j in range(n_lon):
            a, b, c, d = ring[r, j], ring[r + 1, j], ring[r + 1, nxt(j)], ring[r, nxt(j)]
            tris.append([a, b, c])
            tris.append([a, c, d])
    return (np.asarray(tris) + np.asarray(center)).astype(np.float32)


def torus_tris(R, r, n_major=96, n_minor=48, center=(0.0, 0.0, 0.0)):
    """Closed torus around the z axis (major radius R, tube radius r), outward triangles."""
    u = 2.0 * np.pi * np.arange(n_major) / n_major
    v = 2.0 * np.pi * np.arange(n_minor) / n_minor
    P = np.stack([np.outer(np.cos(u), R + r * np.cos(v)), np.outer(np.sin(u), R + r * np.cos(v)),
                  np.repeat((r * np.sin(v))[None, :], n_major, axis=0)], axis=-1)
    tris = []
    for i in range(n_major):
        i2 = (i + 1) % n_major
        for j in range(n_minor):
            j2 = (j + 1) % n_minor
            a, b, c, d = P[i, j], P[i2, j], P[i2, j2], P[i, j2]
            tris.append([a, b, c])
            tris.append([a, c, d])
    return (np.asarray(tris) + np.asarray(center)).astype(np.float32)


def box_tris(lo, hi, n=8):
    """Closed axis-aligned box with n x n quads (2 outward triangles each) per face."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    tris = []
    for ax in range(3):
        u, v = (ax + 1) % 3, (ax + 2) % 3
        for side in (0, 1):
            for p in range(n):
                for q in range(n):
                    def pt(pp, qq):
                        x = np.zeros(3)
                        x[ax] = hi[ax] if side else lo[ax]
                        x[u] = lo[u] + (hi[u] - lo[u]) * pp / n
                        x[v] = lo[v] + (hi[v] - lo[v]) * qq / n
                        return x
                    a, b, c, d = pt(p, q), pt(p + 1, q), pt(p + 1, q + 1), pt(p, q + 1)
                    tris += ([[a, b, c], [a, c, d]] if side else [[a, c, b], [a, d, c]])
    return np.asarray(tris).astype(np.float32)


def points_per_axis_to_extent(npts, dx=0.05, dd=10, margin=0.2):
    """Largest bbox extent that still gives `npts` grid points on an axis:
    npts = ceiling(ext/dx) + 2 + 2*dd (set3d.f90:143-153), minus `margin` cells of slack."""
    return (npts - 2 - 2 * dd - margin) * dx


def sphere_config(npts=512, dx=0.05, n_lat=101, n_lon=100):
    """BASELINE config 3: ~20k-triangle sphere whose grid has npts points per axis."""
    return sphere_tris(points_per_axis_to_extent(npts, dx), n_lat, n_lon)


def torus_cube_config(npts=(1024, 1024, 1024), dx=0.05):
    """BASELINE configs 4/5: a torus and a disjoint cube whose union bbox gives npts grid points per
    axis (z may be elongated for the weak-scaling sweep)."""
    ex, ey, ez = (points_per_axis_to_extent(n, dx) for n in npts)
    s = min(ex, ey)
    r = 0.12 * s
    R = 0.5 * s - r                     # torus spans [-s/2, s/2] in x and y, centred at the origin
    zt = -0.5 * ez + r                  # torus hugs the low-z face of the bbox
    torus = torus_tris(R, r, center=(0.0, 0.0, zt))
    c = 0.18 * s                        # cube in the high corner, disjoint from the torus in z
    hi = np.array([0.5 * ex, 0.5 * ey, 0.5 * ez])
    cube = box_tris(hi - c, hi)
    lo_x = np.array([-0.5 * ex, -0.5 * ey])
    # make the union bbox exact in x/y even when ex != ey: add a thin sliver box at the low corner
    parts = [torus, cube]
    if abs(ex - s) > 1e-9 or abs(ey - s) > 1e-9:
        parts.append(box_tris([lo_x[0], lo_x[1], 0.5 * ez - c], [lo_x[0] + c, lo_x[1] + c, 0.5 * ez]))
    if zt + r > 0.5 * ez - c:
        raise ValueError("grid too thin in z for a disjoint torus + cube")
    return np.concatenate(parts, axis=0)
