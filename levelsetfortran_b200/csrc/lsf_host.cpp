// lsf_host.cpp -- host-side pieces either side of the grid hot path, callable from the reference's
// Fortran driver through BIND(C): the ParaView .vti writer (set3d.f90:320-351, :539-569), the .s3d mesh writer
// (set3d.f90:584-614) and stlRead's vertex de-duplication (subs.f90:68-93) in O(n) instead of O(ntri * nSurfNode).
// No device work in this file.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lsf_b200.h"

namespace lsf { int set_error(int code, const char *fmt, ...); }
using lsf::set_error;

// ------------------------------------------------------------------------------------------- .vti
// The file is the concatenation of the WRITE(sUnit) items of set3d.f90:336-350: literal strings, the 1024-character
// work strings TRIMmed (trailing blanks only), one raw block `nbytePhi,(((phi(i,j,k),i=0,nx),j=0,ny),k=0,nz)`.
// Kept quirks (the output has to be unchanged): nbytePhi = (nx+1)**3*24 in default INTEGER arithmetic (:330: neither
// the cube nor the 24 matches the 8(nx+1)(ny+1)(nz+1) bytes that follow; wraps for nx >= 446), 4-byte length field.
static std::string rtrim(const std::string &s)
{
    size_t n = s.size();
    while (n > 0 && s[n - 1] == ' ') --n;
    return s.substr(0, n);
}

static std::string fmt_i(long v, int w)            // Iw
{
    char b[64];
    int n = snprintf(b, sizeof b, "%ld", v);
    if (n > w) return std::string((size_t)w, '*');
    return std::string((size_t)(w - n), ' ') + b;
}

static std::string fmt_f(double v, int w, int d)   // Fw.d
{
    char b[512];
    int n = snprintf(b, sizeof b, "%.*f", d, v);
    std::string s(b, (size_t)n);
    if (n > w && s[0] == '0') s = s.substr(1);      // optional leading zero
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)(w - (int)s.size()), ' ') + s;
}

extern "C" int lsf_write_vti(const char *path, const double *phi, int nx, int ny, int nz, const double xLo[3], double dx)
{
    if (!path || !phi || !xLo) return set_error(LSF_ERR_ARG, "null argument");
    std::string extent, origin, spacing;
    const int n3[3] = {nx, ny, nz};
    for (int q = 0; q < 3; ++q) extent += " 0 " + fmt_i(n3[q], 6);                  // '(3(A3,I6))'       :327
    for (int q = 0; q < 3; ++q) origin += fmt_f(xLo[q], 20, 8) + " ";               // '(3(F20.8,A1))'    :328
    for (int q = 0; q < 3; ++q) spacing += fmt_f(dx, 20, 8) + " ";                  //                     :329
    const std::string coffset = fmt_i(0, 16);                                       // '(I16)'            :332
    const int32_t nbytePhi = (int32_t)((uint32_t)(nx + 1) * (uint32_t)(nx + 1) * (uint32_t)(nx + 1) * 24u);   // :330, wrapping
    std::string h;
    h += "<?xml version=\"1.0\"?>\n";
    h += "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
    h += "<ImageData WholeExtent=\"" + rtrim(extent) + "\" Origin=\"" + rtrim(origin) + "\" Spacing=\"" + rtrim(spacing) + "\">\n";
    h += "<Piece Extent=\"" + rtrim(extent) + "\">\n";
    h += "<PointData Scalars=\"phi\">\n";
    h += "<DataArray type=\"Float64\" Name=\"phi\" format=\"appended\" offset=\"" + rtrim(coffset) + "\"/>\n";
    h += "</PointData>\n</Piece>\n</ImageData>\n<AppendedData encoding=\"raw\">\n_";
    FILE *f = fopen(path, "wb");
    if (!f) return set_error(LSF_ERR_ARG, "cannot open %s", path);
    const size_t np = (size_t)(nx + 1) * (size_t)(ny + 1) * (size_t)(nz + 1);
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && fwrite(&nbytePhi, 4, 1, f) == 1 && fwrite(phi, 8, np, f) == np;
    static const char tail[] = "\n</AppendedData>\n</VTKFile>\n";
    ok = ok && fwrite(tail, 1, sizeof tail - 1, f) == sizeof tail - 1;
    ok = (fclose(f) == 0) && ok;
    return ok ? LSF_OK : set_error(LSF_ERR_ARG, "short write to %s", path);
}

// ------------------------------------------------------------------------------------------- .s3d
// set3d.f90:604-614: list-directed records.  WHICH values are written, in which order, comes from the reference; HOW a
// list-directed record is spaced is libgfortran's business (not in the reference's source): every record starts with a
// blank, INTEGER(4) items are right-justified in 11 columns, REAL(8) items are 1PG25.17E3-style (17 significant
// digits, F form with 5 trailing blanks for 0.1 <= |x| < 1e17, else E form with a 3-digit exponent), one blank between
// items.  Restated from libgfortran's documented behaviour; the same routine is used by the oracle's run time.
static void ld_int(std::string &o, long v) { o += ' '; o += fmt_i(v, 11); }

static void ld_real(std::string &o, double v)
{
    char out[64];
    o += ' ';
    if (isnan(v)) { snprintf(out, sizeof out, "%25s", "NaN"); o += out; return; }
    if (isinf(v)) { snprintf(out, sizeof out, "%25s", v > 0 ? "Infinity" : "-Infinity"); o += out; return; }
    if (v == 0.) { snprintf(out, sizeof out, "%20s     ", signbit(v) ? "-0.0000000000000000" : "0.0000000000000000"); o += out; return; }
    char e[64], dig[18];
    snprintf(e, sizeof e, "%.16e", fabs(v));
    dig[0] = e[0];
    memcpy(dig + 1, e + 2, 16);
    dig[17] = 0;
    const int x = atoi(strchr(e, 'e') + 1), e10 = x + 1;
    const std::string sg = v < 0 ? "-" : "";
    std::string body;
    if (e10 >= 0 && e10 <= 17) {
        body = e10 == 0 ? sg + "0." + dig : sg + std::string(dig, (size_t)e10) + "." + std::string(dig + e10);
        snprintf(out, sizeof out, "%20s     ", body.c_str());
    } else {
        char ex[16];
        snprintf(ex, sizeof ex, "E%c%03d", x < 0 ? '-' : '+', x < 0 ? -x : x);
        body = sg + std::string(1, dig[0]) + "." + std::string(dig + 1) + ex;
        snprintf(out, sizeof out, "%25s", body.c_str());
    }
    o += out;
}

// surfElem: (nSurfElem,3) column-major, 0-BASED as the reference makes it at :590-594 before writing; surfXX: (nSurfNode,3)
// column-major; bndNormal: (nBndComp,3) column-major (may be NULL when nBndComp == 0).
extern "C" int lsf_write_s3d(const char *path, int nSurfElem, int nSurfNode, int nBndElem, int nBndComp, const int32_t *surfOrder,
                             const int32_t *surfElem, const int32_t *surfElemTag, const double *surfXX, const double *bndNormal)
{
    if (!path || !surfOrder || !surfElem || !surfElemTag || !surfXX || (nBndComp > 0 && !bndNormal))
        return set_error(LSF_ERR_ARG, "null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return set_error(LSF_ERR_ARG, "cannot open %s", path);
    std::string o;
    o.reserve(1 << 20);
    ld_int(o, nSurfElem); ld_int(o, nSurfNode); ld_int(o, nBndElem); ld_int(o, nBndComp); o += '\n';        // :604
    for (int k = 0; k < nSurfElem; ++k) {                                                                   // :605-607
        ld_int(o, surfOrder[k]);
        for (int c = 0; c < 3; ++c) ld_int(o, surfElem[k + (size_t)c * nSurfElem]);
        ld_int(o, surfElemTag[k]);
        o += '\n';
        if (o.size() > (1 << 20) - 256) { fwrite(o.data(), 1, o.size(), f); o.clear(); }
    }
    for (int n = 0; n < nSurfNode; ++n) {                                                                   // :608-610
        for (int c = 0; c < 3; ++c) ld_real(o, surfXX[n + (size_t)c * nSurfNode]);
        o += '\n';
        if (o.size() > (1 << 20) - 256) { fwrite(o.data(), 1, o.size(), f); o.clear(); }
    }
    for (int n = 0; n < nBndComp; ++n) {                                                                    // :611-613
        for (int c = 0; c < 3; ++c) ld_real(o, bndNormal[n + (size_t)c * nBndComp]);
        o += '\n';
    }
    bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
    ok = (fclose(f) == 0) && ok;
    return ok ? LSF_OK : set_error(LSF_ERR_ARG, "short write to %s", path);
}

// ------------------------------------------------------------------------------------------- stlRead de-duplication
// subs.f90:68-93: vertex p of triangle n is matched against nodes kk = 1..nSurfNode (ascending, first match wins) with
//     abs(nodesT(c,kk) - v(c)) < 1.e-13  for c = 1..3   (REAL*4 difference, REAL*8 literal)
// where nSurfNode is 3 during the first triangle and is refreshed only AFTER each triangle (:93) -- a vertex repeated
// inside one later triangle is therefore stored twice.  For float data the predicate is equality (+0 == -0) unless
// both coordinates are tiny: two distinct floats can be closer than 1e-13 only below 2^23 * 1e-13 = 8.4e-7.  Hash key
// per coordinate: the bit pattern, or one shared marker for every tiny value; candidates of a bucket are scanned in
// node order with the reference's own predicate, so the numbering is the reference's for every input.
namespace {
struct Key {
    uint32_t k[3];
    bool operator==(const Key &o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
};
struct KeyHash {
    size_t operator()(const Key &a) const
    {
        uint64_t h = 1469598103934665603ull;
        for (int c = 0; c < 3; ++c) { h ^= a.k[c]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
inline uint32_t coord_key(float x)
{
    if (fabsf(x) < 8.5e-7f) return 0xffffffffu;      // every tiny value (incl. +-0) shares one key; compared exactly below
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
}
inline bool ref_match(const float *a, const float *b)
{
    for (int c = 0; c < 3; ++c) {
        const float d = a[c] - b[c];                 // REAL*4 subtraction
        if (!((double)fabsf(d) < 1.e-13)) return false;
    }
    return true;
}
}  // namespace

// tri: 9*ntri floats, vertex-major (x,y,z of vertex 1, 2, 3 of triangle 1, ...) -- the reference's triangles(3,ntri*3).
// nodes: work array of 3 * 3*ntri floats receiving nodesT(3,k); surfElem: (ntri,3) column-major, 1-based.
extern "C" int lsf_stl_dedup(const float *tri, int ntri, float *nodes, int32_t *surfElem, int *nSurfNode)
{
    if (!tri || !nodes || !surfElem || !nSurfNode || ntri < 0) return set_error(LSF_ERR_ARG, "null argument");
    std::unordered_map<Key, std::vector<int32_t>, KeyHash> buckets;
    buckets.reserve((size_t)ntri * 2 + 16);
    int32_t k = 0, window = 3;                       // nSurfNode = 3 (:68)
    for (int n = 0; n < ntri; ++n) {
        for (int p = 0; p < 3; ++p) {
            const float *v = tri + 9 * (size_t)n + 3 * p;
            Key key = {{coord_key(v[0]), coord_key(v[1]), coord_key(v[2])}};
            std::vector<int32_t> &cand = buckets[key];
            int32_t share = 0;
            for (int32_t kk : cand) {                // ascending node numbers
                if (kk > window) break;
                if (ref_match(nodes + 3 * (size_t)(kk - 1), v)) { share = kk; break; }
            }
            if (share > 0) surfElem[n + (size_t)p * ntri] = share;
            else {
                ++k;
                memcpy(nodes + 3 * (size_t)(k - 1), v, 12);
                surfElem[n + (size_t)p * ntri] = k;
                cand.push_back(k);
            }
        }
        window = k;                                  // nSurfNode = k (:93)
    }
    *nSurfNode = k;
    return LSF_OK;
}

// Binary STL: 80-byte header, INTEGER*4 ntri, per triangle 12 REAL*4 + INTEGER*2 (subs.f90:36-55).
extern "C" int lsf_stl_count(const char *path, int *ntri)
{
    if (!path || !ntri) return set_error(LSF_ERR_ARG, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return set_error(LSF_ERR_ARG, "cannot open %s", path);
    int32_t n = -1;
    const bool ok = fseek(f, 80, SEEK_SET) == 0 && fread(&n, 4, 1, f) == 1 && n >= 0;
    fclose(f);
    if (!ok) return set_error(LSF_ERR_ARG, "%s: not a binary STL", path);
    *ntri = n;
    return LSF_OK;
}

// tri receives 9*ntri floats (vertex-major); the facet normals are skipped as in the reference (:48, never used).
extern "C" int lsf_stl_read_triangles(const char *path, int ntri, float *tri)
{
    if (!path || !tri) return set_error(LSF_ERR_ARG, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return set_error(LSF_ERR_ARG, "cannot open %s", path);
    bool ok = fseek(f, 84, SEEK_SET) == 0;
    unsigned char rec[50];
    for (int n = 0; ok && n < ntri; ++n) {
        ok = fread(rec, 1, 50, f) == 50;
        if (ok) memcpy(tri + 9 * (size_t)n, rec + 12, 36);
    }
    fclose(f);
    return ok ? LSF_OK : set_error(LSF_ERR_ARG, "%s: truncated STL", path);
}
