// tools/fast_accuracy.cpp -- how far FastArith (with the build's -D switches) drifts from ExactArith over a full reinit.
// Serial Gauss-Seidel sweeps (reference loop nest subs.f90:742-852, boundary block in closed form) on a field read from a
// raw file, both arithmetics side by side from the same start; prints max |fast - exact| every 100 sweeps.  Host code only:
// the same lsf_cell.cuh the kernels compile, with the host fallbacks of the device intrinsics (no MUFU: reciprocals and
// square roots are exact here, so this isolates the ALGEBRAIC changes -- eps rounding, sign tricks -- from the seed error).
//   g++ -O2 -ffp-contract=off -std=c++17 [-DLSF_EPS_MODE=1] -I levelsetfortran_b200/csrc tools/fast_accuracy.cpp -o /tmp/fa
//   /tmp/fa field.bin nx ny nz dx h nsweeps [exact_out.bin]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "lsf_cell.cuh"
using namespace lsf;
static const int TAB[8][3] = {{+1, +1, +1}, {+1, +1, -1}, {+1, -1, -1}, {-1, -1, -1}, {-1, +1, -1}, {-1, -1, +1}, {-1, +1, +1}, {+1, -1, +1}};

template <class AR>
static void sweep(std::vector<double> &phi, const std::vector<double> &phiS, int nx, int ny, int nz, const CellConst &cc, int raster)
{
    const long sx = nx + 1, sxy = (long)(nx + 1) * (ny + 1);
    const int *d = TAB[raster - 1];
    for (int kk = 1; kk <= nz - 1; ++kk) {
        const int k = d[2] > 0 ? kk : nz - kk;
        for (int jj = 1; jj <= ny - 1; ++jj) {
            const int j = d[1] > 0 ? jj : ny - jj;
            for (int ii = 1; ii <= nx - 1; ++ii) {
                const int i = d[0] > 0 ? ii : nx - ii;
                const long c = i + sx * j + sxy * k;
                const bool hi = i >= 4 && i <= nx - 5 && j >= 4 && j <= ny - 5 && k >= 4 && k <= nz - 5;
                double vx[7], vy[7], vz[7];
                for (int m = -3; m <= 3; ++m) {
                    const bool in = hi || (m >= -1 && m <= 1);
                    vx[3 + m] = in ? phi[c + m] : 0.;
                    vy[3 + m] = in ? phi[c + m * sx] : 0.;
                    vz[3 + m] = in ? phi[c + m * sxy] : 0.;
                }
                double g[3], gM;
                bool sens;
                phi[c] = reinit_cell<AR>(vx, vy, vz, phiS[c], hi, cc, g, gM, sens);
            }
        }
    }
    // boundary block, closed form (k_reinit_bc_rms): phi(c) = phi(clamp(c)) + dx applied min(1+H, B) times
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) {
                const int B = (i == 0 || i == nx) + (j == 0 || j == ny) + (k == 0 || k == nz);
                if (!B) continue;
                const int H = (i == nx) + (j == ny) + (k == nz);
                const int m = 1 + H < B ? 1 + H : B;
                const int ic = i < 1 ? 1 : (i > nx - 1 ? nx - 1 : i), jc = j < 1 ? 1 : (j > ny - 1 ? ny - 1 : j), kc = k < 1 ? 1 : (k > nz - 1 ? nz - 1 : k);
                double v = phi[ic + sx * jc + sxy * kc];
                for (int r = 0; r < m; ++r) v = v + cc.dx;
                phi[i + sx * j + sxy * k] = v;
            }
}

int main(int argc, char **argv)
{
    if (argc < 8) { fprintf(stderr, "usage: %s field.bin nx ny nz dx h nsweeps [exact_out.bin]\n", argv[0]); return 2; }
    const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]), ns = atoi(argv[7]);
    const double dx = atof(argv[5]), h = atof(argv[6]);
    const size_t n = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    std::vector<double> s(n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(s.data(), 8, n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    std::vector<double> a = s, b = s;
    CellConst cc; cc.dx = dx; cc.inv_dx = 1. / dx; cc.k12 = 1. / (12. * dx); cc.dx2 = dx * dx; cc.h = h;
    double worst = 0;
    for (int it = 0; it < ns; ++it) {
        const int r = it % 8 + 1;
        sweep<ExactArith>(a, s, nx, ny, nz, cc, r);
        sweep<FastArith>(b, s, nx, ny, nz, cc, r);
        if ((it + 1) % 100 == 0 || it + 1 == ns || it < 8) {
            double e = 0;
            for (size_t q = 0; q < n; ++q) { const double d = fabs(a[q] - b[q]); if (!(d <= e)) e = d; }
            if (e > worst) worst = e;
            printf("sweep %5d  max|fast-exact| %.3e\n", it + 1, e);
            fflush(stdout);
        }
    }
    printf("worst %.3e\n", worst);
    if (argc > 8) { f = fopen(argv[8], "wb"); fwrite(a.data(), 8, n, f); fclose(f); }
    return 0;
}
