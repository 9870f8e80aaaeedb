/*
 * TEST INFRASTRUCTURE -- C restatement of the pmesh window scatter (pm.paint as called from
 * nbodykit/source/mesh/catalog.py:287,295-296), same arithmetic as oracle/pmesh_oracle.py::paint:
 *   g = fl(fl((double)pos * fl(N/L)) + shift); CIC i0=floor(g), TSC i0=floor(g+0.5)-1, PCS i0=floor(g)-1;
 *   weight = ((wx*wy)*wz)*mass accumulated in double; periodic wrap.
 * Used only as the checker at sizes NumPy is too slow for, and as the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference).  Per-particle loop like pmesh's; OpenMP threads share the
 * mesh through atomic adds.  Build: oracle/build_c.py (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static double kern(int sup, double x) {
    x = fabs(x);
    if (sup == 3) { if (x <= 0.5) return 0.75 - x * x; if (x < 1.5) { double t = 1.5 - x; return 0.5 * (t * t); } return 0.0; }
    if (sup == 4) { if (x < 1.0) return (4.0 - 6.0 * x * x + 3.0 * (x * x * x)) / 6.0; if (x < 2.0) { double t = 2.0 - x; return (t * t * t) / 6.0; } return 0.0; }
    return 0.0;
}

static void window(int sup, double g, long long *i0, double *w) {
    if (sup == 1) { *i0 = (long long)floor(g + 0.5); w[0] = 1.0; return; }
    if (sup == 2) { double f = floor(g); double d = g - f; *i0 = (long long)f; w[0] = 1.0 - d; w[1] = d; return; }
    double f = (sup == 3) ? floor(g + 0.5) - 1.0 : floor(g) - 1.0;
    double d = g - f;
    *i0 = (long long)f;
    for (int r = 0; r < sup; r++) w[r] = kern(sup, d - (double)r);
}

static long long wrapi(long long i, long long n) { long long r = i % n; return r < 0 ? r + n : r; }

/* pos: [n][3] f4 (pos_f4 != 0) or f8; mass: [n] f8 or NULL; mesh: [N0][N1][N2] f8, accumulated into */
void oracle_paint(const void *pos, int pos_f4, int64_t n, const double *mass, int sup, double shift,
                  const double *box, const int64_t *nmesh, double *mesh) {
    double scale[3];
    for (int d = 0; d < 3; d++) scale[d] = (double)nmesh[d] / box[d];
    const float *pf = (const float *)pos;
    const double *pd = (const double *)pos;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        long long i0[3];
        double w[3][4];
        for (int d = 0; d < 3; d++) {
            double p = pos_f4 ? (double)pf[3 * i + d] : pd[3 * i + d];
            double g = p * scale[d] + shift;
            window(sup, g, &i0[d], w[d]);
        }
        double m = mass ? mass[i] : 1.0;
        for (int rx = 0; rx < sup; rx++) {
            long long ix = wrapi(i0[0] + rx, nmesh[0]);
            for (int ry = 0; ry < sup; ry++) {
                long long iy = wrapi(i0[1] + ry, nmesh[1]);
                double wxy = w[0][rx] * w[1][ry];
                double *row = mesh + (ix * nmesh[1] + iy) * nmesh[2];
                for (int rz = 0; rz < sup; rz++) {
                    long long iz = wrapi(i0[2] + rz, nmesh[2]);
                    double wt = wxy * w[2][rz] * m;
#pragma omp atomic
                    row[iz] += wt;
                }
            }
        }
    }
}
