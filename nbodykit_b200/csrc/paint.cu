// Particle -> mesh window scatter (pmesh `pm.paint`, called from source/mesh/catalog.py:287,295-296).
//
// Compiled with --fmad=false: the grid coordinate g = fl(fl(double(pos)*fl(N/L)) + shift) and the
// f8 window weights must round exactly as the oracle does; cell indices are
// bit-exact by contract.
//
// Path "direct": one particle per thread, support^3 native L2 reductions (REDG.E.ADD.F32/.F64)
// into the mesh.  sm_100a has no native shared-memory float/64-bit atomic add (ATOMS.CAST.SPIN
// loops), while REDG f32/f64 is native, so the mesh itself is the accumulator and the L2 (126 MB)
// absorbs the read-modify-write of spatially coherent catalogues.
#include "common.cuh"

template <int SUP> struct Window;

// nearest grid point
template <> struct Window<1> {
    __device__ static __forceinline__ void eval(double g, long long &i0, double *w) {
        i0 = (long long)floor(g + 0.5);
        w[0] = 1.0;
    }
};
// CIC: i0 = floor(g), w = (1-d, d)
template <> struct Window<2> {
    __device__ static __forceinline__ void eval(double g, long long &i0, double *w) {
        double f = floor(g);
        double d = g - f;
        i0 = (long long)f;
        w[0] = 1.0 - d;
        w[1] = d;
    }
};
__device__ __forceinline__ double tsc_kernel(double x) {
    x = fabs(x);
    if (x <= 0.5) return 0.75 - x * x;
    if (x < 1.5) { double t = 1.5 - x; return 0.5 * (t * t); }
    return 0.0;
}
// TSC: i0 = floor(g + 0.5) - 1, w_r = K(g - i0 - r)
template <> struct Window<3> {
    __device__ static __forceinline__ void eval(double g, long long &i0, double *w) {
        double f = floor(g + 0.5) - 1.0;
        double d = g - f;
        i0 = (long long)f;
#pragma unroll
        for (int r = 0; r < 3; r++) w[r] = tsc_kernel(d - (double)r);
    }
};
__device__ __forceinline__ double pcs_kernel(double x) {
    x = fabs(x);
    if (x < 1.0) return (4.0 - 6.0 * x * x + 3.0 * (x * x * x)) / 6.0;
    if (x < 2.0) { double t = 2.0 - x; return (t * t * t) / 6.0; }
    return 0.0;
}
template <> struct Window<4> {
    __device__ static __forceinline__ void eval(double g, long long &i0, double *w) {
        double f = floor(g) - 1.0;
        double d = g - f;
        i0 = (long long)f;
#pragma unroll
        for (int r = 0; r < 4; r++) w[r] = pcs_kernel(d - (double)r);
    }
};

struct PaintGeom {
    double scale[3];   // fl(N_d / L_d)
    int n[3];          // Nmesh
    int x_start, x_n;  // owned x planes
};

__device__ __forceinline__ int wrap(long long i, int n) {
    long long r = i % n;
    return (int)(r < 0 ? r + n : r);
}

template <typename PT>
__device__ __forceinline__ bool load_grid(const PT *__restrict__ pos, int64_t i, const PaintGeom &gm,
                                          double shift, double *g) {
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double p = (double)pos[3 * i + d];
        g[d] = p * gm.scale[d] + shift;  // two roundings; this file is built with --fmad=false
        ok = ok && isfinite(g[d]);
    }
    return ok;
}

template <int SUP, typename FT>
__device__ __forceinline__ void scatter(const double *g, double mass, const PaintGeom &gm, FT *__restrict__ mesh) {
    long long i0[3];
    double w[3][SUP];
#pragma unroll
    for (int d = 0; d < 3; d++) Window<SUP>::eval(g[d], i0[d], w[d]);
    int iz[SUP], iy[SUP];
#pragma unroll
    for (int r = 0; r < SUP; r++) {
        iz[r] = wrap(i0[2] + r, gm.n[2]);
        iy[r] = wrap(i0[1] + r, gm.n[1]);
    }
#pragma unroll
    for (int rx = 0; rx < SUP; rx++) {
        int ix = wrap(i0[0] + rx, gm.n[0]) - gm.x_start;
        if (ix < 0 || ix >= gm.x_n) continue;  // ghost semantics: not my plane
#pragma unroll
        for (int ry = 0; ry < SUP; ry++) {
            double wxy = w[0][rx] * w[1][ry];
            int64_t row = ((int64_t)ix * gm.n[1] + iy[ry]) * gm.n[2];
#pragma unroll
            for (int rz = 0; rz < SUP; rz++) {
                double wt = wxy * w[2][rz] * mass;
                atomicAdd(&mesh[row + iz[rz]], (FT)wt);  // REDG.E.ADD.{F32,F64}
            }
        }
    }
}

template <int SUP, typename PT, typename MT, typename FT, bool INTERLACED>
__global__ void __launch_bounds__(256)
k_paint_direct(const PT *__restrict__ pos, const MT *__restrict__ mass, int64_t n, PaintGeom gm, double shift,
               FT *__restrict__ mesh, FT *__restrict__ mesh2) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double g[3];
        double m = mass ? (double)mass[i] : 1.0;
        if (load_grid(pos, i, gm, shift, g)) scatter<SUP, FT>(g, m, gm, mesh);
        if (INTERLACED) {
            if (load_grid(pos, i, gm, 0.5, g)) scatter<SUP, FT>(g, m, gm, mesh2);
        }
    }
}

template <int SUP, typename PT>
__global__ void __launch_bounds__(256)
k_cell_index(const PT *__restrict__ pos, int64_t n, PaintGeom gm, double shift, int32_t *__restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double g[3];
        bool ok = load_grid(pos, i, gm, shift, g);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            long long i0 = 0;
            double w[SUP];
            if (ok) Window<SUP>::eval(g[d], i0, w);
            out[3 * i + d] = ok ? wrap(i0, gm.n[d]) : -1;
        }
    }
}

static int make_geom(const double *box, const int64_t *nmesh, int64_t x_start, int64_t x_n, PaintGeom &gm) {
    for (int d = 0; d < 3; d++) {
        NBK_CHECK_ARG(nmesh[d] > 0 && nmesh[d] < (1 << 30), "paint: bad Nmesh[%d]=%lld", d, (long long)nmesh[d]);
        NBK_CHECK_ARG(box[d] > 0, "paint: bad BoxSize[%d]=%g", d, box[d]);
        gm.n[d] = (int)nmesh[d];
        gm.scale[d] = (double)nmesh[d] / box[d];
    }
    NBK_CHECK_ARG(x_start >= 0 && x_n >= 0 && x_start + x_n <= nmesh[0], "paint: bad slab [%lld,+%lld)",
                  (long long)x_start, (long long)x_n);
    gm.x_start = (int)x_start;
    gm.x_n = (int)x_n;
    return NBK_OK;
}

template <int SUP, typename PT, typename MT, typename FT>
static int launch_paint2(const void *pos, const void *mass, int64_t n, const PaintGeom &gm, double shift, void *mesh,
                         void *mesh2, cudaStream_t s) {
    int g = nbk_grid_for(n, 256, 8);
    if (mesh2)
        k_paint_direct<SUP, PT, MT, FT, true>
            <<<g, 256, 0, s>>>((const PT *)pos, (const MT *)mass, n, gm, shift, (FT *)mesh, (FT *)mesh2);
    else
        k_paint_direct<SUP, PT, MT, FT, false>
            <<<g, 256, 0, s>>>((const PT *)pos, (const MT *)mass, n, gm, shift, (FT *)mesh, (FT *)nullptr);
    NBK_LAUNCHED();
    return NBK_OK;
}

template <int SUP, typename PT, typename MT>
static int launch_paint1(const void *pos, const void *mass, int64_t n, const PaintGeom &gm, double shift, void *mesh,
                         void *mesh2, int mesh_dtype, cudaStream_t s) {
    if (mesh_dtype == NBK_F4) return launch_paint2<SUP, PT, MT, float>(pos, mass, n, gm, shift, mesh, mesh2, s);
    return launch_paint2<SUP, PT, MT, double>(pos, mass, n, gm, shift, mesh, mesh2, s);
}

template <int SUP>
static int launch_paint0(const void *pos, int pos_dtype, const void *mass, int mass_dtype, int64_t n,
                         const PaintGeom &gm, double shift, void *mesh, void *mesh2, int mesh_dtype, cudaStream_t s) {
    bool pf4 = pos_dtype == NBK_F4, mf4 = (mass_dtype == NBK_F4);
    if (pf4 && mf4) return launch_paint1<SUP, float, float>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, s);
    if (pf4) return launch_paint1<SUP, float, double>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, s);
    if (mf4) return launch_paint1<SUP, double, float>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, s);
    return launch_paint1<SUP, double, double>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, s);
}

static int paint_impl(const void *pos, int pos_dtype, int64_t n, const void *mass, int mass_dtype, int window,
                      double shift, const double *box, const int64_t *nmesh, int64_t x_start, int64_t x_n,
                      void *mesh, void *mesh2, int mesh_dtype, void *stream) {
    NBK_CHECK_ARG(pos_dtype == NBK_F4 || pos_dtype == NBK_F8, "paint: bad pos dtype %d", pos_dtype);
    NBK_CHECK_ARG(mesh_dtype == NBK_F4 || mesh_dtype == NBK_F8, "paint: bad mesh dtype %d", mesh_dtype);
    NBK_CHECK_ARG(mass == nullptr || mass_dtype == NBK_F4 || mass_dtype == NBK_F8, "paint: bad mass dtype %d",
                  mass_dtype);
    NBK_CHECK_ARG(n >= 0, "paint: negative particle count");
    NBK_CHECK_ARG(mesh != nullptr, "paint: null mesh");
    PaintGeom gm;
    int rc = make_geom(box, nmesh, x_start, x_n, gm);
    if (rc) return rc;
    if (n == 0 || x_n == 0) return NBK_OK;
    if (mass == nullptr) mass_dtype = NBK_F8;
    cudaStream_t s = (cudaStream_t)stream;
    switch (window) {
        case NBK_WINDOW_NNB: return launch_paint0<1>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, s);
        case NBK_WINDOW_CIC: return launch_paint0<2>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, s);
        case NBK_WINDOW_TSC: return launch_paint0<3>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, s);
        case NBK_WINDOW_PCS: return launch_paint0<4>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, s);
    }
    nbk_set_error("paint: unknown window %d", window);
    return NBK_ERR_ARG;
}

extern "C" int nbk_paint(const void *pos, int pos_dtype, int64_t n, const void *mass, int mass_dtype, int window,
                         double shift, const double *box, const int64_t *nmesh, int64_t x_start, int64_t x_n,
                         void *mesh, int mesh_dtype, void *stream) {
    return paint_impl(pos, pos_dtype, n, mass, mass_dtype, window, shift, box, nmesh, x_start, x_n, mesh, nullptr,
                      mesh_dtype, stream);
}

extern "C" int nbk_paint_interlaced(const void *pos, int pos_dtype, int64_t n, const void *mass, int mass_dtype,
                                    int window, const double *box, const int64_t *nmesh, int64_t x_start,
                                    int64_t x_n, void *mesh1, void *mesh2, int mesh_dtype, void *stream) {
    NBK_CHECK_ARG(mesh2 != nullptr, "paint_interlaced: null mesh2");
    return paint_impl(pos, pos_dtype, n, mass, mass_dtype, window, 0.0, box, nmesh, x_start, x_n, mesh1, mesh2,
                      mesh_dtype, stream);
}

extern "C" int nbk_cell_index(const void *pos, int pos_dtype, int64_t n, int window, double shift, const double *box,
                              const int64_t *nmesh, int32_t *cell_out, void *stream) {
    NBK_CHECK_ARG(pos_dtype == NBK_F4 || pos_dtype == NBK_F8, "cell_index: bad pos dtype %d", pos_dtype);
    PaintGeom gm;
    int rc = make_geom(box, nmesh, 0, nmesh[0], gm);
    if (rc) return rc;
    if (n == 0) return NBK_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int g = nbk_grid_for(n, 256, 8);
#define CI(SUP)                                                                                              \
    if (pos_dtype == NBK_F4) k_cell_index<SUP, float><<<g, 256, 0, s>>>((const float *)pos, n, gm, shift, cell_out); \
    else k_cell_index<SUP, double><<<g, 256, 0, s>>>((const double *)pos, n, gm, shift, cell_out);
    switch (window) {
        case NBK_WINDOW_NNB: CI(1); break;
        case NBK_WINDOW_CIC: CI(2); break;
        case NBK_WINDOW_TSC: CI(3); break;
        case NBK_WINDOW_PCS: CI(4); break;
        default: nbk_set_error("cell_index: unknown window %d", window); return NBK_ERR_ARG;
    }
#undef CI
    NBK_LAUNCHED();
    return NBK_OK;
}
