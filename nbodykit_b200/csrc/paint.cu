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
#include <stdlib.h>

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
    // in-box particles (the overwhelming majority) never pay for the 64-bit modulo
    if ((unsigned long long)i < (unsigned long long)n) return (int)i;
    if (i >= -(long long)n && i < 2ll * n) return (int)(i < 0 ? i + n : i - n);
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


// =============================================================================================
// Path "tiled": bucket particles by 16^3-cell tile, accumulate each tile in shared memory with
// native 32-bit integer atomics (ATOMS.ADD) on a 64-bit fixed-point representation, flush the
// tile once.
//
// Measured on B200 (tools/atomics_bench.cu): shared u32 ATOMS sustain ~2.5e12 op/s chip-wide in a
// CIC pattern (3e11 particles/s), the REDG path of "direct" 5e9 (random) .. 4e10 (cell-sorted)
// particles/s -- so the bucketing passes (HBM streaming) become the bound, independent of the
// particle order.  Fixed point also makes the mesh independent of the order particles arrive in:
// cell = round-to-nearest sum of w_i * 2^31/M, M = power of two >= max|mass| (exact integer
// adds, resolution 4.7e-10 M per deposit).
//
//   pass A  k_tile_count_blk   : CTA c histograms the tile ids of its contiguous particle chunk in shared memory
//                                (native ATOMS.ADD.U32; row c of blk[G][ntiles]); with `clear` it also zeroes the mesh
//   pass B  k_tile_colscan/_scan : per-tile prefix over the CTA histograms, tile offsets
//   pass C  k_tile_scatter_blk : the same CTA re-reads its chunk, evaluates the exact f8 grid coordinate once and
//                                emits a 16-byte record per particle (32-bit fixed-point fractions + in-tile cell)
//   pass D  k_tile_paint       : persistent CTAs pull tiles from a queue; region = (T + halo)^3 cells in shared
//                                memory as two u32 limbs; one TMA bulk reduce-add per z row writes it back
// (k_tile_count / k_tile_scatter are the global-atomic variants of A and C for meshes with more tiles than a shared
// histogram holds.)
// A tile owns the particles whose LEFTMOST stencil cell lies in it, so the halo is one-sided.
// =============================================================================================
#define TILE 16
#define NBK_BLK_SMEM (200 * 1024)   // largest shared-memory tile histogram (CTA-local bucketing)

// Tile-ordered particle record, 16 bytes for every position dtype: the scatter pass evaluates the grid coordinate
// in the exact f8 arithmetic once and stores, per axis, the fraction of (g + A) as a 32-bit fixed-point number
// (truncated: the weights move by < 2^-32, below the 2^-31 deposit quantum) plus the leftmost stencil cell
// relative to its tile (4 bits per axis).  The paint pass needs no floor / wrap / range logic, and the half-cell
// shifted mesh of an interlaced pair follows exactly from frac + 1/2 (carry -> next cell).
typedef uint4 TileRec;     // {frac_x, frac_y, frac_z, lx | ly << 8 | lz << 16}

struct TileGeom {
    PaintGeom gm;
    int G;            // ghost reach below the slab in x (0 when the slab is the whole mesh)
    int nt[3];        // tiles per axis
    int R;            // region edge = TILE + support - 1 (+1 when a half-cell shifted mesh is painted)
    int ntiles;
    int full;         // the slab is the whole mesh (single GPU): no ghost / ownership logic
};

// local x of a wrapped global cell relative to the slab origin, in [-G, Nx - G)
__device__ __forceinline__ int slab_local(int ix, const TileGeom &tg) {
    int lx = ix - tg.gm.x_start;
    if (lx >= tg.gm.n[0] - tg.G) lx -= tg.gm.n[0];
    if (lx < -tg.G) lx += tg.gm.n[0];
    return lx;
}

// leftmost-cell offsets of the windows: i0 = floor(g + OFF_A) + OFF_B
template <int SUP> struct WinOff;
template <> struct WinOff<1> { static constexpr float A = 0.5f; static constexpr int B = 0; };
template <> struct WinOff<2> { static constexpr float A = 0.0f; static constexpr int B = 0; };
template <> struct WinOff<3> { static constexpr float A = 0.5f; static constexpr int B = -1; };
template <> struct WinOff<4> { static constexpr float A = 0.0f; static constexpr int B = -1; };

__device__ __forceinline__ int tile_from_cells(const int *c, const TileGeom &tg) {
    if (tg.full)   // cells are in [0, n)
        return (int)((((unsigned)c[0] / TILE) * tg.nt[1] + (unsigned)c[1] / TILE) * tg.nt[2] + (unsigned)c[2] / TILE);
    int lx = slab_local(c[0], tg);
    if (lx < -tg.G || lx >= tg.gm.x_n) return -1;   // cannot touch my planes
    int tx = (lx + tg.G) / TILE, ty = c[1] / TILE, tz = c[2] / TILE;
    return (tx * tg.nt[1] + ty) * tg.nt[2] + tz;
}

// exact (f8) tile id: the arithmetic of the scatter itself
template <int SUP, typename PT>
__device__ __noinline__ int tile_of_exact(const PT *__restrict__ pos, int64_t i, const TileGeom &tg) {
    double g[3];
    if (!load_grid(pos, i, tg.gm, 0.0, g)) return -1;
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        long long i0;
        double w[SUP];
        Window<SUP>::eval(g[d], i0, w);
        c[d] = wrap(i0, tg.gm.n[d]);
    }
    return tile_from_cells(c, tg);
}

// exact leftmost cell + fixed-point fraction of particle i (the arithmetic of Window<SUP>::eval on the unshifted g).
// Slow path: any magnitude, 64-bit cell arithmetic.
template <int SUP, typename PT>
__device__ __noinline__ int make_record_slow(const PT *__restrict__ pos, int64_t i, const TileGeom &tg, TileRec &rec) {
    double g[3];
    if (!load_grid(pos, i, tg.gm, 0.0, g)) return -1;
    unsigned u[3];
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double a = g[d] + (double)WinOff<SUP>::A;
        double f = floor(a);
        u[d] = __double2uint_rz((a - f) * 4294967296.0);
        c[d] = wrap((long long)f + WinOff<SUP>::B, tg.gm.n[d]);
    }
    int lx = (slab_local(c[0], tg) + tg.G) & (TILE - 1);
    rec = make_uint4(u[0], u[1], u[2], (unsigned)lx | ((unsigned)(c[1] & (TILE - 1)) << 8) | ((unsigned)(c[2] & (TILE - 1)) << 16));
    return tile_from_cells(c, tg);
}

// Fast path for |g| < 2^31 without any float<->int conversion instruction (they issue at a fraction of the FP64
// rate): a + 1.5*2^52 holds rint(a) in its low mantissa word; floor and the truncated 32-bit fraction follow with
// FP64 adds.  Bit-identical to the slow path.
template <int SUP, typename PT>
__device__ __forceinline__ int make_record(const PT *x, const PT *__restrict__ pos, int64_t i, const TileGeom &tg,
                                           TileRec &rec) {
    const double K = 6755399441055744.0;      // 1.5 * 2^52
    unsigned u[3];
    int c[3];
    bool fast = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double a = (double)x[d] * tg.gm.scale[d];
        if (WinOff<SUP>::A != 0.f) a += (double)WinOff<SUP>::A;
        fast = fast && (fabs(a) < 2147483000.0);           // false for NaN / inf as well
        double r = a + K;
        int ri = __double2loint(r);
        double rf = r - K;                                  // rint(a)
        if (rf > a) { rf -= 1.0; ri -= 1; }                 // floor(a)
        u[d] = (unsigned)__double2loint(__dadd_rz((a - rf) * 4294967296.0, 4503599627370496.0));
        int cc = ri + WinOff<SUP>::B;                       // one period of wrap here, anything further in the slow path
        if (cc < 0) cc += tg.gm.n[d];
        else if (cc >= tg.gm.n[d]) cc -= tg.gm.n[d];
        fast = fast && ((unsigned)cc < (unsigned)tg.gm.n[d]);
        c[d] = cc;
    }
    if (!fast) return make_record_slow<SUP, PT>(pos, i, tg, rec);
    int lx = (tg.full ? c[0] : slab_local(c[0], tg) + tg.G) & (TILE - 1);
    rec = make_uint4(u[0], u[1], u[2], (unsigned)lx | ((unsigned)(c[1] & (TILE - 1)) << 8) | ((unsigned)(c[2] & (TILE - 1)) << 16));
    return tile_from_cells(c, tg);      // the tile the bucketing pass counted this particle in (both are exact)
}

// coordinates of the 4 consecutive particles i0 .. i0+3 (i0 % 4 == 0): three 16-byte loads per thread for f4
// positions when the array is 16-byte aligned (a warp then reads one contiguous 1536-byte run), scalar loads otherwise
template <typename PT>
__device__ __forceinline__ void load4(const PT *__restrict__ cpos, int j0, int nv, bool aligned, PT (&x)[4][3]) {
    // cpos = first particle of the CTA's chunk, j0 = chunk-relative index of the quad (j0 % 4 == 0), nv = valid ones
    if (aligned && nv == 4) {
        constexpr int NV = (int)(12 * sizeof(PT) / 16);
        const uint4 *v = reinterpret_cast<const uint4 *>(cpos + 3 * j0);
        uint4 r[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) r[k] = v[k];
        const PT *f = reinterpret_cast<const PT *>(r);
#pragma unroll
        for (int u = 0; u < 4; u++) { x[u][0] = f[3 * u]; x[u][1] = f[3 * u + 1]; x[u][2] = f[3 * u + 2]; }
    } else {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            bool in = u < nv;
            x[u][0] = in ? cpos[3 * (j0 + u)] : (PT)0;
            x[u][1] = in ? cpos[3 * (j0 + u) + 1] : (PT)0;
            x[u][2] = in ? cpos[3 * (j0 + u) + 2] : (PT)0;
        }
    }
}

// Tile id of particle i.  float32 positions take a float32 fast path: g32 = x*scale differs from the f8 grid
// coordinate by < 2^-22 |g|, so unless the fraction of (g32 + A) lies within 3e-7|g| of a cell boundary (or the
// particle is far outside the box) floor() agrees with the exact arithmetic; the rare rest is recomputed in f8.
// The result is therefore ALWAYS the exact leftmost cell -- count, scatter and paint passes agree.
// per-launch constants of the float32 fast path
struct FastTile {
    float sc[3];    // float32 scale N/L
    float lim[3];   // accept when |frac(g) - 0.5| < lim  (frac at least eps away from both cell boundaries)
};

static FastTile make_fast_tile(const TileGeom &tg) {      // host side: passed to the kernels by value
    FastTile f;
    for (int d = 0; d < 3; d++) {
        f.sc[d] = (float)tg.gm.scale[d];
        // |g32 - g_exact| <= 2 float32 roundings of a value below n+2 -> 3e-7 (n+2) + 1e-6 is a safe margin
        f.lim[d] = 0.5f - (3e-7f * (float)(tg.gm.n[d] + 2) + 1e-6f);
    }
    return f;
}

// Tile id of particle i.  float32 in-box positions take a float32 fast path: unless the fraction of
// (x*scale + A) lies within the rounding margin of a cell boundary, floor() agrees with the exact f8 arithmetic;
// everything else (near-boundary, outside the box, f8 positions) is recomputed in f8.  The id is therefore
// ALWAYS the exact leftmost cell's tile -- count, scatter and paint passes agree.
template <int SUP, typename PT>
__device__ __forceinline__ int tile_of(const PT *x, const PT *__restrict__ pos, int64_t i, const TileGeom &tg,
                                       const FastTile &ft) {
    if (sizeof(PT) == 4) {
        int c[3];
        bool ok = true;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float g = (float)x[d] * ft.sc[d];
            if (WinOff<SUP>::A != 0.f) g += WinOff<SUP>::A;
            float f = floorf(g);
            ok = ok && (fabsf((g - f) - 0.5f) < ft.lim[d]);
            c[d] = (int)f + WinOff<SUP>::B;
            ok = ok && ((unsigned)c[d] < (unsigned)tg.gm.n[d]);
        }
        if (ok) return tile_from_cells(c, tg);
    }
    return tile_of_exact<SUP, PT>(pos, i, tg);
}

// one atomic per distinct key per warp; returns this lane's rank within its key group and the group's base
__device__ __forceinline__ unsigned warp_claim(unsigned *counter, int key, bool active) {
    const int lane = threadIdx.x & 31;
    // fast path: the whole warp is active and in one tile (spatially coherent catalogues)
    int k0 = __shfl_sync(0xffffffffu, key, 0);
    if (__all_sync(0xffffffffu, active && key == k0)) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&counter[key], 32u);
        return __shfl_sync(0xffffffffu, base, 0) + lane;
    }
    unsigned mask = __match_any_sync(0xffffffffu, active ? key : -1 - lane);
    if (!active) return 0;
    int leader = __ffs(mask) - 1;
    unsigned rank = __popc(mask & ((1u << lane) - 1));
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&counter[key], (unsigned)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + rank;
}

template <int SUP, typename PT, typename MT>
__global__ void __launch_bounds__(256)
k_tile_count(const PT *__restrict__ pos, const MT *__restrict__ mass, int64_t n, TileGeom tg, FastTile ft,
             unsigned *__restrict__ counts, unsigned *__restrict__ absmax_bits, int *__restrict__ tile_ids) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float mx = 0.f;
    int64_t nround = ((n + stride - 1) / stride) * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        bool in = i < n;
        PT x[3] = {0, 0, 0};
        if (in) { x[0] = pos[3 * i]; x[1] = pos[3 * i + 1]; x[2] = pos[3 * i + 2]; }
        int t = in ? tile_of<SUP, PT>(x, pos, i, tg, ft) : -1;
        if (in) tile_ids[i] = t;          // the scatter pass reuses the id instead of recomputing it
        warp_claim(counts, t, t >= 0);
        if (mass && in && t >= 0) mx = fmaxf(mx, fabsf((float)mass[i]) * 1.0000001f);
    }
    if (mass) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(absmax_bits, __float_as_uint(mx));  // positive floats order as uints
    }
}

// single-CTA exclusive scan (4 counters per thread and round, warp shuffles + one shared hop); also clears the
// cursors and the tile queue head
__global__ void __launch_bounds__(1024)
k_tile_scan(const unsigned *__restrict__ counts, unsigned *__restrict__ offsets, unsigned *__restrict__ cursor,
            unsigned *__restrict__ queue, int ntiles) {
    __shared__ unsigned warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned carry = 0;
    for (int base = 0; base < ntiles; base += 4096) {
        const int i = base + threadIdx.x * 4;
        unsigned v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (i + j < ntiles) ? counts[i + j] : 0u;
        const unsigned s = v[0] + v[1] + v[2] + v[3];
        unsigned inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned x = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned n = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += n;
            }
            warp_tot[lane] = x;
        }
        __syncthreads();
        unsigned run = carry + (wid ? warp_tot[wid - 1] : 0u) + inc - s;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i + j < ntiles) { offsets[i + j] = run; cursor[i + j] = 0; }
            run += v[j];
        }
        carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[ntiles] = carry; queue[0] = 0; }
}

template <int SUP, typename PT, typename MT>
__global__ void __launch_bounds__(256)
k_tile_scatter(const PT *__restrict__ pos, const MT *__restrict__ mass, int64_t n, TileGeom tg,
               const unsigned *__restrict__ offsets, unsigned *__restrict__ cursor,
               TileRec *__restrict__ recs, MT *__restrict__ smass, const int *__restrict__ tile_ids) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t nround = ((n + stride - 1) / stride) * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        bool in = i < n;
        int t = in ? tile_ids[i] : -1;
        TileRec r = make_uint4(0, 0, 0, 0);
        if (t >= 0) {
            PT x[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
            t = make_record<SUP, PT>(x, pos, i, tg, r);
        }
        unsigned slot = warp_claim(cursor, t, t >= 0);
        if (t >= 0) {
            int64_t dst = (int64_t)offsets[t] + slot;
            recs[dst] = r;                       // one 16-byte store per particle
            if (mass) smass[dst] = mass[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-local bucketing (ntiles * 4 B fits in shared memory): CTA c owns the contiguous particle chunk
// [c*chunk, (c+1)*chunk).  Pass A histograms its chunk in shared memory (native ATOMS.ADD.U32, no global
// atomics) and stores the histogram as row c of blk[G][ntiles]; pass B turns every column into an exclusive
// prefix over c and the column totals into tile offsets; pass C reloads row c as shared cursors and scatters the
// same chunk.  Slot order inside a tile is arbitrary, the fixed-point accumulation makes the mesh independent of it.
// ---------------------------------------------------------------------------------------------
// Shared-memory counter claim.  Random catalogues: one native ATOMS per lane.  Spatially coherent catalogues put many
// lanes of a warp on the same counter, which the atomic unit serialises; when neighbouring lanes agree often, the
// warp aggregates equal keys first (one ATOMS per distinct key).  Must be called by all 32 lanes.
__device__ __forceinline__ unsigned smem_claim(unsigned *hist, int key, bool active) {
    const int lane = threadIdx.x & 31;
    int kn = __shfl_xor_sync(0xffffffffu, key, 1);
    unsigned same = __ballot_sync(0xffffffffu, active && kn == key);
    if (__popc(same) < 8) return active ? atomicAdd(&hist[key], 1u) : 0u;
    unsigned mask = __match_any_sync(0xffffffffu, active ? key : -1 - lane);
    if (!active) return 0u;
    int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&hist[key], (unsigned)__popc(mask));
    return __shfl_sync(mask, base, leader) + __popc(mask & ((1u << lane) - 1));
}

// float32 fast path of tile_of without the fallback: ok == false means "recompute exactly"
template <int SUP, typename PT>
__device__ __forceinline__ int tile_fast(const PT *x, const TileGeom &tg, const FastTile &ft, bool &ok) {
    ok = sizeof(PT) == 4;
    int c[3] = {0, 0, 0};
    if (sizeof(PT) == 4) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float g = (float)x[d] * ft.sc[d];
            if (WinOff<SUP>::A != 0.f) g += WinOff<SUP>::A;
            float f = floorf(g);
            ok = ok && (fabsf((g - f) - 0.5f) < ft.lim[d]);
            c[d] = (int)f + WinOff<SUP>::B;
            ok = ok && ((unsigned)c[d] < (unsigned)tg.gm.n[d]);
        }
    }
    return ok ? tile_from_cells(c, tg) : -1;
}

// Claim slots for the (up to) four particles of a thread's quad.  Coherent input: when every lane's quad lies in one
// tile, the quad is claimed as a unit (then usually the whole warp as one ATOMS); otherwise particle by particle.
// Returns the slot of the quad's first particle in `slot[0..3]`.  All 32 lanes must call.
__device__ __forceinline__ void quad_claim(unsigned *hist, const int (&t)[4], unsigned (&slot)[4]) {
    const bool uni = (t[0] == t[1]) && (t[1] == t[2]) && (t[2] == t[3]) && (t[0] >= 0);
    if (__all_sync(0xffffffffu, uni)) {
        const int lane = threadIdx.x & 31;
        unsigned mask = __match_any_sync(0xffffffffu, t[0]);
        int leader = __ffs(mask) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&hist[t[0]], 4u * (unsigned)__popc(mask));
        base = __shfl_sync(mask, base, leader) + 4u * (unsigned)__popc(mask & ((1u << lane) - 1));
#pragma unroll
        for (int u = 0; u < 4; u++) slot[u] = base + u;
        return;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) slot[u] = smem_claim(hist, t[u], t[u] >= 0);
}

template <int SUP, typename PT, typename MT, bool HASM>
__global__ void __launch_bounds__(1024)
k_tile_count_blk(const PT *__restrict__ pos, const MT *__restrict__ mass, int64_t n, int64_t chunk, TileGeom tg,
                 FastTile ft, unsigned *__restrict__ blk, unsigned *__restrict__ absmax_bits, uint4 *__restrict__ zero1,
                 uint4 *__restrict__ zero2, int64_t zero_n) {
    extern __shared__ __align__(16) unsigned s_hist[];
    for (int t = threadIdx.x; t < tg.ntiles; t += blockDim.x) s_hist[t] = 0;
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * chunk;
    const int64_t e = (b + chunk < n) ? b + chunk : n;
    const int cn = e > b ? (int)(e - b) : 0;           // particles of this CTA (chunk-relative 32-bit indices below)
    const PT *cpos = pos + 3 * b;
    const MT *cmass = HASM ? mass + b : nullptr;
    float mx = 0.f;
    // 4 consecutive particles per thread and round: the coordinate loads are issued before the first is consumed (the
    // pass is bound by memory latency at 32 warps / SM) and, for aligned arrays, are 16-byte vectors
    const bool aligned = (reinterpret_cast<uintptr_t>(pos) & 15) == 0;
    // hold=False: this pass also clears the mesh(es) -- zero_n 16-byte vectors each, a contiguous share per CTA, three
    // stores per round riding in the shadow of the (latency-bound) particle loads; the tile pass runs later in-stream
    const int64_t zper = zero1 ? (zero_n + gridDim.x - 1) / gridDim.x : 0;
    const int64_t zbeg = (int64_t)blockIdx.x * zper;
    const int zcnt = zero1 ? (int)((zbeg + zper < zero_n ? zbeg + zper : zero_n) - zbeg) : 0;   // may be <= 0
    uint4 *z1 = zero1 ? zero1 + zbeg : nullptr, *z2 = zero2 ? zero2 + zbeg : nullptr;
    int zi = threadIdx.x;
    const uint4 zz = make_uint4(0, 0, 0, 0);
    for (int base = 0; base < cn; base += 4 * (int)blockDim.x) {     // uniform trip count (warp collectives)
        const int j0 = base + 4 * (int)threadIdx.x;
        const int nv = min(4, max(0, cn - j0));                       // valid particles of my quad
        PT x[4][3];
        MT mv[4];
        load4(cpos, j0, nv, aligned, x);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (zi < zcnt) { z1[zi] = zz; if (z2) z2[zi] = zz; }
            zi += blockDim.x;
        }
        if (HASM) {
#pragma unroll
            for (int u = 0; u < 4; u++) mv[u] = (u < nv) ? cmass[j0 + u] : (MT)0;
        }
        int t[4];
        bool redo = false;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            bool ok;
            t[u] = tile_fast<SUP, PT>(x[u], tg, ft, ok);
            redo = redo || (!ok && u < nv);
        }
        if (redo) {                                  // rare: near a cell boundary / outside the box / f8 positions
#pragma unroll
            for (int u = 0; u < 4; u++) {
                bool ok;
                tile_fast<SUP, PT>(x[u], tg, ft, ok);
                if (!ok && u < nv) t[u] = tile_of_exact<SUP, PT>(cpos, j0 + u, tg);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) if (u >= nv) t[u] = -1;
        unsigned slot[4];
        quad_claim(s_hist, t, slot);
        if (HASM) {
#pragma unroll
            for (int u = 0; u < 4; u++) if (t[u] >= 0) mx = fmaxf(mx, fabsf((float)mv[u]) * 1.0000001f);
        }
    }
    for (; zi < zcnt; zi += blockDim.x) { z1[zi] = zz; if (z2) z2[zi] = zz; }
    if (HASM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(absmax_bits, __float_as_uint(mx));
    }
    __syncthreads();
    unsigned *row = blk + (size_t)blockIdx.x * tg.ntiles;
    for (int t = threadIdx.x; t < tg.ntiles; t += blockDim.x) row[t] = s_hist[t];
}

// column pass: blk[c][t] <- sum_{c' < c} blk[c'][t];  counts[t] <- column total.  16 independent loads in flight
// per thread (the column is a chain of G dependent adds, not of G dependent memory round trips)
__global__ void __launch_bounds__(128)
k_tile_colscan(unsigned *__restrict__ blk, unsigned *__restrict__ counts, int ntiles, int G) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    unsigned run = 0;
    for (int c0 = 0; c0 < G; c0 += 16) {
        unsigned v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = (c0 + j < G) ? blk[(size_t)(c0 + j) * ntiles + t] : 0u;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (c0 + j < G) blk[(size_t)(c0 + j) * ntiles + t] = run;
            run += v[j];
        }
    }
    counts[t] = run;
}

template <int SUP, typename PT, typename MT, bool HASM>
__global__ void __launch_bounds__(1024)
k_tile_scatter_blk(const PT *__restrict__ pos, const MT *__restrict__ mass, int64_t n, int64_t chunk, TileGeom tg,
                   const unsigned *__restrict__ offsets, const unsigned *__restrict__ blk,
                   TileRec *__restrict__ recs, MT *__restrict__ smass) {
    extern __shared__ __align__(16) unsigned s_cur[];
    const int ntiles = tg.ntiles;
    const unsigned *row = blk + (size_t)blockIdx.x * ntiles;
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) s_cur[t] = offsets[t] + row[t];
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * chunk;
    const int64_t e = (b + chunk < n) ? b + chunk : n;
    const int cn = e > b ? (int)(e - b) : 0;
    const PT *cpos = pos + 3 * b;
    const MT *cmass = HASM ? mass + b : nullptr;
    const bool aligned = (reinterpret_cast<uintptr_t>(pos) & 15) == 0;
    for (int base = 0; base < cn; base += 4 * (int)blockDim.x) {     // see k_tile_count_blk
        const int j0 = base + 4 * (int)threadIdx.x;
        const int nv = min(4, max(0, cn - j0));
        PT x[4][3];
        MT mv[4];
        load4(cpos, j0, nv, aligned, x);
        if (HASM) {
#pragma unroll
            for (int u = 0; u < 4; u++) mv[u] = (u < nv) ? cmass[j0 + u] : (MT)0;
        }
        TileRec r[4];
        int t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            r[u] = make_uint4(0, 0, 0, 0);
            t[u] = (u < nv) ? make_record<SUP, PT>(x[u], cpos, j0 + u, tg, r[u]) : -1;
        }
        unsigned slot[4];
        quad_claim(s_cur, t, slot);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (t[u] >= 0) {
                recs[slot[u]] = r[u];
                if (HASM) smass[slot[u]] = mv[u];
            }
        }
    }
}

// 64-bit fixed-point cell = two 32-bit limbs kept in SEPARATE arrays lo[NC] | hi[NC] (so the low-limb atomics, which
// are nearly all of them, spread over all 32 banks).  Deposits are native ATOMS.ADD on the low limb; the carry (seen
// in the returned old value) goes to the high limb.
__device__ __forceinline__ void fixed_add(unsigned *lo, unsigned *hi, int cell, long long q) {
    unsigned ql = (unsigned)q, qh = (unsigned)(q >> 32);
    unsigned old = atomicAdd(&lo[cell], ql);
    unsigned h = qh + (unsigned)(old + ql < old);
    if (h) atomicAdd(&hi[cell], h);
}
// non-negative deposit: 0 <= q <= 2^31 fits the low limb
__device__ __forceinline__ void fixed_add_pos(unsigned *lo, unsigned *hi, int cell, unsigned ql) {
    unsigned old = atomicAdd(&lo[cell], ql);
    if (old + ql < old) atomicAdd(&hi[cell], 1u);
}

__device__ __forceinline__ void tma_reduce_add(double *gdst, const void *ssrc, unsigned bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 :: "l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_reduce_add(float *gdst, const void *ssrc, unsigned bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 :: "l"(gdst), "r"(sa), "r"(bytes) : "memory");
}

// window weights from the stencil-relative offset d = g - i0 (d in [0,1) CIC, [0.5,1.5) TSC, [1,2) PCS)
template <int SUP> struct WinD;
template <> struct WinD<1> { static constexpr double DMIN = 0.0;
    __device__ static __forceinline__ void eval(double, double *w) { w[0] = 1.0; } };
template <> struct WinD<2> { static constexpr double DMIN = 0.0;
    __device__ static __forceinline__ void eval(double d, double *w) { w[0] = 1.0 - d; w[1] = d; } };
template <> struct WinD<3> { static constexpr double DMIN = 0.5;
    __device__ static __forceinline__ void eval(double d, double *w) {
#pragma unroll
        for (int r = 0; r < 3; r++) w[r] = tsc_kernel(d - (double)r);
    } };
template <> struct WinD<4> { static constexpr double DMIN = 1.0;
    __device__ static __forceinline__ void eval(double d, double *w) {
#pragma unroll
        for (int r = 0; r < 4; r++) w[r] = pcs_kernel(d - (double)r);
    } };

// FLUSH 0: per-cell read-add-store (exclusive cells) / REDG (halo)   1: TMA bulk reduce-add, one row per op
template <int SUP, typename MT, typename FT, bool SHIFTED, int FLUSH>
__global__ void __launch_bounds__(256)
k_tile_paint(const TileRec *__restrict__ recs, const MT *__restrict__ smass, TileGeom tg,
             const unsigned *__restrict__ offsets, unsigned *__restrict__ queue,
             const unsigned *__restrict__ absmax_bits, FT *__restrict__ mesh) {
    extern __shared__ __align__(16) unsigned s_acc[];
    constexpr int R = TILE + SUP - 1 + (SHIFTED ? 1 : 0);   // == tg.R
    constexpr int RP = (R + 3) & ~3;                         // row pitch in cells: rows start 16-byte aligned
    constexpr int NC = R * R * RP;                           // multiple of 4
    unsigned *s_lo = s_acc, *s_hi = s_acc + NC;
    __shared__ int s_tile;
    // scale 2^31 / M, M = power of two >= max |mass|
    double M = 1.0;
    if (smass) {
        float mx = __uint_as_float(*absmax_bits);
        int e;
        frexpf(mx, &e);
        M = mx > 0.f ? ldexp(1.0, e) : 1.0;
    }
    const double S = 2147483648.0 / M, invS = M / 2147483648.0;
    constexpr int H = R - TILE;   // cells with a local coordinate < H may also be written by the preceding tile
    constexpr int PER = (NC + 255) / 256;
    for (;;) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(queue, 1u);
        __syncthreads();
        int t = s_tile;
        if (t >= tg.ntiles) break;
        unsigned b = offsets[t], e = offsets[t + 1];
        if (b == e) { __syncthreads(); continue; }
        for (int i = threadIdx.x; i < NC / 2; i += blockDim.x) reinterpret_cast<uint4 *>(s_acc)[i] = make_uint4(0, 0, 0, 0);
        int tz = t % tg.nt[2], ty = (t / tg.nt[2]) % tg.nt[1], tx = t / (tg.nt[2] * tg.nt[1]);
        int o[3] = {tx * TILE - tg.G, ty * TILE, tz * TILE};   // region origin (x: slab-local)
        __syncthreads();
        // the record (and mass) of the next round is requested before this round's deposits are issued
        unsigned p = b + threadIdx.x;
        TileRec rn = make_uint4(0, 0, 0, 0);
        MT mn = (MT)1;
        if (p < e) { rn = recs[p]; if (smass) mn = smass[p]; }
        for (; p < e; p += blockDim.x) {
            const TileRec r = rn;
            const MT mcur = mn;
            if (p + blockDim.x < e) { rn = recs[p + blockDim.x]; if (smass) mn = smass[p + blockDim.x]; }
            unsigned u[3] = {r.x, r.y, r.z};
            int l[3] = {(int)(r.w & 255u), (int)((r.w >> 8) & 255u), (int)((r.w >> 16) & 255u)};
            double w[3][SUP];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                if (SHIFTED) {                       // frac(g + A + 1/2): carry moves the stencil one cell up
                    unsigned v = u[d] + 0x80000000u;
                    l[d] += (v < u[d]) ? 1 : 0;
                    u[d] = v;
                }
                // u * 2^-32 via the 2^52 mantissa trick (no I2F)
                double fr = (__hiloint2double(0x43300000, (int)u[d]) - 4503599627370496.0) * 2.3283064365386963e-10;
                WinD<SUP>::eval(WinD<SUP>::DMIN != 0.0 ? WinD<SUP>::DMIN + fr : fr, w[d]);
            }
            const int base0 = (l[0] * R + l[1]) * RP + l[2];
            const double m = smass ? (double)mcur : 1.0;
            const double mS = m * S;
            double wz[SUP];
#pragma unroll
            for (int rz = 0; rz < SUP; rz++) wz[rz] = w[2][rz] * mS;
            if (m >= 0.0) {
                // all deposits are in [0, 2^31]: round-to-nearest integer = low word of fma(wxy, wz, 2^52)
#pragma unroll
                for (int rx = 0; rx < SUP; rx++)
#pragma unroll
                    for (int ry = 0; ry < SUP; ry++) {
                        double wxy = w[0][rx] * w[1][ry];
#pragma unroll
                        for (int rz = 0; rz < SUP; rz++) {
                            unsigned q = (unsigned)__double2loint(__fma_rn(wxy, wz[rz], 4503599627370496.0));
                            fixed_add_pos(s_lo, s_hi, base0 + (rx * R + ry) * RP + rz, q);
                        }
                    }
            } else {
#pragma unroll
                for (int rx = 0; rx < SUP; rx++)
#pragma unroll
                    for (int ry = 0; ry < SUP; ry++) {
                        double wxy = w[0][rx] * w[1][ry];
#pragma unroll
                        for (int rz = 0; rz < SUP; rz++)
                            fixed_add(s_lo, s_hi, base0 + (rx * R + ry) * RP + rz, __double2ll_rn(wxy * wz[rz]));
                    }
            }
        }
        __syncthreads();
        if (FLUSH == 1) {
            // ---- convert the region to mesh dtype (through registers: the values overlay the limb arrays), then one
            // TMA bulk reduce-add per z row
            FT v[PER];
#pragma unroll
            for (int k = 0; k < PER; k++) {
                int i = threadIdx.x + k * 256;
                v[k] = (FT)0;
                if (i < NC) {
                    unsigned lo = s_lo[i], hi = s_hi[i];
                    v[k] = (FT)((double)(long long)(((unsigned long long)hi << 32) | lo) * invS);
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < PER; k++) {
                int i = threadIdx.x + k * 256;
                if (i < NC) reinterpret_cast<FT *>(s_acc)[i] = v[k];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const int nz1 = min(RP, tg.gm.n[2] - o[2]);      // cells up to the end of the z row; the rest wraps to z = 0
            for (int row = threadIdx.x; row < R * R; row += blockDim.x) {
                int cx = row / R, cy = row - cx * R;
                int gx = o[0] + cx + tg.gm.x_start;
                if (gx < 0) gx += tg.gm.n[0];
                if (gx >= tg.gm.n[0]) gx -= tg.gm.n[0];
                int ix = gx - tg.gm.x_start;
                if (ix < 0 || ix >= tg.gm.x_n) continue;     // not my plane (ghost semantics)
                int iy = o[1] + cy; if (iy >= tg.gm.n[1]) iy -= tg.gm.n[1];
                FT *grow = mesh + ((int64_t)ix * tg.gm.n[1] + iy) * tg.gm.n[2];
                const FT *srow = reinterpret_cast<const FT *>(s_acc) + (size_t)row * RP;
                tma_reduce_add(grow + o[2], srow, (unsigned)(nz1 * sizeof(FT)));
                if (nz1 < RP) tma_reduce_add(grow, srow + nz1, (unsigned)((RP - nz1) * sizeof(FT)));
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be reused
        } else {
            // ---- per-cell flush: lanes run along z
            for (int i = threadIdx.x; i < NC; i += blockDim.x) {
                unsigned lo = s_lo[i], hi = s_hi[i];
                if ((lo | hi) == 0u) continue;
                int cz = i % RP, cy = (i / RP) % R, cx = i / (RP * R);
                int gx = o[0] + cx + tg.gm.x_start;
                if (gx < 0) gx += tg.gm.n[0];
                if (gx >= tg.gm.n[0]) gx -= tg.gm.n[0];
                int ix = gx - tg.gm.x_start;
                if (ix < 0 || ix >= tg.gm.x_n) continue;
                int iy = o[1] + cy; if (iy >= tg.gm.n[1]) iy -= tg.gm.n[1];
                int iz = o[2] + cz; if (iz >= tg.gm.n[2]) iz -= tg.gm.n[2];
                double val = (double)(long long)(((unsigned long long)hi << 32) | lo) * invS;
                FT *dst = mesh + ((int64_t)ix * tg.gm.n[1] + iy) * tg.gm.n[2] + iz;
                bool exclusive = cx >= H && cx < TILE && cy >= H && cy < TILE && cz >= H && cz < TILE;
                if (exclusive) *dst = (FT)((double)*dst + val);
                else atomicAdd(dst, (FT)val);
            }
        }
        __syncthreads();
    }
}

static int make_tile_geom(const PaintGeom &gm, int sup, bool shifted, TileGeom &tg) {
    tg.gm = gm;
    tg.G = (gm.x_n == gm.n[0]) ? 0 : sup + 1;
    tg.full = (gm.x_n == gm.n[0] && gm.x_start == 0) ? 1 : 0;
    tg.R = TILE + sup - 1 + (shifted ? 1 : 0);
    tg.nt[0] = (gm.x_n + tg.G + TILE - 1) / TILE;
    tg.nt[1] = (gm.n[1] + TILE - 1) / TILE;
    tg.nt[2] = (gm.n[2] + TILE - 1) / TILE;
    int64_t nt = (int64_t)tg.nt[0] * tg.nt[1] * tg.nt[2];
    NBK_CHECK_ARG(nt < (1ll << 30), "paint_tiled: too many tiles");
    tg.ntiles = (int)nt;
    return NBK_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int nbk_paint_tiled_supported(const int64_t *nmesh, int64_t x_n, int window) {
    int sup = window;
    if (sup < 1 || sup > 4) return 0;
    int R = TILE + sup;
    // the region must not wrap onto itself; tiles must not straddle the periodic seam (else a wrapped halo
    // would land in another tile's exclusively-owned cells); the slab must hold the ghost reach
    if (nmesh[1] < 2 * TILE || nmesh[2] < 2 * TILE || nmesh[0] < 2 * TILE || R > 2 * TILE) return 0;
    if (nmesh[1] % TILE || nmesh[2] % TILE) return 0;
    if (x_n == nmesh[0]) { if (nmesh[0] % TILE) return 0; }
    else if (x_n < sup + 1) return 0;
    return 1;
}

extern "C" int64_t nbk_paint_tiled_workspace(int64_t n, int pos_dtype, int mass_dtype, const int64_t *nmesh,
                                             int64_t x_n) {
    int64_t G = 8;
    int64_t nt = ((x_n + G + TILE - 1) / TILE) * ((nmesh[1] + TILE - 1) / TILE) * ((nmesh[2] + TILE - 1) / TILE);
    size_t bytes = 256;                                  // header: queue, absmax
    bytes += 3 * align256(sizeof(unsigned) * (nt + 1));  // counts, offsets, cursor
    (void)pos_dtype;
    bytes += align256((size_t)n * sizeof(TileRec));                      // 16-byte records
    bytes += align256((size_t)n * sizeof(int));                          // tile id per particle
    int64_t nh = nt < NBK_BLK_SMEM / 4 ? nt : NBK_BLK_SMEM / 4;
    bytes += align256(sizeof(unsigned) * (size_t)nh * NBK_SM_COUNT);    // per-CTA tile histograms (CTA-local bucketing)
    if (mass_dtype) bytes += align256((size_t)n * (mass_dtype == NBK_F4 ? 4 : 8));
    return (int64_t)bytes;
}

template <int SUP, typename PT, typename MT, typename FT>
static int run_tiled(const void *pos, const void *mass, int64_t n, const PaintGeom &gm, double shift, void *mesh,
                     void *mesh2, void *work, bool clear, cudaStream_t s) {
    TileGeom tg;
    bool shifted = (mesh2 != nullptr) || shift != 0.0;
    int rc = make_tile_geom(gm, SUP, shifted, tg);
    if (rc) return rc;
    const FastTile ft = make_fast_tile(tg);
    char *w = (char *)work;
    unsigned *queue = (unsigned *)w;
    unsigned *absmax = queue + 1;
    w += 256;
    size_t tb = align256(sizeof(unsigned) * (tg.ntiles + 1));
    unsigned *counts = (unsigned *)w; w += tb;
    unsigned *offsets = (unsigned *)w; w += tb;
    unsigned *cursor = (unsigned *)w; w += tb;
    TileRec *spos = (TileRec *)w; w += align256((size_t)n * sizeof(TileRec));
    int *tile_ids = (int *)w; w += align256((size_t)n * sizeof(int));
    MT *smass = mass ? (MT *)w : nullptr;
    if (mass) w += align256((size_t)n * sizeof(MT));
    unsigned *blk = (unsigned *)w;
    static int blk_mode = -1;
    if (blk_mode < 0) {
        const char *e = getenv("NBK_PAINT_BUCKET");
        blk_mode = (e && e[0] == 'g') ? 0 : 1;          // NBK_PAINT_BUCKET=global forces the global-atomic passes
    }
    const size_t hist_bytes = sizeof(unsigned) * (size_t)tg.ntiles;
    const size_t mesh_bytes = (size_t)gm.x_n * gm.n[1] * gm.n[2] * sizeof(FT);   // multiple of 16 (tiled meshes)
    const bool use_blk = blk_mode && hist_bytes <= NBK_BLK_SMEM && n >= 4 * (int64_t)tg.ntiles;
    if (use_blk) {
        const int G = NBK_SM_COUNT;
        const int64_t chunk = (((n + G - 1) / G) + 3) & ~(int64_t)3;   // multiple of 4: threads own aligned quads
        NBK_CUDA(cudaMemsetAsync(work, 0, 256, s));        // header
#define LAUNCH_BLK(HM)                                                                                               \
        do {                                                                                                         \
            NBK_CUDA(cudaFuncSetAttribute(k_tile_count_blk<SUP, PT, MT, HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)hist_bytes));                                                         \
            k_tile_count_blk<SUP, PT, MT, HM><<<G, 1024, hist_bytes, s>>>(                                             \
                (const PT *)pos, (const MT *)mass, n, chunk, tg, ft, blk, absmax, clear ? (uint4 *)mesh : nullptr,   \
                clear ? (uint4 *)mesh2 : nullptr, (int64_t)(mesh_bytes / 16));                                       \
            NBK_LAUNCHED();                                                                                          \
            k_tile_colscan<<<(tg.ntiles + 127) / 128, 128, 0, s>>>(blk, counts, tg.ntiles, G);                        \
            NBK_LAUNCHED();                                                                                          \
            k_tile_scan<<<1, 1024, 0, s>>>(counts, offsets, cursor, queue, tg.ntiles);                                \
            NBK_LAUNCHED();                                                                                          \
            NBK_CUDA(cudaFuncSetAttribute(k_tile_scatter_blk<SUP, PT, MT, HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)hist_bytes));                                                         \
            k_tile_scatter_blk<SUP, PT, MT, HM><<<G, 1024, hist_bytes, s>>>((const PT *)pos, (const MT *)mass, n, chunk, tg, \
                                                                          offsets, blk, spos, smass);                \
            NBK_LAUNCHED();                                                                                          \
        } while (0)
        if (mass) LAUNCH_BLK(true); else LAUNCH_BLK(false);
#undef LAUNCH_BLK
    } else {
        if (clear) {
            NBK_CUDA(cudaMemsetAsync(mesh, 0, mesh_bytes, s));
            if (mesh2) NBK_CUDA(cudaMemsetAsync(mesh2, 0, mesh_bytes, s));
        }
        NBK_CUDA(cudaMemsetAsync(work, 0, 256 + tb, s));   // header + counts
        int g = nbk_grid_for(n, 256, 8);
        k_tile_count<SUP, PT, MT><<<g, 256, 0, s>>>((const PT *)pos, (const MT *)mass, n, tg, ft, counts, absmax, tile_ids);
        NBK_LAUNCHED();
        k_tile_scan<<<1, 1024, 0, s>>>(counts, offsets, cursor, queue, tg.ntiles);
        NBK_LAUNCHED();
        k_tile_scatter<SUP, PT, MT><<<g, 256, 0, s>>>((const PT *)pos, (const MT *)mass, n, tg, offsets, cursor, spos, smass,
                                                      tile_ids);
        NBK_LAUNCHED();
    }
    // tile write-back: TMA bulk reduce-add rows (default) or per-cell stores/REDG (NBK_PAINT_FLUSH=st)
    static int flush_mode = -1;
    if (flush_mode < 0) {
        const char *e = getenv("NBK_PAINT_FLUSH");
        flush_mode = (e && e[0] == 's') ? 0 : 1;
    }
    // the region edge depends on the mesh being painted (one more cell for the half-cell shifted one); tile ids do not
#define LAUNCH_TP(SH, FL, MESHP)                                                                                       \
    do {                                                                                                              \
        const int Rr = TILE + SUP - 1 + ((SH) ? 1 : 0), RPr = (Rr + 3) & ~3;                                          \
        size_t smem = (size_t)2 * Rr * Rr * RPr * sizeof(unsigned);                                                   \
        int per_sm = (int)((220 * 1024) / (smem + 2048));                                                             \
        if (per_sm > 6) per_sm = 6;                                                                                   \
        if (per_sm < 1) per_sm = 1;                                                                                   \
        int grid = NBK_SM_COUNT * per_sm;                                                                             \
        if (grid > tg.ntiles) grid = tg.ntiles;                                                                       \
        NBK_CUDA(cudaFuncSetAttribute(k_tile_paint<SUP, MT, FT, SH, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)smem));                                                                    \
        k_tile_paint<SUP, MT, FT, SH, FL><<<grid, 256, smem, s>>>(spos, smass, tg, offsets, queue, absmax,             \
                                                                  (FT *)(MESHP));                                     \
        NBK_LAUNCHED();                                                                                               \
    } while (0)
    if (shift != 0.0) { if (flush_mode) LAUNCH_TP(true, 1, mesh); else LAUNCH_TP(true, 0, mesh); }
    else { if (flush_mode) LAUNCH_TP(false, 1, mesh); else LAUNCH_TP(false, 0, mesh); }
    if (mesh2) {
        NBK_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned), s));
        if (flush_mode) LAUNCH_TP(true, 1, mesh2); else LAUNCH_TP(true, 0, mesh2);
    }
#undef LAUNCH_TP
    return NBK_OK;
}

template <int SUP, typename PT, typename MT>
static int run_tiled1(const void *pos, const void *mass, int64_t n, const PaintGeom &gm, double shift, void *mesh,
                      void *mesh2, int mesh_dtype, void *work, bool clear, cudaStream_t s) {
    if (mesh_dtype == NBK_F4) return run_tiled<SUP, PT, MT, float>(pos, mass, n, gm, shift, mesh, mesh2, work, clear, s);
    return run_tiled<SUP, PT, MT, double>(pos, mass, n, gm, shift, mesh, mesh2, work, clear, s);
}

template <int SUP>
static int run_tiled0(const void *pos, int pos_dtype, const void *mass, int mass_dtype, int64_t n, const PaintGeom &gm,
                      double shift, void *mesh, void *mesh2, int mesh_dtype, void *work, bool clear, cudaStream_t s) {
    bool pf4 = pos_dtype == NBK_F4, mf4 = (mass_dtype == NBK_F4);
    if (pf4 && mf4) return run_tiled1<SUP, float, float>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear, s);
    if (pf4) return run_tiled1<SUP, float, double>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear, s);
    if (mf4) return run_tiled1<SUP, double, float>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear, s);
    return run_tiled1<SUP, double, double>(pos, mass, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear, s);
}

extern "C" int nbk_paint_tiled(const void *pos, int pos_dtype, int64_t n, const void *mass, int mass_dtype, int window,
                               double shift, const double *box, const int64_t *nmesh, int64_t x_start, int64_t x_n,
                               void *mesh, void *mesh2, int mesh_dtype, void *work, int64_t work_bytes, int clear,
                               void *stream) {
    NBK_CHECK_ARG(pos_dtype == NBK_F4 || pos_dtype == NBK_F8, "paint_tiled: bad pos dtype %d", pos_dtype);
    NBK_CHECK_ARG(mesh_dtype == NBK_F4 || mesh_dtype == NBK_F8, "paint_tiled: bad mesh dtype %d", mesh_dtype);
    NBK_CHECK_ARG(mass == nullptr || mass_dtype == NBK_F4 || mass_dtype == NBK_F8, "paint_tiled: bad mass dtype %d", mass_dtype);
    NBK_CHECK_ARG(n >= 0 && n < (1ll << 32) - 1024, "paint_tiled: particle count %lld out of range", (long long)n);
    NBK_CHECK_ARG(mesh != nullptr && work != nullptr, "paint_tiled: null mesh / workspace");
    NBK_CHECK_ARG(mesh2 == nullptr || shift == 0.0, "paint_tiled: the interlaced pair is painted with shifts (0, 0.5)");
    NBK_CHECK_ARG(shift == 0.0 || shift == 0.5, "paint_tiled: shift must be 0 or 0.5 cells");
    NBK_CHECK_ARG(nbk_paint_tiled_supported(nmesh, x_n, window), "paint_tiled: mesh too small for the tiled path (use nbk_paint)");
    PaintGeom gm;
    int rc = make_geom(box, nmesh, x_start, x_n, gm);
    if (rc) return rc;
    if (x_n == 0) return NBK_OK;
    if (n == 0) {
        if (clear) {
            size_t mb = (size_t)x_n * nmesh[1] * nmesh[2] * (mesh_dtype == NBK_F4 ? 4 : 8);
            NBK_CUDA(cudaMemsetAsync(mesh, 0, mb, (cudaStream_t)stream));
            if (mesh2) NBK_CUDA(cudaMemsetAsync(mesh2, 0, mb, (cudaStream_t)stream));
        }
        return NBK_OK;
    }
    int md = mass ? mass_dtype : 0;
    NBK_CHECK_ARG(work_bytes >= nbk_paint_tiled_workspace(n, pos_dtype, md, nmesh, x_n), "paint_tiled: workspace too small");
    if (mass == nullptr) mass_dtype = NBK_F8;
    cudaStream_t s = (cudaStream_t)stream;
    switch (window) {
        case NBK_WINDOW_NNB: return run_tiled0<1>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear != 0, s);
        case NBK_WINDOW_CIC: return run_tiled0<2>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear != 0, s);
        case NBK_WINDOW_TSC: return run_tiled0<3>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear != 0, s);
        case NBK_WINDOW_PCS: return run_tiled0<4>(pos, pos_dtype, mass, mass_dtype, n, gm, shift, mesh, mesh2, mesh_dtype, work, clear != 0, s);
    }
    nbk_set_error("paint_tiled: unknown window %d", window);
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

// =============================================================================================
// readout (gather): value_p = sum over the stencil of W * mesh[cell] -- pmesh `RealField.readout`, the transpose of the
// scatter above (algorithms/fftrecon.py:239-244 reads the displacement field at the particle positions).  Same grid
// coordinate, window and slab semantics as `scatter`: planes outside [x_start, x_start + x_n) contribute nothing,
// so with x slabs every rank produces the partial sum of its own planes.  One particle per thread, support^3 cached
// loads, f8 accumulation in the order (x, y, z) of the stencil.
// =============================================================================================
template <int SUP, typename PT, typename FT, typename OT>
__global__ void __launch_bounds__(256)
k_readout(const PT *__restrict__ pos, int64_t n, PaintGeom gm, double shift, const FT *__restrict__ mesh,
          OT *__restrict__ out, int accumulate) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double g[3];
        double acc = 0.0;
        if (load_grid(pos, i, gm, shift, g)) {
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
                if (ix < 0 || ix >= gm.x_n) continue;
#pragma unroll
                for (int ry = 0; ry < SUP; ry++) {
                    double wxy = w[0][rx] * w[1][ry];
                    int64_t row = ((int64_t)ix * gm.n[1] + iy[ry]) * gm.n[2];
#pragma unroll
                    for (int rz = 0; rz < SUP; rz++) acc += (wxy * w[2][rz]) * (double)mesh[row + iz[rz]];
                }
            }
        }
        out[i] = accumulate ? (OT)((double)out[i] + acc) : (OT)acc;
    }
}

template <int SUP>
static int launch_readout(const void *pos, int pos_dtype, int64_t n, const PaintGeom &gm, double shift, const void *mesh,
                          int mesh_dtype, void *out, int out_dtype, int accumulate, cudaStream_t s) {
    int g = nbk_grid_for(n, 256, 8);
#define RO(PT, FT, OT) k_readout<SUP, PT, FT, OT><<<g, 256, 0, s>>>((const PT *)pos, n, gm, shift, (const FT *)mesh, (OT *)out, accumulate)
    const bool pf = pos_dtype == NBK_F4, ff = mesh_dtype == NBK_F4, of = out_dtype == NBK_F4;
    if (pf) { if (ff) { if (of) RO(float, float, float); else RO(float, float, double); }
              else { if (of) RO(float, double, float); else RO(float, double, double); } }
    else { if (ff) { if (of) RO(double, float, float); else RO(double, float, double); }
           else { if (of) RO(double, double, float); else RO(double, double, double); } }
#undef RO
    NBK_LAUNCHED();
    return NBK_OK;
}

extern "C" int nbk_readout(const void *mesh, int mesh_dtype, const void *pos, int pos_dtype, int64_t n, int window,
                           double shift, const double *box, const int64_t *nmesh, int64_t x_start, int64_t x_n,
                           void *out, int out_dtype, int accumulate, void *stream) {
    NBK_CHECK_ARG(pos_dtype == NBK_F4 || pos_dtype == NBK_F8, "readout: bad pos dtype %d", pos_dtype);
    NBK_CHECK_ARG(mesh_dtype == NBK_F4 || mesh_dtype == NBK_F8, "readout: bad mesh dtype %d", mesh_dtype);
    NBK_CHECK_ARG(out_dtype == NBK_F4 || out_dtype == NBK_F8, "readout: bad out dtype %d", out_dtype);
    NBK_CHECK_ARG(n >= 0 && mesh != nullptr && (n == 0 || out != nullptr), "readout: null mesh / out or negative count");
    PaintGeom gm;
    int rc = make_geom(box, nmesh, x_start, x_n, gm);
    if (rc) return rc;
    if (n == 0) return NBK_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (window) {
        case NBK_WINDOW_NNB: return launch_readout<1>(pos, pos_dtype, n, gm, shift, mesh, mesh_dtype, out, out_dtype, accumulate, s);
        case NBK_WINDOW_CIC: return launch_readout<2>(pos, pos_dtype, n, gm, shift, mesh, mesh_dtype, out, out_dtype, accumulate, s);
        case NBK_WINDOW_TSC: return launch_readout<3>(pos, pos_dtype, n, gm, shift, mesh, mesh_dtype, out, out_dtype, accumulate, s);
        case NBK_WINDOW_PCS: return launch_readout<4>(pos, pos_dtype, n, gm, shift, mesh, mesh_dtype, out, out_dtype, accumulate, s);
    }
    nbk_set_error("readout: unknown window %d", window);
    return NBK_ERR_ARG;
}
