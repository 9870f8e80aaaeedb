// Particle routing for the x-slab decomposition (pmesh `pm.decompose` + `Layout.exchange`, called from
// nbodykit/source/mesh/catalog.py:271-284): which OTHER ranks own a plane within `smoothing` cells of a particle.
// A rank paints ALL of its local particles itself (the scatter kernels drop stencil points outside the slab), so
// only copies for remote slabs travel: for spatially coherent catalogues that is the ghost layer alone.
//   nbk_route_count   : compact list of (particle index, destination bitmask) of the particles that have to travel
//                       (P <= 32) + per-destination counts; particles whose reach stays inside the own slab -- all
//                       but the ghost layer of a slab-local catalogue -- are rejected in float32 and write nothing
//   nbk_route_scatter : copy (pos[, mass]) of the listed particles into per-destination segments of the send buffer
#include "common.cuh"

struct RouteGeom {
    double scale;     // Nx / Lx
    double smoothing;
    int Nx, x_n, P, rank;
};

__device__ __forceinline__ unsigned route_mask(double p, const RouteGeom &g) {
    double gx = p * g.scale;
    if (!isfinite(gx)) return 0u;
    long long lo = (long long)floor(gx - g.smoothing), hi = (long long)floor(gx + g.smoothing);
    unsigned m = 0u;
    if (hi - lo < g.x_n && lo >= -(long long)g.Nx && hi < 2ll * g.Nx) {
        // the reach is narrower than a slab: only the slabs of the two end cells can be touched
        int wl = (int)lo, wh = (int)hi;
        wl = wl < 0 ? wl + g.Nx : (wl >= g.Nx ? wl - g.Nx : wl);
        wh = wh < 0 ? wh + g.Nx : (wh >= g.Nx ? wh - g.Nx : wh);
        m = (1u << (unsigned)(wl / g.x_n)) | (1u << (unsigned)(wh / g.x_n));
    } else {
        for (long long c = lo; c <= hi; c++) {
            long long w = c % g.Nx;
            if (w < 0) w += g.Nx;
            m |= 1u << (unsigned)(w / g.x_n);
        }
    }
    return m & ~(1u << g.rank);
}

template <typename PT>
__global__ void __launch_bounds__(256)
k_route_count(const PT *__restrict__ pos, int64_t n, RouteGeom g, unsigned long long *__restrict__ counts,
              unsigned long long *__restrict__ list) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t nround = ((n + stride - 1) / stride) * stride;
    const int lane = threadIdx.x & 31;
    // float32 rejection: [gx - s, gx + s] strictly inside the own slab (with a margin covering the float32 rounding of
    // gx) cannot touch a remote plane -- no f8 arithmetic, no 64-bit conversions, nothing written
    const float sc = (float)g.scale, sm = (float)g.smoothing;
    const float margin = 1e-3f + 4e-7f * (float)g.Nx;
    const float in_lo = (float)(g.rank * g.x_n) + sm + margin, in_hi = (float)((g.rank + 1) * g.x_n) - sm - margin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        unsigned m = 0u;
        if (i < n) {
            PT x = pos[3 * i];
            float gf = (float)x * sc;
            if (!(gf > in_lo && gf < in_hi)) m = route_mask((double)x, g);
        }
        unsigned any = __ballot_sync(0xffffffffu, m != 0u);
        if (any) {
            for (int r = 0; r < g.P; r++) {
                unsigned b = __ballot_sync(0xffffffffu, (m >> r) & 1u);
                if (lane == 0 && b) atomicAdd(&counts[r], (unsigned long long)__popc(b));
            }
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&counts[g.P], (unsigned long long)__popc(any));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (m) list[base + __popc(any & ((1u << lane) - 1u))] = (unsigned long long)i | ((unsigned long long)m << 32);
        }
    }
}

template <typename PT, typename MT>
__global__ void __launch_bounds__(256)
k_route_scatter(const PT *__restrict__ pos, const MT *__restrict__ mass, const unsigned long long *__restrict__ list,
                int64_t n_list, int P, const long long *__restrict__ offsets, unsigned long long *__restrict__ cursor,
                PT *__restrict__ spos, MT *__restrict__ smass, long long *__restrict__ sindex) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t nround = ((n_list + stride - 1) / stride) * stride;
    const int lane = threadIdx.x & 31;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += stride) {
        unsigned long long ent = (j < n_list) ? list[j] : 0ull;
        const unsigned m = (unsigned)(ent >> 32);
        const int64_t i = (int64_t)(ent & 0xffffffffull);
        for (int r = 0; r < P; r++) {
            unsigned b = __ballot_sync(0xffffffffu, (m >> r) & 1u);
            if (!b) continue;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&cursor[r], (unsigned long long)__popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((m >> r) & 1u) {
                long long dst = offsets[r] + (long long)base + __popc(b & ((1u << lane) - 1u));
                spos[3 * dst] = pos[3 * i];
                spos[3 * dst + 1] = pos[3 * i + 1];
                spos[3 * dst + 2] = pos[3 * i + 2];
                if (mass) smass[dst] = mass[i];
                if (sindex) sindex[dst] = (long long)i;      // where the row came from (for results travelling back)
            }
        }
    }
}

extern "C" int nbk_route_count(const void *pos, int pos_dtype, int64_t n, double smoothing, const double *box,
                               const int64_t *nmesh, int P, int rank, uint64_t *counts, uint64_t *list, void *stream) {
    NBK_CHECK_ARG(pos_dtype == NBK_F4 || pos_dtype == NBK_F8, "route_count: bad pos dtype %d", pos_dtype);
    NBK_CHECK_ARG(P >= 1 && P <= 32 && rank >= 0 && rank < P && nmesh[0] % P == 0, "route_count: bad decomposition");
    NBK_CHECK_ARG(smoothing >= 0 && smoothing < 64, "route_count: bad smoothing");
    NBK_CHECK_ARG(n >= 0 && n < (1ll << 32), "route_count: particle count %lld out of range", (long long)n);
    if (n == 0) return NBK_OK;
    RouteGeom g;
    g.scale = (double)nmesh[0] / box[0];
    g.smoothing = smoothing;
    g.Nx = (int)nmesh[0]; g.x_n = (int)(nmesh[0] / P); g.P = P; g.rank = rank;
    cudaStream_t s = (cudaStream_t)stream;
    int grid = nbk_grid_for(n, 256, 8);
    if (pos_dtype == NBK_F4) k_route_count<float><<<grid, 256, 0, s>>>((const float *)pos, n, g, (unsigned long long *)counts, (unsigned long long *)list);
    else k_route_count<double><<<grid, 256, 0, s>>>((const double *)pos, n, g, (unsigned long long *)counts, (unsigned long long *)list);
    NBK_LAUNCHED();
    return NBK_OK;
}

extern "C" int nbk_route_scatter(const void *pos, int pos_dtype, const void *mass, int mass_dtype, const uint64_t *list,
                                 int64_t n_list, int P, const int64_t *offsets, uint64_t *cursor, void *send_pos,
                                 void *send_mass, int64_t *send_index, void *stream) {
    NBK_CHECK_ARG(pos_dtype == NBK_F4 || pos_dtype == NBK_F8, "route_scatter: bad pos dtype %d", pos_dtype);
    NBK_CHECK_ARG(mass == nullptr || mass_dtype == NBK_F4 || mass_dtype == NBK_F8, "route_scatter: bad mass dtype");
    if (n_list <= 0) return NBK_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int grid = nbk_grid_for(n_list, 256, 8);
    const long long *off = (const long long *)offsets;
    unsigned long long *cur = (unsigned long long *)cursor;
#define RS(PT, MT) k_route_scatter<PT, MT><<<grid, 256, 0, s>>>((const PT *)pos, (const MT *)mass, (const unsigned long long *)list, n_list, P, off, cur, (PT *)send_pos, (MT *)send_mass, (long long *)send_index)
    bool pf = pos_dtype == NBK_F4, mf = (mass != nullptr && mass_dtype == NBK_F4);
    if (pf && mf) RS(float, float); else if (pf) RS(float, double); else if (mf) RS(double, float); else RS(double, double);
#undef RS
    NBK_LAUNCHED();
    return NBK_OK;
}
