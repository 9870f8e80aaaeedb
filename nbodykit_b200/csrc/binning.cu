// Fourier-space passes of the FFTPower path:
//   nbk_compensate        -- Field.apply(Compensate*, kind='circular')  (source/mesh/catalog.py:449-594)
//   nbk_interlace_combine -- source/mesh/catalog.py:345-347
//   nbk_power_bin         -- FFTBase._compute_3d_power (algorithms/fftpower.py:115-128) fused with
//                            project_to_basis (:507-701) / MeshSlab (meshtools.py:104-215)
//
// Built with --fmad=false: the float32 coordinate arithmetic that decides the k / mu bin of a mode
// (k_d = fl32(f32(j_d)*f32(2 pi/L_d)), k^2 = fl32(fl32(kx^2+ky^2)+kz^2), |k| = sqrt_rn, mu = div_rn)
// is part of the bit-exact contract (pinned by nbodykit/tests/data/dataset_2d.json).
#include "common.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <math.h>

#define NBK_MAX_ELL 8

// ---------------------------------------------------------------------------------------------
// index helpers: element e of a slab -> integer frequency labels (jx, jy, jz)
// ---------------------------------------------------------------------------------------------
struct SlabGeom {
    int N[3];        // Nx, Ny, Nz of the full mesh
    int Nzc;         // stored length of the last axis
    int transposed;  // 0: [x_n][Ny][Nzc]   1: [y_n][Nx][Nzc]
    int start, count;  // owned range along the first stored axis
    int D1;          // length of the second stored axis
};

static int make_slab(const int64_t *nmesh, int transposed, int64_t start, int64_t count, int hermitian, SlabGeom &g) {
    for (int d = 0; d < 3; d++) {
        NBK_CHECK_ARG(nmesh[d] > 0 && nmesh[d] < (1 << 24), "bad Nmesh[%d]=%lld", d, (long long)nmesh[d]);
        g.N[d] = (int)nmesh[d];
    }
    g.Nzc = hermitian ? g.N[2] / 2 + 1 : g.N[2];
    g.transposed = transposed ? 1 : 0;
    int D0 = transposed ? g.N[1] : g.N[0];
    g.D1 = transposed ? g.N[0] : g.N[1];
    NBK_CHECK_ARG(start >= 0 && count >= 0 && start + count <= D0, "bad slab range [%lld,+%lld) of %d",
                  (long long)start, (long long)count, D0);
    g.start = (int)start;
    g.count = (int)count;
    return NBK_OK;
}

__device__ __forceinline__ void slab_freqs(const SlabGeom &g, int i0, int i1, int kz, int &jx, int &jy, int &jz) {
    int a = nbk_freq(g.start + i0, g.transposed ? g.N[1] : g.N[0]);
    int b = nbk_freq(i1, g.transposed ? g.N[0] : g.N[1]);
    jx = g.transposed ? b : a;
    jy = g.transposed ? a : b;
    jz = nbk_freq(kz, g.N[2]);
}

// ---------------------------------------------------------------------------------------------
// compensation: v /= prod_d f(w_d), w_d = 2 pi j_d / N_d.  The factor is separable, so three 1-D
// tables of reciprocals are built once per (device, kind, N) and cached.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double comp_factor(int kind, double w) {
    const double PI = 3.14159265358979323846;
    double s = sin(0.5 * w);
    s = s * s;
    switch (kind) {
        case NBK_COMP_CIC:
        case NBK_COMP_TSC:
        case NBK_COMP_PCS: {
            if (w == 0.0) return 1.0;
            double x = 0.5 * w;           // numpy.sinc(0.5*w/pi) = sin(0.5 w)/(0.5 w)
            double sc = sin(x) / x;
            int p = kind == NBK_COMP_CIC ? 2 : (kind == NBK_COMP_TSC ? 3 : 4);
            double r = sc;
            for (int i = 1; i < p; i++) r *= sc;
            (void)PI;
            return r;
        }
        case NBK_COMP_CIC_SHOTNOISE: return sqrt(1.0 - 2.0 / 3.0 * s);
        case NBK_COMP_TSC_SHOTNOISE: return sqrt(1.0 - s + 2.0 / 15.0 * (s * s));
        case NBK_COMP_PCS_SHOTNOISE: return sqrt(1.0 - 4.0 / 3.0 * s + 2.0 / 5.0 * (s * s) - 4.0 / 315.0 * (s * s * s));
    }
    return 1.0;
}

__global__ void k_comp_table(double *tab, int kind, int N) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) {
        const double TWO_PI = 6.28318530717958647692;
        double w = TWO_PI * (double)nbk_freq(j, N) / (double)N;
        tab[j] = 1.0 / comp_factor(kind, w);
    }
}

static std::mutex g_ct_mutex;
static std::map<std::tuple<int, int, int>, double *> g_ct;

static int get_comp_table(int kind, int N, cudaStream_t s, double **out) {
    int dev = 0;
    NBK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ct_mutex);
    auto key = std::make_tuple(dev, kind, N);
    auto it = g_ct.find(key);
    if (it != g_ct.end()) { *out = it->second; return NBK_OK; }
    double *p = nullptr;
    NBK_CUDA(cudaMalloc(&p, sizeof(double) * N));
    k_comp_table<<<(N + 255) / 256, 256, 0, s>>>(p, kind, N);
    NBK_LAUNCHED();
    NBK_CUDA(cudaStreamSynchronize(s));
    g_ct[key] = p;
    *out = p;
    return NBK_OK;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_compensate(T *__restrict__ c, SlabGeom g, const double *__restrict__ t0, const double *__restrict__ t1,
             const double *__restrict__ tz) {
    // t0 / t1: tables for the first / second stored axis
    int64_t rows = (int64_t)g.count * g.D1;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        int i0 = (int)(row / g.D1), i1 = (int)(row - (int64_t)i0 * g.D1);
        double f01 = t0[g.start + i0] * t1[i1];
        T *p = c + row * (int64_t)g.Nzc * 2;
        for (int k = threadIdx.x; k < g.Nzc; k += blockDim.x) {
            double f = f01 * tz[k];
            p[2 * k] = (T)((double)p[2 * k] * f);
            p[2 * k + 1] = (T)((double)p[2 * k + 1] * f);
        }
    }
}

extern "C" int nbk_compensate(void *cplx, int dtype, int kind, const int64_t *nmesh, int transposed, int64_t start,
                              int64_t count, void *stream) {
    NBK_CHECK_ARG(dtype == NBK_F4 || dtype == NBK_F8, "compensate: bad dtype %d", dtype);
    NBK_CHECK_ARG(kind >= NBK_COMP_CIC && kind <= NBK_COMP_PCS_SHOTNOISE, "compensate: unknown kind %d", kind);
    SlabGeom g;
    int rc = make_slab(nmesh, transposed, start, count, 1, g);
    if (rc) return rc;
    if (count == 0) return NBK_OK;
    cudaStream_t s = (cudaStream_t)stream;
    double *tx, *ty, *tz;
    if ((rc = get_comp_table(kind, g.N[0], s, &tx))) return rc;
    if ((rc = get_comp_table(kind, g.N[1], s, &ty))) return rc;
    if ((rc = get_comp_table(kind, g.N[2], s, &tz))) return rc;
    int64_t rows = (int64_t)g.count * g.D1;
    int grid = (int)(rows < (int64_t)NBK_SM_COUNT * 16 ? rows : (int64_t)NBK_SM_COUNT * 16);
    int block = g.Nzc >= 256 ? 256 : 64;
    const double *t0 = transposed ? ty : tx, *t1 = transposed ? tx : ty;
    if (dtype == NBK_F4) k_compensate<float><<<grid, block, 0, s>>>((float *)cplx, g, t0, t1, tz);
    else k_compensate<double><<<grid, block, 0, s>>>((double *)cplx, g, t0, t1, tz);
    NBK_LAUNCHED();
    return NBK_OK;
}

// ---------------------------------------------------------------------------------------------
// interlacing combine: s1 = 0.5 s1 + 0.5 s2 exp(0.5 i sum_d k_d H_d), k_d H_d = 2 pi j_d / N_d
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_interlace(T *__restrict__ c1, const T *__restrict__ c2, SlabGeom g) {
    int64_t rows = (int64_t)g.count * g.D1;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        int i0 = (int)(row / g.D1), i1 = (int)(row - (int64_t)i0 * g.D1);
        int jx, jy, jz;
        slab_freqs(g, i0, i1, 0, jx, jy, jz);
        double base = (double)jx / (double)g.N[0] + (double)jy / (double)g.N[1];
        T *p1 = c1 + row * (int64_t)g.Nzc * 2;
        const T *p2 = c2 + row * (int64_t)g.Nzc * 2;
        for (int k = threadIdx.x; k < g.Nzc; k += blockDim.x) {
            double ph = base + (double)nbk_freq(k, g.N[2]) / (double)g.N[2];  // phase / pi
            double sn, cs;
            sincospi(ph, &sn, &cs);
            double a = p1[2 * k], b = p1[2 * k + 1], x = p2[2 * k], y = p2[2 * k + 1];
            p1[2 * k] = (T)(0.5 * a + 0.5 * (x * cs - y * sn));
            p1[2 * k + 1] = (T)(0.5 * b + 0.5 * (x * sn + y * cs));
        }
    }
}

extern "C" int nbk_interlace_combine(void *c1, const void *c2, int dtype, const int64_t *nmesh, const double *box,
                                     int transposed, int64_t start, int64_t count, void *stream) {
    (void)box;  // k_d H_d = 2 pi j_d / N_d is independent of the box size
    NBK_CHECK_ARG(dtype == NBK_F4 || dtype == NBK_F8, "interlace_combine: bad dtype %d", dtype);
    SlabGeom g;
    int rc = make_slab(nmesh, transposed, start, count, 1, g);
    if (rc) return rc;
    if (count == 0) return NBK_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t rows = (int64_t)g.count * g.D1;
    int grid = (int)(rows < (int64_t)NBK_SM_COUNT * 16 ? rows : (int64_t)NBK_SM_COUNT * 16);
    int block = g.Nzc >= 256 ? 256 : 64;
    if (dtype == NBK_F4) k_interlace<float><<<grid, block, 0, s>>>((float *)c1, (const float *)c2, g);
    else k_interlace<double><<<grid, block, 0, s>>>((double *)c1, (const double *)c2, g);
    NBK_LAUNCHED();
    return NBK_OK;
}

// ---------------------------------------------------------------------------------------------
// power binning
// ---------------------------------------------------------------------------------------------
struct BinParams {
    SlabGeom g;
    int coord_mode;  // 4: f32 coords & mu; 8: f64 coords & mu; 48: f32 coords, f64 mu (numpy-scalar los)
    float kf32[3], los32[3];
    double kf64[3], los64[3];
    int Nx, Nmu, nb;
    int Nell;
    int ells[NBK_MAX_ELL];
    int hermitian, is_p3d, clear_zero, has_c2;
    double volume;
};

// number of edges <= x  (numpy.digitize, right=False, increasing edges)
__device__ __forceinline__ int digitize(const double *__restrict__ edges, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (edges[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double legendre(int ell, double x) {
    if (ell == 0) return 1.0;
    double p0 = 1.0, p1 = x;
    for (int n = 1; n < ell; n++) {
        double p2 = ((2 * n + 1) * x * p1 - n * p0) / (n + 1);
        p0 = p1;
        p1 = p2;
    }
    return p1;
}

template <bool SMEM_ACC>
__device__ __forceinline__ void acc_add(double *p, double v) {
    if (v != 0.0) atomicAdd(p, v);
}

// One warp iteration handles 32 consecutive stored modes; equal-bin runs are combined with a segmented
// shuffle reduction so each run costs one atomic per accumulated quantity.
template <typename T, int NELL, bool SMEM_ACC>
__global__ void __launch_bounds__(256)
k_power_bin(const T *__restrict__ c1, const T *__restrict__ c2, BinParams P, const double *__restrict__ k2edges,
            const double *__restrict__ muedges, unsigned long long *__restrict__ g_nsum, double *__restrict__ g_xsum,
            double *__restrict__ g_musum, double *__restrict__ g_ysum) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared layout: k2edges[Nx+1] | muedges[Nmu+1] | (if SMEM_ACC) xsum[nb] musum[nb] ysum[NELL][nb][2] nsum[nb](u32)
    double *s_k2 = reinterpret_cast<double *>(smem_raw);
    double *s_mu = s_k2 + (P.Nx + 1);
    double *s_x = s_mu + (P.Nmu + 1);
    double *s_m = s_x + (SMEM_ACC ? P.nb : 0);
    double *s_y = s_m + (SMEM_ACC ? P.nb : 0);
    unsigned *s_n = reinterpret_cast<unsigned *>(s_y + (SMEM_ACC ? (size_t)NELL * P.nb * 2 : 0));
    for (int i = threadIdx.x; i <= P.Nx; i += blockDim.x) s_k2[i] = k2edges[i];
    for (int i = threadIdx.x; i <= P.Nmu; i += blockDim.x) s_mu[i] = muedges[i];
    if (SMEM_ACC) {
        int nd = P.nb * (2 + 2 * NELL);
        for (int i = threadIdx.x; i < nd; i += blockDim.x) s_x[i] = 0.0;
        for (int i = threadIdx.x; i < P.nb; i += blockDim.x) s_n[i] = 0u;
    }
    __syncthreads();

    double *a_x = SMEM_ACC ? s_x : g_xsum;
    double *a_m = SMEM_ACC ? s_m : g_musum;
    double *a_y = SMEM_ACC ? s_y : g_ysum;

    const SlabGeom &g = P.g;
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)g.count * g.D1 * g.Nzc;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nit = (total + stride - 1) / stride;
    for (int64_t it = 0; it < nit; it++) {
        int64_t e = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int key = -1;
        unsigned cnt = 0;
        double xs = 0, ms = 0, yr[NELL], yi[NELL];
#pragma unroll
        for (int l = 0; l < NELL; l++) { yr[l] = 0; yi[l] = 0; }
        if (e < total) {
            int64_t row = e / g.Nzc;
            int kz = (int)(e - row * g.Nzc);
            int i0 = (int)(row / g.D1), i1 = (int)(row - (int64_t)i0 * g.D1);
            int jx, jy, jz;
            slab_freqs(g, i0, i1, kz, jx, jy, jz);
            if (!P.hermitian) jz = nbk_freq(kz, g.N[2]);
            double k2d, knorm, mu;
            if (P.coord_mode == 8) {
                double kx = (double)jx * P.kf64[0], ky = (double)jy * P.kf64[1], kzv = (double)jz * P.kf64[2];
                k2d = (kx * kx + ky * ky) + kzv * kzv;
                knorm = sqrt(k2d);
                mu = ((kx * P.los64[0] + ky * P.los64[1]) + kzv * P.los64[2]) / knorm;
                if (knorm == 0.0) mu = 0.0;
            } else {
                float kx = (float)jx * P.kf32[0], ky = (float)jy * P.kf32[1], kzv = (float)jz * P.kf32[2];
                float k2 = (kx * kx + ky * ky) + kzv * kzv;
                float kn = sqrtf(k2);  // IEEE sqrt (-prec-sqrt=true), numpy `** 0.5` on f4 -> sqrtf
                k2d = (double)k2;
                knorm = (double)kn;
                if (P.coord_mode == 4) {
                    float m = ((kx * P.los32[0] + ky * P.los32[1]) + kzv * P.los32[2]) / kn;
                    mu = (kn == 0.0f) ? 0.0 : (double)m;
                } else {
                    double m = (((double)kx * P.los64[0] + (double)ky * P.los64[1]) + (double)kzv * P.los64[2]) / knorm;
                    mu = (kn == 0.0f) ? 0.0 : m;
                }
            }
            int dig_x = digitize(s_k2, P.Nx + 1, k2d);
            int dig_mu = digitize(s_mu, P.Nmu + 1, mu);
            key = dig_x * (P.Nmu + 2) + dig_mu;
            bool nonsing = P.hermitian && (jz > 0);
            double wH = nonsing ? 2.0 : 1.0;
            cnt = nonsing ? 2u : 1u;
            xs = knorm * wH;
            ms = mu * wH;
            // the statistic y
            double a = (double)c1[2 * e], b = (double)c1[2 * e + 1], yre, yim;
            if (P.is_p3d) { yre = a; yim = b; }
            else {
                double c = a, d = b;
                if (P.has_c2) { c = (double)c2[2 * e]; d = (double)c2[2 * e + 1]; }
                yre = (a * c + b * d) * P.volume;   // c1 * conj(c2)
                yim = (b * c - a * d) * P.volume;
                if (P.clear_zero && jx == 0 && jy == 0 && jz == 0) { yre = 0; yim = 0; }
            }
#pragma unroll
            for (int l = 0; l < NELL; l++) {
                int ell = P.ells[l];
                double f = legendre(ell, mu) * (2.0 * ell + 1.0);
                double re = f * yre, im = f * yim;
                if (nonsing) {
                    if (ell & 1) { re = 0.0; im *= 2.0; }
                    else { re *= 2.0; im = 0.0; }
                }
                yr[l] = re;
                yi[l] = im;
            }
        }
        // ---- segmented reduction over equal-key runs
        int prev = __shfl_up_sync(0xffffffffu, key, 1);
        bool head = (lane == 0) || (prev != key);
        unsigned heads = __ballot_sync(0xffffffffu, head);
        if (heads == 1u) {  // whole warp in one bin
            if (key >= 0) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                    xs += __shfl_xor_sync(0xffffffffu, xs, o);
                    ms += __shfl_xor_sync(0xffffffffu, ms, o);
#pragma unroll
                    for (int l = 0; l < NELL; l++) {
                        yr[l] += __shfl_xor_sync(0xffffffffu, yr[l], o);
                        yi[l] += __shfl_xor_sync(0xffffffffu, yi[l], o);
                    }
                }
            }
        } else {
            int seg = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int so = __shfl_down_sync(0xffffffffu, seg, o);
                bool take = (lane + o < 32) && (so == seg);
                unsigned c_o = __shfl_down_sync(0xffffffffu, cnt, o);
                double x_o = __shfl_down_sync(0xffffffffu, xs, o);
                double m_o = __shfl_down_sync(0xffffffffu, ms, o);
                if (take) { cnt += c_o; xs += x_o; ms += m_o; }
#pragma unroll
                for (int l = 0; l < NELL; l++) {
                    double r_o = __shfl_down_sync(0xffffffffu, yr[l], o);
                    double i_o = __shfl_down_sync(0xffffffffu, yi[l], o);
                    if (take) { yr[l] += r_o; yi[l] += i_o; }
                }
            }
        }
        if (head && key >= 0) {
            if (SMEM_ACC) atomicAdd(&s_n[key], cnt);
            else atomicAdd(&g_nsum[key], (unsigned long long)cnt);
            acc_add<SMEM_ACC>(&a_x[key], xs);
            acc_add<SMEM_ACC>(&a_m[key], ms);
#pragma unroll
            for (int l = 0; l < NELL; l++) {
                acc_add<SMEM_ACC>(&a_y[((size_t)l * P.nb + key) * 2], yr[l]);
                acc_add<SMEM_ACC>(&a_y[((size_t)l * P.nb + key) * 2 + 1], yi[l]);
            }
        }
    }
    if (SMEM_ACC) {
        __syncthreads();
        for (int i = threadIdx.x; i < P.nb; i += blockDim.x) {
            unsigned c = s_n[i];
            if (c) atomicAdd(&g_nsum[i], (unsigned long long)c);
            if (s_x[i] != 0.0) atomicAdd(&g_xsum[i], s_x[i]);
            if (s_m[i] != 0.0) atomicAdd(&g_musum[i], s_m[i]);
        }
        for (int i = threadIdx.x; i < NELL * P.nb * 2; i += blockDim.x)
            if (s_y[i] != 0.0) atomicAdd(&g_ysum[i], s_y[i]);
    }
}

// device copies of the edge arrays, cached per (device, content)
struct EdgeCache {
    std::vector<double> host;
    double *dev = nullptr;
};
static std::mutex g_edge_mutex;
static std::map<int, std::vector<EdgeCache>> g_edges;

static int get_edges(const double *host, int n, cudaStream_t s, double **out) {
    int dev = 0;
    NBK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_edge_mutex);
    auto &vec = g_edges[dev];
    for (auto &c : vec)
        if ((int)c.host.size() == n && memcmp(c.host.data(), host, sizeof(double) * n) == 0) { *out = c.dev; return NBK_OK; }
    if (vec.size() > 64) {  // bounded cache
        for (auto &c : vec) cudaFree(c.dev);
        vec.clear();
    }
    EdgeCache c;
    c.host.assign(host, host + n);
    NBK_CUDA(cudaMalloc(&c.dev, sizeof(double) * n));
    NBK_CUDA(cudaMemcpyAsync(c.dev, c.host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
    NBK_CUDA(cudaStreamSynchronize(s));
    vec.push_back(c);
    *out = vec.back().dev;
    return NBK_OK;
}

template <typename T, int NELL>
static int launch_bin(const void *c1, const void *c2, const BinParams &P, const double *d_k2, const double *d_mu,
                      int64_t *nsum, double *xsum, double *musum, double *ysum, cudaStream_t s) {
    size_t edge_bytes = sizeof(double) * (P.Nx + 1 + P.Nmu + 1);
    size_t acc_bytes = (size_t)P.nb * (sizeof(double) * (2 + 2 * NELL) + sizeof(unsigned));
    bool smem_acc = edge_bytes + acc_bytes <= 100 * 1024;
    size_t smem = edge_bytes + (smem_acc ? acc_bytes : 0);
    NBK_CHECK_ARG(smem <= 227 * 1024, "power_bin: too many bin edges for shared memory");
    int64_t total = (int64_t)P.g.count * P.g.D1 * P.g.Nzc;
    int per_sm = smem_acc ? 2 : 4;
    int grid = nbk_grid_for(total, 256, per_sm);
    if (smem_acc) {
        NBK_CUDA(cudaFuncSetAttribute(k_power_bin<T, NELL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_power_bin<T, NELL, true><<<grid, 256, smem, s>>>((const T *)c1, (const T *)c2, P, d_k2, d_mu,
                                                            (unsigned long long *)nsum, xsum, musum, ysum);
    } else {
        NBK_CUDA(cudaFuncSetAttribute(k_power_bin<T, NELL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_power_bin<T, NELL, false><<<grid, 256, smem, s>>>((const T *)c1, (const T *)c2, P, d_k2, d_mu,
                                                             (unsigned long long *)nsum, xsum, musum, ysum);
    }
    NBK_LAUNCHED();
    return NBK_OK;
}

template <typename T>
static int launch_bin_ell(const void *c1, const void *c2, const BinParams &P, const double *d_k2, const double *d_mu,
                          int64_t *nsum, double *xsum, double *musum, double *ysum, cudaStream_t s) {
    switch (P.Nell) {
        case 1: return launch_bin<T, 1>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 2: return launch_bin<T, 2>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 3: return launch_bin<T, 3>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 4: return launch_bin<T, 4>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 5: return launch_bin<T, 5>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 6: return launch_bin<T, 6>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 7: return launch_bin<T, 7>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
        case 8: return launch_bin<T, 8>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
    }
    nbk_set_error("power_bin: Nell=%d unsupported (1..%d)", P.Nell, NBK_MAX_ELL);
    return NBK_ERR_UNSUPPORTED;
}

extern "C" int nbk_power_bin(const void *c1, const void *c2, int dtype, int is_p3d, double volume, int clear_zero,
                             const int64_t *nmesh, const double *box, int transposed, int64_t start, int64_t count,
                             int coord_dtype, const double *k2edges, int Nx, const double *muedges, int Nmu,
                             const double *los, const int *ells, int Nell, int hermitian, int64_t *nsum,
                             double *xsum, double *musum, double *ysum, void *stream) {
    NBK_CHECK_ARG(dtype == NBK_F4 || dtype == NBK_F8, "power_bin: bad dtype %d", dtype);
    NBK_CHECK_ARG(coord_dtype == 4 || coord_dtype == 8 || coord_dtype == 48, "power_bin: bad coord_dtype %d", coord_dtype);
    NBK_CHECK_ARG(Nx >= 0 && Nmu >= 1, "power_bin: need Nx >= 0 and Nmu >= 1");
    NBK_CHECK_ARG(Nell >= 1 && Nell <= NBK_MAX_ELL && ells[0] == 0, "power_bin: ells must start with 0, 1 <= Nell <= %d", NBK_MAX_ELL);
    BinParams P;
    int rc = make_slab(nmesh, transposed, start, count, hermitian, P.g);
    if (rc) return rc;
    if (count == 0) return NBK_OK;
    P.coord_mode = coord_dtype;
    const double TWO_PI = 6.283185307179586476925286766559;
    for (int d = 0; d < 3; d++) {
        NBK_CHECK_ARG(box[d] > 0, "power_bin: bad BoxSize");
        P.kf64[d] = TWO_PI / box[d];
        P.kf32[d] = (float)P.kf64[d];
        P.los64[d] = los[d];
        P.los32[d] = (float)los[d];
    }
    P.Nx = Nx; P.Nmu = Nmu; P.nb = (Nx + 2) * (Nmu + 2);
    P.Nell = Nell;
    for (int l = 0; l < NBK_MAX_ELL; l++) P.ells[l] = l < Nell ? ells[l] : 0;
    for (int l = 0; l < Nell; l++) NBK_CHECK_ARG(ells[l] >= 0 && ells[l] <= 64, "power_bin: bad multipole %d", ells[l]);
    P.hermitian = hermitian ? 1 : 0;
    P.is_p3d = is_p3d ? 1 : 0;
    P.clear_zero = clear_zero ? 1 : 0;
    P.has_c2 = (c2 != nullptr && c2 != c1) ? 1 : 0;
    P.volume = volume;
    cudaStream_t s = (cudaStream_t)stream;
    double *d_k2, *d_mu;
    if ((rc = get_edges(k2edges, Nx + 1, s, &d_k2))) return rc;
    if ((rc = get_edges(muedges, Nmu + 1, s, &d_mu))) return rc;
    if (dtype == NBK_F4) return launch_bin_ell<float>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
    return launch_bin_ell<double>(c1, c2, P, d_k2, d_mu, nsum, xsum, musum, ysum, s);
}
