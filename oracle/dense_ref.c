/* dense_ref.c -- C + OpenMP restatement of oracle/dense_ref.py (fp32 fields,
 * fp64 dot products): the timed CPU arm of bench.py ("cpu_baseline", kind
 * "port") and of `bench.py --impl reference`.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY -- never linked into libsplashy_cuda.so.
 * PARITY UNPINNED: the F# reference cannot be built here (no .NET toolchain,
 * SURVEY.md 8c); this file is validated against dense_ref.py, which is pinned
 * to the reference's divergence KATs and SURVEY App. B through sparse_ref.py.
 *
 * It is an optimistic stand-in for the reference's CPU path: dense arrays
 * instead of Dictionary<Coord,Cell> hashing (Grid.fs:12), Jacobi-PCG instead of
 * MathNet's O(n^3)-class direct solve (Pressure.fs:69), all host threads instead
 * of one (the reference is single-threaded).
 *
 * Layout: idx = (k*ny + j)*nx + i; U on +x faces (Pressure.fs:18-31).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { AIR = 0, FLUID = 1, SOLID = 2 };
enum { B_XM = 1, B_XP = 2, B_YM = 4, B_YP = 8, B_ZM = 16, B_ZP = 32, B_FLUID = 64 };

typedef struct {
  int nx, ny, nz;
  int64_t plane, n;
} box_t;

static inline int nonsolid(uint8_t m) { return m == AIR || m == FLUID; }

int orc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU arm asks for the
 * host's cores explicitly */
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Pressure.fs:34-38,42-43,53 */
void orc_codes(int nx, int ny, int nz, const uint8_t* lab, uint8_t* codes) {
  const int64_t plane = (int64_t)nx * ny;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int64_t idx = (int64_t)k * plane + (int64_t)j * nx + i;
        unsigned c = 0;
        if (i > 0 && nonsolid(lab[idx - 1])) c |= B_XM;
        if (i < nx - 1 && nonsolid(lab[idx + 1])) c |= B_XP;
        if (j > 0 && nonsolid(lab[idx - nx])) c |= B_YM;
        if (j < ny - 1 && nonsolid(lab[idx + nx])) c |= B_YP;
        if (k > 0 && nonsolid(lab[idx - plane])) c |= B_ZM;
        if (k < nz - 1 && nonsolid(lab[idx + plane])) c |= B_ZP;
        if (lab[idx] == FLUID) c |= B_FLUID;
        codes[idx] = (uint8_t)c;
      }
}

/* Pressure.fs:16-32,64-68: out = scale * (in - out) on Fluid cells */
void orc_div_rhs(int nx, int ny, int nz, const uint8_t* codes, const float* u,
                 const float* v, const float* w, float scale, float* out) {
  const int64_t plane = (int64_t)nx * ny, n = plane * nz;
#pragma omp parallel for schedule(static)
  for (int64_t idx = 0; idx < n; ++idx) {
    const unsigned c = codes[idx];
    float d = 0.f;
    if (c & B_FLUID) {
      const float ox = (c & B_XP) ? u[idx] : 0.f, oy = (c & B_YP) ? v[idx] : 0.f,
                  oz = (c & B_ZP) ? w[idx] : 0.f;
      const float ix = (c & B_XM) ? u[idx - 1] : 0.f, iy = (c & B_YM) ? v[idx - nx] : 0.f,
                  iz = (c & B_ZM) ? w[idx - plane] : 0.f;
      d = scale * (((ix - ox) + (iy - oy)) + (iz - oz));
    }
    out[idx] = d;
  }
}

static inline float inv_diag(unsigned c) {
  const int n = __builtin_popcount(c & 63u);
  return ((c & B_FLUID) && n > 0) ? 1.0f / (float)n : 0.f;
}

/* Jacobi-PCG in the operation order of dense_ref.pcg / the CUDA path.
 * fixed_iters > 0 runs exactly that many iterations.  Returns iterations. */
int orc_pcg_jacobi(int nx, int ny, int nz, const uint8_t* codes, const float* b, float* x,
                   double tol, int max_iters, int fixed_iters, double* r_inf, double* b_inf) {
  const int64_t plane = (int64_t)nx * ny, n = plane * nz;
  float* r = (float*)malloc(sizeof(float) * n);
  float* s = (float*)malloc(sizeof(float) * n);
  float* q = (float*)malloc(sizeof(float) * n);
  double sigma = 0.0, bmax = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sigma) reduction(max : bmax)
  for (int64_t i = 0; i < n; ++i) {
    x[i] = 0.f;
    r[i] = (codes[i] & B_FLUID) ? b[i] : 0.f;
    const float z = r[i] * inv_diag(codes[i]);
    s[i] = z;
    sigma += (double)r[i] * (double)z;
    bmax = fmax(bmax, fabs((double)r[i]));
  }
  *b_inf = bmax;
  *r_inf = bmax;
  int it = 0;
  const int limit = fixed_iters > 0 ? fixed_iters : max_iters;
  if (bmax == 0.0 && fixed_iters <= 0) goto done;
  for (it = 1; it <= limit; ++it) {
    double sq = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : sq)
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j) {
        int64_t idx = (int64_t)k * plane + (int64_t)j * nx;
        for (int i = 0; i < nx; ++i, ++idx) {
          const unsigned c = codes[idx];
          float a = 0.f;
          if (c & B_FLUID) {
            const float s0 = s[idx];
            a += (c & B_XM) ? (s0 - s[idx - 1]) : 0.f;
            a += (c & B_XP) ? (s0 - s[idx + 1]) : 0.f;
            a += (c & B_YM) ? (s0 - s[idx - nx]) : 0.f;
            a += (c & B_YP) ? (s0 - s[idx + nx]) : 0.f;
            a += (c & B_ZM) ? (s0 - s[idx - plane]) : 0.f;
            a += (c & B_ZP) ? (s0 - s[idx + plane]) : 0.f;
            sq += (double)s0 * (double)a;
          }
          q[idx] = a;
        }
      }
    if (sq == 0.0) break;
    const float alpha = (float)(sigma / sq);
    double sigma_new = 0.0, rmax = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sigma_new) reduction(max : rmax)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * s[i];
      const float ri = r[i] - alpha * q[i];
      r[i] = ri;
      sigma_new += (double)ri * (double)(ri * inv_diag(codes[i]));
      rmax = fmax(rmax, fabs((double)ri));
    }
    *r_inf = rmax;
    if (fixed_iters <= 0 && rmax <= tol * bmax) break;
    const float beta = (float)(sigma_new / sigma);
    sigma = sigma_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) s[i] = r[i] * inv_diag(codes[i]) + beta * s[i];
  }
  if (it > limit) it = limit;
done:
  free(r);
  free(s);
  free(q);
  return it;
}

/* Pressure.fs:96-129, each face once (SURVEY A.7 intended) */
void orc_grad_sub(int nx, int ny, int nz, const uint8_t* lab, const float* p, float* u,
                  float* v, float* w, float kappa) {
  const int64_t plane = (int64_t)nx * ny;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int64_t idx = (int64_t)k * plane + (int64_t)j * nx + i;
        const uint8_t m0 = lab[idx];
        const int ns0 = nonsolid(m0), fl0 = (m0 == FLUID);
        const float p0 = fl0 ? p[idx] : 0.f;
        const int inb[3] = {i < nx - 1, j < ny - 1, k < nz - 1};
        const int64_t off[3] = {1, nx, plane};
        float* f[3] = {u, v, w};
        for (int a = 0; a < 3; ++a) {
          const uint8_t m1 = inb[a] ? lab[idx + off[a]] : 255;
          const int ns1 = nonsolid(m1), fl1 = (m1 == FLUID);
          if (ns0 && ns1 && (fl0 || fl1)) {
            const float p1 = fl1 ? p[idx + off[a]] : 0.f;
            f[a][idx] = f[a][idx] - kappa * (p1 - p0);
          } else if ((fl0 && !ns1) || (!ns0 && fl1)) {
            f[a][idx] = 0.f;
          }
        }
      }
}

/* Convection.fs:21-39 with the intended corner key; corners outside -> 0 */
static inline float fetchf(const float* F, const box_t* b, int i, int j, int k) {
  if ((unsigned)i >= (unsigned)b->nx || (unsigned)j >= (unsigned)b->ny ||
      (unsigned)k >= (unsigned)b->nz)
    return 0.f;
  return F[(int64_t)k * b->plane + (int64_t)j * b->nx + i];
}

static inline float samplef(const float* F, const box_t* b, int bi, int bj, int bk, float ox,
                            float oy, float oz) {
  const float fx = floorf(ox), fy = floorf(oy), fz = floorf(oz);
  const float wx1 = ox - fx, wy1 = oy - fy, wz1 = oz - fz;
  const float wx[2] = {1.f - wx1, wx1}, wy[2] = {1.f - wy1, wy1}, wz[2] = {1.f - wz1, wz1};
  const int i0 = bi + (int)fx, j0 = bj + (int)fy, k0 = bk + (int)fz;
  float t = 0.f;
  for (int a = 0; a < 2; ++a)
    for (int c = 0; c < 2; ++c)
      for (int d = 0; d < 2; ++d)
        t += fetchf(F, b, i0 + a, j0 + c, k0 + d) * wx[a] * wy[c] * wz[d];
  return t;
}

/* Convection.fs:50-66, intended semantics, Jacobi update */
void orc_advect(int nx, int ny, int nz, const uint8_t* lab, const float* u, const float* v,
                const float* w, float* u2, float* v2, float* w2, float dt_over_h) {
  box_t b = {nx, ny, nz, (int64_t)nx * ny, (int64_t)nx * ny * nz};
  const float c = -dt_over_h, h = 0.5f;
  const float* F[3] = {u, v, w};
  float* G[3] = {u2, v2, w2};
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int64_t idx = (int64_t)k * b.plane + (int64_t)j * nx + i;
        for (int comp = 0; comp < 3; ++comp) {
          float val = F[comp][idx];
          if (lab[idx] == FLUID) {
            const float ax = comp == 0 ? h : 0.f, ay = comp == 1 ? h : 0.f,
                        az = comp == 2 ? h : 0.f;
            const float u0 = samplef(u, &b, i, j, k, ax - h, ay, az);
            const float v0 = samplef(v, &b, i, j, k, ax, ay - h, az);
            const float w0 = samplef(w, &b, i, j, k, ax, ay, az - h);
            const float mx = h * c * u0, my = h * c * v0, mz = h * c * w0;
            const float u1 = samplef(u, &b, i, j, k, mx + ax - h, my + ay, mz + az);
            const float v1 = samplef(v, &b, i, j, k, mx + ax, my + ay - h, mz + az);
            const float w1 = samplef(w, &b, i, j, k, mx + ax, my + ay, mz + az - h);
            val = samplef(F[comp], &b, i, j, k, c * u1, c * v1, c * w1);
          }
          G[comp][idx] = val;
        }
      }
}

/* cell-centred scalar carried by the flow: the trace of Convection.fs:50-57 from the cell
 * centre, trilinear sample of the scalar (Convection.fs:21-39); dense_ref.advect_scalar */
void orc_advect_scalar(int nx, int ny, int nz, const uint8_t* lab, const float* u,
                       const float* v, const float* w, const float* s, float* s2,
                       float dt_over_h) {
  box_t b = {nx, ny, nz, (int64_t)nx * ny, (int64_t)nx * ny * nz};
  const float c = -dt_over_h, h = 0.5f;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int64_t idx = (int64_t)k * b.plane + (int64_t)j * nx + i;
        float val = s[idx];
        if (lab[idx] == FLUID) {
          const float u0 = samplef(u, &b, i, j, k, -h, 0.f, 0.f);
          const float v0 = samplef(v, &b, i, j, k, 0.f, -h, 0.f);
          const float w0 = samplef(w, &b, i, j, k, 0.f, 0.f, -h);
          const float mx = h * c * u0, my = h * c * v0, mz = h * c * w0;
          const float u1 = samplef(u, &b, i, j, k, mx - h, my, mz);
          const float v1 = samplef(v, &b, i, j, k, mx, my - h, mz);
          const float w1 = samplef(w, &b, i, j, k, mx, my, mz - h);
          val = samplef(s, &b, i, j, k, c * u1, c * v1, c * w1);
        }
        s2[idx] = val;
      }
}

/* Pressure.fs:145-149: max |divergence| over Fluid */
double orc_max_div(int nx, int ny, int nz, const uint8_t* codes, const float* u,
                   const float* v, const float* w) {
  const int64_t n = (int64_t)nx * ny * nz;
  float* d = (float*)malloc(sizeof(float) * n);
  orc_div_rhs(nx, ny, nz, codes, u, v, w, 1.0f, d);
  double m = 0.0;
#pragma omp parallel for schedule(static) reduction(max : m)
  for (int64_t i = 0; i < n; ++i) m = fmax(m, fabs((double)d[i]));
  free(d);
  return m;
}

/* One advect -> project step (Simulator.fs:82,88-89) in place on u, v, w; p out.
 * Returns PCG iterations. */
int orc_step(int nx, int ny, int nz, const uint8_t* lab, float* u, float* v, float* w,
             float* p, double dt, double h, double rho, double tol, int max_iters,
             int fixed_iters, double* r_inf, double* b_inf, double* max_div) {
  const int64_t n = (int64_t)nx * ny * nz;
  uint8_t* codes = (uint8_t*)malloc(n);
  float* u2 = (float*)malloc(sizeof(float) * n);
  float* v2 = (float*)malloc(sizeof(float) * n);
  float* w2 = (float*)malloc(sizeof(float) * n);
  float* b = (float*)malloc(sizeof(float) * n);
  orc_codes(nx, ny, nz, lab, codes);
  orc_advect(nx, ny, nz, lab, u, v, w, u2, v2, w2, (float)(dt / h));
  memcpy(u, u2, sizeof(float) * n);
  memcpy(v, v2, sizeof(float) * n);
  memcpy(w, w2, sizeof(float) * n);
  orc_div_rhs(nx, ny, nz, codes, u, v, w, (float)((h * rho) / dt), b);
  const int it = orc_pcg_jacobi(nx, ny, nz, codes, b, p, tol, max_iters, fixed_iters, r_inf,
                                b_inf);
  orc_grad_sub(nx, ny, nz, lab, p, u, v, w, (float)(dt / (h * rho)));
  *max_div = orc_max_div(nx, ny, nz, codes, u, v, w);
  free(codes);
  free(u2);
  free(v2);
  free(w2);
  free(b);
  return it;
}
