// host_setup.cpp — problem gallery, index partitions and smoothed-aggregation setup on the
// host (product code; the hierarchy is "built on the host and uploaded once", BASELINE.json
// north_star).  Algorithm definitions: SURVEY.md Appendix B [DEFINED-HERE]; the reference
// snapshot has no code to follow (/root/reference/README.md:1-2).  Checked bit-exact
// (structure, index maps, aggregates) against oracle/amg_oracle.py in tests/test_setup_parity.py.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <numeric>
#include <stdexcept>
#include <utility>

#include "host.hpp"

namespace pamg {

// An exception must not leave an OpenMP region (that is std::terminate, and the host process -- the Julia
// session -- dies).  Loop bodies that can throw run under this guard: it keeps the first exception and
// rethrows it after the region, where the C ABI's guard() turns it into a status code.
namespace {
struct OmpGuard {
  std::exception_ptr ep;
  template <class F>
  void run(F&& f) noexcept {
    try {
      f();
    } catch (...) {
#pragma omp critical(pamg_omp_guard)
      if (!ep) ep = std::current_exception();
    }
  }
  void rethrow() {
    if (ep) std::rethrow_exception(ep);
  }
};
}  // namespace

// ------------------------------------------------------------------------------------------
// partition  (PartitionedArrays uniform_partition / local_range, App. A)
// ------------------------------------------------------------------------------------------
void local_range(int64_t p, int64_t nparts, int64_t n, int64_t* off, int64_t* len) {
  int64_t l = n / nparts, o = l * p, rem = n % nparts;
  if (rem > 0 && p >= nparts - rem) {  // the LAST rem blocks get one extra item
    o += p - (nparts - rem);
    l += 1;
  }
  *off = o;
  *len = l;
}

void uniform_partition(int ndim, const int64_t* dims, const int32_t* pdims, std::vector<int32_t>& owner) {
  int64_t n = 1;
  for (int a = 0; a < ndim; ++a) n *= dims[a];
  std::vector<std::vector<int32_t>> part_of(ndim);
  for (int a = 0; a < ndim; ++a) {
    part_of[a].resize(dims[a]);
    for (int32_t p = 0; p < pdims[a]; ++p) {
      int64_t off, len;
      local_range(p, pdims[a], dims[a], &off, &len);
      for (int64_t i = off; i < off + len; ++i) part_of[a][i] = p;
    }
  }
  owner.resize(n);
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < n; ++g) {
    int64_t rem = g, ps = 1, o = 0;
    for (int a = 0; a < ndim; ++a) {
      int64_t c = rem % dims[a];
      rem /= dims[a];
      o += part_of[a][c] * ps;
      ps *= pdims[a];
    }
    owner[g] = (int32_t)o;
  }
}

// ------------------------------------------------------------------------------------------
// gallery
// ------------------------------------------------------------------------------------------
template <class RowFn>
static void stencil_matrix(int ndim, const int64_t* dims, Csr& A, RowFn rowfn) {
  int64_t n = 1;
  int64_t stride[3] = {1, 1, 1};
  for (int a = 0; a < ndim; ++a) {
    stride[a] = n;
    n *= dims[a];
  }
  A.nrows = A.ncols = n;
  huge_reserve(A.ptr, (size_t)n + 1);
  A.ptr.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < n; ++g) {
    int64_t rem = g, cnt = 1;
    for (int a = 0; a < ndim; ++a) {
      int64_t c = rem % dims[a];
      rem /= dims[a];
      cnt += (c > 0) + (c < dims[a] - 1);
    }
    A.ptr[g + 1] = cnt;
  }
  for (int64_t g = 0; g < n; ++g) A.ptr[g + 1] += A.ptr[g];
  huge_reserve(A.col, (size_t)A.ptr[n]);
  huge_reserve(A.val, (size_t)A.ptr[n]);
  A.col.resize(A.ptr[n]);
  A.val.resize(A.ptr[n]);
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < n; ++g) {
    int64_t c[3] = {0, 0, 0}, rem = g;
    for (int a = 0; a < ndim; ++a) {
      c[a] = rem % dims[a];
      rem /= dims[a];
    }
    double off[3][2], dg;
    rowfn(g, c, off, &dg);
    int64_t k = A.ptr[g];
    for (int a = ndim - 1; a >= 0; --a)  // ascending gid: -z, -y, -x, self, +x, +y, +z
      if (c[a] > 0) {
        A.col[k] = g - stride[a];
        A.val[k++] = off[a][0];
      }
    A.col[k] = g;
    A.val[k++] = dg;
    for (int a = 0; a < ndim; ++a)
      if (c[a] < dims[a] - 1) {
        A.col[k] = g + stride[a];
        A.val[k++] = off[a][1];
      }
  }
}

void gallery_poisson(int ndim, const int64_t* dims, Csr& A) {
  const double dg0 = 2.0 * ndim;
  stencil_matrix(ndim, dims, A, [=](int64_t, const int64_t*, double off[3][2], double* dg) {
    for (int a = 0; a < 3; ++a) off[a][0] = off[a][1] = -1.0;
    *dg = dg0;
  });
}

void gallery_diffusion_jump(int ndim, const int64_t* dims, int blocks, double kmax, double eps_z, Csr& A) {
  auto kcell = [=](const int64_t* c, int a) {
    int64_t par = 0;
    for (int b = 0; b < ndim; ++b) par += (c[b] * blocks) / dims[b];
    double k = (par % 2 == 0) ? 1.0 : kmax;
    return (ndim == 3 && a == 2) ? eps_z * k : k;
  };
  stencil_matrix(ndim, dims, A, [=](int64_t, const int64_t* c, double off[3][2], double* dg) {
    double d = 0.0;
    for (int a = 0; a < ndim; ++a) {
      double ka = kcell(c, a);
      for (int s = 0; s < 2; ++s) {
        int64_t cn[3] = {c[0], c[1], c[2]};
        cn[a] += s ? 1 : -1;
        bool inside = cn[a] >= 0 && cn[a] < dims[a];
        double kn = inside ? kcell(cn, a) : ka;
        double t = 2.0 * ka * kn / (ka + kn);
        d += t;
        off[a][s] = -t;
      }
    }
    *dg = d;
  });
}


// 3-D linear elasticity, Q1 hexahedra on unit cubes (oracle/amg_oracle.py elasticity_q1): nx x ny x nz FREE
// nodes, 3 DOFs each (gid = 3 node + component); the node layer at i = -1 is clamped and eliminated.
// Entries are (lam * sum NL + mu * sum NM) / 72 with integer element sums, hence bit-identical to the oracle.
void gallery_elasticity(const int64_t* dims, double E, double nu, Csr& A, std::vector<double>& coords) {
  const int64_t nx = dims[0], ny = dims[1], nz = dims[2];
  if (nx < 1 || ny < 2 || nz < 2) throw std::runtime_error("elasticity gallery needs nx >= 1, ny >= 2, nz >= 2");
  const double lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  const double mu = E / (2.0 * (1.0 + nu));
  int G[8][8][3][3];
  for (int a = 0; a < 8; ++a)
    for (int b = 0; b < 8; ++b) {
      const int ab[3] = {a & 1, (a >> 1) & 1, (a >> 2) & 1}, bb[3] = {b & 1, (b >> 1) & 1, (b >> 2) & 1};
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
          int v;
          if (d == e) {
            v = (2 * ab[d] - 1) * (2 * bb[d] - 1) * 2;
            for (int m = 0; m < 3; ++m)
              if (m != d) v *= (ab[m] == bb[m]) ? 2 : 1;
          } else {
            const int m = 3 - d - e;
            v = (2 * ab[d] - 1) * (2 * bb[e] - 1) * ((ab[m] == bb[m]) ? 6 : 3);
          }
          G[a][b][d][e] = v;
        }
    }
  const int64_t nn = nx * ny * nz, n = 3 * nn;
  A.nrows = A.ncols = n;
  huge_reserve(A.ptr, (size_t)n + 1);
  A.ptr.assign(n + 1, 0);
  coords.resize(3 * nn);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < nn; ++v) {
    const int64_t i = v % nx, j = (v / nx) % ny, k = v / (nx * ny);
    coords[3 * v] = (double)(i + 1);
    coords[3 * v + 1] = (double)j;
    coords[3 * v + 2] = (double)k;
    int64_t cnt = 0;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx)
          if (i + dx >= 0 && i + dx < nx && j + dy >= 0 && j + dy < ny && k + dz >= 0 && k + dz < nz) ++cnt;
    for (int d = 0; d < 3; ++d) A.ptr[3 * v + d + 1] = 3 * cnt;
  }
  for (int64_t r = 0; r < n; ++r) A.ptr[r + 1] += A.ptr[r];
  huge_reserve(A.col, (size_t)A.ptr[n]);
  huge_reserve(A.val, (size_t)A.ptr[n]);
  A.col.resize(A.ptr[n]);
  A.val.resize(A.ptr[n]);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < nn; ++v) {
    const int64_t i = v % nx, j = (v / nx) % ny, k = v / (nx * ny);
    int64_t q[3] = {A.ptr[3 * v], A.ptr[3 * v + 1], A.ptr[3 * v + 2]};
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int64_t i2 = i + dx, j2 = j + dy, k2 = k + dz;
          if (i2 < 0 || i2 >= nx || j2 < 0 || j2 >= ny || k2 < 0 || k2 >= nz) continue;
          const int64_t u = i2 + nx * (j2 + ny * k2);
          long long nl[3][3] = {{0}}, nm[3][3] = {{0}};
          for (int64_t cz = std::max(k, k2) - 1; cz <= std::min(k, k2); ++cz) {
            if (cz < 0 || cz > nz - 2) continue;
            for (int64_t cy = std::max(j, j2) - 1; cy <= std::min(j, j2); ++cy) {
              if (cy < 0 || cy > ny - 2) continue;
              for (int64_t cx = std::max(i, i2) - 1; cx <= std::min(i, i2); ++cx) {
                if (cx < -1 || cx > nx - 2) continue;
                const int a = (int)((i - cx) + 2 * (j - cy) + 4 * (k - cz));
                const int b = (int)((i2 - cx) + 2 * (j2 - cy) + 4 * (k2 - cz));
                for (int d = 0; d < 3; ++d)
                  for (int e = 0; e < 3; ++e) {
                    nl[d][e] += G[a][b][d][e];
                    nm[d][e] += G[a][b][e][d];
                    if (d == e) nm[d][e] += G[a][b][0][0] + G[a][b][1][1] + G[a][b][2][2];
                  }
              }
            }
          }
          for (int d = 0; d < 3; ++d)
            for (int e = 0; e < 3; ++e) {
              A.col[q[d]] = 3 * u + e;
              A.val[q[d]++] = (lam * (double)nl[d][e] + mu * (double)nm[d][e]) / 72.0;
            }
        }
  }
}

// rigid-body modes (3 translations, 3 rotations about the axes through the origin), row-major (3 nn) x 6
void rigid_body_modes(const std::vector<double>& coords, std::vector<double>& B) {
  const int64_t nn = (int64_t)coords.size() / 3;
  B.assign((size_t)(18 * nn), 0.0);
  for (int64_t v = 0; v < nn; ++v) {
    const double x = coords[3 * v], y = coords[3 * v + 1], z = coords[3 * v + 2];
    double* r0 = &B[(size_t)(3 * v) * 6];
    double* r1 = r0 + 6;
    double* r2 = r1 + 6;
    r0[0] = 1.0;
    r1[1] = 1.0;
    r2[2] = 1.0;
    r1[3] = -z;
    r2[3] = y;
    r0[4] = z;
    r2[4] = -x;
    r0[5] = -y;
    r1[5] = x;
  }
}

void matvec(const Csr& A, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < A.nrows; ++i) {
    double s = 0.0;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) s += A.val[k] * x[A.col[k]];
    y[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// structural sparse kernels (explicit zeros are never dropped)
// ------------------------------------------------------------------------------------------
// Row accumulator for the structural products: products are added per column in ENCOUNTER order (the order a
// stable sort by column followed by a sequential sum would use, i.e. the oracle's), then the distinct columns of
// the row are emitted in ascending order.  slot[] maps a column to its position in the row-local arrays.
struct RowAccumulator {
  std::vector<int32_t> slot;
  std::vector<int64_t> cols;
  std::vector<double> vals;
  std::vector<int32_t> order;
  explicit RowAccumulator(int64_t ncols) : slot((size_t)ncols, -1) {}
  void clear() {
    for (int64_t c : cols) slot[c] = -1;
    cols.clear();
    vals.clear();
  }
  void add(int64_t c, double v) {
    const int32_t q = slot[c];
    if (q < 0) {
      slot[c] = (int32_t)cols.size();
      cols.push_back(c);
      vals.push_back(v);
    } else {
      vals[q] += v;
    }
  }
  // positions of the row's columns in ascending column order
  const std::vector<int32_t>& sorted() {
    order.resize(cols.size());
    for (size_t q = 0; q < cols.size(); ++q) order[q] = (int32_t)q;
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return cols[x] < cols[y]; });
    return order;
  }
};

static void spgemm(const Csr& A, const Csr& B, Csr& C) {
  const int64_t n = A.nrows;
  C.nrows = n;
  C.ncols = B.ncols;
  C.ptr.assign(n + 1, 0);
  const int nt = omp_get_max_threads();
  std::vector<std::vector<int64_t>> tcol(nt);
  std::vector<std::vector<double>> tval(nt);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = n * t / nt, r1 = n * (t + 1) / nt;
    RowAccumulator acc(B.ncols);
    auto& oc = tcol[t];
    auto& ov = tval[t];
    for (int64_t i = r0; i < r1; ++i) {
      acc.clear();
      for (int64_t ka = A.ptr[i]; ka < A.ptr[i + 1]; ++ka) {
        const int64_t k = A.col[ka];
        const double a = A.val[ka];
        for (int64_t kb = B.ptr[k]; kb < B.ptr[k + 1]; ++kb) acc.add(B.col[kb], a * B.val[kb]);
      }
      for (int32_t q : acc.sorted()) {
        oc.push_back(acc.cols[q]);
        ov.push_back(acc.vals[q]);
      }
      C.ptr[i + 1] = (int64_t)acc.cols.size();
    }
  }
  for (int64_t i = 0; i < n; ++i) C.ptr[i + 1] += C.ptr[i];
  huge_reserve(C.col, (size_t)C.ptr[n]);
  huge_reserve(C.val, (size_t)C.ptr[n]);
  C.col.resize(C.ptr[n]);
  C.val.resize(C.ptr[n]);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = n * t / nt;
    std::copy(tcol[t].begin(), tcol[t].end(), C.col.begin() + C.ptr[r0]);
    std::copy(tval[t].begin(), tval[t].end(), C.val.begin() + C.ptr[r0]);
  }
}

// C = A*B on the GPU when one is there (same accumulation order, same bits), else on the host cores
static void product(const Csr& A, const Csr& B, Csr& C, bool gpu) {
  if (gpu) {
    try {
      gpu_spgemm(A, B, C);
      return;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "[pamg setup] GPU product failed (%s): falling back to the host\n", e.what());
    }
  }
  spgemm(A, B, C);
}

// T = A^T with ascending columns in every row.  Parallel counting sort: the threads own contiguous row ranges of A, the
// per-thread column counts are turned into per-thread write offsets, so that inside a row of T the entries of a lower
// thread (= lower row ids of A) come first and every thread appends in ascending row order: deterministic and sorted.
static void transpose(const Csr& A, Csr& T) {
  T.nrows = A.ncols;
  T.ncols = A.nrows;
  const int64_t nc = A.ncols, nr = A.nrows;
  T.ptr.assign(nc + 1, 0);
  huge_reserve(T.col, (size_t)A.nnz());
  huge_reserve(T.val, (size_t)A.nnz());
  T.col.resize(A.nnz());
  T.val.resize(A.nnz());
  int nt = omp_get_max_threads();
  // per-thread count arrays cost nt * ncols words: fall back to fewer threads for very wide matrices
  while (nt > 1 && (int64_t)nt * nc > ((int64_t)1 << 28)) nt /= 2;
  std::vector<std::vector<int64_t>> cnt(nt);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = nr * t / nt, r1 = nr * (t + 1) / nt;
    cnt[t].assign(nc, 0);
    for (int64_t k = A.ptr[r0]; k < A.ptr[r1]; ++k) cnt[t][A.col[k]]++;
  }
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nc; ++c) {
    int64_t s = 0;
    for (int t = 0; t < nt; ++t) {
      const int64_t v = cnt[t][c];
      cnt[t][c] = s;  // offset of thread t inside row c of T
      s += v;
    }
    T.ptr[c + 1] = s;
  }
  for (int64_t c = 0; c < nc; ++c) T.ptr[c + 1] += T.ptr[c];
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = nr * t / nt, r1 = nr * (t + 1) / nt;
    std::vector<int64_t>& off = cnt[t];
    for (int64_t i = r0; i < r1; ++i)
      for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int64_t c = A.col[k];
        const int64_t q = T.ptr[c] + off[c]++;
        T.col[q] = i;
        T.val[q] = A.val[k];
      }
  }
}

static void diagonal(const Csr& A, std::vector<double>& d) {
  d.assign(A.nrows, 0.0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < A.nrows; ++i)
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (A.col[k] == i) d[i] = A.val[k];
}

// Gershgorin bound of rho(D^-1 A): max_i (sum_j |a_ij|) * |1/d_i|   (App. B item 4, revised)
static double gershgorin_rho(const Csr& A, const std::vector<double>& dinv) {
  double rho = 0.0;
#pragma omp parallel for schedule(static) reduction(max : rho)
  for (int64_t i = 0; i < A.nrows; ++i) {
    double s = 0.0;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) s += std::fabs(A.val[k]);
    rho = std::max(rho, s * std::fabs(dinv[i]));
  }
  return rho;
}

// ------------------------------------------------------------------------------------------
// own lists
// ------------------------------------------------------------------------------------------
struct OwnIndex {
  std::vector<std::vector<int64_t>> own;  // per part, ascending gid
  std::vector<int32_t> lid;               // gid -> own-local id inside its owner
};

static void build_own_index(const std::vector<int32_t>& owner, int32_t nparts, OwnIndex& oi) {
  const int64_t n = (int64_t)owner.size();
  oi.own.assign(nparts, {});
  huge_reserve(oi.lid, (size_t)n);
  oi.lid.resize(n);
  const int nt = omp_get_max_threads();
  std::vector<std::vector<int64_t>> cnt(nt, std::vector<int64_t>(nparts, 0));
  bool bad = false;
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t g0 = n * t / nt, g1 = n * (t + 1) / nt;
    for (int64_t g = g0; g < g1; ++g) {
      if (owner[g] < 0 || owner[g] >= nparts) {
#pragma omp atomic write
        bad = true;
        continue;
      }
      cnt[t][owner[g]]++;
    }
  }
  if (bad) throw std::runtime_error("owner id out of range");
  for (int32_t p = 0; p < nparts; ++p) {  // counts -> first position of thread t inside part p's own list
    int64_t s = 0;
    for (int t = 0; t < nt; ++t) {
      const int64_t v = cnt[t][p];
      cnt[t][p] = s;
      s += v;
    }
    if (s >= INT32_MAX) throw std::runtime_error("part too large for int32 local ids");
    huge_reserve(oi.own[p], (size_t)s);
    oi.own[p].resize(s);
  }
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t g0 = n * t / nt, g1 = n * (t + 1) / nt;
    std::vector<int64_t>& pos = cnt[t];
    for (int64_t g = g0; g < g1; ++g) {  // ascending gid inside every part, as before
      const int64_t q = pos[owner[g]]++;
      oi.lid[g] = (int32_t)q;
      oi.own[owner[g]][q] = g;
    }
  }
}

// ------------------------------------------------------------------------------------------
// aggregation (greedy, 3 passes, per part, ascending local row id; App. B item 2)
// ------------------------------------------------------------------------------------------
static int64_t aggregate_part(const Csr& A, const std::vector<int32_t>& owner, const OwnIndex& oi, int32_t p,
                              double eps, const std::vector<double>& absdiag, std::vector<int32_t>& agg) {
  const auto& own = oi.own[p];
  const int64_t n = (int64_t)own.size();
  agg.assign(n, -1);
  auto strong = [&](int64_t i, int64_t k) {
    const int64_t j = A.col[k];
    if (j == i || owner[j] != p) return false;
    if (eps <= 0.0) return true;
    return std::fabs(A.val[k]) > eps * std::sqrt(absdiag[i] * absdiag[j]);
  };
  int32_t nagg = 0;
  for (int64_t li = 0; li < n; ++li) {  // pass 1
    if (agg[li] != -1) continue;
    const int64_t i = own[li];
    bool ok = true;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1] && ok; ++k)
      if (strong(i, k) && agg[oi.lid[A.col[k]]] != -1) ok = false;
    if (!ok) continue;
    agg[li] = nagg;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k)) agg[oi.lid[A.col[k]]] = nagg;
    ++nagg;
  }
  std::vector<int32_t> agg2(agg);
  for (int64_t li = 0; li < n; ++li) {  // pass 2: first pass-1-aggregated strong neighbour
    if (agg[li] != -1) continue;
    const int64_t i = own[li];
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k) && agg[oi.lid[A.col[k]]] != -1) {
        agg2[li] = agg[oi.lid[A.col[k]]];
        break;
      }
  }
  agg.swap(agg2);
  for (int64_t li = 0; li < n; ++li) {  // pass 3
    if (agg[li] != -1) continue;
    const int64_t i = own[li];
    agg[li] = nagg;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (strong(i, k) && agg[oi.lid[A.col[k]]] == -1) agg[oi.lid[A.col[k]]] = nagg;
    ++nagg;
  }
  return nagg;
}

// every part: on the device when the matrix (or node graph) is resident there (setup_gpu.cu gpu_aggregate:
// same aggregates bit for bit), else one part per thread on the host.  agg_gid[g] = aggregate id inside the owner part.
static void aggregate_all(const Csr& M, const GpuMat* dM, const std::vector<int32_t>& owner, const OwnIndex& oi, int32_t nparts,
                          double eps, const std::vector<double>& absdiag, std::vector<int32_t>& agg_gid,
                          std::vector<int64_t>& counts, bool* on_gpu) {
  const int64_t n = M.nrows;
  huge_reserve(agg_gid, (size_t)n);
  agg_gid.resize(n);
  counts.assign(nparts, 0);
  *on_gpu = false;
  const char* e = std::getenv("PAMG_GPU_AGG");  // 0: host walk; 2: the device or an error (tests); default: device, host as fallback
  const int mode = e ? std::atoi(e) : 1;
  std::string why = "the matrix is not resident on a device";
  if (dM && mode != 0) {
    try {
      std::vector<int64_t> part_off(nparts + 1, 0);
      for (int32_t p = 0; p < nparts; ++p) part_off[p + 1] = part_off[p] + (int64_t)oi.own[p].size();
      std::vector<int32_t> pos(n);
#pragma omp parallel for schedule(static)
      for (int64_t g = 0; g < n; ++g) pos[g] = (int32_t)(part_off[owner[g]] + oi.lid[g]);
      if (gpu_aggregate(dM, owner.data(), pos.data(), part_off.data(), nparts, eps, mode == 2, agg_gid.data(), counts.data())) {
        *on_gpu = true;
        return;
      }
      why = "the strength graph is not symmetric, or its dependency chain is too long";
    } catch (const std::exception& ex) {
      why = ex.what();
      std::fprintf(stderr, "[pamg setup] GPU aggregation failed (%s): falling back to the host\n", ex.what());
    }
  }
  if (mode == 2) throw std::runtime_error("PAMG_GPU_AGG=2 but the aggregation could not run on the device: " + why);
  std::vector<std::vector<int32_t>> aggs(nparts);
  OmpGuard og;
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t p = 0; p < nparts; ++p) og.run([&] { counts[p] = aggregate_part(M, owner, oi, p, eps, absdiag, aggs[p]); });
  og.rethrow();
  for (int32_t p = 0; p < nparts; ++p) {
    const auto& own = oi.own[p];
    const auto& a = aggs[p];
#pragma omp parallel for schedule(static)
    for (int64_t li = 0; li < (int64_t)own.size(); ++li) agg_gid[own[li]] = a[li];
  }
}

// ------------------------------------------------------------------------------------------
// smoothed prolongator P = (I - w D_F^-1 A_F) P0, w = 4/(3 rho_F)       (App. B items 3-4)
// ------------------------------------------------------------------------------------------
// P0: tentative prolongator (n x nc, sorted columns): one unit entry per row for scalar problems, the
// per-aggregate Q factors of the near-nullspace for block problems.
// dA: the device copy of A when the GPU chain is on, else nullptr.  With it the whole smoothing runs on the device
// (S = A_F P0, merge) and *dP_out keeps the smoothed prolongator resident for the transpose and the Galerkin product.
static void build_prolongator(const Csr& A, const Csr& P0, int64_t nc, double eps,
                              const std::vector<double>& absdiag, Csr& P, double* omega_out, bool gpu, const GpuMat* dA,
                              GpuMat** dP_out) {
  const int64_t n = A.nrows;
  // filtered matrix A_F (weak off-diagonals lumped into the diagonal); eps == 0 => A_F = A
  Csr AF;
  const Csr* F = &A;
  if (eps > 0.0) {
    AF.nrows = AF.ncols = n;
    AF.ptr.assign(n + 1, 0);
    std::vector<double> dnew(n, 0.0);
    auto is_strong = [&](int64_t i, int64_t k) {
      const int64_t j = A.col[k];
      return j == i || std::fabs(A.val[k]) > eps * std::sqrt(absdiag[i] * absdiag[j]);
    };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      int64_t cnt = 0;
      double s = 0.0, dii = 0.0;
      for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        if (A.col[k] == i) dii = A.val[k];
        if (is_strong(i, k))
          ++cnt;
        else
          s += A.val[k];
      }
      // weak entries are lumped into the diagonal; the lumped VALUE is clamped to at least half of
      // a_ii so that it cannot be cancelled on coarse levels (the pattern never depends on it)
      const double dl = dii + s;
      const double dc = dii > 0 ? std::max(dl, 0.5 * dii) : std::min(dl, 0.5 * dii);
      AF.ptr[i + 1] = cnt;
      dnew[i] = dii + (dc - dii);  // same arithmetic as the oracle's A_F + diags(dF_new - d0)
    }
    for (int64_t i = 0; i < n; ++i) AF.ptr[i + 1] += AF.ptr[i];
    huge_reserve(AF.col, (size_t)AF.ptr[n]);
    huge_reserve(AF.val, (size_t)AF.ptr[n]);
    AF.col.resize(AF.ptr[n]);
    AF.val.resize(AF.ptr[n]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      int64_t q = AF.ptr[i];
      for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
        const int64_t j = A.col[k];
        if (j == i) {
          AF.col[q] = j;
          AF.val[q++] = dnew[i];
        } else if (is_strong(i, k)) {
          AF.col[q] = j;
          AF.val[q++] = A.val[k];
        }
      }
    }
    F = &AF;
  }
  std::vector<double> dF, dinv(n);
  diagonal(*F, dF);
  for (int64_t i = 0; i < n; ++i) {
    if (dF[i] == 0.0) throw std::runtime_error("zero diagonal in (filtered) matrix");
    dinv[i] = 1.0 / dF[i];
  }
  const double rho = gershgorin_rho(*F, dinv);
  const double omega = 4.0 / (3.0 * rho);
  *omega_out = omega;

  // S = A_F P0 (structural product, encounter-order sums), then P = P0 - omega D_F^-1 S merged by column
  if (gpu && dA && dP_out) {  // device chain: only P0, the weights (and a filtered A_F) go up, only P comes down
    GpuMat *dP0 = nullptr, *dF = nullptr, *dP = nullptr;
    bool ok = false;
    try {
      std::vector<double> w(n);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) w[i] = -(omega * dinv[i]);
      dP0 = gpu_upload(P0);
      if (F != &A) dF = gpu_upload(*F);
      dP = gpu_smooth_prolongator(dF ? dF : dA, dP0, w.data());
      gpu_download(dP, P);
      ok = true;
    } catch (const std::exception& e) {
      std::fprintf(stderr, "[pamg setup] GPU prolongator smoothing failed (%s): falling back to the host\n", e.what());
      gpu_free(dP);
      dP = nullptr;
    }
    gpu_free(dP0);
    gpu_free(dF);
    if (ok) {
      *dP_out = dP;
      return;
    }
  }
  Csr S;
  product(*F, P0, S, gpu);
  P.nrows = n;
  P.ncols = nc;
  P.ptr.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int64_t pk = P0.ptr[i], cnt = 0;
    const int64_t pe = P0.ptr[i + 1];
    for (int64_t k = S.ptr[i]; k < S.ptr[i + 1]; ++k) {
      const int64_t c = S.col[k];
      while (pk < pe && P0.col[pk] < c) {  // tentative entries without an A_F P0 partner
        ++pk;
        ++cnt;
      }
      if (pk < pe && P0.col[pk] == c) ++pk;
      ++cnt;
    }
    P.ptr[i + 1] = cnt + (pe - pk);
  }
  for (int64_t i = 0; i < n; ++i) P.ptr[i + 1] += P.ptr[i];
  huge_reserve(P.col, (size_t)P.ptr[n]);
  huge_reserve(P.val, (size_t)P.ptr[n]);
  P.col.resize(P.ptr[n]);
  P.val.resize(P.ptr[n]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double w = -(omega * dinv[i]);
    int64_t pk = P0.ptr[i], q = P.ptr[i];
    const int64_t pe = P0.ptr[i + 1];
    for (int64_t k = S.ptr[i]; k < S.ptr[i + 1]; ++k) {
      const int64_t c = S.col[k];
      while (pk < pe && P0.col[pk] < c) {
        P.col[q] = P0.col[pk];
        P.val[q++] = P0.val[pk++];
      }
      double v = w * S.val[k];
      if (pk < pe && P0.col[pk] == c) v = P0.val[pk++] + v;
      P.col[q] = c;
      P.val[q++] = v;
    }
    while (pk < pe) {
      P.col[q] = P0.col[pk];
      P.val[q++] = P0.val[pk++];
    }
  }
}

// ------------------------------------------------------------------------------------------
// near-nullspace tentative prolongator (oracle/amg_oracle.py householder_qr / tentative_from_nullspace)
// ------------------------------------------------------------------------------------------
// Thin Householder QR of the m x k row-major block W (overwritten): Q (m x k, row-major) and R (k x k).
// diag(R) >= 0; a sub-column whose norm is <= 1e-12 of the original column's norm counts as zero (no
// reflector); m < k pads Q with zero columns and R with zero rows.
static void householder_qr(std::vector<double>& W, int64_t m, int k, std::vector<double>& Q, std::vector<double>& R) {
  const int r = (int)std::min<int64_t>(m, k);
  std::vector<double> cn(k, 0.0);
  for (int j = 0; j < k; ++j) {
    double s = 0.0;
    for (int64_t i = 0; i < m; ++i) s += W[i * k + j] * W[i * k + j];
    cn[j] = std::sqrt(s);
  }
  std::vector<std::vector<double>> V(r);
  std::vector<double> vn2(r, 0.0);
  for (int j = 0; j < r; ++j) {
    double s = 0.0;
    for (int64_t i = j; i < m; ++i) s += W[i * k + j] * W[i * k + j];
    const double alpha = std::sqrt(s);
    if (alpha <= 1e-12 * cn[j] || cn[j] == 0.0) {
      for (int64_t i = j; i < m; ++i) W[i * k + j] = 0.0;
      continue;  // V[j] stays empty: H_j = I
    }
    std::vector<double>& v = V[j];
    v.resize(m - j);
    for (int64_t i = j; i < m; ++i) v[i - j] = W[i * k + j];
    const double beta = v[0] >= 0.0 ? -alpha : alpha;
    v[0] -= beta;
    double n2 = 0.0;
    for (double t : v) n2 += t * t;
    vn2[j] = n2;
    for (int c = j; c < k; ++c) {
      double dot = 0.0;
      for (int64_t i = j; i < m; ++i) dot += v[i - j] * W[i * k + c];
      const double f = (2.0 / n2) * dot;
      for (int64_t i = j; i < m; ++i) W[i * k + c] -= v[i - j] * f;
    }
    for (int64_t i = j + 1; i < m; ++i) W[i * k + j] = 0.0;
  }
  R.assign((size_t)k * k, 0.0);
  for (int i = 0; i < r; ++i)
    for (int c = i; c < k; ++c) R[(size_t)i * k + c] = W[(size_t)i * k + c];
  Q.assign((size_t)m * k, 0.0);
  for (int i = 0; i < r; ++i) Q[(size_t)i * k + i] = 1.0;
  for (int j = r - 1; j >= 0; --j) {
    if (V[j].empty()) continue;
    const std::vector<double>& v = V[j];
    for (int c = 0; c < r; ++c) {
      double dot = 0.0;
      for (int64_t i = j; i < m; ++i) dot += v[i - j] * Q[i * k + c];
      const double f = (2.0 / vn2[j]) * dot;
      for (int64_t i = j; i < m; ++i) Q[i * k + c] -= v[i - j] * f;
    }
  }
  for (int j = 0; j < r; ++j)
    if (R[(size_t)j * k + j] < 0.0) {
      for (int c = 0; c < k; ++c) R[(size_t)j * k + c] = -R[(size_t)j * k + c];
      for (int64_t i = 0; i < m; ++i) Q[i * k + j] = -Q[i * k + j];
    }
}

// P0 (n x k n_agg; every row of an aggregate stores all k entries), coarse near-nullspace Bc (k n_agg x k)
// and the coarse DOFs whose column is empty (aggregates with fewer rows than k)
static void tentative_from_nullspace(const std::vector<double>& B, int k, int bs, const std::vector<int64_t>& agg_node,
                                     int64_t n_agg, Csr& P0, std::vector<double>& Bc, std::vector<int64_t>& dead) {
  const int64_t nn = (int64_t)agg_node.size(), n = nn * bs;
  std::vector<int64_t> start(n_agg + 1, 0);
  for (int64_t v = 0; v < nn; ++v) start[agg_node[v] + 1]++;
  for (int64_t g = 0; g < n_agg; ++g) start[g + 1] += start[g];
  std::vector<int64_t> members(nn), pos(start.begin(), start.end() - 1);
  for (int64_t v = 0; v < nn; ++v) members[pos[agg_node[v]]++] = v;  // ascending node id inside each aggregate
  P0.nrows = n;
  P0.ncols = k * n_agg;
  P0.ptr.resize(n + 1);
  for (int64_t i = 0; i <= n; ++i) P0.ptr[i] = i * k;
  P0.col.resize((size_t)n * k);
  P0.val.resize((size_t)n * k);
  Bc.assign((size_t)n_agg * k * k, 0.0);
  dead.clear();
  std::vector<char> is_dead((size_t)n_agg * k, 0);
#pragma omp parallel
  {
    std::vector<double> W, Q, R;
#pragma omp for schedule(dynamic, 64)
    for (int64_t g = 0; g < n_agg; ++g) {
      const int64_t m = (start[g + 1] - start[g]) * bs;
      W.resize((size_t)m * k);
      for (int64_t a = start[g]; a < start[g + 1]; ++a)
        for (int d = 0; d < bs; ++d) {
          const int64_t dof = members[a] * bs + d, row = (a - start[g]) * bs + d;
          for (int c = 0; c < k; ++c) W[(size_t)row * k + c] = B[(size_t)dof * k + c];
        }
      householder_qr(W, m, k, Q, R);
      for (int c = 0; c < k * k; ++c) Bc[(size_t)g * k * k + c] = R[c];
      for (int64_t a = start[g]; a < start[g + 1]; ++a)
        for (int d = 0; d < bs; ++d) {
          const int64_t dof = members[a] * bs + d, row = (a - start[g]) * bs + d;
          for (int c = 0; c < k; ++c) {
            P0.col[(size_t)dof * k + c] = g * k + c;
            P0.val[(size_t)dof * k + c] = Q[(size_t)row * k + c];
          }
        }
      for (int64_t c = m; c < k; ++c) is_dead[(size_t)g * k + c] = 1;
    }
  }
  for (int64_t c = 0; c < n_agg * k; ++c)
    if (is_dead[c]) dead.push_back(c);
}

// pattern of the bs x bs blocks of A as a node-level matrix (values unused)
static void node_graph(const Csr& A, int bs, Csr& N) {
  const int64_t nn = A.nrows / bs;
  N.nrows = N.ncols = nn;
  N.ptr.assign(nn + 1, 0);
  std::vector<std::vector<int64_t>> rows(nn);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < nn; ++v) {
    std::vector<int64_t>& r = rows[v];
    for (int d = 0; d < bs; ++d)
      for (int64_t k = A.ptr[v * bs + d]; k < A.ptr[v * bs + d + 1]; ++k) r.push_back(A.col[k] / bs);
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
  }
  for (int64_t v = 0; v < nn; ++v) N.ptr[v + 1] = N.ptr[v] + (int64_t)rows[v].size();
  N.col.resize(N.ptr[nn]);
  N.val.assign(N.ptr[nn], 1.0);
  for (int64_t v = 0; v < nn; ++v) std::copy(rows[v].begin(), rows[v].end(), N.col.begin() + N.ptr[v]);
}

// ------------------------------------------------------------------------------------------
// dense inverse of the coarsest matrix (Gauss-Jordan, partial pivoting)     (App. B item 6)
// ------------------------------------------------------------------------------------------
static void dense_inverse(const Csr& A, std::vector<double>& inv) {
  const int64_t n = A.nrows;
  if (n > 8192) throw std::runtime_error("coarsest level too large for the dense solve (raise max_levels)");
  std::vector<double> M(n * n, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) M[i * n + A.col[k]] = A.val[k];
  inv.assign(n * n, 0.0);
  for (int64_t i = 0; i < n; ++i) inv[i * n + i] = 1.0;
  for (int64_t c = 0; c < n; ++c) {
    int64_t piv = c;
    double best = std::fabs(M[c * n + c]);
    for (int64_t r = c + 1; r < n; ++r)
      if (std::fabs(M[r * n + c]) > best) {
        best = std::fabs(M[r * n + c]);
        piv = r;
      }
    if (best == 0.0) throw std::runtime_error("singular coarsest matrix");
    if (piv != c)
      for (int64_t j = 0; j < n; ++j) {
        std::swap(M[c * n + j], M[piv * n + j]);
        std::swap(inv[c * n + j], inv[piv * n + j]);
      }
    const double d = 1.0 / M[c * n + c];
    for (int64_t j = 0; j < n; ++j) {
      M[c * n + j] *= d;
      inv[c * n + j] *= d;
    }
#pragma omp parallel for schedule(static) if (n > 256)
    for (int64_t r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[r * n + c];
      if (f == 0.0) continue;
      for (int64_t j = 0; j < n; ++j) {
        M[r * n + j] -= f * M[c * n + j];
        inv[r * n + j] -= f * inv[c * n + j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// localisation: index maps, split blocks, halo plans
// ------------------------------------------------------------------------------------------
struct GhostLookup {  // gid -> ghost slot of one part
  std::vector<std::pair<int64_t, int32_t>> sorted;
  void build(const std::vector<int64_t>& g2g) {
    sorted.resize(g2g.size());
    for (size_t s = 0; s < g2g.size(); ++s) sorted[s] = {g2g[s], (int32_t)s};
    std::sort(sorted.begin(), sorted.end());
  }
  int32_t find(int64_t gid) const {
    auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(gid, (int32_t)-1));
    if (it == sorted.end() || it->first != gid) return -1;
    return it->second;
  }
};

static void index_maps(const Csr& A, const std::vector<int32_t>& owner, const OwnIndex& oi, int32_t p, PartLevel& pl) {
  pl.present = true;
  pl.own_to_global = oi.own[p];
  pl.n_own = (int64_t)pl.own_to_global.size();
  std::vector<std::pair<int32_t, int64_t>> gh;
  {  // (inside a parallel loop over parts this region is serialised; with fewer parts than threads it is the parallel one)
    const int64_t no = pl.n_own;
#pragma omp parallel
    {
      std::vector<std::pair<int32_t, int64_t>> loc;
#pragma omp for schedule(static) nowait
      for (int64_t r = 0; r < no; ++r) {
        const int64_t i = pl.own_to_global[r];
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
          if (owner[A.col[k]] != p) loc.emplace_back(owner[A.col[k]], A.col[k]);
      }
      std::sort(loc.begin(), loc.end());
      loc.erase(std::unique(loc.begin(), loc.end()), loc.end());
#pragma omp critical(pamg_index_maps)
      gh.insert(gh.end(), loc.begin(), loc.end());
    }
  }
  std::sort(gh.begin(), gh.end());
  gh.erase(std::unique(gh.begin(), gh.end()), gh.end());
  pl.n_ghost = (int64_t)gh.size();
  pl.ghost_to_global.resize(gh.size());
  pl.ghost_to_owner.resize(gh.size());
  for (size_t s = 0; s < gh.size(); ++s) {
    pl.ghost_to_owner[s] = gh[s].first;
    pl.ghost_to_global[s] = gh[s].second;
  }
}

// rows `rows` (gids) of global M -> (oo, og) with the column partition (owner_c, lid_c, ghosts of p)
static void split_blocks(const Csr& M, const std::vector<int64_t>& rows, const std::vector<int32_t>& owner_c,
                         const std::vector<int32_t>& lid_c, int32_t p, int64_t n_own_c, int64_t n_ghost_c,
                         const GhostLookup& gl, LocalCsr& oo, LocalCsr& og) {
  const int64_t nr = (int64_t)rows.size();
  oo.nrows = og.nrows = nr;
  oo.ncols = n_own_c;
  og.ncols = n_ghost_c;
  huge_reserve(oo.ptr, (size_t)nr + 1);
  oo.ptr.assign(nr + 1, 0);
  og.ptr.assign(nr + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nr; ++r) {
    const int64_t i = rows[r];
    int64_t a = 0, b = 0;
    for (int64_t k = M.ptr[i]; k < M.ptr[i + 1]; ++k) (owner_c[M.col[k]] == p ? a : b)++;
    oo.ptr[r + 1] = a;
    og.ptr[r + 1] = b;
  }
  for (int64_t r = 0; r < nr; ++r) {
    oo.ptr[r + 1] += oo.ptr[r];
    og.ptr[r + 1] += og.ptr[r];
  }
  if (oo.ptr[nr] > INT32_MAX || og.ptr[nr] > INT32_MAX) throw std::runtime_error("block nnz exceeds int32");
  huge_reserve(oo.col, (size_t)oo.ptr[nr]);
  huge_reserve(oo.val, (size_t)oo.ptr[nr]);
  oo.col.resize(oo.ptr[nr]);
  oo.val.resize(oo.ptr[nr]);
  og.col.resize(og.ptr[nr]);
  og.val.resize(og.ptr[nr]);
  bool missing = false;
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, double>> tmp;
#pragma omp for schedule(static)
    for (int64_t r = 0; r < nr; ++r) {
      const int64_t i = rows[r];
      int64_t qa = oo.ptr[r];
      tmp.clear();
      for (int64_t k = M.ptr[i]; k < M.ptr[i + 1]; ++k) {
        const int64_t j = M.col[k];
        if (owner_c[j] == p) {
          oo.col[qa] = lid_c[j];
          oo.val[qa++] = M.val[k];
        } else {
          const int32_t sl = gl.find(j);
          if (sl < 0) {
#pragma omp atomic write
            missing = true;
            continue;
          }
          tmp.emplace_back(sl, M.val[k]);
        }
      }
      std::sort(tmp.begin(), tmp.end(), [](const std::pair<int32_t, double>& x, const std::pair<int32_t, double>& y) { return x.first < y.first; });
      int64_t qb = og.ptr[r];
      for (auto& e : tmp) {
        og.col[qb] = e.first;
        og.val[qb++] = e.second;
      }
    }
  }
  if (missing) throw std::runtime_error("column outside the part's own+ghost set");
}

void build_halo_plans(Level& lev, int32_t nparts) {
  for (int32_t p = 0; p < nparts; ++p) {
    lev.parts[p].recv.clear();
    lev.parts[p].send.clear();
    lev.parts[p].send_idx.clear();
  }
  for (int32_t q = 0; q < nparts; ++q) {
    PartLevel& Q = lev.parts[q];
    int64_t s = 0;
    while (s < Q.n_ghost) {
      const int32_t p = Q.ghost_to_owner[s];
      int64_t e = s;
      while (e < Q.n_ghost && Q.ghost_to_owner[e] == p) ++e;
      if (p < 0 || p >= nparts || p == q) throw std::runtime_error("bad ghost owner");
      if (!Q.recv.empty() && Q.recv.back().part >= p) throw std::runtime_error("ghosts not sorted by (owner, gid)");
      Q.recv.push_back({p, (int32_t)s, (int32_t)(e - s), 0});
      PartLevel& Pp = lev.parts[p];
      Neighbor nb{q, (int32_t)s, (int32_t)(e - s), (int64_t)Pp.send_idx.size()};
      for (int64_t t = s; t < e; ++t) {
        const int64_t gid = Q.ghost_to_global[t];
        auto it = std::lower_bound(Pp.own_to_global.begin(), Pp.own_to_global.end(), gid);
        if (it == Pp.own_to_global.end() || *it != gid) throw std::runtime_error("ghost gid not owned by its owner");
        Pp.send_idx.push_back((int32_t)(it - Pp.own_to_global.begin()));
      }
      Pp.send.push_back(nb);
      s = e;
    }
  }
}

static void fill_diag(PartLevel& pl) {
  const LocalCsr& oo = pl.blk[PAMG_A_OO];
  const LocalCsr& og = pl.blk[PAMG_A_OG];
  pl.diag.assign(pl.n_own, 0.0);
  pl.diag_l1.assign(pl.n_own, 0.0);
  bool zero = false;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < pl.n_own; ++i) {
    for (int64_t k = oo.ptr[i]; k < oo.ptr[i + 1]; ++k)
      if (oo.col[k] == i) pl.diag[i] = oo.val[k];
    double s = 0.0;
    if (!og.ptr.empty())
      for (int64_t k = og.ptr[i]; k < og.ptr[i + 1]; ++k) s += std::fabs(og.val[k]);
    pl.diag_l1[i] = pl.diag[i] + s;
    if (pl.diag[i] == 0.0) {
#pragma omp atomic write
      zero = true;
    }
  }
  if (zero) throw std::runtime_error("zero diagonal");
}

// ------------------------------------------------------------------------------------------
// whole setup
// ------------------------------------------------------------------------------------------
namespace {
struct PhaseTimer {  // PAMG_SETUP_TIMING=1: per-phase wall time of the host setup on stderr
  bool on = std::getenv("PAMG_SETUP_TIMING") != nullptr;
  double t0 = omp_get_wtime();
  void lap(const char* what, int level) {
    if (!on) return;
    const double t1 = omp_get_wtime();
    std::fprintf(stderr, "[pamg setup] level %d %-22s %8.3f s\n", level, what, t1 - t0);
    t0 = t1;
  }
};
}  // namespace

void build_hierarchy(const Csr& A0, const std::vector<int32_t>& owner0, int32_t nparts, const pamg_options& o, Hierarchy& h,
                     int32_t block_size, int32_t ns_k, const std::vector<double>* nullspace) {
  PhaseTimer tm;
  bool gpu = gpu_setup_available();  // the three sparse products of every level run on the GPU when one is there
  GpuMat* dA = nullptr;              // device copy of the current level's matrix (GPU chain)
  struct DaGuard {                   // an exception below must not leak device memory or pinned staging
    GpuMat*& p;
    bool staging;
    ~DaGuard() {
      gpu_free(p);
      if (staging) gpu_setup_end();
    }
  } da_guard{dA, gpu};
  if (gpu) gpu_setup_begin();
  if (tm.on) std::fprintf(stderr, "[pamg setup] sparse products on the %s\n", gpu ? "GPU" : "host");
  if (A0.nrows != (int64_t)owner0.size()) throw std::runtime_error("owner size mismatch");
  const bool use_ns = nullspace && ns_k > 0;
  if (block_size < 1) block_size = 1;
  if (block_size > 1 && !use_ns) throw std::runtime_error("block_size > 1 needs a near-nullspace (pamg_set_near_nullspace)");
  if (use_ns && ((int64_t)nullspace->size() != A0.nrows * ns_k || A0.nrows % block_size))
    throw std::runtime_error("near-nullspace / block size do not match the matrix");
  std::vector<double> Bcur;  // near-nullspace of the current level (row-major n x ns_k)
  if (use_ns) Bcur = *nullspace;
  int bs = block_size;
  h = Hierarchy();
  h.nparts = nparts;
  h.opts = o;

  struct G {
    Csr A_own, P, R;        // A_own: the Galerkin matrix of a coarse level (level 0 borrows the caller's matrix: no 2 GB copy)
    const Csr* Ap = nullptr;
    const Csr& A() const { return *Ap; }
    std::vector<int32_t> owner;
    OwnIndex oi;
    std::vector<int32_t> agg_loc;  // by gid
    double rho = 0, omega_p = 0;
  };
  std::vector<G> g(1);
  g.reserve((size_t)std::max(o.max_levels, 1) + 1);  // no reallocation: G::Ap points into the elements
  g[0].Ap = &A0;
  g[0].owner = owner0;
  while (true) {
    const double eps = o.eps_strength * std::pow(0.5, (double)(g.size() - 1));  // Vanek: eps_l = eps 2^-l
    G& cur = g.back();
    const int64_t n = cur.A().nrows;
    if (gpu && !dA && n > o.coarse_size && (int32_t)g.size() < o.max_levels) {
      try {
        dA = gpu_upload(cur.A());
      } catch (const std::exception& e) {
        std::fprintf(stderr, "[pamg setup] GPU upload failed (%s): sparse products on the host\n", e.what());
        gpu = false;
      }
    }
    build_own_index(cur.owner, nparts, cur.oi);
    std::vector<double> d, dinv(n), absdiag(n);
    diagonal(cur.A(), d);
    bool zero_diag = false;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      if (d[i] == 0.0) {
#pragma omp atomic write
        zero_diag = true;
        continue;
      }
      dinv[i] = 1.0 / d[i];
      absdiag[i] = std::fabs(d[i]);
    }
    if (zero_diag) throw std::runtime_error("zero diagonal");
    cur.rho = gershgorin_rho(cur.A(), dinv);
    tm.lap("own index + diagonal", (int)g.size() - 1);
    if (n <= o.coarse_size || (int32_t)g.size() >= o.max_levels) break;

    std::vector<int32_t> agg_gid;  // aggregate id inside the owner part, by row (scalar) or node (near-nullspace) gid
    std::vector<int64_t> counts(nparts, 0);
    std::vector<int64_t> off(nparts + 1, 0);
    Csr P0;
    std::vector<int64_t> dead;
    std::vector<double> Bc;
    int64_t nc = 0;
    int kdof = 1;
    bool agg_on_gpu = false;
    huge_reserve(cur.agg_loc, (size_t)n);
    cur.agg_loc.resize(n);
    if (!use_ns) {
      aggregate_all(cur.A(), dA, cur.owner, cur.oi, nparts, eps, absdiag, agg_gid, counts, &agg_on_gpu);
      for (int32_t p = 0; p < nparts; ++p) off[p + 1] = off[p] + counts[p];
      nc = off[nparts];
      if (nc >= n) break;
      P0.nrows = n;
      P0.ncols = nc;
      huge_reserve(P0.ptr, (size_t)n + 1);
      huge_reserve(P0.col, (size_t)n);
      huge_reserve(P0.val, (size_t)n);
      P0.ptr.resize(n + 1);
      P0.col.resize(n);
      P0.val.assign(n, 1.0);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i <= n; ++i) {
        P0.ptr[i] = i;
        if (i < n) {
          cur.agg_loc[i] = agg_gid[i];
          P0.col[i] = off[cur.owner[i]] + agg_gid[i];
        }
      }
    } else {
      // nodes (bs DOFs each) are aggregated on the block pattern; tentative P by per-aggregate QR of B
      if (eps > 0.0) throw std::runtime_error("block strength thresholds are not defined: use eps_strength = 0");
      kdof = ns_k;
      const int64_t nn = n / bs;
      std::vector<int32_t> owner_node(nn);
      for (int64_t v = 0; v < nn; ++v) {
        owner_node[v] = cur.owner[v * bs];
        for (int d = 1; d < bs; ++d)
          if (cur.owner[v * bs + d] != owner_node[v]) throw std::runtime_error("the DOFs of a node must share one owner");
      }
      Csr N;
      node_graph(cur.A(), bs, N);
      OwnIndex oin;
      build_own_index(owner_node, nparts, oin);
      std::vector<double> none;
      GpuMat* dN = nullptr;
      if (gpu && dA) {
        try {
          dN = gpu_upload(N);
        } catch (const std::exception&) {
          dN = nullptr;
        }
      }
      try {
        aggregate_all(N, dN, owner_node, oin, nparts, 0.0, none, agg_gid, counts, &agg_on_gpu);
      } catch (...) {
        gpu_free(dN);
        throw;
      }
      gpu_free(dN);
      for (int32_t p = 0; p < nparts; ++p) off[p + 1] = off[p] + counts[p];
      const int64_t nagg = off[nparts];
      nc = nagg * kdof;
      if (nc >= n) break;
      std::vector<int64_t> agg_node(nn);
#pragma omp parallel for schedule(static)
      for (int64_t v = 0; v < nn; ++v) {
        agg_node[v] = off[owner_node[v]] + agg_gid[v];
        for (int d = 0; d < bs; ++d) cur.agg_loc[v * bs + d] = agg_gid[v];
      }
      tentative_from_nullspace(Bcur, kdof, bs, agg_node, nagg, P0, Bc, dead);
    }
    tm.lap(agg_on_gpu ? "aggregation (GPU) + P0" : "aggregation + P0", (int)g.size() - 1);
    GpuMat *dP = nullptr, *dR = nullptr;  // resident copies of P and R (GPU chain)
    build_prolongator(cur.A(), P0, nc, eps, absdiag, cur.P, &cur.omega_p, gpu, dA, &dP);
    tm.lap(dP ? "prolongator smoothing (GPU)" : "prolongator smoothing", (int)g.size() - 1);
    if (dP) {
      try {
        dR = gpu_transpose(dP);
        gpu_download(dR, cur.R);
      } catch (const std::exception& e) {
        std::fprintf(stderr, "[pamg setup] GPU transpose failed (%s): falling back to the host\n", e.what());
        gpu_free(dR);
        dR = nullptr;
      }
    }
    if (!dR) transpose(cur.P, cur.R);
    tm.lap(dR ? "transpose (GPU)" : "transpose", (int)g.size() - 1);
    G nxt;
    bool galerkin_done = false;
    GpuMat* dAc = nullptr;
    if (gpu && dA) {  // A (and usually P, R) resident: A*P stays on the device, only A_c comes down
      GpuMat* dAP = nullptr;
      try {
        if (!dP) dP = gpu_upload(cur.P);
        dAP = gpu_product(dA, dP);
        tm.lap("A*P (GPU)", (int)g.size() - 1);
        if (!dR) dR = gpu_upload(cur.R);
        dAc = gpu_product(dR, dAP);
        gpu_download(dAc, nxt.A_own);
        tm.lap("R*(AP) (GPU)", (int)g.size() - 1);
        galerkin_done = true;
      } catch (const std::exception& e) {
        std::fprintf(stderr, "[pamg setup] GPU Galerkin product failed (%s): falling back to the host\n", e.what());
        gpu_free(dAc);
        dAc = nullptr;
      }
      gpu_free(dAP);
    }
    gpu_free(dP);
    gpu_free(dR);
    gpu_free(dA);  // the fine matrix is not needed on the device any more
    dA = nullptr;
    if (!galerkin_done) {
      Csr AP;
      product(cur.A(), cur.P, AP, gpu);
      tm.lap("A*P", (int)g.size() - 1);
      product(cur.R, AP, nxt.A_own, gpu);
      tm.lap("R*(AP)", (int)g.size() - 1);
    }
    for (int64_t gd : dead) {  // empty coarse column: unit diagonal keeps the Galerkin matrix regular
      bool found = false;
      for (int64_t k = nxt.A_own.ptr[gd]; k < nxt.A_own.ptr[gd + 1]; ++k)
        if (nxt.A_own.col[k] == gd) {
          nxt.A_own.val[k] = 1.0;
          found = true;
        }
      if (!found) throw std::runtime_error("internal: no diagonal slot for an empty coarse column");
    }
    if (!dead.empty() && dAc) {  // the host copy was edited: the device copy is stale
      gpu_free(dAc);
      dAc = nullptr;
    }
    dA = dAc;  // next level's A, already on the device (or nullptr: uploaded at the top of the loop)
    nxt.owner.resize(nc);
    for (int32_t p = 0; p < nparts; ++p)
      for (int64_t c = off[p] * kdof; c < off[p + 1] * kdof; ++c) nxt.owner[c] = p;
    g.push_back(std::move(nxt));
    g.back().Ap = &g.back().A_own;
    if (use_ns) {
      Bcur.swap(Bc);
      bs = kdof;
    }
  }

  const int32_t L = (int32_t)g.size();
  tm.lap("(levels done)", L - 1);
  h.levels.resize(L);
  std::vector<std::vector<GhostLookup>> gl(L, std::vector<GhostLookup>(nparts));
  for (int32_t l = 0; l < L; ++l) {
    Level& lev = h.levels[l];
    lev.n_global = g[l].A().nrows;
    lev.rho = g[l].rho;
    lev.omega_p = g[l].omega_p;
    lev.parts.resize(nparts);
    OmpGuard og;
    // few parts: the loops INSIDE index_maps / split_blocks are the parallel ones (an if(false) region is inactive)
#pragma omp parallel for schedule(dynamic, 1) if (nparts >= omp_get_max_threads())
    for (int32_t p = 0; p < nparts; ++p)
      og.run([&] {
        index_maps(g[l].A(), g[l].owner, g[l].oi, p, lev.parts[p]);
        gl[l][p].build(lev.parts[p].ghost_to_global);
      });
    og.rethrow();
  }
  for (int32_t l = 0; l < L; ++l) {
    Level& lev = h.levels[l];
    OmpGuard og;
#pragma omp parallel for schedule(dynamic, 1) if (nparts >= omp_get_max_threads())
    for (int32_t p = 0; p < nparts; ++p) og.run([&] {
      PartLevel& pl = lev.parts[p];
      split_blocks(g[l].A(), pl.own_to_global, g[l].owner, g[l].oi.lid, p, pl.n_own, pl.n_ghost, gl[l][p],
                   pl.blk[PAMG_A_OO], pl.blk[PAMG_A_OG]);
      fill_diag(pl);
      if (l + 1 < L) {
        const PartLevel& pc = h.levels[l + 1].parts[p];
        pl.n_own_coarse = pc.n_own;
        pl.n_ghost_coarse = pc.n_ghost;
        split_blocks(g[l].P, pl.own_to_global, g[l + 1].owner, g[l + 1].oi.lid, p, pc.n_own, pc.n_ghost, gl[l + 1][p],
                     pl.blk[PAMG_P_OO], pl.blk[PAMG_P_OG]);
        split_blocks(g[l].R, pc.own_to_global, g[l].owner, g[l].oi.lid, p, pl.n_own, pl.n_ghost, gl[l][p],
                     pl.blk[PAMG_R_OO], pl.blk[PAMG_R_OG]);
        pl.agg_local.resize(pl.n_own);
        for (int64_t i = 0; i < pl.n_own; ++i) pl.agg_local[i] = g[l].agg_loc[pl.own_to_global[i]];
      }
    });
    og.rethrow();
    build_halo_plans(lev, nparts);
  }
  tm.lap("localisation", L - 1);
  h.n_coarse = g[L - 1].A().nrows;
  dense_inverse(g[L - 1].A(), h.coarse_inv);
  tm.lap("dense inverse", L - 1);
  h.coarse_part_offset.assign(nparts + 1, 0);
  for (int32_t p = 0; p < nparts; ++p)
    h.coarse_part_offset[p + 1] = h.coarse_part_offset[p] + h.levels[L - 1].parts[p].n_own;
  h.ready = true;
}

void finalize_external(Hierarchy& h) {
  const int32_t L = (int32_t)h.levels.size();
  for (int32_t l = 0; l < L; ++l) {
    Level& lev = h.levels[l];
    int64_t ng = 0;
    for (int32_t p = 0; p < h.nparts; ++p) {
      if (!lev.parts[p].present) throw std::runtime_error("external hierarchy: a (level, part) was never uploaded");
      ng += lev.parts[p].n_own;
      fill_diag(lev.parts[p]);
    }
    lev.n_global = ng;
    build_halo_plans(lev, h.nparts);
  }
  if (h.n_coarse != h.levels[L - 1].n_global) throw std::runtime_error("coarse inverse size != coarsest level size");
  h.coarse_part_offset.assign(h.nparts + 1, 0);
  for (int32_t p = 0; p < h.nparts; ++p)
    h.coarse_part_offset[p + 1] = h.coarse_part_offset[p] + h.levels[L - 1].parts[p].n_own;
  h.ready = true;
}

}  // namespace pamg
