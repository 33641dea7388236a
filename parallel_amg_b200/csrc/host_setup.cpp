// host_setup.cpp — problem gallery, index partitions and smoothed-aggregation setup on the
// host (product code; the hierarchy is "built on the host and uploaded once", BASELINE.json
// north_star).  Algorithm definitions: SURVEY.md Appendix B [DEFINED-HERE]; the reference
// snapshot has no code to follow (/root/reference/README.md:1-2).  Checked bit-exact
// (structure, index maps, aggregates) against oracle/amg_oracle.py in tests/test_setup_parity.py.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <utility>

#include "host.hpp"

namespace pamg {

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
static void spgemm(const Csr& A, const Csr& B, Csr& C) {
  const int64_t n = A.nrows;
  C.nrows = n;
  C.ncols = B.ncols;
  C.ptr.assign(n + 1, 0);
  const int nt = omp_get_max_threads();
  std::vector<std::vector<int64_t>> tcol(nt);
  std::vector<std::vector<double>> tval(nt);
  std::vector<int64_t> tbeg(nt + 1, 0);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = n * t / nt, r1 = n * (t + 1) / nt;
    std::vector<std::pair<int64_t, double>> acc;
    auto& oc = tcol[t];
    auto& ov = tval[t];
    for (int64_t i = r0; i < r1; ++i) {
      acc.clear();
      for (int64_t ka = A.ptr[i]; ka < A.ptr[i + 1]; ++ka) {
        const int64_t k = A.col[ka];
        const double a = A.val[ka];
        for (int64_t kb = B.ptr[k]; kb < B.ptr[k + 1]; ++kb) acc.emplace_back(B.col[kb], a * B.val[kb]);
      }
      std::stable_sort(acc.begin(), acc.end(),
                       [](const std::pair<int64_t, double>& x, const std::pair<int64_t, double>& y) { return x.first < y.first; });
      int64_t cnt = 0;
      for (size_t q = 0; q < acc.size();) {
        int64_t cidx = acc[q].first;
        double s = acc[q].second;
        size_t e = q + 1;
        while (e < acc.size() && acc[e].first == cidx) s += acc[e++].second;
        oc.push_back(cidx);
        ov.push_back(s);
        ++cnt;
        q = e;
      }
      C.ptr[i + 1] = cnt;
    }
  }
  for (int64_t i = 0; i < n; ++i) C.ptr[i + 1] += C.ptr[i];
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

static void transpose(const Csr& A, Csr& T) {
  T.nrows = A.ncols;
  T.ncols = A.nrows;
  T.ptr.assign(T.nrows + 1, 0);
  for (int64_t k = 0; k < A.nnz(); ++k) T.ptr[A.col[k] + 1]++;
  for (int64_t i = 0; i < T.nrows; ++i) T.ptr[i + 1] += T.ptr[i];
  T.col.resize(A.nnz());
  T.val.resize(A.nnz());
  std::vector<int64_t> pos(T.ptr.begin(), T.ptr.end() - 1);
  for (int64_t i = 0; i < A.nrows; ++i)
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) {
      int64_t q = pos[A.col[k]]++;
      T.col[q] = i;
      T.val[q] = A.val[k];
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
  oi.lid.resize(n);
  std::vector<int64_t> cnt(nparts, 0);
  for (int64_t g = 0; g < n; ++g) {
    if (owner[g] < 0 || owner[g] >= nparts) throw std::runtime_error("owner id out of range");
    cnt[owner[g]]++;
  }
  for (int32_t p = 0; p < nparts; ++p) oi.own[p].reserve(cnt[p]);
  for (int64_t g = 0; g < n; ++g) {
    auto& v = oi.own[owner[g]];
    if ((int64_t)v.size() >= INT32_MAX) throw std::runtime_error("part too large for int32 local ids");
    oi.lid[g] = (int32_t)v.size();
    v.push_back(g);
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

// ------------------------------------------------------------------------------------------
// smoothed prolongator P = (I - w D_F^-1 A_F) P0, w = 4/(3 rho_F)       (App. B items 3-4)
// ------------------------------------------------------------------------------------------
static void build_prolongator(const Csr& A, const std::vector<int64_t>& agg_gid, int64_t nc, double eps,
                              const std::vector<double>& absdiag, Csr& P, double* omega_out) {
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

  P.nrows = n;
  P.ncols = nc;
  P.ptr.assign(n + 1, 0);
  const int nt = omp_get_max_threads();
  std::vector<std::vector<int64_t>> tcol(nt);
  std::vector<std::vector<double>> tval(nt);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = n * t / nt, r1 = n * (t + 1) / nt;
    std::vector<std::pair<int64_t, double>> acc;
    for (int64_t i = r0; i < r1; ++i) {
      acc.clear();
      for (int64_t k = F->ptr[i]; k < F->ptr[i + 1]; ++k) acc.emplace_back(agg_gid[F->col[k]], F->val[k]);
      std::stable_sort(acc.begin(), acc.end(),
                       [](const std::pair<int64_t, double>& x, const std::pair<int64_t, double>& y) { return x.first < y.first; });
      const double w = -(omega * dinv[i]);
      const int64_t ci = agg_gid[i];
      bool placed = false;
      int64_t cnt = 0;
      for (size_t q = 0; q < acc.size();) {
        const int64_t c = acc[q].first;
        double s = acc[q].second;
        size_t e = q + 1;
        while (e < acc.size() && acc[e].first == c) s += acc[e++].second;
        if (!placed && ci < c) {  // tentative entry without an A_F P0 partner (no stored diagonal)
          tcol[t].push_back(ci);
          tval[t].push_back(1.0);
          ++cnt;
          placed = true;
        }
        double v = w * s;
        if (c == ci) {
          v = 1.0 + v;
          placed = true;
        }
        tcol[t].push_back(c);
        tval[t].push_back(v);
        ++cnt;
        q = e;
      }
      if (!placed) {
        tcol[t].push_back(ci);
        tval[t].push_back(1.0);
        ++cnt;
      }
      P.ptr[i + 1] = cnt;
    }
  }
  for (int64_t i = 0; i < n; ++i) P.ptr[i + 1] += P.ptr[i];
  P.col.resize(P.ptr[n]);
  P.val.resize(P.ptr[n]);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t r0 = n * t / nt;
    std::copy(tcol[t].begin(), tcol[t].end(), P.col.begin() + P.ptr[r0]);
    std::copy(tval[t].begin(), tval[t].end(), P.val.begin() + P.ptr[r0]);
  }
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
  for (int64_t i : pl.own_to_global)
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
      if (owner[A.col[k]] != p) gh.emplace_back(owner[A.col[k]], A.col[k]);
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
  oo.ptr.assign(nr + 1, 0);
  og.ptr.assign(nr + 1, 0);
  for (int64_t r = 0; r < nr; ++r) {
    const int64_t i = rows[r];
    int64_t a = 0, b = 0;
    for (int64_t k = M.ptr[i]; k < M.ptr[i + 1]; ++k) (owner_c[M.col[k]] == p ? a : b)++;
    oo.ptr[r + 1] = oo.ptr[r] + a;
    og.ptr[r + 1] = og.ptr[r] + b;
  }
  if (oo.ptr[nr] > INT32_MAX || og.ptr[nr] > INT32_MAX) throw std::runtime_error("block nnz exceeds int32");
  oo.col.resize(oo.ptr[nr]);
  oo.val.resize(oo.ptr[nr]);
  og.col.resize(og.ptr[nr]);
  og.val.resize(og.ptr[nr]);
  std::vector<std::pair<int32_t, double>> tmp;
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
        const int32_t s = gl.find(j);
        if (s < 0) throw std::runtime_error("column outside the part's own+ghost set");
        tmp.emplace_back(s, M.val[k]);
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
  for (int64_t i = 0; i < pl.n_own; ++i) {
    for (int64_t k = oo.ptr[i]; k < oo.ptr[i + 1]; ++k)
      if (oo.col[k] == i) pl.diag[i] = oo.val[k];
    double s = 0.0;
    if (!og.ptr.empty())
      for (int64_t k = og.ptr[i]; k < og.ptr[i + 1]; ++k) s += std::fabs(og.val[k]);
    pl.diag_l1[i] = pl.diag[i] + s;
    if (pl.diag[i] == 0.0) throw std::runtime_error("zero diagonal");
  }
}

// ------------------------------------------------------------------------------------------
// whole setup
// ------------------------------------------------------------------------------------------
void build_hierarchy(const Csr& A0, const std::vector<int32_t>& owner0, int32_t nparts, const pamg_options& o, Hierarchy& h) {
  if (A0.nrows != (int64_t)owner0.size()) throw std::runtime_error("owner size mismatch");
  h = Hierarchy();
  h.nparts = nparts;
  h.opts = o;

  struct G {
    Csr A, P, R;
    std::vector<int32_t> owner;
    OwnIndex oi;
    std::vector<int32_t> agg_loc;  // by gid
    double rho = 0, omega_p = 0;
  };
  std::vector<G> g(1);
  g[0].A = A0;
  g[0].owner = owner0;
  while (true) {
    const double eps = o.eps_strength * std::pow(0.5, (double)(g.size() - 1));  // Vanek: eps_l = eps 2^-l
    G& cur = g.back();
    const int64_t n = cur.A.nrows;
    build_own_index(cur.owner, nparts, cur.oi);
    std::vector<double> d, dinv(n), absdiag(n);
    diagonal(cur.A, d);
    for (int64_t i = 0; i < n; ++i) {
      if (d[i] == 0.0) throw std::runtime_error("zero diagonal");
      dinv[i] = 1.0 / d[i];
      absdiag[i] = std::fabs(d[i]);
    }
    cur.rho = gershgorin_rho(cur.A, dinv);
    if (n <= o.coarse_size || (int32_t)g.size() >= o.max_levels) break;

    std::vector<std::vector<int32_t>> aggs(nparts);
    std::vector<int64_t> counts(nparts, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t p = 0; p < nparts; ++p) counts[p] = aggregate_part(cur.A, cur.owner, cur.oi, p, eps, absdiag, aggs[p]);
    std::vector<int64_t> off(nparts + 1, 0);
    for (int32_t p = 0; p < nparts; ++p) off[p + 1] = off[p] + counts[p];
    const int64_t nc = off[nparts];
    if (nc >= n) break;
    std::vector<int64_t> agg_gid(n);
    cur.agg_loc.resize(n);
    for (int32_t p = 0; p < nparts; ++p) {
      const auto& own = cur.oi.own[p];
      for (size_t li = 0; li < own.size(); ++li) {
        cur.agg_loc[own[li]] = aggs[p][li];
        agg_gid[own[li]] = off[p] + aggs[p][li];
      }
    }
    build_prolongator(cur.A, agg_gid, nc, eps, absdiag, cur.P, &cur.omega_p);
    transpose(cur.P, cur.R);
    G nxt;
    {
      Csr AP;
      spgemm(cur.A, cur.P, AP);
      spgemm(cur.R, AP, nxt.A);
    }
    nxt.owner.resize(nc);
    for (int32_t p = 0; p < nparts; ++p)
      for (int64_t c = off[p]; c < off[p + 1]; ++c) nxt.owner[c] = p;
    g.push_back(std::move(nxt));
  }

  const int32_t L = (int32_t)g.size();
  h.levels.resize(L);
  std::vector<std::vector<GhostLookup>> gl(L, std::vector<GhostLookup>(nparts));
  for (int32_t l = 0; l < L; ++l) {
    Level& lev = h.levels[l];
    lev.n_global = g[l].A.nrows;
    lev.rho = g[l].rho;
    lev.omega_p = g[l].omega_p;
    lev.parts.resize(nparts);
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t p = 0; p < nparts; ++p) {
      index_maps(g[l].A, g[l].owner, g[l].oi, p, lev.parts[p]);
      gl[l][p].build(lev.parts[p].ghost_to_global);
    }
  }
  for (int32_t l = 0; l < L; ++l) {
    Level& lev = h.levels[l];
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t p = 0; p < nparts; ++p) {
      PartLevel& pl = lev.parts[p];
      split_blocks(g[l].A, pl.own_to_global, g[l].owner, g[l].oi.lid, p, pl.n_own, pl.n_ghost, gl[l][p],
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
    }
    build_halo_plans(lev, nparts);
  }
  h.n_coarse = g[L - 1].A.nrows;
  dense_inverse(g[L - 1].A, h.coarse_inv);
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
