/* pamg_oracle.c — CPU restatement (plain C + OpenMP) of the AMG-PCG SOLVE phase on emulated
 * parts.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * *** PARITY UNPINNED *** — the reference snapshot (/root/reference/README.md:1-2, LICENSE) has
 * no code, tests or golden vectors and Julia is not installed; this file restates
 * oracle/amg_oracle.py (the normative spec of this repo, see its header) function by function
 * so that the same algorithm can be timed on the host cores at full size:
 *   orc_consistent   <- amg_oracle.consistent      (PartitionedArrays consistent!, App. A)
 *   orc_mul          <- amg_oracle._mul / spmv     (mul!: own-own then own-ghost columns)
 *   orc_jacobi       <- amg_oracle.jacobi_sweep    (damped / l1 Jacobi; w passed in)
 *   orc_vcycle       <- amg_oracle.vcycle
 *   orc_pcg          <- amg_oracle.pcg
 *   orc_fgmres       <- amg_oracle.fgmres          (restarted flexible GMRES, right-preconditioned)
 * It is checked against amg_oracle.py in tests/test_c_oracle.py (<= 1e-12, same iteration counts).
 *
 * Execution model: PartitionedArrays debug backend — all parts in one process, each part reads
 * only its own + ghost entries; OpenMP parallelises over the rows of a part.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t nrows;
  const int64_t* ptr;
  const int32_t* col;
  const double* val;
} blk_t;

typedef struct {
  int64_t n_own, n_ghost, n_own_c;
  blk_t b[6];                 /* A_oo A_og P_oo P_og R_oo R_og (borrowed pointers) */
  const double* w;            /* smoother weight per own row (borrowed) */
  const int32_t* gsrc_part;   /* ghost slot -> owner part */
  const int32_t* gsrc_lid;    /* ghost slot -> own-local id in the owner */
  double *x, *x2, *b_, *t;    /* local vectors: own then ghost */
} part_t;

typedef struct {
  int32_t nparts, nlevels, nu_pre, nu_post;
  int32_t cycle_w;            /* 1: W-cycle (amg_oracle.vcycle, opts["cycle"] == "w") */
  part_t* p;                  /* [nlevels][nparts] */
  int64_t n_coarse;
  const double* inv;          /* row-major, indexed by coarse gid */
  const int64_t** coarse_gid; /* [nparts] own gids on the coarsest level */
  const int64_t** coarse_ggid;/* [nparts] ghost gids on the coarsest level */
  double* cb;                 /* gathered coarse rhs */
} orc_t;

#define PART(o, l, q) ((o)->p[(size_t)(l) * (o)->nparts + (q)])

orc_t* orc_create(int32_t nparts, int32_t nlevels, int32_t nu_pre, int32_t nu_post) {
  orc_t* o = (orc_t*)calloc(1, sizeof(orc_t));
  o->nparts = nparts;
  o->nlevels = nlevels;
  o->nu_pre = nu_pre;
  o->nu_post = nu_post;
  o->p = (part_t*)calloc((size_t)nparts * nlevels, sizeof(part_t));
  o->coarse_gid = (const int64_t**)calloc(nparts, sizeof(void*));
  o->coarse_ggid = (const int64_t**)calloc(nparts, sizeof(void*));
  return o;
}

void orc_set_part(orc_t* o, int32_t level, int32_t part, int64_t n_own, int64_t n_ghost, int64_t n_own_c,
                  const int64_t* const* ptr, const int32_t* const* col, const double* const* val, const double* w,
                  const int32_t* gsrc_part, const int32_t* gsrc_lid) {
  part_t* p = &PART(o, level, part);
  p->n_own = n_own;
  p->n_ghost = n_ghost;
  p->n_own_c = n_own_c;
  for (int k = 0; k < 6; ++k) {
    p->b[k].nrows = (k >= 4) ? n_own_c : n_own;
    p->b[k].ptr = ptr[k];
    p->b[k].col = col[k];
    p->b[k].val = val[k];
  }
  p->w = w;
  p->gsrc_part = gsrc_part;
  p->gsrc_lid = gsrc_lid;
  size_t nl = (size_t)(n_own + n_ghost) + 1;
  p->x = (double*)calloc(nl, sizeof(double));
  p->x2 = (double*)calloc(nl, sizeof(double));
  p->b_ = (double*)calloc(nl, sizeof(double));
  p->t = (double*)calloc(nl, sizeof(double));
}

void orc_set_coarse(orc_t* o, int64_t n, const double* inv, int32_t part, const int64_t* own_gid, const int64_t* ghost_gid) {
  o->n_coarse = n;
  o->inv = inv;
  o->coarse_gid[part] = own_gid;
  o->coarse_ggid[part] = ghost_gid;
  if (!o->cb) o->cb = (double*)calloc((size_t)n + 1, sizeof(double));
}

void orc_set_cycle(orc_t* o, int32_t w_cycle) { o->cycle_w = w_cycle; }

void orc_destroy(orc_t* o) {
  if (!o) return;
  for (size_t i = 0; i < (size_t)o->nparts * o->nlevels; ++i) {
    free(o->p[i].x);
    free(o->p[i].x2);
    free(o->p[i].b_);
    free(o->p[i].t);
  }
  free(o->p);
  free((void*)o->coarse_gid);
  free((void*)o->coarse_ggid);
  free(o->cb);
  free(o);
}

/* which local vector of a part: 0 x, 1 x2, 2 b, 3 t */
static double* vec_of(part_t* p, int which) { return which == 0 ? p->x : which == 1 ? p->x2 : which == 2 ? p->b_ : p->t; }

/* consistent!(v): every ghost entry <- its owner's value.  vs[q] = local vector of part q. */
static void orc_consistent(orc_t* o, int level, double** vs) {
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, level, q);
    double* v = vs[q];
    const int64_t n = p->n_own;
#pragma omp parallel for schedule(static) if (p->n_ghost > 4096)
    for (int64_t s = 0; s < p->n_ghost; ++s) v[n + s] = vs[p->gsrc_part[s]][p->gsrc_lid[s]];
  }
}

/* (M_oo x_own + M_og x_ghost)[i]: own columns (ascending) then ghost columns (ascending) */
static inline double row_mul(const blk_t* oo, const blk_t* og, const double* x, int64_t n_own_cols, int64_t i) {
  double s = 0.0;
  for (int64_t k = oo->ptr[i]; k < oo->ptr[i + 1]; ++k) s += oo->val[k] * x[oo->col[k]];
  if (og->ptr) {
    double g = 0.0;
    const int64_t b = og->ptr[i], e = og->ptr[i + 1];
    if (e > b) {
      for (int64_t k = b; k < e; ++k) g += og->val[k] * x[n_own_cols + og->col[k]];
      s = s + g; /* numpy: y = A_oo@x_own ; y = y + A_og@x_ghost */
    }
  }
  return s;
}

/* y = A x on one level, all parts (halo inside) */
void orc_spmv(orc_t* o, int level, double** xs, double** ys) {
  orc_consistent(o, level, xs);
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, level, q);
    const double* x = xs[q];
    double* y = ys[q];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < p->n_own; ++i) y[i] = row_mul(&p->b[0], &p->b[1], x, p->n_own, i);
  }
}

/* xn = x + w .* (b - A x) */
static void orc_jacobi(orc_t* o, int level, double** xs, double** bs, double** xn) {
  orc_consistent(o, level, xs);
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, level, q);
    const double *x = xs[q], *b = bs[q], *w = p->w;
    double* out = xn[q];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < p->n_own; ++i) out[i] = x[i] + w[i] * (b[i] - row_mul(&p->b[0], &p->b[1], x, p->n_own, i));
  }
}

static void coarse_solve(orc_t* o) {
  const int L = o->nlevels - 1;
  const int64_t n = o->n_coarse;
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, L, q);
    for (int64_t i = 0; i < p->n_own; ++i) o->cb[o->coarse_gid[q][i]] = p->b_[i];
  }
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, L, q);
    for (int64_t r = 0; r < p->n_own + p->n_ghost; ++r) {
      const int64_t g = r < p->n_own ? o->coarse_gid[q][r] : o->coarse_ggid[q][r - p->n_own];
      const double* row = o->inv + (size_t)g * n;
      double s = 0.0;
      for (int64_t j = 0; j < n; ++j) s += row[j] * o->cb[j];
      p->x[r] = s;
    }
  }
}

/* cycle from x = 0 (zero_guess) or from the current part.x; rhs in part.b_, result in part.x (local, ghosts of the
 * result are NOT consistent).  W-cycle: the coarse problem is visited twice, the second time from the first result. */
static void vcycle_from(orc_t* o, int l, int zero_guess);
static void vcycle(orc_t* o, int l) { vcycle_from(o, l, 1); }
static void vcycle_from(orc_t* o, int l, int zero_guess) {
  const int P = o->nparts;
  if (l == o->nlevels - 1) {
    coarse_solve(o);
    return;
  }
  double *cur[256], *nxt[256], *bs[256], *ts[256];
  for (int q = 0; q < P; ++q) {
    part_t* p = &PART(o, l, q);
    cur[q] = p->x;
    nxt[q] = p->x2;
    bs[q] = p->b_;
    ts[q] = p->t;
    if (zero_guess) memset(p->x, 0, sizeof(double) * (size_t)(p->n_own + p->n_ghost));
  }
  for (int s = 0; s < o->nu_pre; ++s) {
    orc_jacobi(o, l, cur, bs, nxt);
    for (int q = 0; q < P; ++q) { double* tmp = cur[q]; cur[q] = nxt[q]; nxt[q] = tmp; }
  }
  /* r = b - A x */
  orc_consistent(o, l, cur);
  for (int q = 0; q < P; ++q) {
    part_t* p = &PART(o, l, q);
    const double *x = cur[q], *b = bs[q];
    double* t = ts[q];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < p->n_own; ++i) t[i] = b[i] - row_mul(&p->b[0], &p->b[1], x, p->n_own, i);
  }
  /* b_c = R r */
  orc_consistent(o, l, ts);
  for (int q = 0; q < P; ++q) {
    part_t* p = &PART(o, l, q);
    part_t* pc = &PART(o, l + 1, q);
    const double* t = ts[q];
    double* bc = pc->b_;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < pc->n_own; ++i) bc[i] = row_mul(&p->b[4], &p->b[5], t, p->n_own, i);
  }
  vcycle_from(o, l + 1, 1);
  if (o->cycle_w && l + 1 < o->nlevels - 1) vcycle_from(o, l + 1, 0);
  /* x += P e_c (coarse result is always in part.x of level l+1) */
  {
    double* ec[256];
    for (int q = 0; q < P; ++q) ec[q] = PART(o, l + 1, q).x;
    if (l + 1 != o->nlevels - 1) orc_consistent(o, l + 1, ec); /* the coarsest solve fills its own ghosts */
    for (int q = 0; q < P; ++q) {
      part_t* p = &PART(o, l, q);
      part_t* pc = &PART(o, l + 1, q);
      double* x = cur[q];
      const double* e = ec[q];
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < p->n_own; ++i) x[i] = x[i] + row_mul(&p->b[2], &p->b[3], e, pc->n_own, i);
    }
  }
  for (int s = 0; s < o->nu_post; ++s) {
    orc_jacobi(o, l, cur, bs, nxt);
    for (int q = 0; q < P; ++q) { double* tmp = cur[q]; cur[q] = nxt[q]; nxt[q] = tmp; }
  }
  for (int q = 0; q < P; ++q) {
    part_t* p = &PART(o, l, q);
    if (cur[q] != p->x) memcpy(p->x, cur[q], sizeof(double) * (size_t)p->n_own);
  }
}

/* z = V(b): b_parts/z_parts are own-length arrays per part */
void orc_vcycle(orc_t* o, const double* const* b_parts, double* const* z_parts) {
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, 0, q);
    memcpy(p->b_, b_parts[q], sizeof(double) * (size_t)p->n_own);
  }
  vcycle(o, 0);
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, 0, q);
    memcpy(z_parts[q], p->x, sizeof(double) * (size_t)p->n_own);
  }
}

static double pdot(orc_t* o, double** us, double** vs) {
  double tot = 0.0; /* per-part partials summed in ascending part order */
  for (int q = 0; q < o->nparts; ++q) {
    part_t* p = &PART(o, 0, q);
    const double *u = us[q], *v = vs[q];
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < p->n_own; ++i) s += u[i] * v[i];
    tot += s;
  }
  return tot;
}

/* PCG from x0 = 0; returns iterations; hist[0..iters]; x_parts own-length outputs. */
/* precond: 0 plain CG, 1 PCG, 2 flexible PCG (beta = z_{k+1}.(r_{k+1} - r_k) / (z_k.r_k), amg_oracle.pcg flexible=True) */
int32_t orc_pcg(orc_t* o, const double* const* b_parts, double* const* x_parts, double rtol, int32_t maxiter,
                int32_t precond, double* hist) {
  const int P = o->nparts;
  const int flexible = precond == 2;
  double *xs[256], *rs[256], *ps[256], *qs[256], *zs[256], *ro[256];
  for (int q = 0; q < P; ++q) {
    part_t* p = &PART(o, 0, q);
    size_t nl = (size_t)(p->n_own + p->n_ghost) + 1;
    xs[q] = (double*)calloc(nl, sizeof(double));
    rs[q] = (double*)calloc(nl, sizeof(double));
    ps[q] = (double*)calloc(nl, sizeof(double));
    qs[q] = (double*)calloc(nl, sizeof(double));
    zs[q] = (double*)calloc(nl, sizeof(double));
    ro[q] = (double*)calloc(nl, sizeof(double));
    memcpy(rs[q], b_parts[q], sizeof(double) * (size_t)p->n_own);
  }
#define APPLY_M()                                                                             \
  do {                                                                                        \
    if (precond) {                                                                            \
      for (int q = 0; q < P; ++q) {                                                           \
        part_t* p = &PART(o, 0, q);                                                           \
        memcpy(p->b_, rs[q], sizeof(double) * (size_t)p->n_own);                              \
      }                                                                                       \
      vcycle(o, 0);                                                                           \
      for (int q = 0; q < P; ++q) {                                                           \
        part_t* p = &PART(o, 0, q);                                                           \
        memcpy(zs[q], p->x, sizeof(double) * (size_t)p->n_own);                               \
      }                                                                                       \
    } else {                                                                                  \
      for (int q = 0; q < P; ++q) memcpy(zs[q], rs[q], sizeof(double) * (size_t)PART(o, 0, q).n_own); \
    }                                                                                         \
  } while (0)
  APPLY_M();
  for (int q = 0; q < P; ++q) memcpy(ps[q], zs[q], sizeof(double) * (size_t)PART(o, 0, q).n_own);
  double rho = pdot(o, rs, zs);
  const double r0 = sqrt(pdot(o, rs, rs));
  if (hist) hist[0] = r0;
  int32_t it = 0;
  if (r0 != 0.0) {
    while (it < maxiter) {
      orc_spmv(o, 0, ps, qs);
      const double alpha = rho / pdot(o, ps, qs);
      if (flexible)
        for (int q = 0; q < P; ++q) memcpy(ro[q], rs[q], sizeof(double) * (size_t)PART(o, 0, q).n_own);
      for (int q = 0; q < P; ++q) {
        part_t* p = &PART(o, 0, q);
        double *x = xs[q], *r = rs[q];
        const double *pp = ps[q], *qq = qs[q];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < p->n_own; ++i) {
          x[i] += alpha * pp[i];
          r[i] -= alpha * qq[i];
        }
      }
      ++it;
      const double rn = sqrt(pdot(o, rs, rs));
      if (hist) hist[it] = rn;
      if (rn <= rtol * r0) break;
      APPLY_M();
      const double rho_new = pdot(o, rs, zs);
      const double beta = flexible ? (rho_new - pdot(o, ro, zs)) / rho : rho_new / rho;
      rho = rho_new;
      for (int q = 0; q < P; ++q) {
        part_t* p = &PART(o, 0, q);
        double* pp = ps[q];
        const double* z = zs[q];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < p->n_own; ++i) pp[i] = z[i] + beta * pp[i];
      }
    }
  }
  for (int q = 0; q < P; ++q) {
    memcpy(x_parts[q], xs[q], sizeof(double) * (size_t)PART(o, 0, q).n_own);
    free(xs[q]);
    free(rs[q]);
    free(ps[q]);
    free(qs[q]);
    free(zs[q]);
    free(ro[q]);
  }
  return it;
}

/* Restarted flexible GMRES, FGMRES(m), right-preconditioned by one cycle, x0 = 0 (amg_oracle.fgmres): Arnoldi with
 * modified Gram-Schmidt on w = A M^-1 v_j, Givens rotations, hist[k] = |g_{k}| (the residual estimate of inner step k,
 * hist[0] = ||b||); stops at estimate <= rtol ||b|| or maxiter inner steps.  Returns the inner iterations. */
int32_t orc_fgmres(orc_t* o, const double* const* b_parts, double* const* x_parts, double rtol, int32_t maxiter,
                   int32_t restart, int32_t precond, double* hist) {
  const int P = o->nparts, m = restart;
  double** V = (double**)calloc((size_t)(m + 1) * P, sizeof(double*)); /* V[j*P + q] */
  double** Z = (double**)calloc((size_t)m * P, sizeof(double*));
  double *xs[256], *rs[256], *ws[256];
  for (int q = 0; q < P; ++q) {
    part_t* p = &PART(o, 0, q);
    size_t nl = (size_t)(p->n_own + p->n_ghost) + 1;
    xs[q] = (double*)calloc(nl, sizeof(double));
    rs[q] = (double*)calloc(nl, sizeof(double));
    ws[q] = (double*)calloc(nl, sizeof(double));
    memcpy(rs[q], b_parts[q], sizeof(double) * (size_t)p->n_own);
    for (int j = 0; j <= m; ++j) V[(size_t)j * P + q] = (double*)calloc(nl, sizeof(double));
    for (int j = 0; j < m; ++j) Z[(size_t)j * P + q] = (double*)calloc(nl, sizeof(double));
  }
  double* H = (double*)calloc((size_t)(m + 1) * m, sizeof(double)); /* H[i*m + j] */
  double* cs = (double*)calloc(m, sizeof(double));
  double* sn = (double*)calloc(m, sizeof(double));
  double* g = (double*)calloc(m + 1, sizeof(double));
  double* y = (double*)calloc(m, sizeof(double));
  const double beta0 = sqrt(pdot(o, rs, rs));
  if (hist) hist[0] = beta0;
  int32_t it = 0;
  int done = (beta0 == 0.0 || maxiter == 0);
  while (!done) {
    const double beta = sqrt(pdot(o, rs, rs));
    for (int q = 0; q < P; ++q) {
      const int64_t n = PART(o, 0, q).n_own;
      double* v = V[q];
      const double* r = rs[q];
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) v[i] = r[i] / beta;
    }
    memset(H, 0, sizeof(double) * (size_t)(m + 1) * m);
    memset(g, 0, sizeof(double) * (size_t)(m + 1));
    g[0] = beta;
    int j = 0;
    while (j < m) {
      double** vj = V + (size_t)j * P;
      double** zj = Z + (size_t)j * P;
      if (precond) {
        for (int q = 0; q < P; ++q) memcpy(PART(o, 0, q).b_, vj[q], sizeof(double) * (size_t)PART(o, 0, q).n_own);
        vcycle(o, 0);
        for (int q = 0; q < P; ++q) memcpy(zj[q], PART(o, 0, q).x, sizeof(double) * (size_t)PART(o, 0, q).n_own);
      } else {
        for (int q = 0; q < P; ++q) memcpy(zj[q], vj[q], sizeof(double) * (size_t)PART(o, 0, q).n_own);
      }
      orc_spmv(o, 0, zj, ws);
      for (int i = 0; i <= j; ++i) {
        double** vi = V + (size_t)i * P;
        const double h = pdot(o, ws, vi);
        H[(size_t)i * m + j] = h;
        for (int q = 0; q < P; ++q) {
          const int64_t n = PART(o, 0, q).n_own;
          double* w = ws[q];
          const double* v = vi[q];
#pragma omp parallel for schedule(static)
          for (int64_t k = 0; k < n; ++k) w[k] = w[k] - h * v[k];
        }
      }
      const double hn = sqrt(pdot(o, ws, ws));
      H[(size_t)(j + 1) * m + j] = hn;
      for (int q = 0; q < P; ++q) {
        const int64_t n = PART(o, 0, q).n_own;
        double* v = V[(size_t)(j + 1) * P + q];
        const double* w = ws[q];
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < n; ++k) v[k] = hn != 0.0 ? w[k] / hn : w[k];
      }
      for (int i = 0; i < j; ++i) { /* previous rotations on the new column */
        const double t = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
        H[(size_t)i * m + j] = t;
      }
      const double d = hypot(H[(size_t)j * m + j], H[(size_t)(j + 1) * m + j]);
      cs[j] = H[(size_t)j * m + j] / d;
      sn[j] = H[(size_t)(j + 1) * m + j] / d;
      H[(size_t)j * m + j] = d;
      H[(size_t)(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      ++j;
      ++it;
      if (hist) hist[it] = fabs(g[j]);
      if (fabs(g[j]) <= rtol * beta0 || it >= maxiter) {
        done = 1;
        break;
      }
    }
    for (int i = j - 1; i >= 0; --i) { /* back substitution, the dot in ascending column order */
      double s = 0.0;
      for (int k = i + 1; k < j; ++k) s += H[(size_t)i * m + k] * y[k];
      y[i] = (g[i] - s) / H[(size_t)i * m + i];
    }
    for (int i = 0; i < j; ++i)
      for (int q = 0; q < P; ++q) {
        const int64_t n = PART(o, 0, q).n_own;
        double* x = xs[q];
        const double* z = Z[(size_t)i * P + q];
        const double yi = y[i];
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < n; ++k) x[k] = x[k] + yi * z[k];
      }
    if (!done) {
      orc_spmv(o, 0, xs, ws);
      for (int q = 0; q < P; ++q) {
        const int64_t n = PART(o, 0, q).n_own;
        double* r = rs[q];
        const double *b = b_parts[q], *a = ws[q];
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < n; ++k) r[k] = b[k] - a[k];
      }
    }
  }
  for (int q = 0; q < P; ++q) {
    memcpy(x_parts[q], xs[q], sizeof(double) * (size_t)PART(o, 0, q).n_own);
    free(xs[q]);
    free(rs[q]);
    free(ws[q]);
    for (int j = 0; j <= m; ++j) free(V[(size_t)j * P + q]);
    for (int j = 0; j < m; ++j) free(Z[(size_t)j * P + q]);
  }
  free(V); free(Z); free(H); free(cs); free(sn); free(g); free(y);
  return it;
}
