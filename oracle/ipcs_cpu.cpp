// CPU ORACLE / CPU BASELINE (test + measurement infrastructure -- never linked into the product).
//
// C++17/OpenMP restatement of the reference's IPCS fractional step
// (/root/reference/src/oasisx/fracstep.py:277-705; algebra in SURVEY.md Appendix A) for problem
// sizes the numpy/SuperLU oracle (oracle/ipcs_oracle.py) cannot reach.  It follows the reference's
// own structure: per-component solves with one shared matrix (fracstep.py:516-524,613-656), the
// "matrix-vector" RHS strategy A.scale/axpy/axpy/mult (fracstep.py:438-469), CSR (int32/FP64)
// storage, BiCGStab+Jacobi for the tentative velocity, CG+Jacobi (or CG + the geometric multigrid of the GPU arm,
// ipcs_cpu_mg_*) for the pressure (null space projected, fracstep.py:573-591) and CG+Jacobi for the mass solves -- the
// Krylov choices of SURVEY.md 8(d),
// because the reference's DOLFINx/PETSc/MUMPS stack cannot be installed here.
//
// PARITY UNPINNED for the same reason as oracle/ipcs_oracle.py; it is pinned against that numpy
// oracle in tests/test_cpu_port.py.  Only tests/ and bench.py's cpu_baseline / --impl reference legs
// load this library.  Build: oracle/build_cpu.py  (g++ -O3 -march=native -fopenmp).
#include <omp.h>
#if defined(__linux__)
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define B2_TABLE_QUAL static const
#include "../oasisx_b200/csrc/ref_tables.h"

namespace {

struct Tables {
  int d, nv, nq;
  const double *MV, *SV, *T, *PX, *GX, *SQ, *MQ, *LV, *LQ;
};

#define TBL(DD, PP)                                                                                          \
  Tables { DD, (int)(sizeof(REF_D##DD##P##PP##_LV) / sizeof(double)), DD + 1, &REF_D##DD##P##PP##_MV[0][0],          \
           &REF_D##DD##P##PP##_SV[0][0][0][0], &REF_D##DD##P##PP##_T[0][0][0][0], &REF_D##DD##P##PP##_PX[0][0][0], \
           &REF_D##DD##P##PP##_GX[0][0][0], &REF_D##DD##P##PP##_SQ[0][0][0][0], &REF_D##DD##P##PP##_MQ[0][0],      \
           &REF_D##DD##P##PP##_LV[0], &REF_D##DD##P##PP##_LQ[0] }

Tables tables_for(int d, int deg) {
  if (d == 2 && deg == 1) return TBL(2, 1);
  if (d == 2 && deg == 2) return TBL(2, 2);
  if (d == 3 && deg == 1) return TBL(3, 1);
  return TBL(3, 2);
}

struct Csr {
  int n_rows = 0, n_cols = 0;
  std::vector<int> ptr, col;
  int find(int r, int c) const {
    const int* b = col.data() + ptr[r];
    const int* e = col.data() + ptr[r + 1];
    return (int)(std::lower_bound(b, e, c) - col.data());
  }
};

struct Geo {
  double Kinv[3][3];
  double detJ;
};

struct Ctx {
  int d, degv;
  Tables t;
  int64_t n_nodes, n_cells, nV, nQ;
  std::vector<double> x;
  std::vector<int> cn, vd, qd;
  // dof -> (cell, local index) adjacency
  std::vector<int> vadj_ptr, vadj_cell, vadj_loc, qadj_ptr, qadj_cell, qadj_loc;
  Csr vv, vq, qv, qq;
  std::vector<double> M, K, A, Ap, P[3], G[3], D[3];
  std::vector<double> u[3], u1[3], u2[3], uab[3], rhs1[3], bfirst[3], b0[3], wrk, b3;
  std::vector<double> ps, p, dp, b2, mQ;
  std::vector<int> bc_dofs[3];
  std::vector<double> bc_vals[3];
  std::vector<char> is_bc;
  std::vector<double> dinvA, dinvM, dinvAp;
  double vol = 0, rtol = 1e-10;
  int maxit = 10000, nonzero = 0, block_rtol = 0, extrapolate = 0, steps_done = 0;
  std::vector<double> delta_prev[3], dp_old;
  int dp_hist = 0;
  double bref2 = 0;  // block_rtol: max_k |b_k|^2 of the current vector solve
  int its_t = 0, its_p = 0, its_u = 0;
  // pressure multigrid (same algorithm as the GPU arm's pc_type=mg: V(1,1), damped Jacobi, exact dense coarse solve);
  // the hierarchy (Galerkin operators, transfers, dense inverse) is built by the Python side (oracle/ipcs_cpu.py)
  struct MgLev {
    Csr A, P, R;  // P: rows = dofs of the finer level, cols = this level; R = P^T
    std::vector<double> Av, Pv, Rv, dinv;
  };
  std::vector<MgLev> mg;
  std::vector<double> mg_dense;  // inverse of (A_last + alpha e e^T), row-major
  double mg_omega = 0.85;
  int pressure_mg = 0;
  // experiments (environment IPCS_GUESS_T / _M / _P, IPCS_DEBUG): order of the time extrapolation of the guesses
  int debug = 0, guess_t = 0, guess_m = 0, guess_p = 1;
  mutable double dbg_rr0 = 0;
  std::vector<double> ustar_hist[3][3], delta_hist[3][3], dp_hist3[3];
  int n_hist = 0, n_dp_hist = 0;
};

Geo geometry(const Ctx& c, int64_t cell) {
  Geo g;
  const int d = c.d;
  double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const int* nodes = &c.cn[cell * (d + 1)];
  for (int k = 0; k < d; ++k)
    for (int dl = 0; dl < d; ++dl) J[k][dl] = c.x[3 * (size_t)nodes[dl + 1] + k] - c.x[3 * (size_t)nodes[0] + k];
  if (d == 2) {
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0], id = 1.0 / det;
    g.Kinv[0][0] = J[1][1] * id; g.Kinv[0][1] = -J[0][1] * id;
    g.Kinv[1][0] = -J[1][0] * id; g.Kinv[1][1] = J[0][0] * id;
    g.detJ = std::fabs(det);
  } else {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
           c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02, id = 1.0 / det;
    g.Kinv[0][0] = c00 * id; g.Kinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    g.Kinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    g.Kinv[1][0] = c01 * id; g.Kinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    g.Kinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    g.Kinv[2][0] = c02 * id; g.Kinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    g.Kinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    g.detJ = std::fabs(det);
  }
  return g;
}

void build_adj(int64_t n_dofs, int64_t n_cells, int nd, const std::vector<int>& cd, std::vector<int>& ptr,
               std::vector<int>& cell, std::vector<int>& loc) {
  ptr.assign(n_dofs + 1, 0);
  for (int64_t i = 0; i < n_cells * nd; ++i) ptr[cd[i] + 1]++;
  for (int64_t i = 0; i < n_dofs; ++i) ptr[i + 1] += ptr[i];
  cell.resize(ptr[n_dofs]);
  loc.resize(ptr[n_dofs]);
  std::vector<int> fill(ptr.begin(), ptr.end() - 1);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int i = 0; i < nd; ++i) {
      int p = fill[cd[c * nd + i]]++;
      cell[p] = (int)c;
      loc[p] = i;
    }
}

// create_matrix: row r couples to every column dof of every cell containing r; sorted columns
void build_pattern(int64_t n_rows, int64_t n_cols, const std::vector<int>& adj_ptr, const std::vector<int>& adj_cell,
                   int ncd, const std::vector<int>& coldofs, Csr& out) {
  out.n_rows = (int)n_rows;
  out.n_cols = (int)n_cols;
  out.ptr.assign(n_rows + 1, 0);
  std::vector<int> cnt(n_rows);
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t r = 0; r < n_rows; ++r) {
      tmp.clear();
      for (int a = adj_ptr[r]; a < adj_ptr[r + 1]; ++a)
        for (int j = 0; j < ncd; ++j) tmp.push_back(coldofs[(size_t)adj_cell[a] * ncd + j]);
      std::sort(tmp.begin(), tmp.end());
      cnt[r] = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  }
  for (int64_t r = 0; r < n_rows; ++r) out.ptr[r + 1] = out.ptr[r] + cnt[r];
  out.col.resize(out.ptr[n_rows]);
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t r = 0; r < n_rows; ++r) {
      tmp.clear();
      for (int a = adj_ptr[r]; a < adj_ptr[r + 1]; ++a)
        for (int j = 0; j < ncd; ++j) tmp.push_back(coldofs[(size_t)adj_cell[a] * ncd + j]);
      std::sort(tmp.begin(), tmp.end());
      auto e = std::unique(tmp.begin(), tmp.end());
      std::copy(tmp.begin(), e, out.col.begin() + out.ptr[r]);
    }
  }
}

void spmv(const Csr& m, const std::vector<double>& v, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < m.n_rows; ++r) {
    double s = 0;
    for (int p = m.ptr[r]; p < m.ptr[r + 1]; ++p) s += v[p] * x[m.col[p]];
    y[r] = s;
  }
}

double dot(int64_t n, const double* a, const double* b) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// PCG with Jacobi; returns iterations (negative: not converged)
int cg(const Ctx& c, const Csr& m, const std::vector<double>& vals, const std::vector<double>& dinv, const double* b,
       double* x) {
  const int64_t n = m.n_rows;
  std::vector<double> r(n), z(n), p(n), q(n);
  if (c.nonzero) {
    spmv(m, vals, x, q.data());
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - q[i];
  } else {
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) { r[i] = b[i]; x[i] = 0; }
  }
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) { z[i] = dinv[i] * r[i]; p[i] = z[i]; }
  double rz = dot(n, r.data(), z.data()), bb = dot(n, b, b), rr = dot(n, r.data(), r.data());
  const double tol2 = std::max(c.rtol * c.rtol * (c.block_rtol && c.bref2 > bb ? c.bref2 : bb), 1e-100);
  c.dbg_rr0 = rr;
  int it = 0;
  while (rr > tol2 && it < c.maxit) {
    spmv(m, vals, p.data(), q.data());
    const double alpha = rz / dot(n, p.data(), q.data());
    double rz2 = 0, rr2 = 0;
#pragma omp parallel for reduction(+ : rz2, rr2)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      z[i] = dinv[i] * r[i];
      rz2 += r[i] * z[i];
      rr2 += r[i] * r[i];
    }
    const double beta = rz2 / rz;
    rz = rz2;
    rr = rr2;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  if (c.debug) std::fprintf(stderr, "  cg: res0 %.2e (vs tol ref) its %d\n", std::sqrt(c.dbg_rr0 / (tol2 / (c.rtol * c.rtol))), it);
  return rr <= tol2 ? it : -it;
}

// z = V(r): V(1,1) cycle on the pressure hierarchy, level 0 = (qq, Ap), levels >= 1 = c.mg[l - 1]; the last level is
// solved exactly with the dense inverse (k_mg_sweep / k_dense_matvec on the GPU arm)
static void mg_vcycle(const Ctx& c, size_t l, const std::vector<double>& b, std::vector<double>& x) {
  const bool fine = (l == 0);
  const Csr& A = fine ? c.qq : c.mg[l - 1].A;
  const std::vector<double>& Av = fine ? c.Ap : c.mg[l - 1].Av;
  const std::vector<double>& dinv = fine ? c.dinvAp : c.mg[l - 1].dinv;
  const int64_t n = A.n_rows;
  x.assign(n, 0.0);
  if (l == c.mg.size()) {  // exact solve
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0;
      const double* row = c.mg_dense.data() + (size_t)i * n;
      for (int64_t j = 0; j < n; ++j) s += row[j] * b[j];
      x[i] = s;
    }
    return;
  }
  const double w = c.mg_omega;
  std::vector<double> r(n), t(n);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] = w * dinv[i] * b[i];  // first sweep from zero
  spmv(A, Av, x.data(), t.data());
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - t[i];
  const Ctx::MgLev& C = c.mg[l];
  std::vector<double> bc(C.A.n_rows), xc;
  spmv(C.R, C.Rv, r.data(), bc.data());
  mg_vcycle(c, l + 1, bc, xc);
  spmv(C.P, C.Pv, xc.data(), t.data());
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] += t[i];
  spmv(A, Av, x.data(), t.data());
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] += w * dinv[i] * (b[i] - t[i]);  // post-smoothing
}

// PCG with the V-cycle as preconditioner (the GPU arm's pcg_mg_solve); true-residual test |r| <= rtol |b|
int cg_mg(const Ctx& c, const double* b, double* x) {
  const Csr& m = c.qq;
  const int64_t n = m.n_rows;
  std::vector<double> r(n), z, p(n), q(n);
  if (c.nonzero) {
    spmv(m, c.Ap, x, q.data());
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - q[i];
  } else {
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) { r[i] = b[i]; x[i] = 0; }
  }
  const double bb = dot(n, b, b);
  double rr = dot(n, r.data(), r.data()), rz = 0;
  const double tol2 = std::max(c.rtol * c.rtol * bb, 1e-100);
  c.dbg_rr0 = rr;
  int it = 0;
  while (rr > tol2 && it < c.maxit) {
    mg_vcycle(c, 0, r, z);
    const double rz2 = dot(n, r.data(), z.data());
    const double beta = it == 0 ? 0.0 : rz2 / rz;
    rz = rz2;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    spmv(m, c.Ap, p.data(), q.data());
    const double alpha = rz / dot(n, p.data(), q.data());
    double rr2 = 0;
#pragma omp parallel for reduction(+ : rr2)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      rr2 += r[i] * r[i];
    }
    rr = rr2;
    ++it;
  }
  if (c.debug) std::fprintf(stderr, "  cg+mg: res0 %.2e its %d\n", std::sqrt(c.dbg_rr0 / std::max(bb, 1e-300)), it);
  return rr <= tol2 ? it : -it;
}

// BiCGStab, left Jacobi (PETSc's default side for bcgs [ext]); convergence on |D^-1 r| <= rtol |D^-1 b|
int bicgstab(const Ctx& c, const Csr& m, const std::vector<double>& vals, const std::vector<double>& dinv,
             const double* b, double* x) {
  const int64_t n = m.n_rows;
  std::vector<double> r(n), rh(n), p(n), v(n), t(n), tmp(n);
  auto apply = [&](const double* in, double* out) {  // out = D^-1 A in
    spmv(m, vals, in, out);
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) out[i] *= dinv[i];
  };
  double bb = 0;
  if (c.nonzero) apply(x, tmp.data());
#pragma omp parallel for reduction(+ : bb)
  for (int64_t i = 0; i < n; ++i) {
    const double bi = dinv[i] * b[i];
    bb += bi * bi;
    if (c.nonzero) r[i] = bi - tmp[i];
    else { r[i] = bi; x[i] = 0; }
    rh[i] = r[i];
    p[i] = r[i];
  }
  double rr = dot(n, r.data(), r.data()), rho = rr, rh2 = rr;
  const double tol2 = std::max(c.rtol * c.rtol * (c.block_rtol && c.bref2 > bb ? c.bref2 : bb), 1e-100);
  const double rr0 = rr;
  int it = 0;
  while (rr > tol2 && it < c.maxit) {
    apply(p.data(), v.data());
    const double alpha = rho / dot(n, rh.data(), v.data());
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) r[i] -= alpha * v[i];  // s
    apply(r.data(), t.data());
    const double tt = dot(n, t.data(), t.data());
    const double omega = tt > 0 ? dot(n, t.data(), r.data()) / tt : 0.0;
    double rr2 = 0, rho2 = 0;
#pragma omp parallel for reduction(+ : rr2, rho2)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i] + omega * r[i];
      r[i] -= omega * t[i];
      rr2 += r[i] * r[i];
      rho2 += rh[i] * r[i];
    }
    rr = rr2;
    ++it;
    if (rr <= tol2) break;
    // (near) breakdown: restart from the current residual (rhat = p = r), as the GPU arm does (linalg.cuh FIN_BCGS_UPDATE)
    if (omega == 0.0 || rho == 0.0 || rho2 * rho2 <= 1e-24 * rr * rh2) {
      rho = rr;
      rh2 = rr;
#pragma omp parallel for
      for (int64_t i = 0; i < n; ++i) { rh[i] = r[i]; p[i] = r[i]; }
      continue;
    }
    const double beta = (rho2 / rho) * (alpha / omega);
    rho = rho2;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
  }
  if (c.debug) std::fprintf(stderr, "  bcgs: res0 %.2e (vs tol ref) its %d\n", std::sqrt(rr0 / (tol2 / (c.rtol * c.rtol))), it);
  return rr <= tol2 ? it : -it;
}

// row-wise (gather) assembly of a square operator on V or Q: kind 0 mass, 1 stiffness, 2 convection
void assemble_square(Ctx& c, bool onV, int kind, std::vector<double>& vals) {
  const Tables& t = c.t;
  const int d = c.d, nd = onV ? t.nv : t.nq;
  const Csr& pat = onV ? c.vv : c.qq;
  const std::vector<int>& cd = onV ? c.vd : c.qd;
  const auto& aptr = onV ? c.vadj_ptr : c.qadj_ptr;
  const auto& acell = onV ? c.vadj_cell : c.qadj_cell;
  const auto& aloc = onV ? c.vadj_loc : c.qadj_loc;
  const double* Mref = onV ? t.MV : t.MQ;
  const double* Sref = onV ? t.SV : t.SQ;
  vals.assign(pat.col.size(), 0.0);
#pragma omp parallel for schedule(dynamic, 512)
  for (int r = 0; r < pat.n_rows; ++r) {
    for (int a = aptr[r]; a < aptr[r + 1]; ++a) {
      const int64_t cell = acell[a];
      const int i = aloc[a];
      const Geo g = geometry(c, cell);
      const int* dofs = &cd[cell * nd];
      double row[10];
      if (kind == 0) {
        for (int j = 0; j < nd; ++j) row[j] = g.detJ * Mref[i * nd + j];
      } else if (kind == 1) {
        double G[3][3];
        for (int a2 = 0; a2 < d; ++a2)
          for (int b2 = 0; b2 < d; ++b2) {
            double s = 0;
            for (int k = 0; k < d; ++k) s += g.Kinv[a2][k] * g.Kinv[b2][k];
            G[a2][b2] = s * g.detJ;
          }
        for (int j = 0; j < nd; ++j) {
          double s = 0;
          for (int a2 = 0; a2 < d; ++a2)
            for (int b2 = 0; b2 < d; ++b2) s += G[a2][b2] * Sref[((a2 * d + b2) * nd + i) * nd + j];
          row[j] = s;
        }
      } else {  // C[i,j] = |detJ| sum_{a,dl} w[a][dl] T[a][dl][i][j]   (fracstep.py:355-358)
        for (int j = 0; j < nd; ++j) row[j] = 0;
        for (int a2 = 0; a2 < nd; ++a2)
          for (int dl = 0; dl < d; ++dl) {
            double w = 0;
            for (int k = 0; k < d; ++k) w += g.Kinv[dl][k] * c.uab[k][dofs[a2]];
            w *= g.detJ;
            const double* Trow = &t.T[(((size_t)a2 * d + dl) * nd + i) * nd];
            for (int j = 0; j < nd; ++j) row[j] += w * Trow[j];
          }
      }
      for (int j = 0; j < nd; ++j) vals[pat.find(r, dofs[j])] += row[j];
    }
  }
}

void assemble_rect(Ctx& c) {
  const Tables& t = c.t;
  const int d = c.d;
  for (int k = 0; k < d; ++k) {
    c.P[k].assign(c.vq.col.size(), 0.0);
    c.G[k].assign(c.vq.col.size(), 0.0);
    c.D[k].assign(c.qv.col.size(), 0.0);
  }
#pragma omp parallel for schedule(dynamic, 512)
  for (int r = 0; r < c.vq.n_rows; ++r)
    for (int a = c.vadj_ptr[r]; a < c.vadj_ptr[r + 1]; ++a) {
      const int64_t cell = c.vadj_cell[a];
      const int j = c.vadj_loc[a];
      const Geo g = geometry(c, cell);
      for (int q = 0; q < t.nq; ++q) {
        const int pos = c.vq.find(r, c.qd[cell * t.nq + q]);
        for (int k = 0; k < d; ++k) {
          double pv = 0, gv = 0;
          for (int dl = 0; dl < d; ++dl) {
            pv += g.Kinv[dl][k] * t.PX[(dl * t.nv + j) * t.nq + q];
            gv += g.Kinv[dl][k] * t.GX[(dl * t.nv + j) * t.nq + q];
          }
          c.P[k][pos] += g.detJ * pv;
          c.G[k][pos] += g.detJ * gv;
        }
      }
    }
#pragma omp parallel for schedule(dynamic, 512)
  for (int r = 0; r < c.qv.n_rows; ++r)
    for (int a = c.qadj_ptr[r]; a < c.qadj_ptr[r + 1]; ++a) {
      const int64_t cell = c.qadj_cell[a];
      const int q = c.qadj_loc[a];
      const Geo g = geometry(c, cell);
      for (int j = 0; j < t.nv; ++j) {
        const int pos = c.qv.find(r, c.vd[cell * t.nv + j]);
        for (int k = 0; k < d; ++k) {
          double pv = 0;
          for (int dl = 0; dl < d; ++dl) pv += g.Kinv[dl][k] * t.PX[(dl * t.nv + j) * t.nq + q];
          c.D[k][pos] += g.detJ * pv;
        }
      }
    }
}

void inv_diag(const Csr& m, const std::vector<double>& v, std::vector<double>& dinv) {
  dinv.assign(m.n_rows, 1.0);
#pragma omp parallel for
  for (int r = 0; r < m.n_rows; ++r) dinv[r] = 1.0 / v[m.find(r, r)];
}

}  // namespace

extern "C" {

int ipcs_cpu_threads(void) { return omp_get_max_threads(); }

// The OpenMP team size is set EXPLICITLY by the caller (bench.py passes the size of the process' CPU affinity mask):
// torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which must not decide how many cores the CPU
// baseline runs on.
void ipcs_cpu_set_threads(int n) {
  if (n > 0) {
    omp_set_dynamic(0);
    omp_set_num_threads(n);
  }
}

// Interleave this process' future page allocations over all NUMA nodes (set_mempolicy(MPOL_INTERLEAVE), raw syscall:
// no libnuma in the image).  The std::vector storage of this port is first touched by the constructing thread, so
// without this every array of a multi-socket box lives on one node and the other socket's threads pull it over the
// interconnect.  Returns the number of nodes interleaved over (<= 1: nothing was changed).
int ipcs_cpu_numa_interleave(void) {
#if defined(__linux__)
  int n_nodes = 0;
  for (int i = 0; i < 64; ++i) {
    char path[96];
    std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d", i);
    if (access(path, F_OK) == 0) n_nodes = i + 1;
  }
  if (n_nodes <= 1) return n_nodes;
  unsigned long mask = n_nodes >= 64 ? ~0ul : ((1ul << n_nodes) - 1);
  const long rc = syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, &mask, (unsigned long)(8 * sizeof(mask)));
  return rc == 0 ? n_nodes : 0;
#else
  return 0;
#endif
}

// STREAM triad a = b + s c on `n` doubles per array, arrays first-touched by the threads that stream them:
// best of `reps`, GB/s counting 24 bytes per element (the host-memory roofline the CPU baseline is held against)
double ipcs_cpu_stream_triad(int64_t n, int reps) {
  double* a = (double*)std::malloc(sizeof(double) * (size_t)n);
  double* b = (double*)std::malloc(sizeof(double) * (size_t)n);
  double* cc = (double*)std::malloc(sizeof(double) * (size_t)n);
  if (!a || !b || !cc) { std::free(a); std::free(b); std::free(cc); return 0.0; }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) { a[i] = 0.0; b[i] = 1.0; cc[i] = 2.0; }
  double best = 0.0;
  for (int r = 0; r < reps; ++r) {
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) a[i] = b[i] + 3.0 * cc[i];
    const double dt = omp_get_wtime() - t0;
    if (dt > 0) best = std::max(best, 24.0 * (double)n / dt / 1e9);
  }
  volatile double sink = a[n / 2];
  (void)sink;
  std::free(a); std::free(b); std::free(cc);
  return best;
}

void* ipcs_cpu_create(int gdim, int deg_v, int64_t n_nodes, const double* x, int64_t n_cells, const int* cell_nodes,
                      int64_t nV, const int* vdofs, int64_t nQ, const int* qdofs) {
  Ctx* c = new Ctx();
  if (const char* e = std::getenv("IPCS_DEBUG")) c->debug = std::atoi(e);
  if (const char* e = std::getenv("IPCS_GUESS_T")) c->guess_t = std::atoi(e);
  if (const char* e = std::getenv("IPCS_GUESS_M")) c->guess_m = std::atoi(e);
  if (const char* e = std::getenv("IPCS_GUESS_P")) c->guess_p = std::atoi(e);
  c->d = gdim;
  c->degv = deg_v;
  c->t = tables_for(gdim, deg_v);
  c->n_nodes = n_nodes;
  c->n_cells = n_cells;
  c->nV = nV;
  c->nQ = nQ;
  c->x.assign(x, x + 3 * n_nodes);
  c->cn.assign(cell_nodes, cell_nodes + n_cells * (gdim + 1));
  c->vd.assign(vdofs, vdofs + n_cells * c->t.nv);
  c->qd.assign(qdofs, qdofs + n_cells * c->t.nq);
  build_adj(nV, n_cells, c->t.nv, c->vd, c->vadj_ptr, c->vadj_cell, c->vadj_loc);
  build_adj(nQ, n_cells, c->t.nq, c->qd, c->qadj_ptr, c->qadj_cell, c->qadj_loc);
  build_pattern(nV, nV, c->vadj_ptr, c->vadj_cell, c->t.nv, c->vd, c->vv);
  build_pattern(nV, nQ, c->vadj_ptr, c->vadj_cell, c->t.nq, c->qd, c->vq);
  build_pattern(nQ, nV, c->qadj_ptr, c->qadj_cell, c->t.nv, c->vd, c->qv);
  build_pattern(nQ, nQ, c->qadj_ptr, c->qadj_cell, c->t.nq, c->qd, c->qq);
  for (int k = 0; k < gdim; ++k)
    for (auto* v : {&c->u[k], &c->u1[k], &c->u2[k], &c->uab[k], &c->rhs1[k], &c->bfirst[k], &c->b0[k]}) v->assign(nV, 0.0);
  c->wrk.assign(nV, 0.0);
  c->b3.assign(nV, 0.0);
  for (auto* v : {&c->ps, &c->p, &c->dp, &c->b2, &c->mQ}) v->assign(nQ, 0.0);
  c->is_bc.assign(nV, 0);
  return c;
}

void ipcs_cpu_destroy(void* h) { delete (Ctx*)h; }

int64_t ipcs_cpu_nnz(void* h, int which) {
  Ctx* c = (Ctx*)h;
  const Csr* p[4] = {&c->vv, &c->vq, &c->qv, &c->qq};
  return (int64_t)p[which]->col.size();
}

void ipcs_cpu_get_pattern(void* h, int which, int* indptr, int* indices) {
  Ctx* c = (Ctx*)h;
  const Csr* p[4] = {&c->vv, &c->vq, &c->qv, &c->qq};
  std::copy(p[which]->ptr.begin(), p[which]->ptr.end(), indptr);
  std::copy(p[which]->col.begin(), p[which]->col.end(), indices);
}

void ipcs_cpu_set_bc(void* h, int comp, int64_t n, const int* dofs) {
  Ctx* c = (Ctx*)h;
  c->bc_dofs[comp].assign(dofs, dofs + n);
  c->bc_vals[comp].assign(n, 0.0);
  if (comp == 0)  // the shared matrix takes component 0's dof set (fracstep.py:470-472)
    for (int64_t i = 0; i < n; ++i) c->is_bc[dofs[i]] = 1;
}

void ipcs_cpu_set_bc_values(void* h, int comp, const double* vals) {
  Ctx* c = (Ctx*)h;
  std::copy(vals, vals + c->bc_vals[comp].size(), c->bc_vals[comp].begin());
}

void ipcs_cpu_set_options(void* h, double rtol, int maxit, int nonzero_guess) {
  Ctx* c = (Ctx*)h;
  c->rtol = rtol;
  c->maxit = maxit;
  c->nonzero = nonzero_guess;
}
// time extrapolation of order `order` from a history (h[0] newest): h0 / 2h0 - h1 / 3h0 - 3h1 + h2, limited by
// the number of stored states
static void extrap(int order, const std::vector<double>* h, int n_avail, int64_t n, double* out, double scale_add = 0.0,
                   const double* add = nullptr) {
  order = std::min(order, n_avail - 1);
  const double c0 = order == 2 ? 3.0 : (order == 1 ? 2.0 : 1.0), c1 = order == 2 ? -3.0 : (order == 1 ? -1.0 : 0.0),
               c2 = order == 2 ? 1.0 : 0.0;
#pragma omp parallel for
  for (int64_t i = 0; i < n; ++i) {
    double v = c0 * h[0][i];
    if (order >= 1) v += c1 * h[1][i];
    if (order >= 2) v += c2 * h[2][i];
    out[i] = v + (add ? scale_add * add[i] : 0.0);
  }
}
static void push_hist(std::vector<double>* h, const std::vector<double>& v) {
  h[2].swap(h[1]);
  h[1].swap(h[0]);
  h[0] = v;
}

// pressure multigrid: one coarser level per call (CSR of its operator, of P from the previous level and of R = P^T),
// then the dense inverse of the last level's shifted operator; ipcs_cpu_set_pressure_mg switches PCG to it
void ipcs_cpu_mg_add_level(void* h, int n, const int* Ap, const int* Ai, const double* Av, int n_fine, const int* Pp,
                           const int* Pi, const double* Pv, const int* Rp, const int* Ri, const double* Rv) {
  Ctx* c = (Ctx*)h;
  c->mg.emplace_back();
  Ctx::MgLev& L = c->mg.back();
  auto fill = [](Csr& m, std::vector<double>& v, int rows, int cols, const int* p, const int* i, const double* a) {
    m.n_rows = rows;
    m.n_cols = cols;
    m.ptr.assign(p, p + rows + 1);
    m.col.assign(i, i + p[rows]);
    v.assign(a, a + p[rows]);
  };
  fill(L.A, L.Av, n, n, Ap, Ai, Av);
  fill(L.P, L.Pv, n_fine, n, Pp, Pi, Pv);
  fill(L.R, L.Rv, n, n_fine, Rp, Ri, Rv);
  L.dinv.assign(n, 1.0);
  for (int r = 0; r < n; ++r)
    for (int q = L.A.ptr[r]; q < L.A.ptr[r + 1]; ++q)
      if (L.A.col[q] == r) L.dinv[r] = 1.0 / L.Av[q];
}
void ipcs_cpu_mg_set_dense(void* h, int n, const double* inv) {
  Ctx* c = (Ctx*)h;
  c->mg_dense.assign(inv, inv + (size_t)n * n);
}
void ipcs_cpu_set_pressure_mg(void* h, int on, double omega) {
  Ctx* c = (Ctx*)h;
  c->pressure_mg = on;
  if (omega > 0) c->mg_omega = omega;
}

// tolerance of the velocity solves relative to max_k |b_k| (same option as the GPU arm's b200_block_rtol)
void ipcs_cpu_set_block_rtol(void* h, int on) { ((Ctx*)h)->block_rtol = on; }
// same initial guesses as the GPU arm's b200_guess=extrapolate: 2u^n - u^{n-1} / u* + (u - u*)^{n-1}
// order 2 = the GPU arm's b200_guess=extrapolate2: quadratic extrapolation of the u* and (u - u*) histories
void ipcs_cpu_set_extrapolate(void* h, int order) {
  Ctx* c = (Ctx*)h;
  c->extrapolate = order > 0;
  if (order >= 2) { c->guess_t = 2; c->guess_m = 2; }
}

// _preassemble (fracstep.py:360-409), no pressure BCs (the Taylor-Green / cavity configuration)
void ipcs_cpu_preassemble(void* h, const double* f) {
  Ctx* c = (Ctx*)h;
  assemble_square(*c, true, 0, c->M);
  assemble_square(*c, true, 1, c->K);
  assemble_square(*c, false, 1, c->Ap);
  assemble_rect(*c);
  c->A.assign(c->vv.col.size(), 0.0);
  const Tables& t = c->t;
  for (int64_t cell = 0; cell < c->n_cells; ++cell) {
    const Geo g = geometry(*c, cell);
    for (int j = 0; j < t.nv; ++j)
      for (int k = 0; k < c->d; ++k) c->b0[k][c->vd[cell * t.nv + j]] += (f ? f[k] : 0.0) * g.detJ * t.LV[j];
    for (int q = 0; q < t.nq; ++q) c->mQ[c->qd[cell * t.nq + q]] += g.detJ * t.LQ[q];
  }
  c->vol = 0;
  for (double v : c->mQ) c->vol += v;
  inv_diag(c->vv, c->M, c->dinvM);
  inv_diag(c->qq, c->Ap, c->dinvAp);
}

static std::vector<double>* vec_of(Ctx* c, int which, int comp) {
  switch (which) {
    case 0: return &c->u[comp];
    case 1: return &c->u1[comp];
    case 2: return &c->u2[comp];
    case 3: return &c->p;
    case 4: return &c->ps;
    case 5: return &c->dp;
    case 6: return &c->rhs1[comp];
    case 7: return &c->bfirst[comp];
    case 8: return &c->b2;
    case 9: return &c->uab[comp];
    default: return nullptr;
  }
}
void ipcs_cpu_set_vec(void* h, int which, int comp, const double* v) {
  auto* d = vec_of((Ctx*)h, which, comp);
  std::copy(v, v + d->size(), d->begin());
}
void ipcs_cpu_get_vec(void* h, int which, int comp, double* v) {
  auto* d = vec_of((Ctx*)h, which, comp);
  std::copy(d->begin(), d->end(), v);
}
void ipcs_cpu_get_matrix(void* h, int which, int comp, double* v) {
  Ctx* c = (Ctx*)h;
  const std::vector<double>* m[7] = {&c->M, &c->K, &c->A, &c->Ap, &c->P[comp], &c->G[comp], &c->D[comp]};
  std::copy(m[which]->begin(), m[which]->end(), v);
}

// demo/assembly_strategies.py:56-152 for ONE scalar field (component `comp` of u1, convecting field uab), as the
// reference times it: the convection matrix is assembled first (untimed there), then
//   matvec strategy (:128-133): A.scale(-0.5); A.axpy(1/dt, M); A.axpy(-nu/2, K); A.mult(u_1, b)  -- four CSR passes
//   action strategy (:137-140): assemble_vector(action(lhs, u_1))                                  -- one cell loop
// out[0..2] = seconds {convection assembly, matvec, action} (best of reps); b_matvec / b_action receive the vectors.
void ipcs_cpu_bench_strategies(void* h, int comp, double dt, double nu, int reps, double* out, double* b_matvec,
                               double* b_action) {
  Ctx* c = (Ctx*)h;
  const Tables& t = c->t;
  const int d = c->d, nd = t.nv;
  const int64_t nV = c->nV, nnz = (int64_t)c->vv.col.size();
  out[0] = out[1] = out[2] = 1e300;
  std::vector<double> A, bm(nV), ba(nV);
  for (int rep = 0; rep < reps; ++rep) {
    double t0 = omp_get_wtime();
    assemble_square(*c, true, 2, A);  // zeroEntries + assemble_matrix(convection)
    out[0] = std::min(out[0], omp_get_wtime() - t0);
    t0 = omp_get_wtime();
#pragma omp parallel for
    for (int64_t p = 0; p < nnz; ++p) A[p] *= -0.5;
#pragma omp parallel for
    for (int64_t p = 0; p < nnz; ++p) A[p] += (1.0 / dt) * c->M[p];
#pragma omp parallel for
    for (int64_t p = 0; p < nnz; ++p) A[p] += (-0.5 * nu) * c->K[p];
    spmv(c->vv, A, c->u1[comp].data(), bm.data());
    out[1] = std::min(out[1], omp_get_wtime() - t0);
    t0 = omp_get_wtime();
#pragma omp parallel for
    for (int64_t i = 0; i < nV; ++i) ba[i] = 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t cell = 0; cell < c->n_cells; ++cell) {
      const Geo g = geometry(*c, cell);
      const int* dofs = &c->vd[cell * nd];
      double G[3][3], w[10][3], ue[10], be[10];
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
          double sacc = 0;
          for (int k = 0; k < d; ++k) sacc += g.Kinv[a][k] * g.Kinv[b][k];
          G[a][b] = sacc * g.detJ;
        }
      for (int a = 0; a < nd; ++a) {
        ue[a] = c->u1[comp][dofs[a]];
        for (int dl = 0; dl < d; ++dl) {
          double sacc = 0;
          for (int k = 0; k < d; ++k) sacc += g.Kinv[dl][k] * c->uab[k][dofs[a]];
          w[a][dl] = sacc * g.detJ;
        }
      }
      for (int i = 0; i < nd; ++i) {
        double acc = 0;
        for (int j = 0; j < nd; ++j) {
          double cv = 0, kv = 0;
          for (int a = 0; a < nd; ++a)
            for (int dl = 0; dl < d; ++dl) cv += w[a][dl] * t.T[(((size_t)a * d + dl) * nd + i) * nd + j];
          for (int a = 0; a < d; ++a)
            for (int b = 0; b < d; ++b) kv += G[a][b] * t.SV[((a * d + b) * nd + i) * nd + j];
          acc += ((1.0 / dt) * g.detJ * t.MV[i * nd + j] - 0.5 * nu * kv - 0.5 * cv) * ue[j];
        }
        be[i] = acc;
      }
      for (int i = 0; i < nd; ++i) {
#pragma omp atomic
        ba[dofs[i]] += be[i];
      }
    }
    out[2] = std::min(out[2], omp_get_wtime() - t0);
  }
  std::copy(bm.begin(), bm.end(), b_matvec);
  std::copy(ba.begin(), ba.end(), b_action);
}

// one time step, fracstep.py:660-696 with max_iter = 1; returns 0 or a negative stage id on divergence
int ipcs_cpu_step(void* h, double dt, double nu, int* its) {
  Ctx* c = (Ctx*)h;
  const int d = c->d;
  const int64_t nV = c->nV, nQ = c->nQ;
  c->ps = c->p;  // :673
  // ---- assemble_first (:411-472)
  for (int k = 0; k < d; ++k) {
#pragma omp parallel for
    for (int64_t i = 0; i < nV; ++i) c->uab[k][i] = 1.5 * c->u1[k][i] - 0.5 * c->u2[k][i];
  }
  assemble_square(*c, true, 2, c->A);  // A = C(uab)
  const int64_t nnz = (int64_t)c->A.size();
#pragma omp parallel for
  for (int64_t p = 0; p < nnz; ++p) c->A[p] = -0.5 * c->A[p] + (1.0 / dt) * c->M[p] + (-0.5 * nu) * c->K[p];  // :438-442
  for (int k = 0; k < d; ++k) {
    spmv(c->vv, c->A, c->u1[k].data(), c->wrk.data());  // :452
#pragma omp parallel for
    for (int64_t i = 0; i < nV; ++i) c->bfirst[k][i] = c->wrk[i] + c->b0[k][i];
  }
#pragma omp parallel for
  for (int64_t p = 0; p < nnz; ++p) c->A[p] = -c->A[p] + (2.0 / dt) * c->M[p];  // :468-469
#pragma omp parallel for
  for (int r = 0; r < c->vv.n_rows; ++r)
    if (c->is_bc[r])
      for (int p = c->vv.ptr[r]; p < c->vv.ptr[r + 1]; ++p) c->A[p] = (c->vv.col[p] == r) ? 1.0 : 0.0;  // :471-472
  inv_diag(c->vv, c->A, c->dinvA);
  // ---- tentative velocity (:474-525)
  c->its_t = 0;
  for (int k = 0; k < d; ++k) {
    spmv(c->vq, c->P[k], c->ps.data(), c->wrk.data());
#pragma omp parallel for
    for (int64_t i = 0; i < nV; ++i) c->rhs1[k][i] = c->bfirst[k][i] + c->wrk[i];
    for (size_t i = 0; i < c->bc_dofs[k].size(); ++i) c->rhs1[k][c->bc_dofs[k][i]] = c->bc_vals[k][i];
  }
  c->bref2 = 0;
  for (int k = 0; k < d; ++k) {
    double s2 = 0;
#pragma omp parallel for reduction(+ : s2)
    for (int64_t i = 0; i < nV; ++i) s2 += (c->dinvA[i] * c->rhs1[k][i]) * (c->dinvA[i] * c->rhs1[k][i]);
    c->bref2 = std::max(c->bref2, s2);
  }
  for (int k = 0; k < d; ++k) {
    if (c->extrapolate && c->nonzero && c->steps_done >= 1) {
      if (c->guess_t > 0 && c->n_hist >= 1) {
        extrap(c->guess_t, c->ustar_hist[k], c->n_hist, nV, c->u[k].data());
      } else {
#pragma omp parallel for
      for (int64_t i = 0; i < nV; ++i)
        c->u[k][i] = 2.0 * c->u1[k][i] - c->u2[k][i] - (c->delta_prev[k].empty() ? 0.0 : c->delta_prev[k][i]);
      }
    }
    int it = bicgstab(*c, c->vv, c->A, c->dinvA, c->rhs1[k].data(), c->u[k].data());
    if (it < 0) return -1;
    c->its_t = std::max(c->its_t, it);
    if (c->guess_t > 0) push_hist(c->ustar_hist[k], c->u[k]);
  }
  if (c->guess_t > 0) c->n_hist = std::min(c->n_hist + 1, 3);
  // ---- pressure correction (:527-605)
  std::fill(c->b2.begin(), c->b2.end(), 0.0);
  std::vector<double> wq(nQ);
  for (int k = 0; k < d; ++k) {
    spmv(c->qv, c->D[k], c->u[k].data(), wq.data());
#pragma omp parallel for
    for (int64_t i = 0; i < nQ; ++i) c->b2[i] += wq[i];
  }
  double mean = 0;
#pragma omp parallel for reduction(+ : mean)
  for (int64_t i = 0; i < nQ; ++i) {
    c->b2[i] *= -1.0 / dt;
    mean += c->b2[i];
  }
  mean /= (double)nQ;
#pragma omp parallel for
  for (int64_t i = 0; i < nQ; ++i) c->b2[i] -= mean;  // nullspace.remove, :573-574
  c->bref2 = 0;
  if (c->extrapolate && c->nonzero) {  // start from 2 dp^{n-1} - dp^{n-2}
    if (c->dp_old.empty()) c->dp_old.assign(nQ, 0.0);
    std::vector<double> prev = c->dp;
    if (c->guess_p != 1 && c->n_dp_hist >= 1) {
      extrap(c->guess_p, c->dp_hist3, c->n_dp_hist, nQ, c->dp.data());
    } else if (c->dp_hist >= 2) {
#pragma omp parallel for
      for (int64_t i = 0; i < nQ; ++i) c->dp[i] = 2.0 * prev[i] - c->dp_old[i];
    }
    c->dp_old = prev;
    c->dp_hist++;
  }
  c->its_p = (c->pressure_mg && !c->mg.empty()) ? cg_mg(*c, c->b2.data(), c->dp.data())
                                                : cg(*c, c->qq, c->Ap, c->dinvAp, c->b2.data(), c->dp.data());
  if (c->its_p < 0) return -2;
  if (c->guess_p != 1) { push_hist(c->dp_hist3, c->dp); c->n_dp_hist = std::min(c->n_dp_hist + 1, 3); }
  const double avg = dot(nQ, c->mQ.data(), c->dp.data()) / c->vol;  // :579-591
#pragma omp parallel for
  for (int64_t i = 0; i < nQ; ++i) {
    c->dp[i] -= avg;
    c->ps[i] = c->p[i] + c->dp[i];  // :604
  }
  // ---- velocity update (:607-658)
  c->its_u = 0;
  std::vector<std::vector<double>> b3s(d, std::vector<double>(nV));
  c->bref2 = 0;
  for (int k = 0; k < d; ++k) {
    spmv(c->vv, c->M, c->u[k].data(), b3s[k].data());
    spmv(c->vq, c->G[k], c->dp.data(), c->wrk.data());
#pragma omp parallel for
    for (int64_t i = 0; i < nV; ++i) b3s[k][i] -= dt * c->wrk[i];
    c->bref2 = std::max(c->bref2, dot(nV, b3s[k].data(), b3s[k].data()));
  }
  for (int k = 0; k < d; ++k) {
    c->b3 = b3s[k];
    std::vector<double> ustar;
    if (c->extrapolate && c->nonzero) {
      if (c->delta_prev[k].empty()) c->delta_prev[k].assign(nV, 0.0);
      ustar = c->u[k];
      if (c->guess_m > 0 && !c->delta_hist[k][0].empty()) {
        int avail = 1 + !c->delta_hist[k][1].empty() + (!c->delta_hist[k][1].empty() && !c->delta_hist[k][2].empty());
        extrap(c->guess_m, c->delta_hist[k], avail, nV, c->u[k].data(), 1.0, ustar.data());
      } else {
#pragma omp parallel for
      for (int64_t i = 0; i < nV; ++i) c->u[k][i] += c->delta_prev[k][i];
      }
    }
    int it = cg(*c, c->vv, c->M, c->dinvM, c->b3.data(), c->u[k].data());
    if (!ustar.empty()) {
#pragma omp parallel for
      for (int64_t i = 0; i < nV; ++i) c->delta_prev[k][i] = c->u[k][i] - ustar[i];
      if (c->guess_m > 0) push_hist(c->delta_hist[k], c->delta_prev[k]);
    }
    if (it < 0) return -3;
    c->its_u = std::max(c->its_u, it);
  }
  for (int k = 0; k < d; ++k) {  // :689-693
    c->u2[k] = c->u1[k];
    c->u1[k] = c->u[k];
  }
  c->p = c->ps;
  c->steps_done++;
  if (its) {
    its[0] = c->its_t;
    its[1] = c->its_p;
    its[2] = c->its_u;
  }
  return 0;
}

}  // extern "C"
