// linalg.cuh -- CSR SpMM (multi-RHS SpMV), the fused "matrix-vector strategy" kernels of
// fracstep.py:438-472, rectangular P2xP1 products, vector kernels and the Krylov recurrences
// (PCG / BiCGStab on up to 3 systems that share one matrix) that replace PETSc Mat/Vec/KSP
// (SURVEY.md N4-N9).
#pragma once
#include "common.cuh"

// ---- Krylov state, device resident ------------------------------------------------------
// One instance per solve; scalars are produced by the last block of the reducing kernel
// (grid_reduce) and consumed by the next kernel on the stream: no host round trip per iteration.
struct KryState {
  double rho[B2_MAXK], alpha[B2_MAXK], beta[B2_MAXK], omega[B2_MAXK];
  double rz[B2_MAXK], bb[B2_MAXK], rr[B2_MAXK], tol2[B2_MAXK];
  int active[B2_MAXK], reason[B2_MAXK], its[B2_MAXK];
  int done, maxit, K, pad;
  double rtol, atol;
};

enum {
  FIN_NONE = 0,
  FIN_CG_INIT,
  FIN_CG_PQ,
  FIN_CG_UPDATE,
  FIN_BCGS_INIT,
  FIN_BCGS_V,
  FIN_BCGS_T,
  FIN_BCGS_UPDATE,
  FIN_STORE  // just store totals into st->rr[0..] (generic sums read back by the host)
};

__device__ __forceinline__ bool b2_bad(double v) { return !(fabs(v) <= 1.79e308); }

__device__ inline void kry_check_done(KryState* st) {
  int any = 0;
  for (int k = 0; k < st->K; ++k) any |= st->active[k];
  st->done = !any;
}

__device__ inline void kry_converge_test(KryState* st, int k) {
  if (!st->active[k]) return;
  double rr = st->rr[k];
  if (b2_bad(rr)) {
    st->reason[k] = -9;  // KSP_DIVERGED_NANORINF
    st->active[k] = 0;
  } else if (rr <= st->tol2[k]) {
    st->reason[k] = (rr <= st->atol * st->atol && st->atol * st->atol >= st->rtol * st->rtol * st->bb[k]) ? 3 : 2;
    st->active[k] = 0;
  } else if (st->its[k] >= st->maxit) {
    st->reason[k] = -3;  // KSP_DIVERGED_ITS
    st->active[k] = 0;
  }
}

// Runs in thread 0 of the last block.  `t` holds the reduced sums in the order documented at
// each case.
__device__ inline void kry_finalize(int fin, KryState* st, const double* t) {
  const int K = st->K;
  switch (fin) {
    case FIN_CG_INIT:  // t = rz[K], bb[K], rr[K]
      for (int k = 0; k < K; ++k) {
        st->rz[k] = t[k];
        st->bb[k] = t[K + k];
        st->rr[k] = t[2 * K + k];
        double a2 = st->atol * st->atol, r2 = st->rtol * st->rtol * st->bb[k];
        st->tol2[k] = r2 > a2 ? r2 : a2;
        st->its[k] = 0;
        st->reason[k] = 0;
        st->active[k] = 1;
        st->beta[k] = 0.0;
        kry_converge_test(st, k);
      }
      kry_check_done(st);
      break;
    case FIN_CG_PQ:  // t = pq[K]
      for (int k = 0; k < K; ++k) {
        if (!st->active[k]) continue;
        double pq = t[k];
        if (b2_bad(pq)) { st->reason[k] = -9; st->active[k] = 0; }
        else if (pq == 0.0) { st->reason[k] = -5; st->active[k] = 0; }  // KSP_DIVERGED_BREAKDOWN
        else if (pq < 0.0) { st->reason[k] = -8; st->active[k] = 0; }   // KSP_DIVERGED_INDEFINITE_PC/MAT
        else st->alpha[k] = st->rz[k] / pq;
      }
      kry_check_done(st);
      break;
    case FIN_CG_UPDATE:  // t = rz_new[K], rr[K]
      for (int k = 0; k < K; ++k) {
        if (!st->active[k]) continue;
        st->beta[k] = t[k] / st->rz[k];
        st->rz[k] = t[k];
        st->rr[k] = t[K + k];
        st->its[k] += 1;
        kry_converge_test(st, k);
      }
      kry_check_done(st);
      break;
    case FIN_BCGS_INIT:  // t = bb[K], rr[K]
      for (int k = 0; k < K; ++k) {
        st->bb[k] = t[k];
        st->rr[k] = t[K + k];
        st->rho[k] = t[K + k];  // rhat = r0
        st->alpha[k] = 1.0;
        st->omega[k] = 1.0;
        st->beta[k] = 0.0;
        double a2 = st->atol * st->atol, r2 = st->rtol * st->rtol * st->bb[k];
        st->tol2[k] = r2 > a2 ? r2 : a2;
        st->its[k] = 0;
        st->reason[k] = 0;
        st->active[k] = 1;
        kry_converge_test(st, k);
      }
      kry_check_done(st);
      break;
    case FIN_BCGS_V:  // t = (rhat . v)[K]
      for (int k = 0; k < K; ++k) {
        if (!st->active[k]) continue;
        double d = t[k];
        if (b2_bad(d)) { st->reason[k] = -9; st->active[k] = 0; }
        else if (d == 0.0) { st->reason[k] = -5; st->active[k] = 0; }
        else st->alpha[k] = st->rho[k] / d;
      }
      kry_check_done(st);
      break;
    case FIN_BCGS_T:  // t = (t . s)[K], (t . t)[K]
      for (int k = 0; k < K; ++k) {
        if (!st->active[k]) continue;
        double tt = t[K + k];
        st->omega[k] = tt > 0.0 ? t[k] / tt : 0.0;
      }
      break;
    case FIN_BCGS_UPDATE:  // t = rr[K], rho_new[K]
      for (int k = 0; k < K; ++k) {
        if (!st->active[k]) continue;
        st->rr[k] = t[k];
        st->its[k] += 1;
        kry_converge_test(st, k);
        if (!st->active[k]) continue;
        double rn = t[K + k];
        if (st->omega[k] == 0.0 || st->rho[k] == 0.0 || rn == 0.0) {
          st->reason[k] = -5;
          st->active[k] = 0;
          continue;
        }
        st->beta[k] = (rn / st->rho[k]) * (st->alpha[k] / st->omega[k]);
        st->rho[k] = rn;
      }
      kry_check_done(st);
      break;
    case FIN_STORE:
    default:
      break;
  }
}

// ---- CSR SpMM: y[row][k] = sum_p vals[p] * x[cols[p]][k] (* colscale[cols[p]]) ---------------
// LPR lanes cooperate on one row (rows of P2 matrices hold ~28 entries, P1 ~15); a warp therefore
// streams 32/LPR consecutive rows: the value/column loads of one warp instruction cover a
// contiguous stretch of the CSR arrays.  Persistent grid (grid-stride over row groups) so the
// number of partial sums for the fused dot products stays small.
//   DOT == 0: no reduction          DOT == 1: sums[k] = y_k . w_k
//   DOT == 2: sums[k] = y_k . w_k , sums[K+k] = y_k . y_k
template <int K, int LPR, bool SCALE, int DOT>
__global__ void __launch_bounds__(256)
k_spmm(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
       const double* __restrict__ vals, const double* __restrict__ x,
       const double* __restrict__ colscale, double* __restrict__ y, const double* __restrict__ w,
       KryState* st, int fin, double* partials, unsigned* counter) {
  if (st != nullptr && st->done) return;
  const int lane = threadIdx.x % LPR;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int ngroups = (gridDim.x * blockDim.x) / LPR;
  constexpr int ND = DOT == 0 ? 1 : DOT * K;
  double dots[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) dots[i] = 0.0;
  const int n_iter = (n_rows + ngroups - 1) / ngroups;
  for (int it = 0; it < n_iter; ++it) {
    const int row = it * ngroups + group;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    if (row < n_rows) {
      const int end = __ldg(rowptr + row + 1);
      for (int p = __ldg(rowptr + row) + lane; p < end; p += LPR) {
        const int c = __ldg(cols + p);
        double v = __ldg(vals + p);
        if constexpr (SCALE) v *= __ldg(colscale + c);
        const double* xc = x + (size_t)c * K;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fma(v, __ldg(xc + k), acc[k]);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (row < n_rows && lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) y[(size_t)row * K + k] = acc[k];
      if constexpr (DOT >= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) dots[k] = fma(acc[k], w[(size_t)row * K + k], dots[k]);
      }
      if constexpr (DOT == 2) {
#pragma unroll
        for (int k = 0; k < K; ++k) dots[K + k] = fma(acc[k], acc[k], dots[K + k]);
      }
    }
  }
  if constexpr (DOT > 0) {
    double total[ND];
    if (grid_reduce<ND>(dots, partials, counter, total) && threadIdx.x == 0) kry_finalize(fin, st, total);
  }
}

// ---- fused "matrix-vector strategy" of assemble_first (fracstep.py:438-472) ------------------
// In:  A = C(uab) (just assembled), M, Kst.   Out, in ONE pass over the nonzeros:
//   b_first[row] = (M/dt - nu/2 K - 1/2 C) u1 + b0 (+ p_surf)        (:438-465)
//   A            =  M/dt + nu/2 K + 1/2 C, unit rows on Dirichlet dofs (:468-472)
//   dinv[row]    = 1 / A[row,row]                                      (Jacobi for the Krylov solve)
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_combine_first(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
                double* __restrict__ A, const double* __restrict__ M, const double* __restrict__ Kst,
                double inv_dt, double half_nu, const double* __restrict__ u1,
                const double* __restrict__ b0, const double* __restrict__ psurf,
                const uint8_t* __restrict__ is_bc_row, double* __restrict__ bfirst,
                double* __restrict__ dinv) {
  const int lane = threadIdx.x % LPR;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int ngroups = (gridDim.x * blockDim.x) / LPR;
  const int n_iter = (n_rows + ngroups - 1) / ngroups;
  for (int it = 0; it < n_iter; ++it) {
    const int row = it * ngroups + group;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    double diag = 0.0;
    if (row < n_rows) {
      const bool bc = is_bc_row[row];
      const int end = __ldg(rowptr + row + 1);
      for (int p = __ldg(rowptr + row) + lane; p < end; p += LPR) {
        const int c = __ldg(cols + p);
        const double m = inv_dt * __ldg(M + p);
        const double kk = half_nu * __ldg(Kst + p);
        const double cv = 0.5 * A[p];
        const double r = (m - cv) - kk;
        double a = (m + cv) + kk;
        const double* xc = u1 + (size_t)c * K;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fma(r, __ldg(xc + k), acc[k]);
        if (bc) a = (c == row) ? 1.0 : 0.0;
        A[p] = a;
        if (c == row) diag = a;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      diag += __shfl_xor_sync(0xffffffffu, diag, o);
    }
    if (row < n_rows && lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double v = acc[k] + b0[(size_t)row * K + k];
        if (psurf != nullptr) v += psurf[(size_t)row * K + k];
        bfirst[(size_t)row * K + k] = v;
      }
      dinv[row] = 1.0 / diag;
    }
  }
}

// ---- rectangular products on the V x Q and Q x V patterns ([nnz][K] values) --------------------
// out[row][k] = add[row][k] + scale * sum_p vals[p][k] * xq[cols[p]]   (P_i ps, G_i dp)
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_rect_vq(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
          const double* __restrict__ vals, const double* __restrict__ xq,
          const double* __restrict__ add, double scale, double* __restrict__ out) {
  const int lane = threadIdx.x % LPR;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  if (row < n_rows) {
    const int end = __ldg(rowptr + row + 1);
    for (int p = __ldg(rowptr + row) + lane; p < end; p += LPR) {
      const double xv = __ldg(xq + __ldg(cols + p));
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fma(__ldg(vals + (size_t)p * K + k), xv, acc[k]);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (row < n_rows && lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double a = add != nullptr ? add[(size_t)row * K + k] : 0.0;
      out[(size_t)row * K + k] = a + scale * acc[k];
    }
  }
}

// out[q] = scale * sum_p sum_k vals[p][k] * xv[cols[p]][k]            (sum_i D_i u_i)
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_rect_qv(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
          const double* __restrict__ vals, const double* __restrict__ xv, double scale,
          const uint8_t* __restrict__ zero_row, double* __restrict__ out) {
  const int lane = threadIdx.x % LPR;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  double acc = 0.0;
  if (row < n_rows) {
    const int end = __ldg(rowptr + row + 1);
    for (int p = __ldg(rowptr + row) + lane; p < end; p += LPR) {
      const double* xc = xv + (size_t)__ldg(cols + p) * K;
#pragma unroll
      for (int k = 0; k < K; ++k) acc = fma(__ldg(vals + (size_t)p * K + k), __ldg(xc + k), acc);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < n_rows && lane == 0) out[row] = (zero_row != nullptr && zero_row[row]) ? 0.0 : scale * acc;
}

// ---- small vector kernels ----------------------------------------------------------------
__global__ void k_lincomb2(int64_t n, double a, const double* __restrict__ x, double b,
                           const double* __restrict__ y, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * x[i] + b * y[i];
}

__global__ void k_fill(int64_t n, double v, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = v;
}

// vec[dofs[i]][comp] = values[i]   (set_bc, bcs.py:135-139)
__global__ void k_set_bc(int64_t n, const int* __restrict__ dofs, const double* __restrict__ values,
                         int K, int comp, double* __restrict__ vec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec[(size_t)dofs[i] * K + comp] = values[i];
}

__global__ void k_mark(int64_t n, const int* __restrict__ dofs, uint8_t* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[dofs[i]] = 1;
}

// strided component copies between the interleaved device layout and per-component host views
__global__ void k_extract(int64_t n, int K, int comp, const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i * K + comp];
}
__global__ void k_insert(int64_t n, int K, int comp, const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i * K + comp] = src[i];
}

// diagonal of a square CSR matrix -> dinv = 1/diag
__global__ void k_inv_diag(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
                           const double* __restrict__ vals, double* __restrict__ dinv) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  double d = 1.0;
  for (int p = rowptr[row]; p < rowptr[row + 1]; ++p)
    if (cols[p] == row) d = vals[p];
  dinv[row] = 1.0 / d;
}

// sums[k] = sum_i w_i * (a[i][k] - b[i][k])^2  (b may be null; w may be null => 1).  FIN_STORE:
// totals land in out[0..K).
template <int K>
__global__ void __launch_bounds__(256)
k_sqdiff(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* out,
         double* partials, unsigned* counter) {
  double s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double d = a[i * K + k] - (b != nullptr ? b[i * K + k] : 0.0);
      s[k] = fma(d, d, s[k]);
    }
  }
  double total[K];
  if (grid_reduce<K>(s, partials, counter, total) && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = total[k];
  }
}

// out[0] = sum_i x[i] ; out[1] = sum_i w[i]*x[i] (w may be null)
__global__ void __launch_bounds__(256)
k_sums(int64_t n, const double* __restrict__ x, const double* __restrict__ w, double* out,
       double* partials, unsigned* counter) {
  double s[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = x[i];
    s[0] += v;
    if (w != nullptr) s[1] = fma(w[i], v, s[1]);
  }
  double total[2];
  if (grid_reduce<2>(s, partials, counter, total) && threadIdx.x == 0) {
    out[0] = total[0];
    out[1] = total[1];
  }
}

// out[0] = sum_i a[i] * b[i]
__global__ void __launch_bounds__(256)
k_dot_all(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* out,
          double* partials, unsigned* counter) {
  double s[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s[0] = fma(a[i], b[i], s[0]);
  double total[1];
  if (grid_reduce<1>(s, partials, counter, total) && threadIdx.x == 0) out[0] = total[0];
}

// x[i] -= sums[which] * inv_norm   (null-space removal :573-574 with which=0, inv_norm = 1/n;
// mass-mean removal :579-591 with which=1, inv_norm = 1/vol); optionally y = z + x (ps = p + dp, :604)
__global__ void k_shift(int64_t n, double* __restrict__ x, const double* __restrict__ sums, int which,
                        double inv_norm, const double* __restrict__ z, double* __restrict__ y) {
  const double s = sums[which] * inv_norm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = x[i] - s;
    x[i] = v;
    if (y != nullptr) y[i] = z[i] + v;
  }
}

// ---- PCG (Jacobi) on K systems sharing one matrix ---------------------------------------------
// init: r = b - q (q = A x0, or r = b and x = 0 when no initial guess), p = dinv r;
//       sums rz, bb, rr
template <int K>
__global__ void __launch_bounds__(256)
k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q,
          const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
          double* __restrict__ p, KryState* st, double* partials, unsigned* counter) {
  double s[3 * K];
#pragma unroll
  for (int i = 0; i < 3 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double bv = b[i * K + k];
      double rv = bv;
      if (q != nullptr) rv -= q[i * K + k];
      else x[i * K + k] = 0.0;
      const double zv = di * rv;
      r[i * K + k] = rv;
      p[i * K + k] = zv;
      s[k] = fma(rv, zv, s[k]);
      s[K + k] = fma(bv, bv, s[K + k]);
      s[2 * K + k] = fma(rv, rv, s[2 * K + k]);
    }
  }
  double total[3 * K];
  if (grid_reduce<3 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_CG_INIT, st, total);
}

// x += alpha p ; r -= alpha q ; sums rz' = r.dinv r , rr = r.r
template <int K>
__global__ void __launch_bounds__(256)
k_cg_update(int64_t n, const double* __restrict__ p, const double* __restrict__ q,
            const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
            KryState* st, double* partials, unsigned* counter) {
  if (st->done) return;
  double alpha[K];
  int act[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    act[k] = st->active[k];
    alpha[k] = st->alpha[k];
  }
  double s[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!act[k]) continue;
      const int64_t j = i * K + k;
      x[j] = fma(alpha[k], p[j], x[j]);
      const double rv = fma(-alpha[k], q[j], r[j]);
      r[j] = rv;
      s[k] = fma(rv * di, rv, s[k]);
      s[K + k] = fma(rv, rv, s[K + k]);
    }
  }
  double total[2 * K];
  if (grid_reduce<2 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_CG_UPDATE, st, total);
}

// p = dinv r + beta p
template <int K>
__global__ void __launch_bounds__(256)
k_cg_p(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
       double* __restrict__ p, const KryState* st) {
  if (st->done) return;
  double beta[K];
  int act[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    act[k] = st->active[k];
    beta[k] = st->beta[k];
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!act[k]) continue;
      const int64_t j = i * K + k;
      p[j] = fma(beta[k], p[j], di * r[j]);
    }
  }
}

// ---- BiCGStab, right-preconditioned with Jacobi, on K systems sharing one matrix ---------------
// init: r = b - q (or r = b, x = 0); rhat = r; p = r; sums bb, rr
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q,
            double* __restrict__ x, double* __restrict__ r, double* __restrict__ rhat,
            double* __restrict__ p, KryState* st, double* partials, unsigned* counter) {
  double s[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int64_t j = i * K + k;
      const double bv = b[j];
      double rv = bv;
      if (q != nullptr) rv -= q[j];
      else x[j] = 0.0;
      r[j] = rv;
      rhat[j] = rv;
      p[j] = rv;
      s[k] = fma(bv, bv, s[k]);
      s[K + k] = fma(rv, rv, s[K + k]);
    }
  }
  double total[2 * K];
  if (grid_reduce<2 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_BCGS_INIT, st, total);
}

// s = r - alpha v   (in place in r)
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_s(int64_t n, const double* __restrict__ v, double* __restrict__ r, const KryState* st) {
  if (st->done) return;
  double alpha[K];
  int act[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    act[k] = st->active[k];
    alpha[k] = st->alpha[k];
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!act[k]) continue;
      const int64_t j = i * K + k;
      r[j] = fma(-alpha[k], v[j], r[j]);
    }
  }
}

// x += dinv (alpha p + omega s) ; r = s - omega t ; sums rr, rho' = rhat . r
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_update(int64_t n, const double* __restrict__ p, const double* __restrict__ t,
              const double* __restrict__ rhat, const double* __restrict__ dinv,
              double* __restrict__ x, double* __restrict__ r, KryState* st, double* partials,
              unsigned* counter) {
  if (st->done) return;
  double alpha[K], omega[K];
  int act[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    act[k] = st->active[k];
    alpha[k] = st->alpha[k];
    omega[k] = st->omega[k];
  }
  double s[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!act[k]) continue;
      const int64_t j = i * K + k;
      const double sv = r[j];
      x[j] = fma(di, fma(alpha[k], p[j], omega[k] * sv), x[j]);
      const double rv = fma(-omega[k], t[j], sv);
      r[j] = rv;
      s[k] = fma(rv, rv, s[k]);
      s[K + k] = fma(rhat[j], rv, s[K + k]);
    }
  }
  double total[2 * K];
  if (grid_reduce<2 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_BCGS_UPDATE, st, total);
}

// p = r + beta (p - omega v)
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_p(int64_t n, const double* __restrict__ r, const double* __restrict__ v,
         double* __restrict__ p, const KryState* st) {
  if (st->done) return;
  double beta[K], omega[K];
  int act[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    act[k] = st->active[k];
    beta[k] = st->beta[k];
    omega[k] = st->omega[k];
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (!act[k]) continue;
      const int64_t j = i * K + k;
      p[j] = fma(beta[k], fma(-omega[k], v[j], p[j]), r[j]);
    }
  }
}
