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

// ---- component-interleaved vectors ---------------------------------------------------------
// Velocity-space vectors hold K components per dof at stride KP = 4 for K = 3 (one aligned 32-byte
// sector per dof: the SpMM gather is ONE 256-bit load per nonzero instead of three 8-byte loads that
// each cost an L1 wavefront), KP = K otherwise.  The pad slot is kept at zero by every kernel.
template <int K>
struct Pad {
  static constexpr int KP = (K == 3) ? 4 : K;
};

template <int K>
__device__ __forceinline__ void ldk(const double* p, double (&v)[Pad<K>::KP]) {
  if constexpr (K == 3) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
  } else if constexpr (K == 2) {
    double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x;
    v[1] = t.y;
  } else {
    v[0] = *p;
  }
}
// read-only (non-coherent) path: for operands no thread of the kernel writes
template <int K>
__device__ __forceinline__ void ldk_nc(const double* p, double (&v)[Pad<K>::KP]) {
  if constexpr (K == 3) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
  } else if constexpr (K == 2) {
    double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v[0] = t.x;
    v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}
template <int K>
__device__ __forceinline__ void stk(double* p, const double (&v)[Pad<K>::KP]) {
  if constexpr (K == 3) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
  } else if constexpr (K == 2) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  } else {
    *p = v[0];
  }
}

// ---- CSR SpMM: y[row][k] = sum_p vals[p] * x[cols[p]][k] --------------------------------------
// LPR lanes cooperate on one row (rows of P2 matrices hold ~28 entries, P1 ~15); a warp therefore
// streams 32/LPR consecutive rows: the value/column loads of one warp instruction cover a
// contiguous stretch of the CSR arrays.  The nonzero loop is unrolled by two so that two
// column->gather chains are in flight per lane.  Persistent grid (grid-stride over row groups) so
// the number of partial sums for the fused dot products stays small.
//   DOT == 0: no reduction          DOT == 1: sums[k] = y_k . w_k
//   DOT == 2: sums[k] = y_k . w_k , sums[K+k] = y_k . y_k
template <int K, int LPR, int DOT>
__global__ void __launch_bounds__(256)
k_spmm(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
       const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
       const double* __restrict__ w, KryState* st, int fin, double* partials, unsigned* counter) {
  constexpr int KP = Pad<K>::KP;
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
    double acc[KP], acc2[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = acc2[k] = 0.0;
    if (row < n_rows) {
      const int end = __ldg(rowptr + row + 1);
      int p = __ldg(rowptr + row) + lane;
      for (; p + LPR < end; p += 2 * LPR) {
        const int c0 = __ldg(cols + p), c1 = __ldg(cols + p + LPR);
        const double v0 = __ldg(vals + p), v1 = __ldg(vals + p + LPR);
        double x0[KP], x1[KP];
        ldk_nc<K>(x + (size_t)c0 * KP, x0);
        ldk_nc<K>(x + (size_t)c1 * KP, x1);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          acc[k] = fma(v0, x0[k], acc[k]);
          acc2[k] = fma(v1, x1[k], acc2[k]);
        }
      }
      if (p < end) {
        const int c0 = __ldg(cols + p);
        const double v0 = __ldg(vals + p);
        double x0[KP];
        ldk_nc<K>(x + (size_t)c0 * KP, x0);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = fma(v0, x0[k], acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] += acc2[k];
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (row < n_rows && lane == 0) {
      if constexpr (KP > K) acc[KP - 1] = 0.0;
      stk<K>(y + (size_t)row * KP, acc);
      if constexpr (DOT >= 1) {
        double wv[KP];
        ldk<K>(w + (size_t)row * KP, wv);
#pragma unroll
        for (int k = 0; k < K; ++k) dots[k] = fma(acc[k], wv[k], dots[k]);
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
//   A            =  D^-1 (M/dt + nu/2 K + 1/2 C), unit rows on Dirichlet dofs (:468-472), stored
//                   ROW-SCALED by its own diagonal D when `scale` (left Jacobi preconditioning, the
//                   PETSc default side for BiCGStab [ext]): the Krylov kernels then need no
//                   preconditioner gather at all.  b2_get_matrix_values undoes the scaling.
//   dinv[row]    = 1 / D[row]  (1 when !scale)
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_combine_first(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
                double* __restrict__ A, const double* __restrict__ M, const double* __restrict__ Kst,
                double inv_dt, double half_nu, const double* __restrict__ u1,
                const double* __restrict__ b0, const double* __restrict__ psurf,
                const uint8_t* __restrict__ is_bc_row, int scale, double* __restrict__ bfirst,
                double* __restrict__ dinv) {
  constexpr int KP = Pad<K>::KP;
  const int lane = threadIdx.x % LPR;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int ngroups = (gridDim.x * blockDim.x) / LPR;
  const int n_iter = (n_rows + ngroups - 1) / ngroups;
  for (int it = 0; it < n_iter; ++it) {
    const int row = it * ngroups + group;
    double acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.0;
    double diag = 0.0;
    bool bc = false;
    int start = 0, end = 0;
    if (row < n_rows) {
      bc = is_bc_row[row];
      start = __ldg(rowptr + row);
      end = __ldg(rowptr + row + 1);
      // phase 1: the diagonal entry of the new left-hand side
      for (int p = start + lane; p < end; p += LPR)
        if (__ldg(cols + p) == row) diag = (inv_dt * __ldg(M + p) + 0.5 * A[p]) + half_nu * __ldg(Kst + p);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) diag += __shfl_xor_sync(0xffffffffu, diag, o);
    if (bc) diag = 1.0;
    const double invd = (scale && row < n_rows) ? 1.0 / diag : 1.0;
    if (row < n_rows) {
      for (int p = start + lane; p < end; p += LPR) {
        const int c = __ldg(cols + p);
        const double m = inv_dt * __ldg(M + p);
        const double kk = half_nu * __ldg(Kst + p);
        const double cv = 0.5 * A[p];
        const double r = (m - cv) - kk;
        double a = ((m + cv) + kk) * invd;
        double xc[KP];
        ldk_nc<K>(u1 + (size_t)c * KP, xc);
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = fma(r, xc[k], acc[k]);
        if (bc) a = (c == row) ? 1.0 : 0.0;
        A[p] = a;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (row < n_rows && lane == 0) {
      double bv[KP];
      ldk_nc<K>(b0 + (size_t)row * KP, bv);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += bv[k];
      if (psurf != nullptr) {
        ldk_nc<K>(psurf + (size_t)row * KP, bv);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += bv[k];
      }
      if constexpr (KP > K) acc[KP - 1] = 0.0;
      stk<K>(bfirst + (size_t)row * KP, acc);
      dinv[row] = invd;
    }
  }
}

// vals[p] *= s[row] (or /= when `divide`): used to hand the caller the unscaled A
__global__ void k_scale_rows(int n_rows, const int* __restrict__ rowptr, const double* __restrict__ s, int divide,
                             const double* __restrict__ in, double* __restrict__ out) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const double f = divide ? 1.0 / s[row] : s[row];
  for (int p = rowptr[row]; p < rowptr[row + 1]; ++p) out[p] = in[p] * f;
}

// ---- rectangular products on the V x Q and Q x V patterns ([nnz][K] values) --------------------
// out[row][k] = add[row][k] + scale * sum_p vals[p][k] * xq[cols[p]]   (P_i ps, G_i dp)
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_rect_vq(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
          const double* __restrict__ vals, const double* __restrict__ xq, const double* add,
          double scale, double* out) {
  constexpr int KP = Pad<K>::KP;
  const int lane = threadIdx.x % LPR;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  double acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc[k] = 0.0;
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
    double a[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) a[k] = 0.0;
    if (add != nullptr) ldk<K>(add + (size_t)row * KP, a);
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = fma(scale, acc[k], a[k]);
    stk<K>(out + (size_t)row * KP, a);
  }
}

// out[q] = scale * sum_p sum_k vals[p][k] * xv[cols[p]][k]            (sum_i D_i u_i)
template <int K, int LPR>
__global__ void __launch_bounds__(256)
k_rect_qv(int n_rows, const int* __restrict__ rowptr, const int* __restrict__ cols,
          const double* __restrict__ vals, const double* __restrict__ xv, double scale,
          const uint8_t* __restrict__ zero_row, double* __restrict__ out) {
  constexpr int KP = Pad<K>::KP;
  const int lane = threadIdx.x % LPR;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  double acc = 0.0;
  if (row < n_rows) {
    const int end = __ldg(rowptr + row + 1);
    for (int p = __ldg(rowptr + row) + lane; p < end; p += LPR) {
      double xc[KP];
      ldk_nc<K>(xv + (size_t)__ldg(cols + p) * KP, xc);
#pragma unroll
      for (int k = 0; k < K; ++k) acc = fma(__ldg(vals + (size_t)p * K + k), xc[k], acc);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < n_rows && lane == 0) out[row] = (zero_row != nullptr && zero_row[row]) ? 0.0 : scale * acc;
}

// ---- small vector kernels ----------------------------------------------------------------
__global__ void k_lincomb2(int64_t n, double a, const double* x, double b, const double* y, double* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * x[i] + b * y[i];
}

__global__ void k_fill(int64_t n, double v, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = v;
}

// vec[dofs[i]][comp] = values[i]   (set_bc, bcs.py:135-139); stride = KP
__global__ void k_set_bc(int64_t n, const int* __restrict__ dofs, const double* __restrict__ values,
                         int stride, int comp, double* __restrict__ vec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec[(size_t)dofs[i] * stride + comp] = values[i];
}

__global__ void k_mark(int64_t n, const int* __restrict__ dofs, uint8_t* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[dofs[i]] = 1;
}

// strided component copies between the interleaved device layout and per-component host views
__global__ void k_extract(int64_t n, int stride, int comp, const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i * stride + comp];
}
__global__ void k_insert(int64_t n, int stride, int comp, const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i * stride + comp] = src[i];
}
// [n][K] (the blocked layout of solver.u) <-> [n][KP]
__global__ void k_repack(int64_t n, int K, int sstride, int dstride, const double* __restrict__ src, double* __restrict__ dst) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * K; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / K;
    int k = (int)(t - i * K);
    dst[i * dstride + k] = src[i * sstride + k];
  }
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

// out[k] = sum_i (a[i][k] - b[i][k])^2  (b may be null)
template <int K>
__global__ void __launch_bounds__(256)
k_sqdiff(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* out,
         double* partials, unsigned* counter) {
  constexpr int KP = Pad<K>::KP;
  double s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double av[KP], bv[KP];
    ldk<K>(a + i * KP, av);
#pragma unroll
    for (int k = 0; k < KP; ++k) bv[k] = 0.0;
    if (b != nullptr) ldk<K>(b + i * KP, bv);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double d = av[k] - bv[k];
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
  constexpr int KP = Pad<K>::KP;
  double s[3 * K];
#pragma unroll
  for (int i = 0; i < 3 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
    double bv[KP], rv[KP], pv[KP];
    ldk<K>(b + i * KP, bv);
#pragma unroll
    for (int k = 0; k < KP; ++k) rv[k] = bv[k];
    if (q != nullptr) {
      double qv[KP];
      ldk<K>(q + i * KP, qv);
#pragma unroll
      for (int k = 0; k < KP; ++k) rv[k] -= qv[k];
    } else {
      double z[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) z[k] = 0.0;
      stk<K>(x + i * KP, z);
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) pv[k] = di * rv[k];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      s[k] = fma(rv[k], pv[k], s[k]);
      s[K + k] = fma(bv[k], bv[k], s[K + k]);
      s[2 * K + k] = fma(rv[k], rv[k], s[2 * K + k]);
    }
    stk<K>(r + i * KP, rv);
    stk<K>(p + i * KP, pv);
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
  constexpr int KP = Pad<K>::KP;
  if (st->done) return;
  double alpha[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) alpha[k] = (k < K && st->active[k]) ? st->alpha[k] : 0.0;
  double s[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
    double pv[KP], qv[KP], xv[KP], rv[KP];
    ldk_nc<K>(p + i * KP, pv);
    ldk_nc<K>(q + i * KP, qv);
    ldk<K>(x + i * KP, xv);
    ldk<K>(r + i * KP, rv);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      xv[k] = fma(alpha[k], pv[k], xv[k]);
      rv[k] = fma(-alpha[k], qv[k], rv[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      s[k] = fma(rv[k] * di, rv[k], s[k]);
      s[K + k] = fma(rv[k], rv[k], s[K + k]);
    }
    stk<K>(x + i * KP, xv);
    stk<K>(r + i * KP, rv);
  }
  double total[2 * K];
  if (grid_reduce<2 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_CG_UPDATE, st, total);
}

// p = dinv r + beta p   (components that have converged keep their p: alpha is then 0 anyway)
template <int K>
__global__ void __launch_bounds__(256)
k_cg_p(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
       double* __restrict__ p, const KryState* st) {
  constexpr int KP = Pad<K>::KP;
  if (st->done) return;
  double beta[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) beta[k] = k < K ? st->beta[k] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
    double rv[KP], pv[KP];
    ldk_nc<K>(r + i * KP, rv);
    ldk<K>(p + i * KP, pv);
#pragma unroll
    for (int k = 0; k < KP; ++k) pv[k] = fma(beta[k], pv[k], di * rv[k]);
    stk<K>(p + i * KP, pv);
  }
}

// ---- BiCGStab on K systems sharing one (already left-preconditioned, row-scaled) matrix ----------
// init: r = dinv b - q (or r = dinv b, x = 0); rhat = r; p = r; sums bb = |dinv b|^2, rr
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_init(int64_t n, const double* __restrict__ b, const double* __restrict__ q,
            const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
            double* __restrict__ rhat, double* __restrict__ p, KryState* st, double* partials,
            unsigned* counter) {
  constexpr int KP = Pad<K>::KP;
  double s[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i];
    double bv[KP], rv[KP];
    ldk<K>(b + i * KP, bv);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      bv[k] *= di;
      rv[k] = bv[k];
    }
    if (q != nullptr) {
      double qv[KP];
      ldk<K>(q + i * KP, qv);
#pragma unroll
      for (int k = 0; k < KP; ++k) rv[k] -= qv[k];
    } else {
      double z[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) z[k] = 0.0;
      stk<K>(x + i * KP, z);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      s[k] = fma(bv[k], bv[k], s[k]);
      s[K + k] = fma(rv[k], rv[k], s[K + k]);
    }
    stk<K>(r + i * KP, rv);
    stk<K>(rhat + i * KP, rv);
    stk<K>(p + i * KP, rv);
  }
  double total[2 * K];
  if (grid_reduce<2 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_BCGS_INIT, st, total);
}

// s = r - alpha v   (in place in r)
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_s(int64_t n, const double* __restrict__ v, double* __restrict__ r, const KryState* st) {
  constexpr int KP = Pad<K>::KP;
  if (st->done) return;
  double alpha[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) alpha[k] = (k < K && st->active[k]) ? st->alpha[k] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double vv[KP], rv[KP];
    ldk_nc<K>(v + i * KP, vv);
    ldk<K>(r + i * KP, rv);
#pragma unroll
    for (int k = 0; k < KP; ++k) rv[k] = fma(-alpha[k], vv[k], rv[k]);
    stk<K>(r + i * KP, rv);
  }
}

// x += alpha p + omega s ; r = s - omega t ; sums rr, rho' = rhat . r
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_update(int64_t n, const double* __restrict__ p, const double* __restrict__ t,
              const double* __restrict__ rhat, double* __restrict__ x, double* __restrict__ r,
              KryState* st, double* partials, unsigned* counter) {
  constexpr int KP = Pad<K>::KP;
  if (st->done) return;
  double alpha[KP], omega[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    const bool a = k < K && st->active[k];
    alpha[k] = a ? st->alpha[k] : 0.0;
    omega[k] = a ? st->omega[k] : 0.0;
  }
  double s[2 * K];
#pragma unroll
  for (int i = 0; i < 2 * K; ++i) s[i] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double pv[KP], tv[KP], hv[KP], xv[KP], rv[KP];
    ldk_nc<K>(p + i * KP, pv);
    ldk_nc<K>(t + i * KP, tv);
    ldk_nc<K>(rhat + i * KP, hv);
    ldk<K>(x + i * KP, xv);
    ldk<K>(r + i * KP, rv);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      xv[k] = fma(alpha[k], pv[k], fma(omega[k], rv[k], xv[k]));
      rv[k] = fma(-omega[k], tv[k], rv[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      s[k] = fma(rv[k], rv[k], s[k]);
      s[K + k] = fma(hv[k], rv[k], s[K + k]);
    }
    stk<K>(x + i * KP, xv);
    stk<K>(r + i * KP, rv);
  }
  double total[2 * K];
  if (grid_reduce<2 * K>(s, partials, counter, total) && threadIdx.x == 0) kry_finalize(FIN_BCGS_UPDATE, st, total);
}

// p = r + beta (p - omega v)   (frozen for converged components)
template <int K>
__global__ void __launch_bounds__(256)
k_bcgs_p(int64_t n, const double* __restrict__ r, const double* __restrict__ v,
         double* __restrict__ p, const KryState* st) {
  constexpr int KP = Pad<K>::KP;
  if (st->done) return;
  double beta[KP], omega[KP];
  bool act[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    act[k] = k < K && st->active[k];
    beta[k] = act[k] ? st->beta[k] : 0.0;
    omega[k] = act[k] ? st->omega[k] : 0.0;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double rv[KP], vv[KP], pv[KP];
    ldk_nc<K>(r + i * KP, rv);
    ldk_nc<K>(v + i * KP, vv);
    ldk<K>(p + i * KP, pv);
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (act[k]) pv[k] = fma(beta[k], fma(-omega[k], vv[k], pv[k]), rv[k]);
    stk<K>(p + i * KP, pv);
  }
}
