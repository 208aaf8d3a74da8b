// ROUND-1 PIPELINE of assemble_first, superseded in round 2 by k_first_cells (oasisx_b200/csrc/elem.cuh) -- NOT compiled.
// zero-fill A -> k_convection (one thread per cell, 100 global FP64 reductions per P2 tetrahedron) -> k_combine_first
// (one pass over A, M, K: b_first, row-scaled A, Dirichlet rows): 14.2 GB of DRAM traffic and 3.3 ms at 96^3
// (profiles/r01_ncu_assemble_first_96cube.txt).
// C[i,j] += |detJ| sum_{a,dl} w[a][dl] T[a][dl][i][j],  w[a][dl] = sum_k Kinv[dl][k] uab_k[dof_a]
// One thread per cell; the element matrix is produced row by row (NV accumulators live in
// registers) and scattered with FP64 reductions (RED.ADD.F64) onto the SELL slots of the row.
template <int D, int DEG>
__global__ void __launch_bounds__(128)
k_convection(int64_t n_cells, const double* __restrict__ x, const int* __restrict__ cell_nodes,
             const int* __restrict__ vdofs, int n_rows_owned, const double* __restrict__ uab, int ld,
             const int* __restrict__ rowptr, const int* __restrict__ cols,
             const int* __restrict__ slice_ptr, const uint8_t* __restrict__ pos8,
             double* __restrict__ Avals) {
  using E = El<D, DEG>;
  constexpr int NV = E::NV;
  constexpr int NVP = (NV + 3) / 4 * 4;
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  Geo<D> g = cell_geometry<D>(x, cell_nodes + c * (D + 1));
  int dofs[NV];
#pragma unroll
  for (int a = 0; a < NV; ++a) dofs[a] = vdofs[c * NV + a];
  double w[NV][D];
#pragma unroll
  for (int a = 0; a < NV; ++a) {
    double u[D];
#pragma unroll
    for (int k = 0; k < D; ++k) u[k] = __ldg(uab + (size_t)k * ld + dofs[a]);
#pragma unroll
    for (int dl = 0; dl < D; ++dl) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < D; ++k) s += g.Kinv[dl][k] * u[k];
      w[a][dl] = s * g.detJ;
    }
  }
  // the loop over the rows is unrolled so that every reference-tensor entry is a compile-time operand of its DFMA
  // (it was 300 indexed constant loads per row before) and its zero entries (40 % for P2 tetrahedra) cost nothing
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int row = dofs[i];
    if (row >= n_rows_owned) continue;
    double r[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) r[j] = 0.0;
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int dl = 0; dl < D; ++dl) {
        const double wv = w[a][dl];
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (E::T(a, dl, i, j) != 0.0) r[j] = fma(wv, E::T(a, dl, i, j), r[j]);
      }
    double* rowbase = Avals + (size_t)__ldg(slice_ptr + (row >> 5)) + (row & 31);
    if (pos8 != nullptr) {
      const uint32_t* pw = reinterpret_cast<const uint32_t*>(pos8 + ((size_t)c * NV + i) * NVP);
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if ((j & 3) == 0) word = __ldg(pw + (j >> 2));
        const int t = (word >> (8 * (j & 3))) & 0xff;
        atomicAdd(rowbase + ((size_t)t << 5), r[j]);
      }
    } else {
      int lo = rowptr[row], hi = rowptr[row + 1];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        int pos = csr_find(cols, lo, hi, dofs[j]);
        atomicAdd(rowbase + ((size_t)(pos - lo) << 5), r[j]);
      }
    }
  }
}

// ---- fused "matrix-vector strategy" of assemble_first (fracstep.py:438-472) ------------------
// In:  A = C(uab) (just assembled), M, Kst, all in SELL slots.   Out, in ONE pass over the slots:
//   b_first[row] = (M/dt - nu/2 K - 1/2 C) u1 + b0 (+ p_surf)        (:438-465)
//   A            =  D^-1 (M/dt + nu/2 K + 1/2 C), unit rows on Dirichlet dofs (:468-472), stored
//                   ROW-SCALED by its own diagonal D when `scale` (left Jacobi preconditioning, the
//                   PETSc default side for BiCGStab [ext]): the Krylov kernels then need no
//                   preconditioner at all.  b2_get_matrix_values undoes the scaling.
//   dinv[row]    = 1 / D[row]  (1 when !scale)
template <int K, int U = 4, int MINB = 4>
__global__ void __launch_bounds__(256, MINB)
k_combine_first(int n_rows, const int* __restrict__ slice_ptr, const int* __restrict__ cols,
                const int* __restrict__ diag_t, double* __restrict__ A, const double* __restrict__ M,
                const double* __restrict__ Kst, const int* __restrict__ order, double inv_dt,
                double half_nu, const double* __restrict__ u1, int ld, const double* __restrict__ b0,
                const double* __restrict__ psurf, const uint8_t* __restrict__ is_bc_row, int scale,
                double* __restrict__ bfirst, double* __restrict__ dinv) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int n_slices = (n_rows + 31) >> 5;
  for (int i = blockIdx.x * wpb + wib; i < n_slices; i += gridDim.x * wpb) {
    const int s = order != nullptr ? __ldg(order + i) : i;
    const int base = __ldg(slice_ptr + s);
    const int len = (__ldg(slice_ptr + s + 1) - base) >> 5;
    const int row = (s << 5) + lane;
    const bool live = row < n_rows;
    const bool bc = live && is_bc_row[row];
    double invd = 1.0;
    if (live && scale && !bc) {
      const size_t pd = (size_t)base + ((size_t)__ldg(diag_t + row) << 5) + lane;
      invd = 1.0 / ((inv_dt * __ldg(M + pd) + 0.5 * A[pd]) + half_nu * __ldg(Kst + pd));
    }
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    const int dt_row = live ? __ldg(diag_t + row) : -1;
    auto body = [&](int t, int c, double mv, double kv, double av, const double (&xu)[K]) {
      const double m = inv_dt * mv;
      const double kk = half_nu * kv;
      const double cv = 0.5 * av;
      const double r = (m - cv) - kk;
      double a = ((m + cv) + kk) * invd;
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fma(r, xu[k], acc[k]);
      if (bc) a = (t == dt_row) ? 1.0 : 0.0;
      A[(size_t)base + ((size_t)t << 5) + lane] = a;
    };
    int t = 0;
    for (; t + U <= len; t += U) {  // U independent (stream -> gather) chains in flight
      int cc[U];
      double mv[U], kv[U], av[U], xu[U][K];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t p = (size_t)base + ((size_t)(t + u) << 5) + lane;
        cc[u] = ld_stream(cols + p);
        mv[u] = ld_stream(M + p);
        kv[u] = ld_stream(Kst + p);
        av[u] = A[p];
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) xu[u][k] = __ldg(u1 + (size_t)k * ld + cc[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) body(t + u, cc[u], mv[u], kv[u], av[u], xu[u]);
    }
    for (; t < len; ++t) {
      const size_t p = (size_t)base + ((size_t)t << 5) + lane;
      const int c = ld_stream(cols + p);
      double xu[K];
#pragma unroll
      for (int k = 0; k < K; ++k) xu[k] = __ldg(u1 + (size_t)k * ld + c);
      body(t, c, ld_stream(M + p), ld_stream(Kst + p), A[p], xu);
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double v = acc[k] + b0[(size_t)k * ld + row];
        if (psurf != nullptr) v += psurf[(size_t)k * ld + row];
        bfirst[(size_t)k * ld + row] = v;
      }
      dinv[row] = invd;
    }
  }
}

