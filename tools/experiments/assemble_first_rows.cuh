// NEGATIVE RESULTS of round 1, kept as reading material -- NOT compiled into libb200ipcs.so.
// Each variant was measured slower than the production kernel (DESIGN.md section 4 has the numbers); the code is
// preserved as it ran (it needs the helpers of oasisx_b200/csrc/{common,linalg,elem}.cuh and the dispatch code that
// round 1's b200ipcs.cu carried: see git history, commit ddf9723).
// row I of the element convection matrix with I a compile-time constant: every reference-tensor entry becomes a
// constant-bank operand of its DFMA (no LDC per multiply-add, zero entries fold away)
template <class E, int I>
__device__ __forceinline__ void conv_row_fixed(const double (&w)[E::NV][E::D], double (&r)[E::NV]) {
#pragma unroll
  for (int a = 0; a < E::NV; ++a)
#pragma unroll
    for (int dl = 0; dl < E::D; ++dl)
#pragma unroll
      for (int j = 0; j < E::NV; ++j)
        if (E::T(a, dl, I, j) != 0.0) r[j] = fma(w[a][dl], E::T(a, dl, I, j), r[j]);  // folds at compile time
}

template <class E, int I = 0>
__device__ __forceinline__ void conv_row(int i, const double (&w)[E::NV][E::D], double (&r)[E::NV]) {
  if constexpr (I < E::NV) {
    if (i == I) conv_row_fixed<E, I>(w, r);
    else conv_row<E, I + 1>(i, w, r);
  }
}

// ---- assemble_first, row-wise and fused (fracstep.py:432-472 in ONE kernel, no atomics) -----------------
// The scatter version above pays 2.6 FP64 reductions per nonzero of the P2 operator (530 M RED at 96^3, bound by the
// REDG issue rate of the SMs) into a zero-filled C, which a second kernel (k_combine_first) then reads back with M
// and K.  Here a warp owns a 32-row slice of the sliced-ELL operator instead: lane = row.  A lane walks the cells
// adjacent to its dof (`adj`, sorted by local index first so that the 32 lanes of a stencil-class slice use the SAME
// row i of the reference tensor: warp-uniform constant-bank operands), forms row i of each cell's convection matrix
// in registers and adds it into its private column of a shared-memory copy of the slice (bank = lane: conflict-free).
// The slice is then combined with M and K exactly as k_combine_first does and written ONCE:
//   A = D^-1 (M/dt + nu/2 K + 1/2 C), Dirichlet rows -> identity;  b_first = (M/dt - nu/2 K - 1/2 C) u1 + b0 (+ p_surf).
// Traffic: M, K, columns read once, A written once (28 B/slot instead of 60), no zero-fill; the arithmetic is the same
// 3.4 kflop per cell plus the per-row recomputation of the cell geometry.  Deterministic (fixed summation order).
__global__ void k_gen_adj_keys(int64_t n_cells, int nv, const int* __restrict__ vdofs, int n_rows_owned,
                               unsigned long long* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells * nv) return;
  const int64_t c = t / nv;
  const int i = (int)(t - c * nv);
  int row = vdofs[t];
  unsigned lo = ((unsigned)i << 28) | (unsigned)c;  // sort by (row, local index, cell); cells < 2^28 (checked by the host)
  if (row >= n_rows_owned) { row = n_rows_owned; lo = 0; }
  keys[t] = ((unsigned long long)(unsigned)row << 32) | lo;
}

template <int D, int DEG>
__global__ void __launch_bounds__(128)
k_assemble_first_rows(int n_rows, const int* __restrict__ slice_ptr, const int* __restrict__ cols,
                      const int* __restrict__ diag_t, const int* __restrict__ order, const int* __restrict__ adj_ptr,
                      const int* __restrict__ adj, const double* __restrict__ x, const int* __restrict__ cell_nodes,
                      const int* __restrict__ vdofs, const double* __restrict__ uab, int ld,
                      const uint8_t* __restrict__ pos8, const double* __restrict__ M, const double* __restrict__ Kst,
                      double* __restrict__ A, double inv_dt, double half_nu, const double* __restrict__ u1,
                      const double* __restrict__ b0, const double* __restrict__ psurf,
                      const uint8_t* __restrict__ is_bc_row, int scale, double* __restrict__ bfirst,
                      double* __restrict__ dinv, int maxlen) {
  using E = El<D, DEG>;
  constexpr int NV = E::NV, K = D;
  constexpr int NVP = (NV + 3) / 4 * 4;
  extern __shared__ double s_acc[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* my = s_acc + (size_t)wib * maxlen * 32 + lane;  // my[t * 32]: entry t of this lane's row
  const int n_slices = (n_rows + 31) >> 5;
  for (int it = blockIdx.x * wpb + wib; it < n_slices; it += gridDim.x * wpb) {
    const int s = order != nullptr ? __ldg(order + it) : it;
    const int base = __ldg(slice_ptr + s);
    const int len = (__ldg(slice_ptr + s + 1) - base) >> 5;
    const int row = (s << 5) + lane;
    const bool live = row < n_rows;
    for (int t = 0; t < len; ++t) my[t << 5] = 0.0;
    // ---- C(uab): rows of the element matrices of the adjacent cells ------------------------------------
    const int a0 = live ? __ldg(adj_ptr + row) : 0, a1 = live ? __ldg(adj_ptr + row + 1) : 0;
    for (int q = a0; q < a1; ++q) {
      const unsigned e = (unsigned)__ldg(adj + q);
      const int i = (int)(e >> 28);
      const int64_t c = (int64_t)(e & 0x0fffffffu);
      const Geo<D> g = cell_geometry<D>(x, cell_nodes + c * (D + 1));
      const int* dofs = vdofs + c * NV;
      double r[NV], w[NV][D];
#pragma unroll
      for (int j = 0; j < NV; ++j) r[j] = 0.0;
#pragma unroll
      for (int a = 0; a < NV; ++a) {
        const int da = __ldg(dofs + a);
        double u[D];
#pragma unroll
        for (int k = 0; k < D; ++k) u[k] = __ldg(uab + (size_t)k * ld + da);
#pragma unroll
        for (int dl = 0; dl < D; ++dl) {
          double wv = 0;
#pragma unroll
          for (int k = 0; k < D; ++k) wv += g.Kinv[dl][k] * u[k];
          w[a][dl] = wv * g.detJ;
        }
      }
      conv_row<E>(i, w, r);
      const uint32_t* pw = reinterpret_cast<const uint32_t*>(pos8 + ((size_t)c * NV + i) * NVP);
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if ((j & 3) == 0) word = __ldg(pw + (j >> 2));
        const int t = (word >> (8 * (j & 3))) & 0xff;
        my[t << 5] += r[j];
      }
    }
    // ---- combine with M and K, Dirichlet rows, Jacobi scaling, b_first (as k_combine_first) -------------
    const bool bc = live && is_bc_row[row];
    const int dt_row = live ? __ldg(diag_t + row) : -1;
    double invd = 1.0;
    if (live && scale && !bc) {
      const size_t pd = (size_t)base + ((size_t)dt_row << 5) + lane;
      invd = 1.0 / ((inv_dt * __ldg(M + pd) + 0.5 * my[dt_row << 5]) + half_nu * __ldg(Kst + pd));
    }
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    auto body = [&](int t, double mv, double kv, double av, const double (&xu)[K]) {
      const double m = inv_dt * mv;
      const double kk = half_nu * kv;
      const double cv = 0.5 * av;
      const double rr = (m - cv) - kk;
      double a = ((m + cv) + kk) * invd;
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fma(rr, xu[k], acc[k]);
      if (bc) a = (t == dt_row) ? 1.0 : 0.0;
      A[(size_t)base + ((size_t)t << 5) + lane] = a;
    };
    int t = 0;
    for (; t + 4 <= len; t += 4) {
      int cc[4];
      double mv[4], kv[4], xu[4][K];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t p = (size_t)base + ((size_t)(t + u) << 5) + lane;
        cc[u] = ld_stream(cols + p);
        mv[u] = ld_stream(M + p);
        kv[u] = ld_stream(Kst + p);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) xu[u][k] = __ldg(u1 + (size_t)k * ld + cc[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) body(t + u, mv[u], kv[u], my[(t + u) << 5], xu[u]);
    }
    for (; t < len; ++t) {
      const size_t p = (size_t)base + ((size_t)t << 5) + lane;
      const int cidx = ld_stream(cols + p);
      double xu[K];
#pragma unroll
      for (int k = 0; k < K; ++k) xu[k] = __ldg(u1 + (size_t)k * ld + cidx);
      body(t, ld_stream(M + p), ld_stream(Kst + p), my[t << 5], xu);
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

