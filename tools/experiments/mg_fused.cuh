// NEGATIVE RESULTS of round 1, kept as reading material -- NOT compiled into libb200ipcs.so.
// Each variant was measured slower than the production kernel (DESIGN.md section 4 has the numbers); the code is
// preserved as it ran (it needs the helpers of oasisx_b200/csrc/{common,linalg,elem}.cuh and the dispatch code that
// round 1's b200ipcs.cu carried: see git history, commit ddf9723).
// ---- fused cycle kernels (tuning "mg_fused", OFF by default: a measured negative result) -------------------------
// Measured at 48^3 (tools/exp_mg_fused.py): same fields to 1e-13, same 9 iterations, but the pressure stage takes
// 1.27 ms instead of 1.01 ms -- re-evaluating the prolongation at each of the ~15 gathered columns costs more inside
// the kernel than the two saved launches (already cheap inside the CUDA graph) return.
// A V(1,1) cycle spends 5 launches per level (first sweep, residual, restriction, prolongation, post-sweep); on levels
// of 10^4..10^6 unknowns each is a few microseconds of work behind a launch.  Two pairs fuse without changing a bit of
// the result, because the first sweep from zero is local (x = omega D^-1 b) and a nested P1 prolongation has at most
// d+1 entries per row, so both can be RE-EVALUATED at the gathered columns instead of read from a finished vector:
//   k_mg_first_resid:    x = omega D^-1 b  and  r = b - A x          (replaces k_mg_first + k_mg_sweep<true>)
//   k_mg_prolong_sweep:  out = xp + omega D^-1 (b - A xp),  xp = x + P xc   (replaces the prolongation + k_mg_sweep<false>)
// Both need D^-1, b resp. P for every COLUMN of the operator, i.e. replicated levels or a single-rank fine level.
__global__ void __launch_bounds__(256)
k_mg_first_resid(int n_rows, const int* __restrict__ slice_ptr, const int* __restrict__ cols,
                 const double* __restrict__ vals, const double* __restrict__ dinv, const double* __restrict__ b,
                 double omega, double* __restrict__ x, double* __restrict__ r) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_slices = (n_rows + 31) >> 5;
  for (int s = warp; s < n_slices; s += nwarps) {
    const int base = __ldg(slice_ptr + s);
    const int len = (__ldg(slice_ptr + s + 1) - base) >> 5;
    const int row = (s << 5) + lane;
    double acc = 0.0;
#pragma unroll 4
    for (int t = 0; t < len; ++t) {
      const int c = ld_stream(cols + base + lane + (t << 5));
      acc = fma(ld_stream(vals + base + lane + (t << 5)), omega * __ldg(dinv + c) * __ldg(b + c), acc);
    }
    if (row < n_rows) {
      x[row] = omega * dinv[row] * b[row];
      r[row] = b[row] - acc;
    }
  }
}

__device__ __forceinline__ double mg_prolonged(const double* __restrict__ x, const int* __restrict__ Pptr,
                                               const int* __restrict__ Pcol, const double* __restrict__ Pval,
                                               const double* __restrict__ xc, int i) {
  double v = __ldg(x + i);
  for (int p = __ldg(Pptr + i); p < __ldg(Pptr + i + 1); ++p) v = fma(__ldg(Pval + p), __ldg(xc + __ldg(Pcol + p)), v);
  return v;
}

__global__ void __launch_bounds__(256)
k_mg_prolong_sweep(int n_rows, const int* __restrict__ slice_ptr, const int* __restrict__ cols,
                   const double* __restrict__ vals, const double* __restrict__ dinv, const double* __restrict__ b,
                   const double* __restrict__ x, const int* __restrict__ Pptr, const int* __restrict__ Pcol,
                   const double* __restrict__ Pval, const double* __restrict__ xc, double omega,
                   double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_slices = (n_rows + 31) >> 5;
  for (int s = warp; s < n_slices; s += nwarps) {
    const int base = __ldg(slice_ptr + s);
    const int len = (__ldg(slice_ptr + s + 1) - base) >> 5;
    const int row = (s << 5) + lane;
    double acc = 0.0;
    for (int t = 0; t < len; ++t) {
      const int c = ld_stream(cols + base + lane + (t << 5));
      const double a = ld_stream(vals + base + lane + (t << 5));
      if (a != 0.0) acc = fma(a, mg_prolonged(x, Pptr, Pcol, Pval, xc, c), acc);  // pads: (col = row, val = 0)
    }
    if (row < n_rows) {
      const double xp = mg_prolonged(x, Pptr, Pcol, Pval, xc, row);
      out[row] = fma(omega * dinv[row], b[row] - acc, xp);
    }
  }
}

