// NEGATIVE RESULTS of round 1, kept as reading material -- NOT compiled into libb200ipcs.so.
// Each variant was measured slower than the production kernel (DESIGN.md section 4 has the numbers); the code is
// preserved as it ran (it needs the helpers of oasisx_b200/csrc/{common,linalg,elem}.cuh and the dispatch code that
// round 1's b200ipcs.cu carried: see git history, commit ddf9723).
// k_spmm_sm (SM-local work queues), k_spmm_tma (TMA-fed matrix stream), run-compressed columns (k_sell_cbase)
// ---- SM-local scheduling variant of the SpMM -----------------------------------------------------
// The slice schedule is cut into one contiguous range per SM (balanced by slot count, `sm_range`); every
// warp resident on SM s pulls the next slice of range s with an atomic counter, so all ~64 warps of an
// SM walk through neighbouring slices together and share the gathered vector lines in that SM's L1.
// Ranges of SMs that received no block (or are slower) are stolen once a warp's own range is drained,
// which also makes the result independent of block placement.  `next` must be zero at launch.
__device__ __forceinline__ unsigned get_smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

template <int K, int DOT, int UNROLL, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2048 / BLOCK)
k_spmm_sm(int n_rows, const int* __restrict__ slice_ptr, const int* __restrict__ cols,
          const double* __restrict__ vals, const int* __restrict__ order, const int* __restrict__ sm_dense,
          int n_ranges, const int* __restrict__ sm_range, int* __restrict__ next, const double* __restrict__ x, int ld,
          double* __restrict__ y, const double* __restrict__ w, KryState* st, int fin, double* partials,
          unsigned* counter, double* red_out) {
  if (st != nullptr && st->done) return;
  const int lane = threadIdx.x & 31;
  constexpr int ND = DOT == 0 ? 1 : DOT * K;
  __shared__ double sdots[DOT == 0 ? 1 : ND][DOT == 0 ? 1 : BLOCK];
  if constexpr (DOT > 0) {
#pragma unroll
    for (int i = 0; i < ND; ++i) sdots[i][threadIdx.x] = 0.0;
  }
  const int home = __ldg(sm_dense + get_smid()) % n_ranges;
  for (int hop = 0; hop < n_ranges; ++hop) {
    const int rg = (home + hop) % n_ranges;
    const int lo = __ldg(sm_range + rg), hi = __ldg(sm_range + rg + 1);
    if (hop > 0 && *((volatile int*)next + rg) >= hi - lo) continue;  // cheap peek before stealing
    for (;;) {
      int i = 0;
      if (lane == 0) i = lo + atomicAdd(next + rg, 1);
      i = __shfl_sync(0xffffffffu, i, 0);
      if (i >= hi) break;
      const int s = order != nullptr ? __ldg(order + i) : i;
      const int base = __ldg(slice_ptr + s);
      const int len = (__ldg(slice_ptr + s + 1) - base) >> 5;
      const int row = (s << 5) + lane;
      const int* cp = cols + base + lane;
      const double* vp = vals + base + lane;
      double acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = 0.0;
      int t = 0;
      for (; t + UNROLL <= len; t += UNROLL) {
        int c[UNROLL];
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          c[u] = ld_stream(cp + ((t + u) << 5));
          v[u] = ld_stream(vp + ((t + u) << 5));
        }
        double xv[UNROLL][K];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int k = 0; k < K; ++k) xv[u][k] = __ldg(x + (size_t)k * ld + c[u]);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fma(v[u], xv[u][k], acc[k]);
      }
      for (; t < len; ++t) {
        const int c = ld_stream(cp + (t << 5));
        const double v = ld_stream(vp + (t << 5));
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fma(v, __ldg(x + (size_t)k * ld + c), acc[k]);
      }
      if (row < n_rows) {
#pragma unroll
        for (int k = 0; k < K; ++k) y[(size_t)k * ld + row] = acc[k];
        if constexpr (DOT >= 1) {
#pragma unroll
          for (int k = 0; k < K; ++k) sdots[k][threadIdx.x] = fma(acc[k], w[(size_t)k * ld + row], sdots[k][threadIdx.x]);
        }
        if constexpr (DOT == 2) {
#pragma unroll
          for (int k = 0; k < K; ++k) sdots[K + k][threadIdx.x] = fma(acc[k], acc[k], sdots[K + k][threadIdx.x]);
        }
      }
    }
  }
  if constexpr (DOT > 0) {
    double dots[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) dots[i] = sdots[i][threadIdx.x];
    reduce_finish<ND>(dots, partials, counter, fin, st, red_out);
  }
}

__global__ void k_probe_smid(int* seen) { if (threadIdx.x == 0) seen[get_smid()] = 1; }

// ---- TMA-fed variant of the SpMM ---------------------------------------------------------------
// The matrix stream (values + columns of one slice, contiguous in SELL storage) is moved by the TMA
// unit: one elected lane per warp issues 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) of
// the next CH steps of its slice into a per-warp ring in shared memory while the warp gathers and
// multiplies the current chunk.  The stream then costs no registers and no LSU load instructions, so
// all of the thread's load slots go to the gathers (CH*K of them in flight).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

template <int K, int DOT, int CH, int STAGES, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_spmm_tma(int n_rows, const int* __restrict__ slice_ptr, const int* __restrict__ cols,
           const double* __restrict__ vals, const int* __restrict__ order, const double* __restrict__ x, int ld,
           double* __restrict__ y, const double* __restrict__ w, KryState* st, int fin, double* partials,
           unsigned* counter, double* red_out) {
  constexpr int WPB = BLOCK / 32;
  constexpr int ND = DOT == 0 ? 1 : DOT * K;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [WPB][STAGES] value chunks (CH*32 doubles), then column chunks (CH*32 ints), then barriers
  double* svals_all = reinterpret_cast<double*>(smem_raw);
  int* scols_all = reinterpret_cast<int*>(smem_raw + (size_t)WPB * STAGES * CH * 32 * sizeof(double));
  unsigned long long* bars_all =
      reinterpret_cast<unsigned long long*>(smem_raw + (size_t)WPB * STAGES * CH * 32 * (sizeof(double) + sizeof(int)));
  double* sdots = reinterpret_cast<double*>(bars_all + WPB * STAGES);  // [ND][BLOCK] when DOT > 0
  if (st != nullptr && st->done) return;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  double* svals = svals_all + (size_t)wib * STAGES * CH * 32;
  int* scols = scols_all + (size_t)wib * STAGES * CH * 32;
  unsigned long long* bars = bars_all + wib * STAGES;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (DOT > 0) {
#pragma unroll
    for (int i = 0; i < ND; ++i) sdots[i * BLOCK + threadIdx.x] = 0.0;
  }
  __syncwarp();
  const int n_slices = (n_rows + 31) >> 5;
  const int stride = gridDim.x * WPB;
  // producer cursor (slice list index pi, step pt) runs STAGES-1 chunks ahead of the consumer
  int pi = blockIdx.x * WPB + wib, pt = 0, pbase = 0, plen = 0;
  auto load_slice = [&](int i, int& base, int& len) {
    if (i < n_slices) {
      const int s = order != nullptr ? __ldg(order + i) : i;
      base = __ldg(slice_ptr + s);
      len = (__ldg(slice_ptr + s + 1) - base) >> 5;
    } else {
      base = 0;
      len = 0;
    }
  };
  load_slice(pi, pbase, plen);
  int issued = 0;
  auto issue = [&]() {  // enqueue the next chunk of this warp's stream, if any
    while (pi < n_slices && pt >= plen) {
      pi += stride;
      pt = 0;
      load_slice(pi, pbase, plen);
    }
    if (pi >= n_slices) return;
    const int tn = min(CH, plen - pt);
    const int stg = issued % STAGES;
    if (lane == 0) {
      mbar_expect_tx(bars + stg, (unsigned)(tn * 32 * 12));
      bulk_g2s(svals + (size_t)stg * CH * 32, vals + (size_t)pbase + ((size_t)pt << 5), (unsigned)(tn * 32 * 8), bars + stg);
      bulk_g2s(scols + (size_t)stg * CH * 32, cols + (size_t)pbase + ((size_t)pt << 5), (unsigned)(tn * 32 * 4), bars + stg);
    }
    pt += tn;
    ++issued;
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue();
  int consumed = 0;
  for (int i = blockIdx.x * WPB + wib; i < n_slices; i += stride) {
    int base, len;
    load_slice(i, base, len);
    const int s = order != nullptr ? __ldg(order + i) : i;
    const int row = (s << 5) + lane;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int t = 0; t < len; t += CH) {
      issue();  // refill the stage consumed in the previous round
      const int tn = min(CH, len - t);
      const int stg = consumed % STAGES;
      mbar_wait(bars + stg, (unsigned)((consumed / STAGES) & 1));
      const double* sv = svals + (size_t)stg * CH * 32 + lane;
      const int* sc = scols + (size_t)stg * CH * 32 + lane;
      if (tn == CH) {
        double xv[CH][K];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int c = sc[u << 5];
#pragma unroll
          for (int k = 0; k < K; ++k) xv[u][k] = __ldg(x + (size_t)k * ld + c);
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const double v = sv[u << 5];
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fma(v, xv[u][k], acc[k]);
        }
      } else {
        for (int u = 0; u < tn; ++u) {
          const int c = sc[u << 5];
          const double v = sv[u << 5];
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fma(v, __ldg(x + (size_t)k * ld + c), acc[k]);
        }
      }
      __syncwarp();  // every lane is done reading this stage before lane 0 lets the TMA overwrite it
      ++consumed;
    }
    if (row < n_rows) {
#pragma unroll
      for (int k = 0; k < K; ++k) y[(size_t)k * ld + row] = acc[k];
      if constexpr (DOT >= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k)
          sdots[k * BLOCK + threadIdx.x] = fma(acc[k], w[(size_t)k * ld + row], sdots[k * BLOCK + threadIdx.x]);
      }
      if constexpr (DOT == 2) {
#pragma unroll
        for (int k = 0; k < K; ++k) sdots[(K + k) * BLOCK + threadIdx.x] = fma(acc[k], acc[k], sdots[(K + k) * BLOCK + threadIdx.x]);
      }
    }
  }
  if constexpr (DOT > 0) {
    double dots[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) dots[i] = sdots[i * BLOCK + threadIdx.x];
    reduce_finish<ND>(dots, partials, counter, fin, st, red_out);
  }
}

// run detection for the compressed SpMM (k_spmm<..., COMP>): one warp per slice column
__global__ void k_sell_cbase(int n_rows, int64_t n_slice_cols, const int* __restrict__ slice_ptr,
                             const int* __restrict__ scols, int* __restrict__ cbase) {
  const int lane = threadIdx.x & 31;
  const int64_t sc = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (sc >= n_slice_cols) return;
  // slice of this slice column: last s with slice_ptr[s] <= 32 * sc
  const int n_slices = (n_rows + 31) >> 5;
  int lo = 0, hi = n_slices - 1;
  const int64_t slot0 = sc << 5;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int64_t)slice_ptr[mid] <= slot0) lo = mid;
    else hi = mid - 1;
  }
  const int c = scols[slot0 + lane];
  const int c0 = __shfl_sync(0xffffffffu, c, 0);
  const bool full = ((lo << 5) + 31) < n_rows;  // rows beyond n_rows have no columns: never a run
  const bool ok = __all_sync(0xffffffffu, c == c0 + lane) && full;
  if (lane == 0) cbase[sc] = ok ? c0 : -1;
}

