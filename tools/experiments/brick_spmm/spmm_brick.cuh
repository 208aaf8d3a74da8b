// NEGATIVE RESULT of round 2, kept as reading material -- NOT compiled into libb200ipcs.so (DESIGN.md section 4 has the
// numbers).  The brick form of the P2xP2 operator: kernels as they last ran (bitwise equal to k_spmm on the GPU, commit
// "brick kernels: byte-offset positions ..."), needing the helpers of oasisx_b200/csrc/{common,linalg}.cuh, the host
// builder in bricks.hpp next to this file and the launch / ABI code in host_glue.txt.
//   k_spmm_brick   2 blocks x 16 warps per SM, register-staged matrix stream            0.81 ms at 96^3 (k_spmm: 0.53)
//   k_spmm_brick2  1 block x 16 warps, per-warp rings fed by TMA bulk copies / cp.async  0.68 / 0.73 ms
// ---- brick SpMM: the same product with the gathered vector staged in shared memory (bricks.hpp) --------------------
// One block walks bricks b = blockIdx.x, blockIdx.x + gridDim.x, ...  Per brick: (1) the x values of the brick's gather
// list go to shared memory, component-major (coalesced: the list is sorted, i.e. runs of consecutive dofs); (2) every
// warp takes slices of the brick and streams values (8 B) and 16-bit list positions (2 B) from HBM, the operands of the
// FMAs come from shared memory.  Same slots, same order of the FMAs as k_spmm: y is bitwise the same.  Two resident
// blocks per SM (K * cap * 8 B of shared memory each) overlap one block's fill with the other's stream.
__device__ __forceinline__ int ld_stream(const unsigned short* p) {
  unsigned short v;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return (int)v;
}

template <int K, int DOT, bool RS, int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK, 2)
k_spmm_brick(int n_rows, const int* __restrict__ slice_ptr, const unsigned short* __restrict__ lcols,
             const double* __restrict__ vals, const int* __restrict__ order, const int* __restrict__ bptr,
             const int* __restrict__ gptr, const int* __restrict__ glist, int n_bricks, int cap,
             const double* __restrict__ x, int ld, double* __restrict__ y, const double* __restrict__ w, KryState* st,
             int fin, double* partials, unsigned* counter, RedCtl red_out, const double* __restrict__ rscale, int diag) {
  // diag (b2_set_tuning "spmm_brick_diag", timing only, results meaningless): 1 = no fill phase, 2 = no stream phase
  if (st != nullptr && st->done) return;
  extern __shared__ double sx[];  // [K][cap]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  constexpr int WPB = BLOCK / 32;
  constexpr int ND = DOT == 0 ? 1 : DOT * K;
  double dots[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) dots[i] = 0.0;
  for (int b = blockIdx.x; b < n_bricks; b += gridDim.x) {
    const int g0 = __ldg(gptr + b), ng = __ldg(gptr + b + 1) - g0;
    __syncthreads();  // the previous brick's operands are no longer read
    if (diag != 1) {  // fill: every list entry of this thread is requested before the first x value is (two memory
                      // latencies per group of FG entries, not two per entry)
      constexpr int F = (B2_BRICK_CAP + BLOCK - 1) / BLOCK, FG = 3;
      int col[F];
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const int i = threadIdx.x + f * BLOCK;
        col[f] = i < ng ? ld_stream(glist + g0 + i) : -1;
      }
#pragma unroll
      for (int f0 = 0; f0 < F; f0 += FG) {
        double xv[FG][K];
#pragma unroll
        for (int f = f0; f < f0 + FG && f < F; ++f)
#pragma unroll
          for (int k = 0; k < K; ++k) xv[f - f0][k] = col[f] >= 0 ? __ldg(x + (size_t)k * ld + col[f]) : 0.0;
#pragma unroll
        for (int f = f0; f < f0 + FG && f < F; ++f)
          if (col[f] >= 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) sx[k * cap + threadIdx.x + f * BLOCK] = xv[f - f0][k];
          }
      }
    }
    __syncthreads();
    const int j1 = diag != 2 ? __ldg(bptr + b + 1) : 0;
    for (int j = __ldg(bptr + b) + wib; j < j1; j += WPB) {
      const int s = order != nullptr ? __ldg(order + j) : j;
      const int base = __ldg(slice_ptr + s);
      const int len = (__ldg(slice_ptr + s + 1) - base) >> 5;
      const int row = (s << 5) + lane;
      const unsigned short* cp = lcols + base + lane;
      const double* vp = vals + base + lane;
      double acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = 0.0;
      // UNROLL entries per trip, the last trip predicated: all its loads are in flight together as well (a scalar
      // remainder loop would expose one memory latency per leftover entry at this kernel's 32 warps per SM)
      for (int t = 0; t < len; t += UNROLL) {
        int c[UNROLL];
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const bool ok = t + u < len;
          c[u] = ok ? ld_stream(cp + ((t + u) << 5)) : 0;
          v[u] = ok ? ld_stream(vp + ((t + u) << 5)) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          if (t + u < len) {
#pragma unroll
            for (int k = 0; k < K; ++k)
              acc[k] = fma(v[u], *reinterpret_cast<const double*>(reinterpret_cast<const char*>(sx + k * cap) + c[u]), acc[k]);
          }
      }
      if (row < n_rows) {
        double rs = 1.0, wv[DOT >= 1 ? K : 1];
        if constexpr (RS) rs = __ldg(rscale + row);
        if constexpr (DOT >= 1) {
#pragma unroll
          for (int k = 0; k < K; ++k) wv[k] = __ldg(w + (size_t)k * ld + row);
        }
        if constexpr (RS) {
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] *= rs;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) y[(size_t)k * ld + row] = acc[k];
        if constexpr (DOT >= 1) {
#pragma unroll
          for (int k = 0; k < K; ++k) dots[k] = fma(acc[k], wv[k], dots[k]);
        }
        if constexpr (DOT == 2) {
#pragma unroll
          for (int k = 0; k < K; ++k) dots[K + k] = fma(acc[k], acc[k], dots[K + k]);
        }
      }
    }
  }
  if constexpr (DOT > 0) reduce_finish<ND>(dots, partials, counter, fin, st, red_out);
}

// ---- brick SpMM, pipelined (k_spmm_brick2) ---------------------------------------------------------------------------
// The first form above is latency-bound: 104 KB of shared memory for x leaves 2 x 16 warps per SM, whose register-staged
// matrix loads drain at every brick barrier.  Here ONE block of 16 warps per SM owns all of the shared memory:
//   * the matrix stream (8-byte values + 16-bit list positions of CH steps of a slice, contiguous in SELL storage) is
//     moved by the TMA unit: lane 0 of every warp issues 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) into
//     the warp's own ring of ST stages.  No registers, no LSU load instructions, and -- because the warp's sequence of
//     chunks is fixed by the host (bricks.hpp: assign_warps) -- the prefetch runs ahead across slices AND across the
//     brick barriers: the HBM pipe never drains;
//   * the fill of a brick's x values is one batch of 8-byte cp.async gathers per thread from list entries that were
//     loaded into registers during the previous brick: one memory latency per brick, during which the rings fill;
//   * the warps of a block get slices of (nearly) equal total length per brick (longest first), so they reach the
//     barrier together.
// Same slots, same order of FMAs per row as k_spmm: y is bitwise the same.
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
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

constexpr size_t brick2_smem_bytes(int K, int cap, int warps, int ch, int st) {
  return sizeof(double) * (size_t)K * cap + (size_t)warps * st * ch * 32 * (sizeof(double) + sizeof(unsigned short)) +
         sizeof(unsigned long long) * (size_t)warps * st;
}

template <int K, int DOT, bool RS, int BLOCK, int CH, int ST, bool TMA>
__global__ void __launch_bounds__(BLOCK, 1)
k_spmm_brick2(int n_rows, const int* __restrict__ slice_ptr, const unsigned short* __restrict__ lcols,
              const double* __restrict__ vals, const int4* __restrict__ wdesc, const int* __restrict__ wseq,
              const int* __restrict__ gptr, const int* __restrict__ glist, int n_bricks, int cap,
              const double* __restrict__ x, int ld, double* __restrict__ y, const double* __restrict__ w, KryState* st,
              int fin, double* partials, unsigned* counter, RedCtl red_out, const double* __restrict__ rscale, int diag) {
  // diag (timing only, results meaningless): 1 = no fill, 2 = no stream
  if (st != nullptr && st->done) return;
  constexpr int WPB = BLOCK / 32;
  constexpr int ND = DOT == 0 ? 1 : DOT * K;
  constexpr int F = (B2_BRICK_CAP + BLOCK - 1) / BLOCK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sx = reinterpret_cast<double*>(smem_raw);  // [K][cap]
  double* svals_all = sx + (size_t)K * B2_BRICK_CAP;  // [WPB][ST][CH * 32]   (cap == B2_BRICK_CAP, checked by the host)
  unsigned short* scols_all = reinterpret_cast<unsigned short*>(svals_all + (size_t)WPB * ST * CH * 32);
  unsigned long long* bars_all = reinterpret_cast<unsigned long long*>(scols_all + (size_t)WPB * ST * CH * 32);
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  double* svals = svals_all + (size_t)wib * ST * CH * 32;
  unsigned short* scols = scols_all + (size_t)wib * ST * CH * 32;
  unsigned long long* bars = bars_all + wib * ST;
  if (TMA && lane == 0) {
#pragma unroll
    for (int s = 0; s < ST; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  // ---- this warp's work list (bricks.hpp: assign_warps): one 16-byte descriptor {brick, slice, first slot, steps} per
  // slice, in the order the warp meets them.  Producer and consumer each read one descriptor AHEAD of the slice they work
  // on, so that no slice starts with a chain of dependent metadata loads (with 16 warps per SM nothing would hide it).
  const int4* desc = wdesc + __ldg(wseq + (size_t)blockIdx.x * WPB + wib);
  const int n_desc = __ldg(wseq + (size_t)blockIdx.x * WPB + wib + 1) - __ldg(wseq + (size_t)blockIdx.x * WPB + wib);
  const int4 none = make_int4(-1, 0, 0, 0);
  // ---- producer cursor: the chunks in the order they will be consumed, running ST - 1 chunks ahead ---------------------
  int pi = 0, pt = 0, issued = 0;
  int4 pcur = (diag != 2 && n_desc > 0) ? __ldg(desc) : none;
  int4 pnxt = (diag != 2 && n_desc > 1) ? __ldg(desc + 1) : none;
  auto issue = [&]() {
    if (pcur.x >= 0) {
      const int tn = min(CH, pcur.w - pt);
      const int stg = issued % ST;
      if constexpr (TMA) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was read by this warp's generic loads
          mbar_expect_tx(bars + stg, (unsigned)(tn * 32 * 10));
          bulk_g2s(svals + (size_t)stg * CH * 32, vals + (size_t)pcur.z + ((size_t)pt << 5), (unsigned)(tn * 32 * 8), bars + stg);
          bulk_g2s(scols + (size_t)stg * CH * 32, lcols + (size_t)pcur.z + ((size_t)pt << 5), (unsigned)(tn * 32 * 2), bars + stg);
        }
      } else {  // 16-byte cp.async.cg per lane: CH * 256 B of values = CH / 2 copies per lane, CH * 64 B of positions = 1
        const char* gv = reinterpret_cast<const char*>(vals + (size_t)pcur.z + ((size_t)pt << 5));
        char* sv = reinterpret_cast<char*>(svals + (size_t)stg * CH * 32);
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) {
          const int off = (lane + 32 * i) * 16;
          if (off < tn * 256) cp_async_16(sv + off, gv + off);
        }
        const int offc = lane * 16;
        if (offc < tn * 64)
          cp_async_16(reinterpret_cast<char*>(scols + (size_t)stg * CH * 32) + offc,
                      reinterpret_cast<const char*>(lcols + (size_t)pcur.z + ((size_t)pt << 5)) + offc);
      }
      pt += tn;
      ++issued;
      if (pt >= pcur.w) {
        ++pi;
        pt = 0;
        pcur = pnxt;
        pnxt = (pi + 1 < n_desc) ? __ldg(desc + pi + 1) : none;
      }
    }
    if constexpr (!TMA) asm volatile("cp.async.commit_group;" ::: "memory");  // one group per call, empty or not
  };
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) issue();
  int consumed = 0;
  double dots[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) dots[i] = 0.0;
  // gather-list entries of the first brick (later ones are loaded during the brick before)
  int col[F];
  int ng = 0;
  if (blockIdx.x < n_bricks) {
    const int g0 = __ldg(gptr + blockIdx.x);
    ng = __ldg(gptr + blockIdx.x + 1) - g0;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const int i = threadIdx.x + f * BLOCK;
      col[f] = i < ng ? __ldg(glist + g0 + i) : -1;
    }
  }
  int ci = 0;
  int4 ccur = (diag != 2 && n_desc > 0) ? __ldg(desc) : none;
  int4 cnxt = (diag != 2 && n_desc > 1) ? __ldg(desc + 1) : none;
  for (int b = blockIdx.x; b < n_bricks; b += gridDim.x) {
    __syncthreads();  // the previous brick's operands are no longer read
    if (diag != 1) {
#pragma unroll
      for (int f = 0; f < F; ++f)
        if (col[f] >= 0) {
#pragma unroll
          for (int k = 0; k < K; ++k) cp_async_8(sx + (size_t)k * B2_BRICK_CAP + threadIdx.x + f * BLOCK, x + (size_t)k * ld + col[f]);
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    {  // the next brick's list entries: in flight during this brick's stream
      const int nb = b + gridDim.x;
      if (nb < n_bricks) {
        const int g0 = __ldg(gptr + nb);
        ng = __ldg(gptr + nb + 1) - g0;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const int i = threadIdx.x + f * BLOCK;
          col[f] = i < ng ? __ldg(glist + g0 + i) : -1;
        }
      }
    }
    if (diag == 2) continue;
    while (ccur.x == b) {
      const int s = ccur.y;
      const int len = ccur.w;
      ++ci;
      ccur = cnxt;
      cnxt = (ci + 1 < n_desc) ? __ldg(desc + ci + 1) : none;
      const int row = (s << 5) + lane;
      // the epilogue's operands are requested now: their latency passes while the slice is processed
      double rs = 1.0, wv[DOT >= 1 ? K : 1];
      if (row < n_rows) {
        if constexpr (RS) rs = __ldg(rscale + row);
        if constexpr (DOT >= 1) {
#pragma unroll
          for (int k = 0; k < K; ++k) wv[k] = __ldg(w + (size_t)k * ld + row);
        }
      }
      double acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = 0.0;
      for (int t = 0; t < len; t += CH) {
        issue();  // refill the stage consumed in the previous round
        const int tn = min(CH, len - t);
        const int stg = consumed % ST;
        if constexpr (TMA) {
          mbar_wait(bars + stg, (unsigned)((consumed / ST) & 1));
        } else {  // all but the ST - 1 newest groups have landed: this chunk's is among them
          asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory");
          __syncwarp();
        }
        const double* sv = svals + (size_t)stg * CH * 32 + lane;
        const unsigned short* sc = scols + (size_t)stg * CH * 32 + lane;
        // positions are stored as byte offsets and the capacity is a compile-time constant: per entry 2 + K shared
        // loads, one add and K FMAs (the first form of this loop spent 37 instructions per step on predicates and
        // index arithmetic, and 16 warps per SM could not issue them fast enough)
        const char* sxb = reinterpret_cast<const char*>(sx);
        if (tn == CH) {
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            const unsigned c = sc[u << 5];
            const double v = sv[u << 5];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fma(v, *reinterpret_cast<const double*>(sxb + c + k * (B2_BRICK_CAP * 8)), acc[k]);
          }
        } else {
          for (int u = 0; u < tn; ++u) {
            const unsigned c = sc[u << 5];
            const double v = sv[u << 5];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fma(v, *reinterpret_cast<const double*>(sxb + c + k * (B2_BRICK_CAP * 8)), acc[k]);
          }
        }
        __syncwarp();  // every lane is done with this stage before it is overwritten
        ++consumed;
      }
      if (row < n_rows) {
        if constexpr (RS) {
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] *= rs;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) y[(size_t)k * ld + row] = acc[k];
        if constexpr (DOT >= 1) {
#pragma unroll
          for (int k = 0; k < K; ++k) dots[k] = fma(acc[k], wv[k], dots[k]);
        }
        if constexpr (DOT == 2) {
#pragma unroll
          for (int k = 0; k < K; ++k) dots[K + k] = fma(acc[k], acc[k], dots[K + k]);
        }
      }
    }
  }
  if constexpr (DOT > 0) reduce_finish<ND>(dots, partials, counter, fin, st, red_out);
}

