// Brick form of a sliced-ELL operator (host-side builder; plain C++, no CUDA: tested on the CPU through
// b2_host_build_bricks).
//
// The SELL-32 SpMM (linalg.cuh: k_spmm) gathers x[col] through L1/L2 for every stored entry: 28.5 gathers per P2 row,
// 3 components each, and the ncu capture of round 2 shows the L2 -> L1 path, not HBM, as its bound.  A BRICK is a run
// of consecutive entries of the slice schedule (all stencil classes of one small spatial neighbourhood, <= max_slices
// 32-row slices) together with the sorted list of the DISTINCT columns its rows touch (the "gather list", <= cap
// entries).  The brick kernel (k_spmm_brick) loads the gather list's x values into shared memory once -- 4 to 5 loads
// per row instead of 28.5 -- and the matrix stream then carries 16-bit positions in that list instead of 32-bit global
// columns: 10 bytes per stored entry instead of 12.  The value array is the SELL-32 array itself (same slots), so the
// assembly kernels and every other consumer are untouched; the arithmetic per row is the same sequence of FMAs, hence
// bitwise the same y.
//
// Bricks never straddle a `hint` boundary (the host passes one hint group per spatial key); inside a group slices are
// packed greedily until the gather list would exceed `cap` or `max_slices` is reached.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace b2bricks {

struct Bricks {
  std::vector<int> brick_ptr;      // n_bricks + 1 offsets into the slice schedule `order`
  std::vector<int> gptr;           // n_bricks + 1 offsets into glist
  std::vector<int> glist;          // gather lists: sorted distinct columns of each brick
  std::vector<uint16_t> lcols;     // one per SELL slot: 8 * (position of the slot's column in its brick's gather list)
  // work lists of the pipelined kernel (k_spmm_brick2), see assign_warps
  std::vector<int> wdesc, wseq;
  int warps = 0, grid = 0;
  int max_gather = 0;              // longest gather list
  int error = 0;                   // 1: a single slice touches more than `cap` distinct columns
  int n_bricks() const { return (int)brick_ptr.size() - 1; }
};

// order: n_slices entries (nullptr = identity); hint_ptr: n_hints + 1 offsets into order (nullptr = one group)
inline void build(int n_rows, int n_cols, const int* slice_ptr, const int* scols, const int* order, const int* hint_ptr,
                  int n_hints, int cap, int max_slices, int n_threads, Bricks& out) {
  const int n_slices = (n_rows + 31) / 32;
  std::vector<int> one_group = {0, n_slices};
  if (hint_ptr == nullptr) {
    hint_ptr = one_group.data();
    n_hints = 1;
  }
  const int64_t slots = slice_ptr[n_slices];
  out = Bricks();
  out.lcols.assign((size_t)slots, 0);
  if (cap > 8191) cap = 8191;  // 8 * position must fit 16 bits
  n_threads = std::max(1, std::min(n_threads, n_hints));
  struct Part {
    std::vector<int> ends, gsize, glist;  // per brick: end offset in order, gather length; concatenated lists
    int error = 0;
  };
  std::vector<Part> parts((size_t)n_threads);
  auto work = [&](int tid) {
    Part& P = parts[(size_t)tid];
    const int g0 = (int)((int64_t)n_hints * tid / n_threads), g1 = (int)((int64_t)n_hints * (tid + 1) / n_threads);
    std::vector<int> stamp((size_t)n_cols, -1), pos((size_t)n_cols, 0), gather, fresh;
    int serial = 0;
    auto slice_of = [&](int j) { return order != nullptr ? order[j] : j; };
    auto close = [&](int j_begin, int j_end) {  // brick = schedule entries [j_begin, j_end), columns in `gather`
      std::sort(gather.begin(), gather.end());
      for (size_t i = 0; i < gather.size(); ++i) pos[(size_t)gather[i]] = (int)i;
      for (int j = j_begin; j < j_end; ++j) {
        const int s = slice_of(j);
        // stored as BYTE offsets into one component's shared-memory array (8 * position < 65536 for cap <= 8191):
        // the kernels add them to the array base without a shift
        for (int64_t p = slice_ptr[s]; p < slice_ptr[s + 1]; ++p) out.lcols[(size_t)p] = (uint16_t)(8 * pos[(size_t)scols[p]]);
      }
      P.ends.push_back(j_end);
      P.gsize.push_back((int)gather.size());
      P.glist.insert(P.glist.end(), gather.begin(), gather.end());
      gather.clear();
      ++serial;
    };
    for (int g = g0; g < g1; ++g) {
      int begin = hint_ptr[g], count = 0;
      for (int j = hint_ptr[g]; j < hint_ptr[g + 1]; ++j) {
        const int s = slice_of(j);
        for (int pass = 0; pass < 2; ++pass) {
          fresh.clear();
          for (int64_t p = slice_ptr[s]; p < slice_ptr[s + 1]; ++p) {
            const int col = scols[p];
            if (stamp[(size_t)col] != serial) {
              stamp[(size_t)col] = serial;
              fresh.push_back(col);
            }
          }
          if (count > 0 && ((int)(gather.size() + fresh.size()) > cap || count >= max_slices)) {
            close(begin, j);  // bumps `serial`: the second pass collects this slice from scratch
            begin = j;
            count = 0;
            continue;
          }
          break;
        }
        if ((int)fresh.size() > cap) {
          P.error = 1;
          return;
        }
        gather.insert(gather.end(), fresh.begin(), fresh.end());
        ++count;
      }
      if (count > 0) close(begin, hint_ptr[g + 1]);
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  out.brick_ptr.push_back(0);
  out.gptr.push_back(0);
  for (const Part& P : parts) {
    if (P.error) out.error = P.error;
    for (size_t b = 0; b < P.ends.size(); ++b) {
      out.brick_ptr.push_back(P.ends[b]);
      out.gptr.push_back(out.gptr.back() + P.gsize[b]);
      out.max_gather = std::max(out.max_gather, P.gsize[b]);
    }
    out.glist.insert(out.glist.end(), P.glist.begin(), P.glist.end());
  }
}

// Work lists of the pipelined kernel for a grid of `grid` blocks of `warps` warps: block g works on bricks g, g + grid,
// ...; inside a brick the slices are dealt to the warps longest-processing-time-first.  wdesc holds one descriptor
// {brick, slice, first slot, length in steps} per slice, grouped by (block, warp) in the order the warp meets them;
// wseq[(g * warps + w)] is where the list of warp w of block g starts.
inline void assign_warps(int n_rows, const int* slice_ptr, const int* order, int warps, int grid, Bricks& B) {
  const int nb = B.n_bricks();
  const int n_slices = (n_rows + 31) / 32;
  grid = std::max(1, std::min(grid, nb));
  B.warps = warps;
  B.grid = grid;
  B.wdesc.assign((size_t)n_slices * 4, 0);
  B.wseq.assign((size_t)grid * warps + 1, 0);
  // pass 1: per brick, the LPT assignment (slice ids per warp)
  std::vector<std::vector<int>> mine((size_t)nb * warps);
  std::vector<std::pair<int, int>> byLen;  // (-length, position in the schedule)
  std::vector<long long> load((size_t)warps);
  for (int b = 0; b < nb; ++b) {
    byLen.clear();
    for (int j = B.brick_ptr[b]; j < B.brick_ptr[b + 1]; ++j) {
      const int s = order != nullptr ? order[j] : j;
      byLen.emplace_back(-(slice_ptr[s + 1] - slice_ptr[s]), j);
    }
    std::sort(byLen.begin(), byLen.end());
    std::fill(load.begin(), load.end(), 0);
    for (const auto& e : byLen) {
      int best = 0;
      for (int w = 1; w < warps; ++w)
        if (load[(size_t)w] < load[(size_t)best]) best = w;
      mine[(size_t)b * warps + best].push_back(order != nullptr ? order[e.second] : e.second);
      load[(size_t)best] += -e.first;
    }
  }
  // pass 2: lay the descriptors out by (block, warp)
  size_t out = 0;
  for (int g = 0; g < grid; ++g)
    for (int w = 0; w < warps; ++w) {
      B.wseq[(size_t)g * warps + w] = (int)out;
      for (int b = g; b < nb; b += grid)
        for (int s : mine[(size_t)b * warps + w]) {
          B.wdesc[4 * out + 0] = b;
          B.wdesc[4 * out + 1] = s;
          B.wdesc[4 * out + 2] = slice_ptr[s];
          B.wdesc[4 * out + 3] = (slice_ptr[s + 1] - slice_ptr[s]) >> 5;
          ++out;
        }
    }
  B.wseq[(size_t)grid * warps] = (int)out;
}

}  // namespace b2bricks
