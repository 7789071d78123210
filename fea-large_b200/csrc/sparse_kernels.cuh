// Sparse side of the hot path for sm_100a: deterministic gather assembly into 3x3-block
// CSR, Dirichlet cancellation, block SpMV and the fused vector kernels of the
// Jacobi-preconditioned CG.  All of this is HBM-bound streaming work: loads are
// lane-contiguous over the value array, reductions are fixed-order (no floating-point
// atomics), and the scalar recurrences of CG live on the device so a solve never waits
// for the host between iterations.
//
// Replaces libspmatrix as used by the reference (fea_solver.c:179,194,251,304,877,966,
// 1055,1255): sp_matrix_element_add -> gather_blocks_kernel, sp_matrix_cross_cancellation
// -> cancel_kernel, sp_matrix_yale_solve_cg -> the pcg_* kernels.
#pragma once
#include <cstdint>

#include "fea_plan.hpp"

namespace fea {

constexpr int RED_THREADS = 256;

// ---------------------------------------------------------------------------------
// fixed-order block reduction of up to 2 values; result valid in thread 0

template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *smem /* [NV][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    if (lane == 0) smem[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nw ? smem[k * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
}

// Grid-wide deterministic sum: every block deposits its partial, the last block to
// arrive adds the partials in index order.  Returns true in thread 0 of that last block
// with the totals in v.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials /* [NV][grid] */,
                                            unsigned int *counter, double *smem) {
  __shared__ bool is_last;
  block_reduce<NV>(v, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(size_t)k * gridDim.x + blockIdx.x] = v[k];
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  double t[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    t[k] = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x)
      t[k] += ((volatile double *)partials)[(size_t)k * gridDim.x + i];
  }
  __syncthreads();
  block_reduce<NV>(t, smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = t[k];
    *counter = 0u;
  }
  return threadIdx.x == 0;
}

// ---------------------------------------------------------------------------------
// SELL-32 block layout (see fea_plan.hpp): slice s = 32 rows (one per lane) x width_s block
// columns; slot = slice_ptr[s] + 32 j + lane; value (slot, c) at 9 (slice_ptr[s] + 32 j) + 32 c + lane.
// One warp walks one slice; every load and store below is 256 contiguous bytes per warp.

struct SellMat {
  int n_slices;
  const int32_t *slice_ptr;   // [n_slices + 1]
  const int32_t *sell_row;    // [n_slices * 32], -1 = padding lane
  const int32_t *bcol;        // [n_slots] column node
  double *vals;               // [n_slots * 9]
};

// 256-bit accesses to the value array (fea_plan.hpp: val_off): a lane's components 0-3 / 4-7 of one column
__device__ __forceinline__ void ld_vals4(const double *p, double &a, double &b, double &c, double &d) {
  asm("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st_vals4(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// the nine values of lane `lane` in the column starting at `colv` (= vals + 288 * column)
__device__ __forceinline__ void store_slot(double *colv, int lane, const double (&v)[9]) {
  st_vals4(colv + 4 * lane, v[0], v[1], v[2], v[3]);
  st_vals4(colv + 128 + 4 * lane, v[4], v[5], v[6], v[7]);
  colv[256 + lane] = v[8];
}
// the same block transposed, into the slot whose value index (288 * column + lane) is `mp`
__device__ __forceinline__ void store_slot_transposed(double *vals, int mp, const double (&v)[9]) {
  const int ml = mp % 288;
  double *colv = vals + (mp - ml);
  st_vals4(colv + 4 * ml, v[0], v[3], v[6], v[1]);
  st_vals4(colv + 128 + 4 * ml, v[4], v[7], v[2], v[5]);
  colv[256 + ml] = v[8];
}

// K3: gather assembly.  Each lane owns one block slot and sums its contributions in the order
// of the precomputed list (ascending global element id = the reference's element-major
// accumulation, fea_solver.c:878-882), so results are bit-reproducible run to run and need
// no atomics.  Optionally applies the Dirichlet cancellation in the same pass, from one flag byte
// per slot (the prescribed DOFs are fixed when the context is created).
// Work item = one slot column of one slice (32 slots, one warp).
// With `cmirror` (value index of the mirror slot, -1 = none; fea_plan.cpp "cell layout") only the upper triangle
// (column >= row) is summed: K_e is symmetric, the list of slot (J, I) is the list of (I, J) with every block
// transposed, so its sum is bit for bit the transpose -- the lane stores it into the mirror slot instead of a
// second lane gathering the same 72-byte blocks again.  Every staged block is then read exactly once; with that the
// kernel moves 5.7 + 2.8 GB of DRAM traffic in 1.79 ms on C3 (4.8 TB/s, 73 % of the measured copy peak:
// profiles/r2s_ncu_full_summary.md) -- it is HBM-bound on the K_e staging it has to read.
// one staged block: the aligned 80 bytes around it (five 16-byte loads)
__device__ __forceinline__ void load_staged(const double *__restrict__ Ke, uint32_t src, double (&w)[10]) {
  const uint32_t idx = src & 0x3fffffffu;       // 55 e + code -> 500 e + 100 pr + 9 pos (fea_plan.hpp); bit 30: see SRC_LAST
  // 72 bytes at an 8-byte boundary: five 16-byte loads of the enclosing aligned 80 bytes (the
  // extra double is the neighbouring block's or the region's pad) instead of nine 8-byte ones: the
  // 32 lanes of a warp read 32 unrelated blocks, so every load costs 32 L1 sector requests.
  // Three 32-byte LDG.256 of the enclosing 96 bytes were measured too: no faster, more data moved.
#if FEA_KE_INTERLEAVED
  const uint32_t el = idx / 55u, code = idx - 55u * el, rg = code / 11u;
  const uint32_t o = 100u * rg + 9u * (code - 11u * rg);       // offset inside the element, doubles
  const double2 *p = reinterpret_cast<const double2 *>(Ke) + ((size_t)(el >> 5) * 250 + (o >> 1)) * 32 + (el & 31u);
#pragma unroll
  for (int h = 0; h < 5; ++h) {
    const double2 t = p[h * 32];
    w[2 * h] = t.x;
    w[2 * h + 1] = t.y;
  }
#else
  const size_t off = (size_t)idx * 9 + idx / 11u;
  const double2 *p = reinterpret_cast<const double2 *>(Ke + (off & ~(size_t)1));
#pragma unroll
  for (int h = 0; h < 5; ++h) {
    const double2 t = p[h];
    w[2 * h] = t.x;
    w[2 * h + 1] = t.y;
  }
#endif
}
__device__ __forceinline__ void add_staged(uint32_t src, const double (&w)[10], double (&acc)[9]) {
  const uint32_t idx = src & 0x3fffffffu;
#if FEA_KE_INTERLEAVED
  const uint32_t code = idx % 55u, rg = code / 11u;
  const bool odd = ((100u * rg + 9u * (code - 11u * rg)) & 1u) != 0;
#else
  const bool odd = ((idx * 9u + idx / 11u) & 1u) != 0;     // parity survives the 32-bit wrap
#endif
  double v[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) v[c] = odd ? w[c + 1] : w[c];
  if (src >> 31) {  // stored block is K_e[b][a]: add its transpose
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] += v[(c % 3) * 3 + c / 3];
  } else {
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] += v[c];
  }
}

// (Two contributions per trip with the next list words prefetched -- fewer dependent round trips, 80-90 registers --
// was measured slower: 2.02-2.16 against 1.84 ms; the kernel wants resident warps, profiles/r2_assembly_variants.md.)
__device__ __forceinline__ void gather_item(const SellMat &A, const int32_t *__restrict__ cptr,
                                            const uint32_t *__restrict__ csrc, const double *__restrict__ Ke,
                                            const uint8_t *__restrict__ sflag, const int32_t *__restrict__ cmirror,
                                            int s, int j, int lane) {
  const int base = A.slice_ptr[s];
  const int slot = base + (j << 5) + lane;
  const int k0 = cptr[slot], k1 = cptr[slot + 1];
  int mp = -1;
  if (cmirror) {
    if (k1 > k0 && A.bcol[slot] < A.sell_row[s * 32 + lane]) return;   // lower triangle: written by its mirror
    mp = cmirror[slot];
  }
  const unsigned f = sflag ? sflag[slot] : 0u;   // bits 0-2: row DOFs prescribed, 3-5: column DOFs, 6: diagonal block
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  {
    for (int k = k0; k < k1; ++k) {
      const uint32_t src = csrc[k];
      double w[10];
      load_staged(Ke, src, w);
      add_staged(src, w, acc);
    }
  }
  if (f & 63u) {
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      const int i = c / 3, jj = c % 3;
      if ((((f >> i) | (f >> (3 + jj))) & 1u) && !((f & 64u) && i == jj)) acc[c] = 0.0;
    }
  }
  store_slot(A.vals + (size_t)(base + (j << 5)) * 9, lane, acc);
  if (mp >= 0) store_slot_transposed(A.vals, mp, acc);
}

// `split` consecutive CTAs per slice, their warps take the slot columns round-robin.  CTAs are
// dispatched in slice order, so the slices in flight are (resident CTAs) / split, and that window
// is what decides the DRAM traffic: every staged block has two readers (slot (i,j) and, transposed,
// slot (j,i)), and the second one only hits L2 if the window's staging (~166 KB per slice) fits
// there.  Measured on C3 (profiles/r1b_gather_variants.md): one warp per slice (9.5 k slices in
// flight) read 17.4 GB from DRAM for 7.2 GB of blocks; one 256-thread CTA per slice (740 in flight)
// 9.2 GB; 512 threads 6.9 GB; 1024 threads 4.9 GB -- but fat CTAs lose to drain/launch gaps, while
// 4 CTAs of 256 threads per slice keep 40 warps per SM on 185 slices: 2.34 -> 2.03 ms, and 8 CTAs
// of 128 threads 2.01 ms (the default).
// Dealing single columns to a persistent grid (no slice affinity: L1 reuse between the columns of a
// row is lost and the warps drift apart, 16 GB) and two warp-per-row mappings with lanes over (column,
// component) (coalesced 72-byte reads: 3.25 instead of 5 sector requests per block, but ~6x the
// instructions per contribution and one row's diagonal list, up to 24 deep, serialises its warp)
// were all slower (profiles/r1b_gather_variants.md).  Round 2 (profiles/r2_assembly_variants.md): the slices can also be
// taken from a list (`slice_list`), which is how the chunked assembly gathers what a chunk of elements completed.
template <int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
gather_blocks_kernel(SellMat A, int split, const int32_t *__restrict__ cptr, const uint32_t *__restrict__ csrc,
                           const double *__restrict__ Ke, const uint8_t *__restrict__ sflag /* may be null */,
                           const int32_t *__restrict__ cmirror /* null: every slot sums its own list */,
                           const int32_t *__restrict__ slice_list /* null: all slices, else n_list of them (chunked assembly) */,
                           int n_list) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5, nwarps = THREADS >> 5;
  // `split` consecutive CTAs share a slice (columns dealt round-robin over their warps): the same
  // number of resident warps then covers `split` times fewer slices, i.e. a smaller L2 footprint
  const int si = blockIdx.x / split, part = blockIdx.x - si * split;
  if (si >= n_list) return;
  const int s = slice_list ? slice_list[si] : si;
  const int width = (A.slice_ptr[s + 1] - A.slice_ptr[s]) >> 5;
  for (int j = part * nwarps + warp; j < width; j += nwarps * split) gather_item(A, cptr, csrc, Ke, sflag, cmirror, s, j, lane);
}

// K3, second mapping: NINE LANES PER BLOCK.  The lane-per-slot kernel above makes every 16-byte load of
// a warp a separate 32-byte sector request (one block per lane): 5.7 sector requests per contribution,
// and ncu pins it there -- the L1 moves ~1.3 sectors (or shared-memory wavefronts) per clock and SM.
// Here lane (g, c), g = lane / 9 < 3, c = lane % 9, owns component c, and group g streams the gather
// lists of slots [11 g, 11 g + 11) of a 32-slot column one contribution after the other: one 8-byte load
// per lane and contribution, the nine lanes of a group reading the 72 contiguous bytes of one staged
// block = exactly 3 sectors (a block stored transposed is the same load with c -> 3 (c % 3) + c / 3).  The
// lists of consecutive slots are consecutive in csrc, so a group walks ONE contiguous range: eight list
// words per group come in with one 32-byte load and are handed round by shuffles, and the eight block
// loads that follow are independent.  Bit 30 of a list word marks the last entry of its slot (set when
// the lists are uploaded), which is all the bookkeeping a contribution needs: add, and on the mark park
// the sum in a [32][9] tile (lane-contiguous, conflict-free) and start the next slot.  That relies on
// the slots with entries forming a prefix of every column -- rows are sorted by length inside a slice, so
// padding only trails; the context checks it when it uploads the lists and keeps the lane-per-slot kernel
// otherwise.  A slot's sum runs in list order in one register, exactly as in gather_item (same bits).
// The Dirichlet flags are applied where the column leaves the tile for the SELL value array (lane = slot
// again: nine coalesced 256-byte stores).
#ifndef FEA_G9_NOPRED
#define FEA_G9_NOPRED 1
#endif
#ifndef FEA_G9_MINCTAS
#define FEA_G9_MINCTAS 10
#endif
constexpr int GATHER9_UNROLL = 8;
constexpr uint32_t SRC_LAST = 0x40000000u;     // device copy of csrc only: last entry of its slot
constexpr uint32_t SRC_IDX_MASK = 0x3fffffffu;

template <int WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
gather_blocks9_kernel(SellMat A, int split, const int32_t *__restrict__ cptr, const uint32_t *__restrict__ csrc,
                      const double *__restrict__ Ke, const uint8_t *__restrict__ sflag /* may be null */) {
  __shared__ double tile_s[WARPS][32 * 9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x / split, part = blockIdx.x - s * split;
  if (s >= A.n_slices) return;
  const int base = A.slice_ptr[s];
  const int width = (A.slice_ptr[s + 1] - base) >> 5;
  const int g = lane / 9, c = lane - 9 * g;
  const double *KeC = Ke + c, *KeT = Ke + (3 * (c % 3) + c / 3);   // plain / transposed component of this lane
  const int s_begin = g < 3 ? 11 * g : 32, s_end = g < 3 ? (g == 2 ? 32 : 11 * g + 11) : 32;
  const int gl0 = 9 * (g < 3 ? g : 0);          // first lane of this lane's group
  const uint32_t last_mask = g < 3 ? SRC_LAST : 0u;   // lanes 27..31 only tag along
  double *tile = tile_s[warp];
  for (int j = part * WARPS + warp; j < width; j += WARPS * split) {
    const int slot0 = base + (j << 5);
    const int cp = cptr[slot0 + lane];
    const int cend = cptr[slot0 + 32];
#pragma unroll
    for (int c2 = 0; c2 < 9; ++c2) tile[c2 * 32 + lane] = 0.0;       // slots without entries stay zero
    int t = __shfl_sync(0xffffffffu, cp, s_begin & 31);
    int te = __shfl_sync(0xffffffffu, cp, s_end & 31);
    if (s_begin == 32) t = cend;
    if (s_end == 32) te = cend;
    double *tp = tile + s_begin * 9 + c;       // (never dereferenced by the idle lanes: their range is empty)
    double acc = 0.0;
    const int tmax = __reduce_max_sync(0xffffffffu, te - t);
    __syncwarp();
    for (int it = 0; it < tmax; it += GATHER9_UNROLL, t += GATHER9_UNROLL) {
      // lanes c < 8 of a group fetch its next eight list words (one sector), everyone gets them by shuffle
      uint32_t mine = 0u;
      if (c < GATHER9_UNROLL && t + c < te) mine = csrc[t + c];
      uint32_t src[GATHER9_UNROLL];
      double v[GATHER9_UNROLL];
#pragma unroll
      for (int u = 0; u < GATHER9_UNROLL; ++u) src[u] = __shfl_sync(0xffffffffu, mine, gl0 + u);
#pragma unroll
      for (int u = 0; u < GATHER9_UNROLL; ++u) {
        const uint32_t idx = src[u] & SRC_IDX_MASK;        // 55 e + code -> 500 e + 100 pr + 9 pos (fea_plan.hpp)
        const size_t off = (size_t)idx * 9 + idx / 11u;
#if FEA_G9_NOPRED
        // beyond its range a group's list word is 0: block 0 of element 0, a harmless (cached) load whose
        // value is added after the group's last mark and never stored
        v[u] = ((src[u] >> 31) ? KeT : KeC)[off];
#else
        v[u] = 0.0;
        if (t + u < te) v[u] = ((src[u] >> 31) ? KeT : KeC)[off];
#endif
      }
#pragma unroll
      for (int u = 0; u < GATHER9_UNROLL; ++u) {
        acc += v[u];                             // beyond the range v = 0 and no mark: harmless
        if (src[u] & last_mask) {
          *tp = acc;
          tp += 9;
          acc = 0.0;
        }
      }
    }
    __syncwarp();
    // lane = slot from here: Dirichlet flags, then nine coalesced stores
    const unsigned f = sflag ? sflag[slot0 + lane] : 0u;   // per-slot flags, see gather_item
    double a9[9];
#pragma unroll
    for (int c2 = 0; c2 < 9; ++c2) {
      double a = tile[lane * 9 + c2];
      const int i = c2 / 3, jj = c2 % 3;
      if ((f & 63u) && (((f >> i) | (f >> (3 + jj))) & 1u) && !((f & 64u) && i == jj)) a = 0.0;
      a9[c2] = a;
    }
    store_slot(A.vals + (size_t)slot0 * 9, lane, a9);
    __syncwarp();
  }
}

// K3, direct (push) assembly: the element kernel has written every contribution of an upper slot into its cell
// (fea_plan.cpp, "cell layout").  One warp per 32-slot SELL column: the column's cells are `kmax` consecutive
// layers, layer k = the k-th contribution of the m_k slots that have one, in the same slot order every layer, so
// the sum over layers runs position by position on the raw memory image: lane l owns 16-byte units l, l+32, ...
// of a layer (unit u = double2 `u % 5` of the cell at position u / 5) -- every load instruction of the warp reads
// one contiguous run (<= 512 bytes), against 32 unrelated 16-byte sectors per load in the pull kernels above.
// A slot's sum still runs in list order, one add per contribution starting from zero: bit-identical to
// gather_item.  The finished image is turned round through a [32 cells][10] shared tile (lane = slot again),
// the Dirichlet flags are applied, and each upper slot stores its block (coalesced along the lanes that are
// upper) and, transposed, the mirror slot of the lower triangle (K_e is symmetric, so is every partial sum).
// Lower and padding slots are never written by their own column: padding stays zero from the allocation.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
gather_cells_kernel(int n_cols, const int32_t *__restrict__ col_order, const int32_t *__restrict__ ccell,
                    const uint16_t *__restrict__ cmeta, const int32_t *__restrict__ cmirror,
                    const double2 *__restrict__ cells, double *__restrict__ vals,
                    const uint8_t *__restrict__ sflag /* may be null */, int dbg /* diagnostics: 1 no mirror stores, 2 no own stores */) {
  __shared__ __align__(16) double tile_s[WARPS][32 * CELL_DOUBLES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x * WARPS + warp;
  if (w >= n_cols) return;
  const int col = col_order[w];
  const size_t slot = (size_t)col * 32 + lane;
  const unsigned meta = cmeta[slot];
  const int n = meta & CELL_MAX_CONTRIB, rk = meta >> CELL_RANK_SHIFT;
  const int kmax = __reduce_max_sync(0xffffffffu, n);
  const double2 *layer = cells + (size_t)ccell[col] * 5;
  double2 acc[5];
#pragma unroll
  for (int h = 0; h < 5; ++h) acc[h] = make_double2(0.0, 0.0);
  int k = 0;
  for (; k + 1 < kmax; k += 2) {   // two layers in flight
    const int m0 = __popc(__ballot_sync(0xffffffffu, n > k)), m1 = __popc(__ballot_sync(0xffffffffu, n > k + 1));
    const double2 *l0 = layer, *l1 = layer + 5 * m0;
    double2 a[5], b[5];
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      const int u = lane + 32 * h;
      a[h] = u < 5 * m0 ? __ldcs(l0 + u) : make_double2(0.0, 0.0);
      b[h] = u < 5 * m1 ? __ldcs(l1 + u) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      const int u = lane + 32 * h;
      if (u < 5 * m0) { acc[h].x += a[h].x; acc[h].y += a[h].y; }
      if (u < 5 * m1) { acc[h].x += b[h].x; acc[h].y += b[h].y; }
    }
    layer += 5 * (m0 + m1);
  }
  if (k < kmax) {
    const int m0 = __popc(__ballot_sync(0xffffffffu, n > k));
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      const int u = lane + 32 * h;
      if (u < 5 * m0) {
        const double2 a = __ldcs(layer + u);
        acc[h].x += a.x;
        acc[h].y += a.y;
      }
    }
  }
  double2 *tile2 = reinterpret_cast<double2 *>(tile_s[warp]);
#pragma unroll
  for (int h = 0; h < 5; ++h) tile2[lane + 32 * h] = acc[h];
  __syncwarp();
  if (n > 0) {
    const double *mine = tile_s[warp] + rk * CELL_DOUBLES;
    double v[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = mine[c];
    const unsigned f = sflag ? sflag[slot] : 0u;   // bits 0-2: row DOFs prescribed, 3-5: column DOFs, 6: diagonal block
    if (f & 63u) {
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const int i = c / 3, jj = c % 3;
        if ((((f >> i) | (f >> (3 + jj))) & 1u) && !((f & 64u) && i == jj)) v[c] = 0.0;
      }
    }
    if (!(dbg & 2)) store_slot(vals + (size_t)col * (32 * 9), lane, v);
    const int mp = (dbg & 1) ? -1 : cmirror[slot];
    if (mp >= 0) store_slot_transposed(vals, mp, v);
  }
}

// residual gather: R[3I+i] = sum over (element, a) touching node I of R_e[a][i]
__global__ void __launch_bounds__(256)
gather_residual_kernel(int n_rows, const int32_t *__restrict__ rptr, const int32_t *__restrict__ rsrc,
                       const double *__restrict__ Re, int ne_pad, double *__restrict__ R,
                       const uint8_t *__restrict__ pflag /* may be null: zero prescribed rows */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n_rows) return;
  const int row = t / 3, i = t - 3 * row;
  double acc = 0.0;
  for (int k = rptr[row]; k < rptr[row + 1]; ++k) {
    const int s = rsrc[k];
    const int e = s / 10, a = s - 10 * e;
    acc += Re[(size_t)(a * 3 + i) * ne_pad + e];
  }
  if (pflag && pflag[t]) acc = 0.0;
  R[t] = acc;
}

// K4: zero rows and columns of prescribed DOFs keeping the diagonal (sp_matrix_cross_cancellation
// as used at fea_solver.c:1255); RHS rows become diag * presc (:1256)
__global__ void __launch_bounds__(256)
cancel_kernel(SellMat A, const uint8_t *__restrict__ sflag) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < A.n_slices; s += warps_per_grid) {
    const int base = A.slice_ptr[s];
    const int width = (A.slice_ptr[s + 1] - base) >> 5;
    for (int j = 0; j < width; ++j) {
      const unsigned f = sflag[base + (j << 5) + lane];   // per-slot flags, see gather_item
      if (!(f & 63u)) continue;
      double *out = A.vals + (size_t)(base + (j << 5)) * 9;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const int i = c / 3, jj = c % 3;
        if ((((f >> i) | (f >> (3 + jj))) & 1u) && !((f & 64u) && i == jj)) out[val_off(c, lane)] = 0.0;
      }
    }
  }
}

// sdiag[row] = 288 * column + lane of the row's diagonal block; its (i,i) entry is component 4 i (fea_plan.hpp: val_off)
__device__ __forceinline__ size_t diag_index(int sd, int i) {
  const int lane = sd % 288;
  return (size_t)(sd - lane) + val_off(4 * i, lane);
}
__global__ void rhs_fix_kernel(int n, const double *__restrict__ vals, const int32_t *__restrict__ sdiag,
                               const uint8_t *__restrict__ pflag, const double *__restrict__ pval,
                               double lambda, double *__restrict__ R) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (pflag[t]) {
    const int row = t / 3, i = t - 3 * row;
    R[t] = vals[diag_index(sdiag[row], i)] * (pval[t] * lambda);
  }
}

__global__ void jacobi_kernel(int n, const double *__restrict__ vals, const int32_t *__restrict__ sdiag,
                              double *__restrict__ dinv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int row = t / 3, i = t - 3 * row;
  const double d = vals[diag_index(sdiag[row], i)];
  dinv[t] = d != 0.0 ? 1.0 / d : 1.0;
}

// ---------------------------------------------------------------------------------
// K5: y = A x, 3x3 blocks in SELL-32.  Lane = row, no shuffles, no index arithmetic beyond
// the slot stride; x is gathered through L1/L2 (24 contiguous bytes per block).  Optionally
// fuses the partial of x_row . y_row (p.Ap of CG).

struct PcgCtl {
  double pq;        // p . A p (global)
  double rz_old;    // r . z of the previous iteration
  double rz_new;
  double rr;        // r . r
  double bb;        // b . b
  double thresh;    // stop when rr <= thresh
  double beta;
  double best_rr;   // smallest r.r seen so far (stagnation guard)
  double rr_saved;  // r.r of the checkpointed iterate (u_saved)
  double rr_exit;   // r.r latched when `done` was set: with several ranks the queued no-op iterations
                    // that follow keep all-reducing pq / rz_new / rr, which are garbage from then on
  int save;         // this iteration's direction kernel must checkpoint u
  int done;         // 1 = tolerance met, 2 = stalled or diverged: the checkpoint is the answer
  int iters;
  int stall;        // iterations since best_rr last improved
  int stall_limit;
};

template <bool FUSE_DOT>
__global__ void __launch_bounds__(256)
spmv_sell_kernel(int n_slices, const int32_t *__restrict__ slice_ptr, const int32_t *__restrict__ sell_row,
                 const int32_t *__restrict__ bcol, const double *__restrict__ vals,
                 const double *__restrict__ x, double *__restrict__ y, double *partials,
                 unsigned int *counter, double *dot_out, const int *done_flag,
                 const int32_t *__restrict__ slice_list /* null: all slices, else n_slices entries */) {
  __shared__ double red[32];
  if (FUSE_DOT && *done_flag) return;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  double dot = 0.0;
  for (int si = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; si < n_slices; si += warps_per_grid) {
    const int s = slice_list ? slice_list[si] : si;
    const int base = slice_ptr[s];
    const int width = (slice_ptr[s + 1] - base) >> 5;
    const int row = sell_row[s * 32 + lane];
    const double *v = vals + (size_t)base * 9 + 4 * lane;
    const int32_t *bc = bcol + base + lane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 2
    for (int j = 0; j < width; ++j) {
      const int col = bc[j << 5];
      const double *xc = x + 3 * (size_t)col;
      const double x0 = xc[0], x1 = xc[1], x2 = xc[2];
      const double *vj = v + (size_t)j * 288;
      double k0, k1, k2, k3, k4, k5, k6, k7;
      ld_vals4(vj, k0, k1, k2, k3);
      ld_vals4(vj + 128, k4, k5, k6, k7);
      const double k8 = __ldg(vj + 256 - 3 * lane);
      a0 = fma(k0, x0, fma(k1, x1, fma(k2, x2, a0)));
      a1 = fma(k3, x0, fma(k4, x1, fma(k5, x2, a1)));
      a2 = fma(k6, x0, fma(k7, x1, fma(k8, x2, a2)));
    }
    if (row >= 0) {
      y[3 * (size_t)row] = a0;
      y[3 * (size_t)row + 1] = a1;
      y[3 * (size_t)row + 2] = a2;
      if (FUSE_DOT)
        dot += a0 * x[3 * (size_t)row] + a1 * x[3 * (size_t)row + 1] + a2 * x[3 * (size_t)row + 2];
    }
  }
  if (FUSE_DOT) {
    double t[1] = {dot};
    if (grid_reduce<1>(t, partials, counter, red)) *dot_out = t[0];
  }
}

// ---------------------------------------------------------------------------------
// PCG vector kernels

// r = b - q (q = A x0) or r = b; z = dinv r; p = z; sums r.z, b.b, r.r
__global__ void __launch_bounds__(RED_THREADS)
pcg_init_kernel(int n, const double *__restrict__ b, const double *__restrict__ q /* null: x0 = 0 */,
                const double *__restrict__ dinv, double *__restrict__ r, double *__restrict__ p,
                double *partials, unsigned int *counter, PcgCtl *ctl, double tol, int abs_tol, int finalize) {
  __shared__ double red[3 * 32];
  double s[3] = {0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double bt = b[t];
    const double rt = q ? bt - q[t] : bt;
    const double zt = dinv[t] * rt;
    r[t] = rt;
    p[t] = zt;
    s[0] += rt * zt;
    s[1] += bt * bt;
    s[2] += rt * rt;
  }
  if (grid_reduce<3>(s, partials, counter, red)) {
    ctl->rz_old = s[0];
    ctl->rz_new = s[0];
    ctl->bb = s[1];
    ctl->rr = s[2];
    ctl->iters = 0;
    ctl->beta = 0.0;
    ctl->stall = 0;
    ctl->best_rr = s[2];
    ctl->rr_saved = s[2];
    ctl->rr_exit = s[2];
    ctl->save = 0;
    if (finalize) {
      ctl->thresh = abs_tol ? tol * tol : tol * tol * s[1];
      ctl->done = (s[1] == 0.0 || s[2] <= ctl->thresh) ? 1 : 0;
    }
  }
}

// multi-rank: fold the all-reduced sums of pcg_init into the control block
__global__ void pcg_init_finalize_kernel(PcgCtl *ctl, double tol, int abs_tol) {
  ctl->rz_old = ctl->rz_new;
  ctl->best_rr = ctl->rr;
  ctl->rr_saved = ctl->rr;
  ctl->rr_exit = ctl->rr;
  ctl->save = 0;
  ctl->thresh = abs_tol ? tol * tol : tol * tol * ctl->bb;
  ctl->done = (ctl->bb == 0.0 || ctl->rr <= ctl->thresh) ? 1 : 0;
}

// Scalar bookkeeping of one CG iteration (one thread).  Besides the recurrences it guards the
// solve against the two ways CG fails on the reference's "analytical" models, whose K is singular
// (free rotation about y) so that a right-hand side at rounding level is inconsistent: ||r|| then
// bottoms out and the iterate is slowly polluted along the null space, finally blowing up.
//  - every time r.r has halved since the last checkpoint the direction kernel copies u aside
//    (at most ~100 cheap copies per solve);
//  - r.r not improving for stall_limit iterations, or exceeding 1e14 x its best (CG's residual norm
//    is not monotone: spikes of 1e4 in ||r|| were seen on healthy 18 M-DOF SPD solves), ends the solve
//    with done = 2 and the host hands back the checkpoint, a clean near-minimum-residual iterate.
__device__ __forceinline__ void pcg_step_control(PcgCtl *ctl) {
  ctl->beta = ctl->rz_new / ctl->rz_old;
  ctl->rz_old = ctl->rz_new;
  ctl->iters += 1;
  ctl->save = 0;
  ctl->rr_exit = ctl->rr;
  if (ctl->rr <= ctl->thresh) {
    ctl->done = 1;
    return;
  }
  if (!(ctl->rr == ctl->rr) || ctl->rr > 1e14 * ctl->best_rr) {   // NaN or diverging
    ctl->done = 2;
    return;
  }
  if (ctl->rr < 0.5 * ctl->rr_saved) {
    ctl->rr_saved = ctl->rr;
    ctl->save = 1;
  }
  if (ctl->rr < 0.999 * ctl->best_rr) {
    ctl->best_rr = ctl->rr;
    ctl->stall = 0;
  } else if (++ctl->stall >= ctl->stall_limit) {
    ctl->done = 2;
  }
}

// u += alpha p; r -= alpha q; sums r.(dinv r), r.r.  With `finalize` (single rank) the last
// block also advances the scalar recurrences.
__global__ void __launch_bounds__(RED_THREADS)
pcg_update_kernel(int n, const double *__restrict__ p, const double *__restrict__ q,
                  const double *__restrict__ dinv, double *__restrict__ u, double *__restrict__ r,
                  double *partials, unsigned int *counter, PcgCtl *ctl, int finalize) {
  __shared__ double red[2 * 32];
  if (ctl->done) return;
  const double alpha = ctl->rz_old / ctl->pq;
  double s[2] = {0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    u[t] = fma(alpha, p[t], u[t]);
    const double rt = fma(-alpha, q[t], r[t]);
    r[t] = rt;
    s[0] += rt * rt * dinv[t];
    s[1] += rt * rt;
  }
  if (grid_reduce<2>(s, partials, counter, red)) {
    ctl->rz_new = s[0];
    ctl->rr = s[1];
    if (finalize) pcg_step_control(ctl);
  }
}

__global__ void pcg_control_kernel(PcgCtl *ctl) {
  if (!ctl->done) pcg_step_control(ctl);
}

// p = dinv r + beta p; when the control block asks for it, checkpoint u
__global__ void __launch_bounds__(256)
pcg_direction_kernel(int n, const double *__restrict__ r, const double *__restrict__ dinv,
                     double *__restrict__ p, const double *__restrict__ u, double *__restrict__ u_saved,
                     const PcgCtl *ctl) {
  if (ctl->done) return;
  const double beta = ctl->beta;
  const bool save = ctl->save != 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    p[t] = fma(beta, p[t], dinv[t] * r[t]);
    if (save) u_saved[t] = u[t];
  }
}

// ---------------------------------------------------------------------------------
// Single-reduction PCG (Chronopoulos & Gear): the recurrences are rearranged so that the three sums
// of an iteration -- gamma = r.z, delta = w.z (w = A z), r.r -- are all known at the same point, i.e. ONE
// all-reduce per iteration across ranks instead of two (p.Ap, then r.z / r.r), and one vector kernel
// instead of two:
//     beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)
//     p = z + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  z = D^-1 r      (pcg2_step_kernel)
//     [halo exchange of z]   w = A z, delta = w.z                                     (spmv_sell_kernel)
//     all-reduce (gamma, delta_a, delta_b, r.r)
// Same iterates as classic PCG in exact arithmetic (checked: equal iteration counts).  The state is
// double-buffered: step k reads S[k & 1] (complete) and fills S[(k + 1) & 1], every block re-deriving
// the same scalars and stop decision from the same inputs, so there is no one-thread control kernel and
// no block ever reads a field another block of the same launch writes.  Guards as in pcg_step_control.

struct Pcg2State {
  double gamma, delta_a, delta_b, rr;    // the four all-reduced sums (contiguous), delta split interior / boundary
  double bb;                             // contiguous with the sums for the 5-double all-reduce of the start
  double gamma_old, alpha_old;
  double thresh, best_rr, rr_saved, rr_exit;
  int done, iters, stall, stall_limit;
};

// r = b - q (q = A x0) or r = b; z = dinv r; p = s = 0; sums r.z, r.r, b.b into S0
__global__ void __launch_bounds__(RED_THREADS)
pcg2_init_kernel(int n, const double *__restrict__ b, const double *__restrict__ q /* null: x0 = 0 */,
                 const double *__restrict__ dinv, double *__restrict__ r, double *__restrict__ z,
                 double *__restrict__ p, double *__restrict__ sv, double *partials, unsigned int *counter,
                 Pcg2State *S0) {
  __shared__ double red[3 * 32];
  double s[3] = {0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double bt = b[t];
    const double rt = q ? bt - q[t] : bt;
    const double zt = dinv[t] * rt;
    r[t] = rt;
    z[t] = zt;
    p[t] = 0.0;
    sv[t] = 0.0;
    s[0] += rt * zt;
    s[1] += rt * rt;
    s[2] += bt * bt;
  }
  if (grid_reduce<3>(s, partials, counter, red)) {
    S0->gamma = s[0];
    S0->delta_a = 0.0;
    S0->delta_b = 0.0;
    S0->rr = s[1];
    S0->bb = s[2];
    S0->done = 0;
  }
}

// after the (all-reduced) sums of the start are in S0
__global__ void pcg2_finalize_kernel(Pcg2State *S0, double tol, int abs_tol, int stall_limit) {
  S0->gamma_old = S0->gamma;
  S0->alpha_old = 1.0;
  S0->thresh = abs_tol ? tol * tol : tol * tol * S0->bb;
  S0->best_rr = S0->rr;
  S0->rr_saved = S0->rr;
  S0->rr_exit = S0->rr;
  S0->iters = 0;
  S0->stall = 0;
  S0->stall_limit = stall_limit;
  S0->done = (S0->bb == 0.0 || S0->rr <= S0->thresh) ? 1 : 0;
}

__global__ void __launch_bounds__(RED_THREADS)
pcg2_step_kernel(int n, const Pcg2State *__restrict__ Sc, Pcg2State *Sn, double *__restrict__ z,
                 const double *__restrict__ w, double *__restrict__ p, double *__restrict__ sv,
                 double *__restrict__ x, double *__restrict__ r, const double *__restrict__ dinv,
                 double *__restrict__ x_saved, double *partials, unsigned int *counter) {
  __shared__ double red[2 * 32];
  const Pcg2State c = *Sc;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
  if (c.done) {   // queued iterations after the end: carry the final state along
    if (lead) *Sn = c;
    return;
  }
  int done = 0;
  const double delta = c.delta_a + c.delta_b;
  double beta = 0.0, denom = delta;
  if (c.iters > 0) {
    beta = c.gamma / c.gamma_old;
    denom = delta - beta * c.gamma / c.alpha_old;
  }
  const double alpha = c.gamma / denom;
  if (c.rr <= c.thresh) done = 1;
  else if (!(c.rr == c.rr) || c.rr > 1e14 * c.best_rr || !(denom > 0.0)) done = 2;   // NaN, divergence, breakdown
  bool save = false;
  Pcg2State nx = c;
  nx.rr_exit = c.rr;
  if (!done) {
    if (c.rr < 0.5 * c.rr_saved) {
      nx.rr_saved = c.rr;
      save = true;
    }
    if (c.rr < 0.999 * c.best_rr) {
      nx.best_rr = c.rr;
      nx.stall = 0;
    } else if (++nx.stall >= c.stall_limit) {
      done = 2;
      save = false;
      nx.rr_saved = c.rr_saved;
    }
  }
  nx.done = done;
  if (done) {
    if (lead) *Sn = nx;
    return;
  }
  if (lead) {   // everything but the three sums this iteration produces
    Sn->bb = c.bb;
    Sn->gamma_old = c.gamma;
    Sn->alpha_old = alpha;
    Sn->thresh = c.thresh;
    Sn->best_rr = nx.best_rr;
    Sn->rr_saved = nx.rr_saved;
    Sn->rr_exit = c.rr;
    Sn->done = 0;
    Sn->iters = c.iters + 1;
    Sn->stall = nx.stall;
    Sn->stall_limit = c.stall_limit;
  }
  double s[2] = {0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double pt = fma(beta, p[t], z[t]);
    const double st = fma(beta, sv[t], w[t]);
    const double xt = x[t];
    if (save) x_saved[t] = xt;       // the iterate r.r = c.rr belongs to
    x[t] = fma(alpha, pt, xt);
    const double rt = fma(-alpha, st, r[t]);
    const double zt = dinv[t] * rt;
    p[t] = pt;
    sv[t] = st;
    r[t] = rt;
    z[t] = zt;
    s[0] += rt * zt;
    s[1] += rt * rt;
  }
  if (grid_reduce<2>(s, partials, counter, red)) {
    Sn->gamma = s[0];
    Sn->rr = s[1];
  }
}

// ---------------------------------------------------------------------------------
// Chebyshev-accelerated Jacobi: the stronger preconditioner behind the task files' PCG_ILU request
// (fea_solver.c:260-280 builds an incomplete factorisation there; triangular solves do not map onto
// 148 SMs, a fixed polynomial in D^-1 A does -- it is a fixed SPD operator, so CG stays CG).
//   z = p_d(D^-1 A) D^-1 r: d steps of the Chebyshev iteration for A z = r on [lmax / ratio, lmax]
//     d_0 = D^-1 r / theta, z = d_0;   d_k = rho_k rho_{k-1} d_{k-1} + (2 rho_k / delta) D^-1 (r - A z),  z += d_k

__global__ void cheb_first_kernel(int n, const double *__restrict__ r, const double *__restrict__ dinv,
                                  double inv_theta, double *__restrict__ d, double *__restrict__ z) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double v = dinv[t] * r[t] * inv_theta;
    d[t] = v;
    z[t] = v;
  }
}

__global__ void cheb_step_kernel(int n, const double *__restrict__ r, const double *__restrict__ w /* A z */,
                                 const double *__restrict__ dinv, double c1, double c2, double *__restrict__ d,
                                 double *__restrict__ z) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double v = fma(c1, d[t], c2 * (dinv[t] * (r[t] - w[t])));
    d[t] = v;
    z[t] += v;
  }
}

// p = z + beta p (general preconditioner); when the control block asks for it, checkpoint u
__global__ void __launch_bounds__(256)
pcg_direction_z_kernel(int n, const double *__restrict__ z, double *__restrict__ p, const double *__restrict__ u,
                       double *__restrict__ u_saved, const PcgCtl *ctl) {
  if (ctl->done) return;
  const double beta = ctl->beta;
  const bool save = ctl->save != 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    p[t] = fma(beta, p[t], z[t]);
    if (save) u_saved[t] = u[t];
  }
}

// v <- D^-1 w / sqrt(*nrm2_prev) ... pieces of the power iteration that bounds the spectrum of D^-1 A
__global__ void scale_dinv_kernel(int n, const double *__restrict__ w, const double *__restrict__ dinv,
                                  const double *__restrict__ nrm2 /* null: no scaling */, double *__restrict__ v) {
  const double s = nrm2 ? rsqrt(*nrm2) : 1.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) v[t] = dinv[t] * w[t] * s;
}
__global__ void power_start_kernel(int n, double *__restrict__ v) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
    v[t] = 1.0 + 0.25 * (double)((t * 2654435761u) >> 29);   // fixed, not orthogonal to anything in particular
}

// out = a . b (fixed order)
__global__ void __launch_bounds__(RED_THREADS)
dot_kernel(int n, const double *__restrict__ a, const double *__restrict__ b, double *partials,
           unsigned int *counter, double *out) {
  __shared__ double red[32];
  double s[1] = {0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) s[0] += a[t] * b[t];
  if (grid_reduce<1>(s, partials, counter, red)) *out = s[0];
}

// ---------------------------------------------------------------------------------
// small vector utilities

__global__ void axpy_kernel(int n, double alpha, const double *__restrict__ x, double *__restrict__ y) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
    y[t] = fma(alpha, x[t], y[t]);
}

// x <- x + alpha (x - x_saved), x_saved <- the old x: secant predictor of a load sequence with equal increments
__global__ void extrapolate_kernel(int n, double alpha, double *__restrict__ x, double *__restrict__ xs) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const double cur = x[t];
    x[t] = fma(alpha, cur - xs[t], cur);
    xs[t] = cur;
  }
}

// x[dof] += lambda * value for the (pre-aggregated) prescribed DOFs (fea_solver.c:1259-1266)
__global__ void increment_kernel(int n, const int32_t *__restrict__ dof, const double *__restrict__ val,
                                 double lambda, double *__restrict__ x) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) x[dof[t]] += val[t] * lambda;
}

__global__ void pack_kernel(int n_nodes, const int32_t *__restrict__ nodes, const double *__restrict__ v,
                            double *__restrict__ buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 3 * n_nodes) buf[t] = v[3 * (size_t)nodes[t / 3] + t % 3];
}

// dst[l] = src[idx[l]] / dst[idx[l]] = src[l] for 3-vectors per node (host <-> local numbering)
__global__ void gather_nodes_kernel(int n_nodes, const int32_t *__restrict__ idx, const double *__restrict__ src,
                                    double *__restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 3 * n_nodes) dst[t] = src[3 * (size_t)idx[t / 3] + t % 3];
}
__global__ void scatter_nodes_kernel(int n_nodes, const int32_t *__restrict__ idx, const double *__restrict__ src,
                                     double *__restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 3 * n_nodes) dst[3 * (size_t)idx[t / 3] + t % 3] = src[t];
}

// [ng*9][ne_pad] -> [n_elems][ng][9] for elements flagged in `take`, scattered to global ids
__global__ void state_export_kernel(int n_elems, int ne_pad, int ng, const double *__restrict__ soa,
                                    double *__restrict__ aos) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)ng * 9;
  if (t >= (int64_t)n_elems * per) return;
  const int e = (int)(t / per), f = (int)(t - (int64_t)e * per);
  aos[t] = soa[(size_t)f * ne_pad + e];
}

// the same for a list of local elements (-1 = skip): aos[k][ng][9]
__global__ void state_export_list_kernel(int n_list, const int32_t *__restrict__ le, int ne_pad, int ng,
                                         const double *__restrict__ soa, double *__restrict__ aos) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)ng * 9;
  if (t >= (int64_t)n_list * per) return;
  const int k = (int)(t / per), f = (int)(t - (int64_t)k * per);
  const int e = le[k];
  aos[t] = e >= 0 ? soa[(size_t)f * ne_pad + e] : 0.0;
}

// peak probes -------------------------------------------------------------------
__global__ void dfma_probe_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// FP64 tensor-core probe: mma.sync m8n8k4 (256 FMA per warp instruction), eight independent accumulator
// pairs per warp so the issue rate, not the dependency chain, is measured
__global__ void dmma_probe_kernel(double *out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = k * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1])
                   : "d"(a), "d"(b));
  }
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void copy_probe_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
    dst[t] = src[t];
}

}  // namespace fea
