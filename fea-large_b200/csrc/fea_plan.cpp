// Host-side planning for the B200 hot path (no CUDA in this file).
//
//  * partition: recursive coordinate bisection of the nodes into `nranks` parts -- slabs for a bar,
//    boxes for a cube (rows of K follow node ownership; an element is
//    processed by every rank owning one of its nodes, so owned rows assemble locally).
//  * symbolic phase: the sparsity pattern the reference obtains dynamically through
//    sp_matrix_element_add (fea_solver.c:964-969, 1053-1058) -- a full 3x3 block for every
//    node pair sharing an element, explicit zeros included -- is built up front as block
//    CSR with ascending columns, together with the element->nonzero gather lists that make
//    assembly atomic-free and order-deterministic (element-ascending, as the reference's
//    element-major accumulation, fea_solver.c:878-882).
#include "fea_plan.hpp"

#include <omp.h>
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>

#include "../../include/fea_gpu.h"

namespace fea {

// 63-bit Morton key of a point on a 2^21 grid (uniform scale so cells are cubes)
static inline uint64_t spread21(uint64_t v) {
  v &= 0x1fffffULL;
  v = (v | (v << 32)) & 0x1f00000000ffffULL;
  v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
  v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
  v = (v | (v << 2)) & 0x1249249249249249ULL;
  return v;
}

struct MortonBox {
  double lo[3], inv;
  uint64_t key(const double *x) const {
    uint64_t q[3];
    for (int d = 0; d < 3; ++d) {
      double t = (x[d] - lo[d]) * inv;
      t = t < 0 ? 0 : (t > 2097151.0 ? 2097151.0 : t);
      q[d] = (uint64_t)t;
    }
    return spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
  }
};

static MortonBox bounding_box(int32_t n_nodes, const double *X0, double (&hi)[3]) {
  MortonBox b;
  for (int d = 0; d < 3; ++d) { b.lo[d] = 1e300; hi[d] = -1e300; }
  for (int32_t i = 0; i < n_nodes; ++i)
    for (int d = 0; d < 3; ++d) {
      b.lo[d] = std::min(b.lo[d], X0[3 * (size_t)i + d]);
      hi[d] = std::max(hi[d], X0[3 * (size_t)i + d]);
    }
  double ext = 0;
  for (int d = 0; d < 3; ++d) ext = std::max(ext, hi[d] - b.lo[d]);
  b.inv = ext > 0 ? 2097151.0 / ext : 0.0;
  return b;
}

// Recursive coordinate bisection: the node set is cut at the count-median along the axis of its largest
// extent (ties by id), ranks split in proportion, until every part has one rank.  A bar comes out as slabs
// along its length, a cube on 8 ranks as 2 x 2 x 2 boxes; any rank count works.
static void rcb(std::vector<int32_t> &idx, int64_t lo, int64_t hi, int rank0, int nr, const double *X0,
                std::vector<int32_t> &owner) {
  if (nr <= 1) {
    for (int64_t k = lo; k < hi; ++k) owner[(size_t)idx[(size_t)k]] = rank0;
    return;
  }
  double bl[3] = {1e300, 1e300, 1e300}, bh[3] = {-1e300, -1e300, -1e300};
  for (int64_t k = lo; k < hi; ++k)
    for (int d = 0; d < 3; ++d) {
      const double v = X0[3 * (size_t)idx[(size_t)k] + d];
      bl[d] = std::min(bl[d], v);
      bh[d] = std::max(bh[d], v);
    }
  // longest axis; on a tie the higher index wins (y before x, as the slab partition of a cube used to cut)
  int axis = 0;
  for (int d = 1; d < 3; ++d)
    if (bh[d] - bl[d] >= bh[axis] - bl[axis]) axis = d;
  const int nl = nr / 2;
  const int64_t mid = lo + (hi - lo) * nl / nr;
  std::sort(idx.begin() + lo, idx.begin() + hi, [&](int32_t a, int32_t b) {
    const double xa = X0[3 * (size_t)a + axis], xb = X0[3 * (size_t)b + axis];
    return xa < xb || (xa == xb && a < b);
  });
  rcb(idx, lo, mid, rank0, nl, X0, owner);
  rcb(idx, mid, hi, rank0 + nl, nr - nl, X0, owner);
}

// owner[] by recursive coordinate bisection; pos_in_owner[] = Morton rank of the node among its
// owner's nodes (every rank computes the same two arrays from the global mesh)
static void partition_nodes(Plan &p, int32_t n_nodes, const double *X0, int nranks, const MortonBox &box) {
  p.owner.assign((size_t)n_nodes, 0);
  std::vector<int32_t> order((size_t)n_nodes);
  std::iota(order.begin(), order.end(), 0);
  if (nranks > 1) rcb(order, 0, n_nodes, 0, nranks, X0, p.owner);
  std::vector<uint64_t> key((size_t)n_nodes);
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < n_nodes; ++i) key[(size_t)i] = box.key(X0 + 3 * (size_t)i);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    if (p.owner[(size_t)a] != p.owner[(size_t)b]) return p.owner[(size_t)a] < p.owner[(size_t)b];
    if (key[(size_t)a] != key[(size_t)b]) return key[(size_t)a] < key[(size_t)b];
    return a < b;
  });
  p.pos_in_owner.assign((size_t)n_nodes, 0);
  int32_t run = 0;
  for (int64_t k = 0; k < n_nodes; ++k) {
    const int32_t g = order[(size_t)k];
    if (k > 0 && p.owner[(size_t)g] != p.owner[(size_t)order[(size_t)k - 1]]) run = 0;
    p.pos_in_owner[(size_t)g] = run++;
  }
}

// Launchers export OMP_NUM_THREADS=1 for every rank (torchrun does); the planner is a one-off
// host phase, so it takes its fair share of the cores this process may run on instead.
static int plan_threads(int nranks) {
  cpu_set_t set;
  int cores = 1;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
  if (const char *s = getenv("FEA_PLAN_THREADS")) return std::max(1, atoi(s));
  return std::max(1, std::min(32, cores / std::max(1, nranks)));
}

struct PhaseTimer {   // FEA_PLAN_TIMING=1: where the host planning time goes (stderr)
  bool on = getenv("FEA_PLAN_TIMING") != nullptr;
  double t0 = omp_get_wtime();
  int rank;
  explicit PhaseTimer(int r) : rank(r) {}
  void lap(const char *what) {
    if (!on) return;
    const double t = omp_get_wtime();
    fprintf(stderr, "[plan rank %d] %-28s %.3f s\n", rank, what, t - t0);
    t0 = t;
  }
};

void build_plan(Plan &p, int32_t n_nodes, int32_t n_elems, const double *X0,
                const int32_t *conn, int rank, int nranks) {
  omp_set_num_threads(plan_threads(nranks));
  PhaseTimer tm(rank);
  if (n_nodes <= 0 || n_elems <= 0 || !X0 || !conn) throw std::runtime_error("empty mesh");
  if (rank < 0 || nranks < 1 || rank >= nranks) throw std::runtime_error("bad rank/nranks");
  for (int64_t k = 0; k < (int64_t)n_elems * NEN; ++k)
    if (conn[k] < 0 || conn[k] >= n_nodes) throw std::runtime_error("connectivity out of range");
  p.rank = rank;
  p.nranks = nranks;
  p.n_nodes_global = n_nodes;
  p.n_elems_global = n_elems;
  double box_hi[3];
  const MortonBox box = bounding_box(n_nodes, X0, box_hi);
  tm.lap("checks + bounding box");
  partition_nodes(p, n_nodes, X0, nranks, box);
  tm.lap("partition + numbering");
  const std::vector<int32_t> &owner = p.owner;
  const std::vector<int32_t> &pos = p.pos_in_owner;

  // ---- local elements, local node numbering ---------------------------------
  std::vector<uint8_t> used((size_t)n_nodes, 0);
  p.elem_gid.clear();
  for (int32_t e = 0; e < n_elems; ++e) {
    const int32_t *c = conn + (size_t)e * NEN;
    bool mine = false;
    for (int a = 0; a < NEN; ++a) mine |= (owner[(size_t)c[a]] == rank);
    if (!mine) continue;
    p.elem_gid.push_back(e);
    for (int a = 0; a < NEN; ++a) used[(size_t)c[a]] = 1;
  }
  {
    std::vector<std::pair<uint64_t, int32_t>> ek(p.elem_gid.size());
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)ek.size(); ++k) {
      const int32_t *c = conn + (size_t)p.elem_gid[(size_t)k] * NEN;
      double cen[3] = {0, 0, 0};
      for (int a = 0; a < 4; ++a)
        for (int d = 0; d < 3; ++d) cen[d] += 0.25 * X0[3 * (size_t)c[a] + d];
      ek[(size_t)k] = {box.key(cen), p.elem_gid[(size_t)k]};
    }
    std::sort(ek.begin(), ek.end());
    for (size_t k = 0; k < ek.size(); ++k) p.elem_gid[k] = ek[k].second;
  }
  p.n_elems = (int32_t)p.elem_gid.size();
  if ((int64_t)p.n_elems * NTRI >= (int64_t)0x7fffffff)
    throw std::runtime_error("too many local elements for 31-bit gather indices");

  p.node_gid.clear();
  for (int32_t i = 0; i < n_nodes; ++i)
    if (owner[(size_t)i] == rank) p.node_gid.push_back(i);
  p.n_own = (int32_t)p.node_gid.size();
  {
    std::vector<int32_t> mine(p.node_gid);
    for (int32_t g : mine) p.node_gid[(size_t)pos[(size_t)g]] = g;   // owned: by Morton rank
    std::vector<int32_t> ghosts;
    for (int32_t i = 0; i < n_nodes; ++i)
      if (used[(size_t)i] && owner[(size_t)i] != rank) ghosts.push_back(i);
    std::sort(ghosts.begin(), ghosts.end(), [&](int32_t a, int32_t b) {
      return owner[(size_t)a] < owner[(size_t)b] ||
             (owner[(size_t)a] == owner[(size_t)b] && pos[(size_t)a] < pos[(size_t)b]);
    });
    p.node_gid.insert(p.node_gid.end(), ghosts.begin(), ghosts.end());
  }
  p.n_local = (int32_t)p.node_gid.size();
  std::vector<int32_t> g2l((size_t)n_nodes, -1);
  for (int32_t l = 0; l < p.n_local; ++l) g2l[(size_t)p.node_gid[(size_t)l]] = l;

  p.conn.resize((size_t)p.n_elems * NEN);
  p.elem_owned.resize((size_t)p.n_elems);
  for (int32_t le = 0; le < p.n_elems; ++le) {
    const int32_t *c = conn + (size_t)p.elem_gid[(size_t)le] * NEN;
    for (int a = 0; a < NEN; ++a) p.conn[(size_t)le * NEN + a] = g2l[(size_t)c[a]];
    p.elem_owned[(size_t)le] = owner[(size_t)c[0]] == rank;
  }

  tm.lap("local elements + nodes");
  // ---- node -> (element, local node) lists for owned nodes (residual gather) ---
  const int32_t n_own = p.n_own;
  p.rptr.assign((size_t)n_own + 1, 0);
  for (int32_t le = 0; le < p.n_elems; ++le)
    for (int a = 0; a < NEN; ++a) {
      int32_t l = p.conn[(size_t)le * NEN + a];
      if (l < n_own) p.rptr[(size_t)l + 1]++;
    }
  for (int32_t i = 0; i < n_own; ++i) p.rptr[(size_t)i + 1] += p.rptr[(size_t)i];
  p.rsrc.resize((size_t)p.rptr[(size_t)n_own]);
  {
    std::vector<int32_t> cur(p.rptr.begin(), p.rptr.end() - 1);
    std::vector<int32_t> by_gid((size_t)p.n_elems);
    std::iota(by_gid.begin(), by_gid.end(), 0);
    std::sort(by_gid.begin(), by_gid.end(),
              [&](int32_t a, int32_t b) { return p.elem_gid[(size_t)a] < p.elem_gid[(size_t)b]; });
    for (int32_t le : by_gid)      // ascending GLOBAL element id: the reference's accumulation order
      for (int a = 0; a < NEN; ++a) {
        int32_t l = p.conn[(size_t)le * NEN + a];
        if (l < n_own) p.rsrc[(size_t)cur[(size_t)l]++] = le * NEN + a;
      }
  }

  tm.lap("residual lists");
  // ---- block pattern of the owned rows ---------------------------------------
  p.browptr.assign((size_t)n_own + 1, 0);
  std::vector<int32_t> rowlen((size_t)n_own, 0);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int32_t i = 0; i < n_own; ++i) {
      tmp.clear();
      tmp.push_back(i);
      for (int32_t k = p.rptr[(size_t)i]; k < p.rptr[(size_t)i + 1]; ++k) {
        const int32_t *c = &p.conn[(size_t)(p.rsrc[(size_t)k] / NEN) * NEN];
        tmp.insert(tmp.end(), c, c + NEN);
      }
      std::sort(tmp.begin(), tmp.end());
      rowlen[(size_t)i] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  }
  int64_t nnzb = 0;
  for (int32_t i = 0; i < n_own; ++i) {
    nnzb += rowlen[(size_t)i];
    if (nnzb * 9 >= (int64_t)0x7fffffff) throw std::runtime_error("local matrix exceeds 2^31 scalar nonzeros");
    p.browptr[(size_t)i + 1] = (int32_t)nnzb;
  }
  p.bcol.resize((size_t)nnzb);
  p.diag.resize((size_t)n_own);
  p.cptr.assign((size_t)nnzb + 1, 0);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 1024)
    for (int32_t i = 0; i < n_own; ++i) {
      tmp.clear();
      tmp.push_back(i);
      for (int32_t k = p.rptr[(size_t)i]; k < p.rptr[(size_t)i + 1]; ++k) {
        const int32_t *c = &p.conn[(size_t)(p.rsrc[(size_t)k] / NEN) * NEN];
        tmp.insert(tmp.end(), c, c + NEN);
      }
      std::sort(tmp.begin(), tmp.end());
      int32_t len = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
      int32_t *row = &p.bcol[(size_t)p.browptr[(size_t)i]];
      std::copy(tmp.begin(), tmp.begin() + len, row);
      p.diag[(size_t)i] = p.browptr[(size_t)i] + (int32_t)(std::lower_bound(row, row + len, i) - row);
      // contribution counts (row-private range of cptr)
      for (int32_t k = p.rptr[(size_t)i]; k < p.rptr[(size_t)i + 1]; ++k) {
        const int32_t *c = &p.conn[(size_t)(p.rsrc[(size_t)k] / NEN) * NEN];
        for (int b = 0; b < NEN; ++b) {
          int32_t pos = p.browptr[(size_t)i] + (int32_t)(std::lower_bound(row, row + len, c[b]) - row);
          p.cptr[(size_t)pos + 1]++;
        }
      }
    }
  }
  for (int64_t k = 0; k < nnzb; ++k) {
    int64_t s = (int64_t)p.cptr[(size_t)k] + p.cptr[(size_t)k + 1];
    if (s >= (int64_t)0x7fffffff) throw std::runtime_error("gather map exceeds 2^31 entries");
    p.cptr[(size_t)k + 1] = (int32_t)s;
  }
  p.csrc.resize((size_t)p.cptr[(size_t)nnzb]);
#pragma omp parallel
  {
    std::vector<int32_t> cur;
#pragma omp for schedule(dynamic, 1024)
    for (int32_t i = 0; i < n_own; ++i) {
      const int32_t r0 = p.browptr[(size_t)i], len = p.browptr[(size_t)i + 1] - r0;
      const int32_t *row = &p.bcol[(size_t)r0];
      cur.assign(p.cptr.begin() + r0, p.cptr.begin() + r0 + len);
      for (int32_t k = p.rptr[(size_t)i]; k < p.rptr[(size_t)i + 1]; ++k) {   // element-ascending
        const int32_t le = p.rsrc[(size_t)k] / NEN, a = p.rsrc[(size_t)k] % NEN;
        const int32_t *c = &p.conn[(size_t)le * NEN];
        for (int b = 0; b < NEN; ++b) {
          int32_t pos = (int32_t)(std::lower_bound(row, row + len, c[b]) - row);
          uint32_t src = (uint32_t)le * NTRI + (uint32_t)ke_code(std::min(a, b), std::max(a, b));
          if (a > b) src |= SRC_TRANSPOSE;   // stored block is K_e[b][a]; K_e[a][b] is its transpose
          p.csrc[(size_t)cur[(size_t)pos]++] = src;
        }
      }
    }
  }

  tm.lap("pattern + gather map");
  // ---- SELL-32-sigma layout of the same pattern -----------------------------------
  {
    std::vector<int32_t> perm((size_t)n_own);
    std::iota(perm.begin(), perm.end(), 0);
    for (int32_t w0 = 0; w0 < n_own; w0 += SELL_SIGMA) {
      const int32_t w1 = std::min(n_own, w0 + SELL_SIGMA);
      std::stable_sort(perm.begin() + w0, perm.begin() + w1,
                       [&](int32_t a, int32_t b) { return rowlen[(size_t)a] > rowlen[(size_t)b]; });
    }
    p.n_slices = (n_own + SELL_C - 1) / SELL_C;
    p.sell_row.assign((size_t)p.n_slices * SELL_C, -1);
    p.row_lane.assign((size_t)n_own, 0);
    p.slice_ptr.assign((size_t)p.n_slices + 1, 0);
    int64_t slots = 0;
    for (int32_t s = 0; s < p.n_slices; ++s) {
      int32_t width = 0;
      for (int l = 0; l < SELL_C; ++l) {
        const int64_t k = (int64_t)s * SELL_C + l;
        if (k >= n_own) break;
        const int32_t r = perm[(size_t)k];
        p.sell_row[(size_t)k] = r;
        p.row_lane[(size_t)r] = (int32_t)k;
        width = std::max(width, rowlen[(size_t)r]);
      }
      slots += (int64_t)width * SELL_C;
      if (slots * 9 >= (int64_t)0x7fffffff) throw std::runtime_error("local SELL matrix exceeds 2^31 values");
      p.slice_ptr[(size_t)s + 1] = (int32_t)slots;
    }
    p.sbcol.assign((size_t)slots, 0);
    p.scptr.assign((size_t)slots + 1, 0);
    p.sdiag.assign((size_t)n_own, 0);
    // slot of (row r, j-th block)
    auto slot_of = [&](int32_t r, int32_t j) {
      const int32_t k = p.row_lane[(size_t)r];
      return p.slice_ptr[(size_t)(k / SELL_C)] + j * SELL_C + (k % SELL_C);
    };
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < p.n_slices; ++s) {
      const int32_t width = (p.slice_ptr[(size_t)s + 1] - p.slice_ptr[(size_t)s]) / SELL_C;
      for (int l = 0; l < SELL_C; ++l) {
        const int32_t r = p.sell_row[(size_t)s * SELL_C + l];
        const int32_t len = r >= 0 ? rowlen[(size_t)r] : 0;
        for (int32_t j = 0; j < width; ++j) {
          const int32_t slot = p.slice_ptr[(size_t)s] + j * SELL_C + l;
          if (j < len) {
            const int32_t pcsr = p.browptr[(size_t)r] + j;
            p.sbcol[(size_t)slot] = p.bcol[(size_t)pcsr];
            p.scptr[(size_t)slot + 1] = p.cptr[(size_t)pcsr + 1] - p.cptr[(size_t)pcsr];
          } else {
            p.sbcol[(size_t)slot] = r >= 0 ? r : 0;   // padding: zero value times a finite x
          }
        }
        if (r >= 0) {
          const int32_t jd = p.diag[(size_t)r] - p.browptr[(size_t)r];
          p.sdiag[(size_t)r] = 9 * (p.slice_ptr[(size_t)s] + jd * SELL_C) + l;
        }
      }
    }
    for (int64_t k = 0; k < slots; ++k) p.scptr[(size_t)k + 1] += p.scptr[(size_t)k];
    p.scsrc.resize(p.csrc.size());
#pragma omp parallel for schedule(dynamic, 1024)
    for (int32_t r = 0; r < n_own; ++r)
      for (int32_t j = 0; j < rowlen[(size_t)r]; ++j) {
        const int32_t pcsr = p.browptr[(size_t)r] + j, slot = slot_of(r, j);
        std::copy(p.csrc.begin() + p.cptr[(size_t)pcsr], p.csrc.begin() + p.cptr[(size_t)pcsr + 1],
                  p.scsrc.begin() + p.scptr[(size_t)slot]);
      }
  }

  tm.lap("SELL layout");
  // ---- cell layout of the direct (push) assembly ---------------------------------------------
  // Upper slots only (column >= row in local numbering; ghost columns come after the owned rows, so every
  // slot with a ghost column is upper).  A 32-slot SELL column keeps its contributions ("cells", one 3x3
  // block each, CELL_DOUBLES doubles) as consecutive layers: layer k holds the k-th contribution (ascending
  // global element id, the reference's accumulation order, fea_solver.c:878-882) of every slot that has
  // one, the slots ordered by descending contribution count (ties by lane) -- so the same slot sits at the
  // same position of every layer and a layer is one contiguous run of memory.  The element kernel writes each
  // K_e block straight into its cell (transposed when the element's node order is the other way round), the
  // gather adds the layers position by position with fully coalesced loads, and writes the finished block to
  // its slot and, transposed, to the mirror slot of the lower triangle.
  {
    const int64_t n_slots = (int64_t)p.sbcol.size();
    const int64_t n_cols = n_slots / SELL_C;
    p.cmeta.assign((size_t)n_slots, 0);
    p.cmirror.assign((size_t)n_slots, -1);
    p.ccell.assign((size_t)n_cols + 1, 0);
    p.col_ready.assign((size_t)n_cols, -1);
    const int32_t ne_pad = (p.n_elems + 31) / 32 * 32;   // SoA pitch of the element arrays (whole element-kernel CTAs)
    p.edest.assign((size_t)NTRI * ne_pad, 0xffffffffu);
    std::vector<int32_t> slot_row((size_t)n_slots, -1);   // row of each real slot
#pragma omp parallel for schedule(static)
    for (int32_t s = 0; s < p.n_slices; ++s) {
      const int32_t width = (p.slice_ptr[(size_t)s + 1] - p.slice_ptr[(size_t)s]) / SELL_C;
      for (int l = 0; l < SELL_C; ++l) {
        const int32_t r = p.sell_row[(size_t)s * SELL_C + l];
        const int32_t len = r >= 0 ? rowlen[(size_t)r] : 0;
        for (int32_t j = 0; j < std::min(len, width); ++j) slot_row[(size_t)p.slice_ptr[(size_t)s] + (size_t)j * SELL_C + l] = r;
      }
    }
    // cells per column
    std::vector<int32_t> ncell((size_t)n_cols, 0);
    bool too_many = false;
#pragma omp parallel for schedule(static) reduction(|| : too_many)
    for (int64_t col = 0; col < n_cols; ++col) {
      int32_t tot = 0;
      for (int l = 0; l < SELL_C; ++l) {
        const int64_t slot = col * SELL_C + l;
        const int32_t r = slot_row[(size_t)slot];
        if (r < 0 || p.sbcol[(size_t)slot] < r) continue;
        const int32_t n = p.scptr[(size_t)slot + 1] - p.scptr[(size_t)slot];
        if (n > CELL_MAX_CONTRIB) too_many = true;
        tot += n;
      }
      ncell[(size_t)col] = tot;
    }
    if (too_many) throw std::runtime_error("a node pair is shared by more than 2047 elements");
    int64_t run = 0;
    for (int64_t col = 0; col < n_cols; ++col) {
      p.ccell[(size_t)col] = (int32_t)run;
      run += ncell[(size_t)col];
      if (run >= (int64_t)0x7fffffff) throw std::runtime_error("more than 2^31 stiffness cells on one rank");
    }
    p.ccell[(size_t)n_cols] = (int32_t)run;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t col = 0; col < n_cols; ++col) {
      int32_t n[SELL_C], order[SELL_C], rank[SELL_C];
      for (int l = 0; l < SELL_C; ++l) {
        const int64_t slot = col * SELL_C + l;
        const int32_t r = slot_row[(size_t)slot];
        n[l] = (r < 0 || p.sbcol[(size_t)slot] < r) ? 0 : p.scptr[(size_t)slot + 1] - p.scptr[(size_t)slot];
        order[l] = l;
      }
      std::stable_sort(order, order + SELL_C, [&](int a, int b) { return n[a] > n[b]; });
      for (int i = 0; i < SELL_C; ++i) rank[order[i]] = i;
      int32_t layer_base[CELL_MAX_CONTRIB + 2];
      {
        int32_t off = p.ccell[(size_t)col];
        const int32_t kmax = n[order[0]];
        int m = SELL_C;
        for (int32_t k = 0; k < kmax; ++k) {
          while (m > 0 && n[order[m - 1]] <= k) --m;   // slots with more than k contributions
          layer_base[k] = off;
          off += m;
        }
      }
      int32_t ready = -1;
      for (int l = 0; l < SELL_C; ++l) {
        const int64_t slot = col * SELL_C + l;
        if (n[l] == 0) continue;
        p.cmeta[(size_t)slot] = (uint16_t)(n[l] | (rank[l] << CELL_RANK_SHIFT));
        const int32_t r = slot_row[(size_t)slot], cnode = p.sbcol[(size_t)slot];
        for (int32_t k = 0; k < n[l]; ++k) {
          const uint32_t src = p.scsrc[(size_t)p.scptr[(size_t)slot] + (size_t)k];
          const uint32_t idx = src & 0x7fffffffu;
          const int32_t e = (int32_t)(idx / NTRI), code = (int32_t)(idx - (uint32_t)e * NTRI);
          p.edest[(size_t)code * ne_pad + e] = (uint32_t)(layer_base[k] + rank[l]) | (src & SRC_TRANSPOSE);
          ready = std::max(ready, e);
        }
        if (cnode > r && cnode < n_own) {   // mirror slot (cnode, r): binary search in cnode's ascending columns
          const int32_t lo = p.browptr[(size_t)cnode], hi = p.browptr[(size_t)cnode + 1];
          const int32_t *f = std::lower_bound(p.bcol.data() + lo, p.bcol.data() + hi, r);
          const int32_t jm = (int32_t)(f - (p.bcol.data() + lo));
          const int32_t km = p.row_lane[(size_t)cnode];
          p.cmirror[(size_t)slot] = 9 * (p.slice_ptr[(size_t)(km / SELL_C)] + jm * SELL_C) + (km % SELL_C);
        }
      }
      p.col_ready[(size_t)col] = ready;
    }
    // columns in the order their last contributing element is produced (chunked assembly: a column is
    // gathered as soon as the element tiles before that point are done, while its cells are still in L2)
    p.col_order.clear();
    for (int64_t col = 0; col < n_cols; ++col)
      if (p.col_ready[(size_t)col] >= 0) p.col_order.push_back((int32_t)col);
    std::stable_sort(p.col_order.begin(), p.col_order.end(),
                     [&](int32_t a, int32_t b) { return p.col_ready[(size_t)a] < p.col_ready[(size_t)b]; });
  }

  tm.lap("cell layout");
  // ---- halo lists ---------------------------------------------------------------
  p.nbr_rank.clear();
  p.send_ptr.assign(1, 0);
  p.recv_ptr.assign(1, 0);
  p.send_nodes.clear();
  if (nranks > 1) {
    std::vector<std::vector<int32_t>> send((size_t)nranks);
    for (int32_t e = 0; e < n_elems; ++e) {
      const int32_t *c = conn + (size_t)e * NEN;
      int own[NEN];
      bool has_me = false, has_other = false;
      for (int a = 0; a < NEN; ++a) {
        own[a] = owner[(size_t)c[a]];
        has_me |= own[a] == rank;
        has_other |= own[a] != rank;
      }
      if (!(has_me && has_other)) continue;
      for (int a = 0; a < NEN; ++a) {
        if (own[a] != rank) continue;
        for (int b = 0; b < NEN; ++b)
          if (own[b] != rank) send[(size_t)own[b]].push_back(c[a]);
      }
    }
    int32_t ghost_pos = 0;
    for (int q = 0; q < nranks; ++q) {
      std::vector<int32_t> &s = send[(size_t)q];
      std::sort(s.begin(), s.end(), [&](int32_t a, int32_t b) { return pos[(size_t)a] < pos[(size_t)b]; });
      s.erase(std::unique(s.begin(), s.end()), s.end());
      int32_t nrecv = 0;
      while (p.n_own + ghost_pos + nrecv < p.n_local &&
             owner[(size_t)p.node_gid[(size_t)(p.n_own + ghost_pos + nrecv)]] == q)
        ++nrecv;
      if (s.empty() && nrecv == 0) continue;
      p.nbr_rank.push_back(q);
      for (int32_t g : s) p.send_nodes.push_back(g2l[(size_t)g]);
      p.send_ptr.push_back((int32_t)p.send_nodes.size());
      ghost_pos += nrecv;
      p.recv_ptr.push_back(ghost_pos);
    }
  }
  tm.lap("halo lists");
}

}  // namespace fea

// ------------------------------------------------------------------------------
// C-ABI: planning and mesh generation (host only)

struct fea_plan {
  fea::Plan plan;
};

static thread_local std::string g_plan_error;

extern "C" int fea_plan_create(fea_plan_handle *out, int32_t n_nodes, int32_t n_elems,
                               const double *X0, const int32_t *conn, int32_t rank,
                               int32_t nranks) {
  if (!out) return FEA_GPU_ERR_ARG;
  fea_plan *p = new fea_plan();
  try {
    fea::build_plan(p->plan, n_nodes, n_elems, X0, conn, rank, nranks);
  } catch (const std::exception &e) {
    g_plan_error = e.what();
    delete p;
    return FEA_GPU_ERR_MESH;
  }
  *out = p;
  return FEA_GPU_OK;
}

extern "C" int fea_plan_destroy(fea_plan_handle p) {
  delete p;
  return FEA_GPU_OK;
}

namespace fea {
void plan_counts(const Plan &pl, int64_t out[16]) {
  std::memset(out, 0, sizeof(int64_t) * 16);
  out[0] = pl.n_own;
  out[1] = pl.n_local;
  out[2] = pl.n_elems;
  out[3] = pl.nnzb();
  out[4] = (int64_t)pl.csrc.size();
  out[5] = (int64_t)pl.nbr_rank.size();
  out[6] = (int64_t)pl.send_nodes.size();
  out[7] = pl.n_local - pl.n_own;
  out[8] = pl.n_nodes_global;
  out[9] = pl.n_elems_global;
  out[10] = pl.n_slots();
  out[11] = pl.n_slices;
  out[12] = pl.n_cells();
  out[13] = (int64_t)pl.col_order.size();
}
}  // namespace fea

extern "C" int fea_plan_counts(fea_plan_handle p, int64_t out[16]) {
  if (!p || !out) return FEA_GPU_ERR_ARG;
  fea::plan_counts(p->plan, out);
  return FEA_GPU_OK;
}

template <class T, class U>
static void copy_out(U *dst, const std::vector<T> &src) {
  if (dst && !src.empty()) std::memcpy(dst, src.data(), sizeof(T) * src.size());
}

extern "C" int fea_plan_arrays(fea_plan_handle p, int32_t *local_node_gid,
                               int32_t *local_elem_gid, int32_t *browptr, int32_t *bcol,
                               int32_t *cptr, uint32_t *csrc, int32_t *nbr_rank,
                               int32_t *send_ptr, int32_t *send_nodes, int32_t *recv_ptr) {
  if (!p) return FEA_GPU_ERR_ARG;
  const fea::Plan &pl = p->plan;
  copy_out(local_node_gid, pl.node_gid);
  copy_out(local_elem_gid, pl.elem_gid);
  copy_out(browptr, pl.browptr);
  copy_out(bcol, pl.bcol);
  copy_out(cptr, pl.cptr);
  copy_out(csrc, pl.csrc);
  copy_out(nbr_rank, pl.nbr_rank);
  copy_out(send_ptr, pl.send_ptr);
  copy_out(send_nodes, pl.send_nodes);
  copy_out(recv_ptr, pl.recv_ptr);
  return FEA_GPU_OK;
}

extern "C" int fea_plan_sell_arrays(fea_plan_handle p, int32_t *slice_ptr, int32_t *sell_row, int32_t *sbcol,
                                    int32_t *scptr, uint32_t *scsrc, int32_t *sdiag) {
  if (!p) return FEA_GPU_ERR_ARG;
  const fea::Plan &pl = p->plan;
  copy_out(slice_ptr, pl.slice_ptr);
  copy_out(sell_row, pl.sell_row);
  copy_out(sbcol, pl.sbcol);
  copy_out(scptr, pl.scptr);
  copy_out(scsrc, pl.scsrc);
  copy_out(sdiag, pl.sdiag);
  return FEA_GPU_OK;
}

extern "C" int fea_plan_cell_arrays(fea_plan_handle p, uint16_t *cmeta, int32_t *ccell, int32_t *cmirror, uint32_t *edest,
                                    int32_t *col_ready, int32_t *col_order) {
  if (!p) return FEA_GPU_ERR_ARG;
  const fea::Plan &pl = p->plan;
  copy_out(cmeta, pl.cmeta);
  copy_out(ccell, pl.ccell);
  copy_out(cmirror, pl.cmirror);
  copy_out(edest, pl.edest);
  copy_out(col_ready, pl.col_ready);
  copy_out(col_order, pl.col_order);
  return FEA_GPU_OK;
}

extern "C" int fea_plan_node_owner(fea_plan_handle p, int32_t *owner) {
  if (!p || !owner) return FEA_GPU_ERR_ARG;
  copy_out(owner, p->plan.owner);
  return FEA_GPU_OK;
}

// Kuhn / Freudenthal subdivision: each cube is cut into the 6 tets that share the body
// diagonal; all half-grid points become nodes (vertices + edge midpoints), so the node
// count is exactly (2nx+1)(2ny+1)(2nz+1).  Local node order follows the
// reference's shape functions (fea_solver.c:1287-1300): 0..3 vertices, 4=(0,1) 5=(1,2)
// 6=(0,2) 7=(0,3) 8=(1,3) 9=(2,3); vertices are ordered for a positive Jacobian.
extern "C" int fea_mesh_block(int32_t nx, int32_t ny, int32_t nz, double lx, double ly,
                              double lz, double y0, int32_t bc_style, double dy,
                              int64_t *n_nodes, int64_t *n_elems, int64_t *n_presc,
                              double *nodes, int32_t *conn, int32_t *presc_node,
                              int32_t *presc_type, double *presc_vals) {
  if (nx < 1 || ny < 1 || nz < 1) return FEA_GPU_ERR_ARG;
  const int64_t px = 2 * (int64_t)nx + 1, py = 2 * (int64_t)ny + 1, pz = 2 * (int64_t)nz + 1;
  const int64_t nn = px * py * pz, ne = 6 * (int64_t)nx * ny * nz, np = 2 * px * pz;
  if (nn >= 0x7fffffff || ne >= 0x7fffffff) return FEA_GPU_ERR_ARG;
  if (n_nodes) *n_nodes = nn;
  if (n_elems) *n_elems = ne;
  if (n_presc) *n_presc = np;
  auto nid = [&](int64_t jx, int64_t jy, int64_t jz) { return (int32_t)((jy * pz + jz) * px + jx); };
  if (nodes) {
#pragma omp parallel for schedule(static)
    for (int64_t jy = 0; jy < py; ++jy)
      for (int64_t jz = 0; jz < pz; ++jz)
        for (int64_t jx = 0; jx < px; ++jx) {
          double *x = nodes + 3 * (size_t)nid(jx, jy, jz);
          x[0] = lx * (double)jx / (double)(px - 1);
          x[1] = y0 + ly * (double)jy / (double)(py - 1);
          x[2] = lz * (double)jz / (double)(pz - 1);
        }
  }
  if (conn) {
    // the 6 axis orders; odd permutations get vertices 2,3 ... handled by a sign test below
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    static const int edge[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
#pragma omp parallel for schedule(static)
    for (int64_t cy = 0; cy < ny; ++cy)
      for (int64_t cz = 0; cz < nz; ++cz)
        for (int64_t cx = 0; cx < nx; ++cx) {
          const int64_t cube = (cy * nz + cz) * nx + cx;
          for (int t = 0; t < 6; ++t) {
            int64_t v[4][3];
            v[0][0] = 2 * cx; v[0][1] = 2 * cy; v[0][2] = 2 * cz;
            for (int s = 0; s < 3; ++s) {
              for (int d = 0; d < 3; ++d) v[s + 1][d] = v[s][d];
              v[s + 1][perm[t][s]] += 2;
            }
            // orientation: det[v1-v0; v2-v0; v3-v0] > 0 (reference J = dN.x, :690-696)
            int64_t a[3][3];
            for (int r = 0; r < 3; ++r)
              for (int d = 0; d < 3; ++d) a[r][d] = v[r + 1][d] - v[0][d];
            int64_t det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
                          a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                          a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
            if (det < 0)
              for (int d = 0; d < 3; ++d) std::swap(v[1][d], v[2][d]);
            int32_t *c = conn + (size_t)(cube * 6 + t) * 10;
            for (int k = 0; k < 4; ++k) c[k] = nid(v[k][0], v[k][1], v[k][2]);
            for (int k = 0; k < 6; ++k) {
              const int64_t *p0 = v[edge[k][0]], *p1 = v[edge[k][1]];
              c[4 + k] = nid((p0[0] + p1[0]) / 2, (p0[1] + p1[1]) / 2, (p0[2] + p1[2]) / 2);
            }
          }
        }
  }
  if (presc_node && presc_type && presc_vals) {
    int64_t k = 0;
    for (int side = 0; side < 2; ++side) {
      const int64_t jy = side ? py - 1 : 0;
      for (int64_t jz = 0; jz < pz; ++jz)
        for (int64_t jx = 0; jx < px; ++jx, ++k) {
          presc_node[k] = nid(jx, jy, jz);
          int type = bc_style == 1 ? 7 : 2;
          if (bc_style != 1 && side == 0 && jx == 0 && jz == 0) type = 7;  // pin one corner
          // style 2: also hold the next bottom corner along x in z -- removes the free rotation about
          // y that style 0 (the reference's "analytical" files) leaves in K, and is still satisfied by
          // the homogeneous uniaxial state
          if (bc_style == 2 && side == 0 && jx == px - 1 && jz == 0) type = 6;
          presc_type[k] = type;
          presc_vals[3 * k + 0] = 0.0;
          presc_vals[3 * k + 1] = side ? dy : 0.0;
          presc_vals[3 * k + 2] = 0.0;
        }
    }
  }
  return FEA_GPU_OK;
}

// Hollow cylinder: the Kuhn split of fea_mesh_block applied in (r, theta, z) index space, periodic
// in theta.  Every Kuhn edge joins vertex v and v + m with m a 0/1 vector, so a half-grid node with
// odd-coordinate mask m is the midpoint of the straight edge between (node - m) and (node + m).
extern "C" int fea_mesh_cylinder(int32_t nr, int32_t nt, int32_t nz, double r_in, double r_out, double length,
                                 double delta, int64_t *n_nodes, int64_t *n_elems, int64_t *n_presc,
                                 double *nodes, int32_t *conn, int32_t *presc_node, int32_t *presc_type,
                                 double *presc_vals) {
  if (nr < 1 || nt < 3 || nz < 1 || !(r_in > 0.0) || !(r_out > r_in) || !(length > 0.0)) return FEA_GPU_ERR_ARG;
  const int64_t pr = 2 * (int64_t)nr + 1, pt = 2 * (int64_t)nt, pz = 2 * (int64_t)nz + 1;
  const int64_t nn = pr * pt * pz, ne = 6 * (int64_t)nr * nt * nz;
  if (nn >= 0x7fffffff || ne >= 0x7fffffff) return FEA_GPU_ERR_ARG;
  auto nid = [&](int64_t ir, int64_t jt, int64_t kz) {
    return (int32_t)((kz * pt + ((jt % pt + pt) % pt)) * pr + ir);
  };
  const double two_pi = 6.283185307179586476925286766559;
  auto vertex = [&](int64_t ir, int64_t jt, int64_t kz, double *x) {   // all indices even
    const double r = r_in + (r_out - r_in) * (double)(ir / 2) / (double)nr;
    const double th = two_pi * (double)(((jt % pt + pt) % pt) / 2) / (double)nt;
    x[0] = r * std::cos(th);
    x[1] = r * std::sin(th);
    x[2] = length * (double)(kz / 2) / (double)nz;
  };
  // prescribed nodes: inner + outer wall (all theta, z), plus the interior of the two end faces
  int64_t np = 0;
  for (int64_t kz = 0; kz < pz; ++kz)
    for (int64_t jt = 0; jt < pt; ++jt)
      for (int64_t ir = 0; ir < pr; ++ir)
        if (ir == 0 || ir == pr - 1 || kz == 0 || kz == pz - 1) ++np;
  if (n_nodes) *n_nodes = nn;
  if (n_elems) *n_elems = ne;
  if (n_presc) *n_presc = np;
  if (nodes) {
#pragma omp parallel for schedule(static)
    for (int64_t kz = 0; kz < pz; ++kz)
      for (int64_t jt = 0; jt < pt; ++jt)
        for (int64_t ir = 0; ir < pr; ++ir) {
          const int64_t mr = ir & 1, mt = jt & 1, mz = kz & 1;
          double a[3], b[3];
          vertex(ir - mr, jt - mt, kz - mz, a);
          vertex(ir + mr, jt + mt, kz + mz, b);
          double *x = nodes + 3 * (size_t)nid(ir, jt, kz);
          for (int d = 0; d < 3; ++d) x[d] = 0.5 * (a[d] + b[d]);
        }
  }
  if (conn) {
    static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    static const int edge[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
#pragma omp parallel for schedule(static)
    for (int64_t cz = 0; cz < nz; ++cz)
      for (int64_t ct = 0; ct < nt; ++ct)
        for (int64_t cr = 0; cr < nr; ++cr) {
          const int64_t cell = (cz * nt + ct) * nr + cr;
          for (int t = 0; t < 6; ++t) {
            int64_t v[4][3];   // (ir, jt, kz) in doubled indices
            v[0][0] = 2 * cr; v[0][1] = 2 * ct; v[0][2] = 2 * cz;
            for (int s = 0; s < 3; ++s) {
              for (int d = 0; d < 3; ++d) v[s + 1][d] = v[s][d];
              v[s + 1][perm[t][s]] += 2;
            }
            // (r, theta, z) -> (x, y, z) preserves orientation, so the sign test works in index space
            int64_t a[3][3];
            for (int r = 0; r < 3; ++r)
              for (int d = 0; d < 3; ++d) a[r][d] = v[r + 1][d] - v[0][d];
            const int64_t det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
                                a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                                a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
            if (det < 0)
              for (int d = 0; d < 3; ++d) std::swap(v[1][d], v[2][d]);
            int32_t *c = conn + (size_t)(cell * 6 + t) * 10;
            for (int k = 0; k < 4; ++k) c[k] = nid(v[k][0], v[k][1], v[k][2]);
            for (int k = 0; k < 6; ++k) {
              const int64_t *p0 = v[edge[k][0]], *p1 = v[edge[k][1]];
              c[4 + k] = nid((p0[0] + p1[0]) / 2, (p0[1] + p1[1]) / 2, (p0[2] + p1[2]) / 2);
            }
          }
        }
  }
  if (nodes && presc_node && presc_type && presc_vals) {
    int64_t k = 0;
    for (int64_t kz = 0; kz < pz; ++kz)
      for (int64_t jt = 0; jt < pt; ++jt)
        for (int64_t ir = 0; ir < pr; ++ir) {
          const bool wall_in = ir == 0, wall_out = ir == pr - 1, face = kz == 0 || kz == pz - 1;
          if (!(wall_in || wall_out || face)) continue;
          const int32_t id = nid(ir, jt, kz);
          const double *x = nodes + 3 * (size_t)id;
          const double r = std::sqrt(x[0] * x[0] + x[1] * x[1]);
          presc_node[k] = id;
          presc_type[k] = ((wall_in || wall_out) ? 3 : 0) | (face ? 4 : 0);
          presc_vals[3 * k + 0] = wall_in ? delta * x[0] / r : 0.0;
          presc_vals[3 * k + 1] = wall_in ? delta * x[1] / r : 0.0;
          presc_vals[3 * k + 2] = 0.0;
          ++k;
        }
  }
  return FEA_GPU_OK;
}
