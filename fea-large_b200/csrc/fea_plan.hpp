// Host-side plan: node-range partition, local numbering, block-CSR pattern and the
// element -> nonzero gather map.  Pure C++ (no CUDA) so the N>1 logic is testable on a
// CPU box.  Replaces what libspmatrix did implicitly in sp_matrix_element_add
// (reference fea_solver.c:966,1055): the pattern is known before the first assembly.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fea {

constexpr int NEN = 10;                 // nodes per TETRAHEDRA10
constexpr int NTRI = 55;                // upper-triangular (a<=b) node pairs per element
constexpr uint32_t SRC_TRANSPOSE = 0x80000000u;

// K_e staging layout (what element_kernel writes and gather_blocks_kernel reads).  The 55 a<=b
// blocks of an element are grouped in five regions of 11: region pr = min(a, 9-a) holds row pr
// (10-pr blocks) and row 9-pr (pr+1 blocks) -- the work of one warp of the element kernel.  Inside
// a region the blocks come in consecutive pairs (a,b),(a,b+1) (first row's pairs, then the second
// row's: always five pairs) and the one block left over, (a,9) of the row with an odd count, last.
// code = 11 pr + pos; the block lives at double offset 500 e + 100 pr + 9 pos = 9 idx + idx / 11
// with idx = 55 e + code, so every pair starts on a 16-byte boundary (one pad double per region).
constexpr int KE_STRIDE = 500;          // doubles per element
// FEA_KE_INTERLEAVED: the 250 16-byte chunks of an element are interleaved across the 32 elements of an
// element-kernel CTA -- chunk c of element e at double offset ((e / 32) 250 + c) 64 + 2 (e % 32) -- so that
// the element kernel stores straight from registers (a warp store = 512 contiguous bytes) instead of
// transposing through shared memory; the gather then reads its five chunks at a 512-byte stride.
// Default since round 2: with the upper-triangle gather (every staged block read once) the strided reads cost the
// gather nothing (1.81 vs 1.84 ms) and the element kernel drops from 1.38 to 1.19 ms (profiles/r2_assembly_variants.md).
// -DFEA_KE_INTERLEAVED=0 builds the flat layout (Makefile: variant-flat).
#ifndef FEA_KE_INTERLEAVED
#define FEA_KE_INTERLEAVED 1
#endif
#if defined(__CUDACC__)
#define FEA_HD __host__ __device__
#else
#define FEA_HD
#endif
FEA_HD inline int ke_pos(int a, int b) {   // a <= b
  const int pr = a < NEN - 1 - a ? a : NEN - 1 - a;
  const int n0 = NEN - pr;                 // blocks of row pr
  const int k = b - a;
  if (a == pr) return k < (n0 & ~1) ? k : 10;
  const int n1 = pr + 1;                   // blocks of row 9-pr
  return k < (n1 & ~1) ? (n0 & ~1) + k : 10;
}
FEA_HD inline int ke_code(int a, int b) { return 11 * (a < NEN - 1 - a ? a : NEN - 1 - a) + ke_pos(a, b); }

// Direct (push) assembly: one cell per contribution to an upper slot, CELL_DOUBLES doubles (the 3x3 block and
// one pad double, so every cell starts on a 16-byte boundary); layout in fea_plan.cpp ("cell layout")
constexpr int CELL_DOUBLES = 10;
constexpr int CELL_MAX_CONTRIB = 2047;  // contributions per slot: low 11 bits of Plan::cmeta
constexpr int CELL_RANK_SHIFT = 11;     // high 5 bits: position of the slot inside every layer of its column
constexpr uint32_t CELL_NONE = 0xffffffffu;

constexpr int SELL_C = 32;              // rows per SELL slice = one warp
// Value layout of a 32-slot SELL column (288 doubles at 288 * column): components c = 3 i + j of lane l sit at
//   c 0-3: 4 l + c        c 4-7: 128 + 4 l + (c - 4)        c 8: 256 + l
// i.e. every lane owns two whole 32-byte sectors per column plus one double: the SpMV reads a column with two
// 256-bit loads and one 64-bit load per lane (1024 + 1024 + 256 contiguous bytes per warp), and a single slot can be
// written from anywhere (the transposed store into the lower triangle) without touching a sector another lane
// owns -- with one component per 256-byte row those stores were partial sectors (read-modify-write in DRAM).
FEA_HD inline int val_off(int c, int lane) { return c < 8 ? ((c >> 2) << 7) + (lane << 2) + (c & 3) : 256 + lane; }
constexpr int SELL_SIGMA = 2048;        // rows per length-sorting window

struct Plan {
  int rank = 0, nranks = 1;
  int64_t n_nodes_global = 0, n_elems_global = 0;

  // local numbering: owned nodes in Morton (Z-curve) order of their reference coordinates,
  // then ghosts grouped by owner, each group in that owner's own order.  Elements are local
  // in Morton order of their centroid.  (L2 locality for the gather and the x-gathers.)
  int32_t n_own = 0, n_local = 0, n_elems = 0;
  std::vector<int32_t> node_gid;        // [n_local]
  std::vector<int32_t> elem_gid;        // [n_elems] ascending
  std::vector<int32_t> conn;            // [n_elems][10] local node ids
  std::vector<uint8_t> elem_owned;      // [n_elems] 1 if this rank owns the element (owner of its node 0)
  std::vector<int32_t> owner;           // [n_nodes_global]
  std::vector<int32_t> pos_in_owner;    // [n_nodes_global] index of the node in its owner's numbering

  // block CSR over owned rows
  std::vector<int32_t> browptr;         // [n_own+1]
  std::vector<int32_t> bcol;            // [nnzb] local node ids ascending
  std::vector<int32_t> diag;            // [n_own] position of the diagonal block
  // stiffness gather map
  std::vector<int32_t> cptr;            // [nnzb+1]
  std::vector<uint32_t> csrc;           // [ncontrib]
  // SELL-32-sigma copy of the block pattern: what the device kernels use.  Rows are sorted by
  // length inside windows of SELL_SIGMA rows and cut into slices of 32; slot (row r = lane l of
  // slice s, j-th block of the row) = slice_ptr[s] + 32 j + l, its 9 values live at
  // 9 (slice_ptr[s] + 32 j) + val_off(c, l)  (c = 3 i + j').
  int32_t n_slices = 0;
  std::vector<int32_t> sell_row;        // [n_slices*32] local row of each lane, -1 = padding lane
  std::vector<int32_t> row_lane;        // [n_own] slice*32 + lane of each row
  std::vector<int32_t> slice_ptr;       // [n_slices+1] first slot of each slice
  std::vector<int32_t> sbcol;           // [n_slots] column node of each slot (padding: the row's own node)
  std::vector<int32_t> scptr;           // [n_slots+1] gather map in slot order
  std::vector<uint32_t> scsrc;          // [ncontrib]
  std::vector<int32_t> sdiag;           // [n_own] 9 * (first slot of the diagonal block's column) + lane
  int64_t n_slots() const { return (int64_t)sbcol.size(); }

  // direct (push) assembly, see "cell layout" in fea_plan.cpp
  std::vector<uint16_t> cmeta;          // [n_slots] upper slots: contributions | rank << 11; 0 = lower triangle / padding
  std::vector<int32_t> ccell;           // [n_slots/32 + 1] first cell of each 32-slot column
  std::vector<int32_t> cmirror;         // [n_slots] value index of the (0,0) entry of the mirror slot, -1 = none
  std::vector<uint32_t> edest;          // [55][ne_pad] cell of staged block (code, element) | SRC_TRANSPOSE; CELL_NONE = row not owned
  std::vector<int32_t> col_ready;       // [n_slots/32] last local element contributing to the column, -1 = no upper slot
  std::vector<int32_t> col_order;       // columns with upper slots, ascending col_ready
  int64_t n_cells() const { return ccell.empty() ? 0 : (int64_t)ccell.back(); }

  // residual gather map (node -> (element, local node))
  std::vector<int32_t> rptr;            // [n_own+1]
  std::vector<int32_t> rsrc;            // [.] elem*10 + a

  // halo
  std::vector<int32_t> nbr_rank;        // neighbour ranks ascending
  std::vector<int32_t> send_ptr;        // [nbr+1]
  std::vector<int32_t> send_nodes;      // local ids of owned nodes, per neighbour
  std::vector<int32_t> recv_ptr;        // [nbr+1] offsets into the ghost range

  int64_t nnzb() const { return (int64_t)bcol.size(); }
};

// throws std::runtime_error on bad meshes
void build_plan(Plan &p, int32_t n_nodes, int32_t n_elems, const double *X0,
                const int32_t *conn, int rank, int nranks);

}  // namespace fea
