// C-ABI implementation of include/fea_gpu.h: device context, kernel launches, NCCL halo
// exchange and scalar all-reduces.  One context = one GPU = one rank.  No CPU fallback:
// every numerical entry point launches the sm_100a kernels in element_kernels.cuh and
// sparse_kernels.cuh or fails with an error code.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include "../../include/fea_gpu.h"
#include "element_kernels.cuh"
#include "fea_plan.hpp"
#include "sparse_kernels.cuh"

namespace fea {
void plan_counts(const Plan &pl, int64_t out[16]);
}

using fea::PcgCtl;

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

extern "C" const char *fea_gpu_last_error(void) { return g_err.c_str(); }
extern "C" int64_t fea_gpu_launch_count(void) { return g_launches.load(); }

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
      return FEA_GPU_ERR_CUDA;                                                           \
    }                                                                                    \
  } while (0)
#define NC(call)                                                                         \
  do {                                                                                   \
    ncclResult_t r_ = (call);                                                            \
    if (r_ != ncclSuccess) {                                                             \
      g_err = std::string(#call) + ": " + ncclGetErrorString(r_);                        \
      return FEA_GPU_ERR_NCCL;                                                           \
    }                                                                                    \
  } while (0)
#define LAUNCHED()                                                                       \
  do {                                                                                   \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    CU(cudaGetLastError());                                                              \
  } while (0)
#define TRY(expr)                                                                        \
  do {                                                                                   \
    int rc_ = (expr);                                                                    \
    if (rc_ != FEA_GPU_OK) return rc_;                                                   \
  } while (0)

enum { PH_ELEM = 0, PH_GATHER_K, PH_GATHER_R, PH_BC, PH_PCG, PH_SPMV, PH_HALO, PH_COUNT };
constexpr int SPMV_EVENT_POOL = 128;
constexpr int PHASE_EVENT_POOL = 32;
constexpr int MAX_PARTIALS = 4096;
constexpr size_t FLUSH_BYTES = 512ull << 20;

struct fea_gpu_group;

struct fea_gpu_ctx {
  fea_gpu_group *group = nullptr;   // set when this context is one rank of a single-process multi-GPU handle
  fea::Plan plan;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // device->host copy of the residual, overlapped with the K gather
  cudaEvent_t ev_copy = nullptr;
  cudaStream_t comm_stream = nullptr;   // halo exchange of the CG direction, beside the interior SpMV
  cudaEvent_t ev_vec = nullptr, ev_halo = nullptr;
  ncclComm_t comm = nullptr;
  bool has_comm = false;
  int model = 0, ng = 5;
  double lambda = 0, mu = 0;
  int n_own = 0, n_local = 0, n_elems = 0, ne_pad = 0;
  int64_t nnzb = 0;
  int64_t n_slots = 0;
  int pcg_batch = 32;
  int pcg_stall = 0;               // 0 = automatic
  int pcg_variant = 0;             // 0 = classic PCG (two reductions per iteration, default), 1 = single reduction
                                   // (Chronopoulos-Gear: measured 1 % faster at N = 2 and NOT robust -- its alpha
                                   // denominator cancels when ||r|| peaks: breakdown at iteration 220 on a 16 k-DOF bar
                                   // that classic PCG solves in 1039, reproduced in numpy)
  int pcg_overlap = 0;             // 1 = halo exchange + boundary slices on a second stream beside the interior SpMV (measured slower at N = 2: 0.748 vs 0.684 ms per iteration, the cross-stream events cost more than the 19 us halo they hide)
  int last_exit = 0;               // exit state of the last solve: 0 max_iter, 1 tolerance, 2 stall/divergence guard
  int precond = 0;                 // 0 = Jacobi (north star), 1 = Chebyshev-accelerated Jacobi (the PCG_ILU request)
  int cheb_degree = 4;
  double cheb_ratio = 100.0;       // the polynomial targets [lmax / ratio, lmax] of D^-1 A
  double last_lmax = 0.0;
  double *cz = nullptr, *cd = nullptr;   // Chebyshev iterate z ([n_local][3]: it is an SpMV input) and increment d
  int gather_threads = 128;
  bool elem_ratio = true;          // A5: use the lambda/mu form of the block (set_param "elem_ratio" 0 = generic)
  int gather_split = 8;            // CTAs per slice (L2 footprint of the gather, sparse_kernels.cuh)
  // direct (push) assembly (gather_mode 2): cells of the upper slots written by the element kernel itself
  uint16_t *cmeta = nullptr;
  int32_t *ccell = nullptr, *cmirror = nullptr, *col_order = nullptr;
  uint32_t *edest = nullptr;
  double2 *cells = nullptr;
  int n_cols_active = 0;
  int cells_dbg = 0;               // diagnostics (tools/push_ab.py): 1 no mirror stores, 2 no own stores, 4 no cell stores in the element kernel
  int chunk_tiles = 0;             // 0 = one element launch + one gather launch; > 0 = chunks of that many 32-element tiles
  std::vector<int32_t> chunk_cols; // [chunks + 1] prefix of col_order each chunk may gather
  std::vector<int32_t> chunk_slices, slice_ready_sorted;   // the same for the pull gathers: prefix of slice_order
  int32_t *slice_order = nullptr;  // slices by the last element any of their upper slots waits for
  int chunk_overlap = 1;           // 1 = a chunk's gather runs on a second stream beside the next chunk's elements, 0 = one stream
  int chunk_cols_for = -1;         // chunk_tiles the list was built for
  cudaStream_t asm_stream = nullptr;   // the gathers of a chunked assembly run beside the next chunk's elements
  std::vector<cudaEvent_t> chunk_ev;
  struct AsmGraph { cudaGraphExec_t exec = nullptr; int launches = 0; };
  std::map<int64_t, AsmGraph> asm_graphs;  // captured chunk sequences, one per (residual, Dirichlet, chunk size) variant
  cudaEvent_t ev_asm = nullptr;
  int gather_sym = 1;              // pull gather: 1 = sum the upper triangle only and store each block into its mirror slot too, 0 = every slot sums its own list
  int gather_mode = 1;             // 1 = lane per slot (gather_blocks_kernel, default), 9 = nine lanes per block (gather_blocks9_kernel: 41 % fewer L1 sectors, same time -- DESIGN 4)
  bool gather9_ok = false;         // the uploaded lists satisfy what gather_blocks9_kernel assumes

  double *X0 = nullptr, *x = nullptr;
  double *x_saved = nullptr;       // fea_gpu_save_nodes (increment control: roll a trial step back)
  int32_t *conn_soa = nullptr;
  double *F_soa = nullptr, *S_soa = nullptr, *Ke = nullptr, *Re = nullptr;
  int32_t *slice_ptr = nullptr, *sell_row = nullptr, *bcol = nullptr, *cptr = nullptr, *rptr = nullptr,
          *rsrc = nullptr, *sdiag = nullptr;
  uint32_t *csrc = nullptr;
  double *vals = nullptr, *vals_saved = nullptr;
  double *R = nullptr, *u = nullptr, *p = nullptr, *q = nullptr, *r = nullptr, *dinv = nullptr, *u_saved = nullptr;
  double *pd = nullptr, *sv = nullptr;   // single-reduction PCG: direction p and s = A p (c->p then holds z, c->q holds w)
  fea::Pcg2State *st2 = nullptr;         // [2] double-buffered state
  fea::Pcg2State *st2_host = nullptr;
  int32_t *sl_inner = nullptr, *sl_bound = nullptr;   // SELL slices without / with ghost columns
  int n_inner = 0, n_bound = 0;
  uint8_t *pflag = nullptr;
  uint8_t *sflag = nullptr;        // [n_slots] bits 0-2 row DOFs prescribed, 3-5 column DOFs, 6 diagonal block
  double *pval = nullptr;
  int32_t *inc_dof = nullptr;
  double *inc_val = nullptr;
  int n_inc = 0;
  bool any_presc_value = false;
  int32_t *send_nodes = nullptr;
  double *send_buf = nullptr;
  double *partials = nullptr;
  double *partials_b = nullptr;    // grid-reduction scratch of the boundary SpMV (it runs beside the interior one)
  unsigned int *counters = nullptr;
  PcgCtl *ctl = nullptr;
  PcgCtl *ctl_host = nullptr;
  double *scalar = nullptr;        // device scratch scalar
  unsigned long long *bad = nullptr;
  void *flush = nullptr;
  double *export_buf = nullptr;    // lazily allocated [n_elems][ng][9]
  double *stage_h = nullptr;       // pinned staging for the fallback host-buffer path
  // host-buffer fast path: the local nodes live inside the global id range [io_lo, io_lo+io_cnt),
  // copied with one DMA and permuted on the device; owned nodes likewise for the way back
  int64_t io_lo = 0, io_cnt = 0, own_lo = 0, own_cnt = 0;
  bool io_range = false, own_range = false;
  int32_t *io_idx = nullptr, *own_idx = nullptr;   // local node -> offset inside the range
  double *io_buf = nullptr;

  // per phase: a ring of event pairs, one per call since the last fea_gpu_phase_ms (which reports the
  // average over them: the per-kernel durations of a whole timed region, no sync inside it)
  cudaEvent_t ev_a[PH_COUNT][PHASE_EVENT_POOL] = {}, ev_b[PH_COUNT][PHASE_EVENT_POOL] = {};
  int ev_n[PH_COUNT] = {};
  cudaEvent_t sp_a[SPMV_EVENT_POOL] = {}, sp_b[SPMV_EVENT_POOL] = {};
  int sp_used = 0;
  int last_iters = 0;
  cudaEvent_t tm_a = nullptr, tm_b = nullptr;
  std::vector<int32_t> own_count;  // owned nodes per rank
  std::vector<int32_t> elem_g2l;   // lazily built: global element id -> local element (-1 = not on this rank)
  double *ag_send = nullptr, *ag_recv = nullptr;   // all-gather buffers of the read-backs (allocated once)
  double *ag_host = nullptr;                       // pinned
};

// ---------------------------------------------------------------------------------
// Single-process multi-GPU handle (fea_gpu_create_multi): one context per GPU, one persistent host thread
// per context.  Every C-ABI call on the handle runs on all ranks at once -- the calls are the same
// rank-local calls a one-process-per-GPU launcher makes, NCCL included -- so the reference's one-process
// boundary (do_main -> solve, fea_solver.c:100-242) can drive the whole box.

struct fea_gpu_group {
  std::vector<fea_gpu_ctx *> ctx;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::function<int(int)> job;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> err;
};

static thread_local bool t_worker = false;   // true on the group's rank threads: calls go straight to the context

static void group_worker(fea_gpu_group *G, int i) {
  t_worker = true;
  uint64_t seen = 0;
  for (;;) {
    std::function<int(int)> job;
    {
      std::unique_lock<std::mutex> lk(G->mu);
      G->cv_job.wait(lk, [&] { return G->stop || G->generation != seen; });
      if (G->stop) return;
      seen = G->generation;
      job = G->job;
    }
    const int rc = job(i);
    {
      std::lock_guard<std::mutex> lk(G->mu);
      G->rc[(size_t)i] = rc;
      G->err[(size_t)i] = rc != FEA_GPU_OK ? g_err : std::string();
      if (--G->pending == 0) G->cv_done.notify_all();
    }
  }
}

// run f(rank) on every rank thread, wait for all; first non-zero return code wins
static int group_run(fea_gpu_group *G, std::function<int(int)> f) {
  std::unique_lock<std::mutex> lk(G->mu);
  G->job = std::move(f);
  G->pending = (int)G->threads.size();
  ++G->generation;
  G->cv_job.notify_all();
  G->cv_done.wait(lk, [&] { return G->pending == 0; });
  for (size_t i = 0; i < G->rc.size(); ++i)
    if (G->rc[i] != FEA_GPU_OK) {
      g_err = "rank " + std::to_string(i) + ": " + G->err[i];
      return G->rc[i];
    }
  return FEA_GPU_OK;
}

// first statement of every entry point that takes a handle: on a multi-GPU handle run `expr` (written in
// terms of the rank's context `ci` and its rank `gi`) on all rank threads
#define GROUP(c, ...)                                                                    \
  do {                                                                                   \
    if ((c) && (c)->group && !t_worker) {                                                \
      fea_gpu_group *G_ = (c)->group;                                                    \
      return group_run(G_, [&](int gi) -> int {                                          \
        fea_gpu_ctx *ci = G_->ctx[(size_t)gi];                                           \
        (void)ci;                                                                        \
        return __VA_ARGS__;                                                              \
      });                                                                                \
    }                                                                                    \
  } while (0)

template <class T>
static int dev_alloc(T **p, size_t n) {
  CU(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
  return FEA_GPU_OK;
}
template <class T>
static int dev_upload(T **p, const std::vector<T> &v, cudaStream_t s) {
  TRY(dev_alloc(p, v.size()));
  if (!v.empty()) CU(cudaMemcpyAsync(*p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s));
  return FEA_GPU_OK;
}

// Tet10 shape-function derivatives and the Gauss tables, stated as in the reference
// (fea_solver.c:32-54 and :1306-1361) so the constants are the same doubles.
static void host_tables(int ng, fea::ElemTables &t) {
  double gp[5][4];
  if (ng == 4) {
    const double a = 0.58541020, b = 0.13819660, w = (1 / 4.) / 6.;
    const double tab[4][4] = {{w, a, b, b}, {w, b, a, b}, {w, b, b, a}, {w, b, b, b}};
    std::memcpy(gp, tab, sizeof(tab));
  } else {
    const double wc = (-4 / 5.) / 6., w = (9 / 20.) / 6.;
    const double tab[5][4] = {{wc, 1 / 4., 1 / 4., 1 / 4.}, {w, 1 / 2., 1 / 6., 1 / 6.},
                              {w, 1 / 6., 1 / 2., 1 / 6.}, {w, 1 / 6., 1 / 6., 1 / 2.},
                              {w, 1 / 6., 1 / 6., 1 / 6.}};
    std::memcpy(gp, tab, sizeof(tab));
  }
  std::memset(&t, 0, sizeof(t));
  for (int g = 0; g < ng; ++g) {
    const double r = gp[g][1], s = gp[g][2], u = gp[g][3];
    t.w[g] = gp[g][0];
    const double d0 = 4 * u + 4 * s + 4 * r - 3;
    const double dr[10] = {d0, 4 * r - 1, 0, 0, -4 * u - 4 * s - 8 * r + 4, 4 * s, -4 * s, -4 * u, 4 * u, 0};
    const double ds[10] = {d0, 0, 4 * s - 1, 0, -4 * r, 4 * r, -4 * u - 8 * s - 4 * r + 4, -4 * u, 0, 4 * u};
    const double dt[10] = {d0, 0, 0, 4 * u - 1, -4 * r, 0, -4 * s, -8 * u - 4 * s - 4 * r + 4, 4 * r, 4 * s};
    for (int a = 0; a < 10; ++a) {
      t.dN[g][0][a] = dr[a];
      t.dN[g][1][a] = ds[a];
      t.dN[g][2][a] = dt[a];
    }
  }
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

static void phase_begin(fea_gpu_ctx *c, int ph) {
  cudaEventRecord(c->ev_a[ph][c->ev_n[ph] % PHASE_EVENT_POOL], c->stream);
}
static void phase_end(fea_gpu_ctx *c, int ph) {
  cudaEventRecord(c->ev_b[ph][c->ev_n[ph] % PHASE_EVENT_POOL], c->stream);
  c->ev_n[ph]++;
}

// ---------------------------------------------------------------------------------
// halo exchange of a [n_local][3] vector: owners push their interface entries to the
// ranks that hold them as ghosts.  Ghosts of one owner are contiguous, so
// ncclRecv lands in place.

static int halo_exchange(fea_gpu_ctx *c, double *vec, cudaStream_t st = nullptr) {
  const fea::Plan &pl = c->plan;
  if (!c->has_comm || pl.nbr_rank.empty()) return FEA_GPU_OK;
  if (!st) st = c->stream;
  const int n_send = (int)pl.send_nodes.size();
  if (n_send) {
    fea::pack_kernel<<<cdiv(3 * (int64_t)n_send, 256), 256, 0, st>>>(n_send, c->send_nodes, vec, c->send_buf);
    LAUNCHED();
  }
  NC(ncclGroupStart());
  for (size_t i = 0; i < pl.nbr_rank.size(); ++i) {
    const int ns = pl.send_ptr[i + 1] - pl.send_ptr[i], nr = pl.recv_ptr[i + 1] - pl.recv_ptr[i];
    if (ns) NC(ncclSend(c->send_buf + 3 * (size_t)pl.send_ptr[i], 3 * (size_t)ns, ncclDouble, pl.nbr_rank[i], c->comm, st));
    if (nr) NC(ncclRecv(vec + 3 * ((size_t)c->n_own + pl.recv_ptr[i]), 3 * (size_t)nr, ncclDouble, pl.nbr_rank[i], c->comm, st));
  }
  NC(ncclGroupEnd());
  return FEA_GPU_OK;
}

static int allreduce_sum(fea_gpu_ctx *c, double *dev, int count) {
  if (!c->has_comm) return FEA_GPU_OK;
  NC(ncclAllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, c->comm, c->stream));
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------

extern "C" int fea_gpu_nccl_unique_id(void *out128) {
  if (!out128) return FEA_GPU_ERR_ARG;
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(out128, &id, 128);
  return FEA_GPU_OK;
}

static int create_impl(fea_gpu_ctx *c, int32_t n_nodes, int32_t n_elems, const double *X0,
                       const int32_t *conn, int32_t n_presc, const int32_t *presc_node,
                       const int32_t *presc_type, const double *presc_vals, int rank, int nranks,
                       const void *nccl_id) {
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev <= 0 || c->device >= ndev) {
    g_err = "no CUDA device (this library has no CPU path)";
    return FEA_GPU_ERR_CUDA;
  }
  CU(cudaSetDevice(c->device));
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_vec, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));

  try {
    fea::build_plan(c->plan, n_nodes, n_elems, X0, conn, rank, nranks);
  } catch (const std::exception &e) {
    g_err = e.what();
    return FEA_GPU_ERR_MESH;
  }
  const fea::Plan &pl = c->plan;
  c->n_own = pl.n_own;
  c->n_local = pl.n_local;
  c->n_elems = pl.n_elems;
  c->ne_pad = (pl.n_elems + 31) / 32 * 32;
  c->nnzb = pl.nnzb();
  c->n_slots = pl.n_slots();
  // tuning knobs for A/B runs of the whole test suite (fea_gpu_set_param does the same per context)
  if (const char *s = getenv("FEA_GATHER_THREADS")) fea_gpu_set_param(c, "gather_threads", atof(s));
  if (const char *s = getenv("FEA_GATHER_SPLIT")) fea_gpu_set_param(c, "gather_split", atof(s));
  if (const char *s = getenv("FEA_GATHER_MODE")) fea_gpu_set_param(c, "gather_mode", atof(s));
  if (const char *s = getenv("FEA_CHUNK_TILES")) fea_gpu_set_param(c, "chunk_tiles", atof(s));
  if (const char *s = getenv("FEA_GATHER_SYM")) fea_gpu_set_param(c, "gather_sym", atof(s));
  if (const char *s = getenv("FEA_PCG_VARIANT")) fea_gpu_set_param(c, "pcg_variant", atof(s));
  if (const char *s = getenv("FEA_PCG_OVERLAP")) fea_gpu_set_param(c, "pcg_overlap", atof(s));
  if (const char *s = getenv("FEA_PCG_BATCH")) {
    int v = atoi(s);
    if (v >= 1 && v <= 4096) c->pcg_batch = v;
  }

  if (nranks > 1) {
    if (!nccl_id) {
      g_err = "nranks > 1 needs a ncclUniqueId";
      return FEA_GPU_ERR_ARG;
    }
    ncclUniqueId id;
    std::memcpy(&id, nccl_id, 128);
    NC(ncclCommInitRank(&c->comm, nranks, id, rank));
    c->has_comm = true;
  }
  c->own_count.assign((size_t)nranks, 0);
  for (int32_t i = 0; i < n_nodes; ++i) c->own_count[(size_t)pl.owner[(size_t)i]]++;

  // coordinates in local numbering
  {
    std::vector<double> xl(3 * (size_t)pl.n_local);
    for (int32_t l = 0; l < pl.n_local; ++l)
      for (int d = 0; d < 3; ++d) xl[3 * (size_t)l + d] = X0[3 * (size_t)pl.node_gid[(size_t)l] + d];
    TRY(dev_upload(&c->X0, xl, c->stream));
    TRY(dev_upload(&c->x, xl, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  {
    std::vector<int32_t> soa(10 * (size_t)c->ne_pad, 0);
    for (int32_t e = 0; e < pl.n_elems; ++e)
      for (int a = 0; a < 10; ++a) soa[(size_t)a * c->ne_pad + e] = pl.conn[(size_t)e * 10 + a];
    TRY(dev_upload(&c->conn_soa, soa, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  TRY(dev_upload(&c->slice_ptr, pl.slice_ptr, c->stream));
  TRY(dev_upload(&c->sell_row, pl.sell_row, c->stream));
  TRY(dev_upload(&c->bcol, pl.sbcol, c->stream));
  TRY(dev_upload(&c->cptr, pl.scptr, c->stream));
  {
    // device copy of the gather lists: bit 30 marks the last entry of every slot (gather_blocks9_kernel).
    // That kernel also needs the slots with entries to form a prefix of every 32-slot column (padding only
    // trails: rows are sorted by length inside a slice) and 30-bit block indices; checked here, once.
    std::vector<uint32_t> marked(pl.scsrc);
    bool ok = (int64_t)pl.n_elems * fea::NTRI < (int64_t)fea::SRC_LAST;
    const int64_t n_cols = pl.n_slots() / fea::SELL_C;
    for (int64_t col = 0; col < n_cols && ok; ++col) {
      bool ended = false;
      for (int l = 0; l < fea::SELL_C; ++l) {
        const int64_t slot = col * fea::SELL_C + l;
        const int32_t k0 = pl.scptr[(size_t)slot], k1 = pl.scptr[(size_t)slot + 1];
        if (k1 > k0) {
          if (ended) ok = false;
          marked[(size_t)k1 - 1] |= fea::SRC_LAST;
        } else {
          ended = true;
        }
      }
    }
    c->gather9_ok = ok;
    TRY(dev_upload(&c->csrc, ok ? marked : pl.scsrc, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  TRY(dev_upload(&c->rptr, pl.rptr, c->stream));
  TRY(dev_upload(&c->rsrc, pl.rsrc, c->stream));
  TRY(dev_upload(&c->sdiag, pl.sdiag, c->stream));
  TRY(dev_upload(&c->send_nodes, pl.send_nodes, c->stream));
  TRY(dev_alloc(&c->send_buf, 3 * pl.send_nodes.size()));
  {
    // slices whose rows reference no ghost column can be multiplied while the halo is still in flight
    std::vector<int32_t> inner, bound;
    for (int32_t sl = 0; sl < pl.n_slices; ++sl) {
      bool ghost = false;
      for (int32_t k = pl.slice_ptr[(size_t)sl]; k < pl.slice_ptr[(size_t)sl + 1] && !ghost; ++k)
        ghost = pl.sbcol[(size_t)k] >= pl.n_own;
      (ghost ? bound : inner).push_back(sl);
    }
    c->n_inner = (int)inner.size();
    c->n_bound = (int)bound.size();
    TRY(dev_upload(&c->sl_inner, inner, c->stream));
    TRY(dev_upload(&c->sl_bound, bound, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }

  // prescribed DOFs: flags for owned+ghost DOFs, summed increments (the reference adds
  // every list entry, fea_solver.c:1210-1240), last value for the RHS (:1256)
  {
    std::vector<int32_t> g2l((size_t)n_nodes, -1);
    for (int32_t l = 0; l < pl.n_local; ++l) g2l[(size_t)pl.node_gid[(size_t)l]] = l;
    std::vector<uint8_t> flag(3 * (size_t)pl.n_local, 0);
    std::vector<double> val(3 * (size_t)pl.n_local, 0.0), inc(3 * (size_t)pl.n_local, 0.0);
    for (int32_t k = 0; k < n_presc; ++k) {
      if (presc_node[k] < 0 || presc_node[k] >= n_nodes) {
        g_err = "prescribed node id out of range";
        return FEA_GPU_ERR_MESH;
      }
      const int32_t l = g2l[(size_t)presc_node[k]];
      if (l < 0) continue;
      const int type = presc_type[k];
      if (type < 0 || type > 7) continue;  // the reference matches only the 8 enum values
      for (int d = 0; d < 3; ++d)
        if (type & (1 << d)) {
          flag[3 * (size_t)l + d] = 1;
          val[3 * (size_t)l + d] = presc_vals[3 * (size_t)k + d];
          inc[3 * (size_t)l + d] += presc_vals[3 * (size_t)k + d];
          if (presc_vals[3 * (size_t)k + d] != 0.0) c->any_presc_value = true;
        }
    }
    std::vector<int32_t> idof;
    std::vector<double> ival;
    for (size_t t = 0; t < flag.size(); ++t)
      if (flag[t]) {
        idof.push_back((int32_t)t);
        ival.push_back(inc[t]);
      }
    c->n_inc = (int)idof.size();
    {
      std::vector<uint8_t> sf((size_t)pl.n_slots(), 0);
      for (int32_t sl = 0; sl < pl.n_slices; ++sl) {
        const int32_t base = pl.slice_ptr[(size_t)sl], width = (pl.slice_ptr[(size_t)sl + 1] - base) / fea::SELL_C;
        for (int l = 0; l < fea::SELL_C; ++l) {
          const int32_t row = pl.sell_row[(size_t)sl * fea::SELL_C + l];
          if (row < 0) continue;
          const unsigned rf = flag[3 * (size_t)row] | (flag[3 * (size_t)row + 1] << 1) | (flag[3 * (size_t)row + 2] << 2);
          for (int32_t j = 0; j < width; ++j) {
            const size_t slot = (size_t)base + (size_t)j * fea::SELL_C + l;
            const int32_t col = pl.sbcol[slot];
            const unsigned cf = flag[3 * (size_t)col] | (flag[3 * (size_t)col + 1] << 1) | (flag[3 * (size_t)col + 2] << 2);
            sf[slot] = (uint8_t)(rf | (cf << 3) | (col == row ? 64u : 0u));
          }
        }
      }
      TRY(dev_upload(&c->sflag, sf, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    TRY(dev_upload(&c->pflag, flag, c->stream));
    TRY(dev_upload(&c->pval, val, c->stream));
    TRY(dev_upload(&c->inc_dof, idof, c->stream));
    TRY(dev_upload(&c->inc_val, ival, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }

  {
    // ranges of global ids covering the local / the owned nodes (slab partitions of meshes numbered
    // along the slab axis give tight ranges; otherwise the host-side gather is used)
    int32_t lo = INT32_MAX, hi = -1, olo = INT32_MAX, ohi = -1;
    for (int32_t l = 0; l < pl.n_local; ++l) {
      const int32_t g = pl.node_gid[(size_t)l];
      lo = std::min(lo, g);
      hi = std::max(hi, g);
      if (l < pl.n_own) {
        olo = std::min(olo, g);
        ohi = std::max(ohi, g);
      }
    }
    c->io_lo = lo;
    c->io_cnt = (int64_t)hi - lo + 1;
    c->own_lo = olo;
    c->own_cnt = (int64_t)ohi - olo + 1;
    c->io_range = c->io_cnt <= 2 * (int64_t)pl.n_local;
    c->own_range = c->own_cnt == (int64_t)pl.n_own;          // must be exact: the copy back overwrites the range
    if (c->io_range) {
      std::vector<int32_t> idx((size_t)pl.n_local), oidx((size_t)pl.n_own);
      for (int32_t l = 0; l < pl.n_local; ++l) idx[(size_t)l] = pl.node_gid[(size_t)l] - lo;
      for (int32_t l = 0; l < pl.n_own; ++l) oidx[(size_t)l] = pl.node_gid[(size_t)l] - olo;
      TRY(dev_upload(&c->io_idx, idx, c->stream));
      TRY(dev_upload(&c->own_idx, oidx, c->stream));
      TRY(dev_alloc(&c->io_buf, 3 * (size_t)c->io_cnt));
      CU(cudaStreamSynchronize(c->stream));
    }
  }
  const size_t n3 = 3 * (size_t)c->n_own, nl3 = 3 * (size_t)c->n_local;
  TRY(dev_alloc(&c->F_soa, (size_t)c->ng * 9 * c->ne_pad));
  TRY(dev_alloc(&c->S_soa, (size_t)c->ng * 9 * c->ne_pad));
  // the K_e staging (pull gathers) or the cells (direct assembly) are allocated on first use: ensure_ke / ensure_cells
  TRY(dev_alloc(&c->Re, (size_t)30 * c->ne_pad));
  TRY(dev_alloc(&c->vals, (size_t)c->n_slots * 9));
  TRY(dev_alloc(&c->R, n3));
  TRY(dev_alloc(&c->u, nl3));
  TRY(dev_alloc(&c->p, nl3));
  TRY(dev_alloc(&c->q, n3));
  TRY(dev_alloc(&c->r, n3));
  TRY(dev_alloc(&c->dinv, n3));
  TRY(dev_alloc(&c->u_saved, n3));
  TRY(dev_alloc(&c->partials, 3 * (size_t)MAX_PARTIALS));
  TRY(dev_alloc(&c->partials_b, (size_t)MAX_PARTIALS));
  TRY(dev_alloc(&c->counters, 8));
  TRY(dev_alloc(&c->ctl, 1));
  TRY(dev_alloc(&c->scalar, 4));
  TRY(dev_alloc(&c->bad, 1));
  CU(cudaHostAlloc((void **)&c->ctl_host, sizeof(PcgCtl), cudaHostAllocDefault));
  std::memset(c->ctl_host, 0, sizeof(PcgCtl));
  TRY(dev_alloc(&c->st2, 2));
  CU(cudaHostAlloc((void **)&c->st2_host, sizeof(fea::Pcg2State), cudaHostAllocDefault));
  std::memset(c->st2_host, 0, sizeof(fea::Pcg2State));
  CU(cudaMemsetAsync(c->F_soa, 0, sizeof(double) * (size_t)c->ng * 9 * c->ne_pad, c->stream));
  CU(cudaMemsetAsync(c->S_soa, 0, sizeof(double) * (size_t)c->ng * 9 * c->ne_pad, c->stream));
  CU(cudaMemsetAsync(c->vals, 0, sizeof(double) * (size_t)c->n_slots * 9, c->stream));
  CU(cudaMemsetAsync(c->R, 0, sizeof(double) * n3, c->stream));
  CU(cudaMemsetAsync(c->u, 0, sizeof(double) * nl3, c->stream));
  CU(cudaMemsetAsync(c->p, 0, sizeof(double) * nl3, c->stream));
  CU(cudaMemsetAsync(c->counters, 0, sizeof(unsigned int) * 8, c->stream));
  CU(cudaMemsetAsync(c->ctl, 0, sizeof(PcgCtl), c->stream));
  CU(cudaMemsetAsync(c->bad, 0, sizeof(unsigned long long), c->stream));

  // both quadrature rules, unconditionally: another live context on this device may use the other one
  for (int k = 0; k < 2; ++k) {
    fea::ElemTables tab;
    host_tables(k ? 5 : 4, tab);
    CU(cudaMemcpyToSymbolAsync(fea::c_tabs, &tab, sizeof(tab), sizeof(tab) * k, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // `tab` is pageable stack memory
  }

  for (int i = 0; i < PH_COUNT; ++i) {
    for (int k = 0; k < PHASE_EVENT_POOL; ++k) {
      CU(cudaEventCreate(&c->ev_a[i][k]));
      CU(cudaEventCreate(&c->ev_b[i][k]));
    }
    c->ev_n[i] = 0;
  }
  for (int i = 0; i < SPMV_EVENT_POOL; ++i) {
    CU(cudaEventCreate(&c->sp_a[i]));
    CU(cudaEventCreate(&c->sp_b[i]));
  }
  CU(cudaEventCreate(&c->tm_a));
  CU(cudaEventCreate(&c->tm_b));
  CU(cudaStreamSynchronize(c->stream));
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_create(fea_gpu_handle *out, int32_t n_nodes, int32_t n_elems, const double *X0,
                              const int32_t *conn, int32_t model_type, double lambda, double mu,
                              int32_t n_gauss, int32_t n_presc, const int32_t *presc_node,
                              const int32_t *presc_type, const double *presc_vals, int32_t rank,
                              int32_t nranks, const void *nccl_unique_id, int32_t device) {
  if (!out || !X0 || !conn || n_nodes <= 0 || n_elems <= 0) {
    g_err = "null or empty mesh";
    return FEA_GPU_ERR_ARG;
  }
  if (model_type != FEA_MODEL_A5 && model_type != FEA_MODEL_COMPRESSIBLE_NEOHOOKEAN) {
    g_err = "unknown model_type";
    return FEA_GPU_ERR_ARG;
  }
  if (n_gauss != 4 && n_gauss != 5) {  // fea_solver.c:1495-1504
    g_err = "gauss nodes count must be 4 or 5";
    return FEA_GPU_ERR_ARG;
  }
  if (n_presc < 0 || (n_presc > 0 && (!presc_node || !presc_type || !presc_vals))) {
    g_err = "bad prescribed-displacement arrays";
    return FEA_GPU_ERR_ARG;
  }
  fea_gpu_ctx *c = new fea_gpu_ctx();
  c->device = device;
  c->model = model_type;
  c->ng = n_gauss;
  c->lambda = lambda;
  c->mu = mu;
  int rc = create_impl(c, n_nodes, n_elems, X0, conn, n_presc, presc_node, presc_type, presc_vals,
                       rank, nranks, nccl_unique_id);
  if (rc != FEA_GPU_OK) {
    fea_gpu_destroy(c);
    return rc;
  }
  *out = c;
  return FEA_GPU_OK;
}

static void group_stop(fea_gpu_group *G) {
  {
    std::lock_guard<std::mutex> lk(G->mu);
    G->stop = true;
  }
  G->cv_job.notify_all();
  for (std::thread &t : G->threads) t.join();
  delete G;
}

// One process, n_gpus GPUs: the same arguments as fea_gpu_create without the rank plumbing.  The handle
// that comes back stands for all ranks; every other entry point accepts it.
extern "C" int fea_gpu_create_multi(fea_gpu_handle *out, int32_t n_nodes, int32_t n_elems, const double *X0,
                                    const int32_t *conn, int32_t model_type, double lambda, double mu,
                                    int32_t n_gauss, int32_t n_presc, const int32_t *presc_node,
                                    const int32_t *presc_type, const double *presc_vals, int32_t n_gpus,
                                    const int32_t *devices) {
  if (!out || n_gpus < 1) return FEA_GPU_ERR_ARG;
  if (n_gpus == 1)
    return fea_gpu_create(out, n_nodes, n_elems, X0, conn, model_type, lambda, mu, n_gauss, n_presc, presc_node,
                          presc_type, presc_vals, 0, 1, nullptr, devices ? devices[0] : 0);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev < n_gpus) {
    g_err = "fea_gpu_create_multi: " + std::to_string(n_gpus) + " GPUs asked, " + std::to_string(ndev) + " visible";
    return FEA_GPU_ERR_CUDA;
  }
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  fea_gpu_group *G = new fea_gpu_group();
  G->ctx.assign((size_t)n_gpus, nullptr);
  G->rc.assign((size_t)n_gpus, 0);
  G->err.assign((size_t)n_gpus, std::string());
  for (int i = 0; i < n_gpus; ++i) G->threads.emplace_back(group_worker, G, i);
  const int rc = group_run(G, [&](int gi) -> int {
    return fea_gpu_create(&G->ctx[(size_t)gi], n_nodes, n_elems, X0, conn, model_type, lambda, mu, n_gauss, n_presc,
                          presc_node, presc_type, presc_vals, gi, n_gpus, &id, devices ? devices[gi] : gi);
  });
  if (rc != FEA_GPU_OK) {
    const std::string why = g_err;
    group_run(G, [&](int gi) -> int { return G->ctx[(size_t)gi] ? fea_gpu_destroy(G->ctx[(size_t)gi]) : FEA_GPU_OK; });
    group_stop(G);
    g_err = why;
    return rc;
  }
  for (fea_gpu_ctx *ci : G->ctx) ci->group = G;
  *out = G->ctx[0];
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_destroy(fea_gpu_handle c) {
  if (!c) return FEA_GPU_OK;
  if (c->group && !t_worker) {
    fea_gpu_group *G = c->group;
    group_run(G, [&](int gi) -> int { return fea_gpu_destroy(G->ctx[(size_t)gi]); });
    group_stop(G);
    return FEA_GPU_OK;
  }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->has_comm) ncclCommDestroy(c->comm);
  void *ptrs[] = {c->X0, c->x, c->conn_soa, c->F_soa, c->S_soa, c->Ke, c->Re, c->slice_ptr, c->sell_row, c->bcol,
                  c->cptr, c->rptr, c->rsrc, c->sdiag, c->csrc, c->vals, c->vals_saved, c->R, c->u,
                  c->p, c->q, c->r, c->dinv, c->u_saved, c->pflag, c->sflag, c->pval, c->inc_dof, c->inc_val,
                  c->send_nodes, c->send_buf, c->io_idx, c->own_idx, c->io_buf, c->partials, c->partials_b, c->counters, c->ctl, c->scalar, c->bad,
                  c->flush, c->export_buf, c->x_saved, c->ag_send, c->ag_recv, c->cz, c->cd, c->pd, c->sv, c->st2, c->sl_inner, c->sl_bound,
                  c->cmeta, c->ccell, c->cmirror, c->col_order, c->edest, c->cells, c->slice_order};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (c->ctl_host) cudaFreeHost(c->ctl_host);
  if (c->st2_host) cudaFreeHost(c->st2_host);
  if (c->ag_host) cudaFreeHost(c->ag_host);
  if (c->ev_vec) cudaEventDestroy(c->ev_vec);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->stage_h) cudaFreeHost(c->stage_h);
  for (int i = 0; i < PH_COUNT; ++i) {
    for (int k = 0; k < PHASE_EVENT_POOL; ++k) {
      if (c->ev_a[i][k]) cudaEventDestroy(c->ev_a[i][k]);
      if (c->ev_b[i][k]) cudaEventDestroy(c->ev_b[i][k]);
    }
  }
  for (int i = 0; i < SPMV_EVENT_POOL; ++i) {
    if (c->sp_a[i]) cudaEventDestroy(c->sp_a[i]);
    if (c->sp_b[i]) cudaEventDestroy(c->sp_b[i]);
  }
  if (c->tm_a) cudaEventDestroy(c->tm_a);
  if (c->tm_b) cudaEventDestroy(c->tm_b);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  for (cudaEvent_t ev : c->chunk_ev) cudaEventDestroy(ev);
  for (auto &kv : c->asm_graphs) cudaGraphExecDestroy(kv.second.exec);
  if (c->ev_asm) cudaEventDestroy(c->ev_asm);
  if (c->asm_stream) cudaStreamDestroy(c->asm_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return FEA_GPU_OK;
}

#define CHECK_H(h)                                                                       \
  do {                                                                                   \
    if (!(h)) {                                                                          \
      g_err = "null handle";                                                             \
      return FEA_GPU_ERR_ARG;                                                            \
    }                                                                                    \
    CU(cudaSetDevice((h)->device));                                                      \
  } while (0)

// ---------------------------------------------------------------------------------
// vectors between the caller's global numbering and the rank-local device layout

// host_global == nullptr: take part in the collective, place nothing (ranks > 0 of a multi-GPU handle)
static int gather_owned(fea_gpu_ctx *c, const double *dev_vec, double *host_global) {
  const fea::Plan &pl = c->plan;
  if (!c->has_comm) {
    if (!host_global) return FEA_GPU_OK;
    std::vector<double> tmp(3 * (size_t)c->n_own);
    CU(cudaMemcpyAsync(tmp.data(), dev_vec, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int32_t l = 0; l < c->n_own; ++l)
      std::memcpy(host_global + 3 * (size_t)pl.node_gid[(size_t)l], tmp.data() + 3 * (size_t)l, 3 * sizeof(double));
    return FEA_GPU_OK;
  }
  // all-gather of equally padded owned blocks, then placement by the shared owner map.  The buffers are
  // allocated once: cudaMalloc / cudaFree around a collective would serialise the ranks of a one-process box
  int32_t maxown = 0;
  for (int32_t v : c->own_count) maxown = std::max(maxown, v);
  const size_t blk = 3 * (size_t)maxown;
  if (!c->ag_send) {
    TRY(dev_alloc(&c->ag_send, blk));
    TRY(dev_alloc(&c->ag_recv, blk * (size_t)pl.nranks));
    CU(cudaHostAlloc((void **)&c->ag_host, sizeof(double) * blk * (size_t)pl.nranks, cudaHostAllocDefault));
  }
  CU(cudaMemsetAsync(c->ag_send, 0, sizeof(double) * blk, c->stream));
  CU(cudaMemcpyAsync(c->ag_send, dev_vec, sizeof(double) * 3 * (size_t)c->n_own, cudaMemcpyDeviceToDevice, c->stream));
  NC(ncclAllGather(c->ag_send, c->ag_recv, blk, ncclDouble, c->comm, c->stream));
  if (host_global)
    CU(cudaMemcpyAsync(c->ag_host, c->ag_recv, sizeof(double) * blk * (size_t)pl.nranks, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (!host_global) return FEA_GPU_OK;
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < pl.n_nodes_global; ++g) {  // every rank knows every owner's numbering
    const int o = pl.owner[(size_t)g];
    std::memcpy(host_global + 3 * (size_t)g, c->ag_host + blk * (size_t)o + 3 * (size_t)pl.pos_in_owner[(size_t)g],
                3 * sizeof(double));
  }
  return FEA_GPU_OK;
}

static int scatter_local(fea_gpu_ctx *c, const double *host_global, double *dev_vec, int n_nodes_local) {
  const fea::Plan &pl = c->plan;
  std::vector<double> tmp(3 * (size_t)n_nodes_local);
  for (int32_t l = 0; l < n_nodes_local; ++l)
    std::memcpy(tmp.data() + 3 * (size_t)l, host_global + 3 * (size_t)pl.node_gid[(size_t)l], 3 * sizeof(double));
  CU(cudaMemcpyAsync(dev_vec, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_set_nodes(fea_gpu_handle c, const double *x) {
  GROUP(c, fea_gpu_set_nodes(ci, x));
  CHECK_H(c);
  if (!x) return FEA_GPU_ERR_ARG;
  return scatter_local(c, x, c->x, c->n_local);
}
extern "C" int fea_gpu_get_nodes(fea_gpu_handle c, double *x) {
  GROUP(c, fea_gpu_get_nodes(ci, gi ? nullptr : x));
  CHECK_H(c);
  if (!x && !(c->group && t_worker)) return FEA_GPU_ERR_ARG;
  return gather_owned(c, c->x, x);
}
extern "C" int fea_gpu_get_forces(fea_gpu_handle c, double *R) {
  GROUP(c, fea_gpu_get_forces(ci, gi ? nullptr : R));
  CHECK_H(c);
  if (!R && !(c->group && t_worker)) return FEA_GPU_ERR_ARG;
  return gather_owned(c, c->R, R);
}
extern "C" int fea_gpu_set_forces(fea_gpu_handle c, const double *R) {
  GROUP(c, fea_gpu_set_forces(ci, R));
  CHECK_H(c);
  if (!R) return FEA_GPU_ERR_ARG;
  return scatter_local(c, R, c->R, c->n_own);
}
extern "C" int fea_gpu_get_solution(fea_gpu_handle c, double *u) {
  GROUP(c, fea_gpu_get_solution(ci, gi ? nullptr : u));
  CHECK_H(c);
  if (!u && !(c->group && t_worker)) return FEA_GPU_ERR_ARG;
  return gather_owned(c, c->u, u);
}

extern "C" int fea_gpu_apply_increment(fea_gpu_handle c, double lambda) {
  GROUP(c, fea_gpu_apply_increment(ci, lambda));
  CHECK_H(c);
  if (c->n_inc) {
    fea::increment_kernel<<<cdiv(c->n_inc, 256), 256, 0, c->stream>>>(c->n_inc, c->inc_dof, c->inc_val, lambda, c->x);
    LAUNCHED();
  }
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_update_nodes_scaled(fea_gpu_handle c, double eta) {
  GROUP(c, fea_gpu_update_nodes_scaled(ci, eta));
  CHECK_H(c);
  const int n = 3 * c->n_own;
  fea::axpy_kernel<<<std::min(cdiv(n, 256), 148 * 8), 256, 0, c->stream>>>(n, eta, c->u, c->x);
  LAUNCHED();
  phase_begin(c, PH_HALO);
  TRY(halo_exchange(c, c->x));
  phase_end(c, PH_HALO);
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_update_nodes(fea_gpu_handle c) { return fea_gpu_update_nodes_scaled(c, 1.0); }

extern "C" int fea_gpu_save_nodes(fea_gpu_handle c) {
  GROUP(c, fea_gpu_save_nodes(ci));
  CHECK_H(c);
  const size_t nl3 = 3 * (size_t)c->n_local;
  if (!c->x_saved) TRY(dev_alloc(&c->x_saved, nl3));
  CU(cudaMemcpyAsync(c->x_saved, c->x, sizeof(double) * nl3, cudaMemcpyDeviceToDevice, c->stream));
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_extrapolate_nodes(fea_gpu_handle c, double alpha) {
  GROUP(c, fea_gpu_extrapolate_nodes(ci, alpha));
  CHECK_H(c);
  if (!c->x_saved) {
    g_err = "no saved nodes";
    return FEA_GPU_ERR_ARG;
  }
  const int n = 3 * c->n_local;     // ghosts included: both arrays hold consistent ghost values
  fea::extrapolate_kernel<<<std::min(cdiv(n, 256), 148 * 8), 256, 0, c->stream>>>(n, alpha, c->x, c->x_saved);
  LAUNCHED();
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_restore_nodes(fea_gpu_handle c) {
  GROUP(c, fea_gpu_restore_nodes(ci));
  CHECK_H(c);
  if (!c->x_saved) {
    g_err = "no saved nodes";
    return FEA_GPU_ERR_ARG;
  }
  CU(cudaMemcpyAsync(c->x, c->x_saved, sizeof(double) * 3 * (size_t)c->n_local, cudaMemcpyDeviceToDevice, c->stream));
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------
// element pass

template <int MODEL, int NG, bool RATIO>
static int launch_element(fea_gpu_ctx *c, bool with_k, bool with_r, const fea::ElemArgs &args, int n_tiles) {
  const int grid = n_tiles;
  const bool push = with_k && args.cells != nullptr;
  const size_t smem = sizeof(double) * NG * fea::FLD_DOUBLES +
                      sizeof(double2) * NG * (push ? fea::PUSH_TILE_D2 : fea::TILE_D2) + sizeof(int) * 9 * 32 +
                      (push ? sizeof(uint32_t) * fea::NTRI * 32 : 0);
#define FEA_LAUNCH(K, Rr, P)                                                                       \
  do {                                                                                             \
    auto kern = fea::element_kernel<MODEL, NG, K, Rr, RATIO, P>;                                   \
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kern<<<grid, NG * 32, smem, c->stream>>>(args);                                                \
  } while (0)
  if (push && with_r) FEA_LAUNCH(true, true, true);
  else if (push) FEA_LAUNCH(true, false, true);
  else if (with_k && with_r) FEA_LAUNCH(true, true, false);
  else if (with_k) FEA_LAUNCH(true, false, false);
  else if (with_r) FEA_LAUNCH(false, true, false);
  else FEA_LAUNCH(false, false, false);
#undef FEA_LAUNCH
  LAUNCHED();
  return FEA_GPU_OK;
}

// device side of the cell layout, on first use of gather_mode 2
static int ensure_cells(fea_gpu_ctx *c) {
  if (c->cells) return FEA_GPU_OK;
  const fea::Plan &pl = c->plan;
  TRY(dev_upload(&c->cmeta, pl.cmeta, c->stream));
  TRY(dev_upload(&c->ccell, pl.ccell, c->stream));
  if (!c->cmirror) TRY(dev_upload(&c->cmirror, pl.cmirror, c->stream));
  TRY(dev_upload(&c->col_order, pl.col_order, c->stream));
  TRY(dev_upload(&c->edest, pl.edest, c->stream));
  c->n_cols_active = (int)pl.col_order.size();
  TRY(dev_alloc(&c->cells, (size_t)std::max<int64_t>(pl.n_cells(), 1) * 5));
  CU(cudaStreamSynchronize(c->stream));
  return FEA_GPU_OK;
}
static int ensure_ke(fea_gpu_ctx *c) {
  if (c->Ke) return FEA_GPU_OK;
  return dev_alloc(&c->Ke, (size_t)c->ne_pad * fea::KE_STRIDE);   // ne_pad: whole CTAs of 32 elements
}

// prefix of col_order every chunk of `chunk_tiles` element tiles may gather: the columns whose last
// contributing element lies before the end of the chunk
static void build_chunks(fea_gpu_ctx *c) {
  if (!c->asm_stream) {
    cudaStreamCreateWithFlags(&c->asm_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->ev_asm, cudaEventDisableTiming);
  }
  if (c->chunk_cols_for == c->chunk_tiles) return;
  const fea::Plan &pl = c->plan;
  const int n_tiles = cdiv(c->n_elems, fea::ELEMS_PER_CTA);
  const int n_chunks = cdiv(n_tiles, c->chunk_tiles);
  c->chunk_cols.assign((size_t)n_chunks + 1, 0);
  size_t k = 0;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int64_t e_end = std::min<int64_t>((int64_t)(ch + 1) * c->chunk_tiles * fea::ELEMS_PER_CTA, c->n_elems);
    while (k < pl.col_order.size() && pl.col_ready[(size_t)pl.col_order[k]] < e_end) ++k;
    c->chunk_cols[(size_t)ch + 1] = (int32_t)k;
  }
  // pull gathers work slice by slice: a slice is ready when the last of its columns is
  if (!c->slice_order) {
    std::vector<int32_t> ready((size_t)pl.n_slices, -1), order((size_t)pl.n_slices);
    for (int32_t sl = 0; sl < pl.n_slices; ++sl)
      for (int32_t col = pl.slice_ptr[(size_t)sl] / fea::SELL_C; col < pl.slice_ptr[(size_t)sl + 1] / fea::SELL_C; ++col)
        ready[(size_t)sl] = std::max(ready[(size_t)sl], pl.col_ready[(size_t)col]);
    for (int32_t sl = 0; sl < pl.n_slices; ++sl) order[(size_t)sl] = sl;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ready[(size_t)a] < ready[(size_t)b]; });
    c->slice_ready_sorted.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) c->slice_ready_sorted[i] = ready[(size_t)order[i]];
    if (dev_upload(&c->slice_order, order, c->stream) == FEA_GPU_OK) cudaStreamSynchronize(c->stream);
  }
  c->chunk_slices.assign((size_t)n_chunks + 1, 0);
  k = 0;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int64_t e_end = ch + 1 == n_chunks ? INT64_MAX : (int64_t)(ch + 1) * c->chunk_tiles * fea::ELEMS_PER_CTA;
    while (k < c->slice_ready_sorted.size() && c->slice_ready_sorted[k] < e_end) ++k;
    c->chunk_slices[(size_t)ch + 1] = (int32_t)k;
  }
  while ((int)c->chunk_ev.size() < n_chunks) {
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    c->chunk_ev.push_back(ev);
  }
  c->chunk_cols_for = c->chunk_tiles;
}

static int launch_gather_cells(fea_gpu_ctx *c, int col0, int col1, bool with_bc, cudaStream_t st) {
  if (col1 <= col0) return FEA_GPU_OK;
  constexpr int W = 4;
  fea::gather_cells_kernel<W><<<cdiv(col1 - col0, W), W * 32, 0, st>>>(
      col1 - col0, c->col_order + col0, c->ccell, c->cmeta, c->cmirror, c->cells, c->vals, with_bc ? c->sflag : nullptr, c->cells_dbg);
  LAUNCHED();
  return FEA_GPU_OK;
}

static int element_launch(fea_gpu_ctx *c, bool with_k, bool with_r, fea::ElemArgs &a, int tile0, int n_tiles);
static int launch_gather_pull(fea_gpu_ctx *c, bool with_bc, const int32_t *list, int n, cudaStream_t st);
static int ensure_mirror(fea_gpu_ctx *c);
static fea::SellMat sell_mat(fea_gpu_ctx *c);

static int element_pass(fea_gpu_ctx *c, bool with_k, bool with_r, bool chunk_bc = false, bool *gathered = nullptr) {
  fea::ElemArgs a;
  a.n_elems = c->n_elems;
  a.ne_pad = c->ne_pad;
  a.conn_soa = c->conn_soa;
  a.X0 = c->X0;
  a.x = c->x;
  a.lambda = c->lambda;
  a.mu = c->mu;
  a.F_soa = c->F_soa;
  a.S_soa = c->S_soa;
  const bool push = with_k && c->gather_mode == 2;
  if (push) TRY(ensure_cells(c));
  else if (with_k) TRY(ensure_ke(c));
  a.Ke = c->Ke;
  a.edest = push ? c->edest : nullptr;
  a.cells = push ? c->cells : nullptr;
  a.dbg = c->cells_dbg;
  a.tile0 = 0;
  a.Re = c->Re;
  a.bad = c->bad;
  CU(cudaMemsetAsync(c->bad, 0, sizeof(unsigned long long), c->stream));
  const int n_tiles = cdiv(c->n_elems, fea::ELEMS_PER_CTA);
  if (gathered) *gathered = false;
  const bool pull_chunks = with_k && c->gather_mode == 1 && c->gather_sym;
  if (pull_chunks) TRY(ensure_mirror(c));
  if ((push || pull_chunks) && gathered && c->chunk_tiles > 0 && c->chunk_tiles < n_tiles) {
    // Chunked assembly: the element kernel runs chunk by chunk on the context's stream, and the columns that a
    // chunk completes are gathered on a second stream beside the next chunk's elements -- their cells were
    // written microseconds ago and are read back from L2 instead of HBM.
    build_chunks(c);
    phase_begin(c, PH_ELEM);
    // The ~2 x chunks launches and their cross-stream dependencies are captured once per variant into a CUDA
    // graph: issued one by one from the host they cost more than the kernels take.
    // everything a captured launch sequence depends on besides the (fixed) device pointers and material constants
    const int64_t key = (with_r ? 1 : 0) | (chunk_bc ? 2 : 0) | ((c->cells_dbg & 7) << 2) | (push ? 32 : 0) | (c->chunk_overlap ? 64 : 0) |
                        ((c->gather_split & 15) << 7) | (c->elem_ratio ? 1 << 11 : 0) | ((int64_t)(c->gather_threads >> 7) << 12) |
                        ((int64_t)c->chunk_tiles << 16);
    auto it = c->asm_graphs.find(key);
    if (it == c->asm_graphs.end()) {
      const int n_chunks = (int)c->chunk_cols.size() - 1;
      const int64_t l0 = g_launches.load();
      cudaGraph_t graph = nullptr;
      CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      int rc = FEA_GPU_OK;
      auto body = [&]() -> int {
        CU(cudaEventRecord(c->ev_asm, c->stream));
        CU(cudaStreamWaitEvent(c->asm_stream, c->ev_asm, 0));
        for (int ch = 0; ch < n_chunks; ++ch) {
          const int t0 = ch * c->chunk_tiles, nt = std::min(c->chunk_tiles, n_tiles - t0);
          TRY(element_launch(c, with_k, with_r, a, t0, nt));
          cudaStream_t gs = c->stream;
          if (c->chunk_overlap) {
            CU(cudaEventRecord(c->chunk_ev[(size_t)ch], c->stream));
            CU(cudaStreamWaitEvent(c->asm_stream, c->chunk_ev[(size_t)ch], 0));
            gs = c->asm_stream;
          }
          if (push)
            TRY(launch_gather_cells(c, c->chunk_cols[(size_t)ch], c->chunk_cols[(size_t)ch + 1], chunk_bc, gs));
          else
            TRY(launch_gather_pull(c, chunk_bc, c->slice_order + c->chunk_slices[(size_t)ch],
                                   c->chunk_slices[(size_t)ch + 1] - c->chunk_slices[(size_t)ch], gs));
        }
        CU(cudaEventRecord(c->ev_asm, c->asm_stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_asm, 0));
        return FEA_GPU_OK;
      };
      rc = body();
      const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
      if (rc != FEA_GPU_OK) return rc;
      CU(ce);
      fea_gpu_ctx::AsmGraph ag;
      CU(cudaGraphInstantiate(&ag.exec, graph, 0));
      CU(cudaGraphDestroy(graph));
      ag.launches = (int)(g_launches.load() - l0);   // kernels per graph launch (the capture itself counted them once)
      it = c->asm_graphs.emplace(key, ag).first;
    }
    CU(cudaGraphLaunch(it->second.exec, c->stream));
    g_launches.fetch_add(it->second.launches, std::memory_order_relaxed);
    phase_end(c, PH_ELEM);
    *gathered = true;   // the stiffness gather is done
    return FEA_GPU_OK;
  }
  phase_begin(c, PH_ELEM);
  const int rc = element_launch(c, with_k, with_r, a, 0, n_tiles);
  phase_end(c, PH_ELEM);
  return rc;
}

static int element_launch(fea_gpu_ctx *c, bool with_k, bool with_r, fea::ElemArgs &a, int tile0, int n_tiles) {
  a.tile0 = tile0;
  int rc;
  // A5 with mu != 0: lam' / mu' is the same at every Gauss point (element_kernels.cuh, RATIO)
  a.rho = c->mu != 0.0 ? c->lambda / c->mu : 0.0;
  const bool ratio = c->model == FEA_MODEL_A5 && c->mu != 0.0 && std::isfinite(a.rho) && c->elem_ratio;
  if (c->model == FEA_MODEL_A5 && ratio)
    rc = c->ng == 5 ? launch_element<0, 5, true>(c, with_k, with_r, a, n_tiles) : launch_element<0, 4, true>(c, with_k, with_r, a, n_tiles);
  else if (c->model == FEA_MODEL_A5)
    rc = c->ng == 5 ? launch_element<0, 5, false>(c, with_k, with_r, a, n_tiles) : launch_element<0, 4, false>(c, with_k, with_r, a, n_tiles);
  else
    rc = c->ng == 5 ? launch_element<1, 5, false>(c, with_k, with_r, a, n_tiles) : launch_element<1, 4, false>(c, with_k, with_r, a, n_tiles);
  return rc;
}

static fea::SellMat sell_mat(fea_gpu_ctx *c) {
  fea::SellMat A;
  A.n_slices = c->plan.n_slices;
  A.slice_ptr = c->slice_ptr;
  A.sell_row = c->sell_row;
  A.bcol = c->bcol;
  A.vals = c->vals;
  return A;
}

// pull gather over all slices (list == nullptr) or over `n` entries of the device list `list`
static int launch_gather_pull(fea_gpu_ctx *c, bool with_bc, const int32_t *list, int n, cudaStream_t st) {
  if (n <= 0) return FEA_GPU_OK;
  const uint8_t *pf = with_bc ? c->sflag : nullptr;
  const int32_t *mir = c->gather_sym ? c->cmirror : nullptr;
  const int sp = c->gather_split;
  const int grid = n * sp;   // CTAs are dispatched in slice order
  if (c->gather_mode == 9 && c->gather9_ok && !FEA_KE_INTERLEAVED && !list)
    fea::gather_blocks9_kernel<4, FEA_G9_MINCTAS><<<grid, 128, 0, st>>>(sell_mat(c), sp, c->cptr, c->csrc, c->Ke, pf);
  else
    switch (c->gather_threads) {
      case 1024: fea::gather_blocks_kernel<1024, 1><<<grid, 1024, 0, st>>>(sell_mat(c), sp, c->cptr, c->csrc, c->Ke, pf, mir, list, n); break;
      case 512: fea::gather_blocks_kernel<512, 2><<<grid, 512, 0, st>>>(sell_mat(c), sp, c->cptr, c->csrc, c->Ke, pf, mir, list, n); break;
      case 128: fea::gather_blocks_kernel<128, 8><<<grid, 128, 0, st>>>(sell_mat(c), sp, c->cptr, c->csrc, c->Ke, pf, mir, list, n); break;
      default: fea::gather_blocks_kernel<256, 5><<<grid, 256, 0, st>>>(sell_mat(c), sp, c->cptr, c->csrc, c->Ke, pf, mir, list, n); break;
    }
  LAUNCHED();
  return FEA_GPU_OK;
}

static int ensure_mirror(fea_gpu_ctx *c) {
  if (c->gather_sym && !c->cmirror) {
    TRY(dev_upload(&c->cmirror, c->plan.cmirror, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return FEA_GPU_OK;
}

static int gather_stiffness(fea_gpu_ctx *c, bool with_bc) {
  phase_begin(c, PH_GATHER_K);
  if (c->gather_mode == 2) {
    TRY(launch_gather_cells(c, 0, c->n_cols_active, with_bc, c->stream));
  } else {
    TRY(ensure_mirror(c));
    TRY(launch_gather_pull(c, with_bc, nullptr, c->plan.n_slices, c->stream));
  }
  phase_end(c, PH_GATHER_K);
  return FEA_GPU_OK;
}

static int gather_residual(fea_gpu_ctx *c) {
  phase_begin(c, PH_GATHER_R);
  fea::gather_residual_kernel<<<cdiv(3 * (int64_t)c->n_own, 256), 256, 0, c->stream>>>(
      c->n_own, c->rptr, c->rsrc, c->Re, c->ne_pad, c->R, nullptr);
  LAUNCHED();
  phase_end(c, PH_GATHER_R);
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_update_state(fea_gpu_handle c) {
  GROUP(c, fea_gpu_update_state(ci));
  CHECK_H(c);
  return element_pass(c, false, false);
}
extern "C" int fea_gpu_assemble_stiffness(fea_gpu_handle c) {
  GROUP(c, fea_gpu_assemble_stiffness(ci));
  CHECK_H(c);
  bool gathered = false;
  TRY(element_pass(c, true, false, false, &gathered));
  return gathered ? FEA_GPU_OK : gather_stiffness(c, false);
}
extern "C" int fea_gpu_assemble_residual(fea_gpu_handle c) {
  GROUP(c, fea_gpu_assemble_residual(ci));
  CHECK_H(c);
  TRY(element_pass(c, false, true));
  return gather_residual(c);
}
extern "C" int fea_gpu_assemble_all(fea_gpu_handle c, int32_t flags) {
  GROUP(c, fea_gpu_assemble_all(ci, flags));
  CHECK_H(c);
  const bool with_k = (flags & FEA_ASSEMBLE_STIFFNESS) != 0, fuse_bc = (flags & FEA_ASSEMBLE_FUSE_BC) != 0;
  bool gathered = false;
  TRY(element_pass(c, with_k, true, fuse_bc, &gathered));
  if (with_k && !gathered) TRY(gather_stiffness(c, fuse_bc));
  if (!fuse_bc) return gather_residual(c);
  // solver_apply_prescribed_bc(self, 0) folded into the two gathers: rows and columns of prescribed
  // DOFs cancelled keeping the diagonal, their right-hand side rows zero (fea_solver.c:1244-1257)
  phase_begin(c, PH_GATHER_R);
  fea::gather_residual_kernel<<<cdiv(3 * (int64_t)c->n_own, 256), 256, 0, c->stream>>>(
      c->n_own, c->rptr, c->rsrc, c->Re, c->ne_pad, c->R, c->pflag);
  LAUNCHED();
  phase_end(c, PH_GATHER_R);
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_bad_points(fea_gpu_handle c, int64_t *count) {
  if (c && c->group && !t_worker) {   // sum over the ranks (interface elements are seen by every rank that computes them)
    if (!count) return FEA_GPU_ERR_ARG;
    std::vector<int64_t> part(c->group->ctx.size(), 0);
    const int rc = group_run(c->group, [&](int gi) { return fea_gpu_bad_points(c->group->ctx[(size_t)gi], &part[(size_t)gi]); });
    *count = 0;
    for (int64_t v : part) *count += v;
    return rc;
  }
  CHECK_H(c);
  if (!count) return FEA_GPU_ERR_ARG;
  unsigned long long v = 0;
  CU(cudaMemcpyAsync(&v, c->bad, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *count = (int64_t)v;
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------
// SpMV / BC / PCG

// y = K x over all slices, or over the slices of `list`; with `dot_out` the partial of x . y of those rows
// is reduced (fixed order) into *dot_out and the launch is a no-op once *done_flag is set
static int launch_spmv_list(fea_gpu_ctx *c, const double *x, double *y, double *dot_out, const int *done_flag,
                            const int32_t *list, int n_list, bool side = false) {
  cudaStream_t st = side ? c->comm_stream : c->stream;          // side: the boundary slices, beside the interior ones
  double *scratch = side ? c->partials_b : c->partials;
  unsigned int *counter = c->counters + (side ? 4 : 0);
  if (n_list <= 0) {
    if (dot_out) CU(cudaMemsetAsync(dot_out, 0, sizeof(double), st));
    return FEA_GPU_OK;
  }
  const int grid = std::min(cdiv((int64_t)n_list * 32, 256), MAX_PARTIALS);
  if (dot_out)
    fea::spmv_sell_kernel<true><<<grid, 256, 0, st>>>(n_list, c->slice_ptr, c->sell_row, c->bcol, c->vals, x, y, scratch,
                                                      counter, dot_out, done_flag, list);
  else
    fea::spmv_sell_kernel<false><<<grid, 256, 0, st>>>(n_list, c->slice_ptr, c->sell_row, c->bcol, c->vals, x, y, scratch,
                                                       counter, nullptr, nullptr, list);
  LAUNCHED();
  return FEA_GPU_OK;
}
static int launch_spmv(fea_gpu_ctx *c, const double *x, double *y, bool fuse_dot) {
  return launch_spmv_list(c, x, y, fuse_dot ? &c->ctl->pq : nullptr, fuse_dot ? &c->ctl->done : nullptr, nullptr,
                          c->plan.n_slices);
}

extern "C" int fea_gpu_apply_bc(fea_gpu_handle c, double lambda) {
  GROUP(c, fea_gpu_apply_bc(ci, lambda));
  CHECK_H(c);
  phase_begin(c, PH_BC);
  const int n = 3 * c->n_own;
  if (lambda != 0.0 && c->any_presc_value) {
    // R -= K . (presc * lambda): the column sweep of solver_apply_single_bc (fea_solver.c:1250-1252)
    CU(cudaMemsetAsync(c->p, 0, sizeof(double) * 3 * (size_t)c->n_local, c->stream));
    fea::axpy_kernel<<<cdiv(3 * (int64_t)c->n_local, 256), 256, 0, c->stream>>>(3 * c->n_local, lambda, c->pval, c->p);
    LAUNCHED();
    TRY(launch_spmv(c, c->p, c->q, false));
    fea::axpy_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(n, -1.0, c->q, c->R);
    LAUNCHED();
  }
  const int grid = std::min(cdiv((int64_t)c->plan.n_slices * 32, 256), 148 * 32);
  fea::cancel_kernel<<<grid, 256, 0, c->stream>>>(sell_mat(c), c->sflag);
  LAUNCHED();
  fea::rhs_fix_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->vals, c->sdiag, c->pflag, c->pval, lambda, c->R);
  LAUNCHED();
  phase_end(c, PH_BC);
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_save_stiffness(fea_gpu_handle c) {
  GROUP(c, fea_gpu_save_stiffness(ci));
  CHECK_H(c);
  if (!c->vals_saved) TRY(dev_alloc(&c->vals_saved, (size_t)c->n_slots * 9));
  CU(cudaMemcpyAsync(c->vals_saved, c->vals, sizeof(double) * (size_t)c->n_slots * 9, cudaMemcpyDeviceToDevice, c->stream));
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_restore_stiffness(fea_gpu_handle c) {
  GROUP(c, fea_gpu_restore_stiffness(ci));
  CHECK_H(c);
  if (!c->vals_saved) {
    g_err = "no saved stiffness";
    return FEA_GPU_ERR_ARG;
  }
  CU(cudaMemcpyAsync(c->vals, c->vals_saved, sizeof(double) * (size_t)c->n_slots * 9, cudaMemcpyDeviceToDevice, c->stream));
  return FEA_GPU_OK;
}

static int stall_limit_of(fea_gpu_ctx *c) {
  // PCG's ||r|| is not monotone; on slender domains it plateaus for O(sqrt(cond)) iterations
  // (2000+ on a 55x220x55 bar), so the window is generous -- inconsistent singular systems are
  // caught much earlier by the divergence test in pcg_step_control
  const double ndof = 3.0 * (double)c->plan.n_nodes_global;
  return c->pcg_stall > 0 ? c->pcg_stall : std::max(500, (int)(50.0 * std::cbrt(ndof)));
}

// w = K z with the halo exchange of z hidden behind the slices that need no ghost value
static int spmv_with_halo(fea_gpu_ctx *c, double *z, double *w, double *dot_a, double *dot_b, const int *done_flag) {
  const bool timed = c->sp_used < SPMV_EVENT_POOL;
  if (!c->has_comm || c->plan.nbr_rank.empty() || !c->pcg_overlap || c->n_bound == 0) {
    TRY(halo_exchange(c, z));
    if (timed) cudaEventRecord(c->sp_a[c->sp_used], c->stream);
    TRY(launch_spmv_list(c, z, w, dot_a, done_flag, nullptr, c->plan.n_slices));
    if (timed) cudaEventRecord(c->sp_b[c->sp_used++], c->stream);
    if (dot_b) CU(cudaMemsetAsync(dot_b, 0, sizeof(double), c->stream));
    return FEA_GPU_OK;
  }
  // comm stream: pack, send / receive, then the few slices that need the ghost values; main stream: all the
  // others meanwhile.  The two products write disjoint rows of w and reduce their dot partials separately.
  CU(cudaEventRecord(c->ev_vec, c->stream));
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_vec, 0));
  TRY(halo_exchange(c, z, c->comm_stream));
  TRY(launch_spmv_list(c, z, w, dot_b, done_flag, c->sl_bound, c->n_bound, true));
  CU(cudaEventRecord(c->ev_halo, c->comm_stream));
  if (timed) cudaEventRecord(c->sp_a[c->sp_used], c->stream);
  TRY(launch_spmv_list(c, z, w, dot_a, done_flag, c->sl_inner, c->n_inner));
  if (timed) cudaEventRecord(c->sp_b[c->sp_used++], c->stream);   // the interior product (the boundary one runs beside it)
  CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  return FEA_GPU_OK;
}

struct SolveExit {
  int done = 0, iters = 0;
  double rr_final = 0, bb = 0;
};

// classic PCG: p.Ap, then r.z and r.r -- two reductions per iteration
static int solve_classic(fea_gpu_ctx *c, double tol, int32_t max_iter, int32_t flags, SolveExit *ex) {
  const int n = 3 * c->n_own;
  const bool multi = c->has_comm;
  const int abs_tol = (flags & FEA_SOLVE_ABS_TOL) ? 1 : 0;
  const int vgrid = std::min(cdiv(n, fea::RED_THREADS), MAX_PARTIALS / 4);
  const int stall_limit = stall_limit_of(c);
  CU(cudaMemcpyAsync(&c->ctl->stall_limit, &stall_limit, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (flags & FEA_SOLVE_X0_RHS) {
    CU(cudaMemcpyAsync(c->u, c->R, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    TRY(halo_exchange(c, c->u));
    TRY(launch_spmv(c, c->u, c->q, false));
    fea::pcg_init_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, c->R, c->q, c->dinv, c->r, c->p, c->partials,
                                                                    c->counters + 1, c->ctl, tol, abs_tol, !multi);
  } else {
    CU(cudaMemsetAsync(c->u, 0, sizeof(double) * (size_t)n, c->stream));
    fea::pcg_init_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, c->R, nullptr, c->dinv, c->r, c->p,
                                                                    c->partials, c->counters + 1, c->ctl, tol, abs_tol,
                                                                    !multi);
  }
  LAUNCHED();
  if (multi) {
    TRY(allreduce_sum(c, &c->ctl->rz_new, 3));
    fea::pcg_init_finalize_kernel<<<1, 1, 0, c->stream>>>(c->ctl, tol, abs_tol);
    LAUNCHED();
  }

  // the starting iterate is the first checkpoint
  CU(cudaMemcpyAsync(c->u_saved, c->u, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  int launched = 0;
  bool done = false;
  while (!done) {
    const int batch = std::min(c->pcg_batch, max_iter - launched);
    for (int b = 0; b < batch; ++b) {
      TRY(halo_exchange(c, c->p));
      const bool timed = c->sp_used < SPMV_EVENT_POOL;
      if (timed) cudaEventRecord(c->sp_a[c->sp_used], c->stream);
      TRY(launch_spmv(c, c->p, c->q, true));
      if (timed) cudaEventRecord(c->sp_b[c->sp_used++], c->stream);
      if (multi) TRY(allreduce_sum(c, &c->ctl->pq, 1));
      fea::pcg_update_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, c->p, c->q, c->dinv, c->u, c->r, c->partials,
                                                                        c->counters + 2, c->ctl, !multi);
      LAUNCHED();
      if (multi) {
        TRY(allreduce_sum(c, &c->ctl->rz_new, 2));
        fea::pcg_control_kernel<<<1, 1, 0, c->stream>>>(c->ctl);
        LAUNCHED();
      }
      fea::pcg_direction_kernel<<<std::min(cdiv(n, 256), 148 * 8), 256, 0, c->stream>>>(n, c->r, c->dinv, c->p, c->u, c->u_saved, c->ctl);
      LAUNCHED();
    }
    launched += batch;
    CU(cudaMemcpyAsync(c->ctl_host, c->ctl, sizeof(PcgCtl), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    done = c->ctl_host->done || launched >= max_iter;
  }
  ex->done = c->ctl_host->done;
  ex->iters = c->ctl_host->iters;
  ex->bb = c->ctl_host->bb;
  ex->rr_final = c->ctl_host->done == 2 ? c->ctl_host->rr_saved : c->ctl_host->rr_exit;
  return FEA_GPU_OK;
}

// single-reduction PCG (sparse_kernels.cuh: Pcg2State): z lives in c->p (it needs the ghost range), w in c->q
static int solve_single_reduction(fea_gpu_ctx *c, double tol, int32_t max_iter, int32_t flags, SolveExit *ex) {
  const int n = 3 * c->n_own;
  const bool multi = c->has_comm;
  const int abs_tol = (flags & FEA_SOLVE_ABS_TOL) ? 1 : 0;
  const int vgrid = std::min(cdiv(n, fea::RED_THREADS), MAX_PARTIALS / 4);
  if (!c->pd) TRY(dev_alloc(&c->pd, (size_t)n));
  if (!c->sv) TRY(dev_alloc(&c->sv, (size_t)n));
  fea::Pcg2State *S = c->st2;
  const double *q0 = nullptr;
  if (flags & FEA_SOLVE_X0_RHS) {
    CU(cudaMemcpyAsync(c->u, c->R, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    TRY(halo_exchange(c, c->u));
    TRY(launch_spmv(c, c->u, c->q, false));
    q0 = c->q;
  } else {
    CU(cudaMemsetAsync(c->u, 0, sizeof(double) * (size_t)n, c->stream));
  }
  fea::pcg2_init_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, c->R, q0, c->dinv, c->r, c->p, c->pd, c->sv,
                                                                   c->partials, c->counters + 1, &S[0]);
  LAUNCHED();
  TRY(spmv_with_halo(c, c->p, c->q, &S[0].delta_a, &S[0].delta_b, &S[0].done));
  if (multi) TRY(allreduce_sum(c, &S[0].gamma, 5));
  fea::pcg2_finalize_kernel<<<1, 1, 0, c->stream>>>(&S[0], tol, abs_tol, stall_limit_of(c));
  LAUNCHED();
  CU(cudaMemcpyAsync(c->u_saved, c->u, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  int launched = 0;
  bool done = false;
  while (!done) {
    const int batch = std::min(c->pcg_batch, max_iter - launched);
    for (int b = 0; b < batch; ++b) {
      fea::Pcg2State *Sc = &S[(launched + b) & 1], *Sn = &S[(launched + b + 1) & 1];
      fea::pcg2_step_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, Sc, Sn, c->p, c->q, c->pd, c->sv, c->u, c->r,
                                                                       c->dinv, c->u_saved, c->partials, c->counters + 2);
      LAUNCHED();
      TRY(spmv_with_halo(c, c->p, c->q, &Sn->delta_a, &Sn->delta_b, &Sn->done));
      if (multi) TRY(allreduce_sum(c, &Sn->gamma, 4));
    }
    launched += batch;
    CU(cudaMemcpyAsync(c->st2_host, &S[launched & 1], sizeof(fea::Pcg2State), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // the stop test of the iterate just produced runs in the next step: apply it here too, so a solve
    // that has converged does not queue another batch of no-op iterations
    done = c->st2_host->done || c->st2_host->rr <= c->st2_host->thresh || launched >= max_iter;
  }
  // the state read back is the one the NEXT step would consume: its own stop test has not run yet
  const fea::Pcg2State &h = *c->st2_host;
  ex->done = h.done;
  ex->iters = h.iters;
  ex->bb = h.bb;
  ex->rr_final = h.done == 2 ? h.rr_saved : (h.done ? h.rr_exit : h.rr);
  if (!h.done && h.rr <= h.thresh) ex->done = 1;   // met on the very last permitted iteration
  // mirror into the classic control block so fea_gpu_phase_ms reports one exit record
  c->ctl_host->done = ex->done;
  c->ctl_host->iters = h.iters;
  c->ctl_host->bb = h.bb;
  c->ctl_host->best_rr = h.best_rr;
  c->ctl_host->rr_exit = h.done ? h.rr_exit : h.rr;
  c->ctl_host->rr_saved = h.rr_saved;
  c->ctl_host->stall = h.stall;
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------
// Chebyshev-Jacobi PCG (the PCG_ILU request of a task file): classic recurrences, general preconditioner

static int dot_to(fea_gpu_ctx *c, const double *a, const double *b, double *dev_out) {
  const int n = 3 * c->n_own;
  fea::dot_kernel<<<std::min(cdiv(n, fea::RED_THREADS), MAX_PARTIALS), fea::RED_THREADS, 0, c->stream>>>(
      n, a, b, c->partials, c->counters + 3, dev_out);
  LAUNCHED();
  return allreduce_sum(c, dev_out, 1);
}

// upper bound of the spectrum of D^-1 A: power iteration (it converges from below; the safety factor and the
// fact that the top of an FE spectrum is dense make 1.2 x the estimate an upper bound in practice)
static int estimate_lmax(fea_gpu_ctx *c, double *lmax) {
  const int n = 3 * c->n_own, grid = std::min(cdiv(n, 256), 148 * 8);
  fea::power_start_kernel<<<grid, 256, 0, c->stream>>>(n, c->cz);
  LAUNCHED();
  TRY(dot_to(c, c->cz, c->cz, c->scalar));
  for (int it = 0; it < 30; ++it) {
    TRY(halo_exchange(c, c->cz));
    TRY(launch_spmv(c, c->cz, c->q, false));
    fea::scale_dinv_kernel<<<grid, 256, 0, c->stream>>>(n, c->q, c->dinv, c->scalar, c->cz);   // D^-1 A v / |v|
    LAUNCHED();
    TRY(dot_to(c, c->cz, c->cz, c->scalar));
  }
  double nrm2 = 0;
  CU(cudaMemcpyAsync(&nrm2, c->scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  // after the last step cz = D^-1 A v_prev / |v_prev| with v_prev the previous (unnormalised) iterate: |cz| -> lmax
  *lmax = std::sqrt(nrm2);
  return FEA_GPU_OK;
}

// z = p_d(D^-1 A) D^-1 r  (z in c->cz)
static int chebyshev_apply(fea_gpu_ctx *c, const double *r, double lmax) {
  const int n = 3 * c->n_own, grid = std::min(cdiv(n, 256), 148 * 8);
  const double b = lmax, a = lmax / c->cheb_ratio, theta = 0.5 * (b + a), delta = 0.5 * (b - a), sigma = theta / delta;
  fea::cheb_first_kernel<<<grid, 256, 0, c->stream>>>(n, r, c->dinv, 1.0 / theta, c->cd, c->cz);
  LAUNCHED();
  double rho = 1.0 / sigma;
  for (int k = 1; k < c->cheb_degree; ++k) {
    const double rho2 = 1.0 / (2.0 * sigma - rho);
    TRY(halo_exchange(c, c->cz));
    TRY(launch_spmv(c, c->cz, c->q, false));
    fea::cheb_step_kernel<<<grid, 256, 0, c->stream>>>(n, r, c->q, c->dinv, rho2 * rho, 2.0 * rho2 / delta, c->cd, c->cz);
    LAUNCHED();
    rho = rho2;
  }
  return FEA_GPU_OK;
}

static int solve_chebyshev(fea_gpu_ctx *c, double tol, int32_t max_iter, int32_t flags, SolveExit *ex) {
  const int n = 3 * c->n_own;
  const int abs_tol = (flags & FEA_SOLVE_ABS_TOL) ? 1 : 0;
  const int vgrid = std::min(cdiv(n, fea::RED_THREADS), MAX_PARTIALS / 4);
  if (!c->cz) TRY(dev_alloc(&c->cz, 3 * (size_t)c->n_local));
  if (!c->cd) TRY(dev_alloc(&c->cd, (size_t)n));
  CU(cudaMemsetAsync(c->cz, 0, sizeof(double) * 3 * (size_t)c->n_local, c->stream));
  double lmax = 0;
  TRY(estimate_lmax(c, &lmax));
  lmax *= 1.2;
  c->last_lmax = lmax;
  const int stall_limit = stall_limit_of(c);
  CU(cudaMemcpyAsync(&c->ctl->stall_limit, &stall_limit, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  const double *q0 = nullptr;
  if (flags & FEA_SOLVE_X0_RHS) {
    CU(cudaMemcpyAsync(c->u, c->R, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    TRY(halo_exchange(c, c->u));
    TRY(launch_spmv(c, c->u, c->q, false));
    q0 = c->q;
  } else {
    CU(cudaMemsetAsync(c->u, 0, sizeof(double) * (size_t)n, c->stream));
  }
  // r, b.b, r.r from the Jacobi start kernel (finalize = 0: the sums are completed below); then z, p = z, r.z
  fea::pcg_init_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, c->R, q0, c->dinv, c->r, c->p, c->partials,
                                                                  c->counters + 1, c->ctl, tol, abs_tol, 0);
  LAUNCHED();
  TRY(allreduce_sum(c, &c->ctl->rz_new, 3));    // (rz_new, rr, bb) are contiguous; rz_new is replaced next
  TRY(chebyshev_apply(c, c->r, lmax));
  CU(cudaMemcpyAsync(c->p, c->cz, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  TRY(dot_to(c, c->r, c->cz, &c->ctl->rz_new));
  fea::pcg_init_finalize_kernel<<<1, 1, 0, c->stream>>>(c->ctl, tol, abs_tol);
  LAUNCHED();
  CU(cudaMemcpyAsync(c->u_saved, c->u, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  int launched = 0;
  bool done = false;
  const int batch_cap = std::max(1, c->pcg_batch / c->cheb_degree);
  while (!done) {
    const int batch = std::min(batch_cap, max_iter - launched);
    for (int bi = 0; bi < batch; ++bi) {
      TRY(halo_exchange(c, c->p));
      const bool timed = c->sp_used < SPMV_EVENT_POOL;
      if (timed) cudaEventRecord(c->sp_a[c->sp_used], c->stream);
      TRY(launch_spmv(c, c->p, c->q, true));
      if (timed) cudaEventRecord(c->sp_b[c->sp_used++], c->stream);
      TRY(allreduce_sum(c, &c->ctl->pq, 1));
      fea::pcg_update_kernel<<<vgrid, fea::RED_THREADS, 0, c->stream>>>(n, c->p, c->q, c->dinv, c->u, c->r, c->partials,
                                                                        c->counters + 2, c->ctl, 0);
      LAUNCHED();
      TRY(allreduce_sum(c, &c->ctl->rr, 1));
      TRY(chebyshev_apply(c, c->r, lmax));      // (a few wasted products after convergence inside a batch)
      TRY(dot_to(c, c->r, c->cz, &c->ctl->rz_new));
      fea::pcg_control_kernel<<<1, 1, 0, c->stream>>>(c->ctl);
      LAUNCHED();
      fea::pcg_direction_z_kernel<<<std::min(cdiv(n, 256), 148 * 8), 256, 0, c->stream>>>(n, c->cz, c->p, c->u, c->u_saved, c->ctl);
      LAUNCHED();
    }
    launched += batch;
    CU(cudaMemcpyAsync(c->ctl_host, c->ctl, sizeof(PcgCtl), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    done = c->ctl_host->done || launched >= max_iter;
  }
  ex->done = c->ctl_host->done;
  ex->iters = c->ctl_host->iters;
  ex->bb = c->ctl_host->bb;
  ex->rr_final = c->ctl_host->done == 2 ? c->ctl_host->rr_saved : c->ctl_host->rr_exit;
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_solve(fea_gpu_handle c, double tol, int32_t max_iter, int32_t flags, int32_t *iters,
                             double *relres) {
  if (c && c->group && !t_worker) {   // every rank returns the same (all-reduced) record
    std::vector<int32_t> it(c->group->ctx.size(), 0);
    std::vector<double> rr(c->group->ctx.size(), 0.0);
    const int rc = group_run(c->group, [&](int gi) {
      return fea_gpu_solve(c->group->ctx[(size_t)gi], tol, max_iter, flags, &it[(size_t)gi], &rr[(size_t)gi]);
    });
    if (iters) *iters = it[0];
    if (relres) *relres = rr[0];
    return rc;
  }
  CHECK_H(c);
  if (max_iter < 0 || !(tol >= 0.0)) return FEA_GPU_ERR_ARG;
  const int n = 3 * c->n_own;
  phase_begin(c, PH_PCG);
  c->sp_used = 0;
  fea::jacobi_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->vals, c->sdiag, c->dinv);
  LAUNCHED();
  SolveExit ex;
  const bool single_reduction = c->pcg_variant == 1;
  if (c->precond == 1)
    TRY(solve_chebyshev(c, tol, max_iter, flags, &ex));
  else if (single_reduction)
    TRY(solve_single_reduction(c, tol, max_iter, flags, &ex));
  else
    TRY(solve_classic(c, tol, max_iter, flags, &ex));
  if (ex.done == 2)   // stalled or diverged: the checkpoint is the answer (see pcg_step_control)
    CU(cudaMemcpyAsync(c->u, c->u_saved, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  phase_end(c, PH_PCG);
  c->last_iters = ex.iters;
  c->last_exit = ex.done;
  if (iters) *iters = ex.iters;
  const double rel = ex.bb > 0.0 ? std::sqrt(ex.rr_final / ex.bb) : 0.0;
  if (relres) *relres = rel;
  if (!ex.done) {
    g_err = "PCG reached max_iter";
    return FEA_GPU_ERR_NOT_CONVERGED;
  }
  if (ex.done == 2 && !(flags & FEA_SOLVE_ACCEPT_STALL)) {
    char buf[200];
    snprintf(buf, sizeof(buf), "PCG stopped on its stall/divergence guard after %d iterations at relative residual %.3e; "
             "u holds the best checkpointed iterate", ex.iters, rel);
    g_err = buf;
    return FEA_GPU_ERR_STALLED;
  }
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_dot_R_u(fea_gpu_handle c, double *out) {
  if (c && c->group && !t_worker) {
    if (!out) return FEA_GPU_ERR_ARG;
    std::vector<double> v(c->group->ctx.size(), 0.0);
    const int rc = group_run(c->group, [&](int gi) { return fea_gpu_dot_R_u(c->group->ctx[(size_t)gi], &v[(size_t)gi]); });
    *out = v[0];
    return rc;
  }
  CHECK_H(c);
  if (!out) return FEA_GPU_ERR_ARG;
  const int n = 3 * c->n_own;
  fea::dot_kernel<<<std::min(cdiv(n, fea::RED_THREADS), MAX_PARTIALS), fea::RED_THREADS, 0, c->stream>>>(
      n, c->R, c->u, c->partials, c->counters + 3, c->scalar);
  LAUNCHED();
  TRY(allreduce_sum(c, c->scalar, 1));
  CU(cudaMemcpyAsync(out, c->scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_spmv(fea_gpu_handle c, const double *x, double *y) {
  GROUP(c, fea_gpu_spmv(ci, x, gi ? nullptr : y));
  CHECK_H(c);
  if (!x || (!y && !(c->group && t_worker))) return FEA_GPU_ERR_ARG;
  TRY(scatter_local(c, x, c->p, c->n_local));
  TRY(launch_spmv(c, c->p, c->q, false));
  return gather_owned(c, c->q, y);
}

extern "C" int fea_gpu_bench_spmv(fea_gpu_handle c, int32_t reps, double *ms_per_spmv) {
  if (c && c->group && !t_worker) {
    std::vector<double> v(c->group->ctx.size(), 0.0);
    const int rc = group_run(c->group, [&](int gi) { return fea_gpu_bench_spmv(c->group->ctx[(size_t)gi], reps, &v[(size_t)gi]); });
    if (ms_per_spmv) *ms_per_spmv = *std::max_element(v.begin(), v.end());
    return rc;
  }
  CHECK_H(c);
  if (reps < 1 || !ms_per_spmv) return FEA_GPU_ERR_ARG;
  TRY(launch_spmv(c, c->p, c->q, false));  // warm
  CU(cudaEventRecord(c->tm_a, c->stream));
  for (int i = 0; i < reps; ++i) TRY(launch_spmv(c, c->p, c->q, false));
  CU(cudaEventRecord(c->tm_b, c->stream));
  CU(cudaEventSynchronize(c->tm_b));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, c->tm_a, c->tm_b));
  *ms_per_spmv = (double)ms / reps;
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------
// read-back

extern "C" int fea_gpu_get_state(fea_gpu_handle c, double *graddefs, double *stresses) {
  GROUP(c, fea_gpu_get_state(ci, graddefs, stresses));
  CHECK_H(c);
  const fea::Plan &pl = c->plan;
  const size_t per = (size_t)c->ng * 9;
  if (!c->export_buf) TRY(dev_alloc(&c->export_buf, per * (size_t)c->n_elems));
  std::vector<double> tmp(per * (size_t)c->n_elems);
  for (int which = 0; which < 2; ++which) {
    double *dst = which ? stresses : graddefs;
    if (!dst) continue;
    fea::state_export_kernel<<<cdiv((int64_t)per * c->n_elems, 256), 256, 0, c->stream>>>(
        c->n_elems, c->ne_pad, c->ng, which ? c->S_soa : c->F_soa, c->export_buf);
    LAUNCHED();
    CU(cudaMemcpyAsync(tmp.data(), c->export_buf, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int32_t e = 0; e < c->n_elems; ++e)
      if (pl.elem_owned[(size_t)e])
        std::memcpy(dst + per * (size_t)pl.elem_gid[(size_t)e], tmp.data() + per * (size_t)e, sizeof(double) * per);
  }
  return FEA_GPU_OK;
}

static void ensure_elem_map(fea_gpu_ctx *c) {
  const fea::Plan &pl = c->plan;
  if (!c->elem_g2l.empty()) return;
  c->elem_g2l.assign((size_t)pl.n_elems_global, -1);
  for (int32_t le = 0; le < pl.n_elems; ++le) c->elem_g2l[(size_t)pl.elem_gid[(size_t)le]] = le;
}

// graddefs / stresses of a LIST of elements (global ids): what fea_gpu_get_state returns, without moving
// the tensors of the whole mesh (0.7 GB per million elements).  found[k] = 1 if element k is local to this
// rank (owned or an interface element it computes redundantly), else its output rows are left untouched.
extern "C" int fea_gpu_get_state_elems(fea_gpu_handle c, int32_t n, const int32_t *elems, double *graddefs,
                                       double *stresses, int32_t *found) {
  if (c && c->group && !t_worker) {   // rank by rank: a later rank overwrites identical values, masks are OR-ed
    std::vector<int32_t> f((size_t)std::max(n, 0), 0);
    if (found) std::fill(found, found + std::max(n, 0), 0);
    for (fea_gpu_ctx *ci : c->group->ctx) {
      t_worker = true;
      const int rc = fea_gpu_get_state_elems(ci, n, elems, graddefs, stresses, f.data());
      t_worker = false;
      if (rc != FEA_GPU_OK) return rc;
      if (found)
        for (int32_t k = 0; k < n; ++k) found[k] |= f[(size_t)k];
    }
    return FEA_GPU_OK;
  }
  CHECK_H(c);
  if (n < 0 || (n > 0 && !elems)) return FEA_GPU_ERR_ARG;
  if (n == 0) return FEA_GPU_OK;
  ensure_elem_map(c);
  const size_t per = (size_t)c->ng * 9;
  std::vector<int32_t> le((size_t)n);
  for (int32_t k = 0; k < n; ++k) {
    if (elems[k] < 0 || elems[k] >= c->plan.n_elems_global) return FEA_GPU_ERR_ARG;
    le[(size_t)k] = c->elem_g2l[(size_t)elems[k]];
    if (found) found[k] = le[(size_t)k] >= 0;
  }
  int32_t *dle = nullptr;
  double *dout = nullptr;
  std::vector<double> tmp(per * (size_t)n);
  auto body = [&]() -> int {
    TRY(dev_alloc(&dle, (size_t)n));
    TRY(dev_alloc(&dout, per * (size_t)n));
    CU(cudaMemcpyAsync(dle, le.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    for (int which = 0; which < 2; ++which) {
      double *dst = which ? stresses : graddefs;
      if (!dst) continue;
      fea::state_export_list_kernel<<<cdiv((int64_t)per * n, 256), 256, 0, c->stream>>>(n, dle, c->ne_pad, c->ng,
                                                                                       which ? c->S_soa : c->F_soa, dout);
      LAUNCHED();
      CU(cudaMemcpyAsync(tmp.data(), dout, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      for (int32_t k = 0; k < n; ++k)
        if (le[(size_t)k] >= 0) std::memcpy(dst + per * (size_t)k, tmp.data() + per * (size_t)k, sizeof(double) * per);
    }
    return FEA_GPU_OK;
  };
  const int rc = body();
  cudaFree(dle);
  cudaFree(dout);
  return rc;
}

// Diagnostic read-back of one staged element matrix: what the last element pass with stiffness left in the
// staging buffer for `element` (global id), expanded from the a <= b block triangle to the dense 30 x 30
// [3a+i][3b+j] the reference builds in solver_local_constitutive_part + solver_local_initial_stess_part
// (fea_solver.c:887-1068; compare with the capture at its sp_matrix_element_add call sites).
extern "C" int fea_gpu_get_element_matrix(fea_gpu_handle c, int32_t element, double *ke900) {
  if (c && c->group && !t_worker) {
    int rc = FEA_GPU_ERR_ARG;
    for (fea_gpu_ctx *ci : c->group->ctx) {
      t_worker = true;
      rc = fea_gpu_get_element_matrix(ci, element, ke900);
      t_worker = false;
      if (rc != FEA_GPU_ERR_ARG) break;
    }
    return rc;
  }
  CHECK_H(c);
  if (!ke900 || element < 0 || element >= c->plan.n_elems_global) return FEA_GPU_ERR_ARG;
  ensure_elem_map(c);
  const int32_t le = c->elem_g2l[(size_t)element];
  if (le < 0) {
    g_err = "element is not local to this rank";
    return FEA_GPU_ERR_ARG;
  }
  double st[fea::KE_STRIDE];
  if (c->gather_mode == 2) {
    // direct assembly: the element's blocks sit in the cells of their slots (blocks of rows owned by another
    // rank are not stored at all: they read as zero here)
    if (!c->cells) {
      g_err = "no stiffness assembled yet";
      return FEA_GPU_ERR_ARG;
    }
    std::memset(st, 0, sizeof(st));
    for (int code = 0; code < fea::NTRI; ++code) {
      const uint32_t d = c->plan.edest[(size_t)code * c->ne_pad + le];
      if (d == fea::CELL_NONE) continue;
      double cell[9], *blk = st + 100 * (code / 11) + 9 * (code % 11);
      CU(cudaMemcpyAsync(cell, c->cells + (size_t)(d & 0x7fffffffu) * 5, sizeof(cell), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      for (int k = 0; k < 9; ++k) blk[k] = (d >> 31) ? cell[(k % 3) * 3 + k / 3] : cell[k];
    }
  } else {
    if (!c->Ke) {
      g_err = "no stiffness assembled yet";
      return FEA_GPU_ERR_ARG;
    }
    if (FEA_KE_INTERLEAVED)   // 16-byte chunk k of the element at double offset ((le / 32) 250 + k) 64 + 2 (le % 32)
      CU(cudaMemcpy2DAsync(st, 16, c->Ke + (size_t)(le / 32) * (32 * fea::KE_STRIDE) + 2 * (le % 32), 512, 16, fea::KE_STRIDE / 2,
                           cudaMemcpyDeviceToHost, c->stream));
    else
      CU(cudaMemcpyAsync(st, c->Ke + (size_t)le * fea::KE_STRIDE, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  for (int a = 0; a < fea::NEN; ++a)
    for (int b = a; b < fea::NEN; ++b) {
      const int code = fea::ke_code(a, b);
      const double *blk = st + 100 * (code / 11) + 9 * (code % 11);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          ke900[(3 * a + i) * 30 + 3 * b + j] = blk[3 * i + j];
          ke900[(3 * b + j) * 30 + 3 * a + i] = blk[3 * i + j];
        }
    }
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_get_csr(fea_gpu_handle c, int64_t *n_rows, int64_t *nnz, int32_t *rows, int32_t *rowptr,
                               int32_t *colidx, double *vals) {
  if (c && c->group && !t_worker) {
    g_err = "fea_gpu_get_csr returns the rows of ONE rank: not available on a multi-GPU handle";
    return FEA_GPU_ERR_ARG;
  }
  CHECK_H(c);
  const fea::Plan &pl = c->plan;
  // a node that belongs to no element carries an internal lone diagonal block (so Jacobi has a
  // diagonal to invert); the reference's container has no entry there, so the export hides it
  int64_t isolated = 0;
  for (int32_t l = 0; l < c->n_own; ++l) isolated += pl.rptr[(size_t)l + 1] == pl.rptr[(size_t)l];
  if (n_rows) *n_rows = 3 * (int64_t)c->n_own;
  if (nnz) *nnz = 9 * (c->nnzb - isolated);
  // rows are returned in ascending GLOBAL id (the device keeps them in Morton / SELL order)
  std::vector<int32_t> by_gid((size_t)c->n_own);
  for (int32_t l = 0; l < c->n_own; ++l) by_gid[(size_t)l] = l;
  std::sort(by_gid.begin(), by_gid.end(),
            [&](int32_t a, int32_t b) { return pl.node_gid[(size_t)a] < pl.node_gid[(size_t)b]; });
  if (rows)
    for (int32_t k = 0; k < c->n_own; ++k)
      for (int i = 0; i < 3; ++i) rows[3 * (size_t)k + i] = 3 * pl.node_gid[(size_t)by_gid[(size_t)k]] + i;
  if (!rowptr && !colidx && !vals) return FEA_GPU_OK;
  std::vector<double> sv;
  if (vals) {
    sv.resize(9 * (size_t)c->n_slots);
    CU(cudaMemcpyAsync(sv.data(), c->vals, sizeof(double) * sv.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  // columns are ascending in LOCAL ids; re-sort each block row by GLOBAL id for the export
  std::vector<std::pair<int32_t, int32_t>> order;
  int64_t out = 0;
  for (int32_t k = 0; k < c->n_own; ++k) {
    const int32_t l = by_gid[(size_t)k];
    const int32_t b0 = pl.browptr[(size_t)l], b1 = pl.browptr[(size_t)l + 1];
    const int32_t rl = pl.row_lane[(size_t)l];
    const int64_t sbase = pl.slice_ptr[(size_t)(rl / fea::SELL_C)];
    const int lane = rl % fea::SELL_C;
    order.clear();
    if (pl.rptr[(size_t)l + 1] != pl.rptr[(size_t)l])
      for (int32_t q = b0; q < b1; ++q) order.emplace_back(pl.node_gid[(size_t)pl.bcol[(size_t)q]], q - b0);
    std::sort(order.begin(), order.end());
    for (int i = 0; i < 3; ++i) {
      if (rowptr) rowptr[3 * (size_t)k + i] = (int32_t)out;
      for (auto &pr : order)
        for (int j = 0; j < 3; ++j, ++out) {
          if (colidx) colidx[out] = 3 * pr.first + j;
          if (vals) vals[out] = sv[(size_t)(9 * (sbase + (int64_t)pr.second * fea::SELL_C) + fea::val_off(3 * i + j, lane))];
        }
    }
  }
  if (rowptr) rowptr[3 * (size_t)c->n_own] = (int32_t)out;
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------
// host-buffer path (end-to-end step)

extern "C" int fea_gpu_host_alloc(void **out, uint64_t bytes) {
  if (!out) return FEA_GPU_ERR_ARG;
  CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_host_free(void *p) {
  if (p) CU(cudaFreeHost(p));
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_step_from_host(fea_gpu_handle c, const double *x, int32_t with_stiffness, double *R,
                                      uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
  if (c && c->group && !t_worker) {   // every rank fills its own rows of R
    std::vector<uint64_t> a(c->group->ctx.size(), 0), b(c->group->ctx.size(), 0);
    const int rc = group_run(c->group, [&](int gi) {
      return fea_gpu_step_from_host(c->group->ctx[(size_t)gi], x, with_stiffness, R, &a[(size_t)gi], &b[(size_t)gi]);
    });
    if (h2d_bytes) { *h2d_bytes = 0; for (uint64_t v : a) *h2d_bytes += v; }
    if (d2h_bytes) { *d2h_bytes = 0; for (uint64_t v : b) *d2h_bytes += v; }
    return rc;
  }
  CHECK_H(c);
  if (!x || !R) return FEA_GPU_ERR_ARG;
  const fea::Plan &pl = c->plan;
  const size_t nl3 = 3 * (size_t)c->n_local, n3 = 3 * (size_t)c->n_own;
  uint64_t h2d = 0, d2h = 0;
  if (c->io_range) {
    // one DMA of the global id range that covers the local nodes, permuted on the device
    h2d = sizeof(double) * 3 * (uint64_t)c->io_cnt;
    CU(cudaMemcpyAsync(c->io_buf, x + 3 * (size_t)c->io_lo, h2d, cudaMemcpyHostToDevice, c->stream));
    fea::gather_nodes_kernel<<<cdiv(3 * (int64_t)c->n_local, 256), 256, 0, c->stream>>>(c->n_local, c->io_idx, c->io_buf, c->x);
    LAUNCHED();
  } else {
    if (!c->stage_h) CU(cudaHostAlloc((void **)&c->stage_h, sizeof(double) * nl3, cudaHostAllocDefault));
    for (int32_t l = 0; l < c->n_local; ++l)
      std::memcpy(c->stage_h + 3 * (size_t)l, x + 3 * (size_t)pl.node_gid[(size_t)l], 3 * sizeof(double));
    h2d = sizeof(double) * nl3;
    CU(cudaMemcpyAsync(c->x, c->stage_h, h2d, cudaMemcpyHostToDevice, c->stream));
  }
  bool gathered = false;
  TRY(element_pass(c, with_stiffness != 0, true, true, &gathered));
  // residual first: its way back to the host (copy stream) overlaps the stiffness gather
  phase_begin(c, PH_GATHER_R);
  fea::gather_residual_kernel<<<cdiv(3 * (int64_t)c->n_own, 256), 256, 0, c->stream>>>(
      c->n_own, c->rptr, c->rsrc, c->Re, c->ne_pad, c->R, c->pflag);   // prescribed rows -> 0 (lambda = 0)
  LAUNCHED();
  phase_end(c, PH_GATHER_R);
  d2h = sizeof(double) * n3;
  const bool ranged = c->io_range && c->own_range;
  if (ranged) {
    fea::scatter_nodes_kernel<<<cdiv(3 * (int64_t)c->n_own, 256), 256, 0, c->stream>>>(c->n_own, c->own_idx, c->R, c->io_buf);
    LAUNCHED();
  } else if (!c->stage_h) {
    CU(cudaHostAlloc((void **)&c->stage_h, sizeof(double) * nl3, cudaHostAllocDefault));
  }
  CU(cudaEventRecord(c->ev_copy, c->stream));
  CU(cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
  if (ranged)
    CU(cudaMemcpyAsync(R + 3 * (size_t)c->own_lo, c->io_buf, d2h, cudaMemcpyDeviceToHost, c->copy_stream));
  else
    CU(cudaMemcpyAsync(c->stage_h, c->R, d2h, cudaMemcpyDeviceToHost, c->copy_stream));
  if (with_stiffness && !gathered) TRY(gather_stiffness(c, true));   // Dirichlet cancellation fused into the gather
  CU(cudaStreamSynchronize(c->copy_stream));
  CU(cudaStreamSynchronize(c->stream));
  if (!ranged)
    for (int32_t l = 0; l < c->n_own; ++l)
      std::memcpy(R + 3 * (size_t)pl.node_gid[(size_t)l], c->stage_h + 3 * (size_t)l, 3 * sizeof(double));
  if (h2d_bytes) *h2d_bytes = h2d;
  if (d2h_bytes) *d2h_bytes = d2h;
  return FEA_GPU_OK;
}

// ---------------------------------------------------------------------------------
// introspection / measurement

extern "C" int fea_gpu_counts(fea_gpu_handle c, int64_t out[16]) {
  if (!c || !out) return FEA_GPU_ERR_ARG;
  if (c->group && !t_worker) {        // sums over the ranks where a sum means something, else the maximum
    int64_t tmp[16];
    std::memset(out, 0, sizeof(int64_t) * 16);
    for (fea_gpu_ctx *ci : c->group->ctx) {
      fea::plan_counts(ci->plan, tmp);
      tmp[12] = ci->gather9_ok ? 1 : 0;
      for (int k : {0, 3, 4, 6, 7, 10, 11}) out[k] += tmp[k];
      for (int k : {1, 2, 5}) out[k] = std::max(out[k], tmp[k]);
      out[8] = tmp[8];
      out[9] = tmp[9];
      out[12] = ci == c->group->ctx[0] ? tmp[12] : std::min(out[12], tmp[12]);
    }
    out[13] = (int64_t)c->group->ctx.size();
    return FEA_GPU_OK;
  }
  fea::plan_counts(c->plan, out);
  out[12] = c->gather9_ok ? 1 : 0;
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_sync(fea_gpu_handle c) {
  GROUP(c, fea_gpu_sync(ci));
  CHECK_H(c);
  CU(cudaStreamSynchronize(c->stream));
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_timer_start(fea_gpu_handle c) {
  GROUP(c, fea_gpu_timer_start(ci));
  CHECK_H(c);
  CU(cudaEventRecord(c->tm_a, c->stream));
  return FEA_GPU_OK;
}
extern "C" int fea_gpu_timer_stop(fea_gpu_handle c, double *ms) {
  if (c && c->group && !t_worker) {
    std::vector<double> v(c->group->ctx.size(), 0.0);
    const int rc = group_run(c->group, [&](int gi) { return fea_gpu_timer_stop(c->group->ctx[(size_t)gi], &v[(size_t)gi]); });
    if (ms) *ms = *std::max_element(v.begin(), v.end());
    return rc;
  }
  CHECK_H(c);
  CU(cudaEventRecord(c->tm_b, c->stream));
  CU(cudaEventSynchronize(c->tm_b));
  float f = 0;
  CU(cudaEventElapsedTime(&f, c->tm_a, c->tm_b));
  if (ms) *ms = f;
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_phase_ms(fea_gpu_handle c, double out[16]) {
  if (c && c->group && !t_worker) {
    if (!out) return FEA_GPU_ERR_ARG;
    std::vector<double> v(16 * c->group->ctx.size(), 0.0);
    const int rc = group_run(c->group, [&](int gi) { return fea_gpu_phase_ms(c->group->ctx[(size_t)gi], &v[16 * (size_t)gi]); });
    for (int k = 0; k < 16; ++k) {
      out[k] = v[(size_t)k];
      if (k < 8)
        for (size_t r = 1; r < c->group->ctx.size(); ++r) out[k] = std::max(out[k], v[16 * r + (size_t)k]);
    }
    return rc;
  }
  CHECK_H(c);
  if (!out) return FEA_GPU_ERR_ARG;
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 16; ++i) out[i] = 0.0;
  for (int i = 0; i < PH_COUNT; ++i) {
    if (i == PH_SPMV) continue;
    const int n = std::min(c->ev_n[i], PHASE_EVENT_POOL);
    double sum = 0;
    int got = 0;
    for (int k = 0; k < n; ++k) {
      float f = 0;
      if (cudaEventElapsedTime(&f, c->ev_a[i][k], c->ev_b[i][k]) == cudaSuccess) {
        sum += f;
        ++got;
      }
    }
    if (got) out[i] = sum / got;     // average ms per call since the last read
    if (i == PH_ELEM) out[14] = got;
    c->ev_n[i] = 0;
  }
  double sp = 0;
  for (int i = 0; i < c->sp_used; ++i) {
    float f = 0;
    if (cudaEventElapsedTime(&f, c->sp_a[i], c->sp_b[i]) == cudaSuccess) sp += f;
  }
  out[PH_SPMV] = c->sp_used ? sp / c->sp_used : 0.0;  // average ms per in-solve SpMV launch
  out[8] = c->sp_used;
  out[9] = c->last_iters;
  if (c->ctl_host) {   // exit state of the last solve
    const double bb = c->ctl_host->bb > 0.0 ? c->ctl_host->bb : 1.0;
    out[10] = c->ctl_host->done;                          // 0 = max_iter, 1 = tolerance, 2 = stall / divergence guard
    out[11] = std::sqrt(c->ctl_host->best_rr / bb);       // best relative residual seen
    out[12] = std::sqrt(c->ctl_host->rr_exit / bb);       // relative residual of the last iterate
    out[13] = c->ctl_host->stall;
  }
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_set_param(fea_gpu_handle c, const char *name, double value) {
  if (!c || !name) return FEA_GPU_ERR_ARG;
  if (c->group && !t_worker) {
    for (fea_gpu_ctx *ci : c->group->ctx) {
      t_worker = true;
      const int rc = fea_gpu_set_param(ci, name, value);
      t_worker = false;
      if (rc != FEA_GPU_OK) return rc;
    }
    return FEA_GPU_OK;
  }
  const std::string k(name);
  const int v = (int)value;
  if (k == "gather_threads" && (v == 128 || v == 256 || v == 512 || v == 1024)) c->gather_threads = v;
  else if (k == "elem_ratio" && (v == 0 || v == 1)) c->elem_ratio = v != 0;
  else if (k == "gather_split" && v >= 1 && v <= 8) c->gather_split = v;
  else if (k == "gather_mode" && (v == 1 || v == 9 || v == 2)) c->gather_mode = v;
  else if (k == "chunk_tiles" && v >= 0) c->chunk_tiles = v;
  else if (k == "cells_dbg" && v >= 0) c->cells_dbg = v;
  else if (k == "gather_sym" && (v == 0 || v == 1)) c->gather_sym = v;
  else if (k == "chunk_overlap" && (v == 0 || v == 1)) c->chunk_overlap = v;
  else if (k == "pcg_batch" && v >= 1 && v <= 4096) c->pcg_batch = v;
  else if (k == "pcg_stall" && v >= 0) c->pcg_stall = v;
  else if (k == "pcg_variant" && v >= 0 && v <= 1) c->pcg_variant = v;
  else if (k == "precond" && (v == 0 || v == 1)) c->precond = v;
  else if (k == "cheb_degree" && v >= 1 && v <= 16) c->cheb_degree = v;
  else if (k == "cheb_ratio" && value >= 2.0 && value <= 1e4) c->cheb_ratio = value;
  else if (k == "pcg_overlap" && (v == 0 || v == 1)) c->pcg_overlap = v;
  else {
    g_err = "unknown parameter or value out of range: " + k;
    return FEA_GPU_ERR_ARG;
  }
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_measure_dmma(int32_t device, double *dmma_tflops) {
  if (!dmma_tflops) return FEA_GPU_ERR_ARG;
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double *out = nullptr;
  CU(cudaMalloc(&out, sizeof(double) * blocks * threads));
  double best = 0;
  float ms = 0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(a));
    fea::dmma_probe_kernel<<<blocks, threads>>>(out, iters);
    g_launches.fetch_add(1);
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    CU(cudaEventElapsedTime(&ms, a, b));
    // 8 mma.sync.m8n8k4 per iteration and warp, 8*8*4 FMA each
    const double tf = 2.0 * 256.0 * 8.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *dmma_tflops = best;
  cudaFree(out);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return FEA_GPU_OK;
}

// average device time of the two collectives of a PCG iteration, each timed alone on the context's
// stream: the halo exchange of a [n_local][3] vector and the all-reduce of the four iteration sums
extern "C" int fea_gpu_bench_comm(fea_gpu_handle c, int32_t reps, double *halo_ms, double *allreduce_ms) {
  if (c && c->group && !t_worker) {
    std::vector<double> a(c->group->ctx.size(), 0.0), b(c->group->ctx.size(), 0.0);
    const int rc = group_run(c->group, [&](int gi) {
      return fea_gpu_bench_comm(c->group->ctx[(size_t)gi], reps, &a[(size_t)gi], &b[(size_t)gi]);
    });
    if (halo_ms) *halo_ms = *std::max_element(a.begin(), a.end());
    if (allreduce_ms) *allreduce_ms = *std::max_element(b.begin(), b.end());
    return rc;
  }
  CHECK_H(c);
  if (reps < 1) return FEA_GPU_ERR_ARG;
  float ms = 0;
  if (halo_ms) {
    TRY(halo_exchange(c, c->p));
    CU(cudaEventRecord(c->tm_a, c->stream));
    for (int i = 0; i < reps; ++i) TRY(halo_exchange(c, c->p));
    CU(cudaEventRecord(c->tm_b, c->stream));
    CU(cudaEventSynchronize(c->tm_b));
    CU(cudaEventElapsedTime(&ms, c->tm_a, c->tm_b));
    *halo_ms = (double)ms / reps;
  }
  if (allreduce_ms) {
    CU(cudaMemsetAsync(c->st2, 0, 2 * sizeof(fea::Pcg2State), c->stream));
    TRY(allreduce_sum(c, &c->st2[0].gamma, 4));
    CU(cudaEventRecord(c->tm_a, c->stream));
    for (int i = 0; i < reps; ++i) TRY(allreduce_sum(c, &c->st2[0].gamma, 4));
    CU(cudaEventRecord(c->tm_b, c->stream));
    CU(cudaEventSynchronize(c->tm_b));
    CU(cudaEventElapsedTime(&ms, c->tm_a, c->tm_b));
    *allreduce_ms = (double)ms / reps;
  }
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_flush_l2(fea_gpu_handle c) {
  GROUP(c, fea_gpu_flush_l2(ci));
  CHECK_H(c);
  if (!c->flush) CU(cudaMalloc(&c->flush, FLUSH_BYTES));
  CU(cudaMemsetAsync(c->flush, 1, FLUSH_BYTES, c->stream));
  return FEA_GPU_OK;
}

extern "C" int fea_gpu_measure_peaks(int32_t device, double *dfma_tflops, double *copy_gbs) {
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  float ms = 0;
  if (dfma_tflops) {
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *out = nullptr;
    CU(cudaMalloc(&out, sizeof(double) * blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      CU(cudaEventRecord(a));
      fea::dfma_probe_kernel<<<blocks, threads>>>(out, iters);
      g_launches.fetch_add(1);
      CU(cudaEventRecord(b));
      CU(cudaEventSynchronize(b));
      CU(cudaEventElapsedTime(&ms, a, b));
      const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    *dfma_tflops = best;
    cudaFree(out);
  }
  if (copy_gbs) {
    const size_t n = (size_t)1 << 26;  // 2 x 1 GiB
    double2 *s = nullptr, *d = nullptr;
    CU(cudaMalloc(&s, sizeof(double2) * n));
    CU(cudaMalloc(&d, sizeof(double2) * n));
    CU(cudaMemset(s, 0, sizeof(double2) * n));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
      CU(cudaEventRecord(a));
      fea::copy_probe_kernel<<<prop.multiProcessorCount * 16, 512>>>(s, d, n);
      g_launches.fetch_add(1);
      CU(cudaEventRecord(b));
      CU(cudaEventSynchronize(b));
      CU(cudaEventElapsedTime(&ms, a, b));
      const double gbs = 2.0 * sizeof(double2) * n / (ms * 1e-3) / 1e9;
      if (rep > 0 && gbs > best) best = gbs;
    }
    *copy_gbs = best;
    cudaFree(s);
    cudaFree(d);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return FEA_GPU_OK;
}
