// Element phase of the hot path for sm_100a: per (element, Gauss point) geometry, F,
// Cauchy stress, tangent coefficients, then K_e (upper-triangular 3x3 blocks) and R_e.
//
// Replaces, fused into one pass over the elements:
//   solver_shape_gradients_alloc   fea_solver.c:656-722   J, det J, J^-1, grad N
//   solver_element_gauss_graddef   fea_solver.c:1131-1152 F through F^-1
//   fea_model_stress_*             fea_model.c:26-107
//   fea_model_ctensor_*            fea_model.c:110-148 (closed form of the symmetrised tensor)
//   solver_local_constitutive_part fea_solver.c:887-983
//   solver_local_initial_stess_part fea_solver.c:986-1068
//   solver_local_residual_forces   fea_solver.c:1072-1114
//
// Mapping: a CTA handles 32 elements with NG warps; warp w owns Gauss point w, lane l owns
// element l.  Phase A (one thread per (element, Gauss point)) keeps everything in
// registers; its results cross to phase B through shared memory laid out
// [gauss][field][lane] in 16-byte fields (conflict-free LDS.128/STS.128).  Phase B gives warp
// w the eleven a<=b blocks of staging region w (rows w and 9-w, fea_plan.hpp) of its lane's
// element, two consecutive blocks at a time, summed over the Gauss points.  38 doubles per (element,
// Gauss point) cross over (73 KB of shared memory per CTA); phase B wants ~150 registers, so two
// CTAs (10 warps) per SM.  (One block at a time under a 128-register cap, three CTAs per SM, was
// measured too: 1.47 against 1.39 ms.)
//
// Closed form used for the tangent (both models):
//   c^_ikjl = lam' d_ik d_jl + mu' (d_ij d_kl + d_il d_jk)
//   NH: lam' = lam/J, mu' = (mu - lam ln J)/J      A5: lam' = lam/J, mu' = mu/J   (J = det F)
//   K_ab[i][j] = wd ( lam' g_ai g_bj + mu' g_aj g_bi + d_ij (mu' g_a.g_b + g_a.sigma g_b) )
// with g_a = grad N_a in the current configuration and wd = w_g |det J|.
#pragma once
#include <cstdint>

#include "fea_plan.hpp"

namespace fea {

constexpr int ELEMS_PER_CTA = 32;
// Hand-over fields per Gauss point, [field][lane] (lane-contiguous: conflict-free):
//   double2 GA[b] = (g0,g1) for the 10 nodes, three double2 of the symmetric
//   s = wd sigma: (s00,s01), (s02,s11), (s12,s22), double2 LM = (lam' wd, mu' wd),
//   then plain doubles g2[b].  g = grad N_b in the current configuration, wd = w |det J|.
// The column side of a block needs only g (24 bytes per Gauss point); the row side forms
// t_a = (mu' wd I + s) g_a itself (9 FMAs + 3 adds per Gauss point and row) instead of reading it.
constexpr int FLD_DOUBLES = 38 * 32;          // doubles per Gauss point
constexpr int FLD_SG = 10 * 64, FLD_LM = 13 * 64, FLD_G2 = 14 * 64;
#ifndef FEA_KE_TMA_STORE
#define FEA_KE_TMA_STORE 0
#endif
constexpr int TILE_D2 = (FEA_KE_TMA_STORE ? 2 : 1) * 9 * 32;   // double2 per warp: the store tile of one block pair (two with the bulk-copy engine)
// PUSH kernels (direct assembly, fea_plan.cpp "cell layout"): a block pair leaves as two 80-byte cells; the
// tile row of an element is 11 double2 (10 used): a 44-word stride keeps the 16-byte stores of a quarter warp
// on distinct banks
constexpr int PUSH_ROW_D2 = 11;
constexpr int PUSH_TILE_D2 = PUSH_ROW_D2 * 32;

struct ElemTables {
  double dN[5][3][10];  // shape-function derivatives at the Gauss points (fea_solver.c:503-535)
  double w[5];          // weights incl. the tetrahedron's 1/6 (:32-54)
};
// both rules stay resident (index NG == 5): contexts with different n_gauss can be alive on one device
__constant__ ElemTables c_tabs[2];

struct ElemArgs {
  int n_elems;
  int ne_pad;                  // element count rounded up to 32 (SoA pitch)
  const int32_t *conn_soa;     // [10][ne_pad] local node ids
  const double *X0;            // [n_local][3]
  const double *x;             // [n_local][3]
  double lambda, mu;
  double rho;                  // lambda / mu (RATIO kernels only)
  double *F_soa;               // [ng*9][ne_pad]   (may be null)
  double *S_soa;               // [ng*9][ne_pad]
  double *Ke;                  // [n_elems][KE_STRIDE]  a<=b node-pair blocks, layout in fea_plan.hpp
  const uint32_t *edest;       // PUSH: [55][ne_pad] cell of block (code, element) | SRC_TRANSPOSE, CELL_NONE = not stored
  double2 *cells;              // PUSH: [n_cells][5] the cells of the upper slots, layer by layer (fea_plan.cpp)
  int tile0;                   // first 32-element tile of this launch (chunked assembly)
  int dbg;                     // diagnostics: 4 = no cell stores
  double *Re;                  // [30][ne_pad]
  unsigned long long *bad;     // count of points with det J <= 0, det F <= 0 or non-finite
};

__device__ __forceinline__ double det3(const double (&m)[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
         m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

// cofactor inverse, same formula as the reference's inv3x3 (dense_matrix.c:34-60)
__device__ __forceinline__ void inv3(const double (&m)[3][3], double det, double (&o)[3][3]) {
  const double id = 1.0 / det;
  o[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id;
  o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id;
  o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id;
  o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
}

// A5 with mu != 0: lam' / mu' = lambda / mu at every Gauss point, so u_a = rho v_a and with
// P = sum_q v_a (x) g_b (accumulated in k) the block is rho P + P^T (+ the diagonal term): 12 instead
// of 21 FMAs per Gauss point and no u in registers
__device__ __forceinline__ void ratio_block(double (&k)[9], double rho) {
  const double p01 = k[1], p02 = k[2], p12 = k[5];
  k[0] = fma(rho, k[0], k[0]);
  k[4] = fma(rho, k[4], k[4]);
  k[8] = fma(rho, k[8], k[8]);
  k[1] = fma(rho, p01, k[3]);
  k[3] = fma(rho, k[3], p01);
  k[2] = fma(rho, p02, k[6]);
  k[6] = fma(rho, k[6], p02);
  k[5] = fma(rho, p12, k[7]);
  k[7] = fma(rho, k[7], p12);
}

template <int MODEL, int NG, bool WITH_K, bool WITH_R, bool RATIO, bool PUSH = false>
__global__ void __launch_bounds__(NG * 32, 2) element_kernel(ElemArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const ElemTables &c_tab = c_tabs[NG == 5 ? 1 : 0];
  double *fld = reinterpret_cast<double *>(smraw);                       // [NG][FLD_DOUBLES]
  double2 *tiles = reinterpret_cast<double2 *>(fld + NG * FLD_DOUBLES);  // [NG][TILE_D2] store tiles
  int *goff = reinterpret_cast<int *>(tiles + NG * (PUSH ? PUSH_TILE_D2 : TILE_D2));   // [9][32] store offsets
  uint32_t *dsm = reinterpret_cast<uint32_t *>(goff + 9 * 32);          // PUSH: [55][32] cells of this CTA's blocks
  const int lane = threadIdx.x & 31;
  const int gp = threadIdx.x >> 5;
  const int e0 = (A.tile0 + blockIdx.x) * ELEMS_PER_CTA;
  const int e = e0 + lane;
  const bool live = e < A.n_elems;
  if (PUSH) {
    // where this CTA's 32 x 55 blocks go: fetched now, beside the node loads of phase 0, and first needed in phase B
    // (read straight from global memory at the point of use they cost a DRAM round trip per block pair: 38 % of
    // the kernel's stall samples, profiles/r2_push_ncu_summary.md)
    uint32_t d[(NTRI + NG - 1) / NG];
#pragma unroll
    for (int i = 0; i < (NTRI + NG - 1) / NG; ++i) {
      const int code = gp + i * NG;
      d[i] = (live && code < NTRI) ? A.edest[(size_t)code * A.ne_pad + e] : CELL_NONE;
    }
#pragma unroll
    for (int i = 0; i < (NTRI + NG - 1) / NG; ++i) {
      const int code = gp + i * NG;
      if (code < NTRI) dsm[code * 32 + lane] = d[i];
    }
  }

  // ------------------------------ phase 0 ------------------------------------
  // the NG warps of the CTA fetch the 10 nodes of the 32 elements once (coalesced connectivity,
  // gathered coordinates) into the region the store tiles will use later: [node][x0..2,X0..2][lane]
  double *coords = reinterpret_cast<double *>(tiles);
  {
    constexpr int NA = (10 + NG - 1) / NG;   // nodes per warp: both index loads, then all coordinate
    int node[NA];                            // loads, are in flight together (one round trip each)
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int a = gp + i * NG;
      node[i] = (live && a < 10) ? A.conn_soa[(size_t)a * A.ne_pad + e] : 0;
    }
    double cx[NA][6];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const double *xk = A.x + 3 * (size_t)node[i], *Xk = A.X0 + 3 * (size_t)node[i];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        cx[i][d] = xk[d];
        cx[i][3 + d] = Xk[d];
      }
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int a = gp + i * NG;
      if (a < 10) {
#pragma unroll
        for (int d = 0; d < 6; ++d) coords[((a * 6 + d) * 32) + lane] = cx[i][d];
      }
    }
  }
  if (WITH_K && !PUSH && gp == 0) {
    // where the 16-byte pieces of a block pair go: piece f = 32 it + lane of a warp's
    // [32 elements][9 double2] tile belongs to element f / 9 (same table for every warp and pair)
#pragma unroll
    for (int it = 0; it < 9; ++it) {
      const int f = it * 32 + lane, le = f / 9;
      goff[f] = le * KE_STRIDE + 2 * (f - 9 * le);
    }
  }
  __syncthreads();

  // ------------------------------ phase A ------------------------------------
  {
    // J[i][j] = sum_k dN_k/dxi_i * x_k,j      (fea_solver.c:690-696)
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const double x0 = coords[(k * 6 + 0) * 32 + lane], x1 = coords[(k * 6 + 1) * 32 + lane],
                   x2 = coords[(k * 6 + 2) * 32 + lane];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double d = c_tab.dN[gp][i][k];
        J[i][0] = fma(d, x0, J[i][0]);
        J[i][1] = fma(d, x1, J[i][1]);
        J[i][2] = fma(d, x2, J[i][2]);
      }
    }
    const double detJ = det3(J);
    bool ok = live && (detJ != 0.0);  // reference skips the point when det J == 0 exactly (:697)
    double Ji[3][3];
    inv3(J, ok ? detJ : 1.0, Ji);

    // g[i][a] = sum_k J^-1[i][k] dN[k][a]       (:714-718), handed over node by node, and
    // F^-1[i][j] = sum_k g[j][k] X0_k,i, then inverted  (:1141-1152)
    double *my = fld + gp * FLD_DOUBLES;
    double Fi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int a = 0; a < 10; ++a) {
      const double d0 = c_tab.dN[gp][0][a], d1 = c_tab.dN[gp][1][a], d2 = c_tab.dN[gp][2][a];
      double ga[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ga[i] = Ji[i][0] * d0 + Ji[i][1] * d1 + Ji[i][2] * d2;
      const double X0 = coords[(a * 6 + 3) * 32 + lane], X1 = coords[(a * 6 + 4) * 32 + lane],
                   X2 = coords[(a * 6 + 5) * 32 + lane];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Fi[0][j] = fma(ga[j], X0, Fi[0][j]);
        Fi[1][j] = fma(ga[j], X1, Fi[1][j]);
        Fi[2][j] = fma(ga[j], X2, Fi[2][j]);
      }
      reinterpret_cast<double2 *>(my + a * 64)[lane] = make_double2(ok ? ga[0] : 0.0, ok ? ga[1] : 0.0);
      my[FLD_G2 + a * 32 + lane] = ok ? ga[2] : 0.0;
    }
    const double detFi = det3(Fi);
    double F[3][3];
    if (detFi != 0.0) {
      inv3(Fi, detFi, F);
    } else {  // the reference ignores the failed inversion and keeps F^-1 (:1152)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) F[i][j] = Fi[i][j];
    }
    const double Jf = det3(F);

    // Cauchy stress and tangent coefficients
    double S[3][3], lam1, mu1;
    if (MODEL == 1) {  // compressible Neo-Hookean, fea_model.c:79-107, 129-148
      const double lnJ = log(Jf);
      const double iJ = 1.0 / Jf;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double b = F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2];
          S[i][j] = A.mu * (b - (i == j ? 1.0 : 0.0)) * iJ + (i == j ? A.lambda * lnJ * iJ : 0.0);
        }
      lam1 = A.lambda * iJ;
      mu1 = (A.mu - A.lambda * lnJ) * iJ;
    } else {  // A5 = St.Venant-Kirchhoff, Cauchy form, fea_model.c:26-77, 110-127
      double E[3][3], trE = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          E[i][j] = 0.5 * (F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j] - (i == j ? 1.0 : 0.0));
      trE = E[0][0] + E[1][1] + E[2][2];
      const double iJ = 1.0 / Jf;
      double P[3][3], Q[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P[i][j] = ((i == j ? A.lambda * trE : 0.0) + 2.0 * A.mu * E[i][j]) * iJ;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Q[i][j] = F[i][0] * P[0][j] + F[i][1] * P[1][j] + F[i][2] * P[2][j];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[i][j] = Q[i][0] * F[j][0] + Q[i][1] * F[j][1] + Q[i][2] * F[j][2];
      lam1 = A.lambda * iJ;
      mu1 = A.mu * iJ;
    }

    if (live) {
      if (A.F_soa) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          A.F_soa[(size_t)(gp * 9 + c) * A.ne_pad + e] = F[c / 3][c % 3];
          A.S_soa[(size_t)(gp * 9 + c) * A.ne_pad + e] = S[c / 3][c % 3];
        }
      }
      if (!(detJ > 0.0) || !(Jf > 0.0) || !isfinite(S[0][0] + S[1][1] + S[2][2]))
        atomicAdd(A.bad, 1ULL);
    }

    // hand over to phase B: wd sigma and the scaled coefficients
    const double wd = ok ? c_tab.w[gp] * fabs(detJ) : 0.0;  // fabs: fea_solver.c:958,1047,1104
    const double lw = lam1 * wd, mw = mu1 * wd;
    reinterpret_cast<double2 *>(my + FLD_SG)[lane] = make_double2(ok ? wd * S[0][0] : 0.0, ok ? wd * S[0][1] : 0.0);
    reinterpret_cast<double2 *>(my + FLD_SG + 64)[lane] = make_double2(ok ? wd * S[0][2] : 0.0, ok ? wd * S[1][1] : 0.0);
    reinterpret_cast<double2 *>(my + FLD_SG + 128)[lane] = make_double2(ok ? wd * S[1][2] : 0.0, ok ? wd * S[2][2] : 0.0);
    reinterpret_cast<double2 *>(my + FLD_LM)[lane] = make_double2(lw, mw);
  }
  __syncthreads();

  // ------------------------------ phase B ------------------------------------
  // Warp w takes staging region w (fea_plan.hpp): rows w and 9-w of the a<=b block triangle, eleven
  // blocks.  With s' symmetric, g_a . t_b = t_a . g_b, so
  //   K_ab[i][j] = sum_q  u_ai g_bj + v_aj g_bi + d_ij t_a . g_b,   u = lam' wd g_a, v = mu' wd g_a
  // and the column side streams 24 bytes per Gauss point from shared memory; u, v, t of the row stay
  // in registers (RATIO kernels: A5 with mu != 0, where u = (lambda / mu) v, keep only v and t).
  // The row's residual R_e[a] = -sum_q wd sigma g_a (fea_solver.c:1094-1109) falls out of the same loads.
#define GA2(q, b) reinterpret_cast<const double2 *>(fld + (q)*FLD_DOUBLES + (b)*64)[lane]
#define SG2(q, h) reinterpret_cast<const double2 *>(fld + (q)*FLD_DOUBLES + FLD_SG + (h)*64)[lane]
#define LM2(q) reinterpret_cast<const double2 *>(fld + (q)*FLD_DOUBLES + FLD_LM)[lane]
#define G2D(q, b) fld[(q)*FLD_DOUBLES + FLD_G2 + (b)*32 + lane]

  if (WITH_K || WITH_R) {
    double2 *tile = tiles + gp * (PUSH ? PUSH_TILE_D2 : TILE_D2);
    const int n_here = min(ELEMS_PER_CTA, A.n_elems - e0);
    unsigned vmask = 0;   // bit it: piece 32 it + lane belongs to an element that exists
#pragma unroll
    for (int it = 0; it < 9; ++it)
      if ((it * 32 + lane) / 9 < n_here) vmask |= 1u << it;
    double *kcta = A.Ke + (size_t)e0 * KE_STRIDE;
#if FEA_KE_TMA_STORE
    int tma_flip = 0;
#endif
    for (int pr = gp; pr < 5; pr += NG)
      for (int half = 0; half < 2; ++half) {
        const int a = half ? 9 - pr : pr;
        double ua[RATIO ? 1 : NG][3], va[NG][3], ta[NG][3];
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          const double2 G = GA2(q, a), S0 = SG2(q, 0), S1 = SG2(q, 1), S2 = SG2(q, 2), LM = LM2(q);
          const double g2 = G2D(q, a);
          // explicitly rounded where the compiler could fuse differently from one instantiation to
          // the next: t, v and the residual must not depend on whether K is built in the same pass
          // wd sigma g_a: the residual integrand itself (fea_solver.c:1094-1109), so R_e carries rounding
          // of size eps |sigma|, not eps mu' (the two differ by mu / |sigma|, ~1e3 at small strain)
          const double ts0 = fma(S0.x, G.x, fma(S0.y, G.y, __dmul_rn(S1.x, g2)));
          const double ts1 = fma(S0.y, G.x, fma(S1.y, G.y, __dmul_rn(S2.x, g2)));
          const double ts2 = fma(S1.x, G.x, fma(S2.x, G.y, __dmul_rn(S2.y, g2)));
          if (!RATIO) {
            ua[q][0] = LM.x * G.x;
            ua[q][1] = LM.x * G.y;
            ua[q][2] = LM.x * g2;
          }
          va[q][0] = __dmul_rn(LM.y, G.x);
          va[q][1] = __dmul_rn(LM.y, G.y);
          va[q][2] = __dmul_rn(LM.y, g2);
          ta[q][0] = __dadd_rn(ts0, va[q][0]);   // t_a = (mu' wd I + wd sigma) g_a for the d_ij term of K
          ta[q][1] = __dadd_rn(ts1, va[q][1]);
          ta[q][2] = __dadd_rn(ts2, va[q][2]);
          if (WITH_R) {
            r0 = __dadd_rn(r0, ts0);
            r1 = __dadd_rn(r1, ts1);
            r2 = __dadd_rn(r2, ts2);
          }
        }
        if (WITH_R && live) {
          A.Re[(size_t)(3 * a + 0) * A.ne_pad + e] = -r0;
          A.Re[(size_t)(3 * a + 1) * A.ne_pad + e] = -r1;
          A.Re[(size_t)(3 * a + 2) * A.ne_pad + e] = -r2;
        }
        // Blocks (a,b), (a,b+1) at an even staging position are 144 contiguous, 16-byte aligned
        // bytes of K_e.  Each thread drops its pair into its row of the warp's tile and the warp
        // stores it as 16-byte pieces walking those 144-byte chunks with consecutive lanes.  (v1
        // stored 8 bytes per lane at a 3960-byte stride: 27 L2 sectors per request; v3 transposed but spent ~25 instructions of index arithmetic per 8-byte store,
        // profiles/r1_v3_ncu_full_summary.md -- here the per-lane offsets come from a table.)
        if (WITH_K)
        for (int b = a; b < 10; b += 2) {
          const bool two = b + 1 < 10;   // warp-uniform
          uint32_t d0 = CELL_NONE, d1 = CELL_NONE;
          if (PUSH) {   // where this element's two blocks go (the blocks of a pair have consecutive codes, fea_plan.hpp)
            const uint32_t *dp = dsm + ke_code(a, b) * 32 + lane;
            d0 = dp[0];
            if (two) d1 = dp[32];
          }
          double k0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, k1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            {
              const double2 G = GA2(q, b);
              const double gb[3] = {G.x, G.y, G2D(q, b)};
              s0 = fma(ta[q][0], gb[0], fma(ta[q][1], gb[1], fma(ta[q][2], gb[2], s0)));
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  k0[3 * i + j] = RATIO ? fma(va[q][i], gb[j], k0[3 * i + j])
                                        : fma(ua[q][i], gb[j], fma(va[q][j], gb[i], k0[3 * i + j]));
            }
            if (two) {
              const double2 G = GA2(q, b + 1);
              const double gb[3] = {G.x, G.y, G2D(q, b + 1)};
              s1 = fma(ta[q][0], gb[0], fma(ta[q][1], gb[1], fma(ta[q][2], gb[2], s1)));
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  k1[3 * i + j] = RATIO ? fma(va[q][i], gb[j], k1[3 * i + j])
                                        : fma(ua[q][i], gb[j], fma(va[q][j], gb[i], k1[3 * i + j]));
            }
          }
          if (RATIO) {
            ratio_block(k0, A.rho);
            ratio_block(k1, A.rho);
          }
          k0[0] += s0;
          k0[4] += s0;
          k0[8] += s0;
          k1[0] += s1;
          k1[4] += s1;
          k1[8] += s1;
          if (PUSH) {
            // Each thread parks its blocks -- transposed where the slot wants K_e[b][a] -- as 80-byte cells in its row
            // of the warp's tile; the warp then writes cell by cell, five consecutive lanes per cell, the cell index
            // handed over by shuffle from the lane that owns the element.
            double2 *t = tile + lane * PUSH_ROW_D2;
            {
              const bool tr = (d0 >> 31) != 0;
              t[0] = make_double2(k0[0], tr ? k0[3] : k0[1]);
              t[1] = make_double2(tr ? k0[6] : k0[2], tr ? k0[1] : k0[3]);
              t[2] = make_double2(k0[4], tr ? k0[7] : k0[5]);
              t[3] = make_double2(tr ? k0[2] : k0[6], tr ? k0[5] : k0[7]);
              t[4] = make_double2(k0[8], 0.0);
            }
            if (two) {
              const bool tr = (d1 >> 31) != 0;
              t[5] = make_double2(k1[0], tr ? k1[3] : k1[1]);
              t[6] = make_double2(tr ? k1[6] : k1[2], tr ? k1[1] : k1[3]);
              t[7] = make_double2(k1[4], tr ? k1[7] : k1[5]);
              t[8] = make_double2(tr ? k1[2] : k1[6], tr ? k1[5] : k1[7]);
              t[9] = make_double2(k1[8], 0.0);
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 10; ++it) {
              if (it >= 5 && !two) break;
              const int sel = it / 5, f = (it - 5 * sel) * 32 + lane, le = f / 5, part = f - 5 * le;
              const uint32_t d = __shfl_sync(0xffffffffu, sel ? d1 : d0, le);
              const double2 v = tile[le * PUSH_ROW_D2 + sel * 5 + part];
              if (d != CELL_NONE && !(A.dbg & 4)) A.cells[(size_t)(d & 0x7fffffffu) * 5 + part] = v;
            }
            __syncwarp();
            continue;
          }
#if FEA_KE_INTERLEAVED
          if (live) {
            double2 *dst2 = reinterpret_cast<double2 *>(A.Ke + (size_t)(e0 / ELEMS_PER_CTA) * (32 * KE_STRIDE)) +
                            (size_t)((100 * pr + 9 * ke_pos(a, b)) >> 1) * 32 + lane;
            dst2[0 * 32] = make_double2(k0[0], k0[1]);
            dst2[1 * 32] = make_double2(k0[2], k0[3]);
            dst2[2 * 32] = make_double2(k0[4], k0[5]);
            dst2[3 * 32] = make_double2(k0[6], k0[7]);
            dst2[4 * 32] = make_double2(k0[8], two ? k1[0] : 0.0);
            if (two) {
              dst2[5 * 32] = make_double2(k1[1], k1[2]);
              dst2[6 * 32] = make_double2(k1[3], k1[4]);
              dst2[7 * 32] = make_double2(k1[5], k1[6]);
              dst2[8 * 32] = make_double2(k1[7], k1[8]);
            }
          }
          continue;
#endif
          double *dst = kcta + 100 * pr + 9 * ke_pos(a, b);
#if FEA_KE_TMA_STORE
          // Bulk-copy engine variant: every thread parks its 144 (80) bytes in its own row of the warp's
          // tile and hands them to the TMA unit with one cp.async.bulk (shared -> global); no LDS / STG
          // through the LSU pipe at all.  Two tile rows per thread alternate, so the wait for the engine to
          // have READ a row comes one block pair late.
          {
            double2 *t = tile + ((tma_flip & 1) * 32 + lane) * 9;
            if (tma_flip >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            t[0] = make_double2(k0[0], k0[1]);
            t[1] = make_double2(k0[2], k0[3]);
            t[2] = make_double2(k0[4], k0[5]);
            t[3] = make_double2(k0[6], k0[7]);
            t[4] = make_double2(k0[8], two ? k1[0] : 0.0);
            if (two) {
              t[5] = make_double2(k1[1], k1[2]);
              t[6] = make_double2(k1[3], k1[4]);
              t[7] = make_double2(k1[5], k1[6]);
              t[8] = make_double2(k1[7], k1[8]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (live) {
              const unsigned src = (unsigned)__cvta_generic_to_shared(t);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + lane * KE_STRIDE),
                           "r"(src), "r"(two ? 144 : 80)
                           : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            ++tma_flip;
          }
          continue;
#endif
          if (two) {
            double2 *t = tile + lane * 9;
            t[0] = make_double2(k0[0], k0[1]);
            t[1] = make_double2(k0[2], k0[3]);
            t[2] = make_double2(k0[4], k0[5]);
            t[3] = make_double2(k0[6], k0[7]);
            t[4] = make_double2(k0[8], k1[0]);
            t[5] = make_double2(k1[1], k1[2]);
            t[6] = make_double2(k1[3], k1[4]);
            t[7] = make_double2(k1[5], k1[6]);
            t[8] = make_double2(k1[7], k1[8]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 9; ++it) {
              const double2 v = tile[it * 32 + lane];
              const int off = goff[it * 32 + lane];
              if ((vmask >> it) & 1u) *reinterpret_cast<double2 *>(dst + off) = v;
            }
            __syncwarp();
          } else {   // the region's last block (a,9): 9 doubles + the pad, 5 double2 per element
            double2 *t = tile + lane * 5;
            t[0] = make_double2(k0[0], k0[1]);
            t[1] = make_double2(k0[2], k0[3]);
            t[2] = make_double2(k0[4], k0[5]);
            t[3] = make_double2(k0[6], k0[7]);
            t[4] = make_double2(k0[8], 0.0);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 5; ++it) {
              const int f = it * 32 + lane, le = f / 5;
              const double2 v = tile[f];
              if (le < n_here) *reinterpret_cast<double2 *>(dst + le * KE_STRIDE + 2 * (f - 5 * le)) = v;
            }
            __syncwarp();
          }
        }
      }
  }
#if FEA_KE_TMA_STORE
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the engine must have read the tile before the CTA retires
#endif
#undef GA2
#undef SG2
#undef LM2
#undef G2D
}

}  // namespace fea
