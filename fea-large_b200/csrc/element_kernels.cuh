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
// [gauss][field][lane] (conflict-free, lane-contiguous).  Phase B gives every thread 55/NG
// of the a<=b node-pair blocks of its lane's element and loops over the Gauss points.
//
// Closed form used for the tangent (both models):
//   c^_ikjl = lam' d_ik d_jl + mu' (d_ij d_kl + d_il d_jk)
//   NH: lam' = lam/J, mu' = (mu - lam ln J)/J      A5: lam' = lam/J, mu' = mu/J   (J = det F)
//   K_ab[i][j] = wd ( lam' g_ai g_bj + mu' g_aj g_bi + d_ij (mu' g_a.g_b + g_a.sigma g_b) )
// with g_a = grad N_a in the current configuration and wd = w_g |det J|.
#pragma once
#include <cstdint>

namespace fea {

constexpr int ELEMS_PER_CTA = 32;
constexpr int NFIELD = 62;  // g[3][10], t[3][10], lam', mu'
constexpr int TILE_LD = 19;  // per-lane pitch of the store-transpose tile (odd: conflict-free)

struct ElemTables {
  double dN[5][3][10];  // shape-function derivatives at the Gauss points (fea_solver.c:503-535)
  double w[5];          // weights incl. the tetrahedron's 1/6 (:32-54)
};
__constant__ ElemTables c_tab;

struct ElemArgs {
  int n_elems;
  int ne_pad;                  // element count rounded up to 32 (SoA pitch)
  const int32_t *conn_soa;     // [10][ne_pad] local node ids
  const double *X0;            // [n_local][3]
  const double *x;             // [n_local][3]
  double lambda, mu;
  double *F_soa;               // [ng*9][ne_pad]   (may be null)
  double *S_soa;               // [ng*9][ne_pad]
  double *Ke;                  // [n_elems][55][9]  upper-triangular node-pair blocks
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

template <int MODEL, int NG, bool WITH_K, bool WITH_R>
__global__ void __launch_bounds__(NG * 32, 2) element_kernel(ElemArgs A) {
  extern __shared__ double sm[];  // [NG][NFIELD][32] fields, then [NG][32][TILE_LD] store tiles
  const int lane = threadIdx.x & 31;
  const int gp = threadIdx.x >> 5;
  const int e = blockIdx.x * ELEMS_PER_CTA + lane;
  const bool live = e < A.n_elems;
  double *my = sm + (size_t)gp * NFIELD * 32 + lane;  // field f at my[f*32]

  // ------------------------------ phase 0 ------------------------------------
  // the NG warps of the CTA fetch the 10 nodes of the 32 elements once (coalesced connectivity,
  // gathered coordinates) into the region the store tiles will use later: [node][x0..2,X0..2][lane]
  double *coords = sm + (size_t)NG * NFIELD * 32;
  for (int a = gp; a < 10; a += NG) {
    const int node = live ? A.conn_soa[(size_t)a * A.ne_pad + e] : 0;
    const double *xk = A.x + 3 * (size_t)node, *Xk = A.X0 + 3 * (size_t)node;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      coords[((a * 6 + d) * 32) + lane] = xk[d];
      coords[((a * 6 + 3 + d) * 32) + lane] = Xk[d];
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

    // g[i][a] = sum_k J^-1[i][k] dN[k][a]       (:714-718)
    double g[3][10];
#pragma unroll
    for (int a = 0; a < 10; ++a) {
      const double d0 = c_tab.dN[gp][0][a], d1 = c_tab.dN[gp][1][a], d2 = c_tab.dN[gp][2][a];
#pragma unroll
      for (int i = 0; i < 3; ++i) g[i][a] = Ji[i][0] * d0 + Ji[i][1] * d1 + Ji[i][2] * d2;
    }

    // F^-1[i][j] = sum_k g[j][k] X0_k,i, then invert  (:1141-1152)
    double Fi[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const double X0 = coords[(k * 6 + 3) * 32 + lane], X1 = coords[(k * 6 + 4) * 32 + lane],
                   X2 = coords[(k * 6 + 5) * 32 + lane];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Fi[0][j] = fma(g[j][k], X0, Fi[0][j]);
        Fi[1][j] = fma(g[j][k], X1, Fi[1][j]);
        Fi[2][j] = fma(g[j][k], X2, Fi[2][j]);
      }
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

    // hand over to phase B: g, t = (mu' I + sigma) g scaled by wd, and the scaled coefficients
    const double wd = ok ? c_tab.w[gp] * fabs(detJ) : 0.0;  // fabs: fea_solver.c:958,1047,1104
    const double lw = lam1 * wd, mw = mu1 * wd;
#pragma unroll
    for (int a = 0; a < 10; ++a) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double sg = S[i][0] * g[0][a] + S[i][1] * g[1][a] + S[i][2] * g[2][a];
        my[(i * 10 + a) * 32] = ok ? g[i][a] : 0.0;
        my[(30 + i * 10 + a) * 32] = ok ? fma(wd, sg, mw * g[i][a]) : 0.0;
      }
    }
    my[60 * 32] = lw;
    my[61 * 32] = mw;
  }
  __syncthreads();

  // ------------------------------ phase B ------------------------------------
  const double *col = sm + lane;  // field f of Gauss point q at col[(q*NFIELD + f)*32]
#define FLD(q, f) col[((q)*NFIELD + (f)) * 32]

  if (WITH_R) {
    // R_e[a][i] = -sum_g wd (sigma g_a)_i = -sum_g (t_ai - mu' g_ai)   (fea_solver.c:1094-1109)
    for (int idx = gp; idx < 30; idx += NG) {
      const int a = idx / 3, i = idx - 3 * a;
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < NG; ++q)
        acc += FLD(q, 30 + i * 10 + a) - FLD(q, 61) * FLD(q, i * 10 + a);
      if (live) A.Re[(size_t)idx * A.ne_pad + e] = -acc;
    }
  }

  if (WITH_K) {
    // Each thread builds two consecutive blocks (a,b), (a,b+1) of its element -- consecutive in
    // the packed upper triangle, i.e. 144 contiguous bytes of K_e staging -- then the warp
    // transposes them through a padded tile so its stores walk those 144-byte chunks with
    // consecutive lanes (8-byte stores at a 3960-byte lane stride cost 27 L2 sectors per
    // request in v1; see profiles/r1_v1_ncu_full_summary.md).
    double *tile = sm + (size_t)NG * NFIELD * 32 + (size_t)gp * 32 * TILE_LD;
    const int e0 = blockIdx.x * ELEMS_PER_CTA;
    // rows a and 9-a of the upper triangle hold 11 blocks together: one such pair per warp
    // when NG == 5 (pairs are dealt round-robin otherwise)
    for (int pr = gp; pr < 5; pr += NG)
      for (int half = 0; half < 2; ++half) {
        const int a = half ? 9 - pr : pr;
        double ga[NG][3], ua[NG][3], va[NG][3];
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          const double lw = FLD(q, 60), mw = FLD(q, 61);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            ga[q][i] = FLD(q, i * 10 + a);
            ua[q][i] = lw * ga[q][i];
            va[q][i] = mw * ga[q][i];
          }
        }
        for (int b = a; b < 10; b += 2) {
          const bool two = b + 1 < 10;   // warp-uniform
          double k0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, k1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            {
              const double gb[3] = {FLD(q, b), FLD(q, 10 + b), FLD(q, 20 + b)};
              const double s = ga[q][0] * FLD(q, 30 + b) + ga[q][1] * FLD(q, 40 + b) + ga[q][2] * FLD(q, 50 + b);
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  k0[3 * i + j] = fma(ua[q][i], gb[j], fma(va[q][j], gb[i], k0[3 * i + j]));
              k0[0] += s;
              k0[4] += s;
              k0[8] += s;
            }
            if (two) {
              const int b1 = b + 1;
              const double gb[3] = {FLD(q, b1), FLD(q, 10 + b1), FLD(q, 20 + b1)};
              const double s = ga[q][0] * FLD(q, 30 + b1) + ga[q][1] * FLD(q, 40 + b1) + ga[q][2] * FLD(q, 50 + b1);
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  k1[3 * i + j] = fma(ua[q][i], gb[j], fma(va[q][j], gb[i], k1[3 * i + j]));
              k1[0] += s;
              k1[4] += s;
              k1[8] += s;
            }
          }
#pragma unroll
          for (int c = 0; c < 9; ++c) {
            tile[lane * TILE_LD + c] = k0[c];
            tile[lane * TILE_LD + 9 + c] = k1[c];
          }
          __syncwarp();
          const int tri = a * 10 - (a * (a - 1)) / 2 + (b - a);
          const int nd = two ? 18 : 9;
          for (int it = 0; it < nd; ++it) {
            const int f = it * 32 + lane;
            const int le = two ? f / 18 : f / 9;
            const int cc = f - le * nd;
            if (e0 + le < A.n_elems) A.Ke[((size_t)(e0 + le) * 55 + tri) * 9 + cc] = tile[le * TILE_LD + cc];
          }
          __syncwarp();
        }
      }
  }
#undef FLD
}

}  // namespace fea
