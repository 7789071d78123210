/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- see oracle_fea.h.
 *
 * CPU restatement of the reference's finite-strain hot path.  Every function
 * names the reference lines it follows; loop nests and the order of floating
 * point operations are kept so that, compiled without FMA contraction, the
 * numbers match the reference-compiled objects (oracle/_ref) bit for bit on
 * the element arithmetic.  This file is a checker: nothing under
 * fea-large_b200/ links or calls it.
 */
#define _POSIX_C_SOURCE 199309L
#include "oracle_fea.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NEN 10 /* nodes per element, TETRAHEDRA10 */
#define DELTA(i, j) ((i) == (j) ? 1 : 0)

/* ------------------------------------------------------------------ */
/* tables                                                               */

/* fea_solver.c:32-54.  The 4-point abscissae are the reference's 8-digit
 * literals; weights already carry the tetrahedron's 1/6. */
void orc_gauss_table(int count, double *out) {
  int g, k;
  if (count == 4) {
    const double a = 0.58541020, b = 0.13819660, w = (1 / 4.) / 6.;
    const double tab[4][4] = {{w, a, b, b}, {w, b, a, b}, {w, b, b, a}, {w, b, b, b}};
    for (g = 0; g < 4; ++g) for (k = 0; k < 4; ++k) out[4 * g + k] = tab[g][k];
  } else {
    const double wc = (-4 / 5.) / 6., w = (9 / 20.) / 6.;
    const double tab[5][4] = {{wc, 1 / 4., 1 / 4., 1 / 4.},
                              {w, 1 / 2., 1 / 6., 1 / 6.},
                              {w, 1 / 6., 1 / 2., 1 / 6.},
                              {w, 1 / 6., 1 / 6., 1 / 2.},
                              {w, 1 / 6., 1 / 6., 1 / 6.}};
    for (g = 0; g < 5; ++g) for (k = 0; k < 4; ++k) out[4 * g + k] = tab[g][k];
  }
}

/* fea_solver.c:1287-1361: node 0 <-> (1-r-s-t), 1 <-> r, 2 <-> s, 3 <-> t,
 * mid-sides 4=(0,1) 5=(1,2) 6=(0,2) 7=(0,3) 8=(1,3) 9=(2,3). */
void orc_shape_functions(double r, double s, double t, double *N, double *dN) {
  double *dr = dN, *ds = dN + NEN, *dt = dN + 2 * NEN;
  N[0] = (2 * (1 - r - s - t) - 1) * (1 - r - s - t);
  N[1] = (2 * r - 1) * r;
  N[2] = (2 * s - 1) * s;
  N[3] = (2 * t - 1) * t;
  N[4] = 4 * r * (1 - r - s - t);
  N[5] = 4 * r * s;
  N[6] = 4 * s * (1 - r - s - t);
  N[7] = 4 * t * (1 - r - s - t);
  N[8] = 4 * r * t;
  N[9] = 4 * s * t;
  /* d/dr :1306-1323 */
  dr[0] = 4 * t + 4 * s + 4 * r - 3; dr[1] = 4 * r - 1; dr[2] = 0; dr[3] = 0;
  dr[4] = -4 * t - 4 * s - 8 * r + 4; dr[5] = 4 * s; dr[6] = -4 * s;
  dr[7] = -4 * t; dr[8] = 4 * t; dr[9] = 0;
  /* d/ds :1325-1342 */
  ds[0] = 4 * t + 4 * s + 4 * r - 3; ds[1] = 0; ds[2] = 4 * s - 1; ds[3] = 0;
  ds[4] = -4 * r; ds[5] = 4 * r; ds[6] = -4 * t - 8 * s - 4 * r + 4;
  ds[7] = -4 * t; ds[8] = 0; ds[9] = 4 * t;
  /* d/dt :1344-1361 */
  dt[0] = 4 * t + 4 * s + 4 * r - 3; dt[1] = 0; dt[2] = 0; dt[3] = 4 * t - 1;
  dt[4] = -4 * r; dt[5] = 0; dt[6] = -4 * s; dt[7] = -8 * t - 4 * s - 4 * r + 4;
  dt[8] = 4 * r; dt[9] = 4 * s;
}

/* ------------------------------------------------------------------ */
/* dense 3x3 (row-major 9 doubles)                                      */

/* dense_matrix.c:25-32 */
double orc_det3(const double *m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* dense_matrix.c:34-60; EQUAL(det,0) (defines.h:50) is true only for det==0 */
int orc_inv3(double *m, double *det) {
  double c[9];
  int k;
  *det = orc_det3(m);
  if (fabs(*det - 0.0) <= fmax(fabs(*det), fabs(0.0)) * 2.2204460492503131e-16)
    return 0;
  c[0] = (m[4] * m[8] - m[5] * m[7]) / (*det);
  c[1] = (m[2] * m[7] - m[1] * m[8]) / (*det);
  c[2] = (m[1] * m[5] - m[2] * m[4]) / (*det);
  c[3] = (m[5] * m[6] - m[3] * m[8]) / (*det);
  c[4] = (m[0] * m[8] - m[2] * m[6]) / (*det);
  c[5] = (m[2] * m[3] - m[0] * m[5]) / (*det);
  c[6] = (m[3] * m[7] - m[4] * m[6]) / (*det);
  c[7] = (m[1] * m[6] - m[0] * m[7]) / (*det);
  c[8] = (m[0] * m[4] - m[1] * m[3]) / (*det);
  for (k = 0; k < 9; ++k) m[k] = c[k];
  return 1;
}

/* dense_matrix.c:62-76 */
void orc_mul3(const double *A, const double *B, double *R) {
  int i, j, k;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double sum = 0.0;
      for (k = 0; k < 3; ++k) sum += A[3 * i + k] * B[3 * k + j];
      R[3 * i + j] = sum;
    }
}
/* dense_matrix.c:79-93 */
void orc_mul3_tn(const double *A, const double *B, double *R) {
  int i, j, k;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double sum = 0.0;
      for (k = 0; k < 3; ++k) sum += A[3 * k + i] * B[3 * k + j];
      R[3 * i + j] = sum;
    }
}
/* dense_matrix.c:96-110 */
void orc_mul3_nt(const double *A, const double *B, double *R) {
  int i, j, k;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double sum = 0.0;
      for (k = 0; k < 3; ++k) sum += A[3 * i + k] * B[3 * j + k];
      R[3 * i + j] = sum;
    }
}

/* ------------------------------------------------------------------ */
/* per (element, Gauss point)                                           */

/* fea_solver.c:672-719: J = dN . x, in-place inverse, g = J^-1 . dN */
int orc_shape_gradients(const double *dN, const double *xe, double *g, double *detJ) {
  double J[9];
  int i, j, k;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (k = 0; k < NEN; ++k) acc += dN[NEN * i + k] * xe[3 * k + j];
      J[3 * i + j] = acc;
    }
  if (!orc_inv3(J, detJ)) return 0;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < NEN; ++j) {
      double acc = 0.0;
      for (k = 0; k < 3; ++k) acc += J[3 * i + k] * dN[NEN * k + j];
      g[NEN * i + j] = acc;
    }
  return 1;
}

/* fea_solver.c:1141-1152: F^-1[i][j] = sum_k g[j][k] X0[k][i], then invert
 * (a failed inversion is ignored by the reference and leaves F^-1 in place) */
void orc_graddef(const double *g, const double *X0e, double *F) {
  int i, j, k;
  double det;
  for (i = 0; i < 3; ++i)
    for (j = 0; j < 3; ++j) {
      double acc = 0;
      for (k = 0; k < NEN; ++k) acc += g[NEN * j + k] * X0e[3 * k + i];
      F[3 * i + j] = acc;
    }
  orc_inv3(F, &det);
}

/* fea_model.c:26-77 (A5 = St.Venant-Kirchhoff pushed forward) and :79-107 */
void orc_stress(int model, double lambda, double mu, const double *F, double *S) {
  int i, j, k;
  if (model == ORC_MODEL_A5) {
    double C[9], G[9], Sn[9], detF, I1 = 0;
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) {
        G[3 * i + j] = 0;
        for (k = 0; k < 3; ++k) G[3 * i + j] += F[3 * k + i] * F[3 * k + j];
      }
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) C[3 * i + j] = 0.5 * (G[3 * i + j] - DELTA(i, j));
    for (i = 0; i < 3; ++i) I1 += C[3 * i + i];
    detF = orc_det3(F);
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j)
        Sn[3 * i + j] = (lambda * I1 * DELTA(i, j) + 2 * mu * C[3 * i + j]) / detF;
    orc_mul3(F, Sn, C);
    orc_mul3_nt(C, F, S);
  } else {
    double B[9], J = orc_det3(F);
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j) {
        B[3 * i + j] = 0;
        for (k = 0; k < 3; ++k) B[3 * i + j] += F[3 * i + k] * F[3 * j + k];
      }
    for (i = 0; i < 3; ++i)
      for (j = 0; j < 3; ++j)
        S[3 * i + j] = mu * (B[3 * i + j] - DELTA(i, j)) / J +
                       lambda * log(J) * DELTA(i, j) / J;
  }
}

/* fea_model.c:110-127 (A5: material SVK tensor / det F, NOT pushed forward)
 * and :129-148 (NH: stored without minor symmetry; symmetrised by the caller) */
void orc_ctensor(int model, double lambda, double mu, const double *F, double *c) {
  int i, j, k, l;
  double J = orc_det3(F);
  if (model == ORC_MODEL_A5) {
    for (i = 0; i < 3; ++i) for (j = 0; j < 3; ++j)
      for (k = 0; k < 3; ++k) for (l = 0; l < 3; ++l)
        c[((i * 3 + j) * 3 + k) * 3 + l] =
            (lambda * DELTA(i, j) * DELTA(k, l) + mu * DELTA(i, k) * DELTA(j, l) +
             mu * DELTA(i, l) * DELTA(j, k)) / J;
  } else {
    double lambda1 = lambda / J, mu1 = (mu - lambda * log(J)) / J;
    for (i = 0; i < 3; ++i) for (j = 0; j < 3; ++j)
      for (k = 0; k < 3; ++k) for (l = 0; l < 3; ++l)
        c[((i * 3 + j) * 3 + k) * 3 + l] =
            lambda1 * DELTA(i, j) * DELTA(k, l) + 2 * mu1 * DELTA(i, k) * DELTA(j, l);
  }
}

/* ------------------------------------------------------------------ */
/* solver object                                                        */

struct orc_solver {
  int n_nodes, n_elems, n_presc, ng, model;
  double lambda, mu;
  double *X0, *x;          /* [n][3] reference / current coordinates */
  int *conn;               /* [e][10] */
  int *presc_node, *presc_type;
  double *presc_vals;      /* [p][3] */
  double gauss[5][4];
  double N[5][NEN], dN[5][3 * NEN];
  double *g;               /* [e][g][3][10] current-configuration gradients */
  double *detJ;            /* [e][g] */
  unsigned char *gvalid;   /* [e][g] 0 when J was singular (element skipped) */
  double *F, *S;           /* [e][g][9] */
  int n;                   /* 3 * n_nodes */
  int *rowptr, *colidx;    /* scalar CSR, columns ascending, full 3x3 blocks */
  double *K, *Ksaved;
  double *R, *u;
};

static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* sparsity pattern = every (3a+i, 3b+j) for node pairs sharing an element,
 * explicit zeros included (fea_solver.c:1035-1058 adds them) */
static void build_pattern(orc_solver *s) {
  int nn = s->n_nodes, e, a, i, j, k;
  int *cnt = (int *)calloc((size_t)nn + 1, sizeof(int));
  int *adj, *fill, *nbr_ptr, *nbr;
  long total = 0;
  for (e = 0; e < s->n_elems; ++e)
    for (a = 0; a < NEN; ++a) cnt[s->conn[NEN * e + a] + 1]++;
  for (i = 0; i < nn; ++i) cnt[i + 1] += cnt[i];
  adj = (int *)malloc(sizeof(int) * (size_t)cnt[nn]);
  fill = (int *)malloc(sizeof(int) * (size_t)nn);
  memcpy(fill, cnt, sizeof(int) * (size_t)nn);
  for (e = 0; e < s->n_elems; ++e)
    for (a = 0; a < NEN; ++a) adj[fill[s->conn[NEN * e + a]]++] = e;
  nbr_ptr = (int *)calloc((size_t)nn + 1, sizeof(int));
  nbr = NULL;
  {
    int cap = 0, used = 0;
    int *tmp = (int *)malloc(sizeof(int) * 4096);
    int tmpcap = 4096;
    for (i = 0; i < nn; ++i) {
      int m = 0, u;
      int need = (cnt[i + 1] - cnt[i]) * NEN;
      if (need > tmpcap) { tmpcap = need; tmp = (int *)realloc(tmp, sizeof(int) * (size_t)tmpcap); }
      for (k = cnt[i]; k < cnt[i + 1]; ++k)
        for (a = 0; a < NEN; ++a) tmp[m++] = s->conn[NEN * adj[k] + a];
      qsort(tmp, (size_t)m, sizeof(int), cmp_int);
      u = 0;
      for (j = 0; j < m; ++j) if (j == 0 || tmp[j] != tmp[j - 1]) tmp[u++] = tmp[j];
      if (used + u > cap) { cap = (used + u) * 2 + 1024; nbr = (int *)realloc(nbr, sizeof(int) * (size_t)cap); }
      memcpy(nbr + used, tmp, sizeof(int) * (size_t)u);
      used += u;
      nbr_ptr[i + 1] = used;
    }
    free(tmp);
  }
  s->rowptr = (int *)malloc(sizeof(int) * ((size_t)s->n + 1));
  s->rowptr[0] = 0;
  for (i = 0; i < nn; ++i)
    for (j = 0; j < 3; ++j) {
      total += 3L * (nbr_ptr[i + 1] - nbr_ptr[i]);
      s->rowptr[3 * i + j + 1] = (int)total;
    }
  s->colidx = (int *)malloc(sizeof(int) * (size_t)total);
  for (i = 0; i < nn; ++i)
    for (j = 0; j < 3; ++j) {
      int *dst = s->colidx + s->rowptr[3 * i + j];
      for (k = nbr_ptr[i]; k < nbr_ptr[i + 1]; ++k) {
        *dst++ = 3 * nbr[k]; *dst++ = 3 * nbr[k] + 1; *dst++ = 3 * nbr[k] + 2;
      }
    }
  s->K = (double *)calloc((size_t)total, sizeof(double));
  s->Ksaved = (double *)calloc((size_t)total, sizeof(double));
  free(cnt); free(adj); free(fill); free(nbr_ptr); free(nbr);
}

static int csr_find(const orc_solver *s, int row, int col) {
  int lo = s->rowptr[row], hi = s->rowptr[row + 1] - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1, v = s->colidx[mid];
    if (v == col) return mid;
    if (v < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

orc_solver *orc_create(int n_nodes, const double *nodes, int n_elems,
                       const int *conn, int n_presc, const int *presc_node,
                       const int *presc_type, const double *presc_vals,
                       int model, double lambda, double mu, int gauss_count) {
  orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
  int g;
  size_t eg;
  s->n_nodes = n_nodes; s->n_elems = n_elems; s->n_presc = n_presc;
  s->ng = gauss_count; s->model = model; s->lambda = lambda; s->mu = mu;
  s->n = 3 * n_nodes;
  s->X0 = (double *)malloc(sizeof(double) * 3 * (size_t)n_nodes);
  s->x = (double *)malloc(sizeof(double) * 3 * (size_t)n_nodes);
  memcpy(s->X0, nodes, sizeof(double) * 3 * (size_t)n_nodes);
  memcpy(s->x, nodes, sizeof(double) * 3 * (size_t)n_nodes); /* fea_solver.c:400 */
  s->conn = (int *)malloc(sizeof(int) * NEN * (size_t)n_elems);
  memcpy(s->conn, conn, sizeof(int) * NEN * (size_t)n_elems);
  s->presc_node = (int *)malloc(sizeof(int) * (size_t)(n_presc + 1));
  s->presc_type = (int *)malloc(sizeof(int) * (size_t)(n_presc + 1));
  s->presc_vals = (double *)malloc(sizeof(double) * 3 * (size_t)(n_presc + 1));
  if (n_presc) {
    memcpy(s->presc_node, presc_node, sizeof(int) * (size_t)n_presc);
    memcpy(s->presc_type, presc_type, sizeof(int) * (size_t)n_presc);
    memcpy(s->presc_vals, presc_vals, sizeof(double) * 3 * (size_t)n_presc);
  }
  orc_gauss_table(gauss_count, &s->gauss[0][0]);
  for (g = 0; g < gauss_count; ++g) /* fea_solver.c:503-535 */
    orc_shape_functions(s->gauss[g][1], s->gauss[g][2], s->gauss[g][3], s->N[g], s->dN[g]);
  eg = (size_t)n_elems * (size_t)gauss_count;
  s->g = (double *)calloc(eg * 30, sizeof(double));
  s->detJ = (double *)calloc(eg, sizeof(double));
  s->gvalid = (unsigned char *)calloc(eg, 1);
  s->F = (double *)calloc(eg * 9, sizeof(double)); /* zero-initialised, :434-435 */
  s->S = (double *)calloc(eg * 9, sizeof(double));
  s->R = (double *)calloc((size_t)s->n, sizeof(double));
  s->u = (double *)calloc((size_t)s->n, sizeof(double));
  build_pattern(s);
  return s;
}

void orc_destroy(orc_solver *s) {
  if (!s) return;
  free(s->X0); free(s->x); free(s->conn); free(s->presc_node); free(s->presc_type);
  free(s->presc_vals); free(s->g); free(s->detJ); free(s->gvalid); free(s->F);
  free(s->S); free(s->R); free(s->u); free(s->rowptr); free(s->colidx);
  free(s->K); free(s->Ksaved); free(s);
}

void orc_set_nodes(orc_solver *s, const double *x) {
  memcpy(s->x, x, sizeof(double) * 3 * (size_t)s->n_nodes);
}
void orc_get_nodes(const orc_solver *s, double *x) {
  memcpy(x, s->x, sizeof(double) * 3 * (size_t)s->n_nodes);
}

/* fea_solver.c:1205-1242: walk prescribed nodes, pick DOFs by the type bitmask
 * (1 = x, 2 = y, 4 = z; enum fea_solver.h:74-83), call `apply` per DOF */
typedef void (*orc_bc_fn)(orc_solver *s, int index, double value);

static void bc_walk(orc_solver *s, orc_bc_fn apply, double lambda) {
  int p, d;
  for (p = 0; p < s->n_presc; ++p) {
    int type = s->presc_type[p];
    for (d = 0; d < 3; ++d)
      if (type >= 0 && type <= 7 && (type & (1 << d)))
        apply(s, 3 * s->presc_node[p] + d, s->presc_vals[3 * p + d] * lambda);
  }
}

/* fea_solver.c:1259-1266 */
static void bc_move_node(orc_solver *s, int index, double value) {
  s->x[index] += value;
}
void orc_apply_increment(orc_solver *s, double lambda) { bc_walk(s, bc_move_node, lambda); }

/* fea_solver.c:787-861 with CURRENT_SHAPE_GRADIENTS: gradients in the current
 * configuration for every (e,g), then F and the Cauchy stress.  A singular J
 * keeps the previous gradients (:808 `if (grads)`). */
void orc_update_state(orc_solver *s) {
  int e, g, a, d;
  for (e = 0; e < s->n_elems; ++e) {
    double xe[30], X0e[30];
    for (a = 0; a < NEN; ++a)
      for (d = 0; d < 3; ++d) {
        xe[3 * a + d] = s->x[3 * s->conn[NEN * e + a] + d];
        X0e[3 * a + d] = s->X0[3 * s->conn[NEN * e + a] + d];
      }
    for (g = 0; g < s->ng; ++g) {
      size_t eg = (size_t)e * s->ng + g;
      double gnew[30], det;
      if (orc_shape_gradients(s->dN[g], xe, gnew, &det)) {
        memcpy(s->g + eg * 30, gnew, sizeof(gnew));
        s->detJ[eg] = det;
        s->gvalid[eg] = 1;
      }
    }
    for (g = 0; g < s->ng; ++g) {
      size_t eg = (size_t)e * s->ng + g;
      orc_graddef(s->g + eg * 30, X0e, s->F + eg * 9);
      orc_stress(s->model, s->lambda, s->mu, s->F + eg * 9, s->S + eg * 9);
    }
  }
}

void orc_get_state(const orc_solver *s, double *F, double *S) {
  size_t n = (size_t)s->n_elems * s->ng * 9;
  memcpy(F, s->F, sizeof(double) * n);
  memcpy(S, s->S, sizeof(double) * n);
}
void orc_get_gradients(const orc_solver *s, double *g, double *detJ) {
  size_t n = (size_t)s->n_elems * s->ng;
  memcpy(g, s->g, sizeof(double) * n * 30);
  memcpy(detJ, s->detJ, sizeof(double) * n);
}

/* One element's scatter stream, in the reference's order: constitutive part
 * for all Gauss points (fea_solver.c:918-973), then the initial-stress part
 * (:1014-1062).  `sink(ctx, I_local, J_local, a, b, value)` receives every
 * term the reference hands to sp_matrix_element_add. */
typedef void (*orc_sink)(void *ctx, int li, int lj, double v);

static void element_terms(const orc_solver *s, int e, int part, orc_sink sink, void *ctx) {
  int gp, a, b, i, j, k, l;
  if (part == 0 || part == 1)
    for (gp = 0; gp < s->ng; ++gp) {
      size_t eg = (size_t)e * s->ng + gp;
      const double *g = s->g + eg * 30;
      double c[81];
      orc_ctensor(s->model, s->lambda, s->mu, s->F + eg * 9, c);
      if (!s->gvalid[eg]) continue;
      for (a = 0; a < NEN; ++a)
        for (b = 0; b < NEN; ++b)
          for (i = 0; i < 3; ++i)
            for (j = 0; j < 3; ++j) {
              double sum = 0.0;
              for (k = 0; k < 3; ++k)
                for (l = 0; l < 3; ++l) {
                  double cikjl = (c[((i * 3 + k) * 3 + j) * 3 + l] + c[((i * 3 + k) * 3 + l) * 3 + j] +
                                  c[((k * 3 + i) * 3 + j) * 3 + l] + c[((k * 3 + i) * 3 + l) * 3 + j]) / 4.;
                  sum += g[NEN * k + a] * cikjl * g[NEN * l + b];
                }
              sum *= fabs(s->detJ[eg]);
              sum *= s->gauss[gp][0];
              sink(ctx, 3 * a + i, 3 * b + j, sum);
            }
    }
  if (part == 0 || part == 2)
    for (gp = 0; gp < s->ng; ++gp) {
      size_t eg = (size_t)e * s->ng + gp;
      const double *g = s->g + eg * 30;
      const double *sg = s->S + eg * 9;
      if (!s->gvalid[eg]) continue;
      for (a = 0; a < NEN; ++a)
        for (b = 0; b < NEN; ++b)
          for (i = 0; i < 3; ++i)
            for (j = 0; j < 3; ++j) {
              double sum = 0.0;
              for (k = 0; k < 3; ++k)
                for (l = 0; l < 3; ++l)
                  sum += g[NEN * k + a] * sg[3 * k + l] * g[NEN * l + b] * DELTA(i, j);
              sum *= fabs(s->detJ[eg]);
              sum *= s->gauss[gp][0];
              sink(ctx, 3 * a + i, 3 * b + j, sum);
            }
    }
}

static void sink_dense(void *ctx, int li, int lj, double v) {
  ((double *)ctx)[30 * li + lj] += v;
}
void orc_element_matrix(const orc_solver *s, int element, double *ke, int part) {
  memset(ke, 0, sizeof(double) * 900);
  element_terms(s, element, part, sink_dense, ke);
}

typedef struct { orc_solver *s; int e; int pos[900]; } sink_csr_ctx;
static void sink_csr(void *ctx, int li, int lj, double v) {
  sink_csr_ctx *c = (sink_csr_ctx *)ctx;
  c->s->K[c->pos[30 * li + lj]] += v;
}

/* fea_solver.c:873-883: clear, then element-major accumulation */
void orc_assemble_stiffness(orc_solver *s) {
  int e, a, b, i, j;
  sink_csr_ctx ctx;
  memset(s->K, 0, sizeof(double) * (size_t)s->rowptr[s->n]);
  ctx.s = s;
  for (e = 0; e < s->n_elems; ++e) {
    ctx.e = e;
    for (a = 0; a < NEN; ++a)
      for (i = 0; i < 3; ++i)
        for (b = 0; b < NEN; ++b)
          for (j = 0; j < 3; ++j)
            ctx.pos[30 * (3 * a + i) + 3 * b + j] =
                csr_find(s, 3 * s->conn[NEN * e + a] + i, 3 * s->conn[NEN * e + b] + j);
    element_terms(s, e, 0, sink_csr, &ctx);
  }
}

/* fea_solver.c:863-870 and :1072-1114 */
void orc_assemble_residual(orc_solver *s) {
  int e, gp, a, i, j;
  memset(s->R, 0, sizeof(double) * (size_t)s->n);
  for (e = 0; e < s->n_elems; ++e)
    for (gp = 0; gp < s->ng; ++gp) {
      size_t eg = (size_t)e * s->ng + gp;
      const double *g = s->g + eg * 30;
      const double *sg = s->S + eg * 9;
      if (!s->gvalid[eg]) continue;
      for (a = 0; a < NEN; ++a)
        for (i = 0; i < 3; ++i) {
          double sum = 0.0;
          for (j = 0; j < 3; ++j) sum += sg[3 * i + j] * g[NEN * j + a];
          sum *= fabs(s->detJ[eg]);
          sum *= s->gauss[gp][0];
          s->R[3 * s->conn[NEN * e + a] + i] += -sum;
        }
    }
}

/* fea_solver.c:1244-1257 + libspmatrix's sp_matrix_cross_cancellation as its
 * call site defines it: move the column's contribution to the RHS, zero the
 * row and column keeping the diagonal, RHS[index] = diag * presc */
static void bc_cancel(orc_solver *s, int index, double presc) {
  int k;
  double diag = 0.0;
  for (k = s->rowptr[index]; k < s->rowptr[index + 1]; ++k) {
    int r = s->colidx[k];                 /* symmetric pattern: row r has column `index` */
    int p = csr_find(s, r, index);        /* entry K[r][index] (the stored column) */
    s->R[r] -= s->K[p] * presc;
  }
  for (k = s->rowptr[index]; k < s->rowptr[index + 1]; ++k) {
    int r = s->colidx[k];
    if (r == index) { diag = s->K[k]; continue; }
    s->K[csr_find(s, r, index)] = 0.0;
    s->K[k] = 0.0;
  }
  s->R[index] = diag * presc;
}
void orc_apply_bc(orc_solver *s, double lambda) { bc_walk(s, bc_cancel, lambda); }

long orc_nnz(const orc_solver *s) { return s->rowptr[s->n]; }
void orc_get_csr(const orc_solver *s, int *rowptr, int *colidx, double *vals) {
  memcpy(rowptr, s->rowptr, sizeof(int) * ((size_t)s->n + 1));
  memcpy(colidx, s->colidx, sizeof(int) * (size_t)s->rowptr[s->n]);
  memcpy(vals, s->K, sizeof(double) * (size_t)s->rowptr[s->n]);
}
void orc_get_forces(const orc_solver *s, double *R) { memcpy(R, s->R, sizeof(double) * (size_t)s->n); }
void orc_set_forces(orc_solver *s, const double *R) { memcpy(s->R, R, sizeof(double) * (size_t)s->n); }
void orc_get_solution(const orc_solver *s, double *u) { memcpy(u, s->u, sizeof(double) * (size_t)s->n); }

/* Stand-in for the libspmatrix solve behind solver_solve_slae
 * (fea_solver.c:300-321): Jacobi-PCG from x0 = 0 on K u = R, relative
 * residual stop plus a stagnation guard at the rounding floor. */
int orc_solve_slae(orc_solver *s, double rel_tol, int max_iter) {
  int n = s->n, i, k, it = 0, stall = 0;
  double *r = (double *)malloc(sizeof(double) * (size_t)n * 5);
  double *z = r + n, *p = z + n, *q = p + n, *dinv = q + n;
  double bb = 0.0, rz = 0.0, rr, best = 1e300;
  for (i = 0; i < n; ++i) {
    int d = csr_find(s, i, i);
    double dv = d >= 0 ? s->K[d] : 1.0;
    dinv[i] = dv != 0.0 ? 1.0 / dv : 1.0;
    s->u[i] = 0.0;
    r[i] = s->R[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += r[i] * r[i];
    rz += r[i] * z[i];
  }
  rr = bb;
  if (bb > 0.0)
    for (it = 0; it < max_iter; ++it) {
      double pq = 0.0, alpha, rz_new = 0.0, beta;
      if (sqrt(rr) <= rel_tol * sqrt(bb)) break;
      for (i = 0; i < n; ++i) {
        double acc = 0.0;
        for (k = s->rowptr[i]; k < s->rowptr[i + 1]; ++k) acc += s->K[k] * p[s->colidx[k]];
        q[i] = acc;
        pq += p[i] * acc;
      }
      alpha = rz / pq;
      rr = 0.0;
      for (i = 0; i < n; ++i) {
        s->u[i] += alpha * p[i];
        r[i] -= alpha * q[i];
        z[i] = dinv[i] * r[i];
        rz_new += r[i] * z[i];
        rr += r[i] * r[i];
      }
      beta = rz_new / rz;
      rz = rz_new;
      for (i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
      if (rr < best * 0.999) { best = rr; stall = 0; }
      else if (++stall > 200) break;
    }
  free(r);
  return it;
}

/* fea_solver.c:1270-1279 */
void orc_update_with_solution(orc_solver *s) {
  int i;
  for (i = 0; i < s->n; ++i) s->x[i] += s->u[i];
}
/* dense_matrix.c:16-23 as used at fea_solver.c:208 */
double orc_dot_forces_solution(const orc_solver *s) {
  double acc = 0;
  int i;
  for (i = 0; i < s->n; ++i) acc += s->R[i] * s->u[i];
  return acc;
}

/* fea_solver.c:163-236 */
int orc_newton_solve(orc_solver *s, int load_increments, double desired_tol,
                     int modified_newton, int max_newton, double lin_tol,
                     int lin_max_iter, double *trace_u, double *trace_tol,
                     int trace_cap, int *n_trace) {
  int step, nt = 0;
  size_t nnz = (size_t)s->rowptr[s->n];
  for (step = 0; step < load_increments; ++step) {
    int it = 0;
    double tol;
    orc_apply_increment(s, 1);
    orc_update_state(s);
    orc_assemble_stiffness(s);
    memcpy(s->Ksaved, s->K, sizeof(double) * nnz);
    do {
      it++;
      orc_assemble_residual(s);
      if (modified_newton) memcpy(s->K, s->Ksaved, sizeof(double) * nnz);
      else orc_assemble_stiffness(s);
      orc_apply_bc(s, 0);
      orc_solve_slae(s, lin_tol, lin_max_iter);
      tol = orc_dot_forces_solution(s);
      if (nt < trace_cap) {
        if (trace_u) memcpy(trace_u + (size_t)nt * s->n, s->u, sizeof(double) * (size_t)s->n);
        if (trace_tol) trace_tol[nt] = tol;
      }
      nt++;
      orc_update_with_solution(s);
      orc_update_state(s);
    } while (fabs(tol) > desired_tol && it < max_newton);
    if (it == max_newton) break; /* :225-231: step rolled back, loop exits */
  }
  if (n_trace) *n_trace = nt;
  return step;
}

double orc_time_assembly(orc_solver *s, int reps) {
  struct timespec t0, t1;
  int r;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (r = 0; r < reps; ++r) {
    orc_update_state(s);
    orc_assemble_stiffness(s);
    orc_assemble_residual(s);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
