/* Reader for the reference's .sexp task files (the five models under solver-large/data).
 *
 * The reference walks a tree built by libsexp (sexp_loader.c:249-327); libsexp is not
 * available, and a tree of 17 M node forms would not be welcome anyway, so this is a
 * single-pass recursive-descent reader: node and element rows go straight into flat arrays,
 * every other form is a head symbol followed by `:keyword value` pairs and nested forms.
 * Grammar and keys: SURVEY 8f-2 / sexp_loader.c:32-247.  Keywords and symbols are matched
 * case-insensitively, `;` starts a comment, key order is free.
 */
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "sexp_loader.h"

typedef struct {
  const char *cur, *end;
  char tok[256];
  int failed;
  char why[200];
  fea_task *task;
  fea_solution_params *params;
  nodes_array *nodes;
  elements_array *elements;
  presc_bnd_array *presc;
  int presc_cap;
  int have_solution, have_model_params;
} reader;

enum { T_EOF, T_OPEN, T_CLOSE, T_ATOM };

static void fail(reader *r, const char *why) {
  if (!r->failed) {
    r->failed = 1;
    strncpy(r->why, why, sizeof(r->why) - 1);
  }
}

static int next_token(reader *r) {
  size_t n = 0;
  for (;;) {
    while (r->cur < r->end && isspace((unsigned char)*r->cur)) r->cur++;
    if (r->cur < r->end && *r->cur == ';') {
      while (r->cur < r->end && *r->cur != '\n') r->cur++;
      continue;
    }
    break;
  }
  if (r->cur >= r->end) return T_EOF;
  if (*r->cur == '(') { r->cur++; return T_OPEN; }
  if (*r->cur == ')') { r->cur++; return T_CLOSE; }
  if (*r->cur == '"') {
    r->cur++;
    while (r->cur < r->end && *r->cur != '"') {
      if (n + 1 < sizeof(r->tok)) r->tok[n++] = *r->cur;
      r->cur++;
    }
    if (r->cur < r->end) r->cur++;
  } else {
    while (r->cur < r->end && !isspace((unsigned char)*r->cur) && *r->cur != '(' && *r->cur != ')') {
      if (n + 1 < sizeof(r->tok)) r->tok[n++] = *r->cur;
      r->cur++;
    }
  }
  r->tok[n] = 0;
  return T_ATOM;
}

static int sym_is(const char *tok, const char *name) { return strcasecmp(tok, name) == 0; }

static double to_real(reader *r, const char *tok) {
  char *e;
  double v;
  errno = 0;
  v = strtod(tok, &e);
  if (e == tok || *e) fail(r, "expected a number");
  return v;
}
static int to_int(reader *r, const char *tok) {
  char *e;
  long v = strtol(tok, &e, 10);
  if (e == tok || *e) fail(r, "expected an integer");
  return (int)v;
}

/* (nodes (x y z) ...): sexp_loader.c:169-189 */
static void read_nodes(reader *r) {
  int cap = 1024, n = 0, t;
  real *flat = (real *)malloc(sizeof(real) * 3 * (size_t)cap);
  while ((t = next_token(r)) == T_OPEN) {
    int k = 0;
    if (n == cap) { cap *= 2; flat = (real *)realloc(flat, sizeof(real) * 3 * (size_t)cap); }
    while ((t = next_token(r)) == T_ATOM) {
      if (k < 3) flat[3 * (size_t)n + k] = to_real(r, r->tok);
      ++k;
    }
    if (k != 3 || t != T_CLOSE) fail(r, "a node must be (x y z)");   /* assert at :183 */
    ++n;
  }
  if (t != T_CLOSE) fail(r, "unterminated nodes form");
  if (!r->failed && nodes_array_reserve(r->nodes, n))
    memcpy(r->nodes->nodes[0], flat, sizeof(real) * 3 * (size_t)n);
  free(flat);
}

/* (elements (i0 ... i9) ...): sexp_loader.c:191-212; ids are 0-based (exporter.py:480) */
static void read_elements(reader *r) {
  const int npe = r->params->nodes_per_element;
  int cap = 1024, n = 0, t;
  int *flat = (int *)malloc(sizeof(int) * (size_t)npe * (size_t)cap);
  while ((t = next_token(r)) == T_OPEN) {
    int k = 0;
    if (n == cap) { cap *= 2; flat = (int *)realloc(flat, sizeof(int) * (size_t)npe * (size_t)cap); }
    while ((t = next_token(r)) == T_ATOM) {
      if (k < npe) flat[(size_t)npe * n + k] = to_int(r, r->tok);
      ++k;
    }
    if (k != npe || t != T_CLOSE) fail(r, "element row length differs from :nodes-count");   /* :204 */
    ++n;
  }
  if (t != T_CLOSE) fail(r, "unterminated elements form");
  if (!r->failed && elements_array_reserve(r->elements, n, npe))
    memcpy(r->elements->elements[0], flat, sizeof(int) * (size_t)npe * (size_t)n);
  free(flat);
}

static void read_form(reader *r);

/* generic body: `:key value` pairs handed to `on_key`, nested forms recursed into */
typedef void (*key_fn)(reader *r, const char *key, const char *value);

static void read_body(reader *r, key_fn on_key) {
  int t;
  char key[64];
  while ((t = next_token(r)) != T_CLOSE) {
    if (t == T_EOF) { fail(r, "unexpected end of file"); return; }
    if (t == T_OPEN) { read_form(r); continue; }
    if (r->tok[0] == ':') {
      strncpy(key, r->tok + 1, sizeof(key) - 1);
      key[sizeof(key) - 1] = 0;
      if (next_token(r) != T_ATOM) { fail(r, "keyword without a value"); return; }
      if (on_key) on_key(r, key, r->tok);
    }
  }
}

static void key_model(reader *r, const char *k, const char *v) {   /* :32-53 */
  if (!sym_is(k, "name")) return;
  if (sym_is(v, "A5")) r->task->model.model = MODEL_A5;
  else if (sym_is(v, "COMPRESSIBLE_NEOHOOKEAN")) r->task->model.model = MODEL_COMPRESSIBLE_NEOHOOKEAN;
  else printf("unknown model type '%s'\n", v);
  r->task->model.parameters_count = 2;
}
static void key_model_parameters(reader *r, const char *k, const char *v) {   /* :55-72 */
  if (sym_is(k, "lambda")) { r->task->model.parameters[0] = to_real(r, v); r->have_model_params |= 1; }
  else if (sym_is(k, "mu")) { r->task->model.parameters[1] = to_real(r, v); r->have_model_params |= 2; }
}
static void key_solution(reader *r, const char *k, const char *v) {   /* :74-94 */
  if (sym_is(k, "desired-tolerance")) { r->task->desired_tolerance = to_real(r, v); r->have_solution |= 1; }
  else if (sym_is(k, "task-type")) { if (sym_is(v, "CARTESIAN3D")) r->task->type = CARTESIAN3D; r->have_solution |= 2; }
  else if (sym_is(k, "load-increments-count")) { r->task->load_increments_count = to_int(r, v); r->have_solution |= 4; }
  else if (sym_is(k, "modified-newton")) { r->task->modified_newton = (sym_is(v, "YES") || sym_is(v, "TRUE")) ? TRUE : FALSE; r->have_solution |= 8; }
  else if (sym_is(k, "max-newton-count")) r->task->max_newton_count = to_int(r, v);
}
static void key_slae_solver(reader *r, const char *k, const char *v) {   /* :96-135 */
  if (sym_is(k, "type")) {
    if (sym_is(v, "CG")) r->task->solver_type = CG;
    else if (sym_is(v, "PCG_ILU")) r->task->solver_type = PCG_ILU;
    else if (sym_is(v, "CHOLESKY")) r->task->solver_type = CHOLESKY;
    else printf("unknown solver type '%s'\n", v);
  } else if (sym_is(k, "tolerance")) r->task->solver_tolerance = to_real(r, v);
  else if (sym_is(k, "max-iterations")) r->task->solver_max_iter = to_int(r, v);
}
static void key_element_type(reader *r, const char *k, const char *v) {   /* :138-151 */
  if (sym_is(k, "gauss-nodes-count")) r->params->gauss_nodes_count = to_int(r, v);
  else if (sym_is(k, "nodes-count")) r->params->nodes_per_element = to_int(r, v);
  else if (sym_is(k, "name") && sym_is(v, "TETRAHEDRA10")) r->task->ele_type = TETRAHEDRA10;
}
static void key_line_search(reader *r, const char *k, const char *v) { if (sym_is(k, "max")) r->task->linesearch_max = to_int(r, v); }
static void key_arc_length(reader *r, const char *k, const char *v) { if (sym_is(k, "max")) r->task->arclength_max = to_int(r, v); }

static void key_presc_node(reader *r, const char *k, const char *v) {   /* :214-247 */
  prescribed_bnd_node *p = &r->presc->prescribed_nodes[r->presc->prescribed_nodes_count];
  if (sym_is(k, "node-id")) p->node_number = to_int(r, v);
  else if (sym_is(k, "x")) p->values[0] = to_real(r, v);
  else if (sym_is(k, "y")) p->values[1] = to_real(r, v);
  else if (sym_is(k, "z")) p->values[2] = to_real(r, v);
  else if (sym_is(k, "type")) p->type = (presc_boundary_type)to_int(r, v);
}

static void read_form(reader *r) { /* '(' already consumed */
  int t = next_token(r);
  if (t == T_CLOSE) return;
  if (t != T_ATOM) { if (t == T_OPEN) { read_form(r); read_body(r, NULL); } return; }
  if (sym_is(r->tok, "nodes")) read_nodes(r);
  else if (sym_is(r->tok, "elements")) read_elements(r);
  else if (sym_is(r->tok, "model")) read_body(r, key_model);
  else if (sym_is(r->tok, "model-parameters")) read_body(r, key_model_parameters);
  else if (sym_is(r->tok, "solution")) read_body(r, key_solution);
  else if (sym_is(r->tok, "slae-solver")) {
    /* defaults as sexp_loader.c:100-103 */
    r->task->solver_type = CG;
    r->task->solver_tolerance = MAX_ITERATIVE_TOLERANCE;
    r->task->solver_max_iter = MAX_ITERATIVE_ITERATIONS;
    read_body(r, key_slae_solver);
  } else if (sym_is(r->tok, "element-type")) read_body(r, key_element_type);
  else if (sym_is(r->tok, "line-search")) read_body(r, key_line_search);
  else if (sym_is(r->tok, "arc-length")) read_body(r, key_arc_length);
  else if (sym_is(r->tok, "presc-node")) {
    prescribed_bnd_node *p;
    if (r->presc->prescribed_nodes_count == r->presc_cap) {
      r->presc_cap = r->presc_cap ? 2 * r->presc_cap : 256;
      r->presc->prescribed_nodes = (prescribed_bnd_node *)realloc(
          r->presc->prescribed_nodes, sizeof(prescribed_bnd_node) * (size_t)r->presc_cap);
    }
    p = &r->presc->prescribed_nodes[r->presc->prescribed_nodes_count];
    memset(p, 0, sizeof(*p));
    p->node_number = -1;
    read_body(r, key_presc_node);
    if (p->node_number < 0) fail(r, "presc-node without :node-id");
    r->presc->prescribed_nodes_count++;
  } else {
    read_body(r, NULL); /* task, input-data, geometry, boundary-conditions, prescribed-displacements, unknown */
  }
}

BOOL sexp_data_load(char *filename, fea_task **task, fea_solution_params **fea_params,
                    nodes_array **nodes, elements_array **elements,
                    presc_bnd_array **presc_boundary) {
  /* The file is mapped, not read: a 50 M-DOF model is a few GB of text, the reader walks it once
   * front to back and the kernel pages it in (and drops it) behind the cursor. */
  int fd = open(filename, O_RDONLY);
  struct stat sb;
  size_t size;
  const char *text;
  reader r;
  if (fd < 0) {
    fprintf(stderr, "Error, could not open file %s\n", filename);   /* :291 */
    return FALSE;
  }
  if (fstat(fd, &sb) != 0 || sb.st_size <= 0) {
    close(fd);
    printf("Error: unable to parse SEXP input\n");                  /* :298 */
    return FALSE;
  }
  size = (size_t)sb.st_size;
  text = (const char *)mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (text == (const char *)MAP_FAILED) {
    printf("Error: unable to parse SEXP input\n");
    return FALSE;
  }
  madvise((void *)text, size, MADV_SEQUENTIAL);
  memset(&r, 0, sizeof(r));
  r.cur = text;
  r.end = text + size;
  r.task = fea_task_alloc();
  r.params = fea_solution_params_alloc();
  r.nodes = nodes_array_alloc();
  r.elements = elements_array_alloc();
  r.presc = presc_bnd_array_alloc();
  if (next_token(&r) != T_OPEN || next_token(&r) != T_ATOM || !sym_is(r.tok, "task"))
    fail(&r, "file does not start with (task");
  else
    read_body(&r, NULL);
  munmap((void *)text, size);
  if (!r.failed && r.have_solution != 15) fail(&r, "solution form lacks a mandatory attribute");   /* asserts :78-89 */
  if (!r.failed && r.have_model_params != 3) fail(&r, "model-parameters needs :lambda and :mu");      /* asserts :63,66 */
  if (!r.failed && (r.nodes->nodes_count == 0 || r.elements->elements_count == 0)) fail(&r, "no geometry");
  if (r.failed) {
    printf("Error: unable to parse SEXP input (%s)\n", r.why);
    fea_task_free(r.task);
    fea_solution_params_free(r.params);
    nodes_array_free(r.nodes);
    elements_array_free(r.elements);
    presc_bnd_array_free(r.presc);
    return FALSE;
  }
  *task = r.task;
  *fea_params = r.params;
  *nodes = r.nodes;
  *elements = r.elements;
  *presc_boundary = r.presc;
  return TRUE;
}
