/*
 * qs_oracle.c -- CPU restatement of the reference's hot path (see qs_oracle.h).
 * TEST INFRASTRUCTURE ONLY.
 *
 * Structure mirrors the reference: work is cut into "storage blocks" of
 * g_block_rows tuples, one work order per block, executed by a pool of worker
 * threads (query_execution/Worker.cpp:54-139).  Inside a block everything is
 * column-at-a-time: a predicate yields a TupleIdSequence bitmap, every scalar
 * node yields a materialised vector (storage/StorageBlock.cpp:363-388), and
 * aggregation walks the filtered vector sequentially
 * (ArithmeticBinaryOperators.hpp:714-744).  Block results are merged in block
 * order, which is one of the orders the reference's scheduler can produce.
 */
#define _GNU_SOURCE
#include "qs_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------ configuration */
static int g_workers = 1;
static uint64_t g_block_rows = 65536;

void qso_set_num_workers(int n) { g_workers = n < 1 ? 1 : n; }
int qso_get_num_workers(void) { return g_workers; }
void qso_set_block_rows(uint64_t rows) {
  if (rows < 64) rows = 64;
  g_block_rows = (rows + 63) & ~(uint64_t)63;   /* block bitmaps concatenate word-aligned */
}

/* ------------------------------------------------------------------ bitmaps */
/* utility/BitVector.hpp:934: bit i is TopBit >> (i & 63) of word i >> 6. */
static inline int bm_get(const uint64_t *w, uint64_t i) { return (int)((w[i >> 6] << (i & 63)) >> 63); }
static inline void bm_set(uint64_t *w, uint64_t i) { w[i >> 6] |= 0x8000000000000000ull >> (i & 63); }
static inline uint64_t bm_words(uint64_t bits) { return (bits + 63) / 64; }

/* --------------------------------------------------------------- LIP filter */
uint64_t qso_lip_words(const qso_lip *f) {
  const uint64_t bits = f->kind == QS_LIP_BITVECTOR_EXACT ? (uint64_t)(f->max_value - f->min_value) + 1 : f->cardinality;
  return bm_words(bits);
}
/* BitVectorExactFilter::insert / contains (…ExactFilter.hpp:152-176),
 * SingleIdentityHashFilter::insert / contains (…HashFilter.hpp:156-171). */
static inline void lip_insert(qso_lip *f, int64_t v) {
  if (f->kind == QS_LIP_BITVECTOR_EXACT) {
    if (v < f->min_value || v > f->max_value) return;
    const uint64_t bit = (uint64_t)(v - f->min_value);
    __atomic_fetch_or(&f->words[bit >> 6], 0x8000000000000000ull >> (bit & 63), __ATOMIC_RELAXED);
  } else {
    const uint64_t bit = (uint64_t)v % f->cardinality;
    __atomic_fetch_or(&f->words[bit >> 6], 0x8000000000000000ull >> (bit & 63), __ATOMIC_RELAXED);
  }
}
static inline int lip_contains(const qso_lip *f, int64_t v) {
  if (f->kind == QS_LIP_BITVECTOR_EXACT) {
    if (v < f->min_value || v > f->max_value) return f->is_anti != 0;
    const int set = bm_get(f->words, (uint64_t)(v - f->min_value));
    return f->is_anti ? !set : set;
  }
  return bm_get(f->words, (uint64_t)v % f->cardinality);
}

/* ------------------------------------------------------- vectors and typing */
typedef struct vec {
  int type;          /* QS_INT .. QS_DATE */
  uint32_t width;
  void *data;        /* n values (or 1 when is_const) */
  int owned;
  int is_const;
} vec;

typedef struct blockctx {
  const qs_expr_set *ex;
  const qso_table *t;      /* scanned / probe-side table */
  const qso_table *build;  /* join build side (attribute nodes with b == 2), may be NULL */
  uint64_t row0, n;        /* block = rows [row0, row0 + n) */
  int error;
} blockctx;

static void vec_free(vec *v) { if (v->owned) free(v->data); v->data = NULL; v->owned = 0; }

static uint32_t type_width(int type, uint32_t width) {
  switch (type) {
    case QS_INT: case QS_FLOAT: return 4;
    case QS_LONG: case QS_DOUBLE: case QS_DATE: return 8;
    default: return width;
  }
}

/* TypeFactory::GetUnifyingType, numeric pairs (types/TypeFactory.cpp:159-180). */
static int unify_type(int a, int b) {
  if (a == b) return a;
  if (a == QS_DOUBLE || b == QS_DOUBLE) return QS_DOUBLE;
  if ((a == QS_LONG && b == QS_FLOAT) || (a == QS_FLOAT && b == QS_LONG)) return QS_DOUBLE;
  if (a == QS_FLOAT || b == QS_FLOAT) return QS_FLOAT;
  return QS_LONG;
}
static int is_numeric(int t) { return t == QS_INT || t == QS_LONG || t == QS_FLOAT || t == QS_DOUBLE; }

/* Materialise `v` as a non-const vector of numeric type `to` (static_cast). */
static vec vec_cast(const vec *v, int to, uint64_t n) {
  vec r = {to, type_width(to, 0), NULL, 1, 0};
  r.data = malloc((n ? n : 1) * r.width);
#define CAST_LOOP(ST, DT)                                                      \
  do {                                                                         \
    const ST *s = (const ST *)v->data;                                         \
    DT *d = (DT *)r.data;                                                      \
    if (v->is_const) { const DT c = (DT)s[0]; for (uint64_t i = 0; i < n; ++i) d[i] = c; } \
    else for (uint64_t i = 0; i < n; ++i) d[i] = (DT)s[i];                     \
  } while (0)
#define CAST_FROM(ST)                                                          \
  switch (to) {                                                                \
    case QS_INT: CAST_LOOP(ST, int32_t); break;                                \
    case QS_LONG: CAST_LOOP(ST, int64_t); break;                               \
    case QS_FLOAT: CAST_LOOP(ST, float); break;                                \
    default: CAST_LOOP(ST, double); break;                                     \
  }
  switch (v->type) {
    case QS_INT: CAST_FROM(int32_t); break;
    case QS_LONG: CAST_FROM(int64_t); break;
    case QS_FLOAT: CAST_FROM(float); break;
    default: CAST_FROM(double); break;
  }
  return r;
}

static vec eval_scalar(blockctx *c, int32_t idx);

static const qs_node *get_node(blockctx *c, int32_t idx) {
  if (idx < 0 || (uint32_t)idx >= c->ex->n_nodes) { c->error = 1; return NULL; }
  return &c->ex->nodes[idx];
}

/* ArithmeticBinaryOperators.hpp:51-159: both operands promoted to the unified
 * type, then the plain C operator; one output vector per node. */
static vec eval_binary(blockctx *c, const qs_node *n) {
  vec a = eval_scalar(c, n->a), b = eval_scalar(c, n->b);
  vec r = {QS_LONG, 8, NULL, 0, 0};
  if (c->error || !is_numeric(a.type) || !is_numeric(b.type)) { c->error = 1; vec_free(&a); vec_free(&b); return r; }
  const int T = unify_type(a.type, b.type);
  vec ca = vec_cast(&a, T, c->n), cb = vec_cast(&b, T, c->n);
  vec_free(&a); vec_free(&b);
  r.type = T; r.width = type_width(T, 0); r.owned = 1;
  r.data = malloc((c->n ? c->n : 1) * r.width);
#define BIN_LOOP(TT, EXPR)                                                     \
  do {                                                                         \
    const TT *x = (const TT *)ca.data, *y = (const TT *)cb.data;               \
    TT *o = (TT *)r.data;                                                      \
    for (uint64_t i = 0; i < c->n; ++i) o[i] = (EXPR);                          \
  } while (0)
#define BIN_TYPE(TT, MODEXPR)                                                  \
  switch (n->op) {                                                             \
    case QS_ADD: BIN_LOOP(TT, x[i] + y[i]); break;                             \
    case QS_SUB: BIN_LOOP(TT, x[i] - y[i]); break;                             \
    case QS_MUL: BIN_LOOP(TT, x[i] * y[i]); break;                             \
    case QS_DIV: BIN_LOOP(TT, DIVEXPR); break;                                 \
    default: BIN_LOOP(TT, MODEXPR); break;                                     \
  }
  switch (T) {
#define DIVEXPR (y[i] == 0 ? 0 : x[i] / y[i])
    case QS_INT: BIN_TYPE(int32_t, (y[i] == 0 ? 0 : x[i] % y[i])); break;
    case QS_LONG: BIN_TYPE(int64_t, (y[i] == 0 ? 0 : x[i] % y[i])); break;
#undef DIVEXPR
#define DIVEXPR (x[i] / y[i])
    case QS_FLOAT: BIN_TYPE(float, fmodf(x[i], y[i])); break;
    default: BIN_TYPE(double, fmod(x[i], y[i])); break;
#undef DIVEXPR
  }
  vec_free(&ca); vec_free(&cb);
  return r;
}

static vec eval_scalar(blockctx *c, int32_t idx) {
  vec r = {QS_LONG, 8, NULL, 0, 0};
  const qs_node *n = get_node(c, idx);
  if (!n) return r;
  switch (n->kind) {
    case QS_N_LITERAL:
      r.type = n->type; r.width = type_width(n->type, n->width); r.is_const = 1;
      if (n->type == QS_CHAR) r.data = (void *)(c->ex->str_pool + n->lit.pool_offset);
      else r.data = (void *)&n->lit;
      return r;
    case QS_N_ATTRIBUTE: {
      const qso_table *t = n->b == 2 ? c->build : c->t;
      if (!t || (uint32_t)n->a >= t->n_cols) { c->error = 1; return r; }
      const qso_column *col = &t->cols[n->a];
      r.type = col->type; r.width = col->width;
      r.data = (void *)((const char *)col->data + c->row0 * col->width);
      return r;
    }
    case QS_N_SHARED: return eval_scalar(c, n->a);   /* ColumnVectorCache only avoids recomputation */
    case QS_N_UNARY: {
      vec a = eval_scalar(c, n->a);
      if (c->error || !is_numeric(a.type)) { c->error = 1; vec_free(&a); return r; }
      if (n->op == QS_CAST) {
        if (!is_numeric(n->type)) { c->error = 1; vec_free(&a); return r; }
        r = vec_cast(&a, n->type, c->n);
        vec_free(&a);
        return r;
      }
      if (n->op != QS_NEGATE) { c->error = 1; vec_free(&a); return r; }
      r = vec_cast(&a, a.type, c->n);
      vec_free(&a);
      switch (r.type) {
        case QS_INT: { int32_t *d = r.data; for (uint64_t i = 0; i < c->n; ++i) d[i] = -d[i]; break; }
        case QS_LONG: { int64_t *d = r.data; for (uint64_t i = 0; i < c->n; ++i) d[i] = -d[i]; break; }
        case QS_FLOAT: { float *d = r.data; for (uint64_t i = 0; i < c->n; ++i) d[i] = -d[i]; break; }
        default: { double *d = r.data; for (uint64_t i = 0; i < c->n; ++i) d[i] = -d[i]; break; }
      }
      return r;
    }
    case QS_N_BINARY: return eval_binary(c, n);
    default: c->error = 1; return r;
  }
}

/* DateLit ordering, types/DatetimeLit.hpp:65-93. */
typedef struct { int32_t year; uint8_t month, day, pad[2]; } date_lit;
static inline int date_cmp(const date_lit *a, const date_lit *b) {
  if (a->year != b->year) return a->year < b->year ? -1 : 1;
  if (a->month != b->month) return a->month < b->month ? -1 : 1;
  if (a->day != b->day) return a->day < b->day ? -1 : 1;
  return 0;
}
static inline int apply_cmp(int op, int c3) {
  switch (op) {
    case QS_EQ: return c3 == 0;
    case QS_NE: return c3 != 0;
    case QS_LT: return c3 < 0;
    case QS_LE: return c3 <= 0;
    case QS_GT: return c3 > 0;
    default: return c3 >= 0;
  }
}
/* AsciiStringUncheckedComparator::strcmpHelper (AsciiStringComparators.hpp:218-251). */
static int char_cmp(const char *l, uint32_t ll, const char *r, uint32_t rl) {
  if (rl > ll) {
    int res = strncmp(l, r, ll);
    if (res) return res;
    return strnlen(r, rl) > ll ? -1 : 0;
  } else if (ll > rl) {
    int res = strncmp(l, r, rl);
    if (res) return res;
    return strnlen(l, ll) > rl ? 1 : 0;
  }
  return strncmp(l, r, ll);
}

/* Evaluates predicate `idx` over the block into bm (block-relative bits). */
static void eval_pred(blockctx *c, int32_t idx, uint64_t *bm) {
  const qs_node *n = get_node(c, idx);
  const uint64_t nw = bm_words(c->n);
  if (!n) return;
  switch (n->kind) {
    case QS_N_TRUE:
      for (uint64_t i = 0; i < c->n; ++i) bm_set(bm, i);
      return;
    case QS_N_FALSE: return;
    case QS_N_NEGATION: {
      uint64_t *t = calloc(nw ? nw : 1, 8);
      eval_pred(c, n->a, t);
      for (uint64_t i = 0; i < c->n; ++i) if (!bm_get(t, i)) bm_set(bm, i);
      free(t);
      return;
    }
    case QS_N_CONJUNCTION:
    case QS_N_DISJUNCTION: {
      uint64_t *x = calloc(nw ? nw : 1, 8), *y = calloc(nw ? nw : 1, 8);
      eval_pred(c, n->a, x);
      eval_pred(c, n->b, y);
      for (uint64_t w = 0; w < nw; ++w) bm[w] |= n->kind == QS_N_CONJUNCTION ? (x[w] & y[w]) : (x[w] | y[w]);
      free(x); free(y);
      return;
    }
    case QS_N_COMPARISON: {
      vec a = eval_scalar(c, n->a), b = eval_scalar(c, n->b);
      if (c->error) { vec_free(&a); vec_free(&b); return; }
      if (a.type == QS_CHAR && b.type == QS_CHAR) {
        for (uint64_t i = 0; i < c->n; ++i) {
          const char *l = (const char *)a.data + (a.is_const ? 0 : i * a.width);
          const char *r = (const char *)b.data + (b.is_const ? 0 : i * b.width);
          if (apply_cmp(n->op, char_cmp(l, a.width, r, b.width))) bm_set(bm, i);
        }
      } else if (a.type == QS_DATE && b.type == QS_DATE) {
        for (uint64_t i = 0; i < c->n; ++i) {
          const date_lit *l = (const date_lit *)a.data + (a.is_const ? 0 : i);
          const date_lit *r = (const date_lit *)b.data + (b.is_const ? 0 : i);
          if (apply_cmp(n->op, date_cmp(l, r))) bm_set(bm, i);
        }
      } else if (is_numeric(a.type) && is_numeric(b.type)) {
        /* LiteralComparators.hpp:36-72: C++ comparison after the usual promotion. */
        const int T = unify_type(a.type, b.type);
        vec ca = vec_cast(&a, T, c->n), cb = vec_cast(&b, T, c->n);
#define CMP_LOOP(TT)                                                           \
  do {                                                                         \
    const TT *x = (const TT *)ca.data, *y = (const TT *)cb.data;               \
    for (uint64_t i = 0; i < c->n; ++i) {                                      \
      int ok;                                                                  \
      switch (n->op) {                                                         \
        case QS_EQ: ok = x[i] == y[i]; break;                                  \
        case QS_NE: ok = x[i] != y[i]; break;                                  \
        case QS_LT: ok = x[i] < y[i]; break;                                   \
        case QS_LE: ok = x[i] <= y[i]; break;                                  \
        case QS_GT: ok = x[i] > y[i]; break;                                   \
        default: ok = x[i] >= y[i]; break;                                     \
      }                                                                        \
      if (ok) bm_set(bm, i);                                                   \
    }                                                                          \
  } while (0)
        switch (T) {
          case QS_INT: CMP_LOOP(int32_t); break;
          case QS_LONG: CMP_LOOP(int64_t); break;
          case QS_FLOAT: CMP_LOOP(float); break;
          default: CMP_LOOP(double); break;
        }
        vec_free(&ca); vec_free(&cb);
      } else {
        c->error = 1;
      }
      vec_free(&a); vec_free(&b);
      return;
    }
    default: c->error = 1;
  }
}

static int64_t col_int(const qso_column *col, uint64_t row) {
  return col->type == QS_INT ? (int64_t)((const int32_t *)col->data)[row] : ((const int64_t *)col->data)[row];
}

/* Predicate, then LIPFilterAdaptiveProber::filterValueAccessor: the adaptive
 * batching only reorders filters, the surviving set is the intersection
 * (utility/lip_filter/LIPFilterAdaptiveProber.hpp:89-232). */
static void block_matches(blockctx *c, int32_t pred_root, uint32_t n_probe, const qso_lip_ref *probe, uint64_t *bm) {
  if (pred_root >= 0) eval_pred(c, pred_root, bm);
  else for (uint64_t i = 0; i < c->n; ++i) bm_set(bm, i);
  for (uint32_t f = 0; f < n_probe; ++f) {
    const qso_column *col = &c->t->cols[probe[f].attr];
    for (uint64_t i = 0; i < c->n; ++i)
      if (bm_get(bm, i) && !lip_contains(probe[f].lip, col_int(col, c->row0 + i)))
        bm[i >> 6] &= ~(0x8000000000000000ull >> (i & 63));
  }
}

/* ------------------------------------------------------- block work orders */
typedef void (*block_fn)(void *arg, uint64_t block, uint64_t row0, uint64_t n);
typedef struct pool_job { block_fn fn; void *arg; uint64_t n_rows, n_blocks, next; } pool_job;

static void *pool_worker(void *p) {
  pool_job *j = p;
  for (;;) {
    const uint64_t b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (b >= j->n_blocks) break;
    const uint64_t row0 = b * g_block_rows;
    const uint64_t n = j->n_rows - row0 < g_block_rows ? j->n_rows - row0 : g_block_rows;
    j->fn(j->arg, b, row0, n);
  }
  return NULL;
}
static uint64_t num_blocks(uint64_t n_rows) { return (n_rows + g_block_rows - 1) / g_block_rows; }
static void run_blocks(uint64_t n_rows, block_fn fn, void *arg) {
  pool_job j = {fn, arg, n_rows, num_blocks(n_rows), 0};
  int nt = g_workers;
  if ((uint64_t)nt > j.n_blocks) nt = (int)j.n_blocks;
  if (nt <= 1) { pool_worker(&j); return; }
  pthread_t *th = malloc(sizeof(pthread_t) * nt);
  for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, pool_worker, &j);
  for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
  free(th);
}

/* ----------------------------------------------------- predicate / scalar */
typedef struct { const qs_expr_set *ex; int32_t root; const qso_table *t; uint64_t *bm; void *out; int out_type; int error; int64_t *counts; } simple_job;

static void pred_block(void *arg, uint64_t b, uint64_t row0, uint64_t n) {
  simple_job *j = arg;
  blockctx c = {j->ex, j->t, NULL, row0, n, 0};
  uint64_t *bm = j->bm + row0 / 64;
  eval_pred(&c, j->root, bm);
  int64_t cnt = 0;
  for (uint64_t i = 0; i < n; ++i) cnt += bm_get(bm, i);
  j->counts[b] = cnt;
  if (c.error) j->error = 1;
}
int64_t qso_predicate(const qs_expr_set *ex, int32_t root, const qso_table *t, uint64_t *bitmap_words) {
  memset(bitmap_words, 0, bm_words(t->n_rows) * 8);
  simple_job j = {ex, root, t, bitmap_words, NULL, 0, 0, calloc(num_blocks(t->n_rows) + 1, 8)};
  run_blocks(t->n_rows, pred_block, &j);
  int64_t total = 0;
  for (uint64_t b = 0; b < num_blocks(t->n_rows); ++b) total += j.counts[b];
  free(j.counts);
  return j.error ? -1 : total;
}

static void scalar_block(void *arg, uint64_t b, uint64_t row0, uint64_t n) {
  (void)b;
  simple_job *j = arg;
  blockctx c = {j->ex, j->t, NULL, row0, n, 0};
  vec v = eval_scalar(&c, j->root);
  if (c.error) { j->error = 1; vec_free(&v); return; }
  j->out_type = v.type;
  char *dst = (char *)j->out + row0 * v.width;
  if (v.is_const) for (uint64_t i = 0; i < n; ++i) memcpy(dst + i * v.width, v.data, v.width);
  else memcpy(dst, v.data, n * v.width);
  vec_free(&v);
}
int qso_scalar(const qs_expr_set *ex, int32_t root, const qso_table *t, void *out) {
  simple_job j = {ex, root, t, NULL, out, -1, 0, NULL};
  run_blocks(t->n_rows, scalar_block, &j);
  return j.error ? -1 : j.out_type;
}

/* ---------------------------------------------------------- BuildLIPFilter */
typedef struct { const qs_expr_set *ex; int32_t pred; const qso_table *t; uint32_t n_probe; const qso_lip_ref *probe; uint32_t n_build; const qso_lip_ref *build; int error; } lip_job;
static void lip_block(void *arg, uint64_t b, uint64_t row0, uint64_t n) {
  (void)b;
  lip_job *j = arg;
  blockctx c = {j->ex, j->t, NULL, row0, n, 0};
  uint64_t *bm = calloc(bm_words(n) + 1, 8);
  block_matches(&c, j->pred, j->n_probe, j->probe, bm);
  for (uint32_t f = 0; f < j->n_build; ++f) {
    const qso_column *col = &j->t->cols[j->build[f].attr];
    for (uint64_t i = 0; i < n; ++i) if (bm_get(bm, i)) lip_insert(j->build[f].lip, col_int(col, row0 + i));
  }
  free(bm);
  if (c.error) j->error = 1;
}
int qso_build_lip_filter(const qs_expr_set *ex, int32_t predicate_root, const qso_table *t, uint32_t n_probe,
                         const qso_lip_ref *probe, uint32_t n_build, const qso_lip_ref *build) {
  lip_job j = {ex, predicate_root, t, n_probe, probe, n_build, build, 0};
  run_blocks(t->n_rows, lip_block, &j);
  return j.error ? -1 : 0;
}

/* ------------------------------------------------------------------ Select */
typedef struct {
  const qs_expr_set *ex; int32_t pred; const qso_table *t; uint32_t n_probe; const qso_lip_ref *probe;
  uint32_t n_project; const int32_t *roots; void *const *out; uint64_t *bm; int64_t *counts; int error; int phase;
  uint64_t *offsets;
} select_job;

static void select_block(void *arg, uint64_t b, uint64_t row0, uint64_t n) {
  select_job *j = arg;
  blockctx c = {j->ex, j->t, NULL, row0, n, 0};
  uint64_t *bm = j->bm + row0 / 64;
  if (j->phase == 0) {
    block_matches(&c, j->pred, j->n_probe, j->probe, bm);
    int64_t cnt = 0;
    for (uint64_t i = 0; i < n; ++i) cnt += bm_get(bm, i);
    j->counts[b] = cnt;
  } else {
    /* StorageBlock::select: one ColumnVector per projected scalar, then the
     * matching tuples are inserted in tuple order. */
    for (uint32_t p = 0; p < j->n_project; ++p) {
      vec v = eval_scalar(&c, j->roots[p]);
      if (c.error) { vec_free(&v); break; }
      char *dst = (char *)j->out[p] + j->offsets[b] * v.width;
      for (uint64_t i = 0; i < n; ++i) {
        if (!bm_get(bm, i)) continue;
        memcpy(dst, (const char *)v.data + (v.is_const ? 0 : i * v.width), v.width);
        dst += v.width;
      }
      vec_free(&v);
    }
  }
  if (c.error) j->error = 1;
}

int64_t qso_select(const qs_expr_set *ex, int32_t predicate_root, const qso_table *t, uint32_t n_probe,
                   const qso_lip_ref *probe, uint32_t n_project, const int32_t *project_roots,
                   void *const *out_cols) {
  const uint64_t nb = num_blocks(t->n_rows);
  select_job j = {ex, predicate_root, t, n_probe, probe, n_project, project_roots, out_cols,
                  calloc(bm_words(t->n_rows) + 1, 8), calloc(nb + 1, 8), 0, 0, calloc(nb + 1, 8)};
  run_blocks(t->n_rows, select_block, &j);
  uint64_t total = 0;
  for (uint64_t b = 0; b < nb; ++b) { j.offsets[b] = total; total += (uint64_t)j.counts[b]; }
  j.phase = 1;
  if (!j.error) run_blocks(t->n_rows, select_block, &j);
  free(j.bm); free(j.counts); free(j.offsets);
  return j.error ? -1 : (int64_t)total;
}

/* ------------------------------------------------------------- aggregation */
enum { K_SUM_F64, K_SUM_I64, K_MIN_F64, K_MAX_F64, K_MIN_I64, K_MAX_I64, K_COUNT };

typedef struct gtable {
  uint32_t key_bytes, n_agg;
  uint64_t cap, n;           /* slots (pow2), groups */
  int64_t *slot;             /* slot -> group index or -1 */
  uint8_t *keys;             /* [n][key_bytes] */
  uint64_t *vals;            /* [n][n_agg] raw words */
  int64_t *counts;           /* [n] */
  uint64_t gcap;
} gtable;

static uint64_t hash_bytes(const uint8_t *p, uint32_t n) {
  uint64_t h = 1469598103934665603ull;
  for (uint32_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  h ^= h >> 29;
  return h;
}
static void gt_init(gtable *g, uint32_t key_bytes, uint32_t n_agg) {
  memset(g, 0, sizeof(*g));
  g->key_bytes = key_bytes; g->n_agg = n_agg; g->cap = 64; g->gcap = 16;
  g->slot = malloc(g->cap * 8);
  for (uint64_t i = 0; i < g->cap; ++i) g->slot[i] = -1;
  g->keys = malloc(g->gcap * (key_bytes ? key_bytes : 1));
  g->vals = malloc(g->gcap * (n_agg ? n_agg : 1) * 8);
  g->counts = malloc(g->gcap * 8);
}
static void gt_free(gtable *g) { free(g->slot); free(g->keys); free(g->vals); free(g->counts); }
static uint64_t ident(int kind) {
  union { double d; uint64_t u; int64_t i; } x;
  switch (kind) {
    case K_MIN_F64: x.d = INFINITY; return x.u;
    case K_MAX_F64: x.d = -INFINITY; return x.u;
    case K_MIN_I64: x.i = INT64_MAX; return x.u;
    case K_MAX_I64: x.i = INT64_MIN; return x.u;
    default: return 0;
  }
}
static uint64_t gt_find_or_insert(gtable *g, const uint8_t *key, const int *kinds) {
  if ((g->n + 1) * 2 > g->cap) {
    g->cap *= 2;
    g->slot = realloc(g->slot, g->cap * 8);
    for (uint64_t i = 0; i < g->cap; ++i) g->slot[i] = -1;
    for (uint64_t q = 0; q < g->n; ++q) {
      uint64_t h = hash_bytes(g->keys + q * g->key_bytes, g->key_bytes) & (g->cap - 1);
      while (g->slot[h] >= 0) h = (h + 1) & (g->cap - 1);
      g->slot[h] = (int64_t)q;
    }
  }
  uint64_t h = hash_bytes(key, g->key_bytes) & (g->cap - 1);
  while (g->slot[h] >= 0) {
    if (memcmp(g->keys + (uint64_t)g->slot[h] * g->key_bytes, key, g->key_bytes) == 0) return (uint64_t)g->slot[h];
    h = (h + 1) & (g->cap - 1);
  }
  if (g->n == g->gcap) {
    g->gcap *= 2;
    g->keys = realloc(g->keys, g->gcap * (g->key_bytes ? g->key_bytes : 1));
    g->vals = realloc(g->vals, g->gcap * (g->n_agg ? g->n_agg : 1) * 8);
    g->counts = realloc(g->counts, g->gcap * 8);
  }
  const uint64_t q = g->n++;
  g->slot[h] = (int64_t)q;
  memcpy(g->keys + q * g->key_bytes, key, g->key_bytes);
  for (uint32_t a = 0; a < g->n_agg; ++a) g->vals[q * g->n_agg + a] = ident(kinds[a]);
  g->counts[q] = 0;
  return q;
}
static inline uint64_t combine(int kind, uint64_t a, uint64_t b) {
  union { double d; uint64_t u; int64_t i; } x, y;
  x.u = a; y.u = b;
  switch (kind) {
    case K_SUM_F64: x.d = x.d + y.d; return x.u;
    case K_SUM_I64: x.i = x.i + y.i; return x.u;
    case K_MIN_F64: return y.d < x.d ? b : a;
    case K_MAX_F64: return y.d > x.d ? b : a;
    case K_MIN_I64: return y.i < x.i ? b : a;
    case K_MAX_I64: return y.i > x.i ? b : a;
    default: return a;
  }
}

typedef struct {
  const qs_expr_set *ex; int32_t pred; uint32_t n_aggs; const qs_aggregate *aggs; uint32_t n_group;
  const int32_t *group_roots; const qso_table *t; uint32_t n_probe; const qso_lip_ref *probe;
  int kinds[16]; int arg_type[16]; uint32_t key_bytes; uint32_t key_width[16]; uint32_t key_attr[16];
  gtable *block_tables; int error; int shared_table; gtable shared;
} agg_job;

/* AggregationOperationState::aggregateBlock (…State.cpp:428-474): predicate ->
 * LIP -> argument vectors -> per-tuple upsert in tuple order. */
static void agg_block_into(agg_job *j, gtable *g, uint64_t row0, uint64_t n) {
  blockctx c = {j->ex, j->t, NULL, row0, n, 0};
  uint64_t *bm = calloc(bm_words(n) + 1, 8);
  block_matches(&c, j->pred, j->n_probe, j->probe, bm);
  vec args[16];
  for (uint32_t a = 0; a < j->n_aggs; ++a) {
    args[a].data = NULL; args[a].owned = 0;
    if (j->kinds[a] == K_COUNT) continue;
    vec v = eval_scalar(&c, j->aggs[a].argument_root);
    if (c.error) { vec_free(&v); continue; }
    /* SUM/AVG accumulate INT as LONG and FLOAT as DOUBLE (AggregationHandleSum.cpp:49-64). */
    const int fp = j->kinds[a] == K_SUM_F64 || j->kinds[a] == K_MIN_F64 || j->kinds[a] == K_MAX_F64;
    args[a] = vec_cast(&v, fp ? QS_DOUBLE : QS_LONG, n);
    vec_free(&v);
  }
  uint8_t key[64];
  if (!c.error) {
    for (uint64_t i = 0; i < n; ++i) {
      if (!bm_get(bm, i)) continue;
      /* ThreadPrivateCompactKeyHashTable::ConstructKeyCode: memcpy key k at the
       * running byte offset of a zeroed code (…CompactKeyHashTable.hpp:125-142). */
      uint32_t off = 0;
      for (uint32_t k = 0; k < j->n_group; ++k) {
        const qso_column *col = &j->t->cols[j->key_attr[k]];
        memcpy(key + off, (const char *)col->data + (row0 + i) * col->width, col->width);
        off += col->width;
      }
      const uint64_t q = gt_find_or_insert(g, key, j->kinds);
      g->counts[q] += 1;
      for (uint32_t a = 0; a < j->n_aggs; ++a) {
        if (j->kinds[a] == K_COUNT) continue;
        g->vals[q * g->n_agg + a] = combine(j->kinds[a], g->vals[q * g->n_agg + a], ((const uint64_t *)args[a].data)[i]);
      }
    }
  }
  for (uint32_t a = 0; a < j->n_aggs; ++a) vec_free(&args[a]);
  free(bm);
  if (c.error) j->error = 1;
}
static void agg_block(void *arg, uint64_t b, uint64_t row0, uint64_t n) {
  agg_job *j = arg;
  gt_init(&j->block_tables[b], j->key_bytes, j->n_aggs);
  agg_block_into(j, &j->block_tables[b], row0, n);
}

static int key_compare_bytes;
static const uint8_t *key_compare_base;
static int key_compare(const void *a, const void *b) {
  return memcmp(key_compare_base + *(const uint64_t *)a * key_compare_bytes,
                key_compare_base + *(const uint64_t *)b * key_compare_bytes, key_compare_bytes);
}

int qso_aggregate(const qs_expr_set *ex, int32_t predicate_root, uint32_t n_aggregates,
                  const qs_aggregate *aggregates, uint32_t n_group_by, const int32_t *group_by_roots,
                  const qso_table *t, uint32_t n_probe, const qso_lip_ref *probe, qso_agg_result *out) {
  memset(out, 0, sizeof(*out));
  if (n_aggregates > 16 || n_group_by > 16) return -1;
  agg_job j;
  memset(&j, 0, sizeof(j));
  j.ex = ex; j.pred = predicate_root; j.n_aggs = n_aggregates; j.aggs = aggregates; j.n_group = n_group_by;
  j.group_roots = group_by_roots; j.t = t; j.n_probe = n_probe; j.probe = probe;
  for (uint32_t k = 0; k < n_group_by; ++k) {
    const qs_node *n = &ex->nodes[group_by_roots[k]];
    if (n->kind != QS_N_ATTRIBUTE || (uint32_t)n->a >= t->n_cols) return -1;
    j.key_attr[k] = (uint32_t)n->a;
    j.key_width[k] = t->cols[n->a].width;
    j.key_bytes += j.key_width[k];
  }
  if (j.key_bytes > 64) return -1;
  /* argument types decide the state type */
  for (uint32_t a = 0; a < n_aggregates; ++a) {
    if (aggregates[a].function == QS_AGG_COUNT) { j.kinds[a] = K_COUNT; continue; }
    blockctx c = {ex, t, NULL, 0, 0, 0};
    vec v = eval_scalar(&c, aggregates[a].argument_root);
    const int fp = v.type == QS_FLOAT || v.type == QS_DOUBLE;
    j.arg_type[a] = v.type;
    vec_free(&v);
    if (c.error) return -1;
    switch (aggregates[a].function) {
      case QS_AGG_SUM: case QS_AGG_AVG: j.kinds[a] = fp ? K_SUM_F64 : K_SUM_I64; break;
      case QS_AGG_MIN: j.kinds[a] = fp ? K_MIN_F64 : K_MIN_I64; break;
      case QS_AGG_MAX: j.kinds[a] = fp ? K_MAX_F64 : K_MAX_I64; break;
      default: return -1;
    }
  }
  gtable final;
  gt_init(&final, j.key_bytes, n_aggregates);
  if (j.key_bytes <= 8) {
    /* no GROUP BY (aggregateBlockSingleState + mergeSingleState) and the
     * thread-private compact-key table: one private state per block, merged in
     * block order (ThreadPrivateCompactKeyHashTable::mergeFrom, …Table.cpp:306-363). */
    const uint64_t nb = num_blocks(t->n_rows);
    j.block_tables = calloc(nb + 1, sizeof(gtable));
    run_blocks(t->n_rows, agg_block, &j);
    if (n_group_by == 0) { uint8_t k0 = 0; gt_find_or_insert(&final, &k0, j.kinds); }
    for (uint64_t b = 0; b < nb; ++b) {
      gtable *g = &j.block_tables[b];
      for (uint64_t q = 0; q < g->n; ++q) {
        const uint64_t d = gt_find_or_insert(&final, g->keys + q * g->key_bytes, j.kinds);
        final.counts[d] += g->counts[q];
        for (uint32_t a = 0; a < n_aggregates; ++a)
          final.vals[d * n_aggregates + a] = combine(j.kinds[a], final.vals[d * n_aggregates + a], g->vals[q * n_aggregates + a]);
      }
      gt_free(g);
    }
    free(j.block_tables);
  } else {
    /* shared PackedPayloadHashTable / CollisionFreeVectorTable: every tuple
     * updates the one table; sequential block order here. */
    for (uint64_t row0 = 0; row0 < t->n_rows; row0 += g_block_rows) {
      const uint64_t n = t->n_rows - row0 < g_block_rows ? t->n_rows - row0 : g_block_rows;
      agg_block_into(&j, &final, row0, n);
    }
  }
  if (j.error) { gt_free(&final); return -1; }
  /* finalize: groups in key order */
  const uint64_t G = final.n;
  uint64_t *order = malloc((G ? G : 1) * 8);
  for (uint64_t q = 0; q < G; ++q) order[q] = q;
  key_compare_bytes = (int)j.key_bytes;
  key_compare_base = final.keys;
  if (j.key_bytes) qsort(order, G, 8, key_compare);
  out->n_groups = G;
  out->key_bytes = j.key_bytes;
  out->keys = malloc((G ? G : 1) * (j.key_bytes ? j.key_bytes : 1));
  out->values = malloc((G ? G : 1) * (n_aggregates ? n_aggregates : 1) * 8);
  out->counts = malloc((G ? G : 1) * 8);
  for (uint64_t o = 0; o < G; ++o) {
    const uint64_t q = order[o];
    memcpy(out->keys + o * j.key_bytes, final.keys + q * j.key_bytes, j.key_bytes);
    out->counts[o] = final.counts[q];
    for (uint32_t a = 0; a < n_aggregates; ++a) {
      union { double d; uint64_t u; int64_t i; } x;
      x.u = final.vals[q * n_aggregates + a];
      const int kind = j.kinds[a];
      const int fp = kind == K_SUM_F64 || kind == K_MIN_F64 || kind == K_MAX_F64;
      if (aggregates[a].function == QS_AGG_COUNT) { x.i = final.counts[q]; out->is_double[a] = 0; }
      else if (aggregates[a].function == QS_AGG_AVG) {
        /* AggregationHandleAvg::finalize: sum / static_cast<double>(count) (…Avg.cpp:144-155) */
        const double s = fp ? x.d : (double)x.i;
        x.d = final.counts[q] ? s / (double)final.counts[q] : 0.0;
        out->is_double[a] = 1;
      } else {
        if (final.counts[q] == 0) x.u = 0;
        out->is_double[a] = (uint8_t)fp;
      }
      out->values[(uint64_t)a * G + o] = x.u;
      if (n_group_by == 0 && final.counts[q] == 0 && aggregates[a].function != QS_AGG_COUNT) out->is_null[a] = 1;
    }
  }
  free(order);
  gt_free(&final);
  return 0;
}

void qso_agg_result_free(qso_agg_result *r) {
  free(r->keys); free(r->values); free(r->counts);
  memset(r, 0, sizeof(*r));
}

/* --------------------------------------------------------------- hash join */
/* Does the scalar rooted at `i` read a build-side attribute?  (is_selection_on_build of
 * HashOuterJoinWorkOrder, relational_operators/HashJoinOperator.hpp:724: such output columns are NULL for
 * probe tuples without a match.) */
static int refs_build_side(const qs_expr_set *ex, int32_t i) {
  if (i < 0 || (uint32_t)i >= ex->n_nodes) return 0;
  const qs_node *n = &ex->nodes[i];
  switch (n->kind) {
    case QS_N_ATTRIBUTE: return n->b == 2;
    case QS_N_UNARY: case QS_N_SHARED: return refs_build_side(ex, n->a);
    case QS_N_BINARY: return refs_build_side(ex, n->a) || refs_build_side(ex, n->b);
    default: return 0;
  }
}

static int64_t hash_join_impl(const qs_expr_set *ex, const qso_table *build, int32_t build_predicate_root,
                      uint32_t build_key_attr, const qso_table *probe, int32_t probe_predicate_root,
                      uint32_t probe_key_attr, uint32_t n_probe_lip, const qso_lip_ref *probe_lip,
                      uint32_t join_type, int32_t residual_root, uint32_t n_project,
                      const int32_t *project_roots, void *const *out_cols, uint64_t out_capacity,
                      uint64_t *out_nulls) {
  if (join_type == QS_JOIN_LEFT_OUTER && residual_root >= 0) return -1;   /* DCHECK in the reference (HashJoinOperator.hpp:139-141) */
  /* BuildHashWorkOrder: predicate, then putValueAccessor(key -> tuple reference). */
  uint64_t *bbm = calloc(bm_words(build->n_rows) + 1, 8);
  {
    blockctx c = {ex, build, NULL, 0, build->n_rows, 0};
    block_matches(&c, build_predicate_root, 0, NULL, bbm);
    if (c.error) { free(bbm); return -1; }
  }
  uint64_t cap = 64;
  while (cap < build->n_rows * 2 + 2) cap <<= 1;
  int64_t *head = malloc(cap * 8), *next = malloc((build->n_rows + 1) * 8), *tail = malloc(cap * 8);
  for (uint64_t i = 0; i < cap; ++i) head[i] = tail[i] = -1;
  const qso_column *bk = &build->cols[build_key_attr];
  for (uint64_t r = 0; r < build->n_rows; ++r) {
    next[r] = -1;
    if (!bm_get(bbm, r)) continue;
    const int64_t key = col_int(bk, r);
    const uint64_t h = ((uint64_t)key * 0x9e3779b97f4a7c15ull >> 20) & (cap - 1);
    if (tail[h] < 0) head[h] = (int64_t)r; else next[tail[h]] = (int64_t)r;   /* chains keep build order */
    tail[h] = (int64_t)r;
  }
  free(bbm);
  /* probe side */
  uint64_t *pbm = calloc(bm_words(probe->n_rows) + 1, 8);
  {
    blockctx c = {ex, probe, NULL, 0, probe->n_rows, 0};
    block_matches(&c, probe_predicate_root, n_probe_lip, probe_lip, pbm);
    if (c.error) { free(pbm); free(head); free(next); free(tail); return -1; }
  }
  const qso_column *pk = &probe->cols[probe_key_attr];
  uint64_t pair_cap = 1024, n_pairs = 0;
  uint64_t *pp = malloc(pair_cap * 8), *pb = malloc(pair_cap * 8);
  for (uint64_t r = 0; r < probe->n_rows; ++r) {
    if (!bm_get(pbm, r)) continue;
    const int64_t key = col_int(pk, r);
    const uint64_t h = ((uint64_t)key * 0x9e3779b97f4a7c15ull >> 20) & (cap - 1);
    for (int64_t q = head[h]; q >= 0; q = next[q]) {
      if (col_int(bk, (uint64_t)q) != key) continue;
      if (n_pairs == pair_cap) { pair_cap *= 2; pp = realloc(pp, pair_cap * 8); pb = realloc(pb, pair_cap * 8); }
      pp[n_pairs] = r; pb[n_pairs] = (uint64_t)q; ++n_pairs;
    }
  }
  free(head); free(next); free(tail);
  /* gather both sides of every candidate pair into a pair table so residual
   * predicate and projections are ordinary column-at-a-time evaluations
   * (Scalar::getAllValuesForJoin, HashJoinOperator.cpp:527-536). */
  qso_column *gp = malloc(sizeof(qso_column) * (probe->n_cols + 1)), *gb = malloc(sizeof(qso_column) * (build->n_cols + 1));
  for (uint32_t c = 0; c < probe->n_cols; ++c) {
    gp[c] = probe->cols[c];
    char *d = malloc((n_pairs ? n_pairs : 1) * gp[c].width);
    for (uint64_t i = 0; i < n_pairs; ++i) memcpy(d + i * gp[c].width, (const char *)probe->cols[c].data + pp[i] * gp[c].width, gp[c].width);
    gp[c].data = d;
  }
  for (uint32_t c = 0; c < build->n_cols; ++c) {
    gb[c] = build->cols[c];
    char *d = malloc((n_pairs ? n_pairs : 1) * gb[c].width);
    for (uint64_t i = 0; i < n_pairs; ++i) memcpy(d + i * gb[c].width, (const char *)build->cols[c].data + pb[i] * gb[c].width, gb[c].width);
    gb[c].data = d;
  }
  qso_table tp = {gp, probe->n_cols, n_pairs}, tb = {gb, build->n_cols, n_pairs};
  blockctx pc = {ex, &tp, &tb, 0, n_pairs, 0};
  uint64_t *ok = calloc(bm_words(n_pairs) + 1, 8);
  if (residual_root >= 0) eval_pred(&pc, residual_root, ok);
  else for (uint64_t i = 0; i < n_pairs; ++i) bm_set(ok, i);
  int64_t n_out = 0;
  int error = pc.error;
  if (!error && (join_type == QS_JOIN_INNER || join_type == QS_JOIN_LEFT_OUTER)) {
    for (uint32_t p = 0; p < n_project && !error; ++p) {
      vec v = eval_scalar(&pc, project_roots[p]);
      if (pc.error) { error = 1; vec_free(&v); break; }
      uint64_t o = 0;
      for (uint64_t i = 0; i < n_pairs; ++i) {
        if (!bm_get(ok, i)) continue;
        if (o < out_capacity) memcpy((char *)out_cols[p] + o * v.width, (const char *)v.data + (v.is_const ? 0 : i * v.width), v.width);
        ++o;
      }
      n_out = (int64_t)o;
      vec_free(&v);
    }
    if (n_project == 0) for (uint64_t i = 0; i < n_pairs; ++i) n_out += bm_get(ok, i);
    if (out_nulls) for (int64_t o = 0; o < n_out && (uint64_t)o < out_capacity; ++o) out_nulls[o] = 0;
    if (!error && join_type == QS_JOIN_LEFT_OUTER) {
      /* HashOuterJoinWorkOrder::execute second half (HashJoinOperator.cpp:1060-1099): probe tuples that passed
       * the filters and found no match are emitted with the probe-side selection evaluated and every
       * build-side selection NULL. */
      uint64_t *matched = calloc(bm_words(probe->n_rows) + 1, 8);
      for (uint64_t i = 0; i < n_pairs; ++i) bm_set(matched, pp[i]);
      uint64_t n_un = 0;
      for (uint64_t r = 0; r < probe->n_rows; ++r) n_un += bm_get(pbm, r) && !bm_get(matched, r);
      qso_column *up = malloc(sizeof(qso_column) * (probe->n_cols + 1)), *ub = malloc(sizeof(qso_column) * (build->n_cols + 1));
      for (uint32_t c = 0; c < probe->n_cols; ++c) {
        up[c] = probe->cols[c];
        char *d = malloc((n_un ? n_un : 1) * up[c].width);
        uint64_t o = 0;
        for (uint64_t r = 0; r < probe->n_rows; ++r)
          if (bm_get(pbm, r) && !bm_get(matched, r)) { memcpy(d + o * up[c].width, (const char *)probe->cols[c].data + r * up[c].width, up[c].width); ++o; }
        up[c].data = d;
      }
      for (uint32_t c = 0; c < build->n_cols; ++c) { ub[c] = build->cols[c]; ub[c].data = calloc(n_un ? n_un : 1, ub[c].width); }
      qso_table tup = {up, probe->n_cols, n_un}, tub = {ub, build->n_cols, n_un};
      blockctx uc = {ex, &tup, &tub, 0, n_un, 0};
      uint64_t null_bits = 0;
      for (uint32_t p = 0; p < n_project; ++p) if (refs_build_side(ex, project_roots[p])) null_bits |= 1ull << p;
      for (uint32_t p = 0; p < n_project && !error; ++p) {
        vec v = eval_scalar(&uc, project_roots[p]);
        if (uc.error) { error = 1; vec_free(&v); break; }
        for (uint64_t i = 0; i < n_un; ++i) {
          const uint64_t o = (uint64_t)n_out + i;
          if (o >= out_capacity) continue;
          if (null_bits >> p & 1) memset((char *)out_cols[p] + o * v.width, 0, v.width);
          else memcpy((char *)out_cols[p] + o * v.width, (const char *)v.data + (v.is_const ? 0 : i * v.width), v.width);
        }
        vec_free(&v);
      }
      if (out_nulls) for (uint64_t i = 0; i < n_un; ++i) if ((uint64_t)n_out + i < out_capacity) out_nulls[(uint64_t)n_out + i] = null_bits;
      n_out += (int64_t)n_un;
      for (uint32_t c = 0; c < probe->n_cols; ++c) free((void *)up[c].data);
      for (uint32_t c = 0; c < build->n_cols; ++c) free((void *)ub[c].data);
      free(up); free(ub); free(matched);
    }
  } else if (!error) {
    /* semi / anti: existence bitmap over probe rows (HashJoinOperator.cpp:673-987) */
    uint64_t *matched = calloc(bm_words(probe->n_rows) + 1, 8);
    for (uint64_t i = 0; i < n_pairs; ++i) if (bm_get(ok, i)) bm_set(matched, pp[i]);
    uint64_t *sel = calloc(bm_words(probe->n_rows) + 1, 8);
    for (uint64_t r = 0; r < probe->n_rows; ++r) {
      if (!bm_get(pbm, r)) continue;
      const int m = bm_get(matched, r);
      if ((join_type == QS_JOIN_LEFT_SEMI && m) || (join_type == QS_JOIN_LEFT_ANTI && !m)) bm_set(sel, r);
    }
    blockctx sc = {ex, probe, NULL, 0, probe->n_rows, 0};
    for (uint32_t p = 0; p < n_project && !error; ++p) {
      vec v = eval_scalar(&sc, project_roots[p]);
      if (sc.error) { error = 1; vec_free(&v); break; }
      uint64_t o = 0;
      for (uint64_t r = 0; r < probe->n_rows; ++r) {
        if (!bm_get(sel, r)) continue;
        if (o < out_capacity) memcpy((char *)out_cols[p] + o * v.width, (const char *)v.data + (v.is_const ? 0 : r * v.width), v.width);
        ++o;
      }
      n_out = (int64_t)o;
      vec_free(&v);
    }
    if (n_project == 0) for (uint64_t r = 0; r < probe->n_rows; ++r) n_out += bm_get(sel, r);
    free(matched); free(sel);
  }
  for (uint32_t c = 0; c < probe->n_cols; ++c) free((void *)gp[c].data);
  for (uint32_t c = 0; c < build->n_cols; ++c) free((void *)gb[c].data);
  free(gp); free(gb); free(ok); free(pp); free(pb); free(pbm);
  return error ? -1 : n_out;
}

int64_t qso_hash_join(const qs_expr_set *ex, const qso_table *build, int32_t build_predicate_root,
                      uint32_t build_key_attr, const qso_table *probe, int32_t probe_predicate_root,
                      uint32_t probe_key_attr, uint32_t n_probe_lip, const qso_lip_ref *probe_lip,
                      uint32_t join_type, int32_t residual_root, uint32_t n_project,
                      const int32_t *project_roots, void *const *out_cols, uint64_t out_capacity) {
  return hash_join_impl(ex, build, build_predicate_root, build_key_attr, probe, probe_predicate_root, probe_key_attr,
                        n_probe_lip, probe_lip, join_type, residual_root, n_project, project_roots, out_cols,
                        out_capacity, NULL);
}

int64_t qso_hash_join_nulls(const qs_expr_set *ex, const qso_table *build, int32_t build_predicate_root,
                            uint32_t build_key_attr, const qso_table *probe, int32_t probe_predicate_root,
                            uint32_t probe_key_attr, uint32_t n_probe_lip, const qso_lip_ref *probe_lip,
                            uint32_t join_type, int32_t residual_root, uint32_t n_project,
                            const int32_t *project_roots, void *const *out_cols, uint64_t out_capacity,
                            uint64_t *out_nulls) {
  return hash_join_impl(ex, build, build_predicate_root, build_key_attr, probe, probe_predicate_root, probe_key_attr,
                        n_probe_lip, probe_lip, join_type, residual_root, n_project, project_roots, out_cols,
                        out_capacity, out_nulls);
}

/* ------------------------------------------------------------------- top-k */
static const qso_table *g_sort_table;
static const qs_sort_key *g_sort_keys;
static uint32_t g_sort_nkeys;
static int sort_compare(const void *pa, const void *pb) {
  const uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
  for (uint32_t k = 0; k < g_sort_nkeys; ++k) {
    const qso_column *col = &g_sort_table->cols[g_sort_keys[k].attr];
    int c = 0;
    switch (col->type) {
      case QS_INT: { int32_t x = ((const int32_t *)col->data)[a], y = ((const int32_t *)col->data)[b]; c = (x > y) - (x < y); break; }
      case QS_LONG: { int64_t x = ((const int64_t *)col->data)[a], y = ((const int64_t *)col->data)[b]; c = (x > y) - (x < y); break; }
      case QS_FLOAT: { float x = ((const float *)col->data)[a], y = ((const float *)col->data)[b]; c = (x > y) - (x < y); break; }
      case QS_DOUBLE: { double x = ((const double *)col->data)[a], y = ((const double *)col->data)[b]; c = (x > y) - (x < y); break; }
      case QS_DATE: c = date_cmp((const date_lit *)col->data + a, (const date_lit *)col->data + b); break;
      case QS_CHAR: {   /* strncmp order (types/operations/comparisons/AsciiStringComparators.hpp:218-251) */
        const int r = strncmp((const char *)col->data + a * col->width, (const char *)col->data + b * col->width, col->width);
        c = (r > 0) - (r < 0);
        break;
      }
      default: c = 0;
    }
    if (c) return g_sort_keys[k].descending ? -c : c;
  }
  return (a > b) - (a < b);
}
int64_t qso_topk(const qso_table *t, uint32_t n_keys, const qs_sort_key *keys, uint64_t limit, uint64_t *row_ids) {
  uint64_t *idx = malloc((t->n_rows ? t->n_rows : 1) * 8);
  for (uint64_t i = 0; i < t->n_rows; ++i) idx[i] = i;
  g_sort_table = t; g_sort_keys = keys; g_sort_nkeys = n_keys;
  qsort(idx, t->n_rows, 8, sort_compare);
  const uint64_t n = limit < t->n_rows ? limit : t->n_rows;
  memcpy(row_ids, idx, n * 8);
  free(idx);
  return (int64_t)n;
}

/* ------------------------------------------------------------ K0 decoders */
static inline uint32_t code_at(const void *codes, uint64_t i, uint32_t cw) {
  switch (cw) {
    case 1: return ((const uint8_t *)codes)[i];
    case 2: return ((const uint16_t *)codes)[i];
    default: return ((const uint32_t *)codes)[i];
  }
}
void qso_decode_dict(void *dst, const void *codes, const void *dict, uint64_t n, uint32_t cw, uint32_t vw) {
  for (uint64_t i = 0; i < n; ++i) memcpy((char *)dst + i * vw, (const char *)dict + (uint64_t)code_at(codes, i, cw) * vw, vw);
}
void qso_decode_truncated(void *dst, const void *codes, uint64_t n, uint32_t cw, uint32_t vw) {
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t c = code_at(codes, i, cw);
    if (vw == 8) ((int64_t *)dst)[i] = (int64_t)c; else ((int32_t *)dst)[i] = (int32_t)c;
  }
}
void qso_decode_strided(void *dst, const void *slots, uint64_t n, uint32_t stride, uint32_t vw) {
  for (uint64_t i = 0; i < n; ++i) memcpy((char *)dst + i * vw, (const char *)slots + i * stride, vw);
}

uint32_t qso_partition_of(int64_t key, uint32_t n_parts) {
  uint64_t x = (uint64_t)key;
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return (uint32_t)(((x >> 32) * (uint64_t)n_parts) >> 32);   /* multiply-shift range reduction of the high word */
}
