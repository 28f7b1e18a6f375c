/*
 * qs_oracle.h -- CPU restatement of Quickstep's data-parallel relational
 * operators.  TEST INFRASTRUCTURE ONLY: nothing under quickstep_b200/ links,
 * imports or executes this; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, as the checker or the
 * timed CPU baseline.
 *
 * Each function restates the reference algorithm named in its comment
 * (file:line under /root/reference, UWQuickstep/quickstep @ fee4c630) in plain
 * C: column-at-a-time evaluation with one materialised vector per expression
 * node, TupleIdSequence bitmaps (MSB-first 64-bit words), sequential
 * per-block accumulation merged in block order.
 *
 * Pinning: tests/golden/ holds the outputs of the unmodified reference binary
 * (quickstep_cli_shell) for TPC-H Q1/Q3/Q6 on dbgen data; tests/test_oracle_golden.py
 * checks this oracle against them.
 *
 * The expression encoding (qs_node) and the type / operation ids are the ones
 * of include/qsgpu.h, which are the reference's own enum values.
 */
#ifndef QS_ORACLE_H_
#define QS_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#include "qsgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qso_column {
  const void *data;
  uint16_t type;   /* QS_INT .. QS_DATE */
  uint16_t width;  /* bytes per value */
} qso_column;

typedef struct qso_table {
  const qso_column *cols;
  uint32_t n_cols;
  uint64_t n_rows;
} qso_table;

/* Number of worker threads used by the block-parallel entry points
 * (the reference's --num_workers, cli/Flags.cpp:44-65). */
void qso_set_num_workers(int n);
int qso_get_num_workers(void);
/* Rows per "storage block" = rows per work order (4 MB lineitem blocks hold
 * ~63k tuples). */
void qso_set_block_rows(uint64_t rows);

/* LIP filter (utility/lip_filter/BitVectorExactFilter.hpp:61-176,
 * SingleIdentityHashFilter.hpp:62-171), bits MSB-first in 64-bit words. */
typedef struct qso_lip {
  uint32_t kind;       /* QS_LIP_* */
  uint32_t is_anti;
  int64_t min_value, max_value;
  uint64_t cardinality;
  uint64_t *words;     /* caller-owned, zeroed, qso_lip_words() entries */
} qso_lip;
uint64_t qso_lip_words(const qso_lip *f);
typedef struct qso_lip_ref { qso_lip *lip; uint32_t attr; } qso_lip_ref;

/* ComparisonPredicate / ConjunctionPredicate ::getAllMatches
 * (expressions/predicate/ComparisonPredicate.cpp:115-340,
 * ConjunctionPredicate.cpp:109-143): bitmap of matching rows; returns the
 * number of matches or -1 on an unsupported tree. */
int64_t qso_predicate(const qs_expr_set *ex, int32_t root, const qso_table *t, uint64_t *bitmap_words);

/* Scalar::getAllValues (storage/StorageBlock.cpp:363-388): evaluates the scalar
 * for every row into `out` (native width of the result type); returns the
 * result type id or -1. */
int qso_scalar(const qs_expr_set *ex, int32_t root, const qso_table *t, void *out);

/* BuildLIPFilterWorkOrder::execute (relational_operators/BuildLIPFilterOperator.cpp:146-172). */
int qso_build_lip_filter(const qs_expr_set *ex, int32_t predicate_root, const qso_table *t,
                         uint32_t n_probe, const qso_lip_ref *probe, uint32_t n_build, const qso_lip_ref *build);

/* SelectWorkOrder::execute (relational_operators/SelectOperator.cpp:161-195).
 * out_cols[j] must hold n_rows values of the projected type; returns the
 * number of output rows (input order preserved) or -1. */
int64_t qso_select(const qs_expr_set *ex, int32_t predicate_root, const qso_table *t, uint32_t n_probe,
                   const qso_lip_ref *probe, uint32_t n_project, const int32_t *project_roots,
                   void *const *out_cols);

/* Aggregation: AggregationOperationState::aggregateBlock + finalizeAggregate
 * (storage/AggregationOperationState.cpp:428-948).  Groups come back sorted by
 * their packed key bytes.  out_keys: n_groups * key_bytes; out_values:
 * per aggregate an array of n_groups 8-byte values (int64 or double, see
 * out_is_double); out_null[j] set when a no-GROUP-BY aggregate saw zero rows. */
typedef struct qso_agg_result {
  uint64_t n_groups;
  uint32_t key_bytes;
  uint8_t *keys;            /* malloc'd */
  uint64_t *values;         /* malloc'd, [n_aggregates][n_groups] raw 8-byte words */
  uint8_t is_double[16];
  uint8_t is_null[16];
  int64_t *counts;          /* malloc'd, rows per group */
} qso_agg_result;
int qso_aggregate(const qs_expr_set *ex, int32_t predicate_root, uint32_t n_aggregates,
                  const qs_aggregate *aggregates, uint32_t n_group_by, const int32_t *group_by_roots,
                  const qso_table *t, uint32_t n_probe, const qso_lip_ref *probe, qso_agg_result *out);
void qso_agg_result_free(qso_agg_result *r);

/* BuildHash + HashInnerJoin/Semi/Anti (relational_operators/BuildHashOperator.cpp:162-207,
 * HashJoinOperator.cpp:450-987).  Attribute nodes with b == 2 read the build
 * table.  Output pairs are produced probe-row-major, build rows in build
 * order.  Returns output rows or -1. */
int64_t qso_hash_join(const qs_expr_set *ex, const qso_table *build, int32_t build_predicate_root,
                      uint32_t build_key_attr, const qso_table *probe, int32_t probe_predicate_root,
                      uint32_t probe_key_attr, uint32_t n_probe_lip, const qso_lip_ref *probe_lip,
                      uint32_t join_type, int32_t residual_root, uint32_t n_project,
                      const int32_t *project_roots, void *const *out_cols, uint64_t out_capacity);

/* Same, plus the NULL mask of every output row (bit p = projected column p is NULL): what a LEFT OUTER
 * join (HashOuterJoinWorkOrder, relational_operators/HashJoinOperator.cpp:989-1099) produces for probe
 * tuples without a match -- build-side selections are NULL, stored as zero bytes. */
int64_t qso_hash_join_nulls(const qs_expr_set *ex, const qso_table *build, int32_t build_predicate_root,
                            uint32_t build_key_attr, const qso_table *probe, int32_t probe_predicate_root,
                            uint32_t probe_key_attr, uint32_t n_probe_lip, const qso_lip_ref *probe_lip,
                            uint32_t join_type, int32_t residual_root, uint32_t n_project,
                            const int32_t *project_roots, void *const *out_cols, uint64_t out_capacity,
                            uint64_t *out_nulls);

/* Sort + LIMIT (SortRunGeneration/SortMergeRun): returns the row ids of the
 * first `limit` rows in order. */
int64_t qso_topk(const qso_table *t, uint32_t n_keys, const qs_sort_key *keys, uint64_t limit, uint64_t *row_ids);

/* Storage-format decode (K0 oracle):
 *   dictionary   compression/CompressionDictionaryLite.hpp:40-51
 *   truncation   storage/CompressedBlockBuilder.cpp:434-506
 *   row store    storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179 */
void qso_decode_dict(void *dst, const void *codes, const void *dict, uint64_t n, uint32_t code_width, uint32_t value_width);
void qso_decode_truncated(void *dst, const void *codes, uint64_t n, uint32_t code_width, uint32_t value_width);
void qso_decode_strided(void *dst, const void *slots, uint64_t n, uint32_t stride, uint32_t value_width);

/* Partition id used by the device radix partition (no reference equivalent;
 * stated here so tests can check placement). */
uint32_t qso_partition_of(int64_t key, uint32_t n_parts);

#ifdef __cplusplus
}
#endif
#endif
