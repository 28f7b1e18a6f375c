/*
 * qsgpu.h -- C ABI of libqsgpu.so: the B200 (sm_100a) execution path for
 * Quickstep's data-parallel relational operators.
 *
 * The reference (UWQuickstep/quickstep @ fee4c630) has no FFI: its seam is the
 * pair of C++ virtuals RelationalOperator::getAllWorkOrders()
 * (relational_operators/RelationalOperator.hpp:132) and WorkOrder::execute()
 * (relational_operators/WorkOrder.hpp:251).  This header is what the bodies of
 * GPU work orders call *under* execute(); every entry point cites the
 * reference code it replaces.  Plain pointers and sizes only; every function
 * returns 0 on success and a non-zero qsgpu_status otherwise (the C++ wrapper
 * turns non-zero into LOG(FATAL), the reference's error convention,
 * relational_operators/BuildHashOperator.cpp:205).
 *
 * Type ids, comparison ids and operation ids use the reference's own enum
 * values (types/TypeID.hpp:33-45, types/operations/comparisons/ComparisonID.hpp,
 * types/operations/binary_operations/BinaryOperationID.hpp) so that a
 * serialization::Predicate / serialization::Scalar proto lowers 1:1 into a
 * qs_node array.
 *
 * There is NO CPU fallback behind this ABI: with no CUDA device every compute
 * entry point fails with QSGPU_ERR_NO_DEVICE.
 */
#ifndef QSGPU_H_
#define QSGPU_H_

#include <stddef.h>
#include <stdint.h>

#include "qsgpu_types.h"   /* status codes and the reference's enum values */

#ifdef __cplusplus
extern "C" {
#endif
/* Thread-local text of the last error raised on the calling thread. */
const char *qsgpu_last_error(void);


/*
 * Flattened expression tree: serialization::Predicate and
 * serialization::Scalar (expressions/Expressions.proto:29-137) in one node
 * array.  Children are referenced by index and must precede their parent.
 */
enum {
  QS_N_LITERAL = 0,     /* Scalar LITERAL:  type/width, value in lit           */
  QS_N_ATTRIBUTE = 1,   /* Scalar ATTRIBUTE: a = attribute id, b = join side   */
                        /*   (0 none / probe side, 2 = build side, matching    */
                        /*   ScalarAttribute::JoinSide RIGHT_SIDE)             */
  QS_N_UNARY = 2,       /* op = QS_NEGATE | QS_CAST (to `type`), a = operand   */
  QS_N_BINARY = 3,      /* op = QS_ADD.., a = left, b = right                  */
  QS_N_SHARED = 5,      /* ScalarSharedExpression: a = operand, b = share id   */
  QS_N_TRUE = 16,       /* Predicate TRUE                                      */
  QS_N_FALSE = 17,
  QS_N_COMPARISON = 18, /* op = QS_EQ.., a = left scalar, b = right scalar     */
  QS_N_NEGATION = 19,   /* a = operand predicate                               */
  QS_N_CONJUNCTION = 20,/* a, b = operand predicates (n-ary lists are folded   */
  QS_N_DISJUNCTION = 21 /*   into left-deep binary chains by the caller)       */
};

typedef struct qs_node {
  uint16_t kind;   /* QS_N_*                                                   */
  uint16_t op;     /* comparison / binary / unary id                           */
  uint16_t type;   /* result type of a scalar node (QS_INT..QS_DATE)           */
  uint16_t width;  /* byte width for QS_CHAR (attribute or literal)            */
  int32_t a;
  int32_t b;
  union {
    int32_t i32;
    int64_t i64;
    float f32;
    double f64;
    struct { int32_t year; uint8_t month, day, pad[2]; } date;
    uint64_t pool_offset; /* QS_CHAR literal: offset into the string pool      */
  } lit;
} qs_node;

/* An expression "program": the node array plus the CHAR literal pool. */
typedef struct qs_expr_set {
  const qs_node *nodes;
  uint32_t n_nodes;
  const char *str_pool;
  uint32_t str_pool_bytes;
} qs_expr_set;

/* --------------------------------------------------------- runtime/devices */
/* Called once per process before anything else (cli/QuickstepCli.cpp:183-284
 * is where the reference brings up its workers).  dev_ids == NULL means
 * devices 0..n_dev-1; n_dev == 0 means "all visible". */
int qsgpu_init(int n_dev, const int *dev_ids);
int qsgpu_shutdown(void);
int qsgpu_device_count(int *n_dev);
/* Block until all work queued on `dev` by this library has completed. */
int qsgpu_synchronize(int dev);
/* The CUDA stream (cudaStream_t) all work of this library on `dev` is queued on.  A caller that runs its
 * own device work between two calls (e.g. an NCCL collective on partial aggregation states) can queue it
 * on this stream instead of synchronising the device around it. */
int qsgpu_stream(int dev, void **stream);
/* Number of kernels this library has launched since init (bench.py's
 * gpu_launches); never reset by the library. */
int qsgpu_launch_count(uint64_t *n);

/* Raw device memory (StorageManager-side allocations for staged columns). */
int qsgpu_malloc(int dev, size_t bytes, void **dptr);
int qsgpu_free(int dev, void *dptr);
int qsgpu_memcpy_h2d(int dev, void *dst, const void *src, size_t bytes);
int qsgpu_memcpy_d2h(int dev, void *dst, const void *src, size_t bytes);
int qsgpu_memcpy_d2d(int dev, void *dst, const void *src, size_t bytes);
/* Same copy, only enqueued on the library stream of `dev`: call qsgpu_synchronize
 * before another stream (e.g. NCCL's) touches dst. */
int qsgpu_memcpy_d2d_async(int dev, void *dst, const void *src, size_t bytes);
/* Pinned host memory for staging buffers. */
int qsgpu_host_alloc(size_t bytes, void **hptr);
int qsgpu_host_free(void *hptr);

/* ---------------------------------------------- device-resident relations */
/*
 * A device relation is the HBM image of a CatalogRelation's blocks: one
 * contiguous, 256-byte aligned, native-width buffer per attribute (the
 * ColumnVector-equivalent the north star asks for), padded so that 16 bytes
 * past the last row are always readable.  Temporary relations produced by
 * Select / HashJoin / FinalizeAggregation are device relations too (the
 * analogue of the temporary SplitRowStore relations written through
 * InsertDestination, storage/InsertDestination.cpp:202-216).
 */
typedef struct qsgpu_relation *qsgpu_relation_t;

typedef struct qs_attr {
  uint16_t type;   /* QS_INT..QS_DATE */
  uint16_t width;  /* bytes per value (4, 8, 4, 8, n, -, 8) */
} qs_attr;

int qsgpu_relation_create(int dev, uint32_t n_attrs, const qs_attr *attrs,
                          uint64_t capacity_rows, qsgpu_relation_t *out);
int qsgpu_relation_destroy(qsgpu_relation_t rel);
int qsgpu_relation_num_rows(qsgpu_relation_t rel, uint64_t *n_rows);
int qsgpu_relation_set_num_rows(qsgpu_relation_t rel, uint64_t n_rows);
/* Device pointer of attribute `attr` (row 0). */
int qsgpu_relation_column(qsgpu_relation_t rel, uint32_t attr, void **dptr);
/* Wrap caller-owned device buffers (e.g. generated in place) without copy.
 * Buffers must be 16-byte aligned and readable 16 bytes past the last row. */
int qsgpu_relation_wrap(int dev, uint32_t n_attrs, const qs_attr *attrs,
                        void *const *dptrs, uint64_t n_rows,
                        qsgpu_relation_t *out);
/* Copy rows [row_begin, row_begin+n_rows) of one attribute to the host
 * (InsertDestination::bulkInsertTuples direction). */
int qsgpu_relation_read(qsgpu_relation_t rel, uint32_t attr, uint64_t row_begin,
                        uint64_t n_rows, void *host_out);

/*
 * Dictionary-coded attributes: the compressed block format as a device format.
 *
 * The reference evaluates a comparison against a literal directly on the codes of a dictionary-compressed
 * stripe: CompressedTupleStorageSubBlock::getMatchesForPredicate turns the literal into limit codes of the
 * block's sorted dictionary and scans the 1/2/4-byte codes (storage/CompressedTupleStorageSubBlock.cpp:160-251,
 * compression/CompressionDictionary.hpp:241-318); values are materialised through
 * CompressionDictionaryLite::getUntypedValueForCode only where a scalar needs them
 * (compression/CompressionDictionaryLite.hpp:40-51).
 *
 * qsgpu_relation_set_dictionary declares that attribute `attr` of `rel` is resident in HBM as `code_width`-byte
 * codes (1, 2 or 4) into ONE relation-wide dictionary of `n_entries` native values, strictly increasing in the
 * attribute's order (CompressionDictionaryBuilder order: numeric / DateLit / strncmp).  Call it on an empty
 * relation from qsgpu_relation_create (its column is re-allocated at code width; qsgpu_stage_blocks then re-codes
 * the blocks' QS_ENC_DICT stripes from their per-block dictionaries into relation codes), or on a relation from
 * qsgpu_relation_wrap whose buffer for `attr` already holds the codes.  From then on every scan of the relation
 * (select, aggregation, join build / probe, LIP build) moves code_width bytes per row of that attribute:
 * comparisons with literals run on the codes (one unsigned range test per row), scalar expressions look values
 * up in the dictionary, and group-by keys / pass-through projections / join keys are decoded per tile in shared
 * memory.  Results are bit-identical to the native column's.  qsgpu_relation_read returns native values;
 * qsgpu_relation_column returns the code buffer.  Operators that read whole native columns of their input
 * (top-k, partitioning, build-side projections of a join) refuse coded attributes with QSGPU_ERR_UNSUPPORTED.
 */
int qsgpu_relation_set_dictionary(qsgpu_relation_t rel, uint32_t attr, uint32_t code_width,
                                  const void *dict_values, uint32_t n_entries);
/*
 * The codes of a sorted dictionary that satisfy `attribute <cmp> literal` (cmp = QS_EQ..QS_GE, `literal` a
 * QS_N_LITERAL node; CHAR literals take their bytes from str_pool): the range [*first, *first + *count), or its
 * complement when *negate is set.  Host arithmetic only, no device needed -- it is what the lowering applies to every
 * comparison on a coded attribute, with the reference's type promotion (an INT attribute against a DOUBLE literal
 * compares as DOUBLE) and NaN semantics; a binding can use it to prune blocks the way
 * CompressedTupleStorageSubBlock::getMatchesForPredicate short-cuts an always-false comparison
 * (storage/CompressedTupleStorageSubBlock.cpp:183-199): *count == 0 without *negate means no tuple can match.
 */
int qsgpu_dictionary_code_range(uint16_t attr_type, uint16_t attr_width, const void *dict_values, uint32_t n_entries,
                                uint32_t cmp, const qs_node *literal, const char *str_pool, uint32_t str_pool_bytes,
                                uint32_t *first, uint32_t *count, int *negate);
/* code_width 0 = native attribute; dict_out (optional) receives n_entries native values. */
int qsgpu_relation_dictionary(qsgpu_relation_t rel, uint32_t attr, uint32_t *code_width, uint32_t *n_entries,
                              void *dict_out);

/* NULL mask of rows [row_begin, row_begin+n_rows): bit j of word i = attribute j of row row_begin+i is NULL.
 * All zeros for a relation without NULL-able attributes. */
int qsgpu_relation_read_nulls(qsgpu_relation_t rel, uint64_t row_begin, uint64_t n_rows,
                              uint64_t *host_out);
/*
 * NULL-able attributes of a base relation.  The reference types every attribute as nullable or not
 * (types/Type.hpp:129 isNullable) and picks NULL-checking code paths from that; qsgpu_relation_set_nullable declares
 * the set (bit a = attribute a may be NULL) before the relation is staged or written.  From then on every operator
 * that reads one of these attributes also reads the relation's per-row NULL mask (8 bytes per row, moved by the same
 * TMA pipeline as the columns):
 *   - a comparison with a NULL operand is false (LiteralComparators-inl.hpp:168-223), NOT complements it
 *     (NegationPredicate::getAllMatches), arithmetic over a NULL is NULL;
 *   - SUM / AVG / MIN / MAX / COUNT(x) skip NULL arguments (AggregationHandleSum.hpp:117-127) and are NULL (COUNT: 0)
 *     for a group without a non-NULL argument -- declare such aggregates in qs_agg_spec.nullable_arguments;
 *   - rows with a NULL join key neither enter a join table nor match (storage/HashTable.hpp:1384,1903): inner / semi
 *     joins drop such a probe row, an anti join emits it and a left outer join emits it NULL-padded (:1999-2003);
 *     they are not inserted into / are rejected by LIP filters -- anti filters included (filterBatchInternal<true> skips a
 *     NULL value before it looks at the filter, utility/lip_filter/BitVectorExactFilter.hpp:130-146);
 *   - Select and the probe side of an inner join carry the NULL-ness of what they project into the output relation.
 *   - rows whose GROUP BY key is NULL belong to no group (storage/PackedPayloadHashTable.hpp:861-866: the reference
 *     prints no NULL group).
 *   - NULL-able attributes of a join's build side are read through the matched build row's mask (projections, scalars
 *     and the residual predicate alike).
 *   - ORDER BY (qsgpu_topk) puts NULLs first or last as qs_sort_key says.
 * Refused with QSGPU_ERR_UNSUPPORTED (never silently wrong): partitioning on a NULL-able attribute, and NULL-able
 * attributes held as relation-wide dictionary codes.
 * qsgpu_relation_write_nulls sets the masks of rows written through qsgpu_relation_column / qsgpu_relation_wrap
 * (their stored values should be zero bytes); qsgpu_stage_blocks fills them from the block formats' own NULL
 * representations (qs_stage_desc.null_kind).
 */
int qsgpu_relation_set_nullable(qsgpu_relation_t rel, uint64_t attr_mask);
int qsgpu_relation_nullable(qsgpu_relation_t rel, uint64_t *attr_mask);
int qsgpu_relation_write_nulls(qsgpu_relation_t rel, uint64_t row_begin, uint64_t n_rows, const uint64_t *masks);

/* All attributes at once: host_out[a] receives rows [row_begin, row_begin+n_rows) of attribute a; the
 * copies are queued back to back and waited for once (result relations are a handful of rows wide and
 * tall: one synchronisation instead of one per column). */
int qsgpu_relation_read_all(qsgpu_relation_t rel, uint64_t row_begin, uint64_t n_rows,
                            void *const *host_out);

/*
 * A small RESULT relation in one round trip: up to max_rows rows of every attribute (host_out[a] has room for
 * max_rows values), the relation's row count (*n_rows; may exceed max_rows, in which case only max_rows rows were
 * copied) and, when null_masks != NULL, the rows' NULL masks -- packed by one kernel, moved by ONE device-to-host
 * copy and waited for once.  Operators upstream only enqueue (the row count of a temporary relation lives on the
 * device), so this is the single host wait of a query; an error a kernel raised on the way (QSGPU_ERR_CAPACITY)
 * is reported here.  Limited to results of <= 256 KB; larger ones are read with qsgpu_relation_read.
 */
int qsgpu_relation_read_rows(qsgpu_relation_t rel, uint64_t max_rows, void *const *host_out, uint64_t *n_rows,
                             uint64_t *null_masks);

/*
 * K0 -- staging of one storage block's attribute into a device relation,
 * appended at the relation's current end.  Physical encodings of the
 * reference's sub-blocks:
 *   QS_ENC_PLAIN      BasicColumnStore stripe
 *                     (storage/BasicColumnStoreTupleStorageSubBlock.cpp:100-183)
 *   QS_ENC_STRIDED    fixed-width attribute inside SplitRowStore tuple slots
 *                     (storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179):
 *                     value i at host + i*stride
 *   QS_ENC_DICT       CompressedColumnStore stripe of 1/2/4-byte codes into an
 *                     ordered dictionary (compression/CompressionDictionaryLite.hpp:40-51,
 *                     storage/CompressedColumnStoreTupleStorageSubBlock.cpp:203-215)
 *   QS_ENC_TRUNCATED  CompressedColumnStore stripe of a non-negative INT/LONG
 *                     truncated to 1/2/4 bytes
 *                     (storage/CompressedBlockBuilder.cpp:434-506)
 * The block is decoded to native width on the device by the K0 kernels.
 */
/* enum { QS_ENC_PLAIN, QS_ENC_STRIDED, QS_ENC_DICT, QS_ENC_TRUNCATED }: qsgpu_types.h */

typedef struct qs_stage_desc {
  uint32_t attr;         /* target attribute                                   */
  uint32_t encoding;     /* QS_ENC_*                                           */
  const void *host;      /* host stripe / first slot                           */
  uint32_t code_width;   /* DICT/TRUNCATED: 1, 2 or 4                          */
  uint32_t stride;       /* STRIDED: tuple slot bytes                          */
  const void *dict;      /* DICT: values array (native width, sorted)          */
  uint32_t dict_entries; /* DICT: number of codes                              */
  /* NULL values of the stripe (QS_NULL_*, qsgpu_types.h; batched staging only).  The attribute must have been
   * declared with qsgpu_relation_set_nullable.  A NULL value is stored as zero bytes and its bit is set in the
   * relation's per-row NULL mask. */
  uint32_t null_kind;
  uint32_t null_arg;        /* CODE: the NULL code; BITMAP: bit of row 0; SLOT_WORD: bit inside the word, from the MSB */
  uint32_t null_stride;     /* BITMAP: bits from one row to the next; SLOT_WORD: bytes from one row's word to the next */
  uint32_t null_width;      /* SLOT_WORD: bytes of the word (1, 2, 4 or 8)                                             */
  uint32_t reserved;
  const void *null_bitmap;  /* BITMAP / SLOT_WORD: inside the block image                                              */
} qs_stage_desc;

/* Stage `n_desc` attributes of one block of `n_rows` tuples; all attributes
 * of the relation must be staged by the same call or by calls made before
 * the next work order reads it. */
int qsgpu_stage_block(qsgpu_relation_t rel, uint64_t n_rows,
                      const qs_stage_desc *descs, uint32_t n_desc);

/*
 * Batched form for a run of storage blocks (what a GPU work order that covers
 * many 4 MB blocks stages at once).  Each block is described by the extent of
 * its memory (StorageBlock::getMemory / getSize, storage/StorageBlockBase.hpp)
 * and the per-attribute stripes inside it; the whole image travels in ONE
 * host-to-device copy (runs of blocks that are contiguous in host memory are
 * merged into one copy), and ONE kernel decodes every stripe of the batch to
 * native width.  descs[i].host / .dict must point inside [host, host+bytes).
 */
typedef struct qs_block_image {
  const void *host;            /* first byte of the block image                 */
  uint64_t bytes;              /* bytes of the image to copy                    */
  uint64_t n_rows;             /* tuples in the block                           */
  const qs_stage_desc *descs;  /* n_desc entries, one per attribute             */
} qs_block_image;

int qsgpu_stage_blocks(qsgpu_relation_t rel, uint32_t n_blocks,
                       const qs_block_image *blocks, uint32_t n_desc);
/*
 * Column pruning: a descriptor with encoding QS_ENC_SKIP leaves its attribute on the host (neither copied
 * nor decoded; the device column keeps whatever it held).  qsgpu_stage_columns stages further attributes of
 * blocks whose rows already exist on the device: the same blocks, in the same order, starting at row
 * `first_row` -- the device cache is keyed by (block, attribute), so a query pays only for the attributes its
 * operators reference, and a later query that needs more stages just those.
 */
int qsgpu_stage_columns(qsgpu_relation_t rel, uint64_t first_row, uint32_t n_blocks,
                        const qs_block_image *blocks, uint32_t n_desc);

/* ----------------------------------------------------------- LIP filters  */
/*
 * K4.  utility/lip_filter/BitVectorExactFilter.hpp:61-176 (bit v-min; probe
 * out of range -> is_anti) and SingleIdentityHashFilter.hpp:62-171 (bit
 * (uint64)v % cardinality).  Bits live in 64-bit words, MSB first, exactly as
 * BarrieredReadWriteConcurrentBitVector (utility/
 * BarrieredReadWriteConcurrentBitVector.hpp:128-147) so a device filter can be
 * compared with (or handed to) the host structure bit for bit.
 */
typedef struct qsgpu_lip *qsgpu_lip_t;

int qsgpu_lip_create(int dev, uint32_t kind, uint32_t attr_type /*QS_INT|QS_LONG*/,
                     int64_t min_value, int64_t max_value, /* exact filter     */
                     uint64_t cardinality,                 /* identity hash    */
                     int is_anti, qsgpu_lip_t *out);
int qsgpu_lip_destroy(qsgpu_lip_t lip);
int qsgpu_lip_num_words(qsgpu_lip_t lip, uint64_t *n_words);
int qsgpu_lip_read(qsgpu_lip_t lip, uint64_t *host_words);
/* Device address of the bit words (for NCCL all-reduce(BOR) across GPUs). */
int qsgpu_lip_device_words(qsgpu_lip_t lip, void **dptr);
/*
 * LIPFilterAdaptiveProber (utility/lip_filter/LIPFilterAdaptiveProber.hpp:89-232) keeps, per filter, how many tuples
 * probed it and how many it rejected, and re-sorts the filters by miss rate between batches.  Every scan kernel
 * that probes a filter adds its counts to the filter (rows already rejected by an earlier filter of the same scan
 * are not probed and not counted); a caller with several probe filters passes them most-selective-first in the
 * qs_scan of its next work order.  The order never changes a result.  Waits for the queued work.
 */
int qsgpu_lip_probe_stats(qsgpu_lip_t lip, uint64_t *probes, uint64_t *misses);

/* A filter bound to the attribute it is built from / probed with
 * (LIPFilterDeployment, utility/lip_filter/LIPFilterDeployment.cpp). */
typedef struct qs_lip_ref {
  qsgpu_lip_t lip;
  uint32_t attr;      /* attribute of the scanned relation */
  uint32_t reserved;
} qs_lip_ref;

/* ------------------------------------------------------------ scan source */
/* What every operator scans: a row range of a device relation, an optional
 * predicate (root index into exprs, -1 = none) and LIP filters to probe. */
typedef struct qs_scan {
  qsgpu_relation_t input;
  uint64_t row_begin, row_end;     /* row_end == UINT64_MAX -> all rows       */
  const qs_expr_set *exprs;        /* may be NULL when nothing refers to it   */
  int32_t predicate_root;          /* -1: no predicate                        */
  uint32_t n_lip_probe;
  const qs_lip_ref *lip_probe;     /* LIPFilterAdaptiveProber::filterValueAccessor */
} qs_scan;

/* --------------------------------------------------------- BuildLIPFilter */
/* BuildLIPFilterWorkOrder::execute (relational_operators/
 * BuildLIPFilterOperator.cpp:146-172): predicate -> probe upstream filters ->
 * insert survivors into the target filters. */
int qsgpu_build_lip_filter(const qs_scan *scan, uint32_t n_build,
                           const qs_lip_ref *build);

/* ------------------------------------------------------------------ Select */
/*
 * K3.  SelectWorkOrder::execute (relational_operators/SelectOperator.cpp:161-195):
 * predicate -> LIP probe -> projection -> output relation.  Projection column
 * j is the scalar rooted at project_roots[j] (a bare attribute node is the
 * selectSimple path, StorageBlock.cpp:390-398).  Rows are appended to `output`
 * (rows of one tile stay in input order; tiles interleave, like blocks of
 * concurrent work orders do in the reference).
 */
int qsgpu_select(const qs_scan *scan, uint32_t n_project,
                 const int32_t *project_roots, qsgpu_relation_t output);

/* ------------------------------------------------------------- aggregation */
/*
 * AggregationOperationState (storage/AggregationOperationState.hpp:72-320):
 * created once per query from the serialized description
 * (query_execution/QueryContext.cpp:66-77), fed by one qsgpu_agg_run per work
 * order (AggregationWorkOrder::execute -> aggregateBlock,
 * storage/AggregationOperationState.cpp:428-474), finalized by
 * FinalizeAggregationWorkOrder (…:641-948) and freed by
 * DestroyAggregationStateWorkOrder.
 *
 * Strategy (ExecutionGenerator.cpp:1924-1965 decides in the reference):
 *   QS_AGG_SINGLE_STATE     no GROUP BY        (K1; aggregateBlockSingleState)
 *   QS_AGG_COMPACT_KEY      keys total <= 8 B  (K2; ThreadPrivateCompactKeyHashTable)
 *   QS_AGG_SEPARATE_CHAINING generic keys <= 32 B (K7; PackedPayloadHashTable)
 *   QS_AGG_COLLISION_FREE   one INT/LONG key used as the array index
 *                                              (K7d; CollisionFreeVectorTable)
 */

typedef struct qs_aggregate {
  uint32_t function;     /* QS_AGG_*                                           */
  int32_t argument_root; /* scalar root in exprs; -1 for COUNT(*)              */
} qs_aggregate;

typedef struct qs_agg_spec {
  int dev;
  uint32_t strategy;
  const qs_expr_set *exprs;          /* predicate + arguments + group-by      */
  int32_t predicate_root;            /* -1: none                              */
  uint32_t n_aggregates;
  const qs_aggregate *aggregates;
  uint32_t n_group_by;
  const int32_t *group_by_roots;     /* attribute nodes                       */
  uint64_t estimated_num_entries;    /* table sizing (proto field 5)          */
  int64_t collision_free_max_key;    /* QS_AGG_COLLISION_FREE: num_entries-1  */
  /* bit j: the argument of aggregate j has a NULL-able type.  The reference picks the handle's code path from the
   * argument type (AggregateFunctionSum::createHandle; AggregationHandleSum.hpp:117-127 skips NULL values, the AVG
   * and COUNT(x) handles count the non-NULL ones); here such an aggregate gets one more state word, its count of
   * non-NULL arguments, which finalization uses instead of the group's row count.                             */
  uint64_t nullable_arguments;
} qs_agg_spec;

typedef struct qsgpu_agg_state *qsgpu_agg_state_t;

int qsgpu_agg_create(const qs_agg_spec *spec, qsgpu_agg_state_t *out);
/* One work order: aggregate rows [row_begin,row_end) of `input`; the state's
 * own predicate is applied, plus the LIP probes given here. */
int qsgpu_agg_run(qsgpu_agg_state_t state, qsgpu_relation_t input,
                  uint64_t row_begin, uint64_t row_end,
                  uint32_t n_lip_probe, const qs_lip_ref *lip_probe);
/* Number of groups currently in the state (1 for SINGLE_STATE). */
int qsgpu_agg_num_groups(qsgpu_agg_state_t state, uint64_t *n_groups);
/*
 * Raw partial state for cross-GPU merging (AggregationHandle::mergeStates):
 * a device array of n_groups x (n_aggregates + 1) 64-bit words, row-major,
 * word 0 = row count of the group, word 1+j = value word of aggregate j
 * (int64 for integer SUM / COUNT / integer MIN,MAX; double otherwise), and the
 * packed group keys (n_groups x key_words 64-bit words).
 */
int qsgpu_agg_partial(qsgpu_agg_state_t state, void **d_states, void **d_keys,
                      uint64_t *n_groups, uint32_t *words_per_group,
                      uint32_t *key_words);
/* The same pointers and layout WITHOUT waiting for queued work orders, for the strategies whose state is a
 * fixed-size array (SINGLE_STATE: 1 row, COMPACT_KEY: 256 rows; unused rows have a zero row count and are
 * skipped by qsgpu_agg_merge_partial).  Contents are valid in stream order after the queued work orders. */
int qsgpu_agg_partial_layout(qsgpu_agg_state_t state, void **d_states, void **d_keys,
                             uint64_t *rows, uint32_t *words_per_group, uint32_t *key_words);
/* Merge a partial state (device pointers on the state's device, same layout
 * as qsgpu_agg_partial returns) into `state`. */
int qsgpu_agg_merge_partial(qsgpu_agg_state_t state, const void *d_states,
                            const void *d_keys, uint64_t n_groups);
/*
 * CollisionFreeVectorTable::getExistenceMap (storage/CollisionFreeVectorTable.hpp:123-128), COLLISION_FREE
 * states only: an exact bit-vector filter over the table's key range, owned by the state.
 * BuildAggregationExistenceMapWorkOrder::execute (relational_operators/BuildAggregationExistenceMapOperator.cpp:176-212)
 * is qsgpu_build_lip_filter with this filter as the build target and the build attribute as its key; a group
 * whose bit is set is finalized even when no input row reached it (COUNT 0, SUM 0) -- the fused
 * "LEFT OUTER JOIN ... GROUP BY left key" plans of ExecutionGenerator.cpp:2142-2180.
 */
int qsgpu_agg_existence_map(qsgpu_agg_state_t state, qsgpu_lip_t *out);
/*
 * finalizeAggregate: output relation gets one row per group: the group-by
 * attributes in order, then one column per aggregate (SUM(int)->LONG,
 * SUM(float/double)->DOUBLE, AVG->DOUBLE, COUNT->LONG, MIN/MAX->argument
 * type).  `*out` is created by the call.  For an aggregate over zero rows
 * (SQL NULL, AggregationHandleSum.cpp:134-143) the value is 0 and the
 * matching bit of *null_mask (bit j = aggregate j, SINGLE_STATE only) is set; the same information is in the
 * output relation's per-row NULL mask (qsgpu_relation_read_nulls / read_rows: bit = output column).
 * NULL-able arguments (qs_agg_spec.nullable_arguments) follow what the reference engine prints: without GROUP BY an
 * aggregate that saw no non-NULL value is NULL (COUNT(x): 0); with GROUP BY only MIN / MAX are, while SUM is 0 and
 * AVG is 0 / 0.0 = NaN for such a group (the hash-table payload of SUM / AVG is the bare running sum,
 * AggregationHandleSum.hpp:176-178, AggregationHandleAvg.hpp:180-189).
 * With null_mask == NULL the call only enqueues, for every strategy: the live group count is read on the device (the
 * state's group counter, or the length of the list of occupied table slots a collect kernel writes), the output is
 * sized for a bound the host knows (256 groups; the table's slots / the rows a hash table was fed) and its row count
 * stays device-side until somebody asks -- so a query's tail (finalize, the wrapping Select, the sort) is queued while
 * the scan kernel is still running.
 */
int qsgpu_agg_finalize(qsgpu_agg_state_t state, qsgpu_relation_t *out,
                       uint64_t *null_mask);
int qsgpu_agg_destroy(qsgpu_agg_state_t state);

/* --------------------------------------------------------------- hash join */
/*
 * K5/K6.  JoinHashTable (storage/HashTable.hpp:1284) keyed by one INT/LONG
 * attribute, duplicates allowed.  The stored value is the build row id (the
 * TupleReference of BuildHashOperator.cpp:48-61); projected build-side
 * attributes are gathered through it at probe time.
 */
typedef struct qsgpu_join_table *qsgpu_join_table_t;

int qsgpu_join_create(int dev, uint32_t key_type /*QS_INT|QS_LONG*/,
                      uint64_t estimated_num_entries, qsgpu_join_table_t *out);
/* Dense ("CollisionFreeVector-style") table for build keys bounded by exact statistics [min_key, max_key]
 * (what \analyze records, catalog/CatalogRelationStatistics.hpp): heads[key - min_key] -> chain of build
 * rows.  No hashing and no key comparison; duplicates are chained.  Same build / probe entry points; a
 * build key outside the declared range raises QSGPU_ERR_INVALID. */
int qsgpu_join_create_dense(int dev, uint32_t key_type, int64_t min_key, int64_t max_key,
                            qsgpu_join_table_t *out);
/* BuildHashWorkOrder::execute (BuildHashOperator.cpp:162-207): predicate ->
 * LIP build -> put(key -> row id).  The build relation must outlive the table. */
int qsgpu_join_build(qsgpu_join_table_t table, const qs_scan *scan,
                     uint32_t key_attr, uint32_t n_lip_build,
                     const qs_lip_ref *lip_build);
int qsgpu_join_num_entries(qsgpu_join_table_t table, uint64_t *n);
/* Composite join key (HashTable::putValueAccessorCompositeKey / getAllFromValueAccessorCompositeKey,
 * storage/HashTable.hpp:1469,2183): n_keys = 2 INT attributes, compared component-wise; the table must be an
 * open-addressing table created with QS_LONG keys (the pair is packed into one 64-bit key).  n_keys = 1 is
 * qsgpu_join_build / qsgpu_join_probe. */
int qsgpu_join_build_composite(qsgpu_join_table_t table, const qs_scan *scan, uint32_t n_keys,
                               const uint32_t *key_attrs, uint32_t n_lip_build,
                               const qs_lip_ref *lip_build);
int qsgpu_join_probe_composite(qsgpu_join_table_t table, const qs_scan *probe, uint32_t n_keys,
                               const uint32_t *probe_key_attrs, uint32_t join_type,
                               int32_t residual_root, uint32_t n_project,
                               const int32_t *project_roots, qsgpu_relation_t output);
/*
 * HashInnerJoinWorkOrder / Semi / Anti / Outer (HashJoinOperator.cpp:450-1099): LIP
 * probe -> hash probe -> residual predicate over both sides -> projection.
 * In `exprs`, attribute nodes with b == 2 refer to the build relation.
 * residual_root == -1: none.  Output rows are appended to `output`.
 * QS_JOIN_LEFT_OUTER (no residual predicate, as in the reference): probe rows without a match are emitted
 * too, their build-side projections NULL (zero bytes + the row's bit in qsgpu_relation_read_nulls).
 */
int qsgpu_join_probe(qsgpu_join_table_t table, const qs_scan *probe,
                     uint32_t probe_key_attr, uint32_t join_type,
                     int32_t residual_root, uint32_t n_project,
                     const int32_t *project_roots, qsgpu_relation_t output);
int qsgpu_join_destroy(qsgpu_join_table_t table);

/* ----------------------------------------------------------------- top-k   */
/* K9.  SortRunGeneration + SortMergeRun with LIMIT (§8f row 1): order rows of
 * `input` by up to 4 sort attributes and keep the first `limit` rows. */
/* descending: bit 0 = DESC.  Bits 1-2 order NULLs of a NULL-able sort attribute: QS_SORT_NULLS_FIRST / QS_SORT_NULLS_LAST
 * as written in the query, 0 = the reference's default (NULLs first iff descending, parser/ParseOrderBy.hpp:53-66). */
#define QS_SORT_DESCENDING 1u
#define QS_SORT_NULLS_FIRST 2u
#define QS_SORT_NULLS_LAST 4u
typedef struct qs_sort_key { uint32_t attr; uint32_t descending; } qs_sort_key;
int qsgpu_topk(qsgpu_relation_t input, uint32_t n_keys, const qs_sort_key *keys,
               uint64_t limit, qsgpu_relation_t *out);

/* ------------------------------------------------------- radix partition   */
/*
 * K8.  Partition rows of `input` by hash(key attr) mod n_parts into
 * `output` (same schema, rows grouped by partition) and report the partition
 * start offsets (n_parts + 1 entries, host).  The exchange that follows is an
 * all-to-all over NVLink (PartitionAwareInsertDestination is the reference's
 * repartitioning sink, storage/InsertDestination.cpp:471-722).
 */
int qsgpu_radix_partition(qsgpu_relation_t input, uint32_t key_attr,
                          uint32_t n_parts, qsgpu_relation_t output,
                          uint64_t *host_offsets);

/* Same regrouping with the reference's OWN partition function for a relation hash-partitioned on one INT / LONG
 * attribute (HashPartitionSchemeHeader::getPartitionId, catalog/PartitionSchemeHeader.hpp:200-214, over the identity
 * hash of an inline scalar, types/TypedValue.hpp:575-607): partition = value & (n_parts - 1) when n_parts is a power of
 * two, value % n_parts otherwise.  What a PartitionAwareInsertDestination (storage/InsertDestination.cpp:471-722)
 * does tuple by tuple; rows land in the partition the reference would put them in. */
int qsgpu_hash_partition(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts, qsgpu_relation_t output,
                         uint64_t *host_offsets);

/* Same regrouping by KEY RANGE: partition p = clamp((key - min_key) / part_width, 0, n_parts-1).  Used to make
 * a large join cache-resident (radix join): build and probe sides are range-partitioned, so the slice of the
 * dense join table and of the build relation that one probe partition touches is contiguous and fits in L2. */
int qsgpu_range_partition(qsgpu_relation_t input, uint32_t key_attr, int64_t min_key,
                          uint64_t part_width, uint32_t n_parts, qsgpu_relation_t output,
                          uint64_t *host_offsets);

/*
 * K8 fused with the all-to-all that follows it (one process per GPU, peers mapped over NVLink with CUDA IPC):
 *   1. every rank counts its rows per destination      qsgpu_partition_count      (partition id = hash, as above)
 *   2. the ranks exchange the counts (a few words) and derive, for every destination, the first row each
 *      sender writes at (exclusive prefix over the senders)
 *   3. every rank scatters its rows STRAIGHT INTO the destinations' receive relations
 *                                                      qsgpu_partition_scatter_peers
 *      peer_cols[p * n_attrs + a] = device address of attribute a of destination p's receive relation (the
 *      local one for p == own rank, an IPC mapping otherwise); rows of destination p start at first_rows[p].
 *   4. a barrier across ranks; the receive relations then hold the shuffled rows.
 * The partition kernel IS the transfer: no send buffers, no separate collective, NVLink traffic overlapped with
 * the partitioning itself.  qsgpu_ipc_* give device memory other processes can map.
 */
typedef struct qs_ipc_handle { unsigned char bytes[64]; } qs_ipc_handle;
int qsgpu_ipc_alloc(int dev, size_t bytes, void **dptr, qs_ipc_handle *handle);
int qsgpu_ipc_open(int dev, const qs_ipc_handle *handle, void **dptr);
int qsgpu_ipc_close(int dev, void *dptr);
int qsgpu_ipc_free(int dev, void *dptr);
int qsgpu_partition_count(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts,
                          uint64_t *host_counts);
int qsgpu_partition_scatter_peers(qsgpu_relation_t input, uint32_t key_attr, uint32_t n_parts,
                                  void *const *peer_cols, const uint64_t *first_rows);

/* The partitioning that makes a join on `table` cache-resident (radix join): rows are grouped by the slice of
 * the table their key lands in -- leading bits of the home slot for an open-addressing table, equal key ranges
 * for a dense one.  Partition the build relation before qsgpu_join_build and the probe relation before
 * probing one partition (row range) at a time; n_parts must be a power of two. */
int qsgpu_join_partition(qsgpu_join_table_t table, qsgpu_relation_t input, uint32_t key_attr,
                         uint32_t n_parts, qsgpu_relation_t output, uint64_t *host_offsets);

/* ----------------------------------------------------------------- multi-GPU */
/*
 * One process per GPU; partition id <-> device id.  The reference keeps ONE aggregation state / hash table / LIP
 * filter per query (or per partition) inside one process (query_execution/QueryContext.cpp:66-97) and merges
 * per-thread partial states with AggregationHandle*::mergeStates
 * (expressions/aggregation/AggregationHandleSum.cpp:109-117) and ThreadPrivateCompactKeyHashTable::mergeFrom
 * (storage/ThreadPrivateCompactKeyHashTable.cpp:306-363).  Across GPUs the same merges are collectives over
 * NVLink (NCCL, resolved at run time: libnccl.so.2, or the path in QSGPU_NCCL_LIB), queued on the library
 * stream behind the kernels that produce their input.  Every rank must issue the same collectives in the same
 * order.  A NULL communicator (or one of a single rank) turns each call into a no-op / local copy.
 */
typedef struct qs_comm_id { unsigned char bytes[128]; } qs_comm_id;    /* an ncclUniqueId */
typedef struct qsgpu_comm *qsgpu_comm_t;
/* Rank 0 creates the id; the caller carries it to the other ranks (any host channel), then every rank creates
 * its communicator for its device.  qsgpu_init must have been called for `dev`. */
int qsgpu_comm_unique_id(qs_comm_id *id);
int qsgpu_comm_create(int dev, int rank, int n_ranks, const qs_comm_id *id, qsgpu_comm_t *out);
int qsgpu_comm_destroy(qsgpu_comm_t comm);
int qsgpu_comm_rank(qsgpu_comm_t comm, int *rank, int *n_ranks);
/* *enabled = 1 when the ranks of `comm` have mapped one another's mailbox (CUDA IPC over NVLink / NVSwitch, set up by
 * qsgpu_comm_create when every rank can reach every other; QSGPU_PEER_MERGE=0 turns it off): qsgpu_agg_merge_all then
 * merges SINGLE_STATE / COMPACT_KEY states with ONE kernel that stores its partial state into the peers' memory,
 * waits for theirs and folds -- no NCCL call on that path.  0: every collective goes through NCCL. */
int qsgpu_comm_peer_memory(qsgpu_comm_t comm, int *enabled);
/* All ranks have finished the work queued so far (device-side all-reduce of one word, then a host wait). */
int qsgpu_comm_barrier(qsgpu_comm_t comm);
/* Host values reduced over all ranks in place (op: 0 = sum, 1 = min, 2 = max): global row counts and the exact
 * min / max statistics \analyze records, when every rank loaded only its partition. */
int qsgpu_comm_allreduce_i64(qsgpu_comm_t comm, int64_t *values, uint32_t n, uint32_t op);
/*
 * mergeStates across GPUs.  SINGLE_STATE / COMPACT_KEY: one all-gather of the state's fixed-size
 * [states | packed keys] block and ONE kernel that folds all ranks' blocks in rank order, keyed by the packed
 * group key (the same group may sit in a different row on every rank) -- every rank ends with bit-identical
 * totals, no host synchronisation.  SEPARATE_CHAINING / COLLISION_FREE: the live groups of every rank are
 * all-gathered (padded to the largest rank) and the foreign ones upserted into the local table.
 * After the call every rank holds the merged state; finalize as usual.
 */
int qsgpu_agg_merge_all(qsgpu_agg_state_t state, qsgpu_comm_t comm);
/* Bitwise OR of a LIP filter's words over all ranks, in place (every GPU filled its copy from its share of the
 * build side; utility/lip_filter/BitVectorExactFilter.hpp:152-176 inserts into ONE shared filter). */
int qsgpu_lip_allreduce(qsgpu_lip_t lip, qsgpu_comm_t comm);
/* *out (created by the call) = the rows of `local` of rank 0, then rank 1, ...: the broadcast build side of a
 * hash join (every GPU builds its own copy of the table and probes its lineitem partition locally), or the
 * top-k candidates of every rank.  Row counts are exchanged first (one host synchronisation). */
int qsgpu_relation_allgather(qsgpu_relation_t local, qsgpu_comm_t comm, qsgpu_relation_t *out);
/* The same for a relation every rank contributes at most max_rows_per_rank rows to (the same value on every rank: the
 * LIMIT of a top-k).  When the ranks have a peer-memory mailbox (qsgpu_comm_peer_memory) and a rank's share fits a
 * slot, ONE kernel stores each rank's rows into the peers' memory, waits for theirs and assembles the result in rank
 * order; the row counts stay on the device, so the call only enqueues.  Otherwise it is qsgpu_relation_allgather. */
int qsgpu_relation_allgather_small(qsgpu_relation_t local, qsgpu_comm_t comm, uint64_t max_rows_per_rank,
                                   qsgpu_relation_t *out);

/* ---------------------------------------------------------- instrumentation */
/* CUDA-event time (ms) of the most recent kernel of the given family (launched by any thread of the process)
 * while timing is on, for bench.py's roofline block.  Timing makes every launch wait for its kernel. */
enum { QS_K_SCAN_AGG = 0, QS_K_SELECT = 1, QS_K_LIP = 2, QS_K_JOIN_BUILD = 3,
       QS_K_JOIN_PROBE = 4, QS_K_GROUPBY = 5, QS_K_PARTITION = 6, QS_K_TOPK = 7,
       QS_K_STAGE = 8, QS_K_FAMILIES = 9 };
int qsgpu_set_timing(int enabled);
/* CUDA events on the library's stream of `dev`: device time of everything
 * queued between start and stop (bench.py's timed region). */
int qsgpu_timer_start(int dev);
int qsgpu_timer_stop(int dev, float *ms);
int qsgpu_last_kernel_ms(uint32_t family, float *ms);
/* The same, accumulated over every launch of the family (by any thread) since timing was last switched on:
 * the last launch, the longest one, their sum and their number.  Any output pointer may be NULL. */
int qsgpu_kernel_ms_stats(uint32_t family, float *last, float *max, float *sum, uint32_t *count);


/* ------------------------------------------------------------ query compiler */
/*
 * The scan kernels are compiled per query shape at first use (NVRTC, sm_100a)
 * and cached in memory and under QSGPU_JIT_CACHE (default: jitcache/ next to
 * the library).  qsgpu_jit_selfcheck compiles representative work order
 * `which` (0..QSGPU_JIT_SELFCHECK_CASES-1: Q6-style single-state aggregate,
 * Q1-style compact-key group-by, select with LIP probes, BuildLIPFilter, join
 * build, inner probe with residual, anti probe, hash group-by, dense group-by,
 * dense join build, dense inner probe, LEFT OUTER probe, Q6 / Q1 / a join probe over dictionary codes, and
 * aggregates / Select / an inner join probe over NULL-able attributes)
 * WITHOUT a device -- the "does every kernel family still compile" check of
 * build() and the CPU test suite.  The generated CUDA source and the NVRTC log
 * are copied into the optional buffers.
 */
#define QSGPU_JIT_SELFCHECK_CASES 19
int qsgpu_jit_selfcheck(uint32_t which, char *source_out, size_t source_bytes,
                        char *log_out, size_t log_bytes);
/* NVRTC compilations / disk-cache hits / in-memory hits since load. */
int qsgpu_jit_stats(uint64_t *compiled, uint64_t *disk_hits, uint64_t *mem_hits);

#ifdef __cplusplus
}
#endif
#endif  /* QSGPU_H_ */
