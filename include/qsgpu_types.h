/*
 * qsgpu_types.h -- constants shared by the C ABI (qsgpu.h), the host library
 * and the device code of libqsgpu (this header is also what the query compiler
 * hands to NVRTC, so it declares no functions).
 */
#ifndef QSGPU_TYPES_H_
#define QSGPU_TYPES_H_

/* ------------------------------------------------------------------ status */
typedef enum qsgpu_status {
  QSGPU_OK = 0,
  QSGPU_ERR_NO_DEVICE = 1,      /* no CUDA device / init not called           */
  QSGPU_ERR_CUDA = 2,           /* a CUDA runtime call failed                 */
  QSGPU_ERR_INVALID = 3,        /* bad argument / malformed expression tree   */
  QSGPU_ERR_UNSUPPORTED = 4,    /* valid in the reference, not lowered (yet)  */
  QSGPU_ERR_CAPACITY = 5,       /* table / output relation capacity exceeded  */
  QSGPU_ERR_OOM = 6
} qsgpu_status;

/* ------------------------------------------------------------------- types */
/* Values follow types/TypeID.hpp:33-45. */
enum {
  QS_INT = 0,      /* int32                                                    */
  QS_LONG = 1,     /* int64                                                    */
  QS_FLOAT = 2,    /* float                                                    */
  QS_DOUBLE = 3,   /* double (SQL DECIMAL parses to this, SqlParser.ypp:791)   */
  QS_CHAR = 4,     /* fixed width, NUL padded, strncmp order                   */
  QS_VARCHAR = 5,  /* never staged on device (QSGPU_ERR_UNSUPPORTED)           */
  QS_DATE = 6      /* DateLit {int32 year; u8 month; u8 day; 2 pad} = 8 bytes, */
                   /* lexicographic order (types/DatetimeLit.hpp:38-93)        */
};

/* Comparison ids (types/operations/comparisons/ComparisonID.hpp). */
enum { QS_EQ = 0, QS_NE = 1, QS_LT = 2, QS_LE = 3, QS_GT = 4, QS_GE = 5 };

/* Binary operation ids (binary_operations/BinaryOperationID.hpp). */
enum { QS_ADD = 0, QS_SUB = 1, QS_MUL = 2, QS_DIV = 3, QS_MOD = 4 };

/* Unary operation ids used on the path. */
enum { QS_NEGATE = 0, QS_CAST = 1 };

/* Aggregate function ids (expressions/aggregation/AggregationID.hpp). */
enum { QS_AGG_AVG = 0, QS_AGG_COUNT = 1, QS_AGG_MAX = 2, QS_AGG_MIN = 3, QS_AGG_SUM = 4 };

/* LIP filter kinds (utility/lip_filter/LIPFilter.proto:24-62). */
enum { QS_LIP_BITVECTOR_EXACT = 0, QS_LIP_SINGLE_IDENTITY_HASH = 1 };

/* Aggregation strategies (see qsgpu_agg_create in qsgpu.h). */
enum {
  QS_AGG_SINGLE_STATE = 0,
  QS_AGG_COMPACT_KEY = 1,
  QS_AGG_SEPARATE_CHAINING = 2,
  QS_AGG_COLLISION_FREE = 3
};

/* Physical encodings of a staged block stripe (see qsgpu_stage_block in qsgpu.h). */
enum { QS_ENC_PLAIN = 0, QS_ENC_STRIDED = 1, QS_ENC_DICT = 2, QS_ENC_TRUNCATED = 3,
       QS_ENC_SKIP = 4 /* batched staging only: leave this attribute on the host (column pruning) */ };

/* How a staged stripe marks NULL values (qs_stage_desc.null_kind):
 *   QS_NULL_CODE       dictionary-compressed stripe: the code null_arg means NULL (CompressionDictionaryLite keeps it
 *                      behind the number of codes, compression/CompressionDictionaryLite.hpp:40-51; the builder
 *                      assigns it in CompressionDictionaryBuilder::buildDictionary)
 *   QS_NULL_BITMAP     a BitVector<false> (64-bit words, most significant bit first, utility/BitVector.hpp:934): bit
 *                      null_arg + row * null_stride.  Column stores keep one bitmap per NULL-able uncompressed
 *                      attribute (stride 1; storage/BasicColumnStoreTupleStorageSubBlock.hpp:230,
 *                      storage/CompressedColumnStoreTupleStorageSubBlock.cpp:252-312)
 *   QS_NULL_SLOT_WORD  SplitRowStore: every tuple slot starts with a BitVector<true> over the relation's NULL-able
 *                      attributes (a 1/2/4-byte word up to 32 of them, 64-bit words beyond;
 *                      storage/SplitRowStoreTupleStorageSubBlock.cpp:130,348): bit null_arg, counted from the most
 *                      significant bit, of the null_width-byte word at null_bitmap + row * null_stride           */
enum { QS_NULL_NONE = 0, QS_NULL_CODE = 1, QS_NULL_BITMAP = 2, QS_NULL_SLOT_WORD = 3 };

/* Join types (relational_operators/HashJoinOperator.hpp:82-87). */
enum { QS_JOIN_INNER = 0, QS_JOIN_LEFT_SEMI = 1, QS_JOIN_LEFT_ANTI = 2, QS_JOIN_LEFT_OUTER = 3 };

#endif  /* QSGPU_TYPES_H_ */
