// Operator descriptors and kernel launchers shared between the .cu files and
// the C-ABI implementation (capi.cu).
#pragma once

#include "qs_agg.cuh"
#include "qs_common.cuh"

namespace qs {

// Output side of Select / HashJoin probe / BuildLIPFilter / BuildHash scans.
struct SinkDesc {
  uint32_t n_out;
  uint8_t out_width[kMaxOut];
  char *out[kMaxOut];               // output column base pointers
  uint64_t capacity;                // rows the output relation can hold
  unsigned long long *counter;      // device row counter of the output relation
  uint32_t n_lip_build;             // LIPFilterBuilder::insertValueAccessor targets
  LipDesc lip_build[kMaxLip];
  uint16_t lip_build_col[kMaxLip];  // staged column slot of the build attribute
  uint8_t lip_build_ltype[kMaxLip];
  uint32_t *error_flag;
  // LEFT OUTER join: per-output-row NULL mask (bit j = projected column j is NULL) and the mask of the
  // columns whose scalar reads the build side (NULL for probe rows without a match)
  unsigned long long *null_out;
  uint64_t null_bits;
};

struct __align__(16) JoinSlot {                   // 16 bytes, one vector load per probe step
  int64_t key;
  unsigned long long row;           // build row id; ~0 = empty
};

struct JoinDesc {
  JoinSlot *slots;                  // open addressing: {key, build row} slots
  // Dense ("collision-free vector") table: heads[key - min_key] is the most recently inserted build row
  // with that key (~0 = none), next[row] chains the earlier ones.  No hashing, no key compare, 8 bytes
  // per key of the declared range -- the join-side analogue of CollisionFreeVectorTable
  // (storage/CollisionFreeVectorTable.hpp:56), usable when exact min/max statistics bound the build key.
  unsigned long long *heads;
  unsigned long long *next;         // indexed by build row id; sized to the build relation's capacity
  int64_t min_key;
  uint32_t dense;                   // 0 = open addressing, 1 = dense
  uint64_t cap;                     // open addressing: power of two;  dense: max_key - min_key + 1
  unsigned long long *n_entries;
  uint16_t key_col;                 // staged column slot of the key
  uint8_t key_ltype;                // V_I32 / V_I64
  // Composite key of two INT attributes (HashTable::putValueAccessorCompositeKey, storage/HashTable.hpp:1469):
  // the pair is packed into one 64-bit key (first attribute in the low word), so equality of the packed key
  // is equality of both components.  key2_present = 0: single-attribute key.
  uint8_t key2_present;
  uint16_t key2_col;
  uint8_t join_type;                // QS_JOIN_*
  // Probe side with a NULL-able key: the staged slot of the probe relation's per-row NULL mask (0xffff = the key
  // cannot be NULL) and the key attributes' bits in it.  A row with a NULL key passes the predicate but does not
  // search: it matches nothing, so an anti join emits it and an outer join emits it NULL-padded
  // (HashTable::runOverKeysFromValueAccessor, storage/HashTable.hpp:1999-2003).
  uint16_t null_col;
  uint64_t key_null_bits;
  const unsigned long long *build_nulls;   // per-row NULL masks of the build relation (nullptr: it has no NULL-able attribute)
  uint32_t n_build_cols;
  ColDesc build_cols[kMaxCols];     // build relation columns for LEAF_BUILD / raw emits
  uint32_t *error_flag;
};

struct FinalizeDesc {
  uint32_t n_key_cols, key_words;
  uint8_t key_width[kMaxKeyCols];
  uint8_t key_off[kMaxKeyCols];
  char *key_out[kMaxKeyCols];
  uint32_t n_out;
  uint8_t function[kMaxOut];   // QS_AGG_*
  uint8_t word[kMaxOut];       // state word holding the value (0 = row count)
  uint8_t nn_word[kMaxOut];    // state word counting the aggregate's non-NULL arguments (0 = the row count)
  uint8_t out_vtype[kMaxOut];  // VType of the output column
  uint8_t word_is_f64[kMaxOut];
  char *out[kMaxOut];
  int keys_are_slots;          // collision free: key value == slot index
  unsigned long long *rows_out;   // device row counter of the output relation (set to n by the kernel)
  // Fixed-size states: the live group count stays on the device (no host wait before the launch); the kernel
  // finalizes min(n, *d_n_groups) rows.  nullptr: n is exact.
  const uint32_t *d_n_groups;
  // SQL NULL of an aggregate over zero rows (AggregationHandleSum.cpp:134-143): per-row mask of the output
  // relation, bit j = output column j is NULL; null_bits = the mask of a group whose row count is zero.
  unsigned long long *null_out;
  uint64_t null_bits;
};

// One (block, attribute) stripe of a batched staging call (qsgpu_stage_blocks).
struct __align__(16) StageSeg {
  char *dst;                   // native-width destination (column base + first row * vw)
  const char *src;             // stripe / first slot inside the device copy of the block image
  const char *dict;            // QS_ENC_DICT: sorted values
  uint64_t n_rows;
  uint64_t tile_begin;         // first tile (kStageTileRows rows each) of this segment in the batch
  uint32_t encoding;           // QS_ENC_*
  uint32_t cw, vw, stride, dict_entries;
  uint32_t aligned;            // src (and dict) are vw-aligned: whole-value loads are legal
  // Stripe of an attribute the relation holds as codes (qsgpu_relation_set_dictionary): `dict` is this block's
  // re-coding table (block code -> relation code, vw bytes each), filled on the device by k_build_recode from
  // the block's own sorted dictionary `bdict` (native values of bw bytes, inside the block image) and the
  // relation-wide dictionary `gdict`.  gdict == nullptr for every other stripe.
  const char *bdict;
  const char *gdict;
  uint32_t g_entries, qtype, bw, pad;
  // NULL values of the stripe (QS_NULL_*): a NULL row's value is stored as zero bytes and `null_bit` is set in
  // null_dst[row], the relation's per-row NULL mask
  const unsigned char *null_src;
  unsigned long long *null_dst;
  uint64_t null_bit;
  uint32_t null_kind, null_arg, null_stride, null_width;
};
constexpr uint32_t kStageTileRows = 4096;

#ifndef __CUDACC_RTC__
// k_agg.cu
size_t agg_smem_extra(int hot, int n_agg, bool grouped, uint32_t words, bool priv = false);
int agg_hot_groups(const AggDesc &A);
cudaError_t launch_fill_identity(uint64_t *states, uint64_t n_rows, const AggDesc &A, cudaStream_t st);
// one launch that brings a fresh SINGLE_STATE / COMPACT_KEY state to its initial contents: per-CTA partial rows and
// totals = identities, key directory empty, counters zero (ctl = the 3 x 256-byte counter block)
cudaError_t launch_agg_init(const AggDesc &A, uint64_t partial_sets, void *ctl, cudaStream_t st);
// rows [0, min(max_rows, *d_rows)) of every column, the row count, the device error word and (optionally) the
// per-row NULL masks packed into one buffer: what one device-to-host copy brings back (qsgpu_relation_read_rows)
cudaError_t launch_pack_rows(char *dst, const ColDesc *cols, uint32_t n_cols, uint64_t max_rows,
                             const unsigned long long *d_rows, const unsigned long long *d_nulls, bool with_nulls,
                             uint32_t *error_flag, cudaStream_t st);
cudaError_t launch_merge_partials(const AggDesc &A, uint32_t n_ctas, cudaStream_t st);
cudaError_t launch_merge_foreign_compact(const AggDesc &A, const uint64_t *f_states, const uint64_t *f_keys,
                                         uint32_t f_groups, cudaStream_t st);
// k_groupby.cu
cudaError_t launch_rehash(const AggDesc &from, const AggDesc &to, cudaStream_t st);
cudaError_t launch_merge_foreign_table(const AggDesc &A, const uint64_t *f_states, const uint64_t *f_keys,
                                       uint64_t f_groups, cudaStream_t st);
cudaError_t launch_collect_slots(const uint64_t *states, uint32_t words, uint64_t cap, uint64_t *out_idx,
                                 unsigned long long *counter, const uint64_t *exist_words, cudaStream_t st);
cudaError_t launch_gather_rows(const uint64_t *states, const uint64_t *keys, uint32_t words, uint32_t kw,
                               const uint64_t *idx, uint64_t n, uint64_t *o_states, uint64_t *o_keys,
                               int keys_are_slots, cudaStream_t st);
cudaError_t launch_finalize(const uint64_t *states, const uint64_t *keys, uint32_t words, const uint64_t *idx,
                            uint64_t n, const FinalizeDesc &F, cudaStream_t st);
// k_join.cu
cudaError_t launch_join_clear(const JoinDesc &J, cudaStream_t st);
cudaError_t launch_join_rehash(const JoinSlot *from, uint64_t from_cap, JoinSlot *to, uint64_t to_cap, cudaStream_t st);
cudaError_t launch_fill_u64(unsigned long long *p, uint64_t n, unsigned long long v, cudaStream_t st);
// k_misc.cu
cudaError_t launch_decode_dict(void *dst, const void *codes, const void *dict, uint64_t n, uint32_t code_width,
                               uint32_t value_width, uint32_t dict_entries, cudaStream_t st);
cudaError_t launch_decode_truncated(void *dst, const void *codes, uint64_t n, uint32_t code_width,
                                    uint32_t value_width, cudaStream_t st);
cudaError_t launch_decode_strided(void *dst, const void *slots, uint64_t n, uint32_t stride,
                                  uint32_t value_width, cudaStream_t st);
cudaError_t launch_zero_date_padding(void *col, uint64_t n, cudaStream_t st);
cudaError_t launch_zero_after_nul(void *col, uint64_t n, uint32_t w, cudaStream_t st);
cudaError_t launch_build_recode(const StageSeg *d_segs, uint32_t n_segs, uint32_t *error_flag, cudaStream_t st);
cudaError_t launch_decode_segments(const StageSeg *d_segs, uint32_t n_segs, uint64_t n_tiles, int sm_count,
                                   cudaStream_t st);
#endif  // !__CUDACC_RTC__

}  // namespace qs
