// Host storage blocks in the reference's three physical layouts, and their
// device residency (kernel family K0).
//
// Stands in for storage/StorageManager.hpp + the TupleStorageSubBlock formats
// so the GPU operators can be driven and tested without the reference's buffer
// pool.  What it keeps of the reference:
//   * a relation is a list of blocks (CatalogRelation::getBlocksSnapshot);
//   * a block is one contiguous piece of memory whose stripes are found from
//     the sub-block header:
//       BasicColumnStore       stripes back to back, stripe i = n x width_i
//                              (storage/BasicColumnStoreTupleStorageSubBlock.cpp:100-183)
//       CompressedColumnStore  [dictionaries][stripes]; per attribute and per
//                              block the builder keeps the smaller of truncation
//                              (non-negative INT/LONG to 1/2/4 bytes), an ordered
//                              dictionary with 1/2/4-byte codes, or the native
//                              stripe (storage/CompressedBlockBuilder.cpp:434-506,
//                              compression/CompressionDictionaryLite.hpp:40-51)
//       SplitRowStore          fixed-width attributes back to back inside
//                              tuple slots (storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179)
//   * getBlock()-style access by block id.
// What it adds: deviceRelation(), the HBM image of a relation's blocks, staged
// by ONE qsgpu_stage_blocks call per batch of blocks and cached (the analogue
// of the warm buffer pool); and device-only temporary relations.
#pragma once

#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "QsTypes.hpp"

namespace quickstep {

enum class TupleStoreLayout { kBasicColumnStore = 0, kCompressedColumnStore = 1, kSplitRowStore = 2 };

struct StorageBlock {
  block_id id = 0;
  relation_id relation = -1;
  tuple_id num_tuples = 0;
  const char *memory = nullptr;          // block image (inside the relation's slab)
  std::size_t size = 0;                  // bytes of the image
  std::vector<qs_stage_desc> stripes;    // one per attribute, pointers into `memory`
};

// Row range of a device relation: what a GPU work order scans.
struct DeviceExtent {
  qsgpu_relation_t relation = nullptr;
  std::uint64_t row_begin = 0, row_end = UINT64_MAX;
};

constexpr std::size_t kSlotSizeBytes = 0x200000;     // storage/StorageConstants.hpp:50

// Read side of a SplitRowStore block of fixed-length, non-nullable attributes (what
// SplitRowStoreTupleStorageSubBlock::getAttributeValue does): where tuple t's attribute a lives.
class SplitRowStoreReader {
 public:
  SplitRowStoreReader(const char *block_memory, const std::vector<qs_attr> &schema);
  std::uint64_t numTuples() const { return num_tuples_; }
  const char *value(std::uint64_t tuple, std::size_t attr) const { return slots_ + tuple * slot_bytes_ + offsets_[attr]; }

 private:
  std::uint64_t num_tuples_ = 0;
  const char *slots_ = nullptr;
  std::size_t slot_bytes_ = 0;
  std::vector<std::size_t> offsets_;
};

class StorageManager {
 public:
  // pinned_blocks = false keeps the block images in ordinary host memory: what the device-free unit tests of
  // the block builder use (staging from unpinned memory works too, only slower).
  explicit StorageManager(int device = 0, bool pinned_blocks = true) : device_(device), pinned_(pinned_blocks) {}
  ~StorageManager();
  int device() const { return device_; }

  // The loader's job in the reference (TextScan -> InsertDestination -> blocks):
  // cut `n_rows` tuples given as native-width columns into blocks of
  // `rows_per_block` tuples in the requested layout and register them with `rel`.
  // null_flags (optional): one entry per attribute, nullptr or one byte per row (non-zero = the value is NULL) for
  // the attributes the catalog declares NULL-able.  Each layout records NULLs the way the reference does: a per-tuple
  // BitVector<true> at the head of a SplitRowStore slot, one BitVector<false> per NULL-able attribute in a column
  // store, the dictionary's null code (= number of codes) in a dictionary-compressed stripe.
  void loadRelation(CatalogRelation *rel, const std::vector<const void *> &columns, std::uint64_t n_rows,
                    std::uint64_t rows_per_block, TupleStoreLayout layout,
                    const std::vector<const std::uint8_t *> &null_flags = {});
  const StorageBlock &getBlock(block_id id) const;
  std::uint64_t hostBytes(const CatalogRelation &rel) const;      // bytes of all block images

  // HBM image of the blocks of a stored relation, in block-list order.  The device cache is keyed by
  // (block, attribute): only the attributes in `needed_attrs` (bit a = attribute a; default all) are
  // guaranteed resident on return; attributes staged by earlier calls stay, missing ones are staged now
  // (qsgpu_stage_blocks / qsgpu_stage_columns with QS_ENC_SKIP for the rest).
  qsgpu_relation_t deviceRelation(const CatalogRelation &rel, std::uint64_t needed_attrs = ~0ull);
  DeviceExtent blockExtent(block_id id);                           // rows of one block inside it
  // deviceRelation() + the row ranges a scan of the whole stored relation is cut into: runs of adjacent blocks of at
  // most `max_rows` rows (0 = one run).  Cached with the image: a query over 9,525 resident blocks (lineitem at
  // SF100) costs one map lookup here, not one per block.
  std::vector<DeviceExtent> stagedExtents(const CatalogRelation &rel, std::uint64_t needed_attrs, std::uint64_t max_rows);
  void evict(const CatalogRelation &rel);                          // drop the HBM image

  // Keep dictionary-compressed attributes as CODES in HBM: an attribute whose stripe
  // is dictionary-compressed in every block and whose block dictionaries union to <= 65,536 values narrower than
  // their codes is declared with qsgpu_relation_set_dictionary, and qsgpu_stage_blocks re-codes the blocks into that
  // relation-wide dictionary instead of decoding them.  Takes effect the next time a relation's image is built.
  void setCodeResident(bool on) { code_resident_ = on; }
  bool codeResident() const { return code_resident_; }
  // (code width, dictionary entries) the image of `rel` uses for attribute `attr`; (0, 0) = native
  std::pair<std::uint32_t, std::uint32_t> residentCoding(const CatalogRelation &rel, std::uint32_t attr);

  // ---- several devices (one process per GPU; partition id <-> device id) ----------
  // The reference keeps one block list per partition of a relation (catalog/PartitionScheme.hpp) and one
  // state / hash table per partition (query_execution/QueryContext.cpp:66-97).  Here every process holds ONE
  // partition of each partitioned relation; a relation is either PARTITIONED across the devices (base
  // relations loaded while a communicator is set, and whatever is derived from them row by row) or
  // REPLICATED (the same rows on every device: merged aggregation results, all-gathered build sides).
  void setCommunicator(qsgpu_comm_t comm) { comm_ = comm; }
  qsgpu_comm_t communicator() const { return comm_; }
  bool multiDevice() const { int r = 0, n = 1; qsgpu_comm_rank(comm_, &r, &n); return n > 1; }
  void setPartitioned(relation_id id, bool partitioned);
  bool isPartitioned(relation_id id) const;
  // All-gathered copy of a partitioned temporary relation (the broadcast build side of a hash join); built on
  // first use, owned by the manager, dropped together with the temporary.
  qsgpu_relation_t replicated(const CatalogRelation &rel);

  // ---- InsertDestination::bulkInsertTuples (storage/InsertDestination.cpp:202-216), the GPU -> host hand-off -------
  // Rows that leave the device (the result relation of a query, or the input of an operator that stays on the CPU)
  // arrive as host column vectors and are written into blocks of the output relation in the layout the reference
  // gives temporary relations: SplitRowStore inside a one-slot (2 MB) StorageBlock,
  //   [int header_len][StorageBlockHeader proto][Header{num_tuples, max_tid, var_bytes, compact}][occupancy bitmap][slots]
  // (storage/StorageBlockLayout.proto:103-124, storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179, .hpp:356-361),
  // bit-compatible with what the reference's own accessors read.  Block buffers come from a small pool and are
  // recycled by dropTemporary().  Returns the ids of the blocks filled, in order.
  std::vector<block_id> insertTuples(const CatalogRelation &rel, const std::vector<const void *> &columns, std::uint64_t n_rows);
  std::vector<block_id> hostBlocksOf(const CatalogRelation &rel) const;

  // Device-only temporary relation (output of Select / HashJoin / Finalize);
  // its single pseudo block id stands for "every row produced so far".
  block_id createTemporary(const CatalogRelation &rel, std::uint64_t capacity_rows);
  void adoptTemporary(const CatalogRelation &rel, qsgpu_relation_t handle);   // takes ownership
  qsgpu_relation_t temporary(const CatalogRelation &rel);
  // PartitionAwareInsertDestination: regroup the rows of a temporary relation by HashPartitionSchemeHeader's
  // partition function on `partition_attribute` (K8, qsgpu_hash_partition) and hand back one pseudo block per
  // partition; blockExtent() of block p is the row range of partition p.
  std::vector<block_id> repartitionTemporary(const CatalogRelation &rel, attribute_id partition_attribute, std::size_t num_partitions);
  void dropTemporary(const CatalogRelation &rel);

 private:
  struct Slab { char *base = nullptr; std::size_t bytes = 0; };
  struct Resident {
    qsgpu_relation_t handle = nullptr;
    std::size_t n_blocks_staged = 0;
    std::uint64_t rows = 0;
    std::uint64_t staged_attrs = 0;      // bit a: attribute a of every staged block is in HBM
    std::vector<DeviceExtent> runs;      // stagedExtents() cache, valid for runs_max_rows while the image stands
    std::uint64_t runs_max_rows = ~0ull;
  };
  // union of the block dictionaries of one attribute, sorted in the attribute's order (cached per relation and
  // block count: blocks are immutable once built)
  struct RelationDictionary { std::uint32_t code_width = 0, n_entries = 0; std::vector<char> values; };
  const std::vector<RelationDictionary> &relationDictionaries(const CatalogRelation &rel, const std::vector<block_id> &ids);
  std::map<std::pair<relation_id, std::size_t>, std::vector<RelationDictionary>> dictionaries_;
  bool code_resident_ = false;
  int device_;
  bool pinned_;
  mutable std::mutex mu_;
  block_id next_block_ = 1;
  std::unordered_map<block_id, StorageBlock> blocks_;
  std::unordered_map<block_id, std::pair<relation_id, std::uint64_t>> block_first_row_;
  std::map<relation_id, std::vector<Slab>> slabs_;
  std::map<relation_id, Resident> resident_;
  std::map<relation_id, qsgpu_relation_t> temporaries_;
  std::map<relation_id, qsgpu_relation_t> replicas_;
  std::map<relation_id, std::vector<block_id>> result_blocks_;     // host blocks written by insertTuples()
  std::vector<char *> block_pool_;                                  // recycled 2 MB block buffers
  std::map<relation_id, bool> partitioned_;
  qsgpu_comm_t comm_ = nullptr;
  std::map<relation_id, block_id> temporary_block_;
  std::map<block_id, DeviceExtent> partition_extents_;                 // pseudo blocks of repartitioned temporaries
  std::map<relation_id, std::vector<block_id>> partition_blocks_;
};

}  // namespace quickstep
