#include "StorageManager.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>

namespace quickstep {

namespace {

std::size_t pad16(std::size_t n) { return (n + 15) & ~static_cast<std::size_t>(15); }

// What the block builder decided for one attribute of one block.
struct StripePlan {
  std::uint32_t encoding = QS_ENC_PLAIN;
  std::uint32_t code_width = 0;
  std::vector<char> dict;            // sorted distinct values, native width
  std::size_t stripe_bytes = 0;
  bool has_nulls = false;            // some value of the block is NULL: null code (dictionary) or a bitmap (native stripe)
};

// CompressedBlockBuilder's rule (storage/CompressedBlockBuilder.cpp:470-498): keep the
// smaller of {truncated (or native) stripe} and {dictionary + code stripe}.
// Value order first, bit pattern as the tie-break (-0.0 and 0.0 are distinct dictionary entries,
// exactly one code per stored bit pattern, so decode(encode(x)) is bit-identical to x).
// A strict weak order also in the presence of NaNs (they go last, ordered by bit pattern among themselves): std::sort
// and std::lower_bound need one, and `a < b` alone is not (NaN is unordered with everything).
template <class T>
bool valueLess(const T &a, const T &b) {
  if constexpr (std::is_floating_point<T>::value) {
    const bool an = a != a, bn = b != b;
    if (an || bn) return (an && bn) ? std::memcmp(&a, &b, sizeof(T)) < 0 : bn;
  }
  if (a < b) return true;
  if (b < a) return false;
  return std::memcmp(&a, &b, sizeof(T)) < 0;
}

template <class T>
StripePlan planCompressed(const T *v_all, std::uint64_t n_all, bool is_integer, const std::uint8_t *nulls) {
  StripePlan p;
  const std::size_t w = sizeof(T);
  std::size_t truncated = w;
  // NULL rows take no part in the dictionary (their code is the number of codes) and rule truncation out
  // (storage/CompressedBlockBuilder.cpp:357-381: "if the new value is null or less than zero, we can't truncate it")
  std::vector<T> non_null;
  bool any_null = false;
  if (nulls) {
    for (std::uint64_t i = 0; i < n_all; ++i) { if (nulls[i]) any_null = true; else non_null.push_back(v_all[i]); }
  }
  const T *v = any_null ? non_null.data() : v_all;
  const std::uint64_t n = any_null ? non_null.size() : n_all;
  p.has_nulls = any_null;
  if constexpr (std::is_integral<T>::value) {
    if (is_integer && n > 0 && !any_null) {
      bool nonneg = true;
      std::uint64_t mx = 0;
      for (std::uint64_t i = 0; i < n; ++i) {
        if (v[i] < 0) { nonneg = false; break; }
        mx = std::max<std::uint64_t>(mx, static_cast<std::uint64_t>(v[i]));
      }
      if (nonneg) {
        if (mx < (1ull << 8)) truncated = 1;
        else if (mx < (1ull << 16)) truncated = 2;
        else if (mx < (1ull << 32) && w > 4) truncated = 4;
      }
    }
  }
  std::vector<T> d(v, v + n);
  std::sort(d.begin(), d.end(), valueLess<T>);
  d.erase(std::unique(d.begin(), d.end(), [](const T &a, const T &b) { return std::memcmp(&a, &b, sizeof(T)) == 0; }), d.end());
  const std::size_t codes = d.size() + 1;                    // one code is reserved for NULL
  const std::size_t cw = codes <= (1u << 8) ? 1 : codes <= (1u << 16) ? 2 : 4;
  const std::size_t dict_bytes = 8 + d.size() * w + n_all * cw;   // [u32 num_codes][u32 null_code][values] + codes
  if (truncated * n_all < dict_bytes) {
    if (truncated < w) { p.encoding = QS_ENC_TRUNCATED; p.code_width = static_cast<std::uint32_t>(truncated); p.stripe_bytes = n_all * truncated; }
    else { p.encoding = QS_ENC_PLAIN; p.stripe_bytes = n_all * w; }
  } else {
    p.encoding = QS_ENC_DICT;
    p.code_width = static_cast<std::uint32_t>(cw);
    p.stripe_bytes = n_all * cw;
    p.dict.resize(d.size() * w);
    std::memcpy(p.dict.data(), d.data(), d.size() * w);
  }
  return p;
}

template <class T>
void writeCodes(char *dst, const T *v, std::uint64_t n, const StripePlan &p, const std::uint8_t *nulls) {
  if (p.encoding == QS_ENC_TRUNCATED) {
    if constexpr (std::is_integral<T>::value) {
      for (std::uint64_t i = 0; i < n; ++i) {
        const std::uint64_t x = static_cast<std::uint64_t>(v[i]);
        std::memcpy(dst + i * p.code_width, &x, p.code_width);          // little endian
      }
    }
    return;
  }
  const T *d = reinterpret_cast<const T *>(p.dict.data());
  const std::size_t nd = p.dict.size() / sizeof(T);
  for (std::uint64_t i = 0; i < n; ++i) {
    // ordered dictionary: code = rank of the value (CompressionDictionaryLite: code order = value order)
    // a NULL is the code one past the dictionary (CompressionDictionaryBuilder assigns null_code = number of codes)
    const std::uint32_t c = (nulls && nulls[i]) ? static_cast<std::uint32_t>(nd)
                                                : static_cast<std::uint32_t>(std::lower_bound(d, d + nd, v[i], valueLess<T>) - d);
    std::memcpy(dst + i * p.code_width, &c, p.code_width);
  }
}

struct DateKey {           // DateLit ordered as (year, month, day): types/DatetimeLit.hpp:65-93
  std::uint64_t raw;
  std::int64_t key() const {
    return static_cast<std::int64_t>(static_cast<std::int32_t>(raw & 0xffffffffu)) * 65536 +
           static_cast<std::int64_t>(((raw >> 32) & 0xff) << 8 | ((raw >> 40) & 0xff));
  }
  bool operator<(const DateKey &o) const { return key() < o.key(); }
};

StripePlan planAttr(const qs_attr &a, const char *col, std::uint64_t n, const std::uint8_t *nulls) {
  switch (a.type) {
    case QS_INT: return planCompressed(reinterpret_cast<const std::int32_t *>(col), n, true, nulls);
    case QS_LONG: return planCompressed(reinterpret_cast<const std::int64_t *>(col), n, true, nulls);
    case QS_FLOAT: return planCompressed(reinterpret_cast<const float *>(col), n, false, nulls);
    case QS_DOUBLE: return planCompressed(reinterpret_cast<const double *>(col), n, false, nulls);
    case QS_DATE: return planCompressed(reinterpret_cast<const DateKey *>(col), n, false, nulls);
    default: {            // CHAR(n): kept native here (the reference would dictionary-encode long strings)
      StripePlan p; p.encoding = QS_ENC_PLAIN; p.stripe_bytes = n * a.width;
      for (std::uint64_t i = 0; nulls && i < n; ++i) p.has_nulls = p.has_nulls || nulls[i] != 0;
      return p;
    }
  }
}

void writeAttr(const qs_attr &a, char *dst, const char *col, std::uint64_t n, const StripePlan &p, const std::uint8_t *nulls) {
  if (p.encoding == QS_ENC_PLAIN) { std::memcpy(dst, col, n * a.width); return; }
  switch (a.type) {
    case QS_INT: writeCodes(dst, reinterpret_cast<const std::int32_t *>(col), n, p, nulls); break;
    case QS_LONG: writeCodes(dst, reinterpret_cast<const std::int64_t *>(col), n, p, nulls); break;
    case QS_FLOAT: writeCodes(dst, reinterpret_cast<const float *>(col), n, p, nulls); break;
    case QS_DOUBLE: writeCodes(dst, reinterpret_cast<const double *>(col), n, p, nulls); break;
    case QS_DATE: writeCodes(dst, reinterpret_cast<const DateKey *>(col), n, p, nulls); break;
    default: QS_CHECK(false);
  }
}

// BitVector<false>::BytesNeeded (utility/BitVector.hpp:107-125) and BitVector<true> for the per-tuple bitmap
std::size_t bitVectorBytes(std::size_t bits) { return ((bits + 63) / 64) * 8; }
std::size_t shortBitVectorBytes(std::size_t bits) {
  return bits == 0 ? 0 : bits < 9 ? 1 : bits < 17 ? 2 : bits < 33 ? 4 : bitVectorBytes(bits);
}
// bit i of a most-significant-bit-first bit string held in little-endian words of `word_bytes` bytes
void setMsbFirstBit(char *words, std::size_t word_bytes, std::size_t i) {
  const std::size_t word = i / (8 * word_bytes), k = i % (8 * word_bytes);
  unsigned char *b = reinterpret_cast<unsigned char *>(words) + word * word_bytes + (word_bytes - 1 - (k >> 3));
  *b = static_cast<unsigned char>(*b | (0x80u >> (k & 7)));
}

}  // namespace

namespace {

// SplitRowStore geometry of a fixed-length, non-nullable schema inside a one-slot block
struct SplitRowStoreGeometry {
  std::size_t header_len = 17;                       // serialized StorageBlockHeader (see writeBlockHeader)
  std::size_t sub_block = 4 + 17;                    // offset of the tuple-store sub-block
  std::size_t sub_block_bytes = 0, slot_bytes = 0, max_tuples = 0, occupancy_bytes = 0, slots = 0;
  std::vector<std::size_t> offsets;
  explicit SplitRowStoreGeometry(const std::vector<qs_attr> &schema) {
    for (const qs_attr &a : schema) { offsets.push_back(slot_bytes); slot_bytes += a.width; }
    if (slot_bytes == 0) slot_bytes = 1;
    sub_block_bytes = kSlotSizeBytes - sub_block;
    max_tuples = (sub_block_bytes - 16) / slot_bytes;                    // sizeof(Header) = 16
    occupancy_bytes = ((max_tuples + 63) / 64) * 8;                      // BitVector<false>::BytesNeeded
    slots = sub_block + 16 + occupancy_bytes;
    max_tuples = std::min(max_tuples, (kSlotSizeBytes - slots) / slot_bytes);
  }
};

// StorageBlockHeader{layout{num_slots = 1, tuple_store_description{sub_block_type = SPLIT_ROW_STORE}},
// tuple_store_size (fixed64)} in protobuf wire format, preceded by its length
void writeBlockHeader(char *block, std::uint64_t tuple_store_size) {
  const unsigned char hdr[9] = {0x0A, 0x06, 0x08, 0x01, 0x12, 0x02, 0x08, 0x03, 0x11};
  const std::int32_t len = 17;
  std::memcpy(block, &len, 4);
  std::memcpy(block + 4, hdr, 9);
  std::memcpy(block + 13, &tuple_store_size, 8);
}

}  // namespace

SplitRowStoreReader::SplitRowStoreReader(const char *block_memory, const std::vector<qs_attr> &schema) {
  const SplitRowStoreGeometry g(schema);
  std::int32_t n = 0;
  std::memcpy(&n, block_memory + g.sub_block, 4);
  num_tuples_ = static_cast<std::uint64_t>(n);
  slots_ = block_memory + g.slots;
  slot_bytes_ = g.slot_bytes;
  offsets_ = g.offsets;
}

std::vector<block_id> StorageManager::insertTuples(const CatalogRelation &rel, const std::vector<const void *> &columns,
                                                   std::uint64_t n_rows) {
  const std::vector<qs_attr> schema = rel.schema();
  QS_CHECK(columns.size() == schema.size());
  const SplitRowStoreGeometry g(schema);
  std::vector<block_id> out;
  std::uint64_t row = 0;
  do {
    const std::uint64_t n = std::min<std::uint64_t>(n_rows - row, g.max_tuples);
    char *mem = nullptr;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!block_pool_.empty()) { mem = block_pool_.back(); block_pool_.pop_back(); }
    }
    if (!mem) { mem = static_cast<char *>(std::calloc(1, kSlotSizeBytes)); QS_CHECK(mem != nullptr); }
    // a recycled buffer may hold an older block: clear what this one will use (headers, occupancy bits, its slots)
    std::memset(mem, 0, g.slots + n * g.slot_bytes);
    writeBlockHeader(mem, g.sub_block_bytes);
    const std::int32_t num = static_cast<std::int32_t>(n), max_tid = num - 1;
    const std::uint32_t var_bytes = 0;
    std::memcpy(mem + g.sub_block, &num, 4);
    std::memcpy(mem + g.sub_block + 4, &max_tid, 4);
    std::memcpy(mem + g.sub_block + 8, &var_bytes, 4);
    mem[g.sub_block + 12] = 1;                                            // variable_length_storage_compact
    std::uint64_t *occ = reinterpret_cast<std::uint64_t *>(mem + g.sub_block + 16);   // unaligned by design (memcpy below)
    for (std::uint64_t w = 0; w * 64 < n; ++w) {
      const std::uint64_t bits = std::min<std::uint64_t>(64, n - w * 64);
      const std::uint64_t word = bits == 64 ? ~0ull : (~0ull << (64 - bits));      // MSB-first (utility/BitVector.hpp:934)
      std::memcpy(reinterpret_cast<char *>(occ) + w * 8, &word, 8);
    }
    for (std::size_t a = 0; a < schema.size(); ++a) {
      const std::size_t w = schema[a].width;
      const char *src = static_cast<const char *>(columns[a]) + row * w;
      char *dst = mem + g.slots + g.offsets[a];
      for (std::uint64_t i = 0; i < n; ++i) std::memcpy(dst + i * g.slot_bytes, src + i * w, w);
    }
    StorageBlock B;
    B.relation = rel.getID();
    B.num_tuples = static_cast<tuple_id>(n);
    B.memory = mem;
    B.size = kSlotSizeBytes;
    {
      std::lock_guard<std::mutex> lk(mu_);
      B.id = next_block_++;
      out.push_back(B.id);
      result_blocks_[rel.getID()].push_back(B.id);
      blocks_.emplace(B.id, std::move(B));
    }
    row += n;
  } while (row < n_rows);
  return out;
}

std::vector<block_id> StorageManager::hostBlocksOf(const CatalogRelation &rel) const {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = result_blocks_.find(rel.getID());
  return it == result_blocks_.end() ? std::vector<block_id>() : it->second;
}

StorageManager::~StorageManager() {
  for (auto &kv : result_blocks_) for (block_id b : kv.second) std::free(const_cast<char *>(blocks_.at(b).memory));
  for (char *p : block_pool_) std::free(p);
  for (auto &kv : replicas_) if (kv.second) qsgpu_relation_destroy(kv.second);
  for (auto &kv : resident_) if (kv.second.handle) qsgpu_relation_destroy(kv.second.handle);
  for (auto &kv : temporaries_) if (kv.second) qsgpu_relation_destroy(kv.second);
  for (auto &kv : slabs_) for (Slab &s : kv.second) if (s.base) { if (pinned_) qsgpu_host_free(s.base); else std::free(s.base); }
}

void StorageManager::loadRelation(CatalogRelation *rel, const std::vector<const void *> &columns, std::uint64_t n_rows,
                                  std::uint64_t rows_per_block, TupleStoreLayout layout,
                                  const std::vector<const std::uint8_t *> &null_flags) {
  const std::vector<qs_attr> schema = rel->schema();
  QS_CHECK(columns.size() == schema.size());
  QS_CHECK(null_flags.empty() || null_flags.size() == schema.size());
  if (rows_per_block == 0) rows_per_block = std::max<std::uint64_t>(n_rows, 1);
  const std::uint64_t n_blocks = (n_rows + rows_per_block - 1) / rows_per_block;
  // NULL-able attributes in catalog order: attribute a is bit nullable_index[a] of a tuple's NULL bitmap
  const std::uint64_t nullable = rel->nullableMask();
  std::vector<int> nullable_index(schema.size(), -1);
  std::size_t n_nullable = 0;
  for (std::size_t a = 0; a < schema.size(); ++a) {
    if ((nullable >> a) & 1) nullable_index[a] = static_cast<int>(n_nullable++);
    else QS_CHECK(null_flags.empty() || null_flags[a] == nullptr);       // NULLs only where the catalog allows them
  }
  auto flags_of = [&](std::size_t a, std::uint64_t r0) -> const std::uint8_t * {
    return (null_flags.empty() || !null_flags[a]) ? nullptr : null_flags[a] + r0;
  };
  const std::size_t tuple_null_bytes = shortBitVectorBytes(n_nullable);
  const std::size_t tuple_null_word = n_nullable > 32 ? 8 : tuple_null_bytes;
  std::size_t slot_bytes = tuple_null_bytes;
  for (const qs_attr &a : schema) slot_bytes += a.width;

  // ---- pass 1 (parallel): decide the physical form of every stripe, size the images
  std::vector<std::vector<StripePlan>> plans(n_blocks, std::vector<StripePlan>(schema.size()));
  std::vector<std::size_t> image_bytes(n_blocks, 0);
  const unsigned n_threads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32u));
  auto for_blocks = [&](auto &&fn) {
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < n_threads; ++t)
      ts.emplace_back([&, t] { for (std::uint64_t b = t; b < n_blocks; b += n_threads) fn(b); });
    for (auto &t : ts) t.join();
  };
  for_blocks([&](std::uint64_t b) {
    const std::uint64_t r0 = b * rows_per_block, n = std::min(rows_per_block, n_rows - r0);
    std::size_t bytes = 0;
    if (layout == TupleStoreLayout::kSplitRowStore) {
      bytes = n * slot_bytes + 8;        // (+8: the last slot's NULL word may be read as a whole 8-byte word)
    } else {
      for (std::size_t a = 0; a < schema.size(); ++a) {
        const char *col = static_cast<const char *>(columns[a]) + r0 * schema[a].width;
        StripePlan &p = plans[b][a];
        if (layout == TupleStoreLayout::kCompressedColumnStore) {
          p = planAttr(schema[a], col, n, flags_of(a, r0));
          // an uncompressed stripe with NULLs carries a bitmap (CompressedBlockBuilder.cpp:439-452)
          if (p.has_nulls && p.encoding != QS_ENC_DICT) bytes += bitVectorBytes(n);
        } else {
          p.encoding = QS_ENC_PLAIN; p.stripe_bytes = n * schema[a].width;
          // BasicColumnStore: one bitmap per NULL-able attribute, NULLs or not (…SubBlock.cpp:152-166)
          if (nullable_index[a] >= 0) bytes += bitVectorBytes(n);
        }
        bytes += p.dict.size() + p.stripe_bytes;
      }
    }
    image_bytes[b] = pad16(bytes);
  });

  // ---- one pinned slab for the relation's blocks (contiguous images => one H2D copy per batch)
  std::size_t total = 0;
  std::vector<std::size_t> off(n_blocks);
  for (std::uint64_t b = 0; b < n_blocks; ++b) { off[b] = total; total += image_bytes[b]; }
  Slab slab;
  slab.bytes = total;
  void *hp = nullptr;
  if (pinned_) QS_CHECK_GPU(qsgpu_host_alloc(std::max<std::size_t>(total, 16), &hp));
  else hp = std::aligned_alloc(4096, (std::max<std::size_t>(total, 16) + 4095) & ~static_cast<std::size_t>(4095));
  QS_CHECK(hp != nullptr);
  slab.base = static_cast<char *>(hp);

  // ---- pass 2 (parallel): write the images
  std::vector<StorageBlock> built(n_blocks);
  for_blocks([&](std::uint64_t b) {
    const std::uint64_t r0 = b * rows_per_block, n = std::min(rows_per_block, n_rows - r0);
    StorageBlock &B = built[b];
    B.relation = rel->getID();
    B.num_tuples = static_cast<tuple_id>(n);
    B.memory = slab.base + off[b];
    B.size = image_bytes[b];
    B.stripes.resize(schema.size());
    char *w = slab.base + off[b];
    std::memset(w, 0, image_bytes[b]);
    if (layout == TupleStoreLayout::kSplitRowStore) {
      // slot = [BitVector<true> over the NULL-able attributes][fixed-length attributes back to back]
      // (storage/SplitRowStoreTupleStorageSubBlock.cpp:130,348)
      std::size_t attr_off = tuple_null_bytes;
      for (std::size_t a = 0; a < schema.size(); ++a) {
        const std::uint32_t vw = schema[a].width;
        const char *col = static_cast<const char *>(columns[a]) + r0 * vw;
        for (std::uint64_t i = 0; i < n; ++i) std::memcpy(w + i * slot_bytes + attr_off, col + i * vw, vw);
        qs_stage_desc &s = B.stripes[a];
        s = qs_stage_desc{};
        s.attr = static_cast<std::uint32_t>(a); s.encoding = QS_ENC_STRIDED; s.host = w + attr_off;
        s.stride = static_cast<std::uint32_t>(slot_bytes);
        if (nullable_index[a] >= 0) {
          const std::size_t k = static_cast<std::size_t>(nullable_index[a]);
          if (const std::uint8_t *f = flags_of(a, r0))
            for (std::uint64_t i = 0; i < n; ++i) if (f[i]) setMsbFirstBit(w + i * slot_bytes, tuple_null_word, k);
          s.null_kind = QS_NULL_SLOT_WORD;
          s.null_bitmap = w + (k / (8 * tuple_null_word)) * tuple_null_word;
          s.null_arg = static_cast<std::uint32_t>(k % (8 * tuple_null_word));
          s.null_stride = static_cast<std::uint32_t>(slot_bytes);
          s.null_width = static_cast<std::uint32_t>(tuple_null_word);
        }
        attr_off += vw;
      }
      return;
    }
    // dictionaries first, then the NULL bitmaps, then the stripes (CompressedTupleStorageSubBlock.cpp:281-342,
    // CompressedColumnStoreTupleStorageSubBlock.cpp:755-798; BasicColumnStoreTupleStorageSubBlock.cpp:152-175)
    std::vector<const char *> dict_at(schema.size(), nullptr);
    for (std::size_t a = 0; a < schema.size(); ++a) {
      const StripePlan &p = plans[b][a];
      if (!p.dict.empty()) { std::memcpy(w, p.dict.data(), p.dict.size()); dict_at[a] = w; w += p.dict.size(); }
    }
    std::vector<const char *> bitmap_at(schema.size(), nullptr);
    for (std::size_t a = 0; a < schema.size(); ++a) {
      const StripePlan &p = plans[b][a];
      const bool wants = layout == TupleStoreLayout::kCompressedColumnStore ? (p.has_nulls && p.encoding != QS_ENC_DICT)
                                                                           : nullable_index[a] >= 0;
      if (!wants) continue;
      if (const std::uint8_t *f = flags_of(a, r0))
        for (std::uint64_t i = 0; i < n; ++i) if (f[i]) setMsbFirstBit(w, 8, i);
      bitmap_at[a] = w;
      w += bitVectorBytes(n);
    }
    for (std::size_t a = 0; a < schema.size(); ++a) {
      const StripePlan &p = plans[b][a];
      const char *col = static_cast<const char *>(columns[a]) + r0 * schema[a].width;
      writeAttr(schema[a], w, col, n, p, flags_of(a, r0));
      qs_stage_desc &s = B.stripes[a];
      s = qs_stage_desc{};
      s.attr = static_cast<std::uint32_t>(a); s.encoding = p.encoding; s.host = w; s.code_width = p.code_width;
      if (p.encoding == QS_ENC_DICT) {
        s.dict = dict_at[a]; s.dict_entries = static_cast<std::uint32_t>(p.dict.size() / schema[a].width);
        if (p.has_nulls) { s.null_kind = QS_NULL_CODE; s.null_arg = s.dict_entries; }
      }
      if (bitmap_at[a]) { s.null_kind = QS_NULL_BITMAP; s.null_bitmap = bitmap_at[a]; s.null_arg = 0; s.null_stride = 1; }
      w += p.stripe_bytes;
    }
  });

  std::lock_guard<std::mutex> lk(mu_);
  slabs_[rel->getID()].push_back(slab);
  for (StorageBlock &B : built) {
    B.id = next_block_++;
    rel->addBlock(B.id);
    blocks_.emplace(B.id, std::move(B));
  }
}

const StorageBlock &StorageManager::getBlock(block_id id) const {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = blocks_.find(id);
  QS_CHECK(it != blocks_.end());
  return it->second;
}

std::uint64_t StorageManager::hostBytes(const CatalogRelation &rel) const {
  std::lock_guard<std::mutex> lk(mu_);
  std::uint64_t n = 0;
  auto it = slabs_.find(rel.getID());
  if (it != slabs_.end()) for (const Slab &s : it->second) n += s.bytes;
  return n;
}

namespace {

// a < b in the attribute's own order (the order CompressionDictionaryBuilder sorts a block dictionary in,
// compression/CompressionDictionaryBuilder.cpp:120-160): numeric, DateLit lexicographic, strncmp
bool valueLess(std::uint16_t type, std::uint32_t width, const char *a, const char *b) {
  switch (type) {
    case QS_INT: { std::int32_t x, y; std::memcpy(&x, a, 4); std::memcpy(&y, b, 4); return x < y; }
    case QS_LONG: { std::int64_t x, y; std::memcpy(&x, a, 8); std::memcpy(&y, b, 8); return x < y; }
    case QS_FLOAT: { float x, y; std::memcpy(&x, a, 4); std::memcpy(&y, b, 4); return x < y; }
    case QS_DOUBLE: { double x, y; std::memcpy(&x, a, 8); std::memcpy(&y, b, 8); return x < y; }
    case QS_DATE: {
      std::int32_t x, y; std::memcpy(&x, a, 4); std::memcpy(&y, b, 4);
      if (x != y) return x < y;
      if (a[4] != b[4]) return static_cast<unsigned char>(a[4]) < static_cast<unsigned char>(b[4]);
      return static_cast<unsigned char>(a[5]) < static_cast<unsigned char>(b[5]);
    }
    default: return std::strncmp(a, b, width) < 0;
  }
}

}  // namespace

const std::vector<StorageManager::RelationDictionary> &StorageManager::relationDictionaries(
    const CatalogRelation &rel, const std::vector<block_id> &ids) {
  const auto key = std::make_pair(rel.getID(), ids.size());
  auto it = dictionaries_.find(key);
  if (it != dictionaries_.end()) return it->second;
  const std::vector<qs_attr> schema = rel.schema();
  std::vector<RelationDictionary> out(schema.size());
  for (std::size_t a = 0; a < schema.size() && !ids.empty(); ++a) {
    const std::uint32_t w = schema[a].width;
    bool all_dict = true;
    std::size_t total = 0;
    for (block_id id : ids) {
      const qs_stage_desc &s = blocks_.at(id).stripes[a];
      if (s.encoding != QS_ENC_DICT) { all_dict = false; break; }
      total += s.dict_entries;
    }
    if (!all_dict || w <= 1) continue;            // a CHAR(1) code is no narrower than its value
    if (schema[a].type == QS_FLOAT || schema[a].type == QS_DOUBLE) {
      // The relation-wide dictionary is ordered and searched NUMERICALLY (comparisons on codes need code order = value
      // order): -0.0 would collapse onto 0.0 (decode no longer bit-identical) and a NaN never compares equal (its block
      // entry would be "missing").  Such attributes stay at native width.
      bool special = false;
      for (block_id id : ids) {
        const qs_stage_desc &s = blocks_.at(id).stripes[a];
        for (std::uint32_t e = 0; e < s.dict_entries && !special; ++e) {
          const char *v = static_cast<const char *>(s.dict) + static_cast<std::size_t>(e) * w;
          if (w == 4) { float f; std::uint32_t u; std::memcpy(&f, v, 4); std::memcpy(&u, v, 4); special = f != f || u == 0x80000000u; }
          else { double f; std::uint64_t u; std::memcpy(&f, v, 8); std::memcpy(&u, v, 8); special = f != f || u == 0x8000000000000000ull; }
        }
        if (special) break;
      }
      if (special) continue;
    }
    std::vector<const char *> entries;
    entries.reserve(total);
    for (block_id id : ids) {
      const qs_stage_desc &s = blocks_.at(id).stripes[a];
      for (std::uint32_t e = 0; e < s.dict_entries; ++e) entries.push_back(static_cast<const char *>(s.dict) + static_cast<std::size_t>(e) * w);
    }
    const std::uint16_t type = schema[a].type;
    std::sort(entries.begin(), entries.end(), [&](const char *x, const char *y) { return valueLess(type, w, x, y); });
    entries.erase(std::unique(entries.begin(), entries.end(),
                              [&](const char *x, const char *y) { return !valueLess(type, w, x, y) && !valueLess(type, w, y, x); }),
                  entries.end());
    if (entries.size() > 65536) continue;          // 4-byte codes into a multi-megabyte dictionary: a random gather per row
    const std::uint32_t cw = entries.size() <= 256 ? 1 : 2;
    if (cw >= w) continue;
    RelationDictionary &D = out[a];
    D.code_width = cw;
    D.n_entries = static_cast<std::uint32_t>(entries.size());
    D.values.resize(entries.size() * w);
    for (std::size_t e = 0; e < entries.size(); ++e) std::memcpy(&D.values[e * w], entries[e], w);
  }
  return dictionaries_.emplace(key, std::move(out)).first->second;
}

std::pair<std::uint32_t, std::uint32_t> StorageManager::residentCoding(const CatalogRelation &rel, std::uint32_t attr) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = resident_.find(rel.getID());
  if (it == resident_.end() || !it->second.handle) return {0, 0};
  std::uint32_t cw = 0, n = 0;
  QS_CHECK_GPU(qsgpu_relation_dictionary(it->second.handle, attr, &cw, &n, nullptr));
  return {cw, n};
}

qsgpu_relation_t StorageManager::deviceRelation(const CatalogRelation &rel, std::uint64_t needed_attrs) {
  std::lock_guard<std::mutex> lk(mu_);
  const std::vector<block_id> ids = rel.getBlocksSnapshot();
  const std::vector<qs_attr> schema = rel.schema();
  QS_CHECK(schema.size() <= 64);
  const std::uint64_t all = schema.size() == 64 ? ~0ull : ((1ull << schema.size()) - 1);
  needed_attrs &= all;
  Resident &R = resident_[rel.getID()];
  const bool rebuild = !R.handle || R.n_blocks_staged != ids.size();
  if (!rebuild && (needed_attrs & ~R.staged_attrs) == 0) return R.handle;
  if (rebuild) {
    // first use, or the relation grew since the image was built (blocks are append-only): (re)build it
    if (R.handle) QS_CHECK_GPU(qsgpu_relation_destroy(R.handle));
    std::uint64_t rows = 0;
    for (block_id id : ids) rows += static_cast<std::uint64_t>(blocks_.at(id).num_tuples);
    QS_CHECK_GPU(qsgpu_relation_create(device_, static_cast<std::uint32_t>(schema.size()), schema.data(),
                                       std::max<std::uint64_t>(rows, 1), &R.handle));
    if (rel.nullableMask()) QS_CHECK_GPU(qsgpu_relation_set_nullable(R.handle, rel.nullableMask()));
    if (code_resident_) {
      const std::vector<RelationDictionary> &dicts = relationDictionaries(rel, ids);
      for (std::size_t a = 0; a < dicts.size(); ++a)
        if (dicts[a].code_width && !((rel.nullableMask() >> a) & 1))      // NULL-able attributes stay at native width
          QS_CHECK_GPU(qsgpu_relation_set_dictionary(R.handle, static_cast<std::uint32_t>(a), dicts[a].code_width,
                                                     dicts[a].values.data(), dicts[a].n_entries));
    }
    R.n_blocks_staged = 0;
    R.rows = 0;
    R.staged_attrs = 0;
    R.runs.clear();
    for (block_id id : ids) {
      const StorageBlock &B = blocks_.at(id);
      block_first_row_[B.id] = {rel.getID(), R.rows};
      R.rows += static_cast<std::uint64_t>(B.num_tuples);
    }
  }
  const std::uint64_t to_stage = needed_attrs & ~R.staged_attrs;
  std::vector<std::vector<qs_stage_desc>> descs(ids.size());
  std::vector<qs_block_image> images(ids.size());
  for (std::size_t i = 0; i < ids.size(); ++i) {
    const StorageBlock &B = blocks_.at(ids[i]);
    descs[i] = B.stripes;
    for (std::size_t a = 0; a < descs[i].size(); ++a)
      if (!((to_stage >> a) & 1)) descs[i][a].encoding = QS_ENC_SKIP;
    images[i].host = B.memory; images[i].bytes = B.size; images[i].n_rows = static_cast<std::uint64_t>(B.num_tuples);
    images[i].descs = descs[i].data();
  }
  if (!images.empty()) {
    if (rebuild)
      QS_CHECK_GPU(qsgpu_stage_blocks(R.handle, static_cast<std::uint32_t>(images.size()), images.data(),
                                      static_cast<std::uint32_t>(schema.size())));
    else
      QS_CHECK_GPU(qsgpu_stage_columns(R.handle, 0, static_cast<std::uint32_t>(images.size()), images.data(),
                                       static_cast<std::uint32_t>(schema.size())));
  }
  R.n_blocks_staged = ids.size();
  R.staged_attrs |= to_stage;
  return R.handle;
}

std::vector<DeviceExtent> StorageManager::stagedExtents(const CatalogRelation &rel, std::uint64_t needed_attrs,
                                                        std::uint64_t max_rows) {
  qsgpu_relation_t h = deviceRelation(rel, needed_attrs);
  std::lock_guard<std::mutex> lk(mu_);
  Resident &R = resident_[rel.getID()];
  if (R.runs_max_rows == max_rows && !R.runs.empty() && R.runs.front().relation == h) return R.runs;
  R.runs.clear();
  DeviceExtent cur;
  std::uint64_t row = 0;
  for (block_id b : rel.getBlocksSnapshot()) {
    const std::uint64_t n = static_cast<std::uint64_t>(blocks_.at(b).num_tuples);
    if (cur.relation && (max_rows == 0 || row + n - cur.row_begin <= max_rows)) {
      cur.row_end = row + n;                       // grow the run of adjacent blocks
    } else {
      if (cur.relation) R.runs.push_back(cur);
      cur.relation = h; cur.row_begin = row; cur.row_end = row + n;
    }
    row += n;
  }
  if (cur.relation) R.runs.push_back(cur);
  R.runs_max_rows = max_rows;
  return R.runs;
}

std::vector<block_id> StorageManager::repartitionTemporary(const CatalogRelation &rel, attribute_id partition_attribute,
                                                           std::size_t num_partitions) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = temporaries_.find(rel.getID());
  QS_CHECK(it != temporaries_.end());
  std::uint64_t n = 0;
  QS_CHECK_GPU(qsgpu_relation_num_rows(it->second, &n));            // the producer has finished: its count is final
  const std::vector<qs_attr> schema = rel.schema();
  qsgpu_relation_t out = nullptr;
  QS_CHECK_GPU(qsgpu_relation_create(device_, static_cast<std::uint32_t>(schema.size()), schema.data(), std::max<std::uint64_t>(n, 1), &out));
  std::vector<std::uint64_t> offsets(num_partitions + 1, 0);
  QS_CHECK_GPU(qsgpu_hash_partition(it->second, static_cast<std::uint32_t>(partition_attribute), static_cast<std::uint32_t>(num_partitions),
                                    out, offsets.data()));
  QS_CHECK_GPU(qsgpu_relation_set_num_rows(out, n));
  QS_CHECK_GPU(qsgpu_relation_destroy(it->second));
  it->second = out;                                                  // the temporary IS the partitioned relation from now on
  std::vector<block_id> blocks;
  for (std::size_t p = 0; p < num_partitions; ++p) {
    DeviceExtent e;
    e.relation = out;
    e.row_begin = offsets[p];
    e.row_end = offsets[p + 1];
    const block_id b = next_block_++;
    partition_extents_[b] = e;
    blocks.push_back(b);
  }
  partition_blocks_[rel.getID()] = blocks;
  return blocks;
}

DeviceExtent StorageManager::blockExtent(block_id id) {
  std::lock_guard<std::mutex> lk(mu_);
  auto pe = partition_extents_.find(id);
  if (pe != partition_extents_.end()) return pe->second;
  auto t = blocks_.find(id);
  if (t == blocks_.end()) {          // pseudo block of a temporary relation: every row produced so far
    for (auto &kv : temporary_block_)
      if (kv.second == id) { DeviceExtent e; e.relation = temporaries_.at(kv.first); return e; }
    QS_CHECK(false);
  }
  auto f = block_first_row_.find(id);
  QS_CHECK(f != block_first_row_.end());     // deviceRelation() first
  DeviceExtent e;
  e.relation = resident_.at(f->second.first).handle;
  e.row_begin = f->second.second;
  e.row_end = e.row_begin + static_cast<std::uint64_t>(t->second.num_tuples);
  return e;
}

void StorageManager::evict(const CatalogRelation &rel) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = resident_.find(rel.getID());
  if (it == resident_.end()) return;
  if (it->second.handle) QS_CHECK_GPU(qsgpu_relation_destroy(it->second.handle));
  resident_.erase(it);
}

block_id StorageManager::createTemporary(const CatalogRelation &rel, std::uint64_t capacity_rows) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = temporaries_.find(rel.getID());
  if (it == temporaries_.end()) {
    const std::vector<qs_attr> schema = rel.schema();
    qsgpu_relation_t h = nullptr;
    QS_CHECK_GPU(qsgpu_relation_create(device_, static_cast<std::uint32_t>(schema.size()), schema.data(),
                                       std::max<std::uint64_t>(capacity_rows, 1), &h));
    temporaries_[rel.getID()] = h;
    temporary_block_[rel.getID()] = next_block_++;
  }
  return temporary_block_.at(rel.getID());
}

void StorageManager::adoptTemporary(const CatalogRelation &rel, qsgpu_relation_t handle) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = temporaries_.find(rel.getID());
  if (it != temporaries_.end() && it->second) QS_CHECK_GPU(qsgpu_relation_destroy(it->second));
  temporaries_[rel.getID()] = handle;
  if (!temporary_block_.count(rel.getID())) temporary_block_[rel.getID()] = next_block_++;
}

qsgpu_relation_t StorageManager::temporary(const CatalogRelation &rel) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = temporaries_.find(rel.getID());
  QS_CHECK(it != temporaries_.end());
  return it->second;
}

void StorageManager::setPartitioned(relation_id id, bool partitioned) {
  std::lock_guard<std::mutex> lk(mu_);
  partitioned_[id] = partitioned;
}

bool StorageManager::isPartitioned(relation_id id) const {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = partitioned_.find(id);
  return it != partitioned_.end() && it->second;
}

qsgpu_relation_t StorageManager::replicated(const CatalogRelation &rel) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = replicas_.find(rel.getID());
  if (it != replicas_.end()) return it->second;
  auto t = temporaries_.find(rel.getID());
  QS_CHECK(t != temporaries_.end());
  qsgpu_relation_t all = nullptr;
  QS_CHECK_GPU(qsgpu_relation_allgather(t->second, comm_, &all));
  replicas_[rel.getID()] = all;
  return all;
}

void StorageManager::dropTemporary(const CatalogRelation &rel) {
  std::lock_guard<std::mutex> lk(mu_);
  partitioned_.erase(rel.getID());
  auto pb = partition_blocks_.find(rel.getID());
  if (pb != partition_blocks_.end()) {
    for (block_id b : pb->second) partition_extents_.erase(b);
    partition_blocks_.erase(pb);
  }
  auto rb = result_blocks_.find(rel.getID());
  if (rb != result_blocks_.end()) {          // host blocks written by insertTuples(): buffers go back to the pool
    for (block_id b : rb->second) {
      block_pool_.push_back(const_cast<char *>(blocks_.at(b).memory));
      blocks_.erase(b);
    }
    result_blocks_.erase(rb);
  }
  auto rp = replicas_.find(rel.getID());
  if (rp != replicas_.end()) {
    if (rp->second) QS_CHECK_GPU(qsgpu_relation_destroy(rp->second));
    replicas_.erase(rp);
  }
  auto it = temporaries_.find(rel.getID());
  if (it == temporaries_.end()) return;
  if (it->second) QS_CHECK_GPU(qsgpu_relation_destroy(it->second));
  temporaries_.erase(it);
  temporary_block_.erase(rel.getID());
}

}  // namespace quickstep
