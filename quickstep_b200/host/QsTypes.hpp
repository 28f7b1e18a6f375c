// Stand-ins for the handful of reference types the operator layer names in its
// signatures.  In an in-tree build these are the reference's
// own classes; here they carry exactly the members the hot path reads, so the
// GPU operators compile and are tested without the reference's CMake build.
//
//   block_id, relation_id, attribute_id, partition_id   storage/StorageBlockInfo.hpp:38-60,
//                                                       catalog/CatalogTypedefs.hpp:39-60
//   CatalogAttribute / CatalogRelation                  catalog/CatalogRelation.hpp:62
//   tmb::client_id / tmb::MessageBus                    third_party/src/tmb/include/tmb/id_typedefs.h
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "qsgpu.h"

namespace quickstep {

typedef std::uint64_t block_id;
typedef int relation_id;
typedef int attribute_id;
typedef std::size_t partition_id;
typedef std::int64_t tuple_id;

namespace tmb_standin {
typedef std::uint32_t client_id;
class MessageBus;   // the GPU operators pass it through untouched, like the CPU ones
}  // namespace tmb_standin
namespace tmb = tmb_standin;

// The reference's error convention on this path: LOG(FATAL) / CHECK abort
// (relational_operators/BuildHashOperator.cpp:205-206).  A failed C-ABI call is
// fatal for the query; nothing retries a work order.
[[noreturn]] inline void qs_fatal(const char *what, int status) {
  std::fprintf(stderr, "FATAL %s: status %d: %s\n", what, status, qsgpu_last_error());
  std::abort();
}
#define QS_CHECK_GPU(call)                              \
  do {                                                  \
    const int st__ = (call);                            \
    if (st__ != 0) ::quickstep::qs_fatal(#call, st__);  \
  } while (0)
#define QS_CHECK(cond)                                                          \
  do {                                                                          \
    if (!(cond)) { std::fprintf(stderr, "FATAL CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); std::abort(); } \
  } while (0)

struct CatalogAttribute {
  std::string name;
  qs_attr type;          // {QS_INT.., byte width}
  bool nullable = false; // Type::isNullable() (types/Type.hpp:129)
};

typedef std::vector<block_id> BlocksInPartition;

// catalog/CatalogRelation.hpp: id, attributes, block list (getBlocksSnapshot),
// partition count.  Temporary relations start with no blocks.
class CatalogRelation {
 public:
  CatalogRelation(relation_id id, std::string name, std::vector<CatalogAttribute> attrs, bool temporary = false)
      : id_(id), name_(std::move(name)), attrs_(std::move(attrs)), temporary_(temporary) {}
  relation_id getID() const { return id_; }
  const std::string &getName() const { return name_; }
  bool isTemporary() const { return temporary_; }
  std::size_t size() const { return attrs_.size(); }
  const CatalogAttribute &getAttributeById(attribute_id a) const { return attrs_[static_cast<std::size_t>(a)]; }
  std::vector<qs_attr> schema() const {
    std::vector<qs_attr> s;
    for (const auto &a : attrs_) s.push_back(a.type);
    return s;
  }
  // bit a = attribute a has a NULL-able type (CatalogRelationSchema::hasNullableAttributes / numNullableAttributes)
  std::uint64_t nullableMask() const {
    std::uint64_t m = 0;
    for (std::size_t a = 0; a < attrs_.size() && a < 64; ++a) if (attrs_[a].nullable) m |= 1ull << a;
    return m;
  }
  std::size_t getNumPartitions() const { return 1u; }
  bool hasPartitionScheme() const { return false; }
  void addBlock(block_id b) { std::lock_guard<std::mutex> lk(mu_); blocks_.push_back(b); }
  std::vector<block_id> getBlocksSnapshot() const { std::lock_guard<std::mutex> lk(mu_); return blocks_; }

 private:
  relation_id id_;
  std::string name_;
  std::vector<CatalogAttribute> attrs_;
  bool temporary_;
  mutable std::mutex mu_;
  std::vector<block_id> blocks_;
};

}  // namespace quickstep
