// TEST INFRASTRUCTURE: opens block images written by THIS repository's host layer (StorageManager::insertTuples, the GPU ->
// host hand-off of result rows; dumped by quickstep_b200/lib/qshost_unittest <dir>) with the REFERENCE's real StorageBlock
// class, linked from the unmodified engine's build tree.  The constructor parses and validates the StorageBlockHeader
// (storage/StorageBlock.cpp:114-156: MalformedBlock otherwise) and builds the real SplitRowStoreTupleStorageSubBlock over
// the bytes; every tuple is then read through the sub-block's own accessors and compared with what the unit test wrote.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "catalog/CatalogAttribute.hpp"
#include "catalog/CatalogRelation.hpp"
#include "storage/StorageBlock.hpp"
#include "storage/StorageBlockInfo.hpp"
#include "storage/StorageBlockLayout.hpp"
#include "storage/TupleStorageSubBlock.hpp"
#include "types/DatetimeLit.hpp"
#include "types/TypeFactory.hpp"
#include "types/TypeID.hpp"

using namespace quickstep;  // NOLINT

#define EXPECT(cond)                                                              \
  do {                                                                            \
    if (!(cond)) {                                                                \
      std::fprintf(stderr, "%s:%d: EXPECT(%s) failed\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                               \
    }                                                                             \
  } while (0)

int main(int argc, char **argv) {
  EXPECT(argc >= 2);
  // the relation of host_unittest.cpp::testInsertTuples
  CatalogRelation rel(nullptr, "result", 9, true);
  rel.addAttribute(new CatalogAttribute(&rel, "a", TypeFactory::GetType(kInt, false)));
  rel.addAttribute(new CatalogAttribute(&rel, "b", TypeFactory::GetType(kDouble, false)));
  rel.addAttribute(new CatalogAttribute(&rel, "c", TypeFactory::GetType(kChar, 10, false)));
  rel.addAttribute(new CatalogAttribute(&rel, "d", TypeFactory::GetType(kDate, false)));
  rel.addAttribute(new CatalogAttribute(&rel, "e", TypeFactory::GetType(kLong, false)));
  std::unique_ptr<StorageBlockLayout> layout(StorageBlockLayout::GenerateDefaultLayout(rel, false));   // unused for an existing block

  std::uint64_t row = 0;
  for (int file_no = 0;; ++file_no) {
    std::ifstream in(std::string(argv[1]) + "/block_" + std::to_string(file_no) + ".bin", std::ios::binary);
    if (!in.good()) break;
    std::vector<char> image((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    EXPECT(image.size() == kSlotSizeBytes);                       // one slot: the size of a temporary relation's block
    void *memory = nullptr;
    EXPECT(posix_memalign(&memory, 64, image.size()) == 0);
    std::memcpy(memory, image.data(), image.size());
    {
      StorageBlock block(rel, BlockIdUtil::GetBlockId(1, 1 + file_no), *layout, false, memory, image.size());     // throws MalformedBlock
      const TupleStorageSubBlock &store = block.getTupleStorageSubBlock();
      EXPECT(store.getTupleStorageSubBlockType() == kSplitRowStore);
      EXPECT(store.isPacked());
      const tuple_id n = store.numTuples();
      EXPECT(n > 0 && store.getMaxTupleID() == n - 1);
      for (tuple_id t = 0; t < n; ++t, ++row) {
        EXPECT(store.hasTupleWithID(t));
        EXPECT(*static_cast<const std::int32_t *>(store.getAttributeValue(t, 0)) == static_cast<std::int32_t>(row) - 50000);
        EXPECT(*static_cast<const double *>(store.getAttributeValue(t, 1)) == static_cast<double>(row) * 0.25 - 7.5);
        char expect[10] = {0};
        std::snprintf(expect, 10, "row%llu", static_cast<unsigned long long>(row % 1000));
        EXPECT(std::strncmp(static_cast<const char *>(store.getAttributeValue(t, 2)), expect, 10) == 0);
        const DateLit &d = *static_cast<const DateLit *>(store.getAttributeValue(t, 3));
        EXPECT(d.year == 1992 + static_cast<std::int32_t>(row % 7) && d.month == 1 + row % 12 && d.day == 1 + row % 28);
        EXPECT(store.getAttributeValueTyped(t, 4).getLiteral<std::int64_t>() == static_cast<std::int64_t>(row) * 1000003ll - (1ll << 40));
      }
    }
    std::free(memory);
  }
  EXPECT(row == 100000);
  std::printf("reference StorageBlock read %llu tuples\n", static_cast<unsigned long long>(row));
  return 0;
}
