// Test driver (tests/test_parser_asan.py): parses every blob of a corpus file with the two wire decoders,
// built with -fsanitize=address,undefined so that any out-of-bounds read or UB in the decoders aborts.
// Corpus format: repeated { uint32 size, bytes }.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "summary_map.h"
#include "vi_map_reader.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint8_t> all;
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) all.insert(all.end(), buf, buf + n);
  std::fclose(f);
  size_t at = 0, parsed = 0, accepted = 0;
  while (at + 4 <= all.size()) {
    uint32_t size;
    std::memcpy(&size, all.data() + at, 4);
    at += 4;
    if (at + size > all.size()) return 3;
    // exact-size heap copy: reads past the end are caught by the sanitizer
    std::vector<uint8_t> blob(all.begin() + at, all.begin() + at + size);
    at += size;
    std::string err;
    mlc::SummaryMap sm;
    if (sm.Parse(blob.data(), blob.size(), &err)) {
      ++accepted;
      mlc::SummaryMapImages images;
      mlc::GroupSummaryMapByObserver(sm, &images, &err);
      std::vector<uint8_t> again;
      sm.Serialize(&again);
      if (again.size() != sm.SerializedSize()) return 4;
    }
    mlc::ViMapVertices vm;
    if (vm.Parse(blob.data(), blob.size(), &err)) ++accepted;
    mlc::ViMapMissions missions;
    if (missions.Parse(blob.data(), blob.size(), &err)) ++accepted;
    ++parsed;
  }
  std::printf("%zu blobs, %zu accepted\n", parsed, accepted);
  return 0;
}
