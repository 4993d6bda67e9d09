// difacto_b200/host/tests/batch_dump.cc -- prints order-sensitive checksums of the minibatches a BatchReader produces
// (test tool: tests/test_host_cpp.py compares them with the compiled reference's BatchReader, batch by batch).
// usage: batch_dump <file> <part> <nparts> <batch_size> <shuffle_buf> <neg_sampling> <nbatches> <reference|seeded>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "difacto_b200/data.h"

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage: %s file part nparts batch shuffle_buf neg_sampling nbatches reference|seeded\n", argv[0]); return 2; }
  using namespace difacto;  // NOLINT
  try {
    const ShuffleOrder order = strcmp(argv[8], "reference") == 0 ? ShuffleOrder::kReference : ShuffleOrder::kSeeded;
    BatchReader reader(argv[1], "libsvm", static_cast<unsigned>(atoi(argv[2])), static_cast<unsigned>(atoi(argv[3])),
                       static_cast<unsigned>(atoi(argv[4])), static_cast<unsigned>(atoi(argv[5])), static_cast<float>(atof(argv[6])),
                       0, 0, order);
    const int nb = atoi(argv[7]);
    for (int b = 0; b < nb && reader.Next(); ++b) {
      const auto blk = reader.Value();
      const size_t nnz = blk.offset[blk.size];
      unsigned long long hi = 0, hl = 0;
      for (size_t i = 0; i < nnz; ++i) hi = hi * 1000003ULL + blk.index[i];                  // order-sensitive
      for (size_t r = 0; r < blk.size; ++r) hl = hl * 1000003ULL + static_cast<unsigned long long>(blk.label[r] > 0 ? 1 : 2);
      printf("%d %zu %zu %llu %llu %d\n", b, static_cast<size_t>(blk.size), nnz, hi, hl, blk.value ? 1 : 0);
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "batch_dump: %s\n", e.what());
    return 1;
  }
  return 0;
}
