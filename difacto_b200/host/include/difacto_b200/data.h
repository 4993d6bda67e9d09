// difacto_b200/host/include/difacto_b200/data.h -- host-side input path of the SGD learner:
// libsvm reader with dmlc-style file partitioning, BatchReader (src/reader/batch_reader.{h,cc}) and
// Localizer (src/data/localizer.{h,cc}) restated for the host.  These feed the C-ABI; the GPU
// localizer replaces Localizer::Compact when raw ids are handed to the engine.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "api.h"

namespace difacto {

/** one part of a text file: the lines that START inside byte range [size*i/n, size*(i+1)/n) */
class LibsvmPartReader {
 public:
  LibsvmPartReader(const std::string& path, unsigned part, unsigned nparts) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw Error("failed to open " + path);
    const size_t size = static_cast<size_t>(f.tellg());
    size_t begin = size * part / nparts, end = size * (part + 1) / nparts;
    auto align = [&](size_t pos) {     // first line start at or after pos
      if (pos == 0 || pos >= size) return std::min(pos, size);
      f.seekg(static_cast<std::streamoff>(pos - 1));
      char c;
      while (f.get(c)) { if (c == '\n') break; }
      return f ? static_cast<size_t>(f.tellg()) : size;
    };
    begin = align(begin);
    f.clear();
    end = align(end);
    f.clear();
    buf_.resize(end > begin ? end - begin : 0);
    f.seekg(static_cast<std::streamoff>(begin));
    if (!buf_.empty()) f.read(&buf_[0], static_cast<std::streamsize>(buf_.size()));
  }
  /** parse everything: "label idx:val idx:val ..." per line (dmlc LibSVMParser semantics) */
  void ParseAll(RowBlockContainer<feaid_t>* out) const {
    out->Clear();
    const char* p = buf_.data();
    const char* e = p + buf_.size();
    while (p < e) {
      const char* le = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(e - p)));
      if (!le) le = e;
      const char* q = p;
      while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
      if (q < le) {
        char* nx = nullptr;
        const float label = strtof(q, &nx);
        if (nx != q) {
          q = nx;
          while (q < le) {
            while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
            if (q >= le) break;
            const feaid_t idx = strtoull(q, &nx, 10);
            if (nx == q) break;
            q = nx;
            float val = 1.f;
            if (q < le && *q == ':') { ++q; val = strtof(q, &nx); q = nx; }
            out->index.push_back(idx);
            out->value.push_back(val);
          }
          out->label.push_back(label);
          out->offset.push_back(out->index.size());
        }
      }
      p = le + 1;
    }
  }

 private:
  std::string buf_;
};

/**
 * fixed-size minibatches with the reference's options (src/reader/batch_reader.cc:8-78):
 * shuffle_buf_size > 0 draws the batch from a shuffled window of that many rows; neg_sampling < 1
 * drops negative rows with probability 1 - neg_sampling (rand_r, seed 0, as the reference);
 * all-ones value arrays are dropped (:71-73).  The shuffle order uses std::mt19937 instead of the
 * reference's std::random_shuffle (whose order is implementation-defined): same distribution,
 * different permutation.
 */
class BatchReader {
 public:
  BatchReader(const std::string& uri, const std::string& format, unsigned part, unsigned nparts,
              unsigned batch_size, unsigned shuffle_buf_size = 0, float neg_sampling = 1.0f, unsigned epoch = 0)
      : batch_size_(batch_size), shuf_buf_(shuffle_buf_size), neg_sampling_(neg_sampling),
        seed_(epoch * 2654435761u + part) {
    if (format != "libsvm") throw Error("unknown format " + format + " (this build reads libsvm)");
    if (shuf_buf_) DFB_CHECK(shuf_buf_ >= batch_size_);
    LibsvmPartReader(uri, part, nparts).ParseAll(&all_);
    order_.resize(all_.Size());
    std::iota(order_.begin(), order_.end(), 0u);
    if (shuf_buf_) {
      // the reference's std::random_shuffle advances one global generator, so every epoch sees another order;
      // here the order is a function of (epoch, part): different per epoch, reproducible per run
      std::mt19937 gen(epoch * 2654435761u + part * 40503u + 1u);
      for (size_t b = 0; b < order_.size(); b += shuf_buf_) {
        const size_t e = std::min(order_.size(), b + shuf_buf_);
        std::shuffle(order_.begin() + static_cast<std::ptrdiff_t>(b), order_.begin() + static_cast<std::ptrdiff_t>(e), gen);
      }
    }
  }
  bool Next() {
    batch_.Clear();
    while (batch_.Size() < batch_size_ && cursor_ < order_.size()) {
      const unsigned j = order_[cursor_++];
      if (shuf_buf_ != 0 || neg_sampling_ != 1.0f) {
        const float p = static_cast<float>(rand_r(&seed_)) / static_cast<float>(RAND_MAX);
        if (neg_sampling_ < 1.0f && all_.label[j] <= 0 && p > 1 - neg_sampling_) continue;
      }
      for (size_t t = all_.offset[j]; t < all_.offset[j + 1]; ++t) {
        batch_.index.push_back(all_.index[t]);
        batch_.value.push_back(all_.value[t]);
      }
      batch_.label.push_back(all_.label[j]);
      batch_.offset.push_back(batch_.index.size());
    }
    bool binary = true;
    for (float f : batch_.value) if (f != 1) { binary = false; break; }
    if (binary) batch_.value.clear();
    return batch_.Size() > 0;
  }
  dmlc::RowBlock<feaid_t> Value() const { return batch_.GetBlock(); }

 private:
  unsigned batch_size_, shuf_buf_;
  float neg_sampling_;
  unsigned int seed_;
  RowBlockContainer<feaid_t> all_, batch_;
  std::vector<unsigned> order_;
  size_t cursor_ = 0;
};

/**
 * Localizer::Compact (src/data/localizer.h:41-51, localizer.cc:11-103): key = ReverseBytes(id %
 * max_index); sorted unique keys, occurrence counts, indices remapped to ranks.  Bit-exact.
 */
class Localizer {
 public:
  explicit Localizer(feaid_t max_index = std::numeric_limits<feaid_t>::max(), int nthreads = 2)
      : max_index_(max_index) { (void)nthreads; }
  void Compact(const dmlc::RowBlock<feaid_t>& blk, RowBlockContainer<unsigned>* compacted,
               std::vector<feaid_t>* uniq_idx = nullptr, std::vector<real_t>* idx_frq = nullptr) {
    compacted->Clear();
    if (uniq_idx) uniq_idx->clear();
    if (idx_frq) idx_frq->clear();
    if (blk.size == 0) return;
    const size_t nnz = blk.offset[blk.size] - blk.offset[0];
    const size_t o0 = blk.offset[0];
    std::vector<std::pair<feaid_t, unsigned>> pr(nnz);
    for (size_t i = 0; i < nnz; ++i) pr[i] = std::make_pair(ReverseBytes(blk.index[o0 + i] % max_index_), static_cast<unsigned>(i));
    std::sort(pr.begin(), pr.end());
    compacted->index.resize(nnz);
    std::vector<feaid_t> keys;
    unsigned rank = 0;
    for (size_t i = 0; i < nnz; ++i) {
      if (i == 0 || pr[i].first != pr[i - 1].first) {
        if (i) ++rank;
        keys.push_back(pr[i].first);
        if (idx_frq) idx_frq->push_back(0);
      }
      if (idx_frq) idx_frq->back() += 1;
      compacted->index[pr[i].second] = rank;
    }
    compacted->offset.resize(blk.size + 1);
    for (size_t i = 0; i <= blk.size; ++i) compacted->offset[i] = blk.offset[i] - o0;
    if (blk.value) compacted->value.assign(blk.value + o0, blk.value + o0 + nnz);
    if (blk.label) compacted->label.assign(blk.label, blk.label + blk.size);
    if (uniq_idx) uniq_idx->swap(keys);
  }

 private:
  feaid_t max_index_;
};

}  // namespace difacto
