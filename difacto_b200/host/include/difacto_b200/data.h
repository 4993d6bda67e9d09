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

/**
 * one part of a text file: the lines that START inside byte range [size*i/n, size*(i+1)/n), parsed as
 * "label idx[:val] idx[:val] ..." (dmlc LibSVMParser semantics).  The part is streamed: the file is read in
 * chunks of kChunkBytes and rows are handed out on demand, so the memory held is one chunk plus the caller's rows
 * (the reference's Reader streams 64 MB chunks the same way, batch_reader.cc:26).
 */
class LibsvmPartReader {
 public:
  static constexpr size_t kChunkBytes = 8u << 20;
  LibsvmPartReader(const std::string& path, unsigned part, unsigned nparts, size_t chunk_bytes = kChunkBytes)
      : f_(path, std::ios::binary | std::ios::ate), chunk_(chunk_bytes ? chunk_bytes : kChunkBytes) {
    if (!f_) throw Error("failed to open " + path);
    const size_t size = static_cast<size_t>(f_.tellg());
    size_t begin = size * part / nparts, end = size * (part + 1) / nparts;
    auto align = [&](size_t pos) {     // first line start at or after pos
      if (pos == 0 || pos >= size) return std::min(pos, size);
      f_.seekg(static_cast<std::streamoff>(pos - 1));
      char c;
      while (f_.get(c)) { if (c == '\n') break; }
      const size_t at = f_ ? static_cast<size_t>(f_.tellg()) : size;
      f_.clear();
      return at;
    };
    next_ = align(begin);
    end_ = align(end);
    f_.seekg(static_cast<std::streamoff>(next_));
  }
  /** append up to max_rows parsed rows to out; returns the number appended (0: the part is exhausted) */
  size_t ReadRows(RowBlockContainer<feaid_t>* out, size_t max_rows) {
    size_t got = 0;
    while (got < max_rows) {
      const char* p = buf_.data() + pos_;
      const char* e = buf_.data() + buf_.size();
      const char* le = p < e ? static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(e - p))) : nullptr;
      if (!le) {
        if (next_ < end_) { Fill(); continue; }     // the line continues in the next chunk
        if (p >= e) break;
        le = e;                                      // last line of the part without a trailing newline
      }
      got += ParseLine(p, le, out);
      pos_ = static_cast<size_t>(le - buf_.data()) + 1;
      if (pos_ > buf_.size()) pos_ = buf_.size();
    }
    return got;
  }
  /** parse everything that is left */
  void ParseAll(RowBlockContainer<feaid_t>* out) {
    out->Clear();
    while (ReadRows(out, 1u << 20)) {}
  }

 private:
  void Fill() {
    buf_.erase(0, pos_);
    pos_ = 0;
    const size_t n = std::min(chunk_, end_ - next_);
    const size_t old = buf_.size();
    buf_.resize(old + n);
    f_.read(&buf_[old], static_cast<std::streamsize>(n));
    next_ += n;
  }
  static size_t ParseLine(const char* q, const char* le, RowBlockContainer<feaid_t>* out) {
    while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
    if (q >= le) return 0;
    // strtof / strtoull stop at the newline (or at the terminating NUL std::string keeps behind the last byte)
    char* nx = nullptr;
    const float label = strtof(q, &nx);
    if (nx == q) return 0;
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
    return 1;
  }
  std::ifstream f_;
  std::string buf_;
  size_t chunk_, pos_ = 0, next_ = 0, end_ = 0;
};

/**
 * fixed-size minibatches with the reference's options (src/reader/batch_reader.cc:8-78):
 * shuffle_buf_size > 0 draws the batch from a shuffled window of that many rows; neg_sampling < 1
 * drops negative rows with probability 1 - neg_sampling (rand_r, seed 0, as the reference);
 * all-ones value arrays are dropped (:71-73).  Rows are streamed window by window (a shuffle window, or
 * max(batch_size, 4096) rows): the file part is never held in memory as a whole.
 *
 * Two shuffle orders:
 *  kReference  the reference's own order on this toolchain: std::random_shuffle (batch_reader.cc:45) is, in libstdc++,
 *              "for i in 1..n-1: swap(a[i], a[rand() % (i + 1)])" on the process-wide rand() state; the permutation
 *              vector persists from window to window and is reset only when the window size changes (:40-44); the
 *              down-sampling draws start from rand_r seed 0 (:16).  Every epoch sees another order because the global
 *              generator moves on.  Checked against the compiled reference (tests/test_host_cpp.py).  rand() is not
 *              for concurrent callers: one reader thread at a time.
 *  kSeeded     std::mt19937 seeded by (epoch, part): another order every epoch, reproducible whatever else runs in
 *              the process -- used where several readers parse concurrently (num_gpus > 1).
 */
enum class ShuffleOrder { kSeeded, kReference };

class BatchReader {
 public:
  BatchReader(const std::string& uri, const std::string& format, unsigned part, unsigned nparts,
              unsigned batch_size, unsigned shuffle_buf_size = 0, float neg_sampling = 1.0f, unsigned epoch = 0,
              size_t chunk_bytes = 0, ShuffleOrder order = ShuffleOrder::kSeeded)
      : batch_size_(batch_size), shuf_buf_(shuffle_buf_size), neg_sampling_(neg_sampling),
        seed_(order == ShuffleOrder::kReference ? 0u : epoch * 2654435761u + part),
        src_(Open(uri, format, part, nparts, chunk_bytes)), mode_(order),
        gen_(epoch * 2654435761u + part * 40503u + 1u) {
    if (shuf_buf_) DFB_CHECK(shuf_buf_ >= batch_size_);
  }
  bool Next() {
    batch_.Clear();
    while (batch_.Size() < batch_size_) {
      if (cursor_ == order_.size() && !Refill()) break;
      const unsigned j = order_[cursor_++];
      if (shuf_buf_ != 0 || neg_sampling_ != 1.0f) {
        const float p = static_cast<float>(rand_r(&seed_)) / static_cast<float>(RAND_MAX);
        if (neg_sampling_ < 1.0f && win_.label[j] <= 0 && p > 1 - neg_sampling_) continue;
      }
      for (size_t t = win_.offset[j]; t < win_.offset[j + 1]; ++t) {
        batch_.index.push_back(win_.index[t]);
        batch_.value.push_back(win_.value[t]);
      }
      batch_.label.push_back(win_.label[j]);
      batch_.offset.push_back(batch_.index.size());
    }
    bool binary = true;
    for (float f : batch_.value) if (f != 1) { binary = false; break; }
    if (binary) batch_.value.clear();
    return batch_.Size() > 0;
  }
  dmlc::RowBlock<feaid_t> Value() const { return batch_.GetBlock(); }

 private:
  static LibsvmPartReader Open(const std::string& uri, const std::string& format, unsigned part, unsigned nparts,
                               size_t chunk_bytes) {
    if (format != "libsvm") throw Error("unknown format " + format + " (this build reads libsvm)");
    return LibsvmPartReader(uri, part, nparts, chunk_bytes);
  }
  // the next window of rows: exactly shuffle_buf_size rows (the last one shorter), shuffled; or the next rows in file order
  bool Refill() {
    win_.Clear();
    const size_t want = shuf_buf_ ? shuf_buf_ : std::max<size_t>(batch_size_, 4096);
    while (win_.Size() < want && src_.ReadRows(&win_, want - win_.Size())) {}
    if (shuf_buf_ && mode_ == ShuffleOrder::kReference) {
      if (order_.size() != win_.Size()) {
        order_.resize(win_.Size());
        std::iota(order_.begin(), order_.end(), 0u);
      }
      for (size_t i = 1; i < order_.size(); ++i) {       // libstdc++'s std::random_shuffle(first, last)
        const size_t j = static_cast<size_t>(std::rand()) % (i + 1);
        if (i != j) std::swap(order_[i], order_[j]);
      }
    } else {
      order_.resize(win_.Size());
      std::iota(order_.begin(), order_.end(), 0u);
      if (shuf_buf_) std::shuffle(order_.begin(), order_.end(), gen_);
    }
    cursor_ = 0;
    return !order_.empty();
  }
  unsigned batch_size_, shuf_buf_;
  float neg_sampling_;
  unsigned int seed_;
  LibsvmPartReader src_;
  ShuffleOrder mode_;
  std::mt19937 gen_;
  RowBlockContainer<feaid_t> win_, batch_;
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
