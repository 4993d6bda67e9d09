// difacto_b200/host/include/difacto_b200/api.h
//
// Host-side mirror of the reference's plugin interface for the SGD path, namespace `difacto`:
// the same class names, method names, argument meaning and "Init returns the unconsumed kwargs"
// convention as include/difacto/{base,sarray,loss,updater,store,learner}.h of dmlc/difacto, so
// that code (and tests) written against the reference read the same here.  Nothing is copied:
// the containers are small self-contained restatements (SArray here is a shared_ptr-backed
// array, not ps::SArray; RowBlock is a plain struct with dmlc::RowBlock's fields).
//
// What sits behind these classes is the C-ABI of include/difacto_b200.h (sm_100a CUDA); there is
// no CPU implementation in this tree outside oracle/ (test infrastructure).
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace difacto {

typedef float real_t;        // include/difacto/base.h:16
typedef uint64_t feaid_t;    // include/difacto/base.h:20
typedef std::vector<std::pair<std::string, std::string>> KWArgs;   // base.h:24

/** the reference CHECK()s abort the process; here they throw so that a host can recover */
struct Error : public std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};
/** dmlc::ParamError equivalent (missing / out-of-range parameter) */
struct ParamError : public Error {
  explicit ParamError(const std::string& m) : Error(m) {}
};

#define DFB_CHECK(cond)                                                                      \
  if (!(cond)) throw ::difacto::Error(std::string("Check failed: " #cond " at ") + __FILE__ + ":" + \
                                      std::to_string(__LINE__))

/** reverse the 16 nibbles of a feature id (ReverseBytes, include/difacto/base.h:39-51) */
inline feaid_t ReverseBytes(feaid_t x) {
  feaid_t r = 0;
  for (int i = 0; i < 16; ++i) { r = (r << 4) | (x & 0xF); x >>= 4; }
  return r;
}
/** EncodeFeaGrpID / DecodeFeaGrpID, base.h:60-72 */
inline feaid_t EncodeFeaGrpID(feaid_t x, int gid, int nbits) {
  DFB_CHECK(gid >= 0 && gid < (1 << nbits));
  return (x << nbits) | static_cast<feaid_t>(gid);
}
inline feaid_t DecodeFeaGrpID(feaid_t x, int nbits) { return x % (1u << nbits); }

/**
 * shared array with pointer-copy semantics (the role ps::SArray plays in the reference):
 * copies share the buffer; SArray<char>(SArray<T>) / SArray<T>(SArray<char>) reinterpret without
 * copying (used for Loss::Predict's `param` vector); SArray(ptr, n) wraps foreign memory.
 */
template <typename T>
class SArray {
 public:
  SArray() {}
  explicit SArray(size_t n, T init = T()) { resize(n, init); }
  SArray(T* data, size_t n, bool deletable = false) {
    if (deletable) ptr_.reset(data, [](T* p) { delete[] p; });
    else ptr_.reset(data, [](T*) {});
    size_ = cap_ = n;
  }
  explicit SArray(const std::vector<T>& v) { CopyFrom(v.data(), v.size()); }
  explicit SArray(const std::shared_ptr<std::vector<T>>& v) {
    ptr_ = std::shared_ptr<T>(v, v->data());
    size_ = cap_ = v->size();
  }
  template <typename U>
  SArray(const SArray<U>& o) {   // NOLINT: zero-copy reinterpretation, like ps::SArray
    ptr_ = std::shared_ptr<T>(o.ptr(), reinterpret_cast<T*>(o.data()));
    size_ = cap_ = o.size() * sizeof(U) / sizeof(T);
  }
  void CopyFrom(const T* src, size_t n) {
    resize(n);
    if (n) memcpy(data(), src, n * sizeof(T));
  }
  void CopyFrom(const SArray<T>& o) { CopyFrom(o.data(), o.size()); }
  void resize(size_t n, T init = T()) {
    if (n > cap_) {
      T* p = new T[n + 8];
      if (size_) memcpy(p, ptr_.get(), size_ * sizeof(T));
      ptr_.reset(p, [](T* q) { delete[] q; });
      cap_ = n;
    }
    for (size_t i = size_; i < n; ++i) ptr_.get()[i] = init;
    size_ = n;
  }
  void clear() { size_ = 0; }
  bool empty() const { return size_ == 0; }
  size_t size() const { return size_; }
  T* data() const { return ptr_.get(); }
  T* begin() const { return data(); }
  T* end() const { return data() + size_; }
  T& operator[](size_t i) const { return ptr_.get()[i]; }
  const std::shared_ptr<T>& ptr() const { return ptr_; }

 private:
  std::shared_ptr<T> ptr_;
  size_t size_ = 0, cap_ = 0;
};

}  // namespace difacto

namespace dmlc {
typedef float real_t;
/** the fields of dmlc::RowBlock<IndexType> (dmlc-core/include/dmlc/data.h:137-186) */
template <typename IndexType>
struct RowBlock {
  size_t size = 0;
  const size_t* offset = nullptr;
  const real_t* label = nullptr;
  const real_t* weight = nullptr;
  const IndexType* index = nullptr;
  const real_t* value = nullptr;
};
}  // namespace dmlc

namespace difacto {

/** owning CSR container (the role of dmlc::data::RowBlockContainer) */
template <typename IndexType>
struct RowBlockContainer {
  std::vector<size_t> offset{0};
  std::vector<real_t> label;
  std::vector<IndexType> index;
  std::vector<real_t> value;
  void Clear() { offset.assign(1, 0); label.clear(); index.clear(); value.clear(); }
  size_t Size() const { return offset.size() - 1; }
  dmlc::RowBlock<IndexType> GetBlock() const {
    dmlc::RowBlock<IndexType> b;
    b.size = Size();
    b.offset = offset.data();
    b.label = label.data();
    b.index = index.data();
    b.value = value.empty() ? nullptr : value.data();
    return b;
  }
};

/** include/difacto/loss.h:18-87 */
class Loss {
 public:
  static Loss* Create(const std::string& type, int nthreads = 2);
  virtual ~Loss() {}
  virtual KWArgs Init(const KWArgs& kwargs) = 0;
  /** param = {weights, w_pos, V_pos}; pred is accumulated into */
  virtual void Predict(const dmlc::RowBlock<unsigned>& data, const std::vector<SArray<char>>& param,
                       SArray<real_t>* pred) = 0;
  virtual real_t Evaluate(dmlc::real_t const* label, const SArray<real_t>& pred) const = 0;
  /** param = {weights, w_pos, V_pos, pred}; grad is accumulated into */
  virtual void CalcGrad(const dmlc::RowBlock<unsigned>& data, const std::vector<SArray<char>>& param,
                        SArray<real_t>* grad) = 0;
  void set_nthreads(int n) { nthreads_ = n; }   // kept for source compatibility; unused on the GPU
  int nthreads_ = 2;
};

/** include/difacto/updater.h:18-71 */
class Updater {
 public:
  virtual ~Updater() {}
  virtual KWArgs Init(const KWArgs& kwargs) = 0;
  virtual void Load(std::istream* fi, bool* has_aux) = 0;
  virtual void Save(bool save_aux, std::ostream* fo) const = 0;
  virtual void Get(const SArray<feaid_t>& fea_ids, int data_type, SArray<real_t>* data,
                   SArray<int>* data_offset) = 0;
  virtual void Update(const SArray<feaid_t>& fea_ids, int data_type, const SArray<real_t>& data,
                      const SArray<int>& data_offset) = 0;
};

/** include/difacto/store.h:19-104 */
class Store {
 public:
  static Store* Create();
  virtual ~Store() {}
  static const int kFeaCount = 1;
  static const int kWeight = 2;
  static const int kGradient = 3;
  virtual KWArgs Init(const KWArgs& kwargs) = 0;
  virtual int Push(const SArray<feaid_t>& fea_ids, int val_type, const SArray<real_t>& vals,
                   const SArray<int>& lens, const std::function<void()>& on_complete = nullptr) = 0;
  virtual int Pull(const SArray<feaid_t>& fea_ids, int val_type, SArray<real_t>* vals, SArray<int>* lens,
                   const std::function<void()>& on_complete = nullptr) = 0;
  virtual void Wait(int time) = 0;
  virtual int NumWorkers() = 0;
  virtual int NumServers() = 0;
  virtual int Rank() = 0;
  void SetUpdater(const std::shared_ptr<Updater>& updater) { updater_ = updater; }
  std::shared_ptr<Updater> updater() { return updater_; }

 protected:
  std::shared_ptr<Updater> updater_;
};

/** include/difacto/learner.h:20-73 (the Tracker indirection is collapsed: one process, one job queue) */
class Learner {
 public:
  static Learner* Create(const std::string& type);
  virtual ~Learner() {}
  virtual KWArgs Init(const KWArgs& kwargs) = 0;
  void Run() { RunScheduler(); }
  virtual void Stop() {}

 protected:
  virtual void RunScheduler() = 0;
};

}  // namespace difacto
