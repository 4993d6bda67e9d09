/*
 * oracle/ref_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" wrapper around the UNMODIFIED reference implementation
 * (dmlc/difacto @ /root/reference).  It is compiled by oracle/Makefile from the
 * reference sources where they lie (nothing is copied into this repo) into
 * oracle/_ref/libdifacto_ref.so.  It exists so that
 *   (1) the plain-C restatement in oracle/fm_oracle.c can be pinned against the
 *       real reference on arbitrary inputs,
 *   (2) golden fixtures under tests/golden/ can be generated, and
 *   (3) bench.py --impl reference / cpu_baseline can time the reference's own
 *       CPU SGD path on the GPU box's host cores.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load it.
 *
 * Reference entry points wrapped (file:line in /root/reference):
 *   ReverseBytes                 include/difacto/base.h:39-51
 *   BatchReader                  src/reader/batch_reader.cc:8-78
 *   Localizer::Compact           src/data/localizer.h:41-51, localizer.cc:11-103
 *   SGDUpdater::{Init,Get,Update} src/sgd/sgd_updater.cc:9-101
 *   FMLoss::{Predict,CalcGrad}   src/loss/fm_loss.h:67-119,148-199
 *   Loss::Evaluate               include/difacto/loss.h:57-66
 *   BinClassMetric::AUC          src/loss/bin_class_metric.h:35-56
 *   SGDLearner (whole learner)   src/sgd/sgd_learner.cc:31-273
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "difacto/base.h"
#include "difacto/loss.h"
#include "difacto/store.h"
#include "difacto/sarray.h"
#include "data/localizer.h"
#include "loss/fm_loss.h"
#include "loss/bin_class_metric.h"
#include "reader/batch_reader.h"
#include "sgd/sgd_updater.h"
#include "sgd/sgd_learner.h"
#include "dmlc/timer.h"

using namespace difacto;  // NOLINT

namespace {

KWArgs MakeKWArgs(const char* const* keys, const char* const* vals, int n) {
  KWArgs kw;
  for (int i = 0; i < n; ++i) kw.push_back(std::make_pair(keys[i], vals[i]));
  return kw;
}

struct RefBatch {
  dmlc::data::RowBlockContainer<feaid_t> blk;
};

/* restatement of the private SGDLearner::GetPos (sgd_learner.cc:113-127); it is
 * private in the reference so it cannot be called; 10 lines of integer scan. */
void GetPosLocal(const SArray<int>& len, SArray<int>* w_pos, SArray<int>* V_pos) {
  size_t n = len.size();
  w_pos->resize(n);
  V_pos->resize(n);
  int p = 0;
  for (size_t i = 0; i < n; ++i) {
    int l = len[i];
    (*w_pos)[i] = l == 0 ? -1 : p;
    (*V_pos)[i] = l > 1 ? p + 1 : -1;
    p += l;
  }
}

/* restatement of the private SGDLearner::EvaluatePenalty (sgd_learner.cc:249-273) */
real_t PenaltyLocal(const SGDUpdaterParam& param, const SArray<real_t>& weights,
                    const SArray<int>& w_pos, const SArray<int>& V_pos) {
  real_t objv = 0;
  if (w_pos.size()) {
    for (int p : w_pos) {
      if (p == -1) continue;
      real_t w = weights[p];
      objv += param.l1 * fabs(w) + .5 * param.l2 * w * w;
    }
    for (int p : V_pos) {
      if (p == -1) continue;
      for (int i = 0; i < param.V_dim; ++i) {
        real_t V = weights[p + i];
        objv += .5 * param.V_l2 * V * V;
      }
    }
  } else {
    for (auto w : weights) objv += param.l1 * fabs(w) + .5 * param.l2 * w * w;
  }
  return objv;
}

struct RefEngine {
  SGDUpdater updater;
  FMLoss loss;
  int nthreads = DEFAULT_NTHREADS;
};

}  // namespace

extern "C" {

uint64_t ref_reverse_bytes(uint64_t x) { return ReverseBytes(x); }

/* ---------------- data: BatchReader on a libsvm file ---------------- */

void* ref_read_batch(const char* path, const char* format, unsigned part,
                     unsigned nparts, unsigned batch_size, unsigned shuffle_buf,
                     float neg_sampling, unsigned which_batch) {
  BatchReader reader(path, format, part, nparts, batch_size, shuffle_buf, neg_sampling);
  for (unsigned b = 0; b <= which_batch; ++b) {
    if (!reader.Next()) return nullptr;
  }
  auto* out = new RefBatch();
  out->blk.Push(reader.Value());
  return out;
}
size_t ref_batch_rows(void* h) { return static_cast<RefBatch*>(h)->blk.offset.size() - 1; }
size_t ref_batch_nnz(void* h) { return static_cast<RefBatch*>(h)->blk.index.size(); }
int ref_batch_has_value(void* h) { return !static_cast<RefBatch*>(h)->blk.value.empty(); }
void ref_batch_copy(void* h, uint64_t* offset, float* label, uint64_t* index, float* value) {
  auto& b = static_cast<RefBatch*>(h)->blk;
  for (size_t i = 0; i < b.offset.size(); ++i) offset[i] = b.offset[i];
  memcpy(label, b.label.data(), b.label.size() * sizeof(float));
  memcpy(index, b.index.data(), b.index.size() * sizeof(uint64_t));
  if (value && !b.value.empty()) memcpy(value, b.value.data(), b.value.size() * sizeof(float));
}
void ref_batch_free(void* h) { delete static_cast<RefBatch*>(h); }

/* ---------------- Localizer::Compact ---------------- */
/* returns number of unique keys; out arrays must hold nnz entries */
size_t ref_localize(size_t nrows, const uint64_t* offset, const uint64_t* index,
                    const float* value_or_null, const float* label, uint64_t max_index,
                    int nthreads, uint32_t* out_index, uint64_t* out_offset,
                    uint64_t* out_keys, float* out_cnt_or_null) {
  static_assert(sizeof(size_t) == sizeof(uint64_t), "LP64 only");
  dmlc::RowBlock<feaid_t> blk;
  blk.size = nrows;
  blk.offset = reinterpret_cast<const size_t*>(offset);
  blk.label = label;
  blk.weight = nullptr;
  blk.index = index;
  blk.value = value_or_null;
  dmlc::data::RowBlockContainer<unsigned> compact;
  std::vector<feaid_t> uidx;
  std::vector<real_t> freq;
  Localizer lc(max_index, nthreads);
  lc.Compact(blk, &compact, &uidx, out_cnt_or_null ? &freq : nullptr);
  memcpy(out_index, compact.index.data(), compact.index.size() * sizeof(unsigned));
  for (size_t i = 0; i < compact.offset.size(); ++i) out_offset[i] = compact.offset[i];
  memcpy(out_keys, uidx.data(), uidx.size() * sizeof(feaid_t));
  if (out_cnt_or_null) memcpy(out_cnt_or_null, freq.data(), freq.size() * sizeof(real_t));
  return uidx.size();
}

/* ---------------- engine = SGDUpdater + FMLoss ---------------- */

void* ref_engine_create(const char* const* keys, const char* const* vals, int n, int nthreads) {
  auto* e = new RefEngine();
  KWArgs kw = MakeKWArgs(keys, vals, n);
  auto remain = e->updater.Init(kw);
  /* SGDLearner::Init re-injects V_dim for the loss (sgd_learner.cc:236) */
  remain.push_back(std::make_pair("V_dim", std::to_string(e->updater.param().V_dim)));
  e->loss.Init(remain);
  e->nthreads = nthreads;
  e->loss.set_nthreads(nthreads);
  return e;
}
void ref_engine_destroy(void* h) { delete static_cast<RefEngine*>(h); }

/* SGDUpdater::Get. vals_out must hold n*(1+V_dim); lens_out n. returns #vals, *nlens_out = lens.size() */
size_t ref_updater_get(void* h, const uint64_t* keys, size_t n, float* vals_out,
                       int* lens_out, size_t* nlens_out) {
  auto* e = static_cast<RefEngine*>(h);
  SArray<feaid_t> k(const_cast<uint64_t*>(keys), n);
  SArray<real_t> vals;
  SArray<int> lens;
  e->updater.Get(k, Store::kWeight, &vals, &lens);
  memcpy(vals_out, vals.data(), vals.size() * sizeof(real_t));
  memcpy(lens_out, lens.data(), lens.size() * sizeof(int));
  *nlens_out = lens.size();
  return vals.size();
}

/* SGDUpdater::Update; type: 1 = kFeaCount, 3 = kGradient */
void ref_updater_update(void* h, const uint64_t* keys, size_t n, int type,
                        const float* vals, size_t nvals, const int* lens, size_t nlens) {
  auto* e = static_cast<RefEngine*>(h);
  SArray<feaid_t> k(const_cast<uint64_t*>(keys), n);
  SArray<real_t> v(const_cast<float*>(vals), nvals);
  SArray<int> l(const_cast<int*>(lens), nlens);
  e->updater.Update(k, type, v, l);
}

static dmlc::RowBlock<unsigned> MakeBlock(size_t nrows, const uint64_t* offset,
                                          const uint32_t* index, const float* value,
                                          const float* label) {
  dmlc::RowBlock<unsigned> d;
  d.size = nrows;
  d.offset = reinterpret_cast<const size_t*>(offset);
  d.label = label;
  d.weight = nullptr;
  d.index = index;
  d.value = value;
  return d;
}

/* FMLoss::Predict (accumulates into pred; caller zero-fills) */
void ref_fm_predict(void* h, size_t nrows, const uint64_t* offset, const uint32_t* index,
                    const float* value_or_null, const float* label, const float* weights,
                    size_t nweights, const int* w_pos, const int* V_pos, size_t npos,
                    float* pred) {
  auto* e = static_cast<RefEngine*>(h);
  auto d = MakeBlock(nrows, offset, index, value_or_null, label);
  SArray<real_t> w(const_cast<float*>(weights), nweights);
  SArray<int> wp(const_cast<int*>(w_pos), w_pos ? npos : 0);
  SArray<int> vp(const_cast<int*>(V_pos), V_pos ? npos : 0);
  SArray<real_t> p(pred, nrows);
  e->loss.Predict(d, w, wp, vp, &p);
}

/* FMLoss::CalcGrad -- must follow ref_fm_predict on the same batch (XV_/XX_ member state) */
void ref_fm_calc_grad(void* h, size_t nrows, const uint64_t* offset, const uint32_t* index,
                      const float* value_or_null, const float* label, const float* weights,
                      size_t nweights, const int* w_pos, const int* V_pos, size_t npos,
                      const float* pred, float* grad) {
  auto* e = static_cast<RefEngine*>(h);
  auto d = MakeBlock(nrows, offset, index, value_or_null, label);
  SArray<real_t> w(const_cast<float*>(weights), nweights);
  SArray<int> wp(const_cast<int*>(w_pos), w_pos ? npos : 0);
  SArray<int> vp(const_cast<int*>(V_pos), V_pos ? npos : 0);
  SArray<real_t> p(const_cast<float*>(pred), nrows);
  SArray<real_t> g(grad, nweights);
  e->loss.CalcGrad(d, w, wp, vp, p, &g);
}

float ref_evaluate(void* h, const float* label, const float* pred, size_t n) {
  auto* e = static_cast<RefEngine*>(h);
  SArray<real_t> p(const_cast<float*>(pred), n);
  return e->loss.Evaluate(label, p);
}

float ref_auc(const float* label, const float* pred, size_t n) {
  BinClassMetric m(label, pred, n, 2);
  return m.AUC();
}

/* One minibatch of SGDLearner::IterateData (sgd_learner.cc:129-227) without file
 * I/O: Localizer::Compact -> [Update(kFeaCount)] -> Get -> GetPos -> Predict ->
 * Evaluate -> penalty -> AUC -> [CalcGrad -> Update(kGradient)].
 * progress[5] = {loss, penalty, auc, nnz_w(unused), nrows} is ACCUMULATED.
 * seconds[6] (optional) accumulates {localize, feacnt, get, predict, calcgrad, update}. */
void ref_sgd_step(void* h, size_t nrows, const uint64_t* offset, const uint64_t* index,
                  const float* value_or_null, const float* label, int is_train,
                  int push_cnt, float* progress, double* seconds) {
  auto* e = static_cast<RefEngine*>(h);
  double t0 = dmlc::GetTime();
  dmlc::RowBlock<feaid_t> blk;
  blk.size = nrows;
  blk.offset = reinterpret_cast<const size_t*>(offset);
  blk.label = label;
  blk.weight = nullptr;
  blk.index = index;
  blk.value = value_or_null;
  dmlc::data::RowBlockContainer<unsigned> data_c;
  auto feaids = std::make_shared<std::vector<feaid_t>>();
  auto feacnt = std::make_shared<std::vector<real_t>>();
  Localizer lc(-1, e->nthreads);
  lc.Compact(blk, &data_c, feaids.get(), push_cnt ? feacnt.get() : nullptr);
  SArray<feaid_t> keys(feaids);
  double t1 = dmlc::GetTime();
  if (push_cnt) {
    e->updater.Update(keys, Store::kFeaCount, SArray<real_t>(feacnt), SArray<int>());
  }
  double t2 = dmlc::GetTime();
  SArray<real_t> values;
  SArray<int> lengths;
  e->updater.Get(keys, Store::kWeight, &values, &lengths);
  double t3 = dmlc::GetTime();
  auto data = data_c.GetBlock();
  progress[4] += data.size;
  SArray<real_t> pred(data.size);
  SArray<int> w_pos, V_pos;
  GetPosLocal(lengths, &w_pos, &V_pos);
  e->loss.Predict(data, values, w_pos, V_pos, &pred);
  progress[0] += e->loss.Evaluate(data.label, pred);
  progress[1] += PenaltyLocal(e->updater.param(), values, w_pos, V_pos);
  BinClassMetric metric(data.label, pred.data(), pred.size(), e->nthreads);
  progress[2] += metric.AUC();
  double t4 = dmlc::GetTime(), t5 = t4, t6 = t4;
  if (is_train) {
    SArray<real_t> grads(values.size());
    e->loss.CalcGrad(data, values, w_pos, V_pos, pred, &grads);
    t5 = dmlc::GetTime();
    /* StoreLocal::Push copies then calls Update (store_local.h:24-34) */
    SArray<real_t> vals_copy; vals_copy.CopyFrom(grads);
    SArray<int> lens_copy; lens_copy.CopyFrom(lengths);
    e->updater.Update(keys, Store::kGradient, vals_copy, lens_copy);
    t6 = dmlc::GetTime();
  }
  if (seconds) {
    seconds[0] += t1 - t0; seconds[1] += t2 - t1; seconds[2] += t3 - t2;
    seconds[3] += t4 - t3; seconds[4] += t5 - t4; seconds[5] += t6 - t5;
  }
}

/* ---------------- the whole reference learner, for end-to-end goldens ---------------- */
/* runs SGDLearner with the given kwargs; records per-epoch {loss, penalty, auc, nnz_w, nrows}
 * for train then val into out[epoch*10 .. +10]. returns #epochs run. */
int ref_sgd_learner_run(const char* const* keys, const char* const* vals, int n,
                        float* out, int max_epochs) {
  SGDLearner learner;
  learner.Init(MakeKWArgs(keys, vals, n));
  int nrun = 0;
  learner.AddEpochEndCallback(
      [&](int epoch, const sgd::Progress& train, const sgd::Progress& val) {
        if (epoch >= max_epochs) return;
        float* o = out + epoch * 10;
        o[0] = train.loss; o[1] = train.penalty; o[2] = train.auc; o[3] = train.nnz_w; o[4] = train.nrows;
        o[5] = val.loss; o[6] = val.penalty; o[7] = val.auc; o[8] = val.nnz_w; o[9] = val.nrows;
        nrun = epoch + 1;
      });
  learner.Run();
  return nrun;
}

}  // extern "C"
