// difacto_b200/host/include/difacto_b200/sgd_learner.h -- SGDLearner over the GPU engine.
// Mirrors src/sgd/sgd_learner.{h,cc} + sgd_utils.h of the reference: epochs -> jobs (file parts) ->
// minibatches; per-epoch Progress, epoch-end callbacks and the two stop criteria are identical.
// The minibatch itself (the pull_callback of IterateData, sgd_learner.cc:138-177) is either
//   fused = 1 (default): one dfb_train_step_raw per batch on the reader's raw block (Localizer::Compact
//              on the device too), the model never leaves HBM;
//   fused = 2: host Localizer::Compact, then one dfb_train_step per batch;
//   fused = 0: the reference's own sequence of plugin calls Store::Pull -> GetPos -> Loss::Predict ->
//              Evaluate -> penalty -> AUC -> Loss::CalcGrad -> Store::Push through the adapter classes.
// num_gpus = N > 1 (with fused = 1): the worker/server split SGDLearner::RunEpoch was written for
// (sgd_learner.cc:78-89): N workers, each reading its own file parts (parsed on helper threads) and owning GPU
// r's shard of the model, one collective sharded step per round of minibatches, enqueued by one host thread in
// interleaved phases (dfb_shard_begin_async + dfb_shard_phase; the NVLink-sharded store, csrc/shard.cu).
#pragma once
#include <cmath>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "api.h"
#include "config.h"
#include "data.h"
#include "engine_adapters.h"

namespace difacto {
namespace sgd {

/** sgd::Job, src/sgd/sgd_utils.h:16-38 */
struct Job {
  static const int kLoadModel = 1, kSaveModel = 2, kTraining = 3, kValidation = 4, kEvaluation = 5;
  int type = kTraining, num_parts = 1, part_idx = 0, epoch = 0;
};

/** sgd::Progress, src/sgd/sgd_utils.h:40-75 */
struct Progress {
  real_t loss = 0, penalty = 0, auc = 0, nnz_w = 0, nrows = 0;
  std::string TextString() const {
    std::stringstream ss;
    ss << "loss = " << loss << ", AUC = " << auc / nrows;
    return ss.str();
  }
  void Merge(const Progress& o) {
    loss += o.loss; penalty += o.penalty; auc += o.auc; nnz_w += o.nnz_w; nrows += o.nrows;
  }
};

}  // namespace sgd

class SGDLearner : public Learner {
 public:
  typedef std::function<void(int epoch, const sgd::Progress& train, const sgd::Progress& val)> EpochCallback;
  SGDLearner() {}
  ~SGDLearner() override { delete loss_; delete store_; }
  KWArgs Init(const KWArgs& kwargs) override;
  void AddEpochEndCallback(const EpochCallback& cb) { epoch_end_callback_.push_back(cb); }
  GpuSGDUpdater* GetUpdater() { return static_cast<GpuSGDUpdater*>(store_->updater().get()); }
  /** task=predict (a `LOG(FATAL) << "TODO"` in the reference, main.cc:61-62): load model_in, run the forward pass
   * over data_in in file order and write one prediction (the clamped FM score, fm_loss.h:118) per row to pred_out;
   * returns the Progress (loss, AUC*n, nrows) of the pass */
  sgd::Progress Predict();
  /** set false to silence the per-epoch log lines */
  bool verbose = true;

 protected:
  void RunScheduler() override;

 private:
  void RunEpochs();
  void RunEpoch(int epoch, int job_type, sgd::Progress* prog);
  void IterateData(const sgd::Job& job, sgd::Progress* prog);
  /** num_gpus > 1: one epoch over all N * num_jobs_per_epoch file parts, part i handled by worker i % N */
  void RunEpochSharded(int epoch, int job_type, sgd::Progress* prog);
  void LoadShards();
  /** engine of GPU r (r = 0: the updater's own) */
  const std::shared_ptr<GpuEngine>& ShardEngine(int r) { return r == 0 ? GetUpdater()->engine() : shard_engines_[r - 1]; }
  void BatchFused(const RowBlockContainer<unsigned>& data, const std::vector<feaid_t>& keys,
                  const std::vector<real_t>* cnt, bool train, sgd::Progress* prog);
  void BatchPluginCalls(const RowBlockContainer<unsigned>& data, const std::vector<feaid_t>& keys,
                        const std::vector<real_t>* cnt, bool train, sgd::Progress* prog);
  static void GetPos(const SArray<int>& len, SArray<int>* w_pos, SArray<int>* V_pos);
  real_t EvaluatePenalty(const SArray<real_t>& weights, const SArray<int>& w_pos, const SArray<int>& V_pos);

  Store* store_ = nullptr;
  Loss* loss_ = nullptr;
  SGDLearnerParam param_;
  std::vector<EpochCallback> epoch_end_callback_;
  std::vector<std::shared_ptr<GpuEngine>> shard_engines_;     // GPUs 1..N-1 (num_gpus > 1)
};

}  // namespace difacto
