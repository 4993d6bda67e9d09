// difacto_b200/host/src/sgd_learner.cc -- see sgd_learner.h.  Control flow follows
// src/sgd/sgd_learner.cc of the reference (cited per function); the arithmetic lives on the GPU.
#include "difacto_b200/sgd_learner.h"

#include <cstdio>
#include <fstream>
#include <thread>

namespace difacto {

Loss* Loss::Create(const std::string& type, int nthreads) {   // src/loss/loss.cc:13-26
  if (type != "fm") throw Error("unknown loss type: " + type + " (the B200 engine provides \"fm\")");
  Loss* l = new GpuFMLoss();
  l->set_nthreads(nthreads);
  return l;
}

Store* Store::Create() { return new GpuStore(); }   // src/store/store.cc:8-15

Learner* Learner::Create(const std::string& type) {   // src/learner.cc:15-26
  if (type == "sgd") return new SGDLearner();
  throw Error("unknown learner type: " + type + " (the B200 engine accelerates the sgd learner)");
}

// SGDLearner::Init, sgd_learner.cc:229-246
KWArgs SGDLearner::Init(const KWArgs& kwargs) {
  KWArgs remain = param_.InitAllowUnknown(kwargs);
  auto* updater = new GpuSGDUpdater();
  std::shared_ptr<Updater> holder(updater);
  remain = updater->Init(remain);
  remain.push_back(std::make_pair("V_dim", std::to_string(updater->param().V_dim)));   // :236
  store_ = Store::Create();
  store_->SetUpdater(holder);
  remain = store_->Init(remain);
  loss_ = Loss::Create(param_.loss, 2);
  static_cast<GpuFMLoss*>(loss_)->AttachEngine(updater->engine());
  remain = loss_->Init(remain);
  if (param_.num_gpus > 1) {
    // one engine (= one shard of the model) per GPU, the same hyper-parameters everywhere; wire the mailboxes
    if (param_.fused != 1) throw ParamError("num_gpus > 1 needs fused = 1 (raw blocks through the sharded store)");
    if (param_.num_gpus > 8) throw ParamError("num_gpus must be <= 8");
    for (int r = 1; r < param_.num_gpus; ++r) {
      KWArgs kw;
      for (const auto& kv : kwargs) if (kv.first != "device") kw.push_back(kv);
      kw.push_back(std::make_pair("device", std::to_string(r)));
      KWArgs engine_kw;     // only what dfb_create knows: drop the learner's keys (it would report them as unknown)
      SGDLearnerParam tmp;
      engine_kw = tmp.InitAllowUnknown(kw);
      shard_engines_.push_back(std::make_shared<GpuEngine>(engine_kw));
    }
    const int N = param_.num_gpus;
    const size_t max_rows = static_cast<size_t>(std::max(param_.batch_size, 65536));     // validation reads 65536-row chunks
    const size_t max_nnz = param_.shard_max_nnz > 0 ? static_cast<size_t>(param_.shard_max_nnz) : max_rows * 256;
    std::vector<void*> boxes(N, nullptr);
    for (int r = 0; r < N; ++r) {
      const auto& e = ShardEngine(r);
      e->Check(dfb_shard_init(e->handle(), r, N, max_rows, max_nnz, 0, 0, nullptr), "dfb_shard_init");
      unsigned char ipc_handle[64];      // unused here (same process); other processes would open it with dfb_peer_open
      e->Check(dfb_shard_export(e->handle(), &boxes[r], getenv("DFB_NO_IPC_EXPORT") ? nullptr : ipc_handle), "dfb_shard_export");
    }
    for (int r = 0; r < N; ++r) ShardEngine(r)->Check(dfb_shard_connect(ShardEngine(r)->handle(), boxes.data()), "dfb_shard_connect");
  }
  return remain;
}

// model_in -> the engine(s): one file, or <name>_part-<r> per shard when num_gpus = N > 1
void SGDLearner::LoadShards() {
  for (int r = 0; r < param_.num_gpus; ++r) {
    const std::string fn = param_.num_gpus > 1 ? param_.model_in + "_part-" + std::to_string(r) : param_.model_in;
    std::ifstream fi(fn, std::ios::binary);
    if (!fi) throw Error("failed to open model_in " + fn);
    std::string blob((std::istreambuf_iterator<char>(fi)), std::istreambuf_iterator<char>());
    int aux = 0;
    ShardEngine(r)->Check(dfb_restore(ShardEngine(r)->handle(), blob.data(), blob.size(), &aux), "dfb_restore");
    if (verbose) printf("Loaded model from %s (%s aux data)\n", fn.c_str(), aux ? "with" : "without");
  }
}

// SGDLearner::RunScheduler, sgd_learner.cc:31-68 (+ model_in / model_out, declared in
// sgd_param.h:53-54 but never acted on by the reference: Job::kLoadModel/kSaveModel are not issued)
void SGDLearner::RunScheduler() {
  // num_gpus = N > 1: one snapshot per shard, <name>_part-<r> (the format of Updater::Save per shard)
  auto shard_name = [&](const std::string& base, int r) { return param_.num_gpus > 1 ? base + "_part-" + std::to_string(r) : base; };
  if (!param_.model_in.empty()) LoadShards();
  RunEpochs();
  if (!param_.model_out.empty()) {
    for (int r = 0; r < param_.num_gpus; ++r) {
      const std::string fn = shard_name(param_.model_out, r);
      std::ofstream fo(fn, std::ios::binary);
      if (!fo) throw Error("failed to open model_out " + fn);
      size_t n = 0;
      ShardEngine(r)->Check(dfb_snapshot_size(ShardEngine(r)->handle(), 1, &n), "dfb_snapshot_size");
      std::string blob(n, '\0');
      ShardEngine(r)->Check(dfb_snapshot(ShardEngine(r)->handle(), 1, &blob[0], n), "dfb_snapshot");
      fo.write(blob.data(), static_cast<std::streamsize>(n));
      if (verbose) printf("Saved model to %s\n", fn.c_str());
    }
  }
}

sgd::Progress SGDLearner::Predict() {
  if (param_.model_in.empty()) throw Error("task=predict needs model_in");
  if (param_.num_gpus > 1) {
    // the shards of a num_gpus = N run, scored by the same N GPUs (loss / AUC only: the rows of a step are spread
    // over the workers, so there is no single prediction file to write)
    if (!param_.pred_out.empty()) throw Error("pred_out needs num_gpus=1 (the predictions of a sharded run are spread over the workers)");
    LoadShards();
    sgd::Progress prog;
    param_.data_val = param_.data_in;
    RunEpochSharded(0, sgd::Job::kValidation, &prog);
    if (verbose) printf(" - Prediction: %s, rows = %g\n", prog.TextString().c_str(), prog.nrows);
    return prog;
  }
  std::ifstream fi(param_.model_in, std::ios::binary);
  if (!fi) throw Error("failed to open model_in " + param_.model_in);
  bool has_aux = false;
  GetUpdater()->Load(&fi, &has_aux);
  FILE* fo = param_.pred_out.empty() ? nullptr : fopen(param_.pred_out.c_str(), "w");
  if (!param_.pred_out.empty() && !fo) throw Error("failed to open pred_out " + param_.pred_out);
  sgd::Progress prog;
  const auto& eng = GetUpdater()->engine();
  BatchReader reader(param_.data_in, param_.data_format, 0, 1, 65536u);
  std::vector<real_t> pred;
  while (reader.Next()) {
    const auto blk = reader.Value();
    pred.assign(blk.size, 0.f);
    dfb_progress pr;
    eng->Check(dfb_train_step_raw(eng->handle(), blk.size, reinterpret_cast<const uint64_t*>(blk.offset), blk.index,
                                  blk.value, blk.label, 0, 0, &pr, pred.data()), "dfb_train_step_raw");
    prog.loss += pr.loss; prog.penalty += pr.penalty; prog.auc += pr.auc; prog.nrows += pr.nrows;
    if (fo) for (real_t p : pred) fprintf(fo, "%.9g\n", p);
  }
  if (fo) fclose(fo);
  if (verbose) printf(" - Prediction: %s, rows = %g\n", prog.TextString().c_str(), prog.nrows);
  return prog;
}

void SGDLearner::RunEpochs() {
  real_t pre_loss = 0, pre_val_auc = 0;
  for (int k = 0; k < param_.max_num_epochs; ++k) {
    sgd::Progress train_prog, val_prog;
    if (verbose) printf("Start epoch %d\n", k);
    RunEpoch(k, sgd::Job::kTraining, &train_prog);
    if (verbose) printf(" - Training: %s\n", train_prog.TextString().c_str());
    if (!param_.data_val.empty()) {
      RunEpoch(k, sgd::Job::kValidation, &val_prog);
      if (verbose) printf(" - Validation: %s\n", val_prog.TextString().c_str());
    }
    for (const auto& cb : epoch_end_callback_) cb(k, train_prog, val_prog);
    real_t eps = std::fabs(train_prog.loss - pre_loss) / pre_loss;
    if (eps < param_.stop_rel_objv) {
      if (verbose) printf("Change of loss [%g] < stop_rel_objv [%g]\n", eps, param_.stop_rel_objv);
      break;
    }
    if (val_prog.auc > 0) {
      eps = (val_prog.auc - pre_val_auc) / val_prog.nrows;
      if (eps < param_.stop_val_auc) {
        if (verbose) printf("Change of validation AUC [%g] < stop_val_auc [%g]\n", eps, param_.stop_val_auc);
        break;
      }
    }
    if (k + 1 >= param_.max_num_epochs && verbose) printf("Reach maximal number of epochs\n");
    pre_loss = train_prog.loss;
    pre_val_auc = val_prog.auc;
    if (verbose) fflush(stdout);
  }
}

// SGDLearner::RunEpoch, sgd_learner.cc:70-111: n = NumWorkers * num_jobs_per_epoch file parts,
// executed in order (the reference's LocalTracker runs them serially on one thread too)
void SGDLearner::RunEpoch(int epoch, int job_type, sgd::Progress* prog) {
  if (param_.num_gpus > 1) { RunEpochSharded(epoch, job_type, prog); return; }
  const int n = store_->NumWorkers() * param_.num_jobs_per_epoch;
  for (int i = 0; i < n; ++i) {
    sgd::Job job;
    job.type = job_type; job.epoch = epoch; job.num_parts = n; job.part_idx = i;
    sgd::Progress p;
    IterateData(job, &p);
    prog->Merge(p);
  }
}

// num_gpus = N > 1: the epoch of sgd_learner.cc:70-111 with N workers.  File part i of the N * num_jobs_per_epoch
// parts goes to worker i % N (the reference's tracker hands parts to whichever worker is free; the assignment is fixed
// here so that runs are reproducible).  Every round each worker contributes its next minibatch (or an empty one when
// its parts are exhausted) to ONE collective sharded step.  One host thread drives all N engines and interleaves
// the five enqueue phases of the step (dfb_shard_begin_async + dfb_shard_phase): nothing in a step waits for
// the host, so the GPUs run concurrently; the readers parse their next batches on helper threads meanwhile.
void SGDLearner::RunEpochSharded(int epoch, int job_type, sgd::Progress* prog) {
  const int N = param_.num_gpus;
  const int nparts = N * param_.num_jobs_per_epoch;
  const bool train = job_type == sgd::Job::kTraining;
  const bool push_cnt = train && epoch == 0;      // sgd_learner.cc:201-202
  struct Worker {
    int part = 0;
    std::unique_ptr<BatchReader> reader;
    bool has = false;
    std::string error;
  };
  std::vector<Worker> w(N);
  for (int r = 0; r < N; ++r) w[r].part = r;
  auto next = [&](int r) {          // host only (file parsing): runs on a helper thread per worker
    Worker& me = w[r];
    me.has = false;
    try {
      for (;;) {
        if (!me.reader) {
          if (me.part >= nparts) return;
          me.reader.reset(new BatchReader(train ? param_.data_in : param_.data_val, param_.data_format,
                                          static_cast<unsigned>(me.part), static_cast<unsigned>(nparts),
                                          train ? static_cast<unsigned>(param_.batch_size) : 65536u,
                                          train ? static_cast<unsigned>(param_.batch_size) * static_cast<unsigned>(param_.shuffle) : 0u,
                                          train ? param_.neg_sampling : 1.0f, static_cast<unsigned>(epoch)));
        }
        if (me.reader->Next()) { me.has = true; return; }
        me.reader.reset();
        me.part += N;
      }
    } catch (const std::exception& e) {
      me.error = e.what();
    }
  };
  for (;;) {
    {
      std::vector<std::thread> th;
      for (int r = 0; r < N; ++r) th.emplace_back(next, r);
      for (auto& t : th) t.join();
    }
    bool any = false;
    for (int r = 0; r < N; ++r) {
      if (!w[r].error.empty()) throw Error("worker " + std::to_string(r) + ": " + w[r].error);
      any = any || w[r].has;
    }
    if (!any) break;
    for (int r = 0; r < N; ++r) {
      const auto& eng = ShardEngine(r);
      if (w[r].has) {
        const auto blk = w[r].reader->Value();
        eng->Check(dfb_shard_begin_async(eng->handle(), blk.size, reinterpret_cast<const uint64_t*>(blk.offset), blk.index,
                                         blk.value, blk.label, push_cnt ? 1 : 0, train ? 1 : 0), "dfb_shard_begin_async");
      } else {
        eng->Check(dfb_shard_begin_async(eng->handle(), 0, nullptr, nullptr, nullptr, nullptr, push_cnt ? 1 : 0, train ? 1 : 0),
                   "dfb_shard_begin_async");
      }
    }
    for (int ph = 0; ph < 5; ++ph)
      for (int r = 0; r < N; ++r) ShardEngine(r)->Check(dfb_shard_phase(ShardEngine(r)->handle(), ph), "dfb_shard_phase");
    for (int r = 0; r < N; ++r) {
      dfb_progress pr;
      ShardEngine(r)->Check(dfb_wait_step(ShardEngine(r)->handle(), &pr), "dfb_wait_step");
      prog->loss += pr.loss; prog->penalty += pr.penalty; prog->auc += pr.auc; prog->nrows += pr.nrows;
    }
  }
}

// SGDLearner::GetPos, sgd_learner.cc:113-127
void SGDLearner::GetPos(const SArray<int>& len, SArray<int>* w_pos, SArray<int>* V_pos) {
  const size_t n = len.size();
  w_pos->resize(n);
  V_pos->resize(n);
  int p = 0;
  for (size_t i = 0; i < n; ++i) {
    const int l = len[i];
    (*w_pos)[i] = l == 0 ? -1 : p;
    (*V_pos)[i] = l > 1 ? p + 1 : -1;
    p += l;
  }
}

// SGDLearner::EvaluatePenalty, sgd_learner.cc:249-273
real_t SGDLearner::EvaluatePenalty(const SArray<real_t>& weights, const SArray<int>& w_pos,
                                   const SArray<int>& V_pos) {
  const auto& param = GetUpdater()->param();
  real_t objv = 0;
  if (w_pos.size()) {
    for (int p : w_pos) {
      if (p == -1) continue;
      const real_t w = weights[p];
      objv += param.l1 * std::fabs(w) + .5 * param.l2 * w * w;
    }
    for (int p : V_pos) {
      if (p == -1) continue;
      for (int i = 0; i < param.V_dim; ++i) {
        const real_t V = weights[p + i];
        objv += .5 * param.V_l2 * V * V;
      }
    }
  } else {
    for (real_t w : weights) objv += param.l1 * std::fabs(w) + .5 * param.l2 * w * w;
  }
  return objv;
}

// the pull_callback of IterateData (sgd_learner.cc:138-177) as ONE fused device step
void SGDLearner::BatchFused(const RowBlockContainer<unsigned>& data, const std::vector<feaid_t>& keys,
                            const std::vector<real_t>* cnt, bool train, sgd::Progress* prog) {
  const auto& eng = GetUpdater()->engine();
  dfb_progress pr;
  eng->Check(dfb_train_step(eng->handle(), data.Size(), reinterpret_cast<const uint64_t*>(data.offset.data()),
                            data.index.data(), data.value.empty() ? nullptr : data.value.data(), data.label.data(),
                            keys.data(), keys.size(), cnt ? cnt->data() : nullptr, train ? 1 : 0, &pr, nullptr),
             "dfb_train_step");
  prog->loss += pr.loss; prog->penalty += pr.penalty; prog->auc += pr.auc; prog->nrows += pr.nrows;
}

// the same minibatch through the reference's own plugin calls (sgd_learner.cc:138-177, :214-217)
void SGDLearner::BatchPluginCalls(const RowBlockContainer<unsigned>& data_c, const std::vector<feaid_t>& keys,
                                  const std::vector<real_t>* cnt, bool train, sgd::Progress* prog) {
  SArray<feaid_t> feaids(const_cast<feaid_t*>(keys.data()), keys.size());
  if (cnt) {
    SArray<real_t> feacnt(const_cast<real_t*>(cnt->data()), cnt->size());
    store_->Wait(store_->Push(feaids, Store::kFeaCount, feacnt, SArray<int>()));
  }
  SArray<real_t> values;
  SArray<int> lengths;
  store_->Pull(feaids, Store::kWeight, &values, &lengths);
  auto data = data_c.GetBlock();
  prog->nrows += data.size;
  SArray<real_t> pred(data.size);
  SArray<int> w_pos, V_pos;
  GetPos(lengths, &w_pos, &V_pos);
  std::vector<SArray<char>> inputs = {SArray<char>(values), SArray<char>(w_pos), SArray<char>(V_pos)};
  loss_->Predict(data, inputs, &pred);
  prog->loss += loss_->Evaluate(data.label, pred);
  prog->penalty += EvaluatePenalty(values, w_pos, V_pos);
  prog->auc += static_cast<GpuFMLoss*>(loss_)->AUC(data.label, pred);
  if (train) {
    SArray<real_t> grads(values.size());
    inputs.push_back(SArray<char>(pred));
    loss_->CalcGrad(data, inputs, &grads);
    store_->Push(feaids, Store::kGradient, grads, lengths);
  }
}

// SGDLearner::IterateData, sgd_learner.cc:129-227
void SGDLearner::IterateData(const sgd::Job& job, sgd::Progress* progress) {
  const bool train = job.type == sgd::Job::kTraining;
  // training: BatchReader(batch_size, shuffle window = batch_size*shuffle, neg_sampling) (:182-188);
  // validation: plain reader, here in chunks of 65536 rows (:190-194 reads 256 MB chunks)
  BatchReader reader(train ? param_.data_in : param_.data_val, param_.data_format,
                     static_cast<unsigned>(job.part_idx), static_cast<unsigned>(job.num_parts),
                     train ? static_cast<unsigned>(param_.batch_size) : 65536u,
                     train ? static_cast<unsigned>(param_.batch_size) * static_cast<unsigned>(param_.shuffle) : 0u,
                     train ? param_.neg_sampling : 1.0f, static_cast<unsigned>(job.epoch), 0,
                     ShuffleOrder::kReference);      // one reader at a time: the reference's own shuffle order
  while (reader.Next()) {
    const bool push_cnt = train && job.epoch == 0;   // :201-202
    if (param_.fused == 1) {
      // Localizer::Compact + [kFeaCount push] + Pull/Predict/.../Push in one device call on the raw block
      const auto blk = reader.Value();
      const auto& eng = GetUpdater()->engine();
      dfb_progress pr;
      eng->Check(dfb_train_step_raw(eng->handle(), blk.size, reinterpret_cast<const uint64_t*>(blk.offset),
                                    blk.index, blk.value, blk.label, push_cnt ? 1 : 0, train ? 1 : 0, &pr, nullptr),
                 "dfb_train_step_raw");
      progress->loss += pr.loss; progress->penalty += pr.penalty; progress->auc += pr.auc; progress->nrows += pr.nrows;
      continue;
    }
    RowBlockContainer<unsigned> data;
    std::vector<feaid_t> feaids;
    std::vector<real_t> feacnt;
    Localizer lc(-1, 2);
    lc.Compact(reader.Value(), &data, &feaids, push_cnt ? &feacnt : nullptr);
    if (param_.fused == 2) BatchFused(data, feaids, push_cnt ? &feacnt : nullptr, train, progress);
    else BatchPluginCalls(data, feaids, push_cnt ? &feacnt : nullptr, train, progress);
  }
}

}  // namespace difacto
