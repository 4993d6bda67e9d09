// difacto_b200/csrc/engine_internal.cuh -- the engine object behind a dfb_handle and the host-side helpers
// shared by engine.cu (C-ABI, single-GPU step) and shard.cu (the NVLink-sharded store).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/difacto_b200.h"
#include "dfb_internal.cuh"

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct ShardState;   // shard.cu

struct dfb_engine {
  dfb::Params prm;
  int device = 0;
  int compute_auc = 1;
  int force_generic = 0;
  int scatter_sorted = 1;   // 1: atomic-free sorted reduction (deterministic); 0: red.global atomics
  int overlap_auc = 1;      // run the AUC kernels on the auxiliary stream, concurrently with the update
  int has_aux = 1;          // false after restoring a snapshot saved without aux data (sgd_updater.h:88)
  int l2_hints = 1;         // per-load L2 eviction policies in the gather kernels
  int id_bits = 0;          // > 0: feature ids are < 2^id_bits (fixes the radix-sort bit range of the GPU localizer)
  int loc_begin_bit = -1;   // auto mode: lowest significant bit of the reversed keys seen so far (sticky)
  long long shard_timeout_ms = 20000;
  int k1_tma = 0;           // 1: validation batches use the bulk-copy staged gather kernel (kernels_fm_tma.cu)
  int long_row_nnz = 1024;  // rows with at least this many nonzeros are walked by a whole CTA (k_fm_long); 0: off
  int hot_split = 256;      // occurrence lists longer than this are pre-reduced in chunks by separate warps (0: off)
  cudaStream_t stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;
  cudaEvent_t ev_fm_done = nullptr, ev_auc_done = nullptr;
  dfb::Table tab;
  std::string err;
  std::vector<std::pair<std::string, std::string>> unknown;
  uint64_t launches = 0;

  // workspaces (grown on demand)
  DevBuf u_wv;
  DevBuf keys, cnt, slot, u_w, u_vrow, flags, pos, lens, cub, gw, gxxp, gV, pred, vals;
  DevBuf auc_k, auc_v, auc_tmp;
  DevBuf pxv, p_row, occ, occ_sorted, lidx_sorted, col_start, col_end;
  DevBuf hot_map, hot_info, hot_part, hot_ps, hot_cnt;
  DevBuf l_rkeys, l_skeys, l_pos, l_spos, l_head, l_rank, l_nnzrow, l_tmp;
  // outputs of the GPU localizer, double-buffered: the localizer of batch t+1 runs on loc_stream
  // while the step of batch t (which reads set t) runs on the main stream
  struct LocSet {
    DevBuf keys, lidx, cnt, occ_sorted, col_start, col_end;
    DevBuf scal;                          // {OR of all reversed keys, number of unique keys} (u64 x 2)
    cudaEvent_t done = nullptr, consumed = nullptr;
    bool used = false;
    const unsigned long long* dU() const { return scal.as<unsigned long long>() + 1; }
  } loc[2];
  uint64_t loc_seq = 0;
  cudaStream_t loc_stream = nullptr;
  unsigned long long* h_scal = nullptr;   // pinned: {or_all, n_unique} of a synchronous localize
  unsigned long long* h_or = nullptr;     // pinned ring [2]: or_all of earlier asynchronous localizes
  cudaEvent_t ev_or[2] = {nullptr, nullptr};
  bool or_pending[2] = {false, false};
  uint64_t or_seq = 0;
  DevBuf a_off, a_idx, a_val, a_lab, a_w, a_wpos, a_vpos, a_pred, a_grad;
  DevBuf scal, hasv, rV, rcg, nvals;
  // double-buffered inputs of the pipelined step
  struct InSet {
    DevBuf off, idx, val, lab, keys, cnt, ids;
    cudaEvent_t copied = nullptr, consumed = nullptr;
    const void* pre_ids = nullptr;     // host batch already staged by dfb_prefetch_raw (identity check)
    size_t pre_nrows = 0, pre_nnz = 0;
  } in[2];
  uint64_t seq = 0;
  // per-step Progress snapshots of the pipelined path (pinned ring + completion events)
  static constexpr int kRing = 8;
  dfb::DevProgress* h_ring = nullptr;
  cudaEvent_t ring_done[kRing] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t submitted = 0, collected = 0;
  dfb::DevProgress backlog;             // snapshots folded in when the ring was full
  // optional per-stage CUDA-event timing (bench.py's roofline numbers)
  // 0 lookup+pull, 1 fm, 2 auc, 3 csc sort, 4 update(+initv), 5 GPU localizer (raw batches),
  // sharded store: 6 worker slice+scatter, 7 owner lookup+partial sums, 8 worker reduce, 9 owner updates
  static constexpr int kStages = DFB_NUM_STAGES;
  static constexpr int kProfRing = 32;
  int profile = 0;
  std::vector<cudaEvent_t> pev;          // [kProfRing][kStages][2]
  std::vector<char> pev_used;            // [kProfRing][kStages]
  uint64_t prof_steps = 0;
  double stage_ms[kStages] = {};
  uint64_t stage_n[kStages] = {};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_join = nullptr;   // dfb_time_mark
  dfb::DevProgress* h_prog = nullptr;   // pinned
  unsigned long long* h_nvals = nullptr;  // pinned
  ShardState* shard = nullptr;          // the NVLink-sharded store this engine is a rank of (shard.cu)

  int fail(int code, const std::string& msg) { err = msg; return code; }
  int cuda_fail(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return DFB_ERR_CUDA;
  }
  // grow a workspace; frees/reallocs only after the streams that may still read it have drained
  int ensure(DevBuf& b, size_t bytes);
};

#define DFB_CUDA(h, call)                                                  \
  do {                                                                     \
    cudaError_t _e = (call);                                               \
    if (_e != cudaSuccess) return (h)->cuda_fail(_e, #call);               \
  } while (0)
#define DFB_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != 0) return _rc;      \
  } while (0)

// CUDA-event pair around one stage of a step (only while dfb_profile is on), on the stream the stage runs on
struct StageTimer {
  dfb_engine* h; int st; cudaStream_t s; cudaEvent_t* e = nullptr;
  StageTimer(dfb_engine* h_, int st_, cudaStream_t s_ = nullptr) : h(h_), st(st_), s(s_ ? s_ : h_->stream) {
    if (!h->profile) return;
    const int r = (int)(h->prof_steps % dfb_engine::kProfRing);
    e = &h->pev[(r * dfb_engine::kStages + st) * 2];
    cudaEventRecord(e[0], s);
  }
  void stop() {
    if (!e) return;
    cudaEventRecord(e[1], s);
    h->pev_used[((h->prof_steps % dfb_engine::kProfRing)) * dfb_engine::kStages + st] = 1;
    e = nullptr;
  }
  ~StageTimer() { stop(); }
};

// host-side helpers defined in engine.cu and used by shard.cu
namespace dfbh {
int prof_drain(dfb_engine* h);
// workspaces of the hot-key pre-reduction for a batch of nkeys keys / nnz non-zeros (hot_split > 0)
int hot_ws(dfb_engine* h, size_t nkeys, size_t nnz, dfb::HotWs* ws);
int join_streams(dfb_engine* h);        // make the main stream wait for everything enqueued on the others
int ensure_key_ws(dfb_engine* h, size_t n);
// Localizer::Compact on the device into L (no host synchronisation unless need_host_U):
//   *U_out = the unique-key count when need_host_U, else the capacity (nnz); L.dU() holds the count on the device
int localize_dev(dfb_engine* h, size_t nrows, size_t nnz, const uint64_t* d_off, const uint64_t* d_ids,
                 const float* d_val, uint64_t max_index, dfb_engine::LocSet& L, cudaStream_t s, bool need_host_U,
                 bool exact_range, size_t* U_out);
int check_csr(dfb_engine* h, size_t nrows, const uint64_t* offset);
int stage_raw(dfb_engine* h, dfb_engine::InSet& in, size_t nrows, size_t nnz, const uint64_t* offset,
              const uint64_t* ids, const float* value, const float* label);
int collect_one(dfb_engine* h, dfb::DevProgress* acc);
void shard_destroy(dfb_engine* h);      // shard.cu
int shard_sync(dfb_engine* h);          // shard.cu: drain the shard's streams
}  // namespace dfbh
