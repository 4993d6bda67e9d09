/*
 * include/difacto_b200.h -- the drop-in boundary of difacto-b200.
 *
 * A C-ABI (plain pointers and sizes, no C++/torch types) over the B200-native
 * FM-SGD engine.  Each entry point replaces one interface of the reference
 * (dmlc/difacto @ 78e3562; paths below are relative to that tree).  The reference
 * itself has no C ABI on this path -- it has four C++ abstract classes
 * (include/difacto/{loss,updater,store,learner}.h); INTEGRATION.md shows the
 * adapter classes a maintainer adds there to bind these symbols.
 *
 * Conventions
 *  - every function returns 0 on success or a negative dfb_status; the message is
 *    available from dfb_last_error(h).  Nothing aborts the process (the reference
 *    CHECK()s -> abort(), Makefile:13 -DDMLC_LOG_FATAL_THROW=0; the C++ adapter maps a
 *    non-zero code back to LOG(FATAL)).
 *  - a handle owns one CUDA device, one compute stream and one copy stream.  It is
 *    thread-compatible: at most one call in flight per handle (the reference's Loss is
 *    not re-entrant either: fm_loss.h:202-203 member state).
 *  - "host" pointers are ordinary host memory (pinned memory from dfb_host_alloc makes
 *    the copies asynchronous); "dev" pointers are device memory on the handle's device.
 *  - feature keys are the reference's *reversed* feature ids (ReverseBytes,
 *    include/difacto/base.h:39-51), sorted ascending and unique, exactly what
 *    Localizer::Compact emits (src/data/localizer.cc:36-49).
 *  - all floating point is fp32 (real_t, include/difacto/base.h:16).
 */
#ifndef DIFACTO_B200_H_
#define DIFACTO_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfb_engine* dfb_handle;

typedef enum {
  DFB_OK = 0,
  DFB_ERR_INVALID = -1,   /* bad argument / a CHECK of the reference would have fired */
  DFB_ERR_CUDA = -2,      /* CUDA runtime error */
  DFB_ERR_CAPACITY = -3,  /* key table or V-row pool exhausted */
  DFB_ERR_PARAM = -4,     /* parameter missing / out of range (dmlc::ParamError in the reference) */
  DFB_ERR_NCCL = -5,
  DFB_ERR_TIMEOUT = -6    /* a rank of the sharded store did not show up within shard_timeout_ms */
} dfb_status;

/* value types of Store::Push / Pull: include/difacto/store.h:33-35 */
enum { DFB_FEA_COUNT = 1, DFB_WEIGHT = 2, DFB_GRADIENT = 3 };

/* sgd::Progress, src/sgd/sgd_utils.h:40-45 (+ device-side counters) */
typedef struct {
  float loss;      /* sum_i log(1+exp(-y_i pred_i))            loss.h:57-66            */
  float penalty;   /* l1|w| + .5 l2 w^2 + .5 V_l2 V^2 (pulled)  sgd_learner.cc:249-273  */
  float auc;       /* AUC * nrows                               bin_class_metric.h:35-56 */
  float nnz_w;     /* unused by the SGD learner (kept for layout compatibility)          */
  float nrows;
  /* extras (not in the reference) */
  uint64_t new_keys;    /* keys inserted by this step                 */
  uint64_t new_vrows;   /* V rows allocated (InitV) by this step      */
} dfb_progress;

/* ---------------------------------------------------------------------------------
 * life cycle.  Replaces SGDUpdater::Init (src/sgd/sgd_updater.cc:9-11) + FMLoss::Init
 * (src/loss/fm_loss.h:37-39) + Store::Init.  Consumes the reference's SGDUpdaterParam
 * keys with the same names, defaults and ranges (src/sgd/sgd_param.h:66-107):
 *   l1 l2 V_l2 lr lr_beta V_lr V_lr_beta V_init_scale V_dim V_threshold seed
 * plus engine-only keys (unknown to the reference, ignored by it with a warning):
 *   device (0)            CUDA device ordinal
 *   table_capacity (1<<20) max number of distinct keys this shard can hold
 *   V_capacity (=table_capacity) max number of keys with an allocated V row
 *   compute_auc (1)       evaluate AUC on device each step (sgd_learner.cc:150-153)
 *   scatter (sorted)      sorted = gradient reduced per key through a stable radix-sorted CSC view of
 *                         the batch, fused with the FTRL/AdaGrad step (no atomics, bit-reproducible,
 *                         row order per key like SpMM::TransTimes); atomic = fp32 red.global scatter
 *                         into dense gradient rows followed by a separate update kernel
 *   overlap_auc (1)       run the AUC kernels on an auxiliary stream, overlapped with the update
 *   l2_hints (1)          per-load L2 eviction policies in the gather kernels (V rows evict_first, the per-nnz
 *                         {w, row index} view evict_last)
 *   id_bits (0)           > 0: feature ids are < 2^id_bits; fixes the bit range the GPU localizer's radix sort
 *                         visits.  0: the range is learned from the first raw batch (one 8-byte sync) and
 *                         verified on the device afterwards; a later batch with wider ids is dropped with
 *                         DFB_ERR_INVALID (pass id_bits, or use the synchronous dfb_train_step_raw)
 *   shard_timeout_ms (20000)  how long a rank of the sharded store waits for a peer
 *   hot_split (256)       a key with more occurrences in a minibatch than this has its gradient pre-reduced in
 *                         chunks by separate warps, then finished in chunk order (0: one warp per key)
 *   long_row_nnz (1024)   an example with at least this many nonzeros is walked by a whole CTA in the forward
 *                         kernel, its chunks added in a fixed order (0: one warp per example whatever its length)
 *   force_generic (0)     1 = use the any-V_dim kernels even where a specialised one exists (tests)
 * Keys that are neither are returned through dfb_unknown_kwarg, mirroring the
 * "return the unconsumed kwargs" convention (updater.h:34, main.cc:25-31).
 * ------------------------------------------------------------------------------- */
int dfb_create(const char* const* keys, const char* const* vals, int n, dfb_handle* out);
int dfb_destroy(dfb_handle h);
const char* dfb_last_error(dfb_handle h);   /* h may be NULL: error of the failed dfb_create */
int dfb_num_unknown_kwargs(dfb_handle h);
int dfb_unknown_kwarg(dfb_handle h, int i, const char** key, const char** val);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
uint64_t dfb_launch_count(dfb_handle h);
/* number of keys / V rows currently stored */
int dfb_table_stats(dfb_handle h, uint64_t* n_keys, uint64_t* n_vrows, uint64_t* capacity,
                    uint64_t* v_capacity);

/* pinned host memory so that H2D/D2H copies are asynchronous DMA */
int dfb_host_alloc(void** ptr, size_t bytes);
int dfb_host_free(void* ptr);

/* ---------------------------------------------------------------------------------
 * (A) API-faithful path: host buffers in, host buffers out; one call = one reference
 * call.  Used by the adapter classes GpuStore/GpuSGDUpdater/GpuFMLoss.
 * ------------------------------------------------------------------------------- */

/* Store::Push(keys, kFeaCount, cnt)  store_local.h:24-34 -> SGDUpdater::Update
 * sgd_updater.cc:62-73: fea_cnt += cnt; allocate V when w!=0 && fea_cnt>V_threshold. */
int dfb_push_feacnt(dfb_handle h, const uint64_t* keys, size_t n, const float* cnt);

/* Store::Pull(keys, kWeight, &vals, &lens)  store_local.h:36-44 -> SGDUpdater::Get
 * sgd_updater.cc:32-56.  vals_out (capacity vals_cap floats, n*(1+V_dim) always suffices)
 * receives the ragged [w_i, V_i...] rows; lens_out[n] receives 1 or 1+V_dim;
 * *nlens = 0 when V_dim == 0 (sgd_updater.cc:40), n otherwise. */
int dfb_pull(dfb_handle h, const uint64_t* keys, size_t n, float* vals_out, size_t vals_cap,
             int* lens_out, size_t* nvals, size_t* nlens);

/* Store::Push(keys, kGradient, grads, lens) -> SGDUpdater::Update sgd_updater.cc:74-101:
 * FTRL on w (UpdateW :104-127), AdaGrad on V (UpdateV :129-138), lazy InitV (:140-147).
 * nlens == 0 means w-only (lens.empty()). */
int dfb_push_grad(dfb_handle h, const uint64_t* keys, size_t n, const float* grads,
                  size_t nvals, const int* lens, size_t nlens);

/* FMLoss::Predict  src/loss/fm_loss.h:67-119.  CSR<u32> data (dmlc::RowBlock<unsigned>:
 * offset[nrows+1] of size_t, index, optional value), weights in the pulled ragged layout
 * with w_pos/V_pos (-1 = absent; both NULL = direct indexing, the V_dim==0 path).
 * pred[nrows] is ACCUMULATED into, as in the reference (caller zero-fills). */
int dfb_predict(dfb_handle h, size_t nrows, const uint64_t* offset, const uint32_t* index,
                const float* value_or_null, const float* weights, size_t nweights,
                const int* w_pos, const int* V_pos, size_t npos, float* pred);

/* FMLoss::CalcGrad  src/loss/fm_loss.h:148-199.  grad[nweights] (same ragged layout as
 * weights) is ACCUMULATED into.  Unlike the reference it does not depend on member state
 * left behind by Predict (XV_ is recomputed on device). */
int dfb_calc_grad(dfb_handle h, size_t nrows, const uint64_t* offset, const uint32_t* index,
                  const float* value_or_null, const float* label, const float* weights,
                  size_t nweights, const int* w_pos, const int* V_pos, size_t npos,
                  const float* pred, float* grad);

/* Loss::Evaluate  include/difacto/loss.h:57-66 */
int dfb_evaluate(dfb_handle h, const float* label, const float* pred, size_t n, float* objv);

/* BinClassMetric::AUC  src/loss/bin_class_metric.h:35-56 (returns AUC * n; ties keep row order) */
int dfb_auc(dfb_handle h, const float* label, const float* pred, size_t n, float* auc_times_n);

/* ---------------------------------------------------------------------------------
 * (B) fused device-resident step: replaces the whole pull_callback of
 * SGDLearner::IterateData (src/sgd/sgd_learner.cc:138-177 + the kFeaCount push :214-217):
 *   [fea_cnt push] -> Pull -> GetPos -> Predict -> Evaluate -> penalty -> AUC ->
 *   [CalcGrad -> Push(kGradient)]
 * on a batch already localized by Localizer::Compact (CSR<u32> + sorted unique reversed
 * keys [+ counts]).  The table never leaves HBM; the host sees only dfb_progress.
 * cnt_or_null != NULL reproduces the epoch-0 feature-count push; is_train == 0 is a
 * validation batch (no CalcGrad/Push, sgd_learner.cc:158-171).
 * pred_out_or_null receives the nrows clamped predictions if not NULL.
 * ------------------------------------------------------------------------------- */
int dfb_train_step(dfb_handle h, size_t nrows, const uint64_t* offset, const uint32_t* index,
                   const float* value_or_null, const float* label, const uint64_t* keys,
                   size_t nkeys, const float* cnt_or_null, int is_train, dfb_progress* out,
                   float* pred_out_or_null);

/* same, but every array is already resident in device memory on the handle's device and
 * the call only enqueues work on the handle's stream (no host synchronisation);
 * dfb_sync() / dfb_read_progress() collect the result. */
int dfb_train_step_dev(dfb_handle h, size_t nrows, size_t nnz, const uint64_t* d_offset,
                       const uint32_t* d_index, const float* d_value_or_null,
                       const float* d_label, const uint64_t* d_keys, size_t nkeys,
                       const float* d_cnt_or_null, int is_train);
int dfb_sync(dfb_handle h);
/* blocks until enqueued steps are done; returns the Progress ACCUMULATED since the last
 * read (the reference merges per-batch Progress by addition, sgd_utils.h:66-71) */
int dfb_read_progress(dfb_handle h, dfb_progress* out);

/* pipelined host-buffer variant: enqueue H2D (from pinned memory: truly asynchronous)
 * + the step, double-buffered so that the copy of batch t+1 overlaps the compute of
 * batch t (the reference keeps <=2 batches in flight, sgd_learner.cc:220-223). */
int dfb_train_step_async(dfb_handle h, size_t nrows, const uint64_t* offset,
                         const uint32_t* index, const float* value_or_null, const float* label,
                         const uint64_t* keys, size_t nkeys, const float* cnt_or_null,
                         int is_train);

/* ---------------------------------------------------------------------------------
 * raw (un-localized) minibatches: Localizer::Compact on the device.
 * dfb_localize is Localizer(max_index).Compact (src/data/localizer.h:41-51, localizer.cc:11-103)
 * with host buffers in and out, bit-exact: index_out[nnz] = rank of every nnz's key, keys_out =
 * ascending unique ReverseBytes(id % max_index), cnt_out = occurrence counts (may be NULL).
 * dfb_train_step_raw* take the CSR<uint64> block the reader produced (BatchReader::Value) and run
 * Localizer(-1).Compact + [kFeaCount push when push_cnt] + the fused step; the CSC view the
 * gradient kernel needs falls out of the same radix sort.
 * ------------------------------------------------------------------------------- */
int dfb_localize(dfb_handle h, size_t nrows, const uint64_t* offset, const uint64_t* index,
                 uint64_t max_index, uint32_t* index_out, uint64_t* keys_out, float* cnt_out,
                 size_t* nkeys);
int dfb_train_step_raw(dfb_handle h, size_t nrows, const uint64_t* offset, const uint64_t* ids,
                       const float* value_or_null, const float* label, int push_cnt, int is_train,
                       dfb_progress* out, float* pred_out_or_null);
int dfb_train_step_raw_async(dfb_handle h, size_t nrows, const uint64_t* offset, const uint64_t* ids,
                             const float* value_or_null, const float* label, int push_cnt,
                             int is_train);
/* optional double-buffering hint: start the H2D copy of the batch the NEXT dfb_train_step_raw_async call
 * will be given (same pointers), so that the copy overlaps the step submitted before it */
int dfb_prefetch_raw(dfb_handle h, size_t nrows, const uint64_t* offset, const uint64_t* ids,
                     const float* value_or_null, const float* label);
int dfb_train_step_raw_dev(dfb_handle h, size_t nrows, size_t nnz, const uint64_t* d_offset,
                           const uint64_t* d_ids, const float* d_value_or_null, const float* d_label,
                           int push_cnt, int is_train);

/* blocks until the OLDEST not yet collected dfb_train_step_async step has finished and
 * returns that step's Progress (each async step snapshots its Progress into a pinned ring,
 * so collecting step t does not drain step t+1 from the pipeline).  Up to 8 steps may be outstanding;
 * beyond that the oldest snapshots are folded into the sum dfb_read_progress returns. */
int dfb_wait_step(dfb_handle h, dfb_progress* out);

/* per-stage device timing with CUDA events on the stream each stage runs on (bench.py's roofline):
 * stage 0 = key lookup + pull, 1 = FM forward kernel, 2 = AUC, 3 = CSC sort of the batch,
 * 4 = per-key gradient reduce + FTRL/AdaGrad update (+ InitV pass), 5 = GPU localizer (raw batches);
 * sharded store: 6 = worker: slice the batch + scatter to the owners, 7 = owner: lookup + partial interaction
 * sums of all workers' rows, 8 = worker: reduce the partials (pred / loss / p, p*XV back), 9 = owner: the
 * workers' updates (incl. waiting for their p*XV).
 * dfb_profile_read returns the accumulated milliseconds and launch counts of the DFB_NUM_STAGES
 * stages and resets them.  While profiling, AUC runs on the main stream (otherwise it overlaps
 * the update on an auxiliary stream).
 * dfb_time_mark(h, 0) / (h, 1) bracket a region with CUDA events that cover ALL streams of the handle
 * (mark 1 joins them); dfb_time_elapsed_ms synchronises and returns the device time between the marks. */
#define DFB_NUM_STAGES 10
int dfb_time_mark(dfb_handle h, int which);
int dfb_time_elapsed_ms(dfb_handle h, float* ms);
int dfb_profile(dfb_handle h, int enable);
int dfb_profile_read(dfb_handle h, double* stage_ms, uint64_t* stage_count);

/* ---------------------------------------------------------------------------------
 * model inspection (tests, checkpointing): read entries of the table.
 * scal_out[n][4] = {fea_cnt, w, sqrt_g, z} (SGDEntry, sgd_updater.h:19-29),
 * has_V_out[n] in {-1 absent, 0 no V, 1 V}; V_out/cg_out [n][V_dim] (may be NULL).
 * ------------------------------------------------------------------------------- */
int dfb_read_entries(dfb_handle h, const uint64_t* keys, size_t n, float* scal_out,
                     int* has_V_out, float* V_out, float* cg_out);
/* checkpoint = Updater::Save / Updater::Load (include/difacto/updater.h:40-47; TODO stubs in the
 * reference's SGDUpdater, sgd_updater.h:44-50, so the byte format is this library's: header +
 * records in ascending key order {key, fea_cnt, w, [sqrt_g, z], has_V, V[V_dim], [cg[V_dim]]}, the
 * bracketed fields only with save_aux).  dfb_restore needs an empty table with the same V_dim; a
 * snapshot without aux data restores a model that can predict but not train ("no aux data",
 * sgd_updater.cc:75). */
int dfb_snapshot_size(dfb_handle h, int save_aux, size_t* bytes);
int dfb_snapshot(dfb_handle h, int save_aux, void* buf, size_t bytes);
int dfb_restore(dfb_handle h, const void* buf, size_t bytes, int* has_aux_out);
/* state of the InitV random stream (SGDUpdaterParam::seed after the rand_r calls so far) */
int dfb_rng_state(dfb_handle h, uint32_t* seed);

/* ---------------------------------------------------------------------------------
 * key-range sharding (multi-GPU).  ps-lite Postoffice::GetServerKeyRanges
 * (ps-lite/src/postoffice.cc:127-136) + KVWorker::DefaultSlicer (kv_app.h:406-460):
 * shard i owns reversed keys in [UINT64_MAX/S*i, UINT64_MAX/S*(i+1)); the tail above
 * UINT64_MAX/S*S is clamped into shard S-1.  Because keys are sorted, every shard's keys
 * are one contiguous segment: bounds_out[S+1] are the segment boundaries.
 * ------------------------------------------------------------------------------- */
uint32_t dfb_key_owner(uint64_t reversed_key, uint32_t num_shards);
int dfb_shard_bounds(const uint64_t* sorted_keys, size_t n, uint32_t num_shards,
                     size_t* bounds_out);

/* device-side building blocks for the sharded store (all pointers are device memory, work is
 * enqueued on the handle's stream; difacto_b200/sharded.py holds the protocol):
 *   dfb_dev_feacnt     owner: Update(kFeaCount) for the keys one worker sent
 *   dfb_dev_pull_rows  owner side of Pull: keys -> dense rows {w[n], has_V[n] (-1/1), V[n][ks]}
 *                      packed for the all-to-all (SGDUpdater::Get, dense instead of ragged)
 *   dfb_dev_fm_step    worker: Predict/Evaluate/penalty/AUC/CalcGrad on the pulled dense buffers;
 *                      writes the COMPLETE gradient rows gw[n], gV[n][ks] (incl. the -V*XXp term
 *                      computed with the pulled V, fm_loss.h:181-188) for the keys with has_V >= 0
 *   dfb_dev_push_rows  owner side of Push(kGradient): FTRL/AdaGrad from the rows one worker sent;
 *                      has_V is the worker's pull-time view (lens[i] > 1, sgd_updater.cc:91)   */
int dfb_row_stride(dfb_handle h);  /* ks: V_dim rounded up to a multiple of 4 floats */
int dfb_dev_feacnt(dfb_handle h, const uint64_t* d_keys, size_t n, const float* d_cnt);
int dfb_dev_pull_rows(dfb_handle h, const uint64_t* d_keys, size_t n, float* d_w_out,
                      int* d_hasv_out, float* d_V_out);
int dfb_dev_fm_step(dfb_handle h, size_t nrows, size_t nnz, const uint64_t* d_offset,
                    const uint32_t* d_index, const float* d_value_or_null, const float* d_label,
                    size_t nkeys, const float* d_w, const int* d_hasv, const float* d_V,
                    int is_train, float* d_gw_out, float* d_gV_out);
int dfb_dev_push_rows(dfb_handle h, const uint64_t* d_keys, size_t n, const float* d_gw,
                      const int* d_hasv, const float* d_gV);
/* NVLink peer-store variants (one process per GPU, buffers shared through CUDA IPC):
 *   dfb_peer_alloc / dfb_peer_open   a device buffer + its 64-byte cudaIpcMemHandle_t; a peer process opens
 *                                    the handle and gets a pointer its kernels can store through
 *   dfb_dev_pull_rows_peer           owner: gather {w, has_V, V} for one requester's key segment and store the
 *                                    rows straight into the requester's pull buffer (fused gather + transfer);
 *                                    the pull-time has_V flags are also kept locally for the later push
 *   dfb_dev_fm_step_peer             worker: as dfb_dev_fm_step (training), but the gradient rows of key
 *                                    segment s (keys [seg_bounds[s], seg_bounds[s+1])) are stored straight into
 *                                    owner s's receive buffer (peer_gw[s], peer_gV[s] = where the segment starts) */
int dfb_peer_alloc(dfb_handle h, size_t bytes, void** ptr, unsigned char* handle64);
int dfb_peer_open(dfb_handle h, const unsigned char* handle64, void** ptr);
int dfb_peer_close(dfb_handle h, void* ptr);
int dfb_peer_free(dfb_handle h, void* ptr);
int dfb_dev_pull_rows_peer(dfb_handle h, const uint64_t* d_keys, size_t n, float* peer_w_out,
                           int* peer_hasv_out, float* peer_V_out, int* d_hasv_local_out);
int dfb_dev_fm_step_peer(dfb_handle h, size_t nrows, size_t nnz, const uint64_t* d_offset,
                         const uint32_t* d_index, const float* d_value_or_null, const float* d_label,
                         size_t nkeys, const float* d_w, const int* d_hasv, const float* d_V, int nseg,
                         const size_t* seg_bounds, float* const* peer_gw, float* const* peer_gV,
                         int first_seg /* segment to start with (own rank): staggers peer traffic */);

/* ---------------------------------------------------------------------------------
 * The NVLink-sharded store behind the C-ABI: N engines (one per GPU; one process per GPU with CUDA IPC, or
 * all of them in one process with peer access) form one model, rank r owning the r-th range of the reversed key space exactly
 * like ps-lite's servers (postoffice.cc:127-136).  It replaces, for this path, the worker/server exchange of
 * the reference (SGDLearner's workers calling Store::Pull / Push on KVWorker, the servers running
 * SGDUpdater; src/sgd/sgd_learner.cc:78-89,138-177, ps-lite/include/ps/kv_app.h:406-460).
 * Every rank calls dfb_shard_step_* once per minibatch, COLLECTIVELY (same number of calls, same push_cnt /
 * is_train; a rank that ran out of data passes nrows = 0).  A step is bulk-synchronous like the
 * all_to_all protocol of dfb_dev_*: every worker's forward sees the model after all updates of the previous
 * step; an owner applies the workers' gradients as separate Updates in rank order (sgd_updater.cc:74-98),
 * each taken at the V the worker "pulled" at step start.  What crosses NVLink is not the k-wide rows but the
 * per-example partial interaction sums (see csrc/kernels_shard.cu); nothing is synchronised with the host.
 *   dfb_shard_init     allocate this rank's mailbox (capacity: max_rows x max_nnz per minibatch; seg_keys /
 *                      seg_nnz = capacity of one (worker, owner) key segment, 0 = twice the even share)
 *   dfb_shard_export   the mailbox pointer (peers in the same process) and its CUDA IPC handle (peers in
 *                      other processes open it with dfb_peer_open)
 *   dfb_shard_connect  the peers' mailbox pointers, indexed by rank (entry [rank] is ignored)
 *   dfb_shard_step_dev   one minibatch of raw CSR<uint64> already on the device (enqueue only)
 *   dfb_shard_step_async the same from host buffers (pinned: asynchronous), double-buffered like
 *                        dfb_train_step_raw_async; dfb_prefetch_raw may stage the next batch early
 * Results: dfb_wait_step (per step) / dfb_read_progress (accumulated), as for the other async entry points;
 * Progress is this rank's own minibatch (loss, AUC, penalty of the weights it used).  DFB_ERR_TIMEOUT: a peer
 * did not reach the step within shard_timeout_ms (kwarg, default 20000).
 * ------------------------------------------------------------------------------- */
int dfb_shard_init(dfb_handle h, int rank, int nranks, size_t max_rows, size_t max_nnz, size_t seg_keys,
                   size_t seg_nnz, size_t* mailbox_bytes_out);
int dfb_shard_export(dfb_handle h, void** mailbox_ptr, unsigned char* handle64);
int dfb_shard_connect(dfb_handle h, void* const* peer_mailbox);
int dfb_shard_step_dev(dfb_handle h, size_t nrows, size_t nnz, const uint64_t* d_offset, const uint64_t* d_ids,
                       const float* d_value_or_null, const float* d_label, int push_cnt, int is_train);
int dfb_shard_step_async(dfb_handle h, size_t nrows, const uint64_t* offset, const uint64_t* ids,
                         const float* value_or_null, const float* label, int push_cnt, int is_train);
/* The same step split into its five enqueue phases (0 worker: localize + scatter the slices, 1 owner: lookup +
 * partial sums, 2 worker: reduce + p*XV back, 3 owner: updates, 4 finish): dfb_shard_step_* is begin + phases
 * 0..4.  Needed whenever ONE host thread drives several ranks (the C++ CLI's num_gpus = N, the one-device tests):
 * the phases must be interleaved -- phase p of every rank before phase p+1 of any -- so that every device-side
 * wait refers to work that is already enqueued.  With monolithic steps, rank 0's pollers would wait for work
 * the thread has not enqueued yet; on one device that never completes, and on several devices it deadlocks as
 * soon as a driver call in between blocks (a lazily loaded kernel synchronises its context).  Ranks that have a
 * host thread or a process of their own just call dfb_shard_step_*.
 * On the wire every minibatch is valued (x = 1 for a binary one), so the workers of a step may hold any mix of
 * binary, valued and empty minibatches. */
int dfb_shard_begin_async(dfb_handle h, size_t nrows, const uint64_t* offset, const uint64_t* ids,
                          const float* value_or_null, const float* label, int push_cnt, int is_train);
int dfb_shard_begin_dev(dfb_handle h, size_t nrows, size_t nnz, const uint64_t* d_offset, const uint64_t* d_ids,
                        const float* d_value_or_null, const float* d_label, int push_cnt, int is_train);
int dfb_shard_phase(dfb_handle h, int phase);
int dfb_shard_info(dfb_handle h, int* rank, int* nranks, size_t* seg_keys, size_t* seg_nnz, uint64_t* steps);

/* the CUDA stream (cudaStream_t) the handle enqueues on, for event interop with torch */
void* dfb_stream(dfb_handle h);

#ifdef __cplusplus
}
#endif
#endif  /* DIFACTO_B200_H_ */
