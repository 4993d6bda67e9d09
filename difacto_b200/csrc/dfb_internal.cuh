// difacto_b200/csrc/dfb_internal.cuh -- shared declarations of the sm_100a engine.
//
// Data layout in HBM (one shard = one GPU):
//   Entry tab[cap]        32-byte open-addressing hash entries, one DRAM sector each:
//                         {u64 reversed key, i32 vrow, i32 pad, f32 fea_cnt, w, sqrt_g, z}
//                         == the reference's unordered_map<feaid_t,SGDEntry> (sgd_updater.h:19-29,84)
//   float VV[vcap][2*ks]  one row per key with an embedding: [V (ks) | AdaGrad accumulators (ks)],
//                         ks = V_dim rounded up to 4 floats -- the reference's `new real_t[2n]`
//                         (sgd_updater.cc:142).  The gather reads the first half (one contiguous
//                         4k-byte chunk); the update reads and writes the whole 8k-byte row.
// V rows are allocated lazily from a bump pool, exactly when the reference calls InitV.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dfb {

struct __align__(32) Entry {
  unsigned long long key;
  int vrow;       // -1: no V allocated (SGDEntry::V == nullptr)
  int pad;
  float fea_cnt, w, sqrt_g, z;
};
static_assert(sizeof(Entry) == 32, "Entry must be one 32-byte sector");

constexpr unsigned long long kEmptyKey = ~0ULL;  // ReverseBytes(id % (2^64-1)) never yields it

// SGDUpdaterParam (src/sgd/sgd_param.h:66-107)
struct Params {
  float l1, l2, V_l2, lr, lr_beta, V_lr, V_lr_beta, V_init_scale;
  int V_dim, V_threshold;
  unsigned seed;
};

// accumulated on device, read back by dfb_read_progress
struct DevProgress {
  double loss, penalty, auc;
  unsigned long long nrows, new_keys, new_vrows;
  int err;        // sticky dfb_status raised by a kernel (capacity, invalid lens ...)
  int pad;
};

// device-resident mutable scalars of the table
struct TableState {
  unsigned long long n_keys;
  unsigned long long n_vrows;
  unsigned int seed;     // SGDUpdaterParam::seed as advanced by rand_r (sgd_updater.cc:144)
  unsigned int pad;
};

struct Table {
  Entry* tab = nullptr;
  uint64_t cap = 0, mask = 0, max_keys = 0;
  float* V = nullptr;
  float* Vcg = nullptr;
  uint64_t vcap = 0;
  int ks = 0;  // V_dim rounded up to a multiple of 4 floats
  int rs = 0;  // stride between table rows in floats (2*ks); Vcg == V + ks
  TableState* state = nullptr;
  DevProgress* prog = nullptr;
};

// how a kernel finds w / V (and where it puts the gradient) for local key u of a batch
struct FmView {
  // w = wbase[w_pos ? w_pos[u] : u]; w_pos[u] == -1 -> 0
  const float* wbase;
  const int* w_pos;
  // optional packed pulled view {bits(w), vrow} per key: one 8-byte load per nnz instead of two
  // 4-byte loads from different arrays (fast kernel only; overrides wbase/w_pos/v_pos when set)
  const int2* wv;
  // V row = vbase + (int64)r * vstride, r = v_pos[u] (or u when dense && v_pos[u] >= 0); -1 -> absent
  const float* vbase;
  const int* v_pos;
  long long vstride;
  int dense;
  // gradient destinations (TRAIN): gw[gw_pos ? gw_pos[u] : u], gV row = gvbase + r' * gvstride with
  // r' = (gv_pos ? gv_pos[u] : u); gxxp[u] (only when values are present)
  float* gwbase;
  const int* gw_pos;
  float* gvbase;
  const int* gv_pos;
  long long gvstride;
  float* gxxp;
  // 1: per-load L2 policies -- the per-nnz {w,vrow} view and the p*XV rows are re-read (evict_last), the
  // V rows are touched once (evict_first)
  int l2hint;
};

// where the worker's gradient rows go: locally contiguous (nseg == 0) or, per owner segment of the
// sorted key list, straight into that owner's receive buffer over NVLink (peer pointers)
struct SegDst {
  int nseg;
  int rot;            // first key (multiple of 32) this rank starts with: staggers the peers' inbound traffic
  int bounds[9];      // key index boundaries of the segments (nseg + 1 used)
  float* gw[8];       // per segment: destination of gw for the segment's first key
  float* gV[8];       // per segment: destination row of gV for the segment's first key
};

// owner side of the fused sharded store (SRC 2 of k_bwd_update): what differs from the single-GPU update
struct ShardApply {
  const float* w_pulled;        // w of every key as pulled at step start (penalty of the pulled weights)
  const unsigned char* conf;    // != 0: a lower-rank worker's push already updated this key in this step ...
  const float* vsave;           // ... and this is its V row as pulled at step start ([n][K])
};

// very hot keys (a Criteo feature that occurs in most rows of a batch): their occurrence lists are cut into chunks
// that are reduced by separate warps beforehand (k_hot_reduce), in a fixed order, so that no single warp walks a
// 65536-row list; k_bwd_update then adds the chunk partials in chunk order (bit-reproducible)
struct HotPart {
  const int* hotmap;        // per key: first chunk id (only valid for keys longer than `split`), < 0: not pre-reduced
  const float* part;        // [chunks][K] partial sum_occ x * pXV[row]
  const float2* part_s;     // [chunks] partial {sum x p, sum x^2 p}
  int split, chunk;         // lists longer than split are cut into chunks of `chunk` occurrences; part == nullptr: off
};
struct HotWs {              // device workspaces of the pre-reduction
  int* hotmap; int2* info; float* part; float2* part_s; unsigned long long* counter; int cap;
};

// forward "partials" of the fused sharded store (MODE 3 of k_fm_fast): the owner of a key segment computes,
// for every row of every worker's minibatch, the part of the FM interaction that is linear in ITS rows
//   XV_i^(s) = sum_{j in segment s} x_ij V_j,  sum_j (x_ij V_j)^2,  sum_j x_ij w_j
// and stores it straight into that worker's mailbox (peer memory over NVLink).
// what a worker tells the owner of a key segment about its minibatch (written into the owner's mailbox)
struct ShardHdr {
  unsigned long long nkeys;   // keys of the worker's batch that fall into this owner's range
  unsigned long long nnz;     // non-zeros over those keys
  unsigned long long nrows;   // rows of the worker's batch
  unsigned long long flags;   // bit 0: is_train, bit 1: push_cnt, bit 2: valued
  unsigned long long step;
  unsigned long long pad[3];
};
static_assert(sizeof(ShardHdr) == 64, "ShardHdr is one 64-byte record");
__host__ __device__ inline unsigned long long shard_hdr_nrows(const ShardHdr* h) { return h->nrows; }

struct PartSrc {
  const uint64_t* rowptr;   // [nrows+1] offsets into ridx (this worker's rows restricted to my key segment)
  const uint32_t* ridx;     // key index relative to the segment
  const float* rval;        // values (nullptr: binary)
  const int2* wv;           // pulled view {bits(w), vrow} of the segment's keys
  const ShardHdr* hdr;      // nrows of the worker's batch
  float* out_xv;            // [nrows][K]   (peer memory)
  float2* out_sc;           // [nrows] {sum (xV)^2, sum x w}
};
struct PartArgs {
  int nsrc, rot;            // rot: first source this owner starts with (staggers the peers' inbound traffic)
  unsigned long long bcap;  // rows are enumerated as src * bcap + row
  PartSrc s[8];
};

struct FmBatch {
  size_t nrows;
  const uint64_t* offset;   // size_t offsets, as in dmlc::RowBlock (dmlc/data.h:141)
  const uint32_t* index;
  const float* value;       // nullptr: binary features
  const float* label;
  const float* pred_in;     // CalcGrad with a caller-provided pred (fm_loss.h:155-161), else nullptr
  float* pred_io;           // output; accumulated into when pred_acc
  int pred_acc;
  int V_dim;
  int train;
  DevProgress* prog;        // loss / nrows accumulated here when non-null
  // emit mode (train && emit): instead of scattering with atomics the kernel writes
  int emit;
  float* p_out;             //   p_i = -y/(1+exp(y pred_i))            [nrows]
  float* pxv_out;           //   p_i * XV_i                             [nrows][V_dim]
  uint32_t* occ_row;        //   row of every nnz (binary data)         [nnz]
  unsigned long long* occ_rowx;  // (row << 32 | bits(x)) (valued data) [nnz]
  // long rows (predict / emit on the fast path): rows of >= long_nnz nonzeros are left to k_fm_long, where a whole
  // CTA walks one row (0 = one warp per row whatever its length); nnz_hint = the batch's nnz when the host knows it
  unsigned long_nnz;
  size_t nnz_hint;
};

// true when launch_fm would take the 16-byte-lane fast path for this V_dim / view
bool fm_fast_supported(int V_dim);

// ---- launchers (kernels_fm.cu) ----
// returns number of kernel launches performed, or <0 on invalid configuration
int launch_fm(const FmBatch& b, const FmView& v, int force_generic, cudaStream_t s);
// the bulk-copy (cp.async.bulk / UBLKCP) variant of the predict kernel for V_dim = 64 (kernels_fm_tma.cu, an A/B);
// returns 0 when the configuration is outside the experiment
int launch_fm_tma_predict(const FmBatch& b, const FmView& v, cudaStream_t s);
// MODE 3: partial interaction sums of every source's rows (see PartArgs); v.vbase/vstride = the table rows
int launch_fm_partial(int V_dim, bool valued, const FmView& v, const PartArgs& pa, cudaStream_t s);
int launch_grad_finalize(int V_dim, size_t nkeys, const float* weights, const int* V_pos,
                         const float* xxp, float* grad, cudaStream_t s);

int launch_grad_finalize_dense(int V_dim, int ks, size_t nkeys, const int* hasv, const float* V,
                               const float* xxp, float* gV, cudaStream_t s);

// ---- launchers (kernels_table.cu) ----
extern int g_lookup_ilp, g_lookup_ctas, g_update_persistent;
int launch_table_init(Table& t, unsigned seed, cudaStream_t s);
// find (or insert) keys; slot_out[i] = hash position or -1.  When pull outputs are non-null also
// emits w and vrow of each entry.
// n is a host count, or (dn != nullptr) the capacity with the actual count read from the device
int launch_lookup(Table& t, const uint64_t* keys, size_t n, const unsigned long long* dn, bool insert,
                  int* slot_out, float* w_out, int* vrow_out, int2* wv_out, cudaStream_t s);
// SGDUpdater::Update(kFeaCount) incl. the InitV pass.  flags/pos/cub_tmp are workspaces of n ints.
// counts: cnt[n], or (cnt == nullptr) the differences of cnt_cols[n+1].  ws: (n+31)/32 + 2 ints.
int launch_feacnt(Table& t, const Params& p, const int* slot, size_t n, const unsigned long long* dn,
                  const float* cnt, const int* cnt_cols, int* flags, int* ws, cudaStream_t s);
size_t scan_tmp_bytes(size_t n);
size_t sort_tmp_bytes(size_t n);
// InitV for flagged keys, consuming the rand_r stream in key order (sgd_updater.cc:140-147)
int launch_initv(Table& t, const Params& p, const int* slot, size_t n, const unsigned long long* dn,
                 int* flags, int* ws, cudaStream_t s);
// SGDUpdater::Get packing: lens -> scan -> ragged [w, V...] (sgd_updater.cc:32-56)
int launch_pack_ragged(Table& t, const Params& p, const int* slot, size_t n, int* lens, int* pos,
                       float* vals, unsigned long long* nvals_out, void* cub_tmp, size_t cub_bytes,
                       cudaStream_t s);
// dense rows for the sharded store: w[n], hasv[n] (-1/1), V[n][ks]
int launch_gather_rows(Table& t, const int* slot, size_t n, float* w_out, int* hasv_out,
                       int* hasv_out2, float* V_out, cudaStream_t s);
// FTRL/AdaGrad from per-key dense gradient rows (fused path and sharded push).
//   pull_vrow[i] >= 0 <=> the worker saw a V row at pull time (lens[i] > 1); when
//   vrow_is_flag the actual row is taken from the entry.
//   xxp_mode 0: gV is the complete gradient; 1: grad_V = gV - V*gxxp[i]; 2: grad_V = gV - V*gw[i]
//   (binary data, where XXp == grad_w).
int launch_update_dense(Table& t, const Params& p, const int* slot, const int* pull_vrow,
                        int vrow_is_flag, size_t n, const float* gw, const float* gxxp,
                        const float* gV, int* flags, int accumulate_penalty, int xxp_mode,
                        cudaStream_t s);
// FTRL/AdaGrad from the reference's ragged gradient layout (sgd_updater.cc:74-98)
int launch_update_ragged(Table& t, const Params& p, const int* slot, size_t n, const float* grads,
                         const int* lens_or_null, const int* pos, int* flags, cudaStream_t s);
// penalty over a pulled view without updating (validation batches, sharded worker)
int launch_penalty(const Params& p, DevProgress* prog, const float* w_arr, const int* vrow,
                   const float* V, int ks, int dense, size_t n, const unsigned long long* dn, cudaStream_t s);
int launch_pull_view(Table& t, const int* slot, size_t n, const unsigned long long* dn, float* w_out,
                     int* vrow_out, int2* wv_out, cudaStream_t s);
int launch_lens_scan(const int* lens, size_t n, int* pos, void* cub_tmp, size_t cub_bytes,
                     cudaStream_t s);
// Loss::Evaluate / BinClassMetric::AUC on device; results added to prog (or written to out)
int launch_evaluate(const float* label, const float* pred, size_t n, double* out, cudaStream_t s);
int launch_auc(const float* label, const float* pred, size_t n, float* key_tmp2, float* val_tmp2,
               void* cub_tmp, size_t cub_bytes, double* out_add, cudaStream_t s);
int launch_restore(Table& t, const uint64_t* keys, size_t n, const float* scal, const int* vrow,
                   unsigned long long n_vrows, unsigned seed, cudaStream_t s);
int launch_read_entries(Table& t, const int* slot, size_t n, float* scal, int* hasv, float* V,
                        float* cg, int k, cudaStream_t s);

// ---- GPU localizer (kernels_localize.cu): Localizer::Compact + the CSC view from one radix sort ----
size_t localize_sort_tmp_bytes(size_t nnz);
// hi32: sort on the upper 32 key bits only (the caller knows the lower 32 are zero: ids < 2^32); verified on the device
int launch_localize_keys(const uint64_t* ids, size_t nnz, uint64_t max_index, unsigned long long* rkeys,
                         uint32_t* pos, unsigned long long* or_all, const uint64_t* offset, size_t nrows,
                         uint32_t* nnz_row, bool hi32, cudaStream_t s);
int launch_localize_sort(const unsigned long long* rkeys, const uint32_t* pos, size_t nnz, int begin_bit, bool hi32,
                         unsigned long long* skeys, uint32_t* spos, int* head, int* rank1, void* tmp,
                         size_t tmp_bytes, const uint32_t* nnz_row, const float* value, uint64_t* keys_out,
                         int* col_start, int* col_end, uint32_t* lidx_out, void* occ_sorted,
                         unsigned long long* scal /* {or_all, n_unique} */, DevProgress* prog, cudaStream_t s);
int launch_cnt_from_cols(const int* col_start, const int* col_end, size_t n, float* cnt, cudaStream_t s);

// ---- sorted (atomic-free, deterministic) gradient reduction ----
// CSC view of the batch: stable radix sort of (local key id -> occurrence payload); afterwards
// key u owns occ_sorted[col_start[u] .. col_end[u]) in row order (the reference's summation
// order per column, spmm.h:140-156).
size_t csc_tmp_bytes(size_t nnz, bool valued);
int launch_csc_build(const uint32_t* lidx, const void* occ, bool valued, size_t nnz, size_t nkeys,
                     uint32_t* lidx_sorted, void* occ_sorted, int* col_start, int* col_end,
                     void* cub_tmp, size_t cub_bytes, DevProgress* prog, cudaStream_t s);
// per key: grad = sum_occ x * pXV[row] - V * XXp, then either FTRL/AdaGrad in place (apply) or
// complete dense gradient rows out (gw_out, gV_out = sum - V_pulled*XXp; the sharded worker).
int launch_bwd_update(Table& t, const Params& p, const int* slot, const int* pull_vrow, size_t n,
                      const unsigned long long* dn, const int* col_start, const int* col_end,
                      const void* occ_sorted, bool valued, const float* p_row, const float* pxv, int* flags,
                      int accumulate_penalty, const ShardApply* shard, const HotPart* hot, cudaStream_t s);
// pre-reduction of the very hot keys' occurrence lists (split > 0); fills *hp for launch_bwd_update
int launch_hot_prereduce(int V_dim, size_t n, const unsigned long long* dn, const int* col_start, const int* col_end,
                         const void* occ_sorted, bool valued, const float* p_row, const float* pxv, int split,
                         const HotWs& ws, HotPart* hp, cudaStream_t s);
int launch_bwd_dense(const Params& p, DevProgress* prog, int ks, const float* w_pulled,
                     const int* hasv, size_t n, const int* col_start, const int* col_end,
                     const void* occ_sorted, bool valued, const float* p_row, const float* pxv,
                     float* gw_out, const float* V_pulled, float* gV_out, int accumulate_penalty,
                     const SegDst* seg, cudaStream_t s);
// owner side of the sharded Push for V_dim in {8,16,32,64,128}: FTRL/AdaGrad from complete dense
// gradient rows; returns -1 for other V_dim (use launch_update_dense)
int launch_update_pushed(Table& t, const Params& p, const int* slot, const int* hasv, size_t n,
                         const float* gw, const float* gV, int* flags, cudaStream_t s);

}  // namespace dfb
