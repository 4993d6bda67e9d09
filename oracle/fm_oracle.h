/*
 * oracle/fm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (scalar, single-thread) restatement of the reference's FM-SGD hot path
 * (dmlc/difacto @ 78e3562).  Every function cites the reference file:line it
 * follows.  It is the parity checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (difacto_b200/) never links, imports or calls it.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against the reference's
 * own golden values (tests/cpp/fm_loss_test.cc:35,39,78,82;
 * tests/cpp/localizer_test.cc:26-27,48-49; tests/cpp/sgd_learner_test.cc:10-30)
 * through the committed fixtures in tests/golden/, and
 * tests/test_oracle_vs_ref.py checks it bit-for-bit against the compiled
 * reference (oracle/_ref/libdifacto_ref.so) on random inputs when that library
 * is present.
 */
#ifndef ORACLE_FM_ORACLE_H_
#define ORACLE_FM_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SGDUpdaterParam, src/sgd/sgd_param.h:66-107 (same names, same defaults) */
typedef struct {
  float l1, l2, V_l2;
  float lr, lr_beta, V_lr, V_lr_beta;
  float V_init_scale;
  int V_dim;
  int V_threshold;
  unsigned int seed;
} orc_param;

void orc_param_default(orc_param* p);

/* include/difacto/base.h:39-51 */
uint64_t orc_reverse_bytes(uint64_t x);

/* glibc rand_r restated (stdlib/rand_r.c of glibc 2.x): used by InitV, sgd_updater.cc:144 */
int orc_rand_r(unsigned int* seed);

/* ps-lite key-range owner: src/postoffice.cc:127-136 + DefaultSlicer kv_app.h:406-460.
 * owner = min(S-1, key / (UINT64_MAX / S)) applied to the (already reversed) key. */
uint32_t orc_owner(uint64_t reversed_key, uint32_t num_shards);

/* Localizer::Compact, src/data/localizer.cc:11-103.  Returns #unique keys.
 * out_index[nnz] (rank of each nnz's key), out_keys[<=nnz] ascending reversed keys,
 * out_cnt[<=nnz] occurrence counts (may be NULL). */
size_t orc_localize(size_t nrows, const uint64_t* offset, const uint64_t* index,
                    uint64_t max_index, uint32_t* out_index, uint64_t* out_keys,
                    float* out_cnt);

/* ------------ the model: SGDUpdater, src/sgd/sgd_updater.{h,cc} ------------ */
typedef struct orc_model orc_model;
orc_model* orc_model_create(const orc_param* p);
void orc_model_free(orc_model* m);
size_t orc_model_size(const orc_model* m);
unsigned int orc_model_seed(const orc_model* m);

/* SGDUpdater::Get, sgd_updater.cc:32-56.  vals_out holds n*(1+V_dim); returns #vals
 * written; lens_out[n] filled when V_dim>0 (*nlens = n) else *nlens = 0. */
size_t orc_get(orc_model* m, const uint64_t* keys, size_t n, float* vals_out,
               int* lens_out, size_t* nlens);
/* SGDUpdater::Update(kFeaCount), sgd_updater.cc:62-73 */
void orc_update_feacnt(orc_model* m, const uint64_t* keys, size_t n, const float* cnt);
/* SGDUpdater::Update(kGradient), sgd_updater.cc:74-98.  nlens==0 => w_only.
 * returns 0, or -1 if a CHECK of the reference would have fired. */
int orc_update_grad(orc_model* m, const uint64_t* keys, size_t n, const float* grads,
                    size_t nvals, const int* lens, size_t nlens);
/* read one entry: scal[4] = {fea_cnt, w, sqrt_g, z}; V2k[2*V_dim] = {V, cg} if allocated.
 * returns -1 if absent, 0 if no V, 1 if V allocated */
int orc_model_lookup(const orc_model* m, uint64_t key, float* scal, float* V2k);

/* SGDLearner::GetPos, src/sgd/sgd_learner.cc:113-127 */
void orc_get_pos(const int* lens, size_t n, int* w_pos, int* V_pos);

/* FMLoss::Predict, src/loss/fm_loss.h:67-119 (+ spmv.h:108-134, spmm.h:94-122).
 * pred is ACCUMULATED into (caller zero-fills).  w_pos/V_pos may be NULL (direct
 * indexing, the V_dim==0 path).  XV_out (nrows*V_dim) receives the XV_ member state,
 * may be NULL. */
void orc_fm_predict(int V_dim, size_t nrows, const uint64_t* offset, const uint32_t* index,
                    const float* value, const float* weights, size_t nweights,
                    const int* w_pos, const int* V_pos, size_t npos, float* pred,
                    float* XV_out);
/* Loss::Evaluate, include/difacto/loss.h:57-66.  _mt mirrors the OpenMP static-schedule
 * float reduction over nthreads chunks; orc_evaluate uses the reference default (2). */
float orc_evaluate_mt(const float* label, const float* pred, size_t n, int nthreads);
float orc_evaluate(const float* label, const float* pred, size_t n);
/* FMLoss::CalcGrad, src/loss/fm_loss.h:148-199 (+ spmv.h:140-171, spmm.h:128-159).
 * grad (nweights) is ACCUMULATED into (caller zero-fills). */
void orc_fm_calc_grad(int V_dim, size_t nrows, const uint64_t* offset, const uint32_t* index,
                      const float* value, const float* label, const float* weights,
                      size_t nweights, const int* w_pos, const int* V_pos, size_t npos,
                      const float* pred, float* grad);
/* BinClassMetric::AUC, src/loss/bin_class_metric.h:35-56.  The reference uses an
 * unstable std::sort, so the result for tied predictions across classes is
 * implementation-defined there; here ties keep original row order (stable). */
float orc_auc(const float* label, const float* pred, size_t n);
/* SGDLearner::EvaluatePenalty, src/sgd/sgd_learner.cc:249-273 */
float orc_penalty(const orc_param* p, const float* weights, size_t nweights,
                  const int* w_pos, const int* V_pos, size_t npos);

/* one minibatch of SGDLearner::IterateData (sgd_learner.cc:129-227) without file I/O.
 * progress[5] = {loss, penalty, auc, nnz_w, nrows} accumulated. */
void orc_sgd_step(orc_model* m, size_t nrows, const uint64_t* offset, const uint64_t* index,
                  const float* value, const float* label, int is_train, int push_cnt,
                  float* progress);

#ifdef __cplusplus
}
#endif
#endif  /* ORACLE_FM_ORACLE_H_ */
