/*
 * oracle/fm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See fm_oracle.h.
 *
 * Scalar C restatement of the reference FM-SGD hot path.  Arithmetic types are
 * mirrored operation by operation (float where the reference computes in real_t,
 * double where a double literal promotes the expression) so that on the same
 * libm this file is bit-identical to the compiled reference for everything that
 * is not an OpenMP reduction.  Compile with -ffp-contract=off.
 */
#include "fm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* parameters: src/sgd/sgd_param.h:94-106                                     */
void orc_param_default(orc_param* p) {
  p->l1 = 1.0f;
  p->l2 = 0.0f;
  p->V_l2 = 0.01f;
  p->lr = 0.01f;
  p->lr_beta = 1.0f;
  p->V_lr = 0.01f;
  p->V_lr_beta = 1.0f;
  p->V_init_scale = 0.01f;
  p->V_threshold = 10;
  p->V_dim = 0; /* required field in the reference (no default) */
  p->seed = 0;
}

/* ------------------------------------------------------------------------- */
/* include/difacto/base.h:39-51: reverse the order of the 16 nibbles           */
uint64_t orc_reverse_bytes(uint64_t x) {
  uint64_t r = 0;
  for (int nib = 0; nib < 16; ++nib) {
    r = (r << 4) | (x & 0xFULL);
    x >>= 4;
  }
  return r;
}

/* glibc rand_r (stdlib/rand_r.c): three LCG steps, 11+10+10 result bits      */
int orc_rand_r(unsigned int* seed) {
  unsigned int next = *seed;
  unsigned int result;
  next = next * 1103515245u + 12345u;
  result = (next / 65536u) % 2048u;
  next = next * 1103515245u + 12345u;
  result = (result << 10) ^ ((next / 65536u) % 1024u);
  next = next * 1103515245u + 12345u;
  result = (result << 10) ^ ((next / 65536u) % 1024u);
  *seed = next;
  return (int)result;
}

/* ps-lite Postoffice::GetServerKeyRanges (src/postoffice.cc:127-136): server i owns
 * [kMaxKey/S*i, kMaxKey/S*(i+1)); kMaxKey = UINT64_MAX (ps/base.h:20).  Keys at or
 * above kMaxKey/S*S fall outside every range there; we clamp them to S-1. */
uint32_t orc_owner(uint64_t key, uint32_t S) {
  uint64_t width = UINT64_MAX / (uint64_t)S;
  uint64_t o = key / width;
  return (uint32_t)(o >= S ? S - 1 : o);
}

/* ------------------------------------------------------------------------- */
/* Localizer: src/data/localizer.cc:11-103                                    */
typedef struct { uint64_t k; uint32_t i; } kpair;

static int kpair_cmp(const void* a, const void* b) {
  uint64_t ka = ((const kpair*)a)->k, kb = ((const kpair*)b)->k;
  return ka < kb ? -1 : (ka > kb ? 1 : 0);
}

size_t orc_localize(size_t nrows, const uint64_t* offset, const uint64_t* index,
                    uint64_t max_index, uint32_t* out_index, uint64_t* out_keys,
                    float* out_cnt) {
  if (nrows == 0) return 0;                       /* localizer.cc:16 */
  size_t nnz = (size_t)offset[nrows];
  if (nnz == 0) return 0;
  kpair* pr = (kpair*)malloc(nnz * sizeof(kpair));
  for (size_t i = 0; i < nnz; ++i) {              /* localizer.cc:22-26 */
    pr[i].k = orc_reverse_bytes(index[i] % max_index);
    pr[i].i = (uint32_t)i;
  }
  qsort(pr, nnz, sizeof(kpair), kpair_cmp);       /* localizer.cc:28-29 */
  size_t nu = 0;                                  /* localizer.cc:36-49 */
  uint64_t curr = pr[0].k;
  float cnt = 0;
  for (size_t i = 0; i < nnz; ++i) {
    if (pr[i].k != curr) {
      out_keys[nu] = curr;
      if (out_cnt) out_cnt[nu] = cnt;
      ++nu;
      curr = pr[i].k;
      cnt = 0;
    }
    cnt += 1.0f;
    out_index[pr[i].i] = (uint32_t)nu;            /* RemapIndex localizer.cc:64-77: rank of key */
  }
  out_keys[nu] = curr;
  if (out_cnt) out_cnt[nu] = cnt;
  ++nu;
  free(pr);
  return nu;
}

/* ------------------------------------------------------------------------- */
/* model: unordered_map<feaid_t, SGDEntry> (sgd_updater.h:19-29,84)            */
typedef struct {
  uint64_t key;
  int used;
  float fea_cnt, w, sqrt_g, z;
  float* V; /* 2*V_dim floats: V then AdaGrad accumulators (sgd_updater.cc:142) */
} entry;

struct orc_model {
  orc_param p;
  entry* tab;
  size_t cap, size;
};

static uint64_t mix64(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
  return h;
}

orc_model* orc_model_create(const orc_param* p) {
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  m->p = *p;
  m->cap = 1024;
  m->tab = (entry*)calloc(m->cap, sizeof(entry));
  return m;
}

void orc_model_free(orc_model* m) {
  if (!m) return;
  for (size_t i = 0; i < m->cap; ++i) free(m->tab[i].V);
  free(m->tab);
  free(m);
}

size_t orc_model_size(const orc_model* m) { return m->size; }
unsigned int orc_model_seed(const orc_model* m) { return m->p.seed; }

static entry* probe(entry* tab, size_t cap, uint64_t key) {
  size_t h = (size_t)mix64(key) & (cap - 1);
  while (tab[h].used && tab[h].key != key) h = (h + 1) & (cap - 1);
  return &tab[h];
}

static void grow(orc_model* m) {
  size_t ncap = m->cap * 2;
  entry* nt = (entry*)calloc(ncap, sizeof(entry));
  for (size_t i = 0; i < m->cap; ++i) {
    if (m->tab[i].used) *probe(nt, ncap, m->tab[i].key) = m->tab[i];
  }
  free(m->tab);
  m->tab = nt;
  m->cap = ncap;
}

/* model_[key]: find or default-construct (all-zero entry, V == nullptr) */
static entry* find_or_insert(orc_model* m, uint64_t key) {
  entry* e = probe(m->tab, m->cap, key);
  if (e->used) return e;
  if ((m->size + 1) * 2 > m->cap) {
    grow(m);
    e = probe(m->tab, m->cap, key);
  }
  memset(e, 0, sizeof(*e));
  e->key = key;
  e->used = 1;
  ++m->size;
  return e;
}

int orc_model_lookup(const orc_model* m, uint64_t key, float* scal, float* V2k) {
  const entry* e = probe(m->tab, m->cap, key);
  if (!e->used) return -1;
  scal[0] = e->fea_cnt; scal[1] = e->w; scal[2] = e->sqrt_g; scal[3] = e->z;
  if (!e->V) return 0;
  if (V2k) memcpy(V2k, e->V, 2 * (size_t)m->p.V_dim * sizeof(float));
  return 1;
}

/* SGDUpdater::InitV, sgd_updater.cc:140-147 */
static void init_V(orc_model* m, entry* e) {
  int n = m->p.V_dim;
  e->V = (float*)malloc(2 * (size_t)n * sizeof(float));
  for (int i = 0; i < n; ++i) {
    /* rand_r()/(real_t)RAND_MAX is a float division; "- 0.5" promotes to double */
    float u = (float)orc_rand_r(&m->p.seed) / (float)2147483647;
    e->V[i] = (float)(((double)u - 0.5) * (double)m->p.V_init_scale);
  }
  memset(e->V + n, 0, (size_t)n * sizeof(float));
}

/* SGDUpdater::Get, sgd_updater.cc:32-56 */
size_t orc_get(orc_model* m, const uint64_t* keys, size_t n, float* vals_out,
               int* lens_out, size_t* nlens) {
  int k = m->p.V_dim;
  size_t p = 0;
  *nlens = k == 0 ? 0 : n;
  for (size_t i = 0; i < n; ++i) {
    entry* e = find_or_insert(m, keys[i]);
    vals_out[p++] = e->w;
    if (e->V) {
      memcpy(vals_out + p, e->V, (size_t)k * sizeof(float));
      p += (size_t)k;
      lens_out[i] = k + 1;
    } else if (k != 0) {
      lens_out[i] = 1;
    }
  }
  return p;
}

/* SGDUpdater::Update, kFeaCount branch, sgd_updater.cc:62-73 */
void orc_update_feacnt(orc_model* m, const uint64_t* keys, size_t n, const float* cnt) {
  for (size_t i = 0; i < n; ++i) {
    entry* e = find_or_insert(m, keys[i]);
    e->fea_cnt += cnt[i];
    if (m->p.V_dim > 0 && e->V == NULL && e->w != 0 &&
        e->fea_cnt > (float)m->p.V_threshold) {
      init_V(m, e);
    }
  }
}

/* SGDUpdater::UpdateW (FTRL-proximal), sgd_updater.cc:104-127 */
static void update_w(orc_model* m, float gw, entry* e) {
  const orc_param* P = &m->p;
  float sg = e->sqrt_g;
  float w = e->w;
  gw += w * P->l2;
  e->sqrt_g = sqrtf(sg * sg + gw * gw);
  e->z -= gw - (e->sqrt_g - sg) / P->lr * w;
  float z = e->z;
  float l1 = P->l1;
  if (z <= l1 && z >= -l1) {
    e->w = 0;
  } else {
    float eta = (P->lr_beta + e->sqrt_g) / P->lr;
    e->w = (z > 0 ? z - l1 : z + l1) / eta;
  }
  if (w == 0 && e->w != 0) {
    if (P->V_dim > 0 && e->V == NULL && e->fea_cnt > (float)P->V_threshold) init_V(m, e);
  }
}

/* SGDUpdater::UpdateV (AdaGrad), sgd_updater.cc:129-138 */
static void update_V(orc_model* m, const float* gV, entry* e) {
  const orc_param* P = &m->p;
  int n = P->V_dim;
  for (int i = 0; i < n; ++i) {
    float g = gV[i] + P->V_l2 * e->V[i];
    float cg = e->V[i + n];
    e->V[i + n] = sqrtf(cg * cg + g * g);
    float eta = P->V_lr / (e->V[i + n] + P->V_lr_beta);
    e->V[i] -= eta * g;
  }
}

/* SGDUpdater::Update, kGradient branch, sgd_updater.cc:74-98 */
int orc_update_grad(orc_model* m, const uint64_t* keys, size_t n, const float* grads,
                    size_t nvals, const int* lens, size_t nlens) {
  int w_only = nlens == 0;
  if (w_only) { if (nvals != n) return -1; } else { if (nlens != n) return -1; }
  size_t p = 0;
  for (size_t i = 0; i < n; ++i) {
    entry* e = find_or_insert(m, keys[i]);
    update_w(m, grads[p++], e);
    if (!w_only && lens[i] > 1) {
      if (lens[i] != m->p.V_dim + 1) return -1;
      if (e->V == NULL) return -1;
      update_V(m, grads + p, e);
      p += (size_t)m->p.V_dim;
    }
  }
  return p == nvals ? 0 : -1;
}

/* ------------------------------------------------------------------------- */
/* SGDLearner::GetPos, sgd_learner.cc:113-127 */
void orc_get_pos(const int* lens, size_t n, int* w_pos, int* V_pos) {
  int p = 0;
  for (size_t i = 0; i < n; ++i) {
    int l = lens[i];
    w_pos[i] = l == 0 ? -1 : p;
    V_pos[i] = l > 1 ? p + 1 : -1;
    p += l;
  }
}

/* position helpers: spmv.h:173-191, spmm.h:161-169 */
static inline int pos_of(const int* pos, size_t idx, int k) {
  if (pos) return pos[idx];        /* -1 = absent */
  return (int)(idx * (size_t)k);   /* direct indexing */
}

/* FMLoss::Predict, fm_loss.h:67-119 */
void orc_fm_predict(int V_dim, size_t nrows, const uint64_t* offset, const uint32_t* index,
                    const float* value, const float* weights, size_t nweights,
                    const int* w_pos, const int* V_pos, size_t npos, float* pred,
                    float* XV_out) {
  (void)nweights; (void)npos;
  /* pred += X * w  (SpMV::Times spmv.h:108-134) */
  for (size_t i = 0; i < nrows; ++i) {
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      int p = pos_of(w_pos, index[j], 1);
      float xj = p < 0 ? 0.0f : weights[p];
      if (xj == 0) continue;                       /* spmv.h:125 */
      if (value) pred[i] += xj * value[j]; else pred[i] += xj;
    }
  }
  if (V_dim == 0) return;                          /* fm_loss.h:77: before the clamp */
  int k = V_dim;
  float* XV = (float*)calloc(nrows * (size_t)k, sizeof(float));
  float* XXVV = (float*)calloc(nrows * (size_t)k, sizeof(float));
  for (size_t i = 0; i < nrows; ++i) {
    float* t = XV + i * (size_t)k;
    float* tt = XXVV + i * (size_t)k;
    /* XV = X*V (spmm.h:94-122) */
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      int p = pos_of(V_pos, index[j], k);
      if (p < 0) continue;
      const float* Vj = weights + p;
      if (value) { float v = value[j]; for (int l = 0; l < k; ++l) t[l] += Vj[l] * v; }
      else       { for (int l = 0; l < k; ++l) t[l] += Vj[l]; }
    }
    /* XXVV = (X.*X)*(V.*V) (fm_loss.h:86-105) */
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      int p = pos_of(V_pos, index[j], k);
      if (p < 0) continue;
      const float* Vj = weights + p;
      if (value) {
        float v = value[j]; v *= v;
        for (int l = 0; l < k; ++l) { float vv = Vj[l] * Vj[l]; tt[l] += vv * v; }
      } else {
        for (int l = 0; l < k; ++l) { float vv = Vj[l] * Vj[l]; tt[l] += vv; }
      }
    }
    /* pred += .5 * sum(XV.^2 - XXVV) (fm_loss.h:108-115; ".5 * s" is a double product) */
    float s = 0;
    for (int l = 0; l < k; ++l) s += t[l] * t[l] - tt[l];
    pred[i] = (float)((double)pred[i] + .5 * (double)s);
    /* projection (fm_loss.h:118) */
    pred[i] = pred[i] > 20 ? 20 : (pred[i] < -20 ? -20 : pred[i]);
  }
  if (XV_out) memcpy(XV_out, XV, nrows * (size_t)k * sizeof(float));
  free(XV);
  free(XXVV);
}

/* Loss::Evaluate, loss.h:57-66.  The reference sums with
 * "#pragma omp parallel for reduction(+:objv) num_threads(nthreads_)": libgomp's static
 * schedule gives thread t a contiguous chunk (the first n%T threads get one extra row),
 * each thread sums its chunk in float from 0, and the partials are added to objv.  The
 * combination order of the partials is unspecified for T>2; we add them in thread order. */
float orc_evaluate_mt(const float* label, const float* pred, size_t n, int nthreads) {
  float objv = 0;
  size_t T = nthreads < 1 ? 1 : (size_t)nthreads;
  size_t q = n / T, r = n % T, begin = 0;
  for (size_t t = 0; t < T; ++t) {
    size_t len = q + (t < r ? 1 : 0);
    float part = 0;
    for (size_t i = begin; i < begin + len; ++i) {
      float y = label[i] > 0 ? 1.0f : -1.0f;
      part += logf(1 + expf(-y * pred[i]));
    }
    objv += part;
    begin += len;
  }
  return objv;
}
/* DEFAULT_NTHREADS = 2 (include/difacto/base.h:28; sgd_learner.h:90) */
float orc_evaluate(const float* label, const float* pred, size_t n) {
  return orc_evaluate_mt(label, pred, n, 2);
}

/* FMLoss::CalcGrad, fm_loss.h:148-199 */
void orc_fm_calc_grad(int V_dim, size_t nrows, const uint64_t* offset, const uint32_t* index,
                      const float* value, const float* label, const float* weights,
                      size_t nweights, const int* w_pos, const int* V_pos, size_t npos,
                      const float* pred, float* grad) {
  (void)nweights;
  int k = V_dim;
  float* p = (float*)malloc(nrows * sizeof(float));
  for (size_t i = 0; i < nrows; ++i) {             /* fm_loss.h:155-161 */
    float y = label[i] > 0 ? 1.0f : -1.0f;
    p[i] = -y / (1 + expf(y * pred[i]));
  }
  /* grad_w = X' p at w_pos (spmv.h:140-171; column accumulation is in row order) */
  for (size_t i = 0; i < nrows; ++i) {
    float xi = p[i];
    if (xi == 0) continue;                         /* spmv.h:153 */
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      int q = pos_of(w_pos, index[j], 1);
      if (q < 0) continue;
      if (value) grad[q] += xi * value[j]; else grad[q] += xi;
    }
  }
  if (k == 0) { free(p); return; }
  /* number of columns: V_pos.size() (fm_loss.h:177) */
  size_t ncol = npos;
  float* XXp = (float*)calloc(ncol ? ncol : 1, sizeof(float));
  for (size_t i = 0; i < nrows; ++i) {             /* XXp = (X.*X)' p, fm_loss.h:171-178 */
    float xi = p[i];
    if (xi == 0) continue;
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      uint32_t c = index[j];
      if (value) { float xx = value[j] * value[j]; XXp[c] += xi * xx; }
      else XXp[c] += xi;
    }
  }
  for (size_t c = 0; c < ncol; ++c) {              /* grad_V -= diag(XXp) V, fm_loss.h:181-188 */
    int q = V_pos[c];
    if (q < 0) continue;
    for (int l = 0; l < k; ++l) grad[q + l] -= weights[q + l] * XXp[c];
  }
  /* XV_ recomputed here (the reference keeps it as member state from Predict) then
   * scaled by p (fm_loss.h:191-195) */
  float* XV = (float*)calloc(nrows * (size_t)k, sizeof(float));
  for (size_t i = 0; i < nrows; ++i) {
    float* t = XV + i * (size_t)k;
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      int q = pos_of(V_pos, index[j], k);
      if (q < 0) continue;
      const float* Vj = weights + q;
      if (value) { float v = value[j]; for (int l = 0; l < k; ++l) t[l] += Vj[l] * v; }
      else       { for (int l = 0; l < k; ++l) t[l] += Vj[l]; }
    }
    for (int l = 0; l < k; ++l) t[l] *= p[i];
  }
  /* grad_V += X' diag(p) X V at V_pos (spmm.h:128-159, row order per column) */
  for (size_t i = 0; i < nrows; ++i) {
    const float* t = XV + i * (size_t)k;
    for (uint64_t j = offset[i]; j < offset[i + 1]; ++j) {
      int q = pos_of(V_pos, index[j], k);
      if (q < 0) continue;
      float* g = grad + q;
      if (value) { float v = value[j]; for (int l = 0; l < k; ++l) g[l] += t[l] * v; }
      else       { for (int l = 0; l < k; ++l) g[l] += t[l]; }
    }
  }
  free(XV);
  free(XXp);
  free(p);
}

/* BinClassMetric::AUC, bin_class_metric.h:35-56 */
typedef struct { float label, predict; uint32_t i; } auc_ent;
static int auc_cmp(const void* a, const void* b) {
  const auc_ent* x = (const auc_ent*)a; const auc_ent* y = (const auc_ent*)b;
  if (x->predict < y->predict) return -1;
  if (x->predict > y->predict) return 1;
  return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}
float orc_auc(const float* label, const float* pred, size_t n) {
  auc_ent* b = (auc_ent*)malloc((n ? n : 1) * sizeof(auc_ent));
  for (size_t i = 0; i < n; ++i) { b[i].label = label[i]; b[i].predict = pred[i]; b[i].i = (uint32_t)i; }
  qsort(b, n, sizeof(auc_ent), auc_cmp);
  float area = 0, cum_tp = 0;
  for (size_t i = 0; i < n; ++i) {
    if (b[i].label > 0) cum_tp += 1; else area += cum_tp;
  }
  free(b);
  if (cum_tp == 0 || cum_tp == (float)n) return 1;
  area /= cum_tp * ((float)n - cum_tp);
  return (area < 0.5 ? 1 - area : area) * (float)n;
}

/* SGDLearner::EvaluatePenalty, sgd_learner.cc:249-273 */
float orc_penalty(const orc_param* P, const float* weights, size_t nweights,
                  const int* w_pos, const int* V_pos, size_t npos) {
  float objv = 0;
  if (npos) {
    for (size_t i = 0; i < npos; ++i) {
      int p = w_pos[i];
      if (p == -1) continue;
      float w = weights[p];
      objv = (float)((double)objv + ((double)(P->l1 * fabsf(w)) + .5 * (double)P->l2 * (double)w * (double)w));
    }
    for (size_t i = 0; i < npos; ++i) {
      int p = V_pos[i];
      if (p == -1) continue;
      for (int l = 0; l < P->V_dim; ++l) {
        float V = weights[p + l];
        objv = (float)((double)objv + .5 * (double)P->V_l2 * (double)V * (double)V);
      }
    }
  } else {
    for (size_t i = 0; i < nweights; ++i) {
      float w = weights[i];
      objv = (float)((double)objv + ((double)(P->l1 * fabsf(w)) + .5 * (double)P->l2 * (double)w * (double)w));
    }
  }
  return objv;
}

/* SGDLearner::IterateData body, sgd_learner.cc:138-177,196-224 */
void orc_sgd_step(orc_model* m, size_t nrows, const uint64_t* offset, const uint64_t* index,
                  const float* value, const float* label, int is_train, int push_cnt,
                  float* progress) {
  size_t nnz = nrows ? (size_t)offset[nrows] : 0;
  int k = m->p.V_dim;
  uint32_t* lidx = (uint32_t*)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
  uint64_t* keys = (uint64_t*)malloc((nnz ? nnz : 1) * sizeof(uint64_t));
  float* cnt = (float*)malloc((nnz ? nnz : 1) * sizeof(float));
  size_t U = orc_localize(nrows, offset, index, UINT64_MAX, lidx, keys, cnt);
  if (push_cnt) orc_update_feacnt(m, keys, U, cnt);
  float* vals = (float*)malloc((U ? U : 1) * (size_t)(1 + k) * sizeof(float));
  int* lens = (int*)malloc((U ? U : 1) * sizeof(int));
  size_t nlens = 0;
  size_t nvals = orc_get(m, keys, U, vals, lens, &nlens);
  int* w_pos = NULL; int* V_pos = NULL;
  if (nlens) {
    w_pos = (int*)malloc(U * sizeof(int));
    V_pos = (int*)malloc(U * sizeof(int));
    orc_get_pos(lens, U, w_pos, V_pos);
  }
  progress[4] += (float)nrows;
  float* pred = (float*)calloc(nrows ? nrows : 1, sizeof(float));
  orc_fm_predict(k, nrows, offset, lidx, value, vals, nvals, w_pos, V_pos, nlens, pred, NULL);
  progress[0] += orc_evaluate(label, pred, nrows);
  progress[1] += orc_penalty(&m->p, vals, nvals, w_pos, V_pos, nlens);
  progress[2] += orc_auc(label, pred, nrows);
  if (is_train) {
    float* grads = (float*)calloc(nvals ? nvals : 1, sizeof(float));
    orc_fm_calc_grad(k, nrows, offset, lidx, value, label, vals, nvals, w_pos, V_pos, nlens, pred, grads);
    orc_update_grad(m, keys, U, grads, nvals, lens, nlens);
    free(grads);
  }
  free(pred); free(w_pos); free(V_pos); free(lens); free(vals);
  free(cnt); free(keys); free(lidx);
}
