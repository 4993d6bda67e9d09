// difacto_b200/csrc/dfb_device.cuh -- device-side helpers shared by the table and shard kernels.
//
// The per-key arithmetic of SGDUpdater (src/sgd/sgd_updater.cc of the reference) lives here so that
// the fused single-GPU kernels (kernels_table.cu) and the owner side of the NVLink-sharded store
// (kernels_shard.cu) run literally the same code.
#pragma once
#include "dfb_internal.cuh"

#include "../../include/difacto_b200.h"

namespace dfb {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ uint64_t hash64(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
  h ^= h >> 33;
  return h;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

__device__ __forceinline__ void raise_err(DevProgress* prog, int code) { atomicCAS(&prog->err, 0, code); }

// the count of a launch: a host value, or (when dn != nullptr) a value a previous kernel left on the
// device -- the sizes of a localized batch never visit the host (no synchronisation inside a step)
__device__ __forceinline__ size_t dev_count(size_t n_host, const unsigned long long* dn) {
  if (dn == nullptr) return n_host;
  const unsigned long long v = *dn;
  return v < (unsigned long long)n_host ? (size_t)v : n_host;   // n_host is the capacity the buffers were sized for
}

// ---- FTRL-proximal on w: SGDUpdater::UpdateW, sgd_updater.cc:104-127 ----
// returns true when w went 0 -> nonzero (the InitV trigger :121-126)
__device__ __forceinline__ bool ftrl_step(const Params& p, float gw, float& w, float& sqrt_g, float& z) {
  const float sg = sqrt_g, w0 = w;
  gw = __fadd_rn(gw, __fmul_rn(w0, p.l2));
  sqrt_g = __fsqrt_rn(__fadd_rn(__fmul_rn(sg, sg), __fmul_rn(gw, gw)));
  // z -= gw - (sqrt_g' - sg) / lr * w
  z = __fsub_rn(z, __fsub_rn(gw, __fmul_rn(__fdiv_rn(__fsub_rn(sqrt_g, sg), p.lr), w0)));
  const float l1 = p.l1;
  if (z <= l1 && z >= -l1) {
    w = 0.f;
  } else {
    const float eta = __fdiv_rn(__fadd_rn(p.lr_beta, sqrt_g), p.lr);
    w = __fdiv_rn(z > 0.f ? __fsub_rn(z, l1) : __fadd_rn(z, l1), eta);
  }
  return w0 == 0.f && w != 0.f;
}

// ---- AdaGrad on one V component: SGDUpdater::UpdateV, sgd_updater.cc:129-138 ----
__device__ __forceinline__ void adagrad_step(const Params& p, float gV, float& v, float& cg) {
  const float g = __fadd_rn(gV, __fmul_rn(p.V_l2, v));
  cg = __fsqrt_rn(__fadd_rn(__fmul_rn(cg, cg), __fmul_rn(g, g)));
  const float eta = __fdiv_rn(p.V_lr, __fadd_rn(cg, p.V_lr_beta));
  v = __fsub_rn(v, __fmul_rn(eta, g));
}

// penalty of one w (sgd_learner.cc:257 ; evaluated in fp32 here, double there)
__device__ __forceinline__ float pen_w(const Params& p, float w) {
  return p.l1 * fabsf(w) + 0.5f * p.l2 * w * w;
}

// one 32-byte table entry fetched with ONE 256-bit request (LDG.256, sm_100), L2-only so that an entry another
// thread of the same launch just inserted is never served stale from L1
struct Entry256 {
  unsigned long long key, q1, q2, q3;      // {key | vrow, pad | fea_cnt, w | sqrt_g, z}
  __device__ __forceinline__ int vrow() const { return (int)(unsigned)(q1 & 0xffffffffULL); }
  __device__ __forceinline__ float w() const { return __uint_as_float((unsigned)(q2 >> 32)); }
};
__device__ __forceinline__ Entry256 load_entry(const Entry* p) {
  Entry256 e;
  asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];"
               : "=l"(e.key), "=l"(e.q1), "=l"(e.q2), "=l"(e.q3)
               : "l"(p));
  return e;
}

// model_[key] (sgd_updater.cc:43-45,65-67,86-88): find or default-construct.  Returns the slot or -1.
// On a hit (or a fresh insert) *lo receives the first half of the entry {key, vrow, pad} as loaded.
template <bool INSERT>
__device__ __forceinline__ int table_find(const Table& t, unsigned long long key, uint64_t h) {
  for (uint64_t probe = 0; probe <= t.mask; ++probe) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&t.tab[h].key);
    if (cur == key) return (int)h;
    if (cur == kEmptyKey) {
      if (!INSERT) return -1;
      const unsigned long long prev = atomicCAS(&t.tab[h].key, kEmptyKey, key);
      if (prev == kEmptyKey) {
        const unsigned long long nk = atomicAdd(&t.state->n_keys, 1ULL) + 1;
        atomicAdd(&t.prog->new_keys, 1ULL);
        if (nk > t.max_keys) raise_err(t.prog, DFB_ERR_CAPACITY);
        return (int)h;
      }
      if (prev == key) return (int)h;
    }
    h = (h + 1) & t.mask;
  }
  if (INSERT) raise_err(t.prog, DFB_ERR_CAPACITY);
  return -1;
}

}  // namespace dfb
