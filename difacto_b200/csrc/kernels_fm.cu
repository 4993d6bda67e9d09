// difacto_b200/csrc/kernels_fm.cu -- FM forward (+ fused backward scatter) for sm_100a.
//
// Restates, for the GPU, FMLoss::Predict (src/loss/fm_loss.h:67-119: SpMV::Times,
// 2x SpMM::Times, row reduce, clamp), Loss::Evaluate (include/difacto/loss.h:57-66) and
// FMLoss::CalcGrad (fm_loss.h:148-199: p = -y/(1+exp(y pred)), SpMV::TransTimes,
// SpMM::TransTimes) of the reference as ONE pass per CSR row:
//
//   phase 1  gather the k-wide V rows of the row's active features (coalesced 16-byte lanes,
//            UNR independent loads in flight per lane), accumulate XV = sum x V and
//            sum (x V)^2 in registers, warp-shuffle reduce over k -> pred, logloss
//   phase 2  (training) p from the clamped pred, then scatter x p XV into the per-key gradient
//            rows with vector fp32 reductions (red.global.add.v4.f32), x p into grad_w and
//            x^2 p into XXp.  The "- V_j XXp_j" term (fm_loss.h:181-188) needs V once per
//            unique key, not per nnz, so it is applied where V is read anyway: in the fused
//            FTRL/AdaGrad kernel (kernels_table.cu) or in grad_finalize below.
//
// The whole path is HBM-bound (about 0.5 FLOP/B): no tensor cores on purpose.
#include "dfb_internal.cuh"

#include <math.h>
#include <string.h>

namespace dfb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float4 ldg128(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// per-load L2 eviction policies (createpolicy + ld.global.nc.L2::cache_hint)
__device__ __forceinline__ uint64_t l2_policy(int kind) {   // 0 normal, 1 evict_first, 2 evict_last
  uint64_t pol;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg128_pol(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int2 ldg64_pol(const int2* p, uint64_t pol) {
  int2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

__device__ __forceinline__ void red_add(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// logloss of one row: log(1 + exp(-y pred)), y = label > 0 ? 1 : -1 (loss.h:62-63)
__device__ __forceinline__ float row_logloss(float label, float pred) {
  float y = label > 0.f ? 1.f : -1.f;
  return logf(1.f + expf(-y * pred));
}

// p = -y / (1 + exp(y pred))  (fm_loss.h:157-161)
__device__ __forceinline__ float row_p(float label, float pred) {
  float y = label > 0.f ? 1.f : -1.f;
  return -y / (1.f + expf(y * pred));
}

__device__ __forceinline__ void block_add_loss(float loss_acc, size_t nrows, DevProgress* prog) {
  __shared__ float red_s[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // lane 0 of each warp holds that warp's partial
  if (lane == 0) red_s[wid] = loss_acc;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    float v = lane < nw ? red_s[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0 && prog) {
      atomicAdd(&prog->loss, (double)v);
      if (blockIdx.x == 0) atomicAdd(&prog->nrows, (unsigned long long)nrows);
    }
  }
}

// ---------------------------------------------------------------------------------------
// fast path: K in {8,16,32,64,128}, 16-byte aligned rows
// ---------------------------------------------------------------------------------------
// MODE 0: predict only.  MODE 1: training with fp32 red.global scatter into dense gradient rows.
// MODE 2: training, "emit": writes p_i, p_i*XV_i and the (row[,x]) payload of every nnz so that
//         the gradient can be reduced per key without atomics (kernels_table.cu: k_bwd_update).
// MODE 3: owner side of the fused sharded store: partial sums of every worker's rows over this owner's
//         key segment, stored into the workers' mailboxes (see PartArgs).
#ifndef DFB_FM_MINBLOCKS
#define DFB_FM_MINBLOCKS 4
#endif
#ifndef DFB_FM_PF_MAXK
#define DFB_FM_PF_MAXK 16       // metadata prefetch (one chunk ahead) for V_dim <= this
#endif
#ifndef DFB_FM_MINBLOCKS32
#define DFB_FM_MINBLOCKS32 DFB_FM_MINBLOCKS     // V_dim = 32 (tuning builds: 3 removes its ~130-byte spills)
#endif
template <int K, int MODE, bool HAS_VAL>
__global__ void __launch_bounds__(256, K == 32 ? DFB_FM_MINBLOCKS32 : DFB_FM_MINBLOCKS) k_fm_fast(FmBatch b, FmView v, PartArgs pa) {
  constexpr bool TRAIN = MODE == 1;
  constexpr int LPR = K / 4;                 // lanes per V row, one float4 each
  constexpr int G = 32 / LPR;                // V rows fetched by one warp-wide load
  constexpr int UNR = (32 / G) < 8 ? (32 / G) : 8;   // independent loads in flight per lane
  constexpr bool PF = K <= DFB_FM_PF_MAXK;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  float loss_acc = 0.f;
  const uint64_t pol_v = l2_policy(v.l2hint ? 1 : 0), pol_wv = l2_policy(v.l2hint ? 2 : 0);
  const size_t total_rows = MODE == 3 ? (size_t)pa.nsrc * (size_t)pa.bcap : b.nrows;

  for (size_t vrow_id = warp0; vrow_id < total_rows; vrow_id += nwarps) {
    size_t row = vrow_id;
    const uint64_t* offp = b.offset;
    const uint32_t* idxp = b.index;
    const float* valp = b.value;
    const int2* wvp = v.wv;
    int src = 0;
    if (MODE == 3) {
      const int s0 = (int)(vrow_id / pa.bcap);
      row = vrow_id - (size_t)s0 * pa.bcap;
      src = s0 + pa.rot;
      if (src >= pa.nsrc) src -= pa.nsrc;
      if (row >= (size_t)shard_hdr_nrows(pa.s[src].hdr)) continue;
      offp = pa.s[src].rowptr; idxp = pa.s[src].ridx; valp = pa.s[src].rval; wvp = pa.s[src].wv;
    }
    const uint64_t o0 = offp[row], o1 = offp[row + 1];
    if ((MODE == 0 || MODE == 2) && b.long_nnz && o1 - o0 >= b.long_nnz) continue;   // k_fm_long's row
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc2 = 0.f, wsum = 0.f;

    // the per-nnz metadata {x, w, V-row index} of a 32-nnz chunk hangs on a dependent chain (index -> pulled view);
    // it is fetched one chunk ahead, so that chain overlaps the V-row gathers of the current chunk
    auto load_meta = [&](uint64_t c, float& x, float& w, int& vr) {
      const uint64_t j = c + lane;
      x = 0.f; w = 0.f; vr = -1;
      if (j < o1) {
        const uint32_t u = __ldg(idxp + j);
        x = HAS_VAL ? __ldg(valp + j) : 1.f;
        if (wvp) {
          const int2 t = ldg64_pol(wvp + u, pol_wv);
          w = __int_as_float(t.x);
          vr = t.y;
        } else {
          const int wp = v.w_pos ? __ldg(v.w_pos + u) : (int)u;
          w = wp >= 0 ? __ldg(v.wbase + wp) : 0.f;
          vr = __ldg(v.v_pos + u);
          if (v.dense && vr >= 0) vr = (int)u;
        }
        if (MODE == 2 && b.occ_row != nullptr) {   // nullptr: the CSC view already exists (GPU localizer)
          if (HAS_VAL) b.occ_rowx[j] = ((unsigned long long)row << 32) | (unsigned long long)__float_as_uint(x);
          else b.occ_row[j] = (uint32_t)row;
        }
      }
    };
    float x_n = 0.f, w_n = 0.f;
    int vr_n = -1;
    if (PF) load_meta(o0, x_n, w_n, vr_n);
    for (uint64_t c = o0; c < o1; c += 32) {
      float x, w;
      int vr;
      if (PF) {
        x = x_n; w = w_n; vr = vr_n;
        if (c + 32 < o1) load_meta(c + 32, x_n, w_n, vr_n);
      } else {
        load_meta(c, x, w, vr);
      }
      wsum = fmaf(x, w, wsum);
      const int cnt = (int)((o1 - c) < 32 ? (o1 - c) : 32);
      for (int t0 = 0; t0 < cnt; t0 += G * UNR) {
        float4 vv[UNR];
        float xs[UNR];
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
          const int t = t0 + q * G + grp;
          const int vr_t = __shfl_sync(kFull, vr, t & 31);
          const float x_t = __shfl_sync(kFull, x, t & 31);
          const bool ok = (t < cnt) && (vr_t >= 0);
          xs[q] = ok ? x_t : 0.f;
          vv[q] = ok ? ldg128_pol(v.vbase + (long long)vr_t * v.vstride + sub * 4, pol_v)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < UNR; ++q) {
          const float a0 = xs[q] * vv[q].x, a1 = xs[q] * vv[q].y;
          const float a2 = xs[q] * vv[q].z, a3 = xs[q] * vv[q].w;
          xv.x += a0; xv.y += a1; xv.z += a2; xv.w += a3;
          acc2 = fmaf(a0, a0, acc2); acc2 = fmaf(a1, a1, acc2);
          acc2 = fmaf(a2, a2, acc2); acc2 = fmaf(a3, a3, acc2);
        }
      }
    }
    // combine the G row-groups: afterwards every group holds the full XV for its 4 dims
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      xv.x += __shfl_xor_sync(kFull, xv.x, o);
      xv.y += __shfl_xor_sync(kFull, xv.y, o);
      xv.z += __shfl_xor_sync(kFull, xv.z, o);
      xv.w += __shfl_xor_sync(kFull, xv.w, o);
    }
    acc2 = warp_sum(acc2);
    wsum = warp_sum(wsum);
    if (MODE == 3) {
      // ||XV||^2 is not linear in the rows: the worker squares the SUM of the owners' partials
      if (grp == 0) *reinterpret_cast<float4*>(pa.s[src].out_xv + row * (size_t)K + sub * 4) = xv;
      if (lane == 0) pa.s[src].out_sc[row] = make_float2(acc2, wsum);
      continue;
    }
    float s1 = grp == 0 ? (xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w) : 0.f;
    s1 = warp_sum(s1);
    float pred = (b.pred_acc ? b.pred_io[row] : 0.f) + wsum;
    pred += 0.5f * (s1 - acc2);
    pred = pred > 20.f ? 20.f : (pred < -20.f ? -20.f : pred);   // fm_loss.h:118
    const float label = b.label ? __ldg(b.label + row) : 0.f;
    if (lane == 0) {
      if (b.pred_io) b.pred_io[row] = pred;
      if (b.label) loss_acc += row_logloss(label, pred);
    }
    if (MODE == 2) {
      const float p = row_p(label, pred);
      if (lane == 0) b.p_out[row] = p;
      // XV_ *= p (fm_loss.h:191-195), one coalesced 4K-byte row
      if (grp == 0)
        *reinterpret_cast<float4*>(b.pxv_out + row * (size_t)K + sub * 4) =
            make_float4(p * xv.x, p * xv.y, p * xv.z, p * xv.w);
    }

    if (TRAIN) {
      const float p = row_p(label, b.pred_in ? __ldg(b.pred_in + row) : pred);
      const float4 gx = make_float4(p * xv.x, p * xv.y, p * xv.z, p * xv.w);
      for (uint64_t c = o0; c < o1; c += 32) {
        const uint64_t j = c + lane;
        const bool valid = j < o1;
        uint32_t u = 0;
        float x = 0.f;
        int vr = -1;
        if (valid) {
          u = __ldg(b.index + j);
          x = HAS_VAL ? __ldg(b.value + j) : 1.f;
          vr = __ldg(v.v_pos + u);
          const int gp = v.gw_pos ? __ldg(v.gw_pos + u) : (int)u;
          if (gp >= 0) red_add(v.gwbase + gp, x * p);
          if (HAS_VAL && vr >= 0) red_add(v.gxxp + u, x * x * p);
        }
        const int cnt = (int)((o1 - c) < 32 ? (o1 - c) : 32);
        for (int t0 = 0; t0 < cnt; t0 += G) {
          const int t = t0 + grp;
          const int vr_t = __shfl_sync(kFull, vr, t & 31);
          const float x_t = __shfl_sync(kFull, x, t & 31);
          const uint32_t u_t = __shfl_sync(kFull, u, t & 31);
          if (t < cnt && vr_t >= 0) {
            const long long r = v.gv_pos ? (long long)__ldg(v.gv_pos + u_t) : (long long)u_t;
            red_add_v4(v.gvbase + r * v.gvstride + sub * 4, x_t * gx.x, x_t * gx.y, x_t * gx.z,
                       x_t * gx.w);
          }
        }
      }
    }
  }
  if (MODE != 3) block_add_loss(loss_acc, b.nrows, b.prog);
}

// ---------------------------------------------------------------------------------------
// long rows (gisette: ~5000 nonzeros per example): one CTA per row.  A single warp would walk such a row as
// ~nnz/16 dependent gather rounds with 32 lanes of loads in flight; here the 8 warps of a CTA take the row's
// 32-nnz chunks round-robin, and their partial (XV, sum (xV)^2, sum xw) are added in warp order through shared
// memory (fixed order: bit-reproducible).  Rows shorter than b.long_nnz are k_fm_fast's.  MODE 0 / 2 as above.
// ---------------------------------------------------------------------------------------
constexpr int kLongWarps = 8;

template <int K, int MODE, bool HAS_VAL>
__global__ void __launch_bounds__(kLongWarps * 32) k_fm_long(FmBatch b, FmView v, int scan_rows) {
  constexpr int LPR = K / 4;
  constexpr int G = 32 / LPR;
  constexpr int UNR = (32 / G) < 8 ? (32 / G) : 8;
  __shared__ float4 s_xv[kLongWarps][LPR];
  __shared__ float2 s_sc[kLongWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const uint64_t pol_v = l2_policy(v.l2hint ? 1 : 0), pol_wv = l2_policy(v.l2hint ? 2 : 0);
  float loss_acc = 0.f;                       // thread 0 only

  // a CTA looks at scan_rows (<= 256, one per thread) row lengths at a time: with many rows per CTA the scan is one
  // coalesced load (and almost always finds nothing); with few rows (a batch of long rows only) scan_rows is 1 and
  // every CTA gets its own row
  __shared__ unsigned char s_long[kLongWarps * 32];
  for (size_t base = (size_t)blockIdx.x * scan_rows; base < b.nrows; base += (size_t)gridDim.x * scan_rows) {
    const size_t mine = base + threadIdx.x;
    bool is_long = false;
    if ((int)threadIdx.x < scan_rows && mine < b.nrows) is_long = b.offset[mine + 1] - b.offset[mine] >= b.long_nnz;
    s_long[threadIdx.x] = is_long ? 1 : 0;
    if (!__syncthreads_or(is_long)) continue;
    for (int r = 0; r < scan_rows; ++r) {                // in row order (the loss sum stays reproducible)
      if (!s_long[r]) continue;                          // CTA-uniform
      const size_t row = base + r;
      const uint64_t o0 = b.offset[row], o1 = b.offset[row + 1];
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      float acc2 = 0.f, wsum = 0.f;
      for (uint64_t c = o0 + 32u * wid; c < o1; c += 32u * kLongWarps) {
        const uint64_t j = c + lane;
        float x = 0.f, w = 0.f;
        int vr = -1;
        if (j < o1) {
          const uint32_t u = __ldg(b.index + j);
          x = HAS_VAL ? __ldg(b.value + j) : 1.f;
          if (v.wv) {
            const int2 t = ldg64_pol(v.wv + u, pol_wv);
            w = __int_as_float(t.x);
            vr = t.y;
          } else {
            const int wp = v.w_pos ? __ldg(v.w_pos + u) : (int)u;
            w = wp >= 0 ? __ldg(v.wbase + wp) : 0.f;
            vr = __ldg(v.v_pos + u);
            if (v.dense && vr >= 0) vr = (int)u;
          }
          if (MODE == 2 && b.occ_row != nullptr) {
            if (HAS_VAL) b.occ_rowx[j] = ((unsigned long long)row << 32) | (unsigned long long)__float_as_uint(x);
            else b.occ_row[j] = (uint32_t)row;
          }
        }
        wsum = fmaf(x, w, wsum);
        const int cnt = (int)((o1 - c) < 32 ? (o1 - c) : 32);
        for (int t0 = 0; t0 < cnt; t0 += G * UNR) {
          float4 vv[UNR];
          float xs[UNR];
  #pragma unroll
          for (int q = 0; q < UNR; ++q) {
            const int t = t0 + q * G + grp;
            const int vr_t = __shfl_sync(kFull, vr, t & 31);
            const float x_t = __shfl_sync(kFull, x, t & 31);
            const bool ok = (t < cnt) && (vr_t >= 0);
            xs[q] = ok ? x_t : 0.f;
            vv[q] = ok ? ldg128_pol(v.vbase + (long long)vr_t * v.vstride + sub * 4, pol_v)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
  #pragma unroll
          for (int q = 0; q < UNR; ++q) {
            const float a0 = xs[q] * vv[q].x, a1 = xs[q] * vv[q].y;
            const float a2 = xs[q] * vv[q].z, a3 = xs[q] * vv[q].w;
            xv.x += a0; xv.y += a1; xv.z += a2; xv.w += a3;
            acc2 = fmaf(a0, a0, acc2); acc2 = fmaf(a1, a1, acc2);
            acc2 = fmaf(a2, a2, acc2); acc2 = fmaf(a3, a3, acc2);
          }
        }
      }
  #pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        xv.x += __shfl_xor_sync(kFull, xv.x, o);
        xv.y += __shfl_xor_sync(kFull, xv.y, o);
        xv.z += __shfl_xor_sync(kFull, xv.z, o);
        xv.w += __shfl_xor_sync(kFull, xv.w, o);
      }
      acc2 = warp_sum(acc2);
      wsum = warp_sum(wsum);
      if (grp == 0) s_xv[wid][sub] = xv;
      if (lane == 0) s_sc[wid] = make_float2(acc2, wsum);
      __syncthreads();
      if (wid == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        float a2 = 0.f, ws = 0.f;
  #pragma unroll
        for (int q = 0; q < kLongWarps; ++q) {
          const float4 e = s_xv[q][sub];
          t.x += e.x; t.y += e.y; t.z += e.z; t.w += e.w;
          a2 += s_sc[q].x; ws += s_sc[q].y;
        }
        float s1 = grp == 0 ? (t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w) : 0.f;
        s1 = warp_sum(s1);
        float pred = (b.pred_acc ? b.pred_io[row] : 0.f) + ws;
        pred += 0.5f * (s1 - a2);
        pred = pred > 20.f ? 20.f : (pred < -20.f ? -20.f : pred);   // fm_loss.h:118
        const float label = b.label ? __ldg(b.label + row) : 0.f;
        if (lane == 0) {
          if (b.pred_io) b.pred_io[row] = pred;
          if (b.label) loss_acc += row_logloss(label, pred);
        }
        if (MODE == 2) {
          const float p = row_p(label, pred);
          if (lane == 0) b.p_out[row] = p;
          if (grp == 0)
            *reinterpret_cast<float4*>(b.pxv_out + row * (size_t)K + sub * 4) = make_float4(p * t.x, p * t.y, p * t.z, p * t.w);
        }
      }
      __syncthreads();
    }
    __syncthreads();                                      // s_long is rewritten by the next scan
  }
  // (nrows is counted by k_fm_fast)
  if (threadIdx.x == 0 && b.prog && loss_acc != 0.f) atomicAdd(&b.prog->loss, (double)loss_acc);
}

// ---------------------------------------------------------------------------------------
// generic path: any V_dim (including 0), unaligned / ragged rows (the pulled layout of the
// reference: [w_0,(V_0..)][w_1,(V_1..)] with w_pos/V_pos, sgd_updater.cc:46-53)
// ---------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(128) k_fm_generic(FmBatch b, FmView v) {
  extern __shared__ float smem[];
  const int k = b.V_dim;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* xv = smem + (size_t)wid * k;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const bool has_val = b.value != nullptr;
  float loss_acc = 0.f;

  for (size_t row = warp0; row < b.nrows; row += nwarps) {
    const uint64_t o0 = b.offset[row], o1 = b.offset[row + 1];
    for (int l = lane; l < k; l += 32) xv[l] = 0.f;
    float acc2 = 0.f, wsum = 0.f;
    for (uint64_t c = o0; c < o1; c += 32) {
      const uint64_t j = c + lane;
      const bool valid = j < o1;
      uint32_t u = 0;
      float x = 0.f, w = 0.f;
      long long vr = -1;
      if (valid) {
        u = b.index[j];
        x = has_val ? b.value[j] : 1.f;
        const int wp = v.w_pos ? v.w_pos[u] : (int)u;
        w = wp >= 0 ? v.wbase[wp] : 0.f;
        if (k > 0 && v.v_pos) {
          vr = v.v_pos[u];
          if (v.dense && vr >= 0) vr = u;
        }
      }
      wsum = fmaf(x, w, wsum);
      if (k > 0) {
        const int cnt = (int)((o1 - c) < 32 ? (o1 - c) : 32);
        for (int t = 0; t < cnt; ++t) {
          const long long vr_t = __shfl_sync(kFull, vr, t);
          const float x_t = __shfl_sync(kFull, x, t);
          if (vr_t < 0) continue;   // warp-uniform
          const float* rowp = v.vbase + vr_t * v.vstride;
          for (int l = lane; l < k; l += 32) {
            const float a = x_t * rowp[l];
            xv[l] += a;
            acc2 = fmaf(a, a, acc2);
          }
        }
      }
    }
    float pred = (b.pred_acc ? b.pred_io[row] : 0.f);
    wsum = warp_sum(wsum);
    pred += wsum;
    if (k > 0) {   // V_dim == 0 returns before the clamp (fm_loss.h:77)
      float s1 = 0.f;
      for (int l = lane; l < k; l += 32) s1 = fmaf(xv[l], xv[l], s1);
      s1 = warp_sum(s1);
      acc2 = warp_sum(acc2);
      pred += 0.5f * (s1 - acc2);
      pred = pred > 20.f ? 20.f : (pred < -20.f ? -20.f : pred);
    }
    const float label = b.label ? b.label[row] : 0.f;
    if (lane == 0) {
      if (b.pred_io) b.pred_io[row] = pred;
      if (b.label) loss_acc += row_logloss(label, pred);
    }
    if (TRAIN) {
      const float p = row_p(label, b.pred_in ? b.pred_in[row] : pred);
      for (uint64_t c = o0; c < o1; c += 32) {
        const uint64_t j = c + lane;
        const bool valid = j < o1;
        uint32_t u = 0;
        float x = 0.f;
        long long gr = -1;
        if (valid) {
          u = b.index[j];
          x = has_val ? b.value[j] : 1.f;
          const int gp = v.gw_pos ? v.gw_pos[u] : (int)u;
          if (gp >= 0) red_add(v.gwbase + gp, x * p);
          if (k > 0 && v.v_pos && v.v_pos[u] >= 0) {
            gr = v.gv_pos ? (long long)v.gv_pos[u] : (long long)u;
            if (v.gxxp) red_add(v.gxxp + u, (has_val ? x * x : 1.f) * p);
          }
        }
        if (k > 0) {
          const int cnt = (int)((o1 - c) < 32 ? (o1 - c) : 32);
          for (int t = 0; t < cnt; ++t) {
            const long long gr_t = __shfl_sync(kFull, gr, t);
            const float x_t = __shfl_sync(kFull, x, t);
            if (gr_t < 0) continue;
            float* g = v.gvbase + gr_t * v.gvstride;
            const float xp = x_t * p;
            for (int l = lane; l < k; l += 32) red_add(g + l, xp * xv[l]);
          }
        }
      }
    }
  }
  block_add_loss(loss_acc, b.nrows, b.prog);
}

// grad_V -= diag(XXp) V for the ragged layout (fm_loss.h:181-188)
__global__ void k_grad_finalize(int k, size_t nkeys, const float* weights, const int* V_pos,
                                const float* xxp, float* grad) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t u = warp0; u < nkeys; u += nwarps) {
    const int p = V_pos[u];
    if (p < 0) continue;
    const float s = xxp[u];
    for (int l = lane; l < k; l += 32) grad[p + l] -= weights[p + l] * s;
  }
}

// dense pulled layout: gV[u][ks] -= V[u][ks] * xxp[u] for keys with a V row (fm_loss.h:181-188)
__global__ void k_grad_finalize_dense(int k, int ks, size_t nkeys, const int* hasv, const float* V,
                                      const float* xxp, float* gV) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t u = warp0; u < nkeys; u += nwarps) {
    if (hasv[u] < 0) continue;
    const float s = xxp[u];
    for (int l = lane; l < k; l += 32) gV[u * (size_t)ks + l] -= V[u * (size_t)ks + l] * s;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int K>
int launch_fast_k(const FmBatch& b, const FmView& v, cudaStream_t s) {
  const int threads = 256;
  size_t need = (b.nrows + 7) / 8;
  int grid = (int)(need < (size_t)(148 * 8) ? (need ? need : 1) : (size_t)(148 * 8));
  const bool hv = b.value != nullptr;
  PartArgs pa;
  memset(&pa, 0, sizeof(pa));
  if (b.train && b.emit) {
    if (hv) k_fm_fast<K, 2, true><<<grid, threads, 0, s>>>(b, v, pa);
    else    k_fm_fast<K, 2, false><<<grid, threads, 0, s>>>(b, v, pa);
  } else if (b.train) {
    if (hv) k_fm_fast<K, 1, true><<<grid, threads, 0, s>>>(b, v, pa);
    else    k_fm_fast<K, 1, false><<<grid, threads, 0, s>>>(b, v, pa);
  } else {
    if (hv) k_fm_fast<K, 0, true><<<grid, threads, 0, s>>>(b, v, pa);
    else    k_fm_fast<K, 0, false><<<grid, threads, 0, s>>>(b, v, pa);
  }
  // rows of >= long_nnz nonzeros: one CTA each (the scan over the offsets costs ~1 us when there is none)
  if (b.long_nnz && (!b.train || b.emit) && (b.nnz_hint == 0 || b.nnz_hint >= b.long_nnz)) {
    const int lgrid = (int)(b.nrows < (size_t)(148 * 4) ? b.nrows : (size_t)(148 * 4));
    int scan = 1;
    while (scan < kLongWarps * 32 && (size_t)lgrid * scan * 2 <= b.nrows) scan *= 2;
    if (b.train) {
      if (hv) k_fm_long<K, 2, true><<<lgrid, kLongWarps * 32, 0, s>>>(b, v, scan);
      else    k_fm_long<K, 2, false><<<lgrid, kLongWarps * 32, 0, s>>>(b, v, scan);
    } else {
      if (hv) k_fm_long<K, 0, true><<<lgrid, kLongWarps * 32, 0, s>>>(b, v, scan);
      else    k_fm_long<K, 0, false><<<lgrid, kLongWarps * 32, 0, s>>>(b, v, scan);
    }
    return 2;
  }
  return 1;
}

template <int K>
int launch_partial_k(bool valued, const FmView& v, const PartArgs& pa, cudaStream_t s) {
  FmBatch b;
  memset(&b, 0, sizeof(b));
  b.V_dim = K;
  const int grid = 148 * 8;
  if (valued) k_fm_fast<K, 3, true><<<grid, 256, 0, s>>>(b, v, pa);
  else        k_fm_fast<K, 3, false><<<grid, 256, 0, s>>>(b, v, pa);
  return 1;
}

}  // namespace

bool fm_fast_supported(int k) { return k == 8 || k == 16 || k == 32 || k == 64 || k == 128; }

int launch_fm_partial(int V_dim, bool valued, const FmView& v, const PartArgs& pa, cudaStream_t s) {
  switch (V_dim) {
    case 8: return launch_partial_k<8>(valued, v, pa, s);
    case 16: return launch_partial_k<16>(valued, v, pa, s);
    case 32: return launch_partial_k<32>(valued, v, pa, s);
    case 64: return launch_partial_k<64>(valued, v, pa, s);
    case 128: return launch_partial_k<128>(valued, v, pa, s);
  }
  return -1;
}

int launch_fm(const FmBatch& b, const FmView& v, int force_generic, cudaStream_t s) {
  const int k = b.V_dim;
  bool fast = !force_generic && (k == 8 || k == 16 || k == 32 || k == 64 || k == 128) &&
              v.v_pos != nullptr && aligned16(v.vbase) && (v.vstride % 4 == 0);
  if (fast && b.train && !b.emit) {
    fast = aligned16(v.gvbase) && (v.gvstride % 4 == 0) && (b.value == nullptr || v.gxxp != nullptr);
  }
  if (b.train && b.emit && !fast) return -1;   // the emit mode exists only on the fast path
  if (fast) {
    switch (k) {
      case 8: return launch_fast_k<8>(b, v, s);
      case 16: return launch_fast_k<16>(b, v, s);
      case 32: return launch_fast_k<32>(b, v, s);
      case 64: return launch_fast_k<64>(b, v, s);
      case 128: return launch_fast_k<128>(b, v, s);
    }
  }
  if (k > 0 && v.v_pos == nullptr) return -1;   // V_pos is required when V_dim > 0
  const int threads = 128;
  const size_t smem = (size_t)4 * k * sizeof(float);
  if (smem > 200 * 1024) return -1;
  size_t need = (b.nrows + 3) / 4;
  int grid = (int)(need < (size_t)(148 * 16) ? (need ? need : 1) : (size_t)(148 * 16));
  if (b.train) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_fm_generic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_fm_generic<true><<<grid, threads, smem, s>>>(b, v);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_fm_generic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_fm_generic<false><<<grid, threads, smem, s>>>(b, v);
  }
  return 1;
}

int launch_grad_finalize_dense(int V_dim, int ks, size_t nkeys, const int* hasv, const float* V,
                               const float* xxp, float* gV, cudaStream_t s) {
  if (V_dim == 0 || nkeys == 0) return 0;
  size_t need = (nkeys + 7) / 8;
  int grid = (int)(need < (size_t)(148 * 8) ? need : (size_t)(148 * 8));
  k_grad_finalize_dense<<<grid, 256, 0, s>>>(V_dim, ks, nkeys, hasv, V, xxp, gV);
  return 1;
}

int launch_grad_finalize(int V_dim, size_t nkeys, const float* weights, const int* V_pos,
                         const float* xxp, float* grad, cudaStream_t s) {
  if (V_dim == 0 || nkeys == 0) return 0;
  size_t need = (nkeys + 7) / 8;
  int grid = (int)(need < (size_t)(148 * 8) ? need : (size_t)(148 * 8));
  k_grad_finalize<<<grid, 256, 0, s>>>(V_dim, nkeys, weights, V_pos, xxp, grad);
  return 1;
}

}  // namespace dfb
