// difacto_b200/csrc/kernels_localize.cu -- Localizer::Compact on the GPU.
//
// GPU restatement of src/data/localizer.cc:11-103 of the reference:
//   key_j = ReverseBytes(id_j % max_index)                (CountUniqIndex :22-26, base.h:39-51)
//   sort (key, position) pairs by key                      (:28-29; here a stable LSD radix sort)
//   unique keys ascending + occurrence counts              (:36-49)
//   index_j = rank of key_j among the unique keys          (RemapIndex :53-103)
// Integer work only: bit-exact by construction (tests compare with the oracle and the reference).
//
// The same sort also yields the CSC view of the batch (per key: the rows that contain it, in row
// order) that the atomic-free gradient kernel needs, so a raw-id step skips the separate CSC sort.
#include "dfb_device.cuh"

#include <cub/cub.cuh>

namespace dfb {

namespace {

__device__ __forceinline__ unsigned long long reverse_nibbles(unsigned long long x) {
  // byte reversal, then swap the two nibbles of every byte  == reversing the 16 nibbles
  x = ((unsigned long long)__byte_perm((unsigned)(x & 0xffffffffULL), 0, 0x0123) << 32) |
      (unsigned long long)__byte_perm((unsigned)(x >> 32), 0, 0x0123);
  return ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x & 0xF0F0F0F0F0F0F0F0ULL) >> 4);
}

// HI32: the caller knows that only the upper 32 bits of the reversed keys can be non-zero (ids < 2^32): the sort
// then runs on 32-bit keys (a third less traffic per radix pass); the OR mask still covers all 64 bits, so a wrong
// assumption is detected by k_emit like any other range violation
template <bool HI32>
__global__ void k_rev_keys(const uint64_t* __restrict__ ids, size_t n, unsigned long long max_index,
                           unsigned long long* __restrict__ rkeys, uint32_t* __restrict__ pos,
                           unsigned long long* __restrict__ or_all) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long k = 0;
  if (j < n) {
    unsigned long long x = ids[j];
    if (max_index == ~0ULL) x = (x == ~0ULL) ? 0ULL : x;   // x % (2^64-1)
    else x = x % max_index;
    k = reverse_nibbles(x);
    if (HI32) reinterpret_cast<uint32_t*>(rkeys)[j] = (uint32_t)(k >> 32);
    else rkeys[j] = k;
    pos[j] = (uint32_t)j;
  }
  // OR of all keys: the radix sort only needs the bit range that is not constant zero
  // (warp shuffle -> shared memory -> one atomic per block)
  __shared__ unsigned long long s_or[8];
  for (int o = 16; o > 0; o >>= 1) k |= __shfl_xor_sync(0xffffffffu, k, o);
  if ((threadIdx.x & 31) == 0) s_or[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a |= s_or[w];
    if (a) atomicOr(or_all, a);
  }
}

// row id of every nnz position (one warp per row, coalesced)
__global__ void k_expand_rows(const uint64_t* __restrict__ offset, size_t nrows, uint32_t* __restrict__ nnz_row) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp0; r < nrows; r += nwarps) {
    const uint64_t o0 = offset[r], o1 = offset[r + 1];
    for (uint64_t j = o0 + lane; j < o1; j += 32) nnz_row[j] = (uint32_t)r;
  }
}

template <typename KT>
__global__ void k_heads(const KT* __restrict__ skeys, size_t n, int* __restrict__ head) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  head[j] = (j == 0 || skeys[j] != skeys[j - 1]) ? 1 : 0;
}

// rank1 = inclusive scan of head.  Emits unique keys, per-key segment bounds, the remapped CSR index
// and the per-occurrence payload (row [, x]) in key-then-row order.
template <typename KT>
__global__ void k_emit(const KT* __restrict__ skeys, const uint32_t* __restrict__ spos,
                       const int* __restrict__ head, const int* __restrict__ rank1, size_t n,
                       const uint32_t* __restrict__ nnz_row, const float* __restrict__ value,
                       uint64_t* __restrict__ keys_out, int* __restrict__ col_start, int* __restrict__ col_end,
                       uint32_t* __restrict__ lidx_out, uint32_t* __restrict__ occ_row,
                       unsigned long long* __restrict__ occ_rowx, unsigned long long* __restrict__ scal,
                       int begin_bit, DevProgress* prog) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int r = rank1[j] - 1;
  const uint32_t p = spos[j];
  lidx_out[p] = (uint32_t)r;
  const uint32_t row = nnz_row[p];
  if (value) occ_rowx[j] = ((unsigned long long)row << 32) | (unsigned long long)__float_as_uint(value[p]);
  else occ_row[j] = row;
  if (head[j]) {
    keys_out[r] = sizeof(KT) == 4 ? ((uint64_t)skeys[j] << 32) : (uint64_t)skeys[j];
    col_start[r] = (int)j;
    if (j > 0) col_end[r - 1] = (int)j;
  }
  if (j == n - 1) {
    col_end[r] = (int)n;
    col_start[r + 1] = (int)n;      // col_start doubles as the U+1 column offsets of the CSC view
    // scal = {OR of all keys, number of unique keys}.  A key bit below the sorted range means the range was
    // assumed too narrow (id_bits / the range learned from earlier batches): the batch is NOT sorted, so it is
    // dropped (zero keys) and the error is reported instead of training on garbage.
    const unsigned long long low = begin_bit > 0 ? (scal[0] & ((1ULL << begin_bit) - 1ULL)) : 0ULL;
    if (low != 0ULL) {
      scal[1] = 0ULL;
      if (prog) raise_err(prog, DFB_ERR_INVALID);
    } else {
      scal[1] = (unsigned long long)(r + 1);
    }
  }
}

__global__ void k_cnt_from_cols(const int* __restrict__ col_start, const int* __restrict__ col_end, size_t n,
                                float* __restrict__ cnt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cnt[i] = (float)(col_end[i] - col_start[i]);   // idx_frq, localizer.cc:41-46
}

}  // namespace

size_t localize_sort_tmp_bytes(size_t nnz) {
  size_t b1 = 0, b2 = 0;
  const int n = (int)(nnz ? nnz : 1);
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
  size_t b3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b3, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, n);
  if (b3 > b1) b1 = b3;
  cub::DeviceScan::InclusiveSum(nullptr, b2, (const int*)nullptr, (int*)nullptr, n);
  return b1 > b2 ? b1 : b2;
}

int launch_localize_keys(const uint64_t* ids, size_t nnz, uint64_t max_index, unsigned long long* rkeys,
                         uint32_t* pos, unsigned long long* or_all, const uint64_t* offset, size_t nrows,
                         uint32_t* nnz_row, bool hi32, cudaStream_t s) {
  cudaMemsetAsync(or_all, 0, sizeof(unsigned long long), s);
  if (nnz == 0) return 0;
  if (hi32) k_rev_keys<true><<<(int)((nnz + 255) / 256), 256, 0, s>>>(ids, nnz, max_index, rkeys, pos, or_all);
  else      k_rev_keys<false><<<(int)((nnz + 255) / 256), 256, 0, s>>>(ids, nnz, max_index, rkeys, pos, or_all);
  size_t need = (nrows + 7) / 8;
  int grid = (int)(need < (size_t)(148 * 8) ? (need ? need : 1) : (size_t)(148 * 8));
  k_expand_rows<<<grid, 256, 0, s>>>(offset, nrows, nnz_row);
  return 2;
}

// hi32 must be what launch_localize_keys was given (it implies begin_bit >= 32)
int launch_localize_sort(const unsigned long long* rkeys, const uint32_t* pos, size_t nnz, int begin_bit, bool hi32,
                         unsigned long long* skeys, uint32_t* spos, int* head, int* rank1, void* tmp, size_t tmp_bytes,
                         const uint32_t* nnz_row, const float* value, uint64_t* keys_out, int* col_start,
                         int* col_end, uint32_t* lidx_out, void* occ_sorted, unsigned long long* scal,
                         DevProgress* prog, cudaStream_t s) {
  cudaMemsetAsync(scal + 1, 0, sizeof(unsigned long long), s);
  if (nnz == 0) return 0;
  const int grid = (int)((nnz + 255) / 256);
  if (hi32) {
    const uint32_t* rk = reinterpret_cast<const uint32_t*>(rkeys);
    uint32_t* sk = reinterpret_cast<uint32_t*>(skeys);
    const int bb = begin_bit >= 32 ? begin_bit - 32 : 0;
    if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, rk, sk, pos, spos, (int)nnz, bb, 32, s) != cudaSuccess) return -1;
    k_heads<uint32_t><<<grid, 256, 0, s>>>(sk, nnz, head);
    if (cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, rank1, (int)nnz, s) != cudaSuccess) return -1;
    k_emit<uint32_t><<<grid, 256, 0, s>>>(sk, spos, head, rank1, nnz, nnz_row, value, keys_out, col_start, col_end, lidx_out,
                                          reinterpret_cast<uint32_t*>(occ_sorted),
                                          reinterpret_cast<unsigned long long*>(occ_sorted), scal,
                                          begin_bit >= 32 ? begin_bit : 32, prog);
    return 5 + (32 - bb + 7) / 8;
  }
  if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, rkeys, skeys, pos, spos, (int)nnz, begin_bit, 64, s) != cudaSuccess) return -1;
  k_heads<unsigned long long><<<grid, 256, 0, s>>>(skeys, nnz, head);
  if (cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, rank1, (int)nnz, s) != cudaSuccess) return -1;
  k_emit<unsigned long long><<<grid, 256, 0, s>>>(skeys, spos, head, rank1, nnz, nnz_row, value, keys_out, col_start, col_end,
                                                  lidx_out, reinterpret_cast<uint32_t*>(occ_sorted),
                                                  reinterpret_cast<unsigned long long*>(occ_sorted), scal, begin_bit, prog);
  return 5 + (64 - begin_bit + 7) / 8;
}

int launch_cnt_from_cols(const int* col_start, const int* col_end, size_t n, float* cnt, cudaStream_t s) {
  if (n == 0) return 0;
  k_cnt_from_cols<<<(int)((n + 255) / 256), 256, 0, s>>>(col_start, col_end, n, cnt);
  return 1;
}

}  // namespace dfb
