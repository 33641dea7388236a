// setup_gpu.cu — SURVEY.md 8(f1): the sparse products of the AMG setup (A*P, R*(A*P), A_F*P0) on the GPU.
//
// C = A * B by expand / sort / compress, arranged so that every output entry is summed in EXACTLY the order the host
// (host_setup.cpp spgemm) and the oracle (amg_oracle.spgemm_structural) use: the products a_ik * b_kj that land on
// column j of row i are added in encounter order (k ascending along row i of A, then along row k of B).
//   1. expand   one thread per A entry: writes its products (key = chunk-local row << 32 | column, value = a*b rounded
//               once) at the positions an exclusive scan of the row-of-B lengths assigns -> encounter order = position;
//   2. sort     stable LSD radix sort of (key, value) pairs (cub::DeviceRadixSort, library code: the only non-hand-written
//               step; stability keeps the encounter order inside runs of equal keys);
//   3. compress run heads -> exclusive scan -> one thread per run adds its values sequentially (no FMA: the products
//               were rounded in step 1), writes (column, sum); rows are counted from the heads.
// Structure (sorted columns, explicit zeros kept) and values are therefore bit-identical to the host product.
// Rows are processed in chunks bounded by a product budget, so the scratch memory stays a few GB at any size.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "host.hpp"

namespace pamg {

namespace {

#define GK(call)                                                                                                  \
  do {                                                                                                            \
    cudaError_t e_ = (call);                                                                                      \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("gpu setup: ") + #call + " -> " + cudaGetErrorString(e_)); \
  } while (0)

template <class T>
struct Buf {
  T* p = nullptr;
  size_t n = 0;
  ~Buf() {
    if (p) cudaFree(p);
  }
  void alloc(size_t count) {
    if (count <= n) return;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    GK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void upload(const T* h, size_t count) {
    alloc(count);
    if (count) GK(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice));
  }
};

constexpr int TB = 256;

// products of A entry e: len(B row col_A[e]); rows [r0, r1) of A only
__global__ void k_count(const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col, const int64_t* __restrict__ b_ptr,
                        int64_t e0, int64_t ne, int64_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= ne) return;
  const int32_t k = a_col[e0 + i];
  cnt[i] = b_ptr[k + 1] - b_ptr[k];
}

// row of every A entry in [e0, e0 + ne) (binary search in a_ptr), then its products at off[i] ...
__global__ void k_expand(const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col, const double* __restrict__ a_val,
                         const int64_t* __restrict__ b_ptr, const int32_t* __restrict__ b_col, const double* __restrict__ b_val,
                         int64_t r0, int64_t r1, int64_t e0, int64_t ne, const int64_t* __restrict__ off,
                         unsigned long long* __restrict__ keys, double* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= ne) return;
  const int64_t e = e0 + i;
  int64_t lo = r0, hi = r1;  // last row with a_ptr[row] <= e
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (a_ptr[mid] <= e)
      lo = mid;
    else
      hi = mid;
  }
  const unsigned long long rowkey = (unsigned long long)(lo - r0) << 32;
  const int32_t k = a_col[e];
  const double a = a_val[e];
  int64_t q = off[i];
  for (int64_t kb = b_ptr[k]; kb < b_ptr[k + 1]; ++kb, ++q) {
    keys[q] = rowkey | (unsigned long long)(uint32_t)b_col[kb];
    vals[q] = __dmul_rn(a, b_val[kb]);
  }
}

__global__ void k_heads(const unsigned long long* __restrict__ keys, int64_t n, int64_t* __restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per run: sequential sum in sorted (= encounter) order; also counts the entries of each row
__global__ void k_compress(const unsigned long long* __restrict__ keys, const double* __restrict__ vals, int64_t n,
                           const int64_t* __restrict__ head, const int64_t* __restrict__ pos, int32_t* __restrict__ out_col,
                           double* __restrict__ out_val, unsigned long long* __restrict__ row_cnt) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n || !head[i]) return;
  const unsigned long long key = keys[i];
  double s = vals[i];
  for (int64_t j = i + 1; j < n && keys[j] == key; ++j) s = __dadd_rn(s, vals[j]);
  const int64_t o = pos[i];
  out_col[o] = (int32_t)(uint32_t)(key & 0xffffffffull);
  out_val[o] = s;
  atomicAdd(row_cnt + (key >> 32), 1ull);
}

int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

}  // namespace

bool gpu_setup_available() {
  const char* e = getenv("PAMG_GPU_SETUP");
  if (e && atoi(e) == 0) return false;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return false;
  }
  return true;
}

void gpu_spgemm(const Csr& A, const Csr& B, Csr& C) {
  if (A.ncols != B.nrows) throw std::runtime_error("gpu_spgemm: shape mismatch");
  if (B.ncols >= (int64_t)1 << 31 || A.ncols >= (int64_t)1 << 31) throw std::runtime_error("gpu_spgemm: more than 2^31 columns");
  const int64_t n = A.nrows;
  C.nrows = n;
  C.ncols = B.ncols;
  C.ptr.assign(n + 1, 0);
  C.col.clear();
  C.val.clear();
  // device copies (column ids as int32)
  std::vector<int32_t> a_col32(A.col.begin(), A.col.end()), b_col32(B.col.begin(), B.col.end());
  Buf<int64_t> dA_ptr, dB_ptr;
  Buf<int32_t> dA_col, dB_col;
  Buf<double> dA_val, dB_val;
  dA_ptr.upload(A.ptr.data(), A.ptr.size());
  dA_col.upload(a_col32.data(), a_col32.size());
  dA_val.upload(A.val.data(), A.val.size());
  dB_ptr.upload(B.ptr.data(), B.ptr.size());
  dB_col.upload(b_col32.data(), b_col32.size());
  dB_val.upload(B.val.data(), B.val.size());
  std::vector<int32_t>().swap(a_col32);
  std::vector<int32_t>().swap(b_col32);

  // products per row on the host (cheap, also drives the chunking)
  std::vector<int64_t> row_prod(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int64_t s = 0;
    for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; ++k) s += B.ptr[A.col[k] + 1] - B.ptr[A.col[k]];
    row_prod[i + 1] = s;
  }
  for (int64_t i = 0; i < n; ++i) row_prod[i + 1] += row_prod[i];
  const char* be = getenv("PAMG_GPU_SETUP_BUDGET");
  const int64_t budget = be ? std::max<int64_t>(1 << 16, atoll(be)) : ((int64_t)384 << 20);  // products per chunk

  Buf<int64_t> d_cnt, d_off, d_head, d_pos;
  Buf<unsigned long long> d_keys, d_keys2, d_rowcnt;
  Buf<double> d_vals, d_vals2, d_oval;
  Buf<int32_t> d_ocol;
  Buf<char> d_tmp;
  std::vector<unsigned long long> h_rowcnt;
  std::vector<int32_t> h_col;
  std::vector<double> h_val;
  std::vector<std::vector<int32_t>> chunk_col;
  std::vector<std::vector<double>> chunk_val;

  int64_t r0 = 0;
  while (r0 < n) {
    int64_t r1 = r0 + 1;
    while (r1 < n && row_prod[r1 + 1] - row_prod[r0] <= budget) ++r1;
    const int64_t T = row_prod[r1] - row_prod[r0];
    if (T >= (int64_t)1 << 40) throw std::runtime_error("gpu_spgemm: a single row produces too many products");
    const int64_t e0 = A.ptr[r0], ne = A.ptr[r1] - e0, nr = r1 - r0;
    std::vector<int32_t> ccol;
    std::vector<double> cval;
    h_rowcnt.assign(nr, 0);
    if (T > 0) {
      d_cnt.alloc(ne + 1);
      d_off.alloc(ne + 1);
      d_keys.alloc(T);
      d_keys2.alloc(T);
      d_vals.alloc(T);
      d_vals2.alloc(T);
      d_head.alloc(T);
      d_pos.alloc(T);
      d_rowcnt.alloc(nr);
      GK(cudaMemset(d_rowcnt.p, 0, nr * sizeof(unsigned long long)));
      const int g_e = (int)((ne + TB - 1) / TB), g_t = (int)((T + TB - 1) / TB);
      k_count<<<g_e, TB>>>(dA_ptr.p, dA_col.p, dB_ptr.p, e0, ne, d_cnt.p);
      size_t tmp_bytes = 0, need = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, need, d_cnt.p, d_off.p, ne);
      tmp_bytes = need;
      cub::DeviceScan::ExclusiveSum(nullptr, need, d_head.p, d_pos.p, T);
      tmp_bytes = std::max(tmp_bytes, need);
      const int end_bit = 32 + bits_for((uint64_t)std::max<int64_t>(nr - 1, 1));
      cub::DeviceRadixSort::SortPairs(nullptr, need, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, T, 0, end_bit);
      tmp_bytes = std::max(tmp_bytes, need);
      d_tmp.alloc(tmp_bytes);
      size_t tb = d_tmp.n;
      GK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_cnt.p, d_off.p, ne));
      k_expand<<<g_e, TB>>>(dA_ptr.p, dA_col.p, dA_val.p, dB_ptr.p, dB_col.p, dB_val.p, r0, r1, e0, ne, d_off.p, d_keys.p, d_vals.p);
      tb = d_tmp.n;
      GK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, T, 0, end_bit));
      k_heads<<<g_t, TB>>>(d_keys2.p, T, d_head.p);
      tb = d_tmp.n;
      GK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_head.p, d_pos.p, T));
      int64_t last_pos = 0, last_head = 0;
      GK(cudaMemcpy(&last_pos, d_pos.p + (T - 1), sizeof(int64_t), cudaMemcpyDeviceToHost));
      GK(cudaMemcpy(&last_head, d_head.p + (T - 1), sizeof(int64_t), cudaMemcpyDeviceToHost));
      const int64_t nout = last_pos + last_head;
      d_ocol.alloc(nout);
      d_oval.alloc(nout);
      k_compress<<<g_t, TB>>>(d_keys2.p, d_vals2.p, T, d_head.p, d_pos.p, d_ocol.p, d_oval.p, d_rowcnt.p);
      GK(cudaGetLastError());
      ccol.resize(nout);
      cval.resize(nout);
      GK(cudaMemcpy(ccol.data(), d_ocol.p, nout * sizeof(int32_t), cudaMemcpyDeviceToHost));
      GK(cudaMemcpy(cval.data(), d_oval.p, nout * sizeof(double), cudaMemcpyDeviceToHost));
      GK(cudaMemcpy(h_rowcnt.data(), d_rowcnt.p, nr * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    for (int64_t i = 0; i < nr; ++i) C.ptr[r0 + i + 1] = (int64_t)h_rowcnt[i];
    chunk_col.push_back(std::move(ccol));
    chunk_val.push_back(std::move(cval));
    r0 = r1;
  }
  for (int64_t i = 0; i < n; ++i) C.ptr[i + 1] += C.ptr[i];
  C.col.resize(C.ptr[n]);
  C.val.resize(C.ptr[n]);
  int64_t q = 0;
  for (size_t k = 0; k < chunk_col.size(); ++k) {
    const int64_t m = (int64_t)chunk_col[k].size();
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; ++j) {
      C.col[q + j] = chunk_col[k][j];
      C.val[q + j] = chunk_val[k][j];
    }
    q += m;
    std::vector<int32_t>().swap(chunk_col[k]);
    std::vector<double>().swap(chunk_val[k]);
  }
  if (q != C.ptr[n]) throw std::runtime_error("gpu_spgemm: internal count mismatch");
}

}  // namespace pamg
