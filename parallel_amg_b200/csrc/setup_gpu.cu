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
#include <cstring>
#include <memory>
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
  void upload(const T* h, size_t count);
  void download(T* h, size_t count) const;
};

// Host arrays of the setup are pageable std::vectors (a few GB at 256^3): a plain cudaMemcpy moves them at 3-4 GB/s.
// Staged through two pinned 64 MB buffers (parallel memcpy into one while the DMA drains the other) they move at PCIe speed.
struct Staging {
  static constexpr size_t BYTES = (size_t)64 << 20;
  char* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t st = nullptr;
  bool ok = false;
  Staging() {
    if (cudaHostAlloc((void**)&buf[0], BYTES, cudaHostAllocDefault) != cudaSuccess) return;
    if (cudaHostAlloc((void**)&buf[1], BYTES, cudaHostAllocDefault) != cudaSuccess) return;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess) return;
    ok = true;
  }
  ~Staging() {
    for (int k = 0; k < 2; ++k) {
      if (ev[k]) cudaEventDestroy(ev[k]);
      if (buf[k]) cudaFreeHost(buf[k]);
    }
    if (st) cudaStreamDestroy(st);
  }
};
Staging*& staging_slot() {
  static Staging* s = nullptr;
  return s;
}
void par_copy(char* dst, const char* src, size_t bytes) {
  const int64_t nb = (int64_t)((bytes + 4095) / 4096);
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < nb; ++b) {
    const size_t o = (size_t)b * 4096;
    std::memcpy(dst + o, src + o, std::min<size_t>(4096, bytes - o));
  }
}
void staged_h2d(void* dev, const void* host, size_t bytes) {
  Staging* S = staging_slot();
  if (!S || !S->ok || bytes < ((size_t)8 << 20)) {
    GK(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
    return;
  }
  GK(cudaDeviceSynchronize());  // the device buffer may have been produced on the default stream
  int k = 0;
  for (size_t o = 0; o < bytes; o += Staging::BYTES, k ^= 1) {
    const size_t m = std::min(Staging::BYTES, bytes - o);
    GK(cudaEventSynchronize(S->ev[k]));  // the DMA that last used this buffer is done
    par_copy(S->buf[k], (const char*)host + o, m);
    GK(cudaMemcpyAsync((char*)dev + o, S->buf[k], m, cudaMemcpyHostToDevice, S->st));
    GK(cudaEventRecord(S->ev[k], S->st));
  }
  GK(cudaStreamSynchronize(S->st));
}
void staged_d2h(void* host, const void* dev, size_t bytes) {
  Staging* S = staging_slot();
  if (!S || !S->ok || bytes < ((size_t)8 << 20)) {
    GK(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
    return;
  }
  GK(cudaDeviceSynchronize());
  // chunk c is copied device -> pinned[c & 1] while chunk c - 1 is copied pinned -> pageable on the host
  const size_t nchunks = (bytes + Staging::BYTES - 1) / Staging::BYTES;
  for (size_t c = 0; c <= nchunks; ++c) {
    if (c < nchunks) {
      const size_t o = c * Staging::BYTES, m = std::min(Staging::BYTES, bytes - o);
      GK(cudaMemcpyAsync(S->buf[c & 1], (const char*)dev + o, m, cudaMemcpyDeviceToHost, S->st));
      GK(cudaEventRecord(S->ev[c & 1], S->st));
    }
    if (c > 0) {
      const size_t o = (c - 1) * Staging::BYTES, m = std::min(Staging::BYTES, bytes - o);
      GK(cudaEventSynchronize(S->ev[(c - 1) & 1]));
      par_copy((char*)host + o, S->buf[(c - 1) & 1], m);
    }
  }
}
template <class T>
void Buf<T>::upload(const T* h, size_t count) {
  alloc(count);
  if (count) staged_h2d(p, h, count * sizeof(T));
}
template <class T>
void Buf<T>::download(T* h, size_t count) const {
  if (count) staged_d2h(h, p, count * sizeof(T));
}

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

void gpu_setup_begin() {
  if (!staging_slot()) staging_slot() = new Staging;
}
void gpu_setup_end() {
  delete staging_slot();
  staging_slot() = nullptr;
}

bool gpu_setup_available() {
  // on whenever a device is there (PAMG_GPU_SETUP=0: host products).  256^3 on the GPU box: 6.1 s with the device chain
  // (smoothing, transpose, A*P, R*(AP) resident; pinned staging) vs 9.6 s on 16 host cores, profiles/r02_setup_timing.txt
  const char* e = getenv("PAMG_GPU_SETUP");
  if (e && atoi(e) == 0) return false;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return false;
  }
  return true;
}

// device-resident CSR of the setup chain: int64 row pointers, int32 columns, fp64 values
struct GpuMat {
  int64_t nrows = 0, ncols = 0, nnz = 0;
  Buf<int64_t> ptr;
  Buf<int32_t> col;
  Buf<double> val;
};

GpuMat* gpu_upload(const Csr& A) {
  if (A.ncols >= (int64_t)1 << 31) throw std::runtime_error("gpu setup: more than 2^31 columns");
  std::unique_ptr<GpuMat> m(new GpuMat);
  m->nrows = A.nrows;
  m->ncols = A.ncols;
  m->nnz = A.nnz();
  std::vector<int32_t> col32;
  huge_reserve(col32, A.col.size());
  col32.resize(A.col.size());
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)A.col.size(); ++k) col32[k] = (int32_t)A.col[k];
  m->ptr.upload(A.ptr.data(), A.ptr.size());
  m->col.upload(col32.data(), col32.size());
  m->val.upload(A.val.data(), A.val.size());
  return m.release();
}

void gpu_download(const GpuMat* m, Csr& C) {
  C.nrows = m->nrows;
  C.ncols = m->ncols;
  huge_reserve(C.ptr, (size_t)m->nrows + 1);
  huge_reserve(C.col, (size_t)m->nnz);
  huge_reserve(C.val, (size_t)m->nnz);
  C.ptr.resize(m->nrows + 1);
  C.col.resize(m->nnz);
  C.val.resize(m->nnz);
  std::vector<int32_t> col32;
  huge_reserve(col32, (size_t)m->nnz);
  col32.resize(m->nnz);
  m->ptr.download(C.ptr.data(), (size_t)m->nrows + 1);
  if (m->nnz) {
    m->col.download(col32.data(), (size_t)m->nnz);
    m->val.download(C.val.data(), (size_t)m->nnz);
  }
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < m->nnz; ++k) C.col[k] = col32[k];
}

void gpu_free(GpuMat* m) { delete m; }

__global__ void k_row_products(const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_col, const int64_t* __restrict__ b_ptr,
                               int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  int64_t s = 0;
  for (int64_t k = a_ptr[i]; k < a_ptr[i + 1]; ++k) s += b_ptr[a_col[k] + 1] - b_ptr[a_col[k]];
  out[i] = s;
}
__global__ void k_set_u64_to_i64(const unsigned long long* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i];
}

// C = A * B, everything on the device
GpuMat* gpu_product(const GpuMat* A, const GpuMat* B) {
  if (A->ncols != B->nrows) throw std::runtime_error("gpu_product: shape mismatch");
  const int64_t n = A->nrows;
  std::unique_ptr<GpuMat> C(new GpuMat);
  C->nrows = n;
  C->ncols = B->ncols;
  // products per row -> chunking (host copy of the per-row counts: 8 bytes per row)
  Buf<int64_t> d_rp;
  d_rp.alloc(n + 1);
  std::vector<int64_t> row_prod(n + 1, 0);
  if (n) {
    k_row_products<<<(int)((n + TB - 1) / TB), TB>>>(A->ptr.p, A->col.p, B->ptr.p, n, d_rp.p);
    GK(cudaMemcpy(row_prod.data() + 1, d_rp.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  for (int64_t i = 0; i < n; ++i) row_prod[i + 1] += row_prod[i];
  std::vector<int64_t> a_ptr(n + 1);
  GK(cudaMemcpy(a_ptr.data(), A->ptr.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  const char* be = getenv("PAMG_GPU_SETUP_BUDGET");
  const int64_t budget = be ? std::max<int64_t>(1 << 16, atoll(be)) : ((int64_t)384 << 20);  // products per chunk

  Buf<int64_t> d_cnt, d_off, d_head, d_pos;
  Buf<unsigned long long> d_keys, d_keys2, d_rowcnt;
  Buf<double> d_vals, d_vals2;
  Buf<char> d_tmp;
  d_rowcnt.alloc(std::max<int64_t>(n, 1));
  GK(cudaMemset(d_rowcnt.p, 0, std::max<int64_t>(n, 1) * sizeof(unsigned long long)));
  struct Chunk {
    Buf<int32_t> col;
    Buf<double> val;
    int64_t n = 0;
  };
  std::vector<std::unique_ptr<Chunk>> chunks;
  int64_t r0 = 0, total = 0;
  while (r0 < n) {
    int64_t r1 = r0 + 1;
    while (r1 < n && row_prod[r1 + 1] - row_prod[r0] <= budget) ++r1;
    const int64_t T = row_prod[r1] - row_prod[r0];
    if (T >= (int64_t)1 << 40) throw std::runtime_error("gpu_product: a single row produces too many products");
    const int64_t e0 = a_ptr[r0], ne = a_ptr[r1] - e0, nr = r1 - r0;
    std::unique_ptr<Chunk> ch(new Chunk);
    if (T > 0) {
      d_cnt.alloc(ne + 1);
      d_off.alloc(ne + 1);
      d_keys.alloc(T);
      d_keys2.alloc(T);
      d_vals.alloc(T);
      d_vals2.alloc(T);
      d_head.alloc(T);
      d_pos.alloc(T);
      const int g_e = (int)((ne + TB - 1) / TB), g_t = (int)((T + TB - 1) / TB);
      k_count<<<g_e, TB>>>(A->ptr.p, A->col.p, B->ptr.p, e0, ne, d_cnt.p);
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
      k_expand<<<g_e, TB>>>(A->ptr.p, A->col.p, A->val.p, B->ptr.p, B->col.p, B->val.p, r0, r1, e0, ne, d_off.p, d_keys.p, d_vals.p);
      tb = d_tmp.n;
      GK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, T, 0, end_bit));
      k_heads<<<g_t, TB>>>(d_keys2.p, T, d_head.p);
      tb = d_tmp.n;
      GK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_head.p, d_pos.p, T));
      int64_t last_pos = 0, last_head = 0;
      GK(cudaMemcpy(&last_pos, d_pos.p + (T - 1), sizeof(int64_t), cudaMemcpyDeviceToHost));
      GK(cudaMemcpy(&last_head, d_head.p + (T - 1), sizeof(int64_t), cudaMemcpyDeviceToHost));
      ch->n = last_pos + last_head;
      ch->col.alloc(ch->n);
      ch->val.alloc(ch->n);
      k_compress<<<g_t, TB>>>(d_keys2.p, d_vals2.p, T, d_head.p, d_pos.p, ch->col.p, ch->val.p, d_rowcnt.p + r0);
      GK(cudaGetLastError());
    }
    total += ch->n;
    chunks.push_back(std::move(ch));
    r0 = r1;
  }
  // row pointers from the per-row counts, then the chunks back to back
  C->nnz = total;
  C->ptr.alloc(n + 1);
  C->col.alloc(total);
  C->val.alloc(total);
  {
    Buf<int64_t> d_c64;
    d_c64.alloc(n + 1);
    GK(cudaMemset(d_c64.p, 0, (n + 1) * sizeof(int64_t)));
    if (n) k_set_u64_to_i64<<<(int)((n + TB - 1) / TB), TB>>>(d_rowcnt.p, n, d_c64.p);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_c64.p, C->ptr.p, n + 1);
    d_tmp.alloc(need);
    size_t tb = d_tmp.n;
    GK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_c64.p, C->ptr.p, n + 1));
  }
  int64_t q = 0;
  for (auto& ch : chunks) {
    if (ch->n) {
      GK(cudaMemcpy(C->col.p + q, ch->col.p, ch->n * sizeof(int32_t), cudaMemcpyDeviceToDevice));
      GK(cudaMemcpy(C->val.p + q, ch->val.p, ch->n * sizeof(double), cudaMemcpyDeviceToDevice));
    }
    q += ch->n;
  }
  GK(cudaDeviceSynchronize());
  return C.release();
}

// ---- prolongator smoothing on the device: P = P0 - omega D^-1 S, S = A_F P0, merged by column (host_setup.cpp build_prolongator)
__global__ void k_merge_count(const int64_t* __restrict__ s_ptr, const int32_t* __restrict__ s_col, const int64_t* __restrict__ p_ptr,
                              const int32_t* __restrict__ p_col, int64_t n, int64_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  int64_t pk = p_ptr[i], c = 0;
  const int64_t pe = p_ptr[i + 1];
  for (int64_t k = s_ptr[i]; k < s_ptr[i + 1]; ++k) {
    const int32_t col = s_col[k];
    while (pk < pe && p_col[pk] < col) {
      ++pk;
      ++c;
    }
    if (pk < pe && p_col[pk] == col) ++pk;
    ++c;
  }
  cnt[i] = c + (pe - pk);
}
__global__ void k_merge_fill(const int64_t* __restrict__ s_ptr, const int32_t* __restrict__ s_col, const double* __restrict__ s_val,
                             const int64_t* __restrict__ p_ptr, const int32_t* __restrict__ p_col, const double* __restrict__ p_val,
                             const double* __restrict__ w, int64_t n, const int64_t* __restrict__ o_ptr, int32_t* __restrict__ o_col,
                             double* __restrict__ o_val) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  int64_t pk = p_ptr[i], q = o_ptr[i];
  const int64_t pe = p_ptr[i + 1];
  const double wi = w[i];
  for (int64_t k = s_ptr[i]; k < s_ptr[i + 1]; ++k) {
    const int32_t col = s_col[k];
    while (pk < pe && p_col[pk] < col) {
      o_col[q] = p_col[pk];
      o_val[q++] = p_val[pk++];
    }
    double v = __dmul_rn(wi, s_val[k]);  // rounded product, then a rounded add: the host's arithmetic, no FMA
    if (pk < pe && p_col[pk] == col) v = __dadd_rn(p_val[pk++], v);
    o_col[q] = col;
    o_val[q++] = v;
  }
  while (pk < pe) {
    o_col[q] = p_col[pk];
    o_val[q++] = p_val[pk++];
  }
}

void exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, Buf<char>& tmp) {
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n);
  tmp.alloc(need);
  size_t tb = tmp.n;
  GK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n));
}

// P = P0 + diag(w) (F P0) with w_i = -(omega / d_F,i); F and P0 resident, w from the host
GpuMat* gpu_smooth_prolongator(const GpuMat* F, const GpuMat* P0, const double* w_host) {
  std::unique_ptr<GpuMat> S(gpu_product(F, P0));
  const int64_t n = F->nrows;
  std::unique_ptr<GpuMat> P(new GpuMat);
  P->nrows = n;
  P->ncols = P0->ncols;
  Buf<double> d_w;
  d_w.upload(w_host, (size_t)n);
  Buf<int64_t> d_cnt;
  Buf<char> tmp;
  d_cnt.alloc(n + 1);
  GK(cudaMemset(d_cnt.p, 0, (n + 1) * sizeof(int64_t)));
  const int g = (int)((n + TB - 1) / TB);
  if (n) k_merge_count<<<g, TB>>>(S->ptr.p, S->col.p, P0->ptr.p, P0->col.p, n, d_cnt.p);
  P->ptr.alloc(n + 1);
  exclusive_scan_i64(d_cnt.p, P->ptr.p, n + 1, tmp);
  GK(cudaMemcpy(&P->nnz, P->ptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost));
  P->col.alloc(P->nnz);
  P->val.alloc(P->nnz);
  if (n) k_merge_fill<<<g, TB>>>(S->ptr.p, S->col.p, S->val.p, P0->ptr.p, P0->col.p, P0->val.p, d_w.p, n, P->ptr.p, P->col.p, P->val.p);
  GK(cudaGetLastError());
  GK(cudaDeviceSynchronize());
  return P.release();
}

// ---- R = P^T on the device: keys (column << 32 | row) are unique, so sorting them IS the transpose with ascending columns
__global__ void k_tr_keys(const int64_t* __restrict__ ptr, const int32_t* __restrict__ col, int64_t n, unsigned long long* __restrict__ keys,
                          unsigned long long* __restrict__ col_cnt) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) {
    keys[k] = ((unsigned long long)(uint32_t)col[k] << 32) | (unsigned long long)(uint32_t)i;
    atomicAdd(col_cnt + col[k], 1ull);
  }
}
__global__ void k_tr_unpack(const unsigned long long* __restrict__ keys, int64_t nnz, int32_t* __restrict__ out_col) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i < nnz) out_col[i] = (int32_t)(uint32_t)(keys[i] & 0xffffffffull);
}
GpuMat* gpu_transpose(const GpuMat* A) {
  std::unique_ptr<GpuMat> T(new GpuMat);
  T->nrows = A->ncols;
  T->ncols = A->nrows;
  T->nnz = A->nnz;
  const int64_t n = A->nrows, nc = A->ncols, nnz = A->nnz;
  Buf<unsigned long long> keys, keys2, cnt;
  Buf<int64_t> cnt64;
  Buf<char> tmp;
  keys.alloc(nnz);
  keys2.alloc(nnz);
  cnt.alloc(nc + 1);
  GK(cudaMemset(cnt.p, 0, (nc + 1) * sizeof(unsigned long long)));
  if (n) k_tr_keys<<<(int)((n + TB - 1) / TB), TB>>>(A->ptr.p, A->col.p, n, keys.p, cnt.p);
  T->val.alloc(nnz);
  T->col.alloc(nnz);
  if (nnz) {
    size_t need = 0;
    const int end_bit = 32 + bits_for((uint64_t)std::max<int64_t>(nc - 1, 1));
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys.p, keys2.p, A->val.p, T->val.p, nnz, 0, end_bit);
    tmp.alloc(need);
    size_t tb = tmp.n;
    GK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, A->val.p, T->val.p, nnz, 0, end_bit));
    k_tr_unpack<<<(int)((nnz + TB - 1) / TB), TB>>>(keys2.p, nnz, T->col.p);
  }
  cnt64.alloc(nc + 1);
  k_set_u64_to_i64<<<(int)((nc + 1 + TB - 1) / TB), TB>>>(cnt.p, nc + 1, cnt64.p);
  T->ptr.alloc(nc + 1);
  exclusive_scan_i64(cnt64.p, T->ptr.p, nc + 1, tmp);
  GK(cudaGetLastError());
  GK(cudaDeviceSynchronize());
  return T.release();
}

// ------------------------------------------------------------------------------------------------
// the greedy three-pass aggregation on the GPU, bit-identical to host_setup.cpp aggregate_part
// ------------------------------------------------------------------------------------------------
// The host algorithm walks the rows of a part in ascending order; what it decides for row i depends only on rows before
// i.  Restated without the walk (S = strong own-own neighbours, symmetric; all parts at once: no strong edge joins two
// parts, and inside a part ascending global id IS ascending local id):
//   pass 1  i is a ROOT  <=>  no ROOT r < i within distance 2 of i in S   (r's closed neighbourhood already holds one
//           of i's).  That is the lexicographically first maximal independent set of S^2.  Fixed point: an undecided i
//           becomes NOT as soon as one earlier vertex within distance 2 is a ROOT, ROOT as soon as all of them are NOT.
//           Every decision rests on final facts only, so any evaluation order (rounds, stale reads) gives the same set.
//           A blocked vertex remembers its LARGEST undecided earlier neighbour and is not rescanned until that one is
//           decided (the front moves in id order, so the largest one is the last to go).
//           aggregate id = rank of the root among the roots of its part (exclusive scan in part-major order);
//           a non-root takes the id of the root it is adjacent to (closed neighbourhoods of roots are disjoint).
//   pass 2  a vertex pass 1 left out joins the first neighbour (ascending column) that pass 1 aggregated.
//   pass 3  is empty when S is symmetric: the root set is MAXIMAL in S^2, so every vertex has a root within distance 2 --
//           at distance 1 pass 1 aggregated it, at distance 2 one of its neighbours is adjacent to that root and pass 2
//           picks it up (isolated vertices are roots).  The device checks that nothing is left and that S is symmetric;
//           otherwise (a non-symmetric operator, values within rounding of the threshold) the host walks the rows.
// The strength test repeats the host's arithmetic (|a_ij| > eps sqrt(|a_ii| |a_jj|), every operation rounded once).
namespace {

enum : unsigned char { ST_U = 0, ST_ROOT = 1, ST_NOT = 2 };

__device__ __forceinline__ unsigned char ld_state(const unsigned char* p) { return __ldcg(p); }

__global__ void k_absdiag(const int64_t* __restrict__ ptr, const int32_t* __restrict__ col, const double* __restrict__ val, int64_t n,
                          double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  double d = 0.0;
  for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
    if (col[k] == i) d = val[k];
  out[i] = fabs(d);
}

__device__ __forceinline__ bool is_strong(int64_t i, int32_t j, double v, const int32_t* __restrict__ owner, double eps,
                                          const double* __restrict__ ad) {
  if ((int64_t)j == i || owner[j] != owner[i]) return false;
  if (eps <= 0.0) return true;
  return fabs(v) > __dmul_rn(eps, __dsqrt_rn(__dmul_rn(ad[i], ad[j])));
}

// FILL = false: cnt[i] = strong neighbours of row i ; FILL = true: their columns (ascending, A's order) at s_ptr[i]
template <bool FILL>
__global__ void k_strong(const int64_t* __restrict__ ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                         const int32_t* __restrict__ owner, double eps, const double* __restrict__ ad, int64_t n,
                         int64_t* __restrict__ cnt, const int64_t* __restrict__ s_ptr, int32_t* __restrict__ s_col) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  int64_t q = FILL ? s_ptr[i] : 0;
  for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
    if (is_strong(i, col[k], val[k], owner, eps, ad)) {
      if (FILL) s_col[q] = col[k];
      ++q;
    }
  if (!FILL) cnt[i] = q;
}

// every strong edge i -> j must have its twin j -> i (binary search in row j)
__global__ void k_sym_check(const int64_t* __restrict__ sp, const int32_t* __restrict__ sc, int64_t n, unsigned long long* bad) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (i >= n) return;
  for (int64_t q = sp[i]; q < sp[i + 1]; ++q) {
    const int32_t j = sc[q];
    int64_t lo = sp[j], hi = sp[j + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)sc[mid] < i)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo >= sp[j + 1] || (int64_t)sc[lo] != i) atomicAdd(bad, 1ull);
  }
}

// The fixed point is event driven.  A work list holds the vertices that may be decidable; a round is two launches:
//   decide  every listed vertex looks at its earlier vertices within distance 2: a ROOT among them -> NOT; none undecided ->
//           ROOT; else it records the LARGEST undecided one as its blocker (the front moves in id order, so that one goes
//           last) and sleeps.  Reads of `state` may be stale inside the launch: that only delays a decision.
//   wake    (all states of the round are final now) a listed vertex that is still undecided re-queues itself if its blocker
//           was decided meanwhile; a vertex decided in this round queues every later vertex within distance 2 that sleeps
//           on it.  `stamp` keeps a vertex from entering a list twice.
// Invariant: an undecided vertex is listed or sleeps on an undecided blocker; the smallest undecided vertex has no
// undecided earlier vertex, so it is listed: an empty list means everything is decided.  One warp per listed vertex: lane l
// walks the neighbourhoods of the neighbours l, l + 32, ... (rows of 7 and rows of 80 entries get the same short chain).
template <class F>
__device__ __forceinline__ void scan_dist2(const int64_t* __restrict__ sp, const int32_t* __restrict__ sc, int64_t i, int lane, F&& f) {
  const int64_t e0 = sp[i + 1];
  for (int64_t q = sp[i] + lane; q < e0; q += 32) {
    const int32_t v = sc[q];
    f(v);
    const int64_t e1 = sp[v + 1];
    for (int64_t q2 = sp[v]; q2 < e1; ++q2) f(sc[q2]);
  }
}

__global__ void k_agg_decide(const int64_t* __restrict__ sp, const int32_t* __restrict__ sc, const int32_t* __restrict__ list,
                             const unsigned* __restrict__ cnt_ptr, int64_t cnt_fixed, unsigned char* state, int32_t* __restrict__ blocker) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (TB / 32);
  const int64_t cnt = list ? (int64_t)*cnt_ptr : cnt_fixed;
  for (int64_t w = (int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); w < cnt; w += nw) {
    const int64_t i = list ? (int64_t)list[w] : w;
    if (ld_state(state + i) != ST_U) continue;
    bool has_root = false;
    int32_t maxu = -1;
    scan_dist2(sp, sc, i, lane, [&](int32_t k) {
      if ((int64_t)k < i) {
        const unsigned char s = ld_state(state + k);
        if (s == ST_ROOT)
          has_root = true;
        else if (s == ST_U)
          maxu = max(maxu, k);
      }
    });
    has_root = __any_sync(0xffffffffu, has_root);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxu = max(maxu, __shfl_xor_sync(0xffffffffu, maxu, o));
    if (lane == 0) {
      if (has_root)
        state[i] = ST_NOT;
      else if (maxu < 0)
        state[i] = ST_ROOT;
      else
        blocker[i] = maxu;
    }
  }
}

__device__ __forceinline__ void agg_push(int32_t j, int32_t* stamp, int32_t tag, int32_t* list_out, unsigned* cnt_out) {
  if (atomicExch(stamp + j, tag) != tag) list_out[atomicAdd(cnt_out, 1u)] = j;  // at most n entries: one per vertex and list
}

__global__ void k_agg_wake(const int64_t* __restrict__ sp, const int32_t* __restrict__ sc, const int32_t* __restrict__ list,
                           const unsigned* __restrict__ cnt_ptr, int64_t cnt_fixed, const unsigned char* state,
                           const int32_t* __restrict__ blocker, int32_t* stamp, int32_t tag, int32_t* __restrict__ list_out,
                           unsigned* cnt_out) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (TB / 32);
  const int64_t cnt = list ? (int64_t)*cnt_ptr : cnt_fixed;
  for (int64_t w = (int64_t)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); w < cnt; w += nw) {
    const int64_t i = list ? (int64_t)list[w] : w;
    if (ld_state(state + i) == ST_U) {  // asleep: unless its blocker went in this very round
      if (lane == 0 && ld_state(state + blocker[i]) != ST_U) agg_push((int32_t)i, stamp, tag, list_out, cnt_out);
      continue;
    }
    scan_dist2(sp, sc, i, lane, [&](int32_t k) {
      if ((int64_t)k > i && ld_state(state + k) == ST_U && blocker[k] == (int32_t)i) agg_push(k, stamp, tag, list_out, cnt_out);
    });
  }
}

__global__ void k_count_undecided(const unsigned char* __restrict__ state, int64_t n, unsigned long long* out) {
  const int64_t i = (int64_t)blockIdx.x * TB + threadIdx.x;
  const bool u = i < n && state[i] == ST_U;
  const unsigned m = __ballot_sync(0xffffffffu, u);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// flags[pos[g]] = 1 for the roots (part-major order: its exclusive scan ranks the roots inside every part); flags[n] = 0
__global__ void k_root_flags(const unsigned char* __restrict__ state, const int32_t* __restrict__ pos, int64_t n, int64_t* __restrict__ flags) {
  const int64_t g = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (g < n) flags[pos[g]] = state[g] == ST_ROOT ? 1 : 0;
  if (g == n) flags[n] = 0;
}
__global__ void k_gather_i64(const int64_t* __restrict__ src, const int64_t* __restrict__ idx, int n, int64_t* __restrict__ out) {
  const int i = blockIdx.x * TB + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}
// pass 1 ids: the root's rank inside its part, for the root and its neighbours; -1 elsewhere
__global__ void k_agg_pass1(const int64_t* __restrict__ sp, const int32_t* __restrict__ sc, int64_t n, const unsigned char* __restrict__ state,
                            const int32_t* __restrict__ pos, const int32_t* __restrict__ owner, const int64_t* __restrict__ rank,
                            const int64_t* __restrict__ base, int32_t* __restrict__ agg) {
  const int64_t v = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (v >= n) return;
  int64_t r = -1;
  if (state[v] == ST_ROOT) {
    r = v;
  } else {
    for (int64_t q = sp[v]; q < sp[v + 1]; ++q)
      if (state[sc[q]] == ST_ROOT) {
        r = sc[q];
        break;
      }
  }
  agg[v] = r < 0 ? -1 : (int32_t)(rank[pos[r]] - base[owner[r]]);
}
// pass 2 (reads pass-1 ids only); *left counts what would remain for pass 3 (none when S is symmetric)
__global__ void k_agg_pass2(const int64_t* __restrict__ sp, const int32_t* __restrict__ sc, int64_t n, const int32_t* __restrict__ agg1,
                            int32_t* __restrict__ agg2, unsigned long long* left) {
  const int64_t v = (int64_t)blockIdx.x * TB + threadIdx.x;
  if (v >= n) return;
  int32_t a = agg1[v];
  if (a == -1)
    for (int64_t q = sp[v]; q < sp[v + 1]; ++q)
      if (agg1[sc[q]] != -1) {
        a = agg1[sc[q]];
        break;
      }
  agg2[v] = a;
  if (a == -1) atomicAdd(left, 1ull);
}
// runs the fixed point to the end; false: gave up (a dependency chain longer than max_rounds) or something is left undecided
bool mis_rounds(const int64_t* sp, const int32_t* sc, int64_t n, unsigned char* state, int32_t* blocker, int64_t max_rounds, int64_t* rounds_out) {
  constexpr int CHECK_EVERY = 32;
  Buf<int32_t> list[2], stamp;
  Buf<unsigned> cnt;
  list[0].alloc(n);
  list[1].alloc(n);
  stamp.alloc(n);
  cnt.alloc(2);
  GK(cudaMemset(stamp.p, 0xff, n * sizeof(int32_t)));
  GK(cudaMemset(cnt.p, 0, 2 * sizeof(unsigned)));
  const int wpb = TB / 32;
  const int grid0 = (int)std::max<int64_t>(1, std::min<int64_t>((n + wpb - 1) / wpb, (int64_t)1 << 20));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + wpb - 1) / wpb, 148 * 16));
  int64_t round = 0;
  while (true) {
    // list of round r: all vertices (r == 0), else list[r & 1] with cnt[r & 1] entries; round r fills list[(r + 1) & 1]
    const int32_t* in = round == 0 ? nullptr : list[round & 1].p;
    const unsigned* cin = cnt.p + (round & 1);
    unsigned* cout = cnt.p + ((round + 1) & 1);
    k_agg_decide<<<round == 0 ? grid0 : grid, TB>>>(sp, sc, in, cin, n, state, blocker);
    GK(cudaMemsetAsync(cout, 0, sizeof(unsigned)));
    k_agg_wake<<<round == 0 ? grid0 : grid, TB>>>(sp, sc, in, cin, n, state, blocker, stamp.p, (int32_t)(round & 0x3fffffff), list[(round + 1) & 1].p, cout);
    ++round;
    if (round % CHECK_EVERY == 0) {
      GK(cudaGetLastError());
      unsigned next = 0;
      GK(cudaMemcpy(&next, cout, sizeof(next), cudaMemcpyDeviceToHost));
      if (next == 0) break;
      if (round > max_rounds) return false;
    }
  }
  Buf<unsigned long long> left;
  left.alloc(1);
  GK(cudaMemset(left.p, 0, sizeof(unsigned long long)));
  k_count_undecided<<<(int)((n + TB - 1) / TB), TB>>>(state, n, left.p);
  unsigned long long l = 0;
  GK(cudaMemcpy(&l, left.p, sizeof(l), cudaMemcpyDeviceToHost));
  if (rounds_out) *rounds_out = round;
  return l == 0;
}

}  // namespace

bool gpu_aggregate(const GpuMat* A, const int32_t* owner, const int32_t* pos, const int64_t* part_off, int32_t nparts, double eps,
                   bool force, int32_t* agg_by_gid, int64_t* counts) {
  const int64_t n = A->nrows;
  if (n == 0) {
    for (int32_t p = 0; p < nparts; ++p) counts[p] = 0;
    return true;
  }
  if (n >= ((int64_t)1 << 31) - 1) return false;
  // Where it pays (B200, profiles/r03_setup_timing.txt): the fixed point needs ~1500-4000 rounds whatever the size, so a
  // small level or one with long rows (Galerkin matrices: 30-70 strong neighbours, 900-4900 vertices within distance 2) is
  // done sooner by the host's single walk.  force: tests (PAMG_GPU_AGG=2).
  if (!force && (n < 1000000 || A->nnz > 12 * n)) return false;
  if (!force && n > 64000000) return false;  // measured and parity-tested up to 256^3 (16.8 M vertices); 512^3 keeps the host walk
  const int g_n = (int)((n + TB - 1) / TB), g_n1 = (int)((n + 1 + TB - 1) / TB);
  Buf<int32_t> d_owner, d_pos;
  d_owner.upload(owner, (size_t)n);
  d_pos.upload(pos, (size_t)n);
  Buf<double> d_ad;
  if (eps > 0.0) {
    d_ad.alloc(n);
    k_absdiag<<<g_n, TB>>>(A->ptr.p, A->col.p, A->val.p, n, d_ad.p);
  }
  // strength graph S
  Buf<int64_t> d_cnt, s_ptr;
  Buf<int32_t> s_col;
  Buf<char> tmp;
  d_cnt.alloc(n + 1);
  GK(cudaMemset(d_cnt.p, 0, (n + 1) * sizeof(int64_t)));
  k_strong<false><<<g_n, TB>>>(A->ptr.p, A->col.p, A->val.p, d_owner.p, eps, d_ad.p, n, d_cnt.p, nullptr, nullptr);
  s_ptr.alloc(n + 1);
  exclusive_scan_i64(d_cnt.p, s_ptr.p, n + 1, tmp);
  int64_t s_nnz = 0;
  GK(cudaMemcpy(&s_nnz, s_ptr.p + n, sizeof(int64_t), cudaMemcpyDeviceToHost));
  s_col.alloc(std::max<int64_t>(s_nnz, 1));
  k_strong<true><<<g_n, TB>>>(A->ptr.p, A->col.p, A->val.p, d_owner.p, eps, d_ad.p, n, nullptr, s_ptr.p, s_col.p);
  Buf<unsigned long long> d_flag;
  d_flag.alloc(2);
  GK(cudaMemset(d_flag.p, 0, 2 * sizeof(unsigned long long)));
  k_sym_check<<<g_n, TB>>>(s_ptr.p, s_col.p, n, d_flag.p);
  unsigned long long bad = 0;
  GK(cudaMemcpy(&bad, d_flag.p, sizeof(bad), cudaMemcpyDeviceToHost));
  if (bad) return false;  // not symmetric (values near the threshold, or a non-symmetric operator): the host walks the rows
  // pass 1
  Buf<unsigned char> state;
  Buf<int32_t> blocker, agg1, agg2;
  state.alloc(n);
  blocker.alloc(n);
  agg1.alloc(n);
  agg2.alloc(n);
  GK(cudaMemset(state.p, ST_U, n));
  GK(cudaMemset(blocker.p, 0xff, n * sizeof(int32_t)));
  const char* mr = getenv("PAMG_GPU_AGG_MAX_ROUNDS");
  const int64_t max_rounds = mr ? atoll(mr) : 40000;
  int64_t rounds1 = 0;
  if (!mis_rounds(s_ptr.p, s_col.p, n, state.p, blocker.p, max_rounds, &rounds1)) return false;
  Buf<int64_t> flags, rank, d_off, base1;
  flags.alloc(n + 1);
  rank.alloc(n + 1);
  d_off.upload(part_off, (size_t)nparts + 1);
  base1.alloc(nparts + 1);
  k_root_flags<<<g_n1, TB>>>(state.p, d_pos.p, n, flags.p);
  exclusive_scan_i64(flags.p, rank.p, n + 1, tmp);
  k_gather_i64<<<(nparts + 1 + TB - 1) / TB, TB>>>(rank.p, d_off.p, nparts + 1, base1.p);
  k_agg_pass1<<<g_n, TB>>>(s_ptr.p, s_col.p, n, state.p, d_pos.p, d_owner.p, rank.p, base1.p, agg1.p);
  std::vector<int64_t> b1(nparts + 1);
  GK(cudaMemcpy(b1.data(), base1.p, (nparts + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  // pass 2
  GK(cudaMemset(d_flag.p, 0, 2 * sizeof(unsigned long long)));
  k_agg_pass2<<<g_n, TB>>>(s_ptr.p, s_col.p, n, agg1.p, agg2.p, d_flag.p + 1);
  unsigned long long left = 0;
  GK(cudaMemcpy(&left, d_flag.p + 1, sizeof(left), cudaMemcpyDeviceToHost));
  if (left) return false;  // cannot happen with a symmetric S (header comment); the host's pass 3 handles it
  GK(cudaGetLastError());
  GK(cudaDeviceSynchronize());
  agg2.download(agg_by_gid, (size_t)n);
  for (int32_t p = 0; p < nparts; ++p) counts[p] = b1[p + 1] - b1[p];
  if (getenv("PAMG_SETUP_TIMING"))
    std::fprintf(stderr, "[pamg setup] GPU aggregation: %lld vertices, %lld strong edges, %lld rounds\n", (long long)n, (long long)s_nnz,
                 (long long)rounds1);
  return true;
}

void gpu_spgemm(const Csr& A, const Csr& B, Csr& C) {
  std::unique_ptr<GpuMat> a(gpu_upload(A)), b(gpu_upload(B));
  std::unique_ptr<GpuMat> c(gpu_product(a.get(), b.get()));
  gpu_download(c.get(), C);
}

}  // namespace pamg
