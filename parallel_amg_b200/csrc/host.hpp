// host.hpp — host-side data model of the partitioned AMG hierarchy (product code).
//
// Mirrors the PartitionedArrays.jl objects the solve phase consumes (SURVEY.md App. A,
// [RECALL-UNVERIFIED]; the reference snapshot /root/reference/README.md:1-2 has no code):
//   PRange / index partition  -> PartLevel::{own_to_global, ghost_to_global, ghost_to_owner}
//   PSparseMatrix split blocks -> LocalCsr A_oo/A_og (+ P, R of the AMG level)
//   exchange plan of consistent!/assemble! -> PartLevel::{recv, send, send_idx}
#pragma once
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>
#ifdef __linux__
#include <sys/mman.h>
#endif

#include "../../include/pamg.h"

namespace pamg {

// The setup allocates a few GB of fresh vectors (matrices of every level, global and split); their first touch costs a page fault
// per 4 KB.  Where the kernel offers transparent huge pages on request (THP = madvise) a large buffer can be reserved first and
// advised, so that the value-initialising resize / assign behind it faults in 2 MB steps.  MEASURED AND LEFT OFF: in the build
// container 480 MB take 0.13-0.19 s instead of 0.31-0.42 s and the 160^3 host setup 4.1-4.9 s instead of 4.5-5.1 s, but on the GPU
// box the 256^3 setup with the device chain went 8.96 -> 11.1 / 11.6 s (profiles/r03_huge_pages.log: the gallery halves, the phases
// that move matrices through pinned staging lose more).  PAMG_HUGE_PAGES=1 turns the hint on.
inline bool huge_pages_wanted() {
  static const bool on = [] {
    const char* e = std::getenv("PAMG_HUGE_PAGES");
    return e && std::atoi(e) != 0;
  }();
  return on;
}
template <class T>
inline void huge_reserve(std::vector<T>& v, size_t n) {
  if (n * sizeof(T) < ((size_t)8 << 20) || v.capacity() >= n || !huge_pages_wanted()) return;
  v.reserve(n);
#if defined(__linux__) && defined(MADV_HUGEPAGE)
  const uintptr_t two_mb = (uintptr_t)2 << 20;
  const uintptr_t a = ((uintptr_t)v.data() + two_mb - 1) & ~(two_mb - 1);
  const uintptr_t e = ((uintptr_t)v.data() + n * sizeof(T)) & ~(two_mb - 1);
  if (e > a) (void)madvise((void*)a, e - a, MADV_HUGEPAGE);
#endif
}

struct Csr {  // global matrix, global ids, sorted columns
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> ptr, col;
  std::vector<double> val;
  int64_t nnz() const { return (int64_t)col.size(); }
};

struct LocalCsr {  // one block of the split format; local int32 column ids
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> ptr;
  std::vector<int32_t> col;
  std::vector<double> val;
  int64_t nnz() const { return (int64_t)col.size(); }
};

struct Neighbor {
  int32_t part;    // the other part
  int32_t slot0;   // recv: first ghost slot here; send: first ghost slot in the peer
  int32_t count;
  int64_t offset;  // send: offset into send_idx
};

struct PartLevel {
  bool present = false;
  bool full = true;                 // false: loaded as metadata only (another rank drives this part)
  std::vector<int64_t> nnz_meta;    // [6] block nnz when the blocks themselves were not loaded
  int64_t n_send_meta = -1;         // send_idx.size() when send_idx was not loaded
  int64_t block_nnz(int b) const { return full || nnz_meta.empty() ? blk[b].nnz() : nnz_meta[b]; }
  int64_t n_send_entries() const { return full || n_send_meta < 0 ? (int64_t)send_idx.size() : n_send_meta; }
  int64_t n_own = 0, n_ghost = 0, n_own_coarse = 0, n_ghost_coarse = 0;
  std::vector<int64_t> own_to_global, ghost_to_global;
  std::vector<int32_t> ghost_to_owner;
  LocalCsr blk[6];  // PAMG_A_OO .. PAMG_R_OG
  std::vector<double> diag, diag_l1;
  std::vector<int32_t> agg_local;
  std::vector<Neighbor> recv, send;
  std::vector<int32_t> send_idx;
};

struct Level {
  int64_t n_global = 0;
  double rho = 0.0, omega_p = 0.0;
  std::vector<PartLevel> parts;
};

struct Hierarchy {
  int32_t nparts = 0;
  pamg_options opts{};
  std::vector<Level> levels;
  int64_t n_coarse = 0;
  std::vector<double> coarse_inv;            // row-major n_coarse x n_coarse
  std::vector<int64_t> coarse_part_offset;   // nparts+1, own range of each part on the coarsest level
  bool ready = false;
};

// ---- gallery / partition (host_setup.cpp) ---------------------------------------------------
void local_range(int64_t p, int64_t nparts, int64_t n, int64_t* off, int64_t* len);
void uniform_partition(int ndim, const int64_t* dims, const int32_t* pdims, std::vector<int32_t>& owner);
void gallery_poisson(int ndim, const int64_t* dims, Csr& A);
void gallery_diffusion_jump(int ndim, const int64_t* dims, int blocks, double kmax, double eps_z, Csr& A);
void matvec(const Csr& A, const double* x, double* y);

// ---- sparse products of the setup on the GPU (setup_gpu.cu) ----------------------------------
// gpu_setup_available: a CUDA device is present and PAMG_GPU_SETUP != 0.  gpu_spgemm: C = A*B with the host product's
// accumulation order (bit-identical structure and values); throws on CUDA errors (the caller falls back to the host).
bool gpu_setup_available();
void gpu_spgemm(const Csr& A, const Csr& B, Csr& C);
// the same with device-resident operands, so that a chain of products (A_F*P0, A*P, R*(A*P), the next level's A) pays the
// PCIe transfer of every matrix at most once
struct GpuMat;
GpuMat* gpu_upload(const Csr& A);
GpuMat* gpu_product(const GpuMat* A, const GpuMat* B);
void gpu_download(const GpuMat* m, Csr& C);
void gpu_free(GpuMat* m);
GpuMat* gpu_smooth_prolongator(const GpuMat* F, const GpuMat* P0, const double* w_host);  // P0 + diag(w) (F P0), merged by column
GpuMat* gpu_transpose(const GpuMat* A);                                                   // ascending columns in every row
// the greedy 3-pass aggregation of ALL parts on the device, bit-identical to the host's aggregate_part.
// A: the level matrix (or the node graph) on the device; owner[g]; pos[g] = first row of the owner part + own-local id
// (part-major position); part_off[nparts + 1]; agg_by_gid[g] = aggregate id inside the owner part; counts[p] = aggregates
// of part p.  Returns false (nothing written) when the strength graph is not symmetric, the dependency chain is too long, or
// (force == false) the level is one the host's walk finishes sooner (small, or long rows): the caller then walks the rows.
bool gpu_aggregate(const GpuMat* A, const int32_t* owner, const int32_t* pos, const int64_t* part_off, int32_t nparts, double eps,
                   bool force, int32_t* agg_by_gid, int64_t* counts);
void gpu_setup_begin();  // pinned staging buffers for the transfers of one setup
void gpu_setup_end();

// ---- setup ------------------------------------------------------------------------------------
// Builds every level (global matrices internally, then the per-part split format).
// Throws std::runtime_error on bad input; the C ABI catches.
// block_size / ns_k / nullspace: DOFs per node on level 0 and the near-nullspace (row-major n x ns_k, e.g. the
// 6 rigid-body modes); nullspace == nullptr: scalar smoothed aggregation (piecewise-constant tentative P)
void build_hierarchy(const Csr& A, const std::vector<int32_t>& owner, int32_t nparts,
                     const pamg_options& o, Hierarchy& h, int32_t block_size = 1, int32_t ns_k = 0,
                     const std::vector<double>* nullspace = nullptr);
void gallery_elasticity(const int64_t* dims, double E, double nu, Csr& A, std::vector<double>& coords);
void rigid_body_modes(const std::vector<double>& coords, std::vector<double>& B);
// halo plans from the index maps of all parts of a level (also used for external hierarchies)
void build_halo_plans(Level& lev, int32_t nparts);
void finalize_external(Hierarchy& h);
// binary hand-off of a built hierarchy between the ranks of one node (hierarchy_io.cpp)
void save_hierarchy(const Hierarchy& h, const std::string& path);
void load_hierarchy(Hierarchy& h, const std::string& path, int32_t keep_part);

}  // namespace pamg
