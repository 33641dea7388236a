// hierarchy_io.cpp — binary save / load of a host-built Hierarchy (product code).
//
// One process per GPU: the smoothed-aggregation setup is deterministic, but repeating it in every
// rank multiplies host time and memory by the number of GPUs (256^3: 14 s and 13 GB per rank).  Rank 0
// builds the hierarchy once with every host core, saves it to a shared-memory file, and the other
// ranks load it — each one only its own part in full.  For the parts it does not drive a rank needs
// metadata only (sizes, index-partition neighbours, halo plans, block nnz), except on the small
// levels of the replicated coarse tail, whose matrices every GPU holds whole.
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "host.hpp"

namespace pamg {
namespace {

constexpr uint64_t MAGIC = 0x50414d4748494552ull;  // "PAMGHIER"
constexpr uint32_t VERSION = 2;

struct File {
  FILE* f = nullptr;
  ~File() {
    if (f) std::fclose(f);
  }
};

void wr(FILE* f, const void* p, size_t n) {
  if (n && std::fwrite(p, 1, n, f) != n) throw std::runtime_error("hierarchy save: short write");
}
void rd(FILE* f, void* p, size_t n) {
  if (n && std::fread(p, 1, n, f) != n) throw std::runtime_error("hierarchy load: truncated file");
}
template <class T>
void wr_pod(FILE* f, const T& v) {
  wr(f, &v, sizeof(T));
}
template <class T>
T rd_pod(FILE* f) {
  T v;
  rd(f, &v, sizeof(T));
  return v;
}
template <class T>
void wr_vec(FILE* f, const std::vector<T>& v) {
  wr_pod<uint64_t>(f, (uint64_t)v.size());
  wr(f, v.data(), v.size() * sizeof(T));
}
// keep == false: skip the payload (the vector stays empty)
template <class T>
void rd_vec(FILE* f, std::vector<T>& v, bool keep) {
  const uint64_t n = rd_pod<uint64_t>(f);
  if (keep) {
    v.resize(n);
    rd(f, v.data(), n * sizeof(T));
  } else {
    v.clear();
    if (n && fseeko(f, (off_t)(n * sizeof(T)), SEEK_CUR) != 0) throw std::runtime_error("hierarchy load: seek failed");
  }
}

void wr_csr(FILE* f, const LocalCsr& m) {
  wr_pod<int64_t>(f, m.nrows);
  wr_pod<int64_t>(f, m.ncols);
  wr_vec(f, m.ptr);
  wr_vec(f, m.col);
  wr_vec(f, m.val);
}
void rd_csr(FILE* f, LocalCsr& m, bool keep) {
  m.nrows = rd_pod<int64_t>(f);
  m.ncols = rd_pod<int64_t>(f);
  rd_vec(f, m.ptr, keep);
  rd_vec(f, m.col, keep);
  rd_vec(f, m.val, keep);
}

}  // namespace

void save_hierarchy(const Hierarchy& h, const std::string& path) {
  if (!h.ready) throw std::runtime_error("hierarchy save: no hierarchy");
  File F;
  F.f = std::fopen(path.c_str(), "wb");
  if (!F.f) throw std::runtime_error("hierarchy save: cannot open " + path);
  FILE* f = F.f;
  wr_pod(f, MAGIC);
  wr_pod(f, VERSION);
  wr_pod<int32_t>(f, h.nparts);
  wr_pod(f, h.opts);
  wr_pod<int32_t>(f, (int32_t)h.levels.size());
  wr_pod<int64_t>(f, h.n_coarse);
  wr_vec(f, h.coarse_inv);
  wr_vec(f, h.coarse_part_offset);
  for (const Level& lev : h.levels) {
    wr_pod<int64_t>(f, lev.n_global);
    wr_pod<double>(f, lev.rho);
    wr_pod<double>(f, lev.omega_p);
    for (const PartLevel& pl : lev.parts) {
      wr_pod<int64_t>(f, pl.n_own);
      wr_pod<int64_t>(f, pl.n_ghost);
      wr_pod<int64_t>(f, pl.n_own_coarse);
      wr_pod<int64_t>(f, pl.n_ghost_coarse);
      for (int b = 0; b < 6; ++b) wr_pod<int64_t>(f, pl.block_nnz(b));
      wr_pod<uint64_t>(f, (uint64_t)pl.send_idx.size());
      wr_vec(f, pl.recv);
      wr_vec(f, pl.send);
      // heavy part
      wr_vec(f, pl.own_to_global);
      wr_vec(f, pl.ghost_to_global);
      wr_vec(f, pl.ghost_to_owner);
      for (int b = 0; b < 6; ++b) wr_csr(f, pl.blk[b]);
      wr_vec(f, pl.diag);
      wr_vec(f, pl.diag_l1);
      wr_vec(f, pl.agg_local);
      wr_vec(f, pl.send_idx);
    }
  }
  if (std::fflush(f) != 0) throw std::runtime_error("hierarchy save: flush failed");
}

// keep_part >= 0: arrays of the other parts are loaded only on levels of at most `full_rows` global rows (the
// replicated coarse tail needs them); their metadata is always loaded.  keep_part < 0: everything.
void load_hierarchy(Hierarchy& h, const std::string& path, int32_t keep_part) {
  File F;
  F.f = std::fopen(path.c_str(), "rb");
  if (!F.f) throw std::runtime_error("hierarchy load: cannot open " + path);
  FILE* f = F.f;
  if (rd_pod<uint64_t>(f) != MAGIC || rd_pod<uint32_t>(f) != VERSION) throw std::runtime_error("hierarchy load: not a pamg hierarchy file");
  h = Hierarchy();
  h.nparts = rd_pod<int32_t>(f);
  h.opts = rd_pod<pamg_options>(f);
  if (h.opts.struct_size != (int32_t)sizeof(pamg_options)) throw std::runtime_error("hierarchy load: options layout mismatch");
  const int32_t L = rd_pod<int32_t>(f);
  if (h.nparts < 1 || L < 1 || L > 16 || keep_part >= h.nparts) throw std::runtime_error("hierarchy load: bad header");
  h.n_coarse = rd_pod<int64_t>(f);
  rd_vec(f, h.coarse_inv, true);
  rd_vec(f, h.coarse_part_offset, true);
  h.levels.resize(L);
  const int64_t full_rows = std::max<int64_t>(h.opts.tail_rows, 0);
  for (int32_t l = 0; l < L; ++l) {
    Level& lev = h.levels[l];
    lev.n_global = rd_pod<int64_t>(f);
    lev.rho = rd_pod<double>(f);
    lev.omega_p = rd_pod<double>(f);
    lev.parts.resize(h.nparts);
    const bool small = lev.n_global <= full_rows || l == L - 1;
    for (int32_t p = 0; p < h.nparts; ++p) {
      PartLevel& pl = lev.parts[p];
      pl.present = true;
      pl.n_own = rd_pod<int64_t>(f);
      pl.n_ghost = rd_pod<int64_t>(f);
      pl.n_own_coarse = rd_pod<int64_t>(f);
      pl.n_ghost_coarse = rd_pod<int64_t>(f);
      pl.nnz_meta.resize(6);
      for (int b = 0; b < 6; ++b) pl.nnz_meta[b] = rd_pod<int64_t>(f);
      pl.n_send_meta = (int64_t)rd_pod<uint64_t>(f);
      rd_vec(f, pl.recv, true);
      rd_vec(f, pl.send, true);
      const bool keep = keep_part < 0 || p == keep_part || small;
      pl.full = keep;
      rd_vec(f, pl.own_to_global, keep);
      rd_vec(f, pl.ghost_to_global, keep);
      rd_vec(f, pl.ghost_to_owner, keep);
      for (int b = 0; b < 6; ++b) rd_csr(f, pl.blk[b], keep);
      rd_vec(f, pl.diag, keep);
      rd_vec(f, pl.diag_l1, keep);
      rd_vec(f, pl.agg_local, keep);
      rd_vec(f, pl.send_idx, keep);
    }
  }
  h.ready = true;
}

}  // namespace pamg
