// Direct b-tree page writer for the node-response table of cnl-dyna's sqlite database (see dbpages.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cnld {
namespace dbpages {

// rows [first_rowid, first_rowid + nf*P*N) of patch_to_node_freq_resp, written in (frequency, source, node) order
struct NodeBlock {
  uint64_t first_rowid = 0;
  uint32_t nf = 0, P = 0, N = 0;
  std::vector<double> freqs;
};

// Return codes: 0 done, 1 not applicable to this file (nothing was modified; the caller takes the SQL path),
// -1 error (cnld::set_error holds the message).  The caller holds sqlite's exclusive lock on the file and has
// no page of it cached that it will use afterwards.
int append_node_rows(const char *file, uint32_t root_page, uint32_t nf, const double *freqs, const double *wavenums,
                     uint32_t P, uint32_t N, const double *x_node /* [nf][P][N] (re, im) */,
                     const double *nodes /* [N][3] */, int threads, uint64_t *first_rowid);
// index on (source_patch, node_id, frequency) of the rows the blocks describe; *root_page = root of the new b-tree
int build_node_index(const char *file, const std::vector<NodeBlock> &blocks, int threads, uint32_t *root_page);

// rows of patch_to_patch_freq_resp from x_patch[nf][P][P] (re, im), times[nf] (nullable: 0); its index
// (source_patch, dest_patch, frequency) = build_node_index with blocks whose N is P
int append_patch_rows(const char *file, uint32_t root_page, uint32_t nf, const double *freqs, const double *wavenums,
                      uint32_t P, const double *x_patch, const double *times, int threads, uint64_t *first_rowid);

// rows (source_patch, dest_patch, time, displacement) of patch_to_patch_imp_resp from h[P][P][nt], and the
// index on (source_patch, dest_patch, time) of exactly those rows (times strictly ascending)
int append_imp_resp_rows(const char *file, uint32_t root_page, uint32_t P, uint32_t nt, const double *times, const double *h,
                         int threads, uint64_t *first_rowid);
int build_imp_resp_index(const char *file, uint64_t first_rowid, uint32_t P, uint32_t nt, const double *times, int threads,
                         uint32_t *root_page);

}  // namespace dbpages
}  // namespace cnld
