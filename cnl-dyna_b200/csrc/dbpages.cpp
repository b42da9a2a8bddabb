// Direct b-tree page writer for the node-response table of cnl-dyna's sqlite database.
//
// cnld/scripts/generate_db.py:114-127 stores, per frequency and source patch, one row per mesh node in
// patch_to_node_freq_resp (cnld/database.py:38-45): 2.13 M rows = 150 MB per frequency at the 4x4 array.
// Through SQL -- even prepared statements staged in parallel and appended by raw record copy
// (dbwriter.cpp) -- that is 0.3 s of single-writer time per frequency, while the GPU delivers a
// frequency every 0.13 s.  The rows are append-only with consecutive rowids, so their b-tree can be written
// as what it is: this file encodes the records into table-leaf pages in parallel, appends the pages to the
// database file, and rebuilds the (small) interior levels above old and new leaves; the index on
// (source_patch, node_id, frequency) is generated the same way, already sorted, from the block
// descriptors -- no scan and no sort.  The result is an ordinary database in SQLite's documented file
// format (sqlite.org/fileformat2.html: record format with the REAL-as-integer storage rule, table / index
// b-tree pages, freelist trunk pages, the lock-byte page at 1 GiB); tests/test_dbpages_cpu.py checks
// PRAGMA integrity_check and row-for-row equality with the SQL path.
//
// Not applicable (the caller falls back to SQL): WAL mode, auto-vacuum, a hot journal, reserved bytes that
// leave less than 480 usable bytes, page 1 as the root.
#include "dbpages.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>

namespace cnld {
void set_error(const char *fmt, ...);
namespace dbpages {
namespace {

inline uint32_t be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
inline void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

inline int put_varint(uint8_t *p, uint64_t v) {
  if (v <= 0x7f) { p[0] = (uint8_t)v; return 1; }
  if (v <= 0x3fff) { p[0] = (uint8_t)((v >> 7) | 0x80); p[1] = (uint8_t)(v & 0x7f); return 2; }
  if (v & 0xff00000000000000ull) {       // nine bytes: the last one carries eight bits
    p[8] = (uint8_t)v;
    v >>= 8;
    for (int i = 7; i >= 0; i--) { p[i] = (uint8_t)((v & 0x7f) | 0x80); v >>= 7; }
    return 9;
  }
  uint8_t b[10];
  int n = 0;
  do { b[n++] = (uint8_t)((v & 0x7f) | 0x80); v >>= 7; } while (v);
  b[0] &= 0x7f;
  for (int i = 0; i < n; i++) p[i] = b[n - 1 - i];
  return n;
}
// v = [p, p + n)  (vector::assign from raw pointers trips a gcc 13 -Wnonnull false positive)
inline void set_bytes(std::vector<uint8_t> &v, const uint8_t *p, size_t n) {
  v.resize(n);
  if (n) memcpy(v.data(), p, n);
}
inline int get_varint(const uint8_t *p, uint64_t *v) {
  uint64_t r = 0;
  for (int i = 0; i < 8; i++) {
    r = (r << 7) | (p[i] & 0x7f);
    if (!(p[i] & 0x80)) { *v = r; return i + 1; }
  }
  *v = (r << 8) | p[8];
  return 9;
}

// record format: serial type + big-endian body of an integer, smallest representation (types 8 / 9 = the
// constants 0 / 1 need schema format 4)
inline int enc_int(int64_t v, bool fmt4, uint8_t *type, uint8_t *data) {
  if (fmt4 && (v == 0 || v == 1)) { *type = (uint8_t)(8 + v); return 0; }
  const uint64_t u = v < 0 ? ~(uint64_t)v : (uint64_t)v;
  int n;
  if (u <= 127) { *type = 1; n = 1; }
  else if (u <= 32767) { *type = 2; n = 2; }
  else if (u <= 8388607) { *type = 3; n = 3; }
  else if (u <= 2147483647) { *type = 4; n = 4; }
  else if (u <= 140737488355327ull) { *type = 5; n = 6; }
  else { *type = 6; n = 8; }
  for (int i = 0; i < n; i++) data[i] = (uint8_t)((uint64_t)v >> (8 * (n - 1 - i)));
  return n;
}
// a value of a REAL-affinity column: SQLite stores integral values that fit 48 bits as integers, NaN as NULL
inline int enc_real(double r, bool fmt4, uint8_t *type, uint8_t *data) {
  if (std::isnan(r)) { *type = 0; return 0; }
  if (r >= -140737488355328.0 && r <= 140737488355327.0) {
    const int64_t i = (int64_t)r;
    if ((double)i == r) return enc_int(i, fmt4, type, data);
  }
  uint64_t b;
  memcpy(&b, &r, 8);
  *type = 7;
  for (int i = 0; i < 8; i++) data[i] = (uint8_t)(b >> (8 * (7 - i)));
  return 8;
}

struct File {
  int fd = -1;
  uint32_t page = 0, usable = 0, pending = 0;
  uint8_t hdr[100];
  uint32_t dbsize = 0;
  bool fmt4 = false;
  std::string err;
  ~File() { if (fd >= 0) close(fd); }
  // usable pages are numbered without the lock-byte page: v -> page number
  uint32_t v2p(uint64_t v) const { return (uint32_t)(v < pending ? v : v + 1); }
  uint64_t vsize() const { return dbsize < pending ? dbsize : dbsize - 1; }
  uint32_t size_for(uint64_t vcount) const { return (uint32_t)(vcount < pending ? vcount : vcount + 1); }
  bool rd(uint32_t pg, uint8_t *buf) const { return pread(fd, buf, page, (off_t)(pg - 1) * page) == (ssize_t)page; }
  bool wr(uint32_t pg, const uint8_t *buf, size_t npages = 1) const {
    size_t left = npages * (size_t)page;
    off_t off = (off_t)(pg - 1) * page;
    while (left) {
      ssize_t w = pwrite(fd, buf, left, off);
      if (w <= 0) { if (w < 0 && errno == EINTR) continue; return false; }
      buf += w; off += w; left -= (size_t)w;
    }
    return true;
  }
  // `n` consecutive usable pages starting at virtual index v (skips the lock-byte page)
  bool wr_virtual(uint64_t v, const uint8_t *buf, size_t n) const {
    if (v < pending && v + n > pending) {
      const size_t a = (size_t)(pending - v);
      return wr(v2p(v), buf, a) && wr(v2p(pending), buf + a * (size_t)page, n - a);
    }
    return wr(v2p(v), buf, n);
  }
};

// 0 ok, 1 not applicable, -1 error
int open_file(const char *file, File &F) {
  F.fd = open(file, O_RDWR | O_CLOEXEC);
  if (F.fd < 0) { set_error("dbpages: cannot open %s: %s", file, strerror(errno)); return -1; }
  if (pread(F.fd, F.hdr, 100, 0) != 100 || memcmp(F.hdr, "SQLite format 3", 16) != 0) return 1;
  F.page = be16(F.hdr + 16);
  if (F.page == 1) F.page = 65536;
  if (F.page < 512 || (F.page & (F.page - 1))) return 1;
  F.usable = F.page - F.hdr[20];
  if (F.usable < 480) return 1;
  if (F.hdr[18] != 1 || F.hdr[19] != 1) return 1;          // WAL
  if (be32(F.hdr + 52) != 0 || be32(F.hdr + 64) != 0) return 1;   // auto-vacuum / incremental vacuum: pointer-map pages
  F.fmt4 = be32(F.hdr + 44) >= 4;
  struct stat st;
  for (const char *suffix : {"-journal", "-wal"}) {
    const std::string j = std::string(file) + suffix;
    if (stat(j.c_str(), &st) == 0 && st.st_size > 0) return 1;
  }
  if (fstat(F.fd, &st) != 0) return 1;
  const uint32_t in_header = be32(F.hdr + 28);
  F.dbsize = (in_header && be32(F.hdr + 24) == be32(F.hdr + 92)) ? in_header : (uint32_t)(st.st_size / F.page);
  if (F.dbsize < 1) return 1;
  F.pending = (uint32_t)(1073741824u / F.page) + 1;
  return 0;
}

// freed page = a freelist trunk page without leaves, chained in front
bool free_page(File &F, uint32_t pg) {
  std::vector<uint8_t> buf(F.page, 0);
  put32(buf.data(), be32(F.hdr + 32));
  put32(buf.data() + 4, 0);
  if (!F.wr(pg, buf.data())) return false;
  put32(F.hdr + 32, pg);
  put32(F.hdr + 36, be32(F.hdr + 36) + 1);
  return true;
}

bool write_header(File &F, uint64_t vcount) {
  const uint32_t cc = be32(F.hdr + 24) + 1;
  put32(F.hdr + 24, cc);
  put32(F.hdr + 28, F.size_for(vcount));
  put32(F.hdr + 92, cc);
  return pwrite(F.fd, F.hdr, 100, 0) == 100;
}

// cells of one b-tree page, collected in key order; emit() lays them out at the end of the page
struct PageBuilder {
  uint32_t usable, page, hdr;
  std::vector<uint8_t> cells;
  std::vector<uint32_t> off;
  PageBuilder(const File &F, bool interior) : usable(F.usable), page(F.page), hdr(interior ? 12 : 8) {
    cells.reserve(F.page);
    off.reserve(F.page / 16);
  }
  size_t n() const { return off.size(); }
  bool fits(size_t c) const { return hdr + 2 * (off.size() + 1) + cells.size() + c <= usable; }
  void add(const uint8_t *c, size_t len) {
    off.push_back((uint32_t)cells.size());
    cells.insert(cells.end(), c, c + len);
  }
  void pop(std::vector<uint8_t> *cell) {
    if (cell) set_bytes(*cell, cells.data() + off.back(), cells.size() - off.back());
    cells.resize(off.back());
    off.pop_back();
  }
  void emit(uint8_t *out, uint8_t type, uint32_t right = 0) {
    memset(out, 0, page);
    const uint32_t content = usable - (uint32_t)cells.size();
    out[0] = type;
    put16(out + 3, (uint32_t)off.size());
    put16(out + 5, content == 65536 ? 0 : content);
    if (hdr == 12) put32(out + 8, right);
    for (size_t i = 0; i < off.size(); i++) put16(out + hdr + 2 * i, content + off[i]);
    if (!cells.empty()) memcpy(out + content, cells.data(), cells.size());
    cells.clear();
    off.clear();
  }
};

struct Child {
  uint32_t pg;
  uint64_t key;     // table b-tree: an upper bound of the rowids below pg
};

// largest rowid of a table leaf page (0 for an empty page)
uint64_t leaf_max_rowid(const uint8_t *p) {
  const uint32_t n = be16(p + 3);
  if (!n) return 0;
  const uint8_t *c = p + be16(p + 8 + 2 * (n - 1));
  uint64_t len, rowid;
  c += get_varint(c, &len);
  get_varint(c, &rowid);
  return rowid;
}

// Leaves (in key order) and interior pages (other than the root) of the table b-tree under `root`.
// Keys of the parent cells bound their left children; the right-most leaf's bound is its own last rowid.
int walk_table(const File &F, uint32_t root, std::vector<Child> &leaves, std::vector<uint32_t> &interior, uint64_t *maxrowid) {
  std::vector<uint8_t> buf(F.page);
  int depth = 1;
  for (uint32_t pg = root;;) {
    if (pg < 2 || pg > F.dbsize || !F.rd(pg, buf.data())) return 1;
    if (buf[0] == 0x0D) break;
    if (buf[0] != 0x05 || ++depth > 20) return 1;
    const uint32_t n = be16(buf.data() + 3);
    pg = n ? be32(buf.data() + be16(buf.data() + 12)) : be32(buf.data() + 8);
  }
  // right spine: the largest rowid
  for (uint32_t pg = root;;) {
    if (pg < 2 || pg > F.dbsize || !F.rd(pg, buf.data())) return 1;
    if (buf[0] == 0x0D) { *maxrowid = leaf_max_rowid(buf.data()); break; }
    pg = be32(buf.data() + 8);
  }
  if (depth == 1) return 0;      // the root is a leaf: the caller moves its cells
  struct Frame { uint32_t pg; int level; uint64_t bound; };
  std::vector<Frame> stack{{root, 1, *maxrowid}};
  // depth-first, children pushed in reverse so that leaves come out in key order
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    if (f.pg < 2 || f.pg > F.dbsize || !F.rd(f.pg, buf.data()) || buf[0] != 0x05) return 1;
    if (f.pg != root) interior.push_back(f.pg);
    const uint32_t n = be16(buf.data() + 3);
    std::vector<Child> kids(n + 1);
    for (uint32_t i = 0; i < n; i++) {
      const uint8_t *c = buf.data() + be16(buf.data() + 12 + 2 * i);
      kids[i].pg = be32(c);
      get_varint(c + 4, &kids[i].key);
    }
    kids[n] = {be32(buf.data() + 8), f.bound};
    if (f.level == depth - 1) leaves.insert(leaves.end(), kids.begin(), kids.end());
    else for (size_t i = kids.size(); i-- > 0;) stack.push_back({kids[i].pg, f.level + 1, kids[i].key});
  }
  return 0;
}

struct Alloc {
  std::vector<uint32_t> pool;      // pages of the old interior levels, reused first
  std::atomic<uint64_t> next;      // next free virtual page index
  const File &F;
  Alloc(const File &f) : next(f.vsize() + 1), F(f) {}
  uint32_t one() {
    if (!pool.empty()) { const uint32_t p = pool.back(); pool.pop_back(); return p; }
    return F.v2p(next.fetch_add(1));
  }
};

// interior levels of a table b-tree over `kids`; the top page is written to `root`
bool build_table_interior(File &F, Alloc &A, std::vector<Child> kids, uint32_t root) {
  std::vector<uint8_t> out(F.page);
  uint8_t cell[16];
  while (true) {
    std::vector<Child> parents;
    const size_t n = kids.size();
    size_t i = 0;
    while (i < n) {
      PageBuilder pb(F, true);
      const size_t start = i;
      while (i + 1 < n) {
        put32(cell, kids[i].pg);
        const int len = 4 + put_varint(cell + 4, kids[i].key);
        if (!pb.fits(len)) break;
        pb.add(cell, len);
        i++;
      }
      if (n - (i + 1) == 1) {            // never leave a single child for the next page
        if (pb.n() < 2) { set_error("dbpages: page too small for an interior node"); return false; }
        pb.pop(nullptr);
        i--;
      }
      const bool top = start == 0 && i + 1 == n;
      const uint32_t pg = top ? root : A.one();
      pb.emit(out.data(), 0x05, kids[i].pg);
      if (!F.wr(pg, out.data())) { set_error("dbpages: write failed: %s", strerror(errno)); return false; }
      parents.push_back({pg, kids[i].key});
      i++;
    }
    if (parents.size() == 1) return true;
    kids.swap(parents);
  }
}

int nthreads_for(int threads, size_t tasks) {
  int t = threads > 0 ? threads : 1;
  if (t > 64) t = 64;
  if ((size_t)t > tasks) t = (int)tasks;
  return t < 1 ? 1 : t;
}

struct Enc {
  uint8_t type, len, data[8];
};


// ---- generic bulk append to a rowid table -------------------------------------------------------------
// `ntasks` tasks of `rows_per_task` rows each, rowids consecutive in (task, row) order behind the table's
// largest rowid; make_row(task, j, payload) writes the record (header + bodies, < 120 bytes) of row j of a
// task and returns its length.  The leaves of a task start on a fresh page, so tasks are independent:
// threads encode them into private buffers and take page numbers from one atomic counter.
template <class MakeRow>
int append_rows_impl(const char *file, uint32_t root, size_t ntasks, size_t rows_per_task, int threads, uint64_t *first_rowid,
                     MakeRow make_row) {
  if (!ntasks || !rows_per_task) return 1;
  File F;
  if (int rc = open_file(file, F)) return rc;
  if (root < 2 || root > F.dbsize) return 1;
  std::vector<Child> leaves;
  Alloc A(F);
  uint64_t maxrowid = 0;
  if (walk_table(F, root, leaves, A.pool, &maxrowid)) return 1;
  std::vector<uint8_t> tmp(F.page);
  if (leaves.empty()) {                  // the root is a leaf; its cells (if any) move to a page of their own
    if (!F.rd(root, tmp.data()) || tmp[0] != 0x0D) return 1;
    if (be16(tmp.data() + 3)) {
      const uint32_t pg = A.one();
      if (!F.wr(pg, tmp.data())) { set_error("dbpages: write failed: %s", strerror(errno)); return -1; }
      leaves.push_back({pg, maxrowid});
    }
  }
  const uint64_t row0 = maxrowid + 1;
  if (first_rowid) *first_rowid = row0;

  struct Result { uint64_t vbase = 0; std::vector<uint64_t> last; };
  std::vector<Result> res(ntasks);
  std::atomic<size_t> next_task{0};
  std::atomic<int> failed{0};
  auto work = [&] {
    PageBuilder pb(F, false);
    std::vector<uint8_t> buf;
    buf.reserve(rows_per_task * 96 + 2 * (size_t)F.page);
    uint8_t cell[160];
    for (size_t q; (q = next_task.fetch_add(1)) < ntasks && !failed.load(std::memory_order_relaxed);) {
      Result &R = res[q];
      buf.clear();
      uint64_t rowid = row0 + q * (uint64_t)rows_per_task;
      for (size_t j = 0; j < rows_per_task; j++, rowid++) {
        uint8_t payload[128];
        const size_t plen = make_row(q, j, payload);
        uint8_t *c = cell;
        *c++ = (uint8_t)plen;                // < 128: one varint byte
        c += put_varint(c, rowid);
        memcpy(c, payload, plen); c += plen;
        const size_t clen = (size_t)(c - cell);
        if (!pb.fits(clen)) {
          buf.resize(buf.size() + F.page);
          pb.emit(buf.data() + buf.size() - F.page, 0x0D);
          R.last.push_back(rowid - 1);
        }
        pb.add(cell, clen);
      }
      buf.resize(buf.size() + F.page);
      pb.emit(buf.data() + buf.size() - F.page, 0x0D);
      R.last.push_back(rowid - 1);
      R.vbase = A.next.fetch_add(R.last.size());
      if (!F.wr_virtual(R.vbase, buf.data(), R.last.size())) failed.store(errno ? errno : EIO);
    }
  };
  const int nt = nthreads_for(threads, ntasks);
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  if (failed.load()) { set_error("dbpages: writing %s failed: %s", file, strerror(failed.load())); return -1; }

  for (size_t q = 0; q < ntasks; q++)
    for (size_t k = 0; k < res[q].last.size(); k++) leaves.push_back({F.v2p(res[q].vbase + k), res[q].last[k]});
  if (leaves.size() == 1) {              // a single leaf is the root itself
    if (!F.rd(leaves[0].pg, tmp.data()) || !F.wr(root, tmp.data())) { set_error("dbpages: write failed: %s", strerror(errno)); return -1; }
    A.pool.push_back(leaves[0].pg);
  } else if (!build_table_interior(F, A, leaves, root)) {
    return -1;
  }
  while (!A.pool.empty()) {
    const uint32_t pg = A.pool.back();
    A.pool.pop_back();
    if (!free_page(F, pg)) { set_error("dbpages: write failed: %s", strerror(errno)); return -1; }
  }
  if (!write_header(F, A.next.load() - 1)) { set_error("dbpages: header write failed: %s", strerror(errno)); return -1; }
  return 0;
}

// ---- generic bulk build of an index ---------------------------------------------------------------------
// `groups` x `per_group` entries, already in key order as (group, k); make_entry(group, k, payload) writes
// the index record (key columns + rowid) and returns its length (< 120).  Tasks = contiguous ranges of
// groups; a task emits [separator] leaf (separator leaf)* -- separators are the entries that go to the
// interior pages -- so the levels above are assembled from the tasks' results.
template <class MakeEntry>
int build_index_impl(const char *file, size_t groups, size_t per_group, int threads, uint32_t *root_out, MakeEntry make_entry) {
  if (groups * per_group == 0) return 1;
  File F;
  if (int rc = open_file(file, F)) return rc;
  const size_t want = std::max<size_t>(2, 200000 / per_group);      // ~4 MB of leaves per task
  const size_t ntasks = groups * per_group < 2 ? 1 : std::max<size_t>(1, std::min((groups + want - 1) / want, std::max<size_t>(1, groups / 2)));
  struct Result { std::vector<uint8_t> lead; uint64_t vbase = 0; size_t npages = 0; std::vector<std::vector<uint8_t>> seps; };
  std::vector<Result> res(ntasks);
  Alloc A(F);
  std::atomic<size_t> next_task{0};
  std::atomic<int> failed{0};
  auto work = [&] {
    PageBuilder pb(F, false);
    std::vector<uint8_t> buf;
    uint8_t cell[160];
    for (size_t t; (t = next_task.fetch_add(1)) < ntasks && !failed.load(std::memory_order_relaxed);) {
      Result &R = res[t];
      const size_t a = groups * t / ntasks, b = groups * (t + 1) / ntasks;
      buf.clear();
      bool want_lead = t > 0;
      for (size_t g = a; g < b; g++)
        for (size_t k = 0; k < per_group; k++) {
          uint8_t *p = cell + 1;
          const size_t plen = make_entry(g, k, p);
          cell[0] = (uint8_t)plen;                   // < 128
          const bool last = g + 1 == b && k + 1 == per_group;
          if (want_lead && !last) { set_bytes(R.lead, p, plen); want_lead = false; continue; }
          if (pb.fits(plen + 1)) { pb.add(cell, plen + 1); continue; }
          // the page is full: this entry separates it from the next leaf -- unless it is the task's last
          // entry, which must sit in a leaf: then the page's last cell becomes the separator instead
          std::vector<uint8_t> sep;
          if (last) {
            pb.pop(&sep);
            sep.erase(sep.begin());                  // separators are kept as payloads (no length varint)
          } else {
            set_bytes(sep, p, plen);
          }
          buf.resize(buf.size() + F.page);
          pb.emit(buf.data() + buf.size() - F.page, 0x0A);
          R.seps.push_back(std::move(sep));
          if (last) pb.add(cell, plen + 1);
        }
      if (want_lead) { failed.store(EINVAL); break; }       // a task of t > 0 holds at least two entries
      buf.resize(buf.size() + F.page);
      pb.emit(buf.data() + buf.size() - F.page, 0x0A);
      R.npages = buf.size() / F.page;
      R.vbase = A.next.fetch_add(R.npages);
      if (!F.wr_virtual(R.vbase, buf.data(), R.npages)) failed.store(errno ? errno : EIO);
    }
  };
  const int nt = nthreads_for(threads, ntasks);
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
  if (failed.load()) { set_error("dbpages: writing an index of %s failed: %s", file, strerror(failed.load())); return -1; }

  std::vector<uint32_t> kids;
  std::vector<std::vector<uint8_t>> seps;
  for (size_t t = 0; t < ntasks; t++) {
    Result &R = res[t];
    if (t > 0) seps.push_back(std::move(R.lead));
    for (size_t k = 0; k < R.npages; k++) {
      kids.push_back(F.v2p(R.vbase + k));
      if (k + 1 < R.npages) seps.push_back(std::move(R.seps[k]));
    }
  }
  if (kids.empty() || seps.size() + 1 != kids.size()) { set_error("dbpages: index layout mismatch"); return -1; }
  std::vector<uint8_t> out(F.page), cell;
  while (kids.size() > 1) {
    std::vector<uint32_t> parents;
    std::vector<std::vector<uint8_t>> psep;
    const size_t n = kids.size();
    size_t i = 0;
    while (i < n) {
      PageBuilder pb(F, true);
      while (i + 1 < n) {
        const std::vector<uint8_t> &sp = seps[i];
        cell.resize(4 + 9 + sp.size());
        put32(cell.data(), kids[i]);
        const int vl = put_varint(cell.data() + 4, sp.size());
        memcpy(cell.data() + 4 + vl, sp.data(), sp.size());
        const size_t len = 4 + vl + sp.size();
        if (!pb.fits(len)) break;
        pb.add(cell.data(), len);
        i++;
      }
      if (n - (i + 1) == 1) {            // never leave a single child for the next page
        if (pb.n() < 2) { set_error("dbpages: page too small for an interior index node"); return -1; }
        pb.pop(nullptr);
        i--;
      }
      const uint32_t pg = A.one();
      pb.emit(out.data(), 0x02, kids[i]);
      if (!F.wr(pg, out.data())) { set_error("dbpages: write failed: %s", strerror(errno)); return -1; }
      parents.push_back(pg);
      if (i + 1 < n) psep.push_back(std::move(seps[i]));
      i++;
    }
    kids.swap(parents);
    seps.swap(psep);
  }
  *root_out = kids[0];
  if (!write_header(F, A.next.load() - 1)) { set_error("dbpages: header write failed: %s", strerror(errno)); return -1; }
  return 0;
}

// schema format of the file (serial types 8 / 9 need format 4); 0 ok, 1 not applicable, -1 error
int file_format4(const char *file, bool *fmt4) {
  File F;
  if (int rc = open_file(file, F)) return rc;
  *fmt4 = F.fmt4;
  return 0;
}

}  // namespace

int append_node_rows(const char *file, uint32_t root, uint32_t nf, const double *freqs, const double *wavenums, uint32_t P,
                     uint32_t N, const double *x_node, const double *nodes, int threads, uint64_t *first_rowid) {
  if (!nf || !P || !N) return 1;
  bool fmt4 = false;
  if (int rc = file_format4(file, &fmt4)) return rc;
  // encodings that do not change from row to row
  std::vector<Enc> es(P), ev(N), ef((size_t)nf * 2), exyz((size_t)N * 3);
  for (uint32_t s = 0; s < P; s++) es[s].len = (uint8_t)enc_int(s, fmt4, &es[s].type, es[s].data);
  for (uint32_t v = 0; v < N; v++) {
    ev[v].len = (uint8_t)enc_int(v, fmt4, &ev[v].type, ev[v].data);
    for (int d = 0; d < 3; d++) {
      Enc &e = exyz[(size_t)v * 3 + d];
      e.len = (uint8_t)enc_real(nodes[(size_t)v * 3 + d], fmt4, &e.type, e.data);
    }
  }
  for (uint32_t i = 0; i < nf; i++) {
    ef[2 * i].len = (uint8_t)enc_real(freqs[i], fmt4, &ef[2 * i].type, ef[2 * i].data);
    ef[2 * i + 1].len = (uint8_t)enc_real(wavenums[i], fmt4, &ef[2 * i + 1].type, ef[2 * i + 1].data);
  }
  // per node: the four serial types and the bodies of (node_id, x, y, z), contiguous
  struct NodePart { uint8_t types[4], len, data[28]; };
  std::vector<NodePart> np(N);
  for (uint32_t v = 0; v < N; v++) {
    NodePart &n = np[v];
    n.types[0] = ev[v].type;
    uint8_t *d = n.data;
    memcpy(d, ev[v].data, ev[v].len); d += ev[v].len;
    for (int k = 0; k < 3; k++) {
      const Enc &e = exyz[(size_t)v * 3 + k];
      n.types[1 + k] = e.type;
      memcpy(d, e.data, e.len); d += e.len;
    }
    n.len = (uint8_t)(d - n.data);
  }
  // one task = the N rows of one (frequency, source patch): columns source_patch, node_id, x, y, z,
  // frequency, wavenumber, displacement_real, displacement_imag (database.py:38-45)
  return append_rows_impl(file, root, (size_t)nf * P, N, threads, first_rowid, [&](size_t q, size_t v, uint8_t *out) -> size_t {
    const uint32_t i = (uint32_t)(q / P), s = (uint32_t)(q % P);
    const double *x = x_node + 2 * (q * (size_t)N + v);
    const Enc &e_s = es[s], &e_f = ef[2 * i], &e_w = ef[2 * i + 1];
    const NodePart &n = np[v];
    uint8_t *b = out + 10;
    out[0] = 10; out[1] = e_s.type;
    memcpy(out + 2, n.types, 4);
    out[6] = e_f.type; out[7] = e_w.type;
    memcpy(b, e_s.data, e_s.len); b += e_s.len;
    memcpy(b, n.data, n.len); b += n.len;
    memcpy(b, e_f.data, e_f.len); b += e_f.len;
    memcpy(b, e_w.data, e_w.len); b += e_w.len;
    b += enc_real(x[0], fmt4, &out[8], b);
    b += enc_real(x[1], fmt4, &out[9], b);
    return (size_t)(b - out);
  });
}

int build_node_index(const char *file, const std::vector<NodeBlock> &blocks, int threads, uint32_t *root_out) {
  if (blocks.empty()) return 1;
  const uint32_t P = blocks[0].P, N = blocks[0].N;
  for (const NodeBlock &b : blocks)
    if (b.P != P || b.N != N || b.freqs.size() != b.nf) return 1;
  if (!P || !N) return 1;
  bool fmt4 = false;
  if (int rc = file_format4(file, &fmt4)) return rc;
  // all (frequency, first rowid of that frequency's rows) in index order
  struct Fq { double f; uint64_t base; Enc e; };
  std::vector<Fq> fq;
  for (const NodeBlock &b : blocks)
    for (uint32_t i = 0; i < b.nf; i++) {
      Fq x{b.freqs[i], b.first_rowid + (uint64_t)i * P * N, {}};
      if (std::isnan(x.f)) return 1;           // NULL keys sort first: leave that to SQL
      x.e.len = (uint8_t)enc_real(x.f, fmt4, &x.e.type, x.e.data);
      fq.push_back(x);
    }
  std::sort(fq.begin(), fq.end(), [](const Fq &a, const Fq &b) { return a.f != b.f ? a.f < b.f : a.base < b.base; });
  std::vector<Enc> es(P), ev(N);
  for (uint32_t s = 0; s < P; s++) es[s].len = (uint8_t)enc_int(s, fmt4, &es[s].type, es[s].data);
  for (uint32_t v = 0; v < N; v++) ev[v].len = (uint8_t)enc_int(v, fmt4, &ev[v].type, ev[v].data);
  // group = (source, node) pair, entries of a group = the frequencies in ascending order
  return build_index_impl(file, (size_t)P * N, fq.size(), threads, root_out, [&](size_t pr, size_t k, uint8_t *p) -> size_t {
    const Enc &e_s = es[pr / N], &e_v = ev[pr % N];
    const Fq &f = fq[k];
    p[0] = 5; p[1] = e_s.type; p[2] = e_v.type; p[3] = f.e.type;
    uint8_t *d = p + 5;
    memcpy(d, e_s.data, e_s.len); d += e_s.len;
    memcpy(d, e_v.data, e_v.len); d += e_v.len;
    memcpy(d, f.e.data, f.e.len); d += f.e.len;
    d += enc_int((int64_t)(f.base + pr), fmt4, &p[4], d);
    return (size_t)(d - p);
  });
}

// patch_to_patch_freq_resp (database.py:22-29): rows (source_patch, dest_patch, frequency, wavenumber,
// displacement_real, displacement_imag, time_solve) in (frequency, source, dest) order; one task per
// frequency.  Its index (source_patch, dest_patch, frequency) has the shape of the node index with the
// destination patch in the place of the node: build_node_index with N = P.
int append_patch_rows(const char *file, uint32_t root, uint32_t nf, const double *freqs, const double *wavenums, uint32_t P,
                      const double *x_patch, const double *times, int threads, uint64_t *first_rowid) {
  if (!nf || !P) return 1;
  bool fmt4 = false;
  if (int rc = file_format4(file, &fmt4)) return rc;
  std::vector<Enc> ep(P), ef((size_t)nf * 3);
  for (uint32_t s = 0; s < P; s++) ep[s].len = (uint8_t)enc_int(s, fmt4, &ep[s].type, ep[s].data);
  for (uint32_t i = 0; i < nf; i++) {
    ef[3 * i].len = (uint8_t)enc_real(freqs[i], fmt4, &ef[3 * i].type, ef[3 * i].data);
    ef[3 * i + 1].len = (uint8_t)enc_real(wavenums[i], fmt4, &ef[3 * i + 1].type, ef[3 * i + 1].data);
    ef[3 * i + 2].len = (uint8_t)enc_real(times ? times[i] : 0.0, fmt4, &ef[3 * i + 2].type, ef[3 * i + 2].data);
  }
  const size_t per = (size_t)P * P;
  return append_rows_impl(file, root, nf, per, threads, first_rowid, [&](size_t i, size_t j, uint8_t *out) -> size_t {
    const Enc &e_s = ep[j / P], &e_d = ep[j % P], &e_f = ef[3 * i], &e_w = ef[3 * i + 1], &e_t = ef[3 * i + 2];
    const double *x = x_patch + 2 * (i * per + j);
    out[0] = 8; out[1] = e_s.type; out[2] = e_d.type; out[3] = e_f.type; out[4] = e_w.type; out[7] = e_t.type;
    uint8_t *b = out + 8;
    memcpy(b, e_s.data, e_s.len); b += e_s.len;
    memcpy(b, e_d.data, e_d.len); b += e_d.len;
    memcpy(b, e_f.data, e_f.len); b += e_f.len;
    memcpy(b, e_w.data, e_w.len); b += e_w.len;
    b += enc_real(x[0], fmt4, &out[5], b);
    b += enc_real(x[1], fmt4, &out[6], b);
    memcpy(b, e_t.data, e_t.len); b += e_t.len;
    return (size_t)(b - out);
  });
}

// patch_to_patch_imp_resp (database.py:47-53): rows (source_patch, dest_patch, time, displacement) in
// (source, dest, time) order -- which is also the order of its index (source_patch, dest_patch, time) when
// the times ascend, so the index entries are the rows' keys with their rowids.
int append_imp_resp_rows(const char *file, uint32_t root, uint32_t P, uint32_t nt, const double *times, const double *h,
                         int threads, uint64_t *first_rowid) {
  if (!P || !nt) return 1;
  bool fmt4 = false;
  if (int rc = file_format4(file, &fmt4)) return rc;
  std::vector<Enc> ep(P), et(nt);
  for (uint32_t s = 0; s < P; s++) ep[s].len = (uint8_t)enc_int(s, fmt4, &ep[s].type, ep[s].data);
  for (uint32_t k = 0; k < nt; k++) et[k].len = (uint8_t)enc_real(times[k], fmt4, &et[k].type, et[k].data);
  const size_t per_src = (size_t)P * nt;
  return append_rows_impl(file, root, P, per_src, threads, first_rowid, [&](size_t s, size_t j, uint8_t *out) -> size_t {
    const size_t d = j / nt, k = j % nt;
    const Enc &e_s = ep[s], &e_d = ep[d], &e_t = et[k];
    out[0] = 5; out[1] = e_s.type; out[2] = e_d.type; out[3] = e_t.type;
    uint8_t *b = out + 5;
    memcpy(b, e_s.data, e_s.len); b += e_s.len;
    memcpy(b, e_d.data, e_d.len); b += e_d.len;
    memcpy(b, e_t.data, e_t.len); b += e_t.len;
    b += enc_real(h[s * per_src + j], fmt4, &out[4], b);
    return (size_t)(b - out);
  });
}

int build_imp_resp_index(const char *file, uint64_t first_rowid, uint32_t P, uint32_t nt, const double *times, int threads,
                         uint32_t *root_out) {
  if (!P || !nt) return 1;
  for (uint32_t k = 0; k < nt; k++)
    if (std::isnan(times[k]) || (k && !(times[k] > times[k - 1]))) return 1;      // not ascending: let SQL sort
  bool fmt4 = false;
  if (int rc = file_format4(file, &fmt4)) return rc;
  std::vector<Enc> ep(P), et(nt);
  for (uint32_t s = 0; s < P; s++) ep[s].len = (uint8_t)enc_int(s, fmt4, &ep[s].type, ep[s].data);
  for (uint32_t k = 0; k < nt; k++) et[k].len = (uint8_t)enc_real(times[k], fmt4, &et[k].type, et[k].data);
  return build_index_impl(file, (size_t)P * P, nt, threads, root_out, [&](size_t pr, size_t k, uint8_t *p) -> size_t {
    const Enc &e_s = ep[pr / P], &e_d = ep[pr % P], &e_t = et[k];
    p[0] = 5; p[1] = e_s.type; p[2] = e_d.type; p[3] = e_t.type;
    uint8_t *d = p + 5;
    memcpy(d, e_s.data, e_s.len); d += e_s.len;
    memcpy(d, e_d.data, e_d.len); d += e_d.len;
    memcpy(d, e_t.data, e_t.len); d += e_t.len;
    d += enc_int((int64_t)(first_rowid + pr * (uint64_t)nt + k), fmt4, &p[4], d);
    return (size_t)(d - p);
  });
}

}  // namespace dbpages
}  // namespace cnld
