// Bulk writer for cnl-dyna's sqlite database (C ABI: cnld_b200_db_*).
//
// The reference appends the results of ONE source patch of ONE frequency per call, through pandas /
// executemany under a process-wide lock (cnld/scripts/generate_db.py:101-130, cnld/database.py:144-194):
// Python tuples for N x P = 2.13 M rows per frequency at the 4x4 array.  A GPU that solves seven
// frequencies a second needs the rows written without the interpreter: this file binds libsqlite3
// directly (dlopen -- the image ships the library but no header; the five prototypes below are sqlite's
// stable C API), prepares each INSERT once, and writes a whole block of frequencies in one transaction
// straight from the arrays the sweep returned.  Tables, columns and row contents are the reference's
// (database.py:22-92); only the two freq-resp indexes are built after the bulk load instead of being
// maintained row by row (cnld_b200_db_finish), which leaves the same schema in the finished file.
#include <dlfcn.h>

#include <unistd.h>

#include <atomic>
#include <climits>
#include <cstdlib>
#include <map>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cnld_b200.h"
#include "dbpages.h"

namespace cnld {
void set_error(const char *fmt, ...);
}
using cnld::set_error;

namespace {
struct sqlite3;
struct sqlite3_stmt;
struct Api {
  void *lib = nullptr;
  int (*open)(const char *, sqlite3 **) = nullptr;
  int (*close)(sqlite3 *) = nullptr;
  int (*exec)(sqlite3 *, const char *, int (*)(void *, int, char **, char **), void *, char **) = nullptr;
  int (*prepare)(sqlite3 *, const char *, int, sqlite3_stmt **, const char **) = nullptr;
  int (*bind_int64)(sqlite3_stmt *, int, long long) = nullptr;
  int (*bind_double)(sqlite3_stmt *, int, double) = nullptr;
  int (*step)(sqlite3_stmt *) = nullptr;
  int (*reset)(sqlite3_stmt *) = nullptr;
  int (*finalize)(sqlite3_stmt *) = nullptr;
  const char *(*errmsg)(sqlite3 *) = nullptr;
  int (*busy_timeout)(sqlite3 *, int) = nullptr;
  long long (*column_int64)(sqlite3_stmt *, int) = nullptr;
};
constexpr int SQLITE_OK = 0, SQLITE_ROW = 100, SQLITE_DONE = 101;

Api *api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
    auto sym = [&](const char *n) { return dlsym(a.lib, n); };
    a.open = (decltype(a.open))sym("sqlite3_open");
    a.close = (decltype(a.close))sym("sqlite3_close");
    a.exec = (decltype(a.exec))sym("sqlite3_exec");
    a.prepare = (decltype(a.prepare))sym("sqlite3_prepare_v2");
    a.bind_int64 = (decltype(a.bind_int64))sym("sqlite3_bind_int64");
    a.bind_double = (decltype(a.bind_double))sym("sqlite3_bind_double");
    a.step = (decltype(a.step))sym("sqlite3_step");
    a.reset = (decltype(a.reset))sym("sqlite3_reset");
    a.finalize = (decltype(a.finalize))sym("sqlite3_finalize");
    a.errmsg = (decltype(a.errmsg))sym("sqlite3_errmsg");
    a.busy_timeout = (decltype(a.busy_timeout))sym("sqlite3_busy_timeout");
    a.column_int64 = (decltype(a.column_int64))sym("sqlite3_column_int64");
    if (!a.open || !a.close || !a.exec || !a.prepare || !a.bind_int64 || !a.bind_double || !a.step || !a.reset ||
        !a.finalize || !a.errmsg || !a.column_int64) {
      dlclose(a.lib);
      a.lib = nullptr;
    }
  });
  return a.lib ? &a : nullptr;
}

struct Db {
  Api *A = nullptr;
  sqlite3 *con = nullptr;
  ~Db() { if (con) A->close(con); }
  int open(const char *file) {
    A = api();
    if (!A) { set_error("libsqlite3.so.0 could not be loaded: %s", dlerror()); return -1; }
    if (A->open(file, &con) != SQLITE_OK) { set_error("sqlite: cannot open %s: %s", file, con ? A->errmsg(con) : "?"); return -1; }
    if (A->busy_timeout) A->busy_timeout(con, 60000);
    return 0;
  }
  int exec(const char *sql) {
    char *msg = nullptr;
    if (A->exec(con, sql, nullptr, nullptr, &msg) != SQLITE_OK) { set_error("sqlite: %s: %s", sql, msg ? msg : A->errmsg(con)); return -1; }
    return 0;
  }
};
// first column of the first row of `sql` as an integer (SQL NULL / no row: 0)
int query_int(Db &db, const char *sql, long long *out) {
  sqlite3_stmt *st = nullptr;
  if (db.A->prepare(db.con, sql, -1, &st, nullptr) != SQLITE_OK) { set_error("sqlite: %s: %s", sql, db.A->errmsg(db.con)); return -1; }
  const int rc = db.A->step(st);
  *out = rc == SQLITE_ROW ? db.A->column_int64(st, 0) : 0;
  db.A->finalize(st);
  if (rc != SQLITE_ROW && rc != SQLITE_DONE) { set_error("sqlite: %s: %s", sql, db.A->errmsg(db.con)); return -1; }
  return 0;
}
struct Stmt {
  Api *A;
  sqlite3_stmt *s = nullptr;
  explicit Stmt(Api *a) : A(a) {}
  ~Stmt() { if (s) A->finalize(s); }
};

const char *NODE_TABLE =
    "CREATE TABLE patch_to_node_freq_resp (source_patch integer, node_id integer, x float, y float, z float, "
    "frequency float, wavenumber float, displacement_real float, displacement_imag float)";

// rows of the (frequency, source patch) pairs [q0, q1) of a block, in the reference's order
// (generate_db.py:114-127: per source patch all nodes)
int insert_node_rows(Db &db, const char *schema, size_t q0, size_t q1, cnld_uint P, cnld_uint N, const double *freqs,
                     const double *wavenums, const cnld_field *x_node, const double *nodes) {
  Api *A = db.A;
  Stmt pn(A);
  std::string sql = std::string("INSERT INTO ") + schema + "patch_to_node_freq_resp VALUES (?,?,?,?,?,?,?,?,?)";
  if (A->prepare(db.con, sql.c_str(), -1, &pn.s, nullptr) != SQLITE_OK) { set_error("sqlite: %s", A->errmsg(db.con)); return -1; }
  for (size_t q = q0; q < q1; q++) {
    const cnld_uint i = (cnld_uint)(q / P), s = (cnld_uint)(q % P);
    const cnld_field *xn = x_node + q * N;
    A->bind_int64(pn.s, 1, s);
    A->bind_double(pn.s, 6, freqs[i]);
    A->bind_double(pn.s, 7, wavenums[i]);
    for (cnld_uint v = 0; v < N; v++) {
      A->bind_int64(pn.s, 2, v);
      A->bind_double(pn.s, 3, nodes[3 * (size_t)v]);
      A->bind_double(pn.s, 4, nodes[3 * (size_t)v + 1]);
      A->bind_double(pn.s, 5, nodes[3 * (size_t)v + 2]);
      A->bind_double(pn.s, 8, xn[v].re);
      A->bind_double(pn.s, 9, xn[v].im);
      if (A->step(pn.s) != SQLITE_DONE) { set_error("sqlite: %s", A->errmsg(db.con)); return -1; }
      A->reset(pn.s);      // bindings persist across reset: columns 1, 6, 7 stay
    }
  }
  return 0;
}

// Parallel staging: SQLite has one writer per file and spends ~0.7 us per prepared-statement row, i.e.
// 1.5 s for the 2.13 M node rows of one C3 frequency.  Worker threads therefore encode slices of the block
// into private scratch databases (same table definition, no index), and the target appends them with
// INSERT ... SELECT, which SQLite executes as a raw record copy when the table definitions match and the
// target has no index to maintain (measured 6.5 M rows/s) -- the reason the indexes are dropped for the
// bulk load and rebuilt by cnld_b200_db_finish.
struct Stage {
  std::string path;
  size_t q0 = 0, q1 = 0;
  int rc = 0;
};
std::string scratch_dir(const char *file) {
  if (access("/dev/shm", W_OK) == 0) return "/dev/shm";
  std::string f(file);
  const size_t p = f.find_last_of('/');
  return p == std::string::npos ? "." : f.substr(0, p);
}

// Page path (dbpages.cpp): between cnld_b200_db_begin_bulk and cnld_b200_db_finish this process remembers
// which rowid ranges of patch_to_node_freq_resp it wrote as pages; if they cover the whole table, finish
// generates the index from the descriptors instead of letting SQLite scan and sort the table.
struct BulkState {
  bool covered = false;                              // every row of the node table is described by `blocks`
  std::vector<cnld::dbpages::NodeBlock> blocks;
  bool pcovered = false;                             // the same for patch_to_patch_freq_resp (blocks with N = P)
  std::vector<cnld::dbpages::NodeBlock> pblocks;
};
std::mutex g_bulk_mutex;
std::atomic<unsigned long long> g_rows_paged{0}, g_rows_sql{0}, g_index_direct{0}, g_index_sql{0}, g_imp_paged{0}, g_imp_sql{0},
    g_patch_paged{0}, g_patch_sql{0};
std::map<std::string, BulkState> g_bulk;
std::string canonical(const char *file) {
  char buf[PATH_MAX];
  return realpath(file, buf) ? std::string(buf) : std::string(file);
}
bool sql_only() {
  const char *m = getenv("CNLD_B200_DB_MODE");       // "sql": never write pages directly (tests, debugging)
  return m && strcmp(m, "sql") == 0;
}
const char *NODE_ROOT_SQL = "SELECT rootpage FROM sqlite_master WHERE type='table' AND name='patch_to_node_freq_resp'";
const char *NODE_NINDEX_SQL = "SELECT count(*) FROM sqlite_master WHERE type='index' AND tbl_name='patch_to_node_freq_resp'";
const char *PATCH_ROOT_SQL = "SELECT rootpage FROM sqlite_master WHERE type='table' AND name='patch_to_patch_freq_resp'";
const char *PATCH_NINDEX_SQL = "SELECT count(*) FROM sqlite_master WHERE type='index' AND tbl_name='patch_to_patch_freq_resp'";

int register_index(const char *file, const char *name, const char *table, uint32_t root, const char *create_sql);

// node rows (x_node may be null) and patch rows of a block as b-tree pages, under SQLite's exclusive lock.
// *node_done / *patch_done: written as pages; otherwise the caller inserts them through SQL.  -1 error
int append_block_pages(const char *file, cnld_uint nf, const double *freqs, const double *wavenums, cnld_uint P, cnld_uint N,
                       const cnld_field *x_node, const double *nodes, const cnld_field *x_patch, const double *times, int threads,
                       bool *node_done, bool *patch_done) {
  *node_done = *patch_done = false;
  if (sql_only()) return 0;
  Db lock;
  if (lock.open(file)) return -1;
  if (lock.exec("BEGIN EXCLUSIVE")) return -1;
  long long root = 0, nindex = 0, proot = 0, pnindex = 0;
  if (query_int(lock, NODE_ROOT_SQL, &root) || query_int(lock, NODE_NINDEX_SQL, &nindex) || query_int(lock, PATCH_ROOT_SQL, &proot) ||
      query_int(lock, PATCH_NINDEX_SQL, &pnindex))
    return -1;
  int rc = 1, prc = 1;
  uint64_t first = 0, pfirst = 0;
  if (x_node && root >= 2 && nindex == 0)
    rc = cnld::dbpages::append_node_rows(file, (uint32_t)root, nf, freqs, wavenums, P, N, (const double *)x_node, nodes,
                                         threads, &first);
  if (rc >= 0 && proot >= 2 && pnindex == 0)
    prc = cnld::dbpages::append_patch_rows(file, (uint32_t)proot, nf, freqs, wavenums, P, (const double *)x_patch, times, threads, &pfirst);
  lock.exec("ROLLBACK");      // nothing was written through this connection; its cache dies with it
  if (rc < 0 || prc < 0) return -1;
  *node_done = rc == 0;
  *patch_done = prc == 0;
  if (*node_done) g_rows_paged += (unsigned long long)nf * P * N;
  if (*patch_done) g_patch_paged += (unsigned long long)nf * P * P;
  std::lock_guard<std::mutex> g(g_bulk_mutex);
  auto it = g_bulk.find(canonical(file));
  if (it != g_bulk.end()) {
    cnld::dbpages::NodeBlock b;
    b.nf = nf; b.P = P;
    b.freqs.assign(freqs, freqs + nf);
    if (*node_done) {
      b.first_rowid = first; b.N = N;
      it->second.blocks.push_back(b);
    } else if (x_node) {
      it->second.covered = false;                    // these rows go in through SQL
    }
    if (*patch_done) {
      b.first_rowid = pfirst; b.N = P;
      it->second.pblocks.push_back(b);
    } else {
      it->second.pcovered = false;
    }
  }
  return 0;
}

// index (first two integer columns, frequency) of a table whose rows the blocks describe completely, generated
// from the descriptors and registered under `name`.  1: done, 0: not applicable (caller uses CREATE INDEX), -1 error
int direct_index(const char *file, std::vector<cnld::dbpages::NodeBlock> &blocks, const char *table, const char *name,
                 const char *create_sql) {
  if (blocks.empty() || sql_only()) return 0;
  std::sort(blocks.begin(), blocks.end(), [](const cnld::dbpages::NodeBlock &a, const cnld::dbpages::NodeBlock &b) { return a.first_rowid < b.first_rowid; });
  uint64_t next = 1;
  for (const auto &b : blocks) {
    if (b.first_rowid != next) return 0;
    next += (uint64_t)b.nf * b.P * b.N;
  }
  uint32_t root = 0;
  {
    Db lock;
    if (lock.open(file) || lock.exec("BEGIN EXCLUSIVE")) return -1;
    char q1[200], q2[200];
    snprintf(q1, sizeof(q1), "SELECT max(rowid) FROM %s", table);
    snprintf(q2, sizeof(q2), "SELECT count(*) FROM sqlite_master WHERE type='index' AND tbl_name='%s'", table);
    long long maxrow = 0, nindex = 0;
    if (query_int(lock, q1, &maxrow) || query_int(lock, q2, &nindex)) return -1;
    int rc = 1;
    if ((uint64_t)maxrow + 1 == next && nindex == 0) {
      int threads = (int)std::thread::hardware_concurrency();
      if (threads > 32) threads = 32;
      rc = cnld::dbpages::build_node_index(file, blocks, threads, &root);
    }
    lock.exec("ROLLBACK");
    if (rc < 0) return -1;
    if (rc != 0) return 0;
  }
  return register_index(file, name, table, root, create_sql) ? -1 : 1;
}

// the schema row CREATE INDEX would have written for a b-tree that already exists, and a new schema cookie
int register_index(const char *file, const char *name, const char *table, uint32_t root, const char *create_sql) {
  Db db;
  if (db.open(file)) return -1;
  long long cookie = 0;
  if (query_int(db, "PRAGMA schema_version", &cookie)) return -1;
  char sql[800], bump[64];
  snprintf(sql, sizeof(sql), "INSERT INTO sqlite_master (type, name, tbl_name, rootpage, sql) VALUES ('index', '%s', '%s', %u, '%s')",
           name, table, root, create_sql);
  snprintf(bump, sizeof(bump), "PRAGMA schema_version=%lld", cookie + 1);
  return (db.exec("PRAGMA writable_schema=ON") || db.exec("BEGIN IMMEDIATE") || db.exec(sql) || db.exec(bump) || db.exec("COMMIT") ||
          db.exec("PRAGMA writable_schema=OFF")) ? -1 : 0;
}
}  // namespace

extern "C" {

/* x_patch[nf][P(src)][P(dst)], x_node[nf][P(src)][N] (nullable), nodes[N][3], times[nf] (nullable, the
 * reference's time_solve column), job_ids[nf] (nullable: rows of the progress table to mark complete).
 * One transaction.  fast != 0: synchronous = OFF and an in-memory journal for this connection (a crash
 * mid-block can then leave the file unusable; the progress table is written inside the same transaction);
 * fast = n > 1 additionally stages the node rows with n threads (scratch databases in /dev/shm or next to
 * the file) and appends them by raw record copy -- needs the indexes dropped (cnld_b200_db_begin_bulk). */
int cnld_b200_db_append_block(const char *file, cnld_uint nf, const double *freqs, const double *wavenums, cnld_uint P,
                              const cnld_field *x_patch, const double *times, cnld_uint N, const cnld_field *x_node,
                              const double *nodes, const cnld_uint *job_ids, int fast) {
  if (!file || !freqs || !wavenums || !x_patch) { set_error("db_append_block: null argument"); return -1; }
  if (x_node && !nodes) { set_error("db_append_block: node coordinates are needed with x_node"); return -1; }
  bool patch_paged = false;
  if (fast > 1) {
    // node and patch rows first, as pages; the progress marks (and whatever the page path could not take)
    // follow in the SQL transaction below, so a frequency is marked complete only when its rows are in the file
    bool node_paged = false;
    if (append_block_pages(file, nf, freqs, wavenums, P, N, x_node, nodes, x_patch, times, fast, &node_paged, &patch_paged)) return -1;
    if (node_paged) x_node = nullptr;
  } else {
    std::lock_guard<std::mutex> g(g_bulk_mutex);
    auto it = g_bulk.find(canonical(file));
    if (it != g_bulk.end()) {
      if (x_node) it->second.covered = false;
      it->second.pcovered = false;
    }
  }
  Db db;
  if (db.open(file)) return -1;
  Api *A = db.A;
  if (fast && (db.exec("PRAGMA synchronous=OFF") || db.exec("PRAGMA journal_mode=MEMORY"))) return -1;
  if (db.exec("PRAGMA cache_size=-262144")) return -1;
  // stage the node rows in parallel (see Stage above), attach the scratch databases, then one transaction
  std::vector<Stage> stages;
  struct Cleanup {
    std::vector<Stage> &st;
    ~Cleanup() { for (auto &s : st) unlink(s.path.c_str()); }
  } cleanup{stages};
  const size_t nq = (size_t)nf * P;
  int nthreads = fast > 1 ? fast : 1;          // fast = 1: plain single-connection path; fast = n > 1: n staging threads
  if (nthreads > 8) nthreads = 8;              // SQLITE_MAX_ATTACHED is 10 by default
  if (x_node && nodes && nthreads > 1 && nq >= (size_t)nthreads) {
    static std::atomic<unsigned> serial{0};
    const std::string dir = scratch_dir(file);
    stages.resize(nthreads);
    for (int t = 0; t < nthreads; t++) {
      char name[96];
      snprintf(name, sizeof(name), "/cnld_b200_stage_%d_%u_%d.db", (int)getpid(), serial.fetch_add(1), t);
      stages[t].path = dir + name;
      stages[t].q0 = nq * t / nthreads;
      stages[t].q1 = nq * (t + 1) / nthreads;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
      th.emplace_back([&, t] {
        Stage &S = stages[t];
        unlink(S.path.c_str());
        Db sd;
        S.rc = sd.open(S.path.c_str()) || sd.exec("PRAGMA synchronous=OFF") || sd.exec("PRAGMA journal_mode=OFF") ||
               sd.exec("PRAGMA cache_size=-131072") || sd.exec(NODE_TABLE) || sd.exec("BEGIN") ||
               insert_node_rows(sd, "", S.q0, S.q1, P, N, freqs, wavenums, x_node, nodes) || sd.exec("COMMIT");
      });
    for (auto &t : th) t.join();
    for (auto &S : stages) if (S.rc) return -1;
    for (size_t t = 0; t < stages.size(); t++) {
      char sql[400];
      snprintf(sql, sizeof(sql), "ATTACH DATABASE '%s' AS stage%zu", stages[t].path.c_str(), t);
      if (db.exec(sql)) return -1;
    }
  }
  if (db.exec("BEGIN IMMEDIATE")) return -1;
  Stmt pp(A), pr(A);
  int rc = 0;
  do {
    if (A->prepare(db.con, "INSERT INTO patch_to_patch_freq_resp VALUES (?,?,?,?,?,?,?)", -1, &pp.s, nullptr) != SQLITE_OK) { rc = -1; break; }
    if (!patch_paged) g_patch_sql += (unsigned long long)nf * P * P;
    for (cnld_uint i = 0; i < nf && !rc && !patch_paged; i++) {
      const cnld_field *xp = x_patch + (size_t)i * P * P;
      for (cnld_uint s = 0; s < P && !rc; s++)
        for (cnld_uint d = 0; d < P; d++) {
          A->bind_int64(pp.s, 1, s);
          A->bind_int64(pp.s, 2, d);
          A->bind_double(pp.s, 3, freqs[i]);
          A->bind_double(pp.s, 4, wavenums[i]);
          A->bind_double(pp.s, 5, xp[(size_t)s * P + d].re);
          A->bind_double(pp.s, 6, xp[(size_t)s * P + d].im);
          A->bind_double(pp.s, 7, times ? times[i] : 0.0);
          if (A->step(pp.s) != SQLITE_DONE) { rc = -1; break; }
          A->reset(pp.s);
        }
    }
    if (rc) break;
    if (x_node) {
      g_rows_sql += (unsigned long long)nf * P * N;
      if (stages.empty()) {
        if (insert_node_rows(db, "", 0, (size_t)nf * P, P, N, freqs, wavenums, x_node, nodes)) { rc = -2; break; }
      } else {
        for (size_t t = 0; t < stages.size() && !rc; t++) {
          char sql[160];
          snprintf(sql, sizeof(sql), "INSERT INTO main.patch_to_node_freq_resp SELECT * FROM stage%zu.patch_to_node_freq_resp", t);
          if (db.exec(sql)) rc = -2;
        }
        if (rc) break;
      }
    }
    if (job_ids) {
      if (A->prepare(db.con, "UPDATE progress SET is_complete=1 WHERE job_id=?", -1, &pr.s, nullptr) != SQLITE_OK) { rc = -1; break; }
      for (cnld_uint i = 0; i < nf; i++) {
        A->bind_int64(pr.s, 1, job_ids[i]);
        if (A->step(pr.s) != SQLITE_DONE) { rc = -1; break; }
        A->reset(pr.s);
      }
    }
  } while (0);
  if (rc == -1) set_error("sqlite: %s", A->errmsg(db.con));
  if (rc) { db.exec("ROLLBACK"); return -1; }
  return db.exec("COMMIT");   // scratch databases are detached when the connection closes, then unlinked
}

/* Bulk-load bracket: begin drops the two freq-resp indexes (a resumed sweep may find them), finish
 * (re)builds them -- the finished file has the reference's schema (database.py:31-45). */
int cnld_b200_db_begin_bulk(const char *file) {
  Db db;
  if (db.open(file)) return -1;
  if (db.exec("DROP INDEX IF EXISTS patch_to_patch_freq_resp_index") || db.exec("DROP INDEX IF EXISTS patch_to_node_freq_resp_index")) return -1;
  long long maxrow = 0, pmaxrow = 0;
  if (query_int(db, "SELECT max(rowid) FROM patch_to_node_freq_resp", &maxrow) ||
      query_int(db, "SELECT max(rowid) FROM patch_to_patch_freq_resp", &pmaxrow))
    return -1;
  std::lock_guard<std::mutex> g(g_bulk_mutex);
  BulkState &st = g_bulk[canonical(file)];
  st.blocks.clear();
  st.pblocks.clear();
  st.covered = maxrow == 0;          // rows of an earlier run are not described: finish indexes through SQL then
  st.pcovered = pmaxrow == 0;
  return 0;
}
int cnld_b200_db_finish(const char *file) {
  BulkState st;
  {
    std::lock_guard<std::mutex> g(g_bulk_mutex);
    auto it = g_bulk.find(canonical(file));
    if (it != g_bulk.end()) { st = std::move(it->second); g_bulk.erase(it); }
  }
  const char *node_index_sql = "CREATE INDEX patch_to_node_freq_resp_index ON patch_to_node_freq_resp (source_patch, node_id, frequency)";
  const char *patch_index_sql = "CREATE INDEX patch_to_patch_freq_resp_index ON patch_to_patch_freq_resp (source_patch, dest_patch, frequency)";
  // the indexes straight from the block descriptors when they describe the whole table
  int direct = st.covered ? direct_index(file, st.blocks, "patch_to_node_freq_resp", "patch_to_node_freq_resp_index", node_index_sql) : 0;
  if (direct < 0) return -1;
  int pdirect = st.pcovered ? direct_index(file, st.pblocks, "patch_to_patch_freq_resp", "patch_to_patch_freq_resp_index", patch_index_sql) : 0;
  if (pdirect < 0) return -1;
  if (direct) g_index_direct++;
  Db db;
  if (db.open(file)) return -1;
  if (db.exec("PRAGMA cache_size=-1048576")) return -1;
  if (!pdirect && db.exec("CREATE INDEX IF NOT EXISTS patch_to_patch_freq_resp_index ON patch_to_patch_freq_resp (source_patch, dest_patch, frequency)")) return -1;
  if (!direct) {
    if (db.exec("CREATE INDEX IF NOT EXISTS patch_to_node_freq_resp_index ON patch_to_node_freq_resp (source_patch, node_id, frequency)")) return -1;
    g_index_sql++;
  }
  return 0;
}
/* Rows (source_patch, dest_patch, time, displacement) of patch_to_patch_imp_resp from h[P][P][nt] -- what
 * generate_db.py:143-164 appends through pandas after the sweep.  Into an empty table of a plain database the
 * rows and their index go in as pages (dbpages.cpp); otherwise prepared statements in one transaction. */
int cnld_b200_db_append_imp_resp(const char *file, cnld_uint P, cnld_uint nt, const double *times, const double *h, int threads) {
  if (!file || !times || !h || !P || !nt) { set_error("db_append_imp_resp: null argument"); return -1; }
  const char *index_sql = "CREATE INDEX patch_to_patch_imp_resp_index ON patch_to_patch_imp_resp (source_patch, dest_patch, time)";
  bool paged = false;
  if (!sql_only() && threads > 0) {
    long long root = 0, maxrow = 0, other = 0;
    {
      Db db;
      if (db.open(file)) return -1;
      if (query_int(db, "SELECT rootpage FROM sqlite_master WHERE type='table' AND name='patch_to_patch_imp_resp'", &root) ||
          query_int(db, "SELECT max(rowid) FROM patch_to_patch_imp_resp", &maxrow) ||
          query_int(db, "SELECT count(*) FROM sqlite_master WHERE type='index' AND tbl_name='patch_to_patch_imp_resp' AND "
                        "name<>'patch_to_patch_imp_resp_index'", &other))
        return -1;
      bool ascending = true;
      for (cnld_uint k = 1; k < nt; k++) ascending = ascending && times[k] > times[k - 1];
      if (root >= 2 && maxrow == 0 && other == 0 && ascending && times[0] == times[0]) {
        if (db.exec("DROP INDEX IF EXISTS patch_to_patch_imp_resp_index")) return -1;
      } else {
        root = 0;
      }
    }
    if (root >= 2) {
      Db lock;
      if (lock.open(file) || lock.exec("BEGIN EXCLUSIVE")) return -1;
      if (query_int(lock, "SELECT rootpage FROM sqlite_master WHERE type='table' AND name='patch_to_patch_imp_resp'", &root)) return -1;
      uint64_t first = 0;
      uint32_t iroot = 0;
      int rc = cnld::dbpages::append_imp_resp_rows(file, (uint32_t)root, P, nt, times, h, threads, &first);
      int rci = rc == 0 ? cnld::dbpages::build_imp_resp_index(file, first, P, nt, times, threads, &iroot) : 1;
      lock.exec("ROLLBACK");
      if (rc < 0 || rci < 0) return -1;
      paged = rc == 0;
      if (paged) g_imp_paged += (unsigned long long)P * P * nt;
      if (rci == 0 && register_index(file, "patch_to_patch_imp_resp_index", "patch_to_patch_imp_resp", iroot, index_sql)) return -1;
    }
  }
  Db db;
  if (db.open(file)) return -1;
  Api *A = db.A;
  if (!paged) {
    if (db.exec("PRAGMA synchronous=OFF") || db.exec("PRAGMA journal_mode=MEMORY") || db.exec("PRAGMA cache_size=-262144") || db.exec("BEGIN IMMEDIATE")) return -1;
    Stmt st(A);
    int rc = 0;
    if (A->prepare(db.con, "INSERT INTO patch_to_patch_imp_resp VALUES (?,?,?,?)", -1, &st.s, nullptr) != SQLITE_OK) rc = -1;
    for (cnld_uint s = 0; s < P && !rc; s++)
      for (cnld_uint d = 0; d < P && !rc; d++) {
        const double *hh = h + ((size_t)s * P + d) * nt;
        A->bind_int64(st.s, 1, s);
        A->bind_int64(st.s, 2, d);
        for (cnld_uint k = 0; k < nt; k++) {
          A->bind_double(st.s, 3, times[k]);
          A->bind_double(st.s, 4, hh[k]);
          if (A->step(st.s) != SQLITE_DONE) { rc = -1; break; }
          A->reset(st.s);
        }
      }
    if (rc) { set_error("sqlite: %s", A->errmsg(db.con)); db.exec("ROLLBACK"); return -1; }
    if (db.exec("COMMIT")) return -1;
    g_imp_sql += (unsigned long long)P * P * nt;
  }
  char sql[300];
  snprintf(sql, sizeof(sql), "CREATE INDEX IF NOT EXISTS %s", index_sql + strlen("CREATE INDEX "));
  return db.exec(sql) ? -1 : 0;
}
/* out[8]: node rows written as pages / through SQL, node indexes generated from descriptors / by CREATE INDEX,
 * impulse-response rows written as pages / through SQL, patch rows written as pages / through SQL (this
 * process, since load) */
int cnld_b200_db_counters(unsigned long long *out) {
  if (!out) { set_error("db_counters: null argument"); return -1; }
  out[0] = g_rows_paged; out[1] = g_rows_sql; out[2] = g_index_direct; out[3] = g_index_sql;
  out[4] = g_imp_paged; out[5] = g_imp_sql;
  out[6] = g_patch_paged; out[7] = g_patch_sql;
  return 0;
}

}  // extern "C"
