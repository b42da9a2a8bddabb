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
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cnld_b200.h"

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
};
constexpr int SQLITE_OK = 0, SQLITE_DONE = 101;

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
    if (!a.open || !a.close || !a.exec || !a.prepare || !a.bind_int64 || !a.bind_double || !a.step || !a.reset ||
        !a.finalize || !a.errmsg) {
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
    for (cnld_uint i = 0; i < nf && !rc; i++) {
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
      if (!nodes) { set_error("db_append_block: node coordinates are needed with x_node"); rc = -2; break; }
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
  return (db.exec("DROP INDEX IF EXISTS patch_to_patch_freq_resp_index") || db.exec("DROP INDEX IF EXISTS patch_to_node_freq_resp_index")) ? -1 : 0;
}
int cnld_b200_db_finish(const char *file) {
  Db db;
  if (db.open(file)) return -1;
  if (db.exec("PRAGMA cache_size=-1048576")) return -1;
  return (db.exec("CREATE INDEX IF NOT EXISTS patch_to_patch_freq_resp_index ON patch_to_patch_freq_resp (source_patch, dest_patch, frequency)") ||
          db.exec("CREATE INDEX IF NOT EXISTS patch_to_node_freq_resp_index ON patch_to_node_freq_resp (source_patch, node_id, frequency)"))
             ? -1 : 0;
}

}  // extern "C"
