// Inter-GPU plumbing of the frequency sweep through the C ABI (cnld_b200_comm_*), no PyTorch.
//
// The reference's only "distribution" is a process pool over frequencies with a lock around the sqlite
// appends (cnld/scripts/generate_db.py:229-242).  Here one process per GPU takes a shard of the
// frequency list; nothing is exchanged inside a frequency.  What does cross NVLink: rank 0's
// frequency-independent setup (mesh, FEM operators, loads), broadcast once, and the patch responses,
// gathered once at the end -- both with NCCL (ncclBroadcast, grouped ncclSend/ncclRecv), called from here
// so the sweep driver needs no torch.distributed.  libnccl.so.2 is loaded at run time (the image has
// 2.27.3 system-wide and 2.28.9 inside torch; either serves).  The NCCL unique id travels from rank 0 to
// the other ranks over one TCP connection each (address / port from the launcher's MASTER_ADDR /
// MASTER_PORT convention).
#include <arpa/inet.h>
#include <dlfcn.h>
#include <nccl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <thread>

#include "../../include/cnld_b200.h"
#include "common.cuh"

using cnld::set_error;

namespace {
struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};
Nccl *nccl() {
  static Nccl n;
  static bool tried = false;
  if (tried) return n.lib ? &n : nullptr;
  tried = true;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  if (!n.lib) return nullptr;
#define CNLD_SYM(f) n.f = (decltype(n.f))dlsym(n.lib, "nccl" #f)
  CNLD_SYM(GetUniqueId); CNLD_SYM(CommInitRank); CNLD_SYM(CommDestroy); CNLD_SYM(Broadcast); CNLD_SYM(AllReduce);
  CNLD_SYM(Send); CNLD_SYM(Recv); CNLD_SYM(GroupStart); CNLD_SYM(GroupEnd); CNLD_SYM(GetErrorString); CNLD_SYM(GetVersion);
#undef CNLD_SYM
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.Broadcast || !n.AllReduce || !n.Send || !n.Recv ||
      !n.GroupStart || !n.GroupEnd || !n.GetErrorString) {
    dlclose(n.lib);
    n.lib = nullptr;
    return nullptr;
  }
  return &n;
}

bool send_all(int fd, const void *buf, size_t n) {
  const char *p = (const char *)buf;
  while (n) { ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL); if (k <= 0) return false; p += k; n -= (size_t)k; }
  return true;
}
bool recv_all(int fd, void *buf, size_t n) {
  char *p = (char *)buf;
  while (n) { ssize_t k = ::recv(fd, p, n, 0); if (k <= 0) return false; p += k; n -= (size_t)k; }
  return true;
}

// rank 0 -> every other rank: `bytes` of payload over TCP (the NCCL unique id)
int tcp_share(int rank, int world, const char *addr, int port, void *payload, size_t bytes) {
  if (world <= 1) return 0;
  if (rank == 0) {
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    sa.sin_port = htons((uint16_t)port);
    if (ls < 0 || bind(ls, (sockaddr *)&sa, sizeof(sa)) != 0 || listen(ls, world) != 0) {
      set_error("comm: rank 0 cannot listen on port %d (set CNLD_B200_COMM_PORT)", port);
      if (ls >= 0) close(ls);
      return -1;
    }
    int served = 0;
    timeval tv{120, 0};
    setsockopt(ls, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    while (served < world - 1) {
      int fd = accept(ls, nullptr, nullptr);
      if (fd < 0) { set_error("comm: only %d of %d ranks connected to rank 0", served, world - 1); close(ls); return -1; }
      const bool ok = send_all(fd, payload, bytes);
      close(fd);
      if (!ok) { set_error("comm: sending the rendezvous payload failed"); close(ls); return -1; }
      served++;
    }
    close(ls);
    return 0;
  }
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  char ps[16];
  snprintf(ps, sizeof(ps), "%d", port);
  if (getaddrinfo(addr, ps, &hints, &res) != 0 || !res) { set_error("comm: cannot resolve %s", addr); return -1; }
  for (int attempt = 0; attempt < 600; attempt++) {      // rank 0 may still be building its setup
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd >= 0 && connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
      const bool ok = recv_all(fd, payload, bytes);
      close(fd);
      freeaddrinfo(res);
      if (!ok) { set_error("comm: receiving the rendezvous payload failed"); return -1; }
      return 0;
    }
    if (fd >= 0) close(fd);
    std::this_thread::sleep_for(std::chrono::milliseconds(200));
  }
  freeaddrinfo(res);
  set_error("comm: rank %d could not reach rank 0 at %s:%d", rank, addr, port);
  return -1;
}
}  // namespace

struct cnld_comm {
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t st = nullptr;
  Nccl *N = nullptr;
  double bytes_moved = 0.0, seconds = 0.0;   // payload through NCCL and the device time it took (events)
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

#define CNLD_NCCL(c, call)                                                                         \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) { set_error("%s -> %s", #call, (c)->N->GetErrorString(r__)); return -1; } \
  } while (0)

extern "C" {

cnld_comm *cnld_b200_comm_create(int rank, int world, int device, const char *addr, int port) {
  if (world < 1 || rank < 0 || rank >= world) { set_error("comm: bad rank %d / world %d", rank, world); return nullptr; }
  cnld_comm *c = new cnld_comm();
  c->rank = rank; c->world = world; c->device = device;
  if (world == 1) return c;                      // single rank: every collective is a no-op
  if (cnld_b200_device_count() <= device) { set_error("comm: no CUDA device %d (NCCL needs one GPU per rank)", device); delete c; return nullptr; }
  c->N = nccl();
  if (!c->N) { set_error("comm: libnccl.so.2 could not be loaded: %s", dlerror()); delete c; return nullptr; }
  ncclUniqueId id;
  memset(&id, 0, sizeof(id));
  bool ok = cudaSetDevice(device) == cudaSuccess;
  if (ok && rank == 0) ok = c->N->GetUniqueId(&id) == ncclSuccess;
  ok = ok && tcp_share(rank, world, addr && *addr ? addr : "127.0.0.1", port, &id, sizeof(id)) == 0;
  ok = ok && c->N->CommInitRank(&c->comm, world, id, rank) == ncclSuccess;
  ok = ok && cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) == cudaSuccess &&
       cudaEventCreate(&c->e0) == cudaSuccess && cudaEventCreate(&c->e1) == cudaSuccess;
  if (!ok) {
    if (!*cnld_b200_last_error()) set_error("comm: NCCL communicator setup failed on rank %d", rank);
    cnld_b200_comm_destroy(c);
    return nullptr;
  }
  return c;
}

void cnld_b200_comm_destroy(cnld_comm *c) {
  if (!c) return;
  if (c->comm) { cudaSetDevice(c->device); c->N->CommDestroy(c->comm); }
  if (c->st) cudaStreamDestroy(c->st);
  if (c->e0) cudaEventDestroy(c->e0);
  if (c->e1) cudaEventDestroy(c->e1);
  delete c;
}

int cnld_b200_comm_rank(const cnld_comm *c) { return c ? c->rank : -1; }
int cnld_b200_comm_world(const cnld_comm *c) { return c ? c->world : -1; }

/* host buffer of `bytes` bytes: on return every rank holds root's content (staged through the device, NVLink) */
int cnld_b200_comm_bcast(cnld_comm *c, void *buf, size_t bytes, int root) {
  if (!c) { set_error("null communicator"); return -1; }
  if (c->world == 1 || bytes == 0) return 0;
  CNLD_CK(cudaSetDevice(c->device));
  void *d = nullptr;
  CNLD_CK(cudaMalloc(&d, bytes));
  int rc = -1;
  do {
    if (c->rank == root && cudaMemcpyAsync(d, buf, bytes, cudaMemcpyHostToDevice, c->st) != cudaSuccess) break;
    cudaEventRecord(c->e0, c->st);
    ncclResult_t r = c->N->Broadcast(d, d, bytes, ncclChar, root, c->comm, c->st);
    cudaEventRecord(c->e1, c->st);
    if (r != ncclSuccess) { set_error("ncclBroadcast -> %s", c->N->GetErrorString(r)); rc = -2; break; }
    if (c->rank != root && cudaMemcpyAsync(buf, d, bytes, cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
    if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, c->e0, c->e1);
    c->bytes_moved += (double)bytes; c->seconds += ms * 1e-3;
    rc = 0;
  } while (0);
  if (rc == -1) set_error("comm_bcast: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
  return rc ? -1 : 0;
}

/* every rank contributes `bytes` bytes; root receives world * bytes in rank order (recv may be NULL elsewhere) */
int cnld_b200_comm_gather(cnld_comm *c, const void *send, size_t bytes, void *recv, int root) {
  if (!c) { set_error("null communicator"); return -1; }
  if (c->world == 1) { if (recv && recv != send && bytes) memcpy(recv, send, bytes); return 0; }
  if (bytes == 0) return 0;
  CNLD_CK(cudaSetDevice(c->device));
  char *ds = nullptr, *dr = nullptr;
  CNLD_CK(cudaMalloc(&ds, bytes));
  int rc = -1;
  do {
    if (c->rank == root && cudaMalloc(&dr, bytes * c->world) != cudaSuccess) break;
    if (cudaMemcpyAsync(ds, send, bytes, cudaMemcpyHostToDevice, c->st) != cudaSuccess) break;
    cudaEventRecord(c->e0, c->st);
    ncclResult_t r = c->N->GroupStart();
    if (r == ncclSuccess && c->rank != root) r = c->N->Send(ds, bytes, ncclChar, root, c->comm, c->st);
    if (c->rank == root)
      for (int p = 0; p < c->world && r == ncclSuccess; p++)
        if (p != root) r = c->N->Recv(dr + (size_t)p * bytes, bytes, ncclChar, p, c->comm, c->st);
    if (r == ncclSuccess) r = c->N->GroupEnd(); else c->N->GroupEnd();
    cudaEventRecord(c->e1, c->st);
    if (r != ncclSuccess) { set_error("NCCL gather -> %s", c->N->GetErrorString(r)); rc = -2; break; }
    if (c->rank == root) {
      if (cudaMemcpyAsync(dr + (size_t)root * bytes, ds, bytes, cudaMemcpyDeviceToDevice, c->st) != cudaSuccess) break;
      if (cudaMemcpyAsync(recv, dr, bytes * c->world, cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
    }
    if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, c->e0, c->e1);
    c->bytes_moved += (double)bytes * (c->rank == root ? c->world - 1 : 1); c->seconds += ms * 1e-3;
    rc = 0;
  } while (0);
  if (rc == -1) set_error("comm_gather: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(ds); cudaFree(dr);
  return rc ? -1 : 0;
}

/* element-wise max over ranks of n doubles (in place); with n = 0 a barrier */
int cnld_b200_comm_allreduce_max(cnld_comm *c, double *vals, int n) {
  if (!c) { set_error("null communicator"); return -1; }
  if (c->world == 1) return 0;
  CNLD_CK(cudaSetDevice(c->device));
  const int m = n > 0 ? n : 1;
  double *d = nullptr, zero = 0.0;
  CNLD_CK(cudaMalloc(&d, sizeof(double) * m));
  int rc = -1;
  do {
    if (cudaMemcpyAsync(d, n > 0 ? vals : &zero, sizeof(double) * m, cudaMemcpyHostToDevice, c->st) != cudaSuccess) break;
    ncclResult_t r = c->N->AllReduce(d, d, m, ncclDouble, ncclMax, c->comm, c->st);
    if (r != ncclSuccess) { set_error("ncclAllReduce -> %s", c->N->GetErrorString(r)); rc = -2; break; }
    if (n > 0 && cudaMemcpyAsync(vals, d, sizeof(double) * m, cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
    if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc == -1) set_error("comm_allreduce: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
  return rc ? -1 : 0;
}

int cnld_b200_comm_barrier(cnld_comm *c) { return cnld_b200_comm_allreduce_max(c, nullptr, 0); }

/* out[3]: payload bytes this rank moved through NCCL, device seconds of those collectives, NCCL version */
int cnld_b200_comm_stats(cnld_comm *c, double *out) {
  if (!c || !out) { set_error("null argument"); return -1; }
  out[0] = c->bytes_moved; out[1] = c->seconds; out[2] = 0.0;
  int v = 0;
  if (c->N && c->N->GetVersion && c->N->GetVersion(&v) == ncclSuccess) out[2] = v;
  return 0;
}

}  // extern "C"
