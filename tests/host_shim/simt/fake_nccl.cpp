// fake_nccl.cpp — an in-process stand-in for libnccl.so.2, for the emulated multi-GPU test only (tests/test_emulated_mgpu_cpu.py):
// the "ranks" are threads of one process, each with its own OxcContext of the SIMT-emulated library; a communicator is a group
// of threads that rendezvous on a barrier.  Implements exactly the entry points oxcull.cu dlsym()s.
#include <pthread.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "nccl.h"

namespace {
struct Group {
  int n = 0, joined = 0;
  pthread_barrier_t bar;
  std::vector<const void*> send;
  std::vector<std::vector<uint8_t>> tmp;
};
struct Comm { Group* g; int rank; };
std::mutex g_mu;
std::map<std::string, Group*> g_groups;
uint64_t g_next_id = 1;

size_t type_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    default: return 8;
  }
}
template <typename T>
void reduce(T* acc, const T* v, size_t n, ncclRedOp_t op) {
  for (size_t i = 0; i < n; i++) {
    switch (op) {
      case ncclMax: acc[i] = v[i] > acc[i] ? v[i] : acc[i]; break;
      case ncclMin: acc[i] = v[i] < acc[i] ? v[i] : acc[i]; break;
      case ncclSum: acc[i] = acc[i] + v[i]; break;
      default: acc[i] = acc[i] * v[i]; break;
    }
  }
}
} // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(id, 0, sizeof *id);
  const uint64_t v = g_next_id++;
  memcpy(id->internal, &v, sizeof v);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
  Group* g;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const std::string key(id.internal, sizeof id.internal);
    Group*& slot = g_groups[key];
    if (!slot) {
      slot = new Group();
      slot->n = nranks;
      pthread_barrier_init(&slot->bar, nullptr, (unsigned)nranks);
      slot->send.assign((size_t)nranks, nullptr);
      slot->tmp.resize((size_t)nranks);
    }
    g = slot;
    g->joined++;
  }
  *comm = reinterpret_cast<ncclComm_t>(new Comm{g, rank});
  pthread_barrier_wait(&g->bar); // collective, like the real one
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c) { delete reinterpret_cast<Comm*>(c); return ncclSuccess; }
ncclResult_t ncclCommCount(const ncclComm_t c, int* n) { *n = reinterpret_cast<const Comm*>(c)->g->n; return ncclSuccess; }
ncclResult_t ncclCommUserRank(const ncclComm_t c, int* r) { *r = reinterpret_cast<const Comm*>(c)->rank; return ncclSuccess; }
const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "error (fake nccl)"; }

ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t c, cudaStream_t) {
  Comm* cm = reinterpret_cast<Comm*>(c);
  Group* g = cm->g;
  const size_t bytes = count * type_size(type);
  g->send[(size_t)cm->rank] = send;
  pthread_barrier_wait(&g->bar); // every rank's input is published (in place: nobody has written yet)
  std::vector<uint8_t>& acc = g->tmp[(size_t)cm->rank];
  acc.assign(static_cast<const uint8_t*>(g->send[0]), static_cast<const uint8_t*>(g->send[0]) + bytes);
  for (int r = 1; r < g->n; r++) {
    if (type == ncclUint32) reduce(reinterpret_cast<uint32_t*>(acc.data()), static_cast<const uint32_t*>(g->send[(size_t)r]), count, op);
    else if (type == ncclUint64) reduce(reinterpret_cast<uint64_t*>(acc.data()), static_cast<const uint64_t*>(g->send[(size_t)r]), count, op);
    else if (type == ncclInt32) reduce(reinterpret_cast<int32_t*>(acc.data()), static_cast<const int32_t*>(g->send[(size_t)r]), count, op);
    else return ncclUnhandledCudaError;
  }
  pthread_barrier_wait(&g->bar); // every rank has read every input
  memcpy(recv, acc.data(), bytes);
  pthread_barrier_wait(&g->bar);
  return ncclSuccess;
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, ncclDataType_t type, ncclComm_t c, cudaStream_t) {
  Comm* cm = reinterpret_cast<Comm*>(c);
  Group* g = cm->g;
  const size_t bytes = count * type_size(type);
  g->send[(size_t)cm->rank] = send;
  pthread_barrier_wait(&g->bar);
  for (int r = 0; r < g->n; r++) memcpy(static_cast<uint8_t*>(recv) + (size_t)r * bytes, g->send[(size_t)r], bytes);
  pthread_barrier_wait(&g->bar);
  return ncclSuccess;
}

} // extern "C"
