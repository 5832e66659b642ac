// Multi-GPU exchange steps of the C ABI (include/minimc_b200.h "multi-GPU"): one process per GPU, NCCL over
// NVLink / NVSwitch.  SURVEY.md 8(e): a fixed-source run needs one all-reduce of the integer tallies at its end; a
// k-eigenvalue generation needs one exchange step -- an all-gather of the ranks' fission-bank counts (global offsets
// and k), then only the parts of each rank's needed site range that live on other ranks move, as grouped
// ncclSend / ncclRecv.  The reference has no counterpart (its KEigenvalue::Solve is a stub, KEigenvalue.cpp:36-62, and
// its workers are threads of one process, FixedSource.cpp:22-36).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy a host process such as PyTorch already loaded, else
// the system's), so libminimc_b200.so loads and runs single-GPU without it.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>  // types and prototypes only

#include "../../include/minimc_b200.h"

namespace mmc {
int set_last_error(int status, const std::string& message);  // capi.cu
}

namespace {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string error;
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* lib = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(lib, name);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  });
  return api;
}

int fail(int status, const std::string& message) { return mmc::set_last_error(status, message); }

#define MMC_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    const ncclResult_t r_ = (expr);                                                                      \
    if (r_ != ncclSuccess) return fail(MMC_ERR_CUDA, std::string(#expr ": ") + nccl().GetErrorString(r_)); \
  } while (0)
#define MMC_CUDA_(expr)                                                                           \
  do {                                                                                            \
    const cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) return fail(MMC_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_)); \
  } while (0)

// floor(a * b / c) with a 128-bit product
uint64_t muldiv(uint64_t a, uint64_t b, uint64_t c) {
  return static_cast<uint64_t>(static_cast<unsigned __int128>(a) * b / c);
}

// Global fission-site range [first, first + count) that the next sources of `rank` are drawn from: source i uses site
// floor(i * M / N) for i in the rank's source range [rank*N/P, (rank+1)*N/P) (DESIGN.md "k-eigenvalue", step 3).
void needed_sites(uint64_t n_total, uint64_t m_total, int rank, int nranks, uint64_t& first, uint64_t& count) {
  const uint64_t i_lo = muldiv(static_cast<uint64_t>(rank), n_total, static_cast<uint64_t>(nranks));
  const uint64_t i_hi = muldiv(static_cast<uint64_t>(rank) + 1, n_total, static_cast<uint64_t>(nranks));
  first = count = 0;
  if (i_hi == i_lo || m_total == 0) return;
  first = muldiv(i_lo, m_total, n_total);
  count = muldiv(i_hi - 1, m_total, n_total) - first + 1;
}

}  // namespace

struct mmc_comm {
  ncclComm_t comm = nullptr;  // null for a single rank
  int rank = 0, nranks = 1, device = 0;
  uint64_t* d_words = nullptr;  // [2 + 2 * nranks]: what this rank contributes, then what every rank contributed
  uint64_t* h_words = nullptr;  // pinned mirror
  cudaStream_t stream = nullptr;
  // device time of the mmc_bank_exchange calls so far (CUDA events on their stream, resolved lazily: the call itself
  // does not wait for its sends and receives)
  double exchange_ms = 0;
  bool pending = false;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  void resolve() {
    if (!pending) return;
    float ms = 0;
    if (cudaEventSynchronize(ev_end) == cudaSuccess && cudaEventElapsedTime(&ms, ev_begin, ev_end) == cudaSuccess) exchange_ms += ms;
    pending = false;
  }
};

extern "C" {

int mmc_comm_unique_id(void* id, size_t cap) {
  if (!id || cap < MMC_COMM_ID_BYTES) return fail(MMC_ERR_INVALID, "mmc_comm_unique_id: buffer smaller than MMC_COMM_ID_BYTES");
  static_assert(sizeof(ncclUniqueId) <= MMC_COMM_ID_BYTES, "ncclUniqueId does not fit MMC_COMM_ID_BYTES");
  if (!nccl().error.empty()) return fail(MMC_ERR_NO_DEVICE, nccl().error);
  ncclUniqueId u;
  MMC_NCCL(nccl().GetUniqueId(&u));
  std::memset(id, 0, MMC_COMM_ID_BYTES);
  std::memcpy(id, &u, sizeof(u));
  return MMC_OK;
}

int mmc_comm_create(int nranks, int rank, const void* id, int device, mmc_comm** out) {
  if (!out) return fail(MMC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MMC_ERR_INVALID, "bad rank / nranks");
  if (mmc_device_count() < 1) return fail(MMC_ERR_NO_DEVICE, "no CUDA device visible: minimc_b200 has no CPU transport path");
  if (device < 0) MMC_CUDA_(cudaGetDevice(&device));
  MMC_CUDA_(cudaSetDevice(device));
  auto* c = new mmc_comm;
  c->rank = rank;
  c->nranks = nranks;
  c->device = device;
  cudaError_t e = cudaMalloc(&c->d_words, (2 + 2 * static_cast<size_t>(nranks)) * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMallocHost(&c->h_words, (2 + 2 * static_cast<size_t>(nranks)) * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev_begin);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev_end);
  if (e != cudaSuccess) {
    mmc_comm_destroy(c);
    return fail(MMC_ERR_CUDA, std::string("mmc_comm_create: ") + cudaGetErrorString(e));
  }
  if (nranks > 1) {
    if (!id) {
      mmc_comm_destroy(c);
      return fail(MMC_ERR_INVALID, "mmc_comm_create: id is NULL");
    }
    if (!nccl().error.empty()) {
      mmc_comm_destroy(c);
      return fail(MMC_ERR_NO_DEVICE, nccl().error);
    }
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    const ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, u, rank);
    if (r != ncclSuccess) {
      c->comm = nullptr;
      mmc_comm_destroy(c);
      return fail(MMC_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
    }
  }
  if (nranks > 1) {
    // NCCL sets its channels up lazily, at the first collective and at the first send / receive between each ordered
    // pair of ranks (measured on B200s: 470 ms at the first exchange, 150 ms the first time a pair exchanged in the
    // other direction): make every connection now, not inside a generation's exchange step
    const NcclApi& api = nccl();
    ncclResult_t r = api.AllGather(c->d_words, c->d_words + 2, 2, ncclUint64, c->comm, c->stream);
    if (r == ncclSuccess) r = api.GroupStart();
    for (int peer = 0; peer < nranks && r == ncclSuccess; peer++) {
      if (peer == rank) continue;
      r = api.Send(c->d_words, 8, ncclChar, peer, c->comm, c->stream);
      if (r == ncclSuccess) r = api.Recv(c->d_words + 2 + 2 * peer, 8, ncclChar, peer, c->comm, c->stream);
    }
    if (r == ncclSuccess) r = api.GroupEnd();
    if (r == ncclSuccess) r = api.AllReduce(c->d_words, c->d_words, 2, ncclUint64, ncclSum, c->comm, c->stream);
    e = cudaStreamSynchronize(c->stream);
    if (r != ncclSuccess || e != cudaSuccess) {
      mmc_comm_destroy(c);
      return fail(MMC_ERR_CUDA, std::string("mmc_comm_create (connection warm-up): ") +
                                    (r != ncclSuccess ? api.GetErrorString(r) : cudaGetErrorString(e)));
    }
  }
  *out = c;
  return MMC_OK;
}

void mmc_comm_destroy(mmc_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm) nccl().CommDestroy(c->comm);
  if (c->ev_begin) cudaEventDestroy(c->ev_begin);
  if (c->ev_end) cudaEventDestroy(c->ev_end);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaFree(c->d_words);
  if (c->h_words) cudaFreeHost(c->h_words);
  delete c;
}

int mmc_comm_rank(const mmc_comm* c) { return c ? c->rank : 0; }
int mmc_comm_size(const mmc_comm* c) { return c ? c->nranks : 1; }
double mmc_comm_exchange_ms(mmc_comm* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  c->resolve();
  return c->exchange_ms;
}

int mmc_nccl_version(void) {
  int v = 0;
  if (nccl().error.empty() && nccl().GetVersion) nccl().GetVersion(&v);
  return v;
}

int mmc_tally_allreduce(mmc_comm* c, uint64_t* d_words, size_t n_words, void* stream) {
  if (!c) return fail(MMC_ERR_INVALID, "comm is NULL");
  if (n_words && !d_words) return fail(MMC_ERR_INVALID, "d_words is NULL");
  if (c->nranks == 1 || n_words == 0) return MMC_OK;
  MMC_CUDA_(cudaSetDevice(c->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  MMC_NCCL(nccl().AllReduce(d_words, d_words, n_words, ncclUint64, ncclSum, c->comm, s));
  return MMC_OK;
}

int mmc_exchange_plan(const uint64_t* counts, int nranks, int rank, uint64_t n_total, uint64_t* need_first,
                      uint64_t* need_count, uint64_t* send, uint64_t* recv) {
  if (!counts || !need_first || !need_count || !send || !recv || nranks < 1 || rank < 0 || rank >= nranks || n_total == 0)
    return fail(MMC_ERR_INVALID, "bad arguments to mmc_exchange_plan");
  std::vector<uint64_t> offsets(static_cast<size_t>(nranks) + 1, 0);
  for (int r = 0; r < nranks; r++) offsets[r + 1] = offsets[r] + counts[r];
  const uint64_t m_total = offsets[nranks];
  needed_sites(n_total, m_total, rank, nranks, *need_first, *need_count);
  for (int peer = 0; peer < nranks; peer++) {
    send[2 * peer] = send[2 * peer + 1] = recv[2 * peer] = recv[2 * peer + 1] = 0;
    // what `peer` needs from this rank's ordered bank (global sites [offsets[rank], offsets[rank + 1]))
    uint64_t p_first, p_count;
    needed_sites(n_total, m_total, peer, nranks, p_first, p_count);
    uint64_t lo = std::max(p_first, offsets[rank]), hi = std::min(p_first + p_count, offsets[rank + 1]);
    if (hi > lo) send[2 * peer] = lo - offsets[rank], send[2 * peer + 1] = hi - lo;
    // what this rank needs from `peer`
    lo = std::max(*need_first, offsets[peer]), hi = std::min(*need_first + *need_count, offsets[peer + 1]);
    if (hi > lo) recv[2 * peer] = lo - *need_first, recv[2 * peer + 1] = hi - lo;
  }
  return MMC_OK;
}

int mmc_bank_exchange(mmc_comm* c, const mmc_site* d_bank_local, const uint64_t* d_n_local, uint64_t local_status,
                      uint64_t n_total, mmc_site* d_slice, uint64_t slice_capacity, uint64_t* counts, uint64_t* statuses,
                      uint64_t* slice_first, uint64_t* slice_n, void* stream) {
  if (!c) return fail(MMC_ERR_INVALID, "comm is NULL");
  if (!d_bank_local || !d_n_local || !d_slice || !counts || !slice_first || !slice_n || n_total == 0)
    return fail(MMC_ERR_INVALID, "bad arguments to mmc_bank_exchange");
  MMC_CUDA_(cudaSetDevice(c->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  const int P = c->nranks;
  c->resolve();
  MMC_CUDA_(cudaEventRecord(c->ev_begin, s));
  // ---- all-gather of {sites banked, status word} of every rank: global offsets, k, and a collective error decision
  c->h_words[1] = local_status;
  MMC_CUDA_(cudaMemcpyAsync(c->d_words, d_n_local, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
  MMC_CUDA_(cudaMemcpyAsync(c->d_words + 1, c->h_words + 1, sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  if (P > 1) MMC_NCCL(nccl().AllGather(c->d_words, c->d_words + 2, 2, ncclUint64, c->comm, s));
  else MMC_CUDA_(cudaMemcpyAsync(c->d_words + 2, c->d_words, 2 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
  MMC_CUDA_(cudaMemcpyAsync(c->h_words + 2, c->d_words + 2, 2 * static_cast<size_t>(P) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  MMC_CUDA_(cudaStreamSynchronize(s));
  uint64_t any_status = 0, m_total = 0;
  for (int r = 0; r < P; r++) {
    counts[r] = c->h_words[2 + 2 * r];
    if (statuses) statuses[r] = c->h_words[3 + 2 * r];
    any_status |= c->h_words[3 + 2 * r];
    m_total += counts[r];
  }
  *slice_first = *slice_n = 0;
  // every rank sees the same words: all of them stop here together, nobody is left waiting in the exchange below
  if (any_status || m_total == 0) {
    MMC_CUDA_(cudaEventRecord(c->ev_end, s));
    c->pending = true;
    return MMC_OK;
  }
  std::vector<uint64_t> send(2 * static_cast<size_t>(P)), recv(2 * static_cast<size_t>(P));
  if (int st = mmc_exchange_plan(counts, P, c->rank, n_total, slice_first, slice_n, send.data(), recv.data())) return st;
  if (*slice_n > slice_capacity)
    return fail(MMC_ERR_CAPACITY, "mmc_bank_exchange: the rank's site range does not fit d_slice (raise bank_capacity_factor)");
  // ---- the pieces: this rank's own part is a device copy, the others grouped ncclSend / ncclRecv
  if (recv[2 * c->rank + 1])
    MMC_CUDA_(cudaMemcpyAsync(d_slice + recv[2 * c->rank], d_bank_local + send[2 * c->rank],
                              recv[2 * c->rank + 1] * sizeof(mmc_site), cudaMemcpyDeviceToDevice, s));
  if (P > 1) {
    MMC_NCCL(nccl().GroupStart());
    for (int peer = 0; peer < P; peer++) {
      if (peer == c->rank) continue;
      if (send[2 * peer + 1])
        MMC_NCCL(nccl().Send(d_bank_local + send[2 * peer], send[2 * peer + 1] * sizeof(mmc_site), ncclChar, peer, c->comm, s));
      if (recv[2 * peer + 1])
        MMC_NCCL(nccl().Recv(d_slice + recv[2 * peer], recv[2 * peer + 1] * sizeof(mmc_site), ncclChar, peer, c->comm, s));
    }
    MMC_NCCL(nccl().GroupEnd());
  }
  MMC_CUDA_(cudaEventRecord(c->ev_end, s));
  c->pending = true;
  return MMC_OK;
}

}  // extern "C"
