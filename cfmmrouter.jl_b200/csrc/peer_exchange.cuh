// peer_exchange.cuh -- sum-allreduce of [Ψ ; acc] (n_tokens+1 fp64) across the
// GPUs of one box, over NVLink peer memory (no NCCL on this path).
//
// Why not NCCL: the message is 8 KB .. 400 KB, i.e. purely latency-bound, and a
// sharded sweep kernel runs for only a few microseconds; NCCL's small-message
// latency would dominate the step.  Here every rank publishes its partial
// vector in a peer-mapped slot, raises one flag per peer, and pulls the other
// ranks' partials straight over NVLink, summing in rank order -- so every rank
// ends with the bitwise-identical vector (the replicated L-BFGS-B drivers stay
// in lock-step), with a single flag round per reduction.
//
// Slots are double-buffered by epoch parity: a rank can only start writing
// epoch e+2 after it saw every peer's flag for e+1, and a peer raises e+1 only
// after its epoch-e kernel (the reader of slot e) has completed in stream
// order, so no reader can still be on the slot that is being overwritten.
//
// Mapping: one process per GPU -> cudaIpc handles exchanged by the host side;
// several contexts in one process -> raw pointers + cudaDeviceEnablePeerAccess.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <unistd.h>

#include <cstring>
#include <string>

namespace cfmm {

constexpr int kMaxPeers = 16;
constexpr int kExchangeThreads = 512;

struct PeerHandle {
  cudaIpcMemHandle_t ipc;  // 64 B
  uint64_t raw_ptr;        // same-process shortcut
  int64_t len;             // doubles per slot
  int32_t pid;
  int32_t device;
};

// Layout of a rank's exchange buffer (all offsets in bytes from base):
//   [0, 1024)                 flags[kMaxPeers] (uint64, one 64 B line each)
//   [1024, 1024 + 2*len*8)    slot[0], slot[1]
struct ExchangeView {
  unsigned long long* flags;             // local: flags[p*8] = last epoch published by rank p
  double* my_slot[2];                    // local slots
  const double* peer_slot[kMaxPeers][2]; // peer-mapped (own entry = local)
  unsigned long long* peer_flags[kMaxPeers];  // peer-mapped flag arrays
  unsigned int* arrive;                  // local grid-arrival counter
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p,
                                               unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(
    const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// grid must be fully co-resident (it spins); launched with <= sm_count CTAs.
__global__ void __launch_bounds__(kExchangeThreads)
    peer_allreduce_kernel(ExchangeView x, double* __restrict__ data, int64_t len,
                          unsigned long long epoch) {
  const int par = (int)(epoch & 1ull);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;

  // phase 1: publish my partial
  double* slot = x.my_slot[par];
  for (int64_t j = tid; j < len; j += stride) slot[j] = data[j];
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(x.arrive, 1u);
    s_last = (prev == (unsigned int)(epoch * gridDim.x) - 1u);
  }
  __syncthreads();
  if (s_last && threadIdx.x < x.world) {
    // every CTA of this rank has published (their fences precede the atomic)
    __threadfence_system();
    st_release_sys(x.peer_flags[threadIdx.x] + 8 * x.rank, epoch);
  }

  // phase 2: wait for every rank's flag, then pull and sum in rank order
  if (threadIdx.x < x.world) {
    const unsigned long long* f = x.flags + 8 * threadIdx.x;
    while (ld_acquire_sys(f) < epoch) {
    }
  }
  __syncthreads();
  for (int64_t j = tid; j < len; j += stride) {
    double v[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < x.world) v[p] = ld_relaxed_sys(x.peer_slot[p][par] + j);
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < x.world) s += v[p];
    data[j] = s;
  }
}

class PeerExchange {
 public:
  bool attached() const { return attached_ && world_ > 1; }
  const std::string& error() const { return err_; }
  int launches_per_reduce() const { return 1; }

  bool export_handle(int64_t len, PeerHandle* out) {
    if (!base_) {
      len_ = len;
      bytes_ = 1024 + 2 * (size_t)len * sizeof(double);
      if (!ok(cudaMalloc(&base_, bytes_), "cudaMalloc(exchange)")) return false;
      if (!ok(cudaMemset(base_, 0, bytes_), "cudaMemset(exchange)")) return false;
      if (!ok(cudaMalloc(&arrive_, sizeof(unsigned int)), "cudaMalloc(arrive)")) return false;
      if (!ok(cudaMemset(arrive_, 0, sizeof(unsigned int)), "cudaMemset(arrive)")) return false;
    }
    memset(out, 0, sizeof(*out));
    if (!ok(cudaIpcGetMemHandle(&out->ipc, base_), "cudaIpcGetMemHandle")) return false;
    out->raw_ptr = (uint64_t)(uintptr_t)base_;
    out->len = len_;
    out->pid = (int32_t)getpid();
    int dev = 0;
    cudaGetDevice(&dev);
    out->device = dev;
    return true;
  }

  bool attach(int world, int rank, const unsigned char* handles, size_t stride,
              int sm_count) {
    if (!base_) {
      err_ = "cfmm_comm_export must be called before cfmm_comm_attach";
      return false;
    }
    world_ = world;
    rank_ = rank;
    grid_ = sm_count < 64 ? sm_count : 64;
    int my_dev = 0;
    cudaGetDevice(&my_dev);
    memset(&view_, 0, sizeof(view_));
    view_.world = world;
    view_.rank = rank;
    view_.arrive = arrive_;
    view_.flags = (unsigned long long*)base_;
    view_.my_slot[0] = (double*)((char*)base_ + 1024);
    view_.my_slot[1] = view_.my_slot[0] + len_;
    for (int p = 0; p < world; ++p) {
      PeerHandle h;
      memcpy(&h, handles + (size_t)p * stride, sizeof(h));
      if (h.len != len_) {
        err_ = "rank " + std::to_string(p) + " exported a different vector length";
        return false;
      }
      void* pbase = nullptr;
      if (p == rank) {
        pbase = base_;
      } else if (h.pid == (int32_t)getpid()) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, my_dev, h.device);
        if (!can) {
          err_ = "no peer access between device " + std::to_string(my_dev) +
                 " and " + std::to_string(h.device);
          return false;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return ok(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
        pbase = (void*)(uintptr_t)h.raw_ptr;
      } else {
        if (!ok(cudaIpcOpenMemHandle(&pbase, h.ipc, cudaIpcMemLazyEnablePeerAccess),
                "cudaIpcOpenMemHandle"))
          return false;
        opened_[p] = pbase;
      }
      view_.peer_flags[p] = (unsigned long long*)pbase;
      view_.peer_slot[p][0] = (const double*)((char*)pbase + 1024);
      view_.peer_slot[p][1] = view_.peer_slot[p][0] + len_;
    }
    epoch_ = 0;
    attached_ = true;
    return true;
  }

  bool all_reduce(double* data, int64_t len, cudaStream_t st) {
    if (len != len_) {
      err_ = "all_reduce length mismatch";
      return false;
    }
    ++epoch_;
    peer_allreduce_kernel<<<grid_, kExchangeThreads, 0, st>>>(view_, data, len, epoch_);
    return ok(cudaGetLastError(), "peer_allreduce_kernel launch");
  }

  void detach() {
    for (int p = 0; p < kMaxPeers; ++p)
      if (opened_[p]) {
        cudaIpcCloseMemHandle(opened_[p]);
        opened_[p] = nullptr;
      }
    if (base_) cudaFree(base_);
    if (arrive_) cudaFree(arrive_);
    base_ = nullptr;
    arrive_ = nullptr;
    attached_ = false;
    world_ = 1;
  }

 private:
  bool ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err_ = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
  void* base_ = nullptr;
  unsigned int* arrive_ = nullptr;
  void* opened_[kMaxPeers] = {};
  size_t bytes_ = 0;
  int64_t len_ = 0;
  int world_ = 1, rank_ = 0, grid_ = 64;
  unsigned long long epoch_ = 0;
  bool attached_ = false;
  ExchangeView view_;
  std::string err_;
};

}  // namespace cfmm
