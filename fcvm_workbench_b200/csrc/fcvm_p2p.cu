// Peer-memory exchanges of the multi-GPU PCG iteration over NVLink / NVSwitch.
//
// One process per GPU; every rank owns an "arena" in device memory that its peers map through CUDA IPC.  An
// exchange is ONE kernel per rank (a single block for the small sums, a few blocks for the halo -- all resident
// while they wait) that (1) stores its contribution straight into the peers' arenas
// (st.global on mapped peer pointers), (2) publishes an epoch flag with release semantics at system scope,
// (3) spins on its own flags until every contributor's epoch has arrived (acquire), and (4) combines the
// contributions in ascending rank order -- so every rank computes bit-identical sums and takes the same
// decisions.  Two exchanges per PCG iteration replace four NCCL all-reduces (and their launch latency, which
// dominated the iteration at 8 GPUs):
//
//   k_xchg_sum      coarse solution: every rank multiplies the column panel of E^-1 its own right-hand side lives
//                   in, the products (6 x boxes doubles per rank) are summed
//   k_xchg_halo     rows of w = K u at shared nodes -- neighbours only (Partition.p2p_plan), not a dense global
//                   interface vector -- together with the three per-rank scalars of the iteration
//
// Buffers are double-buffered by the parity of an epoch that every rank advances with every launch, so a peer
// can never overwrite data that is still being read: to be two exchanges ahead it would need this rank's flag
// of the exchange in between.  (The no-op launches after convergence are skipped by all ranks together -- the
// flag they test is itself a bit-identical sum -- and the next real exchange follows a host synchronisation.)
// A rank that waits longer than ~6 s raises a status flag instead of hanging.  NCCL stays for everything outside
// the iteration.
#include <algorithm>

#include "fcvm_common.cuh"
#include "fcvm_pcg.cuh"

using namespace fcvm;

namespace fcvm {

constexpr int P2P_MAX_RANKS = 8;

struct P2PDev {
  int world, rank, npeers, n_if;
  char *peer[P2P_MAX_RANKS];        // base of every rank's arena (peer[rank] = own)
  unsigned long long off_flags, off_ticket, off_status, off_scal, off_slots, off_halo;
  long long slot_n, halo_cap;       // doubles per slot, nodes per halo buffer
  const int32_t *peer_rank, *send_ptr, *send_node;
  const int64_t *remote_off;
  const int32_t *if_node, *if_ptr;
  const int64_t *if_src;
};

struct P2PState {
  P2PDev d;
  char *arena = nullptr;
  size_t bytes = 0;
  bool opened[P2P_MAX_RANKS] = {false};
  int32_t *peer_rank = nullptr, *send_ptr = nullptr, *send_node = nullptr, *if_node = nullptr, *if_ptr = nullptr;
  int64_t *remote_off = nullptr, *if_src = nullptr;
  unsigned long long *h_status = nullptr;   // pinned
  int64_t n_send = 0;
  unsigned long long epoch[2] = {0, 0};     // per channel, advanced by the host with every launch (all ranks alike)
};

}  // namespace fcvm

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr long long SPIN_LIMIT = 12000000000LL;    // SM cycles (~6 s)
constexpr int XT = 1024;

// publish epoch e on channel ch to every rank (the block whose `publish` is set), then wait for everyone's
__device__ __forceinline__ bool signal_and_wait(const P2PDev &d, int ch, unsigned long long e, bool publish, int *bad) {
  if ((int)threadIdx.x < d.world) {
    const int q = (int)threadIdx.x;
    if (publish) {
      unsigned long long *theirs = (unsigned long long *)(d.peer[q] + d.off_flags) + ch * P2P_MAX_RANKS + d.rank;
      st_release_sys(theirs, e);
    }
    const unsigned long long *mine = (const unsigned long long *)(d.peer[d.rank] + d.off_flags) + ch * P2P_MAX_RANKS + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < e) {
      if (clock64() - t0 > SPIN_LIMIT) {
        *bad = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (*bad) {
    if (threadIdx.x == 0) *(unsigned long long *)(d.peer[d.rank] + d.off_status) = 1ull;
    return false;
  }
  return true;
}

// out[i] = sum over ranks of their src[i], i < n (ascending rank)
__global__ void __launch_bounds__(XT)
k_xchg_sum(P2PDev d, unsigned long long e, const double *src, long long n, double *out, const double *sc, int done_slot) {
  if (sc && sc[done_slot] >= 0.0) return;            // every rank holds the same flag
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  const long long par = (long long)(e & 1ull);
  for (int p = 0; p < d.world; p++) {
    double *dst = (double *)(d.peer[p] + d.off_slots) + (par * d.world + d.rank) * d.slot_n;
    for (long long i = threadIdx.x; i < n; i += XT) dst[i] = src[i];
  }
  // release pattern: the block's stores happen before the barrier, the flag writers fence at system scope after it
  // (fences are cumulative) -- one fence per flag writer instead of one per thread
  __syncthreads();
  if ((int)threadIdx.x < d.world) __threadfence_system();
  if (!signal_and_wait(d, 0, e, true, &bad)) return;
  const double *mine = (const double *)(d.peer[d.rank] + d.off_slots) + par * d.world * d.slot_n;
  for (long long i = threadIdx.x; i < n; i += XT) {
    double s = 0.0;
    for (int r = 0; r < d.world; r++) s += __ldcg(mine + r * d.slot_n + i);
    out[i] = s;
  }
}

// v[shared nodes] = sum over the ranks that hold them (ascending rank); the three per-rank scalars of the
// iteration (r.u, r.r, w.u in sc[L_RU], sc[L_RR], sc[L_WU]) are summed on the way and stored like k_tail_get did.
// Several blocks: all push, the block that finishes its pushes last (ticket) publishes the flags, every block
// waits for the peers' flags itself and unpacks its share -- no grid barrier.
constexpr int HT = 256;
__global__ void __launch_bounds__(HT)
k_xchg_halo(P2PDev d, unsigned long long e, double *v, double *sc, int gamma_slot, int rr_slot, int with_scalars, int done_check) {
  if (done_check && sc[S_ITERS] >= 0.0) return;
  __shared__ int bad, publish;
  if (threadIdx.x == 0) bad = 0;
  const long long par = (long long)(e & 1ull);
  const long long stride = (long long)gridDim.x * HT, t0 = blockIdx.x * (long long)HT + threadIdx.x;
  for (int pi = 0; pi < d.npeers; pi++) {
    const int q = d.peer_rank[pi];
    double *dst = (double *)(d.peer[q] + d.off_halo) + 3 * (par * d.halo_cap + d.remote_off[pi]);
    const int32_t k0 = d.send_ptr[pi], k1 = d.send_ptr[pi + 1];
    for (long long i = t0; i < 3LL * (k1 - k0); i += stride) {
      const long long k = i / 3;
      const int c = (int)(i - 3 * k);
      dst[i] = v[3 * (long long)d.send_node[k0 + k] + c];
    }
  }
  if (with_scalars && blockIdx.x == 0 && threadIdx.x < 3) {
    const double val = sc[threadIdx.x == 0 ? L_RU : (threadIdx.x == 1 ? L_RR : L_WU)];
    for (int p = 0; p < d.world; p++) ((double *)(d.peer[p] + d.off_scal))[(par * P2P_MAX_RANKS + d.rank) * 4 + threadIdx.x] = val;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();        // this block's pushes (ordered before the barrier) before its ticket
    unsigned int *ticket = (unsigned int *)(d.peer[d.rank] + d.off_ticket);
    publish = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    if (publish) *ticket = 0u;
  }
  __syncthreads();
  if (publish && (int)threadIdx.x < d.world) __threadfence_system();   // the other blocks' pushes (seen through the ticket) before the flags
  // the scalars go to every rank, the rows to the neighbours only: one flag round over all ranks covers both
  if (!signal_and_wait(d, 1, e, publish != 0, &bad)) return;
  const double *halo = (const double *)(d.peer[d.rank] + d.off_halo) + 3 * par * d.halo_cap;
  for (long long i = t0; i < 3LL * d.n_if; i += stride) {
    const long long k = i / 3;
    const int c = (int)(i - 3 * k);
    const long long node = d.if_node[k];
    const double own = v[3 * node + c];
    double s = 0.0;
    for (int32_t j = d.if_ptr[k]; j < d.if_ptr[k + 1]; j++) {
      const long long src = d.if_src[j];
      s += src < 0 ? own : __ldcg(halo + 3 * src + c);
    }
    v[3 * node + c] = s;
  }
  if (with_scalars && blockIdx.x == 0 && threadIdx.x < 3) {
    const double *scal = (const double *)(d.peer[d.rank] + d.off_scal) + par * P2P_MAX_RANKS * 4;
    double s = 0.0;
    for (int r = 0; r < d.world; r++) s += __ldcg(scal + r * 4 + threadIdx.x);
    if (threadIdx.x == 0 && gamma_slot >= 0) sc[gamma_slot] = s;
    if (threadIdx.x == 1 && rr_slot >= 0) sc[rr_slot] = s;
    if (threadIdx.x == 2) sc[S_DELTA] = s;
  }
}

template <typename T>
int upload(T **dst, const T *src, int64_t n) {
  *dst = nullptr;
  FCVM_CUDA(cudaMalloc((void **)dst, sizeof(T) * (size_t)std::max<int64_t>(n, 1)));
  if (n > 0) FCVM_CUDA(cudaMemcpy(*dst, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
  return FCVM_OK;
}

}  // namespace

namespace fcvm {

bool p2p_ready(const fcvm_ctx *c) { return c->p2p != nullptr && c->p2p_attached; }

int p2p_allreduce_sum(fcvm_ctx *c, double *v, int64_t n, const double *sc, int done_slot) {
  P2PState *s = c->p2p;
  FCVM_CHECK(n <= s->d.slot_n, FCVM_E_ARG, "p2p exchange: %lld doubles exceed the slot size %lld", (long long)n, (long long)s->d.slot_n);
  ProfScope ps(c, 7);
  k_xchg_sum<<<1, XT, 0, c->stream>>>(s->d, ++s->epoch[0], v, n, v, sc, done_slot);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

int p2p_halo(fcvm_ctx *c, double *v, double *sc, int gamma_slot, int rr_slot, bool with_scalars, bool done_check) {
  P2PState *s = c->p2p;
  // every block must be resident while it waits for the peers: far fewer blocks than SMs
  const int64_t work = 3 * std::max<int64_t>(s->n_send, s->d.n_if);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(48, (work + 4 * HT - 1) / (4 * HT)));
  ProfScope ps(c, 7);
  k_xchg_halo<<<grid, HT, 0, c->stream>>>(s->d, ++s->epoch[1], v, sc, gamma_slot, rr_slot, with_scalars ? 1 : 0, done_check ? 1 : 0);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

// after a batch has been waited for: did an exchange give up?
int p2p_check(fcvm_ctx *c) {
  P2PState *s = c->p2p;
  FCVM_CUDA(cudaMemcpyAsync(s->h_status, s->arena + s->d.off_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  FCVM_CHECK(*s->h_status == 0ull, FCVM_E_NCCL, "peer-memory exchange timed out: a rank did not reach the same exchange (rank %d)", c->rank);
  return FCVM_OK;
}

void p2p_free(fcvm_ctx *c) {
  P2PState *s = c->p2p;
  if (!s) return;
  for (int r = 0; r < P2P_MAX_RANKS; r++)
    if (s->opened[r]) cudaIpcCloseMemHandle(s->d.peer[r]);
  cudaFree(s->arena);
  cudaFree(s->peer_rank); cudaFree(s->send_ptr); cudaFree(s->send_node); cudaFree(s->remote_off);
  cudaFree(s->if_node); cudaFree(s->if_ptr); cudaFree(s->if_src);
  if (s->h_status) cudaFreeHost(s->h_status);
  delete s;
  c->p2p = nullptr;
  c->p2p_attached = false;
}

}  // namespace fcvm

// Allocates this rank's arena (receive area for n_recv_nodes shared-node rows, world slots of slot_n doubles) and
// returns its 64-byte CUDA IPC handle for the peers.
extern "C" int fcvm_p2p_create(fcvm_ctx *c, int64_t n_recv_nodes, int64_t slot_n, void *handle64) {
  FCVM_CHECK(c && handle64 && n_recv_nodes >= 0 && slot_n > 0, FCVM_E_ARG, "fcvm_p2p_create: bad argument");
  FCVM_CHECK(c->world > 1 && c->world <= P2P_MAX_RANKS, FCVM_E_ARG, "fcvm_p2p_create: needs 2..%d ranks (fcvm_comm_init first)", P2P_MAX_RANKS);
  FCVM_CUDA(cudaSetDevice(c->device));
  p2p_free(c);
  P2PState *s = new P2PState();
  c->p2p = s;
  P2PDev &d = s->d;
  memset(&d, 0, sizeof(d));
  d.world = c->world;
  d.rank = c->rank;
  d.slot_n = (slot_n + 15) / 16 * 16;
  d.halo_cap = std::max<int64_t>(n_recv_nodes, 1);
  d.off_flags = 0;
  d.off_ticket = 2 * P2P_MAX_RANKS * 8;
  d.off_status = d.off_ticket + 16;
  d.off_scal = 256;
  d.off_slots = 1024;
  d.off_halo = d.off_slots + sizeof(double) * 2 * (size_t)d.world * (size_t)d.slot_n;
  s->bytes = d.off_halo + sizeof(double) * 2 * 3 * (size_t)d.halo_cap;
  FCVM_CUDA(cudaMalloc((void **)&s->arena, s->bytes));
  FCVM_CUDA(cudaMemset(s->arena, 0, s->bytes));
  FCVM_CUDA(cudaMallocHost((void **)&s->h_status, sizeof(unsigned long long)));
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  FCVM_CUDA(cudaIpcGetMemHandle(&h, s->arena));
  memcpy(handle64, &h, 64);
  d.peer[c->rank] = s->arena;
  return FCVM_OK;
}

// Maps the peers' arenas (handles: world x 64 bytes, rank order) and uploads the exchange lists of
// Partition.p2p_plan.  All ranks must have returned from this call before the first solve (the caller barriers).
extern "C" int fcvm_p2p_attach(fcvm_ctx *c, const void *handles, int npeers, const int32_t *peer_rank, const int32_t *send_ptr,
                               const int32_t *send_node, const int64_t *remote_off, int n_if, const int32_t *if_node,
                               const int32_t *if_ptr, const int64_t *if_src) {
  FCVM_CHECK(c && c->p2p && handles, FCVM_E_ARG, "fcvm_p2p_attach: call fcvm_p2p_create first");
  P2PState *s = c->p2p;
  P2PDev &d = s->d;
  FCVM_CUDA(cudaSetDevice(c->device));
  for (int r = 0; r < d.world; r++) {
    if (r == d.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * r, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("fcvm_p2p_attach: cannot map the arena of rank %d (%s)", r, cudaGetErrorString(e));
      cudaGetLastError();
      return FCVM_E_NCCL;
    }
    d.peer[r] = (char *)p;
    s->opened[r] = true;
  }
  d.npeers = npeers;
  d.n_if = n_if;
  s->n_send = send_ptr ? send_ptr[npeers] : 0;
  FCVM_TRY(upload(&s->peer_rank, peer_rank, npeers));
  FCVM_TRY(upload(&s->send_ptr, send_ptr, npeers + 1));
  FCVM_TRY(upload(&s->send_node, send_node, send_ptr ? send_ptr[npeers] : 0));
  FCVM_TRY(upload(&s->remote_off, remote_off, npeers));
  FCVM_TRY(upload(&s->if_node, if_node, n_if));
  FCVM_TRY(upload(&s->if_ptr, if_ptr, n_if + 1));
  FCVM_TRY(upload(&s->if_src, if_src, if_ptr ? if_ptr[n_if] : 0));
  d.peer_rank = s->peer_rank; d.send_ptr = s->send_ptr; d.send_node = s->send_node; d.remote_off = s->remote_off;
  d.if_node = s->if_node; d.if_ptr = s->if_ptr; d.if_src = s->if_src;
  c->p2p_attached = true;
  return FCVM_OK;
}

// v[shared nodes] = sum over ranks through the peer-memory halo (test entry point; fcvm_interface_sum is the NCCL one)
extern "C" int fcvm_p2p_interface_sum(fcvm_ctx *c, double *v) {
  FCVM_CHECK(c && v && p2p_ready(c), FCVM_E_ARG, "fcvm_p2p_interface_sum: peer exchange not attached");
  FCVM_TRY(p2p_halo(c, v, c->red_out, -1, -1, false, false));
  return p2p_check(c);
}
