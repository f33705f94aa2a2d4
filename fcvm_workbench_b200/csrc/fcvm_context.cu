// Context, mesh upload, sparsity pattern (fixed once per mesh) and nodal-vector kernels.
#include <stdarg.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <cusolverDn.h>
#include <numeric>

#include "fcvm_common.cuh"
#include "fcvm_reduce.cuh"

namespace fcvm {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

template <typename T>
static int dalloc(T **p, int64_t n) {
  *p = nullptr;
  if (n <= 0) n = 1;
  FCVM_CUDA(cudaMalloc((void **)p, sizeof(T) * (size_t)n));
  return FCVM_OK;
}

template <typename T>
static void dfree(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

}  // namespace fcvm

using namespace fcvm;

extern "C" const char *fcvm_last_error(void) { return fcvm::g_err; }
extern "C" int fcvm_version(void) { return 100; }

// ------------------------------------------------------------------------------------------
// kernels: pattern construction
// ------------------------------------------------------------------------------------------
__global__ void k_n2e_keys(int64_t ne, const int32_t *__restrict__ conn, int32_t *keys, int32_t *vals,
                           int32_t *count) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 10 * ne) return;
  int64_t e = i / 10;
  int j = (int)(i - 10 * e);
  int32_t nd = conn[(int64_t)j * ne + e];
  keys[i] = nd;
  vals[i] = (int32_t)i;
  atomicAdd(&count[nd], 1);
}

// one key per (element, local row node a, local column node b)
__global__ void k_pair_keys(int64_t ne, const int32_t *__restrict__ conn, uint64_t *keys, uint32_t *vals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 100 * ne) return;
  int64_t e = i / 100;
  int ab = (int)(i - 100 * e);
  int a = ab / 10, b = ab - 10 * a;
  uint32_t ra = (uint32_t)conn[(int64_t)a * ne + e], cb = (uint32_t)conn[(int64_t)b * ne + e];
  keys[i] = ((uint64_t)ra << 32) | cb;
  int hi = a >= b ? a : b, lo = a >= b ? b : a;
  uint32_t pair = (uint32_t)(hi * (hi + 1) / 2 + lo);
  uint32_t T = a >= b ? 0u : 1u;   // stored block is K[hi][lo]; K[a][b] with a < b is its transpose
  vals[i] = (uint32_t)(((uint64_t)pair * (uint64_t)ne + (uint64_t)e) << 1) | T;
}

__global__ void k_heads(int64_t n, const uint64_t *__restrict__ keys, int32_t *head) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// head_scan = exclusive prefix sum of head: block id of contribution i is head_scan[i] + head[i] - 1
__global__ void k_block_table(int64_t n, const uint64_t *__restrict__ keys, const int32_t *__restrict__ head,
                              const int32_t *__restrict__ head_scan, uint64_t *blk_key, uint32_t *blk_first,
                              int32_t *rowlen) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) {
    int32_t b = head_scan[i];
    blk_key[b] = keys[i];
    blk_first[b] = (uint32_t)i;
    atomicAdd(&rowlen[(int32_t)(keys[i] >> 32)], 1);
  }
}

__global__ void k_sell_init(int64_t nslices, const int32_t *__restrict__ slice_ptr,
                            const int32_t *__restrict__ slot_node, int32_t *colidx, uint32_t *blk_cnt,
                            uint32_t *blk_first) {
  int64_t s = blockIdx.x;
  int lane = threadIdx.x;
  if (s >= nslices) return;
  int32_t nd = slot_node[s * SELL_C + lane];
  int32_t col = nd >= 0 ? nd : 0;
  for (int32_t k = slice_ptr[s]; k < slice_ptr[s + 1]; k++) {
    int64_t pos = (int64_t)k * SELL_C + lane;
    colidx[pos] = col;
    blk_cnt[pos] = 0u;
    blk_first[pos] = 0u;
  }
}

__global__ void k_sell_fill(int64_t nblk, int64_t ncontrib, const uint64_t *__restrict__ blk_key,
                            const uint32_t *__restrict__ first_real, const int32_t *__restrict__ row_first,
                            const int32_t *__restrict__ node_slot, const int32_t *__restrict__ slice_ptr,
                            int32_t *colidx, uint32_t *blk_first, uint32_t *blk_cnt, int32_t *diag_pos,
                            int32_t *row_cols) {
  int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nblk) return;
  uint64_t key = blk_key[b];
  int32_t row = (int32_t)(key >> 32), col = (int32_t)(key & 0xffffffffu);
  int32_t k = (int32_t)(b - row_first[row]);
  int32_t slot = node_slot[row];
  int64_t pos = ((int64_t)slice_ptr[slot / SELL_C] + k) * SELL_C + (slot % SELL_C);
  uint32_t f = first_real[b];
  uint32_t nxt = (b + 1 < nblk) ? first_real[b + 1] : (uint32_t)ncontrib;
  colidx[pos] = col;
  blk_first[pos] = f;
  blk_cnt[pos] = nxt - f;
  row_cols[b] = col;
  if (row == col) diag_pos[row] = (int32_t)pos;
}

// ------------------------------------------------------------------------------------------
// kernels: nodal vectors
// ------------------------------------------------------------------------------------------
__global__ void k_axpby(int64_t n, double a, const double *x, double b, double *y) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = (b == 0.0) ? a * x[i] : a * x[i] + b * y[i];
}

__global__ void k_axpbypcz(int64_t n, double a, const double *x, double b, const double *y, double c, double *z) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = (c == 0.0) ? a * x[i] + b * y[i] : a * x[i] + b * y[i] + c * z[i];
}

__global__ void __launch_bounds__(RED_THREADS) k_dot(int64_t n, const double *__restrict__ x,
                                                       const double *__restrict__ y,
                                                       const double *__restrict__ w, double *red_part,
                                                       unsigned int *counter, double *out) {
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[0] += (w ? w[i] : 1.0) * x[i] * y[i];
  block_reduce_publish<1>(v, red_part, counter, out);
}

// r = fixdof * (lbd * glv - qin), partial sums of r^2   (fcVM.py:1329-1338 / 1446-1447)
__global__ void __launch_bounds__(RED_THREADS) k_residual(int64_t n, double lbd, const double *__restrict__ glv,
                                                            const double *__restrict__ qin,
                                                            const double *__restrict__ fixdof,
                                                            const double *__restrict__ w, double *r, double *red_part,
                                                            unsigned int *counter, double *out) {
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double ri = fixdof[i] * (lbd * glv[i] - qin[i]);
    r[i] = ri;
    v[0] += (w ? w[i] : 1.0) * ri * ri;
  }
  block_reduce_publish<1>(v, red_part, counter, out);
}

// max over nodes of |u|^2 : max is order-independent, so a plain two-stage max is deterministic
__global__ void __launch_bounds__(RED_THREADS) k_max_node_disp(int64_t nnodes, const double *__restrict__ u,
                                                                 double *red_part, unsigned int *counter,
                                                                 double *out) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnodes;
       i += (int64_t)gridDim.x * blockDim.x) {
    double a = u[3 * i], b = u[3 * i + 1], c = u[3 * i + 2];
    m = fmax(m, a * a + b * b + c * c);
  }
  __shared__ double sm[RED_THREADS / 32];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RED_THREADS / 32; w++) m = fmax(m, sm[w]);
    red_part[blockIdx.x] = m;
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    double s = 0.0;
    for (int b = threadIdx.x; b < RED_BLOCKS; b += 32) s = fmax(s, __ldcg(&red_part[b]));
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_down_sync(0xffffffffu, s, o));
    if (threadIdx.x == 0) {
      out[0] = s;
      *counter = 0u;
    }
  }
}

// Gauss-point layout: device SoA [(c*4+ip)*ne + el]  <->  reference AoS [(4*el+ip)*ncomp + c]
__global__ void k_gp_soa_to_aos(int64_t ne, int ncomp, const double *__restrict__ soa, double *aos) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // index into aos
  if (i >= ne * 4 * ncomp) return;
  int c = (int)(i % ncomp);
  int64_t g = i / ncomp;
  int ip = (int)(g & 3);
  int64_t el = g >> 2;
  aos[i] = soa[((int64_t)c * 4 + ip) * ne + el];
}

__global__ void k_gp_aos_to_soa(int64_t ne, int ncomp, const double *__restrict__ aos, double *soa) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // index into soa
  if (i >= ne * 4 * ncomp) return;
  int64_t el = i % ne;
  int64_t q = i / ne;
  int ip = (int)(q & 3);
  int c = (int)(q >> 2);
  soa[i] = aos[(4 * el + ip) * ncomp + c];
}

__global__ void k_pgp_soa_to_aos(int64_t ne, const uint8_t *__restrict__ soa, uint8_t *aos) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 4 * ne) return;
  aos[i] = soa[(i & 3) * ne + (i >> 2)];
}

// the same conversions for the elements [a, b) only (chunks of the pipelined host-buffer path); the AoS side is a
// full-size staging array, so indices are absolute
__global__ void k_gp_aos_to_soa_range(int64_t ne, int64_t a, int64_t b, int ncomp, const double *__restrict__ aos, double *soa) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, len = b - a;
  if (t >= len * 4 * ncomp) return;
  const int64_t el = a + t % len, q = t / len;
  const int ip = (int)(q & 3), c = (int)(q >> 2);
  soa[q * ne + el] = aos[(4 * el + ip) * ncomp + c];
}
__global__ void k_gp_soa_to_aos_range(int64_t ne, int64_t a, int64_t b, int ncomp, const double *__restrict__ soa, double *aos) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (b - a) * 4 * ncomp) return;
  const int64_t i = 4 * a * ncomp + t;                 // absolute AoS index
  const int c = (int)(i % ncomp);
  const int64_t g = i / ncomp;
  aos[i] = soa[((int64_t)c * 4 + (g & 3)) * ne + (g >> 2)];
}
__global__ void k_pgp_soa_to_aos_range(int64_t ne, int64_t a, int64_t b, const uint8_t *__restrict__ soa, uint8_t *aos) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 4 * (b - a)) return;
  const int64_t i = 4 * a + t;
  aos[i] = soa[(i & 3) * ne + (i >> 2)];
}

namespace fcvm {
int launch_gp_in(fcvm_ctx *c, int64_t a, int64_t b, int ncomp, const double *aos, double *soa) {
  k_gp_aos_to_soa_range<<<grid_for((b - a) * 4 * ncomp, 256), 256, 0, c->stream>>>(c->ne, a, b, ncomp, aos, soa);
  c->launches++;
  return FCVM_OK;
}
int launch_gp_out(fcvm_ctx *c, int64_t a, int64_t b, int ncomp, const double *soa, double *aos) {
  k_gp_soa_to_aos_range<<<grid_for((b - a) * 4 * ncomp, 256), 256, 0, c->stream>>>(c->ne, a, b, ncomp, soa, aos);
  c->launches++;
  return FCVM_OK;
}
int launch_pgp_out(fcvm_ctx *c, int64_t a, int64_t b, const uint8_t *soa, uint8_t *aos) {
  k_pgp_soa_to_aos_range<<<grid_for(4 * (b - a), 256), 256, 0, c->stream>>>(c->ne, a, b, soa, aos);
  c->launches++;
  return FCVM_OK;
}
}  // namespace fcvm

__global__ void k_fill(int64_t n, double v, double *x) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

__global__ void __launch_bounds__(RED_THREADS) k_count_u8(int64_t n, const uint8_t *__restrict__ x, double *red_part,
                                                            unsigned int *counter, double *out) {
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[0] += x[i] ? 1.0 : 0.0;
  block_reduce_publish<1>(v, red_part, counter, out);
}

__global__ void k_if_pack(int64_t n_if, const int32_t *__restrict__ node, const int32_t *__restrict__ slot,
                          const double *__restrict__ v, double *buf) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * n_if) return;
  int64_t k = i / 3;
  int c = (int)(i - 3 * k);
  buf[3 * (int64_t)slot[k] + c] = v[3 * (int64_t)node[k] + c];
}

__global__ void k_if_unpack(int64_t n_if, const int32_t *__restrict__ node, const int32_t *__restrict__ slot,
                            const double *__restrict__ buf, double *v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * n_if) return;
  int64_t k = i / 3;
  int c = (int)(i - 3 * k);
  v[3 * (int64_t)node[k] + c] = buf[3 * (int64_t)slot[k] + c];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int read_scalars(fcvm_ctx *c, int n) {
  FCVM_CUDA(cudaMemcpyAsync(c->h_scalars, c->red_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  return FCVM_OK;
}

extern "C" int fcvm_comm_allreduce_sum(fcvm_ctx *c, double *dev, int64_t n);
extern "C" int fcvm_comm_allreduce_max(fcvm_ctx *c, double *dev, int64_t n);

static int finish_scalar(fcvm_ctx *c, int n, bool sum_over_ranks) {
  if (sum_over_ranks && c->world > 1) FCVM_TRY(fcvm_comm_allreduce_sum(c, c->red_out, n));
  return read_scalars(c, n);
}

extern "C" int fcvm_create(fcvm_ctx **out, int device) {
  FCVM_CHECK(out != nullptr, FCVM_E_ARG, "fcvm_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("fcvm_create: no CUDA device (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return FCVM_E_CUDA;
  }
  FCVM_CHECK(device >= 0 && device < ndev, FCVM_E_ARG, "fcvm_create: device %d out of range (0..%d)", device,
             ndev - 1);
  FCVM_CUDA(cudaSetDevice(device));
  fcvm_ctx *c = new fcvm_ctx();
  c->device = device;
  FCVM_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  FCVM_CUDA(cudaEventCreate(&c->ev0));
  FCVM_CUDA(cudaEventCreate(&c->ev1));
  FCVM_CUDA(cudaEventCreate(&c->pev0));
  FCVM_CUDA(cudaEventCreate(&c->pev1));
  FCVM_TRY(dalloc(&c->red_part, 8 * RED_BLOCKS));
  FCVM_TRY(dalloc(&c->red_out, 16));
  FCVM_TRY(dalloc(&c->red_counter, 4));
  FCVM_TRY(dalloc(&c->d_arg, 4));
  FCVM_CUDA(cudaMemset(c->red_counter, 0, sizeof(unsigned int) * 4));
  FCVM_CUDA(cudaMemset(c->red_out, 0, sizeof(double) * 16));
  FCVM_CUDA(cudaMallocHost((void **)&c->h_scalars, sizeof(double) * 16));
  FCVM_CUDA(cudaMallocHost((void **)&c->h_arg, sizeof(int64_t) * 4));
  *out = c;
  return FCVM_OK;
}

namespace fcvm {
void p2p_free(fcvm_ctx *c);
int matfree_set_constraints(fcvm_ctx *c);
int matfree_set_mesh(fcvm_ctx *c);
void deflation_free(fcvm_ctx *c);
}

static void free_mesh(fcvm_ctx *c) {
  deflation_free(c);
  dfree(c->conn); dfree(c->xyz); dfree(c->n2e_ptr); dfree(c->n2e_idx); dfree(c->elv);
  dfree(c->fixmask); dfree(c->fixval); dfree(c->movmask);
  for (int i = 0; i < FCVM_BUF_COUNT; i++) {
    if (c->buf[i]) cudaFree(c->buf[i]);
    c->buf[i] = nullptr;
  }
  dfree(c->slice_ptr); dfree(c->slot_node); dfree(c->node_slot); dfree(c->colidx); dfree(c->vals); dfree(c->vals2);
  dfree(c->blk_first); dfree(c->blk_cnt); dfree(c->src); dfree(c->diag_pos); dfree(c->row_first);
  dfree(c->row_cols); dfree(c->cooK); dfree(c->minv);
  dfree(c->pcg_r); dfree(c->pcg_z); dfree(c->pcg_p); dfree(c->pcg_q); dfree(c->pcg_s); dfree(c->spmv_part);
  dfree(c->dof_weight); dfree(c->if_node); dfree(c->if_slot); dfree(c->if_buf);
  dfree(c->bslices); dfree(c->islices); dfree(c->tail3);
  for (int i = 0; i < 2; i++) { dfree(c->hist_b[i]); dfree(c->hist_x[i]); }
  c->hist_n = 0;
  dfree(c->emask); dfree(c->ga_ticket); dfree(c->ga_group_part); dfree(c->egeo); dfree(c->tile_affine);
  c->n_affine_tiles = 0;
  dfree(c->hs_in); dfree(c->hs_out); dfree(c->hs_pgp);
  dfree(c->h_du); dfree(c->h_disp); dfree(c->h_qin); dfree(c->diag9); dfree(c->gp_tmp);
  c->assembled = false;
  c->have_bcs = false;
  c->un_nodes = -1;
}

extern "C" int fcvm_comm_destroy_(fcvm_ctx *c);

extern "C" int fcvm_destroy(fcvm_ctx *c) {
  if (!c) return FCVM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  p2p_free(c);
  fcvm_comm_destroy_(c);
  free_mesh(c);
  dfree(c->red_part); dfree(c->red_out); dfree(c->red_counter); dfree(c->d_arg); dfree(c->d_arg_part);
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  if (c->h_arg) cudaFreeHost(c->h_arg);
  if (c->cusolver) cusolverDnDestroy((cusolverDnHandle_t)c->cusolver);
  if (c->cus_work) cudaFree(c->cus_work);
  if (c->cus_info) cudaFree(c->cus_info);
  if (c->h_in_stream) cudaStreamDestroy(c->h_in_stream);
  if (c->h_out_stream) cudaStreamDestroy(c->h_out_stream);
  for (int i = 0; i < 16; i++) {
    if (c->h_ev_in[i]) cudaEventDestroy(c->h_ev_in[i]);
    if (c->h_ev_k[i]) cudaEventDestroy(c->h_ev_k[i]);
  }
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->ev_boundary) cudaEventDestroy(c->ev_boundary);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->pev0); cudaEventDestroy(c->pev1);
  for (auto &s : c->prof_pool) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
  cudaStreamDestroy(c->own_stream);
  delete c;
  return FCVM_OK;
}

extern "C" int fcvm_set_stream(fcvm_ctx *c, void *s) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return FCVM_OK;
}

extern "C" int fcvm_synchronize(fcvm_ctx *c) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  return FCVM_OK;
}

extern "C" int64_t fcvm_num_elements(const fcvm_ctx *c) { return c ? c->ne : 0; }
extern "C" int64_t fcvm_num_nodes(const fcvm_ctx *c) { return c ? c->nn : 0; }
extern "C" int64_t fcvm_launch_count(fcvm_ctx *c) { return c ? c->launches : 0; }
extern "C" int fcvm_copy_bytes(fcvm_ctx *c, int64_t *h2d, int64_t *d2h) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  if (h2d) *h2d = c->h2d_bytes;
  if (d2h) *d2h = c->d2h_bytes;
  return FCVM_OK;
}

static int bits_for(uint64_t n) {
  int b = 1;
  while ((1ull << b) < n) b++;
  return b;
}

extern "C" int fcvm_set_mesh(fcvm_ctx *c, int64_t ne, int64_t nn, const int64_t *elNodes, const double *nocoord,
                             double E, double nu, double density) {
  FCVM_CHECK(c && elNodes && nocoord, FCVM_E_ARG, "fcvm_set_mesh: null argument");
  FCVM_CHECK(ne > 0 && nn > 0, FCVM_E_ARG, "fcvm_set_mesh: empty mesh (ne=%lld, nn=%lld)", (long long)ne,
             (long long)nn);
  FCVM_CHECK(ne < 21000000 && nn < 2000000000, FCVM_E_ARG, "fcvm_set_mesh: mesh too large for 32-bit indices");
  FCVM_CUDA(cudaSetDevice(c->device));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  free_mesh(c);
  c->ne = ne; c->nn = nn; c->E = E; c->nu = nu; c->density = density;
  cudaStream_t st = c->stream;

  // connectivity: 1-based AoS int64 -> 0-based SoA int32 (validated here, once)
  std::vector<int32_t> conn((size_t)10 * ne);
  for (int64_t e = 0; e < ne; e++)
    for (int j = 0; j < 10; j++) {
      int64_t nd = elNodes[10 * e + j];
      if (nd < 1 || nd > nn) {
        set_error("fcvm_set_mesh: element %lld node %d = %lld outside 1..%lld", (long long)e, j, (long long)nd,
                  (long long)nn);
        return FCVM_E_MESH;
      }
      conn[(size_t)j * ne + e] = (int32_t)(nd - 1);
    }
  FCVM_TRY(dalloc(&c->conn, 10 * ne));
  FCVM_TRY(dalloc(&c->xyz, 3 * nn));
  FCVM_CUDA(cudaMemcpyAsync(c->conn, conn.data(), sizeof(int32_t) * 10 * ne, cudaMemcpyHostToDevice, st));
  FCVM_CUDA(cudaMemcpyAsync(c->xyz, nocoord, sizeof(double) * 3 * nn, cudaMemcpyHostToDevice, st));
  FCVM_CUDA(cudaStreamSynchronize(st));
  conn.clear(); conn.shrink_to_fit();

  // state buffers
  const int64_t n24 = 24 * ne, n4 = 4 * ne, n3 = 3 * nn;
  for (int i = FCVM_BUF_SIG_OLD; i <= FCVM_BUF_SIG_TEST; i++) {
    FCVM_TRY(dalloc((double **)&c->buf[i], n24));
    FCVM_CUDA(cudaMemsetAsync(c->buf[i], 0, sizeof(double) * n24, st));
  }
  for (int i = FCVM_BUF_SIG_YIELD; i <= FCVM_BUF_ECR; i++) {
    FCVM_TRY(dalloc((double **)&c->buf[i], n4));
    FCVM_CUDA(cudaMemsetAsync(c->buf[i], 0, sizeof(double) * n4, st));
  }
  FCVM_TRY(dalloc((uint8_t **)&c->buf[FCVM_BUF_PGP], n4));
  FCVM_CUDA(cudaMemsetAsync(c->buf[FCVM_BUF_PGP], 0, n4, st));
  for (int i = FCVM_BUF_MODF; i <= FCVM_BUF_FIXDOF; i++) {
    FCVM_TRY(dalloc((double **)&c->buf[i], n3));
    FCVM_CUDA(cudaMemsetAsync(c->buf[i], 0, sizeof(double) * n3, st));
  }
  k_fill<<<grid_for(n3, 256), 256, 0, st>>>(n3, 1.0, (double *)c->buf[FCVM_BUF_FIXDOF]);
  FCVM_TRY(dalloc(&c->elv, 30 * ne));
  FCVM_TRY(dalloc(&c->fixmask, n3));
  FCVM_TRY(dalloc(&c->fixval, n3));
  FCVM_TRY(dalloc(&c->movmask, n3));
  FCVM_CUDA(cudaMemsetAsync(c->fixmask, 0, n3, st));
  FCVM_CUDA(cudaMemsetAsync(c->fixval, 0, sizeof(double) * n3, st));
  FCVM_CUDA(cudaMemsetAsync(c->movmask, 0, sizeof(double) * n3, st));
  FCVM_TRY(dalloc(&c->pcg_r, n3)); FCVM_TRY(dalloc(&c->pcg_z, n3));
  FCVM_TRY(dalloc(&c->pcg_p, n3)); FCVM_TRY(dalloc(&c->pcg_q, n3)); FCVM_TRY(dalloc(&c->pcg_s, n3));

  // ---- node -> (element, local node) map, ascending element within a node ---------------
  {
    const int64_t n10 = 10 * ne;
    int32_t *keys_in, *keys_out, *vals_in, *count;
    FCVM_TRY(dalloc(&keys_in, n10)); FCVM_TRY(dalloc(&keys_out, n10));
    FCVM_TRY(dalloc(&vals_in, n10)); FCVM_TRY(dalloc(&c->n2e_idx, n10));
    FCVM_TRY(dalloc(&count, nn + 1)); FCVM_TRY(dalloc(&c->n2e_ptr, nn + 1));
    FCVM_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (nn + 1), st));
    k_n2e_keys<<<grid_for(n10, 256), 256, 0, st>>>(ne, c->conn, keys_in, vals_in, count);
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, vals_in, c->n2e_idx, (int)n10, 0,
                                    bits_for((uint64_t)nn), st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, count, c->n2e_ptr, (int)(nn + 1), st);
    void *tmp;
    FCVM_CUDA(cudaMalloc(&tmp, std::max(tb, tb2)));
    cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys_out, vals_in, c->n2e_idx, (int)n10, 0,
                                    bits_for((uint64_t)nn), st);
    cub::DeviceScan::ExclusiveSum(tmp, tb2, count, c->n2e_ptr, (int)(nn + 1), st);
    FCVM_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp); cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(count);
  }

  // ---- block sparsity pattern: sort the 100*ne (row node, col node) pairs ----------------
  const int64_t ncon = 100 * ne;
  uint64_t *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  FCVM_TRY(dalloc(&keys_a, ncon)); FCVM_TRY(dalloc(&keys_b, ncon));
  FCVM_TRY(dalloc(&vals_a, ncon)); FCVM_TRY(dalloc(&vals_b, ncon));
  k_pair_keys<<<grid_for(ncon, 256), 256, 0, st>>>(ne, c->conn, keys_a, vals_a);
  {
    cub::DoubleBuffer<uint64_t> dk(keys_a, keys_b);
    cub::DoubleBuffer<uint32_t> dv(vals_a, vals_b);
    size_t tb = 0;
    const int nb = bits_for((uint64_t)nn);
    // two passes over 32-bit halves keep the radix sort on the bits that vary
    cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)ncon, 0, nb, st);
    void *tmp;
    FCVM_CUDA(cudaMalloc(&tmp, tb));
    cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)ncon, 0, nb, st);          // by column (stable)
    cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (int)ncon, 32, 32 + nb, st);    // then by row (stable)
    FCVM_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    if (dk.Current() != keys_a) std::swap(keys_a, keys_b);
    if (dv.Current() != vals_a) std::swap(vals_a, vals_b);
  }
  cudaFree(keys_b);
  cudaFree(vals_b);
  c->src = vals_a;   // sorted contributions, kept for every (re)assembly

  int32_t *head, *head_scan, *rowlen;
  FCVM_TRY(dalloc(&head, ncon)); FCVM_TRY(dalloc(&head_scan, ncon + 1)); FCVM_TRY(dalloc(&rowlen, nn + 1));
  FCVM_CUDA(cudaMemsetAsync(rowlen, 0, sizeof(int32_t) * (nn + 1), st));
  k_heads<<<grid_for(ncon, 256), 256, 0, st>>>(ncon, keys_a, head);
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, head, head_scan, (int)ncon, st);
    void *tmp;
    FCVM_CUDA(cudaMalloc(&tmp, tb));
    cub::DeviceScan::ExclusiveSum(tmp, tb, head, head_scan, (int)ncon, st);
    FCVM_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  int32_t last_scan = 0, last_head = 0;
  FCVM_CUDA(cudaMemcpy(&last_scan, head_scan + ncon - 1, sizeof(int32_t), cudaMemcpyDeviceToHost));
  FCVM_CUDA(cudaMemcpy(&last_head, head + ncon - 1, sizeof(int32_t), cudaMemcpyDeviceToHost));
  const int64_t nblk = (int64_t)last_scan + last_head;
  c->nblk_real = nblk;
  uint64_t *blk_key;
  uint32_t *first_real;
  FCVM_TRY(dalloc(&blk_key, nblk)); FCVM_TRY(dalloc(&first_real, nblk));
  k_block_table<<<grid_for(ncon, 256), 256, 0, st>>>(ncon, keys_a, head, head_scan, blk_key, first_real, rowlen);
  FCVM_TRY(dalloc(&c->row_first, nn + 1));
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, rowlen, c->row_first, (int)(nn + 1), st);
    void *tmp;
    FCVM_CUDA(cudaMalloc(&tmp, tb));
    cub::DeviceScan::ExclusiveSum(tmp, tb, rowlen, c->row_first, (int)(nn + 1), st);
    FCVM_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  cudaFree(keys_a); cudaFree(head); cudaFree(head_scan);

  // ---- SELL-32-sigma layout (host: nn integers) ------------------------------------------
  std::vector<int32_t> h_rowlen((size_t)nn);
  FCVM_CUDA(cudaMemcpy(h_rowlen.data(), rowlen, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
  cudaFree(rowlen);
  const int64_t nslices = (nn + SELL_C - 1) / SELL_C;
  c->nslices = nslices;
  std::vector<int32_t> slot_node((size_t)nslices * SELL_C, -1), node_slot((size_t)nn), slice_ptr((size_t)nslices + 1);
  {
    std::vector<int32_t> order((size_t)nn);
    std::iota(order.begin(), order.end(), 0);
    for (int64_t w0 = 0; w0 < nn; w0 += SELL_SIGMA) {
      int64_t w1 = std::min<int64_t>(nn, w0 + SELL_SIGMA);
      std::stable_sort(order.begin() + w0, order.begin() + w1,
                       [&](int32_t a, int32_t b) { return h_rowlen[a] > h_rowlen[b]; });
    }
    for (int64_t s = 0; s < nn; s++) {
      slot_node[s] = order[s];
      node_slot[order[s]] = (int32_t)s;
    }
    int64_t acc = 0;
    for (int64_t s = 0; s < nslices; s++) {
      slice_ptr[s] = (int32_t)acc;
      int32_t w = 0;
      for (int l = 0; l < SELL_C; l++) {
        int32_t nd = slot_node[s * SELL_C + l];
        if (nd >= 0) w = std::max(w, h_rowlen[nd]);
      }
      acc += w;
      FCVM_CHECK(acc * SELL_C < 2147483647LL, FCVM_E_ARG, "fcvm_set_mesh: matrix too large for 32-bit positions");
    }
    slice_ptr[nslices] = (int32_t)acc;
    c->nblk_stored = acc * SELL_C;
  }
  FCVM_TRY(dalloc(&c->slice_ptr, nslices + 1)); FCVM_TRY(dalloc(&c->slot_node, nslices * SELL_C));
  FCVM_TRY(dalloc(&c->node_slot, nn));
  FCVM_CUDA(cudaMemcpy(c->slice_ptr, slice_ptr.data(), sizeof(int32_t) * (nslices + 1), cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->slot_node, slot_node.data(), sizeof(int32_t) * nslices * SELL_C, cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->node_slot, node_slot.data(), sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
  FCVM_TRY(dalloc(&c->colidx, c->nblk_stored)); FCVM_TRY(dalloc(&c->blk_first, c->nblk_stored));
  FCVM_TRY(dalloc(&c->blk_cnt, c->nblk_stored)); FCVM_TRY(dalloc(&c->vals, 9 * c->nblk_stored));
  FCVM_TRY(dalloc(&c->diag_pos, nn)); FCVM_TRY(dalloc(&c->row_cols, nblk));
  FCVM_TRY(dalloc(&c->minv, 9 * nn));
  FCVM_CUDA(cudaMemsetAsync(c->diag_pos, 0xff, sizeof(int32_t) * nn, st));
  FCVM_CUDA(cudaMemsetAsync(c->vals, 0, sizeof(double) * 9 * c->nblk_stored, st));
  k_sell_init<<<(unsigned)nslices, SELL_C, 0, st>>>(nslices, c->slice_ptr, c->slot_node, c->colidx, c->blk_cnt,
                                                   c->blk_first);
  k_sell_fill<<<grid_for(nblk, 256), 256, 0, st>>>(nblk, ncon, blk_key, first_real, c->row_first, c->node_slot,
                                                   c->slice_ptr, c->colidx, c->blk_first, c->blk_cnt, c->diag_pos,
                                                   c->row_cols);
  FCVM_CUDA(cudaStreamSynchronize(st));
  FCVM_CUDA(cudaGetLastError());
  cudaFree(blk_key); cudaFree(first_real);
  FCVM_TRY(dalloc(&c->cooK, (int64_t)55 * 9 * ne));
  return matfree_set_mesh(c);
}

extern "C" int fcvm_set_constraints(fcvm_ctx *c, const uint8_t *fixmask, const double *fixval) {
  FCVM_CHECK(c && c->nn > 0 && fixmask && fixval, FCVM_E_ARG, "fcvm_set_constraints: call fcvm_set_mesh first");
  const int64_t n3 = 3 * c->nn;
  std::vector<double> fixdof((size_t)n3), mov((size_t)n3);
  for (int64_t i = 0; i < n3; i++) {
    fixdof[i] = fixmask[i] ? 0.0 : 1.0;
    mov[i] = (fixmask[i] && fixval[i] != 0.0) ? 1.0 : 0.0;      // movdof, fcVM.py:256-258
  }
  // a value on a free dof means nothing (modf = K * fixval would pick it up): only prescribed dofs carry one
  std::vector<double> val((size_t)n3);
  for (int64_t i = 0; i < n3; i++) val[i] = fixmask[i] ? fixval[i] : 0.0;
  FCVM_CUDA(cudaMemcpy(c->fixmask, fixmask, n3, cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->fixval, val.data(), sizeof(double) * n3, cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->movmask, mov.data(), sizeof(double) * n3, cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->buf[FCVM_BUF_FIXDOF], fixdof.data(), sizeof(double) * n3, cudaMemcpyHostToDevice));
  c->have_bcs = true;
  c->assembled = false;
  c->matrix_elastic = false;
  FCVM_TRY(matfree_set_constraints(c));
  // the sparsity of K Z (which blocks vanish) depends on the prescribed dofs: rebuilt at the next assembly
  c->defl_structure = false;
  c->defl_ready = false;
  return FCVM_OK;
}

// New nodal coordinates on the same mesh topology (the imperfect geometry of fcVM.py:1240): the sparsity pattern
// and all index structures stay, the geometry-dependent data is recomputed, the matrix must be assembled again.
extern "C" int fcvm_set_coordinates(fcvm_ctx *c, const double *nocoord) {
  FCVM_CHECK(c && c->nn > 0 && nocoord, FCVM_E_ARG, "fcvm_set_coordinates: call fcvm_set_mesh first");
  FCVM_CUDA(cudaMemcpyAsync(c->xyz, nocoord, sizeof(double) * 3 * c->nn, cudaMemcpyHostToDevice, c->stream));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  c->assembled = false;
  c->matrix_elastic = false;
  c->defl_ready = c->defl_structure = false;
  dfree(c->egeo); dfree(c->tile_affine);
  return matfree_set_mesh(c);
}

extern "C" int fcvm_set_interface(fcvm_ctx *c, const double *dof_weight, int64_t n_if_local,
                                  const int64_t *if_local_node, const int64_t *if_global_slot, int64_t n_if_global) {
  FCVM_CHECK(c && c->nn > 0, FCVM_E_ARG, "fcvm_set_interface: call fcvm_set_mesh first");
  dfree(c->dof_weight); dfree(c->if_node); dfree(c->if_slot); dfree(c->if_buf);
  dfree(c->bslices); dfree(c->islices); dfree(c->tail3);
  c->n_bslices = c->n_islices = 0;
  const int64_t n3 = 3 * c->nn;
  if (dof_weight) {
    FCVM_TRY(dalloc(&c->dof_weight, n3));
    FCVM_CUDA(cudaMemcpy(c->dof_weight, dof_weight, sizeof(double) * n3, cudaMemcpyHostToDevice));
  }
  c->n_if_local = n_if_local;
  c->n_if_global = n_if_global;
  if (n_if_local > 0) {
    std::vector<int32_t> nd((size_t)n_if_local), sl((size_t)n_if_local);
    for (int64_t i = 0; i < n_if_local; i++) {
      FCVM_CHECK(if_local_node[i] >= 0 && if_local_node[i] < c->nn && if_global_slot[i] >= 0 &&
                     if_global_slot[i] < n_if_global,
                 FCVM_E_ARG, "fcvm_set_interface: entry %lld out of range", (long long)i);
      nd[i] = (int32_t)if_local_node[i];
      sl[i] = (int32_t)if_global_slot[i];
    }
    FCVM_TRY(dalloc(&c->if_node, n_if_local)); FCVM_TRY(dalloc(&c->if_slot, n_if_local));
    FCVM_CUDA(cudaMemcpy(c->if_node, nd.data(), sizeof(int32_t) * n_if_local, cudaMemcpyHostToDevice));
    FCVM_CUDA(cudaMemcpy(c->if_slot, sl.data(), sizeof(int32_t) * n_if_local, cudaMemcpyHostToDevice));
  }
  FCVM_TRY(dalloc(&c->if_buf, 3 * n_if_global + 4));
  FCVM_TRY(dalloc(&c->tail3, 4));
  FCVM_CUDA(cudaMemset(c->tail3, 0, sizeof(double) * 4));
  {
    // slices that hold at least one interface row ("boundary") and the rest ("interior")
    std::vector<int32_t> node_slot((size_t)c->nn);
    FCVM_CUDA(cudaMemcpy(node_slot.data(), c->node_slot, sizeof(int32_t) * c->nn, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> isb((size_t)c->nslices, 0);
    for (int64_t i = 0; i < n_if_local; i++) isb[(size_t)(node_slot[(size_t)if_local_node[i]] / SELL_C)] = 1;
    std::vector<int32_t> bl, il;
    for (int64_t s = 0; s < c->nslices; s++) (isb[(size_t)s] ? bl : il).push_back((int32_t)s);
    c->n_bslices = (int64_t)bl.size();
    c->n_islices = (int64_t)il.size();
    FCVM_TRY(dalloc(&c->bslices, c->n_bslices));
    FCVM_TRY(dalloc(&c->islices, c->n_islices));
    if (!bl.empty())
      FCVM_CUDA(cudaMemcpy(c->bslices, bl.data(), sizeof(int32_t) * bl.size(), cudaMemcpyHostToDevice));
    if (!il.empty())
      FCVM_CUDA(cudaMemcpy(c->islices, il.data(), sizeof(int32_t) * il.size(), cudaMemcpyHostToDevice));
  }
  return FCVM_OK;
}

extern "C" int fcvm_set_un_nodes(fcvm_ctx *c, int64_t n) {
  FCVM_CHECK(c && c->nn > 0 && n >= 0 && n <= c->nn, FCVM_E_ARG, "fcvm_set_un_nodes: bad argument");
  c->un_nodes = n;
  return FCVM_OK;
}

namespace fcvm {
int comm_allreduce_on(fcvm_ctx *c, double *dev, int64_t n, cudaStream_t st);
}

static int interface_sum_impl(fcvm_ctx *c, double *v, cudaStream_t st) {
  FCVM_CHECK(c && v, FCVM_E_ARG, "fcvm_interface_sum: null argument");
  if (c->world <= 1) return FCVM_OK;
  FCVM_CHECK(c->if_buf, FCVM_E_ARG, "fcvm_interface_sum: call fcvm_set_interface first");
  FCVM_CUDA(cudaMemsetAsync(c->if_buf, 0, sizeof(double) * 3 * c->n_if_global, st));
  if (c->n_if_local > 0)
    k_if_pack<<<grid_for(3 * c->n_if_local, 256), 256, 0, st>>>(c->n_if_local, c->if_node, c->if_slot, v, c->if_buf);
  FCVM_TRY(comm_allreduce_on(c, c->if_buf, 3 * c->n_if_global, st));
  if (c->n_if_local > 0)
    k_if_unpack<<<grid_for(3 * c->n_if_local, 256), 256, 0, st>>>(c->n_if_local, c->if_node, c->if_slot, c->if_buf,
                                                                 v);
  c->launches += 2;
  return FCVM_OK;
}

extern "C" int fcvm_interface_sum(fcvm_ctx *c, double *v) {
  if (c && c->world > 1 && c->n_if_global == 0) return FCVM_OK;
  return interface_sum_impl(c, v, c->stream);
}

namespace fcvm {
// the exchange on the communication stream (the caller orders it against the compute stream with events)
int interface_sum_on_comm_stream(fcvm_ctx *c, double *v) {
  if (c && c->world > 1 && c->n_if_global == 0) return FCVM_OK;      // a partition without shared nodes
  return interface_sum_impl(c, v, c->comm_stream);
}
}

// ---- vectors ----------------------------------------------------------------------------
extern "C" int fcvm_vec_alloc(fcvm_ctx *c, int64_t n, double **out) {
  FCVM_CHECK(c && out && n > 0, FCVM_E_ARG, "fcvm_vec_alloc: bad argument");
  FCVM_CUDA(cudaSetDevice(c->device));
  FCVM_TRY(dalloc(out, n));
  FCVM_CUDA(cudaMemsetAsync(*out, 0, sizeof(double) * n, c->stream));
  return FCVM_OK;
}

extern "C" int fcvm_vec_free(fcvm_ctx *c, double *v) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  if (v) {
    FCVM_CUDA(cudaStreamSynchronize(c->stream));
    FCVM_CUDA(cudaFree(v));
  }
  return FCVM_OK;
}

extern "C" int fcvm_buf(fcvm_ctx *c, int which, void **out, int64_t *n) {
  FCVM_CHECK(c && out && which >= 0 && which < FCVM_BUF_COUNT && c->nn > 0, FCVM_E_ARG, "fcvm_buf: bad argument");
  *out = c->buf[which];
  if (n) *n = which <= FCVM_BUF_SIG_TEST ? 24 * c->ne : (which <= FCVM_BUF_PGP ? 4 * c->ne : 3 * c->nn);
  return FCVM_OK;
}

extern "C" int fcvm_h2d(fcvm_ctx *c, void *dst, const void *src, int64_t bytes) {
  FCVM_CHECK(c && dst && src && bytes >= 0, FCVM_E_ARG, "fcvm_h2d: bad argument");
  FCVM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  c->h2d_bytes += bytes;
  return FCVM_OK;
}

extern "C" int fcvm_d2h(fcvm_ctx *c, void *dst, const void *src, int64_t bytes) {
  FCVM_CHECK(c && dst && src && bytes >= 0, FCVM_E_ARG, "fcvm_d2h: bad argument");
  FCVM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  c->d2h_bytes += bytes;
  return FCVM_OK;
}

extern "C" int fcvm_vec_zero(fcvm_ctx *c, int64_t n, double *x) {
  FCVM_CHECK(c && x, FCVM_E_ARG, "fcvm_vec_zero: null argument");
  FCVM_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, c->stream));
  return FCVM_OK;
}

extern "C" int fcvm_vec_copy(fcvm_ctx *c, int64_t n, const double *x, double *y) {
  FCVM_CHECK(c && x && y, FCVM_E_ARG, "fcvm_vec_copy: null argument");
  FCVM_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  return FCVM_OK;
}

extern "C" int fcvm_vec_axpby(fcvm_ctx *c, int64_t n, double a, const double *x, double b, double *y) {
  FCVM_CHECK(c && x && y, FCVM_E_ARG, "fcvm_vec_axpby: null argument");
  k_axpby<<<grid_for(n, 256), 256, 0, c->stream>>>(n, a, x, b, y);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

extern "C" int fcvm_vec_axpbypcz(fcvm_ctx *c, int64_t n, double a, const double *x, double b, const double *y,
                                 double cc, double *z) {
  FCVM_CHECK(c && x && y && z, FCVM_E_ARG, "fcvm_vec_axpbypcz: null argument");
  k_axpbypcz<<<grid_for(n, 256), 256, 0, c->stream>>>(n, a, x, b, y, cc, z);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

extern "C" int fcvm_vec_dot(fcvm_ctx *c, int64_t n, const double *x, const double *y, double *out) {
  FCVM_CHECK(c && x && y && out, FCVM_E_ARG, "fcvm_vec_dot: null argument");
  const double *w = (n == 3 * c->nn) ? c->dof_weight : nullptr;
  k_dot<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, x, y, w, c->red_part, c->red_counter, c->red_out);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  FCVM_TRY(finish_scalar(c, 1, true));
  *out = c->h_scalars[0];
  return FCVM_OK;
}

extern "C" int fcvm_residual(fcvm_ctx *c, double lbd, const double *glv, const double *qin, double *r,
                             double *rnorm) {
  FCVM_CHECK(c && glv && qin && r && rnorm, FCVM_E_ARG, "fcvm_residual: null argument");
  k_residual<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(3 * c->nn, lbd, glv, qin,
                                                        (const double *)c->buf[FCVM_BUF_FIXDOF], c->dof_weight, r,
                                                        c->red_part, c->red_counter, c->red_out);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  FCVM_TRY(finish_scalar(c, 1, true));
  *rnorm = sqrt(c->h_scalars[0]);
  return FCVM_OK;
}

extern "C" int fcvm_max_node_disp(fcvm_ctx *c, const double *disp, double *out) {
  FCVM_CHECK(c && disp && out, FCVM_E_ARG, "fcvm_max_node_disp: null argument");
  // (ndof - 1) // 3 nodes, as the reference (fcVM.py:1494-1497); on a partitioned mesh the rank that
  // holds the last global node leaves it out (fcvm_set_un_nodes) and the maximum is taken over ranks
  const int64_t nodes = c->un_nodes >= 0 ? c->un_nodes : (3 * c->nn - 1) / 3;
  k_max_node_disp<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(nodes, disp, c->red_part, c->red_counter, c->red_out);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  if (c->world > 1) FCVM_TRY(fcvm_comm_allreduce_max(c, c->red_out, 1));
  FCVM_TRY(read_scalars(c, 1));
  *out = sqrt(c->h_scalars[0]);
  return FCVM_OK;
}

extern "C" int fcvm_reaction(fcvm_ctx *c, const double *qin, double *out) {
  FCVM_CHECK(c && qin && out, FCVM_E_ARG, "fcvm_reaction: null argument");
  k_dot<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(3 * c->nn, c->movmask, qin, c->dof_weight, c->red_part,
                                                   c->red_counter, c->red_out);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  FCVM_TRY(finish_scalar(c, 1, true));
  *out = c->h_scalars[0];
  return FCVM_OK;
}

// ---- Gauss-point arrays -------------------------------------------------------------------
static int ensure_gp_tmp(fcvm_ctx *c) {
  if (c->gp_tmp) return FCVM_OK;
  return dalloc(&c->gp_tmp, 24 * c->ne);
}

// pinned host memory for the callers of the fcvm_host_* entry points
extern "C" int fcvm_host_alloc(int64_t bytes, void **out) {
  FCVM_CHECK(out && bytes > 0, FCVM_E_ARG, "fcvm_host_alloc: bad argument");
  FCVM_CUDA(cudaMallocHost(out, (size_t)bytes));
  return FCVM_OK;
}

extern "C" int fcvm_host_free(void *p) {
  if (p) FCVM_CUDA(cudaFreeHost(p));
  return FCVM_OK;
}

extern "C" int fcvm_gp_to_host(fcvm_ctx *c, const double *dev_soa, int ncomp, double *host_aos) {
  FCVM_CHECK(c && dev_soa && host_aos && (ncomp == 6 || ncomp == 1), FCVM_E_ARG, "fcvm_gp_to_host: bad argument");
  const int64_t n = 4 * c->ne * ncomp;
  FCVM_TRY(ensure_gp_tmp(c));
  k_gp_soa_to_aos<<<grid_for(n, 256), 256, 0, c->stream>>>(c->ne, ncomp, dev_soa, c->gp_tmp);
  c->launches++;
  return fcvm_d2h(c, host_aos, c->gp_tmp, sizeof(double) * n);
}

extern "C" int fcvm_gp_from_host(fcvm_ctx *c, const double *host_aos, int ncomp, double *dev_soa) {
  FCVM_CHECK(c && dev_soa && host_aos && (ncomp == 6 || ncomp == 1), FCVM_E_ARG, "fcvm_gp_from_host: bad argument");
  const int64_t n = 4 * c->ne * ncomp;
  FCVM_TRY(ensure_gp_tmp(c));
  FCVM_TRY(fcvm_h2d(c, c->gp_tmp, host_aos, sizeof(double) * n));
  k_gp_aos_to_soa<<<grid_for(n, 256), 256, 0, c->stream>>>(c->ne, ncomp, c->gp_tmp, dev_soa);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

extern "C" int fcvm_gp_fill(fcvm_ctx *c, int which, double value) {
  FCVM_CHECK(c && which >= 0 && which <= FCVM_BUF_ECR && c->ne > 0, FCVM_E_ARG, "fcvm_gp_fill: bad buffer");
  const int64_t n = which <= FCVM_BUF_SIG_TEST ? 24 * c->ne : 4 * c->ne;
  k_fill<<<grid_for(n, 256), 256, 0, c->stream>>>(n, value, (double *)c->buf[which]);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

extern "C" int fcvm_pgp_to_host(fcvm_ctx *c, uint8_t *host) {
  FCVM_CHECK(c && host && c->ne > 0, FCVM_E_ARG, "fcvm_pgp_to_host: bad argument");
  FCVM_TRY(ensure_gp_tmp(c));
  uint8_t *tmp = (uint8_t *)c->gp_tmp;
  k_pgp_soa_to_aos<<<grid_for(4 * c->ne, 256), 256, 0, c->stream>>>(c->ne, (const uint8_t *)c->buf[FCVM_BUF_PGP],
                                                                   tmp);
  c->launches++;
  return fcvm_d2h(c, host, tmp, 4 * c->ne);
}

extern "C" int fcvm_pgp_count(fcvm_ctx *c, int64_t *n_plastic) {
  FCVM_CHECK(c && n_plastic && c->ne > 0, FCVM_E_ARG, "fcvm_pgp_count: bad argument");
  k_count_u8<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(4 * c->ne, (const uint8_t *)c->buf[FCVM_BUF_PGP],
                                                        c->red_part, c->red_counter, c->red_out);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  FCVM_TRY(finish_scalar(c, 1, true));
  *n_plastic = (int64_t)(c->h_scalars[0] + 0.5);
  return FCVM_OK;
}

// ---- timing ---------------------------------------------------------------------------------
extern "C" int fcvm_timer_start(fcvm_ctx *c) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  FCVM_CUDA(cudaEventRecord(c->ev0, c->stream));
  return FCVM_OK;
}

extern "C" int fcvm_timer_stop_ms(fcvm_ctx *c, float *ms) {
  FCVM_CHECK(c && ms, FCVM_E_ARG, "null argument");
  FCVM_CUDA(cudaEventRecord(c->ev1, c->stream));
  FCVM_CUDA(cudaEventSynchronize(c->ev1));
  FCVM_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return FCVM_OK;
}

static int resolve_samples(fcvm_ctx *c) {
  if (c->prof_used == 0) return FCVM_OK;
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < c->prof_used; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->prof_pool[i].e0, c->prof_pool[i].e1) == cudaSuccess) {
      c->prof.ms[c->prof_pool[i].which] += ms;
      c->prof.launches[c->prof_pool[i].which] += 1;
    }
  }
  c->prof_used = 0;
  return FCVM_OK;
}

extern "C" int fcvm_profile_enable(fcvm_ctx *c, int on) {
  FCVM_CHECK(c && on >= 0, FCVM_E_ARG, "fcvm_profile_enable: bad argument");
  FCVM_TRY(resolve_samples(c));
  if (on >= 2 && c->prof_pool.empty()) {
    c->prof_pool.resize(PROF_POOL);
    for (auto &s : c->prof_pool) {
      FCVM_CUDA(cudaEventCreate(&s.e0));
      FCVM_CUDA(cudaEventCreate(&s.e1));
    }
  }
  c->profiling = on >= 2 ? 2 : on;
  if (on >= 2) c->prof_stride = on;      // on = stride of the sampled mode (>= 2)
  return FCVM_OK;
}

extern "C" int fcvm_profile_get(fcvm_ctx *c, int which, double *ms, int64_t *launches) {
  FCVM_CHECK(c && which >= 0 && which < NUM_PROFILE, FCVM_E_ARG, "fcvm_profile_get: bad index");
  FCVM_TRY(resolve_samples(c));
  if (ms) *ms = c->prof.ms[which];
  if (launches) *launches = c->prof.launches[which];
  return FCVM_OK;
}

extern "C" int fcvm_profile_reset(fcvm_ctx *c) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  FCVM_TRY(resolve_samples(c));
  memset(&c->prof, 0, sizeof(c->prof));
  return FCVM_OK;
}

extern "C" int64_t fcvm_profile_seen(fcvm_ctx *c, int which) {
  return (c && which >= 0 && which < NUM_PROFILE) ? c->prof.seen[which] : 0;
}

extern "C" int fcvm_matrix_stats(fcvm_ctx *c, int64_t *stored, int64_t *real, int64_t *bytes) {
  FCVM_CHECK(c && c->nn > 0, FCVM_E_ARG, "fcvm_matrix_stats: no mesh");
  if (stored) *stored = c->nblk_stored;
  if (real) *real = c->nblk_real;
  if (bytes) *bytes = c->nblk_stored * (9 * 8 + 4);
  return FCVM_OK;
}
