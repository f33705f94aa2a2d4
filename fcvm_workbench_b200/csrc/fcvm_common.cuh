// Shared definitions of the fcvm_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/fcvm_b200.h"

namespace fcvm {

void set_error(const char *fmt, ...);

#define FCVM_CUDA(call)                                                                         \
  do {                                                                                          \
    cudaError_t err__ = (call);                                                                 \
    if (err__ != cudaSuccess) {                                                                 \
      fcvm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
      return FCVM_E_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

#define FCVM_CHECK(cond, code, ...)   \
  do {                                \
    if (!(cond)) {                    \
      fcvm::set_error(__VA_ARGS__);   \
      return (code);                  \
    }                                 \
  } while (0)

#define FCVM_TRY(expr)            \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != FCVM_OK) return rc__; \
  } while (0)

constexpr int SELL_C = 32;           // rows per SELL slice = one warp
constexpr int SELL_SIGMA = 4096;     // sorting window (rows)
constexpr int RED_BLOCKS = 592;      // 4 x 148 SMs: fixed shape of every two-stage reduction
constexpr int RED_THREADS = 256;
constexpr int NUM_PROFILE = 12;

// Gauss points of the 10-node tetrahedron (fcVM.py:589-596)
constexpr double GP_A = 0.138196601125011;
constexpr double GP_B = 0.585410196624968;
constexpr double GP_W = 0.041666666666667;

struct Profile {
  double ms[NUM_PROFILE];        // summed duration of the timed launches
  int64_t launches[NUM_PROFILE]; // launches that were timed
  int64_t seen[NUM_PROFILE];     // all launches since the reset
};

// asynchronous sampling: event pairs recorded around every stride-th launch of a family,
// resolved (one synchronisation) when the profile is read -- the timed region is not disturbed
constexpr int PROF_POOL = 4096;
struct ProfSample {
  int which;
  cudaEvent_t e0, e1;
};

struct P2PState;                 // peer-memory exchange of the multi-GPU PCG (fcvm_p2p.cu)

}  // namespace fcvm

struct fcvm_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // fcvm_timer_*
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;    // profiling
  int profiling = 0;             // 0 off, 1 every launch (synchronous), 2 sampled (asynchronous)
  int prof_stride = 8;
  fcvm::Profile prof{};
  std::vector<fcvm::ProfSample> prof_pool;
  size_t prof_used = 0;
  int64_t launches = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;   // bytes moved over PCIe by fcvm_h2d / fcvm_d2h and the fcvm_host_* entry points

  // mesh
  int64_t ne = 0, nn = 0;
  double E = 0, nu = 0, density = 0;
  int32_t *conn = nullptr;      // [10][ne] zero-based
  double *xyz = nullptr;        // [nn][3]
  int32_t *n2e_ptr = nullptr;   // [nn+1]
  int32_t *n2e_idx = nullptr;   // [10*ne]  e*10+j, ascending e within a node
  double *elv = nullptr;        // [ne][30] element vectors (scratch of the gather)

  // constraints
  uint8_t *fixmask = nullptr;   // [3nn]
  double *fixval = nullptr;     // [3nn]
  double *movmask = nullptr;    // [3nn] 1.0 where a non-zero value is prescribed
  bool have_bcs = false;

  // Gauss-point state and named nodal buffers
  void *buf[FCVM_BUF_COUNT] = {nullptr};

  // matrix: block-SELL
  int64_t nslices = 0;          // ceil(nn/32)
  int64_t nblk_real = 0;        // distinct (row node, col node) pairs
  int64_t nblk_stored = 0;      // incl. padding = 32 * sum(slice widths)
  int32_t *slice_ptr = nullptr; // [nslices+1] in units of block-columns (x32 blocks)
  int32_t *slot_node = nullptr; // [nslices*32] node of SELL slot, -1 = padding row
  int32_t *node_slot = nullptr; // [nn]
  int32_t *colidx = nullptr;    // [nblk_stored] block column (node)
  double *vals = nullptr;       // [nblk_stored/32][9][32]
  double *vals2 = nullptr;      // same layout: geometric stiffness G of the linear buckling analysis
  uint32_t *blk_first = nullptr;// [nblk_stored] first contribution in src
  uint32_t *blk_cnt = nullptr;  // [nblk_stored] number of contributions (0 for padding)
  uint32_t *src = nullptr;      // [100*ne] ((pair*ne + e) << 1) | transpose
  int32_t *diag_pos = nullptr;  // [nn] position of the diagonal block
  int32_t *row_first = nullptr; // [nn+1] CSR-like row pointer over real blocks (for export)
  int32_t *row_cols = nullptr;  // [nblk_real] sorted block columns (for export)
  double *cooK = nullptr;       // [55][ne][9] element stiffness blocks (lower block triangle)
  double *minv = nullptr;       // [nn][9] inverse diagonal blocks
  bool assembled = false;
  bool matrix_elastic = false;  // the assembled operator is calcGSM's elastic one: the PCG may apply it matrix-free
  uint32_t *emask = nullptr;    // [ne] bit 3k+c: dof c of local node k prescribed (matrix-free product)
  unsigned int *ga_ticket = nullptr;  // [1 + groups] tickets of the gather's two-level dot-product finish (zero at rest)
  double *ga_group_part = nullptr;    // [2][groups]
  double *egeo = nullptr;       // [10][ne] inverse Jacobian (9) and w|J| (1) of the straight-sided elements
  uint8_t *tile_affine = nullptr; // [ceil(ne/32)] all elements of the 32-element tile are straight-sided
  int64_t n_affine_tiles = 0;

  // PCG work vectors
  double *pcg_r = nullptr, *pcg_z = nullptr, *pcg_p = nullptr, *pcg_q = nullptr, *pcg_s = nullptr;
  double *spmv_part = nullptr;  // block partials of the dot product fused into the SpMV

  // recycled start vectors: the last (right-hand side, solution) pairs of solves with the present matrix
  double *hist_b[2] = {nullptr, nullptr}, *hist_x[2] = {nullptr, nullptr};
  int hist_n = 0;               // pairs held (0..2), oldest first
  double hist_gram[2][2] = {{0, 0}, {0, 0}};   // x_i . b_j

  // reductions
  double *red_part = nullptr;   // [8][RED_BLOCKS]
  double *red_out = nullptr;    // [16] device scalars
  unsigned int *red_counter = nullptr;
  double *h_scalars = nullptr;  // pinned [16]
  int64_t *d_arg = nullptr;     // argmax result
  int64_t *d_arg_part = nullptr;// [RED_BLOCKS] argmax partials
  int64_t *h_arg = nullptr;     // pinned

  // multi-GPU
  void *nccl_comm = nullptr;
  int rank = 0, world = 1;
  int64_t un_nodes = -1;        // nodes entering fcvm_max_node_disp (-1: nn - 1, the reference's range)
  double *dof_weight = nullptr; // [3nn] 1/multiplicity (nullptr = 1)
  int64_t n_if_local = 0, n_if_global = 0;
  int32_t *if_node = nullptr;   // [n_if_local]
  int32_t *if_slot = nullptr;   // [n_if_local]
  double *if_buf = nullptr;     // [3*n_if_global]
  // overlapped interface exchange of the PCG: slices holding interface rows go first, their exchange
  // runs on comm_stream while the interior slices are multiplied
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
  int32_t *bslices = nullptr, *islices = nullptr;
  int64_t n_bslices = 0, n_islices = 0;
  double *tail3 = nullptr;      // [4] per-rank PCG sums on their way through the scalar all-reduce
  fcvm::P2PState *p2p = nullptr; // arenas mapped between the ranks of one box (CUDA IPC over NVLink)
  bool p2p_attached = false;

  // rigid-body-mode deflation of the PCG (fcvm_deflation.cu): box clusters of nodes, six modes each
  int dn[3] = {0, 0, 0};        // clusters per direction (0 = deflation off)
  double dlo[3] = {0, 0, 0}, dh[3] = {1, 1, 1}, dscale = 1.0;
  int64_t ncl = 0;              // dn[0]*dn[1]*dn[2]
  int32_t *d_cid = nullptr;     // [nn] cluster of each node
  uint8_t *cl_active = nullptr; // [ncl] box carries its six modes
  int32_t *cl_ptr = nullptr, *cl_nodes = nullptr;     // nodes of each cluster, ascending
  int8_t *kz_rel = nullptr;     // [nn][8] relative position code (0..26) of the cluster a slot couples to, -1 unused
  double *kz_val = nullptr;     // [18][nent] (K Z)_(i, cluster) 3 x 6, entry-ordered, component-major
  int32_t *ent_ptr = nullptr, *ent = nullptr;         // per target cluster: its entries (node of each), ascending
  int32_t *ent_inv = nullptr;   // [nn][8] (node, slot) -> entry, -1 = dropped
  int64_t nent = 0;
  double *dE = nullptr, *dEinv = nullptr;             // [6 ncl][6 ncl]
  double *d_rhs = nullptr, *d_lam = nullptr;          // [6 ncl]
  double *spmv_part2 = nullptr; // per-slice partials of r.u
  // single-precision copies of the coarse operators (they only shape the preconditioner)
  float *kz32 = nullptr;        // [18][nent]
  float *einv32 = nullptr;      // [6 ncl][einv_ld]
  int64_t einv_ld = 0;          // row stride of einv32: 6 ncl rounded up to 4 floats, padding zero
  double *rhs_part = nullptr;   // [RHS_SPLIT][6 ncl] shares of the coarse right-hand side
  int64_t col0 = 0, col1 = 0;   // columns of E^-1 this rank's right-hand side can be non-zero in
  int64_t local_boxes = 0;      // boxes that hold nodes of this rank
  void *cusolver = nullptr;
  double *cus_work = nullptr;
  int cus_lwork = 0;
  int *cus_info = nullptr;
  bool defl_structure = false, defl_ready = false;

  // host staging / device scratch of fcvm_host_*
  double *gp_tmp = nullptr;     // [24*ne] device scratch of the Gauss-point layout conversions
  double *h_du = nullptr, *h_disp = nullptr, *h_qin = nullptr;
  // pipelined fcvm_host_update_stress_load: full-size device staging in the reference layout (sig 24 + sig_yield 4
  // in, sig_new 24 + sig_test 24 out, flags), a copy-in and a copy-out stream, one event pair per chunk
  double *hs_in = nullptr, *hs_out = nullptr;
  uint8_t *hs_pgp = nullptr;
  cudaStream_t h_in_stream = nullptr, h_out_stream = nullptr;
  cudaEvent_t h_ev_in[16] = {nullptr}, h_ev_k[16] = {nullptr};
  double *diag9 = nullptr;      // [3][nn][3] assembled diagonal blocks, row-wise
};

namespace fcvm {

inline int grid_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// profiling wrapper: CUDA events on the launching stream around one launch
struct ProfScope {
  fcvm_ctx *c;
  int which;
  fcvm::ProfSample *smp = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;        // synchronous mode: own pair, so scopes may nest
  ProfScope(fcvm_ctx *ctx, int w) : c(ctx), which(w) {
    if (!c->profiling) return;
    const int64_t k = c->prof.seen[which]++;
    if (c->profiling == 1) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, c->stream);
    } else if (((k % c->prof_stride) == 0 || which == 1 || which == 2 || (which >= 4 && which <= 6)) &&
               c->prof_used < c->prof_pool.size()) {
      // the families that run once per Newton iteration (or per assembly) are timed every time, the per-PCG-iteration
      // ones every stride-th launch
      smp = &c->prof_pool[c->prof_used++];
      smp->which = which;
      cudaEventRecord(smp->e0, c->stream);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      c->prof.ms[which] += ms;
      c->prof.launches[which] += 1;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    } else if (smp) {
      cudaEventRecord(smp->e1, c->stream);
    }
  }
};

// ---------------------------------------------------------------------------------------
// 10-node tetrahedron, run-time Gauss point (local derivative table of fcVM.py:390-424): the non-zero
// entries of the derivative table are ten numbers that depend
// on the point only through (xi, eta, zeta).  One code path for all four points keeps the kernel
// small enough for the instruction cache (four compile-time copies of the stress update did not).
// ---------------------------------------------------------------------------------------
struct GPCoef {
  double a4, d01, d04, d12, d16, d23, d27, x4, e4, z4;
};

__device__ __forceinline__ GPCoef gp_coef(int gp) {
  const double xi = gp == 1 ? GP_B : GP_A, et = gp == 2 ? GP_B : GP_A, ze = gp == 3 ? GP_B : GP_A;
  GPCoef c;
  c.a4 = 1.0 - 4.0 * (1.0 - xi - et - ze);
  c.d01 = 4.0 * xi - 1.0;
  c.d04 = 4.0 * (1.0 - 2.0 * xi - et - ze);
  c.d12 = 4.0 * et - 1.0;
  c.d16 = 4.0 * (1.0 - xi - 2.0 * et - ze);
  c.d23 = 4.0 * ze - 1.0;
  c.d27 = 4.0 * (1.0 - xi - et - 2.0 * ze);
  c.x4 = 4.0 * xi;
  c.e4 = 4.0 * et;
  c.z4 = 4.0 * ze;
  return c;
}

// out[i][j] = sum_k v_k[i] * dN[j][k] with the nodal values read from a [30][stride] tile (row 3k+i)
__device__ __forceinline__ void local_gradient_tile(const GPCoef &c, const double *tile, int stride,
                                                    double (&out)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double v0 = tile[(0 + i) * stride], v1 = tile[(3 + i) * stride], v2 = tile[(6 + i) * stride];
    const double v3 = tile[(9 + i) * stride], v4 = tile[(12 + i) * stride], v5 = tile[(15 + i) * stride];
    const double v6 = tile[(18 + i) * stride], v7 = tile[(21 + i) * stride], v8 = tile[(24 + i) * stride];
    const double v9 = tile[(27 + i) * stride];
    out[i][0] = c.a4 * v0 + c.d01 * v1 + c.d04 * v4 + c.e4 * (v5 - v6) + c.z4 * (v8 - v7);
    out[i][1] = c.a4 * v0 + c.d12 * v2 + c.x4 * (v5 - v4) + c.d16 * v6 + c.z4 * (v9 - v7);
    out[i][2] = c.a4 * v0 + c.d23 * v3 - c.x4 * v4 - c.e4 * v6 + c.d27 * v7 + c.x4 * v8 + c.e4 * v9;
  }
}

// F[k][i] = sum_j T[i][j] * dN[j][k] stored straight to a [30][stride] tile (row 3k+i)
__device__ __forceinline__ void store_gradient_tile(const GPCoef &c, const double (&T)[3][3], double *tile,
                                                    int stride) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double t0 = T[i][0], t1 = T[i][1], t2 = T[i][2];
    tile[(0 + i) * stride] = c.a4 * (t0 + t1 + t2);
    tile[(3 + i) * stride] = c.d01 * t0;
    tile[(6 + i) * stride] = c.d12 * t1;
    tile[(9 + i) * stride] = c.d23 * t2;
    tile[(12 + i) * stride] = c.d04 * t0 - c.x4 * (t1 + t2);
    tile[(15 + i) * stride] = c.e4 * t0 + c.x4 * t1;
    tile[(18 + i) * stride] = c.d16 * t1 - c.e4 * (t0 + t2);
    tile[(21 + i) * stride] = c.d27 * t2 - c.z4 * (t0 + t1);
    tile[(24 + i) * stride] = c.z4 * t0 + c.x4 * t2;
    tile[(27 + i) * stride] = c.z4 * t1 + c.e4 * t2;
  }
}

// the same for two Gauss points at once: every staged value is read once and feeds both
__device__ __forceinline__ void local_gradient_tile2(const GPCoef &a, const GPCoef &b, const double *tile, int stride,
                                                     double (&oa)[3][3], double (&ob)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double v0 = tile[(0 + i) * stride], v1 = tile[(3 + i) * stride], v2 = tile[(6 + i) * stride];
    const double v3 = tile[(9 + i) * stride], v4 = tile[(12 + i) * stride], v5 = tile[(15 + i) * stride];
    const double v6 = tile[(18 + i) * stride], v7 = tile[(21 + i) * stride], v8 = tile[(24 + i) * stride];
    const double v9 = tile[(27 + i) * stride];
    const double d56 = v5 - v6, d87 = v8 - v7, d54 = v5 - v4, d97 = v9 - v7;
    oa[i][0] = a.a4 * v0 + a.d01 * v1 + a.d04 * v4 + a.e4 * d56 + a.z4 * d87;
    oa[i][1] = a.a4 * v0 + a.d12 * v2 + a.x4 * d54 + a.d16 * v6 + a.z4 * d97;
    oa[i][2] = a.a4 * v0 + a.d23 * v3 - a.x4 * v4 - a.e4 * v6 + a.d27 * v7 + a.x4 * v8 + a.e4 * v9;
    ob[i][0] = b.a4 * v0 + b.d01 * v1 + b.d04 * v4 + b.e4 * d56 + b.z4 * d87;
    ob[i][1] = b.a4 * v0 + b.d12 * v2 + b.x4 * d54 + b.d16 * v6 + b.z4 * d97;
    ob[i][2] = b.a4 * v0 + b.d23 * v3 - b.x4 * v4 - b.e4 * v6 + b.d27 * v7 + b.x4 * v8 + b.e4 * v9;
  }
}

// F[3k+i] (+)= sum_j T[i][j] * dN[j][k] in registers
template <bool ACC>
__device__ __forceinline__ void gradient_to_regs(const GPCoef &c, const double (&T)[3][3], double (&F)[30]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double t0 = T[i][0], t1 = T[i][1], t2 = T[i][2];
    const double f[10] = {c.a4 * (t0 + t1 + t2), c.d01 * t0, c.d12 * t1, c.d23 * t2, c.d04 * t0 - c.x4 * (t1 + t2),
                          c.e4 * t0 + c.x4 * t1, c.d16 * t1 - c.e4 * (t0 + t2), c.d27 * t2 - c.z4 * (t0 + t1),
                          c.z4 * t0 + c.x4 * t2, c.z4 * t1 + c.e4 * t2};
#pragma unroll
    for (int k = 0; k < 10; k++) F[3 * k + i] = ACC ? F[3 * k + i] + f[k] : f[k];
  }
}

// determinant and inverse of the Jacobian xs[i][j] = d x_i / d xi_j   (fcVM.py:428-453)
__device__ __forceinline__ double invert_jacobian(const double (&xs)[3][3], double (&xsi)[3][3]) {
  double xsj = (xs[0][0] * xs[1][1] * xs[2][2] - xs[0][0] * xs[1][2] * xs[2][1] + xs[0][2] * xs[1][0] * xs[2][1] -
                xs[0][2] * xs[1][1] * xs[2][0] + xs[0][1] * xs[1][2] * xs[2][0] - xs[0][1] * xs[1][0] * xs[2][2]);
  double inv = 1.0 / xsj;
  xsi[0][0] = (xs[1][1] * xs[2][2] - xs[2][1] * xs[1][2]) * inv;
  xsi[0][1] = (xs[0][2] * xs[2][1] - xs[0][1] * xs[2][2]) * inv;
  xsi[0][2] = (xs[0][1] * xs[1][2] - xs[0][2] * xs[1][1]) * inv;
  xsi[1][0] = (xs[1][2] * xs[2][0] - xs[1][0] * xs[2][2]) * inv;
  xsi[1][1] = (xs[0][0] * xs[2][2] - xs[0][2] * xs[2][0]) * inv;
  xsi[1][2] = (xs[1][0] * xs[0][2] - xs[0][0] * xs[1][2]) * inv;
  xsi[2][0] = (xs[1][0] * xs[2][1] - xs[2][0] * xs[1][1]) * inv;
  xsi[2][1] = (xs[2][0] * xs[0][1] - xs[0][0] * xs[2][1]) * inv;
  xsi[2][2] = (xs[0][0] * xs[1][1] - xs[1][0] * xs[0][1]) * inv;
  return xsj;
}

// ---------------------------------------------------------------------------------------
// Bulk asynchronous copies (the TMA unit without a tensor map: cp.async.bulk, UBLKCP in SASS) and the mbarrier
// that counts their bytes.  One elected thread issues; everybody waits on the barrier's phase.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

// block-level deterministic sum: fixed tree over the warp, then over warps in order
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace fcvm
