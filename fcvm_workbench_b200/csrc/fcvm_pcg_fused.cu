// The whole preconditioned-CG loop of x = factor(b) (fcVM.py:1130, 1401) in ONE persistent cooperative
// kernel.
//
// One block of 256 threads per co-resident slot (4 per SM, 592 on a B200); phases are separated by grid
// barriers instead of kernel boundaries, the scalars of the single-reduction (Chronopoulos-Gear) recurrence are
// recomputed by every block from the same block partials in the same order (bit-identical everywhere, so all
// blocks take the same exits), and the loop runs to convergence without the host.  Per iteration:
//
//   [deflation]  coarse partials   Z^T r and -(K Z)^T y over fixed chunks of the box lists     (streams K Z)
//                coarse finish     chunk partials -> right-hand side of the coarse problem
//                coarse product    lam = E^-1 rhs, rows x column quarters over the warps          (streams E^-1)
//                expand            u = y + Z lam
//   product      w = K u on the block-SELL matrix, slices dealt round by round to workers of 1..8 warps (streams K)
//                + block partials of w.u and r.u
//   step         p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; y = D^-1 r ; partials of r.r
//
// The coarse operators K Z and E^-1 only shape the preconditioner, so the kernel streams single-precision
// copies of them (half the bytes); every vector, the matrix and all recurrences stay FP64.  All reductions have
// a fixed shape and order: repeated runs are bit-identical.  Device time per phase is accumulated from
// %globaltimer by one thread (fcvm_pcg_phase_times).
#include <cooperative_groups.h>

#include <algorithm>

#include "fcvm_common.cuh"
#include "fcvm_deflation.cuh"
#include "fcvm_pcg.cuh"

namespace cg = cooperative_groups;
using namespace fcvm;

namespace {

#ifndef FUSED_FT
#define FUSED_FT 256
#endif
#ifndef FUSED_BPS
#define FUSED_BPS 2
#endif
constexpr int FT = FUSED_FT;     // threads per block
constexpr int FW = FT / 32;      // warps per block
constexpr int FUSED_BLOCKS_PER_SM = FUSED_BPS;
constexpr int COARSE_CHUNK = 2048;   // list entries per coarse work item

struct FusedArgs {
  int64_t nn, nslices;
  int max_iter, defl, split;
  // matrix (never written here)
  const int32_t *slice_ptr, *slot_node, *colidx;
  const double *vals, *minv, *wt;
  // vectors (read and written across phases: plain pointers, no read-only path)
  double *x, *r, *u, *p, *s, *w;
  double *part;                  // [4][grid] block partials: w.u, r.u | r.u (step), r.r
  double *sc;
  // deflation level
  Grid g;
  int64_t ncl, n6, nent, einv_ld;
  int n_items;
  const int32_t *cid, *cl_nodes, *ent_node, *it_box, *it_lo, *it_hi, *box_item_ptr;
  const uint8_t *it_kind;
  const double *xyz, *fixdof;
  const void *kz, *einv;         // CT = float (default) or double copies
  double *item_part, *rhs, *lam4;
  unsigned long long *phase_ns;
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// block partial sums of NV values -> part[v * G + block]; fixed tree
template <int NV>
__device__ __forceinline__ void block_partial(double (&v)[NV], double *part, int G, double *sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const double w = warp_sum(v[i]);
    if (lane == 0) sh[i * FW + warp] = w;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < FW; w++) s += sh[threadIdx.x * FW + w];
    part[(int64_t)threadIdx.x * G + blockIdx.x] = s;
  }
  __syncthreads();
}

// sum of the G block partials, the same additions in the same order in every block
__device__ __forceinline__ double grid_sum(const double *part, int G, double *sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v = 0.0;
  for (int i = threadIdx.x; i < G; i += FT) v += __ldcg(part + i);
  v = warp_sum(v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < FW; w++) t += sh[w];
  __syncthreads();
  return t;
}

// ---- deflation phases -----------------------------------------------------------------------------------
template <typename CT>
__device__ __forceinline__ void coarse_partials(const FusedArgs &a, double *sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const CT *kzs = (const CT *)a.kz;
  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int c = a.it_box[item], lo = a.it_lo[item], hi = a.it_hi[item];
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (a.it_kind[item] == 0) {
      for (int idx = lo + tid; idx < hi; idx += FT) {
        const int64_t i = a.cl_nodes[idx];
        double Z[3][6];
        z_of(a.g, c, a.xyz, a.fixdof, i, Z);
        const double w = a.wt ? a.wt[3 * i] : 1.0;
        const double r0 = w * a.r[3 * i], r1 = w * a.r[3 * i + 1], r2 = w * a.r[3 * i + 2];
#pragma unroll
        for (int m = 0; m < 6; m++) v[m] += Z[0][m] * r0 + Z[1][m] * r1 + Z[2][m] * r2;
      }
    } else {
      for (int idx = lo + tid; idx < hi; idx += FT) {
        const int64_t i = a.ent_node[idx];
        const CT *kz = kzs + idx;
        const double y0 = a.u[3 * i], y1 = a.u[3 * i + 1], y2 = a.u[3 * i + 2];
#pragma unroll
        for (int m = 0; m < 6; m++)
          v[m] -= (double)__ldcs(kz + m * a.nent) * y0 + (double)__ldcs(kz + (6 + m) * a.nent) * y1 +
                  (double)__ldcs(kz + (12 + m) * a.nent) * y2;
      }
    }
#pragma unroll
    for (int m = 0; m < 6; m++) {
      const double s = warp_sum(v[m]);
      if (lane == 0) sh[m * FW + warp] = s;
    }
    __syncthreads();
    if (tid < 6) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < FW; w++) s += sh[tid * FW + w];
      a.item_part[6 * (int64_t)item + tid] = s;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void coarse_finish(const FusedArgs &a) {
  const int64_t stride = (int64_t)gridDim.x * FT;
  for (int64_t q = blockIdx.x * (int64_t)FT + threadIdx.x; q < a.n6; q += stride) {
    const int c = (int)(q / 6), m = (int)(q - 6 * (int64_t)c);
    double s = 0.0;
    for (int item = a.box_item_ptr[c]; item < a.box_item_ptr[c + 1]; item++) s += __ldcg(a.item_part + 6 * (int64_t)item + m);
    a.rhs[q] = s;
  }
}

// lam4[part][row] = sum over the part-th quarter of the columns of E^-1[row][col] rhs[col]: one warp per item
template <typename CT>
__device__ __forceinline__ void coarse_gemv(const FusedArgs &a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const CT *E = (const CT *)a.einv;
  const int64_t n6 = a.n6, nw = (int64_t)gridDim.x * FW;
  for (int64_t q = blockIdx.x * (int64_t)FW + warp; q < 4 * n6; q += nw) {
    const int64_t row = q >> 2;
    const int part = (int)(q & 3);
    const int64_t c0 = part * n6 / 4, c1 = (part + 1) * n6 / 4;
    const CT *e = E + row * a.einv_ld;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;            // four loads in flight per lane, fixed interleave
    int64_t col = c0 + lane;
    for (; col + 96 < c1; col += 128) {
      s0 += (double)__ldcs(e + col) * a.rhs[col];
      s1 += (double)__ldcs(e + col + 32) * a.rhs[col + 32];
      s2 += (double)__ldcs(e + col + 64) * a.rhs[col + 64];
      s3 += (double)__ldcs(e + col + 96) * a.rhs[col + 96];
    }
    for (; col < c1; col += 32) s0 += (double)__ldcs(e + col) * a.rhs[col];
    double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) a.lam4[part * n6 + row] = s;
  }
}

__device__ __forceinline__ void coarse_expand(const FusedArgs &a) {
  const int64_t stride = (int64_t)gridDim.x * FT, n6 = a.n6;
  // two nodes per thread and trip: their dependent load chains (box -> coefficients) overlap
  for (int64_t i0 = blockIdx.x * (int64_t)FT + threadIdx.x; i0 < a.nn; i0 += 2 * stride) {
    const int64_t i1 = i0 + stride;
    const bool two = i1 < a.nn;
    const int32_t c0 = a.cid[i0], c1 = two ? a.cid[i1] : c0;
    double l0[6], l1[6];
#pragma unroll
    for (int m = 0; m < 6; m++) {
      const int64_t q0 = 6 * (int64_t)c0 + m, q1 = 6 * (int64_t)c1 + m;
      l0[m] = ((a.lam4[q0] + a.lam4[n6 + q0]) + a.lam4[2 * n6 + q0]) + a.lam4[3 * n6 + q0];
      l1[m] = ((a.lam4[q1] + a.lam4[n6 + q1]) + a.lam4[2 * n6 + q1]) + a.lam4[3 * n6 + q1];
    }
    double u0[3] = {a.u[3 * i0], a.u[3 * i0 + 1], a.u[3 * i0 + 2]}, u1[3] = {0, 0, 0};
    if (two) { u1[0] = a.u[3 * i1]; u1[1] = a.u[3 * i1 + 1]; u1[2] = a.u[3 * i1 + 2]; }
    double Z[3][6];
    z_of(a.g, c0, a.xyz, a.fixdof, i0, Z);
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double sacc = u0[r];
#pragma unroll
      for (int m = 0; m < 6; m++) sacc += Z[r][m] * l0[m];
      a.u[3 * i0 + r] = sacc;
    }
    if (two) {
      z_of(a.g, c1, a.xyz, a.fixdof, i1, Z);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double sacc = u1[r];
#pragma unroll
        for (int m = 0; m < 6; m++) sacc += Z[r][m] * l1[m];
        a.u[3 * i1 + r] = sacc;
      }
    }
  }
}

// ---- product: w = K u over the worker's slices; leaves this thread's share of w.u and r.u ------------------
__device__ __forceinline__ void spmv_phase(const FusedArgs &a, double (*spart)[FW][3][32], double &dsum, double &rsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int split = a.split, wpb = FW / split;
  const int wk = blockIdx.x * wpb + warp / split, sw = warp % split;
  const int barid = 1 + warp / split, nthr = 32 * split;
  const double *u = a.u;
  int buf = 0;
  dsum = 0.0;
  rsum = 0.0;
  // Slices are dealt to the workers round by round: in round j worker wk takes slice j*nwk + (wk + 37 j) mod nwk.
  // All workers then stream from one moving window of the matrix (as a grid of one block per slice would), which
  // keeps the DRAM pages hot; the rotation by 37 moves a worker through the sorted sigma-windows of the SELL
  // layout so that no worker collects only wide (or only narrow) slices.  The mapping is fixed: bit-reproducible.
  const int nwk = gridDim.x * wpb;
  const int64_t nsl = a.nslices;
  for (int64_t j = 0; j * nwk < nsl; j++, buf ^= 1) {
    const int64_t s = j * nwk + (int64_t)((wk + 37 * j) % nwk);
    if (s >= nsl) break;                      // last, partial round (uniform over the warps of a worker)
    const int32_t k0 = a.slice_ptr[s], k1 = a.slice_ptr[s + 1];
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    const int32_t *ci = a.colidx + (int64_t)(k0 + sw) * SELL_C + lane;
    const double *v = a.vals + (int64_t)(k0 + sw) * 9 * SELL_C + lane;
    const int64_t cstep = (int64_t)split * SELL_C, vstep = (int64_t)split * 9 * SELL_C;
#pragma unroll 4
    for (int32_t k = k0 + sw; k < k1; k += split, ci += cstep, v += vstep) {
      const int64_t c3 = 3 * (int64_t)__ldcs(ci);
      const double a0 = __ldcs(v), a1 = __ldcs(v + SELL_C), a2 = __ldcs(v + 2 * SELL_C);
      const double a3 = __ldcs(v + 3 * SELL_C), a4 = __ldcs(v + 4 * SELL_C), a5 = __ldcs(v + 5 * SELL_C);
      const double a6 = __ldcs(v + 6 * SELL_C), a7 = __ldcs(v + 7 * SELL_C), a8 = __ldcs(v + 8 * SELL_C);
      const double x0 = u[c3], x1 = u[c3 + 1], x2 = u[c3 + 2];
      y0 += a0 * x0 + a1 * x1 + a2 * x2;
      y1 += a3 * x0 + a4 * x1 + a5 * x2;
      y2 += a6 * x0 + a7 * x1 + a8 * x2;
    }
    if (split > 1) {
      // the worker's warps add their partial rows in warp order (fixed); two staging buffers, one barrier per slice
      if (sw > 0) {
        spart[buf][warp][0][lane] = y0;
        spart[buf][warp][1][lane] = y1;
        spart[buf][warp][2][lane] = y2;
      }
      named_bar(barid, nthr);
      if (sw == 0)
        for (int q = 1; q < split; q++) {
          y0 += spart[buf][warp + q][0][lane];
          y1 += spart[buf][warp + q][1][lane];
          y2 += spart[buf][warp + q][2][lane];
        }
    }
    if (sw == 0) {
      const int32_t row = a.slot_node[(int64_t)s * SELL_C + lane];
      if (row >= 0) {
        const int64_t r3 = 3 * (int64_t)row;
        a.w[r3] = y0;
        a.w[r3 + 1] = y1;
        a.w[r3 + 2] = y2;
        const double u0 = u[r3], u1 = u[r3 + 1], u2 = u[r3 + 2];
        dsum += y0 * u0 + y1 * u1 + y2 * u2;
        if (a.defl) rsum += (a.wt ? a.wt[r3] : 1.0) * (a.r[r3] * u0 + a.r[r3 + 1] * u1 + a.r[r3 + 2] * u2);
      }
    }
  }
}

#define FCVM_STAMP(k)                \
  if (timer) {                       \
    const unsigned long long t = gtime(); \
    acc##k += t - t_prev;            \
    t_prev = t;                      \
  }

template <typename CT>
__global__ void __launch_bounds__(FT, FUSED_BLOCKS_PER_SM) k_pcg_fused(FusedArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[6 * FW];
  __shared__ double spart[2][FW][3][32];
  const int tid = threadIdx.x, G = gridDim.x;
  const bool timer = (blockIdx.x == 0 && tid == 0 && a.phase_ns != nullptr);
  unsigned long long t_prev = timer ? gtime() : 0ull, acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0;
  double rr = a.sc[S_RR], gam = a.sc[S_GAMMA], gam_prev = 0.0, alpha_prev = 0.0;
  const double thr = a.sc[S_THR];
  int it = 0, status = PCG_RUNNING;
  for (;;) {
    if (a.defl) {
      coarse_partials<CT>(a, sh);
      grid.sync();
      FCVM_STAMP(1)
      coarse_finish(a);
      grid.sync();
      FCVM_STAMP(2)
      coarse_gemv<CT>(a);
      grid.sync();
      FCVM_STAMP(3)
      coarse_expand(a);
      grid.sync();
      FCVM_STAMP(4)
    }
    double d[2];
    spmv_phase(a, spart, d[0], d[1]);
    block_partial<2>(d, a.part, G, sh);
    grid.sync();
    FCVM_STAMP(5)
    const double delta = grid_sum(a.part, G, sh);
    if (a.defl) gam = grid_sum(a.part + G, G, sh);
    // every block holds the same bits in rr, gam, delta: the exits below are taken by all of them together
    if (rr <= thr) { status = PCG_CONVERGED; break; }
    if (it >= a.max_iter) { status = PCG_MAXITER; break; }
    double beta = 0.0, den = delta;
    if (it > 0) {
      beta = gam / gam_prev;
      den = delta - beta * gam / alpha_prev;
    }
    if (!(gam > 0.0) || !(den > 0.0)) { status = PCG_BREAKDOWN; break; }   // operator or preconditioner not positive definite
    const double alpha = gam / den;
    double v[2] = {0.0, 0.0};
    {
      const int64_t stride = (int64_t)G * FT;
      for (int64_t n = blockIdx.x * (int64_t)FT + tid; n < a.nn; n += stride) {
        const int64_t dd = 3 * n;
        double rn[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double pc = a.u[dd + c] + beta * a.p[dd + c];
          const double sn = a.w[dd + c] + beta * a.s[dd + c];
          a.p[dd + c] = pc;
          a.s[dd + c] = sn;
          a.x[dd + c] += alpha * pc;
          rn[c] = a.r[dd + c] - alpha * sn;
          a.r[dd + c] = rn[c];
        }
        const double *m = a.minv + 9 * n;
        const double z0 = m[0] * rn[0] + m[1] * rn[1] + m[2] * rn[2];
        const double z1 = m[3] * rn[0] + m[4] * rn[1] + m[5] * rn[2];
        const double z2 = m[6] * rn[0] + m[7] * rn[1] + m[8] * rn[2];
        a.u[dd] = z0;
        a.u[dd + 1] = z1;
        a.u[dd + 2] = z2;
        const double wt = a.wt ? a.wt[dd] : 1.0;
        v[0] += wt * (rn[0] * z0 + rn[1] * z1 + rn[2] * z2);
        v[1] += wt * (rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]);
      }
    }
    block_partial<2>(v, a.part + 2 * (int64_t)G, G, sh);
    grid.sync();
    FCVM_STAMP(0)
    gam_prev = gam;
    alpha_prev = alpha;
    if (!a.defl) gam = grid_sum(a.part + 2 * (int64_t)G, G, sh);    // deflated: r.u follows the coarse correction
    rr = grid_sum(a.part + 3 * (int64_t)G, G, sh);
    it++;
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.sc[S_ITERS] = (double)it;
    a.sc[S_RR] = a.sc[S_RR + 1] = rr;
    a.sc[S_STATUS] = (double)status;
    if (a.phase_ns) {
      a.phase_ns[0] += acc0; a.phase_ns[1] += acc1; a.phase_ns[2] += acc2;
      a.phase_ns[3] += acc3; a.phase_ns[4] += acc4; a.phase_ns[5] += acc5;
      a.phase_ns[8] += (unsigned long long)it;
    }
  }
}

__global__ void k_to_float(int64_t n, const double *__restrict__ src, float *__restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

// rows of n doubles -> rows of ld floats, padding zero
__global__ void k_to_float_rows(int64_t n, int64_t ld, const double *__restrict__ src, float *__restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * ld) return;
  const int64_t r = i / ld, q = i - r * ld;
  dst[i] = q < n ? (float)src[r * n + q] : 0.0f;
}

template <typename T>
int realloc_dev(T **p, int64_t n) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  FCVM_CUDA(cudaMalloc((void **)p, sizeof(T) * (size_t)std::max<int64_t>(n, 1)));
  return FCVM_OK;
}

}  // namespace

namespace fcvm {

// K Z and E^-1 are streamed in single precision by default (they only shape the preconditioner);
// FCVM_COARSE_FP64=1 keeps the double-precision copies (comparison runs)
bool coarse_fp32() {
  static const bool v = !(getenv("FCVM_COARSE_FP64") && atoi(getenv("FCVM_COARSE_FP64")) != 0);
  return v;
}

// Opt-in (FCVM_PCG_FUSED=1), an EXPERIMENT kept for the record: measured on B200 at 1M elements the persistent
// kernel is correct (the whole GPU suite passed with it) but slower than one launch per phase -- 16 resident warps
// per SM at the 128 registers its largest phase needs do not keep enough loads in flight (product 0.53 ms against
// 0.46 ms for the dedicated kernel at 48 warps per SM, coarse phases 1.5-2.5x); profiles/README.md has the numbers.
// It applies the assembled operator and predates the matrix-free product and the bulk-copy kernels of the
// production path.
bool pcg_fused_enabled(const fcvm_ctx *c) {
  static const bool on = getenv("FCVM_PCG_FUSED") && atoi(getenv("FCVM_PCG_FUSED")) != 0;
  return on && c->world == 1;
}

// Work lists of the coarse phases (fixed chunks of every box's node list and entry list) -- structure only,
// built once per deflation structure from the host copies of the two pointer arrays.
int fused_build_items(fcvm_ctx *c, const std::vector<int32_t> &cl_ptr, const std::vector<int32_t> &ent_ptr) {
  const int64_t ncl = c->ncl;
  std::vector<int32_t> box, lo, hi, bptr((size_t)ncl + 1, 0);
  std::vector<uint8_t> kind;
  for (int64_t cidx = 0; cidx < ncl; cidx++) {
    bptr[(size_t)cidx] = (int32_t)box.size();
    for (int k = 0; k < 2; k++) {
      const int32_t b0 = k == 0 ? cl_ptr[(size_t)cidx] : ent_ptr[(size_t)cidx];
      const int32_t b1 = k == 0 ? cl_ptr[(size_t)cidx + 1] : ent_ptr[(size_t)cidx + 1];
      const int32_t len = b1 - b0;
      if (len <= 0) continue;
      const int32_t nch = (len + COARSE_CHUNK - 1) / COARSE_CHUNK;
      for (int32_t q = 0; q < nch; q++) {
        box.push_back((int32_t)cidx);
        kind.push_back((uint8_t)k);
        lo.push_back(b0 + (int32_t)((int64_t)len * q / nch));
        hi.push_back(b0 + (int32_t)((int64_t)len * (q + 1) / nch));
      }
    }
  }
  bptr[(size_t)ncl] = (int32_t)box.size();
  const int64_t ni = (int64_t)box.size();
  c->n_items = ni;
  FCVM_TRY(realloc_dev(&c->it_box, ni)); FCVM_TRY(realloc_dev(&c->it_lo, ni)); FCVM_TRY(realloc_dev(&c->it_hi, ni));
  FCVM_TRY(realloc_dev(&c->it_kind, ni)); FCVM_TRY(realloc_dev(&c->box_item_ptr, ncl + 1));
  FCVM_TRY(realloc_dev(&c->item_part, 6 * ni)); FCVM_TRY(realloc_dev(&c->lam4, 4 * 6 * ncl));
  if (ni > 0) {
    FCVM_CUDA(cudaMemcpy(c->it_box, box.data(), sizeof(int32_t) * ni, cudaMemcpyHostToDevice));
    FCVM_CUDA(cudaMemcpy(c->it_lo, lo.data(), sizeof(int32_t) * ni, cudaMemcpyHostToDevice));
    FCVM_CUDA(cudaMemcpy(c->it_hi, hi.data(), sizeof(int32_t) * ni, cudaMemcpyHostToDevice));
    FCVM_CUDA(cudaMemcpy(c->it_kind, kind.data(), (size_t)ni, cudaMemcpyHostToDevice));
  }
  FCVM_CUDA(cudaMemcpy(c->box_item_ptr, bptr.data(), sizeof(int32_t) * (ncl + 1), cudaMemcpyHostToDevice));
  return FCVM_OK;
}

// single-precision copies of K Z and E^-1 for the fused kernel (values change with every assembly)
int fused_refresh_coarse(fcvm_ctx *c) {
  if (!coarse_fp32()) return FCVM_OK;
  const int64_t nkz = 18 * c->nent, n6 = 6 * c->ncl, ne = n6 * c->einv_ld;
  if (!c->kz32) FCVM_TRY(realloc_dev(&c->kz32, nkz));
  if (!c->einv32) FCVM_TRY(realloc_dev(&c->einv32, ne));
  k_to_float<<<grid_for(nkz, 256), 256, 0, c->stream>>>(nkz, c->kz_val, c->kz32);
  k_to_float_rows<<<grid_for(ne, 256), 256, 0, c->stream>>>(n6, c->einv_ld, c->dEinv, c->einv32);
  c->launches += 2;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

void fused_free(fcvm_ctx *c) {
  cudaFree(c->it_box); cudaFree(c->it_lo); cudaFree(c->it_hi); cudaFree(c->it_kind); cudaFree(c->box_item_ptr);
  cudaFree(c->item_part); cudaFree(c->kz32); cudaFree(c->einv32); cudaFree(c->lam4);
  c->it_box = c->it_lo = c->it_hi = c->box_item_ptr = nullptr;
  c->it_kind = nullptr;
  c->item_part = c->lam4 = nullptr;
  c->kz32 = c->einv32 = nullptr;
  c->n_items = 0;
}

void fused_free_mesh(fcvm_ctx *c) {
  cudaFree(c->fused_part);
  c->fused_part = nullptr;
  c->wk_grid = c->wk_split = 0;
}

static int fused_prepare(fcvm_ctx *c) {
  if (c->fused_grid == 0) {
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    FCVM_CUDA(cudaGetDevice(&dev));
    FCVM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FCVM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    FCVM_CHECK(coop, FCVM_E_CUDA, "fused PCG: the device does not support cooperative launches");
    FCVM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_fused<float>, FT, 0));
    int per_sm64 = 0;
    FCVM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm64, k_pcg_fused<double>, FT, 0));
    per_sm = std::min(std::min(per_sm, per_sm64), FUSED_BLOCKS_PER_SM);
    FCVM_CHECK(per_sm >= 1, FCVM_E_CUDA, "fused PCG: kernel does not fit an SM");
    c->fused_grid = per_sm * sms;
  }
  const int G = c->fused_grid;
  if (c->wk_grid != G) {
    // 1..8 warps per SpMV worker so that every worker still walks several slices when a rank holds few rows
    // (slices are dealt to the workers round by round inside the kernel)
    int split = 8;
    for (int s = 1; s <= 8; s *= 2)
      if (c->nslices >= 6 * (int64_t)G * (FW / s)) { split = s; break; }
    if (getenv("FCVM_FUSED_SPLIT")) split = std::max(1, std::min(8, atoi(getenv("FCVM_FUSED_SPLIT"))));   // experiments
    FCVM_TRY(realloc_dev(&c->fused_part, 4 * (int64_t)G));
    c->wk_grid = G;
    c->wk_split = split;
  }
  if (!c->phase_ns) {
    FCVM_CUDA(cudaMalloc((void **)&c->phase_ns, sizeof(unsigned long long) * 16));
    FCVM_CUDA(cudaMemset(c->phase_ns, 0, sizeof(unsigned long long) * 16));
  }
  return FCVM_OK;
}

// The iteration loop of fcvm_pcg_solve after its set-up (r, y = D^-1 r, p = s = 0, scalars): launches the
// persistent kernel once; the caller reads the scalars back.
int pcg_fused_loop(fcvm_ctx *c, double *x, int max_iter) {
  FCVM_TRY(fused_prepare(c));
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.nn = c->nn;
  a.nslices = c->nslices;
  a.max_iter = max_iter;
  a.defl = c->defl_ready ? 1 : 0;
  a.split = c->wk_split;
  a.slice_ptr = c->slice_ptr; a.slot_node = c->slot_node; a.colidx = c->colidx;
  a.vals = c->vals; a.minv = c->minv; a.wt = c->dof_weight;
  a.x = x; a.r = c->pcg_r; a.u = c->pcg_z; a.p = c->pcg_p; a.s = c->pcg_s; a.w = c->pcg_q;
  a.part = c->fused_part;
  a.sc = c->red_out;
  a.phase_ns = c->phase_ns;
  const bool f32 = coarse_fp32();
  if (a.defl) {
    a.g = grid_of(c);
    a.ncl = c->ncl; a.n6 = 6 * c->ncl; a.nent = c->nent;
    a.einv_ld = f32 ? c->einv_ld : 6 * c->ncl;
    a.n_items = (int)c->n_items;
    a.cid = c->d_cid; a.cl_nodes = c->cl_nodes; a.ent_node = c->ent;
    a.it_box = c->it_box; a.it_lo = c->it_lo; a.it_hi = c->it_hi; a.box_item_ptr = c->box_item_ptr; a.it_kind = c->it_kind;
    a.xyz = c->xyz; a.fixdof = (const double *)c->buf[FCVM_BUF_FIXDOF];
    a.kz = f32 ? (const void *)c->kz32 : (const void *)c->kz_val;
    a.einv = f32 ? (const void *)c->einv32 : (const void *)c->dEinv;
    a.item_part = c->item_part; a.rhs = c->d_rhs; a.lam4 = c->lam4;
  }
  void *params[] = {(void *)&a};
  const void *fn = f32 ? (const void *)k_pcg_fused<float> : (const void *)k_pcg_fused<double>;
  FCVM_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)c->fused_grid), dim3(FT), params, 0, c->stream));
  c->launches++;
  return FCVM_OK;
}

}  // namespace fcvm

// Device time of the fused kernel per phase since the last reset, in ms:
// [0] step, [1] coarse partials, [2] coarse finish, [3] coarse product, [4] expand, [5] product (SpMV);
// iterations = PCG iterations these cover.
extern "C" int fcvm_pcg_phase_times(fcvm_ctx *c, double *ms6, int64_t *iterations, int reset) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  unsigned long long h[16] = {0};
  if (c->phase_ns) {
    FCVM_CUDA(cudaStreamSynchronize(c->stream));
    FCVM_CUDA(cudaMemcpy(h, c->phase_ns, sizeof(h), cudaMemcpyDeviceToHost));
    if (reset) FCVM_CUDA(cudaMemset(c->phase_ns, 0, sizeof(h)));
  }
  if (ms6)
    for (int i = 0; i < 6; i++) ms6[i] = (double)h[i] * 1e-6;
  if (iterations) *iterations = (int64_t)h[8];
  return FCVM_OK;
}
