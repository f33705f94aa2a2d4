// Block-SELL SpMV and the on-device preconditioned conjugate-gradient solve that replaces
// factor = cholesky(gsm); x = factor(b)   (fcVM.py:1121-1135, 1264-1278, 1369-1406).
//
// Single-reduction (Chronopoulos-Gear) preconditioned CG: per iteration one product w = K u -- which also
// leaves the partial sums of w.u (and r.u) -- and one vector kernel (k_pcg_step_bulk: a bulk-copy stream; k_pcg_step
// for unaligned or tiny problems).  The product is matrix-free (fcvm_matfree.cu) while the matrix is calcGSM's
// elastic one, else the block-SELL SpMV below.  The loop runs without host round trips: step lengths live in
// device scalars, every dot product is a fixed-shape reduction (bit-reproducible), and a device flag turns the
// remaining launches of a batch into no-ops once the tolerance is met; the host only polls every CHECK_EVERY
// iterations.  Preconditioner: block-Jacobi, plus the rigid-body-mode deflation level of fcvm_deflation.cu when
// switched on.  Start vector: zero, the caller's, or the projection onto the last two solutions (use_x0 = 2).  On a
// partitioned mesh the ranks exchange through mapped peer memory (fcvm_p2p.cu); without it the interface
// all-reduce of w overlaps the interior part of the SpMV on a communication stream (NCCL).
#include <cmath>

#include "fcvm_common.cuh"
#include "fcvm_pcg.cuh"
#include "fcvm_reduce.cuh"

using namespace fcvm;

extern "C" int fcvm_comm_allreduce_sum(fcvm_ctx *c, double *dev, int64_t n);
namespace fcvm {
int fcvm_comm_allreduce_oop(fcvm_ctx *c, const double *send, double *recv, int64_t n);
}

namespace {

constexpr int CHECK_EVERY = 16;

// y = K x on the block-SELL matrix.  One block per 32-row slice, SPMV_SPLIT warps per block: warp w
// walks columns k0+w, k0+w+SPLIT, ... of the slice, so the column index and each of the nine block
// entries are read as coalesced 128/256-byte lines; warp 0 then adds the partial rows in warp order
// (fixed, so bit-reproducible).  Matrix data is streamed (evict-first) to keep x resident in L2.
// Four warps per slice keep enough loads in flight even when a rank holds only a few thousand
// slices (one warp per slice measured 3.5 TB/s at 5k slices and 5.9 TB/s at 43k; this layout 5.2 and
// 6.5).  With `dot_part` warp 0 also leaves the slice's share of y.x (the w.u of the PCG iteration,
// summed by k_dot_finish).  `slices` (optional) lists the slices to process: the boundary /
// interior split of the overlapped interface exchange.
constexpr int SPMV_SPLIT = 4;
__global__ void __launch_bounds__(32 * SPMV_SPLIT)
k_spmv_sell(int64_t nlist, const int32_t *__restrict__ slices, const int32_t *__restrict__ slice_ptr,
                  const int32_t *__restrict__ slot_node, const int32_t *__restrict__ colidx,
                  const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
                  const double *__restrict__ sc, int rr_slot, double *dot_part, const double *__restrict__ rvec,
                  const double *__restrict__ wt, double *dot_part2) {
  if (sc && (sc[S_ITERS] >= 0.0 || sc[rr_slot] <= sc[S_THR])) return;
  __shared__ double part[SPMV_SPLIT - 1][3][32];
  const int64_t s = slices ? slices[blockIdx.x] : (int64_t)blockIdx.x;
  (void)nlist;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int32_t k0 = slice_ptr[s], k1 = slice_ptr[s + 1];
  double y0 = 0.0, y1 = 0.0, y2 = 0.0;
  const int32_t *ci = colidx + (int64_t)(k0 + w) * SELL_C + lane;
  const double *v = vals + (int64_t)(k0 + w) * 9 * SELL_C + lane;
#pragma unroll 4
  for (int32_t k = k0 + w; k < k1; k += SPMV_SPLIT, ci += SPMV_SPLIT * SELL_C, v += SPMV_SPLIT * 9 * SELL_C) {
    const int64_t c3 = 3 * (int64_t)__ldcs(ci);
    const double a0 = __ldcs(v), a1 = __ldcs(v + SELL_C), a2 = __ldcs(v + 2 * SELL_C);
    const double a3 = __ldcs(v + 3 * SELL_C), a4 = __ldcs(v + 4 * SELL_C), a5 = __ldcs(v + 5 * SELL_C);
    const double a6 = __ldcs(v + 6 * SELL_C), a7 = __ldcs(v + 7 * SELL_C), a8 = __ldcs(v + 8 * SELL_C);
    const double x0 = x[c3], x1 = x[c3 + 1], x2 = x[c3 + 2];
    y0 += a0 * x0 + a1 * x1 + a2 * x2;
    y1 += a3 * x0 + a4 * x1 + a5 * x2;
    y2 += a6 * x0 + a7 * x1 + a8 * x2;
  }
  if (w > 0) {
    part[w - 1][0][lane] = y0;
    part[w - 1][1][lane] = y1;
    part[w - 1][2][lane] = y2;
  }
  __syncthreads();
  if (w > 0) return;
#pragma unroll
  for (int q = 0; q < SPMV_SPLIT - 1; q++) {
    y0 += part[q][0][lane];
    y1 += part[q][1][lane];
    y2 += part[q][2][lane];
  }
  const int32_t row = slot_node[s * SELL_C + lane];
  double dsum = 0.0, rsum = 0.0;
  if (row >= 0) {
    const int64_t r3 = 3 * (int64_t)row;
    y[r3] = y0;
    y[r3 + 1] = y1;
    y[r3 + 2] = y2;
    if (dot_part) dsum = y0 * x[r3] + y1 * x[r3 + 1] + y2 * x[r3 + 2];
    if (dot_part2) rsum = (wt ? wt[r3] : 1.0) * (rvec[r3] * x[r3] + rvec[r3 + 1] * x[r3 + 1] + rvec[r3 + 2] * x[r3 + 2]);
  }
  if (dot_part) {
    dsum = warp_sum(dsum);
    if (lane == 0) dot_part[s] = dsum;
  }
  if (dot_part2) {
    // r.u with the deflated preconditioner (u is only complete after the coarse correction); on a
    // partitioned mesh shared rows count once (weight 1/multiplicity)
    rsum = warp_sum(rsum);
    if (lane == 0) dot_part2[s] = rsum;
  }
}

// delta = sum of the per-slice partials of w.u, in slice order (one block, fixed shape), and with
// `part2` gamma = r.u likewise; with a communicator the per-rank sums go to the tail buffer of the
// scalar all-reduce instead
__global__ void __launch_bounds__(1024)
k_dot_finish(int64_t n, const double *__restrict__ part, const double *__restrict__ part2, double *sc, int delta_slot,
             int gamma_slot, double *tail) {
  if (sc[S_ITERS] >= 0.0) return;
  // four independent partial sums per thread (fixed interleave) keep the loads in flight; the order of the
  // additions is fixed by n alone
  double ta[4] = {0, 0, 0, 0}, ga[4] = {0, 0, 0, 0};
  for (int64_t i0 = threadIdx.x; i0 < n; i0 += 4 * 1024) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int64_t i = i0 + u * 1024;
      if (i < n) {
        ta[u] += part[i];
        if (part2) ga[u] += part2[i];
      }
    }
  }
  double t = (ta[0] + ta[1]) + (ta[2] + ta[3]), g = (ga[0] + ga[1]) + (ga[2] + ga[3]);
  __shared__ double sm[2][32];
  t = warp_sum(t);
  g = warp_sum(g);
  if ((threadIdx.x & 31) == 0) {
    sm[0][threadIdx.x >> 5] = t;
    sm[1][threadIdx.x >> 5] = g;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0, gam = 0.0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      tot += sm[0][w];
      gam += sm[1][w];
    }
    sc[delta_slot] = tot;
    if (part2 && gamma_slot >= 0) sc[gamma_slot] = gam;
    if (tail) {
      tail[0] = part2 ? gam : sc[L_RU];
      tail[1] = sc[L_RR];
      tail[2] = tot;
    }
  }
}

__device__ __forceinline__ void apply_minv(const double *__restrict__ m, double r0, double r1, double r2, double &z0,
                                           double &z1, double &z2) {
  z0 = m[0] * r0 + m[1] * r1 + m[2] * r2;
  z1 = m[3] * r0 + m[4] * r1 + m[5] * r2;
  z2 = m[6] * r0 + m[7] * r1 + m[8] * r2;
}

// r = b - q (q = K x0, or absent), u = M^-1 r, p = s = 0 ; sums r.u, b.b, r.r
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_init(int64_t nn, const double *__restrict__ b, const double *__restrict__ q, const double *__restrict__ minv,
           const double *__restrict__ w, double *r, double *u, double *p, double *s, double *red_part,
           unsigned int *counter, double *out, Slots<3> sl) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = 3 * n;
    double r0 = b[d], r1 = b[d + 1], r2 = b[d + 2];
    const double wt = w ? w[d] : 1.0;
    v[1] += wt * (r0 * r0 + r1 * r1 + r2 * r2);
    if (q) { r0 -= q[d]; r1 -= q[d + 1]; r2 -= q[d + 2]; }
    double z0, z1, z2;
    apply_minv(minv + 9 * n, r0, r1, r2, z0, z1, z2);
    r[d] = r0; r[d + 1] = r1; r[d + 2] = r2;
    u[d] = z0; u[d + 1] = z1; u[d + 2] = z2;
    p[d] = 0.0; p[d + 1] = 0.0; p[d + 2] = 0.0;
    s[d] = 0.0; s[d + 1] = 0.0; s[d + 2] = 0.0;
    v[0] += wt * (r0 * z0 + r1 * z1 + r2 * z2);
    v[2] += wt * (r0 * r0 + r1 * r1 + r2 * r2);
  }
  block_reduce_publish<3>(v, red_part, counter, out, sl);      // r.u, b.b, r.r
}

__global__ void k_pcg_scalars(double rtol, double *sc) {
  sc[S_THR] = rtol * rtol * sc[S_BB];
  sc[S_ITERS] = -1.0;
  sc[S_STATUS] = (double)PCG_RUNNING;
  sc[S_ALPHA] = sc[S_ALPHA + 1] = 0.0;
  sc[S_RR + 1] = sc[S_RR];              // both parity slots start from the initial residual
}

// One iteration of the single-reduction (Chronopoulos-Gear) form of preconditioned CG, vector part:
//   beta = gamma_it / gamma_(it-1), alpha = gamma_it / (delta_it - beta gamma_it / alpha_(it-1))
//   p = u + beta p ; s = w + beta s (= K p) ; x += alpha p ; r -= alpha s ; u = M^-1 r
// and the sums r.u (-> gamma_(it+1)) and r.r.  w = K u and delta = w.u come from the SpMV that follows.
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_step(int64_t nn, int it, int defl, const double *__restrict__ w, const double *__restrict__ minv,
           const double *__restrict__ wt_, double *__restrict__ x, double *__restrict__ r, double *__restrict__ u,
           double *__restrict__ p, double *__restrict__ s, double *red_part,
           unsigned int *counter, double *sc, Slots<2> sl) {
  const int cur = it & 1, prv = cur ^ 1;
  if (sc[S_ITERS] >= 0.0) return;                    // converged earlier in this batch (sticky)
  if (sc[S_RR + cur] <= sc[S_THR]) {                 // converged after `it` iterations
    if (blockIdx.x == 0 && threadIdx.x == 0 && sc[S_ITERS] < 0.0) sc[S_ITERS] = (double)it;
    return;
  }
  const double gam = sc[S_GAMMA + cur];
  double beta = 0.0, den = sc[S_DELTA];
  if (it > 0) {
    const double gprev = sc[S_GAMMA + prv], aprev = sc[S_ALPHA + prv];
    beta = gprev != 0.0 ? gam / gprev : 0.0;
    den = sc[S_DELTA] - (aprev != 0.0 ? beta * gam / aprev : 0.0);
  }
  if (!(gam > 0.0) || !(den > 0.0)) {
    // operator or preconditioner not positive definite (a tangent at or past buckling, a floating model):
    // stop here with a sticky flag instead of iterating on to the limit
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      sc[S_STATUS] = (double)PCG_BREAKDOWN;
      sc[S_ITERS] = (double)it;
    }
    return;
  }
  const double alpha = gam / den;
  double v[2] = {0.0, 0.0};
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = 3 * n;
    double rn[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double pc = u[d + c] + beta * p[d + c];
      const double sn = w[d + c] + beta * s[d + c];
      p[d + c] = pc;
      s[d + c] = sn;
      x[d + c] += alpha * pc;
      rn[c] = r[d + c] - alpha * sn;
      r[d + c] = rn[c];
    }
    double z0, z1, z2;
    apply_minv(minv + 9 * n, rn[0], rn[1], rn[2], z0, z1, z2);
    u[d] = z0; u[d + 1] = z1; u[d + 2] = z2;
    const double wt = wt_ ? wt_[d] : 1.0;
    if (!defl) v[0] += wt * (rn[0] * z0 + rn[1] * z1 + rn[2] * z2);   // deflated: r.u follows the coarse correction
    v[1] += wt * (rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) sc[S_ALPHA + cur] = alpha;
  block_reduce_publish<2>(v, red_part, counter, sc, sl);       // gamma_(it+1), rr_(it+1)
}

// ---------------------------------------------------------------------------------------------------------------
// The same vector step as a TMA-style stream: the kernel is a pure pass over eleven nodal arrays, so instead of
// every thread holding its own loads in registers, one elected thread per block moves whole tiles with bulk
// asynchronous copies (cp.async.bulk global -> shared, completion counted on an mbarrier), the block updates the
// tile in place in shared memory, and the five result arrays leave through bulk stores (shared -> global, bulk
// groups).  Two stages per block, two blocks per SM: up to 221 kB of loads in flight per SM with no registers tied
// up, which is what an HBM3e stream needs.  Persistent grid, tiles dealt round-robin: the order of all sums is
// fixed.  Tile = 256 nodes: six 6 kB vector pieces and 18 kB of inverse diagonal blocks.
constexpr int ST_T = 256;
constexpr int ST_STAGES = 2;
struct StepTile {
  double w[3 * ST_T], u[3 * ST_T], p[3 * ST_T], s[3 * ST_T], x[3 * ST_T], r[3 * ST_T], m[9 * ST_T];
};
static_assert(sizeof(StepTile) % 16 == 0, "bulk copies move multiples of 16 bytes");

__global__ void __launch_bounds__(ST_T, 2)
k_pcg_step_bulk(int64_t nn, int it, int defl, const double *__restrict__ w, const double *__restrict__ minv,
                const double *__restrict__ wt_, double *x, double *r, double *u, double *p, double *s, double *red_part,
                unsigned int *counter, double *sc, Slots<2> sl) {
  extern __shared__ __align__(128) unsigned char st_raw[];
  StepTile *tiles = (StepTile *)st_raw;
  __shared__ uint64_t full[ST_STAGES];
  const int cur = it & 1, prv = cur ^ 1, tid = threadIdx.x;
  if (sc[S_ITERS] >= 0.0) return;                    // converged earlier in this batch (sticky)
  if (sc[S_RR + cur] <= sc[S_THR]) {                 // converged after `it` iterations
    if (blockIdx.x == 0 && tid == 0 && sc[S_ITERS] < 0.0) sc[S_ITERS] = (double)it;
    return;
  }
  const double gam = sc[S_GAMMA + cur];
  double beta = 0.0, den = sc[S_DELTA];
  if (it > 0) {
    const double gprev = sc[S_GAMMA + prv], aprev = sc[S_ALPHA + prv];
    beta = gprev != 0.0 ? gam / gprev : 0.0;
    den = sc[S_DELTA] - (aprev != 0.0 ? beta * gam / aprev : 0.0);
  }
  if (!(gam > 0.0) || !(den > 0.0)) {                // see k_pcg_step
    if (blockIdx.x == 0 && tid == 0) {
      sc[S_STATUS] = (double)PCG_BREAKDOWN;
      sc[S_ITERS] = (double)it;
    }
    return;
  }
  const double alpha = gam / den;
  if (tid == 0) {
    for (int q = 0; q < ST_STAGES; q++) mbar_init(&full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = nn / ST_T;                  // full tiles; the remainder is walked with plain accesses below
  auto issue = [&](int64_t t, int stage) {           // thread 0 only
    StepTile &T = tiles[stage];
    const int64_t d0 = 3 * t * ST_T;
    mbar_expect_tx(&full[stage], (uint32_t)sizeof(StepTile));
    bulk_load(T.w, w + d0, sizeof(T.w), &full[stage]);
    bulk_load(T.u, u + d0, sizeof(T.u), &full[stage]);
    bulk_load(T.p, p + d0, sizeof(T.p), &full[stage]);
    bulk_load(T.s, s + d0, sizeof(T.s), &full[stage]);
    bulk_load(T.x, x + d0, sizeof(T.x), &full[stage]);
    bulk_load(T.r, r + d0, sizeof(T.r), &full[stage]);
    bulk_load(T.m, minv + 9 * t * ST_T, sizeof(T.m), &full[stage]);
  };
  double v[2] = {0.0, 0.0};
  auto node = [&](double *pw, double *pu, double *pp, double *ps, double *px, double *pr, const double *pm, double wgt) {
    double rn[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double pc = pu[c] + beta * pp[c];
      const double sn = pw[c] + beta * ps[c];
      pp[c] = pc;
      ps[c] = sn;
      px[c] += alpha * pc;
      rn[c] = pr[c] - alpha * sn;
      pr[c] = rn[c];
    }
    const double z0 = pm[0] * rn[0] + pm[1] * rn[1] + pm[2] * rn[2];
    const double z1 = pm[3] * rn[0] + pm[4] * rn[1] + pm[5] * rn[2];
    const double z2 = pm[6] * rn[0] + pm[7] * rn[1] + pm[8] * rn[2];
    pu[0] = z0; pu[1] = z1; pu[2] = z2;
    if (!defl) v[0] += wgt * (rn[0] * z0 + rn[1] * z1 + rn[2] * z2);
    v[1] += wgt * (rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]);
  };
  if (tid == 0 && (int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int k = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, k++) {
    const int stage = k % ST_STAGES;
    if (tid == 0 && t + gridDim.x < ntiles) {
      // the other stage was stored by the previous trip: its bulk stores must have read shared memory
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue(t + gridDim.x, stage ^ 1);
    }
    mbar_wait(&full[stage], (uint32_t)((k / ST_STAGES) & 1));
    StepTile &T = tiles[stage];
    const int64_t n = t * ST_T + tid;
    node(T.w + 3 * tid, T.u + 3 * tid, T.p + 3 * tid, T.s + 3 * tid, T.x + 3 * tid, T.r + 3 * tid, T.m + 9 * tid,
         wt_ ? wt_[3 * n] : 1.0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes before the async-proxy reads
    __syncthreads();
    if (tid == 0) {
      const int64_t d0 = 3 * t * ST_T;
      bulk_store(p + d0, T.p, sizeof(T.p));
      bulk_store(s + d0, T.s, sizeof(T.s));
      bulk_store(x + d0, T.x, sizeof(T.x));
      bulk_store(r + d0, T.r, sizeof(T.r));
      bulk_store(u + d0, T.u, sizeof(T.u));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (blockIdx.x == gridDim.x - 1) {                 // the nodes beyond the last full tile
    const int64_t n = ntiles * ST_T + tid;
    if (n < nn) {
      double mm[9];
#pragma unroll
      for (int q = 0; q < 9; q++) mm[q] = minv[9 * n + q];
      double lw[3] = {w[3 * n], w[3 * n + 1], w[3 * n + 2]};
      node(lw, u + 3 * n, p + 3 * n, s + 3 * n, x + 3 * n, r + 3 * n, mm, wt_ ? wt_[3 * n] : 1.0);
    }
  }
  if (blockIdx.x == 0 && tid == 0) sc[S_ALPHA + cur] = alpha;
  // block partials, then the block that finishes last adds them in block order
  __shared__ double sm[2][ST_T / 32];
  __shared__ bool last;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const double ws = warp_sum(v[i]);
    if (lane == 0) sm[i][warp] = ws;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < ST_T / 32; q++) t += sm[i][q];
      red_part[i * RED_BLOCKS + blockIdx.x] = t;
    }
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    for (int i = warp; i < 2; i += ST_T / 32) {
      double t = 0.0;
      for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&red_part[i * RED_BLOCKS + b]);
      t = warp_sum(t);
      if (lane == 0) sc[sl.s[i]] = t;
    }
    if (tid == 0) *counter = 0u;
  }
}

// multi-GPU: the three per-rank sums ride at the tail of the interface vector (one all-reduce per iteration)
__global__ void k_tail_get(const double *__restrict__ tail, double *sc, int gamma_slot, int rr_slot) {
  if (sc[S_ITERS] >= 0.0) return;
  if (gamma_slot >= 0) sc[gamma_slot] = tail[0];
  if (rr_slot >= 0) sc[rr_slot] = tail[1];
  sc[S_DELTA] = tail[2];
}

}  // namespace

static void spmv_launch(fcvm_ctx *c, const double *x, double *y, const double *sc, int rr_slot, double *dot_part,
                        const int32_t *list = nullptr, int64_t nlist = -1, const double *rvec = nullptr,
                        double *dot_part2 = nullptr, const double *vals = nullptr) {
  const int64_t nb = list ? nlist : c->nslices;
  if (nb <= 0) return;
  k_spmv_sell<<<(unsigned)nb, 32 * SPMV_SPLIT, 0, c->stream>>>(nb, list, c->slice_ptr, c->slot_node, c->colidx,
                                                              vals ? vals : c->vals, x, y, sc, rr_slot, dot_part, rvec,
                                                              c->dof_weight, dot_part2);
}

namespace fcvm {
// the same product with another value array on the context's pattern (geometric stiffness)
int launch_spmv_values(fcvm_ctx *c, const double *vals, const double *x, double *y) {
  ProfScope ps(c, 0);
  spmv_launch(c, x, y, nullptr, 0, nullptr, nullptr, -1, nullptr, nullptr, vals);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}
int launch_spmv(fcvm_ctx *c, const double *x, double *y) {
  ProfScope ps(c, 0);
  spmv_launch(c, x, y, nullptr, 0, nullptr);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}
}  // namespace fcvm

extern "C" int fcvm_spmv(fcvm_ctx *c, const double *x, double *y) {
  FCVM_CHECK(c && c->assembled && x && y, FCVM_E_ARG, "fcvm_spmv: assemble first / null argument");
  FCVM_TRY(launch_spmv(c, x, y));
  return fcvm_interface_sum(c, y);
}

namespace fcvm {
bool matfree_active(const fcvm_ctx *c);
int64_t matfree_parts(const fcvm_ctx *c);
int launch_matfree(fcvm_ctx *c, const double *x, double *y, const double *sc, int rr_slot, int iters_slot, int thr_slot,
                   double *dot_part, const double *rvec, double *dot_part2, double *sc_out, int delta_slot,
                   int gamma_slot);
bool p2p_ready(const fcvm_ctx *c);
int p2p_halo(fcvm_ctx *c, double *v, double *sc, int gamma_slot, int rr_slot, bool with_scalars, bool done_check);
int p2p_check(fcvm_ctx *c);
int interface_sum_on_comm_stream(fcvm_ctx *c, double *v);
int comm_allreduce_on(fcvm_ctx *c, double *dev, int64_t n, cudaStream_t st);
int deflation_correct(fcvm_ctx *c, const double *r, const double *y, const double *base, double *out, const double *sc,
                      int done_slot);
}

// Start vector of a repeated solve with the same matrix (use_x0 == 2): the Galerkin projection of the new right-hand
// side onto the last two solutions d_j, whose images K d_j = b_j are known to the solver tolerance without another
// product:  x0 = sum_j y_j d_j  with  (d_i . b_j) y = (d_i . b).  In the modified Newton iteration successive
// corrections are strongly correlated, so the PCG starts ~0.8 decades closer (measured on the platen sweep: 153 ->
// 131 iterations per solve).  The stopping test stays relative to ||b||.  Returns 1 when x holds a start vector.
static int recycled_start(fcvm_ctx *c, const double *b, double *x, int *have) {
  *have = 0;
  const int m = c->hist_n;
  if (m == 0) return FCVM_OK;
  const int64_t n3 = 3 * c->nn;
  double rhs[2] = {0, 0};
  for (int i = 0; i < m; i++) FCVM_TRY(fcvm_vec_dot(c, n3, c->hist_x[i], b, &rhs[i]));
  double y[2] = {0, 0};
  const double (*G)[2] = c->hist_gram;
  if (m == 1) {
    if (!(G[0][0] > 0.0)) return FCVM_OK;
    y[0] = rhs[0] / G[0][0];
  } else {
    const double g01 = 0.5 * (G[0][1] + G[1][0]), det = G[0][0] * G[1][1] - g01 * g01;
    if (!(G[0][0] > 0.0) || !(G[1][1] > 0.0)) return FCVM_OK;
    if (det > 1e-12 * G[0][0] * G[1][1]) {
      y[0] = (G[1][1] * rhs[0] - g01 * rhs[1]) / det;
      y[1] = (G[0][0] * rhs[1] - g01 * rhs[0]) / det;
    } else {
      y[1] = rhs[1] / G[1][1];                    // the two solutions are parallel: the newer one alone
    }
  }
  if (!std::isfinite(y[0]) || !std::isfinite(y[1])) return FCVM_OK;
  FCVM_TRY(fcvm_vec_axpby(c, n3, y[0], c->hist_x[0], 0.0, x));
  if (m == 2) FCVM_TRY(fcvm_vec_axpby(c, n3, y[1], c->hist_x[1], 1.0, x));
  *have = 1;
  return FCVM_OK;
}

static int remember_solution(fcvm_ctx *c, const double *b, const double *x) {
  const int64_t n3 = 3 * c->nn;
  for (int i = 0; i < 2; i++) {
    if (!c->hist_b[i]) FCVM_TRY(fcvm_vec_alloc(c, n3, &c->hist_b[i]));
    if (!c->hist_x[i]) FCVM_TRY(fcvm_vec_alloc(c, n3, &c->hist_x[i]));
  }
  int slot = c->hist_n;
  if (slot == 2) {                                  // drop the oldest pair, keep its storage
    std::swap(c->hist_b[0], c->hist_b[1]);
    std::swap(c->hist_x[0], c->hist_x[1]);
    c->hist_gram[0][0] = c->hist_gram[1][1];
    slot = 1;
  }
  FCVM_TRY(fcvm_vec_copy(c, n3, b, c->hist_b[slot]));
  FCVM_TRY(fcvm_vec_copy(c, n3, x, c->hist_x[slot]));
  c->hist_n = slot + 1;
  FCVM_TRY(fcvm_vec_dot(c, n3, c->hist_x[slot], c->hist_b[slot], &c->hist_gram[slot][slot]));
  if (slot == 1) {
    FCVM_TRY(fcvm_vec_dot(c, n3, c->hist_x[0], c->hist_b[1], &c->hist_gram[0][1]));
    FCVM_TRY(fcvm_vec_dot(c, n3, c->hist_x[1], c->hist_b[0], &c->hist_gram[1][0]));
  }
  return FCVM_OK;
}

extern "C" int fcvm_pcg_solve(fcvm_ctx *c, const double *b, double *x, double rtol, int max_iter, int use_x0,
                              int *iters, double *relres) {
  FCVM_CHECK(c && c->assembled && b && x, FCVM_E_ARG, "fcvm_pcg_solve: assemble first / null argument");
  FCVM_CHECK(rtol > 0.0 && max_iter > 0, FCVM_E_ARG, "fcvm_pcg_solve: rtol and max_iter must be positive");
  const bool recycle = use_x0 == 2;
  if (recycle) {
    static const bool off = getenv("FCVM_RECYCLE") && atoi(getenv("FCVM_RECYCLE")) == 0;
    int have = 0;
    if (!off) FCVM_TRY(recycled_start(c, b, x, &have));
    use_x0 = have;
  }
  const int64_t nn = c->nn, n3 = 3 * nn;
  // CG ends within ndof iterations in exact arithmetic; a small multiple of that is the most any caller can want
  if ((int64_t)max_iter > 10 * n3 + 100) max_iter = (int)(10 * n3 + 100);
  cudaStream_t st = c->stream;
  double *sc = c->red_out;
  const double *w = c->dof_weight;
  const bool multi = c->world > 1;
  if (!c->spmv_part) {
    FCVM_CUDA(cudaMalloc((void **)&c->spmv_part, sizeof(double) * (size_t)(c->nslices + 8)));
    FCVM_CUDA(cudaMemsetAsync(c->spmv_part, 0, sizeof(double) * (size_t)(c->nslices + 8), st));
  }
  double *r = c->pcg_r, *u = c->pcg_z, *p = c->pcg_p, *wv = c->pcg_q, *s = c->pcg_s;
  const bool mfree = matfree_active(c);
  const bool defl = c->defl_ready;
  double *part2 = defl ? c->spmv_part2 : nullptr;
  const double *q0 = nullptr;
  // K x for the start vectors: the same operator the iteration uses
  auto apply = [&](const double *xin, double *yout) -> int {
    if (!mfree) return fcvm_spmv(c, xin, yout);
    ProfScope ps(c, 0);
    FCVM_TRY(launch_matfree(c, xin, yout, nullptr, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, 0));
    return multi ? p2p_halo(c, yout, sc, -1, -1, false, false) : FCVM_OK;
  };
  if (use_x0) {
    FCVM_TRY(apply(x, wv));
    q0 = wv;
  } else {
    FCVM_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n3, st));
  }
  auto init = [&]() -> int {
    ProfScope ps(c, 3);
    Slots<3> sl = multi ? Slots<3>{{L_RU, L_BB, L_RR}} : Slots<3>{{S_GAMMA, S_BB, S_RR}};
    k_pcg_init<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nn, b, q0, c->minv, w, r, u, p, s, c->red_part, c->red_counter,
                                                   sc, sl);
    c->launches++;
    return FCVM_OK;
  };
  FCVM_TRY(init());
  if (defl) {
    // start from x + Z E^-1 Z^T (b - K x): the residual of the deflated iteration is orthogonal to Z
    FCVM_TRY(deflation_correct(c, r, nullptr, x, x, nullptr, 0));
    FCVM_TRY(apply(x, wv));
    q0 = wv;
    FCVM_TRY(init());
  }
  if (multi) {
    FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_RU, sc + S_GAMMA, 1));
    FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_RR, sc + S_RR, 1));
    FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_BB, sc + S_BB, 1));
  }
  k_pcg_scalars<<<1, 1, 0, st>>>(rtol, sc);
  c->launches++;
  // w = K u and delta = w.u (deflated: first u = y + Z E^-1 (Z^T r - (K Z)^T y), and gamma = r.u with it).
  // With a communicator the slices that hold interface rows are multiplied first; their exchange (pack,
  // all-reduce, unpack) runs on the communication stream while the interior slices are multiplied, and
  // only the three-scalar all-reduce stays on the critical path.
  auto spmv_dot = [&](int it_next) -> int {
    const int flag = S_RR + ((it_next + (multi ? 1 : 0)) & 1);
    // early-out test: r.r of the newest iterate whose global value is known -- the one this product
    // belongs to on one GPU; with a communicator that sum is still in flight, so the one before
    if (defl) {
      ProfScope ps(c, 3);
      FCVM_TRY(deflation_correct(c, r, u, u, u, sc, S_ITERS));
    }
    if (multi && mfree) {
      // partitioned mesh inside one box: per-rank products, then ONE peer-memory exchange that completes the
      // shared rows (neighbours only) and sums the three scalars of the iteration in rank order
      {
        ProfScope ps(c, 0);
        FCVM_TRY(launch_matfree(c, u, wv, sc, flag, S_ITERS, S_THR, c->spmv_part, r, part2, sc, L_WU, defl ? L_RU : -1));
      }
      ProfScope ps(c, 3);
      return p2p_halo(c, wv, sc, (it_next > 0 || defl) ? S_GAMMA + (it_next & 1) : -1,
                      it_next > 0 ? S_RR + (it_next & 1) : -1, true, true);
    } else if (!multi && mfree) {
      // elastic operator recomputed element by element instead of streaming the assembled matrix
      ProfScope ps(c, 0);
      FCVM_TRY(launch_matfree(c, u, wv, sc, flag, S_ITERS, S_THR, c->spmv_part, r, part2, sc, S_DELTA,
                              S_GAMMA + (it_next & 1)));
      return FCVM_OK;                          // delta (and gamma) are published by the gather's last block
    } else if (!multi) {
      ProfScope ps(c, 0);
      spmv_launch(c, u, wv, sc, flag, c->spmv_part, nullptr, -1, r, part2);
    } else {
      {
        ProfScope ps(c, 0);
        spmv_launch(c, u, wv, sc, flag, c->spmv_part, c->bslices, c->n_bslices, r, part2);
      }
      FCVM_CUDA(cudaEventRecord(c->ev_boundary, st));
      FCVM_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_boundary, 0));
      FCVM_TRY(interface_sum_on_comm_stream(c, wv));
      FCVM_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
      {
        ProfScope ps(c, 0);
        spmv_launch(c, u, wv, sc, flag, c->spmv_part, c->islices, c->n_islices, r, part2);
      }
      c->launches++;
    }
    {
      ProfScope ps(c, 3);
      k_dot_finish<<<1, 1024, 0, st>>>(mfree ? matfree_parts(c) : c->nslices, c->spmv_part, part2, sc, multi ? L_WU : S_DELTA,
                                       multi ? -1 : S_GAMMA + (it_next & 1), multi ? c->tail3 : nullptr);
    }
    c->launches += 2;
    if (multi) {
      FCVM_TRY(comm_allreduce_on(c, c->tail3, 3, st));
      FCVM_CUDA(cudaStreamWaitEvent(st, c->ev_halo, 0));
      // gamma / rr of the iterate that the next vector step will test; the first call brings delta (and,
      // deflated, gamma) only: rr of the start vector is already global
      k_tail_get<<<1, 1, 0, st>>>(c->tail3, sc, (it_next > 0 || defl) ? S_GAMMA + (it_next & 1) : -1,
                                  it_next > 0 ? S_RR + (it_next & 1) : -1);
      c->launches++;
    }
    return FCVM_OK;
  };
  // the vector step as a bulk-copy stream (k_pcg_step_bulk) when every array sits on a 16-byte boundary and the mesh
  // has full tiles to stream; FCVM_STEP_BULK=0 keeps the register version
  int step_grid = 0;
  {
    static const bool off = getenv("FCVM_STEP_BULK") && atoi(getenv("FCVM_STEP_BULK")) == 0;
    const uintptr_t all = (uintptr_t)x | (uintptr_t)r | (uintptr_t)u | (uintptr_t)p | (uintptr_t)s | (uintptr_t)wv | (uintptr_t)c->minv;
    if (!off && (all & 15) == 0 && nn >= 4 * ST_T) {
      static int sms = 0;
      if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        FCVM_CUDA(cudaFuncSetAttribute(k_pcg_step_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ST_STAGES * sizeof(StepTile))));
      }
      step_grid = (int)std::min<int64_t>(2 * sms, std::min<int64_t>(RED_BLOCKS, nn / ST_T));
    }
  }
  int it = 0, n_it = -1;
  FCVM_TRY(spmv_dot(0));
  while (n_it < 0 && it < max_iter) {
    const int batch_end = std::min(max_iter, it + CHECK_EVERY);
    for (; it < batch_end; it++) {
      {
        ProfScope ps(c, 3);
        ProfScope ps11(c, 11);
        const int nxt = (it + 1) & 1;
        // deflated: the vector step's own r.u is void (slot L_RU as a sink), gamma comes with the product
        Slots<2> sl = multi ? Slots<2>{{L_RU, L_RR}} : Slots<2>{{defl ? L_RU : S_GAMMA + nxt, S_RR + nxt}};
        if (step_grid > 0)
          k_pcg_step_bulk<<<step_grid, ST_T, ST_STAGES * sizeof(StepTile), st>>>(nn, it, defl ? 1 : 0, wv, c->minv, w, x, r, u, p, s,
                                                                                c->red_part, c->red_counter, sc, sl);
        else
          k_pcg_step<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nn, it, defl ? 1 : 0, wv, c->minv, w, x, r, u, p, s, c->red_part,
                                                         c->red_counter, sc, sl);
        c->launches++;
      }
      FCVM_TRY(spmv_dot(it + 1));
    }
    FCVM_CUDA(cudaGetLastError());
    FCVM_CUDA(cudaMemcpyAsync(c->h_scalars, sc, sizeof(double) * 16, cudaMemcpyDeviceToHost, st));
    FCVM_CUDA(cudaStreamSynchronize(st));
    if (multi && p2p_ready(c)) FCVM_TRY(p2p_check(c));
    if ((int)c->h_scalars[S_STATUS] == PCG_BREAKDOWN) {
      it = (int)c->h_scalars[S_ITERS];
      break;
    }
    if (c->h_scalars[S_ITERS] >= 0.0)
      n_it = (int)c->h_scalars[S_ITERS];
    else if (c->h_scalars[S_RR + (it & 1)] <= c->h_scalars[S_THR])
      n_it = it;
  }
  const double bb = c->h_scalars[S_BB];
  const bool conv = n_it >= 0;
  if (!conv) n_it = it;
  const double rr = c->h_scalars[S_RR + (n_it & 1)];
  if (iters) *iters = n_it;
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  if (recycle && conv) FCVM_TRY(remember_solution(c, b, x));
  if (!conv && (int)c->h_scalars[S_STATUS] == PCG_BREAKDOWN) {
    set_error("fcvm_pcg_solve: breakdown after %d iterations (p.Kp or r.M^-1 r not positive: the matrix is not positive "
              "definite -- the reference's 'singular stiffness matrix', fcVM.py:1367-1381); relative residual %.3e",
              n_it, bb > 0.0 ? sqrt(rr / bb) : 0.0);
    return FCVM_E_INDEFINITE;
  }
  if (!conv) {
    set_error("fcvm_pcg_solve: no convergence in %d iterations (relative residual %.3e, target %.3e)", max_iter,
              bb > 0.0 ? sqrt(rr / bb) : 0.0, rtol);
    return FCVM_E_NOCONV;
  }
  return FCVM_OK;
}
