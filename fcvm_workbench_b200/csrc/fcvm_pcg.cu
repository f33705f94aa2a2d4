// Block-SELL SpMV and the on-device preconditioned conjugate-gradient solve that replaces
// factor = cholesky(gsm); x = factor(b)   (fcVM.py:1121-1135, 1264-1278, 1369-1406).
//
// The loop runs without host round trips: step lengths live in device scalars, every dot
// product is a fixed-shape two-stage reduction (bit-reproducible), and a device flag turns
// the remaining launches of a batch into no-ops once the tolerance is met.  The host only
// polls that flag every CHECK_EVERY iterations.
#include "fcvm_common.cuh"
#include "fcvm_reduce.cuh"

using namespace fcvm;

extern "C" int fcvm_comm_allreduce_sum(fcvm_ctx *c, double *dev, int64_t n);
namespace fcvm {
int fcvm_comm_allreduce_oop(fcvm_ctx *c, const double *send, double *recv, int64_t n);
}

namespace {

constexpr int CHECK_EVERY = 16;

// Device scalar slots in ctx->red_out.  (r.z, r.r) live in two pairs that alternate as
// old/new between iterations; L_* are per-rank partial sums that NCCL reduces (out of place)
// into the shared slots when a communicator is attached.
enum {
  S_PQ = 0, S_RZ_A = 1, S_RR_A = 2, S_RZ_B = 3, S_RR_B = 4, S_BB = 5, S_THR = 6, S_DONE = 7, S_ITERS = 8,
  S_RRFIN = 9, L_PQ = 10, L_RZ = 11, L_RR = 12, L_BB = 13
};

// y = K x.  One thread per block row (SELL slot); a warp walks its slice column by column, so
// the column index and each of the nine block entries are read as coalesced lines.  Matrix
// data is streamed (evict-first) to keep x resident in L2.
__global__ void __launch_bounds__(256)
k_spmv_sell(int64_t nslices, const int32_t *__restrict__ slice_ptr, const int32_t *__restrict__ slot_node,
            const int32_t *__restrict__ colidx, const double *__restrict__ vals, const double *__restrict__ x,
            double *__restrict__ y, const double *__restrict__ done) {
  if (done && *done != 0.0) return;
  const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t s = slot / SELL_C;
  if (s >= nslices) return;
  const int lane = (int)(slot % SELL_C);
  const int32_t k0 = slice_ptr[s], k1 = slice_ptr[s + 1];
  double y0 = 0.0, y1 = 0.0, y2 = 0.0;
  const int32_t *ci = colidx + (int64_t)k0 * SELL_C + lane;
  const double *v = vals + (int64_t)k0 * 9 * SELL_C + lane;
#pragma unroll 4
  for (int32_t k = k0; k < k1; k++, ci += SELL_C, v += 9 * SELL_C) {
    const int64_t c3 = 3 * (int64_t)__ldcs(ci);
    const double a0 = __ldcs(v), a1 = __ldcs(v + SELL_C), a2 = __ldcs(v + 2 * SELL_C);
    const double a3 = __ldcs(v + 3 * SELL_C), a4 = __ldcs(v + 4 * SELL_C), a5 = __ldcs(v + 5 * SELL_C);
    const double a6 = __ldcs(v + 6 * SELL_C), a7 = __ldcs(v + 7 * SELL_C), a8 = __ldcs(v + 8 * SELL_C);
    const double x0 = x[c3], x1 = x[c3 + 1], x2 = x[c3 + 2];
    y0 += a0 * x0 + a1 * x1 + a2 * x2;
    y1 += a3 * x0 + a4 * x1 + a5 * x2;
    y2 += a6 * x0 + a7 * x1 + a8 * x2;
  }
  const int32_t row = slot_node[slot];
  if (row >= 0) {
    y[3 * (int64_t)row] = y0;
    y[3 * (int64_t)row + 1] = y1;
    y[3 * (int64_t)row + 2] = y2;
  }
}

__device__ __forceinline__ void apply_minv(const double *__restrict__ m, double r0, double r1, double r2, double &z0,
                                           double &z1, double &z2) {
  z0 = m[0] * r0 + m[1] * r1 + m[2] * r2;
  z1 = m[3] * r0 + m[4] * r1 + m[5] * r2;
  z2 = m[6] * r0 + m[7] * r1 + m[8] * r2;
}

// r = b - q (q = K x0, or absent), z = M^-1 r, p = z ; sums r.z, b.b, r.r
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_init(int64_t nn, const double *__restrict__ b, const double *__restrict__ q, const double *__restrict__ minv,
           const double *__restrict__ w, double *r, double *z, double *p, double *red_part, unsigned int *counter,
           double *out, Slots<3> sl) {
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
    z[d] = z0; z[d + 1] = z1; z[d + 2] = z2;
    p[d] = z0; p[d + 1] = z1; p[d + 2] = z2;
    v[0] += wt * (r0 * z0 + r1 * z1 + r2 * z2);
    v[2] += wt * (r0 * r0 + r1 * r1 + r2 * r2);
  }
  block_reduce_publish<3>(v, red_part, counter, out, sl);      // r.z, b.b, r.r
}

__global__ void k_pcg_scalars(double rtol, double *sc) {
  const double thr = rtol * rtol * sc[S_BB];
  sc[S_THR] = thr;
  sc[S_ITERS] = 0.0;
  sc[S_RRFIN] = sc[S_RR_A];
  sc[S_DONE] = (sc[S_RR_A] <= thr) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(RED_THREADS)
k_pcg_dot_pq(int64_t n, const double *__restrict__ p, const double *__restrict__ q, const double *__restrict__ w,
             double *red_part, unsigned int *counter, double *sc, Slots<1> sl) {
  if (sc[S_DONE] != 0.0) return;
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[0] += (w ? w[i] : 1.0) * p[i] * q[i];
  block_reduce_publish<1>(v, red_part, counter, sc, sl);
}

// alpha = rz/pq ; x += alpha p ; r -= alpha q ; z = M^-1 r ; sums r.z (new) and r.r
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_update1(int64_t nn, int rz_old, const double *__restrict__ p, const double *__restrict__ q,
              const double *__restrict__ minv, const double *__restrict__ w, double *x, double *r, double *z,
              double *red_part, unsigned int *counter, double *sc, Slots<2> sl) {
  if (sc[S_DONE] != 0.0) return;
  const double pq = sc[S_PQ];
  const double alpha = (pq != 0.0) ? sc[rz_old] / pq : 0.0;
  double v[2] = {0.0, 0.0};
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < nn; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = 3 * n;
    const double r0 = r[d] - alpha * q[d], r1 = r[d + 1] - alpha * q[d + 1], r2 = r[d + 2] - alpha * q[d + 2];
    x[d] += alpha * p[d];
    x[d + 1] += alpha * p[d + 1];
    x[d + 2] += alpha * p[d + 2];
    double z0, z1, z2;
    apply_minv(minv + 9 * n, r0, r1, r2, z0, z1, z2);
    r[d] = r0; r[d + 1] = r1; r[d + 2] = r2;
    z[d] = z0; z[d + 1] = z1; z[d + 2] = z2;
    const double wt = w ? w[d] : 1.0;
    v[0] += wt * (r0 * z0 + r1 * z1 + r2 * z2);
    v[1] += wt * (r0 * r0 + r1 * r1 + r2 * r2);
  }
  block_reduce_publish<2>(v, red_part, counter, sc, sl);       // r.z (new), r.r
}

// convergence test, then beta = rz_new/rz_old ; p = z + beta p
__global__ void __launch_bounds__(256)
k_pcg_update2(int64_t n, int rz_old_slot, int rz_new_slot, int iter, const double *__restrict__ z, double *p,
              double *sc) {
  if (sc[S_DONE] != 0.0) return;
  const double rr = sc[rz_new_slot + 1];
  const bool first = (blockIdx.x == 0 && threadIdx.x == 0);
  if (rr <= sc[S_THR]) {
    if (first) {
      sc[S_RRFIN] = rr;
      sc[S_ITERS] = (double)(iter + 1);
      sc[S_DONE] = 1.0;
    }
    return;
  }
  if (first) {
    sc[S_RRFIN] = rr;
    sc[S_ITERS] = (double)(iter + 1);
  }
  const double rz_old = sc[rz_old_slot], rz_new = sc[rz_new_slot];
  const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = z[i] + beta * p[i];
}

}  // namespace

namespace fcvm {
int launch_spmv(fcvm_ctx *c, const double *x, double *y) {
  ProfScope ps(c, 0);
  k_spmv_sell<<<grid_for(c->nslices * SELL_C, 256), 256, 0, c->stream>>>(c->nslices, c->slice_ptr, c->slot_node,
                                                                         c->colidx, c->vals, x, y, nullptr);
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

extern "C" int fcvm_pcg_solve(fcvm_ctx *c, const double *b, double *x, double rtol, int max_iter, int use_x0,
                              int *iters, double *relres) {
  FCVM_CHECK(c && c->assembled && b && x, FCVM_E_ARG, "fcvm_pcg_solve: assemble first / null argument");
  FCVM_CHECK(rtol > 0.0 && max_iter > 0, FCVM_E_ARG, "fcvm_pcg_solve: rtol and max_iter must be positive");
  const int64_t nn = c->nn, n3 = 3 * nn;
  cudaStream_t st = c->stream;
  double *sc = c->red_out;
  const double *w = c->dof_weight;
  const bool multi = c->world > 1;
  const int vgrid = grid_for(n3, 256);
  const double *q0 = nullptr;
  if (use_x0) {
    FCVM_TRY(fcvm_spmv(c, x, c->pcg_q));
    q0 = c->pcg_q;
  } else {
    FCVM_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n3, st));
  }
  {
    ProfScope ps(c, 3);
    Slots<3> sl = multi ? Slots<3>{{L_RZ, L_BB, L_RR}} : Slots<3>{{S_RZ_A, S_BB, S_RR_A}};
    k_pcg_init<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nn, b, q0, c->minv, w, c->pcg_r, c->pcg_z, c->pcg_p, c->red_part,
                                                   c->red_counter, sc, sl);
    c->launches++;
  }
  if (multi) {
    FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_RZ, sc + S_RZ_A, 2));
    FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_BB, sc + S_BB, 1));
  }
  k_pcg_scalars<<<1, 1, 0, st>>>(rtol, sc);
  c->launches++;
  int it = 0;
  bool done = false;
  while (!done && it < max_iter) {
    const int batch_end = std::min(max_iter, it + CHECK_EVERY);
    for (; it < batch_end; it++) {
      const int rz_old = (it & 1) ? S_RZ_B : S_RZ_A, rz_new = (it & 1) ? S_RZ_A : S_RZ_B;
      {
        ProfScope ps(c, 0);
        k_spmv_sell<<<grid_for(c->nslices * SELL_C, 256), 256, 0, st>>>(c->nslices, c->slice_ptr, c->slot_node,
                                                                       c->colidx, c->vals, c->pcg_p, c->pcg_q,
                                                                       sc + S_DONE);
      }
      if (multi) FCVM_TRY(fcvm_interface_sum(c, c->pcg_q));
      {
        ProfScope ps(c, 3);
        k_pcg_dot_pq<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n3, c->pcg_p, c->pcg_q, w, c->red_part, c->red_counter, sc,
                                                         Slots<1>{{multi ? L_PQ : S_PQ}});
      }
      if (multi) FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_PQ, sc + S_PQ, 1));
      {
        ProfScope ps(c, 3);
        Slots<2> sl = multi ? Slots<2>{{L_RZ, L_RR}} : Slots<2>{{rz_new, rz_new + 1}};
        k_pcg_update1<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nn, rz_old, c->pcg_p, c->pcg_q, c->minv, w, x, c->pcg_r,
                                                          c->pcg_z, c->red_part, c->red_counter, sc, sl);
      }
      if (multi) FCVM_TRY(fcvm_comm_allreduce_oop(c, sc + L_RZ, sc + rz_new, 2));
      {
        ProfScope ps(c, 3);
        k_pcg_update2<<<vgrid, 256, 0, st>>>(n3, rz_old, rz_new, it, c->pcg_z, c->pcg_p, sc);
      }
      c->launches += 4;
    }
    FCVM_CUDA(cudaGetLastError());
    FCVM_CUDA(cudaMemcpyAsync(c->h_scalars, sc, sizeof(double) * 16, cudaMemcpyDeviceToHost, st));
    FCVM_CUDA(cudaStreamSynchronize(st));
    done = c->h_scalars[S_DONE] != 0.0;
  }
  if (it == 0) {
    FCVM_CUDA(cudaMemcpyAsync(c->h_scalars, sc, sizeof(double) * 16, cudaMemcpyDeviceToHost, st));
    FCVM_CUDA(cudaStreamSynchronize(st));
  }
  const double bb = c->h_scalars[S_BB];
  if (iters) *iters = (int)c->h_scalars[S_ITERS];
  if (relres) *relres = bb > 0.0 ? sqrt(c->h_scalars[S_RRFIN] / bb) : 0.0;
  if (!(c->h_scalars[S_DONE] != 0.0)) {
    set_error("fcvm_pcg_solve: no convergence in %d iterations (relative residual %.3e, target %.3e)", max_iter,
              bb > 0.0 ? sqrt(c->h_scalars[S_RRFIN] / bb) : 0.0, rtol);
    return FCVM_E_NOCONV;
  }
  return FCVM_OK;
}
