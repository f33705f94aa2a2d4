// Rigid-body-mode deflation of the preconditioned CG solve (second level of the preconditioner).
//
// Block-Jacobi PCG needs O(L/h) iterations on an elasticity problem because the smooth, low-energy
// displacement fields are only reached slowly.  Here the nodes are grouped into box clusters and
// the six rigid-body modes of every cluster (three translations, three rotations about the box
// centre, rows of prescribed dofs zeroed) span a coarse space Z.  With E = Z^T K Z (dense, a few
// thousand unknowns, inverted once per assembly) the preconditioner becomes
//     u = y + Z E^-1 (Z^T r - (K Z)^T y),   y = D^-1 r          ("A-DEF2", Tang/Nabben/Vuik/Erlangga 2009)
// and the start vector x0 + Z E^-1 Z^T (b - K x0); the iteration count then scales with the cluster
// size H/h instead of the domain size L/h.  Everything is deterministic: fixed lists, fixed-shape
// reductions, no floating-point atomics.  K Z is stored sparsely per node (a node couples to at most
// 2 x 2 x 2 clusters because a cluster is at least two elements wide).
#include <cusolverDn.h>

#include <algorithm>

#include "fcvm_common.cuh"
#include "fcvm_deflation.cuh"
#include "fcvm_reduce.cuh"

using namespace fcvm;

extern "C" int fcvm_comm_allreduce_sum(fcvm_ctx *c, double *dev, int64_t n);
extern "C" int fcvm_comm_allreduce_max(fcvm_ctx *c, double *dev, int64_t n);

namespace {

// (K Z)_(i, c') for every block row i: one thread per row, walking its SELL slice.  Slot t of a row is
// the t-th distinct box met along the row (purely structural).  Structure pass (ent_inv == nullptr):
// only the relative box codes are written, slots whose 3x6 block vanishes -- a rigid motion of a box
// leaves the nodes in its interior force-free -- get -1.  Value pass: the blocks of the kept slots go to
// the entry-ordered, component-major array kzs[q][entry] that the per-iteration kernel streams.
__global__ void k_build_kz(int64_t nslices, Grid g, const int32_t *__restrict__ slice_ptr, const int32_t *__restrict__ slot_node,
                           const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                           const double *__restrict__ xyz, const double *__restrict__ fixdof,
                           const int32_t *__restrict__ cid, int8_t *__restrict__ kz_rel,
                           const int32_t *__restrict__ ent_inv, double *__restrict__ kzs, int64_t nent,
                           int *__restrict__ err) {
  const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t s = slot / SELL_C;
  if (s >= nslices) return;
  const int lane = (int)(slot % SELL_C);
  const int32_t row = slot_node[slot];
  if (row < 0) return;
  const int32_t ci = cid[row];
  int8_t codes[8];
  double acc[8][18];
  int used = 0;
  double amax = 0.0;
  for (int t = 0; t < 8; t++) {
    codes[t] = -1;
    for (int q = 0; q < 18; q++) acc[t][q] = 0.0;
  }
  for (int32_t k = slice_ptr[s]; k < slice_ptr[s + 1]; k++) {
    const int64_t pos = (int64_t)k * SELL_C + lane;
    const double *v = vals + (int64_t)k * 9 * SELL_C + lane;
    double a[9];
    for (int q = 0; q < 9; q++) {
      a[q] = v[q * SELL_C];
      amax = fmax(amax, fabs(a[q]));
    }
    const int32_t j = colidx[pos];                   // padding entries point at the row's own node
    const int32_t cj = cid[j];
    const int code = rel_code(g, ci, cj);
    if (code < 0) { atomicExch(err, 1); continue; }
    int t = 0;
    while (t < used && codes[t] != code) t++;
    if (t == used) {
      if (used == 8) { atomicExch(err, 2); continue; }
      codes[used++] = (int8_t)code;
    }
    double Z[3][6];
    z_of(g, cj, xyz, fixdof, j, Z);
    for (int r = 0; r < 3; r++)
      for (int m = 0; m < 6; m++) acc[t][6 * r + m] += a[3 * r] * Z[0][m] + a[3 * r + 1] * Z[1][m] + a[3 * r + 2] * Z[2][m];
  }
  if (!ent_inv) {
    for (int t = 0; t < 8; t++) {
      double mx = 0.0;
      for (int q = 0; q < 18; q++) mx = fmax(mx, fabs(acc[t][q]));
      kz_rel[8 * (int64_t)row + t] = (t < used && mx > 1e-13 * amax) ? codes[t] : (int8_t)-1;
    }
  } else {
    for (int t = 0; t < 8; t++) {
      const int32_t idx = ent_inv[8 * (int64_t)row + t];
      if (idx < 0) continue;
      for (int q = 0; q < 18; q++) kzs[(int64_t)q * nent + idx] = acc[t][q];
    }
  }
}

// E(c, c') = sum over the nodes i of cluster c of Z_i^T (K Z)_(i,c'): one block per cluster, nodes in
// list order, thread (t, entry) adds slot t's 6x6 contribution to the accumulator of its neighbour code
__global__ void __launch_bounds__(288)
k_build_e(Grid g, int64_t ncl, const int32_t *__restrict__ cl_ptr, const int32_t *__restrict__ cl_nodes,
          const int8_t *__restrict__ kz_rel, const int32_t *__restrict__ ent_inv, const double *__restrict__ kzs,
          int64_t nent, const double *__restrict__ xyz, const double *__restrict__ fixdof, double *__restrict__ E) {
  __shared__ double acc[27][36];
  const int c = blockIdx.x;
  for (int q = threadIdx.x; q < 27 * 36; q += blockDim.x) (&acc[0][0])[q] = 0.0;
  __syncthreads();
  const int t = threadIdx.x / 36, e = threadIdx.x % 36, a = e / 6, b = e % 6;
  for (int32_t idx = cl_ptr[c]; idx < cl_ptr[c + 1]; idx++) {
    const int64_t i = cl_nodes[idx];
    const int code = kz_rel[8 * i + t];
    if (code >= 0) {
      double Z[3][6];
      z_of(g, c, xyz, fixdof, i, Z);
      const double *kz = kzs + ent_inv[8 * i + t];
      acc[code][e] += Z[0][a] * kz[b * nent] + Z[1][a] * kz[(6 + b) * nent] + Z[2][a] * kz[(12 + b) * nent];
    }
    __syncthreads();
  }
  const int64_t n6 = 6 * ncl;
  for (int q = threadIdx.x; q < 27 * 36; q += blockDim.x) {
    const int code = q / 36, ee = q % 36;
    const int32_t cn = neighbour(g, c, code);
    if (cn >= 0) E[(6 * (int64_t)c + ee / 6) * n6 + 6 * (int64_t)cn + ee % 6] = acc[code][ee];
  }
}

// E <- (E + E^T)/2 on the lower triangle, unit diagonal where a mode has no free dof
__global__ void k_sym_guard(int64_t n, double *E) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int64_t r = idx / n, q = idx % n;
  if (r < q) return;
  double v = 0.5 * (E[r * n + q] + E[q * n + r]);
  if (r == q && v == 0.0) v = 1.0;
  E[r * n + q] = v;
}
__global__ void k_mirror(int64_t n, const double *__restrict__ L, double *__restrict__ F) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int64_t r = idx / n, q = idx % n;
  // the factorisation worked on the row-major lower triangle (cuSOLVER's column-major "upper")
  F[idx] = r >= q ? L[r * n + q] : L[q * n + r];
}

// rhs_c = sum_{i in c} w_i Z_i^T r_i  -  sum_{(i,t) -> c} (K Z)_(i,t)^T y_i        (one block per cluster)
// One block of RHS_T threads per box.  The time of a box (its ~1300 nodes and ~4500 entries, every trip a chain
// entry -> node -> y of dependent loads) is a floor that does not shrink when a rank holds only a few boxes, and
// with 256 threads it was 18 trips long: 512 threads at two blocks per SM halve it.  Optionally `split` blocks per
// box (a fixed share of the lists each; the block that finishes last adds the shares in order): measured no faster
// at any N, kept for boxes far larger than these.
constexpr int RHS_SPLIT = 4;
constexpr int RHS_T = 512;
template <typename CT>
__global__ void __launch_bounds__(RHS_T, 2)
k_coarse_rhs(Grid g, const int32_t *__restrict__ cl_ptr, const int32_t *__restrict__ cl_nodes,
             const int32_t *__restrict__ ent_ptr, const int32_t *__restrict__ ent_node, const CT *__restrict__ kzs,
             int64_t nent,
             const double *__restrict__ xyz, const double *__restrict__ fixdof, const double *__restrict__ wt,
             const double *__restrict__ r, const double *__restrict__ y, double *__restrict__ rhs,
             double *rhs_part, unsigned int *ticket, int split, const double *__restrict__ sc, int done_slot) {
  if (sc && sc[done_slot] >= 0.0) return;
  const int c = blockIdx.x / split, share = blockIdx.x % split;
  double v[6] = {0, 0, 0, 0, 0, 0};
  const int32_t nb = cl_ptr[c], nlen = cl_ptr[c + 1] - nb;
  const int32_t n0 = nb + (int32_t)((int64_t)nlen * share / split), n1 = nb + (int32_t)((int64_t)nlen * (share + 1) / split);
  for (int32_t idx = n0 + threadIdx.x; idx < n1; idx += RHS_T) {
    const int64_t i = cl_nodes[idx];
    double Z[3][6];
    z_of(g, c, xyz, fixdof, i, Z);
    const double w = wt ? wt[3 * i] : 1.0;
    const double r0 = w * r[3 * i], r1 = w * r[3 * i + 1], r2 = w * r[3 * i + 2];
#pragma unroll
    for (int m = 0; m < 6; m++) v[m] += Z[0][m] * r0 + Z[1][m] * r1 + Z[2][m] * r2;
  }
  if (y) {
    // entries of this box: consecutive lanes read consecutive doubles of each of the 18 component planes
    // two entries per trip: the dependent chains (entry -> node -> y) of both are in flight together
    const int32_t eb = ent_ptr[c], elen = ent_ptr[c + 1] - eb;
    const int32_t e1 = eb + (int32_t)((int64_t)elen * (share + 1) / split);
    int32_t idx = eb + (int32_t)((int64_t)elen * share / split) + threadIdx.x;
    for (; idx + RHS_T < e1; idx += 2 * RHS_T) {
      const int64_t i = ent_node[idx], j = ent_node[idx + RHS_T];
      const CT *kz = kzs + idx;
      const double y0 = y[3 * i], y1 = y[3 * i + 1], y2 = y[3 * i + 2];
      const double z0 = y[3 * j], z1 = y[3 * j + 1], z2 = y[3 * j + 2];
      double ka[18], kb[18];
#pragma unroll
      for (int q = 0; q < 18; q++) {
        ka[q] = (double)__ldcs(kz + q * nent);
        kb[q] = (double)__ldcs(kz + q * nent + RHS_T);
      }
#pragma unroll
      for (int m = 0; m < 6; m++) {
        v[m] -= ka[m] * y0 + ka[6 + m] * y1 + ka[12 + m] * y2;
        v[m] -= kb[m] * z0 + kb[6 + m] * z1 + kb[12 + m] * z2;
      }
    }
    for (; idx < e1; idx += RHS_T) {
      const int64_t i = ent_node[idx];
      const CT *kz = kzs + idx;
      const double y0 = y[3 * i], y1 = y[3 * i + 1], y2 = y[3 * i + 2];
#pragma unroll
      for (int m = 0; m < 6; m++)
        v[m] -= (double)__ldcs(kz + m * nent) * y0 + (double)__ldcs(kz + (6 + m) * nent) * y1 +
                (double)__ldcs(kz + (12 + m) * nent) * y2;
    }
  }
  __shared__ double sm[6][RHS_T / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const double s = warp_sum(v[m]);
    if (lane == 0) sm[m][warp] = s;
  }
  __syncthreads();
  const int64_t n6 = 6 * (int64_t)(gridDim.x / split);
  if (threadIdx.x < 6) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < RHS_T / 32; w++) s += sm[threadIdx.x][w];
    (split == 1 ? rhs : rhs_part + (int64_t)share * n6)[6 * (int64_t)c + threadIdx.x] = s;
  }
  if (split == 1) return;
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    for (int64_t q0 = threadIdx.x; q0 < n6; q0 += 4 * RHS_T) {
      double p[4][RHS_SPLIT];
#pragma unroll
      for (int t = 0; t < 4; t++)
#pragma unroll
        for (int k = 0; k < RHS_SPLIT; k++)
          p[t][k] = (k < split && q0 + RHS_T * t < n6) ? __ldcg(rhs_part + k * n6 + q0 + RHS_T * t) : 0.0;
#pragma unroll
      for (int t = 0; t < 4; t++)
        if (q0 + RHS_T * t < n6) rhs[q0 + RHS_T * t] = ((p[t][0] + p[t][1]) + p[t][2]) + p[t][3];
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// y = Einv[:, col0:col1) x[col0:col1) for all n rows: four warps per row (a quarter of the column range each, four
// loads in flight per lane), partial sums added in warp order.  On one GPU the range is everything; on a
// partitioned mesh a rank's right-hand side is non-zero only for the boxes around its own nodes, so it multiplies
// just that column panel and the ranks' products are summed (linearity) -- one exchange instead of two.
template <typename CT>
__global__ void __launch_bounds__(256)
k_gemv(int64_t n, int64_t ld, int64_t col0, int64_t col1, const CT *__restrict__ A, const double *__restrict__ x,
       double *__restrict__ y, const double *__restrict__ sc, int done_slot) {
  if (sc && sc[done_slot] >= 0.0) return;
  __shared__ double part[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, quarter = warp & 3;
  const int64_t row = blockIdx.x * 2 + (warp >> 2);
  double s = 0.0;
  if (row < n) {
    const CT *a = A + row * ld;
    const int64_t w = col1 - col0, c0 = col0 + quarter * w / 4, c1 = col0 + (quarter + 1) * w / 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t q = c0 + lane;
    for (; q + 96 < c1; q += 128) {
      s0 += (double)__ldcs(a + q) * x[q];
      s1 += (double)__ldcs(a + q + 32) * x[q + 32];
      s2 += (double)__ldcs(a + q + 64) * x[q + 64];
      s3 += (double)__ldcs(a + q + 96) * x[q + 96];
    }
    for (; q < c1; q += 32) s0 += (double)__ldcs(a + q) * x[q];
    s = warp_sum((s0 + s1) + (s2 + s3));
  }
  if (lane == 0) part[warp] = s;
  __syncthreads();
  if (row < n && quarter == 0 && lane == 0) y[row] = ((part[warp] + part[warp + 1]) + part[warp + 2]) + part[warp + 3];
}

// The same product as a bulk-copy stream over the single-precision copy: one block per SM, the panel of x staged
// once in shared memory, rows of E^-1 arriving RS at a time through an S-stage ring of bulk asynchronous copies
// (cp.async.bulk + mbarrier) -- the kernel is a pure stream of 4 x rows x columns bytes and the per-warp loads of
// k_gemv kept only ~25 kB in flight per SM.  Rows dealt round-robin to the blocks; every sum has a fixed shape.
// ld = row stride in floats (multiple of 4), [col0, col1) multiples of 4.
template <int RS, int S>
__global__ void __launch_bounds__(256, 1)
k_gemv_bulk(int64_t n, int64_t ld, int64_t col0, int64_t col1, const float *__restrict__ A, const double *__restrict__ x,
            double *__restrict__ y, const double *__restrict__ sc, int done_slot) {
  if (sc && sc[done_slot] >= 0.0) return;
  extern __shared__ __align__(128) unsigned char gm_raw[];
  const int cols = (int)(col1 - col0), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *xs = (double *)gm_raw;                                   // [cols]
  float *ring = (float *)(gm_raw + sizeof(double) * (size_t)cols); // [S][RS][cols]
  __shared__ uint64_t full[S], xbar;
  __shared__ double red[RS][8];
  if (tid == 0) {
    for (int q = 0; q < S; q++) mbar_init(&full[q], 1);
    mbar_init(&xbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ngroups = (n + RS - 1) / RS;                       // groups of RS consecutive rows
  auto issue = [&](int64_t g, int stage) {                         // thread 0 only
    const int64_t r0 = g * RS;
    const int nr = (int)min((int64_t)RS, n - r0);
    mbar_expect_tx(&full[stage], (uint32_t)(nr * cols * sizeof(float)));
    for (int q = 0; q < nr; q++)
      bulk_load(ring + ((size_t)stage * RS + q) * cols, A + (r0 + q) * ld + col0, (uint32_t)(cols * sizeof(float)), &full[stage]);
  };
  if (tid == 0) {
    mbar_expect_tx(&xbar, (uint32_t)(cols * sizeof(double)));
    bulk_load(xs, x + col0, (uint32_t)(cols * sizeof(double)), &xbar);
    for (int q = 0; q < S; q++)
      if (blockIdx.x + (int64_t)q * gridDim.x < ngroups) issue(blockIdx.x + (int64_t)q * gridDim.x, q);
  }
  mbar_wait(&xbar, 0);
  int k = 0;
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x, k++) {
    const int stage = k % S;
    mbar_wait(&full[stage], (uint32_t)((k / S) & 1));
    const float *a = ring + (size_t)stage * RS * cols;
    double acc[RS];
#pragma unroll
    for (int q = 0; q < RS; q++) acc[q] = 0.0;
    for (int cidx = tid; cidx < cols; cidx += 256) {
      const double xv = xs[cidx];
#pragma unroll
      for (int q = 0; q < RS; q++) acc[q] += (double)a[(size_t)q * cols + cidx] * xv;
    }
#pragma unroll
    for (int q = 0; q < RS; q++) {
      const double ws = warp_sum(acc[q]);
      if (lane == 0) red[q][warp] = ws;
    }
    __syncthreads();                                               // every thread has read the stage: it may be refilled
    if (tid == 0 && g + (int64_t)S * gridDim.x < ngroups) issue(g + (int64_t)S * gridDim.x, stage);
    if (tid < RS && g * RS + tid < n) {
      double t = 0.0;
#pragma unroll
      for (int wq = 0; wq < 8; wq++) t += red[tid][wq];
      y[g * RS + tid] = t;
    }
    __syncthreads();                                               // red[] is free again
  }
}

// out_i = (base ? base_i : 0) + Z_i lam_(cluster of i)
__global__ void k_expand(int64_t nn, Grid g, const int32_t *__restrict__ cid, const double *__restrict__ xyz,
                         const double *__restrict__ fixdof, const double *__restrict__ lam,
                         const double *__restrict__ base, double *__restrict__ out, const double *__restrict__ sc,
                         int done_slot) {
  if (sc && sc[done_slot] >= 0.0) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const int32_t c = cid[i];
  double Z[3][6];
  z_of(g, c, xyz, fixdof, i, Z);
  const double *l = lam + 6 * (int64_t)c;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    double s = base ? base[3 * i + r] : 0.0;
#pragma unroll
    for (int m = 0; m < 6; m++) s += Z[r][m] * l[m];
    out[3 * i + r] = s;
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
int dalloc2(T **p, int64_t n) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  FCVM_CUDA(cudaMalloc((void **)p, sizeof(T) * (size_t)std::max<int64_t>(n, 1)));
  return FCVM_OK;
}

}  // namespace

extern "C" int fcvm_set_deflation(fcvm_ctx *c, int ncx, int ncy, int ncz, const int32_t *cid, const double *lo,
                                  const double *h, const uint8_t *active) {
  FCVM_CHECK(c && c->nn > 0, FCVM_E_ARG, "fcvm_set_deflation: call fcvm_set_mesh first");
  c->defl_ready = c->defl_structure = false;
  if (ncx <= 0 || ncy <= 0 || ncz <= 0 || !cid) {          // switch off
    c->dn[0] = c->dn[1] = c->dn[2] = 0;
    c->ncl = 0;
    return FCVM_OK;
  }
  FCVM_CHECK(lo && h && h[0] > 0 && h[1] > 0 && h[2] > 0, FCVM_E_ARG, "fcvm_set_deflation: bad grid");
  const int64_t ncl = (int64_t)ncx * ncy * ncz, nn = c->nn;
  FCVM_CHECK(6 * ncl <= 16384, FCVM_E_ARG, "fcvm_set_deflation: %lld clusters make a coarse matrix beyond 16384 unknowns",
             (long long)ncl);
  c->dn[0] = ncx; c->dn[1] = ncy; c->dn[2] = ncz;
  c->ncl = ncl;
  for (int d = 0; d < 3; d++) { c->dlo[d] = lo[d]; c->dh[d] = h[d]; }
  c->dscale = std::max(h[0], std::max(h[1], h[2]));
  std::vector<int32_t> ptr((size_t)ncl + 1, 0), nodes((size_t)nn);
  for (int64_t i = 0; i < nn; i++) {
    FCVM_CHECK(cid[i] >= 0 && cid[i] < ncl, FCVM_E_ARG, "fcvm_set_deflation: cluster of node %lld out of range", (long long)i);
    ptr[(size_t)cid[i] + 1]++;
  }
  for (int64_t k = 0; k < ncl; k++) ptr[(size_t)k + 1] += ptr[(size_t)k];
  {
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < nn; i++) nodes[(size_t)fill[(size_t)cid[i]]++] = (int32_t)i;
  }
  FCVM_TRY(dalloc2(&c->d_cid, nn)); FCVM_TRY(dalloc2(&c->cl_ptr, ncl + 1)); FCVM_TRY(dalloc2(&c->cl_nodes, nn));
  FCVM_TRY(dalloc2(&c->cl_active, ncl));
  if (active)
    FCVM_CUDA(cudaMemcpy(c->cl_active, active, (size_t)ncl, cudaMemcpyHostToDevice));
  else
    FCVM_CUDA(cudaMemset(c->cl_active, 1, (size_t)ncl));
  FCVM_CUDA(cudaMemcpy(c->d_cid, cid, sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->cl_ptr, ptr.data(), sizeof(int32_t) * (ncl + 1), cudaMemcpyHostToDevice));
  FCVM_CUDA(cudaMemcpy(c->cl_nodes, nodes.data(), sizeof(int32_t) * nn, cudaMemcpyHostToDevice));
  FCVM_TRY(dalloc2(&c->kz_rel, 8 * nn));
  FCVM_TRY(dalloc2(&c->dE, 36 * ncl * ncl)); FCVM_TRY(dalloc2(&c->dEinv, 36 * ncl * ncl));
  c->einv_ld = (6 * ncl + 3) / 4 * 4;             // row stride of the single-precision E^-1 (16-byte rows)
  FCVM_TRY(dalloc2(&c->d_rhs, c->einv_ld)); FCVM_TRY(dalloc2(&c->d_lam, 6 * ncl));
  FCVM_CUDA(cudaMemset(c->d_rhs, 0, sizeof(double) * c->einv_ld));      // the padding columns stay zero
  FCVM_TRY(dalloc2(&c->rhs_part, RHS_SPLIT * 6 * ncl));
  {
    // Boxes this rank's right-hand side can touch: the boxes of its nodes and their neighbours -- a contiguous
    // range of box numbers that covers them.  The coarse product of a rank only needs those columns of E^-1.
    int64_t lo_c = ncl, hi_c = -1;
    for (int64_t i = 0; i < nn; i++) {
      lo_c = std::min<int64_t>(lo_c, cid[i]);
      hi_c = std::max<int64_t>(hi_c, cid[i]);
    }
    c->local_boxes = 0;
    for (int64_t k = 0; k < ncl; k++) c->local_boxes += ptr[(size_t)k + 1] > ptr[(size_t)k] ? 1 : 0;
    const int64_t reach = 1 + ncx + (int64_t)ncx * ncy;
    c->col0 = 6 * std::max<int64_t>(0, lo_c - reach);
    c->col1 = 6 * (std::min<int64_t>(ncl - 1, hi_c + reach) + 1);
  }
  FCVM_TRY(dalloc2(&c->spmv_part2, c->nslices + 8));
  FCVM_CUDA(cudaMemset(c->spmv_part2, 0, sizeof(double) * (c->nslices + 8)));
  return FCVM_OK;
}

// boxes and stored (node, box) entries of K Z (roofline arithmetic of the coarse kernels)
extern "C" int fcvm_deflation_stats(fcvm_ctx *c, int64_t *boxes, int64_t *entries) {
  FCVM_CHECK(c, FCVM_E_ARG, "null context");
  if (boxes) *boxes = c->ncl;
  if (entries) *entries = c->defl_structure ? c->nent : 0;
  return FCVM_OK;
}

namespace fcvm {

// K Z and E^-1 are streamed in single precision by default (they only shape the preconditioner);
// FCVM_COARSE_FP64=1 keeps the double-precision copies (comparison runs)
bool coarse_fp32() {
  static const bool v = !(getenv("FCVM_COARSE_FP64") && atoi(getenv("FCVM_COARSE_FP64")) != 0);
  return v;
}

// single-precision copies of K Z and E^-1 (values change with every assembly); rows of E^-1 padded to 16 bytes
static int refresh_coarse_fp32(fcvm_ctx *c) {
  if (!coarse_fp32()) return FCVM_OK;
  const int64_t nkz = 18 * c->nent, n6 = 6 * c->ncl, ne = n6 * c->einv_ld;
  if (!c->kz32) FCVM_CUDA(cudaMalloc((void **)&c->kz32, sizeof(float) * (size_t)std::max<int64_t>(nkz, 1)));
  if (!c->einv32) FCVM_CUDA(cudaMalloc((void **)&c->einv32, sizeof(float) * (size_t)std::max<int64_t>(ne, 1)));
  k_to_float<<<grid_for(nkz, 256), 256, 0, c->stream>>>(nkz, c->kz_val, c->kz32);
  k_to_float_rows<<<grid_for(ne, 256), 256, 0, c->stream>>>(n6, c->einv_ld, c->dEinv, c->einv32);
  c->launches += 2;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}
bool p2p_ready(const fcvm_ctx *c);
int p2p_allreduce_sum(fcvm_ctx *c, double *v, int64_t n, const double *sc, int done_slot);

// K Z, E = Z^T K Z and its inverse for the matrix now in the context (called at the end of fcvm_assemble)
int deflation_build(fcvm_ctx *c) {
  c->defl_ready = false;
  if (c->ncl == 0) return FCVM_OK;
  cudaStream_t st = c->stream;
  const Grid g = grid_of(c);
  const int64_t nn = c->nn, ncl = c->ncl, n6 = 6 * ncl;
  const double *fixdof = (const double *)c->buf[FCVM_BUF_FIXDOF];
  if (!c->cus_info) FCVM_CUDA(cudaMalloc((void **)&c->cus_info, sizeof(int) * 2));
  int *derr = c->cus_info + 1;                    // structure-error flag of the K Z kernels
  FCVM_CUDA(cudaMemsetAsync(derr, 0, sizeof(int), st));
  auto check = [&]() -> int {
    int herr = 0;
    FCVM_CUDA(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
    FCVM_CUDA(cudaStreamSynchronize(st));
    if (c->world > 1) {
      // every rank must take the same exit (collectives follow): the worst flag of all ranks decides
      double flag = (double)herr;
      FCVM_CUDA(cudaMemcpyAsync(c->d_rhs, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
      FCVM_TRY(fcvm_comm_allreduce_max(c, c->d_rhs, 1));
      FCVM_CUDA(cudaMemcpyAsync(&flag, c->d_rhs, sizeof(double), cudaMemcpyDeviceToHost, st));
      FCVM_CUDA(cudaStreamSynchronize(st));
      herr = (int)flag;
    }
    FCVM_CHECK(herr == 0, FCVM_E_ARG,
               "deflation: a node couples to %s -- the clusters must be at least two elements wide in every direction",
               herr == 1 ? "a cluster that is not a neighbour of its own" : "more than eight clusters");
    return FCVM_OK;
  };
  if (!c->defl_structure) {
    // structure pass: which (node, slot) pairs carry a non-vanishing block, and per target box the list
    // of those entries in ascending (node, slot) order -- the fixed order of every later reduction
    k_build_kz<<<grid_for(c->nslices * SELL_C, 128), 128, 0, st>>>(c->nslices, g, c->slice_ptr, c->slot_node, c->colidx,
                                                                  c->vals, c->xyz, fixdof, c->d_cid, c->kz_rel, nullptr,
                                                                  nullptr, 0, derr);
    FCVM_TRY(check());
    std::vector<int8_t> rel((size_t)8 * nn);
    std::vector<int32_t> cid((size_t)nn);
    FCVM_CUDA(cudaMemcpy(rel.data(), c->kz_rel, (size_t)8 * nn, cudaMemcpyDeviceToHost));
    FCVM_CUDA(cudaMemcpy(cid.data(), c->d_cid, sizeof(int32_t) * nn, cudaMemcpyDeviceToHost));
    const int nx = c->dn[0], ny = c->dn[1];
    auto target = [&](int32_t from, int code) {
      return (from % nx + code % 3 - 1) + nx * (((from / nx) % ny + (code / 3) % 3 - 1) + ny * (from / (nx * ny) + code / 9 - 1));
    };
    std::vector<int32_t> ptr((size_t)ncl + 1, 0);
    for (int64_t it = 0; it < 8 * nn; it++)
      if (rel[(size_t)it] >= 0) ptr[(size_t)target(cid[(size_t)(it >> 3)], rel[(size_t)it]) + 1]++;
    for (int64_t k = 0; k < ncl; k++) ptr[(size_t)k + 1] += ptr[(size_t)k];
    const int64_t nent = std::max<int64_t>(ptr[(size_t)ncl], 1);
    std::vector<int32_t> ent_node((size_t)nent, 0), inv((size_t)8 * nn, -1), fill(ptr.begin(), ptr.end() - 1);
    for (int64_t it = 0; it < 8 * nn; it++)
      if (rel[(size_t)it] >= 0) {
        const int32_t idx = fill[(size_t)target(cid[(size_t)(it >> 3)], rel[(size_t)it])]++;
        ent_node[(size_t)idx] = (int32_t)(it >> 3);
        inv[(size_t)it] = idx;
      }
    c->nent = nent;
    FCVM_TRY(dalloc2(&c->ent_ptr, ncl + 1)); FCVM_TRY(dalloc2(&c->ent, nent)); FCVM_TRY(dalloc2(&c->ent_inv, 8 * nn));
    FCVM_TRY(dalloc2(&c->kz_val, 18 * nent));
    FCVM_CUDA(cudaMemcpy(c->ent_ptr, ptr.data(), sizeof(int32_t) * (ncl + 1), cudaMemcpyHostToDevice));
    FCVM_CUDA(cudaMemcpy(c->ent, ent_node.data(), sizeof(int32_t) * nent, cudaMemcpyHostToDevice));
    FCVM_CUDA(cudaMemcpy(c->ent_inv, inv.data(), sizeof(int32_t) * 8 * nn, cudaMemcpyHostToDevice));
    cudaFree(c->kz32); cudaFree(c->einv32);       // sized by nent / ncl: reallocated by the refresh below
    c->kz32 = c->einv32 = nullptr;
    c->defl_structure = true;
  }
  k_build_kz<<<grid_for(c->nslices * SELL_C, 128), 128, 0, st>>>(c->nslices, g, c->slice_ptr, c->slot_node, c->colidx,
                                                                c->vals, c->xyz, fixdof, c->d_cid, c->kz_rel, c->ent_inv,
                                                                c->kz_val, c->nent, derr);
  FCVM_TRY(check());
  FCVM_CUDA(cudaMemsetAsync(c->dE, 0, sizeof(double) * n6 * n6, st));
  k_build_e<<<(unsigned)ncl, 288, 0, st>>>(g, ncl, c->cl_ptr, c->cl_nodes, c->kz_rel, c->ent_inv, c->kz_val, c->nent, c->xyz,
                                           fixdof, c->dE);
  if (c->world > 1) FCVM_TRY(fcvm_comm_allreduce_sum(c, c->dE, n6 * n6));
  k_sym_guard<<<grid_for(n6 * n6, 256), 256, 0, st>>>(n6, c->dE);
  FCVM_CUDA(cudaGetLastError());
  // dense inverse (cuSOLVER Cholesky; column-major view of the symmetric matrix: "upper" = our lower triangle)
  cusolverDnHandle_t h = (cusolverDnHandle_t)c->cusolver;
  if (!h) {
    FCVM_CHECK(cusolverDnCreate(&h) == CUSOLVER_STATUS_SUCCESS, FCVM_E_CUDA, "cusolverDnCreate failed");
    c->cusolver = (void *)h;
  }
  cusolverDnSetStream(h, st);
  int lw1 = 0, lw2 = 0;
  cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_UPPER, (int)n6, c->dE, (int)n6, &lw1);
  cusolverDnDpotri_bufferSize(h, CUBLAS_FILL_MODE_UPPER, (int)n6, c->dE, (int)n6, &lw2);
  const int lw = std::max(lw1, lw2);
  if (lw > c->cus_lwork) {
    if (c->cus_work) cudaFree(c->cus_work);
    FCVM_CUDA(cudaMalloc((void **)&c->cus_work, sizeof(double) * (size_t)lw));
    c->cus_lwork = lw;
  }
  int info = 0;
  FCVM_CHECK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_UPPER, (int)n6, c->dE, (int)n6, c->cus_work, lw, c->cus_info) ==
                 CUSOLVER_STATUS_SUCCESS, FCVM_E_CUDA, "cusolverDnDpotrf failed");
  FCVM_CUDA(cudaMemcpyAsync(&info, c->cus_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  FCVM_CUDA(cudaStreamSynchronize(st));
  FCVM_CHECK(info == 0, FCVM_E_ARG, "deflation: coarse matrix is not positive definite (potrf info %d)", info);
  FCVM_CHECK(cusolverDnDpotri(h, CUBLAS_FILL_MODE_UPPER, (int)n6, c->dE, (int)n6, c->cus_work, lw, c->cus_info) ==
                 CUSOLVER_STATUS_SUCCESS, FCVM_E_CUDA, "cusolverDnDpotri failed");
  k_mirror<<<grid_for(n6 * n6, 256), 256, 0, st>>>(n6, c->dE, c->dEinv);
  FCVM_CUDA(cudaGetLastError());
  c->launches += 5;
  FCVM_TRY(refresh_coarse_fp32(c));
  c->defl_ready = true;
  return FCVM_OK;
}

// lam = E^-1 (Z^T r - (K Z)^T y); out = base + Z lam.   y / base may be null.
int deflation_correct(fcvm_ctx *c, const double *r, const double *y, const double *base, double *out, const double *sc,
                      int done_slot) {
  cudaStream_t st = c->stream;
  const Grid g = grid_of(c);
  const double *fixdof = (const double *)c->buf[FCVM_BUF_FIXDOF];
  const int64_t n6 = 6 * c->ncl;
  // the coarse operators only shape the preconditioner: single-precision copies halve their traffic
  const bool f32 = coarse_fp32() && c->kz32 && c->einv32;
  {
  ProfScope ps8(c, 8);
  static const int split_env = getenv("FCVM_RHS_SPLIT") ? std::max(1, std::min(RHS_SPLIT, atoi(getenv("FCVM_RHS_SPLIT")))) : 1;
  const int split = split_env;
  if (f32)
    k_coarse_rhs<float><<<(unsigned)(split * c->ncl), RHS_T, 0, st>>>(g, c->cl_ptr, c->cl_nodes, c->ent_ptr, c->ent, c->kz32,
                                                                   c->nent, c->xyz, fixdof, c->dof_weight, r, y, c->d_rhs,
                                                                   c->rhs_part, c->red_counter + 3, split, sc, done_slot);
  else
    k_coarse_rhs<double><<<(unsigned)(split * c->ncl), RHS_T, 0, st>>>(g, c->cl_ptr, c->cl_nodes, c->ent_ptr, c->ent, c->kz_val,
                                                                    c->nent, c->xyz, fixdof, c->dof_weight, r, y, c->d_rhs,
                                                                    c->rhs_part, c->red_counter + 3, split, sc, done_slot);
  }
  {
    // every rank holds E^-1 and multiplies the column panel its own right-hand side lives in (all columns on one GPU)
    const bool multi = c->world > 1;
    const int64_t col0 = multi ? c->col0 : 0, col1 = multi ? c->col1 : n6;
    ProfScope ps9(c, 9);
    // the panel widened to 16-byte boundaries: the extra columns are zeros of the right-hand side (outside the
    // hull of this rank's boxes, or padding)
    const int64_t b0 = col0 / 4 * 4, b1 = std::min<int64_t>((col1 + 3) / 4 * 4, c->einv_ld), bcols = b1 - b0;
    static const bool no_bulk = getenv("FCVM_GEMV_BULK") && atoi(getenv("FCVM_GEMV_BULK")) == 0;
    // rows per stage: as many as the panel leaves room for in 220 kB of shared memory (narrow panels of a rank
    // with few boxes would otherwise pay a barrier per two short rows)
    const size_t smem_cap = 220 * 1024;
    auto need = [&](int rs, int stages) { return (size_t)bcols * (8 + (size_t)stages * rs * 4); };
    if (f32 && !no_bulk && need(1, 2) <= smem_cap) {
      static int sms = 0;
      if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        FCVM_CUDA(cudaFuncSetAttribute(k_gemv_bulk<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        FCVM_CUDA(cudaFuncSetAttribute(k_gemv_bulk<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        FCVM_CUDA(cudaFuncSetAttribute(k_gemv_bulk<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        FCVM_CUDA(cudaFuncSetAttribute(k_gemv_bulk<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
      }
      if (need(8, 3) <= smem_cap)
        k_gemv_bulk<8, 3><<<sms, 256, need(8, 3), st>>>(n6, c->einv_ld, b0, b1, c->einv32, c->d_rhs, c->d_lam, sc, done_slot);
      else if (need(4, 3) <= smem_cap)
        k_gemv_bulk<4, 3><<<sms, 256, need(4, 3), st>>>(n6, c->einv_ld, b0, b1, c->einv32, c->d_rhs, c->d_lam, sc, done_slot);
      else if (need(2, 3) <= smem_cap)
        k_gemv_bulk<2, 3><<<sms, 256, need(2, 3), st>>>(n6, c->einv_ld, b0, b1, c->einv32, c->d_rhs, c->d_lam, sc, done_slot);
      else
        k_gemv_bulk<1, 2><<<sms, 256, need(1, 2), st>>>(n6, c->einv_ld, b0, b1, c->einv32, c->d_rhs, c->d_lam, sc, done_slot);
    } else if (f32)
      k_gemv<float><<<grid_for(n6, 2), 256, 0, st>>>(n6, c->einv_ld, col0, col1, c->einv32, c->d_rhs, c->d_lam, sc, done_slot);
    else
      k_gemv<double><<<grid_for(n6, 2), 256, 0, st>>>(n6, n6, col0, col1, c->dEinv, c->d_rhs, c->d_lam, sc, done_slot);
  }
  if (c->world > 1) {
    // lam = sum over ranks of their products: inside one box a peer-memory exchange (fcvm_p2p.cu: sums in rank
    // order, identical on all ranks), else an NCCL all-reduce
    if (p2p_ready(c))
      FCVM_TRY(p2p_allreduce_sum(c, c->d_lam, n6, sc, done_slot));
    else
      FCVM_TRY(fcvm_comm_allreduce_sum(c, c->d_lam, n6));
  }
  ProfScope ps10(c, 10);
  k_expand<<<grid_for(c->nn, 256), 256, 0, st>>>(c->nn, g, c->d_cid, c->xyz, fixdof, c->d_lam, base, out, sc, done_slot);
  c->launches += 3;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

void deflation_free(fcvm_ctx *c) {
  cudaFree(c->kz32); cudaFree(c->einv32);
  c->kz32 = c->einv32 = nullptr;
  cudaFree(c->cl_active); c->cl_active = nullptr;
  cudaFree(c->d_cid); cudaFree(c->cl_ptr); cudaFree(c->cl_nodes); cudaFree(c->kz_rel); cudaFree(c->kz_val);
  cudaFree(c->ent_ptr); cudaFree(c->ent); cudaFree(c->ent_inv); c->ent_inv = nullptr; cudaFree(c->dE); cudaFree(c->dEinv); cudaFree(c->d_rhs); cudaFree(c->d_lam);
  cudaFree(c->spmv_part2); cudaFree(c->rhs_part);
  c->rhs_part = nullptr;
  c->d_cid = c->cl_ptr = c->cl_nodes = c->ent_ptr = c->ent = nullptr;
  c->kz_rel = nullptr;
  c->kz_val = c->dE = c->dEinv = c->d_rhs = c->d_lam = c->spmv_part2 = nullptr;
  c->ncl = 0;
  c->dn[0] = c->dn[1] = c->dn[2] = 0;
  c->defl_ready = c->defl_structure = false;
}

}  // namespace fcvm
