// Matrix-free product y = K x for the ELASTIC stiffness of calcGSM (fcVM.py:620-816) inside the PCG solve.
//
// The assembled block-SELL matrix costs 76 bytes of HBM traffic per stored 3x3 block (3.0 GB per product at
// 1M elements) and the SpMV over it already runs at the copy bandwidth, so the only way to a faster product is
// not to read the matrix: K x = sum over elements of  sum_gp w|J| B^T D B x_e  is recomputed from the nodal
// coordinates (L2-resident, 33 MB) with the same Gauss-point kinematics as the stress update
// (fcVM.py:2256-2454 without the plastic correction), ~2.5 kFLOP and 0.3 kB of HBM traffic per element:
//
//   k_elastic_apply   block = 32 elements x 2 warps, two Gauss points per thread; nodal coordinates and the
//                     masked vector x gathered once into a conflict-free shared tile; element vectors leave as
//                     one contiguous 7.7 kB run (the element-vector scratch of the internal-force assembly)
//   k_gather_apply    y[dof] = sum of the element vectors around the node in ascending element order (no float
//                     atomics, bit-reproducible), rows of prescribed dofs = (elements at the node) * x -- exactly
//                     the constrained operator fcvm_assemble builds (fcVM.py:773-787) -- and the block partials
//                     of y.x and r.x that the single-reduction PCG needs
//
// Used for the geometrically linear analysis, where the matrix is the elastic one for the whole run; the
// tangent of the large-displacement branch changes every Newton iteration and keeps the assembled SpMV.
#include <algorithm>

#include "fcvm_common.cuh"

using namespace fcvm;

namespace {

constexpr int MF_E = 32;            // elements per block
constexpr int MF_THREADS = 64;      // two warps: warp p integrates Gauss points 2p and 2p+1
constexpr int MF_ROW = 35;          // doubles per element of the nodal staging (30 used)
constexpr int MF_PAD = 33;          // row stride of the force staging

// emask[e]: bit 3k+c set when dof c of local node k is prescribed
__global__ void k_elem_mask(int64_t ne, const int32_t *__restrict__ conn, const uint8_t *__restrict__ fixmask,
                            uint32_t *__restrict__ emask) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  uint32_t m = 0;
  for (int k = 0; k < 10; k++) {
    const int64_t n = conn[(int64_t)k * ne + e];
    for (int c = 0; c < 3; c++)
      if (fixmask[3 * n + c]) m |= 1u << (3 * k + c);
  }
  emask[e] = m;
}

__global__ void __launch_bounds__(MF_THREADS, 8)
k_elastic_apply(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                const double *__restrict__ x, const uint32_t *__restrict__ emask, double lambda, double mu,
                double *__restrict__ elv, const double *__restrict__ sc, int rr_slot, int iters_slot, int thr_slot,
                const uint8_t *__restrict__ tile_affine) {
  if (sc && (sc[iters_slot] >= 0.0 || sc[rr_slot] <= sc[thr_slot])) return;      // batch already converged
  if (tile_affine && tile_affine[blockIdx.x]) return;                            // k_elastic_apply_affine's tile
  __shared__ double smem[2 * MF_E * MF_ROW];      // nodal staging [2][32][35], then force staging [2][30][33]
  double *sX = smem, *sU = smem + MF_E * MF_ROW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * MF_E;
  const int gpa = 2 * warp, gpb = gpa + 1;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    // item q = 3*(32*node + element) + component (see k_stress_update: component-adjacent lanes, conflict-free tile)
    constexpr int NQ = 8;
    int64_t d[NQ];
    uint32_t keep[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = min(tid + (half * NQ + r) * MF_THREADS, 30 * MF_E - 1);
      const int p = q / 3, cpt = q - 3 * p;
      const int64_t el = min(e0 + (p & 31), ne - 1);
      d[r] = 3 * (int64_t)conn[(int64_t)(p >> 5) * ne + el] + cpt;
      keep[r] = ((emask[el] >> (3 * (p >> 5) + cpt)) & 1u) ^ 1u;
    }
    double xv[NQ], uv[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      xv[r] = xyz[d[r]];
      uv[r] = x[d[r]];
    }
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = tid + (half * NQ + r) * MF_THREADS;
      if (q < 30 * MF_E) {
        const int p = q / 3;
        const int at = (p & 31) * MF_ROW + 3 * (p >> 5) + (q - 3 * p);
        sX[at] = xv[r];
        sU[at] = keep[r] ? uv[r] : 0.0;            // columns of prescribed dofs are eliminated
      }
    }
  }
  __syncthreads();
  const GPCoef ca = gp_coef(gpa), cb = gp_coef(gpb);
  double xsiA[3][3], xsiB[3][3], gA[3][3], gB[3][3], wA, wB;
  {
    double xa[3][3], xb[3][3];
    local_gradient_tile2(ca, cb, sX + lane * MF_ROW, 1, xa, xb);
    wA = GP_W * fabs(invert_jacobian(xa, xsiA));
    wB = GP_W * fabs(invert_jacobian(xb, xsiB));
    local_gradient_tile2(ca, cb, sU + lane * MF_ROW, 1, xa, xb);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int mm = 0; mm < 3; mm++) {
        gA[i][mm] = xa[i][0] * xsiA[0][mm] + xa[i][1] * xsiA[1][mm] + xa[i][2] * xsiA[2][mm];
        gB[i][mm] = xb[i][0] * xsiB[0][mm] + xb[i][1] * xsiB[1][mm] + xb[i][2] * xsiB[2][mm];
      }
  }
  __syncthreads();      // staged nodal data consumed: the force staging may overwrite it
  double F[30];
  {
    // sigma = lambda tr(eps) I + 2 mu eps (Hooke matrix of fcVM.py:574-582); T = w sigma xsi^T
    double T[3][3];
    auto stress_T = [&](const double (&g)[3][3], const double (&xsi)[3][3], double w) {
      const double tr = lambda * (g[0][0] + g[1][1] + g[2][2]);
      double S[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[i][j] = mu * (g[i][j] + g[j][i]) + (i == j ? tr : 0.0);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[i][j] = w * (S[i][0] * xsi[j][0] + S[i][1] * xsi[j][1] + S[i][2] * xsi[j][2]);
    };
    stress_T(gA, xsiA, wA);
    gradient_to_regs<false>(ca, T, F);
    stress_T(gB, xsiB, wB);
    gradient_to_regs<true>(cb, T, F);
  }
  double *sF = smem + (warp * 30) * MF_PAD + lane;
#pragma unroll
  for (int k = 0; k < 30; k++) sF[k * MF_PAD] = F[k];
  __syncthreads();
  const int nlive = (int)min((int64_t)MF_E, ne - e0) * 30;
  double *out = elv + 30 * e0;
  for (int idx = tid; idx < nlive; idx += MF_THREADS) {
    const int el = idx / 30, k3 = idx - 30 * el;
    const double *f = smem + k3 * MF_PAD + el;
    out[idx] = f[0] + f[30 * MF_PAD];
  }
}

// Straight-sided elements (every mid-side node at the mean of its two corners to round-off -- the interior of any
// mesh, all of a structured one) have one Jacobian for all four Gauss points.  Their inverse Jacobian and w|J| are
// stored once per mesh (80 bytes per element); a tile made of such elements skips the coordinate gather, the four
// Jacobians and their inverses: a third of the FP64 work of the general path.
__global__ void k_elem_geometry(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                                double *__restrict__ egeo, uint8_t *__restrict__ affine) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double X[30], scale = 0.0;
  for (int k = 0; k < 10; k++) {
    const int64_t n = conn[(int64_t)k * ne + e];
    for (int c = 0; c < 3; c++) {
      X[3 * k + c] = xyz[3 * n + c];
      scale = fmax(scale, fabs(X[3 * k + c]));
    }
  }
  // mid-side nodes 4..9 sit on the edges (0,1) (1,2) (0,2) (0,3) (1,3) (2,3)
  const int ea[6] = {0, 1, 0, 0, 1, 2}, eb[6] = {1, 2, 2, 3, 3, 3};
  bool aff = true;
  for (int m = 0; m < 6; m++)
    for (int c = 0; c < 3; c++)
      if (fabs(X[3 * (4 + m) + c] - 0.5 * (X[3 * ea[m] + c] + X[3 * eb[m] + c])) > 8.0 * 2.220446049250313e-16 * scale) aff = false;
  affine[e] = aff ? 1 : 0;
  const GPCoef cf = gp_coef(0);
  double xs[3][3], xsi[3][3];
  local_gradient_tile(cf, X, 1, xs);
  const double xsj = invert_jacobian(xs, xsi);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) egeo[(int64_t)(3 * i + j) * ne + e] = xsi[i][j];
  egeo[(int64_t)9 * ne + e] = GP_W * fabs(xsj);
}

__global__ void k_tile_affine(int64_t ne, const uint8_t *__restrict__ affine, uint8_t *__restrict__ tile) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t * MF_E >= ne) return;
  uint8_t all = 1;
  for (int64_t e = t * MF_E; e < min(ne, (t + 1) * MF_E); e++) all &= affine[e];
  tile[t] = all;
}

// the product on a tile of straight-sided elements (same layout and order of operations as k_elastic_apply)
#ifndef MF_AFF_MINB
#define MF_AFF_MINB 10
#endif
__global__ void __launch_bounds__(MF_THREADS, MF_AFF_MINB)
k_elastic_apply_affine(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ egeo,
                       const uint8_t *__restrict__ tile_affine, const double *__restrict__ x,
                       const uint32_t *__restrict__ emask, double lambda, double mu, double *__restrict__ elv,
                       const double *__restrict__ sc, int rr_slot, int iters_slot, int thr_slot) {
  if (sc && (sc[iters_slot] >= 0.0 || sc[rr_slot] <= sc[thr_slot])) return;
  if (!tile_affine[blockIdx.x]) return;            // the general kernel takes this tile
  __shared__ double smem[2 * 30 * MF_PAD];         // x staging [32][35] first, then force staging [2][30][33]
  double *sU = smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * MF_E;
  const int64_t e = min(e0 + lane, ne - 1);
  const int gpa = 2 * warp, gpb = gpa + 1;
  double xsi[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) xsi[i][j] = egeo[(int64_t)(3 * i + j) * ne + e];
  const double w = egeo[(int64_t)9 * ne + e];
  {
    constexpr int NQ = 15;                         // 960 items / 64 threads
    int64_t d[NQ];
    uint32_t keep[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = tid + r * MF_THREADS;
      const int p = q / 3, cpt = q - 3 * p;
      const int64_t el = min(e0 + (p & 31), ne - 1);
      d[r] = 3 * (int64_t)conn[(int64_t)(p >> 5) * ne + el] + cpt;
      keep[r] = ((emask[el] >> (3 * (p >> 5) + cpt)) & 1u) ^ 1u;
    }
    double uv[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) uv[r] = x[d[r]];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = tid + r * MF_THREADS;
      const int p = q / 3;
      sU[(p & 31) * MF_ROW + 3 * (p >> 5) + (q - 3 * p)] = keep[r] ? uv[r] : 0.0;
    }
  }
  __syncthreads();
  const GPCoef ca = gp_coef(gpa), cb = gp_coef(gpb);
  double gA[3][3], gB[3][3];
  {
    double ha[3][3], hb[3][3];
    local_gradient_tile2(ca, cb, sU + lane * MF_ROW, 1, ha, hb);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int mm = 0; mm < 3; mm++) {
        gA[i][mm] = ha[i][0] * xsi[0][mm] + ha[i][1] * xsi[1][mm] + ha[i][2] * xsi[2][mm];
        gB[i][mm] = hb[i][0] * xsi[0][mm] + hb[i][1] * xsi[1][mm] + hb[i][2] * xsi[2][mm];
      }
  }
  __syncthreads();
  double F[30];
  {
    double T[3][3];
    auto stress_T = [&](const double (&g)[3][3]) {
      const double tr = lambda * (g[0][0] + g[1][1] + g[2][2]);
      double S[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) S[i][j] = mu * (g[i][j] + g[j][i]) + (i == j ? tr : 0.0);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[i][j] = w * (S[i][0] * xsi[j][0] + S[i][1] * xsi[j][1] + S[i][2] * xsi[j][2]);
    };
    stress_T(gA);
    gradient_to_regs<false>(ca, T, F);
    stress_T(gB);
    gradient_to_regs<true>(cb, T, F);
  }
  double *sF = smem + (warp * 30) * MF_PAD + lane;
#pragma unroll
  for (int k = 0; k < 30; k++) sF[k * MF_PAD] = F[k];
  __syncthreads();
  const int nlive = (int)min((int64_t)MF_E, ne - e0) * 30;
  double *out = elv + 30 * e0;
  for (int idx = tid; idx < nlive; idx += MF_THREADS) {
    const int el = idx / 30, k3 = idx - 30 * el;
    const double *f = smem + k3 * MF_PAD + el;
    out[idx] = f[0] + f[30 * MF_PAD];
  }
}

// one thread per dof (three lanes share a node's 24-byte pieces of the element vectors): y, the constrained rows,
// the dot-product partials of the block.  Nodes stay in their natural order -- walking them sorted by degree was
// measured twice as slow (the locality of the element vectors is worth more than even warps).
constexpr int GA_THREADS = 256;
constexpr int GA_GROUP = 64;        // blocks per group of the two-level dot-product finish
__global__ void __launch_bounds__(GA_THREADS)
k_gather_apply(int64_t nn, const int32_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e_idx,
               const double *__restrict__ elv, const uint8_t *__restrict__ fixmask, const double *__restrict__ x,
               double *__restrict__ y, const double *sc, int rr_slot, int iters_slot, int thr_slot,
               double *dot_part, const double *__restrict__ rvec, const double *__restrict__ wt, double *dot_part2,
               unsigned int *ticket, double *group_part, double *sc_out, int delta_slot, int gamma_slot) {
  if (sc && (sc[iters_slot] >= 0.0 || sc[rr_slot] <= sc[thr_slot])) return;
  const int64_t d = blockIdx.x * (int64_t)GA_THREADS + threadIdx.x;
  double dsum = 0.0, rsum = 0.0;
  if (d < 3 * nn) {
    const int64_t n = d / 3;
    const int cpt = (int)(d - 3 * n);
    const double *ev = elv + cpt;
    double s = 0.0;
    const int32_t b = n2e_ptr[n], eend = n2e_ptr[n + 1];
    int32_t k = b;
    // four element vectors in flight; the additions stay in ascending element order
    for (; k + 3 < eend; k += 4) {
      const double a0 = ev[3 * (int64_t)n2e_idx[k]], a1 = ev[3 * (int64_t)n2e_idx[k + 1]];
      const double a2 = ev[3 * (int64_t)n2e_idx[k + 2]], a3 = ev[3 * (int64_t)n2e_idx[k + 3]];
      s = (((s + a0) + a1) + a2) + a3;
    }
    for (; k < eend; k++) s += ev[3 * (int64_t)n2e_idx[k]];
    const double xd = x[d];
    if (fixmask[d]) s = (double)(eend - b) * xd;
    y[d] = s;
    if (dot_part) dsum = s * xd;
    if (dot_part2) rsum = (wt ? wt[d] : 1.0) * rvec[d] * xd;
  }
  if (dot_part || dot_part2) {
    __shared__ double sm[2][GA_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    dsum = warp_sum(dsum);
    rsum = warp_sum(rsum);
    if (lane == 0) {
      sm[0][warp] = dsum;
      sm[1][warp] = rsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, bsum = 0.0;
#pragma unroll
      for (int w = 0; w < GA_THREADS / 32; w++) {
        a += sm[0][w];
        bsum += sm[1][w];
      }
      if (dot_part) dot_part[blockIdx.x] = a;
      if (dot_part2) dot_part2[blockIdx.x] = bsum;
    }
    if (ticket) {
      // Two-level finish without a separate launch and without a long serial tail: the block that completes a
      // group of GA_GROUP blocks adds that group's partials, the block that completes the last group adds the
      // group sums and publishes delta = y.x (and gamma = r.x).  The shape of the sums is fixed, whichever
      // blocks happen to do them.
      __shared__ int role;
      const int grp = blockIdx.x / GA_GROUP, ngroups = (gridDim.x + GA_GROUP - 1) / GA_GROUP;
      const int in_group = min(GA_GROUP, (int)gridDim.x - grp * GA_GROUP);
      if (threadIdx.x == 0) {
        __threadfence();
        role = (atomicAdd(ticket + 1 + grp, 1u) == (unsigned)(in_group - 1)) ? 1 : 0;
      }
      __syncthreads();
      if (role) {
        double t = 0.0, g = 0.0;
        if ((int)threadIdx.x < in_group) {
          if (dot_part) t = __ldcg(dot_part + grp * GA_GROUP + threadIdx.x);
          if (dot_part2) g = __ldcg(dot_part2 + grp * GA_GROUP + threadIdx.x);
        }
        t = warp_sum(t);
        g = warp_sum(g);
        __syncthreads();
        if (lane == 0) {
          sm[0][warp] = t;
          sm[1][warp] = g;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          double tot = 0.0, gam = 0.0;
#pragma unroll
          for (int w = 0; w < GA_GROUP / 32; w++) {
            tot += sm[0][w];
            gam += sm[1][w];
          }
          group_part[grp] = tot;
          group_part[ngroups + grp] = gam;
          ticket[1 + grp] = 0u;
          __threadfence();
          role = (atomicAdd(ticket, 1u) == (unsigned)(ngroups - 1)) ? 2 : 1;
        }
        __syncthreads();
        if (role == 2) {
          t = 0.0;
          g = 0.0;
          for (int i = threadIdx.x; i < ngroups; i += GA_THREADS) {
            t += __ldcg(group_part + i);
            g += __ldcg(group_part + ngroups + i);
          }
          t = warp_sum(t);
          g = warp_sum(g);
          __syncthreads();
          if (lane == 0) {
            sm[0][warp] = t;
            sm[1][warp] = g;
          }
          __syncthreads();
          if (threadIdx.x == 0) {
            double tot = 0.0, gam = 0.0;
#pragma unroll
            for (int w = 0; w < GA_THREADS / 32; w++) {
              tot += sm[0][w];
              gam += sm[1][w];
            }
            sc_out[delta_slot] = tot;
            if (dot_part2 && gamma_slot >= 0) sc_out[gamma_slot] = gam;
            *ticket = 0u;
          }
        }
      }
    }
  }
}

}  // namespace

namespace fcvm {

int64_t matfree_parts(const fcvm_ctx *c) { return (3 * c->nn + GA_THREADS - 1) / GA_THREADS; }

// per-mesh data of the matrix-free product: element geometry of the straight-sided elements
int matfree_set_mesh(fcvm_ctx *c) {
  const int64_t ne = c->ne, tiles = (ne + MF_E - 1) / MF_E;
  cudaStream_t st = c->stream;
  uint8_t *aff = nullptr;
  FCVM_CUDA(cudaMalloc((void **)&c->egeo, sizeof(double) * 10 * (size_t)ne));
  FCVM_CUDA(cudaMalloc((void **)&c->tile_affine, (size_t)tiles));
  FCVM_CUDA(cudaMalloc((void **)&aff, (size_t)ne));
  k_elem_geometry<<<grid_for(ne, 128), 128, 0, st>>>(ne, c->conn, c->xyz, c->egeo, aff);
  k_tile_affine<<<grid_for(tiles, 128), 128, 0, st>>>(ne, aff, c->tile_affine);
  std::vector<uint8_t> ht((size_t)tiles);
  FCVM_CUDA(cudaMemcpyAsync(ht.data(), c->tile_affine, (size_t)tiles, cudaMemcpyDeviceToHost, st));
  FCVM_CUDA(cudaStreamSynchronize(st));
  cudaFree(aff);
  c->n_affine_tiles = 0;
  for (uint8_t v : ht) c->n_affine_tiles += v;
  c->launches += 2;
  return FCVM_OK;
}

int matfree_set_constraints(fcvm_ctx *c) {
  if (!c->emask) FCVM_CUDA(cudaMalloc((void **)&c->emask, sizeof(uint32_t) * (size_t)c->ne));
  if (!c->ga_ticket) {
    const int64_t ngroups = (matfree_parts(c) + GA_GROUP - 1) / GA_GROUP;
    FCVM_CUDA(cudaMalloc((void **)&c->ga_ticket, sizeof(unsigned int) * (size_t)(ngroups + 1)));
    FCVM_CUDA(cudaMemsetAsync(c->ga_ticket, 0, sizeof(unsigned int) * (size_t)(ngroups + 1), c->stream));
    FCVM_CUDA(cudaMalloc((void **)&c->ga_group_part, sizeof(double) * 2 * (size_t)ngroups));
  }
  k_elem_mask<<<grid_for(c->ne, 256), 256, 0, c->stream>>>(c->ne, c->conn, c->fixmask, c->emask);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

// The assembled operator is the elastic one on the undeformed mesh and the product may be recomputed instead
// of streamed.  FCVM_MATFREE=0 keeps the assembled SpMV everywhere (comparison runs).
bool p2p_ready(const fcvm_ctx *c);

// On a partitioned mesh the per-rank products are completed by the peer-memory halo (fcvm_p2p.cu).
bool matfree_active(const fcvm_ctx *c) {
  static const bool off = getenv("FCVM_MATFREE") && atoi(getenv("FCVM_MATFREE")) == 0;
  return !off && c->matrix_elastic && (c->world == 1 || p2p_ready(c)) && c->emask != nullptr;
}


// y = K x; with sc: the early-out test of the PCG batch; dot_part / dot_part2: block partials of y.x and r.x;
// with sc_out the last block of the gather also publishes their sums (delta, gamma)
int launch_matfree(fcvm_ctx *c, const double *x, double *y, const double *sc, int rr_slot, int iters_slot, int thr_slot,
                   double *dot_part, const double *rvec, double *dot_part2, double *sc_out, int delta_slot,
                   int gamma_slot) {
  const double E = c->E, nu = c->nu;
  const double dm = E * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu);
  const double lambda = dm * (nu / (1.0 - nu));
  const double mu = dm * (0.5 * (1.0 - 2.0 * nu) / (1.0 - nu));
  const int tiles = grid_for(c->ne, MF_E);
  static const bool no_affine = getenv("FCVM_MATFREE_AFFINE") && atoi(getenv("FCVM_MATFREE_AFFINE")) == 0;
  const bool aff = c->n_affine_tiles > 0 && !no_affine;
  if (aff)
    k_elastic_apply_affine<<<tiles, MF_THREADS, 0, c->stream>>>(c->ne, c->conn, c->egeo, c->tile_affine, x, c->emask,
                                                                lambda, mu, c->elv, sc, rr_slot, iters_slot, thr_slot);
  if (!aff || c->n_affine_tiles < tiles)
    k_elastic_apply<<<tiles, MF_THREADS, 0, c->stream>>>(c->ne, c->conn, c->xyz, x, c->emask, lambda, mu, c->elv, sc,
                                                         rr_slot, iters_slot, thr_slot, aff ? c->tile_affine : nullptr);
  k_gather_apply<<<(unsigned)matfree_parts(c), GA_THREADS, 0, c->stream>>>(
      c->nn, c->n2e_ptr, c->n2e_idx, c->elv, c->fixmask, x, y, sc, rr_slot, iters_slot, thr_slot, dot_part, rvec,
      c->dof_weight, dot_part2, sc_out ? c->ga_ticket : nullptr, c->ga_group_part, sc_out, delta_slot, gamma_slot);
  c->launches += 2;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

}  // namespace fcvm

// y = K x with the elastic operator recomputed element by element (the product the PCG uses in the geometrically
// linear analysis); equals fcvm_spmv after an elastic fcvm_assemble to round-off.
extern "C" int fcvm_matfree_apply(fcvm_ctx *c, const double *x, double *y) {
  FCVM_CHECK(c && c->ne > 0 && c->have_bcs && c->emask && x && y, FCVM_E_ARG,
             "fcvm_matfree_apply: set the mesh and the constraints first / null argument");
  ProfScope ps(c, 0);
  FCVM_TRY(launch_matfree(c, x, y, nullptr, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, 0));
  return fcvm_interface_sum(c, y);
}
