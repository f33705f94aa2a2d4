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
                double *__restrict__ elv, const double *__restrict__ sc, int rr_slot, int iters_slot, int thr_slot) {
  if (sc && (sc[iters_slot] >= 0.0 || sc[rr_slot] <= sc[thr_slot])) return;      // batch already converged
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

// one thread per node: the three components of y, the constrained rows, the dot-product partials of the block
constexpr int GA_THREADS = 256;
__global__ void __launch_bounds__(GA_THREADS)
k_gather_apply(int64_t nn, const int32_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e_idx,
               const double *__restrict__ elv, const uint8_t *__restrict__ fixmask, const double *__restrict__ x,
               double *__restrict__ y, const double *__restrict__ sc, int rr_slot, int iters_slot, int thr_slot,
               double *dot_part, const double *__restrict__ rvec, const double *__restrict__ wt, double *dot_part2) {
  if (sc && (sc[iters_slot] >= 0.0 || sc[rr_slot] <= sc[thr_slot])) return;
  const int64_t n = blockIdx.x * (int64_t)GA_THREADS + threadIdx.x;
  double dsum = 0.0, rsum = 0.0;
  if (n < nn) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const int32_t b = n2e_ptr[n], eend = n2e_ptr[n + 1];
    for (int32_t k = b; k < eend; k++) {
      const double *f = elv + 3 * (int64_t)n2e_idx[k];
      s0 += f[0];
      s1 += f[1];
      s2 += f[2];
    }
    const int64_t r3 = 3 * n;
    const double x0 = x[r3], x1 = x[r3 + 1], x2 = x[r3 + 2];
    const double cnt = (double)(eend - b);
    if (fixmask[r3]) s0 = cnt * x0;
    if (fixmask[r3 + 1]) s1 = cnt * x1;
    if (fixmask[r3 + 2]) s2 = cnt * x2;
    y[r3] = s0;
    y[r3 + 1] = s1;
    y[r3 + 2] = s2;
    if (dot_part) dsum = s0 * x0 + s1 * x1 + s2 * x2;
    if (dot_part2) rsum = (wt ? wt[r3] : 1.0) * (rvec[r3] * x0 + rvec[r3 + 1] * x1 + rvec[r3 + 2] * x2);
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
  }
}

}  // namespace

namespace fcvm {

int matfree_set_constraints(fcvm_ctx *c) {
  if (!c->emask) FCVM_CUDA(cudaMalloc((void **)&c->emask, sizeof(uint32_t) * (size_t)c->ne));
  k_elem_mask<<<grid_for(c->ne, 256), 256, 0, c->stream>>>(c->ne, c->conn, c->fixmask, c->emask);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

// The assembled operator is the elastic one on the undeformed mesh and the product may be recomputed instead
// of streamed.  FCVM_MATFREE=0 keeps the assembled SpMV everywhere (comparison runs).
bool matfree_active(const fcvm_ctx *c) {
  static const bool off = getenv("FCVM_MATFREE") && atoi(getenv("FCVM_MATFREE")) == 0;
  return !off && c->matrix_elastic && c->world == 1 && c->emask != nullptr;
}

int64_t matfree_parts(const fcvm_ctx *c) { return (c->nn + GA_THREADS - 1) / GA_THREADS; }

// y = K x; with sc: the early-out test of the PCG batch; dot_part / dot_part2: block partials of y.x and r.x
int launch_matfree(fcvm_ctx *c, const double *x, double *y, const double *sc, int rr_slot, int iters_slot, int thr_slot,
                   double *dot_part, const double *rvec, double *dot_part2) {
  const double E = c->E, nu = c->nu;
  const double dm = E * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu);
  const double lambda = dm * (nu / (1.0 - nu));
  const double mu = dm * (0.5 * (1.0 - 2.0 * nu) / (1.0 - nu));
  k_elastic_apply<<<grid_for(c->ne, MF_E), MF_THREADS, 0, c->stream>>>(c->ne, c->conn, c->xyz, x, c->emask, lambda, mu,
                                                                       c->elv, sc, rr_slot, iters_slot, thr_slot);
  k_gather_apply<<<(unsigned)matfree_parts(c), GA_THREADS, 0, c->stream>>>(
      c->nn, c->n2e_ptr, c->n2e_idx, c->elv, c->fixmask, x, y, sc, rr_slot, iters_slot, thr_slot, dot_part, rvec,
      c->dof_weight, dot_part2);
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
  FCVM_TRY(launch_matfree(c, x, y, nullptr, 0, 0, 0, nullptr, nullptr, nullptr));
  return fcvm_interface_sum(c, y);
}
