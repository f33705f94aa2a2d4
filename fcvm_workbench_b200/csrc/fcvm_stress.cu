// Gauss-point stress update, von Mises radial return and internal-force vector.
// Replaces update_stress_load (fcVM.py:2196-2464), vmises_original_optimised
// (fcVM.py:2468-2492), update_PEEQ_CSR (fcVM.py:2084-2137) and mapStresses (fcVM.py:2496-2554).
#include "fcvm_common.cuh"
#include "fcvm_reduce.cuh"

using namespace fcvm;

namespace {

struct Material {
  double d_diag, d_off, d_shear;   // Hooke matrix entries (fcVM.py:574-582)
  double G, H;                     // shear modulus, hardening modulus (fcVM.py:2231-2234)
};

__host__ Material make_material(double E, double nu, double Et_E) {
  Material m;
  double dm = E * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu);
  double od = nu / (1.0 - nu);
  double sd = 0.5 * (1.0 - 2.0 * nu) / (1.0 - nu);
  m.d_diag = 1.0 * dm;
  m.d_off = od * dm;
  m.d_shear = sd * dm;
  m.G = E / 2.0 / (1 + nu);
  if (Et_E > 0.95) Et_E = 0.95;
  double Et = Et_E * E;
  m.H = Et / (1.0 - Et_E);
  return m;
}

#ifndef SU_MINBLOCKS
#define SU_MINBLOCKS 6
#endif
constexpr int SU_E = 32;            // elements per block: one lane per element
constexpr int SU_THREADS = 128;     // four warps: warp w integrates Gauss point w of the 32 elements
constexpr int SU_ROW = 35;          // doubles per element of the nodal staging (30 used)
constexpr int SU_PAD = 33;          // row stride of the force staging (conflict-free in both phases)

// One Gauss point of one element (lane = element, GP = warp): (convected) old stress, elastic test
// stress, radial return; leaves this point's share of the element force vector in shared memory.
template <bool LD>
__device__ __forceinline__ void gauss_point_T(int64_t ne, int64_t e, int GP, bool live, const double (&xsi)[3][3],
                                              double xsj, const double (&g)[3][3], double (&sc)[6], double sy,
                                              const Material &m, double *__restrict__ sig_new,
                                              double *__restrict__ sig_test, uint8_t *__restrict__ pgp,
                                              double (&T)[3][3]) {
  const double deps0 = g[0][0], deps1 = g[1][1], deps2 = g[2][2];
  const double deps3 = g[0][1] + g[1][0], deps4 = g[0][2] + g[2][0], deps5 = g[1][2] + g[2][1];
  if (LD) {
    // convected stress sigma <- F sigma F^T / det F with F = I + grad(du)   (fcVM.py:2383-2429)
    double Fd[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Fd[i][j] = g[i][j] + (i == j ? 1.0 : 0.0);
    double rr = (Fd[0][0] * Fd[1][1] * Fd[2][2] - Fd[0][0] * Fd[1][2] * Fd[2][1] + Fd[0][2] * Fd[1][0] * Fd[2][1] -
                 Fd[0][2] * Fd[1][1] * Fd[2][0] + Fd[0][1] * Fd[1][2] * Fd[2][0] - Fd[0][1] * Fd[1][0] * Fd[2][2]);
    rr = 1.0 / rr;
    const double S[3][3] = {{sc[0], sc[3], sc[4]}, {sc[3], sc[1], sc[5]}, {sc[4], sc[5], sc[2]}};
    double FS[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int l = 0; l < 3; l++) FS[i][l] = Fd[i][0] * S[0][l] + Fd[i][1] * S[1][l] + Fd[i][2] * S[2][l];
    auto con = [&](int i, int k) { return rr * (FS[i][0] * Fd[k][0] + FS[i][1] * Fd[k][1] + FS[i][2] * Fd[k][2]); };
    sc[0] = con(0, 0); sc[1] = con(1, 1); sc[2] = con(2, 2);
    sc[3] = con(0, 1); sc[4] = con(0, 2); sc[5] = con(1, 2);
  }
  // elastic test stress (fcVM.py:2434-2441)
  double st0 = sc[0] + m.d_diag * deps0 + m.d_off * deps1 + m.d_off * deps2;
  double st1 = sc[1] + m.d_off * deps0 + m.d_diag * deps1 + m.d_off * deps2;
  double st2 = sc[2] + m.d_off * deps0 + m.d_off * deps1 + m.d_diag * deps2;
  double st3 = sc[3] + m.d_shear * deps3;
  double st4 = sc[4] + m.d_shear * deps4;
  double st5 = sc[5] + m.d_shear * deps5;
  if (live) {
    __stcs(&sig_test[((int64_t)0 * 4 + GP) * ne + e], st0);
    __stcs(&sig_test[((int64_t)1 * 4 + GP) * ne + e], st1);
    __stcs(&sig_test[((int64_t)2 * 4 + GP) * ne + e], st2);
    __stcs(&sig_test[((int64_t)3 * 4 + GP) * ne + e], st3);
    __stcs(&sig_test[((int64_t)4 * 4 + GP) * ne + e], st4);
    __stcs(&sig_test[((int64_t)5 * 4 + GP) * ne + e], st5);
  }
  // radial return to the von Mises surface (fcVM.py:2468-2492)
  const double p = (st0 + st1 + st2) / 3.0;
  st0 -= p; st1 -= p; st2 -= p;
  const double svm = sqrt(1.5 * (st0 * st0 + st1 * st1 + st2 * st2) + 3.0 * (st3 * st3 + st4 * st4 + st5 * st5));
  double fac = 1.0;
  uint8_t pp = 0;
  if (!(sy > svm)) {
    fac = (1.0 - (1.0 - sy / svm) * 3.0 * m.G / (m.H + 3 * m.G));
    pp = 1;
  }
  const double sxx = fac * st0 + p, syy = fac * st1 + p, szz = fac * st2 + p;
  const double sxy = fac * st3, szx = fac * st4, syz = fac * st5;
  if (live) {
    __stcs(&sig_new[((int64_t)0 * 4 + GP) * ne + e], sxx);
    __stcs(&sig_new[((int64_t)1 * 4 + GP) * ne + e], syy);
    __stcs(&sig_new[((int64_t)2 * 4 + GP) * ne + e], szz);
    __stcs(&sig_new[((int64_t)3 * 4 + GP) * ne + e], sxy);
    __stcs(&sig_new[((int64_t)4 * 4 + GP) * ne + e], szx);
    __stcs(&sig_new[((int64_t)5 * 4 + GP) * ne + e], syz);
    pgp[(int64_t)GP * ne + e] = pp;
  }
  // element force of this Gauss point: F[k][i] = w|J| sum_mm S[i][mm] dshpg[mm][k]   (fcVM.py:2448-2454)
  const double w = GP_W * fabs(xsj);
  const double S[3][3] = {{sxx, sxy, szx}, {sxy, syy, syz}, {szx, syz, szz}};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[i][j] = w * (S[i][0] * xsi[j][0] + S[i][1] * xsi[j][1] + S[i][2] * xsi[j][2]);
}

template <bool LD>
__device__ __forceinline__ void gauss_point(int64_t ne, int64_t e, int GP, bool live, const GPCoef &cf,
                                            const double (&xsi)[3][3], double xsj, const double (&g)[3][3],
                                            double (&sc)[6], double sy, const Material &m,
                                            double *__restrict__ sig_new, double *__restrict__ sig_test,
                                            uint8_t *__restrict__ pgp, double *sF) {
  double T[3][3];
  gauss_point_T<LD>(ne, e, GP, live, xsi, xsj, g, sc, sy, m, sig_new, sig_test, pgp, T);
  store_gradient_tile(cf, T, sF, SU_PAD);
}

// Block = 32 consecutive elements x 4 Gauss points.
//   phase 0  coordinates and displacement increments of the 320 (element, node) pairs are gathered
//            once into shared memory
//   phase 1  warp w = Gauss point w (one code path, ten run-time coefficients), lane = element:
//            every Gauss-point array is read and written as full 256-byte lines
//   phase 2  the 960 entries of the 32 element force vectors are summed over the Gauss points in
//            order 0..3 (the reference's order, fcVM.py:2300) and written as one contiguous run
template <bool LD>
__global__ void __launch_bounds__(SU_THREADS, LD ? 5 : SU_MINBLOCKS)
k_stress_update(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                const double *__restrict__ disp, const double *__restrict__ du, Material m,
                const double *__restrict__ sig_old, const double *__restrict__ sig_yield, double yield_scale,
                double *__restrict__ sig_new, double *__restrict__ sig_test, uint8_t *__restrict__ pgp,
                double *__restrict__ elv, int tile0) {
  __shared__ double smem[4 * 30 * SU_PAD];        // nodal staging [2][32][35], then force staging [4][30][33]
  double *sX = smem, *sU = smem + SU_E * SU_ROW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t e0 = ((int64_t)blockIdx.x + tile0) * SU_E;
  const bool live = e0 + lane < ne;
  const int64_t e = min(e0 + lane, ne - 1);
  // the Gauss-point state is requested first, so that its HBM latency overlaps the gather and the kinematics
  double sc[6];
#pragma unroll
  for (int c = 0; c < 6; c++) sc[c] = __ldcs(&sig_old[((int64_t)c * 4 + warp) * ne + e]);
  const double sy = yield_scale * __ldcs(&sig_yield[(int64_t)warp * ne + e]);
  {
    // item q = 3*(32*node + element) + component, 960 per array: the three components of a node sit in
    // adjacent lanes, so a warp-wide load touches ~11 nodal 24-byte segments instead of 32 separate
    // ones.  All index loads are issued before the first dependent load, all loads before the first
    // store (two exposed latencies).  Staging layout [element][SU_ROW]: conflict-free for these
    // stores (3*el + i distinct) and for the per-element reads of phase 1 (stride 35 is odd).
    constexpr int NQ = (30 * SU_E + SU_THREADS - 1) / SU_THREADS;
    int64_t d[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = min(tid + r * SU_THREADS, 30 * SU_E - 1);
      const int p = q / 3;
      d[r] = 3 * (int64_t)conn[(int64_t)(p >> 5) * ne + min(e0 + (p & 31), ne - 1)] + (q - 3 * p);
    }
    double xv[NQ], uv[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      xv[r] = xyz[d[r]];
      if (LD) xv[r] += disp[d[r]];               // updated geometry (fcVM.py:2256-2260)
      uv[r] = du[d[r]];
    }
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = tid + r * SU_THREADS;
      if (q < 30 * SU_E) {
        const int p = q / 3;
        const int at = (p & 31) * SU_ROW + 3 * (p >> 5) + (q - 3 * p);
        sX[at] = xv[r];
        sU[at] = uv[r];
      }
    }
  }
  __syncthreads();
  const GPCoef cf = gp_coef(warp);
  double xsi[3][3], g[3][3], xsj;
  {
    double xs[3][3], Hl[3][3];
    local_gradient_tile(cf, sX + lane * SU_ROW, 1, xs);
    xsj = invert_jacobian(xs, xsi);
    local_gradient_tile(cf, sU + lane * SU_ROW, 1, Hl);         // Hl[i][j] = sum_k du_k[i] dN[j][k]
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int mm = 0; mm < 3; mm++) g[i][mm] = Hl[i][0] * xsi[0][mm] + Hl[i][1] * xsi[1][mm] + Hl[i][2] * xsi[2][mm];
  }
  __syncthreads();      // every warp has consumed the staged nodal data: the force staging may overwrite it
  gauss_point<LD>(ne, e, warp, live, cf, xsi, xsj, g, sc, sy, m, sig_new, sig_test, pgp,
                  smem + (warp * 30) * SU_PAD + lane);
  __syncthreads();
  const int nlive = (int)min((int64_t)SU_E, ne - e0) * 30;
  double *out = elv + 30 * e0;
  for (int idx = tid; idx < nlive; idx += SU_THREADS) {
    const int el = idx / 30, k3 = idx - 30 * el;
    const double *f = smem + k3 * SU_PAD + el;
    out[idx] = ((f[0] + f[30 * SU_PAD]) + f[60 * SU_PAD]) + f[90 * SU_PAD];
  }
}

// Variant with two Gauss points per thread: block = 32 elements x 2 warps, warp p integrates points 2p and
// 2p+1.  The staged nodal values are read once for both points, the two force contributions are added in
// registers, so half as much data crosses the shared-memory pipe (the unit that bounds the four-warp
// kernel); the price is twice the work and ~2x the registers per thread.
constexpr int SP_THREADS = 64;
template <bool LD>
__global__ void __launch_bounds__(SP_THREADS, 6)
k_stress_update_pair(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                     const double *__restrict__ disp, const double *__restrict__ du, Material m,
                     const double *__restrict__ sig_old, const double *__restrict__ sig_yield, double yield_scale,
                     double *__restrict__ sig_new, double *__restrict__ sig_test, uint8_t *__restrict__ pgp,
                     double *__restrict__ elv, int tile0) {
  __shared__ double smem[2 * SU_E * SU_ROW];      // nodal staging [2][32][35], then force staging [2][30][33]
  double *sX = smem, *sU = smem + SU_E * SU_ROW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t e0 = ((int64_t)blockIdx.x + tile0) * SU_E;
  const bool live = e0 + lane < ne;
  const int64_t e = min(e0 + lane, ne - 1);
  const int gpa = 2 * warp, gpb = gpa + 1;
  // state of the first point is requested before the gather
  double sc[6];
#pragma unroll
  for (int c = 0; c < 6; c++) sc[c] = __ldcs(&sig_old[((int64_t)c * 4 + gpa) * ne + e]);
  double sy = yield_scale * __ldcs(&sig_yield[(int64_t)gpa * ne + e]);
#pragma unroll
  for (int half = 0; half < 2; half++) {
    constexpr int NQ = 8;                          // 960 items / 64 threads = 15 = 8 + 7
    int64_t d[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = min(tid + (half * NQ + r) * SP_THREADS, 30 * SU_E - 1);
      const int p = q / 3;
      d[r] = 3 * (int64_t)conn[(int64_t)(p >> 5) * ne + min(e0 + (p & 31), ne - 1)] + (q - 3 * p);
    }
    double xv[NQ], uv[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      xv[r] = xyz[d[r]];
      if (LD) xv[r] += disp[d[r]];
      uv[r] = du[d[r]];
    }
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = tid + (half * NQ + r) * SP_THREADS;
      if (q < 30 * SU_E) {
        const int p = q / 3;
        const int at = (p & 31) * SU_ROW + 3 * (p >> 5) + (q - 3 * p);
        sX[at] = xv[r];
        sU[at] = uv[r];
      }
    }
  }
  __syncthreads();
  const GPCoef ca = gp_coef(gpa), cb = gp_coef(gpb);
  double xsiA[3][3], xsiB[3][3], gA[3][3], gB[3][3], xsjA, xsjB;
  {
    double xa[3][3], xb[3][3];
    local_gradient_tile2(ca, cb, sX + lane * SU_ROW, 1, xa, xb);
    xsjA = invert_jacobian(xa, xsiA);
    xsjB = invert_jacobian(xb, xsiB);
    local_gradient_tile2(ca, cb, sU + lane * SU_ROW, 1, xa, xb);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int mm = 0; mm < 3; mm++) {
        gA[i][mm] = xa[i][0] * xsiA[0][mm] + xa[i][1] * xsiA[1][mm] + xa[i][2] * xsiA[2][mm];
        gB[i][mm] = xb[i][0] * xsiB[0][mm] + xb[i][1] * xsiB[1][mm] + xb[i][2] * xsiB[2][mm];
      }
  }
  __syncthreads();      // staged nodal data consumed: the force staging may overwrite it
  double F[30], T[3][3];
  gauss_point_T<LD>(ne, e, gpa, live, xsiA, xsjA, gA, sc, sy, m, sig_new, sig_test, pgp, T);
  gradient_to_regs<false>(ca, T, F);
  // (requesting the second point's state ahead of the first point's arithmetic measured slower: registers)
#pragma unroll
  for (int c = 0; c < 6; c++) sc[c] = __ldcs(&sig_old[((int64_t)c * 4 + gpb) * ne + e]);
  sy = yield_scale * __ldcs(&sig_yield[(int64_t)gpb * ne + e]);
  gauss_point_T<LD>(ne, e, gpb, live, xsiB, xsjB, gB, sc, sy, m, sig_new, sig_test, pgp, T);
  gradient_to_regs<true>(cb, T, F);
  double *sF = smem + (warp * 30) * SU_PAD + lane;
#pragma unroll
  for (int k = 0; k < 30; k++) sF[k * SU_PAD] = F[k];
  __syncthreads();
  const int nlive = (int)min((int64_t)SU_E, ne - e0) * 30;
  double *out = elv + 30 * e0;
  for (int idx = tid; idx < nlive; idx += SP_THREADS) {
    const int el = idx / 30, k3 = idx - 30 * el;
    const double *f = smem + k3 * SU_PAD + el;
    out[idx] = f[0] + f[30 * SU_PAD];
  }
}

}  // namespace

// Deterministic assembly of nodal vectors: dof d = 3*node+c sums the element vectors of the
// elements around the node in ascending element order -- the same order in which the
// reference's element loop accumulates into qin (fcVM.py:2456-2462).  No atomics.
__global__ void k_node_gather(int64_t nn, const int32_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e_idx,
                              const double *__restrict__ elv, double *__restrict__ out, int accumulate) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= 3 * nn) return;
  const int64_t n = d / 3;
  const int c = (int)(d - 3 * n);
  double s = accumulate ? out[d] : 0.0;
  const int32_t b = n2e_ptr[n], eend = n2e_ptr[n + 1];
  for (int32_t k = b; k < eend; k++) s += elv[3 * (int64_t)n2e_idx[k] + c];
  out[d] = s;
}

namespace fcvm {
int launch_node_gather(fcvm_ctx *c, double *out, int accumulate) {
  ProfScope ps(c, 2);
  k_node_gather<<<grid_for(3 * c->nn, 256), 256, 0, c->stream>>>(c->nn, c->n2e_ptr, c->n2e_idx, c->elv, out,
                                                                 accumulate);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}
}  // namespace fcvm

namespace fcvm {
// the Gauss-point pass over the 32-element tiles [tile0, tile0 + ntiles): element vectors to the scratch
int launch_stress_tiles(fcvm_ctx *c, const double *disp_new, const double *du, double Et_E, int LD, double yield_scale,
                        int64_t tile0, int64_t ntiles) {
  const Material m = make_material(c->E, c->nu, Et_E);
  const double *so = (const double *)c->buf[FCVM_BUF_SIG_OLD], *sy = (const double *)c->buf[FCVM_BUF_SIG_YIELD];
  double *sn = (double *)c->buf[FCVM_BUF_SIG_NEW], *stt = (double *)c->buf[FCVM_BUF_SIG_TEST];
  uint8_t *pg = (uint8_t *)c->buf[FCVM_BUF_PGP];
  ProfScope ps(c, 1);
  const int grid = (int)ntiles, t0 = (int)tile0;
  // default: two Gauss points per thread (0.265 ms at 1M elements); FCVM_STRESS_PAIR=0 selects the
  // four-warp kernel (0.277 ms) for comparison
  static const bool pair = !(getenv("FCVM_STRESS_PAIR") && atoi(getenv("FCVM_STRESS_PAIR")) == 0);
  if (pair) {
    if (LD)
      k_stress_update_pair<true><<<grid, SP_THREADS, 0, c->stream>>>(c->ne, c->conn, c->xyz, disp_new, du, m, so, sy,
                                                                   yield_scale, sn, stt, pg, c->elv, t0);
    else
      k_stress_update_pair<false><<<grid, SP_THREADS, 0, c->stream>>>(c->ne, c->conn, c->xyz, disp_new, du, m, so, sy,
                                                                    yield_scale, sn, stt, pg, c->elv, t0);
  } else if (LD)
    k_stress_update<true><<<grid, SU_THREADS, 0, c->stream>>>(c->ne, c->conn, c->xyz, disp_new, du, m, so, sy,
                                                              yield_scale, sn, stt, pg, c->elv, t0);
  else
    k_stress_update<false><<<grid, SU_THREADS, 0, c->stream>>>(c->ne, c->conn, c->xyz, disp_new, du, m, so, sy,
                                                               yield_scale, sn, stt, pg, c->elv, t0);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}
}  // namespace fcvm

extern "C" int fcvm_update_stress_load(fcvm_ctx *c, const double *disp_new, const double *du, double *qin,
                                       double Et_E, int LD, double yield_scale) {
  FCVM_CHECK(c && c->ne > 0 && du && qin, FCVM_E_ARG, "fcvm_update_stress_load: null argument / no mesh");
  FCVM_CHECK(!LD || disp_new, FCVM_E_ARG, "fcvm_update_stress_load: LD needs disp_new");
  FCVM_TRY(launch_stress_tiles(c, disp_new, du, Et_E, LD, yield_scale, 0, grid_for(c->ne, SU_E)));
  FCVM_TRY(launch_node_gather(c, qin, 0));
  return fcvm_interface_sum(c, qin);
}

// ---- update_PEEQ_CSR -------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
k_peeq_csr(int64_t ne, double G, double H, double Et, double alpha, const double *__restrict__ sig_test,
           const double *__restrict__ sig_new, double *__restrict__ sig_yield, double *__restrict__ peeq,
           double *__restrict__ csr, double *__restrict__ triax, double *__restrict__ pressure,
           double *__restrict__ sigmises, double *__restrict__ ecr, double *red_part, int64_t *arg_part,
           unsigned int *counter, double *out7, int64_t *arg_out) {
  const int64_t n = 4 * ne;
  double best = -1.0, best_peeq = 0.0;
  int64_t best_ref = INT64_MAX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i % ne;
    const int ip = (int)(i / ne);
    double t[6], s[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
      t[c] = sig_test[((int64_t)c * 4 + ip) * ne + e];
      s[c] = sig_new[((int64_t)c * 4 + ip) * ne + e];
    }
    const double p_t = (t[0] + t[1] + t[2]) / 3.0, p_n = (s[0] + s[1] + s[2]) / 3.0;
    t[0] -= p_t; t[1] -= p_t; t[2] -= p_t;
    s[0] -= p_n; s[1] -= p_n; s[2] -= p_n;
    const double smt = sqrt(1.5 * (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) +
                            3.0 * (t[3] * t[3] + t[4] * t[4] + t[5] * t[5]));
    const double smn = sqrt(1.5 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) +
                            3.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]));
    double sy = sig_yield[i], pq = peeq[i], DL = 0.0;
    if (smt > sy) {
      DL = (smt - sy) / (3.0 * G + H);
      pq += DL;
      sy += Et * DL;
      peeq[i] = pq;
      sig_yield[i] = sy;
    }
    const double T = p_n / sy;
    pressure[i] = p_n;
    sigmises[i] = smn;
    triax[i] = T;
    double ce = alpha * exp(-1.5 * T);
    if (ce < 1.0e-6) ce = 1.0e-6;
    ecr[i] = ce;
    const double cs = csr[i] + DL / ce;
    csr[i] = cs;
    const int64_t ref = 4 * e + ip;               // Gauss-point number in the reference layout
    if (cs > best || (cs == best && ref < best_ref)) {
      best = cs;
      best_ref = ref;
    }
    best_peeq = fmax(best_peeq, pq);
  }
  // block argmax (first maximum in reference numbering, like np.argmax) and max(peeq)
  __shared__ double sv[RED_THREADS / 32], sp[RED_THREADS / 32];
  __shared__ int64_t si[RED_THREADS / 32];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_down_sync(0xffffffffu, best, o);
    int64_t oi = __shfl_down_sync(0xffffffffu, best_ref, o);
    double op = __shfl_down_sync(0xffffffffu, best_peeq, o);
    if (ov > best || (ov == best && oi < best_ref)) { best = ov; best_ref = oi; }
    best_peeq = fmax(best_peeq, op);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = best_ref; sp[warp] = best_peeq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RED_THREADS / 32; w++) {
      if (sv[w] > best || (sv[w] == best && si[w] < best_ref)) { best = sv[w]; best_ref = si[w]; }
      best_peeq = fmax(best_peeq, sp[w]);
    }
    red_part[blockIdx.x] = best;
    red_part[RED_BLOCKS + blockIdx.x] = best_peeq;
    arg_part[blockIdx.x] = best_ref;
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    best = -1.0; best_ref = INT64_MAX; best_peeq = 0.0;
    for (int b = 0; b < RED_BLOCKS; b++) {
      const double v = __ldcg(&red_part[b]);
      const int64_t r = __ldcg(&arg_part[b]);
      if (v > best || (v == best && r < best_ref)) { best = v; best_ref = r; }
      best_peeq = fmax(best_peeq, __ldcg(&red_part[RED_BLOCKS + b]));
    }
    const int64_t i = (best_ref & 3) * ne + (best_ref >> 2);
    out7[0] = best;
    out7[1] = pressure[i];
    out7[2] = sigmises[i];
    out7[3] = triax[i];
    out7[4] = ecr[i];
    out7[5] = peeq[i];
    out7[6] = best_peeq;
    arg_out[0] = best_ref;
    *counter = 0u;
  }
}

extern "C" int fcvm_update_peeq_csr(fcvm_ctx *c, double ultimate_strain, double Et_E, int64_t *argmax_gp,
                                    double *out7) {
  FCVM_CHECK(c && c->ne > 0, FCVM_E_ARG, "fcvm_update_peeq_csr: no mesh");
  const double G = c->E / 2.0 / (1 + c->nu);
  if (Et_E > 0.95) Et_E = 0.95;
  const double Et = Et_E * c->E;
  const double H = Et / (1.0 - Et_E);
  if (ultimate_strain == 0.0) ultimate_strain = 1.0e12;
  const double alpha = sqrt(exp(1.0)) * ultimate_strain;
  if (!c->d_arg_part) {
    FCVM_CUDA(cudaMalloc((void **)&c->d_arg_part, sizeof(int64_t) * RED_BLOCKS));
  }
  k_peeq_csr<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(
      c->ne, G, H, Et, alpha, (const double *)c->buf[FCVM_BUF_SIG_TEST], (const double *)c->buf[FCVM_BUF_SIG_NEW],
      (double *)c->buf[FCVM_BUF_SIG_YIELD], (double *)c->buf[FCVM_BUF_PEEQ], (double *)c->buf[FCVM_BUF_CSR],
      (double *)c->buf[FCVM_BUF_TRIAX], (double *)c->buf[FCVM_BUF_PRESSURE], (double *)c->buf[FCVM_BUF_SIGMISES],
      (double *)c->buf[FCVM_BUF_ECR], c->red_part, c->d_arg_part, c->red_counter, c->red_out, c->d_arg);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  FCVM_CUDA(cudaMemcpyAsync(c->h_scalars, c->red_out, sizeof(double) * 7, cudaMemcpyDeviceToHost, c->stream));
  FCVM_CUDA(cudaMemcpyAsync(c->h_arg, c->d_arg, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  if (out7) memcpy(out7, c->h_scalars, sizeof(double) * 7);
  if (argmax_gp) *argmax_gp = c->h_arg[0];
  return FCVM_OK;
}

__global__ void k_scale_step(int64_t n, double fac, const double *__restrict__ so, double *sn, double *st) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double o = so[i];
  sn[i] = o + fac * (sn[i] - o);
  st[i] = o + fac * (st[i] - o);
}

extern "C" int fcvm_scale_step_stress(fcvm_ctx *c, double fac) {
  FCVM_CHECK(c && c->ne > 0, FCVM_E_ARG, "fcvm_scale_step_stress: no mesh");
  const int64_t n = 24 * c->ne;
  k_scale_step<<<grid_for(n, 256), 256, 0, c->stream>>>(n, fac, (const double *)c->buf[FCVM_BUF_SIG_OLD],
                                                        (double *)c->buf[FCVM_BUF_SIG_NEW],
                                                        (double *)c->buf[FCVM_BUF_SIG_TEST]);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

// ---- mapStresses ---------------------------------------------------------------------------------
// corner nodes: the Gauss-point values of the adjacent elements, in ascending element order
__global__ void k_map_corners(int64_t ne, int64_t nn, int averaged, double sig_yield,
                              const int32_t *__restrict__ n2e_ptr, const int32_t *__restrict__ n2e_idx,
                              const int16_t *__restrict__ noce, const double *__restrict__ sig,
                              const double *__restrict__ peeq, const double *__restrict__ svm,
                              const double *__restrict__ csr, double *t_stress, double *t_peeq, double *t_csr,
                              double *t_svm, double *t_triax) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  double s[6] = {0, 0, 0, 0, 0, 0}, a = 0, b = 0, v = 0, t = 0;
  const double cnt = noce ? (double)noce[n] : (double)(n2e_ptr[n + 1] - n2e_ptr[n]);
  for (int32_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; k++) {
    const int32_t idx = n2e_idx[k];
    const int64_t e = idx / 10;
    const int j = idx - 10 * (int)e;
    if (j >= 4) continue;
    double g[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
      g[c] = sig[((int64_t)c * 4 + j) * ne + e];
      s[c] += g[c] / cnt;
    }
    const double tr = (g[0] + g[1] + g[2]) / 3.0 / sig_yield;
    const int64_t gi = (int64_t)j * ne + e;
    if (averaged) {
      a += peeq[gi] / cnt; b += csr[gi] / cnt; v += svm[gi] / cnt; t += tr / cnt;
    } else {
      a = fmax(a, peeq[gi]); b = fmax(b, csr[gi]); v = fmax(v, svm[gi]); t = fmax(t, tr);
    }
  }
#pragma unroll
  for (int c = 0; c < 6; c++) t_stress[6 * n + c] = s[c];
  t_peeq[n] = a; t_csr[n] = b; t_svm[n] = v; t_triax[n] = t;
}

// mid-side nodes: mean of the two corner nodes of their edge (fcVM.py:2500-2552)
__global__ void k_map_mids(int64_t ne, const int32_t *__restrict__ conn, double *t_stress, double *t_peeq,
                           double *t_csr, double *t_svm, double *t_triax) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 6 * ne) return;
  const int64_t e = i / 6;
  const int mI = (int)(i - 6 * e);
  const int ca[6] = {0, 1, 0, 0, 1, 2}, cb[6] = {1, 2, 2, 3, 3, 3};
  const int64_t na = conn[(int64_t)ca[mI] * ne + e], nb = conn[(int64_t)cb[mI] * ne + e];
  const int64_t nm = conn[(int64_t)(4 + mI) * ne + e];
  // every element around the edge writes the same value: benign
#pragma unroll
  for (int c = 0; c < 6; c++) t_stress[6 * nm + c] = 0.5 * t_stress[6 * na + c] + 0.5 * t_stress[6 * nb + c];
  t_peeq[nm] = 0.5 * t_peeq[na] + 0.5 * t_peeq[nb];
  t_csr[nm] = 0.5 * t_csr[na] + 0.5 * t_csr[nb];
  t_svm[nm] = 0.5 * t_svm[na] + 0.5 * t_svm[nb];
  t_triax[nm] = 0.5 * t_triax[na] + 0.5 * t_triax[nb];
}

extern "C" int fcvm_map_stresses(fcvm_ctx *c, int averaged, double sig_yield, const int16_t *noce,
                                 double *tet10stress, double *tet10peeq, double *tet10csr, double *tet10svm,
                                 double *tet10triax) {
  FCVM_CHECK(c && c->ne > 0 && tet10stress && tet10peeq && tet10csr && tet10svm && tet10triax, FCVM_E_ARG,
             "fcvm_map_stresses: null argument / no mesh");
  const int64_t nn = c->nn;
  double *d;
  int16_t *dn = nullptr;
  FCVM_CUDA(cudaMalloc((void **)&d, sizeof(double) * 10 * nn));
  FCVM_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 10 * nn, c->stream));
  if (noce) {
    FCVM_CUDA(cudaMalloc((void **)&dn, sizeof(int16_t) * nn));
    FCVM_CUDA(cudaMemcpyAsync(dn, noce, sizeof(int16_t) * nn, cudaMemcpyHostToDevice, c->stream));
  }
  double *ts = d, *tp = d + 6 * nn, *tc = d + 7 * nn, *tv = d + 8 * nn, *tt = d + 9 * nn;
  k_map_corners<<<grid_for(nn, 128), 128, 0, c->stream>>>(
      c->ne, nn, averaged, sig_yield, c->n2e_ptr, c->n2e_idx, dn, (const double *)c->buf[FCVM_BUF_SIG_NEW],
      (const double *)c->buf[FCVM_BUF_PEEQ], (const double *)c->buf[FCVM_BUF_SIGMISES],
      (const double *)c->buf[FCVM_BUF_CSR], ts, tp, tc, tv, tt);
  k_map_mids<<<grid_for(6 * c->ne, 256), 256, 0, c->stream>>>(c->ne, c->conn, ts, tp, tc, tv, tt);
  c->launches += 2;
  FCVM_CUDA(cudaGetLastError());
  int rc = fcvm_d2h(c, tet10stress, ts, sizeof(double) * 6 * nn);
  if (rc == FCVM_OK) rc = fcvm_d2h(c, tet10peeq, tp, sizeof(double) * nn);
  if (rc == FCVM_OK) rc = fcvm_d2h(c, tet10csr, tc, sizeof(double) * nn);
  if (rc == FCVM_OK) rc = fcvm_d2h(c, tet10svm, tv, sizeof(double) * nn);
  if (rc == FCVM_OK) rc = fcvm_d2h(c, tet10triax, tt, sizeof(double) * nn);
  cudaFree(d);
  if (dn) cudaFree(dn);
  return rc;
}
