// Element stiffness by Gauss-point integration and deterministic assembly.
// Replaces calcGSM (fcVM.py:620-816) and the nstep>1 branch of calcTSM (fcVM.py:819-1079),
// scipy's COO->CSC conversion (fcVM.py:1111) and the constraint bookkeeping `modf`.
//
// Flow:  k_elem_stiffness  -> cooK[55][ne][9]   (3x3 blocks of the lower block triangle)
//        k_coo_reduce      -> block-SELL values (fixed contribution lists, ascending element
//                             order, no floating-point atomics)
//        SpMV with fixval  -> modf;  k_apply_constraints -> rows/cols of prescribed dofs
//        k_block_inverse   -> block-Jacobi preconditioner
#include "fcvm_common.cuh"

using namespace fcvm;

namespace fcvm {
int deflation_build(fcvm_ctx *c);
int launch_node_gather(fcvm_ctx *c, double *out, int accumulate);
int launch_spmv(fcvm_ctx *c, const double *x, double *y);
}  // namespace fcvm

namespace {

constexpr int KE_E = 32;            // elements per block: one lane per element
constexpr int KE_THREADS = 128;     // four warps
constexpr int KE_ROW = 35;          // doubles per element of the coordinate staging (30 used)
constexpr int KE_PAD = 33;          // row stride of the gradient tiles

struct TangentArgs {
  const double *sig_old;   // SoA (tangent: stress at the start of the step; buckling: the elastic stress state)
  const uint8_t *pgp;      // SoA
  double G, H;
  const uint32_t *emask;   // buckling: prescribed dofs of every element (bit 3k+c)
  double sigma;            // buckling: shift
};

// what k_elem_stiffness integrates
//   KE_ELASTIC     B^T D B                                   calcGSM, fcVM.py:620-816
//   KE_TANGENT     B^T (D - pmat) B at plastic points        calcTSM nstep > 1, fcVM.py:983-999
//   KE_BUCKLING_M  K - sigma G of the linear buckling analysis (calcTSM nstep == 1, fcVM.py:1002-1006, 1063-1073):
//                  K = B^T D B NOT eliminated, diagonal entries of prescribed dofs times 100; G = -nsm
//   KE_BUCKLING_G  G = -nsm, nsm = GM^T (sigma_ij x I3) GM the geometric stiffness of the stress state
enum { KE_ELASTIC = 0, KE_TANGENT = 1, KE_BUCKLING_M = 2, KE_BUCKLING_G = 3 };

// rows a of the lower block triangle handled by each warp (a+1 blocks per row): 14, 14, 14, 13 blocks
__constant__ int c_rows[4][4] = {{9, 3, -1, -1}, {8, 4, -1, -1}, {7, 5, -1, -1}, {6, 2, 1, 0}};

// Block = 32 consecutive elements.
//   phase 0  nodal coordinates (+ displacements for the updated geometry) gathered once into shared memory
//   phase A  warp w = Gauss point w, lane = element: Jacobian, then the 30 scaled shape-function
//            gradients sqrt(w|J|) dN_k/dx_m of that point into a shared tile; weights and (tangent)
//            the deviator / pmat factor of plastic points beside them
//   phase B  the 55 blocks K_ab (a >= b) = lambda P + mu P^T + mu tr(P) I (- plastic correction),
//            P = sum_gp g_a g_b^T, split over the warps by rows of equal work; each block row of 32
//            elements x 9 entries leaves through a per-warp staging tile as one contiguous 2304-byte run
template <int MODE>
__global__ void __launch_bounds__(KE_THREADS, 4)
k_elem_stiffness(int64_t ne, const int32_t *__restrict__ conn, const double *__restrict__ xyz,
                 const double *__restrict__ disp, double lambda, double mu, TangentArgs ta, double gx, double gy,
                 double gz, double *__restrict__ cooK, double *__restrict__ elv) {
  __shared__ double sG[4 * 30 * KE_PAD];          // coordinate staging [32][35] first, then gradient tiles [4][30][33]
  __shared__ double sO[4][KE_E * 9];              // per-warp output staging
  __shared__ double sW[4][KE_E];                  // w|J| per Gauss point
  constexpr bool TANGENT = MODE == KE_TANGENT, GEOM = MODE >= KE_BUCKLING_M;
  __shared__ double sS[(TANGENT || GEOM) ? 4 * 7 * KE_E : 1];   // [gp][6 deviator (tangent) / stress (buckling) comps + factor][element]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * KE_E;
  const int64_t e = min(e0 + lane, ne - 1);
  {
    constexpr int NQ = (30 * KE_E + KE_THREADS - 1) / KE_THREADS;
    int64_t d[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = min(tid + r * KE_THREADS, 30 * KE_E - 1);
      const int p = q / 3;
      d[r] = 3 * (int64_t)conn[(int64_t)(p >> 5) * ne + min(e0 + (p & 31), ne - 1)] + (q - 3 * p);
    }
    double xv[NQ];
#pragma unroll
    for (int r = 0; r < NQ; r++) xv[r] = xyz[d[r]] + (disp ? disp[d[r]] : 0.0);
#pragma unroll
    for (int r = 0; r < NQ; r++) {
      const int q = tid + r * KE_THREADS;
      if (q < 30 * KE_E) {
        const int p = q / 3;
        sG[(p & 31) * KE_ROW + 3 * (p >> 5) + (q - 3 * p)] = xv[r];
      }
    }
  }
  __syncthreads();
  const GPCoef cf = gp_coef(warp);
  double xsi[3][3], w;
  {
    double xs[3][3];
    local_gradient_tile(cf, sG + lane * KE_ROW, 1, xs);
    w = GP_W * fabs(invert_jacobian(xs, xsi));
  }
  __syncthreads();      // coordinates consumed: the gradient tiles may overwrite them
  {
    const double sw = sqrt(w);
    double T[3][3];       // T[m][j] = sqrt(w|J|) * xsi[j][m]  ->  tile row 3k+m = sqrt(w|J|) dN_k/dx_m
#pragma unroll
    for (int mm = 0; mm < 3; mm++)
#pragma unroll
      for (int j = 0; j < 3; j++) T[mm][j] = sw * xsi[j][mm];
    store_gradient_tile(cf, T, sG + (warp * 30) * KE_PAD + lane, KE_PAD);
    sW[warp][lane] = w;
    if (TANGENT) {
      // plastic Gauss points: deviator of sig_old and the factor of pmat (fcVM.py:983-997)
      double sd[6] = {0, 0, 0, 0, 0, 0}, pf = 0.0;
      if (ta.pgp[(int64_t)warp * ne + e]) {
#pragma unroll
        for (int c = 0; c < 6; c++) sd[c] = ta.sig_old[((int64_t)c * 4 + warp) * ne + e];
        const double p = (sd[0] + sd[1] + sd[2]) / 3.0;
        sd[0] -= p; sd[1] -= p; sd[2] -= p;
        double svm = sqrt(1.5 * (sd[0] * sd[0] + sd[1] * sd[1] + sd[2] * sd[2]) +
                          3.0 * (sd[3] * sd[3] + sd[4] * sd[4] + sd[5] * sd[5]));
        if (svm == 0.0) svm = 1.0;
        pf = 3.0 * ta.G / (1.0 + ta.H / 3.0 / ta.G) / (svm * svm);
      }
#pragma unroll
      for (int c = 0; c < 6; c++) sS[(warp * 7 + c) * KE_E + lane] = sd[c];
      sS[(warp * 7 + 6) * KE_E + lane] = pf;
    }
    if (GEOM) {
#pragma unroll
      for (int c = 0; c < 6; c++) sS[(warp * 7 + c) * KE_E + lane] = ta.sig_old[((int64_t)c * 4 + warp) * ne + e];
    }
  }
  __syncthreads();
  const bool live = e0 + lane < ne;
  if (elv && warp == 3 && live) {
    // gravity: gamma[3k+i] = g_i * rho * sum_gp N_k w|J|   (fcVM.py:757-759); g pre-multiplied by rho
    double *out = elv + 30 * e;
    const double w0 = sW[0][lane], w1 = sW[1][lane], w2 = sW[2][lane], w3 = sW[3][lane];
    constexpr double A = GP_A, B = GP_B;
    constexpr double c0 = 1.0 - 3.0 * A, c1 = 1.0 - 2.0 * A - B;     // 1 - xi - eta - zeta at point 0 / points 1..3
    // N_k at the four points, in Gauss-point order (fcVM.py:364-380)
    const double N[10][4] = {
        {(2 * c0 - 1) * c0, (2 * c1 - 1) * c1, (2 * c1 - 1) * c1, (2 * c1 - 1) * c1},
        {A * (2 * A - 1), B * (2 * B - 1), A * (2 * A - 1), A * (2 * A - 1)},
        {A * (2 * A - 1), A * (2 * A - 1), B * (2 * B - 1), A * (2 * A - 1)},
        {A * (2 * A - 1), A * (2 * A - 1), A * (2 * A - 1), B * (2 * B - 1)},
        {4 * A * c0, 4 * B * c1, 4 * A * c1, 4 * A * c1},
        {4 * A * A, 4 * B * A, 4 * A * B, 4 * A * A},
        {4 * A * c0, 4 * A * c1, 4 * B * c1, 4 * A * c1},
        {4 * A * c0, 4 * A * c1, 4 * A * c1, 4 * B * c1},
        {4 * A * A, 4 * B * A, 4 * A * A, 4 * A * B},
        {4 * A * A, 4 * A * A, 4 * B * A, 4 * A * B}};
#pragma unroll
    for (int k = 0; k < 10; k++) {
      const double gam = ((N[k][0] * w0 + N[k][1] * w1) + N[k][2] * w2) + N[k][3] * w3;
      out[3 * k] = gx * gam;
      out[3 * k + 1] = gy * gam;
      out[3 * k + 2] = gz * gam;
    }
  }
  const int nlive9 = (int)min((int64_t)KE_E, ne - e0) * 9;
  double *so = sO[warp];
#pragma unroll 1
  for (int ri = 0; ri < 4; ri++) {
    const int a = c_rows[warp][ri];
    if (a < 0) break;
    double ga[4][3];
#pragma unroll
    for (int gp = 0; gp < 4; gp++)
#pragma unroll
      for (int mm = 0; mm < 3; mm++) ga[gp][mm] = sG[(gp * 30 + 3 * a + mm) * KE_PAD + lane];
#pragma unroll 1
    for (int b = 0; b <= a; b++) {
      double P[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      double Q[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      double sab = 0.0;                  // buckling: sum_gp w grad N_a . sigma grad N_b
#pragma unroll
      for (int gp = 0; gp < 4; gp++) {
        double gb[3];
#pragma unroll
        for (int mm = 0; mm < 3; mm++) gb[mm] = sG[(gp * 30 + 3 * b + mm) * KE_PAD + lane];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) P[i][j] += ga[gp][i] * gb[j];
        if (GEOM) {
          double sg[6];
#pragma unroll
          for (int c = 0; c < 6; c++) sg[c] = sS[(gp * 7 + c) * KE_E + lane];
          sab += ga[gp][0] * (sg[0] * gb[0] + sg[3] * gb[1] + sg[4] * gb[2]) +
                 ga[gp][1] * (sg[3] * gb[0] + sg[1] * gb[1] + sg[5] * gb[2]) +
                 ga[gp][2] * (sg[4] * gb[0] + sg[5] * gb[1] + sg[2] * gb[2]);
        }
        if (TANGENT) {
          const double f = sS[(gp * 7 + 6) * KE_E + lane];
          if (f != 0.0) {
            // B_a^T s = S_dev g_a  (s in Voigt order xx yy zz xy zx yz)
            double sd[6];
#pragma unroll
            for (int c = 0; c < 6; c++) sd[c] = sS[(gp * 7 + c) * KE_E + lane];
            const double tax = sd[0] * ga[gp][0] + sd[3] * ga[gp][1] + sd[4] * ga[gp][2];
            const double tay = sd[3] * ga[gp][0] + sd[1] * ga[gp][1] + sd[5] * ga[gp][2];
            const double taz = sd[4] * ga[gp][0] + sd[5] * ga[gp][1] + sd[2] * ga[gp][2];
            const double tbx = sd[0] * gb[0] + sd[3] * gb[1] + sd[4] * gb[2];
            const double tby = sd[3] * gb[0] + sd[1] * gb[1] + sd[5] * gb[2];
            const double tbz = sd[4] * gb[0] + sd[5] * gb[1] + sd[2] * gb[2];
            Q[0][0] += f * tax * tbx; Q[0][1] += f * tax * tby; Q[0][2] += f * tax * tbz;
            Q[1][0] += f * tay * tbx; Q[1][1] += f * tay * tby; Q[1][2] += f * tay * tbz;
            Q[2][0] += f * taz * tbx; Q[2][1] += f * taz * tby; Q[2][2] += f * taz * tbz;
          }
        }
      }
      const double tr = mu * (P[0][0] + P[1][1] + P[2][2]);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          double v = lambda * P[i][j] + mu * P[j][i] + (i == j ? tr : 0.0);
          if (TANGENT) v -= Q[i][j];
          if (MODE == KE_BUCKLING_M) {
            if (a == b && i == j && ((ta.emask[e] >> (3 * a + i)) & 1u)) v *= 100.0;      // fcVM.py:1071-1072
            if (i == j) v += ta.sigma * sab;                                              // K - sigma G, G = -nsm
          }
          if (MODE == KE_BUCKLING_G) v = (i == j) ? -sab : 0.0;
          so[lane * 9 + 3 * i + j] = v;
        }
      __syncwarp();
      double *out = cooK + ((int64_t)(a * (a + 1) / 2 + b) * ne + e0) * 9;
#pragma unroll
      for (int m = 0; m < 9; m++) {
        const int wd = lane + 32 * m;
        if (wd < nlive9) __stcs(&out[wd], so[wd]);
      }
      __syncwarp();
    }
  }
}

// One warp per SELL slice: every stored block sums its contribution list (ascending element).
__global__ void __launch_bounds__(SELL_C)
k_coo_reduce(int64_t nslices, const int32_t *__restrict__ slice_ptr,
             const uint32_t *__restrict__ blk_first, const uint32_t *__restrict__ blk_cnt,
             const uint32_t *__restrict__ src, const double *__restrict__ cooK, double *__restrict__ vals) {
  const int64_t s = blockIdx.x;
  const int lane = threadIdx.x;
  for (int32_t k = slice_ptr[s]; k < slice_ptr[s + 1]; k++) {
    const int64_t pos = (int64_t)k * SELL_C + lane;
    const uint32_t f = blk_first[pos], n = blk_cnt[pos];
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t code = src[f + i];
      const double *p = cooK + (int64_t)(code >> 1) * 9;
      if (code & 1u) {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int q = 0; q < 3; q++) acc[3 * r + q] += p[3 * q + r];
      } else {
#pragma unroll
        for (int q = 0; q < 9; q++) acc[q] += p[q];
      }
    }
    double *o = vals + (int64_t)k * 9 * SELL_C + lane;
#pragma unroll
    for (int q = 0; q < 9; q++) o[q * SELL_C] = acc[q];
  }
}

// rows / columns of prescribed dofs: zero off-diagonal, diagonal = number of elements at the
// node -- the sum of the 1.0 entries the reference emits per element (fcVM.py:773-777)
__global__ void __launch_bounds__(SELL_C)
k_apply_constraints(int64_t nslices, const int32_t *__restrict__ slice_ptr, const int32_t *__restrict__ slot_node,
                    const int32_t *__restrict__ colidx, const int32_t *__restrict__ diag_pos,
                    const uint8_t *__restrict__ fixmask, const int32_t *__restrict__ n2e_ptr,
                    double *__restrict__ vals) {
  const int64_t s = blockIdx.x;
  const int lane = threadIdx.x;
  const int32_t row = slot_node[s * SELL_C + lane];
  if (row < 0) return;
  const bool fr[3] = {fixmask[3 * (int64_t)row] != 0, fixmask[3 * (int64_t)row + 1] != 0,
                      fixmask[3 * (int64_t)row + 2] != 0};
  const double cnt = (double)(n2e_ptr[row + 1] - n2e_ptr[row]);
  const int64_t dpos = diag_pos[row];       // padding entries also carry col == row: only this one is the diagonal
  for (int32_t k = slice_ptr[s]; k < slice_ptr[s + 1]; k++) {
    const int64_t pos = (int64_t)k * SELL_C + lane;
    const int32_t col = colidx[pos];
    const bool fc[3] = {fixmask[3 * (int64_t)col] != 0, fixmask[3 * (int64_t)col + 1] != 0,
                        fixmask[3 * (int64_t)col + 2] != 0};
    if (!(fr[0] | fr[1] | fr[2] | fc[0] | fc[1] | fc[2])) continue;
    double *o = vals + (int64_t)k * 9 * SELL_C + lane;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
        if (fr[i] || fc[j]) o[(3 * i + j) * SELL_C] = (pos == dpos && i == j) ? cnt : 0.0;
  }
}

// diagonal 3x3 blocks, stored row-wise as three nodal vectors so that the interface exchange
// can sum them over ranks like any other nodal vector
__global__ void k_extract_diag(int64_t nn, const int32_t *__restrict__ diag_pos, const double *__restrict__ vals,
                               double *__restrict__ diag9) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const int32_t pos = diag_pos[n];
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (pos >= 0) {
    const double *p = vals + (int64_t)(pos / SELL_C) * 9 * SELL_C + (pos % SELL_C);
#pragma unroll
    for (int q = 0; q < 9; q++) a[q] = p[q * SELL_C];
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) diag9[((int64_t)i * nn + n) * 3 + j] = a[3 * i + j];
}

// modf = -(K u_fix) on free dofs, K_dd * u_fix on prescribed dofs, where K_dd is the number of
// elements at the node: the sum of the 1.0 entries the reference emits (fcVM.py:773-787)
__global__ void k_modf(int64_t nn, const double *__restrict__ kufix, const uint8_t *__restrict__ fixmask,
                       const double *__restrict__ fixval, const double *__restrict__ diag9, double *modf) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= 3 * nn) return;
  const int64_t n = d / 3;
  const int cpt = (int)(d - 3 * n);
  modf[d] = fixmask[d] ? diag9[((int64_t)cpt * nn + n) * 3 + cpt] * fixval[d] : -kufix[d];
}

__global__ void k_block_inverse(int64_t nn, const double *__restrict__ diag9, double *__restrict__ minv) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  double a[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) a[3 * i + j] = diag9[((int64_t)i * nn + n) * 3 + j];
  const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  double *o = minv + 9 * n;
  if (det == 0.0) {          // node without stiffness (not referenced by any element): identity
    o[0] = o[4] = o[8] = 1.0;
    o[1] = o[2] = o[3] = o[5] = o[6] = o[7] = 0.0;
    return;
  }
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c01 * id;
  o[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c02 * id;
  o[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// full 30x30 element matrices from the stored block triangle (parity checks)
__global__ void k_expand_esm(int64_t ne, const double *__restrict__ cooK, double *__restrict__ esm) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ne * 900) return;
  const int64_t e = i / 900;
  const int rc = (int)(i - 900 * e);
  const int r = rc / 30, cidx = rc - 30 * r;
  const int a = r / 3, ii = r - 3 * a, b = cidx / 3, jj = cidx - 3 * b;
  double v;
  if (a >= b)
    v = cooK[((int64_t)(a * (a + 1) / 2 + b) * ne + e) * 9 + 3 * ii + jj];
  else
    v = cooK[((int64_t)(b * (b + 1) / 2 + a) * ne + e) * 9 + 3 * jj + ii];
  esm[i] = v;
}

// real blocks in CSR order (row node ascending, column node ascending) for the export
__global__ void k_gather_blocks(int64_t nn, const int32_t *__restrict__ row_first,
                                const int32_t *__restrict__ node_slot, const int32_t *__restrict__ slice_ptr,
                                const double *__restrict__ vals, double *__restrict__ out) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const int32_t slot = node_slot[n];
  const int32_t k0 = slice_ptr[slot / SELL_C];
  const int lane = slot % SELL_C;
  for (int32_t b = row_first[n]; b < row_first[n + 1]; b++) {
    const double *p = vals + (int64_t)(k0 + (b - row_first[n])) * 9 * SELL_C + lane;
#pragma unroll
    for (int q = 0; q < 9; q++) out[9 * (int64_t)b + q] = p[q * SELL_C];
  }
}

int run_elem_stiffness(fcvm_ctx *c, int tangent, const double *disp, double Et_E, bool gravity, double gx,
                       double gy, double gz, double sigma = 0.0) {
  const double E = c->E, nu = c->nu;
  const double dm = E * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu);
  const double lambda = dm * (nu / (1.0 - nu));
  const double mu = dm * (0.5 * (1.0 - 2.0 * nu) / (1.0 - nu));
  TangentArgs ta;
  ta.sig_old = (const double *)c->buf[FCVM_BUF_SIG_OLD];
  ta.pgp = (const uint8_t *)c->buf[FCVM_BUF_PGP];
  ta.G = E / (1.0 + nu) / 2.0;
  if (Et_E > 0.95) Et_E = 0.95;
  ta.H = (Et_E * E) / (1.0 - Et_E);
  ta.emask = c->emask;
  ta.sigma = sigma;
  if (tangent >= KE_BUCKLING_M) ta.sig_old = (const double *)c->buf[FCVM_BUF_SIG_NEW];   // the elastic stress state
  const size_t smem = 0;
  const int grid = grid_for(c->ne, KE_E);
  ProfScope ps(c, 5);
  double *elv = gravity ? c->elv : nullptr;
  const double rho = c->density;
  if (tangent == KE_TANGENT)
    k_elem_stiffness<KE_TANGENT><<<grid, KE_THREADS, smem, c->stream>>>(c->ne, c->conn, c->xyz, disp, lambda, mu, ta,
                                                                        gx * rho, gy * rho, gz * rho, c->cooK, elv);
  else if (tangent == KE_BUCKLING_M)
    k_elem_stiffness<KE_BUCKLING_M><<<grid, KE_THREADS, smem, c->stream>>>(c->ne, c->conn, c->xyz, disp, lambda, mu, ta,
                                                                           0.0, 0.0, 0.0, c->cooK, nullptr);
  else if (tangent == KE_BUCKLING_G)
    k_elem_stiffness<KE_BUCKLING_G><<<grid, KE_THREADS, smem, c->stream>>>(c->ne, c->conn, c->xyz, disp, lambda, mu, ta,
                                                                           0.0, 0.0, 0.0, c->cooK, nullptr);
  else
    k_elem_stiffness<KE_ELASTIC><<<grid, KE_THREADS, smem, c->stream>>>(c->ne, c->conn, c->xyz, disp, lambda, mu, ta,
                                                                        gx * rho, gy * rho, gz * rho, c->cooK, elv);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

}  // namespace

extern "C" int fcvm_assemble(fcvm_ctx *c, int tangent, const double *disp, double Et_E, double grav_x, double grav_y,
                             double grav_z, double *glv) {
  FCVM_CHECK(c && c->ne > 0, FCVM_E_ARG, "fcvm_assemble: no mesh");
  FCVM_CHECK(c->have_bcs, FCVM_E_ARG, "fcvm_assemble: call fcvm_set_constraints first");
  ProfScope ps(c, 4);
  const bool gravity = glv != nullptr;
  FCVM_TRY(run_elem_stiffness(c, tangent ? KE_TANGENT : KE_ELASTIC, disp, Et_E, gravity, grav_x, grav_y, grav_z));
  if (gravity) {
    FCVM_TRY(launch_node_gather(c, glv, 1));       // glv += gravity  (fcVM.py:763-767)
    // shared nodes: the caller passes surface loads already summed once; gravity parts are per rank
  }
  {
    ProfScope ps6(c, 6);
    k_coo_reduce<<<(unsigned)c->nslices, SELL_C, 0, c->stream>>>(c->nslices, c->slice_ptr, c->blk_first, c->blk_cnt,
                                                                c->src, c->cooK, c->vals);
  }
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  // modf needs the unconstrained operator: K * u_fix before rows/columns are eliminated
  c->assembled = true;
  c->matrix_elastic = !tangent && disp == nullptr;
  c->hist_n = 0;                                   // recycled solutions belong to the previous matrix
  FCVM_TRY(launch_spmv(c, c->fixval, c->pcg_q));
  FCVM_TRY(fcvm_interface_sum(c, c->pcg_q));
  k_apply_constraints<<<(unsigned)c->nslices, SELL_C, 0, c->stream>>>(c->nslices, c->slice_ptr, c->slot_node,
                                                                     c->colidx, c->diag_pos, c->fixmask, c->n2e_ptr,
                                                                     c->vals);
  if (!c->diag9) FCVM_CUDA(cudaMalloc((void **)&c->diag9, sizeof(double) * 9 * c->nn));
  k_extract_diag<<<grid_for(c->nn, 128), 128, 0, c->stream>>>(c->nn, c->diag_pos, c->vals, c->diag9);
  for (int i = 0; i < 3; i++) FCVM_TRY(fcvm_interface_sum(c, c->diag9 + (int64_t)i * 3 * c->nn));
  k_modf<<<grid_for(3 * c->nn, 256), 256, 0, c->stream>>>(c->nn, c->pcg_q, c->fixmask, c->fixval, c->diag9,
                                                          (double *)c->buf[FCVM_BUF_MODF]);
  k_block_inverse<<<grid_for(c->nn, 128), 128, 0, c->stream>>>(c->nn, c->diag9, c->minv);
  c->launches += 4;
  FCVM_CUDA(cudaGetLastError());
  return deflation_build(c);       // K Z and (Z^T K Z)^-1 of the second preconditioner level, when switched on
}

// Matrices of the linear buckling analysis (the nstep == 1 branch of calcTSM, fcVM.py:1002-1006 and 1063-1073, used
// at fcVM.py:1199-1212): K = elastic stiffness of the undeformed mesh, NOT eliminated -- the diagonal entries of
// prescribed dofs are multiplied by 100 instead -- and G = -nsm, the geometric stiffness of the stress state in
// SIG_NEW.  The context's matrix becomes M = K - sigma G (what the shift-invert iteration solves with, through
// fcvm_pcg_solve: block-Jacobi PCG, no deflation), G goes to a second value array on the same pattern
// (fcvm_spmv_geometric).
extern "C" int fcvm_assemble_buckling(fcvm_ctx *c, double sigma) {
  FCVM_CHECK(c && c->ne > 0 && c->have_bcs, FCVM_E_ARG, "fcvm_assemble_buckling: set the mesh and the constraints first");
  FCVM_CHECK(c->world == 1, FCVM_E_ARG, "fcvm_assemble_buckling: single GPU only");
  ProfScope ps(c, 4);
  if (!c->vals2) FCVM_CUDA(cudaMalloc((void **)&c->vals2, sizeof(double) * 9 * (size_t)c->nblk_stored));
  FCVM_TRY(run_elem_stiffness(c, KE_BUCKLING_G, nullptr, 0.0, false, 0, 0, 0, sigma));
  k_coo_reduce<<<(unsigned)c->nslices, SELL_C, 0, c->stream>>>(c->nslices, c->slice_ptr, c->blk_first, c->blk_cnt, c->src,
                                                              c->cooK, c->vals2);
  FCVM_TRY(run_elem_stiffness(c, KE_BUCKLING_M, nullptr, 0.0, false, 0, 0, 0, sigma));
  k_coo_reduce<<<(unsigned)c->nslices, SELL_C, 0, c->stream>>>(c->nslices, c->slice_ptr, c->blk_first, c->blk_cnt, c->src,
                                                              c->cooK, c->vals);
  if (!c->diag9) FCVM_CUDA(cudaMalloc((void **)&c->diag9, sizeof(double) * 9 * c->nn));
  k_extract_diag<<<grid_for(c->nn, 128), 128, 0, c->stream>>>(c->nn, c->diag_pos, c->vals, c->diag9);
  k_block_inverse<<<grid_for(c->nn, 128), 128, 0, c->stream>>>(c->nn, c->diag9, c->minv);
  c->launches += 4;
  FCVM_CUDA(cudaGetLastError());
  c->assembled = true;
  c->matrix_elastic = false;
  c->hist_n = 0;
  c->defl_ready = false;            // the coarse level belongs to the eliminated operator
  return FCVM_OK;
}

namespace fcvm {
int launch_spmv_values(fcvm_ctx *c, const double *vals, const double *x, double *y);
}

// y = G x with the geometric stiffness of fcvm_assemble_buckling
extern "C" int fcvm_spmv_geometric(fcvm_ctx *c, const double *x, double *y) {
  FCVM_CHECK(c && c->vals2 && x && y, FCVM_E_ARG, "fcvm_spmv_geometric: call fcvm_assemble_buckling first");
  return launch_spmv_values(c, c->vals2, x, y);
}

extern "C" int fcvm_element_matrices(fcvm_ctx *c, int tangent, const double *disp, double Et_E, double *esm_dev) {
  FCVM_CHECK(c && c->ne > 0 && esm_dev, FCVM_E_ARG, "fcvm_element_matrices: null argument / no mesh");
  FCVM_TRY(run_elem_stiffness(c, tangent ? KE_TANGENT : KE_ELASTIC, disp, Et_E, false, 0, 0, 0));
  k_expand_esm<<<grid_for(c->ne * 900, 256), 256, 0, c->stream>>>(c->ne, c->cooK, esm_dev);
  c->launches++;
  FCVM_CUDA(cudaGetLastError());
  return FCVM_OK;
}

extern "C" int fcvm_export_csc_lower(fcvm_ctx *c, int64_t *nnz_out, int64_t *indptr, int64_t *indices, double *data) {
  FCVM_CHECK(c && c->assembled && nnz_out, FCVM_E_ARG, "fcvm_export_csc_lower: assemble first");
  const int64_t nn = c->nn, nb = c->nblk_real;
  std::vector<int32_t> row_first((size_t)nn + 1), cols((size_t)nb);
  std::vector<uint8_t> fm((size_t)3 * nn);
  FCVM_CUDA(cudaStreamSynchronize(c->stream));
  FCVM_CUDA(cudaMemcpy(row_first.data(), c->row_first, sizeof(int32_t) * (nn + 1), cudaMemcpyDeviceToHost));
  FCVM_CUDA(cudaMemcpy(cols.data(), c->row_cols, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost));
  FCVM_CUDA(cudaMemcpy(fm.data(), c->fixmask, 3 * nn, cudaMemcpyDeviceToHost));
  std::vector<double> bv;
  if (indices) {
    double *d;
    FCVM_CUDA(cudaMalloc((void **)&d, sizeof(double) * 9 * nb));
    k_gather_blocks<<<grid_for(nn, 128), 128, 0, c->stream>>>(nn, c->row_first, c->node_slot, c->slice_ptr, c->vals,
                                                             d);
    bv.resize((size_t)9 * nb);
    FCVM_CUDA(cudaMemcpyAsync(bv.data(), d, sizeof(double) * 9 * nb, cudaMemcpyDeviceToHost, c->stream));
    FCVM_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d);
  }
  // lower-triangular CSC column j == upper-triangular CSR row j of the symmetric matrix
  int64_t nnz = 0;
  for (int64_t r = 0; r < nn; r++)
    for (int a = 0; a < 3; a++) {
      const int64_t j = 3 * r + a;
      if (indptr) indptr[j] = nnz;
      for (int32_t b = row_first[r]; b < row_first[r + 1]; b++) {
        const int64_t cn = cols[b];
        if (cn < r) continue;
        for (int q = 0; q < 3; q++) {
          const int64_t i = 3 * cn + q;
          if (i < j) continue;
          const bool keep = (i == j) ? true : (!fm[i] && !fm[j]);
          if (!keep) continue;
          if (indices) {
            indices[nnz] = i;
            data[nnz] = bv[9 * (size_t)b + 3 * a + q];
          }
          nnz++;
        }
      }
    }
  if (indptr) indptr[3 * nn] = nnz;
  *nnz_out = nnz;
  return FCVM_OK;
}
