/*
 * TEST INFRASTRUCTURE -- CPU oracle for the fcVM Newton/load-stepping hot path.
 *
 * A plain-C restatement of the numba-jitted element routines of the reference
 * (HarryvL/fcVM-workbench, "source code/fcVM.py").  Every function cites the
 * reference lines it follows.  Loop order and floating-point association follow
 * the reference statement by statement, so that this file can stand in for it on
 * machines where the reference (FreeCAD + numba + CHOLMOD) is not installed.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path (fcvm_workbench_b200)
 * never does: it fails loudly when its CUDA library is missing.
 *
 * Parity pin: oracle/gen_golden.py runs the UNMODIFIED reference (through
 * oracle/ref_harness.py) and this oracle on the same inputs; the outputs of the
 * reference are committed as tests/golden/ fixtures and
 * tests/test_oracle_vs_golden.py checks this file against them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Gauss points: fcVM.py:586-613 (gaussPoints) ------------------------- */
static const double GP10[4][4] = {
    {0.138196601125011, 0.138196601125011, 0.138196601125011, 0.041666666666667},
    {0.585410196624968, 0.138196601125011, 0.138196601125011, 0.041666666666667},
    {0.138196601125011, 0.585410196624968, 0.138196601125011, 0.041666666666667},
    {0.138196601125011, 0.138196601125011, 0.585410196624968, 0.041666666666667}};
static const double GP6[6][3] = {
    {0.445948490915965, 0.445948490915965, 0.111690794839005},
    {0.10810301816807, 0.445948490915965, 0.111690794839005},
    {0.445948490915965, 0.10810301816807, 0.111690794839005},
    {0.091576213509771, 0.091576213509771, 0.054975871827661},
    {0.816847572980458, 0.091576213509771, 0.054975871827661},
    {0.091576213509771, 0.816847572980458, 0.054975871827661}};
static const double GP2[2][2] = {{-0.5773502691896257, 1.0}, {0.5773502691896257, 1.0}};

/* ---- fcVM.py:364-380 (shp10tet) ------------------------------------------ */
static void shp10tet(double xi, double et, double ze, double shp[10]) {
  double a = 1.0 - xi - et - ze;
  shp[0] = (2.0 * a - 1.0) * a;
  shp[1] = xi * (2.0 * xi - 1.0);
  shp[2] = et * (2.0 * et - 1.0);
  shp[3] = ze * (2.0 * ze - 1.0);
  shp[4] = 4.0 * xi * a;
  shp[5] = 4.0 * xi * et;
  shp[6] = 4.0 * et * a;
  shp[7] = 4.0 * ze * a;
  shp[8] = 4.0 * xi * ze;
  shp[9] = 4.0 * et * ze;
}

/* local derivatives: fcVM.py:390-424 */
static void dshp_local(double xi, double et, double ze, double d[3][10]) {
  memset(d, 0, sizeof(double) * 30);
  double a4 = 1.0 - 4.0 * (1.0 - xi - et - ze);
  d[0][0] = a4;
  d[0][1] = 4.0 * xi - 1.0;
  d[0][4] = 4.0 * (1.0 - 2.0 * xi - et - ze);
  d[0][5] = 4.0 * et;
  d[0][6] = -4.0 * et;
  d[0][7] = -4.0 * ze;
  d[0][8] = 4.0 * ze;
  d[1][0] = a4;
  d[1][2] = 4.0 * et - 1.0;
  d[1][4] = -4.0 * xi;
  d[1][5] = 4.0 * xi;
  d[1][6] = 4.0 * (1.0 - xi - 2.0 * et - ze);
  d[1][7] = -4.0 * ze;
  d[1][9] = 4.0 * ze;
  d[2][0] = a4;
  d[2][3] = 4.0 * ze - 1.0;
  d[2][4] = -4.0 * xi;
  d[2][6] = -4.0 * et;
  d[2][7] = 4.0 * (1.0 - xi - et - 2.0 * ze);
  d[2][8] = 4.0 * xi;
  d[2][9] = 4.0 * et;
}

/* ---- fcVM.py:383-480 (dshp10tet): xsj, global derivatives, B matrix ------- */
static double dshp10tet(double xi, double et, double ze, const double xl[10][3], double bmat[6][30],
                        double dshpg[3][10]) {
  double dshp[3][10], xs[3][3], xsi[3][3];
  dshp_local(xi, et, ze, dshp);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      xs[i][j] = 0.0;
      for (int k = 0; k < 10; k++) xs[i][j] += xl[k][i] * dshp[j][k];
    }
  double xsj = (xs[0][0] * xs[1][1] * xs[2][2] - xs[0][0] * xs[1][2] * xs[2][1] + xs[0][2] * xs[1][0] * xs[2][1] -
                xs[0][2] * xs[1][1] * xs[2][0] + xs[0][1] * xs[1][2] * xs[2][0] - xs[0][1] * xs[1][0] * xs[2][2]);
  xsi[0][0] = (xs[1][1] * xs[2][2] - xs[2][1] * xs[1][2]) / xsj;
  xsi[0][1] = (xs[0][2] * xs[2][1] - xs[0][1] * xs[2][2]) / xsj;
  xsi[0][2] = (xs[0][1] * xs[1][2] - xs[0][2] * xs[1][1]) / xsj;
  xsi[1][0] = (xs[1][2] * xs[2][0] - xs[1][0] * xs[2][2]) / xsj;
  xsi[1][1] = (xs[0][0] * xs[2][2] - xs[0][2] * xs[2][0]) / xsj;
  xsi[1][2] = (xs[1][0] * xs[0][2] - xs[0][0] * xs[1][2]) / xsj;
  xsi[2][0] = (xs[1][0] * xs[2][1] - xs[2][0] * xs[1][1]) / xsj;
  xsi[2][1] = (xs[2][0] * xs[0][1] - xs[0][0] * xs[2][1]) / xsj;
  xsi[2][2] = (xs[0][0] * xs[1][1] - xs[1][0] * xs[0][1]) / xsj;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 10; j++) {
      dshpg[i][j] = 0.0;
      for (int k = 0; k < 3; k++) dshpg[i][j] += xsi[k][i] * dshp[k][j];
    }
  if (bmat) {
    /* the reference never clears bmat; its zero pattern is fixed (fcVM.py:465-478) */
    for (int i = 0; i < 10; i++) {
      int i3 = 3 * i;
      double d00 = dshpg[0][i], d10 = dshpg[1][i], d20 = dshpg[2][i];
      bmat[0][i3] = d00;
      bmat[1][i3 + 1] = d10;
      bmat[2][i3 + 2] = d20;
      bmat[3][i3] = d10;
      bmat[3][i3 + 1] = d00;
      bmat[4][i3] = d20;
      bmat[4][i3 + 2] = d00;
      bmat[5][i3 + 1] = d20;
      bmat[5][i3 + 2] = d10;
    }
  }
  return xsj;
}

/* ---- fcVM.py:484-541 (shape6tri): shape functions, |J| and unit normal ---- */
static double shape6tri(double xi, double et, const double xl[3][6], double shp[6], double xp[3]) {
  double dshp[2][6];
  shp[0] = (1.0 - xi - et) * (1.0 - 2.0 * xi - 2.0 * et);
  shp[1] = xi * (2.0 * xi - 1.0);
  shp[2] = et * (2.0 * et - 1.0);
  shp[3] = 4.0 * xi * (1.0 - xi - et);
  shp[4] = 4.0 * xi * et;
  shp[5] = 4.0 * et * (1 - xi - et);
  dshp[0][0] = -3.0 + 4.0 * et + 4.0 * xi;
  dshp[0][1] = -1.0 + 4.0 * xi;
  dshp[0][2] = 0.0;
  dshp[0][3] = -4.0 * (-1.0 + et + 2.0 * xi);
  dshp[0][4] = 4.0 * et;
  dshp[0][5] = -4.0 * et;
  dshp[1][0] = -3.0 + 4.0 * et + 4.0 * xi;
  dshp[1][1] = 0.0;
  dshp[1][2] = -1.0 + 4.0 * et;
  dshp[1][3] = -4.0 * xi;
  dshp[1][4] = 4.0 * xi;
  dshp[1][5] = -4.0 * (-1.0 + 2.0 * et + xi);
  double xs[2][3];
  for (int a = 0; a < 2; a++)
    for (int c = 0; c < 3; c++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += dshp[a][k] * xl[c][k];
      xs[a][c] = s;
    }
  xp[0] = xs[0][1] * xs[1][2] - xs[0][2] * xs[1][1];
  xp[1] = xs[0][2] * xs[1][0] - xs[0][0] * xs[1][2];
  xp[2] = xs[0][0] * xs[1][1] - xs[0][1] * xs[1][0];
  double xsj = sqrt(xp[0] * xp[0] + xp[1] * xp[1] + xp[2] * xp[2]);
  xp[0] /= xsj;
  xp[1] /= xsj;
  xp[2] /= xsj;
  return xsj;
}

/* ---- fcVM.py:544-565 (shape2lin) ------------------------------------------ */
static double shape2lin(double xi, const double xle[3][3], double shp[3]) {
  double dshp[3];
  shp[0] = -0.5 * (1.0 - xi) * xi;
  shp[1] = 0.5 * (1.0 + xi) * xi;
  shp[2] = (1.0 + xi) * (1.0 - xi);
  dshp[0] = xi - 0.5;
  dshp[1] = xi + 0.5;
  dshp[2] = -2.0 * xi;
  double dx = xle[0][0] * dshp[0] + xle[0][1] * dshp[1] + xle[0][2] * dshp[2];
  double dy = xle[1][0] * dshp[0] + xle[1][1] * dshp[1] + xle[1][2] * dshp[2];
  double dz = xle[2][0] * dshp[0] + xle[2][1] * dshp[1] + xle[2][2] * dshp[2];
  return sqrt(dx * dx + dy * dy + dz * dz);
}

/* ---- fcVM.py:570-582 (hooke) ---------------------------------------------- */
static void hooke(double e, double nu, double dmat[6][6]) {
  double dm = e * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu);
  double od = nu / (1.0 - nu);
  double sd = 0.5 * (1.0 - 2.0 * nu) / (1.0 - nu);
  memset(dmat, 0, sizeof(double) * 36);
  dmat[0][0] = dmat[1][1] = dmat[2][2] = 1.0;
  dmat[3][3] = dmat[4][4] = dmat[5][5] = sd;
  dmat[0][1] = dmat[0][2] = dmat[1][2] = od;
  dmat[1][0] = dmat[2][0] = dmat[2][1] = od;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) dmat[i][j] *= dm;
}

/* ---- external load vector: fcVM.py:647-727 (calcGSM) / 856-938 (calcTSM) ---
 * disp may be NULL (calcGSM) or the converged displacement (calcTSM: pressure
 * follows the stretched surface, fcVM.py:866-868).  Tables carry the reference's
 * dummy first row. */
void fcvm_oracle_load_vector(int64_t nn, const double *nocoord, const double *disp, int64_t n_press,
                             const int64_t *loadfaces, const double *pressure, int64_t n_vert,
                             const int64_t *loadvertices, const double *vertexloads, int64_t n_edge,
                             const int64_t *loadedges, const double *edgeloads, int64_t n_funi,
                             const int64_t *loadfaces_uni, const double *faceloads, double *glv) {
  double xlf[3][6], xle[3][3], shp[6], xp[3];
  (void)nn;
  for (int64_t face = 0; face < n_press - 1; face++) {
    const int64_t *nda = loadfaces + 6 * (face + 1);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 6; j++) {
        int64_t nd = nda[j];
        xlf[i][j] = nocoord[3 * (nd - 1) + i] + (disp ? disp[3 * (nd - 1) + i] : 0.0);
      }
    for (int index = 0; index < 6; index++) {
      double xsj = shape6tri(GP6[index][0], GP6[index][1], xlf, shp, xp);
      for (int i = 0; i < 6; i++) {
        int64_t iglob3 = 3 * (nda[i] - 1);
        for (int k = 0; k < 3; k++) {
          double load = shp[i] * pressure[face + 1] * xp[k] * fabs(xsj) * GP6[index][2];
          glv[iglob3 + k] += load;
        }
      }
    }
  }
  for (int64_t v = 0; v < n_vert - 1; v++) {
    int64_t iglob3 = 3 * (loadvertices[v + 1] - 1);
    for (int k = 0; k < 3; k++) glv[iglob3 + k] += vertexloads[3 * (v + 1) + k];
  }
  for (int64_t face = 0; face < n_funi - 1; face++) {
    const int64_t *nda = loadfaces_uni + 6 * (face + 1);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 6; j++) xlf[i][j] = nocoord[3 * (nda[j] - 1) + i];
    for (int index = 0; index < 6; index++) {
      double xsj = shape6tri(GP6[index][0], GP6[index][1], xlf, shp, xp);
      for (int i = 0; i < 6; i++) {
        int64_t iglob3 = 3 * (nda[i] - 1);
        for (int k = 0; k < 3; k++) {
          double load = shp[i] * faceloads[3 * (face + 1) + k] * fabs(xsj) * GP6[index][2];
          glv[iglob3 + k] += load;
        }
      }
    }
  }
  for (int64_t edge = 0; edge < n_edge - 1; edge++) {
    const int64_t *nda = loadedges + 3 * (edge + 1);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) xle[i][j] = nocoord[3 * (nda[j] - 1) + i];
    for (int index = 0; index < 2; index++) {
      double xsj = shape2lin(GP2[index][0], xle, shp);
      for (int i = 0; i < 3; i++) {
        int64_t iglob3 = 3 * (nda[i] - 1);
        for (int k = 0; k < 3; k++) {
          double load = shp[i] * edgeloads[3 * (edge + 1) + k] * fabs(xsj) * GP2[index][1];
          glv[iglob3 + k] += load;
        }
      }
    }
  }
}

/* esm += B^T (D) B * w : fcVM.py:756 / 998-1000 */
static void add_btdb(double esm[30][30], const double bmat[6][30], const double d[6][6], double w) {
  double db[6][30];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 30; j++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += d[i][k] * bmat[k][j];
      db[i][j] = s;
    }
  for (int i = 0; i < 30; i++)
    for (int j = 0; j < 30; j++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += bmat[k][i] * db[k][j];
      esm[i][j] += s * w;
    }
}

/* COO emission with displacement boundary conditions: fcVM.py:771-796 / 1022-1050 */
static void emit_coo(const double esm[30][30], const int64_t dof[30], const uint8_t *fixmask, const double *fixval,
                     int64_t *row, int64_t *col, double *stm, double *modf, int64_t *ppos) {
  int64_t pos = *ppos;
  for (int i = 0; i < 30; i++) {
    int64_t dofi = dof[i];
    if (fixmask[dofi]) {
      row[pos] = dofi;
      col[pos] = dofi;
      stm[pos] = 1.0;
      modf[dofi] += fixval[dofi];
      pos++;
      for (int j = 0; j < i; j++) {
        int64_t dofj = dof[j];
        if (!fixmask[dofj]) modf[dofj] -= esm[i][j] * fixval[dofi];
      }
    } else {
      for (int j = 0; j <= i; j++) {
        int64_t dofj = dof[j];
        if (fixmask[dofj]) {
          modf[dofi] -= esm[i][j] * fixval[dofj];
        } else {
          if (dofi > dofj) {
            row[pos] = dofi;
            col[pos] = dofj;
          } else {
            row[pos] = dofj;
            col[pos] = dofi;
          }
          stm[pos] = esm[i][j];
          pos++;
        }
      }
    }
  }
  *ppos = pos;
}

/* ---- fcVM.py:620-816 (calcGSM), volume-element part ------------------------
 * glv must come in holding the surface/edge/vertex loads (fcvm_oracle_load_vector)
 * or zeros; gravity is added here.  row/col/stm need 465*ne entries.  esm_out
 * (optional, ne*900) receives the raw element matrices for element-level parity. */
int64_t fcvm_oracle_calc_gsm(int64_t ne, int64_t nn, const int64_t *elNodes, const double *nocoord, double E,
                             double nu, double density, const uint8_t *fixmask, const double *fixval, double grav_x,
                             double grav_y, double grav_z, double *glv, int64_t *row, int64_t *col, double *stm,
                             double *modf, double *x, double *V_out, double *esm_out) {
  double dmat[6][6], bmatV[6][30], dshpg[3][10], xlv[10][3], shp[10];
  double(*esm)[30] = malloc(sizeof(double) * 900);
  double gamma[30];
  int64_t dof[30];
  int64_t pos = 0;
  double V = 0.0;
  (void)nn;
  memset(bmatV, 0, sizeof(bmatV));
  hooke(E, nu, dmat);
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    memset(esm, 0, sizeof(double) * 900);
    memset(gamma, 0, sizeof(gamma));
    for (int j = 0; j < 10; j++)
      for (int i = 0; i < 3; i++) xlv[j][i] = nocoord[3 * (nodes[j] - 1) + i];
    for (int ip = 0; ip < 4; ip++) {
      double xi = GP10[ip][0], et = GP10[ip][1], ze = GP10[ip][2], w = GP10[ip][3];
      shp10tet(xi, et, ze, shp);
      double xsj = dshp10tet(xi, et, ze, xlv, bmatV, dshpg);
      add_btdb(esm, bmatV, dmat, w * fabs(xsj));
      for (int k = 0; k < 10; k++) {
        gamma[3 * k] += grav_x * density * shp[k] * w * fabs(xsj);
        gamma[3 * k + 1] += grav_y * density * shp[k] * w * fabs(xsj);
        gamma[3 * k + 2] += grav_z * density * shp[k] * w * fabs(xsj);
      }
      V += xsj * w;
      for (int c = 0; c < 3; c++) {
        double s = 0.0;
        for (int k = 0; k < 10; k++) s += xlv[k][c] * shp[k];
        x[3 * (4 * el + ip) + c] = s;
      }
    }
    for (int i = 0; i < 10; i++) {
      int64_t nd = nodes[i] - 1;
      glv[3 * nd] += gamma[3 * i];
      glv[3 * nd + 1] += gamma[3 * i + 1];
      glv[3 * nd + 2] += gamma[3 * i + 2];
      for (int j = 0; j < 3; j++) dof[3 * i + j] = 3 * nd + j;
    }
    if (esm_out) memcpy(esm_out + 900 * el, esm, sizeof(double) * 900);
    emit_coo((const double(*)[30])esm, dof, fixmask, fixval, row, col, stm, modf, &pos);
  }
  free(esm);
  if (V_out) *V_out = V;
  return pos;
}

/* ---- fcVM.py:819-1079 (calcTSM), nstep > 1 branch: consistent tangent -------
 * D - pmat at plastic Gauss points (fcVM.py:983-1000), geometry updated with
 * disp_new (fcVM.py:962-967).  The linear-buckling branch (nstep == 1) follows below. */
int64_t fcvm_oracle_calc_tsm(int64_t ne, int64_t nn, const int64_t *elNodes, const double *nocoord, double E,
                             double nu, double density, const uint8_t *fixmask, const double *fixval, double grav_x,
                             double grav_y, double grav_z, const double *disp_new, const double *sig_old,
                             const uint8_t *pgp, double Et_E, double *glv, int64_t *row, int64_t *col, double *stm,
                             double *modf, double *esm_out) {
  double dmat[6][6], dp[6][6], bmatV[6][30], dshpg[3][10], xlv[10][3], shp[10];
  double(*esm)[30] = malloc(sizeof(double) * 900);
  double gamma[30];
  int64_t dof[30];
  int64_t pos = 0;
  (void)nn;
  memset(bmatV, 0, sizeof(bmatV));
  hooke(E, nu, dmat);
  double G = E / (1.0 + nu) / 2.0;
  if (Et_E > 0.95) Et_E = 0.95;
  double Et = Et_E * E;
  double H = Et / (1.0 - Et_E);
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    memset(esm, 0, sizeof(double) * 900);
    memset(gamma, 0, sizeof(gamma));
    for (int j = 0; j < 10; j++)
      for (int i = 0; i < 3; i++) {
        int64_t idof = 3 * (nodes[j] - 1) + i;
        xlv[j][i] = nocoord[idof] + disp_new[idof];
      }
    for (int ip = 0; ip < 4; ip++) {
      int64_t ip4 = 4 * el + ip, ip24 = 24 * el + 6 * ip;
      double xi = GP10[ip][0], et = GP10[ip][1], ze = GP10[ip][2], w = GP10[ip][3];
      shp10tet(xi, et, ze, shp);
      double xsj = dshp10tet(xi, et, ze, xlv, bmatV, dshpg);
      if (pgp[ip4]) {
        double s[6];
        for (int c = 0; c < 6; c++) s[c] = sig_old[ip24 + c];
        double p = (s[0] + s[1] + s[2]) / 3.0;
        s[0] -= p;
        s[1] -= p;
        s[2] -= p;
        double svm = sqrt(1.5 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) +
                          3.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]));
        if (svm == 0.0) svm = 1.0;
        double fac = 3.0 * G / (1.0 + H / 3.0 / G) / (svm * svm);
        for (int i1 = 0; i1 < 6; i1++)
          for (int i2 = 0; i2 < 6; i2++) dp[i1][i2] = dmat[i1][i2] - fac * s[i1] * s[i2];
        add_btdb(esm, bmatV, dp, w * fabs(xsj));
      } else {
        add_btdb(esm, bmatV, dmat, w * fabs(xsj));
      }
      for (int k = 0; k < 10; k++) {
        gamma[3 * k] += grav_x * density * shp[k] * w * fabs(xsj);
        gamma[3 * k + 1] += grav_y * density * shp[k] * w * fabs(xsj);
        gamma[3 * k + 2] += grav_z * density * shp[k] * w * fabs(xsj);
      }
    }
    for (int i = 0; i < 10; i++) {
      int64_t nd = nodes[i] - 1;
      glv[3 * nd] += gamma[3 * i];
      glv[3 * nd + 1] += gamma[3 * i + 1];
      glv[3 * nd + 2] += gamma[3 * i + 2];
      for (int j = 0; j < 3; j++) dof[3 * i + j] = 3 * nd + j;
    }
    if (esm_out) memcpy(esm_out + 900 * el, esm, sizeof(double) * 900);
    emit_coo((const double(*)[30])esm, dof, fixmask, fixval, row, col, stm, modf, &pos);
  }
  free(esm);
  return pos;
}

/* ---- fcVM.py:819-1079 (calcTSM), nstep == 1 branch: linear buckling ----------
 * Material stiffness esm (D, or D - pmat at plastic points) and geometric stiffness
 * nsm = sum_gp w|J| GM^T SM GM with GM = kron(dshpg, I3), SM = kron(sig, I3)
 * (fcVM.py:1002-1006) on the undeformed geometry; the full 30 x 30 element matrices go
 * out as COO triplets, rows and columns of prescribed dofs are kept and their diagonal
 * is stiffened by 100 (fcVM.py:1052-1062).  row/col/stms/stmg need 900*ne entries. */
int64_t fcvm_oracle_calc_tsm_buckling(int64_t ne, const int64_t *elNodes, const double *nocoord, double E, double nu,
                                      const uint8_t *fixmask, const double *sig_old, const uint8_t *pgp, double Et_E,
                                      int64_t *row, int64_t *col, double *stms, double *stmg) {
  double dmat[6][6], dp[6][6], bmatV[6][30], dshpg[3][10], xlv[10][3];
  double(*esm)[30] = malloc(sizeof(double) * 900);
  double(*nsm)[30] = malloc(sizeof(double) * 900);
  int64_t dof[30];
  int64_t pos = 0;
  memset(bmatV, 0, sizeof(bmatV));
  hooke(E, nu, dmat);
  double G = E / (1.0 + nu) / 2.0;
  if (Et_E > 0.95) Et_E = 0.95;
  double Et = Et_E * E;
  double H = Et / (1.0 - Et_E);
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    memset(esm, 0, sizeof(double) * 900);
    memset(nsm, 0, sizeof(double) * 900);
    for (int j = 0; j < 10; j++)
      for (int i = 0; i < 3; i++) xlv[j][i] = nocoord[3 * (nodes[j] - 1) + i];
    for (int ip = 0; ip < 4; ip++) {
      int64_t ip4 = 4 * el + ip, ip24 = 24 * el + 6 * ip;
      double xi = GP10[ip][0], et = GP10[ip][1], ze = GP10[ip][2], w = GP10[ip][3];
      double xsj = dshp10tet(xi, et, ze, xlv, bmatV, dshpg);
      const double *s0 = sig_old + ip24;
      if (pgp[ip4]) {
        double s[6];
        for (int c = 0; c < 6; c++) s[c] = s0[c];
        double p = (s[0] + s[1] + s[2]) / 3.0;
        s[0] -= p;
        s[1] -= p;
        s[2] -= p;
        double svm = sqrt(1.5 * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) +
                          3.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]));
        if (svm == 0.0) svm = 1.0;
        double fac = 3.0 * G / (1.0 + H / 3.0 / G) / (svm * svm);
        for (int i1 = 0; i1 < 6; i1++)
          for (int i2 = 0; i2 < 6; i2++) dp[i1][i2] = dmat[i1][i2] - fac * s[i1] * s[i2];
        add_btdb(esm, bmatV, dp, w * fabs(xsj));
      } else {
        add_btdb(esm, bmatV, dmat, w * fabs(xsj));
      }
      /* sig as a 3 x 3 tensor (Voigt order xx yy zz xy zx yz, fcVM.py:970-972) */
      const double sg[3][3] = {{s0[0], s0[3], s0[4]}, {s0[3], s0[1], s0[5]}, {s0[4], s0[5], s0[2]}};
      for (int a = 0; a < 10; a++)
        for (int b = 0; b < 10; b++) {
          double v = 0.0;
          for (int m = 0; m < 3; m++)
            for (int n = 0; n < 3; n++) v += dshpg[m][a] * sg[m][n] * dshpg[n][b];
          v *= w * fabs(xsj);
          for (int i = 0; i < 3; i++) nsm[3 * a + i][3 * b + i] += v;
        }
    }
    for (int i = 0; i < 10; i++)
      for (int j = 0; j < 3; j++) dof[3 * i + j] = 3 * (nodes[i] - 1) + j;
    for (int i = 0; i < 30; i++)
      for (int j = 0; j < 30; j++) {
        row[pos] = dof[i];
        col[pos] = dof[j];
        stms[pos] = esm[i][j];
        stmg[pos] = nsm[i][j];
        if (i == j && fixmask[dof[i]]) stms[pos] *= 100.0;
        pos++;
      }
  }
  free(esm);
  free(nsm);
  return pos;
}

/* ---- fcVM.py:2468-2492 (vmises_original_optimised) ------------------------- */
static int vmises(const double st_in[6], double sy, double H, double G, double out[6]) {
  double st0 = st_in[0], st1 = st_in[1], st2 = st_in[2], st3 = st_in[3], st4 = st_in[4], st5 = st_in[5];
  double p = (st0 + st1 + st2) / 3.0;
  st0 -= p;
  st1 -= p;
  st2 -= p;
  double sig_mises = sqrt(1.5 * (st0 * st0 + st1 * st1 + st2 * st2) + 3.0 * (st3 * st3 + st4 * st4 + st5 * st5));
  double fac;
  int pp;
  if (sy > sig_mises) {
    fac = 1.0;
    pp = 0;
  } else {
    fac = (1.0 - (1.0 - sy / sig_mises) * 3.0 * G / (H + 3 * G));
    pp = 1;
  }
  out[0] = fac * st0 + p;
  out[1] = fac * st1 + p;
  out[2] = fac * st2 + p;
  out[3] = fac * st3;
  out[4] = fac * st4;
  out[5] = fac * st5;
  return pp;
}

/* ---- fcVM.py:2196-2464 (update_stress_load) --------------------------------
 * svm_test (optional, 4*ne) receives the von Mises value of the elastic test
 * stress so that tests can exclude Gauss points inside the yield-surface band. */
void fcvm_oracle_update_stress_load(int64_t ne, int64_t nn, const int64_t *elNodes, const double *nocoord, double E,
                                    double nu, const double *sig_yield, const double *disp_new, const double *du,
                                    const double *sig, double *sig_update, double *sig_test_global, double *qin,
                                    double Et_E, int LD, uint8_t *pgp) {
  double dmat[6][6], xlv[10][3], du10[10][3], elv[30], dshpg[3][10];
  (void)nn;
  hooke(E, nu, dmat);
  double G = E / 2.0 / (1 + nu);
  if (Et_E > 0.95) Et_E = 0.95;
  double Et = Et_E * E;
  double H = Et / (1.0 - Et_E);
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    int64_t elpos = 24 * el;
    memset(elv, 0, sizeof(elv));
    for (int k = 0; k < 10; k++) {
      int64_t n3 = 3 * (nodes[k] - 1);
      for (int c = 0; c < 3; c++) {
        du10[k][c] = du[n3 + c];
        xlv[k][c] = nocoord[n3 + c] + (LD ? disp_new[n3 + c] : 0.0);
      }
    }
    for (int i = 0; i < 4; i++) {
      int64_t ipp = 4 * el + i, ippos = elpos + 6 * i;
      double sy = sig_yield[ipp];
      double xsj = dshp10tet(GP10[i][0], GP10[i][1], GP10[i][2], xlv, NULL, dshpg);
      double deps[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 0; j < 10; j++) {
        double d0 = dshpg[0][j], d1 = dshpg[1][j], d2 = dshpg[2][j];
        deps[0] += d0 * du10[j][0];
        deps[1] += d1 * du10[j][1];
        deps[2] += d2 * du10[j][2];
        deps[3] += d1 * du10[j][0] + d0 * du10[j][1];
        deps[4] += d2 * du10[j][0] + d0 * du10[j][2];
        deps[5] += d2 * du10[j][1] + d1 * du10[j][2];
      }
      double sigc[6];
      if (LD) {
        /* convected stress, fcVM.py:2383-2429 */
        double st[3][3] = {{sig[ippos], sig[ippos + 3], sig[ippos + 4]},
                           {sig[ippos + 3], sig[ippos + 1], sig[ippos + 5]},
                           {sig[ippos + 4], sig[ippos + 5], sig[ippos + 2]}};
        double F[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int k = 0; k < 10; k++)
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) F[a][b] += du10[k][a] * dshpg[b][k];
        double rr = (F[0][0] * F[1][1] * F[2][2] - F[0][0] * F[1][2] * F[2][1] + F[0][2] * F[1][0] * F[2][1] -
                     F[0][2] * F[1][1] * F[2][0] + F[0][1] * F[1][2] * F[2][0] - F[0][1] * F[1][0] * F[2][2]);
        rr = 1.0 / rr;
        double sc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int a = 0; a < 3; a++)
          for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++)
              for (int l = 0; l < 3; l++) sc[a][k] += F[a][j] * st[j][l] * F[k][l];
        sigc[0] = rr * sc[0][0];
        sigc[1] = rr * sc[1][1];
        sigc[2] = rr * sc[2][2];
        sigc[3] = rr * sc[0][1];
        sigc[4] = rr * sc[0][2];
        sigc[5] = rr * sc[1][2];
      } else {
        for (int c = 0; c < 6; c++) sigc[c] = sig[ippos + c];
      }
      double sig_test[6], s[6];
      for (int j = 0; j < 6; j++) {
        double tmp = sigc[j];
        for (int k = 0; k < 6; k++) tmp += dmat[j][k] * deps[k];
        sig_test[j] = tmp;
      }
      for (int c = 0; c < 6; c++) sig_test_global[ippos + c] = sig_test[c];
      pgp[ipp] = (uint8_t)vmises(sig_test, sy, H, G, s);
      for (int c = 0; c < 6; c++) sig_update[ippos + c] = s[c];
      double ipxsj = GP10[i][3] * fabs(xsj);
      double sxx = s[0], syy = s[1], szz = s[2], sxy = s[3], szx = s[4], syz = s[5];
      for (int j = 0; j < 10; j++) {
        double d0 = dshpg[0][j], d1 = dshpg[1][j], d2 = dshpg[2][j];
        elv[3 * j] += (d0 * sxx + d1 * sxy + d2 * szx) * ipxsj;
        elv[3 * j + 1] += (d1 * syy + d0 * sxy + d2 * syz) * ipxsj;
        elv[3 * j + 2] += (d2 * szz + d0 * szx + d1 * syz) * ipxsj;
      }
    }
    for (int i = 0; i < 10; i++) {
      int64_t iglob3 = 3 * (nodes[i] - 1);
      for (int k = 0; k < 3; k++) qin[iglob3 + k] += elv[3 * i + k];
    }
  }
}

/* ---- fcVM.py:2084-2137 (update_PEEQ_CSR) ----------------------------------- */
void fcvm_oracle_update_peeq_csr(int64_t nelem, double E, double nu, const double *sig_test, const double *sig_new,
                                 double *sig_yield, double ultimate_strain, double *peeq, double *csr, double *triax,
                                 double *pressure, double *sigmises, double *ecr, double Et_E) {
  double G = E / 2.0 / (1 + nu);
  if (Et_E > 0.95) Et_E = 0.95;
  double Et = Et_E * E;
  double H = Et / (1.0 - Et_E);
  if (ultimate_strain == 0.0) ultimate_strain = 1.0e12;
  double alpha = sqrt(M_E) * ultimate_strain;
  double beta = 1.5;
  for (int64_t el = 0; el < nelem; el++)
    for (int ip = 0; ip < 4; ip++) {
      int64_t ipos1 = 4 * el + ip, ipos2 = 24 * el + 6 * ip;
      double st0 = sig_test[ipos2], st1 = sig_test[ipos2 + 1], st2 = sig_test[ipos2 + 2];
      double st3 = sig_test[ipos2 + 3], st4 = sig_test[ipos2 + 4], st5 = sig_test[ipos2 + 5];
      double sn0 = sig_new[ipos2], sn1 = sig_new[ipos2 + 1], sn2 = sig_new[ipos2 + 2];
      double sn3 = sig_new[ipos2 + 3], sn4 = sig_new[ipos2 + 4], sn5 = sig_new[ipos2 + 5];
      double p_t = (st0 + st1 + st2) / 3.0, p_n = (sn0 + sn1 + sn2) / 3.0;
      st0 -= p_t;
      st1 -= p_t;
      st2 -= p_t;
      sn0 -= p_n;
      sn1 -= p_n;
      sn2 -= p_n;
      double smt = sqrt(1.5 * (st0 * st0 + st1 * st1 + st2 * st2) + 3.0 * (st3 * st3 + st4 * st4 + st5 * st5));
      double smn = sqrt(1.5 * (sn0 * sn0 + sn1 * sn1 + sn2 * sn2) + 3.0 * (sn3 * sn3 + sn4 * sn4 + sn5 * sn5));
      double DL = 0.0;
      if (smt > sig_yield[ipos1]) {
        DL = (smt - sig_yield[ipos1]) / (3.0 * G + H);
        peeq[ipos1] += DL;
        sig_yield[ipos1] += Et * DL;
      }
      double T = p_n / sig_yield[ipos1];
      pressure[ipos1] = p_n;
      sigmises[ipos1] = smn;
      triax[ipos1] = T;
      double critical_strain = alpha * exp(-beta * T);
      if (critical_strain < 1.0e-6) critical_strain = 1.0e-6;
      ecr[ipos1] = critical_strain;
      csr[ipos1] += DL / critical_strain;
    }
}

/* ---- fcVM.py:2496-2554 (mapStresses) ---------------------------------------- */
void fcvm_oracle_map_stresses(int averaged, int64_t ne, int64_t nn, const int64_t *elNodes, const double *sig,
                              const double *peeq, const double *sigvm, const double *csr, const int16_t *noce,
                              double sig_yield, double *tet10stress, double *tet10peeq, double *tet10csr,
                              double *tet10svm, double *tet10triax) {
  static const double map_inter[6][4] = {{0.5, 0.5, 0.0, 0.0}, {0.0, 0.5, 0.5, 0.0}, {0.5, 0.0, 0.5, 0.0},
                                         {0.5, 0.0, 0.0, 0.5}, {0.0, 0.5, 0.0, 0.5}, {0.0, 0.0, 0.5, 0.5}};
  memset(tet10stress, 0, sizeof(double) * 6 * nn);
  memset(tet10peeq, 0, sizeof(double) * nn);
  memset(tet10csr, 0, sizeof(double) * nn);
  memset(tet10svm, 0, sizeof(double) * nn);
  memset(tet10triax, 0, sizeof(double) * nn);
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    for (int k = 0; k < 4; k++) {
      int64_t nd = nodes[k] - 1;
      for (int c = 0; c < 6; c++) tet10stress[6 * nd + c] += sig[24 * el + 6 * k + c] / noce[nd];
    }
  }
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    for (int k = 0; k < 4; k++) {
      int64_t nd = nodes[k] - 1, g = 4 * el + k;
      double tr = (sig[24 * el + 6 * k] + sig[24 * el + 6 * k + 1] + sig[24 * el + 6 * k + 2]) / 3.0 / sig_yield;
      if (averaged) {
        tet10peeq[nd] += peeq[g] / noce[nd];
        tet10csr[nd] += csr[g] / noce[nd];
        tet10svm[nd] += sigvm[g] / noce[nd];
        tet10triax[nd] += tr / noce[nd];
      } else {
        tet10peeq[nd] = fmax(tet10peeq[nd], peeq[g]);
        tet10csr[nd] = fmax(tet10csr[nd], csr[g]);
        tet10svm[nd] = fmax(tet10svm[nd], sigvm[g]);
        tet10triax[nd] = fmax(tet10triax[nd], tr);
      }
    }
  }
  for (int64_t el = 0; el < ne; el++) {
    const int64_t *nodes = elNodes + 10 * el;
    for (int m = 0; m < 6; m++) {
      int64_t ni = nodes[4 + m] - 1;
      double s[6] = {0, 0, 0, 0, 0, 0}, a = 0, b = 0, c = 0, d = 0;
      for (int k = 0; k < 4; k++) {
        int64_t nc = nodes[k] - 1;
        double wgt = map_inter[m][k];
        for (int q = 0; q < 6; q++) s[q] += wgt * tet10stress[6 * nc + q];
        a += wgt * tet10peeq[nc];
        b += wgt * tet10csr[nc];
        c += wgt * tet10svm[nc];
        d += wgt * tet10triax[nc];
      }
      for (int q = 0; q < 6; q++) tet10stress[6 * ni + q] = s[q];
      tet10peeq[ni] = a;
      tet10csr[ni] = b;
      tet10svm[ni] = c;
      tet10triax[ni] = d;
    }
  }
}
