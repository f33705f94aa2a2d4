// HOST-buffer entry points with the reference routines' argument lists: copy in, run the
// same device kernels, copy out.  These are what a ctypes binding inside fcVM.py calls.
#include "fcvm_common.cuh"

using namespace fcvm;

namespace {
int ensure_vec(fcvm_ctx *c, double **v, int64_t n) {
  if (*v) return FCVM_OK;
  return fcvm_vec_alloc(c, n, v);
}
}  // namespace

// update_stress_load(gp10, elNodes, nocoord, materialbyElement, sig_yield, disp_new, du, sig,
//                    sig_update, sig_test_global, qin, Et_E, LD, pgp)         fcVM.py:2196
extern "C" int fcvm_host_update_stress_load(fcvm_ctx *c, const double *sig_yield, const double *disp_new,
                                            const double *du, const double *sig, double *sig_update,
                                            double *sig_test_global, double *qin, double Et_E, int LD,
                                            uint8_t *pgp) {
  FCVM_CHECK(c && c->ne > 0 && sig_yield && du && sig && sig_update && sig_test_global && qin && pgp, FCVM_E_ARG,
             "fcvm_host_update_stress_load: null argument / no mesh");
  const int64_t n3 = 3 * c->nn;
  FCVM_TRY(ensure_vec(c, &c->h_du, n3));
  FCVM_TRY(ensure_vec(c, &c->h_disp, n3));
  FCVM_TRY(ensure_vec(c, &c->h_qin, n3));
  FCVM_TRY(fcvm_gp_from_host(c, sig_yield, 1, (double *)c->buf[FCVM_BUF_SIG_YIELD]));
  FCVM_TRY(fcvm_gp_from_host(c, sig, 6, (double *)c->buf[FCVM_BUF_SIG_OLD]));
  FCVM_TRY(fcvm_h2d(c, c->h_du, du, sizeof(double) * n3));
  if (disp_new) FCVM_TRY(fcvm_h2d(c, c->h_disp, disp_new, sizeof(double) * n3));
  // the reference accumulates into the qin it is given (fcVM.py:2462): q = qin + assembled forces
  FCVM_TRY(fcvm_update_stress_load(c, c->h_disp, c->h_du, c->h_qin, Et_E, LD, 1.0));
  FCVM_TRY(fcvm_h2d(c, c->h_du, qin, sizeof(double) * n3));
  FCVM_TRY(fcvm_vec_axpby(c, n3, 1.0, c->h_du, 1.0, c->h_qin));
  FCVM_TRY(fcvm_gp_to_host(c, (const double *)c->buf[FCVM_BUF_SIG_NEW], 6, sig_update));
  FCVM_TRY(fcvm_gp_to_host(c, (const double *)c->buf[FCVM_BUF_SIG_TEST], 6, sig_test_global));
  FCVM_TRY(fcvm_pgp_to_host(c, pgp));
  return fcvm_d2h(c, qin, c->h_qin, sizeof(double) * n3);
}

// x = factor(b)                                                                fcVM.py:1130, 1401
extern "C" int fcvm_host_solve(fcvm_ctx *c, const double *b, double *x, double rtol, int max_iter, int recycle, int *iters,
                               double *relres) {
  FCVM_CHECK(c && c->assembled && b && x, FCVM_E_ARG, "fcvm_host_solve: assemble first / null argument");
  const int64_t n3 = 3 * c->nn;
  FCVM_TRY(ensure_vec(c, &c->h_du, n3));
  FCVM_TRY(ensure_vec(c, &c->h_qin, n3));
  FCVM_TRY(fcvm_h2d(c, c->h_du, b, sizeof(double) * n3));
  int rc = fcvm_pcg_solve(c, c->h_du, c->h_qin, rtol, max_iter, recycle ? 2 : 0, iters, relres);
  if (rc != FCVM_OK && rc != FCVM_E_NOCONV) return rc;
  FCVM_TRY(fcvm_d2h(c, x, c->h_qin, sizeof(double) * n3));
  return rc;
}
