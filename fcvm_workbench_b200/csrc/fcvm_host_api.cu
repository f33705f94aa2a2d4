// HOST-buffer entry points with the reference routines' argument lists: copy in, run the
// same device kernels, copy out.  These are what a ctypes binding inside fcVM.py calls.
#include <algorithm>

#include "fcvm_common.cuh"

using namespace fcvm;

namespace {
int ensure_vec(fcvm_ctx *c, double **v, int64_t n) {
  if (*v) return FCVM_OK;
  return fcvm_vec_alloc(c, n, v);
}
}  // namespace

namespace fcvm {
int launch_stress_tiles(fcvm_ctx *c, const double *disp_new, const double *du, double Et_E, int LD, double yield_scale,
                        int64_t tile0, int64_t ntiles);
int launch_node_gather(fcvm_ctx *c, double *out, int accumulate);
int launch_gp_in(fcvm_ctx *c, int64_t a, int64_t b, int ncomp, const double *aos, double *soa);
int launch_gp_out(fcvm_ctx *c, int64_t a, int64_t b, int ncomp, const double *soa, double *aos);
int launch_pgp_out(fcvm_ctx *c, int64_t a, int64_t b, const uint8_t *soa, uint8_t *aos);
}  // namespace fcvm

// update_stress_load(gp10, elNodes, nocoord, materialbyElement, sig_yield, disp_new, du, sig,
//                    sig_update, sig_test_global, qin, Et_E, LD, pgp)         fcVM.py:2196
//
// The call moves 0.8 GB over PCIe at 1M elements (the reference's Gauss-point arrays live on the host), four times
// what the kernels cost.  It is therefore a three-stage pipeline over chunks of elements: chunk k+1 is copied in on
// one stream while chunk k is converted, updated and converted back on the compute stream and chunk k-1 is copied
// out on a third stream -- host->device and device->host copies run at the same time on the two copy engines.
extern "C" int fcvm_host_update_stress_load(fcvm_ctx *c, const double *sig_yield, const double *disp_new,
                                            const double *du, const double *sig, double *sig_update,
                                            double *sig_test_global, double *qin, double Et_E, int LD,
                                            uint8_t *pgp) {
  FCVM_CHECK(c && c->ne > 0 && sig_yield && du && sig && sig_update && sig_test_global && qin && pgp, FCVM_E_ARG,
             "fcvm_host_update_stress_load: null argument / no mesh");
  const int64_t ne = c->ne, n3 = 3 * c->nn;
  cudaStream_t st = c->stream;
  FCVM_TRY(ensure_vec(c, &c->h_du, n3));
  FCVM_TRY(ensure_vec(c, &c->h_disp, n3));
  FCVM_TRY(ensure_vec(c, &c->h_qin, n3));
  if (!c->hs_in) {
    FCVM_CUDA(cudaMalloc((void **)&c->hs_in, sizeof(double) * 28 * (size_t)ne));
    FCVM_CUDA(cudaMalloc((void **)&c->hs_out, sizeof(double) * 48 * (size_t)ne));
    FCVM_CUDA(cudaMalloc((void **)&c->hs_pgp, 4 * (size_t)ne));
    FCVM_CUDA(cudaStreamCreateWithFlags(&c->h_in_stream, cudaStreamNonBlocking));
    FCVM_CUDA(cudaStreamCreateWithFlags(&c->h_out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; i++) {
      FCVM_CUDA(cudaEventCreateWithFlags(&c->h_ev_in[i], cudaEventDisableTiming));
      FCVM_CUDA(cudaEventCreateWithFlags(&c->h_ev_k[i], cudaEventDisableTiming));
    }
  }
  double *in_sig = c->hs_in, *in_sy = c->hs_in + 24 * ne, *out_new = c->hs_out, *out_test = c->hs_out + 24 * ne;
  // nodal vectors first (small), on the compute stream
  FCVM_CUDA(cudaMemcpyAsync(c->h_du, du, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
  if (disp_new) FCVM_CUDA(cudaMemcpyAsync(c->h_disp, disp_new, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
  const int64_t tiles = (ne + 31) / 32;
  const int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(8, tiles / 512));   // >= 16k elements per chunk
  // everything issued before must be done before the side streams touch the staging of a previous call
  FCVM_CUDA(cudaStreamSynchronize(c->h_out_stream));
  for (int k = 0; k < nchunk; k++) {
    const int64_t a = std::min(ne, 32 * (tiles * k / nchunk)), b = std::min(ne, 32 * (tiles * (k + 1) / nchunk));
    FCVM_CUDA(cudaMemcpyAsync(in_sig + 24 * a, sig + 24 * a, sizeof(double) * 24 * (b - a), cudaMemcpyHostToDevice, c->h_in_stream));
    FCVM_CUDA(cudaMemcpyAsync(in_sy + 4 * a, sig_yield + 4 * a, sizeof(double) * 4 * (b - a), cudaMemcpyHostToDevice, c->h_in_stream));
    FCVM_CUDA(cudaEventRecord(c->h_ev_in[k], c->h_in_stream));
  }
  for (int k = 0; k < nchunk; k++) {
    const int64_t t0 = tiles * k / nchunk, t1 = tiles * (k + 1) / nchunk;
    const int64_t a = std::min(ne, 32 * t0), b = std::min(ne, 32 * t1);
    FCVM_CUDA(cudaStreamWaitEvent(st, c->h_ev_in[k], 0));
    FCVM_TRY(launch_gp_in(c, a, b, 6, in_sig, (double *)c->buf[FCVM_BUF_SIG_OLD]));
    FCVM_TRY(launch_gp_in(c, a, b, 1, in_sy, (double *)c->buf[FCVM_BUF_SIG_YIELD]));
    FCVM_TRY(launch_stress_tiles(c, c->h_disp, c->h_du, Et_E, LD, 1.0, t0, t1 - t0));
    FCVM_TRY(launch_gp_out(c, a, b, 6, (const double *)c->buf[FCVM_BUF_SIG_NEW], out_new));
    FCVM_TRY(launch_gp_out(c, a, b, 6, (const double *)c->buf[FCVM_BUF_SIG_TEST], out_test));
    FCVM_TRY(launch_pgp_out(c, a, b, (const uint8_t *)c->buf[FCVM_BUF_PGP], c->hs_pgp));
    FCVM_CUDA(cudaEventRecord(c->h_ev_k[k], st));
    FCVM_CUDA(cudaStreamWaitEvent(c->h_out_stream, c->h_ev_k[k], 0));
    FCVM_CUDA(cudaMemcpyAsync(sig_update + 24 * a, out_new + 24 * a, sizeof(double) * 24 * (b - a), cudaMemcpyDeviceToHost, c->h_out_stream));
    FCVM_CUDA(cudaMemcpyAsync(sig_test_global + 24 * a, out_test + 24 * a, sizeof(double) * 24 * (b - a), cudaMemcpyDeviceToHost, c->h_out_stream));
    FCVM_CUDA(cudaMemcpyAsync(pgp + 4 * a, c->hs_pgp + 4 * a, 4 * (size_t)(b - a), cudaMemcpyDeviceToHost, c->h_out_stream));
  }
  // internal force: gather of all element vectors, shared nodes, then the reference's accumulation into the qin it
  // is given (fcVM.py:2462): q = qin + assembled forces
  FCVM_TRY(launch_node_gather(c, c->h_qin, 0));
  FCVM_TRY(fcvm_interface_sum(c, c->h_qin));
  FCVM_CUDA(cudaMemcpyAsync(c->h_du, qin, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
  FCVM_TRY(fcvm_vec_axpby(c, n3, 1.0, c->h_du, 1.0, c->h_qin));
  FCVM_CUDA(cudaMemcpyAsync(qin, c->h_qin, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
  FCVM_CUDA(cudaStreamSynchronize(st));
  FCVM_CUDA(cudaStreamSynchronize(c->h_out_stream));
  c->h2d_bytes += (int64_t)sizeof(double) * (28 * ne + (disp_new ? 3 : 2) * n3);
  c->d2h_bytes += (int64_t)sizeof(double) * (48 * ne + n3) + 4 * ne;
  return FCVM_OK;
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
