// Device scalar slots of the PCG solve (ctx->red_out), shared by the launch-per-phase solver
// (fcvm_pcg.cu) and its helpers.
#pragma once

namespace fcvm {

// gamma = r.u, rr = r.r and alpha live in pairs indexed by the parity of the iteration; L_* are per-rank
// partial sums on their way through the scalar all-reduce when a communicator is attached.
enum {
  S_GAMMA = 0,   // [2]
  S_RR = 2,      // [2]
  S_ALPHA = 4,   // [2]
  S_DELTA = 6, S_BB = 7, S_THR = 8, S_ITERS = 9, L_RU = 10, L_RR = 11, L_WU = 12, L_BB = 13,
  S_STATUS = 14  // 0 running, 1 converged, 2 iteration limit, 3 breakdown (operator or preconditioner not positive definite)
};

enum { PCG_RUNNING = 0, PCG_CONVERGED = 1, PCG_MAXITER = 2, PCG_BREAKDOWN = 3, PCG_COMM_TIMEOUT = 4 };

}  // namespace fcvm
