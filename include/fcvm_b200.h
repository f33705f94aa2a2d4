/*
 * fcvm_b200.h -- C ABI of the B200-native fcVM Newton / load-stepping hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * Each entry point names the reference routine it replaces
 * (HarryvL/fcVM-workbench, "source code/fcVM.py", cited as fcVM.py:line).
 * A maintainer binds these with ctypes from fcVM.py itself; INTEGRATION.md shows
 * the stub.  All functions return 0 on success and a negative FCVM_E_* code on
 * failure; fcvm_last_error() gives the message.  There is no CPU fallback: when
 * no CUDA device is usable every compute call fails with FCVM_E_CUDA.
 *
 * Conventions shared with the reference
 *   - node numbers in elNodes are 1-based, local order as after setUpInput's swap
 *     (fcVM.py:338-341); dof = 3*(node-1)+component (fcVM.py:230)
 *   - nodal vectors (du, disp, qin, glv, ...) are interleaved, length 3*nn
 *   - Gauss-point arrays on the HOST side use the reference layout:
 *     stress[24*el + 6*ip + c], scalar[4*el + ip] (fcVM.py:2237, 2269-2272)
 *   - on the DEVICE Gauss-point arrays are structure-of-arrays:
 *     stress[(c*4 + ip)*ne + el], scalar[ip*ne + el]; fcvm_gp_* convert.
 *
 * Two families of calls:
 *   fcvm_*            operate on device-resident state owned by the context
 *                     (device pointers are plain `double*` obtained from
 *                     fcvm_vec_alloc / fcvm_buf); nothing crosses PCIe.
 *   fcvm_host_*       take HOST buffers with the reference routine's argument
 *                     list, copy in, run the same kernels, copy out.
 */
#ifndef FCVM_B200_H
#define FCVM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fcvm_ctx fcvm_ctx;

enum {
  FCVM_OK = 0,
  FCVM_E_ARG = -1,     /* bad argument / call order */
  FCVM_E_CUDA = -2,    /* CUDA runtime error (no device, out of memory, launch failure) */
  FCVM_E_MESH = -3,    /* inconsistent mesh (node number outside 1..nn).  An element with a non-positive Jacobian
                          is NOT rejected: like the reference (fcVM.py:450, abs(xsj)) it is integrated with |J| */
  FCVM_E_NOCONV = -4,  /* PCG hit max_iter before reaching the tolerance */
  FCVM_E_NCCL = -5,    /* NCCL / peer-memory failure (multi-GPU) */
  FCVM_E_INDEFINITE = -6 /* PCG breakdown: the matrix is not positive definite (the reference's 'singular stiffness
                          matrix', fcVM.py:1367-1381) */
};

/* Named device buffers owned by the context (fcvm_buf). */
enum {
  FCVM_BUF_SIG_OLD = 0,  /* 24*ne  stress at the start of the step        (sig_old,  fcVM.py:1148) */
  FCVM_BUF_SIG_NEW = 1,  /* 24*ne  updated stress                         (sig_new,  fcVM.py:1147) */
  FCVM_BUF_SIG_TEST = 2, /* 24*ne  elastic test stress                    (sig_test, fcVM.py:1150) */
  FCVM_BUF_SIG_YIELD = 3,/* 4*ne   current yield stress                   (sig_yield,fcVM.py:1149) */
  FCVM_BUF_PEEQ = 4,     /* 4*ne   equivalent plastic strain              (fcVM.py:1151) */
  FCVM_BUF_CSR = 5,      /* 4*ne   critical strain ratio                  (fcVM.py:1157) */
  FCVM_BUF_TRIAX = 6,    /* 4*ne                                           (fcVM.py:1152) */
  FCVM_BUF_PRESSURE = 7, /* 4*ne                                           (fcVM.py:1153) */
  FCVM_BUF_SIGMISES = 8, /* 4*ne                                           (fcVM.py:1154) */
  FCVM_BUF_ECR = 9,      /* 4*ne   critical plastic strain                (fcVM.py:1156) */
  FCVM_BUF_PGP = 10,     /* 4*ne   uint8 plastic flag                     (pgp, fcVM.py:1155) */
  FCVM_BUF_MODF = 11,    /* 3*nn   rhs modification for prescribed dofs   (modf, fcVM.py:641) */
  FCVM_BUF_GLV = 12,     /* 3*nn   global load vector                     (glv,  fcVM.py:637) */
  FCVM_BUF_FIXDOF = 13,  /* 3*nn   double 1.0 = free, 0.0 = prescribed    (fixdof, fcVM.py:224) */
  FCVM_BUF_COUNT = 14
};

const char *fcvm_last_error(void);
int fcvm_version(void);

/* ---- context, mesh and constraints ----------------------------------------------------- */
int fcvm_create(fcvm_ctx **out, int device);
int fcvm_destroy(fcvm_ctx *ctx);
/* Launch every kernel of this context on an existing CUDA stream (cudaStream_t as void*);
 * NULL returns to the context's own stream. */
int fcvm_set_stream(fcvm_ctx *ctx, void *cuda_stream);
int fcvm_synchronize(fcvm_ctx *ctx);

/* The arrays setUpInput returns (fcVM.py:343): connectivity, coordinates, material of
 * element 0 (the reference uses one material for the whole mesh, fcVM.py:736-737).  Builds
 * the node->element map and the sparsity pattern (fixed once per mesh). */
int fcvm_set_mesh(fcvm_ctx *ctx, int64_t ne, int64_t nn, const int64_t *elNodes, const double *nocoord, double E,
                  double nu, double density);
/* `fix` of fcVM.py:222 as dense arrays over the 3*nn dofs: mask[d] != 0 -> dof d prescribed to val[d]. */
int fcvm_set_constraints(fcvm_ctx *ctx, const uint8_t *fixmask, const double *fixval);
/* Multi-GPU (element-partitioned): for every local node the number of ranks that hold it
 * (NULL = 1 everywhere) and, for the n_if local nodes shared with other ranks, their index in
 * the global interface vector.  See fcvm_comm_init. */
int fcvm_set_interface(fcvm_ctx *ctx, const double *dof_weight, int64_t n_if_local, const int64_t *if_local_node,
                       const int64_t *if_global_slot, int64_t n_if_global);

/* New nodal coordinates on the same topology (nocoord += imper, fcVM.py:1240); assemble again afterwards. */
int fcvm_set_coordinates(fcvm_ctx *ctx, const double *nocoord);

int64_t fcvm_num_elements(const fcvm_ctx *ctx);
int64_t fcvm_num_nodes(const fcvm_ctx *ctx);

/* ---- device vectors -------------------------------------------------------------------- */
int fcvm_vec_alloc(fcvm_ctx *ctx, int64_t n, double **out);   /* zero-filled */
int fcvm_vec_free(fcvm_ctx *ctx, double *v);
int fcvm_buf(fcvm_ctx *ctx, int which, void **out, int64_t *n);
int fcvm_h2d(fcvm_ctx *ctx, void *dst_dev, const void *src_host, int64_t bytes);
int fcvm_d2h(fcvm_ctx *ctx, void *dst_host, const void *src_dev, int64_t bytes);
int fcvm_vec_zero(fcvm_ctx *ctx, int64_t n, double *x);
int fcvm_vec_copy(fcvm_ctx *ctx, int64_t n, const double *x, double *y);                 /* y = x       */
int fcvm_vec_axpby(fcvm_ctx *ctx, int64_t n, double a, const double *x, double b, double *y); /* y = a x + b y */
int fcvm_vec_axpbypcz(fcvm_ctx *ctx, int64_t n, double a, const double *x, double b, const double *y, double c,
                      double *z);                                                         /* z = a x + b y + c z */
/* Deterministic dot product (fixed-shape two-stage reduction; weighted by the interface
 * multiplicity and summed over ranks when a communicator is attached). */
int fcvm_vec_dot(fcvm_ctx *ctx, int64_t n, const double *x, const double *y, double *out);
/* r = fixdof * (lbd * glv - qin), returns ||r||_2   (fcVM.py:1329-1338, 1446-1447) */
int fcvm_residual(fcvm_ctx *ctx, double lbd, const double *glv, const double *qin, double *r, double *rnorm);
/* sqrt(max over nodes 0..nn-2 of |u_node|^2): the `un` of fcVM.py:1494-1497 (the reference's
 * range((ndof-1)//3) leaves the last node out; kept). */
int fcvm_max_node_disp(fcvm_ctx *ctx, const double *disp, double *out);
/* sum(movdof * qin): reaction on the moving boundary (fcVM.py:1523) */
int fcvm_reaction(fcvm_ctx *ctx, const double *qin, double *out);

/* Gauss-point layout conversion, device <-> host(reference layout). ncomp = 6 or 1. */
int fcvm_gp_to_host(fcvm_ctx *ctx, const double *dev_soa, int ncomp, double *host_aos);
int fcvm_gp_from_host(fcvm_ctx *ctx, const double *host_aos, int ncomp, double *dev_soa);
int fcvm_gp_fill(fcvm_ctx *ctx, int which, double value);
int fcvm_pgp_to_host(fcvm_ctx *ctx, uint8_t *host);
int fcvm_pgp_count(fcvm_ctx *ctx, int64_t *n_plastic);

/* ---- stiffness: calcGSM (fcVM.py:620-816) / calcTSM nstep>1 (fcVM.py:819-1079) ------------ */
/* Element matrices by Gauss-point integration, deterministic COO->SELL reduction, constraint
 * elimination and rhs modification `modf`; gravity added to `glv` (which must already hold the
 * surface loads).  tangent != 0 integrates D - pmat at plastic Gauss points of SIG_OLD/PGP on
 * the geometry nocoord + disp (disp may be NULL). */
int fcvm_assemble(fcvm_ctx *ctx, int tangent, const double *disp, double Et_E, double grav_x, double grav_y,
                  double grav_z, double *glv);
/* Linear buckling analysis (calcTSM with nstep == 1, fcVM.py:1002-1006, 1063-1073; used at fcVM.py:1199-1212):
 * K = elastic stiffness, not eliminated, diagonal entries of prescribed dofs x 100; G = -(geometric stiffness of the
 * stress state in SIG_NEW).  After the call the context's matrix is K - sigma G (fcvm_pcg_solve / fcvm_spmv work on
 * it) and fcvm_spmv_geometric multiplies with G: the two operators of the shift-invert eigen-iteration that stands
 * in for eigsh(K, k, M=G, sigma, mode='buckling').  Single GPU. */
int fcvm_assemble_buckling(fcvm_ctx *ctx, double sigma);
int fcvm_spmv_geometric(fcvm_ctx *ctx, const double *x, double *y);
/* Raw element matrices (ne*900 doubles, device) for element-level parity checks. */
int fcvm_element_matrices(fcvm_ctx *ctx, int tangent, const double *disp, double Et_E, double *esm_dev);
/* Lower-triangular CSC of the assembled matrix exactly as scipy builds it at fcVM.py:1111:
 * call with indices == NULL to get nnz, then with host buffers. */
int fcvm_export_csc_lower(fcvm_ctx *ctx, int64_t *nnz, int64_t *indptr, int64_t *indices, double *data);
/* y = K x with the assembled matrix (block-SELL SpMV). */
int fcvm_spmv(fcvm_ctx *ctx, const double *x, double *y);
/* y = K x with calcGSM's elastic operator (constraints eliminated as in fcvm_assemble) recomputed element by
 * element from the mesh instead of read from the assembled matrix: the product the PCG uses in the geometrically
 * linear analysis.  Needs the mesh and the constraints only; equals fcvm_spmv after an elastic fcvm_assemble to
 * round-off. */
int fcvm_matfree_apply(fcvm_ctx *ctx, const double *x, double *y);

/* ---- linear solve: replaces factor = cholesky(gsm); x = factor(b) (fcVM.py:1121-1135, 1401) -- */
/* Preconditioned CG on the device (single-reduction form; block-Jacobi, plus the deflation level below
 * when switched on).  x is overwritten; use_x0 = 0: start from zero, 1: from the x passed in, 2: from the Galerkin
 * projection of b onto the last two solutions obtained with the present matrix (kept on the device; for the repeated
 * solves of the modified Newton iteration, fcVM.py:1401) -- and remember this solution.  Converged when the
 * recursively updated residual satisfies ||r|| <= rtol * ||b||; returns FCVM_E_NOCONV after max_iter. */
int fcvm_pcg_solve(fcvm_ctx *ctx, const double *b, double *x, double rtol, int max_iter, int use_x0, int *iters,
                   double *relres);

/* Second preconditioner level (optional): deflation of the rigid-body modes of box clusters of nodes.
 * The clusters form an ncx x ncy x ncz grid of boxes lo + (i,j,k)*h over the (global) bounding box;
 * cid[n] = ix + ncx*(iy + ncy*iz) is the box of local node n.  A box must be at least two elements
 * wide in every direction (a node then couples to at most 2 x 2 x 2 boxes).  Takes effect at the next fcvm_assemble (K Z and (Z^T K Z)^-1 are built
 * there); ncx = 0 switches it off.  6*ncx*ncy*ncz <= 16384.  active[box] = 0 (optional, NULL = all 1)
 * drops the modes of a box that holds too few free nodes for six independent rigid-body modes. */
int fcvm_set_deflation(fcvm_ctx *ctx, int ncx, int ncy, int ncz, const int32_t *cid, const double *lo,
                       const double *h, const uint8_t *active);

/* Boxes and stored (node, box) entries of K Z, for the roofline arithmetic of the coarse kernels. */
int fcvm_deflation_stats(fcvm_ctx *ctx, int64_t *boxes, int64_t *entries);

/* ---- stress update: update_stress_load (fcVM.py:2196-2464) ---------------------------------- */
/* Reads SIG_OLD / SIG_YIELD, writes SIG_NEW / SIG_TEST / PGP, and qin = internal force vector
 * (overwritten, not accumulated: the reference always passes zeros, fcVM.py:1324, 1441).
 * yield_scale multiplies SIG_YIELD on the fly (the 1.0e6 of fcVM.py:1195). */
int fcvm_update_stress_load(fcvm_ctx *ctx, const double *disp_new, const double *du, double *qin, double Et_E,
                            int LD, double yield_scale);
/* update_PEEQ_CSR (fcVM.py:2084-2137) on SIG_TEST / SIG_NEW; also returns max(csr), its Gauss
 * point (reference numbering 4*el+ip, first maximum), the state there and max(peeq)
 * (fcVM.py:1546-1554). out7 = {csr_max, pressure, sigmises, triax, ecr, peeq, peeq_max}. */
int fcvm_update_peeq_csr(fcvm_ctx *ctx, double ultimate_strain, double Et_E, int64_t *argmax_gp, double *out7);
/* sig_new = sig_old + fac (sig_new - sig_old), same for sig_test (fcVM.py:1490-1491) */
int fcvm_scale_step_stress(fcvm_ctx *ctx, double fac);
/* mapStresses (fcVM.py:2496-2554): host outputs tet10stress (nn*6) and four nn-vectors. */
int fcvm_map_stresses(fcvm_ctx *ctx, int averaged, double sig_yield, const int16_t *noce, double *tet10stress,
                      double *tet10peeq, double *tet10csr, double *tet10svm, double *tet10triax);

/* ---- multi-GPU: one process per GPU, element-partitioned, NCCL over NVLink ------------------- */
/* unique_id: 128 bytes from fcvm_comm_unique_id on rank 0, broadcast by the caller. */
int fcvm_comm_unique_id(void *id128);
int fcvm_comm_init(fcvm_ctx *ctx, const void *id128, int rank, int world);
int fcvm_comm_allreduce_sum(fcvm_ctx *ctx, double *dev, int64_t n);
int fcvm_comm_allreduce_max(fcvm_ctx *ctx, double *dev, int64_t n);
/* Number of leading local nodes that enter fcvm_max_node_disp: nn_local, minus one on the rank that
 * holds the last global node (the reference leaves it out, fcVM.py:1494-1497). */
int fcvm_set_un_nodes(fcvm_ctx *ctx, int64_t n);
/* v[shared nodes] = sum over ranks (interface exchange used after SpMV and after the
 * internal-force gather). */
int fcvm_interface_sum(fcvm_ctx *ctx, double *v);

/* Peer-memory exchanges inside one box (CUDA IPC over NVLink): the halo of the PCG product between neighbouring
 * ranks and the small all-to-all sums of the iteration, each ONE kernel that stores into the peers' mapped memory,
 * flags, waits and adds in rank order -- instead of NCCL all-reduces over a dense global interface vector.
 * fcvm_p2p_create allocates this rank's arena (room for n_recv_nodes received rows and slots of slot_n doubles)
 * and returns its 64-byte IPC handle; fcvm_p2p_attach maps the peers' arenas (handles in rank order, world x 64
 * bytes) and takes the exchange lists of Partition.p2p_plan.  Optional: without it the NCCL path is used. */
int fcvm_p2p_create(fcvm_ctx *ctx, int64_t n_recv_nodes, int64_t slot_n, void *handle64);
int fcvm_p2p_attach(fcvm_ctx *ctx, const void *handles, int npeers, const int32_t *peer_rank, const int32_t *send_ptr,
                    const int32_t *send_node, const int64_t *remote_off, int n_if, const int32_t *if_node,
                    const int32_t *if_ptr, const int64_t *if_src);
/* v[shared nodes] = sum over ranks through the peer-memory halo (fcvm_interface_sum is the NCCL form). */
int fcvm_p2p_interface_sum(fcvm_ctx *ctx, double *v);

/* ---- HOST-buffer drop-ins with the reference's argument lists ----------------------------- */
/* Page-locked host memory for the arrays handed to fcvm_host_* (pageable memory works too, at
 * roughly a third of the PCIe rate). */
int fcvm_host_alloc(int64_t bytes, void **out);
int fcvm_host_free(void *p);
/* update_stress_load(gp10, elNodes, nocoord, materialbyElement, sig_yield, disp_new, du, sig,
 *                    sig_update, sig_test_global, qin, Et_E, LD, pgp)      fcVM.py:2196 */
int fcvm_host_update_stress_load(fcvm_ctx *ctx, const double *sig_yield, const double *disp_new, const double *du,
                                 const double *sig, double *sig_update, double *sig_test_global, double *qin,
                                 double Et_E, int LD, uint8_t *pgp);
/* x = factor(b)                                                            fcVM.py:1130, 1401
 * recycle != 0: start from the projection onto the last two solutions with this matrix (see fcvm_pcg_solve). */
int fcvm_host_solve(fcvm_ctx *ctx, const double *b, double *x, double rtol, int max_iter, int recycle, int *iters,
                    double *relres);

/* ---- timing helpers (CUDA events on the context's stream) ---------------------------------- */
int fcvm_timer_start(fcvm_ctx *ctx);
int fcvm_timer_stop_ms(fcvm_ctx *ctx, float *ms);
/* Device time per kernel family, measured with CUDA events on the launching stream;
 * which: 0 = spmv, 1 = stress update, 2 = node gather, 3 = pcg vector kernels, 4 = assembly (whole call),
 * 5 = element stiffness, 6 = COO->SELL reduction, 7 = peer-memory exchanges of the multi-GPU PCG iteration,
 * 8 / 9 / 10 = coarse right-hand side / coarse product / expansion of the deflation level, 11 = the vector step (8-11 are also part of 3).
 * on = 0 off; 1 = every launch, synchronising after each (exact, slows the run);
 * on >= 2 = every on-th launch of a family, asynchronously (event pairs from a pool, resolved
 * when read): the timed region keeps running undisturbed.
 * fcvm_profile_get returns the summed duration and the number of launches that were timed;
 * fcvm_profile_seen the number of all launches of the family since the reset. */
int fcvm_profile_enable(fcvm_ctx *ctx, int on);
int fcvm_profile_get(fcvm_ctx *ctx, int which, double *ms, int64_t *launches);
int fcvm_profile_reset(fcvm_ctx *ctx);
int64_t fcvm_profile_seen(fcvm_ctx *ctx, int which);
int64_t fcvm_launch_count(fcvm_ctx *ctx);
/* Bytes this context has moved host->device / device->host (fcvm_h2d, fcvm_d2h, fcvm_gp_*, fcvm_host_*), counted
 * where the copies are issued. */
int fcvm_copy_bytes(fcvm_ctx *ctx, int64_t *h2d, int64_t *d2h);

/* Matrix storage facts for roofline arithmetic: stored 3x3 blocks (incl. padding) and real blocks. */
int fcvm_matrix_stats(fcvm_ctx *ctx, int64_t *blocks_stored, int64_t *blocks_real, int64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif /* FCVM_B200_H */
