// Box-cluster geometry of the rigid-body-mode deflation level, shared by the set-up kernels
// and the per-iteration kernels (fcvm_deflation.cu).
#pragma once

#include "fcvm_common.cuh"

namespace fcvm {

struct Grid {
  int n[3];
  double lo[3], h[3], scale;
  const uint8_t *active;     // per box: 0 = too few free nodes for six independent modes, its columns of Z are zero
};

// Z_j: 3 x 6 = [ I | (e_k x rel)/scale ], rows of prescribed dofs zeroed
__device__ __forceinline__ void z_of(const Grid &g, int32_t cl, const double *__restrict__ xyz, const double *__restrict__ fixdof,
                                     int64_t j, double (&Z)[3][6]) {
  const int ix = cl % g.n[0], iy = (cl / g.n[0]) % g.n[1], iz = cl / (g.n[0] * g.n[1]);
  const double rx = (xyz[3 * j] - (g.lo[0] + (ix + 0.5) * g.h[0])) / g.scale;
  const double ry = (xyz[3 * j + 1] - (g.lo[1] + (iy + 0.5) * g.h[1])) / g.scale;
  const double rz = (xyz[3 * j + 2] - (g.lo[2] + (iz + 0.5) * g.h[2])) / g.scale;
  const double on = g.active[cl] ? 1.0 : 0.0;
  const double f0 = on * fixdof[3 * j], f1 = on * fixdof[3 * j + 1], f2 = on * fixdof[3 * j + 2];
  // e_x x r = (0, -rz, ry), e_y x r = (rz, 0, -rx), e_z x r = (-ry, rx, 0)
  Z[0][0] = f0; Z[0][1] = 0;  Z[0][2] = 0;  Z[0][3] = 0;        Z[0][4] = f0 * rz;  Z[0][5] = -f0 * ry;
  Z[1][0] = 0;  Z[1][1] = f1; Z[1][2] = 0;  Z[1][3] = -f1 * rz; Z[1][4] = 0;        Z[1][5] = f1 * rx;
  Z[2][0] = 0;  Z[2][1] = 0;  Z[2][2] = f2; Z[2][3] = f2 * ry;  Z[2][4] = -f2 * rx; Z[2][5] = 0;
}

__device__ __forceinline__ int rel_code(const Grid &g, int32_t from, int32_t to) {
  const int nx = g.n[0], ny = g.n[1];
  const int dx = to % nx - from % nx, dy = (to / nx) % ny - (from / nx) % ny, dz = to / (nx * ny) - from / (nx * ny);
  if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) return -1;
  return (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);
}

__device__ __forceinline__ int32_t neighbour(const Grid &g, int32_t cl, int code) {
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  const int ix = cl % nx + code % 3 - 1, iy = (cl / nx) % ny + (code / 3) % 3 - 1, iz = cl / (nx * ny) + code / 9 - 1;
  if (ix < 0 || ix >= nx || iy < 0 || iy >= ny || iz < 0 || iz >= nz) return -1;
  return ix + nx * (iy + ny * iz);
}

inline Grid grid_of(const fcvm_ctx *c) {
  Grid g;
  for (int d = 0; d < 3; d++) {
    g.n[d] = c->dn[d];
    g.lo[d] = c->dlo[d];
    g.h[d] = c->dh[d];
  }
  g.scale = c->dscale;
  g.active = c->cl_active;
  return g;
}

}  // namespace fcvm
