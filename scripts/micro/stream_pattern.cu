// Micro-benchmark: achievable HBM rate of the stress update's access pattern (28 SoA read streams,
// 52 SoA write streams, 8 B per lane) without any arithmetic, for several block mappings.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// mapping A: block = 32 elements x 4 gp warps (the kernel's mapping)
__global__ void __launch_bounds__(128) k_a(int64_t ne, const double *__restrict__ so, const double *__restrict__ sy,
                                           double *sn, double *st, uint8_t *pg) {
  const int lane = threadIdx.x & 31, gp = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  if (e >= ne) return;
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; c++) s[c] = __ldcs(&so[((int64_t)c * 4 + gp) * ne + e]);
  const double y = __ldcs(&sy[(int64_t)gp * ne + e]);
#pragma unroll
  for (int c = 0; c < 6; c++) {
    __stcs(&st[((int64_t)c * 4 + gp) * ne + e], s[c] + y);
    __stcs(&sn[((int64_t)c * 4 + gp) * ne + e], s[c] - y);
  }
  pg[(int64_t)gp * ne + e] = y > s[0];
}
// mapping B: one thread per element, all four gp
__global__ void __launch_bounds__(128) k_b(int64_t ne, const double *__restrict__ so, const double *__restrict__ sy,
                                           double *sn, double *st, uint8_t *pg) {
  const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= ne) return;
#pragma unroll
  for (int gp = 0; gp < 4; gp++) {
    double s[6];
#pragma unroll
    for (int c = 0; c < 6; c++) s[c] = __ldcs(&so[((int64_t)c * 4 + gp) * ne + e]);
    const double y = __ldcs(&sy[(int64_t)gp * ne + e]);
#pragma unroll
    for (int c = 0; c < 6; c++) {
      __stcs(&st[((int64_t)c * 4 + gp) * ne + e], s[c] + y);
      __stcs(&sn[((int64_t)c * 4 + gp) * ne + e], s[c] - y);
    }
    pg[(int64_t)gp * ne + e] = y > s[0];
  }
}
// mapping C: element-blocked AoSoA: [e/32][28][32] in, [e/32][48][32] out -> one contiguous run per block
__global__ void __launch_bounds__(128) k_c(int64_t ne, const double *__restrict__ in, double *out, uint8_t *pg) {
  const int lane = threadIdx.x & 31, gp = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const double *pi = in + b * 28 * 32;
  double *po = out + b * 48 * 32;
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; c++) s[c] = __ldcs(&pi[(c * 4 + gp) * 32 + lane]);
  const double y = __ldcs(&pi[(24 + gp) * 32 + lane]);
#pragma unroll
  for (int c = 0; c < 6; c++) {
    __stcs(&po[(c * 4 + gp) * 32 + lane], s[c] + y);
    __stcs(&po[(24 + c * 4 + gp) * 32 + lane], s[c] - y);
  }
  pg[b * 128 + gp * 32 + lane] = y > s[0];
}

int main(int argc, char **argv) {
  int64_t ne = argc > 1 ? atoll(argv[1]) : 998250;
  int64_t nep = (ne + 31) / 32 * 32;
  double *so, *sy, *sn, *st, *in, *out; uint8_t *pg;
  CK(cudaMalloc(&so, 24 * nep * 8)); CK(cudaMalloc(&sy, 4 * nep * 8)); CK(cudaMalloc(&sn, 24 * nep * 8));
  CK(cudaMalloc(&st, 24 * nep * 8)); CK(cudaMalloc(&pg, 4 * nep)); CK(cudaMalloc(&in, 28 * nep * 8));
  CK(cudaMalloc(&out, 48 * nep * 8));
  CK(cudaMemset(so, 0, 24 * nep * 8)); CK(cudaMemset(sy, 0, 4 * nep * 8)); CK(cudaMemset(in, 0, 28 * nep * 8));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const double bytes = (double)ne * (28 * 8 + 48 * 8 + 4);
  for (int variant = 0; variant < 5; variant++) {
    float best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
      cudaEventRecord(a);
      if (variant == 0) k_a<<<(unsigned)((ne + 31) / 32), 128>>>(ne, so, sy, sn, st, pg);
      if (variant == 1) k_a<<<(unsigned)((nep + 31) / 32), 128>>>(nep, so, sy, sn, st, pg);   // aligned stride
      if (variant == 2) k_b<<<(unsigned)((ne + 127) / 128), 128>>>(ne, so, sy, sn, st, pg);
      if (variant == 3) k_b<<<(unsigned)((nep + 127) / 128), 128>>>(nep, so, sy, sn, st, pg);
      if (variant == 4) k_c<<<(unsigned)(nep / 32), 128>>>(nep, in, out, pg);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (rep > 0 && ms < best) best = ms;
    }
    const char *names[5] = {"A 32el x 4gp warps, stride ne", "A, stride padded to 32", "B thread/element, stride ne",
                            "B, stride padded", "C blocked AoSoA"};
    printf("%-34s %.3f ms  %.0f GB/s\n", names[variant], best, bytes / best / 1e6);
  }
  CK(cudaGetLastError());
  return 0;
}
