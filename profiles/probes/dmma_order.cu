// Probe: in which order does mma.sync.m8n8k4.f64 (the FP64 tensor-core instruction, "DMMA") sum its four products, and
// is every step a fused multiply-add?  Compares D = A B + C on random data (wide exponent range, so that a different
// order or an unfused product shows) with the candidate orders, evaluated per output with scalar fma() on the device.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o dmma_order dmma_order.cu && ./dmma_order
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__global__ void probe(const double* A, const double* B, const double* C, double* D, int ntiles) {
  // tile t: A [8][4] row-major, B [4][8] (k-major), C / D [8][8]
  const int lane = threadIdx.x & 31;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const double a = A[t * 32 + (lane >> 2) * 4 + (lane & 3)];
    const double b = B[t * 32 + (lane & 3) * 8 + (lane >> 2)];
    double c0 = C[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3)], c1 = C[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1];
    double d0, d1;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
    D[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3)] = d0;
    D[t * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1] = d1;
  }
}

int main() {
  const int nt = 4096;
  double *hA = (double*)malloc(nt * 32 * 8), *hB = (double*)malloc(nt * 32 * 8), *hC = (double*)malloc(nt * 64 * 8),
         *hD = (double*)malloc(nt * 64 * 8);
  srand(1);
  auto rnd = [] { double m = (double)rand() / RAND_MAX * 2 - 1; int e = rand() % 41 - 20; return ldexp(m, e); };
  for (int i = 0; i < nt * 32; i++) { hA[i] = rnd(); hB[i] = rnd(); }
  for (int i = 0; i < nt * 64; i++) hC[i] = rnd();
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, nt * 32 * 8); cudaMalloc(&dB, nt * 32 * 8); cudaMalloc(&dC, nt * 64 * 8); cudaMalloc(&dD, nt * 64 * 8);
  cudaMemcpy(dA, hA, nt * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, nt * 32 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, hC, nt * 64 * 8, cudaMemcpyHostToDevice);
  probe<<<64, 32>>>(dA, dB, dC, dD, nt);
  if (cudaMemcpy(hD, dD, nt * 64 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
  long n = 0, eq_fwd = 0, eq_bwd = 0, eq_pair = 0, eq_unfused = 0, eq_prodfirst = 0;
  for (int t = 0; t < nt; t++)
    for (int m = 0; m < 8; m++)
      for (int c = 0; c < 8; c++) {
        const double* a = hA + t * 32 + m * 4;
        double b[4];
        for (int k = 0; k < 4; k++) b[k] = hB[t * 32 + k * 8 + c];
        const double c0 = hC[t * 64 + m * 8 + c], d = hD[t * 64 + m * 8 + c];
        double f = c0; for (int k = 0; k < 4; k++) f = fma(a[k], b[k], f);                 // increasing k, fused
        double r = c0; for (int k = 3; k >= 0; k--) r = fma(a[k], b[k], r);                // decreasing k
        double p = fma(a[1], b[1], a[0] * b[0]) + fma(a[3], b[3], a[2] * b[2]) + c0;       // pairwise
        double u = c0; for (int k = 0; k < 4; k++) u = u + a[k] * b[k];                    // unfused
        double q = a[0] * b[0]; for (int k = 1; k < 4; k++) q = fma(a[k], b[k], q); q += c0;   // products first, then + c
        n++; eq_fwd += (d == f); eq_bwd += (d == r); eq_pair += (d == p); eq_unfused += (d == u); eq_prodfirst += (d == q);
      }
  printf("{\"outputs\": %ld, \"equal_fma_chain_increasing_k\": %ld, \"equal_fma_chain_decreasing_k\": %ld, \"equal_pairwise\": %ld, "
         "\"equal_unfused_increasing_k\": %ld, \"equal_products_first\": %ld}\n", n, eq_fwd, eq_bwd, eq_pair, eq_unfused, eq_prodfirst);
  return 0;
}
