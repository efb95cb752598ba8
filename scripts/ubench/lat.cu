// Latency microbenchmarks (single warp): dependent chains of the ops the DP kernel is made of.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(double* out, long long* cyc, double seed, int sel) {
  double x = seed + threadIdx.x, y = seed * 0.5, z = 1.0000001;
  unsigned long long kx = (unsigned long long)threadIdx.x * 7919u + sel, ky = 12345;
  int ix = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = __dadd_rn(x, y);                               // DADD chain
    if (OP == 1) { if (x >= y) { y = x; } x = __dadd_rn(y, z); }    // DSETP+FSEL+DADD
    if (OP == 2) x = __shfl_xor_sync(0xffffffffu, x, 1);            // 2x SHFL chain
    if (OP == 3) { kx = (kx > ky) ? kx : ky + i; }                  // 64-bit int compare/select
    if (OP == 4) { float f = (float)x; f = __fadd_rn(f, 1.5f); x = (double)f; }  // F2F chain
    if (OP == 5) ix = __shfl_xor_sync(0xffffffffu, ix, 1) + 1;      // SHFL int chain
    if (OP == 6) x = fmax(x, y) + 0.0;                              // DMNMX?
    if (OP == 7) { x = __dmul_rn(x, z); }                           // DMUL
    if (OP == 8) { ix = ix * 3 + i; }                               // IMAD chain
    if (OP == 9) { x = (x > y) ? x : z; y = __dadd_rn(y, z);}       // DSETP+FSEL only chain on x, dadd indep
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y + (double)kx + ix;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DADD dep chain", "DSETP+FSEL+DADD", "SHFL f64 (2x) chain", "u64 cmp+sel", "F2F f64->f32 FADD f32->f64",
                         "SHFL int + IADD", "fmax + DADD", "DMUL chain", "IMAD chain", "DSETP+FSEL chain"};
  for (int op = 0; op < 10; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<1, 32>>>(out, cyc, 1.0, rep); break; case 1: k<1><<<1, 32>>>(out, cyc, 1.0, rep); break;
        case 2: k<2><<<1, 32>>>(out, cyc, 1.0, rep); break; case 3: k<3><<<1, 32>>>(out, cyc, 1.0, rep); break;
        case 4: k<4><<<1, 32>>>(out, cyc, 1.0, rep); break; case 5: k<5><<<1, 32>>>(out, cyc, 1.0, rep); break;
        case 6: k<6><<<1, 32>>>(out, cyc, 1.0, rep); break; case 7: k<7><<<1, 32>>>(out, cyc, 1.0, rep); break;
        case 8: k<8><<<1, 32>>>(out, cyc, 1.0, rep); break; case 9: k<9><<<1, 32>>>(out, cyc, 1.0, rep); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-32s %.1f cycles/iter\n", names[op], (double)c / N);
  }
  return 0;
}
