"""Micro-benchmark: how fast can a B200 take the OUTPUT of a row-per-thread kernel (forward kinematics: 666 doubles per
state, one state per thread), as a function of the store pattern and the occupancy?
  A  every lane stores 32-byte sectors of ITS OWN row (the shipped form: one warp instruction = 32 sectors in 32 lines)
  B  the four lanes of a quad store the four sectors of one 128-byte line, quads walk over 4 states (one warp
     instruction = 8 whole lines)
  C  the warp stores 1 KB of one row (fully coalesced reference)
Rows of 672 doubles (5 376 bytes, line aligned), 2^18 rows. Usage: python tools/micro/store_patterns.py"""
import ctypes
import os
import subprocess
import tempfile

SRC = r'''
#include <cuda_runtime.h>
#include <cstdio>
#define N 672
__device__ __forceinline__ void st4(double* p, double a) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %1, %1, %1};" :: "l"(p), "d"(a) : "memory");
}
template <int PATTERN>
__global__ void __launch_bounds__(128) k(double* out, long rows, double v) {
  extern __shared__ double pad[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long base = ((long)blockIdx.x * 4 + warp) * 32;   // first row of the warp
  if (base >= rows) return;
  if (PATTERN == 0) {
    double* row = out + (base + lane) * N;
    #pragma unroll 8
    for (int s = 0; s < N / 4; s++) st4(row + 4 * s, v + s);
  } else if (PATTERN == 1) {
    #pragma unroll 1
    for (int k4 = 0; k4 < 4; k4++) {
      double* row = out + (base + 4 * (lane >> 2) + k4) * N + 4 * (lane & 3);
      #pragma unroll 6
      for (int l = 0; l < N / 16; l++) st4(row + 16 * l, v + l);
    }
  } else {
    #pragma unroll 1
    for (int r = 0; r < 32; r++) {
      double* row = out + (base + r) * N + 4 * lane;
      #pragma unroll
      for (int c = 0; c < N / 128; c++) st4(row + 128 * c, v + c);
      if (lane < (N % 128) / 4) st4(row + 128 * (N / 128), v);
    }
  }
}
extern "C" float run(int pattern, long rows, int smem) {
  double* out; cudaMalloc(&out, rows * N * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto launch = [&](void) {
    const int grid = (int)((rows + 127) / 128);
    if (pattern == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<0><<<grid, 128, smem>>>(out, rows, 1.0); }
    if (pattern == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<1><<<grid, 128, smem>>>(out, rows, 1.0); }
    if (pattern == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<2><<<grid, 128, smem>>>(out, rows, 1.0); }
  };
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  cudaFree(out); return best;
}
'''
d = tempfile.mkdtemp()
cu, so = os.path.join(d, "sp.cu"), os.path.join(d, "sp.so")
open(cu, "w").write(SRC)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", so, cu])
lib = ctypes.CDLL(so)
lib.run.restype = ctypes.c_float
lib.run.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_int]
rows = 1 << 18
for smem, occ in ((112 * 1024, "2 CTAs/SM (8 warps)"), (56 * 1024, "4 CTAs/SM (16 warps)"), (0, "16 CTAs/SM (64 warps)")):
    for p, name in ((0, "A own-row sectors"), (1, "B quad lines"), (2, "C coalesced")):
        ms = lib.run(p, rows, smem)
        print("%-22s %-18s %7.3f ms  %6.2f TB/s" % (occ, name, ms, rows * 672 * 8 / ms / 1e9))
