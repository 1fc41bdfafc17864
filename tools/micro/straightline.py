"""Micro-benchmark: how fast can sm_100a issue a STRAIGHT-LINE (never repeated) FP64 instruction stream?
Generates kernels with N independent-chain DFMAs (8 chains), no memory traffic, and times them at the
launch shape of the dynamics kernels (128 threads, 255 registers, 2 CTAs/SM). Compared with the same
work in a loop (instruction-cache resident). Usage: python tools/micro/straightline.py"""
import os, subprocess, sys, tempfile
import torch
from torch.utils.cpp_extension import load_inline  # noqa: F401  (not used: plain nvcc + ctypes)
import ctypes

N_LIST = [4000, 16000]
MIX = int(os.environ.get("MIX", "0"))
CHAINS = int(os.environ.get("CHAINS", "8"))  # independent dependency chains per thread (ILP)
src = ['#include <cuda_runtime.h>\n#include <cstdio>\n']
for n in N_LIST:
    body = []
    for i in range(n):
        k = i % CHAINS
        body.append(f"x{k} = fma(x{k}, a, b);")
        if MIX:
            body.append(f"j{k} = j{k} * 3 + {i};")  # one integer IMAD per DFMA
    src.append(f'''
extern "C" __global__ void __launch_bounds__(128, 2) straight{n}(double* out, double a, double b) {{
  double x0=threadIdx.x, x1=x0+1, x2=x0+2, x3=x0+3, x4=x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
  int j0=threadIdx.x, j1=j0+1, j2=j0+2, j3=j0+3, j4=j0+4, j5=j0+5, j6=j0+6, j7=j0+7;
  {" ".join(body)}
  double s=x0+x1+x2+x3+x4+x5+x6+x7+(double)(j0^j1^j2^j3^j4^j5^j6^j7); if (s==-1.2345) out[0]=s;
}}
extern "C" __global__ void __launch_bounds__(128, 2) looped{n}(double* out, double a, double b) {{
  double x0=threadIdx.x, x1=x0+1, x2=x0+2, x3=x0+3, x4=x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
  int j0=threadIdx.x, j1=j0+1, j2=j0+2, j3=j0+3, j4=j0+4, j5=j0+5, j6=j0+6, j7=j0+7;
  for (int i = 0; i < {n // 64}; i++) {{
    #pragma unroll
    for (int k = 0; k < 8; k++) {{ x0=fma(x0,a,b); x1=fma(x1,a,b); x2=fma(x2,a,b); x3=fma(x3,a,b); x4=fma(x4,a,b); x5=fma(x5,a,b); x6=fma(x6,a,b); x7=fma(x7,a,b);
      if ({MIX}) {{ j0=j0*3+i; j1=j1*3+i; j2=j2*3+i; j3=j3*3+i; j4=j4*3+i; j5=j5*3+i; j6=j6*3+i; j7=j7*3+i; }} }}
  }}
  double s=x0+x1+x2+x3+x4+x5+x6+x7+(double)(j0^j1^j2^j3^j4^j5^j6^j7); if (s==-1.2345) out[0]=s;
}}
''')
src.append('''
extern "C" float run(int which, int n, int grid) {
  double* out; cudaMalloc(&out, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    switch (n) {
''')
for n in N_LIST:
    src.append(f'      case {n}: if (which) looped{n}<<<grid,128>>>(out,1.0000001,1e-7); else straight{n}<<<grid,128>>>(out,1.0000001,1e-7); break;\n')
src.append('''    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaFree(out); return best;
}
''')
d = tempfile.mkdtemp()
cu = os.path.join(d, "sl.cu"); so = os.path.join(d, "sl.so")
open(cu, "w").write("".join(src))
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", so, cu])
lib = ctypes.CDLL(so); lib.run.restype = ctypes.c_float
grid = 8192
for n in N_LIST:
    for which, name in ((0, "straight"), (1, "looped")):
        ms = lib.run(which, n, grid)
        warps = grid * 4
        ipc = n * warps / (ms * 1e-3 * 1.965e9 * 148 * 4)
        print(f"MIX={MIX} CHAINS={CHAINS if which == 0 else 8} {name:9s} n={n:6d}  {ms:8.3f} ms   DFMA per SMSP-cycle = {ipc:.3f}  total instr per SMSP-cycle = {ipc*(1+MIX):.3f}  ({2*n*grid*128/ms/1e9:.1f} TFLOP/s)")
