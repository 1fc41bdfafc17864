// Registry of ahead-of-time compiled per-model kernel sets. Every generated translation unit
// (csrc/generated/<model>.cu, written by the model compiler tool `grbda_modelc`) registers one
// ModelKernels record keyed by the hash of the model description it was generated from; the C ABI
// looks kernels up by that hash, so a model built at run time from a URDF or a schedule finds its
// kernels iff it is exactly the model that was compiled.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../kernels/batched_kernel.cuh"

namespace grbda_runtime
{
    using grbda_kernels::LaunchArgs;
    typedef cudaError_t (*LaunchFn)(const LaunchArgs &);

    struct GenArgs
    {
        uint64_t seed;
        int64_t first_index, count;
        double *q, *yd, *aux;
        int32_t *flags;
        cudaStream_t stream;
    };
    typedef cudaError_t (*GenFn)(const GenArgs &);
    struct StepArgs
    {
        const double *q, *yd, *ydd;
        double dt;
        int64_t count;
        double *q_out, *yd_out;
        int32_t *flags;
        cudaStream_t stream;
    };
    typedef cudaError_t (*StepFn)(const StepArgs &);

    enum
    {
        KERNEL_VARIANTS = 4 // launch-shape variants of each kernel (see generated files)
    };

    struct AlgoKernels
    {
        LaunchFn f64[KERNEL_VARIANTS];
        LaunchFn f32[KERNEL_VARIANTS];
        int n_in[3], n_out[3];
        int64_t counts[8]; // nodes, add, mul, div, sqrt, sin, cos, fusable
    };

    struct ModelKernels
    {
        uint64_t hash;
        const char *name;
        int nq, nv, nb, nc;
        AlgoKernels algo[7]; // id, fd, fk, h, phi, gfa, gfs (compiler::ALGO_COUNT)
        GenFn generate;
        StepFn integrate;
    };

    // Records are assembled from the per-algorithm translation units at load time.
    ModelKernels *modelRecord(uint64_t hash, const char *name, int nq, int nv, int nb, int nc);
    const ModelKernels *findModelKernels(uint64_t hash);
    int numRegisteredModels();
    const ModelKernels *registeredModel(int i);

    struct AlgoRegistrar
    {
        AlgoRegistrar(uint64_t hash, const char *name, int nq, int nv, int nb, int nc, int algo,
                      const AlgoKernels *k)
        {
            modelRecord(hash, name, nq, nv, nb, nc)->algo[algo] = *k;
        }
    };
    struct GenRegistrar
    {
        GenRegistrar(uint64_t hash, const char *name, int nq, int nv, int nb, int nc, GenFn fn, StepFn step = nullptr)
        {
            ModelKernels *k = modelRecord(hash, name, nq, nv, nb, nc);
            k->generate = fn;
            k->integrate = step;
        }
    };
} // namespace grbda_runtime
