// Batched dynamics benchmark, C++ (the "batched benchmark in Benchmarking/src" of the north star; the
// reference's own benchmarks, Benchmarking/src/pinocchioBenchmark.cpp:480-640, time one state at a time on
// the CPU and write per-model CSV lines). For every model: B random valid states generated on the device,
// inverseDynamics / forwardDynamics / getMassMatrix / forwardKinematics timed with CUDA events over
// `steps` launches after `warmup`, one CSV line per model:
//   model,bodies,clusters,nq,nv,depth,batch,id_ms,fd_ms,h_ms,fk_ms,id_evals_per_s,fd_evals_per_s
// Usage: batchedBenchmark [--batch N] [--steps K] [--warmup W] [--urdf-dir DIR] model-or-urdf ...
//   a model is a robot name ("tello_with_arms", "revolute_chain_with_rotor_24", ...) or a path to a URDF+ file
//   (e.g. the reference's Benchmarking/urdfs families); models that were not compiled ahead of time are
//   compiled at run time.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "host/robots.h"

using namespace grbda;

static void cudaOk(cudaError_t e, const char *what)
{
    if (e != cudaSuccess)
    {
        std::fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
        std::exit(2);
    }
}

static int depthOf(const ClusterTreeModel &m)
{
    std::vector<int> d(m.getNumClusters(), 1);
    int mx = 0;
    for (const ClusterTreeNode &c : m.clusters())
    {
        d[c.index_] = c.parent_index_ >= 0 ? d[c.parent_index_] + 1 : 1;
        mx = std::max(mx, d[c.index_]);
    }
    return mx;
}

template <typename F>
static double timeMs(F f, int warmup, int steps)
{
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    for (int i = 0; i < warmup; i++)
        f();
    cudaOk(cudaDeviceSynchronize(), "warm-up");
    cudaEventRecord(t0);
    for (int i = 0; i < steps; i++)
        f();
    cudaEventRecord(t1);
    cudaOk(cudaEventSynchronize(t1), "timed launches");
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return ms / steps;
}

int main(int argc, char **argv)
{
    int64_t batch = 1 << 18;
    int steps = 10, warmup = 3;
    std::string urdf_dir = "generalized_rbda_b200/robot-models";
    std::vector<std::string> models;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        if (a == "--batch" && i + 1 < argc)
            batch = std::atoll(argv[++i]);
        else if (a == "--steps" && i + 1 < argc)
            steps = std::atoi(argv[++i]);
        else if (a == "--warmup" && i + 1 < argc)
            warmup = std::atoi(argv[++i]);
        else if (a == "--urdf-dir" && i + 1 < argc)
            urdf_dir = argv[++i];
        else
            models.push_back(a);
    }
    if (models.empty())
        models = {"tello_with_arms", "mit_humanoid", "mini_cheetah"};
    std::printf("model,bodies,clusters,nq,nv,depth,batch,id_ms,fd_ms,h_ms,fk_ms,id_evals_per_s,fd_evals_per_s\n");
    for (const std::string &name : models)
    {
        try
        {
            const bool is_file = name.size() > 5 && name.substr(name.size() - 5) == ".urdf";
            const ClusterTreeModel model = is_file ? ClusterTreeModel(name) : buildRobotByName(name, urdf_dir);
            const int nq = model.getNumPositions(), nv = model.getNumDegreesOfFreedom(), nb = model.getNumBodies();
            double *q, *yd, *aux, *out, *H, *p, *R, *v;
            // mass matrix and kinematics outputs are large: time them on a quarter of the batch
            const int64_t bs = std::max<int64_t>(1, batch / 4);
            cudaOk(cudaMalloc(&q, batch * nq * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&yd, batch * nv * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&aux, batch * nv * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&out, batch * nv * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&H, bs * nv * nv * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&p, bs * nb * 3 * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&R, bs * nb * 9 * 8), "cudaMalloc");
            cudaOk(cudaMalloc(&v, bs * nb * 6 * 8), "cudaMalloc");
            model.randomStatesBatch(0x6772626461ull, 0, batch, q, yd, aux);
            const double id_ms = timeMs([&] { model.inverseDynamicsBatch(q, yd, aux, out, batch); }, warmup, steps);
            const double fd_ms = timeMs([&] { model.forwardDynamicsBatch(q, yd, aux, out, batch); }, warmup, steps);
            const double h_ms = timeMs([&] { model.massMatrixBatch(q, H, bs); }, warmup, steps) * (double)batch / bs;
            const double fk_ms = timeMs([&] { model.forwardKinematicsBatch(q, yd, p, R, v, bs); }, warmup, steps) * (double)batch / bs;
            std::printf("%s,%d,%d,%d,%d,%d,%lld,%.4f,%.4f,%.4f,%.4f,%.4g,%.4g\n", name.c_str(), nb, model.getNumClusters(), nq,
                        nv, depthOf(model), (long long)batch, id_ms, fd_ms, h_ms, fk_ms, batch / (id_ms * 1e-3),
                        batch / (fd_ms * 1e-3));
            std::fflush(stdout);
            for (double *x : {q, yd, aux, out, H, p, R, v})
                cudaFree(x);
        }
        catch (const std::exception &e)
        {
            std::fprintf(stderr, "%s: %s\n", name.c_str(), e.what());
            return 1;
        }
    }
    return 0;
}
