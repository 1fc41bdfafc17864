// Compiled C++ test of the batched extension of the host ClusterTreeModel (csrc/host/model.h):
// models are built through the reference's construction API (registerBody /
// appendRegisteredBodiesAsCluster<ClusterJoints::...>, Robot::buildClusterTreeModel, ClusterTreeModel(urdf)),
// evaluated with inverseDynamicsBatch / forwardDynamicsBatch / massMatrixBatch / forwardKinematicsBatch on
// the GPU and compared with the CPU oracle (oracle/liboracle.so, test infrastructure) at 1e-10.
// Reference interface: include/grbda/Dynamics/ClusterTreeModel.h:23-97,158-165; the per-state loop this
// replaces is the body of UnitTests/testRigidBodyDynamicsAlgos.cpp:113-239.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "host/robots.h"

extern "C"
{
    void *oracle_model_create(const char *name);
    const char *oracle_last_error();
    int oracle_inverse_dynamics(void *h, const double *q, const double *yd, const double *ydd, double *tau, int64_t batch, int threads);
    int oracle_forward_dynamics(void *h, const double *q, const double *yd, const double *tau, double *ydd, int64_t batch, int threads);
    int oracle_mass_matrix(void *h, const double *q, double *H, int64_t batch, int threads);
    int oracle_forward_kinematics(void *h, const double *q, const double *yd, double *p, double *R, double *v, int64_t batch, int threads);
}

using namespace grbda;

#define CUDA_OK(x)                                                                      \
    do                                                                                  \
    {                                                                                   \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess)                                                          \
        {                                                                               \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));               \
            std::exit(2);                                                               \
        }                                                                               \
    } while (0)

static double relRows(const std::vector<double> &a, const std::vector<double> &b, int64_t rows, int n)
{
    double worst = 0;
    for (int64_t r = 0; r < rows; r++)
    {
        double d = 0, m = 1e-3;
        for (int i = 0; i < n; i++)
        {
            d = std::fmax(d, std::fabs(a[r * n + i] - b[r * n + i]));
            m = std::fmax(m, std::fabs(b[r * n + i]));
        }
        worst = std::fmax(worst, d / m);
    }
    return worst;
}

struct DeviceArray
{
    double *p = nullptr;
    size_t n = 0;
    explicit DeviceArray(size_t n_) : n(n_) { CUDA_OK(cudaMalloc(&p, std::max<size_t>(1, n) * sizeof(double))); }
    ~DeviceArray() { cudaFree(p); }
    std::vector<double> host() const
    {
        std::vector<double> h(n);
        CUDA_OK(cudaMemcpy(h.data(), p, n * sizeof(double), cudaMemcpyDeviceToHost));
        return h;
    }
};

// the same model twice: the product's (GPU) and the oracle's (CPU) hand-coded builder
static int checkModel(const char *label, const ClusterTreeModel &model, const char *oracle_name, int64_t B)
{
    void *o = oracle_model_create(oracle_name);
    if (!o)
    {
        std::fprintf(stderr, "oracle: %s\n", oracle_last_error());
        return 1;
    }
    const int nq = model.getNumPositions(), nv = model.getNumDegreesOfFreedom(), nb = model.getNumBodies();
    DeviceArray q(B * nq), yd(B * nv), aux(B * nv), tau(B * nv), ydd(B * nv), H(B * nv * nv), p(B * 3 * nb), R(B * 9 * nb), v(B * 6 * nb);
    model.randomStatesBatch(0x6772626461ull, 0, B, q.p, yd.p, aux.p);
    model.inverseDynamicsBatch(q.p, yd.p, aux.p, tau.p, B);
    model.forwardDynamicsBatch(q.p, yd.p, aux.p, ydd.p, B);
    model.massMatrixBatch(q.p, H.p, B);
    model.forwardKinematicsBatch(q.p, yd.p, p.p, R.p, v.p, B);
    CUDA_OK(cudaDeviceSynchronize());
    const std::vector<double> hq = q.host(), hyd = yd.host(), haux = aux.host();
    std::vector<double> otau(B * nv), oydd(B * nv), oH(B * nv * nv), op(B * 3 * nb), oR(B * 9 * nb), ov(B * 6 * nb);
    if (oracle_inverse_dynamics(o, hq.data(), hyd.data(), haux.data(), otau.data(), B, 0) ||
        oracle_forward_dynamics(o, hq.data(), hyd.data(), haux.data(), oydd.data(), B, 0) ||
        oracle_mass_matrix(o, hq.data(), oH.data(), B, 0) ||
        oracle_forward_kinematics(o, hq.data(), hyd.data(), op.data(), oR.data(), ov.data(), B, 0))
    {
        std::fprintf(stderr, "oracle: %s\n", oracle_last_error());
        return 1;
    }
    const double e_id = relRows(tau.host(), otau, B, nv), e_fd = relRows(ydd.host(), oydd, B, nv),
                 e_h = relRows(H.host(), oH, B, nv * nv), e_p = relRows(p.host(), op, B, 3 * nb),
                 e_r = relRows(R.host(), oR, B, 9 * nb), e_v = relRows(v.host(), ov, B, 6 * nb);
    std::printf("%-28s nq %2d nv %2d bodies %2d  rel err  ID %.1e  FD %.1e  H %.1e  FK %.1e %.1e %.1e\n", label, nq, nv,
                nb, e_id, e_fd, e_h, e_p, e_r, e_v);
    const double tol = 1e-10;
    return (e_id < tol && e_fd < tol && e_h < tol && e_p < tol && e_r < tol && e_v < tol) ? 0 : 1;
}

int main(int argc, char **argv)
{
    const std::string urdf_dir = argc > 1 ? argv[1] : "generalized_rbda_b200/robot-models";
    int failures = 0;
    const int64_t B = 777;

    // 1. a robot class of the reference (Robot::buildClusterTreeModel), implicit clusters included
    failures += checkModel("TelloWithArms", TelloWithArms().buildClusterTreeModel(), "tello_with_arms", B);

    // 2. a model assembled by hand through registerBody / appendRegisteredBodiesAsCluster, the way
    //    src/Robots/SerialChains/RevoluteChainWithRotor.cpp:45-109 does (three links, not a compiled size:
    //    its kernels come from the run-time compiler)
    {
        using namespace ClusterJoints;
        ClusterTreeModel model;
        model.setGravity({9.81, 0., 0.});
        const Mat3 I3 = ori::identity3();
        const SpatialInertia link(1., {0.5, 0., 0.}, {0, 0, 0, 0, 0, 0, 0, 0, 1.});
        const SpatialInertia rotor(0., {0., 0., 0.}, {0, 0, 0, 0, 0, 0, 0, 0, 1e-4});
        std::string parent = "ground";
        for (int i = 0; i < 3; i++)
        {
            const spatial::Transform X(I3, {i == 0 ? 0. : 1., 0., 0.});
            const std::string k = std::to_string(i);
            Body l = model.registerBody("link-" + k, link, parent, X);
            Body r = model.registerBody("rotor-" + k, rotor, parent, X);
            GearedTransmissionModule module{l, r, "link-joint-" + k, "rotor-joint-" + k, ori::CoordinateAxis::Z,
                                            ori::CoordinateAxis::Z, 6.};
            model.appendRegisteredBodiesAsCluster<RevoluteWithRotor>("cluster-" + k, module);
            parent = "link-" + k;
        }
        failures += checkModel("hand-built rotor chain (3)", model, "revolute_chain_with_rotor_3", B);
        // modifying the model drops the device-side copy: the next call sees the new gravity
        DeviceArray q(B * 3), yd(B * 3), z(B * 3), c0(B * 3), c1(B * 3);
        model.randomStatesBatch(1, 0, B, q.p, yd.p, z.p);
        CUDA_OK(cudaMemset(z.p, 0, B * 3 * sizeof(double)));
        model.biasForceBatch(q.p, yd.p, z.p, c0.p, B);
        model.setGravity({0., 0., 0.});
        model.biasForceBatch(q.p, yd.p, z.p, c1.p, B);
        CUDA_OK(cudaDeviceSynchronize());
        if (relRows(c0.host(), c1.host(), B, 3) < 1e-3)
        {
            std::printf("setGravity did not reach the device-side model\n");
            failures++;
        }
    }

    // 3. ClusterTreeModel(urdf_file)
    failures += checkModel("mini_cheetah.urdf", ClusterTreeModel(urdf_dir + "/mini_cheetah.urdf"), "mini_cheetah", B);

    // 4. errors surface as std::runtime_error, as in the reference
    try
    {
        const ClusterTreeModel chain = RevoluteChainWithRotor(2).buildClusterTreeModel();
        chain.massMatrixBatch(nullptr, nullptr, 1);
        std::printf("expected an exception for a null pointer\n");
        failures++;
    }
    catch (const std::runtime_error &)
    {
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
