// Batched extension of the host ClusterTreeModel (model.h): thin C++ layer over the C ABI, the shape the
// reference-side binding of INTEGRATION.md has. Reference interface it extends:
// include/grbda/Dynamics/ClusterTreeModel.h:91-97,143-165 (setState, inverseDynamics, forwardDynamics,
// getMassMatrix, getBiasForceVector), TreeModel.h:62 (forwardKinematics).
#include <cuda_runtime.h>
#include <stdexcept>
#include "../../../include/grbda_cuda.h"
#include "schedule.h"

namespace grbda
{
    namespace
    {
        void check(grbda_status s)
        {
            if (s != GRBDA_OK)
                throw std::runtime_error(grbda_cuda_last_error_string());
        }
    } // namespace

    ::grbda_model *ClusterTreeModel::deviceModel() const
    {
        if (!device_model_)
        {
            int device = 0;
            const cudaError_t e = cudaGetDevice(&device);
            if (e != cudaSuccess)
                throw std::runtime_error(std::string("batched ClusterTreeModel needs a CUDA device: ") + cudaGetErrorString(e));
            ScheduleStorage s;
            toSchedule(*this, s);
            ::grbda_model *h = nullptr;
            check(grbda_cuda_model_create(&s.view, device, &h));
            device_model_ = std::shared_ptr<void>(h, [](void *p) { grbda_cuda_model_destroy((::grbda_model *)p); });
        }
        return (::grbda_model *)device_model_.get();
    }

    void ClusterTreeModel::inverseDynamicsBatch(const double *q, const double *yd, const double *ydd, double *tau,
                                                int64_t batch, void *stream) const
    {
        check(grbda_cuda_inverse_dynamics_f64(deviceModel(), q, yd, ydd, tau, batch, stream));
    }
    void ClusterTreeModel::forwardDynamicsBatch(const double *q, const double *yd, const double *tau, double *ydd,
                                                int64_t batch, void *stream) const
    {
        check(grbda_cuda_forward_dynamics_f64(deviceModel(), q, yd, tau, ydd, batch, stream));
    }
    void ClusterTreeModel::massMatrixBatch(const double *q, double *H, int64_t batch, void *stream) const
    {
        check(grbda_cuda_mass_matrix_f64(deviceModel(), q, H, batch, stream));
    }
    void ClusterTreeModel::forwardKinematicsBatch(const double *q, const double *yd, double *p, double *R, double *v,
                                                  int64_t batch, void *stream) const
    {
        check(grbda_cuda_forward_kinematics_f64(deviceModel(), q, yd, p, R, v, batch, stream));
    }
    void ClusterTreeModel::biasForceBatch(const double *q, const double *yd, const double *zeros, double *C,
                                          int64_t batch, void *stream) const
    {
        check(grbda_cuda_inverse_dynamics_f64(deviceModel(), q, yd, zeros, C, batch, stream));
    }
    void ClusterTreeModel::randomStatesBatch(uint64_t seed, int64_t first_index, int64_t count, double *q, double *yd,
                                             double *aux, void *stream) const
    {
        check(grbda_cuda_generate_states(deviceModel(), seed, first_index, count, q, yd, aux, nullptr, stream));
    }
} // namespace grbda
