#include "model.h"
namespace grbda
{
    void ClusterTreeModel::buildModelFromURDF(const std::string &urdf_filename)
    {
        throw std::runtime_error("URDF parsing not available yet: " + urdf_filename);
    }
}
