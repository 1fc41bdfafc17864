// Name -> model (the reference's robot classes and the URDF+ files of robot-models/).
#include <fstream>
#include "robots.h"

namespace grbda
{
    ClusterTreeModel buildRobotByName(const std::string &name, const std::string &urdf_dir)
    {
        if (name == "tello")
            return Tello().buildClusterTreeModel();
        if (name == "tello_with_arms")
            return TelloWithArms().buildClusterTreeModel();
        const std::string a = "revolute_chain_with_rotor_", b = "revolute_pair_chain_with_rotor_";
        if (name.compare(0, a.size(), a) == 0)
            return RevoluteChainWithRotor(std::stoi(name.substr(a.size()))).buildClusterTreeModel();
        if (name.compare(0, b.size(), b) == 0)
            return RevolutePairChainWithRotor(std::stoi(name.substr(b.size()))).buildClusterTreeModel();
        const std::string c = "revolute_pair_chain_", d = "revolute_triple_chain_with_rotor_";
        if (name.compare(0, c.size(), c) == 0)
            return RevolutePairChain(std::stoi(name.substr(c.size()))).buildClusterTreeModel();
        if (name.compare(0, d.size(), d) == 0)
            return RevoluteTripleChainWithRotor(std::stoi(name.substr(d.size()))).buildClusterTreeModel();
        // URDF+ model: <urdf_dir>/<name>.urdf (mini_cheetah, mit_humanoid, four_bar, ...)
        const std::string path = urdf_dir + "/" + name + ".urdf";
        std::ifstream f(path);
        if (!f)
            throw std::runtime_error("unknown robot '" + name + "' (no builder and no " + path + ")");
        return ClusterTreeModel(path);
    }
} // namespace grbda
