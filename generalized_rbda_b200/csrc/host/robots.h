// Robot builders with the reference's interface (reference: include/grbda/Robots/Robot.h:10-54:
// `virtual ClusterTreeModel buildClusterTreeModel() const`).
#pragma once
#include "model.h"

namespace grbda
{
    class Robot
    {
    public:
        virtual ~Robot() {}
        virtual ClusterTreeModel buildClusterTreeModel() const = 0;
    };

    // reference: include/grbda/Robots/Tello.hpp, src/Robots/Tello.cpp
    class Tello : public Robot
    {
    public:
        ClusterTreeModel buildClusterTreeModel() const override;
    };

    // reference: include/grbda/Robots/TelloWithArms.hpp, src/Robots/TelloWithArms.cpp
    class TelloWithArms : public Tello
    {
    public:
        ClusterTreeModel buildClusterTreeModel() const override;
    };

    // reference: src/Robots/SerialChains/RevoluteChainWithRotor.cpp:45-109 (uniform model)
    class RevoluteChainWithRotor : public Robot
    {
    public:
        explicit RevoluteChainWithRotor(int N) : N_(N) {}
        ClusterTreeModel buildClusterTreeModel() const override;

    private:
        int N_;
    };

    // reference: src/Robots/SerialChains/RevolutePairChainWithRotor.cpp:62-128 (uniform model)
    class RevolutePairChainWithRotor : public Robot
    {
    public:
        explicit RevolutePairChainWithRotor(int N) : N_(N) {}
        ClusterTreeModel buildClusterTreeModel() const override;

    private:
        int N_;
    };

    // reference: src/Robots/SerialChains/RevolutePairChain.cpp:49-109 (uniform model; RevolutePair clusters, no rotors)
    class RevolutePairChain : public Robot
    {
    public:
        explicit RevolutePairChain(int N) : N_(N) {}
        ClusterTreeModel buildClusterTreeModel() const override;

    private:
        int N_;
    };

    // reference: src/Robots/SerialChains/RevoluteTripleChainWithRotor.cpp:8-83. The reference only builds this
    // chain with rand() parameters (buildUniformClusterTreeModel throws); here the same recipe is driven by a
    // seeded generator (random coordinate rotations and axes, random link / rotor inertias, gear and belt
    // ratios in 1..5), so that the model is reproducible.
    class RevoluteTripleChainWithRotor : public Robot
    {
    public:
        explicit RevoluteTripleChainWithRotor(int N, uint64_t seed = 1) : N_(N), seed_(seed) {}
        ClusterTreeModel buildClusterTreeModel() const override;

    private:
        int N_;
        uint64_t seed_;
    };

    // URDF+ based robots (host/urdf.cpp); `urdf_dir` is the directory holding the URDF files
    ClusterTreeModel buildRobotByName(const std::string &name, const std::string &urdf_dir);

} // namespace grbda
