// ORACLE — TEST INFRASTRUCTURE ONLY (see scalar.h).
//
// Robot builders restated from the reference's hand-coded models:
//   src/Robots/{Tello,TelloWithArms,MiniCheetah,MIT_Humanoid}.cpp + include/grbda/Robots/*.hpp
//   src/Robots/SerialChains/{RevoluteChainWithRotor,RevolutePairChainWithRotor}.cpp
// Contact points / end effectors are not part of the dynamics hot path and are omitted.
#pragma once
#include "model.h"

namespace grbda_oracle
{
    template <typename T>
    using JointVec = std::vector<std::shared_ptr<SingleJoint<T>>>;

    template <typename T>
    Mat<T> V3(double x, double y, double z) { return vec3<T>(T(x), T(y), T(z)); }

    template <typename T>
    void appendRevoluteWithRotor(ClusterTreeModel<T> &model, const std::string &cluster_name,
                                 const std::string &link_name, const std::string &rotor_name,
                                 const std::string &parent_name, const Mat<T> &link_inertia,
                                 const Mat<T> &rotor_inertia, const Transform<T> &link_Xtree,
                                 const Transform<T> &rotor_Xtree, Axis link_axis, Axis rotor_axis,
                                 double gear_ratio)
    {
        Body<T> link = model.registerBody(link_name, link_inertia, parent_name, link_Xtree);
        Body<T> rotor = model.registerBody(rotor_name, rotor_inertia, parent_name, rotor_Xtree);
        GearedTransmissionModule<T> module{link, rotor, link_axis, rotor_axis, T(gear_ratio)};
        model.appendRegisteredBodiesAsCluster(
            cluster_name, std::make_shared<RevoluteWithRotorCluster<T>>(module));
    }

    ////////////////////////////////////////////////////////////////////////////////////////////
    // Tello (src/Robots/Tello.cpp:6-277, include/grbda/Robots/Tello.hpp)
    ////////////////////////////////////////////////////////////////////////////////////////////
    // Hip differential phi (Tello.cpp:139-154). `3021 / 160000` is an integer division (= 0) in
    // the reference and is reproduced as 0.
    template <typename S>
    std::vector<S> telloHipDifferentialPhi(const std::vector<S> &q)
    {
        const double N = 6.0;
        S ql_1 = q[0], ql_2 = q[1], y_1 = q[2] / N, y_2 = q[3] / N;
        std::vector<S> out(2);
        out[0] = (57. * sin(y_1)) / 2500. - (49. * cos(ql_1)) / 5000. - (399. * sin(ql_1)) / 20000. -
                 (8. * cos(y_1) * cos(ql_2)) / 625. - (57. * cos(ql_1) * sin(ql_2)) / 2500. -
                 (7. * sin(y_1) * sin(ql_1)) / 625. + (7. * sin(ql_1) * sin(ql_2)) / 625. -
                 (8. * cos(ql_1) * sin(y_1) * sin(ql_2)) / 625. + (double)(3021 / 160000);
        out[1] = (57. * sin(y_2)) / 2500. - (49. * cos(ql_1)) / 5000. + (399. * sin(ql_1)) / 20000. -
                 (8. * cos(y_2) * cos(ql_2)) / 625. - (57. * cos(ql_1) * sin(ql_2)) / 2500. +
                 (7. * sin(y_2) * sin(ql_1)) / 625. - (7. * sin(ql_1) * sin(ql_2)) / 625. -
                 (8. * cos(ql_1) * sin(y_2) * sin(ql_2)) / 625. + (double)(3021 / 160000);
        return out;
    }

    // Knee-ankle differential phi (Tello.cpp:237-252); `163349 / 6250000` = 0 (integer division).
    template <typename S>
    std::vector<S> telloKneeAnkleDifferentialPhi(const std::vector<S> &q)
    {
        const double N = 6.0;
        S ql_1 = q[0], ql_2 = q[1], y_1 = q[2] / N, y_2 = q[3] / N;
        std::vector<S> out(2);
        out[0] = (21. * cos(y_1 / 2. - y_2 / 2. + (1979 * 3.1415) / 4500)) / 6250. -
                 (13. * cos(y_1 / 2. - y_2 / 2. + (493 * 3.1415) / 1500)) / 625. -
                 S((273 * std::cos(3.1415 / 9)) / 12500) -
                 (7. * sin(y_1 / 2. - y_2 / 2. + ql_2 + (231 * 3.1415) / 500)) / 2500. +
                 (91. * sin(ql_2 + (2 * 3.1415) / 15)) / 5000. -
                 (147. * sin(ql_2 + 3.1415 / 45)) / 50000. + (double)(163349 / 6250000);
        out[1] = ql_1 - y_2 / 2. - y_1 / 2.;
        return out;
    }

    template <typename T>
    struct TelloParams
    {
        // Tello.hpp
        Mat<T> R_down = mat3<T>({1., 0., 0., 0., -1., 0., 0., 0., -1.});
        Mat<T> R_left = mat3<T>({-1., 0., 0., 0., 0., 1., 0., 1., 0.});
        Mat<T> R_right = mat3<T>({1., 0., 0., 0., 0., 1., 0., -1., 0.});
        Mat<T> I3 = Mat<T>::Identity(3);
        double grav = -9.81;
        double gear_ratio = 6.0;
    };

    template <typename T>
    ClusterTreeModel<T> buildTello()
    {
        using S = Taylor2<T>;
        TelloParams<T> P;
        ClusterTreeModel<T> model;
        model.setGravity(0., 0., P.grav);

        const std::string base = "torso";
        Mat<T> torso_inertia = spatialInertia<T>(
            T(2.3008), V3<T>(0.0073, -0.0013, -0.0023),
            mat3<T>({0.0366, 0., -0.0006, 0., 0.0142, -0.0002, -0.0006, -0.0002, 0.0291}));
        {
            Body<T> torso = model.registerBody(base, torso_inertia, "ground", Transform<T>());
            model.appendRegisteredBodiesAsCluster(base, std::make_shared<FreeCluster<T>>(torso, true));
        }

        Mat<T> hip_clamp_inertia = spatialInertia<T>(
            T(1.3289), V3<T>(-0.0010, 0., -0.0069),
            mat3<T>({0.0032, 0., 0.0001, 0., 0.0033, 0., 0.0001, 0., 0.0027}));
        Mat<T> gimbal_inertia = spatialInertia<T>(
            T(0.4433), V3<T>(-0.0027, 0., 0.0258),
            mat3<T>({0.0018, 0., 0., 0., 0.0017, 0., 0., 0., 0.0015}));
        Mat<T> thigh_inertia = spatialInertia<T>(
            T(1.5424), V3<T>(0.003, -0.0001, -0.0323),
            mat3<T>({0.0103, 0., -0.0005, 0., 0.0097, 0., -0.0005, 0., 0.0027}));
        Mat<T> shin_inertia = spatialInertia<T>(
            T(0.3072), V3<T>(0.0047, -0.0003, -0.1043),
            mat3<T>({0.0054, -0., -0.0002, -0., 0.0054, 0., -0.0002, 0., 0.0001}));
        Mat<T> foot_inertia = spatialInertia<T>(
            T(0.1025), V3<T>(0.0042, -0., -0.0251),
            mat3<T>({0.094e-3, -0., -0.0038e-3, -0., 0.1773e-3, 0., -0.0038e-3, 0., 0.0901e-3}));
        Mat<T> rotor_inertia = spatialInertia<T>(
            T(0.07), V3<T>(0., 0., 0.),
            mat3<T>({2.5984e-5, 0., 0., 0., 2.5984e-5, 0., 0., 0., 5.1512e-5}));

        const char *sides[2] = {"left", "right"};
        for (int i = 0; i < 2; i++)
        {
            const std::string side = sides[i];
            const double sy = i == 0 ? 1.0 : -1.0;

            // Hip clamp cluster (Tello.cpp:36-78)
            appendRevoluteWithRotor<T>(
                model, side + "-hip-clamp", side + "-hip-clamp", side + "-hip-clamp-rotor", base,
                hip_clamp_inertia, rotor_inertia,
                Transform<T>(P.I3, V3<T>(0., sy * 126e-3, -87e-3)),
                Transform<T>(P.R_down, V3<T>(0., sy * 126e-3, -26e-3)), Axis::Z, Axis::Z,
                P.gear_ratio);

            // Hip differential cluster (Tello.cpp:80-162). Rotor placement is identical on both
            // sides (Tello.hpp:44-48, 75-79).
            {
                const std::string parent = side + "-hip-clamp";
                Body<T> r1 = model.registerBody(side + "-hip-rotor-1", rotor_inertia, parent,
                                                Transform<T>(P.R_left, V3<T>(0., 0.04, 0.)));
                Body<T> r2 = model.registerBody(side + "-hip-rotor-2", rotor_inertia, parent,
                                                Transform<T>(P.R_right, V3<T>(0., -0.04, 0.)));
                Body<T> gimbal = model.registerBody(side + "-gimbal", gimbal_inertia, parent,
                                                    Transform<T>(P.I3, V3<T>(0., 0., -142.5e-3)));
                Body<T> thigh = model.registerBody(side + "-thigh", thigh_inertia, side + "-gimbal",
                                                   Transform<T>(P.I3, V3<T>(0., 0., 0.)));
                std::vector<Body<T>> bodies = {r1, r2, gimbal, thigh};
                JointVec<T> joints = {std::make_shared<SingleRevolute<T>>(Axis::Z),
                                      std::make_shared<SingleRevolute<T>>(Axis::Z),
                                      std::make_shared<SingleRevolute<T>>(Axis::X),
                                      std::make_shared<SingleRevolute<T>>(Axis::Y)};
                auto lc = std::make_shared<GenericImplicitConstraint<T>>(
                    std::vector<bool>{true, true, false, false}, telloHipDifferentialPhi<S>);
                model.appendRegisteredBodiesAsCluster(
                    side + "-hip-differential",
                    std::make_shared<GenericCluster<T>>(bodies, joints, lc));
            }

            // Knee-ankle differential cluster (Tello.cpp:164-260)
            {
                const std::string parent = side + "-thigh";
                Body<T> r1 = model.registerBody(side + "-knee-ankle-rotor-1", rotor_inertia, parent,
                                                Transform<T>(P.R_right, V3<T>(0., 26.55e-3, 0.)));
                Body<T> r2 = model.registerBody(side + "-knee-ankle-rotor-2", rotor_inertia, parent,
                                                Transform<T>(P.R_left, V3<T>(0., -26.55e-3, 0.)));
                Body<T> shin = model.registerBody(side + "-shin", shin_inertia, parent,
                                                  Transform<T>(P.I3, V3<T>(0., 0., -226.8e-3)));
                Body<T> foot = model.registerBody(side + "-foot", foot_inertia, side + "-shin",
                                                  Transform<T>(P.I3, V3<T>(0., 0., -260e-3)));
                std::vector<Body<T>> bodies = {r1, r2, shin, foot};
                JointVec<T> joints = {std::make_shared<SingleRevolute<T>>(Axis::Z),
                                      std::make_shared<SingleRevolute<T>>(Axis::Z),
                                      std::make_shared<SingleRevolute<T>>(Axis::Y),
                                      std::make_shared<SingleRevolute<T>>(Axis::Y)};
                auto lc = std::make_shared<GenericImplicitConstraint<T>>(
                    std::vector<bool>{true, true, false, false}, telloKneeAnkleDifferentialPhi<S>);
                model.appendRegisteredBodiesAsCluster(
                    side + "-knee-ankle-differential",
                    std::make_shared<GenericCluster<T>>(bodies, joints, lc));
            }
        }
        return model;
    }

    ////////////////////////////////////////////////////////////////////////////////////////////
    // TelloWithArms (src/Robots/TelloWithArms.cpp:6-171, TelloWithArms.hpp)
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    ClusterTreeModel<T> buildTelloWithArms()
    {
        ClusterTreeModel<T> model = buildTello<T>();
        const Mat<T> I3 = Mat<T>::Identity(3);

        Mat<T> shoulderRyRotInertia = mat3<T>({0.0013678, 0.0000266, 0.0000021, 0.0000266, 0.0007392,
                                              -0.0000012, 0.0000021, -0.0000012, 0.000884});
        Mat<T> shoulderRxRotInertia = mat3<T>({0.0011524, 0.0000007, 0.0000396, 0.0000007, 0.0011921,
                                              0.0000014, 0.0000396, 0.0000014, 0.0012386});
        Mat<T> shoulderRzRotInertia = mat3<T>({0.0012713, 0.000001, -0.000008, 0.000001, 0.0017477,
                                              -0.0000225, -0.000008, -0.0000225, 0.0008191});
        Mat<T> elbowRotInertia = mat3<T>({0.001570, 0.0000002, 0.0000335, 0.0000002, 0.0016167,
                                         0.000003, 0.0000335, 0.000003, 0.0000619});
        Mat<T> smallRotorZ = mat3<T>({1.084e-4, 0, 0, 0, 1.084e-4, 0, 0, 0, 1.6841e-4});
        Mat<T> RY = coordinateRotation<T>(Axis::Y, T(M_PI / 2));
        Mat<T> RX = coordinateRotation<T>(Axis::X, T(-M_PI / 2));
        Mat<T> smallRotorX = RY.transpose() * smallRotorZ * RY;
        Mat<T> smallRotorY = RX.transpose() * smallRotorZ * RX;
        Mat<T> zero3 = V3<T>(0., 0., 0.);

        auto lr_vec = [](double x, double y, double z, int side)
        { return V3<T>(x, side == 0 ? y : -y, z); };
        auto lr_inertia = [](const Mat<T> &I, int side)
        { return side == 0 ? I : flipAlongAxis(I, Axis::Y); };

        const char *sides[2] = {"left", "right"};
        for (int arm = 0; arm < 2; arm++)
        {
            const std::string s = std::string(sides[arm]) + "-";
            appendRevoluteWithRotor<T>(
                model, s + "shoulder-ry", s + "shoulder-ry", s + "shoulder-ry-rotor", "torso",
                lr_inertia(spatialInertia<T>(T(0.788506), V3<T>(0.009265, 0.052623, -0.0001249),
                                             shoulderRyRotInertia), arm),
                lr_inertia(spatialInertia<T>(T(0.), zero3, smallRotorY), arm),
                Transform<T>(I3, lr_vec(0.01346, 0.17608, 0.24657, arm)),
                Transform<T>(I3, lr_vec(0.01346, 0.16, 0.24657, arm)), Axis::Y, Axis::Y, 6.0);
            appendRevoluteWithRotor<T>(
                model, s + "shoulder-rx", s + "shoulder-rx", s + "shoulder-rx-rotor",
                s + "shoulder-ry",
                lr_inertia(spatialInertia<T>(T(0.80125), V3<T>(0.0006041, 0.0001221, -0.082361),
                                             shoulderRxRotInertia), arm),
                lr_inertia(spatialInertia<T>(T(0.), zero3, smallRotorX), arm),
                Transform<T>(I3, lr_vec(0.0, 0.0575, 0.0, arm)),
                Transform<T>(I3, lr_vec(0, 0.0575, 0, arm)), Axis::X, Axis::X, 6.0);
            appendRevoluteWithRotor<T>(
                model, s + "shoulder-rz", s + "shoulder-rz-link", s + "shoulder-rz-rotor",
                s + "shoulder-rx",
                lr_inertia(spatialInertia<T>(T(0.905588), V3<T>(0.0001703, -0.016797, -0.060),
                                             shoulderRzRotInertia), arm),
                lr_inertia(spatialInertia<T>(T(0.), zero3, smallRotorZ), arm),
                Transform<T>(I3, lr_vec(0.0, 0.0, -0.10250, arm)),
                Transform<T>(I3, lr_vec(0., 0., -0.1025, arm)), Axis::Z, Axis::Z, 6.0);
            appendRevoluteWithRotor<T>(
                model, s + "elbow", s + "elbow-link", s + "elbow-rotor", s + "shoulder-rz-link",
                lr_inertia(spatialInertia<T>(T(0.34839), V3<T>(-0.0059578, 0.000111, -0.0426735),
                                             elbowRotInertia), arm),
                lr_inertia(spatialInertia<T>(T(0.), zero3, smallRotorY), arm),
                Transform<T>(I3, lr_vec(0.0, 0.0, -0.1455, arm)),
                Transform<T>(I3, lr_vec(0., -0.0325, -0.06, arm)), Axis::Y, Axis::Y, 9.0);
        }
        return model;
    }

    ////////////////////////////////////////////////////////////////////////////////////////////
    // MiniCheetah (src/Robots/MiniCheetah.cpp:6-139, MiniCheetah.hpp)
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    ClusterTreeModel<T> buildMiniCheetah(bool quaternion = true)
    {
        ClusterTreeModel<T> model;
        const Mat<T> I3 = Mat<T>::Identity(3);
        Mat<T> RY = coordinateRotation<T>(Axis::Y, T(M_PI / 2));
        Mat<T> RX = coordinateRotation<T>(Axis::X, T(M_PI / 2));
        const T em6 = T(1e-6);
        Mat<T> bodyRotI = mat3<T>({11253, 0, 0, 0, 36203, 0, 0, 0, 42673}) * em6;
        Mat<T> abadRotI = mat3<T>({381, 58, 0.45, 58, 560, 0.95, 0.45, 0.95, 444}) * em6;
        Mat<T> hipRotI = mat3<T>({1983, 245, 13, 245, 2103, 1.5, 13, 1.5, 408}) * em6;
        Mat<T> kneeRotIRotated = mat3<T>({6, 0, 0, 0, 248, 0, 0, 0, 245}) * em6;
        Mat<T> rotorZ = em6 * mat3<T>({33, 0, 0, 0, 33, 0, 0, 0, 63});
        Mat<T> rotorX = RY * rotorZ * RY.transpose();
        Mat<T> rotorY = RX * rotorZ * RX.transpose();
        Mat<T> zero3 = V3<T>(0, 0, 0);

        const std::string torso = "Floating Base";
        {
            Body<T> b = model.registerBody(torso, spatialInertia<T>(T(3.3), zero3, bodyRotI),
                                           "ground", Transform<T>());
            model.appendRegisteredBodiesAsCluster(torso,
                                                  std::make_shared<FreeCluster<T>>(b, quaternion));
        }

        auto legSigns = [](double x, double y, double z, int leg)
        {
            switch (leg)
            {
            case 0: return V3<T>(x, -y, z);
            case 1: return V3<T>(x, y, z);
            case 2: return V3<T>(-x, -y, z);
            default: return V3<T>(-x, y, z);
            }
        };
        const char *prefix[4] = {"FR_", "FL_", "HR_", "HL_"};
        auto lr_inertia = [](const Mat<T> &I, int sideSign)
        { return sideSign <= 0 ? flipAlongAxis(I, Axis::Y) : I; };

        int sideSign = -1;
        for (int leg : {2, 3, 0, 1})
        {
            const std::string p = prefix[leg];
            appendRevoluteWithRotor<T>(
                model, p + "abad", p + "abad_link", p + "abad_rotor", torso,
                lr_inertia(spatialInertia<T>(T(0.54), V3<T>(0, 0.036, 0), abadRotI), sideSign),
                lr_inertia(spatialInertia<T>(T(0.055), zero3, rotorX), sideSign),
                Transform<T>(I3, legSigns(0.38 * 0.5, 0.098 * 0.5, 0 * 0.5, leg)),
                Transform<T>(I3, legSigns(0.125, 0.049, 0, leg)), Axis::X, Axis::X, 6);

            Mat<T> RZ = coordinateRotation<T>(Axis::Z, T(M_PI));
            appendRevoluteWithRotor<T>(
                model, p + "hip", p + "hip_link", p + "hip_rotor", p + "abad_link",
                lr_inertia(spatialInertia<T>(T(0.634), V3<T>(0, 0.016, -0.02), hipRotI), sideSign),
                lr_inertia(spatialInertia<T>(T(0.055), zero3, rotorY), sideSign),
                Transform<T>(RZ, legSigns(0, 0.062, 0, leg)),
                Transform<T>(RZ, legSigns(0, 0.04, 0, leg)), Axis::Y, Axis::Y, 6);

            appendRevoluteWithRotor<T>(
                model, p + "knee", p + "knee_link", p + "knee_rotor", p + "hip_link",
                lr_inertia(spatialInertia<T>(T(0.064), V3<T>(0, 0, -0.061), kneeRotIRotated),
                           sideSign),
                lr_inertia(spatialInertia<T>(T(0.055), zero3, rotorY), sideSign),
                Transform<T>(I3, legSigns(0, 0, -0.209, leg)),
                Transform<T>(I3, legSigns(0, 0, 0, leg)), Axis::Y, Axis::Y, 9.33);
            sideSign *= -1;
        }
        return model;
    }

    ////////////////////////////////////////////////////////////////////////////////////////////
    // MIT_Humanoid (src/Robots/MIT_Humanoid.cpp:6-358, MIT_Humanoid.hpp)
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    ClusterTreeModel<T> buildMitHumanoid(bool quaternion = true)
    {
        ClusterTreeModel<T> model;
        const Mat<T> I3 = Mat<T>::Identity(3);
        Mat<T> torsoRotI = mat3<T>({0.172699, 0.001419, 0.004023, 0.001419, 0.105949, -0.001672,
                                   0.004023, -0.001672, 0.091906});
        Mat<T> hipRzRotI = mat3<T>({0.0015373, 0.0000011, 0.0005578, 0.0000011, 0.0014252,
                                   0.0000024, 0.0005578, 0.0000024, 0.0012028});
        Mat<T> hipRxRotI = mat3<T>({0.0017535, -0.0000063, -0.000080, -0.0000063, 0.003338,
                                   -0.000013, -0.000080, -0.000013, 0.0019927});
        Mat<T> hipRyRotI = mat3<T>({0.0243761, 0.0000996, 0.0006548, 0.0000996, 0.0259015,
                                   0.0026713, 0.0006548, 0.0026713, 0.0038929});
        Mat<T> kneeRotI = mat3<T>({0.003051, 0.000000, 0.0000873, 0.000000, 0.003033, 0.0000393,
                                  0.0000873, 0.0000393, 0.0002529});
        Mat<T> ankleRotI = mat3<T>({0.0000842, 0.000000, -0.0000488, 0.000000, 0.0007959,
                                   -0.000000, -0.0000488, -0.000000, 0.0007681});
        Mat<T> shoulderRyRotI = mat3<T>({0.0013678, 0.0000266, 0.0000021, 0.0000266, 0.0007392,
                                        -0.0000012, 0.0000021, -0.0000012, 0.000884});
        Mat<T> shoulderRxRotI = mat3<T>({0.0011524, 0.0000007, 0.0000396, 0.0000007, 0.0011921,
                                        0.0000014, 0.0000396, 0.0000014, 0.0012386});
        Mat<T> shoulderRzRotI = mat3<T>({0.0012713, 0.000001, -0.000008, 0.000001, 0.0017477,
                                        -0.0000225, -0.000008, -0.0000225, 0.0008191});
        Mat<T> elbowRotI = mat3<T>({0.001570, 0.0000002, 0.0000335, 0.0000002, 0.0016167, 0.000003,
                                   0.0000335, 0.000003, 0.0000619});
        Mat<T> largeRotorZ = mat3<T>({3.443e-4, 0, 0, 0, 3.443e-4, 0, 0, 0, 5.548e-4});
        Mat<T> smallRotorZ = mat3<T>({1.084e-4, 0, 0, 0, 1.084e-4, 0, 0, 0, 1.6841e-4});
        Mat<T> RY = coordinateRotation<T>(Axis::Y, T(M_PI / 2));
        Mat<T> RX = coordinateRotation<T>(Axis::X, T(-M_PI / 2));
        Mat<T> smallRotorX = RY.transpose() * smallRotorZ * RY;
        Mat<T> smallRotorY = RX.transpose() * smallRotorZ * RX;
        Mat<T> largeRotorY = RX.transpose() * largeRotorZ * RX;
        Mat<T> zero3 = V3<T>(0, 0, 0);
        const double hipRzPitch = -0.174533, hipRxPitch = 0.436332;
        const double hipRyPitch = -(hipRxPitch + hipRzPitch);
        const double smallRotorMass = 0.05, largeRotorMass = 0.1;

        const std::string torso = "Floating Base";
        {
            Body<T> b = model.registerBody(
                torso, spatialInertia<T>(T(8.52), V3<T>(0.009896, 0.004771, 0.100522), torsoRotI),
                "ground", Transform<T>());
            model.appendRegisteredBodiesAsCluster(torso,
                                                  std::make_shared<FreeCluster<T>>(b, quaternion));
        }

        auto lr_vec = [](double x, double y, double z, int side)
        { return V3<T>(x, side == 0 ? y : -y, z); };
        auto lr_name = [](const std::string &s, int side)
        { return (side == 0 ? "right_" : "left_") + s; };
        auto lr_inertia = [](const Mat<T> &I, int side)
        { return side == 0 ? flipAlongAxis(I, Axis::Y) : I; };

        auto appendLeg = [&](int leg)
        {
            Mat<T> Xrot_HipZ = coordinateRotation<T>(Axis::Y, T(hipRzPitch));
            appendRevoluteWithRotor<T>(
                model, lr_name("hip_rz", leg), lr_name("hip_rz_link", leg),
                lr_name("hip_rz_rotor", leg), torso,
                lr_inertia(spatialInertia<T>(T(0.84563), V3<T>(-0.064842, -0.000036, -0.063090),
                                             hipRzRotI), leg),
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorZ), leg),
                Transform<T>(Xrot_HipZ, lr_vec(-0.00565, -0.082, -0.05735, leg)),
                Transform<T>(Xrot_HipZ, lr_vec(-0.00842837, -0.082, -0.041593, leg)), Axis::Z,
                Axis::Z, 6.0);
            Mat<T> Xrot_HipX = coordinateRotation<T>(Axis::Y, T(hipRxPitch));
            appendRevoluteWithRotor<T>(
                model, lr_name("hip_rx", leg), lr_name("hip_rx_link", leg),
                lr_name("hip_rx_rotor", leg), lr_name("hip_rz_link", leg),
                lr_inertia(spatialInertia<T>(T(1.20868), V3<T>(0.067232, -0.013018, 0.0001831),
                                             hipRxRotI), leg),
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorX), leg),
                Transform<T>(Xrot_HipX, lr_vec(-0.06435, 0.0, -.07499, leg)),
                Transform<T>(Xrot_HipX, lr_vec(-0.0827, 0.0, -0.066436, leg)), Axis::X, Axis::X,
                6.0);
            Mat<T> Xrot_HipY = coordinateRotation<T>(Axis::Y, T(hipRyPitch));
            appendRevoluteWithRotor<T>(
                model, lr_name("hip_ry", leg), lr_name("hip_ry_link", leg),
                lr_name("hip_ry_rotor", leg), lr_name("hip_rx_link", leg),
                lr_inertia(spatialInertia<T>(T(2.64093), V3<T>(0.0132054, 0.0269864, -0.096021),
                                             hipRyRotI), leg),
                lr_inertia(spatialInertia<T>(T(largeRotorMass), zero3, largeRotorY), leg),
                Transform<T>(Xrot_HipY, lr_vec(0.071, 0.0018375, 0.0, leg)),
                Transform<T>(Xrot_HipY, lr_vec(0.071, 0.024, 0.0, leg)), Axis::Y, Axis::Y, 6.0);

            // Knee + ankle cluster (MIT_Humanoid.cpp:126-192); registration order is
            // ankle_rotor, knee_link, knee_rotor, ankle_link (:172-179)
            const std::string knee_parent = lr_name("hip_ry_link", leg);
            Mat<T> knee_link_inertia = lr_inertia(
                spatialInertia<T>(T(0.35435), V3<T>(0.00528, 0.0014762, -0.13201), kneeRotI), leg);
            Mat<T> knee_rotor_inertia =
                lr_inertia(spatialInertia<T>(T(largeRotorMass), zero3, largeRotorY), leg);
            Mat<T> ankle_link_inertia = lr_inertia(
                spatialInertia<T>(T(0.280951), V3<T>(0.022623, 0.0, -0.012826), ankleRotI), leg);
            Mat<T> ankle_rotor_inertia =
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorY), leg);
            Body<T> ankle_rotor = model.registerBody(
                lr_name("ankle_rotor", leg), ankle_rotor_inertia, knee_parent,
                Transform<T>(I3, lr_vec(.01563, -.0454, -.13354, leg)));
            Body<T> knee_link = model.registerBody(lr_name("knee_link", leg), knee_link_inertia,
                                                   knee_parent,
                                                   Transform<T>(I3, lr_vec(0.0, 0.0, -0.267, leg)));
            Body<T> knee_rotor = model.registerBody(
                lr_name("knee_rotor", leg), knee_rotor_inertia, knee_parent,
                Transform<T>(I3, lr_vec(0.013, -0.0497, -0.0178, leg)));
            Body<T> ankle_link = model.registerBody(
                lr_name("ankle_link", leg), ankle_link_inertia, lr_name("knee_link", leg),
                Transform<T>(I3, lr_vec(0.0, 0.0, -0.2785, leg)));
            ParallelBeltTransmissionModule<T> knee_module{knee_link, knee_rotor, Axis::Y, Axis::Y,
                                                          T(6.0), {T(2.0)}};
            ParallelBeltTransmissionModule<T> ankle_module{ankle_link, ankle_rotor, Axis::Y,
                                                           Axis::Y, T(6.0), {T(2.0), T(1.0)}};
            model.appendRegisteredBodiesAsCluster(
                lr_name("knee_and_ankle", leg),
                std::make_shared<RevolutePairWithRotorCluster<T>>(knee_module, ankle_module));
        };

        auto appendArm = [&](int arm)
        {
            appendRevoluteWithRotor<T>(
                model, lr_name("shoulder_ry", arm), lr_name("shoulder_ry_link", arm),
                lr_name("shoulder_ry_rotor", arm), torso,
                lr_inertia(spatialInertia<T>(T(0.788506), V3<T>(0.009265, 0.052623, -0.0001249),
                                             shoulderRyRotI), arm),
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorY), arm),
                Transform<T>(I3, lr_vec(0.01346, -0.17608, 0.24657, arm)),
                Transform<T>(I3, lr_vec(0.01346, -0.16, 0.24657, arm)), Axis::Y, Axis::Y, 6.0);
            appendRevoluteWithRotor<T>(
                model, lr_name("shoulder_rx", arm), lr_name("shoulder_rx_link", arm),
                lr_name("shoulder_rx_rotor", arm), lr_name("shoulder_ry_link", arm),
                lr_inertia(spatialInertia<T>(T(0.80125), V3<T>(0.0006041, 0.0001221, -0.082361),
                                             shoulderRxRotI), arm),
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorX), arm),
                Transform<T>(I3, lr_vec(0.0, -0.0575, 0.0, arm)),
                Transform<T>(I3, lr_vec(0, -0.0575, 0, arm)), Axis::X, Axis::X, 6.0);
            appendRevoluteWithRotor<T>(
                model, lr_name("shoulder_rz", arm), lr_name("shoulder_rz_link", arm),
                lr_name("shoulder_rz_rotor", arm), lr_name("shoulder_rx_link", arm),
                lr_inertia(spatialInertia<T>(T(0.905588), V3<T>(0.0001703, -0.016797, -0.060),
                                             shoulderRzRotI), arm),
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorZ), arm),
                Transform<T>(I3, lr_vec(0.0, 0.0, -0.10250, arm)),
                Transform<T>(I3, lr_vec(0., 0., -0.1025, arm)), Axis::Z, Axis::Z, 6.0);
            appendRevoluteWithRotor<T>(
                model, lr_name("elbow", arm), lr_name("elbow_link", arm),
                lr_name("elbow_rotor", arm), lr_name("shoulder_rz_link", arm),
                lr_inertia(spatialInertia<T>(T(0.34839), V3<T>(-0.0059578, 0.000111, -0.0426735),
                                             elbowRotI), arm),
                lr_inertia(spatialInertia<T>(T(smallRotorMass), zero3, smallRotorY), arm),
                Transform<T>(I3, lr_vec(0.0, 0.0, -0.1455, arm)),
                Transform<T>(I3, lr_vec(0., 0.0325, -0.06, arm)), Axis::Y, Axis::Y, 9.0);
        };

        appendArm(0);
        appendLeg(0);
        appendArm(1);
        appendLeg(1);
        return model;
    }

    ////////////////////////////////////////////////////////////////////////////////////////////
    // Uniform serial chains (RevoluteChainWithRotor.cpp:45-109, RevolutePairChainWithRotor.cpp:62-128)
    ////////////////////////////////////////////////////////////////////////////////////////////
    ////////////////////////////////////////////////////////////////////////////////////////////
    // MIT_Humanoid_Leg (src/Robots/MIT_Humanoid_Leg.cpp:6-164): one leg of the MIT humanoid on a fixed base,
    // with the raw (unsigned) constants of MIT_Humanoid.hpp:19-65,74-143 and MASSLESS rotors (:25,52,77,104,117).
    // The reference compares it with robot-models/mit_humanoid_leg.urdf (UnitTests/testClusterTreeModel.cpp:100-113).
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    ClusterTreeModel<T> buildMitHumanoidLeg()
    {
        ClusterTreeModel<T> model;
        const Mat<T> I3 = Mat<T>::Identity(3);
        Mat<T> hipRzRotI = mat3<T>({0.0015373, 0.0000011, 0.0005578, 0.0000011, 0.0014252,
                                   0.0000024, 0.0005578, 0.0000024, 0.0012028});
        Mat<T> hipRxRotI = mat3<T>({0.0017535, -0.0000063, -0.000080, -0.0000063, 0.003338,
                                   -0.000013, -0.000080, -0.000013, 0.0019927});
        Mat<T> hipRyRotI = mat3<T>({0.0243761, 0.0000996, 0.0006548, 0.0000996, 0.0259015,
                                   0.0026713, 0.0006548, 0.0026713, 0.0038929});
        Mat<T> kneeRotI = mat3<T>({0.003051, 0.000000, 0.0000873, 0.000000, 0.003033, 0.0000393,
                                  0.0000873, 0.0000393, 0.0002529});
        Mat<T> ankleRotI = mat3<T>({0.0000842, 0.000000, -0.0000488, 0.000000, 0.0007959,
                                   -0.000000, -0.0000488, -0.000000, 0.0007681});
        Mat<T> largeRotorZ = mat3<T>({3.443e-4, 0, 0, 0, 3.443e-4, 0, 0, 0, 5.548e-4});
        Mat<T> smallRotorZ = mat3<T>({1.084e-4, 0, 0, 0, 1.084e-4, 0, 0, 0, 1.6841e-4});
        Mat<T> RY = coordinateRotation<T>(Axis::Y, T(M_PI / 2));
        Mat<T> RX = coordinateRotation<T>(Axis::X, T(-M_PI / 2));
        Mat<T> smallRotorX = RY.transpose() * smallRotorZ * RY;
        Mat<T> smallRotorY = RX.transpose() * smallRotorZ * RX;
        Mat<T> largeRotorY = RX.transpose() * largeRotorZ * RX;
        Mat<T> zero3 = V3<T>(0, 0, 0);
        const double hipRzPitch = -0.174533, hipRxPitch = 0.436332;
        const double hipRyPitch = -(hipRxPitch + hipRzPitch);

        Mat<T> Xrot_HipZ = coordinateRotation<T>(Axis::Y, T(hipRzPitch));
        appendRevoluteWithRotor<T>(model, "hip_rz", "hip_rz_link", "hip_rz_rotor", "ground",
                                   spatialInertia<T>(T(0.84563), V3<T>(-0.064842, -0.000036, -0.063090), hipRzRotI),
                                   spatialInertia<T>(T(0.0), zero3, smallRotorZ),
                                   Transform<T>(Xrot_HipZ, V3<T>(-0.00565, -0.082, -0.05735)),
                                   Transform<T>(Xrot_HipZ, V3<T>(-0.00842837, -0.082, -0.041593)), Axis::Z, Axis::Z, 6.0);
        Mat<T> Xrot_HipX = coordinateRotation<T>(Axis::Y, T(hipRxPitch));
        appendRevoluteWithRotor<T>(model, "hip_rx", "hip_rx_link", "hip_rx_rotor", "hip_rz_link",
                                   spatialInertia<T>(T(1.20868), V3<T>(0.067232, -0.013018, 0.0001831), hipRxRotI),
                                   spatialInertia<T>(T(0.0), zero3, smallRotorX),
                                   Transform<T>(Xrot_HipX, V3<T>(-0.06435, 0.0, -.07499)),
                                   Transform<T>(Xrot_HipX, V3<T>(-0.0827, 0.0, -0.066436)), Axis::X, Axis::X, 6.0);
        Mat<T> Xrot_HipY = coordinateRotation<T>(Axis::Y, T(hipRyPitch));
        appendRevoluteWithRotor<T>(model, "hip_ry", "hip_ry_link", "hip_ry_rotor", "hip_rx_link",
                                   spatialInertia<T>(T(2.64093), V3<T>(0.0132054, 0.0269864, -0.096021), hipRyRotI),
                                   spatialInertia<T>(T(0.0), zero3, largeRotorY),
                                   Transform<T>(Xrot_HipY, V3<T>(0.071, 0.0018375, 0.0)),
                                   Transform<T>(Xrot_HipY, V3<T>(0.071, 0.024, 0.0)), Axis::Y, Axis::Y, 6.0);
        // knee + ankle: registration order knee_link, ankle_rotor, knee_rotor, ankle_link (:134-141)
        Body<T> knee_link = model.registerBody(
            "knee_link", spatialInertia<T>(T(0.35435), V3<T>(0.00528, 0.0014762, -0.13201), kneeRotI), "hip_ry_link",
            Transform<T>(I3, V3<T>(0.0, 0.0, -0.267)));
        Body<T> ankle_rotor = model.registerBody("ankle_rotor", spatialInertia<T>(T(0.0), zero3, smallRotorY), "hip_ry_link",
                                                 Transform<T>(I3, V3<T>(.01563, -.0454, -.13354)));
        Body<T> knee_rotor = model.registerBody("knee_rotor", spatialInertia<T>(T(0.0), zero3, largeRotorY), "hip_ry_link",
                                                Transform<T>(I3, V3<T>(0.013, -0.0497, -0.0178)));
        Body<T> ankle_link = model.registerBody(
            "ankle_link", spatialInertia<T>(T(0.280951), V3<T>(0.022623, 0.0, -0.012826), ankleRotI), "knee_link",
            Transform<T>(I3, V3<T>(0.0, 0.0, -0.2785)));
        ParallelBeltTransmissionModule<T> knee_module{knee_link, knee_rotor, Axis::Y, Axis::Y, T(6.0), {T(2.0)}};
        ParallelBeltTransmissionModule<T> ankle_module{ankle_link, ankle_rotor, Axis::Y, Axis::Y, T(6.0), {T(2.0), T(1.0)}};
        model.appendRegisteredBodiesAsCluster("knee_and_ankle",
                                              std::make_shared<RevolutePairWithRotorCluster<T>>(knee_module, ankle_module));
        return model;
    }

    template <typename T>
    ClusterTreeModel<T> buildRevoluteChainWithRotor(int N)
    {
        ClusterTreeModel<T> model;
        const Mat<T> I3 = Mat<T>::Identity(3);
        model.setGravity(9.81, 0., 0.);
        const double I = 1., Irot = 1e-4, m = 1., l = 1., c = 0.5, gr = 2., br = 3.;
        Mat<T> link_inertia = spatialInertia<T>(T(m), V3<T>(c, 0., 0.),
                                                mat3<T>({0., 0., 0., 0., 0., 0., 0., 0., I}));
        Mat<T> rotor_inertia = spatialInertia<T>(T(0.), V3<T>(0., 0., 0.),
                                                 mat3<T>({0., 0., 0., 0., 0., 0., 0., 0., Irot}));
        std::string prev = "ground";
        for (int i = 0; i < N; i++)
        {
            Transform<T> Xtree = i == 0 ? Transform<T>(I3, V3<T>(0., 0., 0.))
                                        : Transform<T>(I3, V3<T>(l, 0., 0.));
            const std::string link = "link-" + std::to_string(i);
            appendRevoluteWithRotor<T>(model, "cluster-" + std::to_string(i), link,
                                       "rotor-" + std::to_string(i), prev, link_inertia,
                                       rotor_inertia, Xtree, Xtree, Axis::Z, Axis::Z, gr * br);
            prev = link;
        }
        return model;
    }

    template <typename T>
    ClusterTreeModel<T> buildRevolutePairChainWithRotor(int N)
    {
        ClusterTreeModel<T> model;
        const Mat<T> I3 = Mat<T>::Identity(3);
        model.setGravity(9.81, 0., 0.);
        const double I = 1., Irot = 1e-4, m = 1., l = 1., c = 0.5, gr = 2., br = 3.;
        Mat<T> link_inertia = spatialInertia<T>(T(m), V3<T>(c, 0., 0.),
                                                mat3<T>({0., 0., 0., 0., 0., 0., 0., 0., I}));
        Mat<T> rotor_inertia = spatialInertia<T>(T(0.), V3<T>(0., 0., 0.),
                                                 mat3<T>({0., 0., 0., 0., 0., 0., 0., 0., Irot}));
        const Transform<T> Xtree2(I3, V3<T>(l, 0., 0.));
        std::string parent = "ground";
        for (int i = 0; i < N / 2; i++)
        {
            const Transform<T> Xtree1 = i == 0 ? Transform<T>(I3, V3<T>(0., 0., 0.)) : Xtree2;
            const std::string is = std::to_string(i);
            Body<T> linkA = model.registerBody("link-A-" + is, link_inertia, parent, Xtree1);
            Body<T> rotorA = model.registerBody("rotor-A-" + is, rotor_inertia, parent, Xtree1);
            Body<T> rotorB = model.registerBody("rotor-B-" + is, rotor_inertia, parent, Xtree1);
            Body<T> linkB = model.registerBody("link-B-" + is, link_inertia, "link-A-" + is, Xtree2);
            ParallelBeltTransmissionModule<T> mA{linkA, rotorA, Axis::Z, Axis::Z, T(gr), {T(br)}};
            ParallelBeltTransmissionModule<T> mB{linkB, rotorB, Axis::Z, Axis::Z, T(gr),
                                                 {T(br), T(1.)}};
            model.appendRegisteredBodiesAsCluster(
                "cluster-" + is, std::make_shared<RevolutePairWithRotorCluster<T>>(mA, mB));
            parent = "link-B-" + is;
        }
        return model;
    }

    ////////////////////////////////////////////////////////////////////////////////////////////
    // Generic re-build of any model (UnitTests/testHelpers.hpp:10-45, extractGenericJointModel):
    // every cluster becomes a ClusterJoints::Generic with the same bodies, joints and constraint.
    ////////////////////////////////////////////////////////////////////////////////////////////
    template <typename T>
    ClusterTreeModel<T> extractGenericJointModel(const ClusterTreeModel<T> &model)
    {
        ClusterTreeModel<T> generic;
        generic.gravity = model.gravity;
        for (auto &cluster : model.nodes)
        {
            std::vector<Body<T>> bodies;
            for (auto &b : cluster->bodies)
            {
                std::string parent =
                    b.parent_index == -1 ? "ground" : model.bodies[b.parent_index].name;
                bodies.push_back(generic.registerBody(b.name, b.inertia, parent, b.Xtree));
            }
            generic.appendRegisteredBodiesAsCluster(
                cluster->name,
                std::make_shared<GenericCluster<T>>(bodies, cluster->joint->single_joints,
                                                    cluster->joint->loop_constraint));
        }
        return generic;
    }

} // namespace grbda_oracle
