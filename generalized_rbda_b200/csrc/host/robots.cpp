// Hand-coded robot models of the reference that have no URDF+ description:
//   Tello / TelloWithArms     reference: src/Robots/Tello.cpp:6-277, TelloWithArms.cpp:6-171
//                             (tello_humanoid.urdf carries no <loop>/<coupling> tags)
//   uniform serial chains     reference: src/Robots/SerialChains/RevoluteChainWithRotor.cpp:45-109,
//                             RevolutePairChainWithRotor.cpp:62-128 (known-answer models)
// MiniCheetah / MIT_Humanoid / four_bar / revolute_rotor_chain / JVRC1 are built from the URDF+
// files (host/urdf.cpp), as the reference's ClusterTreeModel(urdf_file) constructor does.
//
// Written table-driven (one row per body) rather than statement by statement; the numbers are the
// robot's physical parameters from include/grbda/Robots/{Tello,TelloWithArms}.hpp.
#include "robots.h"

namespace grbda
{
    using ori::CoordinateAxis;
    using namespace ClusterJoints;

    namespace
    {
        const Mat3 I3 = {1, 0, 0, 0, 1, 0, 0, 0, 1};

        struct LinkRow
        {
            const char *name;
            double mass;
            Vec3 com;
            Mat3 inertia;
        };

        SpatialInertia inertiaOf(const LinkRow &r) { return SpatialInertia(r.mass, r.com, r.inertia); }

        void appendGeared(ClusterTreeModel &model, const std::string &cluster, const std::string &link,
                          const std::string &rotor, const std::string &parent, const SpatialInertia &link_I,
                          const SpatialInertia &rotor_I, const spatial::Transform &link_X,
                          const spatial::Transform &rotor_X, CoordinateAxis axis, double gear_ratio)
        {
            Body l = model.registerBody(link, link_I, parent, link_X);
            Body r = model.registerBody(rotor, rotor_I, parent, rotor_X);
            GearedTransmissionModule module{l, r, link + "-joint", rotor + "-joint", axis, axis, gear_ratio};
            model.appendRegisteredBodiesAsCluster<RevoluteWithRotor>(cluster, module);
        }
    } // namespace

    // -----------------------------------------------------------------------------------------
    // Tello differential constraints written over sym::Sym (counterpart of the casadi::SX lambdas).
    // The C++ integer divisions `3021 / 160000` and `163349 / 6250000` of the reference evaluate
    // to 0 and are therefore absent; they do not affect K, G, k, g.
    // -----------------------------------------------------------------------------------------
    static std::vector<sym::Sym> telloHipDifferentialPhi(const std::vector<sym::Sym> &q)
    {
        using sym::Sym;
        using sym::sin;
        using sym::cos;
        const Sym N(6.0);
        const Sym ql1 = q[0], ql2 = q[1], y1 = q[2] / N, y2 = q[3] / N;
        auto row = [&](const Sym &y, double sgn)
        {
            return Sym(57.) * sin(y) / Sym(2500.) - Sym(49.) * cos(ql1) / Sym(5000.) -
                   Sym(sgn * 399.) * sin(ql1) / Sym(20000.) - Sym(8.) * cos(y) * cos(ql2) / Sym(625.) -
                   Sym(57.) * cos(ql1) * sin(ql2) / Sym(2500.) - Sym(sgn * 7.) * sin(y) * sin(ql1) / Sym(625.) +
                   Sym(sgn * 7.) * sin(ql1) * sin(ql2) / Sym(625.) -
                   Sym(8.) * cos(ql1) * sin(y) * sin(ql2) / Sym(625.);
        };
        return {row(y1, 1.0), row(y2, -1.0)};
    }

    static std::vector<sym::Sym> telloKneeAnkleDifferentialPhi(const std::vector<sym::Sym> &q)
    {
        using sym::Sym;
        using sym::sin;
        using sym::cos;
        const Sym N(6.0), half(0.5);
        const double pi = 3.1415; // sic (Tello.cpp:247)
        const Sym ql1 = q[0], ql2 = q[1], y1 = q[2] / N, y2 = q[3] / N;
        const Sym d = half * y1 - half * y2;
        const Sym r0 = Sym(21.) * cos(d + Sym(1979 * pi / 4500)) / Sym(6250.) -
                       Sym(13.) * cos(d + Sym(493 * pi / 1500)) / Sym(625.) -
                       Sym(273 * std::cos(pi / 9) / 12500) -
                       Sym(7.) * sin(d + ql2 + Sym(231 * pi / 500)) / Sym(2500.) +
                       Sym(91.) * sin(ql2 + Sym(2 * pi / 15)) / Sym(5000.) -
                       Sym(147.) * sin(ql2 + Sym(pi / 45)) / Sym(50000.);
        const Sym r1 = ql1 - half * y2 - half * y1;
        return {r0, r1};
    }

    ClusterTreeModel Tello::buildClusterTreeModel() const
    {
        ClusterTreeModel model;
        model.setGravity({0., 0., -9.81});

        const LinkRow torso{"torso", 2.3008, {0.0073, -0.0013, -0.0023},
                            {0.0366, 0., -0.0006, 0., 0.0142, -0.0002, -0.0006, -0.0002, 0.0291}};
        const LinkRow hip_clamp{"hip-clamp", 1.3289, {-0.0010, 0., -0.0069},
                                {0.0032, 0., 0.0001, 0., 0.0033, 0., 0.0001, 0., 0.0027}};
        const LinkRow gimbal{"gimbal", 0.4433, {-0.0027, 0., 0.0258},
                             {0.0018, 0., 0., 0., 0.0017, 0., 0., 0., 0.0015}};
        const LinkRow thigh{"thigh", 1.5424, {0.003, -0.0001, -0.0323},
                            {0.0103, 0., -0.0005, 0., 0.0097, 0., -0.0005, 0., 0.0027}};
        const LinkRow shin{"shin", 0.3072, {0.0047, -0.0003, -0.1043},
                           {0.0054, -0., -0.0002, -0., 0.0054, 0., -0.0002, 0., 0.0001}};
        const LinkRow foot{"foot", 0.1025, {0.0042, -0., -0.0251},
                           {0.094e-3, -0., -0.0038e-3, -0., 0.1773e-3, 0., -0.0038e-3, 0., 0.0901e-3}};
        const LinkRow rotor{"rotor", 0.07, {0., 0., 0.},
                            {2.5984e-5, 0., 0., 0., 2.5984e-5, 0., 0., 0., 5.1512e-5}};
        const Mat3 R_down = {1., 0., 0., 0., -1., 0., 0., 0., -1.};
        const Mat3 R_left = {-1., 0., 0., 0., 0., 1., 0., 1., 0.};
        const Mat3 R_right = {1., 0., 0., 0., 0., 1., 0., -1., 0.};
        const double gear_ratio = 6.0;

        model.appendBody<Free>("torso", inertiaOf(torso), "ground", spatial::Transform{}, true);

        for (const std::string side : {"left", "right"})
        {
            const double sy = side == "left" ? 1.0 : -1.0;
            appendGeared(model, side + "-hip-clamp", side + "-hip-clamp", side + "-hip-clamp-rotor", "torso",
                         inertiaOf(hip_clamp), inertiaOf(rotor),
                         spatial::Transform(I3, {0., sy * 126e-3, -87e-3}),
                         spatial::Transform(R_down, {0., sy * 126e-3, -26e-3}), CoordinateAxis::Z, gear_ratio);

            // hip differential: two rotors driving gimbal (Rx) and thigh (Ry) through a linkage
            {
                const std::string parent = side + "-hip-clamp";
                std::vector<Body> bodies = {
                    model.registerBody(side + "-hip-rotor-1", inertiaOf(rotor), parent,
                                       spatial::Transform(R_left, {0., 0.04, 0.})),
                    model.registerBody(side + "-hip-rotor-2", inertiaOf(rotor), parent,
                                       spatial::Transform(R_right, {0., -0.04, 0.})),
                    model.registerBody(side + "-gimbal", inertiaOf(gimbal), parent,
                                       spatial::Transform(I3, {0., 0., -142.5e-3})),
                    model.registerBody(side + "-thigh", inertiaOf(thigh), side + "-gimbal",
                                       spatial::Transform(I3, {0., 0., 0.}))};
                const std::vector<CoordinateAxis> axes = {CoordinateAxis::Z, CoordinateAxis::Z,
                                                          CoordinateAxis::X, CoordinateAxis::Y};
                LoopConstraint::GenericImplicit lc({true, true, false, false}, telloHipDifferentialPhi);
                model.appendRegisteredBodiesAsCluster<Generic>(side + "-hip-differential", bodies, axes, lc);
            }
            // knee-ankle differential
            {
                const std::string parent = side + "-thigh";
                std::vector<Body> bodies = {
                    model.registerBody(side + "-knee-ankle-rotor-1", inertiaOf(rotor), parent,
                                       spatial::Transform(R_right, {0., 26.55e-3, 0.})),
                    model.registerBody(side + "-knee-ankle-rotor-2", inertiaOf(rotor), parent,
                                       spatial::Transform(R_left, {0., -26.55e-3, 0.})),
                    model.registerBody(side + "-shin", inertiaOf(shin), parent,
                                       spatial::Transform(I3, {0., 0., -226.8e-3})),
                    model.registerBody(side + "-foot", inertiaOf(foot), side + "-shin",
                                       spatial::Transform(I3, {0., 0., -260e-3}))};
                const std::vector<CoordinateAxis> axes = {CoordinateAxis::Z, CoordinateAxis::Z,
                                                          CoordinateAxis::Y, CoordinateAxis::Y};
                LoopConstraint::GenericImplicit lc({true, true, false, false}, telloKneeAnkleDifferentialPhi);
                model.appendRegisteredBodiesAsCluster<Generic>(side + "-knee-ankle-differential", bodies,
                                                               axes, lc);
            }
        }
        return model;
    }

    ClusterTreeModel TelloWithArms::buildClusterTreeModel() const
    {
        ClusterTreeModel model = Tello::buildClusterTreeModel();

        const Mat3 small_rotor_Z = {1.084e-4, 0, 0, 0, 1.084e-4, 0, 0, 0, 1.6841e-4};
        const Mat3 RY = ori::coordinateRotation(CoordinateAxis::Y, M_PI / 2);
        const Mat3 RX = ori::coordinateRotation(CoordinateAxis::X, -M_PI / 2);
        const Mat3 small_rotor_X = ori::mul(ori::mul(ori::transpose(RY), small_rotor_Z), RY);
        const Mat3 small_rotor_Y = ori::mul(ori::mul(ori::transpose(RX), small_rotor_Z), RX);

        struct ArmRow
        {
            const char *cluster, *link, *rotor, *parent;
            LinkRow inertial;
            const Mat3 *rotor_inertia;
            Vec3 location, rotor_location;
            CoordinateAxis axis;
            double gear_ratio;
        };
        const ArmRow rows[4] = {
            {"shoulder-ry", "shoulder-ry", "shoulder-ry-rotor", "torso",
             {"", 0.788506, {0.009265, 0.052623, -0.0001249},
              {0.0013678, 0.0000266, 0.0000021, 0.0000266, 0.0007392, -0.0000012, 0.0000021, -0.0000012, 0.000884}},
             &small_rotor_Y, {0.01346, 0.17608, 0.24657}, {0.01346, 0.16, 0.24657}, CoordinateAxis::Y, 6.0},
            {"shoulder-rx", "shoulder-rx", "shoulder-rx-rotor", "shoulder-ry",
             {"", 0.80125, {0.0006041, 0.0001221, -0.082361},
              {0.0011524, 0.0000007, 0.0000396, 0.0000007, 0.0011921, 0.0000014, 0.0000396, 0.0000014, 0.0012386}},
             &small_rotor_X, {0.0, 0.0575, 0.0}, {0, 0.0575, 0}, CoordinateAxis::X, 6.0},
            {"shoulder-rz", "shoulder-rz-link", "shoulder-rz-rotor", "shoulder-rx",
             {"", 0.905588, {0.0001703, -0.016797, -0.060},
              {0.0012713, 0.000001, -0.000008, 0.000001, 0.0017477, -0.0000225, -0.000008, -0.0000225, 0.0008191}},
             &small_rotor_Z, {0.0, 0.0, -0.10250}, {0., 0., -0.1025}, CoordinateAxis::Z, 6.0},
            {"elbow", "elbow-link", "elbow-rotor", "shoulder-rz-link",
             {"", 0.34839, {-0.0059578, 0.000111, -0.0426735},
              {0.001570, 0.0000002, 0.0000335, 0.0000002, 0.0016167, 0.000003, 0.0000335, 0.000003, 0.0000619}},
             &small_rotor_Y, {0.0, 0.0, -0.1455}, {0., -0.0325, -0.06}, CoordinateAxis::Y, 9.0}};

        for (int arm = 0; arm < 2; arm++)
        {
            const std::string s = arm == 0 ? "left-" : "right-";
            auto mirror = [&](const Vec3 &v) { return Vec3{v[0], arm == 0 ? v[1] : -v[1], v[2]}; };
            auto mirrorI = [&](const SpatialInertia &I)
            { return arm == 0 ? I : I.flipAlongAxis(CoordinateAxis::Y); };
            for (const ArmRow &r : rows)
            {
                const std::string parent = std::string(r.parent) == "torso" ? "torso" : s + r.parent;
                appendGeared(model, s + r.cluster, s + r.link, s + r.rotor, parent,
                             mirrorI(inertiaOf(r.inertial)),
                             mirrorI(SpatialInertia(0., {0., 0., 0.}, *r.rotor_inertia)),
                             spatial::Transform(I3, mirror(r.location)),
                             spatial::Transform(I3, mirror(r.rotor_location)), r.axis, r.gear_ratio);
            }
        }
        return model;
    }

    // -----------------------------------------------------------------------------------------
    // Uniform chains: m = 1, c = (0.5, 0, 0), Izz = 1, rotor Izz = 1e-4, l = 1, gear 2 x belt 3,
    // z axes, gravity (+9.81, 0, 0)
    // -----------------------------------------------------------------------------------------
    namespace
    {
        const SpatialInertia &chainLinkInertia()
        {
            static const SpatialInertia I(1., {0.5, 0., 0.}, {0., 0., 0., 0., 0., 0., 0., 0., 1.});
            return I;
        }
        const SpatialInertia &chainRotorInertia()
        {
            static const SpatialInertia I(0., {0., 0., 0.}, {0., 0., 0., 0., 0., 0., 0., 0., 1e-4});
            return I;
        }
    } // namespace

    ClusterTreeModel RevoluteChainWithRotor::buildClusterTreeModel() const
    {
        ClusterTreeModel model;
        model.setGravity({9.81, 0., 0.});
        std::string prev = "ground";
        for (int i = 0; i < N_; i++)
        {
            const spatial::Transform X(I3, {i == 0 ? 0. : 1., 0., 0.});
            const std::string k = std::to_string(i);
            appendGeared(model, "cluster-" + k, "link-" + k, "rotor-" + k, prev, chainLinkInertia(),
                         chainRotorInertia(), X, X, CoordinateAxis::Z, 2. * 3.);
            prev = "link-" + k;
        }
        return model;
    }

    ClusterTreeModel RevolutePairChainWithRotor::buildClusterTreeModel() const
    {
        ClusterTreeModel model;
        model.setGravity({9.81, 0., 0.});
        const spatial::Transform X2(I3, {1., 0., 0.});
        std::string parent = "ground";
        for (int i = 0; i < N_ / 2; i++)
        {
            const spatial::Transform X1 = i == 0 ? spatial::Transform(I3, {0., 0., 0.}) : X2;
            const std::string k = std::to_string(i);
            Body linkA = model.registerBody("link-A-" + k, chainLinkInertia(), parent, X1);
            Body rotorA = model.registerBody("rotor-A-" + k, chainRotorInertia(), parent, X1);
            Body rotorB = model.registerBody("rotor-B-" + k, chainRotorInertia(), parent, X1);
            Body linkB = model.registerBody("link-B-" + k, chainLinkInertia(), "link-A-" + k, X2);
            ParallelBeltTransmissionModule mA{linkA, rotorA, CoordinateAxis::Z, CoordinateAxis::Z, 2., {3.}};
            ParallelBeltTransmissionModule mB{linkB, rotorB, CoordinateAxis::Z, CoordinateAxis::Z, 2., {3., 1.}};
            model.appendRegisteredBodiesAsCluster<RevolutePairWithRotor>("cluster-" + k, mA, mB);
            parent = "link-B-" + k;
        }
        return model;
    }

    ClusterTreeModel RevolutePairChain::buildClusterTreeModel() const
    {
        ClusterTreeModel model;
        model.setGravity({9.81, 0., 0.});
        const spatial::Transform X2(I3, {1., 0., 0.});
        std::string parent = "ground";
        for (int i = 0; i < N_ / 2; i++)
        {
            const spatial::Transform X1 = i == 0 ? spatial::Transform(I3, {0., 0., 0.}) : X2;
            const std::string k = std::to_string(i);
            Body linkA = model.registerBody("link-A-" + k, chainLinkInertia(), parent, X1);
            Body linkB = model.registerBody("link-B-" + k, chainLinkInertia(), "link-A-" + k, X2);
            model.appendRegisteredBodiesAsCluster<RevolutePair>("cluster-" + k, linkA, linkB, CoordinateAxis::Z,
                                                                CoordinateAxis::Z);
            parent = "link-B-" + k;
        }
        return model;
    }

    namespace
    {
        // splitmix64: seeded stand-in for the reference's rand() / Eigen Random() parameter draws
        struct ParamRng
        {
            uint64_t s;
            uint64_t next()
            {
                uint64_t z = (s += 0x9e3779b97f4a7c15ull);
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
                return z ^ (z >> 31);
            }
            double unit() { return (double)(next() >> 11) / 9007199254740992.0; } // [0, 1)
            double sym() { return 2.0 * unit() - 1.0; }                           // [-1, 1)
            int upTo(int n) { return (int)(next() % (uint64_t)n); }
            CoordinateAxis axis() { return (CoordinateAxis)upTo(3); }
            // spatial::randomSpatialRotation (include/grbda/Utils/Spatial.h:43-48): r and rpy uniform in [-1, 1]
            spatial::Transform transform()
            {
                const Mat3 E = ori::rpyToRotMat({sym(), sym(), sym()});
                return spatial::Transform(E, {sym(), sym(), sym()});
            }
            // SpatialInertia::createRandomInertia (SpatialInertia.h:144-151)
            SpatialInertia inertia(double scaling)
            {
                const double mass = scaling * unit();
                const Vec3 com = {scaling * sym(), scaling * sym(), scaling * sym()};
                Mat3 A;
                for (double &x : A)
                    x = sym();
                return SpatialInertia(mass, com, ori::scale(ori::mul(A, ori::transpose(A)), scaling));
            }
        };
    } // namespace

    ClusterTreeModel RevoluteTripleChainWithRotor::buildClusterTreeModel() const
    {
        ClusterTreeModel model;
        ParamRng rng{seed_};
        std::string parent = "ground";
        for (int i = 0; i < N_ / 3; i++)
        {
            const std::string k = std::to_string(i);
            const char *tag[3] = {"A", "B", "C"};
            Body link[3], rotor[3];
            CoordinateAxis link_axis[3], rotor_axis[3];
            std::string prev = parent;
            for (int j = 0; j < 3; j++)
            {
                const std::string name = std::string("link-") + tag[j] + "-" + k;
                const spatial::Transform X = rng.transform();
                const SpatialInertia I = rng.inertia(1.0);
                link_axis[j] = rng.axis();
                link[j] = model.registerBody(name, I, prev, X);
                prev = name;
            }
            for (int j = 0; j < 3; j++)
            {
                const spatial::Transform X = rng.transform();
                const SpatialInertia I = rng.inertia(1e-4);
                rotor_axis[j] = rng.axis();
                rotor[j] = model.registerBody(std::string("rotor-") + tag[j] + "-" + k, I, parent, X);
            }
            ParallelBeltTransmissionModule m[3];
            for (int j = 0; j < 3; j++)
            {
                std::vector<double> belts;
                const double gear = (double)(rng.upTo(5) + 1);
                for (int b = 0; b <= j; b++)
                    belts.push_back((double)(rng.upTo(5) + 1));
                m[j] = ParallelBeltTransmissionModule{link[j], rotor[j], link_axis[j], rotor_axis[j], gear, belts};
            }
            model.appendRegisteredBodiesAsCluster<RevoluteTripleWithRotor>("cluster-" + k, m[0], m[1], m[2]);
            parent = prev;
        }
        return model;
    }

} // namespace grbda
