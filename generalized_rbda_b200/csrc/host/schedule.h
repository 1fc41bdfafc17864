// ClusterTreeModel <-> flattened structure-of-arrays schedule (include/grbda_cuda.h: grbda_schedule)
// and the model hash that keys the ahead-of-time compiled kernel registry.
#pragma once
#include <cstring>
#include "../../../include/grbda_cuda.h"
#include "model.h"

namespace grbda
{
    // Owns the arrays a grbda_schedule points to.
    struct ScheduleStorage
    {
        std::vector<int32_t> body_parent, body_joint_axis;
        std::vector<double> body_xtree_E, body_xtree_r, body_inertia;
        std::vector<uint8_t> body_independent;
        std::vector<int32_t> cluster_type, cluster_num_bodies, cluster_num_independent, cluster_G_offset,
            cluster_phi_offset, cluster_phi_count, cluster_phi_out_offset, cluster_num_constraints;
        std::vector<double> G_values;
        std::vector<grbda_phi_op> phi_ops;
        std::vector<int32_t> phi_outputs;
        grbda_schedule view{};

        void finalize(const Vec3 &gravity)
        {
            view.num_bodies = (int32_t)body_parent.size();
            view.num_clusters = (int32_t)cluster_type.size();
            for (int i = 0; i < 3; i++)
                view.gravity[i] = gravity[i];
            view.body_parent = body_parent.data();
            view.body_joint_axis = body_joint_axis.data();
            view.body_xtree_E = body_xtree_E.data();
            view.body_xtree_r = body_xtree_r.data();
            view.body_inertia = body_inertia.data();
            view.body_independent = body_independent.data();
            view.cluster_type = cluster_type.data();
            view.cluster_num_bodies = cluster_num_bodies.data();
            view.cluster_num_independent = cluster_num_independent.data();
            view.cluster_G_offset = cluster_G_offset.data();
            view.G_values = G_values.data();
            view.cluster_phi_offset = cluster_phi_offset.data();
            view.cluster_phi_count = cluster_phi_count.data();
            view.cluster_phi_out_offset = cluster_phi_out_offset.data();
            view.cluster_num_constraints = cluster_num_constraints.data();
            view.phi_ops = phi_ops.data();
            view.phi_outputs = phi_outputs.data();
        }
    };

    inline void toSchedule(const ClusterTreeModel &m, ScheduleStorage &s)
    {
        s = ScheduleStorage();
        for (const ClusterTreeNode &c : m.clusters())
        {
            const ClusterDesc &d = c.joint_;
            s.cluster_type.push_back((int32_t)d.type);
            s.cluster_num_bodies.push_back(d.num_bodies);
            s.cluster_num_independent.push_back(d.num_velocities);
            s.cluster_num_constraints.push_back(d.num_constraints);
            s.cluster_G_offset.push_back((int32_t)s.G_values.size());
            if (d.type == ClusterType::Explicit)
                s.G_values.insert(s.G_values.end(), d.G.begin(), d.G.end());
            s.cluster_phi_offset.push_back((int32_t)s.phi_ops.size());
            s.cluster_phi_count.push_back((int32_t)d.phi.ops.size());
            s.cluster_phi_out_offset.push_back((int32_t)s.phi_outputs.size());
            for (const PhiOp &o : d.phi.ops)
                s.phi_ops.push_back(grbda_phi_op{o.op, o.a, o.b, o.val});
            s.phi_outputs.insert(s.phi_outputs.end(), d.phi.outputs.begin(), d.phi.outputs.end());
            for (int i = 0; i < d.num_bodies; i++)
            {
                const Body &b = c.bodies_[i];
                s.body_parent.push_back(b.parent_index_);
                s.body_joint_axis.push_back((int32_t)d.axes[i]);
                s.body_xtree_E.insert(s.body_xtree_E.end(), b.Xtree_.E.begin(), b.Xtree_.E.end());
                s.body_xtree_r.insert(s.body_xtree_r.end(), b.Xtree_.r.begin(), b.Xtree_.r.end());
                const Mat6 &I = b.inertia_.getMatrix();
                s.body_inertia.insert(s.body_inertia.end(), I.begin(), I.end());
                s.body_independent.push_back(d.type == ClusterType::Implicit ? (uint8_t)d.independent[i] : 1);
            }
        }
        s.finalize(m.getGravity());
    }

    // Rebuild a ClusterTreeModel from a schedule through the same registerBody / appendCluster path
    // (so every ClusterTreeModel rule is re-checked: parents registered first, one parent cluster ...).
    inline ClusterTreeModel fromSchedule(const grbda_schedule &s)
    {
        ClusterTreeModel m;
        m.setGravity({s.gravity[0], s.gravity[1], s.gravity[2]});
        int body = 0;
        auto bodyName = [](int i) { return i < 0 ? std::string("ground") : "body-" + std::to_string(i); };
        for (int c = 0; c < s.num_clusters; c++)
        {
            ClusterDesc d;
            const int type = s.cluster_type[c];
            if (type < 0 || type > 3)
                throw std::runtime_error("schedule: invalid cluster_type");
            d.type = (ClusterType)type;
            d.num_bodies = s.cluster_num_bodies[c];
            d.num_velocities = s.cluster_num_independent[c];
            d.num_constraints = s.cluster_num_constraints ? s.cluster_num_constraints[c] : 0;
            d.joint_type_name = "Schedule";
            for (int i = 0; i < d.num_bodies; i++, body++)
            {
                if (body >= s.num_bodies)
                    throw std::runtime_error("schedule: cluster body counts exceed num_bodies");
                Mat3 E;
                Vec3 r;
                Mat6 I;
                std::memcpy(E.data(), s.body_xtree_E + 9 * body, sizeof(E));
                std::memcpy(r.data(), s.body_xtree_r + 3 * body, sizeof(r));
                std::memcpy(I.data(), s.body_inertia + 36 * body, sizeof(I));
                if (s.body_parent[body] >= body)
                    throw std::runtime_error("schedule: a body's parent must be registered before it");
                m.registerBody(bodyName(body), SpatialInertia(I), bodyName(s.body_parent[body]),
                               spatial::Transform(E, r));
                const int axis = s.body_joint_axis[body];
                if (axis < 0 || axis > 2)
                    throw std::runtime_error("schedule: joint axis must be 0, 1 or 2");
                d.axes.push_back((ori::CoordinateAxis)axis);
                if (d.type == ClusterType::Implicit)
                    d.independent.push_back(s.body_independent[body] != 0);
            }
            switch (d.type)
            {
            case ClusterType::FreeQuaternion:
            case ClusterType::FreeRollPitchYaw:
                if (d.num_bodies != 1 || d.num_velocities != 6)
                    throw std::runtime_error("schedule: a free cluster has one body and six velocities");
                d.num_positions = d.type == ClusterType::FreeQuaternion ? 7 : 6;
                break;
            case ClusterType::Explicit:
                d.num_positions = d.num_velocities;
                d.G.assign(s.G_values + s.cluster_G_offset[c],
                           s.G_values + s.cluster_G_offset[c] + d.num_bodies * d.num_velocities);
                break;
            case ClusterType::Implicit:
            {
                d.num_positions = d.num_bodies;
                d.phi.num_inputs = d.num_bodies;
                for (int i = 0; i < s.cluster_phi_count[c]; i++)
                {
                    const grbda_phi_op &o = s.phi_ops[s.cluster_phi_offset[c] + i];
                    d.phi.ops.push_back(PhiOp{o.op, o.a, o.b, o.val});
                }
                for (int i = 0; i < d.num_constraints; i++)
                    d.phi.outputs.push_back(s.phi_outputs[s.cluster_phi_out_offset[c] + i]);
                int n_ind = 0;
                for (bool b : d.independent)
                    n_ind += b;
                if (n_ind != d.num_velocities || d.num_constraints != d.num_bodies - n_ind)
                    throw std::runtime_error("schedule: implicit cluster needs one phi row per dependent "
                                             "coordinate");
                break;
            }
            }
            m.appendCluster("cluster-" + std::to_string(c), d);
        }
        if (body != s.num_bodies)
            throw std::runtime_error("schedule: cluster body counts do not add up to num_bodies");
        return m;
    }

    // FNV-1a over the numerical content of the model (names excluded).
    inline uint64_t modelHash(const ClusterTreeModel &m)
    {
        ScheduleStorage s;
        toSchedule(m, s);
        uint64_t h = 1469598103934665603ull;
        auto mix = [&h](const void *p, size_t n)
        {
            const unsigned char *b = (const unsigned char *)p;
            for (size_t i = 0; i < n; i++)
            {
                h ^= b[i];
                h *= 1099511628211ull;
            }
        };
        auto mixv = [&](const auto &v) {
            if (!v.empty())
                mix(v.data(), v.size() * sizeof(v[0]));
        };
        mix(s.view.gravity, sizeof(s.view.gravity));
        mixv(s.body_parent), mixv(s.body_joint_axis), mixv(s.body_xtree_E), mixv(s.body_xtree_r);
        mixv(s.body_inertia), mixv(s.body_independent), mixv(s.cluster_type), mixv(s.cluster_num_bodies);
        mixv(s.cluster_num_independent), mixv(s.cluster_num_constraints), mixv(s.G_values);
        for (const grbda_phi_op &o : s.phi_ops)
        {
            mix(&o.op, 4), mix(&o.a, 4), mix(&o.b, 4), mix(&o.val, 8);
        }
        mixv(s.phi_outputs);
        return h;
    }

} // namespace grbda
