// URDF+ front end: XML -> links / joints / constraints -> clusters -> ClusterTreeModel.
//
// Replaces, for the hot path's model construction:
//   * the mit-biomimetics/urdfdom fork (un-vendored dependency of the reference, pinned at
//     0425417a68c331ae308a86fc752474ee55d7b041 by scripts/install_dependencies.sh:52-70): XML
//     parsing, the `independent` joint attribute, <loop> / <coupling> constraints, nearest common
//     ancestors, clustering by strongly connected components and the child-cluster order. Its
//     behaviour is restated from the reference's own tests: UnitTests/testUrdfParser.cpp:92-396
//     (children ordered by joint name; loop links predecessor -> successor side and successor ->
//     predecessor side below the nearest common ancestor; clusters = SCCs) and
//     UnitTests/testClusterTreeModel.cpp:100-230 (cluster / body order must reproduce the manual
//     builders for mini_cheetah, mit_humanoid, revolute_rotor_chain, planar_leg_linkage);
//   * src/Dynamics/ClusterTreeParsing.cpp:5-440: DFS over clusters, single-link clusters ->
//     Revolute / Free, multi-link clusters -> Generic with a GenericImplicit (<loop>) or Static
//     (<coupling>) constraint. The casadi::SX constraint function becomes a recorded sym::Sym
//     program; casadi's which_depends becomes "the recorded row is not a constant".
#include <cctype>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include "../compiler/spatial_sym.h"
#include "model.h"

namespace grbda
{
    namespace
    {
        // ------------------------------------------------------------------------------------
        // minimal XML reader (elements, attributes, comments, declarations; no entities needed)
        // ------------------------------------------------------------------------------------
        struct XmlNode
        {
            std::string name;
            std::map<std::string, std::string> attr;
            std::vector<XmlNode> children;
            const XmlNode *child(const std::string &n) const
            {
                for (const XmlNode &c : children)
                    if (c.name == n)
                        return &c;
                return nullptr;
            }
            std::string get(const std::string &k, const std::string &def = "") const
            {
                auto it = attr.find(k);
                return it == attr.end() ? def : it->second;
            }
        };

        class XmlParser
        {
        public:
            explicit XmlParser(const std::string &text) : s_(text) {}
            XmlNode parseDocument()
            {
                XmlNode root;
                root.name = "#document";
                while (skipMisc())
                    root.children.push_back(parseElement());
                return root;
            }

        private:
            // skip whitespace, text, comments, declarations; true when an element starts at pos_
            bool skipMisc()
            {
                while (pos_ < s_.size())
                {
                    if (s_[pos_] != '<')
                    {
                        pos_++;
                        continue;
                    }
                    if (s_.compare(pos_, 4, "<!--") == 0)
                    {
                        const size_t e = s_.find("-->", pos_ + 4);
                        if (e == std::string::npos)
                            throw std::runtime_error("URDF: unterminated comment");
                        pos_ = e + 3;
                    }
                    else if (s_.compare(pos_, 2, "<?") == 0 || s_.compare(pos_, 2, "<!") == 0)
                    {
                        const size_t e = s_.find('>', pos_);
                        if (e == std::string::npos)
                            throw std::runtime_error("URDF: unterminated declaration");
                        pos_ = e + 1;
                    }
                    else if (s_.compare(pos_, 2, "</") == 0)
                        return false;
                    else
                        return true;
                }
                return false;
            }
            XmlNode parseElement()
            {
                XmlNode n;
                pos_++; // '<'
                while (pos_ < s_.size() && !isspace((unsigned char)s_[pos_]) && s_[pos_] != '>' && s_[pos_] != '/')
                    n.name += s_[pos_++];
                for (;;)
                {
                    while (pos_ < s_.size() && isspace((unsigned char)s_[pos_]))
                        pos_++;
                    if (pos_ >= s_.size())
                        throw std::runtime_error("URDF: unterminated tag <" + n.name);
                    if (s_[pos_] == '/')
                    {
                        pos_ += 2; // "/>"
                        return n;
                    }
                    if (s_[pos_] == '>')
                    {
                        pos_++;
                        break;
                    }
                    std::string key;
                    while (pos_ < s_.size() && s_[pos_] != '=' && !isspace((unsigned char)s_[pos_]))
                        key += s_[pos_++];
                    while (pos_ < s_.size() && (isspace((unsigned char)s_[pos_]) || s_[pos_] == '='))
                        pos_++;
                    const char quote = s_[pos_++];
                    if (quote != '"' && quote != '\'')
                        throw std::runtime_error("URDF: attribute value of '" + key + "' is not quoted");
                    std::string val;
                    while (pos_ < s_.size() && s_[pos_] != quote)
                        val += s_[pos_++];
                    pos_++;
                    n.attr[key] = val;
                }
                while (skipMisc())
                    n.children.push_back(parseElement());
                // closing tag
                const size_t e = s_.find('>', pos_);
                if (e == std::string::npos)
                    throw std::runtime_error("URDF: missing closing tag of <" + n.name + ">");
                pos_ = e + 1;
                return n;
            }
            const std::string &s_;
            size_t pos_ = 0;
        };

        std::vector<double> numbers(const std::string &s, size_t expected, const std::vector<double> &def)
        {
            if (s.empty())
                return def;
            std::istringstream is(s);
            std::vector<double> v;
            double x;
            while (is >> x)
                v.push_back(x);
            if (v.size() != expected)
                throw std::runtime_error("URDF: expected " + std::to_string(expected) + " numbers in '" + s + "'");
            return v;
        }

        // ------------------------------------------------------------------------------------
        // the urdf::ModelInterface subset the reference consumes
        // ------------------------------------------------------------------------------------
        struct Pose
        {
            Vec3 xyz{0, 0, 0}, rpy{0, 0, 0};
            // spatial::Transform(urdf::Pose) (src/Utils/SpatialTransforms.cpp:17-23): E is the
            // coordinate transformation of the pose, r its position
            spatial::Transform toTransform() const { return spatial::Transform(ori::rpyToRotMat(rpy), xyz); }
        };
        Pose parsePose(const XmlNode *origin)
        {
            Pose p;
            if (!origin)
                return p;
            const auto xyz = numbers(origin->get("xyz"), 3, {0, 0, 0}), rpy = numbers(origin->get("rpy"), 3, {0, 0, 0});
            p.xyz = {xyz[0], xyz[1], xyz[2]};
            p.rpy = {rpy[0], rpy[1], rpy[2]};
            return p;
        }

        struct UJoint
        {
            std::string name, type, parent, child;
            bool independent = true;
            Pose origin;
            Vec3 axis{1, 0, 0};
        };
        struct UConstraint
        {
            std::string name;
            bool is_loop = false;
            std::string predecessor, successor, nca;
            Pose pred_origin, succ_origin;
            double ratio = 1.0;
            std::vector<std::string> nca_to_pred, nca_to_succ; // link names below the NCA
        };
        struct ULink
        {
            std::string name;
            bool has_inertial = false;
            double mass = 0;
            Vec3 com{0, 0, 0};
            Mat3 inertia{};
            std::string parent; // empty for the root
            const UJoint *parent_joint = nullptr;
            std::vector<std::string> child_links; // ordered by joint name (urdfdom joints_ map)
            std::vector<std::string> loop_links;
            std::vector<int> constraints; // indices into UModel::constraints
        };
        struct UCluster
        {
            std::vector<std::string> links;
            int parent = -1;
            std::vector<int> children;
        };
        struct UModel
        {
            std::map<std::string, ULink> links;
            std::map<std::string, UJoint> joints;
            std::vector<UConstraint> constraints;
            std::string root;
            std::vector<UCluster> clusters;
            std::map<std::string, int> containing_cluster;
        };

        // ori::urdfAxisToCoordinateAxis (OrientationTools.h:70-93): unit coordinate axes only, sign dropped
        ori::CoordinateAxis coordinateAxis(const Vec3 &a)
        {
            if (std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) != 1)
                throw std::runtime_error("Error: Joint axis must be a unit vector");
            if (a[0] == 1 || a[0] == -1)
                return ori::CoordinateAxis::X;
            if (a[1] == 1 || a[1] == -1)
                return ori::CoordinateAxis::Y;
            if (a[2] == 1 || a[2] == -1)
                return ori::CoordinateAxis::Z;
            throw std::runtime_error("Error: Joint axis not defined");
        }

        std::vector<std::string> chainToRoot(const UModel &m, const std::string &link)
        {
            std::vector<std::string> chain; // link, parent, ..., root
            std::string n = link;
            while (!n.empty())
            {
                chain.push_back(n);
                n = m.links.at(n).parent;
            }
            return chain;
        }

        // The <link> / <joint> / <loop> / <coupling> elements of one file, added to `m`. Several files make one model
        // (the fork's parseURDFFiles, ClusterTreeModel.h:48-53; UnitTests/testUrdfParser.cpp:398-445 splits Mini
        // Cheetah into a base and four legs): every file names the links it attaches to as empty <link/> stubs, so
        // a link may be declared more than once as long as at most one declaration carries an <inertial>; joints
        // and constraints must be unique.
        void readRobotElements(const std::string &path, UModel &m)
        {
            std::ifstream f(path);
            if (!f)
                throw std::runtime_error("Could not parse URDF file: cannot open " + path);
            std::stringstream ss;
            ss << f.rdbuf();
            const std::string text = ss.str();
            const XmlNode doc = XmlParser(text).parseDocument();
            const XmlNode *robot = doc.child("robot");
            if (!robot)
                throw std::runtime_error("Could not parse URDF file: no <robot> element in " + path);

            for (const XmlNode &n : robot->children)
            {
                if (n.name == "link")
                {
                    ULink l;
                    l.name = n.get("name");
                    if (const XmlNode *in = n.child("inertial"))
                    {
                        l.has_inertial = true;
                        if (const XmlNode *ms = in->child("mass"))
                            l.mass = numbers(ms->get("value"), 1, {0})[0];
                        l.com = parsePose(in->child("origin")).xyz; // inertial rpy ignored (SpatialInertia.h:105-116)
                        if (const XmlNode *it = in->child("inertia"))
                        {
                            auto g = [&](const char *k) { return numbers(it->get(k), 1, {0})[0]; };
                            l.inertia = {g("ixx"), g("ixy"), g("ixz"), g("ixy"), g("iyy"), g("iyz"),
                                         g("ixz"), g("iyz"), g("izz")};
                        }
                    }
                    const auto seen = m.links.find(l.name);
                    if (seen == m.links.end())
                        m.links[l.name] = l;
                    else if (!l.has_inertial)
                        ; // a stub for a link another declaration defines
                    else if (!seen->second.has_inertial)
                        seen->second = l;
                    else
                        throw std::runtime_error("URDF: link '" + l.name + "' is not unique");
                }
                else if (n.name == "joint")
                {
                    UJoint j;
                    j.name = n.get("name");
                    j.type = n.get("type");
                    {
                        // "true"/"false" in the hand-written models, "True"/"False" where xacro evaluated a
                        // Python expression (Benchmarking/urdfs/parallel_chains)
                        std::string flag = n.get("independent", "true");
                        for (char &ch : flag)
                            ch = (char)std::tolower((unsigned char)ch);
                        j.independent = !(flag == "false" || flag == "0");
                    }
                    if (!n.child("parent") || !n.child("child"))
                        throw std::runtime_error("URDF: joint '" + j.name + "' needs <parent> and <child>");
                    j.parent = n.child("parent")->get("link");
                    j.child = n.child("child")->get("link");
                    j.origin = parsePose(n.child("origin"));
                    if (const XmlNode *ax = n.child("axis"))
                    {
                        const auto a = numbers(ax->get("xyz"), 3, {1, 0, 0});
                        j.axis = {a[0], a[1], a[2]};
                    }
                    if (m.joints.count(j.name))
                        throw std::runtime_error("URDF: joint '" + j.name + "' is not unique");
                    m.joints[j.name] = j;
                }
                else if (n.name == "loop" || n.name == "coupling")
                {
                    UConstraint c;
                    c.name = n.get("name");
                    c.is_loop = n.name == "loop";
                    const XmlNode *p = n.child("predecessor"), *s = n.child("successor");
                    if (!p || !s)
                        throw std::runtime_error("URDF: constraint '" + c.name + "' needs predecessor and successor");
                    c.predecessor = p->get("link");
                    c.successor = s->get("link");
                    c.pred_origin = parsePose(p->child("origin"));
                    c.succ_origin = parsePose(s->child("origin"));
                    if (const XmlNode *r = n.child("ratio"))
                        c.ratio = numbers(r->get("value"), 1, {1})[0];
                    for (const UConstraint &other : m.constraints)
                        if (other.name == c.name)
                            throw std::runtime_error("URDF: constraint '" + c.name + "' is not unique");
                    m.constraints.push_back(c);
                }
            }
        }

        UModel parseUrdf(const std::vector<std::string> &paths)
        {
            if (paths.empty())
                throw std::runtime_error("Could not parse URDF file: no file given");
            UModel m;
            for (const std::string &path : paths)
                readRobotElements(path, m);
            // tree (urdfdom initTree: joints_ is a name-keyed map, so children follow joint-name order)
            for (auto &kv : m.joints)
            {
                UJoint &j = kv.second;
                if (!m.links.count(j.parent) || !m.links.count(j.child))
                    throw std::runtime_error("URDF: joint '" + j.name + "' refers to an unknown link");
                ULink &child = m.links[j.child];
                if (!child.parent.empty())
                    throw std::runtime_error("URDF: link '" + j.child + "' has two parent joints");
                child.parent = j.parent;
                child.parent_joint = &j;
                m.links[j.parent].child_links.push_back(j.child);
            }
            for (auto &kv : m.links)
                if (kv.second.parent.empty())
                {
                    if (!m.root.empty())
                        throw std::runtime_error("URDF: two root links found: " + m.root + " and " + kv.first);
                    m.root = kv.first;
                }
            if (m.root.empty())
                throw std::runtime_error("URDF: no root link");

            // constraints: nearest common ancestor, sub-chains, loop links
            for (size_t ci = 0; ci < m.constraints.size(); ci++)
            {
                UConstraint &c = m.constraints[ci];
                if (!m.links.count(c.predecessor) || !m.links.count(c.successor))
                    throw std::runtime_error("URDF: constraint '" + c.name + "' refers to an unknown link");
                const auto cp = chainToRoot(m, c.predecessor), cs = chainToRoot(m, c.successor);
                for (const std::string &a : cp)
                    if (std::find(cs.begin(), cs.end(), a) != cs.end())
                    {
                        c.nca = a;
                        break;
                    }
                for (const std::string &a : cp)
                {
                    if (a == c.nca)
                        break;
                    c.nca_to_pred.insert(c.nca_to_pred.begin(), a);
                }
                for (const std::string &a : cs)
                {
                    if (a == c.nca)
                        break;
                    c.nca_to_succ.insert(c.nca_to_succ.begin(), a);
                }
                if (c.nca_to_pred.empty() || c.nca_to_succ.empty())
                    throw std::runtime_error("URDF: constraint '" + c.name + "' connects a link to its own ancestor");
                // testUrdfParser.cpp:274-337: predecessor -> successor side, successor -> predecessor side
                m.links[c.predecessor].loop_links.push_back(c.nca_to_succ.front() == c.successor
                                                                ? c.successor
                                                                : c.nca_to_succ.front());
                m.links[c.successor].loop_links.push_back(c.nca_to_pred.front());
                m.links[c.predecessor].constraints.push_back((int)ci);
                m.links[c.successor].constraints.push_back((int)ci);
            }

            // clusters = strongly connected components (Tarjan) over child links then loop links;
            // components are listed in reverse completion order, which reproduces the cluster order
            // of the reference's manual builders (testClusterTreeModel.cpp:100-230)
            std::map<std::string, int> index, low;
            std::map<std::string, bool> on_stack;
            std::vector<std::string> stack;
            std::vector<std::vector<std::string>> sccs;
            int counter = 0;
            std::function<void(const std::string &)> strong = [&](const std::string &v)
            {
                index[v] = low[v] = counter++;
                stack.push_back(v);
                on_stack[v] = true;
                std::vector<std::string> nbrs = m.links[v].child_links;
                nbrs.insert(nbrs.end(), m.links[v].loop_links.begin(), m.links[v].loop_links.end());
                for (const std::string &w : nbrs)
                {
                    if (!index.count(w))
                    {
                        strong(w);
                        low[v] = std::min(low[v], low[w]);
                    }
                    else if (on_stack[w])
                        low[v] = std::min(low[v], index[w]);
                }
                if (low[v] == index[v])
                {
                    std::vector<std::string> comp;
                    for (;;)
                    {
                        const std::string w = stack.back();
                        stack.pop_back();
                        on_stack[w] = false;
                        comp.push_back(w);
                        if (w == v)
                            break;
                    }
                    sccs.push_back(comp);
                }
            };
            strong(m.root);
            std::reverse(sccs.begin(), sccs.end());
            for (auto &comp : sccs)
            {
                // links of a cluster are kept in link-name order (urdfdom links_ is a name-keyed map);
                // registration later defers a link until its parent is registered
                std::sort(comp.begin(), comp.end());
                UCluster c;
                c.links = comp;
                for (const std::string &l : comp)
                    m.containing_cluster[l] = (int)m.clusters.size();
                m.clusters.push_back(c);
            }
            if (m.containing_cluster.size() != m.links.size())
                throw std::runtime_error("URDF: some links are not connected to the root link");
            for (size_t ci = 0; ci < m.clusters.size(); ci++)
                for (const std::string &l : m.clusters[ci].links)
                {
                    const std::string &p = m.links[l].parent;
                    if (p.empty())
                        continue;
                    const int pc = m.containing_cluster[p];
                    if (pc == (int)ci)
                        continue;
                    if (m.clusters[ci].parent >= 0 && m.clusters[ci].parent != pc)
                        throw std::runtime_error("The parents of all bodies in a cluster must have parents in the "
                                                 "current cluster OR in the same parent cluster");
                    if (m.clusters[ci].parent < 0)
                    {
                        m.clusters[ci].parent = pc;
                        m.clusters[pc].children.push_back((int)ci);
                    }
                }
            return m;
        }

        SpatialInertia linkInertia(const ULink &l)
        {
            // SpatialInertia(urdf::Inertial) (SpatialInertia.h:105-116)
            return SpatialInertia(l.mass, l.com, l.inertia);
        }

        // dense solve A X = B (row-major, partial pivoting) for the constant coupling Jacobians
        std::vector<double> solveDense(std::vector<double> A, std::vector<double> B, int n, int nrhs)
        {
            for (int k = 0; k < n; k++)
            {
                int p = k;
                for (int i = k + 1; i < n; i++)
                    if (std::fabs(A[i * n + k]) > std::fabs(A[p * n + k]))
                        p = i;
                if (A[p * n + k] == 0.0)
                    throw std::runtime_error("URDF: coupling constraints are singular in the dependent coordinates");
                for (int j = 0; j < n; j++)
                    std::swap(A[k * n + j], A[p * n + j]);
                for (int j = 0; j < nrhs; j++)
                    std::swap(B[k * nrhs + j], B[p * nrhs + j]);
                for (int i = k + 1; i < n; i++)
                {
                    const double f = A[i * n + k] / A[k * n + k];
                    for (int j = k; j < n; j++)
                        A[i * n + j] -= f * A[k * n + j];
                    for (int j = 0; j < nrhs; j++)
                        B[i * nrhs + j] -= f * B[k * nrhs + j];
                }
            }
            for (int i = n - 1; i >= 0; i--)
                for (int j = 0; j < nrhs; j++)
                {
                    double s = B[i * nrhs + j];
                    for (int k = i + 1; k < n; k++)
                        s -= A[i * n + k] * B[k * nrhs + j];
                    B[i * nrhs + j] = s / A[i * n + i];
                }
            return B;
        }
    } // namespace

    // What the front end made of the files, as JSON: the quantities the reference's parser tests pin
    // (UnitTests/testUrdfParser.cpp:39-396: link order, parents, children in order, supporting chains, neighbours =
    // children followed by loop links, clusters with parent / child clusters).
    std::string ClusterTreeModel::describeURDF(const std::vector<std::string> &urdf_filenames)
    {
        const UModel um = parseUrdf(urdf_filenames);
        auto quote = [](const std::string &x) { return "\"" + x + "\""; };
        auto list = [&](const std::vector<std::string> &v)
        {
            std::string o = "[";
            for (size_t i = 0; i < v.size(); i++)
                o += (i ? ", " : "") + quote(v[i]);
            return o + "]";
        };
        std::ostringstream os;
        // link order = the order buildModelFromURDF registers the bodies in: clusters depth first, child clusters in
        // their stored order, the links of a cluster in theirs
        std::vector<std::string> order;
        std::function<void(int)> visit = [&](int ci)
        {
            for (const std::string &l : um.clusters[ci].links)
                order.push_back(l);
            for (int ch : um.clusters[ci].children)
                visit(ch);
        };
        visit(um.containing_cluster.at(um.root));
        os << "{\"root\": " << quote(um.root) << ", \"link_order\": " << list(order) << ", \"links\": {";
        bool first = true;
        for (const auto &kv : um.links)
        {
            const ULink &l = kv.second;
            std::vector<std::string> chain = chainToRoot(um, l.name); // link ... root
            chain.pop_back();
            std::reverse(chain.begin(), chain.end());
            os << (first ? "" : ", ") << quote(l.name) << ": {\"parent\": " << (l.parent.empty() ? "null" : quote(l.parent))
               << ", \"children\": " << list(l.child_links) << ", \"loop_links\": " << list(l.loop_links)
               << ", \"supporting_chain\": " << list(chain) << ", \"cluster\": " << um.containing_cluster.at(l.name) << "}";
            first = false;
        }
        os << "}, \"clusters\": [";
        for (size_t c = 0; c < um.clusters.size(); c++)
        {
            const UCluster &uc = um.clusters[c];
            os << (c ? ", " : "") << "{\"links\": " << list(uc.links) << ", \"parent\": " << uc.parent << ", \"children\": [";
            for (size_t i = 0; i < uc.children.size(); i++)
                os << (i ? ", " : "") << uc.children[i];
            os << "]}";
        }
        os << "], \"num_joints\": " << um.joints.size() << ", \"num_constraints\": " << um.constraints.size() << "}";
        return os.str();
    }

    // reference: ClusterTreeModel.h:41-46 + ClusterTreeParsing.cpp:5-43
    void ClusterTreeModel::buildModelFromURDF(const std::string &urdf_filename)
    {
        buildModelFromURDF(std::vector<std::string>{urdf_filename});
    }
    // reference: ClusterTreeModel.h:48-53 (several files, one model)
    void ClusterTreeModel::buildModelFromURDF(const std::vector<std::string> &urdf_filenames)
    {
        using namespace ClusterJoints;
        const UModel um = parseUrdf(urdf_filenames);
        // the root link is the ground (ClusterTreeParsing.cpp:13-19)
        body_name_to_body_index_[um.root] = -1;
        const UCluster &root_cluster = um.clusters[um.containing_cluster.at(um.root)];
        if (root_cluster.links.size() != 1)
            throw std::runtime_error("The root cluster may only contain one body");

        std::function<void(int)> appendCluster = [&](int ci)
        {
            const UCluster &uc = um.clusters[ci];
            if (uc.links.size() == 1)
            {
                // ClusterTreeParsing.cpp:56-76, 232-258
                const ULink &l = um.links.at(uc.links[0]);
                const UJoint &j = *l.parent_joint;
                const spatial::Transform xtree = j.origin.toTransform();
                if (j.type == "revolute" || j.type == "continuous")
                    appendBody<Revolute>(l.name, linkInertia(l), l.parent, xtree, coordinateAxis(j.axis), j.name);
                else if (j.type == "floating")
                {
                    if (!cluster_nodes_.empty())
                        throw std::runtime_error("Floating joint must be the first joint in the system");
                    appendBody<Free>(l.name, linkInertia(l), l.parent, xtree, true, j.name);
                }
                else
                    throw std::runtime_error("The only joint in a cluster with one link must be revolute or floating");
            }
            else
            {
                // registerBodiesInUrdfCluster (:260-307): a link waits until its parent is registered
                std::vector<Body> bodies;
                std::vector<ori::CoordinateAxis> axes;
                std::vector<bool> independent;
                std::map<std::string, int> sub_index;
                std::vector<std::string> pending = uc.links;
                while (!pending.empty())
                {
                    std::vector<std::string> next;
                    for (const std::string &name : pending)
                    {
                        const ULink &l = um.links.at(name);
                        if (!body_name_to_body_index_.count(l.parent))
                        {
                            next.push_back(name);
                            continue;
                        }
                        const UJoint &j = *l.parent_joint;
                        if (j.type != "revolute" && j.type != "continuous")
                            throw std::runtime_error("Joints inside a multi-link cluster must be revolute");
                        sub_index[name] = (int)bodies.size();
                        bodies.push_back(registerBody(name, linkInertia(l), l.parent, j.origin.toTransform()));
                        axes.push_back(coordinateAxis(j.axis));
                        independent.push_back(j.independent);
                    }
                    if (next.size() == pending.size())
                        throw std::runtime_error("URDF: cluster contains a link whose parent is never registered");
                    pending = next;
                }
                // constraints of the cluster, in the order of its links (:88-95)
                std::vector<int> cons;
                for (const std::string &name : uc.links)
                    for (int c : um.links.at(name).constraints)
                        if (std::find(cons.begin(), cons.end(), c) == cons.end())
                            cons.push_back(c);
                if (cons.empty())
                    throw std::runtime_error("Cluster must have at least one constraint");
                const bool loops = um.constraints[cons[0]].is_loop;
                for (int c : cons)
                    if (um.constraints[c].is_loop != loops)
                        throw std::runtime_error("All constraints in cluster must be of same class type");
                const int N = (int)bodies.size();
                const std::string cluster_name = "cluster-" + std::to_string(cluster_nodes_.size());
                auto subChain = [&](const std::vector<std::string> &names) {
                    std::vector<int> s;
                    for (const std::string &n : names)
                    {
                        auto it = sub_index.find(n);
                        if (it == sub_index.end())
                            throw std::runtime_error("URDF: constraint chain leaves its cluster at link '" + n + "'");
                        s.push_back(it->second);
                    }
                    return s;
                };
                if (loops)
                {
                    // implicitPositionConstraint (:310-376) over sym::Sym
                    struct Capture
                    {
                        std::vector<int> pred, succ;
                        spatial::Transform pred_X, succ_X;
                    };
                    std::vector<Capture> caps;
                    for (int c : cons)
                    {
                        const UConstraint &uc2 = um.constraints[c];
                        caps.push_back({subChain(uc2.nca_to_pred), subChain(uc2.nca_to_succ),
                                        uc2.pred_origin.toTransform(), uc2.succ_origin.toTransform()});
                    }
                    std::vector<spatial::Transform> xtree;
                    for (const Body &b : bodies)
                        xtree.push_back(b.Xtree_);
                    auto phi_all = [=](const std::vector<sym::Sym> &q) {
                        using namespace compiler;
                        std::vector<sym::Sym> rows;
                        auto toXf = [](const spatial::Transform &T) {
                            Xf X;
                            X.E = constM3(T.E);
                            X.r = constV3(T.r);
                            return X;
                        };
                        for (const Capture &cap : caps)
                        {
                            auto through = [&](const std::vector<int> &chain, const spatial::Transform &origin) {
                                Xf X = Xf::identity();
                                for (int sub : chain)
                                {
                                    Xf XJ = Xf::identity();
                                    XJ.E = coordinateRotation(axes[sub], sym::sin(q[sub]), sym::cos(q[sub]));
                                    X = XJ * toXf(xtree[sub]) * X;
                                }
                                X = toXf(origin) * X;
                                return X.r;
                            };
                            const V3 rp = through(cap.pred, cap.pred_X), rs = through(cap.succ, cap.succ_X);
                            for (int k = 0; k < 3; k++)
                                rows.push_back(rp[k] - rs[k]);
                        }
                        return rows;
                    };
                    // keep the rows that depend on the joint coordinates (:358-371, casadi which_depends)
                    const PhiProgram all = PhiProgram::record(N, phi_all);
                    std::vector<int> keep;
                    for (size_t r = 0; r < all.outputs.size(); r++)
                        if (all.ops[all.outputs[r]].op != sym::OP_CONST)
                            keep.push_back((int)r);
                    auto phi = [=](const std::vector<sym::Sym> &q) {
                        const std::vector<sym::Sym> rows = phi_all(q);
                        std::vector<sym::Sym> out;
                        for (int r : keep)
                            out.push_back(rows[r]);
                        return out;
                    };
                    LoopConstraint::GenericImplicit lc(independent, phi);
                    appendRegisteredBodiesAsCluster<Generic>(cluster_name, bodies, axes, lc);
                }
                else
                {
                    // explicitRollingConstraint (:378-440)
                    const int nc = (int)cons.size();
                    std::vector<double> K(nc * N, 0.0);
                    for (int i = 0; i < nc; i++)
                    {
                        const UConstraint &uc2 = um.constraints[cons[i]];
                        for (int s : subChain(uc2.nca_to_pred))
                            K[i * N + s] = uc2.ratio;
                        for (int s : subChain(uc2.nca_to_succ))
                            K[i * N + s] = -1.0;
                    }
                    std::vector<int> ind, dep;
                    for (int i = 0; i < N; i++)
                        (independent[i] ? ind : dep).push_back(i);
                    const int n = (int)ind.size();
                    if ((int)dep.size() != nc)
                        throw std::runtime_error("URDF: a coupling cluster needs one constraint per dependent joint");
                    std::vector<double> Kd(nc * nc), Ki(nc * n);
                    for (int i = 0; i < nc; i++)
                    {
                        for (int j = 0; j < nc; j++)
                            Kd[i * nc + j] = K[i * N + dep[j]];
                        for (int j = 0; j < n; j++)
                            Ki[i * n + j] = K[i * N + ind[j]];
                    }
                    const std::vector<double> X = solveDense(Kd, Ki, nc, n); // Kd^-1 Ki
                    std::vector<double> G(N * n, 0.0);
                    for (int j = 0; j < n; j++)
                        G[ind[j] * n + j] = 1.0;
                    for (int i = 0; i < nc; i++)
                        for (int j = 0; j < n; j++)
                            G[dep[i] * n + j] = -X[i * n + j];
                    LoopConstraint::Static lc(G, N, n, K, nc);
                    appendRegisteredBodiesAsCluster<Generic>(cluster_name, bodies, axes, lc);
                }
            }
            for (int ch : uc.children)
                appendCluster(ch);
        };
        for (int ch : root_cluster.children)
            appendCluster(ch);
    }

} // namespace grbda
