"""URDF+ front end against the reference's parser tests, UnitTests/testUrdfParser.cpp, restated.

The reference checks the mit-biomimetics/urdfdom fork's ModelInterface (an un-vendored dependency): parents (:86-135),
children as sets (:137-201), supporting chains (:203-266), neighbours = children followed by loop links, order
significant (:268-337), cluster parent / child consistency over seven files (:339-396) and the union of several files
into one model (:398-445). The golden tables below are that file's tables, written as data. (Its `GetLinkOrders`
table, :39-84, is used by no test of the reference and is not pinned here either: the body order that matters is the
one of the manual builders, tests/test_compiler_cpu.py / tests/test_reference_builders.py.)
No GPU: grbda_cuda_describe_urdf and host-only model handles."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = os.path.join(ROOT, "generalized_rbda_b200", "robot-models")
CORPUS = os.path.join(ROOT, "tests", "urdf_corpus")


def urdf(name):
    for d in (MODELS, CORPUS):
        p = os.path.join(d, name + ".urdf")
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)


# testUrdfParser.cpp:14-25
TEST_URDF_FILES = ["mini_cheetah", "mini_cheetah_leg", "four_bar", "six_bar", "planar_leg_linkage",
                   "revolute_rotor_chain", "mit_humanoid_leg"]

PARENTS = {  # :92-113
    "four_bar": {"link1": "base_link", "link2": "link1", "link3": "base_link"},
    "mini_cheetah_leg": {"abduct": "base", "abduct_rotor": "base", "thigh": "abduct", "hip_rotor": "abduct",
                         "shank": "thigh", "knee_rotor": "thigh"},
}
CHILDREN = {  # :143-166
    "four_bar": {"base_link": ["link1", "link3"], "link1": ["link2"], "link2": [], "link3": []},
    "mini_cheetah_leg": {"base": ["abduct", "abduct_rotor"], "abduct": ["thigh", "hip_rotor"], "abduct_rotor": [],
                         "thigh": ["shank", "knee_rotor"], "hip_rotor": [], "shank": [], "knee_rotor": []},
}
SUPPORTING_CHAINS = {  # :209-229
    "four_bar": {"link1": ["link1"], "link2": ["link1", "link2"], "link3": ["link3"]},
    "mini_cheetah_leg": {"abduct": ["abduct"], "abduct_rotor": ["abduct_rotor"], "thigh": ["abduct", "thigh"],
                         "hip_rotor": ["abduct", "hip_rotor"], "shank": ["abduct", "thigh", "shank"],
                         "knee_rotor": ["abduct", "thigh", "knee_rotor"]},
}
NEIGHBORS = {  # :274-296, order significant
    "four_bar": {"base_link": ["link1", "link3"], "link1": ["link2"], "link2": ["link3"], "link3": ["link1"]},
    "mini_cheetah_leg": {"base": ["abduct", "abduct_rotor"], "abduct": ["hip_rotor", "thigh", "abduct_rotor"],
                         "abduct_rotor": ["abduct"], "thigh": ["shank", "knee_rotor", "hip_rotor"],
                         "hip_rotor": ["thigh"], "shank": ["knee_rotor"], "knee_rotor": ["shank"]},
}


@pytest.mark.parametrize("name", TEST_URDF_FILES)
def test_parses(grbda, name):
    d = grbda.describe_urdf(urdf(name))
    assert d["root"] and d["link_order"][0] == d["root"] and len(d["link_order"]) == len(d["links"])


@pytest.mark.parametrize("name", sorted(PARENTS))
def test_structure_goldens(grbda, name):
    links = grbda.describe_urdf(urdf(name))["links"]
    for link, parent in PARENTS[name].items():
        assert links[link]["parent"] == parent
    for link, children in CHILDREN[name].items():
        assert sorted(links[link]["children"]) == sorted(children)
    for link, chain in SUPPORTING_CHAINS[name].items():
        assert links[link]["supporting_chain"] == chain
    for link, neighbors in NEIGHBORS[name].items():
        assert links[link]["children"] + links[link]["loop_links"] == neighbors


@pytest.mark.parametrize("name", TEST_URDF_FILES)
def test_cluster_parents_and_children(grbda, name):
    d = grbda.describe_urdf(urdf(name))
    links, clusters = d["links"], d["clusters"]
    assert sorted(l for c in clusters for l in c["links"]) == sorted(links)
    for link, info in links.items():
        mine = info["cluster"]
        assert link in clusters[mine]["links"]
        if info["parent"] is not None:  # :351-373
            theirs = links[info["parent"]]["cluster"]
            assert theirs == mine or theirs == clusters[mine]["parent"]
        for child in info["children"]:  # :375-396
            c = links[child]["cluster"]
            assert c == mine or c in clusters[mine]["children"]


SPLIT = ["mini_cheetah_base", "mini_cheetah_fr_leg", "mini_cheetah_fl_leg", "mini_cheetah_hr_leg", "mini_cheetah_hl_leg"]


def test_combined_parse(grbda):
    """:398-445 - the five files describe the robot of mini_cheetah.urdf."""
    combined = grbda.describe_urdf([urdf(n) for n in SPLIT])
    whole = grbda.describe_urdf(urdf("mini_cheetah"))
    assert len(combined["links"]) == len(whole["links"]) == 26
    assert combined["num_joints"] == whole["num_joints"] and combined["num_constraints"] == whole["num_constraints"]
    assert combined["root"] == whole["root"]
    assert set(combined["links"]) == set(whole["links"])
    # beyond the reference's test: the same structure, and - built as models - the same model bit for bit
    assert combined["links"] == whole["links"] and combined["link_order"] == whole["link_order"]
    a = grbda.ClusterTreeModel.from_urdf([urdf(n) for n in SPLIT], device=None)
    b = grbda.ClusterTreeModel.from_urdf(urdf("mini_cheetah"), device=None)
    assert (a.nq, a.nv, a.nb, a.nc) == (b.nq, b.nv, b.nb, b.nc) == (19, 18, 25, 13)
    assert a.hash == b.hash
    for x, y in zip(a.bodies(), b.bodies()):
        assert x["name"] == y["name"] and x["parent"] == y["parent"] and x["cluster"] == y["cluster"]
        assert np.array_equal(x["E"], y["E"]) and np.array_equal(x["r"], y["r"]) and np.array_equal(x["inertia"], y["inertia"])
    # the order of the files does not matter
    c = grbda.ClusterTreeModel.from_urdf([urdf(n) for n in reversed(SPLIT)], device=None)
    assert c.hash == b.hash


def test_combined_parse_errors(grbda, tmp_path):
    leg = urdf("mini_cheetah_fr_leg")
    with pytest.raises(grbda.GrbdaError, match="not unique"):
        grbda.describe_urdf([urdf("mini_cheetah_base"), leg, leg])  # every joint twice
    with pytest.raises(grbda.GrbdaError, match="two root links|not connected"):
        grbda.describe_urdf([urdf("mini_cheetah_base"), urdf("four_bar")])
    with pytest.raises(grbda.GrbdaError, match="cannot open"):
        grbda.describe_urdf([urdf("mini_cheetah_base"), str(tmp_path / "missing.urdf")])
    # a link defined (with an <inertial>) in two files is an error, a stub next to a definition is not
    twice = tmp_path / "twice.urdf"
    twice.write_text('<robot name="x"><link name="Floating Base"><inertial><mass value="1"/>'
                     '<inertia ixx="1" ixy="0" ixz="0" iyy="1" iyz="0" izz="1"/></inertial></link></robot>')
    with pytest.raises(grbda.GrbdaError, match="link 'Floating Base' is not unique"):
        grbda.describe_urdf([urdf("mini_cheetah_base"), str(twice)])


@pytest.mark.gpu
def test_combined_model_on_gpu(grbda, oracle):
    """The model merged from five files finds the ahead-of-time kernels of mini_cheetah (same hash) and agrees with
    the oracle's manual MiniCheetah builder."""
    import torch
    m = grbda.ClusterTreeModel.from_urdf([urdf(n) for n in SPLIT])
    o = oracle.OracleModel("mini_cheetah")
    q, yd, tau = o.generate_states(64, seed=5)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    got = m.forwardDynamics(dev(q), dev(yd), dev(tau)).cpu().numpy()
    want = o.forward_dynamics(q, yd, tau)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-10
    got = m.inverseDynamics(dev(q), dev(yd), dev(tau)).cpu().numpy()
    want = o.inverse_dynamics(q, yd, tau)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-10
