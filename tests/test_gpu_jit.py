"""GPU tests of the run-time compiler (runtime/jit.cpp): models that build.py did not compile ahead of time
get their kernels from NVRTC when an entry point is first used. The boundary under test is the one a
maintainer of the reference binds: grbda_cuda_model_create(schedule) /
grbda_cuda_model_create_from_urdf for ARBITRARY models
(reference: include/grbda/Dynamics/ClusterTreeModel.h:27-53, src/Dynamics/ClusterTreeParsing.cpp:5-43)."""
import os

import numpy as np
import pytest

from test_gpu_parity import TOL64, relrows

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CORPUS = os.path.join(HERE, "urdf_corpus")


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def check_against_oracle(m, o, B=1000, seed=7, fk=True):
    q, yd, aux, flags = m.generateStates(B, seed=seed)
    assert int(flags.sum()) == 0
    qn, ydn, auxn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    assert o.validate_states(qn).all()
    assert relrows(m.inverseDynamics(q, yd, aux).cpu().numpy(), o.inverse_dynamics(qn, ydn, auxn)) < TOL64
    assert relrows(m.forwardDynamics(q, yd, aux).cpu().numpy(), o.forward_dynamics(qn, ydn, auxn)) < TOL64
    assert relrows(m.getMassMatrix(q).cpu().numpy(), o.mass_matrix(qn)) < TOL64
    if fk:
        p, R, v = m.forwardKinematics(q, yd)
        po, Ro, vo = o.forward_kinematics(qn, ydn)
        assert relrows(p.cpu().numpy(), po) < TOL64 and relrows(R.cpu().numpy(), Ro) < TOL64
        assert relrows(v.cpu().numpy(), vo) < TOL64
    return q, yd, aux


@pytest.mark.parametrize("urdf", ["four_bar_branch_2_3.urdf", "revolute_rotor_pair_branch_2_3.urdf",
                                  "implicit_parallel_chains_depth5_loop_size5.urdf",
                                  "explicit_parallel_chains_depth5_loop_size4.urdf"])
def test_urdf_that_was_never_compiled_ahead_of_time(grbda, oracle, torch, urdf):
    """A URDF+ file outside build.py's MODELS: created on the device, kernels from NVRTC, 1e-10 against the oracle."""
    from mirror import mirror_to_oracle
    m = grbda.ClusterTreeModel.from_urdf(os.path.join(CORPUS, urdf))
    assert m.kernel_info(grbda.ALGO_ID)["source"] == "jit"
    o = mirror_to_oracle(m, oracle)
    check_against_oracle(m, o)
    info = m.kernel_info(grbda.ALGO_FD)
    assert info["ready"] and info["source"] == "jit"
    if any(c["type"] == 3 for c in m.clusters()):
        assert float(m.constraintViolation(m.generateStates(100)[0]).max()) < 1e-8


def test_tello_from_a_raw_schedule(grbda, oracle, torch):
    """TelloWithArms handed over as a grbda_schedule (the arrays a reference-side binding fills, including the
    phi programs of the four implicit clusters), with one inertia entry moved by one ulp so that no
    ahead-of-time kernel can match: compiled at run time, 1e-10 against the oracle's hand-coded TelloWithArms."""
    host = grbda.ClusterTreeModel.from_robot("tello_with_arms", device=None)
    s = host.to_schedule()
    assert s.cluster_phi_count.sum() > 0 and (s.body_independent == 0).sum() == 8
    s.body_inertia = s.body_inertia.copy()
    k = int(np.flatnonzero(s.body_inertia)[5])
    s.body_inertia[k] = np.nextafter(s.body_inertia[k], np.inf)
    m = grbda.ClusterTreeModel.from_schedule(s)
    assert m.hash != host.hash and (m.nq, m.nv, m.nb, m.nc) == (33, 24, 37, 15)
    assert m.kernel_info(grbda.ALGO_FD)["source"] == "jit"
    check_against_oracle(m, oracle.OracleModel("tello_with_arms"))
    # the unperturbed schedule finds the ahead-of-time kernels
    m0 = grbda.ClusterTreeModel.from_schedule(host.to_schedule())
    assert m0.hash == host.hash and m0.kernel_info(grbda.ALGO_FD)["source"] == "aot"


def test_run_time_kernels_equal_ahead_of_time_kernels(grbda, torch, monkeypatch):
    """Same model compiler, same shells, nvcc ahead of time vs NVRTC at run time."""
    aot = grbda.ClusterTreeModel.from_robot("mit_humanoid")
    monkeypatch.setenv("GRBDA_JIT", "force")
    jit = grbda.ClusterTreeModel.from_robot("mit_humanoid")
    monkeypatch.delenv("GRBDA_JIT")
    assert aot.kernel_info(0)["source"] == "aot" and jit.kernel_info(0)["source"] == "jit"
    q, yd, aux, _ = aot.generateStates(5000, seed=3)
    q2, yd2, aux2, _ = jit.generateStates(5000, seed=3)
    assert torch.equal(q, q2) and torch.equal(yd, yd2) and torch.equal(aux, aux2)
    for f in ("inverseDynamics", "forwardDynamics"):
        a, b = getattr(aot, f)(q, yd, aux), getattr(jit, f)(q, yd, aux)
        assert relrows(b.cpu().numpy(), a.cpu().numpy()) < 1e-12
    assert relrows(jit.getMassMatrix(q).cpu().numpy(), aot.getMassMatrix(q).cpu().numpy()) < 1e-12
    for a, b in zip(aot.forwardKinematics(q, yd), jit.forwardKinematics(q, yd)):
        assert relrows(b.cpu().numpy(), a.cpu().numpy()) < 1e-12
    # FP32 entry points are compiled on demand as well
    a = aot.inverseDynamics(q.float(), yd.float(), aux.float())
    b = jit.inverseDynamics(q.float(), yd.float(), aux.float())
    assert relrows(b.cpu().numpy(), a.cpu().numpy()) < 1e-5
    # pointers the bulk-copy engine cannot take (8-byte offset) go through the software-staged pass
    big = torch.zeros(q.numel() + 1, dtype=torch.float64, device=q.device)
    qo = big[1:].view_as(q)
    qo.copy_(q)
    assert qo.data_ptr() % 16 == 8
    assert relrows(jit.inverseDynamics(qo, yd, aux).cpu().numpy(), aot.inverseDynamics(q, yd, aux).cpu().numpy()) < 1e-12
    # disabled run-time compilation: a model without ahead-of-time kernels is refused at creation
    monkeypatch.setenv("GRBDA_JIT", "0")
    with pytest.raises(grbda.GrbdaError) as ei:
        grbda.ClusterTreeModel.from_urdf(os.path.join(CORPUS, "four_bar_branch_1_1.urdf"))
    assert ei.value.status == 3


def test_disk_cache_is_used_on_the_second_creation(grbda, torch, tmp_path, monkeypatch):
    monkeypatch.setenv("GRBDA_CACHE_DIR", str(tmp_path))
    path = os.path.join(CORPUS, "revolute_rotor_branch_2_3.urdf")
    m1 = grbda.ClusterTreeModel.from_urdf(path)
    m1.prepare(grbda.ALGO_ID)
    i1 = m1.kernel_info(grbda.ALGO_ID)
    assert i1["ready"] and not i1["from_cache"] and len(os.listdir(tmp_path)) == 1
    m2 = grbda.ClusterTreeModel.from_urdf(path)
    m2.prepare(grbda.ALGO_ID)
    assert m2.kernel_info(grbda.ALGO_ID)["from_cache"]
    q, yd, aux, _ = m1.generateStates(256)
    assert torch.equal(m1.inverseDynamics(q, yd, aux), m2.inverseDynamics(q, yd, aux))
