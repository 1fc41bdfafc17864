"""GPU parity tests (run on the B200 box: pytest -m gpu). Every call goes through the C ABI
(generalized_rbda_b200 -> libgrbda_cuda.so); results are compared with the CPU oracle on the same
seeded states. Tolerances: <= 1e-10 relative (FP64), <= 1e-4 (FP32 variant) — BASELINE.json north_star."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
# product robot -> oracle robot (hand-coded builder restated from the reference) or None = the oracle is
# assembled from the product's topology (tests/mirror.py; URDF-only models)
ROBOTS = {"tello": "tello", "tello_with_arms": "tello_with_arms",
          "mini_cheetah": "mini_cheetah", "mit_humanoid": "mit_humanoid",
          "four_bar": None, "revolute_rotor_chain": None, "six_bar": None, "planar_leg_linkage": None,
          "mit_humanoid_leg": None, "jvrc1_humanoid": None,
          "revolute_chain_with_rotor_2": "revolute_chain_with_rotor_2",
          "revolute_chain_with_rotor_4": "revolute_chain_with_rotor_4",
          "revolute_chain_with_rotor_8": "revolute_chain_with_rotor_8",
          "revolute_chain_with_rotor_16": "revolute_chain_with_rotor_16",
          "revolute_pair_chain_with_rotor_2": "revolute_pair_chain_with_rotor_2",
          "revolute_pair_chain_with_rotor_4": "revolute_pair_chain_with_rotor_4",
          # oracle: restated ClusterJoints::RevolutePair / RevoluteTripleWithRotor on the product's topology
          "revolute_pair_chain_4": None, "revolute_triple_chain_with_rotor_6": None}


def oracle_for(oracle, m, robot):
    if ROBOTS[robot] is not None:
        return oracle.OracleModel(ROBOTS[robot])
    from mirror import mirror_to_oracle
    return mirror_to_oracle(m, oracle)
TOL64, TOL32 = 1e-10, 1e-4


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def relrows(a, b, floor=1e-3):
    """largest per-state relative error: each state is normalised by its own largest entry, with an
    absolute floor so that quantities that are identically zero (the bias force of the parallelogram
    four-bar, for instance) are compared absolutely"""
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    return (np.abs(a - b).max(1) / np.maximum(floor, np.abs(b).max(1))).max()


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("robot", sorted(ROBOTS))
def test_state_generator_matches_oracle(grbda, oracle, torch, robot):
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    q, yd, aux, flags = m.generateStates(777, seed=123, first_index=1000)
    assert int(flags.sum()) == 0
    qo, ydo, auxo = o.generate_states(777, seed=123, first_index=1000)
    # positions: identical draws; implicit clusters go through Newton (with redraws on failure) on both
    # sides. A few states of a multi-loop linkage take a different number of redraws or land on another
    # assembly branch in the last bits of the iteration: those are excluded from the bit-exact comparison
    # of the draws that follow; every state is checked for validity below.
    close = np.abs(q.cpu().numpy() - qo).max(1) < 1e-9
    assert close.mean() > (0.97 if robot == "six_bar" else 0.99)
    assert np.array_equal(yd.cpu().numpy()[close], ydo[close]) and np.array_equal(aux.cpu().numpy()[close], auxo[close])
    if robot != "six_bar":
        assert np.array_equal(yd.cpu().numpy(), ydo) and np.array_equal(aux.cpu().numpy(), auxo)
    assert float(m.constraintViolation(q).max()) < 1e-8
    assert o.validate_states(q.cpu().numpy()).all()


@pytest.mark.parametrize("robot", sorted(ROBOTS))
def test_dynamics_parity_f64(grbda, oracle, torch, robot):
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    B = 1000  # not a multiple of the CTA size: exercises the ragged last tile
    q, yd, aux, _ = m.generateStates(B, seed=42)
    qn, ydn, auxn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    tau_o = o.inverse_dynamics(qn, ydn, auxn)
    assert relrows(m.inverseDynamics(q, yd, aux).cpu().numpy(), tau_o) < TOL64
    assert relrows(m.forwardDynamics(q, yd, aux).cpu().numpy(), o.forward_dynamics(qn, ydn, auxn)) < TOL64
    assert relrows(m.getMassMatrix(q).cpu().numpy(), o.mass_matrix(qn)) < TOL64
    p, R, v = m.forwardKinematics(q, yd)
    po, Ro, vo = o.forward_kinematics(qn, ydn)
    assert rel(p.cpu().numpy(), po) < TOL64 and rel(R.cpu().numpy(), Ro) < TOL64 and rel(v.cpu().numpy(), vo) < TOL64
    # bias force: for the parallelogram four-bar it is a difference of O(1) terms that cancel to ~1e-3, so
    # the error is measured against the size of the state's generalized forces, not against |C| alone
    C = m.getBiasForceVector(q, yd).cpu().numpy()
    C_o = o.inverse_dynamics(qn, ydn, np.zeros_like(ydn))
    scale = np.maximum(np.abs(tau_o).max(1), np.abs(C_o).max(1))
    assert (np.abs(C - C_o).max(1) / np.maximum(1e-3, scale)).max() < TOL64


@pytest.mark.parametrize("robot", ["tello_with_arms", "mit_humanoid", "four_bar", "revolute_pair_chain_with_rotor_4"])
def test_every_compiled_kernel_variant(grbda, oracle, torch, robot, monkeypatch):
    """Every launch shape / program compiled for a model (GRBDA_KERNEL_VARIANT): the forward-dynamics slots
    hold both programs (CRBA + sparse LTDL, and the articulated-body sweep); all must agree with the oracle."""
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    q, yd, aux, _ = m.generateStates(777, seed=11)
    qn, ydn, auxn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    ydd_o, tau_o = o.forward_dynamics(qn, ydn, auxn), o.inverse_dynamics(qn, ydn, auxn)
    ran = 0
    for variant in range(4):
        monkeypatch.setenv("GRBDA_KERNEL_VARIANT", str(variant))
        try:
            ydd = m.forwardDynamics(q, yd, aux)
        except grbda.GrbdaError:
            continue  # fewer than 4 variants compiled for this model
        ran += 1
        assert relrows(ydd.cpu().numpy(), ydd_o) < TOL64
        try:
            assert relrows(m.inverseDynamics(q, yd, aux).cpu().numpy(), tau_o) < TOL64
        except grbda.GrbdaError:
            pass
    assert ran >= 2


@pytest.mark.parametrize("robot", ["tello_with_arms", "mini_cheetah", "mit_humanoid", "four_bar", "jvrc1_humanoid"])
def test_external_forces(grbda, oracle, torch, robot):
    """grbda_cuda_{inverse,forward}_dynamics_ext_f64: world-frame spatial forces on the terminal links
    (TreeModel::setExternalForces) against the oracle's RNEA / ABA with external forces."""
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    bodies = m.externalForceBodies()
    B = 300
    q, yd, aux, _ = m.generateStates(B, seed=17)
    g = torch.Generator(device="cuda").manual_seed(3)
    f = (torch.rand((B, len(bodies), 6), dtype=torch.float64, device="cuda", generator=g) - 0.5) * 40.0
    qn, ydn, auxn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    f_full = np.zeros((B, o.nb, 6))
    f_full[:, bodies, :] = f.cpu().numpy()
    tau = m.inverseDynamics(q, yd, aux, f_ext=f)
    assert relrows(tau.cpu().numpy(), o.dynamics_with_external_forces(qn, ydn, auxn, f_full, forward=False)) < TOL64
    ydd = m.forwardDynamics(q, yd, aux, f_ext=f)
    assert relrows(ydd.cpu().numpy(), o.dynamics_with_external_forces(qn, ydn, auxn, f_full, forward=True)) < TOL64
    # consistency: ID_ext(FD_ext(tau)) = tau, and zero forces reproduce the plain entry points bit for bit
    back = m.inverseDynamics(q, yd, ydd, f_ext=f)
    assert float(((back - aux).abs().amax(1) / aux.abs().amax(1)).median()) < 1e-9
    zero = torch.zeros_like(f)
    assert torch.equal(m.forwardDynamics(q, yd, aux, f_ext=zero), m.forwardDynamics(q, yd, aux))


@pytest.mark.parametrize("robot", ["tello_with_arms", "mini_cheetah", "revolute_rotor_chain"])
def test_external_forces_on_every_body(grbda, oracle, torch, robot):
    """The reference's own external-force test (UnitTests/testRigidBodyDynamicsAlgos.cpp:201-236): a random
    spatial force on EVERY body, rotors included. The force programs are specialised for the chosen body set
    at run time (NVRTC) even though the model's other kernels were built ahead of time."""
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    m.setExternalForceBodies(list(range(m.nb)))
    B = 300
    q, yd, aux, _ = m.generateStates(B, seed=19)
    g = torch.Generator(device="cuda").manual_seed(4)
    f = (torch.rand((B, m.nb, 6), dtype=torch.float64, device="cuda", generator=g) - 0.5) * 40.0
    qn, ydn, auxn, fn = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy(), f.cpu().numpy()
    tau = m.inverseDynamics(q, yd, aux, f_ext=f)
    assert relrows(tau.cpu().numpy(), o.dynamics_with_external_forces(qn, ydn, auxn, fn, forward=False)) < TOL64
    ydd = m.forwardDynamics(q, yd, aux, f_ext=f)
    assert relrows(ydd.cpu().numpy(), o.dynamics_with_external_forces(qn, ydn, auxn, fn, forward=True)) < TOL64
    # back to the default set: the ahead-of-time programs again
    m.setExternalForceBodies([])
    bodies = m.externalForceBodies()
    f_def = f[:, bodies, :].contiguous()
    f_full = np.zeros_like(fn)
    f_full[:, bodies, :] = fn[:, bodies, :]
    tau = m.inverseDynamics(q, yd, aux, f_ext=f_def)
    assert relrows(tau.cpu().numpy(), o.dynamics_with_external_forces(qn, ydn, auxn, f_full, forward=False)) < TOL64


@pytest.mark.parametrize("robot", ["tello_with_arms", "mit_humanoid", "mini_cheetah"])
def test_operational_space(grbda, oracle, torch, robot):
    """Batched contact kinematics, contact Jacobians, applyTestForce and inverse operational-space inertia matrix
    (ClusterTreeDynamics.cpp:10-77,193-435) on the feet / hands of the three BASELINE robots, against
    the oracle (restated Jacobians; J H^-1 J^T as the reference's own tests define the expected values)."""
    from test_compiler_cpu import contact_set
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    bodies, offsets, ee = contact_set(m)
    m.setContactPoints(bodies, offsets, ee)
    o.set_contact_points(bodies, offsets, ee)
    B = 300
    q, yd, aux, _ = m.generateStates(B, seed=31)
    qn, ydn = q.cpu().numpy(), yd.cpu().numpy()
    p, v = m.contactKinematics(q, yd)
    po, vo = o.contact_kinematics(qn, ydn)
    assert relrows(p.cpu().numpy(), po) < TOL64 and relrows(v.cpu().numpy(), vo) < TOL64
    J = m.contactJacobians(q)
    assert relrows(J.cpu().numpy(), o.contact_jacobians(qn, world=True)) < TOL64
    g = torch.Generator(device="cuda").manual_seed(5)
    f = torch.rand((B, m.ncp, 3), dtype=torch.float64, device="cuda", generator=g) - 0.5
    d, lam = m.applyTestForce(q, f)
    do, lamo = o.apply_test_force(qn, f.cpu().numpy())
    assert relrows(d.cpu().numpy(), do) < 1e-8 and relrows(lam.cpu().numpy(), lamo, floor=1e-12) < 1e-8
    # the device's own mass matrix agrees: H dstate = J_lin^T f
    H = m.getMassMatrix(q)
    lhs = torch.einsum("bij,bcj->bci", H, d)
    rhs = torch.einsum("bcik,bci->bck", J[:, :, 3:, :], f)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-9
    L = m.inverseOperationalSpaceInertiaMatrix(q)
    assert relrows(L.cpu().numpy(), o.inverse_osim(qn)) < 1e-8
    # a new contact set replaces the programs
    m.setContactPoints(bodies[:1], offsets[:1], [1])
    o.set_contact_points(bodies[:1], offsets[:1], [1])
    assert relrows(m.inverseOperationalSpaceInertiaMatrix(q).cpu().numpy(), o.inverse_osim(qn)) < 1e-8


@pytest.mark.parametrize("robot", ["tello_with_arms", "mit_humanoid", "four_bar", "revolute_chain_with_rotor_4"])
def test_integration_step(grbda, oracle, torch, robot):
    """grbda_cuda_integrate_f64 / grbda_cuda_step_f64 against the oracle's restatement (ori::integrateQuat,
    OrientationTools.h:387-413; implicit clusters projected back onto phi = 0), a short trajectory of the whole
    simulation step, and energy behaviour of the unforced system."""
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    B, dt = 500, 1e-3
    q, yd, aux, _ = m.generateStates(B, seed=29)
    ydd = m.forwardDynamics(q, yd, aux)
    q1, yd1, flags = m.integrate(q, yd, ydd, dt)
    assert int(flags.sum()) == 0
    qo, ydo, fo = o.integrate(q.cpu().numpy(), yd.cpu().numpy(), ydd.cpu().numpy(), dt)
    assert not fo.any()
    assert np.abs(q1.cpu().numpy() - qo).max() < 1e-12 and np.abs(yd1.cpu().numpy() - ydo).max() < 1e-13
    assert float(m.constraintViolation(q1).max()) < 1e-10
    # step = forwardDynamics + integrate; ten steps on the device against ten steps of the oracle
    qs, yds = q.clone(), yd.clone()
    qn, ydn, taun = q.cpu().numpy(), yd.cpu().numpy(), aux.cpu().numpy()
    for _ in range(10):
        qs, yds, fl = m.step(qs, yds, aux, dt)
        assert int(fl.sum()) == 0
        qn, ydn, _ = o.integrate(qn, ydn, o.forward_dynamics(qn, ydn, taun), dt)
    assert np.abs(qs.cpu().numpy() - qn).max() < 1e-8 and relrows(yds.cpu().numpy(), ydn) < 1e-8
    # in place (q_out = q, yd_out = yd) gives the same result
    qa, yda = q.clone(), yd.clone()
    m.integrate(qa, yda, ydd, dt, out=(qa, yda))
    assert torch.equal(qa, q1) and torch.equal(yda, yd1)


@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_angles_beyond_the_fast_sincos_range(grbda, oracle, torch, dtype_name):
    """Joint angles beyond the range of the branch-free sin/cos reduction (arguments up to 1e12 in FP64, 1e6 in FP32):
    the fast kernel flags the 128-state tile, the second pass recomputes it with the library forms.
    Tiles 0 and 2 of a 3.x-tile batch are flagged, tile 1 and the ragged tail stay on the fast path."""
    m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
    o = oracle.OracleModel("tello_with_arms")
    dtype = getattr(torch, dtype_name)
    B = 3 * 128 + 40
    q, yd, aux, _ = m.generateStates(B, seed=5)
    q = q.clone()
    turns = 4.0e11 if dtype_name == "float64" else 4.0e5       # whole turns: 2.5e12 rad / 2.5e6 rad
    hip_clamp = m.clusters()[1]["position_index"]
    arm = m.clusters()[7]["position_index"]
    q[5, hip_clamp] += 2.0 * np.pi * turns
    q[300, arm] -= 2.0 * np.pi * turns
    # large but inside the fast range (tile 1 stays on the fast path): the FMA reduction must hold up
    q[200, hip_clamp] += 2.0 * np.pi * (1.0e8 if dtype_name == "float64" else 1.0e3)
    q, yd, aux = q.to(dtype), yd.to(dtype), aux.to(dtype)
    qn, ydn, auxn = (x.double().cpu().numpy() for x in (q, yd, aux))
    tol = TOL64 if dtype_name == "float64" else TOL32
    before = grbda.launch_count()
    tau = m.inverseDynamics(q, yd, aux)
    assert grbda.launch_count() - before == 2       # fast kernel + flagged-tile pass
    assert relrows(tau.double().cpu().numpy(), o.inverse_dynamics(qn, ydn, auxn)) < tol
    assert relrows(m.getMassMatrix(q).double().cpu().numpy(), o.mass_matrix(qn)) < tol
    if dtype_name == "float64":
        # (FP32 angles of this size carry ~0.1 rad of rounding in the rotor angle 6 q: FK and FD are FP64-only here)
        p, R, v = m.forwardKinematics(q, yd)
        po, Ro, vo = o.forward_kinematics(qn, ydn)
        assert rel(R.cpu().numpy(), Ro) < tol and rel(v.cpu().numpy(), vo) < tol
        assert relrows(m.forwardDynamics(q, yd, aux).cpu().numpy(), o.forward_dynamics(qn, ydn, auxn)) < tol
    # NaN / inf positions: flagged like any other out-of-range value, results are NaN for that state only
    q2 = q.clone()
    q2[130, hip_clamp] = float("nan")
    tau2 = m.inverseDynamics(q2, yd, aux)
    assert torch.isnan(tau2[130]).any() and not torch.isnan(tau2[131]).any() and not torch.isnan(tau2[0]).any()


@pytest.mark.parametrize("robot", ["tello_with_arms", "mini_cheetah", "mit_humanoid", "revolute_chain_with_rotor_2",
                                   "revolute_chain_with_rotor_16", "revolute_rotor_chain", "jvrc1_humanoid"])
def test_dynamics_parity_f32(grbda, oracle, torch, robot):
    """FP32 variant against the FP64 oracle on float-rounded states."""
    m = grbda.ClusterTreeModel.from_robot(robot)
    o = oracle_for(oracle, m, robot)
    q, yd, aux, _ = m.generateStates(512, seed=4)
    q32, yd32, aux32 = q.float(), yd.float(), aux.float()
    qn, ydn, auxn = (x.double().cpu().numpy() for x in (q32, yd32, aux32))
    # well conditioned comparison: ID and H (FD amplifies rounding by cond(H))
    assert relrows(m.inverseDynamics(q32, yd32, aux32).double().cpu().numpy(), o.inverse_dynamics(qn, ydn, auxn)) < TOL32
    assert relrows(m.getMassMatrix(q32).double().cpu().numpy(), o.mass_matrix(qn)) < TOL32
    ydd32 = m.forwardDynamics(q32, yd32, aux32).double().cpu().numpy()
    ydd = o.forward_dynamics(qn, ydn, auxn)
    # forward dynamics: north_star's FP32 bound (<= 1e-4) on EVERY state, the typical state an order of magnitude
    # better, and the conditioning-scaled bound of a backward-stable solve of H ydd = tau - C: err <= 2 eps32 cond(H)
    # (measured, profiles/r2_fp32_fd_errors.jsonl: medians 1e-7 .. 1e-6, maxima <= 2e-5, err / (eps cond) <= 0.65)
    err = np.abs(ydd32 - ydd).max(1) / np.abs(ydd).max(1)
    assert err.max() < 1e-4 and np.median(err) < 1e-5
    cond = np.linalg.cond(o.mass_matrix(qn))
    assert (err / (np.finfo(np.float32).eps * cond)).max() < 2.0
    # and through the residual tau = ID(ydd) in FP64 (the residual is the error times |H|: link inertias next to rotor
    # inertias 1e3 times smaller)
    tau_back = o.inverse_dynamics(qn, ydn, ydd32)
    res = np.abs(tau_back - auxn).max(1) / np.abs(auxn).max(1)
    assert np.median(res) < 1e-3


def test_golden_vectors_on_gpu(grbda, torch):
    """The reference's own generated closed-form dynamics (tests/golden/codegen_kat.json)."""
    kat = json.load(open(os.path.join(HERE, "golden", "codegen_kat.json")))
    for case in kat["cases"]:
        m = grbda.ClusterTreeModel.from_robot(case["robot"])
        y = torch.tensor([s["y"] for s in case["states"]], dtype=torch.float64, device="cuda")
        yd = torch.tensor([s["yd"] for s in case["states"]], dtype=torch.float64, device="cuda")
        tau = torch.tensor([s["tau"] for s in case["states"]], dtype=torch.float64, device="cuda")
        ydd_ref = np.array([s["ydd_fwd"] for s in case["states"]])
        tau_ref = np.array([s["tau_inv_of_ydd"] for s in case["states"]])
        assert relrows(m.forwardDynamics(y, yd, tau).cpu().numpy(), ydd_ref) < TOL64
        ydd = torch.tensor(ydd_ref, dtype=torch.float64, device="cuda")
        assert relrows(m.inverseDynamics(y, yd, ydd).cpu().numpy(), tau_ref) < TOL64


def test_full_size_round_trip_and_sharding(grbda, torch):
    """BASELINE size (2^20 Tello states): ID(FD(tau)) = tau; shards reproduce the global batch."""
    m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
    B = 1 << 20
    q, yd, tau, flags = m.generateStates(B)
    assert int(flags.sum()) == 0
    ydd = m.forwardDynamics(q, yd, tau)
    tau2 = m.inverseDynamics(q, yd, ydd)
    err = ((tau2 - tau).abs().amax(1) / tau.abs().amax(1))
    assert float(err.median()) < 1e-11 and float(torch.quantile(err[:1 << 18], 0.99)) < 1e-8
    # the few large round-trip errors are conditioning (cond(H) from rotor inertias ~1e-5 and nearly
    # singular loop Jacobians), not the kernels: the CPU oracle loses the same digits on those states
    worst = torch.topk(err, 16).indices
    qw, ydw, tw = q[worst].cpu().numpy(), yd[worst].cpu().numpy(), tau[worst].cpu().numpy()
    from oracle import binding
    o = binding.OracleModel("tello_with_arms")
    err_o = np.abs(o.inverse_dynamics(qw, ydw, o.forward_dynamics(qw, ydw, tw)) - tw).max(1) / np.abs(tw).max(1)
    assert np.median(err_o) > 1e-3 * float(err[worst].median())
    H = m.getMassMatrix(q[:4096])
    C = m.getBiasForceVector(q[:4096], yd[:4096])
    res = torch.einsum("bij,bj->bi", H, ydd[:4096]) + C - tau[:4096]
    assert float((res.abs().amax(1) / tau[:4096].abs().amax(1)).median()) < 1e-10
    # two half shards generated independently == the global batch; checksums add up
    qa, yda, ta, _ = m.generateStates(B // 2, first_index=0)
    qb, ydb, tb, _ = m.generateStates(B // 2, first_index=B // 2)
    assert torch.equal(torch.cat([qa, qb]), q) and torch.equal(torch.cat([ta, tb]), tau)
    ya = m.forwardDynamics(qa, yda, ta)
    yb = m.forwardDynamics(qb, ydb, tb)
    assert torch.equal(torch.cat([ya, yb]), ydd)
    sa, sb, s = grbda.checksum(ya), grbda.checksum(yb), grbda.checksum(ydd)
    assert abs(sa[1] + sb[1] - s[1]) <= 1e-9 * s[1]


def test_host_buffer_path(grbda, oracle, torch, monkeypatch):
    m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
    o = oracle.OracleModel("tello_with_arms")
    monkeypatch.setenv("GRBDA_HOST_CHUNK", "16384")
    B = 70000  # five chunks of 16 384 states: stream rotation, buffer reuse, a ragged last chunk
    q, yd, tau, _ = m.generateStates(B, seed=8)
    qh, ydh, tauh = (x.cpu().pin_memory() for x in (q, yd, tau))
    out = torch.empty((B, m.nv), dtype=torch.float64).pin_memory()
    m.dynamics_host(1, qh, ydh, tauh, out)
    assert torch.equal(out, m.forwardDynamics(q, yd, tau).cpu())
    sel = slice(0, 256)
    assert relrows(out[sel].numpy(), o.forward_dynamics(qh[sel].numpy(), ydh[sel].numpy(), tauh[sel].numpy())) < TOL64
    ydd_h = out.clone()
    m.dynamics_host(0, qh, ydh, ydd_h, out)
    rt = np.abs(out.numpy() - tauh.numpy()).max(1) / np.abs(tauh.numpy()).max(1)
    assert np.median(rt) < 1e-11  # ID(FD(tau)) = tau; the tail is conditioning (see the full-size test)
    a, b = torch.empty_like(out), torch.empty_like(out)
    m.forward_inverse_host(qh, ydh, tauh, a, b)
    assert torch.equal(a, ydd_h) and torch.equal(b, out)
    # pageable (numpy) buffers are staged through the handle's pinned buffers: same bits, both outputs
    monkeypatch.delenv("GRBDA_HOST_CHUNK")
    B2 = 3 * (1 << 15) + 1234  # more than three staged chunks: every stream's buffer is reused
    q2, yd2, tau2, _ = m.generateStates(B2, seed=11)
    qn, ydn, taun = (x.cpu().numpy().copy() for x in (q2, yd2, tau2))
    an, bn = np.empty_like(taun), np.empty_like(taun)
    m.forward_inverse_host(qn, ydn, taun, an, bn)
    ydd_dev = m.forwardDynamics(q2, yd2, tau2)
    assert np.array_equal(an, ydd_dev.cpu().numpy())
    assert np.array_equal(bn, m.inverseDynamics(q2, yd2, ydd_dev).cpu().numpy())
    # malformed host arrays are refused before any copy is issued
    with pytest.raises(ValueError):
        m.forward_inverse_host(qn.astype(np.float32), ydn, taun, an, bn)
    with pytest.raises(ValueError):
        m.forward_inverse_host(qn[:, ::-1], ydn, taun, an, bn)
    with pytest.raises(ValueError):
        m.forward_inverse_host(qn, ydn[:-1], taun, an, bn)


def test_cuda_graph_capture(grbda, oracle, torch):
    """The batched entry points only enqueue kernels on the caller's stream (their scratch is allocated on
    first use), so a control-loop sized step can be captured once in a CUDA graph and replayed."""
    m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
    o = oracle.OracleModel("tello_with_arms")
    B = 4096
    q, yd, tau, _ = m.generateStates(B, seed=9)
    ydd, back = torch.empty_like(tau), torch.empty_like(tau)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):               # warm-up on the capture stream: scratch buffers exist afterwards
        m.forwardDynamics(q, yd, tau, out=ydd)
        m.inverseDynamics(q, yd, ydd, out=back)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        m.forwardDynamics(q, yd, tau, out=ydd)
        m.inverseDynamics(q, yd, ydd, out=back)
    ydd.zero_(), back.zero_()
    q2, yd2, tau2, _ = m.generateStates(B, seed=10)
    q.copy_(q2), yd.copy_(yd2), tau.copy_(tau2)   # new inputs in the captured buffers
    graph.replay()
    torch.cuda.synchronize()
    assert relrows(ydd.cpu().numpy(), o.forward_dynamics(q.cpu().numpy(), yd.cpu().numpy(), tau.cpu().numpy())) < TOL64
    assert float(((back - tau).abs().amax(1) / tau.abs().amax(1)).median()) < 1e-9


def test_concurrent_streams_share_a_model(grbda, torch):
    """Launches of one model handle on different streams use different scratch buffers (tile flags) and may
    overlap; results equal the single-stream ones bit for bit."""
    m = grbda.ClusterTreeModel.from_robot("tello_with_arms")
    B = 1 << 16
    q, yd, tau, _ = m.generateStates(B, seed=23)
    q = q.clone()
    q[::4099, m.clusters()[1]["position_index"]] = 5.0e13       # some tiles need the second pass
    ref_fd, ref_id = m.forwardDynamics(q, yd, tau), m.inverseDynamics(q, yd, tau)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    for rep in range(4):
        for s in streams:
            with torch.cuda.stream(s):
                outs.append((m.forwardDynamics(q, yd, tau), m.inverseDynamics(q, yd, tau)))
    torch.cuda.synchronize()
    for fd, idn in outs:
        assert torch.equal(fd, ref_fd) and torch.equal(idn, ref_id)


def test_empty_and_error_paths(grbda, torch):
    m = grbda.ClusterTreeModel.from_robot("tello")
    q = torch.zeros((0, m.nq), dtype=torch.float64, device="cuda")
    yd = torch.zeros((0, m.nv), dtype=torch.float64, device="cuda")
    assert m.forwardDynamics(q, yd, yd).shape == (0, m.nv)
    with pytest.raises(ValueError):
        m.forwardDynamics(torch.zeros((4, m.nq + 1), dtype=torch.float64, device="cuda"), yd, yd)
    with pytest.raises(grbda.GrbdaError):
        grbda.ClusterTreeModel.from_robot("tello", device=99)
