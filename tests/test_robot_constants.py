"""Transcription-independent check of the hand-coded robots: the constants parsed mechanically out of the reference's
include/grbda/Robots/{Tello,TelloWithArms}.hpp (tests/golden/robot_constants.json, made by
tests/golden/make_robot_constants.py) against the body tables of BOTH the product (csrc/host/robots.cpp) and the
oracle (oracle/grbda_oracle/robots.h). The body-name <-> constant-name correspondence is the reference builders' own
naming rule (src/Robots/Tello.cpp:33-60: body "<side>-hip-clamp" takes R_<side>_hip_clamp, p_<side>_hip_clamp and
hip_clamp_{mass,CoM,inertia}; src/Robots/TelloWithArms.cpp:20-160 with withLeftRightSigns)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "robot_constants.json")))


def skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def spatial_inertia(m, c, I):  # SpatialInertia(mass, com, inertia), include/grbda/Utils/SpatialInertia.h:74-82
    C = skew(np.asarray(c))
    M = np.zeros((6, 6))
    M[:3, :3] = np.asarray(I) + m * C @ C.T
    M[:3, 3:] = m * C
    M[3:, :3] = m * C.T
    M[3:, 3:] = m * np.eye(3)
    return M


def flip_y(v):
    return np.array([v[0], -v[1], v[2]])


def flip_inertia_y(I):  # SpatialInertia::flipAlongAxis(Y) on the rotational inertia: products with y change sign
    I = np.array(I, dtype=float)
    I[0, 1] = -I[0, 1]; I[1, 0] = -I[1, 0]; I[1, 2] = -I[1, 2]; I[2, 1] = -I[2, 1]
    return I


def body_tables(grbda, oracle, robot):
    yield "product", {b["name"]: b for b in grbda.ClusterTreeModel.from_robot(robot, device=None).bodies()}
    yield "oracle", {b["name"]: b for b in oracle.OracleModel(robot).bodies()}


@pytest.mark.parametrize("robot", ["tello", "tello_with_arms"])
def test_tello_legs_against_the_reference_header(grbda, oracle, robot):
    T = GOLD["tello"]
    for which, bodies in body_tables(grbda, oracle, robot):
        torso = bodies["torso"]
        assert np.allclose(torso["inertia"], spatial_inertia(T["torso_mass"], T["torso_CoM"], T["torso_inertia"]), rtol=0, atol=1e-15), which
        checked = 0
        for side in ("left", "right"):
            for part in ("hip_clamp", "gimbal", "thigh", "shin", "foot", "hip_clamp_rotor", "hip_rotor_1", "hip_rotor_2",
                         "knee_ankle_rotor_1", "knee_ankle_rotor_2"):
                b = bodies["%s-%s" % (side, part.replace("_", "-"))]
                assert np.array_equal(np.asarray(b["E"]).reshape(3, 3), np.array(T["R_%s_%s" % (side, part)])), (which, side, part)
                assert np.array_equal(np.asarray(b["r"]).reshape(3), np.array(T["p_%s_%s" % (side, part)])), (which, side, part)
                want = spatial_inertia(T[part + "_mass"], T[part + "_CoM"], T[part + "_inertia"])
                assert np.allclose(np.asarray(b["inertia"]).reshape(6, 6), want, rtol=0, atol=1e-17), (which, side, part)
                checked += 1
        assert checked == 20
    assert T["grav"] == -9.81 and T["gear_ratio"] == 6.0


def test_tello_arms_against_the_reference_header(grbda, oracle):
    A = GOLD["arms"]
    links = {"shoulder-ry": "_shoulderRy", "shoulder-rx": "_shoulderRx", "shoulder-rz-link": "_shoulderRz", "elbow-link": "_elbow"}
    rotors = {"shoulder-ry-rotor": "_shoulderRy", "shoulder-rx-rotor": "_shoulderRx", "shoulder-rz-rotor": "_shoulderRz",
              "elbow-rotor": "_elbow"}
    small = np.sort(np.diag(np.array(A["_smallRotorRotationalInertiaZ"])))
    for which, bodies in body_tables(grbda, oracle, "tello_with_arms"):
        for s, side in enumerate(("left", "right")):
            sign = (lambda v: np.asarray(v)) if s == 0 else flip_y
            for name, var in links.items():
                b = bodies["%s-%s" % (side, name)]
                assert np.array_equal(np.asarray(b["E"]).reshape(3, 3), np.eye(3)), (which, side, name)
                assert np.array_equal(np.asarray(b["r"]).reshape(3), sign(A[var + "Location"])), (which, side, name)
                I = np.array(A[var + "RotInertia"]) if s == 0 else flip_inertia_y(A[var + "RotInertia"])
                want = spatial_inertia(A[var + "Mass"], sign(A[var + "COM"]), I)
                assert np.allclose(np.asarray(b["inertia"]).reshape(6, 6), want, rtol=0, atol=1e-17), (which, side, name)
            for name, var in rotors.items():
                b = bodies["%s-%s" % (side, name)]
                assert np.array_equal(np.asarray(b["r"]).reshape(3), sign(A[var + "RotorLocation"])), (which, side, name)
                M = np.asarray(b["inertia"]).reshape(6, 6)
                # massless rotors: rotational inertia only, the Z inertia rotated onto the joint axis
                assert np.all(M[3:, :] == 0) and np.all(M[:, 3:] == 0), (which, side, name)
                assert np.allclose(np.sort(np.diag(M[:3, :3])), small, rtol=0, atol=1e-19), (which, side, name)
                assert np.abs(M[:3, :3] - np.diag(np.diag(M[:3, :3]))).max() < 1e-19, (which, side, name)
