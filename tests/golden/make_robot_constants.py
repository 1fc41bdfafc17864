"""Reads the numeric constants of the reference's Tello / TelloWithArms robots MECHANICALLY out of
/root/reference/include/grbda/Robots/{Tello,TelloWithArms}.hpp (regular expressions + arithmetic on the literals;
no table is typed by hand) and writes tests/golden/robot_constants.json. tests/test_robot_constants.py compares
them with the product's and the oracle's body tables - a check that is independent of both transcriptions
(csrc/host/robots.cpp and oracle/grbda_oracle/robots.h).
Run here (the reference tree does not exist on the GPU box): python tests/golden/make_robot_constants.py"""
import json
import os
import re

import numpy as np

REF = "/root/reference/include/grbda/Robots"
HERE = os.path.dirname(os.path.abspath(__file__))
NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)"


def numbers(text):
    return [float(x) for x in re.findall(NUM, text)]


def parse_tello():
    src = open(os.path.join(REF, "Tello.hpp")).read()
    env = {}
    # const Mat3<Scalar> X = (Mat3<Scalar>() << a, b, ... ).finished();   |  = Mat3<Scalar>::Identity();  |  = OTHER;
    for name, rhs in re.findall(r"const\s+Mat3<Scalar>\s+(\w+)\s*=\s*(.*?);", src, re.S):
        if "<<" in rhs:
            env[name] = np.array(numbers(rhs.split("<<", 1)[1])).reshape(3, 3)
        elif "Identity" in rhs:
            env[name] = np.eye(3)
        else:
            env[name] = env[rhs.strip()]
    for name, rhs in re.findall(r"const\s+Vec3<Scalar>\s+(\w+)\s*=\s*(.*?);", src, re.S):
        if "Zero" in rhs:
            env[name] = np.zeros(3)
        elif "{" in rhs:
            env[name] = np.array(numbers(rhs[rhs.index("{"):]))
        else:
            env[name] = env[rhs.strip()]
    for name, rhs in re.findall(r"const\s+Scalar\s+(\w+)\s*=\s*(.*?);", src, re.S):
        rhs = rhs.strip()
        env[name] = env[rhs] if rhs in env else float(rhs)
    return env


def parse_arms():
    src = open(os.path.join(REF, "TelloWithArms.hpp")).read()
    env = {}
    for name, rhs in re.findall(r"(\w+)\s*<<\s*(.*?);", src, re.S):
        env[name] = np.array(numbers(rhs)).reshape(3, 3)
    for name, rhs in re.findall(r"Vec3<Scalar>\s+(\w+)\s*=\s*Vec3<Scalar>\((.*?)\);", src):
        env[name] = np.array(numbers(rhs))
    for name, rhs in re.findall(r"\bScalar\s+(\w+)\s*=\s*(" + NUM + r");", src):
        env[name] = float(rhs)
    return env


def main():
    out = {"source": "include/grbda/Robots/Tello.hpp, TelloWithArms.hpp (parsed by tests/golden/make_robot_constants.py)",
           "tello": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in parse_tello().items()},
           "arms": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in parse_arms().items()}}
    with open(os.path.join(HERE, "robot_constants.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("tello: %d constants, arms: %d constants" % (len(out["tello"]), len(out["arms"])))


if __name__ == "__main__":
    main()
