"""Rebuild a product model inside the oracle through the oracle's registerBody / append* builder
(test infrastructure). Used for URDF-derived models that have no hand-coded builder in the
reference (four_bar, revolute_rotor_chain with the URDF's default gravity, ...): topology, inertias,
Xtree, G and the recorded phi program come from the product's introspection API; the dynamics are
then evaluated independently by the oracle's dense cluster algorithms."""
import numpy as np


def mirror_to_oracle(model, oracle, generic=False):
    b = oracle.OracleBuilder(gravity=tuple(model.getGravity()))
    bodies = model.bodies()
    for c, cl in enumerate(model.clusters()):
        first, n = cl["first_body"], cl["num_bodies"]
        axes = []
        for body in bodies[first:first + n]:
            parent = "ground" if body["parent"] < 0 else bodies[body["parent"]]["name"]
            b.register_body(body["name"], parent, body["inertia"], body["E"], body["r"])
            axes.append(body["axis"])
        name = "cluster-%d" % c
        if cl["type"] in (0, 1):
            b.append_simple(name, cl["type"])
        elif cl["type"] == 2 and cl["joint_type"] == "RevolutePair" and not generic:
            b.append_simple(name, 7, axes)
        elif cl["type"] == 2 and cl["joint_type"] == "RevoluteTripleWithRotor" and not generic:
            # gear / belt ratios back out of G (6 x 3): row 3 + i = gear_i * cumprod(belts_i); any factorisation
            # gives the same G, so gear = 1 and belts = successive quotients
            G = cl["G"]
            belts = [G[3, 0], G[4, 0], G[4, 1] / G[4, 0], G[5, 0], G[5, 1] / G[5, 0], G[5, 2] / G[5, 1]]
            b.append_revolute_triple_with_rotor(name, axes, [1.0, 1.0, 1.0], belts)
        elif cl["type"] == 2:
            G = cl["G"]
            K = np.zeros((n - G.shape[1], n))  # K does not enter the dynamics
            b.append_generic_static(name, axes, K, G)
        else:
            b.append_generic_phi(name, axes, cl["independent"], cl["phi_ops"], cl["phi_outputs"])
    return b.finish(generic=generic)
