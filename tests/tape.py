"""Numpy interpreter for the programs emitted by the device-side model compiler
(format: generalized_rbda_b200/csrc/compiler/compile.h, writeTape). Test infrastructure: lets the
CPU test-suite check the model compiler against the oracle without a GPU. Vectorised over states."""
import numpy as np


def load_tape(path):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:24], dtype=np.int32)
    assert hdr[0] == 0x47524244, "bad tape magic"
    n, off = int(hdr[1]), 24
    cols = []
    for _ in range(5):
        cols.append(np.frombuffer(raw[off:off + 4 * n], dtype=np.int32))
        off += 4 * n
    val = np.frombuffer(raw[off:off + 8 * n], dtype=np.float64)
    off += 8 * n
    outs = []
    for _ in range(int(hdr[2])):
        c = int(np.frombuffer(raw[off:off + 4], dtype=np.int32)[0])
        off += 4
        outs.append(np.frombuffer(raw[off:off + 4 * c], dtype=np.int32))
        off += 4 * c
    return dict(op=cols[0], a=cols[1], b=cols[2], c=cols[3], e=cols[4], val=val, outs=outs, n_in=hdr[3:6])


def run_tape(t, inputs, dtype=np.float64):
    """inputs: list of [batch, n_k] arrays (q, yd, aux). Returns one [batch, n] array per output."""
    B = inputs[0].shape[0]
    v = [None] * len(t["op"])
    for i, (op, a, b, c, e) in enumerate(zip(t["op"], t["a"], t["b"], t["c"], t["e"])):
        if op == 0:
            v[i] = np.full(B, t["val"][i], dtype=dtype)
        elif op == 1:
            v[i] = inputs[a][:, b].astype(dtype)
        elif op == 2:
            v[i] = v[a] + v[b]
        elif op == 3:
            v[i] = v[a] - v[b]
        elif op == 4:
            v[i] = v[a] * v[b]
        elif op == 5:
            v[i] = v[a] / v[b]
        elif op == 6:
            v[i] = -v[a]
        elif op == 7:
            v[i] = np.sin(v[a])
        elif op == 8:
            v[i] = np.cos(v[a])
        elif op == 9:
            v[i] = np.sqrt(v[a])
        elif op == 10:
            v[i] = np.where(v[a] > v[b], v[c], v[e])
        else:
            raise ValueError("unknown op %d" % op)
    return [np.stack([v[j] for j in o], axis=1) if len(o) else np.zeros((B, 0)) for o in t["outs"]]


def load_role_tape(path):
    """Limb-parallel programs (writeRoleTape in generalized_rbda_b200/csrc/compiler/compile.h)."""
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:40], dtype=np.int32)
    assert hdr[0] == 0x47524245, "bad role tape magic"
    n, W, slots = int(hdr[1]), int(hdr[2]), int(hdr[3])
    off = 40
    cols = []
    for _ in range(5):
        cols.append(np.frombuffer(raw[off:off + 4 * n], dtype=np.int32))
        off += 4 * n
    val = np.frombuffer(raw[off:off + 8 * n], dtype=np.float64)
    off += 8 * n
    roles = []
    for _ in range(W):
        no = int(np.frombuffer(raw[off:off + 4], dtype=np.int32)[0])
        off += 4
        ops = np.frombuffer(raw[off:off + 12 * no], dtype=np.int32).reshape(no, 3)
        off += 12 * no
        nout = int(np.frombuffer(raw[off:off + 4], dtype=np.int32)[0])
        off += 4
        outs = np.frombuffer(raw[off:off + 12 * nout], dtype=np.int32).reshape(nout, 3)
        off += 12 * nout
        roles.append(dict(ops=ops, outs=outs))
    return dict(op=cols[0], a=cols[1], b=cols[2], c=cols[3], e=cols[4], val=val, W=W, slots=slots,
                n_in=hdr[4:7], n_out=hdr[7:10], roles=roles)


def _eval_node(t, v, i, inputs, B):
    op, a, b, c, e = t["op"][i], t["a"][i], t["b"][i], t["c"][i], t["e"][i]
    if op == 0:
        return np.full(B, t["val"][i])
    if op == 1:
        return inputs[a][:, b]
    if op == 2:
        return v[a] + v[b]
    if op == 3:
        return v[a] - v[b]
    if op == 4:
        return v[a] * v[b]
    if op == 5:
        return v[a] / v[b]
    if op == 6:
        return -v[a]
    if op == 7:
        return np.sin(v[a])
    if op == 8:
        return np.cos(v[a])
    if op == 9:
        return np.sqrt(v[a])
    if op == 10:
        return np.where(v[a] > v[b], v[c], v[e])
    raise ValueError("unknown op %d" % op)


def run_role_tape(t, inputs):
    """Simulates the W warps of a CTA: every role runs to its barrier with a PRIVATE value table
    (a role may only see what it computed itself or loaded from a communication slot), then the
    second phase. Returns the output arrays and per-role instruction counts."""
    B = inputs[0].shape[0]
    W = t["W"]
    comm = {}
    tables = [dict() for _ in range(W)]
    pcs = [0] * W

    class View(dict):
        pass

    def ensure(r, i):
        # constants and negations are inlined by the emitter; evaluate them on demand, privately
        tab = tables[r]
        if i in tab:
            return tab[i]
        op = t["op"][i]
        if op == 0:
            tab[i] = np.full(B, t["val"][i])
        elif op == 6:
            tab[i] = -ensure(r, t["a"][i])
        else:
            raise AssertionError("role %d uses node %d (op %d) that it neither computed nor loaded" % (r, i, op))
        return tab[i]

    class Lookup:
        def __init__(self, r):
            self.r = r

        def __getitem__(self, i):
            return ensure(self.r, int(i))

    def run_phase(r, stop_at_barrier):
        ops = t["roles"][r]["ops"]
        while pcs[r] < len(ops):
            kind, i, slot = (int(x) for x in ops[pcs[r]])
            pcs[r] += 1
            if kind == 2:
                if stop_at_barrier:
                    return
                continue
            if kind == 1:
                comm[slot] = ensure(r, i)
            elif kind == 3:
                assert slot in comm, "communication slot read before it was written"
                tables[r][i] = comm[slot]
            else:
                if t["op"][i] in (0, 6):
                    continue
                tables[r][i] = _eval_node(t, Lookup(r), i, inputs, B)

    for r in range(W):
        run_phase(r, True)
    for r in range(W):
        run_phase(r, False)
    outs = [np.full((B, int(n)), np.nan) for n in t["n_out"]]
    for r in range(W):
        for i, arr, el in t["roles"][r]["outs"]:
            outs[int(arr)][:, int(el)] = ensure(r, int(i))
    return outs
