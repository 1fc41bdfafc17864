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
