"""Shared helpers for the parity tests."""
import hashlib
import os

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ints_to_u64(vals):
    raw = b"".join(int(v % P).to_bytes(32, "little") for v in vals)
    return np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4).copy()


def u64_to_ints(arr):
    raw = np.ascontiguousarray(arr, dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def random_inputs(meta, B, seed, field_bits=253):
    """uint64 [B, n_inputs, 4] honouring each input's declared width."""
    rng = np.random.default_rng(seed)
    total = sum(d["size"] for d in meta["inputs"])
    inp = np.zeros((B, total, 4), dtype=np.uint64)
    for d in meta["inputs"]:
        sl = slice(d["offset"], d["offset"] + d["size"])
        if d["bits"]:
            hi = (1 << d["bits"]) - 1
            inp[:, sl, 0] = rng.integers(0, hi, size=(B, d["size"]), dtype=np.uint64, endpoint=True)
        else:
            inp[:, sl, :3] = rng.integers(0, (1 << 64) - 1, size=(B, d["size"], 3), dtype=np.uint64, endpoint=True)
            inp[:, sl, 3] = rng.integers(0, (1 << (field_bits - 192)) - 1, size=(B, d["size"]), dtype=np.uint64,
                                         endpoint=True)
    return inp


def input_dict(meta, row):
    """uint64 [n_inputs, 4] -> {name: nested list of ints} for the Python oracle."""
    vals = u64_to_ints(row)
    out = {}
    for d in meta["inputs"]:
        out[d["name"]] = vals[d["offset"]:d["offset"] + d["size"]]
    return out


def witness_digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr, dtype=np.uint64).tobytes()).hexdigest()
