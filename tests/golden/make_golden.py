"""Generates tests/golden/*.json.  Run HERE (needs /root/reference and artifacts/ built by
__graft_entry__.build()):   python tests/golden/make_golden.py

For each case the reference's circom sources are interpreted by the Python oracle
(oracle/circom_oracle.py); the resulting signal values are arranged in the product's wire order
BY NAME through the compiler's .sym file, and the fixture stores: the inputs, the SHA-256 of the
.wtns data section (32-byte little-endian canonical values, wire 0 = 1), the public signals and
a few thousand sampled (wire, value) pairs.  The fixtures travel to the GPU box; the reference
does not."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import circom_oracle as co  # noqa: E402
import formats  # noqa: E402
from passport_zk_circuits_b200.passports import C3, PassportFactory  # noqa: E402
from passport_zk_circuits_b200.poseidon import poseidon  # noqa: E402

ART = os.path.join(ROOT, "artifacts")
OUT = os.path.dirname(os.path.abspath(__file__))


def golden(name, sym_path, inputs_list, n_pub, n_samples=4096):
    circ = co.Circuit(os.path.join(ART, "_mains", name + ".circom"))
    names = circ.signal_names()
    sym = formats.read_sym(sym_path)
    cases = []
    for inputs in inputs_list:
        w = circ.calculate_witness(inputs, check=True)
        wires = [0] * (len(names) + 1)
        wires[0] = 1
        for n, v in zip(names, w):
            wires[sym[n]] = v
        h = hashlib.sha256()
        for v in wires:
            h.update(v.to_bytes(32, "little"))
        step = max(1, len(wires) // n_samples)
        cases.append({
            "inputs": {k: ("".join(v) if (isinstance(v, list) and v and all(x in ("0", "1") for x in v) and len(v) > 64)
                           else v) for k, v in inputs.items()},
            "n_wires": len(wires),
            "n_constraints": circ.constraint_count,
            "wtns_data_sha256": h.hexdigest(),
            "public": [str(wires[i]) for i in range(1, 1 + n_pub)],
            "samples": [[i, str(wires[i])] for i in range(0, len(wires), step)],
        })
        print(name, "wires", len(wires), "digest", h.hexdigest()[:16])
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump({"circuit": name, "cases": cases}, f)


def main():
    from passport_zk_circuits_b200.artifacts import C4_VARIANTS as _ALL_C4
    which = sys.argv[1:] or (["poseidon2", "sha256_1", "smt80", "query80", "query80_td1", "c3"] + list(_ALL_C4))
    if "poseidon2" in which:
        golden("poseidon2", os.path.join(ART, "poseidon2.sym"), [{"in": ["1", "2"]}, {"in": ["0", str(co.P - 1)]}], 1)
    if "sha256_1" in which:
        msgs = [b"abc", b"", b"passport-zk-circuits b200 witness generator 0123456789abcdef!"[:55]]
        ins = []
        for m in msgs:
            pad = m + b"\x80" + b"\x00" * (64 - len(m) - 9) + (len(m) * 8).to_bytes(8, "big")
            ins.append({"in": [str((b >> (7 - i)) & 1) for b in pad for i in range(8)]})
        golden("sha256_1", os.path.join(ART, "sha256_1.sym"), ins, 256)
    if "smt80" in which:
        ins = []
        for key in (12345, 2 ** 200 + 17):
            ins.append({"root": str(poseidon([key, key, 1])), "leaf": str(key), "key": str(key), "siblings": ["0"] * 80})
        golden("smt80", os.path.join(ART, "smt80.sym"), ins, 2)
    from passport_zk_circuits_b200.artifacts import C4_VARIANTS
    for name, prm in C4_VARIANTS.items():
        if name in which:
            fac = PassportFactory(prm, seed=42, n_sig_keys=1, n_aa_keys=1)
            sym = os.path.join(ART, name + ".sym")
            if not os.path.exists(sym):
                sym += ".local"
            golden(name, sym, [fac.make(0).inputs], 5, n_samples=2048)
    if "query80" in which:
        from passport_zk_circuits_b200.query_inputs import make_query_input
        golden("query80", os.path.join(ART, "query80.sym"),
               [make_query_input(0, seed=7, selector=39), make_query_input(1, seed=7, selector=255)], 23)
    if "query80_td1" in which:
        from passport_zk_circuits_b200.query_inputs import make_query_input
        golden("query80_td1", os.path.join(ART, "query80_td1.sym"),
               [make_query_input(0, seed=7, selector=39, td1=True), make_query_input(1, seed=7, selector=255, td1=True)], 24)
    if "c3" in which:
        fac = PassportFactory(C3, seed=42, n_sig_keys=2, n_aa_keys=2)
        sym = os.path.join(ART, "c3.sym")
        if not os.path.exists(sym):
            sym += ".local"
        golden("c3", sym, [fac.make(0).inputs, fac.make(1).inputs], 5)


if __name__ == "__main__":
    main()
