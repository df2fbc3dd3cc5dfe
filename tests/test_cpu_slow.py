"""About a minute (PZK_SLOW=0 skips it): the compile-time row proofs at registerIdentity scale.  Every one of the
2 250 656 rows of the .r1cs is evaluated in Python on witnesses the compiled program produces for a valid and
for tampered passports; a failing row must be a run-time row (rowkind 0) and the first one the reported
first_bad.  Needs artifacts/c3.r1cs(.local) and artifacts/c3.rowkind, i.e. a build where /root/reference exists."""
import json
import os

import numpy as np
import pytest

import formats
import ref as oracle_ref
from util import ROOT, u64_to_ints

pytestmark = pytest.mark.skipif(os.environ.get("PZK_SLOW") == "0", reason="PZK_SLOW=0")


def test_c3_every_row_evaluated_independently():
    from passport_zk_circuits_b200 import witness as W
    art = os.path.join(ROOT, "artifacts")
    r1_path = next((p for p in (os.path.join(art, "c3.r1cs"), os.path.join(art, "c3.r1cs.local")) if os.path.exists(p)), None)
    if r1_path is None or not os.path.exists(os.path.join(art, "c3.rowkind")):
        pytest.skip("c3.r1cs / c3.rowkind not built here")
    r1 = formats.read_r1cs(r1_path)
    kinds = np.fromfile(os.path.join(art, "c3.rowkind"), dtype=np.uint8)
    prog = oracle_ref.RefProgram(W.artifact("c3"))
    assert len(kinds) == len(r1["constraints"]) == prog.n_constraints
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "c3.json")))
    size = {d["name"]: d["size"] for d in prog.meta["inputs"]}
    obj = {k: (list(v) if isinstance(v, str) and size[k] > 1 else v) for k, v in g["cases"][0]["inputs"].items()}
    base = W.pack_inputs_fast(prog.meta, [obj])[0]
    d = {x["name"]: x for x in prog.meta["inputs"]}
    P = r1["prime"]

    def failing(w):
        bad = []
        for i, (A, B, C) in enumerate(r1["constraints"]):
            a = sum(c * w[x] for x, c in A) % P
            b = sum(c * w[x] for x, c in B) % P
            cc = sum(c * w[x] for x, c in C) % P
            if (a * b - cc) % P:
                bad.append(i)
        return bad
    n_fail = 0
    for name, k, bit in ((None, 0, 0), ("signature", 5, 3), ("dg1", 300, 0), ("encapsulatedContent", 900, 0), ("pubkey", 2, 9)):
        row = base.copy()
        if name:
            row[d[name]["offset"] + k, 0] ^= np.uint64(1 << bit)
        st, fb, wit = prog.witness(row)
        bad = failing(u64_to_ints(wit))
        assert all(kinds[i] == 0 for i in bad), (name, [(i, int(kinds[i])) for i in bad if kinds[i]][:5])
        assert fb == (bad[0] if bad else -1) and bool(st & 2) == bool(bad)
        n_fail += bool(bad)
    assert n_fail == 4
