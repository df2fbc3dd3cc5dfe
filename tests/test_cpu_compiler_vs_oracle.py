"""CPU parity: the Python oracle (direct interpreter of the circom sources) against the compiled
program evaluated by the plain-C oracle evaluator, signal by signal, mapped BY NAME through the
.sym file.  This validates the product compiler without a GPU."""
import os

import numpy as np
import pytest

import circom_oracle as co
import formats
import ref as oracle_ref
from conftest import REFERENCE, has_reference
from util import ROOT, input_dict, ints_to_u64, random_inputs, u64_to_ints

OWN = {"t_mix": "tests/circuits/mix.circom", "t_bigdiv": "tests/circuits/bigdiv.circom",
       "t_earlyret": "tests/circuits/earlyret.circom", "t_modinv": "tests/circuits/modinv.circom",
       "t_muladd": "tests/circuits/muladd.circom"}
REF_SMALL = ["poseidon2", "sha256_1", "smt80", "babyjub"]


def compare(prog_prefix, circom_main, inputs_u64, expect_ok=True):
    prog = oracle_ref.RefProgram(prog_prefix + ".pzkp")
    sym = formats.read_sym(prog_prefix + ".sym")
    circ = co.Circuit(circom_main)
    names = circ.signal_names()
    assert len(names) + 1 == prog.n_wires
    for row in inputs_u64:
        st, fb, wit = prog.witness(row)
        w = circ.calculate_witness(input_dict(prog.meta, row), check=expect_ok)
        got = u64_to_ints(wit)
        assert got[0] == 1
        bad = [(n, got[sym[n]], v) for n, v in zip(names, w) if got[sym[n]] != v]
        assert not bad, bad[:5]
        if expect_ok:
            assert st == 0 and fb == -1
    return prog


def mix_inputs(meta, B, seed):
    inp = random_inputs(meta, B, seed)
    return inp


def test_own_mix(artifacts_dir):
    prog = oracle_ref.RefProgram(os.path.join(artifacts_dir, "t_mix.pzkp"))
    inp = mix_inputs(prog.meta, 6, 11)
    # make x == y in one row to hit the zero branch of the inversion
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[0, d["x"]["offset"]] = inp[0, d["y"]["offset"]]
    compare(os.path.join(artifacts_dir, "t_mix"), os.path.join(ROOT, OWN["t_mix"]), inp)


def test_own_bigdiv_intrinsic_equals_function(artifacts_dir):
    prog = oracle_ref.RefProgram(os.path.join(artifacts_dir, "t_bigdiv.pzkp"))
    inp = random_inputs(prog.meta, 8, 5)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[:, d["b"]["offset"] + 1, 0] |= np.uint64(1)        # top limb of the divisor non-zero
    inp[1, d["b"]["offset"] + 1, 0] = np.uint64(1)          # tiny top limb: many correction steps
    inp[2, d["a"]["offset"]:d["a"]["offset"] + 3, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    inp[2, d["b"]["offset"] + 1, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    # the quotient must fit m+1 = 2 limbs: a[2] < b[1] is not required, q < 2^128 always holds here
    compare(os.path.join(artifacts_dir, "t_bigdiv"), os.path.join(ROOT, OWN["t_bigdiv"]), inp)


def earlyret_inputs(meta, B, seed):
    inp = random_inputs(meta, B, seed)
    d = {x["name"]: x for x in meta["inputs"]}
    inp[:, d["b"]["offset"] + 1, 0] |= np.uint64(1)
    inp[0, d["x"]["offset"]] = 0                            # inv_or_zero leaves early
    for i in range(1, 6):                                   # first_set returns 0, 1, 2, 3, 4
        inp[i, d["v"]["offset"]:d["v"]["offset"] + (i - 1), 0] = 0
    inp[7, d["v"]["offset"]:d["v"]["offset"] + 4, 0] = 0
    return inp, d


def test_own_early_returns_under_data_dependent_conditions(artifacts_dir):
    """Functions that `return` inside signal-dependent ifs (the shape of short_div_norm / mod_inv /
    long_sub_mod, /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:290-312, 430-446,
    516-522) are if-converted: every path's value must survive with its path condition."""
    prog = oracle_ref.RefProgram(os.path.join(artifacts_dir, "t_earlyret.pzkp"))
    inp, d = earlyret_inputs(prog.meta, 48, 3)
    compare(os.path.join(artifacts_dir, "t_earlyret"), os.path.join(ROOT, OWN["t_earlyret"]), inp)
    # a dividend limb that only fits 64 bits at run time is narrowed under an assertion
    row = inp[9].copy()
    row[d["a"]["offset"] + 2, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    row[d["c"]["offset"], 0] = 1
    st, _, _ = prog.witness(row)
    assert st & 1


def test_own_fused_multiply_add_records(artifacts_dir):
    """A product whose only reader is a sum and that is no signal is computed inside the sum's record
    (pzk_program.h, PZK_F_MULADD / PZK_Z_MULADD): every sign combination, field and wide-integer class."""
    from passport_zk_circuits_b200 import witness as W
    prefix = os.path.join(artifacts_dir, "t_muladd")
    hist = W.program_histogram(prefix + ".pzkp")["records"]
    assert hist.get("F_MULADD", 0) >= 5 and hist.get("Z_MULADD", 0) >= 6, hist
    prog = oracle_ref.RefProgram(prefix + ".pzkp")
    assert prog.meta["stats"]["fused_muladd_wire_products"] >= 3     # fp, zp, zq: second results (PZK_FLAG_DST2)
    assert prog.meta["stats"]["z_u_operands"] >= 6                   # 64-bit factors read as U words in place
    inp = random_inputs(prog.meta, 12, 21)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[0, d["c"]["offset"], 0] = 0                                   # c - a b < 0 with c = 0
    inp[1, d["a"]["offset"], 0] = np.uint64(0xFFFFFFFFFFFFFFFF)       # the largest 64 x 64 product
    inp[1, d["b"]["offset"], 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    inp[2, d["a"]["offset"], 0] = 0                                   # a b = 0
    inp[3, d["x"]["offset"]] = 0                                      # x y = 0 in the field
    compare(prefix, os.path.join(ROOT, OWN["t_muladd"]), inp)


def modinv_inputs(meta, B, seed):
    inp = random_inputs(meta, B, seed)
    inp[0, :, 0] = 0                                                                # a == 0
    inp[1, 0, 0] = np.uint64(0xFFFFFFFFFFFFFFFF); inp[1, 1, 0] = np.uint64(33554431)  # a == p
    inp[2, 0, 0] = 1; inp[2, 1, 0] = 0                                              # a == 1
    inp[3, :, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)                                    # a == 2^128 - 1 >= p
    return inp


def test_own_modinv_intrinsic_equals_function(artifacts_dir, tmp_path):
    """A function named mod_inv with a constant prime modulus (bigIntFunc.circom:430-465 in the reference)
    becomes one MODINV record; same signals as the interpreted square-and-multiply, and as the program
    compiled with the intrinsic switched off."""
    prog = oracle_ref.RefProgram(os.path.join(artifacts_dir, "t_modinv.pzkp"))
    assert prog.meta["stats"]["modinv"] == 1
    inp = modinv_inputs(prog.meta, 24, 3)
    compare(os.path.join(artifacts_dir, "t_modinv"), os.path.join(ROOT, OWN["t_modinv"]), inp)
    from passport_zk_circuits_b200 import witness as W
    plain = W.compile_circuit(os.path.join(ROOT, OWN["t_modinv"]), str(tmp_path / "plain"), {"a": 64}, intrinsics=False)
    unrolled = oracle_ref.RefProgram(plain)
    assert unrolled.meta["stats"]["modinv"] == 0
    for row in inp:
        a, b = prog.witness(row), unrolled.witness(row)
        assert a[:2] == b[:2] == (0, -1) and np.array_equal(a[2], b[2])


@pytest.mark.skipif(not has_reference(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("name", REF_SMALL)
def test_reference_small(artifacts_dir, name):
    prefix = os.path.join(artifacts_dir, name)
    main = os.path.join(artifacts_dir, "_mains", name + ".circom")
    prog = oracle_ref.RefProgram(prefix + ".pzkp")
    if name == "smt80":
        from passport_zk_circuits_b200.poseidon import poseidon
        rows = []
        for key in (12345, 2**200 + 17):
            vals = {"root": [poseidon([key, key, 1])], "leaf": [key], "key": [key], "siblings": [0] * 80}
            flat = []
            for d in prog.meta["inputs"]:
                flat += vals[d["name"]]
            rows.append(ints_to_u64(flat))
        compare(prefix, main, np.stack(rows))
        # a wrong root still yields a witness (isVerified = 0) and all constraints hold
        bad = rows[0].copy()
        d = {x["name"]: x for x in prog.meta["inputs"]}
        bad[d["root"]["offset"], 0] ^= np.uint64(1)
        p = compare(prefix, main, np.stack([bad]))
        st, fb, wit = p.witness(bad)
        assert u64_to_ints(wit)[1] == 0
    elif name == "babyjub":
        inp = random_inputs(prog.meta, 2, 3, field_bits=248)
        compare(prefix, main, inp)
    else:
        compare(prefix, main, random_inputs(prog.meta, 2, 9))


def _failing_rows(r1cs, witness):
    p = r1cs["prime"]
    bad = []
    for i, (A, B, C) in enumerate(r1cs["constraints"]):
        a = sum(c * witness[w] for w, c in A) % p
        b = sum(c * witness[w] for w, c in B) % p
        cc = sum(c * witness[w] for w, c in C) % p
        if (a * b - cc) % p != 0:
            bad.append(i)
    return bad


PROOF_CASES = ["t_mix", "t_earlyret", "t_bigdiv", "t_modinv", "t_muladd", "poseidon2", "sha256_1", "babyjub", "smt80"]


@pytest.mark.parametrize("name", PROOF_CASES)
def test_statically_discharged_rows_never_fail(artifacts_dir, name):
    """Soundness of the compile-time row proofs (alias / truth table / symbolic expansion, <prefix>.rowkind):
    on arbitrary in-range inputs - most of which violate plenty of constraints - every row of the .r1cs is
    evaluated independently in Python on the witness the compiled program produces; a row that fails must be
    one the program checks at run time, and the first of them must be what the program reports."""
    prefix = os.path.join(artifacts_dir, name)
    if not os.path.exists(prefix + ".rowkind"):
        pytest.skip("artifact not built here")
    prog = oracle_ref.RefProgram(prefix + ".pzkp")
    kinds = np.fromfile(prefix + ".rowkind", dtype=np.uint8)
    r1cs = formats.read_r1cs(prefix + ".r1cs")
    assert len(kinds) == len(r1cs["constraints"]) == prog.n_constraints
    st_ = prog.meta["stats"]
    assert int((kinds == 1).sum()) == st_["static_rows"] and int((kinds == 2).sum()) == st_["table_rows"]
    assert int((kinds == 3).sum()) == st_["symbolic_rows"] and int((kinds == 4).sum()) == 0
    assert int((kinds == 5).sum()) == st_["view_rows"]
    n_lanes = 3 if prog.n_constraints > 50000 else 12
    inp = random_inputs(prog.meta, n_lanes, 77, field_bits=253)
    seen_fail = 0
    for row in inp:
        st, fb, wit = prog.witness(row)
        bad = _failing_rows(r1cs, u64_to_ints(wit))
        assert all(kinds[i] == 0 for i in bad), [(i, int(kinds[i])) for i in bad if kinds[i]][:5]
        assert fb == (bad[0] if bad else -1)
        assert bool(st & 2) == bool(bad)
        seen_fail += bool(bad)
    if name == "smt80":
        assert seen_fail > 0      # arbitrary siblings really do break constraints


@pytest.mark.skipif(not has_reference(), reason="/root/reference is not mounted here")
def test_static_proofs_hold_on_tampered_query_inputs(artifacts_dir):
    """queryIdentity(80): a valid input and three tampered ones (identity-tree root, a DG1 bit, the secret key).
    Every row of the .r1cs is evaluated in Python on the produced witness: failing rows must be run-time rows,
    the first of them the reported first_bad."""
    import json
    from passport_zk_circuits_b200 import witness as W
    prefix = os.path.join(artifacts_dir, "query80")
    prog = oracle_ref.RefProgram(prefix + ".pzkp")
    kinds = np.fromfile(prefix + ".rowkind", dtype=np.uint8)
    r1cs = formats.read_r1cs(prefix + ".r1cs")
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "query80.json")))
    size = {d["name"]: d["size"] for d in prog.meta["inputs"]}
    obj = {k: (list(v) if isinstance(v, str) and size[k] > 1 else v) for k, v in g["cases"][0]["inputs"].items()}
    base = W.pack_inputs_fast(prog.meta, [obj])[0]
    d = {x["name"]: x for x in prog.meta["inputs"]}
    rows = [base]
    for name, k, bit in (("idStateRoot", 0, 0), ("dg1", 200, 0), ("skIdentity", 0, 3)):
        r = base.copy()
        r[d[name]["offset"] + k, 0] ^= np.uint64(1 << bit)
        rows.append(r)
    n_bad = 0
    for i, row in enumerate(rows):
        st, fb, wit = prog.witness(row)
        bad = _failing_rows(r1cs, u64_to_ints(wit))
        assert all(kinds[j] == 0 for j in bad), (i, [(j, int(kinds[j])) for j in bad if kinds[j]][:5])
        assert fb == (bad[0] if bad else -1) and bool(st & 2) == bool(bad)
        n_bad += bool(bad)
    assert n_bad >= 2 and not _failing_rows(r1cs, u64_to_ints(prog.witness(base)[2]))


@pytest.mark.skipif(not has_reference(), reason="/root/reference is not mounted here")
def test_reference_p256_doubling(artifacts_dir):
    """EllipticCurveDouble over P-256 (curve.circom:281-313), the building block of the ECDSA variant: the
    compiled program - MODINV and BIGDIV intrinsics, predicated returns of short_div_norm / long_sub_mod - equals
    the Python interpreter of the reference's sources on all 10 065 signals, for points on the curve (all
    constraints hold, result = 2P) and for an arbitrary input (same signals, constraints fail in both)."""
    from passport_zk_circuits_b200.passports import P256_G, _ec_add, _ec_mul, chunks_le
    prefix = os.path.join(artifacts_dir, "p256dbl")
    main = os.path.join(artifacts_dir, "_mains", "p256dbl.circom")
    prog = oracle_ref.RefProgram(prefix + ".pzkp")
    assert prog.meta["stats"]["modinv"] == 1 and prog.meta["stats"]["bigdiv"] >= 1
    pts = [_ec_mul(0xC0FFEE, P256_G)]
    rows = [ints_to_u64(chunks_le(x, 64, 4) + chunks_le(y, 64, 4)) for x, y in pts]
    compare(prefix, main, np.stack(rows))
    for (x, y), row in zip(pts, rows):
        st, fb, wit = prog.witness(row)
        w = u64_to_ints(wit)
        x2, y2 = _ec_add((x, y), (x, y))
        assert w[1:9] == chunks_le(x2, 64, 4) + chunks_le(y2, 64, 4)
    off = ints_to_u64(chunks_le(pts[0][0], 64, 4) + chunks_le(pts[0][1] ^ 5, 64, 4))
    compare(prefix, main, np.stack([off]), expect_ok=False)
    assert prog.witness(off)[0] & 2


@pytest.mark.skipif(not has_reference(), reason="/root/reference is not mounted here")
def test_translated_functions_equal_the_tree_walker():
    """oracle/circom_oracle.py runs circom `function` bodies as translated Python (FunctionTranslator) instead of walking
    their syntax trees; both execution modes of the reference's own big-integer hint functions
    (/root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom) must return identical values on random operands,
    including the data-dependent early returns of short_div_norm / long_gt and the 380-round mod_inv."""
    import random
    import circom_oracle as co
    main = os.path.join(ROOT, "artifacts", "_mains", "p256dbl.circom")
    fast, slow = co.Circuit(main), co.Circuit(main, translate_functions=False)
    rng = random.Random(12)
    P256 = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF

    def limbs(x, k, pad=200):
        return [(x >> (64 * i)) & (2 ** 64 - 1) for i in range(k)] + [0] * (pad - k)
    cases = []
    for _ in range(6):
        a, b = rng.randrange(P256), rng.randrange(1, P256)
        cases += [("long_add_mod", [64, 4, limbs(a, 4), limbs(b, 4), limbs(P256, 4)]),
                  ("long_sub_mod", [64, 4, limbs(a, 4), limbs(b, 4), limbs(P256, 4)]),
                  ("prod_mod", [64, 4, limbs(a, 4), limbs(b, 4), limbs(P256, 4)]),
                  ("prod", [64, 4, limbs(a, 4), limbs(b, 4)]),
                  ("long_gt", [64, 4, limbs(a, 4), limbs(b, 4)]),
                  ("long_div", [64, 4, 4, limbs(a * b, 8), limbs(P256, 4)]),
                  ("long_scalar_mult", [64, 4, a & (2 ** 64 - 1), limbs(b, 4)]),
                  ("log_ceil", [rng.randrange(1, 2 ** 40)]), ("div_ceil", [rng.randrange(1000), rng.randrange(1, 50)]),
                  ("is_negative", [rng.choice([a, co.P - 5, 5])])]
    cases += [("mod_inv", [64, 4, limbs(rng.randrange(1, P256), 4), limbs(P256, 4)]) for _ in range(2)]
    cases += [("mod_inv", [64, 4, limbs(0, 4), limbs(P256, 4)])]
    translated = set()
    for name, args in cases:
        got = fast.translator.call(name, [co._copy(x) for x in args])
        want = slow.interpret_function(name, [co._copy(x) for x in args])
        assert got == want, name
        if fast.translator.fns.get(name):
            translated.add(name)
    assert {"mod_inv", "long_div", "prod", "long_sub_mod", "long_gt"} <= translated   # really the fast path
    # and a whole circuit both ways: one P-256 doubling, every signal
    G = [0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
         0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5]
    inp = {"in": [[str((G[i] >> (64 * j)) & (2 ** 64 - 1)) for j in range(4)] for i in range(2)]}
    assert fast.calculate_witness(inp, check=True) == slow.calculate_witness(inp, check=True)
