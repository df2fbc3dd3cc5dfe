"""GPU parity for the north-star circuit registerIdentity (SHA-256 + RSA-2048),
RegisterIdentityBuilder(1,256,3,4,600,248,1,1496,3,256) of /root/reference/hardhat.config.ts:29."""
import os

import numpy as np
import pytest

import ref as oracle_ref
from util import ROOT

pytestmark = pytest.mark.gpu

from passport_zk_circuits_b200 import witness as W  # noqa: E402
from passport_zk_circuits_b200.passports import C3, PassportFactory  # noqa: E402


@pytest.fixture(scope="module")
def c3():
    prog = W.artifact("c3")
    return prog, W.WitnessCalculator(prog, device=0), oracle_ref.RefProgram(prog)


def test_c3_batch_vs_oracle(c3):
    prog, calc, ref = c3
    fac = PassportFactory(C3, seed=3, n_sig_keys=2, n_aa_keys=2)
    B = 40
    inp = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(B)])
    # negative controls: flip one signature bit / one dg1 bit
    d = {x["name"]: x for x in calc.meta["inputs"]}
    inp[5, d["signature"]["offset"] + 3, 0] ^= np.uint64(1 << 17)
    inp[9, d["dg1"]["offset"] + 100, 0] ^= np.uint64(1)
    export = [0, 5, 39]
    res = calc.calculateWitnessBatch(inp, export_lanes=export)
    for b in range(B):
        want = b in export
        st, fb, wit = ref.witness(inp[b], want_witness=True)
        assert int(res.status[b]) == st, (b, int(res.status[b]), st)
        assert int(res.first_bad[b]) == fb, (b, int(res.first_bad[b]), fb)
        assert np.array_equal(res.public[b], wit[1:1 + calc.n_public])
        if want:
            got = res.witnesses[export.index(b)]
            if not np.array_equal(got, wit):
                idx = np.nonzero((got != wit).any(axis=1))[0]
                raise AssertionError(f"lane {b}: {len(idx)} wires differ, first {idx[:5]}")
    # the same batch without witness export takes the lean path (values that never leave the
    # shared-memory operand cache are not stored to HBM): identical verdicts and public signals
    res2 = calc.calculateWitnessBatch(inp)
    assert np.array_equal(res2.status, res.status) and np.array_equal(res2.first_bad, res.first_bad)
    assert np.array_equal(res2.public, res.public)
    # packed input records (bits as bytes, limbs as u64): same results
    res3 = calc.calculateWitnessBatchPacked(calc.pack(inp))
    assert np.array_equal(res3.status, res.status) and np.array_equal(res3.first_bad, res.first_bad)
    assert np.array_equal(res3.public, res.public)
    ok = np.ones(B, dtype=bool)
    ok[[5, 9]] = False
    assert (res.status[ok] == 0).all()
    assert (res.status[~ok] & W.STATUS_CONSTRAINT).all()


def test_c3_golden_wtns_bytes(c3):
    """The .wtns data section of two fixed synthetic passports, as computed from the reference's
    circom sources by the Python oracle (tests/golden/c3.json), must come out of the GPU byte for byte."""
    import hashlib
    import json
    import os
    from util import ROOT
    prog, calc, ref = c3
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "c3.json")))
    size = {d["name"]: d["size"] for d in calc.meta["inputs"]}
    ins = [{k: (list(v) if isinstance(v, str) and size[k] > 1 else v) for k, v in c["inputs"].items()} for c in g["cases"]]
    res = calc.calculateWitnessBatch(ins, export_lanes=range(len(ins)))
    for j, case in enumerate(g["cases"]):
        assert res.status[j] == 0 and case["n_wires"] == calc.n_wires
        assert hashlib.sha256(res.witnesses[j].tobytes()).hexdigest() == case["wtns_data_sha256"]
        assert [str(v) for v in res.public_ints(j)] == case["public"]
        for wire, val in case["samples"]:
            assert int.from_bytes(res.witnesses[j][wire].tobytes(), "little") == int(val)
    # calculateWTNSBin: header + the same data section
    blob = calc.calculateWTNSBin(ins[0])
    assert blob[:4] == b"wtns" and len(blob) == 76 + 32 * calc.n_wires
    assert hashlib.sha256(blob[76:]).hexdigest() == g["cases"][0]["wtns_data_sha256"]


def _golden_inputs(calc, name):
    import json
    import os
    from util import ROOT
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    size = {d["name"]: d["size"] for d in calc.meta["inputs"]}
    ins = [{k: (list(v) if isinstance(v, str) and size[k] > 1 else v) for k, v in c["inputs"].items()} for c in g["cases"]]
    return g, ins


@pytest.mark.parametrize("name", ["c4_sig3", "c4_sig10", "c4_sig13", "c4_sig20", "c4_sig2", "c4_sig4", "c4_sig11", "c4_sig12",
                                  "c4_sig14", "c4_sig21", "c4_sig24", "c4_na", "c4_ecaa", "c4_td1"])
def test_config4_variants(name):
    """Every arm of the reference's dispatch (signatureVerification.circom:26-116, identity.circom:26-84): SHA-1 + RSA
    PKCS#1 v1.5 (SIG 3), RSA-4096 (2), RSA-3072 e = 37187 with the non-Karatsuba multiplier (4), RSA-PSS e = 3 (10),
    e = 65537 (11), salt 64 (12), SHA-384 with 1024-bit blocks (13), 3072 bits (14), ECDSA P-256 (20), brainpoolP256r1
    (21), secp224r1 in 7 x 32-bit chunks (24), a document without DG15, an EC active-authentication key, a TD1 document:
    golden .wtns bytes from the Python oracle, then a small synthetic batch against the C oracle."""
    import hashlib
    from passport_zk_circuits_b200.artifacts import C4_VARIANTS
    prog = W.artifact(name)
    calc = W.WitnessCalculator(prog, device=0)
    if os.path.exists(os.path.join(ROOT, "tests", "golden", name + ".json")):
        g, ins = _golden_inputs(calc, name)
        res = calc.calculateWitnessBatch(ins, export_lanes=range(len(ins)))
        for j, case in enumerate(g["cases"]):
            assert res.status[j] == 0
            assert hashlib.sha256(res.witnesses[j].tobytes()).hexdigest() == case["wtns_data_sha256"]
        del res
    fac = PassportFactory(C4_VARIANTS[name], seed=11, n_sig_keys=1, n_aa_keys=1)
    B = 6 if name not in ("c4_sig20", "c4_sig21", "c4_sig24") else 4
    inp = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(B)])
    d = {x["name"]: x for x in calc.meta["inputs"]}
    inp[2, d["signature"]["offset"] + 1, 0] ^= np.uint64(4)
    out = calc.calculateWitnessBatch(inp)
    ref = oracle_ref.RefProgram(prog)
    for b in range(B):
        st, fb, wit = ref.witness(inp[b], want_witness=True)
        assert int(out.status[b]) == st and int(out.first_bad[b]) == fb, (b, int(out.status[b]), st)
        assert np.array_equal(out.public[b], wit[1:1 + calc.n_public])
    assert out.status[2] != 0 and (np.delete(out.status, 2) == 0).all()
    calc.close()


@pytest.mark.parametrize("name,td1", [("query80", False), ("query80_td1", True)])
def test_config2_query_identity(name, td1):
    """queryIdentity(80) for TD3 passports and queryIdentityTD1 for identity cards (760-bit DG1, hashed document /
    personal numbers): golden public signals / .wtns bytes, all selectors, enforced SMT inclusion."""
    import hashlib
    from passport_zk_circuits_b200.query_inputs import make_query_input
    prog = W.artifact(name)
    calc = W.WitnessCalculator(prog, device=0)
    g, ins = _golden_inputs(calc, name)
    res = calc.calculateWitnessBatch(ins, export_lanes=range(len(ins)))
    for j, case in enumerate(g["cases"]):
        assert res.status[j] == 0
        assert hashlib.sha256(res.witnesses[j].tobytes()).hexdigest() == case["wtns_data_sha256"]
        assert [str(v) for v in res.public_ints(j)] == case["public"]
    B = 150
    objs = [make_query_input(i, seed=3, selector=(i * 37) & 0xFF, td1=td1) for i in range(B)]
    objs[17]["idStateRoot"] = str(int(objs[17]["idStateRoot"]) ^ 1)        # not in the tree any more
    inp = W.pack_inputs_fast(calc.meta, objs)
    out = calc.calculateWitnessBatch(inp, export_lanes=[0, 17, B - 1])
    ref = oracle_ref.RefProgram(prog)
    for b in range(B):
        st, fb, wit = ref.witness(inp[b], want_witness=True)
        assert int(out.status[b]) == st and int(out.first_bad[b]) == fb
        assert np.array_equal(out.public[b], wit[1:1 + calc.n_public])
        if b in (0, 17, B - 1):
            assert np.array_equal(out.witnesses[[0, 17, B - 1].index(b)], wit)
    assert out.status[17] & W.STATUS_CONSTRAINT and (np.delete(out.status, 17) == 0).all()
    calc.close()


def test_c3_lean_program_same_verdicts(c3):
    """`c3_lean` (definitional rows discharged at compile time, pzk.h PZK_COMPILE_STATIC_DEF_ROWS): the device
    returns the same status / first_bad / public signals as the all-rows program and the oracle."""
    prog, calc, ref = c3
    lean = W.WitnessCalculator(W.artifact("c3_lean"), device=0)
    fac = PassportFactory(C3, seed=5, n_sig_keys=2, n_aa_keys=2)
    B = 24
    inp = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(B)])
    d = {x["name"]: x for x in calc.meta["inputs"]}
    inp[3, d["signature"]["offset"] + 2, 0] ^= np.uint64(16)
    inp[8, d["dg1"]["offset"] + 99, 0] ^= np.uint64(1)
    inp[13, d["slaveMerkleRoot"]["offset"], 0] ^= np.uint64(2)     # not enforced by this circuit: still valid
    inp[20, d["dg1"]["offset"] + 5, 0] = np.uint64(3)
    a = calc.calculateWitnessBatch(inp)
    b = lean.calculateWitnessBatch(inp, export_lanes=[0, 3])
    assert np.array_equal(a.status, b.status) and np.array_equal(a.first_bad, b.first_bad)
    assert np.array_equal(a.public, b.public)
    for j, lane in enumerate((0, 3)):
        st, fb, wit = ref.witness(inp[lane], want_witness=True)
        assert int(b.status[lane]) == st and int(b.first_bad[lane]) == fb
        assert np.array_equal(b.witnesses[j], wit)
    # passportVerificationBuilder.circom:239 leaves `smtVerifier.isVerified === 1` commented out
    assert (a.status[[3, 8, 20]] != 0).all() and (np.delete(a.status, [3, 8, 20]) == 0).all()
    assert a.status[20] & W.STATUS_INPUT_RANGE
    lean.close()


def test_c3_static_proofs_against_the_full_r1cs_stream_check(c3):
    """The program checks 4 % of the rows at run time and discharges the rest by compile-time proofs.  Here the
    full witnesses of valid and tampered passports go through the stand-alone `wtns check` kernel, which
    evaluates all 2 250 656 rows of the .r1cs: verdict and first failing row must be what the program said."""
    prog, calc, ref = c3
    fac = PassportFactory(C3, seed=21, n_sig_keys=2, n_aa_keys=2)
    B = 10
    inp = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(B)])
    d = {x["name"]: x for x in calc.meta["inputs"]}
    inp[1, d["signature"]["offset"] + 9, 0] ^= np.uint64(1 << 40)
    inp[2, d["pubkey"]["offset"] + 31, 0] ^= np.uint64(1)
    inp[3, d["dg1"]["offset"] + 500, 0] ^= np.uint64(1)
    inp[4, d["encapsulatedContent"]["offset"] + 1200, 0] ^= np.uint64(1)
    inp[5, d["signedAttributes"]["offset"] + 77, 0] ^= np.uint64(1)
    inp[6, d["dg15"]["offset"] + 300, 0] ^= np.uint64(1)
    inp[7, d["skIdentity"]["offset"], 0] ^= np.uint64(1 << 17)        # any secret key is a valid witness
    res = calc.calculateWitnessBatch(inp, export_lanes=range(B))
    ok, first, _ = W.r1cs_check_batch(W.artifact_r1cs("c3"), res.witnesses)
    for b in range(B):
        assert bool(ok[b]) == (res.status[b] == 0), b
        assert int(first[b]) == int(res.first_bad[b]), (b, int(first[b]), int(res.first_bad[b]))
    assert [int(s != 0) for s in res.status] == [0, 1, 1, 1, 1, 1, 1, 0, 0, 0]


def test_real_passport_front_end_to_device():
    """process_passport.py (port of the reference's processPassport) on CMS SignedData SODs built here with
    `cryptography`: extracted parameters name the prebuilt program `c3_cms`, extracted inputs go through the
    device; valid documents pass every constraint, a document whose DG1 was altered after signing fails, and
    the device agrees with the oracle lane by lane."""
    from passport_zk_circuits_b200 import process_passport as PP
    from passport_zk_circuits_b200.artifacts import CMS_PARAMS
    from test_cpu_host import _synthetic_passport_json
    objs = []
    for seed in (1, 2, 3, 4):
        passport = _synthetic_passport_json(seed)[0]
        if seed == 3:
            raw = bytearray(bytes.fromhex(passport["dg1"]))
            raw[40] ^= 1
            # the SOD still holds the hash of the original DG1: find the shift with the original, feed the altered one
            params, inputs, _ = PP.process_passport(passport)
            from passport_zk_circuits_b200.passports import bytes_to_bits, sha_pad
            inputs["dg1"] = [str(b) for b in bytes_to_bits(sha_pad(bytes(raw), 512))]
        else:
            params, inputs, _ = PP.process_passport(passport)
        assert params == CMS_PARAMS
        objs.append(inputs)
    prog = W.artifact("c3_cms")
    calc = W.WitnessCalculator(prog, device=0)
    inp = W.pack_inputs_fast(calc.meta, objs)
    out = calc.calculateWitnessBatch(inp)
    ref = oracle_ref.RefProgram(prog)
    for b in range(len(objs)):
        st, fb, wit = ref.witness(inp[b], want_witness=True)
        assert int(out.status[b]) == st and int(out.first_bad[b]) == fb
        assert np.array_equal(out.public[b], wit[1:1 + calc.n_public])
    assert [int(s != 0) for s in out.status] == [0, 0, 1, 0]
    calc.close()
