"""GPU tests of the round-2 boundary additions: witness digest, stream-async / multi-device calls, the R1CS
handle and the device-resident hand-off from the evaluator to the `wtns check` kernel
(/root/reference/test/automatisationTest.js:40-51 as two kernels on one device)."""
import os

import numpy as np
import pytest

import formats
import ref as oracle_ref
from util import ROOT, random_inputs

pytestmark = pytest.mark.gpu

from passport_zk_circuits_b200 import witness as W  # noqa: E402


def oracle_digests(ref, inp):
    out = np.zeros((len(inp), 4), dtype=np.uint64)
    for b in range(len(inp)):
        _, _, wit = ref.witness(inp[b], want_witness=True)
        out[b] = W.witness_digest(wit)
    return out


@pytest.mark.parametrize("name,B,tile", [("t_mix", 301, 128), ("t_mix", 40, 0), ("smt80", 70, 0), ("poseidon2", 33, 0),
                                         ("sha256_1", 20, 0), ("babyjub", 9, 0), ("t_bigdiv", 50, 0)])
def test_witness_digest_equals_the_digest_of_the_oracle_witness(name, B, tile):
    """every wire of every lane is folded on the device (no witness is exported) and must give the checksum of
    the oracle's full witness; tampered inputs change it"""
    prog = W.artifact(name)
    ref = oracle_ref.RefProgram(prog)
    calc = W.WitnessCalculator(prog, 0)
    if tile:
        calc.set_tile_lanes(tile)
    inp = random_inputs(ref.meta, B, 17, field_bits=248)
    calc.set_digest(True)
    calc.upload(inp)
    calc.run(True)
    got = calc.download_digest()
    want = oracle_digests(ref, inp)
    assert np.array_equal(got, want), np.nonzero((got != want).any(axis=1))[0][:8]
    # the packed streamed call returns the same digests, and so does a second run on the same handle
    rec, bad = calc.pack(inp, on_range="mask")
    res = calc.calculateWitnessBatchPacked(rec, range_mask=bad, digest=True)
    assert np.array_equal(res.digest, want)
    plain = calc.calculateWitnessBatch(inp)
    assert np.array_equal(res.status, plain.status) and np.array_equal(res.public, plain.public)
    # digest off: the lean path (optional stores skipped) gives the same verdicts
    calc.set_digest(False)
    lean = calc.calculateWitnessBatchPacked(rec, range_mask=bad)
    assert np.array_equal(lean.status, plain.status) and np.array_equal(lean.public, plain.public)
    calc.close()


def test_async_call_and_single_process_multi_device():
    prog = W.artifact("t_mix")
    ref = oracle_ref.RefProgram(prog)
    inp = random_inputs(ref.meta, 1000, 23)
    a = W.WitnessCalculator(prog, 0)
    ndev = W.lib().pzk_device_count()
    b = W.WitnessCalculator(prog, 1 if ndev > 1 else 0)   # a second handle: another device when there is one
    for c in (a, b):
        c.set_digest(True)
        c.set_tile_lanes(256)                                # several tiles per share: exercises the copy overlap
    rec = a.pack(inp)
    one = a.calculateWitnessBatchPacked(rec, digest=True)
    two = W.witness_batch_packed_multi([a, b], rec, digest=True)
    for f in ("status", "first_bad", "public", "digest"):
        assert np.array_equal(getattr(one, f), getattr(two, f)), f
    # async: nothing is read before pzk_sync
    st = np.zeros(1000, dtype=np.uint32)
    fb = np.zeros(1000, dtype=np.int64)
    pub = np.zeros((1000, a.n_public, 4), dtype=np.uint64)
    dg = np.zeros((1000, 4), dtype=np.uint64)
    assert a._L.pzk_witness_batch_packed_async(a._h, rec.ctypes.data, 1000, st.ctypes.data, fb.ctypes.data,
                                               pub.ctypes.data, dg.ctypes.data) == 0
    a.sync()
    fb[(st & W.STATUS_CONSTRAINT) == 0] = -1
    assert np.array_equal(st, one.status) and np.array_equal(fb, one.first_bad)
    assert np.array_equal(pub, one.public) and np.array_equal(dg, one.digest)
    want = oracle_digests(ref, inp[:25])
    assert np.array_equal(one.digest[:25], want)
    a.close()
    b.close()


def test_r1cs_handle_and_device_resident_hand_off():
    """pzk_r1cs_open once, then: explicit witnesses from the host, a .wtns image, and the evaluator's own witnesses
    handed over on the device; verdicts and first failing rows must agree with the Python `wtns check` loop."""
    prefix = os.path.join(ROOT, "artifacts", "smt80")
    ref = oracle_ref.RefProgram(prefix + ".pzkp")
    from test_gpu_parity import smt_inputs
    keys = [4242 + 13 * i + (i << 100) for i in range(200)]
    inp = smt_inputs(ref.meta, keys)
    calc = W.WitnessCalculator(prefix + ".pzkp", 0)
    r = W.R1cs(prefix + ".r1cs", 0)
    assert (r.n_wires, r.n_constraints) == (calc.n_wires, calc.n_constraints)
    res = calc.calculateWitnessBatch(inp, export_lanes=range(200))
    wit = res.witnesses.copy()
    bad_lanes = {0: 5, 31: 777, 32: 90000, 150: 12345, 199: 3}
    for lane, wire in bad_lanes.items():
        wit[lane, wire, 0] ^= np.uint64(1)
    r1 = formats.read_r1cs(prefix + ".r1cs")
    for rep in range(2):                           # the handle is reusable
        ok, first, ms = r.check(wit)
        for lane in range(200):
            if lane in bad_lanes:
                w = [int.from_bytes(wit[lane, i].tobytes(), "little") for i in range(calc.n_wires)]
                assert (bool(ok[lane]), int(first[lane])) == formats.wtns_check(r1, w), lane
            else:
                assert ok[lane] and first[lane] == -1
    good = [int.from_bytes(res.witnesses[7, i].tobytes(), "little") for i in range(calc.n_wires)]
    assert r.check_wtns(formats.write_wtns(good)) == (True, -1)
    # hand-off: evaluate on the device, check every row of the .r1cs on the device
    calc.upload(inp)
    lanes = [0, 1, 5, 31, 32, 33, 64, 100, 199]
    ok, first, t_eval, t_check = r.check_circuit(calc, lanes)
    assert ok.all() and (first == -1).all() and t_eval > 0 and t_check > 0
    ok, first, _, _ = r.check_circuit(calc, range(200))       # lane-parallel export path
    assert ok.all()
    # a lane whose inputs violate a constraint: the fused rows and the stand-alone kernel name the same row
    brk = smt_inputs(ref.meta, keys[:40], break_lane=11)
    d = {x["name"]: x for x in ref.meta["inputs"]}
    brk[17, d["siblings"]["offset"] + 79, 0] = 5                 # last sibling must be zero
    calc.upload(brk)
    calc.run(True)
    fused = calc.download()
    ok, first, _, _ = r.check_circuit(calc, range(40))
    assert np.array_equal(ok, (fused.status & W.STATUS_CONSTRAINT) == 0)
    assert np.array_equal(first[~ok], fused.first_bad[~ok])
    assert not ok[17]
    # an .r1cs that does not belong to the program is refused by its wire count
    other = W.R1cs(os.path.join(ROOT, "artifacts", "t_mix.r1cs"), 0)
    with pytest.raises(W.PzkError, match="Invalid witness length"):
        other.check_circuit(calc, [0])
    with pytest.raises(W.PzkError, match="Invalid witness length"):
        other.check(wit[:2])
    other.close()
    r.close()
    calc.close()


def test_c3_digest_and_hand_off():
    """registerIdentity: digests of 6 lanes against the oracle's 2.25 M-wire witnesses; the same lanes handed to the
    stream kernel on the device (every one of the 2 250 656 rows evaluated)."""
    from passport_zk_circuits_b200.passports import C3, PassportFactory
    prog = W.artifact("c3")
    ref = oracle_ref.RefProgram(prog)
    calc = W.WitnessCalculator(prog, 0)
    fac = PassportFactory(C3, seed=11, n_sig_keys=2, n_aa_keys=2)
    B = 300
    uniq = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(6)])
    inp = np.tile(uniq, (B // 6, 1, 1))
    d = {x["name"]: x for x in calc.meta["inputs"]}
    inp[4, d["signature"]["offset"] + 1, 0] ^= np.uint64(1 << 9)       # invalid signature: still a full witness
    calc.set_digest(True)
    res = calc.calculateWitnessBatchPacked(calc.pack(inp), digest=True)
    want = oracle_digests(ref, inp[:6])
    assert np.array_equal(res.digest[:6], want)
    assert np.array_equal(res.digest[6:12][[0, 1, 2, 3, 5]], want[[0, 1, 2, 3, 5]])
    assert (res.status[:4] == 0).all() and res.status[4] & W.STATUS_CONSTRAINT
    r = W.R1cs(W.artifact_r1cs("c3"), 0)
    calc.set_digest(False)
    calc.upload(inp)
    lanes = list(range(0, 96))
    ok, first, t_eval, t_check = r.check_circuit(calc, lanes)
    assert np.array_equal(ok, (res.status[:96] & W.STATUS_CONSTRAINT) == 0)
    assert np.array_equal(first[~ok], res.first_bad[:96][~ok])
    r.close()
    calc.close()


def test_external_sym_renumbers_the_witness(tmp_path):
    """pzk_circuit_open_ex: the wires follow an EXTERNAL .sym (what `circom --O2` would have written): a permuted
    numbering with a third of the signals optimised away (-1) and aliases merged into one wire.  The witness,
    the .wtns bytes and the digest must be those of the own-layout witness re-indexed BY NAME."""
    prefix = os.path.join(ROOT, "artifacts", "smt80")
    own = formats.read_sym(prefix + ".sym")                      # name -> own wire
    calc0 = W.WitnessCalculator(prefix + ".pzkp", 0)
    n_pub = calc0.n_public
    rng = np.random.default_rng(9)
    names = sorted(own, key=lambda k: own[k])
    keep = [nm for nm in names if own[nm] <= n_pub or rng.random() > 0.33]      # public wires stay, in place
    inner = [nm for nm in keep if own[nm] > n_pub]
    perm = rng.permutation(len(inner))
    ext = {nm: own[nm] for nm in keep if own[nm] <= n_pub}
    for j, nm in zip(perm, inner):
        ext[nm] = n_pub + 1 + int(j)
    # merge: two kept signals that are aliases of each other in the circuit share one external wire
    # (smt80: hashers[i].in[..] aliases) - emulate with the dropped list: a dropped name pointing at a kept wire
    dropped = [nm for nm in names if nm not in ext]
    sym_path = tmp_path / "ext.sym"
    with open(sym_path, "w") as f:
        for i, nm in enumerate(names):
            f.write(f"{i + 1},{ext.get(nm, -1)},0,{nm}\n")
    calc = W.WitnessCalculator(prefix + ".pzkp", 0, external_sym=str(sym_path))
    assert calc.n_wires == len(ext) + 1
    ref = oracle_ref.RefProgram(prefix + ".pzkp")
    from test_gpu_parity import smt_inputs
    inp = smt_inputs(ref.meta, [99 + 7 * i for i in range(20)])
    a = calc0.calculateWitnessBatch(inp, export_lanes=range(20))
    b = calc.calculateWitnessBatch(inp, export_lanes=range(20))
    want = np.zeros((20, calc.n_wires, 4), dtype=np.uint64)
    want[:, 0, 0] = 1
    for nm, w in ext.items():
        want[:, w] = a.witnesses[:, own[nm]]
    assert np.array_equal(b.witnesses, want)
    assert np.array_equal(a.status, b.status) and np.array_equal(a.public, b.public)
    calc.set_digest(True)
    calc.upload(inp)
    calc.run(True)
    dg = calc.download_digest()
    for lane in range(20):
        assert np.array_equal(dg[lane], W.witness_digest(want[lane]))
    blob = calc.calculateWTNSBin({k: [int.from_bytes(inp[3, d["offset"] + j].tobytes(), "little") for j in range(d["size"])]
                                  if d["size"] > 1 else int.from_bytes(inp[3, d["offset"]].tobytes(), "little")
                                  for d in ref.meta["inputs"] for k in [d["name"]]})
    assert blob == formats.write_wtns([int.from_bytes(want[3, i].tobytes(), "little") for i in range(calc.n_wires)])
    # an external .sym that names a signal the program does not have is refused
    with open(sym_path, "a") as f:
        f.write("999999,5,0,main.no_such_signal\n")
    with pytest.raises(W.PzkError, match="not a signal of this program"):
        W.WitnessCalculator(prefix + ".pzkp", 0, external_sym=str(sym_path))
    calc.close()
    calc0.close()


@pytest.mark.parametrize("name", ["t_mix", "smt80", "query80"])
def test_o1_layout_on_the_device(name):
    """The O1-simplified system (pzk.h PZK_COMPILE_EMIT_O1) end to end on the device: the program opened with
    <name>.O1.sym produces witnesses in the O1 numbering, every row of <name>.O1.r1cs holds on them - through explicit
    witnesses, a .wtns image and the device-resident hand-off - and tampered lanes fail in both systems."""
    prefix = os.path.join(ROOT, "artifacts", name)
    ref = oracle_ref.RefProgram(prefix + ".pzkp")
    if name == "smt80":
        from test_gpu_parity import smt_inputs
        inp = smt_inputs(ref.meta, [31 + 5 * i for i in range(40)])
        d = {x["name"]: x for x in ref.meta["inputs"]}
        inp[9, d["siblings"]["offset"] + 79, 0] = 5                 # last sibling must be zero
    elif name == "query80":
        from passport_zk_circuits_b200.query_inputs import make_query_input
        objs = [make_query_input(i, seed=9, selector=39) for i in range(40)]
        objs[9]["idStateRoot"] = str(int(objs[9]["idStateRoot"]) ^ 1)
        inp = W.pack_inputs_fast(ref.meta, objs)
    else:
        inp = random_inputs(ref.meta, 40, 3)
        d = {x["name"]: x for x in ref.meta["inputs"]}
        inp[9, d["x"]["offset"], 0] ^= np.uint64(1)
    own = W.WitnessCalculator(prefix + ".pzkp", 0)
    o1 = W.WitnessCalculator(prefix + ".pzkp", 0, external_sym=prefix + ".O1.sym", program_sym=prefix + ".sym")
    r1 = W.R1cs(prefix + ".O1.r1cs", 0)
    assert o1.n_wires == r1.n_wires < own.n_wires
    a = own.calculateWitnessBatch(inp)
    b = o1.calculateWitnessBatch(inp, export_lanes=range(40))
    assert np.array_equal(a.status, b.status) and np.array_equal(a.public, b.public)
    ok, first, _ = r1.check(b.witnesses)
    assert np.array_equal(ok, (a.status & W.STATUS_CONSTRAINT) == 0)
    o1.upload(inp)
    ok2, _, _, _ = r1.check_circuit(o1, range(40))
    assert np.array_equal(ok2, ok)
    o1.set_digest(True)
    o1.run(True)
    dg = o1.download_digest()
    for lane in (0, 9, 39):
        assert np.array_equal(dg[lane], W.witness_digest(b.witnesses[lane]))
    for c in (own, o1):
        c.close()
    r1.close()
