"""GPU parity (run on the B200 box): the CUDA path, called through the C ABI, against the oracle
(oracle/ssa_ref.c evaluating the same program; that evaluator is itself pinned to the Python
interpreter of the circom sources by tests/test_cpu_compiler_vs_oracle.py and tests/golden/).
Bar: bit-exact - every wire of every exported lane, every status word, every public signal."""
import os

import numpy as np
import pytest

import ref as oracle_ref
from util import ROOT, ints_to_u64, random_inputs, u64_to_ints

pytestmark = pytest.mark.gpu

from passport_zk_circuits_b200 import witness as W  # noqa: E402


def run_and_compare(name, inp, export=None, tile=None):
    prog = W.artifact(name)
    calc = W.WitnessCalculator(prog, device=0)
    if tile:
        calc.set_tile_lanes(tile)
    B = inp.shape[0]
    export = list(range(B)) if export is None else list(export)
    res = calc.calculateWitnessBatch(inp, export_lanes=export)
    ref = oracle_ref.RefProgram(prog)
    exp_pos = {lane: j for j, lane in enumerate(export)}
    for b in range(B):
        st, fb, wit = ref.witness(inp[b], want_witness=True)
        assert int(res.status[b]) == st, (name, b, int(res.status[b]), st)
        assert int(res.first_bad[b]) == fb, (name, b, int(res.first_bad[b]), fb)
        npub = calc.n_public
        assert np.array_equal(res.public[b], wit[1:1 + npub]), (name, b, "public signals")
        if b in exp_pos:
            got = res.witnesses[exp_pos[b]]
            if not np.array_equal(got, wit):
                idx = np.nonzero((got != wit).any(axis=1))[0]
                raise AssertionError(f"{name} lane {b}: {len(idx)} wires differ, first {idx[:5]}")
    lean = calc.calculateWitnessBatch(inp)      # no export: optional global stores are skipped
    assert np.array_equal(lean.status, res.status) and np.array_equal(lean.first_bad, res.first_bad)
    assert np.array_equal(lean.public, res.public)
    # the packed path must flag exactly the lanes the unpacked path flags (values too wide for their packed
    # section are caught by pack(), values wider than declared by the kernel)
    rec, too_wide = calc.pack(inp, on_range="mask")
    packed = calc.calculateWitnessBatchPacked(rec, range_mask=too_wide)
    assert np.array_equal(packed.status & W.STATUS_INPUT_RANGE, res.status & W.STATUS_INPUT_RANGE)
    in_range = (res.status & W.STATUS_INPUT_RANGE) == 0
    assert np.array_equal(packed.status[in_range], res.status[in_range])
    assert np.array_equal(packed.public[in_range], res.public[in_range])
    calc.close()
    return res


def test_mix_ragged_batch_and_tiles():
    prog = oracle_ref.RefProgram(W.artifact("t_mix"))
    inp = random_inputs(prog.meta, 301, 21)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[5, d["x"]["offset"]] = inp[5, d["y"]["offset"]]
    run_and_compare("t_mix", inp)                 # one tile, 301 lanes (not a multiple of 128)
    run_and_compare("t_mix", inp, tile=128)       # three tiles, last one ragged
    run_and_compare("t_mix", inp[:1])             # batch of one


def test_mix_out_of_range_input_is_flagged():
    prog = oracle_ref.RefProgram(W.artifact("t_mix"))
    inp = random_inputs(prog.meta, 40, 4)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[7, d["bits"]["offset"] + 2, 0] = 2       # a "bit" input that is not a bit
    inp[9, d["u"]["offset"], 1] = 1              # a 16-bit input with a high limb set
    res = run_and_compare("t_mix", inp)
    assert res.status[7] & W.STATUS_INPUT_RANGE and res.status[9] & W.STATUS_INPUT_RANGE
    assert (np.delete(res.status, [7, 9]) == 0).all()


def test_packed_path_rejects_values_that_do_not_fit_the_record():
    """a 'bit' of 256 or a limb above 2^64 must not be narrowed silently on the packed path."""
    calc = W.WitnessCalculator(W.artifact("t_mix"), 0)
    inp = random_inputs(calc.meta, 16, 5)
    d = {x["name"]: x for x in calc.meta["inputs"]}
    inp[3, d["bits"]["offset"], 0] = 256         # wraps to 0 as a byte
    inp[4, d["u"]["offset"], 2] = 7              # limb 2 of a 16-bit input
    with pytest.raises(W.PzkError, match="declared range"):
        calc.pack(inp)
    rec, bad = calc.pack(inp, on_range="mask")
    assert list(np.nonzero(bad)[0]) == [3, 4]
    res = calc.calculateWitnessBatchPacked(rec, range_mask=bad)
    ref = calc.calculateWitnessBatch(inp)
    assert np.array_equal(res.status & W.STATUS_INPUT_RANGE, ref.status & W.STATUS_INPUT_RANGE)
    assert res.status[3] & W.STATUS_INPUT_RANGE and res.status[4] & W.STATUS_INPUT_RANGE
    calc.close()


def test_bigdiv_intrinsic():
    prog = oracle_ref.RefProgram(W.artifact("t_bigdiv"))
    inp = random_inputs(prog.meta, 257, 8)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[:, d["b"]["offset"] + 1, 0] |= np.uint64(1)
    inp[1, d["b"]["offset"] + 1, 0] = np.uint64(1)
    inp[2, :, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    inp[3, d["b"]["offset"] + 1, 0] = np.uint64(0)   # precondition violated -> status 8
    res = run_and_compare("t_bigdiv", inp)
    assert res.status[3] & W.STATUS_BIGDIV


def test_early_returns_and_narrowed_limb():
    prog = oracle_ref.RefProgram(W.artifact("t_earlyret"))
    inp = random_inputs(prog.meta, 200, 3)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[:, d["b"]["offset"] + 1, 0] |= np.uint64(1)
    inp[0, d["x"]["offset"]] = 0
    for i in range(1, 6):
        inp[i, d["v"]["offset"]:d["v"]["offset"] + (i - 1), 0] = 0
    inp[9, d["a"]["offset"] + 2, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)   # a[2] + c = 2^64 does not fit a limb
    inp[9, d["c"]["offset"], 0] = 1
    res = run_and_compare("t_earlyret", inp)
    assert res.status[9] & W.STATUS_ASSERT
    assert (np.delete(res.status, 9) == 0).all()


def test_program_of_another_format_version_is_refused(tmp_path):
    """The record semantics change with the format version (pzk_program.h PZK_VERSION): a stale program must be
    refused when it is opened, not mis-executed."""
    import struct
    blob = bytearray(open(W.artifact("t_mix"), "rb").read())
    struct.pack_into("<I", blob, 4, struct.unpack_from("<I", blob, 4)[0] - 1)
    stale = tmp_path / "stale.pzkp"
    stale.write_bytes(bytes(blob))
    with pytest.raises(W.PzkError, match="magic/version"):
        W.WitnessCalculator(str(stale), device=0)


def test_fused_multiply_add_records():
    """PZK_F_MULADD / PZK_Z_MULADD in every sign combination against the C oracle evaluator, wire by wire."""
    prog = oracle_ref.RefProgram(W.artifact("t_muladd"))
    inp = random_inputs(prog.meta, 300, 21)
    d = {x["name"]: x for x in prog.meta["inputs"]}
    inp[0, d["c"]["offset"], 0] = 0
    inp[1, d["a"]["offset"], 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    inp[1, d["b"]["offset"], 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    inp[2, d["a"]["offset"], 0] = 0
    inp[3, d["x"]["offset"]] = 0
    res = run_and_compare("t_muladd", inp)
    assert (res.status == 0).all()


def test_modinv_intrinsic():
    prog = oracle_ref.RefProgram(W.artifact("t_modinv"))
    inp = random_inputs(prog.meta, 300, 9)
    inp[0, :, 0] = 0
    inp[1, 0, 0] = np.uint64(0xFFFFFFFFFFFFFFFF); inp[1, 1, 0] = np.uint64(33554431)
    inp[2, 0, 0] = 1; inp[2, 1, 0] = 0
    inp[3, :, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    res = run_and_compare("t_modinv", inp)
    assert (res.status == 0).all()
    rem = res.public[:, 2, 0]                        # outputs: inv[2], rem
    assert rem[0] == 0 and rem[1] == 0 and (rem[2:] == 1).all()


@pytest.mark.parametrize("name,B", [("poseidon2", 200), ("sha256_1", 130), ("babyjub", 66), ("p256dbl", 40)])
def test_reference_small(name, B):
    prog = oracle_ref.RefProgram(W.artifact(name))
    inp = random_inputs(prog.meta, B, 33, field_bits=248)
    run_and_compare(name, inp, export=range(0, B, 7))


def smt_inputs(meta, keys, break_lane=None):
    from passport_zk_circuits_b200.poseidon import poseidon
    rows = []
    for i, key in enumerate(keys):
        vals = {"root": [poseidon([key, key, 1])], "leaf": [key], "key": [key], "siblings": [0] * 80}
        if i == break_lane:
            vals["root"] = [vals["root"][0] ^ 1]
        flat = []
        for d in meta["inputs"]:
            flat += vals[d["name"]]
        rows.append(ints_to_u64(flat))
    return np.stack(rows)


def test_smt80_config1():
    prog = oracle_ref.RefProgram(W.artifact("smt80"))
    keys = [12345 + 977 * i + (i << 190) for i in range(150)]
    inp = smt_inputs(prog.meta, keys, break_lane=11)
    res = run_and_compare("smt80", inp, export=[0, 11, 149])
    verified = res.public[:, 0, 0]
    assert verified[11] == 0 and (np.delete(verified, 11) == 1).all()
    assert (res.status == 0).all()


def test_operator_surface_and_wtns_check(tmp_path):
    """calculateWitness / calculateWTNSBin / checkConstraints / `wtns check`, the reference's
    operator surface (/root/reference/test/automatisationTest.js:37-51, gen-witness.sh:25)."""
    import formats
    import os
    from util import ROOT
    prefix = os.path.join(ROOT, "artifacts", "t_mix")
    circuit = W.wasm_tester(prefix)
    inp = {"x": 1234567, "y": 99, "u": [5, 7, 11, 13], "bits": [1, 0, 1, 1, 0, 0, 1, 0]}
    w = circuit.calculateWitness(inp, True)
    assert w[0] == 1 and w[circuit.symbols()["main.prod"]] == 1234567 * 99 + 7
    assert circuit.checkConstraints(w)
    blob = circuit.calc.calculateWTNSBin(inp)
    assert blob == formats.write_wtns(w)                       # byte-identical .wtns
    ok, first = W.wtns_check(prefix + ".r1cs", blob)
    assert ok and first == -1
    r1 = formats.read_r1cs(prefix + ".r1cs")
    for wire in (circuit.symbols()["main.x2"], circuit.symbols()["main.q"], 3):
        bad = list(w)
        bad[wire] = (bad[wire] + 1) % W.P
        ok, first = W.wtns_check(prefix + ".r1cs", formats.write_wtns(bad))
        assert (ok, first) == formats.wtns_check(r1, bad)       # same verdict, same first failing index
        assert not ok
        with pytest.raises(W.PzkError, match="Constraint doesn't match"):
            circuit.checkConstraints(bad)
    # an input that violates a constraint -> "Assert Failed." like the wasm
    with pytest.raises(W.PzkError, match="declared range"):
        circuit.calculateWitness(dict(inp, u=[5, 7, 70000, 13]))   # u is declared 16 bits wide
    smt = W.WitnessCalculator(W.artifact("smt80"), 0)
    with pytest.raises(W.PzkError, match="Assert Failed"):
        # last sibling must be zero: (isZero[N-1].out - 1) === 0 fails -> the wasm throws "Assert Failed."
        smt.calculateWitness({"root": 1, "leaf": 2, "key": 2, "siblings": [0] * 79 + [5]})
    with pytest.raises(W.PzkError, match="Curve of the witness"):
        W.wtns_check(prefix + ".r1cs", blob[:28] + bytes(32) + blob[60:])


@pytest.mark.parametrize("name", ["poseidon2", "sha256_1", "smt80"])
def test_golden_small(name):
    import hashlib
    import json
    import os
    from util import ROOT
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    calc = W.WitnessCalculator(W.artifact(name), 0)
    size = {d["name"]: d["size"] for d in calc.meta["inputs"]}
    ins = [{k: (list(v) if isinstance(v, str) and size[k] > 1 else v) for k, v in c["inputs"].items()} for c in g["cases"]]
    res = calc.calculateWitnessBatch(ins, export_lanes=range(len(ins)))
    for j, case in enumerate(g["cases"]):
        assert res.status[j] == 0
        assert hashlib.sha256(res.witnesses[j].tobytes()).hexdigest() == case["wtns_data_sha256"]


def test_r1cs_stream_kernel_smt80():
    """Stand-alone `wtns check` kernel (rows-parallel, TMA-staged A/B/C stream) on 70 explicit
    witnesses of the SMT(80) circuit; corrupted wires must give the first failing constraint index
    the Python restatement of snarkjs' loop gives."""
    import formats
    import os
    from util import ROOT
    prefix = os.path.join(ROOT, "artifacts", "smt80")
    prog = oracle_ref.RefProgram(prefix + ".pzkp")
    keys = [777 + 31 * i + (i << 120) for i in range(70)]
    inp = smt_inputs(prog.meta, keys)
    calc = W.WitnessCalculator(prefix + ".pzkp", 0)
    res = calc.calculateWitnessBatch(inp, export_lanes=range(70))
    wit = res.witnesses.copy()
    r1 = formats.read_r1cs(prefix + ".r1cs")
    rng = np.random.default_rng(5)
    bad_lanes = {3: 100, 40: 90000, 69: int(rng.integers(2, prog.n_wires))}
    for lane, wire in bad_lanes.items():
        wit[lane, wire, 0] ^= np.uint64(1)
    ok, first, ms = W.r1cs_check_batch(prefix + ".r1cs", wit)
    for lane in range(70):
        if lane in bad_lanes:
            w = [int.from_bytes(wit[lane, i].tobytes(), "little") for i in range(prog.n_wires)]
            assert (bool(ok[lane]), int(first[lane])) == formats.wtns_check(r1, w), lane
            assert not ok[lane]
        else:
            assert ok[lane] and first[lane] == -1
    assert ms > 0


@pytest.mark.parametrize("name", ["mont_test", "inv_test", "tma_test"])
def test_device_unit(name):
    """tests/cuda/*.cu: PTX Montgomery product == plain formulation, inverse * x == 1, TMA + mbarrier staging."""
    import os
    import subprocess
    from util import ROOT
    exe = os.path.join(ROOT, "tests", "bin", name)
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


def test_imad_peak_microbenchmark():
    """tests/cuda/imad_peak.cu: the measured denominators bench.py's roofline uses (dependent-free IMAD rate,
    Montgomery products per second with register operands)."""
    import json
    import os
    import subprocess
    from util import ROOT
    out = subprocess.run([os.path.join(ROOT, "tests", "bin", "imad_peak")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    # 148 SMs x 64 IMAD / clock at ~1.9 GHz is 1.8e13; anything far below means the benchmark did not fill the GPU
    assert 5e12 < d["imad_per_s"] < 4e13 and d["fr_mul_per_s"] * 136 <= d["imad_per_s"] * 1.05, d
