"""GPU parity for the north-star circuit registerIdentity (SHA-256 + RSA-2048),
RegisterIdentityBuilder(1,256,3,4,600,248,1,1496,3,256) of /root/reference/hardhat.config.ts:29."""
import numpy as np
import pytest

import ref as oracle_ref

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
    ok = np.ones(B, dtype=bool)
    ok[[5, 9]] = False
    assert (res.status[ok] == 0).all()
    assert (res.status[~ok] & W.STATUS_CONSTRAINT).all()
