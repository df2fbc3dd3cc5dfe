"""Small driver for ncu captures: one C3 wave through the evaluator with the witness digest on
(`python tests/ncu_probe.py [lanes] [r1cs_lanes]`); with r1cs_lanes > 0 the device-resident hand-off to the
R1CS stream kernel follows."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from passport_zk_circuits_b200 import witness as W
from passport_zk_circuits_b200.passports import C3, PassportFactory
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
R = int(sys.argv[2]) if len(sys.argv) > 2 else 0
calc = W.WitnessCalculator(W.artifact("c3"), 0)
fac = PassportFactory(C3, seed=1, n_sig_keys=2, n_aa_keys=2)
uniq = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(32)])
inp = np.tile(uniq, ((B + 31) // 32, 1, 1))[:B]
calc.set_digest(True)
calc.upload(inp)
calc.run(True)
res = calc.download()
assert (res.status == 0).all()
if R:
    calc.set_digest(False)
    r = W.R1cs(W.artifact_r1cs("c3"), 0)
    ok, fb, t_eval, t_check = r.check_circuit(calc, np.arange(R) * (B // R))
    assert ok.all()
    print("r1cs", R, "lanes", round(t_check, 2), "ms")
print("ok", B)
