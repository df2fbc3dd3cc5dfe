"""Timing of the R1CS stream kernel fed by the evaluator on the device (`python tests/r1cs_bench.py [lanes] [batch]`):
evaluate a batch of registerIdentity passports, hand `lanes` of them to the checker, print the CUDA-event time of the
stream kernel (second call: the first one pays first-touch effects)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from passport_zk_circuits_b200 import witness as W
from passport_zk_circuits_b200.passports import C3, PassportFactory
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
calc = W.WitnessCalculator(W.artifact("c3"), 0)
fac = PassportFactory(C3, seed=1, n_sig_keys=2, n_aa_keys=2)
uniq = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(32)])
calc.upload(np.tile(uniq, ((B + 31) // 32, 1, 1))[:B])
r = W.R1cs(W.artifact_r1cs("c3"), 0)
lanes = np.arange(R) * (B // R)
for it in range(3):
    ok, fb, t_eval, t_check = r.check_circuit(calc, lanes)
    assert ok.all()
    gbs = r.n_terms * 32 * R / (t_check / 1e3) / 1e9
    print(f"run {it}: {R} lanes, eval+export {t_eval:.1f} ms, stream check {t_check:.2f} ms = {R / (t_check / 1e3):.0f} witness checks/s, "
          f"{gbs:.0f} GB/s algorithmic")
