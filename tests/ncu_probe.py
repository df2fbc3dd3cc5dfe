"""Small driver for ncu captures: one C3 tile, a few segments' worth of launches."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from passport_zk_circuits_b200 import witness as W
from passport_zk_circuits_b200.passports import C3, PassportFactory
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
calc = W.WitnessCalculator(W.artifact("c3"), 0)
fac = PassportFactory(C3, seed=1, n_sig_keys=2, n_aa_keys=2)
uniq = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(32)])
inp = np.tile(uniq, ((B + 31) // 32, 1, 1))[:B]
calc.upload(inp)
calc.run(True)
res = calc.download()
assert (res.status == 0).all()
print("ok", B)
