"""Stand-alone R1CS stream kernel on registerIdentity witnesses: N witnesses exported by the evaluator,
checked against the circuit's .r1cs (2 250 656 constraints, 6.2 M terms)."""
import lzma, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from passport_zk_circuits_b200 import witness as W
from passport_zk_circuits_b200.passports import C3, PassportFactory
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
r1 = os.path.join(W.ARTIFACT_DIR, "c3.r1cs")
if not os.path.exists(r1):
    with lzma.open(r1 + ".xz", "rb") as g, open(r1, "wb") as f:
        while True:
            b = g.read(1 << 24)
            if not b: break
            f.write(b)
calc = W.WitnessCalculator(W.artifact("c3"), 0)
fac = PassportFactory(C3, seed=1, n_sig_keys=2, n_aa_keys=2)
uniq = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(min(N, 16))])
inp = np.tile(uniq, ((N + len(uniq) - 1) // len(uniq), 1, 1))[:N]
t = time.time(); res = calc.calculateWitnessBatch(inp, export_lanes=range(N)); print("export s", round(time.time() - t, 2))
wit = res.witnesses
wit[1, 1000, 0] ^= np.uint64(1)
t = time.time(); ok, first, ms = W.r1cs_check_batch(r1, wit); wall = time.time() - t
terms = 6207122
gb = (terms * 32 * N) / 1e9
print(f"N={N} kernel_ms={ms:.1f} wall={wall:.1f}s verdicts ok={int(ok.sum())} bad_first={int(first[1])} "
      f"witness-checks/s={N / (ms / 1e3):.0f} constraints/s={N * 2250656 / (ms / 1e3):.3e} algorithmic GB/s={gb / (ms / 1e3):.0f}")
