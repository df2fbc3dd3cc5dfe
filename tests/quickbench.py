import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from passport_zk_circuits_b200 import witness as W
from passport_zk_circuits_b200.passports import C3, PassportFactory
name = sys.argv[1]; B = int(sys.argv[2])
WAVES = int(sys.argv[3]) if len(sys.argv) > 3 else 0
DIGEST = len(sys.argv) > 4 and sys.argv[4] == 'digest'
calc = W.WitnessCalculator(W.artifact(name), 0)
if DIGEST: calc.set_digest(True)
name = "c3" if name.startswith("c3") else name
if WAVES: B = WAVES * calc.wave_lanes()
print(name, 'wave', calc.wave_lanes(), 'B', B, 'wires', calc.n_wires, 'constraints', calc.n_constraints, calc.stats())
if name == 'c3':
    fac = PassportFactory(C3, seed=1, n_sig_keys=2, n_aa_keys=2)
    uniq = W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(64)])
    inp = np.tile(uniq, ((B + 63) // 64, 1, 1))[:B]
elif name.startswith('c4_'):
    from passport_zk_circuits_b200.artifacts import C4_VARIANTS
    fac = PassportFactory(C4_VARIANTS[name], seed=1, n_sig_keys=2, n_aa_keys=2)
    uniq = calc.pack(W.pack_inputs_fast(calc.meta, [fac.make(i).inputs for i in range(16)]))
    inp = None
    packed = np.tile(uniq, ((B + 15) // 16, 1))[:B].copy()
else:
    sys.path.insert(0, '/root/repo/tests')
    from util import random_inputs
    inp = random_inputs(calc.meta, B, 1, field_bits=248)
if inp is None:
    calc.upload_packed(packed)
else:
    calc.upload(inp)
calc.profile(enable=True, reset=True)
for it in range(3):
    t = time.time(); calc.run(True); dt = time.time() - t
    pr = calc.profile(); calc.profile(reset=True)
    print('run', it, 'lanes', B, 'tile', calc.tile_lanes(), 'sec', round(dt, 4), 'witness/s', round(B / dt, 1), {k: (round(v[0], 1), v[1]) for k, v in pr.items()})
import ctypes
nseg = calc.meta['stats']['segments']
arr = (ctypes.c_double * nseg)()
calc.profile(enable=True, reset=True); calc.run(True)
calc._L.pzk_profile_segments.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_uint32]
calc._L.pzk_profile_segments(calc._h, arr, nseg)
seg = list(arr); tot = sum(seg)
print('segments', nseg, 'total ms', round(tot, 1))
import statistics
print('median seg ms', round(statistics.median(seg), 2), 'top:', sorted([(round(v, 1), i) for i, v in enumerate(seg)], reverse=True)[:25])
buckets = {}
for i, v in enumerate(seg): buckets[i * 20 // nseg] = buckets.get(i * 20 // nseg, 0) + v
print('by 5% of program:', [round(buckets.get(k, 0)) for k in range(20)])
res = calc.download()
print('status ok', int((res.status == 0).sum()), 'of', B)
