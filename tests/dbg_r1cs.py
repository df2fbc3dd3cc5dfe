import sys, os, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from passport_zk_circuits_b200 import witness as W
import ref as R, formats
prefix = '/root/repo/artifacts/t_mix'
prog = R.RefProgram(prefix + '.pzkp')
inp = W.pack_inputs_fast(prog.meta, [{"x": 1234567, "y": 99, "u": [5, 7, 11, 13], "bits": [1, 0, 1, 1, 0, 0, 1, 0]}])
st, fb, wit = prog.witness(inp[0])
print('ref ok', st, flush=True)
w = wit[None, :, :].copy()
print('calling r1cs_check_batch', flush=True)
ok, first, ms = W.r1cs_check_batch(prefix + '.r1cs', w)
print('result', ok, first, ms, flush=True)
