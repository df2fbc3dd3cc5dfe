"""Developer tool: compile a circuit twice (with and without views / packed truth tables) and compare the two
programs wire by wire on random in-range inputs with the C oracle evaluator (no GPU needed).
usage: python tests/tools/diff_views.py <main.circom> [name:bits ...] [--lanes N]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref as oracle_ref  # noqa: E402
from util import random_inputs  # noqa: E402


def main():
    args = sys.argv[1:]
    lanes = 4
    if "--lanes" in args:
        i = args.index("--lanes"); lanes = int(args[i + 1]); del args[i:i + 2]
    src, bits = args[0], args[1:]
    pzkc = os.path.join(ROOT, "build", "pzkc")
    out = "/tmp/w"
    os.makedirs(out, exist_ok=True)
    base = os.path.splitext(os.path.basename(src))[0]
    flags = []
    for b in bits:
        flags += ["--bits", b]
    for tag, extra in (("v", []), ("s", ["--no-views"])):
        r = subprocess.run([pzkc, src, f"{out}/{base}_{tag}"] + flags + extra, capture_output=True, text=True)
        print(tag, r.stdout.strip().replace("\n", " | "), r.stderr[-2000:])
        if r.returncode:
            sys.exit(1)
    pv, ps = oracle_ref.RefProgram(f"{out}/{base}_v.pzkp"), oracle_ref.RefProgram(f"{out}/{base}_s.pzkp")
    print({k: pv.meta["stats"][k] for k in ("op_records", "static_rows", "table_rows", "symbolic_rows", "view_rows", "range_rows",
                                            "vlut", "vlut_lanes", "view_signals", "tabview_signals", "extracts", "i64_rows", "int_rows", "field_rows")})
    print("scalar op_records", ps.meta["stats"]["op_records"])
    inp = random_inputs(pv.meta, lanes, 1234)
    for b in range(lanes):
        a = pv.witness(inp[b]); c = ps.witness(inp[b])
        if a[0] != c[0] or a[1] != c[1]:
            print("lane", b, "status/first_bad differ", a[:2], c[:2])
        if not np.array_equal(a[2], c[2]):
            bad = np.nonzero((a[2] != c[2]).any(axis=1))[0]
            print("lane", b, "wires differ:", len(bad), bad[:10], [(hex(int(a[2][w][0])), hex(int(c[2][w][0]))) for w in bad[:5]])
            sys.exit(1)
    print("OK", lanes, "lanes,", pv.n_wires, "wires identical")


if __name__ == "__main__":
    main()
