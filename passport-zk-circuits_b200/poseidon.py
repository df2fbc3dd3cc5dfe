"""Host-side Poseidon over BN254 Fr (x^5, RF=8, circomlib round numbers).

Used by the synthetic-passport generator for the fake slave-tree root,
root = Poseidon3(pkHash, pkHash, 1), the rule of
/root/reference/test/process_passport.js:628-657 (getFakeIdenData), where the
reference calls /root/reference/test/poseidon.js:134.

The reference ships the constants as a 25k-line table
(test/poseidon_constants.js).  They are not copied here: the round constants
and the MDS matrix are re-derived from the Poseidon paper's Grain-LFSR
parameter generator (field=1, sbox=0, n=254, t, RF, RP), which is how circomlib
produced them.  tests/test_cpu_host.py::test_poseidon_constants_match_reference_tables checks the derived tables against
the reference's tables when /root/reference is present.  The un-optimised
permutation below yields the same digests as the reference's optimised one.
"""
from functools import lru_cache

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_ROUNDS_F = 8
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]


def _grain_bits(t, rf, rp, n=254):
    state = []
    for v, w in ((1, 2), (0, 4), (n, 12), (t, 12), (rf, 10), (rp, 10)):
        state.extend(int(c) for c in format(v, "0%db" % w))
    state.extend([1] * 30)

    def step():
        nb = state[62] ^ state[51] ^ state[38] ^ state[23] ^ state[13] ^ state[0]
        state.pop(0)
        state.append(nb)
        return nb

    for _ in range(160):
        step()
    while True:
        b = step()
        while b == 0:
            step()
            b = step()
        yield step()


@lru_cache(maxsize=None)
def constants(t):
    """(C, M): (RF+RP)*t round constants and the t x t MDS matrix (row-major,
    out[i] = sum_j M[i][j] * in[j])."""
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    g = _grain_bits(t, rf, rp)

    def take(n=254):
        v = 0
        for _ in range(n):
            v = (v << 1) | next(g)
        return v

    C = []
    while len(C) < (rf + rp) * t:
        v = take()
        if v < P:
            C.append(v)
    while True:
        xy = [take() % P for _ in range(2 * t)]
        if len(set(xy)) == 2 * t:
            break
    xs, ys = xy[:t], xy[t:]
    M = [[pow((xs[i] + ys[j]) % P, -1, P) for j in range(t)] for i in range(t)]
    return C, M


def poseidon(inputs):
    """poseidon(inputs) of /root/reference/test/poseidon.js:134 (initial state 0,
    one output)."""
    t = len(inputs) + 1
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    C, M = constants(t)
    s = [0] + [int(x) % P for x in inputs]
    for r in range(rf + rp):
        s = [(x + C[r * t + i]) % P for i, x in enumerate(s)]
        if r < rf // 2 or r >= rf // 2 + rp:
            s = [pow(x, 5, P) for x in s]
        else:
            s[0] = pow(s[0], 5, P)
        s = [sum(M[i][j] * s[j] for j in range(t)) % P for i in range(t)]
    return s[0]
