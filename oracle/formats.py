"""ORACLE (test infrastructure): pure-Python readers for the iden3 binary formats and the
`snarkjs wtns check` loop with Python integers (SURVEY.md section 8b restates the formats;
the reference's call site is /root/reference/test/automatisationTest.js:51 and the CLI of
/root/reference/circuits/scripts/prove.sh:25-27).  Never imported by the product."""
import struct

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def read_r1cs(path):
    with open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"r1cs"
    version, nsec = struct.unpack_from("<II", data, 4)
    assert version == 1
    pos = 12
    sections = {}
    for _ in range(nsec):
        typ, ln = struct.unpack_from("<IQ", data, pos)
        pos += 12
        sections[typ] = (pos, ln)
        pos += ln
    hp, _ = sections[1]
    n8 = struct.unpack_from("<I", data, hp)[0]
    prime = int.from_bytes(data[hp + 4:hp + 4 + n8], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in = struct.unpack_from("<IIII", data, hp + 4 + n8)
    n_labels, n_constraints = struct.unpack_from("<QI", data, hp + 4 + n8 + 16)
    cp, _ = sections[2]
    cons = []
    q = cp
    for _ in range(n_constraints):
        lcs = []
        for _ in range(3):
            n = struct.unpack_from("<I", data, q)[0]
            q += 4
            lc = []
            for _ in range(n):
                w = struct.unpack_from("<I", data, q)[0]
                c = int.from_bytes(data[q + 4:q + 4 + n8], "little")
                q += 4 + n8
                lc.append((w, c))
            lcs.append(lc)
        cons.append(tuple(lcs))
    return {"prime": prime, "n_wires": n_wires, "n_pub_out": n_pub_out, "n_pub_in": n_pub_in,
            "n_prv_in": n_prv_in, "n_labels": n_labels, "constraints": cons}


def read_wtns(data: bytes):
    assert data[:4] == b"wtns"
    version, nsec = struct.unpack_from("<II", data, 4)
    assert version == 2
    pos = 12
    sections = {}
    for _ in range(nsec):
        typ, ln = struct.unpack_from("<IQ", data, pos)
        pos += 12
        sections[typ] = (pos, ln)
        pos += ln
    hp, _ = sections[1]
    n8 = struct.unpack_from("<I", data, hp)[0]
    prime = int.from_bytes(data[hp + 4:hp + 4 + n8], "little")
    n = struct.unpack_from("<I", data, hp + 4 + n8)[0]
    dp, dl = sections[2]
    assert dl == n * n8
    return prime, [int.from_bytes(data[dp + i * n8:dp + (i + 1) * n8], "little") for i in range(n)]


def write_wtns(witness) -> bytes:
    """calculateWTNSBin layout."""
    n = len(witness)
    out = bytearray(b"wtns")
    out += struct.pack("<II", 2, 2)
    out += struct.pack("<IQ", 1, 40) + struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<I", n)
    out += struct.pack("<IQ", 2, 32 * n)
    for v in witness:
        out += int(v).to_bytes(32, "little")
    return bytes(out)


def read_sym(path):
    out = {}
    with open(path) as f:
        for line in f:
            lab, wi, ci, name = line.rstrip("\n").split(",", 3)
            out[name] = int(wi)
    return out


def wtns_check(r1cs, witness, prime_w=P):
    """Returns (verdict, first_bad): first constraint i with A.w * B.w - C.w != 0 mod r."""
    if r1cs["prime"] != prime_w:
        raise ValueError("Curve of the witness does not match the curve of the r1cs")
    p = r1cs["prime"]
    for i, (A, B, C) in enumerate(r1cs["constraints"]):
        a = sum(c * witness[w] for w, c in A) % p
        b = sum(c * witness[w] for w, c in B) % p
        cc = sum(c * witness[w] for w, c in C) % p
        if (a * b - cc) % p != 0:
            return False, i
    return True, -1
