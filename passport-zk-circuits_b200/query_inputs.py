"""Synthetic inputs for the query circuit (BASELINE config 2), following the recipe the reference
uses for its generated tests (/root/reference/helpers/generateRegisterIdentityTest.js:186-230) and
README.md:107-136: identity tree with zero siblings, root = Poseidon3(index, value, 1),
index = Poseidon2(pkPassportHash, Poseidon2(sk * Base8)), value = Poseidon3(dgCommit, counter, timestamp),
dgCommit = Poseidon5(4 x 186-bit little-endian DG1 chunks, Poseidon1(sk))
(/root/reference/circuits/identityManagement/queryIdentity.circom:37-229)."""
from __future__ import annotations

import random

from .poseidon import P, poseidon

# BabyJubjub (twisted Edwards a x^2 + y^2 = 1 + d x^2 y^2), /root/reference/circuits/lib/circuits/babyjubjub
A, D = 168700, 168696
BASE8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
         16950150798460657717958625567821834550301663161624707787222815936182638968203)


def ed_add(p, q):
    (x1, y1), (x2, y2) = p, q
    t = D * x1 * x2 * y1 * y2 % P
    x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, P) % P
    y3 = (y1 * y2 - A * x1 * x2) * pow(1 - t, -1, P) % P
    return x3, y3


def ed_mul(k, p):
    acc = (0, 1)
    while k:
        if k & 1:
            acc = ed_add(acc, p)
        p = ed_add(p, p)
        k >>= 1
    return acc


_NAT = ["UKR", "USA", "DEU", "FRA", "GEO", "POL", "ESP", "ITA"]
_AL = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def td3_dg1(rng) -> bytes:
    """93-byte DG1 of a TD3 passport: tag/len 61 5B 5F 1F 58 + two 44-character MRZ lines."""
    nat = rng.choice(_NAT)
    name = ("".join(rng.choice(_AL) for _ in range(rng.randint(3, 9))) + "<<" +
            "".join(rng.choice(_AL) for _ in range(rng.randint(3, 9))))
    line1 = ("P<" + nat + name).ljust(44, "<")[:44]
    docnum = "".join(rng.choice(_AL + "0123456789") for _ in range(9))
    dob = "%02d%02d%02d" % (rng.randint(50, 99), rng.randint(1, 12), rng.randint(1, 28))
    exp = "%02d%02d%02d" % (rng.randint(27, 35), rng.randint(1, 12), rng.randint(1, 28))
    line2 = (docnum + "0" + nat + dob + "0" + rng.choice("MF") + exp + "0").ljust(43, "<")[:43] + "0"
    return bytes([0x61, 0x5B, 0x5F, 0x1F, 0x58]) + (line1 + line2).encode()


def td1_dg1(rng) -> bytes:
    """95-byte DG1 of a TD1 identity card: tag/len 61 5D 5F 1F 5A + three 30-character MRZ lines
    (ICAO 9303 part 5; /root/reference/circuits/identityManagement/queryIdentityTD1.circom:75 reads 760 bits)."""
    nat = rng.choice(_NAT)
    docnum = "".join(rng.choice(_AL + "0123456789") for _ in range(9))
    line1 = ("ID" + nat + docnum + "0").ljust(30, "<")[:30]
    dob = "%02d%02d%02d" % (rng.randint(50, 99), rng.randint(1, 12), rng.randint(1, 28))
    exp = "%02d%02d%02d" % (rng.randint(27, 35), rng.randint(1, 12), rng.randint(1, 28))
    line2 = (dob + "0" + rng.choice("MF") + exp + "0" + nat).ljust(29, "<")[:29] + "0"
    name = ("".join(rng.choice(_AL) for _ in range(rng.randint(3, 9))) + "<<" +
            "".join(rng.choice(_AL) for _ in range(rng.randint(3, 9))))
    line3 = name.ljust(30, "<")[:30]
    return bytes([0x61, 0x5D, 0x5F, 0x1F, 0x5A]) + (line1 + line2 + line3).encode()


def make_query_input(index: int, seed: int = 1, selector: int = 39, td1: bool = False) -> dict:
    rng = random.Random((seed << 20) ^ index)
    dg1 = td1_dg1(rng) if td1 else td3_dg1(rng)
    bits = [(b >> (7 - i)) & 1 for b in dg1 for i in range(8)]
    sk = rng.getrandbits(248)
    pk_passport_hash = rng.getrandbits(250)
    timestamp, counter = 1713436475 + index, 1
    cs = 190 if td1 else 186      # DG1 commitment chunk size (queryIdentityTD1.circom:203-212)
    chunks = [sum(bits[i * cs + j] << j for j in range(cs)) for i in range(4)]
    dg_commit = poseidon(chunks + [poseidon([sk])])
    value = poseidon([dg_commit, counter, timestamp])
    pk_hash = poseidon(list(ed_mul(sk, BASE8)))
    idx = poseidon([pk_passport_hash, pk_hash])
    root = poseidon([idx, value, 1])
    return {
        "dg1": [str(b) for b in bits], "eventID": "0x1234567890", "eventData": "0x12345678901234567890",
        "idStateRoot": str(root), "idStateSiblings": ["0"] * 80, "pkPassportHash": str(pk_passport_hash),
        "selector": str(selector), "skIdentity": str(sk), "timestamp": str(timestamp),
        "currentDate": "0x323430383230", "identityCounter": str(counter), "timestampLowerbound": "0",
        "timestampUpperbound": "19000000000", "identityCounterLowerbound": "0", "identityCounterUpperbound": "1000",
        "birthDateLowerbound": "0x303030303030", "birthDateUpperbound": "0x303030303030",
        "expirationDateLowerbound": "0x303030303030", "expirationDateUpperbound": "0x303030303030",
        "citizenshipMask": "0",
    }
