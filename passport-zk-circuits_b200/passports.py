"""Batched synthetic-passport generator.

Replaces the reference's per-file input pipeline
(/root/reference/test/process_passport.js:674-816 `processPassport`) for the
benchmark and the tests: instead of parsing one real SOD/DG JSON it builds,
from a seed, any number of passports whose DG1 / DG15 / encapsulated content /
signed attributes have the hashes at the bit offsets the circuit parameters
name, signs the signed attributes and emits exactly the input object
`writeToJson` (process_passport.js:659-672) would write:

    dg1, dg15, signedAttributes, encapsulatedContent : "0"/"1" strings, message
        bytes MSB-first, SHA-padded to whole blocks (padding(), :11-91)
    pubkey, signature : decimal strings, 64-bit little-endian chunks
        (bigintToArrayString, :125-135)
    skIdentity        : "0x" + first 62 hex digits of SHA-256(EC) (:630,667)
    slaveMerkleRoot   : "0x" + Poseidon3(pkHash, pkHash, 1) (:642-654)
    slaveMerkleInclusionBranches : 80 x "0"

No reference code is imported; the layout rules are restated from SURVEY.md
section 8(d).
"""
from __future__ import annotations

import hashlib
import math
import random
from dataclasses import dataclass

from .poseidon import poseidon

TREE_DEPTH = 80


@dataclass(frozen=True)
class CircuitParams:
    """The 10 template parameters of RegisterIdentityBuilder
    (/root/reference/circuits/identityManagement/registerIdentityBuilder.circom:41-52)."""
    sig_type: int = 1
    dg_hash: int = 256
    doc_type: int = 3
    ec_blocks: int = 4
    ec_shift: int = 600
    dg1_shift: int = 248
    aa_algo: int = 1
    dg15_shift: int = 1496
    dg15_blocks: int = 3
    aa_shift: int = 256

    def as_tuple(self):
        return (self.sig_type, self.dg_hash, self.doc_type, self.ec_blocks, self.ec_shift,
                self.dg1_shift, self.aa_algo, self.dg15_shift, self.dg15_blocks, self.aa_shift)

    @property
    def name(self):
        return "registerIdentity_" + "_".join(str(x) for x in self.as_tuple())

    def main_source(self, include_path):
        """The generated root file, as writeToCircom does (process_passport.js:573-588)."""
        args = ", ".join(str(x) for x in self.as_tuple())
        return ("pragma circom 2.1.6;\n\n"
                f'include "{include_path}";\n\n'
                f"component main {{ public [slaveMerkleRoot] }} = RegisterIdentityBuilder({args});\n")


# canonical north-star circuit (/root/reference/hardhat.config.ts:29)
C3 = CircuitParams()


# --------------------------------------------------------------------------- primes / keys
def _is_probable_prime(n, rng, rounds=24):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for _ in range(rounds):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _gen_prime(bits, rng, e=65537):
    while True:
        p = rng.getrandbits(bits) | (3 << (bits - 2)) | 1
        if math.gcd(p - 1, e) != 1:     # e need not be prime (37187 = 41 * 907, SIGNATURE_TYPE 4)
            continue
        if _is_probable_prime(p, rng):
            return p


class RsaKey:
    def __init__(self, bits, rng, e=65537):
        while True:
            p = _gen_prime(bits // 2, rng, e)
            q = _gen_prime(bits // 2, rng, e)
            if p != q and (p * q).bit_length() == bits:
                break
        self.n, self.e, self.p, self.q = p * q, e, p, q
        d = pow(e, -1, (p - 1) * (q - 1))
        self.dp, self.dq, self.qinv = d % (p - 1), d % (q - 1), pow(q, -1, p)
        self.bits = bits

    def private_op(self, m):
        m1 = pow(m % self.p, self.dp, self.p)
        m2 = pow(m % self.q, self.dq, self.q)
        h = (self.qinv * (m1 - m2)) % self.p
        return m2 + h * self.q


# sha1 PKCS#1 DigestInfo is checked by rsa.circom:101-110
_DIGEST_INFO = {
    "sha256": bytes.fromhex("3031300d060960864801650304020105000420"),
    "sha1": bytes.fromhex("3021300906052b0e03021a05000414"),
}


def pkcs1v15_sign(key: RsaKey, msg: bytes, hash_name="sha256"):
    k = key.bits // 8
    t = _DIGEST_INFO[hash_name] + hashlib.new(hash_name, msg).digest()
    em = b"\x00\x01" + b"\xff" * (k - len(t) - 3) + b"\x00" + t
    return key.private_op(int.from_bytes(em, "big"))


def mgf1(seed: bytes, length: int, hash_name: str) -> bytes:
    out = b""
    counter = 0
    while len(out) < length:
        out += hashlib.new(hash_name, seed + counter.to_bytes(4, "big")).digest()
        counter += 1
    return out[:length]


def pss_sign(key: RsaKey, msg: bytes, hash_name: str, salt_len: int, rng) -> int:
    """RSASSA-PSS (EMSA-PSS, MGF1 with the message hash), the encoding
    /root/reference/circuits/lib/circuits/signatures/rsaPss.circom:18-254 unpacks."""
    em_bits = key.bits - 1
    em_len = (em_bits + 7) // 8
    m_hash = hashlib.new(hash_name, msg).digest()
    h_len = len(m_hash)
    salt = bytes(rng.randrange(256) for _ in range(salt_len))
    h = hashlib.new(hash_name, b"\x00" * 8 + m_hash + salt).digest()
    db = b"\x00" * (em_len - salt_len - h_len - 2) + b"\x01" + salt
    mask = mgf1(h, em_len - h_len - 1, hash_name)
    masked = bytearray(a ^ b for a, b in zip(db, mask))
    masked[0] &= 0xFF >> (8 * em_len - em_bits)
    em = bytes(masked) + h + b"\xbc"
    return key.private_op(int.from_bytes(em, "big"))


# Curves of the ECDSA signature types (/root/reference/circuits/signatureVerifier/signatureVerification.circom:
# 177-191 SIG 20 = NIST P-256, :192-205 SIG 21 = brainpoolP256r1, :234-247 SIG 24 = secp224r1 in 7 chunks of 32 bits;
# the domain parameters are the published ones (FIPS 186-4 D.1.2, RFC 5639 3.4) and are checked against the
# `cryptography` package by tests/test_cpu_host.py)
@dataclass(frozen=True)
class Curve:
    name: str
    p: int
    a: int
    b: int
    n: int
    gx: int
    gy: int
    chunk_bits: int
    chunks: int


P256_P = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
P256_A = P256_P - 3
P256_B = 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B
P256_N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
P256_G = (0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296,
          0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5)
CURVE_P256 = Curve("secp256r1", P256_P, P256_A, P256_B, P256_N, P256_G[0], P256_G[1], 64, 4)
CURVE_BP256 = Curve(
    "brainpoolP256r1",
    0xA9FB57DBA1EEA9BC3E660A909D838D726E3BF623D52620282013481D1F6E5377,
    0x7D5A0975FC2C3057EEF67530417AFFE7FB8055C126DC5C6CE94A4B44F330B5D9,
    0x26DC5C6CE94A4B44F330B5D9BBD77CBF958416295CF7E1CE6BCCDC18FF8C07B6,
    0xA9FB57DBA1EEA9BC3E660A909D838D718C397AA3B561A6F7901E0E82974856A7,
    0x8BD2AEB9CB7E57CB2C4B482FFC81B7AFB9DE27E1E3BD23C23A4453BD9ACE3262,
    0x547EF835C3DAC4FD97F8461A14611DC9C27745132DED8E545C1D54C72F046997, 64, 4)
CURVE_P224 = Curve(
    "secp224r1",
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF000000000000000000000001,
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFF000000000000000000000001 - 3,
    0xB4050A850C04B3ABF54132565044B0B7D7BFD8BA270B39432355FFB4,
    0xFFFFFFFFFFFFFFFFFFFFFFFFFFFF16A2E0B8F03E13DD29455C5C2A3D,
    0xB70E0CBD6BB4BF7F321390B94A03C1D356C21122343280D6115C1D21,
    0xBD376388B5F723FB4C22DFE6CD4375A05A07476444D5819985007E34, 32, 7)
SIG_CURVES = {20: CURVE_P256, 21: CURVE_BP256, 24: CURVE_P224}


def _ec_add(p1, p2, a=P256_A, p=P256_P):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = (3 * x1 * x1 + a) * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return x3, (lam * (x1 - x3) - y1) % p


def _ec_mul(k, pt, curve=CURVE_P256):
    acc = None
    while k:
        if k & 1:
            acc = _ec_add(acc, pt, curve.a, curve.p)
        pt = _ec_add(pt, pt, curve.a, curve.p)
        k >>= 1
    return acc


def _ecdsa_z(msg: bytes, hash_name: str, curve: Curve) -> int:
    """leftmost min(hash bits, bit length of n) bits of the digest (FIPS 186-4 6.4)"""
    d = hashlib.new(hash_name, msg).digest()
    z = int.from_bytes(d, "big")
    extra = len(d) * 8 - curve.n.bit_length()
    return z >> extra if extra > 0 else z


class EcKey:
    """ECDSA signer key; `n` is kept so the RSA-shaped call sites (`key.n`) read the x coordinate."""

    def __init__(self, rng, curve: Curve = CURVE_P256):
        self.curve = curve
        self.d = rng.randrange(1, curve.n)
        self.x, self.y = _ec_mul(self.d, (curve.gx, curve.gy), curve)
        self.n = self.x

    def sign(self, msg: bytes, hash_name: str, rng):
        cv = self.curve
        z = _ecdsa_z(msg, hash_name, cv)
        while True:
            k = rng.randrange(1, cv.n)
            r = _ec_mul(k, (cv.gx, cv.gy), cv)[0] % cv.n
            s = pow(k, -1, cv.n) * (z + r * self.d) % cv.n
            if r and s:
                return r, s


def ecdsa_verify(x, y, msg: bytes, hash_name: str, r, s, curve: Curve = CURVE_P256):
    z = _ecdsa_z(msg, hash_name, curve)
    w = pow(s, -1, curve.n)
    pt = _ec_add(_ec_mul(z * w % curve.n, (curve.gx, curve.gy), curve), _ec_mul(r * w % curve.n, (x, y), curve),
                 curve.a, curve.p)
    return pt is not None and pt[0] % curve.n == r


def ec_pubkey_hash(x: int, y: int):
    """Poseidon2 of the low 248 bits of each coordinate
    (/root/reference/circuits/passportVerification/passportVerificationBuilder.circom:193-231)."""
    m = (1 << 248) - 1
    return poseidon([x & m, y & m])


# SIGNATURE_TYPE -> (modulus bits, scheme, signature hash bits, public exponent, PSS salt length)
# (/root/reference/circuits/signatureVerifier/signatureVerification.circom:13-116)
SIG_SCHEMES = {
    1: (2048, "pkcs1", 256, 65537, 0), 2: (4096, "pkcs1", 256, 65537, 0), 3: (2048, "pkcs1", 160, 65537, 0),
    10: (2048, "pss", 256, 3, 32), 11: (2048, "pss", 256, 65537, 32), 12: (2048, "pss", 256, 65537, 64),
    4: (3072, "pkcs1", 160, 37187, 0),
    13: (2048, "pss", 384, 65537, 48), 14: (3072, "pss", 256, 65537, 32),
    20: (256, "ecdsa", 256, 0, 0), 21: (256, "ecdsa", 256, 0, 0), 24: (224, "ecdsa", 224, 0, 0),
}

_KEY_CACHE = {}


def key_pool(bits, count, seed, e=65537):
    """Deterministic pool of RSA keys (the signer certificates of the synthetic state)."""
    k = (bits, count, seed, e)
    if k not in _KEY_CACHE:
        rng = random.Random((seed << 16) ^ bits ^ 0x5A5A ^ (e << 40))
        _KEY_CACHE[k] = [RsaKey(bits, rng, e) for _ in range(count)]
    return _KEY_CACHE[k]


# --------------------------------------------------------------------------- padding / packing
def sha_pad(msg: bytes, block_bits=512) -> bytes:
    """padding() of process_passport.js:11-91 (0x80, zeros, big-endian bit length)."""
    bb = block_bits // 8
    lb = 8 if block_bits == 512 else 16
    pad = (bb - ((len(msg) + 1 + lb) % bb)) % bb
    return msg + b"\x80" + b"\x00" * pad + (len(msg) * 8).to_bytes(lb, "big")


def bytes_to_bits(b: bytes):
    return [(x >> (7 - i)) & 1 for x in b for i in range(8)]


def chunks_le(x: int, n: int, k: int):
    """bigintToArray(n, k, x) of process_passport.js:113-123."""
    mask = (1 << n) - 1
    return [(x >> (n * i)) & mask for i in range(k)]


def rsa_pubkey_hash(n: int):
    """pk_hash of getFakeIdenData (process_passport.js:642-651) ==
    passportVerificationBuilder.circom:182-191."""
    c = chunks_le(n, 64, 15)
    return poseidon([(c[3 * i] << 128) + (c[3 * i + 1] << 64) + c[3 * i + 2] for i in range(5)])


_MRZ = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789<"


def _hash_name(bits):
    return {160: "sha1", 224: "sha224", 256: "sha256", 384: "sha384", 512: "sha512"}[bits]


@dataclass
class Passport:
    """One synthetic passport: raw messages + the circuit input object."""
    dg1: bytes
    dg15: bytes
    ec: bytes
    sa: bytes
    key_index: int
    signature: int
    inputs: dict


class PassportFactory:
    """Seeded generator for one circuit parameter set (RSA PKCS#1 v1.5 families)."""

    def __init__(self, params: CircuitParams = C3, seed: int = 1, n_sig_keys: int = 4,
                 n_aa_keys: int = 4):
        if params.sig_type not in SIG_SCHEMES:
            raise NotImplementedError("synthetic generator: RSA PKCS#1 v1.5 / PSS (SIG 1-4, 10-14) and ECDSA (SIG 20, 21, 24)")
        self.params = params
        self.seed = seed
        self.key_bits, self.scheme, self.sig_hash, self.e, self.salt_len = SIG_SCHEMES[params.sig_type]
        self.block = 512 if self.sig_hash <= 256 else 1024
        if (512 if params.dg_hash <= 256 else 1024) != self.block:
            raise ValueError("DG_HASH_TYPE and the signature hash must share a block size (SURVEY.md appendix D)")
        # the encapsulated content is hashed with the signature hash, except for SIG 24 where it stays SHA-256
        # (passportVerificationBuilder.circom:51-59: EC_HASH_TYPE is fixed before HASH_TYPE becomes 224)
        self.ec_hash = 256 if params.sig_type == 24 else self.sig_hash
        if self.scheme == "ecdsa":
            krng = random.Random((seed << 20) ^ 0xEC)
            self.curve = SIG_CURVES[params.sig_type]
            self.sig_keys = [EcKey(krng, self.curve) for _ in range(n_sig_keys)]
        else:
            self.sig_keys = key_pool(self.key_bits, n_sig_keys, seed, self.e)
        if 0 < params.aa_algo < 20:
            self.aa_keys = key_pool(1024, n_aa_keys, seed + 7)
        elif params.aa_algo >= 20:
            arng = random.Random((seed << 20) ^ 0xAA)
            self.aa_keys = [EcKey(arng, CURVE_P256) for _ in range(n_aa_keys)]   # identity.circom:51-79: x, y at AA_SHIFT
        else:
            self.aa_keys = []
        if self.scheme == "ecdsa":
            self._pkhash = [ec_pubkey_hash(k.x, k.y) for k in self.sig_keys]
        else:
            self._pkhash = [rsa_pubkey_hash(k.n) for k in self.sig_keys]
        self._roots = [poseidon([h, h, 1]) for h in self._pkhash]

    # -- message builders -------------------------------------------------
    def _dg1(self, rng):
        p = self.params
        if p.doc_type == 3:
            body = bytes([0x61, 0x5B, 0x5F, 0x1F, 0x58]) + "".join(
                rng.choice(_MRZ) for _ in range(88)).encode()
        else:
            body = bytes([0x61, 0x5D, 0x5F, 0x1F, 0x5A]) + "".join(
                rng.choice(_MRZ) for _ in range(90)).encode()
        return body

    def _dg15(self, rng):
        p = self.params
        if not p.aa_algo:
            return b""
        key = self.aa_keys[rng.randrange(len(self.aa_keys))]
        if p.aa_algo >= 20:
            # EC active-authentication key: uncompressed point, x at AA_SHIFT, y right behind it
            body = bytes(rng.randrange(256) for _ in range(p.aa_shift // 8 - 1)) + b"\x04" + \
                key.x.to_bytes(32, "big") + key.y.to_bytes(32, "big")
            lo, hi = self._len_range(p.dg15_blocks)
            if len(body) > hi:
                raise ValueError("DG15 does not fit the requested block count")
            return body + bytes(rng.randrange(256) for _ in range(max(0, lo - len(body))))
        # DER SubjectPublicKeyInfo (RSA-1024) inside tag 6F: the modulus starts at byte 32
        hdr = bytes.fromhex("6f81a230819f300d06092a864886f70d010101050003818d0030818902818100")
        body = hdr + key.n.to_bytes(128, "big") + bytes.fromhex("0203010001")
        lead = p.aa_shift // 8 - 32
        if lead < 0:
            raise ValueError("AA_SHIFT below 256 bits is not supported by the generator")
        body = bytes(rng.randrange(256) for _ in range(lead)) + body
        lo, hi = self._len_range(p.dg15_blocks)
        if not (lo <= len(body) <= hi):
            if len(body) < lo:
                body += bytes(rng.randrange(256) for _ in range(lo - len(body)))
            else:
                raise ValueError("DG15 does not fit the requested block count")
        return body

    def _len_range(self, blocks):
        """Message lengths (bytes) whose SHA padding gives exactly `blocks` blocks; the upper
        end keeps clear of len % 64 == 56, see SURVEY.md appendix C.4."""
        bb = self.block // 8
        lb = 8 if self.block == 512 else 16
        return (blocks - 1) * bb - lb + 1, blocks * bb - lb - 1

    def make(self, index: int) -> Passport:
        p = self.params
        rng = random.Random((self.seed << 32) ^ (index * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF))
        dgh = _hash_name(p.dg_hash)
        sgh = _hash_name(self.sig_hash)
        dg1 = self._dg1(rng)
        dg15 = self._dg15(rng)
        hlen = p.dg_hash // 8
        # encapsulated content (LDS security object)
        lo, hi = self._len_range(p.ec_blocks)
        need = p.dg1_shift // 8 + hlen
        if p.aa_algo:
            need = max(need, p.dg15_shift // 8 + hlen)
        lo = max(lo, need)
        if lo > hi:
            raise ValueError("encapsulated content does not fit EC_BLOCK_NUMBER")
        ec = bytearray(rng.randrange(256) for _ in range(rng.randint(lo, hi)))
        ec[0] = 0x30
        o = p.dg1_shift // 8
        ec[o:o + hlen] = hashlib.new(dgh, dg1).digest()
        if p.aa_algo:
            o = p.dg15_shift // 8
            ec[o - 3:o] = bytes([0x0F, 0x04, hlen])
            ec[o:o + hlen] = hashlib.new(dgh, dg15).digest()
        ec = bytes(ec)
        # signed attributes
        slen = self.ec_hash // 8
        lo, hi = self._len_range(1024 // self.block)
        lo = max(lo, p.ec_shift // 8 + slen)
        if lo > hi:
            raise ValueError("signed attributes do not fit their 1024-bit input")
        sa = bytearray(rng.randrange(256) for _ in range(rng.randint(lo, hi)))
        sa[0] = 0x31
        o = p.ec_shift // 8
        sa[o:o + slen] = hashlib.new(_hash_name(self.ec_hash), ec).digest()
        sa = bytes(sa)
        ki = rng.randrange(len(self.sig_keys))
        key = self.sig_keys[ki]
        if self.scheme == "ecdsa":
            sig = key.sign(sa, sgh, rng)
            cb, cn = self.curve.chunk_bits, self.curve.chunks
            pub_chunks = chunks_le(key.x, cb, cn) + chunks_le(key.y, cb, cn)
            sig_chunks = chunks_le(sig[0], cb, cn) + chunks_le(sig[1], cb, cn)
        else:
            sig = pkcs1v15_sign(key, sa, sgh) if self.scheme == "pkcs1" else pss_sign(key, sa, sgh, self.salt_len, rng)
            pub_chunks = chunks_le(key.n, 64, self.key_bits // 64)
            sig_chunks = chunks_le(sig, 64, self.key_bits // 64)
        sk = hashlib.sha256(ec).hexdigest()[:62]
        inputs = {
            "dg1": [str(b) for b in bytes_to_bits(sha_pad(dg1, self.block))],
            "dg15": [str(b) for b in bytes_to_bits(sha_pad(dg15, self.block))] if p.aa_algo else [],
            "signedAttributes": [str(b) for b in bytes_to_bits(sha_pad(sa, self.block))],
            "encapsulatedContent": [str(b) for b in bytes_to_bits(sha_pad(ec, self.block))],
            "pubkey": [str(c) for c in pub_chunks],
            "signature": [str(c) for c in sig_chunks],
            "skIdentity": "0x" + sk,
            "slaveMerkleRoot": "0x" + format(self._roots[ki], "x"),
            "slaveMerkleInclusionBranches": ["0"] * TREE_DEPTH,
        }
        return Passport(dg1, dg15, ec, sa, ki, sig, inputs)

    def batch(self, start: int, count: int):
        return [self.make(start + i) for i in range(count)]
