"""Real-passport front end: passport JSON (dg1, dg15, sod) -> circuit parameters + circuit input object.

Host-side mirror of `processPassport` (/root/reference/test/process_passport.js:674-816) and of the
extraction helpers it calls (:269-571), on top of a small DER reader that produces the node shape the
reference's asn1.js gives them (name, sub, content, dump; OCTET / BIT STRINGs that encapsulate DER are
opened).  The output feeds the batched path unchanged: `CircuitParams` names the program to compile or
load, `inputs` is what `WitnessCalculator.calculateWitnessBatch` takes.

Deliberate deviation, documented: the reference hands AA_SHIFT to `writeToCircom` in BYTES while the
circuit (identity.circom) and the reference's own circuit names use bits (`..._3_256` for a modulus that
starts at byte 32); this port returns bits everywhere.
"""
from __future__ import annotations

import base64
import binascii
import hashlib
import re

from .passports import TREE_DEPTH, CircuitParams, bytes_to_bits, chunks_le, sha_pad
from .poseidon import poseidon

_RE_HEX = re.compile(r"^\s*(?:[0-9A-Fa-f][0-9A-Fa-f]\s*)+$")
_UNIVERSAL = {1: "BOOLEAN", 2: "INTEGER", 3: "BIT_STRING", 4: "OCTET_STRING", 5: "NULL", 6: "OBJECT_IDENTIFIER",
              12: "UTF8String", 16: "SEQUENCE", 17: "SET", 19: "PrintableString", 22: "IA5String", 23: "UTCTime",
              24: "GeneralizedTime"}
_HASH_BY_LEN = {20: "sha1", 28: "sha224", 32: "sha256", 48: "sha384", 64: "sha512"}


class Asn1Error(ValueError):
    pass


class Node:
    """One TLV.  `name` follows asn1.js ("SEQUENCE", "OCTET_STRING", "[0]", "Application_23" ...),
    `content` is the value bytes (for a BIT STRING: without the unused-bits octet), `dump` the whole TLV."""
    __slots__ = ("tag_class", "constructed", "number", "name", "content", "dump", "sub", "unused_bits")

    def __init__(self):
        self.sub = []
        self.unused_bits = 0

    @property
    def length(self):
        return len(self.content)

    def int_value(self):
        return int.from_bytes(self.content, "big", signed=True)

    def walk(self, parent=None):
        yield self, parent
        for c in self.sub:
            yield from c.walk(self)


def _read_tlv(buf: bytes, pos: int, end: int):
    if pos + 2 > end:
        raise Asn1Error("truncated TLV")
    t = buf[pos]
    cls, constructed, number = t >> 6, bool(t & 0x20), t & 0x1F
    p = pos + 1
    if number == 0x1F:
        number = 0
        while True:
            if p >= end:
                raise Asn1Error("truncated tag")
            number = (number << 7) | (buf[p] & 0x7F)
            p += 1
            if not buf[p - 1] & 0x80:
                break
    if p >= end:
        raise Asn1Error("truncated length")
    ln = buf[p]
    p += 1
    if ln & 0x80:
        k = ln & 0x7F
        if k == 0 or k > 4 or p + k > end:
            raise Asn1Error("unsupported length form")
        ln = int.from_bytes(buf[p:p + k], "big")
        p += k
    if p + ln > end:
        raise Asn1Error("length runs past the buffer")
    return cls, constructed, number, p, p + ln


def _try_children(buf: bytes, start: int, end: int):
    """All of [start, end) as a run of TLVs, or None when it does not parse exactly."""
    out, p = [], start
    try:
        while p < end:
            n, p = _decode(buf, p, end)
            out.append(n)
    except Asn1Error:
        return None
    return out if p == end and out else None


def _decode(buf: bytes, pos: int, end: int):
    cls, constructed, number, vs, ve = _read_tlv(buf, pos, end)
    n = Node()
    n.tag_class, n.constructed, n.number = cls, constructed, number
    if cls == 0:
        n.name = _UNIVERSAL.get(number, f"Universal_{number}")
    elif cls == 2:
        n.name = f"[{number}]"
    elif cls == 1:
        n.name = f"Application_{number}"
    else:
        n.name = f"Private_{number}"
    n.dump = bytes(buf[pos:ve])
    n.content = bytes(buf[vs:ve])
    if constructed:
        kids = _try_children(buf, vs, ve)
        if kids is None and ve > vs:
            raise Asn1Error("constructed value does not parse")
        n.sub = kids or []
    elif cls == 0 and number == 4:          # OCTET STRING that encapsulates DER
        looks = ve - vs >= 2 and buf[vs] in (0x30, 0x31)
        n.sub = (_try_children(buf, vs, ve) or []) if looks else []
    elif cls == 0 and number == 3 and ve > vs:  # BIT STRING: first octet = unused bits
        n.unused_bits = buf[vs]
        n.content = bytes(buf[vs + 1:ve])
        looks = ve - vs >= 3 and buf[vs] == 0 and buf[vs + 1] in (0x30, 0x31)
        n.sub = (_try_children(buf, vs + 1, ve) or []) if looks else []
    return n, ve


def decoded(data) -> Node:
    """`decoded(json.sod)` of asn1.js: hex or (armored) base64 text, or raw bytes -> tree."""
    if isinstance(data, str):
        raw = to_bytes(data)
    else:
        raw = bytes(data)
    node, end = _decode(raw, 0, len(raw))
    if end != len(raw):
        raise Asn1Error("trailing bytes after the ASN.1 value")
    return node


def to_bytes(text: str) -> bytes:
    if _RE_HEX.match(text):
        return binascii.unhexlify(re.sub(r"\s+", "", text))
    body = re.sub(r"-----[A-Z ]+-----", "", text)
    return base64.b64decode(re.sub(r"\s+", "", body))


def compute_hash(out_len: int, data: bytes) -> bytes:
    """computeHash(outLen, input), process_passport.js:93-111."""
    if out_len not in _HASH_BY_LEN:
        raise ValueError("Invalid hash output length. Use 20, 28, 32, 48, or 64 bytes.")
    return hashlib.new(_HASH_BY_LEN[out_len], data).digest()


# ---- extraction (process_passport.js:269-571) ----------------------------------------------------------
def get_first_octet_string(asn1: Node):
    for n, _ in asn1.walk():
        if n.name == "OCTET_STRING":
            return n
    return None


def extract_encapsulated_content(asn1: Node):
    ec = get_first_octet_string(asn1)
    if ec is None or not ec.sub:
        raise Asn1Error("SOD: no encapsulated LDS security object")
    hash_type = ec.sub[0].sub[2].sub[0].sub[1].length
    return ec.content, hash_type


def get_zero(asn1: Node):
    """The signed attributes: the first [0] whose last element is SEQUENCE {OID, SET {OCTET STRING}} -
    the messageDigest attribute (process_passport.js:321-358)."""
    for n, _ in asn1.walk():
        if n.name != "[0]" or not n.sub:
            continue
        last = n.sub[-1]
        if last.name == "SEQUENCE" and len(last.sub) == 2 and last.sub[0].name == "OBJECT_IDENTIFIER" \
                and last.sub[1].name == "SET" and len(last.sub[1].sub) == 1 \
                and last.sub[1].sub[0].name == "OCTET_STRING":
            return n
    return None


def extract_signed_attributes(asn1: Node):
    sa = get_zero(asn1)
    if sa is None:
        raise Asn1Error("SOD: signed attributes not found")
    hash_type = sa.sub[-1].sub[-1].sub[0].length
    return b"\x31" + sa.dump[1:], hash_type        # the [0] IMPLICIT tag is hashed as SET OF


def find_parent_of_last_octet_string(asn1: Node):
    result = parent = None
    for n, p in asn1.walk():
        if n.name == "OCTET_STRING":
            result, parent = n, p
    return result, parent


def extract_signature(asn1: Node):
    octet, parent = find_parent_of_last_octet_string(asn1)
    if octet is None or parent is None:
        raise Asn1Error("SOD: signature not found")
    salt = 0
    try:                                        # RSASSA-PSS-params ... [2] saltLength INTEGER
        alg = parent.sub[-2]
        leaf = alg.sub[-1].sub[-1].sub[0]
        if leaf.name == "INTEGER":
            salt = leaf.int_value()
    except (IndexError, AttributeError):
        salt = 0
    if octet.sub:                               # ECDSA: OCTET STRING { SEQUENCE { r, s } }
        return {"r": octet.sub[0].sub[0].int_value(), "s": octet.sub[0].sub[1].int_value()}
    return {"n": int.from_bytes(octet.content, "big"), "salt": salt}


def extract_rsa_pubkey(asn1: Node):
    for n, _ in asn1.walk():
        if n.name == "BIT_STRING":
            for c in n.sub:
                if c.name == "SEQUENCE" and len(c.sub) == 2 and c.sub[0].name == "INTEGER" and c.sub[1].name == "INTEGER":
                    return {"n": c.sub[0].int_value(), "exp": c.sub[1].int_value()}
    raise Asn1Error("SOD: RSA public key not found")


_NAMED_CURVE_A = {  # OID value bytes -> coefficient a (the reference matches explicit parameters by a)
    bytes.fromhex("2a8648ce3d030107"): "FFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFC",  # prime256v1
    bytes.fromhex("2b2403030208010107"): "7D5A0975FC2C3057EEF67530417AFFE7FB8055C126DC5C6CE94A4B44F330B5D9",  # brainpoolP256r1
}


def extract_ecdsa_pubkey(asn1: Node):
    for n, _ in asn1.walk():
        if len(n.sub) >= 2 and n.sub[1].name == "BIT_STRING" and n.sub[1].content[:1] == b"\x04":
            pt = n.sub[1].content[1:]
            x, y = pt[:len(pt) // 2], pt[len(pt) // 2:]
            params = n.sub[0].sub[1]
            if params.sub:                      # explicit ECParameters: curve SEQUENCE { a, b [, seed] }
                a_hex = params.sub[2].sub[0].content.hex().upper()
            else:                               # named curve (extension: the reference only knows secp521r1 by name)
                a_hex = _NAMED_CURVE_A.get(params.content, "")
            return {"x": int.from_bytes(x, "big"), "y": int.from_bytes(y, "big"), "param": a_hex, "bytes": len(x)}
    raise Asn1Error("SOD: ECDSA public key not found")


def get_sig_type(pk, sig, hash_type: int) -> int:
    """getSigType, process_passport.js:157-244."""
    if "r" in sig:
        return {"7D5A0975FC2C3057EEF67530417AFFE7FB8055C126DC5C6CE94A4B44F330B5D9": 21,
                "FFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFC": 20,
                "FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFFFFFFFFFFFFFFFFFFFE": 24}.get(pk["param"], 0)
    n_hex = len("%x" % pk["n"])
    e, salt = pk["exp"], sig["salt"]
    if salt:
        table = {(512, 3, 32, 32): 10, (512, 65537, 32, 32): 11, (512, 65537, 64, 32): 12,
                 (512, 65537, 48, 48): 13, (768, 65537, 32, 32): 14}
        return table.get((n_hex, e, salt, hash_type), 0)
    return {(512, 65537, 32): 1, (1024, 65537, 32): 2, (512, 65537, 20): 3}.get((n_hex, e, hash_type), 0)


def extract_from_dg15(dg15: bytes):
    """extractFromDg15, process_passport.js:497-571: (public key, AA shift in BYTES, AA signature type)."""
    if not dg15:
        return None, 0, 0
    d = decoded(dg15)
    spki = d.sub[0]
    key = spki.sub[1]
    if key.content[:1] == b"\x04" and not key.sub:
        pt = key.content[1:]
        x = pt[:len(pt) // 2]
        p_hex = spki.sub[0].sub[1].sub[1].sub[1].content.hex().upper().lstrip("0") if spki.sub[0].sub[1].sub else ""
        aa_type = {"A9FB57DBA1EEA9BC3E660A909D838D718C397AA3B561A6F7901E0E82974856A7": 21,
                   "FFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF": 20,
                   "FFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEFFFFFFFFFFFFFFFF": 23}.get(p_hex, 0)
        return {"x": int.from_bytes(x, "big")}, d.dump.index(x), aa_type
    rsa_key = key.sub[0]
    n = rsa_key.sub[0].int_value()
    n_bytes = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return {"n": n, "exp": rsa_key.sub[1].int_value()}, d.dump.index(n_bytes), 1


def _blocks(n_bytes: int, block_bits: int) -> int:
    return -(-(n_bytes + 8) // (block_bits // 8))    # Math.ceil((len + 8) / 64|128), process_passport.js:771-790


def process_passport(passport: dict):
    """processPassport(filePath) on an already parsed JSON object {dg1, dg15, sod}.
    Returns (CircuitParams, inputs, name): the RegisterIdentityBuilder parameters writeToCircom would emit,
    the object writeToJson would write, and the reference's circuit name."""
    dg1 = to_bytes(passport["dg1"]) if passport.get("dg1") else b""
    dg15 = to_bytes(passport["dg15"]) if passport.get("dg15") else b""
    asn1 = decoded(passport["sod"])
    ec, dg_hash_len = extract_encapsulated_content(asn1)
    sa, hash_len = extract_signed_attributes(asn1)
    dg_block = 512 if dg_hash_len <= 32 else 1024
    sa_block = 512 if hash_len <= 32 else 1024
    sig = extract_signature(asn1)
    pk = extract_ecdsa_pubkey(asn1) if "r" in sig else extract_rsa_pubkey(asn1)
    sig_type = get_sig_type(pk, sig, hash_len)
    if sig_type == 0:
        raise NotImplementedError("UNKNOWN TECHNOLOGY: signature scheme not covered by the circuits")
    dg1_shift = ec.index(compute_hash(dg_hash_len, dg1)) * 8
    ec_shift = sa.index(compute_hash(hash_len, ec)) * 8
    dg15_shift = ec.index(compute_hash(dg_hash_len, dg15)) * 8 if dg15 else 0
    _aa_pk, aa_shift_bytes, aa_sig_type = extract_from_dg15(dg15)
    if "r" in sig:
        k = -(-pk["bytes"] // 8)
        pub = chunks_le(pk["x"], 64, k) + chunks_le(pk["y"], 64, k)
        sg = chunks_le(sig["r"], 64, k) + chunks_le(sig["s"], 64, k)
        m = (1 << 248) - 1
        pk_hash = poseidon([pk["x"] & m, pk["y"] & m]) if pk["bytes"] * 2 > 62 else poseidon([pk["x"], pk["y"]])
    else:
        k = -(-len("%x" % pk["n"]) // 16)
        pub, sg = chunks_le(pk["n"], 64, k), chunks_le(sig["n"], 64, k)
        arr = chunks_le(pk["n"], 64, 15)
        pk_hash = poseidon([(arr[3 * i] << 128) + (arr[3 * i + 1] << 64) + arr[3 * i + 2] for i in range(5)])
    root = poseidon([pk_hash, pk_hash, 1])
    params = CircuitParams(sig_type, dg_hash_len * 8, 3 if len(dg1) == 93 else 1, _blocks(len(ec), sa_block),
                           ec_shift, dg1_shift, aa_sig_type, dg15_shift,
                           _blocks(len(dg15), dg_block) if dg15 else 0, aa_shift_bytes * 8)
    inputs = {
        "dg1": [str(b) for b in bytes_to_bits(sha_pad(dg1, dg_block))],
        "dg15": [str(b) for b in bytes_to_bits(sha_pad(dg15, dg_block))] if dg15 else [],
        "signedAttributes": [str(b) for b in bytes_to_bits(sha_pad(sa, sa_block))],
        "encapsulatedContent": [str(b) for b in bytes_to_bits(sha_pad(ec, sa_block))],
        "pubkey": [str(c) for c in pub],
        "signature": [str(c) for c in sg],
        "skIdentity": "0x" + hashlib.sha256(ec).hexdigest()[:62],
        "slaveMerkleRoot": "0x" + format(root, "x"),
        "slaveMerkleInclusionBranches": ["0"] * TREE_DEPTH,
    }
    name = "registerIdentity_" + "_".join(str(v) for v in (
        params.sig_type, params.dg_hash, params.doc_type, params.ec_blocks, params.ec_shift, params.dg1_shift)) + \
        ("_NA" if not dg15 else f"_{params.aa_algo}_{params.dg15_shift}_{params.dg15_blocks}_{params.aa_shift}")
    return params, inputs, name
