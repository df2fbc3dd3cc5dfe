"""Builds the compiled programs under artifacts/ (git-ignored, shipped with the repo snapshot).

Reference circuits are compiled from /root/reference when it is mounted (this container);
the GPU box only uses the prebuilt files.  The root files generated here are the 3-line
`component main = T(...)` wrappers that test/process_passport.js:573-588 (writeToCircom)
writes for the reference - no reference source is copied."""
from __future__ import annotations

import os

from . import witness as W
from .passports import C3, CircuitParams

REFERENCE = "/root/reference"
_TESTS = os.path.join(W._ROOT, "tests", "circuits")


def _wrapper(name, body):
    d = os.path.join(W.ARTIFACT_DIR, "_mains")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name + ".circom")
    with open(path, "w") as f:
        f.write(body)
    return path


# SIGNATURE_TYPE 3 (RSA-2048 PKCS#1 v1.5 + SHA-1, SHA-1 data groups), 10 (RSA-PSS e=3, SHA-256),
# 13 (RSA-PSS SHA-384 with 1024-bit hash blocks and a different EC shift / block counts),
# 20 (ECDSA over NIST P-256 + SHA-256: 5.5 M constraints, long_div2 / mod_inv hint functions)
# the parameter set process_passport.py extracts from a CMS SignedData SOD as cryptography's PKCS#7 builder lays it
# out (RSA-2048 + SHA-256 signer, LDS security object with DG1 / DG2 / DG15): the front-end test circuit
CMS_PARAMS = CircuitParams(1, 256, 3, 3, 600, 240, 1, 864, 3, 256)

C4_VARIANTS = {
    "c4_sig3": CircuitParams(3, 160, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig10": CircuitParams(10, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig13": CircuitParams(13, 384, 3, 2, 320, 248, 1, 1496, 2, 256),
    "c4_sig20": CircuitParams(20, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    # the remaining arms of the dispatch (signatureVerification.circom:26-116, identity.circom:26-84): RSA-4096,
    # RSA-3072 with e = 37187 (48 chunks: the non-Karatsuba multiplier), PSS with e = 65537 / salt 64 / 3072 bits,
    # brainpoolP256r1, secp224r1 (7 chunks of 32 bits, SHA-224 over the signed attributes), a document without
    # DG15 (AA_SIGNATURE_ALGO = 0), an EC active-authentication key in DG15, a TD1 document
    "c4_sig2": CircuitParams(2, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig4": CircuitParams(4, 160, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig11": CircuitParams(11, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig12": CircuitParams(12, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig14": CircuitParams(14, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig21": CircuitParams(21, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_sig24": CircuitParams(24, 256, 3, 4, 600, 248, 1, 1496, 3, 256),
    "c4_na": CircuitParams(1, 256, 3, 3, 600, 248, 0, 0, 0, 0),
    "c4_ecaa": CircuitParams(1, 256, 3, 4, 600, 248, 20, 1496, 2, 256),
    "c4_td1": CircuitParams(1, 256, 1, 4, 600, 248, 1, 1496, 3, 256),
}


def reference_circuits():
    lib = REFERENCE + "/circuits/lib/circuits"
    return {
        # config 1: Poseidon + SparseMerkleTree inclusion
        "smt80": (f'pragma circom 2.1.6;\ninclude "{REFERENCE}/circuits/merkleTree/SMTVerifier.circom";\n'
                  "component main {public [root]} = SMTVerifier(80);\n", {}),
        "poseidon2": (f'pragma circom 2.1.6;\ninclude "{lib}/hasher/poseidon/poseidon.circom";\n'
                      "component main = PoseidonHash(2);\n", {}),
        "sha256_1": (f'pragma circom 2.1.6;\ninclude "{lib}/hasher/hash.circom";\n'
                     "component main = ShaHashChunks(1, 256);\n", {"in": 1}),
        "babyjub": (f'pragma circom 2.1.6;\ninclude "{lib}/babyjubjub/curve.circom";\n'
                    "component main = BabyjubjubBase8Multiplication();\n", {}),
        "rsa2048": (f'pragma circom 2.1.6;\ninclude "{lib}/signatures/rsa.circom";\n'
                    "component main = RsaVerifyPkcs1v15(64, 32, 65537, 256);\n",
                    {"signature": 64, "pubkey": 64, "hashed": 1}),
        # config 2: queryIdentity (selective disclosure, date utilities, enforced SMT inclusion); BabyPbk is
        # bound to the in-tree multiplication by tests/circuits/shims/babypbk.circom
        "query80": (f'pragma circom 2.1.6;\ninclude "{W._ROOT}/tests/circuits/shims/babypbk.circom";\n'
                    f'include "{REFERENCE}/circuits/identityManagement/queryIdentity.circom";\n'
                    "component main { public [eventID, eventData, idStateRoot, selector, currentDate, "
                    "timestampLowerbound, timestampUpperbound, identityCounterLowerbound, identityCounterUpperbound, "
                    "birthDateLowerbound, birthDateUpperbound, expirationDateLowerbound, expirationDateUpperbound, "
                    "citizenshipMask] } = QueryIdentity(80);\n", {"dg1": 1}),
        # config 2 for TD1 identity cards (760-bit DG1, 190-bit commitment chunks, hashed document / personal numbers)
        "query80_td1": (f'pragma circom 2.1.6;\ninclude "{W._ROOT}/tests/circuits/shims/babypbk.circom";\n'
                        f'include "{REFERENCE}/circuits/identityManagement/queryIdentityTD1.circom";\n'
                        "component main { public [eventID, eventData, idStateRoot, selector, currentDate, "
                        "timestampLowerbound, timestampUpperbound, identityCounterLowerbound, identityCounterUpperbound, "
                        "birthDateLowerbound, birthDateUpperbound, expirationDateLowerbound, expirationDateUpperbound, "
                        "citizenshipMask] } = QueryIdentity(80);\n", {"dg1": 1}),
        # one P-256 point doubling of the ECDSA verifier (lib/circuits/ec/curve.circom:281-313): mod_inv, long_div,
        # long_div2 / short_div with their data-dependent returns, PointOnTangent / PointOnCurve constraints
        "p256dbl": (f'pragma circom 2.1.6;\ninclude "{lib}/ec/curve.circom";\n'
                    "component main = EllipticCurveDouble(64, 4, [18446744073709551612, 4294967295, 0, "
                    "18446744069414584321], [4309448131093880907, 7285987128567378166, 12964664127075681980, "
                    "6540974713487397863], [18446744073709551615, 4294967295, 0, 18446744069414584321]);\n", {"in": 64}),
        # config 3: the north-star circuit (hardhat.config.ts:29)
        "c3": (C3.main_source(REFERENCE + "/circuits/identityManagement/registerIdentityBuilder.circom"),
               W.REGISTER_IDENTITY_BITS),
        # config 3 again, compiled with the rows of `x <== e` discharged at compile time (pzk.h:
        # PZK_COMPILE_STATIC_DEF_ROWS); same wires, same verdicts, 353 k instead of 1.29 M run-time rows
        "c3_lean": (C3.main_source(REFERENCE + "/circuits/identityManagement/registerIdentityBuilder.circom"),
                    W.REGISTER_IDENTITY_BITS),
        "c3_cms": (CMS_PARAMS.main_source(REFERENCE + "/circuits/identityManagement/registerIdentityBuilder.circom"),
                   W.REGISTER_IDENTITY_BITS),
        # config 3 with alias proofs only (PZK_COMPILE_NO_TABLE_PROOFS): every other row runs on the device
        "c3_allrows": (C3.main_source(REFERENCE + "/circuits/identityManagement/registerIdentityBuilder.circom"),
                       W.REGISTER_IDENTITY_BITS),
        # config 4: SHA-1 / RSA-PSS variants with other hash types and shifts
        **{name: (prm.main_source(REFERENCE + "/circuits/identityManagement/registerIdentityBuilder.circom"),
                  W.REGISTER_IDENTITY_BITS) for name, prm in C4_VARIANTS.items()},
    }


OWN_CIRCUITS = {"t_mix": ("mix.circom", {"u": 16, "bits": 1}),
                "t_bigdiv": ("bigdiv.circom", {"a": 64, "b": 64}),
                "t_earlyret": ("earlyret.circom", {"v": 16, "a": 64, "b": 64, "c": 1}),
                "t_modinv": ("modinv.circom", {"a": 64}),
                "t_muladd": ("muladd.circom", {"a": 64, "b": 64, "c": 64})}

COMPILE_OPTS = {"c3_lean": {"static_def_rows": True},
                "c3_allrows": {"table_proofs": False, "segment_ops": 16384},
                # the small circuits also get the O1-simplified system (<name>.O1.r1cs / .O1.sym, pzk.h PZK_COMPILE_EMIT_O1)
                "smt80": {"emit_o1": True}, "sha256_1": {"emit_o1": True}, "poseidon2": {"emit_o1": True},
                "query80": {"emit_o1": True}}

BIG = {"c3", "c3_lean", "c3_allrows", "c3_cms"} | set(C4_VARIANTS)  # ship only the xz-packed program for these


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_all(verbose=True, only=None):
    os.makedirs(W.ARTIFACT_DIR, exist_ok=True)
    # programs depend on the compiler, not on the kernels that share the library with it
    csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
    deps = [os.path.join(csrc, f) for f in ("compiler.cpp", "compiler.hpp", "circom_front.hpp", "u256.hpp")]
    deps.append(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pzk_program.h"))
    for name, (fname, bits) in OWN_CIRCUITS.items():
        prefix = os.path.join(W.ARTIFACT_DIR, name)
        src = os.path.join(_TESTS, fname)
        if _stale(prefix + ".pzkp", deps + [src]):
            if verbose:
                print("compiling", name)
            W.compile_circuit(src, prefix, bits, emit_o1=True)
    if not os.path.isdir(REFERENCE):
        return
    for name, (body, bits) in reference_circuits().items():
        if only is not None and name not in only:
            continue
        prefix = os.path.join(W.ARTIFACT_DIR, name)
        final = prefix + (".pzkp.xz" if name in BIG else ".pzkp")
        if not _stale(final, deps):
            continue
        if verbose:
            print("compiling", name)
        main = _wrapper(name, body)
        W.compile_circuit(main, prefix, bits, **COMPILE_OPTS.get(name, {}))
        if name in BIG:
            W.pack_artifact(prefix + ".pzkp")
            if name == "c3":  # the stand-alone R1CS stream kernel is measured on this one
                import lzma
                with open(prefix + ".r1cs", "rb") as f, lzma.open(prefix + ".r1cs.xz", "wb", preset=1) as g:
                    while True:
                        chunk = f.read(1 << 24)
                        if not chunk:
                            break
                        g.write(chunk)
            for ext in (".r1cs", ".sym"):
                # too large to ship unpacked; regenerate with compile_circuit when needed here
                if os.path.exists(prefix + ext):
                    os.replace(prefix + ext, prefix + ext + ".local")


def program_for(params: CircuitParams, compile_if_missing: bool = True) -> str:
    """Circuit parameters (what process_passport.py extracts from a document) -> path of the compiled program:
    a prebuilt artifact when one was built for exactly these parameters, else a program compiled on first use
    and cached in artifacts/ under the reference's circuit name (needs the reference's circom sources)."""
    known = {"c3": C3, "c3_cms": CMS_PARAMS, **C4_VARIANTS}
    for name, prm in known.items():
        if prm == params:
            return W.artifact(name)
    prefix = os.path.join(W.ARTIFACT_DIR, params.name)
    if os.path.exists(prefix + ".pzkp"):
        return prefix + ".pzkp"
    if not compile_if_missing or not os.path.isdir(REFERENCE):
        raise W.PzkError(f"no compiled program for {params.name} and the reference's circom sources are not available")
    os.makedirs(os.path.join(W.ARTIFACT_DIR, "_mains"), exist_ok=True)
    main = os.path.join(W.ARTIFACT_DIR, "_mains", params.name + ".circom")
    with open(main, "w") as f:
        f.write(params.main_source(REFERENCE + "/circuits/identityManagement/registerIdentityBuilder.circom"))
    return W.compile_circuit(main, prefix, W.REGISTER_IDENTITY_BITS)
