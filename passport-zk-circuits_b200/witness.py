"""Host-side mirror of the reference's operator surface, over the C ABI (include/pzk.h).

Reference interface mirrored (all of it lives in un-vendored dependencies there):
  * circom's generated `witness_calculator.js`: `calculateWitness(input, sanityCheck)`,
    `calculateWTNSBin(input, sanityCheck)` - call sites
    /root/reference/test/automatisationTest.js:40-50 and
    /root/reference/circuits/scripts/gen-witness.sh:25;
  * `circom_tester`'s `wasm_tester(path)` object: `.calculateWitness`, `.checkConstraints`
    (/root/reference/test/automatisationTest.js:37-51);
  * `snarkjs wtns check`.
Same names, same argument meaning, same error strings.  Python stands in for the Node.js
host layer because this image has no node (INTEGRATION.md shows the N-API binding); all
compute happens in libpzk.so on the GPU - there is no CPU fallback here.
"""
from __future__ import annotations

import ctypes
import json
import lzma
import os
import subprocess
from struct import error as struct_error
from dataclasses import dataclass

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# PZK_LIB_PATH selects another build of the same library (kernel A/B measurements); it is still the CUDA library
LIB_PATH = os.environ.get("PZK_LIB_PATH") or os.path.join(_HERE, "lib", "libpzk.so")
ARTIFACT_DIR = os.path.join(_ROOT, "artifacts")

PZK_ENODEVICE = -3
STATUS_ASSERT, STATUS_CONSTRAINT, STATUS_INPUT_RANGE, STATUS_BIGDIV = 1, 2, 4, 8


class PzkError(RuntimeError):
    pass


# ----------------------------------------------------------------------------- build
def build_library(force=False, verbose=False):
    """nvcc build of libpzk.so for sm_100a (in-tree, so it travels to the GPU box)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, "pzk_api.cu"), os.path.join(csrc, "compiler.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in os.listdir(csrc)] + \
        [os.path.join(_ROOT, "include", f) for f in os.listdir(os.path.join(_ROOT, "include"))]
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(_ROOT, "include"), "-I" + csrc,
           "-o", LIB_PATH] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise PzkError("nvcc failed:\n" + r.stderr[-4000:])
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib():
    """The CUDA library.  Fails loudly when it is missing - nothing falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PzkError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, u32, u64, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64
    L.pzk_version.restype = cp
    L.pzk_device_count.restype = ctypes.c_int
    L.pzk_compile.restype = ctypes.c_int
    L.pzk_compile.argtypes = [cp, cp, ctypes.POINTER(cp), ctypes.POINTER(ctypes.c_int), ctypes.c_int, u32, cp,
                              ctypes.c_size_t]
    L.pzk_compile_ex.restype = ctypes.c_int
    L.pzk_compile_ex.argtypes = [cp, cp, ctypes.POINTER(cp), ctypes.POINTER(ctypes.c_int), ctypes.c_int, u32, u32,
                                 cp, ctypes.c_size_t]
    L.pzk_circuit_open.restype = ctypes.c_int
    L.pzk_circuit_open.argtypes = [cp, ctypes.c_int, ctypes.POINTER(vp)]
    L.pzk_circuit_open_ex.restype = ctypes.c_int
    L.pzk_circuit_open_ex.argtypes = [cp, cp, cp, ctypes.c_int, ctypes.POINTER(vp)]
    L.pzk_circuit_close.argtypes = [vp]
    L.pzk_last_error.restype = cp
    L.pzk_last_error.argtypes = [vp]
    for f in ("pzk_witness_size", "pzk_input_size", "pzk_public_size", "pzk_constraint_count"):
        getattr(L, f).restype = u32
        getattr(L, f).argtypes = [vp]
    L.pzk_circuit_meta_json.restype = cp
    L.pzk_circuit_meta_json.argtypes = [vp]
    L.pzk_circuit_stats.argtypes = [vp] + [ctypes.POINTER(u64)] * 6
    L.pzk_wtns_size.restype = u64
    L.pzk_wtns_size.argtypes = [vp]
    L.pzk_calculate_witness.argtypes = [vp, vp, vp, ctypes.POINTER(u32), ctypes.POINTER(i64)]
    L.pzk_calculate_wtns_bin.argtypes = [vp, vp, vp, ctypes.POINTER(u32), ctypes.POINTER(i64)]
    L.pzk_witness_batch.argtypes = [vp, vp, u64, vp, vp, vp, vp, u64, vp]
    L.pzk_batch_upload.argtypes = [vp, vp, u64]
    L.pzk_batch_run.argtypes = [vp, ctypes.c_int]
    L.pzk_batch_download.argtypes = [vp, vp, vp, vp]
    L.pzk_profile_get.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64)]
    L.pzk_profile_reset.argtypes = [vp]
    L.pzk_profile_enable.argtypes = [vp, ctypes.c_int]
    L.pzk_set_tile_lanes.argtypes = [vp, u64]
    L.pzk_get_tile_lanes.restype = u64
    L.pzk_get_tile_lanes.argtypes = [vp]
    L.pzk_wave_lanes.restype = u64
    L.pzk_wave_lanes.argtypes = [vp]
    L.pzk_packed_stride.restype = u32
    L.pzk_packed_stride.argtypes = [vp]
    L.pzk_packed_layout.argtypes = [vp, vp, vp]
    L.pzk_witness_batch_packed.argtypes = [vp, vp, u64, vp, vp, vp]
    L.pzk_batch_upload_packed.argtypes = [vp, vp, u64]
    L.pzk_wtns_check.argtypes = [cp, vp, u64, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(i64), cp,
                                 ctypes.c_size_t]
    L.pzk_r1cs_check_batch.argtypes = [cp, vp, u64, ctypes.c_int, vp, vp, ctypes.POINTER(ctypes.c_double), cp,
                                       ctypes.c_size_t]
    L.pzk_witness_batch_packed_async.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.pzk_witness_batch_packed_digest.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.pzk_witness_batch_packed_multi.argtypes = [ctypes.POINTER(vp), ctypes.c_int, vp, u64, vp, vp, vp, vp]
    L.pzk_sync.argtypes = [vp]
    L.pzk_batch_set_digest.argtypes = [vp, ctypes.c_int]
    L.pzk_batch_download_digest.argtypes = [vp, vp]
    L.pzk_digest_weight_of.restype = u32
    L.pzk_digest_weight_of.argtypes = [u32]
    L.pzk_r1cs_open.argtypes = [cp, ctypes.c_int, ctypes.POINTER(vp), cp, ctypes.c_size_t]
    L.pzk_r1cs_close.argtypes = [vp]
    for f in ("pzk_r1cs_wires", "pzk_r1cs_constraints"):
        getattr(L, f).restype = u32
        getattr(L, f).argtypes = [vp]
    L.pzk_r1cs_terms.restype = u64
    L.pzk_r1cs_terms.argtypes = [vp]
    L.pzk_r1cs_check.argtypes = [vp, vp, u64, vp, vp, ctypes.POINTER(ctypes.c_double), cp, ctypes.c_size_t]
    L.pzk_r1cs_check_wtns.argtypes = [vp, vp, u64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(i64), cp, ctypes.c_size_t]
    L.pzk_r1cs_check_circuit.argtypes = [vp, vp, vp, u64, vp, vp, ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_double), cp, ctypes.c_size_t]
    _lib = L
    return L


# ----------------------------------------------------------------------------- compile
# declared widths of the reference's input pipeline (process_passport.js writes bits and
# 64-bit chunks, /root/reference/test/process_passport.js:659-672, 590-626)
REGISTER_IDENTITY_BITS = {"dg1": 1, "dg15": 1, "encapsulatedContent": 1, "signedAttributes": 1,
                          "pubkey": 64, "signature": 64}


COMPILE_STATIC_DEF_ROWS, COMPILE_NO_INTRINSICS, COMPILE_NO_TABLE_PROOFS, COMPILE_NO_VIEWS, COMPILE_NO_VECTORIZE = 1, 2, 4, 8, 16
COMPILE_EMIT_O1 = 32


def compile_circuit(main_path, out_prefix, input_bits=None, segment_ops=0, static_def_rows=False, intrinsics=True,
                    table_proofs=True, views=True, vectorize=True, emit_o1=False):
    """circom -> program (.pzkp) + .r1cs + .sym; the role of
    `circom <file> --r1cs --wasm --sym` (/root/reference/circuits/scripts/compile-circuit.sh:34).
    emit_o1: also <prefix>.O1.r1cs / .O1.sym, the system after an O1-style simplification (pzk.h)."""
    L = lib()
    input_bits = input_bits or {}
    names = (ctypes.c_char_p * max(1, len(input_bits)))(*[k.encode() for k in input_bits])
    widths = (ctypes.c_int * max(1, len(input_bits)))(*list(input_bits.values()))
    err = ctypes.create_string_buffer(4096)
    os.makedirs(os.path.dirname(os.path.abspath(out_prefix)), exist_ok=True)
    flags = ((COMPILE_STATIC_DEF_ROWS if static_def_rows else 0) | (0 if intrinsics else COMPILE_NO_INTRINSICS) |
             (0 if table_proofs else COMPILE_NO_TABLE_PROOFS) | (0 if views else COMPILE_NO_VIEWS) |
             (0 if vectorize else COMPILE_NO_VECTORIZE) | (COMPILE_EMIT_O1 if emit_o1 else 0))
    rc = L.pzk_compile_ex(os.fsencode(main_path), os.fsencode(out_prefix), names, widths, len(input_bits),
                          segment_ops, flags, err, len(err))
    if rc != 0:
        raise PzkError("compile failed: " + err.value.decode(errors="replace"))
    return out_prefix + ".pzkp"


def constraint_source(program_path, index):
    """(template, file, line) of a constraint that is checked at run time, from <prefix>.rowsrc written by
    pzk_compile - what the circom runtime prints behind "Assert Failed." (witness_calculator.js: "Error in template
    X line: N").  None when the file is missing or the row is one the compiler discharged (those cannot fail)."""
    import struct
    program_path = os.fsdecode(program_path)
    base = program_path[:-5] if program_path.endswith(".pzkp") else program_path
    path = base + ".rowsrc"
    if index is None or index < 0 or not os.path.exists(path):
        return None
    blob = open(path, "rb").read()
    if blob[:4] != b"PZKS":
        return None
    pos = 4

    def strings(pos):
        n = struct.unpack_from("<I", blob, pos)[0]
        pos += 4
        out = []
        for _ in range(n):
            ln = struct.unpack_from("<I", blob, pos)[0]
            out.append(blob[pos + 4:pos + 4 + ln].decode(errors="replace"))
            pos += 4 + ln
        return out, pos
    files, pos = strings(pos)
    templates, pos = strings(pos)
    n = struct.unpack_from("<I", blob, pos)[0]
    pos += 4
    ent = np.frombuffer(blob, dtype=np.dtype([("row", "<u4"), ("line", "<u4"), ("file", "<u2"), ("tmpl", "<u2")]),
                        count=n, offset=pos)
    k = int(np.searchsorted(ent["row"], index))
    if k >= n or int(ent["row"][k]) != index:
        return None
    e = ent[k]
    fname = files[int(e["file"])] if int(e["file"]) < len(files) else "?"
    return templates[int(e["tmpl"])], fname, int(e["line"])


def pack_artifact(program_path):
    """xz-compress a program for transport (artifacts/ travels with the repo snapshot)."""
    with open(program_path, "rb") as f, lzma.open(program_path + ".xz", "wb", preset=1) as g:
        while True:
            chunk = f.read(1 << 24)
            if not chunk:
                break
            g.write(chunk)
    return program_path + ".xz"


def artifact(name):
    """Path of a prebuilt program in artifacts/ (unpacked on first use)."""
    path = os.path.join(ARTIFACT_DIR, name + ".pzkp")
    if os.path.exists(path):
        return path
    if os.path.exists(path + ".xz"):
        tmp = path + ".tmp%d" % os.getpid()
        with lzma.open(path + ".xz", "rb") as g, open(tmp, "wb") as f:
            while True:
                chunk = g.read(1 << 24)
                if not chunk:
                    break
                f.write(chunk)
        os.replace(tmp, path)
        return path
    raise PzkError(f"artifact {name} is missing from {ARTIFACT_DIR}: run __graft_entry__.build() where "
                   "/root/reference is mounted")


def artifact_r1cs(name):
    """Path of a prebuilt .r1cs in artifacts/ (large ones travel xz-packed)."""
    path = os.path.join(ARTIFACT_DIR, name + ".r1cs")
    if os.path.exists(path):
        return path
    for packed in (path + ".xz",):
        if os.path.exists(packed):
            tmp = path + ".tmp%d" % os.getpid()
            with lzma.open(packed, "rb") as g, open(tmp, "wb") as f:
                while True:
                    chunk = g.read(1 << 24)
                    if not chunk:
                        break
                    f.write(chunk)
            os.replace(tmp, path)
            return path
    raise PzkError(f"{name}.r1cs is missing from {ARTIFACT_DIR}")


# ----------------------------------------------------------------------------- inputs
def _flatten(v, out):
    if isinstance(v, (list, tuple, np.ndarray)):
        for x in v:
            _flatten(x, out)
    elif isinstance(v, str):
        out.append(int(v, 16) if v[:2] in ("0x", "0X") else int(v))
    else:
        out.append(int(v))


def flatten_input(meta, inp: dict) -> bytes:
    """Object keyed by main input signal name -> n_inputs x 32-byte LE field elements, with the
    error behaviour of witness_calculator.js."""
    by_name = {d["name"]: d for d in meta["inputs"]}
    total = sum(d["size"] for d in meta["inputs"])
    buf = bytearray(32 * total)
    n_set = 0
    for name, val in inp.items():
        d = by_name.get(name)
        if d is None:
            raise PzkError(f"Signal not found: {name}")
        flat = []
        _flatten(val, flat)
        if len(flat) < d["size"]:
            raise PzkError(f"Not enough values for input signal {name}")
        if len(flat) > d["size"]:
            raise PzkError(f"Too many values for input signal {name}")
        off = d["offset"]
        for i, x in enumerate(flat):
            buf[32 * (off + i):32 * (off + i + 1)] = (x % P).to_bytes(32, "little")
        n_set += len(flat)
    if n_set != total:
        raise PzkError(f"Not all inputs have been set. Only {n_set} out of {total}")
    return bytes(buf)


def pack_inputs_fast(meta, inputs: list) -> np.ndarray:
    """Batch of input objects -> uint64 array [B, n_inputs, 4] (values < 2^64 take the fast
    path; the general path goes through flatten_input)."""
    total = sum(d["size"] for d in meta["inputs"])
    B = len(inputs)
    out = np.zeros((B, total, 4), dtype=np.uint64)
    for b, inp in enumerate(inputs):
        if set(inp.keys()) != {d["name"] for d in meta["inputs"]}:
            out[b] = np.frombuffer(flatten_input(meta, inp), dtype=np.uint64).reshape(total, 4)
            continue
        for d in meta["inputs"]:
            flat = []
            _flatten(inp[d["name"]], flat)
            if len(flat) != d["size"]:
                raise PzkError(f"{'Not enough' if len(flat) < d['size'] else 'Too many'} values for input signal "
                               f"{d['name']}")
            off = d["offset"]
            if d["bits"] and d["bits"] <= 63:
                out[b, off:off + d["size"], 0] = np.array(flat, dtype=np.uint64)
            else:
                for i, x in enumerate(flat):
                    x %= P
                    for limb in range(4):
                        out[b, off + i, limb] = (x >> (64 * limb)) & 0xFFFFFFFFFFFFFFFF
    return out


@dataclass
class BatchResult:
    status: np.ndarray      # uint32 [B]
    first_bad: np.ndarray   # int64 [B]; -1 when every constraint holds
    public: np.ndarray      # uint64 [B, n_public, 4] canonical little-endian limbs
    witnesses: np.ndarray | None = None  # uint64 [n_export, n_wires, 4]
    digest: np.ndarray | None = None     # uint64 [B, 4] witness digest (set_digest(True))

    def public_ints(self, lane):
        return [int.from_bytes(self.public[lane, i].tobytes(), "little") for i in range(self.public.shape[1])]


class WitnessCalculator:
    """`new WitnessCalculator(wasm)` of circom's witness_calculator.js, with a program instead of
    the wasm.  One instance is reusable across calls; it is not thread-safe."""

    def __init__(self, program_path, device=0, external_sym=None, program_sym=None):
        """external_sym: a .sym written by the circom compiler for this circuit (any optimisation level): the wires
        are then numbered after it (pzk_circuit_open_ex); program_sym defaults to the .sym next to the program."""
        self._L = lib()
        self._h = ctypes.c_void_p()
        self._program_path = program_path
        if external_sym is not None:
            if program_sym is None:
                base = program_path[:-5] if program_path.endswith(".pzkp") else program_path
                program_sym = base + ".sym" if os.path.exists(base + ".sym") else base + ".sym.local"
            rc = self._L.pzk_circuit_open_ex(os.fsencode(program_path), os.fsencode(program_sym), os.fsencode(external_sym),
                                             device, ctypes.byref(self._h))
        else:
            rc = self._L.pzk_circuit_open(os.fsencode(program_path), device, ctypes.byref(self._h))
        if rc != 0:
            msg = self._L.pzk_last_error(self._h).decode() if self._h else ""
            if self._h:
                self._L.pzk_circuit_close(self._h)
                self._h = ctypes.c_void_p()
            if rc == PZK_ENODEVICE:
                raise PzkError("no CUDA device: the witness generator has no CPU fallback")
            raise PzkError(f"pzk_circuit_open failed ({rc}): {msg}")
        self.meta = json.loads(self._L.pzk_circuit_meta_json(self._h).decode())
        self.n_wires = self._L.pzk_witness_size(self._h)
        self.n_inputs = self._L.pzk_input_size(self._h)
        self.n_public = self._L.pzk_public_size(self._h)
        self.n_constraints = self._L.pzk_constraint_count(self._h)

    def close(self):
        if self._h:
            self._L.pzk_circuit_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PzkError(f"pzk error {rc}: {self._L.pzk_last_error(self._h).decode()}")

    def stats(self):
        v = [ctypes.c_uint64() for _ in range(6)]
        self._L.pzk_circuit_stats(self._h, *[ctypes.byref(x) for x in v])
        keys = ("op_records", "f_mul", "f_inv", "rows", "terms", "bytes_per_lane")
        return {k: x.value for k, x in zip(keys, v)}

    # ---- reference operator surface
    def calculateWitness(self, inp, sanityCheck=True):
        """Returns the witness as a list of ints (w[0] == 1).  Raises "Assert Failed." when an
        assert / constraint does not hold, as the wasm does (automatisationTest.js:53-56)."""
        flat = flatten_input(self.meta, inp)
        wit = np.zeros((self.n_wires, 4), dtype=np.uint64)
        st, fb = ctypes.c_uint32(), ctypes.c_int64()
        self._check(self._L.pzk_calculate_witness(self._h, flat, wit.ctypes.data, ctypes.byref(st), ctypes.byref(fb)))
        self._raise_status(st.value, fb.value, sanityCheck)
        raw = wit.tobytes()
        return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(self.n_wires)]

    def calculateWTNSBin(self, inp, sanityCheck=True) -> bytes:
        flat = flatten_input(self.meta, inp)
        out = ctypes.create_string_buffer(self._L.pzk_wtns_size(self._h))
        st, fb = ctypes.c_uint32(), ctypes.c_int64()
        self._check(self._L.pzk_calculate_wtns_bin(self._h, flat, out, ctypes.byref(st), ctypes.byref(fb)))
        self._raise_status(st.value, fb.value, sanityCheck)
        return out.raw

    def _raise_status(self, st, fb, sanity):
        if st & STATUS_INPUT_RANGE:
            raise PzkError("Input out of its declared range (status 4)")
        if st & STATUS_BIGDIV:
            raise PzkError("Assert Failed. long division precondition violated (status 8)")
        if st & (STATUS_ASSERT | STATUS_CONSTRAINT):
            try:
                src = constraint_source(self._program_path, fb) if st & STATUS_CONSTRAINT else None
            except (OSError, ValueError, IndexError, struct_error):
                src = None   # a damaged side file must not hide the verdict
            where = f"\nError in template {src[0]} line: {src[2]} ({os.path.basename(src[1])})" if src else ""
            raise PzkError(f"Assert Failed. (status {st}, first failing constraint {fb}){where}")

    # ---- batched product path
    def calculateWitnessBatch(self, inputs, export_lanes=()) -> BatchResult:
        """inputs: uint64 array [B, n_inputs, 4] (pack_inputs_fast) or list of input objects."""
        if not isinstance(inputs, np.ndarray):
            inputs = pack_inputs_fast(self.meta, list(inputs))
        inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
        B = inputs.shape[0]
        status = np.zeros(B, dtype=np.uint32)
        first_bad = np.zeros(B, dtype=np.int64)
        public = np.zeros((B, self.n_public, 4), dtype=np.uint64)
        lanes = np.ascontiguousarray(np.array(list(export_lanes), dtype=np.uint64))
        wit = np.zeros((len(lanes), self.n_wires, 4), dtype=np.uint64) if len(lanes) else None
        self._check(self._L.pzk_witness_batch(self._h, inputs.ctypes.data, B, status.ctypes.data, first_bad.ctypes.data,
                                              public.ctypes.data, lanes.ctypes.data if len(lanes) else None,
                                              len(lanes), wit.ctypes.data if wit is not None else None))
        first_bad[(status & STATUS_CONSTRAINT) == 0] = -1
        return BatchResult(status, first_bad, public, wit)

    # ---- packed inputs (bits as bytes, 64-bit limbs as 8 bytes, field elements as 32 bytes)
    def packed_layout(self):
        if not hasattr(self, "_layout"):
            kind = np.zeros(self.n_inputs, dtype=np.uint32)
            off = np.zeros(self.n_inputs, dtype=np.uint32)
            self._check(self._L.pzk_packed_layout(self._h, kind.ctypes.data, off.ctypes.data))
            self._layout = (kind, off, int(self._L.pzk_packed_stride(self._h)))
        return self._layout

    def pack(self, inputs: np.ndarray, on_range: str = "raise"):
        """uint64 [B, n_inputs, 4] -> uint8 [B, stride] packed records.

        A packed record has room for one byte / eight bytes per narrow input, so a value that does not fit its
        section cannot cross the boundary; it is rejected HERE instead of being narrowed silently (the kernel's
        IN_U range check only sees what is in the record).  on_range = "raise": PzkError naming the first offending
        lane and input; "mask": returns (packed, bad) with bad[b] = True for lanes holding such a value - pass it to
        calculateWitnessBatchPacked(range_mask=bad), which reports PZK_STATUS_INPUT_RANGE for them exactly as the
        unpacked path does.  Values that fit the section but exceed the declared width (a 'bit' of 2) are flagged
        by the kernel."""
        kind, off, stride = self.packed_layout()
        inputs = np.asarray(inputs)
        B = inputs.shape[0]
        out = np.zeros((B, stride), dtype=np.uint8)
        bad = np.zeros(B, dtype=bool)
        k8 = np.nonzero(kind == 0)[0]
        k64 = np.nonzero(kind == 1)[0]
        narrow = np.nonzero(kind != 2)[0]
        if len(narrow):
            bad |= (inputs[:, narrow, 1:] != 0).any(axis=(1, 2))
        if len(k8):
            bad |= (inputs[:, k8, 0] > 255).any(axis=1)
            out[:, off[k8]] = inputs[:, k8, 0].astype(np.uint8)
        if bad.any() and on_range == "raise":
            b = int(np.nonzero(bad)[0][0])
            wide = [int(k) for k in narrow if inputs[b, k, 1:].any() or (kind[k] == 0 and inputs[b, k, 0] > 255)]
            raise PzkError(f"Input out of its declared range (status {STATUS_INPUT_RANGE}): lane {b}, flattened input "
                           f"{wide[0]} does not fit its packed section")
        if len(k64):
            lo = int(off[k64[0]])
            assert (off[k64] == lo + 8 * np.arange(len(k64))).all()
            out[:, lo:lo + 8 * len(k64)] = np.ascontiguousarray(inputs[:, k64, 0]).view(np.uint8).reshape(B, -1)
        kf = np.nonzero(kind == 2)[0]
        if len(kf):
            lo = int(off[kf[0]])
            assert (off[kf] == lo + 32 * np.arange(len(kf))).all()
            out[:, lo:lo + 32 * len(kf)] = np.ascontiguousarray(inputs[:, kf, :]).view(np.uint8).reshape(B, -1)
        if on_range == "mask":
            return out, bad
        return out

    def calculateWitnessBatchPacked(self, packed: np.ndarray, range_mask=None, digest=False) -> BatchResult:
        """range_mask: the per-lane mask pack(..., on_range="mask") returned; those lanes get INPUT_RANGE.
        digest=True (after set_digest(True)): the result carries the per-lane witness digest."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        B = packed.shape[0]
        status = np.zeros(B, dtype=np.uint32)
        first_bad = np.zeros(B, dtype=np.int64)
        public = np.zeros((B, self.n_public, 4), dtype=np.uint64)
        dig = np.zeros((B, 4), dtype=np.uint64) if digest else None
        if digest:
            self._check(self._L.pzk_witness_batch_packed_digest(self._h, packed.ctypes.data, B, status.ctypes.data,
                                                                first_bad.ctypes.data, public.ctypes.data, dig.ctypes.data))
        else:
            self._check(self._L.pzk_witness_batch_packed(self._h, packed.ctypes.data, B, status.ctypes.data,
                                                         first_bad.ctypes.data, public.ctypes.data))
        self._B = B
        if range_mask is not None:
            status[np.asarray(range_mask, dtype=bool)] |= STATUS_INPUT_RANGE
        first_bad[(status & STATUS_CONSTRAINT) == 0] = -1
        res = BatchResult(status, first_bad, public)
        res.digest = dig
        return res

    def upload_packed(self, packed: np.ndarray):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self._check(self._L.pzk_batch_upload_packed(self._h, packed.ctypes.data, packed.shape[0]))
        self._B = packed.shape[0]

    # ---- device-resident measurement path
    def upload(self, inputs: np.ndarray):
        inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
        self._check(self._L.pzk_batch_upload(self._h, inputs.ctypes.data, inputs.shape[0]))
        self._B = inputs.shape[0]

    def run(self, check_rows=True):
        self._check(self._L.pzk_batch_run(self._h, 1 if check_rows else 0))

    def download(self) -> BatchResult:
        B = self._B
        status = np.zeros(B, dtype=np.uint32)
        first_bad = np.zeros(B, dtype=np.int64)
        public = np.zeros((B, self.n_public, 4), dtype=np.uint64)
        self._check(self._L.pzk_batch_download(self._h, status.ctypes.data, first_bad.ctypes.data, public.ctypes.data))
        first_bad[(status & STATUS_CONSTRAINT) == 0] = -1
        return BatchResult(status, first_bad, public)

    def profile(self, enable=None, reset=False):
        if enable is not None:
            self._L.pzk_profile_enable(self._h, 1 if enable else 0)
        if reset:
            self._L.pzk_profile_reset(self._h)
        out = {}
        for i, k in enumerate(("eval", "check", "export", "run", "digest")):
            ms, n = ctypes.c_double(), ctypes.c_uint64()
            self._L.pzk_profile_get(self._h, i, ctypes.byref(ms), ctypes.byref(n))
            out[k] = (ms.value, n.value)
        return out

    # ---- witness digest (pzk.h): every wire of every lane folded on the device
    def set_digest(self, on=True):
        self._check(self._L.pzk_batch_set_digest(self._h, 1 if on else 0))
        self._digest = bool(on)

    def download_digest(self) -> np.ndarray:
        out = np.zeros((self._B, 4), dtype=np.uint64)
        self._check(self._L.pzk_batch_download_digest(self._h, out.ctypes.data))
        return out

    def sync(self):
        self._check(self._L.pzk_sync(self._h))

    def set_tile_lanes(self, lanes):
        self._check(self._L.pzk_set_tile_lanes(self._h, lanes))

    def tile_lanes(self):
        return self._L.pzk_get_tile_lanes(self._h)

    def wave_lanes(self):
        return self._L.pzk_wave_lanes(self._h)


def wtns_check(r1cs_path, wtns: bytes, device=0):
    """`snarkjs wtns check <r1cs> <wtns>`: (True, -1) or (False, first failing constraint)."""
    L = lib()
    verdict, fb = ctypes.c_int(), ctypes.c_int64()
    err = ctypes.create_string_buffer(1024)
    rc = L.pzk_wtns_check(os.fsencode(r1cs_path), wtns, len(wtns), device, ctypes.byref(verdict), ctypes.byref(fb),
                          err, len(err))
    if rc != 0:
        raise PzkError(err.value.decode(errors="replace") or f"pzk error {rc}")
    return bool(verdict.value), fb.value


def program_histogram(program_path: str) -> dict:
    """Record counts per opcode of a compiled program (host-side read of the file; names from pzk_program.h) plus
    the derived per-witness figures the roofline accounting uses: explicit Fr products, products inside the
    hint intrinsics, one product per quadratic field row, narrow (64-bit) records."""
    import re
    import struct
    blob = open(program_path, "rb").read()
    n_segments = struct.unpack_from("<I", blob, 36)[0]
    n_list = struct.unpack_from("<I", blob, 52)[0]
    n_fpool, n_coef = struct.unpack_from("<II", blob, 40)
    n_rec = struct.unpack_from("<Q", blob, 56)[0]

    def al(x):
        return (x + 15) & ~15
    pos = al(160) + al(n_segments * 48)
    ops = np.frombuffer(blob, dtype=np.uint32, count=n_rec * 4, offset=pos).reshape(-1, 4)
    pos = al(pos + n_rec * 16)
    pos = al(pos + n_fpool * 32)
    pos = al(pos + n_coef * 96)
    lst = np.frombuffer(blob, dtype=np.uint32, count=n_list, offset=pos)
    hdr = open(os.path.join(_ROOT, "include", "pzk_program.h")).read()
    names = {int(m.group(2)): m.group(1) for m in re.finditer(r"PZK_(\w+) = (\d+)", hdr)}
    hist = {}
    quad_field_rows = 0
    bjj_products = 0
    z_imad = 0
    n_dig = 0
    pc = 0
    w0 = ops[:, 0]
    while pc < n_rec:
        w = int(w0[pc])
        opc, fl = w & 0xff, (w >> 8) & 0xff
        nm = names.get(opc, str(opc))
        hist[nm] = hist.get(nm, 0) + 1
        if opc in (56, 57, 58):
            if opc == 57 and (w >> 16) and (int(ops[pc, 2]) & 0xffff):
                quad_field_rows += 1
            pc += int(ops[pc, 3])
        elif opc == 37:
            n = int(lst[int(ops[pc, 2])])
            bjj_products += (2 * n - 1) * (13 + 4)      # 13 per projective addition, 1 prefix + 3 to normalise
        elif opc in (66, 71):                           # Z_MUL / Z_MULADD: la x lb 64-bit limbs = 4 la lb 32-bit multiply-adds
            la, lb = (w >> 16) & 15, (w >> 20) & 15
            z_imad += 4 * la * lb if la and lb else 40
        if fl & 4:
            pc += 1
        if opc == 23 and fl & 32:
            pc += 1
        if fl & 128 and opc not in (56, 57, 58):      # digest descriptor behind the op (PZK_FLAG_DIG)
            pc += 1
            n_dig += 1
        if opc in (70, 71) and fl & 32:               # fused multiply-add whose product is a wire (PZK_FLAG_DIG2)
            pc += 1
            n_dig += 1
        pc += 1
    if pc != n_rec:
        raise PzkError(f"{program_path}: the record walk ends at {pc}, the program has {n_rec} records")
    narrow = sum(v for k, v in hist.items() if k.startswith(("U_", "I_", "V_")) or k in ("N_BIT", "N_LOW", "N_FITS", "IN_U", "CHECK_I64", "CHECK_INT", "CHECK_RANGE",
                                                                                          "Z_ADD", "Z_SUB", "Z_FROM_U", "Z_FROM_I"))
    explicit = hist.get("F_MUL", 0) + hist.get("F_MULADD", 0)
    fr_products = explicit + bjj_products + quad_field_rows
    return {"records": hist, "fr_products": fr_products, "explicit_f_mul": explicit,
            "intrinsic_products": bjj_products, "quadratic_field_rows": quad_field_rows, "narrow_records": narrow,
            "z_mul_records": hist.get("Z_MUL", 0) + hist.get("Z_MULADD", 0), "z_mul_imad": z_imad, "digest_descriptors": n_dig,
            "algorithmic_imad": 136 * fr_products + z_imad + narrow}


def digest_weights(n_wires: int) -> np.ndarray:
    """c(i) of the witness digest (pzk.h: (splitmix64(i) >> 32) | 1) for wires 0..n_wires-1, as uint64."""
    with np.errstate(over="ignore"):
        z = np.arange(n_wires, dtype=np.uint64) + np.uint64(0x9e3779b97f4a7c15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
        return ((z ^ (z >> np.uint64(31))) >> np.uint64(32)) | np.uint64(1)


def witness_digest(witness: np.ndarray) -> np.ndarray:
    """Host restatement of the device digest: witness uint64 [n_wires, 4] canonical -> uint64 [4], the canonical
    little-endian limbs of  sum_i c(i) * w_i mod p.  Exact: the 32-bit halves of every limb are weighted and
    summed separately (32 x 32-bit products, their halves summed in 64 bits), then recombined with Python ints."""
    w = np.ascontiguousarray(witness, dtype=np.uint64)
    c = digest_weights(w.shape[0])
    total = 0
    m32 = np.uint64(0xffffffff)
    for limb in range(4):
        for half in range(2):
            part = (w[:, limb] >> np.uint64(32 * half)) & m32
            prod = part * c                                    # < 2^64, exact
            s = int((prod & m32).sum(dtype=np.uint64)) + (int((prod >> np.uint64(32)).sum(dtype=np.uint64)) << 32)
            total += s << (64 * limb + 32 * half)
    total %= P
    return np.frombuffer(total.to_bytes(32, "little"), dtype=np.uint64).copy()


def witness_batch_packed_multi(calcs, packed: np.ndarray, digest=False) -> "BatchResult":
    """One host thread, several devices: `calcs` are WitnessCalculators of the same program opened on different
    devices; contiguous shares of the batch run concurrently (pzk_witness_batch_packed_multi)."""
    L = lib()
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    B = packed.shape[0]
    c0 = calcs[0]
    status = np.zeros(B, dtype=np.uint32)
    first_bad = np.zeros(B, dtype=np.int64)
    public = np.zeros((B, c0.n_public, 4), dtype=np.uint64)
    dig = np.zeros((B, 4), dtype=np.uint64) if digest else None
    hs = (ctypes.c_void_p * len(calcs))(*[c._h for c in calcs])
    rc = L.pzk_witness_batch_packed_multi(hs, len(calcs), packed.ctypes.data, B, status.ctypes.data, first_bad.ctypes.data,
                                          public.ctypes.data, dig.ctypes.data if digest else None)
    if rc != 0:
        msgs = [L.pzk_last_error(c._h).decode(errors="replace") for c in calcs]
        raise PzkError(next((m for m in msgs if m), f"pzk error {rc}"))
    first_bad[(status & STATUS_CONSTRAINT) == 0] = -1
    res = BatchResult(status, first_bad, public)
    res.digest = dig
    return res


class R1cs:
    """An .r1cs parsed and resident on the device (pzk_r1cs_open): `wtns check` for any number of batches."""

    def __init__(self, r1cs_path, device=0):
        self._L = lib()
        self._h = ctypes.c_void_p()
        err = ctypes.create_string_buffer(1024)
        rc = self._L.pzk_r1cs_open(os.fsencode(r1cs_path), device, ctypes.byref(self._h), err, len(err))
        if rc != 0:
            self._h = None
            raise PzkError(err.value.decode(errors="replace") or f"pzk error {rc}")
        self.n_wires = self._L.pzk_r1cs_wires(self._h)
        self.n_constraints = self._L.pzk_r1cs_constraints(self._h)
        self.n_terms = self._L.pzk_r1cs_terms(self._h)

    def close(self):
        if self._h:
            self._L.pzk_r1cs_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, witnesses: np.ndarray):
        """witnesses: uint64 [B, n_wires, 4] canonical -> (verdicts bool[B], first_bad int64[B], kernel ms)."""
        witnesses = np.ascontiguousarray(witnesses, dtype=np.uint64)
        B = witnesses.shape[0]
        if witnesses.shape[1] != self.n_wires:
            raise PzkError(f"Invalid witness length. Circuit: {self.n_wires}, witness: {witnesses.shape[1]}")
        verdicts = np.zeros(B, dtype=np.int32)
        fb = np.zeros(B, dtype=np.int64)
        ms = ctypes.c_double()
        err = ctypes.create_string_buffer(1024)
        rc = self._L.pzk_r1cs_check(self._h, witnesses.ctypes.data, B, verdicts.ctypes.data, fb.ctypes.data,
                                    ctypes.byref(ms), err, len(err))
        if rc != 0:
            raise PzkError(err.value.decode(errors="replace") or f"pzk error {rc}")
        return verdicts.astype(bool), fb, ms.value

    def check_wtns(self, wtns: bytes):
        verdict, fb = ctypes.c_int(), ctypes.c_int64()
        err = ctypes.create_string_buffer(1024)
        rc = self._L.pzk_r1cs_check_wtns(self._h, wtns, len(wtns), ctypes.byref(verdict), ctypes.byref(fb), err, len(err))
        if rc != 0:
            raise PzkError(err.value.decode(errors="replace") or f"pzk error {rc}")
        return bool(verdict.value), fb.value

    def check_circuit(self, calc: "WitnessCalculator", lanes):
        """calculateWitness -> checkConstraints on the device: evaluates the batch resident in `calc`
        (upload / upload_packed) and checks every row of this .r1cs on the witnesses of `lanes`.
        -> (verdicts bool[n], first_bad int64[n], eval ms, check ms)"""
        lanes = np.ascontiguousarray(np.array(list(lanes), dtype=np.uint64))
        n = len(lanes)
        verdicts = np.zeros(n, dtype=np.int32)
        fb = np.zeros(n, dtype=np.int64)
        t_eval, t_check = ctypes.c_double(), ctypes.c_double()
        err = ctypes.create_string_buffer(1024)
        rc = self._L.pzk_r1cs_check_circuit(self._h, calc._h, lanes.ctypes.data, n, verdicts.ctypes.data, fb.ctypes.data,
                                            ctypes.byref(t_eval), ctypes.byref(t_check), err, len(err))
        if rc != 0:
            raise PzkError(err.value.decode(errors="replace") or f"pzk error {rc}")
        return verdicts.astype(bool), fb, t_eval.value, t_check.value


def r1cs_check_batch(r1cs_path, witnesses: np.ndarray, device=0):
    """witnesses: uint64 [B, n_wires, 4] canonical -> (verdicts bool[B], first_bad int64[B], kernel ms)."""
    L = lib()
    witnesses = np.ascontiguousarray(witnesses, dtype=np.uint64)
    B = witnesses.shape[0]
    verdicts = np.zeros(B, dtype=np.int32)
    fb = np.zeros(B, dtype=np.int64)
    ms = ctypes.c_double()
    err = ctypes.create_string_buffer(1024)
    rc = L.pzk_r1cs_check_batch(os.fsencode(r1cs_path), witnesses.ctypes.data, B, device, verdicts.ctypes.data,
                                fb.ctypes.data, ctypes.byref(ms), err, len(err))
    if rc != 0:
        raise PzkError(err.value.decode(errors="replace") or f"pzk error {rc}")
    return verdicts.astype(bool), fb, ms.value


class wasm_tester:
    """Shape of circom_tester's object (/root/reference/test/automatisationTest.js:37-51):
    `circuit = wasm_tester(path); w = circuit.calculateWitness(input); circuit.checkConstraints(w)`.
    `path` is a prebuilt program prefix (".pzkp/.r1cs/.sym") or a .circom file to compile."""

    def __init__(self, path, input_bits=None, device=0, workdir=None):
        if path.endswith(".circom"):
            workdir = workdir or os.path.join(ARTIFACT_DIR, "_tester")
            prefix = os.path.join(workdir, os.path.splitext(os.path.basename(path))[0])
            compile_circuit(path, prefix, input_bits)
        else:
            prefix = path[:-5] if path.endswith(".pzkp") else path
        self.prefix = prefix
        self.calc = WitnessCalculator(prefix + ".pzkp", device)
        self.device = device
        self._sym = None

    def calculateWitness(self, inp, sanityCheck=True):
        return self.calc.calculateWitness(inp, sanityCheck)

    def checkConstraints(self, witness):
        w = np.zeros((1, len(witness), 4), dtype=np.uint64)
        raw = b"".join(int(x).to_bytes(32, "little") for x in witness)
        w[0] = np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4)
        ok, fb, _ = r1cs_check_batch(self.prefix + ".r1cs", w, self.device)
        if not ok[0]:
            raise PzkError(f"Constraint doesn't match (constraint {int(fb[0])})")
        return True

    def symbols(self):
        """name -> witness index, from the .sym file (map by name)."""
        if self._sym is None:
            self._sym = {}
            with open(self.prefix + ".sym") as f:
                for line in f:
                    lab, wi, ci, name = line.rstrip("\n").split(",", 3)
                    self._sym[name] = int(wi)
        return self._sym
