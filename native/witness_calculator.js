// witness_calculator.js - the JS face of the shim, shaped like circom's generated witness_calculator.js and
// circom_tester's wasm_tester (what /root/reference/test/automatisationTest.js:37-51 and
// circuits/scripts/gen-witness.sh:25 call).  Not runnable in this image (no node); the Python mirror with the
// same checks and messages is passport-zk-circuits_b200/witness.py and is what the tests exercise.
"use strict";
const pzk = require("./build/Release/pzk.node");

const P = 21888242871839275222246405745257275088548364400416034343698204186575808495617n;

function flatten(v, out) {
  if (Array.isArray(v)) for (const x of v) flatten(x, out); else out.push(BigInt(v));
  return out;
}

class WitnessCalculator {
  constructor(programPath, device = 0, programSym = undefined, externalSym = undefined) {
    this.h = pzk.open(programPath, device, programSym, externalSym);
    this.meta = JSON.parse(pzk.meta(this.h));
    this.nInputs = this.meta.inputs.reduce((a, d) => a + d.size, 0);
  }
  // same errors as witness_calculator.js: unknown / missing / short / long input signals
  _inputBytes(input) {
    const buf = new Uint8Array(this.nInputs * 32);
    const known = new Map(this.meta.inputs.map((d) => [d.name, d]));
    for (const k of Object.keys(input)) if (!known.has(k)) throw new Error(`Signal not found: ${k}`);
    let set = 0;
    for (const d of this.meta.inputs) {
      if (!(d.name in input)) continue;
      const vals = flatten(input[d.name], []);
      if (vals.length < d.size) throw new Error(`Not enough values for input signal ${d.name}`);
      if (vals.length > d.size) throw new Error(`Too many values for input signal ${d.name}`);
      vals.forEach((x, j) => {
        let v = ((x % P) + P) % P;
        for (let b = 0; b < 32; b++) { buf[(d.offset + j) * 32 + b] = Number(v & 0xffn); v >>= 8n; }
      });
      set += d.size;
    }
    if (set < this.nInputs) throw new Error(`Not all inputs have been set. Only ${set} out of ${this.nInputs}`);
    return buf;
  }
  async calculateWTNSBin(input) { return pzk.calculateWTNSBin(this.h, this._inputBytes(input)); }
  async calculateWitness(input) {
    const raw = pzk.calculateWitness(this.h, this._inputBytes(input));
    const w = new Array(raw.length / 32);
    for (let i = 0; i < w.length; i++) {
      let v = 0n;
      for (let b = 31; b >= 0; b--) v = (v << 8n) | BigInt(raw[i * 32 + b]);
      w[i] = v;
    }
    return w;
  }
  witnessBatchPacked(packed, batch, wantDigest = false) { return pzk.witnessBatchPacked(this.h, packed, batch, wantDigest); }
  close() { pzk.close(this.h); }
}

// circom_tester shape: const circuit = await wasm_tester(prefix); w = await circuit.calculateWitness(input, true);
// await circuit.checkConstraints(w)
async function wasm_tester(prefix, device = 0) {
  const calc = new WitnessCalculator(prefix + ".pzkp", device);
  return {
    calculateWitness: (input, _sanity) => calc.calculateWitness(input),
    checkConstraints: async (w) => {
      const n = w.length, body = new Uint8Array(n * 32);
      w.forEach((x, i) => { let v = BigInt(x); for (let b = 0; b < 32; b++) { body[i * 32 + b] = Number(v & 0xffn); v >>= 8n; } });
      const hdr = new Uint8Array(12 + 12 + 40 + 12), dv = new DataView(hdr.buffer);
      hdr.set([0x77, 0x74, 0x6e, 0x73]); dv.setUint32(4, 2, true); dv.setUint32(8, 2, true);
      dv.setUint32(12, 1, true); dv.setBigUint64(16, 40n, true); dv.setUint32(24, 32, true);
      let p = P; for (let b = 0; b < 32; b++) { hdr[28 + b] = Number(p & 0xffn); p >>= 8n; }
      dv.setUint32(60, n, true); dv.setUint32(64, 2, true); dv.setBigUint64(68, BigInt(n * 32), true);
      const wt = new Uint8Array(hdr.length + body.length); wt.set(hdr); wt.set(body, hdr.length);
      const r = pzk.wtnsCheck(prefix + ".r1cs", wt, device);
      if (!r.ok) throw new Error(`Constraint doesn't match (constraint ${r.firstBad})`);
      return true;
    },
    release: () => calc.close(),
  };
}

module.exports = { WitnessCalculator, wasm_tester };
