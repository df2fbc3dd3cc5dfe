pragma circom 2.1.6;
// Shim written for this repository.  The reference's query circuits use circomlib's `BabyPbk`
// (private key -> public key on BabyJubjub, Ax/Ay = sk * Base8) without vendoring circomlib
// (identityStateVerifier.circom:19; SURVEY.md section 8c).  The outputs are pinned by the maths;
// the internal signals of circomlib's template are not reproducible offline, so the name is
// bound to the reference's own in-tree multiplication (babyjubjub/curve.circom:143).
template BabyPbk() {
    signal input in;
    signal output Ax;
    signal output Ay;
    component mul = BabyjubjubBase8Multiplication();
    mul.scalar <== in;
    Ax <== mul.out[0];
    Ay <== mul.out[1];
}
