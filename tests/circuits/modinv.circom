pragma circom 2.1.6;
// Test circuit written for this repository: a witness-time function named mod_inv with a constant
// prime modulus (here the Mersenne prime 2^89 - 1 in two 64-bit limbs) - the compiler replaces it by
// the MODINV intrinsic; the Python oracle interprets the square-and-multiply below.

function mod_inv(n, k, a, p) {
    var A = 0;
    var P = 0;
    for (var i = k - 1; i >= 0; i--) { A = A * (1 << n) + a[i]; P = P * (1 << n) + p[i]; }
    A = A % P;
    var e = P - 2;
    var r = 1;
    for (var i = 88; i >= 0; i--) {
        r = (r * r) % P;
        if ((e >> i) & 1 == 1) { r = (r * A) % P; }
    }
    if (A == 0) { r = 0; }
    var out[200];
    for (var i = 0; i < k; i++) { out[i] = r % (1 << n); r = r \ (1 << n); }
    return out;
}

template Inv() {
    signal input a[2];          // 64-bit limbs, any value below 2^128 (reduced mod p by the function)
    signal output inv[2];
    signal output rem;
    var P[2] = [18446744073709551615, 33554431];   // 2^89 - 1
    var r[200] = mod_inv(64, 2, a, P);
    inv[0] <-- r[0];
    inv[1] <-- r[1];
    // a * inv = q * p + rem over the integers (all below 2^218 < the field modulus)
    var PP = P[0] + P[1] * (1 << 64);
    signal A <== a[0] + a[1] * (1 << 64);
    signal I <== inv[0] + inv[1] * (1 << 64);
    signal prod <== A * I;
    signal q <-- prod \ PP;
    rem <-- prod % PP;
    q * PP + rem === prod;
}

component main = Inv();
