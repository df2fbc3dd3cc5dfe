pragma circom 2.1.6;
// Test circuit written for this repository: a 4x64-bit by 2x64-bit long division done by a
// witness-time function named long_div, to exercise the compiler's intrinsic and the constraint
// q * b + r == a checked limb-wise with carries folded into one field equation.

function long_div(n, k, m, a, b) {
    // schoolbook base-2^n long division through single-limb trial quotients is what real
    // circuits carry; this test only needs *a* correct definition: repeated subtraction on
    // the packed value is exact for the small limb counts used here.
    var out[2][200];
    var A = 0;
    var B = 0;
    for (var i = k + m - 1; i >= 0; i--) { A = A * (1 << n) + a[i]; }
    for (var i = k - 1; i >= 0; i--) { B = B * (1 << n) + b[i]; }
    var Q = A \ B;
    var R = A % B;
    for (var i = 0; i <= m; i++) { out[0][i] = Q % (1 << n); Q = Q \ (1 << n); }
    for (var i = 0; i < k; i++) { out[1][i] = R % (1 << n); R = R \ (1 << n); }
    return out;
}

template Div() {
    signal input a[3];
    signal input b[2];
    signal output q[2];
    signal output r[2];
    var d[2][200] = long_div(64, 2, 1, a, b);
    for (var i = 0; i < 2; i++) { q[i] <-- d[0][i]; r[i] <-- d[1][i]; }
    // (q0 + q1 X)(b0 + b1 X) + r0 + r1 X == a0 + a1 X + a2 X^2 at X = 2^64 (192 bits < p)
    signal t0 <== q[0] * b[0];
    signal t1 <== q[0] * b[1];
    signal t2 <== q[1] * b[0];
    signal t3 <== q[1] * b[1];
    t0 + r[0] + (t1 + t2 + r[1]) * (1 << 64) + t3 * (1 << 128) === a[0] + a[1] * (1 << 64) + a[2] * (1 << 128);
}

component main = Div();
