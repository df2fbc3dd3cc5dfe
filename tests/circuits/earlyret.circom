pragma circom 2.1.6;
// Test circuit written for this repository: witness-time functions that leave through `return`
// under signal-dependent conditions (nested ifs, a return inside a loop body, an early exit in
// front of more work), and a long_div whose dividend limb is only known to fit 64 bits at run time.

// three-way classification with the shape of a trial-quotient correction step
function correct(q, lo, hi) {
    var t = q * 3;
    if (t > hi) {
        t = t - lo;
        if (t > hi) {
            return q - 2;
        } else {
            return q - 1;
        }
    } else {
        return q;
    }
}

// index of the first non-zero entry (4 when there is none): a return inside a loop body
function first_set(v) {
    for (var i = 0; i < 4; i++) {
        if (v[i] != 0) {
            return i;
        }
    }
    return 4;
}

// early exit in front of the general path; the general path runs speculatively
function inv_or_zero(x) {
    if (x == 0) {
        var z = 0;
        return z;
    }
    var y = 1 / x;
    return y;
}

// arrays leave through both arms
function ordered(a, b) {
    if (a[0] + a[1] > b[0] + b[1]) {
        return [b[0], b[1], a[0], a[1]];
    } else {
        if (a[0] > b[0]) {
            return [b[0], a[1], a[0], b[1]];
        }
        return [a[0], a[1], b[0], b[1]];
    }
}

function long_div(n, k, m, a, b) {
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

template EarlyRet() {
    signal input v[4];      // 16-bit
    signal input x;         // field
    signal input a[3];      // 64-bit
    signal input b[2];      // 64-bit
    signal input c;         // bit
    signal output cls;
    signal output idx;
    signal output inv;
    signal output ord[4];
    signal output q[2];
    signal output r[2];

    cls <-- correct(v[0], v[1], v[2] + v[3]);
    idx <-- first_set(v);
    inv <-- inv_or_zero(x);
    // inv is 1/x or 0, and x*inv is a bit that is 1 exactly when x != 0
    signal nz <== x * inv;
    nz * (1 - nz) === 0;
    x * (1 - nz) === 0;
    var o[4] = ordered([v[0], v[1]], [v[2], v[3]]);
    for (var i = 0; i < 4; i++) { ord[i] <-- o[i]; }
    ord[0] + ord[1] + ord[2] + ord[3] === v[0] + v[1] + v[2] + v[3];

    // the top dividend limb a[2] + c may be 2^64: narrowed at run time, asserted to fit
    var aa[3];
    aa[0] = a[0]; aa[1] = a[1]; aa[2] = a[2] + c;
    var d[2][200] = long_div(64, 2, 1, aa, b);
    for (var i = 0; i < 2; i++) { q[i] <-- d[0][i]; r[i] <-- d[1][i]; }
    signal t0 <== q[0] * b[0];
    signal t1 <== q[0] * b[1];
    signal t2 <== q[1] * b[0];
    signal t3 <== q[1] * b[1];
    t0 + r[0] + (t1 + t2 + r[1]) * (1 << 64) + t3 * (1 << 128) === a[0] + a[1] * (1 << 64) + (a[2] + c) * (1 << 128);
}

component main = EarlyRet();
