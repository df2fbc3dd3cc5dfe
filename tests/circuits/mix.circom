pragma circom 2.1.6;
// Test circuit written for this repository (not derived from the reference): exercises field
// arithmetic, inversion, bit decomposition, comparisons, integer operators and selection.

template Bits(n) {
    signal input in;
    signal output out[n];
    var acc = 0;
    for (var i = 0; i < n; i++) {
        out[i] <-- (in >> i) & 1;
        out[i] * (out[i] - 1) === 0;
        acc += out[i] * (1 << i);
    }
    acc === in;
}

template NonZeroInv() {
    signal input in;
    signal output isz;
    signal inv;
    inv <-- in != 0 ? 1 / in : 0;
    isz <== 1 - in * inv;
    in * isz === 0;
}

template Less(n) {
    signal input a;
    signal input b;
    signal output lt;
    component d = Bits(n + 1);
    d.in <== a + (1 << n) - b;
    lt <== 1 - d.out[n];
}

template Mix() {
    signal input x;          // field element
    signal input y;          // field element
    signal input u[4];       // 16-bit values
    signal input bits[8];    // bits
    signal output prod;
    signal output q;
    signal output r;
    signal output sel;
    signal output parity;
    signal output cube;

    prod <== x * y + 7;
    signal x2 <== x * x;
    cube <== x2 * x - y;

    component nz = NonZeroInv();
    nz.in <== x - y;

    // integer quotient / remainder of small values, constrained
    var s = u[0] + u[1] * 3;
    q <-- s \ (u[2] + 1);
    r <-- s % (u[2] + 1);
    q * (u[2] + 1) + r === s;
    component lt = Less(20);
    lt.a <== r;
    lt.b <== u[2] + 1;
    lt.lt === 1;

    // selection on a signal-dependent condition
    signal pick;
    pick <-- u[3] > u[0] ? u[3] : u[0];
    component l2 = Less(17);
    l2.a <== u[0];
    l2.b <== u[3];
    pick === l2.lt * (u[3] - u[0]) + u[0];
    sel <== pick + nz.isz;

    // xor chain on bits
    signal acc[8];
    acc[0] <== bits[0] * bits[0];
    for (var i = 1; i < 8; i++) {
        acc[i] <== acc[i - 1] + bits[i] - 2 * acc[i - 1] * bits[i];
    }
    parity <== acc[7];

    component xb = Bits(254);
    xb.in <== x;
}

component main {public [y]} = Mix();
