pragma circom 2.1.6;
// Test circuit written for this repository (not derived from the reference): sums whose product term is not a
// signal, in every sign combination, on field elements and on 64-bit integers - the shapes the compiler turns into
// fused multiply-add records (pzk_program.h: PZK_F_MULADD, PZK_Z_MULADD; PZK_FLAG_DST2 when the product is a signal).

template MulAdd() {
    signal input x;      // field elements
    signal input y;
    signal input z;
    signal input a;      // 64-bit integers
    signal input b;
    signal input c;
    signal output fa;
    signal output fb;
    signal output fc;
    signal output fd;
    signal output za;
    signal output zb;
    signal output zc;
    signal output zd;

    fa <-- x * y + z;
    fb <-- z - x * y;
    fc <-- x * y - z;
    fd <-- z + y * 12345678901234567890123;   // product with a pool constant
    fa + fb === 2 * z;
    fa - fc === 2 * z;

    za <-- a * b + c;
    zb <-- c - a * b;
    zc <-- a * b - c;
    zd <-- (a * b + c) * b + a;               // a chain: the inner sum feeds a second product
    za + zb === 2 * c;
    za - zc === 2 * c;

    // products that ARE signals and have the sum as their only reader: second results of the fused record
    signal output fp;
    signal output fs;
    signal output zp;
    signal output zs;
    signal output zt;
    fp <== y * z;
    fs <== fp + x;
    zp <== b * c;
    zs <== zp + a;
    signal zq <== a * c;
    zt <== b - zq;
}

component main = MulAdd();
