// Test-only host build of the device headers (carry flag emulated), so the limb-level algorithms
// can be checked against the oracle on a box without a GPU.  Never part of the shipped library.
#include <cstring>
#include "../../accumulation_b200/csrc/fp.cuh"
using namespace accmsm;

template <int F> static void binop(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe_t x, y, r;
        memcpy(x.l, a + 8 * i, 32);
        if (b) memcpy(y.l, b + 8 * i, 32);
        switch (op) {
            case 0: r = Fp<F>::mul(x, y); break;
            case 1: r = Fp<F>::add(x, y); break;
            case 2: r = Fp<F>::sub(x, y); break;
            case 3: r = Fp<F>::sqr(x); break;
            case 4: r = Fp<F>::inv(x); break;
            case 5: r = Fp<F>::from_mont(x); break;
            case 6: r = Fp<F>::to_mont(x); break;
            case 7: r = Fp<F>::neg(x); break;
            default: r = x;
        }
        memcpy(o + 8 * i, r.l, 32);
    }
}
extern "C" void host_fe_op(int field, int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) {
    if (field == 0) binop<0>(op, a, b, o, n); else binop<1>(op, a, b, o, n);
}
