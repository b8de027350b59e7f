// Host build of simple_zk_rollups_b200/csrc/fp.cuh + ec.cuh for tests/test_fp_host.py (g++ only, no CUDA).  Under g++ the
// four carry-chain row primitives are portable C (the PTX blocks are the same adds in hardware, proven by the GPU tests);
// everything composed from them -- the CIOS, the multi-product CIOS with one reduction, Fq2 (Karatsuba and schoolbook-lazy),
// the XYZZ mixed addition in both forms -- is the very code the kernels run, checked here against Python integers.
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "../../simple_zk_rollups_b200/csrc/ec.cuh"

thread_local int zkr::zkr_host_carry_lost = 0;

using namespace zkr;

extern "C" int fp_host_carry_lost() {
    int v = zkr_host_carry_lost;
    zkr_host_carry_lost = 0;
    return v;
}

template <class F>
static void field_ops(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out, int n) {
    for (int k = 0; k < n; k++) {
        F A = F::load(a + 8 * k), B = F::load(b + 8 * k), Cc = F::load(c + 8 * k), D = F::load(d + 8 * k), r;
        switch (op) {
            case 0: r = A * B; break;
            case 1: r = A + B; break;
            case 2: r = A - B; break;
            case 3: r = F::mul2(A, B, Cc, D); break;          // a b + c d
            case 4: r = F::msub(A, B, Cc, D); break;          // a b - c d
            case 5: r = A.neg_lazy(); break;                  // p - a in (0, p]
            case 6: r = A.sqr(); break;
            case 8: r = A.sqr_fast(); break;                  // mont_sqr_raw
            default: {                                        // 7: four products a b + c d + a d + c b
                mont_mul4_raw<typename std::conditional<std::is_same<F, Fq>::value, FqParams, FrParams>::type>(
                    r.v, A.v, B.v, Cc.v, D.v, A.v, D.v, Cc.v, B.v);
            }
        }
        r.store(out + 8 * k);
    }
}

// field: 0 = Fq, 1 = Fr; operands / results n x 8 little-endian u32 limbs (Montgomery residues, any value <= p)
extern "C" void fp_host_op(int field, int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d,
                           uint32_t* out, int n) {
    if (field == 0) field_ops<Fq>(op, a, b, c, d, out, n);
    else field_ops<Fr>(op, a, b, c, d, out, n);
}

// Fq2 values are n x 16 limbs (c0 | c1).  op 0: Karatsuba a b, 1: lazy a b, 2: a b - c d (Karatsuba), 3: msub(a, b, c, d), 4: a^2
extern "C" void fq2_host_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out, int n) {
    for (int k = 0; k < n; k++) {
        Fq2 A = Fq2::load(a + 16 * k), B = Fq2::load(b + 16 * k), Cc = Fq2::load(c + 16 * k), D = Fq2::load(d + 16 * k), r;
        switch (op) {
            case 0: r = A * B; break;
            case 1: r = Fq2::mul_l(A, B); break;
            case 2: r = A * B - Cc * D; break;
            case 3: r = Fq2::msub(A, B, Cc, D); break;
            default: r = A.sqr();
        }
        r.store(out + 16 * k);
    }
}

template <class F>
static void madd_ops(int lazy, const uint32_t* acc, const uint32_t* pt, uint32_t* out, int n) {
    constexpr int W = sizeof(F) / 4;
    for (int k = 0; k < n; k++) {
        XYZZ<F> A = XYZZ<F>::load(acc + 4 * W * k);
        Affine<F> P = Affine<F>::load(pt + 2 * W * k);
        if (lazy == 2) A.template madd_lazy<true>(P);
        else if (lazy) A.template madd_lazy<false>(P);
        else A.madd(P);
        A.store(out + 4 * W * k);
    }
}

// group: 1 = G1 (Fq), 2 = G2 (Fq2); acc = n XYZZ accumulators, pt = n affine points, out = acc + pt
extern "C" void xyzz_host_madd(int group, int lazy, const uint32_t* acc, const uint32_t* pt, uint32_t* out, int n) {
    if (group == 1) madd_ops<Fq>(lazy, acc, pt, out, n);
    else madd_ops<Fq2>(lazy, acc, pt, out, n);
}

template <class F>
static void add_ops(const uint32_t* a, const uint32_t* b, uint32_t* out, int n) {
    constexpr int W = sizeof(F) / 4;
    for (int k = 0; k < n; k++) {
        XYZZ<F> A = XYZZ<F>::load(a + 4 * W * k), B = XYZZ<F>::load(b + 4 * W * k);
        A.add(B);
        A.store(out + 4 * W * k);
    }
}

// full XYZZ addition a + b (the form ZKR_LAZY_TAIL selects at compile time: the test builds this file both ways)
extern "C" void xyzz_host_add(int group, const uint32_t* a, const uint32_t* b, uint32_t* out, int n) {
    if (group == 1) add_ops<Fq>(a, b, out, n);
    else add_ops<Fq2>(a, b, out, n);
}
template <class F>
static void dbl_ops(const uint32_t* a, uint32_t* out, int n) {
    constexpr int W = sizeof(F) / 4;
    for (int k = 0; k < n; k++) {
        XYZZ<F> A = XYZZ<F>::load(a + 4 * W * k);
        A = A.dbl();
        A.store(out + 4 * W * k);
    }
}
extern "C" void xyzz_host_dbl(int group, const uint32_t* a, uint32_t* out, int n) {
    if (group == 1) dbl_ops<Fq>(a, out, n);
    else dbl_ops<Fq2>(a, out, n);
}
extern "C" int fp_host_lazy_tail() { return ZKR_LAZY_TAIL; }
