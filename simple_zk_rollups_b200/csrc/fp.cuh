// BN254 Fq / Fr Montgomery arithmetic on 8 x 32-bit limbs for sm_100a.
//
// Replaces the snarkjs bigInt / ZqField and websnark WASM f1m_* inner loops that
// /root/reference/operator/src/snarks/common.ts:29 (groth16GenProof) runs on the CPU.
// Representation matches the wire format of /root/reference/operator/src/utils/binarify.ts:68-90:
// little-endian limbs, Montgomery radix 2^256, values fully reduced to [0, p).
//
// Multiplication is an 8x8 CIOS split into an "even" and an "odd" accumulator so that every
// 32x32->64 product lands 64-bit aligned and each row is one uninterrupted carry chain of
// mad.lo.cc / madc.hi.cc pairs; ptxas fuses each pair into one IMAD.WIDE.U32(.X) with carry
// predicate, i.e. 128 IMAD.WIDE + 8 IMAD per modmul (the "136 IMAD" of SURVEY.md 8(d)).
#pragma once
#include <cstdint>

#include "fp_inv.cuh"

namespace zkr {

struct FqParams {   // base field q (binarify.ts:80)
    static __host__ __device__ __forceinline__ constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;   // -q^-1 mod 2^32
    // R mod q, R^2 mod q (R = 2^256)
    static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
};

struct FrParams {   // scalar field r (binarify.ts:87)
    static __host__ __device__ __forceinline__ constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;   // -r^-1 mod 2^32
    static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
};

// ------------------------------------------------------------------ carry-chain building blocks
// NOTE on constraints: every asm block below is several instructions long and reads inputs after
// it has started writing outputs, so write-only outputs MUST be early-clobber ("=&r"); with plain
// "=r" the compiler may give an output the register of a not-yet-consumed input (seen in practice:
// out-of-line instantiations produced garbage while inlined ones happened to work).
#ifndef __CUDACC__
// Host build (g++, tests/csrc/fp_host.cpp): the same row primitives in portable C so that everything composed from them
// -- the CIOS, the multi-product CIOS, Fq2, the curve formulas -- is checked against Python integers without a GPU.
// out = in + {x0,x2,x4,x6} * k + cin over 256 bits, returns the carry out (asserting callers pass 0 where they claim it).
inline uint32_t host_mad4(uint32_t* out, const uint32_t* in, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t k,
                          uint32_t cin) {
    const uint32_t x[4] = {x0, x2, x4, x6};
    uint64_t carry = cin;
    for (int j = 0; j < 4; j++) {
        const uint64_t prod = (uint64_t)x[j] * k;
        uint64_t lo = (uint64_t)in[2 * j] + (uint32_t)prod + carry;
        out[2 * j] = (uint32_t)lo;
        uint64_t hi = (uint64_t)in[2 * j + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        out[2 * j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    return (uint32_t)carry;
}
extern thread_local int zkr_host_carry_lost;   // set when a "carry out known to be zero" chain produced a carry (tests read it)
#endif
// acc[0..7] += {x0,x2,x4,x6} * k laid out 64-bit aligned; the carry out is added into *top.
__device__ __forceinline__ void mad_row_carry(uint32_t* acc, uint32_t& top, uint32_t x0, uint32_t x2,
                                              uint32_t x4, uint32_t x6, uint32_t k) {
#ifndef __CUDACC__
    top += host_mad4(acc, acc, x0, x2, x4, x6, k, 0);
#else
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(k));
#endif
}

// same, carry out known to be zero (the odd accumulator never exceeds 256 bits)
__device__ __forceinline__ void mad_row(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4,
                                        uint32_t x6, uint32_t k) {
#ifndef __CUDACC__
    if (host_mad4(acc, acc, x0, x2, x4, x6, k, 0)) zkr_host_carry_lost = 1;   // "known to be zero": the host tests assert it
#else
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
          "+r"(acc[6]), "+r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(k));
#endif
}

// e0 += o[1] (carry into the chain); o <- (o >> 64) + {x1,x3,x5,x7} * k
__device__ __forceinline__ void mad_row_shift(uint32_t& e0, uint32_t* o, uint32_t x1, uint32_t x3,
                                              uint32_t x5, uint32_t x7, uint32_t k) {
#ifndef __CUDACC__
    const uint64_t s0 = (uint64_t)e0 + o[1];
    e0 = (uint32_t)s0;
    const uint32_t sh[8] = {o[2], o[3], o[4], o[5], o[6], o[7], 0, 0};
    if (host_mad4(o, sh, x1, x3, x5, x7, k, (uint32_t)(s0 >> 32))) zkr_host_carry_lost = 1;
#else
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]),
          "+r"(o[6]), "+r"(o[7])
        : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(k));
#endif
}

__device__ __forceinline__ void mul_row(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4,
                                        uint32_t x6, uint32_t k) {
#ifndef __CUDACC__
    const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    host_mad4(acc, z, x0, x2, x4, x6, k, 0);
#else
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=&r"(acc[0]), "=&r"(acc[1]), "=&r"(acc[2]), "=&r"(acc[3]), "=&r"(acc[4]), "=&r"(acc[5]),
          "=&r"(acc[6]), "=&r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(k));
#endif
}

template <class P>
__device__ __forceinline__ void mont_row(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi, bool first) {
    if (first) {
        mul_row(o, a[1], a[3], a[5], a[7], bi);
        mul_row(e, a[0], a[2], a[4], a[6], bi);
    } else {
        mad_row_shift(e[0], o, a[1], a[3], a[5], a[7], bi);
        mad_row_carry(e, o[7], a[0], a[2], a[4], a[6], bi);
    }
    uint32_t mi = e[0] * P::INV;
    mad_row(o, P::mod(1), P::mod(3), P::mod(5), P::mod(7), mi);
    mad_row_carry(e, o[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), mi);
}

// r = r - p if r >= p
template <class P>
__device__ __forceinline__ void final_sub(uint32_t* r) {
    uint32_t t[8], borrow;
#ifndef __CUDACC__
    uint32_t pm[8];
    for (int i = 0; i < 8; i++) pm[i] = P::mod(i);
    borrow = 0u - u256_sub(t, r, pm);
#else
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]),
          "=&r"(t[7]), "=&r"(borrow)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)),
          "r"(P::mod(5)), "r"(P::mod(6)), "r"(P::mod(7)));
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

// even += odd >> 32: folds the two accumulators after the last row (the value is < 2p < 2^256: no carry out)
__device__ __forceinline__ void fold_even_odd(uint32_t* even, const uint32_t* odd) {
#ifndef __CUDACC__
    const uint32_t sh[8] = {odd[1], odd[2], odd[3], odd[4], odd[5], odd[6], odd[7], 0};
    if (u256_add(even, even, sh)) zkr_host_carry_lost = 1;
#else
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(even[0]), "+r"(even[1]), "+r"(even[2]), "+r"(even[3]), "+r"(even[4]), "+r"(even[5]),
          "+r"(even[6]), "+r"(even[7])
        : "r"(odd[1]), "r"(odd[2]), "r"(odd[3]), "r"(odd[4]), "r"(odd[5]), "r"(odd[6]), "r"(odd[7]));
#endif
}

template <class P>
__device__ __forceinline__ void mont_mul_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t even[8], odd[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mont_row<P>(even, odd, a, b[i], i == 0);
        mont_row<P>(odd, even, a, b[i + 1], false);
    }
    fold_even_odd(even, odd);
    final_sub<P>(even);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = even[i];
}

// ------------------------------------------------------------------ sums of products with ONE reduction
// (a0 b0 + a1 b1 [+ a2 b2 + a3 b3]) / R mod p in a single CIOS pass ("lazy reduction" across products): every row adds the
// N partial products a_k * b_k[i] into the same even / odd accumulators and reduces once, so N products cost
// 8 (8 N + 8) + 8 multiplier instructions instead of N * 136: 200 instead of 272 for two, 328 instead of 544 for four.
// Bounds (q < 0.1891 * 2^256, r smaller still; operands <= p): the running value t obeys t' < t / 2^32 + (N + 1) p, so before
// a row's shift the accumulators hold less than (N + 1) p (2^32 + 1) < 2^288 for N <= 4 -- the odd accumulator never
// carries out of its 256 bits, exactly as in the single product -- and the result (sum + m p) / R < p (1 + N p / R) < 2 p:
// one conditional subtraction finishes it.  tests/test_fp_host.py runs this code on the host against Python integers and
// asserts that no "known to be zero" carry was ever lost, on random and on all-maximal operands.
template <class P>
__device__ __forceinline__ void mont_row_first(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi, bool first) {
    if (first) {
        mul_row(o, a[1], a[3], a[5], a[7], bi);
        mul_row(e, a[0], a[2], a[4], a[6], bi);
    } else {
        mad_row_shift(e[0], o, a[1], a[3], a[5], a[7], bi);
        mad_row_carry(e, o[7], a[0], a[2], a[4], a[6], bi);
    }
}
__device__ __forceinline__ void mont_row_more(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
    mad_row(o, a[1], a[3], a[5], a[7], bi);
    mad_row_carry(e, o[7], a[0], a[2], a[4], a[6], bi);
}
template <class P>
__device__ __forceinline__ void mont_row_reduce(uint32_t* e, uint32_t* o) {
    uint32_t mi = e[0] * P::INV;
    mad_row(o, P::mod(1), P::mod(3), P::mod(5), P::mod(7), mi);
    mad_row_carry(e, o[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), mi);
}
template <class P>
__device__ __forceinline__ void mont_mul2_raw(uint32_t* r, const uint32_t* a0, const uint32_t* b0, const uint32_t* a1,
                                              const uint32_t* b1) {
    uint32_t even[8], odd[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mont_row_first<P>(even, odd, a0, b0[i], i == 0);
        mont_row_more(even, odd, a1, b1[i]);
        mont_row_reduce<P>(even, odd);
        mont_row_first<P>(odd, even, a0, b0[i + 1], false);
        mont_row_more(odd, even, a1, b1[i + 1]);
        mont_row_reduce<P>(odd, even);
    }
    fold_even_odd(even, odd);
    final_sub<P>(even);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = even[i];
}
template <class P>
__device__ __forceinline__ void mont_mul4_raw(uint32_t* r, const uint32_t* a0, const uint32_t* b0, const uint32_t* a1,
                                              const uint32_t* b1, const uint32_t* a2, const uint32_t* b2, const uint32_t* a3,
                                              const uint32_t* b3) {
    uint32_t even[8], odd[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mont_row_first<P>(even, odd, a0, b0[i], i == 0);
        mont_row_more(even, odd, a1, b1[i]);
        mont_row_more(even, odd, a2, b2[i]);
        mont_row_more(even, odd, a3, b3[i]);
        mont_row_reduce<P>(even, odd);
        mont_row_first<P>(odd, even, a0, b0[i + 1], false);
        mont_row_more(odd, even, a1, b1[i + 1]);
        mont_row_more(odd, even, a2, b2[i + 1]);
        mont_row_more(odd, even, a3, b3[i + 1]);
        mont_row_reduce<P>(odd, even);
    }
    fold_even_odd(even, odd);
    final_sub<P>(even);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = even[i];
}

// ------------------------------------------------------------------ squaring: 100 instead of 128 wide multiplies
// a^2 = sum_i a_i 2^(32 i) * B_i  with  B_i = a_i 2^(32 i) + 2 A_(>i)  (A_(>i) = the limbs of a above i): every cross product
// a_i a_j (i < j) is taken once, doubled through the operand.  With d = 2 a (a < p < 2^254, so d fits 8 limbs) the limbs of
// B_i are: 0 below i, a_i at i, d_(i+1) & ~1 at i + 1 (the bit a_i shifts into d_(i+1) belongs to 2 a_i, not to A_(>i)),
// d_q from i + 2 up.  Row i of the CIOS multiplies the scalar a_i by that vector instead of by a: 8 - i products instead
// of 8, the reduction stays interleaved (no 512-bit intermediate, no reduction-only pass).  A skipped product in the
// in-place group costs nothing; in the group whose MADs also move the window (mad_row_shift) it becomes two plain
// add-with-carry instructions -- 24 of them per squaring, on the ALU pipe, which the accumulation kernel leaves mostly idle.
// Bounds: the vector is < 2 p, so the running value obeys the two-product bound of the sums of products above.
// (Host build: the same rows in portable C; tests/test_fp_host.py checks the result against Python integers.)
template <int S>   // products s >= S of {x0,x2,x4,x6} * k into acc (in place), carry out into top
__device__ __forceinline__ void mad_row_carry_from(uint32_t* acc, uint32_t& top, uint32_t x0, uint32_t x2, uint32_t x4,
                                                   uint32_t x6, uint32_t k) {
    static_assert(S >= 0 && S <= 4, "S");
#ifndef __CUDACC__
    if (S < 4) top += host_mad4(acc, acc, S > 0 ? 0 : x0, S > 1 ? 0 : x2, S > 2 ? 0 : x4, x6, k, 0);
#else
    if (S == 0) {
        mad_row_carry(acc, top, x0, x2, x4, x6, k);
    } else if (S == 1) {
        asm("mad.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(k));
    } else if (S == 2) {
        asm("mad.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(k));
    } else if (S == 3) {
        asm("mad.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
              "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(k));
    }
#endif
}
template <int S>   // e0 += o[1]; o <- (o >> 64) + (products s >= S of {x1,x3,x5,x7} * k); the skipped ones only move the window
__device__ __forceinline__ void mad_row_shift_from(uint32_t& e0, uint32_t* o, uint32_t x1, uint32_t x3, uint32_t x5,
                                                   uint32_t x7, uint32_t k) {
    static_assert(S >= 0 && S <= 3, "S");
#ifndef __CUDACC__
    const uint64_t s0 = (uint64_t)e0 + o[1];
    e0 = (uint32_t)s0;
    const uint32_t sh[8] = {o[2], o[3], o[4], o[5], o[6], o[7], 0, 0};
    if (host_mad4(o, sh, S > 0 ? 0 : x1, S > 1 ? 0 : x3, S > 2 ? 0 : x5, x7, k, (uint32_t)(s0 >> 32))) zkr_host_carry_lost = 1;
#else
    if (S == 0) {
        mad_row_shift(e0, o, x1, x3, x5, x7, k);
    } else if (S == 1) {
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.cc.u32 %1, %3, 0;\n\t"
            "addc.cc.u32 %2, %4, 0;\n\t"
            "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
            "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
            "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
            "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
            "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
            "madc.hi.u32 %8, %12, %13, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(k));
    } else if (S == 2) {
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.cc.u32 %1, %3, 0;\n\t"
            "addc.cc.u32 %2, %4, 0;\n\t"
            "addc.cc.u32 %3, %5, 0;\n\t"
            "addc.cc.u32 %4, %6, 0;\n\t"
            "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
            "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
            "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
            "madc.hi.u32 %8, %12, %13, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(k));
    } else {
        asm("add.cc.u32 %0, %0, %2;\n\t"
            "addc.cc.u32 %1, %3, 0;\n\t"
            "addc.cc.u32 %2, %4, 0;\n\t"
            "addc.cc.u32 %3, %5, 0;\n\t"
            "addc.cc.u32 %4, %6, 0;\n\t"
            "addc.cc.u32 %5, %7, 0;\n\t"
            "addc.cc.u32 %6, %8, 0;\n\t"
            "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
            "madc.hi.u32 %8, %12, %13, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(k));
    }
#endif
}
// one squaring row: frame-relative operand limb f of row I is  0 (f < I), a_I (f == I), d_f & ~1 (f == I + 1), d_f (above)
template <int I, int F>
__device__ __forceinline__ uint32_t sqr_limb(const uint32_t* a, const uint32_t* d) {
    return F < I ? 0u : F == I ? a[I] : F == I + 1 ? (d[F] & ~1u) : d[F];
}
template <class P, int I>
__device__ __forceinline__ void mont_sqr_row(uint32_t* e, uint32_t* o, const uint32_t* a, const uint32_t* d) {
    if (I == 0) {
        mul_row(o, sqr_limb<I, 1>(a, d), sqr_limb<I, 3>(a, d), sqr_limb<I, 5>(a, d), sqr_limb<I, 7>(a, d), a[I]);
        mul_row(e, sqr_limb<I, 0>(a, d), sqr_limb<I, 2>(a, d), sqr_limb<I, 4>(a, d), sqr_limb<I, 6>(a, d), a[I]);
    } else {
        mad_row_shift_from<I / 2>(e[0], o, sqr_limb<I, 1>(a, d), sqr_limb<I, 3>(a, d), sqr_limb<I, 5>(a, d),
                                  sqr_limb<I, 7>(a, d), a[I]);
        mad_row_carry_from<(I + 1) / 2>(e, o[7], sqr_limb<I, 0>(a, d), sqr_limb<I, 2>(a, d), sqr_limb<I, 4>(a, d),
                                        sqr_limb<I, 6>(a, d), a[I]);
    }
    mont_row_reduce<P>(e, o);
}
template <class P>
__device__ __forceinline__ void mont_sqr_raw(uint32_t* r, const uint32_t* a) {
    uint32_t d[8];   // 2 a as a plain 256-bit integer
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = (a[i] << 1) | (i ? a[i - 1] >> 31 : 0u);
    uint32_t even[8], odd[8];
    mont_sqr_row<P, 0>(even, odd, a, d);
    mont_sqr_row<P, 1>(odd, even, a, d);
    mont_sqr_row<P, 2>(even, odd, a, d);
    mont_sqr_row<P, 3>(odd, even, a, d);
    mont_sqr_row<P, 4>(even, odd, a, d);
    mont_sqr_row<P, 5>(odd, even, a, d);
    mont_sqr_row<P, 6>(even, odd, a, d);
    mont_sqr_row<P, 7>(odd, even, a, d);
    fold_even_odd(even, odd);
    final_sub<P>(even);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = even[i];
}

// ------------------------------------------------------------------ field element
template <class P>
struct Fp {
    using Params = P;
    uint32_t v[8];

    static __device__ __forceinline__ Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    static __device__ __forceinline__ Fp one() {   // Montgomery 1
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r;
    }
    static __device__ __forceinline__ Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
        return r;
    }
    __device__ __forceinline__ bool is_zero() const {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) x |= v[i];
        return x == 0;
    }
    __device__ __forceinline__ bool operator==(const Fp& o) const {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) x |= v[i] ^ o.v[i];
        return x == 0;
    }
    __device__ __forceinline__ bool operator!=(const Fp& o) const { return !(*this == o); }

    friend __device__ __forceinline__ Fp operator*(const Fp& a, const Fp& b) {
        Fp r;
        mont_mul_raw<P>(r.v, a.v, b.v);
        return r;
    }
    __device__ __forceinline__ Fp sqr() const { return *this * *this; }
    __device__ __forceinline__ Fp sqr_fast() const {   // mont_sqr_raw: 100 instead of 128 wide multiplies
        Fp r;
        mont_sqr_raw<P>(r.v, v);
        return r;
    }

    friend __device__ __forceinline__ Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
#ifndef __CUDACC__
        u256_add(r.v, a.v, b.v);            // a, b < p < 2^254: no carry out
#else
        asm("add.cc.u32 %0, %8, %16;\n\t"
            "addc.cc.u32 %1, %9, %17;\n\t"
            "addc.cc.u32 %2, %10, %18;\n\t"
            "addc.cc.u32 %3, %11, %19;\n\t"
            "addc.cc.u32 %4, %12, %20;\n\t"
            "addc.cc.u32 %5, %13, %21;\n\t"
            "addc.cc.u32 %6, %14, %22;\n\t"
            "addc.u32 %7, %15, %23;"
            : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
              "=&r"(r.v[6]), "=&r"(r.v[7])
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
              "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
              "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#endif
        final_sub<P>(r.v);
        return r;
    }
    friend __device__ __forceinline__ Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t borrow;
#ifndef __CUDACC__
        borrow = 0u - u256_sub(r.v, a.v, b.v);
        uint32_t pm[8];
        for (int i = 0; i < 8; i++) pm[i] = P::mod(i) & borrow;
        u256_add(r.v, r.v, pm);
#else
        asm("sub.cc.u32 %0, %9, %17;\n\t"
            "subc.cc.u32 %1, %10, %18;\n\t"
            "subc.cc.u32 %2, %11, %19;\n\t"
            "subc.cc.u32 %3, %12, %20;\n\t"
            "subc.cc.u32 %4, %13, %21;\n\t"
            "subc.cc.u32 %5, %14, %22;\n\t"
            "subc.cc.u32 %6, %15, %23;\n\t"
            "subc.cc.u32 %7, %16, %24;\n\t"
            "subc.u32 %8, 0, 0;"
            : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
              "=&r"(r.v[6]), "=&r"(r.v[7]), "=&r"(borrow)
            : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
              "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
              "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
        // borrow is 0 or 0xffffffff: add back p & borrow
        asm("add.cc.u32 %0, %0, %8;\n\t"
            "addc.cc.u32 %1, %1, %9;\n\t"
            "addc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\t"
            "addc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\t"
            "addc.u32 %7, %7, %15;"
            : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]),
              "+r"(r.v[6]), "+r"(r.v[7])
            : "r"(P::mod(0) & borrow), "r"(P::mod(1) & borrow), "r"(P::mod(2) & borrow),
              "r"(P::mod(3) & borrow), "r"(P::mod(4) & borrow), "r"(P::mod(5) & borrow),
              "r"(P::mod(6) & borrow), "r"(P::mod(7) & borrow));
#endif
        return r;
    }
    __device__ __forceinline__ Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    // p - a without the zero test: a value in (0, p], i.e. NOT canonical for a == 0 -- only ever an operand of the
    // multi-product CIOS below (which accepts operands <= p and returns canonical results)
    __device__ __forceinline__ Fp neg_lazy() const {
        Fp r;
#ifndef __CUDACC__
        uint32_t pm[8];
        for (int i = 0; i < 8; i++) pm[i] = P::mod(i);
        u256_sub(r.v, pm, v);
#else
        asm("sub.cc.u32 %0, %8, %16;\n\t"
            "subc.cc.u32 %1, %9, %17;\n\t"
            "subc.cc.u32 %2, %10, %18;\n\t"
            "subc.cc.u32 %3, %11, %19;\n\t"
            "subc.cc.u32 %4, %12, %20;\n\t"
            "subc.cc.u32 %5, %13, %21;\n\t"
            "subc.cc.u32 %6, %14, %22;\n\t"
            "subc.u32 %7, %15, %23;"
            : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]),
              "=&r"(r.v[6]), "=&r"(r.v[7])
            : "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)),
              "r"(P::mod(6)), "r"(P::mod(7)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
              "r"(v[6]), "r"(v[7]));
#endif
        return r;
    }
    // a b + c d and a b - c d with one reduction (mont_mul2_raw)
    static __device__ __forceinline__ Fp mul2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        Fp r;
        mont_mul2_raw<P>(r.v, a.v, b.v, c.v, d.v);
        return r;
    }
    static __device__ __forceinline__ Fp msub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        return mul2(a, b, c.neg_lazy(), d);
    }
    static __device__ __forceinline__ Fp mul_l(const Fp& a, const Fp& b) { return a * b; }   // Fq2 has a lazy product
    __device__ __forceinline__ Fp dbl() const { return *this + *this; }

    // standard form <-> Montgomery
    __device__ __forceinline__ Fp to_mont() const { return *this * r2(); }
    __device__ __forceinline__ Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }
    // a^(p-2); ~380 modmuls, used only O(1) times per proof / per setup element
    __device__ __forceinline__ Fp inverse() const {
        Fp base = *this, acc = one();
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = P::mod(i);
        e[0] -= 2;   // no borrow for either modulus (low limb >= 2)
        for (int i = 0; i < 254; i++) {
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * base;
            base = base.sqr();
        }
        return acc;
    }
    // Same result as inverse(), by the binary extended Euclid of fp_inv.cuh (variable time, no multiplier): for the
    // O(1) inversions that run on ONE thread at the very end of a proof (k_finish).  Not for full grids: the data-
    // dependent loops would diverge.  binary_inverse works on plain integers: (aR)^-1 = a^-1 R^-1, and two
    // Montgomery products with R^2 bring it back to a^-1 R.
    __device__ __forceinline__ Fp inverse_vartime() const {
        Fp t;
        binary_inverse<P>(t.v, v);
        return (t * r2()) * r2();
    }
    // is the standard-form value < p ?
    __device__ __forceinline__ bool in_range() const {
        uint32_t t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = v[i];
        final_sub<P>(t);
        bool same = true;
#pragma unroll
        for (int i = 0; i < 8; i++) same &= (t[i] == v[i]);
        return same;   // unchanged by the conditional subtract  <=>  v < p
    }

#ifdef __CUDACC__
    // 2 x 128-bit global memory access (callers guarantee 16-byte alignment)
    static __device__ __forceinline__ Fp load(const void* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = q[0], b = q[1];
        Fp r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    static __device__ __forceinline__ Fp load_ro(const void* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = __ldg(q), b = __ldg(q + 1);
        Fp r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(void* p) const {
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = make_uint4(v[0], v[1], v[2], v[3]);
        q[1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
#else
    static Fp load(const void* p) {
        Fp r;
        for (int i = 0; i < 8; i++) r.v[i] = reinterpret_cast<const uint32_t*>(p)[i];
        return r;
    }
    static Fp load_ro(const void* p) { return load(p); }
    void store(void* p) const {
        for (int i = 0; i < 8; i++) reinterpret_cast<uint32_t*>(p)[i] = v[i];
    }
#endif
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

// ------------------------------------------------------------------ Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
    Fq c0, c1;
    static __device__ __forceinline__ Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static __device__ __forceinline__ Fq2 one() { return {Fq::one(), Fq::zero()}; }
    __device__ __forceinline__ bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    __device__ __forceinline__ bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    __device__ __forceinline__ bool operator!=(const Fq2& o) const { return !(*this == o); }
    friend __device__ __forceinline__ Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
    friend __device__ __forceinline__ Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
    // Karatsuba: 3 Fq modmuls
    friend __device__ __forceinline__ Fq2 operator*(const Fq2& a, const Fq2& b) {
        Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
        Fq t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
        return {t0 - t1, t2 - t0 - t1};
    }
    // Schoolbook with lazy reduction: c0 = a0 b0 + a1 (-b1), c1 = a0 b1 + a1 b0 as two two-product CIOS passes.  The same
    // 384 wide multiplies as Karatsuba's three products, but 2 reductions' worth of final subtractions instead of 3 and
    // none of Karatsuba's five additions / subtractions (each a carry chain + conditional correction).
    static __device__ __forceinline__ Fq2 mul_l(const Fq2& a, const Fq2& b) {
        return {Fq::mul2(a.c0, b.c0, a.c1, b.c1.neg_lazy()), Fq::mul2(a.c0, b.c1, a.c1, b.c0)};
    }
    // a b - c d: each component is one four-product CIOS pass
    static __device__ __forceinline__ Fq2 msub(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) {
        const Fq nb1 = b.c1.neg_lazy(), nc0 = c.c0.neg_lazy(), nc1 = c.c1.neg_lazy();
        Fq2 r;
        // c0 = a0 b0 - a1 b1 - c0 d0 + c1 d1 ;  c1 = a0 b1 + a1 b0 - c0 d1 - c1 d0
        mont_mul4_raw<FqParams>(r.c0.v, a.c0.v, b.c0.v, a.c1.v, nb1.v, nc0.v, d.c0.v, c.c1.v, d.c1.v);
        mont_mul4_raw<FqParams>(r.c1.v, a.c0.v, b.c1.v, a.c1.v, b.c0.v, nc0.v, d.c1.v, nc1.v, d.c0.v);
        return r;
    }
    __device__ __forceinline__ Fq2 sqr_fast() const { return sqr(); }   // no Fq squaring inside an Fq2 squaring
    __device__ __forceinline__ Fq2 sqr() const {   // 2 modmuls
        Fq t = c0 * c1;
        return {(c0 + c1) * (c0 - c1), t + t};
    }
    __device__ __forceinline__ Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    __device__ __forceinline__ Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    __device__ __forceinline__ Fq2 inverse() const {
        Fq n = (c0.sqr() + c1.sqr()).inverse();
        return {c0 * n, (c1 * n).neg()};
    }
    __device__ __forceinline__ Fq2 inverse_vartime() const {   // see Fp::inverse_vartime
        Fq n = (c0.sqr() + c1.sqr()).inverse_vartime();
        return {c0 * n, (c1 * n).neg()};
    }
    __device__ __forceinline__ Fq2 to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
    __device__ __forceinline__ Fq2 from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
    static __device__ __forceinline__ Fq2 load(const void* p) {
        return {Fq::load(p), Fq::load(reinterpret_cast<const char*>(p) + 32)};
    }
    static __device__ __forceinline__ Fq2 load_ro(const void* p) {
        return {Fq::load_ro(p), Fq::load_ro(reinterpret_cast<const char*>(p) + 32)};
    }
    __device__ __forceinline__ void store(void* p) const {
        c0.store(p);
        c1.store(reinterpret_cast<char*>(p) + 32);
    }
};

}  // namespace zkr
