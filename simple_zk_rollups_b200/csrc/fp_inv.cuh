// Variable-time modular inversion for the O(1) inversions that sit on a proof's critical path.
//
// The last kernel of every proof (k_finish, prover.cu) turns pi_a, pi_b, pi_c into affine form: one field
// inversion per point, each on a single thread with nothing left to overlap it.  Fermat's a^(p-2) is 254
// dependent squarings + ~127 multiplications (~70 k dependent integer instructions, ~0.2 ms for a lone warp);
// the binary extended Euclid below needs <= 2*254 halvings and subtractions of plain 256-bit integers
// (shifts and adds only, no multiplier).  It is data dependent, which is fine here: the inputs are proof
// points that are published anyway, and one thread per block runs it, so there is no warp divergence to pay.
// Bulk inversions (table precomputation, one per thread of a full grid) keep Fp::inverse(): there the
// divergent loops would serialise the warp.
//
// Replaces the affine normalisation inside websnark groth16GenProof (/root/reference/operator/src/snarks/
// common.ts:29: the proof comes back as affine decimal strings).  Plain C++ on purpose (no inline PTX) so the
// same code is compiled for the host and checked against big-integer arithmetic in tests/test_fp_inv.py.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace zkr {

// r = a + b, returns the carry out
__host__ __device__ __forceinline__ uint32_t u256_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}

// r = a - b, returns the borrow out (1 if a < b)
__host__ __device__ __forceinline__ uint32_t u256_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint64_t t = (uint64_t)a[i] - b[i] - bw;
        r[i] = (uint32_t)t;
        bw = (t >> 32) & 1;
    }
    return (uint32_t)bw;
}

// a = (top : a) >> 1
__host__ __device__ __forceinline__ void u256_shr1(uint32_t* a, uint32_t top) {
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] = (a[7] >> 1) | (top << 31);
}

__host__ __device__ __forceinline__ bool u256_is_one(const uint32_t* a) {
    uint32_t o = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) o |= a[i];
    return o == 0;
}

// x = x / 2 mod p  (p odd, x < p)
template <class P>
__host__ __device__ __forceinline__ void u256_half_mod(uint32_t* x) {
    uint32_t top = 0;
    if (x[0] & 1) {
        uint32_t p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = P::mod(i);
        top = u256_add(x, x, p);
    }
    u256_shr1(x, top);
}

// out = a^-1 mod p as plain integers, for 0 < a < p (p = P::mod, an odd prime).  a == 0 returns 0.
// Binary extended Euclid (Guide to Elliptic Curve Cryptography, Alg. 2.22): invariants
//   x1 * a == u (mod p),  x2 * a == v (mod p),  gcd(u, v) == 1.
// Force-inlined on the device: out-of-line device functions taking pointers to the caller's locals are exactly
// what DESIGN.md section 7 (toolchain notes) says to avoid with this nvcc / ptxas.
template <class P>
__host__ __device__ __forceinline__ void binary_inverse(uint32_t* out, const uint32_t* a) {
    uint32_t u[8], v[8], x1[8], x2[8], p[8];
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u[i] = a[i];
        v[i] = p[i] = P::mod(i);
        x1[i] = 0;
        x2[i] = 0;
        any |= a[i];
    }
    x1[0] = 1;
    if (!any) {
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = 0;
        return;
    }
    // every pass removes at least one bit from u or v: <= 2 * 256 passes (the bound only guards against a >= p)
    for (int guard = 0; guard < 1024 && !u256_is_one(u) && !u256_is_one(v); guard++) {
        while (!(u[0] & 1)) {
            u256_shr1(u, 0);
            u256_half_mod<P>(x1);
        }
        while (!(v[0] & 1)) {
            u256_shr1(v, 0);
            u256_half_mod<P>(x2);
        }
        uint32_t t[8];
        if (!u256_sub(t, u, v)) {          // u >= v
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u[i] = t[i];
                nz |= t[i];
            }
            if (!nz) break;                // u == v != 1: gcd(a, p) != 1, only for a >= p (invalid input); result 0
            if (u256_sub(x1, x1, x2)) u256_add(x1, x1, p);
        } else {
            u256_sub(v, v, u);
            if (u256_sub(x2, x2, x1)) u256_add(x2, x2, p);
        }
    }
    const bool from_u = u256_is_one(u), from_v = u256_is_one(v);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = from_u ? x1[i] : (from_v ? x2[i] : 0u);
}

}  // namespace zkr
