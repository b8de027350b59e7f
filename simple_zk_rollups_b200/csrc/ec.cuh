// Short-Weierstrass (a = 0) point arithmetic in XYZZ coordinates, templated on the coordinate
// field (Fq for G1, Fq2 for G2).  Replaces snarkjs GCurve / websnark g1m_*, g2m_* used by the
// five multiexps of groth16GenProof (/root/reference/operator/src/snarks/common.ts:29).
//
//   XYZZ:  x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;   identity <=> ZZ == 0  (all-zero words, so
//   cudaMemset(0) produces an array of identities).
//   mixed add 8M+2S, full add 12M+2S, doubling 6M+3S  (EFD madd-2008-s / add-2008-s / dbl-2008-s-1).
// Exceptional inputs (P == Q, P == -Q, identity operands) are handled, not assumed away.
#pragma once
#include "fp.cuh"

namespace zkr {

template <class F>
struct Affine {
    F x, y;
    // websnark loader rule (binarify.ts:92-95 drops z): x == 0 <=> point at infinity
    __device__ __forceinline__ bool is_inf() const { return x.is_zero(); }
    static __device__ __forceinline__ Affine load(const void* p) {
        return {F::load(p), F::load(reinterpret_cast<const char*>(p) + sizeof(F))};
    }
    static __device__ __forceinline__ Affine load_ro(const void* p) {
        return {F::load_ro(p), F::load_ro(reinterpret_cast<const char*>(p) + sizeof(F))};
    }
    __device__ __forceinline__ void store(void* p) const {
        x.store(p);
        y.store(reinterpret_cast<char*>(p) + sizeof(F));
    }
    __device__ __forceinline__ Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    static __device__ __forceinline__ XYZZ identity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    static __device__ __forceinline__ XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return identity();
        return {p.x, p.y, F::one(), F::one()};
    }
    __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }
    __device__ __forceinline__ XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

    static __device__ __forceinline__ XYZZ load(const void* p) {
        const char* c = reinterpret_cast<const char*>(p);
        return {F::load(c), F::load(c + sizeof(F)), F::load(c + 2 * sizeof(F)), F::load(c + 3 * sizeof(F))};
    }
    __device__ __forceinline__ void store(void* p) const {
        char* c = reinterpret_cast<char*>(p);
        x.store(c);
        y.store(c + sizeof(F));
        zz.store(c + 2 * sizeof(F));
        zzz.store(c + 3 * sizeof(F));
    }

    // 2 * (affine point)
    static __device__ __forceinline__ XYZZ dbl_affine(const Affine<F>& p) {
        if (p.y.is_zero()) return identity();
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        F y3 = m * (s - x3) - w * p.y;
        return {x3, y3, v, w};
    }

    __device__ __forceinline__ XYZZ dbl() const {
        if (is_inf() || y.is_zero()) return identity();
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        F x3 = m.sqr() - s.dbl();
        F y3 = m * (s - x3) - w * y;
        return {x3, y3, v * zz, w * zzz};
    }

    // this += affine q  (q must not be infinity)
    __device__ __forceinline__ void madd(const Affine<F>& q) {
        if (is_inf()) {
            x = q.x; y = q.y; zz = F::one(); zzz = F::one();
            return;
        }
        F p = q.x * zz - x;
        F r = q.y * zzz - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(q);
            else *this = identity();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F qq = x * pp;
        F x3 = r.sqr() - ppp - qq.dbl();
        y = r * (qq - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o
    __device__ __forceinline__ void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        F u1 = x * o.zz, u2 = o.x * zz;
        F s1 = y * o.zzz, s2 = o.y * zzz;
        F p = u2 - u1;
        F r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F qq = u1 * pp;
        F x3 = r.sqr() - ppp - qq.dbl();
        y = r * (qq - x3) - s1 * ppp;
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }

    // affine, Montgomery form; identity -> (0, 0)
    __device__ __forceinline__ Affine<F> to_affine() const {
        if (is_inf()) return {F::zero(), F::zero()};
        // 1/zzz gives both: 1/zz = (1/zzz)^2 * zz^2 ... cheaper: one inversion of zzz, then
        // 1/zz = zzz^-2 * zz^2  (since zz^3 = zzz^2  =>  zz^-1 = zz^2 * zzz^-2)
        F iz3 = zzz.inverse();
        F iz2 = iz3.sqr() * zz.sqr();
        return {x * iz2, y * iz3};
    }
    // same, with the variable-time inversion of fp_inv.cuh: single-thread tails only (k_finish)
    __device__ __forceinline__ Affine<F> to_affine_vartime() const {
        if (is_inf()) return {F::zero(), F::zero()};
        F iz3 = zzz.inverse_vartime();
        F iz2 = iz3.sqr() * zz.sqr();
        return {x * iz2, y * iz3};
    }
};

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;

// Out-of-line helpers for cold paths (block-level bucket reduction).  They are strictly
// MEMORY-TO-MEMORY: every pointer must address shared or global memory and all big values live
// inside the callee.  Passing 256-byte structs across a device call boundary (by value, or as
// pointers to the caller's locals) produced wrong results with nvcc/ptxas 12.9 on sm_100a -- see
// DESIGN.md "toolchain notes" and tools/scratch/variants.cu for the reproducer.
template <class F>
__device__ __noinline__ void xyzz_add_mem(XYZZ<F>* dst, const XYZZ<F>* a, const XYZZ<F>* b) {
    XYZZ<F> x = *a;
    x.add(*b);
    *dst = x;
}
template <class F>
__device__ __noinline__ void xyzz_dbl_mem(XYZZ<F>* dst) {
    XYZZ<F> x = *dst;
    *dst = x.dbl();
}

// k * p, k a 256-bit standard-form integer (LSB-first double-and-add); O(1) uses per proof only.
template <class F>
__device__ __forceinline__ XYZZ<F> scalar_mul(XYZZ<F> p, Fr k) {
    XYZZ<F> acc = XYZZ<F>::identity(), base = p;
    int top = 255;
    while (top >= 0 && !((k.v[top >> 5] >> (top & 31)) & 1)) top--;
    for (int i = 0; i <= top; i++) {
        if ((k.v[i >> 5] >> (i & 31)) & 1) acc.add(base);
        if (i < top) base = base.dbl();
    }
    return acc;
}

}  // namespace zkr
