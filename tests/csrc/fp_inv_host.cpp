// Host build of simple_zk_rollups_b200/csrc/fp_inv.cuh for tests/test_fp_inv.py (g++ only, no CUDA):
// the inversion k_finish runs on the GPU is plain C++, so the very same code is checked here against Python ints.
#include <cstdint>
#include <cstring>

#include "../../simple_zk_rollups_b200/csrc/fp_inv.cuh"

namespace {
struct QParams {   // same limbs as FqParams::mod in fp.cuh (asserted against oracle.bn254.Q by the test)
    static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
};
struct RParams {
    static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
};
}  // namespace

extern "C" void fp_inv_modulus(int field, uint32_t* out) {
    for (int i = 0; i < 8; i++) out[i] = field == 0 ? QParams::mod(i) : RParams::mod(i);
}

// n inversions: in / out are n x 8 little-endian u32 limbs
extern "C" void fp_inv_batch(int field, const uint32_t* in, uint32_t* out, int n) {
    for (int k = 0; k < n; k++) {
        if (field == 0) zkr::binary_inverse<QParams>(out + 8 * k, in + 8 * k);
        else zkr::binary_inverse<RParams>(out + 8 * k, in + 8 * k);
    }
}
