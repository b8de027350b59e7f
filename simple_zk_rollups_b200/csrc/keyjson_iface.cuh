// Internal interface of keyjson.cu used by verify.cu: verification_key.json -> Montgomery-form bytes.
#pragma once
#include <cstdint>
#include <vector>

namespace zkr {

struct VKeyRaw {                     // all coordinates Fq-M, little-endian (same encoding as the binary proving key)
    uint32_t n_public = 0;
    uint8_t alfa1[64] = {};
    uint8_t beta2[128] = {}, gamma2[128] = {}, delta2[128] = {};
    std::vector<uint8_t> ic;         // (n_public + 1) x 64 B
};

// snarkjs verification_key.json text (protocol, nPublic, IC, vk_alfa_1, vk_beta_2, vk_gamma_2,
// vk_delta_2; vk_alfabeta_12 and unknown fields are skipped).  Returns ZKR_OK or ZKR_E_BADKEY (message set).
int vkey_parse_json(const char* json, size_t len, VKeyRaw* out);

}  // namespace zkr
