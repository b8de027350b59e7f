// Shared view of the NTT engine (ntt.cu) for the prover and the synthetic setup.
#pragma once
#include "common.cuh"
#include "fp.cuh"

namespace zkr {

struct NttRoots {
    Fr w, wi;        // omega_N, omega_N^-1
    Fr g, gi;        // omega_2N (coset shift), inverse
    Fr ninv;         // 1/N
    Fr hconst;       // R/(2N): mont_mul(x/R, .) = x/(2N)   (H pipeline final scaling)
    Fr one;
};

struct NttTables {
    int log_n = 0, kw = 0, lb = 0;
    NttRoots* roots = nullptr;
    Fr *wsub_f = nullptr, *wsub_i = nullptr;                              // omega_{2^kw}^{+-j}, j < 2^(kw-1)
    Fr *tw_lo_f = nullptr, *tw_hi_f = nullptr, *tw_lo_i = nullptr, *tw_hi_i = nullptr;  // omega_N^{+-X}
    Fr *cs_lo = nullptr, *cs_hi = nullptr, *cs_hi_n = nullptr;            // g^j ; hi / hi * 1/N
    Fr *ci_lo = nullptr, *ci_hi_n = nullptr, *ci_hi_h = nullptr;          // g^-j ; hi * 1/N ; hi * R/(2N)
    size_t bytes = 0;
};

int ntt_get_tables(zkr_ctx* ctx, int log_n, NttTables** out);
// in-place; dit == false: natural in -> bit-reversed out (DIF); dit == true: bit-reversed in -> natural out
int ntt_run(zkr_ctx* ctx, cudaStream_t st, Fr* data, int log_n, bool dit, bool inverse);
int h_pipeline(zkr_ctx* ctx, cudaStream_t st, Fr* A, Fr* B, Fr* S, Fr* h, int log_m, bool bitrev_out);

}  // namespace zkr
