// snarkjs proving-key / witness JSON  ->  websnark binary layouts, natively.
//
// Replaces binarifyProvingKey / binarifyWitness of /root/reference/operator/src/utils/binarify.ts:10-207,
// which the reference runs on EVERY proof (operator/src/snarks/common.ts:27-28) with one `big-integer`
// `times(2^256).mod(p)` per coordinate (binarify.ts:78-90).  Here it is a single streaming pass over the JSON
// text (no DOM, no big-integer objects): decimal strings go straight to 4 x 64-bit limbs in 19-digit chunks,
// Montgomery conversion is one 4x4-limb multiplication by R^2, and the output is byte-identical to the
// reference's ArrayBuffer.  It runs once per circuit; the result feeds zkr_pkey_load_bin.
//
// Host-only translation unit (format conversion, like the reference's: no field math of the prove path).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "keyjson_iface.cuh"

namespace zkr {
namespace {

typedef unsigned __int128 u128;

struct U256 {
    uint64_t l[4];
};

struct Mod {
    uint64_t p[4];
    uint64_t r2[4];    // 2^512 mod p
    uint64_t inv;      // -p^-1 mod 2^64
};

// binarify.ts:80 / :87
const Mod kQ = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full},
                0x87d20782e4866389ull};
const Mod kR = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
                {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull},
                0xc2e1f593efffffffull};

inline bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}

inline void sub_inplace(uint64_t* a, const uint64_t* b) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}

// a * b / 2^256 mod p  (CIOS, a < 2^256 arbitrary, b < p): result < p
inline void mont_mul(const Mod& m, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t k = t[0] * m.inv;
        c = (u128)k * m.p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)k * m.p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
        t[5] = 0;
    }
    // a < 2^256 < 6p: the result is < p + a*b/2^256 < p + p*6p/2^256... bounded by a few p: subtract until < p
    while (t[4] || geq(t, m.p)) {
        u128 borrow = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)t[i] - m.p[i] - (uint64_t)borrow;
            t[i] = (uint64_t)d;
            borrow = (d >> 64) & 1;
        }
        t[4] -= (uint64_t)borrow;
    }
    for (int i = 0; i < 4; i++) out[i] = t[i];
}

// value * 2^256 mod p  (toMontgomeryQ / toMontgomeryR, binarify.ts:78-90; any value < 2^256 accepted)
inline void to_mont(const Mod& m, const U256& v, uint8_t* out32) {
    uint64_t o[4];
    mont_mul(m, v.l, m.r2, o);
    memcpy(out32, o, 32);      // little-endian host: 4 x u64 == 8 x u32 LE (writeBigInt, binarify.ts:68-76)
}

const uint64_t kPow10[20] = {1ull,
                             10ull,
                             100ull,
                             1000ull,
                             10000ull,
                             100000ull,
                             1000000ull,
                             10000000ull,
                             100000000ull,
                             1000000000ull,
                             10000000000ull,
                             100000000000ull,
                             1000000000000ull,
                             10000000000000ull,
                             100000000000000ull,
                             1000000000000000ull,
                             10000000000000000ull,
                             100000000000000000ull,
                             1000000000000000000ull,
                             10000000000000000000ull};

// decimal digits -> 256-bit integer; false on a non-digit, an empty string or overflow past 2^256
inline bool parse_decimal(const char* s, size_t n, U256& out) {
    if (n == 0) return false;
    uint64_t a[4] = {0, 0, 0, 0};
    size_t i = 0;
    while (i < n) {
        size_t take = std::min<size_t>(19, n - i);
        uint64_t chunk = 0;
        for (size_t k = 0; k < take; k++) {
            unsigned d = (unsigned char)s[i + k] - '0';
            if (d > 9) return false;
            chunk = chunk * 10 + d;
        }
        u128 c = chunk;
        const uint64_t mul = kPow10[take];
        for (int j = 0; j < 4; j++) {
            c += (u128)a[j] * mul;
            a[j] = (uint64_t)c;
            c >>= 64;
        }
        if (c) return false;
        i += take;
    }
    memcpy(out.l, a, 32);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Minimal streaming JSON reader, specialised for the snarkjs key schema.
struct Reader {
    const char* p;
    const char* end;
    const char* begin;
    char msg[200];
    bool fail(const char* what) {
        snprintf(msg, sizeof msg, "%s at byte %zu", what, (size_t)(p - begin));
        return false;
    }
    inline void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
    }
    inline bool eat(char c) {
        ws();
        if (p < end && *p == c) {
            p++;
            return true;
        }
        return false;
    }
    inline bool peek(char c) {
        ws();
        return p < end && *p == c;
    }
    // "..." (no escapes expected inside keys / numbers; escapes are skipped verbatim)
    bool string(const char*& s, size_t& n) {
        ws();
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        s = p;
        const char* q = (const char*)memchr(p, '"', end - p);
        while (q && q > s && q[-1] == '\\') q = (const char*)memchr(q + 1, '"', end - q - 1);
        if (!q) return fail("unterminated string");
        n = q - s;
        p = q + 1;
        return true;
    }
    // big integer given as "123" or bare 123 (stringifyBigInts / plain JSON number)
    bool bigint(U256& v) {
        ws();
        const char* s;
        size_t n;
        if (p < end && *p == '"') {
            if (!string(s, n)) return false;
        } else {
            s = p;
            while (p < end && *p >= '0' && *p <= '9') p++;
            n = p - s;
        }
        if (!parse_decimal(s, n, v)) return fail("expected a decimal integer < 2^256");
        return true;
    }
    bool u32(uint32_t& v) {
        U256 x;
        if (!bigint(x)) return false;
        if (x.l[1] | x.l[2] | x.l[3] || x.l[0] > 0xffffffffull) return fail("integer does not fit 32 bits");
        v = (uint32_t)x.l[0];
        return true;
    }
    bool skip_value() {
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '"') {
            const char* s;
            size_t n;
            return string(s, n);
        }
        if (*p == '{' || *p == '[') {
            int depth = 0;
            while (p < end) {
                char c = *p;
                if (c == '"') {
                    const char* s;
                    size_t n;
                    if (!string(s, n)) return false;
                    continue;
                }
                if (c == '{' || c == '[') depth++;
                if (c == '}' || c == ']') depth--;
                p++;
                if (depth == 0) return true;
            }
            return fail("unterminated container");
        }
        while (p < end && *p != ',' && *p != '}' && *p != ']') p++;
        return true;
    }
    bool null_literal() {
        ws();
        if (end - p >= 4 && memcmp(p, "null", 4) == 0) {
            p += 4;
            return true;
        }
        return false;
    }
};

typedef std::vector<uint8_t> Bytes;

inline uint8_t* grow(Bytes& b, size_t n) {
    size_t o = b.size();
    b.resize(o + n);
    return b.data() + o;
}

// [x, y, z] -> x|y Fq-M (writePoint, binarify.ts:92-95: z is dropped)
bool g1_point(Reader& r, uint8_t* out) {
    if (!r.eat('[')) return r.fail("expected G1 point [x,y,z]");
    U256 v;
    for (int i = 0;; i++) {
        if (!r.bigint(v)) return false;
        if (i < 2) to_mont(kQ, v, out + 32 * i);
        if (r.eat(',')) continue;
        if (r.eat(']')) {
            if (i < 1) return r.fail("G1 point needs two coordinates");
            return true;
        }
        return r.fail("malformed G1 point");
    }
}

// [[x0,x1],[y0,y1],[z0,z1]] -> x0|x1|y0|y1 Fq-M (writePoint2, binarify.ts:97-102)
bool g2_point(Reader& r, uint8_t* out) {
    if (!r.eat('[')) return r.fail("expected G2 point [[x0,x1],[y0,y1],[z0,z1]]");
    U256 v;
    for (int i = 0;; i++) {
        if (!r.eat('[')) return r.fail("expected Fq2 element [c0,c1]");
        for (int j = 0; j < 2; j++) {
            if (!r.bigint(v)) return false;
            if (i < 2) to_mont(kQ, v, out + 64 * i + 32 * j);
            if (j == 0 && !r.eat(',')) return r.fail("Fq2 element needs two coefficients");
        }
        if (!r.eat(']')) return r.fail("Fq2 element has more than two coefficients");
        if (r.eat(',')) continue;
        if (r.eat(']')) {
            if (i < 1) return r.fail("G2 point needs two coordinates");
            return true;
        }
        return r.fail("malformed G2 point");
    }
}

// array of points; `first`: leading entries may be null (C[0..l], A.3) and are not emitted.
// Entry i >= first that is null is an error (the reference would throw on p[0] of null).
bool point_array(Reader& r, int group, Bytes& out, uint64_t& count, uint64_t& leading_nulls) {
    if (!r.eat('[')) return r.fail("expected array of points");
    const size_t pb = group == 1 ? 64 : 128;
    count = 0;
    leading_nulls = 0;
    if (r.eat(']')) return true;
    for (;;) {
        if (r.null_literal()) {
            if (count != leading_nulls) return r.fail("null point after a non-null one");
            leading_nulls++;
        } else {
            uint8_t* o = grow(out, pb);
            if (!(group == 1 ? g1_point(r, o) : g2_point(r, o))) return false;
        }
        count++;
        if (r.eat(',')) continue;
        if (r.eat(']')) return true;
        return r.fail("malformed point array");
    }
}

// polsA / polsB: [ {"row": "coef", ...}, ... ] -> per signal: u32 count, count x (u32 row, Fr-M coef)
// (writeTransformedPolynomial, binarify.ts:104-113; Object.keys order = ascending integer keys)
bool pol_array(Reader& r, Bytes& out, uint64_t& count) {
    if (!r.eat('[')) return r.fail("expected array of polynomials");
    count = 0;
    if (r.eat(']')) return true;
    std::vector<std::pair<uint32_t, U256>> ent;
    for (;;) {
        if (!r.eat('{')) return r.fail("expected polynomial object");
        ent.clear();
        bool sorted = true;
        if (!r.eat('}')) {
            for (;;) {
                const char* ks;
                size_t kn;
                if (!r.string(ks, kn)) return false;
                U256 k;
                if (!parse_decimal(ks, kn, k) || (k.l[1] | k.l[2] | k.l[3]) || k.l[0] > 0xfffffffeull)
                    return r.fail("polynomial key is not a row index");
                if (!r.eat(':')) return r.fail("expected ':'");
                U256 v;
                if (!r.bigint(v)) return false;
                if (!ent.empty() && ent.back().first >= (uint32_t)k.l[0]) sorted = false;
                ent.emplace_back((uint32_t)k.l[0], v);
                if (r.eat(',')) continue;
                if (r.eat('}')) break;
                return r.fail("malformed polynomial object");
            }
        }
        if (!sorted) {
            // JS objects enumerate integer keys in ascending order and keep the LAST value of a repeated key
            std::stable_sort(ent.begin(), ent.end(),
                             [](const std::pair<uint32_t, U256>& a, const std::pair<uint32_t, U256>& b) { return a.first < b.first; });
            size_t w = 0;
            for (size_t i = 0; i < ent.size(); i++) {
                if (i + 1 < ent.size() && ent[i + 1].first == ent[i].first) continue;
                ent[w++] = ent[i];
            }
            ent.resize(w);
        }
        uint8_t* o = grow(out, 4 + 36 * ent.size());
        const uint32_t cnt = (uint32_t)ent.size();
        memcpy(o, &cnt, 4);
        o += 4;
        for (auto& e : ent) {
            memcpy(o, &e.first, 4);
            to_mont(kR, e.second, o + 4);
            o += 36;
        }
        count++;
        if (r.eat(',')) continue;
        if (r.eat(']')) return true;
        return r.fail("malformed polynomial array");
    }
}

int bad(const Reader& r) {
    set_error("proving key JSON: %s", r.msg);
    return ZKR_E_BADKEY;
}

}  // namespace
}  // namespace zkr

int zkr::vkey_parse_json(const char* json, size_t len, VKeyRaw* out) {
    Reader r;
    r.p = r.begin = json;
    r.end = json + len;
    r.msg[0] = 0;
    auto bad_vk = [&]() {
        set_error("verification key JSON: %s", r.msg);
        return ZKR_E_BADKEY;
    };
    bool have[6] = {false, false, false, false, false, false};   // nPublic, IC, alfa1, beta2, gamma2, delta2
    uint64_t n_ic = 0;
    try {
        if (!r.eat('{')) {
            r.fail("expected '{'");
            return bad_vk();
        }
        if (!r.eat('}')) {
            for (;;) {
                const char* ks;
                size_t kn;
                if (!r.string(ks, kn)) return bad_vk();
                if (!r.eat(':')) {
                    r.fail("expected ':'");
                    return bad_vk();
                }
                auto is = [&](const char* name) { return kn == strlen(name) && memcmp(ks, name, kn) == 0; };
                if (is("nPublic")) {
                    if (!r.u32(out->n_public)) return bad_vk();
                    have[0] = true;
                } else if (is("IC")) {
                    uint64_t nulls = 0;
                    out->ic.clear();
                    if (!point_array(r, 1, out->ic, n_ic, nulls)) return bad_vk();
                    if (nulls) {
                        r.fail("null point in IC");
                        return bad_vk();
                    }
                    have[1] = true;
                } else if (is("vk_alfa_1")) {
                    if (!g1_point(r, out->alfa1)) return bad_vk();
                    have[2] = true;
                } else if (is("vk_beta_2")) {
                    if (!g2_point(r, out->beta2)) return bad_vk();
                    have[3] = true;
                } else if (is("vk_gamma_2")) {
                    if (!g2_point(r, out->gamma2)) return bad_vk();
                    have[4] = true;
                } else if (is("vk_delta_2")) {
                    if (!g2_point(r, out->delta2)) return bad_vk();
                    have[5] = true;
                } else if (!r.skip_value()) {
                    return bad_vk();
                }
                if (r.eat(',')) continue;
                if (r.eat('}')) break;
                r.fail("malformed object");
                return bad_vk();
            }
        }
    } catch (const std::bad_alloc&) {
        set_error("out of host memory while parsing the verification key JSON");
        return ZKR_E_NOMEM;
    }
    static const char* const names[6] = {"nPublic", "IC", "vk_alfa_1", "vk_beta_2", "vk_gamma_2", "vk_delta_2"};
    for (int i = 0; i < 6; i++)
        if (!have[i]) {
            set_error("verification key JSON: field %s missing", names[i]);
            return ZKR_E_BADKEY;
        }
    if (n_ic != (uint64_t)out->n_public + 1) {
        set_error("verification key JSON: IC has %llu entries, nPublic + 1 = %llu expected", (unsigned long long)n_ic,
                  (unsigned long long)out->n_public + 1);
        return ZKR_E_BADKEY;
    }
    return ZKR_OK;
}

using namespace zkr;

extern "C" int zkr_pkey_json_to_bin(const char* json, size_t len, void** out_buf, size_t* out_len) {
    if (!json || !out_buf || !out_len) {
        set_error("zkr_pkey_json_to_bin: null argument");
        return ZKR_E_INVALID;
    }
    *out_buf = nullptr;
    *out_len = 0;
    Reader r;
    r.p = r.begin = json;
    r.end = json + len;
    r.msg[0] = 0;
    try {
        uint32_t n_vars = 0, n_public = 0, domain = 0;
        bool have_n = false, have_l = false, have_m = false;
        Bytes vk[5];                 // alfa1, beta1, delta1 (64 B), beta2, delta2 (128 B)
        Bytes sec[7];                // polsA, polsB, A, B1, B2, C, hExps
        uint64_t cnt[7] = {0, 0, 0, 0, 0, 0, 0}, c_nulls = 0;
        bool have_sec[7] = {false, false, false, false, false, false, false};
        bool have_vk[5] = {false, false, false, false, false};
        static const char* const sec_names[7] = {"polsA", "polsB", "A", "B1", "B2", "C", "hExps"};
        static const char* const vk_names[5] = {"vk_alfa_1", "vk_beta_1", "vk_delta_1", "vk_beta_2", "vk_delta_2"};
        if (!r.eat('{')) {
            r.fail("expected '{'");
            return bad(r);
        }
        if (!r.eat('}')) {
            for (;;) {
                const char* ks;
                size_t kn;
                if (!r.string(ks, kn)) return bad(r);
                if (!r.eat(':')) {
                    r.fail("expected ':'");
                    return bad(r);
                }
                auto is = [&](const char* name) { return kn == strlen(name) && memcmp(ks, name, kn) == 0; };
                bool done = false;
                if (is("nVars")) {
                    if (!r.u32(n_vars)) return bad(r);
                    have_n = done = true;
                } else if (is("nPublic")) {
                    if (!r.u32(n_public)) return bad(r);
                    have_l = done = true;
                } else if (is("domainSize")) {
                    if (!r.u32(domain)) return bad(r);
                    have_m = done = true;
                }
                for (int i = 0; i < 7 && !done; i++) {
                    if (!is(sec_names[i])) continue;
                    sec[i].clear();
                    if (i < 2) {
                        sec[i].reserve(len / 8);
                        if (!pol_array(r, sec[i], cnt[i])) return bad(r);
                    } else {
                        uint64_t nulls = 0;
                        if (!point_array(r, i == 4 ? 2 : 1, sec[i], cnt[i], nulls)) return bad(r);
                        if (i == 5) c_nulls = nulls;
                        else if (nulls) {
                            r.fail("null point outside C");
                            return bad(r);
                        }
                    }
                    have_sec[i] = done = true;
                }
                for (int i = 0; i < 5 && !done; i++) {
                    if (!is(vk_names[i])) continue;
                    vk[i].assign(i < 3 ? 64 : 128, 0);
                    if (!(i < 3 ? g1_point(r, vk[i].data()) : g2_point(r, vk[i].data()))) return bad(r);
                    have_vk[i] = done = true;
                }
                if (!done && !r.skip_value()) return bad(r);     // protocol, domainBits, polsC, ...
                if (r.eat(',')) continue;
                if (r.eat('}')) break;
                r.fail("malformed object");
                return bad(r);
            }
        }
        if (!have_n || !have_l || !have_m) {
            set_error("proving key JSON: nVars / nPublic / domainSize missing");
            return ZKR_E_BADKEY;
        }
        for (int i = 0; i < 7; i++)
            if (!have_sec[i]) {
                set_error("proving key JSON: field %s missing", sec_names[i]);
                return ZKR_E_BADKEY;
            }
        for (int i = 0; i < 5; i++)
            if (!have_vk[i]) {
                set_error("proving key JSON: field %s missing", vk_names[i]);
                return ZKR_E_BADKEY;
            }
        if ((uint64_t)n_public + 1 > n_vars) {
            set_error("proving key JSON: nPublic %u >= nVars %u", n_public, n_vars);
            return ZKR_E_BADKEY;
        }
        // the reference indexes [0, nVars) / [nPublic+1, nVars) / [0, domainSize) and would throw on a short array
        const uint64_t want[7] = {n_vars, n_vars, n_vars, n_vars, n_vars, n_vars, domain};
        for (int i = 0; i < 7; i++) {
            if (cnt[i] < want[i]) {
                set_error("proving key JSON: %s has %llu entries, %llu needed", sec_names[i], (unsigned long long)cnt[i],
                          (unsigned long long)want[i]);
                return ZKR_E_BADKEY;
            }
        }
        if (c_nulls > (uint64_t)n_public + 1) {
            set_error("proving key JSON: C has %llu leading nulls, at most nPublic + 1 = %u allowed",
                      (unsigned long long)c_nulls, n_public + 1);
            return ZKR_E_BADKEY;
        }
        // byte ranges actually serialised (arrays may be longer than the header says; the reference ignores the rest)
        size_t take[7];
        {
            // pols: walk the per-signal records to find the end of signal nVars-1
            for (int i = 0; i < 2; i++) {
                size_t off = 0;
                for (uint32_t s = 0; s < n_vars; s++) {
                    uint32_t k;
                    memcpy(&k, sec[i].data() + off, 4);
                    off += 4 + 36 * (size_t)k;
                }
                take[i] = off;
            }
            take[2] = 64 * (size_t)n_vars;
            take[3] = 64 * (size_t)n_vars;
            take[4] = 128 * (size_t)n_vars;
            take[6] = 64 * (size_t)domain;
        }
        // C: emitted entries start at index c_nulls; the layout wants indices nPublic+1 .. nVars-1
        const size_t c_skip = 64 * ((size_t)n_public + 1 - c_nulls);
        take[5] = 64 * ((size_t)n_vars - n_public - 1);
        size_t total = 40 + 3 * 64 + 2 * 128;
        for (int i = 0; i < 7; i++) total += take[i];
        if (total > 0xffffffffull) {
            set_error("proving key JSON: binary layout needs %zu bytes; websnark offsets are 32-bit", total);
            return ZKR_E_UNSUPPORTED;
        }
        uint8_t* out = (uint8_t*)malloc(total);
        if (!out) {
            set_error("out of host memory (%zu bytes)", total);
            return ZKR_E_NOMEM;
        }
        uint32_t hdr[10] = {n_vars, n_public, domain, 0, 0, 0, 0, 0, 0, 0};
        size_t off = 40;
        for (int i = 0; i < 5; i++) {
            memcpy(out + off, vk[i].data(), vk[i].size());
            off += vk[i].size();
        }
        for (int i = 0; i < 7; i++) {
            hdr[3 + i] = (uint32_t)off;
            memcpy(out + off, sec[i].data() + (i == 5 ? c_skip : 0), take[i]);
            off += take[i];
        }
        memcpy(out, hdr, 40);
        *out_buf = out;
        *out_len = total;
        return ZKR_OK;
    } catch (const std::bad_alloc&) {
        set_error("out of host memory while parsing the proving key JSON");
        return ZKR_E_NOMEM;
    }
}

extern "C" void zkr_buf_free(void* buf) { free(buf); }

extern "C" int zkr_witness_json_to_bin(const char* json, size_t len, void** out_buf, size_t* out_len) {
    if (!json || !out_buf || !out_len) {
        set_error("zkr_witness_json_to_bin: null argument");
        return ZKR_E_INVALID;
    }
    *out_buf = nullptr;
    *out_len = 0;
    Reader r;
    r.p = r.begin = json;
    r.end = json + len;
    r.msg[0] = 0;
    try {
        Bytes w;
        w.reserve(len / 2);
        if (!r.eat('[')) {
            r.fail("expected '['");
            set_error("witness JSON: %s", r.msg);
            return ZKR_E_INVALID;
        }
        if (!r.eat(']')) {
            for (;;) {
                U256 v;
                if (!r.bigint(v)) {
                    set_error("witness JSON: %s", r.msg);
                    return ZKR_E_INVALID;
                }
                memcpy(grow(w, 32), v.l, 32);   // written as is (binarify.ts:18-26); zkr_prove rejects values >= r
                if (r.eat(',')) continue;
                if (r.eat(']')) break;
                r.fail("malformed array");
                set_error("witness JSON: %s", r.msg);
                return ZKR_E_INVALID;
            }
        }
        uint8_t* out = (uint8_t*)malloc(w.size() ? w.size() : 1);
        if (!out) {
            set_error("out of host memory");
            return ZKR_E_NOMEM;
        }
        memcpy(out, w.data(), w.size());
        *out_buf = out;
        *out_len = w.size();
        return ZKR_OK;
    } catch (const std::bad_alloc&) {
        set_error("out of host memory while parsing the witness JSON");
        return ZKR_E_NOMEM;
    }
}

extern "C" int zkr_pkey_load_json(zkr_ctx* ctx, const char* json, size_t len, zkr_pkey** out) {
    void* buf = nullptr;
    size_t blen = 0;
    ZKR_TRY(zkr_pkey_json_to_bin(json, len, &buf, &blen));
    int rc = zkr_pkey_load_bin(ctx, buf, blen, out);
    free(buf);
    return rc;
}
