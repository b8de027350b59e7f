// Witness generation on the GPU for forward-solvable R1CS (C-ABI: zkr_wprog_*, zkr_witness_solve).
//
// SURVEY.md 8(f) rank 4.  In the reference the witness comes from circom's generated calculator,
// `circuit.calculateWitness(circuitInputs)` (/root/reference/operator/src/snarks/common.ts:12-17), which interprets the
// circuit component by component in JavaScript; ~80 % of the rollup circuit's signals are MiMC-Feistel round values
// (/root/reference/prover/circuits/hasher.circom:8, MiMCSponge(length, 220, 1): per round t2 = t*t, t4 = t2*t2,
// xL' = xR + t4*t).  Once the prove itself takes milliseconds that interpreter is the whole latency of
// createProofGenerator.  circom cannot run here, so this is not a port of the calculator: it is a solver for the class
// of constraint systems those circuits compile to on their arithmetic side --
//     every constraint  (A.w) * (B.w) = (C.w)  introduces at most ONE new signal, the one with the largest index,
//     and that signal occurs in C only,
// i.e. new = ((A.w)(B.w) - C'.w) / c_new.  MiMC rounds, Feistel chains, products and linear combinations all have this
// form.  Signals that no constraint defines this way (circuit inputs; bits constrained by b (b - 1) = 0, circom's `<--`
// hints) are GIVEN by the caller.  The synthetic rollup-shaped circuits of simple_zk_rollups_b200/synth.py are of this
// class; for the reference's real circuits the hints (Num2Bits, comparators, EdDSA) stay with the TypeScript host.
//
// Build (once per circuit, host): CSC-by-signal -> CSR-by-row, the defined signal of every row, its level in the
// dependency graph (1 + the deepest operand), operations sorted by level.  Solve (per proof, device): scatter the given
// values, then walk the levels -- all Feistel chains advance one step per level -- in ONE persistent launch with a grid
// barrier between levels (one launch per level is kept behind ZKR_WITNESS_PER_LEVEL=1), with the witness left resident
// in HBM in the layout zkr_prove_dev takes (n x 32 B, standard form): no 27 MB host round trip per proof.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fp.cuh"

using namespace zkr;

struct zkr_wprog {
    zkr_ctx* ctx = nullptr;
    uint32_t n_vars = 0, n_rows = 0, n_given = 0, n_ops = 0, n_levels = 0, n_pool = 0;
    std::vector<uint32_t> given;            // host: signals the caller supplies, ascending
    std::vector<uint32_t> level_ofs;        // host: ops of level l are [level_ofs[l], level_ofs[l + 1])
    // device
    uint32_t *ptr[3] = {}, *sig[3] = {}, *cid[3] = {};   // CSR by row of A, B, C
    uint32_t *op_row = nullptr, *op_out = nullptr, *op_kcid = nullptr;   // sorted by level
    uint32_t* d_given = nullptr;
    uint32_t* d_level_ofs = nullptr;
    unsigned int* bar = nullptr;
    Fr *pool = nullptr, *pool_inv = nullptr;              // Montgomery form
    Fr* given_vals = nullptr;                              // staging, n_given
    int* err = nullptr;
    size_t bytes = 0;
};

namespace {

__global__ void k_pool_prepare(Fr* pool, Fr* pool_inv, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr v = Fr::load(pool + i).to_mont();
    v.store(pool + i);
    (v.is_zero() ? v : v.inverse()).store(pool_inv + i);
}

__global__ void k_scatter_given(Fr* __restrict__ w, const uint32_t* __restrict__ sigs, const Fr* __restrict__ vals, uint32_t n,
                                int* err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr v = Fr::load(vals + i);
    if (!v.in_range()) *err = 1;
    v.store(w + sigs[i]);
}

__device__ __forceinline__ Fr row_lc(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ sig,
                                     const uint32_t* __restrict__ cid, uint32_t row, const Fr* __restrict__ pool,
                                     const Fr* __restrict__ w, uint32_t skip) {
    Fr acc = Fr::zero();
    const uint32_t e = ptr[row + 1];
    for (uint32_t k = ptr[row]; k < e; k++) {
        const uint32_t s = sig[k];
        if (s == skip) continue;
        // w is written by other CTAs of the same launch (persistent solve): read it through L2, never a stale L1 line
        const uint4* q = reinterpret_cast<const uint4*>(w + s);
        const uint4 x0 = __ldcg(q), x1 = __ldcg(q + 1);
        Fr ws;
        ws.v[0] = x0.x; ws.v[1] = x0.y; ws.v[2] = x0.z; ws.v[3] = x0.w;
        ws.v[4] = x1.x; ws.v[5] = x1.y; ws.v[6] = x1.z; ws.v[7] = x1.w;
        acc = acc + ws * Fr::load_ro(pool + cid[k]);     // standard x Montgomery -> standard
    }
    return acc;
}

struct SolveArgs {
    const uint32_t *op_row, *op_out, *op_kcid, *level_ofs;
    const uint32_t *pa, *sa, *ca, *pb, *sb, *cb, *pc, *sc, *cc;
    const Fr *pool, *pool_inv;
    Fr* w;
    uint32_t n_levels;
    unsigned int* bar;     // grid barrier counter (zeroed before the launch)
    int* err;
};

// ALL levels in one launch: a persistent grid of one CTA per SM walks the levels, a counter barrier in global memory
// between them (every CTA is resident: the grid is never larger than the SM count, the kernel uses no shared memory and
// few registers).  663 launches of ~11 us each (launch latency + one dependent-load chain) become 663 barriers of ~3 us.
// The spin is bounded (2 s): a CTA that cannot see its peers sets the error flag instead of hanging the GPU.
__global__ void __launch_bounds__(128) k_solve_all(SolveArgs a) {
    const unsigned nblk = gridDim.x;
    for (uint32_t l = 0; l < a.n_levels; l++) {
        const uint32_t lo = a.level_ofs[l], hi = a.level_ofs[l + 1];
        for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += nblk * blockDim.x) {
            const uint32_t row = a.op_row[i], out = a.op_out[i];
            const Fr la = row_lc(a.pa, a.sa, a.ca, row, a.pool, a.w, 0xffffffffu);
            const Fr lb = row_lc(a.pb, a.sb, a.cb, row, a.pool, a.w, 0xffffffffu);
            const Fr lc = row_lc(a.pc, a.sc, a.cc, row, a.pool, a.w, out);
            const Fr u = (la * lb).to_mont() - lc;
            (u * Fr::load_ro(a.pool_inv + a.op_kcid[i])).store(a.w + out);
        }
        if (l + 1 == a.n_levels) break;
        // grid barrier: this level's stores must be visible to every CTA before the next level reads them
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(a.bar, 1u);
            const unsigned target = (l + 1) * nblk;
            const long long t0 = clock64();
            while (*(volatile unsigned int*)a.bar < target) {
                if (clock64() - t0 > 4000000000ll) {      // ~2 s at 2 GHz
                    *a.err = 2;
                    break;
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
}

// one level of the dependency graph: w[out] = ((A.w)(B.w) - C'.w) / c_out
__global__ void k_solve_level(const uint32_t* __restrict__ op_row, const uint32_t* __restrict__ op_out,
                              const uint32_t* __restrict__ op_kcid, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ pa,
                              const uint32_t* __restrict__ sa, const uint32_t* __restrict__ ca, const uint32_t* __restrict__ pb,
                              const uint32_t* __restrict__ sb, const uint32_t* __restrict__ cb, const uint32_t* __restrict__ pc,
                              const uint32_t* __restrict__ sc, const uint32_t* __restrict__ cc, const Fr* __restrict__ pool,
                              const Fr* __restrict__ pool_inv, Fr* w) {
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint32_t row = op_row[i], out = op_out[i];
    const Fr la = row_lc(pa, sa, ca, row, pool, w, 0xffffffffu);
    const Fr lb = row_lc(pb, sb, cb, row, pool, w, 0xffffffffu);
    const Fr lc = row_lc(pc, sc, cc, row, pool, w, out);
    const Fr u = (la * lb).to_mont() - lc;                  // (la lb / R) R - lc, standard form
    (u * Fr::load_ro(pool_inv + op_kcid[i])).store(w + out);
}

int up(const void* h, size_t bytes, void** d, cudaStream_t st, size_t* total) {
    ZKR_CUDA(cudaMalloc(d, bytes ? bytes : 16));
    if (bytes) ZKR_CUDA(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st));
    ZKR_CUDA(cudaStreamSynchronize(st));
    *total += bytes;
    return ZKR_OK;
}

void wprog_release(zkr_wprog* p) {
    if (!p) return;
    for (int m = 0; m < 3; m++) {
        cudaFree(p->ptr[m]);
        cudaFree(p->sig[m]);
        cudaFree(p->cid[m]);
    }
    void* ps[] = {p->op_row, p->op_out, p->op_kcid, p->d_given, p->pool, p->pool_inv, p->given_vals, p->err, p->d_level_ofs, p->bar};
    for (void* q : ps) cudaFree(q);
    delete p;
}

}  // namespace

extern "C" int zkr_wprog_build(zkr_ctx* ctx, const zkr_r1cs_csc* r, zkr_wprog** out) {
    if (!ctx || !r || !out || r->n_vars == 0 || !r->pool || !r->ptr_a || !r->ptr_b || !r->ptr_c) return ZKR_E_INVALID;
    *out = nullptr;
    const uint32_t n = r->n_vars, nc = r->n_constraints;
    const uint32_t* cp[3] = {r->ptr_a, r->ptr_b, r->ptr_c};
    const uint32_t* cr[3] = {r->row_a, r->row_b, r->row_c};
    const uint32_t* cc[3] = {r->cid_a, r->cid_b, r->cid_c};
    // CSC by signal -> CSR by row
    std::vector<uint32_t> ptr[3], sig[3], cid[3];
    for (int m = 0; m < 3; m++) {
        const uint32_t nnz = cp[m][n];
        ptr[m].assign((size_t)nc + 1, 0);
        for (uint32_t e = 0; e < nnz; e++) {
            if (cr[m][e] >= nc || cc[m][e] >= r->n_pool) {
                set_error("zkr_wprog_build: entry %u of matrix %d is out of range", e, m);
                return ZKR_E_INVALID;
            }
            ptr[m][cr[m][e] + 1]++;
        }
        for (uint32_t c = 0; c < nc; c++) ptr[m][c + 1] += ptr[m][c];
        sig[m].resize(nnz);
        cid[m].resize(nnz);
        std::vector<uint32_t> cur(ptr[m].begin(), ptr[m].end() - 1);
        for (uint32_t s = 0; s < n; s++)
            for (uint32_t e = cp[m][s]; e < cp[m][s + 1]; e++) {
                const uint32_t d = cur[cr[m][e]]++;
                sig[m][d] = s;
                cid[m][d] = cc[m][e];
            }
    }
    // the defined signal of every row, levels
    const uint32_t kUnknown = 0xffffffffu;
    std::vector<uint32_t> level(n, kUnknown);           // kUnknown = not solved by any row (given, unless proven otherwise)
    std::vector<uint8_t> solved(n, 0);
    std::vector<uint32_t> op_row, op_out, op_kcid, op_level;
    for (uint32_t row = 0; row < nc; row++) {
        uint32_t dmax = 0;
        bool any = false;
        for (int m = 0; m < 3; m++)
            for (uint32_t e = ptr[m][row]; e < ptr[m][row + 1]; e++) {
                dmax = std::max(dmax, sig[m][e]);
                any = true;
            }
        if (!any || solved[dmax]) continue;              // empty row, or a check on signals that are already known
        bool in_ab = false;
        for (int m = 0; m < 2; m++)
            for (uint32_t e = ptr[m][row]; e < ptr[m][row + 1]; e++) in_ab |= sig[m][e] == dmax;
        uint32_t kc = kUnknown, times = 0;
        for (uint32_t e = ptr[2][row]; e < ptr[2][row + 1]; e++)
            if (sig[2][e] == dmax) {
                kc = cid[2][e];
                times++;
            }
        if (in_ab || times != 1) continue;               // not of the solvable form: dmax stays a given signal (a hint)
        uint32_t lvl = 0;
        for (int m = 0; m < 3; m++)
            for (uint32_t e = ptr[m][row]; e < ptr[m][row + 1]; e++) {
                const uint32_t s = sig[m][e];
                if (s != dmax && solved[s]) lvl = std::max(lvl, level[s] + 1);
            }
        solved[dmax] = 1;
        level[dmax] = lvl;
        op_row.push_back(row);
        op_out.push_back(dmax);
        op_kcid.push_back(kc);
        op_level.push_back(lvl);
    }
    // a signal used by a solving row must be known when that row runs: every operand is either given or solved by an
    // EARLIER row (rows are taken in order and operands have smaller indices than the defined signal, but a smaller
    // index may still be defined by a later row)
    {
        std::vector<uint32_t> def_pos(n, kUnknown);
        for (size_t i = 0; i < op_out.size(); i++) def_pos[op_out[i]] = (uint32_t)i;
        for (size_t i = 0; i < op_row.size(); i++)
            for (int m = 0; m < 3; m++)
                for (uint32_t e = ptr[m][op_row[i]]; e < ptr[m][op_row[i] + 1]; e++) {
                    const uint32_t s = sig[m][e];
                    if (s != op_out[i] && def_pos[s] != kUnknown && def_pos[s] > i) {
                        set_error("zkr_wprog_build: row %u uses signal %u before the row that defines it: not forward-solvable in row order",
                                  op_row[i], s);
                        return ZKR_E_UNSUPPORTED;
                    }
                }
    }
    zkr_wprog* p = new zkr_wprog();
    p->ctx = ctx;
    p->n_vars = n;
    p->n_rows = nc;
    p->n_pool = r->n_pool;
    p->n_ops = (uint32_t)op_row.size();
    for (uint32_t s = 0; s < n; s++)
        if (!solved[s]) p->given.push_back(s);
    p->n_given = (uint32_t)p->given.size();
    uint32_t nl = 0;
    for (uint32_t l : op_level) nl = std::max(nl, l + 1);
    p->n_levels = nl;
    // counting sort by level
    p->level_ofs.assign((size_t)nl + 1, 0);
    for (uint32_t l : op_level) p->level_ofs[l + 1]++;
    for (uint32_t l = 0; l < nl; l++) p->level_ofs[l + 1] += p->level_ofs[l];
    std::vector<uint32_t> s_row(p->n_ops), s_out(p->n_ops), s_kc(p->n_ops), cur(p->level_ofs.begin(), p->level_ofs.end() - (nl ? 1 : 0));
    for (size_t i = 0; i < op_row.size(); i++) {
        const uint32_t d = cur[op_level[i]]++;
        s_row[d] = op_row[i];
        s_out[d] = op_out[i];
        s_kc[d] = op_kcid[i];
    }
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->s[0];
    int rc = ZKR_OK;
#define WP_TRY(expr)            \
    do {                        \
        rc = (expr);            \
        if (rc != ZKR_OK) {     \
            wprog_release(p);   \
            return rc;          \
        }                       \
    } while (0)
    for (int m = 0; m < 3; m++) {
        WP_TRY(up(ptr[m].data(), ptr[m].size() * 4, (void**)&p->ptr[m], st, &p->bytes));
        WP_TRY(up(sig[m].data(), sig[m].size() * 4, (void**)&p->sig[m], st, &p->bytes));
        WP_TRY(up(cid[m].data(), cid[m].size() * 4, (void**)&p->cid[m], st, &p->bytes));
    }
    WP_TRY(up(s_row.data(), s_row.size() * 4, (void**)&p->op_row, st, &p->bytes));
    WP_TRY(up(s_out.data(), s_out.size() * 4, (void**)&p->op_out, st, &p->bytes));
    WP_TRY(up(s_kc.data(), s_kc.size() * 4, (void**)&p->op_kcid, st, &p->bytes));
    WP_TRY(up(p->given.data(), p->given.size() * 4, (void**)&p->d_given, st, &p->bytes));
    WP_TRY(up(p->level_ofs.data(), p->level_ofs.size() * 4, (void**)&p->d_level_ofs, st, &p->bytes));
    if (cudaMalloc(&p->bar, sizeof(unsigned int)) != cudaSuccess) {
        wprog_release(p);
        return ZKR_E_NOMEM;
    }
    WP_TRY(up(r->pool, 32 * (size_t)r->n_pool, (void**)&p->pool, st, &p->bytes));
    cudaError_t e1 = cudaMalloc(&p->pool_inv, 32 * (size_t)(r->n_pool ? r->n_pool : 1));
    cudaError_t e2 = cudaMalloc(&p->given_vals, 32 * (size_t)(p->n_given ? p->n_given : 1));
    cudaError_t e3 = cudaMalloc(&p->err, sizeof(int));
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        wprog_release(p);
        return ZKR_E_NOMEM;
    }
    cudaMemsetAsync(p->err, 0, sizeof(int), st);
    k_pool_prepare<<<ceil_div(r->n_pool, 64), 64, 0, st>>>(p->pool, p->pool_inv, r->n_pool);
    ctx->launches++;
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        wprog_release(p);
        return cuda_fail(cudaGetLastError(), "zkr_wprog_build", __FILE__, __LINE__);
    }
#undef WP_TRY
    *out = p;
    return ZKR_OK;
}

extern "C" void zkr_wprog_free(zkr_wprog* p) {
    if (!p) return;
    DeviceGuard g(p->ctx->device);
    cudaDeviceSynchronize();
    wprog_release(p);
}

extern "C" int zkr_wprog_info(const zkr_wprog* p, uint32_t* n_vars, uint32_t* n_given, uint32_t* n_solved, uint32_t* n_levels) {
    if (!p) return ZKR_E_INVALID;
    if (n_vars) *n_vars = p->n_vars;
    if (n_given) *n_given = p->n_given;
    if (n_solved) *n_solved = p->n_ops;
    if (n_levels) *n_levels = p->n_levels;
    return ZKR_OK;
}

extern "C" int zkr_wprog_given(const zkr_wprog* p, uint32_t* out_signals) {
    if (!p || !out_signals) return ZKR_E_INVALID;
    memcpy(out_signals, p->given.data(), 4 * (size_t)p->n_given);
    return ZKR_OK;
}

extern "C" int zkr_witness_solve(zkr_ctx* ctx, const zkr_wprog* p, const void* given_values, void* d_witness) {
    if (!ctx || !p || !given_values || !d_witness || p->ctx != ctx) return ZKR_E_INVALID;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->user_stream;
    ZKR_CUDA(cudaMemcpyAsync(p->given_vals, given_values, 32 * (size_t)p->n_given, cudaMemcpyHostToDevice, st));
    ZKR_LAUNCH(ctx, k_scatter_given, ceil_div(p->n_given, 128), 128, 0, st, (Fr*)d_witness, p->d_given, p->given_vals, p->n_given,
               p->err);
    // ZKR_WITNESS_PER_LEVEL=1 (A/B knob): one launch per level instead of the persistent kernel
    const bool per_level = getenv("ZKR_WITNESS_PER_LEVEL") && atoi(getenv("ZKR_WITNESS_PER_LEVEL")) != 0;   // per call: tests toggle it
    if (per_level) {
        for (uint32_t l = 0; l < p->n_levels; l++) {
            const uint32_t lo = p->level_ofs[l], hi = p->level_ofs[l + 1];
            if (hi == lo) continue;
            ZKR_LAUNCH(ctx, k_solve_level, ceil_div(hi - lo, 64), 64, 0, st, p->op_row, p->op_out, p->op_kcid, lo, hi, p->ptr[0],
                       p->sig[0], p->cid[0], p->ptr[1], p->sig[1], p->cid[1], p->ptr[2], p->sig[2], p->cid[2], p->pool, p->pool_inv,
                       (Fr*)d_witness);
        }
    } else if (p->n_levels) {
        SolveArgs a;
        a.op_row = p->op_row; a.op_out = p->op_out; a.op_kcid = p->op_kcid; a.level_ofs = p->d_level_ofs;
        a.pa = p->ptr[0]; a.sa = p->sig[0]; a.ca = p->cid[0];
        a.pb = p->ptr[1]; a.sb = p->sig[1]; a.cb = p->cid[1];
        a.pc = p->ptr[2]; a.sc = p->sig[2]; a.cc = p->cid[2];
        a.pool = p->pool; a.pool_inv = p->pool_inv;
        a.w = (Fr*)d_witness;
        a.n_levels = p->n_levels;
        a.bar = p->bar;
        a.err = p->err;
        ZKR_CUDA(cudaMemsetAsync(p->bar, 0, sizeof(unsigned int), st));
        // never more CTAs than SMs (all must be resident for the barrier); small circuits use fewer
        uint32_t widest = 0;
        for (uint32_t l = 0; l < p->n_levels; l++) widest = std::max(widest, p->level_ofs[l + 1] - p->level_ofs[l]);
        const int blocks = std::max(1, std::min(ctx->sm_count, ceil_div(widest, 128)));
        ZKR_LAUNCH(ctx, k_solve_all, blocks, 128, 0, st, a);
    }
    // the given values are host memory of the caller: they must have been read before we return; the flag rides along
    int e = 0;
    ZKR_CUDA(cudaMemcpyAsync(&e, p->err, sizeof(int), cudaMemcpyDeviceToHost, st));
    ZKR_CUDA(cudaStreamSynchronize(st));
    if (e) {
        ZKR_CUDA(cudaMemsetAsync(p->err, 0, sizeof(int), st));
        if (e == 2) {
            set_error("witness solve: grid barrier timed out (the persistent kernel's CTAs were not co-resident)");
            return ZKR_E_CUDA;
        }
        set_error("a given witness value is >= r");
        return ZKR_E_WITNESS_RANGE;
    }
    return ZKR_OK;
}
