// Element-wise test hooks and integer-pipe microbenchmarks (C-ABI: zkr_test_*, zkr_microbench).
// These exist so the parity tests can drive the device field / curve arithmetic directly against
// the CPU oracle, and so that roofline fractions can be quoted against a MEASURED IMAD peak
// (MEASURED_PEAKS.json has none).
#include "common.cuh"
#include "ec.cuh"

using namespace zkr;

namespace {

template <class F>
__global__ void k_field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = F::load(a + 8 * i);
    F y = b ? F::load(b + 8 * i) : F::zero();
    F r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.sqr(); break;
        case 4: r = x.inverse(); break;
        case 5: r = x.to_mont(); break;
        case 6: r = x.from_mont(); break;
        case 7: r = F::mul2(x, y, x + y, x - y); break;      // x y + (x + y)(x - y), one reduction (fp.cuh mont_mul2_raw)
        case 8: r = F::msub(x, y, x + y, x - y); break;      // x y - (x + y)(x - y)
        case 10: r = x.sqr_fast(); break;                    // the dedicated squaring (fp.cuh mont_sqr_raw)
        default: {                                           // 9: four products x y + (x+y)(x-y) + x (x-y) + (x+y) y
            F s = x + y, d = x - y;
            mont_mul4_raw<typename F::Params>(r.v, x.v, y.v, s.v, d.v, x.v, d.v, s.v, y.v);
        }
    }
    r.store(out + 8 * i);
}

template <class F>
__global__ void k_curve_op(int op, const char* p, const char* q, char* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr size_t AB = 2 * sizeof(F);
    Affine<F> P = Affine<F>::load(p + AB * i);
    XYZZ<F> acc = XYZZ<F>::from_affine(P);
    if (op == 0) {
        Affine<F> Q = Affine<F>::load(q + AB * i);
        if (!Q.is_inf()) acc.madd(Q);
    } else if (op == 1) {
        acc = acc.dbl();
    } else if (op == 2) {
        acc = scalar_mul(acc, Fr::load(q + 32 * i));
    } else if (op == 3) {
        Affine<F> Q = Affine<F>::load(q + AB * i);
        XYZZ<F> d = acc.dbl();
        XYZZ<F> qq = XYZZ<F>::from_affine(Q).dbl();            // non-trivial zz on both operands
        d.add(qq);                                             // 2P + 2Q
        acc = d;
    } else {                                                   // 4 / 5: 2P + Q through madd / madd_lazy (non-trivial zz)
        Affine<F> Q = Affine<F>::load(q + AB * i);
        acc = acc.dbl();
        if (!Q.is_inf()) {
            if (op == 4) acc.madd(Q);
            else if (op == 5) acc.madd_lazy(Q);
            else acc.template madd_lazy<true>(Q);
        }
    }
    acc.to_affine().store(out + AB * i);
}

__global__ void k_bench_imad(uint32_t* out, int iters, uint32_t a, uint32_t b) {
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    if (s == 0x12345678u) out[0] = s;
}

__global__ void k_bench_imad_wide(uint32_t* out, int iters, uint32_t b) {
    unsigned long long x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t lo = (uint32_t)x[j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[j]) : "r"(lo), "r"(b));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= x[j];
    if (s == 0x12345678ull) out[0] = (uint32_t)s;
}

__global__ void k_bench_modmul(uint32_t* out, int iters) {
    Fq x = Fq::one(), y = Fq::r2(), z = Fq::r2();
    x.v[0] += threadIdx.x;
    z.v[1] ^= threadIdx.x;
    for (int it = 0; it < iters; it++) {
        x = x * y;
        z = z * y;
    }
    Fq s = x + z;
    if (s.v[0] == 0x12345678u && s.v[7] == 1) out[0] = s.v[3];
}

__global__ void k_bench_madd(uint32_t* out, int iters) {
    // acc = (1,2) generator; add 2G-affine-ish distinct point repeatedly: use P = G, acc starts at 2G
    G1Affine g;
    g.x = Fq::one();
    g.y = Fq::one().dbl();
    G1XYZZ acc = G1XYZZ::dbl_affine(g);
    for (int it = 0; it < iters; it++) acc.madd(g);
    if (acc.x.v[0] == 0x12345678u && acc.zz.v[7] == 1) out[0] = acc.y.v[3];
}

// 12 / 13 / 14: register-resident chains of the lazily reduced G1 addition and of the G2 addition in both forms
template <class F, int LAZY>
__global__ void k_bench_madd_v(uint32_t* out, int iters) {
    Affine<F> g;
    g.x = F::one();
    g.y = F::one().dbl();
    XYZZ<F> acc = XYZZ<F>::dbl_affine(g);
    for (int it = 0; it < iters; it++) {
        if (LAZY == 2) acc.template madd_lazy<true>(g);
        else if (LAZY == 1) acc.template madd_lazy<false>(g);
        else acc.madd(g);
    }
    if (acc.is_inf()) acc.store(out + 64);   // never true for this chain; keeps the additions live
}

// ---- what would batched-affine bucket accumulation cost?  (measure, do not cost on paper)
// 8: Fermat inversions (a^(q-2), ~380 dependent modmuls), every thread of a full grid its own value
__global__ void k_bench_inv_fermat(uint32_t* out, int iters) {
    Fq a = Fq::one().dbl();
    a.v[0] += threadIdx.x + blockIdx.x * 7u;
    for (int it = 0; it < iters; it++) a = a.inverse() + Fq::one();
    if (a.v[0] == 0x12345678u && a.v[7] == 1) out[0] = a.v[3];
}
// 9: the binary extended Euclid of fp_inv.cuh (no multiplier, data-dependent loops) on a full grid: the lanes of a warp
// diverge, which is why the library only uses it on single-thread tails
__global__ void k_bench_inv_euclid(uint32_t* out, int iters) {
    Fq a = Fq::one().dbl();
    a.v[0] += threadIdx.x * 2654435761u + blockIdx.x * 7u;
    for (int it = 0; it < iters; it++) a = a.inverse_vartime() + Fq::one();
    if (a.v[0] == 0x12345678u && a.v[7] == 1) out[0] = a.v[3];
}
// 10: affine + affine with Montgomery's trick over a THREAD-LOCAL batch of B pairs (the denominators' prefix products in
// shared memory, one Fermat inversion per thread and batch): lambda = (y2 - y1) / (x2 - x1), x3 = lambda^2 - x1 - x2,
// y3 = lambda (x1 - x3) - y1 = 5 multiplications + 1 squaring per addition + 1 inversion per batch.  Operands are
// re-derived in registers (no table gathers): this is the arithmetic ceiling of the scheme, to be compared with
// which = 3 (XYZZ mixed additions) at equal thread counts.
template <int B>
__global__ void k_bench_batch_affine(uint32_t* out, int iters) {
    extern __shared__ uint32_t bsm[];
    uint32_t* pre = bsm + threadIdx.x;                  // [B][8][blockDim.x] words, word-major: conflict free
    const int stride = blockDim.x;
    G1Affine g;
    g.x = Fq::one();
    g.y = Fq::one().dbl();
    G1Affine p = XYZZ<Fq>::dbl_affine(g).to_affine();   // 2G
    p.x.v[0] ^= 0;                                       // same point in every thread: the arithmetic does not care
    Fq chk = Fq::zero();
    for (int it = 0; it < iters; it++) {
        // forward: d_i = x(Q_i) - x(P_i) with P_i = p, Q_i = (x(p) + i + 1, ...) stand-ins; prefix products
        Fq acc = Fq::one();
#pragma unroll 1
        for (int i = 0; i < B; i++) {
            Fq qx = p.x;
            qx.v[0] += (uint32_t)(i + 1 + it);
            const Fq d = qx - p.x;
#pragma unroll
            for (int l = 0; l < 8; l++) pre[(i * 8 + l) * stride] = acc.v[l];
            acc = acc * d;
        }
        Fq inv = acc.inverse();
        // backward: 1/d_i = inv * pre_i; inv *= d_i; then the addition itself
#pragma unroll 1
        for (int i = B - 1; i >= 0; i--) {
            Fq qx = p.x, qy = p.y;
            qx.v[0] += (uint32_t)(i + 1 + it);
            qy.v[1] ^= (uint32_t)i;
            const Fq d = qx - p.x;
            Fq pi;
#pragma unroll
            for (int l = 0; l < 8; l++) pi.v[l] = pre[(i * 8 + l) * stride];
            const Fq di = inv * pi;
            inv = inv * d;
            const Fq lam = (qy - p.y) * di;
            const Fq x3 = lam.sqr() - p.x - qx;
            const Fq y3 = lam * (p.x - x3) - p.y;
            chk = chk + x3 + y3;
        }
    }
    if (chk.v[0] == 0x12345678u && chk.v[7] == 1) out[0] = chk.v[3];
}

// FP64 pipe probes (B200 keeps the full-rate FP64 unit): can DFMA carry part of the limb products?
__global__ void k_bench_dfma(uint32_t* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = (double)(threadIdx.x + j);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
    if (s == 0.12345678) out[0] = 1;
}

// 16 DFMA + 16 IMAD.WIDE per inner iteration, interleaved: do the two pipes overlap?
__global__ void k_bench_dfma_imadw(uint32_t* out, int iters, double a, double b, uint32_t m) {
    double x[8];
    unsigned long long y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { x[j] = (double)(threadIdx.x + j); y[j] = threadIdx.x + j; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(a), "d"(b));
                uint32_t lo = (uint32_t)y[j];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[j]) : "r"(lo), "r"(m));
            }
        }
    }
    double s = 0;
    unsigned long long t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { s += x[j]; t ^= y[j]; }
    if (s == 0.12345678 && t == 5) out[0] = 1;
}

// 16 DFMA + 16 64-bit integer adds (2 x IADD3 with carry) per inner iteration: Emmart-style accumulation mix
__global__ void k_bench_dfma_iadd(uint32_t* out, int iters, double a, double b) {
    double x[8];
    unsigned long long y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { x[j] = (double)(threadIdx.x + j); y[j] = threadIdx.x + j; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[j]) : "d"(a), "d"(b));
                y[j] += (unsigned long long)__double_as_longlong(x[(j + 4) & 7]);
            }
        }
    }
    double s = 0;
    unsigned long long t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { s += x[j]; t ^= y[j]; }
    if (s == 0.12345678 && t == 5) out[0] = 1;
}

// plain 64-bit add chains (alu pipe): IADD3 + IADD3.X rate
__global__ void k_bench_iadd64(uint32_t* out, int iters, unsigned long long a) {
    unsigned long long y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) y[j] = threadIdx.x + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("add.u64 %0, %0, %1;" : "+l"(y[j]) : "l"(a + j));
        }
    }
    unsigned long long t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) t ^= y[j];
    if (t == 5) out[0] = 1;
}

}  // namespace

extern "C" int zkr_test_field_op(zkr_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n) {
    if (!ctx || !a || !out || n == 0 || op < 0 || op > 10 || (field != 0 && field != 1)) return ZKR_E_INVALID;
    if ((op <= 2 || (op >= 7 && op <= 9)) && !b) return ZKR_E_INVALID;
    DeviceGuard g(ctx->device);
    uint32_t *da = nullptr, *db = nullptr, *dout = nullptr;
    ZKR_CUDA(cudaMalloc(&da, n * 32));
    ZKR_CUDA(cudaMalloc(&dout, n * 32));
    ZKR_CUDA(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->s[0]));
    if (b) {
        ZKR_CUDA(cudaMalloc(&db, n * 32));
        ZKR_CUDA(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->s[0]));
    }
    int grid = ceil_div(n, 128);
    if (field == 0) ZKR_LAUNCH(ctx, k_field_op<Fq>, grid, 128, 0, ctx->s[0], op, da, db, dout, n);
    else ZKR_LAUNCH(ctx, k_field_op<Fr>, grid, 128, 0, ctx->s[0], op, da, db, dout, n);
    ZKR_CUDA(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->s[0]));
    ZKR_CUDA(cudaStreamSynchronize(ctx->s[0]));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return ZKR_OK;
}

extern "C" int zkr_test_curve_op(zkr_ctx* ctx, int group, int op, const void* p, const void* q, void* out, size_t n) {
    if (!ctx || !p || !out || n == 0 || op < 0 || op > 6 || (group != 1 && group != 2)) return ZKR_E_INVALID;
    if (op != 1 && !q) return ZKR_E_INVALID;
    DeviceGuard g(ctx->device);
    size_t ab = group == 1 ? 64 : 128;
    size_t qb = (op == 2) ? 32 : ab;
    char *dp = nullptr, *dq = nullptr, *dout = nullptr;
    ZKR_CUDA(cudaMalloc(&dp, n * ab));
    ZKR_CUDA(cudaMalloc(&dout, n * ab));
    ZKR_CUDA(cudaMemcpyAsync(dp, p, n * ab, cudaMemcpyHostToDevice, ctx->s[0]));
    if (q) {
        ZKR_CUDA(cudaMalloc(&dq, n * qb));
        ZKR_CUDA(cudaMemcpyAsync(dq, q, n * qb, cudaMemcpyHostToDevice, ctx->s[0]));
    }
    int grid = ceil_div(n, 64);
    if (group == 1) ZKR_LAUNCH(ctx, k_curve_op<Fq>, grid, 64, 0, ctx->s[0], op, dp, dq, dout, n);
    else ZKR_LAUNCH(ctx, k_curve_op<Fq2>, grid, 64, 0, ctx->s[0], op, dp, dq, dout, n);
    ZKR_CUDA(cudaMemcpyAsync(out, dout, n * ab, cudaMemcpyDeviceToHost, ctx->s[0]));
    ZKR_CUDA(cudaStreamSynchronize(ctx->s[0]));
    cudaFree(dp);
    cudaFree(dq);
    cudaFree(dout);
    return ZKR_OK;
}

extern "C" int zkr_microbench(zkr_ctx* ctx, int which, int iters, double* ops_per_s, float* ms_out) {
    if (!ctx || which < 0 || which > 15 || iters <= 0 || !ops_per_s) return ZKR_E_INVALID;
    DeviceGuard g(ctx->device);
    uint32_t* dout = nullptr;
    ZKR_CUDA(cudaMalloc(&dout, 1024));
    cudaEvent_t e0, e1;
    ZKR_CUDA(cudaEventCreate(&e0));
    ZKR_CUDA(cudaEventCreate(&e1));
    const int threads = 256;
    const int blocks = ctx->sm_count * 4;   // 32 warps / SM
    double per_thread = 0;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {     // rep 0 = warm-up
        ZKR_CUDA(cudaEventRecord(e0, ctx->s[0]));
        switch (which) {
            case 0: ZKR_LAUNCH(ctx, k_bench_imad, blocks, threads, 0, ctx->s[0], dout, iters, 0x9e3779b9u, 12345u);
                per_thread = 32.0 * iters; break;
            case 1: ZKR_LAUNCH(ctx, k_bench_imad_wide, blocks, threads, 0, ctx->s[0], dout, iters, 0x9e3779b9u);
                per_thread = 32.0 * iters; break;
            case 2: ZKR_LAUNCH(ctx, k_bench_modmul, blocks, threads, 0, ctx->s[0], dout, iters);
                per_thread = 2.0 * iters; break;
            case 3: ZKR_LAUNCH(ctx, k_bench_madd, blocks, threads, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters; break;
            case 4: ZKR_LAUNCH(ctx, k_bench_dfma, blocks, threads, 0, ctx->s[0], dout, iters, 1.0000001, 0.5);
                per_thread = 32.0 * iters; break;
            case 5: ZKR_LAUNCH(ctx, k_bench_dfma_imadw, blocks, threads, 0, ctx->s[0], dout, iters, 1.0000001, 0.5, 0x9e3779b9u);
                per_thread = 16.0 * iters; break;          // pairs (1 DFMA + 1 IMAD.WIDE)
            case 6: ZKR_LAUNCH(ctx, k_bench_dfma_iadd, blocks, threads, 0, ctx->s[0], dout, iters, 1.0000001, 0.5);
                per_thread = 16.0 * iters; break;          // pairs (1 DFMA + 1 64-bit add)
            case 7: ZKR_LAUNCH(ctx, k_bench_iadd64, blocks, threads, 0, ctx->s[0], dout, iters, 0x123456789abcdefull);
                per_thread = 32.0 * iters; break;
            case 8: ZKR_LAUNCH(ctx, k_bench_inv_fermat, blocks, threads, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters; break;           // inversions
            case 9: ZKR_LAUNCH(ctx, k_bench_inv_euclid, blocks, threads, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters; break;
            case 10: {                                      // batched affine additions, B = 16 per inversion, 128-thread CTAs
                ZKR_CUDA(cudaFuncSetAttribute(k_bench_batch_affine<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 32 * 128));
                ZKR_LAUNCH(ctx, k_bench_batch_affine<16>, ctx->sm_count * 3, 128, 16 * 32 * 128, ctx->s[0], dout, iters);
                per_thread = 16.0 * iters * (ctx->sm_count * 3.0 * 128) / ((double)threads * blocks); break;
            }
            case 12: ZKR_LAUNCH(ctx, (k_bench_madd_v<Fq, 1>), blocks, threads, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters; break;
            case 15: ZKR_LAUNCH(ctx, (k_bench_madd_v<Fq, 2>), blocks, threads, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters; break;
            case 13: ZKR_LAUNCH(ctx, (k_bench_madd_v<Fq2, 0>), ctx->sm_count * 2, 128, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters * (ctx->sm_count * 2.0 * 128) / ((double)threads * blocks); break;   // 8 warps / SM, as in the MSM
            case 14: ZKR_LAUNCH(ctx, (k_bench_madd_v<Fq2, 1>), ctx->sm_count * 2, 128, 0, ctx->s[0], dout, iters);
                per_thread = 1.0 * iters * (ctx->sm_count * 2.0 * 128) / ((double)threads * blocks); break;
            case 11: {                                      // B = 64 per inversion, 64-thread CTAs
                ZKR_CUDA(cudaFuncSetAttribute(k_bench_batch_affine<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 32 * 64));
                ZKR_LAUNCH(ctx, k_bench_batch_affine<64>, ctx->sm_count, 64, 64 * 32 * 64, ctx->s[0], dout, iters);
                per_thread = 64.0 * iters * (ctx->sm_count * 64.0) / ((double)threads * blocks); break;
            }
            default: break;
        }
        ZKR_CUDA(cudaEventRecord(e1, ctx->s[0]));
        ZKR_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        ZKR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    *ops_per_s = per_thread * threads * blocks / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dout);
    return ZKR_OK;
}
