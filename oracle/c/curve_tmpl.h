/* Jacobian short-Weierstrass (a = 0) arithmetic, instantiated twice by zkr_oracle.c (G1 over Fq, G2
 * over Fq2) through the F_* macros.  TEST INFRASTRUCTURE ONLY (see zkr_oracle.c header).
 * Formulas: snarkjs@0.1.20 src/gcurve.js structure (double / add / mulScalar by double-and-add),
 * restated from the published EFD forms dbl-2009-l, add-2007-bl, madd-2007-bl.
 * Required macros: F_T, F_MUL(r,a,b), F_SQR(r,a), F_ADD, F_SUB, F_NEG(r,a), F_ISZERO(a), F_EQ(a,b),
 * F_SET_ONE(r), F_SET_ZERO(r), F_INV(r,a), NAME(x) */

typedef struct { F_T x, y; int inf; } NAME(aff);
typedef struct { F_T x, y, z; } NAME(jac);          /* z == 0 <=> infinity */

static inline void NAME(jset_inf)(NAME(jac)* p) { F_SET_ZERO(&p->x); F_SET_ONE(&p->y); F_SET_ZERO(&p->z); }
static inline int NAME(jis_inf)(const NAME(jac)* p) { return F_ISZERO(&p->z); }

static void NAME(jdbl)(NAME(jac)* r, const NAME(jac)* p) {
    if (NAME(jis_inf)(p) || F_ISZERO(&p->y)) { NAME(jset_inf)(r); return; }
    F_T A, B, C, D, E, Fv, t, x3, y3, z3;
    F_SQR(&A, &p->x); F_SQR(&B, &p->y); F_SQR(&C, &B);
    F_ADD(&t, &p->x, &B); F_SQR(&t, &t); F_SUB(&t, &t, &A); F_SUB(&t, &t, &C); F_ADD(&D, &t, &t);
    F_ADD(&E, &A, &A); F_ADD(&E, &E, &A);
    F_SQR(&Fv, &E);
    F_SUB(&x3, &Fv, &D); F_SUB(&x3, &x3, &D);
    F_SUB(&t, &D, &x3); F_MUL(&y3, &E, &t);
    F_ADD(&C, &C, &C); F_ADD(&C, &C, &C); F_ADD(&C, &C, &C);
    F_SUB(&y3, &y3, &C);
    F_MUL(&z3, &p->y, &p->z); F_ADD(&z3, &z3, &z3);
    r->x = x3; r->y = y3; r->z = z3;
}

static void NAME(jadd)(NAME(jac)* r, const NAME(jac)* p, const NAME(jac)* q) {
    if (NAME(jis_inf)(p)) { *r = *q; return; }
    if (NAME(jis_inf)(q)) { *r = *p; return; }
    F_T z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
    F_SQR(&z1z1, &p->z); F_SQR(&z2z2, &q->z);
    F_MUL(&u1, &p->x, &z2z2); F_MUL(&u2, &q->x, &z1z1);
    F_MUL(&t, &q->z, &z2z2); F_MUL(&s1, &p->y, &t);
    F_MUL(&t, &p->z, &z1z1); F_MUL(&s2, &q->y, &t);
    if (F_EQ(&u1, &u2)) {
        if (F_EQ(&s1, &s2)) { NAME(jdbl)(r, p); return; }
        NAME(jset_inf)(r); return;
    }
    F_SUB(&h, &u2, &u1); F_SUB(&rr, &s2, &s1);
    F_SQR(&hh, &h); F_MUL(&hhh, &h, &hh); F_MUL(&v, &u1, &hh);
    F_SQR(&x3, &rr); F_SUB(&x3, &x3, &hhh); F_SUB(&x3, &x3, &v); F_SUB(&x3, &x3, &v);
    F_SUB(&t, &v, &x3); F_MUL(&y3, &rr, &t); F_MUL(&t, &s1, &hhh); F_SUB(&y3, &y3, &t);
    F_MUL(&z3, &p->z, &q->z); F_MUL(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

/* r = p + q, q affine (not infinity) */
static void NAME(jmadd)(NAME(jac)* r, const NAME(jac)* p, const NAME(aff)* q) {
    if (q->inf) { *r = *p; return; }
    if (NAME(jis_inf)(p)) { r->x = q->x; r->y = q->y; F_SET_ONE(&r->z); return; }
    F_T z1z1, u2, s2, h, rr, hh, hhh, v, t, x3, y3, z3;
    F_SQR(&z1z1, &p->z);
    F_MUL(&u2, &q->x, &z1z1);
    F_MUL(&t, &p->z, &z1z1); F_MUL(&s2, &q->y, &t);
    if (F_EQ(&p->x, &u2)) {
        if (F_EQ(&p->y, &s2)) { NAME(jdbl)(r, p); return; }
        NAME(jset_inf)(r); return;
    }
    F_SUB(&h, &u2, &p->x); F_SUB(&rr, &s2, &p->y);
    F_SQR(&hh, &h); F_MUL(&hhh, &h, &hh); F_MUL(&v, &p->x, &hh);
    F_SQR(&x3, &rr); F_SUB(&x3, &x3, &hhh); F_SUB(&x3, &x3, &v); F_SUB(&x3, &x3, &v);
    F_SUB(&t, &v, &x3); F_MUL(&y3, &rr, &t); F_MUL(&t, &p->y, &hhh); F_SUB(&y3, &y3, &t);
    F_MUL(&z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

static void NAME(jneg)(NAME(jac)* r, const NAME(jac)* p) { r->x = p->x; F_NEG(&r->y, &p->y); r->z = p->z; }

/* k: 4 x u64 little-endian standard-form scalar; MSB-first double-and-add (snarkjs mulScalar) */
static void NAME(jmul)(NAME(jac)* r, const NAME(jac)* p, const uint64_t k[4]) {
    NAME(jac) acc; NAME(jset_inf)(&acc);
    int top = 255;
    while (top >= 0 && !((k[top >> 6] >> (top & 63)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        NAME(jdbl)(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) NAME(jadd)(&acc, &acc, p);
    }
    *r = acc;
}

static void NAME(to_affine)(NAME(aff)* r, const NAME(jac)* p) {
    if (NAME(jis_inf)(p)) { r->inf = 1; F_SET_ZERO(&r->x); F_SET_ZERO(&r->y); return; }
    F_T zi, zi2, zi3;
    F_INV(&zi, &p->z); F_SQR(&zi2, &zi); F_MUL(&zi3, &zi2, &zi);
    F_MUL(&r->x, &p->x, &zi2); F_MUL(&r->y, &p->y, &zi3); r->inf = 0;
}

/* out[i] = base + i * step (affine, i < n): a chain of mixed additions, then ONE shared inversion for the n
 * conversions to affine (Montgomery's trick).  Host-side workload generator for the arithmetic-progression MSM
 * check of SURVEY.md 8(d) config 3 (tests only). */
static void NAME(ap_points)(NAME(aff)* out, const NAME(aff)* base, const NAME(aff)* step, size_t n) {
    if (n == 0) return;
    NAME(jac)* j = (NAME(jac)*)malloc(sizeof(NAME(jac)) * n);
    F_T* pre = (F_T*)malloc(sizeof(F_T) * n);
    j[0].x = base->x; j[0].y = base->y; F_SET_ONE(&j[0].z);
    if (base->inf) NAME(jset_inf)(&j[0]);
    for (size_t i = 1; i < n; i++) NAME(jmadd)(&j[i], &j[i - 1], step);
    F_T acc; F_SET_ONE(&acc);
    for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!NAME(jis_inf)(&j[i])) F_MUL(&acc, &acc, &j[i].z); }
    F_T inv; F_INV(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (NAME(jis_inf)(&j[i])) { out[i].inf = 1; F_SET_ZERO(&out[i].x); F_SET_ZERO(&out[i].y); continue; }
        F_T zi, zi2, zi3;
        F_MUL(&zi, &inv, &pre[i]); F_MUL(&inv, &inv, &j[i].z);
        F_SQR(&zi2, &zi); F_MUL(&zi3, &zi2, &zi);
        F_MUL(&out[i].x, &j[i].x, &zi2); F_MUL(&out[i].y, &j[i].y, &zi3); out[i].inf = 0;
    }
    free(pre); free(j);
}

/* ---- multi-scalar multiplication -------------------------------------------------------------
 * mode 0: the snarkjs genProof structure -- one double-and-add per point, summed (prover_groth.js).
 * mode 1: Pippenger buckets with unsigned c-bit windows; windows are spread over `threads`
 *         pthreads ("strong CPU" baseline).                                       */
typedef struct {
    const NAME(aff)* pts; const uint64_t (*sc)[4]; size_t n; int c, nwin, w0, wstep; NAME(jac)* win_out;
} NAME(pip_job);

static inline unsigned NAME(get_bits)(const uint64_t k[4], int lo, int c) {
    int limb = lo >> 6, sh = lo & 63;
    if (limb >= 4) return 0;
    uint64_t v = k[limb] >> sh;
    if (sh + c > 64 && limb + 1 < 4) v |= k[limb + 1] << (64 - sh);
    return (unsigned)(v & ((1ull << c) - 1));
}

static void* NAME(pip_worker)(void* arg) {
    NAME(pip_job)* j = (NAME(pip_job)*)arg;
    size_t nb = ((size_t)1 << j->c) - 1;
    NAME(jac)* bk = (NAME(jac)*)malloc(sizeof(NAME(jac)) * nb);
    for (int w = j->w0; w < j->nwin; w += j->wstep) {
        for (size_t b = 0; b < nb; b++) NAME(jset_inf)(&bk[b]);
        for (size_t i = 0; i < j->n; i++) {
            if (j->pts[i].inf) continue;
            unsigned d = NAME(get_bits)(j->sc[i], w * j->c, j->c);
            if (d) NAME(jmadd)(&bk[d - 1], &bk[d - 1], &j->pts[i]);
        }
        NAME(jac) run, acc; NAME(jset_inf)(&run); NAME(jset_inf)(&acc);
        for (size_t b = nb; b-- > 0;) { NAME(jadd)(&run, &run, &bk[b]); NAME(jadd)(&acc, &acc, &run); }
        j->win_out[w] = acc;
    }
    free(bk);
    return NULL;
}

static void NAME(msm)(NAME(jac)* out, const NAME(aff)* pts, const uint64_t (*sc)[4], size_t n, int mode, int threads) {
    NAME(jset_inf)(out);
    if (n == 0) return;
    if (mode == 0) {
        for (size_t i = 0; i < n; i++) {
            if (pts[i].inf) continue;
            NAME(jac) p, t; p.x = pts[i].x; p.y = pts[i].y; F_SET_ONE(&p.z);
            NAME(jmul)(&t, &p, sc[i]);
            NAME(jadd)(out, out, &t);
        }
        return;
    }
    int c = 4;
    { size_t t = n; while (t > 32 && c < 16) { t >>= 1; c++; } if (c > 4) c -= 2; if (c < 4) c = 4; }
    int nwin = (254 + c - 1) / c;
    NAME(jac)* wins = (NAME(jac)*)malloc(sizeof(NAME(jac)) * nwin);
    if (threads < 1) threads = 1;
    if (threads > nwin) threads = nwin;
    pthread_t th[64]; NAME(pip_job) jobs[64];
    if (threads > 64) threads = 64;
    for (int t = 0; t < threads; t++) {
        jobs[t] = (NAME(pip_job)){pts, sc, n, c, nwin, t, threads, wins};
        pthread_create(&th[t], NULL, NAME(pip_worker), &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    for (int w = nwin - 1; w >= 0; w--) {
        for (int d = 0; d < c; d++) NAME(jdbl)(out, out);
        NAME(jadd)(out, out, &wins[w]);
    }
    free(wins);
}
