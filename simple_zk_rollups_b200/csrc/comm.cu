// Multi-GPU paths over peer memory (C-ABI: zkr_comm_*, zkr_msm_sharded, zkr_ntt_sharded).
//
// North star (BASELINE.json): "the MSMs shard naturally by point range, and per-GPU partial sums are
// gathered and added over NVLink; large NTTs use a four-step decomposition with an NVLink all-to-all
// transpose".  Inside websnark these are the g1_multiexp / g2_multiexp / fft calls of groth16GenProof
// (/root/reference/operator/src/snarks/common.ts:29), which fan out to web workers on one host; here
// the fan-out is across the GPUs of one NVSwitch box, one process (or one context) per GPU.
//
// No NCCL on the data path: peers' slabs are mapped (CUDA IPC / peer access) and written by the
// producing kernels themselves -- the transpose of the four-step NTT is the write-back of the pass
// before it (ntt.cu, k_ntt_pass MODE 1 / 2), so the transfer overlaps the butterflies tile by tile.
#include "comm_iface.cuh"
#include "ec.cuh"
#include "msm_iface.cuh"

using namespace zkr;

namespace {

struct PeerPtrs {
    char* p[kMaxRanks];
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// thread q: publish `epoch` in rank q's flag word [rank], then wait for rank q's word in my flags.
// Kernel boundaries order the preceding kernels' remote stores before the flag store.
__global__ void k_comm_barrier(PeerPtrs slabs, int rank, int world, uint32_t epoch, int* err) {
    const int q = threadIdx.x;
    if (q >= world) return;
    __threadfence_system();
    volatile uint32_t* remote = reinterpret_cast<volatile uint32_t*>(slabs.p[q]) + rank;
    *remote = epoch;
    __threadfence_system();
    volatile uint32_t* mine = reinterpret_cast<volatile uint32_t*>(slabs.p[rank]) + q;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(*mine - epoch) < 0) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > 20000000000ull) {   // 20 s: a peer died; fail instead of hanging the GPU
            *err = 1;
            break;
        }
    }
    __threadfence_system();
}

// block r: copy `bytes` from src into slot `dst_off` of rank r's slab
__global__ void k_push_small(PeerPtrs slabs, size_t dst_off, const uint4* __restrict__ src, int n16) {
    uint4* dst = reinterpret_cast<uint4*>(slabs.p[blockIdx.x] + dst_off);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

// out = sum of `world` XYZZ points at stride kCommSlotBytes, converted to affine standard form
template <class F>
__global__ void k_sum_to_affine_std(const char* slots, int world, char* out) {
    XYZZ<F> acc = XYZZ<F>::load(slots);
#pragma unroll 1
    for (int r = 1; r < world; r++) acc.add(XYZZ<F>::load(slots + (size_t)r * kCommSlotBytes));
    Affine<F> a = acc.to_affine_vartime();   // single thread (launched <<<1, 1>>>)
    a.x.from_mont().store(out);
    a.y.from_mont().store(out + sizeof(F));
}

PeerPtrs peer_ptrs(const zkr_comm* c) {
    PeerPtrs p = {};
    for (int r = 0; r < c->world; r++) p.p[r] = c->peer_slab[r];
    return p;
}

int barrier_cb(void* arg, cudaStream_t st) { return comm_barrier((zkr_comm*)arg, st); }

}  // namespace

namespace zkr {

int comm_barrier(zkr_comm* c, cudaStream_t st) {
    if (c->world == 1) return ZKR_OK;
    c->epoch++;
    ZKR_LAUNCH(c->ctx, k_comm_barrier, 1, 32, 0, st, peer_ptrs(c), c->rank, c->world, c->epoch, c->d_err);
    return ZKR_OK;
}

int comm_allgather_small(zkr_comm* c, cudaStream_t st, const void* d_src, size_t bytes, int* parity_out) {
    if (bytes > kCommSlotBytes || (bytes & 15)) return ZKR_E_INVALID;
    const int parity = (int)(c->gather_seq++ & 1);
    const size_t off = kCommFlagsBytes + ((size_t)parity * kMaxRanks + c->rank) * kCommSlotBytes;
    ZKR_LAUNCH(c->ctx, k_push_small, c->world, 64, 0, st, peer_ptrs(c), off, (const uint4*)d_src, (int)(bytes / 16));
    ZKR_TRY(comm_barrier(c, st));
    *parity_out = parity;
    return ZKR_OK;
}

int comm_check(zkr_comm* c, cudaStream_t st) {
    int e = 0;
    ZKR_CUDA(cudaMemcpyAsync(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    ZKR_CUDA(cudaStreamSynchronize(st));
    if (e) {
        // report once, then clear (like the witness range flags): the flag says "a barrier since the last check timed
        // out".  The ranks' barrier epochs may have diverged at that point -- a rank that bailed out never launched its
        // barrier -- so the communicator must be destroyed and re-created on EVERY rank before it is used again.
        ZKR_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), st));
        c->dead = true;
        set_error("rank %d: a peer did not reach the barrier within 20 s; destroy and re-create the communicator on every rank", c->rank);
        return ZKR_E_NCCL;
    }
    if (c->dead) {
        set_error("rank %d: this communicator saw a barrier timeout earlier; destroy and re-create it on every rank", c->rank);
        return ZKR_E_NCCL;
    }
    return ZKR_OK;
}

}  // namespace zkr

extern "C" int zkr_comm_create(zkr_ctx* ctx, int rank, int world, size_t max_elems_per_rank, zkr_comm** out) {
    if (!ctx || !out || world < 1 || world > kMaxRanks || (world & (world - 1)) || rank < 0 || rank >= world) {
        set_error("zkr_comm_create: world must be 1, 2, 4 or 8 and 0 <= rank < world");
        return ZKR_E_INVALID;
    }
    *out = nullptr;
    DeviceGuard g(ctx->device);
    zkr_comm* c = new zkr_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    while ((1 << c->g) < world) c->g++;
    c->cap_elems = (max_elems_per_rank + 7) & ~size_t(7);
    c->slab_bytes = kCommHeaderBytes + 2 * c->cap_elems * sizeof(Fr);
    cudaError_t e = cudaMalloc(&c->slab, c->slab_bytes);
    if (e != cudaSuccess) {
        delete c;
        return cuda_fail(e, "cudaMalloc(comm slab)", __FILE__, __LINE__);
    }
    if ((e = cudaMalloc(&c->d_err, sizeof(int))) != cudaSuccess || (e = cudaMalloc(&c->d_small, 512)) != cudaSuccess ||
        (e = cudaMemset(c->slab, 0, kCommHeaderBytes)) != cudaSuccess || (e = cudaMemset(c->d_err, 0, sizeof(int))) != cudaSuccess) {
        cudaFree(c->slab);
        cudaFree(c->d_err);
        cudaFree(c->d_small);
        delete c;
        return cuda_fail(e, "zkr_comm_create (flag / staging buffers)", __FILE__, __LINE__);
    }
    c->peer_slab[rank] = c->slab;
    c->connected = world == 1;
    ZKR_CUDA(cudaDeviceSynchronize());
    *out = c;
    return ZKR_OK;
}

extern "C" int zkr_comm_export(const zkr_comm* c, void* handle64) {
    if (!c || !handle64) return ZKR_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == ZKR_IPC_HANDLE_BYTES, "handle size");
    DeviceGuard g(c->ctx->device);
    cudaIpcMemHandle_t h;
    ZKR_CUDA(cudaIpcGetMemHandle(&h, c->slab));
    memcpy(handle64, &h, sizeof(h));
    return ZKR_OK;
}

extern "C" int zkr_comm_connect(zkr_comm* c, const void* handles) {
    if (!c || !handles) return ZKR_E_INVALID;
    DeviceGuard g(c->ctx->device);
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * ZKR_IPC_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        ZKR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_slab[r] = (char*)p;
        c->ipc_open[r] = true;
    }
    c->connected = true;
    return ZKR_OK;
}

extern "C" int zkr_comm_connect_local(zkr_comm* const* comms, int world) {
    if (!comms || world < 1 || world > kMaxRanks) return ZKR_E_INVALID;
    for (int a = 0; a < world; a++)
        if (!comms[a] || comms[a]->world != world || comms[a]->rank != a) return ZKR_E_INVALID;
    for (int a = 0; a < world; a++)
        for (int b = a + 1; b < world; b++)
            if (comms[a]->ctx == comms[b]->ctx ||
                (comms[a]->ctx->device == comms[b]->ctx->device && comms[a]->ctx->user_stream == comms[b]->ctx->user_stream)) {
                set_error("ranks %d and %d share a device and a stream: their barrier kernels would wait on each other "
                          "(give each ctx its own stream with zkr_ctx_set_stream)", a, b);
                return ZKR_E_INVALID;
            }
    for (int a = 0; a < world; a++) {
        zkr_comm* c = comms[a];
        DeviceGuard g(c->ctx->device);
        for (int b = 0; b < world; b++) {
            if (a == b) continue;
            const int db = comms[b]->ctx->device;
            if (db != c->ctx->device) {
                int can = 0;
                ZKR_CUDA(cudaDeviceCanAccessPeer(&can, c->ctx->device, db));
                if (!can) {
                    set_error("device %d cannot access device %d", c->ctx->device, db);
                    return ZKR_E_UNSUPPORTED;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            }
            c->peer_slab[b] = comms[b]->slab;
        }
        c->connected = true;
    }
    return ZKR_OK;
}

extern "C" int zkr_comm_barrier(zkr_comm* c) {
    if (!c || !c->connected) return ZKR_E_INVALID;
    DeviceGuard g(c->ctx->device);
    return comm_barrier(c, c->ctx->user_stream);
}

extern "C" void* zkr_comm_buffer(zkr_comm* c, int which) {
    if (!c || which < 0 || which > 1) return nullptr;
    return comm_xbuf(c, c->rank, which);
}

extern "C" int zkr_comm_info(const zkr_comm* c, int* rank, int* world, uint64_t* elems_per_buffer) {
    if (!c) return ZKR_E_INVALID;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (elems_per_buffer) *elems_per_buffer = c->cap_elems;
    return ZKR_OK;
}

extern "C" void zkr_comm_destroy(zkr_comm* c) {
    if (!c) return;
    DeviceGuard g(c->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; r++)
        if (c->ipc_open[r]) cudaIpcCloseMemHandle(c->peer_slab[r]);
    cudaFree(c->slab);
    cudaFree(c->d_err);
    cudaFree(c->d_small);
    delete c;
}

// ---------------------------------------------------------------------------------------------- MSM
extern "C" int zkr_msm_sharded(zkr_comm* c, const zkr_bases* b, const void* scalars, size_t n_local,
                               int scalars_on_device, void* out_affine) {
    if (c && c->dead) return comm_check(c, c->ctx->user_stream);
    if (!c || !c->connected || !b || !out_affine || (!scalars && n_local) || bases_ctx(b) != c->ctx ||
        n_local != bases_n_src(b)) {
        set_error("zkr_msm_sharded: bad arguments (scalar count must equal this rank's loaded points)");
        return ZKR_E_INVALID;
    }
    zkr_ctx* ctx = c->ctx;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->user_stream;
    const uint32_t* d_sc = (const uint32_t*)scalars;
    if (!scalars_on_device && n_local) {
        void* p;
        ZKR_TRY(ctx->scratch_get("msm_scalars", n_local * 32 + 32, &p));
        ZKR_CUDA(cudaMemcpyAsync(p, scalars, n_local * 32, cudaMemcpyHostToDevice, st));
        d_sc = (const uint32_t*)p;
    }
    const int group = bases_group(b);
    const size_t xb = group == 1 ? 128 : 256, ob = group == 1 ? 64 : 128;
    void* d_part = c->d_small;
    void* d_aff = c->d_small + 256;
    if (group == 1) ZKR_TRY(msm_run_g1(ctx, st, b, d_sc, d_part));
    else ZKR_TRY(msm_run_g2(ctx, st, b, d_sc, d_part));
    int parity = 0;
    ZKR_TRY(comm_allgather_small(c, st, d_part, xb, &parity));
    const char* slots = comm_gather_slot(c, c->rank, parity, 0);
    if (group == 1) ZKR_LAUNCH(ctx, k_sum_to_affine_std<Fq>, 1, 1, 0, st, slots, c->world, (char*)d_aff);
    else ZKR_LAUNCH(ctx, k_sum_to_affine_std<Fq2>, 1, 1, 0, st, slots, c->world, (char*)d_aff);
    ZKR_CUDA(cudaMemcpyAsync(out_affine, d_aff, ob, cudaMemcpyDeviceToHost, st));
    int err = 0;
    ZKR_TRY(bases_range_error(b, st, &err));
    ZKR_TRY(comm_check(c, st));
    if (err) {
        set_error("a scalar is >= r");
        return ZKR_E_WITNESS_RANGE;
    }
    return ZKR_OK;
}

// ---------------------------------------------------------------------------------------------- NTT
extern "C" int zkr_ntt_sharded_rows_log(int log_n, int world) {
    int g = 0;
    while ((1 << g) < world) g++;
    return ntt_sharded_k0(log_n, g);
}

extern "C" int zkr_ntt_sharded(zkr_comm* c, int log_n, int mode, int src_buf) {
    if (!c || !c->connected || src_buf < 0 || src_buf > 1) return ZKR_E_INVALID;
    if (c->dead) return comm_check(c, c->ctx->user_stream);
    const int base_mode = mode & 0xf;
    const bool br_out = mode & ZKR_NTT_BITREV_OUT, br_in = mode & ZKR_NTT_BITREV_IN;
    if (base_mode > 3 || br_out == br_in) {
        set_error("zkr_ntt_sharded: exactly one of ZKR_NTT_BITREV_OUT (COLS -> ROWS) / ZKR_NTT_BITREV_IN (ROWS -> COLS) is required");
        return ZKR_E_INVALID;
    }
    if (log_n < c->g || ((size_t)1 << (log_n - c->g)) > c->cap_elems) {
        set_error("zkr_ntt_sharded: 2^%d / %d ranks exceeds the communicator's %zu elements per rank", log_n, c->world, c->cap_elems);
        return ZKR_E_INVALID;
    }
    zkr_ctx* ctx = c->ctx;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->user_stream;
    NttTables* t;
    ZKR_TRY(ntt_get_tables(ctx, log_n, &t));
    NttXchg x = {};
    x.g = c->g;
    x.rank = c->rank;
    x.k0 = ntt_sharded_k0(log_n, c->g);
    x.s0 = log_n - x.k0;
    for (int r = 0; r < c->world; r++) x.peer[r] = comm_xbuf(c, r, 1 - src_buf);
    Fr* src = comm_xbuf(c, c->rank, src_buf);
    Fr* dst = comm_xbuf(c, c->rank, 1 - src_buf);
    const bool inverse = base_mode == ZKR_NTT_INVERSE || base_mode == ZKR_NTT_COSET_INVERSE;
    const bool coset = base_mode >= ZKR_NTT_COSET_FORWARD;
    const int in_layout = br_in ? 1 : 0, out_layout = 1 - in_layout;
    const size_t nl = (size_t)1 << (log_n - c->g);
    if (coset && !inverse) ZKR_TRY(ntt_scale_pow_sharded(ctx, st, src, x, log_n, in_layout, t->cs_lo, t->cs_hi, t->lb));
    ZKR_TRY(ntt_run_sharded(ctx, st, src, x, log_n, br_in, inverse, barrier_cb, c));
    if (inverse && coset) ZKR_TRY(ntt_scale_pow_sharded(ctx, st, dst, x, log_n, out_layout, t->ci_lo, t->ci_hi_n, t->lb));
    else if (inverse) ZKR_TRY(ntt_scale_const(ctx, st, dst, nl, &t->roots->ninv));
    return ZKR_OK;
}

extern "C" int zkr_comm_check(zkr_comm* c) {
    if (!c) return ZKR_E_INVALID;
    DeviceGuard g(c->ctx->device);
    return comm_check(c, c->ctx->user_stream);
}
