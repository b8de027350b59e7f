// Peer-memory communicator shared by the sharded MSM / NTT / prove paths (comm.cu).
//
// One zkr_comm per rank (= per GPU).  Every rank owns one device slab
//     [ flags 256 B | gather 2 x kMaxRanks x 1 KB | pad | exchange buffer 0 | exchange buffer 1 ]
// that all peers map (CUDA IPC across processes, plain peer access inside one process), so kernels
// store straight into a peer's HBM over NVLink: the four-step NTT's all-to-all transpose is the
// write-back of the pass before it, an MSM's partial point is written into every peer's gather slot
// by the producing rank.  Cross-rank ordering is a flag barrier kernel (bounded spin, no host round trip).
#pragma once
#include "common.cuh"
#include "ntt_iface.cuh"

struct zkr_comm {
    zkr_ctx* ctx = nullptr;
    int rank = 0, world = 1, g = 0;
    size_t cap_elems = 0;                 // Fr elements per exchange buffer
    char* slab = nullptr;
    size_t slab_bytes = 0;
    char* peer_slab[zkr::kMaxRanks] = {};
    bool ipc_open[zkr::kMaxRanks] = {};
    uint32_t epoch = 0;                   // barriers issued so far
    uint32_t gather_seq = 0;              // gathers issued so far (slot parity)
    int* d_err = nullptr;                 // device flag: a barrier timed out
    char* d_small = nullptr;              // 512 B: partial (256) | affine result (256) of a sharded MSM
    bool connected = false;
    bool dead = false;                    // a barrier timed out: epochs may have diverged, the comm must be re-created
};

namespace zkr {
constexpr size_t kCommFlagsBytes = 256;
constexpr size_t kCommSlotBytes = 1024;
constexpr size_t kCommHeaderBytes = 32 * 1024;

inline char* comm_gather_slot(const zkr_comm* c, int on_rank, int parity, int from_rank) {
    return c->peer_slab[on_rank] + kCommFlagsBytes + ((size_t)parity * kMaxRanks + from_rank) * kCommSlotBytes;
}
inline Fr* comm_xbuf(const zkr_comm* c, int on_rank, int which) {
    return reinterpret_cast<Fr*>(c->peer_slab[on_rank] + kCommHeaderBytes) + (size_t)which * c->cap_elems;
}
// flag barrier across the ranks, ordered on `st`
int comm_barrier(zkr_comm* c, cudaStream_t st);
// copy `bytes` (<= 1 KB, multiple of 16) from local device memory into gather slot [rank] of EVERY rank,
// then barrier.  Returns the parity to read the slots with (comm_gather_slot(c, c->rank, parity, r)).
int comm_allgather_small(zkr_comm* c, cudaStream_t st, const void* d_src, size_t bytes, int* parity_out);
// host-side check of the timeout flag (synchronises st)
int comm_check(zkr_comm* c, cudaStream_t st);
}  // namespace zkr
