// Multi-GPU plumbing used by the distributed CG: NCCL is loaded at run time (dlopen) so that
// single-GPU users do not need it.
#pragma once
#include "common.cuh"

struct ob200_comm {
    ob200_context *ctx = nullptr;
    int nranks = 1, rank = 0;
    void *nccl = nullptr;            // ncclComm_t
    int32_t neq = 0;
    // halo description
    int nneigh = 0;
    std::vector< int > neigh_rank;
    std::vector< int64_t > neigh_offset;      // [nneigh+1] into send/recv buffers
    int64_t nshared = 0;                      // total entries over neighbours
    ob200::DevBuf< int32_t > shared_eq;       // [nshared] local equation (0-based) per buffer entry
    ob200::DevBuf< double > sendbuf, recvbuf;
    // canonical-order accumulation per unique shared dof
    int64_t nuniq = 0;
    ob200::DevBuf< int32_t > uniq_eq, uniq_ptr, uniq_idx, uniq_before;
    ob200::DevBuf< unsigned char > owned;     // [neq]

    // ---- peer-memory transport (comm_p2p.cu): every rank owns a "mailbox" in its HBM that the other
    // ranks of the node map through CUDA IPC and write with plain stores over NVLink
    bool p2p = false;
    int64_t cap = 0;                          // halo entries one rank may send to one neighbour
    char *mailbox = nullptr;                  // cudaMalloc'ed, exported
    std::vector< char * > peer_base;          // [nranks] mapped mailboxes (own entry = mailbox)
    uint32_t halo_seq = 0, scal_seq = 0;      // exchange counters (parity selects the buffer half)
    ob200::DevBuf< int64_t > p_off;           // [nneigh+1] = neigh_offset on the device
    ob200::DevBuf< unsigned long long * > p_data;    // [nneigh] neighbour's halo area reserved for this rank (half 0)
    ob200::DevBuf< unsigned long long * > p_sdata;   // [nranks] scalar entries of this rank in every mailbox (half 0)
    ob200::DevBuf< unsigned long long * > push_dst;  // per uniq_idx entry: where the sharer expects this rank's value (half 0)
    ob200::DevBuf< int32_t > uniq_mb;         // per uniq_idx entry: entry of this rank's mailbox the sharer writes
    ob200::DevBuf< int32_t > route;           // [neq] -1 or index of the shared dof (SpmvHalo)
    ob200::DevBuf< double > pull_partials;    // per-CTA partial sums of the pull kernel's dot product
    int pull_grid = 0;
};

// Mailbox layout, identical on every rank (nranks, cap agreed at export).  An entry is two 8-byte
// words {32 data bits | sequence number << 32}: a word is valid as soon as it carries the expected
// sequence number, so neither fences nor separate flags are needed (8-byte stores are atomic).
struct ob200_mailbox_layout {
    int64_t data, sdata, error, bytes;        // byte offsets
    int64_t data_half, sdata_half;            // entries per buffer half
};
constexpr int kScalSlots = 4;
inline ob200_mailbox_layout mailbox_layout(int nranks, int64_t cap)
{
    ob200_mailbox_layout L;
    auto up = [](int64_t v) { return ( v + 255 ) & ~(int64_t) 255; };
    L.data = 0;
    L.data_half = (int64_t) nranks * cap;
    L.sdata = up(2 * L.data_half * 16);
    L.sdata_half = (int64_t) nranks * kScalSlots;
    L.error = up(L.sdata + 2 * L.sdata_half * 16);
    L.bytes = up(L.error + 16);
    return L;
}

namespace ob200 {
int comm_allreduce_sum(ob200_comm *c, double *dev, int n);
int comm_exchange_add(ob200_comm *c, double *y, const int *done = nullptr);
int comm_p2p_prepare(ob200_comm *c);                               // after set_halo / p2p_open
int comm_p2p_exchange_add(ob200_comm *c, double *y, const int *done);
// second half of an exchange whose first half (the push) the SpMV epilogue has done; optionally the
// partial sums of dot_p . y over the shared dofs this rank owns (-> c->pull_partials[0 .. c->pull_grid))
int comm_p2p_push(ob200_comm *c, const double *y, unsigned int seq, const int *done);
int comm_p2p_pull(ob200_comm *c, double *y, unsigned int seq, const double *dot_p, const int *done);
int comm_p2p_check(ob200_comm *c);                                  // negative if a peer wait timed out
}
