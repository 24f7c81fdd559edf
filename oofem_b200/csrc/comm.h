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
};

namespace ob200 {
int comm_allreduce_sum(ob200_comm *c, double *dev, int n);
int comm_exchange_add(ob200_comm *c, double *y);
}
