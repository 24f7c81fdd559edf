// Shared between the SpMV kernel (spmv.cuh), the peer-memory transport (comm_p2p.cu) and the CG (cg.cu).
#pragma once
#include <stdint.h>

namespace ob200 {

// What the distributed SpMV does with the rows shared with other partitions (comm_p2p.cu): route[r] = -1
// for a row only this rank holds, else the index u of the shared dof; the row's local sum is then
// written straight into the mailboxes of the ranks sharing it (dst[uptr[u] .. uptr[u+1])) as two
// self-validating {32 data bits, sequence number} words -- no pack kernel, no fence, no flag.
struct SpmvHalo {
    const int32_t *route;
    const int32_t *uptr;
    unsigned long long *const *dst;        // mailbox entry (2 words) per (shared dof, sharer), buffer half 0
    int64_t half_words;                    // offset of the buffer half in use, in 8-byte words
    unsigned int seq;
};

__device__ __forceinline__ void ll_store(unsigned long long *slot, double v, unsigned int seq)
{
    const unsigned long long b = (unsigned long long) __double_as_longlong(v), f = (unsigned long long) seq << 32;
    asm volatile( "st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"( slot ), "l"( f | ( b & 0xffffffffull ) ), "l"( f | ( b >> 32 ) ) : "memory" );
}

} // namespace ob200
