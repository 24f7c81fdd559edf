// SparseLinearSystemNM "cudacg": the IML++ preconditioned CG template (iml/cg.h:23-72) as
// IMLSolver::solve instantiates it (src/core/iml/imlsolver.C:101-146) with DiagPreconditioner
// (src/core/iml/diagpre.C) or VoidPreconditioner.
//
// All scalars (rho, alpha, beta, residual, the convergence flag) stay on the device; kernels
// become no-ops once the flag is set, so the host enqueues iterations ahead and polls the flag
// every few iterations -- the result is identical to stopping exactly at the reference's test
//   if ((resid = norm(r) / normb) <= tol) return   (cg.h:36, cg.h:60).
// Reductions are two-stage (per-block partials in a fixed order, then one block) and therefore
// deterministic run to run.
#include "common.cuh"
#include "elemset.h"
#include "comm.h"
#include "spmv_halo.h"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace ob200 {

int spmv(ob200_csr *A, const double *x, double *y);
int spmv_fused_dot(ob200_csr *A, const double *x, double *y, double *partials, int *nblocks, const int *done);
bool spmv_halo_supported(const ob200_csr *A);
int spmv_fused_halo(ob200_csr *A, const double *x, double *y, double *partials, int *nblocks, const int *done, const SpmvHalo &halo);

struct CgScalars {
    double normb, rho, rho_1, alpha, beta, resid;
    int iters, done;
    unsigned int ticket;
    int pad;
};

constexpr int kRedMax = 3;            // simultaneous reductions
constexpr int kCgThreads = 256;
constexpr int kCgMaxBlocks = 2048;    // upper bound on the partial sums of one reduction

__device__ __forceinline__ double warp_sum(double s)
{
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 ) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// block-level sum; result valid in thread 0
__device__ __forceinline__ double block_sum(double s, double *scratch /* [32] */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = warp_sum(s);
    __syncthreads();
    if ( lane == 0 ) scratch[wid] = s;
    __syncthreads();
    if ( wid == 0 ) {
        s = lane < ( blockDim.x >> 5 ) ? scratch[lane] : 0.0;
        s = warp_sum(s);
    }
    return s;
}

// "last block done": true in every thread of the CTA that arrives last at the ticket.  The last
// CTA then sums the per-block partials in a fixed order -- deterministic, and no extra launch.
__device__ __forceinline__ bool last_block(unsigned int *ticket)
{
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if ( threadIdx.x == 0 ) {
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = ( t == gridDim.x - 1 );
        if ( is_last ) *ticket = 0;
    }
    __syncthreads();
    return is_last != 0;
}

__device__ __forceinline__ double sum_partials(const double *partials, int P, double *scratch)
{
    double s = 0.0;
    for ( int i = threadIdx.x; i < P; i += blockDim.x ) s += __ldcg(partials + i);
    return block_sum(s, scratch);
}

// the scalar recurrences of iml/cg.h; one thread
// stage 0 (init):  red = {b.b, r.r, r.z};  stage 1 (after SpMV): red = {p.q};
// stage 2 (after x/r update of iteration `iter`): red = {r.r, r.z}
__device__ __forceinline__ void cg_scalars(int stage, int iter, double tol, const double *red, CgScalars *S)
{
    if ( stage == 1 ) {
        S->alpha = S->rho / red[0];                       // alpha = rho / dot(p, q)     (cg.h:54)
        return;
    }
    double rr, rz;
    if ( stage == 0 ) {
        double normb = sqrt(red[0]);                      // Real normb = norm(b)        (cg.h:31)
        if ( normb == 0.0 ) normb = 1;                    //                              (cg.h:34-35)
        S->normb = normb;
        rr = red[1];
        rz = red[2];
    } else {
        rr = red[0];
        rz = red[1];
    }
    double resid = sqrt(rr) / S->normb;
    S->resid = resid;
    if ( resid <= tol ) {                                 // (cg.h:37-41, 60-64)
        S->done = 1;
        S->iters = iter;
        return;
    }
    S->rho_1 = S->rho;                                    // rho_1 = rho                  (cg.h:66)
    S->rho = rz;                                          // rho = dot(r, z)              (cg.h:47)
    S->beta = S->rho / S->rho_1;                          // beta = rho / rho_1           (cg.h:51)
    S->iters = iter;
}

// DiagPreconditioner::init (diagpre.C:41-58): diag = 1 / A(i,i); zero diagonal is an error
__global__ void diag_init_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                 const double *__restrict__ val, double *__restrict__ diag, int *__restrict__ bad)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        int lo = rowptr[i], hi = rowptr[i + 1] - 1;
        double d = 0.0;
        while ( lo <= hi ) {
            int mid = ( lo + hi ) >> 1;
            int v = colind[mid];
            if ( v == i ) { d = val[mid]; break; }
            if ( v < i ) lo = mid + 1; else hi = mid - 1;
        }
        if ( d == 0.0 ) atomicAdd(bad, 1);
        diag[i] = 1.0 / d;
    }
}

// sub-assembled (distributed) diagonal: raw A(i,i), to be summed over partitions before inversion
__global__ void diag_extract_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                    const double *__restrict__ val, double *__restrict__ diag)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        int lo = rowptr[i], hi = rowptr[i + 1] - 1;
        double d = 0.0;
        while ( lo <= hi ) {
            int mid = ( lo + hi ) >> 1;
            int v = colind[mid];
            if ( v == i ) { d = val[mid]; break; }
            if ( v < i ) lo = mid + 1; else hi = mid - 1;
        }
        diag[i] = d;
    }
}

__global__ void diag_invert_kernel(int32_t neq, double *__restrict__ diag, int *__restrict__ bad)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        double d = diag[i];
        if ( d == 0.0 ) atomicAdd(bad, 1);
        diag[i] = 1.0 / d;
    }
}

// r = b - q (q = A x already, halo-summed), partials of b.b, r.r, r.z with z = M^-1 r   (cg.h:32-33)
// FINAL: the last CTA also reduces the partials and runs the stage-0 scalar update.
template< bool FINAL >
__global__ void __launch_bounds__(kCgThreads)
cg_init_kernel(int32_t neq, const double *__restrict__ b, const double *__restrict__ q, double *__restrict__ r,
               const double *__restrict__ diag, const unsigned char *__restrict__ owned,
               double *__restrict__ partials, int P, double *__restrict__ red, double tol, CgScalars *S)
{
    __shared__ double scratch[32];
    double bb = 0.0, rr = 0.0, rz = 0.0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        double bi = b[i], ri = bi - q[i];
        r[i] = ri;
        if ( !owned || owned[i] ) {
            bb += bi * bi;
            rr += ri * ri;
            rz += ri * ( diag ? ri * diag[i] : ri );
        }
    }
    bb = block_sum(bb, scratch);
    if ( threadIdx.x == 0 ) partials[blockIdx.x] = bb;
    rr = block_sum(rr, scratch);
    if ( threadIdx.x == 0 ) partials[P + blockIdx.x] = rr;
    rz = block_sum(rz, scratch);
    if ( threadIdx.x == 0 ) partials[2 * P + blockIdx.x] = rz;
    if ( last_block(&S->ticket) ) {
        double s0 = sum_partials(partials, gridDim.x, scratch);
        double s1 = sum_partials(partials + P, gridDim.x, scratch);
        double s2 = sum_partials(partials + 2 * P, gridDim.x, scratch);
        if ( threadIdx.x == 0 ) {
            red[0] = s0; red[1] = s1; red[2] = s2;
            if ( FINAL ) cg_scalars(0, 0, tol, red, S);
        }
    }
}

// p = z (first iteration) or p = z + beta p, z = M.solve(r)      (cg.h:45-52, diagpre.C:61-68)
__global__ void __launch_bounds__(kCgThreads)
cg_update_p_kernel(int32_t neq, const double *__restrict__ r, const double *__restrict__ diag,
                   double *__restrict__ p, int first, const CgScalars *__restrict__ S)
{
    if ( S->done ) return;
    const double beta = S->beta;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        double z = diag ? r[i] * diag[i] : r[i];
        p[i] = first ? z : z + beta * p[i];
    }
}

// after the SpMV: red[0] = p.q from the per-CTA partials of the SpMV kernel; FINAL: alpha = rho / p.q
template< bool FINAL >
__global__ void __launch_bounds__(kCgThreads)
cg_pq_kernel(const double *__restrict__ partials, int P, double *__restrict__ red, CgScalars *S)
{
    __shared__ double scratch[32];
    if ( S->done ) return;
    double s = sum_partials(partials, P, scratch);
    if ( threadIdx.x == 0 ) {
        red[0] = s;
        if ( FINAL ) cg_scalars(1, 0, 0.0, red, S);
    }
}

// x += alpha p; r -= alpha q; partials of r.r and r.z for the next test / rho       (cg.h:56-58)
// FINAL: the last CTA reduces them and runs the convergence test + rho/beta recurrences.
template< bool FINAL >
__global__ void __launch_bounds__(kCgThreads)
cg_update_xr_kernel(int32_t neq, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                    const double *__restrict__ q, const double *__restrict__ diag,
                    const unsigned char *__restrict__ owned, double *__restrict__ partials, int P,
                    double *__restrict__ red, int iter, double tol, CgScalars *S)
{
    __shared__ double scratch[32];
    if ( S->done ) return;
    const double alpha = S->alpha;
    double rr = 0.0, rz = 0.0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        x[i] += alpha * p[i];
        double ri = r[i] - alpha * q[i];
        r[i] = ri;
        if ( !owned || owned[i] ) {
            rr += ri * ri;
            rz += ri * ( diag ? ri * diag[i] : ri );
        }
    }
    rr = block_sum(rr, scratch);
    if ( threadIdx.x == 0 ) partials[blockIdx.x] = rr;
    rz = block_sum(rz, scratch);
    if ( threadIdx.x == 0 ) partials[P + blockIdx.x] = rz;
    if ( last_block(&S->ticket) ) {
        double s0 = sum_partials(partials, gridDim.x, scratch);
        double s1 = sum_partials(partials + P, gridDim.x, scratch);
        if ( threadIdx.x == 0 ) {
            red[0] = s0; red[1] = s1;
            if ( FINAL ) cg_scalars(2, iter, tol, red, S);
        }
    }
}

// One-GPU iteration tail in ONE cooperative launch (cg.h:54-66 and 45-52 of the next iteration):
//   alpha = rho / p.q (every CTA sums the SpMV's per-CTA partials in the same fixed order),
//   x += alpha p, r -= alpha q, partial sums of r.r and r.z,
//   grid barrier, every CTA sums the partials in the same fixed order -> resid, rho, beta,
//   p = z + beta p for the next iteration (r and p of this CTA's range are still in L2).
// Replaces cg_pq_kernel + cg_update_xr_kernel + cg_update_p_kernel: two launches per iteration instead
// of four, and the second pass over r/p no longer goes to HBM.
__global__ void __launch_bounds__(kCgThreads)
cg_xr_p_kernel(int32_t neq, double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, const double *__restrict__ q,
               const double *__restrict__ diag, const double *__restrict__ spmv_partials, int nb, double *__restrict__ partials,
               int P, double *__restrict__ red, int iter, double tol, CgScalars *S)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ double scratch[32];
    __shared__ double bc[2];
    if ( S->done ) return;               // uniform: S is only written behind the grid barrier
    const double rho = S->rho, normb = S->normb;
    double pq = sum_partials(spmv_partials, nb, scratch);
    if ( threadIdx.x == 0 ) bc[0] = pq;
    __syncthreads();
    const double alpha = rho / bc[0];                    // alpha = rho / dot(p, q)     (cg.h:54)
    double rr = 0.0, rz = 0.0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    // two (three) independent elements per trip: ten (fifteen) loads in flight per thread
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    for ( ; i + 2 * stride < neq; i += 3 * stride ) {
        const int64_t j = i + stride, k = j + stride;
        const double pi = p[i], pj = p[j], pk = p[k], qi = q[i], qj = q[j], qk = q[k];
        const double xi = x[i], xj = x[j], xk = x[k], r0i = r[i], r0j = r[j], r0k = r[k];
        const double di = diag ? diag[i] : 1.0, dj = diag ? diag[j] : 1.0, dk = diag ? diag[k] : 1.0;
        const double ri = r0i - alpha * qi, rj = r0j - alpha * qj, rk = r0k - alpha * qk;
        x[i] = xi + alpha * pi; x[j] = xj + alpha * pj; x[k] = xk + alpha * pk;
        r[i] = ri; r[j] = rj; r[k] = rk;
        rr += ri * ri; rz += ri * ( diag ? ri * di : ri );
        rr += rj * rj; rz += rj * ( diag ? rj * dj : rj );
        rr += rk * rk; rz += rk * ( diag ? rk * dk : rk );
    }
    for ( ; i < neq; i += stride ) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        rr += ri * ri;
        rz += ri * ( diag ? ri * diag[i] : ri );
    }
    rr = block_sum(rr, scratch);
    if ( threadIdx.x == 0 ) partials[P + blockIdx.x] = rr;
    rz = block_sum(rz, scratch);
    if ( threadIdx.x == 0 ) partials[2 * P + blockIdx.x] = rz;
    __threadfence();
    grid.sync();
    const double s0 = sum_partials(partials + P, gridDim.x, scratch);
    __syncthreads();
    const double s1 = sum_partials(partials + 2 * P, gridDim.x, scratch);
    if ( threadIdx.x == 0 ) { bc[0] = s0; bc[1] = s1; }
    __syncthreads();
    const double t0 = bc[0], t1 = bc[1];
    if ( blockIdx.x == 0 && threadIdx.x == 0 ) {
        red[0] = t0; red[1] = t1;
        S->alpha = alpha;
        cg_scalars(2, iter, tol, red, S);               // resid test, rho_1 = rho, rho = r.z, beta     (cg.h:57-66, 47-51)
    }
    if ( sqrt(t0) / normb <= tol ) return;               // converged: every CTA takes the same branch
    const double beta = t1 / rho;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride ) {
        const double z = diag ? r[i] * diag[i] : r[i];
        p[i] = z + beta * p[i];
    }
}

// masked dot product (distributed path: p.q after the halo sum); last CTA leaves the sum in red[0]
__global__ void __launch_bounds__(kCgThreads)
cg_dot_kernel(int32_t neq, const double *__restrict__ a, const double *__restrict__ b,
              const unsigned char *__restrict__ owned, double *__restrict__ partials, double *__restrict__ red, CgScalars *S)
{
    __shared__ double scratch[32];
    if ( S->done ) return;
    double s = 0.0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < neq; i += stride )
        if ( !owned || owned[i] ) s += a[i] * b[i];
    s = block_sum(s, scratch);
    if ( threadIdx.x == 0 ) partials[blockIdx.x] = s;
    if ( last_block(&S->ticket) ) {
        double t = sum_partials(partials, gridDim.x, scratch);
        if ( threadIdx.x == 0 ) red[0] = t;
    }
}

// distributed path: scalar update after the all-reduce of red[]
__global__ void cg_scalars_kernel(int stage, int iter, double tol, const double *__restrict__ red, CgScalars *S)
{
    if ( S->done ) return;
    cg_scalars(stage, iter, tol, red, S);
}

// distributed path over peer memory (comm_p2p.cu): all-gather of the locally reduced sums through the
// mailboxes, summed in rank order by every rank (bit-identical everywhere), then the scalar update --
// replaces ncclAllReduce + cg_scalars_kernel.  nA + nB > 0: red[0] is first formed from the per-CTA
// partial sums of the fused SpMV (pa) and of the halo pull (pb) in a fixed order.  Lane r of warp 0
// talks to rank r; a value travels as two self-validating words (spmv_halo.h: ll_store).
__global__ void __launch_bounds__(kCgThreads)
cg_scalars_p2p_kernel(int stage, int iter, double tol, double *__restrict__ red, int nred, CgScalars *S, int nranks,
                      const double *__restrict__ pa, int nA, const double *__restrict__ pb, int nB,
                      unsigned long long *const *__restrict__ peer_sdata, const unsigned long long *own_sdata,
                      int64_t half_words, unsigned int seq, int *error)
{
    __shared__ double scratch[32];
    if ( S->done ) return;
    if ( nA + nB > 0 ) {
        const double a = sum_partials(pa, nA, scratch);
        __syncthreads();
        const double b = sum_partials(pb, nB, scratch);
        if ( threadIdx.x == 0 ) red[0] = a + b;
        __syncthreads();
    }
    if ( threadIdx.x >= 32 ) return;
    const int lane = threadIdx.x;
    if ( lane < nranks )
        for ( int j = 0; j < nred; j++ ) ll_store(peer_sdata[lane] + half_words + 2 * j, red[j], seq);
    double v[kRedMax] = { 0.0, 0.0, 0.0 };
    bool ok = true;
    if ( lane < nranks ) {
        unsigned long long t0 = 0, t1;
        for ( int j = 0; j < nred && ok; j++ ) {
            const unsigned long long *slot = own_sdata + half_words + 2 * ( (int64_t) lane * 4 + j );
            unsigned long long w0, w1;
            for ( int spin = 0;; spin++ ) {
                asm volatile( "ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"( w0 ), "=l"( w1 ) : "l"( slot ) : "memory" );
                if ( (unsigned int)( w0 >> 32 ) == seq && (unsigned int)( w1 >> 32 ) == seq ) break;
                if ( spin == 64 ) asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t0 ) );
                if ( spin > 64 ) {
                    asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t1 ) );
                    if ( t1 - t0 > 4000000000ull ) { ok = false; break; }
                    __nanosleep(32);
                }
            }
            v[j] = __longlong_as_double((long long)( ( w0 & 0xffffffffull ) | ( w1 << 32 ) ));
        }
    }
    if ( !__all_sync(0xffffffffu, ok) ) {
        if ( lane == 0 ) { atomicExch(error, 1); S->done = 1; }
        return;
    }
    double tot[kRedMax] = { 0.0, 0.0, 0.0 };
    for ( int r = 0; r < nranks; r++ )
#pragma unroll
        for ( int j = 0; j < kRedMax; j++ ) tot[j] += __shfl_sync(0xffffffffu, v[j], r);
    if ( lane == 0 ) {
        for ( int j = 0; j < nred; j++ ) red[j] = tot[j];
        cg_scalars(stage, iter, tol, tot, S);
    }
}


// ---- distributed iteration tail in ONE cooperative launch ------------------------------------------------------------
// Everything of a distributed CG iteration behind the product (which has pushed this rank's sums of the shared rows into
// the sharers' mailboxes in its epilogue): halo pull, all-gather of p.q, x/r update, all-gather of r.r and r.z, p update
// for the next iteration.  Five launches of the round-1 sequence (halo_pull_kernel, cg_scalars_p2p_kernel,
// cg_update_xr_kernel, cg_scalars_p2p_kernel, cg_update_p_kernel) become one; product + tail = two launches per iteration,
// as on one GPU.  The grid barriers are cooperative-groups grid syncs; the two scalar exchanges are done by warp 0 of
// CTA 0 through the mailboxes exactly as in cg_scalars_p2p_kernel (same words, same rank order on every rank; the
// per-CTA partial sums are grouped differently from the six-launch path, so the two paths agree to round-off only).
struct CgTailArgs {
    int32_t neq;
    double *x, *r, *p, *q;
    const double *diag;
    const unsigned char *owned;
    const double *spmv_partials;
    int nA;
    double *partials;                       // [3][P]: pull p.q | r.r | r.z per CTA
    int P;
    double *red;
    int iter;
    double tol;
    CgScalars *S;
    // halo pull (halo_pull_kernel)
    const int32_t *ueq, *uptr, *umb, *ubefore;
    const unsigned long long *mail;
    unsigned int seq_halo;
    int64_t nuniq;
    // scalar all-gathers (cg_scalars_p2p_kernel)
    int nranks;
    unsigned long long *const *peer_sdata;
    const unsigned long long *own_sdata;
    int64_t half1, half2;
    unsigned int seq1, seq2;
    int *error;
};

// wait for the two words of a mailbox entry (comm_p2p.cu: ll_load); false on timeout
__device__ __forceinline__ bool cg_ll_load(const unsigned long long *slot, unsigned int seq, double &v)
{
    unsigned long long w0, w1, t0 = 0, t1;
    for ( int spin = 0;; spin++ ) {
        asm volatile( "ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"( w0 ), "=l"( w1 ) : "l"( slot ) : "memory" );
        if ( (unsigned int)( w0 >> 32 ) == seq && (unsigned int)( w1 >> 32 ) == seq ) break;
        if ( spin == 64 ) asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t0 ) );
        if ( spin > 64 ) {
            asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t1 ) );
            if ( t1 - t0 > 4000000000ull ) return false;
            __nanosleep(32);
        }
    }
    v = __longlong_as_double((long long)( ( w0 & 0xffffffffull ) | ( w1 << 32 ) ));
    return true;
}

// warp 0 of one CTA: all-gather red[0..nred) over the ranks (lane r talks to rank r), sum in rank order, scalar update
__device__ __forceinline__ void cg_scalars_allgather(int stage, int iter, double tol, double *red, int nred, CgScalars *S, int nranks,
                                                     unsigned long long *const *peer_sdata, const unsigned long long *own_sdata,
                                                     int64_t half_words, unsigned int seq, int *error)
{
    const int lane = threadIdx.x;
    if ( lane < nranks )
        for ( int j = 0; j < nred; j++ ) ll_store(peer_sdata[lane] + half_words + 2 * j, red[j], seq);
    double v[kRedMax] = { 0.0, 0.0, 0.0 };
    bool ok = true;
    if ( lane < nranks )
        for ( int j = 0; j < nred && ok; j++ ) ok = cg_ll_load(own_sdata + half_words + 2 * ( (int64_t) lane * 4 + j ), seq, v[j]);
    if ( !__all_sync(0xffffffffu, ok) ) {
        if ( lane == 0 ) { atomicExch(error, 1); S->done = 1; }
        return;
    }
    double tot[kRedMax] = { 0.0, 0.0, 0.0 };
    for ( int r = 0; r < nranks; r++ )
#pragma unroll
        for ( int j = 0; j < kRedMax; j++ ) tot[j] += __shfl_sync(0xffffffffu, v[j], r);
    if ( lane == 0 ) {
        for ( int j = 0; j < nred; j++ ) red[j] = tot[j];
        cg_scalars(stage, iter, tol, tot, S);
    }
}

__global__ void __launch_bounds__(kCgThreads)
cg_tail_p2p_kernel(CgTailArgs a)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ double scratch[32];
    volatile CgScalars *Sv = a.S;
    if ( Sv->done ) return;                  // uniform: S is only written behind grid barriers
    const int64_t stride = (int64_t) gridDim.x * blockDim.x, first = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    // ---- halo pull: q[shared row] = sum over the sharers in rank order; p.q over the shared rows this rank owns ----
    {
        double dot = 0.0;
        bool ok = true;
        for ( int64_t u = first; u < a.nuniq; u += stride ) {
            const int b = a.uptr[u], e = a.uptr[u + 1], nb = a.ubefore[u];
            const int row = a.ueq[u];
            const double own = a.q[row];
            double acc = 0.0;
            for ( int t = b; t < e; t++ ) {
                if ( t - b == nb ) acc += own;
                double v = 0.0;
                ok = cg_ll_load(a.mail + 2 * (int64_t) a.umb[t], a.seq_halo, v) && ok;
                acc += v;
            }
            if ( nb == e - b ) acc += own;
            a.q[row] = acc;
            if ( a.owned[row] ) dot += a.p[row] * acc;
        }
        if ( !ok ) atomicExch(a.error, 1);
        dot = block_sum(dot, scratch);
        if ( threadIdx.x == 0 ) a.partials[blockIdx.x] = dot;
    }
    __threadfence();
    grid.sync();
    // ---- alpha = rho / p.q over all ranks ----
    if ( blockIdx.x == 0 ) {
        const double s0 = sum_partials(a.spmv_partials, a.nA, scratch);
        __syncthreads();
        const double s1 = sum_partials(a.partials, gridDim.x, scratch);
        if ( threadIdx.x == 0 ) a.red[0] = s0 + s1;
        __syncthreads();
        if ( threadIdx.x < 32 )
            cg_scalars_allgather(1, a.iter, a.tol, a.red, 1, a.S, a.nranks, a.peer_sdata, a.own_sdata, a.half1, a.seq1, a.error);
    }
    __threadfence();
    grid.sync();
    if ( Sv->done ) return;                  // a peer wait timed out
    const double alpha = Sv->alpha;
    // ---- x += alpha p, r -= alpha q, partial sums of r.r and r.z over the owned rows ----
    {
        double rr = 0.0, rz = 0.0;
        int64_t i = first;
        // three independent elements per trip: eighteen loads in flight per thread
        for ( ; i + 2 * stride < a.neq; i += 3 * stride ) {
            const int64_t j = i + stride, k = j + stride;
            const double pi = a.p[i], pj = a.p[j], pk = a.p[k], qi = a.q[i], qj = a.q[j], qk = a.q[k];
            const double xi = a.x[i], xj = a.x[j], xk = a.x[k], r0i = a.r[i], r0j = a.r[j], r0k = a.r[k];
            const double di = a.diag ? a.diag[i] : 1.0, dj = a.diag ? a.diag[j] : 1.0, dk = a.diag ? a.diag[k] : 1.0;
            const double oi = a.owned[i] ? 1.0 : 0.0, oj = a.owned[j] ? 1.0 : 0.0, ok = a.owned[k] ? 1.0 : 0.0;
            const double ri = r0i - alpha * qi, rj = r0j - alpha * qj, rk = r0k - alpha * qk;
            a.x[i] = xi + alpha * pi; a.x[j] = xj + alpha * pj; a.x[k] = xk + alpha * pk;
            a.r[i] = ri; a.r[j] = rj; a.r[k] = rk;
            if ( oi != 0.0 ) { rr += ri * ri; rz += ri * ( a.diag ? ri * di : ri ); }
            if ( oj != 0.0 ) { rr += rj * rj; rz += rj * ( a.diag ? rj * dj : rj ); }
            if ( ok != 0.0 ) { rr += rk * rk; rz += rk * ( a.diag ? rk * dk : rk ); }
        }
        for ( ; i < a.neq; i += stride ) {
            a.x[i] += alpha * a.p[i];
            const double ri = a.r[i] - alpha * a.q[i];
            a.r[i] = ri;
            if ( a.owned[i] ) {
                rr += ri * ri;
                rz += ri * ( a.diag ? ri * a.diag[i] : ri );
            }
        }
        rr = block_sum(rr, scratch);
        if ( threadIdx.x == 0 ) a.partials[a.P + blockIdx.x] = rr;
        rz = block_sum(rz, scratch);
        if ( threadIdx.x == 0 ) a.partials[2 * a.P + blockIdx.x] = rz;
    }
    __threadfence();
    grid.sync();
    // ---- residual test, rho, beta over all ranks ----
    if ( blockIdx.x == 0 ) {
        const double s0 = sum_partials(a.partials + a.P, gridDim.x, scratch);
        __syncthreads();
        const double s1 = sum_partials(a.partials + 2 * a.P, gridDim.x, scratch);
        if ( threadIdx.x == 0 ) { a.red[0] = s0; a.red[1] = s1; }
        __syncthreads();
        if ( threadIdx.x < 32 )
            cg_scalars_allgather(2, a.iter, a.tol, a.red, 2, a.S, a.nranks, a.peer_sdata, a.own_sdata, a.half2, a.seq2, a.error);
    }
    __threadfence();
    grid.sync();
    if ( Sv->done ) return;                  // converged (cg.h:60-64), or a peer wait timed out
    const double beta = Sv->beta;
    // ---- p = z + beta p for the next iteration (cg.h:45-52) ----
    int64_t i = first;
    for ( ; i + 2 * stride < a.neq; i += 3 * stride ) {
        const int64_t j = i + stride, k = j + stride;
        const double ri = a.r[i], rj = a.r[j], rk = a.r[k], pi = a.p[i], pj = a.p[j], pk = a.p[k];
        const double di = a.diag ? a.diag[i] : 1.0, dj = a.diag ? a.diag[j] : 1.0, dk = a.diag ? a.diag[k] : 1.0;
        a.p[i] = ( a.diag ? ri * di : ri ) + beta * pi;
        a.p[j] = ( a.diag ? rj * dj : rj ) + beta * pj;
        a.p[k] = ( a.diag ? rk * dk : rk ) + beta * pk;
    }
    for ( ; i < a.neq; i += stride ) {
        const double z = a.diag ? a.r[i] * a.diag[i] : a.r[i];
        a.p[i] = z + beta * a.p[i];
    }
}

struct CgWork {
    double *r, *p, *q, *partials, *red;
    CgScalars *S;
};

static int cg_run(ob200_csr *A, ob200_comm *comm, const double *b_dev, double *x_dev, int precond, int max_iter,
                  double tol, int *iters, double *resid)
{
    ob200_context *ctx = A->ctx;
    const int32_t n = A->neq;
    const int P = kCgMaxBlocks;                         // stride between the partial arrays
    const int G = ctx->shape.grid(n, kCgThreads, 8);    // CTAs of the vector kernels (<= 148*8 <= P)
    const int64_t need = 3 * (int64_t) n + (int64_t)( kRedMax + 1 ) * P + kRedMax + 16;
    if ( A->work.n < need ) OB_CHECK( A->work.alloc(need) );
    CgWork w;
    w.r = A->work.p;
    w.p = w.r + n;
    w.q = w.p + n;
    w.partials = w.q + n;
    w.red = w.partials + (int64_t)( kRedMax + 1 ) * P;         // [kRedMax][P] reduction partials + [P] the product's p.q partials (distributed tail)
    w.S = reinterpret_cast< CgScalars * >( w.red + kRedMax + 1 );
    OB_CUDA( cudaMemsetAsync(w.partials, 0, sizeof( double ) * ( (size_t)( kRedMax + 1 ) * P + kRedMax + 16 ), ctx->stream) );
    const unsigned char *owned = comm ? comm->owned.p : nullptr;
    const bool dist = comm && comm->nranks > 1;

    // a pending lazy zero() must be visible to the preconditioner set-up below (it reads val directly)
    OB_CHECK( ob200_csr_materialize(A) );
    // preconditioner (IMLSolver::solve re-inits when the matrix version changed, imlsolver.C:110-114)
    const double *diag = nullptr;
    if ( precond == OB200_PRECOND_DIAG ) {
        if ( A->diag.n < n || A->diag_version != A->version || comm ) {
            if ( A->diag.n < n ) OB_CHECK( A->diag.alloc(n > 0 ? n : 1) );
            DevBuf< int > bad;
            OB_CHECK( bad.alloc(1) );
            OB_CUDA( cudaMemsetAsync(bad.p, 0, sizeof( int ), ctx->stream) );
            if ( n ) {
                int grid = ctx->shape.grid(n, 256, 8);
                if ( !comm ) {
                    OB_LAUNCH(ctx, diag_init_kernel, grid, 256, 0, n, A->rowptr.p, A->colind.p, A->val.p, A->diag.p, bad.p);
                } else {
                    OB_LAUNCH(ctx, diag_extract_kernel, grid, 256, 0, n, A->rowptr.p, A->colind.p, A->val.p, A->diag.p);
                    OB_CHECK( comm_exchange_add(comm, A->diag.p) );
                    OB_LAUNCH(ctx, diag_invert_kernel, grid, 256, 0, n, A->diag.p, bad.p);
                }
            }
            int h = 0;
            OB_CUDA( cudaMemcpyAsync(&h, bad.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
            OB_CUDA( cudaStreamSynchronize(ctx->stream) );
            OB_REQUIRE(h == 0, OB200_EZERODIAG, "DiagPreconditioner: failed, zero diagonal detected in %d equations", h);
            A->diag_version = A->version;
        }
        diag = A->diag.p;
    } else {
        OB_REQUIRE(precond == OB200_PRECOND_VOID, OB200_EINVAL, "cg_solve: unknown preconditioner type %d", precond);
    }

    // distributed: all-reduce the locally reduced sums, then the scalar update in its own launch
    const bool p2p = dist && comm->p2p;
    const bool allow_fused = !( getenv("OB200_FUSED_HALO") && getenv("OB200_FUSED_HALO")[0] == '0' );
    const bool fused_halo = allow_fused && p2p && comm->nneigh > 0 && comm->route.p && spmv_halo_supported(A);
    const ob200_mailbox_layout ML = mailbox_layout(comm ? comm->nranks : 1, comm ? comm->cap : 1);
    auto finish = [&](int stage, int iter, int nred, const double *pa = nullptr, int nA = 0, const double *pb = nullptr, int nB = 0) -> int {
        if ( !dist ) return OB200_OK;
        if ( p2p ) {
            const unsigned int seq = ++comm->scal_seq;
            const int64_t par = seq & 1u;
            OB_LAUNCH(ctx, cg_scalars_p2p_kernel, 1, kCgThreads, 0, stage, iter, tol, w.red, nred, w.S, comm->nranks, pa, nA, pb, nB,
                      comm->p_sdata.p, reinterpret_cast< const unsigned long long * >( comm->mailbox + ML.sdata ),
                      2 * par * ML.sdata_half, seq, reinterpret_cast< int * >( comm->mailbox + ML.error ));
            return OB200_OK;
        }
        OB_CHECK( comm_allreduce_sum(comm, w.red, nred) );
        OB_LAUNCH(ctx, cg_scalars_kernel, 1, 1, 0, stage, iter, tol, w.red, w.S);
        return OB200_OK;
    };

    // r = b - A x, resid test (cg.h:31-43)
    OB_CHECK( spmv(A, x_dev, w.q) );
    if ( comm ) OB_CHECK( comm_exchange_add(comm, w.q) );
    if ( dist ) OB_LAUNCH(ctx, cg_init_kernel< false >, G, kCgThreads, 0, n, b_dev, w.q, w.r, diag, owned, w.partials, P, w.red, tol, w.S);
    else OB_LAUNCH(ctx, cg_init_kernel< true >, G, kCgThreads, 0, n, b_dev, w.q, w.r, diag, owned, w.partials, P, w.red, tol, w.S);
    OB_CHECK( finish(0, 0, 3) );

    // one GPU: the iteration tail as one cooperative launch, if the device can hold the whole grid at once
    bool coop = false;
    int coop_grid = 0;
    if ( !comm && n > 0 && !( getenv("OB200_CG_COOP") && getenv("OB200_CG_COOP")[0] == '0' ) ) {
        int can = 0, per_sm = 0;
        cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, ctx->device);
        if ( can && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_xr_p_kernel, kCgThreads, 0) == cudaSuccess && per_sm > 0 ) {
            coop_grid = per_sm * ctx->shape.sms;
            if ( coop_grid > G ) coop_grid = G;
            coop = coop_grid <= kCgMaxBlocks;
        }
        cudaGetLastError();
    }

    // distributed over peer memory: the same idea, with the halo pull and the two scalar exchanges inside the launch
    bool dcoop = false;
    if ( fused_halo && n > 0 && !( getenv("OB200_CG_COOP") && getenv("OB200_CG_COOP")[0] == '0' ) ) {
        int can = 0, per_sm = 0;
        cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, ctx->device);
        if ( can && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_tail_p2p_kernel, kCgThreads, 0) == cudaSuccess && per_sm > 0 ) {
            coop_grid = per_sm * ctx->shape.sms;
            if ( coop_grid > G ) coop_grid = G;
            dcoop = coop_grid <= kCgMaxBlocks && coop_grid <= P;
        }
        cudaGetLastError();
    }

    CgScalars h;
    const int poll = ( coop || dcoop ) ? 32 : 8;       // iterations enqueued between two looks at the device-side stop flag
    int it = 0;
    bool done = false;
    while ( !done ) {
        int batch_end = it + poll < max_iter ? it + poll : max_iter;
        for ( ; it < batch_end; ) {
            it++;
            if ( !( coop || dcoop ) || it == 1 ) OB_LAUNCH(ctx, cg_update_p_kernel, G, kCgThreads, 0, n, w.r, diag, w.p, it == 1 ? 1 : 0, w.S);
            if ( dcoop ) {
                // two launches per iteration: the product (p.q partials of the unshared rows, shared rows pushed to the
                // sharers), then the cooperative tail
                const unsigned int seq = ++comm->halo_seq;
                const SpmvHalo hv{ comm->route.p, comm->uniq_ptr.p, comm->push_dst.p, 2 * (int64_t)( seq & 1u ) * ML.data_half, seq };
                int nb = 0;
                OB_CHECK( spmv_fused_halo(A, w.p, w.q, w.partials + 3 * (int64_t) P, &nb, &w.S->done, hv) );
                CgTailArgs ta;
                ta.neq = n; ta.x = x_dev; ta.r = w.r; ta.p = w.p; ta.q = w.q; ta.diag = diag; ta.owned = owned;
                ta.spmv_partials = w.partials + 3 * (int64_t) P; ta.nA = nb; ta.partials = w.partials; ta.P = P; ta.red = w.red;
                ta.iter = it; ta.tol = tol; ta.S = w.S;
                ta.ueq = comm->uniq_eq.p; ta.uptr = comm->uniq_ptr.p; ta.umb = comm->uniq_mb.p; ta.ubefore = comm->uniq_before.p;
                ta.mail = reinterpret_cast< const unsigned long long * >( comm->mailbox + ML.data ) + 2 * (int64_t)( seq & 1u ) * ML.data_half;
                ta.seq_halo = seq; ta.nuniq = comm->nuniq;
                ta.nranks = comm->nranks; ta.peer_sdata = comm->p_sdata.p;
                ta.own_sdata = reinterpret_cast< const unsigned long long * >( comm->mailbox + ML.sdata );
                ta.seq1 = ++comm->scal_seq; ta.half1 = 2 * (int64_t)( ta.seq1 & 1u ) * ML.sdata_half;
                ta.seq2 = ++comm->scal_seq; ta.half2 = 2 * (int64_t)( ta.seq2 & 1u ) * ML.sdata_half;
                ta.error = reinterpret_cast< int * >( comm->mailbox + ML.error );
                void *args[] = { &ta };
                ob200_prof_rec pr = { "cg_tail_p2p_kernel", nullptr, nullptr };
                if ( ctx->profiling ) {
                    pr.start = ctx->prof_event();
                    pr.stop = ctx->prof_event();
                    cudaEventRecord(pr.start, ctx->stream);
                }
                OB_CUDA( cudaLaunchCooperativeKernel((void *) cg_tail_p2p_kernel, dim3(coop_grid), dim3(kCgThreads), args, 0, ctx->stream) );
                if ( ctx->profiling ) {
                    cudaEventRecord(pr.stop, ctx->stream);
                    ctx->prof_pending.push_back(pr);
                }
                ctx->launches++;
                continue;
            }
            if ( coop ) {
                // two launches per iteration: the product with its p.q partials, then everything else
                int nb = 0;
                OB_CHECK( spmv_fused_dot(A, w.p, w.q, w.partials, &nb, &w.S->done) );
                int32_t n_ = n;
                double *x_ = x_dev, *r_ = w.r, *p_ = w.p, *q_ = w.q, *pa_ = w.partials, *red_ = w.red;
                const double *d_ = diag, *sp_ = w.partials;
                int P_ = P, it_ = it;
                double tol_ = tol;
                CgScalars *S_ = w.S;
                void *args[] = { &n_, &x_, &r_, &p_, &q_, &d_, &sp_, &nb, &pa_, &P_, &red_, &it_, &tol_, &S_ };
                ob200_prof_rec pr = { "cg_xr_p_kernel", nullptr, nullptr };
                if ( ctx->profiling ) {
                    pr.start = ctx->prof_event();
                    pr.stop = ctx->prof_event();
                    cudaEventRecord(pr.start, ctx->stream);
                }
                OB_CUDA( cudaLaunchCooperativeKernel((void *) cg_xr_p_kernel, dim3(coop_grid), dim3(kCgThreads), args, 0, ctx->stream) );
                if ( ctx->profiling ) {
                    cudaEventRecord(pr.stop, ctx->stream);
                    ctx->prof_pending.push_back(pr);
                }
                ctx->launches++;
                continue;
            }
            if ( !comm ) {
                int nb = 0;
                OB_CHECK( spmv_fused_dot(A, w.p, w.q, w.partials, &nb, &w.S->done) );      // q = A p, partials of p.q
                OB_LAUNCH(ctx, cg_pq_kernel< true >, 1, kCgThreads, 0, w.partials, nb, w.red, w.S);
            } else if ( fused_halo ) {
                // one kernel: q = A p, p.q over the rows only this rank holds, shared rows pushed to the sharers;
                // the pull sums the sharers' values and adds p.q over the shared rows this rank owns
                const unsigned int seq = ++comm->halo_seq;
                const SpmvHalo hv{ comm->route.p, comm->uniq_ptr.p, comm->push_dst.p, 2 * (int64_t)( seq & 1u ) * ML.data_half, seq };
                int nb = 0;
                OB_CHECK( spmv_fused_halo(A, w.p, w.q, w.partials, &nb, &w.S->done, hv) );
                OB_CHECK( comm_p2p_pull(comm, w.q, seq, w.p, &w.S->done) );
                OB_CHECK( finish(1, it, 1, w.partials, nb, comm->pull_partials.p, comm->pull_grid) );
            } else {
                OB_CHECK( spmv_fused_dot(A, w.p, w.q, nullptr, nullptr, &w.S->done) );
                OB_CHECK( comm_exchange_add(comm, w.q, &w.S->done) );
                OB_LAUNCH(ctx, cg_dot_kernel, G, kCgThreads, 0, n, w.p, w.q, owned, w.partials, w.red, w.S);
                if ( dist ) OB_CHECK( finish(1, it, 1) );
                else OB_LAUNCH(ctx, cg_scalars_kernel, 1, 1, 0, 1, it, tol, w.red, w.S);
            }
            if ( dist ) OB_LAUNCH(ctx, cg_update_xr_kernel< false >, G, kCgThreads, 0, n, x_dev, w.r, w.p, w.q, diag, owned, w.partials, P, w.red, it, tol, w.S);
            else OB_LAUNCH(ctx, cg_update_xr_kernel< true >, G, kCgThreads, 0, n, x_dev, w.r, w.p, w.q, diag, owned, w.partials, P, w.red, it, tol, w.S);
            OB_CHECK( finish(2, it, 2) );
        }
        OB_CUDA( cudaMemcpyAsync(&h, w.S, sizeof( CgScalars ), cudaMemcpyDeviceToHost, ctx->stream) );
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );
        done = h.done || it >= max_iter;
    }
    if ( p2p ) OB_CHECK( comm_p2p_check(comm) );
    *iters = h.done ? h.iters : max_iter;
    *resid = h.resid;
    return h.done ? 0 : 1;        // cg.h: 0 converged, 1 max_iter reached
}

} // namespace ob200

using namespace ob200;

static int cg_entry(ob200_csr *A, ob200_comm *comm, const double *b, double *x, int precond, int max_iter, double tol,
                    int *iters, double *resid, int on_device)
{
    OB_REQUIRE(A && iters && resid, OB200_EINVAL, "cg_solve: null argument");
    OB_REQUIRE(A->rowptr.p, OB200_EINVAL, "cg_solve: matrix has no structure");
    OB_REQUIRE(A->neq == 0 || ( b && x ), OB200_EINVAL, "cg_solve: size mismatch");       // imlsolver.C:105-107
    OB_REQUIRE(max_iter >= 0, OB200_EINVAL, "cg_solve: negative max_iter");
    Staged< double > B;
    StagedOut< double > X;
    OB_CHECK( B.stage(A->ctx, b, A->neq, on_device) );
    OB_CHECK( X.stage(A->ctx, x, A->neq, on_device, true) );
    int flag = cg_run(A, comm, B.d, X.d, precond, max_iter, tol, iters, resid);
    if ( flag < 0 ) return flag;
    OB_CHECK( X.finish(A->ctx) );
    return flag;
}

extern "C" {

int ob200_cg_solve(ob200_csr *A, const double *b, double *x, int precond, int max_iter, double tol, int *iters,
                   double *resid, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    return cg_entry(A, nullptr, b, x, precond, max_iter, tol, iters, resid, on_device);
}

int ob200_cg_solve_dist(ob200_csr *A, ob200_comm *c, const double *b, double *x, int precond, int max_iter, double tol,
                        int *iters, double *resid, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(c, OB200_EINVAL, "cg_solve_dist: null communicator");
    OB_REQUIRE(c->neq == ( A ? A->neq : 0 ), OB200_EINVAL, "cg_solve_dist: halo describes %d equations, matrix has %d", c->neq, A ? A->neq : 0);
    return cg_entry(A, c, b, x, precond, max_iter, tol, iters, resid, on_device);
}

} // extern "C"
