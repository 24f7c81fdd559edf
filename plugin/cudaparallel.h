// Host-side loops of the plugin over all elements / nodes of a domain (location arrays, connectivity, material statuses) on
// several threads.  Plain std::thread, not OpenMP: the reference's headers change class layouts under _OPENMP (omp_lock_t
// members in set.h, bctracker.h, connectivitytable.h ...), so plugin objects compiled with -fopenmp could not be linked with
// a serial build of OOFEM.  What the loops call per element (Element::giveLocationArray, Material::giveStatus,
// DofManager / Dof queries) is what the reference's own `#pragma omp parallel for` loops call concurrently
// (engngm.C:901-929, 1377-1407, structengngmodel.C:305-310).
#ifndef oofem_b200_cudaparallel_h
#define oofem_b200_cudaparallel_h

#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

namespace oofem {
/// Threads the plugin's host loops use: OOFEM_B200_THREADS, else the hardware concurrency, at most 16.
inline int cudaPluginThreads()
{
    static int n = 0;
    if ( n == 0 ) {
        int want = 0;
        if ( const char *e = std :: getenv("OOFEM_B200_THREADS") ) {
            want = std :: atoi(e);
        }
        if ( want <= 0 ) {
            want = ( int ) std :: min(16u, std :: max(1u, std :: thread :: hardware_concurrency() ) );
        }
        n = std :: max(1, std :: min(want, 64) );
    }
    return n;
}

/// fn(begin, end, thread) over [0, n) cut into one contiguous range per thread; serial below `grain` items per thread.
template< class F >
void parallelFor(long n, F &&fn, long grain = 4096)
{
    int nt = ( int ) std :: min< long >(cudaPluginThreads(), n / std :: max< long >(grain, 1) );
    if ( nt <= 1 ) {
        fn(0L, n, 0);
        return;
    }
    std :: vector< std :: thread >pool;
    pool.reserve(nt - 1);
    const long chunk = ( n + nt - 1 ) / nt;
    for ( int t = 1; t < nt; t++ ) {
        const long b = std :: min(n, t * chunk), e = std :: min(n, b + chunk);
        pool.emplace_back([ &fn, b, e, t ] { fn(b, e, t); });
    }
    fn(0L, std :: min(n, chunk), 0);
    for ( auto &th : pool ) {
        th.join();
    }
}
} // namespace oofem
#endif
