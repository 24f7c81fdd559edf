// Process-wide handle of the GPU context of liboofem_b200.so (one process = one GPU, like one MPI
// rank of OOFEM's parallel mode).  Device ordinal: OOFEM_B200_DEVICE (default: LOCAL_RANK, else 0).
#ifndef oofem_b200_cudacontext_h
#define oofem_b200_cudacontext_h

#include "oofem_b200.h"

namespace oofem {
class CudaContext
{
public:
    /// Creates the context on first use; OOFEM_ERROR when there is no CUDA device (no CPU fallback).
    static ob200_context *get();
    /// OOFEM_ERROR with the library's message when rc < 0; returns rc otherwise.
    static int check(int rc, const char *what);
    /// Wall-clock phase timers of the plugin (printed as one "CudaTiming" line at exit when OOFEM_B200_TIMING is set;
    /// bench.py's e2e_executable reads it): t is added to the named phase.
    static void addTime(const char *phase, double seconds);
    static double now();
};
/// adds the life time of the object to a phase
struct CudaPhaseTimer {
    const char *phase;
    double t0;
    explicit CudaPhaseTimer(const char *p) : phase(p), t0(CudaContext :: now()) { }
    ~CudaPhaseTimer() { CudaContext :: addTime(phase, CudaContext :: now() - t0); }
};
} // namespace oofem
#endif
