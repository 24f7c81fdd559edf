#include "cudacontext.h"
#include "error.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

namespace oofem {
static ob200_context *theContext = nullptr;

static std :: map< std :: string, double > &phaseTimes()
{
    static auto *m = new std :: map< std :: string, double >;     // never destroyed: read by the atexit handler below
    return * m;
}

double CudaContext :: now()
{
    return std :: chrono :: duration< double >( std :: chrono :: steady_clock :: now().time_since_epoch() ).count();
}

void CudaContext :: addTime(const char *phase, double seconds) { phaseTimes() [ phase ] += seconds; }

static void releaseContext()
{
    if ( std :: getenv("OOFEM_B200_TIMING") ) {
        std :: fprintf(stderr, "CudaTiming {");
        bool first = true;
        for ( auto &kv : phaseTimes() ) {
            std :: fprintf(stderr, "%s\"%s\": %.6f", first ? "" : ", ", kv.first.c_str(), kv.second);
            first = false;
        }
        std :: fprintf(stderr, "}\n");
    }
    if ( theContext ) {
        ob200_context_destroy(theContext);
        theContext = nullptr;
    }
}

ob200_context *CudaContext :: get()
{
    if ( !theContext ) {
        int dev = 0;
        if ( const char *e = std :: getenv("OOFEM_B200_DEVICE") ) {
            dev = std :: atoi(e);
        } else if ( const char *l = std :: getenv("LOCAL_RANK") ) {
            dev = std :: atoi(l);
        }
        if ( ob200_context_create(dev, & theContext) < 0 ) {
            OOFEM_ERROR("cudacsr/cudacg: %s", ob200_last_error());
        }
        std :: atexit(releaseContext);
    }
    return theContext;
}

int CudaContext :: check(int rc, const char *what)
{
    if ( rc < 0 ) {
        OOFEM_ERROR("%s: %s", what, ob200_last_error());
    }
    return rc;
}
} // namespace oofem
