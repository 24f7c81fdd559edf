#include "cudacontext.h"
#include "error.h"
#include <cstdlib>

namespace oofem {
static ob200_context *theContext = nullptr;

static void releaseContext()
{
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
