// Library identification.
#include "fd_common.cuh"

extern "C" const char* fd_version(void) { return "founddiff_b200 0.1.0 sm_100a"; }
