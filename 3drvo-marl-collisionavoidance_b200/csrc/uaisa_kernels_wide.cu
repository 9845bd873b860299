// The sliced env kernels for CTAs of up to 192 threads (64 egos x 3 neighbour slices: clusters of 4 instead of 8 CTAs for a 256-drone env),
// a translation unit of its own: __launch_bounds__(192, 2) keeps the 168-register footprint of the 128-thread instantiations (see uaisa_kernels.cu).
#define UAISA_TU_WIDE
#include "uaisa_kernels.cu"
