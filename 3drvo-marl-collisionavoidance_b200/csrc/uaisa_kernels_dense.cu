// The many-neighbour instantiations of the env kernel without helper slices (uaisa_env_kernel_d<10|16, 128, *, 0>), compiled as a translation
// unit of their own so that the Makefile can put them under -maxrregcount=128 (16 instead of 12 resident warps per SM; see uaisa_kernels.cu).
#define UAISA_TU_DENSE
#include "uaisa_kernels.cu"
