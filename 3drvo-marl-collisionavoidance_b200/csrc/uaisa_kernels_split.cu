// The owner / helper split instantiation of the env kernel (uaisa_env_kernel<4, 128, true, true>: small batches of few-neighbour warp-local envs),
// a translation unit of its own without a register cap (the batches it serves never fill the machine; see uaisa_kernels.cu).
#define UAISA_TU_SPLIT
#define UAISA_MINB128 2
#include "uaisa_kernels.cu"
