// fp32 instantiation of every engine kernel (sm_100a)
#include "brov_kernels_impl.cuh"
namespace brov { BROV_INSTANTIATE(float) }
