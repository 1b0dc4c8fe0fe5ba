#include "common.cuh"
#include "kernels.h"
namespace sb {
int launch_aux_forward(const AuxArgs& a, cudaStream_t s) { return 99; }
int launch_aux_backward(const AuxArgs& a, cudaStream_t s) { return 99; }
}
