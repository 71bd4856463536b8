// Instantiates the fused-loss kernels for K = 4 decoders (all C; fast/exact, vector/scalar, both modes).
#include "fused_loss_impl.cuh"
namespace uaps { namespace loss {
template int launch_pass1_k<4>(int, int, bool, bool, const LossArgs&, unsigned*, float*, double*, cudaStream_t);
template int launch_pass2_k<4>(int, int, bool, bool, const LossArgs&, const float*, const float*, cudaStream_t);
} }
