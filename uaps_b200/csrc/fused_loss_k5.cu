// Instantiates the fused-loss kernels for K = 5 decoders (all C; vector widths, fast / exact arithmetic, unlabeled / supervised).
#include "fused_loss_impl.cuh"
namespace uaps { namespace loss {
template int launch_loss_k<5>(int, int, bool, bool, const LossArgs&, float*, const float*, const float*, int*, cudaStream_t);
} }
