// ntt_inv.cu -- inverse-direction instantiations of the NTT kernels (own translation unit: compiled in parallel).
#include "ntt_kernels.cuh"
namespace ntt {
cudaError_t run_pass_inv(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st) {
  return run_pass<DIR_INV>(A, k, first, X, st);
}
cudaError_t configure_inv() { return configure_dir<DIR_INV>(); }
}  // namespace ntt
