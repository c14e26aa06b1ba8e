// ntt_fwd.cu -- forward-direction instantiations of the NTT kernels (own translation unit: compiled in parallel).
#include "ntt_kernels.cuh"
namespace ntt {
cudaError_t run_pass_fwd(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st) {
  return run_pass<DIR_FWD>(A, k, first, X, st);
}
cudaError_t configure_fwd() { return configure_dir<DIR_FWD>(); }
}  // namespace ntt
