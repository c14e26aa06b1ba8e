// ntt_tma_inv.cu -- inverse-direction instantiations of the persistent TMA-fed NTT pass (own translation unit).
#include <algorithm>
#include "ntt_tma.cuh"
namespace ntt {
static int sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 1;
}
cudaError_t run_pass_tma_inv(const PassArgs& A, uint32_t k, bool first, const PassExtra& X, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!tma_pass_ok(A, k)) return cudaSuccess;
  const int sms = sm_count();
  // too few tiles to stream: the plain kernel spreads them over more SMs (env ZK_NTT_TMA_MIN_TILES: tests force the path)
  const char* env = getenv("ZK_NTT_TMA_MIN_TILES");  // read per call: tests switch it inside one process
  const long min_tiles = env ? atol(env) : -1L;
  const uint64_t tiles = (uint64_t)((A.nc + TILE_COLS - 1) / TILE_COLS) << (A.log_n - 10);
  if (tiles < (min_tiles >= 0 ? (uint64_t)min_tiles : 2ull * sms)) return cudaSuccess;
  // measured on the log-21 execution shard (7 + 7 + 7 stages, 56..64-column slabs): the tables are rebuilt every few
  // tiles and the short passes are HBM-bound anyway -- the plain kernel is 5 % faster there; wide, deep passes only
  if (min_tiles < 0 && (A.nc < 8 * TILE_COLS || k < 8)) return cudaSuccess;
  cudaError_t e = run_pass_tma_dir<DIR_INV>(A, k, first, X, st, sms);
  if (e == cudaErrorNotSupported) return cudaSuccess;
  *handled = true;
  tma_pass_counter()++;
  return e;
}
cudaError_t configure_tma_inv() { return configure_tma_dir<DIR_INV>(); }
}  // namespace ntt
