// tracegen.cu -- C ABI of the device trace generation (kernels: tracegen.cuh).
#include "zkgpu_internal.cuh"
#include "tracegen.cuh"

namespace {

bool pow2(uint64_t x) { return x && !(x & (x - 1)); }

// events (host or device) -> device pointer on the ctx stream; `scope` owns a staging copy when one is made
int32_t stage_events(zk_ctx* c, DevScope& scope, const uint32_t* host, zk_dptr dev, uint64_t words, const uint32_t** out) {
  if (dev) { *out = (const uint32_t*)dev; return ZK_OK; }
  uint32_t* d = nullptr;
  int32_t rc = scope.alloc(&d, std::max<uint64_t>(words, 1) * 4);
  if (rc) return rc;
  if (words) CK(cudaMemcpyAsync(d, host, words * 4, cudaMemcpyHostToDevice, c->stream));
  *out = d;
  return ZK_OK;
}

int32_t poseidon2_wide(zk_ctx* c, const uint32_t* host, zk_dptr dev, uint64_t n_events, uint64_t rows, int32_t sbox,
                       zk_dptr* out_trace) {
  if (!c || !out_trace || (n_events && !host && !dev)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (!pow2(rows) || n_events > rows) return zk_fail(ZK_ERR_ARG, "rows must be a power of two >= n_events");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "tracegen");
  DevScope scope(c);
  const uint32_t* ev = nullptr;
  int32_t rc = stage_events(c, scope, host, dev, n_events * 16, &ev);
  if (rc) return rc;
  const uint32_t w = sbox ? tg::P2W_WIDTH_SBOX : tg::P2W_WIDTH_NO_SBOX;
  uint32_t* out = nullptr;
  if ((rc = scope.alloc(&out, rows * w * 4ull))) return rc;
  unsigned grid = (unsigned)((rows + tg::P2W_ROWS - 1) / tg::P2W_ROWS);
  if (sbox) ZK_LAUNCH_COOP(tg::poseidon2_wide_rows<true>, grid, tg::P2W_ROWS, 0, c->stream, ev, n_events, rows, out);
  else ZK_LAUNCH_COOP(tg::poseidon2_wide_rows<false>, grid, tg::P2W_ROWS, 0, c->stream, ev, n_events, rows, out);
  CK(cudaGetLastError());
  c->launches++;
  scope.release(out);
  *out_trace = (zk_dptr)out;
  return ZK_OK;
}

int32_t poseidon2_skinny(zk_ctx* c, const uint32_t* host, zk_dptr dev, uint64_t n_events, uint64_t rows, zk_dptr* out_trace) {
  if (!c || !out_trace || (n_events && !host && !dev)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (!pow2(rows) || n_events * tg::P2S_ROWS_PER_EVENT > rows)
    return zk_fail(ZK_ERR_ARG, "rows must be a power of two >= 11 * n_events");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "tracegen");
  DevScope scope(c);
  const uint32_t* ev = nullptr;
  int32_t rc = stage_events(c, scope, host, dev, n_events * 16, &ev);
  if (rc) return rc;
  uint32_t* out = nullptr;
  if ((rc = scope.alloc(&out, rows * tg::P2S_WIDTH * 4ull))) return rc;
  const uint64_t per_cta = tg::P2S_EVENTS * tg::P2S_ROWS_PER_EVENT;
  ZK_LAUNCH_COOP(tg::poseidon2_skinny_rows, (unsigned)((rows + per_cta - 1) / per_cta), tg::P2S_EVENTS, 0, c->stream, ev,
                 n_events, rows, out);
  CK(cudaGetLastError());
  c->launches++;
  scope.release(out);
  *out_trace = (zk_dptr)out;
  return ZK_OK;
}

int32_t cpu_trace(zk_ctx* c, const uint32_t* host, zk_dptr dev, uint64_t n_events, uint64_t rows, zk_dptr* out_trace) {
  if (!c || !out_trace || (n_events && !host && !dev)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (!pow2(rows) || n_events > rows) return zk_fail(ZK_ERR_ARG, "rows must be a power of two >= n_events");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "tracegen");
  DevScope scope(c);
  const uint32_t* ev = nullptr;
  int32_t rc = stage_events(c, scope, host, dev, n_events * tg::CPU_EV_WORDS, &ev);
  if (rc) return rc;
  uint32_t* out = nullptr;
  if ((rc = scope.alloc(&out, rows * tg::CPU_W * 4ull))) return rc;
  ZK_LAUNCH_COOP(tg::cpu_rows, (unsigned)((rows + tg::ROWS - 1) / tg::ROWS), tg::ROWS, 0, c->stream, ev, n_events, rows, out);
  CK(cudaGetLastError());
  c->launches++;
  scope.release(out);
  *out_trace = (zk_dptr)out;
  return ZK_OK;
}

template <class CHIP>
void launch_alu(zk_ctx* c, const uint32_t* ev, uint64_t n_events, uint64_t rows, uint32_t* out) {
  unsigned grid = (unsigned)((rows + tg::ROWS - 1) / tg::ROWS);
  ZK_LAUNCH_COOP(tg::alu_rows<CHIP>, grid, tg::ROWS, 0, c->stream, ev, n_events, rows, out);
}

int32_t alu(zk_ctx* c, int32_t chip, const uint32_t* host, zk_dptr dev, uint64_t n_events, uint64_t rows,
            zk_dptr* out_trace) {
  if (!c || !out_trace || (n_events && !host && !dev)) return zk_fail(ZK_ERR_ARG, "null argument");
  uint32_t w = zk_tracegen_alu_width(chip);
  if (!w) return zk_fail(ZK_ERR_ARG, "unknown chip");
  if (!pow2(rows) || n_events > rows) return zk_fail(ZK_ERR_ARG, "rows must be a power of two >= n_events");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "tracegen");
  DevScope scope(c);
  const uint32_t* ev = nullptr;
  int32_t rc = stage_events(c, scope, host, dev, n_events * 7, &ev);
  if (rc) return rc;
  uint32_t* out = nullptr;
  if ((rc = scope.alloc(&out, rows * w * 4ull))) return rc;
  switch (chip) {
    case ZK_CHIP_ADD_SUB: launch_alu<tg::AddSub>(c, ev, n_events, rows, out); break;
    case ZK_CHIP_BITWISE: launch_alu<tg::Bitwise>(c, ev, n_events, rows, out); break;
    case ZK_CHIP_SHIFT_LEFT: launch_alu<tg::ShiftLeft>(c, ev, n_events, rows, out); break;
    case ZK_CHIP_SHIFT_RIGHT: launch_alu<tg::ShiftRight>(c, ev, n_events, rows, out); break;
    case ZK_CHIP_CLO_CLZ: launch_alu<tg::CloClz>(c, ev, n_events, rows, out); break;
    default: launch_alu<tg::Lt>(c, ev, n_events, rows, out); break;
  }
  CK(cudaGetLastError());
  c->launches++;
  scope.release(out);
  *out_trace = (zk_dptr)out;
  return ZK_OK;
}

}  // namespace

extern "C" uint32_t zk_tracegen_alu_width(int32_t chip) {
  switch (chip) {
    case ZK_CHIP_ADD_SUB: return tg::AddSub::W;
    case ZK_CHIP_BITWISE: return tg::Bitwise::W;
    case ZK_CHIP_LT: return tg::Lt::W;
    case ZK_CHIP_SHIFT_LEFT: return tg::ShiftLeft::W;
    case ZK_CHIP_SHIFT_RIGHT: return tg::ShiftRight::W;
    case ZK_CHIP_CLO_CLZ: return tg::CloClz::W;
    default: return 0;
  }
}
extern "C" uint32_t zk_tracegen_poseidon2_wide_width(int32_t sbox_state) {
  return sbox_state ? tg::P2W_WIDTH_SBOX : tg::P2W_WIDTH_NO_SBOX;
}
extern "C" int32_t zk_tracegen_alu(zk_ctx* c, int32_t chip, const zk_alu_event* events_host, uint64_t n_events,
                                   uint64_t rows, zk_dptr* out_trace) {
  static_assert(sizeof(zk_alu_event) == 28, "AluEvent is 7 words");
  return alu(c, chip, (const uint32_t*)events_host, 0, n_events, rows, out_trace);
}
extern "C" int32_t zk_tracegen_alu_dev(zk_ctx* c, int32_t chip, zk_dptr events_dev, uint64_t n_events, uint64_t rows,
                                       zk_dptr* out_trace) {
  return alu(c, chip, nullptr, events_dev, n_events, rows, out_trace);
}
extern "C" int32_t zk_tracegen_poseidon2_wide(zk_ctx* c, const uint32_t* inputs_host, uint64_t n_events, uint64_t rows,
                                              int32_t sbox_state, zk_dptr* out_trace) {
  return poseidon2_wide(c, inputs_host, 0, n_events, rows, sbox_state, out_trace);
}
extern "C" int32_t zk_tracegen_poseidon2_wide_dev(zk_ctx* c, zk_dptr inputs_dev, uint64_t n_events, uint64_t rows,
                                                  int32_t sbox_state, zk_dptr* out_trace) {
  return poseidon2_wide(c, nullptr, inputs_dev, n_events, rows, sbox_state, out_trace);
}
extern "C" int32_t zk_tracegen_poseidon2_wide_prep(zk_ctx* c, const uint32_t* instrs_host, uint64_t n, uint64_t rows,
                                                   zk_dptr* out_trace) {
  if (!c || !out_trace || (n && !instrs_host)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (!pow2(rows) || n > rows) return zk_fail(ZK_ERR_ARG, "rows must be a power of two >= n");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "tracegen");
  DevScope scope(c);
  const uint32_t* in = nullptr;
  int32_t rc = stage_events(c, scope, instrs_host, 0, n * 48, &in);
  if (rc) return rc;
  uint32_t* out = nullptr;
  if ((rc = scope.alloc(&out, rows * 49 * 4ull))) return rc;
  ZK_LAUNCH(tg::poseidon2_wide_prep_rows, (unsigned)((rows * 49 + 255) / 256), 256, 0, c->stream, in, n, rows, out);
  CK(cudaGetLastError());
  c->launches++;
  scope.release(out);
  *out_trace = (zk_dptr)out;
  return ZK_OK;
}
extern "C" uint32_t zk_tracegen_poseidon2_skinny_width(void) { return tg::P2S_WIDTH; }
extern "C" int32_t zk_tracegen_poseidon2_skinny(zk_ctx* c, const uint32_t* inputs_host, uint64_t n_events, uint64_t rows,
                                                zk_dptr* out_trace) {
  return poseidon2_skinny(c, inputs_host, 0, n_events, rows, out_trace);
}
extern "C" int32_t zk_tracegen_poseidon2_skinny_dev(zk_ctx* c, zk_dptr inputs_dev, uint64_t n_events, uint64_t rows,
                                                    zk_dptr* out_trace) {
  return poseidon2_skinny(c, nullptr, inputs_dev, n_events, rows, out_trace);
}
extern "C" int32_t zk_tracegen_poseidon2_skinny_prep(zk_ctx* c, const uint32_t* instrs_host, uint64_t n, uint64_t rows,
                                                     zk_dptr* out_trace) {
  if (!c || !out_trace || (n && !instrs_host)) return zk_fail(ZK_ERR_ARG, "null argument");
  if (!pow2(rows) || n * 11 > rows) return zk_fail(ZK_ERR_ARG, "rows must be a power of two >= 11 * n");
  std::lock_guard<std::mutex> g(c->mu);
  CK(cudaSetDevice(c->device));
  ProfScope ps(c, "tracegen");
  DevScope scope(c);
  const uint32_t* in = nullptr;
  int32_t rc = stage_events(c, scope, instrs_host, 0, n * 48, &in);
  if (rc) return rc;
  uint32_t* out = nullptr;
  if ((rc = scope.alloc(&out, rows * 51 * 4ull))) return rc;
  ZK_LAUNCH(tg::poseidon2_skinny_prep_rows, (unsigned)((rows * 51 + 255) / 256), 256, 0, c->stream, in, n, rows, out);
  CK(cudaGetLastError());
  c->launches++;
  scope.release(out);
  *out_trace = (zk_dptr)out;
  return ZK_OK;
}
extern "C" uint32_t zk_tracegen_cpu_width(void) { return tg::CPU_W; }
extern "C" int32_t zk_tracegen_cpu(zk_ctx* c, const zk_cpu_event* events_host, uint64_t n_events, uint64_t rows,
                                   zk_dptr* out_trace) {
  static_assert(sizeof(zk_cpu_event) == 4 * tg::CPU_EV_WORDS, "zk_cpu_event is 22 words");
  return cpu_trace(c, (const uint32_t*)events_host, 0, n_events, rows, out_trace);
}
extern "C" int32_t zk_tracegen_cpu_dev(zk_ctx* c, zk_dptr events_dev, uint64_t n_events, uint64_t rows, zk_dptr* out_trace) {
  return cpu_trace(c, nullptr, events_dev, n_events, rows, out_trace);
}
