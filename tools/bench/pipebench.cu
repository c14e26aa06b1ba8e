// tools/bench/pipebench.cu -- issue-rate microbenchmark of the integer instructions the Poseidon2 / NTT kernels
// are made of (IADD3, VIADDMNMX, IMAD, IMAD.WIDE, IMAD.HI, LOP3, SHF) and of two-pipe mixes, on sm_100a.
// Reports cycles per warp instruction per SM sub-partition at 1 and 8 warps per sub-partition, i.e. the cost
// model the instruction-mix decisions in csrc/poseidon2.cuh are made against.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bench/pipebench tools/bench/pipebench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr uint32_t P = 0x7f000001u;
constexpr int CH = 8;      // independent chains per thread
constexpr int UNROLL = 16; // ops per chain per loop iteration

enum Op { IADD3, VMNMX, IMAD, WIDE, HI, HIS, LOP, SHF, MIX_ADD_MAD, MIX_MNMX_MAD, MIX_MONT, MIX_ADD_MNMX, MODADD, MONTMUL,
          DFMA, DADD, MIX_DFMA_IMAD, MIX_DFMA_IADD, MIX_DFMA_IMAD_IADD, MIX_DFMA_WIDE, FFMA, MIX_FFMA_IMAD, MIX_FFMA_IMAD_IADD };

template <int OP>
__device__ __forceinline__ void step(uint32_t& x, uint32_t& y, uint32_t k) {
  if (OP == IADD3) {
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(x) : "r"(y), "r"(k));
  } else if (OP == VMNMX) {
    uint32_t t;
    asm volatile("add.u32 %0, %1, %2;\n\tmin.u32 %1, %1, %0;" : "=r"(t), "+r"(x) : "r"(k));
  } else if (OP == IMAD) {
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(k));
  } else if (OP == WIDE) {
    uint64_t w;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(y));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(x), "=r"(y) : "l"(w));
  } else if (OP == HI) {
    asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(k));
  } else if (OP == HIS) {
    asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(x) : "r"(k));
  } else if (OP == LOP) {
    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(k));
  } else if (OP == SHF) {
    asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y));
  } else if (OP == MIX_ADD_MAD) {  // one alu + one fma instruction
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(x) : "r"(y), "r"(k));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y) : "r"(k), "r"(k));
  } else if (OP == MIX_MNMX_MAD) {
    uint32_t t;
    asm volatile("add.u32 %0, %1, %2;\n\tmin.u32 %1, %1, %0;" : "=r"(t), "+r"(x) : "r"(k));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y) : "r"(k), "r"(k));
  } else if (OP == MIX_ADD_MNMX) {  // two alu instructions = a modular add whose add cannot move to the fma pipe
    uint32_t t;
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(x) : "r"(y), "r"(k));
    asm volatile("add.u32 %0, %1, %2;\n\tmin.u32 %1, %1, %0;" : "=r"(t), "+r"(x) : "r"(k));
  } else if (OP == DFMA || OP == DADD || OP == MIX_DFMA_IMAD || OP == MIX_DFMA_IADD || OP == MIX_DFMA_IMAD_IADD || OP == MIX_DFMA_WIDE) {
    // the double lives in the (x, y) register pair of the chain for the pure tests; the mixes keep a separate integer op
    // on k-derived values so that the FP64 pipe and the integer pipes run side by side
    double d = __hiloint2double((int)y, (int)x);
    if (OP == DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(d) : "d"(1.5));
    else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5));
    x = (uint32_t)__double2loint(d); y = (uint32_t)__double2hiint(d);
  } else if (OP == FFMA || OP == MIX_FFMA_IMAD || OP == MIX_FFMA_IMAD_IADD) {
    float f = __uint_as_float(x);
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0000001f), "f"(0.5f));
    x = __float_as_uint(f);
  } else if (OP == MODADD) {  // what the compiler makes of kb::add (it may pick IMAD.IADD)
    uint32_t s = x + y;
    x = min(s, s - P);
  } else if (OP == MONTMUL || OP == MIX_MONT) {
    uint64_t t = (uint64_t)x * y;
    uint32_t m = (uint32_t)t * 0x81000001u;
    uint32_t u = __umulhi(m, P);
    uint32_t r = (uint32_t)(t >> 32) - u;
    x = min(r, r + P);
    if (OP == MIX_MONT) {  // plus three modular adds, roughly the Poseidon2 external-round mix
      uint32_t s = y + k; y = min(s, s - P);
      s = y + x; y = min(s, s - P);
      s = y + k; y = min(s, s - P);
    }
  }
}
// integer companion of the FP64 / FP32 mixes, on its own registers (z, u)
template <int OP>
__device__ __forceinline__ void step2(uint32_t& z, uint32_t& u, uint32_t k) {
  if (OP == MIX_DFMA_IMAD || OP == MIX_DFMA_IMAD_IADD || OP == MIX_FFMA_IMAD || OP == MIX_FFMA_IMAD_IADD)
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z) : "r"(k), "r"(k));
  if (OP == MIX_DFMA_IADD || OP == MIX_DFMA_IMAD_IADD || OP == MIX_FFMA_IMAD_IADD)
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(u) : "r"(k), "r"(k));
  if (OP == MIX_DFMA_WIDE) {
    uint64_t w;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(z), "r"(u));
    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(z), "=r"(u) : "l"(w));
  }
}
template <int OP> constexpr int sass_per_step() {
  if (OP == MIX_DFMA_IMAD || OP == MIX_DFMA_IADD || OP == MIX_DFMA_WIDE || OP == MIX_FFMA_IMAD) return 2;
  if (OP == MIX_DFMA_IMAD_IADD || OP == MIX_FFMA_IMAD_IADD) return 3;
  return OP == MIX_ADD_MAD || OP == MIX_MNMX_MAD || OP == MIX_ADD_MNMX || OP == MODADD ? 2 : OP == MONTMUL ? 5 : OP == MIX_MONT ? 11 : 1;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cyc, int iters, const uint32_t* kp) {
  const uint32_t k = kp[0];  // in a register, not a constant-bank operand
  uint32_t x[CH], y[CH], z[CH], u[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { z[c] = threadIdx.x + 77u * c; u[c] = blockIdx.x + 3u * c; }
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = threadIdx.x * 2654435761u + c * 40503u + 1; y[c] = (blockIdx.x + c) * 977u + 3; if (OP == MIX_MONT || OP == MODADD || OP == MONTMUL) { x[c] %= P; y[c] %= P; } if (OP >= DFMA && OP <= MIX_DFMA_WIDE) { y[c] = 0x40000000u | (y[c] & 0xfffffu); } }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int un = 0; un < UNROLL; un++)
#pragma unroll
      for (int c = 0; c < CH; c++) { step<OP>(x[c], y[c], k); step2<OP>(z[c], u[c], k); }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) acc ^= x[c] + y[c] + z[c] + u[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, uint32_t* d_out, long long* d_cyc, int sms) {
  const int iters = 256;
  printf("%-34s", name);
  for (int warps_per_smsp : {1, 2, 4, 8}) {
    int threads = 32 * 4 * warps_per_smsp;
    bench<OP><<<sms, threads>>>(d_out, d_cyc, 4, d_out + sms * 1024);
    bench<OP><<<sms, threads>>>(d_out, d_cyc, iters, d_out + sms * 1024);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    double steps = (double)iters * UNROLL * CH * warps_per_smsp;  // warp-steps per SMSP
    printf("  w%d: %6.3f clk/step (%5.3f clk/instr)", warps_per_smsp, c / steps, c / steps / sass_per_step<OP>());
  }
  printf("  [%d SASS/step nominal] %s\n", sass_per_step<OP>(), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs; clk = SM cycles per warp-level step per sub-partition\n", p.name, sms);
  uint32_t* d; cudaMalloc(&d, (size_t)sms * 1024 * 4 + 4);
  { uint32_t kv = P - 7; cudaMemcpy(d + sms * 1024, &kv, 4, cudaMemcpyHostToDevice); }
  long long* c; cudaMalloc(&c, sms * 8);
  run<IADD3>("IADD3 (3 regs)", d, c, sms);
  run<VMNMX>("VIADDMNMX.U32 (min(x, x+imm))", d, c, sms);
  run<IMAD>("IMAD (lo, 3 regs)", d, c, sms);
  run<WIDE>("IMAD.WIDE.U32", d, c, sms);
  run<HI>("IMAD.HI.U32", d, c, sms);
  run<HIS>("IMAD.HI (signed)", d, c, sms);
  run<LOP>("LOP3", d, c, sms);
  run<SHF>("SHF", d, c, sms);
  run<MIX_ADD_MAD>("IADD3 + IMAD", d, c, sms);
  run<MIX_MNMX_MAD>("VIADDMNMX + IMAD", d, c, sms);
  run<MIX_ADD_MNMX>("IADD3 + VIADDMNMX", d, c, sms);
  run<MODADD>("kb::add as compiled", d, c, sms);
  run<MONTMUL>("kb::mul as compiled", d, c, sms);
  run<MIX_MONT>("kb::mul + 3 kb::add", d, c, sms);
  run<DFMA>("DFMA", d, c, sms);
  run<DADD>("DADD", d, c, sms);
  run<MIX_DFMA_IMAD>("DFMA + IMAD", d, c, sms);
  run<MIX_DFMA_IADD>("DFMA + IADD3", d, c, sms);
  run<MIX_DFMA_IMAD_IADD>("DFMA + IMAD + IADD3", d, c, sms);
  run<MIX_DFMA_WIDE>("DFMA + IMAD.WIDE", d, c, sms);
  run<FFMA>("FFMA", d, c, sms);
  run<MIX_FFMA_IMAD>("FFMA + IMAD", d, c, sms);
  run<MIX_FFMA_IMAD_IADD>("FFMA + IMAD + IADD3", d, c, sms);
  return 0;
}
