// launch.cuh -- kernel launch macros.  The product build (nvcc) expands them to <<< >>> launches; the
// test-only emulator build (tests/emu, g++) runs the kernel body on the CPU to debug index math.
#pragma once
#ifdef ZK_EMU
#define ZK_LAUNCH(kern, grid, block, smem, stream, ...) \
  zkemu::launch((unsigned)(grid), (unsigned)(block), [&]() { kern(__VA_ARGS__); }, false)
#define ZK_LAUNCH_COOP(kern, grid, block, smem, stream, ...) \
  zkemu::launch((unsigned)(grid), (unsigned)(block), [&]() { kern(__VA_ARGS__); }, true)
#define ZK_DYN_SMEM(name) static uint32_t name[60 * 1024]
#else
#define ZK_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ZK_LAUNCH_COOP(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ZK_DYN_SMEM(name) extern __shared__ uint32_t name[]
#endif
