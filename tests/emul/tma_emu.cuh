// tma_emu.cuh -- TEST INFRASTRUCTURE ONLY: what csrc/tma.cuh means, for the SIMT emulator.
// An mbarrier is modelled by the number of phases it has completed; a bulk copy is a memcpy that
// completes one phase.  try_wait.parity P succeeds once the phase with parity P is over.
#pragma once
#include "simt_emu.h"

namespace h3dgs {
inline void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
inline void fence_mbar_init() {}
inline void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
inline void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    if (((uintptr_t)smem_dst | (uintptr_t)gmem_src | bytes) & 15) { fprintf(stderr, "tma_load_1d: operands must be 16-byte aligned\n"); abort(); }
    memcpy(smem_dst, gmem_src, bytes);
    (*bar)++;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (((*bar) & 1u) == parity) simt::fiber_yield();
}
}  // namespace h3dgs
