// TEST INFRASTRUCTURE ONLY -- host emulation of the CUDA subset used by neuman_b200/csrc/human_train_kernels.cuh
// (one thread per element, atomicAdd scatter, no shared memory, no warp intrinsics): the kernel bodies are compiled by
// g++ unchanged and executed serially, block by block and thread by thread, so that their algebra can be checked against
// torch autograd on machines without a GPU (tests/test_human_train_emu.py).  Nothing in the product uses this.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define NM_EMU 1
#define NM_KERNEL static
#define NM_DEV static inline
#define NM_ATOMIC_ADD(p, v) (*(p) += (v))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 blockIdx, threadIdx, blockDim, gridDim;

// runs kernel(args...) for every (block, thread) of a 1-D launch
#define EMU_LAUNCH(kernel, grid, block, ...)                    \
  do {                                                          \
    gridDim = {(unsigned)(grid), 1, 1};                         \
    blockDim = {(unsigned)(block), 1, 1};                       \
    for (unsigned _b = 0; _b < (unsigned)(grid); ++_b)          \
      for (unsigned _t = 0; _t < (unsigned)(block); ++_t) {     \
        blockIdx = {_b, 0, 0};                                  \
        threadIdx = {_t, 0, 0};                                 \
        kernel(__VA_ARGS__);                                    \
      }                                                         \
  } while (0)
