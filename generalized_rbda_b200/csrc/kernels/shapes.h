// Shared-memory layouts of the kernel shells as functions of plain sizes. Plain C++ (no CUDA constructs)
// so that the device code (kernels/batched_kernel.cuh, also under NVRTC), the ahead-of-time tool
// (tools/modelc.cpp) and the run-time compiler (runtime/jit.cpp) size tiles with the same code.
#ifndef GRBDA_KERNELS_SHAPES_H // (NVRTC sees this header under two include names: #pragma once is not enough)
#define GRBDA_KERNELS_SHAPES_H
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define GRBDA_HD __host__ __device__
#else
#include <stddef.h>
#define GRBDA_HD
#endif

namespace grbda_kernels
{
    // values per chunk of the chunked output staging (kernels with large outputs: FK, H)
    constexpr int OUT_CHUNK = 16;
    // largest software-staged tile set the flagged-tile pass keeps in shared memory; beyond it the pass uses
    // direct global I/O, and a body of such a program is never parked (it would write into caller memory)
    constexpr size_t SLOW_PASS_STAGED_LIMIT = 200 * 1024;
    // shared memory one SM can give to resident CTAs (227 KB opt-in per CTA, 1 KB reserved per CTA)
    constexpr size_t SM_SHARED_BYTES = 228 * 1024;

    // Row stride (in elements) of a staged tile: odd, so that thread t reading element i of its
    // own state (address t * stride + i) hits 32 distinct banks for 4-byte and 16 distinct bank
    // pairs per half warp for 8-byte elements.
    GRBDA_HD constexpr int oddStride(int n) { return n | 1; }

    // Output array k of a program leaves through the large-output path (256-bit stores of whole sectors of the
    // thread's own row, or chunk staging) rather than as a staged row / plain stores. The emitter
    // (compiler/emit.h) and the host-side alignment check (runtime/capi.cu) share this rule.
    GRBDA_HD constexpr bool shapeLargeOutput(int k, int n) { return n > 64 || (k > 0 && n > OUT_CHUNK); }

    // elem = sizeof(real); the static_asserts in the kernels keep these in step with TileLayout / TmaLayout
    GRBDA_HD constexpr size_t shapeAlign16(size_t x) { return (x + 15) & ~(size_t)15; }
    GRBDA_HD constexpr size_t shapeChunkStageBytes(const int *n_out, int stage_buffers, int block, int elem)
    {
        return (n_out[0] > 64 || n_out[1] > OUT_CHUNK || n_out[2] > OUT_CHUNK)
                   ? (size_t)stage_buffers * block * (OUT_CHUNK + 1) * elem
                   : 0;
    }
    // A parked body may own `park_extra` more slots per thread than its tile rows offer (kernels whose tiles leave
    // shared memory unused: the mass matrix has one input row and holds its results until their chunk is
    // complete). Odd stride: conflict-free for 4- and 8-byte elements.
    GRBDA_HD constexpr size_t shapeParkBytes(int park_extra, int block, int elem)
    {
        return park_extra > 0 ? (size_t)block * oddStride(park_extra) * elem : 0;
    }
    // everything behind the tiles: chunk staging buffers / rings, then the park area
    GRBDA_HD constexpr size_t shapeStageBytes(const int *n_out, int stage_buffers, int block, int elem, int park_extra = 0)
    {
        return shapeChunkStageBytes(n_out, stage_buffers, block, elem) + shapeParkBytes(park_extra, block, elem);
    }
    // direct_out: the (small) output 0 is not staged as a tile but stored by every thread straight to its own row in
    // global memory (models whose rows otherwise keep an SM from holding its CTAs: JVRC1)
    GRBDA_HD constexpr size_t shapeTileBytes(const int *n_in, const int *n_out, int stage_buffers, int block,
                                                        int elem, int park_extra = 0, bool direct_out = false)
    {
        size_t elems = 0;
        for (int i = 0; i < 3; i++)
            elems += n_in[i] ? (size_t)oddStride(n_in[i]) * block : 0;
        elems += n_out[0] <= 64 && !direct_out ? (size_t)oddStride(n_out[0]) * block : 0;
        return shapeAlign16(elems * elem) + shapeStageBytes(n_out, stage_buffers, block, elem, park_extra);
    }
    GRBDA_HD constexpr int shapeTmaStride(int n, int elem)
    {
        return (n >= 12 && (n * elem) % 16 == 0) ? n + 16 / elem : n;
    }
    GRBDA_HD constexpr size_t shapeTmaBytes(const int *n_in, const int *n_out, int stage_buffers, int block,
                                                       int elem, int park_extra = 0, bool direct_out = false)
    {
        size_t off = 16;
        for (int i = 0; i < 3; i++)
            off = shapeAlign16(off + (size_t)shapeTmaStride(n_in[i], elem) * block * elem);
        off = shapeAlign16(off + (n_out[0] <= 64 && !direct_out ? (size_t)shapeTmaStride(n_out[0], elem) * block * elem : 0));
        return off + shapeStageBytes(n_out, stage_buffers, block, elem, park_extra);
    }
} // namespace grbda_kernels
#endif // GRBDA_KERNELS_SHAPES_H
