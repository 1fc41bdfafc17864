// Host stand-ins for what generated bodies use from csrc/kernels/batched_kernel.cuh, so that the text
// the model compiler emits (grbda_cuda_emit_source) can be compiled with g++ and executed on the CPU:
// one "thread", one state. Shared-memory rows become plain arrays (the parked variant reads AND writes
// them, as on the device), the chunk staging buffers are flushed into the output rows.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
static inline void __syncwarp() {}
using std::max;
using std::min;

namespace host_body
{
    constexpr int OUT_CHUNK = 16;
    template <typename real>
    struct OutStage
    {
        real *lane;
        real *warp;
        real *g[3];
        int valid;
        int zero;
        int buf_stride;
        int cls[3]; // vector-store bodies: alignment class of the thread's output rows (state N_OUTk mod 4)
    };
    template <typename real>
    inline void storeRow4(real *p, real a, real b, real c, real d) { p[0] = a, p[1] = b, p[2] = c, p[3] = d; }
    template <typename real>
    inline void storeRow1(real *p, real a) { p[0] = a; }
    template <typename real, int N, int COUNT>
    inline void flushChunk(real *g, int base, const real *stg, int valid)
    {
        for (int st = 0; st < valid; st++)
            for (int el = 0; el < COUNT; el++)
                g[(size_t)st * N + base + el] = stg[st * (OUT_CHUNK + 1) + el];
    }
    template <bool FAST>
    inline void grbda_sincos(double x, double *s, double *c)
    {
        *s = std::sin(x);
        *c = std::cos(x);
    }
    template <bool FAST>
    inline void grbda_sincos(float x, float *s, float *c)
    {
        *s = std::sin(x);
        *c = std::cos(x);
    }
    template <bool FAST, typename real>
    inline real grbda_sin(real x) { return std::sin(x); }
    template <bool FAST, typename real>
    inline real grbda_cos(real x) { return std::cos(x); }
    inline unsigned absKey(double x)
    {
        uint64_t u;
        std::memcpy(&u, &x, 8);
        return (unsigned)(u >> 32) & 0x7fffffffu;
    }
    inline unsigned absKey(float x)
    {
        uint32_t u;
        std::memcpy(&u, &x, 4);
        return u & 0x7fffffffu;
    }
    template <typename real>
    inline real pinAfter(real x, real, int) { return x; }
} // namespace host_body
#define GRBDA_DIV(a, b) ((a) / (b))
#define GRBDA_PIN_IMPL(x, late, zero) host_body::pinAfter(x, late, zero)
