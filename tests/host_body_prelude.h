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
        struct Ring
        {
            real slots[8];
            real *put, *wrap[3], *ga;
            int m;
            void init(real *row, int m_)
            {
                m = m_, put = slots + m, ga = row - m;
                for (int j = 0; j < 3; j++)
                    wrap[j] = slots + ((5 + j + m) & 7);
            }
        };
        mutable Ring ring[3]; // ring-store bodies (the device keeps the slots in shared memory)
        real *park = nullptr; // park area of a parked body (Body::PARK_EXTRA slots)
    };
    template <int E, typename R, typename real>
    inline void ringPut(R &r, real x)
    {
        if ((E & 7) < 5)
            r.put[E & 7] = x;
        else
            *r.wrap[(E & 7) - 5] = x;
    }
    template <typename real, int N0, typename R>
    inline void ringFlush(R &r, real *)
    {
        for (int i = 0; i < 4; i++)
            r.ga[4 * N0 + i] = r.slots[((4 * N0) & 7) + i];
    }
    template <typename real, typename R>
    inline void ringHead(R &r, real *)
    {
        for (int p = r.m; p < 4; p++)
            r.ga[p] = r.slots[p];
    }
    template <typename real, int N, typename R>
    inline void ringTail(R &r, real *)
    {
        for (int p = 4 * (N / 4); p < r.m + N; p++)
            r.ga[p] = r.slots[p & 7];
    }
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
