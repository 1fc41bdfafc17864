// Hand-written kernel shell around the per-model generated algorithm bodies (sm_100a).
//
// Mapping: ONE STATE PER THREAD. The generated body is straight-line FP64 (or FP32) code whose
// control flow is identical for every thread of the grid: topology, cluster-joint types and the
// loop-closure structure were resolved when the model was compiled, so there is no divergence.
// The 6x6 / 6n x 6n spatial algebra lives in registers; values that stay live across the
// backward / forward sweeps of ABA are kept in an explicit per-thread scratch (local memory,
// interleaved by the hardware so that a warp's access to one slot is two 128-byte lines).
//
// Data movement: states are stored contiguously per state (x[b * n + i], the layout of the
// reference's setState vectors), so a warp's states form ONE contiguous span of 32 * n values.
// The CTA's tile of every input array is streamed from HBM with coalesced 128-bit loads into
// shared memory (stage_in), the threads then read their own state from shared memory with an odd
// stride (bank-conflict free), and results go back the same way (stage_out).
#ifndef GRBDA_KERNELS_BATCHED_KERNEL_CUH // (NVRTC sees this header under two include names: #pragma once is not enough)
#define GRBDA_KERNELS_BATCHED_KERNEL_CUH
// The device part of this header is also compiled at run time by NVRTC (runtime/jit.cpp: models that were
// not compiled ahead of time), which has no host or standard-library headers: everything outside the
// `#ifndef __CUDACC_RTC__` blocks must stay free of them.
#ifdef __CUDACC_RTC__
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef unsigned long long uintptr_t;
namespace grbda_std
{
    template <bool B, class T, class F> struct conditional { typedef T type; };
    template <class T, class F> struct conditional<false, T, F> { typedef F type; };
}
#else
#include <algorithm>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
namespace grbda_std
{
    using std::conditional;
}
#endif

#include "shapes.h"

namespace grbda_kernels
{
#ifndef __CUDACC_RTC__
    struct LaunchArgs
    {
        const void *in[3];
        void *out[3];
        int64_t batch;
        cudaStream_t stream;
        int *launched = nullptr; // optional: incremented by the number of kernels enqueued
        // scratch for the per-tile flags of the sin/cos range decision: at least ceil(batch / 32) bytes,
        // owned by the caller, private to `stream` (launches on one stream are ordered)
        unsigned char *flags = nullptr;
        size_t flags_bytes = 0;
    };
#endif

    // ---- scalar math of the generated bodies -------------------------------------------------------
    // The generated per-state program is ONE basic block of several thousand FP64 instructions. Any
    // branch inside it (the library sincos and the IEEE division both carry a slow-path call) cuts it
    // into pieces: ptxas then schedules, hoists loads and allocates registers per piece, and every
    // branch drains the instruction buffer. So the body uses branch-free forms only, and the one range
    // decision is hoisted in front of it: Body::inRange() (generated: every input that feeds a sin/cos
    // argument is bounded so that each argument stays below the limit of the fast reduction) is
    // evaluated once per CTA tile. The FAST kernel runs run<real, true> on tiles that pass and records
    // the others in a per-launch flag array; a second, small launch (run<real, false>: library sincos)
    // scans the flags and recomputes exactly those tiles. Keeping the two bodies in different kernels
    // matters: one kernel holding both (or calling the slow one) compiles to a much worse fast path.
    //
    // FP64 sincos, fast form: Cody-Waite reduction by pi/2 with three FMA steps (pi/2 split into three
    // doubles, 159 bits), then the fdlibm kernel polynomials on [-pi/4, pi/4]. Every step is an FMA, so
    // the products k * pi/2_i are exact and the reduced argument carries an ABSOLUTE error of a few
    // 1e-16 as long as k = round(x 2/pi) is exact in the 2^52 rounding trick and k * 2^-159 is
    // negligible: |x| <= 1e12 leaves orders of magnitude on both (sincosFastLimit). What the fast form
    // gives up against the library beyond ~1e5 is the RELATIVE accuracy of results that are themselves
    // ~1e-16 (arguments within 1e-16 of a multiple of pi/2); the dynamics need absolute accuracy.
    static __constant__ double grbda_sc_k[18] = {
        0.63661977236758138, 6755399441055744.0, -1.5707963267948966, -6.123233995736766e-17,
        1.4973849048591698e-33,
        1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
        -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
        -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
        2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, 1.0e5};
    // largest |argument| the fast reductions are used for (generated inRange() keeps arguments below it)
    template <typename real>
    __host__ __device__ constexpr double sincosFastLimit() { return sizeof(real) == 8 ? 1.0e12 : 1.0e6; }

    template <bool FAST>
    __device__ __forceinline__ void grbda_sincos(double x, double *s, double *c)
    {
        if (!FAST)
        {
            sincos(x, s, c);
            return;
        }
        const double *K = grbda_sc_k; // constant-bank operands: no instruction spent on literals
        const double t = fma(x, K[0], K[1]); // round-to-nearest-integer trick (1.5 * 2^52)
        const int k = __double2loint(t);
        const double kd = t - K[1];
        double r = fma(kd, K[2], x);
        r = fma(kd, K[3], r);
        r = fma(kd, K[4], r);
        const double z = r * r;
        double ps = fma(z, K[5], K[6]);
        ps = fma(z, ps, K[7]);
        ps = fma(z, ps, K[8]);
        ps = fma(z, ps, K[9]);
        ps = fma(z, ps, K[10]);
        const double sr = fma(z * r, ps, r);
        double pc = fma(z, K[11], K[12]);
        pc = fma(z, pc, K[13]);
        pc = fma(z, pc, K[14]);
        pc = fma(z, pc, K[15]);
        pc = fma(z, pc, K[16]);
        const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
        const double a = (k & 1) ? cr : sr, b = (k & 1) ? sr : cr;
        // quadrant signs: flip the sign bit with integer logic on the high word (one shift + one XOR per
        // result) instead of two more 64-bit selects
        *s = __hiloint2double(__double2hiint(a) ^ ((k & 2) << 30), __double2loint(a));
        *c = __hiloint2double(__double2hiint(b) ^ (((k + 1) & 2) << 30), __double2loint(b));
    }
    // FP32: the same scheme (three-constant reduction, 72 bits of pi/2; degree-7/8 minimax kernels);
    // absolute error ~1e-7 for |x| <= 1e6 (k < 2^22 in the 2^23 rounding trick). FP32 parity budget: 1e-4.
    template <bool FAST>
    __device__ __forceinline__ void grbda_sincos(float x, float *s, float *c)
    {
        if (!FAST)
        {
            sincosf(x, s, c);
            return;
        }
        const float t = fmaf(x, 0.636619772f, 12582912.0f); // 1.5 * 2^23
        const int k = __float_as_int(t);
        const float kd = t - 12582912.0f;
        float r = fmaf(kd, -1.57079601e+00f, x);
        r = fmaf(kd, -3.13916473e-07f, r);
        r = fmaf(kd, -5.39030253e-15f, r);
        const float z = r * r;
        float ps = fmaf(z, 2.86567956e-6f, -1.98559923e-4f);
        ps = fmaf(z, ps, 8.33338592e-3f);
        ps = fmaf(z, ps, -1.66666672e-1f);
        const float sr = fmaf(z * r, ps, r);
        float pc = fmaf(z, 2.44677067e-5f, -1.38877297e-3f);
        pc = fmaf(z, pc, 4.16666567e-2f);
        pc = fmaf(z, pc, -0.5f);
        const float cr = fmaf(z, pc, 1.0f);
        const float a = (k & 1) ? cr : sr, b = (k & 1) ? sr : cr;
        *s = __int_as_float(__float_as_int(a) ^ ((k & 2) << 30));
        *c = __int_as_float(__float_as_int(b) ^ (((k + 1) & 2) << 30));
    }
    template <bool FAST, typename real>
    __device__ __forceinline__ real grbda_sin(real x) { real s, c; grbda_sincos<FAST>(x, &s, &c); return s; }
    template <bool FAST, typename real>
    __device__ __forceinline__ real grbda_cos(real x) { real s, c; grbda_sincos<FAST>(x, &s, &c); return c; }

    // Scheduling pin. ptxas (and LLVM before it) may move any computation whose operands are ready as
    // far up the straight-line program as it likes; it does so with the sin/cos evaluations (their
    // operands are inputs), computing dozens of them at the top of the kernel and spilling the results
    // until the joint is reached. pinAfter(x, late, zero) returns x, bit for bit, but through an integer
    // operation on `late` (a value computed just before in program order) that the compiler cannot
    // remove, because `zero` is only known to be 0 at run time: the evaluation stays where the
    // depth-first program put it. One LOP3 per pinned value; NaN / inf in `late` do not propagate.
    __device__ __forceinline__ double pinAfter(double x, double late, int zero)
    {
        return __hiloint2double(__double2hiint(x) ^ (__double2hiint(late) & zero), __double2loint(x));
    }
    __device__ __forceinline__ float pinAfter(float x, float late, int zero)
    {
        return __int_as_float(__float_as_int(x) ^ (__float_as_int(late) & zero));
    }
    // |x| as an ordered unsigned key (IEEE: the magnitude order of finite values is the integer order of
    // their high words); NaN and inf give the largest keys. Used by the generated range checks.
    __device__ __forceinline__ unsigned absKey(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }
    __device__ __forceinline__ unsigned absKey(float x) { return __float_as_uint(x) & 0x7fffffffu; }

#ifdef GRBDA_NO_RANGE_CHECK /* experiment switch: fast forms unconditionally */
#define GRBDA_RANGE_CHECKED(Body) false
#else
#define GRBDA_RANGE_CHECKED(Body) Body::RANGE_CHECKED
#endif
    // Division of the generated bodies. FP64 keeps the IEEE division: measured on B200, the bodies are
    // FASTER with it than with the branch-free grbda_div below (forward dynamics 0.93-1.16 ms against
    // 1.46-2.4 ms per 2^20 Tello states) - the slow-path call sites of the 26 divisions cut the body
    // into pieces at the pivots of the factorisation, which keeps ptxas from hoisting work (and its
    // live values) across them. FP32 bodies are faster with the branch-free form.
    // (GRBDA_EXTRA_NVCCFLAGS=-DGRBDA_FAST_DIV / -DGRBDA_NO_PIN: experiment switches of build.py)
#ifdef GRBDA_FAST_DIV
#define GRBDA_DIV(a, b) grbda_div(a, b)
#else
#define GRBDA_DIV(a, b) grbda_div_default(a, b)
#endif
#ifdef GRBDA_NO_PIN
#define GRBDA_PIN_IMPL(x, late, zero) (x)
#else
#define GRBDA_PIN_IMPL(x, late, zero) pinAfter(x, late, zero)
#endif

    // a / b without the slow-path call of the IEEE division: hardware reciprocal seed (MUFU.RCP64H, about
    // 20 good bits), two Newton steps, one residual correction of the quotient. Error <= ~1 ulp for
    // normal operands (the divisors are pivots of positive definite blocks and constraint Jacobians);
    // zero / infinite / denormal divisors give inf or nan like the division would, without trapping.
    __device__ __forceinline__ double grbda_div(double a, double b)
    {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
        double e = fma(-b, r, 1.0);
        r = fma(r, e, r);
        e = fma(-b, r, 1.0);
        r = fma(r, e, r);
        const double q = a * r;
        return fma(fma(-b, q, a), r, q);
    }
    __device__ __forceinline__ double grbda_div_default(double a, double b) { return a / b; }
    __device__ __forceinline__ float grbda_div(float a, float b);
    __device__ __forceinline__ float grbda_div_default(float a, float b) { return grbda_div(a, b); }
    __device__ __forceinline__ float grbda_div(float a, float b)
    {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
        r = fmaf(r, fmaf(-b, r, 1.0f), r);
        const float q = a * r;
        return fmaf(fmaf(-b, q, a), r, q);
    }


    // Cooperative, coalesced copy of the CTA's [rows x N] tile (global, dense) into shared memory
    // rows of stride oddStride(N). 128-bit loads when the tile base is 16-byte aligned.
    template <typename real, int N, int BLOCK>
    __device__ __forceinline__ void stage_in(const real *__restrict__ g, real *__restrict__ s, int rows)
    {
        constexpr int S = oddStride(N);
        constexpr int VEC = 16 / sizeof(real);
        const int total = rows * N;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0)
        {
            using vec_t = typename grbda_std::conditional<sizeof(real) == 8, double2, float4>::type;
            const vec_t *gv = reinterpret_cast<const vec_t *>(g);
            const int nvec = total / VEC;
#pragma unroll 4
            for (int k = threadIdx.x; k < nvec; k += BLOCK)
            {
                const vec_t v = __ldcs(gv + k); // streaming: every byte is read exactly once
                const real *e = reinterpret_cast<const real *>(&v);
                const int base = k * VEC;
#pragma unroll
                for (int j = 0; j < VEC; j++)
                {
                    const int idx = base + j;
                    s[(idx / N) * S + (idx % N)] = e[j];
                }
            }
            for (int idx = nvec * VEC + threadIdx.x; idx < total; idx += BLOCK)
                s[(idx / N) * S + (idx % N)] = g[idx];
        }
        else
        {
            for (int idx = threadIdx.x; idx < total; idx += BLOCK)
                s[(idx / N) * S + (idx % N)] = g[idx];
        }
    }

    // Inverse of stage_in for results.
    template <typename real, int N, int BLOCK>
    __device__ __forceinline__ void stage_out(real *__restrict__ g, const real *__restrict__ s, int rows)
    {
        constexpr int S = oddStride(N);
        constexpr int VEC = 16 / sizeof(real);
        const int total = rows * N;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0)
        {
            using vec_t = typename grbda_std::conditional<sizeof(real) == 8, double2, float4>::type;
            vec_t *gv = reinterpret_cast<vec_t *>(g);
            const int nvec = total / VEC;
#pragma unroll 4
            for (int k = threadIdx.x; k < nvec; k += BLOCK)
            {
                vec_t v;
                real *e = reinterpret_cast<real *>(&v);
                const int base = k * VEC;
#pragma unroll
                for (int j = 0; j < VEC; j++)
                {
                    const int idx = base + j;
                    e[j] = s[(idx / N) * S + (idx % N)];
                }
                __stcs(gv + k, v);
            }
            for (int idx = nvec * VEC + threadIdx.x; idx < total; idx += BLOCK)
                g[idx] = s[(idx / N) * S + (idx % N)];
        }
        else
        {
            for (int idx = threadIdx.x; idx < total; idx += BLOCK)
                g[idx] = s[(idx / N) * S + (idx % N)];
        }
    }

    // ---- programmatic dependent launch ------------------------------------------------------------
    // Every call is two launches (fast kernel + flagged-tile pass) and calls follow each other on one stream
    // (forward dynamics, then inverse dynamics ...). The kernels are launched with the programmatic stream
    // serialization attribute: a kernel's launch (CTA scheduling, parameter set-up) overlaps the tail of its
    // predecessor instead of waiting for its completion. Every kernel waits (griddepcontrol.wait: predecessor
    // complete and its memory visible) before it touches global memory, and releases its successor once its
    // compute is done (griddepcontrol.launch_dependents; an exiting CTA counts as released). A predecessor that is
    // not one of these kernels never releases early: plain stream order. GRBDA_NO_PDL=1 turns the attribute off.
    __device__ __forceinline__ void gridDependencyWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
    __device__ __forceinline__ void gridLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifndef __CUDACC_RTC__
    inline bool usePdl()
    {
        static const bool on = !(std::getenv("GRBDA_NO_PDL") && std::getenv("GRBDA_NO_PDL")[0] == '1');
        return on;
    }
    inline void pdlConfig(cudaLaunchConfig_t &cfg, cudaLaunchAttribute &attr, unsigned grid, unsigned block, size_t smem,
                          cudaStream_t stream)
    {
        cfg = cudaLaunchConfig_t();
        cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = usePdl() ? 1 : 0;
    }
    template <typename real>
    cudaError_t launchShell(void (*kernel)(const real *, const real *, const real *, real *, real *, real *, int64_t, unsigned char *),
                            unsigned grid, unsigned block, size_t smem, cudaStream_t stream, const real *in0, const real *in1,
                            const real *in2, real *out0, real *out1, real *out2, int64_t batch, unsigned char *flags)
    {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr;
        pdlConfig(cfg, attr, grid, block, smem, stream);
        return cudaLaunchKernelEx(&cfg, kernel, in0, in1, in2, out0, out1, out2, batch, flags);
    }
#endif

    // ---- chunked output staging (kernels with large outputs: FK, H) ------------------------------
    // The generated body hands over `COUNT` consecutive elements of one output array for the 32
    // states of the warp through the warp's staging buffer; this writes them with coalesced stores
    // (for COUNT = 16 doubles every instruction covers two states x 128 contiguous bytes).
    // Row rings (Body::RING_STORES, compiler/emit.h row_stores = 2): a thread assembles the sectors of its own output
    // row in 8 shared-memory slots. m = (state N) mod 4 is the position of the row's first element inside the 32-byte
    // sector grid (FP32: the 16-byte grid); element e lives in slot (e + m) mod 8, so quad n of the GLOBAL sector grid
    // is slots 4 (n mod 2) .. + 3 and leaves as one aligned 256-bit store to ga + 4 n, ga = row - m.
    template <typename real>
    struct RowRing
    {
        real *slots;   // 8 slots (16-byte aligned; rows of different threads 80 / 48 bytes apart)
        real *put;     // slots + m: elements with (e mod 8) < 5 are stored at put[e mod 8] (no wrap-around)
        real *ga;      // global address of the sector that holds the row's first element (row - m)
        int m;
    };
    template <typename real>
    struct RingStride
    {
        static constexpr int value = sizeof(real) == 8 ? 10 : 12; // elements: 128-bit loads stay aligned, few bank conflicts
    };
    template <typename real>
    struct OutStage
    {
        real *lane;   // this thread's row of the warp's staging buffer [32][OUT_CHUNK + 1]
        real *warp;   // the warp's staging buffer
        real *g[3];   // output rows of the warp's first state
        int valid;    // number of states of this warp that exist (tail of the batch)
        int zero;     // 0 at run time, unknown at compile time (see pinAfter)
        int buf_stride; // elements between the staging buffers of consecutive output arrays (Body::STAGE_BUFFERS > 1)
        // Body::VECTOR_STORES: position of this thread's row of output k inside the 32-byte sector grid,
        // (state N_OUTk) mod 4 - warp-uniform by the state mapping of the shells (threadState) - or -1 for a thread
        // without a state of its own (tail of the last tile)
        int cls[3];
        RowRing<real> ring[3]; // Body::RING_STORES
        real *park;            // Body::PARK_EXTRA > 0: this thread's slots of the park area
    };
    // 256-bit store of one whole sector of the thread's own output row (SASS STG.E.256), evict-first
    __device__ __forceinline__ void storeRow4(double *p, double a, double b, double c, double d)
    {
#ifdef GRBDA_NO_V256 // compilers before CUDA 12.9 (PTX ISA 8.8): the same sector as two 128-bit halves
        asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n\tst.global.cs.v2.f64 [%0+16], {%3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#endif
    }
    __device__ __forceinline__ void storeRow4(float *p, float a, float b, float c, float d)
    {
        // FP32 rows sit on a 16-byte grid for the same classes: one 128-bit store
        asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    }
    template <typename real>
    __device__ __forceinline__ void storeRow1(real *p, real a) { __stcs(p, a); }
    template <int E, typename real>
    __device__ __forceinline__ void ringPut(const RowRing<real> &r, real x)
    {
        if constexpr ((E & 7) < 5)
            r.put[E & 7] = x;
        else // may wrap around the ring (three elements in eight): two integer instructions instead of a third pointer
            r.slots[((E & 7) + r.m) & 7] = x;
    }
    __device__ __forceinline__ void ringStoreQuad(double *g, const double *s)
    {
        const double2 lo = *reinterpret_cast<const double2 *>(s), hi = *reinterpret_cast<const double2 *>(s + 2);
        storeRow4(g, lo.x, lo.y, hi.x, hi.y);
    }
    __device__ __forceinline__ void ringStoreQuad(float *g, const float *s)
    {
        const float4 v = *reinterpret_cast<const float4 *>(s);
        storeRow4(g, v.x, v.y, v.z, v.w);
    }
    // quad N0 of the sector grid; emitted after element 4 N0 + 3: complete whatever m is
    template <typename real, int N0>
    __device__ __forceinline__ void ringFlush(const RowRing<real> &r, real *)
    {
        ringStoreQuad(r.ga + 4 * N0, r.slots + ((4 * N0) & 7));
    }
    // first quad (after element 3): whole when the row starts on a sector boundary, else its m leading values belong
    // to the previous state's row
    template <typename real>
    __device__ __forceinline__ void ringHead(const RowRing<real> &r, real *)
    {
        if (r.m == 0)
            ringStoreQuad(r.ga, r.slots);
        else
        {
#pragma unroll
            for (int p = 1; p < 4; p++)
                if (p >= r.m)
                    __stcs(r.ga + p, r.slots[p]);
        }
    }
    // after the last element: positions 4 floor(N / 4) .. m + N - 1 (at most six values) did not form a quad for every m
    template <typename real, int N>
    __device__ __forceinline__ void ringTail(const RowRing<real> &r, real *)
    {
        constexpr int P0 = 4 * (N / 4);
#pragma unroll
        for (int p = P0; p < P0 + 6; p++)
            if (p < r.m + N)
                __stcs(r.ga + p, r.slots[p & 7]);
    }
    // state (row of the tile) a thread works on. Vector-store bodies: warp w takes the states w, w + 4, w + 8, ... of
    // its group of 128, so that the position of a thread's output rows inside the 32-byte sector grid,
    // (state N_OUTk) mod 4, is the same for the whole warp: (w N_OUTk) mod 4.
    template <typename Body, int BLOCK>
    __device__ __forceinline__ int threadState(int tid)
    {
        if constexpr (Body::VECTOR_STORES)
        {
            static_assert(BLOCK % 128 == 0, "vector-store bodies need CTAs of four warps (one per alignment class)");
            return (tid & 127) / 32 + (tid & 31) * 4 + (tid & ~127);
        }
        else
            return tid;
    }
    // (Measured alternative, profiles/README.md: one instance of the body per alignment class - template parameter,
    // no predicates, a quarter of the store instructions per warp - is SLOWER: the four instruction streams of a CTA
    // quadruple the instruction-cache footprint of a kernel that is bound by instruction fetch.)
    template <typename Body, typename real, bool FAST>
    __device__ __forceinline__ void runBody(const real *i0, const real *i1, const real *i2, real *o0, real *o1, real *o2,
                                            const OutStage<real> &stage)
    {
        Body::template run<real, FAST>(i0, i1, i2, o0, o1, o2, stage);
    }
    // Flush of one staged chunk: 32 states x COUNT values, COUNT * sizeof(real) contiguous bytes per
    // state. FK / H bodies flush 40-70 chunks. Two things were measured on the way: fully unrolled
    // inline copies with index arithmetic made the kernels 15-19 k instructions long and
    // instruction-fetch bound (no_instructions 39 %); a shared noinline copy fixed that (FK 0.96 ->
    // 0.74 ms) but every call spilled the caller's live registers around it (500 local loads per
    // thread, long_scoreboard 68 %). Hence: inline; full warps load the 16 values of a lane first and
    // store them afterwards (one shared-memory latency per flush), the ragged last tile uses a plain loop.
    template <typename real, int COUNT>
    __device__ __forceinline__ void flushChunkShared(real *__restrict__ g, int row_stride, const real *__restrict__ stg,
                                                  int valid)
    {
        __syncwarp();
        const int lane = threadIdx.x & 31;
        if (COUNT == OUT_CHUNK)
        {
            // full chunk: lane -> (state lane / 16 + 2 k, element lane % 16)
            const int el = lane & (OUT_CHUNK - 1);
            int st = lane / OUT_CHUNK;
            const real *src = stg + st * (OUT_CHUNK + 1) + el;
            real *dst = g + (size_t)st * row_stride + el;
            if (valid == 32)
            {
                // full warp (every tile but the last): all loads first, then all stores, so that the
                // shared-memory latency is paid once per flush and not once per element
                constexpr int STEPS = OUT_CHUNK; // 32 states / (32 / OUT_CHUNK states per step)
                real v[STEPS];
#pragma unroll
                for (int k = 0; k < STEPS; k++)
                    v[k] = src[k * (32 / OUT_CHUNK) * (OUT_CHUNK + 1)];
#pragma unroll
                for (int k = 0; k < STEPS; k++)
                    __stcs(dst + (size_t)k * (32 / OUT_CHUNK) * row_stride, v[k]);
            }
            else
            {
#pragma unroll 1
                for (; st < valid; st += 32 / OUT_CHUNK)
                {
                    __stcs(dst, *src);
                    src += (32 / OUT_CHUNK) * (OUT_CHUNK + 1);
                    dst += (size_t)(32 / OUT_CHUNK) * row_stride;
                }
            }
        }
        else
        {
#pragma unroll 1
            for (int e = lane; e < 32 * COUNT; e += 32)
            {
                const int st = e / COUNT, el = e - st * COUNT;
                if (st < valid)
                    __stcs(g + (size_t)st * row_stride + el, stg[st * (OUT_CHUNK + 1) + el]);
            }
        }
        __syncwarp();
    }
    template <typename real, int N, int COUNT>
    __device__ __forceinline__ void flushChunk(real *__restrict__ g, int base, const real *__restrict__ stg, int valid)
    {
        flushChunkShared<real, COUNT>(g + base, N, stg, valid);
    }
    template <typename Body>
    __host__ __device__ constexpr bool bodyChunked()
    {
        return Body::N_OUT0 > 64 || Body::N_OUT1 > OUT_CHUNK || Body::N_OUT2 > OUT_CHUNK; // = Emitter's rule
    }
    template <typename Body, typename real, int BLOCK>
    __host__ __device__ constexpr size_t chunkStageBytes()
    {
        return bodyChunked<Body>() ? (size_t)Body::STAGE_BUFFERS * BLOCK * (OUT_CHUNK + 1) * sizeof(real) : 0;
    }
    // chunk staging buffers / rings, then the park area of a parked body (Body::PARK_EXTRA slots per thread)
    template <typename Body, typename real, int BLOCK>
    __host__ __device__ constexpr size_t stageBytes()
    {
        return chunkStageBytes<Body, real, BLOCK>() + shapeParkBytes(Body::PARK_EXTRA, BLOCK, (int)sizeof(real));
    }
    template <typename Body, typename real>
    __device__ __forceinline__ OutStage<real> makeOutStage(unsigned char *stage_base, real *out0, real *out1,
                                                           real *out2, int64_t first, int rows, int64_t batch,
                                                           int ts = 0)
    {
        OutStage<real> o;
        // ts = state of the tile this thread works on (threadState); tiles start at multiples of four states
        o.cls[0] = ts < rows ? (ts * Body::N_OUT0) & 3 : -1;
        o.cls[1] = ts < rows ? (ts * Body::N_OUT1) & 3 : -1;
        o.cls[2] = ts < rows ? (ts * Body::N_OUT2) & 3 : -1;
        o.zero = (int)((uint64_t)batch >> 62); // batch < 2^62: always 0, but the compiler cannot know
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        real *buf = reinterpret_cast<real *>(stage_base) + (size_t)warp * 32 * (OUT_CHUNK + 1);
        o.warp = buf;
        o.lane = buf + lane * (OUT_CHUNK + 1);
        o.buf_stride = (int)blockDim.x * (OUT_CHUNK + 1); // layout [array][warp][32][OUT_CHUNK + 1]
        const int64_t s0 = first + 32 * warp;
        o.g[0] = out0 ? out0 + s0 * Body::N_OUT0 : nullptr;
        o.g[1] = out1 ? out1 + s0 * Body::N_OUT1 : nullptr;
        o.g[2] = out2 ? out2 + s0 * Body::N_OUT2 : nullptr;
        const int v = rows - 32 * warp;
        o.valid = v < 0 ? 0 : (v > 32 ? 32 : v);
        if constexpr (Body::PARK_EXTRA > 0)
            o.park = reinterpret_cast<real *>(stage_base + (bodyChunked<Body>() ? (size_t)Body::STAGE_BUFFERS * blockDim.x * (OUT_CHUNK + 1) * sizeof(real) : 0)) +
                     (size_t)threadIdx.x * oddStride(Body::PARK_EXTRA);
        if constexpr (Body::RING_STORES)
        {
            // (tail threads of the last tile replay the last state: same slots' values to the same addresses)
            const int64_t state = first + (ts < rows ? ts : rows - 1);
            real *outs[3] = {out0, out1, out2};
            const int n[3] = {Body::N_OUT0, Body::N_OUT1, Body::N_OUT2};
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                RowRing<real> &r = o.ring[k];
                r.slots = reinterpret_cast<real *>(stage_base) + ((size_t)k * blockDim.x + threadIdx.x) * RingStride<real>::value;
                r.m = (int)((state * n[k]) & 3);
                r.put = r.slots + r.m;
                r.ga = outs[k] ? outs[k] + state * n[k] - r.m : nullptr;
            }
        }
        return o;
    }

    // A parked body keeps values in its tile rows, so the threads behind the end of the batch (last tile) cannot share
    // the row of the last valid state the way the replicas of an unparked body do. They normally sit the body out -
    // unless its large outputs leave through chunk staging, where flushChunk is a warp-cooperative copy that needs
    // every lane: then each of them gets a private copy of the last valid row (and its stores are masked by `valid`).
    template <typename Body>
    __host__ __device__ constexpr bool parkedCooperative()
    {
        return Body::PARKED && bodyChunked<Body>() && !Body::RING_STORES && !Body::VECTOR_STORES;
    }
    template <typename real, int N>
    __device__ __forceinline__ void replicateRow(real *mine, const real *last)
    {
#pragma unroll 1
        for (int i = 0; i < N; i++)
            mine[i] = last[i];
    }

    // Shared-memory budget of one CTA: all input tiles plus the tile of output array 0 when it is
    // small (dynamics: nv values). Large outputs (FK, H) are streamed chunk-wise by the body itself
    // through OUTk(), see below.
    template <typename Body, typename real, int BLOCK>
    struct TileLayout
    {
        static constexpr int S0 = Body::N_IN0 ? oddStride(Body::N_IN0) : 0;
        static constexpr int S1 = Body::N_IN1 ? oddStride(Body::N_IN1) : 0;
        static constexpr int S2 = Body::N_IN2 ? oddStride(Body::N_IN2) : 0;
        static constexpr bool STAGE_OUT0 = Body::N_OUT0 <= 64 && !Body::DIRECT_OUT0;
        static constexpr int SO = STAGE_OUT0 ? oddStride(Body::N_OUT0) : 0;
        static constexpr int OFF1 = S0 * BLOCK;
        static constexpr int OFF2 = OFF1 + S1 * BLOCK;
        static constexpr int OFFO = OFF2 + S2 * BLOCK;
        static constexpr int ELEMS = OFFO + SO * BLOCK;
        static constexpr size_t TILE_BYTES = ((size_t)ELEMS * sizeof(real) + 15) & ~(size_t)15;
        static constexpr size_t BYTES = TILE_BYTES + stageBytes<Body, real, BLOCK>();
    };

    // One state per thread. Body::run(in0, in1, in2, out0, out1, out2, stage) is generated code.
    // STAGED = true : inN address the thread's own row of the staged shared-memory tiles and out0
    //                 the thread's row of the staged output tile (small outputs);
    // STAGED = false: inN / outN address global memory directly (x + state * n).
    template <typename real, typename Body, int BLOCK, int MIN_BLOCKS, bool STAGED, bool FAST>
    __global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
        grbda_batched_kernel(const real *__restrict__ in0, const real *__restrict__ in1,
                             const real *__restrict__ in2, real *__restrict__ out0,
                             real *__restrict__ out1, real *__restrict__ out2, int64_t batch,
                             unsigned char *__restrict__ flags)
    {
        using L = TileLayout<Body, real, BLOCK>;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        gridDependencyWait();
        const int64_t num_tiles = (batch + BLOCK - 1) / BLOCK;
        // FAST: one tile per CTA (grid = number of tiles). !FAST: a few CTAs scan the flags the FAST
        // kernel left and recompute the flagged tiles with the library forms.
        // FAST: tile = blockIdx.x. !FAST: CTA b owns tiles [b * BLOCK, (b + 1) * BLOCK); its threads read
        // one flag each, and the CTA leaves at once when none is set (the normal case).
        int64_t tile = FAST ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * BLOCK;
        int64_t tile_end = tile + 1;
        if (!FAST)
        {
            tile_end = min(num_tiles, tile + BLOCK);
            const int64_t mine = tile + threadIdx.x;
            if (!__syncthreads_or(mine < num_tiles && flags[mine]))
                return;
        }
        do
        {
            if (!FAST && !flags[tile])
            {
                tile++;
                continue;
            }
            const int64_t first = tile * BLOCK;
            const int64_t remaining = batch - first;
            const int rows = remaining < BLOCK ? (int)remaining : BLOCK;
            // Every thread runs the body (it may contain CTA-wide alignment barriers); threads past the end
            // of the batch recompute the last valid state, so their duplicate stores are benign.
            const int ts = threadState<Body, BLOCK>((int)threadIdx.x);
            const bool has_state = ts < rows;
            const int t = min(ts, rows - 1);
            const int64_t state = first + t;

            if constexpr (STAGED)
            {
                real *smem = reinterpret_cast<real *>(smem_raw);
                const OutStage<real> stage = makeOutStage<Body, real>(smem_raw + L::TILE_BYTES, out0, out1, out2, first, rows, batch, ts);
                if (Body::N_IN0)
                    stage_in<real, Body::N_IN0 ? Body::N_IN0 : 1, BLOCK>(in0 + first * Body::N_IN0, smem, rows);
                if (Body::N_IN1)
                    stage_in<real, Body::N_IN1 ? Body::N_IN1 : 1, BLOCK>(in1 + first * Body::N_IN1,
                                                                        smem + L::OFF1, rows);
                if (Body::N_IN2)
                    stage_in<real, Body::N_IN2 ? Body::N_IN2 : 1, BLOCK>(in2 + first * Body::N_IN2,
                                                                        smem + L::OFF2, rows);
                __syncthreads();
                const int tr = parkedCooperative<Body>() ? ts : t; // row this thread works in
                if (parkedCooperative<Body>() && rows < BLOCK)     // (CTA-uniform)
                {
                    if (!has_state)
                    {
                        replicateRow<real, Body::N_IN0>(smem + tr * L::S0, smem + t * L::S0);
                        replicateRow<real, Body::N_IN1>(smem + L::OFF1 + tr * L::S1, smem + L::OFF1 + t * L::S1);
                        replicateRow<real, Body::N_IN2>(smem + L::OFF2 + tr * L::S2, smem + L::OFF2 + t * L::S2);
                    }
                    __syncthreads();
                }
                const real *i0 = smem + tr * L::S0, *i1 = smem + L::OFF1 + tr * L::S1, *i2 = smem + L::OFF2 + tr * L::S2;
                bool ok = true;
                if (FAST && GRBDA_RANGE_CHECKED(Body))
                {
                    ok = __syncthreads_and(Body::template inRange<real>(i0, i1, i2));
                    if (threadIdx.x == 0)
                        flags[tile] = ok ? 0 : 1;
                }
                if (ok)
                {
                    real *o0 = L::STAGE_OUT0 ? smem + L::OFFO + tr * L::SO : out0 + state * Body::N_OUT0;
                    real *o1 = Body::N_OUT1 ? out1 + state * Body::N_OUT1 : nullptr;
                    real *o2 = Body::N_OUT2 ? out2 + state * Body::N_OUT2 : nullptr;
                    if (!Body::PARKED || has_state || parkedCooperative<Body>()) // parked rows are private: no tail replicas
                        runBody<Body, real, FAST>(i0, i1, i2, o0, o1, o2, stage);
                    if (L::STAGE_OUT0)
                    {
                        __syncthreads();
                        stage_out<real, Body::N_OUT0 ? Body::N_OUT0 : 1, BLOCK>(out0 + first * Body::N_OUT0,
                                                                               smem + L::OFFO, rows);
                    }
                }
            }
            else
            {
                const OutStage<real> stage = makeOutStage<Body, real>(smem_raw, out0, out1, out2, first, rows, batch, ts);
                const real *i0 = in0 + state * Body::N_IN0, *i1 = in1 + state * Body::N_IN1, *i2 = in2 + state * Body::N_IN2;
                bool ok = true;
                if (FAST && GRBDA_RANGE_CHECKED(Body))
                {
                    ok = __syncthreads_and(Body::template inRange<real>(i0, i1, i2));
                    if (threadIdx.x == 0)
                        flags[tile] = ok ? 0 : 1;
                }
                if (ok)
                    runBody<Body, real, FAST>(i0, i1, i2, out0 + state * Body::N_OUT0, out1 + state * Body::N_OUT1,
                                              out2 + state * Body::N_OUT2, stage);
            }
            if (!FAST)
                __syncthreads(); // the shared-memory tiles are reused by the next flagged tile
            tile++;
        } while (!FAST && tile < tile_end);
        gridLaunchDependents();
    }

#ifndef __CUDACC_RTC__
    template <typename Body>
    cudaError_t acquireFlags(const LaunchArgs &a, int64_t tiles, unsigned char **flags)
    {
        *flags = nullptr;
        if (!GRBDA_RANGE_CHECKED(Body))
            return cudaSuccess;
        if (!a.flags || a.flags_bytes < (size_t)tiles)
            return cudaErrorInvalidValue;
        *flags = a.flags;
        return cudaSuccess;
    }

    // second pass: tiles the FAST kernel flagged (a joint angle beyond the range of the fast sin/cos
    // reduction) are recomputed with the library forms by the software-staged shell
    template <typename real, typename Body, int BLOCK, int MIN_BLOCKS>
    cudaError_t launchSlowPass(const LaunchArgs &a, int64_t tiles, unsigned char *flags)
    {
        if (!GRBDA_RANGE_CHECKED(Body))
            return cudaSuccess;
        using L = TileLayout<Body, real, BLOCK>;
        // staged tiles unless the rows of this program do not fit into shared memory (then direct I/O)
        constexpr bool STAGED = L::BYTES <= SLOW_PASS_STAGED_LIMIT;
        static_assert(!Body::PARKED || STAGED, "a parked body writes into its tile rows: it cannot run on caller memory");
        constexpr size_t SMEM = STAGED ? L::BYTES : stageBytes<Body, real, BLOCK>();
        auto kernel = grbda_batched_kernel<real, Body, BLOCK, MIN_BLOCKS, STAGED, false>;
        if (SMEM > 48 * 1024)
        {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
            if (e != cudaSuccess)
                return e;
        }
        const int64_t grid = (tiles + BLOCK - 1) / BLOCK;
        const cudaError_t e = launchShell<real>(kernel, (unsigned)grid, BLOCK, SMEM, a.stream, (const real *)a.in[0],
                                                (const real *)a.in[1], (const real *)a.in[2], (real *)a.out[0],
                                                (real *)a.out[1], (real *)a.out[2], a.batch, flags);
        if (a.launched)
            ++*a.launched;
        return e;
    }

    template <typename real, typename Body, int BLOCK, int MIN_BLOCKS, bool STAGED>
    cudaError_t launchBatched(const LaunchArgs &a)
    {
        static_assert(!Body::PARKED || STAGED, "a parked body writes into its tile rows: it cannot run on caller memory");
        using L = TileLayout<Body, real, BLOCK>;
        auto kernel = grbda_batched_kernel<real, Body, BLOCK, MIN_BLOCKS, STAGED, true>;
        const size_t smem = STAGED ? L::BYTES : stageBytes<Body, real, BLOCK>();
        if (smem > 48 * 1024)
        {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
                return e;
        }
        if (a.batch <= 0)
            return cudaSuccess;
        const int64_t grid = (a.batch + BLOCK - 1) / BLOCK;
        unsigned char *flags = nullptr;
        cudaError_t e = acquireFlags<Body>(a, grid, &flags);
        if (e != cudaSuccess)
            return e;
        e = launchShell<real>(kernel, (unsigned)grid, BLOCK, smem, a.stream, (const real *)a.in[0], (const real *)a.in[1],
                              (const real *)a.in[2], (real *)a.out[0], (real *)a.out[1], (real *)a.out[2], a.batch, flags);
        if (a.launched)
            ++*a.launched;
        const cudaError_t e2 = launchSlowPass<real, Body, BLOCK, MIN_BLOCKS>(a, grid, flags);
        return e != cudaSuccess ? e : e2;
    }

#endif // !__CUDACC_RTC__

    // ---------------------------------------------------------------------------------------------
    // TMA-staged shell (variant 'T'): the CTA's input tiles are brought into shared memory by the
    // bulk-copy engine (cp.async.bulk, SASS UBLKCP) instead of by load/store instructions of the
    // compute warps: the straight-line bodies are instruction-issue bound, and a software copy of
    // 81 + 24 values per state costs ~10 % of all issued instructions. One mbarrier tracks the
    // transaction bytes. Results leave the same way (cp.async.bulk shared -> global).
    //   * an array whose rows are a multiple of 16 bytes (FP64: even N, FP32: N % 4 == 0) is copied row
    //     by row, one bulk copy per thread, into rows padded by 16 bytes (keeps the alignment the bulk
    //     engine needs and breaks the power-of-two stride: at most 2-way bank conflicts);
    //   * any other array is copied as one dense tile (row stride N; conflict-free for odd N).
    // Requires 16-byte aligned array base pointers (checked on the host, which otherwise picks the
    // software-staged shell above).
    // ---------------------------------------------------------------------------------------------
    template <typename real>
    // (rows shorter than 12 elements are copied as one dense tile: one bulk copy per thread for 16-64 bytes
    // costs more than the bank conflicts of an even stride do)
    __host__ __device__ constexpr bool tmaRowWise(int n) { return n >= 12 && (n * sizeof(real)) % 16 == 0; }
    template <typename real>
    __host__ __device__ constexpr int tmaStride(int n) { return tmaRowWise<real>(n) ? n + 16 / (int)sizeof(real) : n; }

    __device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
    __device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void mbarArriveExpectTx(uint64_t *bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
                     : "memory");
    }
    __device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity)
    {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "WAIT_%=:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                     "@p bra DONE_%=;\n"
                     "bra WAIT_%=;\n"
                     "DONE_%=:\n"
                     "}" ::"r"(smemAddr(bar)),
                     "r"(parity)
                     : "memory");
    }
    __device__ __forceinline__ void bulkGlobalToShared(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
    {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smemAddr(dst)),
                     "l"(src), "r"(bytes), "r"(smemAddr(bar))
                     : "memory");
    }
    __device__ __forceinline__ void bulkSharedToGlobal(void *dst, const void *src, uint32_t bytes)
    {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smemAddr(src)),
                     "r"(bytes)
                     : "memory");
    }

    // bytes thread `t` contributes to the tile of one input array, and the copies themselves
    template <typename real, int N>
    __device__ __forceinline__ uint32_t tmaTileBytes(int t, int rows)
    {
        if (N == 0)
            return 0;
        if (!tmaRowWise<real>(N)) // dense tile, issued by thread 0; a tail of < 16 bytes is copied by hand
            return t == 0 ? (uint32_t)(((size_t)rows * N * sizeof(real)) & ~(size_t)15) : 0u;
        return t < rows ? (uint32_t)(N * sizeof(real)) : 0u;
    }
    template <typename real, int N>
    __device__ __forceinline__ void tmaTileIssue(const real *__restrict__ g, real *s, int t, int rows, uint64_t *bar)
    {
        if (N == 0)
            return;
        if (!tmaRowWise<real>(N))
        {
            if (t == 0)
            {
                const size_t total = (size_t)rows * N * sizeof(real);
                const size_t bulk = total & ~(size_t)15;
                if (bulk)
                    bulkGlobalToShared(s, g, (uint32_t)bulk, bar);
                for (size_t k = bulk / sizeof(real); k < total / sizeof(real); k++)
                    s[k] = g[k]; // ordered before the waiters by this thread's mbarrier arrive (release)
            }
        }
        else if (t < rows)
            bulkGlobalToShared(s + (size_t)t * tmaStride<real>(N), g + (size_t)t * N, (uint32_t)(N * sizeof(real)), bar);
    }

    template <typename Body, typename real, int BLOCK>
    struct TmaLayout
    {
        static constexpr int S0 = tmaStride<real>(Body::N_IN0), S1 = tmaStride<real>(Body::N_IN1),
                             S2 = tmaStride<real>(Body::N_IN2);
        static constexpr bool STAGE_OUT0 = Body::N_OUT0 <= 64 && !Body::DIRECT_OUT0;
        static constexpr int SO = STAGE_OUT0 ? tmaStride<real>(Body::N_OUT0) : 0;
        // byte offsets; every tile starts on a 16-byte boundary, the first 16 bytes hold the mbarrier
        static constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
        static constexpr size_t OFF0 = 16;
        static constexpr size_t OFF1 = align16(OFF0 + (size_t)S0 * BLOCK * sizeof(real));
        static constexpr size_t OFF2 = align16(OFF1 + (size_t)S1 * BLOCK * sizeof(real));
        static constexpr size_t OFFO = align16(OFF2 + (size_t)S2 * BLOCK * sizeof(real));
        static constexpr size_t TILE_BYTES = align16(OFFO + (size_t)SO * BLOCK * sizeof(real));
        static constexpr size_t BYTES = TILE_BYTES + stageBytes<Body, real, BLOCK>();
    };

    template <typename Body, typename real, int BLOCK>
    __host__ __device__ constexpr bool shapeMirrorsAgree()
    {
        const int n_in[3] = {Body::N_IN0, Body::N_IN1, Body::N_IN2}, n_out[3] = {Body::N_OUT0, Body::N_OUT1, Body::N_OUT2};
        return shapeTileBytes(n_in, n_out, Body::STAGE_BUFFERS, BLOCK, (int)sizeof(real), Body::PARK_EXTRA, Body::DIRECT_OUT0) == TileLayout<Body, real, BLOCK>::BYTES &&
               shapeTmaBytes(n_in, n_out, Body::STAGE_BUFFERS, BLOCK, (int)sizeof(real), Body::PARK_EXTRA, Body::DIRECT_OUT0) == TmaLayout<Body, real, BLOCK>::BYTES;
    }

    template <typename real, typename Body, int BLOCK, int MIN_BLOCKS>
    __global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
        grbda_batched_kernel_tma(const real *__restrict__ in0, const real *__restrict__ in1,
                                 const real *__restrict__ in2, real *__restrict__ out0,
                                 real *__restrict__ out1, real *__restrict__ out2, int64_t batch,
                                 unsigned char *__restrict__ flags)
    {
        using L = TmaLayout<Body, real, BLOCK>;
        static_assert(shapeMirrorsAgree<Body, real, BLOCK>(), "shape* functions out of step with the tile layouts");
        extern __shared__ __align__(128) unsigned char smem_raw[];
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
        real *s0 = reinterpret_cast<real *>(smem_raw + L::OFF0);
        real *s1 = reinterpret_cast<real *>(smem_raw + L::OFF1);
        real *s2 = reinterpret_cast<real *>(smem_raw + L::OFF2);
        real *so = reinterpret_cast<real *>(smem_raw + L::OFFO);

        const int64_t first = (int64_t)blockIdx.x * BLOCK;
        const int64_t remaining = batch - first;
        const int rows = remaining < BLOCK ? (int)remaining : BLOCK;
        const int tid = threadIdx.x;

        if (tid == 0)
            mbarInit(bar, BLOCK);
        __syncthreads();
        gridDependencyWait();
        {
            const real *g0 = in0 + first * Body::N_IN0, *g1 = in1 + first * Body::N_IN1, *g2 = in2 + first * Body::N_IN2;
            const uint32_t bytes = tmaTileBytes<real, Body::N_IN0>(tid, rows) + tmaTileBytes<real, Body::N_IN1>(tid, rows) +
                                   tmaTileBytes<real, Body::N_IN2>(tid, rows);
            tmaTileIssue<real, Body::N_IN0>(g0, s0, tid, rows, bar);
            tmaTileIssue<real, Body::N_IN1>(g1, s1, tid, rows, bar);
            tmaTileIssue<real, Body::N_IN2>(g2, s2, tid, rows, bar);
            mbarArriveExpectTx(bar, bytes);
        }
        mbarWait(bar, 0);

        // every thread runs the body (alignment barriers inside); tail threads redo the last valid state
        const int ts = threadState<Body, BLOCK>(tid);
        const bool has_state = ts < rows;
        const int t = min(ts, rows - 1);
        const int64_t state = first + t;
        const int tr = parkedCooperative<Body>() ? ts : t; // row this thread works in
        if (parkedCooperative<Body>() && rows < BLOCK)     // (CTA-uniform)
        {
            if (!has_state)
            {
                replicateRow<real, Body::N_IN0>(s0 + tr * L::S0, s0 + t * L::S0);
                replicateRow<real, Body::N_IN1>(s1 + tr * L::S1, s1 + t * L::S1);
                replicateRow<real, Body::N_IN2>(s2 + tr * L::S2, s2 + t * L::S2);
            }
            __syncthreads();
        }
        real *o0 = L::STAGE_OUT0 ? so + tr * L::SO : out0 + state * Body::N_OUT0;
        real *o1 = Body::N_OUT1 ? out1 + state * Body::N_OUT1 : nullptr;
        real *o2 = Body::N_OUT2 ? out2 + state * Body::N_OUT2 : nullptr;
        const OutStage<real> stage = makeOutStage<Body, real>(smem_raw + L::TILE_BYTES, out0, out1, out2, first, rows, batch, ts);
        const real *i0 = s0 + tr * L::S0, *i1 = s1 + tr * L::S1, *i2 = s2 + tr * L::S2;
        if (GRBDA_RANGE_CHECKED(Body))
        {
            const bool ok = __syncthreads_and(Body::template inRange<real>(i0, i1, i2));
            if (tid == 0)
                flags[blockIdx.x] = ok ? 0 : 1;
            if (!ok)
                return; // CTA-uniform: the second pass recomputes this tile
        }
        if (!Body::PARKED || has_state || parkedCooperative<Body>()) // parked rows are private: no tail replicas
            runBody<Body, real, true>(i0, i1, i2, o0, o1, o2, stage);
        gridLaunchDependents();

        if (L::STAGE_OUT0)
        {
            // generic-proxy writes to shared memory -> visible to the bulk-copy (async) proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            real *g = out0 + first * Body::N_OUT0;
            if (!tmaRowWise<real>(Body::N_OUT0))
            {
                if (tid == 0)
                {
                    const size_t total = (size_t)rows * Body::N_OUT0 * sizeof(real);
                    const size_t bulk = total & ~(size_t)15;
                    if (bulk)
                        bulkSharedToGlobal(g, so, (uint32_t)bulk);
                    for (size_t k = bulk / sizeof(real); k < total / sizeof(real); k++)
                        g[k] = so[k];
                }
            }
            else if (tid < rows)
                bulkSharedToGlobal(g + (size_t)tid * Body::N_OUT0, so + (size_t)tid * L::SO,
                                   (uint32_t)(Body::N_OUT0 * sizeof(real)));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }

#ifndef __CUDACC_RTC__
    template <typename real, typename Body, int BLOCK, int MIN_BLOCKS>
    cudaError_t launchBatchedTma(const LaunchArgs &a)
    {
        // the bulk-copy engine needs 16-byte aligned global addresses
        uintptr_t bits = 0;
        for (int i = 0; i < 3; i++)
            bits |= (uintptr_t)a.in[i];
        bits |= (uintptr_t)a.out[0];
        if (bits & 15)
            return launchBatched<real, Body, BLOCK, MIN_BLOCKS, true>(a);
        using L = TmaLayout<Body, real, BLOCK>;
        auto kernel = grbda_batched_kernel_tma<real, Body, BLOCK, MIN_BLOCKS>;
        if (L::BYTES > 48 * 1024)
        {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES);
            if (e != cudaSuccess)
                return e;
        }
        if (a.batch <= 0)
            return cudaSuccess;
        const int64_t grid = (a.batch + BLOCK - 1) / BLOCK;
        unsigned char *flags = nullptr;
        cudaError_t e = acquireFlags<Body>(a, grid, &flags);
        if (e != cudaSuccess)
            return e;
        e = launchShell<real>(kernel, (unsigned)grid, BLOCK, L::BYTES, a.stream, (const real *)a.in[0], (const real *)a.in[1],
                              (const real *)a.in[2], (real *)a.out[0], (real *)a.out[1], (real *)a.out[2], a.batch, flags);
        if (a.launched)
            ++*a.launched;
        const cudaError_t e2 = launchSlowPass<real, Body, BLOCK, MIN_BLOCKS>(a, grid, flags);
        return e != cudaSuccess ? e : e2;
    }
#endif // !__CUDACC_RTC__

} // namespace grbda_kernels
#endif // GRBDA_KERNELS_BATCHED_KERNEL_CUH
