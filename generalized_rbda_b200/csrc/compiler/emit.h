// Device-side model compiler, part 2: turn the expression DAG of one algorithm into
//   (a) a straight-line CUDA body (one state per thread) consumed by kernels/batched_kernel.cuh,
//   (b) a flat "tape" (the same program as plain arrays) for introspection and for the CPU-side
//       compiler self-tests in tests/ (a numpy interpreter replays it against the oracle),
//   (c) exact operation counts of the emitted program.
#pragma once
#include <functional>
#include <cstdlib>
#include <iomanip>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <sstream>
#include <algorithm>
#include <string>
#include "sym.h"
#include "../kernels/shapes.h"

namespace grbda
{
    namespace compiler
    {
        struct ProgramStats
        {
            long n_nodes = 0, n_add = 0, n_mul = 0, n_div = 0, n_sqrt = 0, n_sin = 0, n_cos = 0,
                 n_neg = 0, n_select = 0, n_inputs = 0, n_fusable = 0;
            // flops with every add / mul / div / sqrt counted once (an FMA counts 2)
            long flops() const { return n_add + n_mul + n_div + n_sqrt; }
        };

        struct Tape
        {
            std::vector<int32_t> op, a, b, c, e;
            std::vector<double> val;
            std::vector<std::vector<int32_t>> outputs; // per output array: tape index per element
        };

        // Model constants that do not fit an instruction immediate (a double whose low 32 bits are
        // non-zero costs two UMOV instructions when written as a literal) are collected in a
        // __constant__ table of the translation unit; FP64 instructions then read them as constant-bank
        // operands for free. KT(i) in the emitted text indexes this table.
        struct ConstTable
        {
            std::vector<double> values;
            std::unordered_map<uint64_t, int> index;
            static bool fitsImmediate(double v)
            {
                uint64_t bits;
                std::memcpy(&bits, &v, 8);
                return (bits & 0xffffffffull) == 0;
            }
            std::string ref(double v)
            {
                char buf[64];
                if (fitsImmediate(v))
                {
                    std::snprintf(buf, sizeof(buf), "KC(%.17g)", v);
                    return buf;
                }
                uint64_t bits;
                std::memcpy(&bits, &v, 8);
                auto it = index.find(bits);
                int k;
                if (it == index.end())
                {
                    k = (int)values.size();
                    values.push_back(v);
                    index[bits] = k;
                }
                else
                    k = it->second;
                std::snprintf(buf, sizeof(buf), "KT(%d)", k);
                return buf;
            }
            std::string definition(const std::string &name) const
            {
                std::ostringstream os;
                os.precision(17);
                const size_t n = std::max<size_t>(1, values.size());
                os << "static __constant__ double " << name << "_f64[" << n << "] = {";
                for (size_t i = 0; i < n; i++)
                    os << (i ? ", " : "") << (i < values.size() ? values[i] : 0.0);
                os << "};\nstatic __constant__ float " << name << "_f32[" << n << "] = {";
                for (size_t i = 0; i < n; i++)
                {
                    char buf[48];
                    std::snprintf(buf, sizeof(buf), "%.9g", (double)(float)(i < values.size() ? values[i] : 0.0));
                    std::string t = buf;
                    if (t.find_first_of(".en") == std::string::npos)
                        t += ".0";
                    os << (i ? ", " : "") << t << "f";
                }
                os << "};\n";
                os << "template <typename real> __device__ __forceinline__ real " << name << "(int i);\n";
                os << "template <> __device__ __forceinline__ double " << name << "<double>(int i) { return " << name
                   << "_f64[i]; }\n";
                os << "template <> __device__ __forceinline__ float " << name << "<float>(int i) { return " << name
                   << "_f32[i]; }\n";
                return os.str();
            }
        };

        struct Program
        {
            std::string name;
            int n_in[3] = {0, 0, 0};
            std::vector<std::vector<sym::Sym>> outputs; // up to 3 output arrays
        };

        // Parking (staged shells only): a value that is produced long before its next use - the factor
        // entries of the forward-dynamics factorisation wait for the final back-substitution - would be
        // spilled to local memory by ptxas (through L1 into L2). The thread's own shared-memory row holds
        // slots that are dead for most of the program: an input element after it has been read, an
        // element of the staged output row until the result is stored. The emitter parks long-lived
        // values there (PARK_ST / PARK_LD) instead.
        struct ParkConfig
        {
            // usable elements of the in0 / in1 / in2 / out0 rows (0: row not staged) and of the park area behind the tiles
            int n_slots[5] = {0, 0, 0, 0, 0};
            // park only across gaps of at least this fraction of the program (in graph nodes). Measured on
            // B200 (TelloWithArms forward dynamics, 8.9 k nodes): 0.09 -> 0.763 ms, 0.13 -> 0.770, 0.20 -> 0.741,
            // 0.28 -> 0.729, 0.39 -> 0.766, 0.56 -> 0.782, 0.78 -> 0.841, no parking 0.986 ms
            double min_gap_fraction = 0.28;
            int min_gap = 0;               // absolute override (> 0)
        };

        class Emitter
        {
        public:
            Emitter(const sym::Graph &g, const Program &p, ConstTable *consts = nullptr)
                : g_(g), p_(p), consts_(consts)
            {
                analyse();
            }

            const ProgramStats &stats() const { return stats_; }
            int numParked() const { return num_parked_; }       // after cudaBody(..., park)
            int stageBuffers() const { return stage_buffers_; } // after cudaBody: staging buffers per warp
            bool vectorStores() const { return vector_stores_; } // after cudaBody: outputs leave as 256-bit row stores
            bool ringStores() const { return ring_stores_; }     // after cudaBody: ... through per-thread shared-memory rings

            Tape tape() const
            {
                Tape t;
                std::vector<int32_t> remap(g_.nodes.size(), -1);
                for (size_t i = 0; i < g_.nodes.size(); i++)
                    if (live_[i])
                    {
                        const sym::Node &n = g_.nodes[i];
                        remap[i] = (int32_t)t.op.size();
                        t.op.push_back(n.op);
                        const bool leaf = n.op == sym::OP_CONST || n.op == sym::OP_INPUT;
                        t.a.push_back(leaf ? n.a : (n.a >= 0 ? remap[n.a] : -1));
                        t.b.push_back(leaf ? n.b : (n.b >= 0 ? remap[n.b] : -1));
                        t.c.push_back(n.c >= 0 ? remap[n.c] : -1);
                        t.e.push_back(n.e >= 0 ? remap[n.e] : -1);
                        t.val.push_back(n.val);
                    }
                for (auto &arr : p_.outputs)
                {
                    std::vector<int32_t> o;
                    for (auto &s : arr)
                        o.push_back(remap[s.id]);
                    t.outputs.push_back(o);
                }
                return t;
            }

            // Range condition of the fast sin/cos forms (kernels/batched_kernel.cuh): every live sin/cos
            // argument is bounded as |arg| <= a Y + b when all inputs it depends on satisfy |x| <= Y
            // (affine interval propagation through +, -, negation and products / quotients with
            // constants). cudaRangeCheck() returns the generated test "all such inputs are within Y",
            // with Y = min over the arguments of (limit - b) / a; an argument that is not affine in the
            // inputs (never the case for joint angles) disables the fast forms for the program.
            struct TrigRange
            {
                std::vector<std::pair<int, int>> inputs; // (array, element) feeding a sin/cos argument
                bool affine = true;
                double max_gain = 0.0, max_offset = 0.0;  // over the arguments
                double bound(double limit) const          // admissible |input|
                {
                    double Y = 1e300;
                    for (auto &ab : args)
                        if (ab.first > 0)
                            Y = std::min(Y, (limit - ab.second) / ab.first);
                    return Y;
                }
                std::vector<std::pair<double, double>> args; // (a, b) per argument
            };
            TrigRange trigRange() const
            {
                const size_t n = g_.nodes.size();
                const double INF = 1e300;
                std::vector<double> A(n, 0.0), B(n, 0.0);
                std::vector<std::vector<int32_t>> deps(n); // inputs a node depends on (only kept while small)
                TrigRange tr;
                std::vector<char> is_arg(n, 0), feeds(n, 0);
                for (size_t i = 0; i < n; i++)
                    if (live_[i] && (g_.nodes[i].op == sym::OP_SIN || g_.nodes[i].op == sym::OP_COS))
                        is_arg[g_.nodes[i].a] = 1;
                for (size_t i = 0; i < n; i++)
                {
                    const sym::Node &nd = g_.nodes[i];
                    switch (nd.op)
                    {
                    case sym::OP_CONST: A[i] = 0, B[i] = std::fabs(nd.val); break;
                    case sym::OP_INPUT: A[i] = 1, B[i] = 0; break;
                    case sym::OP_ADD:
                    case sym::OP_SUB: A[i] = A[nd.a] + A[nd.b], B[i] = B[nd.a] + B[nd.b]; break;
                    case sym::OP_NEG: A[i] = A[nd.a], B[i] = B[nd.a]; break;
                    case sym::OP_MUL:
                        if (A[nd.a] == 0) // bounded factor (a constant, or e.g. a sine) times an affine one
                            A[i] = B[nd.a] * A[nd.b], B[i] = B[nd.a] * B[nd.b];
                        else if (A[nd.b] == 0)
                            A[i] = B[nd.b] * A[nd.a], B[i] = B[nd.b] * B[nd.a];
                        else
                            A[i] = INF, B[i] = INF;
                        break;
                    case sym::OP_DIV:
                        if (g_.nodes[nd.b].op == sym::OP_CONST && g_.nodes[nd.b].val != 0.0)
                            A[i] = A[nd.a] / std::fabs(g_.nodes[nd.b].val), B[i] = B[nd.a] / std::fabs(g_.nodes[nd.b].val);
                        else
                            A[i] = INF, B[i] = INF;
                        break;
                    case sym::OP_SIN:
                    case sym::OP_COS: A[i] = 0, B[i] = 1; break;
                    default: A[i] = INF, B[i] = INF; break;
                    }
                    if (A[i] > INF)
                        A[i] = INF;
                    if (B[i] > INF)
                        B[i] = INF;
                }
                // inputs feeding the arguments: backward marking
                for (size_t i = n; i-- > 0;)
                {
                    if (is_arg[i])
                        feeds[i] = 1;
                    if (!feeds[i])
                        continue;
                    const sym::Node &nd = g_.nodes[i];
                    if (nd.op == sym::OP_INPUT)
                    {
                        tr.inputs.push_back({nd.a, nd.b});
                        continue;
                    }
                    if (nd.op == sym::OP_CONST)
                        continue;
                    for (int32_t o : {nd.a, nd.b, nd.c, nd.e})
                        if (o >= 0)
                            feeds[o] = 1;
                }
                std::sort(tr.inputs.begin(), tr.inputs.end());
                for (size_t i = 0; i < n; i++)
                    if (is_arg[i])
                    {
                        if (A[i] >= INF || B[i] >= INF)
                            tr.affine = false;
                        tr.args.push_back({A[i], B[i]});
                        tr.max_gain = std::max(tr.max_gain, A[i]);
                        tr.max_offset = std::max(tr.max_offset, B[i]);
                    }
                return tr;
            }
            // expression (over IN0/IN1/IN2 and `real`) that is true when the fast forms may be used
            std::string cudaRangeCheck(double limit64 = 1.0e12, double limit32 = 1.0e6) const
            {
                const TrigRange tr = trigRange();
                if (tr.args.empty())
                    return "true";
                if (!tr.affine || tr.bound(limit32) <= 0)
                    return "false";
                std::ostringstream os;
                os << std::setprecision(17);
                os << "[&]() { unsigned m = 0u;";
                for (auto &in : tr.inputs)
                    os << " m = max(m, absKey(IN" << in.first << "(" << in.second << ")));";
                os << " return m < absKey(sizeof(real) == 8 ? KC(" << tr.bound(limit64) << ") : KC(" << tr.bound(limit32)
                   << ")); }()";
                return os.str();
            }

            // Body text. Inputs are read through IN0(i)/IN1(i)/IN2(i), results written through
            // OUT0(i, x)/OUT1/OUT2; `real` is the arithmetic type; KC(x) makes a literal of type real.
            // sync_every > 0: emit GRBDA_ALIGN() after every `sync_every` statements. All warps of a CTA
            // run the same straight-line code; keeping them within a few hundred instructions of each
            // other lets them share instruction-cache lines (the kernels are instruction-fetch bound
            // otherwise: ncu 'no_instruction' is the top stall of the unaligned version).
            // out_chunk > 0 (kernels with large outputs: FK, H): results are not stored one by one —
            // a thread's row of such an array is thousands of bytes away from its neighbour's, so
            // direct stores touch 32 sectors per instruction. Instead the values of one chunk of
            // `out_chunk` consecutive elements are held until the chunk is complete, written to the
            // warp's shared-memory staging buffer (STG_PUT) and flushed with coalesced stores
            // (STG_FLUSHk: 32 states x chunk, 128 contiguous bytes per state).
            // vector_stores: large output arrays skip the staging altogether: the shell maps the four warps of a CTA to the states s = 0..3 (mod 4) of
            // its tile, so the position of a row inside the 32-byte sector grid, (s N + e) mod 4, is warp-uniform
            // ("class" m = (warp N) mod 4), and every thread writes its own row with 256-bit stores of whole sectors
            // (STGV4; the one or two ragged quads at the row ends go out as 64-bit stores, STGV1). A quad of a class
            // is stored as soon as its last element exists: about one conditional store per element.
            // row_stores = 2 (rings): the same sector-aligned 256-bit stores without alignment classes in the code. A thread
            // writes the elements of a row, in order, into a ring of 8 slots of its own in shared memory at slot
            // (e + m) mod 8, m = (state N) mod 4 being its row's offset inside the sector grid (RING_PUT: the offset is
            // part of the thread's base pointer, so the store address is base + constant); after every fourth element
            // the quad that is complete for EVERY m is read back with two 128-bit loads and stored as one sector
            // (RING_FLUSH), the ragged quads at the two ends of the row element by element (RING_HEAD / RING_TAIL).
            // One instruction stream for all threads, any CTA shape, no values held in registers: 1.75 instructions
            // per element instead of ~2.5 (chunk staging) or 1 predicated store + register shuffling per class.
            std::string cudaBody(int sync_every = 0, int out_chunk = 0, const ParkConfig *park = nullptr,
                                 int row_stores = 0) const
            {
                const bool vector_stores = row_stores == 1;
                std::ostringstream os;
                int since_sync = 0;
                std::vector<char> done(g_.nodes.size(), 0);
                alias_.assign(g_.nodes.size(), std::string());
                num_parked_ = 0;
                // node id -> list of (array, element) it must be stored to
                std::vector<std::vector<std::pair<int, int>>> stores(g_.nodes.size());
                for (size_t arr = 0; arr < p_.outputs.size(); arr++)
                    for (size_t i = 0; i < p_.outputs[arr].size(); i++)
                        stores[p_.outputs[arr][i].id].push_back({(int)arr, (int)i});

                // chunked output staging state
                const size_t n_arr = p_.outputs.size();
                std::vector<std::vector<char>> ready(n_arr);
                std::vector<std::vector<int>> chunk_missing(n_arr);
                std::vector<char> chunked(n_arr, 0);
                for (size_t arr = 0; arr < n_arr; arr++)
                {
                    const int n = (int)p_.outputs[arr].size();
                    // output 0 is staged as a whole row when it is small (<= 64 values); every other output
                    // with more than one chunk's worth of values goes through the chunked path (direct
                    // stores of a thread's own row touch 32 sectors per instruction)
                    chunked[arr] = out_chunk > 0 && grbda_kernels::shapeLargeOutput((int)arr, n);
                    ready[arr].assign(n, 0);
                    if (chunked[arr])
                    {
                        chunk_missing[arr].assign((n + out_chunk - 1) / out_chunk, 0);
                        for (int i = 0; i < n; i++)
                            chunk_missing[arr][i / out_chunk]++;
                    }
                }
                // Arrays whose elements become ready in ascending order (forward kinematics: body after body)
                // get their own staging buffer and are drained sequentially: a value is written to the
                // buffer (STG_PUTK) the moment it exists instead of being held in a register until its
                // 16-value chunk is complete (three arrays x 16 held values were most of FK's spills).
                std::vector<char> immediate(n_arr, 0);
                std::vector<int> drain_next(n_arr, 0);
                {
                    auto basePos = [&](int32_t id) {
                        while (g_.nodes[id].op == sym::OP_NEG)
                            id = g_.nodes[id].a;
                        return g_.nodes[id].op == sym::OP_CONST ? (int32_t)-1 : id;
                    };
                    int n_imm = 0;
                    for (size_t arr = 0; arr < n_arr; arr++)
                    {
                        if (!chunked[arr])
                            continue;
                        bool ascending = true;
                        int32_t last = -1;
                        for (auto &o : p_.outputs[arr])
                        {
                            const int32_t pos = basePos(o.id);
                            if (pos < 0)
                                continue; // constants are put when the drain reaches them
                            if (pos < last)
                                ascending = false;
                            last = pos;
                        }
                        immediate[arr] = ascending;
                        n_imm += ascending;
                    }
                    stage_buffers_ = 1;
                    if (n_imm > 0)
                        stage_buffers_ = (int)n_arr; // buffer k belongs to array k
                    int n_chunked = 0;
                    for (size_t arr = 0; arr < n_arr; arr++)
                        n_chunked += chunked[arr];
                    // only for programs with arrays whose elements appear in ascending order (forward kinematics): a quad then waits
                    // for at most three values. The mass matrix fills its rows out of order (H_ij and H_ji in different
                    // rows, ancestors last): holding its quads costs registers (measured 0.53 against 0.28 ms per 2^18
                    // Tello states), it keeps the chunk staging.
                    vector_stores_ = vector_stores && n_chunked > 0 && n_imm > 0;
                    if (vector_stores_)
                        stage_buffers_ = 0; // no staging buffers at all
                    // rings need the elements of an array (nearly) in order: an element that is ready early waits for its
                    // turn in a register, one that is late holds up everything behind it
                    bool in_order = n_chunked > 0;
                    for (size_t arr = 0; arr < n_arr; arr++)
                    {
                        if (!chunked[arr])
                            continue;
                        std::vector<int32_t> pos_of; // statement that produces each element (-1: a constant)
                        for (auto &o : p_.outputs[arr])
                            pos_of.push_back(basePos(o.id));
                        // a value that exists before one of its predecessors waits in a register until the ring reaches
                        // it: sum of those waits (in statements) / program length = values held on average
                        double held = 0.0;
                        int32_t high = -1;
                        for (int32_t pos : pos_of)
                        {
                            if (pos >= 0 && pos < high)
                                held += (double)(high - pos);
                            high = std::max(high, pos);
                        }
                        held /= (double)std::max<size_t>(1, g_.nodes.size());
                        if (std::getenv("GRBDA_EMIT_TRACE"))
                            std::fprintf(stderr, "array %d: %d elements, %.2f values wait for their turn on average\n", (int)arr,
                                         (int)pos_of.size(), held);
                        // (forward kinematics: 13-19 for the rotation matrices, whose early entries are the rows a joint
                        // rotation leaves unchanged - alive in their parent's matrix anyway; mass matrix: 45-130)
                        if (held > 30.0)
                            in_order = false;
                    }
                    ring_stores_ = row_stores == 2 && in_order && n_imm > 0;
                    if (ring_stores_)
                        stage_buffers_ = (int)n_arr; // ring k lives in staging buffer k's space
                }
                // vector stores: per array and class, the elements each sector-aligned quad still waits for
                auto classesOf = [](int n) { return n % 4 == 0 ? 1 : (n % 2 == 0 ? 2 : 4); };
                auto classOffset = [&](int n, int c) { return classesOf(n) == 4 ? c : (classesOf(n) == 2 ? 2 * c : 0); };
                std::vector<std::vector<std::vector<int>>> quad_missing(n_arr);
                if (vector_stores_)
                    for (size_t arr = 0; arr < n_arr; arr++)
                    {
                        if (!chunked[arr])
                            continue;
                        const int n = (int)p_.outputs[arr].size();
                        quad_missing[arr].assign(classesOf(n), std::vector<int>((n + 3) / 4 + 1, 0));
                        for (int c = 0; c < classesOf(n); c++)
                            for (int e = 0; e < n; e++)
                                quad_missing[arr][c][(e + classOffset(n, c)) / 4]++;
                    }
                auto vectorStore = [&](int arr, int el) {
                    const int n = (int)p_.outputs[arr].size();
                    auto val = [&](int e) { return ref(p_.outputs[arr][e].id); };
                    for (int c = 0; c < classesOf(n); c++)
                    {
                        const int m = classOffset(n, c); // row start inside the sector grid
                        const int quad = (el + m) / 4;
                        if (--quad_missing[arr][c][quad] != 0)
                            continue;
                        const int first = std::max(0, 4 * quad - m), last = std::min(n - 1, 4 * quad - m + 3);
                        if (last - first + 1 == 4)
                            os << "STGV4(" << arr << ", " << m << ", " << first << ", " << val(first) << ", " << val(first + 1)
                               << ", " << val(first + 2) << ", " << val(first + 3) << ");\n";
                        else
                            for (int e = first; e <= last; e++)
                                os << "STGV1(" << arr << ", " << m << ", " << e << ", " << val(e) << ");\n";
                    }
                };
                auto drainRing = [&](int arr) {
                    const int n = (int)p_.outputs[arr].size();
                    while (drain_next[arr] < n && ready[arr][drain_next[arr]])
                    {
                        const int el = drain_next[arr]++;
                        os << "RING_PUT(" << arr << ", " << el << ", " << ref(p_.outputs[arr][el].id) << ");\n";
                        if (el % 4 == 3)
                        {
                            if (el == 3)
                                os << "RING_HEAD(" << arr << ");\n";
                            else
                                os << "RING_FLUSH(" << arr << ", " << (el - 3) / 4 << ");\n";
                        }
                        if (el + 1 == n)
                            os << "RING_TAIL(" << arr << ");\n";
                    }
                };
                auto drain = [&](int arr) {
                    const int n = (int)p_.outputs[arr].size();
                    while (drain_next[arr] < n && ready[arr][drain_next[arr]])
                    {
                        const int el = drain_next[arr]++;
                        os << "STG_PUTK(" << arr << ", " << el % out_chunk << ", " << ref(p_.outputs[arr][el].id) << ");\n";
                        if ((el + 1) % out_chunk == 0 || el + 1 == n)
                        {
                            const int base = (el / out_chunk) * out_chunk;
                            os << "STG_FLUSHI" << arr << "(" << base << ", " << (el + 1 - base) << ");\n";
                        }
                    }
                };
                auto flushChunk = [&](int arr, int c) {
                    const int n = (int)p_.outputs[arr].size();
                    const int base = c * out_chunk, count = std::min(out_chunk, n - base);
                    for (int j = 0; j < count; j++)
                        os << "STG_PUT(" << j << ", " << ref(p_.outputs[arr][base + j].id) << ");\n";
                    os << "STG_FLUSH" << arr << "(" << base << ", " << count << ");\n";
                };
                auto markReady = [&](int arr, int el) {
                    if (ready[arr][el])
                        return;
                    ready[arr][el] = 1;
                    if (ring_stores_)
                    {
                        drainRing(arr);
                        return;
                    }
                    if (vector_stores_)
                    {
                        vectorStore(arr, el);
                        return;
                    }
                    if (immediate[arr])
                    {
                        drain(arr);
                        return;
                    }
                    if (--chunk_missing[arr][el / out_chunk] == 0)
                        flushChunk(arr, el / out_chunk);
                };
                auto emitStores = [&](size_t id) {
                    for (auto &st : stores[id])
                    {
                        if (chunked[st.first])
                            markReady(st.first, st.second);
                        else
                            os << "OUT" << st.first << "(" << st.second << ", " << ref((int32_t)id) << ");\n";
                    }
                };
                // a negated output is available as soon as its operand is
                auto readyNow = [&](int32_t id) {
                    int32_t k = id;
                    while (g_.nodes[k].op == sym::OP_NEG)
                        k = g_.nodes[k].a;
                    return g_.nodes[k].op == sym::OP_CONST || done[k];
                };
                // constant outputs (structural zeros of H, ...) are ready from the start
                for (size_t arr = 0; arr < n_arr; arr++)
                    if (chunked[arr])
                        for (size_t i = 0; i < p_.outputs[arr].size(); i++)
                        {
                            int32_t k = p_.outputs[arr][i].id;
                            while (g_.nodes[k].op == sym::OP_NEG)
                                k = g_.nodes[k].a;
                            if (g_.nodes[k].op == sym::OP_CONST)
                                markReady((int)arr, (int)i);
                        }
                (void)readyNow;
                // NEG outputs become ready together with their operand
                std::vector<std::vector<int32_t>> neg_outputs_of(g_.nodes.size());
                for (size_t arr = 0; arr < n_arr; arr++)
                    for (size_t i = 0; i < p_.outputs[arr].size(); i++)
                    {
                        const int32_t id = p_.outputs[arr][i].id;
                        if (g_.nodes[id].op != sym::OP_NEG)
                            continue;
                        int32_t k = id;
                        while (g_.nodes[k].op == sym::OP_NEG)
                            k = g_.nodes[k].a;
                        if (g_.nodes[k].op != sym::OP_CONST)
                            neg_outputs_of[k].push_back(id);
                    }

                // ---- sums as FMA chains ------------------------------------------------------------
                // ((a b + c d) + x) compiles to mul, fma, add; (a b + (c d + x)) to two fmas. A sum node
                // whose only consumer is another sum (possibly through a negation) and that is not an
                // output is therefore not emitted on its own: it is merged into its consumer, and the
                // merged n-ary sum is emitted with the non-product terms first and every single-use
                // product after them, so that each product contracts into one FMA of the running sum.
                // (Changes the rounding order, not the value; GRBDA_NO_SUM_CHAINS=1 at model-compile
                // time keeps the binary form for comparison.)
                std::vector<char> merged(g_.nodes.size(), 0);
                std::vector<int32_t> host_of(g_.nodes.size(), -1); // statement a node ends up in
                if (!std::getenv("GRBDA_NO_SUM_CHAINS"))
                {
                    const int32_t N = (int32_t)g_.nodes.size();
                    const char *dist = std::getenv("GRBDA_SUM_CHAIN_DIST"); // tuning experiments
                    const int32_t local = dist ? std::atoi(dist) : 40;
                    std::vector<int32_t> single_user(N, -1);
                    for (int32_t i = 0; i < N; i++)
                    {
                        if (!live_[i])
                            continue;
                        const sym::Node &n = g_.nodes[i];
                        if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                            continue;
                        for (int32_t o : {n.a, n.b, n.c, n.e})
                            if (o >= 0 && uses_[o] == 1)
                                single_user[o] = i;
                    }
                    auto isSum = [&](int32_t id) { return g_.nodes[id].op == sym::OP_ADD || g_.nodes[id].op == sym::OP_SUB; };
                    for (int32_t x = 0; x < N; x++)
                    {
                        if (!live_[x] || !isSum(x) || uses_[x] != 1 || !stores[x].empty() || !neg_outputs_of[x].empty())
                            continue;
                        int32_t u = single_user[x];
                        while (u >= 0 && g_.nodes[u].op == sym::OP_NEG)
                        {
                            if (uses_[u] != 1 || !stores[u].empty())
                            {
                                u = -1;
                                break;
                            }
                            u = single_user[u];
                        }
                        // only local merges: a running sum that collects contributions across the whole
                        // program (Schur complements, forces handed up the tree) must stay a sequence of
                        // statements, or every product waits for the last one (measured: frame 264 B -> 4.6 KB)
                        if (u >= 0 && isSum(u) && u - x <= local)
                            merged[x] = 1, host_of[x] = u;
                    }
                    // a chain of merges must stay local as a whole
                    for (int32_t x = 0; x < N; x++)
                        if (merged[x])
                        {
                            int32_t h = x;
                            while (merged[h])
                                h = host_of[h];
                            if (h - x > 3 * local)
                                merged[x] = 0;
                        }
                }
                auto statementOf = [&](int32_t id) {
                    while (merged[id])
                        id = host_of[id];
                    return id;
                };
                // terms of the n-ary sum rooted at a (not merged) sum node
                std::function<void(int32_t, int, std::vector<std::pair<int, int32_t>> &)> sumTerms =
                    [&](int32_t id, int sign, std::vector<std::pair<int, int32_t>> &terms) {
                        const sym::Node &n = g_.nodes[id];
                        const int32_t ops[2] = {n.a, n.b};
                        const int sg[2] = {sign, n.op == sym::OP_ADD ? sign : -sign};
                        for (int k = 0; k < 2; k++)
                        {
                            int32_t o = ops[k];
                            int s2 = sg[k];
                            while (g_.nodes[o].op == sym::OP_NEG)
                                s2 = -s2, o = g_.nodes[o].a;
                            if (merged[o])
                                sumTerms(o, s2, terms);
                            else
                                terms.push_back({s2, o});
                        }
                    };
                auto sumExpression = [&](int32_t id) {
                    std::vector<std::pair<int, int32_t>> terms, ordered;
                    sumTerms(id, 1, terms);
                    for (auto &t : terms) // non-products first: they seed the chain
                        if (!(g_.nodes[t.second].op == sym::OP_MUL && uses_[t.second] == 1))
                            ordered.push_back(t);
                    for (auto &t : terms)
                        if (g_.nodes[t.second].op == sym::OP_MUL && uses_[t.second] == 1)
                            ordered.push_back(t);
                    std::string e;
                    for (size_t k = 0; k < ordered.size(); k++)
                    {
                        const std::string r = ref(ordered[k].second);
                        if (k == 0)
                            e = ordered[k].first > 0 ? r : "(-" + r + ")";
                        else
                            e = "(" + e + (ordered[k].first > 0 ? " + " : " - ") + r + ")";
                    }
                    return e;
                };

                // ---- parking plan (positions are node ids: statements are emitted in id order) ----
                std::vector<std::vector<Parked>> park_store_after(g_.nodes.size()), park_load_before(g_.nodes.size());
                bool any_chunked = false;
                for (size_t arr = 0; arr < n_arr; arr++)
                    any_chunked = any_chunked || chunked[arr];
                // (register quads of the predicated vector stores read their values at per-class positions: not planned)
                if (park && !(any_chunked && vector_stores_))
                {
                    const int32_t N = (int32_t)g_.nodes.size();
                    const int32_t min_gap = park->min_gap > 0 ? park->min_gap
                                                              : std::max(200, (int)(park->min_gap_fraction * stats_.n_nodes));
                    auto base = [&](int32_t id) {
                        while (g_.nodes[id].op == sym::OP_NEG)
                            id = g_.nodes[id].a;
                        return id;
                    };
                    // accesses of every value: its definition and the statements that read it
                    std::vector<std::vector<int32_t>> access(N);
                    for (int32_t i = 0; i < N; i++)
                    {
                        if (!live_[i])
                            continue;
                        const sym::Node &n = g_.nodes[i];
                        if (n.op == sym::OP_CONST || n.op == sym::OP_NEG || n.op == sym::OP_INPUT)
                            continue;
                        if (merged[i])
                            continue; // its operands are read by the statement it is merged into (below)
                        // a cosine emitted together with its sine is defined at the earlier of the two
                        int32_t pos = i;
                        if ((n.op == sym::OP_SIN || n.op == sym::OP_COS) && partner_[i] >= 0 && live_[partner_[i]])
                            pos = std::min<int32_t>(i, partner_[i]);
                        if (n.op == sym::OP_ADD || n.op == sym::OP_SUB)
                        {
                            std::vector<std::pair<int, int32_t>> terms;
                            sumTerms(i, 1, terms);
                            for (auto &t : terms)
                                if (g_.nodes[t.second].op != sym::OP_CONST && g_.nodes[t.second].op != sym::OP_INPUT)
                                    access[t.second].push_back(pos);
                            continue;
                        }
                        for (int32_t o : {n.a, n.b, n.c, n.e})
                            if (o >= 0)
                            {
                                const int32_t x = base(o);
                                if (g_.nodes[x].op != sym::OP_CONST && g_.nodes[x].op != sym::OP_INPUT)
                                    access[x].push_back(pos);
                            }
                    }
                    // Large outputs are not stored where they are defined: an element of a chunk-staged array is read
                    // when the LAST element of its 16-value chunk exists (flushChunk), an element of an in-order array
                    // (rings, immediate staging) when every element before it exists. Those reads are accesses too -
                    // the values that wait for them are what the mass matrix spills.
                    for (size_t arr = 0; arr < n_arr; arr++)
                    {
                        if (!chunked[arr])
                            continue;
                        const int n = (int)p_.outputs[arr].size();
                        std::vector<int32_t> ready_at(n, -1); // statement after which the element exists (-1: constant)
                        for (int e = 0; e < n; e++)
                        {
                            const int32_t b = base(p_.outputs[arr][e].id);
                            if (g_.nodes[b].op != sym::OP_CONST)
                                ready_at[e] = b; // (a negated output leaves with its operand)
                        }
                        const bool in_order_drain = ring_stores_ || immediate[arr];
                        std::vector<int32_t> read_at(n, -1);
                        if (in_order_drain)
                        {
                            int32_t high = -1;
                            for (int e = 0; e < n; e++)
                                read_at[e] = high = std::max(high, ready_at[e]);
                        }
                        else
                            for (int c0 = 0; c0 < n; c0 += out_chunk)
                            {
                                int32_t last = -1;
                                for (int e = c0; e < std::min(n, c0 + out_chunk); e++)
                                    last = std::max(last, ready_at[e]);
                                for (int e = c0; e < std::min(n, c0 + out_chunk); e++)
                                    read_at[e] = last;
                            }
                        for (int e = 0; e < n; e++)
                        {
                            const int32_t b = base(p_.outputs[arr][e].id);
                            if (g_.nodes[b].op == sym::OP_CONST || g_.nodes[b].op == sym::OP_INPUT || read_at[e] <= b)
                                continue;
                            access[b].push_back(read_at[e]);
                        }
                    }
                    // slot availability: [free_from, free_until)
                    std::vector<Slot> slots;
                    for (int t = 0; t < 3; t++)
                    {
                        std::vector<int32_t> read_at(park->n_slots[t], -1);
                        for (int32_t i = 0; i < N; i++)
                            if (live_[i] && g_.nodes[i].op == sym::OP_INPUT && g_.nodes[i].a == t &&
                                g_.nodes[i].b < park->n_slots[t])
                                read_at[g_.nodes[i].b] = i;
                        for (int k = 0; k < park->n_slots[t]; k++)
                            slots.push_back(Slot{t, k, read_at[k] + 1, N, {}});
                    }
                    if (park->n_slots[3] > 0 && n_arr > 0)
                        for (int k = 0; k < park->n_slots[3] && k < (int)p_.outputs[0].size(); k++)
                        {
                            const int32_t b = base(p_.outputs[0][k].id);
                            if (g_.nodes[b].op == sym::OP_CONST || g_.nodes[b].op == sym::OP_INPUT)
                                continue;
                            slots.push_back(Slot{3, k, 0, b, {}});
                        }
                    for (int k = 0; k < park->n_slots[4]; k++)
                        slots.push_back(Slot{4, k, 0, N, {}});
                    // candidates: the longest gap between consecutive accesses of a value
                    std::vector<Cand> cands;
                    for (int32_t x = 0; x < N; x++)
                    {
                        if (!live_[x] || access[x].empty() || merged[x])
                            continue;
                        const sym::Node &n = g_.nodes[x];
                        if (n.op == sym::OP_CONST || n.op == sym::OP_NEG || n.op == sym::OP_INPUT)
                            continue;
                        std::vector<int32_t> a = access[x];
                        int32_t def = x;
                        if ((n.op == sym::OP_SIN || n.op == sym::OP_COS) && partner_[x] >= 0 && live_[partner_[x]])
                            def = std::min<int32_t>(x, partner_[x]);
                        a.push_back(def);
                        // an output store is an access at the definition (emitStores)
                        std::sort(a.begin(), a.end());
                        int32_t best = 0, from = -1, to = -1;
                        for (size_t k = 0; k + 1 < a.size(); k++)
                            if (a[k + 1] - a[k] > best)
                                best = a[k + 1] - a[k], from = a[k], to = a[k + 1];
                        if (best >= min_gap)
                            cands.push_back(Cand{x, from, to});
                    }
                    std::sort(cands.begin(), cands.end(),
                              [](const Cand &p, const Cand &q) { return (p.to - p.from) > (q.to - q.from); });
                    for (const Cand &c : cands)
                    {
                        // the value is stored after statement `from` and reloaded before statement `to`
                        for (Slot &sl : slots)
                        {
                            if (sl.free_from > c.from || sl.free_until <= c.to)
                                continue;
                            bool clash = false;
                            for (auto &b : sl.busy)
                                if (!(c.to <= b.first || b.second <= c.from))
                                    clash = true;
                            if (clash)
                                continue;
                            sl.busy.push_back({c.from, c.to});
                            park_store_after[c.from].push_back(Parked{c.node, sl.tile, sl.index});
                            park_load_before[c.to].push_back(Parked{c.node, sl.tile, sl.index});
                            num_parked_++;
                            break;
                        }
                    }
                }
                auto emitParkLoads = [&](size_t pos) {
                    for (const Parked &pk : park_load_before[pos])
                    {
                        alias_[pk.node] = "t" + std::to_string(pk.node) + "p";
                        os << "const real " << alias_[pk.node] << " = PARK_LD(" << pk.tile << ", " << pk.slot << ");\n";
                    }
                };
                auto emitParkStores = [&](size_t pos) {
                    for (const Parked &pk : park_store_after[pos])
                        os << "PARK_ST(" << pk.tile << ", " << pk.slot << ", " << ref(pk.node) << ");\n";
                };

                // the two most recent arithmetic results: anchor of GRBDA_PIN (kernels: pinAfter) for the
                // sin/cos evaluations, which would otherwise all be hoisted to the top of the program
                int32_t last_temp[2] = {-1, -1};
                auto pinned = [&](int32_t arg) {
                    const int32_t anchor = last_temp[0] != arg ? last_temp[0] : last_temp[1];
                    if (anchor < 0)
                        return ref(arg);
                    return "GRBDA_PIN(" + ref(arg) + ", t" + std::to_string(anchor) + ")";
                };
                for (size_t i = 0; i < g_.nodes.size(); i++)
                {
                    if (!live_[i])
                        continue;
                    const sym::Node &n = g_.nodes[i];
                    if (n.op == sym::OP_CONST || n.op == sym::OP_NEG)
                    {
                        // direct (non chunked) stores of constants / negations are emitted in place;
                        // chunked ones are handled through markReady above / below
                        for (auto &st : stores[i])
                            if (!chunked[st.first])
                            {
                                if (n.op == sym::OP_CONST || readyNow((int32_t)i))
                                    os << "OUT" << st.first << "(" << st.second << ", " << ref((int32_t)i) << ");\n";
                                else
                                    throw std::runtime_error("emit: negation stored before its operand");
                            }
                        continue;
                    }
                    if (merged[i])
                        continue; // part of the sum statement of statementOf(i)
                    if (done[i])
                    {
                        emitStores(i);
                        for (int32_t ng : neg_outputs_of[i])
                            emitStores(ng);
                        emitParkStores(i);
                        continue;
                    }
                    emitParkLoads(i);
                    switch (n.op)
                    {
                    case sym::OP_INPUT:
                        os << "const real t" << i << " = IN" << n.a << "(" << n.b << ");\n";
                        break;
                    case sym::OP_ADD:
                    case sym::OP_SUB:
                        os << "const real t" << i << " = " << sumExpression((int32_t)i) << ";\n";
                        break;
                    case sym::OP_MUL:
                        os << "const real t" << i << " = " << ref(n.a) << " * " << ref(n.b) << ";\n";
                        break;
                    case sym::OP_DIV:
                        os << "const real t" << i << " = GRBDA_DIV(" << ref(n.a) << ", " << ref(n.b) << ");\n";
                        break;
                    case sym::OP_SQRT:
                        os << "const real t" << i << " = sqrt(" << ref(n.a) << ");\n";
                        break;
                    case sym::OP_SELECT_GT:
                        os << "const real t" << i << " = (" << ref(n.a) << " > " << ref(n.b) << ") ? "
                           << ref(n.c) << " : " << ref(n.e) << ";\n";
                        break;
                    case sym::OP_SIN:
                    case sym::OP_COS:
                    {
                        // pair sin/cos of the same argument into one sincos
                        const int32_t other = partner_[i];
                        if (other >= 0 && live_[other])
                        {
                            const int32_t s = n.op == sym::OP_SIN ? (int32_t)i : other;
                            const int32_t c = n.op == sym::OP_SIN ? other : (int32_t)i;
                            os << "real t" << s << ", t" << c << "; grbda_sincos<FAST>(" << pinned(n.a) << ", &t" << s
                               << ", &t" << c << ");\n";
                            done[other] = 1;
                        }
                        else
                            os << "const real t" << i << " = " << (n.op == sym::OP_SIN ? "grbda_sin<FAST>(" : "grbda_cos<FAST>(")
                               << pinned(n.a) << ");\n";
                        break;
                    }
                    default:
                        throw std::runtime_error("emit: unknown op");
                    }
                    done[i] = 1;
                    if (n.op == sym::OP_ADD || n.op == sym::OP_SUB || n.op == sym::OP_MUL || n.op == sym::OP_DIV)
                    {
                        last_temp[1] = last_temp[0];
                        last_temp[0] = (int32_t)i;
                    }
                    emitStores(i);
                    for (int32_t ng : neg_outputs_of[i])
                        emitStores(ng);
                    emitParkStores(i);
                    if (sync_every > 0 && ++since_sync >= sync_every)
                    {
                        os << "GRBDA_ALIGN();\n";
                        since_sync = 0;
                    }
                }
                return os.str();
            }

        private:
            struct Parked
            {
                int32_t node, tile, slot;
            };
            struct Slot
            {
                int32_t tile, index, free_from, free_until;
                std::vector<std::pair<int32_t, int32_t>> busy;
            };
            struct Cand
            {
                int32_t node, from, to;
            };

            std::string ref(int32_t id) const
            {
                const sym::Node &n = g_.nodes[id];
                if (n.op == sym::OP_CONST)
                {
                    if (consts_)
                        return consts_->ref(n.val);
                    char buf[64];
                    std::snprintf(buf, sizeof(buf), "KC(%.17g)", n.val);
                    return buf;
                }
                if (n.op == sym::OP_NEG)
                    return "(-" + ref(n.a) + ")";
                if ((size_t)id < alias_.size() && !alias_[id].empty())
                    return alias_[id]; // reloaded from its parking slot
                return "t" + std::to_string(id);
            }

            void analyse()
            {
                const size_t N = g_.nodes.size();
                live_.assign(N, 0);
                uses_.assign(N, 0);
                partner_.assign(N, -1);
                for (auto &arr : p_.outputs)
                    for (auto &s : arr)
                        live_[s.id] = 1;
                for (long i = (long)N - 1; i >= 0; i--)
                {
                    if (!live_[i])
                        continue;
                    const sym::Node &n = g_.nodes[i];
                    if (n.op == sym::OP_CONST || n.op == sym::OP_INPUT)
                        continue;
                    for (int32_t c : {n.a, n.b, n.c, n.e})
                        if (c >= 0)
                        {
                            live_[c] = 1;
                            uses_[c]++;
                        }
                }
                std::unordered_map<int32_t, int32_t> sin_of, cos_of;
                for (size_t i = 0; i < N; i++)
                {
                    if (!live_[i])
                        continue;
                    const sym::Node &n = g_.nodes[i];
                    stats_.n_nodes++;
                    switch (n.op)
                    {
                    case sym::OP_ADD:
                    case sym::OP_SUB:
                        stats_.n_add++;
                        // a*b+c with a single-use product contracts into one FMA
                        for (int32_t c : {n.a, n.b})
                        {
                            int32_t m = c;
                            if (g_.nodes[m].op == sym::OP_NEG)
                                m = g_.nodes[m].a;
                            if (g_.nodes[m].op == sym::OP_MUL && uses_[m] == 1 && (m == c || uses_[c] == 1))
                            {
                                stats_.n_fusable++;
                                break;
                            }
                        }
                        break;
                    case sym::OP_MUL: stats_.n_mul++; break;
                    case sym::OP_DIV: stats_.n_div++; break;
                    case sym::OP_SQRT: stats_.n_sqrt++; break;
                    case sym::OP_NEG: stats_.n_neg++; break;
                    case sym::OP_SELECT_GT: stats_.n_select++; break;
                    case sym::OP_INPUT: stats_.n_inputs++; break;
                    case sym::OP_SIN:
                        stats_.n_sin++;
                        sin_of[n.a] = (int32_t)i;
                        break;
                    case sym::OP_COS:
                        stats_.n_cos++;
                        cos_of[n.a] = (int32_t)i;
                        break;
                    default: break;
                    }
                }
                for (auto &kv : sin_of)
                {
                    auto it = cos_of.find(kv.first);
                    if (it != cos_of.end())
                    {
                        partner_[kv.second] = it->second;
                        partner_[it->second] = kv.second;
                    }
                }
            }

            const sym::Graph &g_;
            const Program &p_;
            ConstTable *consts_;
            std::vector<char> live_;
            mutable std::vector<std::string> alias_; // name of a value after it was reloaded from its parking slot
            mutable int num_parked_ = 0;
            mutable int stage_buffers_ = 1;
            mutable bool vector_stores_ = false;
            mutable bool ring_stores_ = false;
            std::vector<int> uses_;
            std::vector<int32_t> partner_;
            ProgramStats stats_;
        };

    } // namespace compiler
} // namespace grbda
