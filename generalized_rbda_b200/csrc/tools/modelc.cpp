// grbda_modelc — the device-side model compiler as a build-time tool.
//   grbda_modelc --model NAME --urdf-dir DIR --out DIR [--algos id,fd,fk,h,phi,gen]
//                [--variants "BLOCK,MINBLOCKS,STAGED;..."] [--no-f32]
// writes one CUDA translation unit per algorithm (DIR/NAME_<algo>.cu) whose kernels are the
// straight-line per-state programs of that model, plus DIR/NAME_gen.cu (state generation) and
// DIR/NAME.json (sizes, hash, operation counts).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include "../compiler/compile.h"
#include "../host/robots.h"
#include "../host/schedule.h"
#include "../kernels/shapes.h"

using namespace grbda;
using namespace grbda::compiler;

namespace
{
    struct Variant
    {
        int sync = -1; // alignment barrier period in statements (-1: the --sync-every default)
        char kind; // 'S' one state per thread, software-staged I/O; 'T' the same with TMA bulk-copy staging;
                   // 'D' direct global I/O
        int block, min_blocks;
        int program = -1; // alternative program of the entry point (-1: the entry's own), e.g. PROGRAM_FD_LTL
        bool park = false; // staged shells: long-lived values are parked in dead slots of the thread's tile row
        bool direct = false; // output 0 goes straight to global memory although it is small (no output tile)
        bool f32_aba = false; // FP32 launcher of this variant runs the articulated-body sweep (deep fixed-base
                              // chains: H^-1 in FP32 loses cond(H) digits, the O(n) recursion does not)
    };

    std::vector<std::string> split(const std::string &s, char sep)
    {
        std::vector<std::string> out;
        std::stringstream ss(s);
        std::string item;
        while (std::getline(ss, item, sep))
            if (!item.empty())
                out.push_back(item);
        return out;
    }

    std::string ident(const std::string &name)
    {
        std::string s = name;
        for (char &c : s)
            if (!isalnum((unsigned char)c))
                c = '_';
        return s;
    }

    void writeIfChanged(const std::string &path, const std::string &text)
    {
        {
            std::ifstream f(path);
            if (f)
            {
                std::stringstream ss;
                ss << f.rdbuf();
                if (ss.str() == text)
                    return;
            }
        }
        std::ofstream f(path);
        if (!f)
            throw std::runtime_error("cannot write " + path);
        f << text;
    }
} // namespace

int main(int argc, char **argv)
{
    std::string model_name, urdf_dir = ".", out_dir = ".", algos = "id,fd,fk,h,phi,gen";
    std::string variants_s = "S,128,2";
    bool f32 = true;
    int sync_every = 0;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        auto next = [&]() -> std::string
        {
            if (i + 1 >= argc)
            {
                std::fprintf(stderr, "missing value for %s\n", a.c_str());
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--model")
            model_name = next();
        else if (a == "--urdf-dir")
            urdf_dir = next();
        else if (a == "--out")
            out_dir = next();
        else if (a == "--algos")
            algos = next();
        else if (a == "--variants")
            variants_s = next();
        else if (a == "--no-f32")
            f32 = false;
        else if (a == "--sync-every")
            sync_every = std::atoi(next().c_str());
        else
        {
            std::fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    try
    {
        // "--variants spec" or "--variants id=spec|fd=spec|..." (per algorithm)
        std::map<std::string, std::string> per_algo;
        if (variants_s.find('=') != std::string::npos)
        {
            for (auto &entry : split(variants_s, '|'))
            {
                const size_t eq = entry.find('=');
                per_algo[entry.substr(0, eq)] = entry.substr(eq + 1);
            }
            variants_s = per_algo.begin()->second;
        }
        auto parseVariants = [&](const std::string &spec) {
        std::vector<Variant> variants;
        for (auto &v : split(spec, ';'))
        {
            auto p = split(v, ',');
            if (p.size() < 3 || p.size() > 8 || p[0].size() != 1 ||
                std::string("SDT").find(p[0][0]) == std::string::npos)
                throw std::runtime_error("bad --variants entry '" + v +
                                         "' (expected KIND,BLOCK,MINBLOCKS[,SYNC][,ltl][,park][,direct][,f32aba])");
            Variant var;
            var.kind = p[0][0];
            var.block = std::atoi(p[1].c_str());
            var.min_blocks = std::atoi(p[2].c_str());
            for (size_t t = 3; t < p.size(); t++)
            {
                if (p[t] == "ltl")
                    var.program = PROGRAM_FD_LTL;
                else if (p[t] == "auto") // forward dynamics: the program the measured rule picks (compile.h)
                    var.program = -2;
                else if (p[t] == "park")
                    var.park = var.kind == 'S' || var.kind == 'T';
                else if (p[t] == "direct")
                    var.direct = var.kind == 'S' || var.kind == 'T';
                else if (p[t] == "f32aba")
                    var.f32_aba = true;
                else
                    var.sync = std::atoi(p[t].c_str());
            }
            variants.push_back(var);
        }
        if (variants.empty() || variants.size() > 4)
            throw std::runtime_error("between 1 and 4 variants are supported");
        return variants;
        };
        std::vector<Variant> variants = parseVariants(variants_s);

        const ClusterTreeModel model = buildRobotByName(model_name, urdf_dir);
        const uint64_t hash = modelHash(model);
        const std::string id = ident(model_name);
        const int nq = model.getNumPositions(), nv = model.getNumDegreesOfFreedom();
        const int nb = model.getNumBodies(), nc = model.getNumClusters();

        std::ostringstream json;
        json << "{\"model\": \"" << model_name << "\", \"hash\": \"" << std::hex << hash << std::dec
             << "\", \"nq\": " << nq << ", \"nv\": " << nv << ", \"nb\": " << nb << ", \"nc\": " << nc
             << ", \"algos\": {";
        bool first_algo = true;

        const std::string header_common =
            "// GENERATED by grbda_modelc from model '" + model_name +
            "' — do not edit. Straight-line per-state program, one state per thread.\n"
            "#include \"kernels/batched_kernel.cuh\"\n#include \"kernels/stategen.cuh\"\n"
            "#include \"runtime/registry.h\"\n\nnamespace\n{\nusing namespace grbda_kernels;\n\n";

        for (const std::string &algo : split(algos, ','))
        {
            if (algo == "gen")
                continue;
            int a = -1;
            for (int k = 0; k < ALGO_COUNT; k++)
                if (algo == algoName(k))
                    a = k;
            if (a < 0)
                throw std::runtime_error("unknown algorithm '" + algo + "'");
            if (per_algo.count(algo))
                variants = parseVariants(per_algo[algo]);
            ConstTable consts;
            for (auto &v : variants)
                if (v.sync < 0)
                    v.sync = sync_every;
            const int out_chunk = grbda_kernels::OUT_CHUNK; // only arrays with more than 64 values use it
            for (auto &v : variants)
            {
                if (v.program == -2)
                    v.program = a == ALGO_FD ? chooseForwardDynamicsProgram(model) : -1;
                if (v.program >= 0 && (algoOfProgram(v.program) != a || v.program == a))
                    v.program = -1; // "ltl" / "auto" only apply to fd
            }
            {
                // a parked body writes into its tile rows: only where the flagged-tile pass stages them too
                int n_in[3], n_out[3];
                algoSizes(model, a, n_in, n_out);
                for (auto &v : variants)
                    if (v.park && grbda_kernels::shapeTileBytes(n_in, n_out, 3, v.block, 8, 0, v.direct) > grbda_kernels::SLOW_PASS_STAGED_LIMIT)
                        v.park = false;
                // Programs with large outputs (mass matrix): parking pays where the unparked body spills the results
                // that wait for their chunk. Measured on B200 (profiles/README.md, per 2^20 states, unparked -> parked):
                // nv = 24: 1.126 -> 1.014 ms (TelloWithArms), 1.127 -> 1.025 (MIT humanoid); nv = 38: 6.20 -> 5.80 (JVRC1);
                // nv = 18: 0.533 -> 0.531 (Mini Cheetah); nv = 16: 0.442 -> 0.467 (Tello); nv = 8: 0.104 -> 0.132.
                // (a mass matrix of up to 64 values is staged as a row and never held: 0.104 -> 0.136 ms parked, nv = 8)
                if ((a == ALGO_H || grbda_kernels::shapeChunkStageBytes(n_out, 1, 32, 8) > 0) &&
                    std::max(n_out[0], std::max(n_out[1], n_out[2])) < 400 && !std::getenv("GRBDA_PARK_SMALL_OUTPUTS"))
                    for (auto &v : variants)
                        v.park = false;
            }
            // vector-store bodies (large outputs) need CTAs of four warps: only if every variant has them
            bool vec_ok = true;
            for (auto &v : variants)
                vec_ok = vec_ok && v.block % 128 == 0;
            auto programOf = [&](const Variant &v) { return v.program >= 0 ? v.program : a; };
            auto bodyKey = [&](const Variant &v) { return ((v.sync * 2 + (v.park ? 1 : 0)) * 2 + (v.direct ? 1 : 0)) * 16 + programOf(v); };
            CompiledAlgo c = compileAlgo(model, programOf(variants[0]), true, variants[0].sync, &consts, out_chunk,
                                         variants[0].park, vec_ok, 0, variants[0].direct);
            // Park area: a parked body of a program with large outputs (mass matrix: one input row, results held until
            // their chunk is complete) gets the shared memory its tiles leave unused, as extra parking slots per
            // thread - as many as every parked variant of the entry point (and its flagged-tile pass) can hold.
            int park_extra = 0;
            // ... and a body whose output tile was given up for a third CTA per SM (`direct`) gets back as park area
            // what the three CTAs leave (JVRC1 inverse dynamics: 0.444 -> 0.427 ms per 2^19 states without, 0.375 with)
            bool any_direct = false;
            for (auto &v : variants)
                any_direct = any_direct || (v.direct && v.park);
            if (grbda_kernels::shapeChunkStageBytes(c.n_out, 1, 32, 8) > 0 || any_direct || std::getenv("GRBDA_PARK_EXTRA_ALL"))
            {
                long slots = 96;
                bool any = false;
                for (auto &v : variants)
                {
                    if (!v.park)
                        continue;
                    any = true;
                    const size_t tiles = v.kind == 'T' ? grbda_kernels::shapeTmaBytes(c.n_in, c.n_out, c.stage_buffers, v.block, 8, 0, v.direct)
                                                       : grbda_kernels::shapeTileBytes(c.n_in, c.n_out, c.stage_buffers, v.block, 8, 0, v.direct);
                    long per_cta = (long)(grbda_kernels::SM_SHARED_BYTES / v.min_blocks) - 1024 - (long)tiles;
                    if (per_cta < 0) // the variant does not reach its CTA count anyway: what one CTA per SM leaves
                        per_cta = (long)grbda_kernels::SM_SHARED_BYTES - 1024 - (long)tiles;
                    const long slow = (long)grbda_kernels::SLOW_PASS_STAGED_LIMIT -
                                      (long)grbda_kernels::shapeTileBytes(c.n_in, c.n_out, c.stage_buffers, v.block, 8, 0, v.direct);
                    slots = std::min(slots, std::min(per_cta, slow) / (long)(v.block * 8));
                }
                if (const char *e = std::getenv("GRBDA_PARK_EXTRA")) // tuning experiments
                    slots = std::min<long>(slots, std::atol(e));
                if (any && slots >= 3)
                    park_extra = (int)(slots % 2 ? slots : slots - 1); // odd: the stride of the area is the slot count
                if (park_extra > 0 && variants[0].park)
                    c = compileAlgo(model, programOf(variants[0]), true, variants[0].sync, &consts, out_chunk, true, vec_ok,
                                    park_extra, variants[0].direct);
            }
            std::map<int, CompiledAlgo> by_sync; // distinct (alignment period, program) bodies
            // FP32 kernels do not spill (half the register footprint) and are faster without parking
            // (measured: forward dynamics 0.386 against 0.414 ms): their launchers use the unparked body
            auto f32Variant = [&](Variant v) {
                v.park = false;
                if (v.f32_aba)
                    v.program = -1;
                return v;
            };
            for (auto &v : variants)
            {
                if (!by_sync.count(bodyKey(v)))
                    by_sync[bodyKey(v)] = compileAlgo(model, programOf(v), true, v.sync, &consts, out_chunk, v.park, vec_ok,
                                                      v.park ? park_extra : 0, v.direct);
                const Variant u = f32Variant(v);
                if (!by_sync.count(bodyKey(u))) // also the body of the direct-I/O fallback
                    by_sync[bodyKey(u)] = compileAlgo(model, programOf(u), true, u.sync, &consts, out_chunk, false, vec_ok);
            }
            if (a == ALGO_PHI && c.n_out[0] == 0)
                continue; // no implicit clusters
            std::ostringstream os;
            os << header_common;
            os << consts.definition("kc_table");
            if (by_sync.empty())
                by_sync[bodyKey(variants[0])] = c;
            for (auto &kv : by_sync)
                emitBodyStruct(os, "Body" + std::to_string(kv.first), kv.second);
            os << "} // namespace\n\n";
            auto launcher = [&](const Variant &v0, const char *real) {
                Variant v = std::string(real) == "float" ? f32Variant(v0) : v0;
                // staged shells keep every input row (and a small output row) of the CTA in shared memory;
                // a program whose rows do not fit into an SM at all falls back to direct global I/O
                if (v.kind == 'T' || v.kind == 'S')
                {
                    const int elem = std::string(real) == "float" ? 4 : 8;
                    const int extra = v.park ? park_extra : 0;
                    const size_t bytes = v.kind == 'T' ? grbda_kernels::shapeTmaBytes(c.n_in, c.n_out, c.stage_buffers, v.block, elem, extra, v.direct)
                                                       : grbda_kernels::shapeTileBytes(c.n_in, c.n_out, c.stage_buffers, v.block, elem, extra, v.direct);
                    if (bytes + 1024 > grbda_kernels::SM_SHARED_BYTES) // not even one CTA per SM
                    {
                        v.kind = 'D';
                        v.park = false;
                        if (!by_sync.count(bodyKey(v)))
                            throw std::runtime_error("internal: direct-I/O body missing");
                    }
                }
                std::ostringstream l;
                if (v.kind == 'T')
                    l << "&launchBatchedTma<" << real << ", Body" << bodyKey(v) << ", " << v.block << ", " << v.min_blocks << ">";
                else
                    l << "&launchBatched<" << real << ", Body" << bodyKey(v) << ", " << v.block << ", " << v.min_blocks << ", "
                      << (v.kind == 'D' ? "false" : "true") << ">";
                return l.str();
            };
            os << "static const grbda_runtime::AlgoKernels k_algo = {\n    {";
            for (int v = 0; v < 4; v++)
                os << (v < (int)variants.size() ? launcher(variants[v], "double") : std::string("nullptr"))
                   << (v < 3 ? ", " : "");
            os << "},\n    {";
            for (int v = 0; v < 4; v++)
                os << (f32 && v < (int)variants.size() ? launcher(variants[v], "float") : std::string("nullptr"))
                   << (v < 3 ? ", " : "");
            os << "},\n    {" << c.n_in[0] << ", " << c.n_in[1] << ", " << c.n_in[2] << "}, {" << c.n_out[0] << ", "
               << c.n_out[1] << ", " << c.n_out[2] << "},\n    {" << c.stats.n_nodes << ", " << c.stats.n_add
               << ", " << c.stats.n_mul << ", " << c.stats.n_div << ", " << c.stats.n_sqrt << ", "
               << c.stats.n_sin << ", " << c.stats.n_cos << ", " << c.stats.n_fusable << "}};\n";
            os << "static grbda_runtime::AlgoRegistrar r_algo(0x" << std::hex << hash << std::dec << "ull, \""
               << model_name << "\", " << nq << ", " << nv << ", " << nb << ", " << nc << ", " << a
               << ", &k_algo);\n";
            writeIfChanged(out_dir + "/" + id + "_" + algo + ".cu", os.str());

            json << (first_algo ? "" : ", ") << "\"" << algo << "\": {\"nodes\": " << c.stats.n_nodes
                 << ", \"add\": " << c.stats.n_add << ", \"mul\": " << c.stats.n_mul
                 << ", \"div\": " << c.stats.n_div << ", \"sqrt\": " << c.stats.n_sqrt
                 << ", \"sin\": " << c.stats.n_sin << ", \"cos\": " << c.stats.n_cos
                 << ", \"fusable\": " << c.stats.n_fusable << ", \"flops\": " << c.stats.flops() << "}";
            first_algo = false;
        }
        json << "}}\n";

        // ---- state generation -------------------------------------------------------------------
        if (algos.find("gen") != std::string::npos)
        {
            std::ostringstream os;
            os << header_common;
            emitGenerator(os, model);
            os << "} // namespace\n\n";
            os << "static cudaError_t launchGenerate(const grbda_runtime::GenArgs &a)\n{\n"
                  "    if (a.count <= 0) return cudaSuccess;\n"
                  "    const int block = 128;\n"
                  "    grbda_generate_kernel<Gen><<<(unsigned)((a.count + block - 1) / block), block, 0, a.stream>>>(\n"
                  "        a.seed, a.first_index, a.count, a.q, a.yd, a.aux, a.flags);\n"
                  "    return cudaGetLastError();\n}\n";
            os << "static cudaError_t launchIntegrate(const grbda_runtime::StepArgs &a)\n{\n"
                  "    if (a.count <= 0) return cudaSuccess;\n"
                  "    const int block = 128;\n"
                  "    grbda_integrate_kernel<Step><<<(unsigned)((a.count + block - 1) / block), block, 0, a.stream>>>(\n"
                  "        a.q, a.yd, a.ydd, a.dt, a.count, a.q_out, a.yd_out, a.flags);\n"
                  "    return cudaGetLastError();\n}\n";
            os << "static grbda_runtime::GenRegistrar r_gen(0x" << std::hex << hash << std::dec << "ull, \""
               << model_name << "\", " << nq << ", " << nv << ", " << nb << ", " << nc << ", &launchGenerate, &launchIntegrate);\n";
            writeIfChanged(out_dir + "/" + id + "_gen.cu", os.str());
        }
        writeIfChanged(out_dir + "/" + id + ".json", json.str());
        std::printf("%s: hash %016llx nq %d nv %d nb %d nc %d\n", model_name.c_str(), (unsigned long long)hash,
                    nq, nv, nb, nc);
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "grbda_modelc: %s\n", e.what());
        return 1;
    }
    return 0;
}
