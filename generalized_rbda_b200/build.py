"""Build libgrbda_cuda.so in-tree.

Steps (all outputs under generalized_rbda_b200/_build and generalized_rbda_b200/csrc/generated):
  1. g++   host model library + the model compiler tool `grbda_modelc`
  2. run   grbda_modelc for every model in MODELS -> csrc/generated/<model>_<algo>.cu
  3. nvcc  -gencode arch=compute_100a,code=sm_100a -lineinfo every generated .cu and the runtime
  4. link  generalized_rbda_b200/libgrbda_cuda.so
Objects are cached by the hash of their (generated) source + flags, so re-builds are incremental.
"""
import concurrent.futures as cf
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
URDF_DIR = os.path.join(HERE, "robot-models")
# kernel experiments: GRBDA_BUILD_TAG=<tag> builds explibs/<tag>/libgrbda_cuda.so (own object / generated dirs) next
# to the product library, GRBDA_BUILD_MODELS=a,b restricts it to some models; tools/ab.py times such builds
# against each other in one GPU session (load with GRBDA_LIB_PATH)
_TAG = os.environ.get("GRBDA_BUILD_TAG")
_OUT = os.path.join(HERE, "..", "explibs", _TAG) if _TAG else HERE
GEN = os.path.join(_OUT, "generated") if _TAG else os.path.join(CSRC, "generated")
BUILD = os.path.join(_OUT, "_build")
LIB = os.path.join(_OUT, "libgrbda_cuda.so")

# model name -> (algorithms, launch variants, build f32 variants). A variant is KIND,BLOCK,MIN_BLOCKS with
# KIND = T (TMA bulk-copy staged tiles), S (software-staged tiles), D (direct global access),
# R (one limb per warp); up to four ';'-separated variants per kernel, the first is the default,
# GRBDA_KERNEL_VARIANT selects another one at run time (tools/sweep_variants.py).
# Kernel variants per entry point: KIND,BLOCK,MIN_BLOCKS[,SYNC][,ltl] separated by ';' (first = default,
# GRBDA_KERNEL_VARIANT=k selects the k-th). KIND: T = TMA-staged tiles, S = software-staged tiles,
# D = direct global I/O. 'ltl' = forward dynamics as CRBA + bias + sparse LTDL (otherwise the
# articulated-body sweep), 'auto' = whichever of the two the measured rule picks for the model
# (compiler/compile.h chooseForwardDynamicsProgram); 'park' = long-lived values parked in dead slots of the thread's
# shared-memory tile row instead of being spilled (T and S only); 'direct' = the (small) output 0 is stored straight to
# global memory instead of through an output tile (frees shared memory for another CTA); 'f32aba' = the FP32 kernel of the variant runs the
# articulated-body sweep. Measured on B200 (profiles/README.md): T,128,2 is the fastest
# and the most device-independent shape for every entry point.
# 'park' on the mass matrix: its results wait in registers for their 16-value chunk (91 spilled doubles on TelloWithArms);
# the body parks them in the park area - the shared memory its single input row leaves unused (kernels/shapes.h)
DEFAULT_VARIANTS = "id=T,128,2;S,128,2|fd=T,128,2,auto,park;T,128,2,ltl,park;T,128,2;S,128,2|fk=T,128,2;S,128,2|h=T,128,2,park;T,128,2;S,128,2|phi=S,128,2|gfa=T,128,2;S,128,2|gfs=T,128,2;S,128,2"
# Inverse dynamics with THREE 128-thread CTAs per SM: without an output tile ('direct') the rows of the mid-size models
# fit three times, and their bodies fit 168 registers without spilling. Measured on B200 per 2^20 states (128 x 2 ->
# 128 x 3): Mini Cheetah 0.205 -> 0.157 ms, 16-link chain 0.206 -> 0.159, Tello 0.238 -> 0.219; no gain or a loss for the
# small models (launch / HBM bound) and for MIT humanoid / TelloWithArms (rows too long, bodies too large:
# profiles/README.md). Forward dynamics needs its registers: 0.313 -> 0.363 ms (Mini Cheetah).
ID3_VARIANTS = DEFAULT_VARIANTS.replace("id=T,128,2;S,128,2", "id=T,128,3,direct;T,128,2;S,128,2")
# Forward dynamics of the SMALL models: bodies that fit 168 registers do not need parking (it costs them) and gain from
# a third CTA per SM. Measured per 2^20 states (parked 128 x 2 -> unparked direct 128 x 3; unparked 128 x 2 in brackets):
# 8-link chain 0.176 -> 0.088 ms (0.100), planar leg linkage 0.033 -> 0.026 (0.032), six-bar 0.045 -> 0.043, MIT
# humanoid leg 0.069 -> 0.066, 4-link pair chain 0.032 -> 0.030; no change for the 2-4 link chains and the four-bar;
# slower for everything larger (16-link chain 0.271 -> 0.380, Tello 0.395 -> 0.602, triple chain 0.358 -> 0.639).
FD3_VARIANTS = DEFAULT_VARIANTS.replace("fd=T,128,2,auto,park;T,128,2,ltl,park;T,128,2;S,128,2",
                                        "fd=T,128,3,auto,direct;T,128,2,auto,park;T,128,2;S,128,2")
SYNC_EVERY = int(os.environ.get("GRBDA_SYNC_EVERY", "0"))  # alignment barriers measured useless (profiles/)
MODELS = {
    "tello_with_arms": ("id,fd,fk,h,phi,gfa,gfs,gen", "id=T,128,2;S,128,2|fd=T,128,2,ltl,park;T,128,2,ltl;S,128,2,ltl,park;S,128,2|fk=T,128,2;S,128,2|h=T,128,2,park;T,128,2;S,128,2|phi=S,128,2|gfa=T,128,2;S,128,2|gfs=T,128,2;S,128,2", True),
    "tello": ("id,fd,fk,h,phi,gfa,gfs,gen", ID3_VARIANTS, False),
    "mini_cheetah": ("id,fd,fk,h,gfa,gfs,gen", ID3_VARIANTS, True),
    "mit_humanoid": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, True),
    "four_bar": ("id,fd,fk,h,phi,gfa,gfs,gen", DEFAULT_VARIANTS, True),
    "six_bar": ("id,fd,fk,h,phi,gfa,gfs,gen", FD3_VARIANTS, False),
    "planar_leg_linkage": ("id,fd,fk,h,phi,gfa,gfs,gen", FD3_VARIANTS, False),
    "mit_humanoid_leg": ("id,fd,fk,h,phi,gfa,gfs,gen", FD3_VARIANTS, False),
    # 58 bodies: the tile rows of a 128-thread CTA take 163 KB (one CTA per SM); 64-thread CTAs fit twice.
    # Both dynamics kernels spill and are parked (ID 1.01 -> 0.86 ms, FD 3.3-4.2 ms per 2^20 states, L2 dependent).
    # Inverse dynamics: without an output tile ('direct': 38 plain stores per thread) THREE 64-thread CTAs fit, and what
    # they leave becomes park area: 0.444 -> 0.375 ms per 2^19 states. Forward dynamics loses more parking slots than
    # the third CTA gives back (1.92 -> 2.17 ms): unchanged.
    "jvrc1_humanoid": ("id,fd,fk,h,phi,gfa,gfs,gen",
                       "id=T,64,3,park,direct;T,64,2,park;T,128,2,park;S,128,2|fd=T,64,2,ltl,park;T,128,2,ltl,park;T,64,2,park;S,128,2|fk=T,128,2;S,128,2|h=T,128,2,park;T,128,2;S,128,2|phi=S,128,2|gfa=T,128,2;S,128,2|gfs=T,128,2;S,128,2", True),
    "revolute_rotor_chain": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, True),
    "revolute_chain_with_rotor_2": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, True),
    "revolute_chain_with_rotor_4": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, False),
    # depth sweep of the serial cluster chain (BASELINE config 5: deep-tree latency)
    "revolute_chain_with_rotor_8": ("id,fd,fk,h,gfa,gfs,gen", FD3_VARIANTS, False),
    # FP32: forward dynamics of the 16-link fixed-base chain (cond(H) ~ 2e4) through the articulated-body sweep:
    # median error 6e-7 instead of 1e-4 with the factorisation (measured, tests/test_gpu_parity.py)
    "revolute_chain_with_rotor_16": ("id,fd,fk,h,gfa,gfs,gen",
                                     ID3_VARIANTS.replace("fd=T,128,2,auto,park;", "fd=T,128,2,auto,park,f32aba;"), True),
    "revolute_pair_chain_with_rotor_2": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, False),
    "revolute_pair_chain_with_rotor_4": ("id,fd,fk,h,gfa,gfs,gen", FD3_VARIANTS, False),
    # the remaining cluster-joint classes of the reference: RevolutePair (RevolutePairChain.cpp) and
    # RevoluteTripleWithRotor (RevoluteTripleChainWithRotor.cpp, seeded parameters)
    "revolute_pair_chain_4": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, False),
    "revolute_triple_chain_with_rotor_6": ("id,fd,fk,h,gfa,gfs,gen", DEFAULT_VARIANTS, False),
}

CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# (-D switches of GRBDA_EXTRA_NVCCFLAGS reach the host compiler too: kernels/shapes.h is shared by both sides)
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-variable", "-Wno-sign-compare"] + \
    [f for f in os.environ.get("GRBDA_EXTRA_NVCCFLAGS", "").split() if f.startswith("-D")]
NVCCFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
             "-Xcompiler", "-fPIC", "-I", CSRC, "-ccbin", CXX,
             "-DGRBDA_DEFAULT_URDF_DIR=\"%s\"" % URDF_DIR] + os.environ.get("GRBDA_EXTRA_NVCCFLAGS", "").split()

HOST_SOURCES = ["host/robots.cpp", "host/robot_factory.cpp", "host/urdf.cpp"]


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:20]


def _headers():
    out = []
    for root, _, files in os.walk(CSRC):
        if root.startswith(GEN):
            continue
        for f in files:
            if f.endswith((".h", ".cuh")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(HERE, "..", "include", "grbda_cuda.h"))
    return sorted(out)


def source_digest():
    """Digest of everything the library is built from (csrc/ minus generated code, the C header, this file's model
    table): written next to the library as libgrbda_cuda.stamp; tests/conftest.py refuses a library whose stamp does
    not match the sources in the tree."""
    paths = list(_headers())
    for root, _, files in os.walk(CSRC):
        if root.startswith(os.path.join(CSRC, "generated")):
            continue
        paths += [os.path.join(root, f) for f in files if f.endswith((".cpp", ".cu"))]
    paths.append(os.path.abspath(__file__))
    return _digest(sorted(set(os.path.abspath(p) for p in paths)))


def _compile_cached(src, obj_dir, compiler_cmd, dep_digest, log):
    """Compile src -> object named by content hash; returns object path."""
    key = _digest([src], dep_digest + " ".join(compiler_cmd))
    obj = os.path.join(obj_dir, os.path.basename(src) + "." + key + ".o")
    if not os.path.exists(obj):
        log("  compile %s" % os.path.relpath(src, HERE))
        _run(compiler_cmd + ["-c", src, "-o", obj])
    return obj


def build(verbose=True, jobs=None, models=None):
    log = (lambda s: print(s, flush=True)) if verbose else (lambda s: None)
    os.makedirs(GEN, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    models = models or MODELS
    if os.environ.get("GRBDA_BUILD_MODELS"):
        models = {k: v for k, v in MODELS.items() if k in os.environ["GRBDA_BUILD_MODELS"].split(",")}
    jobs = jobs or max(1, (os.cpu_count() or 2))
    hdr_digest = _digest(_headers())

    # 1. host objects + modelc
    host_objs = [_compile_cached(os.path.join(CSRC, s), BUILD, [CXX] + CXXFLAGS, hdr_digest, log)
                 for s in HOST_SOURCES]
    modelc_obj = _compile_cached(os.path.join(CSRC, "tools/modelc.cpp"), BUILD, [CXX] + CXXFLAGS, hdr_digest, log)
    modelc = os.path.join(BUILD, "grbda_modelc")
    _run([CXX, "-o", modelc, modelc_obj] + host_objs)

    # 2. generate
    gen_sources = []
    # experiments: GRBDA_VARIANTS="fk=T,128,2;T,256,1|h=..." replaces the variant lists of the named entry points
    override = dict(seg.split("=", 1) for seg in os.environ.get("GRBDA_VARIANTS", "").split("|") if "=" in seg)
    for name, (algos, variants, f32) in models.items():
        if override:
            segs = dict(seg.split("=", 1) for seg in variants.split("|"))
            segs.update(override)
            variants = "|".join("%s=%s" % kv for kv in segs.items())
        cmd = [modelc, "--model", name, "--urdf-dir", URDF_DIR, "--out", GEN, "--algos", algos,
               "--variants", variants, "--sync-every", str(SYNC_EVERY)]
        if not f32:
            cmd.append("--no-f32")
        log("  " + _run(cmd).strip())
        ident = "".join(c if c.isalnum() else "_" for c in name)
        for a in algos.split(","):
            p = os.path.join(GEN, "%s_%s.cu" % (ident, a))
            if os.path.exists(p):
                gen_sources.append(p)

    # 2b. the kernel headers as string literals for the run-time compiler (runtime/jit.cpp hands them to NVRTC)
    inc = []
    for var, rel in (("shapes", "kernels/shapes.h"), ("batched", "kernels/batched_kernel.cuh"),
                     ("stategen", "kernels/stategen.cuh")):
        with open(os.path.join(CSRC, rel)) as f:
            text = f.read()
        assert ')GRBDAHDR"' not in text
        # string literals are limited to 64 KB by some compilers: concatenate pieces
        pieces = [text[i:i + 12000] for i in range(0, len(text), 12000)]
        inc.append("static const char k_jit_header_%s[] =\n%s;\n" % (
            var, "\n".join('R"GRBDAHDR(%s)GRBDAHDR"' % piece for piece in pieces)))
    inc_path = os.path.join(BUILD, "jit_headers.inc")
    inc_text = "// generated by build.py from csrc/kernels — do not edit\n" + "\n".join(inc)
    if not os.path.exists(inc_path) or open(inc_path).read() != inc_text:
        with open(inc_path, "w") as f:
            f.write(inc_text)

    # 3. nvcc (parallel)
    cuda_sources = gen_sources + [os.path.join(CSRC, "runtime/capi.cu")]
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = [ex.submit(_compile_cached, s, BUILD, [NVCC] + NVCCFLAGS, hdr_digest, log) for s in cuda_sources]
        cuda_objs = [f.result() for f in futs]
    reg_obj = _compile_cached(os.path.join(CSRC, "runtime/registry.cpp"), BUILD,
                              [NVCC] + NVCCFLAGS + ["-x", "cu"], hdr_digest, log)
    jit_obj = _compile_cached(os.path.join(CSRC, "runtime/jit.cpp"), BUILD,
                              [NVCC] + NVCCFLAGS + ["-x", "cu", "-I", BUILD], hdr_digest, log)
    batched_obj = _compile_cached(os.path.join(CSRC, "host/batched.cpp"), BUILD,
                                  [NVCC] + NVCCFLAGS + ["-x", "cu"], hdr_digest, log)

    # 4. link
    _run([NVCC, "-shared", "-o", LIB, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", CXX]
         + cuda_objs + [reg_obj, jit_obj, batched_obj] + host_objs + ["-lcudart", "-ldl"])
    # drop stale cached objects so that _build does not grow without bound
    keep = set(cuda_objs + [reg_obj, jit_obj, batched_obj, modelc_obj, modelc] + host_objs)
    for f in os.listdir(BUILD):
        p = os.path.join(BUILD, f)
        if p not in keep and f.endswith(".o"):
            os.remove(p)
    with open(os.path.splitext(LIB)[0] + ".stamp", "w") as f:
        f.write(source_digest() + "\n")
    log("  linked %s" % os.path.relpath(LIB, os.path.join(HERE, "..")))
    return LIB


if __name__ == "__main__":
    build()
