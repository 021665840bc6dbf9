"""TEST INFRASTRUCTURE ONLY - builds build/emu/libfdfd_emu.so: the UNMODIFIED sources of libfdfd_b200
(maxwellfdm.jl_b200/csrc/*.cu, *.cpp) compiled with g++ for the CPU against the shim headers of tests/emu/shim, so
that the indexing / synchronisation logic of the CUDA kernels can be checked against the oracle without a GPU
(tests/emu/README.md).  Two purely syntactic rewrites are applied to a scratch copy of each source file:

    kernel<<<grid, block, smem, stream>>>(args);     ->  emu::launch(grid, block, smem, [&]() { kernel(args); });
    extern __shared__ ... name[];                    ->  unsigned char *name = emu::dyn_smem();

Everything else (device keywords, threadIdx, __syncthreads, shuffles, the PTX helpers of ptx_sm100.cuh, the runtime
API) is supplied by the shim.  The library is never loaded by the product path and reports "EMULATED" in
fdfd_version()."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "maxwellfdm.jl_b200", "csrc")
# FDFD_EMU_FMA=1: a second build with fused multiply-adds contracted as nvcc does (-mfma -ffp-contract=fast), to see
# whether a parity tolerance survives the different rounding of the device code (build/emu_fma/)
FMA = bool(os.environ.get("FDFD_EMU_FMA"))
OUT_DIR = os.path.join(ROOT, "build", "emu_fma" if FMA else "emu")
OUT_LIB = os.path.join(OUT_DIR, "libfdfd_emu.so")
FAKE_DIR = os.path.join(OUT_DIR, "fakelibs")
SOURCES = ["api.cu", "apply_naive.cu", "apply_tiled.cu", "apply_rowpair.cu", "krylov.cu", "qmr.cu", "matparams.cu", "coeffs.cpp", "pattern.cpp", "comm.cpp",
           "peer.cpp", "tmap.cpp", "multi.cpp"]


def _split_top(s):
    """split on top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    out, pos = "", 0
    pat = re.compile(r"([A-Za-z_]\w*(?:<[^<>()]*>)?)\s*<<<")
    while True:
        m = pat.search(src, pos)
        if not m:
            return out + src[pos:]
        end_cfg = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():end_cfg])
        assert 2 <= len(cfg) <= 4, cfg
        i = end_cfg + 3
        while src[i].isspace():
            i += 1
        assert src[i] == "(", src[i:i + 40]
        depth, j = 0, i
        while True:
            depth += src[j] == "("
            depth -= src[j] == ")"
            if depth == 0:
                break
            j += 1
        args = src[i + 1:j]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:m.start()]
        out += f"emu::launch(emu::to_dim3({cfg[0]}), emu::to_dim3({cfg[1]}), {smem}, [&]() {{ {m.group(1)}({args}); }})"
        pos = j + 1


def rewrite(src):
    src = rewrite_launches(src)
    src = re.sub(r"extern\s+__shared__[^;]*?(\w+)\s*\[\s*\]\s*;", r"unsigned char *\1 = emu::dyn_smem();", src)
    src = src.replace('"fdfd_b200 0.1.0 (sm_100a)"', '"fdfd_b200 0.1.0 EMULATED on the CPU (test harness, not a product path)"')
    return src


def _newer(target, deps):
    return os.path.exists(target) and all(os.path.getmtime(d) <= os.path.getmtime(target) for d in deps)


def build(verbose=False):
    """Compile what is out of date (per object) and link; returns the path of the library."""
    os.makedirs(OUT_DIR, exist_ok=True)
    shim = os.path.join(HERE, "shim")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
              [os.path.join(shim, f) for f in os.listdir(shim)] + [os.path.join(ROOT, "include", "fdfd_b200.h"), __file__]
    flags = ["-O1", *(["-g"] if os.environ.get("FDFD_EMU_DEBUG") else []), *(["-O2", "-mfma", "-ffp-contract=fast"] if FMA else []), "-std=c++17", "-fPIC", "-fopenmp",
             "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-attributes",
             "-I", shim, "-I", CSRC, "-I", os.path.join(ROOT, "include")]
    objs, procs = [], []
    for f in SOURCES + ["emu_core.cpp"]:
        src = os.path.join(HERE if f == "emu_core.cpp" else CSRC, f)
        obj = os.path.join(OUT_DIR, f + ".emu.o")
        objs.append(obj)
        if _newer(obj, headers + [src]):
            continue
        if f == "emu_core.cpp":
            cpp = src
        else:
            text = rewrite(open(src).read())
            # headers are found through -I (the scratch copy lives elsewhere); the relative include of the C ABI header too
            text = text.replace('#include "../../include/fdfd_b200.h"', '#include "fdfd_b200.h"')
            cpp = os.path.join(OUT_DIR, f + ".emu.cpp")
            with open(cpp, "w") as fh:
                fh.write(f'#line 1 "{src}"\n' + text)
        procs.append((f, subprocess.Popen(["g++", *flags, "-c", cpp, "-o", obj], stderr=subprocess.PIPE, text=True)))
    failed = False
    for name, p in procs:
        _, err = p.communicate()
        if p.returncode != 0 or (verbose and err):
            sys.stderr.write(f"--- {name}\n{err}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("emulation build failed")
    if procs or not _newer(OUT_LIB, objs):
        subprocess.run(["g++", "-shared", "-o", OUT_LIB, *objs, "-ldl", "-lpthread", "-lgomp", "-lrt"], check=True)
    # stand-ins for the libraries csrc/comm.cpp resolves with dlopen, for the multi-rank runs (LD_LIBRARY_PATH=FAKE_DIR)
    os.makedirs(FAKE_DIR, exist_ok=True)
    for src, lib in (("fake_nccl.cpp", "libnccl.so.2"), ("fake_cuda_driver.cpp", "libcuda.so.1")):
        srcp, libp = os.path.join(HERE, "fakelibs", src), os.path.join(FAKE_DIR, lib)
        if not _newer(libp, [srcp]):
            subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", libp, srcp, "-lpthread"], check=True)
    return OUT_LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
