"""In-tree build of libofdmrx.so (CUDA kernels + C-ABI) and the `decode` host driver for sm_100a.

nvcc cross-compiles without a GPU; the built artefacts stay in modem_b200/ (git-ignored, shipped to the GPU box).
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libofdmrx.so")
DECODE = os.path.join(PKG, "decode")
ENCODE = os.path.join(PKG, "encode")
HOSTTEST = os.path.join(OBJ, "libhosttest.so")
FFTHOST = os.path.join(OBJ, "libffthost.so")   # test helper: the product's FFT plans compiled for the host
STIMHOST = os.path.join(OBJ, "libstimhost.so") # test helper: the stimulus routines (stimulus.cuh) compiled for the host
STIMTSAN = os.path.join(OBJ, "stimulus_cta_tsan") # test helper: the CTA-cooperative routines on host threads under ThreadSanitizer

CU = ["polar.cu", "frontend.cu", "acquire.cu", "demod.cu", "ofdmrx.cu", "stimulus.cu"]
CC = ["host_tables.cc", "tx_tables.cc"]
HDRS = ["common.cuh", "polar.cuh", "frontend.cuh", "fft.cuh", "host_tables.h", "stimulus.cuh", "tx_tables.h",
        os.path.join("..", "..", "include", "ofdmrx.h"), os.path.join("..", "..", "include", "ofdmtx.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"] + [("-D%s=%s" % (k, os.environ[k])) for k in ("OFDMRX_SCL_CTAS", "OFDMRX_SCL_PREFETCH", "OFDMRX_SYNC_TMA", "OFDMRX_MT_TILE", "OFDMRX_SCL_STREAM_LEVEL", "OFDMRX_TS_Y_SMEM", "OFDMRX_TS_HUBER_ITS", "OFDMRX_TS_CONV", "OFDMRX_TS_MIN_ITS") if os.environ.get(k)]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libofdmrx cannot be built (there is no CPU fallback)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build(force=False, verbose=False, helpers=True):
    """helpers=False builds the product only (library + CLI drivers): the GPU test session uses that, so a problem with a
    host-side test helper (emulator, host builds of the FFT / stimulus routines, the ThreadSanitizer harness) cannot stop it."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HDRS] + [os.path.abspath(__file__)]
    log, objs, rebuilt = [], [], False
    for src in CU + CC:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            _run([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], log)
            rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"], log)
    main = os.path.join(CSRC, "host", "decode_main.cc")
    if os.path.exists(main) and (rebuilt or _stale(DECODE, [main, LIB])):
        _run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), main, "-o", DECODE,
              "-L", PKG, "-lofdmrx", "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], log)
    enc = os.path.join(CSRC, "host", "encode_main.cc")
    if os.path.exists(enc) and (rebuilt or _stale(ENCODE, [enc, LIB])):
        _run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), enc, "-o", ENCODE,
              "-L", PKG, "-lofdmrx", "-Wl,-rpath,$ORIGIN", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], log)
    if helpers:
        emu = os.path.join(ROOT, "tests", "scl_emulator.cc")
        if os.path.exists(emu) and (force or _stale(HOSTTEST, [emu, os.path.join(CSRC, "host_tables.cc"), os.path.join(CSRC, "host_tables.h")])):
            _run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", emu, os.path.join(CSRC, "host_tables.cc"), "-o", HOSTTEST], log)
        ffth = os.path.join(ROOT, "tests", "fft_host.cu")
        if os.path.exists(ffth) and (force or _stale(FFTHOST, [ffth, os.path.join(CSRC, "fft.cuh"), os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "host_tables.cc")])):
            _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-Xcompiler", "-fPIC", "-shared", ffth,
                  os.path.join(CSRC, "host_tables.cc"), "-o", FFTHOST], log)
        stim = os.path.join(ROOT, "tests", "stimulus_host.cu")
        stim_deps = [stim] + [os.path.join(CSRC, f) for f in ("stimulus.cuh", "fft.cuh", "common.cuh", "host_tables.cc", "tx_tables.cc", "tx_tables.h")]
        if os.path.exists(stim) and (force or _stale(STIMHOST, stim_deps)):
            _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", stim,
                  os.path.join(CSRC, "host_tables.cc"), os.path.join(CSRC, "tx_tables.cc"), "-o", STIMHOST], log)
        tsan = os.path.join(ROOT, "tests", "stimulus_cta_tsan.cu")
        if os.path.exists(tsan) and (force or _stale(STIMTSAN, [tsan] + stim_deps[1:])):
            _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++20", "-O1", "-g", "-Xcompiler", "-fsanitize=thread,-ffp-contract=off,-pthread", tsan,
                  os.path.join(CSRC, "host_tables.cc"), os.path.join(CSRC, "tx_tables.cc"), "-o", STIMTSAN], log)
    with open(os.path.join(OBJ, "build.log"), "a") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
