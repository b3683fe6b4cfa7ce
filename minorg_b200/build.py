"""Builds libminorg_b200.so (CUDA, sm_100a) in-tree with nvcc.  `python -m minorg_b200.build`.

An object is rebuilt when the CONTENT of its source, of any header or the flags changed (a SHA-256 kept next to it), not when a
file time says so: the prebuilt objects travel with the tree to the GPU box, where checkout times mean nothing."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libminorg_b200.so")
SOURCES = ["genome.cu", "queries.cu", "screen_bitsliced.cu", "screen_general.cu", "screen_gapped.cu", "screen_seedjoin.cu", "enumerate.cu", "maskfind.cu", "fasta_io.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _digest(paths, extra=()):
    h = hashlib.sha256()
    for x in extra:
        h.update(str(x).encode() + b"\0")
    for p in paths:
        h.update(os.path.basename(p).encode() + b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, digest):
    """True unless `target` exists and was built from inputs with this digest (recorded in target + '.hash')."""
    try:
        with open(target + ".hash") as f:
            return not os.path.exists(target) or f.read().strip() != digest
    except OSError:
        return True


def _record(target, digest):
    with open(target + ".hash", "w") as f:
        f.write(digest + "\n")


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "minorg_b200.h"))
    gen = os.path.join(CSRC, "csa_gen.cuh")
    gen_digest = _digest([os.path.join(CSRC, "gen_csa.py")])
    if force or _stale(gen, gen_digest):
        with open(gen, "w") as f:
            subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_csa.py")], stdout=f)
        _record(gen, gen_digest)
    hdrs = sorted(set(hdrs + [gen]))
    objs = []
    procs = []
    digests = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(CSRC, src[:-3] + ".o")
        objs.append(o)
        d = _digest([s] + hdrs, NVCC_FLAGS)
        digests.append(d)
        if force or _stale(o, d):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, o, d, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, o, d, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
        _record(o, d)
    lib_digest = hashlib.sha256("".join(digests).encode()).hexdigest()
    if force or procs or _stale(LIB, lib_digest):
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-lz"])
        _record(LIB, lib_digest)
    return LIB


def source_digest():
    """Digest the shipped library was built from (None if it was not built by this script): tests compare it with the tree."""
    try:
        with open(LIB + ".hash") as f:
            return f.read().strip()
    except OSError:
        return None


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
