"""Build tests/cusim/_build/libb2a_sim.so: the product's .cu sources compiled with g++ against
cusim.h (CUDA threads -> host threads).  TEST INFRASTRUCTURE ONLY -- nothing under
audiotools_b200/ ever loads this library."""
import glob
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "audiotools_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libb2a_sim.so")


def _digest(files):
    h = hashlib.sha256()
    for f in sorted(files):
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(verbose=False):
    """Build (if stale) the simulator library; serialised by a file lock and moved into place atomically, so that the
    two ranks of a gloo test (or pytest-xdist workers) can call it at the same time."""
    import fcntl

    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "cusim.h"),
                                                          os.path.join(REPO, "include", "b2a.h")]
    stamp = os.path.join(OUT_DIR, "stamp")
    dig = _digest(deps)
    if os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s) + ".o")
        cmd = ["g++", "-std=c++20", "-O2", "-g", "-fPIC", "-pthread", "-DB2A_SIM", "-x", "c++",
               "-include", os.path.join(HERE, "cusim.h"), "-Wno-unknown-pragmas", "-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"cusim compile of {s} failed:\n{out.decode()}")
    tmp = OUT + ".tmp.%d" % os.getpid()
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", tmp] + objs)
    os.replace(tmp, OUT)
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
