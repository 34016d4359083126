"""The CPU simulator's own checks (tests/cusim/cusim.h): fibers + cooperative barriers give the right answer for a kernel
with barriers, shuffles and named barriers, and the shuffled-schedule mode (CUSIM_SHUFFLE) exposes a missing barrier
that the fixed visiting order hides."""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = r'''
#include <stdint.h>
extern "C" {
// out[b * n + t] = sum of the block's inputs rotated by one lane, through shared memory WITH the barrier
__global__ void good_kernel(const float* in, float* out, int n) {
  __shared__ float sh[256];
  const int t = threadIdx.x, b = blockIdx.x;
  sh[t] = in[b * n + t];
  __syncthreads();
  float v = sh[(t + n - 1) % n] + __shfl_xor_sync(0xffffffffu, sh[t], 1);
  cusim::named_bar(1, 64);  // the first / second 64 threads each use their own instance
  out[b * n + t] = v;
}
// the same neighbour read WITHOUT the barrier: right only if thread t - 1 happens to run before thread t
__global__ void racy_kernel(const float* in, float* out, int n) {
  __shared__ float sh[256];
  const int t = threadIdx.x, b = blockIdx.x;
  sh[t] = in[b * n + t];
  out[b * n + t] = (t == 0) ? in[b * n + n - 1] : sh[t - 1];
}
void run(const float* in, float* out, int blocks, int n, int racy) {
  if (racy) cusim::launch(dim3(blocks), dim3(n), 0, [&] { racy_kernel(in, out, n); });
  else cusim::launch(dim3(blocks), dim3(n), 0, [&] { good_kernel(in, out, n); });
}
}
'''


def _run(tmp_path, shuffle):
    src = tmp_path / "selftest.cpp"
    src.write_text(SRC)
    so = tmp_path / "selftest.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-DB2A_SIM", "-include",
                           os.path.join(HERE, "cusim", "cusim.h"), str(src), "-o", str(so)])
    code = f'''
import ctypes, numpy as np, sys
lib = ctypes.CDLL({str(so)!r})
blocks, n = 24, 128
x = np.arange(blocks * n, dtype=np.float32)
for racy in (0, 1):
    out = np.zeros_like(x)
    lib.run(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), blocks, n, racy)
    xb = x.reshape(blocks, n)
    if racy:
        ref = np.roll(xb, 1, axis=1)
    else:
        ref = np.roll(xb, 1, axis=1) + xb[:, np.arange(n) ^ 1]
    print(int(np.array_equal(out.reshape(blocks, n), ref)))
'''
    env = dict(os.environ)
    env.pop("CUSIM_SHUFFLE", None)
    if shuffle:
        env["CUSIM_SHUFFLE"] = "7"
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
    return [int(v) for v in res.stdout.split()]


def test_fibers_barriers_shuffles_and_shuffled_schedule(tmp_path):
    good, racy = _run(tmp_path, shuffle=False)
    assert good == 1 and racy == 1      # fixed order 0, 1, 2, ...: the missing barrier goes unnoticed
    good, racy = _run(tmp_path, shuffle=True)
    assert good == 1 and racy == 0      # random order: the barrier-correct kernel is unaffected, the racy one is caught
