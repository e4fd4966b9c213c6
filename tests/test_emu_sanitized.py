"""Bounds / undefined-behaviour evidence for the kernel bodies without compute-sanitizer (which is closed on the GPU pool
this repo is built on): tests/emu compiles the product's __host__ __device__ kernel bodies for the host, and here that
translation unit is built with AddressSanitizer + UndefinedBehaviorSanitizer and driven through every frame mode — all
G-buffer / reservoir / ray-queue / candidate-record / BVH / light-table indexing, the per-thread traversal stacks, the
band halo copies and the boundary moves run under the sanitizers. Any out-of-bounds access, use of an uninitialised
stack slot as an index, misaligned load, signed overflow or invalid shift aborts the child process.
What this cannot see: launch geometry, shared memory and inter-stream ordering of the CUDA build (covered by the
bit-exact GPU parity tests over many frames with the pipeline on, and by the multi-band image checks)."""
import os
import subprocess
import sys

import pytest

import emu_binding as eb

HERE = os.path.dirname(os.path.abspath(__file__))
ASAN = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()


@pytest.mark.skipif(not os.path.isabs(ASAN) or not os.path.exists(ASAN), reason="libasan not installed")
def test_kernel_bodies_run_clean_under_asan_and_ubsan():
    lib = eb.build_emu_sanitized()
    env = dict(os.environ, RB_EMU_LIB=lib, LD_PRELOAD=ASAN, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1", OMP_NUM_THREADS="4",
               PYTHONPATH=os.pathsep.join([os.path.dirname(HERE), HERE, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu_sanitized_runner.py")], env=env, capture_output=True,
                       text=True, timeout=1500)
    assert r.returncode == 0 and "SANITIZED-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-6000:])
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-6000:]
