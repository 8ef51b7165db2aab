"""CPU suite: the REAL host orchestration of the C ABI (zk_b200/csrc/api.cu) executed without a GPU.

api.cu is compiled as plain C++ and linked with tests/cpp/hostmock/ — a stand-in for the CUDA runtime calls it makes
("device" memory is host memory, launches run inline) and for the kernel launchers: the thread-replayable kernels
(sop_kernel.cuh, ntt_sharded_kernels.cuh) run from their real source, the rest are naive models of their documented
contracts.  tests/hostmock_driver.py then drives the library through the Python mirror in a child process (ZK_B200_LIB
points at the mock) and compares with the oracle: prove / prove_partial / verify round loops with the derived-S(1)
claims, the absorb pipeline, evaluate / partial_evaluate chains, the sum-of-products prover, zk_ntt's buffer swap with
its plan, and the multi-GPU NTT's step functions with 2/4/8 virtual ranks (the path that has not yet run on hardware).
A second driver (tests/hostmock_sharded_driver.py) runs the SHARDED entry points with every rank a thread and an
in-process stand-in for NCCL.

TEST INFRASTRUCTURE ONLY: it checks the caller side of every launch.  The product library is untouched by it and keeps
refusing to run without a GPU (test_abi_host.py::test_no_cpu_fallback_without_gpu)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


_HOSTMOCK_SO = None


def _build_hostmock():
    """Builds build/libzk_b200_hostmock.so (+ build/mock/libnccl.so.2) once per test session."""
    global _HOSTMOCK_SO
    if _HOSTMOCK_SO is None:
        _HOSTMOCK_SO = _build_hostmock_once()
    return _HOSTMOCK_SO


def _build_hostmock_once():
    out_dir = os.path.join(ROOT, "build", "mock")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(ROOT, "zk_b200", "csrc")
    mock = os.path.join(ROOT, "tests", "cpp", "hostmock")
    common = ["g++", "-std=c++17", "-O2", "-w", "-fPIC", "-I", "/usr/local/cuda/include", "-I", src]
    jobs = [
        (common + ["-x", "c++", "-fvisibility=hidden", "-c", os.path.join(src, "api.cu"), "-o", os.path.join(out_dir, "api.o")]),
        (common + ["-x", "c++", "-fvisibility=hidden", "-c", os.path.join(src, "api_sumcheck.cu"), "-o", os.path.join(out_dir, "api_sumcheck.o")]),
        (common + ["-x", "c++", "-fvisibility=hidden", "-c", os.path.join(src, "api_ntt.cu"), "-o", os.path.join(out_dir, "api_ntt.o")]),
        (common + ["-fvisibility=hidden", "-c", os.path.join(mock, "mock_kernels.cpp"), "-o", os.path.join(out_dir, "mock_kernels.o")]),
        (common + ["-c", os.path.join(mock, "mock_cudart.cpp"), "-o", os.path.join(out_dir, "mock_cudart.o")]),
        (["g++", "-std=c++17", "-O3", "-mavx512f", "-mavx512vl", "-fPIC", "-fvisibility=hidden", "-c", os.path.join(src, "keccak_avx512.cpp"),
          "-o", os.path.join(out_dir, "keccak_avx512.o")]),
    ]
    procs = [subprocess.Popen(j, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for j in jobs]
    for p, j in zip(procs, jobs):
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, " ".join(j) + "\n" + out[-3000:]
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-shared", os.path.join(mock, "mock_nccl.cpp"),
                        "-o", os.path.join(out_dir, "libnccl.so.2"), "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    so = os.path.join(ROOT, "build", "libzk_b200_hostmock.so")
    r = subprocess.run(["g++", "-shared", "-o", so] + [os.path.join(out_dir, f) for f in ("api.o", "api_sumcheck.o", "api_ntt.o", "mock_kernels.o", "mock_cudart.o", "keccak_avx512.o")]
                       + ["-ldl", "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return so


def test_api_orchestration_against_the_oracle_under_the_host_mock():
    so = _build_hostmock()
    env = dict(os.environ, ZK_B200_LIB=so)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostmock_driver.py")], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["hostmock_orchestration_ok"] and res["checks"] >= 300 and not res["failures"], res


def test_sharded_entry_points_with_ranks_as_threads_under_the_host_mock():
    """world = 2, 4, 8: every rank a thread with its own sharded context, an in-process NCCL stand-in
    (tests/cpp/hostmock/mock_nccl.cpp: collectives = thread rendezvous + copies): the sharded ProductPoly and
    sum-of-products provers (exact all-reduce + narrowing, derived S(1) on rank 0, residual gather at several
    thresholds) and zk_ntt_sharded's send/recv exchanges — the real api.cu code, every rank checked against the oracle."""
    so = _build_hostmock()
    mockdir = os.path.join(ROOT, "build", "mock")
    env = dict(os.environ, ZK_B200_LIB=so, LD_LIBRARY_PATH=mockdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostmock_sharded_driver.py")], capture_output=True, text=True, env=env,
                       timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["hostmock_sharded_ok"] and res["checks"] >= 1000 and not res["failures"], res


def test_gpu_suite_files_replayed_under_the_host_mock():
    """The GPU suite's own test files (the reference's tests re-expressed, the parity cases, the sum-of-products and NTT
    cases) run in a child pytest with the host mock loaded instead of the product library: every argument check, status
    code, error string, buffer hand-over and round loop those tests drive is executed here on the CPU through the real
    api.cu.  (Only the 2^22+ cases are deselected: the mock's arithmetic is word-serial host code.)  This does NOT
    replace the GPU run — the kernels are stand-ins here; it makes sure a red GPU suite can only mean a kernel problem."""
    so = _build_hostmock()
    mockdir = os.path.join(ROOT, "build", "mock")
    env = dict(os.environ, ZK_B200_LIB=so, LD_LIBRARY_PATH=mockdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    files = [os.path.join(ROOT, "tests", f) for f in ("test_gpu_reference_kats.py", "test_gpu_sop.py", "test_gpu_parity.py", "test_gpu_ntt.py")]
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-k",
           "not large_prove_self_consistency and not large_roundtrip and not multi_chunk and not fullsize_digest"] + files
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:] + r.stderr[-2000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 80, tail
    assert "xpassed" not in tail and "xfailed" not in tail, tail
    # the sum-of-products file once more with the deferred-reduction variant of the kernel source (ZK_B200_SOP_WIDE=1)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", files[1]], capture_output=True, text=True,
                       env=dict(env, ZK_B200_SOP_WIDE="1"), timeout=900, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail and int(tail.split(" passed")[0].split()[-1]) >= 15, r.stdout[-3000:] + r.stderr[-2000:]


def test_cpp_mirror_runs_under_the_host_mock():
    """include/zk_b200.hpp (the compiled-language mirror of the reference API) executed on the CPU: the reference's tests
    re-expressed in C++ (tests/cpp/test_reference_kats.cpp, the GPU suite runs the same binary against the product) and
    the GKR layer through SumOfProductsPoly / SumcheckProver / SumcheckVerifier (tests/cpp/test_sop_mirror_compiles.cpp run)."""
    so = _build_hostmock()
    libdir = os.path.dirname(so)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "build", "mock") + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    for src, args, want in (("test_reference_kats.cpp", [], "ALL C++ MIRROR TESTS PASSED"), ("test_sop_mirror_compiles.cpp", ["run"], "GKR LAYER OK")):
        exe = os.path.join(ROOT, "build", "mock", src.replace(".cpp", "_mock"))
        cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", src), "-L", libdir,
               "-lzk_b200_hostmock", f"-Wl,-rpath,{libdir}", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0 and want in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_abi_walk_under_address_leak_and_ub_sanitizers():
    """api.cu + the host mock + the replayed kernel sources built with -fsanitize=address,undefined, then a walk over the C
    ABI (tests/cpp/test_abi_sanitized.cpp) and the two C++ mirror programs with leak detection on: "device" memory is
    host memory under the mock, so an out-of-bounds access of the orchestration or of a replayed kernel, a double free,
    or a leaked handle / event / buffer fails here.  (This run found the two CUDA events absorb_tables used to leak per
    prove()/verify() call.)"""
    import pytest

    probe = subprocess.run(["g++", "-fsanitize=address,undefined", "-x", "c++", "-", "-o", os.devnull], input="int main(){return 0;}",
                           capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("no sanitizer runtime in this toolchain")
    out_dir = os.path.join(ROOT, "build", "asan")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(ROOT, "zk_b200", "csrc")
    mock = os.path.join(ROOT, "tests", "cpp", "hostmock")
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"]
    common = ["g++", "-std=c++17", "-O1", "-w", "-fPIC", "-I", "/usr/local/cuda/include", "-I", src] + san
    jobs = [
        common + ["-x", "c++", "-fvisibility=hidden", "-c", os.path.join(src, "api.cu"), "-o", os.path.join(out_dir, "api.o")],
        common + ["-x", "c++", "-fvisibility=hidden", "-c", os.path.join(src, "api_sumcheck.cu"), "-o", os.path.join(out_dir, "api_sumcheck.o")],
        common + ["-x", "c++", "-fvisibility=hidden", "-c", os.path.join(src, "api_ntt.cu"), "-o", os.path.join(out_dir, "api_ntt.o")],
        common + ["-fvisibility=hidden", "-c", os.path.join(mock, "mock_kernels.cpp"), "-o", os.path.join(out_dir, "mock_kernels.o")],
        common + ["-c", os.path.join(mock, "mock_cudart.cpp"), "-o", os.path.join(out_dir, "mock_cudart.o")],
        ["g++", "-std=c++17", "-O3", "-mavx512f", "-mavx512vl", "-fPIC", "-fvisibility=hidden", "-c", os.path.join(src, "keccak_avx512.cpp"),
         "-o", os.path.join(out_dir, "keccak_avx512.o")],
    ]
    procs = [subprocess.Popen(j, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for j in jobs]
    for p, j in zip(procs, jobs):
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, " ".join(j) + "\n" + out[-3000:]
    so = os.path.join(out_dir, "libzk_b200_hostmock.so")
    r = subprocess.run(["g++", "-shared"] + san + ["-o", so] + [os.path.join(out_dir, f) for f in ("api.o", "api_sumcheck.o", "api_ntt.o", "mock_kernels.o", "mock_cudart.o", "keccak_avx512.o")]
                       + ["-ldl", "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-fvisibility=hidden", "-shared"] + san + [os.path.join(mock, "mock_nccl.cpp"),
                        "-o", os.path.join(out_dir, "libnccl.so.2"), "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1",
               LD_LIBRARY_PATH=out_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    for srcfile, args, want in (("test_abi_sanitized.cpp", [], "ABI WALK OK"), ("test_abi_sanitized.cpp", ["sharded"], "SHARDED WALK OK"),
                                ("test_reference_kats.cpp", [], "ALL C++ MIRROR TESTS PASSED"), ("test_sop_mirror_compiles.cpp", ["run"], "GKR LAYER OK")):
        exe = os.path.join(out_dir, srcfile.replace(".cpp", "_asan"))
        cmd = ["g++", "-std=c++17", "-O1"] + san + ["-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", srcfile), "-L", out_dir,
                                                    "-lzk_b200_hostmock", f"-Wl,-rpath,{out_dir}", "-lpthread", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0 and want in r.stdout and "ERROR: " not in r.stderr, (srcfile, r.stdout[-1500:], r.stderr[-3000:])


def test_the_host_mock_is_not_reachable_from_the_product():
    """Nothing under zk_b200/ (the product) or in the Makefile mentions the mock; only an explicit ZK_B200_LIB does."""
    for base, _, files in os.walk(os.path.join(ROOT, "zk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".map")):
                assert "hostmock" not in open(os.path.join(base, f), errors="ignore").read(), f
    assert "hostmock" not in open(os.path.join(ROOT, "Makefile")).read()
    assert "hostmock" not in open(os.path.join(ROOT, "bench.py")).read()
    assert "hostmock" not in open(os.path.join(ROOT, "__graft_entry__.py")).read()
