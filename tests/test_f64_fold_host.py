"""CPU suite: the FP64 fixed-multiplier product of the fold (zk_b200/csrc/field_f64.cuh) replayed on the host.

The device folds l + r (h - l) with exact double-precision dot products (every intermediate an integer < 2^53), so the
same header compiles for the host and its arithmetic can be checked bit for bit without a GPU against the word-serial
Montgomery multiplier of host_field.hpp: the reference's `left - a * (left - right)`
(polynomial/src/multilinear/evaluation_form.rs:68) for random and extreme operands, both fields."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_f64_fold_matches_host_multiplier(tmp_path):
    exe = str(tmp_path / "test_f64_fold")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "zk_b200", "csrc"),
           os.path.join(ROOT, "tests", "cpp", "test_f64_fold.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, timeout=300, text=True).stdout
    assert "0 mismatches" in out, out
    assert "4800000 checked" in out, out


def test_tensor_path_fold_model_matches_host_field(tmp_path):
    """The INT8 tensor-path fold (zk_b200/csrc/fold_imma.cuh): the real host table builder (fragment order) and a CPU
    restatement of one warp's data flow (staging swizzle, ldmatrix.x4, mma.m16n8k32 fragment layouts, pair words, assembly,
    one Montgomery row) against evaluation_form.rs:68 computed with the host field, both fields, extreme operands; also
    asserts the bounds the exactness argument rests on (column sums < 2^21, pair words < 2^30)."""
    exe = str(tmp_path / "test_fold_imma_host")
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "zk_b200", "csrc"),
           os.path.join(ROOT, "tests", "cpp", "test_fold_imma_host.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, timeout=300, text=True).stdout
    assert "204800 checked, 0 mismatches" in out, out


def test_keccak_avx512_matches_portable(tmp_path):
    """The AVX-512 absorb loop of the host transcript (zk_b200/csrc/keccak_avx512.cpp) against the portable
    Keccak-f[1600] of keccak.hpp: KATs, every length 0..1100 in six chunkings, large messages, digest chaining.
    (Prints "skipped" and passes on a CPU without AVX-512.)"""
    obj, exe = str(tmp_path / "keccak_avx512.o"), str(tmp_path / "test_keccak_avx512")
    src = os.path.join(ROOT, "zk_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O3", "-mavx512f", "-mavx512vl", "-c", os.path.join(src, "keccak_avx512.cpp"), "-o", obj],
                   check=True, capture_output=True, timeout=300)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", src, os.path.join(ROOT, "tests", "cpp", "test_keccak_avx512.cpp"), obj, "-o", exe],
                   check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, timeout=600, text=True).stdout
    assert "skipped" in out or "0 mismatches" in out, out
