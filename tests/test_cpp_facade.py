"""include/claxon_b200.hpp compiled and RUN (g++, linked against the product library): claxon::FlacReader::open ->
blocks() -> read_next_or_eof / read_batch on the reference's pop.flac and on a synthetic stereo file, checked against
the goldens and the generator's PCM."""
import os
import subprocess

import numpy as np
import pytest

from claxon_b200 import _build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tools", "scratch", "facade_test")
SRC = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")


def build_exe():
    lib = _build.build_lib()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "include", "claxon_b200.hpp"), os.path.join(ROOT, "include", "claxon_b200.h"), lib]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", EXE, SRC,
                               lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def test_cpp_facade_compiles_and_links():
    """No GPU needed: the header is valid C++17 and every symbol it uses is exported by the library."""
    build_exe()


@pytest.mark.gpu
def test_cpp_facade_decodes_like_the_goldens(golden, tmp_path):
    exe = build_exe()
    cases = []
    p = tmp_path / "pop.flac"
    p.write_bytes(golden["pop__bytes"].tobytes())
    cases.append((p, 1, 16, golden["pop__pcm"], int((golden["pop__frames"][:, 1] == 0).sum())))
    b = synth.workload("c4", 33)
    q = tmp_path / "synth.flac"
    q.write_bytes(synth.make_file(b, 0, 33, padding=64))
    cases.append((q, 2, 16, b.pcm, 33))
    for path, ch, bits, pcm, frames in cases:
        out = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.strip().splitlines()
        assert len(lines) == 2 and lines[0] == lines[1]  # frame by frame == batched
        got = [int(v) for v in lines[0].split()]
        x = int(np.bitwise_xor.reduce(pcm.astype(np.int32))) if pcm.size else 0
        assert got == [ch, bits, frames, pcm.size, int(pcm.astype(np.int64).sum()), x]
