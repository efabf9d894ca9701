"""include/ppf_b200.hpp (the C++ mirror of Scene / Model) compiles and links against the library; on a box
without a GPU the first CUDA call must surface as a C++ exception, not a process exit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SRC = r'''
#include <cstdio>
#include <vector>
#include "ppf_b200.hpp"
int main() {
    std::vector<float> p = {1,1,1, 2,1,1, 1,2,1, 1,1,2}, n = {0,0,1, 0,1,0, 1,0,0, 0,0,1};
    ppf_b200::CloudView c{p.data(), 3, n.data(), 3, 4};
    try {
        ppf_b200::Model m(c, 0.5f, 0.4f, false, false, false);
        ppf_b200::Scene s(c, 0.5f, 1);
        bool ok = m.ppf_lookup(&s);
        std::printf("lookup ok=%d K=%zu\n", (int)ok, m.votes.size());
    } catch (const std::exception &e) {
        std::printf("exception: %s\n", e.what());
        return 3;
    }
    return 0;
}
'''


def test_cpp_mirror_compiles_links_and_reports_errors(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "objective_slam_b200", "lib")
    r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", lib, "-lppf_b200", f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and "lookup ok=" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "exception:" in r.stdout, r.stdout + r.stderr
