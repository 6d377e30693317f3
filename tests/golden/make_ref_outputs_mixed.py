#!/usr/bin/env python3
"""Generate tests/golden/ref_outputs_mixed.json: outputs of the UNMODIFIED reference instantiated with
float values on double coordinates -- InterpolationFunction<float, D, 3> with its default U = double --
on seeded inputs (a throw-away C++ generator compiled against the reference headers where they lie).
Build container only; the JSON is committed and read by tests/cpp/drop_in_test2.cpp through
tests/test_cpp_dropin.py."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

GEN = r'''
#include <Interpolation.hpp>
#include <cstdio>
#include <random>
#include <vector>
using namespace intp;
static void dump(const char* name, const std::vector<double>& v) {
    std::printf("\"%s\": [", name);
    for (std::size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
    std::printf("]");
}
int main() {
    std::mt19937_64 gen(4242);
    std::uniform_real_distribution<> uni(0., 1.);
    std::printf("{\n");
    {   // 1-D cubic, 23 samples on [0, 2]
        std::vector<float> f(23);
        for (auto& v : f) v = static_cast<float>(2. * uni(gen) - 1.);
        InterpolationFunction<float, 1, 3> g(std::make_pair(f.begin(), f.end()), std::make_pair(0., 2.));
        std::vector<double> fd(f.begin(), f.end()), x, val, d1;
        for (int i = 0; i < 40; ++i) {
            x.push_back(2. * uni(gen));
            val.push_back(g(x.back()));
            d1.push_back(g.derivative({x.back()}, 1));
        }
        dump("mixed1_f", fd); std::printf(",\n"); dump("mixed1_x", x); std::printf(",\n");
        dump("mixed1_val", val); std::printf(",\n"); dump("mixed1_d1", d1); std::printf(",\n");
    }
    {   // 2-D cubic 9 x 11 on [0, 1] x [-1, 1], periodic along y (closing sample implicit)
        Mesh<float, 2> m(9, 11);
        std::vector<double> fd, xy, val;
        for (std::size_t i = 0; i < 9; ++i)
            for (std::size_t j = 0; j < 11; ++j) { m(i, j) = static_cast<float>(2. * uni(gen) - 1.); fd.push_back(m(i, j)); }
        InterpolationFunction<float, 2, 3> g({false, true}, m, std::make_pair(0., 1.), std::make_pair(-1., 1.));
        for (int i = 0; i < 40; ++i) {
            const double x = uni(gen), y = 2. * uni(gen) - 1.;
            xy.push_back(x); xy.push_back(y);
            val.push_back(g(x, y));
        }
        dump("mixed2_f", fd); std::printf(",\n"); dump("mixed2_xy", xy); std::printf(",\n"); dump("mixed2_val", val);
    }
    std::printf("\n}\n");
    return 0;
}
'''

with tempfile.TemporaryDirectory() as tmp:
    src = os.path.join(tmp, "gen.cpp")
    open(src, "w").write(GEN)
    exe = os.path.join(tmp, "gen")
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O2", "-DINTP_PERIODIC_NO_DUMMY_POINT", "-I" + REF + "/src/include",
                           src, "-o", exe])
    data = json.loads(subprocess.check_output([exe]).decode())
with open(os.path.join(HERE, "ref_outputs_mixed.json"), "w") as fh:
    json.dump(data, fh, indent=0)
print("wrote ref_outputs_mixed.json:", {k: len(v) for k, v in data.items()})
