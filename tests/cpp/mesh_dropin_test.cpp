// Mesh / MeshDimension of include/intp_b200/Mesh.hpp (reference: src/include/Mesh.hpp:11-383) -- pure host
// code, runs in the CPU suite: row-major indexing with the last index fastest, bounds-checked
// access, line iterators along every axis (random access), the 1-D iterator-pair constructor,
// resize, allocator-converting copy, util::pow.
#include <Mesh.hpp>

#include <algorithm>
#include <cstdio>
#include <numeric>
#include <vector>

using namespace intp;

static int failures = 0;
static void expect(bool ok, const char* what) {
    if (!ok) { std::printf("FAILED: %s\n", what); ++failures; }
}

int main() {
    Mesh<int, 3> m(4, 5, 6);
    expect(m.size() == 120 && m.dim_size(0) == 4 && m.dim_size(1) == 5 && m.dim_size(2) == 6, "extents");
    std::iota(m.data(), m.data() + m.size(), 0);
    expect(m(1, 2, 3) == 1 * 30 + 2 * 6 + 3, "row-major, last index fastest");
    expect(m({3, 4, 5}) == 119, "array index");
    expect(m.dimension().indexing(2, 0, 1) == 61 && m.dimension().dim_acc_size(2) == 30, "MeshDimension indexing");
    const auto idx = m.dimension().dimwise_indices(77);
    expect(idx[0] == 2 && idx[1] == 2 && idx[2] == 5, "dimwise_indices");
    bool threw = false;
    try { m(1, 5, 0); } catch (const std::exception&) { threw = true; }
    expect(threw, "out-of-range access throws");

    // a line along each axis
    for (std::size_t d = 0; d < 3; ++d) {
        const Mesh<int, 3>::index_type at{2, 3, 4};
        auto b = m.begin(d, at), e = m.end(d, at);
        expect(static_cast<std::size_t>(e - b) == m.dim_size(d), "line length");
        std::size_t k = 0;
        bool ok = true;
        for (auto it = b; it != e; ++it, ++k) {
            auto where = at;
            where[d] = k;
            ok = ok && *it == m(where) && b[static_cast<std::ptrdiff_t>(k)] == *it;
        }
        expect(ok && k == m.dim_size(d), "line iterator visits the line in order");
        expect(*(b + 2) == *(2 + b) && *(e - 1) == b[static_cast<std::ptrdiff_t>(m.dim_size(d)) - 1] && b < e, "random access");
        const auto ii = m.iter_indices(Mesh<int, 3>::line_iterator<const int>(b + 1));
        auto where = at;
        where[d] = 1;
        expect(ii == where, "iter_indices of a line iterator");
    }
    // writing through a line iterator, std algorithms on it
    std::fill(m.begin(1, {0, 0, 2}), m.end(1, {0, 0, 2}), -7);
    expect(m(0, 0, 2) == -7 && m(0, 4, 2) == -7 && m(0, 0, 1) == 1 && m(1, 0, 2) == 32, "std::fill along axis 1");
    const Mesh<int, 3>& cm = m;
    expect(std::count(cm.begin(1, {0, 0, 2}), cm.end(1, {0, 0, 2}), -7) == 5, "const line iterators");
    auto flat = cm.begin();
    std::advance(flat, 45);
    const auto fi = cm.iter_indices(flat);
    expect(fi[0] == 1 && fi[1] == 2 && fi[2] == 3, "iter_indices of the flat iterator");

    // 1-D from an iterator pair; equal-extent constructor; resize; allocator-converting copy
    const std::vector<double> v{1, 1, 2, 3, 5, 8};
    Mesh<double, 1> m1(std::make_pair(v.begin(), v.end()));
    expect(m1.size() == 6 && m1(4) == 5., "1-D from iterators");
    Mesh<int, 4> m4(3);
    expect(m4.size() == util::pow(3u, 4u) && m4.dim_size(3) == 3, "equal extents, util::pow");
    m.resize({2, 3, 4});
    expect(m.size() == 24 && m.dim_size(1) == 3, "resize");
    Mesh<double, 1, std::allocator<double>> copy(m1);
    expect(copy.size() == 6 && copy(5) == 8., "copy from a mesh with another allocator type");

    std::printf("%s\n", failures ? "FAILED" : "all mesh checks passed");
    return failures ? 1 : 0;
}
