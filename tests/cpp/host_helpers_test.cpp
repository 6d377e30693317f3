// Host-side helpers of include/intp_b200/Interpolation.hpp that never touch the device (CPU suite):
// dropping the dummy sample of periodic axes, and carrying value types other than the coordinate
// type (converted scalars, aggregates as fields).
#include <intp_b200/Interpolation.hpp>

#include <cstdio>
#include <vector>

using namespace intp;

static int failures = 0;
static void expect(bool ok, const char* what) {
    if (!ok) { std::printf("FAILED: %s\n", what); ++failures; }
}

struct Vec2f { float x, y; };

int main() {
    {   // strip_dummy: 3 x 4 x 3 with axes 0 and 2 periodic -> 2 x 4 x 2
        Mesh<int, 3> m(3, 4, 3);
        for (std::size_t i = 0; i < m.size(); ++i) m.data()[i] = int(i);
        const auto kept = b200_detail::strip_dummy<int, 3>(m.data(), m.dimension(), {true, false, true});
        bool ok = kept.size() == 2 * 4 * 2;
        std::size_t n = 0;
        for (std::size_t i = 0; i < 2; ++i)
            for (std::size_t j = 0; j < 4; ++j)
                for (std::size_t k = 0; k < 2; ++k) ok = ok && kept[n++] == m(i, j, k);
        expect(ok, "strip_dummy keeps all but the last index of periodic axes, in row-major order");
        const auto same = b200_detail::strip_dummy<int, 3>(m.data(), m.dimension(), {false, false, false});
        expect(same.size() == m.size() && same[17] == 17, "strip_dummy without periodic axes is a copy");
    }
    {   // converted scalars: float values on double coordinates
        static_assert(b200_detail::components_of<float, double>::value == 1, "one field");
        static_assert(!b200_detail::direct_v<float, double> && b200_detail::direct_v<double, double>, "direct only when T == U");
        const float in[3] = {1.5f, -2.25f, 1e-3f};
        const auto d = b200_detail::split_components<float, double>(in, 3);
        expect(d.size() == 3 && d[0] == 1.5 && d[1] == -2.25 && d[2] == double(1e-3f), "float -> double on the way in");
        float out[3] = {};
        const double res[3] = {0.1, 2.0, -7.5};
        b200_detail::merge_component<float, double>(res, 3, 0, out);
        expect(out[0] == 0.1f && out[1] == 2.0f && out[2] == -7.5f, "double -> float on the way out");
        expect(b200_detail::from_components<float, double>(res) == 0.1f, "single value");
    }
    {   // aggregates as fields: [m][K] interleaved <-> [K][m]
        static_assert(b200_detail::components_of<Vec2f, float>::value == 2, "two fields");
        const Vec2f in[3] = {{1, 2}, {3, 4}, {5, 6}};
        const auto f = b200_detail::split_components<Vec2f, float>(in, 3);
        expect(f.size() == 6 && f[0] == 1 && f[1] == 3 && f[2] == 5 && f[3] == 2 && f[4] == 4 && f[5] == 6, "split [m][K] -> [K][m]");
        Vec2f out[3] = {};
        b200_detail::merge_component<Vec2f, float>(f.data(), 3, 0, out);
        b200_detail::merge_component<Vec2f, float>(f.data() + 3, 3, 1, out);
        expect(out[1].x == 3 && out[1].y == 4 && out[2].y == 6, "merge back");
        const float c[2] = {7, 8};
        const Vec2f v = b200_detail::from_components<Vec2f, float>(c);
        expect(v.x == 7 && v.y == 8, "one aggregate from its components");
    }
    std::printf("%s\n", failures ? "FAILED" : "all host helper checks passed");
    return failures ? 1 : 0;
}
