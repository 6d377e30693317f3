// User code as the reference's README.md:35-50 shows it: the installed include line, namespace intp,
// InterpolationFunction<double, 2, 3, double> from a mesh and two ranges.
#include <BSplineInterpolation/Interpolation.hpp>

#include <cmath>
#include <cstdio>

using namespace intp;

int main() {
    Mesh<double, 2> z_mesh(16, 20);
    for (std::size_t i = 0; i < 16; ++i)
        for (std::size_t j = 0; j < 20; ++j) z_mesh(i, j) = std::sin(0.3 * double(i)) * std::cos(0.2 * double(j));
    const double x_min = 0., x_max = 1.5, y_min = -1., y_max = 1.;
    InterpolationFunction<double, 2, 3, double> func(
        // mesh storing z values, an object of type intp::Mesh<double, 2>
        z_mesh,
        // x range
        std::make_pair(x_min, x_max),
        // y range
        std::make_pair(y_min, y_max));
    const double at_node = func(x_min + 5 * (x_max - x_min) / 15, y_min + 7 * (y_max - y_min) / 19);
    std::printf("f(node 5,7) = %.15f (sample %.15f)\n", at_node, z_mesh(5, 7));
    return std::abs(at_node - z_mesh(5, 7)) < 1e-13 ? 0 : 1;
}
