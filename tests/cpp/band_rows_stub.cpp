// TEST INFRASTRUCTURE ONLY (CPU suite): a stand-in for the two library entry points BandLU.hpp calls,
// so that the host containers of include/intp_b200/BandMatrix.hpp / BandLU.hpp -- and the row form
// they hand to bspl_band_solve_rows() -- can be checked without a GPU.  It expands the rows into a
// dense matrix exactly as the header documents them and solves with pivoted Gaussian elimination.
// Never linked into the product.
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "bspline_b200.h"

extern "C" const char* bspl_last_error(void) { return "band_rows_stub: singular matrix or bad argument"; }

extern "C" int bspl_band_solve_rows(int64_t n, int64_t p, int64_t q, int cyclic, const double* rows, double* x,
                                    int64_t n_rhs, int) {
    if (!rows || !x || n < 1 || n_rhs < 1) return BSPL_ERR_INVALID;
    const int64_t w = p + q + 1;
    std::vector<double> a(static_cast<size_t>(n * n), 0.0);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = 0; k < w; ++k) {
            int64_t j = i + k - p;
            if (j < 0 || j >= n) {
                if (!cyclic) continue;
                j += j < 0 ? n : -n;
            }
            a[i * n + j] = rows[i * w + k];
        }
    for (int64_t r = 0; r < n_rhs; ++r) {
        std::vector<double> m(a);
        double* b = x + r * n;
        for (int64_t c = 0; c < n; ++c) {
            int64_t piv = c;
            for (int64_t i = c + 1; i < n; ++i)
                if (std::fabs(m[i * n + c]) > std::fabs(m[piv * n + c])) piv = i;
            if (m[piv * n + c] == 0.0) return BSPL_ERR_INVALID;
            if (piv != c) {
                for (int64_t j = 0; j < n; ++j) std::swap(m[c * n + j], m[piv * n + j]);
                std::swap(b[c], b[piv]);
            }
            for (int64_t i = c + 1; i < n; ++i) {
                const double l = m[i * n + c] / m[c * n + c];
                if (l == 0.0) continue;
                for (int64_t j = c; j < n; ++j) m[i * n + j] -= l * m[c * n + j];
                b[i] -= l * b[c];
            }
        }
        for (int64_t i = n; i-- > 0;) {
            for (int64_t j = i + 1; j < n; ++j) b[i] -= m[i * n + j] * b[j];
            b[i] /= m[i * n + i];
        }
    }
    return BSPL_OK;
}
