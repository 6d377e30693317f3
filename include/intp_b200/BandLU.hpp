// intp_b200/BandLU.hpp -- the reference's band solver interface (src/include/BandLU.hpp:15-262:
// BandLU<BandMatrix<..>> and BandLU<ExtendedBandMatrix<..>>, compute / solve / solve_in_place) on top of
// the device library: the matrix is factored without pivoting in the reference's elimination order
// (BandLU.hpp:103-118, :159-213) and the substitution runs on the GPU through
// bspl_band_solve_rows().  Link libbspline_b200.so.
#ifndef INTP_B200_BAND_LU_HPP
#define INTP_B200_BAND_LU_HPP

#include <cstdint>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../bspline_b200.h"
#include "BandMatrix.hpp"

namespace intp {

template <typename Matrix>
class BandLU {
   public:
    using matrix_type = Matrix;
    using size_type = typename matrix_type::size_type;

    BandLU() noexcept = default;
    template <typename M, typename = std::enable_if_t<std::is_same_v<util::remove_cvref_t<M>, Matrix>>>
    BandLU(M&& mat) { compute(std::forward<M>(mat)); }

    template <typename M>
    void compute(M&& mat) {
        if (computed_) return;  // like the reference, a solver is computed once (BandLU.hpp:33-40)
        mat_ = std::forward<M>(mat);
        rows_.assign(mat_.rows().begin(), mat_.rows().end());
        computed_ = true;
    }

    // returns the solution in a copy of `vec`; a pointer argument is solved in place
    template <typename Vec>
    util::remove_cvref_t<Vec> solve(Vec&& vec) const {
        util::remove_cvref_t<Vec> x(std::forward<Vec>(vec));
        solve_in_place(x);
        return x;
    }

    // `line` is anything indexable (container, pointer, random-access iterator)
    template <typename Line>
    void solve_in_place(Line&& line) const {
        if (!computed_) throw std::runtime_error("BandLU: no matrix has been factored");
        const size_type n = mat_.dim();
        std::vector<double> x(n);
        for (size_type i = 0; i < n; ++i) x[i] = static_cast<double>(line[i]);
        const int rc = bspl_band_solve_rows(static_cast<int64_t>(n), static_cast<int64_t>(mat_.lower_band_width()),
                                            static_cast<int64_t>(mat_.upper_band_width()), Matrix::is_cyclic ? 1 : 0,
                                            rows_.data(), x.data(), 1, device_);
        if (rc == BSPL_ERR_ALLOC) throw std::bad_alloc();
        if (rc != BSPL_OK) throw std::runtime_error(bspl_last_error());
        using elem = util::remove_cvref_t<decltype(line[0])>;
        for (size_type i = 0; i < n; ++i) line[i] = static_cast<elem>(x[i]);
    }

    void set_device(int ordinal) { device_ = ordinal; }

   private:
    bool computed_ = false;
    int device_ = 0;
    matrix_type mat_;
    std::vector<double> rows_;
};

}  // namespace intp

#endif  // INTP_B200_BAND_LU_HPP
