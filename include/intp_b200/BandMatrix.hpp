// intp_b200/BandMatrix.hpp -- host containers for banded and cyclic-banded matrices under the
// reference's names (src/include/BandMatrix.hpp:19-97 BandMatrix, :99-181 ExtendedBandMatrix).
// One storage scheme serves both: an array of rows, row i holding A(i, i-p .. i+q); the cyclic
// matrix keeps its corner entries in the same rows with the column index taken modulo n --
// exactly the form bspl_band_solve_rows() consumes.  Pure host code.
#ifndef INTP_B200_BAND_MATRIX_HPP
#define INTP_B200_BAND_MATRIX_HPP

#include <algorithm>
#include <cstddef>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <vector>

#include "util.hpp"

namespace intp {
namespace b200_detail {

// n x n matrix with p sub- and q super-diagonals; Cyclic adds the entries whose column distance
// from the diagonal, taken modulo n, falls in [-p, q] (the corner blocks of a periodic collocation
// matrix).  slot(i, j) is the only place that knows the layout.
template <typename T, typename Alloc, bool Cyclic>
class RowBand {
   public:
    using size_type = std::size_t;
    using val_type = T;
    using allocator_type = Alloc;
    static constexpr bool is_cyclic = Cyclic;
    static constexpr size_type npos = static_cast<size_type>(-1);

    RowBand(size_type n, size_type sub, size_type super) : n_(n), p_(sub), q_(super), rows_(n * (sub + super + 1), T{}) {}

    size_type dim() const noexcept { return n_; }
    size_type lower_band_width() const noexcept { return p_; }
    size_type upper_band_width() const noexcept { return q_; }
    size_type row_width() const noexcept { return p_ + q_ + 1; }
    // rows()[i * row_width() + k] = A(i, i + k - p), the column wrapping modulo n when cyclic
    const std::vector<T, Alloc>& rows() const noexcept { return rows_; }

    bool in_band(size_type i, size_type j) const { return j + p_ >= i && i + q_ >= j; }
    // position of A(i, j) in rows(), npos when the matrix has no such entry
    size_type slot(size_type i, size_type j) const {
        if (i >= n_ || j >= n_) return npos;
        if (in_band(i, j)) return i * row_width() + (j + p_ - i);
        if (Cyclic && j > i + q_ && j + p_ >= n_ + i) return i * row_width() + (j + p_ - n_ - i);  // upper right corner
        if (Cyclic && i > j + p_ && i + q_ >= n_ + j) return i * row_width() + (j + n_ + p_ - i);  // lower left corner
        return npos;
    }
    T& entry(size_type i, size_type j) { return rows_[checked(i, j)]; }
    T entry(size_type i, size_type j) const { return rows_[checked(i, j)]; }

    // y = A x for any indexable container constructible from a size
    template <typename Vec>
    util::remove_cvref_t<Vec> times(const Vec& x) const {
        util::remove_cvref_t<Vec> y(x.size());
        const size_type w = row_width();
        for (size_type i = 0; i < n_; ++i)
            for (size_type k = 0; k < w; ++k) {
                const bool outside = i + k < p_ || i + k - p_ >= n_;
                if (outside && !Cyclic) continue;
                const size_type j = (i + k + n_ - p_) % n_;
                if (outside && in_band(i, j)) continue;  // a wrapped column that re-enters the band is not stored twice
                y[i] += rows_[i * w + k] * x[j];
            }
        return y;
    }

    void print(std::ostream& os) const {
        for (size_type i = 0; i < n_; ++i)
            for (size_type j = 0; j < n_; ++j) {
                const size_type s = slot(i, j);
                os << (s == npos ? T{} : rows_[s]) << (j + 1 < n_ ? '\t' : '\n');
            }
    }

   private:
    size_type checked(size_type i, size_type j) const {
        const size_type s = slot(i, j);
        if (s == npos) throw std::out_of_range(Cyclic ? "entry outside the band and its corners" : "entry outside the band");
        return s;
    }
    size_type n_, p_, q_;
    std::vector<T, Alloc> rows_;
};

}  // namespace b200_detail

template <typename T, typename Alloc = std::allocator<T>>
class BandMatrix : public b200_detail::RowBand<T, Alloc, false> {
    using rows_type = b200_detail::RowBand<T, Alloc, false>;

   public:
    using matrix_type = BandMatrix<T, Alloc>;
    using typename rows_type::size_type;
    BandMatrix(size_type n, size_type p, size_type q) : rows_type(n, p, q) {}
    BandMatrix() : rows_type(0, 0, 0) {}

    T& operator()(size_type i, size_type j) { return this->entry(i, j); }
    T operator()(size_type i, size_type j) const { return this->entry(i, j); }
    template <typename Vec>
    util::remove_cvref_t<Vec> operator*(const Vec& x) const { return this->times(x); }
    friend std::ostream& operator<<(std::ostream& os, const BandMatrix& m) { m.print(os); return os; }
};

template <typename T, typename Alloc = std::allocator<T>>
class ExtendedBandMatrix : public b200_detail::RowBand<T, Alloc, true> {
    using rows_type = b200_detail::RowBand<T, Alloc, true>;

   public:
    using matrix_type = ExtendedBandMatrix<T, Alloc>;
    using typename rows_type::size_type;
    ExtendedBandMatrix(size_type dim, size_type lower, size_type upper) : rows_type(dim, lower, upper) {}
    ExtendedBandMatrix() : rows_type(1, 0, 0) {}

    T& operator()(size_type i, size_type j) { return this->entry(i, j); }
    T operator()(size_type i, size_type j) const { return this->entry(i, j); }
    // the reference's two accessors: entries of the band proper / of the corner blocks
    T& main_bands_val(size_type i, size_type j) { return this->entry(expect_band(i, j, true), j); }
    T main_bands_val(size_type i, size_type j) const { return this->entry(expect_band(i, j, true), j); }
    T& side_bands_val(size_type i, size_type j) { return this->entry(expect_band(i, j, false), j); }
    T side_bands_val(size_type i, size_type j) const { return this->entry(expect_band(i, j, false), j); }
    template <typename Vec>
    util::remove_cvref_t<Vec> operator*(const Vec& x) const { return this->times(x); }
    friend std::ostream& operator<<(std::ostream& os, const ExtendedBandMatrix& m) { m.print(os); return os; }

   private:
    size_type expect_band(size_type i, size_type j, bool band) const {
        if (this->in_band(i, j) != band) throw std::out_of_range(band ? "not a band entry" : "not a corner entry");
        return i;
    }
};

}  // namespace intp

#endif  // INTP_B200_BAND_MATRIX_HPP
