// intp_b200/BandMatrix.hpp -- host containers for banded and cyclic-banded matrices with the
// reference's names and accessors (src/include/BandMatrix.hpp:19-97 BandMatrix, :99-181
// ExtendedBandMatrix).  Storage differs: one array of rows, row i holding A(i, i-p .. i+q); the
// cyclic matrix keeps its corner entries in the same rows with the column index taken modulo n --
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

template <typename T, typename Alloc = std::allocator<T>>
class BandMatrix {
   public:
    using size_type = std::size_t;
    using val_type = T;
    using allocator_type = Alloc;
    using matrix_type = BandMatrix<val_type, allocator_type>;
    static constexpr bool is_cyclic = false;

    // n x n, p sub-diagonals, q super-diagonals
    BandMatrix(size_type n, size_type p, size_type q) : n_(n), p_(p), q_(q), rows_(n * (p + q + 1), T{}) {}
    BandMatrix() : BandMatrix(0, 0, 0) {}

    size_type dim() const noexcept { return n_; }
    size_type lower_band_width() const noexcept { return p_; }
    size_type upper_band_width() const noexcept { return q_; }
    size_type row_width() const noexcept { return p_ + q_ + 1; }
    // rows()[i * row_width() + k] = A(i, i + k - p)
    const std::vector<T, Alloc>& rows() const noexcept { return rows_; }

    val_type& operator()(size_type i, size_type j) { return rows_[slot(i, j)]; }
    val_type operator()(size_type i, size_type j) const { return rows_[slot(i, j)]; }

    // y = A x for any indexable container constructible from a size
    template <typename Vec>
    util::remove_cvref_t<Vec> operator*(const Vec& x) const {
        util::remove_cvref_t<Vec> y(x.size());
        for (size_type i = 0; i < n_; ++i) {
            const size_type j0 = i > p_ ? i - p_ : 0, j1 = std::min(n_, i + q_ + 1);
            for (size_type j = j0; j < j1; ++j) y[i] += rows_[i * row_width() + (j + p_ - i)] * x[j];
        }
        return y;
    }

    friend std::ostream& operator<<(std::ostream& os, const BandMatrix& m) {
        for (size_type i = 0; i < m.n_; ++i) {
            for (size_type j = 0; j < m.n_; ++j) os << (m.in_band(i, j) ? m(i, j) : T{}) << (j + 1 < m.n_ ? "\t" : "\n");
        }
        return os;
    }

   protected:
    bool in_band(size_type i, size_type j) const { return j + p_ >= i && i + q_ >= j; }
    size_type slot(size_type i, size_type j) const {
        if (i >= n_ || j >= n_ || !in_band(i, j)) throw std::out_of_range("BandMatrix: entry outside the band");
        return i * row_width() + (j + p_ - i);
    }

    size_type n_, p_, q_;
    std::vector<T, Alloc> rows_;
};

// Banded plus the two corner blocks of a periodic collocation matrix: A(i, j) may also be non-zero
// where (j - i) mod n falls in [-p, q].
template <typename T, typename Alloc = std::allocator<T>>
class ExtendedBandMatrix : public BandMatrix<T, Alloc> {
   public:
    using base_type = BandMatrix<T, Alloc>;
    using size_type = typename base_type::size_type;
    using val_type = typename base_type::val_type;
    using allocator_type = typename base_type::allocator_type;
    static constexpr bool is_cyclic = true;

    ExtendedBandMatrix(size_type dim, size_type lower, size_type upper) : base_type(dim, lower, upper) {}
    ExtendedBandMatrix() : ExtendedBandMatrix(1, 0, 0) {}

    val_type& main_bands_val(size_type i, size_type j) { return base_type::operator()(i, j); }
    val_type main_bands_val(size_type i, size_type j) const { return base_type::operator()(i, j); }
    val_type& side_bands_val(size_type i, size_type j) { return rows_[corner_slot(i, j)]; }
    val_type side_bands_val(size_type i, size_type j) const { return rows_[corner_slot(i, j)]; }

    val_type& operator()(size_type i, size_type j) {
        return this->in_band(i, j) ? main_bands_val(i, j) : side_bands_val(i, j);
    }
    val_type operator()(size_type i, size_type j) const {
        return this->in_band(i, j) ? main_bands_val(i, j) : side_bands_val(i, j);
    }

    template <typename Vec>
    util::remove_cvref_t<Vec> operator*(const Vec& x) const {
        util::remove_cvref_t<Vec> y(x.size());
        const size_type w = this->row_width();
        for (size_type i = 0; i < n_; ++i)
            for (size_type k = 0; k < w; ++k) {
                // column i + k - p, wrapped; a wrapped column that re-enters the band is not stored twice
                const size_type j = (i + k + n_ - p_) % n_;
                const bool wrapped = i + k < p_ || i + k - p_ >= n_;
                if (wrapped && this->in_band(i, j)) continue;
                y[i] += rows_[i * w + k] * x[j];
            }
        return y;
    }

   private:
    // corner entry: above the band in the last p columns, or below it in the last q rows
    size_type corner_slot(size_type i, size_type j) const {
        if (i < n_ && j < n_) {
            if (j > i + q_ && j + p_ >= n_ + i) return i * this->row_width() + (j + p_ - n_ - i);
            if (i > j + p_ && i + q_ >= n_ + j) return i * this->row_width() + (j + n_ + p_ - i);
        }
        throw std::out_of_range("ExtendedBandMatrix: entry outside the band and its corners");
    }

    using base_type::n_;
    using base_type::p_;
    using base_type::q_;
    using base_type::rows_;
};

}  // namespace intp

#endif  // INTP_B200_BAND_MATRIX_HPP
