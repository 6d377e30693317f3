// intp_b200/Mesh.hpp -- host container of the drop-in API: the row-major dense N-d array the
// reference passes sampled fields in (src/include/Mesh.hpp:11-116 MeshDimension, :125-383 Mesh).
// Same class names, constructors and accessors; the last index is the fastest one
// (Mesh.hpp:241-246), which is also the layout the device library expects for `f`.
// Pure host code: nothing here touches the GPU.
#ifndef INTP_B200_MESH_HPP
#define INTP_B200_MESH_HPP

#include <array>
#include <cstddef>
#include <functional>
#include <iterator>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "util.hpp"

namespace intp {

template <std::size_t D>
class MeshDimension {
   public:
    using size_type = std::size_t;
    static constexpr size_type dim = D;
    using index_type = std::array<size_type, D>;

    MeshDimension() : extent_{} {}
    MeshDimension(index_type extent) : extent_(extent) {}
    MeshDimension(size_type n) { extent_.fill(n); }
    template <typename... Args, typename = std::enable_if_t<sizeof...(Args) == D && (D > 1)>>
    MeshDimension(Args... n) : extent_{static_cast<size_type>(n)...} {}

    size_type size() const {
        return std::accumulate(extent_.begin(), extent_.end(), size_type{1}, std::multiplies<size_type>());
    }
    size_type dim_size(size_type d) const { return extent_[d]; }
    size_type& dim_size(size_type d) { return extent_[d]; }
    // product of the last k extents (Mesh.hpp:57-64): the element stride of axis D-1-k
    size_type dim_acc_size(size_type k) const {
        size_type s = 1;
        for (size_type d = 0; d < k && d < D; ++d) s *= extent_[D - 1 - d];
        return s;
    }
    operator index_type() const { return extent_; }

    // row-major, last index fastest
    size_type indexing(const index_type& idx) const {
        size_type lin = 0;
        for (size_type d = 0; d < D; ++d) lin = lin * extent_[d] + idx[d];
        return lin;
    }
    template <typename... Idx, typename = std::enable_if_t<(std::is_integral_v<Idx> && ...)>>
    size_type indexing(Idx... idx) const {
        static_assert(sizeof...(Idx) == D, "one index per dimension");
        return indexing(index_type{static_cast<size_type>(idx)...});
    }
    size_type indexing_safe(const index_type& a) const {
        for (size_type d = 0; d < D; ++d)
            if (a[d] >= extent_[d]) throw std::runtime_error("Mesh access out of range at dim " + std::to_string(d));
        return indexing(a);
    }
    template <typename... Idx, typename = std::enable_if_t<(std::is_integral_v<Idx> && ...)>>
    size_type indexing_safe(Idx... idx) const {
        static_assert(sizeof...(Idx) == D, "one index per dimension");
        return indexing_safe(index_type{static_cast<size_type>(idx)...});
    }
    index_type dimwise_indices(size_type lin) const {
        index_type idx{};
        for (size_type d = D; d-- > 0;) { idx[d] = lin % extent_[d]; lin /= extent_[d]; }
        return idx;
    }
    void resize(index_type extent) { extent_ = extent; }

   private:
    index_type extent_;
};

template <typename T, std::size_t D, typename Alloc = std::allocator<T>>
class Mesh {
   public:
    using size_type = std::size_t;
    using val_type = T;
    static constexpr size_type dim = D;
    using index_type = typename MeshDimension<D>::index_type;
    using allocator_type = Alloc;
    using const_iterator = typename std::vector<T, Alloc>::const_iterator;

    // Random-access iterator over one mesh line (all indices fixed but one): element k of the line
    // sits k * stride elements after its first one (Mesh.hpp:138-238, skip_iterator).
    template <typename V>
    class line_iterator {
       public:
        using value_type = std::remove_const_t<V>;
        using difference_type = std::ptrdiff_t;
        using pointer = V*;
        using reference = V&;
        using iterator_category = std::random_access_iterator_tag;

        line_iterator() = default;
        line_iterator(pointer at, difference_type stride) : at_(at), stride_(stride) {}
        operator line_iterator<const V>() const { return {at_, stride_}; }
        explicit operator const V*() const { return at_; }

        reference operator*() const { return *at_; }
        pointer operator->() const { return at_; }
        reference operator[](difference_type k) const { return at_[k * stride_]; }

        line_iterator& operator+=(difference_type k) { at_ += k * stride_; return *this; }
        line_iterator& operator-=(difference_type k) { return *this += -k; }
        line_iterator& operator++() { return *this += 1; }
        line_iterator& operator--() { return *this += -1; }
        line_iterator operator++(int) { line_iterator t(*this); ++*this; return t; }
        line_iterator operator--(int) { line_iterator t(*this); --*this; return t; }
        friend line_iterator operator+(line_iterator it, difference_type k) { return it += k; }
        friend line_iterator operator+(difference_type k, line_iterator it) { return it += k; }
        friend line_iterator operator-(line_iterator it, difference_type k) { return it -= k; }
        friend difference_type operator-(const line_iterator& a, const line_iterator& b) {
            return (a.at_ - b.at_) / a.stride_;
        }
        friend bool operator==(const line_iterator& a, const line_iterator& b) {
            return a.at_ == b.at_ && a.stride_ == b.stride_;
        }
        friend bool operator!=(const line_iterator& a, const line_iterator& b) { return !(a == b); }
        friend bool operator<(const line_iterator& a, const line_iterator& b) { return (b - a) > 0; }
        friend bool operator>(const line_iterator& a, const line_iterator& b) { return b < a; }
        friend bool operator<=(const line_iterator& a, const line_iterator& b) { return !(b < a); }
        friend bool operator>=(const line_iterator& a, const line_iterator& b) { return !(a < b); }

       private:
        pointer at_ = nullptr;
        difference_type stride_ = 1;
    };
    template <typename V>
    using skip_iterator = line_iterator<V>;

    explicit Mesh(const MeshDimension<D>& md, const Alloc& alloc = Alloc()) : dims_(md), data_(md.size(), T{}, alloc) {}
    explicit Mesh(size_type n, const Alloc& alloc = Alloc()) : Mesh(MeshDimension<D>(n), alloc) {}
    template <typename... Args, typename = std::enable_if_t<sizeof...(Args) == D && (D > 1) &&
                                                            (std::is_integral_v<Args> && ...)>>
    explicit Mesh(Args... n) : Mesh(MeshDimension<D>(static_cast<size_type>(n)...)) {}
    // 1-D: from a pair of iterators (Mesh.hpp:265-275)
    template <typename It, typename = std::enable_if_t<D == 1 && std::is_convertible_v<
                               typename std::iterator_traits<It>::iterator_category, std::input_iterator_tag>>>
    explicit Mesh(std::pair<It, It> range, const Alloc& alloc = Alloc())
        : dims_(size_type{0}), data_(range.first, range.second, alloc) {
        dims_ = MeshDimension<D>(data_.size());
    }
    // same content, another allocator (Mesh.hpp:277-283)
    template <typename A2>
    Mesh(const Mesh<T, D, A2>& other, const Alloc& alloc = Alloc())
        : dims_(other.dimension()), data_(other.begin(), other.end(), alloc) {}

    size_type size() const { return data_.size(); }
    size_type dim_size(size_type d) const { return dims_.dim_size(d); }
    const MeshDimension<D>& dimension() const { return dims_; }
    void resize(index_type extent) { dims_.resize(extent); data_.resize(dims_.size()); }

    template <typename... Idx, typename = std::enable_if_t<sizeof...(Idx) == D && (std::is_integral_v<Idx> && ...)>>
    T& operator()(Idx... i) { return data_[dims_.indexing_safe(i...)]; }
    template <typename... Idx, typename = std::enable_if_t<sizeof...(Idx) == D && (std::is_integral_v<Idx> && ...)>>
    const T& operator()(Idx... i) const { return data_[dims_.indexing_safe(i...)]; }
    T& operator()(index_type i) { return data_[dims_.indexing(i)]; }
    const T& operator()(index_type i) const { return data_[dims_.indexing(i)]; }

    const T* data() const { return data_.data(); }
    T* data() { return data_.data(); }
    const_iterator begin() const { return data_.cbegin(); }
    const_iterator end() const { return data_.cend(); }

    // the line along axis `d` through `at` (at[d] is ignored)  Mesh.hpp:342-371
    line_iterator<T> begin(size_type d, index_type at) { return line<T>(data_.data(), d, at, 0); }
    line_iterator<T> end(size_type d, index_type at) { return line<T>(data_.data(), d, at, dims_.dim_size(d)); }
    line_iterator<const T> begin(size_type d, index_type at) const { return line<const T>(data_.data(), d, at, 0); }
    line_iterator<const T> end(size_type d, index_type at) const {
        return line<const T>(data_.data(), d, at, dims_.dim_size(d));
    }

    index_type iter_indices(const_iterator it) const {
        return dims_.dimwise_indices(static_cast<size_type>(std::distance(begin(), it)));
    }
    index_type iter_indices(line_iterator<const T> it) const {
        return dims_.dimwise_indices(static_cast<size_type>(static_cast<const T*>(it) - data()));
    }

   private:
    template <typename V>
    line_iterator<V> line(V* base, size_type d, index_type at, size_type k) const {
        at[d] = 0;
        const auto stride = static_cast<std::ptrdiff_t>(dims_.dim_acc_size(D - 1 - d));
        return line_iterator<V>(base + dims_.indexing_safe(at) + static_cast<std::ptrdiff_t>(k) * stride, stride);
    }

    MeshDimension<D> dims_;
    std::vector<T, Alloc> data_;
};

}  // namespace intp

#endif  // INTP_B200_MESH_HPP
