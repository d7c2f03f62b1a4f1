// aligned_allocator.hpp — allocator behind rakau::f_vector<F> (reference: detail/di_aligned_allocator.hpp,
// tree.hpp:628-631): over-aligned storage, and default-initialisation (not value-initialisation) of trivially
// constructible elements so that resizing a large vector does not memset it.
#ifndef RAKAU_B200_DETAIL_ALIGNED_ALLOCATOR_HPP
#define RAKAU_B200_DETAIL_ALIGNED_ALLOCATOR_HPP

#include <cstddef>
#include <cstdlib>
#include <limits>
#include <new>
#include <type_traits>
#include <utility>

namespace rakau
{
inline namespace detail
{

template <typename T, std::size_t Alignment = 0>
struct di_aligned_allocator {
    static_assert(Alignment == 0 || (Alignment & (Alignment - 1)) == 0, "alignment must be a power of two");
    using value_type = T;
    using size_type = std::size_t;
    using difference_type = std::ptrdiff_t;
    using propagate_on_container_move_assignment = std::true_type;
    using is_always_equal = std::true_type;
    template <typename U>
    struct rebind {
        using other = di_aligned_allocator<U, Alignment>;
    };

    di_aligned_allocator() noexcept = default;
    template <typename U>
    di_aligned_allocator(const di_aligned_allocator<U, Alignment> &) noexcept
    {
    }

    T *allocate(size_type n) const
    {
        if (n > std::numeric_limits<size_type>::max() / sizeof(T)) {
            throw std::bad_alloc{};
        }
        if (n == 0) {
            return nullptr;
        }
        constexpr std::size_t al = Alignment > alignof(T) ? Alignment : alignof(T);
        void *p = nullptr;
        if constexpr (al <= alignof(std::max_align_t)) {
            p = std::malloc(n * sizeof(T));
        } else {
            // aligned_alloc needs the size to be a multiple of the alignment
            const std::size_t bytes = (n * sizeof(T) + al - 1) / al * al;
            p = std::aligned_alloc(al, bytes);
        }
        if (!p) {
            throw std::bad_alloc{};
        }
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_type) const noexcept { std::free(p); }

    // default-init when no arguments are given, regular construction otherwise
    template <typename U>
    void construct(U *p) const noexcept(std::is_nothrow_default_constructible_v<U>)
    {
        ::new (static_cast<void *>(p)) U;
    }
    template <typename U, typename... Args>
    void construct(U *p, Args &&... args) const
    {
        ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }

    friend bool operator==(const di_aligned_allocator &, const di_aligned_allocator &) noexcept { return true; }
    friend bool operator!=(const di_aligned_allocator &, const di_aligned_allocator &) noexcept { return false; }
};

} // namespace detail
} // namespace rakau

#endif
