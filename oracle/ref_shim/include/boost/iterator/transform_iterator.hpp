// Stand-in for boost::transform_iterator / make_transform_iterator (random access, with base()).
#ifndef RAKAU_SHIM_BOOST_TRANSFORM_ITERATOR_HPP
#define RAKAU_SHIM_BOOST_TRANSFORM_ITERATOR_HPP
#include <iterator>
#include <type_traits>
namespace boost
{
template <typename F, typename It>
class transform_iterator
{
    It m_it{};
    F m_f;

public:
    using iterator_category = std::random_access_iterator_tag;
    using difference_type = typename std::iterator_traits<It>::difference_type;
    using reference = decltype(std::declval<const F &>()(*std::declval<It>()));
    using value_type = std::remove_cv_t<std::remove_reference_t<reference>>;
    using pointer = void;
    transform_iterator(It it, F f) : m_it(it), m_f(f) {}
    const It &base() const { return m_it; }
    reference operator*() const { return m_f(*m_it); }
    reference operator[](difference_type n) const { return m_f(*(m_it + n)); }
    transform_iterator &operator++() { ++m_it; return *this; }
    transform_iterator operator++(int) { auto t = *this; ++m_it; return t; }
    transform_iterator &operator--() { --m_it; return *this; }
    transform_iterator operator--(int) { auto t = *this; --m_it; return t; }
    transform_iterator &operator+=(difference_type n) { m_it += n; return *this; }
    transform_iterator &operator-=(difference_type n) { m_it -= n; return *this; }
    friend transform_iterator operator+(transform_iterator a, difference_type n) { return a += n; }
    friend transform_iterator operator+(difference_type n, transform_iterator a) { return a += n; }
    friend transform_iterator operator-(transform_iterator a, difference_type n) { return a -= n; }
    friend difference_type operator-(const transform_iterator &a, const transform_iterator &b) { return a.m_it - b.m_it; }
    friend bool operator==(const transform_iterator &a, const transform_iterator &b) { return a.m_it == b.m_it; }
    friend bool operator!=(const transform_iterator &a, const transform_iterator &b) { return a.m_it != b.m_it; }
    friend bool operator<(const transform_iterator &a, const transform_iterator &b) { return a.m_it < b.m_it; }
    friend bool operator>(const transform_iterator &a, const transform_iterator &b) { return a.m_it > b.m_it; }
    friend bool operator<=(const transform_iterator &a, const transform_iterator &b) { return a.m_it <= b.m_it; }
    friend bool operator>=(const transform_iterator &a, const transform_iterator &b) { return a.m_it >= b.m_it; }
};
template <typename F, typename It>
inline transform_iterator<F, It> make_transform_iterator(It it, F f)
{
    return transform_iterator<F, It>(it, f);
}
} // namespace boost
#endif
