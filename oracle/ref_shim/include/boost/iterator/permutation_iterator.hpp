// Stand-in for boost::permutation_iterator / make_permutation_iterator (random access over an index iterator).
#ifndef RAKAU_SHIM_BOOST_PERMUTATION_ITERATOR_HPP
#define RAKAU_SHIM_BOOST_PERMUTATION_ITERATOR_HPP
#include <iterator>
namespace boost
{
template <typename ElemIt, typename IdxIt>
class permutation_iterator
{
    ElemIt m_e{};
    IdxIt m_i{};

public:
    using iterator_category = std::random_access_iterator_tag;
    using difference_type = typename std::iterator_traits<IdxIt>::difference_type;
    using value_type = typename std::iterator_traits<ElemIt>::value_type;
    using reference = typename std::iterator_traits<ElemIt>::reference;
    using pointer = typename std::iterator_traits<ElemIt>::pointer;
    permutation_iterator() = default;
    permutation_iterator(ElemIt e, IdxIt i) : m_e(e), m_i(i) {}
    reference operator*() const { return *(m_e + static_cast<typename std::iterator_traits<ElemIt>::difference_type>(*m_i)); }
    reference operator[](difference_type n) const
    {
        return *(m_e + static_cast<typename std::iterator_traits<ElemIt>::difference_type>(*(m_i + n)));
    }
    permutation_iterator &operator++() { ++m_i; return *this; }
    permutation_iterator operator++(int) { auto t = *this; ++m_i; return t; }
    permutation_iterator &operator--() { --m_i; return *this; }
    permutation_iterator operator--(int) { auto t = *this; --m_i; return t; }
    permutation_iterator &operator+=(difference_type n) { m_i += n; return *this; }
    permutation_iterator &operator-=(difference_type n) { m_i -= n; return *this; }
    friend permutation_iterator operator+(permutation_iterator a, difference_type n) { return a += n; }
    friend permutation_iterator operator+(difference_type n, permutation_iterator a) { return a += n; }
    friend permutation_iterator operator-(permutation_iterator a, difference_type n) { return a -= n; }
    friend difference_type operator-(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i - b.m_i; }
    friend bool operator==(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i == b.m_i; }
    friend bool operator!=(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i != b.m_i; }
    friend bool operator<(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i < b.m_i; }
    friend bool operator>(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i > b.m_i; }
    friend bool operator<=(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i <= b.m_i; }
    friend bool operator>=(const permutation_iterator &a, const permutation_iterator &b) { return a.m_i >= b.m_i; }
};
template <typename ElemIt, typename IdxIt>
inline permutation_iterator<ElemIt, IdxIt> make_permutation_iterator(ElemIt e, IdxIt i)
{
    return permutation_iterator<ElemIt, IdxIt>(e, i);
}
} // namespace boost
#endif
