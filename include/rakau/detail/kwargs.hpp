// kwargs.hpp — minimal named-argument machinery for the rakau::tree API of this repository.
//
// The reference takes keyword arguments (x_coords = ..., box_size = ..., G = ..., tree.hpp:599-626) through the
// third-party header detail/igor.hpp. This is an independent implementation of the small surface the tree
// needs: `name = value` produces a tagged reference wrapper; a parser over the call's parameter pack answers
// has(name) / operator()(name) / has_duplicates() / has_unnamed_arguments() at compile time.
#ifndef RAKAU_B200_DETAIL_KWARGS_HPP
#define RAKAU_B200_DETAIL_KWARGS_HPP

#include <cstddef>
#include <initializer_list>
#include <tuple>
#include <type_traits>
#include <utility>

namespace rakau
{
namespace kw
{

// A value bound to a keyword. Holds a forwarding reference, so `masses = std::move(v)` keeps its rvalue-ness.
template <typename Tag, typename T>
struct bound_arg {
    using tag_type = Tag;
    using value_type = T &&;
    T &&value;
};

template <typename Tag>
struct keyword {
    using tag_type = Tag;
    constexpr keyword() = default;
    keyword(const keyword &) = delete;
    keyword &operator=(const keyword &) = delete;
    template <typename T>
    constexpr bound_arg<Tag, T> operator=(T &&v) const
    {
        return bound_arg<Tag, T>{std::forward<T>(v)};
    }
    // `name = {1, 2, 3}`: a braced list cannot bind to a forwarding reference, so initializer lists get their
    // own overloads (the list object lives in the caller until the end of the full expression).
    template <typename T>
    constexpr bound_arg<Tag, std::initializer_list<T>> operator=(std::initializer_list<T> &&l) const
    {
        return bound_arg<Tag, std::initializer_list<T>>{std::move(l)};
    }
    template <typename T>
    constexpr bound_arg<Tag, const std::initializer_list<T> &> operator=(const std::initializer_list<T> &l) const
    {
        return bound_arg<Tag, const std::initializer_list<T> &>{l};
    }
};

template <typename T>
struct is_bound_arg : std::false_type {
};
template <typename Tag, typename T>
struct is_bound_arg<bound_arg<Tag, T>> : std::true_type {
};

template <typename T>
using strip_t = std::remove_cv_t<std::remove_reference_t<T>>;

template <typename Tag, typename Arg, bool = is_bound_arg<strip_t<Arg>>::value>
struct matches : std::false_type {
};
template <typename Tag, typename Arg>
struct matches<Tag, Arg, true> : std::is_same<Tag, typename strip_t<Arg>::tag_type> {
};

// Parser over a pack of (references to) bound arguments.
template <typename... Args>
class parser
{
    std::tuple<Args &...> m_args;

    template <typename Tag, std::size_t I = 0>
    static constexpr std::size_t index_of()
    {
        if constexpr (I == sizeof...(Args)) {
            return I;
        } else if constexpr (matches<Tag, std::tuple_element_t<I, std::tuple<Args...>>>::value) {
            return I;
        } else {
            return index_of<Tag, I + 1>();
        }
    }
    template <typename Tag>
    static constexpr std::size_t count_of()
    {
        return (std::size_t(0) + ... + (matches<Tag, Args>::value ? 1u : 0u));
    }

public:
    constexpr explicit parser(Args &... args) : m_args(args...) {}

    template <typename Tag>
    constexpr bool has(const keyword<Tag> &) const
    {
        return index_of<Tag>() < sizeof...(Args);
    }
    template <typename... Tags>
    constexpr bool has_all(const keyword<Tags> &... kws) const
    {
        return (has(kws) && ...);
    }
    constexpr bool has_unnamed_arguments() const
    {
        return (!is_bound_arg<strip_t<Args>>::value || ...) && sizeof...(Args) > 0;
    }
    constexpr bool has_duplicates() const
    {
        return ((arg_count<Args>() > 1u) || ...);
    }
    // The value bound to the keyword, with the value category it was passed with.
    template <typename Tag>
    constexpr decltype(auto) operator()(const keyword<Tag> &) const
    {
        constexpr std::size_t idx = index_of<Tag>();
        static_assert(idx < sizeof...(Args), "keyword argument not present");
        auto &b = std::get<idx>(m_args);
        using value_type = typename strip_t<decltype(b)>::value_type;
        return static_cast<value_type>(b.value);
    }

private:
    template <typename A>
    static constexpr std::size_t arg_count()
    {
        if constexpr (is_bound_arg<strip_t<A>>::value) {
            return count_of<typename strip_t<A>::tag_type>();
        } else {
            return 0;
        }
    }
};

template <typename... Args>
parser(Args &...) -> parser<Args...>;

} // namespace kw
} // namespace rakau

#endif
