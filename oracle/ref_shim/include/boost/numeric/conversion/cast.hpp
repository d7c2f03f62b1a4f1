// Stand-in for boost::numeric_cast (range-checked arithmetic conversion), written from scratch.
#ifndef RAKAU_SHIM_BOOST_NUMERIC_CAST_HPP
#define RAKAU_SHIM_BOOST_NUMERIC_CAST_HPP
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <typeinfo>
namespace boost
{
namespace numeric
{
struct bad_numeric_cast : std::bad_cast {
    const char *what() const noexcept override { return "bad numeric conversion: overflow"; }
};
} // namespace numeric
template <typename To, typename From>
inline To numeric_cast(const From &x)
{
    if constexpr (std::is_integral_v<To> && std::is_integral_v<From>) {
        using L = std::numeric_limits<To>;
        if constexpr (std::is_signed_v<From> && std::is_unsigned_v<To>) {
            if (x < 0 || static_cast<std::make_unsigned_t<From>>(x) > L::max()) {
                throw numeric::bad_numeric_cast{};
            }
        } else if constexpr (std::is_unsigned_v<From> && std::is_signed_v<To>) {
            if (x > static_cast<std::make_unsigned_t<To>>(L::max())) {
                throw numeric::bad_numeric_cast{};
            }
        } else {
            if (x < L::min() || x > L::max()) {
                throw numeric::bad_numeric_cast{};
            }
        }
        return static_cast<To>(x);
    } else if constexpr (std::is_integral_v<To> && std::is_floating_point_v<From>) {
        if (!(x > static_cast<From>(std::numeric_limits<To>::min()) - From(1))
            || !(x < static_cast<From>(std::numeric_limits<To>::max()) + From(1))) {
            throw numeric::bad_numeric_cast{};
        }
        return static_cast<To>(x);
    } else {
        return static_cast<To>(x);
    }
}
} // namespace boost
#endif
