// mini_test.hpp — a few macros standing in for the parts of Catch the reference's tests use
// (TEST_CASE, REQUIRE, REQUIRE_THROWS_WITH(expr, Contains(...)), REQUIRE_THROWS_AS).
#ifndef RAKAU_B200_MINI_TEST_HPP
#define RAKAU_B200_MINI_TEST_HPP

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <functional>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace mini_test
{
struct registry {
    static std::vector<std::pair<std::string, std::function<void()>>> &cases()
    {
        static std::vector<std::pair<std::string, std::function<void()>>> c;
        return c;
    }
};
struct registrar {
    registrar(const char *name, std::function<void()> f) { registry::cases().emplace_back(name, std::move(f)); }
};
inline int &failures()
{
    static int f = 0;
    return f;
}
inline long &checks()
{
    static long c = 0;
    return c;
}
inline void report(bool ok, const char *expr, const char *file, int line)
{
    ++checks();
    if (!ok) {
        ++failures();
        std::fprintf(stderr, "%s:%d: REQUIRE failed: %s\n", file, line, expr);
        if (failures() > 20) {
            std::fprintf(stderr, "too many failures, aborting\n");
            std::exit(1);
        }
    }
}
inline int run_all()
{
    for (auto &c : registry::cases()) {
        std::printf("[ RUN ] %s\n", c.first.c_str());
        std::fflush(stdout);
        try {
            c.second();
        } catch (const std::exception &e) {
            ++failures();
            std::fprintf(stderr, "unexpected exception in '%s': %s\n", c.first.c_str(), e.what());
        }
    }
    std::printf("%ld checks, %d failures\n", checks(), failures());
    return failures() ? 1 : 0;
}
// tuple_for_each of the reference's test_utils.hpp
template <typename Tuple, typename F>
inline void tuple_for_each(Tuple &&t, F &&f)
{
    std::apply([&f](auto &&... items) { (void(f(std::forward<decltype(items)>(items))), ...); }, std::forward<Tuple>(t));
}
} // namespace mini_test

#define MT_CAT2(a, b) a##b
#define MT_CAT(a, b) MT_CAT2(a, b)
#define TEST_CASE(name)                                                                                                \
    static void MT_CAT(mt_case_, __LINE__)();                                                                          \
    static mini_test::registrar MT_CAT(mt_reg_, __LINE__)(name, MT_CAT(mt_case_, __LINE__));                          \
    static void MT_CAT(mt_case_, __LINE__)()
#define REQUIRE(...) mini_test::report(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__)
#define REQUIRE_THROWS_WITH(expr, substr)                                                                              \
    do {                                                                                                               \
        bool mt_ok = false;                                                                                            \
        std::string mt_msg = "(no exception)";                                                                         \
        try {                                                                                                          \
            (void)(expr);                                                                                              \
        } catch (const std::exception &e) {                                                                            \
            mt_msg = e.what();                                                                                         \
            mt_ok = mt_msg.find(substr) != std::string::npos;                                                          \
        }                                                                                                              \
        if (!mt_ok) std::fprintf(stderr, "  got: %s\n", mt_msg.c_str());                                               \
        mini_test::report(mt_ok, "throws with: " #substr, __FILE__, __LINE__);                                         \
    } while (0)
#define REQUIRE_THROWS_AS(expr, type)                                                                                  \
    do {                                                                                                               \
        bool mt_ok = false;                                                                                            \
        try {                                                                                                          \
            (void)(expr);                                                                                              \
        } catch (const type &) {                                                                                       \
            mt_ok = true;                                                                                              \
        } catch (...) {                                                                                                \
        }                                                                                                              \
        mini_test::report(mt_ok, "throws " #type, __FILE__, __LINE__);                                                 \
    } while (0)
#define MINI_TEST_MAIN()                                                                                               \
    int main() { return mini_test::run_all(); }

#endif
