// simple_timer.hpp — scoped timers of the drop-in header, active with -DRAKAU_WITH_TIMER like the reference's
// include/rakau/detail/simple_timer.hpp:21-47 (same "Elapsed time for '<phase>': <n>μs" lines, same phase names:
// tree.hpp:934, 1270, 1333, 1440, 1460, 3297, 3534, 3749, 3785). The host-side scopes measure wall clock including the
// host <-> device copies; the phases that run inside one C-ABI call are reported from the CUDA-event times the call
// returns (rk_build_info / rk_eval_info).
#ifndef RAKAU_B200_DETAIL_SIMPLE_TIMER_HPP
#define RAKAU_B200_DETAIL_SIMPLE_TIMER_HPP

#include <chrono>
#include <iostream>
#include <string>

namespace rakau
{
inline namespace detail
{

class simple_timer
{
public:
#if defined(RAKAU_WITH_TIMER)
    explicit simple_timer(const char *desc) : m_desc(desc), m_start(std::chrono::steady_clock::now()) {}
    ~simple_timer()
    {
        report(m_desc.c_str(), std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - m_start).count());
    }
    // A phase measured on the device (CUDA events), in milliseconds.
    static void report_ms(const char *desc, double ms) { report(desc, ms * 1e3); }

private:
    static void report(const char *desc, double us)
    {
        std::cout << "Elapsed time for '" << desc << "': " << static_cast<long long>(us + 0.5) << u8"μs\n";
    }
    const std::string m_desc;
    const std::chrono::steady_clock::time_point m_start;
#else
    explicit simple_timer(const char *) {}
    static void report_ms(const char *, double) {}
#endif
};

} // namespace detail
} // namespace rakau

#endif
