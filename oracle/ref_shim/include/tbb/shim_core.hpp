// Minimal, from-scratch stand-ins for the part of Intel TBB that rakau/tree.hpp uses. They exist only so that
// the UNMODIFIED reference header compiles offline (no TBB in this image); scheduling is done with OpenMP.
// Not a TBB implementation: just blocked_range, parallel_for/reduce/sort/invoke, task_group,
// concurrent_vector and the partitioner tags, with the semantics the reference relies on.
#ifndef RAKAU_SHIM_TBB_CORE_HPP
#define RAKAU_SHIM_TBB_CORE_HPP

#include <algorithm>
#include <cstddef>
#include <deque>
#include <iterator>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include <omp.h>
#if defined(RAKAU_SHIM_UNSTABLE_SORT)
#include <parallel/algorithm>
#endif

namespace tbb
{

template <typename Value>
class blocked_range
{
public:
    using const_iterator = Value;
    using size_type = std::size_t;
    blocked_range(Value b, Value e, size_type grain = 1) : m_b(b), m_e(e), m_grain(grain ? grain : 1) {}
    Value begin() const { return m_b; }
    Value end() const { return m_e; }
    size_type size() const { return static_cast<size_type>(m_e - m_b); }
    size_type grainsize() const { return m_grain; }
    bool empty() const { return !(m_b < m_e); }

private:
    Value m_b, m_e;
    size_type m_grain;
};

struct simple_partitioner {
};
struct auto_partitioner {
};

namespace shim_detail
{
// Number of chunks a range is cut into: enough for dynamic load balancing, never below the grain size.
template <typename Range>
inline std::size_t n_chunks(const Range &r)
{
    const std::size_t n = r.size();
    if (!n) {
        return 0;
    }
    const std::size_t by_grain = (n + r.grainsize() - 1) / r.grainsize();
    const std::size_t want = static_cast<std::size_t>(omp_get_max_threads()) * 16u;
    return std::max<std::size_t>(1, std::min(by_grain, want));
}
template <typename Range>
inline Range sub_range(const Range &r, std::size_t c, std::size_t nc)
{
    const std::size_t n = r.size();
    const auto b = r.begin() + static_cast<std::ptrdiff_t>(n * c / nc);
    const auto e = r.begin() + static_cast<std::ptrdiff_t>(n * (c + 1) / nc);
    return Range(b, e, r.grainsize());
}
} // namespace shim_detail

template <typename Range, typename Body>
inline void parallel_for(const Range &r, const Body &body)
{
    const std::size_t nc = shim_detail::n_chunks(r);
    if (nc <= 1 || omp_in_parallel()) {
        if (nc) {
            body(r);
        }
        return;
    }
    std::exception_ptr err;
    std::mutex mu;
#pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t c = 0; c < nc; ++c) {
        try {
            body(shim_detail::sub_range(r, c, nc));
        } catch (...) {
            std::lock_guard<std::mutex> lock(mu);
            if (!err) {
                err = std::current_exception();
            }
        }
    }
    if (err) {
        std::rethrow_exception(err);
    }
}
template <typename Range, typename Body, typename Partitioner>
inline void parallel_for(const Range &r, const Body &body, const Partitioner &)
{
    parallel_for(r, body);
}

template <typename Range, typename Value, typename Body, typename Reduction>
inline Value parallel_reduce(const Range &r, const Value &identity, const Body &body, const Reduction &red)
{
    const std::size_t nc = shim_detail::n_chunks(r);
    if (nc <= 1 || omp_in_parallel()) {
        return nc ? body(r, identity) : identity;
    }
    std::vector<Value> partial(nc, identity);
    std::exception_ptr err;
    std::mutex mu;
#pragma omp parallel for schedule(dynamic, 1)
    for (std::size_t c = 0; c < nc; ++c) {
        try {
            partial[c] = body(shim_detail::sub_range(r, c, nc), identity);
        } catch (...) {
            std::lock_guard<std::mutex> lock(mu);
            if (!err) {
                err = std::current_exception();
            }
        }
    }
    if (err) {
        std::rethrow_exception(err);
    }
    Value out = identity;
    for (const auto &p : partial) {
        out = red(out, p);
    }
    return out;
}

// tbb::parallel_sort is unstable; a stable sort is one of its legal outcomes and makes the shimmed reference
// deterministic (the oracle's canonical order). RAKAU_SHIM_UNSTABLE_SORT selects a parallel unstable sort.
template <typename It, typename Compare>
inline void parallel_sort(It b, It e, const Compare &cmp)
{
#if defined(RAKAU_SHIM_UNSTABLE_SORT)
    __gnu_parallel::sort(b, e, cmp); // libstdc++ parallel mode (OpenMP): a parallel unstable sort like TBB's
#else
    std::stable_sort(b, e, cmp);
#endif
}
template <typename It>
inline void parallel_sort(It b, It e)
{
    parallel_sort(b, e, std::less<typename std::iterator_traits<It>::value_type>{});
}

template <typename... Fs>
inline void parallel_invoke(const Fs &... fs)
{
    (fs(), ...);
}

// Tasks run inline: the reference only needs run()/wait() to be correct, not concurrent.
class task_group
{
public:
    template <typename F>
    void run(const F &f)
    {
        f();
    }
    void wait() {}
};

// push_back must be safe from several threads and must not invalidate references (the reference keeps
// `auto &new_tree = *trees.push_back(...)` alive while other tasks push): std::deque + mutex.
template <typename T>
class concurrent_vector
{
    std::deque<T> m_d;
    mutable std::mutex m_mu;

public:
    using iterator = typename std::deque<T>::iterator;
    using const_iterator = typename std::deque<T>::const_iterator;
    using size_type = std::size_t;
    using value_type = T;
    concurrent_vector() = default;
    concurrent_vector(const concurrent_vector &o) : m_d(o.m_d) {}
    concurrent_vector &operator=(const concurrent_vector &o)
    {
        m_d = o.m_d;
        return *this;
    }
    iterator push_back(const T &v)
    {
        std::lock_guard<std::mutex> lock(m_mu);
        m_d.push_back(v);
        return std::prev(m_d.end());
    }
    iterator push_back(T &&v)
    {
        std::lock_guard<std::mutex> lock(m_mu);
        m_d.push_back(std::move(v));
        return std::prev(m_d.end());
    }
    template <typename... Args>
    iterator emplace_back(Args &&... a)
    {
        std::lock_guard<std::mutex> lock(m_mu);
        m_d.emplace_back(std::forward<Args>(a)...);
        return std::prev(m_d.end());
    }
    iterator begin() { return m_d.begin(); }
    iterator end() { return m_d.end(); }
    const_iterator begin() const { return m_d.begin(); }
    const_iterator end() const { return m_d.end(); }
    size_type size() const { return m_d.size(); }
    bool empty() const { return m_d.empty(); }
    void clear() { m_d.clear(); }
    T &operator[](size_type i) { return m_d[i]; }
    const T &operator[](size_type i) const { return m_d[i]; }
    friend bool operator==(const concurrent_vector &a, const concurrent_vector &b) { return a.m_d == b.m_d; }
};

} // namespace tbb

#endif
