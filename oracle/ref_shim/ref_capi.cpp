// ref_capi.cpp — C wrapper around the UNMODIFIED reference header /root/reference/include/rakau/tree.hpp,
// compiled against the dependency stand-ins in ref_shim/include (Boost, TBB, xsimd are not installed in this
// image). Output: oracle/_ref/libref_*.so — test infrastructure: it pins oracle/rakau_oracle.cpp against the
// reference's own code and serves as the "reference" CPU baseline of bench.py. Label wherever reported:
// "reference source, shimmed deps".
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

#include <omp.h>

#include <rakau/tree.hpp>

using namespace rakau;
using namespace rakau::kwargs;

namespace
{
struct iface {
    virtual ~iface() = default;
    virtual std::size_t nparts() const = 0;
    virtual std::size_t nnodes() const = 0;
    virtual double box_size() const = 0;
    virtual void codes(std::uint64_t *) const = 0;
    virtual void perm(int which, std::uint64_t *) const = 0;
    virtual void parts(void *x, void *y, void *z, void *m) const = 0;
    virtual void nodes(void *out) const = 0; // {u64 begin,end,n_children,code,level; F props[4]; F dim; F delta}
    virtual void acc_pot(int Q, double theta, double G, double eps, void *o0, void *o1, void *o2, void *o3,
                         const std::vector<double> &split) const = 0;
    virtual void exact(std::size_t idx, double G, double eps, double *out4) const = 0;
    virtual void update_positions(const void *x, const void *y, const void *z) = 0;
    virtual void update_masses(const void *m) = 0;
};

template <typename F, mac MAC>
struct impl final : iface {
    octree<F, MAC> t;
    impl(const F *x, const F *y, const F *z, const F *m, std::size_t n, double box, bool deduce, std::size_t mln,
         std::size_t nc)
        : t(deduce ? octree<F, MAC>{x_coords = x, y_coords = y, z_coords = z, masses = m, kwargs::nparts = n,
                                    max_leaf_n = mln, ncrit = nc}
                   : octree<F, MAC>{x_coords = x, y_coords = y, z_coords = z, masses = m, kwargs::nparts = n,
                                    kwargs::box_size = static_cast<F>(box), max_leaf_n = mln, ncrit = nc})
    {
    }
    std::size_t nparts() const override { return t.nparts(); }
    std::size_t nnodes() const override { return t.nodes().size(); }
    double box_size() const override { return t.box_size(); }
    void codes(std::uint64_t *o) const override { std::copy(t.c_it_u(), t.c_it_u() + t.nparts(), o); }
    void perm(int which, std::uint64_t *o) const override
    {
        const auto &v = which == 0 ? t.perm() : (which == 1 ? t.last_perm() : t.inv_perm());
        std::copy(v.begin(), v.end(), o);
    }
    void parts(void *x, void *y, void *z, void *m) const override
    {
        void *o[4] = {x, y, z, m};
        const auto p = t.p_its_u();
        for (int j = 0; j < 4; ++j) {
            std::copy(p[j], p[j] + t.nparts(), static_cast<F *>(o[j]));
        }
    }
    void nodes(void *out) const override
    {
        struct rec {
            std::uint64_t begin, end, n_children, code, level;
            F props[4], dim, delta;
        };
        auto *r = static_cast<rec *>(out);
        std::size_t i = 0;
        for (const auto &n : t.nodes()) {
            r[i].begin = n.begin;
            r[i].end = n.end;
            r[i].n_children = n.n_children;
            r[i].code = n.code;
            r[i].level = n.level;
            for (int j = 0; j < 4; ++j) {
                r[i].props[j] = n.props[j];
            }
            if constexpr (MAC == mac::bh) {
                r[i].dim = n.dim2;
                r[i].delta = 0;
            } else {
                r[i].dim = n.dim;
                r[i].delta = n.delta;
            }
            ++i;
        }
    }
    void acc_pot(int Q, double theta, double Gc, double e, void *o0, void *o1, void *o2, void *o3,
                 const std::vector<double> &sp) const override
    {
        // (split = {}: the reference's default, everything on the CPU)
        if (Q == 0) {
            t.accs_u(std::array<F *, 3>{static_cast<F *>(o0), static_cast<F *>(o1), static_cast<F *>(o2)}, F(theta),
                     G = F(Gc), eps = F(e), split = sp);
        } else if (Q == 1) {
            t.pots_u(static_cast<F *>(o0), F(theta), G = F(Gc), eps = F(e), split = sp);
        } else {
            t.accs_pots_u(std::array<F *, 4>{static_cast<F *>(o0), static_cast<F *>(o1), static_cast<F *>(o2),
                                             static_cast<F *>(o3)},
                          F(theta), G = F(Gc), eps = F(e), split = sp);
        }
    }
    void exact(std::size_t idx, double Gc, double e, double *out4) const override
    {
        const auto r = t.exact_acc_pot_u(idx, G = F(Gc), eps = F(e));
        for (int j = 0; j < 4; ++j) {
            out4[j] = r[j];
        }
    }
    void update_positions(const void *x, const void *y, const void *z) override
    {
        const std::size_t n = t.nparts();
        const F *in[3] = {static_cast<const F *>(x), static_cast<const F *>(y), static_cast<const F *>(z)};
        t.update_particles_u([&](const auto &its) {
            for (int j = 0; j < 3; ++j) {
                if (in[j]) {
                    std::copy(in[j], in[j] + n, its[j]);
                }
            }
        });
    }
    void update_masses(const void *m) override
    {
        const std::size_t n = t.nparts();
        t.update_masses_u([&](auto it) { std::copy(static_cast<const F *>(m), static_cast<const F *>(m) + n, it); });
    }
};

struct handle {
    iface *p = nullptr;
    std::string err;
    int fp = 32;
};
} // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API const char *ref_variant()
{
#if defined(RAKAU_SHIM_BRIDGE)
    return "reference source, shimmed deps, RAKAU_WITH_CUDA: cuda_acc_pot_impl = integration/rakau_b200_bridge.cpp";
#elif defined(RAKAU_WITH_CUDA)
    return "reference source, shimmed deps, RAKAU_WITH_CUDA: the reference's own src/rakau_cuda.cu (nvcc, sm_100a)";
#elif defined(RAKAU_DISABLE_SIMD)
    return "reference source, shimmed deps, scalar (RAKAU_DISABLE_SIMD), stable sort";
#elif defined(__AVX512F__)
    return "reference source, shimmed deps (OpenMP-backed TBB stand-in), AVX-512 batches + rsqrt14";
#else
    return "reference source, shimmed deps (OpenMP-backed TBB stand-in), AVX2 batches + rsqrt";
#endif
}
REF_API void ref_set_threads(int n)
{
    omp_set_num_threads(n > 0 ? n : 1);
}
REF_API void *ref_create(int fp, int mac_kind, const void *x, const void *y, const void *z, const void *m,
                         std::size_t n, double box, int deduce, std::size_t mln, std::size_t nc, char *err,
                         std::size_t errlen)
{
    auto *h = new handle;
    h->fp = fp;
    try {
        if (fp == 32 && mac_kind == 0) {
            h->p = new impl<float, mac::bh>(static_cast<const float *>(x), static_cast<const float *>(y),
                                            static_cast<const float *>(z), static_cast<const float *>(m), n, box,
                                            deduce != 0, mln, nc);
        } else if (fp == 32) {
            h->p = new impl<float, mac::bh_geom>(static_cast<const float *>(x), static_cast<const float *>(y),
                                                 static_cast<const float *>(z), static_cast<const float *>(m), n, box,
                                                 deduce != 0, mln, nc);
        } else if (mac_kind == 0) {
            h->p = new impl<double, mac::bh>(static_cast<const double *>(x), static_cast<const double *>(y),
                                             static_cast<const double *>(z), static_cast<const double *>(m), n, box,
                                             deduce != 0, mln, nc);
        } else {
            h->p = new impl<double, mac::bh_geom>(static_cast<const double *>(x), static_cast<const double *>(y),
                                                  static_cast<const double *>(z), static_cast<const double *>(m), n,
                                                  box, deduce != 0, mln, nc);
        }
        return h;
    } catch (const std::exception &e) {
        if (err && errlen) {
            std::strncpy(err, e.what(), errlen - 1);
            err[errlen - 1] = 0;
        }
        delete h;
        return nullptr;
    }
}
REF_API void ref_destroy(void *p)
{
    auto *h = static_cast<handle *>(p);
    if (h) {
        delete h->p;
        delete h;
    }
}
REF_API const char *ref_last_error(void *p) { return static_cast<handle *>(p)->err.c_str(); }
REF_API std::size_t ref_nparts(void *p) { return static_cast<handle *>(p)->p->nparts(); }
REF_API std::size_t ref_nnodes(void *p) { return static_cast<handle *>(p)->p->nnodes(); }
REF_API double ref_box_size(void *p) { return static_cast<handle *>(p)->p->box_size(); }
REF_API void ref_get_codes(void *p, std::uint64_t *o) { static_cast<handle *>(p)->p->codes(o); }
REF_API void ref_get_perm(void *p, int which, std::uint64_t *o) { static_cast<handle *>(p)->p->perm(which, o); }
REF_API void ref_get_parts(void *p, void *x, void *y, void *z, void *m) { static_cast<handle *>(p)->p->parts(x, y, z, m); }
REF_API void ref_get_nodes(void *p, void *o) { static_cast<handle *>(p)->p->nodes(o); }

template <typename Fn>
static int guarded(void *p, Fn &&fn)
{
    auto *h = static_cast<handle *>(p);
    try {
        fn();
        return 0;
    } catch (const std::invalid_argument &e) {
        h->err = e.what();
        return 1;
    } catch (const std::domain_error &e) {
        h->err = e.what();
        return 2;
    } catch (const std::exception &e) {
        h->err = e.what();
        return 4;
    }
}
REF_API int ref_acc_pot(void *p, int Q, double theta, double G, double eps, void *o0, void *o1, void *o2, void *o3)
{
    return guarded(p, [&]() { static_cast<handle *>(p)->p->acc_pot(Q, theta, G, eps, o0, o1, o2, o3, {}); });
}
// The reference's `split` kwarg (tree.hpp:3147-3198): split[0] = CPU share, split[1..] = accelerator shares.
REF_API int ref_acc_pot_split(void *p, int Q, double theta, double G, double eps, void *o0, void *o1, void *o2,
                              void *o3, const double *split, std::size_t nsplit)
{
    return guarded(p, [&]() {
        static_cast<handle *>(p)->p->acc_pot(Q, theta, G, eps, o0, o1, o2, o3, std::vector<double>(split, split + nsplit));
    });
}
REF_API int ref_exact(void *p, std::size_t idx, double G, double eps, double *out4)
{
    return guarded(p, [&]() { static_cast<handle *>(p)->p->exact(idx, G, eps, out4); });
}
REF_API int ref_update_positions(void *p, const void *x, const void *y, const void *z)
{
    return guarded(p, [&]() { static_cast<handle *>(p)->p->update_positions(x, y, z); });
}
REF_API int ref_update_masses(void *p, const void *m)
{
    return guarded(p, [&]() { static_cast<handle *>(p)->p->update_masses(m); });
}
