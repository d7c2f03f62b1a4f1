// ref_cuda_inst.cu — compiles the UNMODIFIED reference CUDA backend /root/reference/src/rakau_cuda.cu (included where it
// lies, never copied) with nvcc for sm_100a and instantiates cuda_acc_pot_impl for the 3-D, 64-bit-code trees that
// ref_capi.cpp (built with -DRAKAU_WITH_CUDA) uses. Test/benchmark infrastructure: bench.py times it as the second
// baseline ("the reference's own CUDA backend on this B200").
#include <rakau_cuda.cu> // found through -I$(REF)/src

namespace rakau
{
inline namespace detail
{
#define RK_REF_INST(Q, F, MAC)                                                                                          \
    template void cuda_acc_pot_impl<Q, 3, F, std::uint64_t, MAC>(                                                      \
        const std::array<F *, tree_nvecs_res<Q, 3>> &, const std::vector<tree_size_t<F>> &,                            \
        const tree_node_t<3, F, std::uint64_t, MAC> *, tree_size_t<F>, const std::array<const F *, 4> &,               \
        const std::uint64_t *, tree_size_t<F>, F, F, F, bool);
RK_REF_INST(0, float, mac::bh)
RK_REF_INST(1, float, mac::bh)
RK_REF_INST(2, float, mac::bh)
RK_REF_INST(0, float, mac::bh_geom)
RK_REF_INST(1, float, mac::bh_geom)
RK_REF_INST(2, float, mac::bh_geom)
RK_REF_INST(0, double, mac::bh)
RK_REF_INST(1, double, mac::bh)
RK_REF_INST(2, double, mac::bh)
RK_REF_INST(0, double, mac::bh_geom)
RK_REF_INST(1, double, mac::bh_geom)
RK_REF_INST(2, double, mac::bh_geom)
#undef RK_REF_INST
} // namespace detail
} // namespace rakau
