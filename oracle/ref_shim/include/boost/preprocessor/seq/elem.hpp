// Stand-in for <boost/preprocessor/seq/elem.hpp>: src/rakau_cuda.cu uses Boost.Preprocessor only to spell its list of
// explicit instantiations (rakau_cuda.cu:530-568). ref_shim/ref_cuda_inst.cu lists the instantiations it needs by
// hand, so the macro only has to exist.
#ifndef RAKAU_SHIM_BOOST_PP_SEQ_ELEM_HPP
#define RAKAU_SHIM_BOOST_PP_SEQ_ELEM_HPP
#define BOOST_PP_SEQ_ELEM(i, seq)
#endif
