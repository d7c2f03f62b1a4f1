// Stand-in for <boost/preprocessor/seq/for_each_product.hpp> (see elem.hpp): the cartesian-product instantiation list
// of src/rakau_cuda.cu:566 expands to nothing; the trailing ';' of that line needs a declaration to close.
#ifndef RAKAU_SHIM_BOOST_PP_SEQ_FOR_EACH_PRODUCT_HPP
#define RAKAU_SHIM_BOOST_PP_SEQ_FOR_EACH_PRODUCT_HPP
#define BOOST_PP_SEQ_FOR_EACH_PRODUCT(macro, seqs) static_assert(true, "instantiations are listed in ref_cuda_inst.cu")
#endif
