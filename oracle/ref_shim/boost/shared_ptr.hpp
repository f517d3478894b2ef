// TEST INFRASTRUCTURE (oracle/ref_shim): boost::shared_ptr as the reference's headers use it (reset / -> / *), on std::shared_ptr.
#ifndef LVB_REF_SHIM_BOOST_SHARED_PTR
#define LVB_REF_SHIM_BOOST_SHARED_PTR
#include <memory>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
}
#endif
