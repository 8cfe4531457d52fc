// TEST STUB (see registration.h)
#ifndef S4_TEST_STUB_PCL_TE_SVD_H_
#define S4_TEST_STUB_PCL_TE_SVD_H_
#include <pcl/registration/registration.h>
namespace pcl { namespace registration {
template <typename S, typename T>
struct TransformationEstimationSVD : public TransformationEstimation<S, T> {};
} }
#endif
