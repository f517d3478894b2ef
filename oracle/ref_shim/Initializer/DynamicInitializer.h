// TEST INFRASTRUCTURE (oracle/ref_shim): stands in front of the reference's include/Initializer/DynamicInitializer.h (the
// include path lists this directory first).  The dynamic initialiser needs Ceres + OpenCV (out of scope);
// LarVio only reaches it through FlexibleInitializer (include/Initializer/FlexibleInitializer.h:38-41, src/FlexibleInitializer.cpp:
// 19-23), so a class with the same constructor that never initialises leaves the static initialiser and the filter as they are.
#ifndef DYNAMIC_INITIALIZER_H
#define DYNAMIC_INITIALIZER_H
#include <vector>
#include <Eigen/Dense>
#include <larvio/feature_msg.h>
#include "larvio/imu_state.h"
#include "sensors/ImuData.hpp"
using namespace std;
namespace larvio {
class DynamicInitializer {
 public:
  DynamicInitializer() = delete;
  DynamicInitializer(const double&, const Eigen::Matrix3d&, const Eigen::Matrix3d&, const Eigen::Matrix3d&, const double&, const double&,
                     const double&, const double&, const Eigen::Matrix3d&, const Eigen::Vector3d&, const double&) {}
  bool tryDynInit(const std::vector<ImuData>&, MonoCameraMeasurementPtr) { return false; }
  void assignInitialState(std::vector<ImuData>&, Eigen::Vector3d&, Eigen::Vector3d&, IMUState&) {}
  bool ifInitialized() { return false; }
};
}  // namespace larvio
#endif
