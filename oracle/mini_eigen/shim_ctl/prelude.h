// Forced include of the `refctl` build (oracle/Makefile) -- TEST INFRASTRUCTURE ONLY: names the reference sources use
// without including their headers (they reach them through Eigen / ROS headers in a real build).
#ifndef MINI_CTL_PRELUDE_H
#define MINI_CTL_PRELUDE_H
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
using std::isnan;
#endif
