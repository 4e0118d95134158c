// Stand-in for <ros/ros.h> (ROS is not installed): the reference's robots/qr_timer.h and
// utils/qr_tools.h only need ros::Time::now().toSec() to parse.  TEST INFRASTRUCTURE ONLY.
#ifndef MINI_ROS_H
#define MINI_ROS_H
#include <time.h>
namespace ros {
struct Time {
    double sec = 0;
    static Time now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        Time t;
        t.sec = ts.tv_sec + 1e-9 * ts.tv_nsec;
        return t;
    }
    double toSec() const { return sec; }
};
}   // namespace ros
#endif
