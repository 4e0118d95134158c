// Stand-in for the reference's fsm/qr_control_fsm_data.hpp (robot, estimator, gait and ROS
// containers).  qr_multitask_projection.hpp and qr_wholebody_impulse_ctrl.hpp include it but the WBC
// translation units compiled by oracle/Makefile (`refwbc`) name nothing it declares except the two
// constants NumMotor / BaseFreedomDim, which reach them through it from the reference's own
// config/qr_config.h -- included here as it lies.
// TEST INFRASTRUCTURE ONLY.
#ifndef MINI_QR_CONTROL_FSM_DATA_HPP
#define MINI_QR_CONTROL_FSM_DATA_HPP
#include "config/qr_config.h"
#endif
