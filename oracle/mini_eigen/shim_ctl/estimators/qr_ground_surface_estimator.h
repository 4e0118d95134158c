// Stand-in for the reference's estimators/qr_ground_surface_estimator.h -- TEST INFRASTRUCTURE ONLY.
// qr_qp_torque_optimizer.cpp reads three things of the ground estimator; they are plain data here.
#ifndef MINI_QR_GROUND_SURFACE_ESTIMATOR_H
#define MINI_QR_GROUND_SURFACE_ESTIMATOR_H
#include "robots/qr_robot.h"
namespace Quadruped {
struct qrTerrain {
    TerrainType terrainType = TerrainType::PLANE;
};
class qrGroundSurfaceEstimator {
public:
    qrTerrain terrain;
    Vec3<float> controlFrameRPY;
    Mat3<float> alignedDirections;
    Vec3<float> GetControlFrameRPY() const { return controlFrameRPY; }
    Mat3<float> GetAlignedDirections() const { return alignedDirections; }
};
}   // namespace Quadruped
#endif
