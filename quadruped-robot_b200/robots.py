"""Robot constants consumed by the MPC path (benchmark inputs).

Values are the ones the reference reads from its yaml files (paths relative to
/root/reference/quadruped/config):
  A1      a1_sim/a1_sim.yaml:4-5,31,40-43        a1_sim/stance_leg_controller.yaml:40
  Lite3   lite3_sim/robot.yaml:5,9,48,69-72      lite3_sim/stance_leg_controller.yaml:40
  Aliengo aliengo_sim/aliengo_sim.yaml:4-5,36-39 aliengo/stance_leg_controller.yaml:36
Common literals: alpha = 4e-6 and mu = 0.45 (qr_mpc_stance_leg_controller.cpp:83-90),
f_max = total_mass * 9.81 (:80).
"""
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np


@dataclass(frozen=True)
class RobotMPC:
    name: str
    mass: float
    inertia: Tuple[float, float, float]
    weights: Tuple[float, ...]
    com_offset: Tuple[float, float, float]
    hip_positions: Tuple[Tuple[float, float, float], ...]
    body_height: float
    mu: float = 0.45
    alpha: float = 4e-6
    # whole-body model geometry (config/<robot>/*.yaml: body_size, hip_l, upper_l, lower_l); the link
    # masses / inertias are the A1 numbers hard-coded in BuildDynamicModel for every robot
    body_size: Tuple[float, float, float] = (0.267, 0.194, 0.114)
    hip_len: float = 0.08505
    upper_len: float = 0.2
    lower_len: float = 0.2

    @property
    def f_max(self) -> float:
        return self.mass * 9.81


A1 = RobotMPC(
    name="a1", mass=13.0, inertia=(0.24, 0.80, 1.0),
    weights=(10, 10, 5, 40, 60, 100, 0.0, 0, 0.5, 5, 5, 1),
    com_offset=(-0.008, 0.005, 0.0),
    hip_positions=((0.185, -0.135, 0), (0.185, 0.135, 0), (-0.185, -0.135, 0), (-0.185, 0.135, 0)),
    body_height=0.27)

LITE3 = RobotMPC(
    name="lite3", mass=8.742, inertia=(0.24, 0.8, 1.0),
    weights=(20, 20, 10, 40, 40, 150, 0.5, 1, 1, 5, 5, 10),
    com_offset=(-0.012, -0.000, 0.0),
    hip_positions=((0.1745, -0.15, 0), (0.1745, 0.15, 0), (-0.1745, -0.15, 0), (-0.1745, 0.15, 0)),
    body_height=0.27, body_size=(0.349, 0.124, 0.15), hip_len=0.0985, upper_len=0.20, lower_len=0.20)

ALIENGO = RobotMPC(
    name="aliengo", mass=20.0, inertia=(0.24, 0.80, 1.0),
    weights=(10, 10, 5, 40, 60, 100, 0.0, 0, 0.5, 5, 5, 0.1),
    com_offset=(0.0, 0.0, 0.0),
    hip_positions=((0.24, -0.135, 0), (0.24, 0.135, 0), (-0.25, -0.135, 0), (-0.25, 0.135, 0)),
    body_height=0.37)

ROBOTS = {"a1": A1, "lite3": LITE3, "aliengo": ALIENGO}

# Gait tables (config/*/openloop_gait_generator.yaml): per-leg initial phase in the full cycle and
# duty factor; "gallop" is not in the reference and is synthesised through the
# same generic parameters.
GAITS = {
    "trot": dict(offsets=(0.5, 0.0, 0.0, 0.5), duty=0.6, stance_duration=0.5),
    "walk": dict(offsets=(0.5, 0.0, 0.75, 0.25), duty=0.75, stance_duration=7.5),
    "gallop": dict(offsets=(0.0, 0.1, 0.5, 0.6), duty=0.4, stance_duration=0.2),
    "stand": dict(offsets=(0.0, 0.0, 0.0, 0.0), duty=1.0, stance_duration=0.3),
}
