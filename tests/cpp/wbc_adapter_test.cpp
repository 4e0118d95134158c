// wbc_adapter_test.cpp -- drives include/qr_gpu_wbc_adapter.hpp the way qrFSMStateLocomotion::Run drives
// qrWbcLocomotionController (UpdateModel -> Run(ctrlData) on every tick, recomputing on every second one) on robots
// read from a binary dump, and prints torques / joint targets for the Python test to compare.
#include <array>
#include <cstdio>
#include <cstring>

#include "qr_gpu_wbc_adapter.hpp"

template <size_t N>
struct Vec {
    std::array<float, N> a;
    const float* data() const { return a.data(); }
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int B = 0;
    qr_wbc_model model;
    if (std::fread(&B, sizeof(int), 1, f) != 1 || std::fread(&model, sizeof(model), 1, f) != 1) return 2;
    if (qr_gpu_init(0) != QR_OK) { std::printf("init failed: %s\n", qr_gpu_last_error()); return 3; }
    for (int i = 0; i < B; ++i) {
        float state[37], cmd[66];
        int contact[4];
        if (std::fread(state, 4, 37, f) != 37 || std::fread(cmd, 4, 66, f) != 66 || std::fread(contact, 4, 4, f) != 4) return 2;
        Quadruped::gpu::WbcController wbc(model);
        Vec<4> quat; Vec<3> pos, wb, vb; Vec<12> q, qd;
        std::memcpy(quat.a.data(), state, 16); std::memcpy(pos.a.data(), state + 4, 12);
        std::memcpy(wb.a.data(), state + 7, 12); std::memcpy(vb.a.data(), state + 10, 12);
        std::memcpy(q.a.data(), state + 13, 48); std::memcpy(qd.a.data(), state + 25, 48);
        Quadruped::gpu::WbcCtrlData d;
        std::memcpy(d.pBody_des, cmd, 12); std::memcpy(d.vBody_des, cmd + 3, 12); std::memcpy(d.aBody_des, cmd + 6, 12);
        std::memcpy(d.pBody_RPY_des, cmd + 9, 12); std::memcpy(d.vBody_Ori_des, cmd + 12, 12);
        std::memcpy(d.pFoot_des, cmd + 15, 48); std::memcpy(d.vFoot_des, cmd + 27, 48);
        std::memcpy(d.aFoot_des, cmd + 39, 48); std::memcpy(d.Fr_des, cmd + 51, 48);
        for (int l = 0; l < 4; ++l) d.contact_state[l] = contact[l] != 0;
        d.allowAfterMPC = true;
        // first tick: the previous orientation-velocity command of a fresh controller is zero, as in the reference
        wbc.UpdateModel(quat, pos, wb, vb, q, qd);
        const int st0 = wbc.Run(d);
        // second tick must NOT recompute (gating), third must
        float keep[12];
        std::memcpy(keep, wbc.jointTorqueCmd, sizeof(keep));
        const int st1 = wbc.Run(d);
        const bool gated = std::memcmp(keep, wbc.jointTorqueCmd, sizeof(keep)) == 0;
        const int st2 = wbc.Run(d);   // now with prev vBody_Ori_des = d.vBody_Ori_des
        std::printf("W %d %d %d %d %d", i, st0, st1, st2, gated ? 1 : 0);
        for (int k = 0; k < 12; ++k) std::printf(" %.9g", wbc.jointTorqueCmd[k]);
        for (int k = 0; k < 12; ++k) std::printf(" %.9g", wbc.desiredJPos[k]);
        std::printf("\n");
    }
    std::fclose(f);
    return 0;
}
