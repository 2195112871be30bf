// sva_cam.cuh — device copy of the reference's pinhole Camera (include/Camera.h:6-21, src/Camera.cpp:15-33) in strict f64: explicit
// round-to-nearest intrinsics, so no FMA contraction can change a truncation.  Shared by literal mode (k_literal.cu) and the
// point-cloud / depth-map consumers (k_cloud.cu).
#pragma once
#include "sva_common.cuh"

struct DevCam { double px, py, pz, f, ps; };

// Camera::inv_project — src/Camera.cpp:25-33
__device__ __forceinline__ void dev_inv_project(const DevCam& c, int u, int v, double& rx, double& ry, double& rz) {
    double vx = __dmul_rn((double)u, c.ps), vy = __dmul_rn((double)v, c.ps), vz = c.f;
    double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
    rx = __ddiv_rn(vx, n); ry = __ddiv_rn(vy, n); rz = __ddiv_rn(vz, n);
}
// Camera::project — src/Camera.cpp:15-22
__device__ __forceinline__ void dev_project(const DevCam& c, double X, double Y, double Z, int& u, int& v) {
    double mult = __ddiv_rn(__ddiv_rn(c.f, __dsub_rn(Z, c.pz)), c.ps);
    u = (int)__dmul_rn(__dsub_rn(X, c.px), mult);
    v = (int)__dmul_rn(__dsub_rn(Y, c.py), mult);
}

