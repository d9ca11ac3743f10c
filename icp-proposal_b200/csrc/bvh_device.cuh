// bvh_device.cuh - device helpers shared by the nearest-primitive kernels (bvh.cu, bvh_wide.cu): the exact FP64
// point-triangle closest point, the leaf test with the lowest-index tie rule and the conservative FP32 box distance.
#pragma once

#include "icp_internal.h"

namespace icp {

// ---------------------------------------------------------------------------------------------------
// exact point-triangle closest point (FP64), classified vertex (0) / edge (1) / face (2)
// ---------------------------------------------------------------------------------------------------
struct Hit {
    double d2;
    double x, y, z;
    int prim;
    int feat;
    int slot;
};

__device__ __forceinline__ void point_triangle(double px, double py, double pz, double ax, double ay, double az,
                                               double bx, double by, double bz, double cx, double cy, double cz,
                                               double &rx, double &ry, double &rz, int &feat) {
    double abx = bx - ax, aby = by - ay, abz = bz - az;
    double acx = cx - ax, acy = cy - ay, acz = cz - az;
    double apx = px - ax, apy = py - ay, apz = pz - az;
    double d1 = abx * apx + aby * apy + abz * apz, d2 = acx * apx + acy * apy + acz * apz;
    if (d1 <= 0.0 && d2 <= 0.0) { rx = ax; ry = ay; rz = az; feat = 0; return; }
    double bpx = px - bx, bpy = py - by, bpz = pz - bz;
    double d3 = abx * bpx + aby * bpy + abz * bpz, d4 = acx * bpx + acy * bpy + acz * bpz;
    if (d3 >= 0.0 && d4 <= d3) { rx = bx; ry = by; rz = bz; feat = 0; return; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
        double v = d1 / (d1 - d3);
        rx = ax + v * abx; ry = ay + v * aby; rz = az + v * abz; feat = 1; return;
    }
    double cpx = px - cx, cpy = py - cy, cpz = pz - cz;
    double d5 = abx * cpx + aby * cpy + abz * cpz, d6 = acx * cpx + acy * cpy + acz * cpz;
    if (d6 >= 0.0 && d5 <= d6) { rx = cx; ry = cy; rz = cz; feat = 0; return; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        double w = d2 / (d2 - d6);
        rx = ax + w * acx; ry = ay + w * acy; rz = az + w * acz; feat = 1; return;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        rx = bx + w * (cx - bx); ry = by + w * (cy - by); rz = bz + w * (cz - bz); feat = 1; return;
    }
    double denom = 1.0 / (va + vb + vc);
    double v = vb * denom, w = vc * denom;
    rx = ax + abx * v + acx * w; ry = ay + aby * v + acy * w; rz = az + abz * v + acz * w; feat = 2;
}

template <int PRIM, bool DYNAMIC>
__device__ __forceinline__ void leaf_test(int slot, const int *__restrict__ prim, const double *__restrict__ prim_data,
                                          const double *__restrict__ Xi, const int *__restrict__ tris, double qx,
                                          double qy, double qz, Hit &h) {
    int p = prim[slot];
    double rx, ry, rz;
    int feat = 0;
    if (PRIM == 0) {
        double ax, ay, az, bx, by, bz, cx, cy, cz;
        if (DYNAMIC) {
            int a = tris[3 * p], b = tris[3 * p + 1], c = tris[3 * p + 2];
            ax = Xi[3 * a]; ay = Xi[3 * a + 1]; az = Xi[3 * a + 2];
            bx = Xi[3 * b]; by = Xi[3 * b + 1]; bz = Xi[3 * b + 2];
            cx = Xi[3 * c]; cy = Xi[3 * c + 1]; cz = Xi[3 * c + 2];
        } else {
            const double2 *t = reinterpret_cast<const double2 *>(prim_data + (size_t)slot * 10);
            double2 t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2), t3 = __ldg(t + 3), t4 = __ldg(t + 4);
            ax = t0.x; ay = t0.y; az = t1.x; bx = t1.y; by = t2.x; bz = t2.y; cx = t3.x; cy = t3.y; cz = t4.x;
        }
        point_triangle(qx, qy, qz, ax, ay, az, bx, by, bz, cx, cy, cz, rx, ry, rz, feat);
    } else {
        if (DYNAMIC) {
            rx = Xi[3 * p]; ry = Xi[3 * p + 1]; rz = Xi[3 * p + 2];
        } else {
            const double2 *t = reinterpret_cast<const double2 *>(prim_data + (size_t)slot * 4);
            double2 t0 = __ldg(t), t1 = __ldg(t + 1);
            rx = t0.x; ry = t0.y; rz = t1.x;
        }
    }
    double dx = qx - rx, dy = qy - ry, dz = qz - rz;
    double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < h.d2 || (d2 == h.d2 && p < h.prim)) {
        h.d2 = d2; h.x = rx; h.y = ry; h.z = rz; h.prim = p; h.feat = feat; h.slot = slot;
    }
}

__device__ __forceinline__ float box_d2(float lx, float ly, float lz, float hx, float hy, float hz, float qx, float qy,
                                        float qz) {
    float dx = fmaxf(fmaxf(lx - qx, qx - hx), 0.f);
    float dy = fmaxf(fmaxf(ly - qy, qy - hy), 0.f);
    float dz = fmaxf(fmaxf(lz - qz, qz - hz), 0.f);
    return (dx * dx + dy * dy + dz * dz) * 0.999999f;  // conservative: never above the true FP64 distance
}

}  // namespace icp
