// icp_device.cuh - device helpers shared by the kernels of libicpcuda.so.
#pragma once

#include "icp_internal.h"

namespace icp {

#define ICP_LOG_2PI 1.8378770664093454835606594728112

__device__ __forceinline__ void pose_matrix(const double *th, double R[9]) {
    // Scalismo Rotation(phi, theta, psi, centre): R = Rz(phi) Ry(theta) Rx(psi)
    double sph, cph, sth, cth, sps, cps;
    sincos(th[4], &sph, &cph);
    sincos(th[5], &sth, &cth);
    sincos(th[6], &sps, &cps);
    R[0] = cth * cph; R[1] = sps * sth * cph - cps * sph; R[2] = sps * sph + cps * sth * cph;
    R[3] = cth * sph; R[4] = cps * cph + sps * sth * sph; R[5] = cps * sth * sph - sps * cph;
    R[6] = -sth;      R[7] = sps * cth;                   R[8] = cps * cth;
}


// FP64 tensor-pipe instruction (SASS DMMA): C (8 x 8) += A (8 x 4) B (4 x 8). Lane l holds a = A[l / 4][l % 4],
// b = B[l % 4][l / 4], c0 / c1 = C[l / 4][2 (l % 4) + 0 / 1].
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Scalismo vertexNormals: normalised unweighted mean of the unit cell normals of the adjacent
// triangles, adjacency in ascending triangle id
__device__ __forceinline__ void vertex_normal_dev(const ModelDev &m, const double *__restrict__ Xc, int v, double &nx, double &ny,
                                  double &nz) {
    double sx = 0, sy = 0, sz = 0;
    int b = m.adj_off[v], e = m.adj_off[v + 1];
    for (int k = b; k < e; k++) {
        int t = m.adj[k];
        int i1 = m.tris[3 * t], i2 = m.tris[3 * t + 1], i3 = m.tris[3 * t + 2];
        double ux = Xc[3 * i2] - Xc[3 * i1], uy = Xc[3 * i2 + 1] - Xc[3 * i1 + 1], uz = Xc[3 * i2 + 2] - Xc[3 * i1 + 2];
        double vx = Xc[3 * i3] - Xc[3 * i1], vy = Xc[3 * i3 + 1] - Xc[3 * i1 + 1], vz = Xc[3 * i3 + 2] - Xc[3 * i1 + 2];
        double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
        double nrm = sqrt(cx * cx + cy * cy + cz * cz);
        sx += cx / nrm; sy += cy / nrm; sz += cz / nrm;
    }
    double cnt = (double)(e - b);
    sx /= cnt; sy /= cnt; sz /= cnt;
    double nrm = sqrt(sx * sx + sy * sy + sz * sz);
    nx = sx / nrm; ny = sy / nrm; nz = sz / nrm;
}


// inverse of poseTransform (NonRigidIcpProposal.scala:142): x -> R^T (x - t - c) + c
__device__ __forceinline__ void inverse_pose(const double *th, const double R[9], double x, double y, double z,
                                             double &ox, double &oy, double &oz) {
    double ax = x - th[1] - th[7], ay = y - th[2] - th[8], az = z - th[3] - th[9];
    ox = (R[0] * ax + R[3] * ay + R[6] * az) + th[7];
    oy = (R[1] * ax + R[4] * ay + R[7] * az) + th[8];
    oz = (R[2] * ax + R[5] * ay + R[8] * az) + th[9];
}

// block-wide sum (blockDim.x <= 1024, multiple of 32); result valid in every thread
__device__ __forceinline__ double block_sum(double v, double *red /* >= 33 doubles of shared memory */) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = l < nw ? red[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// w = A z for a row-major Kp x Kp matrix in global memory, z / w in shared memory: warps deal the rows, lanes the columns
__device__ __forceinline__ void block_matvec_rows(const double *__restrict__ A, int Kp, const double *z_sm, double *w_sm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = warp; i < Kp; i += nw) {
        const double *row = A + (size_t)i * Kp;
        double acc = 0.0;
        for (int j = lane; j < Kp; j += 32) acc = fma(__ldg(row + j), z_sm[j], acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) w_sm[i] = acc;
    }
}

// Philox4x32-10 (Salmon et al. 2011)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// counter layout of the chain runner: key = seed (lo, hi); counter = (chain lo, chain hi, step, block)
__device__ __forceinline__ uint4 chain_philox(unsigned long long seed, unsigned long long chain, unsigned int step,
                                              unsigned int block) {
    return philox4x32_10(make_uint4((unsigned int)chain, (unsigned int)(chain >> 32), step, block),
                         make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
}
// 53-bit uniform in [0, 1) from two 32-bit words
__device__ __forceinline__ double u53(unsigned int hi, unsigned int lo) {
    return (double)(((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

}  // namespace icp
