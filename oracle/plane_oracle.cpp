// plane_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, see msl_oracle.h) restating the plane pre-stage:
// PlaneDetection::readDepthImage (src/PlaneExtractor.cpp:44-76), ImagePointCloud::get
// (include/PlaneExtractor.h:48-56), PlaneSeg ctor + Stats (include/peac/AHCPlaneSeg.hpp:59-181,235-312),
// ParamSet thresholds (include/peac/AHCParamSet.hpp:68-142) and the node/edge initialisation of
// PlaneFitter::initGraph (include/peac/AHCPlaneFitter.hpp:756-928).
// Eigen::SelfAdjointEigenSolver<Matrix3d> (include/peac/eig33sym.hpp:71-75) is not available here; it is
// replaced by a cyclic Jacobi solver ("parity unpinned" for the solver; cross-checked against LAPACK via
// numpy.linalg.eigh in tests/test_oracle_plane.py).
#include "msl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <set>
#include <utility>
#include <vector>

namespace {

// ParamSet defaults (AHCParamSet.hpp:68-75); ManhattanSLAM never overrides them.
const double depthSigma = 1.6e-6, stdTol_init = 5, z_near = 500, z_far = 4000;
const double angle_near = 15.0 * M_PI / 180.0, angle_far = 90.0 * M_PI / 180.0;
const double depthAlpha = 0.04, depthChangeTol = 0.02;
const int windowWidth = 10, windowHeight = 10;  // AHCPlaneFitter.hpp:156-160

inline double T_mse_init(double z) {  // :92 std::pow(x, 2)
    const double v = depthSigma * z * z + stdTol_init;
    return v * v;
}
inline double T_ang_init(double z) {  // :113-118
    double clipped_z = z;
    clipped_z = std::max(clipped_z, z_near);
    clipped_z = std::min(clipped_z, z_far);
    const double factor = (angle_far - angle_near) / (z_far - z_near);
    return std::cos(factor * clipped_z + angle_near - factor * z_near);
}
inline double T_dz(double z) { return depthAlpha * std::fabs(z) + depthChangeTol; }  // :140-142
inline bool depthDisContinuous(double d0, double d1) { return std::fabs(d0 - d1) > T_dz(d0); }

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations; eigenvalues ascending, V columns = vectors.
void eig33sym(const double K[9], double s[3], double V[9]) {
    double a[3][3] = {{K[0], K[1], K[2]}, {K[3], K[4], K[5]}, {K[6], K[7], K[8]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 32; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-60 || off <= 1e-34 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; k++) {  // A <- A*J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {  // A <- J^T*A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    double e[3] = {a[0][0], a[1][1], a[2][2]};
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++)
            if (e[ord[j]] < e[ord[i]]) std::swap(ord[i], ord[j]);
    for (int i = 0; i < 3; i++) {
        s[i] = e[ord[i]];
        for (int k = 0; k < 3; k++) V[k * 3 + i] = v[k][ord[i]];
    }
}


// P1-P5; sums9 (optional): the nine running sums of Stats per block (sx sy sz sxx syy szz sxy syz sxz), 0 for rejected blocks
void prestage(const uint16_t *depth, int w, int h, int dstride_px, float fx, float fy, float cx, float cy, float depthMapFactor,
              double *cloud_xyz, orc_block_stat *blocks, uint8_t *seed, uint8_t *edges, std::vector<double> *sums9) {
    const int W2 = (int)std::ceil(w / 2.0), H2 = (int)std::ceil(h / 2.0);
    std::vector<double> cloudLocal;
    double *cloud = cloud_xyz;
    if (!cloud) {
        cloudLocal.resize((size_t)W2 * H2 * 3);
        cloud = cloudLocal.data();
    }
    // readDepthImage, src/PlaneExtractor.cpp:60-74
    int vertex_idx = 0;
    for (int i = 0; i < h; i += 2)
        for (int j = 0; j < w; j += 2) {
            double z = (double)(depth[(size_t)i * dstride_px + j]) * depthMapFactor;
            double x = ((double)j - cx) * z / fx;
            double y = ((double)i - cy) * z / fy;
            cloud[vertex_idx * 3] = x, cloud[vertex_idx * 3 + 1] = y, cloud[vertex_idx * 3 + 2] = z;
            vertex_idx++;
        }
    auto get = [&](int row, int col, double &x, double &y, double &z) -> bool {  // ImagePointCloud::get
        const int pixIdx = row * W2 + col;
        z = cloud[pixIdx * 3 + 2];
        if (z == 0 || std::isnan(z)) return false;
        x = cloud[pixIdx * 3], y = cloud[pixIdx * 3 + 1];
        return true;
    };
    const int Nh = H2 / windowHeight, Nw = W2 / windowWidth;
    std::vector<orc_block_stat> local((size_t)Nh * Nw);
    std::vector<uint8_t> G((size_t)Nh * Nw, 0);
    if (sums9) sums9->assign((size_t)Nh * Nw * 9, 0.0);
    for (int bi = 0; bi < Nh; ++bi)
        for (int bj = 0; bj < Nw; ++bj) {
            // PlaneSeg ctor, AHCPlaneSeg.hpp:235-312 (INIT_STRICT)
            orc_block_stat &B = local[bi * Nw + bj];
            std::memset(&B, 0, sizeof(B));
            double sx = 0, sy = 0, sz = 0, sxx = 0, syy = 0, szz = 0, sxy = 0, syz = 0, sxz = 0;
            int N = 0;
            bool windowValid = true;
            const int seed_row = bi * windowHeight, seed_col = bj * windowWidth;
            for (int i = seed_row, icnt = 0; icnt < windowHeight && i < H2; ++i, ++icnt) {
                for (int j = seed_col, jcnt = 0; jcnt < windowWidth && j < W2; ++j, ++jcnt) {
                    double x = 0, y = 0, z = 10000;
                    if (!get(i, j, x, y, z)) {
                        windowValid = false;
                        break;
                    }
                    double xn = 0, yn = 0, zn = 10000;
                    if (j + 1 < W2 && (get(i, j + 1, xn, yn, zn) && depthDisContinuous(z, zn))) {
                        windowValid = false;
                        break;
                    }
                    if (i + 1 < H2 && (get(i + 1, j, xn, yn, zn) && depthDisContinuous(z, zn))) {
                        windowValid = false;
                        break;
                    }
                    sx += x, sy += y, sz += z;
                    sxx += x * x, syy += y * y, szz += z * z;
                    sxy += x * y, syz += y * z, sxz += x * z;
                    ++N;
                }
                if (!windowValid) break;
            }
            if (windowValid) {
                B.nouse = 0;
                B.N = N;
                if (sums9) {
                    double *q = sums9->data() + 9 * (size_t)(bi * Nw + bj);
                    q[0] = sx, q[1] = sy, q[2] = sz, q[3] = sxx, q[4] = syy, q[5] = szz, q[6] = sxy, q[7] = syz, q[8] = sxz;
                }
            } else {
                B.N = 0;
                B.nouse = 1;
            }
            if (B.N < 4) {
                B.mse = B.curvature = std::numeric_limits<double>::quiet_NaN();
            } else {
                // Stats::compute, :148-181
                const double sc = 1.0 / N;
                B.center[0] = sx * sc, B.center[1] = sy * sc, B.center[2] = sz * sc;
                double K[9] = {sxx - sx * sx * sc, sxy - sx * sy * sc, sxz - sx * sz * sc, 0, syy - sy * sy * sc,
                               syz - sy * sz * sc, 0, 0, szz - sz * sz * sc};
                K[3] = K[1], K[6] = K[2], K[7] = K[5];
                double sv[3], V[9];
                eig33sym(K, sv, V);
                if (V[0] * B.center[0] + V[3] * B.center[1] + V[6] * B.center[2] <= 0) {
                    B.normal[0] = V[0], B.normal[1] = V[3], B.normal[2] = V[6];
                } else {
                    B.normal[0] = -V[0], B.normal[1] = -V[3], B.normal[2] = -V[6];
                }
                B.mse = sv[0] * sc;
                B.curvature = sv[0] / (sv[0] + sv[1] + sv[2]);
            }
            // initGraph node test, AHCPlaneFitter.hpp:777-778
            G[bi * Nw + bj] = (B.mse < T_mse_init(B.center[2]) && !B.nouse) ? 1 : 0;
        }
    std::vector<uint8_t> E((size_t)Nh * Nw, 0);
    auto sim = [&](int a, int b) {
        const orc_block_stat &A = local[a], &Bk = local[b];
        return std::abs(A.normal[0] * Bk.normal[0] + A.normal[1] * Bk.normal[1] + A.normal[2] * Bk.normal[2]);
    };
    // first pass, row direction (:840-874): bit0 = left neighbour, bit1 = right neighbour
    for (int i = 0; i < Nh; ++i) {
        for (int j = 1; j < Nw; j += 2) {
            const int cidx = i * Nw + j;
            if (G[cidx - 1] == 0) {
                --j;
                continue;
            }
            if (G[cidx] == 0) continue;
            if (j < Nw - 1 && G[cidx + 1] == 0) {
                ++j;
                continue;
            }
            const double similarityTh = T_ang_init(local[cidx].center[2]);
            if ((j < Nw - 1 && sim(cidx - 1, cidx + 1) >= similarityTh) || (j == Nw - 1 && sim(cidx, cidx - 1) >= similarityTh)) {
                E[cidx] |= 1, E[cidx - 1] |= 2;
                if (j < Nw - 1) E[cidx] |= 2, E[cidx + 1] |= 1;
            } else {
                --j;
            }
        }
    }
    // second pass, column direction (:876-910): bit2 = up, bit3 = down
    for (int j = 0; j < Nw; ++j) {
        for (int i = 1; i < Nh; i += 2) {
            const int cidx = i * Nw + j;
            if (G[cidx - Nw] == 0) {
                --i;
                continue;
            }
            if (G[cidx] == 0) continue;
            if (i < Nh - 1 && G[cidx + Nw] == 0) {
                ++i;
                continue;
            }
            const double similarityTh = T_ang_init(local[cidx].center[2]);
            if ((i < Nh - 1 && sim(cidx - Nw, cidx + Nw) >= similarityTh) || (i == Nh - 1 && sim(cidx, cidx - Nw) >= similarityTh)) {
                E[cidx] |= 4, E[cidx - Nw] |= 8;
                if (i < Nh - 1) E[cidx] |= 8, E[cidx + Nw] |= 4;
            } else {
                --i;
            }
        }
    }
    if (blocks) std::memcpy(blocks, local.data(), local.size() * sizeof(orc_block_stat));
    if (seed) std::memcpy(seed, G.data(), G.size());
    if (edges) std::memcpy(edges, E.data(), E.size());
}

}  // namespace

#include "peac_oracle.inc"

extern "C" {

void orc_eig33sym(const double K[9], double s[3], double V[9]) { eig33sym(K, s, V); }

void orc_plane_prestage(const uint16_t *depth, int w, int h, int dstride_px, float fx, float fy, float cx, float cy,
                        float depthMapFactor, double *cloud_xyz, orc_block_stat *blocks, uint8_t *seed, uint8_t *edges) {
    prestage(depth, w, h, dstride_px, fx, fy, cx, cy, depthMapFactor, cloud_xyz, blocks, seed, edges, nullptr);
}

}  // extern "C"
