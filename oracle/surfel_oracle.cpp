// surfel_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, see msl_oracle.h) restating
// src/SurfelFusion.cpp of razayunus/ManhattanSLAM line by line, plus the compaction tail of
// SurfelMapping::fuseMap (src/SurfelMapping.cpp:366-391).
//
// Determinism choices (the reference is racy / has indeterminate fields):
//  * the THREAD_NUM=10 slices are executed sequentially in slice order 0..9 (row-major pixels,
//    ascending seeds); the `return`-instead-of-`continue` in updateSeedsKernel (:473-474) is kept, so
//    the rest of a 480-seed slice is skipped after the first seed that owns no pixel;
//  * SuperpixelSeed fields left uninitialised by `SuperpixelSeed thisSp;` (:548: size, norm*, pos*,
//    viewCos) are pinned to 0, which is what the preceding memset (:806) intends;
//  * image.at<cv::Vec3b>() on the gray CV_8UC1 image (as actually passed by Tracking.cc:227-229)
//    reads bytes y*step+3x..+2 of the gray buffer; bytes past the end of the buffer read as 0;
//  * Eigen fixed-size products are evaluated left to right, Matrix4f/Matrix4d::inverse() by the
//    cofactor formula ("parity unpinned": Eigen is not available here; differences are O(1 ulp)).
#include "msl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const int ITERATION_NUM = 3;  // include/SurfelFusion.h:33-41
const int THREAD_NUM = 10;
const int SP_SIZE = 8;
#define MAX_ANGLE_COS 0.1
#define HUBER_RANGE 0.4
#define BASELINE 0.5
#define DISPARITY_ERROR 4.0
#define MIN_TOLERATE_DIFF 0.1

// 4x4 inverse by cofactors (stand-in for Eigen::Matrix4{f,d}::inverse()), row-major.
template <typename T>
void inverse4(const T *m, T *inv) {
    T a[16];
    a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    T det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
    det = (T)1 / det;
    for (int i = 0; i < 16; i++) inv[i] = a[i] * det;
}

}  // namespace

struct orc_surfel_fusion {
    int imageWidth, imageHeight, spWidth, spHeight;
    float fx, fy, cx, cy, fuseFar, fuseNear;
    const uint8_t *image = nullptr;
    int imageStep = 0;
    const float *depth = nullptr;
    const int32_t *membership = nullptr;
    int memW = 0;
    std::vector<double> spaceMap;
    std::vector<float> normMap;
    std::vector<orc_seed> superpixelSeeds;
    std::vector<int> superpixelIndex;
    std::vector<orc_seed> seedsIter[ITERATION_NUM];
    std::vector<int> indexIter[ITERATION_NUM];
    orc_surfel *local = nullptr;
    int64_t nLocal = 0;

    float depthAt(int r, int c) const { return depth[(size_t)r * imageWidth + c]; }
    int memberAt(int r, int c) const { return membership[(size_t)r * memW + c]; }
    uint8_t grayAt(int r, int c) const { return image[(size_t)r * imageStep + c]; }
    void vec3bAt(int r, int c, int &v0, int &v1, int &v2) const {  // image.at<cv::Vec3b>(r, c) on CV_8UC1 data
        size_t off = (size_t)r * imageStep + (size_t)c * 3;
        size_t total = (size_t)imageHeight * imageStep;
        v0 = off < total ? image[off] : 0;
        v1 = off + 1 < total ? image[off + 1] : 0;
        v2 = off + 2 < total ? image[off + 2] : 0;
    }

    // :80-85
    void backProject(const float &u, const float &v, const float &d, double &x, double &y, double &z) {
        x = (u - cx) / fx * d;
        y = (v - cy) / fy * d;
        z = d;
    }
    // :87-89
    float getWeight(float &d) { return (float)std::min(1.0 / d / d, 1.0); }

    // :91-165
    void getHuberNorm(float &nx, float &ny, float &nz, float &nb, std::vector<float> &points) {
        int pointNum = (int)points.size() / 3;
        float sumX, sumY, sumZ;
        sumX = sumY = sumZ = 0.0;
        for (int i = 0; i < pointNum; i++) {
            sumX += points[i * 3];
            sumY += points[i * 3 + 1];
            sumZ += points[i * 3 + 2];
        }
        sumX /= pointNum;
        sumY /= pointNum;
        sumZ /= pointNum;
        nb = 0;
        for (int i = 0; i < pointNum; i++) {
            points[i * 3] -= sumX;
            points[i * 3 + 1] -= sumY;
            points[i * 3 + 2] -= sumZ;
        }
        for (int gnI = 0; gnI < 5; gnI++) {
            double H[16] = {0}, J[4] = {0};
            for (int i = 0; i < pointNum; i++) {
                const float p0 = points[i * 3], p1 = points[i * 3 + 1], p2 = points[i * 3 + 2];
                float residual = p0 * nx + p1 * ny + p2 * nz + nb;
                if (residual < HUBER_RANGE && residual > -1 * HUBER_RANGE) {
                    J[0] += 2 * residual * p0;
                    J[1] += 2 * residual * p1;
                    J[2] += 2 * residual * p2;
                    J[3] += 2 * residual;
                    H[0] += 2 * p0 * p0;
                    H[1] += 2 * p0 * p1;
                    H[2] += 2 * p0 * p2;
                    H[3] += 2 * p0;
                    H[4] += 2 * p1 * p0;
                    H[5] += 2 * p1 * p1;
                    H[6] += 2 * p1 * p2;
                    H[7] += 2 * p1;
                    H[8] += 2 * p2 * p0;
                    H[9] += 2 * p2 * p1;
                    H[10] += 2 * p2 * p2;
                    H[11] += 2 * p2;
                    H[12] += 2 * p0;
                    H[13] += 2 * p1;
                    H[14] += 2 * p2;
                    H[15] += 2;
                } else if (residual >= HUBER_RANGE) {
                    J[0] += HUBER_RANGE * p0;
                    J[1] += HUBER_RANGE * p1;
                    J[2] += HUBER_RANGE * p2;
                    J[3] += HUBER_RANGE;
                } else if (residual <= -1 * HUBER_RANGE) {
                    J[0] += -1 * HUBER_RANGE * p0;
                    J[1] += -1 * HUBER_RANGE * p1;
                    J[2] += -1 * HUBER_RANGE * p2;
                    J[3] += -1 * HUBER_RANGE;
                }
            }
            H[0] += 5;
            H[5] += 5;
            H[10] += 5;
            H[15] += 5;
            double Hi[16];
            inverse4<double>(H, Hi);
            double u[4];
            for (int r = 0; r < 4; r++) u[r] = ((Hi[r * 4] * J[0] + Hi[r * 4 + 1] * J[1]) + Hi[r * 4 + 2] * J[2]) + Hi[r * 4 + 3] * J[3];
            nx -= u[0];
            ny -= u[1];
            nz -= u[2];
            nb -= u[3];
        }
        nb = nb - (nx * sumX + ny * sumY + nz * sumZ);
        float normLength = std::sqrt(nx * nx + ny * ny + nz * nz);
        nx /= normLength;
        ny /= normLength;
        nz /= normLength;
        nb /= normLength;
    }

    // :333-355
    bool calculateCost(float &nodepthCost, float &depthCost, const float &pixelIntensity,
                       const float &pixelInverseDepth, const int &x, const int &y, const int &spX, const int &spY) {
        int spIndex = spY * spWidth + spX;
        const orc_seed &sp = superpixelSeeds[spIndex];
        nodepthCost = 0;
        float dist = (sp.x - x) * (sp.x - x) + (sp.y - y) * (sp.y - y);
        nodepthCost += dist / ((SP_SIZE / 2) * (SP_SIZE / 2));
        float intensityDiff = (sp.meanIntensity - pixelIntensity);
        nodepthCost += intensityDiff * intensityDiff / 100.0;
        depthCost = nodepthCost;
        if (sp.meanDepth > 0 && pixelInverseDepth > 0) {
            float inverseDepthDiff = 1.0 / sp.meanDepth - pixelInverseDepth;
            depthCost += inverseDepthDiff * inverseDepthDiff * 400.0;
            return true;
        }
        return false;
    }

    // :357-415
    void updatePixelsKernel(int thread, int threadNum) {
        int stepRow = imageHeight / threadNum;
        int startRow = stepRow * thread;
        int endRow = startRow + stepRow;
        if (thread == threadNum - 1) endRow = imageHeight;
        for (int rowI = startRow; rowI < endRow; rowI++)
            for (int colI = 0; colI < imageWidth; colI++) {
                if (memberAt(rowI / 2, colI / 2) != -1) continue;
                if (superpixelSeeds[superpixelIndex[rowI * imageWidth + colI]].stable) continue;
                float myIntensity = grayAt(rowI, colI);
                float myInvDepth = 0.0;
                if (depthAt(rowI, colI) > 0.01) myInvDepth = 1.0 / depthAt(rowI, colI);
                int baseSpX = colI / SP_SIZE;
                int baseSpY = rowI / SP_SIZE;
                float minDistDepth = 1e6;
                int minSpIndexDepth = -1;
                float minDistNodepth = 1e6;
                int minSpIndexNodepth = -1;
                bool allHasDepth = true;
                for (int checkI = -1; checkI <= 1; checkI++)
                    for (int checkJ = -1; checkJ <= 1; checkJ++) {
                        int checkSpX = baseSpX + checkI;
                        int checkSpY = baseSpY + checkJ;
                        int distSpX = std::abs(checkSpX * SP_SIZE + SP_SIZE / 2 - colI);
                        int distSpY = std::abs(checkSpY * SP_SIZE + SP_SIZE / 2 - rowI);
                        if (distSpX < SP_SIZE && distSpY < SP_SIZE && checkSpX >= 0 && checkSpX < spWidth &&
                            checkSpY >= 0 && checkSpY < spHeight) {
                            float distDepth, distNodepth;
                            allHasDepth &= calculateCost(distNodepth, distDepth, myIntensity, myInvDepth, colI, rowI,
                                                         checkSpX, checkSpY);
                            if (distDepth < minDistDepth) {
                                minDistDepth = distDepth;
                                minSpIndexDepth = (baseSpY + checkJ) * spWidth + baseSpX + checkI;
                            }
                            if (distNodepth < minDistNodepth) {
                                minDistNodepth = distNodepth;
                                minSpIndexNodepth = (baseSpY + checkJ) * spWidth + baseSpX + checkI;
                            }
                        }
                    }
                if (allHasDepth) {
                    superpixelIndex[rowI * imageWidth + colI] = minSpIndexDepth;
                    superpixelSeeds[minSpIndexDepth].stable = false;
                } else {
                    superpixelIndex[rowI * imageWidth + colI] = minSpIndexNodepth;
                    superpixelSeeds[minSpIndexNodepth].stable = false;
                }
            }
    }

    // :428-515
    void updateSeedsKernel(int thread, int threadNum) {
        int step = (int)superpixelSeeds.size() / threadNum;
        int beginIndex = step * thread;
        int endIndex = beginIndex + step;
        if (thread == threadNum - 1) endIndex = (int)superpixelSeeds.size();
        for (int seedI = beginIndex; seedI < endIndex; seedI++) {
            if (!superpixelSeeds[seedI].use) continue;
            if (superpixelSeeds[seedI].stable) continue;
            int spX = seedI % spWidth;
            int spY = seedI / spWidth;
            int checkXBegin = spX * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
            int checkYBegin = spY * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
            int checkXEnd = checkXBegin + SP_SIZE * 2;
            int checkYEnd = checkYBegin + SP_SIZE * 2;
            checkXBegin = checkXBegin > 0 ? checkXBegin : 0;
            checkYBegin = checkYBegin > 0 ? checkYBegin : 0;
            checkXEnd = checkXEnd < imageWidth - 1 ? checkXEnd : imageWidth - 1;
            checkYEnd = checkYEnd < imageHeight - 1 ? checkYEnd : imageHeight - 1;
            float sumX = 0;
            float sumY = 0;
            float sumIntensity = 0.0;
            float sumIntensityNum = 0.0;
            float sumDepth = 0.0;
            float sumDepthNum = 0.0;
            std::vector<float> depthVector;
            for (int checkJ = checkYBegin; checkJ < checkYEnd; checkJ++)
                for (int checkI = checkXBegin; checkI < checkXEnd; checkI++) {
                    int pixelIndex = checkJ * imageWidth + checkI;
                    if (superpixelIndex[pixelIndex] == seedI) {
                        sumX += checkI;
                        sumY += checkJ;
                        sumIntensityNum += 1.0;
                        sumIntensity += grayAt(checkJ, checkI);
                        float checkDepth = depthAt(checkJ, checkI);
                        if (checkDepth > 0.1) {
                            depthVector.push_back(checkDepth);
                            sumDepth += checkDepth;
                            sumDepthNum += 1.0;
                        }
                    }
                }
            if (sumIntensityNum == 0) return;  // sic: `return`, not `continue` (:473-474)
            sumIntensity /= sumIntensityNum;
            sumX /= sumIntensityNum;
            sumY /= sumIntensityNum;
            float preIntensity = superpixelSeeds[seedI].meanIntensity;
            float preX = superpixelSeeds[seedI].x;
            float preY = superpixelSeeds[seedI].y;
            superpixelSeeds[seedI].meanIntensity = sumIntensity;
            superpixelSeeds[seedI].x = sumX;
            superpixelSeeds[seedI].y = sumY;
            vec3bAt((int)sumY, (int)sumX, superpixelSeeds[seedI].r, superpixelSeeds[seedI].g, superpixelSeeds[seedI].b);
            float updateDiff = std::fabs(preIntensity - sumIntensity) + std::fabs(preX - sumX) + std::fabs(preY - sumY);
            if (updateDiff < 0.2) superpixelSeeds[seedI].stable = true;
            if (sumDepthNum > 0) {
                float meanDepth = sumDepth / sumDepthNum;
                float sumA, sumB;
                for (int newtonI = 0; newtonI < 5; newtonI++) {
                    sumA = sumB = 0;
                    for (size_t pI = 0; pI < depthVector.size(); pI++) {
                        float residual = meanDepth - depthVector[pI];
                        if (residual < HUBER_RANGE && residual > -HUBER_RANGE) {
                            sumA += 2 * residual;
                            sumB += 2;
                        } else {
                            sumA += residual > 0 ? HUBER_RANGE : -1 * HUBER_RANGE;
                        }
                    }
                    float deltaDepth = -sumA / (sumB + 10.0);
                    meanDepth = meanDepth + deltaDepth;
                    if (deltaDepth < 0.01 && deltaDepth > -0.01) break;
                }
                superpixelSeeds[seedI].meanDepth = meanDepth;
            } else {
                superpixelSeeds[seedI].meanDepth = 0.0;
            }
        }
    }

    // :528-584
    void initializeSeedsKernel(int thread, int threadNum) {
        int step = (int)superpixelSeeds.size() / threadNum;
        int beginIndex = step * thread;
        int endIndex = beginIndex + step;
        if (thread == threadNum - 1) endIndex = (int)superpixelSeeds.size();
        for (int seedI = beginIndex; seedI < endIndex; seedI++) {
            int spX = seedI % spWidth;
            int spY = seedI / spWidth;
            int imageX = spX * SP_SIZE + SP_SIZE / 2;
            int imageY = spY * SP_SIZE + SP_SIZE / 2;
            imageX = imageX < (imageWidth - 1) ? imageX : (imageWidth - 1);
            imageY = imageY < (imageHeight - 1) ? imageY : (imageHeight - 1);
            if (memberAt(imageY / 2, imageX / 2) != -1) {
                superpixelSeeds[seedI].use = false;
                continue;
            }
            orc_seed thisSp;
            std::memset(&thisSp, 0, sizeof(thisSp));  // indeterminate fields pinned to 0
            thisSp.use = true;                        // `bool use = true;` default member initialiser
            thisSp.x = (float)imageX;
            thisSp.y = (float)imageY;
            vec3bAt(imageY, imageX, thisSp.r, thisSp.g, thisSp.b);
            thisSp.meanIntensity = grayAt(imageY, imageX);
            thisSp.fused = false;
            thisSp.stable = false;
            thisSp.meanDepth = depthAt(imageY, imageX);
            if (thisSp.meanDepth < 0.01) {
                int checkXBegin = spX * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
                int checkYBegin = spY * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
                int checkXEnd = checkXBegin + SP_SIZE * 2;
                int checkYEnd = checkYBegin + SP_SIZE * 2;
                checkXBegin = checkXBegin > 0 ? checkXBegin : 0;
                checkYBegin = checkYBegin > 0 ? checkYBegin : 0;
                checkXEnd = checkXEnd < imageWidth - 1 ? checkXEnd : imageWidth - 1;
                checkYEnd = checkYEnd < imageHeight - 1 ? checkYEnd : imageHeight - 1;
                bool findDepth = false;
                for (int checkJ = checkYBegin; checkJ < checkYEnd; checkJ++) {
                    for (int checkI = checkXBegin; checkI < checkXEnd; checkI++) {
                        float thisDepth = depthAt(checkJ, checkI);
                        if (thisDepth > 0.01) {
                            thisSp.meanDepth = thisDepth;
                            findDepth = true;
                            break;
                        }
                    }
                    if (findDepth) break;
                }
            }
            superpixelSeeds[seedI] = thisSp;
        }
    }

    // :597-613
    void calculateSpacesKernel(int thread, int threadNum) {
        int stepRow = imageHeight / threadNum;
        int startRow = stepRow * thread;
        int endRow = startRow + stepRow;
        if (thread == threadNum - 1) endRow = imageHeight;
        for (int rowI = startRow; rowI < endRow; rowI++)
            for (int colI = 0; colI < imageWidth; colI++) {
                int myIndex = rowI * imageWidth + colI;
                float myDepth = depthAt(rowI, colI);
                double x, y, z;
                backProject((float)colI, (float)rowI, myDepth, x, y, z);
                spaceMap[myIndex * 3] = x;
                spaceMap[myIndex * 3 + 1] = y;
                spaceMap[myIndex * 3 + 2] = z;
            }
    }

    // :615-661
    void calculatePixelsNormsKernel(int thread, int threadNum) {
        int stepRow = imageHeight / threadNum;
        int startRow = stepRow * thread;
        startRow = startRow > 1 ? startRow : 1;
        int endRow = startRow + stepRow;
        if (thread == threadNum - 1) endRow = imageHeight - 1;
        for (int rowI = startRow; rowI < endRow; rowI++)
            for (int colI = 1; colI < imageWidth - 1; colI++) {
                int myIndex = rowI * imageWidth + colI;
                float myX, myY, myZ;
                myX = spaceMap[myIndex * 3];
                myY = spaceMap[myIndex * 3 + 1];
                myZ = spaceMap[myIndex * 3 + 2];
                float rightX, rightY, rightZ;
                rightX = spaceMap[myIndex * 3 + 3];
                rightY = spaceMap[myIndex * 3 + 4];
                rightZ = spaceMap[myIndex * 3 + 5];
                float downX, downY, downZ;
                downX = spaceMap[myIndex * 3 + imageWidth * 3];
                downY = spaceMap[myIndex * 3 + imageWidth * 3 + 1];
                downZ = spaceMap[myIndex * 3 + imageWidth * 3 + 2];
                if (myZ < 0.1 || rightZ < 0.1 || downZ < 0.1) continue;
                rightX = rightX - myX;
                rightY = rightY - myY;
                rightZ = rightZ - myZ;
                downX = downX - myX;
                downY = downY - myY;
                downZ = downZ - myZ;
                float normX, normY, normZ, normLength;
                normX = rightY * downZ - rightZ * downY;
                normY = rightZ * downX - rightX * downZ;
                normZ = rightX * downY - rightY * downX;
                normLength = std::sqrt(normX * normX + normY * normY + normZ * normZ);
                normX /= normLength;
                normY /= normLength;
                normZ /= normLength;
                float viewAngle = (normX * myX + normY * myY + normZ * myZ) / std::sqrt(myX * myX + myY * myY + myZ * myZ);
                if (viewAngle > -MAX_ANGLE_COS && viewAngle < MAX_ANGLE_COS) continue;
                normMap[myIndex * 3] = normX;
                normMap[myIndex * 3 + 1] = normY;
                normMap[myIndex * 3 + 2] = normZ;
            }
    }

    // :663-773
    void calculateSpDepthNormsKernel(int thread, int threadNum) {
        int step = (int)superpixelSeeds.size() / threadNum;
        int beginIndex = step * thread;
        int endIndex = beginIndex + step;
        if (thread == threadNum - 1) endIndex = (int)superpixelSeeds.size();
        for (int seedI = beginIndex; seedI < endIndex; seedI++) {
            int spX = seedI % spWidth;
            int spY = seedI / spWidth;
            int checkXBegin = spX * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
            int checkYBegin = spY * SP_SIZE + SP_SIZE / 2 - SP_SIZE;
            std::vector<float> pixelDepth;
            std::vector<float> pixelNorms;
            std::vector<float> pixelPositions;
            std::vector<float> pixelInlierPositions;
            float validDepthNum = 0;
            float maxDist = 0;
            for (int checkJ = checkYBegin; checkJ < (checkYBegin + SP_SIZE * 2); checkJ++) {
                for (int checkI = checkXBegin; checkI < (checkXBegin + SP_SIZE * 2); checkI++) {
                    int pixelIndex = checkJ * imageWidth + checkI;
                    if (pixelIndex < 0 || pixelIndex >= (int)superpixelIndex.size()) continue;
                    if (superpixelIndex[pixelIndex] == seedI) {
                        float xDiff = checkI - superpixelSeeds[seedI].x;
                        float yDiff = checkJ - superpixelSeeds[seedI].y;
                        float dist = xDiff * xDiff + yDiff * yDiff;
                        if (dist > maxDist) maxDist = dist;
                        // depth.at<float>(checkJ, checkI) with checkI possibly <0 or >=W: flat address checkJ*W+checkI
                        float myDepth = depth[pixelIndex];
                        if (myDepth > 0.05) {
                            pixelDepth.push_back(myDepth);
                            pixelNorms.push_back(normMap[pixelIndex * 3]);
                            pixelNorms.push_back(normMap[pixelIndex * 3 + 1]);
                            pixelNorms.push_back(normMap[pixelIndex * 3 + 2]);
                            validDepthNum += 1;
                            pixelPositions.push_back(spaceMap[pixelIndex * 3]);
                            pixelPositions.push_back(spaceMap[pixelIndex * 3 + 1]);
                            pixelPositions.push_back(spaceMap[pixelIndex * 3 + 2]);
                        }
                    }
                }
            }
            if (validDepthNum < 16) continue;
            float meanDepth = superpixelSeeds[seedI].meanDepth;
            float normX, normY, normZ, normB;
            normX = normY = normZ = normB = 0.0;
            float inlierNum = 0;
            for (size_t pI = 0; pI < pixelDepth.size(); pI++) {
                float residual = meanDepth - pixelDepth[pI];
                if (residual < HUBER_RANGE && residual > -HUBER_RANGE) {
                    normX += pixelNorms[pI * 3];
                    normY += pixelNorms[pI * 3 + 1];
                    normZ += pixelNorms[pI * 3 + 2];
                    inlierNum += 1;
                    pixelInlierPositions.push_back(pixelPositions[pI * 3]);
                    pixelInlierPositions.push_back(pixelPositions[pI * 3 + 1]);
                    pixelInlierPositions.push_back(pixelPositions[pI * 3 + 2]);
                }
            }
            if (inlierNum / pixelDepth.size() < 0.8) continue;
            float normLength = std::sqrt(normX * normX + normY * normY + normZ * normZ);
            normX = normX / normLength;
            normY = normY / normLength;
            normZ = normZ / normLength;
            {
                float gnNx = normX, gnNy = normY, gnNz = normZ, gnNb = normB;
                getHuberNorm(gnNx, gnNy, gnNz, gnNb, pixelInlierPositions);
                normX = gnNx;
                normY = gnNy;
                normZ = gnNz;
                normB = gnNb;
            }
            double avgX, avgY, avgZ;
            backProject(superpixelSeeds[seedI].x, superpixelSeeds[seedI].y, meanDepth, avgX, avgY, avgZ);
            {
                float k = -1 * (avgX * normX + avgY * normY + avgZ * normZ) - normB;
                avgX += k * normX;
                avgY += k * normY;
                avgZ += k * normZ;
                meanDepth = avgZ;
            }
            float viewCos = -1.0 * (normX * avgX + normY * avgY + normZ * avgZ) / std::sqrt(avgX * avgX + avgY * avgY + avgZ * avgZ);
            if (viewCos < 0) {
                viewCos *= -1.0;
                normX *= -1.0;
                normY *= -1.0;
                normZ *= -1.0;
            }
            superpixelSeeds[seedI].normX = normX;
            superpixelSeeds[seedI].normY = normY;
            superpixelSeeds[seedI].normZ = normZ;
            superpixelSeeds[seedI].posX = avgX;
            superpixelSeeds[seedI].posY = avgY;
            superpixelSeeds[seedI].posZ = avgZ;
            superpixelSeeds[seedI].meanDepth = meanDepth;
            superpixelSeeds[seedI].viewCos = viewCos;
            superpixelSeeds[seedI].size = std::sqrt(maxDist);
        }
    }

    // :805-816
    void generateSuperPixels() {
        std::memset(superpixelSeeds.data(), 0, superpixelSeeds.size() * sizeof(orc_seed));
        std::fill(superpixelIndex.begin(), superpixelIndex.end(), 0);
        std::fill(normMap.begin(), normMap.end(), 0.f);
        for (int t = 0; t < THREAD_NUM; t++) initializeSeedsKernel(t, THREAD_NUM);
        for (int itI = 0; itI < ITERATION_NUM; itI++) {
            for (int t = 0; t < THREAD_NUM; t++) updatePixelsKernel(t, THREAD_NUM);
            for (int t = 0; t < THREAD_NUM; t++) updateSeedsKernel(t, THREAD_NUM);
            seedsIter[itI] = superpixelSeeds;
            indexIter[itI] = superpixelIndex;
        }
        for (int t = 0; t < THREAD_NUM; t++) calculateSpacesKernel(t, THREAD_NUM);
        for (int t = 0; t < THREAD_NUM; t++) calculatePixelsNormsKernel(t, THREAD_NUM);
        for (int t = 0; t < THREAD_NUM; t++) calculateSpDepthNormsKernel(t, THREAD_NUM);
    }

    // :167-283
    void fuseSurfelsKernel(int thread, int threadNum, int referenceFrameIndex, const float *pose, const float *invPose) {
        int64_t step = nLocal / threadNum;
        int64_t beginIndex = step * thread;
        int64_t endIndex = beginIndex + step;
        if (thread == threadNum - 1) endIndex = nLocal;
        orc_surfel *localSurfels = local;
        for (int64_t i = beginIndex; i < endIndex; i++) {
            if (referenceFrameIndex - localSurfels[i].lastUpdate > 5 && localSurfels[i].updateTimes < 5) {
                localSurfels[i].updateTimes = 0;
                continue;
            }
            if (localSurfels[i].updateTimes == 0) continue;
            const float pw0 = localSurfels[i].px, pw1 = localSurfels[i].py, pw2 = localSurfels[i].pz;
            float pc[3];
            for (int r = 0; r < 3; r++)
                pc[r] = ((invPose[r * 4] * pw0 + invPose[r * 4 + 1] * pw1) + invPose[r * 4 + 2] * pw2) + invPose[r * 4 + 3] * 1.0f;
            if (pc[2] < fuseNear || pc[2] > fuseFar) continue;
            const float nw0 = localSurfels[i].nx, nw1 = localSurfels[i].ny, nw2 = localSurfels[i].nz;
            float normC[3];
            for (int r = 0; r < 3; r++) normC[r] = (invPose[r * 4] * nw0 + invPose[r * 4 + 1] * nw1) + invPose[r * 4 + 2] * nw2;
            float projectU = pc[0] * fx / pc[2] + cx;  // project(), :75-78
            float projectV = pc[1] * fy / pc[2] + cy;
            int pUInt = projectU + 0.5;
            int pVInt = projectV + 0.5;
            if (pUInt < 1 || pUInt > imageWidth - 2 || pVInt < 1 || pVInt > imageHeight - 2) continue;
            if (pc[2] < depthAt(pVInt, pUInt) - 1.0) {
                localSurfels[i].updateTimes = 0;
                continue;
            }
            int spIndex = superpixelIndex[pVInt * imageWidth + pUInt];
            orc_seed &sp = superpixelSeeds[spIndex];
            if (sp.normX == 0 && sp.normY == 0 && sp.normZ == 0) continue;
            if (sp.viewCos < MAX_ANGLE_COS) continue;
            float cameraF = (std::fabs(fx) + std::fabs(fy)) / 2.0;
            float tolerateDiff = pc[2] * pc[2] / (BASELINE * cameraF) * DISPARITY_ERROR;
            tolerateDiff = tolerateDiff < MIN_TOLERATE_DIFF ? MIN_TOLERATE_DIFF : tolerateDiff;
            if (pc[2] < sp.meanDepth - tolerateDiff) continue;
            if (pc[2] > sp.meanDepth + tolerateDiff) continue;
            float normDiffCos = normC[0] * sp.normX + normC[1] * sp.normY + normC[2] * sp.normZ;
            if (normDiffCos < MAX_ANGLE_COS) {
                localSurfels[i].updateTimes = 0;
                continue;
            }
            float oldWeigth = localSurfels[i].weight;
            float newWeight = getWeight(sp.meanDepth);
            float sumWeight = oldWeigth + newWeight;
            float spPW[3];
            for (int r = 0; r < 3; r++)
                spPW[r] = ((pose[r * 4] * sp.posX + pose[r * 4 + 1] * sp.posY) + pose[r * 4 + 2] * sp.posZ) + pose[r * 4 + 3] * 1.0f;
            float fusedPx = (localSurfels[i].px * oldWeigth + newWeight * spPW[0]) / sumWeight;
            float fusedPy = (localSurfels[i].py * oldWeigth + newWeight * spPW[1]) / sumWeight;
            float fusedPz = (localSurfels[i].pz * oldWeigth + newWeight * spPW[2]) / sumWeight;
            float fusedNx = normC[0] * oldWeigth + newWeight * sp.normX;
            float fusedNy = normC[1] * oldWeigth + newWeight * sp.normY;
            float fusedNz = normC[2] * oldWeigth + newWeight * sp.normZ;
            double newNormLength = std::sqrt(fusedNx * fusedNx + fusedNy * fusedNy + fusedNz * fusedNz);
            fusedNx /= newNormLength;
            fusedNy /= newNormLength;
            fusedNz /= newNormLength;
            float newNormW[3];
            for (int r = 0; r < 3; r++) newNormW[r] = (pose[r * 4] * fusedNx + pose[r * 4 + 1] * fusedNy) + pose[r * 4 + 2] * fusedNz;
            localSurfels[i].px = fusedPx;
            localSurfels[i].py = fusedPy;
            localSurfels[i].pz = fusedPz;
            localSurfels[i].r = sp.r;
            localSurfels[i].g = sp.g;
            localSurfels[i].b = sp.b;
            localSurfels[i].nx = newNormW[0];
            localSurfels[i].ny = newNormW[1];
            localSurfels[i].nz = newNormW[2];
            localSurfels[i].weight = sumWeight;
            localSurfels[i].color = sp.meanIntensity;
            float newSize = sp.size * std::fabs(sp.meanDepth / (cameraF * sp.viewCos));
            if (newSize < localSurfels[i].size) localSurfels[i].size = newSize;
            localSurfels[i].lastUpdate = referenceFrameIndex;
            localSurfels[i].updateTimes += 1;
            sp.fused = true;
        }
    }

    // :285-331
    int initializeSurfels(int referenceFrameIndex, const float *pose, orc_surfel *newSurfels, int cap) {
        int n = 0;
        for (size_t i = 0; i < superpixelSeeds.size(); i++) {
            orc_seed &sp = superpixelSeeds[i];
            if (sp.meanDepth == 0) continue;
            if (sp.fused) continue;
            if (sp.viewCos < MAX_ANGLE_COS) continue;
            if (sp.normX == 0 && sp.normY == 0 && sp.normZ == 0) continue;
            float pW[3], nW[3];
            for (int r = 0; r < 3; r++) {
                pW[r] = ((pose[r * 4] * sp.posX + pose[r * 4 + 1] * sp.posY) + pose[r * 4 + 2] * sp.posZ) + pose[r * 4 + 3] * 1.0f;
                nW[r] = (pose[r * 4] * sp.normX + pose[r * 4 + 1] * sp.normY) + pose[r * 4 + 2] * sp.normZ;
            }
            orc_surfel e;
            e.px = pW[0], e.py = pW[1], e.pz = pW[2];
            e.r = sp.r, e.g = sp.g, e.b = sp.b;
            e.nx = nW[0], e.ny = nW[1], e.nz = nW[2];
            float cameraF = (std::fabs(fx) + std::fabs(fy)) / 2.0;
            float newSize = sp.size * std::fabs(sp.meanDepth / (cameraF * sp.viewCos));
            e.size = newSize;
            e.color = sp.meanIntensity;
            e.weight = getWeight(sp.meanDepth);
            e.updateTimes = 1;
            e.lastUpdate = referenceFrameIndex;
            if (n < cap) newSurfels[n] = e;
            n++;
        }
        return n;
    }
};

extern "C" {

orc_surfel_fusion *orc_surfel_create(int w, int h, float fx, float fy, float cx, float cy, float fuseFar,
                                     float fuseNear) {
    orc_surfel_fusion *s = new orc_surfel_fusion();
    s->imageWidth = w, s->imageHeight = h;
    s->spWidth = w / SP_SIZE, s->spHeight = h / SP_SIZE;
    s->fx = fx, s->fy = fy, s->cx = cx, s->cy = cy, s->fuseFar = fuseFar, s->fuseNear = fuseNear;
    s->superpixelSeeds.resize((size_t)s->spWidth * s->spHeight);
    s->superpixelIndex.resize((size_t)w * h);
    s->spaceMap.resize((size_t)w * h * 3);
    s->normMap.resize((size_t)w * h * 3);
    return s;
}
void orc_surfel_destroy(orc_surfel_fusion *s) { delete s; }

int orc_surfel_fuse(orc_surfel_fusion *s, int referenceFrameIndex, const uint8_t *gray, int gray_stride,
                    const float *depth, const int32_t *membership, const float Twc[16], orc_surfel *local,
                    int64_t n_local, orc_surfel *new_surfels, int cap_new, int threads) {
    s->image = gray, s->imageStep = gray_stride, s->depth = depth, s->membership = membership;
    s->memW = (s->imageWidth + 1) / 2;
    s->local = local, s->nLocal = n_local;
    s->generateSuperPixels();
    float invPose[16];
    inverse4<float>(Twc, invPose);
    if (threads <= 1) {
        for (int t = 0; t < THREAD_NUM; t++) s->fuseSurfelsKernel(t, THREAD_NUM, referenceFrameIndex, Twc, invPose);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back(&orc_surfel_fusion::fuseSurfelsKernel, s, t, threads, referenceFrameIndex, Twc, invPose);
        for (auto &t : pool) t.join();
    }
    return s->initializeSurfels(referenceFrameIndex, Twc, new_surfels, cap_new);
}

const int32_t *orc_surfel_index(const orc_surfel_fusion *s) { return s->superpixelIndex.data(); }
const orc_seed *orc_surfel_seeds(const orc_surfel_fusion *s) { return s->superpixelSeeds.data(); }
const float *orc_surfel_normmap(const orc_surfel_fusion *s) { return s->normMap.data(); }
const orc_seed *orc_surfel_seeds_iter(const orc_surfel_fusion *s, int it) { return s->seedsIter[it].data(); }
const int32_t *orc_surfel_index_iter(const orc_surfel_fusion *s, int it) { return s->indexIter[it].data(); }

// SurfelMapping::fuseMap tail, src/SurfelMapping.cpp:366-391
int64_t orc_surfel_compact(orc_surfel *local, int64_t n_local, const orc_surfel *new_surfels, int n_new) {
    std::vector<int64_t> deletedIndex;
    for (int64_t i = 0; i < n_local; i++)
        if (local[i].updateTimes == 0) deletedIndex.push_back(i);
    int64_t size = n_local;
    for (int i = 0; i < n_new; i++) {
        if (new_surfels[i].updateTimes != 0) {
            if (!deletedIndex.empty()) {
                local[deletedIndex.back()] = new_surfels[i];
                deletedIndex.pop_back();
            } else
                local[size++] = new_surfels[i];
        }
    }
    while (!deletedIndex.empty()) {
        local[deletedIndex.back()] = local[size - 1];
        deletedIndex.pop_back();
        size--;
    }
    return size;
}

// ------------------------------------------------------------------------------------------------
// SurfelMapping::moveAddSurfels (src/SurfelMapping.cpp:194-304) with the state it touches: posesDatabase
// (include/SurfelMapping.h:39-46: attachedSurfels, pointsBeginIndex, pointsPoseIndex), pointcloudPoseIndex and
// Map::mvInactiveSurfels.  posesToAdd / posesToRemove (getAddRemovePoses :306-326, a walk over the pose graph)
// are inputs.  linkedPoseIndex / localSurfelsIndexs are the caller's business.
struct orc_surfel_mapping {
    struct PoseElement {
        std::vector<orc_surfel> attachedSurfels;
        int pointsBeginIndex = -1, pointsPoseIndex = -1;
    };
    std::vector<PoseElement> posesDatabase;
    std::vector<int> pointcloudPoseIndex;
    std::vector<orc_surfel> inactive;  // Map::mvInactiveSurfels
};

orc_surfel_mapping *orc_mapping_create() { return new orc_surfel_mapping(); }
void orc_mapping_destroy(orc_surfel_mapping *m) { delete m; }

int64_t orc_move_add_surfels(orc_surfel_mapping *m, orc_surfel *local, int64_t n_local, int64_t cap_local,
                             const int32_t *posesToRemove, int n_remove, const int32_t *posesToAdd, int n_add) {
    auto &posesDatabase = m->posesDatabase;
    auto &pointcloudPoseIndex = m->pointcloudPoseIndex;
    int maxPose = -1;
    for (int i = 0; i < n_remove; i++) maxPose = std::max(maxPose, posesToRemove[i]);
    for (int i = 0; i < n_add; i++) maxPose = std::max(maxPose, posesToAdd[i]);
    if ((int)posesDatabase.size() < maxPose + 1) posesDatabase.resize(maxPose + 1);
    if (n_remove > 0) {  // :200-229
        for (int r = 0; r < n_remove; r++) {
            const int inactiveIndex = posesToRemove[r];
            posesDatabase[inactiveIndex].pointsBeginIndex = (int)m->inactive.size();
            posesDatabase[inactiveIndex].pointsPoseIndex = (int)pointcloudPoseIndex.size();
            pointcloudPoseIndex.push_back(inactiveIndex);
            for (int64_t i = 0; i < n_local; i++) {
                orc_surfel &localSurfel = local[i];
                if (localSurfel.updateTimes > 0 && localSurfel.lastUpdate == inactiveIndex) {
                    posesDatabase[inactiveIndex].attachedSurfels.push_back(localSurfel);
                    m->inactive.push_back(localSurfel);
                    localSurfel.updateTimes = 0;  // delete the surfel from the local map
                }
            }
        }
    }
    if (n_add > 0) {  // :230-303
        std::vector<std::pair<int, int>> removeInfo;
        for (int addI = 0; addI < n_add; addI++) {
            const int addIndex = posesToAdd[addI];
            if (posesDatabase[addIndex].pointsPoseIndex < 0) return -1;  // never moved out: the reference would index [-1]
            removeInfo.push_back(std::make_pair(posesDatabase[addIndex].pointsPoseIndex, addIndex));
        }
        std::sort(removeInfo.begin(), removeInfo.end(),
                  [](const std::pair<int, int> &first, const std::pair<int, int> &second) { return first.first < second.first; });
        int removeBeginIndex = removeInfo[0].second;
        int removePointsSize = (int)posesDatabase[removeBeginIndex].attachedSurfels.size();
        int removePoseSize = 1;
        for (int removeI = 1; removeI <= (int)removeInfo.size(); removeI++) {
            bool needRemove = false;
            if (removeI == (int)removeInfo.size()) needRemove = true;
            if (removeI < (int)removeInfo.size())
                if (removeInfo[removeI].first != (removeInfo[removeI - 1].first + 1)) needRemove = true;
            if (!needRemove) {
                const int thisPoseIndex = removeInfo[removeI].second;
                removePointsSize += (int)posesDatabase[thisPoseIndex].attachedSurfels.size();
                removePoseSize += 1;
                continue;
            }
            const int removeEndIndex = removeInfo[removeI - 1].second;
            auto beginPtr = m->inactive.begin() + posesDatabase[removeBeginIndex].pointsBeginIndex;
            m->inactive.erase(beginPtr, beginPtr + removePointsSize);
            for (int pi = posesDatabase[removeEndIndex].pointsPoseIndex + 1; pi < (int)pointcloudPoseIndex.size(); pi++) {
                posesDatabase[pointcloudPoseIndex[pi]].pointsBeginIndex -= removePointsSize;
                posesDatabase[pointcloudPoseIndex[pi]].pointsPoseIndex -= removePoseSize;
            }
            pointcloudPoseIndex.erase(pointcloudPoseIndex.begin() + posesDatabase[removeBeginIndex].pointsPoseIndex,
                                      pointcloudPoseIndex.begin() + posesDatabase[removeEndIndex].pointsPoseIndex + 1);
            if (removeI < (int)removeInfo.size()) {
                removeBeginIndex = removeInfo[removeI].second;
                removePointsSize = (int)posesDatabase[removeBeginIndex].attachedSurfels.size();
                removePoseSize = 1;
            }
        }
        for (int pi = 0; pi < n_add; pi++) {  // :292-302 append to mvLocalSurfels
            const int pose_index = posesToAdd[pi];
            auto &att = posesDatabase[pose_index].attachedSurfels;
            if (n_local + (int64_t)att.size() > cap_local) return -2;
            for (const orc_surfel &e : att) local[n_local++] = e;
            att.clear();
            posesDatabase[pose_index].pointsBeginIndex = -1;
            posesDatabase[pose_index].pointsPoseIndex = -1;
        }
    }
    return n_local;
}

int64_t orc_mapping_inactive(const orc_surfel_mapping *m, orc_surfel *out, int64_t cap) {
    const int64_t n = (int64_t)m->inactive.size();
    if (out)
        for (int64_t i = 0; i < n && i < cap; i++) out[i] = m->inactive[i];
    return n;
}

}  // extern "C"
