// One-shot depth pre-processing of SRPS::execute (reference SRPS.cu:117-149, devicecalls.cu:95-125)
// without OpenCV: mean over the depth frames, Telea fast-marching inpainting (radius 16) of the pixels
// where any frame is 0, bilateral filter (d=-1, sigmaColor=2, sigmaSpace=2) on depth/max, bicubic (a=-0.75)
// upsampling.  Like the reference, every step runs on the TRANSPOSED low-resolution image (the reference wraps its
// column-major buffer in a cv::Mat(z0_w rows, z0_h cols), SRPS.cu:130-132) -- for these isotropic filters that
// only fixes the memory order.  This is host init ("next" row N2 of SURVEY §8f), off the hot path: it follows the
// published algorithms (Telea 2004; Tomasi-Manduchi; Keys cubic convolution as OpenCV parametrises it), it is not
// bit-identical to OpenCV 3.3 -- tests/test_cpp_host.py bounds the difference against python cv2 on Mitten.
#include <algorithm>
#include <cmath>
#include <queue>
#include <vector>

#include "Utilities.h"

namespace {

// ---- Telea inpainting on a rows x cols float image (row-major) -------------------------------------------------
enum : unsigned char { KNOWN = 0, BAND = 1, INSIDE = 2 };

struct HeapItem { float t; int idx; bool operator>(const HeapItem& o) const { return t > o.t; } };
using MinHeap = std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>>;

inline float eikonal(const std::vector<float>& T, const std::vector<unsigned char>& f, int rows, int cols, int i1, int j1,
                     int i2, int j2) {
    // solve |grad T| = 1 from the two neighbours (i1,j1), (i2,j2) (Telea 2004, fig. 4)
    float sol = 1e6f;
    const bool in1 = i1 >= 0 && i1 < rows && j1 >= 0 && j1 < cols, in2 = i2 >= 0 && i2 < rows && j2 >= 0 && j2 < cols;
    const bool k1 = in1 && f[i1 * cols + j1] == KNOWN, k2 = in2 && f[i2 * cols + j2] == KNOWN;
    const float t1 = in1 ? T[i1 * cols + j1] : 1e6f, t2 = in2 ? T[i2 * cols + j2] : 1e6f;
    if (k1) {
        if (k2) {
            const float r = std::sqrt(std::max(0.f, 2.f - (t1 - t2) * (t1 - t2)));
            float s = (t1 + t2 - r) * 0.5f;
            if (s >= t1 && s >= t2) sol = s;
            else { s += r; if (s >= t1 && s >= t2) sol = s; }
        } else {
            sol = 1.f + t1;
        }
    } else if (k2) {
        sol = 1.f + t2;
    }
    return sol;
}

void fmm_distance(std::vector<float>& T, std::vector<unsigned char>& f, int rows, int cols, float limit) {
    // fast marching of T over the INSIDE pixels of f starting from its BAND (no inpainting): used for the distance
    // field OUTSIDE the hole, which gives grad T a meaning on the known side of the boundary
    MinHeap heap;
    for (int i = 0; i < rows * cols; i++) if (f[i] == BAND) heap.push({T[i], i});
    const int di[4] = {-1, 0, 1, 0}, dj[4] = {0, -1, 0, 1};
    while (!heap.empty()) {
        HeapItem it = heap.top(); heap.pop();
        if (f[it.idx] == KNOWN) continue;
        f[it.idx] = KNOWN;
        if (it.t > limit) continue;
        const int i = it.idx / cols, j = it.idx % cols;
        for (int q = 0; q < 4; q++) {
            const int a = i + di[q], b = j + dj[q];
            if (a < 0 || a >= rows || b < 0 || b >= cols || f[a * cols + b] != INSIDE) continue;
            const float t = std::min(std::min(eikonal(T, f, rows, cols, a - 1, b, a, b - 1), eikonal(T, f, rows, cols, a + 1, b, a, b - 1)),
                                     std::min(eikonal(T, f, rows, cols, a - 1, b, a, b + 1), eikonal(T, f, rows, cols, a + 1, b, a, b + 1)));
            T[a * cols + b] = t;
            f[a * cols + b] = BAND;
            heap.push({t, a * cols + b});
        }
    }
}

void inpaint_telea(std::vector<float>& img, const std::vector<unsigned char>& hole, int rows, int cols, int radius) {
    const int n = rows * cols;
    std::vector<unsigned char> f(n);
    std::vector<float> T(n, 0.f);
    const int di[4] = {-1, 0, 1, 0}, dj[4] = {0, -1, 0, 1};
    bool any = false;
    for (int i = 0; i < n; i++) { f[i] = hole[i] ? INSIDE : KNOWN; any |= hole[i] != 0; }
    if (!any) return;
    // distance field on the known side (negative), within 2*radius of the hole
    {
        std::vector<unsigned char> fo(n);
        std::vector<float> To(n, 0.f);
        for (int i = 0; i < n; i++) { fo[i] = hole[i] ? KNOWN : INSIDE; To[i] = hole[i] ? 0.f : 1e6f; }
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) {
                if (!hole[i * cols + j]) continue;
                for (int q = 0; q < 4; q++) {
                    const int a = i + di[q], b = j + dj[q];
                    if (a >= 0 && a < rows && b >= 0 && b < cols && !hole[a * cols + b]) { fo[i * cols + j] = BAND; break; }
                }
            }
        fmm_distance(To, fo, rows, cols, 2.f * radius);
        for (int i = 0; i < n; i++) if (!hole[i]) T[i] = To[i] < 1e5f ? -To[i] : -2.f * radius;
    }
    MinHeap heap;
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            const int id = i * cols + j;
            if (f[id] == INSIDE) { T[id] = 1e6f; continue; }
            for (int q = 0; q < 4; q++) {
                const int a = i + di[q], b = j + dj[q];
                if (a >= 0 && a < rows && b >= 0 && b < cols && f[a * cols + b] == INSIDE) { f[id] = BAND; T[id] = 0.f; heap.push({0.f, id}); break; }
            }
        }
    while (!heap.empty()) {
        HeapItem it = heap.top(); heap.pop();
        if (f[it.idx] == KNOWN) continue;
        f[it.idx] = KNOWN;
        const int i = it.idx / cols, j = it.idx % cols;
        for (int q = 0; q < 4; q++) {
            const int a = i + di[q], b = j + dj[q];
            if (a < 0 || a >= rows || b < 0 || b >= cols || f[a * cols + b] != INSIDE) continue;
            const float t = std::min(std::min(eikonal(T, f, rows, cols, a - 1, b, a, b - 1), eikonal(T, f, rows, cols, a + 1, b, a, b - 1)),
                                     std::min(eikonal(T, f, rows, cols, a - 1, b, a, b + 1), eikonal(T, f, rows, cols, a + 1, b, a, b + 1)));
            T[a * cols + b] = t;
            // weighted average of the known pixels within `radius` (Telea 2004, eq. 2-3) with the weights and the
            // normalised first-order term as OpenCV's implementation parametrises them (one-sided differences
            // next to unknown pixels, direction term r.gradT not normalised, distance term 1/|r|^3)
            auto fl = [&](int k, int l) -> unsigned char { return (k < 0 || k >= rows || l < 0 || l >= cols) ? (unsigned char)INSIDE : f[k * cols + l]; };
            auto grad = [&](const std::vector<float>& A, int k, int l, int dk, int dl, float both) -> float {
                const bool nx = fl(k + dk, l + dl) != INSIDE, pv = fl(k - dk, l - dl) != INSIDE;
                if (nx) return pv ? (A[(k + dk) * cols + l + dl] - A[(k - dk) * cols + l - dl]) * both
                                  : A[(k + dk) * cols + l + dl] - A[k * cols + l];
                return pv ? A[k * cols + l] - A[(k - dk) * cols + l - dl] : 0.f;
            };
            const float gx = grad(T, a, b, 0, 1, 0.5f), gy = grad(T, a, b, 1, 0, 0.5f);
            double acc = 0.0, wsum = 0.0, Jx = 0.0, Jy = 0.0;
            for (int k = a - radius; k <= a + radius; k++) {
                if (k < 0 || k >= rows) continue;
                for (int l = b - radius; l <= b + radius; l++) {
                    if (l < 0 || l >= cols || f[k * cols + l] == INSIDE) continue;
                    const float ry = (float)(a - k), rx = (float)(b - l), d2 = rx * rx + ry * ry;
                    if (d2 > (float)(radius * radius) || d2 == 0.f) continue;
                    float dir = rx * gx + ry * gy;
                    if (std::fabs(dir) <= 0.01f) dir = 1e-6f;
                    const float dst = 1.f / (d2 * std::sqrt(d2));
                    const float lev = 1.f / (1.f + std::fabs(T[k * cols + l] - t));
                    const double wgt = std::fabs(dir * dst * lev);
                    acc += wgt * img[k * cols + l];
                    Jx -= wgt * grad(img, k, l, 0, 1, 2.0f) * rx;
                    Jy -= wgt * grad(img, k, l, 1, 0, 2.0f) * ry;
                    wsum += wgt;
                }
            }
            if (wsum > 0.0) img[a * cols + b] = (float)(acc / wsum + (Jx + Jy) / (std::sqrt(Jx * Jx + Jy * Jy) + 1.0e-20));
            f[a * cols + b] = BAND;
            heap.push({t, a * cols + b});
        }
    }
}

inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// cv::bilateralFilter(src, dst, -1, sigmaColor, sigmaSpace) for one float channel: radius = round(1.5 sigmaSpace),
// circular support, BORDER_REFLECT_101
void bilateral(const std::vector<float>& src, std::vector<float>& dst, int rows, int cols, float sigma_color, float sigma_space) {
    const int radius = std::max(1, (int)std::lround(sigma_space * 1.5f));
    const float gc = -0.5f / (sigma_color * sigma_color), gs = -0.5f / (sigma_space * sigma_space);
    dst.resize(src.size());
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            const float c = src[(size_t)i * cols + j];
            float acc = 0.f, wsum = 0.f;
            for (int di = -radius; di <= radius; di++)
                for (int dj = -radius; dj <= radius; dj++) {
                    const float r2 = (float)(di * di + dj * dj);
                    if (std::sqrt(r2) > (float)radius) continue;
                    const float v = src[(size_t)reflect101(i + di, rows) * cols + reflect101(j + dj, cols)];
                    const float w = std::exp(r2 * gs + (v - c) * (v - c) * gc);
                    acc += w * v; wsum += w;
                }
            dst[(size_t)i * cols + j] = acc / wsum;
        }
}

// cv::resize(..., INTER_CUBIC) for one float channel: Keys kernel with a = -0.75, pixel-centre mapping, replicated border
void cubic_weights(float t, float w[4]) {
    const float A = -0.75f;
    w[0] = ((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A;
    w[1] = ((A + 2) * t - (A + 3)) * t * t + 1;
    w[2] = ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1;
    w[3] = 1.f - w[0] - w[1] - w[2];
}

void resize_cubic(const std::vector<float>& src, int rows, int cols, std::vector<float>& dst, int orows, int ocols) {
    std::vector<float> tmp((size_t)rows * ocols);
    const double sx = (double)cols / ocols, sy = (double)rows / orows;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++)
        for (int x = 0; x < ocols; x++) {
            const float fx = (float)((x + 0.5) * sx - 0.5);
            const int ix = (int)std::floor(fx);
            float w[4]; cubic_weights(fx - ix, w);
            float acc = 0.f;
            for (int k = 0; k < 4; k++) acc += w[k] * src[(size_t)i * cols + std::min(std::max(ix - 1 + k, 0), cols - 1)];
            tmp[(size_t)i * ocols + x] = acc;
        }
    dst.resize((size_t)orows * ocols);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < orows; y++) {
        const float fy = (float)((y + 0.5) * sy - 0.5);
        const int iy = (int)std::floor(fy);
        float w[4]; cubic_weights(fy - iy, w);
        for (int x = 0; x < ocols; x++) {
            float acc = 0.f;
            for (int k = 0; k < 4; k++) acc += w[k] * tmp[(size_t)std::min(std::max(iy - 1 + k, 0), rows - 1) * ocols + x];
            dst[(size_t)y * ocols + x] = acc;
        }
    }
}

}  // namespace

void preprocess_depth(const float* z0, int z0_h, int z0_w, int z0_n, int I_h, int I_w, std::vector<float>& zs,
                      std::vector<float>& z_full) {
    const int n = z0_h * z0_w;
    // mean over the frames, always divided by the frame count; a pixel is flagged if ANY frame is 0 (devicecalls.cu:95-110)
    std::vector<float> mean(n, 0.f);
    std::vector<unsigned char> hole(n, 0);
    for (int p = 0; p < n; p++) {
        float avg = 0.f;
        for (int c = 0; c < z0_n; c++) {
            const float v = z0[(size_t)c * n + p];
            if (v != 0.f) avg += v; else hole[p] = 1;
        }
        mean[p] = avg / (float)z0_n;
    }
    // the column-major (z0_h x z0_w) buffer seen as a row-major image of z0_w rows x z0_h cols (SRPS.cu:130-132)
    const int rows = z0_w, cols = z0_h;
    inpaint_telea(mean, hole, rows, cols, 16);                                       // SRPS.cu:133
    float mx = mean[0];
    for (float v : mean) mx = std::max(mx, v);                                      // SRPS.cu:137
    std::vector<float> norm(n), sm;
    for (int p = 0; p < n; p++) norm[p] = mean[p] / mx;                             // SRPS.cu:138
    bilateral(norm, sm, rows, cols, 2.f, 2.f);                                       // SRPS.cu:139
    zs.resize(n);
    for (int p = 0; p < n; p++) zs[p] = sm[p] * mx;                                  // SRPS.cu:140
    resize_cubic(zs, rows, cols, z_full, I_w, I_h);                                  // cv::Size(I_h, I_w): I_w rows x I_h cols  SRPS.cu:149
}
