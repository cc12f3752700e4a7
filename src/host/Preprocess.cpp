// One-shot depth pre-processing of SRPS::execute (reference SRPS.cu:117-149, devicecalls.cu:95-125)
// without OpenCV: mean over the depth frames, Telea fast-marching inpainting (radius 16) of the pixels
// where any frame is 0, bilateral filter (d=-1, sigmaColor=2, sigmaSpace=2) on depth/max, bicubic (a=-0.75)
// upsampling.  Like the reference, every step runs on the TRANSPOSED low-resolution image (the reference wraps its
// column-major buffer in a cv::Mat(z0_w rows, z0_h cols), SRPS.cu:130-132) -- for these isotropic filters that
// only fixes the memory order.  This is host init ("next" row N2 of SURVEY §8f), off the hot path: it follows the
// published algorithms (Telea 2004; Tomasi-Manduchi; Keys cubic convolution) in OpenCV's parametrisation and agrees
// with python cv2 to a few ulp (tests/test_cpp_host.py: synthetic scenes with depth drop-outs and Mitten).
#include <algorithm>
#include <cmath>
#include <queue>
#include <vector>

#include <stdexcept>
#include <string>

#include "../../include/srps_c_api.h"
#include "Utilities.h"

namespace {

// ---- Telea inpainting on a rows x cols float image (row-major) -------------------------------------------------
// Fast-marching inpainting (A. Telea, "An image inpainting technique based on the fast marching method", 2004) in the
// parametrisation of cv::inpaint(src, mask, dst, range, INPAINT_TELEA), which is what the reference calls
// (SRPS.cu:133): a one-pixel frame around the image, flags KNOWN / BAND / INSIDE, a stable priority queue (ties
// leave in insertion order), the arrival time T marched outwards over the (2 range + 1)-square dilation of the hole
// first (negated: it gives grad T a meaning on the known side), then inwards; every new pixel is the weighted mean of
// the known pixels within `range` with weight |dir * dst * lev| (dst = 1/|r|^3, lev = 1/(1 + |T_q - T_p|),
// dir = r . grad T, floored at 1e-6) plus the normalised first-order term, one-sided differences next to unknown
// pixels, single-precision accumulators.  tests/test_cpp_host.py holds it to python cv2 within a few ulp.
enum : unsigned char { KNOWN = 0, BAND = 1, INSIDE = 2, CHANGE = 3 };

struct HeapItem {
    float t; unsigned seq; int i, j;
    bool operator>(const HeapItem& o) const { return t > o.t || (t == o.t && seq > o.seq); }
};
struct StableQueue {                 // smallest T first, equal T in insertion order
    std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>> q;
    unsigned seq = 0;
    void push(int i, int j, float t) { q.push({t, seq++, i, j}); }
    bool pop(int& i, int& j) {
        if (q.empty()) return false;
        i = q.top().i; j = q.top().j; q.pop();
        return true;
    }
};

struct Frame {                        // (rows + 2) x (cols + 2) planes
    int er, ec;
    std::vector<unsigned char> f;
    std::vector<float> t;
    unsigned char& F(int i, int j) { return f[(size_t)i * ec + j]; }
    float& T(int i, int j) { return t[(size_t)i * ec + j]; }
};

// arrival time at a pixel from two of its neighbours (upwind solution of |grad T| = 1)
inline float eikonal(Frame& m, std::vector<unsigned char>& flags, int i1, int j1, int i2, int j2) {
    const double a11 = m.T(i1, j1), a22 = m.T(i2, j2), m12 = std::min(a11, a22);
    const bool k1 = flags[(size_t)i1 * m.ec + j1] != INSIDE, k2 = flags[(size_t)i2 * m.ec + j2] != INSIDE;
    double sol;
    if (k1) {
        if (k2) sol = std::fabs(a11 - a22) >= 1.0 ? 1 + m12 : (a11 + a22 + std::sqrt(2 - (a11 - a22) * (a11 - a22))) * 0.5;
        else sol = 1 + a11;
    } else {
        sol = k2 ? 1 + a22 : 1 + m12;
    }
    return (float)sol;
}

inline float eikonal4(Frame& m, std::vector<unsigned char>& flags, int i, int j) {
    return std::min(std::min(eikonal(m, flags, i - 1, j, i, j - 1), eikonal(m, flags, i + 1, j, i, j - 1)),
                    std::min(eikonal(m, flags, i - 1, j, i, j + 1), eikonal(m, flags, i + 1, j, i, j + 1)));
}

void inpaint_telea(std::vector<float>& img, const std::vector<unsigned char>& hole, int rows, int cols, int range) {
    bool any = false;
    for (unsigned char h : hole) any |= h != 0;
    if (!any) return;
    Frame m;
    m.er = rows + 2; m.ec = cols + 2;
    const int er = m.er, ec = m.ec;
    m.f.assign((size_t)er * ec, KNOWN);
    m.t.assign((size_t)er * ec, 1.0e6f);
    std::vector<unsigned char> mask((size_t)er * ec, 0), band((size_t)er * ec, 0), ring((size_t)er * ec, 0);
    auto at = [&](std::vector<unsigned char>& v, int i, int j) -> unsigned char& { return v[(size_t)i * ec + j]; };
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++)
            if (hole[(size_t)i * cols + j]) at(mask, i + 1, j + 1) = INSIDE;
    // narrow band: known pixels 4-adjacent to the hole
    for (int i = 1; i < er - 1; i++)
        for (int j = 1; j < ec - 1; j++)
            if (!at(mask, i, j) && (at(mask, i - 1, j) || at(mask, i + 1, j) || at(mask, i, j - 1) || at(mask, i, j + 1))) at(band, i, j) = 1;
    // ring: known pixels within Chebyshev distance `range` of the hole, band excluded (separable square dilation)
    {
        std::vector<unsigned char> tmp((size_t)er * ec, 0);
        for (int i = 0; i < er; i++)
            for (int j = 0; j < ec; j++) {
                unsigned char v = 0;
                for (int l = std::max(0, j - range); l <= std::min(ec - 1, j + range) && !v; l++) v = at(mask, i, l);
                at(tmp, i, j) = v;
            }
        for (int i = 1; i < er - 1; i++)
            for (int j = 1; j < ec - 1; j++) {
                unsigned char v = 0;
                for (int k = std::max(0, i - range); k <= std::min(er - 1, i + range) && !v; k++) v = at(tmp, k, j);
                if (v && !at(mask, i, j) && !at(band, i, j)) at(ring, i, j) = INSIDE;
            }
    }
    StableQueue heap, outer;
    for (int i = 0; i < er; i++)
        for (int j = 0; j < ec; j++) {
            if (at(band, i, j)) { heap.push(i, j, 0.f); outer.push(i, j, 0.f); m.F(i, j) = BAND; m.T(i, j) = 0.f; }
            if (at(mask, i, j)) m.F(i, j) = INSIDE;
        }
    // T on the known side: march over the ring, then negate
    {
        int ii, jj;
        while (outer.pop(ii, jj)) {
            at(ring, ii, jj) = CHANGE;
            const int ni[4] = {ii - 1, ii, ii + 1, ii}, nj[4] = {jj, jj - 1, jj, jj + 1};
            for (int q = 0; q < 4; q++) {
                const int i = ni[q], j = nj[q];
                if (i <= 0 || j <= 0 || i >= er - 1 || j >= ec - 1) continue;
                if (at(ring, i, j) != INSIDE) continue;
                const float dist = eikonal4(m, ring, i, j);
                m.T(i, j) = dist;
                at(ring, i, j) = BAND;
                outer.push(i, j, dist);
            }
        }
        for (size_t p = 0; p < ring.size(); p++)
            if (ring[p] == CHANGE) m.t[p] = -m.t[p];
    }
    // inwards: T and the value of every hole pixel
    auto I = [&](int k, int l) -> float { return img[(size_t)k * cols + l]; };
    int ii, jj;
    while (heap.pop(ii, jj)) {
        m.F(ii, jj) = KNOWN;
        const int ni[4] = {ii - 1, ii, ii + 1, ii}, nj[4] = {jj, jj - 1, jj, jj + 1};
        for (int q = 0; q < 4; q++) {
            const int i = ni[q], j = nj[q];
            if (i <= 0 || j <= 0 || i > er - 1 || j > ec - 1) continue;
            if (m.F(i, j) != INSIDE) continue;
            const float dist = eikonal4(m, m.f, i, j);
            m.T(i, j) = dist;
            float gTx, gTy;
            if (m.F(i, j + 1) != INSIDE) gTx = m.F(i, j - 1) != INSIDE ? (m.T(i, j + 1) - m.T(i, j - 1)) * 0.5f : m.T(i, j + 1) - m.T(i, j);
            else gTx = m.F(i, j - 1) != INSIDE ? m.T(i, j) - m.T(i, j - 1) : 0.f;
            if (m.F(i + 1, j) != INSIDE) gTy = m.F(i - 1, j) != INSIDE ? (m.T(i + 1, j) - m.T(i - 1, j)) * 0.5f : m.T(i + 1, j) - m.T(i, j);
            else gTy = m.F(i - 1, j) != INSIDE ? m.T(i, j) - m.T(i - 1, j) : 0.f;
            float Ia = 0.f, Jx = 0.f, Jy = 0.f, s = 1.0e-20f;
            for (int k = i - range; k <= i + range; k++) {
                if (k <= 0 || k >= er - 1) continue;
                const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);          // clamped rows of the one-sided differences
                for (int l = j - range; l <= j + range; l++) {
                    if (l <= 0 || l >= ec - 1) continue;
                    if (m.F(k, l) == INSIDE || (l - j) * (l - j) + (k - i) * (k - i) > range * range) continue;
                    const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
                    const float ry = (float)(i - k), rx = (float)(j - l);
                    const float len2 = rx * rx + ry * ry;
                    const float dst = (float)(1.0 / ((double)len2 * std::sqrt((double)len2)));
                    const float lev = (float)(1.0 / (double)(1.f + std::fabs(m.T(k, l) - m.T(i, j))));
                    float dir = rx * gTx + ry * gTy;
                    if (std::fabs((double)dir) <= 0.01) dir = 0.000001f;
                    const float w = std::fabs(dst * lev * dir);
                    float gIx, gIy;
                    if (m.F(k, l + 1) != INSIDE) gIx = m.F(k, l - 1) != INSIDE ? (I(km, lp + 1) - I(km, lm - 1)) * 2.0f : I(km, lp + 1) - I(km, lm);
                    else gIx = m.F(k, l - 1) != INSIDE ? I(km, lp) - I(km, lm - 1) : 0.f;
                    if (m.F(k + 1, l) != INSIDE) gIy = m.F(k - 1, l) != INSIDE ? (I(kp + 1, lm) - I(km - 1, lm)) * 2.0f : I(kp + 1, lm) - I(km, lm);
                    else gIy = m.F(k - 1, l) != INSIDE ? I(kp, lm) - I(km - 1, lm) : 0.f;
                    Ia += w * I(k - 1, l - 1);
                    Jx -= w * (gIx * rx);
                    Jy -= w * (gIy * ry);
                    s += w;
                }
            }
            img[(size_t)(i - 1) * cols + (j - 1)] = Ia / s + (Jx + Jy) / (std::sqrt(Jx * Jx + Jy * Jy) + 1.0e-20f);
            m.F(i, j) = BAND;
            heap.push(i, j, dist);
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

// Default path: the parallel steps on the device (csrc/srps_init.cuh: mean + flags, max, bilateral, bicubic), the
// fast-marching inpainting between them on the host.  Same arithmetic as the host path below.
static void preprocess_depth_device(const float* z0, int z0_h, int z0_w, int z0_n, int I_h, int I_w, std::vector<float>& zs,
                                    std::vector<float>& z_full) {
    const int n = z0_h * z0_w;
    std::vector<float> mean(n);
    std::vector<unsigned char> hole(n);
    if (srps_init_depth_mean(Preferences::deviceId, z0, n, z0_n, mean.data(), hole.data()))
        throw std::runtime_error(std::string("srps_init_depth_mean: ") + srps_last_error(nullptr));
    const int rows = z0_w, cols = z0_h;                                              // SRPS.cu:130-132
    inpaint_telea(mean, hole, rows, cols, 16);                                       // SRPS.cu:133
    zs.resize(n);
    z_full.resize((size_t)I_w * I_h);
    if (srps_init_depth_smooth_upsample(Preferences::deviceId, mean.data(), rows, cols, I_w, I_h, 2.f, 2.f, zs.data(), z_full.data()))
        throw std::runtime_error(std::string("srps_init_depth_smooth_upsample: ") + srps_last_error(nullptr));
}

void preprocess_depth(const float* z0, int z0_h, int z0_w, int z0_n, int I_h, int I_w, std::vector<float>& zs,
                      std::vector<float>& z_full) {
    if (!Preferences::initOnHost) return preprocess_depth_device(z0, z0_h, z0_w, z0_n, I_h, I_w, zs, z_full);
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
