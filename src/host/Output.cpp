// Result output: what the reference shows in three windows and would have written through
// WRITE_MAT_FROM_DEVICE (compiled out on Linux, SRPS.cu:319-333), as files, without OpenCV / matio:
//   normals.png  N_as_opencv_mat    Utilities.cpp:277-298   albedo.png  rho_as_opencv_mat  Utilities.cpp:242-275
//   depth.png    z_as_opencv_mat    Utilities.cpp:300-320   s.mat rho.mat z.mat N.mat      SRPS.cu:329-332 (MAT v5)
// Same recipe, same bytes as srmeetsps-cuda_b200/output.py (tests/test_cpp_host.py compares the two).
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "Utilities.h"

namespace {

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s));
}

void png_chunk(FILE* f, const char tag[4], const std::vector<unsigned char>& data) {
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)data.size());
    fwrite(head.data(), 1, 4, f);
    fwrite(tag, 1, 4, f);
    if (!data.empty()) fwrite(data.data(), 1, data.size(), f);
    uLong crc = crc32(0L, (const Bytef*)tag, 4);
    if (!data.empty()) crc = crc32(crc, data.data(), (uInt)data.size());
    std::vector<unsigned char> tail;
    put_be32(tail, (uint32_t)crc);
    fwrite(tail.data(), 1, 4, f);
}

unsigned char to_u8(float v) {                       // round to nearest even, saturate (numpy rint / cv::saturate_cast)
    const float r = nearbyintf(v * 255.f);
    return (unsigned char)std::min(255.f, std::max(0.f, r));
}

// masked value k of pixel number m (column-major masked order) -> RGB image, row-major h x w x 3
template <class F>
std::vector<unsigned char> scatter_rgb(int h, int w, const unsigned char* mask, F&& pixel) {
    std::vector<unsigned char> img((size_t)h * w * 3, 0);
    size_t m = 0;
    for (int j = 0; j < w; j++)
        for (int i = 0; i < h; i++)
            if (mask[(size_t)i + (size_t)j * h]) {
                unsigned char* px = &img[((size_t)i * w + j) * 3];
                pixel(m, px);
                m++;
            }
    return img;
}

}  // namespace

void write_png_rgb8(const std::string& path, int w, int h, const unsigned char* rgb) {
    std::vector<unsigned char> raw((size_t)h * (1 + (size_t)w * 3));
    for (int y = 0; y < h; y++) {
        raw[(size_t)y * (1 + (size_t)w * 3)] = 0;                                    // filter type 0
        memcpy(&raw[(size_t)y * (1 + (size_t)w * 3) + 1], rgb + (size_t)y * w * 3, (size_t)w * 3);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("png: deflate failed");
    z.resize(zlen);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)w);
    put_be32(ihdr, (uint32_t)h);
    const unsigned char rest[5] = {8, 2, 0, 0, 0};                                   // 8 bit, RGB
    ihdr.insert(ihdr.end(), rest, rest + 5);
    png_chunk(f, "IHDR", ihdr);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    fclose(f);
}

// one single-precision column vector named "x" (the layout Mat_VarCreate("x", MAT_C_SINGLE, ...) of Utilities.cpp:46-63 gives)
void write_mat5_vector(const std::string& path, const float* x, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    char header[128];
    memset(header, ' ', 116);
    const char* text = "MATLAB 5.0 MAT-file, written by srps_cli";
    memcpy(header, text, strlen(text));
    memset(header + 116, 0, 8);
    header[124] = 0x00; header[125] = 0x01;                                          // version 0x0100, little endian
    header[126] = 'I'; header[127] = 'M';
    fwrite(header, 1, 128, f);
    auto element = [&](std::vector<unsigned char>& out, uint32_t type, const void* payload, uint32_t nbytes) {
        const uint32_t head[2] = {type, nbytes};
        const unsigned char* h8 = (const unsigned char*)head;
        out.insert(out.end(), h8, h8 + 8);
        const unsigned char* p = (const unsigned char*)payload;
        out.insert(out.end(), p, p + nbytes);
        out.insert(out.end(), (8 - nbytes % 8) % 8, 0);
    };
    std::vector<unsigned char> body;
    const uint32_t flags[2] = {7u /* mxSINGLE_CLASS */, 0u};
    const int32_t dims[2] = {(int32_t)n, 1};
    element(body, 6, flags, 8);
    element(body, 5, dims, 8);
    element(body, 1, "x", 1);
    element(body, 7, x, (uint32_t)(n * sizeof(float)));
    const uint32_t top[2] = {14u /* miMATRIX */, (uint32_t)body.size()};
    fwrite(top, 4, 2, f);
    fwrite(body.data(), 1, body.size(), f);
    fclose(f);
}

std::vector<unsigned char> render_normals(int h, int w, const unsigned char* mask, size_t npix, const float* N) {
    auto clip = [](float v) { return std::min(1.f, std::max(0.f, v)); };
    float lo = 0.f, hi = 0.f;                                                        // cv::normalize(MINMAX) sees the black background too
    bool any_bg = false;
    for (size_t p = 0; p < (size_t)h * w; p++) any_bg |= !mask[p];
    bool first = !any_bg;
    for (size_t m = 0; m < npix; m++) {
        const float v[3] = {clip(0.5f + 0.5f * N[m]), clip(0.5f + 0.5f * N[npix + m]), clip(0.5f - 0.5f * N[2 * npix + m])};
        for (float x : v) {
            if (first) { lo = hi = x; first = false; }
            lo = std::min(lo, x); hi = std::max(hi, x);
        }
    }
    const float span = hi - lo;
    auto img = scatter_rgb(h, w, mask, [&](size_t m, unsigned char* px) {
        const float v[3] = {clip(0.5f + 0.5f * N[m]), clip(0.5f + 0.5f * N[npix + m]), clip(0.5f - 0.5f * N[2 * npix + m])};
        for (int c = 0; c < 3; c++) px[c] = span > 0.f ? to_u8((v[c] - lo) / span) : 0;
    });
    return img;
}

std::vector<unsigned char> render_albedo(int h, int w, const unsigned char* mask, size_t npix, const float* rho) {
    float cap[3];
    for (int c = 0; c < 3; c++) {
        const float* r = rho + (size_t)c * npix;
        float sum = 0.f, sq = 0.f;
        for (size_t m = 0; m < npix; m++) { sum += r[m]; sq += r[m] * r[m]; }
        const float mean = sum / (float)npix;
        const float sd = std::sqrt(sq / (float)npix - mean * mean);
        std::vector<float> v(r, r + npix);
        std::sort(v.begin(), v.end());
        const float med = npix % 2 ? v[npix / 2] : 0.5f * (v[npix / 2 - 1] + v[npix / 2]);
        cap[c] = med + 5.f * sd;
    }
    return scatter_rgb(h, w, mask, [&](size_t m, unsigned char* px) {
        for (int c = 0; c < 3; c++) px[c] = to_u8(std::min(1.f, std::max(0.f, std::min(cap[c], rho[(size_t)c * npix + m]))));
    });
}

std::vector<unsigned char> render_depth(int h, int w, const unsigned char* mask, size_t npix, const float* z) {
    unsigned char lut[256][3];                                                       // MATLAB bone = (7 gray + fliplr(hot)) / 8
    const int m = 256, n1 = 3 * m / 8;
    for (int k = 0; k < m; k++) {
        const double g = k == m - 1 ? 1.0 : k * (1.0 / (m - 1));                     // numpy.linspace(0, 1, 256)[k], as output.py
        const double hot[3] = {k < n1 ? (k + 1.0) / n1 : 1.0,
                               k < n1 ? 0.0 : (k < 2 * n1 ? (k - n1 + 1.0) / n1 : 1.0),
                               k < 2 * n1 ? 0.0 : (k - 2 * n1 + 1.0) / (m - 2 * n1)};
        for (int c = 0; c < 3; c++) lut[k][c] = (unsigned char)nearbyint((7.0 * g + hot[2 - c]) / 8.0 * 255.0);
    }
    float lo = -z[0], hi = -z[0];
    for (size_t p = 0; p < npix; p++) { lo = std::min(lo, -z[p]); hi = std::max(hi, -z[p]); }
    const float span = hi - lo;
    return scatter_rgb(h, w, mask, [&](size_t p, unsigned char* px) {
        const unsigned char k = span > 0.f ? to_u8((-z[p] - lo) / span) : 0;
        for (int c = 0; c < 3; c++) px[c] = lut[k][c];
    });
}

void save_results(const std::string& dir, int h, int w, const unsigned char* mask, size_t npix, int n_images,
                  const float* z, const float* rho, const float* N, const float* s) {
    if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) throw std::runtime_error("cannot create " + dir);
    write_png_rgb8(dir + "/normals.png", w, h, render_normals(h, w, mask, npix, N).data());
    write_png_rgb8(dir + "/albedo.png", w, h, render_albedo(h, w, mask, npix, rho).data());
    write_png_rgb8(dir + "/depth.png", w, h, render_depth(h, w, mask, npix, z).data());
    write_mat5_vector(dir + "/s.mat", s, (size_t)n_images * 12);
    write_mat5_vector(dir + "/rho.mat", rho, 3 * npix);
    write_mat5_vector(dir + "/z.mat", z, npix);
    write_mat5_vector(dir + "/N.mat", N, 4 * npix);
}
