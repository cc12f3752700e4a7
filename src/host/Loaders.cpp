// Loaders of the reference's two dataset types without OpenCV / matio (zlib only).
//   ImageDataHandler::loadDataFromImages   restates Utilities.cpp:322-395
//   MatFileDataHandler::loadDataFromMatFiles restates Utilities.cpp:124-199 (MAT v5 container)
#include <dirent.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "../../include/srps_snapshot.h"
#include "Utilities.h"

int Preferences::blockX = 256;       // Main.cpp:5
int Preferences::blockY = 4;         // Main.cpp:6
int Preferences::deviceId = 0;       // Main.cpp:7
int Preferences::albedoMode = 0;
int Preferences::maxOuter = 0;
int Preferences::initOnHost = 0;

DataHandler::DataHandler() : I(NULL), K(NULL), mask(NULL), z0(NULL) {}       // Utilities.cpp:142
DataHandler::~DataHandler() { freeMemory(); }
void DataHandler::freeMemory() {                                               // Utilities.cpp:148-157
    delete[] I; delete[] K; delete[] mask; delete[] z0;
    I = K = mask = z0 = NULL;
}

// ------------------------------------------------------------------------------------------------
static std::vector<unsigned char> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static std::vector<unsigned char> inflate_all(const unsigned char* src, size_t n, size_t hint) {
    std::vector<unsigned char> out(std::max<size_t>(hint, 1 << 16));
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) throw std::runtime_error("inflateInit failed");
    zs.next_in = const_cast<unsigned char*>(src);
    zs.avail_in = (uInt)n;
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
        if (zs.total_out >= out.size()) out.resize(out.size() * 2);
        zs.next_out = out.data() + zs.total_out;
        zs.avail_out = (uInt)(out.size() - zs.total_out);
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); throw std::runtime_error("zlib inflate error"); }
    }
    out.resize(zs.total_out);
    inflateEnd(&zs);
    return out;
}

std::vector<std::string> list_sorted(const std::string& dir) {
    std::vector<std::string> out;
    DIR* d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("cannot list " + dir);
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        out.push_back(dir + "/" + n);
    }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

// ---- PNG: non-interlaced, colour types 0/2/4/6, 8 or 16 bit (what the reference's datasets contain) ----
static inline uint32_t be32(const unsigned char* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// 8/16-bit gray, gray+alpha, RGB, RGBA; 1/2/4/8-bit gray and palette; Adam7 interlacing.  Like cv::imread: alpha is
// dropped, palettes are expanded to RGB, gray samples below 8 bits are scaled to the full 8-bit range.
PngImage read_png(const std::string& path) {
    std::vector<unsigned char> f = read_file(path);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (f.size() < 8 || memcmp(f.data(), sig, 8) != 0) throw std::runtime_error(path + ": not a PNG file");
    PngImage img;
    int ctype = -1, interlace = 0;
    std::vector<unsigned char> idat, plte;
    size_t pos = 8;
    while (pos + 8 <= f.size()) {
        uint32_t len = be32(&f[pos]);
        std::string type((const char*)&f[pos + 4], 4);
        const unsigned char* data = &f[pos + 8];
        if (pos + 12 + len > f.size()) throw std::runtime_error(path + ": truncated PNG chunk");
        if (type == "IHDR") {
            if (len < 13) throw std::runtime_error(path + ": bad PNG header");
            const uint32_t w32 = be32(data), h32 = be32(data + 4);
            // a dataset image is at most a few thousand pixels wide: anything else is a corrupt header, not an allocation request
            if (w32 == 0 || h32 == 0 || w32 > 65535u || h32 > 65535u) throw std::runtime_error(path + ": implausible PNG size");
            img.w = (int)w32; img.h = (int)h32; img.bits = data[8]; ctype = data[9]; interlace = data[12];
        } else if (type == "PLTE") {
            plte.assign(data, data + len);
        } else if (type == "IDAT") {
            idat.insert(idat.end(), data, data + len);
        } else if (type == "IEND") {
            break;
        }
        pos += 12 + len;
    }
    if (ctype < 0) throw std::runtime_error(path + ": PNG without a header chunk");
    const int bits = img.bits;
    const bool low = bits == 1 || bits == 2 || bits == 4;
    const bool ok = (ctype == 0 && (low || bits == 8 || bits == 16)) || (ctype == 3 && (low || bits == 8) && !plte.empty()) ||
                    ((ctype == 2 || ctype == 4 || ctype == 6) && (bits == 8 || bits == 16));
    if (!ok || interlace > 1) throw std::runtime_error(path + ": unsupported PNG flavour");
    const int src_ch = (ctype == 0 || ctype == 3) ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : 4;
    img.channels = (ctype == 0 || ctype == 4) ? 1 : 3;            // alpha dropped, palette expanded
    if (ctype == 3) img.bits = 8;
    else if (low) img.bits = 8;
    const int bpp = std::max(1, src_ch * bits / 8);              // filter distance in bytes
    auto line_bytes = [&](int wpx) { return ((size_t)wpx * src_ch * bits + 7) / 8; };

    // pass geometry: one pass for a plain image, seven for Adam7
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole[1] = {{0, 0, 1, 1}};
    const Pass* passes = interlace ? adam7 : whole;
    const int npass = interlace ? 7 : 1;
    size_t need = 0;
    for (int p = 0; p < npass; p++) {
        const int pw = (img.w - passes[p].x0 + passes[p].dx - 1) / passes[p].dx, ph = (img.h - passes[p].y0 + passes[p].dy - 1) / passes[p].dy;
        if (pw > 0 && ph > 0) need += (line_bytes(pw) + 1) * ph;
    }
    std::vector<unsigned char> raw = inflate_all(idat.data(), idat.size(), need);
    if (raw.size() < need) throw std::runtime_error(path + ": short PNG data");
    img.px.assign((size_t)img.w * img.h * img.channels, 0);
    size_t off = 0;
    for (int p = 0; p < npass; p++) {
        const Pass& ps = passes[p];
        const int pw = (img.w - ps.x0 + ps.dx - 1) / ps.dx, ph = (img.h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = line_bytes(pw);
        std::vector<unsigned char> prev(stride, 0), cur(stride);
        for (int y = 0; y < ph; y++) {
            const unsigned char* line = &raw[off];
            off += stride + 1;
            const int ft = line[0];
            for (size_t i = 0; i < stride; i++) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                int pr = 0;
                switch (ft) {
                    case 0: pr = 0; break;
                    case 1: pr = a; break;
                    case 2: pr = b; break;
                    case 3: pr = (a + b) / 2; break;
                    case 4: { int q = a + b - c, pa = std::abs(q - a), pb = std::abs(q - b), pc = std::abs(q - c);
                              pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                    default: throw std::runtime_error(path + ": bad PNG filter");
                }
                cur[i] = (unsigned char)(line[1 + i] + pr);
            }
            const int oy = ps.y0 + y * ps.dy;
            for (int x = 0; x < pw; x++) {
                const int ox = ps.x0 + x * ps.dx;
                uint16_t* dst = &img.px[((size_t)oy * img.w + ox) * img.channels];
                if (low || ctype == 3) {
                    const size_t bit = (size_t)x * bits;
                    const int v = bits == 8 ? cur[x] : (cur[bit / 8] >> (8 - bits - (int)(bit % 8))) & ((1 << bits) - 1);
                    if (ctype == 3) {
                        if ((size_t)v * 3 + 2 >= plte.size()) throw std::runtime_error(path + ": palette index out of range");
                        for (int ch = 0; ch < 3; ch++) dst[ch] = plte[(size_t)v * 3 + ch];
                    } else {
                        dst[0] = (uint16_t)(v * 255 / ((1 << bits) - 1));
                    }
                } else {
                    for (int ch = 0; ch < img.channels; ch++) {
                        const unsigned char* q = &cur[(size_t)x * bpp + (size_t)ch * bits / 8];
                        dst[ch] = bits == 8 ? q[0] : (uint16_t)(q[0] << 8 | q[1]);
                    }
                }
            }
            prev.swap(cur);
        }
    }
    return img;
}

// ---- image folder: Utilities.cpp:349-395 ----------------------------------------------------------------
void ImageDataHandler::loadDataFromImages(const char* dataFolder) {
    freeMemory();
    const std::string root(dataFolder);
    std::vector<std::string> files = list_sorted(root + "/RGB");
    if (files.empty()) throw std::runtime_error("no images in " + root + "/RGB");
    auto read_color = [](const std::string& file) {          // cv::imread(file): always 3 channels, 8 bit
        PngImage im = read_png(file);
        if (im.bits == 16) { for (auto& v : im.px) v >>= 8; im.bits = 8; }
        if (im.channels == 1) {
            PngImage c = im;
            c.channels = 3;
            c.px.resize((size_t)im.w * im.h * 3);
            for (size_t p = 0; p < (size_t)im.w * im.h; p++) c.px[3 * p] = c.px[3 * p + 1] = c.px[3 * p + 2] = im.px[p];
            return c;
        }
        return im;
    };
    PngImage first = read_color(files[0]);
    I_n = (int)files.size(); I_w = first.w; I_h = first.h; I_c = first.channels;
    I = new float[(size_t)I_h * I_w * I_c * I_n];
    for (int n = 0; n < I_n; n++) {
        PngImage im = n == 0 ? first : read_color(files[n]);
        if (im.w != I_w || im.h != I_h || im.channels != I_c) throw std::runtime_error(files[n] + ": size mismatch");
        float* dst = I + (size_t)n * I_w * I_h * I_c;
        // cv::imread gives BGR and the reference stores channel (channels-1-c) (Utilities.cpp:343): plane 0 = R
        for (int c = 0; c < I_c; c++)
            for (int i = 0; i < I_h; i++)
                for (int j = 0; j < I_w; j++)
                    dst[(size_t)i + (size_t)j * I_h + (size_t)c * I_h * I_w] = im.px[((size_t)i * I_w + j) * I_c + c] / 255.f;
    }
    std::ifstream fk(root + "/K.txt");
    if (!fk) throw std::runtime_error("cannot open " + root + "/K.txt");
    std::string line, val;
    K = new float[9];
    for (int i = 0; i < 3; i++) {                              // Utilities.cpp:364-373
        std::getline(fk, line);
        std::istringstream tok(line);
        for (int j = 0; j < 3; j++) { std::getline(tok, val, ','); K[i + 3 * j] = std::stof(val); }
    }
    std::getline(fk, line);
    std::istringstream tok(line);
    std::getline(tok, val, ','); sf = std::stof(val);
    std::getline(tok, val, ','); const float min_z = std::stof(val);
    std::getline(tok, val);      const float max_z = std::stof(val);
    PngImage m = read_png(root + "/mask.png");
    if (m.w != I_w || m.h != I_h) throw std::runtime_error("mask.png: size mismatch");
    mask = new float[(size_t)I_h * I_w];
    for (int i = 0; i < I_h; i++)
        for (int j = 0; j < I_w; j++) {                        // IMREAD_GRAYSCALE: a colour mask is converted (BT.601 weights)
            const uint16_t* px = &m.px[((size_t)i * I_w + j) * m.channels];
            int g = m.channels == 1 ? px[0] : (px[0] * 4899 + px[1] * 9617 + px[2] * 1868 + 8192) >> 14;
            if (m.bits == 16) g >>= 8;
            mask[(size_t)i + (size_t)j * I_h] = g / 255.f;
        }
    files = list_sorted(root + "/Depth");
    z0_n = (int)files.size();
    z0_h = (int)(I_h / sf); z0_w = (int)(I_w / sf);
    z0 = new float[(size_t)z0_h * z0_w * z0_n];
    for (int n = 0; n < z0_n; n++) {
        PngImage d = read_png(files[n]);
        if (d.w != z0_w || d.h != z0_h) throw std::runtime_error(files[n] + ": size mismatch");
        float* dst = z0 + (size_t)n * z0_h * z0_w;
        const float q = d.bits == 16 ? 65535.f : 255.f;
        for (int i = 0; i < z0_h; i++)
            for (int j = 0; j < z0_w; j++)                     // min + v/65535*(max-min)   Utilities.cpp:330
                dst[(size_t)i + (size_t)j * z0_h] = min_z + (d.px[((size_t)i * z0_w + j) * d.channels] / q) * (max_z - min_z);
    }
}

// ---- MAT v5 -----------------------------------------------------------------------------------------------
namespace {
struct MatVar { std::vector<int> dims; std::vector<double> data; };

struct Cursor {
    const unsigned char* p; size_t n, pos = 0;
    // One data element: its payload [data, data + nbytes) is guaranteed to lie inside the buffer on success (a
    // truncated or corrupt file fails here instead of being read past its end).
    bool tag(uint32_t& type, uint32_t& nbytes, const unsigned char*& data) {
        if (pos + 8 > n) return false;
        uint32_t t; memcpy(&t, p + pos, 4);
        if (t >> 16) {                                          // small data element: <= 4 payload bytes inside the tag
            type = t & 0xffff; nbytes = t >> 16; data = p + pos + 4; pos += 8;
            return nbytes <= 4;
        }
        type = t; memcpy(&nbytes, p + pos + 4, 4); data = p + pos + 8;
        if ((size_t)nbytes > n - pos - 8) return false;         // payload must fit
        const size_t adv = 8 + (type == 15 ? (size_t)nbytes : (((size_t)nbytes + 7) / 8) * 8);   // miCOMPRESSED elements are not padded
        pos = std::min(n, pos + adv);                           // the padding of the last element may be cut off
        return true;
    }
};

template <typename T> void append_as_double(std::vector<double>& out, const unsigned char* d, uint32_t nbytes) {
    const size_t cnt = nbytes / sizeof(T);
    for (size_t i = 0; i < cnt; i++) { T v; memcpy(&v, d + i * sizeof(T), sizeof(T)); out.push_back((double)v); }
}

bool parse_matrix(const unsigned char* p, size_t n, std::string& name, MatVar& var) {
    Cursor c{p, n};
    uint32_t type, nb; const unsigned char* d;
    if (!c.tag(type, nb, d)) return false;                      // array flags
    if (!c.tag(type, nb, d)) return false;                      // dimensions
    var.dims.clear();
    for (uint32_t i = 0; i < nb / 4; i++) { int32_t v; memcpy(&v, d + 4 * i, 4); var.dims.push_back(v); }
    if (!c.tag(type, nb, d)) return false;                      // name
    name.assign((const char*)d, nb);
    if (!c.tag(type, nb, d)) return false;                      // real part
    var.data.clear();
    size_t expect = 1;
    for (int v : var.dims) { if (v < 0 || (v > 0 && expect > (size_t)1 << 40)) return false; expect *= (size_t)v; }
    switch (type) {
        case 1: append_as_double<int8_t>(var.data, d, nb); break;
        case 2: append_as_double<uint8_t>(var.data, d, nb); break;
        case 3: append_as_double<int16_t>(var.data, d, nb); break;
        case 4: append_as_double<uint16_t>(var.data, d, nb); break;
        case 5: append_as_double<int32_t>(var.data, d, nb); break;
        case 6: append_as_double<uint32_t>(var.data, d, nb); break;
        case 7: append_as_double<float>(var.data, d, nb); break;
        case 9: append_as_double<double>(var.data, d, nb); break;
        default: return false;
    }
    return var.data.size() == expect;                           // dims and payload must agree
}
}  // namespace

void MatFileDataHandler::loadDataFromMatFiles(const char* filename) {
    freeMemory();
    std::vector<unsigned char> f;
    try { f = read_file(filename); } catch (...) {
        fprintf(stderr, "Error opening MAT file \"%s\"!\n", filename);                 // Utilities.cpp:165-168
        throw std::runtime_error("Failed opening MAT file");
    }
    if (f.size() < 128 || memcmp(f.data(), "MATLAB 5.0", 10) != 0) throw std::runtime_error("Failed opening MAT file");
    std::map<std::string, MatVar> vars;
    Cursor c{f.data() + 128, f.size() - 128};
    uint32_t type, nb; const unsigned char* d;
    while (c.pos + 8 <= c.n && c.tag(type, nb, d)) {
        std::vector<unsigned char> tmp;
        const unsigned char* body = d; size_t blen = nb;
        if (type == 15) {                                       // miCOMPRESSED
            try { tmp = inflate_all(d, nb, (size_t)nb * 4); } catch (...) { continue; }
            if (tmp.size() < 8) continue;
            uint32_t t2, n2; memcpy(&t2, tmp.data(), 4); memcpy(&n2, tmp.data() + 4, 4);
            if (t2 != 14 || (size_t)n2 > tmp.size() - 8) continue;           // the inner length must fit the inflated data
            body = tmp.data() + 8; blen = n2;
        } else if (type != 14) {
            continue;
        }
        std::string name; MatVar v;
        if (parse_matrix(body, blen, name, v)) vars[name] = std::move(v);
    }
    auto need = [&](const char* nm) -> MatVar& {
        auto it = vars.find(nm);
        if (it == vars.end()) {
            fprintf(stderr, "Variable not found, or error reading MAT file\n");      // Utilities.cpp:37-40
            throw std::runtime_error("Failed reading MAT file");
        }
        return it->second;
    };
    auto bad = [](const char* what) -> void {
        fprintf(stderr, "Variable not found, or error reading MAT file\n");          // Utilities.cpp:37-40
        throw std::runtime_error(std::string("Failed reading MAT file: ") + what);
    };
    MatVar& vI = need("I");
    if (vI.dims.size() < 4) throw std::runtime_error("MAT variable I must be h x w x c x n");
    I_h = vI.dims[0]; I_w = vI.dims[1]; I_c = vI.dims[2]; I_n = vI.dims[3];            // Utilities.cpp:171
    if (I_h < 1 || I_w < 1 || I_c < 1 || I_n < 1 || vI.data.size() != (size_t)I_h * I_w * I_c * I_n) bad("I: dims do not match its data");
    I = new float[vI.data.size()];
    for (size_t i = 0; i < vI.data.size(); i++) I[i] = (float)vI.data[i];
    MatVar& vK = need("K");
    if (vK.data.size() < 9) bad("K must be 3 x 3");
    K = new float[9];
    for (int i = 0; i < 9; i++) K[i] = (float)vK.data[i];
    MatVar& vm = need("mask");
    if (vm.data.size() != (size_t)I_h * I_w) bad("mask must be h x w");
    mask = new float[(size_t)I_h * I_w];
    for (size_t i = 0; i < (size_t)I_h * I_w; i++) mask[i] = (float)vm.data[i];
    MatVar& vsf = need("sf");
    if (vsf.data.empty() || !(vsf.data[0] >= 1.0)) bad("sf must be a scalar >= 1");
    sf = (float)vsf.data[0];
    MatVar& vz = need("z0");
    if (vz.dims.size() < 2) bad("z0 must be at least 2-D");
    z0_n = vz.dims.size() > 2 ? vz.dims[2] : 1;                                          // Utilities.cpp:188-190
    z0_h = vz.dims[0]; z0_w = vz.dims[1];
    if (z0_h < 1 || z0_w < 1 || z0_n < 1 || vz.data.size() != (size_t)z0_h * z0_w * z0_n) bad("z0: dims do not match its data");
    z0 = new float[vz.data.size()];
    for (size_t i = 0; i < vz.data.size(); i++) z0[i] = (float)vz.data[i];
}

// ---- post-init snapshot -----------------------------------------------------------------------------------
void SnapshotState::load(const std::string& path) {
    srps::Snapshot s = srps::Snapshot::load(path);
    const int* d = s.at("dims").i32();
    h = d[0]; w = d[1]; sf = d[2];
    const srps::SnapArray& aI = s.at("I");
    n = (int)aI.dims[0]; c = (int)aI.dims[1];
    I.assign(aI.f32(), aI.f32() + aI.count());
    z.assign(s.at("z").f32(), s.at("z").f32() + s.at("z").count());
    z0s.assign(s.at("z0s").f32(), s.at("z0s").f32() + s.at("z0s").count());
    K.assign(s.at("K").f32(), s.at("K").f32() + 9);
    mask.assign(s.at("mask").u8(), s.at("mask").u8() + s.at("mask").count());
}

void SnapshotState::save(const std::string& path) const {
    srps::Snapshot s;
    const int d[3] = {h, w, sf};
    s.put("dims", 1, {3}, d);
    s.put("K", 0, {9}, K.data());
    s.put("mask", 2, {(int64_t)mask.size()}, mask.data());
    s.put("I", 0, {n, c, (int64_t)z.size()}, I.data());
    s.put("z", 0, {(int64_t)z.size()}, z.data());
    s.put("z0s", 0, {(int64_t)z0s.size()}, z0s.data());
    s.save(path);
}
