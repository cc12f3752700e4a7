// Host-side data model of the reference, kept source-compatible for the hot path's callers:
//   Preferences            Utilities.h:224-230, Main.cpp:5-7
//   DataHandler            Utilities.h:166-181   (same fields, same column-major layouts, same ownership)
//   MatFileDataHandler     Utilities.h:183-188, Utilities.cpp:159-199  (MAT v5, variables I,K,mask,sf,z0)
//   ImageDataHandler       Utilities.h:190-192, Utilities.cpp:349-395  (RGB/*.png, Depth/*.png, mask.png, K.txt)
//   Timer                  Utilities.h:194-222   (device-synchronised wall time instead of CPU clock())
// No OpenCV / matio: PNG and MAT5 are decoded by the small readers in Loaders.cpp (zlib only).
#pragma once
#include <chrono>
#include <cstdint>
#include <string>
#include <vector>

struct Preferences {
    static int blockX;        // accepted for CLI compatibility (launch shapes are fixed per kernel)
    static int blockY;
    static int deviceId;
    static int albedoMode;    // extension: 0 closed form (default), 1 the reference's diagonal CG
    static int maxOuter;      // extension: 0 -> the reference's MAX_ITERATIONS = 10
    static int initOnHost;    // extension (--init=host): depth pre-processing on the host cores instead of the device
                              // (tooling without a GPU: --init-only dumps); default 0 = device kernels, Telea on the host
private:
    Preferences() {}
};

struct DataHandler {
    float* I;                 // h x w x c x n, column-major            Utilities.h:168
    int I_w, I_h, I_c, I_n;
    int z0_w, z0_h;
    float* K;                 // 3 x 3 column-major: K[0]=fx K[4]=fy K[6]=cx K[7]=cy
    float* mask;              // h x w, {0,1}
    float sf;
    float* z0;                // (h/sf) x (w/sf) x z0_n
    int z0_n;
    DataHandler();
    ~DataHandler();
    void freeMemory();
};

struct MatFileDataHandler : public DataHandler {
    void loadDataFromMatFiles(const char* filename);
};

struct ImageDataHandler : public DataHandler {
    void loadDataFromImages(const char* dataFolder);
};

// Extension: a post-init loop state (SRPSNAP1, include/srps_snapshot.h) written by `--init-only --dump=...`
// or by python; lets the loop run without the depth pre-processing.
struct SnapshotState {
    int h = 0, w = 0, sf = 1, n = 0, c = 3;
    std::vector<float> K;             // 9
    std::vector<unsigned char> mask;  // h*w column-major
    std::vector<float> I, z, z0s;     // masked layouts: I[n][c][npix], z[npix], z0s[npixs]
    void load(const std::string& path);
    void save(const std::string& path) const;
};

class Timer {                          // Utilities.h:194-222
public:
    void start() { t0 = std::chrono::steady_clock::now(); running = true; }
    void end() { if (running) { sec = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count(); running = false; } }
    float get() { if (running) end(); return sec; }
private:
    std::chrono::steady_clock::time_point t0;
    bool running = false;
    float sec = 0.f;
};

// ---- loaders (Loaders.cpp) ------------------------------------------------------------------------
struct PngImage { int w = 0, h = 0, channels = 0, bits = 0; std::vector<uint16_t> px; };   // row-major, interleaved
PngImage read_png(const std::string& path);
std::vector<std::string> list_sorted(const std::string& dir);                                // cv::glob order

// ---- one-shot depth pre-processing (Preprocess.cpp), SRPS.cu:117-149 + devicecalls.cu:95-125 -------
// zs: inpainted + smoothed LR depth (column-major hs x ws), z_full: bicubic upsample (column-major h x w).
void preprocess_depth(const float* z0, int z0_h, int z0_w, int z0_n, int I_h, int I_w,
                      std::vector<float>& zs, std::vector<float>& z_full);

// ---- result output (Output.cpp), SRPS.cu:319-333 + Utilities.cpp:242-320 without OpenCV / matio ----------
// mask: h*w column-major {0,1}; z[npix], rho[3][npix], N[4][npix], s[n][3][4] in the reference's masked layouts.
void write_png_rgb8(const std::string& path, int w, int h, const unsigned char* rgb);
void write_mat5_vector(const std::string& path, const float* x, size_t n);
std::vector<unsigned char> render_normals(int h, int w, const unsigned char* mask, size_t npix, const float* N);
std::vector<unsigned char> render_albedo(int h, int w, const unsigned char* mask, size_t npix, const float* rho);
std::vector<unsigned char> render_depth(int h, int w, const unsigned char* mask, size_t npix, const float* z);
void save_results(const std::string& dir, int h, int w, const unsigned char* mask, size_t npix, int n_images,
                  const float* z, const float* rho, const float* N, const float* s);
