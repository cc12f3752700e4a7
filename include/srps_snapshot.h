// SRPSNAP1 named-array container (see srmeetsps-cuda_b200/snapshot.py for the layout).
// Header-only C++ reader/writer used by the host CLI (src/host) and the reference replay
// driver (oracle/ref/ref_replay.cu).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace srps {

struct SnapArray {
    int dtype = 0;  // 0=f32 1=i32 2=u8 3=f64
    int ndim = 0;
    int64_t dims[4] = {1, 1, 1, 1};
    std::vector<unsigned char> raw;
    size_t count() const { return (size_t)(dims[0] * dims[1] * dims[2] * dims[3]); }
    const float* f32() const { if (dtype != 0) throw std::runtime_error("snapshot: not float32"); return (const float*)raw.data(); }
    const int* i32() const { if (dtype != 1) throw std::runtime_error("snapshot: not int32"); return (const int*)raw.data(); }
    const unsigned char* u8() const { if (dtype != 2) throw std::runtime_error("snapshot: not uint8"); return raw.data(); }
    const double* f64() const { if (dtype != 3) throw std::runtime_error("snapshot: not float64"); return (const double*)raw.data(); }
};

static inline size_t snap_elem_size(int dtype) { return dtype == 2 ? 1 : (dtype == 3 ? 8 : 4); }

struct Snapshot {
    std::vector<std::string> order;
    std::map<std::string, SnapArray> arrays;

    const SnapArray& at(const std::string& name) const {
        auto it = arrays.find(name);
        if (it == arrays.end()) throw std::runtime_error("snapshot: missing array '" + name + "'");
        return it->second;
    }
    bool has(const std::string& name) const { return arrays.count(name) != 0; }

    void put(const std::string& name, int dtype, std::initializer_list<int64_t> dims, const void* data) {
        SnapArray a;
        a.dtype = dtype;
        a.ndim = (int)dims.size();
        int k = 0;
        for (auto d : dims) a.dims[k++] = d;
        size_t nb = a.count() * snap_elem_size(dtype);
        a.raw.resize(nb);
        if (nb) memcpy(a.raw.data(), data, nb);
        if (!arrays.count(name)) order.push_back(name);
        arrays[name] = std::move(a);
    }

    static Snapshot load(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("snapshot: cannot open " + path);
        Snapshot s;
        char magic[8];
        int32_t count = 0;
        if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SRPSNAP1", 8) != 0 || fread(&count, 4, 1, f) != 1) {
            fclose(f);
            throw std::runtime_error("snapshot: bad header in " + path);
        }
        for (int i = 0; i < count; i++) {
            char name[25] = {0};
            int32_t hdr[2];
            SnapArray a;
            if (fread(name, 1, 24, f) != 24 || fread(hdr, 4, 2, f) != 2 || fread(a.dims, 8, 4, f) != 4) {
                fclose(f);
                throw std::runtime_error("snapshot: truncated entry in " + path);
            }
            a.dtype = hdr[0];
            a.ndim = hdr[1];
            size_t nb = a.count() * snap_elem_size(a.dtype);
            a.raw.resize(nb);
            if (nb && fread(a.raw.data(), 1, nb, f) != nb) {
                fclose(f);
                throw std::runtime_error("snapshot: truncated data in " + path);
            }
            size_t pad = (8 - nb % 8) % 8;
            if (pad) fseek(f, (long)pad, SEEK_CUR);
            s.order.push_back(name);
            s.arrays[name] = std::move(a);
        }
        fclose(f);
        return s;
    }

    void save(const std::string& path) const {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("snapshot: cannot write " + path);
        int32_t count = (int32_t)order.size();
        fwrite("SRPSNAP1", 1, 8, f);
        fwrite(&count, 4, 1, f);
        const char zeros[8] = {0};
        for (auto& n : order) {
            const SnapArray& a = arrays.at(n);
            char name[24] = {0};
            strncpy(name, n.c_str(), 23);
            int32_t hdr[2] = {a.dtype, a.ndim};
            fwrite(name, 1, 24, f);
            fwrite(hdr, 4, 2, f);
            fwrite(a.dims, 8, 4, f);
            if (!a.raw.empty()) fwrite(a.raw.data(), 1, a.raw.size(), f);
            size_t pad = (8 - a.raw.size() % 8) % 8;
            if (pad) fwrite(zeros, 1, pad, f);
        }
        fclose(f);
    }
};

}  // namespace srps
