// Command line of the reference (SRmeetsPS-GPU/Main.cpp:9-44) without cv::CommandLineParser: same keys, same
// defaults, `--key=value` / `-k=value` syntax, help printed (and exit code 0) when --dsloc is missing.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>

#include "../../include/srps_snapshot.h"
#include "SRPS.h"
#include "Utilities.h"

static void printMessage() {
    std::cout << "Usage: srps_cli [params]\n\n"
                 "\t-d, --dsloc\n\t\tpath to dataset mat file or folder containing images\n"
                 "\t-g, --device (value:0)\n\t\tcuda device to run the application on\n"
                 "\t-h, --help, --usage\n\t\tprint help\n"
                 "\t-t, --dstype (value:matlab)\n\t\tdataset type, can be matlab or images (extension: snapshot)\n"
                 "\t-x, --blockx (value:256)\n\t\tblock dimension x\n"
                 "\t-y, --blocky (value:4)\n\t\tblock dimension y\n"
                 "\textensions: --albedo=closed_form|reference_cg  --iters=K  --init-only  --dump-init=F.snap  --out=F.snap\n"
                 "\t            --init=device|host (depth pre-processing: CUDA kernels (default) or host cores, e.g. for --init-only without a GPU)\n"
                 "\t            --outdir=DIR (s/rho/z/N.mat + normals/albedo/depth.png)  --render=F.snap (files from a result snapshot, no GPU)\n";
}

int main(int argc, char* argv[]) {
    std::map<std::string, std::string> alias = {{"t", "dstype"}, {"d", "dsloc"}, {"g", "device"}, {"x", "blockx"},
                                                {"y", "blocky"}, {"h", "help"}, {"usage", "help"}};
    std::map<std::string, std::string> opt = {{"dstype", "matlab"}, {"device", "0"}, {"blockx", "256"}, {"blocky", "4"}};   // Main.cpp:10-17
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.empty() || a[0] != '-') continue;
        a = a.substr(a.find_first_not_of('-'));
        std::string key = a, val = "true";
        const size_t eq = a.find('=');
        if (eq != std::string::npos) { key = a.substr(0, eq); val = a.substr(eq + 1); }
        if (alias.count(key)) key = alias[key];
        opt[key] = val;
    }
    if (opt.count("render") && opt.count("outdir")) {        // extension: files from a result snapshot (--out), CPU only
        try {
            const srps::Snapshot r = srps::Snapshot::load(opt["render"]);
            const int32_t* hw = (const int32_t*)r.at("hw").raw.data();
            const size_t npix = r.at("z").count();
            save_results(opt["outdir"], hw[0], hw[1], r.at("mask").raw.data(), npix, (int)r.at("s").dims[0],
                         (const float*)r.at("z").raw.data(), (const float*)r.at("rho").raw.data(),
                         (const float*)r.at("N").raw.data(), (const float*)r.at("s").raw.data());
        } catch (const std::exception& e) {
            std::cerr << e.what() << std::endl;
            return 1;
        }
        return 0;
    }
    if (opt.count("help") || !opt.count("dsloc")) {          // Main.cpp:19-26
        printMessage();
        return 0;
    }
    Preferences::blockX = atoi(opt["blockx"].c_str());      // Main.cpp:27-29
    Preferences::blockY = atoi(opt["blocky"].c_str());
    Preferences::deviceId = atoi(opt["device"].c_str());
    if (opt.count("albedo")) Preferences::albedoMode = opt["albedo"] == "reference_cg" ? 1 : 0;
    if (opt.count("init")) Preferences::initOnHost = opt["init"] == "host" ? 1 : 0;
    auto configure = [&](SRPS& s) {
        s.init_only = opt.count("init-only") != 0;
        if (opt.count("dump-init")) s.dump_init = opt["dump-init"];
        if (opt.count("out")) s.dump_result = opt["out"];
        if (opt.count("outdir")) s.dump_dir = opt["outdir"];
        if (opt.count("iters")) s.fixed_iters = atoi(opt["iters"].c_str());
    };
    try {
        if (opt["dstype"] == "matlab") {                      // Main.cpp:31-36
            MatFileDataHandler dh;
            dh.loadDataFromMatFiles(opt["dsloc"].c_str());
            SRPS srps(dh);
            configure(srps);
            srps.execute();
        } else if (opt["dstype"] == "images") {               // Main.cpp:37-42
            ImageDataHandler dh;
            dh.loadDataFromImages(opt["dsloc"].c_str());
            SRPS srps(dh);
            configure(srps);
            srps.execute();
        } else if (opt["dstype"] == "snapshot") {             // extension
            SnapshotState st;
            st.load(opt["dsloc"]);
            SRPS srps(st);
            configure(srps);
            srps.execute();
        }                                                     // anything else: silently nothing (Main.cpp:43)
    } catch (const std::exception& e) {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 134;                                           // the reference lets the exception escape main
    }
    return 0;
}
