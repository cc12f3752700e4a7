// class SRPS with the reference's interface (SRmeetsPS-GPU/SRPS.h:10-18): non-owning DataHandler pointer,
// single-shot blocking execute().  The loop body runs in libsrps_b200.so through the C ABI.
#pragma once
#include <string>
#include <vector>

#include "Utilities.h"

class SRPS {
private:
    DataHandler* dh;
    SnapshotState* snap = nullptr;        // extension: start from a post-init snapshot instead of a DataHandler
public:
    SRPS(DataHandler& dh);
    explicit SRPS(SnapshotState& st);
    ~SRPS();
    void execute();

    // extensions (the reference only shows GUI windows and writes nothing on Linux)
    bool init_only = false;               // stop after the one-shot init
    std::string dump_init;                // write the post-init loop state (SRPSNAP1)
    std::string dump_result;              // write z, rho, N, s, the mask and the energies (SRPSNAP1)
    std::string dump_dir;                 // write s/rho/z/N.mat and normals/albedo/depth.png there (Output.cpp)
    int fixed_iters = 0;                  // > 0: ignore the stop rule
    std::vector<float> energies;          // per outer iteration, filled by execute()
};
