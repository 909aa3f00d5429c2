// jit.hpp — run-time compilation (NVRTC) and loading (driver API) of emitted programs.
#pragma once
#include <string>
#include <vector>

namespace rscm {

struct JitProgram {
    void *module[3] = {nullptr, nullptr, nullptr}; // one module per kernel variant, compiled on first use
    void *fn[3] = {nullptr, nullptr, nullptr};     // variants: (write), (logpost), (write + logpost)
};

// full translation unit for an emitted program body (embedded headers + `rscm_jit::Prog`)
std::string jit_source(const std::string &program_body);
// NVRTC -> sm_100a cubin of one kernel variant (+ its lowered name); disk-cached by source hash
bool jit_compile_cubin(const std::string &program_body, int dtype, int variant, std::string &cubin, std::string &lowered,
                       std::string &err);
bool jit_load(const std::string &cubin, const std::string &lowered, int variant, JitProgram &out, std::string &err);
void jit_unload(JitProgram &p);
// returns the CUresult of cuLaunchKernel (0 = success)
int jit_launch(const JitProgram &p, int variant, unsigned gx, unsigned gy, unsigned block, unsigned smem, void *stream, void *kargs);

} // namespace rscm
