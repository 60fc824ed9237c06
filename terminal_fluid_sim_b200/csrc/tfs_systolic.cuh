// tfs_systolic.cuh -- temporally blocked, register-resident systolic wavefront projection.
// (placeholder interface; the kernel is added in a later commit)
#pragma once
#include "tfs_common.cuh"

namespace tfs {

constexpr int kMaxTemporalBlock = 32;

struct SystolicWorkspace {
    void *buf = nullptr;
    size_t bytes = 0;
};

struct SystolicArgs {
    float **u, **v, **p, **u2, **v2, **p2;  // pointers to the handle's buffers (ping-pong swaps them)
    const uint8_t *flags;
    long long W, H;
    int iterations, temporal_block;
    float omega, pc, gdt;
    bool fuse_gravity;
    cudaStream_t stream;
};

inline bool systolic_available() { return false; }
inline cudaError_t systolic_project(const SystolicArgs &, SystolicWorkspace &, uint64_t *, const char **msg) {
    if (msg) *msg = "not built";
    return cudaErrorNotSupported;
}
inline void systolic_release(SystolicWorkspace &ws) {
    if (ws.buf) cudaFree(ws.buf);
    ws.buf = nullptr;
    ws.bytes = 0;
}

}  // namespace tfs
