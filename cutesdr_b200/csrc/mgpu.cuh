// mgpu.cuh -- multi-GPU plumbing of libcutesdr_cuda: one process per GPU, channels shard across ranks, and the ONE
// exchange the path has -- the wideband IQ block that rank 0 received from the host -- is an NCCL broadcast over
// NVLink / NVSwitch. NCCL is resolved at run time (dlopen of libnccl.so.2): a single-GPU host
// needs no NCCL at all.
#pragma once
#include "common.cuh"

struct cutesdr_mgpu {
    void* comm = nullptr;          // ncclComm_t
    int rank = 0, world = 1, device = 0;
    cudaStream_t st_comm = 0;      // NCCL broadcasts
    std::vector<cudaEvent_t> ev_chunk;
    size_t chunk_bytes = 4 << 20;  // measured (2 x B200): 1 MiB chunks leave the broadcast latency-bound at 26 GB/s; CUTESDR_BCAST_CHUNK_KB overrides
    long long blocks = 0, bytes_bcast = 0;
    std::mutex mu;
    ~cutesdr_mgpu();
};

namespace csdr {
// Broadcast `bytes` of rank 0's device buffer `d_buf` (same pointer role on every rank: the local landing buffer).
// Rank 0: `h_src` (pinned host) is copied in chunks on st_copy, chunk k's broadcast (on the comm stream) waits only
// for chunk k's copy, so the H2D of chunk k+1 overlaps the broadcast of chunk k. Other ranks: h_src is ignored.
// On return, `done` has been recorded on the comm stream after the last chunk landed.
int mgpu_bcast_block(cutesdr_mgpu* m, const void* h_src, void* d_buf, size_t bytes, cudaStream_t st_copy, cudaEvent_t done);
}  // namespace csdr
