// mgpu_host.cpp -- a C++ host of the multi-GPU path that uses nothing but the C ABI (include/cutesdr_cuda.h): what a
// maintainer writes instead of bench.py's torch.distributed launcher. One process per GPU; rank, world size and the
// file that carries the 128-byte NCCL id come from the command line:
//     mgpu_host <rank> <world> <id_file> [n_channels] [n_blocks]
// Rank 0 creates the id and writes it to <id_file>; the other ranks wait for the file. Every rank then builds a bank for
// its channel slice and feeds blocks through cutesdr_bank_process_async_bcast (only rank 0 holds samples).
// Without a CUDA device the program reports the library's error and exits with status 3 -- the CPU test suite links it
// against libcutesdr_cuda.so and checks exactly that (tests/test_abi_surface.py); the GPU path is bench.py --gpus N.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <math.h>
#include <vector>

#include "cutesdr_cuda.h"

static int fail(const char* what)
{
    fprintf(stderr, "mgpu_host: %s: %s\n", what, cutesdr_last_error());
    return 3;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: mgpu_host <rank> <world> <id_file> [n_channels] [n_blocks]\n"); return 2; }
    const int rank = atoi(argv[1]), world = atoi(argv[2]);
    const char* id_file = argv[3];
    const int n_channels = argc > 4 ? atoi(argv[4]) : 1024;
    const int n_blocks = argc > 5 ? atoi(argv[5]) : 20;
    const double fs = 100147200.0;

    int n_dev = 0;
    if (cutesdr_device_count(&n_dev) < 0 || n_dev < 1) return fail("no CUDA device");

    unsigned char id[128];
    memset(id, 0, sizeof(id));
    if (world > 1) {
        if (rank == 0) {
            if (cutesdr_mgpu_unique_id(id) < 0) return fail("cutesdr_mgpu_unique_id");
            FILE* f = fopen(id_file, "wb");
            if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) { fprintf(stderr, "mgpu_host: cannot write %s\n", id_file); return 2; }
            fclose(f);
        } else {
            for (int tries = 0; tries < 600; tries++) {
                FILE* f = fopen(id_file, "rb");
                if (f) { const size_t n = fread(id, 1, sizeof(id), f); fclose(f); if (n == sizeof(id)) break; }
                usleep(100000);
            }
        }
    }
    cutesdr_mgpu* mg = nullptr;
    if (cutesdr_mgpu_init(&mg, id, rank, world, rank % n_dev) < 0) return fail("cutesdr_mgpu_init");
    int first = 0, count = 0;
    if (cutesdr_mgpu_channel_slice(n_channels, rank, world, &first, &count) < 0) return fail("cutesdr_mgpu_channel_slice");

    cutesdr_bank* bank = nullptr;
    if (cutesdr_bank_create(&bank, count, fs, rank % n_dev) < 0) return fail("cutesdr_bank_create");
    cutesdr_demod_info info;
    memset(&info, 0, sizeof(info));
    info.HiCut = 5000; info.HiCutmin = 5000; info.HiCutmax = 15000;
    info.LowCut = -5000; info.LowCutmin = -15000; info.LowCutmax = -5000;
    info.AgcThresh = -100; info.AgcManualGain = 30; info.AgcDecay = 200; info.AgcOn = 1;
    for (int c = 0; c < count; c++) {
        if (cutesdr_bank_set_demod(bank, c, CUTESDR_DEMOD_FM, &info) < 0) return fail("cutesdr_bank_set_demod");
        if (cutesdr_bank_set_demod_freq(bank, c, -(78125.0 * (first + c) - 40.0e6)) < 0) return fail("cutesdr_bank_set_demod_freq");
    }
    if (cutesdr_bank_set_audio_rate(bank, 48000.0) < 0) return fail("cutesdr_bank_set_audio_rate");
    int L = 0;
    if (cutesdr_bank_block_length(bank, &L) < 0) return fail("cutesdr_bank_block_length");

    // rank 0's samples: int16 I/Q pairs in pinned memory (a tone per block is enough for a host-side check)
    short* iq = nullptr;
    float* audio[2] = {nullptr, nullptr};
    const int stride = 2304;
    if (rank == 0 && cutesdr_host_alloc((void**)&iq, (size_t)L * 4) < 0) return fail("cutesdr_host_alloc");
    for (int k = 0; k < 2; k++)
        if (cutesdr_host_alloc((void**)&audio[k], (size_t)count * stride * sizeof(float)) < 0) return fail("cutesdr_host_alloc");
    if (rank == 0)
        for (int i = 0; i < L; i++) {
            iq[2 * i] = (short)lrint(8000.0 * cos(2.0 * M_PI * 0.01 * i));
            iq[2 * i + 1] = (short)lrint(8000.0 * sin(2.0 * M_PI * 0.01 * i));
        }
    std::vector<int> n_out(count);
    long long produced = 0;
    for (int k = 0; k < n_blocks; k++) {
        const int m = cutesdr_bank_process_async_bcast(bank, mg, L, rank == 0 ? iq : nullptr, CUTESDR_FMT_CS16, audio[k & 1], stride, n_out.data());
        if (m < 0) return fail("cutesdr_bank_process_async_bcast");
        produced += m;
    }
    if (cutesdr_bank_synchronize(bank) < 0) return fail("cutesdr_bank_synchronize");
    printf("rank %d/%d: channels %d..%d, %d blocks of %d samples, %lld audio samples per channel\n", rank, world, first, first + count - 1,
           n_blocks, L, produced);
    cutesdr_bank_destroy(bank);
    cutesdr_mgpu_destroy(mg);
    if (iq) cutesdr_host_free(iq);
    for (int k = 0; k < 2; k++) cutesdr_host_free(audio[k]);
    return 0;
}
