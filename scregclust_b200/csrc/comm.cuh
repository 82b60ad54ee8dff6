// Exchange layer of a multi-GPU job: one Comm per rank (one rank per GPU; ranks are processes under
// torchrun / MPI-style launchers, or host threads of one process in scrc_grid_fit_multi).
//
// The hot path needs three collectives, all tiny next to the compute:
//   * broadcast of the session precompute (Gram, eigenbasis, initial fit) from rank 0,
//   * an in-place all-reduce(SUM) that merges per-cycle outputs in which every entry is owned by
//     exactly one rank and is zero elsewhere -- module summaries after Step 1, the per-gene
//     assignment / R2 slices after Step 2 (x + 0 + ... + 0 is exact, so the merge is bit-preserving),
//   * a broadcast-per-rank of contiguous coefficient column slices (an all-gather with ragged counts).
// Implementation: NCCL over NVLink 5 / NVSwitch, resolved at run time with dlopen("libnccl.so.2") so that
// single-GPU users need no NCCL installation and a process that already loaded NCCL (PyTorch) shares it.
#pragma once
#include "common.cuh"

namespace scrc {

struct Comm {
    int rank = 0, world = 1, device = 0;
    void* nccl = nullptr;               // ncclComm_t
    ~Comm();
    void init(const char* unique_id_128, int world, int rank, int device);
    void allreduce_sum_f64(double* p, size_t n, cudaStream_t st);
    void allreduce_sum_i32(int* p, size_t n, cudaStream_t st);
    void allreduce_sum_u8(unsigned char* p, size_t n, cudaStream_t st);
    void broadcast_bytes(void* p, size_t bytes, int root, cudaStream_t st);
    void group_begin();
    void group_end();
};

void comm_unique_id(char* out_128);

}  // namespace scrc
