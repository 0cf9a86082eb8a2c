// NCCL communicator behind the C ABI. libnccl.so.2 is opened at run time (the same library torch.distributed already
// loaded when the host process is a torchrun rank), so libb2reg.so carries no link-time NCCL dependency and
// single-GPU users never need it. Used for the one exchange step of the sharded registrations: the all-reduce of the
// 6x6 normal equations per iteration.
#pragma once
#include "b2_common.cuh"

struct b2_comm_s {
    void* comm = nullptr;      // ncclComm_t
    int rank = 0, world = 1, device = 0;
};

namespace b2 {
// sum-all-reduce of n doubles, in place allowed, on stream s
int comm_allreduce_sum_f64(b2_comm_s* c, const double* d_send, double* d_recv, size_t n, cudaStream_t s);
// all-gather of `count` doubles per rank (rank r's block lands at d_recv + r * count; in place when d_send is that block)
int comm_allgather_f64(b2_comm_s* c, const double* d_send, double* d_recv, size_t count, cudaStream_t s);
}
