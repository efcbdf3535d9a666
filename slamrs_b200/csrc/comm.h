// Thin NCCL binding for the one real exchange step of the path: the all-gather of per-particle
// results before resampling (plus a broadcast of the exported map and a stream-ordered barrier).
// NCCL is resolved with dlopen at first use so that single-GPU users need no NCCL at all and a
// host process that already loaded a libnccl.so.2 (e.g. PyTorch's) shares it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string>

namespace slamrs {

struct Comm;  // opaque

// 128-byte unique id for rank 0 to distribute
int comm_unique_id(uint8_t out[128], std::string* err);
// collective; returns nullptr and fills *err on failure
Comm* comm_create(const uint8_t id[128], int rank, int world, std::string* err);
void comm_destroy(Comm* c);

int comm_all_gather(Comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s, std::string* err);
int comm_broadcast(Comm* c, void* buf, size_t bytes, int root, cudaStream_t s, std::string* err);
// stream-ordered barrier across all ranks (a 1-element all-reduce on `scratch`, device memory)
int comm_barrier(Comm* c, int* scratch, cudaStream_t s, std::string* err);

}  // namespace slamrs
