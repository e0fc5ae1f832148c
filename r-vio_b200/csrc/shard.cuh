// shard.cuh -- feature-sharded single stream: NCCL communicator + the two in-stream collectives (shard.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvio {

struct ShardComm { void* comm = nullptr; int rank = 0, world = 1; };

int shard_unique_id(void* id128);
int shard_comm_create(ShardComm* sc, int rank, int world, const void* id128, int device);
void shard_comm_destroy(ShardComm* sc);
int shard_allgather_lk(const ShardComm* sc, cudaStream_t s, float2* lk, float2* un, uint8_t* status, int S);
int shard_allreduce_terms(const ShardComm* sc, cudaStream_t s, double* red, int count);

}  // namespace rvio
