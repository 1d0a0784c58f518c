// GPU BVH8 builder and on-disk BVH cache (bvh_build.cu); host-side interface used by api.cu.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "bvh8.h"

namespace igb {

constexpr int BVH_BUILDER_HOST_SAH = 0, BVH_BUILDER_GPU_LBVH = 1;

// LBVH on the device collapsed to BVH8 with leaves of <= 4 primitives; same output as build_bvh8(boxes, 4). boxes.size() >= 5.
bool build_bvh8_gpu(const std::vector<Box3>& boxes, Bvh8& out, cudaStream_t stream, std::string& err);

uint64_t bvh_cache_hash(const std::vector<Box3>& boxes, int max_leaf, int builder);
bool bvh_cache_load(const std::string& dir, uint64_t hash, size_t n_prims, int builder, Bvh8& out);
bool bvh_cache_store(const std::string& dir, uint64_t hash, int builder, const Bvh8& bvh);

}  // namespace igb
