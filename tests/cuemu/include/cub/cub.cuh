// cuemu shim for the two CUB device-wide primitives the library uses (tests only): same signatures, host execution.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "cuemu.h"

namespace cub {
struct DeviceRadixSort {
    template <typename K, typename V>
    static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, int n,
                                 int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
        if (tmp == nullptr) { tmp_bytes = 256; return cudaSuccess; }
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        const K mask = (end_bit >= (int)sizeof(K) * 8) ? ~K(0) : ((K(1) << end_bit) - 1);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ((keys_in[a] & mask) >> begin_bit) < ((keys_in[b] & mask) >> begin_bit); });
        for (int i = 0; i < n; ++i) { keys_out[i] = keys_in[idx[i]]; vals_out[i] = vals_in[idx[i]]; }
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <typename In, typename Out>
    static cudaError_t ExclusiveSum(void* tmp, size_t& tmp_bytes, const In* in, Out* out, int n, cudaStream_t = nullptr) {
        if (tmp == nullptr) { tmp_bytes = 256; return cudaSuccess; }
        Out acc = 0;
        for (int i = 0; i < n; ++i) { const Out v = (Out)in[i]; out[i] = acc; acc += v; }
        return cudaSuccess;
    }
};
}  // namespace cub
