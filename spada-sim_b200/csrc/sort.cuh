// sort.cuh -- in-register bitonic sorting of 32*E keys per warp, shared by esc.cu and fused.cu.
// GPU restatement of the reference's SortingNetwork (simulator.rs:117-176): a fixed compare-
// exchange network, here over the (column, arrival index) keys of one C row.
#pragma once
#include "common.cuh"

namespace spada {

template <typename K> struct KeyTraits;
template <> struct KeyTraits<uint32_t> { static constexpr uint32_t sentinel = 0xffffffffu; };
template <> struct KeyTraits<uint64_t> { static constexpr uint64_t sentinel = 0xffffffffffffffffull; };

template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// ---- blocked register <-> shared memory transfer (lane l owns elements [l*E, (l+1)*E)) ----
template <typename K, int E>
__device__ __forceinline__ void load_blocked(K (&x)[E], const K* s, int lane) {
    constexpr int BYTES = E * (int)sizeof(K);
    const char* base = reinterpret_cast<const char*>(s + lane * E);
    if constexpr (BYTES % 16 == 0) {
        uint4 tmp[BYTES / 16];
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) tmp[i] = reinterpret_cast<const uint4*>(base)[i];
        memcpy(x, tmp, BYTES);
    } else if constexpr (BYTES == 8) {
        uint2 tmp = *reinterpret_cast<const uint2*>(base);
        memcpy(x, &tmp, 8);
    } else {
        x[0] = s[lane];
    }
}
template <typename K, int E>
__device__ __forceinline__ void store_blocked(const K (&x)[E], K* s, int lane) {
    constexpr int BYTES = E * (int)sizeof(K);
    char* base = reinterpret_cast<char*>(s + lane * E);
    if constexpr (BYTES % 16 == 0) {
        uint4 tmp[BYTES / 16];
        memcpy(tmp, x, BYTES);
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4*>(base)[i] = tmp[i];
    } else if constexpr (BYTES == 8) {
        uint2 tmp;
        memcpy(&tmp, x, 8);
        *reinterpret_cast<uint2*>(base) = tmp;
    } else {
        s[lane] = x[0];
    }
}

__device__ __forceinline__ uint32_t shfl_xor_key(uint32_t v, int m) { return __shfl_xor_sync(FULL, v, m); }
__device__ __forceinline__ uint64_t shfl_xor_key(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint32_t shfl_up_key(uint32_t v) { return __shfl_up_sync(FULL, v, 1); }
__device__ __forceinline__ uint64_t shfl_up_key(uint64_t v) {
    uint32_t lo = __shfl_up_sync(FULL, (uint32_t)v, 1);
    uint32_t hi = __shfl_up_sync(FULL, (uint32_t)(v >> 32), 1);
    return ((uint64_t)hi << 32) | lo;
}

// One phase (all strides KK/2 .. 1) of the bitonic network over the 32*E keys a warp holds in
// registers, element index e = lane*E + r.  `flip` inverts every comparison (descending).
template <typename K, int E, int KK>
__device__ __forceinline__ void bitonic_phase(K (&x)[E], int lane, bool flip) {
#pragma unroll
    for (int j = KK / 2; j > 0; j >>= 1) {
        if (j >= E) {
            const int lj = j / E;
            const bool lower = (lane & lj) == 0;
            const bool up_lane = (KK >= E) ? (((lane * E) & KK) == 0) : true;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const bool up = ((KK >= E) ? up_lane : ((r & KK) == 0)) != flip;
                K y = shfl_xor_key(x[r], lj);
                K lo = x[r] < y ? x[r] : y, hi = x[r] < y ? y : x[r];
                x[r] = (lower == up) ? lo : hi;
            }
        } else {
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if ((r & j) == 0) {
                    const bool up = ((KK >= E) ? (((lane * E) & KK) == 0) : ((r & KK) == 0)) != flip;
                    K a = x[r], b = x[r | j];
                    K lo = a < b ? a : b, hi = a < b ? b : a;
                    x[r] = up ? lo : hi;
                    x[r | j] = up ? hi : lo;
                }
            }
        }
    }
}

template <typename K, int E, int KK>
struct ChunkSort {
    static __device__ __forceinline__ void run(K (&x)[E], int lane, bool flip) {
        ChunkSort<K, E, KK / 2>::run(x, lane, flip);
        bitonic_phase<K, E, KK>(x, lane, flip);
    }
};
template <typename K, int E>
struct ChunkSort<K, E, 1> {
    static __device__ __forceinline__ void run(K (&)[E], int, bool) {}
};

// full sort of the warp's 32*E keys, ascending unless flip
template <typename K, int E>
__device__ __forceinline__ void warp_sort(K (&x)[E], int lane, bool flip) {
    ChunkSort<K, E, 32 * E>::run(x, lane, flip);
}
// last log2(32*E) stages of a larger merge phase: the warp's chunk is bitonic, direction uniform
template <typename K, int E>
__device__ __forceinline__ void warp_merge_tail(K (&x)[E], int lane, bool up) {
    bitonic_phase<K, E, 32 * E>(x, lane, !up);
}

}  // namespace spada
