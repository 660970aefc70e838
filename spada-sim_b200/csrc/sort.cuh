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

// Bitonic network in its all-ascending form: phase KK merges two ascending halves of every block of
// KK keys by first comparing element i with its mirror image i ^ (KK-1) ("flip"), then running
// half-cleaners of strides KK/4 .. 1.  Every compare-exchange keeps the smaller key at the lower
// index, so the register-to-register steps need no direction selects (two VIMNMX per pair) and
// chunks never have to be sorted descending.  Element index e = lane*E + r.
template <typename K>
__device__ __forceinline__ void cmpx(K& a, K& b) {
    const K lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo;
    b = hi;
}

// half-cleaners of strides JS, JS/2, .. 1
template <typename K, int E, int JS>
__device__ __forceinline__ void half_cleaners(K (&x)[E], int lane) {
#pragma unroll
    for (int j = JS; j > 0; j >>= 1) {
        if (j >= E) {
            const int lj = j / E;
            const bool lower = (lane & lj) == 0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const K y = shfl_xor_key(x[r], lj);
                const K lo = x[r] < y ? x[r] : y, hi = x[r] < y ? y : x[r];
                x[r] = lower ? lo : hi;
            }
        } else {
#pragma unroll
            for (int r = 0; r < E; ++r)
                if ((r & j) == 0) cmpx(x[r], x[r | j]);
        }
    }
}

template <typename K, int E, int KK>
__device__ __forceinline__ void bitonic_phase(K (&x)[E], int lane) {
    if constexpr (KK <= E) {
#pragma unroll
        for (int r = 0; r < E; ++r)
            if (r < (r ^ (KK - 1))) cmpx(x[r], x[r ^ (KK - 1)]);
    } else {
        constexpr int LM = KK / E - 1;
        const bool lower = (lane & (KK / (2 * E))) == 0;
        K y[E];
#pragma unroll
        for (int r = 0; r < E; ++r) y[r] = shfl_xor_key(x[E - 1 - r], LM);
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const K lo = x[r] < y[r] ? x[r] : y[r], hi = x[r] < y[r] ? y[r] : x[r];
            x[r] = lower ? lo : hi;
        }
    }
    if constexpr (KK >= 4) half_cleaners<K, E, KK / 4>(x, lane);
}

template <typename K, int E, int KK>
struct ChunkSort {
    static __device__ __forceinline__ void run(K (&x)[E], int lane) {
        ChunkSort<K, E, KK / 2>::run(x, lane);
        bitonic_phase<K, E, KK>(x, lane);
    }
};
template <typename K, int E>
struct ChunkSort<K, E, 1> {
    static __device__ __forceinline__ void run(K (&)[E], int) {}
};

// ascending sort of the warp's 32*E keys
template <typename K, int E>
__device__ __forceinline__ void warp_sort(K (&x)[E], int lane) {
    ChunkSort<K, E, 32 * E>::run(x, lane);
}
// the strides 16*E .. 1 of a larger merge phase whose wider strides were done in shared memory
template <typename K, int E>
__device__ __forceinline__ void warp_merge_tail(K (&x)[E], int lane) {
    half_cleaners<K, E, 16 * E>(x, lane);
}

}  // namespace spada
