// mf_sha1.cuh -- SHA-1 (FIPS 180-4) batch kernel for sm_100a: one lane per blob.
//
// Replaces hashlib.sha1(blob).hexdigest() at metaflow/datastore/content_addressed_store.py:98.
// SHA-1 is a serial chain over 64-byte blocks of ONE blob (Merkle-Damgard), so the only
// parallelism is across blobs: lane i of the grid walks blob i.  The critical path per round
// is rol(a,5) + (f+e+K+W) = SHF + IADD3 (two dependent ALU ops, 8 cycles on sm_100), i.e.
// >= 640 cycles per 64-byte block per stream; aggregate throughput = streams x ~190 MB/s up
// to the INT32 issue limit.  The next block's 64 bytes are prefetched into registers while the
// current block is compressed, so HBM latency stays off the chain.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

struct MfHashDesc {
    const uint8_t *src;   // device pointer, 16-byte aligned
    uint64_t len;
};

__device__ __forceinline__ uint32_t mf_rol(uint32_t x, int s) { return __funnelshift_l(x, x, s); }
__device__ __forceinline__ uint32_t mf_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// The only op on the serial chain is the last one: rol(a,5) + T compiles to a single LEA.HI.
// T = f + (e + K + W) depends on values that are at least one round old.
#define MF_SHA_R(F, K, a, b, c, d, e, wt)            \
    {                                                \
        uint32_t t0_ = e + (K) + (wt);               \
        uint32_t t1_ = (F) + t0_;                    \
        e = mf_rol(a, 5) + t1_;                      \
        b = mf_rol(b, 30);                           \
    }
#define MF_SHA_W(t) (w[(t) & 15] = mf_rol(w[((t) + 13) & 15] ^ w[((t) + 8) & 15] ^ w[((t) + 2) & 15] ^ w[(t) & 15], 1))
#define MF_F_CH(b, c, d) ((d) ^ ((b) & ((c) ^ (d))))
#define MF_F_PAR(b, c, d) ((b) ^ (c) ^ (d))
#define MF_F_MAJ(b, c, d) (((b) & (c)) | ((d) & ((b) | (c))))

// One compression of the 16 big-endian words in w[] into h[].  Fully unrolled so w[] lives in
// registers and the 5-variable rotation is pure renaming.
__device__ __forceinline__ void mf_sha1_compress(uint32_t h[5], uint32_t w[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    // round constants held in registers the compiler cannot see through: otherwise it re-associates
    // "+K" to the end of every round, i.e. onto the serial chain (LEA.HI -> VIADD instead of LEA.HI)
    uint32_t k1, k2, k3, k4;
    asm volatile("mov.u32 %0, 0x5A827999;" : "=r"(k1));
    asm volatile("mov.u32 %0, 0x6ED9EBA1;" : "=r"(k2));
    asm volatile("mov.u32 %0, 0x8F1BBCDC;" : "=r"(k3));
    asm volatile("mov.u32 %0, 0xCA62C1D6;" : "=r"(k4));
#define MF_R5(F, K, t, W0, W1, W2, W3, W4)               \
    MF_SHA_R(F(b, c, d), K, a, b, c, d, e, W0);          \
    MF_SHA_R(F(a, b, c), K, e, a, b, c, d, W1);          \
    MF_SHA_R(F(e, a, b), K, d, e, a, b, c, W2);          \
    MF_SHA_R(F(d, e, a), K, c, d, e, a, b, W3);          \
    MF_SHA_R(F(c, d, e), K, b, c, d, e, a, W4);
    MF_R5(MF_F_CH, k1, 0, w[0], w[1], w[2], w[3], w[4])
    MF_R5(MF_F_CH, k1, 5, w[5], w[6], w[7], w[8], w[9])
    MF_R5(MF_F_CH, k1, 10, w[10], w[11], w[12], w[13], w[14])
    MF_R5(MF_F_CH, k1, 15, w[15], MF_SHA_W(16), MF_SHA_W(17), MF_SHA_W(18), MF_SHA_W(19))
#pragma unroll
    for (int t = 20; t < 40; t += 5) { MF_R5(MF_F_PAR, k2, t, MF_SHA_W(t), MF_SHA_W(t + 1), MF_SHA_W(t + 2), MF_SHA_W(t + 3), MF_SHA_W(t + 4)) }
#pragma unroll
    for (int t = 40; t < 60; t += 5) { MF_R5(MF_F_MAJ, k3, t, MF_SHA_W(t), MF_SHA_W(t + 1), MF_SHA_W(t + 2), MF_SHA_W(t + 3), MF_SHA_W(t + 4)) }
#pragma unroll
    for (int t = 60; t < 80; t += 5) { MF_R5(MF_F_PAR, k4, t, MF_SHA_W(t), MF_SHA_W(t + 1), MF_SHA_W(t + 2), MF_SHA_W(t + 3), MF_SHA_W(t + 4)) }
#undef MF_R5
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

__device__ __forceinline__ void mf_sha1_load_block(const uint4 *p, uint32_t w[16]) {
    uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
    w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
    w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
    w[12] = v3.x; w[13] = v3.y; w[14] = v3.z; w[15] = v3.w;
}

// digests: 20 bytes per blob (big-endian words, i.e. the hashlib.sha1().digest() byte order)
__global__ void __launch_bounds__(128) mf_sha1_kernel(const MfHashDesc *__restrict__ blobs, uint32_t n, uint8_t *__restrict__ digests) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *src = blobs[i].src;
    uint64_t len = blobs[i].len;
    uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    uint64_t nfull = len >> 6;
    const uint4 *p = reinterpret_cast<const uint4 *>(src);
    uint32_t w[16], nx[16];
    if (nfull) mf_sha1_load_block(p, nx);
    for (uint64_t b = 0; b < nfull; b++) {
#pragma unroll
        for (int k = 0; k < 16; k++) w[k] = mf_bswap(nx[k]);
        if (b + 1 < nfull) mf_sha1_load_block(p + 4 * (b + 1), nx);   // prefetch next block
        mf_sha1_compress(h, w);
    }
    // tail: remaining r < 64 bytes, then 0x80, zero pad, 64-bit big-endian bit length.
    // Built in a separate (dynamically indexed, local-memory) array so that w[] above stays in registers.
    uint32_t r = (uint32_t)(len & 63u);
    const uint8_t *tp = src + (nfull << 6);
    uint32_t tw[32];
#pragma unroll
    for (int k = 0; k < 32; k++) tw[k] = 0;
    for (uint32_t k = 0; k < r; k++) tw[k >> 2] |= (uint32_t)tp[k] << (24 - 8 * (k & 3));
    tw[r >> 2] |= 0x80u << (24 - 8 * (r & 3));
    const uint64_t bits = len << 3;
    const int nb = (r >= 56) ? 2 : 1;
    tw[16 * nb - 2] = (uint32_t)(bits >> 32);
    tw[16 * nb - 1] = (uint32_t)bits;
    for (int blk = 0; blk < nb; blk++) {
        uint32_t t2[16];
#pragma unroll
        for (int k = 0; k < 16; k++) t2[k] = tw[16 * blk + k];
        mf_sha1_compress(h, t2);
    }
    uint8_t *out = digests + 20ull * i;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        out[4 * k] = (uint8_t)(h[k] >> 24); out[4 * k + 1] = (uint8_t)(h[k] >> 16);
        out[4 * k + 2] = (uint8_t)(h[k] >> 8); out[4 * k + 3] = (uint8_t)h[k];
    }
}
