// mfcas.cu -- host runtime and C ABI (include/mfcas.h) of the B200 artifact codec.
//
// The batcher side of the design: a batch of blobs (one foreach step's artifact set) is sharded
// round-robin over the context's GPUs (blob i -> device i mod G; independent blobs, no collective),
// each shard is cut into waves that fit the device arena, and every wave runs
//   H2D (cudaMemcpyAsync, coalesced over contiguous host ranges) -> kernels -> D2H
// on the device's streams.  This replaces the reference's process-per-task parallelism
// (metaflow/runtime.py:1517-1540) and multicore_utils.parallel_map (metaflow/multicore_utils.py:140-175)
// for the artifact work.  There is no CPU implementation of the path in this library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cuda.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mfcas.h"
#include "mf_core.h"
#include "mf_sha1.cuh"
#include "mf_deflate.cuh"
#include "mf_deflate2.cuh"
#include "mf_inflate.cuh"
#include "mf_inflate_par.cuh"

#define MFCAS_VERSION_STR "mfcas 0.2 (sm_100a; format MF v1: chunk 32 KiB, segment 1 KiB, member 2 MiB)"

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

struct HostBuf {     // pinned
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

enum class Op { Pack, Hash, Unpack };

// One in-flight wave: its own streams, events and buffers, so that several waves overlap on the device
// (H2D of wave k+1 | kernels of wave k | D2H of wave k-1; the serial SHA-1 kernels of many waves co-run).
struct WaveCtx {
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_in = nullptr, ev_sha = nullptr, ev_cmp = nullptr, ev_res = nullptr, ev_done = nullptr, ev_shares = nullptr;
    cudaEvent_t tl_start = nullptr, tl_in = nullptr, tl_res = nullptr, tl_done = nullptr;   // timeline (MFCAS_PHASE_TIMING)
    DevBuf in_arena, out_arena, d_desc, d_work, slots, phase, phase2, fscratch;
    HostBuf h_desc, h_res;
    // bookkeeping of the submitted wave
    int state = 0;                       // 0 idle, 1 kernels in flight, 2 output D2H in flight
    Op op = Op::Pack;
    uint64_t n = 0, n_out = 0, n_chunks = 0;
    unsigned sha_ctas = 0;
    size_t r_len = 0, r_st = 0, r_sha = 0, res_bytes = 0;
    bool timed_pack = false;
    size_t pos = 0, wn = 0;              // slice [pos, pos+wn) of the shard (host-pointer API)
    std::vector<const void *> w_src; std::vector<void *> w_dst;
    std::vector<uint64_t> w_len, w_cap, w_dlen, w_ioff, w_ooff;
    std::vector<int32_t> w_st; std::vector<uint8_t> w_sha;
};

struct DeviceState {
    int dev = -1;
    cudaStream_t compute = nullptr;      // all chunk-parallel encode kernels of all waves, in submission order
    DevBuf sha_state;                    // 5 words per blob of the running striped call
    DevBuf gate;                         // one word: SHA-1 CTAs resident (device-resident packing, see pack_submit)
    std::vector<WaveCtx> ring;
    std::string err;
    uint64_t launches = 0, h2d = 0, d2h = 0;
    float ms[8] = {};
    cudaEvent_t tl_origin = nullptr;     // MFCAS_PHASE_TIMING: time zero of the current call's wave timeline
};

}  // namespace

struct mfcas_ctx {
    std::vector<DeviceState> devs;
    uint64_t wave_bytes = 1ull << 30;
    uint32_t flags = 0;
    bool numa = true;
    std::string err;
    std::mutex mu;
};

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            D.err = _b;                                                                              \
            (void)cudaGetLastError();                                                                \
            return MFCAS_E_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

// carve helper over one contiguous buffer
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes, size_t align = 256) { off = align_up(off, align); size_t o = off; off += bytes; return o; }
};

#define MF_SHA_DEFAULT_VARIANT 0
#define MF_ENC_DEFAULT_CRC 0
#define MF_DEC_DEFAULT_OCC 4
#define MF_DEC_DEFAULT_CRC 0
#define MF_RING 8
// Encoder CTAs resident at once: two per SM, on the SMs that the SHA-1 CTAs in flight do not own (a SHA-1 CTA
// takes a whole SM: 216 KB of shared memory).  The persistent encoder grid is sized to leave those SMs free;
// otherwise whichever kernel's CTAs are dispatched first owns the SMs and the hash chain (the longest pole of a
// wave) waits for the whole encoder.
#define MF_SM_SLOTS (2u * 148u)
#define MF_ENC_GRID_MIN 168u

int wave_setup(DeviceState &D, WaveCtx &W) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithFlags(&W.stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithPriority(&W.stream2, cudaStreamNonBlocking, hi));   // SHA-1 CTAs are dispatched first
    for (auto &e : W.ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&W.ev_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&W.ev_sha, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&W.ev_cmp, cudaEventDisableTiming));
    CU(cudaEventCreate(&W.tl_start)); CU(cudaEventCreate(&W.tl_in)); CU(cudaEventCreate(&W.tl_res)); CU(cudaEventCreate(&W.tl_done));
    CU(cudaEventCreateWithFlags(&W.ev_res, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&W.ev_shares, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&W.ev_done, cudaEventDisableTiming));
    return MFCAS_OK;
}

int device_setup(DeviceState &D) {
    CU(cudaSetDevice(D.dev));
    CU(cudaStreamCreateWithFlags(&D.compute, cudaStreamNonBlocking));
    D.ring.resize(MF_RING);
    for (auto &W : D.ring) { int rc = wave_setup(D, W); if (rc != MFCAS_OK) return rc; }
    for (auto k : {mf_sha1_kernel<0>, mf_sha1_kernel<1>, mf_sha1_kernel<3>, mf_sha1_kernel<4>}) {
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_SHA_UNITS * sizeof(MfShaSmem))));
        CU(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    CU(cudaFuncSetAttribute(mf_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MfEncSmem)));
    for (auto k : {mf_deflate2_kernel<0>, mf_deflate2_kernel<1>}) {
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MfEnc2Smem)));
        CU(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    for (auto k : {mf_inflate_chunk_kernel<5, 0>, mf_inflate_chunk_kernel<4, 0>, mf_inflate_chunk_kernel<4, 1>})
    {
        static_assert(4u * (MF_DEC_SMEM_BYTES + 1024u) <= 228u * 1024u, "four decoder CTAs per SM must fit the B200's 228 KB of shared memory");
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MF_DEC_SMEM_BYTES));
        CU(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    CU(cudaFuncSetAttribute(mf_fsize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_F_LANES * sizeof(MfSpecLane))));
    CU(cudaFuncSetAttribute(mf_fdecode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MF_F_LANES * sizeof(MfSpecLane))));
    // SHA-1 and the encoder share SMs (one 100+ KB CTA each).  An SM cannot hold CTAs of kernels whose L1/shared
    // split differs, and re-splitting needs the SM idle -- with the driver's per-kernel default the hash CTAs wait
    // for the persistent encoder to drain whenever the encoder happens to be dispatched first.  Same split for both.
    CU(cudaFuncSetAttribute(mf_deflate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    mf_init_tables_kernel<<<2, 256, 0, D.ring[0].stream>>>();
    CU(cudaGetLastError());
    mf_init_tables2_kernel<<<4, 256, 0, D.ring[0].stream>>>();
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(D.ring[0].stream));
    return MFCAS_OK;
}

void device_teardown(DeviceState &D) {
    if (D.dev < 0) return;
    cudaSetDevice(D.dev);
    for (auto &W : D.ring) {
        if (W.stream) cudaStreamSynchronize(W.stream);
        if (W.stream2) cudaStreamSynchronize(W.stream2);
        W.in_arena.release(); W.out_arena.release(); W.d_desc.release(); W.d_work.release(); W.slots.release();
        W.phase.release(); W.phase2.release(); W.fscratch.release(); W.h_desc.release(); W.h_res.release();
        for (auto &e : W.ev) if (e) cudaEventDestroy(e);
        if (D.tl_origin) { cudaEventDestroy(D.tl_origin); D.tl_origin = nullptr; }
        for (cudaEvent_t e : {W.ev_in, W.ev_sha, W.ev_cmp, W.ev_res, W.ev_shares, W.ev_done, W.tl_start, W.tl_in, W.tl_res, W.tl_done}) if (e) cudaEventDestroy(e);
        if (W.stream) cudaStreamDestroy(W.stream);
        if (W.stream2) cudaStreamDestroy(W.stream2);
    }
    D.ring.clear();
    if (D.compute) { cudaStreamSynchronize(D.compute); cudaStreamDestroy(D.compute); D.compute = nullptr; }
    D.sha_state.release();
    D.gate.release();
}

// Copies a set of (host ptr, device ptr, len) with runs of contiguous pairs merged into one memcpy.
struct CopyPlan {
    struct Item { char *h; char *d; size_t n; };
    std::vector<Item> items;
    void add(const void *h, void *d, size_t n) {
        if (!n) return;
        char *hc = (char *)h, *dc = (char *)d;
        if (!items.empty()) {
            Item &l = items.back();
            ptrdiff_t gh = hc - l.h, gd = dc - l.d;
            // same displacement on both sides and the gap after the previous item is small: extend
            if (gh == gd && gh >= (ptrdiff_t)l.n && gh - (ptrdiff_t)l.n <= 4096) { l.n = (size_t)gh + n; return; }
        }
        items.push_back({hc, dc, n});
    }
    // Enqueue the plan.  Runs of equal-sized items at a constant pitch on both sides (the stripes of equally
    // sized, equally spaced blobs -- a foreach's artifacts) go as ONE pitched copy: a wave of 2 MiB stripes is
    // then a single DMA descriptor instead of hundreds of back-to-back small copies.
    cudaError_t run(cudaMemcpyKind kind, cudaStream_t st, uint64_t *bytes) const {
        size_t i = 0;
        while (i < items.size()) {
            const Item &a = items[i];
            size_t j = i + 1;
            ptrdiff_t hp = 0, dp = 0;
            if (j < items.size() && items[j].n == a.n) {
                hp = items[j].h - a.h; dp = items[j].d - a.d;
                const ptrdiff_t lim = 0x7FFFFFFF;
                if (hp >= (ptrdiff_t)a.n && dp >= (ptrdiff_t)a.n && hp <= lim && dp <= lim) {
                    while (j < items.size() && items[j].n == a.n && items[j].h - items[j - 1].h == hp && items[j].d - items[j - 1].d == dp) j++;
                } else j = i + 1;
            }
            cudaError_t e;
            if (j - i >= 2) {
                e = kind == cudaMemcpyHostToDevice ? cudaMemcpy2DAsync(a.d, (size_t)dp, a.h, (size_t)hp, a.n, j - i, kind, st)
                                                   : cudaMemcpy2DAsync(a.h, (size_t)hp, a.d, (size_t)dp, a.n, j - i, kind, st);
            } else {
                j = i + 1;
                e = kind == cudaMemcpyHostToDevice ? cudaMemcpyAsync(a.d, a.h, a.n, kind, st) : cudaMemcpyAsync(a.h, a.d, a.n, kind, st);
            }
            if (e != cudaSuccess) return e;
            if (bytes) *bytes += (uint64_t)a.n * (j - i);
            i = j;
        }
        return cudaSuccess;
    }
};

typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
// the driver entry point, fetched through the runtime so that the library does not link libcuda
static StreamWaitValue32Fn stream_wait_value_fn() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { (void)cudaGetLastError(); return nullptr; }
    return reinterpret_cast<StreamWaitValue32Fn>(fn);
}

static bool phase_timing_on() { static const bool on = getenv("MFCAS_PHASE_TIMING") != nullptr; return on; }

// ------------------------------------------------------------------ pack: enqueue (async) / collect
// Inputs and outputs are DEVICE pointers.  Everything is enqueued on W's streams, including the copy of the
// small result block (dst_len, status, digests) to pinned host memory; W.ev_res fires when that is there.
// `upload` (host-pointer API): the wave's input copies; they are enqueued AFTER the small descriptor copy so
// that nothing small ever queues behind other waves' gigabyte copies on the copy engine.
// `chain` (striped packing): SHA-1 descriptors of message PIECES (state carried across waves); `sha_deps` are
// the ev_sha events of the in-flight waves that hold the previous piece of any of these blobs.
// Where the encoder's members go:
//   PackContiguous  one contiguous RFC 1952 file per blob in d_dst[i]: members are encoded into scratch slots and moved
//                   together by the compaction kernels (mfcas_pack_batch, mfcas_pack_device)
//   PackUnits       every unit IS one member (a stripe of the striped host API): written straight to d_dst[k]
//   PackSlots       device-resident extents: member m of blob i is written straight to d_dst[i] + m * ext_slot();
//                   results (length, status) are per MEMBER, blob-major
enum PackLayout { PackContiguous = 0, PackUnits = 1, PackSlots = 2 };
inline uint64_t ext_slot_bytes() { return align_up(mf_pack_bound(MF_MEMBER_RAW), 256); }
static bool enc_v1() { static const bool on = getenv("MFCAS_ENC") != nullptr && atoi(getenv("MFCAS_ENC")) == 1; return on; }

int pack_submit(DeviceState &D, WaveCtx &W, uint32_t flags, uint64_t n, const void *const *d_src, const uint64_t *src_len,
                void *const *d_dst, const uint64_t *dst_cap, bool want_pack, const CopyPlan *upload,
                const MfHashDesc *chain = nullptr, const std::vector<cudaEvent_t> *sha_deps = nullptr, PackLayout layout = PackContiguous) {
    if (n > 0x7FFFFFFFull) return MFCAS_E_ARG;
    const bool v1 = enc_v1() && layout != PackSlots;
    uint64_t n_members = 0, n_chunks = 0;
    if (want_pack)
        for (uint64_t i = 0; i < n; i++) { n_members += mf_n_members(src_len[i]); n_chunks += mf_n_chunks(src_len[i]); }
    if (layout == PackUnits && n_members != n) return MFCAS_E_ARG;           // a unit is at most one stripe
    const uint64_t max_groups = n_chunks / MF_ENC_GROUP + n_members;      // v1 encoder work units: <= MF_ENC_GROUP chunks of one member
    if (n_chunks > 0x7FFFFFF0ull || n_members > 0x7FFFFFF0ull) return MFCAS_E_ARG;
    const uint64_t n_out = (want_pack && layout == PackSlots) ? n_members : n;   // entries of the length / status result arrays
    Carver hc;
    const size_t o_hash = hc.take(sizeof(MfHashDesc) * n);
    const size_t o_blob = hc.take(std::max(sizeof(MfEncBlob), sizeof(MfCompactBlob)) * n);
    const size_t o_mem = hc.take(std::max(sizeof(MfEncMember), sizeof(MfEnc2Member)) * n_members);
    const size_t o_cm = hc.take(sizeof(uint32_t) * (v1 ? n_chunks : n_members));      // v1: chunk -> member; v2: member -> blob
    const size_t o_grp = hc.take(sizeof(uint32_t) * max_groups);
    const size_t desc_bytes = hc.take(0);
    CU(W.h_desc.ensure(desc_bytes));
    CU(W.d_desc.ensure(desc_bytes));
    uint8_t *hd = W.h_desc.as<uint8_t>();
    MfHashDesc *h_hash = reinterpret_cast<MfHashDesc *>(hd + o_hash);
    uint32_t *h_cm = reinterpret_cast<uint32_t *>(hd + o_cm);
    uint32_t *h_grp = reinterpret_cast<uint32_t *>(hd + o_grp);
    uint32_t n_groups = 0;
    if (chain) {
        memcpy(h_hash, chain, sizeof(MfHashDesc) * n);
    } else {   // SHA-1 streams sorted by length (descending): the 32 streams of a CTA then finish together
        std::vector<uint32_t> order(n);
        for (uint64_t i = 0; i < n; i++) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return src_len[a] > src_len[b]; });
        for (uint64_t k = 0; k < n; k++) {
            const uint32_t i = order[k];
            h_hash[k].src = static_cast<const uint8_t *>(d_src[i]);
            h_hash[k].len = src_len[i];
            h_hash[k].out_index = i;
            h_hash[k].flags = MF_SHA_FIRST | MF_SHA_LAST;
            h_hash[k].total_len = src_len[i];
            h_hash[k].state = nullptr;
        }
    }
    const uint64_t slot_stride = ext_slot_bytes();
    if (want_pack && !v1 && layout == PackContiguous) CU(W.slots.ensure((size_t)n_members * slot_stride + 256));
    uint32_t mi = 0, ci = 0;
    if (want_pack && v1) {
        MfEncBlob *h_blob = reinterpret_cast<MfEncBlob *>(hd + o_blob);
        MfEncMember *h_mem = reinterpret_cast<MfEncMember *>(hd + o_mem);
        for (uint64_t i = 0; i < n; i++) {
            h_blob[i].dst = static_cast<uint8_t *>(d_dst[i]);
            h_blob[i].dst_cap = dst_cap[i];
            h_blob[i].first_member = mi;
            uint64_t nm = mf_n_members(src_len[i]);
            h_blob[i].n_members = (uint32_t)nm;
            for (uint64_t m = 0; m < nm; m++) {
                uint64_t off = m * MF_MEMBER_RAW;
                uint64_t rl = std::min<uint64_t>(MF_MEMBER_RAW, src_len[i] - off);
                MfEncMember &M = h_mem[mi];
                M.src = static_cast<const uint8_t *>(d_src[i]) + off;
                M.raw_len = (uint32_t)rl;
                M.n_chunks = (uint32_t)mf_n_chunks(rl);
                M.first_chunk = ci;
                M.blob = (uint32_t)i;
                for (uint32_t k = 0; k < M.n_chunks; k += MF_ENC_GROUP) h_grp[n_groups++] = ci + k;
                for (uint32_t k = 0; k < M.n_chunks; k++) h_cm[ci++] = mi;
                mi++;
            }
        }
    } else if (want_pack) {
        MfCompactBlob *h_blob = reinterpret_cast<MfCompactBlob *>(hd + o_blob);
        MfEnc2Member *h_mem = reinterpret_cast<MfEnc2Member *>(hd + o_mem);
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t nm = mf_n_members(src_len[i]);
            h_blob[i].dst = static_cast<uint8_t *>(d_dst[i]);
            h_blob[i].dst_cap = dst_cap[i];
            h_blob[i].first_member = mi;
            h_blob[i].n_members = (uint32_t)nm;
            for (uint64_t m = 0; m < nm; m++) {
                const uint64_t off = m * MF_MEMBER_RAW;
                const uint64_t rl = std::min<uint64_t>(MF_MEMBER_RAW, src_len[i] - off);
                MfEnc2Member &M = h_mem[mi];
                M.src = static_cast<const uint8_t *>(d_src[i]) + off;
                M.raw_len = (uint32_t)rl;
                M.n_chunks = (uint32_t)mf_n_chunks(rl);
                M.first_chunk = ci;
                ci += M.n_chunks;
                if (layout == PackContiguous) { M.dst = W.slots.as<uint8_t>() + (size_t)mi * slot_stride; M.dst_cap = slot_stride; M.unit = mi; }
                else if (layout == PackUnits) { M.dst = static_cast<uint8_t *>(d_dst[i]); M.dst_cap = dst_cap[i]; M.unit = (uint32_t)i; }
                else {
                    // a blob whose destination cannot hold all its slots gets no member written at all
                    const bool room = dst_cap[i] >= (nm - 1) * slot_stride + mf_pack_bound(src_len[i] - (nm - 1) * MF_MEMBER_RAW);
                    M.dst = static_cast<uint8_t *>(d_dst[i]) + m * slot_stride; M.dst_cap = room ? slot_stride : 0; M.unit = mi;
                }
                h_cm[mi] = (uint32_t)i;
                mi++;
            }
        }
    }
    Carver wc;
    const size_t o_comp = wc.take(sizeof(uint32_t) * (v1 ? n_chunks : 0));
    const size_t o_hdr = wc.take(sizeof(uint16_t) * (v1 ? n_chunks : 0));
    const size_t o_seg = wc.take(sizeof(uint16_t) * (v1 ? n_chunks : 0) * MF_SEGS_PER_CHUNK);
    const size_t o_crc = wc.take(sizeof(uint32_t) * (v1 ? n_chunks : 0));
    const size_t o_cdst = wc.take(sizeof(uint64_t) * (v1 ? n_chunks : 0));
    const bool compact = want_pack && !v1 && layout == PackContiguous;
    const size_t o_mlen = wc.take(sizeof(uint64_t) * (compact ? n_members : 0));     // member lengths / status / offsets (compaction)
    const size_t o_mst = wc.take(sizeof(int32_t) * (compact ? n_members : 0));
    const size_t o_moff = wc.take(sizeof(uint64_t) * (compact ? n_members : 0));
    const size_t o_ticket = wc.take(sizeof(uint32_t));
    const size_t o_res = wc.take(0);
    Carver rc;
    W.r_len = rc.take(sizeof(uint64_t) * n_out, 16);
    W.r_st = rc.take(sizeof(int32_t) * n_out, 16);
    W.r_sha = rc.take(20 * n, 16);
    W.res_bytes = rc.take(0, 16);
    CU(W.d_work.ensure(o_res + W.res_bytes));
    CU(W.h_res.ensure(W.res_bytes));
    if (want_pack && v1) CU(W.slots.ensure((size_t)n_chunks * MF_SLOT + 64));
    uint8_t *dw = W.d_work.as<uint8_t>();
    uint8_t *dd = W.d_desc.as<uint8_t>();
    uint8_t *dres = dw + o_res;
    cudaStream_t s = W.stream, s2 = W.stream2;
    unsigned long long *phase = nullptr;      // MFCAS_PHASE_TIMING=1: per-chunk clock64 at phase boundaries
    if (phase_timing_on() && want_pack && n_chunks) {
        CU(W.phase.ensure(n_chunks * 16 * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(W.phase.p, 0, n_chunks * 16 * sizeof(unsigned long long), s));
        phase = W.phase.as<unsigned long long>();
    }
    CU(cudaMemcpyAsync(dd, hd, desc_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(dres + W.r_st, 0, sizeof(int32_t) * n_out, s));
    if (compact && n_members) CU(cudaMemsetAsync(dw + o_mst, 0, sizeof(int32_t) * n_members, s));
    CU(cudaMemsetAsync(dw + o_ticket, 0, sizeof(uint32_t), s));
    if (upload) {
        if (phase_timing_on()) CU(cudaEventRecord(W.tl_start, s));
        CU(upload->run(cudaMemcpyHostToDevice, s, &D.h2d));
        if (phase_timing_on()) CU(cudaEventRecord(W.tl_in, s));
    }
    // Device-resident packing is ONE wave: a hash kernel that runs as long as its longest blob beside one long
    // persistent encoder.  A SHA-1 CTA needs an SM of its own (216 KB of shared memory); if the encoder's CTAs are
    // dispatched first they sit on every SM until the encoder is done and the hash chain -- the longest pole --
    // starts late.  Make it deterministic: the encoder's stream waits (cuStreamWaitValue32) until every hash CTA
    // has reported in.  (Host-pointer waves are short and many; there the encoder just leaves SMs free, below.)
    static const StreamWaitValue32Fn wait_value = stream_wait_value_fn();
    uint32_t *gate = nullptr;
    const unsigned sha_ctas = (unsigned)((n + MF_SHA_STREAMS_PER_CTA - 1) / MF_SHA_STREAMS_PER_CTA);
    const size_t sha_smem = MF_SHA_UNITS * sizeof(MfShaSmem);
    if (!upload && !chain && want_pack && sha_ctas <= 64 && wait_value) {
        CU(D.gate.ensure(256));
        gate = D.gate.as<uint32_t>();
        CU(cudaMemsetAsync(gate, 0, sizeof(uint32_t), s));          // before ev_in: both kernels' streams wait on it
    }
    CU(cudaEventRecord(W.ev_in, s));
    // ---- SHA-1 on the high-priority second stream (serial per blob: it overlaps the chunk-parallel encoder)
    CU(cudaStreamWaitEvent(s2, W.ev_in, 0));
    if (sha_deps) for (cudaEvent_t e : *sha_deps) CU(cudaStreamWaitEvent(s2, e, 0));   // earlier pieces of these blobs
    CU(cudaEventRecord(W.ev[0], s2));
    unsigned long long *sha_dbg = nullptr;
    if (phase_timing_on()) { CU(W.phase2.ensure((size_t)sha_ctas * MF_SHA_UNITS * 4 * sizeof(unsigned long long))); sha_dbg = W.phase2.as<unsigned long long>(); }
    W.sha_ctas = sha_ctas;
    {
        static const int variant = getenv("MFCAS_SHA_VARIANT") ? atoi(getenv("MFCAS_SHA_VARIANT")) : MF_SHA_DEFAULT_VARIANT;
        auto kern = variant == 1 ? mf_sha1_kernel<1> : variant == 3 ? mf_sha1_kernel<3> : variant == 4 ? mf_sha1_kernel<4> : mf_sha1_kernel<0>;
        kern<<<sha_ctas, MF_SHA_THREADS, sha_smem, s2>>>(reinterpret_cast<const MfHashDesc *>(dd + o_hash), (uint32_t)n, dres + W.r_sha, sha_dbg, gate, 1u, 1u << 30);
    }
    CU(cudaGetLastError());
    D.launches++;
    CU(cudaEventRecord(W.ev[1], s2));
    CU(cudaEventRecord(W.ev_sha, s2));
    if (want_pack) {
        cudaStream_t sc = D.compute;          // one encoder kernel at a time per device (its grid leaves the SHA-1 SMs free)
        CU(cudaStreamWaitEvent(sc, W.ev_in, 0));
        if (gate && wait_value((CUstream)sc, (CUdeviceptr)(uintptr_t)gate, sha_ctas, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
            D.err = "cuStreamWaitValue32 failed"; return MFCAS_E_CUDA;
        }
        CU(cudaEventRecord(W.ev[2], sc));
        unsigned sha_sms = 0;             // SMs owned by SHA-1 CTAs of this wave and of the waves still in flight
        for (WaveCtx &O : D.ring) if (&O == &W || O.state == 1) sha_sms += O.sha_ctas;
        if (v1) {
            if (n_chunks) {
                const unsigned enc_cap = std::max(MF_ENC_GRID_MIN, MF_SM_SLOTS - 2u * std::min(sha_sms, 148u));
                mf_deflate_kernel<<<(unsigned)std::min<uint64_t>(n_groups, enc_cap), MF_ENC_THREADS, sizeof(MfEncSmem), sc>>>(
                    reinterpret_cast<const MfEncMember *>(dd + o_mem), reinterpret_cast<const uint32_t *>(dd + o_cm), (uint32_t)n_chunks,
                    W.slots.as<uint8_t>(), reinterpret_cast<uint32_t *>(dw + o_comp), reinterpret_cast<uint16_t *>(dw + o_hdr),
                    reinterpret_cast<uint16_t *>(dw + o_seg), reinterpret_cast<uint32_t *>(dw + o_crc), flags, phase, reinterpret_cast<uint32_t *>(dw + o_ticket),
                    reinterpret_cast<const uint32_t *>(dd + o_grp), n_groups);
                CU(cudaGetLastError());
                D.launches++;
            }
            CU(cudaEventRecord(W.ev[3], sc));
            mf_gather_layout_kernel<<<(unsigned)n, MF_GA_THREADS, 0, sc>>>(
                reinterpret_cast<const MfEncBlob *>(dd + o_blob), reinterpret_cast<const MfEncMember *>(dd + o_mem),
                reinterpret_cast<const uint32_t *>(dw + o_comp), reinterpret_cast<const uint16_t *>(dw + o_hdr),
                reinterpret_cast<const uint16_t *>(dw + o_seg), reinterpret_cast<const uint32_t *>(dw + o_crc),
                reinterpret_cast<uint64_t *>(dw + o_cdst), reinterpret_cast<uint64_t *>(dres + W.r_len), reinterpret_cast<int32_t *>(dres + W.r_st));
            CU(cudaGetLastError());
            D.launches++;
            if (n_chunks) {
                mf_gather_copy_kernel<<<(unsigned)n_chunks, 256, 0, sc>>>(
                    reinterpret_cast<const MfEncBlob *>(dd + o_blob), reinterpret_cast<const MfEncMember *>(dd + o_mem),
                    reinterpret_cast<const uint32_t *>(dd + o_cm), (uint32_t)n_chunks, W.slots.as<uint8_t>(),
                    reinterpret_cast<const uint32_t *>(dw + o_comp), reinterpret_cast<const uint64_t *>(dw + o_cdst));
                CU(cudaGetLastError());
                D.launches++;
            }
        } else {
            // the member-owned encoder: one CTA per SM on the SMs the hash CTAs in flight leave free
            // Host waves: stay off the SMs the hash CTAs of the waves in flight need.  The one-wave device-resident call
            // asks for every SM (the hash kernel was launched first, on the high-priority stream, and -- up to 64 hash
            // CTAs -- gated to be resident before the encoder is dispatched): the encoder CTAs that do not fit wait in
            // the hardware queue and take over the hash CTAs' SMs the moment those finish.
            const unsigned enc_cap = (gate || !upload) ? 148u : std::max(74u, 148u - std::min(sha_sms, 148u));
            uint64_t *mlen = compact ? reinterpret_cast<uint64_t *>(dw + o_mlen) : reinterpret_cast<uint64_t *>(dres + W.r_len);
            int32_t *mst = compact ? reinterpret_cast<int32_t *>(dw + o_mst) : reinterpret_cast<int32_t *>(dres + W.r_st);
            if (n_members) {
                static const int crcv = getenv("MFCAS_ENC_CRC") ? atoi(getenv("MFCAS_ENC_CRC")) : MF_ENC_DEFAULT_CRC;
                auto enc = crcv == 1 ? mf_deflate2_kernel<1> : mf_deflate2_kernel<0>;
                enc<<<(unsigned)std::min<uint64_t>(n_members, enc_cap), MF2_THREADS, sizeof(MfEnc2Smem), sc>>>(
                    reinterpret_cast<const MfEnc2Member *>(dd + o_mem), (uint32_t)n_members, mlen, mst, flags, phase, reinterpret_cast<uint32_t *>(dw + o_ticket));
                CU(cudaGetLastError());
                D.launches++;
            }
            CU(cudaEventRecord(W.ev[3], sc));
            if (compact && n) {
                mf_compact_layout_kernel<<<(unsigned)n, 256, 0, sc>>>(reinterpret_cast<const MfCompactBlob *>(dd + o_blob), mlen, mst,
                                                                      reinterpret_cast<uint64_t *>(dw + o_moff), reinterpret_cast<uint64_t *>(dres + W.r_len),
                                                                      reinterpret_cast<int32_t *>(dres + W.r_st));
                CU(cudaGetLastError());
                const unsigned tiles = (unsigned)((slot_stride + MF_COMPACT_TILE - 1) / MF_COMPACT_TILE);
                mf_compact_copy_kernel<<<(unsigned)(n_members * tiles), 256, 0, sc>>>(
                    reinterpret_cast<const MfCompactBlob *>(dd + o_blob), reinterpret_cast<const uint32_t *>(dd + o_cm), W.slots.as<uint8_t>(), slot_stride,
                    mlen, reinterpret_cast<const uint64_t *>(dw + o_moff), reinterpret_cast<const int32_t *>(dres + W.r_st), tiles);
                CU(cudaGetLastError());
                D.launches += 2;
            }
        }
        CU(cudaEventRecord(W.ev[4], sc));
        CU(cudaEventRecord(W.ev_cmp, sc));
        CU(cudaStreamWaitEvent(s, W.ev_cmp, 0));
    }
    // lengths + status follow the encoder on `s` (the caller's bulk copy-back queues right behind them); the
    // digests follow the SHA-1 kernel on `s2`, so a long hash chain never holds back the copy-back
    if (want_pack) CU(cudaMemcpyAsync(W.h_res.p, dres, W.r_sha, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(W.ev_res, s));
    CU(cudaMemcpyAsync(W.h_res.as<uint8_t>() + W.r_sha, dres + W.r_sha, W.res_bytes - W.r_sha, cudaMemcpyDeviceToHost, s2));
    CU(cudaEventRecord(W.ev_shares, s2));
    W.n = n; W.n_out = n_out; W.n_chunks = n_chunks; W.sha_ctas = sha_ctas; W.timed_pack = want_pack;
    return MFCAS_OK;
}

int pack_collect(DeviceState &D, WaveCtx &W, uint64_t *dst_len, uint8_t *sha1, int32_t *status) {
    CU(cudaEventSynchronize(W.ev_res));
    CU(cudaEventSynchronize(W.ev_shares));
    const uint8_t *hr = W.h_res.as<uint8_t>();
    const uint64_t n = W.n;
    if (sha1) memcpy(sha1, hr + W.r_sha, 20 * n);
    if (W.timed_pack) {
        memcpy(dst_len, hr + W.r_len, sizeof(uint64_t) * W.n_out);
        memcpy(status, hr + W.r_st, sizeof(int32_t) * W.n_out);
        cudaEventElapsedTime(&D.ms[1], W.ev[2], W.ev[3]);
        cudaEventElapsedTime(&D.ms[2], W.ev[3], W.ev[4]);
    } else if (status) {
        for (uint64_t i = 0; i < n; i++) status[i] = 0;
    }
    cudaEventElapsedTime(&D.ms[0], W.ev[0], W.ev[1]);
    if (phase_timing_on()) {
        if (W.phase2.p) {
            const unsigned nu = W.sha_ctas * MF_SHA_UNITS;
            std::vector<unsigned long long> h((size_t)nu * 4);
            CU(cudaMemcpy(h.data(), W.phase2.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            double a[4] = {0};
            for (unsigned c = 0; c < nu; c++) for (int k = 0; k < 4; k++) a[k] += (double)h[c * 4 + k];
            fprintf(stderr, "[mfcas phase] sha1 over %u units: producer wait=%.0f total=%.0f | consumer wait=%.0f total=%.0f cycles\n", nu,
                    a[0] / nu, a[1] / nu, a[2] / nu, a[3] / nu);
        }
        if (W.timed_pack && W.n_chunks && W.phase.p) {
            const uint64_t nc = W.n_chunks;
            std::vector<unsigned long long> h(nc * 16);
            CU(cudaMemcpy(h.data(), W.phase.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (!enc_v1()) {
                // member-owned encoder: clock64 of thread 0 at 0 start, 1 chunk loaded, 2 LZ done (warp 0), 3 sweep done,
                // 4 verdict known, 5 codes settled (rebuilt if rejected), 6 image complete, 7 stored
                static const char *nm2[7] = {"wait_load", "lz77(w0)", "sweep(w0)", "verdict", "rebuild", "place", "store"};
                double acc[7] = {0};
                uint64_t rebuilt = 0;
                for (uint64_t c = 0; c < nc; c++) {
                    for (int k = 0; k < 7; k++) acc[k] += (double)(h[c * 16 + k + 1] - h[c * 16 + k]);
                    if (h[c * 16 + 5] - h[c * 16 + 4] > 2000) rebuilt++;
                }
                fprintf(stderr, "[mfcas phase] deflate2, mean cycles per chunk over %llu chunks:", (unsigned long long)nc);
                double tot = 0;
                for (int k = 0; k < 7; k++) { fprintf(stderr, " %s=%.0f", nm2[k], acc[k] / (double)nc); tot += acc[k]; }
                fprintf(stderr, " total=%.0f | chunks with rebuilt codes: %llu\n", tot / (double)nc, (unsigned long long)rebuilt);
                return MFCAS_OK;
            }
            // boundaries in program order: 0 start,1 loaded,2 crc,3 lz,8 litlen code,9 dist code,10 rle,11 zeroed,12 cl code,13 header,4,5,6,7
            static const int seq[14] = {0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 4, 5, 6, 7};
            static const char *nm[13] = {"load", "crc", "lz77", "codes(warp0)", "codes_sync", "hdr_scan", "hdr_tokens", "hdr_clcode", "hdr_write", "sync", "bitcount", "pack", "store"};
            double acc[13] = {0};
            for (uint64_t c = 0; c < nc; c++)
                for (int k = 0; k < 13; k++) acc[k] += (double)(h[c * 16 + seq[k + 1]] - h[c * 16 + seq[k]]);
            fprintf(stderr, "[mfcas phase] deflate, mean cycles per chunk over %llu chunks:", (unsigned long long)nc);
            for (int k = 0; k < 13; k++) fprintf(stderr, " %s=%.0f", nm[k], acc[k] / (double)nc);
            double x0 = 0, x1 = 0, x2 = 0;
            for (uint64_t c = 0; c < nc; c++) { x0 += (double)(h[c * 16 + 14] - h[c * 16 + 9]); x1 += (double)(h[c * 16 + 15] - h[c * 16 + 14]); x2 += (double)(h[c * 16 + 10] - h[c * 16 + 15]); }
            fprintf(stderr, " | hdr_scan split: trim+fill=%.0f fwd=%.0f bwd=%.0f\n", x0 / (double)nc, x1 / (double)nc, x2 / (double)nc);
        }
    }
    return MFCAS_OK;
}

// ------------------------------------------------------------------ unpack: enqueue (async) / collect
// `foreign` (may be null): per blob, 1 when the HOST copy of the blob does not start with an 'MF'-indexed member, i.e.
// a stock gzip stream: those go through the block-parallel foreign decoder (mf_inflate_par.cuh) first.
int unpack_submit(DeviceState &D, WaveCtx &W, uint64_t n, const void *const *d_src, const uint64_t *src_len,
                  void *const *d_dst, const uint64_t *dst_cap, const CopyPlan *upload, const uint8_t *foreign = nullptr) {
    if (n > 0x7FFFFFFFull) return MFCAS_E_ARG;
    uint64_t max_chunks = 0, max_members = 0;
    for (uint64_t i = 0; i < n; i++) {       // the parse kernel enforces exactly these per-blob quotas
        const uint64_t qm = dst_cap[i] / MF_DEC_MIN_MEMBER_RAW + 2;
        max_members += qm; max_chunks += dst_cap[i] / MF_CHUNK + qm;
    }
    if (max_chunks > 0x7FFFFFF0ull) return MFCAS_E_ARG;
    // foreign blobs: descriptors + scratch of the block-parallel decoder
    std::vector<MfFBlob> fbl;
    uint64_t total_ranges = 0, out16_syms = 0;
    if (foreign)
        for (uint64_t i = 0; i < n; i++) {
            if (!foreign[i] || src_len[i] < 18 || total_ranges > 0x60000000ull) continue;
            MfFBlob f;
            f.src = static_cast<const uint8_t *>(d_src[i]); f.src_len = src_len[i];
            f.dst = static_cast<uint8_t *>(d_dst[i]); f.dst_cap = dst_cap[i];
            f.blob = (uint32_t)i;
            f.range_base = (uint32_t)total_ranges;
            f.n_ranges = (uint32_t)((src_len[i] + MF_F_RANGE_BYTES - 1) / MF_F_RANGE_BYTES);
            f.blk_base = (uint32_t)(total_ranges + (uint64_t)fbl.size() * MF_F_EXTRA_BLOCKS);
            f.out16_off = out16_syms;
            total_ranges += f.n_ranges; out16_syms += dst_cap[i] + 8;
            fbl.push_back(f);
        }
    const size_t nf = fbl.size();
    const size_t o_fb = align_up(sizeof(MfDecBlob) * n, 256);
    CU(W.h_desc.ensure(o_fb + sizeof(MfFBlob) * nf));
    MfDecBlob *hb = W.h_desc.as<MfDecBlob>();
    for (uint64_t i = 0; i < n; i++) {
        hb[i].src = static_cast<const uint8_t *>(d_src[i]); hb[i].src_len = src_len[i];
        hb[i].dst = static_cast<uint8_t *>(d_dst[i]); hb[i].dst_cap = dst_cap[i];
    }
    if (nf) memcpy(W.h_desc.as<uint8_t>() + o_fb, fbl.data(), sizeof(MfFBlob) * nf);
    CU(W.d_desc.ensure(o_fb + sizeof(MfFBlob) * nf));
    Carver wc;
    const size_t o_chunks = wc.take(sizeof(MfDecChunk) * max_chunks);
    const size_t o_members = wc.take(sizeof(MfDecMember) * max_members);
    const size_t o_crc = wc.take(sizeof(uint32_t) * max_chunks);
    const size_t o_kind = wc.take(n);
    const size_t o_cnt = wc.take(sizeof(MfDecCounters));
    const size_t o_res = wc.take(0);
    Carver rc;
    W.r_len = rc.take(sizeof(uint64_t) * n, 16);
    W.r_st = rc.take(sizeof(int32_t) * n, 16);
    W.r_sha = 0;
    W.res_bytes = rc.take(0, 16);
    CU(W.d_work.ensure(o_res + W.res_bytes));
    CU(W.h_res.ensure(W.res_bytes));
    uint8_t *dw = W.d_work.as<uint8_t>();
    uint8_t *dres = dw + o_res;
    cudaStream_t s = W.stream;
    CU(cudaMemcpyAsync(W.d_desc.p, hb, o_fb + sizeof(MfFBlob) * nf, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(dw + o_kind, 0, (o_res + W.res_bytes) - o_kind, s));     // kind, counters, dst_len, status
    if (upload) {
        if (phase_timing_on()) CU(cudaEventRecord(W.tl_start, s));
        CU(upload->run(cudaMemcpyHostToDevice, s, &D.h2d));
        if (phase_timing_on()) CU(cudaEventRecord(W.tl_in, s));
    }
    const MfDecBlob *db = W.d_desc.as<MfDecBlob>();
    MfDecCounters *cnt = reinterpret_cast<MfDecCounters *>(dw + o_cnt);
    CU(cudaEventRecord(W.ev[5], s));
    mf_parse_kernel<<<(unsigned)n, 32, 0, s>>>(db, (uint32_t)n, reinterpret_cast<MfDecChunk *>(dw + o_chunks), (uint32_t)max_chunks,
                                               reinterpret_cast<MfDecMember *>(dw + o_members), (uint32_t)max_members, cnt,
                                               dw + o_kind, reinterpret_cast<uint64_t *>(dres + W.r_len), reinterpret_cast<int32_t *>(dres + W.r_st));
    CU(cudaGetLastError());
    static const int dec_occ = getenv("MFCAS_DEC_OCC") ? atoi(getenv("MFCAS_DEC_OCC")) : MF_DEC_DEFAULT_OCC;
    static const int dec_crc = getenv("MFCAS_DEC_CRC") ? atoi(getenv("MFCAS_DEC_CRC")) : MF_DEC_DEFAULT_CRC;
    auto dec_kernel = dec_occ == 5 ? mf_inflate_chunk_kernel<5, 0> : (dec_crc ? mf_inflate_chunk_kernel<4, 1> : mf_inflate_chunk_kernel<4, 0>);
    static const unsigned dec_group = getenv("MFCAS_DEC_GROUP") ? (unsigned)std::max(1, atoi(getenv("MFCAS_DEC_GROUP"))) : MF_DEC_GROUP;
    // resident warps take groups of chunks by ticket: a grid that fills the SMs once (B200: 148 SMs), or less for a small call
    const unsigned dec_grid = (unsigned)std::min<uint64_t>((max_chunks + (uint64_t)MF_DEC_WARPS * dec_group - 1) / ((uint64_t)MF_DEC_WARPS * dec_group),
                                                         148ull * (dec_occ == 5 ? 5u : 4u));
    dec_kernel<<<std::max(dec_grid, 1u), 32 * MF_DEC_WARPS, MF_DEC_SMEM_BYTES, s>>>(
        reinterpret_cast<const MfDecChunk *>(dw + o_chunks), cnt, reinterpret_cast<uint32_t *>(dw + o_crc), reinterpret_cast<int32_t *>(dres + W.r_st), dec_group);
    CU(cudaGetLastError());
    CU(cudaEventRecord(W.ev[6], s));
    mf_verify_kernel<<<(unsigned)((max_members * 32 + 127) / 128), 128, 0, s>>>(
        reinterpret_cast<const MfDecMember *>(dw + o_members), cnt, reinterpret_cast<const uint32_t *>(dw + o_crc), reinterpret_cast<int32_t *>(dres + W.r_st));
    CU(cudaGetLastError());
    if (nf) {
        // ---- block-parallel decoding of the foreign streams (scan, size, walk, decode to 16-bit symbols, resolve, verify);
        // whatever it cannot follow keeps kind 1 and is decoded by the serial kernel below
        const uint64_t blk_slots = total_ranges + (uint64_t)nf * MF_F_EXTRA_BLOCKS;
        Carver fc;
        const size_t f_cnt = fc.take(sizeof(uint32_t) * total_ranges);
        const size_t f_pos = fc.take(sizeof(uint64_t) * total_ranges * MF_F_RANGE_SLOTS);
        const size_t f_cand = fc.take(sizeof(MfSpecCand) * total_ranges);
        const size_t f_blk = fc.take(sizeof(MfSpecBlock) * blk_slots);
        const size_t f_mem = fc.take(sizeof(MfSpecMember) * nf * MF_F_MAX_MEMBERS);
        const size_t f_state = fc.take(sizeof(MfFState) * nf);
        const size_t f_out16 = fc.take(sizeof(uint16_t) * out16_syms);
        CU(W.fscratch.ensure(fc.take(0)));
        uint8_t *fs = W.fscratch.as<uint8_t>();
        const MfFBlob *dfb = reinterpret_cast<const MfFBlob *>(W.d_desc.as<uint8_t>() + o_fb);
        MfSpecCand *dcand = reinterpret_cast<MfSpecCand *>(fs + f_cand);
        MfSpecBlock *dblk = reinterpret_cast<MfSpecBlock *>(fs + f_blk);
        MfSpecMember *dmem = reinterpret_cast<MfSpecMember *>(fs + f_mem);
        MfFState *dstate = reinterpret_cast<MfFState *>(fs + f_state);
        uint8_t *dkind = dw + o_kind;
        uint64_t *dlen = reinterpret_cast<uint64_t *>(dres + W.r_len);
        int32_t *dst_status = reinterpret_cast<int32_t *>(dres + W.r_st);
        const size_t lane_smem = MF_F_LANES * sizeof(MfSpecLane);
        CU(cudaMemsetAsync(fs + f_cnt, 0, sizeof(uint32_t) * total_ranges, s));
        mf_fscan_kernel<<<(unsigned)((total_ranges + 127) / 128), 128, 0, s>>>(dfb, (uint32_t)nf, (uint32_t)total_ranges, reinterpret_cast<uint32_t *>(fs + f_cnt),
                                                                                reinterpret_cast<uint64_t *>(fs + f_pos));
        CU(cudaGetLastError());
        mf_fcompact_kernel<<<(unsigned)nf, 1024, 0, s>>>(dfb, reinterpret_cast<const uint32_t *>(fs + f_cnt), reinterpret_cast<const uint64_t *>(fs + f_pos), dcand, dstate);
        CU(cudaGetLastError());
        mf_fsize_kernel<<<(unsigned)((total_ranges + MF_F_LANES - 1) / MF_F_LANES), MF_F_LANES, lane_smem, s>>>(dfb, (uint32_t)nf, (uint32_t)total_ranges, dcand, dstate);
        CU(cudaGetLastError());
        mf_fwalk_kernel<<<(unsigned)nf, 32, 0, s>>>(dfb, dcand, dstate, dblk, dmem, dkind, dlen, dst_status);
        CU(cudaGetLastError());
        mf_fdecode_kernel<<<(unsigned)((blk_slots + MF_F_LANES - 1) / MF_F_LANES), MF_F_LANES, lane_smem, s>>>(dfb, (uint32_t)nf, (uint32_t)blk_slots, dblk, dstate,
                                                                                                                  reinterpret_cast<uint16_t *>(fs + f_out16));
        CU(cudaGetLastError());
        mf_fresolve_kernel<<<(unsigned)nf, 1024, 0, s>>>(dfb, dblk, dmem, dstate, reinterpret_cast<const uint16_t *>(fs + f_out16), dkind, dlen, dst_status);
        CU(cudaGetLastError());
        mf_fverify_kernel<<<(unsigned)nf, 1024, 0, s>>>(dfb, dmem, dstate, dkind, dlen, dst_status);
        CU(cudaGetLastError());
        D.launches += 7;
    }
    mf_gunzip_lane_kernel<<<(unsigned)((n + 1) / 2), 64, 0, s>>>(db, (uint32_t)n, dw + o_kind, reinterpret_cast<uint64_t *>(dres + W.r_len),
                                                                  reinterpret_cast<int32_t *>(dres + W.r_st));
    CU(cudaGetLastError());
    D.launches += 4;
    CU(cudaEventRecord(W.ev[7], s));
    CU(cudaMemcpyAsync(W.h_res.p, dres, W.res_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(W.ev_res, s));
    W.n = n;
    return MFCAS_OK;
}

int unpack_collect(DeviceState &D, WaveCtx &W, uint64_t *dst_len, int32_t *status) {
    CU(cudaEventSynchronize(W.ev_res));
    const uint8_t *hr = W.h_res.as<uint8_t>();
    memcpy(dst_len, hr + W.r_len, sizeof(uint64_t) * W.n);
    memcpy(status, hr + W.r_st, sizeof(int32_t) * W.n);
    cudaEventElapsedTime(&D.ms[3], W.ev[5], W.ev[6]);
    cudaEventElapsedTime(&D.ms[4], W.ev[6], W.ev[7]);
    return MFCAS_OK;
}

// ------------------------------------------------------------------ host-pointer shards: a ring of waves
struct Shard {
    std::vector<uint64_t> idx;     // blob indices of this device, in order
};

struct ShardIO {
    Op op; const Shard *sh;
    const void *const *src; const uint64_t *src_len; void *const *dst; const uint64_t *dst_cap;
    uint64_t *dst_len; uint8_t *sha1; int32_t *status;
};

// stage + enqueue slice [pos, end) of the shard on wave context W
int wave_submit(mfcas_ctx *ctx, DeviceState &D, WaveCtx &W, const ShardIO &io, size_t pos, size_t end,
                const std::vector<uint64_t> &ioff, const std::vector<uint64_t> &ooff, const std::vector<uint64_t> &caps,
                uint64_t in_bytes, uint64_t out_bytes) {
    const size_t wn = end - pos;
    CU(W.in_arena.ensure(in_bytes + 256));
    if (io.op != Op::Hash) CU(W.out_arena.ensure(out_bytes + 256));
    CopyPlan up;
    W.w_src.resize(wn); W.w_dst.resize(wn); W.w_len.resize(wn); W.w_cap = caps; W.w_ooff = ooff;
    W.w_dlen.assign(wn, 0); W.w_st.assign(wn, 0); W.w_sha.resize(20 * wn);
    for (size_t k = 0; k < wn; k++) {
        uint64_t i = io.sh->idx[pos + k];
        W.w_src[k] = W.in_arena.as<char>() + ioff[k];
        W.w_dst[k] = (io.op == Op::Hash) ? nullptr : (void *)(W.out_arena.as<char>() + ooff[k]);
        W.w_len[k] = io.src_len[i];
        up.add(io.src[i], (void *)W.w_src[k], io.src_len[i]);
    }
    int rc;
    std::vector<uint8_t> foreign;
    if (io.op == Op::Unpack) {
        static const bool par = !(getenv("MFCAS_FOREIGN_PARALLEL") && atoi(getenv("MFCAS_FOREIGN_PARALLEL")) == 0);
        foreign.assign(wn, 0);
        for (size_t k = 0; par && k < wn; k++) {
            const uint64_t i = io.sh->idx[pos + k];
            uint32_t rl, nc;
            foreign[k] = mf_is_indexed_member(static_cast<const uint8_t *>(io.src[i]), io.src_len[i], &rl, &nc) ? 0 : 1;
        }
    }
    if (io.op == Op::Unpack) rc = unpack_submit(D, W, wn, W.w_src.data(), W.w_len.data(), W.w_dst.data(), W.w_cap.data(), &up, foreign.data());
    else rc = pack_submit(D, W, ctx->flags, wn, W.w_src.data(), W.w_len.data(), W.w_dst.data(), W.w_cap.data(), io.op == Op::Pack, &up);
    if (rc != MFCAS_OK) return rc;
    if (phase_timing_on()) CU(cudaEventRecord(W.tl_res, W.stream));
    // Output copies are enqueued right behind the kernels, without a host round trip: the whole destination
    // slot (its capacity) is copied -- for packed output a few per cent more than the bytes produced, for decoded
    // output exactly the decoded size -- so that D2H of this wave overlaps H2D and kernels of the next ones.
    if (io.op != Op::Hash) {
        CopyPlan down;
        for (size_t k = 0; k < wn; k++) down.add(io.dst[io.sh->idx[pos + k]], W.w_dst[k], W.w_cap[k]);
        CU(down.run(cudaMemcpyDeviceToHost, W.stream, &D.d2h));
    }
    CU(cudaEventRecord(W.ev_done, W.stream));
    if (phase_timing_on()) CU(cudaEventRecord(W.tl_done, W.stream));
    W.op = io.op; W.pos = pos; W.wn = wn; W.state = 1;
    return MFCAS_OK;
}

// everything of the wave is done -> hand the per-blob results to the caller
// MFCAS_PHASE_TIMING=1: one line per finished wave, milliseconds since the call's first enqueue.
static void timeline_begin(DeviceState &D, const char *what) {
    if (!phase_timing_on()) return;
    if (!D.tl_origin) cudaEventCreate(&D.tl_origin);
    cudaEventRecord(D.tl_origin, D.ring[0].stream);
    fprintf(stderr, "[mfcas timeline] %s (ms): units: h2d_start h2d_end | sha_start sha_end | kern_start kern_end | results | out_d2h_end\n", what);
}
static void timeline_wave(DeviceState &D, WaveCtx &O, bool unpack, bool pack, size_t units) {
    if (!phase_timing_on() || !D.tl_origin) return;
    float a = 0, b = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0, e = 0, f = 0;
    cudaEvent_t t0 = D.tl_origin;
    cudaEventElapsedTime(&a, t0, O.tl_start); cudaEventElapsedTime(&b, t0, O.tl_in);
    if (!unpack) { cudaEventElapsedTime(&c0, t0, O.ev[0]); cudaEventElapsedTime(&c1, t0, O.ev[1]); }
    if (pack) { cudaEventElapsedTime(&d0, t0, O.ev[2]); cudaEventElapsedTime(&d1, t0, O.ev[4]); }
    if (unpack) { cudaEventElapsedTime(&d0, t0, O.ev[5]); cudaEventElapsedTime(&d1, t0, O.ev[7]); }
    cudaEventElapsedTime(&e, t0, O.tl_res); cudaEventElapsedTime(&f, t0, O.tl_done);
    fprintf(stderr, "  %5zu: %6.1f %6.1f | %6.1f %6.1f | %6.1f %6.1f | %6.1f | %6.1f\n", units, a, b, c0, c1, d0, d1, e, f);
    (void)cudaGetLastError();
}

int wave_finish(DeviceState &D, WaveCtx &W, const ShardIO &io) {
    int rc = (W.op == Op::Unpack) ? unpack_collect(D, W, W.w_dlen.data(), W.w_st.data())
                                  : pack_collect(D, W, W.w_dlen.data(), W.w_sha.data(), W.w_st.data());
    if (rc != MFCAS_OK) return rc;
    CU(cudaEventSynchronize(W.ev_done));
    timeline_wave(D, W, W.op == Op::Unpack, W.op == Op::Pack, W.wn);
    if (W.op != Op::Hash) {
        for (size_t k = 0; k < W.wn; k++) {
            uint64_t i = io.sh->idx[W.pos + k];
            io.dst_len[i] = W.w_dlen[k]; io.status[i] = W.w_st[k];
        }
    } else if (io.status) {
        for (size_t k = 0; k < W.wn; k++) io.status[io.sh->idx[W.pos + k]] = 0;
    }
    if (io.sha1 && W.op != Op::Unpack)
        for (size_t k = 0; k < W.wn; k++) memcpy(io.sha1 + 20 * io.sh->idx[W.pos + k], W.w_sha.data() + 20 * k, 20);
    W.state = 0;
    return MFCAS_OK;
}

int wave_drain(DeviceState &D, WaveCtx &W, const ShardIO &io) {
    if (W.state == 1) return wave_finish(D, W, io);
    return MFCAS_OK;
}

int run_shard(mfcas_ctx *ctx, DeviceState &D, const ShardIO &io) {
    const Shard &sh = *io.sh;
    const uint64_t wave_bytes = ctx->wave_bytes;
    CU(cudaSetDevice(D.dev));
    timeline_begin(D, io.op == Op::Unpack ? "unpack" : "pack");
    size_t pos = 0, w = 0;
    std::vector<uint64_t> ioff, ooff, caps;
    int rc = MFCAS_OK;
    while (pos < sh.idx.size() && rc == MFCAS_OK) {
        // ---- choose the wave: blobs until the staged bytes (in + out) exceed wave_bytes
        size_t end = pos; uint64_t in_bytes = 0, out_bytes = 0;
        ioff.clear(); ooff.clear(); caps.clear();
        while (end < sh.idx.size()) {
            uint64_t i = sh.idx[end];
            uint64_t il = align_up(io.src_len[i], 256) + 256;
            uint64_t oc = 0;
            if (io.op == Op::Pack) oc = std::min<uint64_t>(mf_pack_bound(io.src_len[i]), io.dst_cap[i]);
            else if (io.op == Op::Unpack) oc = io.dst_cap[i];
            uint64_t ol = align_up(oc, 256) + 256;
            if (end > pos && in_bytes + il + out_bytes + ol > wave_bytes) break;
            ioff.push_back(in_bytes); ooff.push_back(out_bytes); caps.push_back(oc);
            in_bytes += il; out_bytes += ol; end++;
        }
        WaveCtx &W = D.ring[w % D.ring.size()];
        rc = wave_drain(D, W, io);                                   // its previous wave (ring full): finish it
        if (rc == MFCAS_OK) rc = wave_submit(ctx, D, W, io, pos, end, ioff, ooff, caps, in_bytes, out_bytes);
        (void)cudaGetLastError();
        pos = end; w++;
    }
    // drain in submission order
    for (size_t k = 0; k < D.ring.size(); k++) {
        WaveCtx &O = D.ring[(w + k) % D.ring.size()];
        int r2 = wave_drain(D, O, io);
        if (rc == MFCAS_OK) rc = r2;
    }
    return rc;
}

// ------------------------------------------------------------------ striped packing (extent output)
// A blob is cut into 2 MiB stripes; every stripe becomes its own gzip member in a fixed slot of the blob's
// destination buffer, and the waves go stripe-major (stripe 0 of every blob, then stripe 1, ...).  The SHA-1
// of a blob is hashed piecewise (state carried in HBM, launches ordered by events), so it progresses while later stripes are still
// on the PCIe bus, and every wave's output goes back as soon as it is encoded.  The packed blob is the
// concatenation of its extents -- a valid multi-member RFC 1952 file.
#define MF_STRIPE_CHUNKS 64u
inline uint64_t stripe_raw() { return (uint64_t)MF_STRIPE_CHUNKS * MF_CHUNK; }
inline uint64_t ext_count(uint64_t len) { return len == 0 ? 1 : (len + stripe_raw() - 1) / stripe_raw(); }
inline uint64_t ext_slot() { return align_up(mf_pack_bound(stripe_raw()), 256); }
inline uint64_t ext_bound(uint64_t len) {
    const uint64_t c = ext_count(len);
    return (c - 1) * ext_slot() + mf_pack_bound(len - (c - 1) * stripe_raw());
}

struct ExtIO {
    const Shard *sh;
    const void *const *src; const uint64_t *src_len; void *const *dst; const uint64_t *dst_cap;
    const uint64_t *ext_base;        // index of blob i's first extent in ext_off / ext_len
    uint64_t *ext_off, *ext_len; uint8_t *sha1; int32_t *status;
};
struct ExtUnit { uint64_t blob; uint32_t m; uint32_t sblob; };

int ext_wave_finish(DeviceState &D, WaveCtx &W, const ExtIO &io, const std::vector<ExtUnit> &units) {
    int rc = pack_collect(D, W, W.w_dlen.data(), W.w_sha.data(), W.w_st.data());
    if (rc != MFCAS_OK) return rc;
    CU(cudaEventSynchronize(W.ev_done));
    timeline_wave(D, W, false, true, W.wn);
    for (size_t k = 0; k < W.wn; k++) {
        const ExtUnit &u = units[W.pos + k];
        io.ext_len[io.ext_base[u.blob] + u.m] = W.w_dlen[k];
        if (W.w_st[k] != 0 && io.status[u.blob] == 0) io.status[u.blob] = W.w_st[k];
        if (u.m + 1 == ext_count(io.src_len[u.blob])) memcpy(io.sha1 + 20 * u.blob, W.w_sha.data() + 20 * k, 20);
    }
    W.state = 0;
    return MFCAS_OK;
}

int run_shard_extents(mfcas_ctx *ctx, DeviceState &D, const ExtIO &io) {
    const Shard &sh = *io.sh;
    CU(cudaSetDevice(D.dev));
    // units in stripe-major order; blobs whose destination is too small are reported and left out
    std::vector<ExtUnit> units;
    uint64_t maxc = 0;
    std::vector<uint8_t> skip(sh.idx.size(), 0);
    for (size_t j = 0; j < sh.idx.size(); j++) {
        const uint64_t i = sh.idx[j];
        io.status[i] = 0;
        const uint64_t c = ext_count(io.src_len[i]);
        for (uint64_t m = 0; m < c; m++) { io.ext_off[io.ext_base[i] + m] = m * ext_slot(); io.ext_len[io.ext_base[i] + m] = 0; }
        if (io.dst_cap[i] < ext_bound(io.src_len[i])) { io.status[i] = MFCAS_E_DST_TOO_SMALL; skip[j] = 1; continue; }
        maxc = std::max(maxc, c);
    }
    for (uint64_t m = 0; m < maxc; m++)
        for (size_t j = 0; j < sh.idx.size(); j++)
            if (!skip[j] && m < ext_count(io.src_len[sh.idx[j]])) units.push_back({sh.idx[j], (uint32_t)m, (uint32_t)j});
    timeline_begin(D, "pack, striped");
    CU(D.sha_state.ensure(sh.idx.size() * 5 * sizeof(uint32_t) + 64));
    uint32_t *d_state = D.sha_state.as<uint32_t>();
    const uint64_t wave_bytes = ctx->wave_bytes;
    auto unit_len = [&](const ExtUnit &u) { return std::min<uint64_t>(stripe_raw(), io.src_len[u.blob] - (uint64_t)u.m * stripe_raw()); };
    // A wave never holds two pieces of one blob (it stays inside one stripe index), and the waves of a stripe
    // index are cut evenly, so that no launch is left with a handful of streams: a SHA-1 launch takes as long
    // as its longest piece whatever its width.
    std::vector<size_t> per_wave(maxc, 0);
    {
        std::vector<uint64_t> g_bytes(maxc, 0), g_count(maxc, 0);
        for (const ExtUnit &u : units) {
            const uint64_t len = unit_len(u);
            g_bytes[u.m] += align_up(len, 256) + 256 + align_up(mf_pack_bound(len), 256) + 256; g_count[u.m]++;
        }
        for (uint64_t m = 0; m < maxc; m++) {
            const uint64_t nw = std::max<uint64_t>(1, (g_bytes[m] + wave_bytes - 1) / wave_bytes);
            per_wave[m] = (size_t)((g_count[m] + nw - 1) / nw);
        }
    }
    WaveCtx *pending = nullptr;
    auto enqueue_down = [&](WaveCtx &O) -> int {
        CU(cudaEventSynchronize(O.ev_res));
        const uint64_t *lens = reinterpret_cast<const uint64_t *>(O.h_res.as<uint8_t>() + O.r_len);
        uint64_t maxlen = 0;
        for (size_t k = 0; k < O.wn; k++) maxlen = std::max(maxlen, std::min<uint64_t>(lens[k], O.w_cap[k]));
        maxlen = align_up(maxlen, 256);
        CopyPlan down;
        for (size_t k = 0; k < O.wn; k++) {
            const ExtUnit &u = units[O.pos + k];
            down.items.push_back({static_cast<char *>(io.dst[u.blob]) + (uint64_t)u.m * ext_slot(), (char *)O.w_dst[k], (size_t)std::min<uint64_t>(maxlen, O.w_cap[k])});
        }
        CU(down.run(cudaMemcpyDeviceToHost, O.stream, &D.d2h));
        CU(cudaEventRecord(O.ev_done, O.stream));
        if (phase_timing_on()) CU(cudaEventRecord(O.tl_done, O.stream));
        O.state = 1;
        pending = nullptr;
        return MFCAS_OK;
    };
    std::vector<long long> last_wave(sh.idx.size(), -1);     // wave that hashed the previous piece of each blob
    std::vector<MfHashDesc> chain;
    std::vector<cudaEvent_t> deps;
    size_t pos = 0, w = 0;
    int rc = MFCAS_OK;
    while (pos < units.size() && rc == MFCAS_OK) {
        size_t end = pos; uint64_t in_bytes = 0, out_bytes = 0;
        std::vector<uint64_t> ioff, ooff, caps, lens;
        while (end < units.size()) {
            const ExtUnit &u = units[end];
            const uint64_t len = unit_len(u);
            const uint64_t il = align_up(len, 256) + 256, oc = mf_pack_bound(len), ol = align_up(oc, 256) + 256;
            if (end > pos && (in_bytes + il + out_bytes + ol > wave_bytes || u.m != units[pos].m || end - pos >= per_wave[u.m])) break;
            ioff.push_back(in_bytes); ooff.push_back(out_bytes); caps.push_back(oc); lens.push_back(len);
            in_bytes += il; out_bytes += ol; end++;
        }
        const size_t wn = end - pos;
        WaveCtx &W = D.ring[w % D.ring.size()];
        if (W.state == 1) rc = ext_wave_finish(D, W, io, units);
        if (rc != MFCAS_OK) break;
        CU(W.in_arena.ensure(in_bytes + 256));
        CU(W.out_arena.ensure(out_bytes + 256));
        CopyPlan up;
        W.w_src.resize(wn); W.w_dst.resize(wn); W.w_len = lens; W.w_cap = caps;
        W.w_dlen.assign(wn, 0); W.w_st.assign(wn, 0); W.w_sha.resize(20 * wn);
        chain.resize(wn);
        for (size_t k = 0; k < wn; k++) {
            const ExtUnit &u = units[pos + k];
            const uint64_t off = (uint64_t)u.m * stripe_raw();
            W.w_src[k] = W.in_arena.as<char>() + ioff[k];
            W.w_dst[k] = W.out_arena.as<char>() + ooff[k];
            up.add(static_cast<const char *>(io.src[u.blob]) + off, (void *)W.w_src[k], lens[k]);
            MfHashDesc &h = chain[k];
            h.src = static_cast<const uint8_t *>(W.w_src[k]);
            h.len = lens[k];
            h.out_index = (uint32_t)k;
            h.flags = (u.m == 0 ? MF_SHA_FIRST : 0u) | (u.m + 1 == ext_count(io.src_len[u.blob]) ? MF_SHA_LAST : 0u);
            h.total_len = io.src_len[u.blob];
            h.state = d_state + 5ull * u.sblob;
        }
        deps.clear();
        {
            std::vector<long long> seen;
            for (size_t k = 0; k < wn; k++) {
                const uint32_t sb = units[pos + k].sblob;
                const long long lw = last_wave[sb];
                last_wave[sb] = (long long)w;
                if (lw < 0 || lw + (long long)D.ring.size() <= (long long)w) continue;      // none, or already collected
                if (std::find(seen.begin(), seen.end(), lw) == seen.end()) { seen.push_back(lw); deps.push_back(D.ring[lw % D.ring.size()].ev_sha); }
            }
        }
        rc = pack_submit(D, W, ctx->flags, wn, W.w_src.data(), W.w_len.data(), W.w_dst.data(), W.w_cap.data(), true, &up, chain.data(), &deps, PackUnits);
        if (rc != MFCAS_OK) break;
        if (phase_timing_on()) CU(cudaEventRecord(W.tl_res, W.stream));
        W.op = Op::Pack; W.pos = pos; W.wn = wn; W.state = 3;          // kernels in flight, copy-back not enqueued yet
        // the copy-back of the PREVIOUS wave goes out now, with the lengths its encoder produced (one small host
        // round trip per wave, hidden behind this wave's upload): packed bytes, not slot capacities, cross the bus
        if (pending) { rc = enqueue_down(*pending); if (rc != MFCAS_OK) break; }
        pending = &W;
        pos = end; w++;
    }
    if (pending && pending->state == 3) { int r2 = enqueue_down(*pending); if (rc == MFCAS_OK) rc = r2; }
    for (size_t k = 0; k < D.ring.size(); k++) {
        WaveCtx &O = D.ring[(w + k) % D.ring.size()];
        if (O.state == 1) { int r2 = ext_wave_finish(D, O, io, units); if (rc == MFCAS_OK) rc = r2; }
    }
    return rc;
}

int run_batch(mfcas_ctx *ctx, Op op, uint64_t n, const void *const *src, const uint64_t *src_len,
              void *const *dst, const uint64_t *dst_cap, uint64_t *dst_len, uint8_t *sha1, int32_t *status) {
    if (!ctx) return MFCAS_E_ARG;
    if (n == 0) return MFCAS_OK;
    if (!src || !src_len) return MFCAS_E_ARG;
    if (op != Op::Hash && (!dst || !dst_cap || !dst_len || !status)) return MFCAS_E_ARG;
    if (op != Op::Unpack && !sha1) return MFCAS_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t G = ctx->devs.size();
    std::vector<Shard> shards(G);
    for (uint64_t i = 0; i < n; i++) shards[i % G].idx.push_back(i);     // round-robin: blob i -> device i mod G
    std::vector<int> rcs(G, MFCAS_OK);
    auto work = [&](size_t g) {
        if (G > 1 && ctx->numa) (void)mfcas_bind_thread_to_device(ctx->devs[g].dev);
        ShardIO io{op, &shards[g], src, src_len, dst, dst_cap, dst_len, sha1, status};
        rcs[g] = run_shard(ctx, ctx->devs[g], io);
    };
    if (G == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(work, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; g++)
        if (rcs[g] != MFCAS_OK) { ctx->err = ctx->devs[g].err; return rcs[g]; }
    return MFCAS_OK;
}

}  // namespace

// ------------------------------------------------------------------ NUMA placement
// The end-to-end path is bound by host memory and PCIe: on a two-socket box a rank whose page-locked buffers sit on
// the other socket pushes every byte across the socket interconnect twice (round 1: e2e scaling 0.56 at 8 GPUs).
static int device_numa_node(int dev) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static bool node_cpu_set(int node, cpu_set_t *set) {
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a = 0, b = 0, n = 0;
    char sep = 0;
    while (fscanf(f, "%d", &a) == 1) {
        b = a;
        int c = fgetc(f);
        if (c == '-') { if (fscanf(f, "%d", &b) != 1) break; c = fgetc(f); }
        for (int k = a; k <= b && k < CPU_SETSIZE; k++) { CPU_SET(k, set); n++; }
        sep = (char)c;
        if (sep != ',') break;
    }
    fclose(f);
    return n > 0;
}

// ====================================================================== C ABI
extern "C" {

int mfcas_bind_thread_to_device(int device) {
    const int node = device_numa_node(device);
    if (node < 0) return -1;
    cpu_set_t want, have, both;
    if (node_cpu_set(node, &want) && sched_getaffinity(0, sizeof have, &have) == 0) {
        CPU_AND(&both, &want, &have);
        if (CPU_COUNT(&both) > 0) (void)sched_setaffinity(0, sizeof both, &both);
    }
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof mask * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        (void)syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
    }
    return node;
}

const char *mfcas_version(void) { return MFCAS_VERSION_STR; }

int mfcas_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

mfcas_ctx *mfcas_create(const int *devices, int ndev, uint64_t wave_bytes, uint32_t flags) {
    // Every in-flight wave has its own streams; with the default of 8 hardware work queues they alias and
    // serialise (wave k+4's H2D waits for wave k's SHA-1).  Effective only if set before the CUDA context exists.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int avail = mfcas_device_count();
    if (avail <= 0) return nullptr;
    if (ndev <= 0) ndev = avail;
    mfcas_ctx *ctx = new (std::nothrow) mfcas_ctx();
    if (!ctx) return nullptr;
    ctx->flags = flags;
    if (wave_bytes) ctx->wave_bytes = std::max<uint64_t>(wave_bytes, 1ull << 20);
    ctx->devs.resize((size_t)ndev);
    // one device: the calling thread (which stages the page-locked buffers) moves next to it; several devices: the
    // per-device worker threads do (run_batch).  MFCAS_NUMA=0 leaves affinity and memory policy alone.
    const char *numa_env = getenv("MFCAS_NUMA");
    ctx->numa = !(numa_env && atoi(numa_env) == 0);
    if (ctx->numa && ndev == 1) (void)mfcas_bind_thread_to_device(devices ? devices[0] : 0);
    for (int g = 0; g < ndev; g++) {
        int d = devices ? devices[g] : g;
        if (d < 0 || d >= avail) { mfcas_destroy(ctx); return nullptr; }
        ctx->devs[(size_t)g].dev = d;
        if (device_setup(ctx->devs[(size_t)g]) != MFCAS_OK) {
            fprintf(stderr, "mfcas_create: %s\n", ctx->devs[(size_t)g].err.c_str());
            mfcas_destroy(ctx);
            return nullptr;
        }
    }
    return ctx;
}

void mfcas_destroy(mfcas_ctx *ctx) {
    if (!ctx) return;
    for (auto &D : ctx->devs) device_teardown(D);
    delete ctx;
}

const char *mfcas_last_error(mfcas_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

void *mfcas_alloc_pinned(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
void mfcas_free_pinned(void *p) { if (p) cudaFreeHost(p); }

uint64_t mfcas_pack_bound(uint64_t src_len) { return mf_pack_bound(src_len); }

int mfcas_pack_batch(mfcas_ctx *ctx, uint64_t n, const void *const *src, const uint64_t *src_len,
                     void *const *dst, const uint64_t *dst_cap, uint64_t *dst_len, uint8_t *sha1, int32_t *status) {
    return run_batch(ctx, Op::Pack, n, src, src_len, dst, dst_cap, dst_len, sha1, status);
}

uint64_t mfcas_pack_extent_count(uint64_t src_len) { return ext_count(src_len); }
uint64_t mfcas_pack_extent_bound(uint64_t src_len) { return ext_bound(src_len); }

int mfcas_pack_batch_extents(mfcas_ctx *ctx, uint64_t n, const void *const *src, const uint64_t *src_len,
                             void *const *dst, const uint64_t *dst_cap, uint64_t *ext_off, uint64_t *ext_len,
                             uint8_t *sha1, int32_t *status) {
    if (!ctx) return MFCAS_E_ARG;
    if (n == 0) return MFCAS_OK;
    if (!src || !src_len || !dst || !dst_cap || !ext_off || !ext_len || !sha1 || !status) return MFCAS_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t G = ctx->devs.size();
    std::vector<Shard> shards(G);
    std::vector<uint64_t> base(n);
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n; i++) { shards[i % G].idx.push_back(i); base[i] = acc; acc += ext_count(src_len[i]); }
    std::vector<int> rcs(G, MFCAS_OK);
    auto work = [&](size_t g) {
        if (G > 1 && ctx->numa) (void)mfcas_bind_thread_to_device(ctx->devs[g].dev);
        ExtIO io{&shards[g], src, src_len, dst, dst_cap, base.data(), ext_off, ext_len, sha1, status};
        rcs[g] = run_shard_extents(ctx, ctx->devs[g], io);
    };
    if (G == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(work, g);
        for (auto &t : th) t.join();
    }
    for (size_t g = 0; g < G; g++)
        if (rcs[g] != MFCAS_OK) { ctx->err = ctx->devs[g].err; return rcs[g]; }
    return MFCAS_OK;
}

int mfcas_hash_batch(mfcas_ctx *ctx, uint64_t n, const void *const *src, const uint64_t *src_len, uint8_t *sha1, int32_t *status) {
    return run_batch(ctx, Op::Hash, n, src, src_len, nullptr, nullptr, nullptr, sha1, status);
}

int mfcas_unpack_batch(mfcas_ctx *ctx, uint64_t n, const void *const *src, const uint64_t *src_len,
                       void *const *dst, const uint64_t *dst_cap, uint64_t *dst_len, int32_t *status) {
    return run_batch(ctx, Op::Unpack, n, src, src_len, dst, dst_cap, dst_len, nullptr, status);
}

int mfcas_unpacked_size(const void *src, uint64_t src_len, uint64_t *size, int *exact) {
    if (!src || !size) return MFCAS_E_ARG;
    const uint8_t *p = static_cast<const uint8_t *>(src);
    uint64_t ip = 0, total = 0;
    int members = 0;
    if (exact) *exact = 0;
    while (ip < src_len) {
        if (members) { while (ip < src_len && p[ip] == 0) ip++; if (ip >= src_len) break; }
        uint32_t rl, nc;
        uint64_t hdr = mf_is_indexed_member(p + ip, src_len - ip, &rl, &nc);
        if (!hdr) {
            // foreign stream: ISIZE of the last member is the best hint there is without decoding
            if (src_len < 18 || p[0] != 0x1f || p[1] != 0x8b) return MFCAS_E_BAD_STREAM;
            *size = total + mf_le32(p + src_len - 4);
            return MFCAS_OK;
        }
        uint64_t payload = 0;
        for (uint32_t k = 0; k < nc; k++) payload += mf_le32(p + ip + MF_HDR_FIXED + 8ull * k);
        if (ip + hdr + payload + MF_TRAILER > src_len) return MFCAS_E_TRUNC;
        total += rl; members++;
        ip += hdr + payload + MF_TRAILER;
    }
    if (!members) return MFCAS_E_TRUNC;
    *size = total;
    if (exact) *exact = 1;
    return MFCAS_OK;
}

int mfcas_pack_device(mfcas_ctx *ctx, int dev_slot, uint64_t n, const void *const *d_src, const uint64_t *src_len,
                      void *const *d_dst, const uint64_t *dst_cap, uint64_t *dst_len, uint8_t *sha1, int32_t *status) {
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return MFCAS_E_ARG;
    if (n && (!d_src || !src_len || !d_dst || !dst_cap || !dst_len || !sha1 || !status)) return MFCAS_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceState &D = ctx->devs[(size_t)dev_slot];
    if (n == 0) return MFCAS_OK;
    cudaSetDevice(D.dev);
    WaveCtx &W = D.ring[0];
    int rc = pack_submit(D, W, ctx->flags, n, d_src, src_len, d_dst, dst_cap, true, nullptr);
    if (rc == MFCAS_OK) rc = pack_collect(D, W, dst_len, sha1, status);
    if (rc != MFCAS_OK) ctx->err = D.err;
    return rc;
}

int mfcas_pack_device_extents(mfcas_ctx *ctx, int dev_slot, uint64_t n, const void *const *d_src, const uint64_t *src_len,
                              void *const *d_dst, const uint64_t *dst_cap, uint64_t *ext_off, uint64_t *ext_len, uint8_t *sha1, int32_t *status) {
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return MFCAS_E_ARG;
    if (n && (!d_src || !src_len || !d_dst || !dst_cap || !ext_off || !ext_len || !sha1 || !status)) return MFCAS_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceState &D = ctx->devs[(size_t)dev_slot];
    if (n == 0) return MFCAS_OK;
    cudaSetDevice(D.dev);
    WaveCtx &W = D.ring[0];
    uint64_t n_members = 0;
    for (uint64_t i = 0; i < n; i++) n_members += ext_count(src_len[i]);
    std::vector<uint64_t> mlen(n_members);
    std::vector<int32_t> mst(n_members);
    int rc = pack_submit(D, W, ctx->flags, n, d_src, src_len, d_dst, dst_cap, true, nullptr, nullptr, nullptr, PackSlots);
    if (rc == MFCAS_OK) rc = pack_collect(D, W, mlen.data(), sha1, mst.data());
    if (rc != MFCAS_OK) { ctx->err = D.err; return rc; }
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; i++) {
        status[i] = 0;
        const uint64_t c = ext_count(src_len[i]);
        for (uint64_t m = 0; m < c; m++, k++) {
            ext_off[k] = m * ext_slot();
            ext_len[k] = mst[k] ? 0 : mlen[k];
            if (mst[k] && !status[i]) status[i] = mst[k];
        }
    }
    return MFCAS_OK;
}

int mfcas_compact_device(mfcas_ctx *ctx, int dev_slot, uint64_t n, void *const *d_dst, const uint64_t *src_len, const uint64_t *ext_len,
                         uint64_t *dst_len) {
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return MFCAS_E_ARG;
    if (n && (!d_dst || !src_len || !ext_len || !dst_len)) return MFCAS_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceState &D = ctx->devs[(size_t)dev_slot];
    if (n == 0) return MFCAS_OK;
    CU(cudaSetDevice(D.dev));
    WaveCtx &W = D.ring[0];
    std::vector<uint64_t> base(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) base[i + 1] = base[i] + ext_count(src_len[i]);
    const uint64_t nm = base[n];
    Carver hc;
    const size_t o_dst = hc.take(sizeof(void *) * n);
    const size_t o_base = hc.take(sizeof(uint64_t) * (n + 1));
    const size_t o_len = hc.take(sizeof(uint64_t) * nm);
    const size_t o_out = hc.take(sizeof(uint64_t) * n);
    const size_t bytes = hc.take(0);
    CU(W.h_desc.ensure(bytes));
    CU(W.d_desc.ensure(bytes));
    uint8_t *hd = W.h_desc.as<uint8_t>(), *dd = W.d_desc.as<uint8_t>();
    memcpy(hd + o_dst, d_dst, sizeof(void *) * n);
    memcpy(hd + o_base, base.data(), sizeof(uint64_t) * (n + 1));
    memcpy(hd + o_len, ext_len, sizeof(uint64_t) * nm);
    cudaStream_t s = W.stream;
    CU(cudaMemcpyAsync(dd, hd, o_out, cudaMemcpyHostToDevice, s));
    mf_compact_inplace_kernel<<<(unsigned)n, 1024, 0, s>>>(reinterpret_cast<uint8_t *const *>(dd + o_dst), reinterpret_cast<const uint64_t *>(dd + o_base),
                                                           reinterpret_cast<const uint64_t *>(dd + o_len), ext_slot(), reinterpret_cast<uint64_t *>(dd + o_out));
    CU(cudaGetLastError());
    D.launches++;
    CU(cudaMemcpyAsync(hd + o_out, dd + o_out, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(dst_len, hd + o_out, sizeof(uint64_t) * n);
    return MFCAS_OK;
}

int mfcas_unpack_device(mfcas_ctx *ctx, int dev_slot, uint64_t n, const void *const *d_src, const uint64_t *src_len,
                        void *const *d_dst, const uint64_t *dst_cap, uint64_t *dst_len, int32_t *status) {
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return MFCAS_E_ARG;
    if (n && (!d_src || !src_len || !d_dst || !dst_cap || !dst_len || !status)) return MFCAS_E_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceState &D = ctx->devs[(size_t)dev_slot];
    if (n == 0) return MFCAS_OK;
    cudaSetDevice(D.dev);
    WaveCtx &W = D.ring[0];
    int rc = unpack_submit(D, W, n, d_src, src_len, d_dst, dst_cap, nullptr);
    if (rc == MFCAS_OK) rc = unpack_collect(D, W, dst_len, status);
    if (rc != MFCAS_OK) ctx->err = D.err;
    return rc;
}

void *mfcas_stream(mfcas_ctx *ctx, int dev_slot) {
    if (!ctx || dev_slot < 0 || (size_t)dev_slot >= ctx->devs.size()) return nullptr;
    return (void *)ctx->devs[(size_t)dev_slot].ring[0].stream;
}

void mfcas_counters(mfcas_ctx *ctx, uint64_t *kernel_launches, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
    uint64_t l = 0, a = 0, b = 0;
    if (ctx) for (auto &D : ctx->devs) { l += D.launches; a += D.h2d; b += D.d2h; }
    if (kernel_launches) *kernel_launches = l;
    if (h2d_bytes) *h2d_bytes = a;
    if (d2h_bytes) *d2h_bytes = b;
}

void mfcas_last_kernel_ms(mfcas_ctx *ctx, float out[8]) {
    for (int i = 0; i < 8; i++) out[i] = 0.f;
    if (!ctx) return;
    for (auto &D : ctx->devs) for (int i = 0; i < 8; i++) out[i] = std::max(out[i], D.ms[i]);
}

}  // extern "C"
