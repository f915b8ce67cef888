/*
 * mfcas.h -- C ABI of the B200-native artifact content-addressed-store codec.
 *
 * This is the drop-in boundary for ONE hot path of Netflix/metaflow 2.19.24:
 *
 *   TaskDataStore.save_artifacts / load_artifacts     metaflow/datastore/task_datastore.py:329-383, 386-442
 *     -> ContentAddressedStore.save_blobs / load_blobs metaflow/datastore/content_addressed_store.py:41-120, 122-204
 *        -> sha1(blob).hexdigest()                    content_addressed_store.py:98      (hashlib / OpenSSL)
 *        -> _pack_v1:  gzip.GzipFile(level 3).write   content_addressed_store.py:211-216 (zlib deflate + crc32)
 *        -> _unpack_v1: gzip.GzipFile("rb").read      content_addressed_store.py:218-220 (zlib inflate + crc32)
 *
 * The reference is pure Python and has no FFI of its own; the entry points below are
 * exactly what a ctypes binding inserted at those three call sites would bind (see
 * INTEGRATION.md for the stub).  Plain pointers and sizes only; no torch / CUDA types.
 * All functions are thread-safe per context; none throws; per-blob outcomes are reported
 * in `status[]`.  There is NO CPU fallback: without a CUDA device mfcas_create fails.
 */
#ifndef MFCAS_H
#define MFCAS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfcas_ctx mfcas_ctx;

/* per-blob status codes (status[i]) and function return codes */
enum {
    MFCAS_OK = 0,
    MFCAS_E_DST_TOO_SMALL = 1, /* dst_cap[i] < bytes needed; dst_len[i] holds the need when known   */
    MFCAS_E_CUDA = 2,          /* a CUDA call failed; see mfcas_last_error()                         */
    MFCAS_E_BAD_STREAM = 3,    /* not gzip / invalid DEFLATE  (reference: BadGzipFile / zlib.error)  */
    MFCAS_E_BAD_CRC = 4,       /* member CRC-32 mismatch      (reference gzip.py:556)                */
    MFCAS_E_BAD_ISIZE = 5,     /* member ISIZE mismatch       (reference gzip.py:559)                */
    MFCAS_E_ARG = 6,           /* bad argument                                                       */
    MFCAS_E_NOMEM = 7,         /* host or device allocation failed                                   */
    MFCAS_E_TRUNC = 8          /* input ends inside a member  (reference: EOFError)                  */
};

/* mfcas_create flags */
enum {
    MFCAS_FLAG_NO_LZ = 1u      /* Huffman-only encoder (no LZ77 matching)                           */
};

/* Number of CUDA devices visible (0 if none / no driver). */
int mfcas_device_count(void);

/* Create a context bound to `ndev` devices (devices == NULL: 0..ndev-1; ndev <= 0: all).
 * `wave_bytes` bounds the raw bytes staged per device per wave (0 = default 1 GiB).
 * Blobs of a batch are sharded round-robin over the context's devices (blob i -> device
 * i mod ndev), replacing the reference's process-per-task parallelism
 * (metaflow/runtime.py:1517-1540) and multicore_utils.parallel_map (multicore_utils.py:140-175).
 * Returns NULL on failure (no device, no memory). */
mfcas_ctx *mfcas_create(const int *devices, int ndev, uint64_t wave_bytes, uint32_t flags);
void mfcas_destroy(mfcas_ctx *ctx);

/* Last error text of this context (valid until the next call on the context). */
const char *mfcas_last_error(mfcas_ctx *ctx);

/* Moves the CALLING THREAD next to CUDA device `device`: CPU affinity restricted to the cores of the device's NUMA
 * node, memory policy "prefer that node" (so page-locked buffers allocated afterwards are local to the GPU's PCIe
 * root).  mfcas_create does this itself for a one-device context unless MFCAS_NUMA=0.  Returns the node, -1 if unknown. */
int mfcas_bind_thread_to_device(int device);

/* Page-locked host memory for zero-staging transfers. */
void *mfcas_alloc_pinned(uint64_t bytes);
void mfcas_free_pinned(void *p);

/* Upper bound of the packed (RFC 1952) size of a blob of `src_len` bytes. */
uint64_t mfcas_pack_bound(uint64_t src_len);

/* save side: replaces sha1() + _pack_v1() of content_addressed_store.py:98,114,211-216.
 * For each blob i: sha1[20*i .. 20*i+20) = SHA-1(src[i]); dst[i][0..dst_len[i]) = a valid
 * RFC 1952 file (cas_version 1) that gzip.GzipFile / the reference's _unpack_v1 decodes to
 * src[i] bit-exactly.  Host pointers (pinned or pageable). */
int mfcas_pack_batch(mfcas_ctx *ctx, uint64_t n,
                     const void *const *src, const uint64_t *src_len,
                     void *const *dst, const uint64_t *dst_cap, uint64_t *dst_len,
                     uint8_t *sha1, int32_t *status);

/* Striped form of mfcas_pack_batch for large blobs.  Every 2 MiB stripe of a blob becomes its own gzip
 * member, written at a fixed slot of dst[i]; the packed blob is the concatenation, in order, of the
 * extents dst[i][ext_off[k] .. ext_off[k] + ext_len[k]) for k = base_i .. base_i + count_i - 1, where
 * count_i = mfcas_pack_extent_count(src_len[i]) and base_i is the sum of the counts of blobs 0..i-1
 * (ext_off / ext_len hold the sum of all counts entries).  dst_cap[i] >= mfcas_pack_extent_bound(src_len[i]).
 * The waves go stripe by stripe across all blobs and the SHA-1 state is carried between them, so hashing,
 * encoding and the copy back overlap the upload instead of trailing the last blob (a writer simply
 * writev()s the extents).  Same digests, same decoded bytes as mfcas_pack_batch. */
uint64_t mfcas_pack_extent_count(uint64_t src_len);
uint64_t mfcas_pack_extent_bound(uint64_t src_len);
int mfcas_pack_batch_extents(mfcas_ctx *ctx, uint64_t n,
                             const void *const *src, const uint64_t *src_len,
                             void *const *dst, const uint64_t *dst_cap,
                             uint64_t *ext_off, uint64_t *ext_len,
                             uint8_t *sha1, int32_t *status);

/* key only: replaces sha1() for raw blobs (save_blobs(raw=True), flow_datastore.py:348-367)
 * and for dedup-before-pack (content_addressed_store.py:107). */
int mfcas_hash_batch(mfcas_ctx *ctx, uint64_t n,
                     const void *const *src, const uint64_t *src_len,
                     uint8_t *sha1, int32_t *status);

/* Decoded size of a packed blob, read from the host copy of the file: exact for files this
 * library wrote (index in the gzip FEXTRA field); for foreign single-member gzip the ISIZE
 * trailer (exact below 4 GiB).  *exact is set to 0 when the value is only a hint. */
int mfcas_unpacked_size(const void *src, uint64_t src_len, uint64_t *size, int *exact);

/* load side: replaces _unpack_v1() of content_addressed_store.py:218-220 -- gunzip of every
 * member with CRC-32 and ISIZE verified.  Decodes files written by this library (chunk-
 * parallel) and by stock Metaflow / any RFC 1952 writer (stream-parallel). */
int mfcas_unpack_batch(mfcas_ctx *ctx, uint64_t n,
                       const void *const *src, const uint64_t *src_len,
                       void *const *dst, const uint64_t *dst_cap, uint64_t *dst_len,
                       int32_t *status);

/* Device-resident variants (src/dst are DEVICE pointers on context device slot `dev_slot`,
 * src 16-byte aligned; dst_len/sha1/status are HOST arrays filled before return).  Used by
 * bench.py for the HBM-resident roofline measurement and by callers that keep artifacts on
 * the GPU.  The work is enqueued on mfcas_stream(ctx, dev_slot) and the call synchronises it.
 * mfcas_unpack_device reads its sources in whole 16-byte quads: d_src[i] must be readable up to the
 * next 16-byte boundary behind d_src[i] + src_len[i] (true of any buffer that ends where a device
 * allocation ends, and of sub-buffers with 15 bytes of slack; the host-path arenas leave 256). */
int mfcas_pack_device(mfcas_ctx *ctx, int dev_slot, uint64_t n,
                      const void *const *d_src, const uint64_t *src_len,
                      void *const *d_dst, const uint64_t *dst_cap, uint64_t *dst_len,
                      uint8_t *sha1, int32_t *status);
/* Device-resident form of mfcas_pack_batch_extents: member m of blob i (its m-th 2 MiB stripe) is written straight
 * to d_dst[i] + ext_off[..] by the encoder -- no scratch, no compaction pass; dst_cap[i] >= mfcas_pack_extent_bound.
 * mfcas_compact_device then moves the extents of every blob together IN PLACE (what writev() into the CAS file does
 * on the host path) so that d_dst[i][0 .. dst_len[i]) is the contiguous RFC 1952 file mfcas_unpack_device takes. */
int mfcas_pack_device_extents(mfcas_ctx *ctx, int dev_slot, uint64_t n,
                              const void *const *d_src, const uint64_t *src_len,
                              void *const *d_dst, const uint64_t *dst_cap,
                              uint64_t *ext_off, uint64_t *ext_len, uint8_t *sha1, int32_t *status);
int mfcas_compact_device(mfcas_ctx *ctx, int dev_slot, uint64_t n, void *const *d_dst, const uint64_t *src_len,
                         const uint64_t *ext_len, uint64_t *dst_len);
int mfcas_unpack_device(mfcas_ctx *ctx, int dev_slot, uint64_t n,
                        const void *const *d_src, const uint64_t *src_len,
                        void *const *d_dst, const uint64_t *dst_cap, uint64_t *dst_len,
                        int32_t *status);

/* cudaStream_t (as void*) the device-resident calls of `dev_slot` run on, for event timing. */
void *mfcas_stream(mfcas_ctx *ctx, int dev_slot);

/* Counters since context creation: kernels launched, bytes moved H2D / D2H. */
void mfcas_counters(mfcas_ctx *ctx, uint64_t *kernel_launches, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/* Milliseconds spent in the encode / decode kernel groups of the most recent device-resident
 * call, measured with CUDA events on the library stream: out[0]=sha1, out[1]=deflate,
 * out[2]=gather, out[3]=inflate, out[4]=verify. */
void mfcas_last_kernel_ms(mfcas_ctx *ctx, float out[8]);

const char *mfcas_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MFCAS_H */
