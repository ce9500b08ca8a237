/* include/ljb_comm.h — multi-GPU entry points for a C host, in lz4-jpeg_b200/libljb_comm.so.
 *
 * One process drives all GPUs of a node: a context per GPU (ljb_ctx, include/lz4jpeg_b200.h) and one NCCL communicator per GPU
 * (ncclCommInitAll).  Both codecs shard by independent units — LZ4 blocks (the reference's thread per block,
 * Algorithms/parallel/LZ4/LZ4.c:724-749), JPEG group rows (thread per group, Algorithms/parallel/JPEG/JPEG.c:1297-1302) — so
 * every GPU encodes its own contiguous range into its own buffer and the ONLY collective is one ncclAllGather of the per-GPU
 * byte totals over NVLink, whose exclusive scan gives every shard its base offset in the global stream (what the reference's
 * serial write_output, LZ4.c:427-441, does implicitly).  This library is the only part of the package that links NCCL.
 */
#ifndef LJB_COMM_H
#define LJB_COMM_H

#include <stddef.h>
#include <stdint.h>

#include "lz4jpeg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ljb_comm ljb_comm;

/* ngpus contexts on devices 0 .. ngpus-1 and their NCCL communicators.  LJB_E_ARG if the node has fewer GPUs. */
int ljb_comm_create(int ngpus, ljb_comm **out);
void ljb_comm_destroy(ljb_comm *comm);
int ljb_comm_size(const ljb_comm *comm);
ljb_ctx *ljb_comm_ctx(ljb_comm *comm, int rank); /* the context of GPU `rank` (for the single-GPU entry points) */

/* The collective: d_totals[r] points at one device uint64 on GPU r (an encode kernel's d_result[0]); after the all-gather
 * bases[r] = sum of the totals of ranks < r and *grand = the sum of all (host values).  Ordered after the work already
 * enqueued on every context's stream. */
int ljb_comm_gather_totals(ljb_comm *comm, const uint64_t *const *d_totals, uint64_t *bases, uint64_t *grand);

/* ljb_lz4_compress over all GPUs: GPU r encodes blocks [r * ceil(nblocks / ngpus), ...) (upload, kernel and download of the
 * GPUs run concurrently), the totals are all-gathered, and every shard lands at its base offset of `out`.  The result is byte
 * for byte the single-GPU stream (same arguments as ljb_lz4_compress, include/lz4jpeg_b200.h). */
int ljb_comm_lz4_compress(ljb_comm *comm, const uint8_t *in, size_t n, size_t block_len, uint8_t *out, size_t out_cap,
                          uint64_t *block_offsets, size_t *out_len, uint64_t *phantom);

/* ljb_jpeg_encode_rgba over all GPUs: GPU r encodes a contiguous range of whole group rows of the one image. */
int ljb_comm_jpeg_encode_rgba(ljb_comm *comm, const uint8_t *rgba, int w, int h, size_t stride, uint8_t *out, size_t out_cap,
                              uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len);
/* the same for r g b pixels of three bytes (ljb_jpeg_encode_rgb) */
int ljb_comm_jpeg_encode_rgb(ljb_comm *comm, const uint8_t *rgb, int w, int h, size_t stride, uint8_t *out, size_t out_cap,
                             uint64_t *group_offsets, uint16_t *group_bits, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
