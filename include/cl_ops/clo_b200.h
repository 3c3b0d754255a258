/*
 * include/cl_ops/clo_b200.h -- additive entry points that have no counterpart in
 * the reference API but that its callers need on this backend (SURVEY.md 8b,
 * "necessary additive exports").
 */
#ifndef CLO_B200_EXT_H
#define CLO_B200_EXT_H

#include <cl_ops/clo_common.h>
#include <cl_ops/clo_rng.h>
#include <cl_ops/clo_sort_abstract.h>
#include <cl_ops/clo_scan_abstract.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bulk generation.  The reference has no host-side generate call; the bulk
 * layout is defined by its benchmark driver, which relaunches a one-number-per-
 * stream kernel `runs` times (src/benchmarks/clo_rng_bench.cl:23-37,
 * src/benchmarks/clo_rng_bench.c:302-324):
 *   out[r * G + g] = next_r(state_g) >> (32 - bits)      when maxint == 0
 *                  = next_r(state_g) % maxint            otherwise
 * for r < runs, g < G = seeds_count.  States are written back, so a second call
 * continues every stream where the first stopped. `out` holds runs*G cl_uint. */
CCLEvent* clo_rng_generate(CloRng* rng, CCLQueue* cq, CCLBuffer* out,
	size_t runs, cl_uint bits, cl_uint maxint, GError** err);
/* Same with a host destination (allocates, generates, copies back, blocks). */
cl_bool clo_rng_generate_host(CloRng* rng, CCLQueue* cq, void* out,
	size_t runs, cl_uint bits, cl_uint maxint, GError** err);
/* Multi-GPU stream partitioning: like clo_rng_new(DEV_GID) but the seeds are
 * those of global work-items [gid_offset, gid_offset + seeds_count)
 * (clo_rng_init.cl:52: seed = get_global_id(0) + main_seed). */
CloRng* clo_rng_new_dev_gid_offset(const char* type, size_t seeds_count,
	cl_ulong gid_offset, cl_ulong main_seed, const char* hash,
	CCLContext* ctx, CCLQueue* cq, GError** err);

/* Key/payload sort with the payload in a separate array (CloType is scalar,
 * so "u64 key + u32 payload" cannot be expressed through clo_sort_new's element
 * type).  Uses the sorter's key type for `keys`; `payload` is cl_uint.  Stable,
 * raw key bits ascending (satradix semantics).  In place (aux buffers cached). */
CCLEvent* clo_sort_pairs_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
	CCLBuffer* keys, CCLBuffer* payload, size_t numel, GError** err);

/* Scan building blocks for the multi-GPU scan (per-GPU total, then a scan
 * with a carry-in held in device memory).  `total_out` / `carry_in` point to one
 * value of the scanner's sum type in device memory; carry_in may be NULL. */
CCLEvent* clo_scan_reduce_with_device_data(CloScan* scanner, CCLQueue* cq_exec,
	CCLBuffer* data_in, CCLBuffer* total_out, size_t numel, GError** err);
CCLEvent* clo_scan_with_device_data_carry(CloScan* scanner, CCLQueue* cq_exec,
	CCLBuffer* data_in, CCLBuffer* data_out, CCLBuffer* carry_in,
	size_t numel, GError** err);

/* Sample-sort building block: stable partition of (keys[, payload]) into
 * `nparts` contiguous buckets by `nparts-1` splitters.  Element i (global index
 * gidx0+i) goes to bucket #{ s : (splitter_key[s], splitter_idx[s]) <= (key_i, gidx0+i) }.
 * counts_out receives nparts cl_ulong bucket sizes (device memory). */
CCLEvent* clo_sort_partition_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
	CCLBuffer* keys_in, CCLBuffer* payload_in, CCLBuffer* keys_out,
	CCLBuffer* payload_out, size_t numel, cl_ulong gidx0,
	CCLBuffer* splitter_keys, CCLBuffer* splitter_idx, cl_uint nparts,
	CCLBuffer* counts_out, GError** err);

/* Sample sort, fused partition + exchange (one process per GPU on one box):
 * stage 1 counts this rank's buckets; the ranks exchange the sizes; stage 2 scatters every
 * bucket straight into the receive buffer of its destination rank over NVLink.
 * dest_ptrs / payload_dest_ptrs: device arrays of nparts raw device addresses (own buffer or
 * peer memory from clo_b200_ipc_import); first_slot[q]: element index in destination q where
 * this rank's bucket q starts; *ok_flag == 0 makes the scatter a no-op. */
CCLEvent* clo_sort_partition_count_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
	CCLBuffer* keys_in, size_t numel, cl_ulong gidx0, CCLBuffer* splitter_keys,
	CCLBuffer* splitter_idx, cl_uint nparts, CCLBuffer* counts_out, GError** err);
CCLEvent* clo_sort_partition_scatter_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
	CCLBuffer* keys_in, CCLBuffer* payload_in, size_t numel, cl_ulong gidx0,
	CCLBuffer* splitter_keys, CCLBuffer* splitter_idx, cl_uint nparts, CCLBuffer* first_slot,
	CCLBuffer* dest_ptrs, CCLBuffer* payload_dest_ptrs, CCLBuffer* ok_flag, GError** err);

/* CUDA IPC for buffers created by ccl_buffer_new: export a 64-byte handle, import it in
 * another process of the same box (peer access is enabled on first use). */
cl_bool clo_b200_ipc_export(CCLBuffer* buf, unsigned char handle[64], GError** err);
CCLBuffer* clo_b200_ipc_import(CCLContext* ctx, const unsigned char handle[64], size_t size, GError** err);

/* Status words of the sorter's last radix call (blocks on the queue): out[0] look-back
 * timeout flag, out[1] number of tiles whose atomic ranks failed verification and were
 * redone with the ballot ranks, out[2..17] optional phase profile (CLO_RADIX_PROFILE=1). */
cl_bool clo_sort_b200_debug(CloSort* sorter, CCLQueue* cq, cl_ulong out[18]);

/* Per-kernel device timing of the sorter's radix calls, taken with CUDA events on the
 * queue's stream (the analogue of cf4ocl's per-event profiling the reference benches use,
 * src/benchmarks/clo_sort_bench.c:201-208).  After a call, get_timing returns
 * out_ms[0] = histogram + bin scan and out_ms[1..] = one value per onesweep pass. */
void clo_sort_b200_set_timing(CloSort* sorter, cl_bool on);
cl_uint clo_sort_b200_get_timing(CloSort* sorter, float* out_ms, cl_uint cap);

/* ---- multi-GPU: one process (or thread) per GPU of one box --------------------------
 * The reference is single-device (clo_sort_abstract.c:335 takes "the first device in the
 * context"); these entry points shard its three operators over the GPUs of a box.  The
 * library moves everything itself over NVLink / NVSwitch through memory mapped with CUDA IPC:
 * the DATA (the partition kernel writes every bucket straight into its destination GPU) and the
 * bookkeeping (sample rows, bucket sizes, scan totals and the closing barrier are small kernels
 * that write into every peer's control buffer and spin, bounded, on epoch flags).  The caller
 * supplies ONE collective: a host all-gather that passes the 64-byte IPC handles round at
 * set-up.  all_gather_dev / barrier_dev are optional: they carry the bookkeeping instead (one
 * NCCL call each) when all_gather_host is missing or CLO_DIST_CTRL=nccl is set. */
typedef struct clo_dist CloDist;
typedef struct clo_dist_comm {
	void* user;
	cl_uint rank, world;                 /* world <= 16 */
	/* optional.  Device all-gather ordered on `cuda_stream`: every rank gives `bytes` bytes at
	 * send_dev and gets world * bytes at recv_dev, in rank order (ncclAllGather).  0 on success. */
	int (*all_gather_dev)(void* user, const void* send_dev, void* recv_dev, size_t bytes, void* cuda_stream);
	/* optional.  Device barrier ordered on `cuda_stream`: what follows it on the stream runs after
	 * everything the other ranks enqueued before their call has completed (a 4-byte ncclAllReduce) */
	int (*barrier_dev)(void* user, void* cuda_stream);
	/* host all-gather of `bytes` bytes per rank (set-up and tear-down only) */
	int (*all_gather_host)(void* user, const void* send, void* recv, size_t bytes);
} CloDistComm;

#define CLO_DIST_GIDX_AUTO (~(cl_ulong) 0)

/* both collective when world > 1 (the control buffers are exchanged / unmapped) */
CloDist* clo_dist_new(CCLContext* ctx, const CloDistComm* comm, GError** err);
void clo_dist_destroy(CloDist* d);

/* Collective.  Allocates this rank's receive buffers (`capacity` elements of key_type, 4 or 8
 * bytes wide, plus cl_uint payloads when with_payload), exchanges their IPC handles and maps
 * every peer's.  capacity must cover the largest slice a rank can receive (the splitters keep
 * slices within a few per cent of numel for any key distribution: 1.25 * numel is ample). */
cl_bool clo_dist_sort_setup(CloDist* d, CloType key_type, size_t capacity, cl_bool with_payload, GError** err);

/* Collective.  Globally stable sort (raw key bits ascending, satradix semantics) of the
 * concatenation of every rank's keys_in[0..numel) in rank order.  Sample sort: regular samples
 * -> all-gather -> world-1 splitters (key, global index), chosen on the device by this
 * library's rank kernel -> clo_partition_count -> all-gather of the bucket sizes ->
 * clo_partition_scatter writes every bucket into the receive buffer of its destination rank
 * -> barrier -> local LSD radix sort into keys_out (and payload_out).  This rank ends up with
 * *numel_out elements: the rank-th slice of the sorted sequence.  gidx0 = global index of this
 * rank's first element, or CLO_DIST_GIDX_AUTO (derived from the gathered counts).  Blocks the
 * host once (the received size).  Fails, writing nothing, when a slice exceeds the capacity
 * given to clo_dist_sort_setup or out_capacity. */
cl_bool clo_dist_sort_with_device_data(CloDist* d, CCLQueue* cq_exec, CCLBuffer* keys_in, CCLBuffer* payload_in,
	size_t numel, cl_ulong gidx0, CCLBuffer* keys_out, CCLBuffer* payload_out, size_t out_capacity,
	size_t* numel_out, GError** err);

/* Collective.  Exclusive scan of the concatenation of every rank's data_in (rank order):
 * per-GPU total -> all-gather of `world` totals -> local scan with the carry-in kept in device
 * memory.  Integer results are bit-exact (wrap-around kept); float carries are summed in f64. */
CCLEvent* clo_dist_scan_with_device_data(CloDist* d, CloScan* scanner, CCLQueue* cq_exec, CCLBuffer* data_in,
	CCLBuffer* data_out, size_t numel, GError** err);

/* RNG streams partition without communication: rank r owns the contiguous work-items
 * [*first, *first + *count) of total_streams; pass them to clo_rng_new_dev_gid_offset. */
void clo_dist_rng_partition(cl_ulong total_streams, cl_uint rank, cl_uint world, cl_ulong* first, cl_ulong* count);

/* Device time of the phases of the last clo_dist_sort call (CUDA events on the queue's stream):
 * samples + all-gather, splitters, count, sizes all-gather + slots, scatter, barrier, local sort.
 * Timing is off until clo_dist_set_timing(d, CL_TRUE).  Returns the number of values written. */
void clo_dist_set_timing(CloDist* d, cl_bool on);
cl_uint clo_dist_get_phases(CloDist* d, float* out_ms, cl_uint cap);
/* what the last sort sent to / received from every rank (world values each) */
cl_bool clo_dist_get_counts(CloDist* d, cl_ulong* sent, cl_ulong* received);

/* Library / device info. */
const char* clo_b200_version(void);
/* number of this library's kernels launched since load (bench evidence) */
cl_ulong clo_b200_launch_count(void);
void clo_b200_error_free(GError* err);

#ifdef __cplusplus
}
#endif
#endif
