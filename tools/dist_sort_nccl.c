/*
 * tools/dist_sort_nccl.c -- a plain C client of the multi-GPU entry points (clo_dist_*,
 * include/cl_ops/clo_b200.h): one process per GPU, NCCL for the three bookkeeping collectives,
 * no Python anywhere.  This is the code INTEGRATION.md shows; it is also a test
 * (tests/test_gpu_sort.py::test_c_client_dist_sort runs it on 2 GPUs).
 *
 *   build:  gcc -O2 -I include -I include/compat -I /usr/local/cuda/include tools/dist_sort_nccl.c \
 *             -o tools/dist_sort_nccl -L cl_ops_b200 -lcl_ops -lnccl -L /usr/local/cuda/lib64 -lcudart \
 *             -Wl,-rpath,'$ORIGIN/../cl_ops_b200'
 *   run:    tools/dist_sort_nccl <gpus> <log2 keys per gpu>
 *
 * The parent forks one child per GPU; rank 0 passes the ncclUniqueId through a file.  Every child sorts
 * its shard of xorshift keys with clo_dist_sort_with_device_data, checks that its slice is
 * sorted, that slices are ordered across ranks and that no key was lost (count and sum), and
 * then scans the same words with clo_dist_scan_with_device_data against a host prefix sum of
 * the rank totals.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>

#include <cuda_runtime.h>
#include <nccl.h>

#include <cl_ops.h>
#include <cl_ops/clo_b200.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
#define NK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { fprintf(stderr, "NCCL %s at %d\n", ncclGetErrorString(r_), __LINE__); exit(2); } } while (0)

typedef struct { ncclComm_t comm; int world; unsigned char* dev_tmp; int* dev_flag; } Comm;

/* ---- the three callbacks of CloDistComm ---- */
static int cb_all_gather_dev(void* user, const void* send, void* recv, size_t bytes, void* stream) {
	Comm* c = (Comm*) user;
	return ncclAllGather(send, recv, bytes, ncclChar, c->comm, (cudaStream_t) stream) == ncclSuccess ? 0 : 1;
}
static int cb_barrier_dev(void* user, void* stream) {
	Comm* c = (Comm*) user;
	return ncclAllReduce(c->dev_flag, c->dev_flag, 1, ncclInt, ncclSum, c->comm, (cudaStream_t) stream) == ncclSuccess ? 0 : 1;
}
static int cb_all_gather_host(void* user, const void* send, void* recv, size_t bytes) {
	Comm* c = (Comm*) user;             /* setup only: through a small device buffer */
	if (bytes > 256) return 1;
	if (cudaMemcpy(c->dev_tmp, send, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
	if (ncclAllGather(c->dev_tmp, c->dev_tmp + 256, bytes, ncclChar, c->comm, 0) != ncclSuccess) return 1;
	if (cudaStreamSynchronize(0) != cudaSuccess) return 1;
	return cudaMemcpy(recv, c->dev_tmp + 256, bytes * c->world, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}

static void die(GError* err, const char* what) {
	fprintf(stderr, "%s: %s\n", what, err ? err->message : "failed");
	exit(3);
}

static int child(int rank, int world, int log2n, const char* idfile) {
	CK(cudaSetDevice(rank));
	ncclUniqueId id;
	if (rank == 0) {
		/* rank 0 makes the id (nothing CUDA- or NCCL-related runs in the parent before the fork) */
		char tmp[96];
		snprintf(tmp, sizeof(tmp), "%s.tmp", idfile);
		NK(ncclGetUniqueId(&id));
		FILE* f = fopen(tmp, "wb");
		if (!f || fwrite(&id, sizeof(id), 1, f) != 1) return 2;
		fclose(f);
		rename(tmp, idfile);
	} else {
		FILE* f = NULL;
		for (int tries = 0; tries < 600 && !(f = fopen(idfile, "rb")); ++tries) usleep(100000);
		if (!f || fread(&id, sizeof(id), 1, f) != 1) { fprintf(stderr, "no id file\n"); return 2; }
		fclose(f);
	}
	Comm c; c.world = world;
	NK(ncclCommInitRank(&c.comm, world, id, rank));
	CK(cudaMalloc((void**) &c.dev_tmp, 256 + 256 * 16));
	CK(cudaMalloc((void**) &c.dev_flag, sizeof(int)));
	CK(cudaMemset(c.dev_flag, 0, sizeof(int)));

	GError* err = NULL;
	int dev_index = rank;
	CCLContext* ctx = ccl_context_new_from_menu_full(&dev_index, &err);
	if (!ctx) die(err, "context");
	CCLQueue* cq = ccl_queue_new(ctx, NULL, 0, &err);
	if (!cq) die(err, "queue");

	CloDistComm comm = { &c, (cl_uint) rank, (cl_uint) world, cb_all_gather_dev, cb_barrier_dev, cb_all_gather_host };
	CloDist* d = clo_dist_new(ctx, &comm, &err);
	if (!d) die(err, "clo_dist_new");

	const size_t n = (size_t) 1 << log2n, cap = n + n / 4;
	if (!clo_dist_sort_setup(d, CLO_UINT, cap, CL_FALSE, &err)) die(err, "clo_dist_sort_setup");

	/* this rank's shard: xorshift32 of the global index (so the union is rank-count independent) */
	cl_uint* h = (cl_uint*) malloc(cap * sizeof(cl_uint));
	unsigned long long my_sum = 0;
	for (size_t i = 0; i < n; ++i) {
		cl_uint x = (cl_uint) ((size_t) rank * n + i) * 2654435761u + 1u;
		x ^= x << 13; x ^= x >> 17; x ^= x << 5;
		h[i] = x; my_sum += x;
	}
	CCLBuffer* in = ccl_buffer_new(ctx, CL_MEM_READ_WRITE, n * sizeof(cl_uint), NULL, &err);
	CCLBuffer* out = ccl_buffer_new(ctx, CL_MEM_READ_WRITE, cap * sizeof(cl_uint), NULL, &err);
	if (!in || !out) die(err, "buffers");
	if (!ccl_buffer_enqueue_write(in, cq, CL_TRUE, 0, n * sizeof(cl_uint), h, NULL, &err)) die(err, "write");

	size_t n_out = 0;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float best = 1e30f;
	for (int it = 0; it < 4; ++it) {
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(e0, 0));
		if (!clo_dist_sort_with_device_data(d, cq, in, NULL, n, (cl_ulong) rank * n, out, NULL, cap, &n_out, &err)) die(err, "clo_dist_sort");
		if (!ccl_queue_finish(cq, &err)) die(err, "finish");
		CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (it > 0 && ms < best) best = ms;
	}
	if (!ccl_buffer_enqueue_read(out, cq, CL_TRUE, 0, n_out * sizeof(cl_uint), h, NULL, &err)) die(err, "read");

	int ok = 1;
	unsigned long long out_sum = 0;
	for (size_t i = 0; i < n_out; ++i) { out_sum += h[i]; if (i && h[i - 1] > h[i]) ok = 0; }
	/* across ranks: [count, sum in, sum out, first, last] */
	unsigned long long mine[5] = { n_out, my_sum, out_sum, n_out ? h[0] : 0xffffffffull, n_out ? h[n_out - 1] : 0 }, all[5 * 16];
	unsigned long long* dv;
	CK(cudaMalloc((void**) &dv, sizeof(mine) * (world + 1)));
	CK(cudaMemcpy(dv, mine, sizeof(mine), cudaMemcpyHostToDevice));
	NK(ncclAllGather(dv, dv + 5, 5, ncclUint64, c.comm, 0));
	CK(cudaStreamSynchronize(0));
	CK(cudaMemcpy(all, dv + 5, sizeof(mine) * world, cudaMemcpyDeviceToHost));
	unsigned long long cnt = 0, si = 0, so = 0;
	for (int r = 0; r < world; ++r) {
		cnt += all[5 * r]; si += all[5 * r + 1]; so += all[5 * r + 2];
		if (r && all[5 * r] && all[5 * (r - 1)] && all[5 * (r - 1) + 4] > all[5 * r + 3]) ok = 0;
	}
	if (cnt != (unsigned long long) n * world || si != so) ok = 0;

	/* ---- scan of the input words (uint -> ulong) over all GPUs */
	CloScan* sc = clo_scan_new("blelloch", NULL, ctx, CLO_UINT, CLO_ULONG, NULL, &err);
	if (!sc) die(err, "clo_scan_new");
	CCLBuffer* sout = ccl_buffer_new(ctx, CL_MEM_READ_WRITE, n * sizeof(cl_ulong), NULL, &err);
	if (!sout) die(err, "scan buffer");
	if (!clo_dist_scan_with_device_data(d, sc, cq, in, sout, n, &err)) die(err, "clo_dist_scan");
	cl_ulong first = 0, last = 0;
	if (!ccl_buffer_enqueue_read(sout, cq, CL_TRUE, 0, sizeof(cl_ulong), &first, NULL, &err)) die(err, "read");
	if (!ccl_buffer_enqueue_read(sout, cq, CL_TRUE, (n - 1) * sizeof(cl_ulong), sizeof(cl_ulong), &last, NULL, &err)) die(err, "read");
	unsigned long long before = 0;
	for (int r = 0; r < rank; ++r) before += all[5 * r + 1];
	{
		cl_uint x = (cl_uint) ((size_t) rank * n + n - 1) * 2654435761u + 1u;
		x ^= x << 13; x ^= x >> 17; x ^= x << 5;
		if (first != before || last != before + my_sum - x) ok = 0;
	}

	if (rank == 0)
		printf("{\"c_client\": \"clo_dist_sort_with_device_data\", \"gpus\": %d, \"keys_per_gpu\": %zu, \"ms\": %.3f, \"gkeys_s\": %.2f, \"ok\": %s}\n",
			world, n, best, (double) n * world / best / 1e6, ok ? "true" : "false");
	else if (!ok) fprintf(stderr, "rank %d: check failed\n", rank);
	fflush(stdout);                     /* the child leaves through _exit */

	clo_scan_destroy(sc);
	ccl_buffer_destroy(sout); ccl_buffer_destroy(in); ccl_buffer_destroy(out);
	clo_dist_destroy(d);
	ccl_queue_destroy(cq); ccl_context_destroy(ctx);
	ncclCommDestroy(c.comm);
	free(h);
	return ok ? 0 : 1;
}

int main(int argc, char** argv) {
	const int world = argc > 1 ? atoi(argv[1]) : 2;
	const int log2n = argc > 2 ? atoi(argv[2]) : 24;
	if (world < 1 || world > 16 || log2n < 1 || log2n > 29) { fprintf(stderr, "usage: %s <gpus 1..16> <log2 keys per gpu>\n", argv[0]); return 2; }
	char idfile[64];
	snprintf(idfile, sizeof(idfile), "/tmp/clo_nccl_id_%d", (int) getpid());
	pid_t pids[16];
	for (int r = 0; r < world; ++r) {
		pids[r] = fork();
		if (pids[r] == 0) _exit(child(r, world, log2n, idfile));
	}
	int rc = 0;
	for (int r = 0; r < world; ++r) { int st = 0; waitpid(pids[r], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st)) rc = 1; }
	unlink(idfile);
	return rc;
}
