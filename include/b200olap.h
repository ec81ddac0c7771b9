/*
 * b200olap.h — C ABI of the B200 columnar-operator library (libb200olap.so).
 *
 * This header is the drop-in boundary that replaces dpu_olap's device layer:
 *   - host/dpuext/dpuext.hpp      (dpu::DpuSet::allocate :710, load :739, exec :638,
 *                                  copy :163/:277/:442-533, async().call/sync :860-898)
 *   - host/dpuext/arrow_utils.cc  (arrow_copy_to_dpus :47-73, arrow_copy_from_dpus* :147-266)
 *   - shared/umq/kernels.h        (enum Kernel :12-20, param structs :27-51)
 *   - dpu/{filter,aggr,take,partition,join}/main.c + dpu/shared/kernels/ (the DPU programs)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only. No C++/torch/Arrow types cross this boundary.
 *   - Every entry point returns an int status (B2_OK == 0). No exception crosses the ABI.
 *     (Reference convention: dpu_error_t + DPU_RETURN_NOT_OK, host/dpuext/status.h:7-12.)
 *   - "_dev" entry points take DEVICE pointers and a cudaStream_t (as void*); they only enqueue
 *     work on that stream and never synchronise (this is the reference's "dpu-work" leg).
 *   - "_host" entry points take HOST pointers, one per Arrow record batch (data buffer #1 of a
 *     non-nullable uint32 column, host/dpuext/arrow_utils.cc:23,60-66); they upload, run and
 *     download inside the call and are synchronous at return (the reference's
 *     copy-to-dpu / dpu-work / copy-from-dpu legs, reported in b2_timings).
 *   - A b2_ctx is bound to ONE GPU and is NOT thread-safe (one ctx per host thread / per rank).
 *   - Memory the library allocates is owned by the ctx / result handle and released by the
 *     matching b2_*_free / b2_ctx_destroy; callers never free() library memory.
 *   - Columns are non-null unless an entry point says "nullable" (the reference assumes non-null:
 *     it passes a nullptr bitmap, host/filter/filter_dpu.cc:91). Nullable entry points take Arrow
 *     validity bitmaps: bit i = row i of the packed column, least significant bit first, pointer
 *     4-byte aligned, buffer padded to a multiple of 4 bytes; NULL = no nulls.
 */
#ifndef B200OLAP_H_
#define B200OLAP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_VERSION 100 /* 0.1.0 */

/* ---- status codes (stable) ------------------------------------------------------------- */
enum b2_status {
  B2_OK = 0,
  B2_ERR_INVALID = 1,      /* bad argument (null pointer, negative size, misaligned buffer) */
  B2_ERR_CUDA = 2,         /* a CUDA runtime call failed; see b2_last_error */
  B2_ERR_OOM = 3,          /* device or pinned-host allocation failed */
  B2_ERR_UNSUPPORTED = 4,  /* valid request this build does not implement */
  B2_ERR_WORKSPACE = 5,    /* caller-provided workspace too small */
  B2_ERR_OVERFLOW = 6      /* result does not fit the caller-provided output capacity */
};

/* 32-bit column types of the typed entry points (the reference fixes `#define T uint32_t`,
 * dpu/shared/common.h:3; uint32 is what every other entry point means). */
enum b2_dtype32 { B2_U32 = 0, B2_I32 = 1, B2_F32 = 2 };
/* 64-bit column types (b2_aggr_64_*); the values continue b2_dtype32's. */
enum b2_dtype64 { B2_U64 = 3, B2_I64 = 4, B2_F64 = 5 /* filter only */ };

typedef struct b2_ctx b2_ctx;

/* Phase timings of the last *_host call, in milliseconds. Names follow the reference's timers
 * (host/filter/filter_dpu.cc:107-110): copy-to-dpu / dpu-work / copy-from-dpu / build-result. */
typedef struct b2_timings {
  double copy_to_dev_ms;   /* host->device, wall time the copies occupied their stream */
  double dev_work_ms;      /* kernels (CUDA events on the compute stream) */
  double copy_from_dev_ms; /* device->host */
  double total_ms;         /* wall clock of the whole call */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int32_t kernel_launches; /* kernels of this library launched by the call */
  int32_t reserved;
} b2_timings;

/* ---- context (replaces dpu::DpuSet::allocate / load / dtor, dpuext.hpp:669-739) ---------- */
int b2_version(void);
const char* b2_strerror(int status);
int b2_device_count(int* count);
int b2_ctx_create(int device, b2_ctx** out);
int b2_ctx_destroy(b2_ctx* ctx);
/* Human-readable detail for the last non-OK status returned on this ctx (never NULL). */
const char* b2_last_error(const b2_ctx* ctx);
/* Kernels launched by this library on this ctx since creation (all entry points). */
int64_t b2_launch_count(const b2_ctx* ctx);
int b2_ctx_device(const b2_ctx* ctx);
int b2_ctx_sm_count(const b2_ctx* ctx);
/* Kernel-selection knobs of ONE ctx (no process-global state). Every setting computes the same,
 * bit-identical result; they exist so that tests can push every input through every kernel variant
 * and so that a deployment can re-tune thresholds without a rebuild. Defaults are the measured best
 * on B200 (profiles/r1_filter.md, profiles/r1_scatter_fanout.md). */
enum b2_tunable {
  B2_TUNE_SCATTER_SECTORS_MIN_BITS = 0, /* log2 fan-out from which the radix scatter stores whole 32 B sectors (8; 0 = always, 11 = never) */
  B2_TUNE_SCATTER_PREFETCH = 1,         /* scatter kernels request the next tile while flushing this one (1) */
  B2_TUNE_SCATTER_SHAPE = 2,            /* plain scatter kernel shape: 0 = 512 thr x 16 rows x 2 CTA/SM, 1, 2, 3, 8 */
  B2_TUNE_FILTER_VARIANT = 3,           /* filter kernel shape 0..7 (6) */
  B2_TUNE_SCATTER_SECTOR_TILE = 4,      /* whole-sector scatter: 0 = 8192-row tiles x 2 CTA/SM, 1 = 16384-row tiles x 1 CTA/SM,
                                           2 = quad-aligned regions, 14336-row tiles x 1 CTA/SM,
                                           3 = quad-aligned regions flushed by the copy engine (cp.async.bulk), 16384-row tiles */
  B2_TUNE_PEER_SCATTER_CTAS = 6,        /* CTA budget of the peer (NVLink) scatter kernel: 0 = one CTA per work unit (all SMs);
                                           n > 0 = at most n CTAs walk the units and leave the other SMs to kernels of
                                           other streams (the overlapped sharded join sets it around the probe side's scatter) */
  B2_TUNE_PEER_SCATTER_KERNEL = 7,      /* peer (NVLink) scatter: 0 = whole 128-byte lines stored by the threads (line carry),
                                           1 = whole 32-byte sectors, one copy-engine bulk copy per (bucket, tile) */
  B2_TUNE_FILTER64_KERNEL = 8,          /* 64-bit filter (filter64.cu): 0 = single pass with decoupled look-back (8 + 8 s bytes per row),
                                           1 = counted two-pass compaction (8 + 8 + 8 s) */
  B2_TUNE_JOIN_DIRECT_MIN_ROWS = 5      /* perfect-hash probe path (join.cu): used when <= 14 hash bits are left below the partition
                                           bits; the planner adds partition bits to get there while partitions keep at least this
                                           many build rows (2048), and takes fewer when 2^14-row partitions are enough. 0 = path off, 1 = always when the bits allow (tests). */
};
int b2_ctx_set_tunable(b2_ctx* ctx, int which, int value);
int b2_ctx_get_tunable(const b2_ctx* ctx, int which, int* value);

/* ---- device set: the GPUs of one node behind ONE handle ----------------------------------------
 * Replaces dpu::DpuSet::allocate(nr_dpus) (host/dpuext/dpuext.hpp:704-739): the reference's set owns
 * every DPU and its operators iterate over them (filter_dpu.cc:127 "batch i -> DPU i"; join_dpu.cc:254
 * groups of nr_dpus partitions, repartitioned through the HOST, partitioner.cc:350-375). A b2_set owns
 * one b2_ctx per GPU of ONE process and enables peer access between them:
 *   - filter / sum / take shard by contiguous batch ranges (one host thread per GPU drives that GPU's
 *     own pipelined host entry point; no data-path collective);
 *   - the join routes both sides by the top log2(N) hash bits with the fused peer-store shuffle
 *     (count -> b2_shuffle_p2p_plan_dev over peer pointers -> NVLink scatter -> local join), ordered
 *     by CUDA events across devices: no NCCL, no host synchronisation between count and local join.
 * devices == NULL means 0..n-1. The join needs n to be a power of two and peer access between all
 * members (b2_set_peer_access); the range-sharded operators take any n. Same threading rule as a
 * ctx: one host thread per set. b2_set_ctx hands out the member contexts for the *_dev / single-GPU
 * entry points. Timings: phases of concurrently running members overlap, so the *_ms fields are the
 * maximum over members, bytes and launches the sum. */
typedef struct b2_set b2_set;
int b2_set_create(const int* devices, int n, b2_set** out);
int b2_set_destroy(b2_set* set);
int b2_set_size(const b2_set* set);
b2_ctx* b2_set_ctx(b2_set* set, int i);
int b2_set_peer_access(const b2_set* set);
const char* b2_set_last_error(const b2_set* set);
int64_t b2_set_launch_count(const b2_set* set);
int b2_set_set_inputs_pinned(b2_set* set, int on);
/* SumDpu::Run over the whole set (aggr_dpu.cc:31-89; one partial per device added on the host, :82-84). */
int b2_set_sum_u32_host(b2_set* set, const uint32_t* const* batch_ptrs, const int64_t* batch_lens, int64_t nbatches,
                        uint64_t* sum, b2_timings* timings);
/* FilterDpu over the whole set; arguments as b2_filter_lt_u32_host / b2_filter_fetch_host. */
int b2_set_filter_lt_u32_host(b2_set* set, const uint32_t* const* batch_ptrs, const int64_t* batch_lens,
                              int64_t nbatches, uint32_t threshold, int64_t* out_counts, uint64_t* total,
                              b2_timings* timings);
int b2_set_filter_fetch_host(b2_set* set, uint32_t* const* out_ptrs, int64_t nbatches, b2_timings* timings);
/* TakeDpu over the whole set; arguments as b2_take_u32_host. */
int b2_set_take_u32_host(b2_set* set, const uint32_t* const* value_ptrs, const int64_t* value_lens,
                         const uint32_t* const* idx_ptrs, const int64_t* idx_lens, int64_t nbatches,
                         uint32_t* const* out_ptrs, b2_timings* timings);
/* JoinDpu::Run over the whole set; arguments as b2_join_u32_host / b2_join_fetch_host. The result is
 * fetched GPU by GPU into out_*[0 .. rows) (row order unspecified, as JoinDpu's). Skewed keys that
 * overflow a receive buffer make the set grow its buffers and repeat the exchange; duplicate build keys
 * that overflow a GPU's output columns make that GPU repeat its local join with the exact capacity. */
int b2_set_join_u32_host(b2_set* set, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                         const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches,
                         uint64_t* out_rows, b2_timings* timings);
int b2_set_join_fetch_host(b2_set* set, uint32_t* out_fk, uint32_t* out_y, uint32_t* out_x, int64_t capacity_rows,
                           b2_timings* timings);
/* The fused pipeline [filter left payload < y_threshold ->] join -> COUNT / SUM / SUM over the whole set
 * (arguments as b2_join_aggr_u32_host): the same exchange as b2_set_join_u32_host, every GPU's local
 * join adds its output rows' payloads instead of materialising them, the partial results are added on
 * the host (as SumDpu adds its per-device partials, aggr_dpu.cc:82-84). */
struct b2_join_aggr;
int b2_set_join_aggr_u32_host(b2_set* set, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                              const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches, int filter_y,
                              uint32_t y_threshold, struct b2_join_aggr* out, b2_timings* timings);

/* ---- pinned host memory (zero-copy Arrow interop at the boundary) -------------------------- */
/* Arrow buffers live in pageable memory; copies from/to them run at a fraction of the PCIe rate.
 * b2_host_register page-locks an EXISTING buffer in place (e.g. the data buffers of the input
 * record batches, once, outside the timed region — the reference's fixtures build their inputs in
 * SetUp too); b2_host_alloc_pinned returns page-locked memory a caller can wrap in an
 * arrow::Buffer for results. Neither needs a ctx. The reference itself lists zero-copy transfers
 * as future work (host/dpuext/arrow_utils.h:28-29). */
/* Promise that every host INPUT pointer later passed to this ctx's *_host calls is page-locked and
 * device-accessible (b2_host_register, b2_host_alloc_pinned, cudaHostAlloc/cudaHostRegister with
 * the mapped flag). A group of batches is then uploaded by ONE kernel that reads host memory over
 * PCIe directly, instead of one DMA per record batch (at 256 KB per batch a DMA's set-up costs as
 * much as its transfer). Breaking the promise faults the kernel; off by default. */
int b2_ctx_set_inputs_pinned(b2_ctx* ctx, int on);
int b2_host_alloc_pinned(size_t bytes, void** out);
int b2_host_free_pinned(void* p);
int b2_host_register(const void* p, size_t bytes);
int b2_host_unregister(const void* p);

/* ---- device-resident columns and the Arrow C Device Data Interface ---------------------------------
 * The reference copies every operator's input to the DPUs and every result back
 * (arrow_copy_to_dpus / arrow_copy_from_dpus*, host/dpuext/arrow_utils.cc:47-73,147-266) and lists
 * "results without a copy" as future work itself (arrow_utils.h:28-29). A b2_col is one packed,
 * non-null uint32 column in HBM plus its batch boundaries (host-side metadata). Operators over b2_cols
 * leave their results in HBM, so chains such as filter -> take -> sum or join -> sum cross PCIe once
 * on the way in and not at all in between; b2_col_export / b2_col_import hand a column to, or take one
 * from, any Arrow consumer on the same GPU as an ArrowDeviceArray WITHOUT a copy:
 *   device_type = ARROW_DEVICE_CUDA, device_id = the ctx's device, array.buffers = {NULL, device pointer},
 *   sync_event = cudaEvent_t* recorded after the kernels that produced the column.
 * The struct definitions below are the ones of the Arrow specification (arrow/c/abi.h), under the
 * specification's own include guards, so this header and Arrow's can be included together.
 * Ownership: b2_col_free drops the handle; an exported array keeps the device memory alive until its
 * release callback has run as well. b2_col_import MOVES the array (its release is set to NULL) and the
 * column calls the producer's release when it dies. All *_col operators run on the ctx's own compute
 * stream and are synchronous at return only where they return host data (sum, the filter's chunk
 * boundaries, the join's row count). */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif
#ifndef ARROW_C_DEVICE_DATA_INTERFACE
#define ARROW_C_DEVICE_DATA_INTERFACE
typedef int32_t ArrowDeviceType;
#define ARROW_DEVICE_CPU 1
#define ARROW_DEVICE_CUDA 2
#define ARROW_DEVICE_CUDA_HOST 3
struct ArrowDeviceArray {
  struct ArrowArray array;
  int64_t device_id;
  ArrowDeviceType device_type;
  void* sync_event;
  int64_t reserved[3];
};
#endif
typedef struct b2_col b2_col;
/* Host batches -> one packed device column (the only H2D copy of a chain). */
int b2_col_upload_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs, const int64_t* batch_lens, int64_t nbatches,
                       b2_col** out);
int b2_col_free(b2_col* col);
int64_t b2_col_rows(const b2_col* col);
int64_t b2_col_nbatches(const b2_col* col);
const uint32_t* b2_col_device_ptr(const b2_col* col);
/* out[0 .. nbatches]: batch boundaries in rows (capacity >= nbatches + 1). */
int b2_col_batch_offsets(const b2_col* col, int64_t* out, int64_t capacity);
/* Batch b of the column -> out_ptrs[b] (the only D2H copy of a chain). Synchronous. */
int b2_col_download_host(b2_ctx* ctx, const b2_col* col, uint32_t* const* out_ptrs, int64_t nbatches);
int b2_col_export(b2_col* col, struct ArrowDeviceArray* out);
/* batch_lens / nbatches describe how the array's rows split into record batches (NULL / 0: one batch). */
int b2_col_import(b2_ctx* ctx, struct ArrowDeviceArray* in, const int64_t* batch_lens, int64_t nbatches, b2_col** out);
/* The operators over device columns (semantics of the *_host entry points above). The filter's result
 * has one chunk per input batch; take is batch-local; the join's three result columns have one chunk. */
int b2_sum_u32_col(b2_ctx* ctx, const b2_col* col, uint64_t* sum);
int b2_filter_lt_u32_col(b2_ctx* ctx, const b2_col* col, uint32_t threshold, b2_col** out);
int b2_take_u32_col(b2_ctx* ctx, const b2_col* values, const b2_col* indices, b2_col** out);
int b2_join_u32_col(b2_ctx* ctx, const b2_col* fk, const b2_col* y, const b2_col* pk, const b2_col* x, b2_col** out_fk,
                    b2_col** out_y, b2_col** out_x);

/* ---- synthetic inputs (replaces host/generator for device-resident benchmarks) ----------- */
/* One array per batch, bit-identical to arrow::random::RandomArrayGenerator's
 * GenerateTypedDataNoNan for uint32 (host/generator/random.cc:103-109):
 *   pcg32_fast rng(data_seed[b]);  std::uniform_int_distribution<uint32_t> dist(lo[b], hi[b]);
 * data_seeds are the seeds ACTUALLY fed to pcg32_fast (i.e. seed()+1, random.cc:111-125,190-196).
 * lo/hi may be NULL (full range). hi-lo+1 must be 2^32 or a power of two (the only shapes the
 * reference's fixtures use; other ranges reject draws and cannot be generated in parallel)
 * else B2_ERR_UNSUPPORTED. data_seeds/lo/hi are HOST arrays of length nbatches. */
int b2_gen_u32_dev(b2_ctx* ctx, const uint64_t* data_seeds, const uint32_t* lo, const uint32_t* hi,
                   int64_t nbatches, int64_t batch_len, uint32_t* d_out, void* stream);
/* out[i] = (uint32_t)(start + i)   — generator::MakeIndexColumn, host/generator/generator.cc:59-71 */
int b2_iota_u32_dev(b2_ctx* ctx, uint64_t start, int64_t n, uint32_t* d_out, void* stream);

/* ---- Sum (replaces dpu/aggr/main.c:44-89 + dpu/shared/kernels/aggr.c:16-33) -------------- */
/* *d_sum = sum of d_in[0..n) as uint64 (mod 2^64). One kernel launch. Every launch has its own
 * scratch slot (a ring of 16 per ctx), so sums of one ctx may be in flight on different streams. */
int b2_sum_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t n, uint64_t* d_sum, void* stream);
/* Fused pipeline filter(v < threshold) -> sum: *d_sum = sum of the rows below the threshold,
 * *d_count (may be NULL) = how many there are. One read of the column, nothing materialised — the
 * plan the reference keeps commented out in host/aggr/aggr_native.cc:59-65. */
int b2_sum_lt_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t n, uint32_t threshold, uint64_t* d_sum,
                      uint64_t* d_count, void* stream);
/* Aggregates over a NULLABLE column in one pass, with the semantics of Arrow's compute kernels
 * (the reference's oracle, host/aggr/aggr_native.cc:68-73): null rows are skipped. The reference's
 * `enum AggregatorType` only has AggrSum (shared/umq/kernels.h:22-25); min / max / count come for
 * free with the same 4 B/row read. count == 0 means every aggregate is null (min = 0xffffffff,
 * max = 0 then). Two kernel launches (init + reduce). */
typedef struct b2_aggr_u32 {
  uint64_t sum;   /* sum of the valid rows, mod 2^64 */
  uint64_t count; /* number of valid rows */
  uint32_t min;
  uint32_t max;
} b2_aggr_u32;
int b2_aggr_u32_dev(b2_ctx* ctx, const uint32_t* d_in, const uint8_t* d_valid, int64_t n,
                    b2_aggr_u32* d_out, void* stream);
/* The same for an int32 column (dtype B2_I32: sum is an int64, min / max are in signed order, all
 * returned as bit patterns in the same struct; count == 0: min = INT32_MAX, max = INT32_MIN).
 * B2_F32 is rejected (B2_ERR_UNSUPPORTED): a float sum's rounding and the sign of a zero min / max
 * depend on the evaluation order, in Arrow itself too, so there is nothing bit-exact to match. */
int b2_aggr_32_dev(b2_ctx* ctx, const void* d_in, int dtype, const uint8_t* d_valid, int64_t n,
                   b2_aggr_u32* d_out, void* stream);
/* Host batches in, aggregates out. valid_ptrs[b] = validity bitmap of batch b starting at bit
 * valid_bit_offsets[b] (Arrow buffer #0 and the array's offset); valid_ptrs, valid_ptrs[b] and
 * valid_bit_offsets may each be NULL (no nulls / offset 0). */
int b2_aggr_u32_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                     const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches,
                     b2_aggr_u32* out, b2_timings* timings);
int b2_aggr_32_host(b2_ctx* ctx, const void* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                    const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches, int dtype,
                    b2_aggr_u32* out, b2_timings* timings);
/* The aggregates over a 64-bit column (dtype B2_U64 / B2_I64; SURVEY.md section 8f-3 "other
 * fixed-width types" — the reference fixes T = uint32_t, dpu/shared/common.h:3). Arrow semantics
 * (aggr_native.cc:68-73 with a uint64 / int64 column): the sum has the column's type and wraps
 * mod 2^64, min / max are in the type's order, null rows are skipped; count == 0: min / max hold the
 * type's largest / smallest value. int64 results travel as bit patterns. d_in must be 8-byte
 * aligned. Argument meaning as b2_aggr_32_dev / b2_aggr_32_host. */
typedef struct b2_aggr_u64 {
  uint64_t sum;
  uint64_t count;
  uint64_t min;
  uint64_t max;
} b2_aggr_u64;
int b2_aggr_64_dev(b2_ctx* ctx, const void* d_in, int dtype, const uint8_t* d_valid, int64_t n,
                   b2_aggr_u64* d_out, void* stream);
int b2_aggr_64_host(b2_ctx* ctx, const void* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                    const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches, int dtype,
                    b2_aggr_u64* out, b2_timings* timings);
/* SumDpu::Run (host/aggr/aggr_dpu.cc:31-89): batches on the host, result on the host. */
int b2_sum_u32_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs, const int64_t* batch_lens,
                    int64_t nbatches, uint64_t* sum, b2_timings* timings);

/* ---- Filter (replaces dpu/shared/kernels/filter.c:57-177, predicate :25) ----------------- */
/* Order-preserving selection of d_in[i] < threshold over nbatches equal-length batches packed
 * back to back in d_in. Output rows are written compacted, in input order, to d_out (capacity
 * must be >= nbatches*batch_len rows unless the caller knows a tighter bound);
 * d_batch_end[b] (int64, nbatches entries) = number of selected rows in batches 0..b, so chunk b
 * of the result (FilterDpu::GetResult returns one chunk per batch, filter_dpu.cc:89-96,162-166)
 * is d_out[d_batch_end[b-1] .. d_batch_end[b]). d_total (may be NULL) receives the grand total.
 * d_carry_in (may be NULL = 0): device int64 holding the output row at which this call starts;
 * d_out indices, d_batch_end and d_total are all offset by it, so a column can be filtered in
 * several calls (streamed chunks) that append to one compacted result without a host round trip
 * (pass call k's d_total as call k+1's d_carry_in).
 * d_ws: workspace of b2_filter_ws_bytes() bytes.
 * Concurrency: the filter kernels (32- and 64-bit) are persistent grids sized to what the GPU holds at
 * once, and a CTA may wait for tiles counted by any other CTA of its grid. Two filter launches that
 * overlap on one GPU (different streams, no ordering between them) can each hold the SM slots the
 * other's unstarted CTAs need and never finish: order filter launches on one GPU (one stream, or
 * events between streams). Kernels that end on their own may share the GPU with a filter; they delay
 * it, they cannot block it. */
size_t b2_filter_ws_bytes(int64_t nbatches, int64_t batch_len);
int b2_filter_lt_u32_dev(b2_ctx* ctx, const uint32_t* d_in, int64_t nbatches, int64_t batch_len,
                         uint32_t threshold, uint32_t* d_out, int64_t* d_batch_end,
                         int64_t* d_total, const int64_t* d_carry_in, void* d_ws, size_t ws_bytes,
                         void* stream);
/* Nullable column: a row is selected iff it is valid AND below the threshold — Arrow's filter drops
 * rows whose predicate is null (the semantics of the reference's oracle, filter_native.cc:52-66;
 * the DPU path itself has no bitmaps). d_valid: validity bitmap over the packed column (see
 * Conventions), NULL = b2_filter_lt_u32_dev. The result has no nulls. Everything else as above. */
int b2_filter_lt_u32_nullable_dev(b2_ctx* ctx, const uint32_t* d_in, const uint8_t* d_valid,
                                  int64_t nbatches, int64_t batch_len, uint32_t threshold,
                                  uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                                  const int64_t* d_carry_in, void* d_ws, size_t ws_bytes, void* stream);
/* Other 32-bit column types (the reference fixes `#define T uint32_t`, dpu/shared/common.h:3):
 * the same kernel comparing `v < threshold` as int32 or float32 (IEEE: a NaN row is never selected,
 * as in Arrow). Values travel as raw 32-bit words; threshold_bits is the threshold's bit pattern;
 * d_valid as above (NULL = no nulls). */
int b2_filter_lt_32_dev(b2_ctx* ctx, const void* d_in, int dtype, uint32_t threshold_bits,
                        const uint8_t* d_valid, int64_t nbatches, int64_t batch_len, void* d_out,
                        int64_t* d_batch_end, int64_t* d_total, const int64_t* d_carry_in, void* d_ws,
                        size_t ws_bytes, void* stream);
/* Ragged variant: batch b occupies d_in[d_batch_off[b] .. d_batch_off[b+1]) (int64 device array
 * of nbatches+1 entries; host copy h_batch_off is needed to size the launch). Empty batches ok. */
size_t b2_filter_ragged_ws_bytes(const int64_t* h_batch_off, int64_t nbatches);
int b2_filter_lt_32_ragged_dev(b2_ctx* ctx, const void* d_in, int dtype, uint32_t threshold_bits,
                               const uint8_t* d_valid, const int64_t* h_batch_off, const int64_t* d_batch_off,
                               int64_t nbatches, void* d_out, int64_t* d_batch_end, int64_t* d_total,
                               const int64_t* d_carry_in, void* d_ws, size_t ws_bytes, void* stream);
int b2_filter_lt_u32_ragged_dev(b2_ctx* ctx, const uint32_t* d_in, const int64_t* h_batch_off,
                                const int64_t* d_batch_off, int64_t nbatches, uint32_t threshold,
                                uint32_t* d_out, int64_t* d_batch_end, int64_t* d_total,
                                const int64_t* d_carry_in, void* d_ws, size_t ws_bytes,
                                void* stream);
/* FilterDpu::GetResult (host/filter/filter_dpu.cc:104-169) in two steps, because the caller can
 * only size its result buffers once the per-batch counts are known (the reference reads
 * "output_buffer_length" first, then allocates and pulls "output_buffer", :57-83):
 *   b2_filter_lt_u32_host  uploads, filters, keeps the result on the device, returns the
 *                          per-batch selected counts in out_counts[nbatches];
 *   b2_filter_fetch_host   downloads chunk b into out_ptrs[b] (capacity out_counts[b] rows).
 * The pending result is dropped by the next *_host call on the ctx or b2_ctx_destroy. */
int b2_filter_lt_u32_host(b2_ctx* ctx, const uint32_t* const* batch_ptrs,
                          const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                          int64_t* out_counts, uint64_t* total, b2_timings* timings);
int b2_filter_fetch_host(b2_ctx* ctx, uint32_t* const* out_ptrs, int64_t nbatches,
                         b2_timings* timings);
/* One-call streaming form for callers that can offer ONE result buffer up front (capacity
 * out_capacity rows; nbatches*batch_len always suffices): the compacted result of all batches
 * is written to out[0 .. *total), chunk b being out[sum(out_counts[0..b)) ..) — the buffers an
 * Arrow ChunkedArray would slice. Upload, kernels and download of successive 64 MiB groups of
 * batches overlap (the reference's per-rank copy/exec/copy-back callbacks, filter_dpu.cc:128-157),
 * and no device memory is allocated after the first call. Use pinned host memory for full PCIe
 * speed. B2_ERR_OVERFLOW if the result does not fit (counts and *total are still valid). */
int b2_filter_lt_u32_host_into(b2_ctx* ctx, const uint32_t* const* batch_ptrs,
                               const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                               uint32_t* out, int64_t out_capacity, int64_t* out_counts,
                               uint64_t* total, b2_timings* timings);
/* Nullable form of b2_filter_lt_u32_host_into (validity arguments as b2_aggr_u32_host). Batches may
 * have different lengths; not chunked. */
int b2_filter_lt_u32_nullable_host_into(b2_ctx* ctx, const uint32_t* const* batch_ptrs,
                                        const uint8_t* const* valid_ptrs, const int64_t* valid_bit_offsets,
                                        const int64_t* batch_lens, int64_t nbatches, uint32_t threshold,
                                        uint32_t* out, int64_t out_capacity, int64_t* out_counts,
                                        uint64_t* total, b2_timings* timings);
/* 64-bit columns (dtype B2_U64 / B2_I64 / B2_F64; SURVEY.md section 8f-3 "other fixed-width types" — the
 * reference fixes T = uint32_t, dpu/shared/common.h:3): `v < threshold` in the column's type, rows keep
 * their order, a null row is dropped, a NaN row is never selected (the Acero filter plan of the
 * reference's oracle, filter_native.cc:52-66, over a 64-bit column). A counted two-pass compaction
 * (csrc/filter64.cu), not the single-pass 32-bit kernel. d_in / d_out: n packed 64-bit values, 8-byte
 * aligned; threshold_bits: the threshold's bit pattern; d_valid: validity bitmap over the packed column
 * or NULL; batches: d_batch_off (device, nbatches + 1 row offsets) or NULL = nbatches x batch_len rows;
 * d_batch_end[b] = rows selected up to the end of batch b, *d_total = rows selected (both on the
 * device). d_ws: b2_filter_64_ws_bytes(n) bytes, 256 B aligned. */
size_t b2_filter_64_ws_bytes(int64_t n);
int b2_filter_lt_64_dev(b2_ctx* ctx, const void* d_in, int dtype, uint64_t threshold_bits, const uint8_t* d_valid,
                        int64_t n, const int64_t* d_batch_off, int64_t nbatches, int64_t batch_len, void* d_out,
                        int64_t* d_batch_end, int64_t* d_total, void* d_ws, size_t ws_bytes, void* stream);
/* ... over host batches of any lengths (validity arguments as b2_aggr_u32_host): upload, filter,
 * download; out receives the selected values back to back (out_capacity values), out_counts[b] the
 * rows of batch b. Not chunked. */
int b2_filter_lt_64_host_into(b2_ctx* ctx, const void* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                              const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches,
                              int dtype, uint64_t threshold_bits, void* out, int64_t out_capacity,
                              int64_t* out_counts, uint64_t* total, b2_timings* timings);
/* The same for int32 / float32 columns (see b2_filter_lt_32_dev): raw 32-bit words in and out. */
int b2_filter_lt_32_host_into(b2_ctx* ctx, const void* const* batch_ptrs, const uint8_t* const* valid_ptrs,
                              const int64_t* valid_bit_offsets, const int64_t* batch_lens, int64_t nbatches,
                              int dtype, uint32_t threshold_bits, void* out, int64_t out_capacity,
                              int64_t* out_counts, uint64_t* total, b2_timings* timings);

/* ---- Take (replaces dpu/shared/kernels/take.c:12-47) ------------------------------------- */
/* Batch-local gather, no bounds check (take.c:36, TakeOptions::NoBoundsCheck take_native.cc:27):
 *   d_out[b*idx_len + j] = d_values[b*values_len + d_indices[b*idx_len + j]]. */
int b2_take_u32_dev(b2_ctx* ctx, const uint32_t* d_values, int64_t values_len,
                    const uint32_t* d_indices, int64_t idx_len, int64_t nbatches, uint32_t* d_out,
                    void* stream);
/* Nullable take with Arrow's semantics (cp::Take, take_native.cc:27): output j is null when index
 * j is null or the value it selects is null. d_values_valid spans the packed values column,
 * d_indices_valid / d_out_valid the packed index / output positions; either input bitmap may be
 * NULL (no nulls); d_out_valid (nbatches*idx_len bits, padded to 4 bytes) may only be NULL when
 * both are. Null slots of d_out hold 0. */
int b2_take_u32_nullable_dev(b2_ctx* ctx, const uint32_t* d_values, const uint8_t* d_values_valid,
                             int64_t values_len, const uint32_t* d_indices, const uint8_t* d_indices_valid,
                             int64_t idx_len, int64_t nbatches, uint32_t* d_out, uint8_t* d_out_valid,
                             void* stream);
/* Ragged variant: device offset tables of nbatches+1 int64 each; processes the packed index
 * positions [idx_begin, idx_end) (0 .. idx_off[nbatches] for the whole column). */
int b2_take_u32_ragged_dev(b2_ctx* ctx, const uint32_t* d_values, const int64_t* d_values_off,
                           const uint32_t* d_indices, const int64_t* d_idx_off, int64_t nbatches,
                           int64_t idx_begin, int64_t idx_end, uint32_t* d_out, void* stream);
/* Nullable take over host batches of equal lengths: out_ptrs[b] (idx_lens[b] rows) and
 * out_valid_ptrs[b] ((idx_lens[b] + 7) / 8 bytes, bit offset 0) receive batch b's result. */
int b2_take_u32_nullable_host(b2_ctx* ctx, const uint32_t* const* value_ptrs,
                              const uint8_t* const* value_valid_ptrs, const int64_t* value_valid_bit_offsets,
                              const int64_t* value_lens, const uint32_t* const* idx_ptrs,
                              const uint8_t* const* idx_valid_ptrs, const int64_t* idx_valid_bit_offsets,
                              const int64_t* idx_lens, int64_t nbatches, uint32_t* const* out_ptrs,
                              uint8_t* const* out_valid_ptrs, b2_timings* timings);
/* The same gather over 64-bit values (uint64 / int64 / float64 as raw 8-byte words; SURVEY.md
 * section 8f-3, the reference fixes T = uint32_t, dpu/shared/common.h:3) with 32-bit indices:
 *   d_out[b*idx_len + j] = d_values[b*values_len + d_indices[b*idx_len + j]]   (8-byte elements).
 * d_values and d_out must be 8-byte aligned. b2_take_64_host: batches of any lengths, non-null. */
int b2_take_64_dev(b2_ctx* ctx, const void* d_values, int64_t values_len, const uint32_t* d_indices,
                   int64_t idx_len, int64_t nbatches, void* d_out, void* stream);
int b2_take_64_host(b2_ctx* ctx, const void* const* value_ptrs, const int64_t* value_lens,
                    const uint32_t* const* idx_ptrs, const int64_t* idx_lens, int64_t nbatches,
                    void* const* out_ptrs, b2_timings* timings);
/* TakeDpu::Run (host/take/take_dpu.cc:34-104): out_ptrs[b] has capacity idx_lens[b]. */
int b2_take_u32_host(b2_ctx* ctx, const uint32_t* const* value_ptrs, const int64_t* value_lens,
                     const uint32_t* const* idx_ptrs, const int64_t* idx_lens, int64_t nbatches,
                     uint32_t* const* out_ptrs, b2_timings* timings);

/* ---- Partition (replaces dpu/shared/kernels/partition.c:296-341) ------------------------- */
/* The reference's hash and bucket: wang_hash_uint32 (partition.c:20-28), radix bucket =
 * hash >> (32 - log2 P) (BUCKET_OF, partition.c:45-46). Host-callable so tests can pin it. */
uint32_t b2_wang_hash_u32(uint32_t key);
/* Radix-partition n rows into nparts (power of two, 1..2^20) partitions:
 *   bucket(key) = (wang_hash(key) << skip_bits) >> (32 - log2 nparts)
 * skip_bits > 0 skips hash bits already consumed by an outer partitioning (GPU id, pass 1).
 * ncols columns (1..16): d_cols_in[0] is the key column; all columns are permuted alike into
 * d_cols_out (key column + permutation first, then one gather per column — the reference's
 * PartitionKernel + TakeKernel split, partitioner.cc:119-207). n must be < 2^32. d_part_off (int64, nparts+1) receives the partition boundaries in rows.
 * Row order inside a partition is unspecified (as in the reference, which scatters under a
 * mutex, partition.c:174-231). d_cols_in/d_cols_out are HOST arrays of device pointers. */
size_t b2_partition_ws_bytes(int64_t n, int nparts);
int b2_partition_u32_dev(b2_ctx* ctx, const uint32_t* const* d_cols_in, uint32_t* const* d_cols_out,
                         int ncols, int64_t n, int nparts, int skip_bits, int64_t* d_part_off,
                         void* d_ws, size_t ws_bytes, void* stream);
/* PartitionDpu::Run (host/partition/partition_dpu.cc:31-135): host batches in; the library keeps
 * the partitioned columns on the device until fetched. part_rows[nparts] receives the sizes;
 * b2_partition_fetch_host copies partition p / column c into out_ptrs[p*ncols + c]. */
int b2_partition_u32_host(b2_ctx* ctx, const uint32_t* const* col_batch_ptrs /*[ncols*nbatches], column-major*/,
                          const int64_t* batch_lens, int64_t nbatches, int ncols, int key_col,
                          int nparts, int64_t* part_rows, b2_timings* timings);
int b2_partition_fetch_host(b2_ctx* ctx, uint32_t* const* out_ptrs, int nparts, int ncols,
                            b2_timings* timings);

/* ---- Join (replaces kernel_hash_build/probe hash_build.c:9-35, hash_probe.c:9-46,
 *            ht_put/ht_get hashtable.c:89-192, and JoinDpu::Run_internal join_dpu.cc:168-400) -- */
/* Inner equi-join L.fk = R.pk with Arrow hash-join semantics (join_native.cc:31-36): every
 * (l, r) pair with equal keys yields one output row (fk, y, x); unmatched rows are dropped;
 * duplicate build keys yield one row per duplicate. Output row order is unspecified
 * (partition-major, as JoinDpu's); parity is on the sorted multiset.
 *   L = (d_fk, d_y) nl rows; R = (d_pk, d_x) nr rows; output columns d_out_fk/d_out_y/d_out_x
 *   with capacity out_capacity rows each; *d_out_rows (uint64 on the device) = rows produced.
 * If more than out_capacity rows match, the rows beyond capacity are not written but
 * *d_out_rows still holds the true count (caller compares it with its capacity).
 * *d_out_rows == UINT64_MAX reports that a hash-space slice overflowed its buffer (only possible
 * when the workspace forced slicing and the key distribution is heavily skewed).
 * hash_skip_bits: top hash bits already consumed by an outer routing step (log2 #GPUs in the
 * sharded join, 0 otherwise).
 * d_ws: 256 B aligned workspace. b2_join_ws_bytes() is the size at which the join runs in one
 * go; with less (down to b2_join_min_ws_bytes()) it runs in 2..64 hash-space slices, re-reading
 * the inputs once per slice.
 * Adjacent output columns: when d_out_fk, d_out_y, d_out_x are ONE allocation (d_out_y == d_out_fk +
 * out_capacity, d_out_x == d_out_y + out_capacity, 32 B aligned, 12 * out_capacity >= 8 * max(nl, nr)),
 * the library uses them as the temporary of the first radix pass — they are dead until the probe
 * writes them (the reference aliases its outputs onto the partitioned left side the same way,
 * join_dpu.cc:315-322). b2_join_ws_bytes_adjacent_outputs() is the one-go workspace size then:
 * 8 * max(nl, nr) bytes less. It is what lets SF=2048 (2^32 rows per side) run unsliced on one B200. */
size_t b2_join_ws_bytes(int64_t nl, int64_t nr);
size_t b2_join_ws_bytes_adjacent_outputs(int64_t nl, int64_t nr);
size_t b2_join_min_ws_bytes(int64_t nl, int64_t nr);
/* b2_join_u32_phased_dev: the same join in two calls, so a caller can start on the build side while
 * the probe side is still being uploaded (b2_join_u32_host does: the probe side's H2D copy runs under
 * the build side's radix passes). phases & 1: reset + build side's passes (reads d_pk / d_x only);
 * phases & 2: probe side's passes, probe, *d_out_rows. Same arguments and workspace in both calls;
 * 3 = b2_join_u32_dev. */
int b2_join_u32_phased_dev(b2_ctx* ctx, const uint32_t* d_fk, const uint32_t* d_y, int64_t nl,
                           const uint32_t* d_pk, const uint32_t* d_x, int64_t nr, uint32_t* d_out_fk,
                           uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                           uint64_t* d_out_rows, int hash_skip_bits, int phases, void* d_ws, size_t ws_bytes,
                           void* stream);
int b2_join_u32_dev(b2_ctx* ctx, const uint32_t* d_fk, const uint32_t* d_y, int64_t nl,
                    const uint32_t* d_pk, const uint32_t* d_x, int64_t nr, uint32_t* d_out_fk,
                    uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                    uint64_t* d_out_rows, int hash_skip_bits, void* d_ws, size_t ws_bytes,
                    void* stream);
/* Fused pipeline [filter ->] join -> aggregate (the reference only hints at
 * operator fusion, host/aggr/aggr_native.cc:59-65):
 *   SELECT COUNT(*), SUM(L.y), SUM(R.x) FROM L JOIN R ON L.fk = R.pk [WHERE L.y < y_threshold]
 * with the join's semantics above. Nothing is materialised: no filtered column, no output columns
 * (48 GiB at SF=2048) and no output-range bookkeeping — the probe kernel adds every output row's y
 * and x to its sums. filter_y == 0 ignores y_threshold. Sums are modulo 2^64; rows == ~0 reports a
 * skewed hash-space slice as b2_join_u32_dev does. Workspace as b2_join_u32_dev. */
typedef struct b2_join_aggr {
  uint64_t rows;
  uint64_t sum_y;
  uint64_t sum_x;
} b2_join_aggr;
int b2_join_aggr_u32_dev(b2_ctx* ctx, const uint32_t* d_fk, const uint32_t* d_y, int64_t nl,
                         const uint32_t* d_pk, const uint32_t* d_x, int64_t nr, int filter_y,
                         uint32_t y_threshold, b2_join_aggr* d_out, int hash_skip_bits, void* d_ws,
                         size_t ws_bytes, void* stream);
/* The same pipeline over host batches (tables laid out as for b2_join_u32_host): one upload, three
 * numbers back. */
int b2_join_aggr_u32_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const int64_t* l_lens,
                          int64_t nl_batches, const uint32_t* const* r_ptrs, const int64_t* r_lens,
                          int64_t nr_batches, int filter_y, uint32_t y_threshold, b2_join_aggr* out,
                          b2_timings* timings);
/* Same join over rows packed as 8-byte pairs, key in the low and payload in the high 32 bits
 * (little endian: {uint32 key; uint32 payload}) — the layout the multi-GPU shuffle delivers. */
int b2_join_pairs_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, int64_t nl, const uint64_t* d_r_pairs,
                      int64_t nr, uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x,
                      int64_t out_capacity, uint64_t* d_out_rows, int hash_skip_bits, void* d_ws,
                      size_t ws_bytes, void* stream);
/* JoinDpu::Run: host batches in; result stays on the device until fetched.
 * l_ptrs = [fk batches..., y batches...] (2*nl_batches), r_ptrs = [pk..., x...] (2*nr_batches). */
int b2_join_u32_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const int64_t* l_lens,
                     int64_t nl_batches, const uint32_t* const* r_ptrs, const int64_t* r_lens,
                     int64_t nr_batches, uint64_t* out_rows, b2_timings* timings);
int b2_join_fetch_host(b2_ctx* ctx, uint32_t* out_fk, uint32_t* out_y, uint32_t* out_x,
                       int64_t capacity_rows, b2_timings* timings);

/* Join phase timers (the reference's JoinDpu timers "build" / "probe" / "take" / "partitionKernel",
 * host/join/join_dpu.cc:146-148, read by join_benchmark.cc). b2_join_trace(ctx, 1) makes every join
 * launched through the ctx — *_dev and *_host entry points alike — record CUDA events at its phase
 * boundaries on the stream it runs on; b2_join_last_phases waits for the events of the LAST join and
 * reports the time between them, summed per phase (a sliced join, or a probe side fed in shares, has
 * several intervals per phase). Where the DPU design has "build" and "probe" as separate launches,
 * the probe kernel here builds a partition's table and probes it in one go: probe_ms covers both.
 * Off by default (no events, no cost). */
typedef struct b2_join_phases {
  double partition_build_ms; /* radix passes over the build side (R: pk, x) */
  double partition_probe_ms; /* radix passes over the probe side (L: fk, y), incl. a pushed-down filter */
  double probe_ms;           /* table build + probe + output rows */
  double take_ms;            /* payload gathers of the multi-column / typed-table joins */
  int32_t intervals;         /* intervals summed */
  int32_t reserved;
} b2_join_phases;
int b2_join_trace(b2_ctx* ctx, int on);
int b2_join_last_phases(b2_ctx* ctx, b2_join_phases* out);

/* The same join with NULLABLE key columns: Arrow's inner hash join — the
 * reference's oracle, join_native.cc:31-36 — never matches a null key on either side (the DPU path has
 * no bitmaps at all). *_key_valid_ptrs[b] = validity bitmap of batch b's KEY column starting at bit
 * *_key_valid_bit_offsets[b] (each may be NULL: no nulls / offset 0); payload columns are non-null.
 * Rows with a null key are dropped on the device before the join (row numbers -> nullable filter ->
 * take); a side without a bitmap skips that. Result via b2_join_fetch_host. */
int b2_join_u32_nullable_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const uint8_t* const* l_key_valid_ptrs,
                              const int64_t* l_key_valid_bit_offsets, const int64_t* l_lens, int64_t nl_batches,
                              const uint32_t* const* r_ptrs, const uint8_t* const* r_key_valid_ptrs,
                              const int64_t* r_key_valid_bit_offsets, const int64_t* r_lens, int64_t nr_batches,
                              uint64_t* out_rows, b2_timings* timings);
/* The same join over ANY number of payload columns per side (JoinDpu partitions every left value
 * column, join_dpu.cc:127-138, and takes every right non-key column, :325-341):
 *   l_ptrs = [fk batches..., payload 0 batches..., payload 1 batches..., ...]  ((1 + nl_payloads) * nl_batches)
 *   r_ptrs likewise. Result columns, in order: fk, the left payloads, the right payloads.
 * One payload per side is the fast path above (the payload travels with the key). Otherwise the pair
 * carries the row number and every payload column is gathered afterwards with the take kernel — the
 * reference's selection-vector + TakeKernel scheme without its host round trips. Sides are limited to
 * 2^32 - 1 rows then. b2_join_cols_fetch_host copies column c to out_cols[c] (capacity_rows each). */
int b2_join_cols_u32_host(b2_ctx* ctx, const uint32_t* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                          int nl_payloads, const uint32_t* const* r_ptrs, const int64_t* r_lens, int64_t nr_batches,
                          int nr_payloads, uint64_t* out_rows, b2_timings* timings);
int b2_join_cols_fetch_host(b2_ctx* ctx, uint32_t* const* out_cols, int ncols, int64_t capacity_rows,
                            b2_timings* timings);

/* The join over a TYPED table: 32- or 64-bit keys (the reference's table can be built with 64-bit keys,
 * HT_64BIT_KEYS, dpu/shared/hashtable/hashtable.h:14-18) and any number of 32- or 64-bit payload
 * columns per side (raw words: uint / int / float alike). Column 0 of each side is its key.
 *   l_ptrs = column-major batch pointers, (nl_cols * nl_batches) entries; l_col_bytes[c] = 4 or 8
 * Rows travel through the radix passes by reference (the pair carries the row number) and every
 * result column is gathered afterwards; a 64-bit key is folded to 32 bits for partitioning / probing
 * and every candidate pair is verified against the full keys on the device. Result columns, in order:
 * the key, the left payloads, the right payloads, each of its own width; b2_join_table_fetch_host copies
 * column c to out_cols[c] (capacity_rows elements each). */
int b2_join_table_host(b2_ctx* ctx, const void* const* l_ptrs, const int64_t* l_lens, int64_t nl_batches,
                       const int* l_col_bytes, int nl_cols, const void* const* r_ptrs, const int64_t* r_lens,
                       int64_t nr_batches, const int* r_col_bytes, int nr_cols, uint64_t* out_rows,
                       b2_timings* timings);
int b2_join_table_fetch_host(b2_ctx* ctx, void* const* out_cols, int ncols, int64_t capacity_rows, b2_timings* timings);

/* ---- multi-GPU join exchange (the step that replaces the reference's host-mediated
 *      DPU->host->DPU repartition, partitioner.cc:350-375 + join_dpu.cc:269,293) -------------- */
/* Destination rank of a key when the join is sharded over nranks (power of two) GPUs:
 * top log2(nranks) bits of wang_hash(key); -1 when nranks is not a power of two. */
int b2_join_dest_rank(uint32_t key, int nranks);
/* Route n (key, val) rows by destination rank: d_pairs_out receives the rows as 8-byte pairs
 * grouped by destination, d_dest_off (int64, nranks+1) the group boundaries — the send counts of
 * the all-to-all. The receiver joins with b2_join_pairs_dev(hash_skip_bits = log2 nranks). */
size_t b2_shuffle_ws_bytes(int64_t n, int nranks);
int b2_shuffle_partition_u32_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n,
                                 int nranks, uint64_t* d_pairs_out, int64_t* d_dest_off, void* d_ws,
                                 size_t ws_bytes, void* stream);

/* ---- fused multi-GPU shuffle over peer memory (NVLink stores from inside the scatter kernel) ----
 * Instead of grouping rows by destination and handing them to an all-to-all, every rank
 *   1. counts its rows per bucket, bucket = top `bits` bits of wang_hash(key), where
 *      bits = log2(nranks) + coarse_bits <= 10: the top log2(nranks) bits pick the destination
 *      rank, the next coarse_bits the coarse partition there (b2_shuffle_p2p_count_dev;
 *      d_bucket_off int64[2^bits + 1] = boundaries of the local buckets);
 *   2. after the ranks have exchanged these counts (one small all-gather) and turned them into
 *      destination addresses, scatters its (key, payload) pairs DIRECTLY to
 *      d_bucket_addr[b] (device array of 2^bits byte addresses, usually inside the peers' receive
 *      buffers: CUDA IPC / symmetric-memory pointers) — b2_shuffle_p2p_scatter_dev. Rows of bucket
 *      b are written contiguously from d_bucket_addr[b] on; nothing is bounds-checked, the caller
 *      sizes the receive buffers from the exchanged counts.
 * The receiver therefore finds its rows ALREADY partitioned into 2^coarse_bits coarse buckets
 * (bucket-major, source-rank-minor) and joins them with b2_join_pairs_seg_dev, which only runs
 * the fine partitioning pass. This replaces b2_shuffle_partition + all-to-all + the receiver's
 * first partitioning pass, i.e. the host-mediated repartition of partitioner.cc:350-375.
 * Both calls must use the same n / bits / workspace; scatter reuses the scanned histogram. */
size_t b2_shuffle_p2p_ws_bytes(int64_t n, int bits);
int b2_shuffle_p2p_count_dev(b2_ctx* ctx, const uint32_t* d_key, int64_t n, int bits, int64_t* d_bucket_off,
                             void* d_ws, size_t ws_bytes, void* stream);
/* Step 2a, on the device: turns every rank's bucket boundaries into THIS rank's destination
 * addresses (one launch; no host round trip, no eager tensor arithmetic on the step's critical path).
 *   d_off_ptrs   device array of nranks pointers; entry s -> int64[2^bits + 1], the d_bucket_off of
 *                source rank s (local copies after an all-gather, or peer pointers inside one process)
 *   d_recv_base  device array of nranks byte addresses: every rank's receive buffer as THIS rank
 *                addresses it (CUDA IPC / peer pointers). Layout of a receive buffer: bucket-major,
 *                source-rank-minor, so every coarse bucket is contiguous
 *   capacity_rows rows a receive buffer holds
 * Outputs (device): d_bucket_addr uint64[2^bits] for b2_shuffle_p2p_scatter_dev; d_seg_off
 * int64[2^(bits - log2 nranks) + 1], boundaries of the coarse buckets THIS rank receives; d_info
 * int64[3] = {rows this rank receives, largest receive count of any rank, overflow flag}. When any
 * rank would receive more than capacity_rows the flag is 1 on EVERY rank (all see the same counts),
 * d_seg_off is all zero, and a scatter / join given &d_info[2] as d_abort stores nothing / reports ~0
 * rows: skew surfaces as an error on all ranks instead of a buffer overrun on one.
 * d_prev_abort (may be NULL): an earlier plan's flag, OR-ed into this one — plan the probe side with
 * NULL, the build side with the probe side's &d_info[2], and hand the build side's flag to the join. */
int b2_shuffle_p2p_plan_dev(b2_ctx* ctx, const int64_t* const* d_off_ptrs, const uint64_t* d_recv_base, int rank,
                            int nranks, int bits, int64_t capacity_rows, uint64_t* d_bucket_addr,
                            int64_t* d_seg_off, int64_t* d_info, const int64_t* d_prev_abort, void* stream);
/* d_abort (may be NULL): device int64; non-zero = store nothing (see b2_shuffle_p2p_plan_dev). */
int b2_shuffle_p2p_scatter_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n,
                               int bits, const uint64_t* d_bucket_addr, const int64_t* d_abort, void* d_ws,
                               size_t ws_bytes, void* stream);
/* The same two steps with a predicate pushed in front of the link (the fused pipeline filter -> join ->
 * aggregate over N GPUs): rows whose VALUE fails `value < val_threshold` are neither counted nor sent, so
 * a 25 % predicate moves a quarter of the probe side over NVLink. filter_val == 0: exactly the calls
 * above (d_val may then be NULL in the count). Count and scatter must agree on the predicate. */
int b2_shuffle_p2p_count_lt_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n, int bits,
                                int filter_val, uint32_t val_threshold, int64_t* d_bucket_off, void* d_ws,
                                size_t ws_bytes, void* stream);
int b2_shuffle_p2p_scatter_lt_dev(b2_ctx* ctx, const uint32_t* d_key, const uint32_t* d_val, int64_t n, int bits,
                                  int filter_val, uint32_t val_threshold, const uint64_t* d_bucket_addr,
                                  const int64_t* d_abort, void* d_ws, size_t ws_bytes, void* stream);
/* Join of sides that are already grouped into 2^seg_bits coarse buckets on hash bits
 * [hash_skip_bits, hash_skip_bits + seg_bits); d_*_seg_off (int64, 2^seg_bits + 1, device) hold the
 * bucket boundaries in rows. PRECONDITION (the fused shuffle establishes it): every row of both sides
 * has the same top hash_skip_bits bits of wang_hash(key), and sits in the coarse bucket its next
 * seg_bits hash bits name — the probe kernel's perfect-hash table identifies a key by its remaining
 * hash bits alone (csrc/join.cu). Same output contract as b2_join_pairs_dev. One fine pass refines a
 * coarse bucket at most 2^10-fold; build sides beyond 2^(seg_bits + 22) rows still join correctly
 * (oversized partitions are built in chunks) but more slowly. */
size_t b2_join_seg_ws_bytes(int64_t nl, int64_t nr, int hash_skip_bits, int seg_bits);
int b2_join_pairs_seg_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off, int64_t nl,
                          const uint64_t* d_r_pairs, const int64_t* d_r_seg_off, int64_t nr, int seg_bits,
                          uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                          uint64_t* d_out_rows, int hash_skip_bits, void* d_ws, size_t ws_bytes,
                          void* stream);

/* The same join when the caller only knows CAPACITIES of its receive buffers (the fused shuffle
 * without a host read-back): nl_cap / nr_cap bound the rows (the real counts are the last entries of
 * the segment tables, on the device), nr_expected (the build rows a rank receives when the hash
 * spreads evenly; 0 = nr_cap) picks the number of fine partitions. d_abort (may be NULL): device
 * int64, non-zero = the exchange was called off, *d_out_rows = UINT64_MAX. */
size_t b2_join_seg_cap_ws_bytes(int64_t nl_cap, int64_t nr_cap, int64_t nr_expected, int hash_skip_bits,
                                int seg_bits);
int b2_join_pairs_seg_cap_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off, int64_t nl_cap,
                              const uint64_t* d_r_pairs, const int64_t* d_r_seg_off, int64_t nr_cap,
                              int64_t nr_expected, int seg_bits, uint32_t* d_out_fk, uint32_t* d_out_y,
                              uint32_t* d_out_x, int64_t out_capacity, uint64_t* d_out_rows, int hash_skip_bits,
                              const int64_t* d_abort, void* d_ws, size_t ws_bytes, void* stream);
/* ... and in PHASES, for a probe side that is still arriving while the build side is already there:
 *   phases & 1  build:  reset the join state, fine partitioning pass of the build side (kept in d_ws)
 *   phases & 2  probe:  fine pass of d_l_pairs + probe; may be called several times, each with another
 *                       share of the probe side (its own buffer and segment table); output rows append
 *   phases & 4  finish: publish *d_out_rows
 * 7 = b2_join_pairs_seg_cap_dev. Same d_ws, build-side arguments, nr_expected and bit counts in every
 * call; nl_cap may differ per probe call (the workspace must cover the largest). Between the calls the
 * caller orders its stream after the probe share's arrival (cudaStreamWaitEvent): the probe side's
 * NVLink scatter then runs under the build side's fine pass, and the second share's scatter under the
 * first share's probe (dpu_olap_b200/sharded.py::P2PShuffleJoin). */
int b2_join_pairs_seg_cap_phased_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off,
                                     int64_t nl_cap, const uint64_t* d_r_pairs, const int64_t* d_r_seg_off,
                                     int64_t nr_cap, int64_t nr_expected, int seg_bits, uint32_t* d_out_fk,
                                     uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_capacity,
                                     uint64_t* d_out_rows, int hash_skip_bits, const int64_t* d_abort, int phases,
                                     void* d_ws, size_t ws_bytes, void* stream);
/* The fused join -> aggregate pipeline (b2_join_aggr_u32_dev) over the receiving half of the shuffle:
 * same buffers, segment tables, phases and abort flag as b2_join_pairs_seg_cap_phased_dev, no output
 * columns — the probe kernel adds every output row's payloads to its sums and, with filter_y, only
 * counts probe rows whose payload is < y_threshold (the rows crossed the link unfiltered, so the
 * predicate is evaluated here). phases & 4 publishes *d_out (rows == ~0: overflow or aborted exchange). */
int b2_join_aggr_pairs_seg_cap_phased_dev(b2_ctx* ctx, const uint64_t* d_l_pairs, const int64_t* d_l_seg_off,
                                          int64_t nl_cap, const uint64_t* d_r_pairs, const int64_t* d_r_seg_off,
                                          int64_t nr_cap, int64_t nr_expected, int seg_bits, int filter_y,
                                          uint32_t y_threshold, b2_join_aggr* d_out, int hash_skip_bits,
                                          const int64_t* d_abort, int phases, void* d_ws, size_t ws_bytes,
                                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200OLAP_H_ */
