// join_l2.cuh — internal (non-ABI) interface of the L2-resident build/probe kernel (join_l2.cu),
// used by the join drivers in join.cu.
#pragma once
#include "partition.cuh"

struct JoinState {  // lives in the workspace header
  unsigned long long out_rows;
  unsigned int overflow;
  unsigned int pad;
  unsigned long long phase_ns[4];  // join_l2_kernel: time in clear / build / probe phases, grid syncs counted
};

// Geometry of the global-memory hash table one group of build rows is inserted into. A bucket is
// one 32-byte sector: [count, pad, (key, value) x 3].
struct JoinL2Geom {
  uint32_t nbuckets = 0;  // buckets of the table
  int64_t cap_rows = 0;   // build rows inserted per chunk (groups beyond it are built in chunks)
  size_t table_bytes = 0;
};
constexpr int kJoinL2SlotsPerBucket = 3;

// Table for groups of `group_rows` expected build rows (mean fill per bucket = fill_x100 / 100).
JoinL2Geom join_l2_geom(int64_t group_rows, int fill_x100);

// Joins ngroups co-partitioned groups: build rows of group g are r[roff[g] .. roff[g+1]), probe
// rows l[loff[g] .. loff[g+1]) (offsets on the device). One persistent cooperative launch walks
// the groups: clear table, insert the group's build rows, stream its probe rows through the table.
// Appends (fk, y, x) rows at st->out_rows (rows beyond out_cap are counted but not stored).
int join_l2_run(b2_ctx* ctx, const PartInput& rin, const int64_t* d_roff, const PartInput& lin,
                const int64_t* d_loff, int64_t ngroups, const JoinL2Geom& geom, void* d_table,
                uint32_t* d_out_fk, uint32_t* d_out_y, uint32_t* d_out_x, int64_t out_cap, JoinState* st,
                cudaStream_t s);
