// pending.h — device-resident result of the last *_host call (internal, non-ABI).
#pragma once
#include <stdint.h>

#include <vector>

struct b2_ctx;
struct b2_pending {
  enum Kind { kNone, kFilter, kPartition, kJoin, kJoinCols, kJoinTable } kind = kNone;
  std::vector<void*> dev;          // device allocations owned by the pending result
  // filter
  uint32_t* d_out = nullptr;
  std::vector<int64_t> batch_end;  // inclusive running counts per batch
  // partition
  std::vector<uint32_t*> d_cols;
  std::vector<int64_t> part_off;
  int ncols = 0;
  // join over a typed table (b2_join_table_host): result columns and their element sizes
  std::vector<void*> t_cols;
  std::vector<int> t_bytes;
  // join
  uint32_t* d_fk = nullptr;
  uint32_t* d_y = nullptr;
  uint32_t* d_x = nullptr;
  uint64_t rows = 0;
};
void b2_pending_free(b2_ctx* ctx);
