// sweep_internal.h — what gather.cu needs from an am_sweep handle (same shared library;
// not part of the C-ABI): the buffers one tick of the shard left behind and the events
// that order their producers and consumers across streams.
#pragma once
#ifndef AMSWEEP_EMULATE
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "sweep_types.h"

struct am_sweep;

namespace amsweep {

struct ShardTick {
  TickOut out;               // bitmap, group_prefix, tile_exc, exc_seg of the handle's last tick
  unsigned long long* acc;   // that tick's statistics accumulators
  uint64_t shard_base, n_records;
  uint32_t n_groups, n_tiles;
  int parity;                // which of the handle's two buffer sets
  int device;
};

// The buffers of the last am_sweep_tick_shard.  Returns false if the handle has not ticked.
bool shard_last_tick(am_sweep* h, ShardTick* out);
// Make stream `s` wait for the handle's last tick (issued on another stream, perhaps) without
// making the handle's next tick wait for `s`: the exchange of tick k overlaps the sweep of k+1.
int shard_order_consumer(am_sweep* h, cudaStream_t s);
// The consumer (exchange + list rebuild on `s`) is done with buffer set `parity`: the next
// tick that reuses the set waits for this point.
int shard_mark_consumed(am_sweep* h, int parity, cudaStream_t s);
// The list rebuild and the statistics publication, launched from sweep.cu (where the kernels
// are compiled) on behalf of the exchange.
int shard_launch_expand(am_sweep* h, const ExpandParams& e, uint32_t groups_x, uint32_t world_y, cudaStream_t s);
int shard_launch_publish(am_sweep* h, unsigned long long* acc, am_tick_stats_t* out_stats, uint64_t n_records,
                         cudaStream_t s);
// n_groups / n_tiles a shard of n records occupies
inline uint32_t tiles_of(uint64_t n) { return (uint32_t)((n + kTile - 1) / kTile); }
inline uint32_t groups_of(uint64_t n) { return (uint32_t)((n + kGroupRecords - 1) / kGroupRecords); }

}  // namespace amsweep
