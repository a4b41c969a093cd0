// emu_gather.cpp — runs the exchange kernels of csrc/gather_kernels.cuh on the CPU emulator
// (cuda_emu.h): `world` ranks as OS threads, each with its own exchange block, pushing
// `epochs` ticks back to back with no host-side barrier in between (as the stream-ordered
// product does), in one of the wire formats.  After every push each rank checks that its
// output is the rank-ordered concatenation of all ranks' lists.
//
//   emu_gather <wire: plain|c3|bm> <world> <idx_bytes: 4|8> <records per rank> <epochs> <ctas> [density% [cap]]
//   EMU_ABSENT_RANK=r (bm only): rank r never pushes; the others must give up after the watchdog
//   time-out (200 ms here) and report kPeerTimeout instead of spinning for ever.
// `cap` (optional) shrinks the output capacity below the total, to exercise the truncation
// paths: only the first `cap` entries of the concatenation exist then.
#include "cuda_emu.h"

#include <cstdio>
#include <cstring>
#include <random>
#include <string>

#include "../../active-monitor_b200/csrc/gather_kernels.cuh"

namespace {

struct Lists { std::vector<uint32_t> idx; std::vector<uint8_t> act; };

// the local emitted list of (rank, epoch): deterministic, so every rank can rebuild all of them
Lists make_list(int rank, int epoch, uint32_t n_records, int density_pct) {
  std::mt19937_64 rng(0x9E3779B97F4A7C15ull * (uint64_t)(rank + 1) + (uint64_t)epoch * 1000003ull);
  Lists l;
  // alternate dense / sparse / empty stretches so that groups with 0, few and 8192 entries occur
  for (uint32_t i = 0; i < n_records; ++i) {
    const uint32_t stretch = (i / 3000u + (uint32_t)rank + (uint32_t)epoch) % 5u;
    int p = density_pct;
    if (stretch == 0) p = 0;
    if (stretch == 1) p = 100;
    if ((int)(rng() % 100) < p) {
      l.idx.push_back(i);
      const uint64_t r = rng() % 1000;
      // mostly the bare submit; some other action bytes, including 0x01-containing and 0x80
      l.act.push_back(r < 970 ? 0x01 : (uint8_t)(r < 985 ? 0x08 : (r < 995 ? 0x23 : 0x80)));
    }
  }
  if (epoch % 4 == 3 && rank == 1) { l.idx.clear(); l.act.clear(); }  // an empty list now and then
  return l;
}

struct Rank {
  unsigned char* block = nullptr;
  uint32_t* out_counts = nullptr;
  void* final_idx[2] = {nullptr, nullptr};
};

int fail(const char* what, int rank, int epoch, uint64_t pos) {
  std::printf("MISMATCH %s rank %d epoch %d pos %llu\n", what, rank, epoch, (unsigned long long)pos);
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 7) { std::printf("usage: emu_gather wire world idx_bytes records epochs ctas [density]\n"); return 2; }
  const std::string wire = argv[1];
  const int world = std::atoi(argv[2]), idx_bytes = std::atoi(argv[3]);
  const uint32_t n_rec = (uint32_t)std::atoll(argv[4]);
  const int epochs = std::atoi(argv[5]), ctas = std::atoi(argv[6]);
  const int density = argc > 7 ? std::atoi(argv[7]) : 33;
  const uint64_t cap_override = argc > 8 ? (uint64_t)std::atoll(argv[8]) : 0;
  if (world < 1 || world > kMaxWorld || (idx_bytes != 4 && idx_bytes != 8)) return 2;

  // shards of slightly different sizes (the last one is ragged), as shard_range() produces
  std::vector<uint64_t> bases(world), sizes(world);
  uint64_t cap_total = 0;
  for (int r = 0; r < world; ++r) { bases[r] = cap_total; sizes[r] = n_rec + (r == world - 1 ? 37 : 0); cap_total += sizes[r]; }
  const uint64_t n_total = cap_total;
  const uint32_t ngroups_max = (uint32_t)((cap_total + kGroupRecords - 1) / kGroupRecords);
  std::vector<uint32_t> ngroups(world);
  std::vector<uint64_t> bm_word0(world);
  uint64_t gb = 0;
  for (int r = 0; r < world; ++r) {
    ngroups[r] = (uint32_t)((sizes[r] + kGroupRecords - 1) / kGroupRecords);
    bm_word0[r] = gb * kGroupWords;
    gb += ngroups[r];
  }
  // block layout (any non-overlapping, aligned layout works: the kernels only see offsets)
  auto align = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t off = align(sizeof(ExchangeHeader));
  size_t off_idx[2], off_act[2], off_gc[2], off_o16[2], off_bm[2];
  for (int b = 0; b < 2; ++b) { off_idx[b] = off; off = align(off + cap_total * 8); }
  for (int b = 0; b < 2; ++b) { off_act[b] = off; off = align(off + cap_total + 4); }
  for (int b = 0; b < 2; ++b) { off_gc[b] = off; off = align(off + (size_t)world * ngroups_max * 4); }
  for (int b = 0; b < 2; ++b) { off_o16[b] = off; off = align(off + cap_total * 2 + 8); }
  for (int b = 0; b < 2; ++b) { off_bm[b] = off; off = align(off + ((size_t)ngroups_max + world) * kGroupWords * 4); }
  const size_t block_bytes = off;
  if (cap_override && cap_override < n_total) cap_total = cap_override;  // buffers stay full-size

  std::vector<Rank> ranks(world);
  for (auto& rk : ranks) {
    rk.block = (unsigned char*)std::aligned_alloc(256, block_bytes);
    std::memset(rk.block, 0, block_bytes);
    // poison the payload areas: stale data must never be mistaken for this epoch's
    std::memset(rk.block + off_idx[0], 0xEE, block_bytes - off_idx[0]);
    rk.out_counts = (uint32_t*)std::calloc(kMaxWorld + 1, 4);
    for (int b = 0; b < 2; ++b) rk.final_idx[b] = std::aligned_alloc(256, align(cap_total * 8));
  }

  std::atomic<int> failures{0};
  const int absent = std::getenv("EMU_ABSENT_RANK") ? std::atoi(std::getenv("EMU_ABSENT_RANK")) : -1;
  auto rank_main = [&](int rank) {
    Rank& me = ranks[rank];
    if (rank == absent) return;
    for (int e = 1; e <= epochs && !failures.load(); ++e) {
      const Lists mine = make_list(rank, e, (uint32_t)sizes[rank], density);
      const uint32_t count = (uint32_t)mine.idx.size();
      const int buf = e & 1;
      if (wire == "plain") {
        PushParams p{};
        for (int r = 0; r < world; ++r) p.peer[r] = ranks[r].block;
        p.idx_local = mine.idx.data(); p.act_local = mine.act.data(); p.count_local = &count;
        p.out_counts = me.out_counts; p.shard_base = bases[rank]; p.cap_total = cap_total;
        for (int b = 0; b < 2; ++b) { p.off_idx[b] = off_idx[b]; p.off_act[b] = off_act[b]; }
        p.epoch = (uint32_t)e; p.rank = rank; p.world = world; p.idx_bytes = idx_bytes;
        emu::launch(gather_push_kernel, dim3(ctas), dim3(256), p);
      } else if (wire == "c3") {
        PushC3Params c{};
        for (int r = 0; r < world; ++r) c.peer[r] = ranks[r].block;
        c.idx_local = mine.idx.data(); c.act_local = mine.act.data(); c.count_local = &count;
        c.out_counts = me.out_counts; c.cap_total = cap_total;
        for (int b = 0; b < 2; ++b) { c.off_act[b] = off_act[b]; c.off_gc[b] = off_gc[b]; c.off_o16[b] = off_o16[b]; }
        c.epoch = (uint32_t)e; c.ngroups_mine = ngroups[rank]; c.ngroups_max = ngroups_max; c.rank = rank; c.world = world;
        emu::launch(gather_push_c3_kernel, dim3(ctas), dim3(256), c);
        DecodeParams d{};
        d.o16 = (const uint16_t*)(me.block + off_o16[buf]); d.gc = (const uint32_t*)(me.block + off_gc[buf]);
        d.counts = me.out_counts; d.final_idx = me.final_idx[buf]; d.cap_total = cap_total;
        d.ngroups_max = ngroups_max; d.world = world; d.idx_bytes = idx_bytes;
        uint32_t ng_used = 1;
        for (int r = 0; r < world; ++r) { d.ngroups[r] = ngroups[r]; d.bases[r] = bases[r]; if (ngroups[r] > ng_used) ng_used = ngroups[r]; }
        emu::launch(gather_decode_kernel, dim3(ng_used, world), dim3(256), d);
      } else {  // bm
        std::memset(me.block + off_act[buf], (int)AM_ACT_SUBMIT_HC, cap_total);  // am_gather_push does this first
        PushBmParams b{};
        for (int r = 0; r < world; ++r) b.peer[r] = ranks[r].block;
        b.idx_local = mine.idx.data(); b.act_local = mine.act.data(); b.count_local = &count;
        b.out_counts = me.out_counts; b.cap_total = cap_total;
        for (int k = 0; k < 2; ++k) { b.off_act[k] = off_act[k]; b.off_gc[k] = off_gc[k]; b.off_bm[k] = off_bm[k]; }
        b.bm_word0 = bm_word0[rank]; b.epoch = (uint32_t)e; b.ngroups_mine = ngroups[rank];
        b.timeout_ns = absent >= 0 ? 200000000ull : 0ull;
        b.ngroups_max = ngroups_max; b.rank = rank; b.world = world;
        emu::launch(gather_push_bm_kernel, dim3(ctas), dim3(256), b);
        ExpandBmParams x{};
        x.bm = (const uint32_t*)(me.block + off_bm[buf]); x.gc = (const uint32_t*)(me.block + off_gc[buf]);
        x.counts = me.out_counts; x.final_idx = me.final_idx[buf]; x.cap_total = cap_total;
        x.ngroups_max = ngroups_max; x.world = world; x.idx_bytes = idx_bytes;
        uint32_t ng_used = 1;
        for (int r = 0; r < world; ++r) {
          x.ngroups[r] = ngroups[r]; x.bases[r] = bases[r]; x.bm_word0[r] = bm_word0[r];
          if (ngroups[r] > ng_used) ng_used = ngroups[r];
        }
        emu::launch(gather_expand_bitmap_kernel, dim3(ng_used, world), dim3(256), x);
        if (absent >= 0) {  // a peer is missing: the watchdog must have fired, nothing else is defined
          if (me.out_counts[world] != kPeerTimeout) failures += fail("expected time-out marker", rank, e, 0);
          return;
        }
      }
      // check: my output == concatenation of every rank's list of this epoch
      const void* out_idx = wire == "plain" ? (const void*)(me.block + off_idx[buf]) : me.final_idx[buf];
      const uint8_t* out_act = me.block + off_act[buf];
      uint64_t pos = 0;
      for (int r = 0; r < world; ++r) {
        const Lists l = r == rank ? mine : make_list(r, e, (uint32_t)sizes[r], density);
        if (me.out_counts[r] != l.idx.size()) { failures += fail("count", rank, e, (uint64_t)r); return; }
        for (size_t k = 0; k < l.idx.size(); ++k, ++pos) {
          if (pos >= cap_total) continue;  // beyond the capacity nothing is defined
          const uint64_t want = bases[r] + l.idx[k];
          const uint64_t got = idx_bytes == 4 ? ((const uint32_t*)out_idx)[pos] : ((const uint64_t*)out_idx)[pos];
          if (got != want) { failures += fail("idx", rank, e, pos); return; }
          if (out_act[pos] != l.act[k]) { failures += fail("act", rank, e, pos); return; }
        }
      }
      if (me.out_counts[world] != (pos < cap_total ? pos : cap_total)) { failures += fail("total", rank, e, pos); return; }
    }
  };
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(rank_main, r);
  for (auto& t : th) t.join();
  if (failures.load()) return 1;
  if (absent >= 0) { std::printf("ok watchdog: rank %d absent, the others gave up\n", absent); return 0; }
  std::printf("ok %s world=%d idx_bytes=%d records=%u epochs=%d ctas=%d\n", wire.c_str(), world, idx_bytes, n_rec, epochs, ctas);
  return 0;
}
