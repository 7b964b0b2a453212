// capi.cu — the extern "C" surface of libdtree_b200.so (include/dtree_b200.h): handle state,
// device mirror of the tree, scratch management and kernel launches. No sampling or social-choice
// arithmetic runs on the host: if CUDA is unusable every sampling entry point fails.
#include <cuda_runtime.h>

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dtree_b200.h"
#include "dtree_b200_debug.h"
#include "host_tree.h"
#include "launch.h"
#include "launch_shape.h"
#include "tree_build.h"

using namespace dtb;

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) return fail(DTREE_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = (dev == prev) || cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Packs ballot `b[0..len)` (truncated to nc-1 preferences: the last one can never matter to
// IRV) into 3 words; unused positions up to nc-1 repeat the first candidate, which is how the
// IRV kernel recognises the end of an observed ballot without a length array.
void pack_ballot(const uint8_t *b, uint32_t len, uint32_t nc, bool pad_with_first, uint64_t out[kMaxKeyWords],
                 uint32_t *packed_len) {
  BallotKey<kMaxKeyWords> k;
  for (int w = 0; w < kMaxKeyWords; ++w) k.w[w] = 0;
  uint32_t L = std::min(len, nc - 1);
  for (uint32_t j = 0; j < L; ++j) key_set<kMaxKeyWords>(k, (int)j, b[j]);
  if (pad_with_first && L > 0)
    for (uint32_t j = L; j < nc - 1; ++j) key_set<kMaxKeyWords>(k, (int)j, b[0]);
  for (int w = 0; w < kMaxKeyWords; ++w) out[w] = k.w[w];
  if (packed_len) *packed_len = L;
}

double falling_factorial(uint32_t n, uint32_t k) {
  double r = 1.0;
  for (uint32_t i = 0; i < k; ++i) r *= (double)(n - i);
  return r;
}

}  // namespace

// Layout of dtree_last_stats (include/dtree_b200.h).
enum OutStat {
  kOutLaunches = 0, kOutElections, kOutEntries, kOutItems, kOutGammas, kOutBinomials, kOutRereads, kOutScratchBytes,
  kOutH2D, kOutD2H, kOutSlots, kOutUrnItems, kOutMode, kOutArenaRuns, kOutBuildLaunches, kOutLaneRetries,
  kOutMirrorUs, kOutPlanUs, kOutLaunchUs, kOutFinishUs, kOutCount = 24
};

// NVTX range over a scope (the reference has no tracing; these show up in Nsight Systems / ncu --nvtx).
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct HostTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  uint64_t us() const {
    return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
  }
};

// Everything a sampling call reads back, contiguous on the device so that one copy into page-locked memory fetches it.
struct ResultBlock {
  unsigned long long wins[kMaxCandidates];
  unsigned long long stats[8];  // kStatCount <= 8
  unsigned int usage[8];        // [0..2] deferred_kernel arena high-water marks, [3] elections re-run in a worst-case
                                // arena, [4] lane_kernel elections handed to the replay, [5..7] lane_kernel high-water marks
  int error;
  unsigned int counter;         // work counter of the running launch
  unsigned int pad_[2];
};
static_assert(kStatCount <= 8, "ResultBlock::stats too small");

// A host range page-locked for direct DMA (the ballot log: it travels to the device whenever the mirror is rebuilt
// from host memory). Failure to lock is not an error: the copy is then staged by the driver as for any pageable memory.
struct PinnedRange {
  void *p = nullptr;
  size_t bytes = 0;
};

static void unpin(PinnedRange &r) {
  if (r.p) {
    cudaHostUnregister(r.p);
    cudaGetLastError();
  }
  r = PinnedRange();
}

static void pin(PinnedRange &r, const void *p, size_t bytes) {
  if (r.p == p && r.bytes == bytes) return;
  unpin(r);
  if (!p || bytes < (1u << 16)) return;  // small logs: not worth the page-locking
  if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) {
    r.p = const_cast<void *>(p);
    r.bytes = bytes;
  } else {
    cudaGetLastError();
  }
}

struct dtree_ballots {
  std::vector<uint8_t> prefs;
  std::vector<uint64_t> offsets{0};
  std::vector<uint32_t> counts;
};

constexpr uint32_t kLenNone = 0xFFFFFFFFu, kLenMixed = 0xFFFFFFFEu;

struct dtree {
  // Observed ballots as given to dtree_update (validated): the device builds the tree from this log
  // (tree_build.cu); the host trie below is brought up to date only for the host-side exports.
  std::vector<uint8_t> log_prefs;
  std::vector<uint64_t> log_offsets{0};
  uint32_t log_len = kLenNone;  // common number of preferences of the ballots in the log (kLenNone: empty, kLenMixed)
  PinnedRange pin_prefs, pin_offsets;  // the two vectors' buffers while page-locked (dropped before they can move)
  uint64_t n_observed = 0, depth_mask = 0, version = 0;  // version: bumped by update/reset
  HostTree tree;              // parameters (tree.P) + lazily synchronised trie
  uint64_t trie_ballots = 0;  // log entries already inserted into `tree`
  uint64_t n_nodes = 1;       // of the current mirror
  bool host_build = false;    // tuning "host_build": mirror = HostTree::flatten + upload (measurement aid)
  int build_mode = 0;         // tuning "build_mode": 0 one cooperative kernel, 1 a launch per phase, 2 exact allocation
  int device;
  int sm_count = 0;
  size_t smem_optin = 0;
  // device mirror
  uint64_t mirrored_version = ~0ull, image_version = ~0ull;
  FlatTree flat;
  std::vector<uint64_t> h_obs_keys;
  std::vector<uint8_t> h_obs_first;
  std::vector<uint32_t> h_obs_tally0;
  DeviceTreeStore dev;  // tree arrays, packed observed ballots, ballot log (tree_build.h)
  // scratch + small device state
  DevBuf<uint8_t> d_scratch;
  DevBuf<ResultBlock> d_res;
  ResultBlock *h_res = nullptr;  // page-locked
  DevBuf<uint32_t> d_entry_count;
  DevBuf<uint8_t> d_orders;
  // deferred kernel: worst-case overflow arenas and their locks; lane kernel: the retry list ([0] count, [4..] elections)
  DevBuf<uint8_t> d_arena;
  DevBuf<unsigned int> d_locks, d_retry;
  struct Sizing {
    bool valid = false;
    uint64_t version = 0;
    uint32_t n_sample = 0, urn_max = 0, min_depth = 0, max_depth = 0;
    double a0 = 0;
    bool vd = false, use_obs = false, short_cuts = false;  // short_cuts: IRV may stop early (smaller elections)
    uint32_t pool = 0, queue = 0, big = 0;              // largest usage seen (items), warp per election
    uint32_t lpool = 0, lqueue = 0, lbig = 0;           // largest usage seen by lane_kernel's own elections
    uint32_t cap_pool = 0, cap_queue = 0, cap_big = 0;  // capacities in force (warp per election)
    uint32_t lane_pool = 0, lane_queue = 0, lane_big = 0;  // capacities in force (thread per election)
    double nodes_per_election = 0;                          // node splits per election, last measured
  } sizing;
  // tuning
  int block_threads = 0, blocks_per_sm = 0;
  uint32_t urn_max = 8;
  uint32_t arena_cap = 0;        // > 0: per-warp / per-thread arenas hold at most this many items (forces the overflow paths)
  int64_t lane_max_splits = -1;  // lane_kernel hands an election to the replay after this many splits (-1 auto, 0 never)
  uint32_t tau_quarters = 3;     // deferred_kernel's resolve threshold (deferred.cuh)
  uint32_t binv_max = 30;        // binomials with n p below this use inversion, the others BTRS
  uint32_t last_lane_retries = 0;  // elections lane_kernel handed to its replay in the last call read back
  uint64_t mem_avail = 0;        // device bytes the plans may use (free + held by this handle), as last queried
  uint32_t lane_width = 0;  // lane_kernel: elections per warp (0: chosen per job)
  uint32_t lane_block_warps = 0;  // lane_kernel: warps per block (0: chosen per job)
  int lane_sync = -1;  // lane_kernel: block-wide meeting points (-1 chosen per job, 0 none, 1 claim + after the root split, 2 + every round, 3 + every split step)
  int sim_phased = -1;  // materialising path as three launches per wave: -1 when it pays off, 0 never, 1 always
  int mode = -1;  // 0: materialise every ballot (simulate_kernel); 1: deferred expansion (deferred_kernel); -1: choose
  // replicas on other devices for dtree_sample_posterior_multi (owned; same parameters, tree copied when stale)
  std::vector<dtree *> replicas;
  // callbacks, stats
  int (*should_abort)(void *) = nullptr;
  void *abort_user = nullptr;
  uint64_t stats[kOutCount] = {0};

  dtree(const HostParams &p, int dev) : tree(p), device(dev) {}
};

namespace {

DeviceParams device_params(const HostParams &P) {
  DeviceParams D;
  memset(&D, 0, sizeof D);
  D.nc = P.nc;
  D.min_depth = P.min_depth;
  D.max_depth = P.max_depth;
  D.leaf_depth = P.max_depth - 1u;  // unsigned wrap when max_depth == 0, as in irv_node.cpp:136
  D.small_alpha = 0;
  D.binv_max = 30.0f;
  for (uint32_t d = 0; d < kMaxCandidates; ++d) {
    double a = P.a0;
    if (P.vd) a *= (d < P.depth_factors.size()) ? P.depth_factors[d] : 1.0;
    D.a_prior[d] = (float)a;
    if (d + 1 < P.nc && a < 1.0) D.small_alpha = 1;
  }
  return D;
}

void unpin_log(dtree *t) {
  if (!t->pin_prefs.p && !t->pin_offsets.p) return;
  DeviceGuard g(t->device);
  unpin(t->pin_prefs);
  unpin(t->pin_offsets);
}

// Brings the host trie up to date with the ballot log.
void sync_trie(dtree *t) {
  const uint64_t n = t->log_offsets.size() - 1;
  if (t->trie_ballots < n) t->tree.insert_ballots(t->log_prefs.data(), t->log_offsets.data(), t->trie_ballots, n);
  t->trie_ballots = n;
}

// Host image of the tree: the CSR arrays plus the packed observed entries, for the host-side exports
// and the "host_build" mirror. Rebuilt only when the tree changes.
int ensure_host_image(dtree *t) {
  if (t->image_version == t->version) return DTREE_OK;
  sync_trie(t);
  t->tree.flatten(t->flat);
  const FlatTree &F = t->flat;
  const uint32_t nc = t->tree.P.nc;
  const int W = key_words_for(nc);
  if (F.slot_count.size() >= 0xFFFFFFF0ull) return fail(DTREE_ERR_STATE, "tree has too many outcome slots for 32-bit indices");
  size_t n = F.obs_counts.size();
  t->h_obs_keys.assign(n * W, 0);
  t->h_obs_first.assign(n, 0);
  t->h_obs_tally0.assign(kMaxCandidates, 0);
  for (size_t i = 0; i < n; ++i) {
    uint64_t k[kMaxKeyWords];
    uint32_t len = (uint32_t)(F.obs_offsets[i + 1] - F.obs_offsets[i]);
    pack_ballot(F.obs_prefs.data() + F.obs_offsets[i], len, nc, true, k, nullptr);
    for (int w = 0; w < W; ++w) t->h_obs_keys[i * W + w] = k[w];
    t->h_obs_first[i] = len ? F.obs_prefs[F.obs_offsets[i]] : 0xFF;
    if (len) t->h_obs_tally0[t->h_obs_first[i]] += F.obs_counts[i];
  }
  t->image_version = t->version;
  return DTREE_OK;
}

// The device mirror. Product path: the tree is built on the device from the ballot log. With the
// tuning "host_build" the host trie is flattened and uploaded instead (kept to measure against).
int ensure_mirror(dtree *t, cudaStream_t s) {
  if (t->mirrored_version == t->version) return DTREE_OK;
  if (!t->host_build) {
    pin(t->pin_prefs, t->log_prefs.data(), t->log_prefs.capacity());
    pin(t->pin_offsets, t->log_offsets.data(), t->log_offsets.capacity() * sizeof(uint64_t));
    BuildStats bs;
    bool too_big = false;
    CUDA_TRY(build_tree_on_device(t->dev, t->tree.P.nc, t->log_prefs.data(), t->log_offsets.data(),
                                  t->log_offsets.size() - 1, s, &bs, &too_big, t->build_mode == 2 ? 0ull : 1ull << 30,
                                  t->build_mode == 0, t->log_len < kLenMixed ? t->log_len : 0xFFFFFFFFu));
    if (too_big) return fail(DTREE_ERR_STATE, "tree has too many outcome slots for 32-bit indices");
    t->stats[kOutH2D] += bs.h2d_bytes;
    t->stats[kOutBuildLaunches] = bs.launches;
    t->n_nodes = t->dev.n_nodes;
    t->mirrored_version = t->version;
    return DTREE_OK;
  }
  int rc = ensure_host_image(t);
  if (rc) return rc;
  const FlatTree &F = t->flat;
  CUDA_TRY(t->dev.node_base.upload(F.node_base, s));
  CUDA_TRY(t->dev.node_total.upload(F.node_total, s));
  CUDA_TRY(t->dev.slot_count.upload(F.slot_count, s));
  CUDA_TRY(t->dev.slot_child.upload(F.slot_child, s));
  CUDA_TRY(t->dev.slot_cand.upload(F.slot_cand, s));
  CUDA_TRY(t->dev.obs_key.upload(t->h_obs_keys, s));
  CUDA_TRY(t->dev.obs_cnt.upload(F.obs_counts, s));
  CUDA_TRY(t->dev.obs_first.upload(t->h_obs_first, s));
  CUDA_TRY(t->dev.obs_tally0.upload(t->h_obs_tally0, s));
  t->dev.n_obs = (uint32_t)F.obs_counts.size();
  t->dev.n_nodes = F.n_nodes();
  t->dev.n_slots = (uint32_t)F.slot_count.size();
  t->dev.log_ballots = t->dev.log_bytes = 0;  // the device log is not maintained on this path
  t->n_nodes = F.n_nodes();
  t->stats[kOutH2D] += (F.node_base.size() + F.node_total.size() + F.slot_count.size() + F.slot_child.size()) * 4 +
                 F.slot_cand.size() + t->h_obs_keys.size() * 8 + F.obs_counts.size() * 5 + kMaxCandidates * 4;
  CUDA_TRY(cudaStreamSynchronize(s));  // pageable sources: the copies must not outlive a later update
  t->mirrored_version = t->version;
  return DTREE_OK;
}

struct Plan {
  int W, block, grid, blocks_per_sm;
  bool log_gamma;
  uint32_t cap_items, cap_entries, cap_queue = 0, cap_big = 0, cap_small = 0;
  uint32_t lane_width = 32;  // lane_kernel: elections per warp
  bool phased = false;  // simulate_kernel's work as three launches per wave of `grid` elections (sim_*_kernel)
  uint32_t n_arenas = 0, arena_items = 0, arena_queue = 0, arena_big = 0;
  uint64_t arena_stride = 0, mem_budget = 0;
  uint64_t stride;        // scratch bytes per block (materialising kernel) or per warp (deferred kernel)
  uint64_t stride_block;  // scratch bytes per block
  bool deferred = false;
  bool lane = false;  // deferred algorithm, one thread per election (lane_kernel)
};

template <int W>
uint64_t scratch_bytes(uint32_t ci, uint32_t ce, uint32_t nobs, uint32_t cs) { return Scratch<W>::bytes(ci, ce, nobs, cs); }

// Launch shape and scratch of the materialising path. allow_phased: the caller runs IRV on every election, so the job
// may be split into the three phase kernels (dumps of a single election always take the fused kernel).
int make_plan(dtree *t, uint32_t n_sample, uint32_t n_elections, bool use_obs, Plan &pl, bool allow_phased = false) {
  const HostParams &P = t->tree.P;
  const uint32_t nc = P.nc;
  pl.W = key_words_for(nc);
  DeviceParams D = device_params(P);
  pl.log_gamma = D.small_alpha != 0;
  // Upper bounds on what one election can hold: never more things than ballots, never more than
  // there are distinct prefixes.
  uint32_t item_depth = std::min<uint32_t>(D.leaf_depth, nc - 2);          // deepest frontier
  uint32_t max_len = std::min<uint32_t>(P.max_depth ? P.max_depth : nc, nc - 1);  // longest sampled ballot
  // the level frontiers hold only items with more than urn_max ballots; the others wait in the small queue, where an
  // item is put once (as many as there are ballots, or prefixes down to the deepest frontier)
  double items = std::min<double>((double)(n_sample / (t->urn_max + 1) + 1), falling_factorial(nc, item_depth));
  double smalls = 0;
  for (uint32_t l = 0; l <= item_depth; ++l) smalls += falling_factorial(nc, l);
  smalls = std::min<double>((double)n_sample, smalls);
  pl.cap_small = (uint32_t)std::max(smalls, 1.0);
  double entries = 0;
  for (uint32_t l = 0; l <= max_len; ++l) entries += falling_factorial(nc, l);
  entries = std::min<double>((double)n_sample, entries);
  pl.cap_items = (uint32_t)std::max(items, 1.0);
  pl.cap_entries = (uint32_t)std::max(entries, 1.0);
  uint32_t nobs = use_obs ? t->dev.n_obs : 0;
  pl.stride = pl.W == 1 ? scratch_bytes<1>(pl.cap_items, pl.cap_entries, nobs, pl.cap_small)
            : pl.W == 2 ? scratch_bytes<2>(pl.cap_items, pl.cap_entries, nobs, pl.cap_small)
                        : scratch_bytes<3>(pl.cap_items, pl.cap_entries, nobs, pl.cap_small);
  pl.stride = (pl.stride + 255) & ~255ull;
  pl.stride_block = pl.stride;

  int block = t->block_threads;
  if (block == 0) {
    // A block works on one election; size it to the widest frontier it is likely to see.
    double work = std::min<double>(smalls, 1.0 + (double)t->n_nodes + (double)n_sample / 8);
    block = work <= 48 ? 32 : work <= 512 ? 64 : work <= 8192 ? 128 : 256;
  }
  bool okb = false;
  for (int i = 0; i < kNumBlockSizes; ++i) okb |= (kBlockSizes[i] == block);
  if (!okb) return fail(DTREE_ERR_ARG, "block_threads must be one of 32, 64, 128, 256");
  pl.block = block;
  int occ = pl.W == 1 ? simulate_blocks_per_sm<1>(nc, block, pl.log_gamma)
          : pl.W == 2 ? simulate_blocks_per_sm<2>(nc, block, pl.log_gamma)
                      : simulate_blocks_per_sm<3>(nc, block, pl.log_gamma);
  if (occ <= 0) return fail(DTREE_ERR_CUDA, "kernel cannot be resident (occupancy query failed: %s)",
                            cudaGetErrorString(cudaGetLastError()));
  if (t->blocks_per_sm > 0) occ = std::min(occ, t->blocks_per_sm);
  pl.blocks_per_sm = occ;
  uint64_t grid = (uint64_t)occ * t->sm_count;
  grid = std::min<uint64_t>(grid, n_elections);
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  uint64_t budget = (uint64_t)((double)(free_b + t->d_scratch.n) * 0.6);
  uint64_t fit = std::max<uint64_t>(1, budget / pl.stride);
  grid = std::max<uint64_t>(1, std::min(grid, fit));
  if (pl.stride > budget) return fail(DTREE_ERR_CUDA, "one election needs %llu bytes of scratch, more than the device has free",
                                      (unsigned long long)pl.stride);
  pl.grid = (int)grid;
  // Three launches per wave instead of one persistent kernel once there is more than a machine-full of elections
  // (tuning key sim_phased: -1 automatic, 0 never, 1 whenever IRV runs): a wave is up to 12 elections per SM (a whole
  // number of rounds of resident blocks for each of the three kernels), one scratch slot each.
  if (allow_phased && (t->sim_phased > 0 || (t->sim_phased < 0 && n_elections > 2 * grid && block == 256))) {
    pl.phased = true;
    const uint64_t round = 2ull * t->sm_count;
    uint64_t wave = std::min<uint64_t>(12ull * t->sm_count, fit);
    if (wave > round) wave -= wave % round;
    pl.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(wave, n_elections));
  }
  return DTREE_OK;
}

// Launch shape and scratch of deferred_kernel: one warp per election. Each warp gets an arena sized
// from measured usage (pl.cap_*); a few shared arenas hold the proven worst case (pl.arena_*): a sampled
// ballot sits in at most one live node, and a node with only observed ballots is a tree node.
constexpr uint32_t kMaxArenas = 8;

int make_plan_deferred(dtree *t, uint32_t n_sample, uint32_t n_elections, bool refresh_memory, Plan &pl) {
  const HostParams &P = t->tree.P;
  pl.W = 1;
  pl.log_gamma = device_params(P).small_alpha != 0;
  pl.block = 128;
  if (t->block_threads == 64 || t->block_threads == 128 || t->block_threads == 256) pl.block = t->block_threads;
  int occ = deferred_blocks_per_sm(pl.block, pl.log_gamma);
  if (occ <= 0) return fail(DTREE_ERR_CUDA, "deferred kernel cannot be resident (%s)", cudaGetErrorString(cudaGetLastError()));
  if (t->blocks_per_sm > 0) occ = std::min(occ, t->blocks_per_sm);
  pl.blocks_per_sm = occ;
  // ... and there are never more live nodes than distinct prefixes
  double prefixes = 0;
  for (uint32_t l = 0; l < P.nc; ++l) prefixes += falling_factorial(P.nc, l);
  uint64_t worst = (uint64_t)n_sample + t->n_nodes + 64;
  if (prefixes + 64 < (double)worst) worst = (uint64_t)prefixes + 64;
  if (worst > 0x7FFFFFF0ull) {
    pl.n_arenas = 0;
    return DTREE_OK;  // caller falls back to the materialising kernel
  }
  pl.arena_items = (uint32_t)worst;
  pl.arena_queue = (uint32_t)worst;
  pl.arena_big = (uint32_t)std::min<uint64_t>((uint64_t)n_sample / (t->urn_max + 1) + 64, worst);
  pl.arena_stride = (deferred_scratch_bytes(pl.arena_items, pl.arena_queue, pl.arena_big) + 255) & ~255ull;
  // Device memory this handle may plan with: what is free plus what it already holds. cudaMemGetInfo costs 0.1 ms
  // (milliseconds while another process polls the driver), so the figure is kept between calls of the same shape.
  if (refresh_memory || t->mem_avail == 0) {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    t->mem_avail = free_b + t->d_scratch.n + t->d_arena.n;
  }
  const uint64_t avail = t->mem_avail;
  pl.n_arenas = (uint32_t)std::min<uint64_t>(kMaxArenas, (uint64_t)((double)avail * 0.25) / pl.arena_stride);
  pl.cap_entries = 0;
  pl.mem_budget = (uint64_t)((double)avail * 0.5);
  (void)n_elections;
  return DTREE_OK;
}

// Sizes the per-warp arenas from pl.cap_* and fits the grid into the memory budget.
void shape_deferred(dtree *t, uint32_t n_elections, Plan &pl) {
  const int wpb = pl.block / 32;
  pl.stride = (deferred_scratch_bytes(pl.cap_items, pl.cap_queue, pl.cap_big) + 255) & ~255ull;  // per warp
  pl.stride_block = pl.stride * wpb;
  uint64_t grid = std::min<uint64_t>((uint64_t)pl.blocks_per_sm * t->sm_count, ((uint64_t)n_elections + wpb - 1) / wpb);
  uint64_t fit = std::max<uint64_t>(1, pl.mem_budget / pl.stride_block);
  pl.grid = (int)std::max<uint64_t>(1, std::min(grid, fit));
}

void fill_args(dtree *t, const Plan &pl, SimArgs &a) {
  memset(&a, 0, sizeof a);
  a.P = device_params(t->tree.P);
  a.T.node_base = t->dev.node_base.p;
  a.T.node_total = t->dev.node_total.p;
  a.T.slot_count = t->dev.slot_count.p;
  a.T.slot_child = t->dev.slot_child.p;
  a.T.slot_cand = t->dev.slot_cand.p;
  a.T.n_nodes = (uint32_t)t->n_nodes;
  a.urn_max = t->urn_max;
  a.P.binv_max = (float)t->binv_max;
  a.tau_quarters = t->tau_quarters;
  a.scratch = t->d_scratch.p;
  a.scratch_stride = pl.stride;
  a.cap_items = pl.cap_items;
  a.cap_small = pl.cap_small;
  a.lane_width = pl.lane_width;
  a.lane_sync = t->lane_sync >= 0 ? (uint32_t)t->lane_sync : (pl.lane && pl.block >= 512 ? 2u : 0u);
  a.cap_entries = pl.cap_entries;
  a.cap_queue = pl.cap_queue;
  a.cap_big = pl.cap_big;
  a.work_counter = &t->d_res.p->counter;
  a.error_flag = &t->d_res.p->error;
  a.stats = t->d_res.p->stats;
}

cudaError_t launch(const Plan &pl, const SimArgs &a, cudaStream_t s) {
  LaunchCfg cfg{pl.block, pl.log_gamma, pl.grid, s};
  if (pl.lane) return launch_lane(a, cfg);
  if (pl.deferred) return launch_deferred(a, cfg);
  if (pl.phased) {
    cfg.grid = (int)a.n_elections;  // block b works on election b of the wave
    return pl.W == 1 ? launch_simulate_phased<1>(a, cfg) : pl.W == 2 ? launch_simulate_phased<2>(a, cfg)
                                                                   : launch_simulate_phased<3>(a, cfg);
  }
  return pl.W == 1 ? launch_simulate<1>(a, cfg) : pl.W == 2 ? launch_simulate<2>(a, cfg) : launch_simulate<3>(a, cfg);
}

int check_sampling_args(dtree *t, uint32_t n_ballots, const char *seed, size_t seed_len) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  if (!seed && seed_len) return fail(DTREE_ERR_ARG, "null seed");
  if (t->tree.P.vd && t->tree.P.max_depth == 0)
    return fail(DTREE_ERR_ARG, "vd requires max_depth >= 1 (the reference reads depthFactors out of bounds here)");
  // The kernels carry the prior parameters a0 * depthFactor[d] and sums of up to 33 Gamma variates of about that size
  // in float: beyond this bound they would overflow to inf (vd at 32 candidates reaches 32! = 2.6e35 times a0).
  {
    const HostParams &P = t->tree.P;
    double worst = P.a0;
    if (P.vd)
      for (double f : P.depth_factors) worst = std::max(worst, P.a0 * f);
    if (!(worst <= 1e36))
      return fail(DTREE_ERR_ARG, "a0 times the largest depth factor is %.3g; the device sampler supports at most 1e36", worst);
  }
  (void)n_ballots;
  return DTREE_OK;
}

void store_kernel_counters(dtree *t, const unsigned long long *st) {
  t->stats[kOutEntries] = st[kStatEntries];
  t->stats[kOutItems] = st[kStatItems];
  t->stats[kOutGammas] = st[kStatGammas];
  t->stats[kOutBinomials] = st[kStatBinomials];
  t->stats[kOutRereads] = st[kStatRereads];
  t->stats[kOutSlots] = st[kStatSlots];
  t->stats[kOutUrnItems] = st[kStatUrnItems];
}

// The device result block and its page-locked host copy.
int ensure_result_block(dtree *t) {
  CUDA_TRY(t->d_res.reserve(1));
  if (!t->h_res) CUDA_TRY(cudaHostAlloc((void **)&t->h_res, sizeof(ResultBlock), cudaHostAllocDefault));
  return DTREE_OK;
}

// The counters a call reports are its own (dtree_last_stats: "of the last sampling call").
void begin_call_stats(dtree *t) {
  for (int i = 0; i < kOutCount; ++i)
    if (i != kOutBuildLaunches && i != kOutScratchBytes) t->stats[i] = 0;
}

// Waits for the stream, fetches the result block with one copy and digests it: kernel-side failure, counters,
// arena usage (which grows the cached sizing if an election came close to its capacity).
int fetch_results(dtree *t, cudaStream_t s, bool copy_enqueued = false) {
  NvtxRange nv("dtree:read_back");
  HostTimer tm;
  if (!copy_enqueued) CUDA_TRY(cudaMemcpyAsync(t->h_res, t->d_res.p, sizeof(ResultBlock), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  t->stats[kOutD2H] += sizeof(ResultBlock);
  t->stats[kOutFinishUs] += tm.us();
  const ResultBlock &R = *t->h_res;
  if (R.error) return fail(DTREE_ERR_CUDA, "internal: per-election scratch overflowed (code %d)", R.error);
  store_kernel_counters(t, R.stats);
  if (t->stats[kOutMode] >= 1) {
    dtree::Sizing &z = t->sizing;
    if (z.valid) {
      z.pool = std::max(z.pool, R.usage[0]);
      z.queue = std::max(z.queue, R.usage[1]);
      z.big = std::max(z.big, R.usage[2]);
      z.lpool = std::max(z.lpool, R.usage[5]);
      z.lqueue = std::max(z.lqueue, R.usage[6]);
      z.lbig = std::max(z.lbig, R.usage[7]);
      if (t->stats[kOutElections] > 0 && t->stats[kOutItems] > 0)
        z.nodes_per_election = (double)t->stats[kOutItems] / (double)t->stats[kOutElections];
    }
    t->stats[kOutArenaRuns] = R.usage[3];
    t->stats[kOutLaneRetries] = R.usage[4];
    t->last_lane_retries = R.usage[4];
  }
  return DTREE_OK;
}

// One sample_posterior job on one device: prepared once, launched wave by wave (so that several devices can be
// driven from one host thread and the abort callback can be polled between waves), then read back.
struct Job {
  dtree *t = nullptr;
  cudaStream_t s = 0;
  Plan pl, pp;    // the main launch; the pilot / replay launch of deferred_kernel
  SimArgs a, ar;  // their arguments (ar: replay of the elections lane_kernel could not finish)
  uint32_t nc = 0;
  uint64_t first_election = 0, n_elections = 0, done = 0, pilot_done = 0, wave = 0, launches = 0;
  bool want_orders = false;
  bool replay_launch = true;  // lane mode: a deferred_kernel launch behind every wave replays what lane_kernel handed over
};

int job_prepare(Job &j, dtree *t, uint64_t first_election, uint32_t n_elections, uint32_t n_ballots, uint32_t n_winners,
                int replace, const char *seed, size_t seed_len, unsigned long long *d_wins, bool want_orders,
                cudaStream_t s, int sc_function = DTREE_SC_IRV) {
  j.t = t;
  j.s = s;
  j.first_election = first_election;
  j.n_elections = n_elections;
  j.want_orders = want_orders;
  int rc = check_sampling_args(t, n_ballots, seed, seed_len);
  if (rc) return rc;
  const uint32_t nc = j.nc = t->tree.P.nc;
  if ((uint64_t)n_ballots < t->n_observed)  // R_tree.cpp:183-186
    return fail(DTREE_ERR_STATE, "`nBallots` must be larger than the number of ballots observed to obtain the posterior.");
  if (n_winners < 1 || n_winners >= nc) return fail(DTREE_ERR_ARG, "n_winners must be in [1, n_candidates)");
  if (sc_function != DTREE_SC_IRV && sc_function != DTREE_SC_PLURALITY)
    return fail(DTREE_ERR_ARG, "sc_function must be DTREE_SC_IRV or DTREE_SC_PLURALITY (stv is not implemented, as in R/social_choice.R:94-96)");
  if (sc_function == DTREE_SC_PLURALITY && (n_winners != 1 || want_orders))
    return fail(DTREE_ERR_ARG, "plurality reports the candidate(s) with the most first preferences: n_winners must be 1 and there is no elimination order");
  if (n_elections == 0) return DTREE_OK;
  {
    NvtxRange nv("dtree:mirror");
    HostTimer tm;
    rc = ensure_mirror(t, s);
    t->stats[kOutMirrorUs] += tm.us();
    if (rc) return rc;
  }
  NvtxRange nv_plan("dtree:plan");
  HostTimer tm_plan;
  const bool use_obs = !replace;
  const uint32_t n_sample = replace ? n_ballots : n_ballots - (uint32_t)t->n_observed;
  Plan &pl = j.pl, &pp = j.pp;
  pl = Plan();
  if (sc_function == DTREE_SC_PLURALITY) {
    // Only the root is split: a thread per election (lane_kernel) in a token arena — at most n_candidates children
    // are parked and nothing ever waits, so no pilot, no replay.
    rc = ensure_result_block(t);
    if (rc) return rc;
    ResultBlock *R = t->d_res.p;
    CUDA_TRY(cudaMemsetAsync(&R->stats[0], 0, sizeof(ResultBlock) - offsetof(ResultBlock, stats), s));
    const HostParams &P = t->tree.P;
    pl.W = 1;
    pl.log_gamma = device_params(P).small_alpha != 0;
    pl.deferred = pl.lane = true;
    pl.cap_items = 64;
    pl.cap_queue = 8;
    pl.cap_big = 16;
    pl.stride = (lane_scratch_bytes(pl.cap_items, pl.cap_queue, pl.cap_big) + 255) & ~255ull;
    const int wps = lane_warps_per_sm(device_params(P), pl.log_gamma);
    if (wps <= 0) return fail(DTREE_ERR_CUDA, "lane kernel cannot be resident");
    {  // one block per SM of as many warps as the job needs, as for IRV (below); the token arenas always fit
      const LaneShape shape = lane_shape(((uint64_t)n_elections + 31) / 32, (uint32_t)t->sm_count, (uint32_t)wps, ~0ull, 0);
      pl.block = (int)(shape.warps_per_block * 32);
      pl.stride_block = pl.stride * pl.block;
      pl.grid = (int)shape.grid;
    }
    CUDA_TRY(t->d_scratch.reserve(pl.stride_block * pl.grid));
    CUDA_TRY(t->d_retry.reserve(8));
    SimArgs &a = j.a;
    fill_args(t, pl, a);
    a.n_obs = 0;
    a.use_obs = use_obs ? 1 : 0;
    a.seed = hash_seed(seed, seed_len);
    a.n_sample = n_sample;
    a.n_winners = 1;
    a.run_irv = 1;
    a.sc_function = DTREE_SC_PLURALITY;
    a.win_counts = d_wins;
    a.retry_list = t->d_retry.p + 4;
    a.retry_count = t->d_retry.p;
    a.retry_cap = 0;
    a.usage = nullptr;
    j.replay_launch = false;
    j.wave = std::max<uint64_t>(1, n_elections);
    j.done = j.pilot_done = 0;
    t->stats[kOutElections] = n_elections;
    t->stats[kOutScratchBytes] = pl.stride_block * pl.grid;
    t->stats[kOutMode] = 2;
    t->stats[kOutPlanUs] += tm_plan.us();
    return DTREE_OK;
  }
  pl.deferred = t->mode == 1 || t->mode == 2;
  bool want_lane = t->mode == 2;
  if (t->mode < 0) pl.deferred = true;  // the deferred kernels are the faster ones on every configuration measured
  const HostParams &P = t->tree.P;
  dtree::Sizing &z = t->sizing;
  // the arena sizing measured for this tree, these parameters and this kind of job is still good
  const bool hit = z.valid && z.version == t->version && z.n_sample == n_sample && z.urn_max == t->urn_max &&
                   z.min_depth == P.min_depth && z.max_depth == P.max_depth && z.a0 == P.a0 && z.vd == P.vd &&
                   z.use_obs == use_obs && z.short_cuts == (n_winners == 1 && !want_orders);
  if (pl.deferred) {
    rc = make_plan_deferred(t, n_sample, n_elections, !hit, pl);
    if (rc) return rc;
    if (pl.n_arenas == 0) pl.deferred = false;  // not even one worst-case arena fits: materialise instead
  }
  if (!pl.deferred) {
    rc = make_plan(t, n_sample, n_elections, use_obs, pl, true);
    if (rc) return rc;
  }
  rc = ensure_result_block(t);
  if (rc) return rc;
  ResultBlock *R = t->d_res.p;
  // error flag, counters and usage start at zero; the win counts are the caller's business (they are added to)
  CUDA_TRY(cudaMemsetAsync(&R->stats[0], 0, sizeof(ResultBlock) - offsetof(ResultBlock, stats), s));
  if (want_orders) CUDA_TRY(t->d_orders.reserve((size_t)n_elections * nc));

  auto fill_job = [&](const Plan &plan, SimArgs &a) {
    fill_args(t, plan, a);
    a.obs_key = t->dev.obs_key.p;
    a.obs_cnt = t->dev.obs_cnt.p;
    a.obs_first = t->dev.obs_first.p;
    a.obs_tally0 = t->dev.obs_tally0.p;
    a.n_obs = use_obs ? t->dev.n_obs : 0;
    a.use_obs = use_obs ? 1 : 0;
    a.seed = hash_seed(seed, seed_len);
    a.n_sample = n_sample;
    a.n_winners = n_winners;
    a.run_irv = 1;
    a.win_counts = d_wins;
  };

  uint32_t pilot_done = 0;
  bool replay = false;
  if (pl.deferred) {
    CUDA_TRY(t->d_arena.reserve(pl.arena_stride * pl.n_arenas));
    // Arena locks start free at every call: a launch that died must not leave one held.
    CUDA_TRY(t->d_locks.reserve(kMaxArenas));
    CUDA_TRY(cudaMemsetAsync(t->d_locks.p, 0, kMaxArenas * sizeof(unsigned int), s));
    if (!hit) {
      // Pilot: a few elections run directly in the worst-case arenas (they cannot overflow and their
      // results count); their high-water marks size everybody else's arena.
      NvtxRange nv("dtree:pilot");
      pp = pl;
      pp.cap_items = pl.arena_items;
      pp.cap_queue = pl.arena_queue;
      pp.cap_big = pl.arena_big;
      pp.stride = pl.arena_stride;
      const int wpb = pp.block / 32;
      pp.grid = (int)((pl.n_arenas + wpb - 1) / wpb);
      pilot_done = std::min<uint32_t>(n_elections, 4 * pl.n_arenas);
      SimArgs ap;
      fill_job(pp, ap);
      ap.scratch = t->d_arena.p;
      ap.scratch_stride = pl.arena_stride;
      ap.max_warps = pl.n_arenas;
      ap.n_arenas = 0;
      ap.usage = R->usage;
      ap.arena_runs = R->usage + 3;
      ap.first_election = first_election;
      ap.n_elections = pilot_done;
      ap.elim_orders = want_orders ? t->d_orders.p : nullptr;
      CUDA_TRY(cudaMemsetAsync(&R->counter, 0, sizeof(unsigned int), s));
      CUDA_TRY(launch(pp, ap, s));
      ++j.launches;
      CUDA_TRY(cudaMemcpyAsync(t->h_res, R, sizeof(ResultBlock), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      t->stats[kOutD2H] += sizeof(ResultBlock);
      const ResultBlock &H = *t->h_res;
      z.valid = true;
      z.version = t->version;
      z.n_sample = n_sample; z.urn_max = t->urn_max; z.min_depth = P.min_depth; z.max_depth = P.max_depth;
      z.a0 = P.a0; z.vd = P.vd; z.use_obs = use_obs; z.short_cuts = n_winners == 1 && !want_orders;
      z.pool = H.usage[0]; z.queue = H.usage[1]; z.big = H.usage[2];
      // lane_kernel's own elections report their usage later; until then the pilot's (which, splitting in batches,
      // holds more at a time than lane_kernel's heaviest-first order) stands in. Its heap also holds small-count
      // nodes that carry many observed ballots, which deferred_kernel keeps in its small-count queue.
      z.lpool = z.pool; z.lqueue = z.queue; z.lbig = z.big + z.queue;
      z.nodes_per_election = (double)H.stats[kStatItems] / std::max<uint32_t>(pilot_done, 1);
    }
    // Capacities get a 2x margin (+ a constant that matters for light elections only) over the largest election seen
    // and are only revised when an election has used more than 70 % of them, so that a new maximum does not
    // reallocate the scratch every time. An election that still outgrows its arena re-runs in a worst-case one.
    if (!hit || 10ull * z.pool > 7ull * z.cap_pool) z.cap_pool = (uint32_t)std::min<uint64_t>(pl.arena_items, 2ull * z.pool + 2048);
    if (!hit || 10ull * z.queue > 7ull * z.cap_queue) z.cap_queue = (uint32_t)std::min<uint64_t>(pl.arena_queue, 2ull * z.queue + 2048);
    if (!hit || 10ull * z.big > 7ull * z.cap_big) z.cap_big = (uint32_t)std::min<uint64_t>(pl.arena_big, 2ull * z.big + 512);
    pl.cap_items = std::min(z.cap_pool, pl.arena_items);
    pl.cap_queue = std::min(z.cap_queue, pl.arena_queue);
    pl.cap_big = std::min(z.cap_big, pl.arena_big);
    // One thread per election (lane_kernel): 32x more arenas in flight, so a tighter margin — twice the largest
    // election lane_kernel itself has finished, revised only when an election has used more than 75 % of it. An
    // election that still outgrows its arena, or needs more node splits than `lane_max_splits`, is replayed by
    // deferred_kernel; without that budget one rare heavy election would keep its thread, and with it the whole
    // launch, busy, and the arenas would have to be sized for it.
    if (!hit || 4ull * z.lpool > 3ull * z.lane_pool) z.lane_pool = (uint32_t)std::min<uint64_t>(pl.arena_items, 2ull * z.lpool + 64);
    if (!hit || 4ull * z.lqueue > 3ull * z.lane_queue) z.lane_queue = (uint32_t)std::min<uint64_t>(pl.arena_queue, 2ull * z.lqueue + 32);
    if (!hit || 4ull * z.lbig > 3ull * z.lane_big) z.lane_big = (uint32_t)std::min<uint64_t>(pl.arena_items, 2ull * z.lbig + 24);
    Plan lp = pl;
    lp.lane = true;
    lp.block = kLaneBlock;
    lp.cap_items = std::min(z.lane_pool, pl.arena_items);
    lp.cap_queue = std::min(z.lane_queue, pl.arena_queue);
    lp.cap_big = std::min(z.lane_big, pl.arena_items);
    if (t->arena_cap) {  // test hook: arenas far too small, so that the overflow / replay paths run
      pl.cap_items = std::min(pl.cap_items, t->arena_cap);
      pl.cap_queue = std::min(pl.cap_queue, t->arena_cap);
      pl.cap_big = std::min(pl.cap_big, t->arena_cap);
      lp.cap_items = std::min(lp.cap_items, t->arena_cap);
      lp.cap_queue = std::min(lp.cap_queue, t->arena_cap);
      lp.cap_big = std::min(lp.cap_big, t->arena_cap);
    }
    lp.stride = (lane_scratch_bytes(lp.cap_items, lp.cap_queue, lp.cap_big) + 255) & ~255ull;  // per election
    {
      int wps = lane_warps_per_sm(device_params(P), pl.log_gamma);  // resident warps per SM
      if (wps <= 0) return fail(DTREE_ERR_CUDA, "lane kernel cannot be resident");
      if (t->blocks_per_sm > 0) wps = std::min(wps, t->blocks_per_sm * (kLaneBlock / 32));
      // Elections per warp: 32 unless the tuning key lane_width says otherwise. Measured (tools/width_probe.py, C3): a
      // warp's batch takes ~0.4 ms however few of its lanes work (one election alone is a 0.25 ms chain of dependent
      // variates), so spreading a small job over more, narrower warps buys at most 5 % (12 500 elections: 0.436 ms at
      // 32 per warp, 0.412 at 8, 0.53 at 4) and costs a lot once the warps outnumber the schedulers' appetite.
      const uint64_t todo = (uint64_t)(n_elections - pilot_done);
      lp.lane_width = t->lane_width ? t->lane_width : 32u;
      const uint64_t batches = (todo + lp.lane_width - 1) / lp.lane_width;  // one per warp and claim
      // Blocks: ONE per SM, of as many warps as the job needs to run in a single pass (at most `wps`, and what the
      // memory budget allows); a job larger than that runs on full blocks whose warps claim batch after batch.
      const uint64_t fit_warps = std::max<uint64_t>(1, pl.mem_budget / (lp.stride * 32));
      const LaneShape shape = lane_shape(batches, (uint32_t)t->sm_count, (uint32_t)wps, fit_warps, t->lane_block_warps);
      const uint64_t wpb = shape.warps_per_block;
      const uint64_t fit = std::max<uint64_t>(1, fit_warps / wpb), resident = (uint64_t)t->sm_count * ((uint64_t)wps / wpb);
      lp.block = (int)(wpb * 32);
      lp.stride_block = lp.stride * lp.block;
      lp.grid = (int)shape.grid;
      if (t->mode < 0) {
        // Automatic choice: a thread per election pays off when there are enough elections to fill the machine
        // (a thread runs its election alone, start to finish) and memory lets most of the threads be resident.
        const uint64_t lane_threads = std::min(resident, fit) * lp.block;
        const uint64_t warp_elections = (uint64_t)pl.blocks_per_sm * t->sm_count * (pl.block / 32);
        // ... and the elections are light: a heavy election (a close race expands tens of thousands of nodes) is
        // better served by the 32 lanes and shared-memory tile of deferred_kernel than by one thread.
        // A launch of either kernel lasts about as long as its slowest warp: lane_kernel ~ one batch of 32 elections
        // per warp however few elections there are, deferred_kernel ~ one election per warp and wave. Measured on the
        // 10-candidate benchmark a lane batch on a partly filled machine takes as long as 3.5 waves of deferred_kernel
        // (0.51 ms vs 0.144 ms), so the thread-per-election kernel wins from about four waves' worth of elections up
        // (9 500), well before the machine is full (12 500 elections per GPU: 0.73 ms with a warp each).
        want_lane = lane_threads >= 16 * warp_elections && (uint64_t)(n_elections - pilot_done) >= 4 * warp_elections &&
                    z.nodes_per_election <= 5000.0;
      }
    }
    shape_deferred(t, std::min<uint32_t>(n_elections - pilot_done, 1u << 20), pl);
    if (want_lane) {
      // The replay launch: the elections lane_kernel gave up on, by deferred_kernel — a warp each in the arenas a
      // warp-per-election launch would use (the two launches follow each other on the stream, so those alias
      // lane_kernel's scratch), with the worst-case arenas behind them.
      pp = pl;
      pl = lp;
      replay = true;
      CUDA_TRY(t->d_retry.reserve((1u << 20) + 4));
    }
  }
  uint64_t scratch = pl.stride_block * pl.grid;
  if (replay) scratch = std::max<uint64_t>(scratch, pp.stride_block * pp.grid);
  CUDA_TRY(t->d_scratch.reserve(scratch));
  SimArgs &a = j.a, &ar = j.ar;
  fill_job(pl, a);
  auto deferred_fields = [&](SimArgs &x) {
    x.arena_scratch = t->d_arena.p;
    x.arena_stride = pl.arena_stride;
    x.n_arenas = pl.n_arenas;
    x.arena_cap_items = pl.arena_items;
    x.arena_cap_queue = pl.arena_queue;
    x.arena_cap_big = pl.arena_big;
    x.arena_locks = t->d_locks.p;
    x.usage = R->usage;
    x.arena_runs = R->usage + 3;
    x.max_warps = 0xFFFFFFFFu;
  };
  if (pl.deferred) deferred_fields(a);
  if (replay) {
    a.retry_list = t->d_retry.p + 4;
    a.retry_count = t->d_retry.p;
    a.retry_cap = 1u << 20;
    a.n_arenas = 0;
    // split budget of a thread: four times what an election typically needs (measured), at least 32
    if (t->lane_max_splits < 0) a.lane_max_splits = (uint32_t)std::min(4.0 * t->sizing.nodes_per_election + 32.0, 4.0e9);
    else a.lane_max_splits = (uint32_t)std::min<int64_t>(t->lane_max_splits, 0xFFFFFFFFll);
    fill_job(pp, ar);
    deferred_fields(ar);
    ar.election_list = t->d_retry.p + 4;
    ar.n_elections_ptr = t->d_retry.p;
  }
  // One launch per wave. Without an abort callback the whole range is one wave (lane mode: at most 2^20
  // elections per wave, the capacity of its replay list).
  j.wave = n_elections;
  if (t->should_abort) j.wave = std::min<uint64_t>(n_elections, (uint64_t)pl.grid * (pl.lane ? 1024 : pl.deferred ? 4096 : 64));
  if (pl.lane) j.wave = std::min<uint64_t>(j.wave, 1u << 20);
  if (pl.phased) j.wave = (uint64_t)pl.grid;  // one scratch slot per election of the wave
  j.wave = std::max<uint64_t>(j.wave, 1);
  j.done = j.pilot_done = pilot_done;
  t->stats[kOutElections] = n_elections;
  t->stats[kOutScratchBytes] = scratch;
  t->stats[kOutMode] = pl.lane ? 2 : pl.deferred ? 1 : 0;
  t->stats[kOutPlanUs] += tm_plan.us();
  return DTREE_OK;
}

// Enqueues the next wave of elections (and, in lane mode, the replay of those it could not finish).
int job_launch_wave(Job &j) {
  if (j.done >= j.n_elections) return DTREE_OK;
  NvtxRange nv(j.pl.lane ? "dtree:lane_kernel" : j.pl.deferred ? "dtree:deferred_kernel" : "dtree:simulate_kernel");
  HostTimer tm;
  dtree *t = j.t;
  ResultBlock *R = t->d_res.p;
  const uint32_t n = (uint32_t)std::min<uint64_t>(j.wave, j.n_elections - j.done);
  SimArgs &a = j.a;
  a.work_counter = &R->counter;
  a.error_flag = &R->error;
  a.stats = R->stats;
  a.first_election = j.first_election + j.done;
  a.n_elections = n;
  a.elim_orders = j.want_orders ? t->d_orders.p + (size_t)j.done * j.nc : nullptr;
  CUDA_TRY(cudaMemsetAsync(&R->counter, 0, sizeof(unsigned int), j.s));
  if (j.pl.lane) CUDA_TRY(cudaMemsetAsync(t->d_retry.p, 0, sizeof(unsigned int), j.s));
  CUDA_TRY(launch(j.pl, a, j.s));
  j.launches += j.pl.phased ? 3 : 1;
  if (j.pl.lane && j.replay_launch) {
    NvtxRange nv2("dtree:replay");
    SimArgs &ar = j.ar;
    ar.work_counter = &R->counter;
    ar.error_flag = &R->error;
    ar.stats = R->stats;
    ar.first_election = a.first_election;
    ar.n_elections = n;
    ar.elim_orders = a.elim_orders;
    CUDA_TRY(cudaMemsetAsync(&R->counter, 0, sizeof(unsigned int), j.s));
    Plan replay = j.pp;
    replay.deferred = true;
    replay.lane = false;
    // The replay's warps pull elections from the retry list until it is empty, so any grid is correct; it is sized
    // for four times what the previous call handed over (+ 64 elections), not for a full machine that usually finds
    // the list empty.
    replay.grid = (int)std::min<uint64_t>((uint64_t)replay.grid, (4ull * t->last_lane_retries + 64 + 3) / 4);
    CUDA_TRY(launch(replay, ar, j.s));
    ++j.launches;
  }
  j.done += n;
  t->stats[kOutLaunches] = j.launches;
  t->stats[kOutLaunchUs] += tm.us();
  return DTREE_OK;
}

// Common body of the sample_posterior entry points. d_wins must be device memory on t->device.
int run_posterior(dtree *t, uint64_t first_election, uint32_t n_elections, uint32_t n_ballots, uint32_t n_winners,
                  int replace, const char *seed, size_t seed_len, unsigned long long *d_wins, uint8_t *h_orders,
                  cudaStream_t s, bool sync_at_end, int sc_function = DTREE_SC_IRV) {
  Job j;
  int rc = job_prepare(j, t, first_election, n_elections, n_ballots, n_winners, replace, seed, seed_len, d_wins,
                       h_orders != nullptr, s, sc_function);
  if (rc || n_elections == 0) return rc;
  while (j.done < j.n_elections) {
    if (t->should_abort && j.done > j.pilot_done) {
      CUDA_TRY(cudaStreamSynchronize(s));
      if (t->should_abort(t->abort_user)) return fail(DTREE_ERR_ABORTED, "aborted by the host");
    }
    rc = job_launch_wave(j);
    if (rc) return rc;
  }
  if (!sync_at_end) return DTREE_OK;
  if (h_orders) {
    CUDA_TRY(cudaMemcpyAsync(h_orders, t->d_orders.p, (size_t)n_elections * j.nc, cudaMemcpyDeviceToHost, s));
    t->stats[kOutD2H] += (uint64_t)n_elections * j.nc;
  }
  return fetch_results(t, s);
}

// Expansion of ONE election without IRV; copies the sampled entries back.
int run_dump(dtree *t, uint64_t election, uint32_t n_sample, const char *seed, size_t seed_len, dtree_ballots *out) {
  const uint32_t nc = t->tree.P.nc;
  cudaStream_t s = 0;
  int rc = ensure_mirror(t, s);
  if (rc) return rc;
  Plan pl;
  rc = make_plan(t, n_sample, 1, false, pl);
  if (rc) return rc;
  pl.grid = 1;
  CUDA_TRY(t->d_scratch.reserve(pl.stride));
  rc = ensure_result_block(t);
  if (rc) return rc;
  CUDA_TRY(t->d_entry_count.reserve(1));
  CUDA_TRY(cudaMemsetAsync(t->d_res.p, 0, sizeof(ResultBlock), s));
  CUDA_TRY(cudaMemsetAsync(t->d_entry_count.p, 0, sizeof(uint32_t), s));
  SimArgs a;
  fill_args(t, pl, a);
  a.seed = hash_seed(seed, seed_len);
  a.first_election = election;
  a.n_elections = 1;
  a.n_sample = n_sample;
  a.n_winners = 1;
  a.run_irv = 0;
  a.win_counts = t->d_res.p->wins;
  a.entry_count_out = t->d_entry_count.p;
  CUDA_TRY(launch(pl, a, s));
  CUDA_TRY(cudaMemcpyAsync(t->h_res, t->d_res.p, sizeof(ResultBlock), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  uint32_t n = 0;
  if (t->h_res->error) return fail(DTREE_ERR_CUDA, "internal: per-election scratch overflowed");
  CUDA_TRY(cudaMemcpy(&n, t->d_entry_count.p, sizeof n, cudaMemcpyDeviceToHost));
  const int W = pl.W;
  std::vector<uint64_t> keys((size_t)n * W + 1);
  std::vector<uint32_t> cnt(n + 1);
  std::vector<uint8_t> len(n + 1);
  // scratch layout of block 0 (Scratch<W>::bind): the entries are records of W key words + count + length
  uint8_t *base = t->d_scratch.p;
  uint8_t *p_ent;
  if (W == 1) { Scratch<1> sc; sc.bind(base, pl.cap_items, pl.cap_entries, pl.cap_small); p_ent = (uint8_t *)sc.ent; }
  else if (W == 2) { Scratch<2> sc; sc.bind(base, pl.cap_items, pl.cap_entries, pl.cap_small); p_ent = (uint8_t *)sc.ent; }
  else { Scratch<3> sc; sc.bind(base, pl.cap_items, pl.cap_entries, pl.cap_small); p_ent = (uint8_t *)sc.ent; }
  const size_t rec = (size_t)W * 8 + 8;
  if (n) {
    std::vector<uint8_t> raw((size_t)n * rec);
    CUDA_TRY(cudaMemcpy(raw.data(), p_ent, raw.size(), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; ++i) {
      const uint8_t *r = raw.data() + (size_t)i * rec;
      memcpy(&keys[(size_t)i * W], r, (size_t)W * 8);
      uint32_t c, l;
      memcpy(&c, r + (size_t)W * 8, 4);
      memcpy(&l, r + (size_t)W * 8 + 4, 4);
      cnt[i] = c;
      len[i] = (uint8_t)l;
    }
  }
  t->stats[kOutD2H] += (uint64_t)n * rec;
  // The kernel appends entries in whatever order its warps finish, which varies from launch to launch; the list is
  // returned in a canonical order (by ballot, lexicographic in the candidate indices, a prefix before its
  // extensions) so that the same seed gives the same R object (tests/testthat/test-determinism.R:5-22).
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; ++i) order[i] = i;
  auto pref = [&](uint32_t i, uint32_t j) {
    return (uint32_t)((keys[(size_t)i * W + j / kPrefsPerWord] >> (kPrefBits * (j % kPrefsPerWord))) & 31u);
  };
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    const uint32_t m = std::min<uint32_t>(len[x], len[y]);
    for (uint32_t j = 0; j < m; ++j) {
      const uint32_t a = pref(x, j), b = pref(y, j);
      if (a != b) return a < b;
    }
    return len[x] < len[y];
  });
  for (uint32_t oi = 0; oi < n; ++oi) {
    const uint32_t i = order[oi];
    for (uint32_t j = 0; j < len[i]; ++j) {
      uint64_t w = keys[(size_t)i * W + j / kPrefsPerWord];
      out->prefs.push_back((uint8_t)((w >> (kPrefBits * (j % kPrefsPerWord))) & 31u));
    }
    out->offsets.push_back(out->prefs.size());
    out->counts.push_back(cnt[i]);
  }
  (void)nc;
  return DTREE_OK;
}

void append_observed(const FlatTree &F, dtree_ballots *out) {
  for (size_t i = 0; i < F.obs_counts.size(); ++i) {
    out->prefs.insert(out->prefs.end(), F.obs_prefs.begin() + F.obs_offsets[i], F.obs_prefs.begin() + F.obs_offsets[i + 1]);
    out->offsets.push_back(out->prefs.size());
    out->counts.push_back(F.obs_counts[i]);
  }
}

}  // namespace

extern "C" {

const char *dtree_last_error(void) { return g_error.c_str(); }
const char *dtree_version(void) { return "dtree_b200 0.1 (sm_100a)"; }

int dtree_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int dtree_create(uint32_t nc, uint32_t min_depth, uint32_t max_depth, double a0, int vd, int device, dtree_t **out) {
  if (!out) return fail(DTREE_ERR_ARG, "null output pointer");
  *out = nullptr;
  if (nc < 2 || nc > DTREE_MAX_CANDIDATES) return fail(DTREE_ERR_ARG, "n_candidates must be in [2, %d]", DTREE_MAX_CANDIDATES);
  if (!(a0 >= 0.0) || !std::isfinite(a0)) return fail(DTREE_ERR_ARG, "`a0` must be >= 0.");
  if (min_depth > max_depth || max_depth > nc)
    return fail(DTREE_ERR_ARG, "min_depth and max_depth must satisfy: 0 <= min_depth <= max_depth <= n_candidates");
  HostParams P;
  P.nc = nc;
  P.min_depth = min_depth;
  P.max_depth = max_depth;
  P.a0 = a0;
  P.vd = vd != 0;
  dtree *t = new dtree(P, device);
  // The device is bound lazily so that a tree can be built and inspected on a host without a GPU;
  // anything that samples resolves it (and fails loudly if there is none).
  *out = t;
  return DTREE_OK;
}

int dtree_destroy(dtree_t *t) {
  if (!t) return DTREE_OK;
  for (dtree *r : t->replicas) dtree_destroy(r);
  t->replicas.clear();
  unpin_log(t);
  if (t->sm_count) {
    DeviceGuard g(t->device);
    t->dev.release_all();
    t->d_scratch.release(); t->d_res.release();
    t->d_entry_count.release(); t->d_orders.release();
    t->d_arena.release(); t->d_locks.release(); t->d_retry.release();
    if (t->h_res) cudaFreeHost(t->h_res);
    t->h_res = nullptr;
  }
  delete t;
  return DTREE_OK;
}

int dtree_set_min_depth(dtree_t *t, uint32_t v) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  t->tree.P.min_depth = v;  // irv_node.h:127: depth factors deliberately NOT recomputed
  return DTREE_OK;
}
int dtree_set_max_depth(dtree_t *t, uint32_t v) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  if (v > t->tree.P.nc) return fail(DTREE_ERR_ARG, "max_depth must be <= n_candidates");
  t->tree.P.max_depth = v;
  t->tree.P.recompute_depth_factors();  // irv_node.h:134-137
  return DTREE_OK;
}
int dtree_set_a0(dtree_t *t, double v) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  if (!(v >= 0.0) || !std::isfinite(v)) return fail(DTREE_ERR_ARG, "`a0` must be >= 0.");
  t->tree.P.a0 = v;
  return DTREE_OK;
}
int dtree_set_vd(dtree_t *t, int v) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  t->tree.P.vd = v != 0;
  return DTREE_OK;
}
uint32_t dtree_n_candidates(const dtree_t *t) { return t->tree.P.nc; }
uint32_t dtree_min_depth(const dtree_t *t) { return t->tree.P.min_depth; }
uint32_t dtree_max_depth(const dtree_t *t) { return t->tree.P.max_depth; }
double dtree_a0(const dtree_t *t) { return t->tree.P.a0; }
int dtree_vd(const dtree_t *t) { return t->tree.P.vd ? 1 : 0; }
uint32_t dtree_depth_factors(const dtree_t *t, double *out, uint32_t cap) {
  const std::vector<double> &f = t->tree.P.depth_factors;
  for (uint32_t i = 0; i < f.size() && i < cap; ++i) out[i] = f[i];
  return (uint32_t)f.size();
}

int dtree_update(dtree_t *t, const uint8_t *prefs, const uint64_t *offsets, uint64_t n, uint64_t *n_short) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  if (n_short) *n_short = 0;
  if (n == 0) return DTREE_OK;
  if (!offsets || (!prefs && offsets[n] > 0)) return fail(DTREE_ERR_ARG, "null ballot buffers");
  if (t->n_observed + n > 0xFFFFFFFFull) return fail(DTREE_ERR_STATE, "more than 2^32-1 observed ballots");
  std::string err;
  uint64_t mask = 0;
  if (!t->tree.P.validate(prefs, offsets, n, n_short, &mask, err)) return fail(DTREE_ERR_ARG, "%s", err.c_str());
  // append to the log, offsets rebased onto it (a buffer that has to grow stops being page-locked first)
  const uint64_t at = t->log_prefs.size(), first = offsets[0];
  if (at + (offsets[n] - first) > t->log_prefs.capacity() || t->log_offsets.size() + n > t->log_offsets.capacity()) unpin_log(t);
  t->log_prefs.insert(t->log_prefs.end(), prefs + first, prefs + offsets[n]);
  for (uint64_t i = 1; i <= n; ++i) {
    t->log_offsets.push_back(at + (offsets[i] - first));
    const uint64_t len = offsets[i] - offsets[i - 1];
    if (t->log_len == kLenNone) t->log_len = (uint32_t)len;
    else if (t->log_len != (uint32_t)len) t->log_len = kLenMixed;
  }
  t->n_observed += n;  // R_tree.cpp:149
  t->depth_mask |= mask;
  ++t->version;
  return DTREE_OK;
}

int dtree_reset(dtree_t *t) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  t->log_prefs.clear();
  t->log_offsets.assign(1, 0);
  t->log_len = kLenNone;
  t->n_observed = t->depth_mask = 0;
  t->tree.reset();
  t->trie_ballots = 0;
  t->dev.log_ballots = t->dev.log_bytes = 0;
  ++t->version;
  return DTREE_OK;
}
uint64_t dtree_n_observed(const dtree_t *t) { return t->n_observed; }
uint64_t dtree_observed_depth_mask(const dtree_t *t) { return t->depth_mask; }
uint64_t dtree_n_nodes(const dtree_t *t) {
  sync_trie(const_cast<dtree_t *>(t));
  return t->tree.n_nodes();
}

int dtree_export_nodes(const dtree_t *t, uint32_t *buf, uint64_t cap, uint64_t *needed) {
  if (!t || !needed) return fail(DTREE_ERR_ARG, "null argument");
  FlatTree F;
  sync_trie(const_cast<dtree_t *>(t));
  t->tree.flatten(F);
  const uint32_t nc = t->tree.P.nc;
  std::vector<uint32_t> out;
  // prefixes by propagation in level order
  std::vector<std::vector<uint8_t>> prefix(F.n_nodes());
  for (uint32_t i = 0; i < F.n_nodes(); ++i) {
    uint32_t depth = F.node_depth[i], K = nc - depth, base = F.node_base[i];
    out.push_back(depth);
    for (uint8_t c : prefix[i]) out.push_back(c);
    out.push_back(K);
    for (uint32_t j = 0; j < K; ++j) out.push_back(F.slot_cand[base + j]);
    for (uint32_t j = 0; j <= K; ++j) out.push_back(F.slot_count[base + j]);
    for (uint32_t j = 0; j < K; ++j) {
      int32_t c = F.slot_child[base + j];
      out.push_back(c >= 0);
      if (c >= 0) {
        prefix[c] = prefix[i];
        prefix[c].push_back(F.slot_cand[base + j]);
      }
    }
    std::vector<uint8_t>().swap(prefix[i]);
  }
  *needed = out.size();
  if (buf && cap >= out.size()) std::copy(out.begin(), out.end(), buf);
  return DTREE_OK;
}

int dtree_observed(const dtree_t *t, dtree_ballots_t **out) {
  if (!t || !out) return fail(DTREE_ERR_ARG, "null argument");
  FlatTree F;
  sync_trie(const_cast<dtree_t *>(t));
  t->tree.flatten(F);
  dtree_ballots *b = new dtree_ballots;
  append_observed(F, b);
  *out = b;
  return DTREE_OK;
}

static int bind_device(dtree_t *t) {
  if (t->sm_count) return DTREE_OK;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(DTREE_ERR_CUDA, "no CUDA device available: libdtree_b200 has no CPU path");
  }
  if (t->device < 0) CUDA_TRY(cudaGetDevice(&t->device));
  if (t->device >= n) return fail(DTREE_ERR_ARG, "device %d does not exist (%d visible)", t->device, n);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, t->device));
  // sm_100a code is not forward compatible: sm_103 / sm_110 / sm_120 devices have no kernel image in this library
  if (prop.major != 10 || prop.minor != 0)
    return fail(DTREE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", t->device, prop.major, prop.minor);
  t->sm_count = prop.multiProcessorCount;
  t->smem_optin = prop.sharedMemPerBlockOptin;
  return DTREE_OK;
}

int dtree_sync_device(dtree_t *t) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  int rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  begin_call_stats(t);
  NvtxRange nv("dtree:mirror");
  return ensure_mirror(t, 0);
}

int dtree_invalidate_device(dtree_t *t) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  t->mirrored_version = ~0ull;
  t->dev.log_ballots = t->dev.log_bytes = 0;  // the next mirror starts from host memory again
  return DTREE_OK;
}

int dtree_sample_posterior(dtree_t *t, uint64_t first_election, uint32_t n_elections, uint32_t n_ballots,
                           uint32_t n_winners, int replace, const char *seed, size_t seed_len, uint64_t *win_counts,
                           uint8_t *elim_orders) {
  if (!t || !win_counts) return fail(DTREE_ERR_ARG, "null argument");
  NvtxRange nv("dtree_sample_posterior");
  int rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  const uint32_t nc = t->tree.P.nc;
  begin_call_stats(t);
  rc = ensure_result_block(t);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(t->d_res.p->wins, 0, sizeof t->d_res.p->wins, 0));
  for (uint32_t c = 0; c < nc; ++c) win_counts[c] = 0;
  rc = run_posterior(t, first_election, n_elections, n_ballots, n_winners, replace, seed, seed_len, t->d_res.p->wins,
                     elim_orders, 0, true);
  if (rc || n_elections == 0) return rc;
  for (uint32_t c = 0; c < nc; ++c) win_counts[c] = t->h_res->wins[c];  // came back with the result block
  return DTREE_OK;
}

int dtree_sample_posterior_sc(dtree_t *t, int sc_function, uint64_t first_election, uint32_t n_elections, uint32_t n_ballots,
                              int replace, const char *seed, size_t seed_len, uint64_t *win_counts) {
  if (!t || !win_counts) return fail(DTREE_ERR_ARG, "null argument");
  if (sc_function == DTREE_SC_IRV)
    return dtree_sample_posterior(t, first_election, n_elections, n_ballots, 1, replace, seed, seed_len, win_counts, nullptr);
  NvtxRange nv("dtree_sample_posterior_sc");
  int rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  const uint32_t nc = t->tree.P.nc;
  begin_call_stats(t);
  rc = ensure_result_block(t);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(t->d_res.p->wins, 0, sizeof t->d_res.p->wins, 0));
  for (uint32_t c = 0; c < nc; ++c) win_counts[c] = 0;
  rc = run_posterior(t, first_election, n_elections, n_ballots, 1, replace, seed, seed_len, t->d_res.p->wins, nullptr, 0,
                     true, sc_function);
  if (rc || n_elections == 0) return rc;
  for (uint32_t c = 0; c < nc; ++c) win_counts[c] = t->h_res->wins[c];
  return DTREE_OK;
}

int dtree_sample_posterior_device(dtree_t *t, uint64_t first_election, uint32_t n_elections, uint32_t n_ballots,
                                  uint32_t n_winners, int replace, const char *seed, size_t seed_len,
                                  uint64_t *d_win_counts, void *stream) {
  if (!t || !d_win_counts) return fail(DTREE_ERR_ARG, "null argument");
  int rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  begin_call_stats(t);
  return run_posterior(t, first_election, n_elections, n_ballots, n_winners, replace, seed, seed_len,
                       (unsigned long long *)d_win_counts, nullptr, (cudaStream_t)stream, false);
}

int dtree_sample_posterior_multi(dtree_t *t, const int *devices, int n_devices, uint64_t first_election,
                                 uint32_t n_elections, uint32_t n_ballots, uint32_t n_winners, int replace,
                                 const char *seed, size_t seed_len, uint64_t *win_counts) {
  if (!t || !win_counts) return fail(DTREE_ERR_ARG, "null argument");
  if (n_devices < 1 || n_devices > 64) return fail(DTREE_ERR_ARG, "n_devices must be in [1, 64]");
  NvtxRange nv("dtree_sample_posterior_multi");
  const int visible = dtree_device_count();
  if (visible < 1) return fail(DTREE_ERR_CUDA, "no CUDA device available: libdtree_b200 has no CPU path");
  // every ordinal is checked before anything is launched
  for (int i = 0; i < n_devices; ++i) {
    const int dev = devices ? devices[i] : i;
    if (dev < 0 || dev >= visible) return fail(DTREE_ERR_ARG, "device %d does not exist (%d visible)", dev, visible);
  }
  const uint32_t nc = t->tree.P.nc;
  for (uint32_t c = 0; c < nc; ++c) win_counts[c] = 0;
  begin_call_stats(t);
  // one replica per listed device; index i of `replicas` serves devices[i]
  while ((int)t->replicas.size() < n_devices) t->replicas.push_back(nullptr);
  const uint32_t per = (uint32_t)(((uint64_t)n_elections + n_devices - 1) / n_devices);  // ceil, as sharding.py
  std::vector<Job> jobs;  // the replicas with work, in device-list order
  jobs.reserve(n_devices);
  int rc = DTREE_OK;
  const dtree *sized = nullptr;  // a replica whose arena sizing (pilot) the later ones can take over
  for (int i = 0; i < n_devices && rc == DTREE_OK; ++i) {
    const int dev = devices ? devices[i] : i;
    dtree *&r = t->replicas[i];
    if (r && r->device != dev) {
      dtree_destroy(r);
      r = nullptr;
    }
    if (!r) r = new dtree(t->tree.P, dev);
    if (r->version != t->version || r->n_observed != t->n_observed) {  // same ballot log, own mirror
      unpin_log(r);
      r->log_prefs = t->log_prefs;
      r->log_offsets = t->log_offsets;
      r->n_observed = t->n_observed;
      r->depth_mask = t->depth_mask;
      r->version = t->version;
      r->tree.reset();
      r->trie_ballots = 0;
      r->dev.log_ballots = r->dev.log_bytes = 0;
    }
    r->tree.P = t->tree.P;
    r->host_build = t->host_build;
    r->build_mode = t->build_mode;
    r->sim_phased = t->sim_phased;
    r->lane_width = t->lane_width;
    r->lane_block_warps = t->lane_block_warps;
    r->lane_sync = t->lane_sync;
    r->mode = t->mode;
    r->urn_max = t->urn_max;
    r->block_threads = t->block_threads;
    r->blocks_per_sm = t->blocks_per_sm;
    r->arena_cap = t->arena_cap;
    r->tau_quarters = t->tau_quarters;
    r->binv_max = t->binv_max;
    r->lane_max_splits = t->lane_max_splits;
    // The abort callback is the parent's: it is polled between waves below, on the calling thread, once for all
    // devices. A replica only needs to know that its job must be cut into waves.
    r->should_abort = t->should_abort;
    r->abort_user = t->abort_user;
    const uint64_t lo = std::min<uint64_t>(n_elections, (uint64_t)i * per), hi = std::min<uint64_t>(n_elections, lo + per);
    if (hi == lo) continue;
    rc = bind_device(r);
    if (rc) break;
    DeviceGuard g(r->device);
    if (!g.ok) {
      rc = fail(DTREE_ERR_CUDA, "cannot select device %d", r->device);
      break;
    }
    begin_call_stats(r);
    rc = ensure_result_block(r);
    if (rc) break;
    if (cudaMemsetAsync(r->d_res.p->wins, 0, sizeof r->d_res.p->wins, 0) != cudaSuccess) {
      rc = fail(DTREE_ERR_CUDA, "clearing win counts on device %d failed", r->device);
      break;
    }
    // Same tree, same parameters, same job shape: the arena sizing measured by the first replica's pilot holds for
    // all of them, so only that one pilot runs (and the devices are not serialised behind one another's).
    if (sized && !(r->sizing.valid && r->sizing.version == r->version)) {
      const dtree::Sizing &z = sized->sizing;
      r->sizing = z;
    }
    jobs.emplace_back();
    rc = job_prepare(jobs.back(), r, first_election + lo, (uint32_t)(hi - lo), n_ballots, n_winners, replace, seed, seed_len,
                     r->d_res.p->wins, false, 0);
    if (rc) {
      jobs.pop_back();
      break;
    }
    if (!sized && r->sizing.valid) sized = r;
  }
  // waves, round-robin over the devices; the abort callback is polled once per round with every device idle
  bool more = rc == DTREE_OK;
  for (uint64_t round = 0; more; ++round) {
    if (t->should_abort && round > 0) {
      for (Job &j : jobs) {
        DeviceGuard g(j.t->device);
        if (cudaStreamSynchronize(j.s) != cudaSuccess) rc = fail(DTREE_ERR_CUDA, "device %d failed", j.t->device);
      }
      if (rc == DTREE_OK && t->should_abort(t->abort_user)) rc = fail(DTREE_ERR_ABORTED, "aborted by the host");
      if (rc) break;
    }
    more = false;
    for (Job &j : jobs) {
      if (j.done >= j.n_elections) continue;
      DeviceGuard g(j.t->device);
      rc = job_launch_wave(j);
      if (rc) break;
      more |= j.done < j.n_elections;
    }
    if (rc) break;
  }
  // collect — also on failure, so that nothing is left running when the call returns
  std::string first_error = rc ? g_error : std::string();
  uint64_t launches = 0;
  std::vector<char> enqueued(jobs.size(), 0);
  for (size_t k = 0; k < jobs.size(); ++k) {  // every device's read-back is in flight before the first one is waited for
    dtree *r = jobs[k].t;
    DeviceGuard g(r->device);
    enqueued[k] = cudaMemcpyAsync(r->h_res, r->d_res.p, sizeof(ResultBlock), cudaMemcpyDeviceToHost, jobs[k].s) == cudaSuccess;
    if (!enqueued[k]) cudaGetLastError();
  }
  for (size_t k = 0; k < jobs.size(); ++k) {
    Job &j = jobs[k];
    dtree *r = j.t;
    DeviceGuard g(r->device);
    int rc2 = fetch_results(r, j.s, enqueued[k] != 0);
    if (rc2 == DTREE_OK && rc == DTREE_OK)
      for (uint32_t c = 0; c < nc; ++c) win_counts[c] += r->h_res->wins[c];
    else if (rc == DTREE_OK) {
      rc = rc2;
      first_error = g_error;
    }
    launches += j.launches;
    for (int k : {kOutEntries, kOutItems, kOutGammas, kOutBinomials, kOutRereads, kOutSlots, kOutUrnItems, kOutArenaRuns,
                  kOutLaneRetries, kOutH2D, kOutD2H, kOutScratchBytes})
      t->stats[k] += r->stats[k];
    t->stats[kOutMode] = r->stats[kOutMode];
  }
  if (rc) g_error = first_error;
  t->stats[kOutLaunches] = launches;
  t->stats[kOutElections] = n_elections;
  return rc;
}

int dtree_sample_posterior_finish(dtree_t *t, void *stream) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  if (!t->sm_count) return DTREE_OK;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  if (!t->d_res.p || !t->h_res) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return DTREE_OK;
  }
  return fetch_results(t, (cudaStream_t)stream);
}

int dtree_sample_predictive(dtree_t *t, uint32_t n_ballots, const char *seed, size_t seed_len, dtree_ballots_t **out) {
  if (!t || !out) return fail(DTREE_ERR_ARG, "null argument");
  *out = nullptr;
  int rc = check_sampling_args(t, n_ballots, seed, seed_len);
  if (rc) return rc;
  rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  dtree_ballots *b = new dtree_ballots;
  rc = n_ballots ? run_dump(t, 0, n_ballots, seed, seed_len, b) : DTREE_OK;
  if (rc) {
    delete b;
    return rc;
  }
  *out = b;
  return DTREE_OK;
}

int dtree_posterior_set(dtree_t *t, uint64_t election, uint32_t n_ballots, int replace, const char *seed,
                        size_t seed_len, dtree_ballots_t **out) {
  if (!t || !out) return fail(DTREE_ERR_ARG, "null argument");
  *out = nullptr;
  int rc = check_sampling_args(t, n_ballots, seed, seed_len);
  if (rc) return rc;
  if (!replace && (uint64_t)n_ballots < t->n_observed)
    return fail(DTREE_ERR_STATE, "`nBallots` must be larger than the number of ballots observed to obtain the posterior.");
  rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  dtree_ballots *b = new dtree_ballots;
  uint32_t n_sample = replace ? n_ballots : n_ballots - (uint32_t)t->n_observed;
  rc = ensure_mirror(t, 0);
  if (!rc && !replace) rc = ensure_host_image(t);
  if (!rc && !replace) append_observed(t->flat, b);
  if (!rc && n_sample) rc = run_dump(t, election, n_sample, seed, seed_len, b);
  if (rc) {
    delete b;
    return rc;
  }
  *out = b;
  return DTREE_OK;
}

int dtree_irv(const uint8_t *prefs, const uint64_t *offsets, const uint32_t *counts, uint64_t n, uint32_t nc,
              const char *seed, size_t seed_len, int device, uint32_t *order_out, uint32_t *tie_rounds) {
  if (!order_out || (n && (!offsets || !counts))) return fail(DTREE_ERR_ARG, "null argument");
  if (nc < 2 || nc > DTREE_MAX_CANDIDATES) return fail(DTREE_ERR_ARG, "n_candidates must be in [2, %d]", DTREE_MAX_CANDIDATES);
  if (n > 0xFFFFFFF0ull) return fail(DTREE_ERR_ARG, "too many distinct ballots");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(DTREE_ERR_CUDA, "no CUDA device available: libdtree_b200 has no CPU path");
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  DeviceGuard g(device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", device);
  const int W = key_words_for(nc);
  std::vector<uint64_t> keys(n * W + 1);
  std::vector<uint8_t> lens(n + 1);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t len = offsets[i + 1] - offsets[i];
    if (len > nc) return fail(DTREE_ERR_ARG, "ballot ranks more candidates than exist");
    for (uint64_t j = offsets[i]; j < offsets[i + 1]; ++j)
      if (prefs[j] >= nc) return fail(DTREE_ERR_ARG, "Invalid candidate found during social-choice evaluation.");
    uint64_t k[kMaxKeyWords];
    uint32_t pl = 0;
    pack_ballot(prefs + offsets[i], (uint32_t)len, nc, false, k, &pl);
    for (int w = 0; w < W; ++w) keys[i * W + w] = k[w];
    lens[i] = (uint8_t)pl;
  }
  uint32_t cap = (uint32_t)std::max<uint64_t>(n, 1);
  uint64_t bytes = W == 1 ? Scratch<1>::bytes(0, cap, 0) : W == 2 ? Scratch<2>::bytes(0, cap, 0) : Scratch<3>::bytes(0, cap, 0);
  DevBuf<uint8_t> scratch, d_order;
  DevBuf<uint32_t> d_ties;
  auto cleanup = [&]() { scratch.release(); d_order.release(); d_ties.release(); };
  cudaError_t e = scratch.reserve(bytes);
  if (e == cudaSuccess) e = d_order.reserve(kMaxCandidates);
  if (e == cudaSuccess) e = d_ties.reserve(1);
  if (e == cudaSuccess) {
    uint8_t *p_ent;
    if (W == 1) { Scratch<1> sc; sc.bind(scratch.p, 0, cap); p_ent = (uint8_t *)sc.ent; }
    else if (W == 2) { Scratch<2> sc; sc.bind(scratch.p, 0, cap); p_ent = (uint8_t *)sc.ent; }
    else { Scratch<3> sc; sc.bind(scratch.p, 0, cap); p_ent = (uint8_t *)sc.ent; }
    if (n) {  // entry records: W key words, count, length (Entry<W>)
      const size_t rec = (size_t)W * 8 + 8;
      std::vector<uint8_t> raw((size_t)n * rec);
      for (uint64_t i = 0; i < n; ++i) {
        uint8_t *r = raw.data() + (size_t)i * rec;
        memcpy(r, &keys[i * W], (size_t)W * 8);
        const uint32_t c = counts[i], l = lens[i];
        memcpy(r + (size_t)W * 8, &c, 4);
        memcpy(r + (size_t)W * 8 + 4, &l, 4);
      }
      e = cudaMemcpy(p_ent, raw.data(), raw.size(), cudaMemcpyHostToDevice);
    }
  }
  if (e == cudaSuccess) {
    IrvArgs ia;
    ia.nc = nc;
    ia.n_entries = (uint32_t)n;
    ia.seed = hash_seed(seed, seed_len);
    ia.scratch = scratch.p;
    ia.cap_entries = cap;
    ia.order_out = d_order.p;
    ia.tie_rounds_out = d_ties.p;
    e = W == 1 ? launch_irv_only<1>(ia, 0) : W == 2 ? launch_irv_only<2>(ia, 0) : launch_irv_only<3>(ia, 0);
  }
  uint8_t order8[kMaxCandidates];
  uint32_t ties = 0;
  if (e == cudaSuccess) e = cudaMemcpy(order8, d_order.p, nc, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&ties, d_ties.p, 4, cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail(DTREE_ERR_CUDA, "dtree_irv: %s", cudaGetErrorString(e));
  for (uint32_t c = 0; c < nc; ++c) order_out[c] = order8[c];
  if (tie_rounds) *tie_rounds = ties;
  return DTREE_OK;
}

uint64_t dtree_ballots_size(const dtree_ballots_t *b) { return b->counts.size(); }
uint64_t dtree_ballots_total_prefs(const dtree_ballots_t *b) { return b->prefs.size(); }
int dtree_ballots_export(const dtree_ballots_t *b, uint8_t *prefs, uint64_t *offsets, uint32_t *counts) {
  if (!b || !offsets) return fail(DTREE_ERR_ARG, "null argument");
  if (prefs && !b->prefs.empty()) memcpy(prefs, b->prefs.data(), b->prefs.size());
  memcpy(offsets, b->offsets.data(), b->offsets.size() * sizeof(uint64_t));
  if (counts && !b->counts.empty()) memcpy(counts, b->counts.data(), b->counts.size() * sizeof(uint32_t));
  return DTREE_OK;
}
void dtree_ballots_free(dtree_ballots_t *b) { delete b; }

int dtree_set_abort_callback(dtree_t *t, int (*should_abort)(void *), void *user) {
  if (!t) return fail(DTREE_ERR_ARG, "null handle");
  t->should_abort = should_abort;
  t->abort_user = user;
  return DTREE_OK;
}

int dtree_set_tuning(dtree_t *t, const char *key, int64_t value) {
  if (!t || !key) return fail(DTREE_ERR_ARG, "null argument");
  if (!strcmp(key, "block_threads")) t->block_threads = (int)value;
  else if (!strcmp(key, "blocks_per_sm")) t->blocks_per_sm = (int)value;
  else if (!strcmp(key, "mode")) {
    if (value < -1 || value > 2)
      return fail(DTREE_ERR_ARG, "mode must be -1 (auto), 0 (materialise), 1 (deferred, warp per election) or 2 (deferred, thread per election)");
    t->mode = (int)value;
  } else if (!strcmp(key, "urn_max")) {
    if (value < 0 || value > 8) return fail(DTREE_ERR_ARG, "urn_max must be in [0, 8]");
    t->urn_max = (uint32_t)value;
  } else if (!strcmp(key, "binv_max")) {
    if (value < 10 || value > 60) return fail(DTREE_ERR_ARG, "binv_max must be in [10, 60] (BTRS needs n p >= 10; inversion is exact to 192 successes)");
    t->binv_max = (uint32_t)value;
  } else if (!strcmp(key, "tau_quarters")) {
    if (value < 1 || value > 3) return fail(DTREE_ERR_ARG, "tau_quarters must be 1, 2 or 3");
    t->tau_quarters = (uint32_t)value;
  } else if (!strcmp(key, "arena_cap")) {
    if (value != 0 && value < 8) return fail(DTREE_ERR_ARG, "arena_cap must be 0 or >= 8");
    if (value < 0 || value > 0x7FFFFFFF) return fail(DTREE_ERR_ARG, "arena_cap must be >= 0");
    t->arena_cap = (uint32_t)value;
  } else if (!strcmp(key, "lane_max_splits")) {
    if (value < -1) return fail(DTREE_ERR_ARG, "lane_max_splits must be -1 (auto), 0 (no limit) or a positive count");
    t->lane_max_splits = value;
  } else if (!strcmp(key, "lane_sync")) {
    if (value < -1 || value > 3) return fail(DTREE_ERR_ARG, "lane_sync must be -1 (auto) or 0..3");
    t->lane_sync = (int)value;
  } else if (!strcmp(key, "lane_block_warps")) {
    if (value < 0 || value > kLaneMaxBlock / 32) return fail(DTREE_ERR_ARG, "lane_block_warps must be 0 (auto) or 1..24");
    t->lane_block_warps = (uint32_t)value;
  } else if (!strcmp(key, "lane_width")) {
    if (value < 0 || value > 32) return fail(DTREE_ERR_ARG, "lane_width must be 0 (auto) or 1..32");
    t->lane_width = (uint32_t)value;
  } else if (!strcmp(key, "sim_phased")) {
    if (value < -1 || value > 1) return fail(DTREE_ERR_ARG, "sim_phased must be -1 (auto), 0 or 1");
    t->sim_phased = (int)value;
  } else if (!strcmp(key, "build_mode")) {
    if (value < 0 || value > 2) return fail(DTREE_ERR_ARG, "build_mode must be 0, 1 or 2");
    t->build_mode = (int)value;
    t->mirrored_version = ~0ull;
  } else if (!strcmp(key, "host_build")) {
    t->host_build = value != 0;
    t->mirrored_version = ~0ull;
  } else return fail(DTREE_ERR_ARG, "unknown tuning key '%s'", key);
  return DTREE_OK;
}

int dtree_debug_device_image(dtree_t *t, int which, void *out, uint64_t cap_bytes, uint64_t *needed_bytes) {
  if (!t || !needed_bytes) return fail(DTREE_ERR_ARG, "null argument");
  int rc = bind_device(t);
  if (rc) return rc;
  DeviceGuard g(t->device);
  if (!g.ok) return fail(DTREE_ERR_CUDA, "cannot select device %d", t->device);
  rc = ensure_mirror(t, 0);
  if (rc) return rc;
  const DeviceTreeStore &S = t->dev;
  const int W = key_words_for(t->tree.P.nc);
  const void *src = nullptr;
  uint64_t bytes = 0;
  switch (which) {
    case 0: src = S.node_base.p; bytes = ((uint64_t)S.n_nodes + 1) * 4; break;
    case 1: src = S.node_total.p; bytes = (uint64_t)S.n_nodes * 4; break;
    case 2: src = S.slot_count.p; bytes = (uint64_t)S.n_slots * 4; break;
    case 3: src = S.slot_child.p; bytes = (uint64_t)S.n_slots * 4; break;
    case 4: src = S.slot_cand.p; bytes = S.n_slots; break;
    case 5: src = S.obs_key.p; bytes = (uint64_t)S.n_obs * W * 8; break;
    case 6: src = S.obs_cnt.p; bytes = (uint64_t)S.n_obs * 4; break;
    case 7: src = S.obs_first.p; bytes = S.n_obs; break;
    case 8: src = S.obs_tally0.p; bytes = kMaxCandidates * 4; break;
    default: return fail(DTREE_ERR_ARG, "unknown device array %d", which);
  }
  *needed_bytes = bytes;
  if (out && cap_bytes >= bytes && bytes) CUDA_TRY(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
  return DTREE_OK;
}

int dtree_last_stats(const dtree_t *t, uint64_t *out, uint32_t cap) {
  if (!t || !out) return fail(DTREE_ERR_ARG, "null argument");
  for (uint32_t i = 0; i < cap && i < (uint32_t)kOutCount; ++i) out[i] = t->stats[i];
  return DTREE_OK;
}

}  // extern "C"
