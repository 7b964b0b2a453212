/* dtree_b200.h — C ABI of libdtree_b200.so, the B200 (sm_100a) implementation of
 * elections.dtree's posterior-simulation hot path.
 *
 * This is the drop-in boundary: the entry points below are what the
 * reference's Rcpp layer (src/R_tree.cpp, src/R_social_choice.cpp) would bind in place
 * of its DirichletTree<IRVNode,IRVBallot,IRVParameters> calls. Each declaration cites
 * the reference interface it replaces. INTEGRATION.md shows the Rcpp-side stub.
 *
 * Conventions
 *  - Plain C: pointers and sizes only. The library never throws, never calls R, never
 *    prints. Every function returns DTREE_OK (0) or an error code; the message is
 *    available from dtree_last_error() (thread-local). The host layer turns a non-zero
 *    status into Rcpp::stop(), mirroring R_tree.cpp:30-31,88-89,109-110,183-186.
 *  - Candidates are 0-based indices < n_candidates (<= 32). Name<->index mapping and
 *    argument validation stay in the host layer (R_tree.cpp:13-42,44-60).
 *  - Ballots cross the boundary as CSR: prefs[offsets[i] .. offsets[i+1]) are ballot i's
 *    candidates in preference order; an empty range is a blank ballot.
 *  - All input buffers are caller-owned HOST memory, copied during the call. Output
 *    buffers are caller-allocated unless a dtree_ballots_t is returned (free it with
 *    dtree_ballots_free).
 *  - There is no CPU fallback: a call that needs the GPU fails with DTREE_ERR_CUDA when
 *    no sm_100 device is usable.
 *  - A handle is not re-entrant (neither is the reference: R_tree.cpp:188 mutates the
 *    tree engine); distinct handles may be used from distinct threads.
 */
#ifndef DTREE_B200_H
#define DTREE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTREE_OK 0
#define DTREE_ERR_ARG 1     /* invalid argument (message says which) */
#define DTREE_ERR_CUDA 2    /* CUDA runtime failure / no device */
#define DTREE_ERR_STATE 3   /* e.g. n_ballots < number of observed ballots */
#define DTREE_ERR_ABORTED 4 /* the abort callback asked to stop */
#define DTREE_MAX_CANDIDATES 32

typedef struct dtree dtree_t;                 /* replaces RDirichletTree's `tree` member, R_tree.h:35 */
typedef struct dtree_ballots dtree_ballots_t; /* replaces std::list<IRVBallotCount>, irv_ballot.h:88 */

const char *dtree_last_error(void);
const char *dtree_version(void);
/* Number of usable CUDA devices (0 when none; never fails). */
int dtree_device_count(void);

/* ---- lifetime: RDirichletTree ctor/dtor, R_tree.cpp:44-66 ------------------------------------
 * device: CUDA ordinal the handle's device mirror lives on, or -1 for the current device. */
int dtree_create(uint32_t n_candidates, uint32_t min_depth, uint32_t max_depth, double a0, int vd,
                 int device, dtree_t **out);
int dtree_destroy(dtree_t *t);

/* ---- parameters: RcppModules.cpp:21-28 properties, irv_node.h:93-153 -------------------------
 * Setters take effect on the next sample and never touch stored counts. The range checks
 * (min <= max etc., R_tree.cpp:87-112) are the host layer's; these only store. set_min_depth
 * does NOT recompute the vd depth factors, set_max_depth does (irv_node.h:127,134-137). */
int dtree_set_min_depth(dtree_t *t, uint32_t v);
int dtree_set_max_depth(dtree_t *t, uint32_t v);
int dtree_set_a0(dtree_t *t, double v);
int dtree_set_vd(dtree_t *t, int v);
uint32_t dtree_n_candidates(const dtree_t *t);
uint32_t dtree_min_depth(const dtree_t *t);
uint32_t dtree_max_depth(const dtree_t *t);
double dtree_a0(const dtree_t *t);
int dtree_vd(const dtree_t *t);
/* irv_node.cpp:13-25; writes min(cap, max_depth) factors, returns max_depth. */
uint32_t dtree_depth_factors(const dtree_t *t, double *out, uint32_t cap);

/* ---- observations ----------------------------------------------------------------------------
 * dtree_update: RDirichletTree::update, R_tree.cpp:125-153 -> DirichletTree::update,
 * dirichlet_tree.h:182-193 -> IRVNode::update, irv_node.cpp:176-221. Every ballot counts once
 * (R_tree.cpp:38). Fails with DTREE_ERR_ARG, changing nothing, if a ballot repeats a candidate
 * or names one >= n_candidates. *n_short (may be NULL) receives how many ballots had
 * 0 < length < min_depth, for the host layer's warning (R_tree.cpp:139-147).
 * The call validates the ballots and appends them to the handle's ballot log; the count tree itself is
 * built on the device, from that log, by the next sampling call or dtree_sync_device (tree_build.cu:
 * one pass per tree level instead of one pointer walk + std::map insert per ballot). */
int dtree_update(dtree_t *t, const uint8_t *prefs, const uint64_t *offsets, uint64_t n_ballots,
                 uint64_t *n_short);
/* RDirichletTree::reset, R_tree.cpp:119-123 */
int dtree_reset(dtree_t *t);
uint64_t dtree_n_observed(const dtree_t *t);
/* Bit mask of observed ballot lengths (bit L set <=> a ballot of length L was observed):
 * RDirichletTree::observedDepths, R_tree.h:48, used by the warning of R_tree.cpp:95-105. */
uint64_t dtree_observed_depth_mask(const dtree_t *t);
uint64_t dtree_n_nodes(const dtree_t *t);

/* Parity hook for the node counts (the reference has no accessor; the oracle reads
 * IRVNode::as / ::children directly). Per node, as uint32 words, nodes in level order:
 *   depth, prefix[depth], K, slot_candidate[K], count[K+1], has_child[K]
 * Writes up to cap words, stores the number of words needed in *needed. */
int dtree_export_nodes(const dtree_t *t, uint32_t *buf, uint64_t cap, uint64_t *needed);
/* The observed multiset as the device sees it (derived from the node counts): one entry per
 * distinct observed ballot, except that a ballot ranking all n candidates and the same ballot
 * without its last preference are one entry of length n-1 (they are indistinguishable to
 * IRVNode::update, irv_node.cpp:210, and to IRV). */
int dtree_observed(const dtree_t *t, dtree_ballots_t **out);

/* ---- sampling --------------------------------------------------------------------------------
 * dtree_sample_posterior: RDirichletTree::samplePosterior, R_tree.cpp:177-257. Simulates
 * elections with GLOBAL ids [first_election, first_election + n_elections): each draws
 * posteriorSet(n_ballots, replace) (dirichlet_tree.h:216-235), runs IRV (irv_ballot.cpp:42-138)
 * and adds 1 to win_counts[c] for each of the n_winners candidates left standing
 * (R_tree.cpp:245-253). Draws are a pure function of (seed, global election id), so splitting a
 * job into ranges — over GPUs, ranks or calls — and summing win_counts gives bit-identical
 * totals. The caller divides by the total number of elections (R_tree.cpp:255).
 *   win_counts  [n_candidates] host, overwritten.
 *   elim_orders NULL, or [n_elections * n_candidates] host: full elimination order of each
 *               election, winner last (what socialChoiceIRV returns).
 * Fails with DTREE_ERR_STATE if n_ballots < n_observed (R_tree.cpp:183-186; like the reference,
 * also when replace is set). The reference's n_threads has no counterpart. */
int dtree_sample_posterior(dtree_t *t, uint64_t first_election, uint32_t n_elections,
                           uint32_t n_ballots, uint32_t n_winners, int replace, const char *seed,
                           size_t seed_len, uint64_t *win_counts, uint8_t *elim_orders);

/* The same posterior simulation under another social choice function (R/social_choice.R:35-97 offers "plurality",
 * "irv" and an unimplemented "stv"; the reference's sample_posterior is IRV only, R_tree.cpp:226-227):
 *   DTREE_SC_IRV        what dtree_sample_posterior does with n_winners = 1;
 *   DTREE_SC_PLURALITY  R/social_choice.R:58-83 on every simulated election: first preferences only, blank ballots
 *                       dropped, win_counts[c] += 1 for EVERY candidate with the maximal tally (ties give several
 *                       winners, as `which(frequencies == max(frequencies))` does; an election without a single first
 *                       preference has none). Only the root of the tree is ever sampled. */
#define DTREE_SC_IRV 0
#define DTREE_SC_PLURALITY 1
int dtree_sample_posterior_sc(dtree_t *t, int sc_function, uint64_t first_election, uint32_t n_elections,
                              uint32_t n_ballots, int replace, const char *seed, size_t seed_len, uint64_t *win_counts);

/* Same, for callers that keep the result on the device (one rank per GPU reducing with NCCL):
 * d_win_counts is DEVICE memory [n_candidates] uint64 on the handle's device and is ADDED to;
 * work is enqueued on `stream` (a cudaStream_t, NULL = legacy default stream) and the call
 * returns without synchronising. */
int dtree_sample_posterior_device(dtree_t *t, uint64_t first_election, uint32_t n_elections,
                                  uint32_t n_ballots, uint32_t n_winners, int replace,
                                  const char *seed, size_t seed_len, uint64_t *d_win_counts,
                                  void *stream);

/* One process, several GPUs (what an R session on a multi-GPU host uses; the thread pool of
 * R_tree.cpp:200-242 becomes a device pool). Elections [first_election, first_election + n_elections) are
 * split into n_devices contiguous ranges; range i runs on CUDA device devices[i] (devices == NULL: ordinals
 * 0..n_devices-1; a device may be listed more than once) against a replica of the tree mirrored there, all
 * devices concurrently; the uint64 win counts are added on the host. Totals are bit-identical to a
 * single-device call. win_counts [n_candidates] host, overwritten. */
int dtree_sample_posterior_multi(dtree_t *t, const int *devices, int n_devices, uint64_t first_election,
                                 uint32_t n_elections, uint32_t n_ballots, uint32_t n_winners, int replace,
                                 const char *seed, size_t seed_len, uint64_t *win_counts);

/* Waits for work enqueued by dtree_sample_posterior_device on `stream`, then reports a kernel-side
 * failure (DTREE_ERR_CUDA) if there was one and refreshes dtree_last_stats. */
int dtree_sample_posterior_finish(dtree_t *t, void *stream);

/* Builds the device mirror of the tree now (otherwise done lazily by the first sample after an
 * update/reset): uploads the part of the ballot log the device has not seen and rebuilds the level-ordered
 * arrays there. dtree_invalidate_device drops mirror and device log, so that the next sample starts from
 * host memory again. */
int dtree_sync_device(dtree_t *t);
int dtree_invalidate_device(dtree_t *t);

/* RDirichletTree::samplePredictive, R_tree.cpp:155-175, WITHOUT the count expansion: returns the
 * aggregated (ballot, count) list of one tree->sample(n_ballots) (dirichlet_tree.h:195-209). */
int dtree_sample_predictive(dtree_t *t, uint32_t n_ballots, const char *seed, size_t seed_len,
                            dtree_ballots_t **out);
/* The complete ballot set of ONE simulated election of dtree_sample_posterior (same seed, same
 * global id => same draws): observed entries first, then the sampled ones. Test/diagnostic hook. */
int dtree_posterior_set(dtree_t *t, uint64_t election, uint32_t n_ballots, int replace,
                        const char *seed, size_t seed_len, dtree_ballots_t **out);

/* ---- social choice: social_choice_irv, R_social_choice.cpp:16-95 -> socialChoiceIRV,
 * irv_ballot.cpp:42-138. order_out[n_candidates] = elimination order, winner last.
 * Ties for the minimal tally are broken uniformly at random from (seed, round); *tie_rounds (may be
 * NULL) = number of rounds with >= 2 standing candidates in which a tie was broken. */
int dtree_irv(const uint8_t *prefs, const uint64_t *offsets, const uint32_t *counts, uint64_t n,
              uint32_t n_candidates, const char *seed, size_t seed_len, int device,
              uint32_t *order_out, uint32_t *tie_rounds);

/* ---- (ballot, count) lists ------------------------------------------------------------------- */
uint64_t dtree_ballots_size(const dtree_ballots_t *b);
uint64_t dtree_ballots_total_prefs(const dtree_ballots_t *b);
/* prefs[total_prefs], offsets[size+1], counts[size] */
int dtree_ballots_export(const dtree_ballots_t *b, uint8_t *prefs, uint64_t *offsets, uint32_t *counts);
void dtree_ballots_free(dtree_ballots_t *b);

/* ---- cooperation with the host: RcppThread::checkUserInterrupt(), R_tree.cpp:222 ------------
 * should_abort(user) is polled on the calling thread between waves of elections; non-zero
 * makes the running call return DTREE_ERR_ABORTED. */
int dtree_set_abort_callback(dtree_t *t, int (*should_abort)(void *), void *user);

/* ---- tuning / introspection (not part of the reference interface) ----------------------------
 * key: "mode" — how dtree_sample_posterior* simulates an election: 1 and 2 expand a tree node only when IRV
 *        needs to look below it, with a warp (1) or a single thread (2) per election; 0 materialises every
 *        sampled ballot first, as the reference does (dirichlet_tree.h:216-235 then irv_ballot.cpp:42-138);
 *        -1 (default) picks 2 when there are enough elections to fill the GPU with one per thread and an election
 *        needs at most a few thousand node splits (measured on the first elections of the call), else 1, and 0 only
 *        if the worst-case scratch of 1/2 does not fit the device. All give bit-identical outcomes; forcing 2 on
 *        elections that need tens of thousands of splits (very close races, 20 candidates) is correct but slow;
 *      "block_threads" (0 = auto), "blocks_per_sm" (0 = auto), "urn_max" (largest count split by the
 *        Polya-urn shortcut instead of Gamma variates, <= 8);
 *      "host_build" (0 default; 1 = build the tree with the host trie and upload it, the pre-device-build
 *        path, kept to measure against; the two mirrors are identical array for array);
 *      "build_mode" (0 default: the device tree build is one cooperative kernel; 1: one launch per phase and level;
 *        2: as 1 with every level's size read back by the host to allocate exactly, the path taken on its own by trees
 *        whose a priori bounds exceed 1 GB; all three produce the same arrays);
 *      "tau_quarters" (1..3, default 3: share of the missing margin mode 1 lets wait when it picks the nodes to split);
 *      "binv_max" (10..60, default 30: binomials with n p below it are drawn by inversion, the others by BTRS);
 *      "arena_cap" (0 default; > 0 = modes 1 and 2 give every election a scratch arena of at most that many node
 *        records, far too small on purpose, so that tests reach the overflow paths: re-run in a shared worst-case
 *        arena (mode 1) and hand-over to the replay launch (mode 2); results do not change);
 *      "lane_max_splits" (-1 default = chosen from the node splits an election typically needs; 0 = no limit;
 *        n > 0 = in mode 2 an election that needs more than n node splits is handed to the replay launch, where a
 *        whole warp finishes it, instead of keeping its thread — and the launch — busy; results do not change);
 *      "sim_phased" (-1 default: mode 0 runs a job of more than two machine-fulls of elections as waves of three
 *        launches — level loop, small-subtree walkers, IRV stage, each at its own occupancy — instead of one
 *        persistent kernel; 0 never; 1 whenever an election is followed by IRV; results do not change);
 *      "lane_width" (0 default = 32; 1..32 = elections a warp of mode 2 claims at a time, the other lanes idle: a
 *        measurement hook for small jobs; results do not change);
 *      "lane_block_warps" (0 default = as many warps per block as make the job one block per SM, at most 24; 1..24 =
 *        warps per block of mode 2; results do not change);
 *      "lane_sync" (-1 default: 2 for blocks of 16 warps and more, else 0; 0..3 = how often the warps of a mode-2 block
 *        wait for each other: never / when they claim elections and after the root split / also at every IRV round /
 *        also at every node split. Warps that stay together share the instruction cache; results do not change).
 * Returns DTREE_ERR_ARG for an unknown key. */
int dtree_set_tuning(dtree_t *t, const char *key, int64_t value);
/* Statistics of the last sampling call on this handle (counters restart at every call; [7] and [14] describe the
 * scratch and tree build in force):
 *  [0] kernel launches, [1] elections simulated, [2] sampled entries written (mode 0) / 32-byte node records
 *  written to the arenas (modes 1, 2), [3] frontier items
 *  processed, [4] Gamma variates drawn, [5] binomial variates drawn, [6] IRV entry re-reads (mode 0) / node
 *  records read back (modes 1, 2),
 *  [7] device bytes of scratch, [8] H2D bytes, [9] D2H bytes, [10] outcome slots read,
 *  [11] nodes split by the urn shortcut, [12] mode used, [13] elections re-run in a worst-case arena,
 *  [14] kernel launches of the last device tree build, [15] elections mode 2 handed to its replay launch,
 *  [16..19] host microseconds spent in the call: device mirror (upload + tree build), planning, enqueueing the
 *  launches, waiting for and reading back the result. */
int dtree_last_stats(const dtree_t *t, uint64_t *out, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DTREE_B200_H */
