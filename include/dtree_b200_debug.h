/* dtree_b200_debug.h — test hooks exported by libdtree_b200.so next to the product ABI
 * (include/dtree_b200.h). They expose the device random-variate generators on their own so that
 * tests/ can check them distributionally (the reference's counterparts are libstdc++'s
 * std::gamma_distribution / std::binomial_distribution, called from distributions.cpp:45-46,63-64).
 * Not part of the drop-in boundary. */
#ifndef DTREE_B200_DEBUG_H
#define DTREE_B200_DEBUG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 of one counter block, evaluated on the host (on_device = 0) or by a one-thread
 * kernel (on_device = 1): the two must agree and match the published known-answer vectors. */
int dtree_debug_philox(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4], int on_device);
/* 64-bit seed derived from a seed string (what every sampling call does with `seed`). */
uint64_t dtree_debug_hash_seed(const char *seed, size_t seed_len);
/* n independent Gamma(alpha, 1) variates (log_space = 0) or their logarithms (log_space = 1). */
int dtree_debug_gamma(double alpha, int log_space, uint64_t seed, uint32_t n, float *out, int device);
/* n independent Binomial(n_trials, p) variates, 0 <= p <= 1. */
int dtree_debug_binomial(uint32_t n_trials, double p, uint64_t seed, uint32_t n, uint32_t *out, int device);
/* The same draws with every accept/reject decision that is not a quick accept made by the double-precision test
 * (exact_only = 1) instead of the guard-banded f32 test with f64 fall-back (exact_only = 0, what the kernels run):
 * the outputs must be identical, which is how the guard bands are validated (csrc/variates.cuh). */
int dtree_debug_gamma_ex(double alpha, int log_space, uint64_t seed, uint32_t n, float *out, int device, int exact_only);
int dtree_debug_binomial_ex(uint32_t n_trials, double p, uint64_t seed, uint32_t n, uint32_t *out, int device,
                            int exact_only);
/* BINV (the sampler for n_trials * p < 30) on EVERY one of its 2^23 possible uniforms: counts[k] = how many of them
 * give k (k <= 192). Divided by 2^23 this is the sampler's exact output law, to compare with the binomial pmf. */
int dtree_debug_binv_law(uint32_t n_trials, double p, uint32_t counts[256], int device);

/* Launch shape of a thread-per-election job (csrc/launch_shape.h; host arithmetic, no device needed): `batches` units of
 * 32 elections on `sms` multiprocessors that hold `warps_per_sm` warps of the kernel each, scratch for `fit_warps`
 * warps, tuning key lane_block_warps = forced_wpb (0: automatic). */
int dtree_debug_lane_shape(uint64_t batches, uint32_t sms, uint32_t warps_per_sm, uint64_t fit_warps, uint32_t forced_wpb,
                           uint32_t *warps_per_block, uint32_t *grid);

/* Copies one array of the handle's device mirror back to the host (building the mirror first if it is
 * stale). which: 0 node_base u32[n_nodes+1], 1 node_total u32[n_nodes], 2 slot_count u32[n_slots],
 * 3 slot_child i32[n_slots], 4 slot_cand u8[n_slots], 5 obs_key u64[n_obs*W], 6 obs_cnt u32[n_obs],
 * 7 obs_first u8[n_obs], 8 obs_tally0 u32[32]. *needed_bytes receives the array size; the copy happens
 * only if cap_bytes is large enough. Used to check the device-built tree against the host trie. */
struct dtree;
int dtree_debug_device_image(struct dtree *t, int which, void *out, uint64_t cap_bytes, uint64_t *needed_bytes);

#ifdef __cplusplus
}
#endif
#endif
