/* oracle_api.h — C interface shared by the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load these libraries, and only as the checker / CPU baseline.
 *
 * Two shared objects export exactly this interface:
 *   oracle/_ref/libdtree_ref.so  — the UNMODIFIED reference C++ core
 *        (/root/reference/src/{distributions,irv_ballot,irv_node}.cpp + headers)
 *        compiled where it lies, driven by oracle/ref_harness.cpp;
 *   oracle/libdtree_oracle.so    — oracle/dtree_oracle.cpp, a flat restatement of
 *        the same algorithm (no std::list, no pointer tree) that consumes the
 *        std::mt19937 stream in the same order, so with the same libstdc++ its
 *        outputs are bit-identical to the reference's (tests/test_oracle_pin.py).
 *
 * Ballots cross this interface as CSR: prefs[offsets[i] .. offsets[i+1]) are the
 * candidate indices of ballot i in preference order; counts[i] its multiplicity.
 */
#ifndef DTREE_ORACLE_API_H
#define DTREE_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_tree orc_tree;       /* one Dirichlet-tree (reference: RDirichletTree minus Rcpp) */
typedef struct orc_ballots orc_ballots; /* a list of (ballot, count) pairs */

const char *orc_kind(void); /* "reference" or "port" */

/* R_tree.cpp:44-60 (ctor), :63-66 (dtor) */
orc_tree *orc_create(uint32_t n_candidates, uint32_t min_depth, uint32_t max_depth, double a0,
                     int vd, const char *seed, size_t seed_len);
void orc_destroy(orc_tree *t);

/* irv_node.h:127-153 (setters; set_min_depth leaves the depth factors stale) */
void orc_set_min_depth(orc_tree *t, uint32_t v);
void orc_set_max_depth(orc_tree *t, uint32_t v);
void orc_set_a0(orc_tree *t, double v);
void orc_set_vd(orc_tree *t, int v);
/* irv_node.cpp:13-25; returns the number of factors (= max_depth) */
uint32_t orc_depth_factors(orc_tree *t, double *out, uint32_t cap);

/* dirichlet_tree.h:172-180 */
void orc_reset(orc_tree *t);
/* R_tree.cpp:125-153 -> dirichlet_tree.h:182-193 -> irv_node.cpp:176-221; every ballot has count 1 */
void orc_update(orc_tree *t, const uint8_t *prefs, const uint64_t *offsets, uint64_t n_ballots);
uint64_t orc_n_observed(orc_tree *t);

/* Node dump, depth-first in slot order. Per node, as uint32 words:
 *   depth, prefix[depth], K, slot_candidate[K], count[K+1], has_child[K]
 * Returns the number of words needed; fills buf when cap is large enough. */
uint64_t orc_dump_nodes(orc_tree *t, uint32_t *buf, uint64_t cap);
/* The observed multiset (dirichlet_tree.h:40, std::map in lexicographic order). */
orc_ballots *orc_observed(orc_tree *t);

/* dirichlet_tree.h:216-235 with the per-thread engine of R_tree.cpp:215-216:
 * std::mt19937 e(engine_seed); e.discard(624*100). */
orc_ballots *orc_posterior_set(orc_tree *t, uint32_t n_ballots, int replace, uint32_t engine_seed);
/* R_tree.cpp:155-175 without the count expansion: setSeed(seed); tree->sample(n). */
orc_ballots *orc_sample_predictive(orc_tree *t, uint32_t n_ballots, const char *seed, size_t seed_len);

uint64_t orc_ballots_size(const orc_ballots *b);
uint64_t orc_ballots_total_prefs(const orc_ballots *b);
void orc_ballots_export(const orc_ballots *b, uint8_t *prefs, uint64_t *offsets, uint32_t *counts);
void orc_ballots_free(orc_ballots *b);

/* R_social_choice.cpp:72-79 -> irv_ballot.cpp:42-138. Engine: seed_seq(seed bytes),
 * discard(624*100). order_out[n_candidates] = elimination order, winner last. */
void orc_irv(const uint8_t *prefs, const uint64_t *offsets, const uint32_t *counts, uint64_t n,
             uint32_t n_candidates, const char *seed, size_t seed_len, uint32_t *order_out);

/* R_tree.cpp:177-257 without Rcpp: seeds, static partition, std::thread pool,
 * aggregation. out[n_candidates] = win frequency; *pool_seconds = steady_clock
 * time around the pool only. Returns 0, or 1 when n_ballots < n_observed. */
int orc_sample_posterior(orc_tree *t, uint32_t n_elections, uint32_t n_ballots, uint32_t n_winners,
                         int replace, uint32_t n_threads, const char *seed, size_t seed_len,
                         double *out, double *pool_seconds);

#ifdef __cplusplus
}
#endif
#endif
