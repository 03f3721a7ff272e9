/* oracle/philox_ref.h -- TEST INFRASTRUCTURE ONLY (checker; never shipped, never timed as product).
 *
 * Independent plain-C statement of Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy
 * as 1, 2, 3", SC'11; Random123 v1.14 philox.h constants).  The reference repo has no counter-based
 * generator (MCTSPlayer.cpp:472-476 uses std::default_random_engine, implementation-defined), so the
 * choice stream is NEW behaviour defined by SURVEY.md 8d:
 *     key     = (seed_lo, seed_hi)
 *     counter = (playout_id_lo, playout_id_hi, ply >> 2, 0)
 *     word    = ply & 3 ;  idx = (uint64(r) * n) >> 32
 * Pinned by the Random123 known-answer vectors in tests/test_philox.py.
 */
#ifndef ORACLE_PHILOX_REF_H
#define ORACLE_PHILOX_REF_H
#include <stdint.h>

static inline void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; ++r) {
		const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		const uint32_t n1 = (uint32_t)p1;
		const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		const uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* The 32-bit choice word for (key, playout_id, ply). */
static inline uint32_t oracle_choice_word(uint64_t key, uint64_t playout_id, uint32_t ply)
{
	uint32_t ctr[4] = { (uint32_t)playout_id, (uint32_t)(playout_id >> 32), ply >> 2, 0u };
	uint32_t k[2] = { (uint32_t)key, (uint32_t)(key >> 32) };
	uint32_t out[4];
	oracle_philox4x32_10(ctr, k, out);
	return out[ply & 3u];
}
/* idx in [0,n) -- multiply-high, no modulo bias argument needed for parity (both sides use it). */
static inline uint32_t oracle_choice_index(uint32_t r, uint32_t n)
{
	return (uint32_t)(((uint64_t)r * (uint64_t)n) >> 32);
}
#endif
