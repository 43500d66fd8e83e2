/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) —
 * plain C restatement of the published algorithm, for the ORACLE side only (test infrastructure).
 * The CUDA path has its own independent implementation in polymer-sarw_b200/csrc; the two are tied
 * together by the Random123 known-answer vectors in tests/test_oracle_philox.py and by parity tests.
 *
 * Stream layout shared by oracle, stand-in RNG and CUDA path (DESIGN.md "RNG"):
 *   key     = (seed & 0xffffffff, seed >> 32)
 *   counter = (chain, round_or_trial, block, kind)      kind 0 = growth, 1 = seeding
 *   growth : trial t (1-based) of chain c in growth round r uses word (t-1)&3 of block (t-1)>>2,
 *            counter = (c, r, (t-1)>>2, 0);           U = word * 2^-32
 *   seeding: trial t of chain c uses counter = (c, t, b, 1), b = 0 -> U0..U3, b = 1 -> U4,U5
 *            (U0..U2 = pos0 draws of sarw.cpp:337-339, U3..U5 = randomUnitStep draws of sarw.cpp:150, in the
 *            order the compiled reference makes them: g++ evaluates constructor arguments right to left,
 *            so U0 -> pos0.z, U1 -> pos0.y, U2 -> pos0.x, U3 -> step.z, U4 -> step.y, U5 -> step.x)
 */
#ifndef SARW_ORACLE_PHILOX_H
#define SARW_ORACLE_PHILOX_H
#include <stdint.h>

static inline void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int round = 0; round < 10; ++round) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define ORACLE_KIND_GROW 0u
#define ORACLE_KIND_SEED 1u

/* U[0,1) of growth trial t (1-based) */
static inline double oracle_u_grow(uint64_t seed, uint32_t chain, uint32_t round, uint32_t trial)
{
	uint32_t ctr[4] = { chain, round, (trial - 1u) >> 2, ORACLE_KIND_GROW };
	uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
	uint32_t w[4];
	oracle_philox4x32_10(ctr, key, w);
	return (double)w[(trial - 1u) & 3u] * (1.0 / 4294967296.0);
}

/* six U[0,1) draws of seeding trial t (1-based) */
static inline void oracle_u_seed(uint64_t seed, uint32_t chain, uint32_t trial, double u[6])
{
	uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
	uint32_t ctr[4] = { chain, trial, 0u, ORACLE_KIND_SEED };
	uint32_t w[4];
	oracle_philox4x32_10(ctr, key, w);
	for (int i = 0; i < 4; ++i) u[i] = (double)w[i] * (1.0 / 4294967296.0);
	ctr[2] = 1u;
	oracle_philox4x32_10(ctr, key, w);
	u[4] = (double)w[0] * (1.0 / 4294967296.0);
	u[5] = (double)w[1] * (1.0 / 4294967296.0);
}
#endif
