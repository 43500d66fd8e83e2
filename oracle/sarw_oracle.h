/* sarw_oracle — plain-C CPU restatement of the reference's chain-growth hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this. The product (libsarw_b200.so) never links, loads or falls back to it.
 *
 * Every function cites the reference lines (under /root/reference) it restates. The restatement is pinned by
 * tests/test_oracle_vs_ref.py: with trig = ORACLE_TRIG_LIBM it reproduces, bit for bit, the walk of the
 * UNMODIFIED reference sources (oracle/_ref/libsarw_ref.so) driven by the same keyed Philox stream, and its
 * lookup reproduces PeriodicLookup::find on adversarial candidates; golden copies of those reference outputs
 * are committed under tests/golden/. With trig = ORACLE_TRIG_PORTABLE (a fixed sequence of IEEE double
 * operations for sin/cos, no libm) it is the bit-exact target of the CUDA path.
 * The reference has no tests, golden vectors or fixtures of its own.
 */
#ifndef SARW_ORACLE_H
#define SARW_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_TRIG_LIBM = 0, ORACLE_TRIG_PORTABLE = 1 };
enum { ORACLE_GRAFT_FILE = 0, ORACLE_GRAFT_Z = 1, ORACLE_GRAFT_ZCENTER = 2, ORACLE_GRAFT_OTHER = 3 };
enum { ORACLE_ORDER_ROUNDROBIN = 0, ORACLE_ORDER_GRAFTED = 1 };
enum { ORACLE_BOUNDARY_PPP = 0, ORACLE_BOUNDARY_PPF = 1, ORACLE_BOUNDARY_PFP = 2, ORACLE_BOUNDARY_FPP = 3 };

typedef struct oracle_params {
	int32_t nChains;
	double box[3];
	double minDist;
	double targetMassDensity;
	int32_t nGraftChains;
	int32_t maxTrials;
	int32_t graftLoc;
	int32_t growthOrder;
	int32_t boundary;
	double growthBias[3];   /* extension (README.md:33 only; no reference code): reject steps with bias.step < 0 */
	int32_t trig;
	uint64_t seed;
	int32_t maxRounds;      /* < 0: to completion */
} oracle_params;

typedef struct oracle_lookup oracle_lookup;
typedef struct oracle_walk oracle_walk;

/* PeriodicLookup.h:26-34, 37-47, 51-75 */
oracle_lookup* oracle_lookup_create(double lx, double ly, double lz, double thresh);
void oracle_lookup_destroy(oracle_lookup* lk);
void oracle_lookup_add(oracle_lookup* lk, const double* xyz, int64_t n);
void oracle_lookup_find(const oracle_lookup* lk, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit);
/* all-pairs version of the same per-pair arithmetic (PeriodicLookup.h:53-55,67-71), no cells */
void oracle_lookup_find_bruteforce(const oracle_lookup* lk, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit);

/* sarw.h:103-104 (64-bit volume; same truncation) */
int64_t oracle_target_count(const double box[3], double targetMassDensity);

/* whole walk: sarw.cpp:66-79 */
oracle_walk* oracle_walk_run(const oracle_params* p, const double* grafts, int64_t nGrafts, const double* obstacles, int64_t nObst);
void oracle_walk_counts(const oracle_walk* w, int64_t* nAtoms, int64_t* nBonds, int64_t* target, int32_t* initiated, int32_t* rounds,
	int32_t* lastProgressed, double* seconds, uint64_t* trials, uint64_t* seedTrials);
void oracle_walk_get(const oracle_walk* w, double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds, int32_t* chainLengths,
	int64_t* countAfterRound, int32_t* acceptedTrial);
void oracle_walk_free(oracle_walk* w);

/* ---- analysis: scripts/rdf.py restated for the counts it accumulates ----
 * rdf.py:13-20 rdfPair: dist = (a1 - a2)*(1/L); dist -= floor(0.5 + dist); dist *= L; r = sqrt(sum(dist^2)); numpy.histogram
 * over edges arange(rMin, ., dr) (bins half open, the last one closed). rdf.py:39-57 rdfChains: intra = all ordered pairs of a
 * chain with itself (self pairs included); inter = 2 x pairs with the first atom taken from the chain with the HIGHER id.
 * edges[nBins+1] as numpy.arange fills them: start + k*((start+dr)-start). chain ids as in polymer.dat (any integers). */
void oracle_rdf_edges(double rMin, double dr, int32_t nBins, double* edges);
void oracle_rdf(const double box[3], const double* xyz, const int32_t* chain, int64_t n, double rMin, double dr, int32_t nBins,
	uint64_t* intra, uint64_t* inter);
/* per-chain end-to-end distance^2 and radius of gyration^2 from unwrapped coordinates in acceptance order (chain ids 1..nChains);
 * sums run over the beads of a chain in chain order, (x+y)+z per bead */
void oracle_chain_shapes(const double* xyz, const int32_t* chain, int64_t n, int32_t nChains, double* ree2, double* rg2);

/* building blocks exposed for unit tests */
void oracle_sincos_portable(double x, double* s, double* c);
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_cone_pos(const double last[3], const double pen[3], double u01, int trig, double out[3]);

#ifdef __cplusplus
}
#endif
#endif
