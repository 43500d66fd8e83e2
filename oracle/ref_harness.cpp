// Harness around the UNMODIFIED reference sources (/root/reference/{sarw.cpp,InputMap.cpp,sarw.h,
// PeriodicLookup.h,InputMap.h}), built by oracle/Makefile into oracle/_ref/libsarw_ref.so.
// TEST INFRASTRUCTURE ONLY: used by tests/ as the truth for parity, and by bench.py's reference arm /
// cpu_baseline as the CPU timing baseline. Nothing in the product path links or loads it.
//
// It drives the reference through its own public surface only:
//   * PeriodicLookup{ctor, addPoint, find}                      PeriodicLookup.h:26,37,51
//   * SARW{ctor, public fields, initiateChain, propagateChain, terminateChain, report,
//          calcAnglesDihedrals, exportLAMMPS/XYZ/XSF}           sarw.h:59-136
//   * the reference's main(), renamed by -Dmain=sarw_reference_main at compile time (sarw.cpp:13)
// and mirrors the call order of main (sarw.cpp:59-86) when it needs a seeded, inspectable run.
#include <core/Random.h>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <unistd.h>
#include "sarw.h"
#include "standin_hooks.h"
#include "../philox.h"

int sarw_reference_main(int argc, char** argv); // = main of sarw.cpp:13 (renamed on the command line)

// ------------------------------------------------------------------ keyed RNG context inference
namespace {
struct KeyedCtx
{	const SARW* s = 0;
	uint64_t seed = 0;
	int phase = 0;          // 0 = idle, 1 = seeding (initiateChain), 2 = growth (propagateChain)
	uint32_t round = 0;     // growth round (0-based), set by the harness before each propagateChain()
	int chain = 0, trial = 0, draws = 0, lastCount = 0;
	double useed[6];
	unsigned long long growTrials = 0, seedTrials = 0;
} K;

// reference call pattern, growth: one Random::seed() then one Random::uniform() per trial (sarw.cpp:161,177);
// trials of one chain run until acceptance (actualCount grows, sarw.cpp:267) or maxTrials (sarw.cpp:259).
void keyedSeed(int)
{	if(K.phase != 2) return; // seeding: the reseed between pos0 and step draws carries no information
	int c = K.s->actualCount();
	if(K.trial > 0 && (c != K.lastCount || K.trial == K.s->maxTrials)) { K.chain++; K.trial = 0; K.lastCount = c; }
	K.trial++;
	K.growTrials++;
}
double keyedUniform01()
{	if(K.phase == 2)
		return oracle_u_grow(K.seed, (uint32_t)K.chain, K.round, (uint32_t)K.trial);
	// seeding: 3 draws (sarw.cpp:337-339) + reseed + 3 draws (sarw.cpp:150) per trial; every accepted
	// chain adds exactly two atoms (sarw.cpp:303-304), so chain = actualCount()/2
	int c = K.s->actualCount() / 2;
	if(c != K.chain || K.trial == 0) { K.chain = c; K.trial = 0; K.draws = 6; }
	if(K.draws == 6)
	{	K.trial++; K.draws = 0; K.seedTrials++;
		oracle_u_seed(K.seed, (uint32_t)K.chain, (uint32_t)K.trial, K.useed);
	}
	return K.useed[K.draws++];
}
} // namespace

// ------------------------------------------------------------------ C surface
extern "C" {

struct ref_params
{	int32_t nChains;
	double box[3];
	double minDist;
	double targetMassDensity;
	int32_t nGraftChains;
	int32_t maxTrials;
	char graftLoc[16];
	char growthOrder[16];
	char boundary[8];
	char polymer[64];
	int32_t rngMode;        // 0 = mt19937_64 with per-trial reseed (upstream behaviour), 1 = keyed Philox
	uint64_t seed;
	int32_t maxRounds;      // < 0: run to completion like main(); >= 0: stop after that many propagateChain() calls
};

struct ref_walk
{	SARW* s;
	int initiated;          // return value of initiateChain()
	int rounds;             // propagateChain() calls made
	int lastProgressed;     // return value of the last propagateChain()
	double seconds;         // steady_clock around initiateChain + growth loop (sarw.cpp:66-78)
	unsigned long long trials;   // growth trials = Random::seed calls during growth
	unsigned long long seedTrials;
	std::vector<int> countAfterRound;
};

// ---- PeriodicLookup 1:1
void* ref_lookup_create(double lx, double ly, double lz, double thresh)
{	return new PeriodicLookup(vector3<>(lx, ly, lz), thresh);
}
void ref_lookup_destroy(void* h) { delete (PeriodicLookup*)h; }
void ref_lookup_add(void* h, const double* xyz, int64_t n)
{	PeriodicLookup* p = (PeriodicLookup*)h;
	for(int64_t i = 0; i < n; i++) p->addPoint(vector3<>(xyz[3*i], xyz[3*i+1], xyz[3*i+2]));
}
void ref_lookup_find(void* h, const double* xyz, const int64_t* ignore, int64_t n, uint8_t* hit)
{	const PeriodicLookup* p = (const PeriodicLookup*)h;
	for(int64_t i = 0; i < n; i++)
		hit[i] = p->find(vector3<>(xyz[3*i], xyz[3*i+1], xyz[3*i+2]), ignore ? (int)ignore[i] : -1) ? 1 : 0;
}

// ---- whole walk through the SARW class, call order of main (sarw.cpp:59-79)
ref_walk* ref_walk_run(const ref_params* p, const double* grafts, int64_t nGrafts, const double* obstacles, int64_t nObst)
{	ref_walk* w = new ref_walk();
	SARW* s = new SARW(p->nChains, vector3<>(p->box[0], p->box[1], p->box[2]), p->minDist);
	w->s = s;
	s->targetMassDensity = p->targetMassDensity;
	s->nGraftChains = p->nGraftChains;
	s->graftLoc = p->graftLoc;
	s->growthOrder = p->growthOrder;
	s->boundary = p->boundary;
	s->maxTrials = p->maxTrials;
	s->polymer = p->polymer;
	s->setLogFlags("no", "no");
	s->obstacle = "none";
	s->nObstacles = 0; // NB: left uninitialised by the reference when no obstacle file is read (sarw.h:75,88-94)
	// what readGrafts (sarw.cpp:110,116-119) and addObstacle (sarw.cpp:135-142) do with file contents:
	for(int64_t i = 0; i < nGrafts && i < p->nGraftChains; i++)
		s->listGrafts.push_back(vector3<>(grafts[3*i], grafts[3*i+1], grafts[3*i+2]));
	for(int64_t i = 0; i < nObst; i++) s->plook.addPoint(vector3<>(obstacles[3*i], obstacles[3*i+1], obstacles[3*i+2]));
	if(nObst > 0) s->nObstacles = (int)nObst;

	standin_rng_reset_counters();
	if(p->rngMode == 1)
	{	K = KeyedCtx(); K.s = s; K.seed = p->seed;
		standin_rng_set_hooks(keyedSeed, keyedUniform01);
	}
	else
	{	standin_rng_set_hooks(0, 0);
		srand((unsigned)p->seed); // main uses srand(time(NULL)) (sarw.cpp:16); fixed here for repeatability
		Random::seed((int)p->seed);
		standin_rng_reset_counters();
	}

	auto t0 = std::chrono::steady_clock::now();
	K.phase = 1; K.chain = -1; K.trial = 0; K.draws = 6;
	w->initiated = s->initiateChain() ? 1 : 0;
	unsigned long long seedCallsAfterInit = standin_rng_seed_calls();
	w->lastProgressed = 1;
	if(w->initiated)
	{	while(s->actualCount() < s->targetCount())
		{	if(p->maxRounds >= 0 && w->rounds >= p->maxRounds) break;
			// chain range of this round (sarw.cpp:230-249), needed only to tell the keyed RNG where the loop starts
			int startChain = 0;
			if(s->growthOrder == "grafted" && s->nGraftChains > 0)
			{	int nGraftAtoms = (s->nGraftChains * s->targetCount() / s->nChains) + (s->nChains - s->nGraftChains) * 2;
				if(!(s->actualCount() < nGraftAtoms)) startChain = s->nGraftChains;
			}
			K.phase = 2; K.round = (uint32_t)w->rounds; K.chain = startChain; K.trial = 0; K.lastCount = s->actualCount();
			bool ok = s->propagateChain();
			w->rounds++;
			w->countAfterRound.push_back(s->actualCount());
			w->lastProgressed = ok ? 1 : 0;
			if(!ok) break;
		}
	}
	K.phase = 0;
	auto t1 = std::chrono::steady_clock::now();
	w->seconds = std::chrono::duration<double>(t1 - t0).count();
	w->trials = standin_rng_seed_calls() - seedCallsAfterInit;
	w->seedTrials = (p->rngMode == 1) ? K.seedTrials : seedCallsAfterInit;
	if(s->actualCount() > 0) s->terminateChain(); // sarw.cpp:79, unconditional in main(); guarded only against the empty-vector write it would do when chain 0 failed
	standin_rng_set_hooks(0, 0);
	return w;
}

void ref_walk_counts(const ref_walk* w, int64_t* nAtoms, int64_t* nBonds, int64_t* target, int32_t* initiated, int32_t* rounds,
	int32_t* lastProgressed, double* seconds, uint64_t* trials, uint64_t* seedTrials)
{	*nAtoms = w->s->actualCount(); *nBonds = w->s->nBonds(); *target = w->s->targetCount();
	*initiated = w->initiated; *rounds = w->rounds; *lastProgressed = w->lastProgressed; *seconds = w->seconds;
	*trials = w->trials; *seedTrials = w->seedTrials;
}

void ref_walk_get(const ref_walk* w, double* xyz, int32_t* chainID, int32_t* type, int64_t* bonds, int32_t* chainLengths, int64_t* countAfterRound)
{	const SARW* s = w->s;
	for(int i = 0; i < s->actualCount(); i++)
	{	const unitedAtom& a = s->polymerChains[i];
		if(xyz) { xyz[3*i] = a.pos[0]; xyz[3*i+1] = a.pos[1]; xyz[3*i+2] = a.pos[2]; }
		if(chainID) chainID[i] = a.chainID;
		if(type) type[i] = a.type;
	}
	if(bonds) for(int i = 0; i < s->nBonds(); i++) { bonds[2*i] = s->listBonds[i].species[0]; bonds[2*i+1] = s->listBonds[i].species[1]; }
	if(chainLengths) for(int i = 0; i < s->nChains; i++) chainLengths[i] = s->chainLengths[i];
	if(countAfterRound) for(size_t i = 0; i < w->countAfterRound.size(); i++) countAfterRound[i] = w->countAfterRound[i];
}

// report + topology + the three exporters, exactly as main calls them (sarw.cpp:82-86), inside directory `dir`
int ref_walk_export(ref_walk* w, const char* dir, const char* logFile)
{	char cwd[4096];
	if(!getcwd(cwd, sizeof cwd)) return -1;
	if(chdir(dir) != 0) return -2;
	FILE* saved = globalLog;
	FILE* lf = logFile ? fopen(logFile, "w") : 0;
	if(lf) globalLog = lf;
	w->s->report();
	w->s->calcAnglesDihedrals();
	w->s->exportLAMMPS();
	w->s->exportXYZ();
	w->s->exportXSF();
	if(lf) { fclose(lf); globalLog = saved; }
	int rc = chdir(cwd);
	return rc;
}

// topology only (sarw.cpp:397-422), for parity of the O(N) host builder
void ref_walk_topology(ref_walk* w, int64_t* nAngles, int64_t* nDihedrals)
{	if(w->s->nAngles() == 0 && w->s->nDihedrals() == 0) w->s->calcAnglesDihedrals();
	*nAngles = w->s->nAngles(); *nDihedrals = w->s->nDihedrals();
}
void ref_walk_get_topology(const ref_walk* w, int64_t* angles, int64_t* dihedrals)
{	for(int i = 0; i < w->s->nAngles(); i++) for(int j = 0; j < 3; j++) angles[3*i+j] = w->s->listAngles[i].species[j];
	for(int i = 0; i < w->s->nDihedrals(); i++) for(int j = 0; j < 4; j++) dihedrals[4*i+j] = w->s->listDihedrals[i].species[j];
}

void ref_walk_free(ref_walk* w) { delete w->s; delete w; }

// the reference's own main(), e.g. {"sarw","-i","pe.in","-o","log"}; RNG = upstream-like mt19937_64
int ref_main(int argc, char** argv)
{	standin_rng_set_hooks(0, 0);
	FILE* saved = globalLog;
	int rc = sarw_reference_main(argc, argv);
	if(globalLog != saved && globalLog != stdout) { fclose(globalLog); }
	globalLog = saved;
	return rc;
}

} // extern "C"
