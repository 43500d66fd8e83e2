// Stand-in for JDFTx core/Random.h — see core/vector3.h in this directory for the why.
//
// Two generators behind the two symbols the reference calls (sarw.cpp:149-150,161,177,337-339):
//  * MODE_MT    (default): what upstream JDFTx does [memory, SURVEY §8c]: one global std::mt19937_64,
//               seed(int) re-initialises it, uniform(a,b) = a + (b-a)*U[0,1). Used for CPU TIMING so that
//               the reference pays its real per-trial reseed cost (73 % of its profile).
//  * MODE_KEYED: counter-based Philox4x32-10 keyed by (seed; chain, round, trial, draw), the same stream
//               the CUDA path and the C oracle use, so the unmodified reference walk can be compared
//               decision-for-decision. The (chain, round, trial) context is inferred by the harness from
//               the reference's own call pattern and public state (see ref_harness.cpp).
#ifndef STANDIN_CORE_RANDOM_H
#define STANDIN_CORE_RANDOM_H
namespace Random
{	void seed(int s);
	double uniform(double start = 0.0, double end = 1.0);
}
#endif
