// Stand-in implementations for the JDFTx symbols the reference links against.
// TEST INFRASTRUCTURE ONLY — lets the unmodified reference sources build into oracle/_ref.
#include <core/Util.h>
#include <core/Random.h>
#include <random>
#include <cstring>
#include "standin_hooks.h"

FILE* globalLog = stdout;
static StandinMPI standinMPI;
StandinMPI* mpiWorld = &standinMPI;

// ${VAR} / $VAR expansion in config values (InputMap.cpp:25,39)
void environmentSubstitute(string& s)
{	string out; size_t i = 0;
	while(i < s.size())
	{	if(s[i] == '$' && i + 1 < s.size())
		{	size_t b = i + 1, e; bool braced = (s[b] == '{');
			if(braced) { b++; e = s.find('}', b); if(e == string::npos) { out += s[i++]; continue; } }
			else { e = b; while(e < s.size() && (isalnum((unsigned char)s[e]) || s[e] == '_')) e++; }
			string name = s.substr(b, e - b);
			const char* val = name.empty() ? 0 : getenv(name.c_str());
			if(name.empty()) { out += s[i++]; continue; }
			if(val) out += val;
			i = braced ? e + 1 : e;
		}
		else out += s[i++];
	}
	s = out;
}

void initSystemCmdline(int argc, char** argv, InitParams& ip)
{	for(int i = 1; i < argc; i++)
	{	if(!strcmp(argv[i], "-i") && i + 1 < argc) ip.inputFilename = argv[++i];
		else if(!strcmp(argv[i], "-o") && i + 1 < argc) ip.logFilename = argv[++i];
		else if(!strcmp(argv[i], "-n")) ip.dryRun = true;
	}
	if(!ip.logFilename.empty())
	{	FILE* fp = fopen(ip.logFilename.c_str(), "w");
		if(fp) globalLog = fp;
	}
	logPrintf("\n*************** %s %s (git hash %s) ***************\n", ip.packageName ? ip.packageName : "",
		ip.versionString ? ip.versionString : "", ip.versionHash ? ip.versionHash : "");
	if(ip.description) logPrintf("%s\n", ip.description);
	if(ip.inputFilename.empty()) die("No input file given (use -i <file>).\n");
}

// ---- RNG ----
static std::mt19937_64 mtGen;
static std::uniform_real_distribution<double> mtDist(0.0, 1.0);
static void mtSeed(int s) { mtGen.seed(s); }
static double mtUniform01() { return mtDist(mtGen); }

static void (*seedHook)(int) = mtSeed;
static double (*uniformHook)() = mtUniform01;
static unsigned long long nSeedCalls = 0, nUniformCalls = 0;

extern "C" void standin_rng_set_hooks(void (*seedFn)(int), double (*uniform01Fn)())
{	seedHook = seedFn ? seedFn : mtSeed;
	uniformHook = uniform01Fn ? uniform01Fn : mtUniform01;
}
extern "C" unsigned long long standin_rng_seed_calls() { return nSeedCalls; }
extern "C" unsigned long long standin_rng_uniform_calls() { return nUniformCalls; }
extern "C" void standin_rng_reset_counters() { nSeedCalls = 0; nUniformCalls = 0; }

namespace Random
{	void seed(int s) { nSeedCalls++; seedHook(s); }
	double uniform(double start, double end) { nUniformCalls++; return start + (end - start) * uniformHook(); }
}
